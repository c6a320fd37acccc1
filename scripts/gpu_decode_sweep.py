"""Decode latency A/B (BASELINE config 2 geometry: 336 px + 256-token prompt, 7B, CUDA-graph decode): ms/token at the
batch given on the command line under the current environment (A3_DEC_PREFETCH_MB, A3_LIB_PATH).

    python scripts/gpu_decode_sweep.py [batch] [tokens]"""
import json, os, sys, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = bench.MODELS["7b"]
dev = torch.device("cuda", 0)
model = MetaModel("llama_b200", {**cfg, "image_size": 336}, None, with_visual=True, max_seq_len=2048, max_batch_size=max(B, 1))
model.llma.reset_parameters(seed=2)
eng = model.llma.engine
g = torch.Generator().manual_seed(7)
prompts = torch.randint(3, 32000, (B, 256), generator=g); prompts[:, 0] = 1
img = torch.randn(B, 3, 336, 336, generator=g).to(dev)
tk = prompts.to(dev)
best = None
ids = None
for rep in range(3):
    eng.destroy_kv_cache() if rep == 0 else None
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    lg = model.llma.forward_inference(tk, 0, img)
    e1.record()
    nxt = lg.argmax(-1, keepdim=True)
    got = [nxt]
    for i in range(NT - 1):
        lg = model.llma.forward_inference(nxt, 256 + i)
        nxt = lg.argmax(-1, keepdim=True)
        got.append(nxt)
    e2.record(); torch.cuda.synchronize()
    dec = e1.elapsed_time(e2) / (NT - 1)
    best = dec if best is None else min(best, dec)
    cur = torch.cat(got, 1).cpu()
    assert ids is None or bool((cur == ids).all()), "decode not repeatable"
    ids = cur
ctx = 579 + 256 + NT // 2
bytes_tok = 2 * (eng.w.numel - 32000 * 4096) + B * ctx * 2 * 32 * 4096 * 2
print(json.dumps({"batch": B, "prefetch_mb": os.environ.get("A3_DEC_PREFETCH_MB", "default"), "lib": os.path.basename(os.environ.get("A3_LIB_PATH", "default")),
                  "decode_ms_per_token": round(best, 4), "prefill_ms": round(e0.elapsed_time(e1), 3), "hbm_frac": round(bytes_tok / best / 1e6 / bench.peaks()["hbm"], 4),
                  "ids_head": ids[0, :8].tolist()}))
