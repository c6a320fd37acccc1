"""BASELINE config 2: batch-1 inference, ViT-L/14 + Llama-2-7B bf16, 336x336 image, 256-token prompt,
128 generated tokens, one B200.  Reports prefill latency, per-token decode latency against the HBM
roofline (13.21 GB of weights + KV per token), and end-to-end generate() time."""
import json, os, sys, time, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200 import ops
from a3vlm_b200.model import MetaModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = bench.MODELS["7b"]
dev = torch.device("cuda", 0)
model = MetaModel("llama_b200", {**cfg, "image_size": 336}, None, with_visual=True, max_seq_len=2048, max_batch_size=64)
model.llma.reset_parameters(seed=2)
eng = model.llma.engine
g = torch.Generator().manual_seed(7)
prompts = torch.randint(3, 32000, (B, 256), generator=g); prompts[:, 0] = 1
img = torch.randn(B, 3, 336, 336, generator=g).to(dev)
plist = prompts.tolist()
def ev():
    return torch.cuda.Event(enable_timing=True)
res = {"config": f"batch {B}, 336px + 256-token prompt, 128 generated tokens, llama-7b + ViT-L/14, 1xB200"}
for graph in (True, False):
    eng.use_decode_graph = graph
    eng.destroy_kv_cache()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, toks, _ = model.generate_tokens(plist, img, max_gen_len=128, eos_id=-1)
        torch.cuda.synchronize(); t_gen = time.perf_counter() - t0
    # split: prefill, then decode steps
    e0, e1, e2 = ev(), ev(), ev()
    tk = prompts.to(dev)
    e0.record()
    lg = model.llma.forward_inference(tk, 0, img)
    e1.record()
    nxt = lg.argmax(-1, keepdim=True)
    for i in range(127):
        lg = model.llma.forward_inference(nxt, 256 + i)
        nxt = lg.argmax(-1, keepdim=True)
    e2.record(); torch.cuda.synchronize()
    pre, dec = e0.elapsed_time(e1), e1.elapsed_time(e2) / 127
    ctx = 579 + 256 + 64
    bytes_tok = 2 * (eng.w.numel - 32000 * 4096) + B * ctx * 2 * 32 * 4096 * 2  # weights once (embedding rows excluded) + KV
    res["graph" if graph else "eager"] = {"generate_s": t_gen, "prefill_ms": pre, "decode_ms_per_token": dec,
        "decode_tokens_per_s": B * 1e3 / dec, "hbm_GBs": bytes_tok / dec / 1e6, "hbm_frac_of_measured_peak": bytes_tok / dec / 1e6 / bench.peaks()["hbm"],
        "generated_tokens": len(out[0])}
print(json.dumps(res))
