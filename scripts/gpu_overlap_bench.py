"""Pipelined optimiser step (Engine.overlap_optimizer) vs the plain one on the headline workload:
7B fine-tune step, batch 8, 336 px + 256 text tokens, one B200.  One model build, K timed steps per
variant (CUDA events, synchronised on both sides), losses printed so the variants can be checked
against each other (same seeds -> same loss trajectory)."""
import json, os, sys, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
BPS = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
cfg = bench.MODELS["7b"]
dev = torch.device("cuda", 0)
B, T, S = 8, 256, 336
model = MetaModel("llama_b200", {**cfg, "image_size": S}, None, with_visual=True, max_seq_len=2048)
eng = model.llma.engine
g = torch.Generator().manual_seed(100)
tok = torch.randint(3, cfg["vocab_size"], (B, T), generator=g); tok[:, 0] = 1
lab = tok.clone(); lab[:, : T // 2] = 0
img = torch.randn(B, 3, S, S, generator=g)
tok, lab, img = tok.to(dev), lab.to(dev), img.to(dev)

def step():
    ws, _ = eng.forward_train(tok, img, lab)
    eng.backward(ws, tok, ws.logits, accumulate=False, scale=1.0)
    eng.optimizer_step(2e-5, 0.0, 8.0)
    return ws.loss

res = {}
for name, overlap, bps in [("plain", False, 1)] + [(f"overlap_bps{b}", True, b) for b in BPS] + [("plain_again", False, 1)]:
    model.llma.reset_parameters(seed=2)
    eng.master = eng.m = eng.v = None
    eng.opt_step = 0
    eng.overlap_optimizer, eng.overlap_adamw_blocks_per_sm = overlap, bps
    losses = []
    for _ in range(3):
        losses.append(step())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        losses.append(step().clone())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res[name] = {"ms_per_step": ms, "img_per_s": B / ms * 1e3, "losses": [round(float(l), 5) for l in losses[3:]]}
    print(name, res[name], flush=True)
eng.overlap_optimizer = False
print(json.dumps(res))
