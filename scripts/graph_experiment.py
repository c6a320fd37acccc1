"""Experiment: how much of the step is launch gaps?  Times fwd+bwd eagerly and as one CUDA graph."""
import sys, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
cfg = bench.MODELS["7b"]
dev = torch.device("cuda", 0)
model = MetaModel("llama_b200", {**cfg, "image_size": 336}, None, with_visual=True, max_seq_len=2048)
model.llma.reset_parameters(seed=2)
eng = model.llma.engine
B, T, S = 8, 256, 336
g = torch.Generator().manual_seed(100)
tok = torch.randint(3, 32000, (B, T), generator=g); tok[:, 0] = 1
lab = tok.clone(); lab[:, :128] = 0
img = torch.randn(B, 3, S, S, generator=g)
tok, lab, img = tok.to(dev), lab.to(dev), img.to(dev)
def fb():
    ws, _ = eng.forward_train(tok, img, lab)
    eng.backward(ws, tok, ws.logits)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager fwd+bwd ms", timeit(fb))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fb(); fb()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    fb()
print("graph fwd+bwd ms", timeit(gr.replay))
print("eager again   ms", timeit(fb))
