"""ncu aid: one eager decode step of the 7B model at batch B between cudaProfilerStart/Stop."""
import sys, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model = MetaModel("llama_b200", {**bench.MODELS["7b"], "image_size": 336}, None, with_visual=True, max_seq_len=2048, max_batch_size=64)
model.llma.reset_parameters(seed=2)
model.llma.engine.use_decode_graph = False
model.llma.engine.decode_fused_max_b = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator().manual_seed(7)
tk = torch.randint(3, 32000, (B, 256), generator=g).cuda(); tk[:, 0] = 1
img = torch.randn(B, 3, 336, 336, generator=g).cuda()
lg = model.llma.forward_inference(tk, 0, img)
nxt = lg.argmax(-1, keepdim=True)
for i in range(3):
    lg = model.llma.forward_inference(nxt, 256 + i); nxt = lg.argmax(-1, keepdim=True)
torch.cuda.synchronize(); torch.cuda.profiler.start()
lg = model.llma.forward_inference(nxt, 259)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
