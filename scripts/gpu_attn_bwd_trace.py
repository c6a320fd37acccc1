"""Debug aid: per-role clock64 stamps of one attn_bwd launch (A3_BW_TRACE_PTR)."""
import math, os, sys, torch
sys.path.insert(0, ".")
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 835
B, H, hd = 8, 32, 128
nk = (L + 127) // 128
trace = torch.zeros(B * H * nk * 384, dtype=torch.int64, device=dev)
os.environ["A3_BW_TRACE_PTR"] = str(trace.data_ptr())
from a3vlm_b200 import ops
W = H * hd
qkv = torch.randn(B * L, 3 * W, device=dev).bfloat16(); dout = torch.randn(B * L, W, device=dev).bfloat16()
o = torch.empty(B * L, W, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, L, device=dev)
sc = 1 / math.sqrt(hd)
q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
ops.attn_fwd(q, k, v, B, H, L, L, hd, sc, True, out=o, lse=lse)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, L, device=dev); dq_acc = torch.empty(B * L, W, device=dev)
for _ in range(2):
    trace.zero_()
    ops.attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, dqkv[:, :W], dqkv[:, W:2 * W], dqkv[:, 2 * W:], B, H, L, hd, sc, True)
torch.cuda.synchronize()
tr = trace.cpu().view(B, H, nk, 3, 128)
for (bb, hh, kt) in ((0, 0, 0), (B - 1, H - 1, 0), (3, 5, 2)):
    t = tr[bb, hh, kt]
    t0 = min(int(x) for x in t.flatten() if x > 0)
    print(f"--- b {bb} h {hh} key tile {kt}")
    for r, name in enumerate(("worker", "issuer", "producer")):
        print(f"{name:9s}", " ".join(str(int(x) - t0) for x in t[r] if x > 0))
