"""Isolated timing of the attention kernels at the bench shape (B 8, H 32, L 835, hd 128) and ViT shape."""
import math, sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B, H, L, hd = 8, 32, 835, 128
W = H * hd
qkv = (torch.randn(B * L, 3 * W, device=dev)).bfloat16()
dout = torch.randn(B * L, W, device=dev).bfloat16()
o = torch.empty(B * L, W, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device=dev)
sc = 1 / math.sqrt(hd)
q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
fwd = lambda: ops.attn_fwd(q, k, v, B, H, L, L, hd, sc, True, out=o, lse=lse)
us = t(fwd)
fl = 4 * L * L * hd * H * B / 2
print(f"attn_fwd causal hd128: {us:.1f} us  {fl/us/1e6:.0f} TFLOP/s")
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, L, device=dev); dq_acc = torch.empty(B * L, W, device=dev)
bwd = lambda: ops.attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, dqkv[:, :W], dqkv[:, W:2 * W], dqkv[:, 2 * W:], B, H, L, hd, sc, True)
us = t(bwd)
print(f"attn_bwd causal hd128 (delta+memset+main+convert): {us:.1f} us  {2.5*fl/us/1e6:.0f} TFLOP/s")
Bv, Hv, N, hdv = 8, 16, 577, 64
Wv = Hv * hdv
qkv2 = torch.randn(Bv * N, 3 * Wv, device=dev).bfloat16(); o2 = torch.empty(Bv * N, Wv, device=dev, dtype=torch.bfloat16)
f2 = lambda: ops.attn_fwd(qkv2[:, :Wv], qkv2[:, Wv:2 * Wv], qkv2[:, 2 * Wv:], Bv, Hv, N, N, hdv, 0.125, False, out=o2)
us = t(f2)
print(f"attn_fwd vit hd64: {us:.1f} us  {4*N*N*hdv*Hv*Bv/us/1e6:.0f} TFLOP/s")
