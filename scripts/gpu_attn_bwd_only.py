"""attn_bwd alone (ncu target): bench shape B 8, H 32, L 835 (or argv[1]), hd 128, causal."""
import math, sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 835
B, H, hd = 8, 32, 128
W = H * hd
qkv = torch.randn(B * L, 3 * W, device=dev).bfloat16()
dout = torch.randn(B * L, W, device=dev).bfloat16()
o = torch.empty(B * L, W, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device=dev)
sc = 1 / math.sqrt(hd)
q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
ops.attn_fwd(q, k, v, B, H, L, L, hd, sc, True, out=o, lse=lse)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, L, device=dev); dq_acc = torch.empty(B * L, W, device=dev)
bwd = lambda: ops.attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, dqkv[:, :W], dqkv[:, W:2 * W], dqkv[:, 2 * W:], B, H, L, hd, sc, True)
for _ in range(3): bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): bwd()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(f"L {L}: attn_bwd (delta + memset + main + convert) {us:.1f} us, {2.5 * 4 * L * L * hd * H * B / 2 / us / 1e6:.0f} TFLOP/s")
