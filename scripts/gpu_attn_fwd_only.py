"""attn_fwd alone (ncu target): bench shape B 8, H 32, L 835 (or argv[1]), hd 128, causal."""
import math, sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 835
B, H, hd = 8, 32, 128
W = H * hd
qkv = torch.randn(B * L, 3 * W, device=dev).bfloat16()
o = torch.empty(B * L, W, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device=dev)
q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
for _ in range(5):
    ops.attn_fwd(q, k, v, B, H, L, L, hd, 1 / math.sqrt(hd), True, out=o, lse=lse)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    ops.attn_fwd(q, k, v, B, H, L, L, hd, 1 / math.sqrt(hd), True, out=o, lse=lse)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"L {L}: {us:.1f} us, {4 * L * L * hd * H * B / 2 / us / 1e6:.0f} TFLOP/s")
