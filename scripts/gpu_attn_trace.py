"""Debug aid: per-role clock64 stamps of one attn_fwd launch (A3_FA_TRACE_PTR), printed as deltas for a few CTAs."""
import math, os, sys, torch
sys.path.insert(0, ".")
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 835
B, H, hd = 8, 32, 128
nsm = torch.cuda.get_device_properties(0).multi_processor_count
trace = torch.zeros(nsm * 768, dtype=torch.int64, device=dev)
os.environ["A3_FA_TRACE_PTR"] = str(trace.data_ptr())
from a3vlm_b200 import ops
W = H * hd
qkv = torch.randn(B * L, 3 * W, device=dev).bfloat16()
o = torch.empty(B * L, W, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device=dev)
q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
for _ in range(3):
    ops.attn_fwd(q, k, v, B, H, L, L, hd, 1 / math.sqrt(hd), True, out=o, lse=lse)
torch.cuda.synchronize()
tr = trace.cpu().view(nsm, 6, 128)
names = ["softmax0", "softmax1", "mma0", "mma1", "kvprod", "epilogue"]
for cta in (0, nsm - 1):
    t = tr[cta]
    t0 = min(int(x) for x in t.flatten() if x > 0)
    print(f"--- CTA {cta} (first 128 stamps per role)")
    for r in range(6):
        row = [int(x) - t0 for x in t[r] if x > 0]
        print(f"{names[r]:9s}", " ".join(f"{x}" for x in row))
