"""attn_fwd must be bit-reproducible run to run (any difference is a race)."""
import math, sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, H, Tq, Tk, hd, causal) in [(8, 32, 835, 835, 128, 1), (3, 4, 70, 70, 128, 1), (8, 16, 577, 577, 64, 0), (3, 4, 17, 17, 64, 0), (2, 2, 1603, 1603, 128, 1), (3, 2, 300, 1000, 128, 1)]:
    W = H * hd
    q = torch.randn(B * Tq, W, device=dev).bfloat16(); k = torch.randn(B * Tk, W, device=dev).bfloat16(); v = torch.randn(B * Tk, W, device=dev).bfloat16()
    lse0 = torch.empty(B, H, Tq, device=dev); o0 = torch.empty(B * Tq, W, device=dev, dtype=torch.bfloat16)
    ops.attn_fwd(q, k, v, B, H, Tq, Tk, hd, 1 / math.sqrt(hd), causal, out=o0, lse=lse0)
    bad = 0
    for it in range(30):
        lse = torch.full((B, H, Tq), float("nan"), device=dev); o = torch.full((B * Tq, W), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attn_fwd(q, k, v, B, H, Tq, Tk, hd, 1 / math.sqrt(hd), causal, out=o, lse=lse)
        if not (torch.equal(o, o0) and torch.equal(lse, lse0)):
            bad += 1
            d = (o.float() - o0.float()).abs()
            idx = d.view(B, Tq, H, hd).amax(-1).nonzero()[:5].tolist()
            print("  mismatch iter", it, "max", d.max().item(), "n", int((d > 0).sum()), "where (b,t,h)", idx, "lse diff", int((lse != lse0).sum()))
    print((B, H, Tq, Tk, hd, causal), "mismatching runs:", bad)
