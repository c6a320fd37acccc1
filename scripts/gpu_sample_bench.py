"""Latency of the device-side token choice: a3_sample_top_p and a3_argmax + a3_generate_step on fp32
logits [B][32000] (CUDA events, 200 launches each after warm-up), next to the torch sequence the
reference runs per token (softmax, sort, cumsum, mask, div, multinomial, gather -- meta.py:457-459,568-583)."""
import json, sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
res = {}
for B in (1, 8, 32):
    g = torch.Generator().manual_seed(B)
    logits = (torch.randn(B, 32000, generator=g) * 3).cuda()
    u = torch.rand(B, generator=g).cuda()
    out = torch.empty(B, dtype=torch.int64, device="cuda")
    def ours():
        ops.sample_top_p(logits, 0.1, 0.75, u, out)
    def ref():
        probs = torch.softmax(logits / 0.1, dim=-1)
        ps, pi = torch.sort(probs, dim=-1, descending=True)
        cs = torch.cumsum(ps, dim=-1)
        ps[cs - ps > 0.75] = 0.0
        ps.div_(ps.sum(dim=-1, keepdim=True))
        return torch.gather(pi, -1, torch.multinomial(ps, num_samples=1))
    def greedy():
        ops.argmax(logits, out)
    r = {}
    for name, fn in (("a3_sample_top_p", ours), ("torch_reference_sequence", ref), ("a3_argmax", greedy)):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        r[name + "_us"] = e0.elapsed_time(e1) * 1e3 / 200
    res[f"B{B}"] = r
print(json.dumps(res))
