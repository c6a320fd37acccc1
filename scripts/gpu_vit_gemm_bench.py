import sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M = 4616
for (N, K) in [(4096, 1024), (1024, 4096), (3072, 1024), (1024, 1024)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16(); res = torch.randn(M, N, device=dev).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    r = {}
    r["plain"] = t(lambda: ops.gemm(a, w, out))
    r["bias"] = t(lambda: ops.gemm(a, w, out, bias=bias))
    r["bias+gelu"] = t(lambda: ops.gemm(a, w, out, bias=bias, act=ops.ACT_GELU))
    r["bias+quickgelu"] = t(lambda: ops.gemm(a, w, out, bias=bias, act=ops.ACT_QUICKGELU))
    r["bias+res"] = t(lambda: ops.gemm(a, w, out, bias=bias, residual=res))
    r["cublas"] = t(lambda: torch.matmul(a, w.t()))
    print(f"M{M} N{N} K{K}: " + "  ".join(f"{k} {v:.1f}us ({fl/v/1e6:.0f} TF)" for k, v in r.items()))
