import sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M, D, F = 6680, 4096, 11008
x = torch.randn(M, D, device=dev).bfloat16(); res = torch.randn(M, D, device=dev).bfloat16()
wqkv = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16(); wo = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
w13 = (torch.randn(2 * F, D, device=dev) * 0.02).bfloat16(); w2 = (torch.randn(D, F, device=dev) * 0.02).bfloat16()
y = torch.randn(M, F, device=dev).bfloat16()
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16); o = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
gu = torch.empty(M, 2 * F, device=dev, dtype=torch.bfloat16); yy = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
dgu = torch.empty_like(gu); dy = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
rope = ops.rope_table(128, 2048, device=dev)
rows = [
 ("qkv plain", 2.0*M*3*D*D, lambda: ops.gemm(x, wqkv, qkv)),
 ("qkv +rope", 2.0*M*3*D*D, lambda: ops.gemm(x, wqkv, qkv, rope=rope, rope_cols=2*D, rope_period=835)),
 ("wo plain", 2.0*M*D*D, lambda: ops.gemm(x, wo, o)),
 ("wo +res", 2.0*M*D*D, lambda: ops.gemm(x, wo, o, residual=res)),
 ("w13 plain", 2.0*M*2*F*D, lambda: ops.gemm(x, w13, gu)),
 ("w13 fused swiglu", 2.0*M*2*F*D, lambda: ops.gemm_swiglu(x, w13, gu, yy)),
 ("w2 plain", 2.0*M*D*F, lambda: ops.gemm(y, w2, o)),
 ("w2 +res", 2.0*M*D*F, lambda: ops.gemm(y, w2, o, residual=res)),
 ("w2 dgrad plain", 2.0*M*D*F, lambda: ops.gemm(x, w2, dy, b_mn=True)),
 ("w2 dgrad fused dswiglu", 2.0*M*D*F, lambda: ops.gemm_dswiglu(x, w2, gu, dgu)),
 ("wgrad w13 (a_mn,b_mn)", 2.0*M*2*F*D, lambda: ops.gemm(gu, x, w13, a_mn=True, b_mn=True)),
]
for name, fl, fn in rows:
    us = t(fn)
    print(f"{name:26s} {us:8.1f} us  {fl/us/1e6:6.0f} TFLOP/s")
