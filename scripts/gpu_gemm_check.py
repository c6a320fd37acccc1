"""Bring-up diagnostic for the tcgen05 GEMM (run on a B200 through gpurun).  Prints, per case, the
max error against torch fp32 matmul of the same bf16 inputs and a coarse error map so a wrong
descriptor / swizzle shows up as a pattern.  Not a test: tests/test_gemm_gpu.py is."""
import ctypes, sys, time
import torch
from a3vlm_b200 import _lib

lib = _lib.load(strict=False)
dev = torch.device("cuda:0")
torch.manual_seed(0)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(A, B, M, N, K, a_mn, b_mn, out_fp32=True, bias=None, residual=None, act=0, accumulate=0,
         C=None, alpha=1.0):
    if C is None:
        C = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.a3_gemm_bf16(P(A), P(B), P(C), M, N, K, A.stride(0), B.stride(0), C.stride(0), a_mn, b_mn,
                          P(bias), P(residual), residual.stride(0) if residual is not None else 0,
                          act, int(out_fp32), accumulate, alpha, None, 0, 1, 0, ctypes.c_void_p(st))
    if rc != 0:
        raise RuntimeError(lib.a3_last_error().decode())
    torch.cuda.synchronize()
    return C


def case(M, N, K, a_mn, b_mn):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    ref = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    try:
        C = gemm(A, B, M, N, K, a_mn, b_mn)
    except Exception as e:
        print(f"M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn}: EXC {e}")
        return False
    err = (C - ref).abs()
    tol = 1e-3 * ref.abs() + 2e-2 * (K ** 0.5) * 1e-2
    ok = bool((err <= tol + 1e-4 * K ** 0.5).all())
    print(f"M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn}: max_err {err.max().item():.4g} ref_max {ref.abs().max().item():.4g} ok={ok}")
    if not ok:
        # error map per 32x32 block (first 128x256 tile)
        e = err[: min(M, 128), : min(N, 256)]
        for r0 in range(0, e.shape[0], 32):
            print("   ", " ".join(f"{e[r0:r0+32, c0:c0+32].max().item():8.2g}" for c0 in range(0, e.shape[1], 32)))
        nz = (C != 0).float().mean().item()
        print(f"    nonzero fraction {nz:.3f}; C[0,:4]={C[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}")
    return ok


allok = True
for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 512), (200, 264, 200),
                  (4616, 1024, 1024), (6680, 12288, 4096)]:
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        if M * N * K > 1e11 and (a_mn, b_mn) != (0, 0):
            continue
        allok &= case(M, N, K, a_mn, b_mn)
print("ALL OK" if allok else "SOME FAILED")

# quick timing of the big shapes
def bench(M, N, K, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): gemm(a, b, M, N, K, 0, 0, out_fp32=False, C=C)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    st = torch.cuda.current_stream().cuda_stream
    e0.record()
    for _ in range(iters):
        lib.a3_gemm_bf16(P(a), P(b), P(C), M, N, K, K, K, N, 0, 0, None, None, 0, 0, 0, 0, 1.0, None, 0, 1, 0, ctypes.c_void_p(st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0 = torch.cuda.Event(True); t1 = torch.cuda.Event(True)
    for _ in range(3): torch.matmul(a, b.t())
    t0.record()
    for _ in range(iters): torch.matmul(a, b.t())
    t1.record(); torch.cuda.synchronize()
    ms2 = t0.elapsed_time(t1) / iters
    print(f"bench M{M} N{N} K{K}: a3 {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.1f} TFLOP/s")

if allok or "--bench" in sys.argv:
    for s in [(6680, 12288, 4096), (6680, 4096, 4096), (6680, 22016, 4096), (6680, 4096, 11008), (4616, 3072, 1024), (4616, 1024, 4096), (8192, 8192, 8192)]:
        bench(*s)
