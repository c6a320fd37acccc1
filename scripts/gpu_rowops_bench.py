"""HBM-bound row kernels at the 7B fine-tune shapes (rows = 8 x 835, D = 4096), isolated, CUDA events,
L2 flushed between launches by cycling through 8 distinct buffer sets (8 x ~270 MB > 126 MB L2)."""
import json, sys, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200 import ops
dev = "cuda"
rows, D, NS = 8 * 835, 4096, 8
hbm = bench.peaks()["hbm"]
bf = torch.bfloat16
sets = [dict(x=torch.randn(rows, D, device=dev, dtype=bf), dy=torch.randn(rows, D, device=dev, dtype=bf),
             dx=torch.randn(rows, D, device=dev, dtype=bf), y=torch.empty(rows, D, device=dev, dtype=bf),
             rstd=torch.rand(rows, device=dev) + 0.5) for _ in range(NS)]
w = torch.ones(D, device=dev, dtype=bf)
nparts = 2 * torch.cuda.get_device_properties(0).multi_processor_count
part = torch.empty(nparts, D, device=dev, dtype=torch.float32)
def timeit(fn, n=40):
    for i in range(NS):
        fn(sets[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(sets[i % NS])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
res = {}
t = timeit(lambda s: ops.rmsnorm_bwd(s["dy"], s["x"], w, s["rstd"], s["dx"], part, add_to_dx=True))
b = 4 * rows * D * 2  # x, dy, dx read + dx write
res["rmsnorm_bwd_add"] = {"us": t, "bytes": b, "GBs": b / t / 1e3, "frac_hbm": b / t / 1e3 / hbm}
t = timeit(lambda s: ops.rmsnorm_bwd(s["dy"], s["x"], w, s["rstd"], s["dx"], part, add_to_dx=False))
b = 3 * rows * D * 2
res["rmsnorm_bwd"] = {"us": t, "bytes": b, "GBs": b / t / 1e3, "frac_hbm": b / t / 1e3 / hbm}
t = timeit(lambda s: ops.rmsnorm_fwd(s["x"], w, 1e-5, out=s["y"], rstd=s["rstd"]))
b = 2 * rows * D * 2
res["rmsnorm_fwd"] = {"us": t, "bytes": b, "GBs": b / t / 1e3, "frac_hbm": b / t / 1e3 / hbm}
out = torch.empty(D, device=dev, dtype=bf)
t = timeit(lambda s: ops.colsum_partials(part, out))
res["colsum_partials_296x4096"] = {"us": t}
print(json.dumps(res))
