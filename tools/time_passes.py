"""Device timings of the HBM-bound passes around the Gram at the bench size (run by hand under gpurun)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vmc_pde_b200 import _kernels
n, Pp = 2 ** 18, 8192
O = torch.randn(n, Pp, device="cuda", dtype=torch.float64)
E = torch.randn(n, device="cuda", dtype=torch.float64); lp = torch.randn(n, device="cuda", dtype=torch.float64)
first = _kernels.zeros(4 + Pp); F = _kernels.zeros(Pp); var = _kernels.zeros(1)
dE, wE, wLp = _kernels.zeros(n), _kernels.zeros(n), _kernels.zeros(n)
meanO = _kernels.zeros(Pp)


def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = n * Pp * 8 / 1e9
t = timeit(lambda: _kernels.moments1(E, lp, O, n, Pp, first)); print(f"moments1      {t:7.3f} ms  {gb / t:7.2f} TB/s (read)")
ref = O.sum(0); first.zero_(); _kernels.moments1(E, lp, O, n, Pp, first); print("  colsum relerr", float((first[4:] - ref).abs().max() / ref.abs().max()))
t = timeit(lambda: _kernels.center_force(O, n, Pp, meanO, E, lp, 0.0, dE, wE, wLp, F, var)); print(f"center_force  {t:7.3f} ms  {2 * gb / t:7.2f} TB/s (read+write)")
