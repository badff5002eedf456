"""Device time of the fused local-terms kernel at the bench size (run by hand under gpurun)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from vmc_pde_b200 import _kernels
vs, eq, T, st = bench.build_ours()
h = vs.net.handle
n = 2 ** 17
key = vs.sampler.next_key()
x, lp = vs.sample_range(key, 0, n, n)
O = _kernels.empty(n, h.Pp)
e = eq.equation_struct(0.0)
def run(): _kernels.local_terms(h, vs._flat, x, e, O=O, ldo=h.Pp, want=("eloc", "logp"))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"local_terms n={n} P={h.P}: {ms:.3f} ms  ({n * h.Pp * 8 / ms / 1e9:.2f} TB/s of O written; x2 for N=2^18)")

def samp(): vs.sample_range(key, 0, n, n)
samp(); torch.cuda.synchronize()
e0.record()
for _ in range(5): samp()
e1.record(); torch.cuda.synchronize()
print(f"sample+logp n={n}: {e0.elapsed_time(e1) / 5:.3f} ms")
