"""Times the tcgen05 split-precision Gram against the FP64 DMMA Gram on synthetic O (one matrix each).
usage (GPU box): python tools/gram_split_bench.py [n] [P]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_pde_b200 import _kernels, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 18
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8187
Pp = _kernels.round_up(P, 128)
g = torch.Generator(device="cuda"); g.manual_seed(0)
O = torch.randn(n, Pp, device="cuda", dtype=torch.float64, generator=g)
O *= 10.0 ** (-6.0 * torch.arange(Pp, device="cuda", dtype=torch.float64) / P)
O[:, P:] = 0
w = torch.rand(n, device="cuda", dtype=torch.float64)
S1, S2 = _kernels.zeros(Pp, Pp), _kernels.zeros(Pp, Pp)
ev = lambda: torch.cuda.Event(enable_timing=True)
res = {"n": n, "P": P}
reps = int(os.environ.get("REPS", "3"))
for name, fn in (("split", lambda: _kernels.gram_split(O, n, Pp, Pp, w, S1)), ("fp64", lambda: _kernels.gram(O, n, Pp, Pp, [w], [S2]))):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps if name == "split" else 1):
        (S1 if name == "split" else S2).zero_()
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ms = best
    res[name + "_ms"] = ms
    res[name + "_fp64_equiv_tflops"] = n * P * (P + 1.0) / (ms * 1e-3) * 1e-12
tiles = Pp // 128
res["split_bf16_tflops"] = tiles * (tiles + 1) / 2 * 128 * 128 * ((n + 127) // 128 * 128) * 2.0 * 6 / (res["split_ms"] * 1e-3) * 1e-12   # executed bf16 flops
d = torch.sqrt(torch.diagonal(S2)[:P])
iu = torch.triu_indices(P, P, device="cuda")
err = ((S1 - S2)[:P, :P][iu[0], iu[1]].abs() / (d[iu[0]] * d[iu[1]])).max()
res["max_scaled_error"] = float(err)
res["fro_rel_error"] = float((torch.triu(S1 - S2)).norm() / torch.triu(S2).norm())
print(json.dumps(res))
