"""Times vmcpde_gram on synthetic O (CUDA events). Used for ncu captures: python tools/gram_bench.py N PP MATS [REPS]"""
import sys, json
import torch
sys.path.insert(0, "/root/repo")
from vmc_pde_b200 import _lib
n, Pp, nm = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = _lib.load()
dev = torch.device("cuda:0")
O = torch.randn(n, Pp, device=dev, dtype=torch.float64)
w = [None] + [torch.rand(n, device=dev, dtype=torch.float64) for _ in range(nm - 1)]
S = [torch.zeros(Pp, Pp, device=dev, dtype=torch.float64) for _ in range(nm)]
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
ts = []
for it in range(reps):
    torch.cuda.synchronize()
    e0.record()
    _lib.check(L.vmcpde_gram(_lib.ptr(O), n, Pp, Pp, nm, _lib.ptr_array(w), _lib.ptr_array(S), _lib.stream()))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
fl = nm * n * Pp * (Pp + 1.0)
print(json.dumps({"n": n, "Pp": Pp, "mats": nm, "ms": ts, "tflops_best": fl / min(ts) * 1e-9}))
