"""BASELINE configs[3] ("C4"): 10-D Fokker-Planck (diffusion), INN depth 4 x (185,), P = 16385, N = 2^20 samples sharded over
the ranks of one box.  One TDVP right-hand side, timed, with the size-independent property checks of
tests/test_gpu_tdvp.py::test_full_size_properties_c3.  usage:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c4.py [log2_samples]"""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp
d, depth, h = 10, 4, 185
N = 2 ** (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
off = np.zeros(d)
smp = sampler.Sampler(dim=d, numChains=30, name="Gauss", mcmc_info={"offset": off, "bound": 0.25})
vs = var_state.VarState(smp, d, 1, depth, network_args={"intmediate": (h,), "offset": off, "latentSpaceName": "Gauss", "dim": d})
eq = evolutionEq.EvolutionEquation(dim=d, name="diffusion")
T = tdvp.TDVP()
times = []
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.time()
    upd, info = T(vs.get_parameters(), 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
    torch.cuda.synchronize()
    times.append(time.time() - t0)
ok = bool(torch.isfinite(upd).all()) and torch.equal(T.S0, T.S0.T) and float(T.ev[0]) > -1e-10 * float(T.ev[-1])
ok = ok and abs(float(T.ev.sum()) / float(torch.diagonal(T.S0).sum()) - 1) < 1e-10
ok = ok and float(T.solverResidual) < 1e-6 and 0 <= float(T.tdvp_error) < 1
ok = ok and abs(float(info["entropy"]) - 0.5 * d * np.log(2 * np.pi * np.e)) < 0.05
checks = [bool(torch.isfinite(upd).all()), bool(torch.equal(T.S0, T.S0.T)), float(T.ev[0]) / float(T.ev[-1]),
          float(T.ev.sum()) / float(torch.diagonal(T.S0).sum()) - 1, float(T.solverResidual), float(T.tdvp_error), float(info["entropy"])]
print(f"[rank {rank}] checks {checks} ok={bool(ok)}", file=sys.stderr, flush=True)
if rank == 0:
    print(json.dumps({"config": "C4", "P": int(vs.numParameters), "N": N, "gpus": world, "rhs_seconds": times,
                      "residual": float(T.solverResidual), "tdvp_error": float(T.tdvp_error), "entropy": float(info["entropy"]),
                      "ev_max": float(T.ev[-1]), "modes_above_cutoff": int((T.ev / T.ev[-1] > 1e-11).sum()), "properties_ok": bool(ok)}, default=str), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if bool(ok) else 1)
