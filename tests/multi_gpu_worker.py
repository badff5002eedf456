"""Worker of tests/test_gpu_multi.py: one TDVP right-hand side (C2 architecture, P = 2053) on `WORLD_SIZE` ranks.
Deliberately does NOT call torch.cuda.set_device: the package binds cuda:LOCAL_RANK itself (global_defs.device)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--out"); ap.add_argument("--pipeline", type=int, default=1); ap.add_argument("--n", type=int, default=2 ** 14)
ap.add_argument("--share", type=float, default=-1.0)
args = ap.parse_args()
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, global_defs
d, depth, h, N = 2, 4, 85, args.n
off = np.zeros(d)
smp = sampler.Sampler(dim=d, numChains=30, name="Gauss", mcmc_info={"offset": off, "bound": 0.25})
vs = var_state.VarState(smp, d, 1, depth, network_args={"intmediate": (h,), "offset": off, "latentSpaceName": "Gauss", "dim": d})
assert torch.cuda.current_device() == local and vs.get_parameters().device.index == local
eq = evolutionEq.EvolutionEquation(dim=d, name="diffusion")
T = tdvp.TDVP(pipelineSolve=bool(args.pipeline), solverShare=None if args.share < 0 else args.share)
KEY0, THETA0 = vs.sampler.key.copy(), vs.get_parameters().clone()
part = T.sample_partition(N, world, vs.numParameters)
upd, info = T(THETA0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
V = T.V   # collective when sharded
vq = torch.linspace(-1, 1, vs.numParameters, device=upd.device, dtype=torch.float64)
q_eager = float(vq @ T.SExp @ vq)
T2 = tdvp.TDVP(computeSExp="lazy", pipelineSolve=bool(args.pipeline), solverShare=None if args.share < 0 else args.share)
vs.sampler.key = KEY0.copy()
upd2, _ = T2(THETA0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
q_lazy = float(vq @ T2.SExp @ vq)
mx = lambda a, b: float((a - b).abs().max())
lazy_diffs = {"S0": mx(T.S0, T2.S0), "F0": mx(T.F0, T2.F0), "ev": mx(T.ev, T2.ev), "VtF": mx(T.VtF, T2.VtF), "invEv": mx(T.invEv, T2.invEv),
              "update": mx(upd, upd2), "S": mx(T.S, T2.S), "VT": mx(T._VT, T2._VT), "ZT": mx(T._ZT, T2._ZT) if T._ZT is not None else -1.0,
              "refl": mx(T._Swork, T2._Swork), "rhoVar": mx(T.rhoVar, T2.rhoVar), "CEO": mx(T._mats(vs.net.handle.Pp)[2], T2._mats(vs.net.handle.Pp)[2])}
fp = torch.stack([T.ev.sum(), T.ev[-1], T.ev[0], upd.norm(), T.VtF.norm(), T.snr.norm(), T.invEv.sum(), T.solverResidual, T.tdvp_error,
                  info["entropy"], info["max_grad"], V.abs().sum()]).to(torch.float64)
# the C-ABI's own NCCL entry points (for hosts without torch.distributed) against torch.distributed on the same data
nccl_dev = 0.0
if world > 1:
    import ctypes as C
    from vmc_pde_b200 import _lib
    L = _lib.load()
    ident = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.vmcpde_nccl_unique_id(ident))
    box = [ident.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ident = C.create_string_buffer(box[0], 128)
    comm = C.c_void_p()
    _lib.check(L.vmcpde_nccl_comm_init(world, rank, ident, C.byref(comm)))
    Pq = 256
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    mats = [torch.randn(Pq, Pq, device="cuda", dtype=torch.float64, generator=g) for _ in range(3)]
    tail = torch.randn(Pq + 8, device="cuda", dtype=torch.float64, generator=g)
    ref = [m.clone() for m in mats] + [tail.clone()]
    for t in ref:
        dist.all_reduce(t)
    pk = torch.empty(3 * L.vmcpde_packed_tiles_len(Pq) + Pq + 8, device="cuda", dtype=torch.float64)
    _lib.check(L.vmcpde_allreduce_moments(comm, _lib.ptr_array(mats), 3, Pq, _lib.ptr(tail), Pq + 8, _lib.ptr(pk), _lib.stream()))
    torch.cuda.synchronize()
    nccl_dev = max(float((torch.triu(a) - torch.triu(b)).abs().max()) for a, b in zip(mats, ref[:3]))
    nccl_dev = max(nccl_dev, float((tail - ref[3]).abs().max()))
    _lib.check(L.vmcpde_nccl_comm_destroy(comm))
fps = [fp]
if world > 1:
    fps = [torch.empty_like(fp) for _ in range(world)]
    dist.all_gather(fps, fp)
if rank == 0:
    torch.save({"update": upd.cpu(), "S0": T.S0.cpu(), "F0": T.F0.cpu(), "ev": T.ev.cpu(), "VtF": T.VtF.cpu(), "snr": T.snr.cpu(),
                "res": T.solverResidual.cpu(), "err": T.tdvp_error.cpu(), "V": V.cpu(), "entropy": info["entropy"].cpu(),
                "covar": info["covar"].cpu(), "q_eager": q_eager, "q_lazy": q_lazy, "upd_lazy": upd2.cpu(), "SExp": T.SExp.cpu(),
                "fingerprints": torch.stack(fps).cpu(), "partition": part, "lazy_diffs": lazy_diffs, "nccl_dev": nccl_dev}, args.out)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
