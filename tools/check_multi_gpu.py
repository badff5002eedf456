"""One TDVP right-hand side at a moderate size, run (a) on 1 rank and (b) under torchrun with R ranks; both write their
results to gpurun_out/ and `--compare` checks that the sharded run (samples AND the post-tridiagonal solve sharded)
reproduces the single-rank one.  usage (on a multi-GPU box):
    python tools/check_multi_gpu.py --out gpurun_out/mg_1.pt
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py --out gpurun_out/mg_2.pt
    python tools/check_multi_gpu.py --compare gpurun_out/mg_1.pt gpurun_out/mg_2.pt"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--out"); ap.add_argument("--compare", nargs=2)
args = ap.parse_args()
if args.compare:
    a, b = torch.load(args.compare[0]), torch.load(args.compare[1])
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    S0 = a["S0"]
    du = b["update"] - a["update"]
    print("S0", rel(b["S0"], a["S0"]), "F0", rel(b["F0"], a["F0"]), "ev", rel(b["ev"], a["ev"]), "VtF^2", rel(b["VtF"] ** 2, a["VtF"] ** 2))
    big = (a["ev"] / a["ev"][-1]).abs() > 1e-8
    print("snr(big)", float((b["snr"][big] / a["snr"][big] - 1).abs().max()), "residual", float(a["res"]), float(b["res"]),
          "tdvp_error", float(a["err"]), float(b["err"]))
    sn = float(du @ S0 @ du) / float(a["update"] @ S0 @ a["update"])
    print("update S-norm rel diff", sn)
    Va, Vb = a["V"], b["V"]
    print("|S V - V ev| (sharded V)", float((S0 @ Vb - Vb * b["ev"]).abs().max() / a["ev"][-1]))
    print("v^T SExp v: 1 rank eager", a["q_eager"], "lazy", a["q_lazy"], "| R ranks eager", b["q_eager"], "lazy", b["q_lazy"],
          "| lazy update == eager update:", bool(torch.equal(a["upd_lazy"], a["update"])), bool(torch.equal(b["upd_lazy"], b["update"])))
    ql = max(abs(a["q_lazy"] / a["q_eager"] - 1), abs(b["q_lazy"] / b["q_eager"] - 1), abs(b["q_eager"] / a["q_eager"] - 1))
    ok = ql < 1e-11 and rel(b["S0"], a["S0"]) < 1e-11 and rel(b["ev"], a["ev"]) < 1e-11 and sn < 1e-14 and abs(float(a["err"]) - float(b["err"])) < 1e-10
    print("MULTI-GPU CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp
d, depth, h, N = 2, 4, 85, 2 ** 14           # BASELINE config C2 architecture (P = 2053)
off = np.zeros(d)
smp = sampler.Sampler(dim=d, numChains=30, name="Gauss", mcmc_info={"offset": off, "bound": 0.25})
vs = var_state.VarState(smp, d, 1, depth, network_args={"intmediate": (h,), "offset": off, "latentSpaceName": "Gauss", "dim": d})
eq = evolutionEq.EvolutionEquation(dim=d, name="diffusion")
T = tdvp.TDVP()
KEY0, THETA0 = vs.sampler.key.copy(), vs.get_parameters().clone()
upd, info = T(THETA0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
V = T.V   # collective when sharded
# SExp through the quadratic form: eager matrix vs the matrix-free operator (collective with several ranks)
vq = torch.linspace(-1, 1, vs.numParameters, device=upd.device, dtype=torch.float64)
q_eager = float(vq @ T.SExp @ vq)
T2 = tdvp.TDVP(computeSExp="lazy")
vs.sampler.key = KEY0.copy()   # same samples for the second evaluation
upd2, _ = T2(THETA0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
q_lazy = float(vq @ T2.SExp @ vq)
# every rank must hold the same solve results (the scripts above only look at rank 0)
fp = torch.stack([T.ev.sum(), T.ev[-1], upd.norm(), T.VtF.norm(), T.snr.norm(), T.invEv.sum(), T.solverResidual, T.tdvp_error,
                  info["entropy"], info["max_grad"], V.abs().sum()]).to(torch.float64)
if world > 1:
    allfp = [torch.empty_like(fp) for _ in range(world)]
    dist.all_gather(allfp, fp)
    dev = max(float(((a - allfp[0]).abs() / (allfp[0].abs() + 1e-300)).max()) for a in allfp)
    if rank == 0:
        print("max relative deviation of the per-rank fingerprints:", dev, flush=True)
        for r_, a in enumerate(allfp):
            print("  rank", r_, [f"{float(x):.15e}" for x in a], flush=True)
if rank == 0:
    torch.save({"update": upd.cpu(), "S0": T.S0.cpu(), "F0": T.F0.cpu(), "ev": T.ev.cpu(), "VtF": T.VtF.cpu(), "snr": T.snr.cpu(),
                "res": T.solverResidual.cpu(), "err": T.tdvp_error.cpu(), "V": V.cpu(), "entropy": info["entropy"].cpu(), "q_eager": q_eager, "q_lazy": q_lazy,
                "upd_lazy": upd2.cpu()}, args.out)
    print("rank 0 wrote", args.out, "P", vs.numParameters, "world", world, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
