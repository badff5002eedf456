#!/usr/bin/env python
"""Benchmark of the TDVP time-step hot path (BASELINE.json metric: TDVP steps/sec at N=2^18 samples, P~8k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE configs[2] / SURVEY 8d "C3": d=6, depth 8, intmediate (36,), different_add couplings (P = 8187),
Gauss latent, offset [1,0,0,1,0,0], evolution 'advection_hamiltonian_wDiss', N = 2^18 samples in total,
TDVP() defaults (svdTol 1e-11, eigen-solve, SExp and SNR computed), one step = FixedStepper(mode='Heun').step =
2 right-hand sides.  N > 1 GPUs shard the samples (strong scaling of the same step; NCCL all-reduce of the packed
first and second moments; the P x P solve is replicated).

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/) on the
host cores on a bounded sample of the same workload (the reference itself needs jax 0.2.18 / flax 0.3.6, which are
not installable here: SURVEY 8c).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(dim=6, depth=8, hidden=36, variant="different_add", latent="Gauss", equation="advection_hamiltonian_wDiss",
          offset=[1.0, 0, 0, 1, 0, 0], n_samples=2 ** 18)
METRIC = "TDVP steps/sec at N=2^18 samples, P~8k params (one step = one Heun step = 2 RHS)"


def workload_config(n_gpus):
    return {"workload": "C3: 6D phase-space Fokker-Planck (advection_hamiltonian_wDiss), INN depth 8 x (36,) different_add, "
                        "P=8187, N=2^18 samples total, TDVP defaults (eigh solve, svdTol=1e-11), FixedStepper Heun",
            "n_samples": C3["n_samples"], "num_params": 8187, "dim": 6, "rhs_per_step": 2,
            "parallelism": f"samples sharded over {n_gpus} GPU(s); tridiagonalisation + divide&conquer replicated, "
                           "back-transformation / SNR / update sharded over eigenvectors",
            "l2": "no explicit flush: each RHS streams the 17 GB centred O matrix (>> 126 MB L2) three times"}


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def build_ours():
    from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper, net
    off = np.asarray(C3["offset"], dtype=np.float64)
    smp = sampler.Sampler(dim=C3["dim"], numChains=30, name=C3["latent"], mcmc_info={"offset": off, "bound": 0.25})
    net.SingleBlock.different_add = True          # main.py:46-48: harmonicOsc uses "DifferentAdd"
    try:
        vs = var_state.VarState(smp, C3["dim"], 1, C3["depth"], network_args={"intmediate": (C3["hidden"],), "offset": off,
                                                                          "latentSpaceName": C3["latent"], "dim": C3["dim"]})
    finally:
        net.SingleBlock.different_add = False
    eq = evolutionEq.EvolutionEquation(dim=C3["dim"], name=C3["equation"])
    T = tdvp.TDVP()
    st = stepper.FixedStepper(timeStep=1e-4, mode='Heun', maxStep=1e-2, increase_fac=1.3)   # main.py:51,113
    return vs, eq, T, st


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vmc_pde_b200 import _kernels, _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n_gpus = world
    vs, eq, T, st = build_ours()
    P, N = vs.numParameters, C3["n_samples"]
    norm_fun = lambda v, S: v @ S @ v
    rhs = dict(evolutionEq=eq, psi=vs, nSamplesTDVP=N, nSamplesObs=N, normFunction=norm_fun, timings=None, integrals=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        y, dt, info = st.step(0, T, vs.get_parameters(), **rhs)
        vs.set_parameters(y)
        return info

    theta_init = vs.get_parameters().clone()
    host_theta = torch.empty(P, dtype=torch.float64).pin_memory()
    host_theta.copy_(vs.get_parameters().cpu())
    host_out = torch.empty(P + 3, dtype=torch.float64).pin_memory()

    def step_e2e():
        """The call a user of the reference makes (main.py:161-162) with HOST parameter buffers: H2D of theta, the step,
        D2H of the new theta and of the logged scalars."""
        y0 = host_theta.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
        y, dt, info = st.step(0, T, y0, **rhs)
        vs.set_parameters(y)
        out = torch.cat([y, torch.stack([info["entropy"], T.solverResidual, T.tdvp_error])])
        host_out.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        host_theta.copy_(host_out[:P])
        return float(host_out[P])

    for _ in range(args.warmup):
        step_device()
    # ---- device-resident timing ----
    gram_events, eigh_events = [], []
    orig_gram, orig_eigh, orig_eigh_cols = _kernels.gram, _kernels.eigh, _kernels.eigh_cols

    def timed_eigh(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_eigh(*a)
        e1.record()
        eigh_events.append((e0, e1))

    def timed_eigh_cols(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_eigh_cols(*a)
        e1.record()
        eigh_events.append((e0, e1))

    def timed_gram(O, n, ldo, Pp, weights, mats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_gram(O, n, ldo, Pp, weights, mats)
        e1.record()
        gram_events.append((e0, e1, n, len(mats)))

    _kernels.gram, _kernels.eigh, _kernels.eigh_cols = timed_gram, timed_eigh, timed_eigh_cols
    import vmc_pde_b200.tdvp as _t
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _kernels.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        info = step_device()
    ev1.record()
    barrier()
    launches = _kernels.launches - launches0
    ms = ev0.elapsed_time(ev1)
    _kernels.gram, _kernels.eigh, _kernels.eigh_cols = orig_gram, orig_eigh, orig_eigh_cols
    eigh_ms = [a.elapsed_time(b) for a, b in eigh_events]
    gram_ms = [a.elapsed_time(b) for a, b, _, _ in gram_events]
    gram_flops = [m * n * P * (P + 1.0) for _, _, n, m in gram_events]   # SYRK convention, true P
    # ---- end-to-end timing through host buffers ----
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        ent = step_e2e()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    clk = clocks.stop() if rank == 0 else None
    # ---- reported beside the headline, never instead of it: SExp as a matrix-free operator (TDVP(computeSExp="lazy")).
    # The reference builds SExp every call (tdvp.py:47) although only AdaptiveHeun reads it, and only as v^T SExp v
    # (stepper.py:71); FixedStepper never does.  Same step, one Gram out of three not formed.
    T.computeSExp = "lazy"
    step_device()
    barrier()
    tl0 = torch.cuda.Event(enable_timing=True); tl1 = torch.cuda.Event(enable_timing=True)
    tl0.record()
    for _ in range(2):
        step_device()
    tl1.record()
    barrier()
    ms_lazy = tl0.elapsed_time(tl1) / 2
    T.computeSExp = True
    # second variant: the shifted-Cholesky solve (north-star item 4 allows "diagonal shift ... blocked Cholesky"): no
    # eigendecomposition, hence no SNR Gram either; the serial fraction of the multi-GPU step all but disappears
    from vmc_pde_b200 import tdvp as _tdvp
    Tc = _tdvp.TDVP(diagonalShift=1e-4, solver="cholesky")
    vs.set_parameters(theta_init)      # timing variant: start again from the initial state with a fresh, small step
    from vmc_pde_b200 import stepper as _stepper
    st = _stepper.FixedStepper(timeStep=1e-4, mode='Heun', maxStep=1e-2, increase_fac=1.3)
    st.step(0, Tc, vs.get_parameters(), **rhs)
    barrier()
    tc0 = torch.cuda.Event(enable_timing=True); tc1 = torch.cuda.Event(enable_timing=True)
    tc0.record()
    for _ in range(2):
        y, _, _ = st.step(0, Tc, vs.get_parameters(), **rhs)
        vs.set_parameters(y)
    tc1.record()
    barrier()
    ms_chol = tc0.elapsed_time(tc1) / 2
    del Tc
    adaptive = None
    if args.adaptive:   # C3 with the adaptive integrator -- one attempt = 5 right-hand sides + the SExp error norm
        from vmc_pde_b200 import stepper as _stp
        out_a = {}
        for mode in (True, "lazy"):
            vs.set_parameters(theta_init)
            Ta = _tdvp.TDVP(computeSExp=mode)
            ah = _stp.AdaptiveHeun(timeStep=1e-4, tol=1e-2, maxStep=1e-2)   # main.py:109-112
            barrier()
            ta0 = torch.cuda.Event(enable_timing=True); ta1 = torch.cuda.Event(enable_timing=True)
            calls0 = _kernels.launches
            ta0.record()
            y, dt_used, _ = ah.step(0, Ta, vs.get_parameters(), **rhs)
            ta1.record()
            barrier()
            out_a["eager_SExp" if mode is True else "lazy_SExp"] = {"seconds_per_accepted_step": ta0.elapsed_time(ta1) * 1e-3, "dt": float(dt_used)}
            del Ta
        adaptive = out_a
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_lazy, ms_chol], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_lazy, ms_chol = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # FP64 tensor peak, measured live on the warm GPU right after the timed steps (MEASURED_PEAKS.json has no FP64 entry)
    peak = _kernels.dmma_peak_tflops()
    value = args.steps / (ms * 1e-3)
    traffic = None
    prof = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    achieved = sum(gram_flops) / (sum(gram_ms) * 1e-3) * 1e-12 if gram_ms else None
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n_gpus),
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": P * 8, "d2h_bytes_per_step": (P + 3) * 8},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gram_kernel (3 weighted FP64 SYRKs per RHS: S0, SExp, SNR covariance)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved and peak else None,
                     "traffic": traffic,
                     "peak_source": "measured live: register-resident DMMA.8x8x4 loop (vmcpde_dmma_peak); MEASURED_PEAKS.json has no FP64 entry",
                     "flops_convention": "n_mats * n * P * (P+1) per launch (SYRK, SURVEY 8d); launches per step: 2",
                     "share_of_step": sum(gram_ms) / ms if gram_ms else None},
        "stages_ms_per_rhs": {"gram": sum(gram_ms) / max(len(eigh_ms), 1), "eigh": sum(eigh_ms) / max(len(eigh_ms), 1),
                              "everything_else": (ms - sum(gram_ms) - sum(eigh_ms)) / max(len(eigh_ms), 1)},
        "variants": {"lazy_SExp_steps_per_s": 1e3 / ms_lazy, "cholesky_shift1e-4_steps_per_s": 1e3 / ms_chol,
                     "note": "TDVP(computeSExp='lazy'): SExp kept as a matrix-free operator on the resident O (2 Grams per RHS "
                             "instead of 3); cholesky: TDVP(diagonalShift=1e-4, solver='cholesky') replaces the eigen-solve by the "
                             "tensor-core blocked Cholesky (S0 and SExp Grams only).  Neither is the headline -- the reference forms "
                             "SExp every call and regularises through the eigendecomposition"},
        "last_entropy": ent,
    }
    if adaptive is not None:
        line["variants"]["adaptive_heun"] = adaptive
    # CPU arm: rank 0 at N=1 only (under torchrun the host threads are pinned to 1 per rank; see --impl reference)
    line["cpu_baseline"] = cpu_baseline(bounded_seconds=True) if world == 1 else None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
_cpu_state = {}


def cpu_step_estimate(n_s=1024, p_s=2048):
    """One Heun step of the CPU restatement (oracle/, the reference's algorithm incl. its three N*P^2 products and host
    eigh) on a bounded sample: local terms and Grams on n_s samples (linear in N), eigh on a p_s x p_s block (cubic in P),
    extrapolated to N = 2^18, P = 8187.  Returns (seconds per step, breakdown)."""
    import torch
    from oracle import flow, tdvp
    if "st" not in _cpu_state:
        ups, downs, _ = flow.make_index_splits(C3["dim"], C3["depth"], 1)
        spec = flow.FlowSpec(dim=C3["dim"], depth=C3["depth"], hidden=(C3["hidden"],), variant=C3["variant"], latent=C3["latent"],
                             offset=np.asarray(C3["offset"]), inds_up=ups, inds_down=downs)
        _cpu_state["st"] = flow.OracleState(spec, flow.init_params(spec, 1))
        _cpu_state["P"] = spec.num_params
    st, P, N = _cpu_state["st"], _cpu_state["P"], C3["n_samples"]
    t0 = time.perf_counter()
    x, lp, _ = st.sample(n_s)
    t1 = time.perf_counter()
    E, O, lp2, _ = tdvp.local_terms(st, x, C3["equation"], 0.0)
    t2 = time.perf_counter()
    On, En, lpn = O.numpy(), E.numpy(), lp2.numpy()
    dO = On - On.mean(0)
    dE = En - En.mean()
    S0 = dO.T @ dO / n_s                                   # tdvp.py:46
    w = lpn[:, None] * dO
    SExp = w.T @ w / n_s                                   # tdvp.py:47
    EO = dE[:, None] * dO
    F = EO.mean(0)
    t3 = time.perf_counter()
    ev, V = np.linalg.eigh(S0[:p_s, :p_s])                 # tdvp.py:61-64 on a block
    t4 = time.perf_counter()
    EOv = EO[:, :p_s] @ V                                  # tdvp.py:68 on the block (N x P x P in the reference)
    t5 = time.perf_counter()
    lin = N / n_s
    rhs = ((t1 - t0) + (t2 - t1) + (t3 - t2)) * lin + (t4 - t3) * (P / p_s) ** 3 + (t5 - t4) * lin * (P / p_s) ** 2
    parts = {"sampling_s": (t1 - t0) * lin, "local_terms_s": (t2 - t1) * lin, "two_grams_and_F_s": (t3 - t2) * lin,
             "eigh_s": (t4 - t3) * (P / p_s) ** 3, "EO_at_V_s": (t5 - t4) * lin * (P / p_s) ** 2}
    return 2.0 * rhs, parts


def cpu_baseline(bounded_seconds=True):
    import torch
    cpu_step_estimate(n_s=256)                            # warm-up (torch.func tracing)
    sec, parts = cpu_step_estimate()
    return {"value": 1.0 / sec, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/ restatement (torch float64 CPU, not JAX): local terms + 2 Grams + F on 1024 of 2^18 samples "
                      "(scaled x256), eigh on a 2048 block (scaled (8187/2048)^3), EOdata@V on the block (scaled); "
                      "seconds per RHS at full size: " + ", ".join(f"{k}={v:.0f}" for k, v in parts.items())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses all the host threads it can
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    cpu_step_estimate(n_s=256)
    for _ in range(args.warmup):
        cpu_step_estimate(n_s=256)
    secs = []
    for _ in range(args.steps):
        s, parts = cpu_step_estimate()
        secs.append(s)
    sec = sum(secs) / len(secs)
    value = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "each step: bounded sample of the C3 workload (1024 of 2^18 samples, 2048-block eigh) "
                                       "extrapolated linearly in N and cubically in P; reference = oracle/ port, JAX is not "
                                       "installable in this image"},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--adaptive", action="store_true",
                    help="also time AdaptiveHeun attempts (5 RHS each, stepper.py:54-66) on the same workload; adds ~30 s")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
