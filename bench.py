#!/usr/bin/env python
"""Benchmark of the TDVP time-step hot path (BASELINE.json metric: TDVP steps/sec at N=2^18 samples, P~8k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]

Default workload = BASELINE configs[2] / SURVEY 8d "C3": d=6, depth 8, intmediate (36,), different_add couplings (P = 8187),
Gauss latent, offset [1,0,0,1,0,0], evolution 'advection_hamiltonian_wDiss', N = 2^18 samples in total, TDVP() defaults
(svdTol 1e-11, eigen-solve, SExp and SNR computed), one step = FixedStepper(mode='Heun').step = 2 right-hand sides.
N > 1 GPUs shard the samples of the same step (strong scaling); the eigensolver's serial stages run on one rank while the
others build the SExp / C_EO Grams, its O(P^3) tail is sharded over eigenvectors.
--config C2 / C4 run BASELINE configs[1] / [3] through the same code; --config C5 is the synthetic Gram + solve sweep.

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/) on the host
cores on a bounded sample of the same workload (the reference itself needs jax 0.2.18 / flax 0.3.6, which are not
installable here: SURVEY 8c).
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

CONFIGS = {
    "C2": dict(dim=2, depth=4, hidden=85, variant="no_add", latent="Gauss", equation="diffusion", offset=[0.0, 0.0],
               n_samples=2 ** 16, num_params=2053, dt=1e-7,
               text="C2: 2D Fokker-Planck (diffusion), INN depth 4 x (85,), P=2053, N=2^16 samples"),
    "C3": dict(dim=6, depth=8, hidden=36, variant="different_add", latent="Gauss", equation="advection_hamiltonian_wDiss",
               offset=[1.0, 0, 0, 1, 0, 0], n_samples=2 ** 18, num_params=8187, dt=1e-4,
               text="C3: 6D phase-space Fokker-Planck (advection_hamiltonian_wDiss), INN depth 8 x (36,) different_add, P=8187, "
                    "N=2^18 samples total"),
    "C4": dict(dim=10, depth=4, hidden=185, variant="no_add", latent="Gauss", equation="diffusion", offset=[0.0] * 10,
               n_samples=2 ** 20, num_params=16385, dt=1e-7,
               text="C4: 10D Fokker-Planck (diffusion), INN depth 4 x (185,), P=16385, N=2^20 samples total"),
}
C3 = CONFIGS["C3"]
METRIC = "TDVP steps/sec at N=2^18 samples, P~8k params (one step = one Heun step = 2 RHS)"
PARITY_FILE = os.path.join(ROOT, "tests", "golden", "bench_parity.json")


def metric_name(cfg_name):
    if cfg_name == "C3":
        return METRIC
    c = CONFIGS[cfg_name]
    return f"TDVP steps/sec at N=2^{int(np.log2(c['n_samples']))} samples, P={c['num_params']} params (one step = one Heun step = 2 RHS)"


def workload_config(n_gpus, cfg_name="C3"):
    c = CONFIGS[cfg_name]
    return {"workload": c["text"] + ", TDVP defaults (eigh solve, svdTol=1e-11), FixedStepper Heun",
            "n_samples": c["n_samples"], "num_params": c["num_params"], "dim": c["dim"], "rhs_per_step": 2,
            "parallelism": f"samples sharded over {n_gpus} GPU(s); tridiagonalisation + divide&conquer on one rank, overlapped with the "
                           "other ranks' SExp / C_EO Grams; back-transformation / SNR / update sharded over eigenvectors",
            "l2": "no explicit flush: each RHS streams the centred O matrix (17 GB at C3, >> 126 MB L2) three times"}


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
def build_ours(cfg_name="C3", **tdvp_args):
    from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper, net
    c = CONFIGS[cfg_name]
    off = np.asarray(c["offset"], dtype=np.float64)
    smp = sampler.Sampler(dim=c["dim"], numChains=30, name=c["latent"], mcmc_info={"offset": off, "bound": 0.25})
    net.SingleBlock.different_add = c["variant"] == "different_add"      # main.py:46-48: harmonicOsc uses "DifferentAdd"
    try:
        vs = var_state.VarState(smp, c["dim"], 1, c["depth"], network_args={"intmediate": (c["hidden"],), "offset": off,
                                                                        "latentSpaceName": c["latent"], "dim": c["dim"]})
    finally:
        net.SingleBlock.different_add = False
    assert vs.numParameters == c["num_params"]
    eq = evolutionEq.EvolutionEquation(dim=c["dim"], name=c["equation"])
    T = tdvp.TDVP(**tdvp_args)
    st = stepper.FixedStepper(timeStep=c["dt"], mode='Heun', maxStep=1e-2, increase_fac=1.3)   # main.py:51,113
    return vs, eq, T, st


class StageTimers:
    """CUDA-event timers around the C-ABI wrappers that make up the stages of a right-hand side (on the launching stream)."""
    NAMES = ("gram", "eigh", "eigh_cols", "eigh_factor", "eigh_backtransform")

    def __init__(self, P):
        self.P, self.events, self.orig = P, {k: [] for k in self.NAMES}, {}
        self.gram_flops = 0.0

    def __enter__(self):
        import torch
        from vmc_pde_b200 import _kernels
        for name in self.NAMES:
            fn = getattr(_kernels, name)
            self.orig[name] = fn

            def timed(*a, _fn=fn, _name=name):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _fn(*a)
                e1.record()
                self.events[_name].append((e0, e1))
                if _name == "gram":       # gram(O, n, ldo, Pp, weights, mats): SYRK convention, true P
                    self.gram_flops += len(a[5]) * a[1] * self.P * (self.P + 1.0)
            setattr(_kernels, name, timed)
        return self

    def __exit__(self, *exc):
        from vmc_pde_b200 import _kernels
        for name, fn in self.orig.items():
            setattr(_kernels, name, fn)

    def ms(self, *names):
        return sum(a.elapsed_time(b) for n in names for a, b in self.events[n])


def parity_probe(vs, eq, T, N):
    """One right-hand side from the fixed initial state: the numbers every run (any rank count) must reproduce."""
    upd, info = T(vs.get_parameters(), 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
    S0u = T.S0 @ upd                                   # checker arithmetic (torch), not the product path
    return {"ev_max": float(T.ev[-1]), "ev_sum": float(T.ev.sum()), "update_S0_update": float(upd @ S0u),
            "F_norm": float(T.F0.norm()), "tdvp_error": float(T.tdvp_error), "entropy": float(info["entropy"]),
            "solver_residual": float(T.solverResidual)}


def parity_report(cfg_name, values, write=False):
    ref = {}
    if os.path.exists(PARITY_FILE):
        ref = json.load(open(PARITY_FILE))
    if write:
        ref[cfg_name] = values
        json.dump(ref, open(PARITY_FILE, "w"), indent=1, sort_keys=True)
    if cfg_name not in ref:
        return {"checked_against": None, "values": values}
    tol = {"ev_max": 1e-10, "ev_sum": 1e-10, "update_S0_update": 1e-8, "F_norm": 1e-10, "tdvp_error": 1e-9, "entropy": 1e-12}
    dev = {k: abs(values[k] - ref[cfg_name][k]) / max(abs(ref[cfg_name][k]), 1e-300) for k in tol}
    dev["tdvp_error"] = abs(values["tdvp_error"] - ref[cfg_name]["tdvp_error"])     # 1 + (...) / <E^2>: an O(1) cancellation, absolute
    return {"checked_against": "tests/golden/bench_parity.json (committed values of the smallest run that holds the config: 1 GPU for C2 / C3, "
                               "2 GPUs with the replicated solve for C4; same seeds, any rank count)",
            "rel_dev": dev, "tolerance": tol, "ok": all(dev[k] <= tol[k] for k in tol) and values["solver_residual"] < 1e-8,
            "values": values}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vmc_pde_b200 import _kernels, _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n_gpus = world
    cfg_name = args.config
    cfg = CONFIGS[cfg_name]
    vs, eq, T, st = build_ours(cfg_name)
    P, N = vs.numParameters, cfg["n_samples"]
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
    key_init = vs.sampler.key.copy()
    # ---- parity: first right-hand side from the fixed initial state against the committed 1-GPU values
    parity = parity_report(cfg_name, parity_probe(vs, eq, T, N), write=args.write_parity and world == 1)
    vs.set_parameters(theta_init)
    vs.sampler.key = key_init.copy()

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
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _kernels.launches
    with StageTimers(P) as stages:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            info = step_device()
        ev1.record()
        barrier()
    launches = _kernels.launches - launches0
    ms = ev0.elapsed_time(ev1)
    n_rhs = 2 * args.steps
    gram_ms, gram_flops = stages.ms("gram"), stages.gram_flops
    serial_ms = stages.ms("eigh", "eigh_cols", "eigh_factor")
    back_ms = stages.ms("eigh_backtransform")
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
    variants = {}
    if not args.no_variants:
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
        from vmc_pde_b200 import tdvp as _tdvp, stepper as _stepper
        Tc = _tdvp.TDVP(diagonalShift=1e-4, solver="cholesky")
        vs.set_parameters(theta_init)      # timing variant: start again from the initial state with a fresh, small step
        stc = _stepper.FixedStepper(timeStep=cfg["dt"], mode='Heun', maxStep=1e-2, increase_fac=1.3)
        stc.step(0, Tc, vs.get_parameters(), **rhs)
        barrier()
        tc0 = torch.cuda.Event(enable_timing=True); tc1 = torch.cuda.Event(enable_timing=True)
        tc0.record()
        for _ in range(2):
            y, _, _ = stc.step(0, Tc, vs.get_parameters(), **rhs)
            vs.set_parameters(y)
        tc1.record()
        barrier()
        ms_chol = tc0.elapsed_time(tc1) / 2
        del Tc
        variants = {"lazy_SExp_ms": ms_lazy, "cholesky_ms": ms_chol}
        # third variant: the north star's "FP32-split path with stated tolerance": SExp and the SNR covariance on tcgen05
        # (bf16 x 3 split operands, TMEM accumulators); S0 stays FP64, theta_dot is bit-identical (tests).  Same step, same work.
        Ts = _tdvp.TDVP(gramPrecision="split")
        vs.set_parameters(theta_init)
        sts = _stepper.FixedStepper(timeStep=cfg["dt"], mode='Heun', maxStep=1e-2, increase_fac=1.3)
        sts.step(0, Ts, vs.get_parameters(), **rhs)
        barrier()
        split_ev = []
        orig_split = _kernels.gram_split

        def timed_split(*a_):
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record(); orig_split(*a_); e1_.record()
            split_ev.append((e0_, e1_, a_[1]))
        _kernels.gram_split = timed_split
        ts0 = torch.cuda.Event(enable_timing=True); ts1 = torch.cuda.Event(enable_timing=True)
        ts0.record()
        for _ in range(2):
            y, _, _ = sts.step(0, Ts, vs.get_parameters(), **rhs)
            vs.set_parameters(y)
        ts1.record()
        barrier()
        _kernels.gram_split = orig_split
        variants["split_ms"] = ts0.elapsed_time(ts1) / 2
        sp_ms = sum(a_.elapsed_time(b_) for a_, b_, _ in split_ev)
        Pp_ = _kernels.round_up(P, 128)
        tl_ = Pp_ // 128
        sp_flops = sum(tl_ * (tl_ + 1) / 2 * 128 * 128 * ((n_ + 127) // 128 * 128) * 2.0 * 6 for _, _, n_ in split_ev)
        variants["split_kernel"] = {"launches": len(split_ev), "ms_per_launch": sp_ms / max(len(split_ev), 1),
                                    "bf16_tflops": sp_flops / (sp_ms * 1e-3) * 1e-12 if sp_ms else None,
                                    "fp64_equivalent_tflops": sum(n_ * P * (P + 1.0) for _, _, n_ in split_ev) / (sp_ms * 1e-3) * 1e-12 if sp_ms else None}
        del Ts
        if args.adaptive:   # C3 with the adaptive integrator -- one attempt = 5 right-hand sides + the SExp error norm
            out_a = {}
            for mode in (True, "lazy"):
                vs.set_parameters(theta_init)
                Ta = _tdvp.TDVP(computeSExp=mode)
                ah = _stepper.AdaptiveHeun(timeStep=cfg["dt"], tol=1e-2, maxStep=1e-2)   # main.py:109-112
                barrier()
                ta0 = torch.cuda.Event(enable_timing=True); ta1 = torch.cuda.Event(enable_timing=True)
                ta0.record()
                y, dt_used, _ = ah.step(0, Ta, vs.get_parameters(), **rhs)
                ta1.record()
                barrier()
                out_a["eager_SExp" if mode is True else "lazy_SExp"] = {"seconds_per_accepted_step": ta0.elapsed_time(ta1) * 1e-3, "dt": float(dt_used)}
                del Ta
            variants["adaptive_heun"] = out_a
    # ---- max over ranks of the timings; per-rank stage times gathered
    mine = torch.tensor([ms, ms_e2e, variants.get("lazy_SExp_ms", 0.0), variants.get("cholesky_ms", 0.0), gram_ms, gram_flops, serial_ms,
                         back_ms, variants.get("split_ms", 0.0)], device="cuda", dtype=torch.float64)
    allv = [mine]
    if world > 1:
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    allv = torch.stack(allv).cpu().numpy()
    ms, ms_e2e, ms_lazy, ms_chol, ms_split = allv[:, 0].max(), allv[:, 1].max(), allv[:, 2].max(), allv[:, 3].max(), allv[:, 8].max()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # FP64 tensor peak: median of warm repetitions of the DMMA probe right after the timed steps, checked against the
    # datapath arithmetic 148 SMs x 64 FMA/clk x 2 x f_SM at the clock nvidia-smi reported during the timed region
    peak, peak_vals = _kernels.dmma_peak_tflops()
    sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    peak_arith = sms * 64 * 2 * sm_mhz * 1e6 * 1e-12
    peak_note = "measured (median of 9 warm DMMA-probe launches)"
    if abs(peak / peak_arith - 1) > 0.03:
        peak_note = (f"probe median {peak:.2f} TFLOP/s is more than 3 % off the datapath arithmetic {peak_arith:.2f} "
                     f"(SMs x 64 FMA/clk x 2 x {sm_mhz:.0f} MHz): the arithmetic value is used")
        peak = peak_arith
    value = args.steps / (ms * 1e-3)
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(prof) and cfg_name == "C3":
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            traffic_src = "profiles/gram_traffic.json: ncu capture of one 3-matrix launch at n=2^18, Pp=8192 on one GPU (not re-measured in this run)"
        except Exception:
            traffic = None
    g_ms, g_fl = allv[:, 4], allv[:, 5]
    achieved = float(g_fl.sum() / (g_ms.sum() * 1e-3) * 1e-12) if g_ms.sum() > 0 else None      # mean per-GPU rate of the kernel
    frac = (achieved / peak) if achieved and peak else None
    line = {
        "metric": metric_name(cfg_name), "value": value, "unit": "steps/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n_gpus, cfg_name),
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": P * 8, "d2h_bytes_per_step": (P + 3) * 8},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gram_kernel (3 weighted FP64 SYRKs per RHS: S0, SExp, SNR covariance)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": frac,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "FP64 DMMA.8x8x4 probe (vmcpde_dmma_probe), MEASURED_PEAKS.json has no FP64 entry; " + peak_note,
                     "peak_probe_values": [round(v, 3) for v in peak_vals], "peak_datapath_arithmetic": peak_arith,
                     "frac_flag": None if (frac is None or frac <= 1.05) else "frac > 1.05: the peak probe ran at a lower clock than the timed kernel; do not trust",
                     "flops_convention": "n_mats * n * P * (P+1) per launch (SYRK, SURVEY 8d), summed over ranks / summed kernel time",
                     "share_of_step": float(g_ms.max() / ms) if ms else None},
        "stages_ms_per_rhs": {"gram_max_over_ranks": float(g_ms.max() / n_rhs), "gram_by_rank": [float(v / n_rhs) for v in g_ms],
                              "eigh": float(allv[:, 6].max() / n_rhs + allv[:, 7].max() / n_rhs),
                              "eigh_serial_on_solver_rank": float(allv[:, 6].max() / n_rhs),
                              "eigh_backtransform_sharded": float(allv[:, 7].max() / n_rhs),
                              "rhs_total": float(ms / n_rhs),
                              "pipelined_solve": bool(getattr(T, "_last_partition", (False,))[0]),
                              "sample_partition": [n for _, n in T.sample_partition(N, world, P, getattr(T, "_last_partition", (None,))[0])]},
        "parity": parity,
        "last_entropy": ent,
    }
    if variants:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        sk = variants.get("split_kernel", {})
        line["variants"] = {"lazy_SExp_steps_per_s": 1e3 / ms_lazy, "cholesky_shift1e-4_steps_per_s": 1e3 / ms_chol,
                            "split_precision_steps_per_s": 1e3 / ms_split,
                            "split_precision": {"what": "TDVP(gramPrecision='split'): SExp and the SNR covariance on tcgen05 (bf16 x 3 split operands, "
                                                        "6 products per logical product, FP32 TMEM accumulation over 128 samples, FP64 sums; stated "
                                                        "tolerance 1e-6); S0 on FP64 DMMA; theta_dot bit-identical to the FP64 run",
                                                "kernel": "gram_split_kernel (rank 0)", **sk,
                                                "roofline": {"bound": "tensor", "achieved": sk.get("bf16_tflops"), "peak": bf16_peak, "unit": "TFLOP/s",
                                                             "frac": (sk.get("bf16_tflops") / bf16_peak) if sk.get("bf16_tflops") else None,
                                                             "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (else fallback 1400)"}},
                            "note": "TDVP(computeSExp='lazy'): SExp kept as a matrix-free operator on the resident O (2 Grams per RHS "
                                    "instead of 3); cholesky: TDVP(diagonalShift=1e-4, solver='cholesky') replaces the eigen-solve by the "
                                    "tensor-core blocked Cholesky (S0 and SExp Grams only).  Neither is the headline -- the reference forms "
                                    "SExp every call and regularises through the eigendecomposition"}
        if "adaptive_heun" in variants:
            line["variants"]["adaptive_heun"] = variants["adaptive_heun"]
    # CPU arm: rank 0 at N=1 only
    line["cpu_baseline"] = cpu_baseline(cfg_name) if (world == 1 and not args.no_cpu) else None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------- C5
def run_c5(args):
    """BASELINE configs[4]: synthetic S-build + solve sweep.  O rows iid N(0,1) x column scale 10^(-6 k/P) (a 12-decade spectrum),
    weights iid; per point: the 3-matrix Gram launch (TFLOP/s, SYRK convention, fraction of the DMMA peak), the eigen-solve and the
    shifted-Cholesky solve.  One GPU per process; with several ranks every rank runs the same sweep on its own sample shard and the
    Gram time is the max over ranks (the all-reduce is not part of the sweep)."""
    import torch
    from vmc_pde_b200 import _kernels, _lib
    _lib.require_cuda()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    Ps = [int(p) for p in args.c5_P.split(",")]
    Ns = [int(n) for n in args.c5_N.split(",")]
    free = torch.cuda.mem_get_info()[0]
    points = []
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for P in Ps:
        Pp = _kernels.round_up(P, 128)
        solve = {}
        for N in Ns:
            n = N // world
            if n * Pp * 8 + 6 * Pp * Pp * 8 > 0.8 * free:
                points.append({"P": P, "N": N, "skipped": "O shard does not fit in HBM"})
                continue
            g = torch.Generator(device="cuda"); g.manual_seed(P + N + rank)
            O = torch.randn(n, Pp, device="cuda", dtype=torch.float64, generator=g)
            O *= 10.0 ** (-6.0 * torch.arange(Pp, device="cuda", dtype=torch.float64) / P)
            O[:, P:] = 0
            w1, w2 = torch.rand(n, device="cuda", dtype=torch.float64), torch.rand(n, device="cuda", dtype=torch.float64)
            mats = [_kernels.zeros(Pp, Pp) for _ in range(3)]
            _kernels.gram(O, n, Pp, Pp, [None, w1, w2], mats)          # warm-up
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(args.c5_reps):
                _kernels.gram(O, n, Pp, Pp, [None, w1, w2], mats)
            e1.record(); e1.synchronize()
            g_ms = e0.elapsed_time(e1) / args.c5_reps
            if world > 1:
                t = torch.tensor([g_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                g_ms = float(t[0])
            pt = {"P": P, "N": N, "gram_ms": g_ms, "gram_tflops_per_gpu": 3.0 * n * P * (P + 1.0) / (g_ms * 1e-3) * 1e-12}
            if not solve:       # the solve depends on P only: once per P, on the S of this N
                S = mats[0].clone()
                _kernels.sym_finalize(S, Pp, 1.0 / (n * (args.c5_reps + 1)))
                F = torch.randn(Pp, device="cuda", dtype=torch.float64, generator=g); F[P:] = 0
                if P <= 25 * 1024:
                    work, evv, VT = S.clone(), _kernels.zeros(Pp), _kernels.zeros(Pp, Pp)
                    ws = _kernels.workspace(_kernels.eigh_workspace_bytes(P, Pp))
                    _kernels.eigh(work, P, Pp, evv, VT, ws)
                    work.copy_(S)
                    e0, e1 = ev(), ev()
                    e0.record(); _kernels.eigh(work, P, Pp, evv, VT, ws); e1.record(); e1.synchronize()
                    solve["eigh_ms"] = e0.elapsed_time(e1)
                    del work, VT
                else:
                    solve["eigh_ms"] = None
                    solve["eigh_note"] = "vmcpde_eigh supports P <= 25600: the top row of the sweep runs the shifted Cholesky only"
                Ssh, upd, info = _kernels.empty(Pp, Pp), _kernels.zeros(Pp), torch.zeros(1, dtype=torch.int32, device="cuda")
                for rep in range(2):
                    _kernels.diag_shift(S, Ssh, Pp, P, 1e-4)
                    e0, e1 = ev(), ev()
                    e0.record(); _kernels.chol_solve(Ssh, P, Pp, F, upd, info); e1.record(); e1.synchronize()
                solve["cholesky_ms"] = e0.elapsed_time(e1)
                solve["cholesky_ok"] = int(info.item()) == 0
                del S, Ssh
            pt.update(solve)
            points.append(pt)
            del O, mats
            torch.cuda.empty_cache()
    if rank == 0:
        peak, vals = _kernels.dmma_peak_tflops()
        for pt in points:
            if "gram_tflops_per_gpu" in pt:
                pt["frac_of_dmma_peak"] = pt["gram_tflops_per_gpu"] / peak
        best = max((p for p in points if "gram_tflops_per_gpu" in p), key=lambda p: p["N"] * p["P"])
        line = {"metric": "S-build TFLOP/s per GPU (3 weighted FP64 SYRKs, SYRK flop convention)", "value": best["gram_tflops_per_gpu"],
                "unit": "TFLOP/s", "n_gpus": world, "steps": args.c5_reps, "warmup": 1, "ms_per_step": best["gram_ms"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C5: synthetic S-build + solve sweep (graded O, 12 decades)", "P": Ps, "N": Ns},
                "roofline": {"bound": "tensor", "achieved": best["gram_tflops_per_gpu"], "peak": peak, "unit": "TFLOP/s",
                             "frac": best["gram_tflops_per_gpu"] / peak, "traffic": None}, "sweep": points}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
_cpu_state = {}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1: give torch AND NumPy's BLAS every core this process may use."""
    import torch
    n = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    torch.set_num_threads(n)
    try:
        from threadpoolctl import threadpool_limits
        _cpu_state["tp"] = threadpool_limits(limits=n)       # kept alive for the rest of the process
    except Exception:
        pass
    return n


def cpu_state(cfg_name):
    from oracle import flow
    key = ("st", cfg_name)
    if key not in _cpu_state:
        c = CONFIGS[cfg_name]
        ups, downs, k = flow.make_index_splits(c["dim"], c["depth"], 1)
        spec = flow.FlowSpec(dim=c["dim"], depth=c["depth"], hidden=(c["hidden"],), variant=c["variant"], latent=c["latent"],
                             offset=np.asarray(c["offset"]), inds_up=ups, inds_down=downs)
        _cpu_state[key] = (flow.OracleState(spec, flow.init_params_flax(spec, k)), spec.num_params)
    return _cpu_state[key]


def cpu_linear_sections(cfg_name, n_s):
    """The N-linear sections of one right-hand side of the CPU restatement (oracle/: the reference's algorithm incl. its
    two Gram products tdvp.py:46-47) on n_s samples.  Returns (seconds by section, S0, EO)."""
    from oracle import tdvp
    st, P = cpu_state(cfg_name)
    c = CONFIGS[cfg_name]
    t0 = time.perf_counter()
    x, lp, _ = st.sample(n_s)
    t1 = time.perf_counter()
    E, O, lp2, _ = tdvp.local_terms(st, x, c["equation"], 0.0)
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
    return {"sampling_s": t1 - t0, "local_terms_s": t2 - t1, "two_grams_and_F_s": t3 - t2}, S0, EO


def cpu_solve_sections(S0, EO, p_block=None):
    """The P-only section (host eigh, tdvp.py:61-64) at full P -- or on a p_block x p_block block scaled cubically when P is
    beyond what the host finishes in a minute (P > 10000) or a block is asked for -- and EOdata @ V (tdvp.py:68; linear in N,
    quadratic in P) on the given samples.  Returns (eigh seconds at full P, EO@V seconds on these samples at full P, note)."""
    P = S0.shape[0]
    p_s = min(P, p_block) if p_block else (P if P <= 10000 else 4096)
    t0 = time.perf_counter()
    ev, V = np.linalg.eigh(S0[:p_s, :p_s])
    t1 = time.perf_counter()
    EOv = EO[:, :p_s] @ V
    t2 = time.perf_counter()
    return (t1 - t0) * (P / p_s) ** 3, (t2 - t1) * (P / p_s) ** 2, ("full P" if p_s == P else f"a {p_s} block scaled (P/{p_s})^3")


def cpu_step_estimate(cfg_name="C3", n_s=2048, p_block=None):
    """One Heun step (2 right-hand sides) of the CPU restatement, seconds: the N-linear sections measured on n_s samples and
    scaled by N / n_s, the eigh measured ONCE per process (cached; full P unless p_block), EOdata@V on the same samples."""
    N = CONFIGS[cfg_name]["n_samples"]
    parts, S0, EO = cpu_linear_sections(cfg_name, n_s)
    key = ("solve", cfg_name, p_block)
    if key not in _cpu_state:
        eigh_s, eov_s, note = cpu_solve_sections(S0, EO, p_block)
        _cpu_state[key] = (eigh_s, eov_s / n_s, note)
    eigh_s, eov_per_sample, note = _cpu_state[key]
    lin = N / n_s
    full = {k: v * lin for k, v in parts.items()}
    full["eigh_s"] = eigh_s
    full["EO_at_V_s"] = eov_per_sample * N
    return 2.0 * sum(full.values()), full, note


def cpu_linearity(cfg_name, sizes=(2 ** 14, 2 ** 15)):
    """the N-linear sections at two sizes; returns seconds per sample at each size."""
    out = {}
    for n in sizes:
        parts, _, _ = cpu_linear_sections(cfg_name, n)
        out[str(n)] = {k: v / n for k, v in parts.items()}
        out[str(n)]["total_per_sample_s"] = sum(parts.values()) / n
    return out


def cpu_baseline(cfg_name="C3", n_s=2048):
    cores = use_all_host_threads()
    cpu_linear_sections(cfg_name, 256)                     # warm-up (torch.func tracing)
    sec, parts, note = cpu_step_estimate(cfg_name, n_s)
    N = CONFIGS[cfg_name]["n_samples"]
    return {"value": 1.0 / sec, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ restatement (torch float64 CPU + NumPy BLAS, not JAX): sampling, local terms, 2 Grams + F measured on {n_s} of "
                      f"{N} samples and scaled linearly; np.linalg.eigh measured once at {note}; EOdata@V measured on the same samples at "
                      "full P; seconds per RHS at full size: " + ", ".join(f"{k}={v:.0f}" for k, v in parts.items())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config if args.config in CONFIGS else "C3"
    cores = use_all_host_threads()
    cpu_linear_sections(cfg_name, 256)
    linearity = None
    if not os.environ.get("VMCPDE_BENCH_QUICK"):
        linearity = cpu_linearity(cfg_name)                # N = 2^14 and 2^15, once
    for _ in range(args.warmup):
        cpu_step_estimate(cfg_name)
    secs = []
    for _ in range(args.steps):
        s, parts, note = cpu_step_estimate(cfg_name)
        secs.append(s)
    sec = sum(secs) / len(secs)
    value = 1.0 / sec
    N = CONFIGS[cfg_name]["n_samples"]
    line = {"impl": "reference", "metric": metric_name(cfg_name), "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, cfg_name),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"each step: sampling, local terms, 2 Grams + F on 2048 of {N} samples scaled linearly in N (linearity of "
                                       f"those sections measured at N = 2^14 and 2^15, see `linearity`), np.linalg.eigh at {note} measured once, "
                                       "EOdata@V at full P on the same samples; reference = oracle/ port (torch CPU + NumPy BLAS on all host "
                                       "cores), JAX is not installable in this image",
                             "seconds_per_rhs_full_size": parts, "linearity_seconds_per_sample": linearity},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--adaptive", action="store_true",
                    help="also time AdaptiveHeun attempts (5 RHS each, stepper.py:54-66) on the same workload; adds ~30 s")
    ap.add_argument("--no-variants", action="store_true", help="skip the lazy-SExp / Cholesky variant timings")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--write-parity", action="store_true", help="(1 GPU) store the parity values of this run as the committed reference")
    ap.add_argument("--c5-P", default="2048,4096,8192,16384,32768")
    ap.add_argument("--c5-N", default="65536,262144,1048576")
    ap.add_argument("--c5-reps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
