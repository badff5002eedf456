"""CPU: host-side logic of the mirror modules (steppers, sharding arithmetic, reductions over gloo with 2 ranks)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import tdvp as otdvp
from vmc_pde_b200 import stepper, mpi_wrapper, util


class MockRHS:
    """y' = A y with a fixed SExp, the interface stepper.py expects from TDVP (callable + .SExp)."""

    def __init__(self, A):
        self.A, self.SExp, self.calls = A, torch.eye(A.shape[0], dtype=torch.float64), []

    def __call__(self, y, t, **kw):
        self.calls.append(kw["intStep"])
        return self.A @ y, {"k": kw["intStep"]}


def test_fixed_stepper_matches_reference_semantics():
    A = torch.tensor([[-1.0, 0.3], [0.0, -2.0]], dtype=torch.float64)
    y0 = torch.tensor([1.0, 2.0], dtype=torch.float64)
    f = MockRHS(A)
    st = stepper.FixedStepper(timeStep=1e-2, mode='Heun', maxStep=1.0, increase_fac=1.3)
    y, dt, info = st.step(0, f, y0, extra=1)
    yo, dto = otdvp.heun_step(lambda y, k: A.numpy() @ y, y0.numpy(), 1e-2, 1.0, 1.3)
    assert abs(dt - dto) < 1e-16 and np.allclose(y.numpy(), yo, atol=1e-15)
    assert f.calls == [0, 1] and info == {"k": 1}          # stepper.py:135-137: info comes from the second call
    assert torch.equal(y0, torch.tensor([1.0, 2.0], dtype=torch.float64))  # input untouched
    st = stepper.FixedStepper(timeStep=1e-2, mode='Euler', maxStep=1.5e-2, increase_fac=2.0)
    y, dt, _ = st.step(0, MockRHS(A), y0)
    assert abs(dt - 1.5e-2) < 1e-16 and np.allclose(y.numpy(), (y0 + dt * (A @ y0)).numpy())


def test_adaptive_heun_matches_reference_semantics():
    A = torch.tensor([[-3.0, 0.5], [0.2, -1.0]], dtype=torch.float64)
    y0 = torch.tensor([1.0, -1.0], dtype=torch.float64)
    f = MockRHS(A)
    norm = lambda v, S: v @ S @ v   # main.py:24-26 (a quadratic form, no square root)
    ah = stepper.AdaptiveHeun(timeStep=0.2, tol=1e-6, maxStep=1.0)
    y, rdt, info = ah.step(0, f, y0, normFunction=norm)
    yo, rdto, ndto = otdvp.adaptive_heun_step(lambda y, k: A.numpy() @ y, y0.numpy(), 0.2, 1e-6, 1.0, lambda v: float(v @ v))
    assert abs(rdt - rdto) < 1e-15 and abs(ah.dt - ndto) < 1e-15 and np.allclose(y.numpy(), yo, atol=1e-14)
    assert len(f.calls) % 5 == 0 and f.calls[:5] == [0, 1, 2, 3, 4] and info == {"k": 0}


def test_distribute_sampling_arithmetic():
    """mpi_wrapper.py:68-126 on one process."""
    assert mpi_wrapper.distribute_sampling(1000) == 1000 and mpi_wrapper.globNumSamples == 1000
    assert mpi_wrapper.distribute_sampling(1000, localDevices=1, numChainsPerDevice=30) == 34
    assert mpi_wrapper.globNumSamples == 34 * 30
    assert mpi_wrapper.first_sample_id() == 0
    assert mpi_wrapper.shard_range(10) == (0, 10)
    assert mpi_wrapper.rank == 0 and mpi_wrapper.commSize == 1


def test_grid_geometry_matches_reference_layout():
    """grid.py:7-29: widths, cell area, origin list in meshgrid('xy') order, symmetric and one-sided variants."""
    from vmc_pde_b200 import grid
    g = grid.Grid(np.array([2.0, 3.0]), 4, sym=True)
    assert g.dim == 2 and np.allclose(g.widths, [1.0, 1.5]) and abs(g.bin_area - 1.5) < 1e-15
    assert g.range == [[-2.0, 2.0], [-3.0, 3.0]] and g.coords.shape == (16, 2)
    xs, ys = np.arange(-2, 2, 1.0), np.arange(-3, 3, 1.5)
    ref = np.moveaxis(np.array(np.meshgrid(xs, ys)), 0, -1).reshape(16, 2)     # the reference's construction
    assert np.array_equal(g.coords, ref)
    g1 = grid.Grid(np.array([2.0, 2.0]), 5, sym=False)
    assert np.allclose(g1.widths, 0.4) and g1.range == [[0.0, 2.0], [0.0, 2.0]] and g1.coords.min() == 0.0
    # a unit Gaussian integrates to ~1 on a wide symmetric grid (midpoint-free Riemann sum as in var_state.py:88-91)
    g2 = grid.Grid(np.array([8.0, 8.0]), 200)
    dens = np.exp(-0.5 * (g2.coords ** 2).sum(-1)) / (2 * np.pi)
    assert abs(g2.bin_area * dens.sum() - 1.0) < 1e-6


def test_timings_and_cov_matrix():
    t = util.Timings()
    t.start_timing("a"); t.stop_timing("a")
    assert t.timing_dict["a"][-1] >= 0
    S = util.build_cov_matrix(torch.tensor([0.5, -0.2, 0.1], dtype=torch.float64), torch.tensor([0.1, 0.0, -0.3], dtype=torch.float64), 3)
    L = np.array([[np.exp(0.1), 0.5, -0.2], [0, 1.0, 0.1], [0, 0, np.exp(-0.3)]])
    assert np.allclose(S.numpy(), L @ L.T)


def test_store_infos_writes_the_flat_hdf5_group_of_the_reference(tmp_path):
    """util.py:29-32 / main.py:205: every key of `infos` (a list of per-step values) becomes one dataset of shape
    (steps, ...) in infos.hdf5 -- the layout the reference's plotting scripts and its stored runs use."""
    from vmc_pde_b200 import _hdf5
    rng = np.random.default_rng(3)
    steps, P, d = 5, 7, 2
    infos = {"times": [0.1 * i for i in range(steps)], "ev": [torch.tensor(rng.normal(size=P)) for _ in range(steps)],
             "covar": [torch.tensor(rng.normal(size=(d, d))) for _ in range(steps)], "entropy": [torch.tensor(1.5 + i) for i in range(steps)],
             "integral_0.5sigma": [np.float64(0.3)] * steps, "dist_params": [torch.zeros(0, dtype=torch.float64)] * steps}
    wdir = str(tmp_path) + "/"
    util.store_infos(wdir, infos)
    back = _hdf5.read(wdir + "infos.hdf5")
    assert set(back) == set(infos) and back["ev"].shape == (steps, P) and back["covar"].shape == (steps, d, d)
    assert back["dist_params"].shape == (steps, 0) and back["entropy"].shape == (steps,)       # as in the stored runs
    assert np.array_equal(back["ev"][3], infos["ev"][3].numpy()) and np.allclose(back["times"], infos["times"])
    again = util.load_infos(wdir)
    assert all(np.array_equal(again[k], back[k]) for k in back)
    with pytest.raises(TypeError):
        util.store_infos(wdir, {"snr": [None, None]})


def test_solve_shard_ranges_partition_the_eigenvectors():
    """tdvp.solve_shard_range: contiguous 128-blocks, every eigenvector exactly once, and the per-slice partial updates
    plus zero-padded slice vectors sum to the unsharded solve (the all-reduce pattern of TDVP._finish)."""
    from vmc_pde_b200.tdvp import solve_shard_range
    for Pp, R in ((128, 1), (256, 2), (1024, 3), (8192, 8), (16384, 5), (2176, 8)):
        covered = np.zeros(Pp, dtype=int)
        for r in range(R):
            row0, nrows = solve_shard_range(Pp, R, r)
            assert row0 % 128 == 0 and nrows % 128 == 0 and nrows > 0
            covered[row0:row0 + nrows] += 1
        assert (covered == 1).all()
    rng = np.random.default_rng(5)
    P, Pp, R = 300, 384, 3
    A = rng.normal(size=(P + 40, P)); S = A.T @ A / (P + 40); F = rng.normal(size=P)
    ev, V = np.linalg.eigh(S)
    coef = (V.T @ F) / ev
    full = V @ coef
    acc, vt = np.zeros(P), np.zeros(Pp)
    for r in range(R):
        row0, nrows = solve_shard_range(Pp, R, r)
        k0, k1 = row0, min(P, row0 + nrows)
        acc += V[:, k0:k1] @ coef[k0:k1]
        vt[k0:k1] += (V.T @ F)[k0:k1]
    assert np.allclose(acc, full, rtol=1e-12, atol=1e-12) and np.allclose(vt[:P], V.T @ F)


def test_sample_partition_of_the_pipelined_solve():
    """TDVP.sample_partition: contiguous, complete, deterministic; equal shards unless the solve is pipelined, then the
    solver rank's share shrinks with the cost of the serial eigensolver stages (none at C3 on 4 and 8 ranks)."""
    from vmc_pde_b200.tdvp import TDVP, solver_seconds
    T = TDVP()
    for N, R, P in ((2 ** 18, 1, 8187), (2 ** 18, 2, 8187), (2 ** 18, 4, 8187), (2 ** 18, 8, 8187), (2 ** 20, 8, 16385), (10007, 3, 2053), (1000, 2, 37)):
        part = T.sample_partition(N, R, P)
        assert len(part) == R and part[0][0] == 0 and sum(n for _, n in part) == N
        assert all(part[i][0] + part[i][1] == part[i + 1][0] for i in range(R - 1)) and part == T.sample_partition(N, R, P)
        if R > 1 and P >= 384:
            others = [n for r, (_, n) in enumerate(part) if r != 0]
            assert part[0][1] <= min(others) and max(others) - min(others) <= 1 and part[0][1] % 16 == 0
    assert [n for _, n in T.sample_partition(1000, 2, 37)] == [500, 500]                       # unblocked solver: equal shards
    assert 0 < T.sample_partition(2 ** 18, 8, 8187)[0][1] < 2 ** 18 // 16 and T.sample_partition(2 ** 18, 4, 8187)[0][1] < 2 ** 18 // 8
    n0 = T.sample_partition(2 ** 18, 2, 8187)[0][1]
    assert 0.25 * 2 ** 18 < n0 < 0.45 * 2 ** 18
    assert [n for _, n in TDVP(pipelineSolve=False).sample_partition(2 ** 18, 4, 8187)] == [2 ** 16] * 4
    assert [n for _, n in TDVP(solverShare=0.5).sample_partition(2 ** 18, 4, 8187)][0] == 2 ** 15
    assert solver_seconds(8187) == pytest.approx(0.33) and solver_seconds(4096) < solver_seconds(5000) < solver_seconds(8187)


def test_sample_partition_invariants_hold_for_any_size():
    """Property test: whatever (N, ranks, P, solver rank, Gram precision, SExp mode), the shares tile [0, N) in rank order, are
    a pure function of their arguments, the solver rank's share is 16-aligned (TMA box rows) and never the largest, and the
    other ranks differ by at most one sample."""
    from hypothesis import given, settings, strategies as st
    from vmc_pde_b200.tdvp import TDVP, solve_shard_range

    @settings(max_examples=300, deadline=None)
    @given(N=st.integers(1, 2 ** 22), R=st.integers(1, 8), P=st.integers(1, 26000), srank=st.integers(0, 7),
           split=st.booleans(), sexp=st.sampled_from([True, "lazy", False]), share=st.sampled_from([None, 0.0, 0.3, 1.0]))
    def check(N, R, P, srank, split, sexp, share):
        T = TDVP(solverRank=srank % R, gramPrecision="split" if split else "fp64", computeSExp=sexp, solverShare=share)
        part = T.sample_partition(N, R, P)
        assert part == T.sample_partition(N, R, P) and len(part) == R
        assert part[0][0] == 0 and all(n >= 0 for _, n in part) and sum(n for _, n in part) == N
        assert all(part[i][0] + part[i][1] == part[i + 1][0] for i in range(R - 1))
        Pp = (P + 127) // 128 * 128
        if T._pipelined(R, P, Pp):
            n0 = part[T.solverRank][1]
            others = [n for r, (_, n) in enumerate(part) if r != T.solverRank]
            assert n0 % 16 == 0 or n0 == N
            assert max(others) - min(others) <= 1 and (share is not None or n0 <= max(others))
        else:
            assert max(n for _, n in part) - min(n for _, n in part) <= 1
        # eigenvector slices of the sharded solve: 128-aligned, disjoint, complete
        rows = [solve_shard_range(Pp, R, r) for r in range(R)]
        assert all(a % 128 == 0 and b % 128 == 0 for a, b in rows) and sum(b for _, b in rows) == Pp
        assert all(rows[i][0] + rows[i][1] == rows[i + 1][0] for i in range(R - 1))
    check()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vmc_pde_b200 import mpi_wrapper as mpi
    N = 11
    first, n = mpi.shard_range(N)
    rng = np.random.default_rng(0)
    full = torch.tensor(rng.normal(size=(1, N, 3)))
    mine = full[:, first:first + n]
    mpi.globNumSamples = N
    out = dict(rank=mpi.rank, size=mpi.commSize, first=first, n=n,
               sum=mpi.global_sum(mine).numpy(), mean=mpi.global_mean(mine).numpy(), var=mpi.global_variance(mine).numpy(),
               spp=mpi.distribute_sampling(N), fsid=mpi.first_sample_id(),
               bcast=mpi.bcast_unknown_size(np.arange(4, dtype=np.float64) if rank == 0 else None))
    # the packed two-collective pattern of tdvp.py: first moments, then second moments about the global mean
    packed = torch.cat([mine.sum(dim=(0, 1)), (mine ** 2).sum(dim=(0, 1))])
    mpi.allreduce_(packed)
    out["packed"] = packed.numpy()
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reductions_over_gloo():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r["rank"])
    [p.join(timeout=60) for p in procs]
    full = np.random.default_rng(0).normal(size=(11, 3))
    assert [r["size"] for r in res] == [2, 2] and [(r["first"], r["n"]) for r in res] == [(0, 6), (6, 5)]
    for r in res:
        assert np.allclose(r["sum"], full.sum(0)) and np.allclose(r["mean"], full.mean(0)) and np.allclose(r["var"], full.var(0))
        assert np.allclose(r["packed"], np.concatenate([full.sum(0), (full ** 2).sum(0)]))
        assert np.array_equal(r["bcast"], np.arange(4.0))
    assert [r["spp"] for r in res] == [6, 5] and [r["fsid"] for r in res] == [0, 6]


def test_bench_reference_arm_runs_on_the_cpu():
    """bench.py --impl reference (the oracle port timed on the host cores): a reduced sample goes through every section
    (sampling, local terms, Grams, eigh, EO@V) and extrapolates to a positive step time; all host threads are claimed even
    under torchrun's OMP_NUM_THREADS=1."""
    import importlib.util, pathlib
    spec = importlib.util.spec_from_file_location("bench", pathlib.Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    cores = bench.use_all_host_threads()
    assert cores == len(os.sched_getaffinity(0)) == torch.get_num_threads()
    sec, parts, note = bench.cpu_step_estimate("C3", n_s=64, p_block=256)
    assert sec > 0 and set(parts) == {"sampling_s", "local_terms_s", "two_grams_and_F_s", "eigh_s", "EO_at_V_s"}
    assert all(v >= 0 for v in parts.values()) and "256 block" in note
    lin = bench.cpu_linearity("C2", sizes=(64, 128))
    assert set(lin) == {"64", "128"} and lin["64"]["total_per_sample_s"] > 0
    cfg = bench.workload_config(4)
    assert cfg["num_params"] == 8187 and cfg["n_samples"] == 2 ** 18 and "workload" in cfg
    assert bench.workload_config(2, "C4")["num_params"] == 16385 and bench.metric_name("C3") == bench.METRIC
    rep = bench.parity_report("C3", {"ev_max": 1.0, "ev_sum": 2.0, "update_S0_update": 3.0, "F_norm": 1.0, "tdvp_error": 0.5,
                                     "entropy": 1.0, "solver_residual": 1e-10})
    assert "values" in rep


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines kept under profiles/ (what DESIGN.md quotes) have every key of the bench.py contract, the metric and
    config of BASELINE.json, and self-consistent numbers (value = 1 / step time; frac = achieved / peak; Gram time < step time)."""
    import json, pathlib
    root = pathlib.Path(__file__).resolve().parents[1]
    base = json.loads((root / "BASELINE.json").read_text())
    for name, n in (("r02b_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)):
        txt = (root / "profiles" / name).read_text()
        d = json.loads(txt[txt.index("{"):])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "roofline", "e2e", "clocks", "gpu_launches", "parity"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
        assert d["unit"] == "steps/s" and "workload" in d["config"] and d["config"]["num_params"] == 8187 and d["config"]["n_samples"] == 2 ** 18
        assert str(base.get("metric", "")).split()[0].lower() in d["metric"].lower()
        assert abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-9 and d["warmup"] >= 3
        r = d["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] <= 1.05
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert d["gpu_launches"] > 0 and d["parity"]["ok"] is True
        assert 2 * d["stages_ms_per_rhs"]["gram_max_over_ranks"] < d["ms_per_step"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            c = d["cpu_baseline"]
            assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] == "port" and c["value"] > 0
