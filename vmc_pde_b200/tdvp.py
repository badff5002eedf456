"""Mirror of vmc_fluids/tdvp.py: the TDVP right-hand side  theta_dot = S^-1 F  with the reference's regularised
eigen-solve, computed entirely on the GPU.

`TDVP.__call__(netParameters, t, psi=, evolutionEq=, nSamplesTDVP=, nSamplesObs=, timings=, ...)` keeps the
reference's signature, required keys and post-call attributes (tdvp.py:96-164).  What changes is
how the numbers are produced:

  * samples are processed in chunks; the N x P matrix O is never required to exist as a whole and the reference's
    other N x P temporaries (centred O, E*O, logp*O, EO@V; tdvp.py:30-34,40-47,68) never exist at all;
  * two passes: first moments (one packed all-reduce), then centring + force vector + the three weighted Grams
    S0, SExp, C_EO on FP64 tensor cores (one packed all-reduce) -- the reference issues ten host-staged Allreduces;
  * the signal-to-noise ratio uses diag(V^T C_EO V) instead of the N x P x P product EOdata @ V (tdvp.py:68-70);
  * eigendecomposition, cut-offs, update, residual and TDVP error run on the device (no host LAPACK round trip).
"""
from dataclasses import dataclass
import math
import time
import numpy as np
import scipy.special
import torch

from . import _kernels, mpi_wrapper as mpi, global_defs


class LazyGram:
    """SExp = (1/N) sum_i logp_i^2 dO_i dO_i^T (tdvp.py:47) as an operator on the centred O that is still resident.

    The reference reads SExp only through `normFunction(v, SExp)` (stepper.py:71; main.py:24-26: v^T S v), which needs
    S.dot(v) -- two streaming passes over O (`vmcpde_gram_matvec`) instead of the N P^2 flops of the matrix.  `S.dot(v)`,
    `S @ v` and `v @ S` are matrix free; anything that needs the entries (`materialize()`, indexing, `.cpu()`) builds the
    P x P matrix once with the Gram kernel and caches it.  Valid until the owning TDVP evaluates its next right-hand side
    (the O buffer is reused); using it later raises.  With several ranks `dot` and `materialize` are collective."""

    def __init__(self, owner, gen, O, n, n_pad, Pp, P, w, N):
        self._owner, self._gen, self._O, self._n, self._n_pad = owner, gen, O, n, n_pad
        self._Pp, self._P, self._w, self._N = Pp, P, w, N
        self._t = None
        self._mat = None
        self.shape = (P, P)
        self.dtype, self.device = O.dtype, O.device

    def _alive(self):
        if self._gen is not None and self._owner._gen != self._gen:
            raise RuntimeError("this SExp refers to the samples of an earlier right-hand side (the O buffer has been reused); "
                               "read it before the next TDVP call or use TDVP(computeSExp=True)")

    def dot(self, v):
        if self._mat is not None:
            return self._mat @ _kernels.as_dev(v)
        self._alive()
        vp = _kernels.zeros(self._Pp)
        vp[:self._P] = _kernels.as_dev(v).reshape(-1)
        out = _kernels.zeros(self._Pp)
        if self._t is None:
            self._t = _kernels.empty(max(self._n, 1))
        if self._n > 0:      # a solver rank of a pipelined solve may own no samples
            _kernels.gram_matvec(self._O, self._n, self._Pp, self._w, vp, self._t, out)
        mpi.allreduce_(out)
        return out[:self._P] / self._N

    __matmul__ = dot

    def __rmatmul__(self, v):  # v @ S: S is symmetric
        return self.dot(v)

    @property
    def T(self):
        return self

    def materialize(self):
        if self._mat is None:
            self._alive()
            M = _kernels.zeros(self._Pp, self._Pp)
            if self._n_pad > self._n:
                self._w[self._n:self._n_pad].zero_()
            _kernels.gram(self._O, self._n_pad, self._Pp, self._Pp, [self._w], [M])
            mpi.allreduce_(M)
            _kernels.sym_finalize(M, self._Pp, 1.0 / self._N)
            self._mat = M[:self._P, :self._P]
        return self._mat

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def cpu(self):
        return self.materialize().cpu()

    def __array__(self, dtype=None):
        a = self.materialize().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


# seconds of the serial eigensolver stages (tridiagonalisation + divide & conquer) on one B200 by matrix size, measured with
# bench.py / tools/probe_eigh.py (profiles/); log-log interpolation in between
_SOLVER_SECONDS = ((384, 0.006), (1024, 0.015), (2053, 0.030), (4096, 0.080), (8187, 0.33), (12000, 0.85), (16385, 2.8), (25600, 10.0))


def solver_seconds(P):
    pts = _SOLVER_SECONDS
    if P <= pts[0][0]:
        return pts[0][1]
    for (p0, t0), (p1, t1) in zip(pts, pts[1:]):
        if P <= p1:
            a = math.log(P / p0) / math.log(p1 / p0)
            return math.exp(math.log(t0) + a * (math.log(t1) - math.log(t0)))
    return pts[-1][1] * (P / pts[-1][0]) ** 3


def solve_shard_range(Pp, world, rank):
    """Eigenvector slice [row0, row0 + nrows) of rank `rank` in a sharded solve: the Pp / 128 blocks of 128 eigenvectors
    are dealt contiguously; ranks beyond the block count get an empty slice."""
    nblk = Pp // 128
    b0, b1 = rank * nblk // world, (rank + 1) * nblk // world
    return 128 * b0, 128 * (b1 - b0)


@dataclass
class TDVP:
    useSNR: bool = False
    snrTol: float = 2e0
    svdTol: float = 1e-11
    diagonalShift: float = 0
    diagonalizeOnDevice: bool = False   # kept for signature parity; the solve is always on the device
    # ---- extensions (not in the reference; defaults reproduce the reference's semantics) ----
    solver: str = "eigh"               # "eigh" | "cholesky" (needs diagonalShift > 0; no ev / V / snr)
    computeSExp: object = True         # True: build SExp every call (the reference, tdvp.py:47); "lazy": SExp is an
                                       # operator on the resident O (LazyGram; stepper.py:71 only needs v^T SExp v); False: skip
    computeSNR: bool = True            # snr is logged by main.py:187 and gates the solve when useSNR
    shardSolve: bool = True            # several ranks: shard the post-tridiagonal O(P^3) stages over eigenvectors
    pipelineSolve: bool = True         # several ranks: one rank runs the serial eigensolver stages while the others build
                                       # the SExp / C_EO Grams (sample shares balanced by sample_partition)
    gramPrecision: str = "fp64"        # "fp64": all three Grams on the FP64 DMMA pipe (the reference's arithmetic);
                                       # "split": SExp and C_EO on tcgen05 with bf16 x 3 split operands (stated tolerance 1e-6;
                                       # S0, F and hence theta_dot are unchanged when useSNR is off)
    solverRank: int = 0
    solverShare: object = None         # None: cost model; else fraction of an equal sample share given to the solver rank
    chunkSamples: int = 0              # samples per chunk (0 = choose from free memory)
    memoryFraction: float = 0.45       # share of free device memory the O buffer may take

    def __post_init__(self):
        if self.solver not in ("eigh", "cholesky"):
            raise ValueError("solver must be 'eigh' or 'cholesky'")
        if self.gramPrecision not in ("fp64", "split"):
            raise ValueError("gramPrecision must be 'fp64' or 'split'")
        self._bufP = None
        self._P = None
        self._vt_range = None
        self._gen = 0              # bumped whenever the O buffer is about to be overwritten
        self._lazy_ok = False      # the whole rank-local O is resident after pass 2 (one chunk)
        self._lazy_src = self._lazy_gen = None
        self.S = self.S0 = self.F0 = self.SExp = self.ev = self.VtF = None
        self.rhoVar = self.snr = self.invEv = None
        self.solverResidual = self.tdvp_error = None
        self.ElocMean = self.ElocMeanAbs = self.ElocVar = None

    @property
    def V(self):
        """Eigenvectors as columns (tdvp.py:59-64).  With a sharded solve every rank holds a slice of eigenvectors;
        the full matrix is assembled (one all-reduce of the zero-padded P x P buffer) on first access only."""
        if self._P is None:
            return None
        if self._vt_range is not None:
            row0, nrows = self._vt_range
            self._VT[:row0].zero_()
            self._VT[row0 + nrows:].zero_()
            mpi.allreduce_(self._VT)
            self._vt_range = None
        return self._VT[:self._P, :self._P].T

    # ------------------------------------------------------------------------------------------------
    def _buffers(self, P, Pp):
        if self._bufP != (P, Pp):
            z = _kernels.zeros
            # [S0 | Fsum | var_sum | SExp | CEO]: one all-reduce, or two (head = what the solve needs first) when pipelined
            self._second = z(3 * Pp * Pp + Pp + 8)
            self._ZT = self._tau = None
            self._Sshift = None
            self._Swork = _kernels.empty(Pp, Pp)
            self._VT = z(Pp, Pp)
            self._vecs = z(8, Pp)                            # ev, VtF, rhoVar, snr, invEv, update, 2 x scratch
            self._scal = z(2)
            self._info = torch.zeros(1, dtype=torch.int32, device=global_defs.device())
            self._bufP = (P, Pp)
        return self._second

    def _mats(self, Pp):
        s = self._second
        head = Pp * Pp + Pp + 8
        S0 = s[0:Pp * Pp].view(Pp, Pp)
        Fsum = s[Pp * Pp:Pp * Pp + Pp]
        var_sum = s[Pp * Pp + Pp:Pp * Pp + Pp + 1]
        SExp = s[head:head + Pp * Pp].view(Pp, Pp)
        CEO = s[head + Pp * Pp:head + 2 * Pp * Pp].view(Pp, Pp)
        return S0, SExp, CEO, Fsum, var_sum

    # ---- pieces of get_tdvp_equation (tdvp.py:36-52) on one chunk ------------------------------------
    def _pass1_chunk(self, E, lp, O, n, ldo, first):
        _kernels.moments1(E, lp, O, n, ldo, first)

    def _pass2_chunk(self, E, lp, O, n, n_pad, ldo, Pp, meanO, meanE, scratch, which="all"):
        """Centring + force vector (which != "rest"), then the weighted Grams: "all" = S0, SExp, C_EO in one launch;
        "s0" = only S0 (what the solve needs first); "rest" = SExp and C_EO on the already centred O."""
        S0, SExp, CEO, Fsum, var_sum = self._mats(Pp)
        dE, wE, wLp = scratch
        if which != "rest":
            if n_pad > n:
                O[n:n_pad].zero_(); wE[n:n_pad].zero_(); wLp[n:n_pad].zero_()
            _kernels.center_force(O, n, ldo, meanO, E, lp, meanE, dE, wE, wLp, Fsum, var_sum)
        mats, weights = ([S0], [None]) if which != "rest" else ([], [])
        split = []
        if which != "s0":
            tgt = split if self.gramPrecision == "split" else None
            if self.computeSExp is True or (self.computeSExp and not self._lazy_ok):
                (tgt.append((wLp, SExp)) if tgt is not None else (mats.append(SExp), weights.append(wLp)))
            if self.computeSNR and self.solver == "eigh":
                (tgt.append((wE, CEO)) if tgt is not None else (mats.append(CEO), weights.append(wE)))
        if mats:
            _kernels.gram(O, n_pad, ldo, Pp, weights, mats)
        for wgt, mat in split:
            _kernels.gram_split(O, n, ldo, Pp, wgt, mat)

    def _finish(self, P, Pp, N, first, pipeline=None, early=None):
        """All-reduce of the second moments, normalisation, regularised solve (tdvp.py:50-51,57-94).

        `pipeline` (several ranks, blocked eigensolver) = (m_head, fn, n_local) with fn(lo, hi) launching the SExp / C_EO Grams
        of the local samples [lo, hi).  Then the S0 block is all-reduced first (the solver rank fills its wait for it with
        fn(0, m_head)), the solver rank runs the serial stages of the eigensolver (tridiagonalisation, divide & conquer) while
        the other ranks build those Grams, and the factors are broadcast for the sharded back-transformation -- the serial
        stages leave the critical path of every rank but one.  `early`: enqueued on every rank right
        before those serial stages (the observables block, which does not depend on the solve)."""
        S0, SExp, CEO, Fsum, var_sum = self._mats(Pp)
        head = Pp * Pp + Pp + 8
        R, rank = mpi.comm.Get_size(), mpi.comm.Get_rank()
        eager_sexp = self.computeSExp is True or (self.computeSExp and not self._lazy_ok)
        use_ceo = self.computeSNR and self.solver == "eigh"
        rest_mats = ([SExp] if eager_sexp else []) + ([CEO] if use_ceo else [])
        if pipeline is not None and pipeline[0] > 0:
            pipeline[1](0, pipeline[0])          # solver rank: fills its wait for the slowest S0 pass
        if R > 1:
            # only the upper-triangular tiles are computed, so only they cross NVLink (packed upper triangles)
            self._allreduce_packed([S0] + ([] if pipeline is not None else rest_mats), Pp, self._second[Pp * Pp:head])
        inv = 1.0 / N
        _kernels.sym_finalize(S0, Pp, inv)
        F = Fsum * inv
        self.ElocVar = (var_sum[0] * inv).clone()
        # S0 / SExp / S (P x P, 0.5 GB each at P = 8187) are views of the persistent buffers: valid until the next call
        self.S0, self.F0 = S0[:P, :P], F[:P]
        S = S0
        if self.diagonalShift > 1e-10:  # tdvp.py:50-51
            if self._Sshift is None:
                self._Sshift = _kernels.empty(Pp, Pp)
            _kernels.diag_shift(S0, self._Sshift, Pp, P, self.diagonalShift)
            S = self._Sshift
        self.S = S[:P, :P]
        meanE2 = float(first[2]) * inv  # mean(Eloc**2) of the un-centred local term (tdvp.py:93); host scalar
        ev, VtF, rhoVar, snr, invEv, update, w0, w1 = [self._vecs[i] for i in range(8)]
        self._Swork.copy_(S)
        if pipeline is not None:
            ws = _kernels.workspace(_kernels.eigh_workspace_bytes(P, Pp))
            if self._ZT is None:
                self._ZT, self._tau = _kernels.empty(Pp, Pp), _kernels.zeros(2 * Pp)
            if early is not None:
                early()       # work that does not depend on the solve (observables, with their small collectives): enqueued
                              # on every rank BEFORE the solver rank's serial stages, i.e. off the critical path
            if rank == self.solverRank:
                _kernels.eigh_factor(self._Swork, P, Pp, ev, self._ZT, self._tau[:Pp], ws)
                self._tau[Pp:].copy_(ev)
            pipeline[1](pipeline[0], pipeline[2])
            self._allreduce_packed(rest_mats, Pp, None)
            mpi.broadcast_(self._Swork, self.solverRank)        # the reflectors
            mpi.broadcast_(self._ZT, self.solverRank)           # eigenvectors of the tridiagonal matrix (rows)
            mpi.broadcast_(self._tau, self.solverRank)          # tau | ev
            ev.copy_(self._tau[Pp:])
        else:
            ws = None
        if eager_sexp:
            _kernels.sym_finalize(SExp, Pp, inv)
        if use_ceo:
            _kernels.sym_finalize(CEO, Pp, inv)
        if eager_sexp:
            self.SExp = SExp[:P, :P]
        elif self.computeSExp:
            O, n, n_pad, w = self._lazy_src
            self.SExp = LazyGram(self, self._lazy_gen, O, n, n_pad, Pp, P, w, N)
        else:
            self.SExp = None
        if self.solver == "eigh":
            if ws is None:
                ws = _kernels.workspace(_kernels.eigh_workspace_bytes(P, Pp))
            nblk = Pp // 128
            self._vt_range = None
            if R > 1 and nblk >= R and self.shardSolve:
                # The O(P^3) stages after the serial ones -- back-transformation, V^T C V for the SNR, the update -- are sharded
                # over the eigenvector index in blocks of 128 and joined by one all-reduce of 5 P doubles.  The serial stages
                # themselves are either replicated (every rank holds the all-reduced S) or, pipelined, run on the solver rank.
                row0, nrows = solve_shard_range(Pp, R, rank)
                vec = self._vecs[1:6]
                vec.zero_()
                if pipeline is not None:
                    _kernels.eigh_backtransform(self._Swork, self._tau[:Pp], self._ZT, P, Pp, self._VT, row0, nrows, ws)
                else:
                    _kernels.eigh_cols(self._Swork, P, Pp, ev, self._VT, row0, nrows, ws)
                _kernels.solve_tail_range(ev, self._VT, P, Pp, F, CEO if use_ceo else None, float(N), self.svdTol, self.snrTol,
                                          self.useSNR, row0, nrows, VtF, rhoVar if use_ceo else None, snr if use_ceo else None,
                                          invEv, update, ws)
                mpi.allreduce_(vec)
                _kernels.solve_scalars(S, S0, P, Pp, F, update, meanE2, self._scal, self._vecs[6:8].reshape(-1))
                self._vt_range = (row0, nrows)
            else:
                _kernels.eigh(self._Swork, P, Pp, ev, self._VT, ws)
                _kernels.solve_tail(ev, self._VT, P, Pp, F, S, S0, CEO if use_ceo else None, float(N), self.svdTol, self.snrTol,
                                    self.useSNR, meanE2, VtF, rhoVar if use_ceo else None, snr if use_ceo else None, invEv, update,
                                    self._scal, ws)
            self._P = P
            # fresh P-vectors: main.py:187-188 appends tdvpEq.ev / .snr to histories (JAX arrays are immutable; views of
            # the persistent work buffer would all end up equal to the last step)
            self.ev, self.VtF, self.invEv = ev[:P].clone(), VtF[:P].clone(), invEv[:P].clone()
            self.rhoVar, self.snr = (rhoVar[:P].clone(), snr[:P].clone()) if use_ceo else (None, None)
        else:
            if not self.diagonalShift > 1e-10:
                raise ValueError("solver='cholesky' needs diagonalShift > 0: S is rank deficient otherwise (SURVEY fact 5)")
            self._info.zero_()
            _kernels.chol_solve(self._Swork, P, Pp, F, update, self._info)
            _kernels.solve_scalars(S, S0, P, Pp, F, update, meanE2, self._scal, self._vecs[6:8].reshape(-1))
            if int(self._info.item()) != 0:
                raise RuntimeError(f"Cholesky failed: non-positive pivot at index {int(self._info.item()) - 1}")
            self._P = None
            self.ev = self.VtF = self.invEv = self.rhoVar = self.snr = None
        self.solverResidual, self.tdvp_error = self._scal[0].clone(), self._scal[1].clone()
        return update[:P].clone()

    def _allreduce_packed(self, mats, Pp, tail):
        """SUM over ranks of the upper tiles of `mats` (+ the `tail` vector) with one all-reduce of the packed buffer."""
        if not mats and tail is None:
            return
        ln = _kernels.packed_tiles_len(Pp)
        n_tail = 0 if tail is None else tail.numel()
        need = len(mats) * ln + n_tail
        if getattr(self, "_packed", None) is None or self._packed.numel() < need:
            self._packed = None
            self._packed = _kernels.empty(3 * ln + Pp + 8)
        buf = self._packed[:need]
        for m, M in enumerate(mats):
            _kernels.pack_upper(M, Pp, buf[m * ln:(m + 1) * ln])
        if n_tail:
            buf[len(mats) * ln:].copy_(tail)
        mpi.allreduce_(buf)
        for m, M in enumerate(mats):
            _kernels.unpack_upper(buf[m * ln:(m + 1) * ln], Pp, M)
        if n_tail:
            tail.copy_(buf[len(mats) * ln:])

    # ---- sample partition of the fused path -----------------------------------------------------------
    def _pipelined(self, R, P, Pp):
        """The solver-rank pipeline applies to a multi-rank eigen-solve on the blocked path with a sharded back-transformation."""
        return (R > 1 and self.pipelineSolve and self.solver == "eigh" and self.shardSolve and Pp // 128 >= R
                and P >= 384 and P <= 25 * 1024)

    def _gram_seconds_per_sample(self, P):
        """(S0, SExp + C_EO) Gram seconds per sample on one B200: FP64 DMMA at 34.5 TFLOP/s, the tcgen05 split path at
        150 TFLOP/s FP64-equivalent (measured)."""
        n_rest = (1 if (self.computeSExp is True) else 0) + (1 if (self.computeSNR and self.solver == "eigh") else 0)
        per = P * (P + 1.0)
        return per / 34.5e12, n_rest * per / (150e12 if self.gramPrecision == "split" else 34.5e12)

    def sample_partition(self, N, R, P, pipelined=None):
        """[(first, n)] per rank.  Equal contiguous shards unless the solve is pipelined: then the solver rank,
        which spends E(P) seconds in the serial stages of the eigensolver while the others build the SExp / C_EO Grams, gets
        n0 samples and the others n1 with (n1 - n0) (g_s0 + g_rest) = E, g_s0 / g_rest = seconds per sample of the S0 and of the
        SExp + C_EO Grams -- all ranks finish together (or, when E dominates, the share the solver rank completes inside the
        others' S0 pass).  Deterministic in (N, R, P): a model of the B200, not a measurement of the run, so the summation
        order of a run is reproducible.  `solverShare` (fraction of an equal share) overrides the model."""
        base, rem = N // R, N % R
        equal = [(r * base + min(r, rem), base + (1 if r < rem else 0)) for r in range(R)]
        Pp = _kernels.round_up(P, 128)
        if pipelined is None:
            pipelined = self._pipelined(R, P, Pp)
        if not pipelined:
            return equal
        if self.solverShare is not None:
            n0 = int(max(0.0, min(1.0, float(self.solverShare))) * N / R)
        else:
            # workers: (g_s0 + g_rest) n1; solver rank: g_s0 n1 (it waits for the slowest S0, filling the wait with the
            # SExp / C_EO Grams of its first samples) + E + the rest of its own SExp / C_EO  =>  (n1 - n0) (g_s0 + g_rest) = E
            g_s0, g_rest = self._gram_seconds_per_sample(P)
            delta = solver_seconds(P) / (g_s0 + g_rest)
            n0 = int(max(0.0, (N - (R - 1) * delta) / R))
            # E dominates (many ranks): the solver rank can still take what it finishes, S0 and SExp / C_EO, inside the other
            # ranks' S0 pass -- n0 (g_s0 + g_rest) = n1 g_s0 -- which shortens that pass for everybody
            frac = g_s0 / (g_s0 + g_rest)
            n0 = max(n0, int(N * frac / ((R - 1) + frac)))
        n0 = min(n0 // 16 * 16, N)
        others = R - 1
        b2, r2 = (N - n0) // others, (N - n0) % others
        out, first, k = [], 0, 0
        for r in range(R):
            if r == self.solverRank:
                n = n0
            else:
                n = b2 + (1 if k < r2 else 0)
                k += 1
            out.append((first, n))
            first += n
        return out

    # ---- reference entry points on materialised arrays -----------------------------------------------
    def get_tdvp_equation(self, Eloc, gradients, logProbs):
        """tdvp.py:36-52 for given (1,n), (1,n,P), (1,n) arrays; returns (S, F, EOdata)."""
        self._solve_materialised(Eloc, gradients, logProbs, solve=False)
        dE = (_kernels.as_dev(Eloc) - self.ElocMean)
        EO = dE[..., None] * (_kernels.as_dev(gradients) - self._gradMean[None, None, :])
        return self.S, self.F0, EO

    def solve(self, Eloc, gradients, logProbs):
        """tdvp.py:73-94: returns (update, solverResidual, tdvp_error)."""
        update = self._solve_materialised(Eloc, gradients, logProbs, solve=True)
        return update, self.solverResidual, self.tdvp_error

    def _solve_materialised(self, Eloc, gradients, logProbs, solve=True):
        E = _kernels.as_dev(Eloc).reshape(-1)
        lp = _kernels.as_dev(logProbs).reshape(-1)
        G = _kernels.as_dev(gradients)
        P = G.shape[-1]
        G = G.reshape(-1, P)
        n = E.shape[0]
        N = mpi.globNumSamples if mpi.globNumSamples else n * mpi.comm.Get_size()
        Pp = _kernels.round_up(P, 128)
        self._buffers(P, Pp).zero_()
        n_pad = _kernels.round_up(max(n, 1), 16)
        O = _kernels.zeros(n_pad, Pp)
        O[:n, :P] = G
        first = _kernels.zeros(4 + Pp)
        self._pass1_chunk(E, lp, O, n, Pp, first)
        mpi.allreduce_(first)
        self._set_first(first, N, P)
        scratch = (_kernels.zeros(n_pad), _kernels.zeros(n_pad), _kernels.zeros(n_pad))
        self._lazy_ok, self._lazy_src, self._lazy_gen = True, (O, n, n_pad, scratch[2]), None   # this O is private: never reused
        self._pass2_chunk(E, lp, O, n, n_pad, Pp, Pp, (first[4:] / N).contiguous(), float(first[0]) / N, scratch)
        return self._finish(P, Pp, N, first)

    def _set_first(self, first, N, P):
        self.ElocMean = first[0] / N         # tdvp.py:37
        self.ElocMeanAbs = first[1] / N      # tdvp.py:38
        self._gradMean = first[4:4 + P] / N  # tdvp.py:41

    # ---- the right-hand side ---------------------------------------------------------------------------
    def __call__(self, netParameters, t, psi, evolutionEq, **rhsArgs):
        nSamplesTDVP = rhsArgs["nSamplesTDVP"]
        nSamplesObs = rhsArgs["nSamplesObs"]
        mpi.globNumSamples = nSamplesTDVP
        psi.set_parameters(netParameters)     # tdvp.py:100-101 (psi is left at the trial parameters, as in the reference)
        timings = rhsArgs["timings"]
        acc = {}

        def tic():
            if timings is not None:
                torch.cuda.synchronize()
                return time.perf_counter()
            return 0.0

        def toc(name, t0):
            if timings is not None:
                torch.cuda.synchronize()
                acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0

        fused = hasattr(psi, "sample_range") and hasattr(evolutionEq, "equation_struct")
        if not fused:  # duck-typed psi / evolutionEq: the reference's materialised flow (tdvp.py:116-127)
            t0 = tic(); sampleConfigs, logProbs = psi.sample(numSamples=nSamplesTDVP); toc("sampling", t0)
            t0 = tic(); Eloc, sampleGradients, logProbs = evolutionEq(psi, sampleConfigs, t); toc("compute Eloc", t0)
            t0 = tic(); update, self.solverResidual, self.tdvp_error = self.solve(Eloc, sampleGradients, logProbs); toc("solve TDVP eqn.", t0)
            x_all, lp_all, E_all = sampleConfigs.reshape(-1, sampleConfigs.shape[-1]), logProbs.reshape(-1), Eloc.reshape(-1)
        else:
            self._early_info = None
            update, x_all, lp_all, E_all = self._fused_rhs(psi, evolutionEq, t, nSamplesTDVP, tic, toc,
                                                           obs_early=nSamplesObs if nSamplesObs <= nSamplesTDVP else None)

        if nSamplesObs > nSamplesTDVP:  # tdvp.py:130-134
            t0 = tic()
            xo, lpo = psi.sample(numSamples=nSamplesObs)
            x_obs, lp_obs = xo.reshape(-1, xo.shape[-1]), lpo.reshape(-1)
            n_obs_glob = nSamplesObs
            toc("sampling observables", t0)
        else:
            x_obs, lp_obs, n_obs_glob = x_all, lp_all, nSamplesTDVP

        if timings is not None:
            for k, v in acc.items():
                timings.timing_dict.setdefault(k, []).append(v)

        if bool(torch.isnan(update).any()):  # tdvp.py:136-141
            print(self.S0)
            print(self.F0)
            print("nan encountered. Exitting.")
            raise SystemExit(1)

        info = getattr(self, "_early_info", None)     # pipelined multi-rank solve: already enqueued before the eigensolve
        if info is None or not fused:
            info = self._observables(psi, x_obs, lp_obs, E_all, n_obs_glob, nSamplesObs)
        self._early_info = None
        return update, info

    def _plan_chunks(self, n_local, Pp):
        """Rows of the O buffer and whether the whole local O fits (then local terms are evaluated once).  Planned once
        per (n_local, Pp, chunkSamples, memoryFraction): the chunking fixes the floating-point summation order, so it must
        not drift between consecutive right-hand sides of a run (after the first call the cached O buffer no longer counts as
        free memory)."""
        plan_key = (n_local, Pp, self.chunkSamples, self.memoryFraction)
        plans = self.__dict__.setdefault("_plans", {})
        if plan_key not in plans:
            plans[plan_key] = self._plan_chunks_now(n_local, Pp)
        return plans[plan_key]

    def _plan_chunks_now(self, n_local, Pp):
        free, _ = torch.cuda.mem_get_info(global_defs.device())
        if getattr(self, "_Obuf", None) is not None:
            free += self._Obuf.numel() * 8        # a buffer of an earlier plan is released before the new one is allocated
        fixed = 8 * Pp * Pp * 8 + (64 << 20)  # S0, SExp, CEO, shifted S, work copy, VT, eigh scratch (2)
        budget = max(int((free - fixed) * self.memoryFraction), 16 * Pp * 8)
        rows = self.chunkSamples if self.chunkSamples else budget // (Pp * 8)
        rows = max(16, rows // 16 * 16)
        n_pad = _kernels.round_up(max(n_local, 1), 16)
        if rows >= n_pad:
            return n_pad, True
        return rows, False

    def _fused_rhs(self, psi, evolutionEq, t, N, tic, toc, obs_early=None):
        h = psi.net.handle
        P, Pp, d = h.P, h.Pp, h.dim
        R, rank = mpi.comm.Get_size(), mpi.comm.Get_rank()
        pipelined = self._pipelined(R, P, Pp)
        if pipelined:   # needs the whole rank-local O resident on every rank: else equal shards and the replicated solve
            part = self.sample_partition(N, R, P, True)
            fits = self._plan_chunks(max(n for _, n in part), Pp)[1]
            flag = torch.tensor([0.0 if fits else 1.0], dtype=torch.float64, device=global_defs.device())
            mpi.allreduce_(flag)
            pipelined = bool(flag.item() == 0.0)
        first_idx, n_local = self.sample_partition(N, R, P, pipelined)[rank]
        self._last_partition = (pipelined, first_idx, n_local)
        key = psi.sampler.next_key()                       # sampler.py:73 (one key per psi.sample call)
        chi2_all = psi.chi2_draws(n_local, first_idx, N)
        eq = evolutionEq.equation_struct(t)
        self._buffers(P, Pp).zero_()
        rows, stored = self._plan_chunks(n_local, Pp)
        self._gen += 1
        self._lazy_ok = stored
        if self.computeSExp and self.computeSExp is not True and mpi.comm.Get_size() > 1:
            # LazyGram.dot is collective: every rank must take the same branch
            flag = torch.tensor([1.0 if self._lazy_ok else 0.0], dtype=torch.float64, device=global_defs.device())
            mpi.allreduce_(flag)
            self._lazy_ok = bool(flag.item() == mpi.comm.Get_size())
        if getattr(self, "_Obuf", None) is None or self._Obuf.shape != (rows, Pp):
            self._Obuf = None
            self._Obuf = _kernels.zeros(rows, Pp)
            self._scratch = (_kernels.zeros(rows), _kernels.zeros(rows), _kernels.zeros(rows))
        O = self._Obuf
        x_all, lp_all, E_all = _kernels.empty(n_local, d), _kernels.empty(n_local), _kernels.empty(n_local)
        first = _kernels.zeros(4 + Pp)
        chunks = [(c0, min(rows, n_local - c0)) for c0 in range(0, n_local, rows)] or [(0, 0)]

        def local_chunk(c0, cn, sample):
            if sample:
                t0 = tic()
                chi2 = chi2_all[c0:c0 + cn] if chi2_all is not None else None
                x, lps = psi.sample_range(key, first_idx + c0, cn, N, chi2)
                x_all[c0:c0 + cn] = x
                toc("sampling", t0)
            t0 = tic()
            out = _kernels.local_terms(h, psi._flat, x_all[c0:c0 + cn], eq, O=O, ldo=Pp, want=("eloc", "logp"))
            E_all[c0:c0 + cn] = out["eloc"]; lp_all[c0:c0 + cn] = out["logp"]
            toc("compute Eloc", t0)

        # pass 1: samples, local terms, first moments (tdvp.py:117-122,37-41)
        for c0, cn in chunks:
            if cn == 0:
                continue
            local_chunk(c0, cn, True)
            t0 = tic()
            self._pass1_chunk(E_all[c0:c0 + cn], lp_all[c0:c0 + cn], O, cn, Pp, first)
            toc("solve TDVP eqn.", t0)
        t0 = tic()
        mpi.allreduce_(first)
        self._set_first(first, N, P)
        meanO = (first[4:] / N).contiguous()
        meanE = float(first[0]) / N
        toc("solve TDVP eqn.", t0)
        # pass 2: centring, force vector, weighted Grams (tdvp.py:40-47); local terms are re-evaluated chunk by chunk
        # when the whole O does not fit in memory (same key and counters -> identical samples)
        pipelined = pipelined and stored
        which = "s0" if pipelined else "all"
        for c0, cn in chunks:
            if cn == 0:
                continue
            if not stored:
                local_chunk(c0, cn, False)
            t0 = tic()
            self._pass2_chunk(E_all[c0:c0 + cn], lp_all[c0:c0 + cn], O, cn, _kernels.round_up(cn, 16), Pp, Pp, meanO, meanE,
                              self._scratch, which)
            toc("solve TDVP eqn.", t0)
        t0 = tic()
        self._lazy_src, self._lazy_gen = (O, n_local, _kernels.round_up(max(n_local, 1), 16), self._scratch[2]), self._gen
        rest = None
        if pipelined:
            n_pad_all = _kernels.round_up(max(n_local, 1), 16)

            def rest_range(lo, hi):
                """SExp / C_EO Grams of the local samples [lo, hi) (lo, and hi unless it is n_local, multiples of 16)."""
                if hi > lo:
                    sc = tuple(a[lo:] for a in self._scratch)
                    self._pass2_chunk(E_all[lo:hi], lp_all[lo:hi], O[lo:], hi - lo, (n_pad_all if hi == n_local else hi) - lo, Pp, Pp,
                                      meanO, meanE, sc, "rest")
            # samples whose SExp / C_EO Grams the solver rank builds while it waits for the other ranks' (longer) S0 pass
            m_head = 0
            if rank == self.solverRank and n_local > 0:
                g_s0, g_rest = self._gram_seconds_per_sample(P)
                n1 = max(n for _, n in self.sample_partition(N, R, P, True))
                if g_rest > 0 and n1 > n_local:
                    m_head = min(n_local, int((n1 - n_local) * g_s0 / g_rest)) // 16 * 16
            rest = (m_head, rest_range, n_local)
        early = None
        if pipelined and obs_early is not None:
            def early():
                self._early_info = self._observables(psi, x_all, lp_all, E_all, N, obs_early)
        update = self._finish(P, Pp, N, first, pipeline=rest, early=early)
        toc("solve TDVP eqn.", t0)
        return update, x_all, lp_all, E_all

    def _observables(self, psi, x, lp, E_all, n_glob, nSamplesObs):
        """tdvp.py:143-162.  With several ranks the sums are all-reduced (the reference uses rank-local means)."""
        d = x.shape[1]
        n = x.shape[0]
        ws = _kernels.obs_workspace(d)
        first = _kernels.zeros(d + 2)
        first[d + 1] = -float("inf")
        same = E_all.shape[0] == n
        if n > 0:            # a solver rank of a pipelined solve may own no samples: its sums stay zero
            _kernels.obs_first(x, lp, E_all if same else None, n, d, first, ws)
        if same:
            mx = first[d + 1].clone()
        else:  # observables were re-sampled: max E_loc still refers to the TDVP samples (tdvp.py:150)
            tmp = _kernels.zeros(3)
            tmp[2] = -float("inf")
            if E_all.shape[0] > 0:
                _kernels.obs_first(E_all.reshape(-1, 1), None, E_all, E_all.shape[0], 1, tmp, ws)
            mx = tmp[2].clone()
        mpi.allreduce_(first[:d + 1])
        mean = (first[:d] / n_glob).contiguous()
        central = _kernels.zeros(d * d + 4 * d)
        if n > 0:
            _kernels.obs_central(x, n, d, mean, central, ws)
        mpi.allreduce_(central)
        central = central / n_glob
        info = {}
        info["x1"] = mean
        info["covar"] = central[:d * d].view(d, d)
        info["entropy"] = -first[d] / n_glob
        for i, m in enumerate([3, 4, 5, 6]):
            info[f"x{m}"] = central[d * d + i * d:d * d + (i + 1) * d]
        if mpi.comm.Get_size() > 1:
            import torch.distributed as dist
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        info["max_grad"] = mx
        # integrate on small cells (tdvp.py:152-162): both draws use psi.sampler.key itself
        first_idx, n_loc = mpi.shard_range(nSamplesObs)
        key = psi.sampler.key
        sums = _kernels.zeros(3)
        lims = [1, 0.5, 0.1]
        T = 10
        for j, lim_normal in enumerate(lims):
            lim = lim_normal * math.sqrt(T)
            pts = _kernels.ball_points(key, first_idx, n_loc, nSamplesObs, d, lim)
            lpb = psi(pts[None, ...])[0] if not hasattr(psi, "net") else _kernels.logp(psi.net.handle, psi._flat, pts)
            _kernels.sum_exp(lpb, n_loc, sums[j:j + 1], ws)
        mpi.allreduce_(sums)
        for j, lim_normal in enumerate(lims):
            lim = lim_normal * math.sqrt(T)
            sphere_volume = math.pi ** (d / 2) / scipy.special.gamma(d / 2 + 1) * lim ** d
            info[f"integral_{lim_normal}sigma"] = sums[j] / nSamplesObs * sphere_volume
        return info
