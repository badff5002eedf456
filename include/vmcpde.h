/* vmcpde.h -- C-ABI of libvmcpde.so: the B200 (sm_100a) implementation of vmc_pde's TDVP time-step
 * hot path.  Plain pointers and sizes only; every `double*` below is DEVICE memory unless the comment
 * says host; `stream` is a cudaStream_t passed as void*.  Functions return 0 on success, a nonzero
 * code otherwise (message via vmcpde_last_error()).  No function allocates device memory behind the
 * caller's back or synchronises the host, except where stated.
 *
 * The reference (RehMoritz/vmc_pde) has no FFI: its boundary is the Python surface used by main.py.
 * Each entry point cites the reference code it replaces (paths relative to vmc_fluids/).  The Python
 * mirror of that surface lives in vmc_pde_b200/*.py; an XLA-FFI shim over the same entry points is in
 * vmc_pde_b200/csrc/xla_ffi_shim.cc.
 */
#ifndef VMCPDE_H_
#define VMCPDE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMCPDE_VERSION 100

/* error codes */
#define VMCPDE_OK 0
#define VMCPDE_EINVAL 1      /* bad argument */
#define VMCPDE_EUNSUPPORTED 2 /* configuration not built (dimension, hidden layers, ...) */
#define VMCPDE_ECUDA 3       /* CUDA runtime error */
#define VMCPDE_ENUMERIC 4    /* numerical failure (non-positive Cholesky pivot, no convergence) */

typedef struct vmcpde_flow vmcpde_flow; /* opaque ansatz description */
typedef void* vmcpde_stream;

/* coupling variants: class defaults of net.py:69-71 (the reference selects them by editing source) */
enum { VMCPDE_NO_ADD = 0, VMCPDE_DIFFERENT_ADD = 1, VMCPDE_JAC_EQ_1 = 2, VMCPDE_ADD_S = 3 };
/* OR-ed into `variant`: SingleBlock.global_change (net.py:72,80-82,115-116,149-150) -- every block owns global_offset[dim] and
 * global_scale[1] (flat order: before s1) and maps  result -> scale * result + offset  after the coupling.  The reference's
 * "inverse" branch (used for sampling) undoes the affine step after the inverse coupling, too; that order is kept as it is. */
#define VMCPDE_GLOBAL_CHANGE 0x100
/* latent densities: net.py:197-198 */
enum { VMCPDE_GAUSS = 0, VMCPDE_STUDENT_T = 1 };
/* evolution equations: evolutionEq.py:54-60 */
enum {
  VMCPDE_DIFFUSION = 0, VMCPDE_DIFFUSION_DRIFT = 1, VMCPDE_DIFFUSION_ANISOTROPIC = 2,
  VMCPDE_ADVECTION_HAMILTONIAN = 3, VMCPDE_ADVECTION_PAPER = 4, VMCPDE_ADVECTION_HAMILTONIAN_WDISS = 5
};

/* net.INNwProb(inds_up, inds_down, intmediate, offset, latentSpaceName, dim), net.py:185-207 */
typedef struct vmcpde_flow_config {
  int32_t dim;              /* d */
  int32_t depth;            /* number of SingleBlocks */
  int32_t n_hidden_layers;  /* len(intmediate): 1 (streaming fast path, width <= 256) to 3 (generic path, widths <= 32) */
  int32_t hidden;           /* intmediate[0] */
  int32_t variant;          /* VMCPDE_NO_ADD ... , optionally | VMCPDE_GLOBAL_CHANGE */
  int32_t latent;           /* VMCPDE_GAUSS | VMCPDE_STUDENT_T */
  const int32_t* ind_up;    /* host, depth * (dim/2)       : var_state.py:116 */
  const int32_t* ind_down;  /* host, depth * (dim - dim/2) : var_state.py:117 */
  const double* offset;     /* host, dim : network_args["offset"] */
  const int32_t* hidden_widths; /* host, n_hidden_layers entries = intmediate (net.py:53-58); may be NULL when
                                   n_hidden_layers == 1 (then `hidden` is used) */
} vmcpde_flow_config;

/* evolutionEq.EvolutionEquation parameters, evolutionEq.py:61-77 */
typedef struct vmcpde_equation {
  int32_t mode;            /* VMCPDE_DIFFUSION ... */
  double D, mu, m, omega, lam, T, gamma;
  double t;                /* time passed to the velocity field */
  const double* tangents;  /* device, dim*dim row-major factor A with D = A^T A; required for
                              VMCPDE_DIFFUSION_ANISOTROPIC (evolutionEq.py:18-20), else NULL */
} vmcpde_equation;

const char* vmcpde_last_error(void);
int vmcpde_version(void);
/* leading dimension (in doubles) used for rows of O and for S/V: num_params rounded up to 128 */
int32_t vmcpde_padded_params(int32_t num_params);

/* ---- ansatz handle ----------------------------------------------------------------------------- */
int vmcpde_flow_create(const vmcpde_flow_config* cfg, vmcpde_flow** out);
void vmcpde_flow_destroy(vmcpde_flow* f);
/* VarState.numParameters, var_state.py:27 */
int32_t vmcpde_flow_num_params(const vmcpde_flow* f);
/* flat layout (var_state.py:106-108): out[0..3] = offsets of L, L_diag, dist_params, mu;
 * out[4+b] = offset of blocks_b (host array of 4+depth ints) */
int vmcpde_flow_param_offsets(const vmcpde_flow* f, int32_t* out);

/* ---- (1) sampler ------------------------------------------------------------------------------- */
/* Replaces Sampler.__call__ exact branch (sampler.py:25-34,72-86) + VarState.sample
 * (var_state.py:76-79) + INNwProb(evaluate=False, inv=True) (net.py:214-217).
 * Draws global sample indices [first, first+n) of the n_total-sample stream of key (key0,key1):
 * xi = normal(key, (1, n_total, d)) in JAX's threefry2x32 counter layout, z = mu + chol(S) xi
 * [* sqrt(nu/chi2) for Student-t] + offset, x = INN^-1(z), logp = log p_lat(z - offset) - logJ.
 * chi2: n chi-square(nu) variates (Student-t only; the reference draws them from NumPy's global RNG,
 * sampler.py:32), NULL for Gauss.  z_out may be NULL. */
int vmcpde_sample(const vmcpde_flow* f, const double* theta, uint32_t key0, uint32_t key1,
                  int64_t first, int64_t n, int64_t n_total, const double* chi2,
                  double* x, double* logp, double* z_out, vmcpde_stream stream);
/* raw N(0,1) draws in the same layout: out[i] = normal(key, (count_total,))[first + i]  (tdvp.py:154) */
int vmcpde_normal(uint32_t key0, uint32_t key1, int64_t first, int64_t n, int64_t count_total,
                  double* out, vmcpde_stream stream);
/* U[0,1) draws, jax.random.uniform float64 layout (tdvp.py:155) */
int vmcpde_uniform(uint32_t key0, uint32_t key1, int64_t first, int64_t n, int64_t count_total,
                   double* out, vmcpde_stream stream);

/* ---- (2) ansatz evaluation and local terms ----------------------------------------------------- */
/* VarState.__call__(mode="eval"), var_state.py:38-43: logp[i] = log p(x[i]) */
int vmcpde_logp(const vmcpde_flow* f, const double* theta, const double* x, int64_t n, double* logp,
                vmcpde_stream stream);
/* Fused VarState.__call__(mode="eval_coordgrads") + VarState.hessian + EvolutionEquation.__call__
 * (var_state.py:55-67, evolutionEq.py:84-119).  Per sample: logp, E_loc = d_t log p, the directional
 * derivatives of log p (grad_x for every mode but anisotropic, where they are A grad_x), the weighted
 * second-derivative sum entering E_loc, and the row O[i,:] = d logp/d theta in flat order.
 * O has leading dimension ldo >= P; columns [P, ldo) are written as zeros.  Any output may be NULL. */
int vmcpde_local_terms(const vmcpde_flow* f, const double* theta, const double* x, int64_t n,
                       const vmcpde_equation* eq, double* eloc, double* logp, double* grad,
                       double* lap, double* O, int64_t ldo, vmcpde_stream stream);
/* INN map alone, INN.__call__(x, inv) (net.py:168-182), as used through net.apply(params, x, evaluate=False, inv)
 * (net.py:214-217; main.py:81-82,202): y = INN(x) or INN^-1(x); logjac = its log-Jacobian; optionally the latent
 * log-pdf of the INPUT, log p_lat(x - offset).  logjac / latent_logpdf_of_input may be NULL. */
int vmcpde_flow_transform(const vmcpde_flow* f, const double* theta, const double* x, int64_t n,
                          int32_t inverse, double* y, double* logjac, double* latent_logpdf_of_input,
                          vmcpde_stream stream);
/* VarState.hessian, var_state.py:66-67: H[i] = d x d Hessian of log p at x[i] (row-major, n*d*d) */
int vmcpde_hessian(const vmcpde_flow* f, const double* theta, const double* x, int64_t n, double* H,
                   vmcpde_stream stream);

/* ---- (3) moments, centring, Gram --------------------------------------------------------------- */
/* Local sums for mpi.global_mean / global_variance (tdvp.py:37-41, mpi_wrapper.py:129-193):
 * sums[0..3] += sum E, sum |E|, sum E^2, sum logp ; sums[4 + p] += sum_i O[i,p] for p < ldo. */
int vmcpde_moments1(const double* eloc, const double* logp, const double* O, int64_t n, int64_t ldo,
                    double* sums, void* workspace, size_t workspace_bytes, vmcpde_stream stream);
/* scratch of vmcpde_moments1 / vmcpde_center_force: one partial row of ldo doubles per block of 512 samples
 * (the column sums are reduced in two fixed-order levels, so results do not depend on the device) */
int vmcpde_moments_workspace_bytes(int64_t n, int64_t ldo, size_t* bytes);
/* tdvp.py:40-45 on one chunk: O[i,:] -= meanO (in place); dE[i] = E[i] - meanE; Fsum[p] += sum_i dE[i]*O[i,p];
 * wE[i] = dE[i]^2, wLp[i] = logp[i]^2 (row weights for the SNR covariance and SExp Grams);
 * var_sum[0] += sum_i dE[i]^2.  meanO, Fsum have ldo entries. */
int vmcpde_center_force(double* O, int64_t n, int64_t ldo, const double* meanO, const double* eloc,
                        const double* logp, double meanE, double* dE, double* wE, double* wLp,
                        double* Fsum, double* var_sum, void* workspace, size_t workspace_bytes,
                        vmcpde_stream stream);
/* Weighted Gram accumulation on FP64 tensor cores (DMMA), replacing mpi.global_covariance /
 * _cov_helper_without_p (mpi_wrapper.py:21-25,248-274; tdvp.py:46-47,68-70):
 * for m < n_mats:  S[m][a,b] += sum_i w[m][i] * O[i,a] * O[i,b]   for the tiles of the UPPER triangle.
 * weights[m] == NULL means w = 1.  weights / S are HOST arrays of device pointers.  Pp = padded
 * parameter count (multiple of 128) = row/column count and leading dimension of every S[m]; O has
 * leading dimension ldo >= Pp with zero padding columns.  n must be a multiple of 16. */
int vmcpde_gram(const double* O, int64_t n, int64_t ldo, int32_t Pp, int32_t n_mats,
                const double* const* weights, double* const* S, vmcpde_stream stream);
/* Matrix-free product with a weighted Gram: out[c] += sum_i w[i] (O[i,:] . v) O[i,c] = ((O^T diag(w) O) v)[c], two
 * streaming passes over O instead of the N P^2 flops of the matrix.  The reference reads SExp (tdvp.py:47) only through
 * normFunction(v, SExp) (stepper.py:71; main.py:24-26: v^T S v), which this serves without building SExp.
 * w may be NULL (w = 1); v, out: ldo doubles; t: n doubles of scratch; workspace: vmcpde_moments_workspace_bytes. */
int vmcpde_gram_matvec(const double* O, int64_t n, int64_t ldo, const double* w, const double* v, double* t,
                       double* out, void* workspace, size_t workspace_bytes, vmcpde_stream stream);
/* S <- scale * S on the upper triangle, mirrored to the lower; then, if shift > 1e-10,
 * S += diag(shift * diag(S)) (tdvp.py:50-51).  S_shifted may alias S or be a second Pp x Pp buffer. */
int vmcpde_sym_finalize(double* S, int32_t Pp, double scale, vmcpde_stream stream);
int vmcpde_diag_shift(const double* S, double* S_shifted, int32_t Pp, int32_t P, double shift,
                      vmcpde_stream stream);

/* General FP64 tensor-core product on the Gram pipeline: Out[M x N] = alpha * X^T Y + beta * Out with X [K x M]
 * (ldx) and Y [K x N] (ldy) row-major.  M, N multiples of 128, K of 16 (pad with zeros). */
int vmcpde_gemm_tn(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Out, int64_t ldo,
                   int32_t M, int32_t N, int64_t K, double alpha, double beta, vmcpde_stream stream);
/* Split-K form for few output tiles and a long contraction: slice s of the K range -> Part + s * M * ldo (alpha 1, beta 0);
 * the caller adds the `splits` (1..16) slices. */
int vmcpde_gemm_tn_splitk(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Part, int64_t ldo,
                          int32_t M, int32_t N, int64_t K, int32_t splits, vmcpde_stream stream);
/* Upper-triangular 128x128 tiles of Out = alpha * X^T X + beta * Out on the same pipeline (X [K x M] row-major, M multiple
 * of 128, K of 16); the lower tiles of Out are not touched. */
int vmcpde_syrk_tn(const double* X, int64_t ldx, double* Out, int64_t ldo, int32_t M, int64_t K, double alpha,
                   double beta, vmcpde_stream stream);
/* Split-precision variant of vmcpde_gram for ONE matrix on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators;
 * the north star's "FP32-split path with stated tolerance"): S (ld = Pp, upper-triangular 128x128 tiles) += O^T diag(w) O with
 * every FP64 operand sqrt(w_row) * O split into three bfloat16 slices, the six slice products with i + j <= 4 accumulated in
 * FP32 in TMEM over at most 128 samples and added in FP64.  Stated tolerance: 1e-6 relative to sqrt(S_ii S_jj).  Meant for
 * SExp (tdvp.py:47) and the SNR covariance (tdvp.py:68-70); S0 stays on vmcpde_gram.  O [n][ldo] row-major, w[n] >= 0 or
 * NULL; workspace (vmcpde_gram_split_workspace_bytes) holds the bf16 slices (3 * Pp * round_up(n, 128) * 2 bytes). */
int vmcpde_gram_split_workspace_bytes(int64_t n, int32_t Pp, size_t* bytes);
int vmcpde_gram_split(const double* O, int64_t n, int64_t ldo, int32_t Pp, const double* w, double* S, void* workspace,
                      size_t workspace_bytes, vmcpde_stream stream);
/* ---- cross-GPU sums (mpi_wrapper.py:129-274 as issued from tdvp.py:37-47) for hosts without torch.distributed ----
 * NCCL is resolved at run time (dlopen libnccl.so.2; the copy already loaded in the process is reused): no link-time
 * dependency.  `nccl_comm` is an ncclComm_t of that library.  Only the upper-triangular 128x128 tiles of a Gram matrix are
 * computed, so only they cross NVLink: vmcpde_packed_tiles_len(Pp) doubles per matrix (272 MB instead of 537 MB at Pp = 8192).
 * vmcpde_allreduce_sum: in-place SUM of `count` doubles (the first moments, P + 4).  vmcpde_allreduce_moments: n_mats (<= 8)
 * Gram matrices + `n_tail` doubles (force vector, variance sums) packed into `packed` (caller-owned, n_mats * packed_len +
 * n_tail doubles), ONE ncclAllReduce, unpacked in place.  mats is a HOST array of device pointers. */
int64_t vmcpde_packed_tiles_len(int32_t Pp);
int vmcpde_pack_upper_tiles(const double* S, int32_t Pp, double* packed, vmcpde_stream stream);
int vmcpde_unpack_upper_tiles(const double* packed, int32_t Pp, double* S, vmcpde_stream stream);
int vmcpde_allreduce_sum(void* nccl_comm, double* buf, int64_t count, vmcpde_stream stream);
int vmcpde_allreduce_moments(void* nccl_comm, double* const* mats, int32_t n_mats, int32_t Pp, double* tail, int64_t n_tail,
                             double* packed, vmcpde_stream stream);
/* communicator plumbing for hosts that have none: 128-byte id (host memory) from rank 0, shipped out of band */
int vmcpde_nccl_unique_id(char* id128);
int vmcpde_nccl_comm_init(int32_t n_ranks, int32_t rank, const char* id128, void** comm_out);
int vmcpde_nccl_comm_destroy(void* comm);
/* One launch of a register-resident DMMA.8x8x4 loop (148 x 4 CTAs x 8 warps): the FP64 tensor-pipe probe that gives the
 * roofline denominator of the S build (MEASURED_PEAKS.json has no FP64 entry).  `scratch`: 8 bytes of device memory.
 * *flops = floating-point operations of the launch; the caller times it with events on `stream`.  No allocation, no sync. */
int vmcpde_dmma_probe(void* scratch, int32_t iters, double* flops, vmcpde_stream stream);

/* ---- (4) regularised solve -------------------------------------------------------------------- */
/* Symmetric eigendecomposition on the device, replacing np.linalg.eigh at tdvp.py:61-64.
 * S (full symmetric, leading dimension ld) is destroyed; ev[n] ascending; VT row k = eigenvector k.
 * S and VT are ld x ld buffers with finite (zero) padding outside n x n when ld is a multiple of 128 >= n -- the
 * layout every caller in this package uses; then the blocked path runs for n >= 384: panel tridiagonalisation in a
 * persistent cooperative kernel (trailing matrix read once per column, rank-2nb update on FP64 tensor cores),
 * divide & conquer, compact-WY back-transformation as tensor-core products.  Otherwise (small n or unpadded ld,
 * S and VT n x ld) the unblocked kernels run.  No host synchronisation.  n <= 25472 in this release. */
int vmcpde_eigh_workspace_bytes(int32_t n, int32_t ld, size_t* bytes);
int vmcpde_eigh(double* S, int32_t n, int32_t ld, double* ev, double* VT, void* workspace,
                size_t workspace_bytes, vmcpde_stream stream);
/* Same decomposition, but only the eigenvectors [col0, col0 + ncols) are back-transformed and written (rows col0..
 * of VT; other rows untouched) -- the slice one rank of a multi-GPU solve needs (tdvp.py:61-71 sharded over the
 * eigenvector index).  col0, ncols: multiples of 128 inside the padded size on the blocked path; the unblocked path
 * writes every row. */
int vmcpde_eigh_cols(double* S, int32_t n, int32_t ld, double* ev, double* VT, int32_t col0, int32_t ncols,
                     void* workspace, size_t workspace_bytes, vmcpde_stream stream);
/* The same eigensolver in two calls, for a multi-GPU solve in which ONE rank runs the serial stages while the others
 * still build Gram matrices (np.linalg.eigh of tdvp.py:61-64 split at its only parallel seam): vmcpde_eigh_factor =
 * scaling + tridiagonalisation + divide & conquer; it leaves the reflectors in S, tau[n], Z^T (rows x ld with rows = n
 * rounded up to 128; row k = eigenvector k of the tridiagonal matrix) and ev[n] in caller-owned buffers that can be
 * broadcast.  vmcpde_eigh_backtransform applies the reflectors to the 128-aligned eigenvector slice [col0, col0+ncols)
 * (ncols <= 0: all) and writes rows col0.. of VT (VT must not alias the inputs).  factor + backtransform of all
 * columns == vmcpde_eigh bit for bit.  Blocked path only (n >= 384, ld a multiple of 128): VMCPDE_EUNSUPPORTED otherwise.
 * Workspace: vmcpde_eigh_workspace_bytes for both. */
int vmcpde_eigh_factor(double* S, int32_t n, int32_t ld, double* ev, double* ZT, double* tau, void* workspace,
                       size_t workspace_bytes, vmcpde_stream stream);
int vmcpde_eigh_backtransform(const double* reflectors, const double* tau, const double* ZT, int32_t n, int32_t ld,
                              double* VT, int32_t col0, int32_t ncols, void* workspace, size_t workspace_bytes,
                              vmcpde_stream stream);
/* number of kernel launches one vmcpde_eigh(n, ld) call issues */
int vmcpde_eigh_launch_count(int32_t n, int32_t ld, int32_t* count);
/* Everything after eigh in TDVP.transform_to_eigenbasis / TDVP.solve (tdvp.py:66-94): VtF = V^T F;
 * rhoVar = diag(V^T CEO V) - VtF^2 and snr = sqrt|N VtF^2 / rhoVar| when CEO (the dE^2-weighted Gram, ld x ld,
 * zero padded) is given; invEv = (|ev/ev_max| > 1e-14) ? 1/ev : 0; regulariser 1/(1+(svdTol/|ev/ev_max|)^6)
 * [* 1/(1+(snrTol/snr)^6)]; update = V (invEv * reg * VtF); scalars[0] = ||S update - F|| / ||F||;
 * scalars[1] = 1 + (update^T S0 update - 2 F^T update) / meanE2. */
int vmcpde_solve_tail_workspace_bytes(int32_t n, int32_t ld, size_t* bytes);
int vmcpde_solve_tail(const double* ev, const double* VT, int32_t n, int32_t ld, const double* F,
                      const double* S, const double* S0, const double* CEO, double n_glob, double svdTol,
                      double snrTol, int32_t useSNR, double meanE2, double* VtF, double* rhoVar, double* snr,
                      double* invEv, double* update, double* scalars, void* workspace,
                      size_t workspace_bytes, vmcpde_stream stream);
/* Eigenvector-local part of the above for the eigenvectors [row0, row0 + nrows) (multiples of 128 when CEO is
 * given): VtF, rhoVar, snr, invEv on that range (entries outside are not written) and
 * update_partial[n] = sum_{k in range} V[:,k] invEv_k reg_k VtF_k.  Ranks of a multi-GPU solve each take a slice and
 * sum the partial updates with one all-reduce; vmcpde_solve_scalars then gives residual and TDVP error. */
int vmcpde_solve_tail_range(const double* ev, const double* VT, int32_t n, int32_t ld, const double* F,
                            const double* CEO, double n_glob, double svdTol, double snrTol, int32_t useSNR,
                            int32_t row0, int32_t nrows, double* VtF, double* rhoVar, double* snr, double* invEv,
                            double* update_partial, void* workspace, size_t workspace_bytes,
                            vmcpde_stream stream);
/* Blocked Cholesky solve S x = F for a shifted, positive definite S (diagonalShift > 0).  S is overwritten
 * by its lower factor.  *info (device int, zero on entry) = 1 + index of the first non-positive pivot. */
int vmcpde_chol_solve(double* S, int32_t n, int32_t ld, const double* F, double* x, int32_t* info,
                      vmcpde_stream stream);
/* scalars = {||S u - F|| / ||F||, 1 + (u^T S0 u - 2 F^T u)/meanE2} for a given update (tdvp.py:93-94);
 * work2n: 2n doubles of scratch. */
int vmcpde_solve_scalars(const double* S, const double* S0, int32_t n, int32_t ld, const double* F,
                         const double* update, double meanE2, double* scalars, double* work2n,
                         vmcpde_stream stream);

/* ---- (5) observables (tdvp.py:143-162) --------------------------------------------------------- */
int vmcpde_observables_workspace_bytes(int32_t d, size_t* bytes);
/* first[0..d) += sum x; first[d] += sum logp; first[d+1] = max(first[d+1], max E_loc) */
int vmcpde_obs_first(const double* x, const double* logp, const double* eloc, int64_t n, int32_t d,
                     double* first, void* ws, vmcpde_stream stream);
/* central[0..d*d) += sum dx dx^T; then d entries each of sum dx^3, dx^4, dx^5, dx^6, dx = x - mean */
int vmcpde_obs_central(const double* x, int64_t n, int32_t d, const double* mean, double* central,
                       void* ws, vmcpde_stream stream);
/* uniform points in the ball of the given radius, from normal(key) and uniform(key) of the SAME key
 * (tdvp.py:154-155): out[i] = radius * xi_i/|xi_i| * u_i^(1/d) for global indices [first, first+n) */
int vmcpde_ball_points(uint32_t key0, uint32_t key1, int64_t first, int64_t n, int64_t n_total, int32_t d,
                       double radius, double* out, vmcpde_stream stream);
/* out[0] += sum_i exp(logp[i]) */
int vmcpde_sum_exp(const double* logp, int64_t n, double* out, void* ws, vmcpde_stream stream);

/* ---- (6) particle reference dynamics (exact_dyn.py) ----------------------------------------------- */
/* One step of exact_dyn.integrate (exact_dyn.py:70-84) for n particles (coords [n][d], in place): the 4-stage stochastic
 * scheme of integrate_single_coord with per-particle keys split(key, n)[i] -> split(., 4) and normal(key, (d,)) noise in
 * JAX's counter layout.  update: 0 = update_fun_phaseSpace, 1 = update_fun_Diff; field: 0 = _velocity_field_hamiltonian
 * (uncoupled), 1 = _velocity_field_fluiddynpaper, -1 = none.  eq supplies D, m, omega, lam, T, gamma, t. */
int vmcpde_particles_step(double* coords, int64_t n, int32_t d, double dt, int32_t update, int32_t field,
                          const vmcpde_equation* eq, uint32_t key0, uint32_t key1, vmcpde_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VMCPDE_H_ */
