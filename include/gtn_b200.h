/*
 * gtn_b200.h -- C ABI of libgtn_b200.so, the sm_100a kernels behind grassmanntn_b200.
 *
 * The reference (ayosprakob/grassmanntn) is pure Python/numpy and has no FFI boundary; its
 * "operator interface" is the numpy / opt_einsum / LAPACK call sites its Grassmann algebra
 * bottoms out in (SURVEY.md section 1, table "L0 call").  Each entry point below replaces one
 * family of those call sites; the citation says which (file:line relative to the reference
 * repository root).  All pointers are DEVICE pointers unless the name ends in _host, all sizes
 * are element counts, every call takes the CUDA stream to launch on, returns 0 on success or a
 * negative gtn_status / positive cudaError_t, never throws and never allocates (the one workspace-taking call,
 * gtn_sector_svd_trunc, has an explicit size query: gtn_workspace_bytes).
 *
 * Data types: GTN_F64 = IEEE double, GTN_C128 = interleaved (re, im) doubles (numpy complex128,
 * torch.complex128).
 */
#ifndef GTN_B200_H
#define GTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTN_F64 0
#define GTN_C128 1

#define GTN_MAX_SUPER 8 /* super-axes per sign-permute job (see gtn_permute_job) */

typedef enum {
  GTN_OK = 0,
  GTN_ERR_BAD_ARG = -1,
  GTN_ERR_UNSUPPORTED = -2,
  GTN_ERR_NOT_CONVERGED = -3
} gtn_status;

/* ------------------------------------------------------------------------------------------
 * Fused sign + permute (+ encoder gather, + conjugate, + scale).
 *
 * Replaces: the sign tensors S1/S3 that einsum_ds materialises element by element and feeds
 * to opt_einsum (__init__.py:1962-1999, :2088-2126, :2295) whenever they only re-order data;
 * dense.switch_format / switch_encoder / switch_parity (__init__.py:1011-1154: `dat*v[ss]`,
 * `dat.take(...)`); join_legs / split_legs (__init__.py:3015, :3027, :3036-3050, :3112-3157);
 * hconjugate's conj-transpose (__init__.py:5428); block <-> dense conversion
 * (__init__.py:311-337, :701-725) and join/split_legs_block's element copies (:3582-3586).
 *
 *   out[out_base + sum_X out_off_X(i_X)] = scale * (+-1) * maybe_conj( in[in_base + sum_X in_off_X(i_X)] )
 *
 * The tensor is presented as <= GTN_MAX_SUPER "super-axes" (one or several fused tensor legs).
 * Every super-axis X has a table of gtn_axis_entry, one per index value i_X: the element
 * offsets on both sides (so strides, diagonals, encoder gathers and block interleaves are all
 * just tables), the Grassmann-parity bits P of the legs it fuses and M = Qsym*P, the GF(2)
 * image of P under the symmetric pair matrix of the sign program.  The sign exponent is the
 * quadratic form
 *     e = const ^ XOR_X intra_X ^ XOR_{X<Y} parity(P_Y & M_X)
 * i.e. exactly (-1)^{sum_a alpha_a p_a + beta_a q_a + sum_{a<b} Q_ab p_a p_b} over the legs'
 * parities p_a = popcount(index)&1 and sigma bits q_a = (popcount(index)>>1)&1, which is what
 * param.gparity / param.sgn (param.py:61-73) and relative_sign (__init__.py:1553-1591) evaluate
 * element by element in the reference.  Sign tensors are never materialised.
 *
 * Super-axis 0 is contiguous on the INPUT side (tile rows are read along it), super-axis 1 is
 * contiguous on the OUTPUT side (tile rows are written along it); the kernel transposes
 * 32x32 tiles through shared memory.  If `transpose` is 0 both sides are contiguous along
 * super-axis 0 and the tile goes straight through registers.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t in_off;  /* element offset contributed on the input side  */
  int64_t out_off; /* element offset contributed on the output side */
  uint32_t P;      /* bits 0..27: parity bits of the fused legs; bit 31: intra-axis sign exponent */
  uint32_t M;      /* bits 0..27: Qsym * P (GF(2)) */
} gtn_axis_entry;

typedef struct {
  int64_t in_base;                 /* element offset of this job inside `in`  */
  int64_t out_base;                /* element offset of this job inside `out` */
  int64_t table_start[GTN_MAX_SUPER]; /* first entry of super-axis X inside `entries` */
  int32_t size[GTN_MAX_SUPER];     /* extent of super-axis X (unused ones = 1) */
  int32_t nsuper;                  /* number of super-axes in use (>= 1) */
  int32_t const_exp;               /* constant sign exponent (0/1) */
  int32_t conj;                    /* 1: complex-conjugate the element */
  int32_t transpose;               /* 1: shared-memory tile transpose between axis 0 and axis 1 */
  int64_t ntiles;                  /* number of 32x32 tiles (CTAs) of this job */
} gtn_permute_job;

/* Launch `njobs` jobs in ONE grid (grid.y = job, grid.x = max_tiles). `jobs` and `entries` are
 * device arrays (built once per plan by the host and cached).  scale_re/scale_im multiply
 * every element (scale_im ignored for GTN_F64). */
int gtn_sign_permute(const void* in, void* out, int dtype, const gtn_permute_job* jobs,
                     const gtn_axis_entry* entries, int njobs, int64_t max_tiles,
                     double scale_re, double scale_im, void* stream);

/* ------------------------------------------------------------------------------------------
 * Grouped GEMM on the FP64 tensor cores (DMMA m8n8k4), row-major operands.
 *
 * Replaces: oe.contract(..) of the two data operands in einsum_ds (__init__.py:2295) and the
 * per-parity-block np.einsum of einsum_block (__init__.py:2781, :2928).
 *
 *   for every group g:  C_g[M x N] (ldc) = alpha_g * A_g[M x K] (lda) * B_g[K x N] (ldb) + beta_g * C_g
 *
 * offsets are in elements from the base pointers; a group may be repeated `batch` times with
 * strides (batch index is the slowest loop).  GTN_C128 is computed as four real DMMA products
 * per tile with the complex arithmetic expanded in registers.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t a_off, b_off, c_off;
  int64_t lda, ldb, ldc;
  int64_t batch_stride_a, batch_stride_b, batch_stride_c;
  int32_t m, n, k, batch;
  double alpha; /* real scalar (block sign S1*S3 = +-1 in einsum_block, __init__.py:2761) */
  double beta;  /* 0 or 1 */
  int64_t tile_start; /* prefix sum of CTA tiles (filled by gtn_gemm_plan_host) */
  int32_t flags;      /* GTN_GEMM_B_CONJ_TRANS: B_g is given as its conjugate transpose, i.e. the kernel reads
                         B_g[k][n] = conj(Bsrc[n * ldb + k]) from the row-major N x K array at b_off (Gram matrices
                         X X^H and products with W^H without materialising the transposed copy) */
  int32_t reserved;   /* TMA configurations: `raster` -- the tile grid of the group is walked in bands of this many row
                         tiles, column-major inside a band (0 = 1); ignored by the cp.async configurations */
} gtn_gemm_group;
#define GTN_GEMM_B_CONJ_TRANS 1

/* Fills tile_start for all groups on the HOST array and returns the total CTA count.
 * config 0: 64x64 CTA tiles (large sector GEMMs); config 1: 32x32 tiles with a 4x deeper K step
 * (skinny l x p x q products of the randomized subspace iteration); config 3: as 1, and EVERY group's B is
 * given as its conjugate transpose (all groups carry GTN_GEMM_B_CONJ_TRANS).  Same config at launch. */
int64_t gtn_gemm_plan_host(gtn_gemm_group* groups_host, int ngroups, int dtype, int config);

int gtn_grouped_gemm(const void* A, const void* B, void* C, int dtype,
                     const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles,
                     int config, void* stream);

/* TMA-staged variant of gtn_grouped_gemm (north_star: "FP64 DMMA tensor-core tiles staged by TMA"): one producer
 * warp feeds a 4-stage shared-memory ring with cp.async.bulk.tensor loads (SASS UTMALDG) completing on mbarriers,
 * the consumer warps (32x32 warp tiles of DMMA.8x8x4) never compute a load address; tiles are dense 128-byte rows in
 * the hardware SWIZZLE_128B pattern; out-of-range parts of a tile are zero-filled by the TMA unit from the true
 * extents in the tensor maps.  config 4: 64x64 CTA tiles, config 12: 128x64 (same values for gtn_gemm_plan_host).
 * The tensor maps (one per group and operand) are encoded on the host at every call from `groups_host` -- the same
 * array as groups_dev -- and travel as a kernel parameter, so ngroups <= GTN_TMA_MAX_GROUPS.  Requires 16-byte aligned
 * operand sub-matrices (always true for GTN_C128; even offsets / leading dimensions for GTN_F64) and no
 * GTN_GEMM_B_CONJ_TRANS: gtn_gemm_tma_check returns 1 when a group list qualifies.  Returns GTN_ERR_UNSUPPORTED when
 * the driver does not export cuTensorMapEncodeTiled or an operand is misaligned.
 * Replaces the same reference call sites as gtn_grouped_gemm (__init__.py:2295, :2781, :2928). */
#define GTN_TMA_MAX_GROUPS 32
int gtn_gemm_tma_check(const gtn_gemm_group* groups_host, int ngroups, int dtype);
int gtn_grouped_gemm_tma(const void* A, const void* B, void* C, int dtype, const gtn_gemm_group* groups_host,
                         const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles, int config,
                         void* stream);

/* Fused GEMM + all-gather for the multi-GPU output-tile sharding: same product as
 * gtn_grouped_gemm (64x64 tiles, beta = 0), but every result element is stored at the same offset into
 * EACH of the npeers output buffers C_peers[0..npeers) -- all ranks' copies of C mapped into this process
 * (NVLink peer memory, e.g. torch symmetric memory), own copy included.  C_peers is a HOST array of
 * device pointers.  The caller brackets the launch with cross-rank barriers. */
#define GTN_MAX_PEERS 8
int gtn_grouped_gemm_bcast(const void* A, const void* B, void* const* C_peers, int npeers, int dtype,
                           const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched one-sided Jacobi SVD (Hestenes) of the parity-sector matrices.
 *
 * Replaces: np.linalg.svd(M, full_matrices=False) in SortedSVD (__init__.py:3932) and
 * SortedEig (__init__.py:4323), i.e. LAPACK gesdd, for the E/O sectors of BlockSVD/BlockEig
 * (:4002-4003, :4394-4395) and decompose_block (:5083-5088).
 *
 * W_b (p_b x q_b row-major, ld = q_b, p_b <= q_b) is overwritten: its rows are rotated until
 * mutually orthogonal; Z_b (p_b x p_b, initialised to identity by the caller or by
 * gtn_jacobi_init) accumulates the rotations.  After convergence  W0 = Z^H * diag(s) * Vh with
 * s_i = ||row_i(W)|| and Vh = rows of W normalised.  gtn_jacobi_finish sorts by descending s,
 * writes s (double), U = (Z^H permuted) (p x p) and Vh (p x q) into caller buffers.
 * The sweep loop runs on the host side of the ABI: gtn_jacobi_sweep launches the p-1 rounds of
 * one sweep for the whole batch and returns; `offdiag` (device double[batch]) receives
 * max |<w_i,w_j>|^2 / (|w_i|^2 |w_j|^2) over the pairs that were rotated in that sweep (0 = converged).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t w_off; /* element offset of W_b inside W */
  int64_t z_off; /* element offset of Z_b inside Z */
  int32_t p, q;
} gtn_svd_problem;

/* rownorm2_dev: double[sum p_b] (problem b at rn_off_dev[b]) receives the squared row norms,
 * fro2_dev: double[2 * nprob] the largest squared row norm seen so far (a lower bound of s_0^2), double buffered by
 * round parity so that a round only reads what the previous round completed (bitwise reproducible results); both
 * are maintained by the sweeps and let a CTA skip a row pair without reading it when one of the
 * rows has decayed below 2e-15 * s_0 (rank-deficient sectors: such rows are discarded by the
 * reference's rank rule s_i/s_0 > 1e-14 anyway). */
int gtn_jacobi_init(const void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev, int nprob,
                    int max_p, double* rownorm2_dev, double* fro2_dev, const int64_t* rn_off_dev,
                    void* stream);
int gtn_jacobi_sweep(void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev, int nprob,
                     int max_p, int max_q, double tol, double* offdiag_dev, double* rownorm2_dev,
                     const double* fro2_dev, const int64_t* rn_off_dev, void* stream);
/* Whole sweep loop in ONE cooperative launch (grid-wide barriers between rounds, convergence decided on
 * the device) for batches whose (max_p/2 x nprob) CTAs are all co-resident; returns
 * GTN_ERR_UNSUPPORTED otherwise (use gtn_jacobi_sweep).  offdiag2_dev: double[2*nprob] scratch,
 * sweeps_dev: int32[4] = {sweeps executed, converged flag, barrier counter, unused}. */
int gtn_jacobi_persistent(void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev, int nprob,
                          int max_p, double tol, double* offdiag2_dev, double* rownorm2_dev,
                          const double* fro2_dev, const int64_t* rn_off_dev, int max_sweeps,
                          int32_t* sweeps_dev, void* stream, double early_stop);
/* early_stop = 0: stop after the first sweep in which no pair exceeded `tol` (a confirming sweep).
 * early_stop > 0: also stop after a sweep whose rotated pairs all had |<w_i,w_j>| / (|w_i||w_j|) <= early_stop --
 * the iteration converges quadratically, so that sweep leaves at most ~early_stop^2 / (relative gap) behind.
 * Used (1e-10) for the projected matrix of the truncated SVD, whose result is checked by the residual
 * certificate anyway; the full SVD of a sector passes 0. */

/* s_out: double[sum p_b] at offsets s_off_b (descending); U_out (p x p row-major) at u_off;
 * Vh_out: rows of W normalised and permuted, at w_off.  order_dev: int32 [sum p_b] receives the
 * permutation; norm_scratch_dev: double [sum p_b] scratch. */
typedef struct {
  int64_t s_off, u_off;
} gtn_svd_out;
int gtn_jacobi_finish(const void* W, const void* Z, void* U_out, void* Vh_out, double* s_out,
                      int dtype, const gtn_svd_problem* probs_dev, const gtn_svd_out* outs_dev,
                      int32_t* order_dev, double* norm_scratch_dev, int nprob, int max_p,
                      int max_q, void* stream);

/* Whitening transform from a small Hermitian Gram matrix (two-sided Jacobi, one CTA per matrix, all
 * in shared memory; n_b <= 80).  For every problem b: G_b (n_b x n_b, row-major) = E L E^H and
 * T_b = L^{-1/2} E^H with the rows of eigenvalues <= rel_thr * L_max zeroed (kept_dev[b] = number
 * retained, evals_dev[e_off[b] + i] = L_i), so that T_b Y orthonormalises the rows of Y when
 * G_b = Y Y^H.  Used by the randomized subspace iteration of the truncated sector SVD -- the
 * "Gram-matrix / randomized-projection" replacement of LAPACK's full gesdd in reference SortedSVD
 * (__init__.py:3932) when only the first `cutoff` triplets survive (:3943-3948). */
int gtn_small_eigh_whiten(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                          const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n,
                          double rel_thr, int32_t* kept_dev, double* evals_dev,
                          const int64_t* e_off_dev, void* stream);

/* Same contract as gtn_small_eigh_whiten with a diagonally pivoted Cholesky factorisation
 * P^T G P = L L^H (stopped at numerical rank r: remaining diagonal <= rel_thr * first pivot):
 * T = [L_r^{-1} 0] P^T, rows >= r zero, kept_dev[b] = r.  Three block barriers per pivot.
 * max_n <= 80: everything in shared memory (scratch may be NULL); 80 < max_n <= 512: the working
 * copies live in `scratch` (complex128 elements, nprob * gtn_chol_whiten_scratch_elems(max_n)).
 * nsplit >= 1: G_b is the SUM of nsplit consecutive n_b x n_b slices at g_off (split-K partial Gram
 * matrices of the grouped GEMM, summed in a fixed order while loading). */
int64_t gtn_chol_whiten_scratch_elems(int max_n);
int gtn_chol_whiten(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                    const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n, int nsplit,
                    double rel_thr, int32_t* kept_dev, void* scratch, void* stream);

/* G_b <- G_b + rel_shift * trace(G_b) * I for the Gram matrices handed to gtn_chol_whiten (same g_off / n / nsplit
 * arguments; the trace is taken over the sum of the nsplit slices, the shift goes onto slice 0): the shift of a SHIFTED
 * Cholesky QR.  A panel whose singular values span more than sqrt(1/eps) has a numerically indefinite Gram matrix --
 * plain (pivoted) Cholesky drops every direction below 3e-7 s_0; with the shift the factorisation exists, T X has
 * condition ~ 1/sqrt(rel_shift) and keeps all directions, and further (shifted, then plain) passes finish the
 * orthonormalisation (shifted CholeskyQR3).  Used by the robust mode of the truncated sector SVD (spectra steeper than
 * a Gram matrix resolves: the third decomposition of an ATRG step, reference gauge2d.py:1843). */
int gtn_gram_shift(void* G, int dtype, const int64_t* g_off_dev, const int32_t* n_dev, int nprob, int nsplit,
                   double rel_shift, void* stream);

/* Pre-rotation for the one-sided Jacobi SVD of a short-and-wide matrix B (n_b <= 80 rows): from the
 * Gram matrix G_b = B B^H compute a UNITARY T_b (n_b x n_b) such that the rows of T_b B are orthogonal
 * up to the accuracy a Gram matrix allows (pivoted Cholesky G = P L L^H P^T, one-sided Jacobi on the
 * rows of L in shared memory, T = Z P^T; one CTA per matrix).  gtn_jacobi_persistent started from
 * T_b B then needs 2-3 sweeps instead of 7-8, and the singular values it returns never go through G.
 * sweeps_dev: int32[nprob] (may be NULL) receives the in-kernel sweep counts. */
int gtn_gram_rotate(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                    const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n, int nsplit,
                    double rel_thr, double tol, int max_sweeps, int32_t* sweeps_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * One-call truncated decomposition of a batch of parity-sector matrices.
 *
 * Replaces, per batch: np.linalg.svd (LAPACK gesdd) + the rank rule s_i / (s_0 + numer_cutoff) > numer_cutoff + the
 * cut to the first k triplets in SortedSVD (__init__.py:3931-3951), BlockSVD (:3998-4003) and decompose_block
 * (:5083-5088); gtn_sector_eigh_trunc adds SortedEig's signed eigenvalues (:4340-4341) for Hermitian sectors.
 *
 * M[b]: DEVICE pointer of the row-major m[b] x n[b] sector matrix (not modified); k[b] <= min(m[b], n[b]) triplets
 * wanted.  Outputs: U_out[b] (m[b] x k[b], row-major), Vh_out[b] (k[b] x n[b]) on the device; S_host (sum_b k[b]
 * doubles, problem b after problem b-1, descending) and rank_host[b] = number of those values that pass the rank rule
 * on the HOST -- the call synchronises `stream` once per certificate check (the decision to stop iterating is the
 * only host-dependent step).  M, U_out, Vh_out are HOST arrays of nb device pointers.
 * The top-k triplets come from randomized subspace iteration on all matrices at once (grouped DMMA GEMMs, pivoted
 * Cholesky whitening, one-sided Jacobi SVD of the l x n projected matrices) and are returned only with a certificate:
 * residuals ||M^H u_i - s_i v_i|| <= 1e-11 s_0 for the kept triplets, no discarded Ritz value + residual above the
 * smallest kept value, and -- for numerically rank-deficient sectors -- a norm bound of the deflated matrix
 * M - U_D U_D^H M below numer_cutoff * s_0.  Returns GTN_OK, or GTN_ERR_NOT_CONVERGED when the certificate cannot be
 * met (flat spectrum at the cut: the caller runs the full SVD, gtn_jacobi_*), or an error.
 * Nothing is allocated: `workspace` (256-byte aligned device memory) must hold gtn_workspace_bytes(op, ...) bytes.
 * info (may be NULL): in  start_iters = iterations to run before the first certificate check (the count the last
 * accepted run of this shape needed; 0 = check right after the range finder), rate = convergence factor measured by
 * an earlier call (0 = unknown); out  iters, checks, sweeps, worst (largest kept residual / s_0), rate.
 * ------------------------------------------------------------------------------------------ */
#define GTN_OP_SECTOR_SVD_TRUNC 0
#define GTN_OP_SECTOR_EIGH_TRUNC 1
typedef struct {
  int32_t start_iters;
  int32_t iters;
  int32_t checks;
  int32_t sweeps;
  int32_t launches; /* out: kernels launched by the call */
  int32_t robust;   /* in: 1 = shifted Cholesky QR orthonormalisations (gtn_gram_shift): spectra that span more than 3e-7
                       inside the cut, where the plain run returns GTN_ERR_NOT_CONVERGED */
  double worst;
  double rate;
} gtn_svd_info;
int64_t gtn_workspace_bytes(int op, int dtype, int nb, const int64_t* m, const int64_t* n, const int32_t* k);
int gtn_sector_svd_trunc(const void* const* M, const int64_t* m, const int64_t* n, int nb, int dtype,
                         const int32_t* k, double numer_cutoff, void* const* U_out, double* S_host,
                         void* const* Vh_out, int32_t* rank_host, void* workspace, int64_t workspace_bytes,
                         gtn_svd_info* info, void* stream);
/* lam_host: interleaved (re, im) doubles, sum_b k[b] entries: lam_c = sum_i s_i (Vh U)_ic. */
int gtn_sector_eigh_trunc(const void* const* M, const int64_t* m, const int64_t* n, int nb, int dtype,
                          const int32_t* k, double numer_cutoff, void* const* U_out, double* S_host,
                          void* const* Vh_out, double* lam_host, int32_t* rank_host, void* workspace,
                          int64_t workspace_bytes, gtn_svd_info* info, void* stream);

/* Diagnostic: clock64 stamps of CTA 0 of the last gtn_chol_whiten ([0..3]: start, factorised,
 * inverted, stored) and gtn_gram_rotate ([4..7]: start, factorised, rotated, stored) launches.
 * Synchronises the device; host_out8 is a HOST array of 8 int64. */
int gtn_debug_phase_clocks(long long* host_out8);

/* ------------------------------------------------------------------------------------------
 * Small element-wise / reduction helpers.
 * ------------------------------------------------------------------------------------------ */
/* out[0] = sum |x_i|^2 (double). Replaces np.linalg.norm in dense.norm / block.norm
 * (__init__.py:891-893, :432-437).  `out` must be zeroed by the caller (or pass zero_first=1). */
int gtn_sumsq(const void* x, int64_t n, int dtype, double* out_dev, int zero_first, void* stream);

/* out[0] += sum |x_i| (complex modulus for GTN_C128).  The reference's Grassmann-evenness tests are L1 means over
 * the odd-parity entries: BlockSVD / BlockEig (__init__.py:3977-3983, :4369-4375) and decompose_block reject a
 * tensor whose odd entries have mean |x| > numer_cutoff. */
int gtn_sumabs(const void* x, int64_t n, int dtype, double* out_dev, int zero_first, void* stream);

/* y[r] = sum_c x[r*cols + c]  (row sums; the summed tail of a trace-type einsum such as
 * 'IJIJ' -- the final reduction of oe.contract at __init__.py:2295 when nothing is left to
 * multiply).  y has `rows` elements of the same dtype. */
int gtn_rowsum(const void* x, void* y, int64_t rows, int64_t cols, int dtype, void* stream);

/* y[i] = sum_{s < nslices} x[s * n + i], slices added in index order (deterministic): completes a split-K grouped
 * GEMM whose slices were written to consecutive copies of the output (Gram-type contractions with a handful of output
 * tiles and a very long contracted range, e.g. the environment matrices of hotrg3dz, gauge2d.py:1964-1992; the
 * reference contracts them inside oe.contract, __init__.py:2295).  x and y 16-byte aligned; GTN_F64 needs an even n when
 * nslices > 1 (every slice starts 16-byte aligned). */
int gtn_sum_slices(const void* x, void* y, int64_t n, int nslices, int dtype, void* stream);

/* out = sum_i a_i * b_i (no conjugation; out has one element of `dtype`).  The fully contracted
 * einsum ('ijkl,klij', 'IJIK,iKiJ': the trace-preservation checks of gauge2d.py:1738, :1856) after
 * the operands were packed with their signs -- the degenerate 1 x K x 1 case of oe.contract
 * (__init__.py:2295).  partial_dev: double[2*nparts] scratch; deterministic two-stage reduction. */
int gtn_dot(const void* a, const void* b, int64_t n, int dtype, void* out, double* partial_dev, int nparts,
            void* stream);

/* y[r] = sum_c |x[r*cols + c]|^2 (double): residual norms of the truncated-SVD certificate. */
int gtn_row_sumsq(const void* x, double* y, int64_t rows, int64_t cols, int dtype, void* stream);

/* x_i <- |x_i| > rcond ? x_i^p : 0   (power_ds, __init__.py:6059-6069; fixed power_block :6071). */
int gtn_pow_rcond(void* x, int64_t n, int dtype, double p, double rcond, void* stream);

/* x_i <- x_i * (s_re + i s_im)   (T.data/Tnorm at gauge2d.py:1748, :1864). */
int gtn_scale(void* x, int64_t n, int dtype, double s_re, double s_im, void* stream);

/* out[0] = max |x_i| over the entries whose Grassmann parity is odd, for a packed sector check
 * (BlockSVD's evenness test __init__.py:3977-3983): x is rows x cols, entry (r,c) counts when
 * ((r ^ c) & 1). */
int gtn_odd_checker(const void* x, int64_t rows, int64_t cols, int dtype, double* out_dev,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Collectives of the sharded coarse-graining step: thin wrappers of NCCL over NVLink /
 * NVSwitch, bound at run time (dlopen of libnccl.so.2; gtn_comm_available() = 0 and GTN_ERR_UNSUPPORTED without it).
 * The reference has no distributed mode; the sharded step (grassmanntn_b200/sharded.py) all-reduces l x p sketch
 * panels and l x l Gram matrices, all-gathers the isometries, broadcasts the owners' small SVDs.
 * One communicator per process (one process per GPU): rank 0 calls gtn_comm_unique_id, the 128-byte id reaches the
 * other ranks by the host's own means, every rank calls gtn_comm_init.  Counts are ELEMENTS of `dtype` (GTN_C128 is
 * sent as two doubles per element; only the sum is defined on it).  op: 0 sum, 1 max, 2 min.  All calls are enqueued on
 * `stream`; buffers are device memory.  NCCL failures are returned as 1000 + ncclResult_t.
 * ------------------------------------------------------------------------------------------ */
#define GTN_COMM_ID_BYTES 128
int gtn_comm_available(void);
int gtn_comm_unique_id(void* id_out_128_bytes);
int gtn_comm_init(const void* id_128_bytes, int rank, int world, void** comm_out);
int gtn_comm_destroy(void* comm);
int gtn_allreduce(void* comm, void* buf, int64_t count, int dtype, int op, void* stream);         /* in place */
int gtn_allgather(void* comm, const void* send, void* recv, int64_t count_per_rank, int dtype, void* stream);
int gtn_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream);        /* in place */

/* Device-side barrier over peer-mapped memory (the fused GEMM + all-reduce of the sharded step): flag_peers is a HOST
 * array of npeers device pointers, entry q = rank q's flag array (>= npeers uint64, zero-initialised, mapped into this
 * process).  Every rank passes the same monotonically increasing `epoch` (> 0).  Enqueued on `stream`. */
int gtn_peer_barrier(void* const* flag_peers, int rank, int npeers, uint64_t epoch, void* stream);

/* Library / device introspection (no device work). */
int gtn_version(void);
const char* gtn_build_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* GTN_B200_H */
