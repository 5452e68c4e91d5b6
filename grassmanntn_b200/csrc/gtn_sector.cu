// gtn_sector.cu -- one-call truncated SVD / eigen-decomposition of a batch of parity-sector matrices, sm_100a.
//
// Replaces, in ONE C-ABI call per batch: np.linalg.svd (LAPACK gesdd) + the rank rule + the cut to the first
// `cutoff` triplets in reference SortedSVD (__init__.py:3931-3951) / BlockSVD (:3998-4003) / decompose_block
// (:5083-5088), and SortedEig / BlockEig (:4314-4395) for the Hermitian case.  SURVEY.md section 8(b) names the
// entry points: gtn_sector_svd_trunc, gtn_sector_eigh_trunc, gtn_workspace_bytes.
//
// The driver loop that used to live in the Python host (_engine.truncated_svd_batch) runs here, on the host side
// of the library, and only enqueues the library's own kernels on the caller's stream:
//
//     Yh = G Wh ; Qh = orth(Yh) ; [ Zh = Qh W ; Ph = orth(Zh) ; Yh = Ph Wh ; Qh = orth(Yh) ]*      (DMMA GEMMs)
//     orth(X) = T X with T from the pivoted Cholesky of the split-K Gram matrix X X^H            (gtn_chol_whiten)
//     B = Qh W (l x q) ; B = Ub S Vh (one-sided Jacobi, persistent kernel) ; Uh = Ub^H Qh
//     certificate rows  Vh W^H - S Uh  ->  residual norms                                        (gtn_row_sumsq)
//
// with ONE host read-back per certificate check (the decision to stop is the only host-dependent step).  The result
// is accepted only with the certificate (residuals of the kept triplets <= 1e-11 s_0, no discarded Ritz value can
// reach above the smallest kept one) and, for numerically rank-deficient sectors, the rank certificate (norm bound of
// the deflated matrix); otherwise GTN_ERR_NOT_CONVERGED is returned and the caller runs the full Jacobi SVD
// (gtn_jacobi_*).  Nothing is allocated: every buffer is carved out of the caller's workspace
// (gtn_workspace_bytes).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gtn_b200.h"

namespace {

constexpr double kTruncTol = 1e-11;     // certificate: max_i ||W^H u_i - s_i v_i|| <= kTruncTol * s_0
constexpr double kRankNoise = 1e-13;    // singular values <= kRankNoise * s_0 are decided by the rank certificate
constexpr double kJacobiTol = 4e-15;
constexpr double kEarlyStop = 1e-10;
constexpr int kJacobiMaxSweeps = 60;
constexpr int kPersistentMaxRows = 160;
constexpr int kLmax = 320;
constexpr int kMaxIters = 20;
constexpr int64_t kMetaBytes = 1 << 20;  // ring of device copies of the GEMM group lists of one schedule pass

struct c128 { double re, im; };

// ---- small kernels ---------------------------------------------------------------------------------------------
// dst (c x r) = op(src[:r, :c]) transposed, op = conj for CONJ; 32x32 tiles through padded shared memory
template <typename T, bool CONJ>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, int64_t ld_src, T* __restrict__ dst,
                                                        int64_t ld_dst, int64_t r, int64_t c) {
  __shared__ T tile[32][33];
  const int64_t r0 = int64_t(blockIdx.y) * 32, c0 = int64_t(blockIdx.x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t rr = r0 + ty + 8 * k, cc = c0 + tx;
    if (rr < r && cc < c) tile[ty + 8 * k][tx] = src[rr * ld_src + cc];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t cc = c0 + ty + 8 * k, rr = r0 + tx;
    if (rr < r && cc < c) {
      T v = tile[tx][ty + 8 * k];
      if constexpr (CONJ && sizeof(T) == 16) v.im = -v.im;
      dst[cc * ld_dst + rr] = v;
    }
  }
}

// counter-based standard normals (splitmix64 + Box-Muller): the sketch matrix is a pure function of (seed, index)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void randn_kernel(double* __restrict__ x, int64_t n_pairs, uint64_t seed) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n_pairs; i += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t a = mix64(seed ^ (uint64_t(i) * 2 + 1)), b = mix64(seed + 0x632BE59BD9B4E019ull + uint64_t(i) * 2);
    const double u1 = (double(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = double(b >> 11) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    x[2 * i] = rad * cs;
    x[2 * i + 1] = rad * sn;
  }
}

// D (l x l, zeroed elsewhere) gets diag(s)
template <typename T>
__global__ void set_diag_kernel(T* __restrict__ D, const double* __restrict__ s, int l) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < l) {
    T v;
    if constexpr (sizeof(T) == 16) { v.re = s[i]; v.im = 0.0; } else { v = s[i]; }
    D[int64_t(i) * l + i] = v;
  }
}

__global__ void pack_out_kernel(double* __restrict__ out, const double* __restrict__ s, const double* __restrict__ res2,
                                const int32_t* __restrict__ kept, const int32_t* __restrict__ sw, int sumL, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < sumL) { out[i] = s[i]; out[sumL + i] = res2[i]; }
  if (i < nb) out[2 * sumL + i] = double(kept[i]);
  if (i < 2) out[2 * sumL + nb + i] = double(sw[i]);
}

// ---- host side -------------------------------------------------------------------------------------------------
inline int subspace_rows(int k) {
  if (k <= 16) return (2 * k + 16 <= 32) ? 32 : 48;
  return 64 * ((2 * k + 63) / 64);
}

struct Mat { int64_t off, r, c; };       // element offset inside the workspace, rows, cols (row-major, ld = c)

struct Layout {
  int nb = 0, dtype = 0, NS = 1;
  size_t esz = 16;
  std::vector<int64_t> P, Q;
  std::vector<int> K, L;
  std::vector<Mat> W, Wh, G, Yh, Qh, Zh, Ph, B, Vk, Z, Ub, UbH, Uh, U, Xh, D, T1, T2, Cp, Cq, Sp, Sq;
  int64_t elems = 0;                      // element region
  int64_t scratchG = 0, scratchG_elems = 0;   // min(p,q)^2 for the rank certificate (largest problem)
  int sumL = 0, maxL = 0;
  int64_t maxQ = 0;
  std::vector<int> soff;
  // byte offsets of the typed tail
  int64_t o_sdev, o_rn2, o_res2, o_nscr, o_fro2, o_offd, o_out, o_acc, o_kept, o_order, o_sw, o_ndev, o_goff, o_toff,
      o_rnoff, o_probs, o_probs_g, o_probs_r, o_outs, o_chol, o_meta, total_bytes;
  bool prerotate = false;                 // wide subspaces (l > 80): Gram pre-rotation of the projected matrix
  int64_t chol_elems = 0;

  Mat add(int64_t r, int64_t c) {
    Mat m{elems, r, c};
    int64_t n = r * c;
    if (dtype == GTN_F64 && (n & 1)) ++n;             // keep every matrix 16-byte aligned (TMA operands)
    elems += n;
    return m;
  }

  void build(int nb_, int dtype_, const int64_t* m, const int64_t* n, const int32_t* k) {
    nb = nb_; dtype = dtype_;
    esz = dtype == GTN_C128 ? 16 : 8;
    P.assign(m, m + nb); Q.assign(n, n + nb); K.assign(k, k + nb);
    L.resize(nb);
    int64_t minpq = INT64_MAX;
    for (int b = 0; b < nb; ++b) {
      L[b] = (int)std::min<int64_t>(std::min(P[b], Q[b]), std::min(subspace_rows(K[b]), kLmax));
      minpq = std::min(minpq, std::min(P[b], Q[b]));
    }
    NS = minpq >= 256 ? 4 : 1;
    auto each = [&](std::vector<Mat>& v, auto rows, auto cols) {
      v.resize(nb);
      for (int b = 0; b < nb; ++b) v[b] = add(rows(b), cols(b));
    };
    auto p_ = [&](int b) { return P[b]; };
    auto q_ = [&](int b) { return Q[b]; };
    auto l_ = [&](int b) { return (int64_t)L[b]; };
    auto nsl = [&](int b) { return (int64_t)NS * L[b]; };
    each(W, p_, q_); each(Wh, q_, p_);
    each(G, l_, q_); each(Yh, l_, p_); each(Qh, l_, p_);
    each(Zh, l_, q_); each(Ph, l_, q_);
    each(B, l_, q_); each(Vk, l_, q_);
    each(Z, l_, l_); each(Ub, l_, l_); each(UbH, l_, l_);
    each(Uh, l_, p_); each(U, p_, l_); each(Xh, l_, p_);
    each(D, l_, l_); each(T1, nsl, l_); each(T2, l_, l_);
    each(Cp, p_, l_); each(Cq, q_, l_); each(Sp, l_, p_); each(Sq, l_, q_);
    scratchG_elems = 0;
    for (int b = 0; b < nb; ++b) {
      const int64_t mn = std::min(P[b], Q[b]);
      scratchG_elems = std::max(scratchG_elems, mn * mn);
    }
    scratchG = elems;
    elems += scratchG_elems + (scratchG_elems & 1);
    sumL = 0; maxL = 0; maxQ = 0;
    soff.resize(nb);
    for (int b = 0; b < nb; ++b) { soff[b] = sumL; sumL += L[b]; maxL = std::max(maxL, L[b]); maxQ = std::max(maxQ, Q[b]); }
    int64_t o = elems * (int64_t)esz;
    auto take = [&](int64_t bytes) { int64_t at = o; o += (bytes + 255) & ~int64_t(255); return at; };
    o = (o + 255) & ~int64_t(255);
    o_sdev = take(8 * sumL); o_rn2 = take(8 * sumL); o_res2 = take(8 * sumL); o_nscr = take(8 * sumL);
    o_fro2 = take(16 * nb); o_offd = take(16 * nb); o_out = take(8 * (2 * sumL + nb + 2)); o_acc = take(64);
    o_kept = take(8 * nb); o_order = take(4 * sumL); o_sw = take(16); o_ndev = take(4 * nb);
    o_goff = take(8 * nb); o_toff = take(8 * nb); o_rnoff = take(8 * nb);
    o_probs = take(sizeof(gtn_svd_problem) * nb); o_outs = take(sizeof(gtn_svd_out) * nb);
    o_probs_g = take(sizeof(gtn_svd_problem) * nb); o_probs_r = take(sizeof(gtn_svd_problem) * nb);
    prerotate = true;
    for (int b = 0; b < nb; ++b) prerotate = prerotate && L[b] > 80;
    chol_elems = gtn_chol_whiten_scratch_elems(maxL);
    o_chol = take(16 * chol_elems * nb);
    o_meta = take(kMetaBytes);
    total_bytes = o;
  }
};

struct Driver {
  Layout lay;
  char* ws = nullptr;
  cudaStream_t st = nullptr;
  int64_t meta_used = 0;
  int err = 0;
  bool persistent = true;
  int host_sweeps = -1;
  int launches = 0;

  char* elem(const Mat& m) const { return ws + m.off * (int64_t)lay.esz; }
  template <typename T> T* at(int64_t byte_off) const { return reinterpret_cast<T*>(ws + byte_off); }
  void note(int rc) { if (rc != 0 && err == 0) err = rc; }
  void note_cuda() { note((int)cudaGetLastError()); }

  // ---- grouped GEMM launch: configuration chosen like the Python host (estimated rate of each tile family)
  void gemm(std::vector<gtn_gemm_group>& g) {
    if (g.empty() || err) return;
    const int n = (int)g.size();
    struct Cfg { int id; double base; int slots, bm, bn; };
    std::vector<Cfg> cands;
    int64_t min_mn = INT64_MAX, min_k = INT64_MAX, max_m = 0;
    double useful = 0;
    for (auto& x : g) {
      min_mn = std::min<int64_t>(min_mn, std::min(x.m, x.n));
      min_k = std::min<int64_t>(min_k, x.k);
      max_m = std::max<int64_t>(max_m, x.m);
      useful += double(x.m) * x.n * std::max(x.batch, 1);
    }
    if (n <= GTN_TMA_MAX_GROUPS && min_mn >= 64 && min_k >= 16 && gtn_gemm_tma_check(g.data(), n, lay.dtype))
      cands.push_back(max_m >= 128 ? Cfg{12, 0.975, 148, 128, 64} : Cfg{4, 0.975, 296, 64, 64});
    cands.push_back(Cfg{1, 0.91, 296, 32, 32});
    cands.push_back(Cfg{0, 0.95, 296, 64, 64});
    int best = 0;
    double best_eff = -1.0;
    for (auto& c : cands) {
      const int64_t tiles = gtn_gemm_plan_host(g.data(), n, lay.dtype, c.id);
      if (tiles <= 0) continue;
      if ((c.id & 4) && tiles < 37) continue;
      const int64_t waves = (tiles + c.slots - 1) / c.slots;
      const double eff = c.base * (useful / (double(tiles) * c.bm * c.bn)) * (double(tiles) / (double(waves) * c.slots));
      if (eff > best_eff * 1.005) { best = c.id; best_eff = eff; }
    }
    if (best & 4)
      for (auto& x : g) x.reserved = (best & 8) ? 8 : 16;
    const int64_t tiles = gtn_gemm_plan_host(g.data(), n, lay.dtype, best);
    const int64_t bytes = (int64_t)sizeof(gtn_gemm_group) * n;
    if (meta_used + bytes > kMetaBytes) {                       // ring full: everything enqueued so far must have
      cudaStreamSynchronize(st);                                // consumed its tables before they are overwritten
      meta_used = 0;
    }
    gtn_gemm_group* dev = at<gtn_gemm_group>(lay.o_meta + meta_used);
    meta_used += (bytes + 255) & ~int64_t(255);
    note((int)cudaMemcpyAsync(dev, g.data(), bytes, cudaMemcpyHostToDevice, st));
    ++launches;
    if (best & 4)
      note(gtn_grouped_gemm_tma(ws, ws, ws, lay.dtype, g.data(), dev, n, tiles, best, st));
    else
      note(gtn_grouped_gemm(ws, ws, ws, lay.dtype, dev, n, tiles, best, st));
  }

  static gtn_gemm_group group(const Mat& a, const Mat& b, const Mat& c, double alpha = 1.0, double beta = 0.0) {
    gtn_gemm_group g;
    memset(&g, 0, sizeof(g));
    g.a_off = a.off; g.b_off = b.off; g.c_off = c.off;
    g.lda = a.c; g.ldb = b.c; g.ldc = c.c;
    g.m = (int32_t)a.r; g.k = (int32_t)a.c; g.n = (int32_t)b.c;
    g.batch = 1; g.alpha = alpha; g.beta = beta;
    return g;
  }

  void gemm_all(const std::vector<Mat>& a, const std::vector<Mat>& b, const std::vector<Mat>& c, double alpha = 1.0,
                double beta = 0.0) {
    std::vector<gtn_gemm_group> g;
    for (int i = 0; i < lay.nb; ++i) g.push_back(group(a[i], b[i], c[i], alpha, beta));
    gemm(g);
  }

  // dst (c x r) = src (r x c)^H
  void ctranspose(const char* src, int64_t ld_src, char* dst, int64_t ld_dst, int64_t r, int64_t c, bool conj = true) {
    if (err || r <= 0 || c <= 0) return;
    dim3 grid((unsigned)((c + 31) / 32), (unsigned)((r + 31) / 32));
    ++launches;
    if (lay.dtype == GTN_C128) {
      if (conj) transpose_kernel<c128, true><<<grid, 256, 0, st>>>((const c128*)src, ld_src, (c128*)dst, ld_dst, r, c);
      else transpose_kernel<c128, false><<<grid, 256, 0, st>>>((const c128*)src, ld_src, (c128*)dst, ld_dst, r, c);
    } else {
      transpose_kernel<double, false><<<grid, 256, 0, st>>>((const double*)src, ld_src, (double*)dst, ld_dst, r, c);
    }
    note_cuda();
  }
  void ctranspose_all(const std::vector<Mat>& src, const std::vector<Mat>& dst) {
    for (int b = 0; b < lay.nb; ++b) ctranspose(elem(src[b]), src[b].c, elem(dst[b]), dst[b].c, src[b].r, src[b].c);
  }

  void load(const void* const* M) {
    for (int b = 0; b < lay.nb; ++b) load_one(M, b);
  }
  void load_one(const void* const* M, int b) {
    note((int)cudaMemcpyAsync(elem(lay.W[b]), M[b], lay.P[b] * lay.Q[b] * (int64_t)lay.esz, cudaMemcpyDeviceToDevice, st));
    ctranspose(elem(lay.W[b]), lay.Q[b], elem(lay.Wh[b]), lay.P[b], lay.P[b], lay.Q[b]);
  }

  void init_tables() {
    const int nb = lay.nb;
    std::vector<int64_t> goff(nb), toff(nb), rnoff(nb);
    std::vector<int32_t> nd(nb);
    std::vector<gtn_svd_problem> pr(nb), pg(nb), prr(nb);
    std::vector<gtn_svd_out> ou(nb);
    for (int b = 0; b < nb; ++b) {
      goff[b] = lay.T1[b].off; toff[b] = lay.T2[b].off; rnoff[b] = lay.soff[b]; nd[b] = lay.L[b];
      pr[b].w_off = lay.B[b].off; pr[b].z_off = lay.Z[b].off; pr[b].p = lay.L[b]; pr[b].q = (int32_t)lay.Q[b];
      // pre-rotation: Jacobi on the l x l Gram matrix (in T2, rotations into UbH), then on B' = T B (in Sq)
      pg[b].w_off = lay.T2[b].off; pg[b].z_off = lay.UbH[b].off; pg[b].p = lay.L[b]; pg[b].q = lay.L[b];
      prr[b] = pr[b]; prr[b].w_off = lay.Sq[b].off;
      ou[b].s_off = lay.soff[b]; ou[b].u_off = lay.Ub[b].off;
    }
    auto up = [&](int64_t off, const void* src, size_t bytes) {
      note((int)cudaMemcpyAsync(ws + off, src, bytes, cudaMemcpyHostToDevice, st));
    };
    up(lay.o_goff, goff.data(), 8 * nb); up(lay.o_toff, toff.data(), 8 * nb); up(lay.o_rnoff, rnoff.data(), 8 * nb);
    up(lay.o_ndev, nd.data(), 4 * nb);
    up(lay.o_probs, pr.data(), sizeof(gtn_svd_problem) * nb); up(lay.o_outs, ou.data(), sizeof(gtn_svd_out) * nb);
    up(lay.o_probs_g, pg.data(), sizeof(gtn_svd_problem) * nb); up(lay.o_probs_r, prr.data(), sizeof(gtn_svd_problem) * nb);
    for (int b = 0; b < nb; ++b) {
      const int64_t n = (int64_t)lay.L[b] * lay.Q[b] * (lay.dtype == GTN_C128 ? 2 : 1);
      randn_kernel<<<296, 256, 0, st>>>((double*)elem(lay.G[b]), (n + 1) / 2, 0x243F6A8885A308D3ull + 7919ull * b + (uint64_t)n);
      note((int)cudaMemsetAsync(elem(lay.D[b]), 0, (int64_t)lay.L[b] * lay.L[b] * lay.esz, st));
    }
    note_cuda();
    cudaStreamSynchronize(st);                 // (the host vectors above go out of scope)
  }

  void gram(const std::vector<Mat>& cur, const std::vector<Mat>& curH) {
    std::vector<gtn_gemm_group> g;
    for (int b = 0; b < lay.nb; ++b) {
      const int64_t l = cur[b].r, K = cur[b].c, kc = (K + lay.NS - 1) / lay.NS;
      for (int sp = 0; sp < lay.NS; ++sp) {
        const int64_t k0 = std::min<int64_t>(sp * kc, K), kk = std::min<int64_t>(kc, K - k0);
        gtn_gemm_group x;
        memset(&x, 0, sizeof(x));
        x.a_off = cur[b].off + k0; x.b_off = curH[b].off + k0 * l; x.c_off = lay.T1[b].off + sp * l * l;
        x.lda = K; x.ldb = l; x.ldc = l; x.m = (int32_t)l; x.n = (int32_t)l; x.k = (int32_t)kk; x.batch = 1;
        x.alpha = 1.0; x.beta = 0.0;
        if (kk <= 0) { x.k = 1; x.alpha = 0.0; }          // empty slice: C = 0 * (one valid column)
        g.push_back(x);
      }
    }
    gemm(g);
  }

  void whiten(int slot, double rel_thr = 1e-13) {
    if (err) return;
    ++launches;
    note(gtn_chol_whiten(ws, ws, lay.dtype, at<int64_t>(lay.o_goff), at<int64_t>(lay.o_toff), at<int32_t>(lay.o_ndev),
                         lay.nb, lay.maxL, lay.NS, rel_thr, at<int32_t>(lay.o_kept) + slot * lay.nb,
                         lay.chol_elems ? (void*)(ws + lay.o_chol) : nullptr, st));
  }

  bool robust = false;      // shifted Cholesky QR: two passes with 1e-9 trace(G) on the diagonal before the plain ones

  // src is Yh (side p) or Zh (side q): its storage serves as the second intermediate once the first pass has read it
  void orth(const std::vector<Mat>& src, const std::vector<Mat>& dst, bool side_p, int passes) {
    const std::vector<Mat>& C = side_p ? lay.Cp : lay.Cq;
    const std::vector<Mat>& S = side_p ? lay.Sp : lay.Sq;
    const std::vector<Mat>* cur = &src;
    const int shifted = robust ? 2 : 0, total = passes + shifted;
    for (int ps = 0; ps < total; ++ps) {
      ctranspose_all(*cur, C);
      gram(*cur, C);
      if (ps < shifted && !err) {
        ++launches;
        note(gtn_gram_shift(ws, lay.dtype, at<int64_t>(lay.o_goff), at<int32_t>(lay.o_ndev), lay.nb, lay.NS, 1e-9, st));
      }
      whiten(ps == 0 ? 0 : 1, ps < shifted ? 1e-15 : 1e-13);
      const std::vector<Mat>& out = (ps == total - 1) ? dst : ((cur == &S) ? src : S);
      gemm_all(lay.T2, *cur, out);
      cur = &out;
    }
  }

  void start(int passes) {
    gemm_all(lay.G, lay.Wh, lay.Yh);
    orth(lay.Yh, lay.Qh, true, passes);
  }
  void iterate(bool last) {
    gemm_all(lay.Qh, lay.W, lay.Zh);
    orth(lay.Zh, lay.Ph, false, 1);
    gemm_all(lay.Ph, lay.Wh, lay.Yh);
    orth(lay.Yh, lay.Qh, true, last ? 2 : 1);
  }

  // B = Qh W, its Jacobi SVD, Ritz vectors, residual norms; results into the pinned-free out buffer (read())
  void check_enqueue() {
    const int nb = lay.nb;
    gemm_all(lay.Qh, lay.W, lay.B);
    if (err) return;
    double* rn2 = at<double>(lay.o_rn2);
    double* fro2 = at<double>(lay.o_fro2);
    double* offd = at<double>(lay.o_offd);
    int64_t* rnoff = at<int64_t>(lay.o_rnoff);
    gtn_svd_problem* probs = at<gtn_svd_problem>(lay.o_probs);
    int32_t* sw = at<int32_t>(lay.o_sw);
    const std::vector<Mat>* Bsrc = &lay.B;
    if (lay.prerotate && persistent) {
      // Gram pre-rotation for wide subspaces (l > 80, where gtn_gram_rotate's shared-memory Jacobi does not fit): the
      // Jacobi kernel diagonalises G = B B^H first -- l^2 numbers per round instead of l q -- and its accumulated
      // rotations T make the rows of B' = T B orthogonal to the accuracy a Gram matrix allows; the Jacobi SVD of the
      // l x q matrix then needs 2 sweeps instead of 8.  Qh is rotated along (Qh' = T Qh), so that B' = Qh' W and
      // everything downstream is unchanged; the singular values never go through G.
      ctranspose_all(lay.B, lay.Cq);
      gram(lay.B, lay.Cq);
      for (int b = 0; b < nb && !err; ++b) {
        note(gtn_sum_slices(elem(lay.T1[b]), elem(lay.T2[b]), (int64_t)lay.L[b] * lay.L[b], lay.NS, lay.dtype, st));
        ++launches;
      }
      gtn_svd_problem* pg = at<gtn_svd_problem>(lay.o_probs_g);
      note(gtn_jacobi_init(ws, ws, lay.dtype, pg, nb, lay.maxL, rn2, fro2, rnoff, st));
      const int rcg = gtn_jacobi_persistent(ws, ws, lay.dtype, pg, nb, lay.maxL, kJacobiTol, offd, rn2, fro2, rnoff,
                                            kJacobiMaxSweeps, sw, st, 1e-6);
      launches += 2;
      if (rcg == 0) {
        gemm_all(lay.UbH, lay.B, lay.Sq);
        gemm_all(lay.UbH, lay.Qh, lay.Sp);
        for (int b = 0; b < nb && !err; ++b)
          note((int)cudaMemcpyAsync(elem(lay.Qh[b]), elem(lay.Sp[b]), (int64_t)lay.L[b] * lay.P[b] * lay.esz,
                                    cudaMemcpyDeviceToDevice, st));
        probs = at<gtn_svd_problem>(lay.o_probs_r);
        Bsrc = &lay.Sq;
      } else if (rcg != GTN_ERR_UNSUPPORTED) {
        note(rcg);
      }
    }
    note(gtn_jacobi_init(ws, ws, lay.dtype, probs, nb, lay.maxL, rn2, fro2, rnoff, st));
    ++launches;
    host_sweeps = -1;
    bool done = false;
    if (persistent && lay.maxL >= 2 && lay.maxL <= kPersistentMaxRows) {
      const int rc = gtn_jacobi_persistent(ws, ws, lay.dtype, probs, nb, lay.maxL, kJacobiTol, offd, rn2, fro2, rnoff,
                                           kJacobiMaxSweeps, sw, st, kEarlyStop);
      if (rc == 0) { done = true; ++launches; }
      else if (rc == GTN_ERR_UNSUPPORTED) persistent = false;
      else note(rc);
    }
    if (!done && !err && lay.maxL >= 2) {
      int sweeps = 0;
      std::vector<double> h(nb);
      for (;;) {
        note(gtn_jacobi_sweep(ws, ws, lay.dtype, probs, nb, lay.maxL, (int)lay.maxQ, kJacobiTol, offd, rn2, fro2, rnoff, st));
        ++sweeps;
        launches += ((lay.maxL + 1) & ~1) - 1;
        note((int)cudaMemcpyAsync(h.data(), offd, 8 * nb, cudaMemcpyDeviceToHost, st));
        cudaStreamSynchronize(st);
        double mx = 0;
        for (double v : h) mx = std::max(mx, v);
        if (err || mx <= kJacobiTol * kJacobiTol) break;
        if (sweeps >= kJacobiMaxSweeps) { note(GTN_ERR_NOT_CONVERGED); break; }
      }
      host_sweeps = sweeps;
    }
    if (err) return;
    double* s_dev = at<double>(lay.o_sdev);
    char* vh_ptr = ws + (lay.Vk[0].off - (*Bsrc)[0].off) * (int64_t)lay.esz;    // Vh_out is addressed with W's offsets
    note(gtn_jacobi_finish(ws, ws, ws, vh_ptr, s_dev, lay.dtype, probs, at<gtn_svd_out>(lay.o_outs),
                           at<int32_t>(lay.o_order), at<double>(lay.o_nscr), nb, lay.maxL, (int)lay.maxQ, st));
    launches += 2;
    ctranspose_all(lay.Ub, lay.UbH);
    gemm_all(lay.UbH, lay.Qh, lay.Uh);                    // Uh = Ub^H Qh (l x p)
    gemm_all(lay.Vk, lay.Wh, lay.Xh);                     // certificate rows  v_i^H W^H - s_i u_i^H
    for (int b = 0; b < nb && !err; ++b) {
      if (lay.dtype == GTN_C128)
        set_diag_kernel<c128><<<(lay.L[b] + 127) / 128, 128, 0, st>>>((c128*)elem(lay.D[b]), s_dev + lay.soff[b], lay.L[b]);
      else
        set_diag_kernel<double><<<(lay.L[b] + 127) / 128, 128, 0, st>>>((double*)elem(lay.D[b]), s_dev + lay.soff[b], lay.L[b]);
    }
    note_cuda();
    launches += nb;
    gemm_all(lay.D, lay.Uh, lay.Xh, -1.0, 1.0);
    double* res2 = at<double>(lay.o_res2);
    for (int b = 0; b < nb && !err; ++b)
      note(gtn_row_sumsq(elem(lay.Xh[b]), res2 + lay.soff[b], lay.L[b], lay.P[b], lay.dtype, st));
    if (err) return;
    launches += nb + 1;
    pack_out_kernel<<<(std::max(lay.sumL, nb) + 127) / 128, 128, 0, st>>>(at<double>(lay.o_out), s_dev, res2,
                                                                          at<int32_t>(lay.o_kept), sw, lay.sumL, nb);
    note_cuda();
  }

  // one host read-back per check; false: the Jacobi SVD of a projected matrix did not converge
  bool read(std::vector<double>& out, int* sweeps) {
    out.resize(2 * lay.sumL + lay.nb + 2);
    note((int)cudaMemcpyAsync(out.data(), at<double>(lay.o_out), 8 * out.size(), cudaMemcpyDeviceToHost, st));
    note((int)cudaStreamSynchronize(st));
    meta_used = 0;                                        // every enqueued launch has consumed its tables
    if (err) return false;
    if (host_sweeps >= 0) { *sweeps = host_sweeps; return true; }
    *sweeps = (int)out[2 * lay.sumL + lay.nb];
    return out[2 * lay.sumL + lay.nb + 1] != 0.0;
  }

  double sumsq(const char* x, int64_t n) {
    double* acc = at<double>(lay.o_acc);
    note(gtn_sumsq(x, n, lay.dtype, acc, 1, st));
    ++launches;
    double h = 0;
    note((int)cudaMemcpyAsync(&h, acc, 8, cudaMemcpyDeviceToHost, st));
    note((int)cudaStreamSynchronize(st));
    meta_used = 0;
    return h;
  }

  // Upper bound of || (I - U_D U_D^H) W_b ||_2 (see _engine.deflated_norm_bound).  DESTROYS W_b and Wh_b (the
  // deflated matrix is formed in place); the caller reloads them from the input when the iteration continues.
  double deflated_bound(int b, int nD, double thr) {
    const Mat& W = lay.W[b];
    const int64_t p = lay.P[b], q = lay.Q[b];
    const int l = lay.L[b];
    ctranspose(elem(lay.Uh[b]), p, elem(lay.U[b]), l, l, p);                    // U = Uh^H (p x l)
    Mat UhD{lay.Uh[b].off, nD, p}, UD{lay.U[b].off, p, l}, X{lay.Zh[b].off, nD, q};
    for (int pass = 0; pass < 2 && nD > 0; ++pass) {                            // "twice is enough"
      std::vector<gtn_gemm_group> g1{group(UhD, W, X)};
      gemm(g1);
      gtn_gemm_group g = group(UD, X, W, -1.0, 1.0);
      g.k = nD;                                                                 // first nD columns of U (ld = l)
      std::vector<gtn_gemm_group> g2{g};
      gemm(g2);
    }
    const double f = sqrt(std::max(sumsq(elem(W), p * q), 0.0));
    if (err || f <= thr || std::min(p, q) < 2) return f;
    if (f > thr * sqrt((double)std::min(p, q))) return f;      // ||N||_2 >= ||N||_F / sqrt(rank): no bound can reach thr
    ctranspose(elem(W), q, elem(lay.Wh[b]), p, p, q);                           // N^H (q x p)
    Mat Gm{lay.scratchG, std::min(p, q), std::min(p, q)};
    std::vector<gtn_gemm_group> g3{p <= q ? group(W, lay.Wh[b], Gm) : group(lay.Wh[b], W, Gm)};
    gemm(g3);
    const double t = sqrt(std::max(sumsq(ws + Gm.off * (int64_t)lay.esz, Gm.r * Gm.c), 0.0));
    return std::min(f, sqrt(t));
  }
};

struct Verdict { bool ok, reject; double worst; };

// port of _engine._trunc_certificate; `out` = [s (sumL) | res2 (sumL) | kept (nb) | ...]; s is lowered in place when
// the rank certificate decides the count of non-zero values
Verdict certificate(Driver& d, std::vector<double>& out, const void* const* M, double numer_cutoff) {
  const Layout& lay = d.lay;
  Verdict v{true, false, 0.0};
  for (int b = 0; b < lay.nb; ++b) {
    double* s = out.data() + lay.soff[b];
    const double* r2 = out.data() + lay.sumL + lay.soff[b];
    const int l = lay.L[b];
    const int kept = (int)out[2 * lay.sumL + b];
    const double s0 = l ? s[0] : 0.0;
    int nnz = 0;
    for (int i = 0; i < l; ++i) nnz += fabs(s[i] / (fabs(s0) + numer_cutoff)) > numer_cutoff;
    int kk = std::min(lay.K[b], nnz);
    const bool band = kk > 0 && s[kk - 1] <= kRankNoise * s0;
    const bool dropped = !d.robust && nnz < lay.K[b] && kept < l && nnz >= kept;
    if (band || dropped) {
      int nD = 0;
      for (int i = 0; i < l; ++i) nD += s[i] > kRankNoise * s0;
      const double thr = numer_cutoff * (fabs(s0) + numer_cutoff);
      if (nD == 0 || nD >= l) { v.ok = false; v.reject = true; return v; }
      const double bound = d.deflated_bound(b, nD, thr);
      d.load_one(M, b);                                   // W_b / Wh_b were consumed by the bound
      if (d.err || !(bound <= thr)) { v.ok = false; v.reject = true; return v; }
      for (int i = nD; i < l; ++i) s[i] = std::min(s[i], bound);
      nnz = nD;
      kk = std::min(lay.K[b], nnz);
    }
    double rmax = 0;
    for (int i = 0; i < kk; ++i) rmax = std::max(rmax, sqrt(std::max(r2[i], 0.0)));
    if (kk > 0) {
      v.worst = std::max(v.worst, rmax / std::max(s0, 1e-300));
      if (rmax > kTruncTol * s0) v.ok = false;
    }
    if (kk > 0 && kk < l) {
      double hi = 0;
      for (int i = kk; i < l; ++i) hi = std::max(hi, s[i] + sqrt(std::max(r2[i], 0.0)));
      const double rk = sqrt(std::max(r2[kk - 1], 0.0));
      if (hi > s[kk - 1] + std::max(kTruncTol * s0, rk)) {
        v.ok = false;
        v.worst = std::max(v.worst, (hi - s[kk - 1]) / std::max(s0, 1e-300));
      }
    }
  }
  return v;
}

}  // namespace

extern "C" int64_t gtn_workspace_bytes(int op, int dtype, int nb, const int64_t* m, const int64_t* n, const int32_t* k) {
  if ((op != GTN_OP_SECTOR_SVD_TRUNC && op != GTN_OP_SECTOR_EIGH_TRUNC) || nb < 1 || (dtype != GTN_C128 && dtype != GTN_F64))
    return GTN_ERR_BAD_ARG;
  for (int b = 0; b < nb; ++b)
    if (m[b] < 1 || n[b] < 1 || k[b] < 1) return GTN_ERR_BAD_ARG;
  Layout lay;
  lay.build(nb, dtype, m, n, k);
  return lay.total_bytes;
}

extern "C" int gtn_sector_svd_trunc(const void* const* M, const int64_t* m, const int64_t* n, int nb, int dtype,
                                    const int32_t* k, double numer_cutoff, void* const* U_out, double* S_host,
                                    void* const* Vh_out, int32_t* rank_host, void* workspace, int64_t workspace_bytes,
                                    gtn_svd_info* info, void* stream) {
  if (nb < 1 || (dtype != GTN_C128 && dtype != GTN_F64) || !M || !U_out || !Vh_out || !S_host || !rank_host || !workspace)
    return GTN_ERR_BAD_ARG;
  if ((uintptr_t)workspace & 255) return GTN_ERR_BAD_ARG;
  for (int b = 0; b < nb; ++b)
    if (m[b] < 1 || n[b] < 1 || k[b] < 1 || k[b] > std::min(m[b], n[b])) return GTN_ERR_BAD_ARG;
  Driver d;
  d.lay.build(nb, dtype, m, n, k);
  if (workspace_bytes < d.lay.total_bytes) return GTN_ERR_BAD_ARG;
  d.ws = (char*)workspace;
  d.st = (cudaStream_t)stream;
  const Layout& lay = d.lay;
  d.robust = info && info->robust != 0;
  d.init_tables();
  d.load(M);
  const int start_it = info ? std::max(0, std::min(info->start_iters, kMaxIters)) : 0;
  std::vector<double> out;
  double prev_worst = -1.0;
  int prev_it = -1, next_check = 0, sweeps = 0, checks = 0;
  double rate_mem = info && info->rate > 0 ? info->rate : 0.0;
  int result = GTN_ERR_NOT_CONVERGED;
  for (int it = 0; it <= kMaxIters && !d.err; ++it) {
    if (it == 0) d.start(start_it == 0 ? 2 : 1);
    else d.iterate(it >= start_it && it >= next_check);
    if (it < start_it || it < next_check) continue;
    d.check_enqueue();
    if (!d.read(out, &sweeps)) break;                       // Jacobi did not converge (or a launch failed)
    ++checks;
    Verdict v = certificate(d, out, M, numer_cutoff);
    if (info) { info->iters = it; info->worst = v.worst; info->sweeps = sweeps; info->checks = checks; info->launches = d.launches; }
    if (d.err || v.reject) break;
    if (prev_worst > 0 && v.worst > 0 && it > prev_it)
      rate_mem = std::min(std::max(pow(v.worst / prev_worst, 1.0 / (it - prev_it)), 1e-3), 0.9);
    if (v.ok) { result = GTN_OK; break; }
    if (prev_worst >= 0 && it >= 2) {
      const double rate = prev_worst > 0 ? pow(v.worst / prev_worst, 1.0 / std::max(it - prev_it, 1)) : 1.0;
      if (rate > 0.6) break;                                // stalled: flat spectrum at the cut
      const double need = log(std::max(kTruncTol * 0.3, 1e-300) / std::max(v.worst, 1e-300)) / log(std::max(rate, 1e-3));
      next_check = it + std::max(1, std::min((int)ceil(need), 6));
      if (next_check > kMaxIters) break;
    } else if (prev_worst < 0 && rate_mem > 0 && v.worst > 0) {
      const double need = log(std::max(kTruncTol * 0.3, 1e-300) / v.worst) / log(rate_mem);
      next_check = std::min(it + std::max(1, std::min((int)ceil(need), 6)), kMaxIters);
    }
    prev_worst = v.worst;
    prev_it = it;
  }
  if (info) { info->rate = rate_mem; info->launches = d.launches; }
  if (d.err) return d.err;
  if (result != GTN_OK) return result;
  // ---- results: U_b = (Uh_b^H)[:, :k], Vh_b = Vk_b[:k], S, rank by the reference's rule on the kept values
  int so = 0;
  for (int b = 0; b < nb; ++b) {
    const int kb = k[b], l = lay.L[b];
    d.ctranspose(d.elem(lay.Uh[b]), lay.P[b], (char*)U_out[b], kb, kb, lay.P[b]);         // first kb rows of Uh -> p x kb
    d.note((int)cudaMemcpyAsync(Vh_out[b], d.elem(lay.Vk[b]), (int64_t)kb * lay.Q[b] * lay.esz, cudaMemcpyDeviceToDevice,
                                d.st));
    const double* s = out.data() + lay.soff[b];
    const double s0 = l ? s[0] : 0.0;
    int nnz = 0;
    for (int i = 0; i < l; ++i) nnz += fabs(s[i] / (fabs(s0) + numer_cutoff)) > numer_cutoff;
    rank_host[b] = std::min(nnz, kb);
    for (int i = 0; i < kb; ++i) S_host[so + i] = s[i];
    so += kb;
  }
  d.note((int)cudaStreamSynchronize(d.st));
  return d.err;
}

// Hermitian sectors: M = U S Vh from the truncated SVD, signed eigenvalues lam_k = sum_i s_i (Vh U)_ik (reference
// SortedEig, __init__.py:4340-4341); lam_host: interleaved (re, im) doubles, sum_b k_b entries.
extern "C" int gtn_sector_eigh_trunc(const void* const* M, const int64_t* m, const int64_t* n, int nb, int dtype,
                                     const int32_t* k, double numer_cutoff, void* const* U_out, double* S_host,
                                     void* const* Vh_out, double* lam_host, int32_t* rank_host, void* workspace,
                                     int64_t workspace_bytes, gtn_svd_info* info, void* stream) {
  for (int b = 0; b < nb; ++b)
    if (m[b] != n[b]) return GTN_ERR_BAD_ARG;
  int rc = gtn_sector_svd_trunc(M, m, n, nb, dtype, k, numer_cutoff, U_out, S_host, Vh_out, rank_host, workspace,
                                workspace_bytes, info, stream);
  if (rc != GTN_OK) return rc;
  // VU_b = Vh_b U_b (k x k) through the same GEMM kernels: U_b / Vh_b are copied back into workspace panels
  Driver d;
  d.lay.build(nb, dtype, m, n, k);
  d.ws = (char*)workspace;
  d.st = (cudaStream_t)stream;
  const Layout& lay = d.lay;
  int so = 0;
  for (int b = 0; b < nb; ++b) {
    const int kb = k[b];
    const int64_t p = lay.P[b];
    Mat Vm{lay.Vk[b].off, kb, p}, Um{lay.Cp[b].off, p, kb}, VU{lay.Z[b].off, kb, kb};
    d.note((int)cudaMemcpyAsync(d.elem(Um), U_out[b], p * kb * (int64_t)lay.esz, cudaMemcpyDeviceToDevice, d.st));
    d.note((int)cudaMemcpyAsync(d.elem(Vm), Vh_out[b], p * kb * (int64_t)lay.esz, cudaMemcpyDeviceToDevice, d.st));
    std::vector<gtn_gemm_group> g{Driver::group(Vm, Um, VU)};
    d.gemm(g);
    std::vector<double> h((size_t)kb * kb * (dtype == GTN_C128 ? 2 : 1));
    d.note((int)cudaMemcpyAsync(h.data(), d.elem(VU), h.size() * 8, cudaMemcpyDeviceToHost, d.st));
    d.note((int)cudaStreamSynchronize(d.st));
    d.meta_used = 0;
    if (d.err) return d.err;
    for (int c = 0; c < kb; ++c) {
      double re = 0, im = 0;
      for (int i = 0; i < kb; ++i) {
        if (dtype == GTN_C128) { re += S_host[so + i] * h[2 * ((size_t)i * kb + c)]; im += S_host[so + i] * h[2 * ((size_t)i * kb + c) + 1]; }
        else re += S_host[so + i] * h[(size_t)i * kb + c];
      }
      lam_host[2 * (so + c)] = re;
      lam_host[2 * (so + c) + 1] = im;
    }
    so += kb;
  }
  return GTN_OK;
}
