// gtn_svd.cu -- batched one-sided Jacobi (Hestenes) SVD of the parity-sector matrices, sm_100a.
//
// Replaces np.linalg.svd (LAPACK gesdd) in SortedSVD / SortedEig (reference __init__.py:3932,
// :4323).  A Gram-matrix route (A^H A) cannot reproduce the reference's rank rule
// s_i/(s_0+1e-14) > 1e-14 (__init__.py:3939-3941): it loses everything below 1e-8*s_0, and the
// Z2 gauge tensors ARE rank deficient (first TRG step keeps 16 of 32 per sector).  One-sided
// Jacobi works on the rows themselves, so converged singular values carry an absolute error of
// O(eps * s_0) like LAPACK and exact-zero directions come out as (numerically) zero rows.
//
// Rows of W (p x q, row-major, p <= q) are rotated pairwise until mutually orthogonal; the same
// unitary 2x2 rotations are accumulated into Z (p x p, starts as identity):
//     W_final = Z * W_0,  rows of W_final orthogonal  =>  W_0 = Z^H diag(s) Vh.
// One CTA per row pair, round-robin (chess-tournament) ordering: P-1 rounds of P/2 disjoint
// pairs per sweep, one launch per round for the WHOLE batch of sector matrices (the E and O
// sectors of both SVDs of a TRG step go in one batch).  Row elements stay in registers between
// the reduction and the rotation when q <= 256*CACHE.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gtn_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int JT = 256;     // threads per CTA
constexpr int CACHE = 8;    // row elements cached per thread

struct c128 {
  double re, im;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}

template <bool CPLX>
struct Elem;
template <>
struct Elem<true> {
  using T = c128;
  static __device__ __forceinline__ T ld(const T* p) {
    double2 t = *reinterpret_cast<const double2*>(p);
    T r; r.re = t.x; r.im = t.y; return r;
  }
  static __device__ __forceinline__ void st(T* p, T v) { *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im); }
};
template <>
struct Elem<false> {
  using T = double;
  static __device__ __forceinline__ T ld(const T* p) { return *p; }
  static __device__ __forceinline__ void st(T* p, T v) { *p = v; }
};

// rotation:  x' = cs*x - sn*(ph*y) ;  y' = sn*x + cs*(ph*y),  ph = e^{i theta}, c = <x,y^*> = |c| e^{i theta}
__device__ __forceinline__ void rot(c128& x, c128& y, double cs, double sn, double phr, double phi) {
  const double yr = phr * y.re - phi * y.im, yi = phr * y.im + phi * y.re;
  const double xr = x.re, xi = x.im;
  x.re = cs * xr - sn * yr; x.im = cs * xi - sn * yi;
  y.re = sn * xr + cs * yr; y.im = sn * xi + cs * yi;
}
__device__ __forceinline__ void rot(double& x, double& y, double cs, double sn, double phr, double) {
  const double yy = phr * y, xx = x;
  x = cs * xx - sn * yy;
  y = sn * xx + cs * yy;
}

// Jacobi rotation for the row pair with squared norms a, b and inner product c = <x, y^*> = cr + i ci:
// returns false when |c| <= tol |x||y| (no rotation).  One rsqrt for |c| (phase, zeta), one sqrt + one
// division for t = tan(theta), one rsqrt for cos(theta): the rotation is computed by a single thread (or
// redundantly by a lane group) on the critical path of every round, so the special-function chain is
// kept short.  *off2 = |c|^2 / (a b).
__device__ __forceinline__ bool jacobi_rotation(double a, double b, double cr, double ci, double tol, double& cs,
                                                double& sn, double& phr, double& phi, double& tcabs, double* off2) {
  const double c2 = cr * cr + ci * ci;
  const double ab = a * b;
  if (!(c2 > tol * tol * ab) || !(ab > 0.0)) return false;
  const double ic = rsqrt(c2);
  const double zeta = (b - a) * 0.5 * ic;
  const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  cs = rsqrt(1.0 + tt * tt);
  sn = cs * tt;
  phr = cr * ic; phi = ci * ic;
  tcabs = tt * c2 * ic;
  if (off2) *off2 = c2 / ab;
  return true;
}

// one (round, pair) step of the one-sided Jacobi iteration for problem `prob`, executed by one CTA
template <bool CPLX>
__device__ __forceinline__ void jacobi_pair_step(typename Elem<CPLX>::T* __restrict__ Wb,
                                                 typename Elem<CPLX>::T* __restrict__ Zb,
                                                 const gtn_svd_problem* __restrict__ probs, int prob, int k,
                                                 int round, double tol, double* __restrict__ offdiag,
                                                 double* __restrict__ rn2, const double* __restrict__ fro2,
                                                 const int64_t* __restrict__ rn_off, int nprob) {
  using T = typename Elem<CPLX>::T;
  const gtn_svd_problem pr = probs[prob];
  const int p = pr.p, q = pr.q;
  const int P = (p + 1) & ~1;
  // The running maximum of the squared row norms (dead-row threshold below) is double buffered by round parity:
  // a round READS the value completed by the previous round and accumulates its own maximum into the other buffer,
  // so what a CTA sees never depends on how far its neighbours of the same round have got -- rotation decisions,
  // and with them every bit of the result, are reproducible from run to run and from GPU to GPU.  (Round 0 of a
  // sweep reads the buffer that is one sweep old: a lower bound of the maximum, which only delays declaring a row
  // dead.)  The carry is done before any early exit.
  const double* fro_rd = fro2 + (round & 1) * nprob;
  double* fro_wr = const_cast<double*>(fro2) + ((round + 1) & 1) * nprob;
  if (k == 0 && threadIdx.x == 0) atomic_max_pos(fro_wr + prob, fro_rd[prob]);
  if (P < 2 || round >= P - 1) return;
    if (k >= P / 2) return;
  int i, j;
  if (k == 0) { i = P - 1; j = round; }
  else { i = (round + k) % (P - 1); j = (round - k + (P - 1)) % (P - 1); }
  if (i >= p || j >= p) return;
  if (i > j) { int tmp = i; i = j; j = tmp; }

  // dead-row test on the stored squared row norms.  A row whose norm has fallen below
  // 2e-15 * (largest row norm seen so far <= s_0) can only carry singular values that the rank rule
  // s_i/s_0 > 1e-14 (reference __init__.py:3939-3941) discards anyway, and its overlap with a live
  // row perturbs that row by O(1e-15) relative -- so such pairs are never read again.  This is what
  // makes rank-deficient sectors (the normal case for the gauge tensors) cheap.
  double* rn = rn2 + rn_off[prob];
  {
    const double as = rn[i], bs = rn[j];
    const double dead = 4e-30 * fro_rd[prob];    // max_i |row_i|^2 as of the end of the previous round (monotone)
    if (fmin(as, bs) <= dead) return;
  }

  T* x = Wb + pr.w_off + int64_t(i) * q;
  T* y = Wb + pr.w_off + int64_t(j) * q;
  const int tid = threadIdx.x;

  T xc[CACHE], yc[CACHE];
  double a = 0, b = 0, cr = 0, ci = 0;
  const bool cached = q <= JT * CACHE;
  if (cached) {
#pragma unroll
    for (int u = 0; u < CACHE; ++u) {
      const int e = tid + u * JT;
      if (e < q) {
        xc[u] = Elem<CPLX>::ld(x + e);
        yc[u] = Elem<CPLX>::ld(y + e);
        if constexpr (CPLX) {
          a += xc[u].re * xc[u].re + xc[u].im * xc[u].im;
          b += yc[u].re * yc[u].re + yc[u].im * yc[u].im;
          cr += xc[u].re * yc[u].re + xc[u].im * yc[u].im;   // x * conj(y)
          ci += xc[u].im * yc[u].re - xc[u].re * yc[u].im;
        } else {
          a += xc[u] * xc[u]; b += yc[u] * yc[u]; cr += xc[u] * yc[u];
        }
      }
    }
  } else {
    for (int e = tid; e < q; e += JT) {
      const T xv = Elem<CPLX>::ld(x + e), yv = Elem<CPLX>::ld(y + e);
      if constexpr (CPLX) {
        a += xv.re * xv.re + xv.im * xv.im;
        b += yv.re * yv.re + yv.im * yv.im;
        cr += xv.re * yv.re + xv.im * yv.im;
        ci += xv.im * yv.re - xv.re * yv.im;
      } else {
        a += xv * xv; b += yv * yv; cr += xv * yv;
      }
    }
  }
  __shared__ double red[4][JT / 32];
  __shared__ double rotp[6];
  a = warp_sum(a); b = warp_sum(b); cr = warp_sum(cr);
  if (CPLX) ci = warp_sum(ci);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = a; red[1][tid >> 5] = b; red[2][tid >> 5] = cr; red[3][tid >> 5] = ci;
  }
  __syncthreads();
  if (tid == 0) {
    double A = 0, B = 0, CR = 0, CI = 0;
#pragma unroll
    for (int w = 0; w < JT / 32; ++w) { A += red[0][w]; B += red[1][w]; CR += red[2][w]; CI += red[3][w]; }
    double cs = 1.0, sn = 0.0, phr = 1.0, phi = 0.0, act = 0.0, tc = 0.0, off2 = 0.0, exact = 0.0;
    rn[i] = A; rn[j] = B;
    if (jacobi_rotation(A, B, CR, CI, tol, cs, sn, phr, phi, tc, &off2)) {
      atomic_max_pos(offdiag + prob, off2);
      act = 1.0;
      // the squared norms after the rotation by their update formula -- exact to eps * (A + B) only: when a row
      // collapses by more than ~1e-3 in one rotation (nearly parallel rows: every rank-deficient or steeply decaying
      // sector), its new norm is recomputed from the rotated elements below, otherwise a row that still carries
      // directions at 1e-8 ... 1e-15 s_0 could be stored as 0, pass the dead-row test and never be rotated again
      const double ni = fmax(A - tc, 0.0), nj = fmax(B + tc, 0.0);
      rn[i] = ni; rn[j] = nj;
      if (ni < 1e-6 * (A + B) || nj < 1e-6 * (A + B)) exact = 1.0;
      atomic_max_pos(fro_wr + prob, fmax(ni, nj));
    }
    rotp[0] = cs; rotp[1] = sn; rotp[2] = phr; rotp[3] = phi; rotp[4] = act; rotp[5] = exact;
  }
  __syncthreads();
  if (rotp[4] == 0.0) return;
  const double cs = rotp[0], sn = rotp[1], phr = rotp[2], phi = rotp[3];
  const bool exact = rotp[5] != 0.0;
  double na = 0.0, nb2 = 0.0;
  if (cached) {
#pragma unroll
    for (int u = 0; u < CACHE; ++u) {
      const int e = tid + u * JT;
      if (e < q) {
        rot(xc[u], yc[u], cs, sn, phr, phi);
        Elem<CPLX>::st(x + e, xc[u]);
        Elem<CPLX>::st(y + e, yc[u]);
        if (exact) {
          if constexpr (CPLX) {
            na += xc[u].re * xc[u].re + xc[u].im * xc[u].im; nb2 += yc[u].re * yc[u].re + yc[u].im * yc[u].im;
          } else { na += xc[u] * xc[u]; nb2 += yc[u] * yc[u]; }
        }
      }
    }
  } else {
    for (int e = tid; e < q; e += JT) {
      T xv = Elem<CPLX>::ld(x + e), yv = Elem<CPLX>::ld(y + e);
      rot(xv, yv, cs, sn, phr, phi);
      Elem<CPLX>::st(x + e, xv);
      Elem<CPLX>::st(y + e, yv);
      if (exact) {
        if constexpr (CPLX) { na += xv.re * xv.re + xv.im * xv.im; nb2 += yv.re * yv.re + yv.im * yv.im; }
        else { na += xv * xv; nb2 += yv * yv; }
      }
    }
  }
  if (exact) {                          // (block-uniform: rotp[5] is shared)
    na = warp_sum(na); nb2 = warp_sum(nb2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = na; red[1][tid >> 5] = nb2; }
    __syncthreads();
    if (tid == 0) {
      double A2 = 0, B2 = 0;
#pragma unroll
      for (int w = 0; w < JT / 32; ++w) { A2 += red[0][w]; B2 += red[1][w]; }
      rn[i] = A2; rn[j] = B2;
    }
  }
  T* zx = Zb + pr.z_off + int64_t(i) * p;
  T* zy = Zb + pr.z_off + int64_t(j) * p;
  for (int e = tid; e < p; e += JT) {
    T xv = Elem<CPLX>::ld(zx + e), yv = Elem<CPLX>::ld(zy + e);
    rot(xv, yv, cs, sn, phr, phi);
    Elem<CPLX>::st(zx + e, xv);
    Elem<CPLX>::st(zy + e, yv);
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(JT)
    jacobi_round_kernel(typename Elem<CPLX>::T* __restrict__ Wb, typename Elem<CPLX>::T* __restrict__ Zb,
                        const gtn_svd_problem* __restrict__ probs, int round, double tol,
                        double* __restrict__ offdiag, double* __restrict__ rn2,
                        const double* __restrict__ fro2, const int64_t* __restrict__ rn_off) {
  jacobi_pair_step<CPLX>(Wb, Zb, probs, blockIdx.y, blockIdx.x, round, tol, offdiag, rn2, fro2, rn_off, gridDim.y);
}

// grid-wide barrier on a monotonically increasing counter (all CTAs are co-resident: cooperative
// launch).  About 3x cheaper than cooperative_groups::grid_group::sync() for the ~100-CTA grids used
// here; the counter lives in global memory and is zeroed by the host wrapper.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (++epoch) * nblocks;
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// Persistent variant for small problems (all CTAs co-resident, cooperative launch): the whole
// sweep loop runs inside ONE kernel with grid-wide barriers between rounds, and convergence is
// decided on the device -- no per-round launches and no per-sweep host round trip.  Used for the
// l x q projected matrices of the truncated path (l <= 80 rows).
template <bool CPLX>
__global__ void __launch_bounds__(JT)
    jacobi_persistent_kernel(typename Elem<CPLX>::T* __restrict__ Wb, typename Elem<CPLX>::T* __restrict__ Zb,
                             const gtn_svd_problem* __restrict__ probs, int nprob, int max_p, double tol,
                             double* __restrict__ offdiag2, double* __restrict__ rn2,
                             const double* __restrict__ fro2, const int64_t* __restrict__ rn_off,
                             int max_sweeps, int32_t* __restrict__ sweeps_out, unsigned int* __restrict__ bar,
                             double early2) {
  const int P = (max_p + 1) & ~1;
  const int prob = blockIdx.y, k = blockIdx.x;
  const unsigned int nblocks = gridDim.x * gridDim.y;
  unsigned int epoch = 0;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    double* off = offdiag2 + (sweep & 1) * nprob;
    double* off_next = offdiag2 + ((sweep + 1) & 1) * nprob;
    for (int round = 0; round < P - 1; ++round) {
      jacobi_pair_step<CPLX>(Wb, Zb, probs, prob, k, round, tol, off, rn2, fro2, rn_off, nprob);
      grid_barrier(bar, nblocks, epoch);
      if (round == 0 && k == 0 && threadIdx.x == 0) off_next[prob] = 0.0;
    }
    // converged when no pair of any problem rotated in this sweep -- or, with early2 > 0, when every pair that was
    // rotated had a normalised inner product^2 <= early2: the cyclic Jacobi iteration converges quadratically, so
    // what this sweep leaves behind is already below the rotation threshold and the confirming sweep is skipped
    double m = 0.0;
    for (int b = 0; b < nprob; ++b) m = fmax(m, *reinterpret_cast<volatile double*>(off + b));
    if (m <= early2) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    sweeps_out[0] = sweep;
    double m = 0.0;
    const double* off = offdiag2 + ((sweep - 1) & 1) * nprob;
    for (int b = 0; b < nprob; ++b) m = fmax(m, off[b]);
    sweeps_out[1] = (m <= early2) ? 1 : 0;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(128)
    jacobi_init_kernel(const typename Elem<CPLX>::T* __restrict__ Wb, typename Elem<CPLX>::T* __restrict__ Zb,
                       const gtn_svd_problem* __restrict__ probs, double* __restrict__ rn2,
                       double* __restrict__ fro2, const int64_t* __restrict__ rn_off) {
  using T = typename Elem<CPLX>::T;
  const gtn_svd_problem pr = probs[blockIdx.y];
  const int r = blockIdx.x;
  if (r >= pr.p) return;
  T* z = Zb + pr.z_off + int64_t(r) * pr.p;
  for (int e = threadIdx.x; e < pr.p; e += blockDim.x) {
    if constexpr (CPLX) { T v; v.re = (e == r) ? 1.0 : 0.0; v.im = 0.0; Elem<CPLX>::st(z + e, v); }
    else Elem<CPLX>::st(z + e, (e == r) ? 1.0 : 0.0);
  }
  const T* x = Wb + pr.w_off + int64_t(r) * pr.q;
  double a = 0;
  for (int e = threadIdx.x; e < pr.q; e += blockDim.x) {
    const T v = Elem<CPLX>::ld(x + e);
    if constexpr (CPLX) a += v.re * v.re + v.im * v.im; else a += v * v;
  }
  __shared__ double red[4];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double A = red[0] + red[1] + red[2] + red[3];
    rn2[rn_off[blockIdx.y] + r] = A;
    atomic_max_pos(fro2 + blockIdx.y, A);
  }
}

// s_out[s_off + i] (unsorted position i) = ||row_i||
template <bool CPLX>
__global__ void __launch_bounds__(JT)
    row_norm_kernel(const typename Elem<CPLX>::T* __restrict__ Wb, const gtn_svd_problem* __restrict__ probs,
                    const gtn_svd_out* __restrict__ outs, double* __restrict__ tmp_norm) {
  using T = typename Elem<CPLX>::T;
  const gtn_svd_problem pr = probs[blockIdx.y];
  const int r = blockIdx.x;
  if (r >= pr.p) return;
  const T* x = Wb + pr.w_off + int64_t(r) * pr.q;
  double a = 0;
  for (int e = threadIdx.x; e < pr.q; e += JT) {
    const T v = Elem<CPLX>::ld(x + e);
    if constexpr (CPLX) a += v.re * v.re + v.im * v.im; else a += v * v;
  }
  __shared__ double red[JT / 32];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double A = 0;
    for (int w = 0; w < JT / 32; ++w) A += red[w];
    tmp_norm[outs[blockIdx.y].s_off + r] = sqrt(A);
  }
}

// rank-sort rows by descending norm (ties: lower index first) and scatter s, Vh, U = Z^H
template <bool CPLX>
__global__ void __launch_bounds__(JT)
    sort_scatter_kernel(const typename Elem<CPLX>::T* __restrict__ Wb, const typename Elem<CPLX>::T* __restrict__ Zb,
                        typename Elem<CPLX>::T* __restrict__ U, typename Elem<CPLX>::T* __restrict__ Vh,
                        double* __restrict__ s_out, const double* __restrict__ tmp_norm,
                        const gtn_svd_problem* __restrict__ probs, const gtn_svd_out* __restrict__ outs,
                        int32_t* __restrict__ order) {
  using T = typename Elem<CPLX>::T;
  const gtn_svd_problem pr = probs[blockIdx.y];
  const gtn_svd_out ou = outs[blockIdx.y];
  const int r = blockIdx.x;
  if (r >= pr.p) return;
  const double* nn = tmp_norm + ou.s_off;
  const double mine = nn[r];
  int cnt = 0;
  for (int e = threadIdx.x; e < pr.p; e += JT) {
    const double o = nn[e];
    cnt += (o > mine || (o == mine && e < r)) ? 1 : 0;
  }
  __shared__ int redi[JT / 32];
  __shared__ int rank_s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) redi[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < JT / 32; ++w) c += redi[w];
    rank_s = c;
    s_out[ou.s_off + c] = mine;
    order[ou.s_off + c] = r;
  }
  __syncthreads();
  const int rank = rank_s;
  const double inv = mine > 0.0 ? 1.0 / mine : 0.0;
  const T* x = Wb + pr.w_off + int64_t(r) * pr.q;
  T* v = Vh + pr.w_off + int64_t(rank) * pr.q;
  for (int e = threadIdx.x; e < pr.q; e += JT) {
    T t = Elem<CPLX>::ld(x + e);
    if constexpr (CPLX) { t.re *= inv; t.im *= inv; } else t *= inv;
    Elem<CPLX>::st(v + e, t);
  }
  // U[e, rank] = conj(Z[r, e])
  const T* z = Zb + pr.z_off + int64_t(r) * pr.p;
  T* u = U + ou.u_off;
  for (int e = threadIdx.x; e < pr.p; e += JT) {
    T t = Elem<CPLX>::ld(z + e);
    if constexpr (CPLX) t.im = -t.im;
    Elem<CPLX>::st(u + int64_t(e) * pr.p + rank, t);
  }
}

}  // namespace

extern "C" int gtn_jacobi_init(const void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev,
                               int nprob, int max_p, double* rownorm2_dev, double* fro2_dev,
                               const int64_t* rn_off_dev, void* stream) {
  if (nprob <= 0 || max_p <= 0) return GTN_OK;
  dim3 grid(max_p, nprob), block(128);
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(fro2_dev, 0, sizeof(double) * 2 * nprob, s);      // two buffers (round parity), see jacobi_pair_step
  if (dtype == GTN_C128)
    jacobi_init_kernel<true><<<grid, block, 0, s>>>((const c128*)W, (c128*)Z, probs_dev, rownorm2_dev, fro2_dev, rn_off_dev);
  else if (dtype == GTN_F64)
    jacobi_init_kernel<false><<<grid, block, 0, s>>>((const double*)W, (double*)Z, probs_dev, rownorm2_dev, fro2_dev, rn_off_dev);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_jacobi_sweep(void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev,
                                int nprob, int max_p, int max_q, double tol, double* offdiag_dev,
                                double* rownorm2_dev, const double* fro2_dev, const int64_t* rn_off_dev,
                                void* stream) {
  (void)max_q;
  if (nprob <= 0 || max_p < 2) return GTN_OK;
  const int P = (max_p + 1) & ~1;
  dim3 grid(P / 2, nprob), block(JT);
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(offdiag_dev, 0, sizeof(double) * nprob, s);
  for (int r = 0; r < P - 1; ++r) {
    if (dtype == GTN_C128)
      jacobi_round_kernel<true><<<grid, block, 0, s>>>((c128*)W, (c128*)Z, probs_dev, r, tol, offdiag_dev,
                                                       rownorm2_dev, fro2_dev, rn_off_dev);
    else if (dtype == GTN_F64)
      jacobi_round_kernel<false><<<grid, block, 0, s>>>((double*)W, (double*)Z, probs_dev, r, tol, offdiag_dev,
                                                        rownorm2_dev, fro2_dev, rn_off_dev);
    else
      return GTN_ERR_BAD_ARG;
  }
  return (int)cudaGetLastError();
}

extern "C" int gtn_jacobi_persistent(void* W, void* Z, int dtype, const gtn_svd_problem* probs_dev,
                                     int nprob, int max_p, double tol, double* offdiag2_dev,
                                     double* rownorm2_dev, const double* fro2_dev,
                                     const int64_t* rn_off_dev, int max_sweeps, int32_t* sweeps_dev,
                                     void* stream, double early_stop) {
  if (nprob <= 0 || max_p < 2) return GTN_ERR_UNSUPPORTED;
  const int P = (max_p + 1) & ~1;
  dim3 grid(P / 2, nprob), block(JT);
  cudaStream_t s = (cudaStream_t)stream;
  const void* fn = dtype == GTN_C128 ? (const void*)jacobi_persistent_kernel<true>
                                     : (const void*)jacobi_persistent_kernel<false>;
  if (dtype != GTN_C128 && dtype != GTN_F64) return GTN_ERR_BAD_ARG;
  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dtype == GTN_C128) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_persistent_kernel<true>, JT, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_persistent_kernel<false>, JT, 0);
  if (!coop || (long long)grid.x * grid.y > (long long)per_sm * sms) return GTN_ERR_UNSUPPORTED;
  cudaMemsetAsync(offdiag2_dev, 0, sizeof(double) * 2 * nprob, s);
  // the barrier counter lives behind the two sweep counters in sweeps_dev (int32[4])
  unsigned int* bar = reinterpret_cast<unsigned int*>(sweeps_dev) + 2;
  cudaMemsetAsync(bar, 0, sizeof(unsigned int), s);
  double early2 = early_stop > 0.0 ? early_stop * early_stop : 0.0;
  void* args[] = {&W, &Z, (void*)&probs_dev, &nprob, &max_p, &tol, &offdiag2_dev, &rownorm2_dev,
                  (void*)&fro2_dev, (void*)&rn_off_dev, &max_sweeps, &sweeps_dev, &bar, &early2};
  cudaError_t e = cudaLaunchCooperativeKernel(fn, grid, block, args, 0, s);
  return (int)e;
}

extern "C" int gtn_jacobi_finish(const void* W, const void* Z, void* U_out, void* Vh_out,
                                 double* s_out, int dtype, const gtn_svd_problem* probs_dev,
                                 const gtn_svd_out* outs_dev, int32_t* order_dev,
                                 double* norm_scratch_dev, int nprob, int max_p, int max_q,
                                 void* stream) {
  (void)max_q;
  if (nprob <= 0 || max_p <= 0) return GTN_OK;
  dim3 grid(max_p, nprob), block(JT);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) {
    row_norm_kernel<true><<<grid, block, 0, s>>>((const c128*)W, probs_dev, outs_dev, norm_scratch_dev);
    sort_scatter_kernel<true><<<grid, block, 0, s>>>((const c128*)W, (const c128*)Z, (c128*)U_out, (c128*)Vh_out,
                                                     s_out, norm_scratch_dev, probs_dev, outs_dev, order_dev);
  } else if (dtype == GTN_F64) {
    row_norm_kernel<false><<<grid, block, 0, s>>>((const double*)W, probs_dev, outs_dev, norm_scratch_dev);
    sort_scatter_kernel<false><<<grid, block, 0, s>>>((const double*)W, (const double*)Z, (double*)U_out,
                                                      (double*)Vh_out, s_out, norm_scratch_dev, probs_dev,
                                                      outs_dev, order_dev);
  } else {
    return GTN_ERR_BAD_ARG;
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
//  small Hermitian eigen-solver (two-sided Jacobi, one CTA per matrix, everything in shared memory)
//  used to orthonormalise the l x p iterates of the randomized subspace iteration through their
//  l x l Gram matrix:  G = E L E^H  ->  T = L^{-1/2} E^H  (rows with L_i <= rel_thr * L_max are zeroed),
//  so that (T Y)(T Y)^H = I on the retained directions.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int EIG_MAXN = 80;
constexpr int EIG_T = 256;

template <bool CPLX>
__global__ void __launch_bounds__(EIG_T)
    small_eigh_whiten_kernel(const typename Elem<CPLX>::T* __restrict__ Gb, typename Elem<CPLX>::T* __restrict__ Tb,
                             const int64_t* __restrict__ g_off, const int64_t* __restrict__ t_off,
                             const int32_t* __restrict__ ns, double rel_thr, int32_t* __restrict__ kept,
                             double* __restrict__ evals, const int64_t* __restrict__ e_off) {
  using T = typename Elem<CPLX>::T;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int n = ns[blockIdx.x];
  const int ld = n + 1;
  c128* G = reinterpret_cast<c128*>(sm_raw);
  c128* E = G + n * ld;
  double* rot = reinterpret_cast<double*>(E + n * ld);   // per pair: cs, sn, phr, phi
  int* pq = reinterpret_cast<int*>(rot + 4 * (EIG_MAXN / 2));
  __shared__ double offmax_s;
  __shared__ double lmax_s;
  const int tid = threadIdx.x;
  const T* Gin = Gb + g_off[blockIdx.x];
  for (int e = tid; e < n * n; e += EIG_T) {
    const int i = e / n, j = e % n;
    c128 v;
    if constexpr (CPLX) { const T t = Elem<CPLX>::ld(Gin + e); v.re = t.re; v.im = t.im; }
    else { v.re = Elem<CPLX>::ld(Gin + e); v.im = 0.0; }
    G[i * ld + j] = v;
    c128 id; id.re = (i == j) ? 1.0 : 0.0; id.im = 0.0;
    E[i * ld + j] = id;
  }
  __syncthreads();
  const int P = (n + 1) & ~1;
  const int npairs = P / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    if (tid == 0) offmax_s = 0.0;
    __syncthreads();
    for (int round = 0; round < P - 1; ++round) {
      if (tid < npairs) {
        int i, j;
        if (tid == 0) { i = P - 1; j = round; }
        else { i = (round + tid) % (P - 1); j = (round - tid + (P - 1)) % (P - 1); }
        if (i > j) { const int t = i; i = j; j = t; }
        double cs = 1.0, sn = 0.0, phr = 1.0, phi = 0.0;
        int act = 0;
        if (j < n) {
          const double a = G[i * ld + i].re, b = G[j * ld + j].re;
          const c128 c = G[i * ld + j];
          const double cabs = sqrt(c.re * c.re + c.im * c.im);
          const double den = sqrt(fabs(a)) * sqrt(fabs(b));
          if (cabs > 0.0 && den > 0.0 && cabs > 1e-15 * den) {
            const double zeta = (b - a) / (2.0 * cabs);
            const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cs = 1.0 / sqrt(1.0 + tt * tt);
            sn = cs * tt;
            phr = c.re / cabs; phi = c.im / cabs;
            act = 1;
            atomicMax(reinterpret_cast<unsigned long long*>(&offmax_s),
                      (unsigned long long)__double_as_longlong(cabs / den));
          }
        }
        rot[4 * tid + 0] = cs; rot[4 * tid + 1] = sn; rot[4 * tid + 2] = phr; rot[4 * tid + 3] = phi;
        pq[2 * tid] = act ? i : -1; pq[2 * tid + 1] = j;
      }
      __syncthreads();
      // rows: [x; y] <- Jr [x; y] on G and E,  Jr = [[cs, -sn ph], [sn, cs ph]]
      for (int w = tid; w < npairs * n * 2; w += EIG_T) {
        const int k = w / (2 * n), r = w % (2 * n);
        const int which = r / n, col = r % n;
        const int i = pq[2 * k], j = pq[2 * k + 1];
        if (i < 0) continue;
        c128* Mx = which ? E : G;
        c128 x = Mx[i * ld + col], y = Mx[j * ld + col];
        {
          const double cs = rot[4 * k], sn = rot[4 * k + 1], phr = rot[4 * k + 2], phi = rot[4 * k + 3];
          const double yr = phr * y.re - phi * y.im, yi = phr * y.im + phi * y.re;
          c128 xn, yn;
          xn.re = cs * x.re - sn * yr; xn.im = cs * x.im - sn * yi;
          yn.re = sn * x.re + cs * yr; yn.im = sn * x.im + cs * yi;
          Mx[i * ld + col] = xn; Mx[j * ld + col] = yn;
        }
      }
      __syncthreads();
      // columns of G: [g_ri, g_rj] <- [g_ri, g_rj] Jr^H
      for (int w = tid; w < npairs * n; w += EIG_T) {
        const int k = w / n, row = w % n;
        const int i = pq[2 * k], j = pq[2 * k + 1];
        if (i < 0) continue;
        const double cs = rot[4 * k], sn = rot[4 * k + 1], phr = rot[4 * k + 2], phi = rot[4 * k + 3];
        const c128 gi = G[row * ld + i], gj = G[row * ld + j];
        // conj(ph) * gj
        const double jr = phr * gj.re + phi * gj.im, ji = phr * gj.im - phi * gj.re;
        c128 ni, nj;
        ni.re = cs * gi.re - sn * jr; ni.im = cs * gi.im - sn * ji;
        nj.re = sn * gi.re + cs * jr; nj.im = sn * gi.im + cs * ji;
        G[row * ld + i] = ni; G[row * ld + j] = nj;
      }
      __syncthreads();
    }
    if (offmax_s <= 1e-15) break;
    __syncthreads();
  }
  // eigenvalues on the diagonal; T = L^{-1/2} Eacc (rows), small ones zeroed
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < n; ++i) m = fmax(m, G[i * ld + i].re);
    lmax_s = m;
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += (G[i * ld + i].re > rel_thr * m) ? 1 : 0;
    kept[blockIdx.x] = cnt;
  }
  __syncthreads();
  T* Tout = Tb + t_off[blockIdx.x];
  double* ev = evals + e_off[blockIdx.x];
  for (int e = tid; e < n * n; e += EIG_T) {
    const int i = e / n, j = e % n;
    const double lam = G[i * ld + i].re;
    const double sc = (lam > rel_thr * lmax_s && lam > 0.0) ? rsqrt(lam) : 0.0;
    const c128 v = E[i * ld + j];
    if constexpr (CPLX) { T t; t.re = v.re * sc; t.im = v.im * sc; Elem<CPLX>::st(Tout + e, t); }
    else Elem<CPLX>::st(Tout + e, v.re * sc);
    if (j == 0) ev[i] = lam;
  }
}

}  // namespace

extern "C" int gtn_small_eigh_whiten(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                                     const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n,
                                     double rel_thr, int32_t* kept_dev, double* evals_dev,
                                     const int64_t* e_off_dev, void* stream) {
  if (nprob <= 0) return GTN_OK;
  if (max_n > EIG_MAXN) return GTN_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = size_t(2) * max_n * (max_n + 1) * 16 + 4 * (EIG_MAXN / 2) * 8 + 2 * (EIG_MAXN / 2) * 4 + 64;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e1 = cudaFuncSetAttribute(small_eigh_whiten_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(small_eigh_whiten_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess) return (int)e1;
    if (e2 != cudaSuccess) return (int)e2;
    attr = smem;
  }
  if (dtype == GTN_C128)
    small_eigh_whiten_kernel<true><<<nprob, EIG_T, smem, s>>>((const c128*)G, (c128*)T, g_off_dev, t_off_dev, n_dev,
                                                             rel_thr, kept_dev, evals_dev, e_off_dev);
  else if (dtype == GTN_F64)
    small_eigh_whiten_kernel<false><<<nprob, EIG_T, smem, s>>>((const double*)G, (double*)T, g_off_dev, t_off_dev,
                                                              n_dev, rel_thr, kept_dev, evals_dev, e_off_dev);
  else
    return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
//  pivoted-Cholesky whitening of a small Gram matrix (one CTA per matrix, shared memory):
//     P^T G P = L L^H (rank r, stop when the largest remaining diagonal <= rel_thr * first pivot)
//     T = [ L_r^{-1}  0 ] P^T        =>   (T Y)(T Y)^H = I_r   when G = Y Y^H
//  ~5 block barriers per pivot instead of the ~120 per sweep of the Jacobi eigen-solver above.
// ------------------------------------------------------------------------------------------------
namespace {

// phase time stamps (clock64) of block 0 of the last chol_whiten / gram_rotate launch: [0..3] / [4..7]
__device__ long long gtn_phase_clk[8];

constexpr int CHOL_MAXN = 512;       // global-scratch variant
constexpr int CHOL_T_SMALL = 256;


// Diagonally pivoted Cholesky of the Hermitian n x n matrix G (row stride ld), ONE block barrier per
// pivot and no data movement for the pivoting.  Every warp runs the pivot search redundantly on
// lane-owned copies of the remaining diagonal (registers), so pivot index and 1/sqrt(pivot) are known to
// all threads without a broadcast; the rank-1 update of the remaining rows/columns reads the pivot row
// of G, scales it on the fly and stores column k of the factor into the SEPARATE array L:
//     L[k*ld + i] = L_k[i]   (i = original row index, defined for the rows not pivoted before step k)
// piv[k] = original index of the k-th pivot; in pivot order L_r[a][b] = L[b*ld + piv[a]], a >= b.
// The pivot is the largest remaining diagonal entry up to its top 32 bits (sign, exponent, 20 mantissa
// bits): one REDUX instead of a shuffle tree; any near-maximal pivot is as good.
template <int NT, int MAXN>
__device__ __forceinline__ int chol_factor(c128* G, c128* L, int n, int ld, int* piv, double rel_thr) {
  constexpr int NW = (MAXN + 31) / 32;
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned done[NW];
  double dreg[NW];
#pragma unroll
  for (int t = 0; t < NW; ++t) {
    done[t] = 0u;
    const int i = lane + 32 * t;
    dreg[t] = i < n ? G[i * ld + i].re : -1.0;
  }
  double first = 0.0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    // ---- pivot search (identical in every warp)
    double best = -1.0; int bidx = 0;
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const bool free_ = !((done[t] >> lane) & 1u);
      if (free_ && dreg[t] > best) { best = dreg[t]; bidx = lane + 32 * t; }
    }
    const unsigned key = best > 0.0 ? (unsigned)(__double_as_longlong(best) >> 32) : 0u;
    const unsigned top = __reduce_max_sync(0xffffffffu, key);
    const int src = __ffs(__ballot_sync(0xffffffffu, key == top)) - 1;
    best = __shfl_sync(0xffffffffu, best, src);
    bidx = __shfl_sync(0xffffffffu, bidx, src);
    if (k == 0) first = best;
    if (!(best > rel_thr * first && best > 0.0)) { rank = k; break; }
    const int pk = bidx;
    const double inv = rsqrt(best);
    const double inv2 = inv * inv;
    const c128* Gk = G + pk * ld;
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      if ((pk >> 5) == t) done[t] |= 1u << (pk & 31);
      const int i = lane + 32 * t;
      if (i < n && !((done[t] >> lane) & 1u)) {
        const c128 g = Gk[i];
        dreg[t] -= (g.re * g.re + g.im * g.im) * inv2;
      }
    }
    c128* Lk = L + k * ld;
    if (tid == 0) {
      piv[k] = pk;
      c128 d; d.re = best * inv; d.im = 0.0;
      Lk[pk] = d;
    }
    // ---- G[i][j] -= L_k[i] conj(L_k[j]),  L_k[i] = conj(G[pk][i]) / sqrt(pivot), over the rows / columns
    //      that are still free; 16 x (NT/16) thread tile, the first tile column also stores L_k
    for (int i = tid / 16; i < n; i += NT / 16) {
      if ((done[i >> 5] >> (i & 31)) & 1u) continue;
      c128 a = Gk[i];
      a.re *= inv; a.im *= -inv;
      if ((tid & 15) == 0) Lk[i] = a;
      for (int j = tid & 15; j < n; j += 16) {
        if ((done[j >> 5] >> (j & 31)) & 1u) continue;
        c128 b = Gk[j];
        b.re *= inv; b.im *= -inv;
        c128 g = G[i * ld + j];
        g.re -= a.re * b.re + a.im * b.im;
        g.im -= a.im * b.re - a.re * b.im;
        G[i * ld + j] = g;
      }
    }
    __syncthreads();
  }
  // the pivots stop at the numerical rank; the remaining rows keep their place behind them
  if (tid == 0 && rank < n) {
    int k = rank;
    for (int i = 0; i < n; ++i)
      if (!((done[i >> 5] >> (i & 31)) & 1u)) piv[k++] = i;
  }
  __syncthreads();
  return rank;
}

// Register-tile variant for n <= 16*TS (shared-memory kernels, 256 threads as a 16 x 16 grid): thread
// (ty, tx) keeps G[ty + 16 t][tx + 16 u], t, u < TS, in registers for the whole factorisation.  Per pivot
// the owners of the pivot row publish it through a double-buffered shared row, everybody applies the
// rank-1 update to its registers: ONE block barrier and no shared-memory read-modify-write per pivot.
// Same outputs as chol_factor (L, piv, rank); G is read straight from global memory.
template <bool CPLX, int TS>
__device__ __forceinline__ int chol_factor_reg(const typename Elem<CPLX>::T* __restrict__ Gin, int nsplit, c128* L,
                                               int n, int ld, int* piv, double rel_thr, c128* rowbuf) {
  using T = typename Elem<CPLX>::T;
  constexpr int NW = (16 * TS + 31) / 32;
  const int tid = threadIdx.x, lane = tid & 31, tx = tid & 15, ty = tid >> 4;
  c128 g[TS][TS];
#pragma unroll
  for (int t = 0; t < TS; ++t)
#pragma unroll
    for (int u = 0; u < TS; ++u) { g[t][u].re = 0.0; g[t][u].im = 0.0; }
  // G = sum of the split-K partial Gram matrices; all loads of one slice are independent
  for (int sp = 0; sp < nsplit; ++sp) {
    const T* Gs = Gin + sp * n * n;
#pragma unroll
    for (int t = 0; t < TS; ++t)
#pragma unroll
      for (int u = 0; u < TS; ++u) {
        const int i = ty + 16 * t, j = tx + 16 * u;
        if (i < n && j < n) {
          const T x = Elem<CPLX>::ld(Gs + i * n + j);
          if constexpr (CPLX) { g[t][u].re += x.re; g[t][u].im += x.im; } else g[t][u].re += x;
        }
      }
  }
  unsigned long long done_lo = 0ull, done_hi = 0ull;      // pivoted rows (n <= 80 < 128)
  auto is_done = [&](int i) -> bool { return ((i < 64 ? done_lo >> i : done_hi >> (i - 64)) & 1ull) != 0ull; };
  double dreg[NW];
#pragma unroll
  for (int t = 0; t < NW; ++t) {
    const int i = lane + 32 * t;
    double d = -1.0;
    if (i < n) {
      d = 0.0;
      for (int sp = 0; sp < nsplit; ++sp) {
        if constexpr (CPLX) d += Elem<CPLX>::ld(Gin + sp * n * n + i * n + i).re;
        else d += Elem<CPLX>::ld(Gin + sp * n * n + i * n + i);
      }
    }
    dreg[t] = d;
  }
  double first = 0.0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    double best = -1.0; int bidx = 0;
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      if (!is_done(lane + 32 * t) && dreg[t] > best) { best = dreg[t]; bidx = lane + 32 * t; }
    }
    const unsigned key = best > 0.0 ? (unsigned)(__double_as_longlong(best) >> 32) : 0u;
    const unsigned top = __reduce_max_sync(0xffffffffu, key);
    const int src = __ffs(__ballot_sync(0xffffffffu, key == top)) - 1;
    best = __shfl_sync(0xffffffffu, best, src);
    bidx = __shfl_sync(0xffffffffu, bidx, src);
    if (k == 0) first = best;
    if (!(best > rel_thr * first && best > 0.0)) { rank = k; break; }
    const int pk = bidx;
    const double inv = rsqrt(best);
    c128* buf = rowbuf + (k & 1) * (16 * TS);
#pragma unroll
    for (int t = 0; t < TS; ++t) {
      if (ty + 16 * t == pk) {               // static register slots: no dynamic indexing of g
#pragma unroll
        for (int u = 0; u < TS; ++u) buf[tx + 16 * u] = g[t][u];
      }
    }
    __syncthreads();
    c128 a[TS], b[TS];
#pragma unroll
    for (int t = 0; t < TS; ++t) {
      a[t] = buf[ty + 16 * t]; a[t].re *= inv; a[t].im *= -inv;
      b[t] = buf[tx + 16 * t]; b[t].re *= inv; b[t].im *= -inv;
    }
#pragma unroll
    for (int t = 0; t < TS; ++t)
#pragma unroll
      for (int u = 0; u < TS; ++u) {
        g[t][u].re -= a[t].re * b[u].re + a[t].im * b[u].im;
        g[t][u].im -= a[t].im * b[u].re - a[t].re * b[u].im;
      }
    const double inv2 = inv * inv;
    if (pk < 64) done_lo |= 1ull << pk; else done_hi |= 1ull << (pk - 64);
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = lane + 32 * t;
      if (i < n && !is_done(i)) {
        const c128 x = buf[i];
        dreg[t] -= (x.re * x.re + x.im * x.im) * inv2;
      }
    }
    c128* Lk = L + k * ld;
    if (tx == 0) {
#pragma unroll
      for (int t = 0; t < TS; ++t) {
        const int i = ty + 16 * t;
        if (i < n && !is_done(i)) Lk[i] = a[t];
      }
    }
    if (tid == 0) {
      piv[k] = pk;
      c128 d; d.re = best * inv; d.im = 0.0;
      Lk[pk] = d;
    }
  }
  if (tid == 0 && rank < n) {
    int k = rank;
    for (int i = 0; i < n; ++i)
      if (!is_done(i)) piv[k++] = i;
  }
  __syncthreads();
  return rank;
}

template <bool CPLX, int NT>
__device__ __forceinline__ void load_gram(c128* G, const typename Elem<CPLX>::T* Gin, int nsplit, int n, int ld) {
  using T = typename Elem<CPLX>::T;
  for (int e = threadIdx.x; e < n * n; e += NT) {
    const int i = e / n, j = e % n;
    c128 v; v.re = 0.0; v.im = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {
      const T t = Elem<CPLX>::ld(Gin + sp * n * n + e);
      if constexpr (CPLX) { v.re += t.re; v.im += t.im; } else v.re += t;
    }
    G[i * ld + j] = v;
  }
}

// Whitening transform T = [L_r^{-1} 0] P^T from the pivoted Cholesky factor.  The inverse of the
// r x r factor is built by groups of INV_G lanes per column (dot products across the group) into LiT
// (transposed).  GLOB=false keeps G and LiT in shared memory (n <= EIG_MAXN), GLOB=true in a
// caller-provided global scratch (L2-resident, n <= CHOL_MAXN).
template <bool CPLX, int NT, bool GLOB, int TS>
__global__ void __launch_bounds__(NT)
    chol_whiten_kernel(const typename Elem<CPLX>::T* __restrict__ Gb, typename Elem<CPLX>::T* __restrict__ Tb,
                       const int64_t* __restrict__ g_off, const int64_t* __restrict__ t_off,
                       const int32_t* __restrict__ ns, int nsplit, double rel_thr, int32_t* __restrict__ kept,
                       c128* __restrict__ scratch, int64_t scratch_stride) {
  using T = typename Elem<CPLX>::T;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int n = ns[blockIdx.x];
  const int ld = n + 1;
  c128* G;
  if constexpr (GLOB) G = scratch + int64_t(blockIdx.x) * scratch_stride;
  else G = reinterpret_cast<c128*>(sm_raw);
  c128* L = G + n * ld;
  __shared__ int perm[GLOB ? CHOL_MAXN : EIG_MAXN];
  __shared__ double invd[GLOB ? CHOL_MAXN : EIG_MAXN];
  const int tid = threadIdx.x;
  const bool stamp = tid == 0 && blockIdx.x == 0;
  if (stamp) gtn_phase_clk[0] = clock64();
  int r;
  if constexpr (TS > 0) {
    __shared__ c128 rowbuf[2 * 16 * (TS > 0 ? TS : 1)];
    r = chol_factor_reg<CPLX, TS>(Gb + g_off[blockIdx.x], nsplit, L, n, ld, perm, rel_thr, rowbuf);
  } else {
    load_gram<CPLX, NT>(G, Gb + g_off[blockIdx.x], nsplit, n, ld);
    __syncthreads();
    r = chol_factor<NT, (GLOB ? CHOL_MAXN : EIG_MAXN)>(G, L, n, ld, perm, rel_thr);
  }
  if (stamp) gtn_phase_clk[1] = clock64();
  for (int i = tid; i < r; i += NT) invd[i] = 1.0 / L[i * ld + perm[i]].re;
  __syncthreads();
  // ---- LiT[c][i] = (L_r^{-1})[i][c] into the (now dead) storage of G; L_r[i][j] = L[j][perm[i]] (i >= j).
  //      INV_G lanes per column: small n -> all columns in flight at once, large n -> longer dot products
  c128* LiT = G;
  constexpr int INV_G = GLOB ? 32 : 4;
  const int sub = tid % INV_G;
  for (int c0 = 0; c0 < r; c0 += NT / INV_G) {
    const int c = c0 + tid / INV_G;
    const bool live = c < r;
    c128* col = LiT + (live ? c : 0) * ld;
    for (int i = c0; i < r; ++i) {              // warp-uniform trip count; columns c > i just idle
      const int pi = perm[i];
      double sr = 0.0, si = 0.0;
      if (live && i > c) {
        for (int j = c + sub; j < i; j += INV_G) {
          const c128 l = L[j * ld + pi], x = col[j];
          sr -= l.re * x.re - l.im * x.im;
          si -= l.re * x.im + l.im * x.re;
        }
      }
#pragma unroll
      for (int o = INV_G / 2; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (live && i >= c && sub == 0) {
        if (i == c) sr += 1.0;
        const double d = invd[i];
        c128 o; o.re = sr * d; o.im = si * d;
        col[i] = o;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (stamp) gtn_phase_clk[2] = clock64();
  if (tid == 0) kept[blockIdx.x] = r;
  T* Tout = Tb + t_off[blockIdx.x];
  for (int e = tid; e < n * n; e += NT) {
    const int i = e / n, c = e % n;       // T[i][perm[c]] = Li[i][c]
    c128 v; v.re = 0.0; v.im = 0.0;
    if (i < r && c <= i) v = LiT[c * ld + i];
    T* dst = Tout + int64_t(i) * n + perm[c];
    if constexpr (CPLX) { T t; t.re = v.re; t.im = v.im; Elem<CPLX>::st(dst, t); }
    else Elem<CPLX>::st(dst, v.re);
  }
  if (stamp) gtn_phase_clk[3] = clock64();
}

// Shared-memory variant for EIG_MAXN < n <= PK_MAXN (chi = 96 ... 128: l = 128 rows of the iterated subspace), where
// the full G + L pair of the kernel above (2 n (n+1) elements) does not fit any more and the global-scratch variant
// pays an L2 round trip per element and pivot (1.3 ms per launch at n = 128: 13 ms of a chi = 128 TRG step, and the
// largest piece of work that is REPLICATED when the step is sharded over several GPUs).  Here the Hermitian matrix is
// kept as its packed lower triangle (n (n+1) / 2 elements, 129 KB at n = 128) and factorised IN PLACE with physical
// diagonal pivoting (symmetric row / column swaps inside the triangle), so the factor ends up in pivot order and the
// trailing update only touches the trailing triangle.  Per pivot: swap + scale + stage the pivot column into a
// contiguous vector (coalesced, conflict-free reads for everybody), rank-1 update of the trailing lower triangle, two
// block barriers.  The inverse is built row by row in place and scattered into T.
constexpr int PK_MAXN = 128;
constexpr int PK_T = 1024;
static_assert(PK_T / 8 >= PK_MAXN, "the in-place inverse of the packed Cholesky kernel runs 8 lanes per column");
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
// GLOB = false: the triangle lives in shared memory (n <= PK_MAXN).  GLOB = true (PK_MAXN < n <= CHOL_MAXN: chi = 192
// ... 256 subspaces): the same algorithm on a triangle in the caller's global scratch (L2-resident: 0.5 MB at n = 256);
// __syncthreads() orders the block's global accesses.  It replaces the mask-pivoted full-matrix kernel for these
// sizes (n^3 element updates through L2: 2.9 ms per launch at n = 160).
template <bool CPLX, bool GLOB>
__global__ void __launch_bounds__(PK_T, 1)
    chol_whiten_packed_kernel(const typename Elem<CPLX>::T* __restrict__ Gb, typename Elem<CPLX>::T* __restrict__ Tb,
                              const int64_t* __restrict__ g_off, const int64_t* __restrict__ t_off,
                              const int32_t* __restrict__ ns, int nsplit, double rel_thr, int32_t* __restrict__ kept,
                              c128* __restrict__ scratch, int64_t scratch_stride) {
  // Round-2 rewrite (after `ncu --set full` of the first version: 4.0 M warp instructions per matrix, issue slots 67 %
  // busy, 13 % FP64 -- bound by index arithmetic: pivoting by masks made every pivot walk the WHOLE triangle, and the
  // inverse chased the pivot order through two indirections per element).  Now the pivoting is PHYSICAL: pivot k swaps
  // row / column k with the pivot's in the packed triangle (n element swaps), so the trailing update runs over the
  // clean trailing triangle only ((n - k)^2 / 2 elements: n^3 / 6 in total instead of n^3 / 2) and the factor ends up
  // in pivot order in place, where the row-wise in-place inverse reads it without any indirection.
  using T = typename Elem<CPLX>::T;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int n = ns[blockIdx.x];
  constexpr int NMAX = GLOB ? CHOL_MAXN : PK_MAXN;
  c128* Gp;
  if constexpr (GLOB) Gp = scratch + int64_t(blockIdx.x) * scratch_stride;
  else Gp = reinterpret_cast<c128*>(sm_raw);
  __shared__ c128 vb[2][NMAX];
  __shared__ int perm[NMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool stamp = tid == 0 && blockIdx.x == 0;
  if (stamp) gtn_phase_clk[0] = clock64();
  const T* Gin = Gb + g_off[blockIdx.x];
  T* Tout = Tb + t_off[blockIdx.x];
  // ---- load the lower triangle (sum of the split-K partial Gram matrices), clear T
  for (int e = tid; e < n * n; e += PK_T) {
    const int i = e / n, j = e - i * n;
    if (j <= i) {
      c128 v; v.re = 0.0; v.im = 0.0;
      for (int sp = 0; sp < nsplit; ++sp) {
        const T t = Elem<CPLX>::ld(Gin + sp * n * n + e);
        if constexpr (CPLX) { v.re += t.re; v.im += t.im; } else v.re += t;
      }
      Gp[tri(i) + j] = v;
    }
    if constexpr (CPLX) { T z; z.re = 0.0; z.im = 0.0; Elem<CPLX>::st(Tout + e, z); } else Elem<CPLX>::st(Tout + e, 0.0);
  }
  if (tid < n) perm[tid] = tid;
  __syncthreads();
  // ---- diagonally pivoted Cholesky with physical symmetric swaps, in place: L_r[a][b] ends up at tri(a) + b
  double first = 0.0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    // pivot search on the remaining diagonal, redundantly by every warp (same data, same result: no broadcast)
    double best = -1.0; int bidx = k;
    for (int i = k + lane; i < n; i += 32) {
      const double d = Gp[tri(i) + i].re;
      if (d > best) { best = d; bidx = i; }
    }
    const unsigned key = best > 0.0 ? (unsigned)(__double_as_longlong(best) >> 32) : 0u;
    const unsigned top = __reduce_max_sync(0xffffffffu, key);
    const int src = __ffs(__ballot_sync(0xffffffffu, key == top)) - 1;
    best = __shfl_sync(0xffffffffu, best, src);
    bidx = __shfl_sync(0xffffffffu, bidx, src);
    if (k == 0) first = best;
    if (!(best > rel_thr * first && best > 0.0)) { rank = k; break; }
    const int p = bidx;
    const double inv = rsqrt(best);
    c128* v = vb[k & 1];
    // swap k <-> p (thread x owns index x), scale column k by 1 / sqrt(pivot), stage it contiguously in v
    if (tid < n) {
      const int x = tid;
      if (p != k) {
        if (x < k) {
          const c128 a = Gp[tri(k) + x]; Gp[tri(k) + x] = Gp[tri(p) + x]; Gp[tri(p) + x] = a;
        } else if (x > k && x < p) {
          c128 a = Gp[tri(x) + k], b = Gp[tri(p) + x];
          a.im = -a.im; b.im = -b.im;
          Gp[tri(x) + k] = b; Gp[tri(p) + x] = a;
        } else if (x > p) {
          const c128 a = Gp[tri(x) + k]; Gp[tri(x) + k] = Gp[tri(x) + p]; Gp[tri(x) + p] = a;
        } else if (x == p) {
          c128 a = Gp[tri(p) + k]; a.im = -a.im; Gp[tri(p) + k] = a;
        } else {                                                    // x == k: the two diagonal entries
          const c128 a = Gp[tri(k) + k]; Gp[tri(k) + k] = Gp[tri(p) + p]; Gp[tri(p) + p] = a;
          const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
      }
      if (x > k) {
        c128 a = Gp[tri(x) + k];
        a.re *= inv; a.im *= inv;
        Gp[tri(x) + k] = a;
        v[x] = a;
      } else if (x == k) {
        c128 d; d.re = best * inv; d.im = 0.0;                      // sqrt(pivot)
        Gp[tri(k) + k] = d;
      }
    }
    __syncthreads();
    // trailing update G[i][j] -= L[i][k] conj(L[j][k]) for k < j <= i: one warp per row, lanes along the row
    for (int i = k + 1 + warp; i < n; i += PK_T / 32) {
      const c128 a = v[i];
      c128* row = Gp + tri(i);
      for (int j = k + 1 + lane; j <= i; j += 32) {
        const c128 b = v[j];
        c128 g = row[j];
        g.re -= a.re * b.re + a.im * b.im;
        g.im -= a.im * b.re - a.re * b.im;
        row[j] = g;
      }
    }
    __syncthreads();
  }
  const int r = rank;
  if (stamp) gtn_phase_clk[1] = clock64();
  // ---- X = L_r^{-1} in place, row by row:  X[a][b] = -(1 / L[a][a]) sum_{j = b .. a-1} L[a][j] X[j][b],  X[a][a] = 1 / L[a][a].
  // All columns advance together (8 lanes per column); row a of L is read by everybody before it is overwritten.
  {
    constexpr int NCH = NMAX / (PK_T / 8);          // column chunks of 128 (one in shared memory, up to 4 in global)
    const int b0 = tid >> 3, sub = tid & 7;
    for (int a = 0; a < r; ++a) {
      const c128* La = Gp + tri(a);
      double sr[NCH], si[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int b = b0 + ch * (PK_T / 8);
        sr[ch] = 0.0; si[ch] = 0.0;
        if (b < a) {
          for (int jj = b + sub; jj < a; jj += 8) {
            const c128 l = La[jj];
            const c128 x = Gp[tri(jj) + b];
            sr[ch] += l.re * x.re - l.im * x.im;
            si[ch] += l.re * x.im + l.im * x.re;
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          sr[ch] += __shfl_xor_sync(0xffffffffu, sr[ch], o);
          si[ch] += __shfl_xor_sync(0xffffffffu, si[ch], o);
        }
      }
      const double d = 1.0 / La[a].re;
      __syncthreads();
      if (sub == 0) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int b = b0 + ch * (PK_T / 8);
          if (b <= a) {
            c128 o;
            if (b == a) { o.re = d; o.im = 0.0; }
            else { o.re = -sr[ch] * d; o.im = -si[ch] * d; }
            Gp[tri(a) + b] = o;
          }
        }
      }
      __syncthreads();
    }
    // T[a][perm[b]] = X[a][b]
    for (int e = tid; e < r * r; e += PK_T) {
      const int a = e / r, bb = e - a * r;
      if (bb <= a) {
        const c128 x = Gp[tri(a) + bb];
        T* dst = Tout + int64_t(a) * n + perm[bb];
        if constexpr (CPLX) { T t; t.re = x.re; t.im = x.im; Elem<CPLX>::st(dst, t); } else Elem<CPLX>::st(dst, x.re);
      }
    }
  }
  if (tid == 0) kept[blockIdx.x] = r;
  if (stamp) { gtn_phase_clk[2] = clock64(); gtn_phase_clk[3] = clock64(); }
}

// Pre-rotation for the one-sided Jacobi SVD of a short-and-wide matrix B (n rows): from G = B B^H
// compute a unitary T such that the rows of T B are orthogonal up to the accuracy a Gram matrix
// allows.  G = P L L^H P^T (pivoted Cholesky, above); the rows of L are orthogonalised by one-sided
// Jacobi entirely in shared memory (8 lanes per row pair, one block barrier per round), Z accumulates
// the rotations and T = Z P^T.  The global Jacobi kernel then starts from T B and needs 2-3 sweeps
// instead of 7-8; the singular values it delivers never go through G.
constexpr int ROT_T = 256;
template <bool CPLX, int TS>
__global__ void __launch_bounds__(ROT_T)
    gram_rotate_kernel(const typename Elem<CPLX>::T* __restrict__ Gb, typename Elem<CPLX>::T* __restrict__ Tb,
                       const int64_t* __restrict__ g_off, const int64_t* __restrict__ t_off,
                       const int32_t* __restrict__ ns, int nsplit, double rel_thr, double tol, int max_sweeps,
                       int32_t* __restrict__ sweeps_out) {
  using T = typename Elem<CPLX>::T;
  constexpr int NT = ROT_T;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int n = ns[blockIdx.x];
  const int ld = n + 1;
  c128* G = reinterpret_cast<c128*>(sm_raw);
  c128* L = G + n * ld;
  __shared__ int perm[EIG_MAXN];
  __shared__ int rotated_s;
  __shared__ unsigned long long maxoff_s;      // bits of the largest |<x,y>|^2 / (|x|^2 |y|^2) rotated in this sweep
  __shared__ double maxn_s;
  const int tid = threadIdx.x;
  __shared__ c128 rowbuf[2 * 16 * TS];
  const bool stamp = tid == 0 && blockIdx.x == 0;
  if (stamp) gtn_phase_clk[4] = clock64();
  const int r = chol_factor_reg<CPLX, TS>(Gb + g_off[blockIdx.x], nsplit, L, n, ld, perm, rel_thr, rowbuf);
  if (stamp) gtn_phase_clk[5] = clock64();
  if (tid == 0) maxn_s = r > 0 ? L[perm[0]].re * L[perm[0]].re : 0.0;      // first pivot = max |row|^2
  __syncthreads();
  // M = L_r (n x r, rows in pivot order) into the dead storage of G
  c128* M = G;
  for (int e = tid; e < n * r; e += NT) {
    const int a = e / r, b = e % r;
    c128 v; v.re = 0.0; v.im = 0.0;
    if (b <= a || a >= r) v = L[b * ld + perm[a]];
    M[a * ld + b] = v;
  }
  __syncthreads();
  c128* Z = L;
  for (int e = tid; e < n * n; e += NT) {
    c128 v; v.re = ((e / n) == (e % n)) ? 1.0 : 0.0; v.im = 0.0;
    Z[(e / n) * ld + (e % n)] = v;
  }
  __syncthreads();
  const int P = (n + 1) & ~1;
  const int sub = tid & 7;
  const double dead = 4e-30 * maxn_s;
  int sweep = 0;
  for (; sweep < max_sweeps && r > 0 && P >= 2; ++sweep) {
    if (tid == 0) { rotated_s = 0; maxoff_s = 0ull; }
    __syncthreads();
    for (int round = 0; round < P - 1; ++round) {
      for (int k0 = 0; k0 < P / 2; k0 += NT / 8) {
        const int k = k0 + (tid >> 3);
        int i = 0, j = 0;
        bool valid = k < P / 2;
        if (valid) {
          if (k == 0) { i = P - 1; j = round; }
          else { i = (round + k) % (P - 1); j = (round - k + (P - 1)) % (P - 1); }
          valid = i < n && j < n;
          if (i > j) { const int t = i; i = j; j = t; }
        }
        c128* x = M + i * ld;
        c128* y = M + j * ld;
        double a = 0, b = 0, cr = 0, ci = 0;
        if (valid) {
          for (int e = sub; e < r; e += 8) {
            const c128 xv = x[e], yv = y[e];
            a += xv.re * xv.re + xv.im * xv.im;
            b += yv.re * yv.re + yv.im * yv.im;
            cr += xv.re * yv.re + xv.im * yv.im;
            ci += xv.im * yv.re - xv.re * yv.im;
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          cr += __shfl_xor_sync(0xffffffffu, cr, o);
          ci += __shfl_xor_sync(0xffffffffu, ci, o);
        }
        double cs, sn, phr, phi, tc, off2 = 0.0;
        if (valid && fmin(a, b) > dead && jacobi_rotation(a, b, cr, ci, tol, cs, sn, phr, phi, tc, &off2)) {
          for (int e = sub; e < r; e += 8) rot(x[e], y[e], cs, sn, phr, phi);
          c128* zx = Z + i * ld;
          c128* zy = Z + j * ld;
          for (int e = sub; e < n; e += 8) rot(zx[e], zy[e], cs, sn, phr, phi);
          if (sub == 0) {
            rotated_s = 1;
            atomicMax(&maxoff_s, static_cast<unsigned long long>(__double_as_longlong(off2)));
          }
        }
      }
      __syncthreads();
    }
    // (An early stop "every rotated pair had a normalised overlap <= sqrt(tol)" -- the rule that took the global Jacobi
    // kernel from 4 sweeps to 1 -- was measured here and rejected: 605 -> 577 us for this kernel, but the rows it
    // leaves are orthogonal to ~1e-9 instead of 1e-12, the global kernel's first sweep then rotates pairs above ITS
    // early-stop threshold and runs a second sweep: 200 -> 364 us, chi = 32 step 2.55 -> 2.84 ms.  maxoff_s is kept
    // for diagnostics.)
    if (!rotated_s) { ++sweep; break; }
    __syncthreads();
  }
  if (tid == 0 && sweeps_out) sweeps_out[blockIdx.x] = sweep;
  if (stamp) gtn_phase_clk[6] = clock64();
  T* Tout = Tb + t_off[blockIdx.x];
  for (int e = tid; e < n * n; e += NT) {
    const int i = e / n, c = e % n;       // T[i][perm[c]] = Z[i][c]
    const c128 v = Z[i * ld + c];
    T* dst = Tout + int64_t(i) * n + perm[c];
    if constexpr (CPLX) { T t; t.re = v.re; t.im = v.im; Elem<CPLX>::st(dst, t); }
    else Elem<CPLX>::st(dst, v.re);
  }
  if (stamp) gtn_phase_clk[7] = clock64();
}

template <typename K>
int set_smem(K kernel, size_t smem) {
  return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

extern "C" int64_t gtn_chol_whiten_scratch_elems(int max_n) {
  return max_n <= EIG_MAXN ? 0 : int64_t(2) * max_n * (max_n + 1);
}

extern "C" int gtn_chol_whiten(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                               const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n, int nsplit,
                               double rel_thr, int32_t* kept_dev, void* scratch, void* stream) {
  if (nprob <= 0) return GTN_OK;
  if (nsplit < 1) return GTN_ERR_BAD_ARG;
  if (max_n > CHOL_MAXN) return GTN_ERR_UNSUPPORTED;
  if (dtype != GTN_C128 && dtype != GTN_F64) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (max_n <= EIG_MAXN) {
    const size_t smem = size_t(2) * max_n * (max_n + 1) * 16 + 64;
    static size_t attr = 0;
    if (smem > attr) {
      int e = set_smem(chol_whiten_kernel<true, CHOL_T_SMALL, false, 3>, smem);
      if (!e) e = set_smem(chol_whiten_kernel<false, CHOL_T_SMALL, false, 3>, smem);
      if (!e) e = set_smem(chol_whiten_kernel<true, CHOL_T_SMALL, false, 5>, smem);
      if (!e) e = set_smem(chol_whiten_kernel<false, CHOL_T_SMALL, false, 5>, smem);
      if (e) return e;
      attr = smem;
    }
    const bool c = dtype == GTN_C128;
#define GTN_CHOL_LAUNCH(CP, TS_, GT)                                                                       \
  chol_whiten_kernel<CP, CHOL_T_SMALL, false, TS_><<<nprob, CHOL_T_SMALL, smem, s>>>(                      \
      (const GT*)G, (GT*)T, g_off_dev, t_off_dev, n_dev, nsplit, rel_thr, kept_dev, nullptr, 0)
    if (max_n <= 48) { if (c) GTN_CHOL_LAUNCH(true, 3, c128); else GTN_CHOL_LAUNCH(false, 3, double); }
    else             { if (c) GTN_CHOL_LAUNCH(true, 5, c128); else GTN_CHOL_LAUNCH(false, 5, double); }
#undef GTN_CHOL_LAUNCH
  } else if (max_n <= PK_MAXN) {
    // packed lower triangle + 32 column buffers in shared memory (199 KB at n = 128)
    const size_t smem = (size_t(max_n) * (max_n + 1) / 2) * 16;        // the packed lower triangle
    static size_t attr = 0;
    if (smem > attr) {
      int e = set_smem(chol_whiten_packed_kernel<true, false>, smem);
      if (!e) e = set_smem(chol_whiten_packed_kernel<false, false>, smem);
      if (e) return e;
      attr = smem;
    }
    if (dtype == GTN_C128)
      chol_whiten_packed_kernel<true, false><<<nprob, PK_T, smem, s>>>((const c128*)G, (c128*)T, g_off_dev, t_off_dev,
                                                                        n_dev, nsplit, rel_thr, kept_dev, nullptr, 0);
    else
      chol_whiten_packed_kernel<false, false><<<nprob, PK_T, smem, s>>>((const double*)G, (double*)T, g_off_dev,
                                                                         t_off_dev, n_dev, nsplit, rel_thr, kept_dev,
                                                                         nullptr, 0);
  } else {
    // PK_MAXN < n <= CHOL_MAXN: the packed physical-pivot kernel on a triangle in the global scratch
    if (!scratch) return GTN_ERR_BAD_ARG;
    const int64_t stride = gtn_chol_whiten_scratch_elems(max_n);
    if (dtype == GTN_C128)
      chol_whiten_packed_kernel<true, true><<<nprob, PK_T, 0, s>>>((const c128*)G, (c128*)T, g_off_dev, t_off_dev, n_dev,
                                                                    nsplit, rel_thr, kept_dev, (c128*)scratch, stride);
    else
      chol_whiten_packed_kernel<false, true><<<nprob, PK_T, 0, s>>>((const double*)G, (double*)T, g_off_dev, t_off_dev,
                                                                     n_dev, nsplit, rel_thr, kept_dev, (c128*)scratch,
                                                                     stride);
  }
  return (int)cudaGetLastError();
}

extern "C" int gtn_gram_rotate(const void* G, void* T, int dtype, const int64_t* g_off_dev,
                               const int64_t* t_off_dev, const int32_t* n_dev, int nprob, int max_n, int nsplit,
                               double rel_thr, double tol, int max_sweeps, int32_t* sweeps_dev, void* stream) {
  if (nprob <= 0) return GTN_OK;
  if (nsplit < 1) return GTN_ERR_BAD_ARG;
  if (max_n > EIG_MAXN) return GTN_ERR_UNSUPPORTED;
  if (dtype != GTN_C128 && dtype != GTN_F64) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = size_t(2) * max_n * (max_n + 1) * 16 + 64;
  static size_t attr = 0;
  if (smem > attr) {
    int e = set_smem(gram_rotate_kernel<true, 3>, smem);
    if (!e) e = set_smem(gram_rotate_kernel<false, 3>, smem);
    if (!e) e = set_smem(gram_rotate_kernel<true, 5>, smem);
    if (!e) e = set_smem(gram_rotate_kernel<false, 5>, smem);
    if (e) return e;
    attr = smem;
  }
  const bool c = dtype == GTN_C128;
#define GTN_ROT_LAUNCH(CP, TS_, GT)                                                                     \
  gram_rotate_kernel<CP, TS_><<<nprob, ROT_T, smem, s>>>((const GT*)G, (GT*)T, g_off_dev, t_off_dev, n_dev, \
                                                         nsplit, rel_thr, tol, max_sweeps, sweeps_dev)
  if (max_n <= 48) { if (c) GTN_ROT_LAUNCH(true, 3, c128); else GTN_ROT_LAUNCH(false, 3, double); }
  else             { if (c) GTN_ROT_LAUNCH(true, 5, c128); else GTN_ROT_LAUNCH(false, 5, double); }
#undef GTN_ROT_LAUNCH
  return (int)cudaGetLastError();
}

/* diagnostic: clock64 stamps of block 0 of the last gtn_chol_whiten ([0..3]: start, factorised, inverted,
 * stored) and gtn_gram_rotate ([4..7]: start, factorised, rotated, stored) launches; synchronises. */
extern "C" int gtn_debug_phase_clocks(long long* host_out8) {
  return (int)cudaMemcpyFromSymbol(host_out8, gtn_phase_clk, sizeof(long long) * 8);
}

