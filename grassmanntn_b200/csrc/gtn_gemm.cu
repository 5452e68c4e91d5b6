// gtn_gemm.cu -- grouped row-major GEMM on the FP64 tensor cores (DMMA.8x8x4) for sm_100a.
//
//   C_g = alpha_g * A_g * B_g + beta_g * C_g      (float64 or complex128)
//
// Blackwell's tcgen05/TMEM path has no FP64 kind, so the FP64 tensor pipe is reached through
// warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  complex128 products are expanded in
// registers into four real DMMAs per (m8,n8,k4) step:
//     Cre += Are*Bre ; Cre += (-Aim)*Bim ; Cim += Are*Bim ; Cim += Aim*Bre
// with ONE 16-byte shared-memory load per complex fragment element (re and im arrive together).
//
// CTA tile 64x64, 4 warps (2x2), warp tile 32x32, K step 128 bytes of A row per stage
// (8 complex / 16 real), 3-stage cp.async (LDGSTS) pipeline with zero-fill predication so any
// M, N, K (the block format has odd sector sizes, e.g. 13 and 12) is handled in-kernel.
// Shared-memory row strides are chosen so every fragment load is bank-conflict free:
//   A rows: 128 B payload + 64 B (complex) / 32 B (real) pad; B rows: 64 elems + 32 B pad.
// Groups (one per output parity block in einsum_block, __init__.py:2688-2790) are scheduled in
// one grid through a prefix sum of tile counts.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gtn_b200.h"

namespace {

constexpr int STAGES = 3, NTHREADS = 128;

// two tile configurations: 64x64 (K step = 128 B of A row) for the large sector GEMMs, and
// 32x32 with a 4x deeper K step for the skinny (l x p x q, l ~ 40) products of the randomized
// subspace iteration, which would otherwise occupy 8 CTAs per matrix and be latency bound.
template <bool CPLX, int BM_, int BN_, int AROW_>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_;
  static constexpr int ELEM = CPLX ? 16 : 8;             // bytes per element
  static constexpr int BK = AROW_ / ELEM;                // elements of K per stage
  static constexpr int A_STRIDE = AROW_ + (CPLX ? 64 : 32);  // bytes per A row in smem
  static constexpr int B_STRIDE = BN * ELEM + 32;        // bytes per B row in smem
  static constexpr int A_BYTES = BM * A_STRIDE;
  static constexpr int B_BYTES = BK * B_STRIDE;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES;
  static constexpr int MT = BM / 16, NT = BN / 16;       // m8 / n8 tiles per warp (2 x 2 warps)
};

__device__ __forceinline__ void cp_async(uint32_t dst, const void* src, int bytes16, bool pred) {
  // bytes16: 16 -> cp.async.cg 16B ; 8 -> cp.async.ca 8B.  src-size 0 zero-fills.
  if (bytes16 == 16) {
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
  } else {
    int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
  }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int find_group(const gtn_gemm_group* g, int ng, int64_t tile) {
  int lo = 0, hi = ng - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (g[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// peer output buffers of the fused GEMM + all-gather (multi-GPU output-tile sharding): every rank's
// copy of C, mapped into this process (NVLink peer memory)
struct PeerBufs {
  int n;
  char* p[GTN_MAX_PEERS];
};

template <bool CPLX, int BM, int BN, int AROW, bool BCAST, bool BTRANS = false>
__global__ void __launch_bounds__(NTHREADS)
    grouped_gemm_kernel(const char* __restrict__ Abase, const char* __restrict__ Bbase,
                        char* __restrict__ Cbase, const gtn_gemm_group* __restrict__ groups,
                        int ngroups, const PeerBufs peers) {
  using C = Cfg<CPLX, BM, BN, AROW>;
  constexpr int MT = C::MT, NT = C::NT, WM = BM / 2, WN = BN / 2;
  extern __shared__ __align__(128) unsigned char smem[];

  const int64_t gtile = blockIdx.x;
  const int gi = find_group(groups, ngroups, gtile);
  const gtn_gemm_group grp = groups[gi];
  const int tiles_m = (grp.m + BM - 1) / BM;
  const int tiles_n = (grp.n + BN - 1) / BN;
  int64_t local = gtile - grp.tile_start;
  const int tn = int(local % tiles_n);
  local /= tiles_n;
  const int tm = int(local % tiles_m);
  const int bidx = int(local / tiles_m);
  if (bidx >= grp.batch) return;

  const int M = grp.m, N = grp.n, K = grp.k;
  const int m0 = tm * BM, n0 = tn * BN;
  const char* A = Abase + (grp.a_off + int64_t(bidx) * grp.batch_stride_a) * C::ELEM;
  const char* B = Bbase + (grp.b_off + int64_t(bidx) * grp.batch_stride_b) * C::ELEM;
  const int64_t c_byte_off = (grp.c_off + int64_t(bidx) * grp.batch_stride_c) * C::ELEM;
  char* Cp = Cbase + c_byte_off;
  const int64_t lda = grp.lda, ldb = grp.ldb, ldc = grp.ldc;
  constexpr bool b_trans = BTRANS;     // whole launch: every group carries GTN_GEMM_B_CONJ_TRANS (config bit 1)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;   // 2 x 2 warps
  const int g = lane >> 2, t = lane & 3;

  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const int ktiles = (K + C::BK - 1) / C::BK;

  auto load_stage = [&](int stage, int kt) {
    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
    const uint32_t sb = sa + C::A_BYTES;
    const int k0 = kt * C::BK;
    // A tile: BM rows x BK elems.  BK elems = 128 B contiguous per row.
    constexpr int A_PER_ROW = C::BK;                       // one cp.async per element
    constexpr int A_TOTAL = BM * A_PER_ROW;
#pragma unroll
    for (int i = tid; i < A_TOTAL; i += NTHREADS) {
      const int r = i / A_PER_ROW, kk = i % A_PER_ROW;
      const bool p = (m0 + r < M) && (k0 + kk < K);
      const char* src = p ? A + (int64_t(m0 + r) * lda + (k0 + kk)) * C::ELEM : A;
      cp_async(sa + r * C::A_STRIDE + kk * C::ELEM, src, C::ELEM, p);
    }
    constexpr int B_TOTAL = C::BK * BN;
    if constexpr (!b_trans) {
#pragma unroll
      for (int i = tid; i < B_TOTAL; i += NTHREADS) {
        const int kk = i / BN, c = i % BN;
        const bool p = (k0 + kk < K) && (n0 + c < N);
        const char* src = p ? B + (int64_t(k0 + kk) * ldb + (n0 + c)) * C::ELEM : B;
        cp_async(sb + kk * C::B_STRIDE + c * C::ELEM, src, C::ELEM, p);
      }
    } else {
      // B given as its conjugate transpose (row-major N x K): consecutive threads walk K, the contiguous
      // direction of the source, and scatter into the K-major tile; the conjugation happens at fragment load
#pragma unroll
      for (int i = tid; i < B_TOTAL; i += NTHREADS) {
        const int c = i / C::BK, kk = i % C::BK;
        const bool p = (k0 + kk < K) && (n0 + c < N);
        const char* src = p ? B + (int64_t(n0 + c) * ldb + (k0 + kk)) * C::ELEM : B;
        cp_async(sb + kk * C::B_STRIDE + c * C::ELEM, src, C::ELEM, p);
      }
    }
  };

  // accumulators: 4 (m8) x 4 (n8) tiles, 2 doubles each, re (+ im)
  double cre[MT][NT][2];
  double cim[CPLX ? MT : 1][CPLX ? NT : 1][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      cre[i][j][0] = cre[i][j][1] = 0.0;
      if (CPLX) cim[CPLX ? i : 0][CPLX ? j : 0][0] = cim[CPLX ? i : 0][CPLX ? j : 0][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage(nk % STAGES, nk);
      cp_commit();
    }
    const unsigned char* sa = smem + (kt % STAGES) * C::STAGE_BYTES;
    const unsigned char* sb = sa + C::A_BYTES;
#pragma unroll
    for (int ks = 0; ks < C::BK / 4; ++ks) {
      if (CPLX) {
        double are[MT], aim[MT], nim[MT], bre[NT], bim[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(
              sa + (wm * WM + i * 8 + g) * C::A_STRIDE + (ks * 4 + t) * 16);
          are[i] = v.x; aim[i] = v.y; nim[i] = -v.y;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double2 v = *reinterpret_cast<const double2*>(
              sb + (ks * 4 + t) * C::B_STRIDE + (wn * WN + j * 8 + g) * 16);
          bre[j] = v.x; bim[j] = b_trans ? -v.y : v.y;
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
            dmma(cim[CPLX ? i : 0][CPLX ? j : 0][0], cim[CPLX ? i : 0][CPLX ? j : 0][1], are[i], bim[j]);
            dmma(cre[i][j][0], cre[i][j][1], nim[i], bim[j]);
            dmma(cim[CPLX ? i : 0][CPLX ? j : 0][0], cim[CPLX ? i : 0][CPLX ? j : 0][1], aim[i], bre[j]);
          }
      } else {
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i)
          a[i] = *reinterpret_cast<const double*>(sa + (wm * WM + i * 8 + g) * C::A_STRIDE + (ks * 4 + t) * 8);
#pragma unroll
        for (int j = 0; j < NT; ++j)
          b[j] = *reinterpret_cast<const double*>(sb + (ks * 4 + t) * C::B_STRIDE + (wn * WN + j * 8 + g) * 8);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma(cre[i][j][0], cre[i][j][1], a[i], b[j]);
      }
    }
  }
  cp_wait<0>();

  // epilogue: thread owns rows (wm*WM + i*8 + g), cols (wn*WN + j*8 + 2t, +1)
  const double alpha = grp.alpha, beta = grp.beta;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int r = m0 + wm * WM + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c = n0 + wn * WN + j * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (c + e >= N) continue;
        const int64_t eoff = (int64_t(r) * ldc + (c + e)) * C::ELEM;
        char* dst = Cp + eoff;
        if (CPLX) {
          double2 v = make_double2(alpha * cre[i][j][e], alpha * cim[CPLX ? i : 0][CPLX ? j : 0][e]);
          if constexpr (BCAST) {
            // the result tile goes straight into every rank's copy of C: the NVLink stores overlap the
            // DMMA work of the other resident CTAs, no separate all-gather pass over the output
#pragma unroll 1
            for (int pr = 0; pr < peers.n; ++pr)
              *reinterpret_cast<double2*>(peers.p[pr] + c_byte_off + eoff) = v;
          } else {
            if (beta != 0.0) {
              const double2 o = *reinterpret_cast<const double2*>(dst);
              v.x += beta * o.x; v.y += beta * o.y;
            }
            *reinterpret_cast<double2*>(dst) = v;
          }
        } else {
          double v = alpha * cre[i][j][e];
          if constexpr (BCAST) {
#pragma unroll 1
            for (int pr = 0; pr < peers.n; ++pr)
              *reinterpret_cast<double*>(peers.p[pr] + c_byte_off + eoff) = v;
          } else {
            if (beta != 0.0) v += beta * *reinterpret_cast<const double*>(dst);
            *reinterpret_cast<double*>(dst) = v;
          }
        }
      }
    }
  }
}

template <bool CPLX, int BM, int BN, int AROW, bool BCAST, bool BTRANS = false>
int launch_cfg(const void* A, const void* B, void* C, const gtn_gemm_group* groups_dev, int ngroups,
               int64_t total_tiles, const PeerBufs& peers, cudaStream_t s) {
  using K = Cfg<CPLX, BM, BN, AROW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(grouped_gemm_kernel<CPLX, BM, BN, AROW, BCAST, BTRANS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  grouped_gemm_kernel<CPLX, BM, BN, AROW, BCAST, BTRANS><<<dim3((unsigned)total_tiles), dim3(NTHREADS), K::SMEM, s>>>(
      (const char*)A, (const char*)B, (char*)C, groups_dev, ngroups, peers);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int64_t gtn_gemm_plan_host(gtn_gemm_group* groups, int ngroups, int dtype, int config) {
  (void)dtype;
  // bit 0: 32x32 deep-K tiles; bit 2: TMA-staged kernel (gtn_gemm_tma.cu), bit 3 with it: 128-row tiles
  const int64_t BM = (config & 8) ? 128 : ((config & 1) ? 32 : 64), BN = (config & 1) ? 32 : 64;
  int64_t acc = 0;
  for (int i = 0; i < ngroups; ++i) {
    groups[i].tile_start = acc;
    const int64_t tm = (groups[i].m + BM - 1) / BM, tn = (groups[i].n + BN - 1) / BN;
    const int64_t b = groups[i].batch > 0 ? groups[i].batch : 0;
    acc += tm * tn * b;
  }
  return acc;
}

extern "C" int gtn_grouped_gemm(const void* A, const void* B, void* C, int dtype,
                                const gtn_gemm_group* groups_dev, int ngroups,
                                int64_t total_tiles, int config, void* stream) {
  if (ngroups <= 0 || total_tiles <= 0) return GTN_OK;
  if (total_tiles > 2147483647LL) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  PeerBufs none; none.n = 0;
  // config bit 0: 32x32 deep-K tiles; bit 1: every group's B is given as its conjugate transpose
  // (GTN_GEMM_B_CONJ_TRANS; only with the 32x32 configuration, where the Gram matrices of the subspace iteration live)
  if (dtype == GTN_C128) {
    if (config == 3) return launch_cfg<true, 32, 32, 512, false, true>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
    if (config == 1) return launch_cfg<true, 32, 32, 512, false>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
    if (config == 0) return launch_cfg<true, 64, 64, 128, false>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
  } else if (dtype == GTN_F64) {
    if (config == 3) return launch_cfg<false, 32, 32, 512, false, true>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
    if (config == 1) return launch_cfg<false, 32, 32, 512, false>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
    if (config == 0) return launch_cfg<false, 64, 64, 128, false>(A, B, C, groups_dev, ngroups, total_tiles, none, s);
  }
  return GTN_ERR_BAD_ARG;
}

extern "C" int gtn_grouped_gemm_bcast(const void* A, const void* B, void* const* C_peers, int npeers, int dtype,
                                      const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles,
                                      void* stream) {
  if (ngroups <= 0 || total_tiles <= 0) return GTN_OK;
  if (total_tiles > 2147483647LL || npeers < 1 || npeers > GTN_MAX_PEERS) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  PeerBufs peers; peers.n = npeers;
  for (int i = 0; i < npeers; ++i) peers.p[i] = (char*)C_peers[i];
  if (dtype == GTN_C128)
    return launch_cfg<true, 64, 64, 128, true>(A, B, nullptr, groups_dev, ngroups, total_tiles, peers, s);
  if (dtype == GTN_F64)
    return launch_cfg<false, 64, 64, 128, true>(A, B, nullptr, groups_dev, ngroups, total_tiles, peers, s);
  return GTN_ERR_BAD_ARG;
}
