// gtn_gemm_tma.cu -- grouped row-major GEMM on the FP64 tensor cores (DMMA.8x8x4), operands staged by TMA.
//
//   C_g = alpha_g * A_g * B_g + beta_g * C_g      (float64 or complex128)
//
// Same product and group list as gtn_gemm.cu (gtn_grouped_gemm); what changes is how the operand tiles reach
// shared memory and how the CTAs walk the tile grid:
//   * one PRODUCER warp (one elected lane) issues cp.async.bulk.tensor (TMA, SASS UTMALDG) loads of the A tile
//     [BM rows x 128 B] and of BN*ELEM/128 B boxes [BK rows x 128 B] per pipeline stage, completion counted on a
//     `full` mbarrier per stage (expect_tx); the CONSUMER warps (one 32x32 warp tile each) wait on it, run their
//     DMMAs and release the stage through an `empty` mbarrier.  No thread of a consumer warp computes an address
//     for a load, there is no zero-fill predication (the tensor maps carry the true extents: TMA fills
//     out-of-range box elements with zeros) and no block-wide barrier in the main loop.
//   * shared-memory tiles are dense 128-byte rows in the hardware SWIZZLE_128B pattern (16-byte chunk c of row r is
//     stored at chunk c ^ (r & 7)).  Rows / columns of an m8n8k4 fragment are assigned to tile rows / columns
//     through fixed permutations (frag_row / frag_col below) chosen so that the lanes of one shared-memory
//     wavefront (8 lanes for the 16-byte complex loads, 16 lanes for the 8-byte real loads) always fall into
//     distinct banks: no padding bytes, no bank conflicts.
//   * the tile grid of every group is walked in bands of `raster` row tiles, column-major inside a band, so that the
//     CTAs resident at one time cover a square-ish region of C and share their A / B panels in L2.
// Tensor maps (one per group and operand, 3-D: inner extent, rows, batch) are encoded on the host per launch and
// passed as a __grid_constant__ kernel parameter: nothing is uploaded, the launch is capturable in a CUDA graph.
//
// Replaces: the data contraction inside oe.contract (reference __init__.py:2295) and the per-parity-block np.einsum
// of einsum_block (:2781, :2928) for the large sector GEMMs; every large product of the truncated-SVD subspace
// iteration.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gtn_b200.h"

namespace {

constexpr int STAGES = 4;
constexpr int MAXG = GTN_TMA_MAX_GROUPS;

struct TmaMaps {
  CUtensorMap a[MAXG];
  CUtensorMap b[MAXG];
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ int find_group(const gtn_gemm_group* g, int ng, int64_t tile) {
  int lo = 0, hi = ng - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (g[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Fragment index x (0..7: the row of an A / C fragment, the column of a B / C fragment) -> position inside an
// aligned group of 8 (complex) tile rows or columns, and -- real case -- inside a 16-column B box.
//   complex: lanes 0-7 of a wavefront are x in {2q, 2q+1}, t = 0..3; chunk (4ks + t) ^ pos must cover all 8 chunks:
//            pos(2q) = q, pos(2q+1) = q + 4.
//   real A:  lanes 0-15 are x = 0..3 (or 4..7), t = 0..3; chunk (2ks + t/2) ^ pos: pos = 0,2,4,6 (1,3,5,7).
//   real B:  16 columns per box; lanes 0-15: x = 0..3, rows 4ks + t: (chunk bit 2, half) must differ between the four x.
template <bool CPLX>
__device__ __forceinline__ int frag_row(int x) {
  return CPLX ? (((x & 1) << 2) | (x >> 1)) : (((x & 3) << 1) | (x >> 2));
}
template <bool CPLX>
__device__ __forceinline__ int frag_col(int j, int x) {     // column inside the warp's 32-column strip
  if (CPLX) return j * 8 + (((x & 1) << 2) | (x >> 1));
  return (j >> 1) * 16 + ((x & 1) | ((x >> 2) << 1) | ((j & 1) << 2) | (((x >> 1) & 1) << 3));
}

template <bool CPLX, int BM, int BN>
struct TCfg {
  static constexpr int ELEM = CPLX ? 16 : 8;
  static constexpr int BK = 128 / ELEM;                 // K elements per stage: one 128-byte row of A
  static constexpr int NCW = (BM / 32) * (BN / 32);     // consumer warps
  static constexpr int NTHREADS = NCW * 32 + 128;       // + the producer warpgroup (one working lane)
  // registers: the kernel is compiled for 65536 / (threads per SM) per thread; the producer warpgroup hands its
  // share back (setmaxnreg.dec 40) and every consumer warpgroup grows by the same total (setmaxnreg.inc)
  static constexpr int MINB = (BM == 128) ? 1 : 2;
  static constexpr int REG_BASE = ((65536 / (NTHREADS * MINB)) / 8) * 8;          // 168 (384 threads) / 128 (2 x 256)
  static constexpr int REG_CONS = ((REG_BASE + (REG_BASE - 40) * 4 / (NCW / 4 * 4)) / 8) * 8;   // 232 / 216
  static constexpr int A_BYTES = BM * 128;
  static constexpr int BOX_COLS = 128 / ELEM;           // columns of B per 128-byte box row
  static constexpr int NBOX = BN / BOX_COLS;
  static constexpr int BOX_BYTES = BK * 128;
  static constexpr int B_BYTES = NBOX * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 2 * STAGES * 8;
};

template <bool CPLX, int BM, int BN>
__global__ void __launch_bounds__(TCfg<CPLX, BM, BN>::NTHREADS, TCfg<CPLX, BM, BN>::MINB)
    grouped_gemm_tma_kernel(const __grid_constant__ TmaMaps maps, char* __restrict__ Cbase,
                            const gtn_gemm_group* __restrict__ groups, int ngroups) {
  using C = TCfg<CPLX, BM, BN>;
  constexpr int MT = 4, NT = 4;                          // 32 x 32 warp tile = 4 x 4 fragments of 8 x 8
  extern __shared__ unsigned char smem_raw[];

  const int64_t gtile = blockIdx.x;
  const int gi = find_group(groups, ngroups, gtile);
  const gtn_gemm_group grp = groups[gi];
  const int tiles_m = (grp.m + BM - 1) / BM;
  const int tiles_n = (grp.n + BN - 1) / BN;
  int64_t local = gtile - grp.tile_start;
  const int64_t per_batch = int64_t(tiles_m) * tiles_n;
  const int bidx = int(local / per_batch);
  if (bidx >= grp.batch) return;
  local -= int64_t(bidx) * per_batch;
  // bands of `raster` row tiles, column-major inside a band
  const int G = grp.reserved > 0 ? grp.reserved : 1;
  const int band = int(local / (int64_t(G) * tiles_n));
  const int first = band * G;
  const int rows_in_band = min(G, tiles_m - first);
  const int rem = int(local - int64_t(band) * G * tiles_n);
  const int tm = first + rem % rows_in_band;
  const int tn = rem / rows_in_band;

  const int M = grp.m, N = grp.n, K = grp.k;
  const int m0 = tm * BM, n0 = tn * BN;
  const int ktiles = (K + C::BK - 1) / C::BK;

  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B atoms are 1024-byte aligned
  const uint32_t bars = base + STAGES * C::STAGE_BYTES;  // full[STAGES], empty[STAGES]
  unsigned char* smem = smem_raw + (base - raw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + s * 8, 1);                        // the producer's arrive.expect_tx
      mbar_init(bars + (STAGES + s) * 8, C::NCW);        // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= C::NCW) {
    // ===== producer warpgroup: one lane feeds the ring, the registers go to the consumers =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == C::NCW && lane == 0) {
      const CUtensorMap* mapA = &maps.a[gi];
      const CUtensorMap* mapB = &maps.b[gi];
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(mapB)) : "memory");
      constexpr int W = CPLX ? 2 : 1;                    // map coordinates count doubles
      for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        const uint32_t full = bars + s * 8, empty = bars + (STAGES + s) * 8;
        mbar_wait(empty, ((kt / STAGES) & 1) ^ 1);       // first round: passes at once
        mbar_expect_tx(full, C::STAGE_BYTES);
        const uint32_t sa = base + s * C::STAGE_BYTES;
        const int k0 = kt * C::BK;
        tma_load_3d(sa, mapA, full, k0 * W, m0, bidx);
#pragma unroll
        for (int bx = 0; bx < C::NBOX; ++bx)
          tma_load_3d(sa + C::A_BYTES + bx * C::BOX_BYTES, mapB, full, (n0 + bx * C::BOX_COLS) * W, k0, bidx);
      }
    }
    return;
  }

  // ===== consumers =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::REG_CONS));
  constexpr int WN_WARPS = BN / 32;
  const int wm = warp / WN_WARPS, wn = warp % WN_WARPS;
  const int g = lane >> 2, t = lane & 3;
  const int ra = frag_row<CPLX>(g);                      // this lane's row inside every 8-row group of A

  double cre[MT][NT][2];
  double cim[CPLX ? MT : 1][CPLX ? NT : 1][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      cre[i][j][0] = cre[i][j][1] = 0.0;
      if (CPLX) cim[CPLX ? i : 0][CPLX ? j : 0][0] = cim[CPLX ? i : 0][CPLX ? j : 0][1] = 0.0;
    }

  // per-lane constant parts of the fragment addresses
  int a_row_off[MT];                                     // byte offset of the lane's row in m-fragment i
#pragma unroll
  for (int i = 0; i < MT; ++i) a_row_off[i] = (wm * 32 + i * 8 + ra) * 128;
  int b_box_off[NT], b_chunk[NT], b_half[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int col = wn * 32 + frag_col<CPLX>(j, g);      // column inside the CTA tile
    const int box = col / C::BOX_COLS, cin = col % C::BOX_COLS;
    b_box_off[j] = box * C::BOX_BYTES;
    b_chunk[j] = CPLX ? cin : (cin >> 1);
    b_half[j] = CPLX ? 0 : (cin & 1) * 8;
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    const int s = kt % STAGES;
    mbar_wait(bars + s * 8, (kt / STAGES) & 1);
    const unsigned char* sa = smem + s * C::STAGE_BYTES;
    const unsigned char* sb = sa + C::A_BYTES;
#pragma unroll
    for (int ks = 0; ks < C::BK / 4; ++ks) {
      const int kk = ks * 4 + t;                         // K index inside the stage
      if (CPLX) {
        double are[MT], aim[MT], nim[MT], bre[NT], bim[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(sa + a_row_off[i] + ((kk ^ ra) << 4));
          are[i] = v.x; aim[i] = v.y; nim[i] = -v.y;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double2 v = *reinterpret_cast<const double2*>(sb + b_box_off[j] + kk * 128 + ((b_chunk[j] ^ (kk & 7)) << 4));
          bre[j] = v.x; bim[j] = v.y;
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
          a[i] = *reinterpret_cast<const double*>(sa + a_row_off[i] + (((kk >> 1) ^ ra) << 4) + (kk & 1) * 8);
#pragma unroll
        for (int j = 0; j < NT; ++j)
          b[j] = *reinterpret_cast<const double*>(sb + b_box_off[j] + kk * 128 + ((b_chunk[j] ^ (kk & 7)) << 4) + b_half[j]);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma(cre[i][j][0], cre[i][j][1], a[i], b[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + (STAGES + s) * 8);  // this warp is done with the stage
  }

  // epilogue: the lane owns rows frag_row(g) of every m-fragment and columns frag_col(j, 2t), frag_col(j, 2t + 1)
  const double alpha = grp.alpha, beta = grp.beta;
  char* Cp = Cbase + (grp.c_off + int64_t(bidx) * grp.batch_stride_c) * C::ELEM;
  const int64_t ldc = grp.ldc;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int r = m0 + wm * 32 + i * 8 + ra;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn * 32 + frag_col<CPLX>(j, 2 * t + e);
        if (c >= N) continue;
        char* dst = Cp + (int64_t(r) * ldc + c) * C::ELEM;
        if (CPLX) {
          double2 v = make_double2(alpha * cre[i][j][e], alpha * cim[CPLX ? i : 0][CPLX ? j : 0][e]);
          if (beta != 0.0) {
            const double2 o = *reinterpret_cast<const double2*>(dst);
            v.x += beta * o.x; v.y += beta * o.y;
          }
          *reinterpret_cast<double2*>(dst) = v;
        } else {
          double v = alpha * cre[i][j][e];
          if (beta != 0.0) v += beta * *reinterpret_cast<const double*>(dst);
          *reinterpret_cast<double*>(dst) = v;
        }
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    (void)cudaGetLastError();
  }
  return fn;
}

// 3-D map over a row-major [batch][rows][inner] operand (strides in elements), box = 128 bytes x box_rows x 1
int encode(CUtensorMap* map, const char* base, int elem, int64_t inner, int64_t rows, int64_t ld, int64_t batch,
           int64_t batch_stride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return GTN_ERR_UNSUPPORTED;
  const int w = elem / 8;                                   // doubles per element
  if (batch < 1) batch = 1;
  if (batch == 1 || batch_stride <= 0) batch_stride = rows * ld;   // any valid stride: the coordinate stays 0
  cuuint64_t dims[3] = {(cuuint64_t)(inner * w), (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * elem), (cuuint64_t)(batch_stride * elem)};
  cuuint32_t box[3] = {16u, (cuuint32_t)box_rows, 1u};      // 16 doubles = 128 bytes
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return GTN_ERR_UNSUPPORTED;
  if (strides[0] >= (1ull << 40) || strides[1] >= (1ull << 40)) return GTN_ERR_UNSUPPORTED;
  if (dims[0] == 0 || dims[1] == 0) return GTN_ERR_BAD_ARG;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<char*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GTN_OK : GTN_ERR_UNSUPPORTED;
}

template <bool CPLX, int BM, int BN>
int launch(const TmaMaps& maps, void* Cm, const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles,
           cudaStream_t s) {
  using C = TCfg<CPLX, BM, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(grouped_gemm_tma_kernel<CPLX, BM, BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  grouped_gemm_tma_kernel<CPLX, BM, BN><<<dim3((unsigned)total_tiles), dim3(C::NTHREADS), C::SMEM, s>>>(
      maps, (char*)Cm, groups_dev, ngroups);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int gtn_gemm_tma_check(const gtn_gemm_group* g, int ngroups, int dtype) {
  if (ngroups < 1 || ngroups > MAXG) return 0;
  if (dtype != GTN_C128 && dtype != GTN_F64) return 0;
  const int64_t unit = dtype == GTN_C128 ? 1 : 2;           // element counts that keep 16-byte alignment
  for (int i = 0; i < ngroups; ++i) {
    if (g[i].flags != 0 || g[i].m < 1 || g[i].n < 1 || g[i].k < 1 || g[i].batch < 1) return 0;
    if (g[i].a_off % unit || g[i].b_off % unit || g[i].lda % unit || g[i].ldb % unit) return 0;
    if (g[i].batch > 1 && (g[i].batch_stride_a % unit || g[i].batch_stride_b % unit)) return 0;
    if (g[i].batch > 1 && (g[i].batch_stride_a <= 0 || g[i].batch_stride_b <= 0)) return 0;
  }
  return 1;
}

extern "C" int gtn_grouped_gemm_tma(const void* A, const void* B, void* Cm, int dtype, const gtn_gemm_group* groups_host,
                                    const gtn_gemm_group* groups_dev, int ngroups, int64_t total_tiles, int config,
                                    void* stream) {
  if (ngroups <= 0 || total_tiles <= 0) return GTN_OK;
  if (total_tiles > 2147483647LL || ngroups > MAXG) return GTN_ERR_BAD_ARG;
  if (!(config & 4)) return GTN_ERR_BAD_ARG;
  const bool cplx = dtype == GTN_C128;
  if (!cplx && dtype != GTN_F64) return GTN_ERR_BAD_ARG;
  const int elem = cplx ? 16 : 8;
  const int BM = (config & 8) ? 128 : 64;
  const int BK = 128 / elem;
  TmaMaps maps;
  for (int i = 0; i < ngroups; ++i) {
    const gtn_gemm_group& g = groups_host[i];
    int rc = encode(&maps.a[i], (const char*)A + g.a_off * elem, elem, g.k, g.m, g.lda, g.batch, g.batch_stride_a, BM);
    if (rc) return rc;
    rc = encode(&maps.b[i], (const char*)B + g.b_off * elem, elem, g.n, g.k, g.ldb, g.batch, g.batch_stride_b, BK);
    if (rc) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (cplx) {
    if (BM == 128) return launch<true, 128, 64>(maps, Cm, groups_dev, ngroups, total_tiles, s);
    return launch<true, 64, 64>(maps, Cm, groups_dev, ngroups, total_tiles, s);
  }
  if (BM == 128) return launch<false, 128, 64>(maps, Cm, groups_dev, ngroups, total_tiles, s);
  return launch<false, 64, 64>(maps, Cm, groups_dev, ngroups, total_tiles, s);
}
