// gtn_permute.cu -- fused Grassmann sign + permute kernel for sm_100a.
//
// out[ob + sum_X out_off_X(i_X)] = scale * (-1)^e(i) * maybe_conj(in[ib + sum_X in_off_X(i_X)])
// with e a GF(2) quadratic form of the index parities (see include/gtn_b200.h).  HBM-bound:
// every element is read once and written once (2 * sizeof(elem) algorithmic bytes); the sign
// tensors the reference materialises (__init__.py:1962-1999, :2088-2126) never exist.
//
// Layout: one CTA = one 32x32 tile spanned by super-axis 0 (contiguous on the input side) and
// super-axis 1 (contiguous on the output side); 256 threads = 32 x 8, four elements per
// thread, all four loads in flight before the first shared-memory store.  Rows are read along
// axis 0 (32 x 16 B = 512 B per warp for complex128) and written along axis 1 after a
// shared-memory transpose; the tile is padded by one 16-byte element per row so that both the
// row-wise stores and the column-wise loads are conflict-free at 128 B per wavefront.
// Parity/sign work per element is 6 integer ops: slow super-axes are folded into CTA-uniform
// (P_s, M_s, e_s) once per tile.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gtn_b200.h"

namespace {

constexpr int TILE = 32;
constexpr int ROWS_PER_PASS = 8;   // blockDim.y
constexpr int NPASS = TILE / ROWS_PER_PASS;

struct c128 {
  double re, im;
};

__device__ __forceinline__ double apply(double v, unsigned e, int /*conj*/, double sr, double /*si*/) {
  return (e & 1u) ? -v * sr : v * sr;
}
__device__ __forceinline__ c128 apply(c128 v, unsigned e, int conj, double sr, double si) {
  if (conj) v.im = -v.im;
  if (e & 1u) {
    v.re = -v.re;
    v.im = -v.im;
  }
  c128 r;
  r.re = v.re * sr - v.im * si;
  r.im = v.re * si + v.im * sr;
  return r;
}

__device__ __forceinline__ c128 ldg(const c128* p) {
  double2 t = __ldg(reinterpret_cast<const double2*>(p));
  c128 r;
  r.re = t.x;
  r.im = t.y;
  return r;
}
__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ void stg(c128* p, c128 v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im);
}
__device__ __forceinline__ void stg(double* p, double v) { *p = v; }

template <typename T>
__global__ void __launch_bounds__(TILE* ROWS_PER_PASS)
    sign_permute_kernel(const T* __restrict__ in, T* __restrict__ out,
                        const gtn_permute_job* __restrict__ jobs,
                        const gtn_axis_entry* __restrict__ entries, double sr, double si) {
  const gtn_permute_job& job = jobs[blockIdx.y];
  const int64_t tile_id = blockIdx.x;
  if (tile_id >= job.ntiles) return;

  __shared__ T tile[TILE][TILE + 1];

  const int nA = job.size[0];
  const int nB = job.nsuper > 1 ? job.size[1] : 1;
  const int tilesA = (nA + TILE - 1) / TILE;
  const int tilesB = (nB + TILE - 1) / TILE;

  // decode tile id: fastest = tile along A, then tile along B, then the slow super-axes
  int64_t rem = tile_id;
  const int a0 = int(rem % tilesA) * TILE;
  rem /= tilesA;
  const int b0 = int(rem % tilesB) * TILE;
  rem /= tilesB;

  int64_t in_base = job.in_base, out_base = job.out_base;
  uint32_t Ps = 0, Ms = 0, es = uint32_t(job.const_exp);
  for (int x = 2; x < job.nsuper; ++x) {
    const int sz = job.size[x];
    const int ix = int(rem % sz);
    rem /= sz;
    const gtn_axis_entry en = entries[job.table_start[x] + ix];
    in_base += en.in_off;
    out_base += en.out_off;
    const uint32_t P = en.P & 0x0fffffffu;
    es ^= (en.P >> 31) ^ uint32_t(__popc(P & Ms));
    Ms ^= en.M;
    Ps |= P;
  }
  (void)Ps;

  const gtn_axis_entry* tabA = entries + job.table_start[0];
  const gtn_axis_entry* tabB = entries + (job.nsuper > 1 ? job.table_start[1] : job.table_start[0]);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int conj = job.conj;

  if (!job.transpose) {
    // both sides contiguous along axis 0: straight through registers
    const int a = a0 + tx;
    if (a >= nA) return;
    const gtn_axis_entry ea = tabA[a];
    const uint32_t PA = ea.P & 0x0fffffffu;
    const uint32_t eA = es ^ (ea.P >> 31) ^ uint32_t(__popc(PA & Ms));
    const uint32_t MsA = Ms ^ ea.M;
    T v[NPASS];
    int64_t oaddr[NPASS];
    uint32_t ev[NPASS];
#pragma unroll
    for (int j = 0; j < NPASS; ++j) {
      const int b = b0 + ty + j * ROWS_PER_PASS;
      oaddr[j] = -1;
      if (b < nB) {
        int64_t ia = in_base + ea.in_off;
        int64_t oa = out_base + ea.out_off;
        uint32_t e = eA;
        if (job.nsuper > 1) {
          const gtn_axis_entry eb = tabB[b];
          ia += eb.in_off;
          oa += eb.out_off;
          e ^= (eb.P >> 31) ^ uint32_t(__popc((eb.P & 0x0fffffffu) & MsA));
        }
        v[j] = ldg(in + ia);
        oaddr[j] = oa;
        ev[j] = e;
      }
    }
#pragma unroll
    for (int j = 0; j < NPASS; ++j)
      if (oaddr[j] >= 0) stg(out + oaddr[j], apply(v[j], ev[j], conj, sr, si));
    return;
  }

  // ---- read phase: rows along A (tx), 4 rows of B per thread
  {
    const int a = a0 + tx;
    gtn_axis_entry ea;
    ea.in_off = 0; ea.out_off = 0; ea.P = 0; ea.M = 0;
    if (a < nA) ea = tabA[a];
    const uint32_t PA = ea.P & 0x0fffffffu;
    const uint32_t eA = es ^ (ea.P >> 31) ^ uint32_t(__popc(PA & Ms));
    const uint32_t MsA = Ms ^ ea.M;
    T v[NPASS];
    uint32_t ev[NPASS];
    bool ok[NPASS];
#pragma unroll
    for (int j = 0; j < NPASS; ++j) {
      const int b = b0 + ty + j * ROWS_PER_PASS;
      ok[j] = (a < nA) && (b < nB);
      if (ok[j]) {
        const gtn_axis_entry eb = tabB[b];
        v[j] = ldg(in + (in_base + ea.in_off + eb.in_off));
        ev[j] = eA ^ (eb.P >> 31) ^ uint32_t(__popc((eb.P & 0x0fffffffu) & MsA));
      }
    }
#pragma unroll
    for (int j = 0; j < NPASS; ++j)
      if (ok[j]) tile[ty + j * ROWS_PER_PASS][tx] = apply(v[j], ev[j], conj, sr, si);
  }
  __syncthreads();
  // ---- write phase: rows along B (tx), 4 rows of A per thread
  {
    const int b = b0 + tx;
    if (b >= nB) return;
    const int64_t ob = out_base + tabB[b].out_off;
#pragma unroll
    for (int j = 0; j < NPASS; ++j) {
      const int al = ty + j * ROWS_PER_PASS;
      const int a = a0 + al;
      if (a < nA) stg(out + (ob + tabA[a].out_off), tile[tx][al]);
    }
  }
}

}  // namespace

extern "C" int gtn_sign_permute(const void* in, void* out, int dtype, const gtn_permute_job* jobs,
                                const gtn_axis_entry* entries, int njobs, int64_t max_tiles,
                                double scale_re, double scale_im, void* stream) {
  if (njobs <= 0 || max_tiles <= 0) return GTN_OK;
  if (njobs > 65535 || max_tiles > 2147483647LL) return GTN_ERR_BAD_ARG;
  dim3 grid((unsigned)max_tiles, (unsigned)njobs, 1), block(TILE, ROWS_PER_PASS, 1);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) {
    sign_permute_kernel<c128><<<grid, block, 0, s>>>((const c128*)in, (c128*)out, jobs, entries,
                                                     scale_re, scale_im);
  } else if (dtype == GTN_F64) {
    sign_permute_kernel<double><<<grid, block, 0, s>>>((const double*)in, (double*)out, jobs,
                                                       entries, scale_re, 0.0);
  } else {
    return GTN_ERR_BAD_ARG;
  }
  return (int)cudaGetLastError();
}
