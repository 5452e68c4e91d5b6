// gtn_misc.cu -- small reductions / element-wise helpers of the Grassmann hot path (sm_100a).
// Each is a single streaming pass (HBM-bound, grid sized to the SM count, grid-stride loops,
// 16-byte loads for complex128).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gtn_b200.h"

namespace {

constexpr int RT = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[RT / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < RT / 32) r = red[threadIdx.x];
  if (threadIdx.x < 32) r = warp_sum(r);
  __syncthreads();
  return r;   // valid in thread 0
}

constexpr int SUMSQ_SLOTS = 16, SUMSQ_MAXGRID = 148 * 8;
__device__ double g_sumsq_part[SUMSQ_SLOTS][SUMSQ_MAXGRID];
__device__ unsigned int g_sumsq_count[SUMSQ_SLOTS];

__global__ void __launch_bounds__(RT) sumsq_kernel(const double* __restrict__ x, int64_t ndoubles,
                                                   double* __restrict__ out, double* __restrict__ part,
                                                   unsigned int* __restrict__ counter) {
  double acc = 0;
  const int64_t stride = int64_t(gridDim.x) * RT;
  const int64_t n2 = ndoubles >> 1;
  const double2* x2 = reinterpret_cast<const double2*>(x);
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n2; i += stride) {
    const double2 v = __ldg(x2 + i);
    acc += v.x * v.x + v.y * v.y;
  }
  if ((ndoubles & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc += x[ndoubles - 1] * x[ndoubles - 1];
  acc = block_sum(acc);
  // Deterministic completion: every CTA parks its partial sum, the CTA that arrives last adds all of them in index
  // order and accumulates into `out` -- the result does not depend on the order in which the CTAs finish (an
  // atomicAdd per CTA made the last bits of Tnorm, and with them every later step, differ from run to run).  The
  // scratch belongs to a slot chosen by the launching stream, so launches on different streams do not share one.
  __shared__ bool last;
  if (threadIdx.x == 0) {
    part[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double tot = 0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += RT) tot += *reinterpret_cast<volatile double*>(part + b);
  // (fixed assignment of partials to threads and a fixed reduction tree: order independent of timing)
  tot = block_sum(tot);
  if (threadIdx.x == 0) { *out += tot; *counter = 0u; }
}

// sum of |x_i| (complex modulus for GTN_C128): the reference's Grassmann-evenness test is an L1 mean
template <bool CPLX>
__global__ void __launch_bounds__(RT) sumabs_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  double acc = 0;
  const int64_t stride = int64_t(gridDim.x) * RT;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n; i += stride) {
    if (CPLX) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(x) + i);
      acc += hypot(v.x, v.y);
    } else {
      acc += fabs(__ldg(x + i));
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

template <bool CPLX>
__global__ void __launch_bounds__(RT) rowsum_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                    int64_t rows, int64_t cols) {
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  double ar = 0, ai = 0;
  if (CPLX) {
    const double2* p = reinterpret_cast<const double2*>(x) + r * cols;
    for (int64_t c = threadIdx.x; c < cols; c += RT) { const double2 v = __ldg(p + c); ar += v.x; ai += v.y; }
  } else {
    const double* p = x + r * cols;
    for (int64_t c = threadIdx.x; c < cols; c += RT) ar += __ldg(p + c);
  }
  ar = block_sum(ar);
  if (CPLX) ai = block_sum(ai);
  if (threadIdx.x == 0) {
    if (CPLX) { y[2 * r] = ar; y[2 * r + 1] = ai; } else y[r] = ar;
  }
}

// out[r] = sum_c |x[r, c]|^2
template <bool CPLX>
__global__ void __launch_bounds__(RT) row_sumsq_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                       int64_t rows, int64_t cols) {
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  double a = 0;
  if (CPLX) {
    const double2* p = reinterpret_cast<const double2*>(x) + r * cols;
    for (int64_t c = threadIdx.x; c < cols; c += RT) { const double2 v = __ldg(p + c); a += v.x * v.x + v.y * v.y; }
  } else {
    const double* p = x + r * cols;
    for (int64_t c = threadIdx.x; c < cols; c += RT) { const double v = __ldg(p + c); a += v * v; }
  }
  a = block_sum(a);
  if (threadIdx.x == 0) y[r] = a;
}

// out = sum_i a_i * b_i (no conjugation): the fully contracted Grassmann einsum ('ijkl,klij', the
// trace checks of gauge2d.py:1738, :1856) after both operands have been packed with their signs.
// Deterministic two-stage reduction: per-CTA partials, then one CTA.
template <bool CPLX>
__global__ void __launch_bounds__(RT) dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                         int64_t n, double* __restrict__ partial) {
  double re = 0, im = 0;
  const int64_t stride = int64_t(gridDim.x) * RT;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n; i += stride) {
    if (CPLX) {
      const double2 x = __ldg(reinterpret_cast<const double2*>(a) + i);
      const double2 y = __ldg(reinterpret_cast<const double2*>(b) + i);
      re += x.x * y.x - x.y * y.y;
      im += x.x * y.y + x.y * y.x;
    } else {
      re += __ldg(a + i) * __ldg(b + i);
    }
  }
  re = block_sum(re);
  if (CPLX) im = block_sum(im);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = re; partial[2 * blockIdx.x + 1] = im; }
}

template <bool CPLX>
__global__ void __launch_bounds__(RT) dot_final_kernel(const double* __restrict__ partial, int nparts,
                                                       double* __restrict__ out) {
  double re = 0, im = 0;
  for (int i = threadIdx.x; i < nparts; i += RT) { re += partial[2 * i]; im += partial[2 * i + 1]; }
  re = block_sum(re);
  im = block_sum(im);
  if (threadIdx.x == 0) { out[0] = re; if (CPLX) out[1] = im; }
}

// complex power: principal branch of z^p, like numpy.power on complex128
__device__ __forceinline__ void cpow(double& re, double& im, double p) {
  const double r = hypot(re, im);
  const double th = atan2(im, re);
  const double rp = pow(r, p);
  double s, c;
  sincos(p * th, &s, &c);
  re = rp * c;
  im = rp * s;
}

template <bool CPLX>
__global__ void __launch_bounds__(RT) pow_kernel(double* __restrict__ x, int64_t n, double p, double rcond) {
  const int64_t stride = int64_t(gridDim.x) * RT;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n; i += stride) {
    if (CPLX) {
      double2 v = reinterpret_cast<double2*>(x)[i];
      if (hypot(v.x, v.y) > rcond) {
        if (v.y == 0.0 && v.x >= 0.0) { v.x = pow(v.x, p); v.y = 0.0; }
        else cpow(v.x, v.y, p);
      } else { v.x = 0.0; v.y = 0.0; }
      reinterpret_cast<double2*>(x)[i] = v;
    } else {
      const double v = x[i];
      x[i] = fabs(v) > rcond ? pow(v, p) : 0.0;
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(RT) scale_kernel(double* __restrict__ x, int64_t n, double sr, double si) {
  const int64_t stride = int64_t(gridDim.x) * RT;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n; i += stride) {
    if (CPLX) {
      double2 v = reinterpret_cast<double2*>(x)[i];
      const double a = v.x * sr - v.y * si, b = v.x * si + v.y * sr;
      reinterpret_cast<double2*>(x)[i] = make_double2(a, b);
    } else {
      x[i] *= sr;
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(RT) odd_checker_kernel(const double* __restrict__ x, int64_t rows, int64_t cols,
                                                         double* __restrict__ out) {
  const int64_t n = rows * cols, stride = int64_t(gridDim.x) * RT;
  double m = 0;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    if ((r ^ c) & 1) {
      const double a = CPLX ? hypot(x[2 * i], x[2 * i + 1]) : fabs(x[i]);
      m = fmax(m, a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0)
    atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

int grid_for(int64_t n) {
  int64_t g = (n + RT - 1) / RT;
  const int64_t cap = 148 * 8;   // 8 resident CTAs of 256 threads per SM, 148 SMs
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int gtn_sumsq(const void* x, int64_t n, int dtype, double* out_dev, int zero_first, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (zero_first) cudaMemsetAsync(out_dev, 0, sizeof(double), s);
  if (n <= 0) return GTN_OK;
  const int64_t nd = dtype == GTN_C128 ? 2 * n : n;
  // scratch slot of this stream (launches on one stream are ordered; different streams get different slots unless
  // more than SUMSQ_SLOTS streams hash alike)
  static double* part_base = nullptr;
  static unsigned int* count_base = nullptr;
  if (!part_base) {
    cudaGetSymbolAddress((void**)&part_base, g_sumsq_part);
    cudaGetSymbolAddress((void**)&count_base, g_sumsq_count);
  }
  const uintptr_t h = reinterpret_cast<uintptr_t>(stream);
  const int slot = (int)(((h >> 4) ^ (h >> 12) ^ (h >> 20)) % SUMSQ_SLOTS);
  sumsq_kernel<<<grid_for(nd / 2 + 1), RT, 0, s>>>((const double*)x, nd, out_dev, part_base + slot * SUMSQ_MAXGRID,
                                                   count_base + slot);
  return (int)cudaGetLastError();
}

extern "C" int gtn_sumabs(const void* x, int64_t n, int dtype, double* out_dev, int zero_first, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (zero_first) cudaMemsetAsync(out_dev, 0, sizeof(double), s);
  if (n <= 0) return GTN_OK;
  if (dtype == GTN_C128) sumabs_kernel<true><<<grid_for(n), RT, 0, s>>>((const double*)x, n, out_dev);
  else if (dtype == GTN_F64) sumabs_kernel<false><<<grid_for(n), RT, 0, s>>>((const double*)x, n, out_dev);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_rowsum(const void* x, void* y, int64_t rows, int64_t cols, int dtype, void* stream) {
  if (rows <= 0) return GTN_OK;
  if (rows > 2147483647LL) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) rowsum_kernel<true><<<(unsigned)rows, RT, 0, s>>>((const double*)x, (double*)y, rows, cols);
  else if (dtype == GTN_F64) rowsum_kernel<false><<<(unsigned)rows, RT, 0, s>>>((const double*)x, (double*)y, rows, cols);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

// y[i] = alpha * sum_s x[s * n + i] (doubles; slices added in index order: deterministic) -- completes a split-K GEMM
__global__ void __launch_bounds__(RT) sum_slices_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                        int64_t nd, int nslices) {
  const int64_t stride = int64_t(gridDim.x) * RT;
  const int64_t n2 = nd >> 1;
  for (int64_t i = int64_t(blockIdx.x) * RT + threadIdx.x; i < n2; i += stride) {
    double2 acc = __ldg(reinterpret_cast<const double2*>(x) + i);
    for (int s = 1; s < nslices; ++s) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(x + int64_t(s) * nd) + i);
      acc.x += v.x; acc.y += v.y;
    }
    reinterpret_cast<double2*>(y)[i] = acc;
  }
  if ((nd & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double acc = x[nd - 1];
    for (int s = 1; s < nslices; ++s) acc += x[int64_t(s) * nd + nd - 1];
    y[nd - 1] = acc;
  }
}

extern "C" int gtn_sum_slices(const void* x, void* y, int64_t n, int nslices, int dtype, void* stream) {
  if (n <= 0 || nslices < 1) return GTN_OK;
  if (dtype != GTN_C128 && dtype != GTN_F64) return GTN_ERR_BAD_ARG;
  const int64_t nd = dtype == GTN_C128 ? 2 * n : n;
  if ((nd & 1) && nslices > 1) return GTN_ERR_BAD_ARG;      // 16-byte loads: slices must start 16-byte aligned
  sum_slices_kernel<<<grid_for(nd / 2 + 1), RT, 0, (cudaStream_t)stream>>>((const double*)x, (double*)y, nd, nslices);
  return (int)cudaGetLastError();
}

// G_b (sum of nsplit slices) gets rel * trace(G_b) added to its diagonal (slice 0): the shift of a shifted Cholesky QR.
template <bool CPLX>
__global__ void __launch_bounds__(RT) gram_shift_kernel(double* __restrict__ G, const int64_t* __restrict__ g_off,
                                                        const int32_t* __restrict__ ns, int nsplit, double rel) {
  const int n = ns[blockIdx.x];
  const int64_t es = CPLX ? 2 : 1;
  double* base = G + g_off[blockIdx.x] * es;
  double tr = 0;
  for (int i = threadIdx.x; i < n; i += RT)
    for (int sp = 0; sp < nsplit; ++sp) tr += base[(int64_t(sp) * n * n + int64_t(i) * n + i) * es];
  tr = block_sum(tr);
  __shared__ double tot;
  if (threadIdx.x == 0) tot = tr;
  __syncthreads();
  const double add = rel * tot;
  for (int i = threadIdx.x; i < n; i += RT) base[(int64_t(i) * n + i) * es] += add;
}

extern "C" int gtn_gram_shift(void* G, int dtype, const int64_t* g_off_dev, const int32_t* n_dev, int nprob, int nsplit,
                              double rel_shift, void* stream) {
  if (nprob <= 0) return GTN_OK;
  if (nsplit < 1) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) gram_shift_kernel<true><<<nprob, RT, 0, s>>>((double*)G, g_off_dev, n_dev, nsplit, rel_shift);
  else if (dtype == GTN_F64) gram_shift_kernel<false><<<nprob, RT, 0, s>>>((double*)G, g_off_dev, n_dev, nsplit, rel_shift);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_row_sumsq(const void* x, double* y, int64_t rows, int64_t cols, int dtype, void* stream) {
  if (rows <= 0) return GTN_OK;
  if (rows > 2147483647LL) return GTN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) row_sumsq_kernel<true><<<(unsigned)rows, RT, 0, s>>>((const double*)x, y, rows, cols);
  else if (dtype == GTN_F64) row_sumsq_kernel<false><<<(unsigned)rows, RT, 0, s>>>((const double*)x, y, rows, cols);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_dot(const void* a, const void* b, int64_t n, int dtype, void* out, double* partial_dev,
                       int nparts, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (nparts < 1) return GTN_ERR_BAD_ARG;
  int g = grid_for(n);
  if (g > nparts) g = nparts;
  if (dtype == GTN_C128) {
    dot_partial_kernel<true><<<g, RT, 0, s>>>((const double*)a, (const double*)b, n, partial_dev);
    dot_final_kernel<true><<<1, RT, 0, s>>>(partial_dev, g, (double*)out);
  } else if (dtype == GTN_F64) {
    dot_partial_kernel<false><<<g, RT, 0, s>>>((const double*)a, (const double*)b, n, partial_dev);
    dot_final_kernel<false><<<1, RT, 0, s>>>(partial_dev, g, (double*)out);
  } else {
    return GTN_ERR_BAD_ARG;
  }
  return (int)cudaGetLastError();
}

extern "C" int gtn_pow_rcond(void* x, int64_t n, int dtype, double p, double rcond, void* stream) {
  if (n <= 0) return GTN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) pow_kernel<true><<<grid_for(n), RT, 0, s>>>((double*)x, n, p, rcond);
  else if (dtype == GTN_F64) pow_kernel<false><<<grid_for(n), RT, 0, s>>>((double*)x, n, p, rcond);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_scale(void* x, int64_t n, int dtype, double s_re, double s_im, void* stream) {
  if (n <= 0) return GTN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GTN_C128) scale_kernel<true><<<grid_for(n), RT, 0, s>>>((double*)x, n, s_re, s_im);
  else if (dtype == GTN_F64) scale_kernel<false><<<grid_for(n), RT, 0, s>>>((double*)x, n, s_re, 0.0);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_odd_checker(const void* x, int64_t rows, int64_t cols, int dtype, double* out_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out_dev, 0, sizeof(double), s);
  if (rows * cols <= 0) return GTN_OK;
  if (dtype == GTN_C128) odd_checker_kernel<true><<<grid_for(rows * cols), RT, 0, s>>>((const double*)x, rows, cols, out_dev);
  else if (dtype == GTN_F64) odd_checker_kernel<false><<<grid_for(rows * cols), RT, 0, s>>>((const double*)x, rows, cols, out_dev);
  else return GTN_ERR_BAD_ARG;
  return (int)cudaGetLastError();
}

extern "C" int gtn_version(void) { return 100; }
extern "C" const char* gtn_build_arch(void) { return "sm_100a"; }
