// Sample moments (count, sum x, sum x x^T) of a device-resident sample block, and their cross-rank sum.
// Not in the reference (its specs compare sample means by hand: GaussianAnalyticPosteriorSpec.scala:34-50); the
// statistical parity checks of the samplers need them at sizes where a host pass would dominate.
// Two-stage and atomics-free, so the result is bit-reproducible run to run: stage 1 writes per-slab partial tiles,
// stage 2 adds them in slab order into the accumulator.
#include "kernels.h"
#include "comm.h"
#include <algorithm>
#include <cmath>

namespace thy {

constexpr int MT = 32;  // output tile edge

// grid (tiles, tiles, nslab); block 16x16, each thread owns a 2x2 patch of the 32x32 tile of sum x x^T.
__global__ void __launch_bounds__(256) k_moments_part(long long n, int d, const double* __restrict__ X, int nslab,
                                                      double* __restrict__ part) {
  __shared__ double si[MT][MT + 1], sj[MT][MT + 1];
  __shared__ double valid[MT];
  const int ti = blockIdx.y, tj = blockIdx.x, slab = blockIdx.z;
  const long long r0 = (n * slab) / nslab, r1 = (n * (slab + 1)) / nslab;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t stride = (size_t)1 + d + (size_t)d * d;
  double* out = part + (size_t)slab * stride;
  double a00 = 0, a01 = 0, a10 = 0, a11 = 0, sx = 0, cnt = 0;
  for (long long rb = r0; rb < r1; rb += MT) {
    // stage 32 rows x 32 columns of both column tiles; a row whose first coordinate is NaN counts as absent
    for (int e = threadIdx.x; e < MT * MT; e += 256) {
      const int rr = e / MT, cc = e % MT;
      const long long r = rb + rr;
      const int ci = ti * MT + cc, cj = tj * MT + cc;
      const bool row_ok = r < r1 && !isnan(X[(size_t)r * d]);
      si[rr][cc] = (row_ok && ci < d) ? X[(size_t)r * d + ci] : 0.0;
      sj[rr][cc] = (row_ok && cj < d) ? X[(size_t)r * d + cj] : 0.0;
      if (cc == 0) valid[rr] = row_ok ? 1.0 : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < MT; ++rr) {
      const double x0 = si[rr][ty * 2], x1 = si[rr][ty * 2 + 1];
      const double y0 = sj[rr][tx * 2], y1 = sj[rr][tx * 2 + 1];
      a00 = fma(x0, y0, a00);
      a01 = fma(x0, y1, a01);
      a10 = fma(x1, y0, a10);
      a11 = fma(x1, y1, a11);
    }
    if (ti == tj && threadIdx.x < MT) {
      for (int rr = 0; rr < MT; ++rr) sx += si[rr][threadIdx.x];
    }
    if (ti == 0 && tj == 0 && threadIdx.x == 0) {
      for (int rr = 0; rr < MT; ++rr) cnt += valid[rr];
    }
    __syncthreads();
  }
  const int i0 = ti * MT + ty * 2, j0 = tj * MT + tx * 2;
  double* S2 = out + 1 + d;
  if (i0 < d && j0 < d) S2[(size_t)i0 * d + j0] = a00;
  if (i0 < d && j0 + 1 < d) S2[(size_t)i0 * d + j0 + 1] = a01;
  if (i0 + 1 < d && j0 < d) S2[(size_t)(i0 + 1) * d + j0] = a10;
  if (i0 + 1 < d && j0 + 1 < d) S2[(size_t)(i0 + 1) * d + j0 + 1] = a11;
  if (ti == tj && threadIdx.x < MT && ti * MT + threadIdx.x < d) out[1 + ti * MT + threadIdx.x] = sx;
  if (ti == 0 && tj == 0 && threadIdx.x == 0) out[0] = cnt;
}

__global__ void __launch_bounds__(256) k_moments_add(size_t stride, int nslab, const double* __restrict__ part,
                                                     double* __restrict__ acc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  double s = acc[i];
  for (int b = 0; b < nslab; ++b) s += part[(size_t)b * stride + i];
  acc[i] = s;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_moments_reset_dev(thy_model_t m, double* acc_dev) {
  thy::ErrScope _err_scope(m);
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!acc_dev) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  THY_TRY(require_device());
  const size_t stride = (size_t)1 + m->d + (size_t)m->d * m->d;
  THY_CUDA_CHECK(cudaMemsetAsync(acc_dev, 0, stride * sizeof(double), m->stream));
  return THY_OK;
}

thy_status thy_moments_accumulate_dev(thy_model_t m, int64_t n, const double* X_dev, double* acc_dev) {
  thy::ErrScope _err_scope(m);
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (n < 0) return fail(THY_ERR_INVALID_ARGUMENT, "negative count");
  if (n == 0) return THY_OK;
  if (!X_dev || !acc_dev) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  THY_TRY(require_device());
  const int d = m->d;
  const int tiles = ceil_div(d, MT);
  const size_t stride = (size_t)1 + d + (size_t)d * d;
  // enough slabs to fill the machine twice over, at least 64 rows each
  long long nslab = std::max<long long>(1, (2LL * sm_count() + (long long)tiles * tiles - 1) / ((long long)tiles * tiles));
  nslab = std::min<long long>(nslab, std::max<long long>(1, n / 64));
  nslab = std::min<long long>(nslab, 64);
  THY_TRY(m->ws.Z.reserve(stride * (size_t)nslab));
  cudaStream_t s = m->stream;
  k_moments_part<<<dim3(tiles, tiles, (unsigned)nslab), 256, 0, s>>>(n, d, X_dev, (int)nslab, m->ws.Z.ptr);
  k_moments_add<<<(unsigned)ceil_div((long long)stride, 256), 256, 0, s>>>(stride, (int)nslab, m->ws.Z.ptr, acc_dev);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status thy_moments_finish(thy_model_t m, thy_comm_t comm, double* acc_dev, double* out_count, double* out_mean,
                              double* out_cov) {
  thy::ErrScope _err_scope(m);
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!acc_dev) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  THY_TRY(require_device());
  const int d = m->d;
  const size_t stride = (size_t)1 + d + (size_t)d * d;
  cudaStream_t s = m->stream;
  THY_TRY(comm_all_reduce_f64(comm, acc_dev, stride, 0, s));
  std::vector<double> h(stride);
  THY_CUDA_CHECK(cudaMemcpyAsync(h.data(), acc_dev, stride * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  const double cnt = h[0];
  if (out_count) *out_count = cnt;
  if (cnt < 1.0) return fail(THY_ERR_STATE, "no samples accumulated");
  for (int i = 0; i < d; ++i) {
    const double mi = h[1 + i] / cnt;
    if (out_mean) out_mean[i] = mi;
    if (out_cov) {
      for (int j = 0; j < d; ++j) {
        const double mj = h[1 + j] / cnt;
        const double c = h[1 + d + (size_t)i * d + j] - cnt * mi * mj;
        out_cov[(size_t)i * d + j] = cnt > 1.0 ? c / (cnt - 1.0) : 0.0;
      }
    }
  }
  return THY_OK;
}

}  // extern "C"
