// CartesianSurface: Gaussian-kernel accumulation of sample curves onto an x-y grid.
// Reference (paths relative to /root/reference/src/main/scala/thylacine/):
//   visualisation/CartesianSurface.scala:110-150  addSamples / addValues / concatenateMapping:
//        value(g) += sum over interpolation points p of exp(-0.5 * scaledDistSquared(p, g) / kernelVariance)
//   visualisation/CartesianSurface.scala:53-83    normaliseColumns: each x column divided by its trapezoid integral over y
//   visualisation/SimplexChain.scala:21-33        linearInterpolationUsing(ds): arc-length resampling, remainder carried
//   visualisation/Simplex.scala:25-47             getInterpolation / getInterpolationPoints (start += ds by repeated addition)
//   visualisation/GraphPoint.scala:26-31          scaledDistSquaredTo
//   util/MathOps.scala:26-55                      trapezoidalQuadrature (sorted by abscissa)
// Quirk kept: the resampled points are wrapped in a SimplexChain whose getPoints returns every simplex START, i.e. the
// LAST element of the point vector is dropped; segments prepend their points, so that element is the FIRST interpolation
// point of the last segment that produced any.  A curve that resamples to fewer than 2 points contributes nothing.
//
// Device layout: curves [S][m] row-major (a consumer of the samplers' output buffers), points as double2, grid values
// [nx][ny].  Resampling is one thread per curve (sequential in the segments, as the remainder carries); accumulation
// is grid-point tiles x point chunks with per-chunk partial sums reduced in a fixed order (deterministic, no atomics).
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <numeric>

struct thy_surface_s {
  int nx = 0, ny = 0;
  std::vector<double> hx, hy;
  std::vector<int> y_order;      // indices of yAbscissa in ascending order
  bool y_distinct = true;
  double xscale = 0.0, yscale = 0.0;
  thy::DevBuf<double> gx, gy, values, partial, curves, absc;
  thy::DevBuf<double2> pts;
  thy::DevBuf<long long> counts;
  cudaStream_t stream = nullptr;
};

namespace thy {

// Walks one curve; emit(p) is called for every point the reference keeps, in segment order.
template <typename Emit>
__device__ __forceinline__ long long surface_walk(int m, const double* __restrict__ ab, const double* __restrict__ v, double ds,
                                                  int skip_segment, Emit emit) {
  long long total = 0;
  double start = 0.0;
  for (int k = 0; k + 1 < m; ++k) {
    const double x0 = ab[k], y0 = v[k], x1 = ab[k + 1], y1 = v[k + 1];
    const double dx = x1 - x0, dy = y1 - y0;
    // Math.sqrt(Math.pow(.,2) + Math.pow(.,2)) and start + differential * t: the JVM does not contract to FMA
    const double len = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    bool first = true;
    while (!(start >= len)) {
      const double t = start / len;
      double px, py;
      if (t <= 0) { px = x0; py = y0; }
      else if (t >= 1) { px = x1; py = y1; }
      else { px = __dadd_rn(x0, __dmul_rn(t, dx)); py = __dadd_rn(y0, __dmul_rn(t, dy)); }
      if (!(first && k == skip_segment)) emit(make_double2(px, py));
      first = false;
      ++total;
      start = start + ds;
    }
    start = start - len;
  }
  return total;
}

// pass 1: number of resampled points per curve and the last segment that produced any
__global__ void __launch_bounds__(128) k_surface_count(long long S, int m, const double* __restrict__ ab,
                                                       const double* __restrict__ curves, double ds,
                                                       long long* __restrict__ counts, int* __restrict__ last_seg) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S) return;
  const double* v = curves + c * m;
  long long total = 0;
  int last = -1;
  double start = 0.0;
  for (int k = 0; k + 1 < m; ++k) {
    const double dx = ab[k + 1] - ab[k], dy = v[k + 1] - v[k];
    const double len = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    long long n = 0;
    while (!(start >= len)) {
      ++n;
      start = start + ds;
    }
    start = start - len;
    if (n > 0) last = k;
    total += n;
  }
  counts[c] = total >= 2 ? total - 1 : 0;   // fewer than 2 points: empty chain; otherwise one point is dropped
  last_seg[c] = total >= 2 ? last : -2;
}

// pass 2: write the kept points of every curve at its offset
__global__ void __launch_bounds__(128) k_surface_points(long long S, int m, const double* __restrict__ ab,
                                                        const double* __restrict__ curves, double ds,
                                                        const long long* __restrict__ offsets, const int* __restrict__ last_seg,
                                                        double2* __restrict__ pts) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S) return;
  const int skip = last_seg[c];
  if (skip == -2) return;
  double2* out = pts + offsets[c];
  long long w = 0;
  surface_walk(m, ab, curves + c * m, ds, skip, [&](double2 p) { out[w++] = p; });
}

constexpr int SURF_T = 256;      // grid points per CTA
constexpr int SURF_STAGE = 1024; // points staged in shared memory per pass

__global__ void __launch_bounds__(SURF_T) k_surface_accumulate(int nx, int ny, const double* __restrict__ gx,
                                                               const double* __restrict__ gy, double inv_xs, double inv_ys,
                                                               double neg_half_inv_var, const double2* __restrict__ pts,
                                                               long long npts, long long chunk, double* __restrict__ partial) {
  __shared__ double2 sp[SURF_STAGE];
  const long long g = (long long)blockIdx.x * SURF_T + threadIdx.x;
  const long long ng = (long long)nx * ny;
  const bool live = g < ng;
  const double x = live ? gx[g / ny] : 0.0, y = live ? gy[g % ny] : 0.0;
  const long long p0 = (long long)blockIdx.y * chunk;
  const long long p1 = (p0 + chunk < npts) ? p0 + chunk : npts;
  double acc = 0.0;
  for (long long base = p0; base < p1; base += SURF_STAGE) {
    const int n = (int)((p1 - base < SURF_STAGE) ? (p1 - base) : SURF_STAGE);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += SURF_T) sp[i] = pts[base + i];
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const double dx = (sp[i].x - x) * inv_xs, dy = (sp[i].y - y) * inv_ys;
        acc += exp(neg_half_inv_var * (dx * dx + dy * dy));
      }
    }
  }
  if (live) partial[(size_t)blockIdx.y * ng + g] = acc;
}

__global__ void __launch_bounds__(256) k_surface_reduce(long long ng, int nchunks, const double* __restrict__ partial,
                                                        double* __restrict__ values) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  double a = values[g];
  for (int c = 0; c < nchunks; ++c) a += partial[(size_t)c * ng + g];   // fixed order
  values[g] = a;
}

// one thread per x column: trapezoid over y in ascending order, divide when positive
__global__ void __launch_bounds__(128) k_surface_normalise(int nx, int ny, const double* __restrict__ gy,
                                                           const int* __restrict__ order, const double* __restrict__ values,
                                                           double* __restrict__ out) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= nx) return;
  const double* col = values + (size_t)ix * ny;
  double integ = 0.0;
  for (int j = 0; j + 1 < ny; ++j) {
    const int a = order[j], b = order[j + 1];
    integ += 0.5 * (col[a] + col[b]) * (gy[b] - gy[a]);
  }
  for (int j = 0; j < ny; ++j) out[(size_t)ix * ny + j] = integ > 0 ? col[j] / integ : col[j];
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_surface_create(int nx, const double* x_abscissa, int ny, const double* y_abscissa, thy_surface_t* out) {
  if (!out) return fail(THY_ERR_INVALID_ARGUMENT, "null out pointer");
  *out = nullptr;
  THY_TRY(require_device());
  if (nx < 1 || ny < 1 || !x_abscissa || !y_abscissa) return fail(THY_ERR_INVALID_ARGUMENT, "empty abscissa");
  auto* s = new (std::nothrow) thy_surface_s();
  if (!s) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  s->nx = nx;
  s->ny = ny;
  s->hx.assign(x_abscissa, x_abscissa + nx);
  s->hy.assign(y_abscissa, y_abscissa + ny);
  s->xscale = *std::max_element(s->hx.begin(), s->hx.end()) - *std::min_element(s->hx.begin(), s->hx.end());
  s->yscale = *std::max_element(s->hy.begin(), s->hy.end()) - *std::min_element(s->hy.begin(), s->hy.end());
  s->y_order.resize(ny);
  std::iota(s->y_order.begin(), s->y_order.end(), 0);
  std::stable_sort(s->y_order.begin(), s->y_order.end(), [&](int a, int b) { return s->hy[a] < s->hy[b]; });
  for (int j = 0; j + 1 < ny; ++j)
    if (s->hy[s->y_order[j]] == s->hy[s->y_order[j + 1]]) s->y_distinct = false;
  thy_status st = THY_OK;
  auto chk = [&](thy_status s2) { if (st == THY_OK) st = s2; };
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamDefault) != cudaSuccess) st = fail(THY_ERR_CUDA, "stream creation failed");
  chk(s->gx.upload(s->hx, s->stream));
  chk(s->gy.upload(s->hy, s->stream));
  chk(s->values.reserve((size_t)nx * ny));
  if (st == THY_OK && cudaMemsetAsync(s->values.ptr, 0, (size_t)nx * ny * sizeof(double), s->stream) != cudaSuccess)
    st = fail(THY_ERR_CUDA, "surface initialisation failed");   // zeroValuesSpec (CartesianSurface.scala:85-88)
  if (st == THY_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) st = fail(THY_ERR_CUDA, "surface initialisation failed");
  if (st != THY_OK) {
    thy_surface_destroy(s);
    return st;
  }
  *out = s;
  return THY_OK;
}

thy_status thy_surface_destroy(thy_surface_t s) {
  if (s && s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return THY_OK;
}

static thy_status surface_add(thy_surface_t s, int m, const double* abscissa_host, int64_t S, const double* curves_dev, double ds,
                              double kernel_variance) {
  if (S == 0 || m < 2) return THY_OK;   // SimplexChain of fewer than two points is empty (SimplexChain.scala:37-43)
  if (!(ds > 0)) return fail(THY_ERR_INVALID_ARGUMENT, "ds must be > 0 (the reference would not terminate)");
  if (!(kernel_variance > 0)) return fail(THY_ERR_INVALID_ARGUMENT, "kernelVariance must be > 0");
  if (!(s->xscale > 0) || !(s->yscale > 0))
    return fail(THY_ERR_INVALID_ARGUMENT, "degenerate abscissa: max - min must be > 0 on both axes");
  cudaStream_t st = s->stream;
  std::vector<double> ab(abscissa_host, abscissa_host + m);
  THY_TRY(s->absc.upload(ab, st));
  THY_TRY(s->counts.reserve((size_t)2 * S + 2));
  DevBuf<int> last_seg;
  THY_TRY(last_seg.reserve((size_t)S));
  long long* counts = s->counts.ptr;
  long long* offsets = s->counts.ptr + S;
  k_surface_count<<<(unsigned)ceil_div(S, 128), 128, 0, st>>>(S, m, s->absc.ptr, curves_dev, ds, counts, last_seg.ptr);
  count_launch();
  // exclusive scan of the per-curve counts on the host side of a single D2H/H2D (S is the number of curves, small next
  // to the S x points x grid accumulation that follows)
  std::vector<long long> hc((size_t)S), ho((size_t)S);
  THY_CUDA_CHECK(cudaMemcpyAsync(hc.data(), counts, (size_t)S * sizeof(long long), cudaMemcpyDeviceToHost, st));
  THY_CUDA_CHECK(cudaStreamSynchronize(st));
  long long npts = 0;
  for (int64_t c = 0; c < S; ++c) {
    ho[c] = npts;
    npts += hc[c];
  }
  if (npts == 0) return THY_OK;
  THY_CUDA_CHECK(cudaMemcpyAsync(offsets, ho.data(), (size_t)S * sizeof(long long), cudaMemcpyHostToDevice, st));
  THY_TRY(s->pts.reserve((size_t)npts));
  k_surface_points<<<(unsigned)ceil_div(S, 128), 128, 0, st>>>(S, m, s->absc.ptr, curves_dev, ds, offsets, last_seg.ptr, s->pts.ptr);
  count_launch();
  const long long ng = (long long)s->nx * s->ny;
  const long long tiles = ceil_div(ng, SURF_T);
  long long nchunks = ceil_div((long long)sm_count() * 8, tiles);
  const long long max_chunks = ceil_div(npts, SURF_STAGE);
  if (nchunks > max_chunks) nchunks = max_chunks;
  if (nchunks > 65535) nchunks = 65535;
  if (nchunks < 1) nchunks = 1;
  const long long chunk = round_up(ceil_div(npts, nchunks), SURF_STAGE);
  nchunks = ceil_div(npts, chunk);
  THY_TRY(s->partial.reserve((size_t)nchunks * ng));
  dim3 grid((unsigned)tiles, (unsigned)nchunks);
  k_surface_accumulate<<<grid, SURF_T, 0, st>>>(s->nx, s->ny, s->gx.ptr, s->gy.ptr, 1.0 / s->xscale, 1.0 / s->yscale,
                                                -0.5 / kernel_variance, s->pts.ptr, npts, chunk, s->partial.ptr);
  k_surface_reduce<<<(unsigned)ceil_div(ng, 256), 256, 0, st>>>(ng, (int)nchunks, s->partial.ptr, s->values.ptr);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  THY_CUDA_CHECK(cudaStreamSynchronize(st));   // last_seg dies with this scope
  return THY_OK;
}

thy_status thy_surface_add_samples_dev(thy_surface_t s, int m, const double* abscissa, int64_t n_samples, const double* samples_dev,
                                       double ds, double kernel_variance) {
  if (!s) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (m < 0 || n_samples < 0 || (m > 0 && !abscissa) || (n_samples > 0 && m > 0 && !samples_dev))
    return fail(THY_ERR_INVALID_ARGUMENT, "bad sample block");
  THY_CUDA_CHECK(cudaDeviceSynchronize());   // the caller's buffer may have been produced on any stream
  return surface_add(s, m, abscissa, n_samples, samples_dev, ds, kernel_variance);
}

thy_status thy_surface_add_samples(thy_surface_t s, int m, const double* abscissa, int64_t n_samples, const double* samples,
                                   double ds, double kernel_variance) {
  if (!s) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (m < 0 || n_samples < 0 || (m > 0 && !abscissa) || (n_samples > 0 && m > 0 && !samples))
    return fail(THY_ERR_INVALID_ARGUMENT, "bad sample block");
  if (n_samples == 0 || m < 2) return THY_OK;
  THY_TRY(s->curves.reserve((size_t)n_samples * m));
  THY_CUDA_CHECK(cudaMemcpyAsync(s->curves.ptr, samples, (size_t)n_samples * m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  return surface_add(s, m, abscissa, n_samples, s->curves.ptr, ds, kernel_variance);
}

thy_status thy_surface_results(thy_surface_t s, int normalise_columns, double* out_values) {
  if (!s || !out_values) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  const size_t ng = (size_t)s->nx * s->ny;
  const double* src = s->values.ptr;
  DevBuf<double> tmp;
  DevBuf<int> order;
  if (normalise_columns) {
    // trapezoidalQuadrature throws on fewer than 2 points or repeated abscissa values (MathOps.scala:39-55)
    if (s->ny < 2 || !s->y_distinct)
      return fail(THY_ERR_INVALID_ARGUMENT, "Malformed abscissa for trapezoidal quadrature (MathOps.scala:53-55)");
    THY_TRY(tmp.reserve(ng));
    THY_TRY(order.upload(s->y_order, s->stream));
    k_surface_normalise<<<(unsigned)ceil_div(s->nx, 128), 128, 0, s->stream>>>(s->nx, s->ny, s->gy.ptr, order.ptr, s->values.ptr,
                                                                             tmp.ptr);
    count_launch();
    THY_CUDA_CHECK(cudaGetLastError());
    src = tmp.ptr;
  }
  THY_CUDA_CHECK(cudaMemcpyAsync(out_values, src, ng * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  THY_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  return THY_OK;
}

}  // extern "C"
