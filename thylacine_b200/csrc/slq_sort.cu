// Exact ascending sort of a round's retired log-likelihoods without a library call.
//
// The quadrature consumes the K values a round retires in ascending order (SlqEngine.scala:139-143 retires `keys.min`
// one at a time; QuadratureAbscissa.scala:33-73 pairs the i-th smallest with the i-th prior-mass shrink).  Round 1 sorted
// the whole pool with cub::DeviceRadixSort; round 2 replaced that by the radix SELECT of slq_select.cu and kept cub only
// for these K values (eight one-sweep passes of ~10 us each on the quadrature stream).  The values of one round lie in a
// narrow interval -- (previous threshold, threshold] -- and nearly uniformly in it, which a comparison-free radix sort
// cannot use and a bucket sort can:
//   k_bs_minmax   min / max of the order-preserving integer images (block reduction + one 64-bit atomic pair per CTA)
//   k_bs_count    bucket = floor((image - min) * NB / (max - min + 1))  (monotone in the value), histogram by atomics
//   k_bs_scan     exclusive scan of the NB counts (one CTA)
//   k_bs_scatter  value -> its bucket's segment (arrival order inside a segment is arbitrary)
//   k_bs_rank     one thread per element: rank inside its segment = #smaller + #equal-and-earlier, written to
//                 out[segment + rank]
//   k_bs_rearm    counters and min / max back to their rest state for the next round
// NB ~ n / 8, so a segment holds ~8 values and the rank step is ~8 comparisons per element; a degenerate input (many equal
// values -> one long segment) costs n_b comparisons per element of that segment, bounded by n per element.  The output is
// the sorted multiset, i.e. bit-identical to what cub produced.  All six kernels are plain launches (graph-capturable).
#include "slq_internal.h"

namespace thy {

namespace {

__device__ __forceinline__ unsigned long long ordered_image(double v) {   // monotone: a < b  <=>  image(a) < image(b)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ int bucket_of(unsigned long long img, unsigned long long lo, double scale, int nb) {
  const int b = (int)((double)(img - lo) * scale);
  return b < nb - 1 ? b : nb - 1;
}

__global__ void __launch_bounds__(1024) k_bs_minmax(const double* __restrict__ in, int n, unsigned long long* __restrict__ mm) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long v = ordered_image(in[i]);
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  __shared__ unsigned long long sl[32], sh[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sl[warp] = lo;
    sh[warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = sl[w] < lo ? sl[w] : lo;
      hi = sh[w] > hi ? sh[w] : hi;
    }
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

__global__ void __launch_bounds__(1024) k_bs_count(const double* __restrict__ in, int n, const unsigned long long* __restrict__ mm,
                                                  int nb, unsigned int* __restrict__ count) {
  const unsigned long long lo = mm[0];
  const double scale = (double)nb / ((double)(mm[1] - lo) + 1.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    atomicAdd(&count[bucket_of(ordered_image(in[i]), lo, scale, nb)], 1u);
}

// offset[b] = sum of count[0, b); offset[nb] = n.  One CTA of 1024 threads, nb a multiple of 1024.
__global__ void __launch_bounds__(1024) k_bs_scan(const unsigned int* __restrict__ count, int nb, unsigned int* __restrict__ offset) {
  __shared__ unsigned int wsum[32];
  const int per = nb / 1024, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  unsigned int local = 0;
  for (int k = 0; k < per; ++k) local += count[t * per + k];
  unsigned int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  unsigned int run = incl - local + (warp > 0 ? wsum[warp - 1] : 0u);
  for (int k = 0; k < per; ++k) {
    offset[t * per + k] = run;
    run += count[t * per + k];
  }
  if (t == 1023) offset[nb] = run;
}

__global__ void __launch_bounds__(1024) k_bs_scatter(const double* __restrict__ in, int n, const unsigned long long* __restrict__ mm,
                                                    int nb, const unsigned int* __restrict__ offset,
                                                    unsigned int* __restrict__ cursor, double* __restrict__ tmp) {
  const unsigned long long lo = mm[0];
  const double scale = (double)nb / ((double)(mm[1] - lo) + 1.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = in[i];
    const int b = bucket_of(ordered_image(v), lo, scale, nb);
    tmp[offset[b] + atomicAdd(&cursor[b], 1u)] = v;
  }
}

__global__ void __launch_bounds__(1024) k_bs_rank(const double* __restrict__ tmp, int n, const unsigned long long* __restrict__ mm,
                                                 int nb, const unsigned int* __restrict__ offset, double* __restrict__ out) {
  const unsigned long long lo = mm[0];
  const double scale = (double)nb / ((double)(mm[1] - lo) + 1.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = tmp[i];
    const unsigned long long img = ordered_image(v);
    const int b = bucket_of(img, lo, scale, nb);
    const unsigned int s0 = offset[b], s1 = offset[b + 1];
    unsigned int rank = 0;
    for (unsigned int j = s0; j < s1; ++j) {
      const unsigned long long o = ordered_image(tmp[j]);
      rank += (o < img || (o == img && j < (unsigned int)i)) ? 1u : 0u;
    }
    out[s0 + rank] = v;
  }
}

// counters and min / max back to their rest state for the next round (its own launch: stream order is the only barrier
// that covers every CTA of k_bs_rank)
__global__ void __launch_bounds__(256) k_bs_rearm(unsigned long long* __restrict__ mm, int nb, unsigned int* __restrict__ count,
                                                  unsigned int* __restrict__ cursor) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) {
    count[i] = 0u;
    cursor[i] = 0u;
  }
  if (i == 0) {
    mm[0] = ~0ull;
    mm[1] = 0ull;
  }
}

}  // namespace

static int slq_bucket_count(int n) {
  int nb = 1024;
  while (nb < n / 8 && nb < SLQ_SORT_MAX_BUCKETS) nb <<= 1;
  return nb;
}

thy_status slq_bucket_sort_init(const BucketSortBuffers& b, cudaStream_t s) {
  k_bs_rearm<<<SLQ_SORT_MAX_BUCKETS / 256, 256, 0, s>>>(b.minmax, SLQ_SORT_MAX_BUCKETS, b.count, b.cursor);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status slq_bucket_sort(const BucketSortBuffers& b, const double* in, double* out, int n, cudaStream_t s) {
  if (n <= 0) return THY_OK;
  const int nb = slq_bucket_count(n);
  // Very few, wide CTAs on purpose: this runs on the quadrature stream beside the replacement walk, whose grid is sized
  // to fill every SM exactly.  Every CTA resident here at the moment the walk's grid is placed costs one walk CTA its
  // slot, and the displaced CTA then queues behind the rest of this chain (measured at 10^6 live points: 1 150 us per
  // round with 148 CTAs per kernel, 1 105 with 32, against 1 045 with the library sort's handful).
  const unsigned grid = (unsigned)std::min(4, (n + 1023) / 1024);
  k_bs_minmax<<<grid, 1024, 0, s>>>(in, n, b.minmax);
  k_bs_count<<<grid, 1024, 0, s>>>(in, n, b.minmax, nb, b.count);
  k_bs_scan<<<1, 1024, 0, s>>>(b.count, nb, b.offset);
  k_bs_scatter<<<grid, 1024, 0, s>>>(in, n, b.minmax, nb, b.offset, b.cursor, b.tmp);
  k_bs_rank<<<2 * grid, 1024, 0, s>>>(b.tmp, n, b.minmax, nb, b.offset, out);
  k_bs_rearm<<<(unsigned)(nb / 256), 256, 0, s>>>(b.minmax, nb, b.count, b.cursor);
  count_launch(6);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
