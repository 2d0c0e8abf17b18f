// DMMA.8x8x4 issue model on sm_100a: clocks per DMMA for one warp as a function of the number of independent accumulator
// chains it interleaves and of the number of warps that share a scheduler.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k(int iters, double a0, double b0, long long* clocks, double* sink) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c0[i] = c1[i] = 0.0;
  double a = a0 + threadIdx.x, b = b0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c0[i], c1[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) sink[0] = s;
  if ((threadIdx.x & 31) == 0) clocks[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int CH>
void run(int warps_per_sm, int sms) {
  const int iters = 4096 / CH * 4;
  long long* d;
  double* sink;
  cudaMalloc(&d, sizeof(long long) * sms * warps_per_sm);
  cudaMalloc(&sink, 8);
  k<CH><<<sms, warps_per_sm * 32>>>(iters, 1.0, 1e-9, d, sink);
  k<CH><<<sms, warps_per_sm * 32>>>(iters, 1.0, 1e-9, d, sink);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, d, sizeof(long long) * sms * warps_per_sm, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < sms * warps_per_sm; ++i) mx = h[i] > mx ? (double)h[i] : mx;
  const double per = mx / ((double)iters * CH);
  printf("{\"chains\": %d, \"warps_per_sm\": %d, \"warps_per_scheduler\": %.2f, \"clocks_per_dmma_per_warp\": %.2f, \"clocks_per_dmma_per_scheduler\": %.2f}\n",
         CH, warps_per_sm, warps_per_sm / 4.0, per, per / (warps_per_sm / 4.0));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  int sms = 148;
  for (int w : {4, 8, 12, 16, 24, 32}) {
    run<1>(w, sms);
    run<2>(w, sms);
    run<4>(w, sms);
    run<8>(w, sms);
  }
  return 0;
}
