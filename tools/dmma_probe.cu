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


// ---- operands from shared memory, as the product kernels have them: GEMM1 of one row block group ----
// JG accumulator chains (row groups), KS K-slices; b = As[(j*8 + g)*LDS + 4*kk + q], a = xf[kk].
// UNROLLED: the form kernels_fused_dmma.cu uses (kk outer, j inner, fully unrolled, a from registers).
// ROLLED:   K loop kept rolled (a from shared memory), JG independent DMMAs per iteration.
constexpr int KS = 26, LDSP = 108;
template <int JG, bool ROLLED>
__global__ void ks(int iters, const double* __restrict__ Ag, long long* clocks, double* sink) {
  extern __shared__ double sm[];
  double* As = sm;                       // [JG*8][LDSP]
  double* xs = sm + JG * 8 * LDSP;       // [warps][KS][32]
  for (int i = threadIdx.x; i < JG * 8 * LDSP; i += blockDim.x) As[i] = Ag[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double xf[KS];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) { xf[kk] = 1.0 + 1e-3 * (kk + lane); xs[(warp * KS + kk) * 32 + lane] = xf[kk]; }
  __syncthreads();
  const int off1 = g * LDSP + q;
  double acc = 0.0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    double z[JG][2];
#pragma unroll
    for (int j = 0; j < JG; ++j) z[j][0] = z[j][1] = 0.0;
    if (ROLLED) {
      const double* bp = As + off1;
      const double* xq = xs + (warp * KS) * 32 + lane;
#pragma unroll 1
      for (int kk = 0; kk < KS; ++kk, bp += 4, xq += 32) {
        const double a = *xq;
        double b[JG];
#pragma unroll
        for (int j = 0; j < JG; ++j) b[j] = bp[j * 8 * LDSP];
#pragma unroll
        for (int j = 0; j < JG; ++j) dmma(z[j][0], z[j][1], a, b[j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
        for (int j = 0; j < JG; ++j) dmma(z[j][0], z[j][1], xf[kk], As[off1 + j * 8 * LDSP + 4 * kk]);
      }
    }
#pragma unroll
    for (int j = 0; j < JG; ++j) acc += z[j][0] + z[j][1];
  }
  const long long t1 = clock64();
  if (acc == 12345.678) sink[0] = acc;
  if (lane == 0) clocks[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
}

template <int JG, bool ROLLED>
void runs(int warps_per_sm, int sms) {
  const int iters = 400;
  long long* d;
  double *sink, *Ag;
  cudaMalloc(&d, sizeof(long long) * sms * warps_per_sm);
  cudaMalloc(&sink, 8);
  cudaMalloc(&Ag, sizeof(double) * JG * 8 * LDSP);
  cudaMemset(Ag, 0, sizeof(double) * JG * 8 * LDSP);
  const size_t smem = sizeof(double) * (JG * 8 * LDSP + warps_per_sm * KS * 32);
  cudaFuncSetAttribute(ks<JG, ROLLED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) ks<JG, ROLLED><<<sms, warps_per_sm * 32, smem>>>(iters, Ag, d, sink);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, d, sizeof(long long) * sms * warps_per_sm, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < sms * warps_per_sm; ++i) mx = h[i] > mx ? (double)h[i] : mx;
  const double per = mx / ((double)iters * JG * KS);
  printf("{\"operands\": \"shared\", \"k_loop\": \"%s\", \"chains\": %d, \"warps_per_sm\": %d, \"clocks_per_dmma_per_warp\": %.2f, \"clocks_per_dmma_per_scheduler\": %.2f, \"err\": \"%s\"}\n",
         ROLLED ? "rolled" : "unrolled", JG, warps_per_sm, per, per / (warps_per_sm / 4.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(sink); cudaFree(Ag);
}

int main() {
  int sms = 148;
  for (int w : {4, 8, 12, 16, 24, 32}) {
    run<1>(w, sms);
    run<2>(w, sms);
    run<4>(w, sms);
    run<8>(w, sms);
  }
  for (int w : {4, 8, 12}) {
    runs<1, false>(w, sms); runs<2, false>(w, sms); runs<4, false>(w, sms);
    runs<1, true>(w, sms); runs<2, true>(w, sms); runs<4, true>(w, sms); runs<8, true>(w, sms);
  }
  return 0;
}
