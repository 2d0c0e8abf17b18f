// fp64 roofline denominators for B200 (MEASURED_PEAKS.json has no fp64 entry):
//   dfma   : 8 independent FMA chains per thread, register-resident
//   dmma   : DMMA.8x8x4 (mma.sync m8n8k4 f64), 8 independent accumulator tiles per warp
//   dmma_lds: the same with one conflict-free LDS.64 per DMMA (the fused kernel's operand pattern)
// Prints one JSON line.  Build: make -C tools
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma_lds(double* out, int iters, double a) {
  __shared__ double sm[32 * 108];
  for (int i = threadIdx.x; i < 32 * 108; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int off = (lane >> 2) * 108 + (lane & 3);
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double b = sm[off + ((it + i) & 15) * 4 + (i & 3) * 8 * 108];
      dmma(c[i][0], c[i][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out;
  CK(cudaMalloc(&out, 64));
  const int iters = 4096;
  printf("{\"sms\": %d", sms);
  for (int bps = 1; bps <= 8; bps *= 2) {  // blocks per SM (256 threads each)
    const int grid = sms * bps;
    float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dfma_tflops_%dcta\": %.2f", bps, tf);
    ms = time_ms([&] { k_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    tf = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;  // 256 FMA per DMMA per warp, 8 warps
    printf(", \"dmma8_tflops_%dcta\": %.2f", bps, tf);
    ms = time_ms([&] { k_dmma<2><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    tf = 2.0 * 256 * 2 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma2_tflops_%dcta\": %.2f", bps, tf);
    ms = time_ms([&] { k_dmma_lds<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
    tf = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma_lds_tflops_%dcta\": %.2f", bps, tf);
  }
  printf("}\n");
  return 0;
}
