// Is fma.rn.f32x2 (FFMA2) issued at the rate of FFMA on B200, i.e. does it double the fp32 FMA throughput per issue slot?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/ffma2 tools/micro/ffma2.cu
// 148 CTAs x 512 threads, 16 independent accumulator chains per thread, 4096 iterations; prints FMA lanes per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) k_ffma(float* out, float a, float b, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) k_ffma2(float* out, float a, float b, int iters) {
  unsigned long long acc[16];
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = threadIdx.x * 0.001f + i, y = x + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(x), "f"(y));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
    s += x + y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int mode = 0; mode < 2; ++mode) {
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_ffma<<<148, 512>>>(out, 0.999f, 0.001f, iters);
      else k_ffma2<<<148, 512>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double instr = 512.0 * 16 * iters;                      // per SM (thread-instructions)
    const double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%s: %.3f ms, %.1f thread-instructions per clock per SM = %.1f fp32 FMA lanes per clock per SM  [%s]\n",
           mode ? "FFMA2" : "FFMA ", ms, instr / clocks, instr / clocks * (mode ? 2 : 1), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
