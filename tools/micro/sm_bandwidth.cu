// Per-SM global-memory rates on B200: what one SM can pull from / push into L2, as a function of how many SMs do it at once.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/sm_bandwidth tools/micro/sm_bandwidth.cu
// Each CTA (256 or 1024 threads, one per SM) streams over its own 512 KB window (L2 resident after the first pass) with 16-byte
// accesses, 8 in flight per thread.  Prints GB/s per SM and for the chip: loads, stores, and read-modify-write.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_load(const float4* p, size_t per_cta, int iters, float* sink) {
  const float4* q = p + (size_t)blockIdx.x * per_cta;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it)
    for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (i + k * blockDim.x < per_cta) ? q[i + k * blockDim.x] : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
    }
  if (acc == 1.2345f) sink[0] = acc;
}
__global__ void __launch_bounds__(1024) k_store(float4* p, size_t per_cta, int iters) {
  float4* q = p + (size_t)blockIdx.x * per_cta;
  for (int it = 0; it < iters; ++it)
    for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) q[i] = make_float4(it, 1.f, 2.f, 3.f);
}
__global__ void __launch_bounds__(1024) k_rmw(float4* p, size_t per_cta, int iters) {
  float4* q = p + (size_t)blockIdx.x * per_cta;
  for (int it = 0; it < iters; ++it)
    for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (i + k * blockDim.x < per_cta) ? q[i + k * blockDim.x] : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (i + k * blockDim.x < per_cta) q[i + k * blockDim.x] = make_float4(v[k].x + 1.f, v[k].y, v[k].z, v[k].w);
    }
}

int main() {
  const size_t per_cta = (512u << 10) / 16;   // float4 elements per CTA window (148 x 512 KB stays L2 resident)
  const int max_ctas = 296;
  float4* buf;
  float* sink;
  cudaMalloc(&buf, per_cta * 16 * max_ctas);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, per_cta * 16 * max_ctas);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 200;
  for (int nt : {256, 1024})
  for (int ctas : {1, 40, 80, 148}) {
    float ms[3];
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {   // first repetition warms L2
        cudaEventRecord(e0);
        if (mode == 0) k_load<<<ctas, nt>>>(buf, per_cta, iters, sink);
        else if (mode == 1) k_store<<<ctas, nt>>>(buf, per_cta, iters);
        else k_rmw<<<ctas, nt>>>(buf, per_cta, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[mode], e0, e1);
      }
    }
    const double gb = (double)per_cta * 16 * iters / 1e9;
    printf("threads %4d CTAs %3d: load %6.1f GB/s per CTA (%7.0f chip) | store %6.1f (%7.0f) | rmw %6.1f + %6.1f (%7.0f)\n", nt, ctas, gb / (ms[0] / 1e3), ctas * gb / (ms[0] / 1e3), gb / (ms[1] / 1e3),
           ctas * gb / (ms[1] / 1e3), gb / (ms[2] / 1e3), gb / (ms[2] / 1e3), 2 * ctas * gb / (ms[2] / 1e3));
  }
  return 0;
}
