// Launch-chain floor: a CUDA graph of N dependent kernels that do (almost) nothing, launched with programmatic stream
// serialization like the step's kernels, with the same block size / dynamic shared memory footprint.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_floor launch_floor.cu && ./launch_floor
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

__global__ void empty_kernel(int* p, int work) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (work) {   // one dependent global round trip per kernel, like a real consumer
    int v = p[blockIdx.x];
    if (threadIdx.x == 0) p[blockIdx.x] = v + 1;
  }
}

static float run(int n, int grid, int block, size_t smem, bool pdl, int work) {
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  int* p;
  cudaMalloc(&p, 4096 * sizeof(int));
  cudaMemset(p, 0, 4096 * sizeof(int));
  cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, empty_kernel, p, work);
  }
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
  return ms * 1e3f / reps;
}

int main() {
  const int n = 203;
  struct { int grid, block; size_t smem; } shapes[] = {{60, 192, 205 * 1024}, {120, 192, 205 * 1024}, {80, 256, 0}, {64, 384, 20 * 1024}};
  for (auto& s : shapes)
    for (int pdl = 0; pdl < 2; ++pdl)
      for (int work = 0; work < 2; ++work) {
        float us = run(n, s.grid, s.block, s.smem, pdl, work);
        printf("n=%d grid=%3d block=%3d smem=%6zu pdl=%d work=%d: %8.1f us per graph, %5.2f us per kernel\n", n, s.grid, s.block,
               s.smem, pdl, work, us, us / n);
      }
  return 0;
}
