// How long does a CTA need to pull one 128 x 384 bf16 activation tile (6 TMA boxes of 128 rows x 64 columns, 96 KB)
// into shared memory, as a function of how many CTAs read the SAME tile and of the global layout?
//   shared   : grid = 5 M tiles x NT CTAs, all NT CTAs of a row read the same tile (what the N-tile CTAs of a GEMM do)
//   distinct : every CTA reads its own tile (no sharing)
//   panel    : same as shared, but the matrix is stored k-block-major ([6][rows][64]), so a box is 16 KB contiguous
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_ingest tma_ingest.cu -lcuda && ./tma_ingest
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64) ingest(const __grid_constant__ CUtensorMap tm, int mode, int rows_per_tile,
                                             int nt, long long* out, int nbox = 6) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[6];
  const int m = blockIdx.x / nt, n = blockIdx.x % nt;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = 0, t1 = 0, tfirst = 0;
  if (threadIdx.x < nbox) {
    const int kb = threadIdx.x;
    // mode 0 (shared): tile m; mode 1 (distinct): tile blockIdx.x; mode 2 (panel): rows of panel kb
    int x = kb * 64, y = (mode == 1 ? blockIdx.x : m) * rows_per_tile;
    if (mode == 2) {
      x = 0;
      y = kb * (gridDim.x / nt) * rows_per_tile + m * rows_per_tile;
    }
    const unsigned lm = (1u << nbox) - 1u;
    __syncwarp(lm);
    t0 = clock64();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[kb])), "r"(16384) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(sm + kb * 16384)),
        "l"((uint64_t)&tm), "r"(smem_u32(&bar[kb])), "r"(x), "r"(y)
        : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar[kb]))
                   : "memory");
    t1 = clock64();
    // first / last box of this CTA
    long long d = t1 - t0;
    long long dmax = d, dmin = d;
    for (int o = 4; o > 0; o >>= 1) {
      if ((int)(threadIdx.x ^ o) >= nbox && nbox != 6) continue;   // partner not in the mask (nbox = 1, 2, 3: handled below)
      long long a = __shfl_xor_sync(lm, dmax, o), b = __shfl_xor_sync(lm, dmin, o);
      if ((int)(threadIdx.x ^ o) < nbox) {
        dmax = a > dmax ? a : dmax;
        dmin = b < dmin ? b : dmin;
      }
    }
    if (threadIdx.x == 0) {
      out[2 * blockIdx.x] = dmin;
      out[2 * blockIdx.x + 1] = dmax;
    }
    (void)tfirst;
  }
  (void)n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const int TILE = 128, K = 384, MAXT = 160;
  __nv_bfloat16* A;
  cudaMalloc(&A, (size_t)MAXT * TILE * K * 2);
  cudaMemset(A, 0, (size_t)MAXT * TILE * K * 2);
  long long* out;
  cudaMalloc(&out, 2 * 160 * sizeof(long long));
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  auto make = [&](CUtensorMap* m, uint64_t cols, uint64_t rows) {
    cuuint64_t gd[2] = {cols, rows}, gs[1] = {cols * 2};
    cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  };
  CUtensorMap rowmajor, panel;
  make(&rowmajor, K, (uint64_t)MAXT * TILE);
  make(&panel, 64, (uint64_t)MAXT * TILE * 6);
  struct Case { const char* name; int mode, mt, nt; } cases[] = {
      {"shared   5 tiles x 1", 0, 5, 1},   {"shared   5 tiles x 6", 0, 5, 6},   {"shared   5 tiles x 12", 0, 5, 12},
      {"shared   5 tiles x 24", 0, 5, 24}, {"distinct 60 tiles", 1, 60, 1},     {"distinct 120 tiles", 1, 120, 1},
      {"panel    5 tiles x 12", 2, 5, 12}, {"panel    5 tiles x 24", 2, 5, 24}, {"shared   1 tile x 1", 0, 1, 1}};
  for (auto& c : cases) {
    const int grid = c.mt * c.nt;
    std::vector<long long> h(2 * grid);
    std::vector<double> mins, maxs;
    for (int rep = 0; rep < 6; ++rep) {
      // touch the matrix so that it sits in L2 like a freshly written activation
      cudaMemset(A, rep, (size_t)(c.mode == 1 ? grid : c.mt) * TILE * K * 2 * (c.mode == 2 ? 1 : 1));
      if (c.mode == 2) cudaMemset(A, rep, (size_t)MAXT * TILE * 64 * 2 * 6 > 0 ? (size_t)c.mt * TILE * K * 2 * 32 : 0);
      ingest<<<grid, 64, 100 * 1024>>>(c.mode == 2 ? panel : rowmajor, c.mode, TILE, c.nt, out);
      cudaDeviceSynchronize();
      cudaMemcpy(h.data(), out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
      if (rep == 0) continue;
      for (int i = 0; i < grid; ++i) {
        mins.push_back(h[2 * i] / 1965.0);
        maxs.push_back(h[2 * i + 1] / 1965.0);
      }
    }
    std::sort(mins.begin(), mins.end());
    std::sort(maxs.begin(), maxs.end());
    printf("%-24s CTAs %3d: first box median %.2f us | last box median %.2f us, p95 %.2f us, max %.2f us\n", c.name, grid,
           mins[mins.size() / 2], maxs[maxs.size() / 2], maxs[maxs.size() * 95 / 100], maxs.back());
  }
  // bytes vs time: 1, 2, 3 and 6 boxes of 16 KB per CTA, 60 CTAs (5 tiles x 12)
  for (int nbox : {1, 2, 6}) {
    const int grid = 60;
    std::vector<long long> h(2 * grid);
    std::vector<double> maxs;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemset(A, rep, (size_t)5 * TILE * K * 2);
      ingest<<<grid, 64, 100 * 1024>>>(rowmajor, 0, TILE, 12, out, nbox);
      cudaDeviceSynchronize();
      cudaMemcpy(h.data(), out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
      if (rep == 0) continue;
      for (int i = 0; i < grid; ++i) maxs.push_back(h[2 * i + 1] / 1965.0);
    }
    std::sort(maxs.begin(), maxs.end());
    printf("boxes per CTA %d (%3d KB): last box median %.2f us, p95 %.2f us\n", nbox, nbox * 16, maxs[maxs.size() / 2],
           maxs[maxs.size() * 95 / 100]);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
