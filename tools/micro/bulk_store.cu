// Per-SM rate of bulk asynchronous stores (cp.async.bulk.global.shared::cta) on B200, next to the st.global rate measured
// by sm_bandwidth.cu (62 GB/s per SM alone, 51 GB/s per SM with all 148 storing).  The question: is an epilogue that
// stages its final rows in shared memory better off handing them to the copy engine than storing them from threads?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/bulk_store tools/micro/bulk_store.cu
// Each CTA (one per SM) owns a 512 KB window (L2 resident) and writes it `iters` times from a 64 KB shared-memory tile:
//   mode 0: 128 copies of 512 B (one per thread: the row-per-copy form a padded staging layout needs)
//   mode 1: 4 copies of 16 KB
//   mode 2: 1 copy of 64 KB
//   mode 3: threads store the tile with 16-byte st.global (reference)
//   mode 4: bulk load of the tile (global -> shared, mbarrier) followed by the bulk store of the same tile (read-modify-write)
// `wait_group.read` before the tile is reused (the source may be overwritten), full wait_group at the end.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int TILE = 64 * 1024;

__global__ void __launch_bounds__(256) k_bulk(uint8_t* p, size_t per_cta, int iters, int mode, long long* cyc) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* q = p + (size_t)blockIdx.x * per_cta;
  for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 1, 2, 3);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int ntile = (int)(per_cta / TILE);
  long long t0 = clock64();
  uint32_t ph = 0;
  for (int it = 0; it < iters; ++it) {
    for (int t = 0; t < ntile; ++t) {
      uint8_t* g = q + (size_t)t * TILE;
      if (mode == 0) {
        if (threadIdx.x < 128) {
          bulk_store(g + threadIdx.x * 512, sm + threadIdx.x * 512, 512);
          bulk_commit();
          bulk_wait_read0();
        }
      } else if (mode == 1) {
        if (threadIdx.x < 4) {
          bulk_store(g + threadIdx.x * 16384, sm + threadIdx.x * 16384, 16384);
          bulk_commit();
          bulk_wait_read0();
        }
      } else if (mode == 2) {
        if (threadIdx.x == 0) {
          bulk_store(g, sm, TILE);
          bulk_commit();
          bulk_wait_read0();
        }
      } else if (mode == 3) {
        for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x)
          reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(sm)[i];
      } else {
        if (threadIdx.x == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(TILE) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)),
                       "l"(g), "r"(TILE), "r"(smem_u32(&bar))
                       : "memory");
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar)), "r"(ph)
                         : "memory");
          ph ^= 1;
          bulk_store(g, sm, TILE);
          bulk_commit();
          bulk_wait_read0();
        }
      }
      __syncthreads();
    }
  }
  long long t1 = clock64();   // sources consumed: the CTA could move on
  bulk_wait0();
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0;
    cyc[1] = t2 - t1;
  }
}

int main() {
  const size_t per_cta = 512u << 10;
  const int max_ctas = 148;
  uint8_t* buf;
  long long* cyc;
  cudaMalloc(&buf, per_cta * max_ctas);
  cudaMalloc(&cyc, 16);
  cudaMemset(buf, 0, per_cta * max_ctas);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 100;
  const char* names[5] = {"128 x 512 B bulk", "4 x 16 KB bulk", "1 x 64 KB bulk", "st.global v4", "bulk load + bulk store"};
  for (int ctas : {1, 80, 148})
    for (int mode = 0; mode < 5; ++mode) {
      float ms = 0.f;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_bulk<<<ctas, 256, TILE>>>(buf, per_cta, iters, mode, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h[2];
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      const double gb = (double)per_cta * iters / 1e9;
      printf("CTAs %3d %-24s: %6.1f GB/s per SM (%7.0f chip)%s  loop %lld cyc, drain %lld cyc  [%s]\n", ctas, names[mode], gb / (ms / 1e3),
             ctas * gb / (ms / 1e3), mode == 4 ? " each way" : "", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
