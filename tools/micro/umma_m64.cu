// Where does tcgen05.mma (cta_group::1, kind::f16) put the accumulator rows when M = 64?
// A[r][0] = r + 1, A[r][1] = 1, B[n][0] = 1, B[n][1] = 128 (n + 1)  ->  D[r][n] = (r + 1) + 128 (n + 1), exactly.
// Every TMEM lane / column is dumped and decoded back to (r, n).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_m64 umma_m64.cu && ./umma_m64
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7))) << 4) + (k & 7) * 2;
}

__global__ void __launch_bounds__(128) test(float* out, int M, int N) {
  __shared__ __align__(1024) uint8_t sA[16384];
  __shared__ __align__(1024) uint8_t sB[16384];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384 / 2; i += 128) {
    reinterpret_cast<__nv_bfloat16*>(sA)[i] = __float2bfloat16(0.f);
    reinterpret_cast<__nv_bfloat16*>(sB)[i] = __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = tid; r < 128; r += 128) {
    *reinterpret_cast<__nv_bfloat16*>(sA + sw_off(r, 0)) = __float2bfloat16((float)(r + 1));
    *reinterpret_cast<__nv_bfloat16*>(sA + sw_off(r, 1)) = __float2bfloat16(1.f);
    *reinterpret_cast<__nv_bfloat16*>(sB + sw_off(r, 0)) = __float2bfloat16(1.f);
    *reinterpret_cast<__nv_bfloat16*>(sB + sw_off(r, 1)) = __float2bfloat16(128.f * (float)(r + 1));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  // sentinel: clear the accumulator columns with an M = 128 MMA over zero operands?  Simpler: store -1 with tcgen05.st
  {
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t m1 = __float_as_uint(-1.f);
    for (int c = 0; c < 64; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(m1) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = sw128_desc(smem_u32(sA)), db = sw128_desc(smem_u32(sB));
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
        "l"(da), "l"(db), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(&bar))
                 : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; ++c) {
      uint32_t v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out[tid * 64 + c] = __uint_as_float(v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  for (int M : {128, 64}) {
    const int N = 32;
    cudaMemset(d, 0, sizeof(h));
    test<<<1, 128>>>(d, M, N);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // per TMEM lane: which row r it holds (from column 0..N-1), or -1
    for (int lane = 0; lane < 128; ++lane) {
      int r_first = -2, consistent = 1, ncols = 0;
      for (int c = 0; c < 64; ++c) {
        const float v = h[lane * 64 + c];
        if (v == -1.f) continue;
        const int n = (int)(v / 128.f) - 1, r = (int)(v - 128.f * (n + 1)) - 1;
        ++ncols;
        if (n != c) consistent = 0;
        if (r_first == -2) r_first = r;
        else if (r != r_first) consistent = 0;
      }
      if (lane < 8 || lane % 16 == 0 || lane % 16 == 15)
        printf("  lane %3d: row %3d  cols written %2d  %s\n", lane, r_first, ncols, consistent ? "col c = n" : "MIXED");
    }
  }
  return 0;
}
