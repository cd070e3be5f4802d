// Row-owner GEMM on CTA pairs for the N = 384 projections of large batches (feed-forward down projection K = 1536,
// attention out projection and conv-module pointwise conv 2 K = 384; reference: conformer_blocks.py:468-482, :812-836):
//
//     x = r + scale * (A W^T + b) ;   r = x   or   r = g1 x / (rms(x) + eps)   ;   then either
//     rb = bf16(r), ss = sum r^2   (the next GEMM applies the following RMSNorm as a row scale), or
//     n = bf16(g2 r / (rms(r) + eps))   (optionally scattered into the [cache | new] rows of layers 14 / 15)
//
// At this size the GEMMs of the model are bound by the L2 -> SM operand ingest (~110 GB/s per SM), not by the tensor
// pipe: a 128 x 128 tile with K = 1536 pulls 786 KB for 50 MFLOP.  Here a 2-CTA cluster owns 256 whole rows
// (tcgen05.mma.cta_group::2, M = 256, N = 256 + 128 into TMEM columns [0, 384)): each CTA loads its own 128 rows of A
// and HALF of every weight K block (40 KB per K block and SM for 6.3 MFLOP: 2.5x fewer operand bytes per FLOP than
// the 128 x 128 tiling), and because a CTA owns whole rows the residual add, the row norms and the sum of squares are
// thread-local in the epilogue - no split-K partials, no norm kernel, one sum-of-squares tile per row.
#pragma once

#include "ff_fused.cuh"

namespace tone {

struct RowGemmArgs {
  int M;                    // valid rows
  int nk;                   // K blocks of 64
  const float* bias;        // [384]
  float scale;
  float* r;                 // [M][384] residual stream (in / out)
  const float* g1;          // nullable: norm applied to r in place
  const float* g2;          // nullable: gain of the RMSNorm that produces n
  bf16* n;                  // nullable: [M][384] bf16 rows out (g2-normalised, or plain bf16(r) when g2 is null)
  bf16* kv;                 // nullable: scatter n rows into [slots][KV_ROWS_MAX][384] at row kv_row_off + t
  const int* slots;
  int rows_per_stream, kv_row_off;
  bf16* rb_out;             // nullable: bf16(r) for a row-scale consumer
  float* ss_out;            // with rb_out: [M][ss_ld], column 0 = sum of squares of the row
  int ss_ld;
};

constexpr int RG_THREADS = 320;                           // warp 0: TMA, warp 1: MMA (leader CTA), warps 2..9: epilogue
constexpr int RG_A_BYTES = 128 * 128;                     // this CTA's 128 rows of one K block
constexpr int RG_W_BYTES = 192 * 128;                     // this CTA's half of the 384 weight rows of one K block
constexpr int RG_STAGE = RG_A_BYTES + RG_W_BYTES;         // 40 KB
constexpr int RG_STAGES = 5;
constexpr int RG_XP = 388;                                // floats per staged row of the epilogue (conflict-free)
constexpr int RG_OPER = RG_STAGES * RG_STAGE;             // 204,800
constexpr int RG_SMEM = RG_OPER + 3 * 384 * 4 + 256 + 1024;
static_assert(128 * RG_XP * 4 <= RG_OPER, "the x tile is staged over the dead operand ring");
static_assert(RG_SMEM <= 232448, "does not fit");

__global__ void __launch_bounds__(RG_THREADS, 1) rowgemm_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmW128,
                                                                      const __grid_constant__ CUtensorMap tmW64,
                                                                      const RowGemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_vec = reinterpret_cast<float*>(smem + RG_OPER);                 // b[384] | g1[384] | g2[384]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_vec + 3 * 384);           // [STAGES] (used on the leader)
  uint64_t* empty = full + RG_STAGES;                                      // [STAGES]
  uint64_t* acc_full = empty + RG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int tile = blockIdx.x;                     // 128-row tile of this CTA (CTAs 2p, 2p + 1 = rows 256p ..)

  PROF_DECL();
  PROF_BEGIN(9);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW128);
    tma_prefetch_desc(&tmW64);
    for (int s = 0; s < RG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  pair_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs): per K block this CTA's A rows and its half of the weight rows (128 rows of
    // the N = 256 MMA, 64 rows of the N = 128 MMA); all bytes complete on the leader's barrier
    bool waited = false;
    for (int kb = 0; kb < a.nk; ++kb) {
      const int s = kb % RG_STAGES;
      mbar_wait(&empty[s], ((kb / RG_STAGES) & 1) ^ 1);
      const uint32_t lbar = map_to_rank(smem_u32(&full[s]), 0);
      uint8_t* st = smem + s * RG_STAGE;
      if (elect_one_sync()) {
        if (leader) mbar_expect_tx(&full[s], 2 * RG_STAGE);
        tma_load_2d_pair(st + RG_A_BYTES, &tmW128, lbar, kb * 64, (int)rank * 128);
        tma_load_2d_pair(st + RG_A_BYTES + 128 * 128, &tmW64, lbar, kb * 64, 256 + (int)rank * 64);
      }
      __syncwarp();
      if (!waited) {           // activations come from the predecessor grid
        pdl_wait();
        waited = true;
        if (lane == 0) PROF_MARK(2);
      }
      if (elect_one_sync()) tma_load_2d_pair(st, &tmA, lbar, kb * 64, tile * 128);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA): per K block N = 256 into columns [0, 256) and N = 128 into [256, 384)
    if (leader) {
      constexpr uint32_t idesc256 = make_idesc_bf16_mn(256, 256);
      constexpr uint32_t idesc128 = make_idesc_bf16_mn(256, 128);
      const uint32_t s_u = smem_u32(smem);
      for (int kb = 0; kb < a.nk; ++kb) {
        const int s = kb % RG_STAGES;
        mbar_wait(&full[s], (kb / RG_STAGES) & 1);
        if (kb == 0 && lane == 0) PROF_MARK(3);
        tc_fence_after();
        const uint64_t da = make_sw128_desc(s_u + s * RG_STAGE);
        const uint64_t d0 = make_sw128_desc(s_u + s * RG_STAGE + RG_A_BYTES);
        const uint64_t d1 = make_sw128_desc(s_u + s * RG_STAGE + RG_A_BYTES + 128 * 128);
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_bf16_pair(tmem_base, da + 2 * ks, d0 + 2 * ks, idesc256, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_bf16_pair(tmem_base + 256, da + 2 * ks, d1 + 2 * ks, idesc128, (kb > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit_pair(acc_full);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue warps 2..9: warp w owns TMEM lanes 32 (w % 4) .. +31 and column half hf = (w - 2) / 4
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    for (int i = et; i < 384; i += EPI_THREADS) {      // constants (weights): before the dependency wait
      s_vec[i] = __ldg(a.bias + i);
      s_vec[384 + i] = a.g1 ? __ldg(a.g1 + i) : 1.f;
      s_vec[768 + i] = a.g2 ? __ldg(a.g2 + i) : 1.f;
    }
    pdl_wait();
    // residual rows of this warp's first batch travel under the main loop (they do not depend on it)
    const int ew = warp - 2;
    float4 x[8][3];
    auto load_rows = [&](int rg) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int grow = tile * 128 + ew * 16 + rg + k;
        if (grow < a.M) {
#pragma unroll
          for (int i = 0; i < 3; ++i) x[k][i] = *reinterpret_cast<const float4*>(a.r + (size_t)grow * D_MODEL + i * 128 + lane * 4);
        }
      }
    };
    load_rows(0);
    mbar_wait(acc_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    // Phase A (thread = row, columns [192 hf, +192)): scale * (acc + b) -> fp32 tile X[128][388] over the dead ring
    float* X = reinterpret_cast<float*>(smem);
    {
      const int c0 = hf * 192;
      const uint32_t xrow = smem_u32(X) + (row_in_tile * RG_XP + c0) * 4;
#pragma unroll 1
      for (int cb = 0; cb < 192; cb += 32) {
        uint32_t acc[32];
        tmem_ld16_async(tmem_base + lane_base + c0 + cb, acc);
        tmem_ld16_async(tmem_base + lane_base + c0 + cb + 16, acc + 16);
        tmem_ld_wait();
        tmem_regs_ready16(acc);
        tmem_regs_ready16(acc + 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(s_vec + c0 + cb + 4 * i);
          sts128(xrow + (cb + 4 * i) * 4,
                 make_float4(a.scale * (__uint_as_float(acc[4 * i]) + b.x), a.scale * (__uint_as_float(acc[4 * i + 1]) + b.y),
                             a.scale * (__uint_as_float(acc[4 * i + 2]) + b.z), a.scale * (__uint_as_float(acc[4 * i + 3]) + b.w)));
        }
      }
    }
    bar_epilogue();
    // Phase B (warp = 16 rows, lanes along the row, 8 rows in flight)
    Vec384 g1v, g2v;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g1v.v[i] = *reinterpret_cast<const float4*>(s_vec + 384 + i * 128 + lane * 4);
      g2v.v[i] = *reinterpret_cast<const float4*>(s_vec + 768 + i * 128 + lane * 4);
    }
    // The eight rows of a batch go through every step together (loads, sums, shuffles, scaling, stores): the per-row
    // chains (shared-memory load -> add -> sum of squares -> five shuffles -> rsqrt -> scale -> store) are independent and
    // only two warps share a scheduler, so walking them one row at a time leaves the SM waiting on latencies.
#pragma unroll 1
    for (int rg = 0; rg < 16; rg += 8) {
      if (rg) load_rows(rg);
      bool ok[8];
      float inv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rt = ew * 16 + rg + k;
        ok[k] = tile * 128 + rt < a.M;
        if (!ok[k]) {
#pragma unroll
          for (int i = 0; i < 3; ++i) x[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 d = lds128(smem_u32(X) + (rt * RG_XP + i * 128 + lane * 4) * 4);
          x[k][i].x += d.x;
          x[k][i].y += d.y;
          x[k][i].z += d.z;
          x[k][i].w += d.w;
        }
      }
      auto sumsq8 = [&](float (&out)[8]) {          // out[k] = sum of squares of row k (all lanes)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) sq += x[k][i].x * x[k][i].x + x[k][i].y * x[k][i].y + x[k][i].z * x[k][i].z + x[k][i].w * x[k][i].w;
          out[k] = sq;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) out[k] += __shfl_xor_sync(0xffffffffu, out[k], o);
        }
      };
      if (a.g1) {
        sumsq8(inv);
#pragma unroll
        for (int k = 0; k < 8; ++k) scale_384(x[k], g1v, 1.0f / (sqrtf(inv[k]) * 0.05103103630798288f + 1e-8f));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!ok[k]) continue;
        float* rr = a.r + (size_t)(tile * 128 + ew * 16 + rg + k) * D_MODEL;
#pragma unroll
        for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(rr + i * 128 + lane * 4) = x[k][i];
      }
      if (a.rb_out) {
        sumsq8(inv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!ok[k]) continue;
          const size_t grow = (size_t)(tile * 128 + ew * 16 + rg + k);
          bf16* rb = a.rb_out + grow * D_MODEL;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            *reinterpret_cast<uint2*>(rb + i * 128 + lane * 4) =
                make_uint2(pack_bf16x2(x[k][i].x, x[k][i].y), pack_bf16x2(x[k][i].z, x[k][i].w));
          if (lane == 0) a.ss_out[grow * a.ss_ld] = inv[k];
        }
      }
      if (a.n) {
        if (a.g2) {
          sumsq8(inv);
#pragma unroll
          for (int k = 0; k < 8; ++k) scale_384(x[k], g2v, 1.0f / (sqrtf(inv[k]) * 0.05103103630798288f + 1e-8f));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!ok[k]) continue;
          const int grow = tile * 128 + ew * 16 + rg + k;
          bf16* nr = a.n + (size_t)grow * D_MODEL;
          bf16* kr = nullptr;
          if (a.kv) {
            const int b = grow / a.rows_per_stream, t = grow - b * a.rows_per_stream;
            kr = a.kv + ((size_t)a.slots[b] * KV_ROWS_MAX + a.kv_row_off + t) * D_MODEL;
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const uint2 p = make_uint2(pack_bf16x2(x[k][i].x, x[k][i].y), pack_bf16x2(x[k][i].z, x[k][i].w));
            *reinterpret_cast<uint2*>(nr + i * 128 + lane * 4) = p;
            if (kr) *reinterpret_cast<uint2*>(kr + i * 128 + lane * 4) = p;
          }
        }
      }
    }
  }
  tc_fence_before();
  pair_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  PROF_END();
}

inline cudaError_t configure_rowgemm() {
  return cudaFuncSetAttribute(rowgemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RG_SMEM);
}

// m_tiles = 128-row tiles; the grid is rounded up to whole pairs (the odd tile's rows are invalid and skipped)
inline cudaError_t launch_rowgemm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW128, const CUtensorMap& tmW64,
                                  const RowGemmArgs& a, int m_tiles, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * ((m_tiles + 1) / 2));
  cfg.blockDim = dim3(RG_THREADS);
  cfg.dynamicSmemBytes = RG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = 2;
  at[na].val.clusterDim.y = 1;
  at[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, rowgemm_pair_kernel, tmA, tmW128, tmW64, a);
}

}  // namespace tone
