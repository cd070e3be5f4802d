// Feed-forward down projection + residual + RMSNorm(s) in ONE kernel, split-K across a thread-block cluster:
//
//     r += 0.5 * (h W2^T + b2) ;  [r = g1 * r / (rms(r) + eps)] ;  n = g2 * r / (rms(r) + eps)      (K = 1536, N = 384)
//
// reference: ConformerFeedForward.forward linear2 (tone/nn/modules/conformer_blocks.py:482) with the residual add and
// the RMSNorms around it in ConformerLayer.forward (:812-816, :832-836).
//
// Replaces the pair "split-K GEMM writing fp16 partial sums to HBM" + "norm kernel that sums them": 116 MB per 64-stream
// step of partial-sum traffic and one kernel launch per feed-forward go away.
//   * a CTA computes ALL 384 output columns of its 128 rows for its K slice (accumulator 128 x 384 fp32 = 384 TMEM
//     columns; per K step one A tile 128 x 64 and three 128-row weight boxes), so a row's sum of squares is local;
//   * the K dimension is split across the SK CTAs of a cluster (SK = 1, 2, 4, 8 chosen so that row tiles x SK fills the
//     GPU); after the main loop the partial accumulators are reduce-scattered over distributed shared memory: CTA k owns
//     rows [k * 128 / SK, (k + 1) * 128 / SK) of the tile, every CTA writes the owner's rows of its partial tile into the
//     owner's shared memory (st.shared::cluster, over the operand ring, which is dead by then), cluster barrier, and the
//     owner sums the SK partials in a fixed order (deterministic), adds bias and residual and applies the norms with one
//     warp per row (all global accesses are row-contiguous);
//   * the weight boxes of the first ring are issued before the programmatic-dependency wait: with SK = 8 (64 streams)
//     the whole K slice of the weights is in flight while the up-projection kernel is still running.
#pragma once

#include "ff_fused.cuh"

namespace tone {

constexpr int FD_THREADS = 320;
constexpr int FD_STAGES = 3;
constexpr int FD_STAGE_BYTES = 128 * 128 + 384 * 128;      // A tile + 3 weight boxes per K step of 64
constexpr int FD_X_PITCH = 388;                            // floats per staged row (thread-per-row stores conflict-free)
constexpr int FD_X_BYTES = 128 * FD_X_PITCH * 4;           // [SK][128 / SK][pitch] fp32
constexpr int FD_OPER_BYTES = FD_STAGES * FD_STAGE_BYTES > FD_X_BYTES ? FD_STAGES * FD_STAGE_BYTES : FD_X_BYTES;
constexpr int FD_SMEM_BYTES = FD_OPER_BYTES + 3 * 384 * 4 + 256 + 1024;
static_assert(FD_SMEM_BYTES <= 232448, "does not fit");

__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sts128_cluster(uint32_t cluster_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// args: FfArgs (ff_fused.cuh) - M, down_bias, r, scale, g1, g2, n, kv scatter; up_bias / ss are unused here
__global__ void __launch_bounds__(FD_THREADS, 1) ffdown_norm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB, const FfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_vec = reinterpret_cast<float*>(smem + FD_OPER_BYTES);                 // b2 | g1 | g2
  uint64_t* full = reinterpret_cast<uint64_t*>(s_vec + 3 * 384);
  uint64_t* empty = full + FD_STAGES;
  uint64_t* acc_full = empty + FD_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SK = (int)cluster_nctarank(), rank = (int)cluster_rank();
  const int tile = blockIdx.x / SK;                    // 128-row tile of this cluster
  const int nk = 24 / SK;                              // K steps of 64 in this CTA's slice
  const int k0 = rank * nk;

  PROF_DECL();
  PROF_BEGIN(8);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < FD_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer: weight boxes of the first ring before the dependency wait, then A, then steady state
    const int npre = nk < FD_STAGES ? nk : FD_STAGES;
    if (lane < npre) {
      mbar_expect_tx(&full[lane], FD_STAGE_BYTES);
      uint8_t* d = smem + lane * FD_STAGE_BYTES + 128 * 128;
#pragma unroll
      for (int j = 0; j < 3; ++j) tma_load_2d(d + j * 128 * 128, &tmB, &full[lane], (k0 + lane) * 64, j * 128);
    }
    pdl_wait();
    if (lane < npre) tma_load_2d(smem + lane * FD_STAGE_BYTES, &tmA, &full[lane], (k0 + lane) * 64, tile * 128);
    __syncwarp();
    for (int it = npre; it < nk; ++it) {
      const int s = it % FD_STAGES;
      mbar_wait(&empty[s], ((it / FD_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], FD_STAGE_BYTES);
        uint8_t* d = smem + s * FD_STAGE_BYTES;
        tma_load_2d(d, &tmA, &full[s], (k0 + it) * 64, tile * 128);
#pragma unroll
        for (int j = 0; j < 3; ++j) tma_load_2d(d + 128 * 128 + j * 128 * 128, &tmB, &full[s], (k0 + it) * 64, j * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: per K step 4 x (N = 256 on weight rows 0..255, N = 128 on rows 256..383)
    constexpr uint32_t idesc256 = make_idesc_bf16_mn(128, 256), idesc128 = make_idesc_bf16_mn(128, 128);
    const uint32_t sm_u = smem_u32(smem);
    for (int it = 0; it < nk; ++it) {
      const int s = it % FD_STAGES;
      mbar_wait(&full[s], (it / FD_STAGES) & 1);
      tc_fence_after();
      const uint64_t da = make_sw128_desc(sm_u + s * FD_STAGE_BYTES);
      const uint64_t d0 = make_sw128_desc(sm_u + s * FD_STAGE_BYTES + 128 * 128);
      const uint64_t d1 = make_sw128_desc(sm_u + s * FD_STAGE_BYTES + 128 * 128 + 256 * 128);
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_bf16(tmem_base, da + 2 * ks, d0 + 2 * ks, idesc256, (it > 0 || ks > 0) ? 1u : 0u);
          umma_bf16(tmem_base + 256, da + 2 * ks, d1 + 2 * ks, idesc128, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else {
    // constants of the epilogue (weights): b2 | g1 | g2 -> smem, before the dependency wait
    const int et = threadIdx.x - 64;
    for (int i = et; i < 384; i += EPI_THREADS) {
      s_vec[i] = __ldg(a.down_bias + i);
      s_vec[384 + i] = a.g1 ? __ldg(a.g1 + i) : 1.f;
      s_vec[768 + i] = a.g2 ? __ldg(a.g2 + i) : 1.f;
    }
    pdl_wait();
    mbar_wait(acc_full, 0);                              // every MMA of this CTA has completed: its operand ring is dead
    tc_fence_after();
  }
  // ---- the operand ring of EVERY CTA of the cluster is dead after this barrier: partial tiles may be written into it
  if (SK > 1) pair_sync_all();
  if (warp >= 2) {
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int R = 128 / SK;                              // rows owned per CTA
    // Phase A (thread = row, columns [192 hf, +192)): this CTA's partial of row r goes to the CTA that owns r, into slot
    // [source rank][r % R] of the owner's X buffer
    {
      const int owner = row_in_tile / R, r_local = row_in_tile - owner * R;
      const int c0 = hf * 192;
      const uint32_t local = smem_u32(smem) + ((rank * R + r_local) * FD_X_PITCH + c0) * 4;
      const uint32_t dst = SK > 1 ? map_to_rank(local, (uint32_t)owner) : local;
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
          const float4 v = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                                       __uint_as_float(acc[4 * i + 3]));
          if (SK > 1) sts128_cluster(dst + (cb + 4 * i) * 4, v);
          else sts128(dst + (cb + 4 * i) * 4, v);
        }
      }
    }
  }
  // ---- every partial has landed in its owner's shared memory
  if (SK > 1) pair_sync_all();
  else __syncthreads();
  if (warp >= 2) {
    // Phase B (warp = rows, lanes along the row): sum the SK partials in rank order, bias, residual, norms
    const int ew = warp - 2;
    const int R = 128 / SK;
    Vec384 g1v, g2v, b2v;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      b2v.v[i] = *reinterpret_cast<const float4*>(s_vec + i * 128 + lane * 4);
      g1v.v[i] = *reinterpret_cast<const float4*>(s_vec + 384 + i * 128 + lane * 4);
      g2v.v[i] = *reinterpret_cast<const float4*>(s_vec + 768 + i * 128 + lane * 4);
    }
    const uint32_t Xu = smem_u32(smem);
    for (int rl0 = ew * 2; rl0 < R; rl0 += 16) {          // two rows per warp and pass (R >= 16)
      float4 x[2][3];
      int grow[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        grow[k] = tile * 128 + rank * R + rl0 + k;
        if (grow[k] < a.M) {
#pragma unroll
          for (int i = 0; i < 3; ++i) x[k][i] = *reinterpret_cast<const float4*>(a.r + (size_t)grow[k] * D_MODEL + i * 128 + lane * 4);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (grow[k] >= a.M) continue;                       // warp-uniform
        float4 s[3] = {b2v.v[0], b2v.v[1], b2v.v[2]};
        for (int src = 0; src < SK; ++src) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float4 d = lds128(Xu + ((src * R + rl0 + k) * FD_X_PITCH + i * 128 + lane * 4) * 4);
            s[i].x += d.x;
            s[i].y += d.y;
            s[i].z += d.z;
            s[i].w += d.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          x[k][i].x = fmaf(a.scale, s[i].x, x[k][i].x);
          x[k][i].y = fmaf(a.scale, s[i].y, x[k][i].y);
          x[k][i].z = fmaf(a.scale, s[i].z, x[k][i].z);
          x[k][i].w = fmaf(a.scale, s[i].w, x[k][i].w);
        }
        float* rr = a.r + (size_t)grow[k] * D_MODEL;
        if (a.g1) scale_384(x[k], g1v, rms_inv_384(x[k]));
#pragma unroll
        for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(rr + i * 128 + lane * 4) = x[k][i];
        if (a.n) {
          if (a.g2) scale_384(x[k], g2v, rms_inv_384(x[k]));
          bf16* nr = a.n + (size_t)grow[k] * D_MODEL;
          bf16* kr = nullptr;
          if (a.kv) {
            const int b = grow[k] / a.rows_per_stream, t = grow[k] - b * a.rows_per_stream;
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
  // no CTA may exit while a peer could still write into its shared memory: the second cluster barrier above is the last
  // remote access, so a CTA-local barrier is enough here
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
  PROF_END();
}

inline cudaError_t configure_ffdown_norm() {
  cudaError_t e = cudaFuncSetAttribute(ffdown_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FD_SMEM_BYTES);
  return e;
}

// m_tiles = 128-row tiles, sk = K split = cluster size (1, 2, 4 or 8)
inline cudaError_t launch_ffdown_norm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const FfArgs& a, int m_tiles,
                                      int sk, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(m_tiles * sk);
  cfg.blockDim = dim3(FD_THREADS);
  cfg.dynamicSmemBytes = FD_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = sk;
  at[na].val.clusterDim.y = 1;
  at[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, ffdown_norm_kernel, tmA, tmB, a);
}

}  // namespace tone
