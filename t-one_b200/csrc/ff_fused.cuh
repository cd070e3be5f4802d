// Fused feed-forward module for large batches (reference: ConformerFeedForward.forward, tone/nn/modules/
// conformer_blocks.py:468-482, with the surrounding residual add and RMSNorms of ConformerLayer.forward :812-814,:832-836):
//
//     r += 0.5 * ( W2 ( silu(W1 a + b1) * (Wv a + bv) ) + b2 ) ;   [r = g1 * r / (rms(r) + eps)] ;   n = g2 * r / (rms(r) + eps)
//
// ONE kernel per 128-row tile (a CTA, or 256 rows on a CTA pair with tcgen05.mma.cta_group::2) instead of the chain
// up-GEMM -> h (bf16, HBM) -> down-GEMM (split-K partials, HBM) -> norm kernel:
//   * the tile's A rows (128 x 384 bf16, 96 KB) are loaded once and stay in shared memory;
//   * the hidden dimension is walked in 24 chunks of 64 units.  Per chunk the up GEMM (N = 128: 64 gate | 64 value
//     columns, K = 384) accumulates in TMEM columns [384, 512); 8 epilogue warps turn it into the bf16 hidden chunk
//     (SiLU gate) written straight into shared memory in the K-major 128-byte-swizzled operand layout; the down GEMM then
//     accumulates h_chunk (128 x 64) . W2[:, chunk]^T into TMEM columns [0, 384).  The hidden activation never leaves
//     the SM and there are no split-K partials;
//   * the MMA warp issues up(j+1) BEFORE down(j): the tensor pipe works on the next chunk while the epilogue warps
//     convert the current one (only the TMEM drain of the single up accumulator is exposed, and down(j-1) covers it);
//   * the final epilogue owns whole rows, so the residual add and both RMSNorms are thread-local: pass 1 forms
//     x = r + 0.5 (acc + b2), parks it back in TMEM and accumulates sum x^2 and sum (g1 x)^2; pass 2 writes r and the
//     normalised bf16 rows (optionally also into the per-stream [cache | new] rows of layers 14 / 15).
// Weight streaming: up weights 96 KB + down weights 48 KB per chunk (1.2 us of MMA) - the pair form halves that per SM
// (each CTA loads half of every weight tile; both tensor cores read both halves).
#pragma once

#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace tone {

struct FfArgs {
  int M;                    // valid rows
  const float* up_bias;     // [3072] interleaved like the weights: per 128-column tile 64 gate | 64 value
  const float* down_bias;   // [384]
  const float* ss;          // nullable: A = bf16(residual), row scale = 1 / (sqrt(sum_k ss[row][k]) / sqrt(384) + eps)
  int ss_ld, ss_tiles;
  float* r;                 // [M][384] residual stream (in / out)
  float scale;              // 0.5
  const float* g1;          // nullable: norm_out applied to r in place
  const float* g2;          // nullable: gain of the RMSNorm that produces n
  bf16* n;                  // nullable: [M][384] bf16 rows out
  bf16* kv;                 // nullable: scatter n rows into [slots][KV_ROWS_MAX][384] at row kv_row_off + t
  const int* slots;
  int rows_per_stream, kv_row_off;
};

constexpr int FF_THREADS = 352;                 // warp 0: up-weight TMA, warp 1: MMA, warps 2..9: epilogue, warp 10: A + down-weight TMA
constexpr int FF_CHUNKS = 24;                    // 1536 hidden units / 64
constexpr int FF_KB = 6;                         // K blocks of 64 in d_model = 384

template <bool PAIR>
struct FfCfg {
  static constexpr int A_BYTES = FF_KB * 128 * 128;                 // 98304
  static constexpr int H_BYTES = 128 * 128;                         // one hidden chunk, 128 rows x 64 bf16
  static constexpr int UP_STG = PAIR ? 64 * 128 : 128 * 128;        // one K block of this CTA's share of an up tile
  static constexpr int S_UP = PAIR ? 6 : 3;                         // pair: a whole chunk of up weights in flight
  static constexpr int DN_SLAB = PAIR ? 192 * 128 : 384 * 128;      // this CTA's share of W2[:, chunk]
  static constexpr int S_DN = PAIR ? 2 : 1;
  static constexpr int VEC_BYTES = 3 * 384 * 4;                     // b2 | g1 | g2
  static constexpr int OPER_BYTES = A_BYTES + H_BYTES + S_UP * UP_STG + S_DN * DN_SLAB;
  static constexpr int X_PITCH = 388;                               // floats per staged row of the final epilogue (conflict-free)
  static constexpr int SMEM_BYTES = OPER_BYTES + VEC_BYTES + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "does not fit");
  static_assert(128 * X_PITCH * 4 <= OPER_BYTES, "the x tile is staged over the dead operand buffers");
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Signals from the epilogue warps to the MMA warp of the leader CTA.  `mbarrier.arrive.release.cluster` costs ~1500 cycles
// per use (measured: it was 40 % of the per-chunk epilogue), so it is kept off the epilogue warps:
//  * "accumulator drained" carries no data written by these threads (the tcgen05 fence orders the TMEM reads): leader
//    warps arrive at CTA scope, peer warps with a relaxed cluster-scope arrive;
//  * "hidden chunk written" is a CTA-scope arrive on the CTA's OWN barrier; in the peer CTA the otherwise idle warp 1
//    relays each completed phase to the leader with one release.cluster arrive.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// remote arrive with the default semantics (release at CTA scope): what CUTLASS' ClusterBarrier::arrive(cta_id) emits
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ float silu_mul(float x, float y) { return silu_f(x) * y; }

template <bool PAIR>
__global__ void __launch_bounds__(FF_THREADS, 1) ff_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmUp,
                                                                   const __grid_constant__ CUtensorMap tmDn,
                                                                   const __grid_constant__ CUtensorMap tmDn64, const FfArgs a) {
  using Cfg = FfCfg<PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sH = sA + Cfg::A_BYTES;
  uint8_t* sUp = sH + Cfg::H_BYTES;
  uint8_t* sDn = sUp + Cfg::S_UP * Cfg::UP_STG;
  float* s_vec = reinterpret_cast<float*>(sDn + Cfg::S_DN * Cfg::DN_SLAB);        // b2[384] | g1[384] | g2[384]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_vec + 3 * 384);
  uint64_t* up_full = bars;                      // [S_UP]
  uint64_t* up_empty = up_full + Cfg::S_UP;      // [S_UP]
  uint64_t* dn_full = up_empty + Cfg::S_UP;      // [S_DN]
  uint64_t* dn_empty = dn_full + Cfg::S_DN;      // [S_DN]
  uint64_t* a_full = dn_empty + Cfg::S_DN;
  uint64_t* upacc_full = a_full + 1;
  uint64_t* upacc_empty = upacc_full + 1;
  uint64_t* h_full = upacc_empty + 1;
  uint64_t* h_empty = h_full + 1;
  uint64_t* dnacc_full = h_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dnacc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const bool leader = rank == 0;
  const int tile = blockIdx.x;                    // 128-row tile of this CTA (pairs: CTAs 2p, 2p + 1 = rows 256p ..)
  constexpr int NCTA = PAIR ? 2 : 1;

  PROF_DECL();
  PROF_BEGIN(7);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmUp);
    tma_prefetch_desc(&tmDn);
    if constexpr (PAIR) tma_prefetch_desc(&tmDn64);
    for (int s = 0; s < Cfg::S_UP; ++s) {
      mbar_init(&up_full[s], 1);
      mbar_init(&up_empty[s], 1);
    }
    for (int s = 0; s < Cfg::S_DN; ++s) {
      mbar_init(&dn_full[s], 1);
      mbar_init(&dn_empty[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(upacc_full, 1);
    mbar_init(upacc_empty, 8 * NCTA);            // one arrive per epilogue warp (of both CTAs)
    mbar_init(h_full, 8 * NCTA);                 // every epilogue warp of the pair arrives on the leader's barrier
    mbar_init(h_empty, 1);
    mbar_init(dnacc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(tmem_slot);
    }
  }
  tc_fence_before();
  if constexpr (PAIR) pair_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dnacc = tmem_base, upacc = tmem_base + 384;

  if (warp == 0 || warp == 10) {
    // ---------------- TMA producers: warp 0 streams the up weights (144 K blocks, weights only: no dependency on the
    // predecessor kernel), warp 10 the A tile and the down-weight slabs.  Each CTA of a pair loads its own rows of A and
    // its half of every weight tile; all bytes complete on the leader's barriers.
    auto load2d = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
      if constexpr (PAIR) tma_load_2d_pair(dst, m, map_to_rank(smem_u32(bar), 0), x, y);
      else tma_load_2d(dst, m, bar, x, y);
    };
    if (warp == 0) {
      for (int cnt = 0; cnt < FF_CHUNKS * FF_KB; ++cnt) {
        const int j = cnt / FF_KB, k = cnt - j * FF_KB, s = cnt % Cfg::S_UP;
        mbar_wait(&up_empty[s], ((cnt / Cfg::S_UP) & 1) ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_expect_tx(&up_full[s], NCTA * Cfg::UP_STG);
          load2d(sUp + s * Cfg::UP_STG, &tmUp, &up_full[s], k * 64, j * 128 + (PAIR ? (int)rank * 64 : 0));
        }
        __syncwarp();
      }
    } else {
      auto load_dn = [&](int j) {
        const int s = j % Cfg::S_DN;
        mbar_wait(&dn_empty[s], ((j / Cfg::S_DN) & 1) ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_expect_tx(&dn_full[s], NCTA * Cfg::DN_SLAB);
          uint8_t* d = sDn + s * Cfg::DN_SLAB;
          if constexpr (PAIR) {
            load2d(d, &tmDn, &dn_full[s], j * 64, (int)rank * 128);                      // rows of the N = 256 MMA
            load2d(d + 128 * 128, &tmDn64, &dn_full[s], j * 64, 256 + (int)rank * 64);   // rows of the N = 128 MMA
          } else {
            load2d(d, &tmDn, &dn_full[s], j * 64, 0);
            load2d(d + 128 * 128, &tmDn, &dn_full[s], j * 64, 128);
            load2d(d + 256 * 128, &tmDn, &dn_full[s], j * 64, 256);
          }
        }
        __syncwarp();
      };
      for (int j = 0; j < Cfg::S_DN; ++j) load_dn(j);      // weights: before the dependency wait
      pdl_wait();
      if (elect_one_sync()) {
        if (leader) mbar_expect_tx(a_full, NCTA * Cfg::A_BYTES);
        for (int k = 0; k < FF_KB; ++k) load2d(sA + k * 128 * 128, &tmA, a_full, k * 64, tile * 128);
      }
      __syncwarp();
      for (int j = Cfg::S_DN; j < FF_CHUNKS; ++j) load_dn(j);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA of a pair): up(0), then per chunk up(j) followed by down(j - 1)
    if (leader) {
      constexpr uint32_t idesc_up = make_idesc_bf16_mn(PAIR ? 256 : 128, 128);
      constexpr uint32_t idesc_d256 = make_idesc_bf16_mn(PAIR ? 256 : 128, 256);
      constexpr uint32_t idesc_d128 = make_idesc_bf16_mn(PAIR ? 256 : 128, 128);
      const uint32_t sA_u = smem_u32(sA), sH_u = smem_u32(sH), sUp_u = smem_u32(sUp), sDn_u = smem_u32(sDn);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma_bf16_pair(d, da, db, idesc, acc);
        else umma_bf16(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      int up_cnt = 0;
#ifdef TONE_PROF
      long long w_up = 0, w_acc = 0, w_h = 0, w_dn = 0, t_;
#define FF_TIMED(acc_, stmt) t_ = clock64(); stmt; acc_ += clock64() - t_
#else
#define FF_TIMED(acc_, stmt) stmt
#endif
      auto down = [&](int i) {
        const int s = i % Cfg::S_DN;
        FF_TIMED(w_h, mbar_wait(h_full, i & 1));
        FF_TIMED(w_dn, mbar_wait(&dn_full[s], (i / Cfg::S_DN) & 1));
        tc_fence_after();
        const uint64_t dh = make_sw128_desc(sH_u);
        const uint64_t d0 = make_sw128_desc(sDn_u + s * Cfg::DN_SLAB);
        const uint64_t d1 = make_sw128_desc(sDn_u + s * Cfg::DN_SLAB + (PAIR ? 128 * 128 : 256 * 128));
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma(dnacc, dh + 2 * ks, d0 + 2 * ks, idesc_d256, (i > 0 || ks > 0) ? 1u : 0u);
            mma(dnacc + 256, dh + 2 * ks, d1 + 2 * ks, idesc_d128, (i > 0 || ks > 0) ? 1u : 0u);
          }
          commit(&dn_empty[s]);
          commit(h_empty);
        }
        __syncwarp();
      };
      mbar_wait(a_full, 0);
#ifdef TONE_PROF
      long long tm[5] = {0, 0, 0, 0, 0};
      long long tch[FF_CHUNKS + 1];
      const long long t_loop0 = clock64();
#endif
      for (int j = 0; j < FF_CHUNKS; ++j) {
#ifdef TONE_PROF
        if (j == 12) tm[0] = clock64();
        tch[j] = clock64();
#endif
        FF_TIMED(w_acc, mbar_wait(upacc_empty, (j & 1) ^ 1));     // the epilogue has drained the previous chunk's accumulator
        tc_fence_after();
        for (int k = 0; k < FF_KB; ++k, ++up_cnt) {
          const int s = up_cnt % Cfg::S_UP;
          FF_TIMED(w_up, mbar_wait(&up_full[s], (up_cnt / Cfg::S_UP) & 1));
          tc_fence_after();
          const uint64_t da = make_sw128_desc(sA_u + k * 128 * 128);
          const uint64_t db = make_sw128_desc(sUp_u + s * Cfg::UP_STG);
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma(upacc, da + 2 * ks, db + 2 * ks, idesc_up, (k > 0 || ks > 0) ? 1u : 0u);
            commit(&up_empty[s]);
          }
          __syncwarp();
        }
#ifdef TONE_PROF
        if (j == 12) tm[1] = clock64();
#endif
        if (elect_one_sync()) commit(upacc_full);
        __syncwarp();
#ifdef TONE_PROF
        if (j == 12) tm[2] = clock64();
#endif
        if (j > 0) down(j - 1);
#ifdef TONE_PROF
        if (j == 12) tm[3] = clock64();
        if (j == 13) tm[4] = clock64();
#endif
      }
      down(FF_CHUNKS - 1);
      if (elect_one_sync()) commit(dnacc_full);
      __syncwarp();
#ifdef TONE_PROF
      if (lane == 0 && blockIdx.x == 0 && g_prof) {
        tch[FF_CHUNKS] = clock64();
        printf("ff loop: start %lld | chunk starts rel: %lld %lld %lld %lld %lld %lld %lld %lld %lld end %lld\n", t_loop0 - g_prof[prof_seq_s].c[0],
               tch[0] - t_loop0, tch[1] - t_loop0, tch[2] - t_loop0, tch[3] - t_loop0, tch[6] - t_loop0, tch[12] - t_loop0,
               tch[18] - t_loop0, tch[22] - t_loop0, tch[23] - t_loop0, tch[24] - t_loop0);
      }
      if (lane == 0 && blockIdx.x == 0 && g_prof)
        printf("ff mma chunk12: wait_upacc_empty+U issue %lld commit %lld down(11) %lld  iter %lld | stalls up %lld acc %lld h %lld dn %lld\n",
               tm[1] - tm[0], tm[2] - tm[1], tm[3] - tm[2], tm[4] - tm[0] - (tm[4] - tm[3]) + (tm[4] - tm[3]), w_up, w_acc, w_h, w_dn);
      if (lane == 0 && blockIdx.x == 0 && g_prof) {   // stall cycles of the MMA warp: up weights | up-acc drain | hidden chunk | down weights
        g_prof[prof_seq_s].c[1] = g_prof[prof_seq_s].c[0] + w_up;
        g_prof[prof_seq_s].c[2] = g_prof[prof_seq_s].c[0] + w_acc;
        g_prof[prof_seq_s].c[3] = g_prof[prof_seq_s].c[0] + w_h + w_dn;
      }
#endif
    }
  } else {
    // ---------------- epilogue warps 2..9: warp w owns TMEM lanes 32 (w % 4) .. +31 and column half hf = (w - 2) / 4
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int row_in_tile = q * 32 + lane;
    const int row = tile * 128 + row_in_tile;
    const bool valid = row < a.M;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    // constants of the final epilogue (weights): b2 | g1 | g2 -> smem, before the dependency wait
    for (int i = et; i < 384; i += EPI_THREADS) {
      s_vec[i] = __ldg(a.down_bias + i);
      s_vec[384 + i] = a.g1 ? __ldg(a.g1 + i) : 1.f;
      s_vec[768 + i] = a.g2 ? __ldg(a.g2 + i) : 1.f;
    }
    pdl_wait();
    float rs = 1.f;                               // row scale of the folded RMSNorm (A = un-normalised residual)
    if (a.ss && valid) {
      const float* sp = a.ss + (size_t)row * a.ss_ld;
      float t = 0.f;
      for (int k = 0; k < a.ss_tiles; ++k) t += sp[k];
      rs = 1.0f / (sqrtf(t) * 0.05103103630798288f + 1e-8f);
    }
    const uint32_t h_row = smem_u32(sH) + row_in_tile * 128;
#ifdef TONE_PROF
    long long ts[6] = {0, 0, 0, 0, 0, 0};
#define FF_TS(i) if (j == 12) ts[i] = clock64()
#else
#define FF_TS(i)
#endif
    for (int j = 0; j < FF_CHUNKS; ++j) {
      // biases of this thread's 32 gate / 32 value columns (identical for every lane: broadcast loads)
      const float4* bg = reinterpret_cast<const float4*>(a.up_bias + j * 128 + hf * 32);
      const float4* bv = reinterpret_cast<const float4*>(a.up_bias + j * 128 + 64 + hf * 32);
      FF_TS(0);
      mbar_wait(upacc_full, j & 1);
      FF_TS(1);
      tc_fence_after();
      uint32_t g[32], v[32];
      tmem_ld16_async(upacc + lane_base + hf * 32, g);
      tmem_ld16_async(upacc + lane_base + hf * 32 + 16, g + 16);
      tmem_ld16_async(upacc + lane_base + 64 + hf * 32, v);
      tmem_ld16_async(upacc + lane_base + 64 + hf * 32 + 16, v + 16);
      tmem_ld_wait();
      tmem_regs_ready16(g);
      tmem_regs_ready16(g + 16);
      tmem_regs_ready16(v);
      tmem_regs_ready16(v + 16);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                        // the next chunk's up GEMM may overwrite the accumulator
        if (PAIR && !leader) mbar_arrive_cluster_relaxed(map_to_rank(smem_u32(upacc_empty), 0));
        else mbar_arrive(upacc_empty);
      }
      FF_TS(2);
      uint32_t hp[16];
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 b0 = __ldg(bg + (c >> 2)), b1 = __ldg(bv + (c >> 2));
        const float gb[4] = {b0.x, b0.y, b0.z, b0.w}, vb[4] = {b1.x, b1.y, b1.z, b1.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o[i] = silu_mul(fmaf(__uint_as_float(g[c + i]), rs, gb[i]), fmaf(__uint_as_float(v[c + i]), rs, vb[i]));
        hp[(c >> 1)] = pack_bf16x2(o[0], o[1]);
        hp[(c >> 1) + 1] = pack_bf16x2(o[2], o[3]);
      }
      FF_TS(3);
      mbar_wait(h_empty, (j & 1) ^ 1);                        // down(j - 1) has read the previous hidden chunk
      FF_TS(4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {                           // 16-byte chunks hf * 4 + i of the row, 128-byte swizzle
        const int phys = (hf * 4 + i) ^ (row_in_tile & 7);
        sts128u(h_row + (phys << 4), make_uint4(hp[4 * i], hp[4 * i + 1], hp[4 * i + 2], hp[4 * i + 3]));
      }
      fence_proxy_async();                                    // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_remote(map_to_rank(smem_u32(h_full), 0));
        else mbar_arrive(h_full);
      }
      FF_TS(5);
    }
#ifdef TONE_PROF
    if (blockIdx.x == 0 && threadIdx.x == 64 && g_prof)
      printf("ff epi chunk12: wait_upacc %lld ld %lld math %lld wait_h %lld store %lld\n", ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2],
             ts[4] - ts[3], ts[5] - ts[4]);
#endif
    // ---- final epilogue.  Phase A (thread = row, columns [192 hf, +192)): scale * (acc + b2) -> fp32 tile X[128][388]
    // staged over the operand buffers, which are dead once the last MMA has completed.
    mbar_wait(dnacc_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    float* X = reinterpret_cast<float*>(smem);
    {
      const int c0 = hf * 192;
      const uint32_t xrow = smem_u32(X) + (row_in_tile * Cfg::X_PITCH + c0) * 4;
#pragma unroll 1
      for (int cb = 0; cb < 192; cb += 32) {
        uint32_t acc[32];
        tmem_ld16_async(dnacc + lane_base + c0 + cb, acc);
        tmem_ld16_async(dnacc + lane_base + c0 + cb + 16, acc + 16);
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
    // Phase B (warp = 16 rows, lanes along the row: every global access is a contiguous 512 / 256 bytes):
    // x = r + X ; [y = g1 x / (rms(x) + eps) -> r] ; n = g2 y / (rms(y) + eps)
    {
      const int ew = warp - 2;
      Vec384 g1v, g2v;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        g1v.v[i] = *reinterpret_cast<const float4*>(s_vec + 384 + i * 128 + lane * 4);
        g2v.v[i] = *reinterpret_cast<const float4*>(s_vec + 768 + i * 128 + lane * 4);
      }
#pragma unroll 1
      for (int rg = 0; rg < 16; rg += 4) {
        float4 x[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // four rows of residual in flight
          const int rt = ew * 16 + rg + k, grow = tile * 128 + rt;
          if (grow < a.M) {
#pragma unroll
            for (int i = 0; i < 3; ++i) x[k][i] = *reinterpret_cast<const float4*>(a.r + (size_t)grow * D_MODEL + i * 128 + lane * 4);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rt = ew * 16 + rg + k, grow = tile * 128 + rt;
          if (grow >= a.M) continue;                        // warp-uniform
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float4 d = lds128(smem_u32(X) + (rt * Cfg::X_PITCH + i * 128 + lane * 4) * 4);
            x[k][i].x += d.x;
            x[k][i].y += d.y;
            x[k][i].z += d.z;
            x[k][i].w += d.w;
          }
          float* rr = a.r + (size_t)grow * D_MODEL;
          if (a.g1) scale_384(x[k], g1v, rms_inv_384(x[k]));
#pragma unroll
          for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(rr + i * 128 + lane * 4) = x[k][i];
          if (a.n) {
            if (a.g2) scale_384(x[k], g2v, rms_inv_384(x[k]));
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
  }
  tc_fence_before();
  if constexpr (PAIR) pair_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else tmem_dealloc<512>(tmem_base);
  }
  PROF_END();
}

template <bool PAIR>
inline cudaError_t configure_ff_fused() {
  return cudaFuncSetAttribute(ff_fused_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<PAIR>::SMEM_BYTES);
}

// m_tiles = 128-row tiles; the pair form rounds the grid up to whole pairs (the odd tile's rows are invalid and skipped)
template <bool PAIR>
inline cudaError_t launch_ff_fused(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmUp, const CUtensorMap& tmDn,
                                   const CUtensorMap& tmDn64, const FfArgs& a, int m_tiles, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(PAIR ? 2 * ((m_tiles + 1) / 2) : m_tiles);
  cfg.blockDim = dim3(FF_THREADS);
  cfg.dynamicSmemBytes = FfCfg<PAIR>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, ff_fused_kernel<PAIR>, tmA, tmUp, tmDn, tmDn64, a);
}

}  // namespace tone
