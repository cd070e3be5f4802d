// Fused attention block of the score-sharing layers for large batches (reference: RotaryMultiHeadAttention.forward with
// shared scores, tone/nn/modules/submodules.py:230-310, and the residual add of ConformerLayer.forward,
// conformer_blocks.py:816-822):
//
//     v = n Wv^T + bv ;  ctx[:, head] = P[head] v[:, head] (per stream) ;  r += ctx Wo^T + bo
//
// ONE kernel per tile of G whole streams x R frames (<= 128 rows) instead of the chain (V projection + P.V per head:
// 8 CTAs per tile, each re-reading the tile's A rows) -> ctx (bf16, HBM) -> out-projection GEMM (3 CTAs per tile):
//   * the tile's A rows (128 x 384 bf16, 96 KB) are loaded once; V = A Wv^T accumulates in TMEM columns [0, 384);
//   * eight epilogue warps walk the heads: V_h (+ bias, optional row-scale RMSNorm) -> fp32 staging (double buffered),
//     ctx_h = P_h V_h per stream on the CUDA cores (P was published by the last recompute layer), written as bf16
//     straight into shared memory in the K-major 128-byte-swizzled operand layout over the dead A tile;
//   * the out projection runs from that ctx tile into the same TMEM columns while its weights have been streaming through
//     the ring since the V GEMM released it; the final epilogue owns whole rows: r += acc + bo, bf16(r) and the row sum
//     of squares for the row-scale RMSNorm of the following GEMM (one ss tile per row).
#pragma once

#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace tone {

struct AttFArgs {
  int B;                   // streams in the batch
  int R, G;                // frames per stream at this layer's rate, streams per tile (128 / R)
  const float* P;          // [B][8][R][R] attention probabilities
  const float* bv;         // [384]
  const float* bo;         // [384]
  const float* ss;         // nullable: A = bf16(residual), row scale = 1 / (sqrt(sum_k ss[row][k]) / sqrt(384) + eps)
  int ss_ld, ss_tiles;
  float* r;                // [B * R][384] residual stream (in / out)
  bf16* rb_out;            // [B * R][384] bf16(r)
  float* ss_out;           // [B * R][ss_ld]: column 0 = sum of squares of the new residual row
};

constexpr int ATF_THREADS = 320;                 // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int ATF_KB = 6;                        // K blocks of 64 in d_model
constexpr int ATF_STAGES = 4;                    // weight ring: 128 rows x 64 k per stage
constexpr int ATF_TILE = 128 * 128;              // bytes of one operand K block
constexpr int ATF_VST = 128 * (2 * 48 + 4) * 4;  // v of two heads, fp32, [128][100]
constexpr int ATF_PST = 2 * 1600 * 4;            // P blocks of two heads: [2][G][R * R padded to 4] (<= 9 x 172 floats each)
constexpr int ATF_XP = 388;                      // floats per staged row of the final epilogue
constexpr int ATF_OPER = ATF_KB * ATF_TILE + ATF_STAGES * ATF_TILE + ATF_VST + ATF_PST;
constexpr int ATF_SMEM = ATF_OPER + 2 * 384 * 4 + 256 + 1024;
static_assert(128 * ATF_XP * 4 <= ATF_OPER, "the x tile is staged over the dead operand buffers");
static_assert(ATF_SMEM <= 232448, "does not fit");

__global__ void __launch_bounds__(ATF_THREADS, 1) att_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmWv,
                                                                   const __grid_constant__ CUtensorMap tmWo, const AttFArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // A rows, later the ctx tile
  uint8_t* sW = sA + ATF_KB * ATF_TILE;
  float* vst = reinterpret_cast<float*>(sW + ATF_STAGES * ATF_TILE);   // [128][100] v of two heads, then the P blocks
  float* s_vec = reinterpret_cast<float*>(smem + ATF_OPER);            // bv[384] | bo[384]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_vec + 2 * 384);
  uint64_t* full = bars;                   // [STAGES]
  uint64_t* empty = full + ATF_STAGES;     // [STAGES]
  uint64_t* a_full = empty + ATF_STAGES;   // [KB]
  uint64_t* vacc_full = a_full + ATF_KB;
  uint64_t* ctx_full = vacc_full + 1;
  uint64_t* oacc_full = ctx_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oacc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int R = a.R, G = a.G;
  const int row0 = tile * G * R;                         // first global row of the tile

  PROF_DECL();
  PROF_BEGIN(8);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmWv);
    tma_prefetch_desc(&tmWo);
    for (int s = 0; s < ATF_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int k = 0; k < ATF_KB; ++k) mbar_init(&a_full[k], 1);
    mbar_init(vacc_full, 1);
    mbar_init(ctx_full, 8);                // one arrive per epilogue warp
    mbar_init(oacc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer: 36 weight tiles through the ring (18 of Wv, then 18 of Wo; K block outer, N tile
    // inner) and the six K blocks of the A rows
    auto w_load = [&](int i) {
      const int s = i % ATF_STAGES;
      mbar_wait(&empty[s], ((i / ATF_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        const int j = i % 18, k = j / 3, nt = j - k * 3;
        mbar_expect_tx(&full[s], ATF_TILE);
        tma_load_2d(sW + s * ATF_TILE, i < 18 ? &tmWv : &tmWo, &full[s], k * 64, nt * 128);
      }
      __syncwarp();
    };
    for (int i = 0; i < ATF_STAGES; ++i) w_load(i);      // weights: before the dependency wait
    pdl_wait();
    if (lane == 0) PROF_MARK(2);
    if (elect_one_sync()) {
      for (int k = 0; k < ATF_KB; ++k) {
        mbar_expect_tx(&a_full[k], ATF_TILE);
        tma_load_2d(sA + k * ATF_TILE, &tmA, &a_full[k], k * 64, row0);
      }
    }
    __syncwarp();
    for (int i = ATF_STAGES; i < 36; ++i) w_load(i);
  } else if (warp == 1) {
    // ---------------- MMA issuer: V = A Wv^T, then (once the ctx tile is in shared memory) O = ctx Wo^T, both into
    // TMEM columns [0, 384)
    constexpr uint32_t idesc = make_idesc_bf16(128);
    const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 1) {
        mbar_wait(ctx_full, 0);
        tc_fence_after();
      }
      for (int k = 0; k < ATF_KB; ++k) {
        if (phase == 0) {
          mbar_wait(&a_full[k], 0);
        }
        for (int nt = 0; nt < 3; ++nt) {
          const int i = phase * 18 + k * 3 + nt, s = i % ATF_STAGES;
          mbar_wait(&full[s], (i / ATF_STAGES) & 1);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(sA_u + k * ATF_TILE);
          const uint64_t db = make_sw128_desc(sW_u + s * ATF_TILE);
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base + nt * 128, da + 2 * ks, db + 2 * ks, idesc, (k > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      }
      if (elect_one_sync()) umma_commit(phase == 0 ? vacc_full : oacc_full);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue warps 2..9: warp w owns TMEM lanes 32 (w % 4) .. +31 and column half hf = (w - 2) / 4
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int rt = q * 32 + lane;                        // row in tile
    const int g = rt / R, t = rt - g * R;
    const int b = tile * G + g;
    const bool valid = g < G && b < a.B;
    const long long grow = (long long)row0 + rt;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    for (int i = et; i < 384; i += EPI_THREADS) {
      s_vec[i] = __ldg(a.bv + i);
      s_vec[384 + i] = __ldg(a.bo + i);
    }
    bar_epilogue();
    pdl_wait();
    float rs = 1.f;
    if (a.ss && valid) {
      const float* sp = a.ss + grow * a.ss_ld;
      float tsum = 0.f;
      for (int k = 0; k < a.ss_tiles; ++k) tsum += sp[k];
      rs = 1.0f / (sqrtf(tsum) * 0.05103103630798288f + 1e-8f);
    }
    // P.V runs on (stream, head, 8-dim group) units with the stream's R x 8 slice of v in registers: every v element is
    // read from shared memory once per unit (not once per output row), which is what bounds this phase (a broadcast
    // LDS.128 costs 4 issue cycles whatever it broadcasts).  Two heads per pass: v of both heads staged in fp32, the
    // heads' R x R probability blocks of the tile's streams staged beside it.
    const int RR = R * R, RRP = (RR + 3) & ~3;           // floats per (stream, head) block of P, padded to 16 B
    float* pst = vst + 128 * (2 * D_HEAD + 4);           // [2 heads][G][RRP] behind the v stage [128][100]
    constexpr int VL2 = 2 * D_HEAD + 4;                  // floats per staged row: two heads + pad
    // P blocks of a pass ((head of the pass, stream): RR contiguous floats in global memory) -> pst with 4-byte cp.async,
    // one block per warp at a time; issued when the stage is free, awaited just before the pass's barrier
    const int ewp = warp - 2;
    auto stage_p = [&](int pass) {
      for (int blk = ewp; blk < 2 * G; blk += 8) {
        const int hh = blk >= G ? 1 : 0, gg = blk - hh * G;
        const int bb = tile * G + gg;
        if (bb < a.B) {
          const float* src = a.P + ((size_t)bb * N_HEADS + 2 * pass + hh) * RR;
          float* dst = pst + blk * RRP;
          for (int x = lane; x < RR; x += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + x)), "l"(src + x) : "memory");
        }
      }
      cp_async_commit();
    };
    stage_p(0);
    mbar_wait(vacc_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    const uint32_t ctx_base = smem_u32(sA);
#pragma unroll 1
    for (int pass = 0; pass < N_HEADS / 2; ++pass) {
      {
        // this thread's 48 columns of the pass: head 2 pass + hf, row rt
        const int c0 = (2 * pass + hf) * D_HEAD;
        uint32_t r0[16], r1[16], r2[16];
        tmem_ld16_async(tmem_base + lane_base + c0, r0);
        tmem_ld16_async(tmem_base + lane_base + c0 + 16, r1);
        tmem_ld16_async(tmem_base + lane_base + c0 + 32, r2);
        tmem_ld_wait();
        tmem_regs_ready16(r0);
        tmem_regs_ready16(r1);
        tmem_regs_ready16(r2);
        float* vr = vst + rt * VL2 + hf * D_HEAD;
        const float* bb = s_vec + c0;
#pragma unroll
        for (int c = 0; c < 48; c += 4) {
          const uint32_t* rr = c < 16 ? r0 + c : (c < 32 ? r1 + (c - 16) : r2 + (c - 32));
          float4 o;
          o.x = fmaf(__uint_as_float(rr[0]), rs, bb[c]);
          o.y = fmaf(__uint_as_float(rr[1]), rs, bb[c + 1]);
          o.z = fmaf(__uint_as_float(rr[2]), rs, bb[c + 2]);
          o.w = fmaf(__uint_as_float(rr[3]), rs, bb[c + 3]);
          *reinterpret_cast<float4*>(vr + c) = o;
        }
      }
      cp_async_wait_all();
      bar_epilogue();                                    // v and P of the pass are staged
      const int nunits = 2 * G * 6;                      // (head of the pass, stream, 8-dim group)
      for (int u = et; u < nunits; u += EPI_THREADS) {
        const int hh = u / (G * 6), rem = u - hh * (G * 6), gg = rem / 6, dg = rem - gg * 6;
        if (tile * G + gg >= a.B) continue;
        float v[VATT_MAX_T][8];
        const float* vsrc = vst + (gg * R) * VL2 + hh * D_HEAD + dg * 8;
#pragma unroll
        for (int j = 0; j < VATT_MAX_T; ++j) {
          if (j < R) {
            const float4 v0 = *reinterpret_cast<const float4*>(vsrc + j * VL2);
            const float4 v1 = *reinterpret_cast<const float4*>(vsrc + j * VL2 + 4);
            v[j][0] = v0.x; v[j][1] = v0.y; v[j][2] = v0.z; v[j][3] = v0.w;
            v[j][4] = v1.x; v[j][5] = v1.y; v[j][6] = v1.z; v[j][7] = v1.w;
          }
        }
        const float* pp = pst + (hh * G + gg) * RRP;
        const int chunk = (2 * pass + hh) * 6 + dg;      // 16-byte chunk of the 768-byte ctx row
        const int kb = chunk >> 3, cc = chunk & 7;
#pragma unroll 1
        for (int t = 0; t < R; ++t) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < VATT_MAX_T; ++j) {
            if (j < R) {
              const float p = pp[t * R + j];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, v[j][i], acc[i]);
            }
          }
          const int row = gg * R + t;
          sts128u(ctx_base + kb * ATF_TILE + row * 128 + ((cc ^ (row & 7)) << 4),
                  make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                             pack_bf16x2(acc[6], acc[7])));
        }
      }
      bar_epilogue();                                    // the stage is free for the next pass
      if (pass + 1 < N_HEADS / 2) stage_p(pass + 1);
    }
    tc_fence_before();                                   // this warp's TMEM reads are done: the out projection may overwrite
    fence_proxy_async();                                 // generic-proxy writes of the ctx tile -> visible to the tensor core
    __syncwarp();
    if (lane == 0) mbar_arrive(ctx_full);
    if (threadIdx.x == 64) PROF_MARK(1);             // head loop done

    // ---- final epilogue.  The residual rows of this warp's first batch are requested before the out projection has
    // finished (they do not depend on it).  Phase A (thread = row, columns [192 hf, +192)): acc + bo -> fp32 tile
    // X[128][388] staged over the operand buffers, which are dead once the last MMA has completed.  Phase B (warp = 16
    // rows, lanes along the row, 8 rows in flight): x = r + X -> r, bf16(x) -> rb, sum x^2 -> ss.
    const int ew = warp - 2;
    const int nrows = G * R;
    float4 x[8][3];
    auto row_ok = [&](int rr) { return rr < nrows && (tile * G + rr / R) < a.B; };
    auto load_rows = [&](int rg) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = ew * 16 + rg + k;
        if (row_ok(rr)) {
          const float* src = a.r + ((size_t)row0 + rr) * D_MODEL;
#pragma unroll
          for (int i = 0; i < 3; ++i) x[k][i] = *reinterpret_cast<const float4*>(src + i * 128 + lane * 4);
        }
      }
    };
    load_rows(0);
    mbar_wait(oacc_full, 0);
    if (threadIdx.x == 64) PROF_MARK(3);             // out projection complete
    tc_fence_after();
    float* X = reinterpret_cast<float*>(smem);
    {
      const int c0 = hf * 192;
      const uint32_t xrow = smem_u32(X) + (rt * ATF_XP + c0) * 4;
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
          const float4 bb = *reinterpret_cast<const float4*>(s_vec + 384 + c0 + cb + 4 * i);
          sts128(xrow + (cb + 4 * i) * 4,
                 make_float4(__uint_as_float(acc[4 * i]) + bb.x, __uint_as_float(acc[4 * i + 1]) + bb.y,
                             __uint_as_float(acc[4 * i + 2]) + bb.z, __uint_as_float(acc[4 * i + 3]) + bb.w));
        }
      }
    }
    bar_epilogue();
#pragma unroll 1
    for (int rg = 0; rg < 16; rg += 8) {
      if (rg) load_rows(rg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = ew * 16 + rg + k;
        if (!row_ok(rr)) continue;                          // warp-uniform
        const size_t gr = (size_t)row0 + rr;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 d = lds128(smem_u32(X) + (rr * ATF_XP + i * 128 + lane * 4) * 4);
          x[k][i].x += d.x;
          x[k][i].y += d.y;
          x[k][i].z += d.z;
          x[k][i].w += d.w;
          sq += x[k][i].x * x[k][i].x + x[k][i].y * x[k][i].y + x[k][i].z * x[k][i].z + x[k][i].w * x[k][i].w;
        }
        float* dst = a.r + gr * D_MODEL;
        bf16* rb = a.rb_out + gr * D_MODEL;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          *reinterpret_cast<float4*>(dst + i * 128 + lane * 4) = x[k][i];
          *reinterpret_cast<uint2*>(rb + i * 128 + lane * 4) =
              make_uint2(pack_bf16x2(x[k][i].x, x[k][i].y), pack_bf16x2(x[k][i].z, x[k][i].w));
        }
        sq = warp_sum(sq);
        if (lane == 0) a.ss_out[gr * a.ss_ld] = sq;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
  PROF_END();
}

inline cudaError_t configure_att_fused() {
  return cudaFuncSetAttribute(att_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATF_SMEM);
}

inline cudaError_t launch_att_fused(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmWv, const CUtensorMap& tmWo,
                                    const AttFArgs& a, bool pdl) {
  const int tiles = (a.B + a.G - 1) / a.G;
  return launch_kernel(att_fused_kernel, dim3(tiles), dim3(ATF_THREADS), ATF_SMEM, st, pdl, tmA, tmWv, tmWo, a);
}

}  // namespace tone
