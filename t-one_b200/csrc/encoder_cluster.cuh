// Latency path of the Conformer stack (reference tone/nn/modules/conformer_blocks.py:799-836,
// conformer.py:216-227): all 16 layers, the temporal reduction / upsampling and the decoder in ONE kernel.
//
// At 64 streams the per-kernel path is bound by ~200 dependent launches, not by the SMs.  Streams are independent
// recurrences, so a thread-block CLUSTER of 8 CTAs takes a group of G streams (R = G*T <= 56 rows) through the whole
// stack without any grid-wide synchronisation:
//   * the model dimension is split 8 ways inside the cluster: CTA c owns attention head c (48 of 384 features),
//     depthwise-conv channels [48c, 48c+48) and 192 of the 1536 feed-forward hidden features;
//   * every sub-block is (column-parallel GEMM from the normalised input, replicated in each CTA's shared memory)
//     -> local epilogue -> (row-parallel GEMM over the CTA's own K slice) -> partial rows sent over DSMEM to the
//     ROW OWNER (row j belongs to CTA j % 8) -> the owner sums the 8 partials in a fixed order, adds bias and residual,
//     applies the RMSNorm(s) and broadcasts the bf16 row into all 8 CTAs' operand buffers (DSMEM all-gather);
//   * GEMMs are tcgen05.mma with TMEM accumulators.  With <= 56 rows the tensor-core cost is proportional to the MMA N
//     extent, so the big GEMMs run "swapped": the WEIGHT tile is the M = 128 operand and the activation rows are N
//     (32..64); the epilogue then has one output feature per thread, all 8 worker warps busy, and DSMEM stores that are
//     contiguous per warp.  The q/k/v, GLU and decoder GEMMs keep rows on M because their epilogues (LayerNorm over a
//     head, RoPE, log-softmax) want a whole row per thread;
//   * weights stream through a TMA ring that a free-running producer warp keeps full across phase boundaries;
//     activations never leave shared memory;
//   * cross-CTA ordering uses cluster-scope mbarriers (partials arrived / operand rows arrived / scratch free), so
//     only the warps that need the data wait.
// Numerics follow the per-kernel path (bf16 GEMM operands, fp32 accumulate, fp32 residual / norms / softmax / RoPE);
// the K-slice partial sums cross the cluster in fp16 (11-bit mantissa, saturating), summed in fp32 in a fixed order.
#pragma once

#include "common.cuh"
#include "kernels.cuh"

namespace tone {

constexpr int CL_CTAS = 8;
constexpr int CL_THREADS = 352;          // warps 0..7 workers, warp 8 TMA producer, warp 9 MMA issuer + TMEM owner, warp 10 L2 prefetcher
constexpr int CL_PREFETCH_AHEAD = 40;    // chunks (<= 16 KB each) the L2 prefetcher runs ahead of the ring producer
constexpr int CL_WORKERS = 256;
constexpr int CL_STAGE_BYTES = 16384;    // one ring stage: up to 128 weight rows x 64 K (bf16)
constexpr int CL_RECV_LD = 392;          // halfs per received partial row (384 + 8)

// tensor-map slots: per layer, then the globals
enum ClMap : int { CM_FF1_UP = 0, CM_FF1_DOWN, CM_QKV, CM_KV, CM_WO, CM_PW1, CM_PW2, CM_FF2_UP, CM_FF2_DOWN, CM_PER_LAYER };
enum ClMapG : int { CG_RED_PW = 16 * CM_PER_LAYER, CG_DEC, CG_KVC14, CG_KVC15, CG_TOTAL };

struct ClLayer {
  const float *ff1_up_b, *ff1_down_b, *ff2_up_b, *ff2_down_b;
  const float *qkv_b;      // l < 14: [q|k|v] (recompute) or [v]; l >= 14: unused
  const float *q_b, *kv_b; // l >= 14
  const float *wo_b, *pw1_b /* cluster packing: per CTA [a 48 | b 48] */, *pw2_b;
  const float *g_ff1, *g_att, *g_out;          // gains applied by the row owner (norm_conv / norm_ff2 are folded into weights)
  const float *qln_w, *qln_b, *kln_w, *kln_b;
  const float *dw_w, *dw_b;                    // [31][384] BN-folded taps, [384]
};

struct ClParams {
  const CUtensorMap* maps;     // [CG_TOTAL] in global memory
  ClLayer L[16];
  const float *red_dw_w, *red_dw_b, *red_pw_b, *dec_b;
  const float *rope_cos, *rope_sin;
  // per-stream state
  bf16 *kv14, *kv15, *conv;    // [slots][KV_ROWS_MAX][384] x2, [slots][16][30][384]
  float* red;                  // [slots][384]
};

// what changes per launch (passed by value: the captured graph keeps it)
struct ClStep {
  const int* slots;            // [B]
  const int* len_in;           // [B]
  const float* r_in;           // [B*T][384] pre-encode output (out_norm applied)
  float* res;                  // [B*T][384] scratch: layer-6 output kept for the upsampling residual
  float* logprobs;             // [B*T][35]
  int* tokens;                 // [B*T]
  float* aux;                  // [B*T][2]
  float* taps;                 // nullable: [17][tap_rows][384]
  long long tap_stride;        // floats between taps
  int B;
  unsigned long long* prof;    // nullable diagnostics: [0,2048) worker marks (ns), [2048,6144) MMA lane (t0, t1, stall clk, chunks)
};

template <int T_, int G_>
struct ClCfg {
  static constexpr int T = T_, G = G_;
  static constexpr int T2 = (T + 1 - 3) / 2 + 1;
  static constexpr int R = G * T, R2 = G * T2;
  static constexpr int RA = (R + 7) / 8 * 8;            // rows held per operand k-block tile
  static constexpr int RO = (R + 7) / 8;                // rows owned per CTA (full rate)
  static constexpr int NP = (R + 15) / 16 * 16;         // MMA N of the swapped GEMMs at full rate
  static constexpr int NP2 = (R2 + 15) / 16 * 16;       // ... at the reduced rate
  static constexpr int KB_BYTES = RA * 128;             // one [RA][64] bf16 k-block tile
  static constexpr int TKMAX = MHSA_S + T;
  // byte offsets from the 1024-aligned base
  static constexpr int OFF_A = 0;                       // normalised input, 6 k-blocks
  static constexpr int OFF_H = OFF_A + 6 * KB_BYTES;    // local hidden slice, 3 k-blocks
  static constexpr int OFF_RING = (OFF_H + 3 * KB_BYTES + 1023) / 1024 * 1024;
  static constexpr int RECV_BYTES = CL_CTAS * RO * CL_RECV_LD * 2;
  static constexpr int QS_LD = 52;                      // floats per q row (16 B aligned, conflict-light)
  static constexpr int SCR_ATT = R * QS_LD * 4 + G * TKMAX * D_HEAD * 4;         // q rows + v rows
  static constexpr int SCR_DW = G * (CONV_S + T) * D_HEAD * 2;                   // bf16 [G][30+T][48]
  static constexpr int UNI_ = RECV_BYTES > SCR_ATT ? (RECV_BYTES > SCR_DW ? RECV_BYTES : SCR_DW)
                                                   : (SCR_ATT > SCR_DW ? SCR_ATT : SCR_DW);
  static constexpr int UNI_BYTES = (UNI_ + 127) / 128 * 128;   // receive buffer of the partial sums / scratch (aliased)
  static constexpr int ROWN_BYTES = RO * D_MODEL * 4;
  static constexpr int P_BYTES = (G * T * TKMAX * 4 + 127) / 128 * 128;
  static constexpr int CONST_FLOATS = 144 + 192 + 96 + 48 + 32 * D_HEAD;   // qkv bias, LN(48) x4, glu bias, dec bias, dw taps+bias
  static constexpr int CONST_BYTES = (CONST_FLOATS * 4 + 127) / 128 * 128;
  static constexpr int FIXED = UNI_BYTES + ROWN_BYTES + P_BYTES + CONST_BYTES + 512;
  static constexpr int SMEM_MAX = 227 * 1024 - 1024;    // minus alignment slack
  static constexpr int NST = (SMEM_MAX - OFF_RING - FIXED) / CL_STAGE_BYTES;
  static constexpr int OFF_UNI = OFF_RING + NST * CL_STAGE_BYTES;
  static constexpr int OFF_ROWN = OFF_UNI + UNI_BYTES;
  static constexpr int OFF_P = OFF_ROWN + ROWN_BYTES;
  static constexpr int OFF_CONST = OFF_P + P_BYTES;
  static constexpr int OFF_BAR = OFF_CONST + CONST_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static_assert(NST >= 3, "weight ring too shallow");
  static_assert(R <= 64, "rows must fit TMEM lane quadrants 0 and 1 / N <= 64");
  static_assert(OFF_H + 2 * KB_BYTES + 16384 <= OFF_UNI, "UMMA reads 128 rows from every k-block base");
  static_assert(RO * D_MODEL * 4 <= 3 * KB_BYTES, "upsample rows are parked in the hidden operand");
  static_assert(R * D_HEAD * 4 <= 6 * KB_BYTES, "reduction slices are parked in the input operand");
  static_assert(4 * NP <= 512 && 144 + 96 * ((G + 1) / 2) <= 512, "TMEM columns");
};

__device__ __forceinline__ unsigned long long cl_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ----------------------------------------------------------------------------------------------- cluster / DSMEM PTX
__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_nclusters() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_map(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st128f(uint32_t raddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void cl_st128u(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cl_st16(uint32_t raddr, uint16_t v) {
  asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(raddr), "h"(v) : "memory");
}
__device__ __forceinline__ void cl_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void cl_arrive_remote(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ bool cl_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cl_wait(uint64_t* bar, uint32_t parity) {
  for (int spins = 0; spins < 400; ++spins)
    if (cl_try_wait(bar, parity)) return;
  printf("tone_b200: cluster barrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {   // exactly one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_workers_mma() { asm volatile("bar.sync 2, 288;" ::: "memory"); }   // workers + MMA warp
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  __nv_bfloat16 b = __float2bfloat16(x);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ uint16_t f16_sat_bits(float x) {
  __half hv = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
  return *reinterpret_cast<uint16_t*>(&hv);
}

// byte offset of 16-byte chunk `ci` (8 bf16 columns 8ci..8ci+7) of row `row` in a K-major SWIZZLE_128B operand made
// of k-block tiles of `kb_bytes`
__device__ __forceinline__ uint32_t sw128_off(int row, int ci, int kb_bytes) {
  return (uint32_t)((ci >> 3) * kb_bytes + (row >> 3) * 1024 + (row & 7) * 128 + (((ci & 7) ^ (row & 7)) << 4));
}
// byte offset of bf16 element (row, col)
__device__ __forceinline__ uint32_t sw128_elem(int row, int col, int kb_bytes) {
  return sw128_off(row, col >> 3, kb_bytes) + ((col & 7) << 1);
}

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(x),
               "r"(y)
               : "memory");
}

// ----------------------------------------------------------------------------------------------- GEMM ops
// One GEMM of the schedule as seen by CTA `c`.  A ring chunk = the weight boxes of one tile and one 64-wide k-block
// (plus, for the cached-context projections, the cached activation rows of that k-block).
struct ClOp {
  const CUtensorMap* m0;     // weight map of box 0 / box 1
  const CUtensorMap* m1;
  int y0, y1;                // first weight row of each box (tile 0)
  int nbox, box_rows;
  int ntile, y_step;         // tiles: weight rows advance by y_step
  int x0, nkb, ksteps_last;  // K start (elements), number of 64-wide k-blocks, k-steps (of 16) in the last block
  int act;                   // activation operand: 0 = normalised input (OFF_A), 1 = local hidden (OFF_H), 2 = cached kv rows (ring)
  int swap;                  // 1: weights are the M = 128 operand and the `np` activation rows are N; 0: rows on M, weights on N
  int np;
  int tmem0, tmem_step;      // accumulator column of tile 0 and step per tile
  const CUtensorMap* amap;   // act == 2: a_ns streams from group stream a_s0, a_rows cached rows each
  int a_rows, a_ns, a_s0;
};

__device__ __forceinline__ ClOp cl_op(const CUtensorMap* m, int y, int box_rows, int ntile, int y_step, int x0, int nkb,
                                      int ksteps_last, int act, int swap, int np, int tmem0, int tmem_step) {
  ClOp o;
  o.m0 = o.m1 = m;
  o.y0 = y;
  o.y1 = 0;
  o.nbox = 1;
  o.box_rows = box_rows;
  o.ntile = ntile;
  o.y_step = y_step;
  o.x0 = x0;
  o.nkb = nkb;
  o.ksteps_last = ksteps_last;
  o.act = act;
  o.swap = swap;
  o.np = np;
  o.tmem0 = tmem0;
  o.tmem_step = tmem_step;
  o.amap = nullptr;
  o.a_rows = o.a_ns = o.a_s0 = 0;
  return o;
}

// feed-forward up projection of CTA c: packed rows [384c, 384c+384) = 3 blocks of [64 gate | 64 value].
// Tiles: gate 0..127, value 0..127 (two 64-row boxes each), gate 128..191, value 128..191 (one box, upper lanes unused).
template <class Role>
__device__ __forceinline__ void cl_ff_up(Role& role, const CUtensorMap* m, int c, int np) {
  ClOp a = cl_op(m, 384 * c, 64, 2, 64, 0, 6, 4, 0, 1, np, 0, np);
  a.nbox = 2;
  a.y1 = 384 * c + 128;
  role.gemm(a);
  role.gemm(cl_op(m, 384 * c + 256, 64, 2, 64, 0, 6, 4, 0, 1, np, 2 * np, np));
}

// The roles walk the same schedule; Role supplies sync_ag / sync_h / gemm / commit.
template <class Cfg, class Role>
__device__ __forceinline__ void cl_schedule(const ClParams& P, int c, Role& role) {
  const CUtensorMap* maps = P.maps;
  for (int l = 0; l < 16; ++l) {
    const CUtensorMap* ml = maps + l * CM_PER_LAYER;
    const bool reduced = l > 6 && l <= 14;
    const int np = reduced ? Cfg::NP2 : Cfg::NP;
    // ---- feed-forward 1
    role.sync_ag();
    cl_ff_up(role, ml + CM_FF1_UP, c, np);
    role.commit();
    role.sync_h();
    role.gemm(cl_op(ml + CM_FF1_DOWN, 0, 128, 3, 128, 192 * c, 3, 4, 1, 1, np, 0, np));
    role.commit();
    // ---- attention: head c.  Accumulator columns: q [0,48) k [48,96) v [96,144); cached rows from 144
    role.sync_ag();
    if (l < 14) {
      const bool rec = (l == 0 || l == 7);
      if (rec) {
        ClOp o = cl_op(ml + CM_QKV, 48 * c, 48, 1, 0, 0, 6, 4, 0, 0, 0, 0, 0);
        o.nbox = 2;
        o.y1 = 384 + 48 * c;
        role.gemm(o);
        role.gemm(cl_op(ml + CM_QKV, 768 + 48 * c, 48, 1, 0, 0, 6, 4, 0, 0, 0, 96, 0));
      } else {
        role.gemm(cl_op(ml + CM_QKV, 48 * c, 48, 1, 0, 0, 6, 4, 0, 0, 0, 96, 0));
      }
    } else {
      ClOp o = cl_op(ml + CM_QKV, 48 * c, 48, 1, 0, 0, 6, 4, 0, 0, 0, 0, 0);   // q | k of the new rows
      o.nbox = 2;
      o.m1 = ml + CM_KV;
      o.y1 = 48 * c;
      role.gemm(o);
      role.gemm(cl_op(ml + CM_KV, 384 + 48 * c, 48, 1, 0, 0, 6, 4, 0, 0, 0, 96, 0));   // v of the new rows
      // k, v of the cached rows: layer 14 = one pass of G x 15 rows, layer 15 = passes of 2 streams x 30 rows
      const int S = (l == 14) ? MHSA_S / 2 : MHSA_S;
      const int per = (l == 14) ? Cfg::G : 2;
      for (int s0 = 0, pass = 0; s0 < Cfg::G; s0 += per, ++pass)
        for (int kv = 0; kv < 2; ++kv) {
          ClOp k = cl_op(ml + CM_KV, 384 * kv + 48 * c, 48, 1, 0, 0, 6, 4, 2, 0, 0, 144 + 96 * pass + 48 * kv, 0);
          k.amap = maps + (l == 14 ? CG_KVC14 : CG_KVC15);
          k.a_rows = S;
          k.a_ns = (Cfg::G - s0 < per) ? Cfg::G - s0 : per;
          k.a_s0 = s0;
          role.gemm(k);
        }
    }
    role.commit();
    role.sync_h();
    role.gemm(cl_op(ml + CM_WO, 0, 128, 3, 128, 48 * c, 1, 3, 1, 1, np, 0, np));
    role.commit();
    // ---- convolution module: channels [48c, 48c+48)
    role.sync_ag();
    role.gemm(cl_op(ml + CM_PW1, 96 * c, 96, 1, 0, 0, 6, 4, 0, 0, 0, 0, 0));
    role.commit();
    role.sync_h();
    role.gemm(cl_op(ml + CM_PW2, 0, 128, 3, 128, 48 * c, 1, 3, 1, 1, np, 0, np));
    role.commit();
    // ---- feed-forward 2
    role.sync_ag();
    cl_ff_up(role, ml + CM_FF2_UP, c, np);
    role.commit();
    role.sync_h();
    role.gemm(cl_op(ml + CM_FF2_DOWN, 0, 128, 3, 128, 192 * c, 3, 4, 1, 1, np, 0, np));
    role.commit();
    if (l == 6) {   // temporal reduction: pointwise 1536 -> 384 over the CTA's 192 depthwise outputs
      role.sync_h();
      role.gemm(cl_op(maps + CG_RED_PW, 0, 128, 3, 128, 192 * c, 3, 4, 1, 1, Cfg::NP2, 0, Cfg::NP2));
      role.commit();
    }
  }
  role.sync_ag();
  role.gemm(cl_op(maps + CG_DEC, 0, 48, 1, 0, 0, 6, 4, 0, 0, 0, 0, 0));
  role.commit();
}

struct ClRing {
  uint8_t* base;
  uint64_t* full;
  uint64_t* empty;
  int nst;
};

template <class Cfg>
struct ClProducer {   // executed by the whole (converged) producer warp; one elected lane issues the TMA loads
  ClRing ring;
  const int* slots;   // slots of this group (global)
  int nvalid;
  int stage = 0;
  uint32_t phase = 0;
  volatile int* progress;   // chunks issued so far (read by the L2 prefetcher)
  int count = 0;
  __device__ __forceinline__ void sync_ag() {}
  __device__ __forceinline__ void sync_h() {}
  __device__ __forceinline__ void commit() {}
  __device__ __forceinline__ void gemm(const ClOp& o) {
    const int wbytes = o.nbox * o.box_rows * 128;
    int abytes = 0, ns = 0;
    if (o.act == 2) {
      ns = o.a_ns;
      if (o.a_s0 + ns > nvalid) ns = nvalid - o.a_s0 > 0 ? nvalid - o.a_s0 : 0;
      abytes = ns * o.a_rows * 128;
    }
    for (int tl = 0; tl < o.ntile; ++tl)
      for (int kb = 0; kb < o.nkb; ++kb) {
        mbar_wait(&ring.empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = ring.base + stage * CL_STAGE_BYTES;
          mbar_expect_tx(&ring.full[stage], wbytes + abytes);
          tma_load_2d(dst, o.m0, &ring.full[stage], o.x0 + kb * 64, o.y0 + tl * o.y_step);
          if (o.nbox == 2)
            tma_load_2d(dst + o.box_rows * 128, o.m1, &ring.full[stage], o.x0 + kb * 64, o.y1 + tl * o.y_step);
          for (int i = 0; i < ns; ++i)
            tma_load_3d(dst + wbytes + i * o.a_rows * 128, o.amap, &ring.full[stage], kb * 64, 0, slots[o.a_s0 + i]);
        }
        __syncwarp();
        if (++stage == ring.nst) {
          stage = 0;
          phase ^= 1;
        }
        ++count;
      }
    if ((threadIdx.x & 31) == 0) *progress = count;
  }
};

// L2 prefetcher: walks the same schedule CL_PREFETCH_AHEAD chunks ahead of the producer and asks the TMA unit to pull
// the weight boxes into L2.  All clusters stream the same weights in near lockstep, so without this every chunk's
// first touch pays DRAM latency in every cluster at once.
template <class Cfg>
struct ClPrefetcher {
  volatile const int* progress;
  int count = 0;
  __device__ __forceinline__ void sync_ag() {}
  __device__ __forceinline__ void sync_h() {}
  __device__ __forceinline__ void commit() {}
  __device__ __forceinline__ void gemm(const ClOp& o) {
    for (int tl = 0; tl < o.ntile; ++tl)
      for (int kb = 0; kb < o.nkb; ++kb) {
        while (count > *progress + CL_PREFETCH_AHEAD) __nanosleep(200);
        if (elect_one()) {
          tma_prefetch_2d(o.m0, o.x0 + kb * 64, o.y0 + tl * o.y_step);
          if (o.nbox == 2) tma_prefetch_2d(o.m1, o.x0 + kb * 64, o.y1 + tl * o.y_step);
        }
        __syncwarp();
        ++count;
      }
  }
};

template <class Cfg>
struct ClMma {   // executed by the whole (converged) MMA warp: control flow and descriptors stay warp-uniform, one
                 // elected lane issues tcgen05.mma / commit (no per-instruction register broadcast)
  ClRing ring;
  uint32_t smem_base;     // shared-window address of the aligned base
  uint32_t tmem_base;
  uint64_t* acc_bar;
  uint64_t* ag_bar;
  int stage = 0;
  uint32_t phase = 0, ag_phase = 0;
  unsigned long long* prof = nullptr;
  int prof_n = 0;
  __device__ __forceinline__ void sync_ag() {   // operand rows of all 8 owners have landed in OFF_A
    cl_wait(ag_bar, ag_phase);
    ag_phase ^= 1;
    fence_proxy_async_all();
    tc_fence_after();
  }
  __device__ __forceinline__ void sync_h() {
    bar_workers_mma();
    fence_proxy_async_all();
    tc_fence_after();
  }
  __device__ __forceinline__ void commit() {
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  __device__ __forceinline__ void gemm(const ClOp& o) {
    const int wrows = o.nbox * o.box_rows;
    const uint32_t idesc = make_idesc_bf16(o.swap ? o.np : wrows);
    unsigned long long t0 = 0;
    if (prof) t0 = cl_gtimer();
    const uint32_t act_base = smem_base + (o.act == 1 ? Cfg::OFF_H : Cfg::OFF_A);
    const uint32_t ring_u32 = smem_u32(ring.base);
    for (int tl = 0; tl < o.ntile; ++tl) {
      const uint32_t d = tmem_base + o.tmem0 + tl * o.tmem_step;
      for (int kb = 0; kb < o.nkb; ++kb) {
        mbar_wait(&ring.full[stage], phase);
        tc_fence_after();
        const uint32_t stg = ring_u32 + stage * CL_STAGE_BYTES;
        const uint32_t act_addr = (o.act == 2) ? stg + wrows * 128 : act_base + kb * Cfg::KB_BYTES;
        const uint64_t dw_ = make_sw128_desc(stg);
        const uint64_t dx = make_sw128_desc(act_addr);
        const uint64_t da = o.swap ? dw_ : dx, db = o.swap ? dx : dw_;
        const int ks = (kb == o.nkb - 1) ? o.ksteps_last : 4;
        if (elect_one()) {
          for (int k = 0; k < ks; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ring.empty[stage]);
        }
        __syncwarp();
        if (++stage == ring.nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (prof && prof_n < 1024 && (threadIdx.x & 31) == 0) {
      unsigned long long* pr = prof + 2048 + 4 * prof_n;
      pr[0] = t0;
      pr[1] = cl_gtimer();
      pr[2] = 0;
      pr[3] = (unsigned long long)(o.ntile * o.nkb);
    }
    ++prof_n;
  }
};

// ----------------------------------------------------------------------------------------------- worker side
template <class Cfg>
struct ClWorker {
  const ClParams& P;
  const ClStep& S;
  uint8_t* sm;              // aligned base
  uint32_t sm_u32;
  uint32_t tmem_base;
  uint64_t *acc_bar, *x_bar, *ag_bar, *z_bar;
  uint32_t acc_phase = 0, x_phase = 0, z_phase = 0;
  int c;                    // CTA rank in the cluster
  int wt, ww, lane, q, h;   // worker thread id, warp, lane, TMEM lane quadrant, half
  int row;                  // TMEM lane of this thread (activation row, or weight row of a swapped tile)
  int grp_stream0;          // first batch position of this group
  int nvalid;               // valid streams in this group
  int slot_of[Cfg::G];
  int len_of[Cfg::G];
  unsigned long long* prof = nullptr;
  int prof_n = 0;
  __device__ __forceinline__ void mark() {
    if (prof && wt == 0 && prof_n < 2048) prof[prof_n++] = cl_gtimer();
  }

  __device__ __forceinline__ ClWorker(const ClParams& p, const ClStep& st) : P(p), S(st) {}

  __device__ __forceinline__ float* r_own() { return reinterpret_cast<float*>(sm + Cfg::OFF_ROWN); }
  __device__ __forceinline__ float* Pbuf() { return reinterpret_cast<float*>(sm + Cfg::OFF_P); }
  __device__ __forceinline__ uint8_t* scr() { return sm + Cfg::OFF_UNI; }
  // staged per-layer constants
  __device__ __forceinline__ float* k_qkvb() { return reinterpret_cast<float*>(sm + Cfg::OFF_CONST); }   // [144] q | k | v bias of head c
  __device__ __forceinline__ float* k_ln() { return k_qkvb() + 144; }                                     // qln_w, qln_b, kln_w, kln_b [48] each
  __device__ __forceinline__ float* k_glub() { return k_ln() + 192; }                                     // [96] a | b bias
  __device__ __forceinline__ float* k_decb() { return k_glub() + 96; }                                    // [48]
  __device__ __forceinline__ float* k_dw() { return k_decb() + 48; }                                      // [31][48] taps, [48] bias

  __device__ __forceinline__ void wait_acc() {
    mbar_wait(acc_bar, acc_phase);
    acc_phase ^= 1;
    tc_fence_after();
  }
  // all workers have finished their (remote) stores: one thread signals every CTA of the cluster
  // Ordering: every worker's (remote) stores happen-before the bar.sync; the elected thread's arrive has release
  // semantics at cluster scope, which is cumulative over what it observed through the CTA barrier.  (An explicit
  // fence.acq_rel.cluster per thread costs ~3 us per signal on B200 - measured - and is not needed.)
  __device__ __forceinline__ void signal_all(uint64_t* bar) {
    bar_workers();
    if (wt < CL_CTAS) cl_arrive_remote(cl_map(smem_u32(bar), wt));
  }
  __device__ __forceinline__ void wait_x() {
    cl_wait(x_bar, x_phase);
    x_phase ^= 1;
  }
  __device__ __forceinline__ void wait_z() {
    cl_wait(z_bar, z_phase);
    z_phase ^= 1;
  }
  // hidden slice written: hand over to the MMA warp
  __device__ __forceinline__ void release_h() {
    fence_proxy_async_all();
    tc_fence_before();
    bar_workers_mma();
  }
  __device__ __forceinline__ uint32_t tmem_row() const { return tmem_base + (static_cast<uint32_t>(q * 32) << 16); }

  // ---- per-layer constants -> shared memory (read later by the attention / GLU / depthwise epilogues)
  __device__ __forceinline__ void stage_consts(int l) {
    const ClLayer& L = P.L[l];
    const bool rec = (l == 0 || l == 7 || l >= 14);
    for (int i = wt; i < Cfg::CONST_FLOATS; i += CL_WORKERS) {
      float v = 0.f;
      if (i < 144) {
        const int part = i / 48, d = i - part * 48;
        if (l < 14) {
          if (rec) v = __ldg(L.qkv_b + part * 384 + 48 * c + d);
          else if (part == 2) v = __ldg(L.qkv_b + 48 * c + d);
        } else {
          v = part == 0 ? __ldg(L.q_b + 48 * c + d) : __ldg(L.kv_b + (part - 1) * 384 + 48 * c + d);
        }
      } else if (i < 336) {
        if (rec) {
          const int j = i - 144, part = j / 48, d = j - part * 48;
          const float* src = part == 0 ? L.qln_w : (part == 1 ? L.qln_b : (part == 2 ? L.kln_w : L.kln_b));
          v = __ldg(src + d);
        }
      } else if (i < 432) {
        v = __ldg(L.pw1_b + 96 * c + (i - 336));
      } else if (i < 480) {
        const int d = i - 432;
        v = d < 35 ? __ldg(P.dec_b + d) : 0.f;
      } else {
        const int j = i - 480, jj = j / 48, d = j - jj * 48;
        v = jj < 31 ? __ldg(L.dw_w + jj * D_MODEL + 48 * c + d) : __ldg(L.dw_b + 48 * c + d);
      }
      reinterpret_cast<float*>(sm + Cfg::OFF_CONST)[i] = v;
    }
  }

  // ---- SwiGLU epilogue of the swapped up projection.  Tiles at columns [0,np) gate 0..127, [np,2np) value 0..127,
  // [2np,3np) gate 128..191, [3np,4np) value 128..191 (lanes 0..63).  Thread = hidden feature, columns = rows.
  // Half h of each lane quadrant takes every other 16-row unit.
  __device__ __forceinline__ void ep_swiglu(const float* __restrict__ bias, int Rl, int np) {
    const int L = row;                               // lane within the tile
    // packed bias: block t = f / 64 holds [64 gate | 64 value]
    const float bg0 = __ldg(bias + 128 * (L >> 6) + (L & 63)), bv0 = __ldg(bias + 128 * (L >> 6) + 64 + (L & 63));
    float bg1 = 0.f, bv1 = 0.f;
    if (L < 64) {
      bg1 = __ldg(bias + 256 + L);
      bv1 = __ldg(bias + 320 + L);
    }
    const uint32_t trow = tmem_row();
    const uint32_t hb = sm_u32 + Cfg::OFF_H;
    for (int u = h; u * 16 < np; u += 2) {
      if (u * 16 >= Rl) break;
      uint32_t g0[16], v0[16], g1[16], v1[16];
      tmem_ld16_async(trow + u * 16, g0);
      tmem_ld16_async(trow + np + u * 16, v0);
      if (q < 2) {                                   // warp-uniform
        tmem_ld16_async(trow + 2 * np + u * 16, g1);
        tmem_ld16_async(trow + 3 * np + u * 16, v1);
      }
      tmem_ld_wait();
      tmem_regs_ready16(g0);
      tmem_regs_ready16(v0);
      if (q < 2) {
        tmem_regs_ready16(g1);
        tmem_regs_ready16(v1);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int r = u * 16 + k;
        if (r < Rl) {
          const float o = silu_f(__uint_as_float(g0[k]) + bg0) * (__uint_as_float(v0[k]) + bv0);
          sts16(hb + sw128_elem(r, L, Cfg::KB_BYTES), bf16_bits(o));
          if (q < 2) {
            const float o1 = silu_f(__uint_as_float(g1[k]) + bg1) * (__uint_as_float(v1[k]) + bv1);
            sts16(hb + sw128_elem(r, 128 + L, Cfg::KB_BYTES), bf16_bits(o1));
          }
        }
      }
    }
  }

  // ---- swapped row-parallel GEMM done: tiles m = 0..2 at columns m*np hold features 128m + lane for all rows.
  // Each value goes, as fp16, to the owner of its row: recv[(sender * RO + row / 8)][feature].
  __device__ __forceinline__ void ep_send(int Rl, int np) {
    const uint32_t trow = tmem_row();
    uint32_t rb[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) rb[p] = cl_map(sm_u32 + Cfg::OFF_UNI + (c * Cfg::RO * CL_RECV_LD + row) * 2, p);
    for (int u = h; u * 16 < np; u += 2) {
      if (u * 16 >= Rl) break;
      uint32_t d0[16], d1[16], d2[16];
      tmem_ld16_async(trow + u * 16, d0);
      tmem_ld16_async(trow + np + u * 16, d1);
      tmem_ld16_async(trow + 2 * np + u * 16, d2);
      tmem_ld_wait();
      tmem_regs_ready16(d0);
      tmem_regs_ready16(d1);
      tmem_regs_ready16(d2);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int r = u * 16 + k;
        if (r < Rl) {
          const uint32_t dst = rb[k & 7] + ((2 * u + (k >> 3)) * CL_RECV_LD) * 2;   // li = r / 8
          cl_st16(dst, f16_sat_bits(__uint_as_float(d0[k])));
          cl_st16(dst + 256, f16_sat_bits(__uint_as_float(d1[k])));
          cl_st16(dst + 512, f16_sat_bits(__uint_as_float(d2[k])));
        }
      }
    }
    tc_fence_before();
    mark();
    signal_all(x_bar);
  }

  // ---- row owner: x = [r +] scale * (sum of 8 partials + bias); optional norm_out in place; operand row n = gain *
  // x / rms (or bf16(x) for the decoder) broadcast to the 8 CTAs; optional copy of n into the stream's kv rows.
  struct OwnerArgs {
    const float* bias;
    float scale;
    bool add_resid;       // r_own += ... (else r_own = ...)
    const float* g_out;   // nullable: x = g_out * x / rms(x) written back as the new residual
    const float* g_next;  // nullable: gain of the following RMSNorm
    bool norm_next;       // false: n = bf16(x) (decoder input)
    bf16* kv;             // nullable: per-slot [KV_ROWS_MAX][384] pool of the layer whose input rows are cached
    int kv_row_off;       // S
    int Tl;               // frames per stream at this rate
    int tap;              // tap index or -1
    bool stash_res;       // S.res rows = x (layer 6 output kept for the upsampling residual)
    bool from_partials;   // false: x = r_own as is (start of the kernel / after the upsample add)
    bool broadcast;       // false: keep x only (layers 6 and 14 continue with an exchange first)
  };
  struct OwnerConsts {
    float b[16], go[16], gn[16];
  };
  // constants of this lane's 16 columns, fetched BEFORE the wait for the partial sums
  __device__ __forceinline__ void owner_prefetch(const OwnerArgs& a, OwnerConsts& k) {
    const int c0 = lane * 8, c1 = 192 + lane * 8;
#pragma unroll
    for (int i = 0; i < 16; ++i) k.b[i] = k.go[i] = k.gn[i] = 0.f;   // lanes 24..31 carry zeros through the row sums
    if (lane < 24) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const int col = (i < 8 ? c0 : c1 - 8) + i;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 vb = (a.from_partials && a.bias) ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : z;
        const float4 vo = a.g_out ? __ldg(reinterpret_cast<const float4*>(a.g_out + col)) : one;
        const float4 vn = a.g_next ? __ldg(reinterpret_cast<const float4*>(a.g_next + col)) : one;
        k.b[i] = vb.x; k.b[i + 1] = vb.y; k.b[i + 2] = vb.z; k.b[i + 3] = vb.w;
        k.go[i] = vo.x; k.go[i + 1] = vo.y; k.go[i + 2] = vo.z; k.go[i + 3] = vo.w;
        k.gn[i] = vn.x; k.gn[i + 1] = vn.y; k.gn[i + 2] = vn.z; k.gn[i + 3] = vn.w;
      }
    }
  }
  __device__ __forceinline__ void owner_rows(const OwnerArgs& a, const OwnerConsts& kc) {
    const int Rl = Cfg::G * a.Tl;
    const __half* rcv = reinterpret_cast<const __half*>(scr());
    for (int li = ww; li < Cfg::RO; li += 8) {
      const int j = li * 8 + c;
      if (j >= Rl) continue;                                 // warp-uniform
      float x[16];
      float* rr = r_own() + li * D_MODEL;
      const bool act = lane < 24;
      const int c0 = lane * 8, c1 = 192 + lane * 8;
      if (act) {
        if (a.from_partials) {
          float s[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) s[k] = kc.b[k];
#pragma unroll
          for (int p = 0; p < CL_CTAS; ++p) {                // fixed order: deterministic
            const __half* pr = rcv + (p * Cfg::RO + li) * CL_RECV_LD;
            const uint4 v0 = *reinterpret_cast<const uint4*>(pr + c0), v1 = *reinterpret_cast<const uint4*>(pr + c1);
            const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
              s[2 * k] += f.x;
              s[2 * k + 1] += f.y;
            }
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float base = a.add_resid ? rr[(k < 8 ? c0 : c1 - 8) + k] : 0.f;
            x[k] = base + a.scale * s[k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) x[k] = rr[(k < 8 ? c0 : c1 - 8) + k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = 0.f;
      }
      if (a.g_out) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) ss += x[k] * x[k];
        ss = warp_sum(ss);
        const float inv = 1.0f / (sqrtf(ss) * 0.05103103630798288f + 1e-8f);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = kc.go[k] * (x[k] * inv);
      }
      if (act && (a.from_partials || a.g_out)) {
#pragma unroll
        for (int k = 0; k < 16; k += 4)
          *reinterpret_cast<float4*>(rr + (k < 8 ? c0 : c1 - 8) + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
      }
      const int s_ = j / a.Tl, t_ = j - s_ * a.Tl;
      if (act && a.stash_res && s_ < nvalid) {
        float* rs = S.res + ((size_t)grp_stream0 * a.Tl + j) * D_MODEL;
#pragma unroll
        for (int k = 0; k < 16; k += 4)
          *reinterpret_cast<float4*>(rs + (k < 8 ? c0 : c1 - 8) + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
      }
      if (a.tap >= 0 && S.taps && act && s_ < nvalid) {
        float* tp = S.taps + (size_t)a.tap * S.tap_stride + ((size_t)grp_stream0 * a.Tl + j) * D_MODEL;
#pragma unroll
        for (int k = 0; k < 16; ++k) tp[(k < 8 ? c0 : c1 - 8) + k] = x[k];
      }
      if (!a.broadcast) continue;
      if (a.norm_next) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) ss += x[k] * x[k];
        ss = warp_sum(ss);
        const float inv = 1.0f / (sqrtf(ss) * 0.05103103630798288f + 1e-8f);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = kc.gn[k] * (x[k] * inv);
      }
      if (act) {
        const uint4 v0 = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
        const uint4 v1 = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
        const uint32_t o0 = sm_u32 + Cfg::OFF_A + sw128_off(j, lane, Cfg::KB_BYTES);
        const uint32_t o1 = sm_u32 + Cfg::OFF_A + sw128_off(j, 24 + lane, Cfg::KB_BYTES);
#pragma unroll
        for (int p = 0; p < CL_CTAS; ++p) {
          cl_st128u(cl_map(o0, p), v0);
          cl_st128u(cl_map(o1, p), v1);
        }
        if (a.kv && s_ < nvalid) {
          bf16* kr = a.kv + ((size_t)slot_of[s_] * KV_ROWS_MAX + a.kv_row_off + t_) * D_MODEL;
          *reinterpret_cast<uint4*>(kr + c0) = v0;
          *reinterpret_cast<uint4*>(kr + c1) = v1;
        }
      }
    }
    mark();
    if (a.broadcast) {
      fence_proxy_async_all();
      signal_all(ag_bar);
    }
  }

  // ---- attention of head c over this group's streams.  TMEM (rows on lanes): q [0,48) k [48,96) v [96,144) of the
  // new rows; layers 14 / 15 additionally k, v of the cached rows at columns 144 + 96 * pass (+48).
  __device__ __forceinline__ void ln_rope(float* x, const float* w, const float* b, int pos, float scale) {
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) mean += x[i];
    mean *= (1.0f / D_HEAD);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) {
      const float d = x[i] - mean;
      var = fmaf(d, d, var);
    }
    const float inv = rsqrtf(var * (1.0f / D_HEAD) + 1e-5f);
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) x[i] = (x[i] - mean) * inv * w[i] + b[i];
    const float* cs = P.rope_cos + (pos + MHSA_S) * 16;
    const float* sn = P.rope_sin + (pos + MHSA_S) * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 cc = __ldg(reinterpret_cast<const float4*>(cs + i)), s4 = __ldg(reinterpret_cast<const float4*>(sn + i));
      const float ca[4] = {cc.x, cc.y, cc.z, cc.w}, sa[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x1 = x[i + k], x2 = x[i + k + 16];
        x[i + k] = x1 * ca[k] - x2 * sa[k];
        x[i + k + 16] = x2 * ca[k] + x1 * sa[k];
      }
    }
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) x[i] *= scale;
  }
  __device__ __forceinline__ void load48(uint32_t taddr, const float* bias_s, float* x) {
    uint32_t r[3][16];
#pragma unroll
    for (int i = 0; i < 3; ++i) tmem_ld16_async(taddr + 16 * i, r[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      tmem_regs_ready16(r[i]);
#pragma unroll
      for (int k = 0; k < 16; ++k) x[16 * i + k] = __uint_as_float(r[i][k]) + bias_s[16 * i + k];
    }
  }
  // key row held in registers -> scores against the T queries of its stream
  __device__ __forceinline__ void scores_for_key(const float* kx, int s_, int key, int Tl, int Tk, int off) {
    const float* qs = reinterpret_cast<const float*>(scr());
    float* Pb = Pbuf();
    const bool masked = key < off;
    for (int t = 0; t < Tl; ++t) {
      const float* qr = qs + (s_ * Tl + t) * Cfg::QS_LD;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D_HEAD; d += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qr + d);
        acc = fmaf(qv.x, kx[d], acc);
        acc = fmaf(qv.y, kx[d + 1], acc);
        acc = fmaf(qv.z, kx[d + 2], acc);
        acc = fmaf(qv.w, kx[d + 3], acc);
      }
      Pb[(s_ * Tl + t) * Tk + key] = masked ? -10000.0f : acc;
    }
  }
  __device__ __forceinline__ void attention(int l, int Tl) {
    const bool rec = (l == 0 || l == 7 || l >= 14);
    const int Sc = (l == 14) ? MHSA_S / 2 : (l == 15 ? MHSA_S : 0);
    const int Tk = Sc + Tl, Rl = Cfg::G * Tl;
    float* qs = reinterpret_cast<float*>(scr());
    float* vs = qs + Cfg::R * Cfg::QS_LD;          // [G][Tk][48]
    const uint32_t trow = tmem_row();
    const bool have_rows = q * 32 < Rl;            // warp-uniform
    const int s_ = row / Tl, t_ = row - s_ * Tl;
    const float* qb = k_qkvb();
    const float* kb = qb + 48;
    const float* vb = qb + 96;
    const float* lnw = k_ln();
    float kx[D_HEAD];
    bool have_k = false;
    int k_s = 0, k_key = 0;
    // phase 1: q rows -> smem (half 0); k rows -> registers, v rows -> smem (half 1)
    if (rec) {
      if (h == 0) {
        if (have_rows) {
          float x[D_HEAD];
          load48(trow + 0, qb, x);
          if (row < Rl) {
            ln_rope(x, lnw, lnw + 48, t_, 0.14433756729740643f);
#pragma unroll
            for (int d = 0; d < D_HEAD; d += 4)
              *reinterpret_cast<float4*>(qs + row * Cfg::QS_LD + d) = make_float4(x[d], x[d + 1], x[d + 2], x[d + 3]);
          }
        }
      } else {
        if (have_rows) {
          float v[D_HEAD];
          load48(trow + 48, kb, kx);
          load48(trow + 96, vb, v);
          if (row < Rl) {
            ln_rope(kx, lnw + 96, lnw + 144, t_, 1.0f);
            have_k = true;
            k_s = s_;
            k_key = Sc + t_;
#pragma unroll
            for (int d = 0; d < D_HEAD; d += 4)
              *reinterpret_cast<float4*>(vs + (s_ * Tk + Sc + t_) * D_HEAD + d) = make_float4(v[d], v[d + 1], v[d + 2], v[d + 3]);
          }
        }
      }
    } else if (h == 0 && have_rows) {
      float v[D_HEAD];
      load48(trow + 96, vb, v);
      if (row < Rl) {
#pragma unroll
        for (int d = 0; d < D_HEAD; d += 4)
          *reinterpret_cast<float4*>(vs + (s_ * Tk + t_) * D_HEAD + d) = make_float4(v[d], v[d + 1], v[d + 2], v[d + 3]);
      }
    }
    // cached rows (layers 14, 15): pass p holds `per` streams x Sc rows in TMEM lanes [0, per*Sc)
    float kx2[D_HEAD];
    bool have_k2 = false;
    int k2_s = 0, k2_key = 0;
    float kx3[D_HEAD];
    bool have_k3 = false;
    int k3_s = 0, k3_key = 0;
    if (Sc > 0) {
      const int per = (l == 14) ? Cfg::G : 2;
      const int npass = (Cfg::G + per - 1) / per;
      // half 0 takes passes 0 and 2, half 1 pass 1 (layer 14 has one pass: half 0)
      for (int pass = h; pass < npass; pass += 2) {
        const int rows_p = per * Sc;
        if (q * 32 < rows_p) {
          float v[D_HEAD];
          float* kdst = (pass < 2) ? kx2 : kx3;
          const uint32_t tc = trow + 144 + 96 * pass;
          load48(tc, kb, kdst);
          load48(tc + 48, vb, v);
          if (row < rows_p) {
            const int sp = row / Sc, i_c = row - sp * Sc;
            const int ss_ = pass * per + sp;
            if (ss_ < Cfg::G) {
              ln_rope(kdst, lnw + 96, lnw + 144, i_c - Sc, 1.0f);
              if (pass < 2) {
                have_k2 = true;
                k2_s = ss_;
                k2_key = i_c;
              } else {
                have_k3 = true;
                k3_s = ss_;
                k3_key = i_c;
              }
#pragma unroll
              for (int d = 0; d < D_HEAD; d += 4)
                *reinterpret_cast<float4*>(vs + (ss_ * Tk + i_c) * D_HEAD + d) = make_float4(v[d], v[d + 1], v[d + 2], v[d + 3]);
            }
          }
        }
      }
    }
    bar_workers();
    if (rec) {
      // phase 2: scores
      if (have_k) {
        const int off = (l == 15) ? MHSA_S - len_of[k_s] : (l == 14 ? (MHSA_S - len_of[k_s]) / 2 : 0);
        scores_for_key(kx, k_s, k_key, Tl, Tk, off);
      }
      if (have_k2) {
        const int off = (l == 15) ? MHSA_S - len_of[k2_s] : (MHSA_S - len_of[k2_s]) / 2;
        scores_for_key(kx2, k2_s, k2_key, Tl, Tk, off);
      }
      if (have_k3) {
        const int off = (l == 15) ? MHSA_S - len_of[k3_s] : (MHSA_S - len_of[k3_s]) / 2;
        scores_for_key(kx3, k3_s, k3_key, Tl, Tk, off);
      }
      bar_workers();
      // phase 3: softmax, one warp per (stream, query) row
      float* Pb = Pbuf();
      for (int r = ww; r < Rl; r += 8) {
        const int rs = r / Tl;
        const int off = (l == 15) ? MHSA_S - len_of[rs] : (l == 14 ? (MHSA_S - len_of[rs]) / 2 : 0);
        float* pr = Pb + r * Tk;
        const float a0 = lane < Tk ? pr[lane] : -INFINITY;
        const float a1 = lane + 32 < Tk ? pr[lane + 32] : -INFINITY;
        const float mx = warp_max(fmaxf(a0, a1));
        const float e0 = lane < Tk ? expf(a0 - mx) : 0.f;
        const float e1 = lane + 32 < Tk ? expf(a1 - mx) : 0.f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        if (lane < Tk) pr[lane] = (lane < off) ? 0.f : e0 * inv;
        if (lane + 32 < Tk) pr[lane + 32] = (lane + 32 < off) ? 0.f : e1 * inv;
      }
      bar_workers();
    }
    // phase 4: ctx = P V -> bf16 into the first k-block of the hidden operand (columns 0..47)
    {
      const float* Pb = Pbuf();
      for (int unit = wt; unit < Rl * 6; unit += CL_WORKERS) {
        const int r = unit / 6, u = unit - r * 6;
        const int rs = r / Tl;
        const float* pr = Pb + r * Tk;
        const float* vr = vs + rs * Tk * D_HEAD + u * 8;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int j = 0; j < Tk; ++j) {
          const float p = pr[j];
          const float4 v0 = *reinterpret_cast<const float4*>(vr + j * D_HEAD);
          const float4 v1 = *reinterpret_cast<const float4*>(vr + j * D_HEAD + 4);
          acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v0.z, acc[2]); acc[3] = fmaf(p, v0.w, acc[3]);
          acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]); acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
        }
        sts128u(sm_u32 + Cfg::OFF_H + sw128_off(r, u, Cfg::KB_BYTES),
                make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7])));
      }
    }
  }

  // ---- convolution module middle: GLU epilogue -> causal depthwise conv k=31 (+ folded BN) -> SiLU, cache roll
  __device__ __forceinline__ void conv_prefetch(int l, uint4 (&pre)[3]) {
    // cached rows [30][48 channels of this CTA] of every stream: G*30 rows x 6 pieces of 16 B
    const int n = Cfg::G * CONV_S * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = wt + k * CL_WORKERS;
      pre[k] = make_uint4(0, 0, 0, 0);
      if (i < n) {
        const int s_ = i / (CONV_S * 6), rem = i - s_ * CONV_S * 6, rrow = rem / 6, pc = rem - rrow * 6;
        if (s_ < nvalid) {
          const bf16* src = P.conv + (((size_t)slot_of[s_] * 16 + l) * CONV_S + rrow) * D_MODEL + 48 * c + pc * 8;
          pre[k] = *reinterpret_cast<const uint4*>(src);
        }
      }
    }
  }
  __device__ __forceinline__ void conv_module(int l, int Tl, const uint4 (&pre)[3]) {
    static_assert(Cfg::G * CONV_S * 6 <= 4 * CL_WORKERS, "cache prefetch registers");
    const int Rl = Cfg::G * Tl, rows_s = CONV_S + Tl;
    bf16* xs = reinterpret_cast<bf16*>(scr());       // [G][30 + Tl][48]
    {
      const int n = Cfg::G * CONV_S * 6;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = wt + k * CL_WORKERS;
        if (i < n) {
          const int s_ = i / (CONV_S * 6), rem = i - s_ * CONV_S * 6, rrow = rem / 6, pc = rem - rrow * 6;
          *reinterpret_cast<uint4*>(xs + (s_ * rows_s + rrow) * D_HEAD + pc * 8) = pre[k];
        }
      }
      // G = 5: 900 pieces > 768 prefetched; the remainder is fetched here
      for (int i = wt + 3 * CL_WORKERS; i < n; i += CL_WORKERS) {
        const int s_ = i / (CONV_S * 6), rem = i - s_ * CONV_S * 6, rrow = rem / 6, pc = rem - rrow * 6;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (s_ < nvalid)
          v = *reinterpret_cast<const uint4*>(P.conv + (((size_t)slot_of[s_] * 16 + l) * CONV_S + rrow) * D_MODEL + 48 * c + pc * 8);
        *reinterpret_cast<uint4*>(xs + (s_ * rows_s + rrow) * D_HEAD + pc * 8) = v;
      }
    }
    // GLU: columns [a 48 | b 48]; half h takes channels [24h, 24h+24)
    if (q * 32 < Rl) {
      const uint32_t trow = tmem_row();
      uint32_t ra[2][16], rb[2][16];
      // 24 channels = 16 + 8: load 16-wide pieces (the second piece overlaps into the other half; only 8 used)
      tmem_ld16_async(trow + 24 * h, ra[0]);
      tmem_ld16_async(trow + 24 * h + 16, ra[1]);
      tmem_ld16_async(trow + 48 + 24 * h, rb[0]);
      tmem_ld16_async(trow + 48 + 24 * h + 16, rb[1]);
      tmem_ld_wait();
      tmem_regs_ready16(ra[0]);
      tmem_regs_ready16(ra[1]);
      tmem_regs_ready16(rb[0]);
      tmem_regs_ready16(rb[1]);
      if (row < Rl) {
        const int s_ = row / Tl, t_ = row - s_ * Tl;
        const float* ba = k_glub() + 24 * h;
        const float* bb = k_glub() + 48 + 24 * h;
        uint32_t pk[12];
#pragma unroll
        for (int k = 0; k < 24; k += 2) {
          const float a0 = __uint_as_float(k < 16 ? ra[0][k] : ra[1][k - 16]) + ba[k];
          const float a1 = __uint_as_float(k + 1 < 16 ? ra[0][k + 1] : ra[1][k + 1 - 16]) + ba[k + 1];
          const float b0 = __uint_as_float(k < 16 ? rb[0][k] : rb[1][k - 16]) + bb[k];
          const float b1 = __uint_as_float(k + 1 < 16 ? rb[0][k + 1] : rb[1][k + 1 - 16]) + bb[k + 1];
          pk[k >> 1] = pack_bf16x2(a0 * sigmoid_f(b0), a1 * sigmoid_f(b1));
        }
        uint4* dst = reinterpret_cast<uint4*>(xs + (s_ * rows_s + CONV_S + t_) * D_HEAD + 24 * h);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
      }
    }
    bar_workers();
    // depthwise: unit = (row, 8 channels); taps and bias staged in shared memory
    const float* tw = k_dw();
    for (int unit = wt; unit < Rl * 6; unit += CL_WORKERS) {
      const int r = unit / 6, u = unit - r * 6;
      const int s_ = r / Tl, t_ = r - s_ * Tl;
      float acc[8];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(tw + 31 * D_HEAD + u * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(tw + 31 * D_HEAD + u * 8 + 4);
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
      }
      const bf16* xr = xs + (s_ * rows_s + t_) * D_HEAD + u * 8;
#pragma unroll
      for (int jj = 0; jj <= CONV_S; ++jj) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + jj * D_HEAD);
        const float4 w0 = *reinterpret_cast<const float4*>(tw + jj * D_HEAD + u * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(tw + jj * D_HEAD + u * 8 + 4);
        const float2 x0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.x));
        const float2 x1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.y));
        const float2 x2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.z));
        const float2 x3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.w));
        acc[0] = fmaf(w0.x, x0.x, acc[0]); acc[1] = fmaf(w0.y, x0.y, acc[1]); acc[2] = fmaf(w0.z, x1.x, acc[2]); acc[3] = fmaf(w0.w, x1.y, acc[3]);
        acc[4] = fmaf(w1.x, x2.x, acc[4]); acc[5] = fmaf(w1.y, x2.y, acc[5]); acc[6] = fmaf(w1.z, x3.x, acc[6]); acc[7] = fmaf(w1.w, x3.y, acc[7]);
      }
      sts128u(sm_u32 + Cfg::OFF_H + sw128_off(r, u, Cfg::KB_BYTES),
              make_uint4(pack_bf16x2(silu_f(acc[0]), silu_f(acc[1])), pack_bf16x2(silu_f(acc[2]), silu_f(acc[3])),
                         pack_bf16x2(silu_f(acc[4]), silu_f(acc[5])), pack_bf16x2(silu_f(acc[6]), silu_f(acc[7]))));
    }
    // new cache = last 30 rows of [cache | new]
    {
      const int n = Cfg::G * CONV_S * 6;
      for (int i = wt; i < n; i += CL_WORKERS) {
        const int s_ = i / (CONV_S * 6), rem = i - s_ * CONV_S * 6, rrow = rem / 6, pc = rem - rrow * 6;
        if (s_ < nvalid) {
          bf16* dst = P.conv + (((size_t)slot_of[s_] * 16 + l) * CONV_S + rrow) * D_MODEL + 48 * c + pc * 8;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(xs + (s_ * rows_s + Tl + rrow) * D_HEAD + pc * 8);
        }
      }
    }
  }

  // ---- after layer 6: owners scatter their rows by column slice (parked in the dead input-operand buffer), every CTA
  // runs the stride-2 depthwise (x4) conv of its 48 channels into the hidden operand (192 columns); the pointwise conv
  // follows as a row-parallel GEMM.
  __device__ __forceinline__ void reduction_exchange() {
    constexpr int T = Cfg::T, T2 = Cfg::T2, R = Cfg::R;
    for (int li = ww; li < Cfg::RO; li += 8) {
      const int j = li * 8 + c;
      if (j >= R) continue;
      const float* rr = r_own() + li * D_MODEL;
      // lane handles 12 consecutive columns = 3 float4; destination CTA = col / 48
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = lane * 12 + k * 4;
        const float4 v = *reinterpret_cast<const float4*>(rr + col);
        const int pc = col / 48, cc = col - pc * 48;
        cl_st128f(cl_map(sm_u32 + Cfg::OFF_A + (j * D_HEAD + cc) * 4, pc), v);
      }
    }
    signal_all(x_bar);
    wait_x();
    const float* rt = reinterpret_cast<const float*>(sm + Cfg::OFF_A);
    for (int unit = wt; unit < Cfg::G * T2 * 6; unit += CL_WORKERS) {
      const int r2 = unit / 6, u = unit - r2 * 6;
      const int s_ = r2 / T2, t2 = r2 - s_ * T2;
      const int ch0 = 48 * c + u * 8;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = ch0 + i;
        float v0;
        if (t2 == 0) v0 = (s_ < nvalid) ? P.red[(size_t)slot_of[s_] * D_MODEL + ch] : 0.f;
        else v0 = rt[(s_ * T + 2 * t2 - 1) * D_HEAD + u * 8 + i];
        const float v1 = rt[(s_ * T + 2 * t2) * D_HEAD + u * 8 + i];
        const float v2 = rt[(s_ * T + 2 * t2 + 1) * D_HEAD + u * 8 + i];
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = __ldg(P.red_dw_b + ch * 4 + k) + __ldg(P.red_dw_w + ch * 12 + k * 3) * v0 +
                 __ldg(P.red_dw_w + ch * 12 + k * 3 + 1) * v1 + __ldg(P.red_dw_w + ch * 12 + k * 3 + 2) * v2;
        pk[2 * i] = pack_bf16x2(o[0], o[1]);
        pk[2 * i + 1] = pack_bf16x2(o[2], o[3]);
      }
      // local hidden columns 32u .. 32u+31 = chunks 4u .. 4u+3
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts128u(sm_u32 + Cfg::OFF_H + sw128_off(r2, 4 * u + k, Cfg::KB_BYTES),
                make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
    }
    bar_workers();   // every read of the old carried column is done
    for (int i = wt; i < Cfg::G * D_HEAD; i += CL_WORKERS) {
      const int s_ = i / D_HEAD, cc = i - s_ * D_HEAD;
      if (s_ < nvalid) P.red[(size_t)slot_of[s_] * D_MODEL + 48 * c + cc] = rt[(s_ * T + T - 1) * D_HEAD + cc];
    }
  }

  // ---- after layer 14: reduced-rate owners send their rows to the owners of the two full-rate rows they feed
  // (parked in the dead hidden-operand buffer); r = upsampled + layer-6 output
  __device__ __forceinline__ void upsample_exchange() {
    constexpr int T = Cfg::T, T2 = Cfg::T2;
    for (int li = ww; li < Cfg::RO; li += 8) {
      const int j2 = li * 8 + c;
      if (j2 >= Cfg::R2) continue;
      const int s_ = j2 / T2, t2 = j2 - s_ * T2;
      const float* rr = r_own() + li * D_MODEL;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = lane * 12 + k * 4;
        const float4 v = *reinterpret_cast<const float4*>(rr + col);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int j = s_ * T + 2 * t2 + d;
          cl_st128f(cl_map(sm_u32 + Cfg::OFF_H + ((j >> 3) * D_MODEL + col) * 4, j & 7), v);
        }
      }
    }
    signal_all(x_bar);
    wait_x();
    const float* up = reinterpret_cast<const float*>(sm + Cfg::OFF_H);
    for (int li = ww; li < Cfg::RO; li += 8) {
      const int j = li * 8 + c;
      if (j >= Cfg::R) continue;
      const int s_ = j / T, t_ = j - s_ * T;
      float* rr = r_own() + li * D_MODEL;
      const float* rs = S.res + ((size_t)grp_stream0 * T + j) * D_MODEL;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = lane * 12 + k * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s_ < nvalid) v = *reinterpret_cast<const float4*>(rs + col);
        if (t_ < 2 * T2) {
          const float4 u = *reinterpret_cast<const float4*>(up + li * D_MODEL + col);
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        *reinterpret_cast<float4*>(rr + col) = v;
      }
    }
    __syncwarp();
  }

  // ---- decoder epilogue: log-softmax over 35 classes + first-max argmax; row j is written by CTA j % 8
  __device__ __forceinline__ void decoder() {
    constexpr int T = Cfg::T, R = Cfg::R;
    if (h != 0 || q * 32 >= R) return;
    uint32_t r[3][16];
    const uint32_t trow = tmem_row();
#pragma unroll
    for (int i = 0; i < 3; ++i) tmem_ld16_async(trow + 16 * i, r[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 3; ++i) tmem_regs_ready16(r[i]);
    if (row >= R || (row & 7) != c) return;
    const int s_ = row / T;
    if (s_ >= nvalid) return;
    const float* db = k_decb();
    float lg[35];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int i = 0; i < 35; ++i) {
      lg[i] = __uint_as_float(r[i >> 4][i & 15]) + db[i];
      if (lg[i] > mx) {   // strict >: first maximum (numpy argmax, tone/decoder.py:57)
        mx = lg[i];
        am = i;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 35; ++i) sum += expf(lg[i] - mx);
    const float lse = mx + logf(sum);
    const size_t orow = (size_t)grp_stream0 * T + row;
    float* out = S.logprobs + orow * 35;
#pragma unroll
    for (int i = 0; i < 35; ++i) out[i] = lg[i] - lse;
    if (S.tokens) S.tokens[orow] = am;
    if (S.aux) *reinterpret_cast<float2*>(S.aux + orow * 2) = make_float2(lg[33] - lse, lg[34] - lse);
  }

  // partial sums -> owner -> operand rows, the common tail of every sub-block
  __device__ __forceinline__ void reduce_tail(const OwnerArgs& a, int np, bool guard_scratch) {
    OwnerConsts kc;
    const int Rl = Cfg::G * a.Tl;
    wait_acc();
    mark();
    owner_prefetch(a, kc);
    mark();
    if (guard_scratch) wait_z();      // every CTA has left the scratch that aliases the receive buffer
    mark();
    ep_send(Rl, np);
    mark();
    wait_x();
    mark();
    owner_rows(a, kc);
    mark();
  }

  // ---- the whole stack for one group
  __device__ __forceinline__ void run_group() {
    constexpr int T = Cfg::T, T2 = Cfg::T2;
    // residual rows of this owner from the pre-encode output
    for (int li = ww; li < Cfg::RO; li += 8) {
      const int j = li * 8 + c;
      if (j >= Cfg::R) continue;
      const int s_ = j / T;
      float* rr = r_own() + li * D_MODEL;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = lane * 12 + k * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s_ < nvalid) v = *reinterpret_cast<const float4*>(S.r_in + ((size_t)grp_stream0 * T + j) * D_MODEL + col);
        *reinterpret_cast<float4*>(rr + col) = v;
      }
    }
    __syncwarp();
    {
      OwnerArgs a{};
      a.from_partials = false;
      a.g_next = P.L[0].g_ff1;
      a.norm_next = true;
      a.Tl = T;
      a.tap = -1;
      a.broadcast = true;
      OwnerConsts kc;
      owner_prefetch(a, kc);
      owner_rows(a, kc);
    }
    mark();
    for (int l = 0; l < 16; ++l) {
      const ClLayer& L = P.L[l];
      const bool reduced = l > 6 && l <= 14;
      const int Tl = reduced ? T2 : T;
      const int Rl = Cfg::G * Tl;
      const int np = reduced ? Cfg::NP2 : Cfg::NP;
      stage_consts(l);
      // ---- feed-forward 1
      wait_acc();
      mark();
      ep_swiglu(L.ff1_up_b + 384 * c, Rl, np);
      mark();
      release_h();
      mark();
      {
        OwnerArgs a{};
        a.from_partials = true;
        a.bias = L.ff1_down_b;
        a.scale = 0.5f;
        a.add_resid = true;
        a.g_next = L.g_att;
        a.norm_next = true;
        a.Tl = Tl;
        a.tap = -1;
        a.broadcast = true;
        if (l >= 14) {
          a.kv = (l == 14) ? P.kv14 : P.kv15;
          a.kv_row_off = (l == 14) ? MHSA_S / 2 : MHSA_S;
        }
        reduce_tail(a, np, false);
      }
      // ---- attention
      wait_acc();
      mark();
      attention(l, Tl);
      signal_all(z_bar);
      mark();
      release_h();
      mark();
      uint4 pre[3];
      conv_prefetch(l, pre);          // cache columns of the next stage travel while the partials are exchanged
      {
        OwnerArgs a{};
        a.from_partials = true;
        a.bias = L.wo_b;
        a.scale = 1.0f;
        a.add_resid = true;
        a.g_next = nullptr;           // norm_conv gain is folded into pointwise conv 1
        a.norm_next = true;
        a.Tl = Tl;
        a.tap = -1;
        a.broadcast = true;
        reduce_tail(a, np, true);
      }
      // ---- convolution module
      wait_acc();
      mark();
      conv_module(l, Tl, pre);
      signal_all(z_bar);
      mark();
      release_h();
      mark();
      {
        OwnerArgs a{};
        a.from_partials = true;
        a.bias = L.pw2_b;
        a.scale = 1.0f;
        a.add_resid = true;
        a.g_next = nullptr;           // norm_feed_forward2 gain is folded into ff2_up
        a.norm_next = true;
        a.Tl = Tl;
        a.tap = -1;
        a.broadcast = true;
        reduce_tail(a, np, true);
      }
      // ---- feed-forward 2 + norm_out (+ what follows the layer)
      wait_acc();
      mark();
      ep_swiglu(L.ff2_up_b + 384 * c, Rl, np);
      mark();
      release_h();
      mark();
      {
        OwnerArgs a{};
        a.from_partials = true;
        a.bias = L.ff2_down_b;
        a.scale = 0.5f;
        a.add_resid = true;
        a.g_out = L.g_out;
        a.Tl = Tl;
        a.tap = (l == 6 || l == 14) ? -1 : 1 + l;
        if (l == 6) {
          a.stash_res = true;
          a.broadcast = false;
        } else if (l == 14) {
          a.broadcast = false;
        } else if (l == 15) {
          a.norm_next = false;
          a.broadcast = true;
        } else {
          a.g_next = P.L[l + 1].g_ff1;
          a.norm_next = true;
          a.broadcast = true;
        }
        reduce_tail(a, np, false);
      }
      if (l == 6) {
        __syncwarp();
        reduction_exchange();
        release_h();
        OwnerArgs a{};
        a.from_partials = true;
        a.bias = P.red_pw_b;
        a.scale = 1.0f;
        a.add_resid = false;
        a.g_next = P.L[7].g_ff1;
        a.norm_next = true;
        a.Tl = T2;
        a.tap = 1 + l;
        a.broadcast = true;
        reduce_tail(a, Cfg::NP2, false);
      } else if (l == 14) {
        __syncwarp();
        upsample_exchange();
        OwnerArgs a{};
        a.from_partials = false;
        a.g_next = P.L[15].g_ff1;
        a.norm_next = true;
        a.Tl = T;
        a.tap = 1 + l;
        a.broadcast = true;
        OwnerConsts kc;
        owner_prefetch(a, kc);
        owner_rows(a, kc);
        mark();
      }
    }
    wait_acc();
    decoder();
    tmem_ld_wait();
    tc_fence_before();
    // a fast CTA must not start the next group's operand broadcast while a peer's decoder GEMM still reads its rows
    signal_all(x_bar);
    wait_x();
    mark();
  }
};

// ----------------------------------------------------------------------------------------------- kernel
template <int T_, int G_>
__global__ void __launch_bounds__(CL_THREADS, 1) encoder_cluster_kernel(const ClParams* __restrict__ Pp, const ClStep S) {
  using Cfg = ClCfg<T_, G_>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const ClParams& P = *Pp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::OFF_BAR);
  uint64_t* full = bars;                    // [NST]
  uint64_t* empty = bars + Cfg::NST;        // [NST]
  uint64_t* acc_bar = bars + 2 * Cfg::NST;
  uint64_t* x_bar = acc_bar + 1;
  uint64_t* ag_bar = acc_bar + 2;
  uint64_t* z_bar = acc_bar + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 4);
  volatile int* progress = reinterpret_cast<volatile int*>(acc_bar + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (int)cl_rank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init(x_bar, CL_CTAS);
    mbar_init(ag_bar, CL_CTAS);
    mbar_init(z_bar, CL_CTAS);
    *progress = 0;
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();                       // every CTA's barriers are initialised before any remote arrive

  const int n_groups = (S.B + Cfg::G - 1) / Cfg::G;
  ClRing ring{sm + Cfg::OFF_RING, full, empty, Cfg::NST};

  if (warp == 8) {
    ClProducer<Cfg> prod;
    prod.ring = ring;
    prod.progress = progress;
    for (int grp = (int)cl_id(); grp < n_groups; grp += (int)cl_nclusters()) {
      prod.slots = S.slots + grp * Cfg::G;
      prod.nvalid = min(Cfg::G, S.B - grp * Cfg::G);
      cl_schedule<Cfg>(P, c, prod);
    }
  } else if (warp == 10) {
    ClPrefetcher<Cfg> pf;
    pf.progress = progress;
    for (int grp = (int)cl_id(); grp < n_groups; grp += (int)cl_nclusters()) cl_schedule<Cfg>(P, c, pf);
  } else if (warp == 9) {
    ClMma<Cfg> mw;
    mw.ring = ring;
    mw.smem_base = smem_u32(sm);
    mw.tmem_base = tmem_base;
    mw.acc_bar = acc_bar;
    mw.ag_bar = ag_bar;
    if (S.prof && blockIdx.x == 0) mw.prof = S.prof;
    for (int grp = (int)cl_id(); grp < n_groups; grp += (int)cl_nclusters()) cl_schedule<Cfg>(P, c, mw);
  } else {
    ClWorker<Cfg> w(P, S);
    w.sm = sm;
    w.sm_u32 = smem_u32(sm);
    w.tmem_base = tmem_base;
    w.acc_bar = acc_bar;
    w.x_bar = x_bar;
    w.ag_bar = ag_bar;
    w.z_bar = z_bar;
    w.c = c;
    w.wt = threadIdx.x;
    w.ww = warp;
    w.lane = lane;
    w.q = warp & 3;
    w.h = warp >> 2;
    w.row = w.q * 32 + lane;
    if (S.prof && blockIdx.x == 0) w.prof = S.prof;
    for (int grp = (int)cl_id(); grp < n_groups; grp += (int)cl_nclusters()) {
      w.grp_stream0 = grp * Cfg::G;
      w.nvalid = min(Cfg::G, S.B - grp * Cfg::G);
#pragma unroll
      for (int s = 0; s < Cfg::G; ++s) {
        w.slot_of[s] = s < w.nvalid ? S.slots[grp * Cfg::G + s] : 0;
        w.len_of[s] = s < w.nvalid ? S.len_in[grp * Cfg::G + s] : MHSA_S;
      }
      w.run_group();
    }
  }
  tc_fence_before();
  __syncwarp();
  __syncthreads();
  cluster_sync_all();                       // no CTA leaves while a peer may still write into its shared memory
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

}  // namespace tone
