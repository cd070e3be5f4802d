// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[128 x BN] (fp32, TMEM) = A[128 x K] * W[BN x K]^T, bf16 operands.
//
// One output tile per CTA, 320 threads: warp 0 = TMA producer (one lane), warp 1 = TMEM owner and
// tcgen05.mma issuer (one lane), warps 2..9 = epilogue: two warps per TMEM lane quarter, each taking half of the tile's
// columns in the accumulator phase and half of the quarter's rows in the coalesced write-out.  Operand tiles are
// K-major rows of 64 bf16 (128 B) in the 128-byte swizzle, staged through a STAGES-deep mbarrier ring.
//
// The A tile is either 128 consecutive rows of a dense [M][K] activation matrix, or a GATHER of G boxes of R
// rows, one box per stream, addressed by slot id through a 3-D tensor map over the per-stream state pool.  The
// gather form is how the conv-subsampling layers run as implicit GEMMs without materialising im2col, and how the
// cached-context K/V projections read [cache | new rows] in place.
#pragma once

#include <string.h>

#include "common.cuh"

namespace tone {

constexpr int GEMM_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int VATT_MAX_T = 13;   // most frames per stream and step (400 ms chunks)

enum GemmKind : int {
  G_STORE_F32 = 0,  // out fp32 = acc + bias                                 (q/k/v, out-linear, reduction pw)
  G_RESID = 1,      // r fp32 += scale * (acc + bias)                        (ff down, attn out, conv pw2)
  G_SWIGLU = 2,     // h bf16 = silu(acc_g + b_g) * (acc_v + b_v)            (ff up; tile = 64 gate | 64 value cols)
  G_GLU = 3,        // g bf16 = (acc_a + b_a) * sigmoid(acc_b + b_b)         (conv pw1; tile = 32 a | 32 b cols)
  G_CONV0 = 4,      // x1[slot] bf16 = silu(acc*alpha + beta)                (gather A from feature rows)
  G_CONV1 = 5,      // c1 bf16 = silu(acc*alpha + beta)                      (gather A from x1 windows)
  G_KV = 6,         // out fp32 = acc + bias, gather A from [cache|new] rows (layers 14, 15)
  G_DECODER = 7,    // logprobs = log_softmax(acc + bias)[0:35], argmax      (BN = 48)
  G_PARTIAL = 8,    // part[z] fp16 (saturating) = acc over K slice z (split-K; blockIdx.z)  (ff down; the norm kernel sums
                    // the slices in fp32 in a fixed order)
  G_VATT = 10,      // score-sharing attention layers in one kernel: v = acc + b for one head (BN = 48) of whole streams,
                    // then ctx bf16 = P v with the probabilities P published by the last recompute layer
};

struct GemmArgs {
  int M;            // dense: valid rows; gather: number of streams in the batch
  int nk;           // K iterations of 64
  int R, G;         // gather: rows per stream box, boxes per 128-row tile
  const int* slots; // gather: slot id per batch position
  void* out;        // output base
  int ldo;          // output leading dimension (elements)
  const float* bias;
  const float* alpha;
  const float* beta;
  float scale;
  long long out_slot_stride;  // G_CONV0: elements between consecutive slots of x1
  int out_row_off;            // G_CONV0: first row written inside a slot (the cached rows come first)
  const float* P;             // G_VATT: [streams][8][R][R] attention probabilities (R = frames per stream at this rate)
  int* tokens;                // G_DECODER: argmax per frame
  float* aux;                 // G_DECODER: [rows][2] = logprob of ' ' (33) and of blank (34), what the phrase splitter needs
  long long z_stride;         // G_PARTIAL: elements between the partial outputs of consecutive K slices
  // Row-scale form of RMSNorm (x g / (rms + eps)) folded around a GEMM: the PRODUCER (G_RESID) also emits the new
  // residual row in bf16 (rb_out) and, per N tile, the sum of squares of its columns (ss_out[row][tile]); the
  // CONSUMER (G_GLU / G_SWIGLU) reads A = rb, has the norm gain folded into its weights, and scales the accumulator
  // row by 1 / (sqrt(sum_tiles ss) * d^-1/2 + eps).  Removes the RMSNorm kernel between the two GEMMs.
  bf16* rb_out;
  float* ss_out;
  int ss_ld;                  // floats per row of ss (= number of N tiles of the producer)
  const float* ss;            // consumer side: nullable
  int ss_tiles;
  // Raw operand views, used only by the SIMT debug kernels (gemm_ref.cuh); the tensor-core path reads through
  // the tensor maps.
  const bf16* A;
  int lda;
  long long a_slot_stride;    // gather: elements between consecutive slots of the A pool
  const bf16* W;
  int ldw;
};

template <int KIND>
struct KindTraits {
  static constexpr bool gather = (KIND == G_CONV0 || KIND == G_CONV1 || KIND == G_KV);
  // rows of a tile are G whole streams x R frames (gather kinds, and dense kinds that need whole streams per tile)
  static constexpr bool stream_rows = gather || (KIND == G_VATT);
};

// DEEP = one CTA per SM with the whole shared memory as the operand ring: the small-batch GEMMs of this model are
// latency bound (few CTAs, short K loops), so all that matters is how many TMA loads are in flight.
template <int BN, bool DEEP = true>
struct TileCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGES = DEEP ? (200 * 1024) / (A_BYTES + B_BYTES) : ((BN > 64) ? 3 : 4);
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  // + barriers (256) + per-column constants (1 KB) + alignment slack
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 + 1024 + 1024;
};

// ---------------------------------------------------------------------------------------------- epilogues
// Row bookkeeping shared by all kinds.
struct RowInfo {
  bool valid;
  long long out_row;  // row index into a dense output, or element offset base for the scatter kind
};

template <int KIND>
__device__ __forceinline__ RowInfo row_info(const GemmArgs& a, int row_in_tile) {
  RowInfo ri;
  if constexpr (KindTraits<KIND>::stream_rows) {
    int g = row_in_tile / a.R;
    int j = row_in_tile - g * a.R;
    int b = blockIdx.x * a.G + g;
    ri.valid = (g < a.G) && (b < a.M);
    if constexpr (KIND == G_CONV0) {
      int slot = ri.valid ? a.slots[b] : 0;
      ri.out_row = (long long)slot * a.out_slot_stride + (long long)(a.out_row_off + j) * a.ldo;
    } else {
      ri.out_row = (long long)b * a.R + j;
    }
  } else {
    int m = blockIdx.x * 128 + row_in_tile;
    ri.valid = m < a.M;
    ri.out_row = m;
  }
  return ri;
}

// Per-tile constants (bias, or folded BatchNorm scale/shift per output column) are staged in shared memory by the
// epilogue warps while the main loop runs; they are weights, so this happens before the PDL wait.
template <int KIND, int BN>
__device__ __forceinline__ void stage_constants(const GemmArgs& a, float* s_c0, float* s_c1, int t /*0..255*/) {
  const int n0 = blockIdx.y * BN;
  if constexpr (KIND == G_PARTIAL) {
    return;
  } else if constexpr (KIND == G_CONV0 || KIND == G_CONV1) {
    constexpr int CH = (KIND == G_CONV0) ? 32 : 64;
    if (t < BN) {
      s_c0[t] = __ldg(a.alpha + (n0 + t) % CH);
      s_c1[t] = __ldg(a.beta + (n0 + t) % CH);
    }
  } else if constexpr (KIND == G_DECODER) {
    if (t < 48) s_c0[t] = t < 35 ? __ldg(a.bias + t) : 0.f;
  } else {
    if (t < BN) s_c0[t] = a.bias ? __ldg(a.bias + n0 + t) : 0.f;
  }
}

// Output geometry of one tile row, in bytes of the FINAL output type.
template <int KIND, int BN>
struct OutCfg {
  static constexpr bool f32 = (KIND == G_STORE_F32 || KIND == G_KV || KIND == G_RESID);
  static constexpr int ROW_BYTES = f32 ? BN * 4 : ((KIND == G_SWIGLU || KIND == G_GLU) ? BN : BN * 2);
  static constexpr int STRIDE = ROW_BYTES + 16;   // +16 B: float4 stores of a quarter-warp hit distinct banks
  static constexpr int LPR = ROW_BYTES / 16;      // lanes per output row in the coalesced phase
  static constexpr int RPI = 32 / LPR;            // rows per warp instruction
  static constexpr int NIT = 16 / RPI;            // write-out iterations per warp (16 rows each: two warps per quarter)
  static_assert(RPI <= 16, "a write-out instruction must stay inside one warp's 16 rows");
};

template <int KIND>
__device__ __forceinline__ char* out_row_ptr(const GemmArgs& a, const RowInfo& ri, int n0_elems) {
  if constexpr (KIND == G_PARTIAL)
    return reinterpret_cast<char*>(reinterpret_cast<__half*>(a.out) + blockIdx.z * a.z_stride + ri.out_row * a.ldo + n0_elems);
  else if constexpr (KIND == G_STORE_F32 || KIND == G_KV || KIND == G_RESID)
    return reinterpret_cast<char*>(reinterpret_cast<float*>(a.out) + ri.out_row * a.ldo + n0_elems);
  else if constexpr (KIND == G_CONV0)
    return reinterpret_cast<char*>(reinterpret_cast<bf16*>(a.out) + ri.out_row + n0_elems);
  else
    return reinterpret_cast<char*>(reinterpret_cast<bf16*>(a.out) + ri.out_row * (long long)a.ldo + n0_elems);
}

// Epilogue of one warp.  Warps (q, hf), hf = 0 | 1, share TMEM lanes / tile rows 32q .. 32q+31.  Two phases:
//  1. each thread owns one accumulator row and HALF of its columns (hf): tcgen05.ld, apply the column-wise math (bias,
//     activation, gating, folded BatchNorm), and park the result in shared memory in its final element type;
//  2. after the 8 epilogue warps have met, warp (q, hf) writes rows 32q + 16hf .. +15 back out with lanes running along
//     the row, so every global store (and the residual read of G_RESID) is a contiguous 16 B-per-lane access.
// The staging area is the operand ring, which is idle once the accumulator barrier has fired.
__device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int KIND, int BN>
__device__ __forceinline__ void epilogue(const GemmArgs& a, uint32_t tmem_row_base, int q, int hf, int lane, char* stage,
                                         const float* s_c0, const float* s_c1, uint64_t* tmem_full) {
  using O = OutCfg<KIND, BN>;
  const int n0 = blockIdx.y * BN;

  if constexpr (KIND == G_DECODER) {
    if (hf) return;
    const RowInfo ri = row_info<KIND>(a, q * 32 + lane);
    float lg[48];
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    tmem_ld16(tmem_row_base + 0, lg);
    tmem_ld16(tmem_row_base + 16, lg + 16);
    tmem_ld16(tmem_row_base + 32, lg + 32);
    if (ri.valid) {
      float mx = -INFINITY;
      int am = 0;
#pragma unroll
      for (int i = 0; i < 35; ++i) {
        lg[i] += s_c0[i];
        if (lg[i] > mx) {  // strict > keeps the first maximum (numpy argmax, tone/decoder.py:57)
          mx = lg[i];
          am = i;
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 35; ++i) sum += expf(lg[i] - mx);
      const float lse = mx + logf(sum);
      float* __restrict__ out = reinterpret_cast<float*>(a.out) + ri.out_row * 35;
#pragma unroll
      for (int i = 0; i < 35; ++i) out[i] = lg[i] - lse;
      if (a.tokens) a.tokens[ri.out_row] = am;
      if (a.aux) *reinterpret_cast<float2*>(a.aux + ri.out_row * 2) = make_float2(lg[33] - lse, lg[34] - lse);
    }
    return;
  } else if constexpr (KIND == G_VATT) {
    // Tile = G whole streams x R frames (rows), one head (48 columns).  Phase 1: v rows (+ bias) -> fp32 staging, and the
    // tile's probability blocks (one contiguous R x R block per stream and head) -> shared memory beside them.
    // Phase 2: unit = (row, 8 dims): ctx[t][d] = sum_j P[t][j] v[j][d] over the R rows of the row's stream.
    static_assert(KIND != G_VATT || BN == 48, "one head per tile");
    constexpr int VLD = 52;                              // floats per staged v row (16 B aligned, conflict-light)
    const int R = a.R, head = blockIdx.y, RR = R * R;
    float* vst = reinterpret_cast<float*>(stage);
    float* pst = vst + 128 * VLD;                        // [G][R * R]
    const int et = (threadIdx.x - 64);
    // This thread's share of the P blocks, fetched (coalesced: consecutive threads, consecutive floats of a block) before
    // the accumulator is ready - they come from an earlier layer.  A per-unit gather of P rows (39 scalar loads per thread,
    // ~5000 sector requests per CTA) used to be what bound this epilogue.
    constexpr int NP = (128 * VATT_MAX_T + EPI_THREADS - 1) / EPI_THREADS;   // G * R * R <= 128 * 13
    float pre[NP];
    const int np = a.G * RR, b0 = blockIdx.x * a.G;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int e = et + i * EPI_THREADS;
      pre[i] = 0.f;
      if (e < np) {
        const int g = e / RR, x = e - g * RR;
        if (b0 + g < a.M) pre[i] = __ldg(a.P + ((size_t)(b0 + g) * 8 + head) * RR + x);
      }
    }
    // row-scale RMSNorm folded around the V projection (A = bf16 residual rows, gain folded into the weights)
    float rsv = 1.f;
    if (a.ss) {
      const RowInfo rme = row_info<KIND>(a, q * 32 + lane);
      if (rme.valid) {
        const float* sp = a.ss + rme.out_row * a.ss_ld;
        float tsum = 0.f;
        for (int k = 0; k < a.ss_tiles; ++k) tsum += sp[k];
        rsv = 1.0f / (sqrtf(tsum) * 0.05103103630798288f + 1e-8f);
      }
    }
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    {
      // 24 columns per half: two 16-wide TMEM loads (the second overlaps into the other half, 8 used)
      uint32_t r0[16], r1[16];
      tmem_ld16_async(tmem_row_base + 24 * hf, r0);
      tmem_ld16_async(tmem_row_base + 24 * hf + 16, r1);
      tmem_ld_wait();
      tmem_regs_ready16(r0);
      tmem_regs_ready16(r1);
      const uint32_t vr = smem_u32(vst) + ((q * 32 + lane) * VLD + 24 * hf) * 4;
#pragma unroll
      for (int c = 0; c < 24; c += 4) {
        float4 o;
        o.x = fmaf(__uint_as_float(c < 16 ? r0[c] : r1[c - 16]), rsv, s_c0[24 * hf + c]);
        o.y = fmaf(__uint_as_float(c + 1 < 16 ? r0[c + 1] : r1[c + 1 - 16]), rsv, s_c0[24 * hf + c + 1]);
        o.z = fmaf(__uint_as_float(c + 2 < 16 ? r0[c + 2] : r1[c + 2 - 16]), rsv, s_c0[24 * hf + c + 2]);
        o.w = fmaf(__uint_as_float(c + 3 < 16 ? r0[c + 3] : r1[c + 3 - 16]), rsv, s_c0[24 * hf + c + 3]);
        sts128(vr + c * 4, o);
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int e = et + i * EPI_THREADS;
      if (e < np) sts32(smem_u32(pst) + e * 4, pre[i]);
    }
    bar_epilogue();
    // unit = (stream, 8-dim group, half of the stream's frames when R > 7): every v element is read once per unit, not
    // once per output row - shared-memory read bandwidth (a 32-lane LDS.128 returns 512 B whatever it broadcasts) is what
    // bounds this phase, and two CTAs share an SM
    {
      const int nth = R > 7 ? 2 : 1, rows_u = (R + nth - 1) / nth;           // frames per unit: <= 7
      const int nunits = a.G * 6 * nth;
      for (int unit = et; unit < nunits; unit += EPI_THREADS) {
        const int g = unit / (6 * nth), rem = unit - g * (6 * nth), uu = rem / nth, th = rem - uu * nth;
        if (b0 + g >= a.M) continue;
        const int t0 = th * rows_u;
        // shared-window addresses (generic pointers cost 64-bit address arithmetic and LD.E per access)
        uint32_t va = smem_u32(vst) + ((g * R) * VLD + uu * 8) * 4;
        uint32_t pa = smem_u32(pst) + (g * RR + t0 * R) * 4;
        float acc[7][8];
#pragma unroll
        for (int t = 0; t < 7; ++t)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
        const int nrows = min(rows_u, R - t0);
        // All seven row slots are computed (rows beyond nrows read whatever follows in the staging area and are never
        // stored): no branches in the loop, and the seven P loads of a step are issued together ahead of the FMAs.
        const uint32_t r4 = (uint32_t)R * 4;
#pragma unroll 2
        for (int j = 0; j < R; ++j, va += VLD * 4, pa += 4) {
          const float4 v0 = lds128(va), v1 = lds128(va + 16);
          float pv[7];
#pragma unroll
          for (int t = 0; t < 7; ++t) pv[t] = lds32(pa + t * r4);
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            const float p = pv[t];
            acc[t][0] = fmaf(p, v0.x, acc[t][0]); acc[t][1] = fmaf(p, v0.y, acc[t][1]);
            acc[t][2] = fmaf(p, v0.z, acc[t][2]); acc[t][3] = fmaf(p, v0.w, acc[t][3]);
            acc[t][4] = fmaf(p, v1.x, acc[t][4]); acc[t][5] = fmaf(p, v1.y, acc[t][5]);
            acc[t][6] = fmaf(p, v1.z, acc[t][6]); acc[t][7] = fmaf(p, v1.w, acc[t][7]);
          }
        }
        bf16* dst = reinterpret_cast<bf16*>(a.out) + ((long long)(b0 + g) * R + t0) * a.ldo + head * 48 + uu * 8;
#pragma unroll
        for (int t = 0; t < 7; ++t)
          if (t < nrows)
            *reinterpret_cast<uint4*>(dst + (long long)t * a.ldo) =
                make_uint4(pack_bf16x2(acc[t][0], acc[t][1]), pack_bf16x2(acc[t][2], acc[t][3]), pack_bf16x2(acc[t][4], acc[t][5]),
                           pack_bf16x2(acc[t][6], acc[t][7]));
      }
    }
    return;
  } else {
    constexpr bool gated = (KIND == G_SWIGLU || KIND == G_GLU);
    // element offset of this tile's first output column
    const int n0_out = gated ? blockIdx.y * (BN / 2) : n0;
    // phase-2 geometry (also used to prefetch the residual before the accumulator is ready)
    const int row0 = q * 32 + hf * 16;                       // first of this warp's 16 write-out rows
    const int sub_row = lane / O::LPR, cb = (lane % O::LPR) * 16;
    float4 rres[O::NIT];
    if constexpr (KIND == G_RESID) {
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const RowInfo ri = row_info<KIND>(a, row0 + it * O::RPI + sub_row);
        if (ri.valid) rres[it] = *reinterpret_cast<const float4*>(out_row_ptr<KIND>(a, ri, n0_out) + cb);
      }
    }
    // consumer side of the row-scale RMSNorm: 1 / (rms + eps) of this thread's row, from the producer's partial sums
    float rs = 1.f;
    if constexpr (gated || KIND == G_STORE_F32) {
      if (a.ss) {
        const RowInfo rme = row_info<KIND>(a, q * 32 + lane);
        if (rme.valid) {
          const float* sp = a.ss + rme.out_row * a.ss_ld;
          float t = 0.f;
          for (int k = 0; k < a.ss_tiles; ++k) t += sp[k];
          rs = 1.0f / (sqrtf(t) * 0.05103103630798288f + 1e-8f);   // 384^-1/2, eps outside the sqrt (submodules.py:50-52)
        }
      }
    }
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();

    // ---- phase 1: half of the accumulator row -> final-type row in shared memory (all TMEM loads in flight, one wait)
    const uint32_t srow = smem_u32(stage) + (q * 32 + lane) * O::STRIDE;
    const uint32_t c0a = smem_u32(s_c0), c1a = smem_u32(s_c1);
    if constexpr (O::f32) {
      constexpr int NC = BN / 2;                 // this warp's columns [hf * NC, hf * NC + NC)
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 4) {
        float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        {
          const float4 bb = lds128(c0a + (cbase + c) * 4);
          if constexpr (KIND == G_STORE_F32) {   // optional row-scale RMSNorm of the A rows (rs = 1 without it)
            o.x = fmaf(o.x, rs, bb.x);
            o.y = fmaf(o.y, rs, bb.y);
            o.z = fmaf(o.z, rs, bb.z);
            o.w = fmaf(o.w, rs, bb.w);
          } else {
            o.x += bb.x;
            o.y += bb.y;
            o.z += bb.z;
            o.w += bb.w;
          }
        }
        if constexpr (KIND == G_RESID) {
          o.x *= a.scale;
          o.y *= a.scale;
          o.z *= a.scale;
          o.w *= a.scale;
        }
        sts128(srow + (cbase + c) * 4, o);
      }
    } else if constexpr (gated) {
      constexpr int HW = BN / 2;  // first half of the tile = gate / a, second half = value / b
      constexpr int NC = HW / 2;  // this warp's output columns [hf * NC, hf * NC + NC)
      const int cbase = hf * NC;
      float ag[NC], av[NC];
      {
        uint32_t* rg = reinterpret_cast<uint32_t*>(ag);
        uint32_t* rv = reinterpret_cast<uint32_t*>(av);
#pragma unroll
        for (int c = 0; c < NC; c += 16) {
          tmem_ld16_async(tmem_row_base + cbase + c, rg + c);
          tmem_ld16_async(tmem_row_base + HW + cbase + c, rv + c);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; c += 16) {
          tmem_regs_ready16(rg + c);
          tmem_regs_ready16(rv + c);
        }
      }
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        const float4 g0 = lds128(c0a + (cbase + c) * 4), g1 = lds128(c0a + (cbase + c) * 4 + 16);
        const float4 u0 = lds128(c0a + (HW + cbase + c) * 4), u1 = lds128(c0a + (HW + cbase + c) * 4 + 16);
        const float gb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float ub[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = fmaf(ag[c + i], rs, gb[i]), y = fmaf(av[c + i], rs, ub[i]);
          r[i] = (KIND == G_SWIGLU) ? silu_f(x) * y : x * sigmoid_f(y);
        }
        sts128u(srow + (cbase + c) * 2, make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                                                   pack_bf16x2(r[6], r[7])));
      }
    } else if constexpr (KIND == G_PARTIAL) {   // raw K-slice sums as saturating fp16
      constexpr int NC = BN / 2;
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half2 hv = __floats2half2_rn(fminf(fmaxf(acc[c + 2 * i], -65504.f), 65504.f),
                                               fminf(fmaxf(acc[c + 2 * i + 1], -65504.f), 65504.f));
          pk[i] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        sts128u(srow + (cbase + c) * 2, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
    } else {  // G_CONV0 / G_CONV1: folded BatchNorm + SiLU per output channel
      constexpr int NC = BN / 2;
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        const float4 a0 = lds128(c0a + (cbase + c) * 4), a1 = lds128(c0a + (cbase + c) * 4 + 16);
        const float4 b0 = lds128(c1a + (cbase + c) * 4), b1 = lds128(c1a + (cbase + c) * 4 + 16);
        const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = silu_exact_f(acc[c + i] * al[i] + be[i]);
        sts128u(srow + (cbase + c) * 2, make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                                                   pack_bf16x2(r[6], r[7])));
      }
    }
    bar_epilogue();   // both halves of every row are staged

    // ---- phase 2: coalesced write-out of this warp's 16 rows
    if constexpr (KIND == G_RESID) {
      // All rows of the warp go through every step together (staging loads, adds, stores, the sum-of-squares shuffles):
      // walking them one row at a time is a chain of ~16 x (LDS -> FADD -> STG, 5 dependent shuffles) latencies with two
      // warps per scheduler to hide them.
      float sq[O::NIT];
      bool ok[O::NIT];
      long long orow[O::NIT];
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const int row = row0 + it * O::RPI + sub_row;
        const RowInfo ri = row_info<KIND>(a, row);
        ok[it] = ri.valid;
        orow[it] = ri.out_row;
        const float4 d = lds128(smem_u32(stage) + row * O::STRIDE + cb);
        float4 o = ri.valid ? rres[it] : make_float4(0.f, 0.f, 0.f, 0.f);
        o.x += d.x;
        o.y += d.y;
        o.z += d.z;
        o.w += d.w;
        rres[it] = o;
        sq[it] = o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
      }
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        if (ok[it]) {
          RowInfo ri;
          ri.valid = true;
          ri.out_row = orow[it];
          *reinterpret_cast<float4*>(out_row_ptr<KIND>(a, ri, n0_out) + cb) = rres[it];
          if (a.rb_out)
            *reinterpret_cast<uint2*>(a.rb_out + ri.out_row * a.ldo + n0_out + cb / 4) =
                make_uint2(pack_bf16x2(rres[it].x, rres[it].y), pack_bf16x2(rres[it].z, rres[it].w));
        }
      }
      if (a.rb_out) {   // reduce every row's LPR lanes; lane 0 of each row group stores the tile's sum
#pragma unroll
        for (int off = O::LPR / 2; off > 0; off >>= 1) {
#pragma unroll
          for (int it = 0; it < O::NIT; ++it) sq[it] += __shfl_xor_sync(0xffffffffu, sq[it], off);
        }
#pragma unroll
        for (int it = 0; it < O::NIT; ++it)
          if (ok[it] && (lane % O::LPR) == 0) a.ss_out[orow[it] * a.ss_ld + blockIdx.y] = sq[it];
      }
    } else {
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const int row = row0 + it * O::RPI + sub_row;
        const RowInfo ri = row_info<KIND>(a, row);
        if (ri.valid) {
          char* dst = out_row_ptr<KIND>(a, ri, n0_out) + cb;
          *reinterpret_cast<uint4*>(dst) = lds128u(smem_u32(stage) + row * O::STRIDE + cb);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- persistent-kernel epilogues
// The same epilogue code with the tile coordinates, the accumulator-barrier parity, the TMEM-release barrier, the
// named-barrier id and a precomputed row scale as arguments.  Kept as a separate copy so that the one-tile kernel
// compiles to exactly the code it had before (threading these through the shared functions cost 64 streams 1.3 %).
template <int KIND, bool PERSIST = true>
__device__ __forceinline__ RowInfo row_info_p(const GemmArgs& a, int tx, int row_in_tile) {
  RowInfo ri;
  if constexpr (KindTraits<KIND>::stream_rows) {
    int g = row_in_tile / a.R;
    int j = row_in_tile - g * a.R;
    int b = tx * a.G + g;
    ri.valid = (g < a.G) && (b < a.M);
    if constexpr (KIND == G_CONV0) {
      int slot = ri.valid ? a.slots[b] : 0;
      ri.out_row = (long long)slot * a.out_slot_stride + (long long)(a.out_row_off + j) * a.ldo;
    } else {
      ri.out_row = (long long)b * a.R + j;
    }
  } else {
    int m = tx * 128 + row_in_tile;
    ri.valid = m < a.M;
    ri.out_row = m;
  }
  return ri;
}

// Named barrier of one epilogue group (ids 1 and 2).
__device__ __forceinline__ void bar_epilogue_p(int id = 1) {   // immediate barrier ids (the id is a constant after inlining)
  if (id == 1) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}

template <int KIND, int BN, bool PERSIST = true>
// (tx, ty) = tile coordinates (arguments); par = phase parity of tmem_full;
// tmem_empty (persistent kernel only): shared::cluster address of the barrier each warp arrives on once its
// accumulator columns are out of TMEM (0 = none).
__device__ __forceinline__ void epilogue_p(const GemmArgs& a, uint32_t tmem_row_base, int q, int hf, int lane, char* stage,
                                         const float* s_c0, const float* s_c1, uint64_t* tmem_full, int tx_in, int ty_in,
                                         uint32_t par_in = 0, uint32_t tmem_empty_in = 0, int bar_id_in = 1,
                                         const float* rs_pre_in = nullptr) {
  using O = OutCfg<KIND, BN>;
  const int tx = PERSIST ? tx_in : (int)blockIdx.x, ty = PERSIST ? ty_in : (int)blockIdx.y;
  const uint32_t par = PERSIST ? par_in : 0u, tmem_empty = PERSIST ? tmem_empty_in : 0u;
  const int bar_id = PERSIST ? bar_id_in : 1;
  const float* rs_pre = PERSIST ? rs_pre_in : nullptr;
  const int n0 = ty * BN;

  if constexpr (KIND == G_DECODER) {
    if (hf) return;
    const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, q * 32 + lane);
    float lg[48];
    mbar_wait(tmem_full, par);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    tmem_ld16(tmem_row_base + 0, lg);
    tmem_ld16(tmem_row_base + 16, lg + 16);
    tmem_ld16(tmem_row_base + 32, lg + 32);
    if (ri.valid) {
      float mx = -INFINITY;
      int am = 0;
#pragma unroll
      for (int i = 0; i < 35; ++i) {
        lg[i] += s_c0[i];
        if (lg[i] > mx) {  // strict > keeps the first maximum (numpy argmax, tone/decoder.py:57)
          mx = lg[i];
          am = i;
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 35; ++i) sum += expf(lg[i] - mx);
      const float lse = mx + logf(sum);
      float* __restrict__ out = reinterpret_cast<float*>(a.out) + ri.out_row * 35;
#pragma unroll
      for (int i = 0; i < 35; ++i) out[i] = lg[i] - lse;
      if (a.tokens) a.tokens[ri.out_row] = am;
      if (a.aux) *reinterpret_cast<float2*>(a.aux + ri.out_row * 2) = make_float2(lg[33] - lse, lg[34] - lse);
    }
    return;
  } else if constexpr (KIND == G_VATT) {
    // Tile = G whole streams x R frames (rows), one head (48 columns).  Phase 1: v rows (+ bias) -> fp32 staging.
    // Phase 2: unit = (row, 8 dims): ctx[t][d] = sum_j P[t][j] v[j][d] over the R rows of the row's stream.
    static_assert(KIND != G_VATT || BN == 48, "one head per tile");
    constexpr int VLD = 52;                              // floats per staged v row (16 B aligned, conflict-light)
    const int R = a.R, head = ty;
    float* vst = reinterpret_cast<float*>(stage);
    // P rows of this thread's units, fetched before the accumulator is ready (they come from an earlier layer)
    constexpr int NU = 3;                                // units per thread: 128 rows x 6 / 256 threads
    const int et = (threadIdx.x - 64);
    float pr[NU][VATT_MAX_T];
    int urow[NU], uu[NU];
    bool uok[NU];
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int unit = et + k * EPI_THREADS;
      urow[k] = unit / 6;
      uu[k] = unit - urow[k] * 6;
      const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, urow[k] < 128 ? urow[k] : 0);
      uok[k] = urow[k] < 128 && ri.valid;
      if (uok[k]) {
        const int b = (int)(ri.out_row / R), t = (int)(ri.out_row - (long long)b * R);
        const float* pp = a.P + (((size_t)b * 8 + head) * R + t) * R;
#pragma unroll
        for (int j = 0; j < VATT_MAX_T; ++j) pr[k][j] = j < R ? __ldg(pp + j) : 0.f;
      }
    }
    float rsv = 1.f;
    if (a.ss) {
      const RowInfo rme = row_info_p<KIND, PERSIST>(a, tx, q * 32 + lane);
      if (rme.valid) {
        const float* sp = a.ss + rme.out_row * a.ss_ld;
        float tsum = 0.f;
        for (int k = 0; k < a.ss_tiles; ++k) tsum += sp[k];
        rsv = 1.0f / (sqrtf(tsum) * 0.05103103630798288f + 1e-8f);
      }
    }
    mbar_wait(tmem_full, par);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();
    {
      // 24 columns per half: two 16-wide TMEM loads (the second overlaps into the other half, 8 used)
      uint32_t r0[16], r1[16];
      tmem_ld16_async(tmem_row_base + 24 * hf, r0);
      tmem_ld16_async(tmem_row_base + 24 * hf + 16, r1);
      tmem_ld_wait();
      tmem_regs_ready16(r0);
      tmem_regs_ready16(r1);
      float* vr = vst + (q * 32 + lane) * VLD + 24 * hf;
#pragma unroll
      for (int c = 0; c < 24; c += 4) {
        float4 o;
        o.x = fmaf(__uint_as_float(c < 16 ? r0[c] : r1[c - 16]), rsv, s_c0[24 * hf + c]);
        o.y = fmaf(__uint_as_float(c + 1 < 16 ? r0[c + 1] : r1[c + 1 - 16]), rsv, s_c0[24 * hf + c + 1]);
        o.z = fmaf(__uint_as_float(c + 2 < 16 ? r0[c + 2] : r1[c + 2 - 16]), rsv, s_c0[24 * hf + c + 2]);
        o.w = fmaf(__uint_as_float(c + 3 < 16 ? r0[c + 3] : r1[c + 3 - 16]), rsv, s_c0[24 * hf + c + 3]);
        *reinterpret_cast<float4*>(vr + c) = o;
      }
    }
    bar_epilogue_p(bar_id);
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      if (!uok[k]) continue;
      const int g = urow[k] / R;                          // stream within the tile
      const float* vb = vst + (g * R) * VLD + uu[k] * 8;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < VATT_MAX_T; ++j) {
        if (j < R) {
          const float4 v0 = *reinterpret_cast<const float4*>(vb + j * VLD);
          const float4 v1 = *reinterpret_cast<const float4*>(vb + j * VLD + 4);
          const float p = pr[k][j];
          acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v0.z, acc[2]); acc[3] = fmaf(p, v0.w, acc[3]);
          acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]); acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
        }
      }
      const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, urow[k]);
      bf16* dst = reinterpret_cast<bf16*>(a.out) + ri.out_row * (long long)a.ldo + head * 48 + uu[k] * 8;
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                  pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
    }
    return;
  } else {
    constexpr bool gated = (KIND == G_SWIGLU || KIND == G_GLU);
    // element offset of this tile's first output column
    const int n0_out = gated ? ty * (BN / 2) : n0;
    // phase-2 geometry (also used to prefetch the residual before the accumulator is ready)
    const int row0 = q * 32 + hf * 16;                       // first of this warp's 16 write-out rows
    const int sub_row = lane / O::LPR, cb = (lane % O::LPR) * 16;
    float4 rres[O::NIT];
    if constexpr (KIND == G_RESID) {
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, row0 + it * O::RPI + sub_row);
        if (ri.valid) rres[it] = *reinterpret_cast<const float4*>(out_row_ptr<KIND>(a, ri, n0_out) + cb);
      }
    }
    // consumer side of the row-scale RMSNorm: 1 / (rms + eps) of this thread's row, from the producer's partial sums
    float rs = 1.f;
    if constexpr (gated || KIND == G_STORE_F32) {
      if (rs_pre) {
        rs = *rs_pre;
      } else if (a.ss) {
        const RowInfo rme = row_info_p<KIND, PERSIST>(a, tx, q * 32 + lane);
        if (rme.valid) {
          const float* sp = a.ss + rme.out_row * a.ss_ld;
          float t = 0.f;
          for (int k = 0; k < a.ss_tiles; ++k) t += sp[k];
          rs = 1.0f / (sqrtf(t) * 0.05103103630798288f + 1e-8f);   // 384^-1/2, eps outside the sqrt (submodules.py:50-52)
        }
      }
    }
    mbar_wait(tmem_full, par);
    if (threadIdx.x == 64) PROF_MARK(4);
    tc_fence_after();

    // ---- phase 1: half of the accumulator row -> final-type row in shared memory (all TMEM loads in flight, one wait)
    const uint32_t srow = smem_u32(stage) + (q * 32 + lane) * O::STRIDE;
    const uint32_t c0a = smem_u32(s_c0), c1a = smem_u32(s_c1);
    if constexpr (O::f32) {
      constexpr int NC = BN / 2;                 // this warp's columns [hf * NC, hf * NC + NC)
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 4) {
        float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        {
          const float4 bb = lds128(c0a + (cbase + c) * 4);
          if constexpr (KIND == G_STORE_F32) {   // optional row-scale RMSNorm of the A rows (rs = 1 without it)
            o.x = fmaf(o.x, rs, bb.x);
            o.y = fmaf(o.y, rs, bb.y);
            o.z = fmaf(o.z, rs, bb.z);
            o.w = fmaf(o.w, rs, bb.w);
          } else {
            o.x += bb.x;
            o.y += bb.y;
            o.z += bb.z;
            o.w += bb.w;
          }
        }
        if constexpr (KIND == G_RESID) {
          o.x *= a.scale;
          o.y *= a.scale;
          o.z *= a.scale;
          o.w *= a.scale;
        }
        sts128(srow + (cbase + c) * 4, o);
      }
    } else if constexpr (gated) {
      constexpr int HW = BN / 2;  // first half of the tile = gate / a, second half = value / b
      constexpr int NC = HW / 2;  // this warp's output columns [hf * NC, hf * NC + NC)
      const int cbase = hf * NC;
      float ag[NC], av[NC];
      {
        uint32_t* rg = reinterpret_cast<uint32_t*>(ag);
        uint32_t* rv = reinterpret_cast<uint32_t*>(av);
#pragma unroll
        for (int c = 0; c < NC; c += 16) {
          tmem_ld16_async(tmem_row_base + cbase + c, rg + c);
          tmem_ld16_async(tmem_row_base + HW + cbase + c, rv + c);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; c += 16) {
          tmem_regs_ready16(rg + c);
          tmem_regs_ready16(rv + c);
        }
      }
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        const float4 g0 = lds128(c0a + (cbase + c) * 4), g1 = lds128(c0a + (cbase + c) * 4 + 16);
        const float4 u0 = lds128(c0a + (HW + cbase + c) * 4), u1 = lds128(c0a + (HW + cbase + c) * 4 + 16);
        const float gb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float ub[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = fmaf(ag[c + i], rs, gb[i]), y = fmaf(av[c + i], rs, ub[i]);
          r[i] = (KIND == G_SWIGLU) ? silu_f(x) * y : x * sigmoid_f(y);
        }
        sts128u(srow + (cbase + c) * 2, make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                                                   pack_bf16x2(r[6], r[7])));
      }
    } else if constexpr (KIND == G_PARTIAL) {   // raw K-slice sums as saturating fp16
      constexpr int NC = BN / 2;
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half2 hv = __floats2half2_rn(fminf(fmaxf(acc[c + 2 * i], -65504.f), 65504.f),
                                               fminf(fmaxf(acc[c + 2 * i + 1], -65504.f), 65504.f));
          pk[i] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        sts128u(srow + (cbase + c) * 2, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
    } else {  // G_CONV0 / G_CONV1: folded BatchNorm + SiLU per output channel
      constexpr int NC = BN / 2;
      const int cbase = hf * NC;
      float acc[NC];
      tmem_load_row<NC>(tmem_row_base + cbase, acc);
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        const float4 a0 = lds128(c0a + (cbase + c) * 4), a1 = lds128(c0a + (cbase + c) * 4 + 16);
        const float4 b0 = lds128(c1a + (cbase + c) * 4), b1 = lds128(c1a + (cbase + c) * 4 + 16);
        const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = silu_exact_f(acc[c + i] * al[i] + be[i]);
        sts128u(srow + (cbase + c) * 2, make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                                                   pack_bf16x2(r[6], r[7])));
      }
    }
    if (tmem_empty) {   // this warp's accumulator columns are in shared memory: the MMA warp may overwrite them
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tmem_empty) : "memory");
    }
    bar_epilogue_p(bar_id);   // both halves of every row are staged

    // ---- phase 2: coalesced write-out of this warp's 16 rows
    if constexpr (KIND == G_RESID) {
      // All rows of the warp go through every step together (staging loads, adds, stores, the sum-of-squares shuffles):
      // walking them one row at a time is a chain of ~16 x (LDS -> FADD -> STG, 5 dependent shuffles) latencies with two
      // warps per scheduler to hide them.
      float sq[O::NIT];
      bool ok[O::NIT];
      long long orow[O::NIT];
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const int row = row0 + it * O::RPI + sub_row;
        const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, row);
        ok[it] = ri.valid;
        orow[it] = ri.out_row;
        const float4 d = lds128(smem_u32(stage) + row * O::STRIDE + cb);
        float4 o = ri.valid ? rres[it] : make_float4(0.f, 0.f, 0.f, 0.f);
        o.x += d.x;
        o.y += d.y;
        o.z += d.z;
        o.w += d.w;
        rres[it] = o;
        sq[it] = o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
      }
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        if (ok[it]) {
          RowInfo ri;
          ri.valid = true;
          ri.out_row = orow[it];
          *reinterpret_cast<float4*>(out_row_ptr<KIND>(a, ri, n0_out) + cb) = rres[it];
          if (a.rb_out)
            *reinterpret_cast<uint2*>(a.rb_out + ri.out_row * a.ldo + n0_out + cb / 4) =
                make_uint2(pack_bf16x2(rres[it].x, rres[it].y), pack_bf16x2(rres[it].z, rres[it].w));
        }
      }
      if (a.rb_out) {   // reduce every row's LPR lanes; lane 0 of each row group stores the tile's sum
#pragma unroll
        for (int off = O::LPR / 2; off > 0; off >>= 1) {
#pragma unroll
          for (int it = 0; it < O::NIT; ++it) sq[it] += __shfl_xor_sync(0xffffffffu, sq[it], off);
        }
#pragma unroll
        for (int it = 0; it < O::NIT; ++it)
          if (ok[it] && (lane % O::LPR) == 0) a.ss_out[orow[it] * a.ss_ld + ty] = sq[it];
      }
    } else {
#pragma unroll
      for (int it = 0; it < O::NIT; ++it) {
        const int row = row0 + it * O::RPI + sub_row;
        const RowInfo ri = row_info_p<KIND, PERSIST>(a, tx, row);
        if (ri.valid) {
          char* dst = out_row_ptr<KIND>(a, ri, n0_out) + cb;
          *reinterpret_cast<uint4*>(dst) = lds128u(smem_u32(stage) + row * O::STRIDE + cb);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- kernel
template <int KIND, int BN, bool DEEP>
__global__ void __launch_bounds__(GEMM_THREADS, DEEP ? 1 : 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmAw,
                                                      const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  using Cfg = TileCfg<BN, DEEP>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_c0 = reinterpret_cast<float*>(sB + STAGES * Cfg::B_BYTES + 256);   // per-column constants of this tile
  float* s_c1 = s_c0 + 128;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  PROF_DECL();
  PROF_BEGIN(1000 + 100 * KIND + BN / 8);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) PROF_MARK(1);   // prologue done

  if (warp == 0) {
    // ---------------- TMA producer.  The first ring is filled by lanes 0..npre-1 in parallel (one stage per lane:
    // issuing a TMA costs ~100+ cycles, and these loads sit on the critical path of a short-K GEMM); lane 0 then
    // runs the steady-state loop.
    int nvalid = 1;
    if constexpr (KindTraits<KIND>::gather) {
      nvalid = a.M - blockIdx.x * a.G;
      nvalid = nvalid > a.G ? a.G : nvalid;
    }
    const uint32_t a_bytes = KindTraits<KIND>::gather ? (uint32_t)(nvalid * a.R * 128) : (uint32_t)Cfg::A_BYTES;
    const int b_row = (KIND == G_CONV1) ? 0 : blockIdx.y * BN;
    const int kz = (KIND == G_PARTIAL) ? blockIdx.z * a.nk * 64 : 0;   // first K element of this CTA's slice
    const int npre = a.nk < STAGES ? a.nk : STAGES;
    // The weight tiles do not depend on the previous kernel: put the first ring of them in flight, then wait for
    // the predecessor grid before touching activations.
    if (lane < npre) {
      mbar_expect_tx(&full[lane], a_bytes + Cfg::B_BYTES);
      tma_load_2d(sB + lane * Cfg::B_BYTES, &tmB, &full[lane], kz + lane * 64, b_row);
    }
    pdl_wait();
    if (lane == 0) PROF_MARK(2);        // predecessor grid complete
    // Gather tiles whose G streams sit in consecutive slots load all of them with ONE wide box per K step
    // (tmAw: same view, box spans G slots); otherwise one box per stream.
    bool wide = false;
    int slot0 = 0;
    if constexpr (KindTraits<KIND>::gather) {
      slot0 = a.slots[blockIdx.x * a.G];
      wide = (nvalid == a.G);
      for (int g = 1; g < nvalid; ++g) wide = wide && (a.slots[blockIdx.x * a.G + g] == slot0 + g);
    }
    auto load_a = [&](int it, int s) {
      uint8_t* dA = sA + s * Cfg::A_BYTES;
      if constexpr (KIND == G_CONV0) {
        // K iteration = kernel row kt; box = 64 mel bins x R frames starting at feature row kt
        if (wide) tma_load_3d(dA, &tmAw, &full[s], 0, it, slot0);
        else
          for (int g = 0; g < nvalid; ++g)
            tma_load_3d(dA + g * a.R * 128, &tmA, &full[s], 0, it, a.slots[blockIdx.x * a.G + g]);
      } else if constexpr (KIND == G_CONV1) {
        // K iteration = (kt, 64-wide piece kc of the 12-position x 32-channel window at f0 = 2*blockIdx.y).
        // x1 is viewed as [slot][row triple][row in triple][44*32]; output frame t reads row 3t + kt, so the
        // box walks R consecutive triples starting at kt/3 with the in-triple row fixed to kt%3.
        const int kt = it / 6, kc = it - kt * 6;
        if (wide) tma_load_4d(dA, &tmAw, &full[s], blockIdx.y * 64 + kc * 64, kt % 3, kt / 3, slot0);
        else
          for (int g = 0; g < nvalid; ++g)
            tma_load_4d(dA + g * a.R * 128, &tmA, &full[s], blockIdx.y * 64 + kc * 64, kt % 3, kt / 3,
                        a.slots[blockIdx.x * a.G + g]);
      } else if constexpr (KIND == G_KV) {
        if (wide) tma_load_3d(dA, &tmAw, &full[s], it * 64, 0, slot0);
        else
          for (int g = 0; g < nvalid; ++g)
            tma_load_3d(dA + g * a.R * 128, &tmA, &full[s], it * 64, 0, a.slots[blockIdx.x * a.G + g]);
      } else {
        tma_load_2d(dA, &tmA, &full[s], kz + it * 64,
                    (KIND == G_VATT) ? blockIdx.x * a.G * a.R : blockIdx.x * 128);
      }
    };
    if (lane < npre) load_a(lane, lane);
    __syncwarp();
    // steady state (ring wraps): warp-uniform loop, one elected lane issues (operands stay in uniform registers)
    for (int it = npre; it < a.nk; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], a_bytes + Cfg::B_BYTES);
        tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full[s], kz + it * 64, b_row);
        load_a(it, s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp walks the K loop (warp-uniform control flow and descriptors), one
    // elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(BN);
    const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
    for (int it = 0; it < a.nk; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full[s], ph);
      if (it == 0 && lane == 0) PROF_MARK(3);      // first operand stage landed
      tc_fence_after();
      const uint64_t da = make_sw128_desc(sA_u32 + s * Cfg::A_BYTES);
      const uint64_t db = make_sw128_desc(sB_u32 + s * Cfg::B_BYTES);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle atom
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);      // frees the smem stage once these MMAs have read it
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(tmem_full);        // accumulator complete
    __syncwarp();
  } else {
    // ---------------- epilogue: warps 2..9; warp w owns TMEM lanes 32*(w%4) .. +31, column half (w-2)/4
    const int q = warp & 3, hf = (warp - 2) >> 2;
    stage_constants<KIND, BN>(a, s_c0, s_c1, threadIdx.x - 64);
    bar_epilogue();                                  // the eight epilogue warps only
    pdl_wait();
    epilogue<KIND, BN>(a, tmem_base + (static_cast<uint32_t>(q * 32) << 16), q, hf, lane, reinterpret_cast<char*>(sA), s_c0,
                       s_c1, tmem_full);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  PROF_END();
}


// ---------------------------------------------------------------------------------------------- persistent kernel
// Large-batch form of the dense kinds (thousands of rows).  At that size the GEMMs of this model are bound by the
// L2 -> SM operand traffic (~6300 B/clk for the whole chip, ~80 GB/s per SM when all 148 pull at once), not by the
// tensor pipe: a 128 x 128 x 64 step needs 32 KB of operands for 2.1 MFLOP.  Three things raise the FLOP per L2 byte
// and hide the rest:
//  * NSUB = 2: the tile is two adjacent 128-column weight tiles wide (MMA N = 256), so one A tile feeds twice the math;
//  * PAIR: a 2-CTA cluster (cta_group::2, MMA M = 256) in which each CTA loads its own 128 rows of A and only ONE of
//    the two 128-row weight tiles; the tensor cores of both SMs read both halves (131 FLOP per L2 byte instead of 64);
//  * persistent CTAs (one per SM) walking tiles t = first, + stride, ... with the N tile fastest; the operand ring runs
//    across tile boundaries and the accumulator is double buffered in TMEM, so the TMA warp prefetches and the MMA warp
//    computes tile i + 1 while the eight epilogue warps drain tile i.
// The epilogue code is the one-tile kernel's, called once per 128-column sub-tile.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void pair_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-CTA forms: the load lands in this CTA's shared memory and completes bytes on the LEADER's barrier; the MMA is
// issued by the leader for both CTAs; the commit arrives on the barrier at the same offset in every CTA of the mask.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int KIND, int NSUB, bool PAIR>
struct PersistCfg {
  static constexpr int BN = 128;                                      // sub-tile width (the epilogue's tile)
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_ROWS = PAIR ? (NSUB * BN) / 2 : NSUB * BN;   // weight rows this CTA loads per K step
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = ((128 * OutCfg<KIND, BN>::STRIDE + 1023) / 1024) * 1024;   // epilogue staging
  // Epilogue groups of 8 warps; group g drains accumulator buffer g.  One sub-tile epilogue is a chain of latencies
  // (TMEM load, activation math, staging, barrier, stores: ~2 us), longer than the sub-tile's MMAs, so two of them run
  // interleaved where the staging area is small enough to have two.
  static constexpr int NGRP = STAGE_BYTES <= 20 * 1024 ? 2 : 1;
  static constexpr int THREADS = 64 + NGRP * EPI_THREADS;
  static constexpr int STAGES = (196 * 1024 - NGRP * STAGE_BYTES) / (A_BYTES + B_BYTES);
  static constexpr int ACC_COLS = NSUB * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + NGRP * STAGE_BYTES + 512 + NGRP * 1024 + 1024;
  static_assert(NSUB == 1 || NSUB == 2, "one or two weight tiles per tile");
  static_assert(!PAIR || NSUB == 2, "the pair form loads one 128-row weight tile per CTA");
  static_assert(STAGES >= 3, "tile does not fit");
};

// grid: a multiple of the cluster size; unit = CTA (or CTA pair); tiles_y counts NSUB-wide column tiles; units walk
// ntiles = (row tiles or row-tile pairs) x tiles_y.
template <int KIND, int NSUB, bool PAIR>
__global__ void __launch_bounds__(PersistCfg<KIND, NSUB, PAIR>::THREADS, 1) gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                          const __grid_constant__ CUtensorMap tmB,
                                                                          const GemmArgs a, int tiles_y, int ntiles) {
  static_assert(!KindTraits<KIND>::stream_rows && KIND != G_DECODER, "dense kinds only");
  using Cfg = PersistCfg<KIND, NSUB, PAIR>;
  constexpr int BN = Cfg::BN;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  char* stage = reinterpret_cast<char*>(sB + STAGES * Cfg::B_BYTES);            // [NGRP][STAGE_BYTES]
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + Cfg::NGRP * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_cbase = reinterpret_cast<float*>(stage + Cfg::NGRP * Cfg::STAGE_BYTES + 512);   // [NGRP][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const bool leader = rank == 0;
  const int unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int nunits = PAIR ? gridDim.x >> 1 : gridDim.x;

  PROF_DECL();
  PROF_BEGIN(1000 + 100 * KIND + BN / 8);
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], PAIR ? 16 : 8);   // one arrive per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
  }
  tc_fence_before();
  if constexpr (PAIR) pair_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) PROF_MARK(1);

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs of a pair): warp-uniform loop, one elected lane issues; the ring does
    // not drain between tiles
    int cnt = 0;
    bool waited = false;
    for (int t = unit; t < ntiles; t += nunits) {
      const int tu = t / tiles_y, ty = t - tu * tiles_y;
      const int tx = PAIR ? 2 * tu + (int)rank : tu;                               // this CTA's 128-row tile
      for (int it = 0; it < a.nk; ++it, ++cnt) {
        const int s = cnt % STAGES;
        const uint32_t ph = (cnt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if constexpr (PAIR) {
          const uint32_t lbar = map_to_rank(smem_u32(&full[s]), 0);
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(&full[s], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
            tma_load_2d_pair(sB + s * Cfg::B_BYTES, &tmB, lbar, it * 64, (ty * NSUB + (int)rank) * BN);
          }
          __syncwarp();
          if (!waited) {   // activations come from the predecessor grid
            pdl_wait();
            waited = true;
          }
          if (elect_one_sync()) tma_load_2d_pair(sA + s * Cfg::A_BYTES, &tmA, lbar, it * 64, tx * 128);
          __syncwarp();
        } else {
          if (elect_one_sync()) {
            mbar_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
            for (int j = 0; j < NSUB; ++j)   // weights: no dependency
              tma_load_2d(sB + s * Cfg::B_BYTES + j * BN * 128, &tmB, &full[s], it * 64, (ty * NSUB + j) * BN);
          }
          __syncwarp();
          if (!waited) {
            pdl_wait();
            waited = true;
            if (lane == 0) PROF_MARK(2);
          }
          if (elect_one_sync()) tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, &full[s], it * 64, tx * 128);
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (the leader CTA of a pair)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(PAIR ? 256 : 128, NSUB * BN);
      const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
      int cnt = 0, i = 0;
      for (int t = unit; t < ntiles; t += nunits, ++i) {
        const int buf = i & 1;
        mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);   // the epilogue of tile i - 2 has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * Cfg::ACC_COLS;
        for (int it = 0; it < a.nk; ++it, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&full[s], ph);
          if (cnt == 0 && lane == 0) PROF_MARK(3);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(sA_u32 + s * Cfg::A_BYTES);
          const uint64_t db = make_sw128_desc(sB_u32 + s * Cfg::B_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (PAIR) umma_bf16_pair(acc, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
              else umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (PAIR) umma_commit_pair(&empty[s]);
            else umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one_sync()) {
          if constexpr (PAIR) umma_commit_pair(&tmem_full[buf]);
          else umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue warps: group g = warps 2 + 8g .. 9 + 8g takes the tiles whose accumulator buffer is g
    // (every tile when there is one group); one pass of the shared epilogue per 128-column sub-tile
    const int g = (warp - 2) >> 3;
    const int q = warp & 3, hf = ((warp - 2) >> 2) & 1;
    const int et = (threadIdx.x - 64) & (EPI_THREADS - 1);
    char* my_stage = stage + g * Cfg::STAGE_BYTES;
    float* s_c0 = s_cbase + g * 256;
    float* s_c1 = s_c0 + 128;
    bool first = true;
    int i = 0;
    for (int t = unit; t < ntiles; t += nunits, ++i) {
      const int buf = i & 1;
      if (Cfg::NGRP == 2 && buf != g) continue;
      const int tu = t / tiles_y, ty = t - tu * tiles_y;
      const int tx = PAIR ? 2 * tu + (int)rank : tu;
      const uint32_t empty_addr = PAIR ? map_to_rank(smem_u32(&tmem_empty[buf]), 0) : smem_u32(&tmem_empty[buf]);
      // per-column constants (bias) of both sub-tiles: weights, loaded here so that their L2 round trip overlaps the
      // row-scale loads and the wait for the accumulator instead of sitting in front of every sub-tile
      float cpre[NSUB];
#pragma unroll
      for (int j = 0; j < NSUB; ++j)
        cpre[j] = (KIND != G_PARTIAL && et < BN && a.bias) ? __ldg(a.bias + (ty * NSUB + j) * BN + et) : 0.f;
      if (first) pdl_wait();
      first = false;
      // row scale of the folded RMSNorm: one value per row, shared by the sub-tiles
      float rs = 1.f;
      if constexpr (KIND == G_SWIGLU || KIND == G_GLU || KIND == G_STORE_F32) {
        if (a.ss) {
          const RowInfo rme = row_info_p<KIND, true>(a, tx, q * 32 + lane);
          if (rme.valid) {
            const float* sp = a.ss + rme.out_row * a.ss_ld;
            float tsum = 0.f;
            for (int k = 0; k < a.ss_tiles; ++k) tsum += sp[k];
            rs = 1.0f / (sqrtf(tsum) * 0.05103103630798288f + 1e-8f);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        // every warp of the group is past phase 1 of the previous sub-tile (its mid barrier), so the constants may
        // change; the barrier below also orders the previous phase-2 reads of the staging area before this phase 1
        if (KIND != G_PARTIAL && et < BN) s_c0[et] = cpre[j];
        bar_epilogue_p(1 + g);
        epilogue_p<KIND, BN, true>(a, tmem_base + buf * Cfg::ACC_COLS + j * BN + (static_cast<uint32_t>(q * 32) << 16), q, hf, lane,
                           my_stage, s_c0, s_c1, &tmem_full[buf], tx, ty * NSUB + j, (i >> 1) & 1,
                           j == NSUB - 1 ? empty_addr : 0u, 1 + g, &rs);
      }
    }
  }
  tc_fence_before();
  if constexpr (PAIR) pair_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  PROF_END();
}

// Launch with (optional) programmatic stream serialization.
template <typename Kern, typename... Args>
inline cudaError_t launch_kernel(Kern kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Must be called once per instantiation (outside stream capture) before the first launch.
template <int KIND, int BN>
inline cudaError_t configure_gemm_tc() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<KIND, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       TileCfg<BN, true>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_tc_kernel<KIND, BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TileCfg<BN, false>::SMEM_BYTES);
}

template <int KIND, int NSUB, bool PAIR>
inline cudaError_t configure_gemm_tc_persist() {
  return cudaFuncSetAttribute(gemm_tc_persist_kernel<KIND, NSUB, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              PersistCfg<KIND, NSUB, PAIR>::SMEM_BYTES);
}

// m_tiles / n_tiles count 128 x 128 tiles; ctas = CTAs to launch at most (the SM count)
template <int KIND, int NSUB, bool PAIR>
inline cudaError_t launch_gemm_tc_persist(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a,
                                          int m_tiles, int n_tiles, bool pdl, int ctas) {
  const int tiles_y = n_tiles / NSUB;
  const int ntiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * tiles_y;
  // as few CTAs (CTA pairs) as reach the same number of rounds: 480 tiles on 148 SMs take 4 rounds whether 148 or 120 CTAs
  // walk them, and the 28 SMs left free serve the other lane's kernels meanwhile
  const int units_max = PAIR ? ctas / 2 : ctas;
  const int rounds = (ntiles + units_max - 1) / units_max;
  const int units = (ntiles + rounds - 1) / rounds;
  int grid = PAIR ? 2 * units : units;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(PersistCfg<KIND, NSUB, PAIR>::THREADS);
  cfg.dynamicSmemBytes = PersistCfg<KIND, NSUB, PAIR>::SMEM_BYTES;
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
  return cudaLaunchKernelEx(&cfg, gemm_tc_persist_kernel<KIND, NSUB, PAIR>, tmA, tmB, a, tiles_y, ntiles);
}

// deep = one CTA per SM, whole smem as the ring (grids that fit in one wave of single CTAs); otherwise two CTAs per SM
// (a grid of 149 .. 296 CTAs then still runs in one round instead of one full round plus a nearly empty one).
template <int KIND, int BN>
inline cudaError_t launch_gemm_tc(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmAw,
                                  const CUtensorMap& tmB, const GemmArgs& a, int m_tiles, int n_tiles, bool pdl,
                                  int num_sms, int splits = 1) {
  const dim3 grid(m_tiles, n_tiles, splits);
  if (m_tiles * n_tiles * splits <= num_sms)
    return launch_kernel(gemm_tc_kernel<KIND, BN, true>, grid, dim3(GEMM_THREADS), TileCfg<BN, true>::SMEM_BYTES, st, pdl, tmA,
                         tmAw, tmB, a);
  return launch_kernel(gemm_tc_kernel<KIND, BN, false>, grid, dim3(GEMM_THREADS), TileCfg<BN, false>::SMEM_BYTES, st, pdl, tmA,
                       tmAw, tmB, a);
}

}  // namespace tone
