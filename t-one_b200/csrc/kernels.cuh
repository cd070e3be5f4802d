// Bandwidth-class kernels of the streaming step (everything that is not a GEMM):
//   begin_step_kernel   PCM -> log-mel -> RMSNorm(64) rows, plus the start-of-step cache rolls       (A1, K1-K3, K13, K21)
//   norm_kernel         RMSNorm(384) variants feeding the GEMMs, optional cache-row scatter          (A3, K7, K13)
//   attention_kernel    per-head LayerNorm + RoPE + scores + masked softmax + P.V, score sharing     (A5, K11)
//   dwconv_kernel       causal depthwise conv k=31 + BN + SiLU with its 30-frame cache               (A4.3, K15)
//   reduction_dw_kernel / upsample_norm_kernel                                                       (K17, K18)
// Section/kernel numbers refer to SURVEY.md Appendix A and §2b.
#pragma once

#include "common.cuh"

namespace tone {

constexpr int D_MODEL = 384;
constexpr int N_HEADS = 8;
constexpr int D_HEAD = 48;
constexpr int N_MELS = 64;
constexpr int N_BINS = 81;
constexpr int WIN = 160;
constexpr int HOP = 80;
constexpr int MHSA_S = 30;
constexpr int CONV_S = 30;
constexpr int SUB1_ROWS = 10;
constexpr int SUB2_ROWS = 8;
constexpr int X1_ROW = 44 * 32;      // elements per conv0-output time row (f-major, channel-minor)
constexpr int FEAT_ROWS_MAX = 50;    // 10 cached + up to 40 new feature rows per slot
constexpr int X1_ROWS_MAX = 48;      // 8 cached + up to 40 new rows per slot
constexpr int KV_ROWS_MAX = 44;      // 30 cached + up to 13 new rows (+1 pad) per stateful layer
constexpr int MAX_FRAMES = 40;       // 400 ms
constexpr int MAX_T = 13;

// ------------------------------------------------------------------------------------------------ begin step
struct BeginArgs {
  const int* slots;          // [B]
  const void* pcm;           // [B][C] int32 or int16 samples (pcm_fmt: 0 | 1), int16 range
  int pcm_fmt;
  const __half* feats_in;    // nullable: feature-input mode (reference skip_preprocessor=True, tone/nn/model.py:151-160):
                             // [B][64][F] fp16 log-mel features replace the waveform front end; `pre` is left untouched
  __half* pre;               // [slots][80] last samples of the previous chunk (fp16-exact values)
  bf16* feat;                // [slots][FEAT_ROWS_MAX][64]
  bf16* x1;                  // [slots][X1_ROWS_MAX][X1_ROW]
  bf16* kv14;                // [slots][KV_ROWS_MAX][384]   layer-14 [cache | new] rows (15 + T2)
  bf16* kv15;                // [slots][KV_ROWS_MAX][384]   layer-15 [cache | new] rows (30 + T)
  int* mhsa_len;             // [slots]
  int* len_in;               // [B] length seen by this step's masks (value entering the step)
  int* cpos;                 // [slots][2] ring positions of the depthwise-conv caches (full / reduced rate)
  int* cpos_in;              // [B][2] positions seen by this step
  const __half* basis;       // [2][168][168] fp16 hi / lo split of the fused pre-emphasis * Hann * DFT basis, row n =
                             // cos bin n (n < 81) / -sin bin n-81, k contiguous, zero padded
  const int* mel_start;      // [65] CSR over mel filters
  const unsigned char* mel_bin;  // [nnz]
  const float* mel_w;        // [nnz]
  const float* pre_norm_g;   // [64]
  int C, F, T, T2;
  int B;                     // streams in the batch
};

// One CTA per stream, 352 threads (11 warps).  The framed DFT runs on the tensor cores as an fp16 GEMM with fp32
// accumulation: the waveform IS fp16 (the reference quantises it, model.py:165), so the A operand is exact, and the
// fused (pre-emphasis x Hann x DFT) basis is split into an fp16 high and low part (b = hi + lo, ~22 bits), i.e. two
// MMAs per tile: spec[f][n] = sum_k u[80 f + k] * (hi[n][k] + lo[n][k]).
constexpr int BEGIN_THREADS = 352;
constexpr int BASIS_N = 168;             // 162 columns (81 cos + 81 sin) padded to 21 n-tiles of 8
constexpr int BASIS_LD = 168;            // halfs per basis row: 160 k + 8 pad (conflict-free fragment loads)

__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Persistent: grid = min(B, SM count) CTAs, each walks streams b = blockIdx.x, + gridDim.x, ... with the DFT basis
// staged ONCE.  The start-of-step rolls of the large per-stream caches (x1 8 rows, layer-15 / layer-14 attention rows:
// 57 KB per stream, contiguous blocks) go through a shared-memory staging buffer as bulk async copies issued by one
// thread: the loads of stream i travel under its front-end arithmetic and the stores drain under stream i + 1's.
constexpr int ROLL_X1_BYTES = SUB2_ROWS * X1_ROW * 2;            // 22,528
constexpr int ROLL_KV15_BYTES = MHSA_S * D_MODEL * 2;            // 23,040
constexpr int ROLL_KV14_BYTES = (MHSA_S / 2) * D_MODEL * 2;      // 11,520
constexpr int ROLL_BYTES = ROLL_X1_BYTES + ROLL_KV15_BYTES + ROLL_KV14_BYTES;

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(BEGIN_THREADS) begin_step_kernel(const BeginArgs a) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  PROF_DECL();
  PROF_BEGIN(1);
  pdl_launch_dependents();
  const int F = a.F, C = a.C;
  const int n_mt = (F + 15) / 16;                       // frame tiles of 16
  const int UH = 16 * n_mt * HOP + HOP + 16;            // halfs: frames of the padded tiles stay in bounds
  __half* basisS = reinterpret_cast<__half*>(sm_raw);   // [2][BASIS_N][BASIS_LD] hi, lo
  unsigned char* rollS = sm_raw + 2 * BASIS_N * BASIS_LD * 2;   // [ROLL_BYTES] x1 | kv15 | kv14 rows in flight
  __half* uh = reinterpret_cast<__half*>(rollS + ROLL_BYTES);   // [UH] carried 80 samples + chunk, zero tail
  float* spec = reinterpret_cast<float*>(uh + ((UH + 7) & ~7));   // [F][162]
  float* featS = spec + F * 162;                        // [F][64]
  unsigned char* pcmS = reinterpret_cast<unsigned char*>(featS + F * N_MELS);   // raw PCM of the NEXT stream | its 80 carried samples
  const int tid = threadIdx.x;
  // constant basis -> smem by ONE bulk async copy (113 KB), issued before the PDL wait and awaited just before the DFT;
  // the mel filterbank (CSR, ~1 KB) is staged the same way so that the mel loop does not chase indices in global memory
  __shared__ uint64_t basis_bar, roll_bar, pcm_bar;
  __shared__ int mel_startS[N_MELS + 1];
  __shared__ unsigned char mel_binS[256];
  __shared__ float mel_wS[256];
  constexpr uint32_t BASIS_BYTES = 2 * BASIS_N * BASIS_LD * 2;
  static_assert(BASIS_BYTES % 16 == 0 && ROLL_X1_BYTES % 16 == 0 && ROLL_KV15_BYTES % 16 == 0 && ROLL_KV14_BYTES % 16 == 0,
                "bulk copy granularity");
  if (tid == 0) {
    mbar_init(&basis_bar, 1);
    mbar_init(&roll_bar, 1);
    mbar_init(&pcm_bar, 1);
    fence_mbar_init();
    mbar_expect_tx(&basis_bar, BASIS_BYTES);
    bulk_load_1d(basisS, a.basis, BASIS_BYTES, &basis_bar);
  }
  if (tid <= N_MELS) mel_startS[tid] = a.mel_start[tid];
  {
    const int nnz = a.mel_start[N_MELS];
    for (int i = tid; i < nnz && i < 256; i += BEGIN_THREADS) {
      mel_binS[i] = a.mel_bin[i];
      mel_wS[i] = a.mel_w[i];
    }
  }
  __syncthreads();                                       // barrier inits visible before anybody waits on them
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);

  // The waveform of a stream (its chunk of the batch's PCM, contiguous, and the 80 samples carried in the slot) is
  // brought into shared memory by bulk async copies one stream ahead, so an iteration starts on data that is already
  // on the SM instead of a slot-id -> address -> samples chain of global round trips.
  const uint32_t pcm_bytes = (uint32_t)C * (a.pcm_fmt ? 2u : 4u);
  auto fetch_pcm = [&](int b) {      // tid 0
    mbar_expect_tx(&pcm_bar, pcm_bytes + HOP * 2);
    bulk_load_1d(pcmS, reinterpret_cast<const unsigned char*>(a.pcm) + (size_t)b * pcm_bytes, pcm_bytes, &pcm_bar);
    bulk_load_1d(pcmS + pcm_bytes, a.pre + (size_t)a.slots[b] * HOP, HOP * 2, &pcm_bar);
  };
  if (tid == 0 && !a.feats_in && (int)blockIdx.x < a.B) fetch_pcm(blockIdx.x);
  int iter = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
  const int slot = a.slots[b];
  bf16* x1p = a.x1 + (size_t)slot * X1_ROWS_MAX * X1_ROW;
  bf16* k15p = a.kv15 + (size_t)slot * KV_ROWS_MAX * D_MODEL;
  bf16* k14p = a.kv14 + (size_t)slot * KV_ROWS_MAX * D_MODEL;
  // ---- start-of-step cache rolls: the last rows of [cache | previous new rows] become the cache
  if (tid == 0) {
    // x1 rows [F, F+8) -> [0, 8); layer-15 rows [T, T+30) -> [0, 30); layer-14 rows [T2, T2+15) -> [0, 15)
    if (iter) bulk_wait_read0();                         // the previous stream's stores have left the staging buffer
    mbar_expect_tx(&roll_bar, ROLL_BYTES);
    bulk_load_1d(rollS, x1p + (size_t)F * X1_ROW, ROLL_X1_BYTES, &roll_bar);
    bulk_load_1d(rollS + ROLL_X1_BYTES, k15p + (size_t)a.T * D_MODEL, ROLL_KV15_BYTES, &roll_bar);
    bulk_load_1d(rollS + ROLL_X1_BYTES + ROLL_KV15_BYTES, k14p + (size_t)a.T2 * D_MODEL, ROLL_KV14_BYTES, &roll_bar);
  }
  {
    // feature rows [F, F+10) -> [0, 10)  (80 x 16 B): disjoint ranges
    uint4* f4 = reinterpret_cast<uint4*>(a.feat + (size_t)slot * FEAT_ROWS_MAX * N_MELS);
    if (tid < SUB1_ROWS * N_MELS / 8) f4[tid] = f4[F * N_MELS / 8 + tid];
    // ---- waveform: int -> /32767 -> fp16 (model.py:164-165), prefixed by the carried 80 samples (feats.py:129-133)
    __half* pre = a.pre + (size_t)slot * HOP;
    if (!a.feats_in) {
      mbar_wait(&pcm_bar, iter & 1);
      const int* pcm32 = reinterpret_cast<const int*>(pcmS);
      const short* pcm16 = reinterpret_cast<const short*>(pcmS);
      const __half* preS = reinterpret_cast<const __half*>(pcmS + pcm_bytes);
      for (int i = tid; i < UH; i += BEGIN_THREADS) {
        __half v = __float2half_rn(0.f);
        if (i < HOP) v = preS[i];
        else if (i < C + HOP)
          v = __float2half_rn(static_cast<float>(a.pcm_fmt ? (int)pcm16[i - HOP] : pcm32[i - HOP]) / 32767.0f);
        uh[i] = v;
      }
    }
    __syncthreads();
    if (tid == 0 && !a.feats_in && b + (int)gridDim.x < a.B) fetch_pcm(b + gridDim.x);   // the staging buffer is free again
    if (tid == 0) {
      int len = a.mhsa_len[slot];
      a.len_in[b] = len;
      a.mhsa_len[slot] = min(len + a.T, MHSA_S);   // conformer_blocks.py:191
      const int p0 = a.cpos[2 * slot], p1 = a.cpos[2 * slot + 1];
      a.cpos_in[2 * b] = p0;
      a.cpos_in[2 * b + 1] = p1;
      a.cpos[2 * slot] = (p0 + a.T) % CONV_S;      // the T (T2) oldest cache rows are replaced by this step's rows
      a.cpos[2 * slot + 1] = (p1 + a.T2) % CONV_S;
    }
    if (!a.feats_in)
      for (int i = tid; i < HOP; i += BEGIN_THREADS) pre[i] = uh[C + i];
  }
  if (threadIdx.x == 0) PROF_MARK(1);

  if (a.feats_in) {   // features come from outside: [64][F] fp16 -> featS[f][m]
    const __half* fi = a.feats_in + (size_t)b * N_MELS * F;
    for (int i = tid; i < F * N_MELS; i += BEGIN_THREADS) {
      const int m = i / F, f = i - m * F;
      featS[f * N_MELS + m] = __half2float(fi[i]);
    }
    __syncthreads();
  }
  // ---- framed DFT on the tensor cores: warp w owns n-tiles w, w+11 (21 tiles of 8 columns)
  mbar_wait(&basis_bar, 0);
  if (!a.feats_in) {
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    for (int nt = warp; nt < BASIS_N / 8; nt += BEGIN_THREADS / 32) {
      const __half* bh = basisS + (nt * 8 + g) * BASIS_LD + 2 * t;
      const __half* bl = bh + BASIS_N * BASIS_LD;
      for (int mt = 0; mt < n_mt; ++mt) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const __half* ua = uh + (mt * 16 + g) * HOP + 2 * t;
#pragma unroll
        for (int ks = 0; ks < WIN / 16; ++ks) {
          uint32_t af[4], bhf[2], blf[2];
          af[0] = *reinterpret_cast<const uint32_t*>(ua + ks * 16);
          af[1] = *reinterpret_cast<const uint32_t*>(ua + 8 * HOP + ks * 16);
          af[2] = *reinterpret_cast<const uint32_t*>(ua + ks * 16 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(ua + 8 * HOP + ks * 16 + 8);
          bhf[0] = *reinterpret_cast<const uint32_t*>(bh + ks * 16);
          bhf[1] = *reinterpret_cast<const uint32_t*>(bh + ks * 16 + 8);
          blf[0] = *reinterpret_cast<const uint32_t*>(bl + ks * 16);
          blf[1] = *reinterpret_cast<const uint32_t*>(bl + ks * 16 + 8);
          mma_f16_16x8x16(acc, af, bhf);
          mma_f16_16x8x16(acc, af, blf);
        }
        const int n = nt * 8 + 2 * t, f0 = mt * 16 + g;
        if (n < 162) {
          if (f0 < F) {
            spec[f0 * 162 + n] = acc[0];
            spec[f0 * 162 + n + 1] = acc[1];
          }
          if (f0 + 8 < F) {
            spec[(f0 + 8) * 162 + n] = acc[2];
            spec[(f0 + 8) * 162 + n + 1] = acc[3];
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) PROF_MARK(3);
  // ---- power -> mel -> log (feats.py:99-101)
  for (int i = tid; i < (a.feats_in ? 0 : F * N_MELS); i += BEGIN_THREADS) {
    const int f = i / N_MELS, m = i - f * N_MELS;
    const float* sp = spec + f * 162;
    float e = 0.f;
    for (int p = mel_startS[m]; p < mel_startS[m + 1]; ++p) {
      const int kb = mel_binS[p];
      const float re = sp[kb], im = sp[N_BINS + kb];
      e = fmaf(mel_wS[p], re * re + im * im, e);
    }
    featS[i] = logf(e + 5.9604644775390625e-08f);   // 2^-24
  }
  // the rolled rows have landed in the staging buffer: send them to the front of their buffers
  if (tid == 0) {
    mbar_wait(&roll_bar, iter & 1);
    bulk_store_1d(x1p, rollS, ROLL_X1_BYTES);
    bulk_store_1d(k15p, rollS + ROLL_X1_BYTES, ROLL_KV15_BYTES);
    bulk_store_1d(k14p, rollS + ROLL_X1_BYTES + ROLL_KV15_BYTES, ROLL_KV14_BYTES);
    bulk_commit();
  }
  __syncthreads();
  if (threadIdx.x == 0) PROF_MARK(4);
  // ---- RMSNorm(64) per frame (conformer_blocks.py:632, submodules.py:45-54), bf16 rows behind the 10 cached rows
  bf16* frow = a.feat + ((size_t)slot * FEAT_ROWS_MAX + SUB1_ROWS) * N_MELS;
  const int warp = tid >> 5, lane = tid & 31;
  for (int f = warp; f < F; f += BEGIN_THREADS / 32) {
    const float x0 = featS[f * N_MELS + lane], x1 = featS[f * N_MELS + 32 + lane];
    const float ss = warp_sum(x0 * x0 + x1 * x1);
    const float inv = 1.0f / (sqrtf(ss) * 0.125f + 1e-8f);
    frow[f * N_MELS + lane] = __float2bfloat16(a.pre_norm_g[lane] * (x0 * inv));
    frow[f * N_MELS + 32 + lane] = __float2bfloat16(a.pre_norm_g[32 + lane] * (x1 * inv));
  }
  }
  if (tid == 0) bulk_wait0();                            // all rolled rows are in global memory before the grid completes
  PROF_END();
}

// ------------------------------------------------------------------------------------------------ RMSNorm(384)
// One warp per row.  x = r[row]; if g1: x = g1*x/(rms+eps) and r[row] = x (the layer's norm_out, in place);
// if g2: y = g2*x/(rms+eps) else y = x; n[row] = bf16(y).  Optionally the same bf16 row is scattered into the
// stream's [cache | new] attention rows (layers 14/15 cache the NORMALISED layer input, submodules.py:295-303).
struct NormArgs {
  float* r;
  const float* g1;
  const float* g2;
  bf16* n;
  int M;
  // optional split-K input: x = r + scale * (sum_z part[z][row] + bias) is formed (and written back to r) first
  const __half* part;
  int nsplit;
  long long part_stride;
  const float* bias;
  float scale;
  bf16* kv;            // nullable
  const int* slots;
  int rows_per_stream; // T of this layer
  int kv_row_off;      // S
};

__device__ __forceinline__ float rms_inv_384(const float4 (&x)[3]) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) ss += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
  ss = warp_sum(ss);
  return 1.0f / (sqrtf(ss) * 0.05103103630798288f + 1e-8f);   // 384^-1/2
}

__device__ __forceinline__ void scale_384(float4 (&x)[3], const float* g, float inv, int lane) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i * 128 + lane * 4);
    x[i].x = gg.x * (x[i].x * inv);
    x[i].y = gg.y * (x[i].y * inv);
    x[i].z = gg.z * (x[i].z * inv);
    x[i].w = gg.w * (x[i].w * inv);
  }
}

// gains / bias of this lane's 12 columns, loaded BEFORE the programmatic-dependency wait (they are weights)
struct Vec384 {
  float4 v[3];
};
__device__ __forceinline__ Vec384 load_vec384(const float* g, int lane, float fill) {
  Vec384 o;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o.v[i] = g ? __ldg(reinterpret_cast<const float4*>(g + i * 128 + lane * 4)) : make_float4(fill, fill, fill, fill);
  return o;
}
__device__ __forceinline__ void scale_384(float4 (&x)[3], const Vec384& g, float inv) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x[i].x = g.v[i].x * (x[i].x * inv);
    x[i].y = g.v[i].y * (x[i].y * inv);
    x[i].z = g.v[i].z * (x[i].z * inv);
    x[i].w = g.v[i].w * (x[i].w * inv);
  }
}

// x += scale * (sum_z part[z] + bias): fixed summation order, so split-K stays deterministic.  All partial loads are
// issued before the first add (a runtime-trip-count loop would serialise nsplit L2 round trips: 3.9 us -> 1.x us).
constexpr int MAX_SPLITS = 8;
template <int MAXS>
__device__ __forceinline__ void add_partials_384(float4 (&x)[3], const __half* part_row, int nsplit, long long stride,
                                                 const Vec384& bias, float scale, int lane) {
  uint2 p[MAXS][3];     // 4 halfs each
#pragma unroll
  for (int z = 0; z < MAXS; ++z) {
    if (z < nsplit) {
#pragma unroll
      for (int i = 0; i < 3; ++i) p[z][i] = *reinterpret_cast<const uint2*>(part_row + z * stride + i * 128 + lane * 4);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float4 s = bias.v[i];
#pragma unroll
    for (int z = 0; z < MAXS; ++z) {
      if (z < nsplit) {
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&p[z][i].x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&p[z][i].y));
        s.x += lo.x;
        s.y += lo.y;
        s.z += hi.x;
        s.w += hi.y;
      }
    }
    x[i].x += scale * s.x;
    x[i].y += scale * s.y;
    x[i].z += scale * s.z;
    x[i].w += scale * s.w;
  }
}

// MAXS = compile-time bound of the split-K factor (registers for the partials): 2 for large batches, 8 for small ones
template <int MAXS>
__global__ void __launch_bounds__(256) norm_kernel(const NormArgs a) {
  PROF_DECL();
  PROF_BEGIN(2);
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const Vec384 g1 = load_vec384(a.g1, lane, 1.f), g2 = load_vec384(a.g2, lane, 1.f);
  const Vec384 bias = load_vec384(a.part ? a.bias : nullptr, lane, 0.f);
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  if (row >= a.M) return;
  float* rr = a.r + (size_t)row * D_MODEL;
  float4 x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = *reinterpret_cast<const float4*>(rr + i * 128 + lane * 4);
  if (a.part) add_partials_384<MAXS>(x, a.part + (size_t)row * D_MODEL, a.nsplit, a.part_stride, bias, a.scale, lane);
  if (a.g1) scale_384(x, g1, rms_inv_384(x));
  if (a.part || a.g1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(rr + i * 128 + lane * 4) = x[i];
  }
  if (a.g2) scale_384(x, g2, rms_inv_384(x));
  if (a.n) {
    bf16* nr = a.n + (size_t)row * D_MODEL;
    bf16* kr = nullptr;
    if (a.kv) {
      const int b = row / a.rows_per_stream, t = row - b * a.rows_per_stream;
      kr = a.kv + ((size_t)a.slots[b] * KV_ROWS_MAX + a.kv_row_off + t) * D_MODEL;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      uint2 p = make_uint2(pack_bf16x2(x[i].x, x[i].y), pack_bf16x2(x[i].z, x[i].w));
      *reinterpret_cast<uint2*>(nr + i * 128 + lane * 4) = p;
      if (kr) *reinterpret_cast<uint2*>(kr + i * 128 + lane * 4) = p;
    }
  }
  PROF_END();
}

// After layer 14: r_full[b,t] = (t < 2*T2 ? norm_out14(r_red[b, t/2]) : 0) + r_full[b,t]  (conformer_blocks.py:955-988),
// followed by layer 15's first RMSNorm -> n (bf16).  One warp per full-rate row.
struct UpsampleArgs {
  float* r_full;
  const float* r_red;
  const __half* part;     // split-K output of layer 14's second feed-forward (added to r_red rows first)
  int nsplit;
  long long part_stride;
  const float* bias;
  float scale;
  const float* g_out;   // norm_out of layer 14
  const float* g_next;  // norm_feed_forward1 of layer 15
  bf16* n;
  int B, T, T2;
};

template <int MAXS>
__global__ void __launch_bounds__(256) upsample_norm_kernel(const UpsampleArgs a) {
  PROF_DECL();
  PROF_BEGIN(3);
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const Vec384 g_out = load_vec384(a.g_out, lane, 1.f), g_next = load_vec384(a.g_next, lane, 1.f);
  const Vec384 bias = load_vec384(a.part ? a.bias : nullptr, lane, 0.f);
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  if (row >= a.B * a.T) return;
  const int b = row / a.T, t = row - b * a.T;
  float* rr = a.r_full + (size_t)row * D_MODEL;
  float4 x[3], y[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = *reinterpret_cast<const float4*>(rr + i * 128 + lane * 4);
  if (t < 2 * a.T2) {
    const float* rs = a.r_red + ((size_t)b * a.T2 + (t >> 1)) * D_MODEL;
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = *reinterpret_cast<const float4*>(rs + i * 128 + lane * 4);
    if (a.part)
      add_partials_384<MAXS>(x, a.part + ((size_t)b * a.T2 + (t >> 1)) * D_MODEL, a.nsplit, a.part_stride, bias,
                             a.scale, lane);
    scale_384(x, g_out, rms_inv_384(x));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      y[i].x += x[i].x;
      y[i].y += x[i].y;
      y[i].z += x[i].z;
      y[i].w += x[i].w;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(rr + i * 128 + lane * 4) = y[i];
  scale_384(y, g_next, rms_inv_384(y));
  bf16* nr = a.n + (size_t)row * D_MODEL;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    *reinterpret_cast<uint2*>(nr + i * 128 + lane * 4) =
        make_uint2(pack_bf16x2(y[i].x, y[i].y), pack_bf16x2(y[i].z, y[i].w));
  PROF_END();
}

// ------------------------------------------------------------------------------------------------ attention core
struct AttnArgs {
  const float* q; int ldq;    // [B*T][ldq]      (used when recompute)
  const float* k; int ldk;    // [B*Tk][ldk]
  const float* v; int ldv;    // [B*Tk][ldv]
  float* P;                   // [B][8][T][Tk] probabilities shared by the following non-recompute layers
  bf16* ctx;                  // [B*T][384]
  const float* q_ln_w; const float* q_ln_b;
  const float* k_ln_w; const float* k_ln_b;
  const float* rope_cos;      // [MHSA_S + MAX_T][16], position p at row p + MHSA_S
  const float* rope_sin;
  const int* len_in;          // [B]
  int T, Tk, S;
  int recompute;
  int mask_mode;              // 0 none, 1: off = 30 - len (layer 15), 2: off = (30 - len) / 2 (layer 14)
};

__device__ __forceinline__ void ln_rope_row(const float* src, const float* w, const float* bia, const float* cs,
                                            const float* sn, float scale, float* dst) {
  float x[D_HEAD];
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(src + i);
    x[i] = t.x; x[i + 1] = t.y; x[i + 2] = t.z; x[i + 3] = t.w;
  }
  // four partial sums: 12-long dependent chains instead of 48
  float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
    m4[0] += x[i];
    m4[1] += x[i + 1];
    m4[2] += x[i + 2];
    m4[3] += x[i + 3];
  }
  const float mean = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / D_HEAD);
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = x[i + k] - mean;
      v4[k] = fmaf(d, d, v4[k]);
    }
  }
  const float var = (v4[0] + v4[1]) + (v4[2] + v4[3]);
  const float inv = rsqrtf(var * (1.0f / D_HEAD) + 1e-5f);   // nn.LayerNorm(48), submodules.py:200-201
#pragma unroll
  for (int i = 0; i < D_HEAD; ++i) x[i] = (x[i] - mean) * inv * w[i] + bia[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) {                              // RoPE on dims [0,32): pairs (i, i+16)
    const float c = cs[i], s = sn[i];
    const float x1 = x[i], x2 = x[i + 16];
    x[i] = x1 * c - x2 * s;
    x[i + 16] = x2 * c + x1 * s;
  }
#pragma unroll
  for (int i = 0; i < D_HEAD; ++i) dst[i] = x[i] * scale;
}

// Score-sharing layers: one CTA per STREAM, all 8 heads.  Recompute layers: two CTAs per stream, 4 heads each.
//   RECOMPUTE (layers 0, 7, 14, 15; 256 threads): thread = (key row j, head) or (query row t, head): per-head LayerNorm
//   + RoPE with the row in registers; query rows go to shared memory, key rows stay in registers and produce their
//   column of the score matrix; one warp per (head, query) row does the masked softmax and publishes P for the
//   score-sharing layers.
//   all layers (384 threads in the score-sharing instantiation): thread = (head, dim) accumulates ctx[t] over the keys,
//   reading V coalesced straight from global memory in batches of 8 keys (each V element is used by exactly one
//   thread), P broadcast from shared memory.
constexpr int ATT_THREADS_REC = 256;
constexpr int ATT_HEADS_REC = 4;            // heads per CTA in the recompute instantiation (grid.y = 2)
constexpr int ATT_THREADS = D_MODEL;       // 384
constexpr int ATT_TK = MHSA_S + MAX_T;     // 43
constexpr int ATT_V_SMEM = ATT_TK * ATT_HEADS_REC * D_HEAD * 4;   // 33,024 B: V tile of the recompute instantiation

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool RECOMPUTE>
__global__ void __launch_bounds__(RECOMPUTE ? ATT_THREADS_REC : ATT_THREADS) attention_kernel(const AttnArgs a) {
  constexpr int NT = RECOMPUTE ? ATT_THREADS_REC : ATT_THREADS;
  constexpr int NH = RECOMPUTE ? ATT_HEADS_REC : N_HEADS;                       // heads handled by this CTA
  constexpr int NC = NH * D_HEAD;                                               // channels handled by this CTA
  // query rows as [t][d / 4][head] float4: the heads of a quarter-warp read consecutive 16 B words (no conflicts)
  __shared__ float4 qs[RECOMPUTE ? MAX_T : 1][D_HEAD / 4][NH];                  // 10 KB (recompute only)
  __shared__ __align__(16) float ps[NH][MAX_T][48];                             // columns >= Tk are zero in the P.V loop
  // recompute instantiation: the CTA's V tile [Tk][NC] fp32 (dynamic shared memory, ATT_V_SMEM bytes), fetched with
  // cp.async right after the dependency wait so that it travels under the LayerNorm / score / softmax phases
  extern __shared__ __align__(16) float vsm[];
  PROF_DECL();
  PROF_BEGIN(4);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int T = a.T, Tk = a.Tk, S = a.S;
  const int h0 = RECOMPUTE ? blockIdx.y * NH : 0;        // first head of this CTA
  float* Pg = a.P + ((size_t)b * N_HEADS + h0) * T * Tk;
  const float* vcol = a.v + (size_t)b * Tk * a.ldv + h0 * D_HEAD + tid;
  // padding columns [Tk, Tk + 8) are read by the 8-wide P.V loop: keep them zero (all other phases touch j < Tk only)
  for (int i = tid; i < NH * MAX_T * 8; i += NT) {
    const int j = Tk + (i & 7);
    if (j < 48) (&ps[0][0][0])[(i >> 3) * 48 + j] = 0.f;
  }
  if constexpr (RECOMPUTE) {
    const float* vrow0 = a.v + (size_t)b * Tk * a.ldv + h0 * D_HEAD;
    for (int i = tid; i < Tk * (NC / 4); i += NT) {
      const int j = i / (NC / 4), c4 = i - j * (NC / 4);
      cp_async16(vsm + j * NC + c4 * 4, vrow0 + (size_t)j * a.ldv + c4 * 4);
    }
    cp_async_commit();
  }

  if constexpr (RECOMPUTE) {
    const int nk = Tk * NH, nq = T * NH;                // <= 172 + 52 <= 256
    float kx[D_HEAD];
    int kj = -1, kh = 0;
    if (tid < nk) {
      kj = tid / NH;
      kh = tid - kj * NH;
      ln_rope_row(a.k + (size_t)(b * Tk + kj) * a.ldk + (h0 + kh) * D_HEAD, a.k_ln_w, a.k_ln_b,
                  a.rope_cos + (kj - S + MHSA_S) * 16, a.rope_sin + (kj - S + MHSA_S) * 16, 1.0f, kx);
    } else if (tid < nk + nq) {
      const int i = tid - nk, t = i / NH, h = i - t * NH;
      float qx[D_HEAD];
      ln_rope_row(a.q + (size_t)(b * T + t) * a.ldq + (h0 + h) * D_HEAD, a.q_ln_w, a.q_ln_b,
                  a.rope_cos + (t + MHSA_S) * 16, a.rope_sin + (t + MHSA_S) * 16,
                  0.14433756729740643f /* 1/sqrt(48) */, qx);
#pragma unroll
      for (int d = 0; d < D_HEAD; d += 4) qs[t][d >> 2][h] = make_float4(qx[d], qx[d + 1], qx[d + 2], qx[d + 3]);
    }
    __syncthreads();
    if (threadIdx.x == 0) PROF_MARK(1);
    int off = 0;
    if (a.mask_mode == 1) off = MHSA_S - a.len_in[b];
    else if (a.mask_mode == 2) off = (MHSA_S - a.len_in[b]) / 2;
    if (kj >= 0) {
      const bool masked = kj < off;                              // cache columns older than the stream
      for (int t = 0; t < T; ++t) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int d = 0; d < D_HEAD; d += 4) {
          const float4 qv = qs[t][d >> 2][kh];
          s0 = fmaf(qv.x, kx[d], s0);
          s1 = fmaf(qv.y, kx[d + 1], s1);
          s2 = fmaf(qv.z, kx[d + 2], s2);
          s3 = fmaf(qv.w, kx[d + 3], s3);
        }
        ps[kh][t][kj] = masked ? -10000.0f : (s0 + s1) + (s2 + s3);   // submodules.py:261
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) PROF_MARK(3);
    // masked softmax: 8 lanes per (head, query) row, NT / 8 rows in flight
    {
      const int g = tid >> 3, l8 = tid & 7;
      for (int r = g; r < nq; r += NT / 8) {
        const int h = r / T, t = r - h * T;
        float* pr = &ps[h][t][0];
        float e[6];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int j = l8 + 8 * k;
          e[k] = j < Tk ? pr[j] : -INFINITY;
          mx = fmaxf(mx, e[k]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          e[k] = (l8 + 8 * k < Tk) ? expf(e[k] - mx) : 0.f;
          sum += e[k];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        float* pg = Pg + (size_t)(h * T + t) * Tk;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int j = l8 + 8 * k;
          if (j < Tk) {
            const float p = (j < off) ? 0.f : e[k] * inv;          // submodules.py:262
            pr[j] = p;
            pg[j] = p;
          }
        }
      }
    }
  } else {
    for (int i = tid; i < NH * T * Tk; i += NT) {
      const int r = i / Tk, j = i - r * Tk, h = r / T, t = r - h * T;
      ps[h][t][j] = Pg[i];
    }
  }
  if constexpr (RECOMPUTE) {
    cp_async_wait_all();
    __syncthreads();
    if (threadIdx.x == 0) PROF_MARK(4);
    if (tid < NC) {
      const int h = tid / D_HEAD;
      float acc[MAX_T];
#pragma unroll
      for (int t = 0; t < MAX_T; ++t) acc[t] = 0.f;
      for (int j0 = 0; j0 < Tk; j0 += 4) {
        float vb4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) vb4[k] = (j0 + k < Tk) ? vsm[(j0 + k) * NC + tid] : 0.f;
#pragma unroll
        for (int t = 0; t < MAX_T; ++t) {
          if (t < T) {
            const float4 p0 = *reinterpret_cast<const float4*>(&ps[h][t][j0]);
            float s0 = fmaf(p0.x, vb4[0], acc[t]), s1 = p0.y * vb4[1];
            s0 = fmaf(p0.z, vb4[2], s0);
            s1 = fmaf(p0.w, vb4[3], s1);
            acc[t] = s0 + s1;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < MAX_T; ++t)
        if (t < T) a.ctx[(size_t)(b * T + t) * D_MODEL + h0 * D_HEAD + tid] = __float2bfloat16(acc[t]);
    }
    PROF_END();
    return;
  }
  // first batch of this thread's V column travels while the barrier is reached
  float vb[8];
  if (tid < NC) {
#pragma unroll
    for (int k = 0; k < 8; ++k) vb[k] = k < Tk ? vcol[(size_t)k * a.ldv] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) PROF_MARK(4);
  if (tid < NC) {
    const int h = tid / D_HEAD;
    float acc[MAX_T];
#pragma unroll
    for (int t = 0; t < MAX_T; ++t) acc[t] = 0.f;
    for (int j0 = 0; j0 < Tk; j0 += 8) {
      float vn[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) vn[k] = (j0 + 8 + k < Tk) ? vcol[(size_t)(j0 + 8 + k) * a.ldv] : 0.f;   // next batch in flight
      // V beyond Tk is zero, so stale P columns in [Tk, 48) do not matter (they are finite: the buffer is zeroed below)
#pragma unroll
      for (int t = 0; t < MAX_T; ++t) {
        if (t < T) {
          const float4 p0 = *reinterpret_cast<const float4*>(&ps[h][t][j0]);
          const float4 p1 = *reinterpret_cast<const float4*>(&ps[h][t][j0 + 4]);
          float s0 = fmaf(p0.x, vb[0], acc[t]), s1 = p0.y * vb[1];
          s0 = fmaf(p0.z, vb[2], s0);
          s1 = fmaf(p0.w, vb[3], s1);
          s0 = fmaf(p1.x, vb[4], s0);
          s1 = fmaf(p1.y, vb[5], s1);
          s0 = fmaf(p1.z, vb[6], s0);
          s1 = fmaf(p1.w, vb[7], s1);
          acc[t] = s0 + s1;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) vb[k] = vn[k];
    }
#pragma unroll
    for (int t = 0; t < MAX_T; ++t)
      if (t < T) a.ctx[(size_t)(b * T + t) * D_MODEL + h0 * D_HEAD + tid] = __float2bfloat16(acc[t]);
  }
  PROF_END();
}

// Large-batch form of the recompute layers (0, 7, 14, 15): persistent CTAs walk whole streams (512 threads, one CTA per
// SM for the cached-context layers 14 / 15; 256 threads, two CTAs per SM for layers 0 / 7).  A stream's q / k / v rows
// are contiguous blocks of the projection outputs, so one warp brings them into shared memory with bulk async copies,
// one per row (rows padded by 16 B: the per-(key, head) LayerNorm reads then spread over the banks).  The k + q rows
// and the v rows have their own buffers and barriers: k / q of stream i + 1 are requested as soon as the LayerNorm phase
// of stream i has put them in registers, v of stream i + 1 when the P.V phase of stream i is over - both travel under
// the arithmetic of the phases that do not need them.  LayerNorm gains and the RoPE table are staged in shared memory
// once per CTA.  The one-CTA-per-(stream, head half) kernel above is a chain of dependent global round trips at two
// CTAs per SM: 7 waves of ~10 us at 1024 streams.
constexpr int ATP_PAD = 16;
constexpr int ATP_QS = MAX_T * (D_HEAD / 4) * N_HEADS * 16;          // 19,968: query rows [t][d / 4][head] float4
constexpr int ATP_PS = N_HEADS * MAX_T * 48 * 4;                     // 19,968: scores / probabilities
constexpr int ATP_PAR = (4 * D_HEAD + 2 * (MHSA_S + MAX_T) * 16) * 4;   // q/k LayerNorm gain + bias, RoPE cos | sin: 6,272
constexpr int ATP_FIXED_SMEM = ATP_QS + ATP_PS + ATP_PAR;

__host__ __device__ inline int atp_kq_bytes(int T, int Tk) { return (Tk + T) * (D_MODEL * 4 + ATP_PAD) + (Tk == T ? T * D_MODEL * 4 : 0); }
__host__ __device__ inline int atp_v_bytes(int Tk) { return Tk * (D_MODEL * 4 + ATP_PAD); }

// LayerNorm(48) + RoPE of one (row, head) slice held in shared memory; parameters in shared memory as float4
__device__ __forceinline__ void ln_rope_row_s(const float* src, const float4* w4, const float4* b4, const float4* cs4,
                                              const float4* sn4, float scale, float* x) {
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(src + i);
    x[i] = t.x; x[i + 1] = t.y; x[i + 2] = t.z; x[i + 3] = t.w;
  }
  float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
    m4[0] += x[i];
    m4[1] += x[i + 1];
    m4[2] += x[i + 2];
    m4[3] += x[i + 3];
  }
  const float mean = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / D_HEAD);
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = x[i + k] - mean;
      v4[k] = fmaf(d, d, v4[k]);
    }
  }
  const float var = (v4[0] + v4[1]) + (v4[2] + v4[3]);
  const float inv = rsqrtf(var * (1.0f / D_HEAD) + 1e-5f);   // nn.LayerNorm(48), submodules.py:200-201
#pragma unroll
  for (int i = 0; i < D_HEAD; i += 4) {
    const float4 w = w4[i >> 2], bb = b4[i >> 2];
    x[i] = (x[i] - mean) * inv * w.x + bb.x;
    x[i + 1] = (x[i + 1] - mean) * inv * w.y + bb.y;
    x[i + 2] = (x[i + 2] - mean) * inv * w.z + bb.z;
    x[i + 3] = (x[i + 3] - mean) * inv * w.w + bb.w;
  }
#pragma unroll
  for (int i = 0; i < 16; i += 4) {                           // RoPE on dims [0,32): pairs (i, i+16)
    const float4 c = cs4[i >> 2], sn = sn4[i >> 2];
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x1 = x[i + k], x2 = x[i + k + 16];
      x[i + k] = x1 * cc[k] - x2 * ss[k];
      x[i + k + 16] = x2 * cc[k] + x1 * ss[k];
    }
  }
#pragma unroll
  for (int i = 0; i < D_HEAD; ++i) x[i] *= scale;
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 1 : 2) attention_pipe_kernel(const AttnArgs a, int B) {
  extern __shared__ __align__(128) unsigned char asm_[];
  __shared__ uint64_t kq_full, v_full;
  PROF_DECL();
  PROF_BEGIN(4);
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.T, Tk = a.Tk, S = a.S;
  const bool cached = S > 0;
  constexpr int RS = D_MODEL * 4 + ATP_PAD;                 // padded row of 384 floats
  float4* qs = reinterpret_cast<float4*>(asm_);                          // [MAX_T][12][8]
  float* ps = reinterpret_cast<float*>(asm_ + ATP_QS);                   // [8][MAX_T][48]
  float* par = reinterpret_cast<float*>(asm_ + ATP_QS + ATP_PS);         // q_w | q_b | k_w | k_b | cos | sin
  const float4* qw4 = reinterpret_cast<const float4*>(par);
  const float4* qb4 = qw4 + D_HEAD / 4;
  const float4* kw4 = qb4 + D_HEAD / 4;
  const float4* kb4 = kw4 + D_HEAD / 4;
  const float4* cos4 = kb4 + D_HEAD / 4;                                  // [MHSA_S + MAX_T][4]
  const float4* sin4 = cos4 + (MHSA_S + MAX_T) * 4;
  unsigned char* kbuf = asm_ + ATP_FIXED_SMEM;                            // k rows [Tk][RS]
  unsigned char* qbuf = kbuf + Tk * RS;                                   // q rows [T][RS]
  unsigned char* vbuf = qbuf + T * RS;                                    // v rows [Tk][RS]
  if (tid == 0) {
    mbar_init(&kq_full, 1);
    mbar_init(&v_full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < N_HEADS * MAX_T * 48; i += NT) ps[i] = 0.f;   // columns >= Tk stay zero for the P.V loop
  for (int i = tid; i < D_HEAD; i += NT) {
    par[i] = __ldg(a.q_ln_w + i);
    par[D_HEAD + i] = __ldg(a.q_ln_b + i);
    par[2 * D_HEAD + i] = __ldg(a.k_ln_w + i);
    par[3 * D_HEAD + i] = __ldg(a.k_ln_b + i);
  }
  for (int i = tid; i < (MHSA_S + MAX_T) * 16; i += NT) {
    par[4 * D_HEAD + i] = __ldg(a.rope_cos + i);
    par[4 * D_HEAD + (MHSA_S + MAX_T) * 16 + i] = __ldg(a.rope_sin + i);
  }
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  const int nmine = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // warp 0: bulk copies of stream number k of this CTA.  Uncached layers: rows of [q | k | v] (ldq = 1152 floats);
  // cached layers: k | v rows from the K/V projection (ldk = 768 floats), q rows from the Q projection.
  auto issue_kq = [&](int k) {
    const int b = blockIdx.x + k * gridDim.x;
    if (lane == 0) mbar_expect_tx(&kq_full, (uint32_t)((Tk + T) * D_MODEL * 4));
    __syncwarp();
    for (int j = lane; j < Tk; j += 32) bulk_load_1d(kbuf + j * RS, a.k + (size_t)(b * Tk + j) * a.ldk, D_MODEL * 4, &kq_full);
    for (int t = lane; t < T; t += 32) bulk_load_1d(qbuf + t * RS, a.q + (size_t)(b * T + t) * a.ldq, D_MODEL * 4, &kq_full);
  };
  auto issue_v = [&](int k) {
    const int b = blockIdx.x + k * gridDim.x;
    if (lane == 0) mbar_expect_tx(&v_full, (uint32_t)(Tk * D_MODEL * 4));
    __syncwarp();
    for (int j = lane; j < Tk; j += 32) bulk_load_1d(vbuf + j * RS, a.v + (size_t)(b * Tk + j) * a.ldv, D_MODEL * 4, &v_full);
  };
  if (warp == 0 && nmine > 0) {
    issue_kq(0);
    issue_v(0);
  }
  const int nk = Tk * N_HEADS, nq = T * N_HEADS;
  for (int k = 0; k < nmine; ++k) {
    const int b = blockIdx.x + k * gridDim.x;
    mbar_wait(&kq_full, k & 1);
    // ---- per-head LayerNorm + RoPE: key rows stay in registers, query rows go to shared memory (items beyond the
    // thread count of the 256-thread form take a second pass)
    float kx[D_HEAD];
    int kj = -1, kh = 0;
    if (tid < nk) {
      kh = tid / Tk;
      kj = tid - kh * Tk;
      ln_rope_row_s(reinterpret_cast<const float*>(kbuf + kj * RS) + kh * D_HEAD, kw4, kb4, cos4 + (kj - S + MHSA_S) * 4,
                    sin4 + (kj - S + MHSA_S) * 4, 1.0f, kx);
    }
    for (int i = tid - nk; i < nq; i += NT) {
      if (i < 0) continue;
      const int h = i / T, t = i - h * T;
      float qx[D_HEAD];
      ln_rope_row_s(reinterpret_cast<const float*>(qbuf + t * RS) + h * D_HEAD, qw4, qb4, cos4 + (t + MHSA_S) * 4,
                    sin4 + (t + MHSA_S) * 4, 0.14433756729740643f /* 1/sqrt(48) */, qx);
#pragma unroll
      for (int d = 0; d < D_HEAD; d += 4) qs[(t * (D_HEAD / 4) + (d >> 2)) * N_HEADS + h] = make_float4(qx[d], qx[d + 1], qx[d + 2], qx[d + 3]);
    }
    __syncthreads();                                         // k / q rows are in registers / qs: their buffers are free
    if (warp == 0 && k + 1 < nmine) issue_kq(k + 1);
    int off = 0;
    if (a.mask_mode == 1) off = MHSA_S - a.len_in[b];
    else if (a.mask_mode == 2) off = (MHSA_S - a.len_in[b]) / 2;
    if (kj >= 0) {
      const bool masked = kj < off;                              // cache columns older than the stream
      for (int t = 0; t < T; ++t) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int d = 0; d < D_HEAD; d += 4) {
          const float4 qv = qs[(t * (D_HEAD / 4) + (d >> 2)) * N_HEADS + kh];
          s0 = fmaf(qv.x, kx[d], s0);
          s1 = fmaf(qv.y, kx[d + 1], s1);
          s2 = fmaf(qv.z, kx[d + 2], s2);
          s3 = fmaf(qv.w, kx[d + 3], s3);
        }
        ps[(kh * MAX_T + t) * 48 + kj] = masked ? -10000.0f : (s0 + s1) + (s2 + s3);   // submodules.py:261
      }
    }
    __syncthreads();
    // ---- masked softmax: 8 lanes per (head, query) row; P is published for the score-sharing layers
    {
      float* Pg = a.P + (size_t)b * N_HEADS * T * Tk;
      const int g = tid >> 3, l8 = tid & 7;
      for (int r = g; r < nq; r += NT / 8) {
        const int h = r / T, t = r - h * T;
        float* pr = ps + (h * MAX_T + t) * 48;
        float e[6];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int j = l8 + 8 * q;
          e[q] = j < Tk ? pr[j] : -INFINITY;
          mx = fmaxf(mx, e[q]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          e[q] = (l8 + 8 * q < Tk) ? expf(e[q] - mx) : 0.f;
          sum += e[q];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        float* pg = Pg + (size_t)(h * T + t) * Tk;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int j = l8 + 8 * q;
          if (j < Tk) {
            const float p = (j < off) ? 0.f : e[q] * inv;          // submodules.py:262
            pr[j] = p;
            pg[j] = p;
          }
        }
      }
    }
    __syncthreads();
    // ---- ctx = P v: thread = (head, dim), V read from its buffer
    mbar_wait(&v_full, k & 1);
    for (int c = tid; c < D_MODEL; c += NT) {
      const int h = c / D_HEAD;
      const float* vcol = reinterpret_cast<const float*>(vbuf) + c;
      float acc[MAX_T];
#pragma unroll
      for (int t = 0; t < MAX_T; ++t) acc[t] = 0.f;
      for (int j0 = 0; j0 < Tk; j0 += 4) {
        float vb4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) vb4[q] = (j0 + q < Tk) ? vcol[(j0 + q) * (RS / 4)] : 0.f;
#pragma unroll
        for (int t = 0; t < MAX_T; ++t) {
          if (t < T) {
            const float4 p0 = *reinterpret_cast<const float4*>(ps + (h * MAX_T + t) * 48 + j0);
            float s0 = fmaf(p0.x, vb4[0], acc[t]), s1 = p0.y * vb4[1];
            s0 = fmaf(p0.z, vb4[2], s0);
            s1 = fmaf(p0.w, vb4[3], s1);
            acc[t] = s0 + s1;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < MAX_T; ++t)
        if (t < T) a.ctx[(size_t)(b * T + t) * D_MODEL + c] = __float2bfloat16(acc[t]);
    }
    __syncthreads();                                         // the v buffer and ps are free again
    if (warp == 0 && k + 1 < nmine) issue_v(k + 1);
  }
  PROF_END();
}

// ------------------------------------------------------------------------------------------------ depthwise conv
struct DwArgs {
  const bf16* g;        // [B*T][384] GLU output of this layer
  bf16* cache;          // [slots][16][30][384] (layer offset already applied), time-major ring: logical row i (0 = oldest)
                        // sits at physical row (pos + i) mod 30; a step overwrites its T oldest rows and advances pos by T
  long long cache_slot_stride;
  const int* slots;
  const int* cpos_in;   // [B][2] (rate offset applied): physical row of the oldest cached frame (the cache is a ring)
  const float* w;       // [31][384] BN-folded taps
  const float* bias;    // [384]     BN-folded bias
  bf16* e;              // [B*T][384]
  int T;
};

// Two CTAs per stream (192 channels each), 384 threads = (channel, half of the output frames).  The
// [30 cached | T new] x 192 tile is staged in shared memory with 16-byte coalesced loads (all in flight at once); each
// thread then owns one channel and ceil(T/2) output frames: taps in registers, inputs streamed from smem.  The new
// cache is rows [T, T+30) of the tile, copied back with coalesced 16-byte stores.
constexpr int DW_CH = 192;                       // channels per CTA
constexpr int DW_THREADS = 2 * DW_CH;
constexpr int DW_ROWS = CONV_S + MAX_T;          // 43
constexpr int DW_RV = DW_CH / 8;                 // uint4 per tile row
constexpr int DW_TH = (MAX_T + 1) / 2;           // output frames per thread (7)

// TH = output frames per thread = ceil(T / 2), a template parameter so that the unrolled tap loop does no work for frames
// that do not exist (T = 10 -> 5, T = 13 -> 7, reduced-rate layers T = 5 / 6 -> 3): the kernel is bound by FMA issue.
template <int TH>
__global__ void __launch_bounds__(DW_THREADS) dwconv_kernel(const DwArgs a) {
  __shared__ __align__(16) bf16 tile[DW_ROWS][DW_CH];
  PROF_DECL();
  PROF_BEGIN(5);
  pdl_launch_dependents();
  const int b = blockIdx.x, c0 = blockIdx.y * DW_CH, tid = threadIdx.x;
  const int T = a.T;
  const int cl = tid % DW_CH, th = tid / DW_CH;  // local channel, frame half
  const int c = c0 + cl;
  // BN-folded taps of this thread's channel: weights, fetched before the PDL wait
  float w[CONV_S + 1];
#pragma unroll
  for (int j = 0; j <= CONV_S; ++j) w[j] = __ldg(a.w + j * D_MODEL + c);
  const float bb = __ldg(a.bias + c);
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  bf16* cache = a.cache + (size_t)a.slots[b] * a.cache_slot_stride + c0;
  const int pos = a.cpos_in[2 * b];
  {
    const bf16* g = a.g + (size_t)b * T * D_MODEL + c0;
    uint4* t4 = reinterpret_cast<uint4*>(&tile[0][0]);
    const int n_cache = CONV_S * DW_RV, n_all = (CONV_S + T) * DW_RV;       // 720, <= 1032
    uint4 tmp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * DW_THREADS;
      if (i < n_all) {
        const int r = i / DW_RV, v = i - r * DW_RV;
        const int pr = r + pos >= CONV_S ? r + pos - CONV_S : r + pos;        // ring: logical row r -> physical row
        tmp[k] = (i < n_cache) ? *reinterpret_cast<const uint4*>(cache + (size_t)pr * D_MODEL + v * 8)
                               : *reinterpret_cast<const uint4*>(g + (size_t)(r - CONV_S) * D_MODEL + v * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * DW_THREADS;
      if (i < n_all) t4[i] = tmp[k];
    }
  }
  __syncthreads();
  // y[t] = b' + sum_j w'[j] x[t + j]  (submodules.py:364-402 with BatchNorm folded), input-stationary order;
  // this thread's frames: t0 .. t0 + DW_TH - 1 (clipped to T)
  const int half = (T + 1) / 2;
  const int t0 = th * half;
  float acc[TH];
#pragma unroll
  for (int t = 0; t < TH; ++t) acc[t] = bb;
#pragma unroll
  for (int i = 0; i < CONV_S + TH; ++i) {             // rows t0 + i of the tile
    if (t0 + i < CONV_S + T) {
      const float x = __bfloat162float(tile[t0 + i][cl]);
#pragma unroll
      for (int t = 0; t < TH; ++t) {
        const int j = i - t;                               // compile-time after unrolling
        if (j >= 0 && j <= CONV_S) acc[t] = fmaf(w[j], x, acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TH; ++t)
    if (t < half && t0 + t < T) a.e[(size_t)(b * T + t0 + t) * D_MODEL + c] = __float2bfloat16(silu_f(acc[t]));
  // new cache = last 30 rows of [cache | g]: the T new rows replace the T oldest rows of the ring
  {
    const uint4* t4 = reinterpret_cast<const uint4*>(&tile[CONV_S][0]);
    const int i = tid;
    if (i < T * DW_RV) {
      const int r = i / DW_RV, v = i - r * DW_RV;
      const int pr = r + pos >= CONV_S ? r + pos - CONV_S : r + pos;
      *reinterpret_cast<uint4*>(cache + (size_t)pr * D_MODEL + v * 8) = t4[i];
    }
  }
  PROF_END();
}

// Large-batch form: persistent CTAs (two per SM) walk whole streams.  A stream's [30 cached | T new] x 384 tile is two
// contiguous blocks in global memory (the layer's cache block of the slot, the stream's rows of g), so one thread brings
// it in with two bulk async copies into a 3-stage ring (the one-CTA-per-(stream, channel half) kernel above has too few
// bytes in flight per SM to reach the memory system's rate once the batch is in the thousands of rows); thread =
// channel, all T output frames, taps in registers; the new cache (rows [T, T + 30) of the tile) leaves as one bulk
// store straight from the tile.
constexpr int DWP_THREADS = D_MODEL;
constexpr int DWP_STAGES = 3;
constexpr int DWP_STAGE_BYTES = (CONV_S + MAX_T) * D_MODEL * 2;      // 33,024
constexpr int DWP_SMEM = DWP_STAGES * DWP_STAGE_BYTES;

template <int TM>     // TM >= T: compile-time bound of the output frames per stream
__global__ void __launch_bounds__(DWP_THREADS, 2) dwconv_pipe_kernel(const DwArgs a, int B) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ uint64_t full[DWP_STAGES];
  PROF_DECL();
  PROF_BEGIN(5);
  pdl_launch_dependents();
  const int tid = threadIdx.x, T = a.T;
  if (tid == 0) {
    for (int s = 0; s < DWP_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  float w[CONV_S + 1];
#pragma unroll
  for (int j = 0; j <= CONV_S; ++j) w[j] = __ldg(a.w + j * D_MODEL + tid);
  const float bb = __ldg(a.bias + tid);
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  constexpr uint32_t CACHE_BYTES = CONV_S * D_MODEL * 2;
  const uint32_t g_bytes = (uint32_t)T * D_MODEL * 2;
  const int nmine = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr uint32_t ROW_BYTES = D_MODEL * 2;
  auto issue = [&](int k) {
    const int b = blockIdx.x + k * gridDim.x, s = k % DWP_STAGES;
    unsigned char* tile = dsm + s * DWP_STAGE_BYTES;
    const bf16* cache = a.cache + (size_t)a.slots[b] * a.cache_slot_stride;
    const int pos = a.cpos_in[2 * b];            // ring: physical rows [pos, 30) are the oldest, then [0, pos)
    mbar_expect_tx(&full[s], CACHE_BYTES + g_bytes);
    bulk_load_1d(tile, cache + (size_t)pos * D_MODEL, (CONV_S - pos) * ROW_BYTES, &full[s]);
    if (pos) bulk_load_1d(tile + (CONV_S - pos) * ROW_BYTES, cache, pos * ROW_BYTES, &full[s]);
    bulk_load_1d(tile + CACHE_BYTES, a.g + (size_t)b * T * D_MODEL, g_bytes, &full[s]);
  };
  if (tid == 0)
    for (int k = 0; k < DWP_STAGES - 1 && k < nmine; ++k) issue(k);
  for (int k = 0; k < nmine; ++k) {
    const int b = blockIdx.x + k * gridDim.x, s = k % DWP_STAGES;
    if (tid == 0 && k + DWP_STAGES - 1 < nmine) {
      bulk_wait_read0();            // the cache store of the previous stream has left the stage being refilled
      issue(k + DWP_STAGES - 1);
    }
    mbar_wait(&full[s], (k / DWP_STAGES) & 1);
    const bf16* tile = reinterpret_cast<const bf16*>(dsm + s * DWP_STAGE_BYTES) + tid;
    float acc[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) acc[t] = bb;
#pragma unroll
    for (int i = 0; i < CONV_S + TM; ++i) {             // tile rows; y[t] = b' + sum_j w'[j] x[t + j]
      if (i < CONV_S + T) {
        const float x = __bfloat162float(tile[i * D_MODEL]);
#pragma unroll
        for (int t = 0; t < TM; ++t) {
          const int j = i - t;                               // compile-time after unrolling
          if (j >= 0 && j <= CONV_S) acc[t] = fmaf(w[j], x, acc[t]);
        }
      }
    }
    bf16* eo = a.e + (size_t)b * T * D_MODEL + tid;
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < T) eo[t * D_MODEL] = __float2bfloat16(silu_f(acc[t]));
    __syncthreads();                                      // every thread is done reading the tile
    if (tid == 0) {                                       // the T new rows replace the T oldest rows of the ring
      bf16* cache = a.cache + (size_t)a.slots[b] * a.cache_slot_stride;
      const unsigned char* gsm = dsm + s * DWP_STAGE_BYTES + CACHE_BYTES;
      const int pos = a.cpos_in[2 * b];
      const int n0 = min(T, CONV_S - pos);
      bulk_store_1d(cache + (size_t)pos * D_MODEL, gsm, n0 * ROW_BYTES);
      if (n0 < T) bulk_store_1d(cache, gsm + n0 * ROW_BYTES, (T - n0) * ROW_BYTES);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();
  PROF_END();
}

// ------------------------------------------------------------------------------------------------ temporal reduction
// m[4c+k, t2] = b[4c+k] + sum_j W[4c+k][j] * v[c][2 t2 + j],  v = [red | r^T]  (conformer_blocks.py:874-911).
struct RedArgs {
  const float* r;       // [B*T][384] layer-6 output (norm_out applied)
  float* red;           // [slots][384]
  const int* slots;
  const float* w;       // [1536][3]
  const float* bias;    // [1536]
  bf16* m;              // [B*T2][1536]
  int T, T2;
};

__global__ void __launch_bounds__(D_MODEL) reduction_dw_kernel(const RedArgs a) {
  PROF_DECL();
  PROF_BEGIN(6);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) PROF_MARK(2);
  const int b = blockIdx.x, c = threadIdx.x;
  float* red = a.red + (size_t)a.slots[b] * D_MODEL;
  float w[12], bia[4];
#pragma unroll
  for (int i = 0; i < 12; ++i) w[i] = a.w[c * 12 + i];
#pragma unroll
  for (int i = 0; i < 4; ++i) bia[i] = a.bias[c * 4 + i];
  float prev = red[c];
  const float* r = a.r + (size_t)b * a.T * D_MODEL + c;
  for (int t2 = 0; t2 < a.T2; ++t2) {
    const float v0 = prev;                                    // v[c][2 t2]
    const float v1 = r[(size_t)(2 * t2) * D_MODEL];           // v[c][2 t2 + 1] = r[2 t2]
    const float v2 = r[(size_t)(2 * t2 + 1) * D_MODEL];       // v[c][2 t2 + 2] = r[2 t2 + 1]
    prev = v2;
    uint32_t p[2];
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = bia[k] + w[k * 3] * v0 + w[k * 3 + 1] * v1 + w[k * 3 + 2] * v2;
    p[0] = pack_bf16x2(o[0], o[1]);
    p[1] = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(a.m + ((size_t)b * a.T2 + t2) * (4 * D_MODEL) + c * 4) = make_uint2(p[0], p[1]);
  }
  red[c] = r[(size_t)(a.T - 1) * D_MODEL];                    // new state = last column of v
  PROF_END();
}

}  // namespace tone
