// Per-stream state: reset, and gather / scatter between the engine's HBM layout (engine.cu header comment) and the
// reference's flat fp16 wire format (tone/nn/model.py:259-267: preproc 80 | mhsa (2,30,384) | conv (16,384,30) | len 1 |
// sub1 (1,10,64) | sub2 (32,8,44) | reduction (384,1); 219,729 elements, tone/onnx_wrapper.py:34).
// Bandwidth kernels off the hot path (slot alloc, checkpoint / migration, the numpy-state mode of the model class).
#pragma once

#include "ctc_phrase.cuh"
#include "kernels.cuh"

namespace tone {

struct StatePool {
  __half* pre;      // [slots][80]
  bf16* feat;       // [slots][FEAT_ROWS_MAX][64]
  bf16* x1;         // [slots][X1_ROWS_MAX][X1_ROW]
  bf16* kv14;       // [slots][KV_ROWS_MAX][384]
  bf16* kv15;
  bf16* conv;       // [slots][16][30][384]
  float* red;       // [slots][384]
  int* len;         // [slots]
  int* cpos;        // [slots][2] ring positions of the conv caches (full-rate layers 0..6, 15 | reduced-rate layers 7..14)
  PhSlot* ph;       // [slots]
  int F, T, T2;
};

namespace st_off {
constexpr int pre = 0, mhsa = 80, conv = mhsa + 2 * 30 * 384, len = conv + 16 * 384 * 30, sub1 = len + 1,
              sub2 = sub1 + 640, red = sub2 + 32 * 8 * 44, end = red + 384;
static_assert(end == 219729, "state layout");
}  // namespace st_off

constexpr int STATE_IO_THREADS = 256;
constexpr int STATE_IO_CHUNK = 32;        // slots per gather / scatter launch (28 MB of fp16 staging)

__device__ __forceinline__ void zero_bytes(void* p, size_t bytes, int tid, int nt) {   // p 16-byte aligned, bytes % 16 == 0
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = tid; i < bytes / 16; i += nt) q[i] = z;
}

// grid (n): the all-zero initial state (tone/nn/model.py:208-267) and an empty splitter state
__global__ void __launch_bounds__(STATE_IO_THREADS) reset_slots_kernel(StatePool p, const int* slots) {
  const size_t s = slots[blockIdx.x];
  const int tid = threadIdx.x, nt = blockDim.x;
  zero_bytes(p.pre + s * HOP, HOP * 2, tid, nt);
  zero_bytes(p.feat + s * FEAT_ROWS_MAX * N_MELS, FEAT_ROWS_MAX * N_MELS * 2, tid, nt);
  zero_bytes(p.x1 + s * X1_ROWS_MAX * X1_ROW, (size_t)X1_ROWS_MAX * X1_ROW * 2, tid, nt);
  zero_bytes(p.kv14 + s * KV_ROWS_MAX * D_MODEL, KV_ROWS_MAX * D_MODEL * 2, tid, nt);
  zero_bytes(p.kv15 + s * KV_ROWS_MAX * D_MODEL, KV_ROWS_MAX * D_MODEL * 2, tid, nt);
  zero_bytes(p.conv + s * 16 * CONV_S * D_MODEL, (size_t)16 * CONV_S * D_MODEL * 2, tid, nt);
  zero_bytes(p.red + s * D_MODEL, D_MODEL * 4, tid, nt);
  if (tid == 0) {
    p.len[s] = 0;
    p.cpos[2 * s] = 0;
    p.cpos[2 * s + 1] = 0;
    PhSlot z;
    z.head = 0;
    z.n = 0;
    z.s = -1;
    z.run = 0;
    z.offset = 0;
    z.pad[0] = z.pad[1] = z.pad[2] = 0;
    p.ph[s] = z;
  }
}

__device__ __forceinline__ __half bf2h(bf16 v) { return __float2half_rn(__bfloat162float(v)); }
__device__ __forceinline__ bf16 h2bf(__half v) { return __float2bfloat16_rn(__half2float(v)); }

// grid (n, parts): flat[n][219729] <- slot state.  Between steps the carried rows sit at the END of [cache | new]:
// feat rows [F, F+10), x1 rows [F, F+8), kv14 rows [T2, T2+15), kv15 rows [T, T+30).
__global__ void __launch_bounds__(STATE_IO_THREADS) export_state_kernel(StatePool p, const int* slots, __half* out) {
  const size_t s = slots[blockIdx.x];
  __half* o = out + (size_t)blockIdx.x * st_off::end;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < st_off::end; i += gridDim.y * blockDim.x) {
    __half v;
    if (i < st_off::mhsa) {
      v = p.pre[s * HOP + i];
    } else if (i < st_off::conv) {
      const int j = i - st_off::mhsa, l = j / (30 * D_MODEL), r = (j / D_MODEL) % 30, c = j % D_MODEL;
      if (l == 0) v = r < 15 ? __float2half_rn(0.f) : bf2h(p.kv14[(s * KV_ROWS_MAX + p.T2 + r - 15) * D_MODEL + c]);
      else v = bf2h(p.kv15[(s * KV_ROWS_MAX + p.T + r) * D_MODEL + c]);
    } else if (i < st_off::len) {
      const int j = i - st_off::conv, l = j / (D_MODEL * CONV_S), c = (j / CONV_S) % D_MODEL, t = j % CONV_S;
      // the cache is a ring: logical row t (0 = oldest) sits at physical row (pos + t) mod 30
      const int pos = p.cpos[2 * s + ((l > 6 && l <= 14) ? 1 : 0)];
      v = bf2h(p.conv[((s * 16 + l) * CONV_S + (pos + t) % CONV_S) * D_MODEL + c]);
    } else if (i < st_off::sub1) {
      v = __float2half_rn((float)p.len[s]);
    } else if (i < st_off::sub2) {
      v = bf2h(p.feat[(s * FEAT_ROWS_MAX + p.F) * N_MELS + (i - st_off::sub1)]);
    } else if (i < st_off::red) {
      const int j = i - st_off::sub2, c = j / (SUB2_ROWS * 44), r = (j / 44) % SUB2_ROWS, f = j % 44;
      v = bf2h(p.x1[(s * X1_ROWS_MAX + p.F + r) * X1_ROW + f * 32 + c]);
    } else {
      v = __float2half_rn(p.red[s * D_MODEL + (i - st_off::red)]);
    }
    o[i] = v;
  }
}

// grid (n, parts): slot state <- flat[n][219729] (values go through the engine's storage types: bf16 caches)
__global__ void __launch_bounds__(STATE_IO_THREADS) import_state_kernel(StatePool p, const int* slots, const __half* in) {
  const size_t s = slots[blockIdx.x];
  const __half* src = in + (size_t)blockIdx.x * st_off::end;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < st_off::end; i += gridDim.y * blockDim.x) {
    const __half v = src[i];
    if (i < st_off::mhsa) {
      p.pre[s * HOP + i] = v;
    } else if (i < st_off::conv) {
      const int j = i - st_off::mhsa, l = j / (30 * D_MODEL), r = (j / D_MODEL) % 30, c = j % D_MODEL;
      if (l == 0) {
        if (r >= 15) p.kv14[(s * KV_ROWS_MAX + p.T2 + r - 15) * D_MODEL + c] = h2bf(v);
      } else {
        p.kv15[(s * KV_ROWS_MAX + p.T + r) * D_MODEL + c] = h2bf(v);
      }
    } else if (i < st_off::len) {
      const int j = i - st_off::conv, l = j / (D_MODEL * CONV_S), c = (j / CONV_S) % D_MODEL, t = j % CONV_S;
      p.conv[((s * 16 + l) * CONV_S + t) * D_MODEL + c] = h2bf(v);
    } else if (i < st_off::sub1) {
      int len = (int)lrintf(__half2float(v));
      p.len[s] = max(0, min(len, MHSA_S));
      p.cpos[2 * s] = 0;                   // imported caches are stored in logical order
      p.cpos[2 * s + 1] = 0;
    } else if (i < st_off::sub2) {
      p.feat[(s * FEAT_ROWS_MAX + p.F) * N_MELS + (i - st_off::sub1)] = h2bf(v);
    } else if (i < st_off::red) {
      const int j = i - st_off::sub2, c = j / (SUB2_ROWS * 44), r = (j / 44) % SUB2_ROWS, f = j % 44;
      p.x1[(s * X1_ROWS_MAX + p.F + r) * X1_ROW + f * 32 + c] = h2bf(v);
    } else {
      p.red[s * D_MODEL + (i - st_off::red)] = __half2float(v);
    }
  }
}

}  // namespace tone
