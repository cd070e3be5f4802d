// SIMT debug GEMMs (gemm_impl = 1): the same operands, packing and epilogue semantics as gemm_tc.cuh, computed
// one output element per thread on the CUDA cores.  They exist to bisect a parity failure between "tcgen05 / TMA /
// descriptor bug" and "packing or step-logic bug"; they are not a product path and not a CPU fallback.
#pragma once

#include "gemm_tc.cuh"

namespace tone {

template <int KIND>
__device__ __forceinline__ const bf16* ref_a_row(const GemmArgs& a, int b_or_m, int j, int ntile, int& seg_len,
                                                 int seg, bool& ok) {
  // Returns a pointer to a contiguous K segment `seg` of the A row; seg_len = its length.  ok=false when done.
  if constexpr (KIND == G_CONV0) {            // 11 segments of 64 mel bins: rows j+kt of the feature buffer
    ok = seg < 11;
    seg_len = 64;
    return a.A + (long long)a.slots[b_or_m] * a.a_slot_stride + (long long)(j + seg) * 64;
  } else if constexpr (KIND == G_CONV1) {     // 11 segments of 12*32: window rows 3t+kt, cols f0..f0+11
    ok = seg < 11;
    seg_len = 384;
    return a.A + (long long)a.slots[b_or_m] * a.a_slot_stride + (long long)((3 * j + seg) * 44 + 2 * ntile) * 32;
  } else if constexpr (KIND == G_KV) {
    ok = seg < 1;
    seg_len = a.nk * 64;
    return a.A + (long long)a.slots[b_or_m] * a.a_slot_stride + (long long)j * a.lda;
  } else {
    ok = seg < 1;
    seg_len = a.nk * 64;
    return a.A + (long long)b_or_m * a.lda;
  }
}

template <int KIND>
__device__ __forceinline__ float ref_dot(const GemmArgs& a, int b_or_m, int j, int ntile, int wrow) {
  float acc = 0.f;
  const bf16* w = a.W + (long long)wrow * a.ldw;
  int koff = 0;
  for (int seg = 0;; ++seg) {
    int len;
    bool ok;
    const bf16* ar = ref_a_row<KIND>(a, b_or_m, j, ntile, len, seg, ok);
    if (!ok) break;
    for (int k = 0; k < len; ++k) acc += __bfloat162float(ar[k]) * __bfloat162float(w[koff + k]);
    koff += len;
  }
  return acc;
}

// grid.x covers rows (dense: M; gather: B*R), grid.y covers output columns.
template <int KIND, int BN>
__global__ void gemm_ref_kernel(const GemmArgs a, int n_out) {
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  const int row = blockIdx.x;
  if (col >= n_out) return;
  int b = row, j = 0;
  if constexpr (KindTraits<KIND>::gather) {
    b = row / a.R;
    j = row - b * a.R;
  }
  if constexpr (KIND == G_PARTIAL) {
    // K slice z: a.nk iterations of 64 starting at z * a.nk * 64
    const int z = blockIdx.z, klen = a.nk * 64;
    const bf16* ar = a.A + (long long)row * a.lda + (long long)z * klen;
    const bf16* w = a.W + (long long)col * a.ldw + (long long)z * klen;
    float acc = 0.f;
    for (int k = 0; k < klen; ++k) acc += __bfloat162float(ar[k]) * __bfloat162float(w[k]);
    reinterpret_cast<__half*>(a.out)[z * a.z_stride + (long long)row * a.ldo + col] =
        __float2half_rn(fminf(fmaxf(acc, -65504.f), 65504.f));
  } else if constexpr (KIND == G_STORE_F32 || KIND == G_KV) {
    float v = ref_dot<KIND>(a, b, j, 0, col) + (a.bias ? a.bias[col] : 0.f);
    reinterpret_cast<float*>(a.out)[(long long)row * a.ldo + col] = v;
  } else if constexpr (KIND == G_RESID) {
    float v = ref_dot<KIND>(a, b, j, 0, col) + a.bias[col];
    reinterpret_cast<float*>(a.out)[(long long)row * a.ldo + col] += a.scale * v;
  } else if constexpr (KIND == G_SWIGLU || KIND == G_GLU) {
    constexpr int HW = BN / 2;               // packed tiles: [HW gate|a rows][HW value|b rows]
    const int tile = col / HW, c = col % HW;
    const int r0 = tile * BN + c, r1 = r0 + HW;
    float rs = 1.f;
    if (a.ss) {   // row-scale RMSNorm folded around the GEMM (see GemmArgs)
      float t = 0.f;
      for (int k = 0; k < a.ss_tiles; ++k) t += a.ss[(long long)row * a.ss_ld + k];
      rs = 1.0f / (sqrtf(t) * 0.05103103630798288f + 1e-8f);
    }
    float x = ref_dot<KIND>(a, b, j, 0, r0) * rs + a.bias[r0];
    float y = ref_dot<KIND>(a, b, j, 0, r1) * rs + a.bias[r1];
    float o = (KIND == G_SWIGLU) ? silu_f(x) * y : x * sigmoid_f(y);
    reinterpret_cast<bf16*>(a.out)[(long long)row * a.ldo + col] = __float2bfloat16(o);
  } else if constexpr (KIND == G_CONV0) {
    float v = ref_dot<KIND>(a, b, j, 0, col);
    const int ch = col % 32;
    v = silu_f(v * a.alpha[ch] + a.beta[ch]);
    long long off = (long long)a.slots[b] * a.out_slot_stride + (long long)(a.out_row_off + j) * a.ldo + col;
    reinterpret_cast<bf16*>(a.out)[off] = __float2bfloat16(v);
  } else if constexpr (KIND == G_CONV1) {
    const int ntile = col / 128, c = col % 128;   // column = f*64 + o ; tile = 2 f positions
    float v = ref_dot<KIND>(a, b, j, ntile, c);
    const int ch = c % 64;
    v = silu_f(v * a.alpha[ch] + a.beta[ch]);
    reinterpret_cast<bf16*>(a.out)[(long long)row * a.ldo + col] = __float2bfloat16(v);
  }
}

__global__ void decoder_ref_kernel(const GemmArgs a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.M) return;
  float lg[35];
  float mx = -INFINITY;
  int am = 0;
  for (int i = 0; i < 35; ++i) {
    lg[i] = ref_dot<G_DECODER>(a, row, 0, 0, i) + a.bias[i];
    if (lg[i] > mx) {
      mx = lg[i];
      am = i;
    }
  }
  float s = 0.f;
  for (int i = 0; i < 35; ++i) s += expf(lg[i] - mx);
  const float lse = mx + logf(s);
  float* out = reinterpret_cast<float*>(a.out) + (long long)row * 35;
  for (int i = 0; i < 35; ++i) out[i] = lg[i] - lse;
  if (a.tokens) a.tokens[row] = am;
  if (a.aux) {
    a.aux[row * 2] = lg[33] - lse;
    a.aux[row * 2 + 1] = lg[34] - lse;
  }
}

// Debug-path producer side of the row-scale RMSNorm: rb = bf16(r), ss[row][0] = sum of squares (one "tile").
__global__ void rowscale_ref_kernel(const float* r, bf16* rb, float* ss, int ss_ld, int M) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float t = 0.f;
  for (int c = 0; c < 384; ++c) {
    const float v = r[(long long)row * 384 + c];
    rb[(long long)row * 384 + c] = __float2bfloat16(v);
    t += v * v;
  }
  ss[(long long)row * ss_ld] = t;
}

template <int KIND, int BN>
inline cudaError_t launch_gemm_ref(cudaStream_t st, const GemmArgs& a, int rows, int n_out, int splits = 1) {
  if constexpr (KIND == G_DECODER) {
    decoder_ref_kernel<<<(rows + 127) / 128, 128, 0, st>>>(a);
  } else {
    gemm_ref_kernel<KIND, BN><<<dim3(rows, (n_out + 127) / 128, splits), 128, 0, st>>>(a, n_out);
  }
  return cudaGetLastError();
}

}  // namespace tone
