// tone_engine: weights, per-stream state pool, step orchestration, CUDA graphs and the C ABI (include/tone_b200.h).
//
// The step follows the reference's Tone.forward_for_export (tone/nn/model.py:101-205) / SURVEY.md Appendix A.
// Data layout in HBM (per slot = per stream, all persistent):
//   pre   [80]            fp16   last 80 samples of the previous chunk                   (A1)
//   feat  [50][64]        bf16   RMSNorm'ed log-mel rows: 10 cached + F new              (A2, conv0 input)
//   x1    [48][44][32]    bf16   conv0 output, channels-last: 8 cached + F new rows      (A2, conv1 input)
//   kv14  [44][384]       bf16   layer-14 attention input rows: 15 cached + T2 new       (A5)
//   kv15  [44][384]       bf16   layer-15 attention input rows: 30 cached + T new        (A5)
//   conv  [16][30][384]   bf16   depthwise-conv caches, time-major                       (A4.3)
//   red   [384]           fp32   last layer-6 output frame                               (A4 reduction)
//   len   int32                  mhsa_len
// Between steps the carried rows sit at the END of [cache | new] (rows [F,F+10) of feat, ...); the next step's
// begin_step_kernel rolls them to the front.  Import/export use the same convention.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tone_b200.h"
#include "gemm_ref.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "ctc_phrase.cuh"
#include "ff_fused.cuh"
#include "att_fused.cuh"
#include "rowgemm.cuh"
#include "state_io.cuh"

// NVTX ranges around the host side of every entry point that enqueues GPU work (visible in Nsight tools, filterable with
// `ncu --nvtx --nvtx-include "tone_step_graph/"`); header-only NVTX 3: a no-op costing nanoseconds when no tool is attached.
struct NvtxScope {
  explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
  ~NvtxScope() { nvtxRangePop(); }
  NvtxScope(const NvtxScope&) = delete;
  NvtxScope& operator=(const NvtxScope&) = delete;
};

using namespace tone;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) return fail(TONE_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

#define RC(x)              \
  do {                     \
    int _rc = (x);         \
    if (_rc) return _rc;   \
  } while (0)

// ------------------------------------------------------------------------------------------------ host bf16/fp16
static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
[[maybe_unused]] static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float h2f(uint16_t h) {
  __half_raw r;
  r.x = h;
  return __half2float(__half(r));
}
static inline uint16_t f2h(float f) {
  __half_raw r = static_cast<__half_raw>(__float2half_rn(f));
  return r.x;
}

// ------------------------------------------------------------------------------------------------ model constants
static const int N_LAYERS = 16, D_FF = 1536, N_CLASSES = 35, DEC_PAD = 48;
static const int SUB_OUT = 2176;  // 34 * 64
static const bool RECOMPUTE[16] = {true, false, false, false, false, false, false, true,
                                   false, false, false, false, false, false, true, true};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct WeightMat {      // a bf16 [N][K] matrix on the device with its TMA map (box 64 x BN)
  bf16* ptr = nullptr;
  int N = 0, K = 0;
  CUtensorMap map;
  CUtensorMap map128;   // same matrix, box 64 x 128: the large-batch tile shape (valid when N % 128 == 0)
  CUtensorMap map64;    // box 64 x 64: one CTA's half of a weight tile in the CTA-pair kernels (valid when N % 64 == 0)
};

struct LayerW {
  WeightMat ff1_up, ff1_down, ff2_up, ff2_down, qkv, q, kv, wo, pw1, pw1w, pw2;   // pw1w: pw1 packed for 128-wide tiles
  float *ff1_up_b, *ff1_down_b, *ff2_up_b, *ff2_down_b, *qkv_b, *q_b, *kv_b, *wo_b, *pw1_b, *pw1w_b, *pw2_b;
  float *n_ff1, *n_att, *n_conv, *n_ff2, *n_out;
  float *qln_w, *qln_b, *kln_w, *kln_b;
  float *dw_w, *dw_b;
  CUtensorMap wv48;     // score-sharing layers: Wv with a 48-row box (one head per N tile of the fused V + P.V kernel)
  // Large-batch path (rows per lane >= BIG_M): the attention input norm is applied as a row scale inside the projection
  // GEMMs (A = bf16 residual rows from the feed-forward down projection's epilogue), so the norm_self_att gain is folded
  // into the input columns of these copies (layers 0..13)
  WeightMat qkv_f;
  CUtensorMap wv48_f;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tone_engine {
  tone_config cfg;
  int C, F, T, T2;
  cudaStream_t stream = nullptr;
  PFN_encodeTiled encode = nullptr;
  std::map<std::string, HostTensor> host_w;
  bool finalized = false;

  // arenas
  char* w_arena = nullptr;
  size_t w_cap = 0, w_used = 0;
  std::vector<void*> allocs;

  // weights
  WeightMat conv0_w, conv1_w, out_w, red_pw, dec_w;
  float *conv0_alpha, *conv0_beta, *conv1_alpha, *conv1_beta, *pre_norm_g, *out_norm_g;
  float *red_dw_w, *red_dw_b, *red_pw_b, *dec_b;
  LayerW L[16];
  __half* basis;
  float *mel_w, *rope_cos, *rope_sin;
  int* mel_start;
  unsigned char* mel_bin;

  // state pool
  __half* st_pre;
  bf16 *st_feat, *st_x1, *st_kv14, *st_kv15, *st_conv;
  float* st_red;
  int* st_len;
  int* st_cpos;         // [slots][2] ring position (oldest row) of the depthwise-conv caches: full-rate / reduced-rate layers

  // phrase splitter state (ctc_phrase.cuh)
  PhSlot* st_ph = nullptr;
  unsigned char* st_ring = nullptr;
  std::vector<int> free_slots;
  std::vector<char> slot_used;
  std::vector<uint32_t> slot_stamp;     // duplicate detection: stamp of the last batch that named the slot
  uint32_t stamp_gen = 0;

  // Step inputs / outputs, per staging set.  Sets 0 .. PIPE-1 form the ring of tone_submit / tone_wait (PCM int16);
  // set PIPE serves the staged / device-pointer / feature / debug entry points (PCM int16 or int32).
  struct IoSet {
    int *d_slots = nullptr, *d_tokens = nullptr, *d_len_in = nullptr, *d_cpos_in = nullptr;
    void* d_pcm = nullptr;
    unsigned char* d_last = nullptr;
    float *d_logprobs = nullptr, *d_aux = nullptr;
    char* d_ph = nullptr;                 // PhHeader | PhRecord[4 B] | text pool
    int* p_slots = nullptr;
    int16_t* p_pcm = nullptr;
    unsigned char* p_last = nullptr;
    float *p_logprobs = nullptr, *p_aux = nullptr;
    int* p_tokens = nullptr;
    char* p_ph = nullptr;
    cudaEvent_t ev_in = nullptr, ev_step = nullptr, ev_out = nullptr;
    bool busy = false;
    int B = 0, outputs = 0, ticket = -1;
    int staged_mode = 0;                  // legacy set: SM_PCM16 if the staged PCM is int16
    bool ph_complete = true;              // the whole text pool of the ticket is on the host
  };
  static const int PIPE = 2;
  IoSet io[PIPE + 1];
  int next_ticket = 0;
  size_t ph_bytes = 0;                    // bytes of one set's phrase area
  __half* d_feats = nullptr;   // [max_batch][64][MAX_FRAMES] feature-input mode staging
  uint16_t* p_feats = nullptr; // pinned
  // small pinned ring for the slot ids of the asynchronous entry points (reset, device step, state io)
  static const int RING = 8;
  int* p_ring = nullptr;
  int* d_ring = nullptr;
  cudaEvent_t ev_ring[RING] = {};
  int ring_pos = 0;
  uint16_t* d_state_io = nullptr;         // [STATE_IO_CHUNK][219729] fp16 staging of the state wire format
  uint16_t* p_state_io = nullptr;
  int rows_alloc = 0;
  int max_splits = 8;
  CUtensorMap m_feat, m_x1, m_kv14, m_kv15;
  CUtensorMap w_feat, w_x1, w_kv14, w_kv15;   // same views with a box spanning the G slots of one tile
  // The batch is cut into up to n_lanes independent sub-batches whose kernel chains run concurrently (fork/join in
  // the captured graph): at small batch the step is bound by kernel-to-kernel latency, not by the SMs.
  struct Lane {
    float *r_full, *r_red, *qkv, *P;
    __half* part;                       // split-K partial sums of the feed-forward down projection (fp16)
    bf16 *n, *h, *ctx, *g, *ebuf, *c1, *m_red, *rb;
    float* ss;                          // [rows][12] per-N-tile sums of squares of the residual rows (row-scale RMSNorm)
    CUtensorMap m_n, m_h, m_ctx, m_e, m_c1, m_mred, m_rb;
    cudaStream_t stream = nullptr;       // lanes > 0 run on their own stream between fork and join
    cudaEvent_t done = nullptr;
    // view of the sub-batch this lane is working on (set per step)
    const int* slots;
    const void* pcm;
    int pcm_fmt;
    const __half* feats;                 // non-null: feature-input mode
    int* len_in;
    int* cpos_in;                        // [B][2] ring positions of the conv caches seen by this step (full / reduced rate)
    float* lp_out;
    int* tok_out;
    float* aux_out;
  };
  std::vector<Lane> lanes;
  // Lanes: a step is cut into concurrent sub-batches when one lane's N = 384 GEMMs (3 column tiles per 128-row tile) no
  // longer fit in one round of single-tile CTAs - 576 streams x 10 frames = 135 tiles run in one lane (1.66 ms; two lanes
  // 1.91 ms), 640 streams = 150 tiles run in two (1.92 ms; one lane 2.15 ms).  tone_config.lane_min_batch > 0 replaces the
  // rule by a fixed streams-per-lane threshold.
  int n_lanes = 2, lane_min_batch = 0;
  cudaEvent_t fork_ev = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;   // H2D / D2H copy streams of the pipelined step
  cudaStream_t s_cap = nullptr;                   // graph capture happens on a stream of its own
  cudaEvent_t ev_sync = nullptr, ev_last = nullptr;   // splice of caller-stream launches into the engine's timeline

  std::unordered_map<uint64_t, cudaGraphExec_t> graphs;
  int launches = 0, launches_per_step = 0;
  bool pdl = true;      // programmatic dependent launch between the kernels of a step (TONE_FLAG_NO_PDL disables)
  // Large dense GEMMs (>= persist_min_tiles output tiles) run as a persistent one-CTA-per-SM kernel with the epilogue
  // overlapped with the next tile's main loop (gemm_tc_persist_kernel); 0 = never.
  int persist_min_tiles = 0, persist_ctas = 0;
  int lane_ctas = 0;    // persistent-kernel CTAs while a step runs in more than one lane (tone_config.persist_ctas)
  int cur_ctas = 148;   // CTAs a persistent kernel of the step being enqueued may occupy
  int cur_lanes = 1;    // lanes of the step being enqueued
  // gated kinds (N % 256 == 0): 0 = 128-wide tiles, 1 = 256-wide, 2 = 256-wide on CTA pairs (cta_group::2).  Measured
  // (profiles/r01_persistent_gemm.md): the pair form runs the feed-forward up GEMM at 73 % of the sustained bf16 peak when
  // it has the GPU to itself, but with two lanes in flight the 256-wide single-CTA form gives the faster step.
  int persist_mode = 1;
  int split_k = 0;         // 0 = fill the SMs once
  // Feed-forward module as ONE kernel per row tile (ff_fused.cuh) from ff_fused_min_rows rows per lane on:
  // 0 = off (default), 1 = one CTA per 128 rows, 2 = CTA pairs (cta_group::2, 256 rows per pair).  Opt-in: parity-green,
  // but at 1024 streams per GPU 80 row tiles cannot fill 148 SMs (profiles/r02_fused_ff.md).
  int ff_fused = 0, ff_fused_min_rows = 2048;
  bool fuse_vatt = true;   // score-sharing layers: V projection + P.V in one kernel
  // score-sharing attention layers as ONE kernel per tile of whole streams (att_fused.cuh) from this many rows per lane (0 = never)
  int att_block_min_rows = 16384;  // pays from ~4096 streams per GPU (profiles/r02_experiments.md)
  // feed-forward 1 adds straight into the residual stream and norm_self_att becomes a row scale inside the projection
  // GEMMs from this many rows per lane (0 = never)
  int lazy_norm_min_rows = 4096;
  // N = 384 projections (feed-forward down, attention out, pointwise conv 2) as the row-owner CTA-pair kernel
  // (rowgemm.cuh) from this many rows per lane (0 = never)
  int rowgemm_min_rows = 8192;     // pays once a lane's grid of row pairs covers half the SMs (2048+ streams per GPU)
  int att_pipe_min_batch = 256;  // streams per lane from which the recompute attention layers run as the pipelined persistent kernel
  int dw_pipe_min_batch = 128;   // streams per lane from which the depthwise conv runs as the pipelined persistent kernel (0 = never)
  int num_sms = 148;
};

// ------------------------------------------------------------------------------------------------ small helpers
template <typename Tp>
static int dev_alloc(tone_engine* e, Tp** p, size_t n) {
  void* q = nullptr;
  CK(cudaMalloc(&q, std::max<size_t>(n * sizeof(Tp), 256)));
  CK(cudaMemset(q, 0, std::max<size_t>(n * sizeof(Tp), 256)));
  e->allocs.push_back(q);
  *p = (Tp*)q;
  return 0;
}
static void* arena_take(tone_engine* e, size_t bytes) {
  size_t off = (e->w_used + 255) & ~size_t(255);
  if (off + bytes > e->w_cap) return nullptr;
  e->w_used = off + bytes;
  return e->w_arena + off;
}
static int upload_f32(tone_engine* e, const std::vector<float>& v, float** out) {
  void* p = arena_take(e, v.size() * 4);
  if (!p) return fail(TONE_ENOMEM, "weight arena exhausted");
  CK(cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  *out = (float*)p;
  return 0;
}
static int make_map(tone_engine* e, CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool weight) {
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = e->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B,
                         weight ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TONE_ECUDA, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
  return 0;
}
static int make_map_2d(tone_engine* e, CUtensorMap* m, void* base, uint64_t rows, uint64_t K, uint32_t box_rows,
                       bool weight) {
  uint64_t dims[2] = {K, rows}, st[1] = {K * 2};
  uint32_t box[2] = {64, box_rows};
  return make_map(e, m, base, 2, dims, st, box, weight);
}
static int upload_mat(tone_engine* e, const std::vector<float>& w, int N, int K, int box_rows, WeightMat* out) {
  std::vector<uint16_t> b((size_t)N * K);
  for (size_t i = 0; i < b.size(); ++i) b[i] = f2bf(w[i]);
  void* p = arena_take(e, b.size() * 2);
  if (!p) return fail(TONE_ENOMEM, "weight arena exhausted");
  CK(cudaMemcpy(p, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  out->ptr = (bf16*)p;
  out->N = N;
  out->K = K;
  if (N % 128 == 0) {
    int rc = make_map_2d(e, &out->map128, p, N, K, 128, true);
    if (rc) return rc;
  }
  if (N % 64 == 0) {
    int rc = make_map_2d(e, &out->map64, p, N, K, 64, true);
    if (rc) return rc;
  }
  return make_map_2d(e, &out->map, p, N, K, box_rows, true);
}
static const HostTensor* W(tone_engine* e, const std::string& name) {
  auto it = e->host_w.find(name);
  return it == e->host_w.end() ? nullptr : &it->second;
}
#define NEEDW(var, name)                                                         \
  const HostTensor* var = W(e, name);                                            \
  if (!var) return fail(TONE_ESTATE, "weight '%s' was not loaded", std::string(name).c_str());

// rows of Wa / Wb interleaved in blocks of HW so that one BN = 2*HW output tile holds matching columns
static std::vector<float> interleave_rows(const std::vector<float>& a, const std::vector<float>& b, int Nh, int K,
                                          int HW) {
  std::vector<float> o((size_t)2 * Nh * K);
  for (int t = 0; t < Nh / HW; ++t)
    for (int r = 0; r < HW; ++r) {
      memcpy(&o[((size_t)t * 2 * HW + r) * K], &a[((size_t)t * HW + r) * K], (size_t)K * 4);
      memcpy(&o[((size_t)t * 2 * HW + HW + r) * K], &b[((size_t)t * HW + r) * K], (size_t)K * 4);
    }
  return o;
}
// W[n][k] * g[k]: a norm gain on the GEMM's input folded into the weight columns
static std::vector<float> fold_cols(std::vector<float> w, const std::vector<float>& g, int N, int K) {
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) w[(size_t)n * K + k] *= g[k];
  return w;
}
static std::vector<float> concat(std::initializer_list<const std::vector<float>*> parts) {
  std::vector<float> o;
  for (auto p : parts) o.insert(o.end(), p->begin(), p->end());
  return o;
}

// ------------------------------------------------------------------------------------------------ create / destroy
static const int BN_SWIGLU = 128, BN_RESID = 32, BN_GLU = 64, BN_STORE = 64, BN_CONV = 128, BN_KV = 64, BN_PART = 128;
// Narrow N tiles fill the SMs when there are only a few M tiles; from BIG_M rows on the 128-wide tile is used for
// every GEMM so that the A tile is not re-read by 12 N-tile CTAs.
static const int BIG_M = 2048;

extern "C" const char* tone_last_error(void) { return g_err; }
// internal: lets server.cu report through the same thread-local buffer
extern "C" void tone_internal_set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

static int create_impl(tone_engine* e, const tone_config* cfg, const cudaDeviceProp& prop);

extern "C" int tone_create(const tone_config* cfg, tone_engine** out) {
  if (!cfg || !out) return fail(TONE_EINVAL, "null argument");
  if (cfg->chunk_samples != 2400 && cfg->chunk_samples != 3200)
    return fail(TONE_EINVAL, "chunk_samples must be 2400 (300 ms) or 3200 (400 ms), got %d", cfg->chunk_samples);
  if (cfg->max_slots < 1 || cfg->max_batch < 1 || cfg->max_batch > cfg->max_slots)
    return fail(TONE_EINVAL, "need 1 <= max_batch <= max_slots");
  if (cfg->lanes < 0 || cfg->lanes > 4 || cfg->persist_mode < 0 || cfg->persist_mode > 3 || cfg->split_k < 0 ||
      cfg->split_k > MAX_SPLITS || cfg->lane_min_batch < 0 || cfg->fused_ff < 0 || cfg->fused_ff > 3 || cfg->fused_ff_min_rows < 0 ||
      cfg->att_block_min_rows < -1 || cfg->lazy_norm_min_rows < -1 || cfg->dw_pipe_min_batch < -1 ||
      cfg->att_pipe_min_batch < -1 || cfg->rowgemm_min_rows < -1)
    return fail(TONE_EINVAL, "tuning field out of range (lanes 0..4, persist_mode 0..3, split_k 0..%d)", (int)MAX_SPLITS);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(TONE_EINVAL, "device %d of %d", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(TONE_ECUDA, "this library is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  tone_engine* e = new tone_engine();
  const int rc = create_impl(e, cfg, prop);
  if (rc) {           // every failure path releases what was allocated so far (g_err keeps the reason)
    tone_destroy(e);
    return rc;
  }
  *out = e;
  return TONE_OK;
}

template <typename Tp>
static int pinned_alloc(Tp** p, size_t n) {
  CK(cudaMallocHost((void**)p, std::max<size_t>(n * sizeof(Tp), 64)));
  memset(*p, 0, std::max<size_t>(n * sizeof(Tp), 64));
  return 0;
}

static int create_impl(tone_engine* e, const tone_config* cfg, const cudaDeviceProp& prop) {
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  e->pdl = !(cfg->flags & TONE_FLAG_NO_PDL);
  e->fuse_vatt = !(cfg->flags & TONE_FLAG_NO_FUSED_VATT);
  if (cfg->att_block_min_rows) e->att_block_min_rows = std::max(0, cfg->att_block_min_rows);      // -1 = never
  if (cfg->lazy_norm_min_rows) e->lazy_norm_min_rows = std::max(0, cfg->lazy_norm_min_rows);
  if (cfg->dw_pipe_min_batch) e->dw_pipe_min_batch = std::max(0, cfg->dw_pipe_min_batch);
  if (cfg->att_pipe_min_batch) e->att_pipe_min_batch = std::max(0, cfg->att_pipe_min_batch);
  if (cfg->rowgemm_min_rows) e->rowgemm_min_rows = std::max(0, cfg->rowgemm_min_rows);
  e->persist_mode = cfg->persist_mode ? cfg->persist_mode - 1 : 1;
  e->split_k = cfg->split_k;
  if (cfg->fused_ff) e->ff_fused = cfg->fused_ff - 1;
  if (cfg->fused_ff_min_rows) e->ff_fused_min_rows = cfg->fused_ff_min_rows;
  if (cfg->lanes) e->n_lanes = cfg->lanes;
  if (cfg->lane_min_batch) e->lane_min_batch = cfg->lane_min_batch;
  e->C = cfg->chunk_samples;
  e->F = e->C / HOP;
  e->T = (e->F + SUB2_ROWS - 11) / 3 + 1;
  e->T2 = (e->T + 1 - 3) / 2 + 1;
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->s_cap, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->ev_sync, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return fail(TONE_ECUDA, "cuTensorMapEncodeTiled not available");
    e->encode = (PFN_encodeTiled)fn;
  }
  // weight arena: ~72M params in bf16 + expanded conv matrices + fp32 vectors
  e->w_cap = (size_t)256 << 20;
  CK(cudaMalloc((void**)&e->w_arena, e->w_cap));
  CK(cudaMemset(e->w_arena, 0, e->w_cap));

  const size_t S = cfg->max_slots;
  int rc = 0;
  rc |= dev_alloc(e, &e->st_pre, S * HOP);
  rc |= dev_alloc(e, &e->st_feat, S * FEAT_ROWS_MAX * N_MELS);
  rc |= dev_alloc(e, &e->st_x1, S * X1_ROWS_MAX * X1_ROW);
  rc |= dev_alloc(e, &e->st_kv14, S * KV_ROWS_MAX * D_MODEL);
  rc |= dev_alloc(e, &e->st_kv15, S * KV_ROWS_MAX * D_MODEL);
  rc |= dev_alloc(e, &e->st_conv, S * N_LAYERS * CONV_S * D_MODEL);
  rc |= dev_alloc(e, &e->st_cpos, S * 2);
  rc |= dev_alloc(e, &e->st_red, S * D_MODEL);
  rc |= dev_alloc(e, &e->st_len, S);
  rc |= dev_alloc(e, &e->st_ph, S);
  rc |= dev_alloc(e, &e->st_ring, S * PH_CAP);
  if (rc) return rc;
  e->slot_used.assign(S, 0);
  e->slot_stamp.assign(S, 0);
  for (int i = (int)S - 1; i >= 0; --i) e->free_slots.push_back(i);

  const size_t Bm = cfg->max_batch;
  e->rows_alloc = (int)(((Bm * MAX_T + 127) / 128) * 128);
  const size_t R = e->rows_alloc;
  // phrase area of one set: header | PH_PER_STREAM records per stream | text pool (a stream's buffer never exceeds
  // ~2050 frames, plus 6 frames of overlap per phrase)
  e->ph_bytes = sizeof(PhHeader) + Bm * PH_PER_STREAM * sizeof(PhRecord) + Bm * 2304;
  for (int k = 0; k <= tone_engine::PIPE; ++k) {
    tone_engine::IoSet& io = e->io[k];
    const size_t pcm_bytes = Bm * e->C * (k == tone_engine::PIPE ? 4 : 2);
    rc |= dev_alloc(e, &io.d_slots, Bm);
    rc |= dev_alloc(e, (char**)&io.d_pcm, pcm_bytes);
    rc |= dev_alloc(e, &io.d_last, Bm);
    rc |= dev_alloc(e, &io.d_len_in, Bm);
    rc |= dev_alloc(e, &io.d_cpos_in, 2 * Bm);
    rc |= dev_alloc(e, &io.d_tokens, R);
    rc |= dev_alloc(e, &io.d_aux, R * 2);
    rc |= dev_alloc(e, &io.d_logprobs, R * N_CLASSES);
    rc |= dev_alloc(e, &io.d_ph, e->ph_bytes);
    if (rc) return rc;
    RC(pinned_alloc(&io.p_slots, Bm));
    RC(pinned_alloc(&io.p_pcm, Bm * e->C));
    RC(pinned_alloc(&io.p_last, Bm));
    RC(pinned_alloc(&io.p_tokens, Bm * MAX_T));
    RC(pinned_alloc(&io.p_logprobs, Bm * MAX_T * N_CLASSES));
    RC(pinned_alloc(&io.p_aux, Bm * MAX_T * 2));
    RC(pinned_alloc(&io.p_ph, e->ph_bytes));
    CK(cudaEventCreateWithFlags(&io.ev_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&io.ev_step, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&io.ev_out, cudaEventDisableTiming));
  }
  rc |= dev_alloc(e, &e->d_feats, Bm * N_MELS * MAX_FRAMES);
  RC(pinned_alloc(&e->p_feats, Bm * N_MELS * MAX_FRAMES));
  rc |= dev_alloc(e, &e->d_ring, (size_t)tone_engine::RING * std::max<size_t>(Bm, STATE_IO_CHUNK));
  RC(pinned_alloc(&e->p_ring, (size_t)tone_engine::RING * std::max<size_t>(Bm, STATE_IO_CHUNK)));
  for (int i = 0; i < tone_engine::RING; ++i) CK(cudaEventCreateWithFlags(&e->ev_ring[i], cudaEventDisableTiming));
  rc |= dev_alloc(e, &e->d_state_io, (size_t)STATE_IO_CHUNK * TONE_STATE_SIZE);
  RC(pinned_alloc(&e->p_state_io, (size_t)STATE_IO_CHUNK * TONE_STATE_SIZE));
  e->lanes.resize(e->n_lanes);
  for (int li = 0; li < e->n_lanes; ++li) {
    tone_engine::Lane& ln = e->lanes[li];
    rc |= dev_alloc(e, &ln.r_full, R * D_MODEL);
    rc |= dev_alloc(e, &ln.r_red, R * D_MODEL);
    rc |= dev_alloc(e, &ln.qkv, std::max(R * 3 * D_MODEL, Bm * (MHSA_S + MAX_T) * 2 * D_MODEL + R * D_MODEL));
    rc |= dev_alloc(e, &ln.P, Bm * N_HEADS * MAX_T * (MHSA_S + MAX_T));
    rc |= dev_alloc(e, &ln.part, (size_t)e->max_splits * R * D_MODEL);
    rc |= dev_alloc(e, &ln.n, R * D_MODEL);
    rc |= dev_alloc(e, &ln.h, R * D_FF);
    rc |= dev_alloc(e, &ln.ctx, R * D_MODEL);
    rc |= dev_alloc(e, &ln.g, R * D_MODEL);
    rc |= dev_alloc(e, &ln.ebuf, R * D_MODEL);
    rc |= dev_alloc(e, &ln.c1, R * SUB_OUT);
    rc |= dev_alloc(e, &ln.m_red, R * D_FF);
    rc |= dev_alloc(e, &ln.rb, R * D_MODEL);
    rc |= dev_alloc(e, &ln.ss, R * 12);
    if (rc) return rc;
    if (li > 0) {
      CK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    }
  }
  CK(cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming));
  if (rc) return rc;

  // activation-side tensor maps
  {
    uint64_t d[3] = {N_MELS, FEAT_ROWS_MAX, S}, s[2] = {N_MELS * 2, (uint64_t)FEAT_ROWS_MAX * N_MELS * 2};
    uint32_t bx[3] = {64, (uint32_t)e->F, 1}, bw[3] = {64, (uint32_t)e->F, (uint32_t)(128 / e->F)};
    if ((rc = make_map(e, &e->m_feat, e->st_feat, 3, d, s, bx, false))) return rc;
    if ((rc = make_map(e, &e->w_feat, e->st_feat, 3, d, s, bw, false))) return rc;
  }
  {  // x1 as [slot][16 row triples][3 rows][1408]: conv1 frame t, kernel row kt reads row 3t+kt = triple t+kt/3, row kt%3
    uint64_t d[4] = {X1_ROW, 3, X1_ROWS_MAX / 3, S};
    uint64_t s[3] = {X1_ROW * 2, 3 * X1_ROW * 2, (uint64_t)X1_ROWS_MAX * X1_ROW * 2};
    uint32_t bx[4] = {64, 1, (uint32_t)e->T, 1}, bw[4] = {64, 1, (uint32_t)e->T, (uint32_t)(128 / e->T)};
    if ((rc = make_map(e, &e->m_x1, e->st_x1, 4, d, s, bx, false))) return rc;
    if ((rc = make_map(e, &e->w_x1, e->st_x1, 4, d, s, bw, false))) return rc;
  }
  {
    uint64_t d[3] = {D_MODEL, KV_ROWS_MAX, S}, s[2] = {D_MODEL * 2, (uint64_t)KV_ROWS_MAX * D_MODEL * 2};
    uint32_t b14[3] = {64, (uint32_t)(MHSA_S / 2 + e->T2), 1}, b15[3] = {64, (uint32_t)(MHSA_S + e->T), 1};
    if ((rc = make_map(e, &e->m_kv14, e->st_kv14, 3, d, s, b14, false))) return rc;
    if ((rc = make_map(e, &e->m_kv15, e->st_kv15, 3, d, s, b15, false))) return rc;
    uint32_t w14[3] = {64, b14[1], 128 / b14[1]}, w15[3] = {64, b15[1], 128 / b15[1]};
    if ((rc = make_map(e, &e->w_kv14, e->st_kv14, 3, d, s, w14, false))) return rc;
    if ((rc = make_map(e, &e->w_kv15, e->st_kv15, 3, d, s, w15, false))) return rc;
  }
  for (auto& ln : e->lanes) {
    if ((rc = make_map_2d(e, &ln.m_n, ln.n, R, D_MODEL, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_h, ln.h, R, D_FF, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_ctx, ln.ctx, R, D_MODEL, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_e, ln.ebuf, R, D_MODEL, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_c1, ln.c1, R, SUB_OUT, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_mred, ln.m_red, R, D_FF, 128, false))) return rc;
    if ((rc = make_map_2d(e, &ln.m_rb, ln.rb, R, D_MODEL, 128, false))) return rc;
  }

  CK((configure_gemm_tc<G_SWIGLU, BN_SWIGLU>()));
  CK((configure_gemm_tc<G_RESID, BN_RESID>()));
  CK((configure_gemm_tc<G_RESID, 64>()));
  CK((configure_gemm_tc<G_GLU, BN_GLU>()));
  CK((configure_gemm_tc<G_STORE_F32, BN_STORE>()));
  CK((configure_gemm_tc<G_STORE_F32, 32>()));
  CK((configure_gemm_tc<G_STORE_F32, 128>()));
  CK((configure_gemm_tc<G_CONV0, BN_CONV>()));
  CK((configure_gemm_tc<G_CONV1, BN_CONV>()));
  CK((configure_gemm_tc<G_KV, BN_KV>()));
  CK((configure_gemm_tc<G_KV, 128>()));
  CK((configure_gemm_tc<G_DECODER, DEC_PAD>()));
  CK((configure_gemm_tc<G_PARTIAL, BN_PART>()));
  CK((configure_gemm_tc<G_VATT, D_HEAD>()));
  CK((configure_gemm_tc<G_RESID, 128>()));
  CK((configure_gemm_tc<G_GLU, 128>()));
  CK((configure_gemm_tc_persist<G_SWIGLU, 1, false>()));
  CK((configure_gemm_tc_persist<G_SWIGLU, 2, false>()));
  CK((configure_gemm_tc_persist<G_SWIGLU, 2, true>()));
  CK((configure_gemm_tc_persist<G_GLU, 1, false>()));
  CK((configure_gemm_tc_persist<G_GLU, 2, false>()));
  CK((configure_gemm_tc_persist<G_GLU, 2, true>()));
  CK((configure_gemm_tc_persist<G_RESID, 1, false>()));
  CK((configure_gemm_tc_persist<G_STORE_F32, 1, false>()));
  CK((configure_gemm_tc_persist<G_PARTIAL, 1, false>()));
  CK((configure_ff_fused<false>()));
  CK((configure_ff_fused<true>()));
  CK(configure_att_fused());
  CK(configure_rowgemm());
  e->persist_ctas = e->num_sms;
  e->lane_ctas = cfg->persist_ctas > 0 ? std::min(cfg->persist_ctas, e->num_sms) : e->num_sms;
  e->persist_min_tiles = cfg->persist_min_tiles < 0 ? 0 : (cfg->persist_min_tiles ? cfg->persist_min_tiles : e->num_sms + 1);
  CK(cudaFuncSetAttribute(begin_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048));
  CK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_V_SMEM));
  CK(cudaFuncSetAttribute(attention_pipe_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
  CK(cudaFuncSetAttribute(attention_pipe_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  CK(cudaFuncSetAttribute(dwconv_pipe_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWP_SMEM));
  CK(cudaFuncSetAttribute(dwconv_pipe_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWP_SMEM));
  CK(cudaFuncSetAttribute(dwconv_pipe_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWP_SMEM));
  CK(cudaFuncSetAttribute(dwconv_pipe_kernel<MAX_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWP_SMEM));
  CK(cudaDeviceSynchronize());   // the memsets above ran on the legacy stream; the engine streams are non-blocking
  return TONE_OK;
}

extern "C" void tone_destroy(tone_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& ln : e->lanes) {
    if (ln.stream) cudaStreamDestroy(ln.stream);
    if (ln.done) cudaEventDestroy(ln.done);
  }
  if (e->fork_ev) cudaEventDestroy(e->fork_ev);
  if (e->ev_sync) cudaEventDestroy(e->ev_sync);
  if (e->ev_last) cudaEventDestroy(e->ev_last);
  for (int i = 0; i < tone_engine::RING; ++i)
    if (e->ev_ring[i]) cudaEventDestroy(e->ev_ring[i]);
  for (void* p : e->allocs) cudaFree(p);
  if (e->w_arena) cudaFree(e->w_arena);
  for (auto& io : e->io) {
    if (io.p_slots) cudaFreeHost(io.p_slots);
    if (io.p_pcm) cudaFreeHost(io.p_pcm);
    if (io.p_last) cudaFreeHost(io.p_last);
    if (io.p_tokens) cudaFreeHost(io.p_tokens);
    if (io.p_logprobs) cudaFreeHost(io.p_logprobs);
    if (io.p_aux) cudaFreeHost(io.p_aux);
    if (io.p_ph) cudaFreeHost(io.p_ph);
    if (io.ev_in) cudaEventDestroy(io.ev_in);
    if (io.ev_step) cudaEventDestroy(io.ev_step);
    if (io.ev_out) cudaEventDestroy(io.ev_out);
  }
  if (e->p_feats) cudaFreeHost(e->p_feats);
  if (e->p_ring) cudaFreeHost(e->p_ring);
  if (e->p_state_io) cudaFreeHost(e->p_state_io);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->s_cap) cudaStreamDestroy(e->s_cap);
  delete e;
}

extern "C" int tone_get_info(const tone_engine* e, tone_info* o) {
  if (!e || !o) return fail(TONE_EINVAL, "null argument");
  o->chunk_samples = e->C;
  o->frames_out = e->T;
  o->n_classes = N_CLASSES;
  o->state_size = TONE_STATE_SIZE;
  o->max_slots = e->cfg.max_slots;
  o->max_batch = e->cfg.max_batch;
  o->launches_per_step = e->launches_per_step;
  o->n_taps = 1 + N_LAYERS;
  o->state_bytes_per_slot = (int64_t)HOP * 2 + FEAT_ROWS_MAX * N_MELS * 2 + (int64_t)X1_ROWS_MAX * X1_ROW * 2 +
                            2LL * KV_ROWS_MAX * D_MODEL * 2 + (int64_t)N_LAYERS * CONV_S * D_MODEL * 2 + D_MODEL * 4 + 4;
  o->weight_bytes = (int64_t)e->w_used;
  o->pipeline_depth = tone_engine::PIPE;
  o->max_phrases_per_step = PH_PER_STREAM * e->cfg.max_batch;
  return TONE_OK;
}

// ------------------------------------------------------------------------------------------------ weights
extern "C" int tone_load_weight(tone_engine* e, const char* name, const float* data, const int64_t* shape,
                                int32_t ndim) {
  if (!e || !name || !data || !shape || ndim < 1 || ndim > 4) return fail(TONE_EINVAL, "bad argument");
  if (e->finalized) return fail(TONE_ESTATE, "weights already finalized");
  std::string nm(name);
  if (nm.rfind("tone.", 0) == 0) nm = nm.substr(5);
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 1) return fail(TONE_EINVAL, "bad shape for %s", name);
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(data, data + n);
  e->host_w[nm] = std::move(t);
  return TONE_OK;
}

static int expect_shape(const HostTensor* t, const char* name, std::initializer_list<int64_t> s) {
  std::vector<int64_t> v(s);
  if (t->shape != v) return fail(TONE_EINVAL, "weight '%s' has an unexpected shape", name);
  return 0;
}

static int bn_fold(tone_engine* e, const std::string& bn, const std::vector<float>& conv_bias, int n,
                   std::vector<float>& alpha, std::vector<float>& beta) {
  NEEDW(w, bn + "weight");
  NEEDW(b, bn + "bias");
  NEEDW(m, bn + "running_mean");
  NEEDW(v, bn + "running_var");
  alpha.resize(n);
  beta.resize(n);
  for (int i = 0; i < n; ++i) {
    float a = w->data[i] / sqrtf(v->data[i] + 1e-5f);  // BatchNorm eval, eps 1e-5
    alpha[i] = a;
    beta[i] = (conv_bias[i] - m->data[i]) * a + b->data[i];
  }
  return 0;
}

static int finalize_frontend(tone_engine* e) {
  // fused (pre-emphasis x symmetric Hann x DFT-160) basis, stored [j][k] (feats.py:66-80)
  const double PI = 3.14159265358979323846;
  std::vector<double> basis((size_t)162 * WIN);
  for (int k = 0; k < N_BINS; ++k)
    for (int j = 0; j < WIN; ++j) {
      double hann = 0.5 - 0.5 * cos(2.0 * PI * j / (WIN - 1));
      double ang = 2.0 * PI * (double)k * j / WIN;
      basis[(size_t)k * WIN + j] = cos(ang) * hann;
      basis[(size_t)(N_BINS + k) * WIN + j] = -sin(ang) * hann;
    }
  // fp16 hi/lo split, [2][BASIS_N][BASIS_LD], zero padded: the frontend kernel feeds it to mma.sync with the fp16
  // waveform; hi + lo carries ~22 bits of the fp32 basis the reference uses
  std::vector<uint16_t> bt((size_t)2 * BASIS_N * BASIS_LD, 0);
  for (int k = 0; k < 162; ++k)
    for (int j = 0; j < WIN; ++j) {
      double v = basis[(size_t)k * WIN + j];
      if (j + 1 < WIN) v -= 0.97 * basis[(size_t)k * WIN + j + 1];
      if (j == 0) v -= 0.97 * basis[(size_t)k * WIN];
      const float vf = (float)v;
      const uint16_t hi = f2h(vf);
      const uint16_t lo = f2h(vf - h2f(hi));
      bt[(size_t)k * BASIS_LD + j] = hi;
      bt[((size_t)BASIS_N + k) * BASIS_LD + j] = lo;
    }
  int rc = 0;
  {
    void* pb = arena_take(e, bt.size() * 2);
    if (!pb) return fail(TONE_ENOMEM, "weight arena exhausted");
    CK(cudaMemcpy(pb, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice));
    e->basis = (__half*)pb;
  }
  // slaney mel filterbank in CSR form (feats.py:82-93; torchaudio melscale_fbanks, slaney scale + norm)
  auto hz2mel = [](double f) {
    const double f_sp = 200.0 / 3, logstep = log(6.4) / 27.0;
    return f >= 1000.0 ? 1000.0 / f_sp + log(f / 1000.0) / logstep : f / f_sp;
  };
  auto mel2hz = [](double m) {
    const double f_sp = 200.0 / 3, logstep = log(6.4) / 27.0, mlm = 1000.0 / f_sp;
    return m >= mlm ? 1000.0 * exp(logstep * (m - mlm)) : f_sp * m;
  };
  std::vector<double> fpts(N_MELS + 2);
  const double m0 = hz2mel(0.0), m1 = hz2mel(4000.0);
  for (int i = 0; i < N_MELS + 2; ++i) fpts[i] = mel2hz(m0 + (m1 - m0) * i / (N_MELS + 1));
  std::vector<int> start(N_MELS + 1, 0);
  std::vector<unsigned char> bins;
  std::vector<float> wts;
  for (int m = 0; m < N_MELS; ++m) {
    start[m] = (int)bins.size();
    for (int k = 0; k < N_BINS; ++k) {
      double fr = 4000.0 * k / (N_BINS - 1);
      double down = (fr - fpts[m]) / (fpts[m + 1] - fpts[m]);
      double up = (fpts[m + 2] - fr) / (fpts[m + 2] - fpts[m + 1]);
      double v = std::max(0.0, std::min(down, up)) * (2.0 / (fpts[m + 2] - fpts[m]));
      if (v > 0.0) {
        bins.push_back((unsigned char)k);
        wts.push_back((float)v);
      }
    }
  }
  start[N_MELS] = (int)bins.size();
  if ((rc = upload_f32(e, wts, &e->mel_w))) return rc;
  void* p = arena_take(e, start.size() * 4);
  CK(cudaMemcpy(p, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
  e->mel_start = (int*)p;
  p = arena_take(e, bins.size());
  CK(cudaMemcpy(p, bins.data(), bins.size(), cudaMemcpyHostToDevice));
  e->mel_bin = (unsigned char*)p;
  // RoPE tables for positions -30 .. 12 (submodules.py:117-139): angle = pos * 10000^(-2i/32), fp32
  std::vector<float> rc_((size_t)(MHSA_S + MAX_T) * 16), rs_((size_t)(MHSA_S + MAX_T) * 16);
  for (int p_ = 0; p_ < MHSA_S + MAX_T; ++p_)
    for (int i = 0; i < 16; ++i) {
      float inv = 1.0f / powf(10000.0f, (float)(2 * i) / 32.0f);
      float ang = (float)(p_ - MHSA_S) * inv;
      rc_[p_ * 16 + i] = cosf(ang);
      rs_[p_ * 16 + i] = sinf(ang);
    }
  if ((rc = upload_f32(e, rc_, &e->rope_cos))) return rc;
  return upload_f32(e, rs_, &e->rope_sin);
}

static int finalize_pre_encode(tone_engine* e) {
  const std::string P = "encoder.pre_encode.";
  int rc;
  NEEDW(pn, P + "pre_norm.weight");
  if ((rc = upload_f32(e, pn->data, &e->pre_norm_g))) return rc;
  NEEDW(on, P + "out_norm.weight");
  if ((rc = upload_f32(e, on->data, &e->out_norm_g))) return rc;
  // conv0 (32,1,11,21) as a banded [f*32+o][kt*64+fin] matrix: the GEMM row is a frame, K walks 11 feature rows
  NEEDW(w0, P + "conv.0.0.weight");
  NEEDW(b0, P + "conv.0.0.bias");
  if ((rc = expect_shape(w0, "conv.0.0.weight", {32, 1, 11, 21}))) return rc;
  {
    std::vector<float> t((size_t)1408 * 704, 0.f);
    for (int f = 0; f < 44; ++f)
      for (int o = 0; o < 32; ++o)
        for (int kt = 0; kt < 11; ++kt)
          for (int kf = 0; kf < 21; ++kf)
            t[(size_t)(f * 32 + o) * 704 + kt * 64 + f + kf] = w0->data[((size_t)o * 11 + kt) * 21 + kf];
    if ((rc = upload_mat(e, t, 1408, 704, BN_CONV, &e->conv0_w))) return rc;
    std::vector<float> al, be;
    if ((rc = bn_fold(e, P + "conv.0.1.", b0->data, 32, al, be))) return rc;
    if ((rc = upload_f32(e, al, &e->conv0_alpha))) return rc;
    if ((rc = upload_f32(e, be, &e->conv0_beta))) return rc;
  }
  // conv1 (64,32,11,11): two adjacent output positions j in {0,1} share a 12-position window:
  //   W2[j*64+o][kt*384 + d*32 + c] = w[o][c][kt][d-j]  for 0 <= d-j < 11
  NEEDW(w1, P + "conv.1.0.weight");
  NEEDW(b1, P + "conv.1.0.bias");
  if ((rc = expect_shape(w1, "conv.1.0.weight", {64, 32, 11, 11}))) return rc;
  {
    const int K = 11 * 384;
    std::vector<float> t((size_t)128 * K, 0.f);
    for (int j = 0; j < 2; ++j)
      for (int o = 0; o < 64; ++o)
        for (int kt = 0; kt < 11; ++kt)
          for (int kf = 0; kf < 11; ++kf)
            for (int c = 0; c < 32; ++c)
              t[(size_t)(j * 64 + o) * K + kt * 384 + (kf + j) * 32 + c] =
                  w1->data[(((size_t)o * 32 + c) * 11 + kt) * 11 + kf];
    if ((rc = upload_mat(e, t, 128, K, BN_CONV, &e->conv1_w))) return rc;
    std::vector<float> al, be;
    if ((rc = bn_fold(e, P + "conv.1.1.", b1->data, 64, al, be))) return rc;
    if ((rc = upload_f32(e, al, &e->conv1_alpha))) return rc;
    if ((rc = upload_f32(e, be, &e->conv1_beta))) return rc;
  }
  // out Linear (384, 2176): reference flattens o*34+f (conformer_blocks.py:649); our rows are f*64+o
  NEEDW(wo, P + "out.weight");
  if ((rc = expect_shape(wo, "out.weight", {384, 2176}))) return rc;
  {
    std::vector<float> t((size_t)384 * SUB_OUT);
    for (int n = 0; n < 384; ++n)
      for (int o = 0; o < 64; ++o)
        for (int f = 0; f < 34; ++f) t[(size_t)n * SUB_OUT + f * 64 + o] = wo->data[(size_t)n * SUB_OUT + o * 34 + f];
    if ((rc = upload_mat(e, t, 384, SUB_OUT, BN_STORE, &e->out_w))) return rc;
  }
  return 0;
}

static int finalize_layer(tone_engine* e, int l) {
  const std::string Lp = "encoder.layers." + std::to_string(l) + ".";
  LayerW& L = e->L[l];
  int rc;
  auto vec = [&](const std::string& nm, float** out) -> int {
    NEEDW(t, Lp + nm);
    return upload_f32(e, t->data, out);
  };
  if ((rc = vec("norm_feed_forward1.weight", &L.n_ff1))) return rc;
  if ((rc = vec("norm_self_att.weight", &L.n_att))) return rc;
  if ((rc = vec("norm_conv.weight", &L.n_conv))) return rc;
  if ((rc = vec("norm_feed_forward2.weight", &L.n_ff2))) return rc;
  if ((rc = vec("norm_out.weight", &L.n_out))) return rc;
  for (int k = 0; k < 2; ++k) {
    const std::string ff = Lp + (k == 0 ? "feed_forward1." : "feed_forward2.");
    NEEDW(w1, ff + "linear1.weight");
    NEEDW(b1, ff + "linear1.bias");
    NEEDW(wv, ff + "linearv.weight");
    NEEDW(bv, ff + "linearv.bias");
    NEEDW(w2, ff + "linear2.weight");
    NEEDW(b2, ff + "linear2.bias");
    if ((rc = expect_shape(w1, "linear1.weight", {D_FF, D_MODEL}))) return rc;
    if ((rc = expect_shape(w2, "linear2.weight", {D_MODEL, D_FF}))) return rc;
    std::vector<float> w1d = w1->data, wvd = wv->data;
    if (k == 1) {   // second feed-forward: its RMSNorm gain is folded into the weight columns (row-scale RMSNorm)
      NEEDW(gff2, Lp + "norm_feed_forward2.weight");
      for (int n = 0; n < D_FF; ++n)
        for (int c = 0; c < D_MODEL; ++c) {
          w1d[(size_t)n * D_MODEL + c] *= gff2->data[c];
          wvd[(size_t)n * D_MODEL + c] *= gff2->data[c];
        }
    }
    std::vector<float> up = interleave_rows(w1d, wvd, D_FF, D_MODEL, BN_SWIGLU / 2);
    std::vector<float> upb = interleave_rows(b1->data, bv->data, D_FF, 1, BN_SWIGLU / 2);
    WeightMat* mu = k == 0 ? &L.ff1_up : &L.ff2_up;
    WeightMat* md = k == 0 ? &L.ff1_down : &L.ff2_down;
    if ((rc = upload_mat(e, up, 2 * D_FF, D_MODEL, BN_SWIGLU, mu))) return rc;
    if ((rc = upload_f32(e, upb, k == 0 ? &L.ff1_up_b : &L.ff2_up_b))) return rc;
    if ((rc = upload_mat(e, w2->data, D_MODEL, D_FF, BN_PART, md))) return rc;
    if ((rc = upload_f32(e, b2->data, k == 0 ? &L.ff1_down_b : &L.ff2_down_b))) return rc;
  }
  const std::string A = Lp + "self_attn.";
  NEEDW(wv, A + "linear_v.weight");
  NEEDW(bv, A + "linear_v.bias");
  NEEDW(wo, A + "linear_out.weight");
  NEEDW(bo, A + "linear_out.bias");
  if ((rc = upload_mat(e, wo->data, D_MODEL, D_MODEL, BN_RESID, &L.wo))) return rc;
  if ((rc = upload_f32(e, bo->data, &L.wo_b))) return rc;
  if (RECOMPUTE[l]) {
    NEEDW(wq, A + "linear_q.weight");
    NEEDW(bq, A + "linear_q.bias");
    NEEDW(wk, A + "linear_k.weight");
    NEEDW(bk, A + "linear_k.bias");
    if (l < 14) {  // q | k | v on the same T rows
      if ((rc = upload_mat(e, concat({&wq->data, &wk->data, &wv->data}), 3 * D_MODEL, D_MODEL, BN_STORE, &L.qkv)))
        return rc;
      if ((rc = upload_f32(e, concat({&bq->data, &bk->data, &bv->data}), &L.qkv_b))) return rc;
      NEEDW(gat, Lp + "norm_self_att.weight");
      if ((rc = upload_mat(e, fold_cols(concat({&wq->data, &wk->data, &wv->data}), gat->data, 3 * D_MODEL, D_MODEL),
                           3 * D_MODEL, D_MODEL, BN_STORE, &L.qkv_f)))
        return rc;
    } else {       // q on the T new rows, k | v on the S+T [cache | new] rows
      if ((rc = upload_mat(e, wq->data, D_MODEL, D_MODEL, BN_STORE, &L.q))) return rc;
      if ((rc = upload_f32(e, bq->data, &L.q_b))) return rc;
      if ((rc = upload_mat(e, concat({&wk->data, &wv->data}), 2 * D_MODEL, D_MODEL, BN_KV, &L.kv))) return rc;
      if ((rc = upload_f32(e, concat({&bk->data, &bv->data}), &L.kv_b))) return rc;
    }
    NEEDW(qw, A + "q_ln.weight");
    NEEDW(qb, A + "q_ln.bias");
    NEEDW(kw, A + "k_ln.weight");
    NEEDW(kb, A + "k_ln.bias");
    if ((rc = upload_f32(e, qw->data, &L.qln_w))) return rc;
    if ((rc = upload_f32(e, qb->data, &L.qln_b))) return rc;
    if ((rc = upload_f32(e, kw->data, &L.kln_w))) return rc;
    if ((rc = upload_f32(e, kb->data, &L.kln_b))) return rc;
  } else {
    if ((rc = upload_mat(e, wv->data, D_MODEL, D_MODEL, BN_STORE, &L.qkv))) return rc;
    if ((rc = upload_f32(e, bv->data, &L.qkv_b))) return rc;
    if ((rc = make_map_2d(e, &L.wv48, L.qkv.ptr, D_MODEL, D_MODEL, D_HEAD, true))) return rc;
    NEEDW(gat, Lp + "norm_self_att.weight");
    if ((rc = upload_mat(e, fold_cols(wv->data, gat->data, D_MODEL, D_MODEL), D_MODEL, D_MODEL, BN_STORE, &L.qkv_f))) return rc;
    if ((rc = make_map_2d(e, &L.wv48_f, L.qkv_f.ptr, D_MODEL, D_MODEL, D_HEAD, true))) return rc;
  }
  // conv module: pw1 rows [0,384) = a, [384,768) = b (GLU = a * sigmoid(b), conformer_blocks.py:422)
  const std::string Cp = Lp + "conv.";
  NEEDW(p1, Cp + "pointwise_conv1.weight");
  NEEDW(p1b, Cp + "pointwise_conv1.bias");
  {
    std::vector<float> a(p1->data.begin(), p1->data.begin() + (size_t)D_MODEL * D_MODEL);
    std::vector<float> b(p1->data.begin() + (size_t)D_MODEL * D_MODEL, p1->data.end());
    {   // norm_conv gain folded into the input columns (row-scale RMSNorm)
      NEEDW(gcv, Lp + "norm_conv.weight");
      for (int n = 0; n < D_MODEL; ++n)
        for (int c = 0; c < D_MODEL; ++c) {
          a[(size_t)n * D_MODEL + c] *= gcv->data[c];
          b[(size_t)n * D_MODEL + c] *= gcv->data[c];
        }
    }
    std::vector<float> ab(p1b->data.begin(), p1b->data.begin() + D_MODEL);
    std::vector<float> bb(p1b->data.begin() + D_MODEL, p1b->data.end());
    if ((rc = upload_mat(e, interleave_rows(a, b, D_MODEL, D_MODEL, BN_GLU / 2), 2 * D_MODEL, D_MODEL, BN_GLU, &L.pw1)))
      return rc;
    if ((rc = upload_f32(e, interleave_rows(ab, bb, D_MODEL, 1, BN_GLU / 2), &L.pw1_b))) return rc;
    if ((rc = upload_mat(e, interleave_rows(a, b, D_MODEL, D_MODEL, 64), 2 * D_MODEL, D_MODEL, 128, &L.pw1w))) return rc;
    if ((rc = upload_f32(e, interleave_rows(ab, bb, D_MODEL, 1, 64), &L.pw1w_b))) return rc;
  }
  NEEDW(dw, Cp + "depthwise_conv.conv.weight");
  NEEDW(dwb, Cp + "depthwise_conv.conv.bias");
  {
    std::vector<float> al, be;
    if ((rc = bn_fold(e, Cp + "batch_norm.", dwb->data, D_MODEL, al, be))) return rc;
    std::vector<float> t((size_t)31 * D_MODEL);
    for (int c = 0; c < D_MODEL; ++c)
      for (int j = 0; j < 31; ++j) t[(size_t)j * D_MODEL + c] = dw->data[(size_t)c * 31 + j] * al[c];
    if ((rc = upload_f32(e, t, &L.dw_w))) return rc;
    if ((rc = upload_f32(e, be, &L.dw_b))) return rc;
  }
  NEEDW(p2, Cp + "pointwise_conv2.weight");
  NEEDW(p2b, Cp + "pointwise_conv2.bias");
  if ((rc = upload_mat(e, p2->data, D_MODEL, D_MODEL, BN_RESID, &L.pw2))) return rc;
  if ((rc = upload_f32(e, p2b->data, &L.pw2_b))) return rc;
  return 0;
}

extern "C" int tone_finalize_weights(tone_engine* e) {
  if (!e) return fail(TONE_EINVAL, "null engine");
  if (e->finalized) return fail(TONE_ESTATE, "weights already finalized");
  CK(cudaSetDevice(e->cfg.device));
  int rc;
  if ((rc = finalize_frontend(e))) return rc;
  if ((rc = finalize_pre_encode(e))) return rc;
  for (int l = 0; l < N_LAYERS; ++l)
    if ((rc = finalize_layer(e, l))) return rc;
  const std::string R = "encoder.temportal_reduction.";
  NEEDW(rw, R + "conv.weight");
  NEEDW(rb, R + "conv.bias");
  NEEDW(pw, R + "conv_pw.weight");
  NEEDW(pb, R + "conv_pw.bias");
  if ((rc = expect_shape(rw, "reduction conv.weight", {4 * D_MODEL, 1, 3}))) return rc;
  if ((rc = upload_f32(e, rw->data, &e->red_dw_w))) return rc;
  if ((rc = upload_f32(e, rb->data, &e->red_dw_b))) return rc;
  if ((rc = upload_mat(e, pw->data, D_MODEL, D_FF, BN_STORE, &e->red_pw))) return rc;
  if ((rc = upload_f32(e, pb->data, &e->red_pw_b))) return rc;
  NEEDW(dw, "decoder.decoder_layers.0.weight");
  NEEDW(db, "decoder.decoder_layers.0.bias");
  if ((rc = expect_shape(dw, "decoder weight", {N_CLASSES, D_MODEL, 1}))) return rc;
  {
    std::vector<float> t((size_t)DEC_PAD * D_MODEL, 0.f);
    memcpy(t.data(), dw->data.data(), (size_t)N_CLASSES * D_MODEL * 4);
    if ((rc = upload_mat(e, t, DEC_PAD, D_MODEL, DEC_PAD, &e->dec_w))) return rc;
    if ((rc = upload_f32(e, db->data, &e->dec_b))) return rc;
  }
  e->host_w.clear();
  CK(cudaDeviceSynchronize());   // pageable H2D copies may still be in flight when cudaMemcpy returns
  e->finalized = true;
  return TONE_OK;
}

// ------------------------------------------------------------------------------------------------ slots
static StatePool state_pool(tone_engine* e) {
  StatePool p;
  p.pre = e->st_pre;
  p.feat = e->st_feat;
  p.x1 = e->st_x1;
  p.kv14 = e->st_kv14;
  p.kv15 = e->st_kv15;
  p.conv = e->st_conv;
  p.red = e->st_red;
  p.len = e->st_len;
  p.cpos = e->st_cpos;
  p.ph = e->st_ph;
  p.F = e->F;
  p.T = e->T;
  p.T2 = e->T2;
  return p;
}

// Range / allocation / uniqueness of the slot ids of one batch (O(B): a per-engine stamp array).
static int validate_slots(tone_engine* e, int n, const int32_t* slots, bool need_alloc = true) {
  if (++e->stamp_gen == 0) {                       // wrapped: restart the stamps
    std::fill(e->slot_stamp.begin(), e->slot_stamp.end(), 0u);
    e->stamp_gen = 1;
  }
  for (int i = 0; i < n; ++i) {
    const int sl = slots[i];
    if (sl < 0 || sl >= e->cfg.max_slots) return fail(TONE_EINVAL, "slot %d out of range [0, %d)", sl, e->cfg.max_slots);
    if (need_alloc && !e->slot_used[sl]) return fail(TONE_ESTATE, "slot %d is not allocated", sl);
    if (e->slot_stamp[sl] == e->stamp_gen) return fail(TONE_EINVAL, "slot %d appears twice in one batch", sl);
    e->slot_stamp[sl] = e->stamp_gen;
  }
  return 0;
}

// Slot ids for an asynchronous kernel: through a small pinned ring (an entry is reused only after the copy that read
// it has completed), H2D on `st`.  Returns the device pointer through *d.
static int ring_slots(tone_engine* e, int n, const int32_t* slots, cudaStream_t st, int** d) {
  const size_t cap = std::max<size_t>(e->cfg.max_batch, STATE_IO_CHUNK);
  if ((size_t)n > cap) return fail(TONE_EINVAL, "%d slots exceed the staging capacity %zu", n, cap);
  const int k = e->ring_pos;
  e->ring_pos = (k + 1) % tone_engine::RING;
  CK(cudaEventSynchronize(e->ev_ring[k]));
  memcpy(e->p_ring + k * cap, slots, (size_t)n * 4);
  CK(cudaMemcpyAsync(e->d_ring + k * cap, e->p_ring + k * cap, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(e->ev_ring[k], st));
  *d = e->d_ring + k * cap;
  return 0;
}

extern "C" int tone_reset_slots(tone_engine* e, int32_t n, const int32_t* slots) {
  if (!e || !slots || n < 0) return fail(TONE_EINVAL, "bad argument");
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, n, slots, false));
  const int cap = (int)std::max<size_t>(e->cfg.max_batch, STATE_IO_CHUNK);
  for (int i0 = 0; i0 < n; i0 += cap) {
    const int m = std::min(cap, n - i0);
    int* d = nullptr;
    RC(ring_slots(e, m, slots + i0, e->stream, &d));
    reset_slots_kernel<<<m, STATE_IO_THREADS, 0, e->stream>>>(state_pool(e), d);
    CK(cudaGetLastError());
  }
  return TONE_OK;   // stream-ordered before any later step of this engine
}

extern "C" int tone_alloc_slots(tone_engine* e, int32_t n, int32_t* out) {
  if (!e || !out || n < 0) return fail(TONE_EINVAL, "bad argument");
  if ((size_t)n > e->free_slots.size())
    return fail(TONE_ENOMEM, "%d slots requested, %zu free of %d", n, e->free_slots.size(), e->cfg.max_slots);
  for (int i = 0; i < n; ++i) {
    out[i] = e->free_slots.back();
    e->free_slots.pop_back();
    e->slot_used[out[i]] = 1;
  }
  return tone_reset_slots(e, n, out);
}

extern "C" int tone_release_slots(tone_engine* e, int32_t n, const int32_t* slots) {
  if (!e || !slots || n < 0) return fail(TONE_EINVAL, "bad argument");
  RC(validate_slots(e, n, slots));
  for (int i = 0; i < n; ++i) {
    e->slot_used[slots[i]] = 0;
    e->free_slots.push_back(slots[i]);
  }
  return TONE_OK;
}

// ------------------------------------------------------------------------------------------------ the step
template <int KIND, int BN>
static int gemm(tone_engine* e, cudaStream_t st, const CUtensorMap& tmA, const WeightMat& w, GemmArgs a, int m_tiles,
                int n_tiles, int ref_rows, int ref_cols, const CUtensorMap* tmAw = nullptr, int splits = 1,
                int box = 0) {   // weight map: 0 = w.map (box rows = the matrix's upload width), 1 = map128, 2 = map64
  a.W = w.ptr;
  a.ldw = w.K;
  cudaError_t err;
  constexpr bool can_persist = BN == 128 && (KIND == G_SWIGLU || KIND == G_GLU || KIND == G_RESID || KIND == G_STORE_F32 ||
                                             KIND == G_PARTIAL);
  bool persist = false;
  if constexpr (can_persist)
    persist = e->cfg.gemm_impl == 0 && splits == 1 && e->persist_min_tiles > 0 &&
              m_tiles * n_tiles >= std::min(e->persist_min_tiles, e->cur_ctas + 1);
  if (persist) {
    if constexpr (can_persist) {
      const CUtensorMap& mb = box == 1 ? w.map128 : w.map;
      constexpr bool wide = (KIND == G_SWIGLU || KIND == G_GLU);   // N a multiple of 256: two weight tiles per tile
      if (wide && e->persist_mode == 2)
        err = launch_gemm_tc_persist<KIND, wide ? 2 : 1, wide>(st, tmA, mb, a, m_tiles, n_tiles, e->pdl, e->cur_ctas);
      else if (wide && e->persist_mode == 1)
        err = launch_gemm_tc_persist<KIND, wide ? 2 : 1, false>(st, tmA, mb, a, m_tiles, n_tiles, e->pdl, e->cur_ctas);
      else err = launch_gemm_tc_persist<KIND, 1, false>(st, tmA, mb, a, m_tiles, n_tiles, e->pdl, e->cur_ctas);
    }
  } else if (e->cfg.gemm_impl == 0)
    err = launch_gemm_tc<KIND, BN>(st, tmA, tmAw ? *tmAw : tmA, box == 1 ? w.map128 : (box == 2 ? w.map64 : w.map), a, m_tiles, n_tiles, e->pdl,
                                   e->num_sms, splits);
  else err = launch_gemm_ref<KIND, BN>(st, a, ref_rows, ref_cols, splits);
  e->launches++;
  if (err != cudaSuccess) return fail(TONE_ECUDA, "gemm kind %d launch: %s", KIND, cudaGetErrorString(err));
  return 0;
}

static GemmArgs dense_args(int M, int K, const bf16* A, void* out, int ldo, const float* bias, float scale) {
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.nk = K / 64;
  a.R = 128;
  a.G = 1;
  a.out = out;
  a.ldo = ldo;
  a.bias = bias;
  a.scale = scale;
  a.A = A;
  a.lda = K;
  return a;
}

#define KLAUNCH(call)                                                                                   \
  do {                                                                                                  \
    e->launches++;                                                                                      \
    cudaError_t _e = (call);                                                                            \
    if (_e != cudaSuccess) return fail(TONE_ECUDA, "launch at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

struct PartIn {            // split-K output waiting to be folded into the residual stream by the next norm kernel
  const __half* part = nullptr;
  int nsplit = 0;
  long long stride = 0;
  const float* bias = nullptr;
  float scale = 0.f;
};

static int run_norm(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, float* r, const float* g1, const float* g2, bf16* n, int M,
                    const PartIn& p = PartIn(), bf16* kv = nullptr, int rows_per_stream = 1, int kv_row_off = 0) {
  NormArgs a{r, g1, g2, n, M, p.part, p.nsplit, p.stride, p.bias, p.scale, kv, ln.slots, rows_per_stream, kv_row_off};
  if (p.nsplit > 2) KLAUNCH(launch_kernel(norm_kernel<MAX_SPLITS>, dim3((M + 7) / 8), dim3(256), 0, st, e->pdl, a));
  else KLAUNCH(launch_kernel(norm_kernel<2>, dim3((M + 7) / 8), dim3(256), 0, st, e->pdl, a));
  return 0;
}

// Feed-forward: h = silu(n W1^T + b1) * (n Wv^T + bv); the down projection runs split-K and leaves its partial
// sums in ln.part; the NEXT norm kernel adds 0.5 * (sum + b2) to the residual stream (conformer_blocks.py:814,834).
static int run_resid_rowscale(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, int M, const bf16* A,
                              const CUtensorMap& mapA, const WeightMat& w, const float* bias, float* r, int* ss_tiles,
                              int K, float scale);
// resid_r != nullptr (large batches, no split-K): the down projection adds 0.5 * (acc + b2) to the residual stream in
// its own epilogue and emits bf16(r) + the row sums of squares (*resid_ss_tiles tiles), so that the NEXT GEMM applies
// the following RMSNorm as a row scale - no partial sums through HBM and no norm kernel.
static int run_ff(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, int M, const WeightMat& up, const float* up_b, const WeightMat& down,
                  const float* down_b, PartIn* out, int ss_tiles = 0, float* resid_r = nullptr, int* resid_ss_tiles = nullptr) {
  const int mt = (M + 127) / 128;
  // ss_tiles > 0: the input is the un-normalised residual in bf16 (ln.rb) with its row sums of squares in ln.ss
  GemmArgs a = dense_args(M, D_MODEL, ss_tiles ? ln.rb : ln.n, ln.h, D_FF, up_b, 1.f);
  if (ss_tiles) {
    a.ss = ln.ss;
    a.ss_ld = 12;
    a.ss_tiles = ss_tiles;
  }
  RC((gemm<G_SWIGLU, BN_SWIGLU>(e, st, ss_tiles ? ln.m_rb : ln.m_n, up, a, mt, 2 * D_FF / BN_SWIGLU, M, D_FF)));
  if (resid_r) {
    *out = PartIn();
    return run_resid_rowscale(e, ln, st, M, ln.h, ln.m_h, down, down_b, resid_r, resid_ss_tiles, D_FF, 0.5f);
  }
  int splits = 1;
  while (splits < e->max_splits && mt * (D_MODEL / BN_PART) * splits * 2 <= e->num_sms) splits *= 2;  // fill the SMs once
  if (e->split_k) splits = e->split_k;
  splits = std::max(1, std::min(splits, std::min(e->max_splits, (int)MAX_SPLITS)));
  GemmArgs b = dense_args(M, D_FF / splits, ln.h, ln.part, D_MODEL, nullptr, 1.f);
  b.lda = D_FF;
  b.z_stride = (long long)e->rows_alloc * D_MODEL;
  RC((gemm<G_PARTIAL, BN_PART>(e, st, ln.m_h, down, b, mt, D_MODEL / BN_PART, M, D_MODEL, nullptr, splits)));
  out->part = ln.part;
  out->nsplit = splits;
  out->stride = b.z_stride;
  out->bias = down_b;
  out->scale = 0.5f;
  return 0;
}

// The whole feed-forward module in one kernel per row tile (ff_fused.cuh): r += 0.5 FF(a); [r = norm(r; g1)];
// n = norm(r; g2) (bf16, optionally scattered into the [cache | new] rows of a stateful attention layer).
static bool use_ff_fused(const tone_engine* e, int M) {
  return e->cfg.gemm_impl == 0 && e->ff_fused > 0 && M >= e->ff_fused_min_rows;
}
static int run_ff_fused(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, int M, float* r, const WeightMat& up,
                        const float* up_b, const WeightMat& down, const float* down_b, int ss_tiles, const float* g1,
                        const float* g2, bf16* n_out, bf16* kv = nullptr, int rows_per_stream = 1, int kv_row_off = 0) {
  FfArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.up_bias = up_b;
  a.down_bias = down_b;
  if (ss_tiles) {
    a.ss = ln.ss;
    a.ss_ld = 12;
    a.ss_tiles = ss_tiles;
  }
  a.r = r;
  a.scale = 0.5f;
  a.g1 = g1;
  a.g2 = g2;
  a.n = n_out;
  a.kv = kv;
  a.slots = ln.slots;
  a.rows_per_stream = rows_per_stream;
  a.kv_row_off = kv_row_off;
  const CUtensorMap& mA = ss_tiles ? ln.m_rb : ln.m_n;
  const int mt = (M + 127) / 128;
  cudaError_t err = e->ff_fused == 2 ? launch_ff_fused<true>(st, mA, up.map64, down.map, down.map64, a, mt, e->pdl)
                                     : launch_ff_fused<false>(st, mA, up.map, down.map, down.map, a, mt, e->pdl);
  e->launches++;
  if (err != cudaSuccess) return fail(TONE_ECUDA, "fused feed-forward launch: %s", cudaGetErrorString(err));
  return 0;
}

// r += A W^T + b through the tensor cores; also emits bf16(r) and the per-tile row sums of squares that let the next
// GEMM apply the following RMSNorm as a row scale.  Returns the number of ss tiles through *ss_tiles.
static int run_resid_rowscale(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, int M, const bf16* A,
                              const CUtensorMap& mapA, const WeightMat& w, const float* bias, float* r, int* ss_tiles,
                              int K, float scale) {
  const int mt = (M + 127) / 128;
  GemmArgs a = dense_args(M, K, A, r, D_MODEL, bias, scale);
  if (e->cfg.gemm_impl == 0) {
    a.rb_out = ln.rb;
    a.ss_out = ln.ss;
    a.ss_ld = 12;
    if (M >= BIG_M && e->cur_lanes == 1 && 2 * mt * (D_MODEL / 128) <= e->num_sms) {
      // 128-wide tiles would fill at most half of the SMs (2560 rows: 60 CTAs): 64-wide tiles, half the epilogue per CTA
      // (with two lanes the 60-CTA kernels of both lanes run side by side instead: measured equal or better)
      *ss_tiles = D_MODEL / 64;
      RC((gemm<G_RESID, 64>(e, st, mapA, w, a, mt, D_MODEL / 64, M, D_MODEL, nullptr, 1, 2)));
    } else if (M >= BIG_M) {
      *ss_tiles = D_MODEL / 128;
      RC((gemm<G_RESID, 128>(e, st, mapA, w, a, mt, D_MODEL / 128, M, D_MODEL, nullptr, 1, true)));
    } else {
      *ss_tiles = D_MODEL / BN_RESID;
      RC((gemm<G_RESID, BN_RESID>(e, st, mapA, w, a, mt, D_MODEL / BN_RESID, M, D_MODEL)));
    }
  } else {   // SIMT debug path: plain residual GEMM, then one helper kernel for rb / ss
    RC((gemm<G_RESID, BN_RESID>(e, st, mapA, w, a, mt, D_MODEL / BN_RESID, M, D_MODEL)));
    rowscale_ref_kernel<<<(M + 127) / 128, 128, 0, st>>>(r, ln.rb, ln.ss, 12, M);
    e->launches++;
    *ss_tiles = 1;
  }
  return 0;
}

// Row-owner CTA-pair GEMM (rowgemm.cuh): x = r + scale * (A W^T + b); r = x or norm(x; g1); then rb / ss (row-scale
// consumer follows) or n = norm(r; g2) (bf16, optional cache-row scatter).
static bool use_rowgemm(const tone_engine* e, int M) {
  return e->cfg.gemm_impl == 0 && e->rowgemm_min_rows > 0 && M >= e->rowgemm_min_rows;
}
static int run_rowgemm(tone_engine* e, tone_engine::Lane& ln, cudaStream_t st, int M, int K, const CUtensorMap& mapA,
                       const WeightMat& w, const float* bias, float scale, float* r, bool emit_rb, const float* g1 = nullptr,
                       const float* g2 = nullptr, bf16* n = nullptr, bf16* kv = nullptr, int rows_per_stream = 1,
                       int kv_row_off = 0) {
  RowGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.nk = K / 64;
  a.bias = bias;
  a.scale = scale;
  a.r = r;
  a.g1 = g1;
  a.g2 = g2;
  a.n = n;
  a.kv = kv;
  a.slots = ln.slots;
  a.rows_per_stream = rows_per_stream;
  a.kv_row_off = kv_row_off;
  if (emit_rb) {
    a.rb_out = ln.rb;
    a.ss_out = ln.ss;
    a.ss_ld = 12;
  }
  cudaError_t err = launch_rowgemm(st, mapA, w.map128, w.map64, a, (M + 127) / 128, e->pdl);
  e->launches++;
  if (err != cudaSuccess) return fail(TONE_ECUDA, "row-owner GEMM launch: %s", cudaGetErrorString(err));
  return 0;
}

// taps: optional host pointer [17][B*T][384]; when set the step synchronises after every layer (debug only)
static int run_step(tone_engine* e, tone_engine::Lane& ln, int B, cudaStream_t st, float* taps) {
  const int T = e->T, T2 = e->T2, F = e->F, C = e->C;
  auto tap = [&](int idx, const float* src, int rows) -> int {
    if (!taps) return 0;
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(taps + (size_t)idx * B * T * D_MODEL, src, (size_t)rows * D_MODEL * 4, cudaMemcpyDeviceToHost));
    return 0;
  };
  {
    BeginArgs a;
    memset(&a, 0, sizeof(a));
    a.slots = ln.slots;
    a.pcm = ln.pcm;
    a.pcm_fmt = ln.pcm_fmt;
    a.feats_in = ln.feats;
    a.pre = e->st_pre;
    a.feat = e->st_feat;
    a.x1 = e->st_x1;
    a.kv14 = e->st_kv14;
    a.kv15 = e->st_kv15;
    a.mhsa_len = e->st_len;
    a.len_in = ln.len_in;
    a.cpos = e->st_cpos;
    a.cpos_in = ln.cpos_in;
    a.basis = e->basis;
    a.mel_start = e->mel_start;
    a.mel_bin = e->mel_bin;
    a.mel_w = e->mel_w;
    a.pre_norm_g = e->pre_norm_g;
    a.C = C;
    a.F = F;
    a.T = T;
    a.T2 = T2;
    a.B = B;
    const int n_mt = (F + 15) / 16, UH = ((16 * n_mt * HOP + HOP + 16) + 7) & ~7;
    const size_t smem = (size_t)2 * BASIS_N * BASIS_LD * 2 + ROLL_BYTES + (size_t)UH * 2 + (size_t)(F * 162 + F * N_MELS) * 4 +
                        (size_t)C * 4 + HOP * 2;   // + staging of the next stream's PCM (int32 at most) and carried samples
    KLAUNCH(launch_kernel(begin_step_kernel, dim3(std::min(B, e->cur_ctas)), dim3(BEGIN_THREADS), smem, st, e->pdl, a));
  }
  {  // conv0: rows = F frames per stream, K = 11 kernel rows x 64 mel bins, N = 44 positions x 32 channels
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.M = B;
    a.nk = 11;
    a.R = F;
    a.G = 128 / F;
    a.slots = ln.slots;
    a.out = e->st_x1;
    a.ldo = X1_ROW;
    a.alpha = e->conv0_alpha;
    a.beta = e->conv0_beta;
    a.out_slot_stride = (long long)X1_ROWS_MAX * X1_ROW;
    a.out_row_off = SUB2_ROWS;
    a.A = e->st_feat;
    a.a_slot_stride = FEAT_ROWS_MAX * N_MELS;
    RC((gemm<G_CONV0, BN_CONV>(e, st, e->m_feat, e->conv0_w, a, (B + a.G - 1) / a.G, X1_ROW / BN_CONV, B * F, X1_ROW,
                               &e->w_feat)));
  }
  {  // conv1: rows = T frames per stream, K = 11 kernel rows x (12 positions x 32 channels), N = 2 positions x 64 ch
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.M = B;
    a.nk = 66;
    a.R = T;
    a.G = 128 / T;
    a.slots = ln.slots;
    a.out = ln.c1;
    a.ldo = SUB_OUT;
    a.alpha = e->conv1_alpha;
    a.beta = e->conv1_beta;
    a.A = e->st_x1;
    a.a_slot_stride = (long long)X1_ROWS_MAX * X1_ROW;
    RC((gemm<G_CONV1, BN_CONV>(e, st, e->m_x1, e->conv1_w, a, (B + a.G - 1) / a.G, SUB_OUT / BN_CONV, B * T, SUB_OUT,
                               &e->w_x1)));
  }
  int M = B * T;
  {
    GemmArgs a = dense_args(M, SUB_OUT, ln.c1, ln.r_full, D_MODEL, nullptr, 1.f);
    if (M >= BIG_M) RC((gemm<G_STORE_F32, 128>(e, st, ln.m_c1, e->out_w, a, (M + 127) / 128, D_MODEL / 128, M, D_MODEL, nullptr, 1, true)));
    else RC((gemm<G_STORE_F32, BN_STORE>(e, st, ln.m_c1, e->out_w, a, (M + 127) / 128, D_MODEL / BN_STORE, M, D_MODEL)));
  }
  RC(run_norm(e, ln, st, ln.r_full, e->out_norm_g, e->L[0].n_ff1, ln.n, M));
  RC(tap(0, ln.r_full, M));

  for (int l = 0; l < N_LAYERS; ++l) {
    LayerW& L = e->L[l];
    const bool reduced = l > 6 && l <= 14;
    const int Tl = reduced ? T2 : T;
    float* r = reduced ? ln.r_red : ln.r_full;
    M = B * Tl;
    const int mt = (M + 127) / 128;
    PartIn ff;
    const bool fused_ff = use_ff_fused(e, M);
    // large batches, layers 0..13: feed-forward 1 goes straight into the residual stream and norm_self_att becomes a row
    // scale inside the q / k / v projections (layers 14 / 15 need the normalised rows themselves for their caches)
    const bool lazy_att = !fused_ff && e->lazy_norm_min_rows > 0 && e->cfg.gemm_impl == 0 && M >= std::max(e->lazy_norm_min_rows, BIG_M) &&
                          l < 14 && (RECOMPUTE[l] || e->fuse_vatt);
    int att_ss_tiles = 0;
    const bool rowg = !fused_ff && use_rowgemm(e, M);
    bool ff2_norm_done = false;       // norm_out (+ the next layer's first norm) done by feed-forward 2's down projection
    bool ff1_norm_done = false;       // the attention input (n, or rb + ss) has already been produced by the down projection
    if (fused_ff) {   // feed-forward 1 + residual + norm_self_att (+ cache-row scatter of layers 14 / 15) in one kernel
      if (l < 14) RC(run_ff_fused(e, ln, st, M, r, L.ff1_up, L.ff1_up_b, L.ff1_down, L.ff1_down_b, 0, nullptr, L.n_att, ln.n));
      else
        RC(run_ff_fused(e, ln, st, M, r, L.ff1_up, L.ff1_up_b, L.ff1_down, L.ff1_down_b, 0, nullptr, L.n_att, ln.n,
                        l == 14 ? e->st_kv14 : e->st_kv15, Tl, l == 14 ? MHSA_S / 2 : MHSA_S));
    } else if (rowg) {
      // up GEMM as usual; the down projection owns whole rows: residual add in its epilogue, then either bf16(r) + sum of
      // squares for the row-scale projections (layers 0..13) or the normalised rows themselves (layers 14 / 15)
      GemmArgs ua = dense_args(M, D_MODEL, ln.n, ln.h, D_FF, L.ff1_up_b, 1.f);
      RC((gemm<G_SWIGLU, BN_SWIGLU>(e, st, ln.m_n, L.ff1_up, ua, mt, 2 * D_FF / BN_SWIGLU, M, D_FF)));
      if (lazy_att) {
        RC(run_rowgemm(e, ln, st, M, D_FF, ln.m_h, L.ff1_down, L.ff1_down_b, 0.5f, r, true));
        att_ss_tiles = 1;
      } else if (l < 14) {
        RC(run_rowgemm(e, ln, st, M, D_FF, ln.m_h, L.ff1_down, L.ff1_down_b, 0.5f, r, false, nullptr, L.n_att, ln.n));
      } else {
        RC(run_rowgemm(e, ln, st, M, D_FF, ln.m_h, L.ff1_down, L.ff1_down_b, 0.5f, r, false, nullptr, L.n_att, ln.n,
                       l == 14 ? e->st_kv14 : e->st_kv15, Tl, l == 14 ? MHSA_S / 2 : MHSA_S));
      }
      ff1_norm_done = true;
    } else if (lazy_att) {
      RC(run_ff(e, ln, st, M, L.ff1_up, L.ff1_up_b, L.ff1_down, L.ff1_down_b, &ff, 0, r, &att_ss_tiles));
    } else {
      RC(run_ff(e, ln, st, M, L.ff1_up, L.ff1_up_b, L.ff1_down, L.ff1_down_b, &ff));
    }
    // ---- attention (its norm kernel first folds the feed-forward output into r)
    AttnArgs at;
    memset(&at, 0, sizeof(at));
    at.P = ln.P;
    at.ctx = ln.ctx;
    at.rope_cos = e->rope_cos;
    at.rope_sin = e->rope_sin;
    at.len_in = ln.len_in;
    at.T = Tl;
    at.recompute = RECOMPUTE[l] ? 1 : 0;
    bool fused_att = false, att_block = false;
    if (l < 14) {
      if (!fused_ff && !lazy_att && !ff1_norm_done) RC(run_norm(e, ln, st, r, nullptr, L.n_att, ln.n, M, ff));
      at.S = 0;
      at.Tk = Tl;
      if (RECOMPUTE[l]) {
        GemmArgs a = dense_args(M, D_MODEL, lazy_att ? ln.rb : ln.n, ln.qkv, 3 * D_MODEL, L.qkv_b, 1.f);
        if (lazy_att) {
          a.ss = ln.ss;
          a.ss_ld = 12;
          a.ss_tiles = att_ss_tiles;
          RC((gemm<G_STORE_F32, 128>(e, st, ln.m_rb, L.qkv_f, a, mt, 3 * D_MODEL / 128, M, 3 * D_MODEL, nullptr, 1, true)));
        } else if (M >= BIG_M) RC((gemm<G_STORE_F32, 128>(e, st, ln.m_n, L.qkv, a, mt, 3 * D_MODEL / 128, M, 3 * D_MODEL, nullptr, 1, true)));
        else RC((gemm<G_STORE_F32, BN_STORE>(e, st, ln.m_n, L.qkv, a, mt, 3 * D_MODEL / BN_STORE, M, 3 * D_MODEL)));
        at.q = ln.qkv;
        at.k = ln.qkv + D_MODEL;
        at.v = ln.qkv + 2 * D_MODEL;
        at.ldq = at.ldk = at.ldv = 3 * D_MODEL;
        at.q_ln_w = L.qln_w;
        at.q_ln_b = L.qln_b;
        at.k_ln_w = L.kln_w;
        at.k_ln_b = L.kln_b;
      } else if (e->cfg.gemm_impl == 0 && e->fuse_vatt && e->att_block_min_rows > 0 && M >= e->att_block_min_rows) {
        // score-sharing layer, large batch: V projection + P.V + out projection + residual in ONE kernel per tile of
        // whole streams
        AttFArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.B = B;
        fa.R = Tl;
        fa.G = 128 / Tl;
        fa.P = ln.P;
        fa.bv = L.qkv_b;
        fa.bo = L.wo_b;
        if (lazy_att) {
          fa.ss = ln.ss;
          fa.ss_tiles = att_ss_tiles;
        }
        fa.ss_ld = 12;
        fa.r = r;
        fa.rb_out = ln.rb;
        fa.ss_out = ln.ss;
        cudaError_t err = launch_att_fused(st, lazy_att ? ln.m_rb : ln.m_n, lazy_att ? L.qkv_f.map128 : L.qkv.map128,
                                           L.wo.map128, fa, e->pdl);
        e->launches++;
        if (err != cudaSuccess) return fail(TONE_ECUDA, "fused attention block launch: %s", cudaGetErrorString(err));
        fused_att = true;
        att_block = true;
      } else if (e->cfg.gemm_impl == 0 && e->fuse_vatt) {
        // score-sharing layer: ctx = P (n Wv^T + bv) per head in ONE kernel; tiles = whole streams x one head
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.M = B;
        a.nk = D_MODEL / 64;
        a.R = Tl;
        a.G = 128 / Tl;
        a.out = ln.ctx;
        a.ldo = D_MODEL;
        a.bias = L.qkv_b;
        a.P = ln.P;
        a.A = lazy_att ? ln.rb : ln.n;
        a.lda = D_MODEL;
        if (lazy_att) {
          a.ss = ln.ss;
          a.ss_ld = 12;
          a.ss_tiles = att_ss_tiles;
        }
        const CUtensorMap& mA = lazy_att ? ln.m_rb : ln.m_n;
        cudaError_t err = launch_gemm_tc<G_VATT, D_HEAD>(st, mA, mA, lazy_att ? L.wv48_f : L.wv48, a, (B + a.G - 1) / a.G,
                                                         N_HEADS, e->pdl, e->num_sms);
        e->launches++;
        if (err != cudaSuccess) return fail(TONE_ECUDA, "fused V + attention launch: %s", cudaGetErrorString(err));
        fused_att = true;
      } else {
        GemmArgs a = dense_args(M, D_MODEL, ln.n, ln.qkv, D_MODEL, L.qkv_b, 1.f);
        if (M >= BIG_M) RC((gemm<G_STORE_F32, 128>(e, st, ln.m_n, L.qkv, a, mt, D_MODEL / 128, M, D_MODEL, nullptr, 1, true)));
        else RC((gemm<G_STORE_F32, BN_STORE>(e, st, ln.m_n, L.qkv, a, mt, D_MODEL / BN_STORE, M, D_MODEL)));
        at.v = ln.qkv;
        at.ldv = D_MODEL;
      }
    } else {
      const int S = (l == 14) ? MHSA_S / 2 : MHSA_S;
      bf16* kvbuf = (l == 14) ? e->st_kv14 : e->st_kv15;
      if (!fused_ff && !ff1_norm_done) RC(run_norm(e, ln, st, r, nullptr, L.n_att, ln.n, M, ff, kvbuf, Tl, S));
      float* qbuf = ln.qkv;
      float* kvout = ln.qkv + (size_t)e->rows_alloc * D_MODEL;
      GemmArgs a = dense_args(M, D_MODEL, ln.n, qbuf, D_MODEL, L.q_b, 1.f);
      if (M >= BIG_M) RC((gemm<G_STORE_F32, 128>(e, st, ln.m_n, L.q, a, mt, D_MODEL / 128, M, D_MODEL, nullptr, 1, true)));
      else RC((gemm<G_STORE_F32, BN_STORE>(e, st, ln.m_n, L.q, a, mt, D_MODEL / BN_STORE, M, D_MODEL)));
      GemmArgs k;
      memset(&k, 0, sizeof(k));
      k.M = B;
      k.nk = D_MODEL / 64;
      k.R = S + Tl;
      k.G = 128 / k.R;
      k.slots = ln.slots;
      k.out = kvout;
      k.ldo = 2 * D_MODEL;
      k.bias = L.kv_b;
      k.A = kvbuf;
      k.lda = D_MODEL;
      k.a_slot_stride = KV_ROWS_MAX * D_MODEL;
      if (B * k.R >= BIG_M)   // large batch: 128-wide tiles halve the re-reads of the gathered A rows
        RC((gemm<G_KV, 128>(e, st, l == 14 ? e->m_kv14 : e->m_kv15, L.kv, k, (B + k.G - 1) / k.G, 2 * D_MODEL / 128,
                            B * k.R, 2 * D_MODEL, l == 14 ? &e->w_kv14 : &e->w_kv15, 1, true)));
      else
        RC((gemm<G_KV, BN_KV>(e, st, l == 14 ? e->m_kv14 : e->m_kv15, L.kv, k, (B + k.G - 1) / k.G, 2 * D_MODEL / BN_KV,
                              B * k.R, 2 * D_MODEL, l == 14 ? &e->w_kv14 : &e->w_kv15)));
      at.S = S;
      at.Tk = S + Tl;
      at.q = qbuf;
      at.ldq = D_MODEL;
      at.k = kvout;
      at.v = kvout + D_MODEL;
      at.ldk = at.ldv = 2 * D_MODEL;
      at.q_ln_w = L.qln_w;
      at.q_ln_b = L.qln_b;
      at.k_ln_w = L.kln_w;
      at.k_ln_b = L.kln_b;
      at.mask_mode = (l == 14) ? 2 : 1;
    }
    if (fused_att) {
    } else if (RECOMPUTE[l] && e->att_pipe_min_batch > 0 && B >= e->att_pipe_min_batch) {
      const size_t smem = (size_t)ATP_FIXED_SMEM + (size_t)(2 * at.Tk + at.T) * (D_MODEL * 4 + ATP_PAD);
      if (at.S > 0)
        KLAUNCH(launch_kernel(attention_pipe_kernel<512>, dim3(std::min(B, e->cur_ctas)), dim3(512), smem, st, e->pdl, at, B));
      else
        KLAUNCH(launch_kernel(attention_pipe_kernel<256>, dim3(std::min(B, 2 * e->cur_ctas)), dim3(256), smem, st, e->pdl, at, B));
    } else if (RECOMPUTE[l]) KLAUNCH(launch_kernel(attention_kernel<true>, dim3(B, N_HEADS / ATT_HEADS_REC), dim3(ATT_THREADS_REC), ATT_V_SMEM, st, e->pdl, at));
    else KLAUNCH(launch_kernel(attention_kernel<false>, dim3(B), dim3(ATT_THREADS), 0, st, e->pdl, at));
    int ss_tiles = 0;
    if (att_block) ss_tiles = 1;      // the fused kernel owns whole rows: one sum of squares per row
    else if (rowg) {
      RC(run_rowgemm(e, ln, st, M, D_MODEL, ln.m_ctx, L.wo, L.wo_b, 1.f, r, true));
      ss_tiles = 1;
    } else RC(run_resid_rowscale(e, ln, st, M, ln.ctx, ln.m_ctx, L.wo, L.wo_b, r, &ss_tiles, D_MODEL, 1.f));
    // ---- convolution module: norm_conv is applied as a row scale inside the pointwise-conv GEMM (A = bf16(r))
    {
      {
        GemmArgs a = dense_args(M, D_MODEL, ln.rb, ln.g, D_MODEL, M >= BIG_M ? L.pw1w_b : L.pw1_b, 1.f);
        a.ss = ln.ss;
        a.ss_ld = 12;
        a.ss_tiles = ss_tiles;
        if (M >= BIG_M) RC((gemm<G_GLU, 128>(e, st, ln.m_rb, L.pw1w, a, mt, 2 * D_MODEL / 128, M, D_MODEL)));
        else RC((gemm<G_GLU, BN_GLU>(e, st, ln.m_rb, L.pw1, a, mt, 2 * D_MODEL / BN_GLU, M, D_MODEL)));
      }
      DwArgs d;
      d.g = ln.g;
      d.cache = e->st_conv + (size_t)l * CONV_S * D_MODEL;
      d.cache_slot_stride = (long long)N_LAYERS * CONV_S * D_MODEL;
      d.slots = ln.slots;
      d.cpos_in = ln.cpos_in + (reduced ? 1 : 0);
      d.w = L.dw_w;
      d.bias = L.dw_b;
      d.e = ln.ebuf;
      d.T = Tl;
      const int dw_half = (Tl + 1) / 2;   // output frames per thread
      if (e->dw_pipe_min_batch > 0 && B >= e->dw_pipe_min_batch) {
        const dim3 grid(std::min(B, 2 * e->cur_ctas));
        if (Tl <= 5) KLAUNCH(launch_kernel(dwconv_pipe_kernel<5>, grid, dim3(DWP_THREADS), DWP_SMEM, st, e->pdl, d, B));
        else if (Tl <= 7) KLAUNCH(launch_kernel(dwconv_pipe_kernel<7>, grid, dim3(DWP_THREADS), DWP_SMEM, st, e->pdl, d, B));
        else if (Tl <= 10) KLAUNCH(launch_kernel(dwconv_pipe_kernel<10>, grid, dim3(DWP_THREADS), DWP_SMEM, st, e->pdl, d, B));
        else KLAUNCH(launch_kernel(dwconv_pipe_kernel<MAX_T>, grid, dim3(DWP_THREADS), DWP_SMEM, st, e->pdl, d, B));
      } else if (dw_half <= 3) KLAUNCH(launch_kernel(dwconv_kernel<3>, dim3(B, D_MODEL / DW_CH), dim3(DW_THREADS), 0, st, e->pdl, d));
      else if (dw_half <= 5) KLAUNCH(launch_kernel(dwconv_kernel<5>, dim3(B, D_MODEL / DW_CH), dim3(DW_THREADS), 0, st, e->pdl, d));
      else KLAUNCH(launch_kernel(dwconv_kernel<DW_TH>, dim3(B, D_MODEL / DW_CH), dim3(DW_THREADS), 0, st, e->pdl, d));
    }
    if (rowg) {
      RC(run_rowgemm(e, ln, st, M, D_MODEL, ln.m_e, L.pw2, L.pw2_b, 1.f, r, true));
      ss_tiles = 1;
    } else RC(run_resid_rowscale(e, ln, st, M, ln.ebuf, ln.m_e, L.pw2, L.pw2_b, r, &ss_tiles, D_MODEL, 1.f));
    // ---- second feed-forward (norm_feed_forward2 as a row scale), norm_out and what follows the layer
    if (fused_ff) {   // feed-forward 2 + residual + norm_out + the next layer's first norm in one kernel
      const float* g1 = l == 14 ? nullptr : L.n_out;
      const float* g2 = (l == 6 || l >= 14) ? nullptr : e->L[l + 1].n_ff1;
      bf16* n_out = (l == 6 || l == 14) ? nullptr : ln.n;
      RC(run_ff_fused(e, ln, st, M, r, L.ff2_up, L.ff2_up_b, L.ff2_down, L.ff2_down_b, ss_tiles, g1, g2, n_out));
      ff = PartIn();
    } else if (rowg && l != 14) {
      // the down projection owns whole rows: residual add, norm_out in place and the next layer's first norm in its epilogue
      GemmArgs ua = dense_args(M, D_MODEL, ln.rb, ln.h, D_FF, L.ff2_up_b, 1.f);
      ua.ss = ln.ss;
      ua.ss_ld = 12;
      ua.ss_tiles = ss_tiles;
      RC((gemm<G_SWIGLU, BN_SWIGLU>(e, st, ln.m_rb, L.ff2_up, ua, mt, 2 * D_FF / BN_SWIGLU, M, D_FF)));
      const float* g2 = (l == 6 || l == 15) ? nullptr : e->L[l + 1].n_ff1;
      RC(run_rowgemm(e, ln, st, M, D_FF, ln.m_h, L.ff2_down, L.ff2_down_b, 0.5f, r, false, L.n_out, g2, l == 6 ? nullptr : ln.n));
      ff = PartIn();
      ff2_norm_done = true;
    } else {
      RC(run_ff(e, ln, st, M, L.ff2_up, L.ff2_up_b, L.ff2_down, L.ff2_down_b, &ff, ss_tiles));
    }
    if (l == 6) {
      if (!fused_ff && !ff2_norm_done) RC(run_norm(e, ln, st, r, L.n_out, nullptr, nullptr, M, ff));   // r_full = layer output = residual kept for layer 14
      RedArgs ra{ln.r_full, e->st_red, ln.slots, e->red_dw_w, e->red_dw_b, ln.m_red, T, T2};
      KLAUNCH(launch_kernel(reduction_dw_kernel, dim3(B), dim3(D_MODEL), 0, st, e->pdl, ra));
      const int M2 = B * T2;
      GemmArgs a = dense_args(M2, D_FF, ln.m_red, ln.r_red, D_MODEL, e->red_pw_b, 1.f);
      if (M2 >= BIG_M) RC((gemm<G_STORE_F32, 128>(e, st, ln.m_mred, e->red_pw, a, (M2 + 127) / 128, D_MODEL / 128, M2, D_MODEL, nullptr, 1, true)));
      else RC((gemm<G_STORE_F32, BN_STORE>(e, st, ln.m_mred, e->red_pw, a, (M2 + 127) / 128, D_MODEL / BN_STORE, M2, D_MODEL)));
      RC(run_norm(e, ln, st, ln.r_red, nullptr, e->L[7].n_ff1, ln.n, M2));
      RC(tap(1 + l, ln.r_red, M2));
    } else if (l == 14) {
      UpsampleArgs ua{ln.r_full, ln.r_red, ff.part, ff.nsplit, ff.stride, ff.bias, ff.scale,
                      L.n_out, e->L[15].n_ff1, ln.n, B, T, T2};
      if (ff.nsplit > 2) KLAUNCH(launch_kernel(upsample_norm_kernel<MAX_SPLITS>, dim3((B * T + 7) / 8), dim3(256), 0, st, e->pdl, ua));
      else KLAUNCH(launch_kernel(upsample_norm_kernel<2>, dim3((B * T + 7) / 8), dim3(256), 0, st, e->pdl, ua));
      RC(tap(1 + l, ln.r_full, B * T));
    } else if (l == 15) {
      if (!fused_ff && !ff2_norm_done) RC(run_norm(e, ln, st, r, L.n_out, nullptr, ln.n, M, ff));
      RC(tap(1 + l, r, M));
    } else {
      if (!fused_ff && !ff2_norm_done) RC(run_norm(e, ln, st, r, L.n_out, e->L[l + 1].n_ff1, ln.n, M, ff));
      RC(tap(1 + l, r, M));
    }
  }
  {
    M = B * T;
    GemmArgs a = dense_args(M, D_MODEL, ln.n, ln.lp_out, N_CLASSES, e->dec_b, 1.f);
    a.tokens = ln.tok_out;
    a.aux = ln.aux_out;
    RC((gemm<G_DECODER, DEC_PAD>(e, st, ln.m_n, e->dec_w, a, (M + 127) / 128, 1, M, N_CLASSES)));
  }
  return 0;
}

static int check_step_args(tone_engine* e, int B) {
  if (!e) return fail(TONE_EINVAL, "null engine");
  if (!e->finalized) return fail(TONE_ESTATE, "weights not finalized");
  if (B < 1 || B > e->cfg.max_batch) return fail(TONE_EINVAL, "batch %d outside [1, %d]", B, e->cfg.max_batch);
  return 0;
}

enum StepMode : int { SM_FEATURES = 1, SM_PHRASES = 2, SM_PCM16 = 4 };

// Cut the batch into lanes and enqueue their kernel chains: lane 0 on `st`, the others on their own streams between
// a fork event and per-lane join events (works both eagerly and under stream capture).  Inputs / outputs are those of
// staging set `io`.
static int enqueue_step(tone_engine* e, tone_engine::IoSet& io, int B, cudaStream_t st, float* taps, int mode) {
  int nl = 1;
  if (!taps) {
    if (e->lane_min_batch > 0) nl = std::min(e->n_lanes, std::max(1, B / e->lane_min_batch));
    else if (((B * e->T + 127) / 128) * (D_MODEL / 128) > e->num_sms) nl = e->n_lanes;
  }
  nl = std::min(nl, B);
  e->cur_ctas = nl > 1 ? e->lane_ctas : e->num_sms;
  e->cur_lanes = nl;
  e->launches = 0;
  const int per = (B + nl - 1) / nl;
  const int pcm_fmt = (mode & SM_PCM16) ? 1 : 0;
  if (nl > 1) CK(cudaEventRecord(e->fork_ev, st));
  for (int li = 0; li < nl; ++li) {
    tone_engine::Lane& ln = e->lanes[li];
    const int b0 = li * per, nb = std::min(per, B - b0);
    if (nb <= 0) continue;
    ln.slots = io.d_slots + b0;
    ln.pcm = (const char*)io.d_pcm + (size_t)b0 * e->C * (pcm_fmt ? 2 : 4);
    ln.pcm_fmt = pcm_fmt;
    ln.feats = (mode & SM_FEATURES) ? e->d_feats + (size_t)b0 * N_MELS * e->F : nullptr;
    ln.len_in = io.d_len_in + b0;
    ln.cpos_in = io.d_cpos_in + 2 * b0;
    ln.lp_out = io.d_logprobs + (size_t)b0 * e->T * N_CLASSES;
    ln.tok_out = io.d_tokens + (size_t)b0 * e->T;
    ln.aux_out = io.d_aux + (size_t)b0 * e->T * 2;
    cudaStream_t ls = li == 0 ? st : ln.stream;
    if (li > 0) CK(cudaStreamWaitEvent(ls, e->fork_ev, 0));
    RC(run_step(e, ln, nb, ls, taps));
    if (li > 0) {
      CK(cudaEventRecord(ln.done, ls));
      CK(cudaStreamWaitEvent(st, ln.done, 0));
    }
  }
  if (mode & SM_PHRASES) {   // device-side phrase splitter + greedy decode over the whole batch (after the lanes joined)
    PhraseArgs pa;
    pa.slots = io.d_slots;
    pa.tokens = io.d_tokens;
    pa.sil = io.d_aux;
    pa.is_last = io.d_last;
    pa.st = e->st_ph;
    pa.ring = e->st_ring;
    pa.hdr = reinterpret_cast<PhHeader*>(io.d_ph);
    pa.rec = reinterpret_cast<PhRecord*>(io.d_ph + sizeof(PhHeader));
    pa.pool = reinterpret_cast<unsigned char*>(io.d_ph + sizeof(PhHeader) + (size_t)B * PH_PER_STREAM * sizeof(PhRecord));
    pa.B = B;
    pa.T = e->T;
    pa.max_rec = B * PH_PER_STREAM;
    pa.pool_cap = B * 2304;
    KLAUNCH(launch_kernel(ctc_phrase_kernel, dim3(1), dim3(PH_THREADS), 0, st, e->pdl, pa));
  }
  e->launches_per_step = e->launches;
  return 0;
}

// Replay (or capture on first use) the step graph of (batch size, staging set, mode) on `st`.
static int launch_step(tone_engine* e, int set, int B, cudaStream_t st, int mode) {
  tone_engine::IoSet& io = e->io[set];
  NvtxScope nvtx_("tone_step_graph");
  if (!e->cfg.use_graph) return enqueue_step(e, io, B, st, nullptr, mode);
  const uint64_t key = (uint64_t)B | ((uint64_t)set << 24) | ((uint64_t)mode << 28);
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    // Capture on a side stream that has nothing in flight, so the capture neither depends on nor disturbs `st`.
    NvtxScope nvtx_cap("tone_step_capture");
    cudaGraph_t graph = nullptr;
    cudaStream_t cs = e->s_cap;
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_step(e, io, B, cs, nullptr, mode);
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return rc ? rc : fail(TONE_ECUDA, "graph capture: %s", cudaGetErrorString(ce));
    }
    cudaGraphExec_t exec = nullptr;
    cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(TONE_ECUDA, "graph instantiate: %s", cudaGetErrorString(ie));
    if (e->graphs.size() >= 64) {   // bound the cache: a dynamic batcher can name many batch sizes
      CK(cudaDeviceSynchronize());   // a graph may still be running on a caller's stream
      for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
      e->graphs.clear();
    }
    it = e->graphs.emplace(key, exec).first;
  }
  CK(cudaGraphLaunch(it->second, st));
  return 0;
}

// A launch on a caller's stream is spliced into the engine's timeline: the caller's stream first waits for everything
// the engine has enqueued so far (staging copies, the previous step), and the engine's stream then waits for the launch.
static int splice_begin(tone_engine* e, cudaStream_t st) {
  if (st == e->stream) return 0;
  CK(cudaEventRecord(e->ev_sync, e->stream));
  CK(cudaStreamWaitEvent(st, e->ev_sync, 0));
  return 0;
}
static int splice_end(tone_engine* e, cudaStream_t st) {
  if (st == e->stream) return 0;
  CK(cudaEventRecord(e->ev_last, st));
  CK(cudaStreamWaitEvent(e->stream, e->ev_last, 0));
  return 0;
}

// int32 samples -> the int16 wire format, range-checked (tone/onnx_wrapper.py:108-113).  Large batches (the synchronous
// reference-shaped call at 1024 streams narrows 2.4 M samples) are cut across a few host threads.
static void narrow_span(const int32_t* src, int16_t* dst, size_t n, int32_t* lo_out, int32_t* hi_out) {
  int32_t lo = n ? src[0] : 0, hi = lo;
  for (size_t i = 0; i < n; ++i) {
    const int32_t v = src[i];
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
    dst[i] = (int16_t)v;
  }
  *lo_out = lo;
  *hi_out = hi;
}
static int narrow_pcm(const int32_t* src, int16_t* dst, size_t n) {
  constexpr int NT = 4;
  int32_t lo[NT], hi[NT];
  int used = 1;
  if (n >= (size_t)1 << 19) {
    const size_t per = (n + NT - 1) / NT;
    std::thread th[NT - 1];
    int started = 0;
    try {
      for (int t = 1; t < NT; ++t) {
        const size_t a = std::min(n, t * per), b = std::min(n, (t + 1) * per);
        th[t - 1] = std::thread(narrow_span, src + a, dst + a, b - a, &lo[t], &hi[t]);
        ++started;
      }
    } catch (...) {   // no more threads to be had: this thread does the remaining spans itself
    }
    narrow_span(src, dst, std::min(n, per), &lo[0], &hi[0]);
    for (int t = started + 1; t < NT; ++t) {
      const size_t a = std::min(n, t * per), b = std::min(n, (t + 1) * per);
      narrow_span(src + a, dst + a, b - a, &lo[t], &hi[t]);
    }
    for (int t = 1; t <= started; ++t) th[t - 1].join();
    used = NT;
  } else {
    narrow_span(src, dst, n, &lo[0], &hi[0]);
  }
  int32_t l = lo[0], h = hi[0];
  for (int t = 1; t < used; ++t) {
    const size_t a = std::min(n, t * ((n + NT - 1) / NT));
    if (a >= n) continue;             // empty span: its lo / hi are the zero defaults of an empty range
    l = std::min(l, lo[t]);
    h = std::max(h, hi[t]);
  }
  if (l < -32768 || h > 32767)
    return fail(TONE_ERANGE, "Samples in 'audio_chunk' must be in range [-32768; 32767], but it is in range [%d; %d]", (int)l, (int)h);
  return 0;
}

static inline tone_engine::IoSet& legacy_set(tone_engine* e) { return e->io[tone_engine::PIPE]; }

// ------------------------------------------------------------------------------------------------ pipelined step
extern "C" int tone_next_staging(tone_engine* e, int32_t** slots, int16_t** pcm16, uint8_t** is_last) {
  if (!e) return fail(TONE_EINVAL, "null engine");
  tone_engine::IoSet& io = e->io[e->next_ticket % tone_engine::PIPE];
  if (io.busy)   // its H2D copies may still be reading the pinned buffers
    return fail(TONE_ESTATE, "ticket %d of the next staging set has not been waited for", io.ticket);
  if (slots) *slots = io.p_slots;
  if (pcm16) *pcm16 = io.p_pcm;
  if (is_last) *is_last = io.p_last;
  return TONE_OK;
}

extern "C" int tone_submit(tone_engine* e, int32_t B, const int32_t* slots, const void* pcm, int32_t pcm_format,
                           const uint8_t* is_last, int32_t outputs, int32_t* ticket_out) {
  NvtxScope nvtx_("tone_submit");
  RC(check_step_args(e, B));
  if (!slots || !pcm || !ticket_out) return fail(TONE_EINVAL, "null argument");
  if (pcm_format != TONE_PCM_I32 && pcm_format != TONE_PCM_I16) return fail(TONE_EINVAL, "unknown pcm_format %d", pcm_format);
  if (outputs & ~(TONE_OUT_LOGPROBS | TONE_OUT_TOKENS | TONE_OUT_SIL | TONE_OUT_PHRASES))
    return fail(TONE_EINVAL, "unknown output bits 0x%x", outputs);
  CK(cudaSetDevice(e->cfg.device));
  const int set = e->next_ticket % tone_engine::PIPE;
  tone_engine::IoSet& io = e->io[set];
  if (io.busy) return fail(TONE_ESTATE, "ticket %d of this staging set has not been waited for", io.ticket);
  RC(validate_slots(e, B, slots));
  const size_t ns = (size_t)B * e->C;
  if (pcm_format == TONE_PCM_I32) RC(narrow_pcm((const int32_t*)pcm, io.p_pcm, ns));
  else if (pcm != io.p_pcm) memcpy(io.p_pcm, pcm, ns * 2);
  if (slots != io.p_slots) memcpy(io.p_slots, slots, (size_t)B * 4);
  const bool phrases = outputs & TONE_OUT_PHRASES;
  if (phrases) {
    if (is_last && is_last != io.p_last) memcpy(io.p_last, is_last, B);
    else if (!is_last) memset(io.p_last, 0, B);
  }
  // H2D on the input copy stream
  CK(cudaMemcpyAsync(io.d_slots, io.p_slots, (size_t)B * 4, cudaMemcpyHostToDevice, e->s_in));
  CK(cudaMemcpyAsync(io.d_pcm, io.p_pcm, ns * 2, cudaMemcpyHostToDevice, e->s_in));
  if (phrases) CK(cudaMemcpyAsync(io.d_last, io.p_last, B, cudaMemcpyHostToDevice, e->s_in));
  CK(cudaEventRecord(io.ev_in, e->s_in));
  // the step on the compute stream
  CK(cudaStreamWaitEvent(e->stream, io.ev_in, 0));
  RC(launch_step(e, set, B, e->stream, SM_PCM16 | (phrases ? SM_PHRASES : 0)));
  CK(cudaEventRecord(io.ev_step, e->stream));
  // D2H of what was asked for on the output copy stream
  CK(cudaStreamWaitEvent(e->s_out, io.ev_step, 0));
  const size_t rows = (size_t)B * e->T;
  if (outputs & TONE_OUT_LOGPROBS)
    CK(cudaMemcpyAsync(io.p_logprobs, io.d_logprobs, rows * N_CLASSES * 4, cudaMemcpyDeviceToHost, e->s_out));
  if (outputs & TONE_OUT_TOKENS) CK(cudaMemcpyAsync(io.p_tokens, io.d_tokens, rows * 4, cudaMemcpyDeviceToHost, e->s_out));
  if (outputs & TONE_OUT_SIL) CK(cudaMemcpyAsync(io.p_aux, io.d_aux, rows * 8, cudaMemcpyDeviceToHost, e->s_out));
  if (phrases) {   // header + records + the first 32 KB of the text pool; the rest (rare) is fetched by tone_wait
    const size_t fixed = sizeof(PhHeader) + (size_t)B * PH_PER_STREAM * sizeof(PhRecord);
    const size_t n = std::min(fixed + (size_t)B * 2304, fixed + (32u << 10));
    CK(cudaMemcpyAsync(io.p_ph, io.d_ph, n, cudaMemcpyDeviceToHost, e->s_out));
  }
  CK(cudaEventRecord(io.ev_out, e->s_out));
  io.busy = true;
  io.B = B;
  io.outputs = outputs;
  io.ticket = e->next_ticket++;
  io.ph_complete = !phrases;
  *ticket_out = io.ticket;
  return TONE_OK;
}

extern "C" int tone_wait(tone_engine* e, int32_t ticket, float* logprobs, int32_t* tokens, float* sil) {
  NvtxScope nvtx_("tone_wait");
  if (!e) return fail(TONE_EINVAL, "null engine");
  tone_engine::IoSet* io = nullptr;
  for (int k = 0; k < tone_engine::PIPE; ++k)
    if (e->io[k].busy && e->io[k].ticket == ticket) io = &e->io[k];
  if (!io) return fail(TONE_ESTATE, "ticket %d is not in flight", ticket);
  if ((logprobs && !(io->outputs & TONE_OUT_LOGPROBS)) || (tokens && !(io->outputs & TONE_OUT_TOKENS)) ||
      (sil && !(io->outputs & TONE_OUT_SIL)))
    return fail(TONE_EINVAL, "an output was asked for that ticket %d did not request at tone_submit", ticket);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaEventSynchronize(io->ev_out));
  io->busy = false;
  const size_t rows = (size_t)io->B * e->T;
  if (logprobs && logprobs != io->p_logprobs) memcpy(logprobs, io->p_logprobs, rows * N_CLASSES * 4);
  if (tokens && tokens != io->p_tokens) memcpy(tokens, io->p_tokens, rows * 4);
  if (sil && sil != io->p_aux) memcpy(sil, io->p_aux, rows * 8);
  if (io->outputs & TONE_OUT_PHRASES) {
    const PhHeader* h = reinterpret_cast<const PhHeader*>(io->p_ph);
    if (h->overflow) return fail(TONE_ECUDA, "phrase records overflowed (%d phrases, %d text bytes)", h->n_phrases, h->pool_used);
    const size_t fixed = sizeof(PhHeader) + (size_t)io->B * PH_PER_STREAM * sizeof(PhRecord);
    if ((size_t)h->pool_used > (32u << 10)) {
      CK(cudaMemcpyAsync(io->p_ph + fixed + (32u << 10), io->d_ph + fixed + (32u << 10), (size_t)h->pool_used - (32u << 10),
                         cudaMemcpyDeviceToHost, e->s_out));
      CK(cudaStreamSynchronize(e->s_out));
    }
    io->ph_complete = true;
  }
  return TONE_OK;
}

extern "C" int tone_ticket_phrases(tone_engine* e, int32_t ticket, const tone_phrase** phrases, int32_t* n_phrases,
                                   const uint8_t** text_pool, int32_t* text_pool_len) {
  if (!e || !phrases || !n_phrases || !text_pool || !text_pool_len) return fail(TONE_EINVAL, "null argument");
  tone_engine::IoSet* io = nullptr;
  if (ticket == -1) io = &legacy_set(e);   // tone_selftest_phrases
  for (int k = 0; k < tone_engine::PIPE && ticket >= 0; ++k)
    if (e->io[k].ticket == ticket) io = &e->io[k];
  if (!io || io->busy || !io->ph_complete || !(io->outputs & TONE_OUT_PHRASES))
    return fail(TONE_ESTATE, "ticket %d has no phrase records to read (not waited for, reused, or PHRASES not requested)", ticket);
  static_assert(sizeof(tone_phrase) == sizeof(PhRecord), "tone_phrase layout");
  const PhHeader* h = reinterpret_cast<const PhHeader*>(io->p_ph);
  *phrases = reinterpret_cast<const tone_phrase*>(io->p_ph + sizeof(PhHeader));
  *n_phrases = h->n_phrases;
  *text_pool = reinterpret_cast<const uint8_t*>(io->p_ph + sizeof(PhHeader) + (size_t)io->B * PH_PER_STREAM * sizeof(PhRecord));
  *text_pool_len = h->pool_used;
  return TONE_OK;
}

// The synchronous reference-shaped call: one ticket through the same pipeline.
extern "C" int tone_step(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm, float* logprobs,
                         int32_t* tokens) {
  int ticket = -1;
  const int outputs = (logprobs ? TONE_OUT_LOGPROBS : 0) | (tokens ? TONE_OUT_TOKENS : 0);
  RC(tone_submit(e, B, slots, pcm, TONE_PCM_I32, nullptr, outputs, &ticket));
  return tone_wait(e, ticket, logprobs, tokens, nullptr);
}

// ------------------------------------------------------------------------------------------------ staged / device forms
extern "C" int tone_stage(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm) {
  NvtxScope nvtx_("tone_stage");
  RC(check_step_args(e, B));
  if (!slots || !pcm) return fail(TONE_EINVAL, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, B, slots));
  tone_engine::IoSet& io = legacy_set(e);
  CK(cudaStreamSynchronize(e->stream));     // the pinned staging of this set may still be read by an earlier copy
  RC(narrow_pcm(pcm, io.p_pcm, (size_t)B * e->C));
  memcpy(io.p_slots, slots, (size_t)B * 4);
  CK(cudaMemcpyAsync(io.d_slots, io.p_slots, (size_t)B * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(io.d_pcm, io.p_pcm, (size_t)B * e->C * 2, cudaMemcpyHostToDevice, e->stream));
  io.B = B;
  io.staged_mode = SM_PCM16;                // format of what is staged
  return TONE_OK;
}

extern "C" int tone_step_staged(tone_engine* e, int32_t B, void* cuda_stream) {
  NvtxScope nvtx_("tone_step_staged");
  RC(check_step_args(e, B));
  CK(cudaSetDevice(e->cfg.device));
  tone_engine::IoSet& io = legacy_set(e);
  if (io.B != B) return fail(TONE_ESTATE, "%d streams are staged, the step names %d", io.B, B);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
  RC(splice_begin(e, st));
  RC(launch_step(e, tone_engine::PIPE, B, st, io.staged_mode));
  return splice_end(e, st);
}

// Device-pointer form for GPU-resident producers/consumers: inputs are copied device-to-device into the engine's
// staging (the captured graph reads fixed addresses), outputs are copied out the same way.  Stream-ordered, no sync.
extern "C" int tone_step_device(tone_engine* e, int32_t B, const int32_t* slots, const void* d_pcm, int32_t pcm_format,
                                float* d_logprobs, int32_t* d_tokens, void* cuda_stream) {
  NvtxScope nvtx_("tone_step_device");
  RC(check_step_args(e, B));
  if (!slots) return fail(TONE_EINVAL, "null argument");
  if (pcm_format != TONE_PCM_I32 && pcm_format != TONE_PCM_I16) return fail(TONE_EINVAL, "unknown pcm_format %d", pcm_format);
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, B, slots));
  tone_engine::IoSet& io = legacy_set(e);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
  RC(splice_begin(e, st));
  int* ds = nullptr;
  RC(ring_slots(e, B, slots, st, &ds));
  CK(cudaMemcpyAsync(io.d_slots, ds, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  if (d_pcm) {
    CK(cudaMemcpyAsync(io.d_pcm, d_pcm, (size_t)B * e->C * (pcm_format == TONE_PCM_I16 ? 2 : 4), cudaMemcpyDeviceToDevice, st));
    io.staged_mode = pcm_format == TONE_PCM_I16 ? SM_PCM16 : 0;
  } else if (io.B != B) {
    return fail(TONE_ESTATE, "no PCM given and %d streams are staged, the step names %d", io.B, B);
  }
  io.B = B;
  RC(launch_step(e, tone_engine::PIPE, B, st, io.staged_mode));
  if (d_logprobs)
    CK(cudaMemcpyAsync(d_logprobs, io.d_logprobs, (size_t)B * e->T * N_CLASSES * 4, cudaMemcpyDeviceToDevice, st));
  if (d_tokens) CK(cudaMemcpyAsync(d_tokens, io.d_tokens, (size_t)B * e->T * 4, cudaMemcpyDeviceToDevice, st));
  return splice_end(e, st);
}

extern "C" int tone_fetch(tone_engine* e, int32_t B, float* logprobs, int32_t* tokens) {
  RC(check_step_args(e, B));
  CK(cudaSetDevice(e->cfg.device));
  tone_engine::IoSet& io = legacy_set(e);
  const size_t nl = (size_t)B * e->T * N_CLASSES * 4, nt = (size_t)B * e->T * 4;
  // e->stream has waited for the last launch wherever it ran (splice_end), so its order covers the outputs
  if (logprobs) CK(cudaMemcpyAsync(io.p_logprobs, io.d_logprobs, nl, cudaMemcpyDeviceToHost, e->stream));
  if (tokens) CK(cudaMemcpyAsync(io.p_tokens, io.d_tokens, nt, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (logprobs) memcpy(logprobs, io.p_logprobs, nl);
  if (tokens) memcpy(tokens, io.p_tokens, nt);
  return TONE_OK;
}

// Greedy fast path: per frame only the argmax token and the two log-probs the phrase splitter looks at
// (tone/logprob_splitter.py:129: speech iff exp(lp[33]) + exp(lp[34]) <= 0.9) cross PCIe - 12 B instead of 140 B.
extern "C" int tone_fetch_greedy(tone_engine* e, int32_t B, int32_t* tokens, float* sil_logprobs) {
  RC(check_step_args(e, B));
  if (!tokens || !sil_logprobs) return fail(TONE_EINVAL, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  tone_engine::IoSet& io = legacy_set(e);
  const size_t nt = (size_t)B * e->T * 4, na = (size_t)B * e->T * 8;
  CK(cudaMemcpyAsync(io.p_tokens, io.d_tokens, nt, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(io.p_aux, io.d_aux, na, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  memcpy(tokens, io.p_tokens, nt);
  memcpy(sil_logprobs, io.p_aux, na);
  return TONE_OK;
}

extern "C" int tone_sync(tone_engine* e) {
  if (!e) return fail(TONE_EINVAL, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaStreamSynchronize(e->s_out));
  return TONE_OK;
}

// Feature-input mode (reference skip_preprocessor=True, tone/nn/model.py:151-160; the Triton ensemble feeds DALI log-mel
// features, triton/preprocessing/1/features_8k_tone.py): feats = host fp16 [B][64][F], F = chunk_samples / 80.
extern "C" int tone_step_features(tone_engine* e, int32_t B, const int32_t* slots, const uint16_t* feats,
                                  float* logprobs, int32_t* tokens) {
  NvtxScope nvtx_("tone_step_features");
  RC(check_step_args(e, B));
  if (!slots || !feats) return fail(TONE_EINVAL, "null argument");
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, B, slots));
  tone_engine::IoSet& io = legacy_set(e);
  CK(cudaStreamSynchronize(e->stream));
  const size_t nf = (size_t)B * N_MELS * e->F;
  memcpy(io.p_slots, slots, (size_t)B * 4);
  memcpy(e->p_feats, feats, nf * 2);
  CK(cudaMemcpyAsync(io.d_slots, io.p_slots, (size_t)B * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_feats, e->p_feats, nf * 2, cudaMemcpyHostToDevice, e->stream));
  io.B = 0;                                 // no PCM is staged
  RC(launch_step(e, tone_engine::PIPE, B, e->stream, SM_FEATURES));
  return tone_fetch(e, B, logprobs, tokens);
}

extern "C" int tone_step_debug(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm, float* logprobs,
                               int32_t* tokens, float* taps) {
  RC(tone_stage(e, B, slots, pcm));
  RC(enqueue_step(e, legacy_set(e), B, e->stream, taps, SM_PCM16));
  return tone_fetch(e, B, logprobs, tokens);
}

// Debug: the device-side splitter alone on host-supplied per-frame inputs, `frames` frames per stream fed in pieces
// of at most 13 (the kernel's chunk bound); is_last applies to the final piece.
extern "C" int tone_selftest_phrases(tone_engine* e, int32_t B, const int32_t* slots, int32_t frames, const int32_t* tokens,
                                     const float* sil, const uint8_t* is_last) {
  RC(check_step_args(e, B));
  if (!slots || !tokens || !sil || frames < 1 || frames > MAX_T)
    return fail(TONE_EINVAL, "need 1 <= frames <= %d per call", (int)MAX_T);
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, B, slots));
  tone_engine::IoSet& io = legacy_set(e);
  CK(cudaStreamSynchronize(e->stream));
  memcpy(io.p_slots, slots, (size_t)B * 4);
  memcpy(io.p_tokens, tokens, (size_t)B * frames * 4);
  memcpy(io.p_aux, sil, (size_t)B * frames * 8);
  if (is_last) memcpy(io.p_last, is_last, B);
  else memset(io.p_last, 0, B);
  CK(cudaMemcpyAsync(io.d_slots, io.p_slots, (size_t)B * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(io.d_tokens, io.p_tokens, (size_t)B * frames * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(io.d_aux, io.p_aux, (size_t)B * frames * 8, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(io.d_last, io.p_last, B, cudaMemcpyHostToDevice, e->stream));
  PhraseArgs pa;
  pa.slots = io.d_slots;
  pa.tokens = io.d_tokens;
  pa.sil = io.d_aux;
  pa.is_last = io.d_last;
  pa.st = e->st_ph;
  pa.ring = e->st_ring;
  pa.hdr = reinterpret_cast<PhHeader*>(io.d_ph);
  pa.rec = reinterpret_cast<PhRecord*>(io.d_ph + sizeof(PhHeader));
  pa.pool = reinterpret_cast<unsigned char*>(io.d_ph + sizeof(PhHeader) + (size_t)B * PH_PER_STREAM * sizeof(PhRecord));
  pa.B = B;
  pa.T = frames;
  pa.max_rec = B * PH_PER_STREAM;
  pa.pool_cap = B * 2304;
  KLAUNCH(launch_kernel(ctc_phrase_kernel, dim3(1), dim3(PH_THREADS), 0, e->stream, false, pa));
  CK(cudaMemcpyAsync(io.p_ph, io.d_ph, e->ph_bytes, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  io.B = B;
  io.busy = false;
  io.ticket = -1;
  io.ph_complete = true;
  io.outputs = TONE_OUT_PHRASES;
  return TONE_OK;
}

// ------------------------------------------------------------------------------------------------ state wire formats
// Flat fp16 order (tone/nn/model.py:259-267): preproc 80 | mhsa (2,30,384) | conv (16,384,30) | len 1 |
// sub1 (1,10,64) | sub2 (32,8,44) | reduction (384,1): gathered / scattered on the device (state_io.cuh), one kernel and
// one copy per STATE_IO_CHUNK slots.  The phrase-splitter state is not part of the model state and is left untouched.
extern "C" int tone_export_states(tone_engine* e, int32_t n, const int32_t* slots, uint16_t* out) {
  NvtxScope nvtx_("tone_export_states");
  if (!e || !slots || !out || n < 0) return fail(TONE_EINVAL, "bad argument");
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, n, slots, false));
  for (int i0 = 0; i0 < n; i0 += STATE_IO_CHUNK) {
    const int m = std::min((int)STATE_IO_CHUNK, n - i0);
    int* d = nullptr;
    RC(ring_slots(e, m, slots + i0, e->stream, &d));
    export_state_kernel<<<dim3(m, 24), STATE_IO_THREADS, 0, e->stream>>>(state_pool(e), d, reinterpret_cast<__half*>(e->d_state_io));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->p_state_io, e->d_state_io, (size_t)m * TONE_STATE_SIZE * 2, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(out + (size_t)i0 * TONE_STATE_SIZE, e->p_state_io, (size_t)m * TONE_STATE_SIZE * 2);
  }
  return TONE_OK;
}

extern "C" int tone_import_states(tone_engine* e, int32_t n, const int32_t* slots, const uint16_t* in) {
  NvtxScope nvtx_("tone_import_states");
  if (!e || !slots || !in || n < 0) return fail(TONE_EINVAL, "bad argument");
  CK(cudaSetDevice(e->cfg.device));
  RC(validate_slots(e, n, slots, false));
  for (int i0 = 0; i0 < n; i0 += STATE_IO_CHUNK) {
    const int m = std::min((int)STATE_IO_CHUNK, n - i0);
    CK(cudaStreamSynchronize(e->stream));   // the pinned staging may still be read by the previous chunk's copy
    memcpy(e->p_state_io, in + (size_t)i0 * TONE_STATE_SIZE, (size_t)m * TONE_STATE_SIZE * 2);
    int* d = nullptr;
    RC(ring_slots(e, m, slots + i0, e->stream, &d));
    CK(cudaMemcpyAsync(e->d_state_io, e->p_state_io, (size_t)m * TONE_STATE_SIZE * 2, cudaMemcpyHostToDevice, e->stream));
    import_state_kernel<<<dim3(m, 24), STATE_IO_THREADS, 0, e->stream>>>(state_pool(e), d, reinterpret_cast<const __half*>(e->d_state_io));
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(e->stream));
  return TONE_OK;
}

// Three-tensor Triton cache layout (tone/scripts/export.py:293-376): cache_last_time (18,384,30) = [mhsa transposed to
// (2,384,30) | conv (16,384,30)]; cache_last_channel (32,8,50) = [sub2 (32,8,44) | a (32,8,6) tail holding preproc 80,
// sub1 640, reduction 384 and zero padding, flattened in that order]; cache_last_chan_len = mhsa_len.
namespace {
const int TR_TIME = 18 * 384 * 30, TR_CHAN = 32 * 8 * 50, TR_TPAD = 6;
void flat_to_triton(const uint16_t* f, uint16_t* tm, uint16_t* ch, int64_t* len) {
  for (int l = 0; l < 2; ++l)
    for (int c = 0; c < 384; ++c)
      for (int t = 0; t < 30; ++t) tm[(l * 384 + c) * 30 + t] = f[st_off::mhsa + (l * 30 + t) * 384 + c];
  memcpy(tm + 2 * 384 * 30, f + st_off::conv, (size_t)16 * 384 * 30 * 2);
  uint16_t tail[32 * 8 * TR_TPAD];
  memset(tail, 0, sizeof(tail));
  memcpy(tail, f + st_off::pre, 80 * 2);
  memcpy(tail + 80, f + st_off::sub1, 640 * 2);
  memcpy(tail + 720, f + st_off::red, 384 * 2);
  for (int cr = 0; cr < 32 * 8; ++cr) {
    memcpy(ch + cr * 50, f + st_off::sub2 + cr * 44, 44 * 2);
    memcpy(ch + cr * 50 + 44, tail + cr * TR_TPAD, TR_TPAD * 2);
  }
  *len = (int64_t)lrintf(h2f(f[st_off::len]));
}
void triton_to_flat(const uint16_t* tm, const uint16_t* ch, int64_t len, uint16_t* f) {
  for (int l = 0; l < 2; ++l)
    for (int c = 0; c < 384; ++c)
      for (int t = 0; t < 30; ++t) f[st_off::mhsa + (l * 30 + t) * 384 + c] = tm[(l * 384 + c) * 30 + t];
  memcpy(f + st_off::conv, tm + 2 * 384 * 30, (size_t)16 * 384 * 30 * 2);
  uint16_t tail[32 * 8 * TR_TPAD];
  for (int cr = 0; cr < 32 * 8; ++cr) {
    memcpy(f + st_off::sub2 + cr * 44, ch + cr * 50, 44 * 2);
    memcpy(tail + cr * TR_TPAD, ch + cr * 50 + 44, TR_TPAD * 2);
  }
  memcpy(f + st_off::pre, tail, 80 * 2);
  memcpy(f + st_off::sub1, tail + 80, 640 * 2);
  memcpy(f + st_off::red, tail + 720, 384 * 2);
  f[st_off::len] = f2h((float)len);
}
}  // namespace

extern "C" int tone_export_states_triton(tone_engine* e, int32_t n, const int32_t* slots, uint16_t* cache_last_time,
                                         uint16_t* cache_last_channel, int64_t* cache_last_chan_len) {
  if (!e || !slots || !cache_last_time || !cache_last_channel || !cache_last_chan_len || n < 0)
    return fail(TONE_EINVAL, "bad argument");
  std::vector<uint16_t> flat((size_t)TONE_STATE_SIZE);
  for (int i = 0; i < n; ++i) {
    RC(tone_export_states(e, 1, slots + i, flat.data()));
    flat_to_triton(flat.data(), cache_last_time + (size_t)i * TR_TIME, cache_last_channel + (size_t)i * TR_CHAN,
                   cache_last_chan_len + i);
  }
  return TONE_OK;
}

extern "C" int tone_import_states_triton(tone_engine* e, int32_t n, const int32_t* slots, const uint16_t* cache_last_time,
                                         const uint16_t* cache_last_channel, const int64_t* cache_last_chan_len) {
  if (!e || !slots || !cache_last_time || !cache_last_channel || !cache_last_chan_len || n < 0)
    return fail(TONE_EINVAL, "bad argument");
  std::vector<uint16_t> flat((size_t)TONE_STATE_SIZE);
  for (int i = 0; i < n; ++i) {
    triton_to_flat(cache_last_time + (size_t)i * TR_TIME, cache_last_channel + (size_t)i * TR_CHAN, cache_last_chan_len[i],
                   flat.data());
    RC(tone_import_states(e, 1, slots + i, flat.data()));
  }
  return TONE_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM self-test
template <int BN>
static int selftest_bn(tone_engine* e, int M, int N, int K, const bf16* dA, const bf16* dW, float* dC) {
  CUtensorMap ma, mw;
  RC(make_map_2d(e, &ma, (void*)dA, M, K, 128, false));
  RC(make_map_2d(e, &mw, (void*)dW, N, K, BN, true));
  GemmArgs a = dense_args(M, K, dA, dC, N, nullptr, 1.f);
  a.W = dW;
  a.ldw = K;
  CK(cudaDeviceSynchronize());   // operands were uploaded with cudaMemcpy from pageable memory
  cudaError_t err = launch_gemm_tc<G_STORE_F32, BN>(e->stream, ma, ma, mw, a, (M + 127) / 128, N / BN, false, e->num_sms);
  if (err != cudaSuccess) return fail(TONE_ECUDA, "selftest launch: %s", cudaGetErrorString(err));
  return 0;
}

extern "C" int tone_selftest_gemm(tone_engine* e, int32_t M, int32_t N, int32_t K, const float* A, const float* Wm,
                                  float* Cout, int32_t block_n) {
  if (!e || !A || !Wm || !Cout) return fail(TONE_EINVAL, "null argument");
  if (K % 64 || N % block_n || M < 1) return fail(TONE_EINVAL, "need K %% 64 == 0 and N %% block_n == 0");
  CK(cudaSetDevice(e->cfg.device));
  std::vector<uint16_t> a((size_t)M * K), w((size_t)N * K);
  for (size_t i = 0; i < a.size(); ++i) a[i] = f2bf(A[i]);
  for (size_t i = 0; i < w.size(); ++i) w[i] = f2bf(Wm[i]);
  bf16 *dA = nullptr, *dW = nullptr;
  float* dC = nullptr;
  const size_t Mp = ((size_t)M + 127) / 128 * 128;
  CK(cudaMalloc((void**)&dA, Mp * K * 2));
  CK(cudaMemset(dA, 0, Mp * K * 2));
  CK(cudaMalloc((void**)&dW, w.size() * 2));
  CK(cudaMalloc((void**)&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  int rc;
  if (block_n == 32) rc = selftest_bn<32>(e, M, N, K, dA, dW, dC);
  else if (block_n == 64) rc = selftest_bn<64>(e, M, N, K, dA, dW, dC);
  else if (block_n == 128) rc = selftest_bn<128>(e, M, N, K, dA, dW, dC);
  else rc = fail(TONE_EINVAL, "block_n must be 32, 64 or 128");
  if (!rc) {
    cudaError_t se = cudaStreamSynchronize(e->stream);
    if (se != cudaSuccess) rc = fail(TONE_ECUDA, "selftest run: %s", cudaGetErrorString(se));
  }
  if (!rc) {
    cudaError_t ce = cudaMemcpy(Cout, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) rc = fail(TONE_ECUDA, "selftest copy: %s", cudaGetErrorString(ce));
  }
  cudaFree(dA);
  cudaFree(dW);
  cudaFree(dC);
  return rc;
}

// ------------------------------------------------------------------------------------------------ timeline (diagnostic build)
#ifdef TONE_PROF
static ProfRec* g_prof_dev = nullptr;
static int g_prof_cap = 0;
extern "C" int tone_prof_start(tone_engine* e, int32_t max_records) {
  if (!e) return fail(TONE_EINVAL, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  if (g_prof_dev) cudaFree(g_prof_dev);
  CK(cudaMalloc((void**)&g_prof_dev, (size_t)max_records * sizeof(ProfRec)));
  CK(cudaMemset(g_prof_dev, 0, (size_t)max_records * sizeof(ProfRec)));
  g_prof_cap = max_records;
  unsigned int zero = 0;
  CK(cudaMemcpyToSymbol(g_prof_n, &zero, sizeof(zero)));
  CK(cudaMemcpyToSymbol(g_prof, &g_prof_dev, sizeof(g_prof_dev)));
  CK(cudaDeviceSynchronize());
  return TONE_OK;
}
// out: [n][10] uint64 = g0, g1, c0..c5, id, grid ; returns the number of records through *n_out and stops recording
extern "C" int tone_prof_read(tone_engine* e, unsigned long long* out, int32_t max_records, int32_t* n_out) {
  if (!e || !out || !n_out) return fail(TONE_EINVAL, "null argument");
  CK(cudaDeviceSynchronize());
  unsigned int n = 0;
  CK(cudaMemcpyFromSymbol(&n, g_prof_n, sizeof(n)));
  ProfRec* none = nullptr;
  CK(cudaMemcpyToSymbol(g_prof, &none, sizeof(none)));
  int m = (int)std::min<unsigned int>(n, (unsigned int)std::min(max_records, g_prof_cap));
  CK(cudaMemcpy(out, g_prof_dev, (size_t)m * sizeof(ProfRec), cudaMemcpyDeviceToHost));
  *n_out = m;
  return TONE_OK;
}
#endif
