/*
 * tone_b200.h - C ABI of the B200-native T-one streaming acoustic-model step.
 *
 * One `tone_engine` per GPU.  It owns the packed weights, a pool of per-stream state slots
 * resident in HBM, the per-step scratch and the captured CUDA graphs.  The entry points are
 * what a host-language binding for the reference's acoustic-model boundary would bind; each
 * one cites the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions: every function returns 0 on success or a negative TONE_E* code and never
 * throws across the ABI; tone_last_error() returns a thread-local description of the last
 * failure.  A handle is not thread-safe: one stepping thread per engine (the reference is a
 * synchronous call as well, tone/onnx_wrapper.py:84-123).  Plain pointers and sizes only.
 */
#ifndef TONE_B200_H
#define TONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TONE_OK 0
#define TONE_EINVAL (-1)   /* bad argument (shape / range / null)              -> ValueError */
#define TONE_ENOMEM (-2)   /* out of slots or device memory                                  */
#define TONE_ECUDA (-3)    /* CUDA runtime / driver failure, see tone_last_error()           */
#define TONE_ESTATE (-4)   /* call order: weights not finalized, slot not allocated, ...     */
#define TONE_ERANGE (-5)   /* PCM sample outside [-32768, 32767] (tone/onnx_wrapper.py:108)  */

#define TONE_STATE_SIZE 219729 /* fp16 elements per stream, tone/onnx_wrapper.py:34 */
#define TONE_N_CLASSES 35      /* 34 labels + CTC blank, tone/decoder.py:23         */

typedef struct tone_engine tone_engine;

/* Replaces the choice of exported model + session options
 * (tone/onnx_wrapper.py:66-78, configs/streaming_acoustic/config.pbtxt:1-44). */
typedef struct tone_config {
  int32_t device;        /* CUDA ordinal                                                    */
  int32_t chunk_samples; /* 2400 (300 ms, tone/onnx_wrapper.py:32) or 3200 (400 ms variant,  */
                         /* dev/triton/client_wer.py:277-278)                                */
  int32_t max_slots;     /* resident streams (state pool capacity)                           */
  int32_t max_batch;     /* largest B accepted by tone_step                                  */
  int32_t gemm_impl;     /* 0 = tcgen05/TMEM/TMA (product path), 1 = SIMT debug kernels      */
  int32_t use_graph;     /* 1 = replay a captured CUDA graph per batch size, 0 = eager       */
  int32_t cluster_max_batch; /* experimental latency path: batches up to this size run the 16  */
                         /* Conformer layers + decoder as ONE thread-block-cluster kernel     */
                         /* (csrc/encoder_cluster.cuh); 0 = off (default)                    */
} tone_config;

/* Shapes a caller needs to size its buffers (tone/onnx_wrapper.py:30-34,
 * configs/streaming_acoustic/config.pbtxt:5-33). */
typedef struct tone_info {
  int32_t chunk_samples;    /* samples per stream per step                   */
  int32_t frames_out;       /* T: logprob frames per step (10 | 13)          */
  int32_t n_classes;        /* 35                                            */
  int32_t state_size;       /* 219729                                        */
  int32_t max_slots, max_batch;
  int32_t launches_per_step; /* kernels launched by one step (filled after the first step) */
  int32_t n_taps;            /* rows of the debug tap buffer: 1 + n_layers   */
  int64_t state_bytes_per_slot;
  int64_t weight_bytes;
} tone_info;

int tone_create(const tone_config* cfg, tone_engine** out);
void tone_destroy(tone_engine* e);
int tone_get_info(const tone_engine* e, tone_info* out);
const char* tone_last_error(void);

/* Weights: one call per tensor with the reference state_dict name (without the `tone.`
 * prefix; tone/training/model_wrapper.py:146,156) and fp32 data in the reference's shape.
 * Replaces ort.InferenceSession(model_path) (tone/onnx_wrapper.py:77).  tone_finalize_weights
 * folds/permutes/rounds them into the device layouts and uploads. */
int tone_load_weight(tone_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int tone_finalize_weights(tone_engine* e);

/* Stream slots: server-resident per-stream state, the role of Triton's implicit sequence
 * state (triton/model/config.pbtxt:26-69).  A fresh slot holds the all-zero initial state
 * (tone/nn/model.py:208-267, tone/onnx_wrapper.py:114-115). */
int tone_alloc_slots(tone_engine* e, int32_t n, int32_t* slots_out);
int tone_release_slots(tone_engine* e, int32_t n, const int32_t* slots);
int tone_reset_slots(tone_engine* e, int32_t n, const int32_t* slots);

/* The step: replaces `_ort_sess.run(None, {"signal","state"})` (tone/onnx_wrapper.py:123).
 *   pcm       host int32 [B][chunk_samples]   ("signal", int16 range)
 *   logprobs  host fp32  [B][frames_out][35]  ("logprobs"), may be NULL
 *   tokens    host int32 [B][frames_out]      per-frame argmax (first max, tone/decoder.py:57), may be NULL
 * State is read from and written back to slots[b] on the device.  Synchronous: returns when
 * the outputs are in the host buffers.  H2D of pcm and D2H of the outputs are part of the call. */
int tone_step(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm,
              float* logprobs, int32_t* tokens);

/* Feature-input form of the step: replaces the exported graph built with `--skip-preprocessor`
 * (tone/nn/model.py:151-160, tone/scripts/export.py:48-49), i.e. the acoustic model behind an external log-mel front
 * end such as the Triton/DALI ensemble (triton/preprocessing/1/features_8k_tone.py).
 *   feats  host fp16 [B][64][chunk_samples / 80]  log-mel features, (B, C = n_mels, T) like the reference
 * The waveform state of the slots is left untouched; everything else is tone_step. */
int tone_step_features(tone_engine* e, int32_t B, const int32_t* slots, const uint16_t* feats_fp16,
                       float* logprobs, int32_t* tokens);

/* Same step with inputs/outputs left in HBM: stage once, then step any number of times on
 * the staged chunk (benchmark "inputs already resident" leg).  tone_fetch copies the last
 * outputs to the host. */
int tone_stage(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm);
int tone_step_staged(tone_engine* e, int32_t B, void* cuda_stream);
int tone_fetch(tone_engine* e, int32_t B, float* logprobs, int32_t* tokens);
int tone_sync(tone_engine* e);

/* Greedy fast path: instead of the full log-probs, fetch per frame the argmax token and the two
 * log-probs the phrase splitter thresholds on - ' ' (id 33) and blank (id 34), tone/logprob_splitter.py:129.
 *   tokens        host int32 [B][frames_out]
 *   sil_logprobs  host fp32  [B][frames_out][2]
 * Greedy decoding of a phrase needs nothing else (tone/decoder.py:57-59), so D2H drops from 140 to 12 bytes/frame. */
int tone_fetch_greedy(tone_engine* e, int32_t B, int32_t* tokens, float* sil_logprobs);

/* Device-pointer form of the step for GPU-resident producers/consumers (the role of Triton's
 * GPU tensors between ensemble stages, triton/ensemble/config.pbtxt:20-55): slots / pcm are
 * device int32 buffers, logprobs / tokens device outputs (any may be NULL = use what is staged /
 * leave in the engine).  Enqueued on `cuda_stream` (NULL = the engine's stream); asynchronous. */
int tone_step_device(tone_engine* e, int32_t B, const int32_t* d_slots, const int32_t* d_pcm,
                     float* d_logprobs, int32_t* d_tokens, void* cuda_stream);

/* Pinned host staging buffers owned by the engine (slots [max_batch], pcm [max_batch][chunk],
 * logprobs [max_batch][T][35], tokens [max_batch][T]).  Passing these pointers to tone_step /
 * tone_stage / tone_fetch skips the intermediate host copy. */
int tone_host_buffers(tone_engine* e, int32_t** slots, int32_t** pcm, float** logprobs, int32_t** tokens);

/* State wire format: the reference's flat fp16 vector ("state"/"state_next",
 * configs/streaming_acoustic/config.pbtxt:12-33; element order = get_initial_state order,
 * tone/nn/model.py:259-267).  Used for parity, checkpoint/resume and stream migration. */
int tone_export_state(tone_engine* e, int32_t slot, uint16_t* fp16_out /* [219729] */);
int tone_import_state(tone_engine* e, int32_t slot, const uint16_t* fp16_in /* [219729] */);

/* Debug: run one eager step that also records the residual stream after pre-encode and after
 * every Conformer layer.  taps: host fp32 [1+n_layers][B*frames_out][384] (rows of reduced
 * layers 7..14 occupy the first B*T2 rows).  Not a product path. */
int tone_step_debug(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm,
                    float* logprobs, int32_t* tokens, float* taps);

/* Debug: C[M][N] = A[M][K] * W[N][K]^T through the product GEMM kernel (bf16 operands given as
 * fp32, rounded on upload; fp32 result).  Used by the GPU unit tests. */
int tone_selftest_gemm(tone_engine* e, int32_t M, int32_t N, int32_t K, const float* A, const float* W,
                       float* C, int32_t block_n);

/* Debug: diagnostics of the experimental cluster (latency) path.  out (nullable): 6144 uint64 of in-kernel
 * timestamps of the last cluster-kernel launch (needs TONE_CL_PROF=1 at tone_create); max_active (nullable):
 * co-resident clusters reported by the occupancy query, small-group * 1000 + large-group instantiation. */
int tone_cluster_prof_read(tone_engine* e, unsigned long long* out, int32_t* max_active);

#ifdef __cplusplus
}
#endif
#endif /* TONE_B200_H */
