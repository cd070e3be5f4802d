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
 * failure.  Plain pointers and sizes only.
 *
 * Threading and stream ordering.  A handle is not thread-safe: one stepping thread per engine
 * (the reference is a synchronous call as well, tone/onnx_wrapper.py:84-123).  All work of an
 * engine is ordered on its own streams; the engine records an event after every launch, and
 * every entry point that reads results (tone_fetch*, tone_wait, tone_export_*) or overwrites
 * inputs waits on the events it depends on.  Entry points that take a caller `cuda_stream` make
 * that stream wait for the engine's pending input copies and previous step before launching, and
 * the engine's later work waits for that launch; the caller only has to keep its own buffers
 * alive and unchanged until the work it enqueued on its stream has completed.
 *
 * Slot ids within one batch must be distinct (two entries of one stream in one step would be a
 * data race on that stream's caches): every entry point validates range, allocation and
 * uniqueness on the host and returns TONE_EINVAL / TONE_ESTATE.
 */
#ifndef TONE_B200_H
#define TONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TONE_OK 0
#define TONE_EINVAL (-1)   /* bad argument (shape / null / duplicate slot)         -> ValueError */
#define TONE_ENOMEM (-2)   /* out of slots or device memory                                      */
#define TONE_ECUDA (-3)    /* CUDA runtime / driver failure, see tone_last_error()               */
#define TONE_ESTATE (-4)   /* call order: weights not finalized, slot not allocated, ticket ...  */
#define TONE_ERANGE (-5)   /* PCM sample outside [-32768, 32767] (tone/onnx_wrapper.py:108-113)  */

#define TONE_STATE_SIZE 219729 /* fp16 elements per stream, tone/onnx_wrapper.py:34 */
#define TONE_N_CLASSES 35      /* 34 labels + CTC blank, tone/decoder.py:23         */

/* PCM sample formats accepted on the wire.  The reference's "signal" tensor is int32 holding int16-range
 * values (configs/streaming_acoustic/config.pbtxt:5-11, tone/onnx_wrapper.py:104-113); telephony audio is
 * int16 natively, so the int16 form halves the H2D bytes and cannot be out of range. */
#define TONE_PCM_I32 0
#define TONE_PCM_I16 1

/* What a step leaves for the host (bit mask, `outputs` argument of tone_submit):
 *   LOGPROBS  fp32 [B][T][35]                        ("logprobs", tone/onnx_wrapper.py:123)
 *   TOKENS    int32 [B][T] per-frame argmax          (tone/decoder.py:57)
 *   SIL       fp32 [B][T][2] log-probs of ' ' / blank (what tone/logprob_splitter.py:129 thresholds on)
 *   PHRASES   finished phrases of the device-side splitter + greedy decoder (tone_phrase records) */
#define TONE_OUT_LOGPROBS 1
#define TONE_OUT_TOKENS 2
#define TONE_OUT_SIL 4
#define TONE_OUT_PHRASES 8

typedef struct tone_engine tone_engine;

/* Replaces the choice of exported model + session options
 * (tone/onnx_wrapper.py:66-78, configs/streaming_acoustic/config.pbtxt:1-44).
 * Tuning fields: 0 selects the engine default everywhere. */
typedef struct tone_config {
  int32_t device;        /* CUDA ordinal                                                    */
  int32_t chunk_samples; /* 2400 (300 ms, tone/onnx_wrapper.py:32) or 3200 (400 ms variant,  */
                         /* dev/triton/client_wer.py:277-278)                                */
  int32_t max_slots;     /* resident streams (state pool capacity)                           */
  int32_t max_batch;     /* largest B accepted by a step                                     */
  int32_t gemm_impl;     /* 0 = tcgen05/TMEM/TMA (product path), 1 = SIMT debug kernels      */
  int32_t use_graph;     /* 1 = replay a captured CUDA graph per batch size, 0 = eager       */
  /* ---- tuning (0 = default) */
  int32_t lanes;             /* concurrent sub-batches a large step is cut into (1..4; default 2)           */
  int32_t lane_min_batch;    /* streams per lane below which the batch is not cut (default: cut when one     */
                             /* lane's 384-column GEMMs exceed one tile per SM, i.e. from 628 streams x 10 frames) */
  int32_t persist_min_tiles; /* dense GEMMs with at least this many 128x128 tiles run as the persistent      */
                             /* kernel (default: SM count + 1; -1 = never)                                   */
  int32_t persist_mode;      /* gated persistent GEMMs: 1 = 128-wide tiles, 2 = 256-wide (default),          */
                             /* 3 = 256-wide on CTA pairs (cta_group::2)                                     */
  int32_t split_k;           /* split-K factor of the feed-forward down projection (default: fill the SMs)   */
  int32_t flags;             /* TONE_FLAG_* bits                                                             */
  int32_t fused_ff;          /* feed-forward module as one kernel per row tile (experimental, opt-in):       */
                             /* 1 = off (default), 2 = one CTA per 128 rows, 3 = CTA pairs                   */
  int32_t fused_ff_min_rows; /* rows per lane from which the fused feed-forward is used (default 2048)       */
  int32_t att_block_min_rows; /* rows per lane from which a score-sharing attention layer runs as ONE kernel  */
                             /* per tile of whole streams: V projection + P.V + out projection + residual     */
                             /* (default 16384: pays from ~4096 streams per GPU; -1 = never)                   */
  int32_t lazy_norm_min_rows; /* rows per lane from which feed-forward 1 adds straight into the residual      */
                             /* stream and norm_self_att becomes a row scale inside the projection GEMMs      */
                             /* (default 4096; -1 = never)                                                    */
  int32_t dw_pipe_min_batch; /* streams per lane from which the depthwise conv runs as the pipelined          */
                             /* persistent kernel (default 128; -1 = never)                                   */
  int32_t att_pipe_min_batch; /* streams per lane from which the recompute attention layers (0, 7, 14, 15) run */
                             /* as the pipelined persistent kernel (default 256; -1 = never)                  */
  int32_t rowgemm_min_rows;  /* rows per lane from which the N = 384 projections (feed-forward down, attention */
                             /* out, pointwise conv 2) run as the row-owner CTA-pair kernel with the residual  */
                             /* add and the norms in its epilogue (default 8192; -1 = never)                   */
  int32_t persist_ctas;      /* CTAs of a persistent kernel when the step runs in more than one lane          */
                             /* (default: the SM count; SM count / lanes gives every lane its own SMs)        */
} tone_config;

#define TONE_FLAG_NO_PDL 1         /* launch the kernels of a step without programmatic dependent launch     */
#define TONE_FLAG_NO_FUSED_VATT 2  /* score-sharing layers: V projection and P.V as two kernels              */

/* Shapes a caller needs to size its buffers (tone/onnx_wrapper.py:30-34,
 * configs/streaming_acoustic/config.pbtxt:5-33). */
typedef struct tone_info {
  int32_t chunk_samples;    /* samples per stream per step                   */
  int32_t frames_out;       /* T: logprob frames per step (10 | 13)          */
  int32_t n_classes;        /* 35                                            */
  int32_t state_size;       /* 219729                                        */
  int32_t max_slots, max_batch;
  int32_t launches_per_step; /* kernels launched by the last enqueued step                   */
  int32_t n_taps;            /* rows of the debug tap buffer: 1 + n_layers   */
  int64_t state_bytes_per_slot;
  int64_t weight_bytes;
  int32_t pipeline_depth;    /* staging sets of tone_submit / tone_wait (2)  */
  int32_t max_phrases_per_step; /* capacity of the phrase records of one step: 4 * max_batch */
} tone_info;

int tone_create(const tone_config* cfg, tone_engine** out);
void tone_destroy(tone_engine* e);
int tone_get_info(const tone_engine* e, tone_info* out);
const char* tone_last_error(void);

/* Weights: one call per tensor with the reference state_dict name (with or without the `tone.`
 * prefix; tone/training/model_wrapper.py:146,156) and fp32 data in the reference's shape.
 * Replaces ort.InferenceSession(model_path) (tone/onnx_wrapper.py:77).  tone_finalize_weights
 * folds/permutes/rounds them into the device layouts and uploads. */
int tone_load_weight(tone_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int tone_finalize_weights(tone_engine* e);

/* Stream slots: server-resident per-stream state, the role of Triton's implicit sequence
 * state (triton/model/config.pbtxt:26-69).  A fresh slot holds the all-zero initial state
 * (tone/nn/model.py:208-267, tone/onnx_wrapper.py:114-115) and an empty phrase-splitter state
 * (tone/logprob_splitter.py:33-38). */
int tone_alloc_slots(tone_engine* e, int32_t n, int32_t* slots_out);
int tone_release_slots(tone_engine* e, int32_t n, const int32_t* slots);
int tone_reset_slots(tone_engine* e, int32_t n, const int32_t* slots);

/* The step: replaces `_ort_sess.run(None, {"signal","state"})` (tone/onnx_wrapper.py:123).
 *   pcm       host int32 [B][chunk_samples]   ("signal"); a sample outside the int16 range -> TONE_ERANGE
 *   logprobs  host fp32  [B][frames_out][35]  ("logprobs"), may be NULL
 *   tokens    host int32 [B][frames_out]      per-frame argmax (first max, tone/decoder.py:57), may be NULL
 * State is read from and written back to slots[b] on the device.  Synchronous: returns when
 * the outputs are in the host buffers.  H2D of pcm and D2H of the outputs are part of the call. */
int tone_step(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm,
              float* logprobs, int32_t* tokens);

/* Pipelined form of the step for serving loops: tone_submit enqueues (H2D on a copy stream -> step -> D2H on a
 * second copy stream) on one of `pipeline_depth` staging sets and returns a ticket immediately; tone_wait blocks
 * until that ticket's outputs are on the host.  With two tickets in flight the copies of step i+1 / i-1 overlap
 * the kernels of step i.  Steps execute in submission order (a stream may appear in consecutive tickets).
 *   pcm         host [B][chunk_samples] int32 or int16 (pcm_format); int32 is range-checked (TONE_ERANGE)
 *   is_last     host uint8 [B] or NULL: 1 = this is the stream's last chunk, flush its unfinished phrase
 *               (`is_last` of tone/logprob_splitter.py:91-97 / tone/pipeline.py:200); only used with PHRASES
 *   outputs     TONE_OUT_* mask: only these are produced / copied back
 * tone_wait copies into the caller's buffers (each nullable; must have been requested): logprobs [B][T][35],
 * tokens [B][T], sil [B][T][2]; phrases are read with tone_ticket_phrases until the set is reused by a later
 * submit.  A set must be waited on before it is submitted again (TONE_ESTATE otherwise). */
int tone_submit(tone_engine* e, int32_t B, const int32_t* slots, const void* pcm, int32_t pcm_format,
                const uint8_t* is_last, int32_t outputs, int32_t* ticket_out);
int tone_wait(tone_engine* e, int32_t ticket, float* logprobs, int32_t* tokens, float* sil);

/* Pinned staging of the set the NEXT tone_submit will use: writing the inputs there (and passing these very
 * pointers to tone_submit with TONE_PCM_I16) skips the intermediate host copy.  PCM always crosses PCIe as int16
 * (int32 input is range-checked and narrowed while it is copied into this staging). */
int tone_next_staging(tone_engine* e, int32_t** slots, int16_t** pcm16, uint8_t** is_last);

/* Device-side phrase splitter + greedy CTC decoder (tone/logprob_splitter.py:60-153, tone/decoder.py:57-59,
 * the frame arithmetic of tone/pipeline.py:149-171 stays with the caller).  One record per finished phrase:
 * the frame interval [start_frame, end_frame) in the stream's own frame count, and the greedy text as label ids
 * (argmax -> repeats collapsed -> blank dropped -> leading/trailing ' ' stripped; ids index tone/decoder.py:23). */
typedef struct tone_phrase {
  int32_t batch_index;   /* position of the stream in the submitted batch                         */
  int32_t start_frame;   /* LogprobPhrase.start_frame (tone/logprob_splitter.py:139)              */
  int32_t end_frame;     /* LogprobPhrase.end_frame                                               */
  int32_t text_offset;   /* first label id of this phrase in the ticket's text pool               */
  int32_t text_len;      /* number of label ids                                                   */
} tone_phrase;
/* After tone_wait(ticket): the ticket's phrase records (in (batch_index, time) order per stream; streams in
 * arbitrary order) and its text pool; the pointers stay valid until the staging set is submitted again. */
int tone_ticket_phrases(tone_engine* e, int32_t ticket, const tone_phrase** phrases, int32_t* n_phrases,
                        const uint8_t** text_pool, int32_t* text_pool_len);

/* Feature-input form of the step: replaces the exported graph built with `--skip-preprocessor`
 * (tone/nn/model.py:151-160, tone/scripts/export.py:48-49), i.e. the acoustic model behind an external log-mel front
 * end such as the Triton/DALI ensemble (triton/preprocessing/1/features_8k_tone.py).
 *   feats  host fp16 [B][64][chunk_samples / 80]  log-mel features, (B, C = n_mels, T) like the reference
 * The waveform state of the slots is left untouched; everything else is tone_step. */
int tone_step_features(tone_engine* e, int32_t B, const int32_t* slots, const uint16_t* feats_fp16,
                       float* logprobs, int32_t* tokens);

/* Same step with inputs/outputs left in HBM: stage once, then step any number of times on
 * the staged chunk.  tone_fetch / tone_fetch_greedy copy the last step's outputs to the host (they wait for
 * the last launch, whichever stream it went to). */
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
 * GPU tensors between ensemble stages, triton/ensemble/config.pbtxt:20-55).
 *   slots       HOST int32 [B] (validated like tone_step; copied to the device by the call)
 *   d_pcm       device [B][chunk_samples] in pcm_format; NULL = step again on what is staged
 *   d_logprobs  device fp32 [B][T][35] or NULL;  d_tokens device int32 [B][T] or NULL
 * Enqueued on `cuda_stream` (NULL = the engine's stream); asynchronous; see the ordering contract above. */
int tone_step_device(tone_engine* e, int32_t B, const int32_t* slots, const void* d_pcm, int32_t pcm_format,
                     float* d_logprobs, int32_t* d_tokens, void* cuda_stream);

/* State wire formats.  Used for parity, checkpoint/resume and stream migration; batched (n slots per call,
 * one device gather/scatter kernel and one copy).
 * Flat: the reference's fp16 vector ("state"/"state_next", configs/streaming_acoustic/config.pbtxt:12-33;
 * element order = get_initial_state order, tone/nn/model.py:259-267). */
int tone_export_states(tone_engine* e, int32_t n, const int32_t* slots, uint16_t* fp16_out /* [n][219729] */);
int tone_import_states(tone_engine* e, int32_t n, const int32_t* slots, const uint16_t* fp16_in /* [n][219729] */);
/* Three-tensor Triton cache layout of the newer export (tone/scripts/export.py:293-376,
 * triton/model/config.pbtxt:44-66): cache_last_time fp16 [n][18][384][30], cache_last_channel fp16 [n][32][8][50],
 * cache_last_chan_len int64 [n]. */
int tone_export_states_triton(tone_engine* e, int32_t n, const int32_t* slots, uint16_t* cache_last_time,
                              uint16_t* cache_last_channel, int64_t* cache_last_chan_len);
int tone_import_states_triton(tone_engine* e, int32_t n, const int32_t* slots, const uint16_t* cache_last_time,
                              const uint16_t* cache_last_channel, const int64_t* cache_last_chan_len);

/* ---------------------------------------------------------------------------------------------------------------
 * Stream server: sequence batcher + stepping thread above one engine.  The reference delegates this role to Triton -
 * `sequence_batching { oldest { max_candidate_sequences 4096 } max_sequence_idle_microseconds 15 s }`
 * (triton/model/config.pbtxt:26-31) and `dynamic_batching { max_queue_delay_microseconds 10000 }`
 * (configs/streaming_acoustic/config.pbtxt:35-37).  Here: any number of producer threads push chunks; ONE native
 * thread owns the engine, forms batches (at most one chunk per stream per step, oldest head-of-queue first, a step is
 * due when max_batch chunks wait or the oldest has waited max_queue_delay), keeps two tickets in flight
 * (tone_submit / tone_wait), allocates a slot on a stream's first chunk, releases it after its last chunk or when it
 * has been idle for idle_timeout, and hands completed batches to tone_server_poll.  All entry points are thread-safe.
 * While a server exists, the engine must not be used directly. */
typedef struct tone_server tone_server;
typedef struct tone_server_config {
  int32_t max_batch;           /* largest step (<= engine max_batch); 0 = engine max_batch                         */
  int32_t max_queue_delay_us;  /* batching window; 0 = 10000 (the reference's Triton setting)                      */
  int32_t idle_timeout_ms;     /* idle streams are closed after this long; 0 = 15000                               */
  int32_t queue_depth;         /* chunks buffered per stream; 0 = 4                                                */
  int32_t outputs;             /* TONE_OUT_LOGPROBS and / or TONE_OUT_PHRASES delivered per chunk; 0 = PHRASES     */
  int32_t prewarm;             /* 1 = capture the step graph of every batch-size bucket at create (about a second) */
} tone_server_config;
typedef struct tone_server_stats {
  int64_t steps, chunks, phrases, streams_opened, streams_closed, streams_reclaimed, rejected;
  int32_t open_streams, queued_chunks;
  double mean_batch;
  double latency_ms_p50, latency_ms_p99, latency_ms_max;   /* push -> results available, per chunk */
  double queue_ms_p50, queue_ms_p99;                        /* push -> batch formed                 */
} tone_server_stats;
/* A finished phrase of a stream (see tone_phrase); text ids are in the batch's text pool. */
typedef struct tone_stream_phrase {
  uint64_t stream_id;
  int32_t start_frame, end_frame, text_offset, text_len;
} tone_stream_phrase;

int tone_server_create(tone_engine* e, const tone_server_config* cfg, tone_server** out);
void tone_server_destroy(tone_server* s);
/* Push n chunks, one per listed stream (a stream may appear once per call): pcm int16 [n][chunk_samples];
 * flags[i] bit 0 = this is the stream's last chunk (nullable = none).  A stream is opened by its first chunk.
 * All or nothing: TONE_ENOMEM if a stream's queue is full or no slot is left for a new stream. */
int tone_server_push(tone_server* s, int32_t n, const uint64_t* stream_ids, const int16_t* pcm, const uint8_t* flags);
/* Take the next completed batch (steps complete in order).  Waits up to timeout_ms for one; *n_chunks = 0 on time-out.
 *   stream_ids [max_batch], seq [max_batch] (chunk number within its stream), latency_ms [max_batch]
 *   logprobs   [max_batch][T][35] or NULL (needs TONE_OUT_LOGPROBS)
 *   phrases    [phrase_cap] or NULL, text [text_cap] or NULL (need TONE_OUT_PHRASES): finished phrases of the batch */
int tone_server_poll(tone_server* s, int32_t timeout_ms, int32_t* n_chunks, uint64_t* stream_ids, int32_t* seq,
                     float* latency_ms, float* logprobs, tone_stream_phrase* phrases, int32_t phrase_cap,
                     int32_t* n_phrases, uint8_t* text, int32_t text_cap, int32_t* text_len);
int tone_server_get_stats(tone_server* s, tone_server_stats* out);

/* Debug: run one eager step that also records the residual stream after pre-encode and after
 * every Conformer layer.  taps: host fp32 [1+n_layers][B*frames_out][384] (rows of reduced
 * layers 7..14 occupy the first B*T2 rows).  Not a product path. */
int tone_step_debug(tone_engine* e, int32_t B, const int32_t* slots, const int32_t* pcm,
                    float* logprobs, int32_t* tokens, float* taps);

/* Debug: C[M][N] = A[M][K] * W[N][K]^T through the product GEMM kernel (bf16 operands given as
 * fp32, rounded on upload; fp32 result).  Used by the GPU unit tests. */
int tone_selftest_gemm(tone_engine* e, int32_t M, int32_t N, int32_t K, const float* A, const float* W,
                       float* C, int32_t block_n);

/* Debug: run the device-side phrase splitter + greedy decoder alone on host-supplied per-frame inputs (one stream
 * per slot, `frames` frames each): tokens int32 [B][frames], sil fp32 [B][frames][2].  Lets the splitter be
 * checked against the reference's StreamingLogprobSplitter + GreedyCTCDecoder on arbitrary log-prob streams.
 * Results are read with tone_ticket_phrases(e, -1, ...). */
int tone_selftest_phrases(tone_engine* e, int32_t B, const int32_t* slots, int32_t frames, const int32_t* tokens,
                          const float* sil, const uint8_t* is_last);

#ifdef __cplusplus
}
#endif
#endif /* TONE_B200_H */
