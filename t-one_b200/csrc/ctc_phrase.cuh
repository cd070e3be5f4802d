// Device-side phrase splitter + greedy CTC decoder.
//
// Restates, as an incremental per-stream state machine, what the reference does on the host per chunk:
//   StreamingLogprobSplitter.forward      tone/logprob_splitter.py:91-153  (is_speech :129, trimming :144-151)
//   StreamingLogprobSplitter._iterate...  tone/logprob_splitter.py:60-89   (20-frame separator rule, 2000-frame forced split)
//   GreedyCTCDecoder.forward              tone/decoder.py:57-59            (argmax -> collapse repeats -> drop blank -> strip)
//
// The reference keeps, per stream, the log-probs since `offset` and re-derives the silence runs of that whole buffer on
// every call.  Equivalent incremental form (the buffer is always trimmed to start at the end of the last emitted phrase,
// so it never contains a completed separator): per stream we keep the frames since `offset` as one byte each
// (argmax token | speech bit << 7) in a ring, the index `s` of the first speech frame of the unfinished phrase (-1: none)
// and the length `run` of the silence run at the end of the buffer.  Per step, for each new frame:
//   speech:  starts the unfinished phrase if there is none; run = 0
//   silence: run += 1; when run reaches 20 inside a phrase, the phrase [s, n - 20) is complete (the leading pad of 20
//            silent frames the reference prepends makes whatever precedes the first speech frame a separator as well)
// then, once per step (the reference evaluates these on the buffer as it stands after the chunk was appended):
//   forced split  while n - s >= 2000: emit [s, s + 2000), continue at the first speech frame at or after s + 2000
//   is_last       the trailing pad of 20 silent frames completes the unfinished phrase at [s, n - run)
//   trimming      buffer := buffer[last emitted end:], or its last 3 frames if no speech is left in it
// A phrase's text is decoded from frames [max(0, start - 3), min(n, end + 3)) of the buffer at the time of the call.
#pragma once

#include "common.cuh"

namespace tone {

constexpr int PH_CAP = 4096;              // ring capacity per stream (frames); the buffer never exceeds ~2050
constexpr int PH_MIN_SILENCE = 20;        // tone/logprob_splitter.py:56
constexpr int PH_EXPAND = 3;              // :57
constexpr int PH_MAX_PHRASE = 2000;       // :58
constexpr int PH_BLANK = 34;              // ids >= len(LABELS) are dropped (tone/decoder.py:59)
constexpr int PH_SPACE = 33;              // LABELS[33] == ' ' (tone/decoder.py:23): what str.strip() removes
constexpr int PH_PER_STREAM = 4;          // most phrases one stream can finish in one step

struct PhSlot {                           // per stream slot, persistent
  int head;                               // ring position of buffer[0]
  int n;                                  // frames in the buffer
  int s;                                  // first speech frame of the unfinished phrase (buffer coordinates), -1 = none
  int run;                                // length of the silence run at the end of the buffer
  int offset;                             // stream frame index of buffer[0]  (StreamingLogprobSplitterState.offset)
  int pad[3];
};

struct PhHeader {                         // per step output header
  int n_phrases;
  int pool_used;
  int overflow;                           // records or text pool exhausted (never with the capacities the engine sizes)
  int pad;
};

struct PhRecord {                         // == tone_phrase (include/tone_b200.h)
  int batch_index, start_frame, end_frame, text_offset, text_len;
};

struct PhraseArgs {
  const int* slots;                       // [B]
  const int* tokens;                      // [B][T] per-frame argmax (decoder epilogue)
  const float* sil;                       // [B][T][2] log-prob of ' ' and of blank
  const unsigned char* is_last;           // [B] or null
  PhSlot* st;                             // [slots]
  unsigned char* ring;                    // [slots][PH_CAP]
  PhHeader* hdr;
  PhRecord* rec;                          // [max_rec]
  unsigned char* pool;                    // [pool_cap]
  int B, T, max_rec, pool_cap;
};

// speech iff exp(lp[' ']) + exp(lp[blank]) <= 0.9 in float32 (tone/logprob_splitter.py:129; numpy compares the float32
// sum against the float32 value of 0.9).  exp is evaluated in double and rounded once, i.e. correctly rounded float32.
__device__ __forceinline__ bool ph_is_speech(float lp_space, float lp_blank) {
  const float a = (float)exp((double)lp_space), b = (float)exp((double)lp_blank);
  return (a + b) <= 0.9f;
}

// One CTA for the whole batch (thread = stream, strided): the per-step work is T frames per stream, and a single CTA
// lets the record / text-pool allocators live in shared memory (no memset node, no global atomics).
constexpr int PH_THREADS = 256;

__global__ void __launch_bounds__(PH_THREADS) ctc_phrase_kernel(const PhraseArgs a) {
  __shared__ int s_nrec, s_pool, s_over;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    s_nrec = 0;
    s_pool = 0;
    s_over = 0;
  }
  pdl_wait();
  __syncthreads();
  for (int b = threadIdx.x; b < a.B; b += PH_THREADS) {
    const int slot = a.slots[b];
    PhSlot st = a.st[slot];
    unsigned char* ring = a.ring + (size_t)slot * PH_CAP;
    int last_end = 0;                                       // `last_phrase` of logprob_splitter.py:134-142

    auto emit = [&](int start, int end) {                   // buffer coordinates; st.n = buffer length of this call
      const int lo = max(0, start - PH_EXPAND), hi = min(st.n, end + PH_EXPAND);
      const int idx = atomicAdd(&s_nrec, 1);
      const int off = atomicAdd(&s_pool, hi - lo);          // worst case: nothing collapses
      last_end = end;
      if (idx >= a.max_rec || off + (hi - lo) > a.pool_cap) {
        s_over = 1;
        return;
      }
      int len = 0, prev = -1;
      for (int i = lo; i < hi; ++i) {
        const int t = ring[(st.head + i) & (PH_CAP - 1)] & 0x7f;
        if (t != prev && t < PH_BLANK && !(len == 0 && t == PH_SPACE)) a.pool[off + len++] = (unsigned char)t;
        prev = t;
      }
      while (len > 0 && a.pool[off + len - 1] == PH_SPACE) --len;
      PhRecord r;
      r.batch_index = b;
      r.start_frame = st.offset + start;
      r.end_frame = st.offset + end;
      r.text_offset = off;
      r.text_len = len;
      a.rec[idx] = r;
    };
    // a completed phrase [s, e): forced 2000-frame pieces first (logprob_splitter.py:85-87), then the rest
    auto emit_completed = [&](int s, int e) {
      while (e - s >= PH_MAX_PHRASE) {
        emit(s, s + PH_MAX_PHRASE);
        s += PH_MAX_PHRASE;
      }
      emit(s, e);
    };

    const int n_after = st.n + a.T;                         // every emit of this call sees the whole appended buffer
    const int n_before = st.n;
    // append the chunk first (the text of a forced piece may reach into frames appended after its end)
    bool sp[16];
    for (int t = 0; t < a.T; ++t) {
      const int tok = a.tokens[(size_t)b * a.T + t];
      const float2 sl = *reinterpret_cast<const float2*>(a.sil + ((size_t)b * a.T + t) * 2);
      sp[t] = ph_is_speech(sl.x, sl.y);
      ring[(st.head + n_before + t) & (PH_CAP - 1)] = (unsigned char)(tok | (sp[t] ? 0x80 : 0));
    }
    st.n = n_after;
    for (int t = 0; t < a.T; ++t) {
      const int i = n_before + t;
      if (sp[t]) {
        if (st.s < 0) st.s = i;
        st.run = 0;
      } else {
        ++st.run;
        if (st.s >= 0 && st.run == PH_MIN_SILENCE) {        // separator complete: phrase [s, start of this run)
          emit_completed(st.s, i + 1 - PH_MIN_SILENCE);
          st.s = -1;
        }
      }
    }
    const bool fin = a.is_last && a.is_last[b];
    if (st.s >= 0) {
      if (fin) {                                            // trailing pad of 20 silent frames closes the phrase
        emit_completed(st.s, st.n - st.run);
        st.s = -1;
      } else {
        bool split = false;
        while (st.n - st.s >= PH_MAX_PHRASE) {              // unfinished phrase: its end is the buffer end (:83)
          emit(st.s, st.s + PH_MAX_PHRASE);
          st.s += PH_MAX_PHRASE;
          split = true;
        }
        if (split) {                                        // next call starts from the trimmed buffer: the phrase
          int i = st.s;                                     // resumes at its first speech frame
          while (i < st.n && !(ring[(st.head + i) & (PH_CAP - 1)] & 0x80)) ++i;
          st.s = i < st.n ? i : -1;
        }
      }
    }
    // trimming (logprob_splitter.py:144-151): no speech left after the last phrase -> keep the last 3 frames only
    int cut = last_end;
    if (st.s < 0) cut = max(cut, st.n - PH_EXPAND);
    cut = max(cut, 0);
    if (cut > 0) {
      st.head = (st.head + cut) & (PH_CAP - 1);
      st.n -= cut;
      st.offset += cut;
      if (st.s >= 0) st.s -= cut;
    }
    if (st.s < 0) st.run = min(st.run, st.n);               // only meaningful inside a phrase; keep it bounded
    a.st[slot] = st;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    PhHeader h;
    h.n_phrases = min(s_nrec, a.max_rec);
    h.pool_used = min(s_pool, a.pool_cap);
    h.overflow = s_over;
    h.pad = 0;
    *a.hdr = h;
  }
}

}  // namespace tone
