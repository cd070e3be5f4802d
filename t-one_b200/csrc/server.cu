// tone_server: sequence batcher + stepping thread above one tone_engine (C ABI in include/tone_b200.h).
//
// The reference has no scheduler of its own; it delegates batching and per-stream state residency to Triton
// (sequence_batching / oldest, triton/model/config.pbtxt:26-31; dynamic_batching with a 10 ms queue delay,
// configs/streaming_acoustic/config.pbtxt:35-37).  This is that role, native and in process: producers push chunks from
// any thread; one worker thread owns the engine, forms batches and keeps two tickets in flight so that batch formation,
// the H2D / D2H copies and the kernels of neighbouring steps overlap.  Built only on the public C ABI of the engine.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/tone_b200.h"

namespace {

using Clock = std::chrono::steady_clock;
static inline double ms_between(Clock::time_point a, Clock::time_point b) {
  return std::chrono::duration<double, std::milli>(b - a).count();
}


struct Stream {
  uint64_t id = 0;
  int32_t slot = -1;           // -1 until the worker schedules its first chunk
  int32_t head = 0, count = 0; // ring of queued chunks (indices into the chunk pool)
  int32_t seq = 0;             // chunks stepped so far
  bool ending = false;         // the last chunk has been pushed
  uint64_t push_mark = 0;      // generation of the last tone_server_push that named this stream (duplicate detection)
  int32_t in_flight = 0;       // tickets in flight that contain a chunk of this stream
  Clock::time_point last_active;
  std::vector<int32_t> ring;   // [queue_depth] chunk-pool indices
  std::vector<Clock::time_point> t_push;
  std::vector<uint8_t> last_flag;
};

struct Batch {                  // one step, from formation to delivery
  int32_t ticket = -1, B = 0;   // B = real chunks (the submitted step may be padded beyond it)
  std::vector<uint64_t> ids;
  std::vector<int32_t> seq, chunk_idx;
  std::vector<Stream*> streams;
  std::vector<uint8_t> last;
  std::vector<Clock::time_point> t_push;
  Clock::time_point t_formed;
  std::vector<float> latency_ms;
  std::vector<float> logprobs;
  std::vector<tone_stream_phrase> phrases;
  std::vector<uint8_t> text;
};

}  // namespace

struct tone_server {
  tone_engine* eng = nullptr;
  tone_server_config cfg{};
  tone_info info{};
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::unordered_map<uint64_t, std::unique_ptr<Stream>> streams;
  std::vector<int16_t> pool;             // [n_chunks][chunk_samples]
  std::vector<int32_t> pool_free;
  std::deque<std::unique_ptr<Batch>> done;
  std::thread worker;
  bool stop = false;
  int32_t queued = 0;                    // chunks waiting
  // streams with at least one queued chunk, keyed by the push time of their head-of-queue chunk (oldest first);
  // a ready stream has exactly one entry
  typedef std::pair<Clock::time_point, Stream*> Ready;
  std::priority_queue<Ready, std::vector<Ready>, std::greater<Ready>> ready;
  // stats
  tone_server_stats st{};
  double batch_sum = 0;
  std::vector<float> lat_samples, queue_samples;
  std::string error;                     // first engine failure seen by the worker
  uint64_t push_gen = 0;
  std::unordered_set<uint64_t> fresh_ids;
  // Batch-size bucketing: every distinct batch size costs a CUDA-graph capture (milliseconds) at first use, so steps
  // are padded up to a small set of sizes with scratch streams the server owns (zero PCM, results dropped).
  std::vector<int32_t> pad_slots;
};

extern "C" void tone_internal_set_error(const char* msg);   // engine.cu: the buffer behind tone_last_error()
static int sfail(int code, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  tone_internal_set_error(buf);
  return code;
}

// sizes a step may have: exact up to 16, then multiples of 16 up to 128, then multiples of 64
static int bucket_size(int B) {
  if (B <= 16) return B;
  if (B <= 128) return (B + 15) / 16 * 16;
  return (B + 63) / 64 * 64;
}
static int padded_size(const tone_server* s, int B) {
  return std::min(std::min(bucket_size(B), B + (int)s->pad_slots.size()), (int)s->cfg.max_batch);
}

static void release_stream(tone_server* s, Stream* st, bool reclaimed) {   // mutex held; the stream is not in flight
  if (st->slot >= 0) tone_release_slots(s->eng, 1, &st->slot);
  for (int i = 0; i < st->count; ++i) s->pool_free.push_back(st->ring[(st->head + i) % (int)st->ring.size()]);
  s->queued -= st->count;
  if (reclaimed) s->st.streams_reclaimed++;
  else s->st.streams_closed++;
  s->streams.erase(st->id);
}

// Collect a completed ticket: outputs -> Batch, stats, stream bookkeeping, delivery.
static void finish_batch(tone_server* s, std::unique_ptr<Batch> b) {
  const int T = s->info.frames_out;
  const bool want_lp = s->cfg.outputs & TONE_OUT_LOGPROBS, want_ph = s->cfg.outputs & TONE_OUT_PHRASES;
  const int Bp = padded_size(s, b->B);   // as submitted
  if (want_lp) b->logprobs.resize((size_t)Bp * T * TONE_N_CLASSES);
  int rc = tone_wait(s->eng, b->ticket, want_lp ? b->logprobs.data() : nullptr, nullptr, nullptr);
  if (want_lp) b->logprobs.resize((size_t)b->B * T * TONE_N_CLASSES);
  if (!rc && want_ph) {
    const tone_phrase* ph = nullptr;
    const uint8_t* pool = nullptr;
    int32_t n = 0, npool = 0;
    rc = tone_ticket_phrases(s->eng, b->ticket, &ph, &n, &pool, &npool);
    if (!rc) {
      for (int i = 0; i < n; ++i)
        if (ph[i].batch_index < b->B)          // padding streams are not anybody's stream
          b->phrases.push_back(tone_stream_phrase{b->ids[ph[i].batch_index], ph[i].start_frame, ph[i].end_frame, ph[i].text_offset, ph[i].text_len});
      b->text.assign(pool, pool + npool);
    }
  }
  const auto now = Clock::now();
  std::lock_guard<std::mutex> lk(s->mu);
  if (rc && s->error.empty()) s->error = tone_last_error();
  b->latency_ms.resize(b->B);
  for (int i = 0; i < b->B; ++i) {
    b->latency_ms[i] = (float)ms_between(b->t_push[i], now);
    if (s->lat_samples.size() < (1u << 20)) {
      s->lat_samples.push_back(b->latency_ms[i]);
      s->queue_samples.push_back((float)ms_between(b->t_push[i], b->t_formed));
    }
    s->pool_free.push_back(b->chunk_idx[i]);
    Stream* st = b->streams[i];
    st->in_flight--;
    st->last_active = now;
    if (b->last[i]) release_stream(s, st, false);      // its last chunk has been stepped (and its phrase flushed)
  }
  s->st.steps++;
  s->st.chunks += b->B;
  s->st.phrases += (int64_t)b->phrases.size();
  s->batch_sum += b->B;
  s->done.push_back(std::move(b));
  s->cv_done.notify_all();
}

static void worker_main(tone_server* s) {
  const int C = s->info.chunk_samples;
  const auto window = std::chrono::microseconds(s->cfg.max_queue_delay_us);
  const auto idle = std::chrono::milliseconds(s->cfg.idle_timeout_ms);
  std::unique_ptr<Batch> pending;
  std::vector<tone_server::Ready> again;
  std::vector<Stream*> fresh;
  std::vector<int32_t> fresh_slots;
  auto last_reclaim = Clock::now();
  bool stopping = false;
  while (!stopping) {
    std::unique_ptr<Batch> b;
    {
      std::unique_lock<std::mutex> lk(s->mu);
      for (;;) {
        if (s->stop) {
          stopping = true;
          break;
        }
        const auto now = Clock::now();
        if (now - last_reclaim > std::chrono::milliseconds(200)) {    // idle reclaim (Triton: max_sequence_idle)
          last_reclaim = now;
          std::vector<Stream*> dead;
          for (auto& kv : s->streams)
            if (kv.second->count == 0 && kv.second->in_flight == 0 && now - kv.second->last_active >= idle) dead.push_back(kv.second.get());
          for (Stream* d : dead) release_stream(s, d, true);
        }
        // a step is due when a full batch waits or the oldest head-of-queue chunk has waited the window
        const bool due = !s->ready.empty() && ((int)s->ready.size() >= s->cfg.max_batch || now - s->ready.top().first >= window);
        if (due) {
          // oldest head-of-queue first, at most one chunk per stream (the step is a recurrence per stream)
          b.reset(new Batch());
          b->t_formed = now;
          again.clear();
          while (!s->ready.empty() && (int)b->ids.size() < s->cfg.max_batch) {
            Stream* st = s->ready.top().second;
            s->ready.pop();
            const int h = st->head;
            b->ids.push_back(st->id);
            b->seq.push_back(st->seq++);
            b->chunk_idx.push_back(st->ring[h]);
            b->t_push.push_back(st->t_push[h]);
            b->last.push_back(st->last_flag[h]);
            b->streams.push_back(st);
            st->head = (h + 1) % (int)st->ring.size();
            st->count--;
            st->in_flight++;
            s->queued--;
            if (st->count > 0) again.emplace_back(st->t_push[st->head], st);   // back in line with its next chunk, after this batch
          }
          for (auto& r : again) s->ready.push(r);
          b->B = (int)b->ids.size();
          // first chunk of a stream: its slot.  One call (one reset kernel) for all new streams of the batch; push()
          // caps the open streams at the slot count, so this cannot run out.
          fresh.clear();
          for (Stream* st : b->streams)
            if (st->slot < 0) fresh.push_back(st);
          if (!fresh.empty()) {
            fresh_slots.resize(fresh.size());
            if (tone_alloc_slots(s->eng, (int32_t)fresh.size(), fresh_slots.data()) == TONE_OK) {
              for (size_t i = 0; i < fresh.size(); ++i) fresh[i]->slot = fresh_slots[i];
            } else if (s->error.empty()) {
              s->error = tone_last_error();
            }
          }
          break;
        }
        if (pending) break;                                           // use the wait to collect the ticket in flight
        if (!s->ready.empty()) s->cv_work.wait_until(lk, s->ready.top().first + window);
        else s->cv_work.wait_for(lk, std::chrono::milliseconds(50));
      }
    }
    if (b && b->B > 0) {
      // fill the pinned staging of the next ticket outside the lock, then submit (asynchronous)
      int32_t* sl = nullptr;
      int16_t* pcm = nullptr;
      uint8_t* last = nullptr;
      tone_next_staging(s->eng, &sl, &pcm, &last);
      const int Bp = padded_size(s, b->B);
      auto fill = [&]() {
        for (int i = 0; i < b->B; ++i) {
          sl[i] = b->streams[i]->slot;
          last[i] = b->last[i];
          memcpy(pcm + (size_t)i * C, s->pool.data() + (size_t)b->chunk_idx[i] * C, (size_t)C * 2);
        }
        for (int i = b->B; i < Bp; ++i) {      // padding: scratch streams, silence
          sl[i] = s->pad_slots[i - b->B];
          last[i] = 0;
          memset(pcm + (size_t)i * C, 0, (size_t)C * 2);
        }
      };
      fill();
      int rc = tone_submit(s->eng, Bp, sl, pcm, TONE_PCM_I16, last, s->cfg.outputs, &b->ticket);
      if (rc == TONE_ESTATE && pending) {      // both staging sets busy: collect the older ticket first
        finish_batch(s, std::move(pending));
        tone_next_staging(s->eng, &sl, &pcm, &last);
        fill();
        rc = tone_submit(s->eng, Bp, sl, pcm, TONE_PCM_I16, last, s->cfg.outputs, &b->ticket);
      }
      if (rc) {
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->error.empty()) s->error = tone_last_error();
        for (int i = 0; i < b->B; ++i) {
          b->streams[i]->in_flight--;
          s->pool_free.push_back(b->chunk_idx[i]);
        }
        s->st.rejected += b->B;
        b.reset();
      }
    }
    if (pending) finish_batch(s, std::move(pending));
    if (b && b->B > 0) pending = std::move(b);
  }
  if (pending) finish_batch(s, std::move(pending));
}

extern "C" int tone_server_create(tone_engine* e, const tone_server_config* cfg, tone_server** out) {
  if (!e || !cfg || !out) return sfail(TONE_EINVAL, "null argument");
  tone_info info;
  int rc = tone_get_info(e, &info);
  if (rc) return rc;
  std::unique_ptr<tone_server> s(new tone_server());
  s->eng = e;
  s->info = info;
  s->cfg = *cfg;
  if (s->cfg.max_batch <= 0 || s->cfg.max_batch > info.max_batch) s->cfg.max_batch = info.max_batch;
  if (s->cfg.max_queue_delay_us <= 0) s->cfg.max_queue_delay_us = 10000;
  if (s->cfg.idle_timeout_ms <= 0) s->cfg.idle_timeout_ms = 15000;
  if (s->cfg.queue_depth <= 0) s->cfg.queue_depth = 4;
  if (s->cfg.outputs == 0) s->cfg.outputs = TONE_OUT_PHRASES;
  if (s->cfg.outputs & ~(TONE_OUT_LOGPROBS | TONE_OUT_PHRASES)) return sfail(TONE_EINVAL, "server outputs: LOGPROBS and / or PHRASES");
  const size_t n_chunks = (size_t)info.max_slots * s->cfg.queue_depth;
  s->pool.resize(n_chunks * info.chunk_samples);
  s->pool_free.reserve(n_chunks);
  for (size_t i = n_chunks; i-- > 0;) s->pool_free.push_back((int32_t)i);
  const int n_pad = std::min(63, info.max_slots / 8);
  if (n_pad > 0) {
    s->pad_slots.resize(n_pad);
    rc = tone_alloc_slots(e, n_pad, s->pad_slots.data());
    if (rc) return rc;
  }
  if (cfg->prewarm) {
    // One silent step per batch-size bucket on scratch slots, so that no CUDA graph is captured while serving (a
    // capture costs tens of milliseconds: it would be the p99 of the first seconds).
    const int nb = std::min(s->cfg.max_batch, info.max_slots - n_pad);
    std::vector<int32_t> tmp(nb);
    if ((rc = tone_alloc_slots(e, nb, tmp.data()))) return rc;
    int last_size = 0;
    for (int b = 1; b <= nb && !rc; ++b) {
      const int size = std::min(bucket_size(b), nb);
      if (size == last_size) continue;
      last_size = size;
      int32_t* sl = nullptr;
      int16_t* pcm = nullptr;
      uint8_t* last = nullptr;
      tone_next_staging(e, &sl, &pcm, &last);
      memcpy(sl, tmp.data(), (size_t)size * 4);
      memset(pcm, 0, (size_t)size * info.chunk_samples * 2);
      memset(last, 0, size);
      int32_t ticket = -1;
      rc = tone_submit(e, size, sl, pcm, TONE_PCM_I16, last, s->cfg.outputs, &ticket);
      if (!rc) rc = tone_wait(e, ticket, nullptr, nullptr, nullptr);
    }
    tone_release_slots(e, nb, tmp.data());
    if (rc) {
      tone_release_slots(e, n_pad, s->pad_slots.data());
      return rc;
    }
  }
  s->worker = std::thread(worker_main, s.get());
  *out = s.release();
  return TONE_OK;
}

extern "C" void tone_server_destroy(tone_server* s) {
  if (!s) return;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->stop = true;
  }
  s->cv_work.notify_all();
  if (s->worker.joinable()) s->worker.join();
  for (auto& kv : s->streams)
    if (kv.second->slot >= 0) tone_release_slots(s->eng, 1, &kv.second->slot);
  if (!s->pad_slots.empty()) tone_release_slots(s->eng, (int32_t)s->pad_slots.size(), s->pad_slots.data());
  delete s;
}

extern "C" int tone_server_push(tone_server* s, int32_t n, const uint64_t* ids, const int16_t* pcm, const uint8_t* flags) {
  if (!s || !ids || !pcm || n < 0) return sfail(TONE_EINVAL, "bad argument");
  const int C = s->info.chunk_samples, D = s->cfg.queue_depth;
  const auto now = Clock::now();
  std::lock_guard<std::mutex> lk(s->mu);
  if (!s->error.empty()) return sfail(TONE_ECUDA, "server stopped: %s", s->error.c_str());
  // all or nothing: check duplicates and capacity first
  int fresh = 0;
  ++s->push_gen;
  s->fresh_ids.clear();
  for (int i = 0; i < n; ++i) {
    auto it = s->streams.find(ids[i]);
    if (it == s->streams.end()) {
      if (!s->fresh_ids.insert(ids[i]).second) return sfail(TONE_EINVAL, "stream %llu appears twice in one push", (unsigned long long)ids[i]);
      ++fresh;
      continue;
    }
    Stream* st = it->second.get();
    if (st->push_mark == s->push_gen) return sfail(TONE_EINVAL, "stream %llu appears twice in one push", (unsigned long long)ids[i]);
    st->push_mark = s->push_gen;
    if (st->count >= D) return sfail(TONE_ENOMEM, "stream %llu has %d chunks queued already", (unsigned long long)ids[i], D);
    if (st->ending) return sfail(TONE_ESTATE, "stream %llu already received its last chunk", (unsigned long long)ids[i]);
  }
  if ((int64_t)s->streams.size() + fresh > s->info.max_slots - (int64_t)s->pad_slots.size()) {
    s->st.rejected += n;
    return sfail(TONE_ENOMEM, "%d new streams, %zu open of %d slots", fresh, s->streams.size(), s->info.max_slots - (int)s->pad_slots.size());
  }
  if ((size_t)n > s->pool_free.size()) return sfail(TONE_ENOMEM, "chunk pool exhausted");
  for (int i = 0; i < n; ++i) {
    auto& up = s->streams[ids[i]];
    if (!up) {
      up.reset(new Stream());
      up->id = ids[i];
      up->ring.resize(D);
      up->t_push.resize(D);
      up->last_flag.resize(D);
      s->st.streams_opened++;
    }
    Stream* st = up.get();
    const int32_t ci = s->pool_free.back();
    s->pool_free.pop_back();
    memcpy(s->pool.data() + (size_t)ci * C, pcm + (size_t)i * C, (size_t)C * 2);
    const int pos = (st->head + st->count) % D;
    st->ring[pos] = ci;
    st->t_push[pos] = now;
    st->last_flag[pos] = flags ? (flags[i] & 1) : 0;
    if (st->last_flag[pos]) st->ending = true;
    if (st->count == 0) s->ready.emplace(now, st);
    st->count++;
    st->last_active = now;
    s->queued++;
  }
  s->cv_work.notify_one();
  return TONE_OK;
}

extern "C" int tone_server_poll(tone_server* s, int32_t timeout_ms, int32_t* n_chunks, uint64_t* ids, int32_t* seq,
                                float* latency_ms, float* logprobs, tone_stream_phrase* phrases, int32_t phrase_cap,
                                int32_t* n_phrases, uint8_t* text, int32_t text_cap, int32_t* text_len) {
  if (!s || !n_chunks) return sfail(TONE_EINVAL, "bad argument");
  std::unique_ptr<Batch> b;
  {
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->done.empty() && timeout_ms > 0)
      s->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !s->done.empty() || !s->error.empty(); });
    if (s->done.empty()) {
      *n_chunks = 0;
      if (n_phrases) *n_phrases = 0;
      if (text_len) *text_len = 0;
      if (!s->error.empty()) return sfail(TONE_ECUDA, "server stopped: %s", s->error.c_str());
      return TONE_OK;
    }
    b = std::move(s->done.front());
    s->done.pop_front();
  }
  *n_chunks = b->B;
  if (ids) memcpy(ids, b->ids.data(), (size_t)b->B * 8);
  if (seq) memcpy(seq, b->seq.data(), (size_t)b->B * 4);
  if (latency_ms) memcpy(latency_ms, b->latency_ms.data(), (size_t)b->B * 4);
  if (logprobs) {
    if (b->logprobs.empty()) return sfail(TONE_EINVAL, "log-probs were not requested at tone_server_create");
    memcpy(logprobs, b->logprobs.data(), b->logprobs.size() * 4);
  }
  const int np = (int)b->phrases.size();
  if (n_phrases) *n_phrases = np;
  if (text_len) *text_len = (int)b->text.size();
  if (phrases && np) {
    if (np > phrase_cap || (int)b->text.size() > text_cap || !text) return sfail(TONE_ENOMEM, "phrase buffers too small (%d phrases, %zu text bytes)", np, b->text.size());
    memcpy(phrases, b->phrases.data(), (size_t)np * sizeof(tone_stream_phrase));
    memcpy(text, b->text.data(), b->text.size());
  }
  return TONE_OK;
}

static void percentiles(std::vector<float> v, double* p50, double* p99, double* mx) {
  if (v.empty()) {
    *p50 = *p99 = 0;
    if (mx) *mx = 0;
    return;
  }
  std::sort(v.begin(), v.end());
  *p50 = v[v.size() / 2];
  *p99 = v[std::min(v.size() - 1, (size_t)(v.size() * 0.99))];
  if (mx) *mx = v.back();
}

extern "C" int tone_server_get_stats(tone_server* s, tone_server_stats* out) {
  if (!s || !out) return sfail(TONE_EINVAL, "bad argument");
  std::vector<float> lat, q;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->st;
    out->open_streams = (int32_t)s->streams.size();
    out->queued_chunks = s->queued;
    out->mean_batch = s->st.steps ? s->batch_sum / (double)s->st.steps : 0.0;
    lat = s->lat_samples;
    q = s->queue_samples;
  }
  percentiles(std::move(lat), &out->latency_ms_p50, &out->latency_ms_p99, &out->latency_ms_max);
  percentiles(std::move(q), &out->queue_ms_p50, &out->queue_ms_p99, nullptr);
  return TONE_OK;
}
