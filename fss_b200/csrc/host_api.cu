// SPDX-License-Identifier: Apache-2.0
//
// host_api.cu -- the host-buffer entry points of include/fssb200.h (`fssb200_*_host`, fssb200_pack_rows): what a
// CPU caller of the reference binds.  Inputs live in HOST memory in the reference's layouts; every call stages them
// to the device, launches the sm_100a kernels of api.cu and brings the results back.  No evaluation happens on the
// CPU: the host threads below only move and re-pack bytes.
//
// Re-entrancy (the reference's Eval is a const pure function used under `#pragma omp parallel for`,
// src/bench_cpu.cu:157-161): a context owns no staging memory.  Every call checks an ARENA (device staging +
// pinned staging + 3 streams + events) out of a process-wide per-device pool, sized from the call's own batch, and
// returns it on every exit path after both the streams and the worker threads are quiescent.  A 1-key call takes a
// 1 MiB arena; concurrent calls take different arenas.
//
// fssb200_eval_host, large batches: ADAPTIVE PACK / DIRECT PIPELINE.  15 of the 32 bytes of a Dpf::Cw /
// HalfTreeDpf::Cw are padding, and the call is bound by the PCIe link (or, with several GPUs, by whatever the
// host gives each link).  Worker threads strip the padding of chunk after chunk (from the front of the batch) into
// a ring of pinned staging slots; the calling thread submits every finished slot as ONE H2D copy + kernel + D2H, and
// whenever the link is about to run dry while no packed chunk is ready it sends a chunk from the BACK of the
// batch in the reference layout straight from the caller's buffer.  So the split between "CPU packs, link moves
// 17 B/level" and "link moves 32 B/level" settles by itself at the point where cores and link finish together,
// for any core count per rank (1 GPU with 16 cores: almost everything packed; 8 ranks with 4 cores each: about
// half).  No barriers: workers claim 512-key blocks with one fetch_add, hand-offs are per-chunk counters.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "ctx.h"

using namespace fssb200;

namespace {

// ---------------------------------------------------------------------------------------------------------------
// row packing (format conversion only)
// ---------------------------------------------------------------------------------------------------------------
// rows [k0, k1): ncw x {16 B s} + 16 B of flag bits (bit i = byte 16 of entry i != 0, i < 128).
// STREAM: non-temporal stores for the whole row, flag word included -- a regular store into a line that is still
// being assembled in a write-combining buffer forces a flush + read-for-ownership and cost 3.5x (measured).
template <bool STREAM>
void pack_rows_range_t(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw) {
  const size_t in_row = size_t(ncw) * 32u, out_row = size_t(ncw) * 16u + 16u;
  const int nflag = ncw < 128 ? ncw : 128;
  for (size_t k = k0; k < k1; ++k) {
    const uint8_t *r = src + k * in_row;
    uint8_t *o = dst + k * out_row;
    uint64_t f0 = 0, f1 = 0;
    for (int i = 0; i < nflag && i < 64; ++i) f0 |= uint64_t(r[32 * i + 16] != 0) << i;
    for (int i = 64; i < nflag; ++i) f1 |= uint64_t(r[32 * i + 16] != 0) << (i - 64);
#if defined(__SSE2__)
    for (int i = 0; i < ncw; ++i) {
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(r + 32 * i));
      if (STREAM) _mm_stream_si128(reinterpret_cast<__m128i *>(o + 16 * i), v);
      else _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 16 * i), v);
    }
    const __m128i fv = _mm_set_epi64x(static_cast<long long>(f1), static_cast<long long>(f0));
    if (STREAM) _mm_stream_si128(reinterpret_cast<__m128i *>(o + size_t(ncw) * 16u), fv);
    else _mm_storeu_si128(reinterpret_cast<__m128i *>(o + size_t(ncw) * 16u), fv);
#else
    for (int i = 0; i < ncw; ++i) std::memcpy(o + 16 * i, r + 32 * i, 16);
    const uint64_t f[2] = {f0, f1};
    std::memcpy(o + size_t(ncw) * 16u, f, 16);
#endif
  }
#if defined(__SSE2__)
  if (STREAM) _mm_sfence();
#endif
}
void pack_rows_range(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw, bool stream_stores) {
  if (stream_stores) pack_rows_range_t<true>(src, dst, k0, k1, ncw);
  else pack_rows_range_t<false>(src, dst, k0, k1, ncw);
}
// plain staging copy (16-byte granules; dst 16-byte aligned); non-temporal when `stream_stores`
void stage_copy(const uint8_t *src, uint8_t *dst, size_t bytes, bool stream_stores) {
#if defined(__SSE2__)
  if (stream_stores && !(bytes & 15u)) {
    for (size_t i = 0; i < bytes; i += 16)
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i)));
    _mm_sfence();
    return;
  }
#endif
  std::memcpy(dst, src, bytes);
}

inline void cpu_relax() {
#if defined(__SSE2__)
  _mm_pause();
#endif
}

int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e && *e ? std::atoi(e) : dflt;
}

int usable_cpus() {
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) return CPU_COUNT(&set);
#endif
  const unsigned h = std::thread::hardware_concurrency();
  return h ? int(h) : 1;
}

// Threads one call of this process may use to stage rows, the calling thread included: the usable cores divided by
// the ranks that share the host (torchrun's LOCAL_WORLD_SIZE), at most 32.  FSSB200_PACK_THREADS overrides (0 or 1: the
// reference layout always crosses the link as it is).
int host_thread_budget() {
  int lws = env_int("LOCAL_WORLD_SIZE", 1);
  if (lws < 1) lws = 1;
  int t = usable_cpus() / lws;
  // Several ranks on one host: leave two CPUs of the share to the rank's other busy threads (its submitter, the
  // framework's) -- measured with 2 ranks on 24 CPUs, barrier-synchronised steps: 12 threads per rank 75.4 ms, 10 threads
  // 68.0 ms, 8 69.2, 6 69.9, 4 74.3 (profiles/r02_host_pipeline.md).  One rank: every CPU helps (16: 46.7, 14: 46.6-48.7, 12: 51.7).
  if (lws >= 2 && t > 4) t -= 2;
  if (t > 32) t = 32;
  t = env_int("FSSB200_PACK_THREADS", t);
  return t < 0 ? 0 : t;
}

// Last-level cache of the host divided by the ranks that share it (0 if unknown).
size_t llc_bytes_per_rank() {
  static const size_t v = [] {
    size_t best = 0;
#if defined(__linux__)
    for (int idx = 2; idx <= 4; ++idx) {
      char path[96];
      std::snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu0/cache/index%d/size", idx);
      if (FILE *f = std::fopen(path, "r")) {
        unsigned long n = 0;
        char unit = 0;
        if (std::fscanf(f, "%lu%c", &n, &unit) >= 1) {
          size_t b = size_t(n) * (unit == 'K' ? 1024u : (unit == 'M' ? 1048576u : 1u));
          if (b > best) best = b;
        }
        std::fclose(f);
      }
    }
#endif
    int lws = env_int("LOCAL_WORLD_SIZE", 1);
    return best / size_t(lws < 1 ? 1 : lws);
  }();
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// worker threads: one process-wide crew, lent to one call at a time
// ---------------------------------------------------------------------------------------------------------------
struct CrewJob {
  virtual void run() = 0;  // called once by every worker lent to the job; returns when the job has nothing left for it
  virtual ~CrewJob() {}
};

// The crew can serve several calls at once (one process driving several GPUs: fssb200_eval_host_multi): a call
// borrows up to `want` idle workers and gets whatever is idle (possibly none).
class Crew {
 public:
  static Crew *get() {
    static Crew *crew = [] {
      const int t = host_thread_budget();
      return t >= 2 ? new (std::nothrow) Crew(t - 1) : nullptr;  // leaked on purpose: no teardown races at exit
    }();
    return crew;
  }
  int workers() const { return int(slot_.size()); }
  // Lends up to `want` idle workers to `job`; returns how many (0: none idle).  The caller must call end(job) before
  // `job` dies.
  int begin(CrewJob *job, int want) {
    std::lock_guard<std::mutex> l(mu_);
    int got = 0;
    for (auto &s : slot_) {
      if (got >= want) break;
      if (!s.job) {
        s.job = job;
        s.fresh = true;
        ++got;
      }
    }
    if (got) cv_start_.notify_all();
    return got;
  }
  // Blocks until every worker lent to `job` has left it.
  void end(CrewJob *job) {
    std::unique_lock<std::mutex> l(mu_);
    cv_done_.wait(l, [&] {
      for (auto &s : slot_)
        if (s.job == job) return false;
      return true;
    });
  }

 private:
  struct Slot {
    CrewJob *job = nullptr;
    bool fresh = false;
  };
  explicit Crew(int n) : slot_(size_t(n)) {
    for (int i = 0; i < n; ++i) std::thread([this, i] { loop(i); }).detach();
  }
  void loop(int i) {
    for (;;) {
      CrewJob *job;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_start_.wait(l, [&] { return slot_[size_t(i)].fresh; });
        slot_[size_t(i)].fresh = false;
        job = slot_[size_t(i)].job;
      }
      job->run();
      {
        std::lock_guard<std::mutex> l(mu_);
        slot_[size_t(i)].job = nullptr;
        cv_done_.notify_all();
      }
    }
  }
  std::vector<Slot> slot_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
};

// Workers one call may borrow: all of them, or this thread's share when a multi-device call set one
thread_local int t_crew_share = 0;  // 0 = no limit
// ... and the devices this process drives at once (they share the last-level cache)
thread_local int t_devices_sharing_host = 1;

// ---------------------------------------------------------------------------------------------------------------
// arena pool
// ---------------------------------------------------------------------------------------------------------------
constexpr int kArenaStreams = 3;
constexpr int kMaxSets = 8;

struct Arena {
  int device = -1;
  uint8_t *dev = nullptr;
  size_t dev_bytes = 0;
  uint8_t *pin = nullptr;
  size_t pin_bytes = 0;
  cudaStream_t stream[kArenaStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t h2d_ev[kMaxSets] = {};
  cudaEvent_t set_ev[kMaxSets] = {};
  cudaEvent_t slot_ev[32] = {};  // one per staging-ring slot (kMaxSlots)
  cudaEvent_t dir_ev[32] = {};   // link events of the direct pieces in flight (at most 24 at a time)
};

size_t round_arena(size_t b) {
  if (b == 0) return 0;
  size_t r = size_t(1) << 20;
  while (r < b && r < (size_t(32) << 20)) r <<= 1;
  return r >= b ? r : align_up(b, size_t(32) << 20);
}

class ArenaPool {
 public:
  static ArenaPool &get() {
    static ArenaPool *p = new ArenaPool();  // leaked on purpose (CUDA may be gone when statics are destroyed)
    return *p;
  }
  // device must be current.  Returns nullptr and *rc on failure.
  Arena *checkout(int device, size_t dev_bytes, size_t pin_bytes, int *rc) {
    dev_bytes = round_arena(dev_bytes);
    pin_bytes = round_arena(pin_bytes);
    Arena *a = nullptr;
    {
      std::lock_guard<std::mutex> l(mu_);
      int best = -1, biggest = -1;
      for (int i = 0; i < int(free_.size()); ++i) {
        Arena *f = free_[i];
        if (f->device != device) continue;
        if (f->dev_bytes >= dev_bytes && f->pin_bytes >= pin_bytes && (best < 0 || f->dev_bytes < free_[best]->dev_bytes))
          best = i;
        if (biggest < 0 || f->dev_bytes > free_[biggest]->dev_bytes) biggest = i;
      }
      const int take = best >= 0 ? best : biggest;
      if (take >= 0) {
        a = free_[take];
        free_.erase(free_.begin() + take);
      }
    }
    if (!a) {
      a = new (std::nothrow) Arena();
      if (!a) { *rc = FSSB200_EINVAL; return nullptr; }
      a->device = device;
      cudaError_t e = cudaSuccess;
      for (int i = 0; i < kArenaStreams && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&a->stream[i], cudaStreamNonBlocking);
      for (int i = 0; i < kMaxSets && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&a->h2d_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->set_ev[i], cudaEventDisableTiming);
      }
      for (int i = 0; i < 32 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&a->slot_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->dir_ev[i], cudaEventDisableTiming);
      }
      if (e != cudaSuccess) { destroy(a); *rc = int(e); return nullptr; }
    }
    if (a->dev_bytes < dev_bytes) {
      if (a->dev) cudaFree(a->dev);
      a->dev = nullptr;
      a->dev_bytes = 0;
      const cudaError_t e = cudaMalloc(&a->dev, dev_bytes);
      if (e != cudaSuccess) { (void)cudaGetLastError(); destroy(a); *rc = int(e); return nullptr; }
      a->dev_bytes = dev_bytes;
    }
    if (a->pin_bytes < pin_bytes) {
      if (a->pin) cudaFreeHost(a->pin);
      a->pin = nullptr;
      a->pin_bytes = 0;
      const cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&a->pin), pin_bytes, cudaHostAllocDefault);
      if (e != cudaSuccess) { (void)cudaGetLastError(); destroy(a); *rc = int(e); return nullptr; }
      a->pin_bytes = pin_bytes;
    }
    *rc = 0;
    return a;
  }
  void checkin(Arena *a) {
    if (!a) return;
    Arena *victim = nullptr;
    {
      std::lock_guard<std::mutex> l(mu_);
      free_.push_back(a);
      int on_dev = 0, smallest = -1;
      for (int i = 0; i < int(free_.size()); ++i) {
        if (free_[i]->device != a->device) continue;
        ++on_dev;
        if (smallest < 0 || free_[i]->dev_bytes + free_[i]->pin_bytes < free_[smallest]->dev_bytes + free_[smallest]->pin_bytes)
          smallest = i;
      }
      if (on_dev > kKeepPerDevice) {
        victim = free_[smallest];
        free_.erase(free_.begin() + smallest);
      }
    }
    if (victim) {
      DeviceGuard g(victim->device);
      destroy(victim);
    }
  }
  void trim() {
    std::vector<Arena *> all;
    {
      std::lock_guard<std::mutex> l(mu_);
      all.swap(free_);
    }
    for (Arena *a : all) {
      DeviceGuard g(a->device);
      destroy(a);
    }
  }
  size_t cached_bytes(size_t *pinned) {
    std::lock_guard<std::mutex> l(mu_);
    size_t d = 0, p = 0;
    for (Arena *a : free_) { d += a->dev_bytes; p += a->pin_bytes; }
    if (pinned) *pinned = p;
    return d;
  }

 private:
  static constexpr int kKeepPerDevice = 4;
  static void destroy(Arena *a) {
    if (a->dev) cudaFree(a->dev);
    if (a->pin) cudaFreeHost(a->pin);
    for (int i = 0; i < kArenaStreams; ++i)
      if (a->stream[i]) cudaStreamDestroy(a->stream[i]);
    for (int i = 0; i < kMaxSets; ++i) {
      if (a->h2d_ev[i]) cudaEventDestroy(a->h2d_ev[i]);
      if (a->set_ev[i]) cudaEventDestroy(a->set_ev[i]);
    }
    for (int i = 0; i < 32; ++i) {
      if (a->slot_ev[i]) cudaEventDestroy(a->slot_ev[i]);
      if (a->dir_ev[i]) cudaEventDestroy(a->dir_ev[i]);
    }
    delete a;
  }
  std::mutex mu_;
  std::vector<Arena *> free_;
};

// One host call's hold on an arena: sets the device, checks the arena out, and on every exit path drains the
// arena's streams before handing it back (a failed call must not leave copies into the caller's buffers in flight).
struct ArenaLease {
  DeviceGuard guard;
  Arena *a = nullptr;
  int rc = 0;
  ArenaLease(int device, size_t dev_bytes, size_t pin_bytes) : guard(device) {
    if (guard.err != cudaSuccess) { rc = int(guard.err); return; }
    a = ArenaPool::get().checkout(device, dev_bytes, pin_bytes, &rc);
  }
  // Synchronises the streams; returns the first error (or `rc_in` if that is already set).
  int drain(int rc_in) {
    if (!a) return rc_in ? rc_in : rc;
    for (int i = 0; i < kArenaStreams; ++i) {
      const cudaError_t e = cudaStreamSynchronize(a->stream[i]);
      if (!rc_in && e != cudaSuccess) rc_in = int(e);
    }
    return rc_in;
  }
  ~ArenaLease() {
    if (a) {
      for (int i = 0; i < kArenaStreams; ++i) cudaStreamSynchronize(a->stream[i]);
      ArenaPool::get().checkin(a);
    }
  }
  ArenaLease(const ArenaLease &) = delete;
  ArenaLease &operator=(const ArenaLease &) = delete;
};

size_t chunk_pref(const fssb200_ctx *c, size_t nkeys, size_t dflt) {
  size_t ck = c->host_chunk_keys.load(std::memory_order_relaxed);
  if (ck == 0) ck = dflt;
  if (ck > nkeys) ck = nkeys;
  return ck ? ck : 1;
}

bool is_pinned_or_device(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
}

// (cudaMemcpyDefault: the "host" arrays may just as well be managed or device memory)
#define H2D(dst, src, bytes, s) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s)
#define D2H(dst, src, bytes, s) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s)
#define TRY_BREAK(expr)                                  \
  {                                                      \
    const cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) { rc = int(e__); break; }    \
  }

// ---------------------------------------------------------------------------------------------------------------
// the adaptive pipeline of fssb200_eval_host
// ---------------------------------------------------------------------------------------------------------------
// Two granularities (measured, profiles/r02_host_pipeline.md):
//   CHUNK  = the keys of one kernel launch + one D2H (default 2^16): large enough that the launch is not latency-bound;
//   PIECE  = the keys of one staging-ring slot + one H2D copy (default 2^12, 2.2 MB packed): small enough that a ring
//            of a few slots stays in the last-level cache between the workers' stores and the copy engine's reads, so
//            the only DRAM traffic of a packed key is the 1056-byte read of the caller's row (the host's DRAM
//            bandwidth, ~185 GB/s on the boxes measured, is what bounds the call once more than one GPU shares it).
// Chunks [0, tail) are staged piece by piece from the front of the batch; chunks [tail, nchunks) were sent in the
// reference layout, taken one at a time from the back whenever the link was about to run dry.
constexpr size_t kPackBlock = 256;  // keys a worker claims at a time
constexpr int kMaxSlots = 32;

struct EvalPipe : CrewJob {
  // the call
  fssb200_ctx *c;
  const uint8_t *cws;
  size_t nkeys;
  // geometry
  size_t ck, nchunks, pk, ppc;  // keys per chunk, chunks, keys per piece, pieces per chunk
  size_t cwb, rowb;
  bool pack;        // staged rows are packed rows (else: a plain copy of the reference layout)
  bool nt_stores;
  uint8_t *stage;   // ring of pinned slots, one piece each
  size_t slot_bytes, nslots;
  // shared state
  struct PieceState {
    std::atomic<uint32_t> next{0};  // next block to claim
    std::atomic<uint32_t> done{0};  // keys staged so far
  };
  std::unique_ptr<PieceState[]> st;  // indexed by global piece number q = chunk * ppc + piece
  std::atomic<size_t> cur{0};        // piece the workers are staging
  std::atomic<size_t> tail{0};       // chunks [tail, nchunks) go in the reference layout
  std::atomic<size_t> free_upto{0};  // staging may write staged-piece ordinal o iff o < free_upto
  std::atomic<bool> stop{false};
  std::mutex claim_mu;

  size_t keys_of_chunk(size_t ch) const { return ch + 1 < nchunks ? ck : nkeys - ch * ck; }
  size_t pieces_of_chunk(size_t ch) const { return (keys_of_chunk(ch) + pk - 1) / pk; }
  size_t keys_of_piece(size_t q) const {
    const size_t ch = q / ppc, pi = q % ppc, kc = keys_of_chunk(ch);
    return pi * pk >= kc ? 0 : std::min(pk, kc - pi * pk);
  }
  size_t staged_row_bytes() const { return pack ? rowb : cwb; }
  // Ordinal of piece q among the NON-EMPTY staged pieces (only the batch's last chunk can have empty pieces, and they
  // trail): ring slot = ordinal % nslots.  All chunks before q's are full, so the ordinal is q itself.
  uint8_t *slot_of(size_t q) const { return stage + (q % nslots) * slot_bytes; }

  // Stages one block if there is one.  Returns false when there is nothing to do right now.
  bool stage_step() {
    const size_t q = cur.load(std::memory_order_acquire);
    if (q / ppc >= tail.load(std::memory_order_acquire)) return false;
    const size_t k = keys_of_piece(q);
    if (k && q >= free_upto.load(std::memory_order_acquire)) return false;
    const uint32_t nblocks = uint32_t((k + kPackBlock - 1) / kPackBlock);
    const uint32_t b = st[q].next.fetch_add(1, std::memory_order_relaxed);
    if (b >= nblocks) {
      std::lock_guard<std::mutex> l(claim_mu);
      if (cur.load(std::memory_order_relaxed) == q) cur.store(q + 1, std::memory_order_release);
      return true;
    }
    const size_t k0 = size_t(b) * kPackBlock, k1 = std::min(k, k0 + kPackBlock);
    const size_t g0 = (q / ppc) * ck + (q % ppc) * pk;  // first key of the piece
    uint8_t *slot = slot_of(q);
    if (pack) pack_rows_range(cws + g0 * cwb, slot, k0, k1, c->ncw, nt_stores);
    else stage_copy(cws + (g0 + k0) * cwb, slot + k0 * cwb, (k1 - k0) * cwb, nt_stores);
    st[q].done.fetch_add(uint32_t(k1 - k0), std::memory_order_release);
    return true;
  }
  void run() override {
    int idle = 0;
    while (!stop.load(std::memory_order_acquire)) {
      if (cur.load(std::memory_order_acquire) / ppc >= tail.load(std::memory_order_acquire)) break;
      if (stage_step()) {
        idle = 0;
      } else if (++idle < 64) {
        cpu_relax();
      } else {
        std::this_thread::yield();
      }
    }
  }
};

// Large batches with worker threads available.  `allow_direct`: the caller's rows are pinned, so chunks may also
// cross the link straight from them.
int eval_host_pipelined(fssb200_ctx *c, int party, const void *seeds_v, const void *cws, const void *ocws_v,
    const void *xs_v, void *ys_v, size_t nkeys, Crew *crew, bool allow_direct, bool ys_pinned, bool stage_nothing = false) {
  const uint8_t *seeds = static_cast<const uint8_t *>(seeds_v), *ocws = static_cast<const uint8_t *>(ocws_v),
                *xs = static_cast<const uint8_t *>(xs_v);
  uint8_t *ys = static_cast<uint8_t *>(ys_v);
  EvalPipe P;
  P.c = c;
  P.cws = static_cast<const uint8_t *>(cws);
  P.nkeys = nkeys;
  P.cwb = size_t(c->ncw) * 32;
  P.rowb = fssb200_packed_row_bytes(c);
  const size_t ib = size_t(c->p.in_bytes);
  P.pack = P.rowb != 0;
  // ring: cache-resident (ordinary stores) when this rank's share of the last-level cache can hold it together with
  // the rows streaming through; else non-temporal stores into a ring that lives in DRAM
  const size_t llc = llc_bytes_per_rank() / size_t(t_devices_sharing_host);
  const bool cached_ring = llc >= (size_t(6) << 20);
  P.nt_stores = env_int("FSSB200_PACK_NT", cached_ring ? 0 : 1) != 0;
  P.ck = chunk_pref(c, nkeys, size_t(1) << env_int("FSSB200_PIPE_CHUNK_BITS", 16));
  P.nchunks = (nkeys + P.ck - 1) / P.ck;
  // piece: 2^14 keys (8.9 MB packed) when the ring of 4 has the cache to itself, 2^13 when ranks share it; smaller
  // pieces lose more to their hand-offs than they gain in residency (2 ranks: 2^13 -> 66-69 ms, 2^12 -> 84-87 ms)
  const int piece_bits = llc >= (size_t(48) << 20) ? 14 : 13;
  P.pk = std::min(P.ck, size_t(1) << env_int("FSSB200_PIPE_PIECE_BITS", piece_bits));
  P.ppc = (P.ck + P.pk - 1) / P.pk;
  P.slot_bytes = align_up(P.pk * P.staged_row_bytes(), 4096);
  {
    size_t want = cached_ring ? 4 : 6;
    want = size_t(std::max(2, std::min(kMaxSlots, env_int("FSSB200_PIPE_SLOTS", int(want)))));
    P.nslots = stage_nothing ? 0 : std::min(want, P.nchunks * P.ppc);
  }
  const int nsets = int(std::min<size_t>(kMaxSets, std::max<size_t>(2, std::min<size_t>(P.nchunks, 4))));
  // device set: rows (reference layout: the larger of the two formats) | seeds | xs | ocws | ys
  const size_t off_seeds = align_up(P.ck * P.cwb, 256), off_xs = off_seeds + align_up(P.ck * 16, 256),
               off_ocws = off_xs + align_up(P.ck * ib, 256), off_ys = off_ocws + (ocws ? align_up(P.ck * 16, 256) : 0);
  const size_t set_bytes = align_up(off_ys + P.ck * 16, 4096);
  const size_t ys_stage = ys_pinned ? 0 : align_up(P.ck * 16, 4096);  // pageable ys: results land in pinned memory first
  ArenaLease L(c->p.device, set_bytes * nsets, P.nslots * P.slot_bytes + ys_stage * nsets);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  P.stage = A.pin;
  uint8_t *ys_pin = A.pin + P.nslots * P.slot_bytes;
  P.st.reset(new (std::nothrow) EvalPipe::PieceState[P.nchunks * P.ppc]);
  if (!P.st) return FSSB200_EINVAL;
  P.tail.store(stage_nothing ? 0 : P.nchunks);  // stage_nothing: every chunk crosses the link as it is, piece by piece
  P.free_upto.store(P.nslots);

  // (no idle worker and pageable inputs: the calling thread stages alone -- still correct, just slower)
  const int lent = (crew && !stage_nothing) ? crew->begin(&P, t_crew_share > 0 ? t_crew_share : crew->workers()) : 0;

  struct SetUse {
    bool busy = false;      // holds a chunk
    bool launched = false;  // ... whose kernel + D2H are queued (set_ev recorded)
    size_t chunk = 0;
  };
  SetUse sets[kMaxSets];
  struct LinkOp {
    cudaEvent_t ev;
    size_t bytes;
    bool piece;
  };
  std::deque<LinkOp> link;   // row copies not yet known to be complete, in issue order
  size_t link_bytes = 0;     // ... and their bytes: how much work the link has queued
  size_t pieces_copied = 0;  // staged pieces whose copy has completed (they complete in order)
  size_t next_piece = 0;     // next staged piece to submit
  int cur_set = -1;          // device set of the staged chunk in progress
  size_t issue_seq = 0;      // chunks started so far (stream round-robin)
  cudaStream_t cur_stream = nullptr;
  uint64_t direct_keys = 0;
  int rc = 0;
  // the link should always have about two pieces queued; below that, rows from the back of the batch go as they are
  const size_t low_water = env_int("FSSB200_PIPE_LOW_WATER_MB", 0) > 0 ? size_t(env_int("FSSB200_PIPE_LOW_WATER_MB", 0)) << 20
                                                                         : P.pk * P.staged_row_bytes() * 7 / 4;

  auto retire_set = [&](int s, bool wait) -> bool {  // results of the chunk in set s are on the host
    if (!sets[s].busy) return true;
    if (!sets[s].launched) return false;  // the staged chunk still being filled
    const cudaError_t e = wait ? cudaEventSynchronize(A.set_ev[s]) : cudaEventQuery(A.set_ev[s]);
    if (e == cudaErrorNotReady) return false;
    if (e != cudaSuccess) { rc = int(e); return false; }
    if (ys_stage)
      std::memcpy(ys + sets[s].chunk * P.ck * 16, ys_pin + size_t(s) * ys_stage, P.keys_of_chunk(sets[s].chunk) * 16);
    sets[s].busy = sets[s].launched = false;
    return true;
  };
  auto acquire_set = [&]() -> int {
    for (int s = 0; s < nsets; ++s)
      if (!sets[s].busy) return s;
    for (int s = 0; s < nsets; ++s)
      if (retire_set(s, false)) return s;
    return -1;
  };
  auto poll_link = [&]() {
    while (!link.empty()) {
      const cudaError_t e = cudaEventQuery(link.front().ev);
      if (e == cudaErrorNotReady) break;
      if (e != cudaSuccess) { rc = int(e); break; }
      link_bytes -= link.front().bytes;
      if (link.front().piece) P.free_upto.store(++pieces_copied + P.nslots, std::memory_order_release);
      link.pop_front();
    }
  };
  // small arrays of a chunk straight from the caller's memory (pageable: the driver stages them, they are 36 B per key)
  auto copy_small = [&](uint8_t *d, size_t ch, cudaStream_t str) -> int {
    const size_t k = P.keys_of_chunk(ch), g0 = ch * P.ck;
    int rc2 = 0;
    do {
      const cudaError_t e1 = H2D(d + off_seeds, seeds + g0 * 16, k * 16, str);
      if (e1 != cudaSuccess) { rc2 = int(e1); break; }
      const cudaError_t e2 = H2D(d + off_xs, xs + g0 * ib, k * ib, str);
      if (e2 != cudaSuccess) { rc2 = int(e2); break; }
      if (ocws) {
        const cudaError_t e3 = H2D(d + off_ocws, ocws + g0 * 16, k * 16, str);
        if (e3 != cudaSuccess) { rc2 = int(e3); break; }
      }
    } while (0);
    return rc2;
  };
  auto launch_chunk = [&](int s, size_t ch, bool packed_rows, cudaStream_t str) {
    const size_t k = P.keys_of_chunk(ch);
    uint8_t *d = A.dev + size_t(s) * set_bytes;
    do {
      rc = packed_rows ? fssb200_eval_packed(c, party, d + off_seeds, d, ocws ? d + off_ocws : nullptr, d + off_xs, d + off_ys, k, str)
                       : fssb200_eval(c, party, d + off_seeds, d, ocws ? d + off_ocws : nullptr, d + off_xs, d + off_ys, k, str);
      if (rc) break;
      TRY_BREAK(D2H(ys_stage ? ys_pin + size_t(s) * ys_stage : ys + ch * P.ck * 16, d + off_ys, k * 16, str));
      TRY_BREAK(cudaEventRecord(A.set_ev[s], str));
    } while (0);
    sets[s].busy = sets[s].launched = true;
    sets[s].chunk = ch;
  };

  // An OPEN chunk is being filled on the device piece by piece; its kernel is launched behind its last piece.
  struct Open {
    int set = -1;
    size_t chunk = 0, next_pi = 0;
    cudaStream_t str = nullptr;
  };
  Open so, dio;              // the staged chunk (front of the batch) and the direct chunk (back) in progress
  size_t direct_next = 0;    // stage_nothing: next chunk to send
  int dir_ev_next = 0;       // direct pieces take their link events from a small ring
  auto open_chunk = [&](Open &o, int s, size_t ch) {  // s: a free device set
    o.set = s;
    o.chunk = ch;
    o.next_pi = 0;
    o.str = A.stream[issue_seq++ % kArenaStreams];
    sets[s].busy = true;  // (held from the first piece on; `launched` once its kernel is queued)
    sets[s].chunk = ch;
    rc = copy_small(A.dev + size_t(s) * set_bytes, ch, o.str);
  };

  int idle = 0;
  while (!rc) {
    poll_link();
    if (rc) break;
    const size_t tl = P.tail.load(std::memory_order_acquire);
    const bool staged_left = next_piece / P.ppc < tl;
    if (!staged_left && dio.set < 0 && (!stage_nothing || direct_next >= P.nchunks)) break;  // every chunk has been issued
    bool progressed = false;
    if (staged_left) {
      const size_t ch = next_piece / P.ppc, pi = next_piece % P.ppc, kp = P.keys_of_piece(next_piece);
      if (kp == 0) {  // trailing empty piece of the batch's last chunk
        ++next_piece;
        continue;
      }
      if (so.set < 0 && P.st[next_piece].done.load(std::memory_order_acquire) == uint32_t(kp)) {
        const int s = acquire_set();  // a staged chunk starts: it needs a device set
        if (rc) break;
        if (s >= 0) open_chunk(so, s, ch);
        if (rc) break;
      }
      if (so.set >= 0 && P.st[next_piece].done.load(std::memory_order_acquire) == uint32_t(kp)) {
        // submit the piece: its rows go behind the chunk's earlier pieces in the set
        const size_t bytes = kp * P.staged_row_bytes();
        const int slot = int(next_piece % P.nslots);
        do {
          TRY_BREAK(H2D(A.dev + size_t(so.set) * set_bytes + pi * P.pk * P.staged_row_bytes(), P.slot_of(next_piece), bytes, so.str));
          TRY_BREAK(cudaEventRecord(A.slot_ev[slot], so.str));
        } while (0);
        if (rc) break;
        link.push_back(LinkOp{A.slot_ev[slot], bytes, true});
        link_bytes += bytes;
        ++next_piece;
        if (pi + 1 == P.pieces_of_chunk(ch)) {  // the chunk is complete on the device side: evaluate it
          launch_chunk(so.set, ch, P.pack, so.str);
          so.set = -1;
          next_piece = (ch + 1) * P.ppc;
        }
        progressed = true;
      }
    }
    // The link is about to run dry and no staged piece is ready (or nothing is left to stage): rows from the back of
    // the batch cross as they are, one piece at a time so that a staged piece never waits long behind them.
    if (!progressed && allow_direct && (link_bytes < low_water || !staged_left) && link.size() < 24) {
      if (dio.set < 0 && (staged_left || stage_nothing)) {
        const int s = acquire_set();  // (before the claim: a chunk taken from the back is never given back)
        if (rc) break;
        if (s >= 0) {
          size_t take = ~size_t(0);
          if (stage_nothing) {  // nothing is staged: the chunks simply go in order
            if (direct_next < P.nchunks) take = direct_next++;
          } else {
            std::lock_guard<std::mutex> l(P.claim_mu);
            const size_t t = P.tail.load(std::memory_order_relaxed);
            if (t > 0 && t - 1 > P.cur.load(std::memory_order_relaxed) / P.ppc && t - 1 > next_piece / P.ppc) {
              take = t - 1;
              P.tail.store(take, std::memory_order_release);
            }
          }
          if (take != ~size_t(0)) open_chunk(dio, s, take);
          if (rc) break;
        }
      }
      if (dio.set >= 0) {
        const size_t kc = P.keys_of_chunk(dio.chunk), k0 = dio.next_pi * P.pk, kp = std::min(P.pk, kc - k0);
        cudaEvent_t ev = A.dir_ev[dir_ev_next];
        dir_ev_next = (dir_ev_next + 1) % 32;
        do {
          TRY_BREAK(H2D(A.dev + size_t(dio.set) * set_bytes + k0 * P.cwb, P.cws + (dio.chunk * P.ck + k0) * P.cwb, kp * P.cwb, dio.str));
          TRY_BREAK(cudaEventRecord(ev, dio.str));
        } while (0);
        if (rc) break;
        link.push_back(LinkOp{ev, kp * P.cwb, false});
        link_bytes += kp * P.cwb;
        direct_keys += kp;
        if (k0 + kp == kc) {
          launch_chunk(dio.set, dio.chunk, false, dio.str);
          dio.set = -1;
        } else {
          ++dio.next_pi;
        }
        progressed = true;
      }
    }
    if (progressed) {
      idle = 0;
      continue;
    }
    if (P.stage_step()) {
      idle = 0;
    } else if (++idle < 64) {
      cpu_relax();
    } else {
      // nothing to submit and nothing to stage: the link or the GPU has to make progress first.  Back off instead of
      // hammering cudaEventQuery -- in a process that drives several GPUs the submitters share the driver's locks
      // (2 GPUs from one process: 90.9 ms per step with a spinning poll)
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
  }
  P.stop.store(true, std::memory_order_release);
  if (lent) crew->end(&P);
  // drain: results of every issued chunk
  for (int s = 0; s < nsets; ++s) {
    const int first = rc;
    rc = 0;
    if (sets[s].busy && sets[s].launched) retire_set(s, true);
    if (first) rc = first;
  }
  rc = L.drain(rc);
  c->last_direct_keys.store(direct_keys, std::memory_order_relaxed);
  c->last_packed_keys.store(nkeys - direct_keys, std::memory_order_relaxed);  // staged: packed, or copied (no padding)
  c->last_pack_threads.store(stage_nothing ? 0 : lent + 1, std::memory_order_relaxed);
  return rc;
}

// Small batches, or no worker threads: chunks of the reference layout through two device sets.
int eval_host_simple(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws, const void *xs,
    void *ys, size_t nkeys) {
  const size_t ck = chunk_pref(c, nkeys, size_t(1) << 18), cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  const size_t set_bytes = align_up(ck * 16, 256) * 3 + align_up(ck * cwb, 256) + align_up(ck * ib, 256);
  const int nsets = nkeys > ck ? 2 : 1;
  ArenaLease L(c->p.device, set_bytes * nsets, 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nkeys - k0);
    const int b = int(chunk & 1);
    cudaStream_t s = A.stream[b];
    uint8_t *d_seeds = A.dev + size_t(b) * set_bytes;
    uint8_t *d_cws = d_seeds + align_up(k * 16, 256);
    uint8_t *d_ocws = d_cws + align_up(k * cwb, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * ib, 256);
    // (a set is reused two chunks later on the same stream: stream order protects it)
    TRY_BREAK(H2D(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, s));
    TRY_BREAK(H2D(d_cws, static_cast<const uint8_t *>(cws) + k0 * cwb, k * cwb, s));
    if (ocws) TRY_BREAK(H2D(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, s));
    TRY_BREAK(H2D(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, s));
    rc = fssb200_eval(c, party, d_seeds, d_cws, ocws ? d_ocws : nullptr, d_xs, d_ys, k, s);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, s));
  }
  rc = L.drain(rc);
  c->last_direct_keys.store(nkeys, std::memory_order_relaxed);
  c->last_packed_keys.store(0, std::memory_order_relaxed);
  c->last_pack_threads.store(0, std::memory_order_relaxed);
  return rc;
}

}  // namespace

extern "C" {

// ---- packed rows: host-side format conversion ---------------------------------------------------------------------
namespace {
struct PackRowsJob : CrewJob {
  const uint8_t *src;
  uint8_t *dst;
  size_t nkeys;
  int ncw;
  bool nt;
  std::atomic<size_t> next{0};
  void run() override {
    for (;;) {
      const size_t b = next.fetch_add(1024, std::memory_order_relaxed);
      if (b >= nkeys) return;
      pack_rows_range(src, dst, b, std::min(nkeys, b + 1024), ncw, nt);
    }
  }
};
}  // namespace

int fssb200_pack_rows(const fssb200_ctx *c, const void *cws, void *rows, size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  if (!fssb200_packed_row_bytes(c)) return FSSB200_ESCHEME;
  if (!cws || !rows) return FSSB200_EINVAL;
  PackRowsJob job;
  job.src = static_cast<const uint8_t *>(cws);
  job.dst = static_cast<uint8_t *>(rows);
  job.nkeys = nkeys;
  job.ncw = c->ncw;
  job.nt = aligned16(rows);
  Crew *crew = nkeys >= 4096 ? Crew::get() : nullptr;
  const int lent = crew ? crew->begin(&job, crew->workers()) : 0;
  job.run();  // the calling thread takes blocks too
  if (lent) crew->end(&job);
  return 0;
}

int fssb200_ctx_host_pack_threads(const fssb200_ctx *c) {
  if (!c || !fssb200_packed_row_bytes(c)) return 0;
  Crew *crew = Crew::get();
  return crew ? crew->workers() + 1 : 0;
}

int fssb200_ctx_host_stats(const fssb200_ctx *c, uint64_t *packed_keys, uint64_t *direct_keys, int *threads) {
  if (!c) return FSSB200_EINVAL;
  if (packed_keys) *packed_keys = c->last_packed_keys.load(std::memory_order_relaxed);
  if (direct_keys) *direct_keys = c->last_direct_keys.load(std::memory_order_relaxed);
  if (threads) *threads = c->last_pack_threads.load(std::memory_order_relaxed);
  return 0;
}

int fssb200_ctx_set_host_mode(fssb200_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 3) return FSSB200_EINVAL;
  c->host_mode.store(mode, std::memory_order_relaxed);
  return 0;
}

int fssb200_ctx_reserve_host(fssb200_ctx *c, size_t max_keys_per_chunk) {
  if (!c) return FSSB200_EINVAL;
  c->host_chunk_keys.store(max_keys_per_chunk, std::memory_order_relaxed);
  return 0;
}

void fssb200_host_trim(void) { ArenaPool::get().trim(); }

int fssb200_host_cached_bytes(uint64_t *device_bytes, uint64_t *pinned_bytes) {
  size_t p = 0;
  const size_t d = ArenaPool::get().cached_bytes(&p);
  if (device_bytes) *device_bytes = d;
  if (pinned_bytes) *pinned_bytes = p;
  return 0;
}

// ---- point evaluation ------------------------------------------------------------------------------------------------
int fssb200_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    const void *xs, void *ys, size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  if (c->p.scheme == FSSB200_SCHEME_GROTTO || c->p.scheme == FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (party != 0 && party != 1) return FSSB200_EINVAL;
  if (!seeds || !cws || !xs || !ys) return FSSB200_EINVAL;
  if (c->p.scheme == FSSB200_SCHEME_HALFTREE && !ocws) return FSSB200_EINVAL;
  if (nkeys == 0) return 0;
  if (nkeys >= 8192) {
    DeviceGuard g(c->p.device);
    if (g.err != cudaSuccess) return int(g.err);
    Crew *crew = Crew::get();
    const bool in_pinned = is_pinned_or_device(cws) && is_pinned_or_device(seeds) && is_pinned_or_device(xs) &&
        (!ocws || is_pinned_or_device(ocws));
    const bool packable = fssb200_packed_row_bytes(c) != 0;
    int mode = c->host_mode.load(std::memory_order_relaxed);
    // Auto (0): packing pays while the LINKS are the bound -- one or two GPUs per host on the boxes measured.  From
    // three GPUs on, the links together ask for more than the host memory system serves (4 GPUs: 167 GB/s for all,
    // 38-52 GB/s each), a packed key costs the same DRAM read as a plain one, and the packing cores only take
    // bandwidth from the copy engines: 137 ms packed vs 119 ms plain per 2^22 keys and rank (profiles/r02_host_pipeline.md).
    if (mode == 0 && in_pinned && std::max(env_int("LOCAL_WORLD_SIZE", 1), t_devices_sharing_host) >= 3) mode = 1;
    if (mode == 3) mode = 0;  // adaptive pipeline whatever the rank count (A/B measurements)
    if (mode == 1 && in_pinned)  // the reference layout crosses as it is, in pieces that keep the link's queue full
      return eval_host_pipelined(c, party, seeds, cws, ocws, xs, ys, nkeys, nullptr, true, is_pinned_or_device(ys), true);
    // (pinned inputs of a scheme without padding: nothing to gain from staging either)
    if (crew && mode != 1 && (packable || !in_pinned))
      return eval_host_pipelined(c, party, seeds, cws, ocws, xs, ys, nkeys, crew, in_pinned && mode != 2,
          is_pinned_or_device(ys));
    if (in_pinned)
      return eval_host_pipelined(c, party, seeds, cws, ocws, xs, ys, nkeys, nullptr, true, is_pinned_or_device(ys), true);
  }
  return eval_host_simple(c, party, seeds, cws, ocws, xs, ys, nkeys);
}

// ---- VDPF ------------------------------------------------------------------------------------------------------------
int fssb200_vdpf_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *cs,
    void *ocws, void *status, size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!s0s || !alphas || !betas || !cws || !cs || !ocws || !status) return FSSB200_EINVAL;
  if (nkeys == 0) return 0;
  const size_t ck = chunk_pref(c, nkeys, size_t(1) << 18), cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  const size_t set_bytes = align_up(ck * 32, 256) + align_up(ck * ib, 256) + align_up(ck * 16, 256) * 2 +
      align_up(ck * cwb, 256) + align_up(ck * 64, 256) + align_up(ck * 4, 256);
  ArenaLease L(c->p.device, set_bytes * (nkeys > ck ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nkeys - k0);
    const int b = int(chunk & 1);
    cudaStream_t s = A.stream[b];
    uint8_t *d_s0s = A.dev + size_t(b) * set_bytes;
    uint8_t *d_al = d_s0s + align_up(k * 32, 256);
    uint8_t *d_be = d_al + align_up(k * ib, 256);
    uint8_t *d_cws = d_be + align_up(k * 16, 256);
    uint8_t *d_cs = d_cws + align_up(k * cwb, 256);
    uint8_t *d_ocws = d_cs + align_up(k * 64, 256);
    uint8_t *d_st = d_ocws + align_up(k * 16, 256);
    TRY_BREAK(H2D(d_s0s, static_cast<const uint8_t *>(s0s) + k0 * 32, k * 32, s));
    TRY_BREAK(H2D(d_al, static_cast<const uint8_t *>(alphas) + k0 * ib, k * ib, s));
    TRY_BREAK(H2D(d_be, static_cast<const uint8_t *>(betas) + k0 * 16, k * 16, s));
    TRY_BREAK(cudaMemsetAsync(d_ocws, 0, k * 16, s));  // Gen leaves ocw untouched when it returns 1
    rc = fssb200_vdpf_gen(c, d_s0s, d_al, d_be, d_cws, d_cs, d_ocws, d_st, k, s);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(cws) + k0 * cwb, d_cws, k * cwb, s));
    TRY_BREAK(D2H(static_cast<uint8_t *>(cs) + k0 * 64, d_cs, k * 64, s));
    TRY_BREAK(D2H(static_cast<uint8_t *>(ocws) + k0 * 16, d_ocws, k * 16, s));
    TRY_BREAK(D2H(static_cast<uint8_t *>(status) + k0 * 4, d_st, k * 4, s));
  }
  return L.drain(rc);
}

int fssb200_vdpf_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs,
    const void *ocws, const void *xs, void *ys, void *pis, size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!seeds || !cws || !cs || !ocws || !xs || !ys || !pis) return FSSB200_EINVAL;
  if (nkeys == 0) return 0;
  const size_t ck = chunk_pref(c, nkeys, size_t(1) << 18), cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  const size_t set_bytes = align_up(ck * 16, 256) * 3 + align_up(ck * cwb, 256) + align_up(ck * 64, 256) * 2 +
      align_up(ck * ib, 256);
  ArenaLease L(c->p.device, set_bytes * (nkeys > ck ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nkeys - k0);
    const int b = int(chunk & 1);
    cudaStream_t s = A.stream[b];
    uint8_t *d_seeds = A.dev + size_t(b) * set_bytes;
    uint8_t *d_cws = d_seeds + align_up(k * 16, 256);
    uint8_t *d_cs = d_cws + align_up(k * cwb, 256);
    uint8_t *d_ocws = d_cs + align_up(k * 64, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * ib, 256);
    uint8_t *d_pis = d_ys + align_up(k * 16, 256);
    TRY_BREAK(H2D(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, s));
    TRY_BREAK(H2D(d_cws, static_cast<const uint8_t *>(cws) + k0 * cwb, k * cwb, s));
    TRY_BREAK(H2D(d_cs, static_cast<const uint8_t *>(cs) + k0 * 64, k * 64, s));
    TRY_BREAK(H2D(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, s));
    TRY_BREAK(H2D(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, s));
    rc = fssb200_vdpf_eval(c, party, d_seeds, d_cws, d_cs, d_ocws, d_xs, d_ys, d_pis, k, s);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, s));
    TRY_BREAK(D2H(static_cast<uint8_t *>(pis) + k0 * 64, d_pis, k * 64, s));
  }
  return L.drain(rc);
}

// ---- level-major host arrays -----------------------------------------------------------------------------------------
int fssb200_eval_levelmajor_host(fssb200_ctx *c, int party, const void *seeds, const void *cw_s, const void *cw_v,
    const void *extra, const void *out_cw, const void *ocws, const void *xs, void *ys, size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  const int scheme = c->p.scheme;
  if (scheme == FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (!seeds || !cw_s || !xs || !ys) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_DCF && (!cw_v || !out_cw)) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_DPF && (!extra || !out_cw)) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_HALFTREE && (!extra || !ocws)) return FSSB200_EINVAL;
  if (nkeys == 0) return 0;
  const size_t ck = chunk_pref(c, nkeys, size_t(1) << 18), ib = size_t(c->p.in_bytes), n = size_t(c->p.in_bits),
               nw = (n + 31) / 32;
  // one set: seeds | cw_s[n][k] | cw_v[n][k] | extra[nw][k] | out_cw | ocws | xs | ys
  const size_t set_bytes = align_up(ck * 16, 256) * 4 + align_up(n * ck * 16, 256) * (cw_v ? 2 : 1) +
      (extra ? align_up(nw * ck * 4, 256) : 0) + align_up(ck * ib, 256);
  ArenaLease L(c->p.device, set_bytes * (nkeys > ck ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  const uint8_t *h_s = static_cast<const uint8_t *>(cw_s), *h_v = static_cast<const uint8_t *>(cw_v),
                *h_e = static_cast<const uint8_t *>(extra);
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nkeys - k0);
    const int b = int(chunk & 1);
    cudaStream_t st = A.stream[b];
    uint8_t *d_seeds = A.dev + size_t(b) * set_bytes;
    uint8_t *d_s = d_seeds + align_up(k * 16, 256);
    uint8_t *d_v = d_s + align_up(n * k * 16, 256);
    uint8_t *d_e = d_v + (cw_v ? align_up(n * k * 16, 256) : 0);
    uint8_t *d_oc = d_e + (extra ? align_up(nw * k * 4, 256) : 0);
    uint8_t *d_ocws = d_oc + align_up(k * 16, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * ib, 256);
    TRY_BREAK(H2D(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, st));
    TRY_BREAK(cudaMemcpy2DAsync(d_s, k * 16, h_s + k0 * 16, nkeys * 16, k * 16, n, cudaMemcpyDefault, st));
    if (cw_v) TRY_BREAK(cudaMemcpy2DAsync(d_v, k * 16, h_v + k0 * 16, nkeys * 16, k * 16, n, cudaMemcpyDefault, st));
    if (extra) TRY_BREAK(cudaMemcpy2DAsync(d_e, k * 4, h_e + k0 * 4, nkeys * 4, k * 4, nw, cudaMemcpyDefault, st));
    if (out_cw) TRY_BREAK(H2D(d_oc, static_cast<const uint8_t *>(out_cw) + k0 * 16, k * 16, st));
    if (ocws) TRY_BREAK(H2D(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, st));
    TRY_BREAK(H2D(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, st));
    rc = fssb200_eval_levelmajor(c, party, d_seeds, d_s, cw_v ? d_v : nullptr, extra ? d_e : nullptr,
        out_cw ? d_oc : nullptr, ocws ? d_ocws : nullptr, d_xs, d_ys, k, st);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, st));
  }
  return L.drain(rc);
}

// ---- key generation ----------------------------------------------------------------------------------------------------
int fssb200_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *ocws,
    size_t nkeys) {
  if (!c) return FSSB200_EINVAL;
  if (c->p.scheme == FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!s0s || !alphas || !cws) return FSSB200_EINVAL;
  const bool grotto = c->p.scheme == FSSB200_SCHEME_GROTTO, half = c->p.scheme == FSSB200_SCHEME_HALFTREE;
  if ((!grotto && !betas) || (half && !ocws)) return FSSB200_EINVAL;
  if (nkeys == 0) return 0;
  const size_t ck = chunk_pref(c, nkeys, size_t(1) << 18), cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  const size_t set_bytes = align_up(ck * 32, 256) + align_up(ck * cwb, 256) + align_up(ck * 16, 256) * 2 + align_up(ck * ib, 256);
  ArenaLease L(c->p.device, set_bytes * (nkeys > ck ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nkeys - k0);
    const int b = int(chunk & 1);
    cudaStream_t s = A.stream[b];
    uint8_t *d_s0s = A.dev + size_t(b) * set_bytes;
    uint8_t *d_cws = d_s0s + align_up(k * 32, 256);
    uint8_t *d_ocws = d_cws + align_up(k * cwb, 256);
    uint8_t *d_al = d_ocws + align_up(k * 16, 256);
    uint8_t *d_be = d_al + align_up(k * ib, 256);
    TRY_BREAK(H2D(d_s0s, static_cast<const uint8_t *>(s0s) + k0 * 32, k * 32, s));
    TRY_BREAK(H2D(d_al, static_cast<const uint8_t *>(alphas) + k0 * ib, k * ib, s));
    if (!grotto) TRY_BREAK(H2D(d_be, static_cast<const uint8_t *>(betas) + k0 * 16, k * 16, s));
    rc = fssb200_gen(c, d_s0s, d_al, grotto ? nullptr : d_be, d_cws, half ? d_ocws : nullptr, k, s);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(cws) + k0 * cwb, d_cws, k * cwb, s));
    if (half) TRY_BREAK(D2H(static_cast<uint8_t *>(ocws) + k0 * 16, d_ocws, k * 16, s));
  }
  return L.drain(rc);
}

// ---- PRG blocks of host seeds (the `prg.Gen(seed)` member of the header shim) -----------------------------------------
int fssb200_prg_gen_host(fssb200_ctx *c, const void *seeds, void *out, int mul, size_t nseeds) {
  if (!c) return FSSB200_EINVAL;
  if (!seeds || !out || (mul != 1 && mul != 2 && mul != 4)) return FSSB200_EINVAL;
  if (nseeds == 0) return 0;
  const size_t ck = std::min<size_t>(nseeds, size_t(1) << 20);
  const size_t set_bytes = align_up(ck * 16, 256) + align_up(ck * 16 * size_t(mul), 256);
  ArenaLease L(c->p.device, set_bytes * (nseeds > ck ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nseeds && !rc; k0 += ck, ++chunk) {
    const size_t k = std::min(ck, nseeds - k0);
    const int b = int(chunk & 1);
    cudaStream_t s = A.stream[b];
    uint8_t *d_in = A.dev + size_t(b) * set_bytes, *d_out = d_in + align_up(k * 16, 256);
    TRY_BREAK(H2D(d_in, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, s));
    rc = fssb200_prg_gen(c, d_in, d_out, mul, k, s);
    if (rc) break;
    TRY_BREAK(D2H(static_cast<uint8_t *>(out) + k0 * 16 * size_t(mul), d_out, k * 16 * size_t(mul), s));
  }
  return L.drain(rc);
}

// ---- full-domain evaluation -------------------------------------------------------------------------------------------
int fssb200_eval_all_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    void *ys, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count) {
  if (!c) return FSSB200_EINVAL;
  if (c->p.scheme == FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!seeds || !cws || !ys) return FSSB200_EINVAL;
  const bool half = c->p.scheme == FSSB200_SCHEME_HALFTREE, grotto = c->p.scheme == FSSB200_SCHEME_GROTTO;
  if (half && !ocws) return FSSB200_EINVAL;
  const int n = c->p.in_bits;
  if (n > 40) return FSSB200_EDOMAIN;
  const uint64_t N = uint64_t(1) << n;
  if (leaf_begin >= N) return FSSB200_ERANGE;
  if (leaf_count == 0) leaf_count = N - leaf_begin;
  if (leaf_count > N - leaf_begin) return FSSB200_ERANGE;
  const uint64_t granule = fssb200_eval_all_granule(c);
  if ((leaf_begin | leaf_count) & (granule - 1)) return FSSB200_ERANGE;
  if (grotto && leaf_begin != 0) return FSSB200_ERANGE;
  if (nkeys == 0) return 0;
  const size_t cwb = size_t(c->ncw) * 32, leaf_bytes = grotto ? 1 : 16;
  // 64 MiB of leaves per set (Grotto: the scan needs whole keys)
  const uint64_t set_leaves = (uint64_t(std::min(4096, std::max(1, env_int("FSSB200_ALL_SET_MB", 64)))) << 20) / leaf_bytes;
  uint64_t leaves_per_chunk = grotto ? leaf_count : std::min<uint64_t>(leaf_count, set_leaves);
  leaves_per_chunk = std::max<uint64_t>(granule, leaves_per_chunk / granule * granule);
  // Small domains: one launch and one copy each way serve as many whole keys as fit the set (the leaves of consecutive keys
  // are contiguous in `ys` exactly when a chunk holds whole ranges).  Large domains: one key at a time, range by range.
  size_t kc = 1;
  if (leaves_per_chunk == leaf_count)
    kc = size_t(std::min<uint64_t>(chunk_pref(c, nkeys, 65536), std::max<uint64_t>(1, set_leaves / leaf_count)));
  // key material of the chunk's keys (seeds | cws | ocws) at the front of each set, leaves behind it
  const size_t off_cws = align_up(kc * 16, 256), off_ocws = off_cws + align_up(kc * cwb, 256),
               hdr = off_ocws + align_up(kc * 16, 256);
  const size_t set_bytes = align_up(hdr + std::max<uint64_t>(leaves_per_chunk, kc * leaf_count) * leaf_bytes, 256);
  const bool many = nkeys > kc || leaf_count > leaves_per_chunk;
  ArenaLease L(c->p.device, set_bytes * (many ? 2 : 1), 0);
  if (!L.a) return L.rc;
  Arena &A = *L.a;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += kc) {
    const size_t k = std::min(kc, nkeys - k0);
    for (uint64_t l0 = 0; l0 < leaf_count && !rc; l0 += leaves_per_chunk, ++chunk) {
      const uint64_t cnt = std::min<uint64_t>(leaves_per_chunk, leaf_count - l0);  // (k > 1 only with cnt == leaf_count)
      const int b = int(chunk & 1);
      cudaStream_t s = A.stream[b];
      uint8_t *d_seed = A.dev + size_t(b) * set_bytes, *d_cws = d_seed + off_cws, *d_ocw = d_seed + off_ocws, *d_ys = d_seed + hdr;
      TRY_BREAK(H2D(d_seed, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, s));
      TRY_BREAK(H2D(d_cws, static_cast<const uint8_t *>(cws) + k0 * cwb, k * cwb, s));
      if (half) TRY_BREAK(H2D(d_ocw, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, s));
      rc = fssb200_eval_all(c, party, d_seed, d_cws, half ? d_ocw : nullptr, d_ys, k, leaf_begin + l0, cnt, s);
      if (rc) break;
      TRY_BREAK(D2H(static_cast<uint8_t *>(ys) + (k0 * leaf_count + l0) * leaf_bytes, d_ys, k * cnt * leaf_bytes, s));
    }
  }
  return L.drain(rc);
}

// ---- one process, several GPUs: host arrays of the WHOLE batch -------------------------------------------------------
// Keys are independent (dpf.cuh:170-214); no collective, no peer traffic.  Two ways to spread them:
//  * static: device d takes the contiguous range key_shard(nkeys, d, ndev); one host thread per device runs the
//    pipeline above on its range with its share of the worker threads (1-2 devices: links independent, packing pays);
//  * balanced: when the keys cross in the reference layout (>= 3 devices share the host, or host mode 1) the LINKS
//    differ -- 22.8 ... 34.4 GB/s at 8 GPUs, the host memory system arbitrates -- and an equal split lasts as long as
//    the slowest link needs.  Devices then claim blocks of keys from one counter (two calls in flight per device so a
//    link never waits for a call to drain; blocks shrink towards the end of the batch), and the batch finishes when
//    the host memory system has served all of it: profiles/r02_host_pipeline.md.
int fssb200_eval_host_multi(fssb200_ctx *const *ctxs, int ndev, int party, const void *seeds, const void *cws,
    const void *ocws, const void *xs, void *ys, size_t nkeys, int *rcs) {
  if (!ctxs || ndev < 1 || ndev > 64) return FSSB200_EINVAL;
  for (int d = 0; d < ndev; ++d) {
    if (!ctxs[d]) return FSSB200_EINVAL;
    const fssb200_params &a = ctxs[0]->p, &b = ctxs[d]->p;
    if (a.scheme != b.scheme || a.in_bits != b.in_bits || a.in_bytes != b.in_bytes || a.group != b.group ||
        a.prg != b.prg)
      return FSSB200_EINVAL;  // one key batch = one parameter set
  }
  if (!seeds || !cws || !xs || !ys) return FSSB200_EINVAL;
  const size_t cwb = size_t(ctxs[0]->ncw) * 32, ib = size_t(ctxs[0]->p.in_bytes);
  Crew *crew = Crew::get();
  std::vector<int> rc(size_t(ndev), 0);
  std::vector<std::thread> th;
  auto call = [&](int d, size_t k0, size_t k) {
    return fssb200_eval_host(ctxs[d], party, static_cast<const uint8_t *>(seeds) + k0 * 16,
        static_cast<const uint8_t *>(cws) + k0 * cwb, ocws ? static_cast<const uint8_t *>(ocws) + k0 * 16 : nullptr,
        static_cast<const uint8_t *>(xs) + k0 * ib, static_cast<uint8_t *>(ys) + k0 * 16, k);
  };
  const int mode0 = ctxs[0]->host_mode.load(std::memory_order_relaxed);
  const size_t min_block = size_t(1) << std::min(24, std::max(8, env_int("FSSB200_MULTI_MIN_BLOCK_BITS", 15)));
  const size_t max_block = std::max(min_block, size_t(1) << std::min(28, std::max(8, env_int("FSSB200_MULTI_MAX_BLOCK_BITS", 17))));
  const bool balanced = ndev >= 2 && (mode0 == 1 || (mode0 == 0 && ndev >= 3)) && nkeys >= 2 * min_block * size_t(ndev) &&
      env_int("FSSB200_MULTI_BALANCE", 1) != 0;
  if (balanced) {
    std::mutex mu;
    size_t next = 0;
    auto claim = [&](size_t *k0, size_t *k) {
      std::lock_guard<std::mutex> l(mu);
      if (next >= nkeys) return false;
      const size_t left = nkeys - next;
      size_t b = left / (size_t(4) * size_t(ndev));        // guided: large blocks first, small ones to even out the end
      b = std::min(max_block, std::max(min_block, b)) & ~size_t(255);
      if (left - std::min(left, b) < min_block / 2) b = left;  // no crumbs
      b = std::min(b, left);
      *k0 = next;
      *k = b;
      next += b;
      return true;
    };
    std::vector<std::atomic<int>> first(static_cast<size_t>(ndev));
    for (auto &f : first) f.store(0);
    auto run = [&](int d) {
      t_crew_share = crew ? std::max(1, crew->workers() / (2 * ndev)) : 0;
      t_devices_sharing_host = ndev;
      size_t k0, k;
      while (first[size_t(d)].load(std::memory_order_relaxed) == 0 && claim(&k0, &k)) {
        const int r = call(d, k0, k);
        int zero = 0;
        if (r) first[size_t(d)].compare_exchange_strong(zero, r);
        if (r) {  // the block still has to be evaluated for the batch to be complete: report, do not hide
          std::lock_guard<std::mutex> l(mu);
          next = nkeys;  // stop handing out blocks; the call fails as a whole
        }
      }
      t_crew_share = 0;
      t_devices_sharing_host = 1;
    };
    for (int d = 0; d < ndev; ++d)
      for (int j = 0; j < 2; ++j)
        if (d || j) th.emplace_back(run, d);
    run(0);
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; ++d) rc[size_t(d)] = first[size_t(d)].load();
  } else {
    const int share = crew ? std::max(1, crew->workers() / ndev - (ndev >= 2 ? 1 : 0)) : 0;  // (a CPU per device for its submitter)
    const size_t base = nkeys / size_t(ndev), rem = nkeys % size_t(ndev);
    auto run = [&](int d) {
      const size_t k0 = size_t(d) * base + std::min<size_t>(size_t(d), rem), k = base + (size_t(d) < rem ? 1 : 0);
      t_crew_share = share;
      t_devices_sharing_host = ndev;
      rc[size_t(d)] = call(d, k0, k);
      t_crew_share = 0;
      t_devices_sharing_host = 1;
    };
    for (int d = 1; d < ndev; ++d) th.emplace_back(run, d);
    run(0);
    for (auto &t : th) t.join();
  }
  int first_rc = 0;
  for (int d = 0; d < ndev; ++d) {
    if (rcs) rcs[d] = rc[size_t(d)];
    if (!first_rc && rc[size_t(d)]) first_rc = rc[size_t(d)];
  }
  return first_rc;
}

}  // extern "C"
