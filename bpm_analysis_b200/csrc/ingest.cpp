// Host side of the ingest: K0 -- audio_data[::downsample_factor],
// bpm_analysis.py:1033 -- done by the host cores, so that only the kept frames cross PCIe.
//
// The reference decimates FIRST: of a 60-minute 48 kHz recording (345.6 MB of int16) the filter
// only ever sees one frame in 159.  Letting the GPU fetch those frames (a strided copy-engine copy
// with 2-byte rows, or a kernel reading mapped memory) costs one PCIe read request per frame and
// saturates the root complex at ~0.7 G requests/s per GPU, ~1.9 G/s for a whole 8-GPU box.  Here a
// small pool of host threads walks the recording with software prefetch (one cache miss per kept
// frame, many in flight per core) and packs the frames into a pinned staging buffer; one plain
// cudaMemcpyAsync of the packed frames (2.2 MB for the recording above) does the rest.
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bpm_host.h"

namespace {

// cores this process may run on (its affinity mask: a container / NUMA binding is respected)
int usable_cores() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return n;
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return hw == 0 ? 4 : static_cast<int>(hw);
}

// Threads one call uses when the caller does not say: BPM_HOST_THREADS, else this process's share of the
// cores -- with one process per GPU (torchrun sets LOCAL_WORLD_SIZE) the ranks of a box divide the cores
// between them instead of each starting one thread per core (8 ranks x 32 threads on 96 cores ran the
// gather 6x slower than one rank alone).
int default_threads() {
  if (const char* e = std::getenv("BPM_HOST_THREADS")) {
    const int v = std::atoi(e);
    if (v > 0) return v;
  }
  int ranks = 1;
  if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {
    const int v = std::atoi(e);
    if (v > 0) ranks = v;
  }
  const int n = usable_cores() / ranks;
  return n < 1 ? 1 : n;
}

// A fixed pool: the threads are created on first use and sleep on a condition variable between
// jobs (creating 16 threads per recording would cost more than the gather itself).
class Pool {
 public:
  static Pool& get() {
    static Pool* p = new Pool;      // never destroyed: its threads sleep on the condition variable until exit
    return *p;
  }
  int size() const { return static_cast<int>(workers_.size()) + 1; }

  // fn(part, parts) for part = 0 .. parts-1, the caller's thread takes part 0
  void run(int parts, const std::function<void(int, int)>& fn) {
    if (parts <= 1) { fn(0, 1); return; }
    std::unique_lock<std::mutex> job_lock(job_mutex_);            // one job at a time
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn;
      parts_ = parts;
      next_.store(1);
      pending_ = parts - 1;
      ++generation_;
    }
    cv_.notify_all();
    fn(0, parts);
    for (;;) {                                        // then helps with whatever parts are left
      const int p = next_.fetch_add(1);
      if (p >= parts) break;
      fn(p, parts);
      std::lock_guard<std::mutex> g(m_);
      --pending_;
    }
    std::unique_lock<std::mutex> g(m_);
    // every part is done AND every worker that saw this job has left it (fn lives on our stack)
    done_cv_.wait(g, [&] { return pending_ == 0 && active_ == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() {
    int n = usable_cores();
    if (n > 32) n = 32;
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    for (auto& t : workers_) t.detach();
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int, int)>* fn;
      int parts;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return generation_ != seen; });
        seen = generation_;
        fn = fn_;
        parts = parts_;
        if (fn == nullptr) continue;                  // woke up after the job was over
        ++active_;
      }
      for (;;) {
        const int p = next_.fetch_add(1);
        if (p >= parts) break;
        (*fn)(p, parts);
        std::lock_guard<std::mutex> g(m_);
        --pending_;
      }
      std::lock_guard<std::mutex> g(m_);
      if (--active_ == 0 && pending_ == 0) done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, job_mutex_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int parts_ = 0, pending_ = 0, active_ = 0;
  std::atomic<int> next_{0};
  unsigned long generation_ = 0;
};

template <class T>
void gather_range(const char* src, int64_t pitch, char* dst, int64_t j0, int64_t j1) {
  constexpr int64_t AHEAD = 24;                       // frames prefetched ahead of the copy
  const char* s = src + j0 * pitch;
  T* d = reinterpret_cast<T*>(dst) + j0;
  for (int64_t j = j0; j < j1; ++j, s += pitch, ++d) {
    if (j + AHEAD < j1) __builtin_prefetch(s + AHEAD * pitch, 0, 0);
    T v;
    memcpy(&v, s, sizeof(T));
    *d = v;
  }
}

}  // namespace

extern "C" {

int bpm_host_threads(void) {
  const int d = default_threads(), p = Pool::get().size();
  return d < p ? d : p;
}

int bpm_host_gather_frames(const void* pcm, int64_t frame_bytes, int64_t n_frames, int64_t stride, void* out,
                           int n_threads) {
  if (!pcm || !out || frame_bytes < 1 || n_frames < 1 || stride < 1) return BPM_HOST_ERR_ARG;
  const int64_t m = (n_frames + stride - 1) / stride;
  const int64_t pitch = frame_bytes * stride;
  const char* src = static_cast<const char*>(pcm);
  char* dst = static_cast<char*>(out);
  Pool& pool = Pool::get();
  int parts = n_threads <= 0 ? default_threads() : n_threads;
  if (parts > pool.size()) parts = pool.size();
  if (m < 4096 * static_cast<int64_t>(parts)) parts = static_cast<int>(m / 4096) + 1;   // tiny recordings: fewer threads
  const std::function<void(int, int)> body = [&](int part, int nparts) {
    const int64_t j0 = m * part / nparts, j1 = m * (part + 1) / nparts;
    switch (frame_bytes) {
      case 1: gather_range<uint8_t>(src, pitch, dst, j0, j1); break;
      case 2: gather_range<uint16_t>(src, pitch, dst, j0, j1); break;
      case 4: gather_range<uint32_t>(src, pitch, dst, j0, j1); break;
      case 8: gather_range<uint64_t>(src, pitch, dst, j0, j1); break;
      default:
        for (int64_t j = j0; j < j1; ++j) memcpy(dst + j * frame_bytes, src + j * pitch, static_cast<size_t>(frame_bytes));
    }
  };
  pool.run(parts, body);
  return BPM_HOST_OK;
}

// 24-bit PCM (3 bytes per sample, little endian) as scipy.io.wavfile.read hands it to the reference
// (bpm_analysis.py:1014): int32 with the 24 bits in the upper three bytes, i.e. the sample << 8.  Such a
// file cannot be memory-mapped as an array, and scipy's read expands all of it (a 60-minute 48 kHz
// recording: 518 MB in, 691 MB out); here the kept frames are expanded straight out of the mapping.
int bpm_host_gather_s24(const void* pcm, int64_t channels, int64_t n_frames, int64_t stride, int32_t* out,
                        int n_threads) {
  if (!pcm || !out || channels < 1 || n_frames < 1 || stride < 1) return BPM_HOST_ERR_ARG;
  const int64_t m = (n_frames + stride - 1) / stride;
  const int64_t pitch = 3 * channels * stride;
  const unsigned char* src = static_cast<const unsigned char*>(pcm);
  Pool& pool = Pool::get();
  int parts = n_threads <= 0 ? default_threads() : n_threads;
  if (parts > pool.size()) parts = pool.size();
  if (m < 4096 * static_cast<int64_t>(parts)) parts = static_cast<int>(m / 4096) + 1;
  const std::function<void(int, int)> body = [&](int part, int nparts) {
    constexpr int64_t AHEAD = 24;
    const int64_t j0 = m * part / nparts, j1 = m * (part + 1) / nparts;
    const unsigned char* s = src + j0 * pitch;
    int32_t* d = out + j0 * channels;
    for (int64_t j = j0; j < j1; ++j, s += pitch, d += channels) {
      if (stride > 1 && j + AHEAD < j1) __builtin_prefetch(s + AHEAD * pitch, 0, 0);
      for (int64_t c = 0; c < channels; ++c) {
        const unsigned char* b = s + 3 * c;
        const uint32_t u = (static_cast<uint32_t>(b[0]) << 8) | (static_cast<uint32_t>(b[1]) << 16) |
                           (static_cast<uint32_t>(b[2]) << 24);
        d[c] = static_cast<int32_t>(u);
      }
    }
  };
  pool.run(parts, body);
  return BPM_HOST_OK;
}

}  // extern "C"
