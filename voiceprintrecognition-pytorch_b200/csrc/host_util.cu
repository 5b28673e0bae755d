// Host-side helpers of the C ABI for the staging half of predict_batch (predict.py:244-255): gather a ragged list of
// waveforms into one zero-padded [n, lmax] matrix with a few worker threads, and -- vp_host_stage_h2d -- push it to the
// device slice by slice while the gather of the next slice is still running.  At B=256 x 3 s the batch is 49 MB: one
// Python-level np.zeros + per-row copy would cost more than the whole GPU step.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/vpb200.h"

namespace {

// host_copy.cpp: non-temporal row copy (run-time AVX dispatch) and the store fence that must precede publication
extern "C" int vpb_copy_stream_level();
extern "C" void vpb_copy_stream(float* dst, const float* src, size_t n);
extern "C" void vpb_copy_stream_fence();

std::atomic<int> g_streaming{0};   // vp_host_gather_streaming: 1 = gather with streaming stores

inline void gather_rows(const float* const* srcs, const int32_t* lens, int lmax, float* dst, int r0, int r1, bool streaming) {
  for (int i = r0; i < r1; ++i) {
    float* row = dst + (size_t)i * lmax;
    if (streaming) vpb_copy_stream(row, srcs[i], (size_t)lens[i]);
    else std::memcpy(row, srcs[i], (size_t)lens[i] * sizeof(float));
    if (lens[i] < lmax) std::memset(row + lens[i], 0, (size_t)(lmax - lens[i]) * sizeof(float));
  }
  if (streaming) vpb_copy_stream_fence();   // before the slice's done flag: the copy engine reads these rows next
}

// Persistent worker pool (created on first use, never joined: the workers sleep on a condition variable and die with the
// process).  One job at a time; the job is "slices 0..n_slices-1 of a gather", handed out through an atomic counter.
struct Pool {
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::vector<std::thread> workers;
  // current job
  uint64_t generation = 0;
  int want_workers = 0;
  const float* const* srcs = nullptr;
  const int32_t* lens = nullptr;
  float* dst = nullptr;
  int n = 0, lmax = 0, slice_rows = 0, n_slices = 0;
  bool streaming = false;
  std::atomic<int> next{0};
  std::vector<std::atomic<int>> done;      // per slice: 1 when gathered
  int remaining = 0;                       // participating workers that have not finished the current job yet

  Pool() : done(4096) {}

  void run_slices() {
    for (;;) {
      const int s = next.fetch_add(1, std::memory_order_acq_rel);
      if (s >= n_slices) break;
      const int r0 = s * slice_rows, r1 = std::min(n, r0 + slice_rows);
      gather_rows(srcs, lens, lmax, dst, r0, r1, streaming);
      done[s].store(1, std::memory_order_release);
      { std::lock_guard<std::mutex> lk(mu); }
      cv_done.notify_all();
    }
  }

  std::atomic<uint64_t> gen_hint{0};     // mirrors `generation` for the lock-free spin below

  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      // a predict_batch issues its staging calls back to back: spin ~50 us for the next job before going to sleep on the
      // condition variable (a futex wake-up of an idle core costs 50-100 us, more than gathering a slice)
      for (int spin = 0; spin < 3000 && gen_hint.load(std::memory_order_acquire) == seen; ++spin) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return generation != seen; });
        seen = generation;
        if (id >= want_workers) continue;            // this job uses fewer workers
      }
      run_slices();
      {
        std::lock_guard<std::mutex> lk(mu);          // every participant checks out: the job returns only after all of
        --remaining;                                 // them have left run_slices (a late waker can never see the next job's
      }                                              // parameters half-written)
      cv_done.notify_all();
    }
  }

  void ensure(int nthreads) {
    while ((int)workers.size() < nthreads) {
      const int id = (int)workers.size();
      workers.emplace_back([this, id] { worker(id); });
      workers.back().detach();
    }
  }
};

Pool* pool() {
  static Pool* p = new Pool();     // leaked on purpose: no destructor races at interpreter exit
  return p;
}
std::mutex g_job_mu;               // one staging job at a time per process

}  // namespace

extern "C" int vp_host_gather_streaming(int on) {
  if (on >= 0) g_streaming.store(on ? 1 : 0, std::memory_order_relaxed);
  return (g_streaming.load(std::memory_order_relaxed) != 0 && vpb_copy_stream_level() > 0) ? 1 : 0;
}

extern "C" int vp_host_gather_pad(const float* const* srcs, const int32_t* lens, int32_t n, int32_t lmax, float* dst,
                                  int32_t n_threads) {
  return vp_host_stage_h2d(srcs, lens, n, lmax, dst, nullptr, n > 0 ? (n + std::max(1, n_threads) - 1) / std::max(1, n_threads) : 1,
                           n_threads, nullptr);
}

extern "C" int vp_host_stage_h2d(const float* const* srcs, const int32_t* lens, int32_t n, int32_t lmax, float* staging,
                                 float* device_dst, int32_t slice_rows, int32_t n_threads, void* copy_stream) {
  if (!srcs || !lens || !staging || n < 0 || lmax < 1 || slice_rows < 1) return VP_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (!srcs[i] || lens[i] < 0 || lens[i] > lmax) return VP_ERR_INVALID;
  if (n == 0) return VP_OK;
  const int n_slices = (n + slice_rows - 1) / slice_rows;
  if (n_slices > 4096) return VP_ERR_INVALID;
  const int nt = std::max(1, std::min<int>(std::min<int>(n_threads, n_slices), 32));
  std::lock_guard<std::mutex> job(g_job_mu);
  const bool streaming = g_streaming.load(std::memory_order_relaxed) != 0 && vpb_copy_stream_level() > 0;
  Pool* P = pool();
  {
    std::lock_guard<std::mutex> lk(P->mu);
    P->ensure(nt - 1);
    P->srcs = srcs; P->lens = lens; P->dst = staging; P->n = n; P->lmax = lmax;
    P->slice_rows = slice_rows; P->n_slices = n_slices;
    P->streaming = streaming;
    for (int s = 0; s < n_slices; ++s) P->done[s].store(0, std::memory_order_relaxed);
    P->next.store(0, std::memory_order_release);
    P->want_workers = nt - 1;
    P->remaining = nt - 1;
    ++P->generation;
    P->gen_hint.store(P->generation, std::memory_order_release);
  }
  P->cv_job.notify_all();
  int rc = VP_OK;
  // The calling thread is a gatherer too (n_threads counts it): between two slices of its own it hands every slice that
  // is complete -- in order, so the device side can consume prefixes -- to the copy engine.  With n_threads == 1 the call
  // degenerates to gather slice k / enqueue copy k, still overlapping the H2D of slice k with the gather of slice k+1.
  // Consecutive finished slices go out as ONE copy: when enqueueing is cheap the copies stay fine grained (first bytes on
  // the bus early), when it is slow -- eight ranks of one host in the driver at once: tens of us per call -- more slices
  // finish meanwhile and the copies grow by themselves.
  int issued = 0;
  auto issue_ready = [&]() {
    int j = issued;
    while (j < n_slices && P->done[j].load(std::memory_order_acquire)) ++j;
    if (j == issued) return;
    if (device_dst != nullptr && rc == VP_OK) {
      const int r0 = issued * slice_rows, r1 = std::min(n, j * slice_rows);
      const size_t off = (size_t)r0 * lmax;
      if (cudaMemcpyAsync(device_dst + off, staging + off, (size_t)(r1 - r0) * lmax * sizeof(float), cudaMemcpyHostToDevice,
                          (cudaStream_t)copy_stream) != cudaSuccess) {
        cudaGetLastError();
        rc = VP_ERR_CUDA;
      }
    }
    issued = j;
  };
  for (;;) {
    issue_ready();
    const int s = P->next.fetch_add(1, std::memory_order_acq_rel);
    if (s >= n_slices) break;
    const int r0 = s * slice_rows, r1 = std::min(n, r0 + slice_rows);
    gather_rows(srcs, lens, lmax, staging, r0, r1, streaming);
    P->done[s].store(1, std::memory_order_release);
  }
  while (issued < n_slices) {                          // the workers' last slices
    for (int spin = 0; spin < 200000 && !P->done[issued].load(std::memory_order_acquire); ++spin) {   // ~ms: slices are short
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (!P->done[issued].load(std::memory_order_acquire)) {
      std::unique_lock<std::mutex> lk(P->mu);
      P->cv_done.wait(lk, [&] { return P->done[issued].load(std::memory_order_acquire) != 0; });
    }
    issue_ready();
  }
  {                                                    // the job's memory must not be touched after we return
    std::unique_lock<std::mutex> lk(P->mu);
    P->cv_done.wait(lk, [&] { return P->remaining == 0; });
  }
  return rc;
}
