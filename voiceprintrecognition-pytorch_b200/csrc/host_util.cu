// Host-side helper of the C ABI: multi-threaded gather of a ragged list of waveforms into one zero-padded [n, lmax]
// staging matrix (the np.zeros + per-item copy loop of predict.py:248-254, which at B=256 x 3 s moves 49 MB and would
// otherwise dominate the end-to-end time of predict_batch).  Plain C++ threads + memcpy; no CUDA calls.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/vpb200.h"

extern "C" int vp_host_gather_pad(const float* const* srcs, const int32_t* lens, int32_t n, int32_t lmax, float* dst,
                                  int32_t n_threads) {
  if (!srcs || !lens || !dst || n < 0 || lmax < 1) return VP_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (!srcs[i] || lens[i] < 0 || lens[i] > lmax) return VP_ERR_INVALID;
  auto work = [&](int r0, int r1) {
    for (int i = r0; i < r1; ++i) {
      float* row = dst + (size_t)i * lmax;
      std::memcpy(row, srcs[i], (size_t)lens[i] * sizeof(float));
      if (lens[i] < lmax) std::memset(row + lens[i], 0, (size_t)(lmax - lens[i]) * sizeof(float));
    }
  };
  int nt = std::max(1, std::min<int>(n_threads, n));
  if (nt == 1) {
    work(0, n);
    return VP_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  const int step = (n + nt - 1) / nt;
  for (int t = 1; t < nt; ++t) {
    const int r0 = t * step, r1 = std::min(n, r0 + step);
    if (r0 < r1) pool.emplace_back(work, r0, r1);
  }
  work(0, std::min(n, step));
  for (auto& th : pool) th.join();
  return VP_OK;
}
