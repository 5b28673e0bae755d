// Streaming (non-temporal) row copy for the staging gather of host_util.cu.  A plain C++ translation unit (no CUDA front
// end) so that the AVX intrinsics are compiled by the host compiler alone; dispatch is at run time, nothing here needs
// -mavx* on the command line.
//
// Why: when several ranks of one host stage their shards at the same time the gather is DRAM-bound (8 ranks x 49 MB on
// a two-socket box: 2.7 ms per rank against 0.56 ms alone, tools/e2e_multi.py).  A regular store first reads the
// destination line for ownership; the pinned staging rows are never read back by a CPU -- the copy engine DMAs them --
// so streaming stores cut the gather's DRAM traffic from 3 to 2 bytes per byte copied.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

namespace {

__attribute__((target("avx512f"))) void copy_stream_512(float* dst, const float* src, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; ++i; }
  for (; i + 64 <= n; i += 64) {
    const __m512 a = _mm512_loadu_ps(src + i), b = _mm512_loadu_ps(src + i + 16);
    const __m512 c = _mm512_loadu_ps(src + i + 32), d = _mm512_loadu_ps(src + i + 48);
    _mm512_stream_ps(dst + i, a);
    _mm512_stream_ps(dst + i + 16, b);
    _mm512_stream_ps(dst + i + 32, c);
    _mm512_stream_ps(dst + i + 48, d);
  }
  for (; i < n; ++i) dst[i] = src[i];
}

__attribute__((target("avx2"))) void copy_stream_256(float* dst, const float* src, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
  for (; i + 32 <= n; i += 32) {
    const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
    const __m256 c = _mm256_loadu_ps(src + i + 16), d = _mm256_loadu_ps(src + i + 24);
    _mm256_stream_ps(dst + i, a);
    _mm256_stream_ps(dst + i + 8, b);
    _mm256_stream_ps(dst + i + 16, c);
    _mm256_stream_ps(dst + i + 24, d);
  }
  for (; i < n; ++i) dst[i] = src[i];
}

int detect() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return 2;
  if (__builtin_cpu_supports("avx2")) return 1;
  return 0;
}

}  // namespace

// 0: no vector streaming stores on this CPU (the callers fall back to memcpy), 1: AVX2, 2: AVX-512
extern "C" int vpb_copy_stream_level() {
  static const int level = detect();
  return level;
}

// dst[0:n] = src[0:n] with non-temporal stores; dst must be 4-byte aligned (a float row), src anything.  The caller issues
// vpb_copy_stream_fence() before it publishes the rows to another agent (thread or DMA engine).
extern "C" void vpb_copy_stream(float* dst, const float* src, size_t n) {
  switch (vpb_copy_stream_level()) {
    case 2: copy_stream_512(dst, src, n); break;
    case 1: copy_stream_256(dst, src, n); break;
    default: std::memcpy(dst, src, n * sizeof(float));
  }
}

extern "C" void vpb_copy_stream_fence() { _mm_sfence(); }

#else

extern "C" int vpb_copy_stream_level() { return 0; }
extern "C" void vpb_copy_stream(float* dst, const float* src, size_t n) { std::memcpy(dst, src, n * sizeof(float)); }
extern "C" void vpb_copy_stream_fence() {}

#endif
