/* Test double of the two CUDA runtime entry points host_util.cu uses, so that the staging call's copy-issuing path runs on
 * a machine without a GPU (tests/test_host_logic.py::test_host_stage_copy_path_with_stub_runtime): "device" memory is host
 * memory, a copy is a memcpy, and every call is logged for the test to inspect. */
#pragma once
#include <cstddef>
#include <cstring>

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1 };

extern "C" {
struct stub_copy { size_t dst_off_bytes, bytes; };
extern stub_copy g_stub_log[8192];
extern int g_stub_n;
extern char* g_stub_base;
extern int g_stub_fail_at;     /* >= 0: that call (0-based) returns an error */
}

inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  const int k = g_stub_n++;
  if (k == g_stub_fail_at) return 1;
  if (k < 8192) g_stub_log[k] = {(size_t)((char*)dst - g_stub_base), bytes};
  std::memcpy(dst, src, bytes);
  return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
