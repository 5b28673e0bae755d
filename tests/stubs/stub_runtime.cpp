#include "cuda_runtime.h"
extern "C" {
stub_copy g_stub_log[8192];
int g_stub_n = 0;
char* g_stub_base = nullptr;
int g_stub_fail_at = -1;
void stub_reset(char* base, int fail_at) { g_stub_n = 0; g_stub_base = base; g_stub_fail_at = fail_at; }
int stub_count() { return g_stub_n; }
void stub_get(int i, size_t* off, size_t* bytes) { *off = g_stub_log[i].dst_off_bytes; *bytes = g_stub_log[i].bytes; }
}
