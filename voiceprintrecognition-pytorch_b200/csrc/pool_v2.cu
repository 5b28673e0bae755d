// One-trip cp.async staging variants of the pooling kernels (default since round 2; VPB_POOL_V2=0 disables them).
//
// Same contracts as asp_pool_smem_kernel / colstats_kernel (pool.cu), different data movement: the whole [rows, 32]
// channel strip of an utterance is brought into shared memory with cp.async (LDGSTS, 16 B per request), ALL requests of
// the CTA in flight at once, instead of register-staged rounds of 8 loads per thread.  Round-1 measurement that
// motivates it: asp_pool reaches 2.25 TB/s and colstats 3.5 TB/s of the 6.5 TB/s copy peak; both are
// bound by the load round trips per CTA, not by bandwidth.  The arithmetic after staging is identical to the v1
// kernels (same order of operations), so results must match them bit for bit.
#include "kernels.cuh"

namespace vpb {

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;                       // src-size 0: the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stage rows [r0, r1) of a [.., ld] matrix, 32 columns starting at column c0 (only columns < C are read), into
// dst[(r - r0) * 32 + col]
__device__ __forceinline__ void stage_strip(float* dst, const float* src, int ld, int r0, int r1, int c0, int C) {
  const int c4 = (threadIdx.x & 7) * 4;
  const bool cok = c0 + c4 < C;                           // C % 4 == 0 on this path
  const float* g = src + c0 + c4;
  for (int r = r0 + (threadIdx.x >> 3); r < r1; r += blockDim.x >> 3)
    cp_async16(dst + (r - r0) * 32 + c4, cok ? g + (size_t)r * ld : src, cok);
}

__global__ void __launch_bounds__(256) asp_pool_v2_kernel(const __grid_constant__ AspParams p) {
  pdl_launch_dependents();
  pdl_wait();                 // first access to mutable global memory comes after this
  extern __shared__ __align__(16) float sm2[];
  __shared__ float red[8][33];
  __shared__ float bc[32];
  float* sx = sm2;                       // [T][32]
  float* sl = sm2 + (size_t)p.T * 32;    // [T][32]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32;
  const int c = c0 + lane;
  const int b = blockIdx.y;
  const bool ok = c < p.C;
  stage_strip(sx, p.x + (size_t)b * p.T * p.x_ld + p.x_coff, p.x_ld, 0, p.T, c0, p.C);
  stage_strip(sl, p.logit + (size_t)b * p.T * p.l_ld + p.l_coff, p.l_ld, 0, p.T, c0, p.C);
  cp_async_wait_all();
  __syncthreads();
  auto block_reduce = [&](float v, bool is_max) -> float {
    red[wid][lane] = v;
    __syncthreads();
    if (wid == 0) {
      float s = red[0][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) s = is_max ? fmaxf(s, red[i][lane]) : s + red[i][lane];
      bc[lane] = s;
    }
    __syncthreads();
    float r = bc[lane];
    __syncthreads();
    return r;
  };
  float mx = -INFINITY;
  for (int t = wid; t < p.T; t += 8) mx = fmaxf(mx, sl[t * 32 + lane]);
  mx = block_reduce(mx, true);
  float se = 0.f, sxe = 0.f;
  for (int t = wid; t < p.T; t += 8) {
    const float e = expf(sl[t * 32 + lane] - mx);
    sl[t * 32 + lane] = e;
    se += e;
    sxe = fmaf(e, sx[t * 32 + lane], sxe);
  }
  se = block_reduce(se, false);
  sxe = block_reduce(sxe, false);
  const float mean = sxe / se;
  float sq = 0.f;
  for (int t = wid; t < p.T; t += 8) {
    const float d = sx[t * 32 + lane] - mean;
    sq = fmaf(sl[t * 32 + lane], d * d, sq);
  }
  sq = block_reduce(sq, false);
  if (wid == 0 && ok) {
    float* o = p.dst + (size_t)b * p.out_ld + p.out_coff;
    o[c] = mean;
    if (!p.mean_only) o[p.C + c] = sqrtf(fmaxf(sq / se, p.eps));
  }
}

bool asp_pool_v2_supported(const AspParams& p) {
  return (size_t)p.T * 32 * 2 * sizeof(float) <= 100 * 1024 &&       // two CTAs per SM keep loads and sweeps overlapped
         (p.C & 3) == 0 && (p.x_ld & 3) == 0 && (p.x_coff & 3) == 0 && (p.l_ld & 3) == 0 && (p.l_coff & 3) == 0;
}

cudaError_t launch_asp_pool_v2(const AspParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)p.T * 32 * 2 * sizeof(float);
  static PerDeviceSmem once;
  if (once.need(100 * 1024)) {
    cudaError_t e = cudaFuncSetAttribute(asp_pool_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    once.set(100 * 1024);
  }
  dim3 grid((p.C + 31) / 32, p.B);
  launch_pdl(asp_pool_v2_kernel, grid, 256, smem, stream, p);
  return cudaGetLastError();
}

// colstats for the modes whose whole [R, 32] strip fits in shared memory (1-D maps): one trip to HBM, then the same
// two-pass mean / centred sum of squares as colstats_kernel, out of shared memory.
__global__ void __launch_bounds__(256) colstats_v2_kernel(const __grid_constant__ StatsParams p) {
  pdl_launch_dependents();
  pdl_wait();                 // first access to mutable global memory comes after this
  extern __shared__ __align__(16) float sm2[];
  __shared__ float red[8][33];
  __shared__ float bc[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32;
  const int c = c0 + lane;
  const int b = blockIdx.y;
  const bool ok = c < p.C;
  stage_strip(sm2, p.src + (size_t)b * p.R * p.in_ld + p.in_coff, p.in_ld, 0, p.R, c0, p.C);
  cp_async_wait_all();
  __syncthreads();
  auto block_sum = [&](float v) -> float {
    red[wid][lane] = v;
    __syncthreads();
    if (wid == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][lane];
      bc[lane] = s;
    }
    __syncthreads();
    return bc[lane];
  };
  float v = 0.f;
  for (int r = wid; r < p.R; r += 8) v += sm2[r * 32 + lane];
  const float mean = block_sum(v) / (float)p.R;
  float* o = p.dst + (size_t)b * p.out_ld + p.out_coff;
  if (p.mode == VP_STATS_MEAN) {
    if (wid == 0 && ok) o[c] = mean;
    return;
  }
  __syncthreads();
  float q = 0.f;
  for (int r = wid; r < p.R; r += 8) { float d = sm2[r * 32 + lane] - mean; q = fmaf(d, d, q); }
  const float ssq = block_sum(q);
  if (wid == 0 && ok) {
    float sd;
    if (p.mode == VP_STATS_MEAN_STD_CLAMP) sd = sqrtf(fmaxf(ssq / (float)p.R, p.eps));
    else if (p.mode == VP_STATS_MEAN_VAR_UNBIASED) sd = ssq / (float)(p.R - 1);
    else if (p.mode == VP_STATS_MEAN_STD_UNBIASED) sd = sqrtf(ssq / (float)(p.R - 1));
    else sd = sqrtf(ssq / (float)(p.R - 1) + p.eps);
    o[c] = mean;
    o[p.C + c] = sd;
  }
}

bool colstats_v2_supported(const StatsParams& p) {
  return p.mode != VP_STATS_SEG_CONTEXT && (size_t)p.R * 32 * sizeof(float) <= 100 * 1024 && (p.C & 3) == 0 &&
         (p.in_ld & 3) == 0 && (p.in_coff & 3) == 0;
}

cudaError_t launch_colstats_v2(const StatsParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)p.R * 32 * sizeof(float);
  static PerDeviceSmem once;
  if (once.need(100 * 1024)) {
    cudaError_t e = cudaFuncSetAttribute(colstats_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    once.set(100 * 1024);
  }
  dim3 grid((p.C + 31) / 32, p.B);
  launch_pdl(colstats_v2_kernel, grid, 256, smem, stream, p);
  return cudaGetLastError();
}

}  // namespace vpb
