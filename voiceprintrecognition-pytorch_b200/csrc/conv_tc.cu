// tcgen05 engine: implicit-GEMM convolution on the 5th-gen tensor cores with fp32-grade accuracy (split TF32, "3xTF32").
//
//   out[m, n] = epi( sum_k A(m, k) * W[n, k] )        same op contract as conv_ffma.cu (common.cuh::ConvParams)
//
// Why split TF32: a single TF32 (or BF16) pass breaks the 1e-4 embedding parity gate (3e-4 / 2.7e-3),
// so every operand is split  x = hi + lo,  hi = tf32(x),  lo = x - hi  and three MMAs are accumulated in fp32 TMEM:
//   D += A_lo * B_hi;  D += A_hi * B_lo;  D += A_hi * B_hi        (the dropped lo*lo term is ~2^-22 relative).
//
// Persistent, warp-specialised CTA (448 threads), one CTA per SM:
//   warps 0-3  epilogue   : tcgen05.ld accumulator rows from TMEM -> fused epilogue (bias / ubias / act / BN affine /
//                           gate / residual / act2) -> float4 stores; TMEM accumulators are double buffered so the
//                           epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-11 A producers: gather the implicit-GEMM A tile straight from the channel-last activation map (any tap /
//                           dilation / stride / reflect or zero padding / add or concat second source / BN-ReLU
//                           prologue: common.cuh::gather_a4) with coalesced 128 B row segments, split into hi / lo and
//                           write both in the UMMA K-major SWIZZLE_128B shared-memory layout
//   warp 12    B loader   : weights are pre-split and pre-tiled on the host in exactly that shared-memory image, so one
//                           bulk-async copy (TMA engine, cp.async.bulk -> UBLKCP) per stage lands B_hi|B_lo
//   warp 13    MMA issuer : one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), commits to mbarriers
// Pipeline: S shared-memory stages (full/empty mbarriers) + 2 TMEM accumulator buffers (tmem_full/tmem_empty).
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"

namespace vpb {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                  // fp32 elements per stage row = 128 B = one SWIZZLE_128B row
constexpr int A_TILE = BM * BK * 4;     // 16 KB per (hi|lo) A tile
#ifndef VPB_EPI_WARPS
#define VPB_EPI_WARPS 4
#endif
// Epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, each taking every other 32-column chunk).  With
// 128 x 256 tiles four warps were the pacing role (ncu, round 2: the MMA warp waited on tmem_empty); the layers whose
// tiles are short in K are pure output streaming and need the store parallelism.
constexpr int EPI_WARPS = VPB_EPI_WARPS;
static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue warps: one or two per TMEM lane quadrant");
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PRODUCER_WARPS = 8;
constexpr int PRODUCER_THREADS = PRODUCER_WARPS * 32;      // 256
constexpr int NUM_THREADS = EPI_THREADS + PRODUCER_THREADS + 64;   // epilogue + 8 producer + loader + MMA warps
constexpr int ROWS_PER_THREAD = BM * 8 / PRODUCER_THREADS;   // 4
constexpr int SMEM_BUDGET = 193 * 1024;            // pipeline stages (+ 32 KB epilogue pads + 1 KB alignment <= 227 KB)
constexpr int EPI_PAD_BYTES = EPI_WARPS * 32 * 32 * 4;      // per epilogue warp: a 32 x 32-float transpose pad (XOR-swizzled chunks)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, kind::f16 (fp16 operands, UMMA_K = 16, fp32 accumulate) -- experimental two-term FP16 split path
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 (16 B) |
// SBO = 1024 B (8 rows x 128 B) | version 1 (sm_100) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __noinline__ float4 act4_slow(float4 v, int act) {
  v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
  return v;
}

template <int MODE> struct Slot;
template <> struct Slot<0> { float4 v[ROWS_PER_THREAD]; };
template <> struct Slot<1> { float4 v[ROWS_PER_THREAD]; float4 u[ROWS_PER_THREAD]; };
template <> struct Slot<2> { float4 v[ROWS_PER_THREAD]; float4 ps, ph; uint32_t okmask; };
struct SlotH { float4 v[ROWS_PER_THREAD]; float4 w[ROWS_PER_THREAD]; };   // F16 path: 8 consecutive K elements per row

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
      "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
      "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
      "f"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct TcArgs {
  const float* w_tc;     // pre-split, pre-tiled, pre-swizzled weights: [n_tile][k_block][hi BN x 128 B | lo BN x 128 B]
  int BN;                // N tile (multiple of 16, <= 256)
  int stages;
  int tmem_cols;         // power of two >= 2*BN
  int m_tiles, n_tiles, k_blocks;
  int kc, n_chunks;      // K blocks per accumulation chunk / chunks per tile (long-K layers: bounded accumulation length)
  int debug;             // experiments only (VPB_TC_DEBUG): 1 = skip B copies, 2 = skip A global loads, 4 = skip MMAs
  float descale;         // F16 path only: 1 / (power-of-two scale folded into the fp16 weight image); keep LAST
};

// F16 = true (experimental, opt-in via VPB_TC_F16=1, MODE 0 only): operands are split into two fp16 terms instead of
// two tf32 terms -- hi = fp16(x), lo = fp16(x - hi), same three MMAs, kind::f16 at twice the tf32 issue rate; a stage
// row of 128 B then holds 64 K elements (BKE) and a producer thread moves 8 of them per row (KPT).
// (A cp.async shared-memory ring for the A producers was built and measured in round 2: parity-green, 5.54 vs 5.56 ms per
// ECAPA step -- no gain, removed.  profiles/r2_staged_variants.md)
template <int MODE, bool F16 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p,
                                                                 const __grid_constant__ TcArgs a) {
  static_assert(!F16 || MODE == 0, "the fp16 split path only implements the plain / concat source mode");
  constexpr int BKE = F16 ? 64 : BK;     // K elements per pipeline stage
  constexpr int KPT = F16 ? 8 : 4;       // K elements per producer thread and row
  using SlotT = typename std::conditional<F16, SlotH, Slot<MODE>>::type;
  pdl_launch_dependents();                 // the next kernel's CTAs may be scheduled as SMs free up (see common.cuh)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t bars[2 * 8 + 4];     // full[S], empty[S], tmem_full[2], tmem_empty[2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int BN = a.BN;
  const uint32_t b_tile = (uint32_t)BN * 128u;                 // bytes of one (hi|lo) B tile
  const uint32_t stage_bytes = 2u * A_TILE + 2u * b_tile;
  const int S = a.stages;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
  const uint32_t tfull0 = smem_u32(&bars[16]), tempty0 = smem_u32(&bars[18]);
  // Thread-block cluster of C CTAs working on C consecutive M tiles of the SAME N tile: every CTA fetches 1/C of each B
  // stage and multicasts it to all C shared memories (L2 -> SM weight traffic / C); a stage is recycled only when the
  // MMAs of all C CTAs have consumed it (multicast tcgen05.commit on every CTA's `empty` barrier).
  const uint32_t C = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, PRODUCER_WARPS + 1);            // one arrival per A-producer warp + 1 expect_tx arrival (B)
      mbar_init(empty0 + 8 * s, C);                            // one tcgen05.commit from every CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, EPI_THREADS);                 // every epilogue thread arrives once per accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();           // peers' mbarriers are initialised before any remote arrive / multicast lands
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // barrier init, TMEM allocation and the cluster handshake above overlap the tail of the previous kernel; everything
  // below touches activations / amax slots / the output
  pdl_wait();

  // work distribution: group g = (m_group, n_tile); CTA `crank` of the cluster takes M tile m_group*C + crank
  const int n_clusters = gridDim.x / C, cluster_id = blockIdx.x / C;
  const int m_groups = (a.m_tiles + C - 1) / C;
  const int total_groups = m_groups * a.n_tiles;

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PRODUCER_WARPS) {
    // =========================== A producers ===========================
    // Thread t owns the 16-byte chunk `chunk` of rows r0 + 32*i: a warp-level load covers 4 rows x 128 contiguous bytes.
    // All tap / channel arithmetic is per K block (uniform over the thread's rows); per row only the bounds test remains.
    const int t = threadIdx.x - EPI_THREADS;
    const int chunk = t & 7;
    const int r0 = t >> 3;                 // 0..31
    const bool pointwise = (p.KT * p.KF == 1);
    // The (tile, K block) pairs of this CTA form ONE flat stream: the gather cursor runs DEPTH items ahead of the publish
    // cursor and crosses tile boundaries, so the load pipeline never drains between tiles.
    const int my_tiles = (total_groups - cluster_id + n_clusters - 1) / n_clusters;
    const int total_items = my_tiles * a.k_blocks;
    RowInfo rows[ROWS_PER_THREAD];
    // Gather cursor.  gather() is always called for consecutive items q = 0, 1, 2, ..., so the (tile, K block, tap,
    // channel) decomposition of q is kept incrementally -- no integer division per K block (ncu, Res2 convs: the old
    // per-call divisions + 64-bit row arithmetic made the producers instruction-issue bound at N = 64).
    int g_tl = 0, g_kb = 0;                    // tile (local index) and K block of the next item to gather
    int g_k = chunk * KPT;                     // this thread's first k of that K block
    int g_ci = 0, g_kt = 0, g_kf = 0;          // channel / tap position of g_k
    auto cursor_reset = [&]() {
      g_kb = 0; g_k = chunk * KPT; g_ci = g_k; g_kt = 0; g_kf = 0;
      if (!pointwise)
        while (g_ci >= p.CinTot) { g_ci -= p.CinTot; if (++g_kf == p.KF) { g_kf = 0; ++g_kt; } }
    };
    cursor_reset();
    {
      // Register prefetch slots.  gather() ONLY issues loads (no instruction may read a loaded register before publish():
      // even a predicated-off consumer stalls on the load's scoreboard and would serialise the loads); add / BN-ReLU
      // prologue are applied in publish().  MODE 0: plain or channel-concat source, 3 K blocks in flight; MODE 1: second
      // source added (x_i + y_{i-1}); MODE 2: per-channel affine(+ReLU) prologue; 2 K blocks in flight for 1 and 2.
      SlotT sl0, sl1, sl2;
      auto gather = [&](SlotT& sl) {
        if (g_kb == 0) {                       // cursor entered a new tile: decode its 4 rows
          const int g = cluster_id + g_tl * n_clusters;
          const int m0 = ((g / a.n_tiles) * (int)C + (int)crank) * BM;
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            rows[i] = decode_row(p, m0 + r0 + 32 * i);
            if (!rows[i].valid) rows[i].t0 = -(1 << 28);        // fails every bounds test below (also after reflection)
          }
        }
        const int ci = g_ci;
        const int dt = g_kt * p.dT, df = g_kf * p.dF;
        const bool kok = g_k < p.K && !(a.debug & 2);
        const bool second = (MODE == 0) && (p.src2_mode == VP_SRC2_CONCAT) && (ci >= p.Cin);
        const char* base = reinterpret_cast<const char*>(second ? p.src2 + p.src2_coff + (ci - p.Cin) : p.src + p.in_coff + ci);
        const uint32_t ldb = (uint32_t)(second ? p.src2_ld : p.in_ld) * 4u;      // row pitch in bytes
        if constexpr (MODE == 2) {
          sl.ps = make_float4(1.f, 1.f, 1.f, 1.f);
          sl.ph = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kok) {
            sl.ps = __ldg(reinterpret_cast<const float4*>(p.pre_s + ci));
            sl.ph = __ldg(reinterpret_cast<const float4*>(p.pre_h + ci));
          }
          sl.okmask = 0;
        }
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          int ti = rows[i].t0 + dt;
          const int fi = rows[i].f0 + df;
          if (p.pad_mode == VP_PAD_REFLECT) {
            if (ti < 0) ti = -ti;
            if (ti >= p.Tin) ti = 2 * (p.Tin - 1) - ti;
          }
          const bool ok = kok && (unsigned)ti < (unsigned)p.Tin && (unsigned)fi < (unsigned)p.Fin;
          const uint32_t r = (uint32_t)(rows[i].base + ti * p.Fin + fi);       // source row index (< 2^31 rows)
          sl.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) sl.v[i] = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)r * ldb));
          if constexpr (F16) {
            sl.w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) sl.w[i] = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)r * ldb + 16));
          }
          if constexpr (MODE == 1) {
            sl.u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok)
              sl.u[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(p.src2 + p.src2_coff + ci) +
                                                              (uint64_t)r * ((uint32_t)p.src2_ld * 4u)));
          }
          if constexpr (MODE == 2) sl.okmask |= ok ? (1u << i) : 0u;
        }
        // advance the cursor by one K block
        if (++g_kb == a.k_blocks) {
          ++g_tl;
          cursor_reset();
        } else {
          g_k += BKE;
          g_ci += BKE;
          if (!pointwise)
            while (g_ci >= p.CinTot) { g_ci -= p.CinTot; if (++g_kf == p.KF) { g_kf = 0; ++g_kt; } }
        }
      };
      constexpr int DEPTH = (MODE == 0 && !F16) ? 3 : 2;
      int p_s = 0;
      uint32_t p_ph = 0;
      // fp16 split: dynamic power-of-two activation scale from the source tensor's amax slot (written by the ops that
      // produced the tensor, complete at kernel start) -> |x * ascale| < 2^14, range-safe for any activation magnitude
      float ascale = 1.f;
      if constexpr (F16) ascale = exp2i(f16_scale_exp(__ldg(p.amax_in)));
      auto publish = [&](int q, SlotT& sl) {
        const int s = p_s;
        const uint32_t ph = p_ph;
        if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t a_hi = smem_base + s * stage_bytes;
        const uint32_t a_lo = a_hi + A_TILE;
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          const int r = r0 + 32 * i;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
          if constexpr (F16) {
            // 8 consecutive K elements -> one 16-byte chunk of fp16 hi and one of fp16 lo (lo = x - hi, exact in fp32)
            const float xs[8] = {sl.v[i].x * ascale, sl.v[i].y * ascale, sl.v[i].z * ascale, sl.v[i].w * ascale,
                                 sl.w[i].x * ascale, sl.w[i].y * ascale, sl.w[i].z * ascale, sl.w[i].w * ascale};
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __half2 h = __floats2half2_rn(xs[2 * j], xs[2 * j + 1]);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(xs[2 * j] - hf.x, xs[2 * j + 1] - hf.y);
              hw[j] = *reinterpret_cast<const uint32_t*>(&h);
              lw[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
            continue;
          }
          float4 x = sl.v[i];
          if constexpr (MODE == 1) { x.x += sl.u[i].x; x.y += sl.u[i].y; x.z += sl.u[i].z; x.w += sl.u[i].w; }
          if constexpr (MODE == 2) {
            if (sl.okmask & (1u << i)) {              // zero padding stays zero: the conv pads the BN-ReLU'd map
              x.x = fmaf(x.x, sl.ps.x, sl.ph.x); x.y = fmaf(x.y, sl.ps.y, sl.ph.y);
              x.z = fmaf(x.z, sl.ps.z, sl.ph.z); x.w = fmaf(x.w, sl.ps.w, sl.ph.w);
              if (p.pre_relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            }
          }
          float4 hi, lo;
          hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
          lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
        }
        fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * s);     // one arrival per producer warp
        if (q + DEPTH < total_items) gather(sl);   // refill this slot (item q + DEPTH): DEPTH K blocks of loads stay in flight
      };
      if (total_items > 0) gather(sl0);
      if (total_items > 1) gather(sl1);
      if constexpr (DEPTH == 3) {
        if (total_items > 2) gather(sl2);
        for (int q = 0; q < total_items; q += 3) {
          publish(q, sl0);
          if (q + 1 < total_items) publish(q + 1, sl1);
          if (q + 2 < total_items) publish(q + 2, sl2);
        }
      } else {
        for (int q = 0; q < total_items; q += 2) {
          publish(q, sl0);
          if (q + 1 < total_items) publish(q + 1, sl1);
        }
      }
    }
  } else if (warp == EPI_WARPS + PRODUCER_WARPS) {
    // =========================== B loader (bulk async copy) ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;                                        // stage / phase cursors, advanced without divisions
      const uint32_t slice = 2u * b_tile / C;                 // this CTA's share of every B stage
      for (int g = cluster_id; g < total_groups; g += n_clusters) {
        const int nt = g % a.n_tiles;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_tc) + (size_t)nt * a.k_blocks * (2u * b_tile) + crank * slice;
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          if (a.debug & 1) {
            mbar_arrive(full0 + 8 * s);
          } else {
            mbar_expect_tx(full0 + 8 * s, 2u * b_tile);        // the whole stage lands here (own slice + peers' multicasts)
            const uint32_t dst = smem_base + s * stage_bytes + 2u * A_TILE + crank * slice;
            if (C > 1) bulk_copy_g2s_mcast(dst, src + (size_t)kb * (2u * b_tile), slice, full0 + 8 * s, cmask);
            else bulk_copy_g2s(dst, src + (size_t)kb * (2u * b_tile), slice, full0 + 8 * s);
          }
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1 + PRODUCER_WARPS) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // kind::tf32 instruction descriptor: D=F32 (bit 4), A=B=TF32 (2 at bits 7 and 10), K-major, N>>3 @17, M>>4 @24
      // (kind::f16: A = B = F16 is format code 0 at bits 7 and 10)
      const uint32_t fmt = F16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t ccount = 0;
      int s = 0;
      uint32_t sph = 0;                                       // stage / phase cursors, advanced without divisions
      for (int g = cluster_id; g < total_groups; g += n_clusters) {
        // Long-K layers are accumulated in chunks of `kc` K blocks, each into a fresh TMEM accumulator; the epilogue
        // warps fold the chunks into a running fp32 sum with correctly rounded adds.  (The tensor core truncates on every
        // accumulate; the bias grows linearly with the number of accumulates: 4.8e-5 at K=1536, measured.)
        for (int ch = 0; ch < a.n_chunks; ++ch, ++ccount) {
          const int acc = ccount & 1;
          mbar_wait(tempty0 + 8 * acc, ((ccount >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(acc * BN);
          const int kb0 = ch * a.kc;
          const int kb1 = (kb0 + a.kc < a.k_blocks) ? kb0 + a.kc : a.k_blocks;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full0 + 8 * s, sph);
            tc_fence_after();
            const uint32_t a_hi = smem_base + s * stage_bytes, a_lo = a_hi + A_TILE;
            const uint32_t b_hi = a_hi + 2u * A_TILE, b_lo = b_hi + b_tile;
            if (!(a.debug & 4))
#pragma unroll
            for (int kc = 0; kc < BK / 8; ++kc) {             // UMMA_K = 8 for tf32 = 32 bytes inside the swizzle row
              const uint64_t dah = smem_desc(a_hi + kc * 32), dal = smem_desc(a_lo + kc * 32);
              const uint64_t dbh = smem_desc(b_hi + kc * 32), dbl = smem_desc(b_lo + kc * 32);
              if constexpr (F16) {                              // UMMA_K = 16 halves = the same 32 bytes of the row
                umma_f16(d, dal, dbh, idesc, (kb != kb0) || (kc != 0));
                umma_f16(d, dah, dbl, idesc, 1);
                umma_f16(d, dah, dbh, idesc, 1);
              } else {
                umma_tf32(d, dal, dbh, idesc, (kb != kb0) || (kc != 0));
                umma_tf32(d, dah, dbl, idesc, 1);
                umma_tf32(d, dah, dbh, idesc, 1);
              }
            }
            if (C > 1) umma_commit_mcast(empty0 + 8 * s, cmask);   // frees the stage in every CTA of the cluster
            else umma_commit(empty0 + 8 * s);                 // frees the smem stage when these MMAs retire
            if (++s == S) { s = 0; sph ^= 1u; }
          }
          umma_commit(tfull0 + 8 * acc);                      // chunk accumulator ready for the epilogue warps
        }
      }
    }
  } else {
    // =========================== epilogue (warps 0-3 <-> TMEM lanes 32*warp ..) ===========================
    // Per 32-column chunk: tcgen05.ld gives each thread (= accumulator row) 32 consecutive columns; the chunk is
    // transposed through a private 32 x 32-float swizzled shared-memory pad so that the fused epilogue and the global stores run
    // with lanes along N: every store/residual instruction covers 4 rows x 128 contiguous bytes, and the per-column
    // parameters are one float4 per chunk.
    // pad: [32 rows][8 chunks of 4 floats], chunk position XOR-swizzled by (row & 7): row-per-thread float4 writes and
    // column-per-lane float4 reads are both bank-conflict free at 128 bytes per row (no padding column)
    float* pad = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)S * stage_bytes) + warp * (32 * 32);
    const uint32_t pad_u32 = smem_base + S * stage_bytes + warp * (32 * 32 * 4);
    const int quad = warp & 3;               // TMEM lane quadrant (rows 32 * quad .. of the tile)
    const int chalf = warp >> 2;             // with 8 epilogue warps: this warp takes the chunks c0 / 32 == chalf (mod 2)
    constexpr int CSTEP = 32 * (EPI_WARPS / 4);
    const int cg = (lane & 7) * 4;           // this thread's 4 columns inside the chunk
    const int rsub = lane >> 3;              // rows rsub + 4*i
    const bool need_urow = (p.gate != nullptr) || (p.ubias != nullptr);
    // The transcendental activations (tanh / sigmoid / SiLU / hardtanh) live in ONE out-of-line function: inlined they
    // were ~100 instructions per element x 32 elements x 2 call sites = more than half of the kernel's 11 k SASS
    // instructions, and ncu showed `stall_no_instruction` (instruction-cache misses) at 1.4 per issued instruction.
    uint32_t ccount = 0;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t run_col = (uint32_t)(2 * BN);   // running-sum accumulator (only when n_chunks > 1)
    float descale = 1.f;                           // fp16 split: 2^-k of the weight image x 2^-s of the activation scale
    if constexpr (F16) descale = a.descale * exp2i(-f16_scale_exp(__ldg(p.amax_in)));
    float tmax = 0.f;                              // running max |y| of everything this thread stores (amax_out slot)
    // Fast path (ncu, round 2: with 128 x 256 tiles the four epilogue warps, ~370 dependent instructions per 32-column
    // chunk, were the pacing role -- the MMA warp spent its time waiting for tmem_empty).  The common epilogue shape
    // -- bias, optional ReLU, optional BN affine, nothing per utterance, no residual -- on a tile without row / column
    // tails needs no predicates, one pointer per thread and three FP instructions per element.
    // ReLU and the ERes2Net family's ReLU(20) = Hardtanh(0, 20) (eres2net.py:12-17) are both "clamp to [0, hi]" with
    // hi = +inf / 20: min(max(x, 0), hi), the same expression and order as apply_act.  (ncu, round 2: with Hardtanh left to
    // the general path every conv of the 55 M ERes2Net ran the rolled row-at-a-time epilogue -- 13.9 ms for a residual
    // 1x1 conv whose twin without activation took 1.6 ms.)
    auto clamps = [](int act) { return act == VP_ACT_RELU || act == VP_ACT_HARDTANH20; };
    const bool simple = !p.ubias && !p.gate && !p.sum && (p.act == VP_ACT_NONE || clamps(p.act)) &&
                        (p.act2 == VP_ACT_NONE || clamps(p.act2));
    const bool relu = clamps(p.act), relu2 = clamps(p.act2);
    const float hi1 = p.act == VP_ACT_HARDTANH20 ? 20.f : INFINITY, hi2 = p.act2 == VP_ACT_HARDTANH20 ? 20.f : INFINITY;
    const size_t row4 = (size_t)4 * p.out_ld;      // floats between the rows rsub + 4i and rsub + 4(i+1)
    const size_t res4 = (size_t)4 * p.res_ld;
    for (int g = cluster_id; g < total_groups; g += n_clusters) {
      const int mbase = ((g / a.n_tiles) * (int)C + (int)crank) * BM + quad * 32;
      const int n0 = (g % a.n_tiles) * BN;
      if (p.res) {
        // Residual tiles are pure streaming (c5: 128 KB per 128 x 256 tile, read once): with only eight 16-byte loads in
        // flight per epilogue thread the reads were latency bound at ~5 GB/s per SM (15 ms for the M = 5 M, K = 72 layers
        // of the 55 M ERes2Net).  prefetch.global.L2 is fire-and-forget: pull the NEXT tile's residual rows into the L2
        // now (and this tile's, on the first iteration), so that the loads below are L2 hits.
        auto prefetch_tile = [&](int gg) {
          const int mb = ((gg / a.n_tiles) * (int)C + (int)crank) * BM + quad * 32 + rsub;
          const float* base = p.res + p.res_coff + (gg % a.n_tiles) * BN + (lane & 7) * 32;
          if ((lane & 7) * 32 < BN)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (mb + 4 * i < p.M)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(mb + 4 * i) * p.res_ld));
        };
        if (g == cluster_id) prefetch_tile(g);
        if (g + n_clusters < total_groups) prefetch_tile(g + n_clusters);
      }
      if (simple && mbase + 32 <= p.M && n0 + BN <= p.N && (BN & 31) == 0) {
        float* o0 = p.dst + (size_t)(mbase + rsub) * p.out_ld + p.out_coff + n0 + cg;
        const float* r0 = p.res ? p.res + (size_t)(mbase + rsub) * p.res_ld + p.res_coff + n0 + cg : nullptr;
        for (int ch = 0; ch + 1 < a.n_chunks; ++ch, ++ccount) {      // chunk folding exactly as in the general path
          const int accf = ccount & 1;
          mbar_wait(tfull0 + 8 * accf, (ccount >> 1) & 1);
          tc_fence_after();
          for (int c0 = 32 * chalf; c0 < BN; c0 += CSTEP) {
            float v[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(accf * BN + c0), v);
            if (ch > 0) {
              float r[32];
              tmem_ld32(tmem_base + lane_base + run_col + (uint32_t)c0, r);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += r[j];
            }
            tmem_st32(tmem_base + lane_base + run_col + (uint32_t)c0, v);
          }
          tc_fence_before();
          mbar_arrive(tempty0 + 8 * accf);
        }
        const int acc = ccount & 1;
        mbar_wait(tfull0 + 8 * acc, (ccount >> 1) & 1);
        ++ccount;
        tc_fence_after();
        for (int c0 = 32 * chalf; c0 < BN; c0 += CSTEP) {
          const int n = n0 + c0 + cg;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = b4;
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          if (p.post_s) {
            s4 = __ldg(reinterpret_cast<const float4*>(p.post_s + n));
            h4 = __ldg(reinterpret_cast<const float4*>(p.post_h + n));
          }
          {
            float v[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(acc * BN + c0), v);
            if (a.n_chunks > 1) {
              float r[32];
              tmem_ld32(tmem_base + lane_base + run_col + (uint32_t)c0, r);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += r[j];
            }
            if (c0 + CSTEP >= BN) {
              tc_fence_before();
              mbar_arrive(tempty0 + 8 * acc);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pad_u32 + (uint32_t)(lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) * 4u), "f"(v[j]),
                           "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          }
          float4 rr[8];
          if (r0) {
            // residual: eight 16-byte loads issued as ONE batch (volatile asm: ptxas otherwise sinks every load next to its
            // use to save registers, which serialises eight DRAM round trips per chunk -- the N = 256, K = 72 layers of the
            // 55 M ERes2Net took 15 ms that way), before the pad is read back so that the latency overlaps the transpose
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rr[i].x), "=f"(rr[i].y), "=f"(rr[i].z), "=f"(rr[i].w)
                           : "l"(r0 + c0 + i * res4));
          }
          __syncwarp();
          float* o = o0 + c0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x = *reinterpret_cast<const float4*>(pad + (rsub + 4 * i) * 32 + ((((lane & 7) ^ ((rsub + 4 * i) & 7))) << 2));
            x.x = fmaf(x.x, descale, b4.x); x.y = fmaf(x.y, descale, b4.y);      // descale == 1 on the tf32 path
            x.z = fmaf(x.z, descale, b4.z); x.w = fmaf(x.w, descale, b4.w);
            if (relu) {
              x.x = fminf(fmaxf(x.x, 0.f), hi1); x.y = fminf(fmaxf(x.y, 0.f), hi1);
              x.z = fminf(fmaxf(x.z, 0.f), hi1); x.w = fminf(fmaxf(x.w, 0.f), hi1);
            }
            if (p.post_s) {
              x.x = fmaf(x.x, s4.x, h4.x); x.y = fmaf(x.y, s4.y, h4.y); x.z = fmaf(x.z, s4.z, h4.z); x.w = fmaf(x.w, s4.w, h4.w);
            }
            if (r0) { x.x += rr[i].x; x.y += rr[i].y; x.z += rr[i].z; x.w += rr[i].w; }
            if (relu2) {
              x.x = fminf(fmaxf(x.x, 0.f), hi2); x.y = fminf(fmaxf(x.y, 0.f), hi2);
              x.z = fminf(fmaxf(x.z, 0.f), hi2); x.w = fminf(fmaxf(x.w, 0.f), hi2);
            }
            *reinterpret_cast<float4*>(o + i * row4) = x;
            tmax = amax4(tmax, x);
          }
          __syncwarp();
        }
        if (32 * chalf >= BN) {                     // this warp has no chunk on tiles this narrow: still release the buffer
          tc_fence_before();
          mbar_arrive(tempty0 + 8 * acc);
        }
        continue;
      }
      // per-tile row validity (rows rsub + 4i of this warp's 32)
      uint32_t rowok = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) rowok |= (mbase + rsub + 4 * i < p.M) ? (1u << i) : 0u;
      // fold all but the last accumulation chunk into the running sum (TMEM columns [2BN, 3BN))
      for (int ch = 0; ch + 1 < a.n_chunks; ++ch, ++ccount) {
        const int accf = ccount & 1;
        mbar_wait(tfull0 + 8 * accf, (ccount >> 1) & 1);
        tc_fence_after();
        for (int c0 = 32 * chalf; c0 < BN; c0 += CSTEP) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + (uint32_t)(accf * BN + c0), v);
          if (ch > 0) {
            float r[32];
            tmem_ld32(tmem_base + lane_base + run_col + (uint32_t)c0, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += r[j];
          }
          tmem_st32(tmem_base + lane_base + run_col + (uint32_t)c0, v);
        }
        tc_fence_before();
        mbar_arrive(tempty0 + 8 * accf);
      }
      const int acc = ccount & 1;
      mbar_wait(tfull0 + 8 * acc, (ccount >> 1) & 1);
      ++ccount;
      tc_fence_after();
      for (int c0 = 32 * chalf; c0 < BN; c0 += CSTEP) {
        const int n = n0 + c0 + cg;
        const bool nok = n < p.N;
        // per-column parameters first: their latency hides behind the TMEM load + transpose
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = b4;
        if (nok) {
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          if (p.post_s) {
            s4 = __ldg(reinterpret_cast<const float4*>(p.post_s + n));
            h4 = __ldg(reinterpret_cast<const float4*>(p.post_h + n));
          }
        }
        {
          float v[32];
          tmem_ld32(tmem_base + lane_base + (uint32_t)(acc * BN + c0), v);
          if (a.n_chunks > 1) {
            float r[32];
            tmem_ld32(tmem_base + lane_base + run_col + (uint32_t)c0, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += r[j];
          }
          if (c0 + CSTEP >= BN) {             // last read of this accumulator buffer: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * acc);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pad_u32 + (uint32_t)(lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) * 4u), "f"(v[j]),
                         "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                         : "memory");
        }
        __syncwarp();
        if (nok) {
          // General epilogue, one row at a time (deliberately NOT unrolled: this path serves the tile tails and the
          // per-utterance / gated / accumulate-into layers; the hot shapes take the fast path above, and keeping this
          // one compact keeps the kernel's instruction footprint -- and its instruction-cache misses -- down).
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            if (!(rowok & (1u << i))) continue;
            const int m = mbase + rsub + 4 * i;
            float4 x = *reinterpret_cast<const float4*>(pad + (rsub + 4 * i) * 32 + ((((lane & 7) ^ ((rsub + 4 * i) & 7))) << 2));
            // acc * descale + bias in one rounding (descale == 1 on the tf32 path: exactly acc + bias) -- the same
            // expression as the fast path, so a row's result does not depend on which path its tile took
            x.x = fmaf(x.x, descale, b4.x); x.y = fmaf(x.y, descale, b4.y);
            x.z = fmaf(x.z, descale, b4.z); x.w = fmaf(x.w, descale, b4.w);
            const int ur = need_urow ? urow_of(p, m) : 0;
            if (p.ubias) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(p.ubias + (size_t)ur * p.N + n));
              x.x += u.x; x.y += u.y; x.z += u.z; x.w += u.w;
            }
            if (p.act == VP_ACT_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            else if (p.act != VP_ACT_NONE) x = act4_slow(x, p.act);
            if (p.post_s) {
              x.x = fmaf(x.x, s4.x, h4.x); x.y = fmaf(x.y, s4.y, h4.y); x.z = fmaf(x.z, s4.z, h4.z); x.w = fmaf(x.w, s4.w, h4.w);
            }
            if (p.gate) {
              const float4 gt = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)ur * p.N + n));
              x.x *= gt.x; x.y *= gt.y; x.z *= gt.z; x.w *= gt.w;
            }
            if (p.res) {
              const float4 r = __ldg(reinterpret_cast<const float4*>(p.res + (size_t)m * p.res_ld + p.res_coff + n));
              x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
            }
            if (p.act2 == VP_ACT_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            else if (p.act2 != VP_ACT_NONE) x = act4_slow(x, p.act2);
            *reinterpret_cast<float4*>(p.dst + (size_t)m * p.out_ld + p.out_coff + n) = x;
            tmax = amax4(tmax, x);
            if (p.sum) {                         // accumulate-into view (Res2 chains): sum[m, n] += y[m, n]
              float4* q = reinterpret_cast<float4*>(p.sum + (size_t)m * p.sum_ld + p.sum_coff + n);
              float4 sv = *q;
              sv.x += x.x; sv.y += x.y; sv.z += x.z; sv.w += x.w;
              *q = sv;
            }
          }
        }
        __syncwarp();
      }
      if (32 * chalf >= BN) {                       // no chunk for this warp on tiles this narrow: still release the buffer
        tc_fence_before();
        mbar_arrive(tempty0 + 8 * acc);
      }
    }
    if (p.amax_out) amax_commit(p.amax_out, tmax);     // one atomicMax per epilogue warp per launch
  }

  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();           // no CTA exits while a peer may still multicast into it
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

}  // namespace tc

// Host-visible tiling rule (mirrored by the Python packer mvector/engine.py::tc_tile_n).  Chunked layers (vp_op.tc_kc > 0,
// see the MMA issuer) need a third TMEM accumulator for the running sum -> N tile <= 128.
static int tc_tile_n(int N, bool chunked) {
  if (N >= 256 && !chunked) return 256;
  if (N >= 128) return 128;
  return (N + 15) & ~15;
}

bool conv_tc_supported(const ConvParams& p) {
  if (p.w_tc == nullptr) return false;
  if (p.M < 1024) return false;                       // tiny-M ops (SE / ASP bias / final FC) stay on the exact FFMA engine
  if (p.N < 16 || (p.N & 3) || (p.K & 3)) return false;
  if (p.tc_kc < 0 || (p.tc_kc & 63)) return false;
  if (p.tc_bn != tc_tile_n(p.N, p.tc_kc > 0)) return false;
  // the BN-ReLU prologue gather (MODE 2) reads one source only: prologue + second source stays on the FFMA engine, whose
  // gather composes both (common.cuh::gather_a4)
  if (p.pre_s != nullptr && p.src2_mode != VP_SRC2_NONE) return false;
  return true;
}

// cudaFuncSetAttribute is per device: remember which devices have been configured (one Engine per device per process is
// the norm, but nothing stops a host from creating handles on several GPUs)
static cudaError_t configure_once(int dev) {
  static bool done[64] = {};
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (done[dev]) return cudaSuccess;
  const int bytes = tc::SMEM_BUDGET + tc::EPI_PAD_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc::conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::conv_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[dev] = true;
  return e;
}

// f16 = true: two-term FP16 split (conv_tc_kernel<0, true>); w_img is then the fp16 weight image and descale the inverse
// of the power-of-two scale folded into it.
static cudaError_t launch_conv_tc_impl(const ConvParams& p, const float* w_img, bool f16, float descale, cudaStream_t stream) {
  using namespace tc;
  TcArgs a;
  a.w_tc = w_img;
  a.descale = descale;
  a.BN = p.tc_bn;
  const int bke = f16 ? 64 : BK;
  const int stage_bytes = 2 * A_TILE + 2 * a.BN * 128;
  // (Round 2 tried a smaller pipeline for narrow multi-tap layers so that the L1 would serve the repeated taps: no effect,
  // removed -- profiles/r2_ncu_conv_tc.md section 4.)
  a.stages = SMEM_BUDGET / stage_bytes;
  if (a.stages > 8) a.stages = 8;
  if (a.stages < 2) return cudaErrorInvalidConfiguration;
  a.m_tiles = (p.M + BM - 1) / BM;
  a.n_tiles = (p.N + a.BN - 1) / a.BN;
  a.k_blocks = (p.K + bke - 1) / bke;
  a.kc = (p.tc_kc > 0 && p.tc_kc < p.K) ? p.tc_kc / bke : a.k_blocks;
  a.n_chunks = (a.k_blocks + a.kc - 1) / a.kc;
  // accumulator regions [acc*BN, +BN) (+ running sum at 2*BN when chunked); tcgen05.ld reads 32 columns at a time, so
  // the last 32-column read of the last region must stay inside the allocation
  const int regions = a.n_chunks > 1 ? 3 : 2;
  const int need = (regions - 1) * a.BN + ((a.BN + 31) / 32) * 32;
  int cols = 32;
  while (cols < need || cols < regions * a.BN) cols <<= 1;
  if (cols > 512) return cudaErrorInvalidConfiguration;
  a.tmem_cols = cols;
  static int debug_flags = -1;
  if (debug_flags < 0) { const char* e = getenv("VPB_TC_DEBUG"); debug_flags = e ? atoi(e) : 0; }
  a.debug = debug_flags;
  const size_t smem = (size_t)a.stages * stage_bytes + EPI_PAD_BYTES + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaError_t e = configure_once(dev);
  if (e != cudaSuccess) return e;
  static int sm_count[64] = {};
  if (sm_count[dev] == 0) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
  if (sm_count[dev] > 0) sms = sm_count[dev];
  static int cluster_pref = -1;
  if (cluster_pref < 0) {
    const char* ev = getenv("VPB_TC_CLUSTER");
    cluster_pref = ev ? atoi(ev) : 2;
    if (cluster_pref != 1 && cluster_pref != 2 && cluster_pref != 4) cluster_pref = 2;
  }
  int C = cluster_pref;
  while (C > 1 && (a.m_tiles < 2 * C || ((2 * a.BN * 128 / C) & 15))) C >>= 1;
  const int groups = ((a.m_tiles + C - 1) / C) * a.n_tiles;
  int clusters = sms / C;
  if (clusters > groups) clusters = groups;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * C);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (f16) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, true>, p, a);
  else if (p.pre_s != nullptr) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<2>, p, a);
  else if (p.src2_mode == VP_SRC2_ADD) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<1>, p, a);
  else e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<0>, p, a);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_conv_tc(const ConvParams& p, cudaStream_t stream) {
  return launch_conv_tc_impl(p, p.w_tc, false, 1.f, stream);
}

// FP16-split engine: plain / concat source only, 8-element channel granularity, the same tiling rule as the tf32 image,
// and a source tensor whose amax slot is tracked (dynamic power-of-two activation scale -> range-safe).
bool conv_tc16_supported(const ConvParams& p) {
  if (!conv_tc_supported(p)) return false;
  if (p.amax_in == nullptr) return false;
  if (p.pre_s != nullptr || p.src2_mode == VP_SRC2_ADD) return false;
  if ((p.Cin & 7) || (p.CinTot & 7) || (p.K & 7)) return false;
  return p.N >= 128;                                  // narrow layers are not tensor bound: nothing to gain
}

cudaError_t launch_conv_tc16(const ConvParams& p, const float* w_tc16, float descale, cudaStream_t stream) {
  return launch_conv_tc_impl(p, w_tc16, true, descale, stream);
}

}  // namespace vpb
