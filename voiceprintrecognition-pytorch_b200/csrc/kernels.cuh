// Kernel parameter blocks + host launchers shared between translation units.
#pragma once
#include "common.cuh"

namespace vpb {

struct FrontendParams {
  const float* wave; float* feats; float* partial;
  const float* window; const double2* twiddle;   // twiddle[k] = exp(-2 pi i k / N) in fp64
  const int* mel_start; const int* mel_count; const int* mel_off; const float* mel_w;
  int B, L, T, kind, N, WL, hop, F, remove_dc, power, use_log, fpb, nblk;
  float preemph, log_floor, db_mult;
  float* cta_max;      // non-null: MFCC mel stage (per-CTA maxima instead of CMN partial sums)
  int n_pass, radix[12], G;   // FFT pass plan + threads per FFT group (frontend_plan)
};
void frontend_plan(FrontendParams& p);

// MFCC tail: mel [B,T,M] dB values -> clamp(max - top_db) -> DCT [M,K] -> feats [B,T,K] + CMN partial sums.
struct MfccParams {
  const float* mel; const float* cta_max; const float* dct; float* feats; float* partial;
  int B, T, M, K, fpb, nblk, n_max;
  float top_db;        // < 0: no clamp
};

struct StatsParams {
  const float* src; float* dst;
  int B, R, C, in_ld, in_coff, out_ld, out_coff, mode, seg_len, n_seg;
  float eps;
};

// x: [B,T,C] view (x_ld/x_coff), logits likewise; dst[b, c] = mean, dst[b, C + c] = std.
struct AspParams {
  const float* x; const float* logit; float* dst;
  int B, T, C, x_ld, x_coff, l_ld, l_coff, out_ld, out_coff, mean_only;
  float eps;
};

struct EwParams {
  const float* x; const float* y; const float* att; const float* gate; const float* res; float* dst;
  long long rows; int C, rows_per_utt;
  int x_ld, x_coff, y_ld, y_coff, att_ld, att_coff, res_ld, res_coff, out_ld, out_coff, mode, act2;
  int C_out;   // PAD_COPY: destination width (>= C, zero filled)
  unsigned* amax_out;
};

struct PoolParams {
  const float* src; float* dst;
  int B, Tin, Fin, Tout, Fout, C, in_ld, in_coff, out_ld, out_coff, KT, KF, sT, sF, padT, padF, mode;
  unsigned* amax_out;
};
cudaError_t launch_pool2d(const PoolParams& p, cudaStream_t stream);
cudaError_t launch_cosine_scores(const float* a, const float* b, float* out, int n, int m, int D, cudaStream_t stream);

cudaError_t launch_frontend(const FrontendParams& p, const int* keep, cudaStream_t stream);
cudaError_t launch_frontend_mfcc(const FrontendParams& p, const MfccParams& m, const int* keep, cudaStream_t stream);
cudaError_t launch_frontend_mfcc_mel(const FrontendParams& p, float* max_out, cudaStream_t stream);
cudaError_t launch_frontend_mfcc_finish(const FrontendParams& p, const MfccParams& m, const int* keep, cudaStream_t stream);
size_t frontend_smem_bytes(int N, int WL, int hop, int fpb);
cudaError_t launch_conv_ffma(const ConvParams& p, cudaStream_t stream);
cudaError_t launch_conv_c1(const ConvParams& p, cudaStream_t stream);
cudaError_t launch_colstats(const StatsParams& p, cudaStream_t stream);
cudaError_t launch_asp_pool(const AspParams& p, cudaStream_t stream);
cudaError_t launch_ew(const EwParams& p, cudaStream_t stream);
// tcgen05 engine (conv_tc.cu)
bool conv_tc_supported(const ConvParams& p);
cudaError_t launch_conv_tc(const ConvParams& p, cudaStream_t stream);
// two-term FP16 split of the same engine (VP_ENGINE_TC16)
bool conv_tc16_supported(const ConvParams& p);
cudaError_t launch_conv_tc16(const ConvParams& p, const float* w_tc16, float descale, cudaStream_t stream);

// opt-in experimental variants (pool_v2.cu, VPB_POOL_V2=1)
bool asp_pool_v2_supported(const AspParams& p);
cudaError_t launch_asp_pool_v2(const AspParams& p, cudaStream_t stream);
bool colstats_v2_supported(const StatsParams& p);
cudaError_t launch_colstats_v2(const StatsParams& p, cudaStream_t stream);

}  // namespace vpb
