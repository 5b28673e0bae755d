// Fused front-end: framing -> (DC removal, pre-emphasis) -> window -> real FFT -> power -> sparse mel -> log,
// one kernel, waveform read once from HBM through a shared-memory stage; then CMN + length mask.
//
// Replaces the reference's per-utterance Python loop over torchaudio.compliance.kaldi.fbank
// (mvector/data_utils/featurizer.py:119-132 -> kaldi.py:514-645: ~250 ATen ops per utterance, window and mel bank
// rebuilt on every call) and torchaudio.transforms.MelSpectrogram (featurizer.py:41-42,76), followed by
// AudioFeaturizer.forward's transpose / mean-subtract / mask (featurizer.py:77-90).
//
// The same kernel serves torchaudio.transforms.Spectrogram (identity "mel" bank, featurizer.py:43-44) and the mel stage
// of torchaudio.transforms.MFCC (featurizer.py:45-46), whose AmplitudeToDB / top_db clamp / DCT-II run in mfcc_post_kernel.
//
// FFT: two real frames are packed into one complex length-N Stockham autosort FFT held in shared memory; N/4 threads
// (<= 256) cooperate on one FFT.  N = 2^a 3^b 5^c (4 | N): radix-4 passes first, one radix-2 pass if needed, then
// generic radix-5 / radix-3 passes (torchaudio's default n_fft = 400 = 4*4*5*5).  Twiddles come from a
// host-computed (fp64 -> fp32) table.  Bound: the algorithmic traffic is 4*L + 4*T*F bytes per utterance (HBM), the
// kernel itself is shared-memory / issue bound.
#include "kernels.cuh"

namespace vpb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) frontend_kernel(const __grid_constant__ FrontendParams p) {
  extern __shared__ __align__(16) float smem[];
  const int N = p.N, WL = p.WL, F = p.F;
  const int G = (N / 4 < 256) ? N / 4 : 256;        // threads per FFT
  const int NG = 256 / G;                           // concurrent FFTs per CTA
  const int span = (p.fpb - 1) * p.hop + WL;

  float* stage = smem;                                              // span floats (rounded up to x4)
  float* win = stage + ((span + 3) & ~3);                           // WL
  float2* tw = reinterpret_cast<float2*>(win + ((WL + 3) & ~3));    // N
  float2* bufA = tw + N;                                            // NG * N
  float2* bufB = bufA + NG * N;                                     // NG * N
  float* means = reinterpret_cast<float*>(bufB + NG * N);           // NG * 2

  float* red = means + NG * 2;                                      // 8 floats: per-warp maxima (MFCC mel stage)

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * p.fpb;
  const int g = tid / G;
  const int t = tid - g * G;
  const int NGe = NG < p.fpb / 2 ? NG : p.fpb / 2;  // groups that have a frame pair to work on
  const bool active = g < NGe;                      // 256 % G != 0 (e.g. N = 400 -> G = 100): the tail threads idle
  float vmax = -INFINITY;
  const float* wv = p.wave + (size_t)b * p.L;

  // ---- stage the waveform span, window and twiddles ----
  const int q0 = f0 * p.hop;
  for (int i = tid; i < span; i += 256) {
    int s = q0 + i;
    float v = 0.f;
    if (p.kind == 1) {                      // torch.stft(center=True, pad_mode='reflect'): functional.py:123-135
      s -= N / 2;
      if (s < 0) s = -s;
      if (s >= p.L) s = 2 * (p.L - 1) - s;
      if (s >= 0 && s < p.L) v = __ldg(wv + s);
    } else if (s < p.L) {
      v = __ldg(wv + s);
    }
    stage[i] = v;
  }
  for (int i = tid; i < WL; i += 256) win[i] = __ldg(p.window + i);
  for (int i = tid; i < N; i += 256) tw[i] = __ldg(p.twiddle + i);
  __syncthreads();

  const int iters = (p.fpb / 2 + NGe - 1) / NGe;
  for (int it = 0; it < iters; ++it) {
    const int pair = it * NGe + g;                  // frame pair of this CTA's tile handled by group g
    const int fa = f0 + pair * 2;                   // frames packed as real (fa) and imaginary (fa + 1) parts
    const int oa = (fa - f0) * p.hop;
    const bool pv = active && pair < p.fpb / 2;
    const bool va = pv && fa < p.T, vb = pv && fa + 1 < p.T;

    // ---- per-frame mean (kaldi.py:183-186), one warp per frame ----
    if (p.kind == 0 && p.remove_dc && active) {
      const int w = t >> 5, lane = t & 31;
      if (w < 2) {
        const int o = oa + w * p.hop;
        float s = 0.f;
        if (w == 0 ? va : vb)
          for (int j = lane; j < WL; j += 32) s += stage[o + j];
        s = warp_sum(s);
        if (lane == 0) means[g * 2 + w] = s / (float)WL;
      }
    }
    __syncthreads();
    float2* src = bufA + (active ? g : 0) * N;
    float2* dst = bufB + (active ? g : 0) * N;
    if (active) {
      float ma = 0.f, mb = 0.f;
      if (p.kind == 0 && p.remove_dc) { ma = means[g * 2]; mb = means[g * 2 + 1]; }
      for (int j = t; j < N; j += G) {
        float ya = 0.f, yb = 0.f;
        if (j < WL) {
          const int jp = j > 0 ? j - 1 : 0;
          const float wj = win[j];
          if (va) {
            float x = stage[oa + j];
            if (p.kind == 0) {
              x = __fsub_rn(x, ma);
              if (p.preemph != 0.f) x = __fsub_rn(x, __fmul_rn(p.preemph, __fsub_rn(stage[oa + jp], ma)));
            }
            ya = __fmul_rn(x, wj);
          }
          if (vb) {
            float x = stage[oa + p.hop + j];
            if (p.kind == 0) {
              x = __fsub_rn(x, mb);
              if (p.preemph != 0.f) x = __fsub_rn(x, __fmul_rn(p.preemph, __fsub_rn(stage[oa + p.hop + jp], mb)));
            }
            yb = __fmul_rn(x, wj);
          }
        }
        src[j] = make_float2(ya, yb);
      }
    }
    __syncthreads();

    // ---- Stockham autosort FFT, radix 4 (+ one radix-2 pass) ----
    for (int Ns = 1; Ns < N;) {
      const int rem = N / Ns;
      const int R = (rem % 4 == 0) ? 4 : (rem % 2 == 0) ? 2 : (rem % 5 == 0) ? 5 : 3;
      const int step = N / (Ns * R);
      if (!active) {
        // nothing: idle tail threads only take part in the barriers
      } else if (R == 4) {
        const int q = N >> 2;
        for (int j = t; j < q; j += G) {
          const int kk = j & (Ns - 1);
          float2 v0 = src[j], v1 = src[j + q], v2 = src[j + 2 * q], v3 = src[j + 3 * q];
          if (Ns > 1) {
            v1 = cmul(v1, tw[kk * step]);
            v2 = cmul(v2, tw[2 * kk * step]);
            v3 = cmul(v3, tw[3 * kk * step]);
          }
          const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
          const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
          const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
          const float2 a3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));      // (v1 - v3) * (-i)
          const int base = (j - kk) * 4 + kk;
          dst[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
          dst[base + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
          dst[base + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
          dst[base + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
        }
      } else if (R == 2) {
        const int q = N >> 1;
        for (int j = t; j < q; j += G) {
          const int kk = j & (Ns - 1);
          float2 v0 = src[j], v1 = cmul(src[j + q], tw[kk * step]);
          const int base = (j - kk) * 2 + kk;
          dst[base] = make_float2(v0.x + v1.x, v0.y + v1.y);
          dst[base + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
      } else {
        // generic odd radix (5 or 3): out[m] = sum_r v[r] * W_R^(r m), W_R^k = tw[k * N / R]; Ns need not be a power of 2
        const int q = N / R;
        for (int j = t; j < q; j += G) {
          const int kk = j % Ns;
          float2 v[5];
#pragma unroll
          for (int r = 0; r < 5; ++r)
            if (r < R) {
              v[r] = src[j + r * q];
              if (r > 0 && Ns > 1) v[r] = cmul(v[r], tw[r * kk * step]);
            }
          const int base = (j - kk) * R + kk;
#pragma unroll
          for (int m = 0; m < 5; ++m)
            if (m < R) {
              float2 acc = v[0];
#pragma unroll
              for (int r = 1; r < 5; ++r)
                if (r < R) {
                  const float2 w = tw[((r * m) % R) * q];
                  acc.x += v[r].x * w.x - v[r].y * w.y;
                  acc.y += v[r].x * w.y + v[r].y * w.x;
                }
              dst[base + m * Ns] = acc;
            }
        }
      }
      __syncthreads();
      float2* tmp = src; src = dst; dst = tmp;
      Ns *= R;
    }

    // ---- split the packed spectrum, power (kaldi.py:616-618) into P[2][N/2+1] (reuses the idle FFT buffer) ----
    const int NB = N / 2 + 1;
    float* P = reinterpret_cast<float*>(dst);
    for (int k = active ? t : NB; k < NB; k += G) {
      const float2 z = src[k];
      const float2 zn = src[k == 0 ? 0 : N - k];
      const float ar = 0.5f * (z.x + zn.x), ai = 0.5f * (z.y - zn.y);
      const float br = 0.5f * (z.y + zn.y), bi = -0.5f * (z.x - zn.x);
      float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
      if (p.power == 1) { pa = sqrtf(pa); pb = sqrtf(pb); }
      P[k] = pa;
      P[NB + k] = pb;
    }
    __syncthreads();

    // ---- sparse triangular mel projection + log floor (kaldi.py:630-633) ----
    for (int idx = active ? t : 2 * F; idx < 2 * F; idx += G) {
      const int fr = idx / F;
      const int m = idx - fr * F;
      const int f = fa + fr;
      if (fr == 0 ? va : vb) {
        const int st = __ldg(p.mel_start + m), cnt = __ldg(p.mel_count + m), off = __ldg(p.mel_off + m);
        const float* pp = P + fr * NB + st;
        float s = 0.f;
        for (int i = 0; i < cnt; ++i) s = fmaf(pp[i], __ldg(p.mel_w + off + i), s);
        if (p.use_log == 1) s = logf(fmaxf(s, p.log_floor));                       // kaldi.py:633
        else if (p.use_log == 2) s = p.db_mult * log10f(fmaxf(s, p.log_floor));   // amplitude_to_DB, functional.py:389-391
        else if (p.use_log == 3) s = logf(s + p.log_floor);                        // MFCC(log_mels=True), transforms MFCC.forward
        vmax = fmaxf(vmax, s);
        p.feats[((size_t)b * p.T + f) * F + m] = s;
      }
    }
    __syncthreads();
  }

  if (p.cta_max) {
    // ---- MFCC mel stage: per-CTA maximum for the call-wide top_db clamp (functional.py:393-399); CMN comes after the DCT
    vmax = warp_max(vmax);
    if ((tid & 31) == 0) red[tid >> 5] = vmax;
    __syncthreads();
    if (tid == 0) {
      float m = red[0];
      for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
      p.cta_max[(size_t)b * p.nblk + blockIdx.x] = m;
    }
    return;
  }
  // ---- per-CTA column sums for the CMN mean (featurizer.py:79), fixed summation order ----
  for (int m = tid; m < F; m += 256) {
    float s = 0.f;
    for (int f = f0; f < f0 + p.fpb && f < p.T; ++f) s += p.feats[((size_t)b * p.T + f) * F + m];
    p.partial[((size_t)b * p.nblk + blockIdx.x) * F + m] = s;
  }
}

// MFCC tail (torchaudio MFCC.forward): clamp the dB mel values to (max over the WHOLE call) - top_db -- torchaudio folds
// the batch axis into the clamp's channel axis, so the maximum is shared by every utterance of the call -- then
// mfcc[t, k] = sum_m mel_db[t, m] * dct[m, k] (create_dct, functional.py:640-667).  One CTA per (frame tile, utterance);
// also emits the per-CTA column sums that cmn_mask_kernel turns into the CMN mean.
__global__ void __launch_bounds__(256) mfcc_post_kernel(const __grid_constant__ MfccParams p) {
  extern __shared__ __align__(16) float sm[];
  float* dct = sm;                              // [M][K]
  float* tile = dct + p.M * p.K;                // [fpb][M]
  float* outt = tile + p.fpb * p.M;             // [fpb][K]
  float* red = outt + p.fpb * p.K;              // 8
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * p.fpb;
  float thr = -INFINITY;
  if (p.top_db >= 0.f) {
    float m = -INFINITY;
    for (int i = tid; i < p.n_max; i += 256) m = fmaxf(m, __ldg(p.cta_max + i));
    m = warp_max(m);
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    thr = m - p.top_db;
  }
  for (int i = tid; i < p.M * p.K; i += 256) dct[i] = __ldg(p.dct + i);
  for (int i = tid; i < p.fpb * p.M; i += 256) {
    const int fr = i / p.M;
    const int f = f0 + fr;
    tile[i] = f < p.T ? fmaxf(__ldg(p.mel + ((size_t)b * p.T + f) * p.M + (i - fr * p.M)), thr) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < p.fpb * p.K; i += 256) {
    const int fr = i / p.K;
    const int k = i - fr * p.K;
    const float* row = tile + fr * p.M;
    float s = 0.f;
    for (int m = 0; m < p.M; ++m) s = fmaf(row[m], dct[m * p.K + k], s);
    outt[i] = s;
    if (f0 + fr < p.T) p.feats[((size_t)b * p.T + f0 + fr) * p.K + k] = s;
  }
  __syncthreads();
  for (int k = tid; k < p.K; k += 256) {
    float s = 0.f;
    for (int fr = 0; fr < p.fpb && f0 + fr < p.T; ++fr) s += outt[fr * p.K + k];
    p.partial[((size_t)b * p.nblk + blockIdx.x) * p.K + k] = s;
  }
}

// feats[b, t, :] -= mean_t(feats[b]) over ALL T frames, then frames t >= keep[b] are zeroed (featurizer.py:79-90).
__global__ void __launch_bounds__(128) cmn_mask_kernel(float* feats, const float* partial, const int* keep, int T, int F,
                                                       int nblk, int rows_per_cta) {
  const int b = blockIdx.y;
  const int kp = keep ? keep[b] : T;
  const int t0 = blockIdx.x * rows_per_cta;
  for (int m = threadIdx.x; m < F; m += blockDim.x) {        // F > 128 only for Spectrogram (n_fft/2 + 1 bins)
    float s = 0.f;
    for (int i = 0; i < nblk; ++i) s += partial[((size_t)b * nblk + i) * F + m];
    const float mean = s / (float)T;
    for (int t = t0; t < t0 + rows_per_cta && t < T; ++t) {
      float* q = feats + ((size_t)b * T + t) * F + m;
      *q = (t < kp) ? (*q - mean) : 0.f;
    }
  }
}

// out[0] = max(v[0..n)): the call-wide (or, sharded, the rank-wide) maximum of the per-CTA maxima of the MFCC mel stage
__global__ void __launch_bounds__(256) max_reduce_kernel(const float* v, int n, float* out) {
  __shared__ float red[8];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, v[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    out[0] = m;
  }
}

size_t frontend_smem_bytes(int N, int WL, int hop, int fpb) {
  const int G = (N / 4 < 256) ? N / 4 : 256;
  const int NG = 256 / G;
  const int span = (fpb - 1) * hop + WL;
  size_t fl = ((span + 3) & ~3) + ((WL + 3) & ~3) + 2 * (size_t)N + 2 * 2 * (size_t)NG * N + 2 * NG + 8 + 4;
  return fl * sizeof(float);
}

// MFCC: mel stage (dB values into p.feats = the temporary mel buffer, maxima into p.cta_max), then clamp + DCT into
// m.feats and the CMN partial sums, then CMN + mask over the K cepstral coefficients.
// Dynamic shared memory above 48 KB must be opted into once per kernel; remember the largest request so far.
static cudaError_t ensure_frontend_smem(size_t smem) {
  static PerDeviceSmem once;
  if (once.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    once.set(smem);
  }
  return cudaSuccess;
}

// MFCC stage 1 only: mel dB values + per-CTA maxima, and their maximum into max_out[0] (device) -- the scalar a sharded
// call all-reduces (MAX) across ranks before stage 2.
cudaError_t launch_frontend_mfcc_mel(const FrontendParams& p, float* max_out, cudaStream_t stream) {
  size_t smem = frontend_smem_bytes(p.N, p.WL, p.hop, p.fpb);
  cudaError_t e = ensure_frontend_smem(smem);
  if (e != cudaSuccess) return e;
  dim3 grid(p.nblk, p.B);
  frontend_kernel<<<grid, 256, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  max_reduce_kernel<<<1, 256, 0, stream>>>(p.cta_max, p.B * p.nblk, max_out);
  return cudaGetLastError();
}

// MFCC stage 2 only: clamp against m.cta_max[0..n_max) (one externally reduced scalar when n_max == 1), DCT, CMN + mask.
cudaError_t launch_frontend_mfcc_finish(const FrontendParams& p, const MfccParams& m, const int* keep, cudaStream_t stream) {
  dim3 grid(p.nblk, p.B);
  size_t smem2 = ((size_t)m.M * m.K + (size_t)m.fpb * (m.M + m.K) + 8) * sizeof(float);
  static PerDeviceSmem once;
  cudaError_t e;
  if (once.need(smem2)) {
    e = cudaFuncSetAttribute(mfcc_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return e;
    once.set(smem2);
  }
  mfcc_post_kernel<<<grid, 256, smem2, stream>>>(m);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int rows = 64;
  dim3 g2((p.T + rows - 1) / rows, p.B);
  cmn_mask_kernel<<<g2, 128, 0, stream>>>(m.feats, m.partial, keep, p.T, m.K, p.nblk, rows);
  return cudaGetLastError();
}

cudaError_t launch_frontend_mfcc(const FrontendParams& p, const MfccParams& m, const int* keep, cudaStream_t stream) {
  size_t smem = frontend_smem_bytes(p.N, p.WL, p.hop, p.fpb);
  cudaError_t e = ensure_frontend_smem(smem);
  if (e != cudaSuccess) return e;
  dim3 grid(p.nblk, p.B);
  frontend_kernel<<<grid, 256, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_frontend_mfcc_finish(p, m, keep, stream);     // clamps against all B*nblk per-CTA maxima (m.n_max)
}

cudaError_t launch_frontend(const FrontendParams& p, const int* keep, cudaStream_t stream) {
  size_t smem = frontend_smem_bytes(p.N, p.WL, p.hop, p.fpb);
  cudaError_t e = ensure_frontend_smem(smem);
  if (e != cudaSuccess) return e;
  dim3 grid(p.nblk, p.B);
  frontend_kernel<<<grid, 256, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int rows = 64;
  dim3 g2((p.T + rows - 1) / rows, p.B);
  cmn_mask_kernel<<<g2, 128, 0, stream>>>(p.feats, p.partial, keep, p.T, p.F, p.nblk, rows);
  return cudaGetLastError();
}

}  // namespace vpb
