/* vpb200.h -- C ABI of the B200-native speaker-embedding extraction path (sm_100a).
 *
 * The reference (yeyupiaoling/VoiceprintRecognition-Pytorch, mvector 1.1.1) is pure Python and has NO FFI /
 * plugin boundary; the seam it offers is two Python callables:
 *     seam 1  AudioFeaturizer.forward(waveforms[, lens_ratio]) -> [B,T,F]   mvector/data_utils/featurizer.py:53-91
 *     seam 2  predictor(features) -> [B, embd_dim]                          mvector/predict.py:228,262
 * driven by MVectorPredictor.predict / predict_batch (mvector/predict.py:214-265).  This header is the C ABI a
 * maintainer binds (ctypes stub in INTEGRATION.md) to replace exactly those two callables:
 *     vp_fbank / vp_melspec   <-> seam 1 (KaldiFbank: featurizer.py:119-132 -> torchaudio kaldi.py:514-645;
 *                                          MelSpectrogram: featurizer.py:41-42,76; CMN + mask: featurizer.py:77-90)
 *     vp_embed                <-> seam 2 (nn.Sequential(build_model(...)).eval(), predict.py:54-63)
 *     vp_embed_wave           <-> seam 1 + seam 2 back to back (predict.py:256-262)
 *
 * Conventions: all tensor arguments are caller-owned DEVICE pointers to float32 (row-major, dense); every call
 * takes the CUDA stream to enqueue on (a cudaStream_t passed as void*, NULL = legacy default stream) and is
 * asynchronous; functions return 0 on success or a VP_ERR_* code, with vp_last_error() giving the message.
 * No hidden device allocation after vp_program_create.  One handle per device; thread-compatible: ONE thread and ONE
 * stream at a time per handle -- all programs of a handle share one workspace arena (sized to the largest of them), so
 * two programs of the same handle must never be in flight on different streams.  There is no CPU fallback anywhere
 * behind this ABI.
 */
#ifndef VPB200_H_
#define VPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_ABI_VERSION 4

enum {
  VP_OK = 0,
  VP_ERR_INVALID = 1,      /* bad argument / inconsistent op */
  VP_ERR_CUDA = 2,         /* a CUDA runtime call failed (message has the cudaError string) */
  VP_ERR_NOMEM = 3,
  VP_ERR_UNSUPPORTED = 4   /* shape/option outside what the kernels implement; never silently emulated */
};

typedef struct vp_handle vp_handle;
typedef struct vp_program vp_program;

/* ------------------------------------------------------------------------------------------------------------
 * Front-end (seam 1).  kind 0 = Kaldi Fbank framing (snip_edges, per-frame DC removal, pre-emphasis, window,
 * zero-pad to n_fft; kaldi.py:154-217), kind 1 = torch.stft framing (reflect-centred frames of n_fft samples,
 * functional.py:123-135).  Then |rFFT|^2 (power 2) or |rFFT| (power 1), sparse triangular mel projection,
 * optional log(max(x, log_floor)), then (featurizer.py:77-90) time-mean subtraction over ALL T frames and zeroing
 * of frames >= keep_frames[b].
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct vp_frontend_desc {
  int32_t kind;         /* 0 kaldi-fbank framing, 1 centred-STFT framing */
  int32_t n_fft;        /* FFT size N = 2^a 3^b 5^c, multiple of 4, in [64, 2048]; kind 0 needs a power of two */
  int32_t win_length;   /* samples taken per frame (<= n_fft); window[] has this many taps */
  int32_t hop;          /* frame shift in samples */
  int32_t n_mels;       /* filter count: <= 128 mel filters, or n_fft/2+1 (<= 1025) pass-through bins for Spectrogram */
  int32_t remove_dc;    /* kind 0: subtract the frame mean */
  float   preemph;      /* kind 0: pre-emphasis coefficient (0 = off) */
  int32_t power;        /* 2 = power spectrum, 1 = magnitude */
  int32_t use_log;      /* 0 none; 1 ln(max(x, log_floor)); 2 db_mult*log10(max(x, log_floor)); 3 ln(x + log_floor) */
  float   log_floor;
  int32_t post;         /* 0: the filter outputs are the features; 1: MFCC = DCT-II over the n_mels log values */
  int32_t n_out;        /* post 1: cepstral coefficients kept (<= n_mels); ignored otherwise */
  float   db_mult;      /* use_log 2 multiplier (10 for power, AmplitudeToDB) */
  float   top_db;       /* post 1: clamp the log values to (max over the whole call) - top_db first; < 0 = no clamp */
} vp_frontend_desc;

int vp_create(int device, vp_handle** out);
void vp_destroy(vp_handle* h);
const char* vp_last_error(const vp_handle* h);
int vp_abi_version(void);
int32_t vp_sizeof_op(void);             /* binding self-check: sizeof(vp_op) */
int32_t vp_sizeof_frontend_desc(void);  /* binding self-check: sizeof(vp_frontend_desc) */

/* window: win_length floats.  Mel bank in CSR-like form: filter m covers FFT bins
 * [mel_start[m], mel_start[m] + mel_count[m]) with weights mel_w[mel_off[m] ...]; dct: [n_mels, n_out] row-major
 * (torchaudio create_dct layout) when desc->post == 1, else NULL; all host pointers, copied. */
int vp_frontend_set(vp_handle* h, const vp_frontend_desc* desc, const float* window, const int32_t* mel_start,
                    const int32_t* mel_count, const int32_t* mel_off, const float* mel_w, int32_t n_w, const float* dct);
/* feature dimension F the configured front-end emits (n_mels, or n_out for MFCC) */
int32_t vp_feature_dim(const vp_handle* h);
/* number of frames T the configured front-end yields for n_samples (kaldi.py:63-67 / torch.stft) */
int32_t vp_num_frames(const vp_handle* h, int32_t n_samples);

/* wave [B, Lpad] (zero padded to the batch max, predict.py:248-254) -> feats [B, T, F], T = vp_num_frames(Lpad).
 * keep_frames: device int32 [B] = round(len_i / Lmax * T) (featurizer.py:82-84) or NULL for no masking.
 * scratch: device floats, at least vp_frontend_scratch_floats(B, Lpad).
 * vp_fbank   = torchaudio.compliance.kaldi.fbank per utterance (featurizer.py:47-48,119-132): kind 0, post 0.
 * vp_melspec = torchaudio.transforms.MelSpectrogram, and Spectrogram with a pass-through bank (featurizer.py:41-44):
 *              kind 1, post 0.
 * vp_mfcc    = torchaudio.transforms.MFCC (featurizer.py:45-46): kind 1, post 1.  The top_db clamp uses the maximum
 *              over ALL B utterances of the call, as torchaudio does for a [B, n_mels, T] input. */
size_t vp_frontend_scratch_floats(const vp_handle* h, int32_t B, int32_t Lpad);
int vp_fbank(vp_handle* h, const float* wave, int32_t B, int32_t Lpad, const int32_t* keep_frames, float* feats,
             float* scratch, void* stream);
int vp_melspec(vp_handle* h, const float* wave, int32_t B, int32_t Lpad, const int32_t* keep_frames, float* feats,
               float* scratch, void* stream);
int vp_mfcc(vp_handle* h, const float* wave, int32_t B, int32_t Lpad, const int32_t* keep_frames, float* feats,
            float* scratch, void* stream);

/* vp_mfcc in two stages, for callers that split ONE reference call over several processes (utterance sharding,
 * SURVEY.md 8e): the top_db clamp of torchaudio's MFCC uses the maximum over the WHOLE call, so a sharded call computes
 *   vp_mfcc_mel    : mel dB values of its shard into `scratch`, their maximum into max_out[0] (device float),
 *   (caller)       : all-reduce(MAX) of that one scalar over the ranks (ncclAllReduce / torch.distributed),
 *   vp_mfcc_finish : clamp to max_in[0] - top_db, DCT-II, CMN, mask -> feats.
 * `scratch` (vp_frontend_scratch_floats(B, Lpad) floats) must be left untouched between the two calls.
 * vp_mfcc == vp_mfcc_mel + vp_mfcc_finish with max_in = max_out. */
int vp_mfcc_mel(vp_handle* h, const float* wave, int32_t B, int32_t Lpad, float* scratch, float* max_out, void* stream);
int vp_mfcc_finish(vp_handle* h, int32_t B, int32_t Lpad, const int32_t* keep_frames, float* feats, float* scratch,
                   const float* max_in, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backbone (seam 2).  The host (Python mirror of the reference's mvector/models modules) lowers a model + a concrete (B, T) to a
 * straight-line program of fused ops over a workspace arena; weights live in one packed arena uploaded once.
 * Activations are channel-last: 1-D maps are [B, T, C], 2-D maps are [B, T, F, C] (T = time = conv2d W axis,
 * F = frequency = conv2d H axis of the reference's [B, C, F, T]).
 * ---------------------------------------------------------------------------------------------------------- */
enum {                       /* vp_op.kind */
  VP_OP_CONV = 1,            /* implicit-GEMM conv1d/conv2d/linear with fused prologue + epilogue */
  VP_OP_CONV_C1 = 2,         /* KTxKF (<= 7x7) strided conv2d with Cin = 1 on the feature map [B,T,F] -> [B,T',F',C] */
  VP_OP_COLSTATS = 3,        /* per-utterance column statistics over rows (mean / mean+std variants / segments) */
  VP_OP_ASP_POOL = 4,        /* softmax over T of logits, attentive mean + std (pooling.py:122-126); mode 1 = mean only (SAP, pooling.py:62-64) */
  VP_OP_EW = 5,              /* elementwise: gate*x + residual, AFF blend, copy */
  VP_OP_POOL2D = 6           /* KTxKF max / average pooling on a channel-last 2-D map (res2net.py:33-34,105); mode 0 = max
                                (implicit -inf padding), 1 = average with count_include_pad (zero padding, divide by KT*KF) */
};
enum { VP_ACT_NONE = 0, VP_ACT_RELU = 1, VP_ACT_HARDTANH20 = 2, VP_ACT_SIGMOID = 3, VP_ACT_TANH = 4, VP_ACT_SILU = 5 };
enum { VP_PAD_ZERO = 0, VP_PAD_REFLECT = 1 };
enum { VP_SRC2_NONE = 0, VP_SRC2_ADD = 1, VP_SRC2_CONCAT = 2 };
enum {                       /* VP_OP_COLSTATS modes (op.mode) */
  VP_STATS_MEAN = 0,               /* out[b, c] = mean over rows                                  (ecapa_tdnn.py:79, resnet_se.py:58-60) */
  VP_STATS_MEAN_STD_CLAMP = 1,     /* [mean ; sqrt(clamp(sum((x-mean)^2)/R, eps))]                (pooling.py:91-94,108) */
  VP_STATS_MEAN_STD_UNBIASED = 2,  /* [mean ; sqrt(sum((x-mean)^2)/(R-1))]                        (campplus.py:27-33) */
  VP_STATS_MEAN_STD_TSTP = 3,      /* [mean ; sqrt(sum((x-mean)^2)/(R-1) + eps)]                  (pooling.py:140-148) */
  VP_STATS_SEG_CONTEXT = 4,        /* out[b, s, c] = mean over rows + mean over segment s (ceil)  (campplus.py:96-111) */
  VP_STATS_MEAN_VAR_UNBIASED = 5   /* [mean ; sum((x-mean)^2)/(R-1)]  (TemporalStatisticsPooling returns the VARIANCE, pooling.py:44-46) */
};
enum { VP_EW_GATE_RES = 0, VP_EW_AFF = 1, VP_EW_COPY = 2,
       VP_EW_PAD_COPY = 3 };  /* dst[r, 0:Cout] = (src[r, 0:Cin], zeros): any Cin / in_ld; Cout % 4 == 0 (odd feature dims) */
enum { VP_BUF_NONE = -1, VP_BUF_INPUT = -2, VP_BUF_OUTPUT = -3 };  /* special values for activation offsets */
enum { VP_ENGINE_AUTO = 0,   /* vp_op.engine: fastest eligible of the three below (TC16 > TC > FFMA) */
       VP_ENGINE_FFMA = 1,   /* exact fp32 FFMA kernels */
       VP_ENGINE_TC = 2,     /* tcgen05, three-pass split TF32 */
       VP_ENGINE_TC16 = 3 }; /* tcgen05, three-pass two-term FP16 split with dynamic power-of-two activation scaling */

typedef struct vp_op {
  int32_t kind;
  int32_t mode;            /* COLSTATS / EW sub-mode */
  int32_t engine;          /* VP_OP_CONV: VP_ENGINE_* */
  int32_t B;               /* utterances */
  /* activation operands: byte offsets into the program workspace, or VP_BUF_* */
  int64_t src, src2, dst, res, gate, ubias;
  /* weight-arena operands: byte offsets (or -1) */
  int64_t w, bias, pre_s, pre_h, post_s, post_h;
  /* optional tensor-core image of w (or -1): split-TF32 hi/lo planes, tiled [n_tile][k_block][hi|lo][tc_bn rows][32]
   * floats with the 16-byte chunks of every 128-byte row XOR-swizzled by (row & 7) -- the UMMA SWIZZLE_128B K-major
   * shared-memory image, so one bulk-async copy lands a pipeline stage (see conv_tc.cu, mvector/engine.py::pack_tc) */
  int64_t w_tc;
  /* optional accumulate-into view (or VP_BUF_NONE): after the epilogue, sum[m, sum_coff + n] += y[m, n].  Lets a Res2
   * chain hand "x_{j+1} + y_j" to the next conv as ONE source (in place over x_{j+1}) instead of gathering two. */
  int64_t sum;
  /* source geometry: rows = B*Tin*Fin, each row in_ld floats, channels [in_coff, in_coff+Cin) */
  int32_t Tin, Fin, Cin, in_ld, in_coff;
  int32_t src2_mode, src2_ld, src2_coff, Cin2;   /* ADD: same Cin; CONCAT: channels Cin..Cin+Cin2 come from src2 */
  /* destination geometry: rows = B*Tout*Fout, row out_ld floats, channels [out_coff, out_coff+Cout) */
  int32_t Tout, Fout, Cout, out_ld, out_coff;
  int32_t res_ld, res_coff;
  /* taps (time, freq), strides, dilations, paddings */
  int32_t KT, KF, sT, sF, dT, dF, padT, padF, pad_mode;
  int32_t w_ld;            /* floats per weight row; weight element (n, (kt*KF+kf)*CinTot + ci) */
  int32_t pre_relu;        /* prologue: a = relu(a*pre_s[ci] + pre_h[ci]) when pre_s >= 0 (campplus.py:141-149) */
  int32_t act, act2;       /* y = act2(act(acc + bias + ubias) * post_s + post_h) * gate + res)  -- see DESIGN.md */
  int32_t seg_len, n_seg;  /* gate / ubias / SEG_CONTEXT rows per utterance: row (b*n_seg + min(t/seg_len, n_seg-1)) */
  float   eps;
  int32_t tc_bn;           /* N tile of w_tc: 256 if Cout >= 256 and tc_kc == 0, else 128 if Cout >= 128, else Cout rounded up to 16 */
  int32_t sum_ld, sum_coff;
  /* fp16 two-term image of w for the kind::f16 variant of the tcgen05 engine (VP_ENGINE_TC16; VPB_TC_F16=0 disables it
   * at run time).  w_tc16_q = (byte offset in the weight arena >> 4) + 1, 0 = none; layout
   * [n_tile][k_block of 64][hi|lo][tc_bn rows][64 halves], 16-byte chunks XOR-swizzled by (row & 7); the image holds
   * w * 2^k, tc16_descale = 2^-k is applied to the accumulator. */
  int32_t w_tc16_q;
  float   tc16_descale;
  /* Dynamic activation range for the fp16 split: "amax slots" are uint32 words (float bits, zeroed at the start of
   * every vp_embed) behind the program workspace.  An op with amax_out = s + 1 atomically maxes |y| over everything it
   * writes into slot s (max is exact and order independent: results stay deterministic); a CONV with amax_in = s + 1 may
   * run on the fp16 split and then scales its gathered activations by the power of two that puts slot s's maximum just
   * below 2^14, so the split is range-safe whatever the magnitude of the activations.  0 = none: the op then never runs
   * on the fp16 split (it stays on split-TF32, which has fp32's range). */
  int32_t amax_out, amax_in;
  /* Accumulation chunk of the tcgen05 engines: 0 = the whole K extent goes into one TMEM accumulator; > 0 (a multiple of
   * 64) = every tc_kc K elements go into a fresh accumulator and the epilogue warps fold the chunks with correctly
   * rounded fp32 adds.  The tensor core truncates on every accumulate (a bias that grows linearly with the chain
   * length); deep networks set a short chunk on their long-K layers.  A chunked op uses tc_bn <= 128 (third TMEM region). */
  int32_t tc_kc;
  int32_t reserved0;
} vp_op;

/* Upload the packed fp32 weight arena (host pointer, copied to the device; replaces any previous arena). */
int vp_weights_load(vp_handle* h, const void* host_blob, size_t nbytes);

/* Validate + own a program for fixed (B, T).  Its workspace is the handle's shared arena, grown (after draining the
 * device) when workspace_bytes exceeds the current arena -- serving ragged lengths therefore costs max, not sum, of the
 * programs' workspaces, and destroying a program frees host memory only. */
int vp_program_create(vp_handle* h, const vp_op* ops, int32_t n_ops, size_t workspace_bytes, size_t input_floats,
                      size_t output_floats, vp_program** out);
void vp_program_destroy(vp_program* p);
/* feats [B,T,F] -> emb [B, embd_dim] */
int vp_embed(vp_program* p, const float* feats, float* emb, void* stream);
/* wave [B,Lpad] -> emb: front-end then program; feats_scratch holds B*T*F floats, fe_scratch as for vp_fbank */
int vp_embed_wave(vp_program* p, const float* wave, int32_t B, int32_t Lpad, const int32_t* keep_frames,
                  float* feats_scratch, float* fe_scratch, float* emb, void* stream);
/* Same as vp_embed but brackets every op with CUDA events on `stream`, synchronises, and writes the per-op device
 * time in milliseconds to the HOST array ms_per_op[n_ops] (bench.py's live roofline measurement). */
int vp_embed_profiled(vp_program* p, const float* feats, float* emb, void* stream, float* ms_per_op);
/* op i of the program: kind, GEMM view (M rows, N = Cout, K = taps*Cin; K = 0 for non-conv ops), resolved engine */
int vp_program_op_info(const vp_program* p, int32_t i, int32_t* kind, int64_t* M, int64_t* N, int64_t* K, int32_t* engine);
/* Host utility (no CUDA): gather n waveforms (host pointers srcs[i], lens[i] samples) into the zero-padded row-major
 * staging matrix dst[n, lmax] with n_threads worker threads -- the pad-to-longest loop of predict.py:248-254. */
int vp_host_gather_pad(const float* const* srcs, const int32_t* lens, int32_t n, int32_t lmax, float* dst,
                       int32_t n_threads);
/* cudaMemsetAsync(device_ptr, 0, nbytes) on `stream`: zero padding of feature batches by hosts that must not bring
 * their own kernels (collate_fn.py:12-19 pads FEATURES, not waveforms). */
int vp_device_zero(void* device_ptr, size_t nbytes, void* stream);

/* Cosine score matrix scores[i, j] = <a_i, b_j> / (|a_i| |b_j|), a [n, D], b [m, D], scores [n, m] (device, row-major):
 * the scoring step of the callers around the embedding path -- retrieval against the enrolled speaker means
 * (predict.py:169-183), evaluate's trial-vs-enrol scores (trainer.py:454-461) and the diarization similarity matrix
 * (infer_utils/speaker_diarization.py:254-257). */
int vp_cosine_scores(vp_handle* h, const float* a, int32_t n, const float* b, int32_t m, int32_t D, float* scores, void* stream);

/* Staging half of predict_batch (predict.py:244-255) as ONE native call: worker threads gather slices of slice_rows
 * utterances into the zero-padded PINNED matrix staging[n, lmax]; the calling thread -- one of the n_threads gatherers --
 * issues cudaMemcpyAsync(staging slice -> device_dst slice) on copy_stream, in slice order, as soon as a slice is
 * complete, so the H2D transfer of slice k overlaps the gather of slice k+1 (also with n_threads == 1).  Returns when the
 * last copy has been ENQUEUED.  device_dst == NULL: gather only (== vp_host_gather_pad). */
int vp_host_stage_h2d(const float* const* srcs, const int32_t* lens, int32_t n, int32_t lmax, float* staging,
                      float* device_dst, int32_t slice_rows, int32_t n_threads, void* copy_stream);
/* Process-wide switch of the staging gather: on != 0 -> rows are written with non-temporal (streaming) stores, which skip
 * the read-for-ownership of the pinned destination lines (the copy engine, not a CPU, reads them next): less DRAM traffic
 * when several ranks of one host stage at the same time.  Returns 1 when streaming stores are in effect (x86-64 with AVX2
 * or AVX-512), 0 otherwise (plain memcpy).  on < 0: query only. */
int vp_host_gather_streaming(int on);
/* bytes of the handle's shared workspace arena right now */
size_t vp_workspace_bytes(const vp_handle* h);
/* number of kernel launches one vp_embed enqueues (bench.py's gpu_launches) */
int32_t vp_program_launches(const vp_program* p);
/* debugging / tests: copy a workspace region to a caller device buffer on the stream */
int vp_program_peek(vp_program* p, int64_t byte_offset, size_t nbytes, void* dst_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPB200_H_ */
