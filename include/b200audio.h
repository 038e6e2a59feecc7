/*
 * b200audio — C ABI of the B200-native STFT / log-mel / iSTFT hot path.
 *
 * This is the drop-in boundary.  The reference (mlx-audio-plus 0.1.8) has no FFI:
 * its seam is the Python module mlx_audio/dsp.py plus each model's log_mel_spectrogram.  Our Python
 * package mirrors that module API and binds THIS header through ctypes
 * (mlx_audio_plus_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All `d_*` pointers are device pointers in the
 *     CURRENT CUDA context (one plan per GPU); `h_*` are host pointers.
 *   - every entry point returns a b2a_status (0 = ok, negative = error); b2a_last_error() returns a
 *     thread-local message.  The three Python ValueErrors of the reference map to
 *     B2A_ERR_UNKNOWN_WINDOW (dsp.py:109,175), B2A_ERR_PAD_MODE (dsp.py:126), B2A_ERR_TOO_SHORT
 *     (dsp.py:132-136); broadcast failures (window longer than n_fft, dsp.py:141/197) -> B2A_ERR_SHAPE.
 *   - launches are stream-ordered on the given cudaStream_t (passed as void*), no hidden syncs;
 *     the caller owns every buffer, the library owns only the plan (twiddles, window, filterbank).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     B2A_ERR_CUDA.
 */
#ifndef B200AUDIO_H
#define B200AUDIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2A_VERSION 100 /* 0.1.0 */

typedef enum b2a_status {
  B2A_OK = 0,
  B2A_ERR_INVALID_ARG = -1,
  B2A_ERR_UNKNOWN_WINDOW = -2, /* ValueError("Unknown window function")  dsp.py:109,175 */
  B2A_ERR_PAD_MODE = -3,       /* ValueError("Invalid pad_mode")         dsp.py:126 */
  B2A_ERR_TOO_SHORT = -4,      /* ValueError("Input is too short")       dsp.py:132-136 */
  B2A_ERR_SHAPE = -5,          /* broadcast / shape mismatch (MLX shape error in the reference) */
  B2A_ERR_CUDA = -6,
  B2A_ERR_UNSUPPORTED = -7,
  B2A_ERR_NOMEM = -8
} b2a_status;

/* ---- enumerations ------------------------------------------------------------------------------ */
enum { B2A_WIN_HANN = 0, B2A_WIN_HAMMING = 1, B2A_WIN_BLACKMAN = 2, B2A_WIN_BARTLETT = 3 };
enum { B2A_PAD_REFLECT = 0, B2A_PAD_CONSTANT = 1 };
enum { /* what is taken from each STFT bin */
  B2A_SPEC_COMPLEX = 0,   /* X (complex64)                      dsp.stft */
  B2A_SPEC_POWER = 1,     /* |X|^2        whisper/audio.py:77, parakeet/audio.py:58 */
  B2A_SPEC_MAGNITUDE = 2, /* |X|          vocos/mel.py:23 */
  B2A_SPEC_SQRT_POWER_EPS = 3 /* sqrt(|X|^2 + eps)  qwen3_tts.py:85 */
};
enum { B2A_LOG_NONE = 0, B2A_LOG_LOG10 = 1, B2A_LOG_LN = 2 };
enum { B2A_DTYPE_F32 = 0, B2A_DTYPE_F16 = 1, B2A_DTYPE_BF16 = 2 }; /* output element types */
enum { B2A_GUARD_NONE = 0, B2A_GUARD_MAX = 1 /* max(x,eps) */, B2A_GUARD_ADD = 2 /* x+eps */ };
enum {
  B2A_CLAMP_NONE = 0,
  B2A_CLAMP_CLIP_MAX = 1,  /* max(y, max_over_clip(y) - v)     whisper/audio.py:83 */
  B2A_CLAMP_BATCH_MAX = 2, /* max(y, max_over_batch(y) - v)    s3tokenizer/utils.py:131 */
  B2A_CLAMP_FIXED = 3      /* max(y, v)                        voxtral_realtime/audio.py:91-92 */
};
enum {
  B2A_NORM_NONE = 0,
  B2A_NORM_PER_FEATURE = 1, /* (y-mean_t)/(std_t+eps) per mel   parakeet/audio.py:66-69 */
  B2A_NORM_GLOBAL = 2       /* one mean/std per clip            parakeet/audio.py:70-73 */
};
enum { B2A_LAYOUT_TM = 0 /* (T, M) */, B2A_LAYOUT_MT = 1 /* (M, T) */ };
enum { B2A_ISTFT_NORM_WINDOW = 0 /* sum w  dsp.py:199 normalized=False */,
       B2A_ISTFT_NORM_WINDOW_SQ = 1 /* sum w^2 */ };
enum { /* how istft treats the envelope: */
  B2A_ISTFT_DIV_WHERE = 0, /* num/den where den>1e-10 else num     dsp.py:207-209 */
  B2A_ISTFT_DIV_CLAMP = 1  /* num/max(den,1e-10)                   dsp.py:344,407 (ISTFTCache) */
};
enum {
  B2A_ISTFT_INPUT_COMPLEX = 0,
  B2A_ISTFT_INPUT_POLAR = 1
};

/* ---- front-end (forward) plan -------------------------------------------------------------------
 * One descriptor expresses every row of SURVEY.md Appendix A "wrapper parameter matrix":
 *   x -> [right pad to `length` with pad_value] -> [preemphasis] -> centre pad (reflect|constant)
 *     -> frames (hop) * window -> rFFT(n_fft) -> spec_kind -> [mel filterbank] -> [guard,log]
 *     -> [clamp] -> [affine (y+add)/div] -> [normalise] -> layout
 * n_mels == 0 means "no mel projection": the output is the (T, n_fft/2+1) spectrum itself
 * (complex64 for B2A_SPEC_COMPLEX == dsp.stft, dsp.py:92-141; float32 otherwise). */
typedef struct b2a_frontend_desc {
  int32_t n_fft;
  int32_t hop;
  int32_t center;   /* dsp.py:128 */
  int32_t pad_mode; /* B2A_PAD_*  dsp.py:118-126 */
  int32_t window_len; /* taps in `window` passed to create (<= n_fft; zero-extended on the RIGHT, dsp.py:114-116) */
  float preemph;      /* 0 = off; y[0]=x[0], y[n]=x[n]-a*x[n-1]  parakeet/audio.py:53-55 */
  int32_t drop_last;  /* drop the final frame  whisper/audio.py:77 */
  int32_t spec_kind;  /* B2A_SPEC_* */
  float spec_eps;
  int32_t n_mels;     /* 0 = spectrum output */
  int32_t log_kind;   /* B2A_LOG_* */
  int32_t guard_kind; /* B2A_GUARD_* */
  float guard_eps;
  int32_t clamp_kind; /* B2A_CLAMP_* */
  float clamp_value;
  float affine_add; /* applied iff affine_div != 0: (y + add) / div   whisper/audio.py:84 */
  float affine_div;
  int32_t norm_kind; /* B2A_NORM_* */
  int32_t norm_ddof; /* 0 (mx.std) or 1 (sortformer.py:105-108) */
  float norm_eps;
  int32_t out_layout; /* B2A_LAYOUT_* */
  /* Kaldi-style PER-FRAME pre-processing (compute_fbank_kaldi, dsp.py:577-676), applied to the first `frame_len`
   * samples of every frame before the window; all off when zero.  Served by the generic kernel. */
  int32_t frame_len;    /* samples of a frame that carry signal (window_size); 0 = n_fft.  Dither, DC removal and
                           per-frame pre-emphasis act on these; the window zero-extends the rest (dsp.py:652-656) */
  int32_t frame_dc;     /* 1: subtract the frame's own mean (dsp.py:624-626) */
  float frame_preemph;  /* y[0]=x[0], y[k]=x[k]-a*x[k-1] WITHIN the frame, after DC removal (dsp.py:628-632) */
  float dither;         /* add dither * N(0,1), drawn independently per frame element (dsp.py:619-622); the stream
                           is Philox keyed by b2a_forward_args.seed — reproducible, not MLX's generator */
  /* B2A_DTYPE_*: element type of the FEATURES.  F16 / BF16 write the encoder's input dtype straight from the fused
   * kernel's epilogue (the `.astype(self.dtype)` of whisper/whisper.py:994-996 without a second pass; bit-identical to
   * casting the float32 result: the clamp commutes with the monotone cast).  Supported by the 400/160 generated-mel
   * kernels (Whisper, FunASR, Voxtral-RT, S3Tokenizer front-ends) without cross-frame normalisation; anything else
   * returns B2A_ERR_UNSUPPORTED at create time. */
  int32_t out_dtype;
} b2a_frontend_desc;

enum { B2A_PCM_F32 = 0, B2A_PCM_I16 = 1 }; /* sample formats (b2a_forward_args.audio_kind, b2a_resample_args.in_kind) */

/* Arguments of one forward launch over `batch` equal-length clips.
 * Long-form frame-range sharding: a rank that owns frames [frame_begin, frame_begin+frame_count)
 * of a signal of GLOBAL length `length` passes a slice whose first element is global sample
 * `sample_offset`; the slice must cover the samples those frames touch (halo of n_fft-hop, +1 with
 * preemphasis); padding applies at the global ends only. */
typedef struct b2a_forward_args {
  const float* audio;    /* device (or host for *_host) pointer to clip 0 */
  int64_t clip_stride;   /* elements between consecutive clips */
  int64_t length;        /* global signal length per clip INCLUDING virtual right padding */
  int64_t valid_length;  /* samples really present (global index < valid_length), rest reads pad_value */
  float pad_value;       /* parakeet pad_to / whisper `padding` (0) */
  int32_t batch;
  int64_t sample_offset; /* global index of audio[0] (0 when not sharded) */
  int64_t frame_begin;   /* first frame to compute */
  int64_t frame_count;   /* number of frames to compute, -1 = all (after drop_last) */
  void* out;             /* features (float32) or spectrum (complex64 / float32), local frames only */
  int64_t out_clip_stride; /* elements (float32, or complex64 for SPEC_COMPLEX) between clips; 0 = dense */
  float* clip_max;       /* optional [batch] float: running max of y before clamp (device); NULL = internal */
  double* feat_sums;     /* optional [batch][2*n_mels] double: sum, sum of squares per mel (device) */
  void* workspace;       /* device scratch of b2a_frontend_call_workspace_bytes() or NULL to let the plan own it */
  size_t workspace_bytes;
  uint64_t seed;         /* dither stream (used only when desc.dither != 0) */
  int32_t audio_kind;    /* B2A_PCM_F32 (0, default): float32 samples.  B2A_PCM_I16: `audio` points at int16 PCM as a decoder
                          * delivers it (audio_io.py:258-262: float32 = int16 / 32768, exact) — b2a_frontend_forward_host only:
                          * half the host-to-device bytes, converted on the device in front of the fused kernel. */
  int32_t reserved0;
} b2a_forward_args;

typedef struct b2a_plan b2a_plan;

/* ---- library ------------------------------------------------------------------------------------ */
int b2a_version(void);
const char* b2a_last_error(void);
int b2a_device_count(void); /* 0 when no CUDA device / driver: compute calls will fail loudly */
unsigned long long b2a_launch_count(void); /* kernels this library has launched in this process (all plans, all streams) */

/* ---- host-side tables (no GPU needed) ------------------------------------------------------------
 * b2a_window: dsp.py:33-79 — float64 cosine evaluated per tap, rounded to float32.
 * b2a_mel_filters: dsp.py:223-296 — float32 arithmetic in the reference's order; `out` is
 *   row-major (n_mels, n_fft/2+1).  f_max <= 0 means sample_rate/2 (dsp.py:264). */
int b2a_window(int kind, int size, int periodic, float* h_out);
int b2a_mel_filters(int sample_rate, int n_fft, int n_mels, double f_min, double f_max,
                    int norm_slaney, int scale_htk, float* h_out);

/* ---- geometry (pure integer arithmetic, no GPU needed) -------------------------------------------
 * b2a_stft_geometry: dsp.py:118-136.  Returns B2A_ERR_TOO_SHORT exactly when the reference raises.
 * b2a_frame_source_index: the bit-exact framing contract — source sample index of tap k of frame t
 *   (-1 = literal zero from constant padding). */
int b2a_stft_geometry(int64_t length, int n_fft, int hop, int center, int pad_mode,
                      int64_t* padded_len, int64_t* num_frames);
int64_t b2a_frame_source_index(int64_t length, int n_fft, int hop, int center, int pad_mode,
                               int64_t t, int k);
int b2a_istft_geometry(int64_t num_frames, int n_fft, int hop, int center, int64_t length /* -1 = None */,
                       int64_t* ola_len, int64_t* out_start, int64_t* out_len);

/* ---- forward: STFT / spectrogram / log-mel front-ends --------------------------------------------
 * h_window: window_len taps; h_filterbank: (n_mels, n_fft/2+1) row-major or NULL when n_mels == 0. */
int b2a_frontend_create(const b2a_frontend_desc* desc, const float* h_window,
                        const float* h_filterbank, b2a_plan** plan);
int b2a_plan_destroy(b2a_plan* plan);
int b2a_frontend_out_frames(const b2a_plan* plan, int64_t length, int64_t* frames /* after drop_last */);
size_t b2a_frontend_workspace_bytes(const b2a_plan* plan, int32_t batch);
/* Scratch ONE call needs (statistics + per-tile minima for these args).  A call whose args->workspace has at least this
 * many bytes touches nothing of the plan's own scratch, so one plan may serve concurrent streams / threads (each with its
 * own workspace; partial() and finalize() of one call share it).  With a smaller or NULL workspace the plan's scratch
 * is used and calls on one plan must not overlap. */
size_t b2a_frontend_call_workspace_bytes(const b2a_plan* plan, const b2a_forward_args* args);
/* forward = partial + finalize (clamp / normalise) in one stream-ordered call */
int b2a_frontend_forward(b2a_plan* plan, const b2a_forward_args* args, void* stream);
/* split form for frame-range sharding: partial() leaves un-clamped / un-normalised values in `out`
 * and the statistics in args->clip_max / args->feat_sums; the caller reduces those across ranks
 * (max / sum) and calls finalize() with the GLOBAL frame count. */
int b2a_frontend_partial(b2a_plan* plan, const b2a_forward_args* args, void* stream);
int b2a_frontend_finalize(b2a_plan* plan, const b2a_forward_args* args, int64_t global_frames, void* stream);
/* host-buffer entry (what a NumPy caller hits): H2D, compute, D2H pipelined on internal streams over
 * chunks of clips; synchronous on return.  `args->audio` and `args->out` are HOST pointers here. */
int b2a_frontend_forward_host(b2a_plan* plan, const b2a_forward_args* args);
/* debug / parity: dump the windowed-or-raw frame matrix (T, n_fft) float32 for the bit-exact framing test */
/* compute_deltas_kaldi (dsp.py:439-483): d[f][t] = sum_{k=-n..n} k * x[f][t+k] / (n(n+1)(2n+1)/3), n = (win_length-1)/2,
 * over `rows` rows of `cols` samples (device pointers); edge = 1 replicates the end samples, 0 pads with zeros. */
int b2a_deltas(const float* x, float* out, int64_t rows, int64_t cols, int32_t win_length, int32_t edge, void* stream);
int b2a_frontend_dump_frames(b2a_plan* plan, const b2a_forward_args* args, int apply_window, void* stream);
/* which kernel family a plan dispatches to: "fast400", "fast512", "generic", "small", ... */
const char* b2a_plan_kernel_name(const b2a_plan* plan);

/* ---- the step in front of the path: PCM -> resampled float32 (mono) ----------------------------------
 * What load_audio does after the decoder (stt/utils.py:21-57: resample_audio = scipy.signal.resample_poly(audio, up, down,
 * padtype="edge") per channel, then mx.array(audio, float32).mean(axis=1); audio_io.py:258-262: int16 / 32768.0).
 * The polyphase filter is designed on the host (scipy's Kaiser(5.0) windowed sinc, zero-padded as resample_poly does) and
 * handed over as h_taps[j][phase] = h_padded[phase + j * up], taps_per_phase = ceil(len(h_padded) / up);
 * pre_remove = resample_poly's n_pre_remove.  Output sample n = sum_j taps[j][t % up] * x_edge[t / up - j], t = (n + pre_remove) * down. */
typedef struct b2a_resampler b2a_resampler;
typedef struct b2a_resample_args {
  const void* in;          /* device: (batch, n_in, channels) interleaved float32 or int16 */
  float* out;              /* device: (batch, n_out) when mono, else (batch, n_out, channels) */
  int64_t n_in;            /* frames per clip */
  int64_t in_clip_stride;  /* elements between clips (0 = n_in * channels) */
  int64_t out_clip_stride; /* elements between clips (0 = dense) */
  int32_t batch;
  int32_t channels;
  int32_t in_kind;         /* B2A_PCM_* */
  int32_t mono;            /* 1: mean over channels (float32) */
} b2a_resample_args;
int b2a_resampler_create(int32_t up, int32_t down, int32_t taps_per_phase, int64_t pre_remove, const float* h_taps,
                         b2a_resampler** out);
int b2a_resampler_destroy(b2a_resampler* r);
int b2a_resampler_out_len(const b2a_resampler* r, int64_t n_in, int64_t* n_out); /* ceil(n_in * up / down) */
int b2a_resample(b2a_resampler* r, const b2a_resample_args* args, void* stream);

/* ---- the steps right after the path: the layouts the encoders read -------------------------------------
 * b2a_rows_pad_cast: whisper/whisper.py:990-996  pad_or_trim(mel[seek : seek + segment_size], N_FRAMES, axis=-2).astype(dtype)
 *   in (batch, *, cols) float32 with in_clip_stride elements between clips; rows [row_begin, row_begin + rows_valid) are
 *   copied, rows up to rows_out are zero; out (batch, rows_out, cols) dense in out_dtype.
 * b2a_lfr: funasr/audio.py:84-139 apply_lfr (+ the precomputed CMVN of apply_cmvn, funasr/audio.py:166-169, when
 *   cmvn_shift / cmvn_scale [lfr_m * n_mels] are given): out (batch, ceil(frames / lfr_n), lfr_m * n_mels). */
int b2a_rows_pad_cast(const float* in, int64_t in_clip_stride, int64_t row_begin, int64_t rows_valid, int32_t cols, void* out,
                      int64_t rows_out, int32_t out_dtype, int32_t batch, void* stream);
int b2a_lfr(const float* in, int64_t in_clip_stride, int64_t frames, int32_t n_mels, int32_t lfr_m, int32_t lfr_n,
            const float* cmvn_shift, const float* cmvn_scale, float* out, int64_t out_clip_stride, int32_t batch, void* stream);
/* (batch, rows, cols) float32 -> (batch, cols, rows_out) with zero columns from `rows` on: the (B, n_mels, T_padded) layout of
 * vad/models/sortformer/sortformer.py:112-118 (`pad_to`) from (B, T, n_mels) features, one pass. */
int b2a_transpose_pad(const float* in, float* out, int64_t rows, int32_t cols, int64_t rows_out, int32_t batch, void* stream);
/* Row-wise zero-mean / unit-variance of waveforms (rows, cols): (x - mean) / den over the first valid[row] samples (device int64
 * array, or NULL = all), pad_value behind them.  den_kind 0: sqrt(var + eps) (transformers zero_mean_unit_var_norm: the Qwen3-ASR
 * extractor's do_normalize); 1: max(std, eps) (vad/models/smart_turn/smart_turn.py:196-199).  ddof 0, float64 accumulation. */
int b2a_rows_normalize(const float* in, float* out, int64_t rows, int64_t cols, const int64_t* valid, int32_t den_kind, float eps,
                       float pad_value, void* stream);
/* Phase unwrap along the last axis of (rows, cols) float32 (tts/models/kokoro/istftnet.py:418-452 mlx_unwrap, discont >= period / 2):
 * one pass, float32 prefix sum of the 2 pi corrections. */
int b2a_unwrap(const float* in, float* out, int64_t rows, int64_t cols, float discont, float period, void* stream);
/* per-utterance CMVN, funasr/audio.py:160-164: out = (x - mean_t) / (std_t + eps) per feature column, mx.std (ddof 0);
 * x, out: (batch, rows, cols) float32 (may alias); stats_ws: device scratch of 2 * cols * batch doubles. */
int b2a_cmvn_utterance(const float* in, float* out, int64_t clip_stride, int64_t rows, int32_t cols, float eps, double* stats_ws,
                       int32_t batch, void* stream);

/* ---- inverse: iSTFT with windowed overlap-add -----------------------------------------------------
 * dsp.istft (dsp.py:144-217) and ISTFTCache.istft (dsp.py:350-417). */
typedef struct b2a_istft_desc {
  int32_t n_fft;      /* == win_length at every reference call site */
  int32_t hop;
  int32_t window_len; /* <= n_fft, zero-extended on the right (dsp.py:180-181) */
  int32_t center;
  int32_t norm_kind;  /* B2A_ISTFT_NORM_* */
  int32_t div_kind;   /* B2A_ISTFT_DIV_* */
  int32_t trim_tail;  /* 1: strip n_fft/2 from BOTH ends when center && length<0 (dsp.py:211-212);
                         0: strip only the front (ISTFTCache, dsp.py:410-412) */
  float div_eps;      /* envelope floor of the division guard; 0 = the reference's 1e-10 (dsp.py:207,344).  The
                         model-local HiFT / CosyVoice iSTFTs use 1e-8 (s3gen/hifigan.py:521, cosyvoice3/hifigan.py:493) */
  int32_t input_form; /* B2A_ISTFT_INPUT_*: COMPLEX = spec (+ optional spec_imag plane); POLAR = `spec` is the MAGNITUDE
                         plane and `spec_imag` the PHASE plane: X = clip(mag) * (cos p + i sin p) is formed in the kernel
                         (kokoro/istftnet.py:500-512, s3gen/hifigan.py:480-485, cosyvoice3/hifigan.py:447-452) */
  float mag_clip_max; /* POLAR: magnitude is clipped to <= this (1e2 in the HiFT heads); <= 0 = no upper clip */
  int32_t mag_clip_min_zero; /* POLAR: 1 = clip magnitude to >= 0 (cosyvoice3/hifigan.py:447 a_min=0.0) */
  int32_t mag_log;    /* POLAR: 1 = the magnitude plane holds LOG-magnitudes, mag = exp(a) before the clip — the Vocos / Soprano
                         iSTFT head `clip(exp(x), max=1e2)` (codec/models/vocos/vocos.py:129-130, soprano/decoder.py:36-37) */
} b2a_istft_desc;

typedef struct b2a_inverse_args {
  const void* spec;     /* interleaved complex64 (batch, F, T) — or real plane when spec_imag != NULL */
  const void* spec_imag; /* NULL, or imag plane float32 (batch, F, T) (ISTFTCache real/imag form) */
  int64_t clip_stride;  /* elements (complex64 or float32) between clips; 0 = dense F*T */
  int64_t num_frames;   /* T */
  int32_t batch;
  int64_t length;       /* -1 = None; else out = full[:length] (istft) / [:, :audio_length] (cache) */
  float* out;           /* (batch, out_len) */
  int64_t out_clip_stride; /* 0 = dense */
} b2a_inverse_args;

int b2a_istft_create(const b2a_istft_desc* desc, const float* h_window, b2a_plan** plan);
int b2a_istft_out_len(const b2a_plan* plan, int64_t num_frames, int64_t length, int64_t* out_len);
int b2a_istft_inverse(b2a_plan* plan, const b2a_inverse_args* args, void* stream);
int b2a_istft_inverse_host(b2a_plan* plan, const b2a_inverse_args* args);

/* ---- microbenchmarks used by bench.py for roofline denominators -----------------------------------
 * FP32 FFMA peak of the current device (TFLOP/s) and a plain device copy bandwidth (GB/s). */
int b2a_measure_fp32_tflops(double* tflops, void* stream);
int b2a_measure_copy_gbs(double* gbs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200AUDIO_H */
