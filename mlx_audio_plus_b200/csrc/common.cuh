// b200audio — shared declarations for the CUDA sources (internal; the public surface is include/b200audio.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b200audio.h"

namespace b2a {

void set_error(const char* fmt, ...);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: one process driving several GPUs (plans are per device)
// must set it on each of them.  One instance per kernel instantiation (function-local static).
struct SmemAttrOnce {
  size_t set[64] = {};
  bool need(int device, size_t smem) {
    size_t& s = set[device & 63];
    if (smem <= s) return false;
    s = smem;
    return true;
  }
};
int cuda_fail(cudaError_t e, const char* what);

#define B2A_CUDA(call)                                        \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return b2a::cuda_fail(_e, #call);  \
  } while (0)

// every kernel launch of the library is counted (b2a_launch_count: bench.py's `gpu_launches` is a measurement, not a claim)
void note_launch();
#define B2A_LAUNCHED()                 \
  do {                                 \
    b2a::note_launch();                \
    B2A_CUDA(cudaGetLastError());      \
  } while (0)

constexpr int kMaxStages = 12;
constexpr int kMaxGenericRadix = 32;

enum PlanKind { PLAN_FRONTEND = 0, PLAN_ISTFT = 1 };
enum KernelFamily { KF_GENERIC = 0, KF_FAST = 1, KF_SMALL = 2 };

// Framing geometry shared by host and device (dsp.py:118-136 incl. the short-input slice quirk).
struct Geometry {
  int64_t length;      // global signal length
  int64_t pad_left;    // samples of centre padding actually produced on the left
  int64_t pad_right;
  int64_t padded_len;
  int64_t num_frames;  // T (before drop_last)
};

__host__ __device__ inline Geometry make_geometry(int64_t length, int n_fft, int hop, int center, int pad_mode) {
  Geometry g;
  g.length = length;
  int64_t p = center ? n_fft / 2 : 0;
  if (center && pad_mode == B2A_PAD_REFLECT) {
    // x[1:p+1][::-1] and x[-(p+1):-1][::-1] silently truncate to length-1 samples when length <= p
    int64_t lim = length > 0 ? length - 1 : 0;
    g.pad_left = p < lim ? p : lim;
    g.pad_right = g.pad_left;
  } else {
    g.pad_left = p;
    g.pad_right = p;
  }
  g.padded_len = length + g.pad_left + g.pad_right;
  g.num_frames = g.padded_len >= n_fft ? 1 + (g.padded_len - n_fft) / hop : 0;
  return g;
}

// padded position -> source index; -1 = literal zero
__host__ __device__ inline int64_t source_index(const Geometry& g, int pad_mode, int64_t q) {
  int64_t s = q - g.pad_left;
  if (s < 0) return pad_mode == B2A_PAD_REFLECT ? -s : -1;
  if (s >= g.length) return pad_mode == B2A_PAD_REFLECT ? 2 * g.length - 2 - s : -1;
  return s;
}

// iSTFT input forms (b2a_istft_desc.input_form): how a kernel turns the two values it read for a bin into X[k]
struct PolarSpec {
  int polar;          // 1: (a, b) = (magnitude, phase) -> clip(a) * (cos b, sin b)
  float clip_max;     // <= 0: none
  int clip_min_zero;
  int log_mag;        // 1: a holds ln(magnitude): exp first (Vocos / Soprano head)
};
#ifdef __CUDACC__
__device__ __forceinline__ float lg2_approx(float x) {  // MUFU.LG2: the fused kernels' logarithm (base change folded into the affine FFMA)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sin and cos of one float32 argument, both within 1.2 ulp of 1 (max abs error 7e-8, checked against float64 over
// |x| <= 1e5): three-term Cody-Waite reduction by pi/2 (hi / mid / lo of pi/2, one FMA each), then the degree-7 / degree-8
// minimax polynomials on [-pi/4, pi/4] and a quadrant swap — ~25 instructions where libdevice's sincosf, which carries
// the Payne-Hanek path for huge arguments inline, costs ~100 (the polar iSTFT kernels evaluate one per spectrum bin:
// C4, 270 M of them, 1.32 -> see DESIGN).  Arguments beyond 65536 rad (and NaN / inf) take libdevice's path.
__device__ __forceinline__ void sincos_cw_reduced(float x, float* sn, float* cs);
__device__ __forceinline__ void sincos_cw(float x, float* sn, float* cs) {
  if (!(fabsf(x) <= 65536.0f)) {
    sincosf(x, sn, cs);
    return;
  }
  sincos_cw_reduced(x, sn, cs);
}
// the branch-free part: valid for |x| <= 65536 (callers that test a whole group of arguments at once use it directly)
__device__ __forceinline__ void sincos_cw_reduced(float x, float* sn, float* cs) {
  const float q = rintf(x * 0.636619747f);
  float r = fmaf(-q, 0x1.921fb6p+0f, x);
  r = fmaf(-q, -0x1.777a5cp-25f, r);
  r = fmaf(-q, -0x1.ee59dap-50f, r);
  const float r2 = r * r;
  float s = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
  s = fmaf(s, r2, -1.66666546e-1f);
  s = fmaf(s * r2, r, r);
  float c = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
  c = fmaf(c, r2, 4.16666457e-2f);
  c = fmaf(c, r2, -0.5f);
  c = fmaf(c, r2, 1.0f);
  const int k = (int)q;
  const float a = (k & 1) ? c : s, b = (k & 1) ? s : c;
  *sn = (k & 2) ? -a : a;
  *cs = ((k + 1) & 2) ? -b : b;
}

__device__ __forceinline__ float2 polar_to_complex(const PolarSpec& ps, float2 v) {
  if (!ps.polar) return v;
  float m = v.x;
  if (ps.log_mag) m = expf(m);
  if (ps.clip_max > 0.0f) m = fminf(m, ps.clip_max);
  if (ps.clip_min_zero) m = fmaxf(m, 0.0f);
  float s, c;
  sincos_cw(v.y, &s, &c);  // libm-grade accuracy (the reference uses float32 cos / sin)
  return make_float2(m * c, m * s);
}
#endif
inline PolarSpec make_polar_spec(const b2a_istft_desc& d) {
  PolarSpec ps;
  ps.polar = d.input_form == B2A_ISTFT_INPUT_POLAR;
  ps.clip_max = d.mag_clip_max;
  ps.clip_min_zero = d.mag_clip_min_zero;
  ps.log_mag = d.mag_log;
  return ps;
}
// Epilogue constants of the fused log-mel kernels: a = max(x + guard_add, guard_floor); y = (use_log ? log2(a) : a) * y_mul + y_add
// (log base change and the affine map folded into one FFMA).  ONE definition: the kernels, the constant-row fill and the
// clamp fix-up's fill of unwritten (all-silent) tiles must produce bit-identical values.
struct EpilogueConsts {
  float guard_add, guard_floor, y_mul, y_add;
  int use_log;
};
inline EpilogueConsts epilogue_consts(const b2a_frontend_desc& d) {
  EpilogueConsts e;
  e.guard_add = d.guard_kind == B2A_GUARD_ADD ? d.guard_eps : 0.0f;
  e.guard_floor = d.guard_kind == B2A_GUARD_MAX ? d.guard_eps : -INFINITY;
  e.use_log = d.log_kind != B2A_LOG_NONE;
  const double lscale = d.log_kind == B2A_LOG_LOG10 ? 0.30102999566398119521 : (d.log_kind == B2A_LOG_LN ? 0.69314718055994530942 : 1.0);
  if (d.affine_div != 0.0f) {  // ((log2(a) * lscale) + add) / div
    e.y_mul = (float)(lscale / (double)d.affine_div);
    e.y_add = (float)((double)d.affine_add / (double)d.affine_div);
  } else {
    e.y_mul = (float)lscale;
    e.y_add = 0.0f;
  }
  return e;
}
#ifdef __CUDACC__
__device__ __forceinline__ float epilogue_of_zero(const EpilogueConsts& e) {  // the value every mel bin of a silent frame gets
  const float a = fmaxf(0.0f + e.guard_add, e.guard_floor);
  return fmaf(e.use_log ? lg2_approx(a) : a, e.y_mul, e.y_add);
}
#endif

inline float istft_div_eps(const b2a_istft_desc& d) { return d.div_eps > 0.0f ? d.div_eps : 1e-10f; }

struct MelCsr {  // filterbank rows as contiguous runs of non-zero taps
  int* d_start = nullptr;  // [M] first bin
  int* d_len = nullptr;    // [M] number of taps
  int* d_off = nullptr;    // [M] offset into d_w
  float* d_w = nullptr;    // [nnz_total]
  int nnz = 0;
  int max_len = 0;
};

}  // namespace b2a

struct b2a_plan {
  int kind;
  int family;
  const char* kernel_name;
  b2a_frontend_desc fd;
  b2a_istft_desc id;
  int n_fft, hop, n_freqs;
  int device, sm_count;
  // device tables
  float2* d_twiddle;  // W_N^k = exp(-2 pi i k / N), k = 0..N-1 (computed in double)
  float2* d_twiddle_passes;  // the same roots as per-pass tables [r - 1][k] = W_(Ns R)^(k r), back to back (N - 1 entries, padded to N)
  float* d_window;    // n_fft taps, zero-extended on the right
  std::vector<float> h_window;
  b2a::MelCsr mel;
  float* d_fb_dense;  // (M, F) dense copy (used by the fast kernels' own packing)
  std::vector<float> h_fb;
  int nstages;
  int radix[b2a::kMaxStages];
  // plan-owned scratch (stats etc.), grown on demand
  void* d_ws;
  size_t ws_bytes;
  // host-entry staging
  void* d_stage_in[2];
  void* d_stage_out[2];
  size_t stage_in_bytes, stage_out_bytes;
  cudaStream_t host_streams[2];
  cudaEvent_t host_events[2];
  void* fast;  // family-specific state
  float* d_tilemin[2];  // per-tile minima (clamp fix-up skips tiles with nothing below the floor)
  size_t tilemin_count[2];
};

namespace b2a {
// generic (any n_fft) kernels — generic.cu
int generic_tile_frames(const b2a_plan* plan, const b2a_forward_args* a);
int generic_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min,
                             double* feat_sums, cudaStream_t st);
int generic_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st);
int dump_frames(b2a_plan* plan, const b2a_forward_args* a, int apply_window, cudaStream_t st);
int frontend_finalize(b2a_plan* plan, const b2a_forward_args* a, int64_t global_frames, float* clip_max,
                      const float* tile_min, int tile_frames, const double* feat_sums, cudaStream_t st);
int init_stats(float* clip_max, double* feat_sums, int batch, int n_mels, cudaStream_t st);
int deltas(const float* x, float* out, int64_t rows, int64_t cols, int win_length, int edge, cudaStream_t st);
size_t generic_smem_limit(const b2a_plan* plan);
// fast (specialised two-stage register FFT) kernels — fast_fwd.cu
bool fast_frontend_supported(const b2a_plan* plan);
int fast_frontend_init(b2a_plan* plan);
void fast_frontend_destroy(b2a_plan* plan);
bool fast_frontend_out16_ok(const b2a_plan* plan);
int64_t fast_const_row0(const b2a_plan* plan, const b2a_forward_args* a);
bool fast_skip_floor_tiles(const b2a_plan* plan);
// the whole clamping forward in one cooperative launch (fast_logmel_tma_kernel<..., FUSED>); returns 1 when not applicable
int fast_frontend_fused(b2a_plan* plan, const b2a_forward_args* a, float* tile_min, float* tile_max, float* clip_max_out, cudaStream_t st);
int fast_const_rows_finalize(const b2a_plan* plan, const b2a_forward_args* a, int64_t row0, float* clip_max, cudaStream_t st);
int fast_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min,
                          double* feat_sums, cudaStream_t st);
// fused iSTFT for n_fft = 4*hop vocoder heads (1024/256) — fast_inv.cu
bool fast_istft_supported(const b2a_plan* plan);
int fast_istft_init(b2a_plan* plan);
void fast_istft_destroy(b2a_plan* plan);
int fast_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st);
// small-n (n_fft 16 / 20) thread-per-frame kernels — small.cu
bool small_istft_supported(const b2a_plan* plan);
int small_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st);
bool small_stft_supported(const b2a_plan* plan);
int small_stft(b2a_plan* plan, const b2a_forward_args* a, cudaStream_t st);
}  // namespace b2a
