// b200audio — host side of the "fast" fused log-mel family: plan tables and dispatch to the per-variant
// translation units (kernel template: fast_fwd.cuh).
#include <stdlib.h>

#include "fast_fwd.cuh"

namespace b2a {

namespace {

// Rows of all-padding frames (Whisper is always called with `padding = N_SAMPLES` zero samples behind the clip,
// whisper/whisper.py: a 30 s clip yields 6000 frames of which the last ~3000 see nothing but zeros): every mel value of such
// a frame is the same constant c = affine(log(guard(0))).  They are not transformed.  c is computed on the device exactly as
// the fused kernel's phase B does.  Without a clamp the rows are written right away; with one, c is only folded into the
// clip max by the partial step, and the finalize step writes max(c, floor) ONCE (no read-modify-write by the fix-up).
struct ConstRows {
  float guard_add, guard_floor, y_mul, y_add;
  int use_log;
  int floor_mode;      // 0: none, 1: clip_max[clip] - delta, 2: fixed
  float floor_delta, floor_fixed;
  __device__ __forceinline__ float value() const {
    const float a = fmaxf(0.0f + guard_add, guard_floor);
    return fmaf(use_log ? lg2_approx(a) : a, y_mul, y_add);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) const_rows_kernel(T* out, int64_t out_clip_stride, int64_t row0, int64_t rows, int M, ConstRows cr,
                                                          float* clip_max) {
  float c = cr.value();
  const int clip = blockIdx.y;
  if (cr.floor_mode == 1) c = fmaxf(c, clip_max[clip] - cr.floor_delta);
  else if (cr.floor_mode == 2) c = fmaxf(c, cr.floor_fixed);
  else if (clip_max && blockIdx.x == 0 && threadIdx.x == 0) atomic_max_f(clip_max + clip, c);  // caller-visible statistic
  T* o = out + (int64_t)clip * out_clip_stride + row0 * M;
  const int64_t total = rows * M;
  T cv;
  if constexpr (sizeof(T) == 4) cv = c;
  else if constexpr (std::is_same<T, __half>::value) cv = __float2half_rn(c);
  else cv = __float2bfloat16_rn(c);
  constexpr int PER = 16 / (int)sizeof(T);  // elements per 16-byte store
  if (reinterpret_cast<uintptr_t>(o) % 16 == 0 && total % PER == 0) {
    union { uint4 v; T e[PER]; } u;
#pragma unroll
    for (int k = 0; k < PER; ++k) u.e[k] = cv;
    uint4* o4 = reinterpret_cast<uint4*>(o);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total / PER; i += (int64_t)gridDim.x * blockDim.x) o4[i] = u.v;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) o[i] = cv;
  }
}

__global__ void const_fold_max_kernel(float* clip_max, int batch, ConstRows cr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) atomic_max_f(clip_max + b, cr.value());
}

ConstRows make_const_rows(const b2a_frontend_desc& d) {
  const EpilogueConsts ec = epilogue_consts(d);
  ConstRows cr;
  cr.guard_add = ec.guard_add;
  cr.guard_floor = ec.guard_floor;
  cr.use_log = ec.use_log;
  cr.y_mul = ec.y_mul;
  cr.y_add = ec.y_add;
  cr.floor_mode = 0;
  cr.floor_delta = cr.floor_fixed = 0.0f;
  return cr;
}

int launch_const_rows(const b2a_plan* plan, const b2a_forward_args* a, int64_t row0, const ConstRows& cr, float* clip_max, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  const int64_t rows = a->frame_count - row0;
  const int64_t stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * d.n_mels;
  int64_t gx = (rows * d.n_mels / 4 + 255) / 256;
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)a->batch);
  if (d.out_dtype == B2A_DTYPE_F16)
    const_rows_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<__half*>(a->out), stride, row0, rows, d.n_mels, cr, clip_max);
  else if (d.out_dtype == B2A_DTYPE_BF16)
    const_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(a->out), stride, row0, rows, d.n_mels, cr, clip_max);
  else
    const_rows_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(a->out), stride, row0, rows, d.n_mels, cr, clip_max);
  B2A_LAUNCHED();
  return B2A_OK;
}

}  // namespace

// May fast_logmel_tma_kernel leave all-silent tiles unwritten (tile_min = -inf) for the clamp fix-up to fill?  Only when a
// clamp fix-up is certain to follow and writes float32 (T, M) rows.  A pure function of the plan: partial and finalize agree.
bool fast_skip_floor_tiles(const b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  static const bool off = getenv("B2A_NO_TILE_SKIP") != nullptr;  // development toggle, read once
  return !off && plan->family == KF_FAST && d.clamp_kind != B2A_CLAMP_NONE && d.norm_kind == B2A_NORM_NONE && d.n_mels > 0 &&
         d.out_layout == B2A_LAYOUT_TM && d.out_dtype == B2A_DTYPE_F32 && d.guard_kind == B2A_GUARD_MAX && d.n_mels % 4 == 0;
}

// First row of the trailing block of all-padding frames that is filled instead of transformed (a multiple of the 32-frame
// tile), or -1.  A pure function of the plan and the call's arguments: the partial and the finalize step both derive it.
// Conditions: zero pad value, no pre-emphasis, power / magnitude spectrum, (T, M) layout, no per-feature sums, unsharded
// call, and the reflected tail must not reach back into the signal.
int64_t fast_const_row0(const b2a_plan* plan, const b2a_forward_args* a) {
  const b2a_frontend_desc& d = plan->fd;
  static const bool no_pad_skip = getenv("B2A_NO_PAD_SKIP") != nullptr;  // development toggle, read once
  if (plan->family != KF_FAST || no_pad_skip) return -1;
  const int64_t pc = d.center ? d.n_fft / 2 : 0;
  const bool tail_ok = !d.center || d.pad_mode == B2A_PAD_CONSTANT || a->length - pc - 1 >= a->valid_length;
  if (!(a->pad_value == 0.0f && d.preemph == 0.0f && d.n_mels > 0 && d.norm_kind == B2A_NORM_NONE && a->feat_sums == nullptr && tail_ok &&
        (d.spec_kind == B2A_SPEC_POWER || d.spec_kind == B2A_SPEC_MAGNITUDE) && d.out_layout == B2A_LAYOUT_TM &&
        a->frame_begin == 0 && a->sample_offset == 0 && a->valid_length < a->length))
    return -1;
  const int64_t first_pad = (a->valid_length + pc + d.hop - 1) / d.hop;  // first frame t with t*hop - pc >= valid_length
  const int64_t r0 = (first_pad + 31) / 32 * 32;
  return r0 + 32 <= a->frame_count ? r0 : -1;
}

// finalize side (clamp configured): the constant rows get max(c, floor) in one write pass
int fast_const_rows_finalize(const b2a_plan* plan, const b2a_forward_args* a, int64_t row0, float* clip_max, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  ConstRows cr = make_const_rows(d);
  const bool affine = d.affine_div != 0.0f;
  if (d.clamp_kind == B2A_CLAMP_FIXED) {
    cr.floor_mode = 2;
    cr.floor_fixed = affine ? (d.clamp_value + d.affine_add) / d.affine_div : d.clamp_value;
  } else {
    cr.floor_mode = 1;
    cr.floor_delta = affine ? d.clamp_value / d.affine_div : d.clamp_value;  // clamp_fixup_kernel: mx - clamp_value / affine_div
  }
  return launch_const_rows(plan, a, row0, cr, clip_max, st);
}

bool fast_frontend_supported(const b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  // the spectrum itself (dsp.stft): complex64 (T, F) rows through fast_stft_kernel
  const bool cplx = d.n_mels == 0 && d.spec_kind == B2A_SPEC_COMPLEX && d.out_layout == B2A_LAYOUT_TM &&
                    d.clamp_kind == B2A_CLAMP_NONE && d.norm_kind == B2A_NORM_NONE && !getenv("B2A_NO_FAST_STFT");
  if (!cplx && (d.n_mels <= 0 || d.spec_kind == B2A_SPEC_COMPLEX)) return false;
  if (d.affine_div < 0.0f) return false;
  if (d.frame_dc || d.frame_preemph != 0.0f || d.dither != 0.0f || d.frame_len != 0) return false;  // Kaldi per-frame steps
  const bool v400 = d.n_fft == 400 && d.hop == 160;
  const bool v512 = d.n_fft == 512 && d.hop == 160;
  const bool v1024 = d.n_fft == 1024 && d.hop == 256;
  const bool v800 = d.n_fft == 800 && d.hop == 200, v1024h320 = d.n_fft == 1024 && d.hop == 320;
  if (!(v400 || v512 || v1024 || v800 || v1024h320)) return false;
  if (d.n_mels > 256) return false;
  return true;
}

bool fast_frontend_out16_ok(const b2a_plan* plan) {
  const FastState* fs = reinterpret_cast<const FastState*>(plan->fast);
  const b2a_frontend_desc& d = plan->fd;
  return fs && fs->variant == 1 && fs->spec > 0 && d.n_mels > 0 && d.out_layout == B2A_LAYOUT_TM && d.norm_kind == B2A_NORM_NONE;
}

int fast_frontend_init(b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = new FastState();
  plan->fast = fs;
  int N1, N2;
  if (d.n_fft == 400) { fs->variant = 1; N1 = 20; N2 = 10; }
  else if (d.n_fft == 512) { fs->variant = 2; N1 = 16; N2 = 16; }
  else if (d.n_fft == 800) { fs->variant = 4; N1 = 20; N2 = 20; }
  else if (d.hop == 320) { fs->variant = 5; N1 = 32; N2 = 16; }
  else { fs->variant = 3; N1 = 32; N2 = 16; }
  const int NC = N1 * N2, N = 2 * NC;
  std::vector<float2> win2(NC);
  std::vector<float2> tw1(NC), twp(NC);
  for (int n2 = 0; n2 < N2; ++n2)
    for (int n1 = 0; n1 < N1; ++n1) {  // 0.5 of the real-FFT post-processing is folded into the window
      const int m = N2 * n1 + n2;
      win2[n2 * N1 + n1] = make_float2(0.5f * plan->h_window[2 * m], 0.5f * plan->h_window[2 * m + 1]);
    }
  auto expand = [](double a) { return make_float2((float)cos(a), (float)sin(a)); };
  for (int n2 = 0; n2 < N2; ++n2)
    for (int k1 = 0; k1 < N1; ++k1) tw1[n2 * N1 + k1] = expand(-2.0 * M_PI * (double)((n2 * k1) % NC) / (double)NC);
  auto wn = [&](int k) { return expand(-2.0 * M_PI * (double)k / (double)N); };
  for (int i = 0; i < NC; ++i) twp[i] = make_float2(0.f, 0.f);
  for (int u = 1; u < N1 / 2; ++u)  // unit u: W_N^(u + N1*k2), k2 = 0..N2-1
    for (int k2 = 0; k2 < N2; ++k2) twp[u * 2 * N2 + k2] = wn(u + N1 * k2);
  // unit 0 (columns 0 and N1/2 pair up within themselves; see the kernel's stage 2): slot s < N2/2 is bin N1*(s+1),
  // slot s >= N2/2 is bin N1/2 + N1*(s - N2/2)
  for (int s = 0; s < N2 / 2; ++s) twp[s] = wn(N1 * (s + 1));
  for (int s = N2 / 2; s < N2; ++s) twp[s] = wn(N1 / 2 + N1 * (s - N2 / 2));
  // mel filterbank in (8 rows per octet) form: per octet the longest row length L and the weights W[j][row]
  const int M = d.n_mels, F = plan->n_freqs;
  const int G = (M + 7) / 8;
  std::vector<int> start(G * 8, 0), len(G * 8, 0), ginfo(2 * G, 0);
  for (int m = 0; m < M; ++m) {
    int lo = -1, hi = -1;
    for (int f = 0; f < F; ++f)
      if (plan->h_fb[(size_t)m * F + f] != 0.0f) {
        if (lo < 0) lo = f;
        hi = f;
      }
    start[m] = lo < 0 ? 0 : lo;
    len[m] = lo < 0 ? 0 : hi - lo + 1;
  }
  std::vector<float> wg;
  for (int g = 0; g < G; ++g) {
    int L = 0;
    for (int l = 0; l < 8; ++l) L = std::max(L, len[g * 8 + l]);
    // rows shorter than L read (zero weighted) bins past their end: keep those reads inside the P row
    for (int l = 0; l < 8; ++l)
      if (start[g * 8 + l] + L > F) start[g * 8 + l] = std::max(0, F - L);
    ginfo[2 * g] = L;
    ginfo[2 * g + 1] = (int)wg.size();
    for (int j = 0; j < L; ++j)
      for (int l = 0; l < 8; ++l) {
        const int m = g * 8 + l;
        float w = 0.0f;
        if (m < M) {
          const int f = start[m] + j;  // start may have been shifted left: index the dense filterbank directly
          if (f < F) w = plan->h_fb[(size_t)m * F + f];
        }
        wg.push_back(w);
      }
  }
  fs->groups = G;
  fs->wg_count = (int)wg.size();
  B2A_CUDA(cudaMalloc(&fs->d_win2, sizeof(float2) * NC));
  B2A_CUDA(cudaMalloc(&fs->d_tw1, sizeof(float2) * NC));
  B2A_CUDA(cudaMalloc(&fs->d_twp, sizeof(float2) * NC));
  B2A_CUDA(cudaMalloc(&fs->d_start, sizeof(int) * G * 8));
  B2A_CUDA(cudaMalloc(&fs->d_ginfo, sizeof(int) * 2 * G));
  B2A_CUDA(cudaMalloc(&fs->d_wg, sizeof(float) * std::max<size_t>(wg.size(), 1)));
  B2A_CUDA(cudaMemcpy(fs->d_win2, win2.data(), sizeof(float2) * NC, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_tw1, tw1.data(), sizeof(float2) * NC, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_twp, twp.data(), sizeof(float2) * NC, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_start, start.data(), sizeof(int) * G * 8, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_ginfo, ginfo.data(), sizeof(int) * 2 * G, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_wg, wg.data(), sizeof(float) * wg.size(), cudaMemcpyHostToDevice));
  static const char* const kLogmel[] = {"", "fast_logmel_400x160", "fast_logmel_512x160", "fast_logmel_1024x256", "fast_logmel_800x200",
                                        "fast_logmel_1024x320"};
  static const char* const kStft[] = {"", "fast_stft_400x160", "fast_stft_512x160", "fast_stft_1024x256", "fast_stft_800x200",
                                      "fast_stft_1024x320"};
  plan->kernel_name = d.spec_kind == B2A_SPEC_COMPLEX ? kStft[fs->variant] : kLogmel[fs->variant];
  switch (fs->variant) {
    case 1: fs->spec = fast_match_400(plan, &fs->spec_name); break;
    case 2: fs->spec = fast_match_512(plan, &fs->spec_name); break;
    case 3: fs->spec = fast_match_1024(plan, &fs->spec_name); break;
    case 4: fs->spec = fast_match_800(plan, &fs->spec_name); break;
    default: fs->spec = fast_match_1024h320(plan, &fs->spec_name); break;
  }
  return B2A_OK;
}

void fast_frontend_destroy(b2a_plan* plan) {
  FastState* fs = reinterpret_cast<FastState*>(plan->fast);
  if (!fs) return;
  cudaFree(fs->d_win2);
  cudaFree(fs->d_tw1);
  cudaFree(fs->d_twp);
  cudaFree(fs->d_start);
  cudaFree(fs->d_ginfo);
  cudaFree(fs->d_wg);
  delete fs;
  plan->fast = nullptr;
}

// launch parameters common to the partial step and the single-launch forward
static void fill_fast_params(const b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min, double* feat_sums, FastParams& p) {
  const b2a_frontend_desc& d = plan->fd;
  const FastState* fs = reinterpret_cast<const FastState*>(plan->fast);
  memset(&p, 0, sizeof(p));
  p.audio = a->audio;
  p.clip_stride = a->clip_stride;
  p.valid_length = a->valid_length;
  p.sample_offset = a->sample_offset;
  p.frame_begin = a->frame_begin;
  p.frame_count = a->frame_count;
  p.pad_value = a->pad_value;
  p.batch = a->batch;
  p.geo = make_geometry(a->length, d.n_fft, d.hop, d.center, d.pad_mode);
  p.pad_mode = d.pad_mode;
  p.preemph = d.preemph;
  p.fast_fill_ok = (reinterpret_cast<uintptr_t>(a->audio) % 8 == 0) && (a->clip_stride % 2 == 0) &&
                   (p.geo.pad_left % 2 == 0) && (a->sample_offset % 2 == 0);
  p.spec_kind = d.spec_kind;
  p.spec_eps = d.spec_kind == B2A_SPEC_SQRT_POWER_EPS ? d.spec_eps : 0.0f;
  p.n_mels = d.n_mels;
  const EpilogueConsts ec = epilogue_consts(d);
  p.guard_add = ec.guard_add;
  p.guard_floor = ec.guard_floor;
  p.use_log = ec.use_log;
  p.y_mul = ec.y_mul;
  p.y_add = ec.y_add;
  // all-silent tiles are left to the clamp fix-up (which then writes them once) whenever a clamp follows
  p.skip_floor_tiles = fast_skip_floor_tiles(plan) ? 1 : 0;
  p.out_layout = d.out_layout;
  p.out_dtype = d.out_dtype;
  p.out = reinterpret_cast<float*>(a->out);
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * (d.n_mels > 0 ? d.n_mels : plan->n_freqs);
  p.clip_max = clip_max;
  p.tile_min = tile_min;
  p.feat_sums = feat_sums;
  p.win2 = fs->d_win2;
  p.tw1 = fs->d_tw1;
  p.twp = fs->d_twp;
  p.mel_start = fs->d_start;
  p.mel_ginfo = fs->d_ginfo;
  p.mel_wg = fs->d_wg;
  p.mel_groups = fs->groups;
  p.mel_wg_count = fs->wg_count;
  p.tiles_per_clip = (int)((a->frame_count + 31) / 32);
  p.tile_min_pitch = p.tiles_per_clip;
}

int fast_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min,
                          double* feat_sums, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = reinterpret_cast<FastState*>(plan->fast);
  FastParams p;
  fill_fast_params(plan, a, clip_max, tile_min, feat_sums, p);
  // Trailing frames that see only virtual zero padding are constant rows (fast_const_row0): transform the frames that
  // touch the signal, rounded up to a tile, and fill the rest — here when there is no clamp, in the finalize step otherwise.
  const int64_t const_row0 = fast_const_row0(plan, a);
  if (const_row0 >= 0) {
    p.frame_count = const_row0;
    p.tiles_per_clip = (int)(const_row0 / 32);
  }
#ifdef B2A_PHASE_CLOCKS
  const char* clk_path = getenv("B2A_CLOCKS");
#else
  const char* clk_path = nullptr;
#endif
  if (clk_path) {
    const int maxg = plan->sm_count * 4;
    B2A_CUDA(cudaMalloc(&p.dbg_clk, sizeof(long long) * 8 * maxg));
    B2A_CUDA(cudaMemset(p.dbg_clk, 0, sizeof(long long) * 8 * maxg));
  }
  int rc;
  if (fs->variant == 1) rc = fast_launch_400(plan, fs, p, st);
  else if (fs->variant == 2) rc = fast_launch_512(plan, fs, p, st);
  else if (fs->variant == 3) rc = fast_launch_1024(plan, fs, p, st);
  else if (fs->variant == 4) rc = fast_launch_800(plan, fs, p, st);
  else rc = fast_launch_1024h320(plan, fs, p, st);
  if (rc == B2A_OK && const_row0 >= 0) {
    const ConstRows cr = make_const_rows(d);
    if (d.clamp_kind == B2A_CLAMP_NONE) {
      rc = launch_const_rows(plan, a, const_row0, cr, clip_max, st);
    } else {  // the rows are written by the finalize step (max(c, floor), one pass); their value counts for the clip max
      const_fold_max_kernel<<<(a->batch + 255) / 256, 256, 0, st>>>(clip_max, a->batch, cr);
      B2A_LAUNCHED();
    }
  }
  if (clk_path && rc == B2A_OK) {  // debugging only: synchronises the stream
    const int maxg = plan->sm_count * 4;
    std::vector<long long> h(8 * (size_t)maxg);
    B2A_CUDA(cudaStreamSynchronize(st));
    B2A_CUDA(cudaMemcpy(h.data(), p.dbg_clk, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg_clk);
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < maxg; ++b)
      for (int i = 0; i < 8; ++i) acc[i] += (double)h[8 * b + i];
    if (FILE* f = fopen(clk_path, "a")) {
      const double t = acc[7] > 0 ? acc[7] : 1;
      fprintf(f, "%s spec=%s tiles=%.0f cycles/tile: wait+sync %.0f stage1 %.0f stage2 %.0f melA %.0f melB/mel %.0f\n", plan->kernel_name,
              fs->spec_name ? fs->spec_name : "-", acc[7], acc[0] / t, acc[1] / t, acc[2] / t, acc[3] / t, acc[4] / t);
      fclose(f);
    }
  }
  return rc;
}

// The whole clamping forward in ONE cooperative launch of the TMA kernel: per-tile maxima instead of atomics on an initialised
// clip_max, a grid-wide barrier, the clamp fix-up by the same CTAs.  Applies to the per-clip clamp of the 400/160 generated-mel
// float32 (T, M) instances without constant padding rows and with at most 512 tiles per clip (a CTA reduces a clip's tile
// maxima and fixes its tiles by itself: long files keep the three-launch path, whose fix-up spreads a clip over the grid).
// Returns 1 when it does not apply (or the grid cannot be launched cooperatively): the caller takes the three-launch path.
int fast_frontend_fused(b2a_plan* plan, const b2a_forward_args* a, float* tile_min, float* tile_max, float* clip_max_out, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = reinterpret_cast<FastState*>(plan->fast);
  static const bool off = getenv("B2A_NO_FUSED_FORWARD") != nullptr;  // development toggle, read once
  if (off || !fs || fs->variant != 1 || fs->spec <= 0 || d.clamp_kind != B2A_CLAMP_CLIP_MAX || !fast_skip_floor_tiles(plan) ||
      a->feat_sums != nullptr || fast_const_row0(plan, a) >= 0 || !tile_min || !tile_max)
    return 1;
  const int64_t tpc = (a->frame_count + 31) / 32;
  if (tpc > 512 || a->batch * tpc >= ((int64_t)1 << 31)) return 1;
  FastParams p;
  fill_fast_params(plan, a, clip_max_out, tile_min, nullptr, p);
  p.tile_max = tile_max;
  p.clamp_delta = d.affine_div != 0.0f ? d.clamp_value / d.affine_div : d.clamp_value;  // clamp_fixup_kernel: mx - clamp_value / affine_div
  return fast_launch_400(plan, fs, p, st);
}

}  // namespace b2a
