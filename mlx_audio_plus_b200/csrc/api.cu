// b200audio — C ABI entry points (include/b200audio.h): plan lifetime, dispatch, host-buffer pipeline,
// roofline microbenchmarks.
#include <math.h>

#include <algorithm>
#include <new>

#include "common.cuh"

using namespace b2a;

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
namespace b2a {
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace b2a

namespace {

// Radix plan of the generic shared-memory FFT (csrc/generic.cu run_fft).  Powers of two first — so that the product of the
// radices already applied stays a power of two and the pass index is a mask, not a modulo — split into ceil(e / 5) passes of
// near-equal radix <= 32; then 15, 25, 5, 3 and the remaining primes <= 31.  A trailing 2 / 4 merges with a 3 / 5 into one
// 6 / 10 / 12 / 20 pass (register codelets for all of them: fft_regs.cuh).  1920 = 16 x 8 x 15, 2048 = 16 x 16 x 8,
// 1280 = 16 x 16 x 5, 512 = 32 x 16, 320 = 16 x 20.
int factorize(int n, int* radix, int* nstages) {
  int cnt = 0;
  auto push = [&](int r) -> bool {
    if (cnt >= kMaxStages) return false;
    radix[cnt++] = r;
    return true;
  };
  int e = 0;
  while (n % 2 == 0) { n /= 2; ++e; }
  if (e > 0) {
    const int passes = (e + 4) / 5, base = e / passes, rem = e % passes;
    for (int i = 0; i < passes; ++i)
      if (!push(1 << (base + (i < rem ? 1 : 0)))) return -1;
  }
  const int first_odd = cnt;
  while (n % 15 == 0) { if (!push(15)) return -1; n /= 15; }
  while (n % 25 == 0) { if (!push(25)) return -1; n /= 25; }
  while (n % 5 == 0) { if (!push(5)) return -1; n /= 5; }
  while (n % 3 == 0) { if (!push(3)) return -1; n /= 3; }
  for (int p = 7; p <= kMaxGenericRadix && n > 1; p += 2)
    while (n % p == 0) { if (!push(p)) return -1; n /= p; }
  if (n != 1) return -1;
  // merge the smallest power-of-two pass (the last one) with the first 3 / 5 pass when the product has a codelet
  if (first_odd > 0 && first_odd < cnt) {
    const int a = radix[first_odd - 1], b = radix[first_odd];
    if ((a == 2 || a == 4) && (b == 3 || b == 5)) {
      radix[first_odd - 1] = a * b;
      for (int i = first_odd; i + 1 < cnt; ++i) radix[i] = radix[i + 1];
      --cnt;
      // the merged (non power of two) pass must come after every power-of-two pass: it already does (it was the last of them)
    }
  }
  if (cnt == 0) { radix[cnt++] = 1; }
  *nstages = cnt;
  return 0;
}

int upload(void** dptr, const void* h, size_t bytes) {
  B2A_CUDA(cudaMalloc(dptr, bytes ? bytes : 4));
  if (bytes) B2A_CUDA(cudaMemcpy(*dptr, h, bytes, cudaMemcpyHostToDevice));
  return B2A_OK;
}

int plan_common_init(b2a_plan* p, int n_fft, int hop, const float* h_window, int window_len) {
  p->n_fft = n_fft;
  p->hop = hop;
  p->n_freqs = n_fft / 2 + 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("b200audio: no CUDA device available — this library has no CPU fallback");
    return B2A_ERR_CUDA;
  }
  p->device = dev;
  B2A_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (factorize(n_fft, p->radix, &p->nstages) != 0) {
    set_error("n_fft=%d has a prime factor > %d: unsupported", n_fft, kMaxGenericRadix);
    return B2A_ERR_UNSUPPORTED;
  }
  if (p->nstages == 1 && p->radix[0] == 1) p->nstages = 0;  // n_fft == 1
  std::vector<float2> tw(n_fft);
  for (int k = 0; k < n_fft; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)n_fft;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  int rc = upload((void**)&p->d_twiddle, tw.data(), sizeof(float2) * n_fft);
  if (rc) return rc;
  // per-pass twiddle tables of the generic kernels' Stockham passes (generic.cu: stockham_pass_fixed)
  std::vector<float2> twp(n_fft > 0 ? n_fft : 1, make_float2(1.0f, 0.0f));
  {
    int Ns = 1;
    for (int s = 0; s < p->nstages; ++s) {
      const int R = p->radix[s];
      if (Ns > 1)
        for (int r = 1; r < R; ++r)
          for (int k = 0; k < Ns; ++k) {
            const double a = -2.0 * M_PI * (double)(((int64_t)k * r) % ((int64_t)Ns * R)) / (double)((int64_t)Ns * R);
            twp[(Ns - 1) + (r - 1) * Ns + k] = make_float2((float)cos(a), (float)sin(a));
          }
      Ns *= R;
    }
  }
  rc = upload((void**)&p->d_twiddle_passes, twp.data(), sizeof(float2) * twp.size());
  if (rc) return rc;
  p->h_window.assign(n_fft, 0.0f);  // zero-extended on the right: dsp.py:114-116 / 180-181
  for (int i = 0; i < window_len; ++i) p->h_window[i] = h_window[i];
  return upload((void**)&p->d_window, p->h_window.data(), sizeof(float) * n_fft);
}

int build_mel_csr(b2a_plan* p, const float* fb, int M, int F) {
  std::vector<int> start(M), len(M), off(M);
  std::vector<float> w;
  int maxlen = 0;
  for (int m = 0; m < M; ++m) {
    int lo = -1, hi = -1;
    for (int f = 0; f < F; ++f)
      if (fb[(size_t)m * F + f] != 0.0f) {
        if (lo < 0) lo = f;
        hi = f;
      }
    off[m] = (int)w.size();
    if (lo < 0) {
      start[m] = 0;
      len[m] = 0;
    } else {
      start[m] = lo;
      len[m] = hi - lo + 1;
      for (int f = lo; f <= hi; ++f) w.push_back(fb[(size_t)m * F + f]);
    }
    maxlen = std::max(maxlen, len[m]);
  }
  p->mel.nnz = (int)w.size();
  p->mel.max_len = maxlen;
  int rc;
  if ((rc = upload((void**)&p->mel.d_start, start.data(), sizeof(int) * M))) return rc;
  if ((rc = upload((void**)&p->mel.d_len, len.data(), sizeof(int) * M))) return rc;
  if ((rc = upload((void**)&p->mel.d_off, off.data(), sizeof(int) * M))) return rc;
  if ((rc = upload((void**)&p->mel.d_w, w.data(), sizeof(float) * w.size()))) return rc;
  p->h_fb.assign(fb, fb + (size_t)M * F);
  return upload((void**)&p->d_fb_dense, fb, sizeof(float) * (size_t)M * F);
}

int ensure_ws(b2a_plan* p, size_t bytes) {
  if (p->ws_bytes >= bytes) return B2A_OK;
  if (p->d_ws) cudaFree(p->d_ws);
  p->d_ws = nullptr;
  p->ws_bytes = 0;
  B2A_CUDA(cudaMalloc(&p->d_ws, bytes));
  p->ws_bytes = bytes;
  return B2A_OK;
}

int ensure_tilemin(b2a_plan* p, int slot, size_t count) {
  if (p->tilemin_count[slot] >= count) return B2A_OK;
  if (p->d_tilemin[slot]) cudaFree(p->d_tilemin[slot]);
  p->d_tilemin[slot] = nullptr;
  p->tilemin_count[slot] = 0;
  B2A_CUDA(cudaMalloc(&p->d_tilemin[slot], count * sizeof(float)));
  p->tilemin_count[slot] = count;
  return B2A_OK;
}

size_t stats_bytes(const b2a_plan* p, int batch) {
  const int M = p->fd.n_mels > 0 ? p->fd.n_mels : p->n_freqs;
  size_t b = (size_t)batch * 2 * sizeof(float);
  b = (b + 15) & ~(size_t)15;
  b += (size_t)batch * M * 2 * sizeof(double);
  return b + 64;
}

struct StatPtrs {
  float* clip_max;
  float* tile_min;
  float* tile_max;  // fast family: [tiles] behind tile_min
  int tile_frames;
  double* feat_sums;
};

int stats_tile_frames(const b2a_plan* p, const b2a_forward_args* a) {
  int tf = p->family == KF_FAST ? 32 : generic_tile_frames(p, a);
  return tf > 0 ? tf : 2;
}
size_t tile_count(const b2a_plan* p, const b2a_forward_args* a) {
  const int tf = stats_tile_frames(p, a);
  const size_t tiles = (size_t)a->batch * (size_t)((a->frame_count + tf - 1) / tf);
  return tiles > 0 ? tiles : 1;
}
// floats of per-tile statistics a call needs: minima, and for the fast family maxima behind them (single-launch forward)
size_t tilemin_count(const b2a_plan* p, const b2a_forward_args* a) { return tile_count(p, a) * (p->family == KF_FAST ? 2 : 1); }
// a caller-owned workspace of this size makes the call independent of the plan's own scratch (re-entrant plans)
size_t call_ws_bytes(const b2a_plan* p, const b2a_forward_args* a) {
  return stats_bytes(p, a->batch) + ((tilemin_count(p, a) * sizeof(float) + 15) & ~(size_t)15);
}

int resolve_args(const b2a_plan* p, const b2a_forward_args* in, b2a_forward_args* a) {
  *a = *in;
  if (!in->audio || !in->out || in->batch <= 0) {
    set_error("forward: null buffer or batch <= 0");
    return B2A_ERR_INVALID_ARG;
  }
  int64_t padded, T;
  int rc = b2a_stft_geometry(in->length, p->fd.n_fft, p->fd.hop, p->fd.center, p->fd.pad_mode, &padded, &T);
  if (rc) return rc;
  T -= p->fd.drop_last ? 1 : 0;
  if (a->valid_length <= 0 || a->valid_length > a->length) a->valid_length = a->length;
  if (a->frame_begin < 0) a->frame_begin = 0;
  if (a->frame_count < 0) a->frame_count = T - a->frame_begin;
  if (a->frame_begin + a->frame_count > T) {
    set_error("forward: frame range [%lld,+%lld) exceeds %lld frames", (long long)a->frame_begin,
              (long long)a->frame_count, (long long)T);
    return B2A_ERR_INVALID_ARG;
  }
  if (a->clip_stride == 0) a->clip_stride = a->valid_length - a->sample_offset;
  return B2A_OK;
}

int locate_stats(b2a_plan* p, const b2a_forward_args* a, StatPtrs* s, int slot = 0) {
  const b2a_frontend_desc& d = p->fd;
  const bool need_max = d.clamp_kind != B2A_CLAMP_NONE;
  const bool need_sums = d.norm_kind != B2A_NORM_NONE;
  s->clip_max = s->tile_min = s->tile_max = nullptr;
  s->feat_sums = nullptr;
  s->tile_frames = stats_tile_frames(p, a);
  if (!need_max && !need_sums && !a->clip_max && !a->feat_sums) return B2A_OK;
  char* base;
  const size_t need = stats_bytes(p, a->batch);
  const bool own_ws = a->workspace && a->workspace_bytes >= need;
  if (own_ws) {
    base = (char*)a->workspace;
  } else {
    int rc = ensure_ws(p, need);
    if (rc) return rc;
    base = (char*)p->d_ws;
  }
  s->clip_max = (float*)base;
  if (a->workspace && a->workspace_bytes >= call_ws_bytes(p, a)) {
    s->tile_min = (float*)((char*)a->workspace + need);  // the whole call runs on the caller's scratch
  } else {
    int rc = ensure_tilemin(p, slot, tilemin_count(p, a));
    if (rc) return rc;
    s->tile_min = p->d_tilemin[slot];
  }
  s->tile_max = p->family == KF_FAST ? s->tile_min + tile_count(p, a) : nullptr;
  size_t off = ((size_t)a->batch * 2 * sizeof(float) + 15) & ~(size_t)15;
  s->feat_sums = (double*)(base + off);
  if (a->clip_max) s->clip_max = a->clip_max;  // caller-visible (sharded) statistics
  if (a->feat_sums) s->feat_sums = a->feat_sums;
  if (!need_sums && !a->feat_sums) s->feat_sums = nullptr;
  return B2A_OK;
}

int partial_impl(b2a_plan* p, const b2a_forward_args* a, const StatPtrs& s, cudaStream_t st, bool init) {
  const int M = p->fd.n_mels > 0 ? p->fd.n_mels : p->n_freqs;
  if (init) {
    int rc = init_stats(s.clip_max, s.feat_sums, a->batch, M, st);
    if (rc) return rc;
  }
  if (a->frame_count == 0) return B2A_OK;
  if (p->family == KF_FAST) return fast_frontend_partial(p, a, s.clip_max, s.tile_min, s.feat_sums, st);
  if (p->family == KF_SMALL) return small_stft(p, a, st);
  return generic_frontend_partial(p, a, s.clip_max, s.tile_min, s.feat_sums, st);
}

}  // namespace

extern "C" {

unsigned long long b2a_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b2a_frontend_create(const b2a_frontend_desc* d, const float* h_window, const float* h_fb, b2a_plan** out) {
  if (!d || !h_window || !out) {
    set_error("frontend_create: null argument");
    return B2A_ERR_INVALID_ARG;
  }
  if (d->n_fft <= 0 || d->hop <= 0) {
    set_error("frontend_create: n_fft=%d hop=%d", d->n_fft, d->hop);
    return B2A_ERR_INVALID_ARG;
  }
  if (d->window_len > d->n_fft || d->window_len <= 0) {  // frames * w would not broadcast (dsp.py:141)
    set_error("window of %d taps cannot be broadcast against frames of n_fft=%d", d->window_len, d->n_fft);
    return B2A_ERR_SHAPE;
  }
  if (d->center && d->pad_mode != B2A_PAD_REFLECT && d->pad_mode != B2A_PAD_CONSTANT) {
    set_error("Invalid pad_mode %d", d->pad_mode);
    return B2A_ERR_PAD_MODE;
  }
  if (d->n_mels > 0 && !h_fb) {
    set_error("frontend_create: n_mels=%d but no filterbank", d->n_mels);
    return B2A_ERR_INVALID_ARG;
  }
  if (d->n_mels > 0 && d->spec_kind == B2A_SPEC_COMPLEX) {
    set_error("frontend_create: mel projection of a complex spectrum is undefined");
    return B2A_ERR_INVALID_ARG;
  }
  if (d->frame_len < 0 || d->frame_len > d->n_fft) {
    set_error("frontend_create: frame_len=%d outside [0, n_fft=%d]", d->frame_len, d->n_fft);
    return B2A_ERR_INVALID_ARG;
  }
  if (d->clamp_kind != B2A_CLAMP_NONE && d->norm_kind != B2A_NORM_NONE) {
    set_error("frontend_create: clamp and normalise together are not a reference configuration");
    return B2A_ERR_UNSUPPORTED;
  }
  b2a_plan* p = new (std::nothrow) b2a_plan();
  if (!p) return B2A_ERR_NOMEM;
  p->kind = PLAN_FRONTEND;
  p->fd = *d;
  p->family = KF_GENERIC;
  p->kernel_name = "generic";
  int rc = plan_common_init(p, d->n_fft, d->hop, h_window, d->window_len);
  if (rc == B2A_OK && d->n_mels > 0) rc = build_mel_csr(p, h_fb, d->n_mels, p->n_freqs);
  if (rc == B2A_OK && fast_frontend_supported(p)) {
    rc = fast_frontend_init(p);
    if (rc == B2A_OK) p->family = KF_FAST;
  } else if (rc == B2A_OK && small_stft_supported(p)) {
    p->family = KF_SMALL;
    p->kernel_name = "stft_small";
  }
  if (rc == B2A_OK && d->out_dtype != B2A_DTYPE_F32 &&
      !((d->out_dtype == B2A_DTYPE_F16 || d->out_dtype == B2A_DTYPE_BF16) && p->family == KF_FAST && fast_frontend_out16_ok(p))) {
    set_error("out_dtype %d: 16-bit features are written by the 400/160 generated-mel kernels only ((T, M) layout, named "
              "filterbank, no cross-frame normalisation)", d->out_dtype);
    rc = B2A_ERR_UNSUPPORTED;
  }
  if (rc != B2A_OK) {
    b2a_plan_destroy(p);
    return rc;
  }
  *out = p;
  return B2A_OK;
}

int b2a_deltas(const float* x, float* out, int64_t rows, int64_t cols, int32_t win_length, int32_t edge, void* stream) {
  if (!x || !out || rows < 0 || cols < 0) {
    set_error("deltas: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  if (win_length < 3) {  // dsp.py:456-457
    set_error("win_length should be >= 3, got %d", win_length);
    return B2A_ERR_INVALID_ARG;
  }
  return deltas(x, out, rows, cols, win_length, edge, (cudaStream_t)stream);
}

int b2a_istft_create(const b2a_istft_desc* d, const float* h_window, b2a_plan** out) {
  if (!d || !h_window || !out) {
    set_error("istft_create: null argument");
    return B2A_ERR_INVALID_ARG;
  }
  if (d->n_fft <= 0 || d->hop <= 0 || (d->n_fft & 1)) {
    set_error("istft_create: n_fft=%d hop=%d (n_fft must be even: irfft of F bins yields 2(F-1) samples)", d->n_fft, d->hop);
    return B2A_ERR_INVALID_ARG;
  }
  if (d->window_len > d->n_fft || d->window_len <= 0) {  // frames_time * w would not broadcast (dsp.py:197)
    set_error("window of %d taps cannot be broadcast against irfft frames of %d samples", d->window_len, d->n_fft);
    return B2A_ERR_SHAPE;
  }
  b2a_plan* p = new (std::nothrow) b2a_plan();
  if (!p) return B2A_ERR_NOMEM;
  p->kind = PLAN_ISTFT;
  p->id = *d;
  p->family = KF_GENERIC;
  p->kernel_name = "generic_istft";
  int rc = plan_common_init(p, d->n_fft, d->hop, h_window, d->window_len);
  if (rc == B2A_OK && small_istft_supported(p)) {
    p->family = KF_SMALL;
    p->kernel_name = "istft_small";
  } else if (rc == B2A_OK && fast_istft_supported(p)) {
    rc = fast_istft_init(p);
    if (rc == B2A_OK) p->family = KF_FAST;
  }
  if (rc != B2A_OK) {
    b2a_plan_destroy(p);
    return rc;
  }
  *out = p;
  return B2A_OK;
}

int b2a_plan_destroy(b2a_plan* p) {
  if (!p) return B2A_OK;
  if (p->fast && p->kind == PLAN_ISTFT) fast_istft_destroy(p);
  if (p->fast) fast_frontend_destroy(p);
  cudaFree(p->d_twiddle);
  cudaFree(p->d_twiddle_passes);
  cudaFree(p->d_window);
  cudaFree(p->mel.d_start);
  cudaFree(p->mel.d_len);
  cudaFree(p->mel.d_off);
  cudaFree(p->mel.d_w);
  cudaFree(p->d_fb_dense);
  cudaFree(p->d_ws);
  for (int i = 0; i < 2; ++i) {
    cudaFree(p->d_tilemin[i]);
    cudaFree(p->d_stage_in[i]);
    cudaFree(p->d_stage_out[i]);
    if (p->host_streams[i]) cudaStreamDestroy(p->host_streams[i]);
  }
  cudaGetLastError();
  delete p;
  return B2A_OK;
}

const char* b2a_plan_kernel_name(const b2a_plan* p) { return p ? p->kernel_name : ""; }

int b2a_frontend_out_frames(const b2a_plan* p, int64_t length, int64_t* frames) {
  if (!p || p->kind != PLAN_FRONTEND) return B2A_ERR_INVALID_ARG;
  int64_t padded, T;
  int rc = b2a_stft_geometry(length, p->fd.n_fft, p->fd.hop, p->fd.center, p->fd.pad_mode, &padded, &T);
  if (rc) return rc;
  if (frames) *frames = T - (p->fd.drop_last ? 1 : 0);
  return B2A_OK;
}

size_t b2a_frontend_workspace_bytes(const b2a_plan* p, int32_t batch) {
  if (!p || p->kind != PLAN_FRONTEND) return 0;
  return stats_bytes(p, batch);
}

size_t b2a_frontend_call_workspace_bytes(const b2a_plan* p, const b2a_forward_args* in) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return 0;
  b2a_forward_args a;
  if (resolve_args(p, in, &a)) return 0;
  return call_ws_bytes(p, &a);
}

static int reject_pcm16(const b2a_forward_args* in) {
  if (in->audio_kind == B2A_PCM_F32) return B2A_OK;
  set_error("audio_kind %d: int16 PCM input is taken by b2a_frontend_forward_host only (device entries read float32)", in->audio_kind);
  return B2A_ERR_UNSUPPORTED;
}

int b2a_frontend_partial(b2a_plan* p, const b2a_forward_args* in, void* stream) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return B2A_ERR_INVALID_ARG;
  if (int prc = reject_pcm16(in)) return prc;
  b2a_forward_args a;
  int rc = resolve_args(p, in, &a);
  if (rc) return rc;
  StatPtrs s;
  if ((rc = locate_stats(p, &a, &s))) return rc;
  return partial_impl(p, &a, s, (cudaStream_t)stream, true);
}

int b2a_frontend_finalize(b2a_plan* p, const b2a_forward_args* in, int64_t global_frames, void* stream) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return B2A_ERR_INVALID_ARG;
  b2a_forward_args a;
  int rc = resolve_args(p, in, &a);
  if (rc) return rc;
  StatPtrs s;
  if ((rc = locate_stats(p, &a, &s))) return rc;
  if (a.frame_count == 0) return B2A_OK;
  return frontend_finalize(p, &a, global_frames, s.clip_max, s.tile_min, s.tile_frames, s.feat_sums, (cudaStream_t)stream);
}

int b2a_frontend_forward(b2a_plan* p, const b2a_forward_args* in, void* stream) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return B2A_ERR_INVALID_ARG;
  if (int prc = reject_pcm16(in)) return prc;
  b2a_forward_args a;
  int rc = resolve_args(p, in, &a);
  if (rc) return rc;
  StatPtrs s;
  if ((rc = locate_stats(p, &a, &s))) return rc;
  if (p->family == KF_FAST && a.frame_count > 0 && s.tile_max) {
    // per-clip clamp on the TMA kernel: statistics, grid-wide barrier and fix-up in ONE cooperative launch (1 = not applicable)
    rc = fast_frontend_fused(p, &a, s.tile_min, s.tile_max, a.clip_max ? s.clip_max : nullptr, (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  if ((rc = partial_impl(p, &a, s, (cudaStream_t)stream, true))) return rc;
  if (a.frame_count == 0) return B2A_OK;
  return frontend_finalize(p, &a, a.frame_count, s.clip_max, s.tile_min, s.tile_frames, s.feat_sums, (cudaStream_t)stream);
}

int b2a_frontend_dump_frames(b2a_plan* p, const b2a_forward_args* in, int apply_window, void* stream) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return B2A_ERR_INVALID_ARG;
  if (int prc = reject_pcm16(in)) return prc;
  b2a_forward_args a;
  int rc = resolve_args(p, in, &a);
  if (rc) return rc;
  // dump ignores drop_last bookkeeping beyond what resolve_args applied
  return dump_frames(p, &a, apply_window, (cudaStream_t)stream);
}

// int16 PCM -> float32 / 32768 (audio_io.py:258-262; exact), 8 samples per thread; n8 = groups of eight, then the tail
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t n8 = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 v = reinterpret_cast<const int4*>(in)[i];
    const int w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = (float)(short)(w[k] & 0xffff) * (1.0f / 32768.0f);
      f[2 * k + 1] = (float)(short)(w[k] >> 16) * (1.0f / 32768.0f);
    }
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 8 * n8)) out[8 * n8 + threadIdx.x] = (float)in[8 * n8 + threadIdx.x] * (1.0f / 32768.0f);
}

// ---- host-buffer pipeline: chunks of clips, two streams, H2D / compute / D2H overlapped -----------------
static int ensure_stage(b2a_plan* p, size_t in_bytes, size_t out_bytes) {
  for (int i = 0; i < 2; ++i) {
    if (!p->host_streams[i]) B2A_CUDA(cudaStreamCreateWithFlags(&p->host_streams[i], cudaStreamNonBlocking));
    if (p->stage_in_bytes < in_bytes) {
      cudaFree(p->d_stage_in[i]);
      p->d_stage_in[i] = nullptr;
      B2A_CUDA(cudaMalloc(&p->d_stage_in[i], in_bytes));
    }
    if (p->stage_out_bytes < out_bytes) {
      cudaFree(p->d_stage_out[i]);
      p->d_stage_out[i] = nullptr;
      B2A_CUDA(cudaMalloc(&p->d_stage_out[i], out_bytes));
    }
  }
  p->stage_in_bytes = std::max(p->stage_in_bytes, in_bytes);
  p->stage_out_bytes = std::max(p->stage_out_bytes, out_bytes);
  return B2A_OK;
}

int b2a_frontend_forward_host(b2a_plan* p, const b2a_forward_args* in) {
  if (!p || p->kind != PLAN_FRONTEND || !in) return B2A_ERR_INVALID_ARG;
  b2a_forward_args a;
  int rc = resolve_args(p, in, &a);
  if (rc) return rc;
  const b2a_frontend_desc& d = p->fd;
  const int M = d.n_mels > 0 ? d.n_mels : p->n_freqs;
  const size_t out_elem = (d.n_mels == 0 && d.spec_kind == B2A_SPEC_COMPLEX) ? 8 : (d.out_dtype != B2A_DTYPE_F32 ? 2 : 4);
  const int64_t in_per_clip = a.valid_length - a.sample_offset;  // samples physically present per clip
  const int64_t out_per_clip = a.frame_count * M;
  const int64_t out_stride = a.out_clip_stride ? a.out_clip_stride : out_per_clip;
  // chunk so that each staging buffer stays <= ~192 MB, at least 8 chunks for overlap on big batches
  int chunk = a.batch;
  const size_t cap = (size_t)192 << 20;
  const size_t per_clip_bytes = std::max((size_t)in_per_clip * 4, (size_t)out_per_clip * out_elem);
  if (d.clamp_kind != B2A_CLAMP_BATCH_MAX) {
    if ((size_t)chunk * per_clip_bytes > cap) chunk = (int)std::max<size_t>(1, cap / per_clip_bytes);
    if (a.batch >= 16 && chunk > (a.batch + 7) / 8) chunk = (a.batch + 7) / 8;
  }
  const bool pcm16 = a.audio_kind == B2A_PCM_I16;
  if (a.audio_kind != B2A_PCM_F32 && !pcm16) {
    set_error("forward_host: audio_kind %d", a.audio_kind);
    return B2A_ERR_INVALID_ARG;
  }
  // int16 input: the PCM lands behind the float32 area of the staging buffer (16-byte aligned) and is converted in place
  const size_t f32_bytes = ((size_t)chunk * in_per_clip * 4 + 15) & ~(size_t)15;
  if ((rc = ensure_stage(p, f32_bytes + (pcm16 ? (size_t)chunk * in_per_clip * 2 : 0), (size_t)chunk * out_per_clip * out_elem))) return rc;
  // per-stream statistic scratch
  const size_t sb = stats_bytes(p, chunk);
  if ((rc = ensure_ws(p, 2 * sb))) return rc;
  // the chunk loop as a lambda: whatever it returns, BOTH streams are drained before this call returns — asynchronous copies
  // must never still be writing into the caller's buffers after an error return
  auto run_chunks = [&]() -> int {
  int idx = 0;
  for (int c0 = 0; c0 < a.batch; c0 += chunk, ++idx) {
    const int nb = std::min(chunk, a.batch - c0);
    const int s = idx & 1;
    cudaStream_t st = p->host_streams[s];
    if (pcm16) {
      const int16_t* hsrc = reinterpret_cast<const int16_t*>(a.audio) + (int64_t)c0 * a.clip_stride;
      int16_t* d16 = reinterpret_cast<int16_t*>((char*)p->d_stage_in[s] + f32_bytes);
      if (a.clip_stride == in_per_clip) {
        B2A_CUDA(cudaMemcpyAsync(d16, hsrc, (size_t)nb * in_per_clip * 2, cudaMemcpyHostToDevice, st));
      } else {
        B2A_CUDA(cudaMemcpy2DAsync(d16, (size_t)in_per_clip * 2, hsrc, (size_t)a.clip_stride * 2, (size_t)in_per_clip * 2, nb,
                                   cudaMemcpyHostToDevice, st));
      }
      const int64_t n = (int64_t)nb * in_per_clip;
      pcm16_to_f32_kernel<<<(unsigned)std::min<int64_t>((n / 8 + 255) / 256 + 1, 4096), 256, 0, st>>>(d16, (float*)p->d_stage_in[s], n);
      B2A_LAUNCHED();
    } else {
    const float* hsrc = a.audio + (int64_t)c0 * a.clip_stride;
    if (a.clip_stride == in_per_clip) {
      B2A_CUDA(cudaMemcpyAsync(p->d_stage_in[s], hsrc, (size_t)nb * in_per_clip * 4, cudaMemcpyHostToDevice, st));
    } else {
      B2A_CUDA(cudaMemcpy2DAsync(p->d_stage_in[s], (size_t)in_per_clip * 4, hsrc, (size_t)a.clip_stride * 4,
                                 (size_t)in_per_clip * 4, nb, cudaMemcpyHostToDevice, st));
    }
    }
    b2a_forward_args c = a;
    c.audio_kind = B2A_PCM_F32;
    c.audio = (const float*)p->d_stage_in[s];
    c.clip_stride = in_per_clip;
    c.batch = nb;
    c.out = p->d_stage_out[s];
    c.out_clip_stride = out_per_clip;
    c.clip_max = nullptr;
    c.feat_sums = nullptr;
    c.workspace = (char*)p->d_ws + s * sb;
    c.workspace_bytes = sb;
    StatPtrs sp;
    if ((rc = locate_stats(p, &c, &sp, s))) return rc;
    if ((rc = partial_impl(p, &c, sp, st, true))) return rc;
    if (c.frame_count > 0 &&
        (rc = frontend_finalize(p, &c, c.frame_count, sp.clip_max, sp.tile_min, sp.tile_frames, sp.feat_sums, st)))
      return rc;
    char* hdst = (char*)a.out + (size_t)c0 * out_stride * out_elem;
    if (out_stride == out_per_clip) {
      B2A_CUDA(cudaMemcpyAsync(hdst, p->d_stage_out[s], (size_t)nb * out_per_clip * out_elem, cudaMemcpyDeviceToHost, st));
    } else {
      B2A_CUDA(cudaMemcpy2DAsync(hdst, (size_t)out_stride * out_elem, p->d_stage_out[s], (size_t)out_per_clip * out_elem,
                                 (size_t)out_per_clip * out_elem, nb, cudaMemcpyDeviceToHost, st));
    }
  }
  return B2A_OK;
  };
  rc = run_chunks();
  const cudaError_t e0 = cudaStreamSynchronize(p->host_streams[0]), e1 = cudaStreamSynchronize(p->host_streams[1]);
  if (rc) return rc;
  B2A_CUDA(e0);
  B2A_CUDA(e1);
  return B2A_OK;
}

// ---- inverse -------------------------------------------------------------------------------------------
int b2a_istft_out_len(const b2a_plan* p, int64_t num_frames, int64_t length, int64_t* out_len) {
  if (!p || p->kind != PLAN_ISTFT) return B2A_ERR_INVALID_ARG;
  const b2a_istft_desc& d = p->id;
  int64_t ola, start, len;
  int rc = b2a_istft_geometry(num_frames, d.n_fft, d.hop, d.center, d.trim_tail ? length : -1, &ola, &start, &len);
  if (rc) return rc;
  if (!d.trim_tail) {
    start = d.center ? d.n_fft / 2 : 0;
    len = std::max<int64_t>(0, ola - start);
    if (length >= 0 && length < len) len = length;
  }
  if (out_len) *out_len = len;
  return B2A_OK;
}

// argument checks shared by the device and the host entry point
static int check_inverse_args(const b2a_plan* p, const b2a_inverse_args* a, const char* who) {
  if (!p || p->kind != PLAN_ISTFT || !a || !a->spec || !a->out || a->batch <= 0 || a->num_frames <= 0) {
    set_error("%s: invalid argument", who);
    return B2A_ERR_INVALID_ARG;
  }
  if (p->id.input_form == B2A_ISTFT_INPUT_POLAR && !a->spec_imag) {
    set_error("%s: the polar input form needs the phase plane in spec_imag", who);
    return B2A_ERR_INVALID_ARG;
  }
  return B2A_OK;
}

int b2a_istft_inverse(b2a_plan* p, const b2a_inverse_args* a, void* stream) {
  if (int vrc = check_inverse_args(p, a, "istft_inverse")) return vrc;
  if (p->family == KF_SMALL) return small_istft(p, a, (cudaStream_t)stream);
  if (p->family == KF_FAST) return fast_istft(p, a, (cudaStream_t)stream);
  return generic_istft(p, a, (cudaStream_t)stream);
}

int b2a_istft_inverse_host(b2a_plan* p, const b2a_inverse_args* in) {
  if (int vrc = check_inverse_args(p, in, "istft_inverse_host")) return vrc;
  const int F = p->n_freqs;
  const int64_t T = in->num_frames;
  int64_t out_len;
  int rc = b2a_istft_out_len(p, T, in->length, &out_len);
  if (rc) return rc;
  const bool planar = in->spec_imag != nullptr;
  const size_t elem = planar ? 4 : 8;
  const int64_t in_per_clip = (int64_t)F * T;
  const int64_t in_stride = in->clip_stride ? in->clip_stride : in_per_clip;
  const int64_t out_stride = in->out_clip_stride ? in->out_clip_stride : out_len;
  int chunk = in->batch;
  const size_t cap = (size_t)192 << 20;
  const size_t per_clip_bytes = std::max((size_t)in_per_clip * 8, (size_t)out_len * 4);
  if ((size_t)chunk * per_clip_bytes > cap) chunk = (int)std::max<size_t>(1, cap / per_clip_bytes);
  if (in->batch >= 16 && chunk > (in->batch + 7) / 8) chunk = (in->batch + 7) / 8;
  if ((rc = ensure_stage(p, (size_t)chunk * in_per_clip * 8, (size_t)chunk * std::max<int64_t>(out_len, 1) * 4))) return rc;
  auto run_chunks = [&]() -> int {  // both streams are drained below whatever this returns
  int idx = 0;
  for (int c0 = 0; c0 < in->batch; c0 += chunk, ++idx) {
    const int nb = std::min(chunk, in->batch - c0);
    const int s = idx & 1;
    cudaStream_t st = p->host_streams[s];
    char* dre = (char*)p->d_stage_in[s];
    char* dim = dre + (size_t)chunk * in_per_clip * 4;
    const char* hre = (const char*)in->spec + (size_t)c0 * in_stride * elem;
    B2A_CUDA(cudaMemcpy2DAsync(dre, (size_t)in_per_clip * elem, hre, (size_t)in_stride * elem,
                               (size_t)in_per_clip * elem, nb, cudaMemcpyHostToDevice, st));
    if (planar) {
      const char* him = (const char*)in->spec_imag + (size_t)c0 * in_stride * elem;
      B2A_CUDA(cudaMemcpy2DAsync(dim, (size_t)in_per_clip * 4, him, (size_t)in_stride * 4, (size_t)in_per_clip * 4, nb,
                                 cudaMemcpyHostToDevice, st));
    }
    b2a_inverse_args c = *in;
    c.spec = dre;
    c.spec_imag = planar ? dim : nullptr;
    c.clip_stride = in_per_clip;
    c.batch = nb;
    c.out = (float*)p->d_stage_out[s];
    c.out_clip_stride = out_len;
    if ((rc = (p->family == KF_SMALL ? small_istft(p, &c, st)
                                     : (p->family == KF_FAST ? fast_istft(p, &c, st) : generic_istft(p, &c, st)))))
      return rc;
    if (out_len > 0)
      B2A_CUDA(cudaMemcpy2DAsync((char*)in->out + (size_t)c0 * out_stride * 4, (size_t)out_stride * 4, p->d_stage_out[s],
                                 (size_t)out_len * 4, (size_t)out_len * 4, nb, cudaMemcpyDeviceToHost, st));
  }
  return B2A_OK;
  };
  rc = run_chunks();
  const cudaError_t e0 = cudaStreamSynchronize(p->host_streams[0]), e1 = cudaStreamSynchronize(p->host_streams[1]);
  if (rc) return rc;
  B2A_CUDA(e0);
  B2A_CUDA(e1);
  return B2A_OK;
}

}  // extern "C"

// ---- roofline microbenchmarks -------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace

extern "C" {

int b2a_measure_fp32_tflops(double* tflops, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev, sms;
  B2A_CUDA(cudaGetDevice(&dev));
  B2A_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, iters = 4096;
  float* d;
  B2A_CUDA(cudaMalloc(&d, (size_t)blocks * 256 * 4));
  cudaEvent_t e0, e1;
  B2A_CUDA(cudaEventCreate(&e0));
  B2A_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    B2A_CUDA(cudaEventRecord(e0, st));
    ffma_peak_kernel<<<blocks, 256, 0, st>>>(d, iters, 1.0000001f, 1e-9f);
    B2A_CUDA(cudaEventRecord(e1, st));
    B2A_CUDA(cudaEventSynchronize(e1));
    float ms;
    B2A_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * 256 * iters * 64 * 2;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (tflops) *tflops = best;
  return B2A_OK;
}

int b2a_measure_copy_gbs(double* gbs, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev, sms;
  B2A_CUDA(cudaGetDevice(&dev));
  B2A_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t bytes = (size_t)2 << 30;
  float4 *a, *b;
  B2A_CUDA(cudaMalloc(&a, bytes));
  B2A_CUDA(cudaMalloc(&b, bytes));
  B2A_CUDA(cudaMemsetAsync(a, 1, bytes, st));
  cudaEvent_t e0, e1;
  B2A_CUDA(cudaEventCreate(&e0));
  B2A_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    B2A_CUDA(cudaEventRecord(e0, st));
    copy_kernel<<<sms * 16, 256, 0, st>>>(a, b, bytes / 16);
    B2A_CUDA(cudaEventRecord(e1, st));
    B2A_CUDA(cudaEventSynchronize(e1));
    float ms;
    B2A_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  if (gbs) *gbs = best;
  return B2A_OK;
}

}  // extern "C"
