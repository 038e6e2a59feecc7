// b200audio — small-n kernels (n_fft in {16, 20}: Kokoro iSTFTNet n_fft=20/hop=5, HiFT-style n_fft=16/hop=4).
//
// ONE THREAD PER FRAME, whole transform in registers (fft_regs.cuh), no shared memory:
//   istft_small_kernel<N>  lanes = consecutive frames, so the (F, T) spectrum is read with fully coalesced
//                          8-byte loads; irfft via the half-size complex DFT (Hermitian pre-twiddle); the
//                          windowed frame stays in registers and the overlap-add of the 4 frames covering an
//                          output hop is done with warp shuffles in ascending frame order (the order of the
//                          reference's sequential scatter-add, dsp.py:203-204); each warp re-computes a 3-frame
//                          halo instead of exchanging through memory (29 useful frames of 32).  HBM traffic:
//                          88 B in (+10 % halo) and 20 B out per frame for n_fft=20 — the kernel is HBM-bound.
//   stft_small_kernel<N>   the forward counterpart for dsp.stft with complex output (Kokoro transform).
#include <algorithm>

#include "common.cuh"
#include "fft_regs.cuh"

namespace b2a {
namespace {

using regs::Dft;
using regs::static_for;

struct SmallInvParams {
  const float2* spec;    // interleaved (B, F, T) or nullptr
  const float* spec_re;  // planar
  const float* spec_im;
  int64_t clip_stride, T;
  int hop;
  int norm_sq, div_clamp;
  int64_t out_start, out_len, out_clip_stride;
  float* out;
  float w[32];   // synthesis window (zero extended)
  float wn[32];  // window / n_fft: the inverse transform's 1 / N folded into the window product
  int warps_per_clip;
  PolarSpec polar;
  float div_eps;
  float rden[8];  // 1 / (full-overlap envelope of output sample j of a hop, summed in ascending frame order)
  int rden_ok;    // every entry of that envelope lies above the division guard
};

// sin / cos for |x| <= 1 without range reduction (near-minimax fits, <= 0.6 ulp of 1.0): the phases Kokoro's generator hands to
// MLXSTFT.inverse are sin(.) of a network output (istftnet.py), i.e. always inside [-1, 1] — 12 instructions per bin instead of 22
__device__ __forceinline__ void sincos_unit(float x, float* sn, float* cs) {
  const float r2 = x * x;
  float s = fmaf(r2, 0x1.27dbb8p-19f, -0x1.9e3becp-13f);
  s = fmaf(s, r2, 0x1.110d8ap-7f);
  s = fmaf(s, r2, -0x1.55554ep-3f);
  *sn = fmaf(s * r2, x, x);
  float c = fmaf(r2, 0x1.783fa2p-22f, 0x1.88ffdp-16f);
  c = fmaf(c, r2, -0x1.6bd36p-10f);
  c = fmaf(c, r2, 0x1.5554c6p-5f);
  c = fmaf(c, r2, -0x1.fffffep-2f);
  *cs = fmaf(c, r2, 1.0f);
}

__device__ __forceinline__ float num_div(float a, float b) { return a / b; }  // IEEE division, as the general path below

template <int N, bool POLAR>
__global__ void __launch_bounds__(256) istft_small_kernel(const SmallInvParams p) {
  constexpr int NC = N / 2, F = NC + 1, HOP = N / 4;
  const int lane = threadIdx.x & 31;
  const int wclip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // warp index within the clip
  if (wclip >= p.warps_per_clip) return;
  const int clip_i = blockIdx.y;
  const int64_t t = (int64_t)wclip * 29 - 3 + lane;  // frame slot; slots < 0 or >= T contribute nothing
  const bool live = t >= 0 && t < p.T;

  // ---- load the frame's spectrum: lanes are consecutive frames -> coalesced; one running pointer per plane ------
  float2 X[F];
  const int64_t base = (int64_t)clip_i * p.clip_stride + (live ? t : 0);
  if (p.spec) {
    const float2* q = p.spec + base;
#pragma unroll
    for (int k = 0; k < F; ++k, q += p.T) X[k] = live ? __ldg(q) : make_float2(0.f, 0.f);
  } else {
    const float *qr = p.spec_re + base, *qi = p.spec_im + base;
#pragma unroll
    for (int k = 0; k < F; ++k, qr += p.T, qi += p.T) X[k] = live ? make_float2(__ldg(qr), __ldg(qi)) : make_float2(0.f, 0.f);
    // magnitude / phase -> complex in a SECOND pass: with the conversion (and its range-reduction branch) inside the load
    // loop the 2 * F loads were issued one bin at a time, each waiting out an HBM round trip (C4: 1.32 ms, see DESIGN)
    if (POLAR && live) {
      // the common form (no clipping, plain magnitudes, |phase| <= 65536: Kokoro) is straight-line code — no per-bin
      // flag tests, no per-bin slow-path branch — so the F independent sin / cos chains interleave
      float big = 0.0f;
#pragma unroll
      for (int k = 0; k < F; ++k) big = fmaxf(big, fabsf(X[k].y));
      const bool plain = !p.polar.log_mag && !(p.polar.clip_max > 0.0f) && !p.polar.clip_min_zero;
      if (plain && big <= 1.0f) {
#pragma unroll
        for (int k = 0; k < F; ++k) {
          float sn, cs;
          sincos_unit(X[k].y, &sn, &cs);
          X[k] = make_float2(X[k].x * cs, X[k].x * sn);
        }
      } else if (plain && big <= 65536.0f) {
#pragma unroll
        for (int k = 0; k < F; ++k) {
          float sn, cs;
          sincos_cw_reduced(X[k].y, &sn, &cs);
          X[k] = make_float2(X[k].x * cs, X[k].x * sn);
        }
      } else {
#pragma unroll
        for (int k = 0; k < F; ++k) X[k] = polar_to_complex(p.polar, X[k]);
      }
    }
  }
  X[0].y = 0.f;   // irfft ignores Im(DC) and Im(Nyquist)
  X[NC].y = 0.f;

  // ---- Hermitian pre-twiddle: Z[k] = E[k] + i O[k], E = (X[k] + conj X[Nc-k])/2, O = (X[k] - conj X[Nc-k]) W_N^-k / 2
  // stored conjugated so that a FORWARD DFT yields conj(Nc * z)
  float2 z[NC];
  static_for<0, NC>([&](auto K_) {
    constexpr int k = decltype(K_)::value;
    const float2 a = X[k], b = X[NC - k];
    constexpr regs::cplx_d w = regs::unit_root(k, N);  // W_N^-k = exp(+2 pi i k / N)
    constexpr float wr = (float)w.re, wi = (float)w.im;
    // packed f32x2 form (FFMA2 / FMUL2, as the fused kernels): 2E conjugated, X[k] - conj X[Nc-k], 2O, then
    // conj(Z) = (ex - oy, -(ey + ox)) with Z = E + iO
    const float2 ec = regs::pfma(a, make_float2(1.0f, -1.0f), b);   // (ex, -ey)
    const float2 d = regs::pfma(b, make_float2(-1.0f, 1.0f), a);    // (dx, dy)
    const float2 o = regs::cmul(d, make_float2(wr, wi));            // (ox, oy)
    z[k] = regs::pfma(regs::pswap(o), make_float2(-1.0f, -1.0f), ec);
  });
  Dft<NC>::run(z);
  // x[2m] = Re z[m], x[2m+1] = Im z[m], z = conj(DFT(conj Z)) / Nc, and the /2 of E,O  => scale 1/N
  float y[N];
#pragma unroll
  for (int m = 0; m < NC; ++m) {  // p.wn = window / N (exact for n_fft 16; one rounding instead of two for n_fft 20)
    y[2 * m] = z[m].x * p.wn[2 * m];
    y[2 * m + 1] = -z[m].y * p.wn[2 * m + 1];
  }

  // ---- overlap-add by shuffles, ascending frame order: t-3, t-2, t-1, t ------------------------------------
  // interior lanes (all four contributing frames exist, the whole hop inside the output range): sums and the envelope of
  // the four taps in the same order as below, without the per-frame existence tests (88 ISETP per warp on the source page)
  const bool inner = t - 3 >= 0 && t < p.T && t * HOP - p.out_start >= 0 && t * HOP + HOP - p.out_start <= p.out_len;
  if (__all_sync(0xffffffffu, inner || lane < 3)) {
    // the warp's 29 hops are ONE contiguous run of the output: results are parked in a per-warp staging row (pitch odd:
    // conflict free) and leave as coalesced 128-byte stores — a thread storing its own hop writes HOP floats at a stride of
    // HOP floats, i.e. every store instruction of the warp touches all ~19 sectors of the run
    constexpr int SP = HOP | 1;
    __shared__ float s_out[8][32 * SP];
    float* const so = s_out[threadIdx.x >> 5];
    const bool rd = p.rden_ok != 0;
#pragma unroll
    for (int j = 0; j < HOP; ++j) {
      const float y3 = __shfl_up_sync(0xffffffffu, y[j + 3 * HOP], 3);
      const float y2 = __shfl_up_sync(0xffffffffu, y[j + 2 * HOP], 2);
      const float y1 = __shfl_up_sync(0xffffffffu, y[j + HOP], 1);
      const float n_ = ((0.f + y3) + y2) + y1 + y[j];
      float r;
      if (rd) {
        // all four frames exist: the envelope is a constant of j; its reciprocal (rounded once from double on the host)
        // replaces the IEEE division (<= 1 ulp from the quotient), as in the 1024 / 256 kernel
        r = n_ * p.rden[j];
      } else {
        const float w3 = p.w[j + 3 * HOP], w2 = p.w[j + 2 * HOP], w1 = p.w[j + HOP], w0 = p.w[j];
        const float d_ = p.norm_sq ? (((0.f + w3 * w3) + w2 * w2) + w1 * w1) + w0 * w0 : (((0.f + w3) + w2) + w1) + w0;
        if (p.div_clamp) r = num_div(n_, fmaxf(d_, p.div_eps));
        else r = d_ > p.div_eps ? num_div(n_, d_) : n_;
      }
      so[lane * SP + j] = r;
    }
    __syncwarp();
    float* const o = p.out + (int64_t)clip_i * p.out_clip_stride + ((int64_t)wclip * 29 * HOP - p.out_start);  // hop of lane 3
#pragma unroll
    for (int i0 = 0; i0 < 29 * HOP; i0 += 32) {
      const int i = i0 + lane;
      if (i < 29 * HOP) o[i] = so[(3 + i / HOP) * SP + i % HOP];
    }
    return;
  }
  float num[HOP], den[HOP];
#pragma unroll
  for (int j = 0; j < HOP; ++j) {
    const float y3 = __shfl_up_sync(0xffffffffu, y[j + 3 * HOP], 3);
    const float y2 = __shfl_up_sync(0xffffffffu, y[j + 2 * HOP], 2);
    const float y1 = __shfl_up_sync(0xffffffffu, y[j + HOP], 1);
    float n_ = 0.f, d_ = 0.f;
    // frames t-q exist iff 0 <= t-q < T
    if (t - 3 >= 0 && t - 3 < p.T) { n_ += y3; const float w = p.w[j + 3 * HOP]; d_ += p.norm_sq ? w * w : w; }
    if (t - 2 >= 0 && t - 2 < p.T) { n_ += y2; const float w = p.w[j + 2 * HOP]; d_ += p.norm_sq ? w * w : w; }
    if (t - 1 >= 0 && t - 1 < p.T) { n_ += y1; const float w = p.w[j + HOP]; d_ += p.norm_sq ? w * w : w; }
    if (live) { n_ += y[j]; const float w = p.w[j]; d_ += p.norm_sq ? w * w : w; }
    num[j] = n_;
    den[j] = d_;
  }
  if (lane < 3 || t < 0 || t >= p.T + 3) return;  // halo lanes / slots beyond the OLA tail
  float* o = p.out + (int64_t)clip_i * p.out_clip_stride;
#pragma unroll
  for (int j = 0; j < HOP; ++j) {
    const int64_t n = t * HOP + j;
    const int64_t jo = n - p.out_start;
    if (jo >= 0 && jo < p.out_len) {
      float r;
      if (p.div_clamp) r = num[j] / fmaxf(den[j], p.div_eps);
      else r = den[j] > p.div_eps ? num[j] / den[j] : num[j];
      o[jo] = r;
    }
  }
}

struct SmallFwdParams {
  const float* audio;
  int64_t clip_stride, valid_length, sample_offset, frame_begin, frame_count;
  float pad_value;
  Geometry geo;
  int hop, pad_mode;
  float preemph;
  float2* out;
  int64_t out_clip_stride;
  float w[32];
};

template <int N>
__global__ void __launch_bounds__(256) stft_small_kernel(const SmallFwdParams p) {
  constexpr int NC = N / 2, F = NC + 1;
  const int64_t lt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lt - (threadIdx.x & 31) >= p.frame_count) return;  // whole warp past the end (a ragged warp keeps all its lanes: see the write-out)
  const bool dead = lt >= p.frame_count;
  const int clip_i = blockIdx.y;
  const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
  const int64_t q0 = (p.frame_begin + lt) * p.hop;
  const int64_t s0 = q0 - p.geo.pad_left;
  float x[N];
  if (dead) {
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = 0.0f;
  } else if (s0 >= 1 && s0 + N <= p.valid_length && s0 >= p.sample_offset + 1 && p.preemph == 0.0f) {
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = __ldg(clip + (s0 - p.sample_offset) + k);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int64_t s = source_index(p.geo, p.pad_mode, q0 + k);
      float v = 0.f;
      if (s >= 0) {
        v = s < p.valid_length ? __ldg(clip + (s - p.sample_offset)) : p.pad_value;
        if (p.preemph != 0.0f && s > 0) {
          const float xm = (s - 1) < p.valid_length ? __ldg(clip + (s - 1 - p.sample_offset)) : p.pad_value;
          v = __fsub_rn(v, __fmul_rn(p.preemph, xm));
        }
      }
      x[k] = v;
    }
  }
  float2 z[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) z[m] = make_float2(x[2 * m] * p.w[2 * m], x[2 * m + 1] * p.w[2 * m + 1]);
  Dft<NC>::run(z);
  // The 32 frames of a warp are ONE contiguous run of the (T, F) output (F complex values each): rows are parked in a per-warp
  // staging block (pitch F: odd, conflict free) and leave as coalesced 256-byte stores — a thread storing its own row writes 8
  // bytes at a stride of 8 F bytes, every store instruction touching all ~90 sectors of the run.
  __shared__ float2 s_rows[8][32 * F];
  const int lane = threadIdx.x & 31;
  float2* const so = s_rows[threadIdx.x >> 5] + lane * F;
  static_for<0, NC / 2 + 1>([&](auto K_) {
    constexpr int k = decltype(K_)::value;
    const float2 a = z[k], b = z[(NC - k) % NC];
    const float ex = 0.5f * (a.x + b.x), ey = 0.5f * (a.y - b.y);
    const float ox = 0.5f * (a.y + b.y), oy = 0.5f * (b.x - a.x);
    constexpr regs::cplx_d w = regs::unit_root(-k, N);  // W_N^k
    constexpr float wr = (float)w.re, wi = (float)w.im;
    const float tx = wr * ox - wi * oy, ty = wr * oy + wi * ox;
    so[k] = make_float2(ex + tx, ey + ty);
    so[NC - k] = make_float2(ex - tx, -(ey - ty));
  });
  __syncwarp();
  const int64_t lt0 = lt - lane;                                      // first frame of this warp
  const int nrows = (int)min((int64_t)32, p.frame_count - lt0);       // the clip's last warp may be ragged
  float2* const o = p.out + (int64_t)clip_i * p.out_clip_stride + lt0 * F;
  const float2* const sw = s_rows[threadIdx.x >> 5];
  for (int i = lane; i < nrows * F; i += 32) o[i] = sw[i];
}

}  // namespace

bool small_istft_supported(const b2a_plan* plan) {
  const b2a_istft_desc& d = plan->id;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  return (d.n_fft == 20 || d.n_fft == 16) && d.hop * 4 == d.n_fft;
}

int small_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st) {
  const b2a_istft_desc& d = plan->id;
  SmallInvParams p;
  memset(&p, 0, sizeof(p));
  const int N = d.n_fft, F = plan->n_freqs;
  if (a->spec_imag) {
    p.spec_re = reinterpret_cast<const float*>(a->spec);
    p.spec_im = reinterpret_cast<const float*>(a->spec_imag);
  } else {
    p.spec = reinterpret_cast<const float2*>(a->spec);
  }
  p.T = a->num_frames;
  p.clip_stride = a->clip_stride ? a->clip_stride : (int64_t)F * a->num_frames;
  p.hop = d.hop;
  p.norm_sq = d.norm_kind == B2A_ISTFT_NORM_WINDOW_SQ;
  p.div_clamp = d.div_kind == B2A_ISTFT_DIV_CLAMP;
  p.polar = make_polar_spec(d);
  p.div_eps = istft_div_eps(d);
  int64_t ola, start, len;
  b2a_istft_geometry(a->num_frames, N, d.hop, d.center, d.trim_tail ? a->length : -1, &ola, &start, &len);
  if (!d.trim_tail) {
    start = d.center ? N / 2 : 0;
    len = std::max<int64_t>(0, ola - start);
    if (a->length >= 0 && a->length < len) len = a->length;
  }
  if (len <= 0) return B2A_OK;
  p.out_start = start;
  p.out_len = len;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : len;
  p.out = a->out;
  for (int i = 0; i < 32; ++i) p.w[i] = i < N ? plan->h_window[i] : 0.0f;
  for (int i = 0; i < 32; ++i) p.wn[i] = p.w[i] * (1.0f / (float)N);
  p.rden_ok = 1;
  for (int j = 0; j < d.hop; ++j) {  // ascending frame order: the oldest frame contributes tap j + 3 * hop
    float sum = 0.0f;
    for (int q = 3; q >= 0; --q) {
      const float w = p.w[j + q * d.hop];
      sum += p.norm_sq ? w * w : w;
    }
    if (!(sum > p.div_eps)) p.rden_ok = 0;  // guard active: keep the exact division path
    p.rden[j] = (float)(1.0 / (double)sum);
  }
  const int64_t slots = a->num_frames + 3;
  p.warps_per_clip = (int)((slots + 28) / 29);
  dim3 grid((p.warps_per_clip + 7) / 8, a->batch);
  if (N == 20 && p.polar.polar) istft_small_kernel<20, true><<<grid, 256, 0, st>>>(p);
  else if (N == 20) istft_small_kernel<20, false><<<grid, 256, 0, st>>>(p);
  else if (p.polar.polar) istft_small_kernel<16, true><<<grid, 256, 0, st>>>(p);
  else istft_small_kernel<16, false><<<grid, 256, 0, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

bool small_stft_supported(const b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  if (d.frame_dc || d.frame_preemph != 0.0f || d.dither != 0.0f || d.frame_len != 0) return false;
  // stft_small_kernel writes plain (T, F) complex rows: anything else a descriptor may ask for goes to the generic kernel
  if (d.out_layout != B2A_LAYOUT_TM || d.clamp_kind != B2A_CLAMP_NONE || d.norm_kind != B2A_NORM_NONE || d.preemph != 0.0f) return false;
  return (d.n_fft == 20 || d.n_fft == 16) && d.n_mels == 0 && d.spec_kind == B2A_SPEC_COMPLEX;
}

int small_stft(b2a_plan* plan, const b2a_forward_args* a, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  SmallFwdParams p;
  memset(&p, 0, sizeof(p));
  p.audio = a->audio;
  p.clip_stride = a->clip_stride;
  p.valid_length = a->valid_length;
  p.sample_offset = a->sample_offset;
  p.frame_begin = a->frame_begin;
  p.frame_count = a->frame_count;
  p.pad_value = a->pad_value;
  p.geo = make_geometry(a->length, d.n_fft, d.hop, d.center, d.pad_mode);
  p.hop = d.hop;
  p.pad_mode = d.pad_mode;
  p.preemph = d.preemph;
  p.out = reinterpret_cast<float2*>(a->out);
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * plan->n_freqs;
  for (int i = 0; i < 32; ++i) p.w[i] = i < d.n_fft ? plan->h_window[i] : 0.0f;
  dim3 grid((unsigned)((a->frame_count + 255) / 256), a->batch);
  if (d.n_fft == 20) stft_small_kernel<20><<<grid, 256, 0, st>>>(p);
  else stft_small_kernel<16><<<grid, 256, 0, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

}  // namespace b2a
