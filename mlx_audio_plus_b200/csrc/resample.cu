// b200audio — the step in front of the path: what load_audio does AFTER the decoder
// (stt/utils.py:21-57, audio_io.py:258-262): interleaved PCM (int16 / 32768, or float32) -> polyphase resampling of every
// channel (scipy.signal.resample_poly(audio, up, down, padtype="edge"), axis 0) -> mean over channels -> float32 mono.
//
// One kernel, one pass: output sample n of a clip is
//     y[n] = sum_j Hp[j][phase] * x_edge[i0 - j],   t = (n + pre_remove) * down,  phase = t mod up,  i0 = t div up,
// with x_edge the input clamped to its first / last sample (upfirdn mode "edge") and Hp the polyphase split of scipy's
// zero-padded Kaiser(5.0) windowed-sinc (designed on the host by the Python layer with NumPy, the same formulas as
// scipy.signal.firwin — tests/test_resample_cpu.py — and handed over like windows and filterbanks are).  HBM traffic is the
// compulsory input once + output once; the <= 61 taps per output sample are shared-memory reads (bank-conflict free: the
// phases of neighbouring outputs differ by `down mod up`) against an L1-resident input window.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace b2a {

struct Resampler {
  int up = 1, down = 1, J = 1;
  int64_t pre_remove = 0;
  float* d_taps = nullptr;  // [J][up]
  size_t taps_bytes = 0;
  float* d_taps_pm = nullptr;  // [up][4 * J4]: phase-major, tap count padded to a multiple of 4 with zeros
  int J4 = 1;
};

namespace {

template <typename T>
__device__ __forceinline__ float load_sample(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, int64_t i) {
  return __ldg(p + i);
}
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, int64_t i) {
  return (float)__ldg(p + i) * (1.0f / 32768.0f);  // exact: samples.astype(float) / 32768.0 (audio_io.py:260)
}

struct ResampleParams {
  const void* in;
  float* out;
  int64_t n_in, n_out, in_clip_stride, out_clip_stride, pre_remove;
  int channels, up, down, J, mono, taps_in_smem;
  const float* taps;
};

// CH > 0: channel count fixed at compile time (1, 2); CH == 0: any count, one channel at a time
template <typename T, int CH>
__global__ void __launch_bounds__(256) resample_kernel(const ResampleParams p) {
  extern __shared__ float s_taps[];
  const int ntaps = p.J * p.up;
  if (p.taps_in_smem) {
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) s_taps[i] = p.taps[i];
    __syncthreads();
  }
  const float* taps = p.taps_in_smem ? s_taps : p.taps;
  const int ch = CH > 0 ? CH : p.channels;
  const T* x = reinterpret_cast<const T*>(p.in) + (int64_t)blockIdx.y * p.in_clip_stride;
  float* y = p.out + (int64_t)blockIdx.y * p.out_clip_stride;
  const int64_t last = p.n_in - 1;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < p.n_out; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = (n + p.pre_remove) * p.down;
    const int64_t i0 = t / p.up;
    const int ph = (int)(t - i0 * p.up);
    const float* w = taps + ph;
    if (CH > 0) {
      float acc[CH > 0 ? CH : 1];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
      if (i0 - (p.J - 1) >= 0 && i0 <= last) {  // interior: no clamping
        const T* xp = x + i0 * CH;
        for (int j = 0; j < p.J; ++j) {
          const float wj = w[j * p.up];
#pragma unroll
          for (int c = 0; c < CH; ++c) acc[c] = fmaf(wj, load_sample<T>(xp, c - (int64_t)j * CH), acc[c]);
        }
      } else {
        for (int j = 0; j < p.J; ++j) {
          int64_t i = i0 - j;
          i = i < 0 ? 0 : (i > last ? last : i);
          const float wj = w[j * p.up];
#pragma unroll
          for (int c = 0; c < CH; ++c) acc[c] = fmaf(wj, load_sample<T>(x, i * CH + c), acc[c]);
        }
      }
      if (p.mono) {  // mx.array(audio, float32).mean(axis=1): float32 sum over channels, then the division
        float s = acc[0];
#pragma unroll
        for (int c = 1; c < CH; ++c) s += acc[c];
        y[n] = CH > 1 ? s / (float)CH : s;
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) y[n * CH + c] = acc[c];
      }
    } else {
      float s = 0.0f;
      for (int c = 0; c < ch; ++c) {
        float acc = 0.0f;
        for (int j = 0; j < p.J; ++j) {
          int64_t i = i0 - j;
          i = i < 0 ? 0 : (i > last ? last : i);
          acc = fmaf(w[j * p.up], load_sample<T>(x, i * ch + c), acc);
        }
        if (p.mono) s += acc;
        else y[n * ch + c] = acc;
      }
      if (p.mono) y[n] = s / (float)ch;
    }
  }
}

// Mono output (load_audio): SAME-PHASE lane mapping.  A tile is 32 * up * R consecutive output samples; a warp takes one
// (q, r) item at a time and its lanes compute outputs n = N0 + q + up * (lane + 32 r).  All 32 share the filter phase
// ((n + pre) * down mod up does not depend on the lane), so the taps are warp-uniform LDS.128 broadcasts of 4 taps each from
// a [phase][tap] table, and the lanes' input windows start `down` samples apart — an odd stride for every common rate pair,
// i.e. bank-conflict free.  Per tap: one input LDS + one FMA + a quarter of a tap load (the direct kernel spends ~6
// instructions and ~4 shared-memory wavefronts per tap).  The tile's input span is converted to float32 and mixed to mono
// ONCE into shared memory (the filter is linear: mean first, then filter); results are staged at an odd pitch and written
// out as contiguous rows.
struct PhaseParams {
  const void* in;
  float* out;
  int64_t n_in, n_out, in_clip_stride, out_clip_stride, pre_remove;
  int channels, up, down, J4, R, tile, ypitch, span_max;
  int direct;  // 1: results go straight to global memory (no staging rows): two CTAs fit one SM
  const float* taps_pm;  // [up][4 * J4], zero padded
};

template <typename T, int CH>  // CH: 1, 2 or 0 (any channel count)
__global__ void __launch_bounds__(512) resample_phase_kernel(const PhaseParams p) {
  extern __shared__ float4 s_mem4[];
  float* const s_taps = reinterpret_cast<float*>(s_mem4);            // [up][4 * J4]
  float* const xs = s_taps + p.up * 4 * p.J4;                         // [span_max]
  float* const ys = xs + ((p.span_max + 3) & ~3);                     // [32 * R * ypitch]
  const int JP = 4 * p.J4;
  for (int i = threadIdx.x; i < p.up * JP; i += blockDim.x) s_taps[i] = p.taps_pm[i];
  const int ch = p.channels;
  const T* x = reinterpret_cast<const T*>(p.in) + (int64_t)blockIdx.y * p.in_clip_stride;
  float* y = p.out + (int64_t)blockIdx.y * p.out_clip_stride;
  const int64_t last = p.n_in - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t tiles = (p.n_out + p.tile - 1) / p.tile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t n0 = tile * p.tile;
    const int cnt = (int)(p.n_out - n0 < p.tile ? p.n_out - n0 : p.tile);  // outputs in this tile
    const int64_t t0 = (n0 + p.pre_remove) * p.down;
    const int64_t b0 = t0 / p.up;
    const int m0 = (int)(t0 - b0 * p.up);
    const int64_t base = b0 - (JP - 1);                                    // input index of xs[0]
    const int span = (int)(((n0 + cnt - 1 + p.pre_remove) * p.down) / p.up - base + 1);
    __syncthreads();  // previous tile: xs / ys free (and s_taps loaded)
    // fill: 8 frames per thread per round with every global load issued before the first use (one load in flight per
    // thread made this phase latency bound: 1.3 ms -> see DESIGN.md K5)
    for (int s0 = threadIdx.x; s0 < span; s0 += 8 * blockDim.x) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int64_t i = base + s0 + k * (int)blockDim.x;
        i = i < 0 ? 0 : (i > last ? last : i);  // upfirdn mode "edge" (indices past the span are clamped too: harmless)
        if (CH == 1) {
          v[k] = load_sample<T>(x, i);
        } else if (CH == 2) {
          v[k] = (load_sample<T>(x, 2 * i) + load_sample<T>(x, 2 * i + 1)) / 2.0f;
        } else {
          float a = load_sample<T>(x, i * ch);
          for (int c = 1; c < ch; ++c) a += load_sample<T>(x, i * ch + c);
          v[k] = a / (float)ch;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (s0 + k * (int)blockDim.x < span) xs[s0 + k * (int)blockDim.x] = v[k];
    }
    __syncthreads();
    const int items = p.up * p.R;
    // two items per round, four accumulator chains each: 16 independent input loads per tap quad pair are in flight (with
    // one item and two chains the loop waited out a shared-memory latency per FMA: short_scoreboard 2.3 per issue)
    for (int item = 2 * warp; item < items; item += 2 * nwarps) {
      const float4* w4[2];
      const float* xp[2];
      int dst[2], gdst[2];
      bool ok[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int it = item + e < items ? item + e : item;
        const int r = it / p.up, q = it - r * p.up;
        const int l = lane + 32 * r;                  // "row" of this lane's output inside the tile
        const unsigned t = (unsigned)m0 + (unsigned)q * (unsigned)p.down;   // < up + up * down: fits 32 bits
        const unsigned bq = t / (unsigned)p.up;
        const int ph = (int)(t - bq * (unsigned)p.up);
        ok[e] = item + e < items && q + p.up * l < cnt;
        w4[e] = reinterpret_cast<const float4*>(s_taps + ph * JP);
        xp[e] = xs + (ok[e] ? (int)bq + p.down * l : 0) + (JP - 1);  // tap j reads xp[-j]
        dst[e] = q + p.ypitch * l;
        gdst[e] = q + p.up * l;
      }
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 2
      for (int j4 = 0; j4 < p.J4; ++j4) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 w = w4[e][j4];
          acc[e][0] = fmaf(w.x, xp[e][-4 * j4], acc[e][0]);
          acc[e][1] = fmaf(w.y, xp[e][-4 * j4 - 1], acc[e][1]);
          acc[e][2] = fmaf(w.z, xp[e][-4 * j4 - 2], acc[e][2]);
          acc[e][3] = fmaf(w.w, xp[e][-4 * j4 - 3], acc[e][3]);
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (ok[e]) {
          const float r = (acc[e][0] + acc[e][1]) + (acc[e][2] + acc[e][3]);
          if (p.direct) y[n0 + gdst[e]] = r;  // lanes are `up` outputs apart: a scattered store, but no staging rows
          else ys[dst[e]] = r;
        }
    }
    if (p.direct) continue;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int l = i / p.up, q = i - l * p.up;
      y[n0 + i] = ys[q + p.ypitch * l];
    }
  }
}

template <typename T>
int launch_resample(const ResampleParams& p, int sm_count, int64_t batch, cudaStream_t st) {
  const size_t smem = p.taps_in_smem ? sizeof(float) * (size_t)p.J * p.up : 0;
  int64_t gx = (p.n_out + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 16;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)batch);
  auto go = [&](auto kernel) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, 256, smem, st>>>(p);
  };
  if (p.channels == 1) go(resample_kernel<T, 1>);
  else if (p.channels == 2) go(resample_kernel<T, 2>);
  else go(resample_kernel<T, 0>);
  B2A_LAUNCHED();
  return B2A_OK;
}

}  // namespace
}  // namespace b2a

using namespace b2a;

extern "C" {

int b2a_resampler_create(int32_t up, int32_t down, int32_t taps_per_phase, int64_t pre_remove, const float* h_taps,
                         b2a_resampler** out) {
  if (!out || !h_taps || up <= 0 || down <= 0 || taps_per_phase <= 0 || pre_remove < 0) {
    set_error("resampler_create: up=%d down=%d taps_per_phase=%d", up, down, taps_per_phase);
    return B2A_ERR_INVALID_ARG;
  }
  Resampler* r = new (std::nothrow) Resampler();
  if (!r) return B2A_ERR_NOMEM;
  r->up = up;
  r->down = down;
  r->J = taps_per_phase;
  r->pre_remove = pre_remove;
  r->taps_bytes = sizeof(float) * (size_t)up * taps_per_phase;
  r->J4 = (taps_per_phase + 3) / 4;
  std::vector<float> pm((size_t)up * 4 * r->J4, 0.0f);
  for (int j = 0; j < taps_per_phase; ++j)
    for (int ph = 0; ph < up; ++ph) pm[(size_t)ph * 4 * r->J4 + j] = h_taps[(size_t)j * up + ph];
  cudaError_t e = cudaMalloc(&r->d_taps, r->taps_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_taps, h_taps, r->taps_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_taps_pm, sizeof(float) * pm.size());
  if (e == cudaSuccess) e = cudaMemcpy(r->d_taps_pm, pm.data(), sizeof(float) * pm.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("resampler_create: %s (no CUDA device? there is no CPU fallback)", cudaGetErrorString(e));
    cudaFree(r->d_taps);
    cudaFree(r->d_taps_pm);
    delete r;
    return B2A_ERR_CUDA;
  }
  *out = reinterpret_cast<b2a_resampler*>(r);
  return B2A_OK;
}

int b2a_resampler_destroy(b2a_resampler* h) {
  Resampler* r = reinterpret_cast<Resampler*>(h);
  if (!r) return B2A_OK;
  cudaFree(r->d_taps);
  cudaFree(r->d_taps_pm);
  delete r;
  return B2A_OK;
}

int b2a_resampler_out_len(const b2a_resampler* h, int64_t n_in, int64_t* n_out) {
  const Resampler* r = reinterpret_cast<const Resampler*>(h);
  if (!r || !n_out || n_in < 0) return B2A_ERR_INVALID_ARG;
  *n_out = (n_in * r->up + r->down - 1) / r->down;  // resample_poly: n_in * up // down + bool(n_in * up % down)
  return B2A_OK;
}

int b2a_resample(b2a_resampler* h, const b2a_resample_args* a, void* stream) {
  Resampler* r = reinterpret_cast<Resampler*>(h);
  if (!r || !a || !a->in || !a->out || a->channels <= 0 || a->batch <= 0 || a->n_in <= 0 ||
      (a->in_kind != B2A_PCM_F32 && a->in_kind != B2A_PCM_I16)) {
    set_error("resample: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("resample: no CUDA device (there is no CPU fallback)");
    return B2A_ERR_CUDA;
  }
  ResampleParams p;
  p.in = a->in;
  p.out = a->out;
  p.n_in = a->n_in;
  p.n_out = (a->n_in * r->up + r->down - 1) / r->down;
  p.channels = a->channels;
  p.in_clip_stride = a->in_clip_stride ? a->in_clip_stride : a->n_in * a->channels;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : p.n_out * (a->mono ? 1 : a->channels);
  p.pre_remove = r->pre_remove;
  p.up = r->up;
  p.down = r->down;
  p.J = r->J;
  p.mono = a->mono;
  p.taps = r->d_taps;
  p.taps_in_smem = r->taps_bytes <= 160 * 1024;
  if (a->batch > 65535) {
    set_error("resample: batch %lld > 65535", (long long)a->batch);
    return B2A_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (a->mono && (int64_t)r->up * r->down < ((int64_t)1 << 30)) {  // same-phase kernel (32-bit phase arithmetic inside a tile)
    PhaseParams q;
    q.in = a->in; q.out = a->out; q.n_in = p.n_in; q.n_out = p.n_out;
    q.in_clip_stride = p.in_clip_stride; q.out_clip_stride = p.out_clip_stride; q.pre_remove = r->pre_remove;
    q.channels = a->channels; q.up = r->up; q.down = r->down; q.J4 = r->J4;
    q.R = (4096 + 32 * r->up - 1) / (32 * r->up);
    if (q.R < 1) q.R = 1;
    q.tile = 32 * r->up * q.R;
    q.ypitch = r->up | 1;
    q.span_max = (int)(((int64_t)q.tile * r->down) / r->up + 4 * r->J4 + 2);
    q.taps_pm = r->d_taps_pm;
    size_t need = sizeof(float) * ((size_t)r->up * 4 * r->J4 + ((q.span_max + 3) & ~3) + (size_t)32 * q.R * q.ypitch) + 16;
    // The kernel is bound by shared-memory latency at one 16-warp CTA per SM (44.1 -> 16 kHz: 116 KB of taps, input span and
    // staging rows).  Without the staging rows (20 KB) two CTAs fit: the results then leave as scattered 4-byte stores, which the
    // L2 merges — the output is a quarter of the traffic.
    const size_t need_direct = need - sizeof(float) * (size_t)32 * q.R * q.ypitch;
    static const int direct_mode = getenv("B2A_X_RS_DIRECT") ? atoi(getenv("B2A_X_RS_DIRECT")) : 1;  // development: 0 keeps the staging rows
    q.direct = direct_mode && need > 113 * 1024 && need_direct <= 113 * 1024;
    if (q.direct) need = need_direct;
    if (need <= 200 * 1024) {
      const int64_t tiles = (p.n_out + q.tile - 1) / q.tile;
      int64_t g = tiles < (int64_t)sms * 2 ? tiles : (int64_t)sms * 2;
      if (g < 1) g = 1;
      dim3 grid((unsigned)g, (unsigned)a->batch);
      auto go = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        kernel<<<grid, 512, need, st>>>(q);
      };
      const bool i16 = a->in_kind == B2A_PCM_I16;
      if (a->channels == 1) i16 ? go(resample_phase_kernel<int16_t, 1>) : go(resample_phase_kernel<float, 1>);
      else if (a->channels == 2) i16 ? go(resample_phase_kernel<int16_t, 2>) : go(resample_phase_kernel<float, 2>);
      else i16 ? go(resample_phase_kernel<int16_t, 0>) : go(resample_phase_kernel<float, 0>);
      B2A_LAUNCHED();
      return B2A_OK;
    }
  }
  return a->in_kind == B2A_PCM_I16 ? launch_resample<int16_t>(p, sms, a->batch, st) : launch_resample<float>(p, sms, a->batch, st);
}

}  // extern "C"
