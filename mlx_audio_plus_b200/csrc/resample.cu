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

#include <new>

#include "common.cuh"

namespace b2a {

struct Resampler {
  int up = 1, down = 1, J = 1;
  int64_t pre_remove = 0;
  float* d_taps = nullptr;  // [J][up]
  size_t taps_bytes = 0;
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

// Single-channel path (load_audio of mono files): the tile's input span is converted to float32 ONCE into shared memory; a block then produces TILE consecutive output samples from it.  Per tap: one
// conflict-free tap read, one input read, one FMA — no global loads, no conversions, no per-channel work in the inner loop.
constexpr int kTile = 2048;  // outputs per block iteration
template <typename T>
__global__ void __launch_bounds__(256) resample_mono_tiled_kernel(const ResampleParams p, int span_max) {
  extern __shared__ float s_mem[];
  float* const s_taps = s_mem;                 // [J][up]
  float* const xs = s_mem + p.J * p.up;        // [span_max]
  for (int i = threadIdx.x; i < p.J * p.up; i += blockDim.x) s_taps[i] = p.taps[i];
  const int ch = p.channels;
  const T* x = reinterpret_cast<const T*>(p.in) + (int64_t)blockIdx.y * p.in_clip_stride;
  float* y = p.out + (int64_t)blockIdx.y * p.out_clip_stride;
  const int64_t last = p.n_in - 1;
  const int64_t tiles = (p.n_out + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t n0 = tile * kTile;
    const int64_t n1 = (n0 + kTile < p.n_out ? n0 + kTile : p.n_out) - 1;   // last output of the tile
    const int64_t base = ((n0 + p.pre_remove) * p.down) / p.up - (p.J - 1);  // lowest input index any output reads
    const int span = (int)(((n1 + p.pre_remove) * p.down) / p.up - base + 1);
    __syncthreads();  // the previous tile's reads of xs are done (and s_taps is loaded)
    for (int s = threadIdx.x; s < span; s += blockDim.x) {
      int64_t i = base + s;
      i = i < 0 ? 0 : (i > last ? last : i);  // upfirdn mode "edge"
      float v = load_sample<T>(x, i * ch);
      for (int c = 1; c < ch; ++c) v += load_sample<T>(x, i * ch + c);
      xs[s] = ch > 1 ? v / (float)ch : v;
    }
    __syncthreads();
    for (int64_t n = n0 + threadIdx.x; n <= n1; n += blockDim.x) {
      const int64_t t = (n + p.pre_remove) * p.down;
      const int64_t i0 = t / p.up;
      const float* w = s_taps + (int)(t - i0 * p.up);
      const float* xp = xs + (int)(i0 - base);
      float a0 = 0.0f, a1 = 0.0f;  // two chains: the FMA latency is not the loop's critical path
      int j = 0;
      for (; j + 1 < p.J; j += 2) {
        a0 = fmaf(w[j * p.up], xp[-j], a0);
        a1 = fmaf(w[(j + 1) * p.up], xp[-j - 1], a1);
      }
      if (j < p.J) a0 = fmaf(w[j * p.up], xp[-j], a0);
      y[n] = a0 + a1;
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
  // tiled path for single-channel input when the taps + one tile's input span fit in shared memory (measured on 1 h of
  // audio: 48 kHz mono 1.53 -> 1.32 ms; 44.1 kHz stereo is faster on the direct kernel, 2.21 vs 2.44 ms)
  if (p.mono && p.channels == 1) {
    const int span_max = (int)(((int64_t)kTile * p.down) / p.up + p.J + 2);
    const size_t need = sizeof(float) * ((size_t)p.J * p.up + span_max);
    if (need <= 200 * 1024) {
      int64_t tiles = (p.n_out + kTile - 1) / kTile;
      int64_t g = tiles < (int64_t)sm_count * 4 ? tiles : (int64_t)sm_count * 4;
      if (g < 1) g = 1;
      static size_t attr = 0;  // per element type (template instance)
      if (need > attr) {
        B2A_CUDA(cudaFuncSetAttribute(resample_mono_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        attr = need;
      }
      resample_mono_tiled_kernel<T><<<dim3((unsigned)g, (unsigned)batch), 256, need, st>>>(p, span_max);
      B2A_CUDA(cudaGetLastError());
      return B2A_OK;
    }
  }
  dim3 grid((unsigned)gx, (unsigned)batch);
  auto go = [&](auto kernel) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, 256, smem, st>>>(p);
  };
  if (p.channels == 1) go(resample_kernel<T, 1>);
  else if (p.channels == 2) go(resample_kernel<T, 2>);
  else go(resample_kernel<T, 0>);
  B2A_CUDA(cudaGetLastError());
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
  cudaError_t e = cudaMalloc(&r->d_taps, r->taps_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_taps, h_taps, r->taps_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("resampler_create: %s (no CUDA device? there is no CPU fallback)", cudaGetErrorString(e));
    cudaFree(r->d_taps);
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
  return a->in_kind == B2A_PCM_I16 ? launch_resample<int16_t>(p, sms, a->batch, st) : launch_resample<float>(p, sms, a->batch, st);
}

}  // extern "C"
