// b200audio — the steps right after the path: the layouts the encoders read.
//   b2a_rows_pad_cast  Whisper's segment builder, whisper/whisper.py:990-996:
//                      pad_or_trim(mel[seek : seek + segment_size], N_FRAMES, axis=-2).astype(dtype)
//                      -> (batch, rows_out, cols) float32 / float16 / bfloat16, zero rows after the valid ones.
//   b2a_lfr            FunASR low-frame-rate stacking, funasr/audio.py:84-139 (first / last frame replicated at the
//                      ends, lfr_m frames stacked every lfr_n), with the precomputed CMVN of apply_cmvn
//                      ((x + shift) * scale, funasr/audio.py:166-169) applied on the way out when given.
// Both are single-pass, HBM-bound copies: every input row is read once (LFR: lfr_m / lfr_n times, from L2), every
// output element written once.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace b2a {
namespace {

template <typename T>
__device__ __forceinline__ T cast_out(float v);
template <>
__device__ __forceinline__ float cast_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half cast_out<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int VEC>
__global__ void __launch_bounds__(256) rows_pad_cast_kernel(const float* __restrict__ in, int64_t in_clip_stride, int64_t row_begin,
                                                             int64_t rows_valid, int cols, T* __restrict__ out, int64_t rows_out) {
  const int64_t per_clip = rows_out * cols, valid = rows_valid * cols;
  const float* src = in + (int64_t)blockIdx.y * in_clip_stride + row_begin * cols;
  T* dst = out + (int64_t)blockIdx.y * per_clip;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < per_clip; i += (int64_t)gridDim.x * blockDim.x * VEC) {
    float v[VEC];
    if (VEC == 4 && i + 3 < valid) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(src + i));
      v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = (i + k < valid) ? __ldg(src + i + k) : 0.0f;
    }
    T o[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) o[k] = cast_out<T>(v[k]);
    if (VEC == 4) {
      if (sizeof(T) == 4) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<float4*>(o);
      else *reinterpret_cast<float2*>(dst + i) = *reinterpret_cast<float2*>(o);
    } else {
      dst[i] = o[0];
    }
  }
}

template <typename T>
int launch_rows(const float* in, int64_t in_clip_stride, int64_t row_begin, int64_t rows_valid, int cols, void* out,
                int64_t rows_out, int batch, cudaStream_t st) {
  const int64_t per_clip = rows_out * cols;
  const bool vec = per_clip % 4 == 0 && (row_begin * cols) % 4 == 0 && in_clip_stride % 4 == 0 &&
                   reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  const int64_t work = vec ? per_clip / 4 : per_clip;
  int64_t gx = (work + 255) / 256;
  if (gx > 148 * 16) gx = 148 * 16;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)batch);
  if (vec) rows_pad_cast_kernel<T, 4><<<grid, 256, 0, st>>>(in, in_clip_stride, row_begin, rows_valid, cols, (T*)out, rows_out);
  else rows_pad_cast_kernel<T, 1><<<grid, 256, 0, st>>>(in, in_clip_stride, row_begin, rows_valid, cols, (T*)out, rows_out);
  B2A_LAUNCHED();
  return B2A_OK;
}

// VEC = 4: four consecutive mel bins of one stacked frame per thread (n_mels % 4 == 0: they never straddle a frame)
template <int VEC>
__global__ void __launch_bounds__(256) lfr_kernel(const float* __restrict__ in, int64_t in_clip_stride, int64_t frames, int n_mels, int lfr_m,
                                                   int lfr_n, const float* __restrict__ shift, const float* __restrict__ scale,
                                                   float* __restrict__ out, int64_t out_clip_stride, int64_t t_lfr) {
  const int width = lfr_m * n_mels, left = (lfr_m - 1) / 2;
  const float* src = in + (int64_t)blockIdx.y * in_clip_stride;
  float* dst = out + (int64_t)blockIdx.y * out_clip_stride;
  const int64_t total = t_lfr * width;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total; i += (int64_t)gridDim.x * blockDim.x * VEC) {
    const int64_t t = i / width;
    const int c = (int)(i - t * width);
    const int k = c / n_mels, m = c - k * n_mels;
    int64_t f = t * lfr_n + k - left;  // index into the unpadded features; the pads replicate frame 0 / frame T-1
    f = f < 0 ? 0 : (f > frames - 1 ? frames - 1 : f);
    if (VEC == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + f * n_mels + m));
      if (shift) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(shift + c)), b = __ldg(reinterpret_cast<const float4*>(scale + c));
        v.x = (v.x + a.x) * b.x; v.y = (v.y + a.y) * b.y; v.z = (v.z + a.z) * b.z; v.w = (v.w + a.w) * b.w;
      }
      *reinterpret_cast<float4*>(dst + i) = v;
    } else {
      float v = __ldg(src + f * n_mels + m);
      if (shift) v = (v + __ldg(shift + c)) * __ldg(scale + c);
      dst[i] = v;
    }
  }
}

// per-utterance CMVN (funasr/audio.py:160-164): mean / std over the frames of each feature column (mx.std: ddof 0),
// (x - mean) / (std + eps).  Pass 1: per-column sum and sum of squares in float64 (one atomicAdd pair per block and column);
// pass 2: apply.  stats: [batch][cols][2] doubles, zeroed by the caller side of the ABI function.
// Both passes: a block owns a slab of rows, a thread owns V adjacent columns (V = 4: 16-byte loads when the rows allow it)
// and walks the slab row by row, so a warp reads one contiguous run per row.
template <int V>
struct CmvnVec;
template <>
struct CmvnVec<1> {
  using T = float;
  static __device__ __forceinline__ void get(const T& v, float* e) { e[0] = v; }
  static __device__ __forceinline__ T make(const float* e) { return e[0]; }
};
template <>
struct CmvnVec<4> {
  using T = float4;
  static __device__ __forceinline__ void get(const T& v, float* e) { e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w; }
  static __device__ __forceinline__ T make(const float* e) { return make_float4(e[0], e[1], e[2], e[3]); }
};

template <int V>
__global__ void __launch_bounds__(256) cmvn_stats_kernel(const float* __restrict__ x, int64_t clip_stride, int64_t rows, int cols, double* stats) {
  using T = typename CmvnVec<V>::T;
  const T* src = reinterpret_cast<const T*>(x + (int64_t)blockIdx.y * clip_stride);
  double* st = stats + (int64_t)blockIdx.y * cols * 2;
  const int cv = cols / V;  // columns in units of V
  const int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  if (r1 <= r0) return;
  for (int c = threadIdx.x; c < cv; c += blockDim.x) {
    double s1[V], s2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s1[k] = s2[k] = 0.0;
#pragma unroll 4
    for (int64_t r = r0; r < r1; ++r) {
      float e[V];
      CmvnVec<V>::get(__ldg(src + r * cv + c), e);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const double v = (double)e[k];
        s1[k] += v;
        s2[k] += v * v;
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      atomicAdd(st + 2 * (c * V + k), s1[k]);
      atomicAdd(st + 2 * (c * V + k) + 1, s2[k]);
    }
  }
}

// (x - mean) / (std + eps) keeps the reference's true division; mean and std + eps are formed once per thread and column
template <int V>
__global__ void __launch_bounds__(256) cmvn_apply_kernel(const float* x, float* out, int64_t clip_stride, int64_t rows, int cols,
                                                          const double* stats, float eps) {
  using T = typename CmvnVec<V>::T;
  const T* src = reinterpret_cast<const T*>(x + (int64_t)blockIdx.y * clip_stride);
  T* dst = reinterpret_cast<T*>(out + (int64_t)blockIdx.y * clip_stride);
  const double* st = stats + (int64_t)blockIdx.y * cols * 2;
  const int cv = cols / V;
  const int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int c = threadIdx.x; c < cv; c += blockDim.x) {
    float m[V], sd[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const double mean = st[2 * (c * V + k)] / (double)rows;
      double var = st[2 * (c * V + k) + 1] / (double)rows - mean * mean;
      if (var < 0.0) var = 0.0;
      m[k] = (float)mean;
      sd[k] = (float)sqrt(var) + eps;
    }
#pragma unroll 4
    for (int64_t r = r0; r < r1; ++r) {
      float e[V];
      CmvnVec<V>::get(src[r * cv + c], e);  // plain load: x and out may alias
#pragma unroll
      for (int k = 0; k < V; ++k) e[k] = (e[k] - m[k]) / sd[k];
      dst[r * cv + c] = CmvnVec<V>::make(e);
    }
  }
}

// (rows, cols) -> (cols, rows_out) per clip through a 32 x 33 shared-memory tile; columns rows .. rows_out-1 of the output are
// zeros (Sortformer's pad_to, sortformer.py:112-118: the (B, n_mels, T_padded) layout its encoder reads).
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int cols,
                                                            int64_t rows_out) {
  __shared__ float tile[32][33];
  const float* src = in + (int64_t)blockIdx.z * rows * cols;
  float* dst = out + (int64_t)blockIdx.z * cols * rows_out;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k;
    const int c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < rows && c < cols) ? __ldg(src + r * cols + c) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const int64_t r = r0 + tx;
    if (c < cols && r < rows_out) dst[(int64_t)c * rows_out + r] = tile[tx][ty + 8 * k];
  }
}

}  // namespace

// ---- waveform-side helpers of the drop-ins (one launch each instead of a chain of eager array ops) -----------------
// block-wide sum of one double per thread (blockDim.x a multiple of 32, <= 1024); every thread gets the total
__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// Row-wise zero-mean / unit-variance of waveforms (B, L): out = (x - mean) / den over the first `valid[b]` samples (all L
// when valid == nullptr), `pad_value` behind them.  den_kind 0: sqrt(var + eps) — transformers' zero_mean_unit_var_norm
// (the Qwen3-ASR extractor's do_normalize); 1: max(std, eps) — smart_turn.py:196-199.  Population variance (ddof 0), mean
// and centred sum of squares accumulated in float64 (two sweeps over a row that stays in L2), one CTA per row.
__global__ void __launch_bounds__(512) rows_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t cols,
                                                              const int64_t* __restrict__ valid, int den_kind, float eps, float pad_value) {
  __shared__ double red[32];
  const float* x = in + (int64_t)blockIdx.x * cols;
  float* o = out + (int64_t)blockIdx.x * cols;
  int64_t n = valid ? valid[blockIdx.x] : cols;
  n = n < 0 ? 0 : (n > cols ? cols : n);
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  const double cnt = n > 0 ? (double)n : 1.0;
  const float mean = (float)(block_sum(s, red) / cnt);
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[i] - mean;
    q += (double)d * (double)d;
  }
  const float var = (float)(block_sum(q, red) / cnt);
  const float den = den_kind == 0 ? sqrtf(var + eps) : fmaxf(sqrtf(var), eps);
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) o[i] = i < n ? __fdiv_rn(x[i] - mean, den) : pad_value;
}

// Phase unwrap along the last axis (kokoro/istftnet.py:418-452, numpy.unwrap's algorithm in float32): out[t] = p[t] +
// cumsum(corr)[t], corr[t] = wrap(p[t] - p[t-1]) - (p[t] - p[t-1]) where |p[t] - p[t-1]| >= discont, else 0.
// One CTA per row; 4 consecutive samples per thread and sweep, block-wide inclusive scan of the thread sums (warp shuffles),
// running carry across sweeps.  (The reference's cumsum is a float32 prefix sum whose rounding depends on the evaluation
// order; this one associates per thread / warp / sweep — same magnitude of rounding error, not the same bits.)
__device__ __forceinline__ float unwrap_corr(float cur, float prev, float discont, float period) {
  const float hi = 0.5f * period, lo = -0.5f * period;
  const float dd = __fsub_rn(cur, prev);
  float ddmod = __fsub_rn(dd, __fmul_rn(period, floorf(__fdiv_rn(__fsub_rn(dd, lo), period))));
  if (fabsf(__fsub_rn(dd, hi)) < 1e-10f && dd > 0.0f) ddmod = hi;
  return fabsf(dd) < discont ? 0.0f : __fsub_rn(ddmod, dd);
}

__global__ void __launch_bounds__(256) unwrap_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t cols, float discont,
                                                          float period) {
  __shared__ float wsum[8];
  __shared__ float s_carry;
  const float* p = in + (int64_t)blockIdx.x * cols;
  float* o = out + (int64_t)blockIdx.x * cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0.0f;
  __syncthreads();
  for (int64_t base = 0; base < cols; base += 1024) {
    const int64_t i0 = base + 4 * (int64_t)threadIdx.x;
    float v[4], c[4];
    float prev = (i0 > 0 && i0 - 1 < cols) ? p[i0 - 1] : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = i0 + k;
      v[k] = i < cols ? p[i] : 0.0f;
      c[k] = (i > 0 && i < cols) ? unwrap_corr(v[k], prev, discont, period) : 0.0f;
      prev = v[k];
    }
    c[1] += c[0];
    c[2] += c[1];
    c[3] += c[2];
    float incl = c[3];  // inclusive scan of the thread totals across the warp, then across the 8 warps
    for (int d = 1; d < 32; d <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    float off = s_carry;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    const float excl = off + (incl - c[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k < cols) o[i0 + k] = v[k] + (excl + c[k]);
    __syncthreads();
    if (threadIdx.x == 255) s_carry = off + incl;
    __syncthreads();
  }
}

}  // namespace b2a

using namespace b2a;

extern "C" {

int b2a_rows_pad_cast(const float* in, int64_t in_clip_stride, int64_t row_begin, int64_t rows_valid, int32_t cols, void* out,
                      int64_t rows_out, int32_t out_dtype, int32_t batch, void* stream) {
  if (!in || !out || cols <= 0 || rows_out <= 0 || batch <= 0 || batch > 65535 || row_begin < 0 || rows_valid < 0) {
    set_error("rows_pad_cast: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  if (rows_valid > rows_out) rows_valid = rows_out;  // trim
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_dtype) {
    case B2A_DTYPE_F32: return launch_rows<float>(in, in_clip_stride, row_begin, rows_valid, cols, out, rows_out, batch, st);
    case B2A_DTYPE_F16: return launch_rows<__half>(in, in_clip_stride, row_begin, rows_valid, cols, out, rows_out, batch, st);
    case B2A_DTYPE_BF16: return launch_rows<__nv_bfloat16>(in, in_clip_stride, row_begin, rows_valid, cols, out, rows_out, batch, st);
    default: set_error("rows_pad_cast: out_dtype %d", out_dtype); return B2A_ERR_INVALID_ARG;
  }
}

int b2a_cmvn_utterance(const float* in, float* out, int64_t clip_stride, int64_t rows, int32_t cols, float eps, double* stats_ws,
                       int32_t batch, void* stream) {
  if (!in || !out || !stats_ws || rows <= 0 || cols <= 0 || batch <= 0 || batch > 65535) {
    set_error("cmvn_utterance: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cs = clip_stride ? clip_stride : rows * cols;
  B2A_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)cols * batch, st));
  const bool vec = cols % 4 == 0 && cs % 4 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  const int cv = vec ? cols / 4 : cols;
  const int threads = cv >= 256 ? 256 : ((cv + 31) / 32) * 32;  // whole warps, no idle warps on narrow feature rows
  const int64_t slab = 64 * (256 / threads);                    // rows per block: ~16 K (row, column-unit) pairs
  int64_t gx = (rows + slab - 1) / slab;
  if (gx > 148 * 16) gx = 148 * 16;
  const dim3 grid((unsigned)gx, (unsigned)batch);
  if (vec) {
    cmvn_stats_kernel<4><<<grid, threads, 0, st>>>(in, cs, rows, cols, stats_ws);
    cmvn_apply_kernel<4><<<grid, threads, 0, st>>>(in, out, cs, rows, cols, stats_ws, eps);
  } else {
    cmvn_stats_kernel<1><<<grid, threads, 0, st>>>(in, cs, rows, cols, stats_ws);
    cmvn_apply_kernel<1><<<grid, threads, 0, st>>>(in, out, cs, rows, cols, stats_ws, eps);
  }
  B2A_LAUNCHED();
  return B2A_OK;
}

int b2a_rows_normalize(const float* in, float* out, int64_t rows, int64_t cols, const int64_t* valid, int32_t den_kind, float eps,
                       float pad_value, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || rows > 2147483647LL || (den_kind != 0 && den_kind != 1)) {
    set_error("rows_normalize: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  rows_normalize_kernel<<<(unsigned)rows, 512, 0, (cudaStream_t)stream>>>(in, out, cols, valid, den_kind, eps, pad_value);
  B2A_LAUNCHED();
  return B2A_OK;
}

int b2a_unwrap(const float* in, float* out, int64_t rows, int64_t cols, float discont, float period, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || rows > 2147483647LL || !(period > 0.0f)) {
    set_error("unwrap: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  unwrap_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(in, out, cols, discont, period);
  B2A_LAUNCHED();
  return B2A_OK;
}

int b2a_transpose_pad(const float* in, float* out, int64_t rows, int32_t cols, int64_t rows_out, int32_t batch, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || rows_out < rows || batch <= 0 || batch > 65535 || (cols + 31) / 32 > 65535) {
    set_error("transpose_pad: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  const dim3 grid((unsigned)((rows_out + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)batch);
  transpose_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, rows_out);
  B2A_LAUNCHED();
  return B2A_OK;
}

int b2a_lfr(const float* in, int64_t in_clip_stride, int64_t frames, int32_t n_mels, int32_t lfr_m, int32_t lfr_n,
            const float* cmvn_shift, const float* cmvn_scale, float* out, int64_t out_clip_stride, int32_t batch, void* stream) {
  if (!in || !out || frames <= 0 || n_mels <= 0 || lfr_m <= 0 || lfr_n <= 0 || batch <= 0 || batch > 65535 ||
      ((cmvn_shift == nullptr) != (cmvn_scale == nullptr))) {
    set_error("lfr: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  const int64_t t_lfr = (frames + lfr_n - 1) / lfr_n;  // ceil(T / lfr_n), funasr/audio.py:114
  const int64_t total = t_lfr * lfr_m * n_mels;
  const int64_t ics = in_clip_stride ? in_clip_stride : frames * n_mels, ocs = out_clip_stride ? out_clip_stride : total;
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  const bool vec = n_mels % 4 == 0 && ics % 4 == 0 && ocs % 4 == 0 && al16(in) && al16(out) && al16(cmvn_shift) && al16(cmvn_scale);
  int64_t gx = ((vec ? total / 4 : total) + 255) / 256;
  if (gx > 148 * 32) gx = 148 * 32;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)batch);
  if (vec) lfr_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(in, ics, frames, n_mels, lfr_m, lfr_n, cmvn_shift, cmvn_scale, out, ocs, t_lfr);
  else lfr_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(in, ics, frames, n_mels, lfr_m, lfr_n, cmvn_shift, cmvn_scale, out, ocs, t_lfr);
  B2A_LAUNCHED();
  return B2A_OK;
}

}  // extern "C"
