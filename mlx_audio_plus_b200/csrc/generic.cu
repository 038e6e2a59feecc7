// b200audio — generic kernels: correct for ANY n_fft whose prime factors are <= 31.
//
// Forward (replaces the MLX op chain of dsp.py:118-141 + each wrapper's epilogue): one CTA owns a tile of
// consecutive frames of one clip.  The tile's contiguous sample span is staged ONCE in shared memory
// (padding, virtual right-pad and pre-emphasis are resolved on the way in), so overlapping frames re-read
// smem, not HBM.  Two real frames are packed as one complex sequence and transformed by a mixed-radix
// Stockham FFT in shared memory; the spectra are separated by Hermitian symmetry, reduced to
// power / magnitude, projected through the (banded, CSR) mel filterbank, logged, staged and written with
// coalesced stores.  Per-clip max/min and per-mel sums are reduced per tile and merged with atomics.
//
// Inverse (dsp.py:183-217 / 350-417): gather-form overlap-add.  A CTA owns a contiguous range of OUTPUT
// samples, inverse-transforms exactly the frames that touch it (two Hermitian spectra packed per complex
// FFT), windows them into shared memory, and each output sample sums its <= ceil(N/hop) frames in frame
// order (the order the reference's sequential scatter-add uses), divides by the window envelope computed
// the same way, and is written once.  No atomics, no materialised index arrays.
#include <algorithm>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"
#include "fft_regs.cuh"

namespace b2a {

namespace {

constexpr int kThreads = 512;

struct FftDesc {
  int n;
  int nstages;
  int radix[kMaxStages];
  const float2* tw_plain;  // W_n^k, k < n, in GLOBAL memory: only the prime-radix pass reads it
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT-class approximation, flush-to-zero
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Shared-memory FFT buffers are SKEWED: element e of a sequence lives at e + (e >> 4) (one float2 of padding per 16), sequences
// are seq_pitch(n) float2 apart.  The first Stockham pass writes R consecutive outputs per thread, i.e. the lanes of a warp
// store at a stride of R float2 — with R = 16 that is one bank pair for the whole warp (16-way conflicts on every store);
// the skew turns every power-of-two stride into an odd one.
__host__ __device__ __forceinline__ int skew(int e) { return e + (e >> 4); }
__host__ __device__ __forceinline__ int seq_pitch(int n) { return n + (n >> 4) + 1; }

// Exact idx / d for idx * d-independent small operands (idx < 2^20, d < 2^12): one multiply-high instead of a division sequence
__device__ __forceinline__ int fast_div(int idx, int d, unsigned magic) {
  (void)d;
  return (int)__umulhi((unsigned)idx, magic);
}
// ceil(2^32 / d) = floor((2^32 - 1) / d) + 1, in 32-bit arithmetic: written as a 64-bit quotient this was a ~90-instruction
// library call per thread in every FFT pass (the kernels compute the magic of a pass's butterfly count on the fly) — 6-9 % of
// the generic kernels' instructions on the ncu source page
__host__ __device__ inline unsigned div_magic(int d) { return d <= 1 ? 0u : 0xFFFFFFFFu / (unsigned)d + 1u; }

// One Stockham pass of radix R over `count` independent length-n sequences laid out back to back.
// src / dst: [count][n] float2.  tw: W_n^k table (SHARED memory: the generic kernels stage it once per CTA — read from global
// memory it put an L2 round trip into every butterfly).  Ns = product of the radices already applied.  The radix-R butterfly is
// the register codelet of fft_regs.cuh (2, 3, 4, 5 hand written; 6, 8, 10, 12, 15, 16, 20, 25, 32 composed at compile time), so a
// 1920-point transform is three passes (16 x 8 x 15) instead of six (4 x 4 x 4 x 2 x 3 x 5), 2048 is 16 x 16 x 8.
// Where a pass reads element e of sequence `seq` from: a skewed shared-memory buffer, or — first pass of the forward kernel —
// the staged sample span itself: two real frames, windowed and packed on the fly (FrameSrc), which removes the window / pack
// sweep over the tile and its barrier (9 % of the S3Gen-mel kernel's time on the ncu source page).
struct SmemSrc {
  const float2* p;
  int pitch;
  __device__ __forceinline__ float2 operator()(int seq, int e) const { return p[(size_t)seq * pitch + skew(e)]; }
};
struct FrameSrc {
  const float* xs;      // the tile's sample span (shared memory)
  const float* window;  // n taps (global memory: a warp reads consecutive taps, L1-resident)
  int hop, nf;          // frames 2 seq and 2 seq + 1 start at (2 seq) * hop, (2 seq + 1) * hop; frames >= nf are zero
  __device__ __forceinline__ float2 operator()(int seq, int e) const {
    const float w = __ldg(window + e);
    const float* x = xs + 2 * seq * hop + e;
    return make_float2(2 * seq < nf ? x[0] * w : 0.0f, 2 * seq + 1 < nf ? x[hop] * w : 0.0f);
  }
};

template <int R, class Src>
__device__ __forceinline__ void stockham_pass_fixed(const Src src, float2* __restrict__ dst,
                                                    const float2* __restrict__ tw, int n, int Ns, int count) {
  const int nb = n / R;            // butterflies per sequence
  const unsigned nb_magic = div_magic(nb);
  const bool ns_pow2 = (Ns & (Ns - 1)) == 0;
  const int total = count * nb;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int seq = nb > 1 ? fast_div(idx, nb, nb_magic) : idx;
    const int j = idx - seq * nb;
    const int k = ns_pow2 ? (j & (Ns - 1)) : (j % Ns);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src(seq, j + r * nb);
    if (Ns > 1) {  // first pass: k == 0, every twiddle is 1
      // per-pass table [r - 1][k] = W_(Ns R)^(k r): the lanes of a warp (consecutive k) read consecutive entries.  Indexed as
      // W_n^(k tstep r) in one table of n roots, a warp's loads were strided by tstep * r entries: the twiddle loads were the
      // bank-conflict replays of the FFT passes (23 % of the forward kernel's shared-memory wavefronts there, 58 % of the inverse's)
      const float2* t = tw + k;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = regs::cmul(v[r], t[(r - 1) * Ns]);  // packed f32x2: two instructions instead of four
    }
    regs::Dft<R>::run(v);
    float2* d = dst + (size_t)seq * seq_pitch(n);
    const int o0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) d[skew(o0 + r * Ns)] = v[r];
  }
}

// arbitrary (prime) radix: O(R^2) per butterfly, twiddles from the same table
__device__ __forceinline__ void stockham_pass_any(const float2* __restrict__ src, float2* __restrict__ dst,
                                                  const float2* __restrict__ tw, int n, int Ns, int count, int R) {
  const int nb = n / R;
  const int tstep = n / (Ns * R);
  const int rstep = n / R;
  for (int idx = threadIdx.x; idx < count * nb; idx += blockDim.x) {
    const int seq = idx / nb, j = idx - seq * nb;
    const int k = j % Ns;
    const float2* sp = src + (size_t)seq * seq_pitch(n);
    float2 v[kMaxGenericRadix];
    for (int r = 0; r < R; ++r) {
      float2 x = sp[skew(j + r * nb)];
      if (r > 0) x = cmul(x, __ldg(tw + k * r * tstep));  // k*r < Ns*R  =>  index < n
      v[r] = x;
    }
    float2* d = dst + (size_t)seq * seq_pitch(n);
    const int o0 = (j - k) * R + k;
    for (int q = 0; q < R; ++q) {
      float2 acc = v[0];
      for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], __ldg(tw + ((q * r) % R) * rstep)));
      d[skew(o0 + q * Ns)] = acc;
    }
  }
}

// Runs all passes; returns the buffer holding the result.  All threads must call.  `tw` lives in shared memory.
// `first_pass_done`: pass 0 has already been run into `a` (fft_first_pass_from_frames).
__device__ float2* run_fft(const FftDesc& fd, float2* a, float2* b, const float2* tw, int count, bool first_pass_done = false) {
  int Ns = first_pass_done ? fd.radix[0] : 1;
  float2 *src_p = a, *dst = b;
  for (int s = first_pass_done ? 1 : 0; s < fd.nstages; ++s) {
    const int R = fd.radix[s];
    const float2* tws = tw + (Ns - 1);  // pass tables back to back: sum over earlier passes of (R - 1) * Ns = Ns - 1
    const SmemSrc src{src_p, seq_pitch(fd.n)};
    switch (R) {
      case 2: stockham_pass_fixed<2>(src, dst, tws, fd.n, Ns, count); break;
      case 3: stockham_pass_fixed<3>(src, dst, tws, fd.n, Ns, count); break;
      case 4: stockham_pass_fixed<4>(src, dst, tws, fd.n, Ns, count); break;
      case 5: stockham_pass_fixed<5>(src, dst, tws, fd.n, Ns, count); break;
      case 6: stockham_pass_fixed<6>(src, dst, tws, fd.n, Ns, count); break;
      case 8: stockham_pass_fixed<8>(src, dst, tws, fd.n, Ns, count); break;
      case 10: stockham_pass_fixed<10>(src, dst, tws, fd.n, Ns, count); break;
      case 12: stockham_pass_fixed<12>(src, dst, tws, fd.n, Ns, count); break;
      case 15: stockham_pass_fixed<15>(src, dst, tws, fd.n, Ns, count); break;
      case 16: stockham_pass_fixed<16>(src, dst, tws, fd.n, Ns, count); break;
      case 20: stockham_pass_fixed<20>(src, dst, tws, fd.n, Ns, count); break;
      case 25: stockham_pass_fixed<25>(src, dst, tws, fd.n, Ns, count); break;
      case 32: stockham_pass_fixed<32>(src, dst, tws, fd.n, Ns, count); break;
      default: stockham_pass_any(src_p, dst, fd.tw_plain, fd.n, Ns, count, R); break;
    }
    __syncthreads();
    Ns *= R;
    float2* t = src_p;
    src_p = dst;
    dst = t;
  }
  return src_p;
}

// First pass of the forward transform straight from the sample span (power-of-two first radices 8 / 16 / 32: what `factorize`
// puts first for every n_fft with a factor 8).  Returns false when the plan's first radix has no fused instance.
__device__ __forceinline__ bool fft_first_pass_fusable(const FftDesc& fd) {
  return fd.nstages >= 1 && (fd.radix[0] == 8 || fd.radix[0] == 16 || fd.radix[0] == 32);
}
__device__ __forceinline__ void fft_first_pass_from_frames(const FftDesc& fd, const FrameSrc& src, float2* a, int count) {
  switch (fd.radix[0]) {
    case 8: stockham_pass_fixed<8>(src, a, nullptr, fd.n, 1, count); break;
    case 16: stockham_pass_fixed<16>(src, a, nullptr, fd.n, 1, count); break;
    default: stockham_pass_fixed<32>(src, a, nullptr, fd.n, 1, count); break;
  }
  __syncthreads();
}

// the twiddle table into shared memory, once per CTA (n float2 behind the kernel's other buffers)
__device__ __forceinline__ void stage_twiddles(float2* tw_s, const float2* __restrict__ tw_g, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw_s[i] = tw_g[i];
  __syncthreads();
}

// Fire-and-forget float max (no read-back, so the issuing warp never waits on an HBM round trip):
// non-negative floats order like signed ints, negative floats order inversely as unsigned ints.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

struct FwdParams {
  const float* audio;
  int64_t clip_stride, valid_length, sample_offset, frame_begin, frame_count;
  float pad_value;
  int batch;
  Geometry geo;
  int n_fft, hop, n_freqs, pad_mode;
  float preemph;
  int spec_kind;
  float spec_eps;
  int n_mels, log_kind, guard_kind;
  float guard_eps;
  int apply_affine;
  float affine_add, affine_div;
  int out_layout;
  void* out;
  int64_t out_clip_stride;
  float *clip_max, *tile_min;
  double* feat_sums;
  const float2* tw;
  const float* window;
  const int *mel_start, *mel_len, *mel_off;
  const float* mel_w;
  int mel_nnz, mel_smem;  // filterbank taps; 1: the CSR tables are staged in shared memory behind the twiddles
  FftDesc fft;
  int frames_per_tile, tiles_per_clip;
  int dump_frames, dump_windowed;
  // Kaldi per-frame pre-processing (compute_fbank_kaldi, dsp.py:619-632)
  int frame_len, frame_dc;
  float frame_preemph, dither;
  unsigned long long seed;
};

// global sample (after right-pad + preemphasis) at source index s >= 0
__device__ __forceinline__ float fetch_sample(const FwdParams& p, const float* clip, int64_t s) {
  float x = s < p.valid_length ? __ldg(clip + (s - p.sample_offset)) : p.pad_value;
  if (p.preemph != 0.0f && s > 0) {
    const int64_t sm = s - 1;
    const float xm = sm < p.valid_length ? __ldg(clip + (sm - p.sample_offset)) : p.pad_value;
    x = __fsub_rn(x, __fmul_rn(p.preemph, xm));  // two roundings, as `x[1:] - a*x[:-1]` (no FMA contraction)
  }
  return x;
}

__global__ void __launch_bounds__(kThreads, 1) frontend_generic_kernel(const FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft, hop = p.hop, F = p.n_freqs, FT = p.frames_per_tile, PAIRS = FT / 2;
  const int span = (FT - 1) * hop + N;
  const int NP = seq_pitch(N);  // skewed sequence pitch of the FFT buffers
  float2* bufA = reinterpret_cast<float2*>(smem_raw);
  float2* bufB = bufA + (size_t)PAIRS * NP;
  float* xs = reinterpret_cast<float*>(bufB + (size_t)PAIRS * NP);
  float2* tw_s = reinterpret_cast<float2*>(xs + ((span + 3) & ~3));  // [N] twiddles, staged once per CTA
  __shared__ float red_max[kThreads / 32], red_min[kThreads / 32];
  // the CSR filterbank (taps, then start / length / offset per row) once per CTA: read from global memory, the taps put an L1
  // round trip into every step of the projection's FMA chains (long-scoreboard stalls were half of that phase's samples)
  float* melw_s = reinterpret_cast<float*>(tw_s + N);
  int* melm_s = reinterpret_cast<int*>(melw_s + ((p.mel_nnz + 3) & ~3));
  if (p.mel_smem) {
    for (int i = threadIdx.x; i < p.mel_nnz; i += blockDim.x) melw_s[i] = p.mel_w[i];
    for (int i = threadIdx.x; i < p.n_mels; i += blockDim.x) {
      melm_s[3 * i + 0] = p.mel_start[i];
      melm_s[3 * i + 1] = p.mel_len[i];
      melm_s[3 * i + 2] = p.mel_off[i];
    }
  }
  if (!p.dump_frames) stage_twiddles(tw_s, p.tw, N);
  else __syncthreads();
  // index splits of the element-wise loops as multiply-high (operands well below 2^20 / 2^12, see fast_div)
  const unsigned mg_N = div_magic(N), mg_M = div_magic(p.n_mels > 0 ? p.n_mels : F);

  const int64_t total_tiles = (int64_t)p.batch * p.tiles_per_clip;
  // (clip, tile) of the CTA's tiles are walked incrementally: a 64-bit division per tile and thread (two with the prefetch of
  // the next tile) was 6 % of the instructions at 8 frames per tile
  const unsigned tpc = (unsigned)p.tiles_per_clip;
  const int step_c = (int)(gridDim.x / tpc), step_t = (int)(gridDim.x - (unsigned)step_c * tpc);
  int clip_i = (int)(blockIdx.x / tpc), tile_i = (int)(blockIdx.x - (unsigned)clip_i * tpc);
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, clip_i += step_c, tile_i += step_t) {
    if (tile_i >= (int)tpc) {
      tile_i -= (int)tpc;
      ++clip_i;
    }
    const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
    const int64_t lt0 = (int64_t)tile_i * FT;  // local frame index of the tile's first frame
    const int64_t t0 = p.frame_begin + lt0;    // global frame index
    const int nf = (int)min((int64_t)FT, p.frame_count - lt0);  // valid frames in this tile

    // ---- stage the sample span (padding / right-pad / pre-emphasis resolved here) -------------------
    const int64_t q0 = t0 * hop;
    const int need = (nf - 1) * hop + N;
    // INTERIOR tile (CTA-uniform): every needed sample is a sample of the signal this rank holds — no padding, no reflection,
    // no virtual right pad — so the span is one linear copy (16-byte loads where the source is aligned).  The general loop
    // below resolves padding per sample (64-bit index arithmetic and three range tests around every 4-byte load: on the ncu
    // source page it was 26 % of the S3Gen-mel kernel's time, waiting on its own loads).
    const int64_t s0 = q0 - p.geo.pad_left;
    const int64_t s_lim = p.geo.length < p.valid_length ? p.geo.length : p.valid_length;
    const bool interior = s0 >= (p.preemph != 0.0f ? 1 : 0) && s0 - (p.preemph != 0.0f ? 1 : 0) >= p.sample_offset && s0 + need <= s_lim;
    if (interior) {
      const float* src = clip + (s0 - p.sample_offset);
      if (p.preemph != 0.0f) {
        for (int i = threadIdx.x; i < need; i += blockDim.x) xs[i] = __fsub_rn(__ldg(src + i), __fmul_rn(p.preemph, __ldg(src + i - 1)));
      } else if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* x4 = reinterpret_cast<float4*>(xs);
        const int n4 = need >> 2;
#pragma unroll 4
        for (int i = threadIdx.x; i < n4; i += blockDim.x) x4[i] = __ldg(s4 + i);
        for (int i = 4 * n4 + threadIdx.x; i < need; i += blockDim.x) xs[i] = __ldg(src + i);
      } else {
#pragma unroll 4
        for (int i = threadIdx.x; i < need; i += blockDim.x) xs[i] = __ldg(src + i);
      }
      for (int i = need + threadIdx.x; i < span; i += blockDim.x) xs[i] = 0.0f;
    } else {
      for (int i = threadIdx.x; i < span; i += blockDim.x) {
        float v = 0.0f;
        if (i < need) {
          const int64_t s = source_index(p.geo, p.pad_mode, q0 + i);
          if (s >= 0) v = fetch_sample(p, clip, s);
        }
        xs[i] = v;
      }
    }
    {  // the next tile's samples are pulled into L2 while this tile is transformed (one 128-byte line per thread)
      const int64_t ntile = tile + gridDim.x;
      if (ntile < total_tiles) {
        int nclip = clip_i + step_c, nti = tile_i + step_t;
        if (nti >= (int)tpc) {
          nti -= (int)tpc;
          ++nclip;
        }
        int64_t a0 = (p.frame_begin + (int64_t)nti * FT) * hop - p.geo.pad_left, a1 = a0 + span;
        a0 = a0 < p.sample_offset ? p.sample_offset : a0;
        a1 = a1 > s_lim ? s_lim : a1;
        const float* nsrc = p.audio + (int64_t)nclip * p.clip_stride - p.sample_offset;
        const uintptr_t l0 = reinterpret_cast<uintptr_t>(nsrc + a0) & ~(uintptr_t)127;
        if (a1 > a0)
          for (uintptr_t q = l0 + 128u * threadIdx.x; q < reinterpret_cast<uintptr_t>(nsrc + a1); q += 128u * blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
      }
    }
    __syncthreads();

    if (p.dump_frames) {  // parity hook: the framed (optionally windowed) matrix itself
      float* o = reinterpret_cast<float*>(p.out) + (int64_t)clip_i * p.out_clip_stride + lt0 * N;
      for (int i = threadIdx.x; i < nf * N; i += blockDim.x) {
        const int f = i / N, k = i - f * N;
        float v = xs[f * hop + k];
        if (p.dump_windowed) v *= p.window[k];
        o[i] = v;
      }
      __syncthreads();
      continue;
    }

    // ---- window + pack two real frames per complex sequence ------------------------------------------
    float2* Z;
    if (p.frame_dc || p.frame_preemph != 0.0f || p.dither != 0.0f) {
      // Kaldi front-end (dsp.py:619-656): per FRAME, on its first W = frame_len samples: + dither * N(0,1)
      // (independent per frame element), - the frame's own mean, pre-emphasis within the frame (sample 0 kept),
      // then the window (zero beyond W).  Raw frames go to bufA, the finished ones to bufB.
      const int W = p.frame_len > 0 ? p.frame_len : N;
      for (int i = threadIdx.x; i < PAIRS * N; i += blockDim.x) {
        const int pr = N > 1 ? fast_div(i, N, mg_N) : i, k = i - pr * N;
        const int fa = 2 * pr, fb = 2 * pr + 1;
        float a = (fa < nf && k < W) ? xs[fa * hop + k] : 0.0f;
        float b = (fb < nf && k < W) ? xs[fb * hop + k] : 0.0f;
        if (p.dither != 0.0f && k < W) {
          curandStatePhilox4_32_10_t rs;  // counter-based: (seed, frame pair, sample) -> two normals, O(1) set-up
          curand_init(p.seed, (unsigned long long)clip_i * (unsigned long long)((p.geo.num_frames + 1) / 2 + 1) +
                                  (unsigned long long)((t0 >> 1) + pr), (unsigned long long)k, &rs);
          const float2 g = curand_normal2(&rs);
          a += p.dither * g.x;
          b += p.dither * g.y;
        }
        bufA[(size_t)pr * NP + skew(k)] = make_float2(a, b);
      }
      __syncthreads();
      float2* means = reinterpret_cast<float2*>(xs);  // the sample span is consumed: reuse it for the frame means
      if (p.frame_dc) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int pr = warp; pr < PAIRS; pr += nw) {
          float sa = 0.0f, sb = 0.0f;
          for (int k = lane; k < W; k += 32) {
            const float2 v = bufA[(size_t)pr * NP + skew(k)];
            sa += v.x;
            sb += v.y;
          }
          for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
          }
          if (lane == 0) means[pr] = make_float2(sa / (float)W, sb / (float)W);
        }
      }
      __syncthreads();
      const float pe = p.frame_preemph;
      for (int i = threadIdx.x; i < PAIRS * N; i += blockDim.x) {
        const int pr = N > 1 ? fast_div(i, N, mg_N) : i, k = i - pr * N;
        float2 v = make_float2(0.0f, 0.0f);
        if (k < W) {
          const float2 m = p.frame_dc ? means[pr] : make_float2(0.0f, 0.0f);
          const float2 x = bufA[(size_t)pr * NP + skew(k)];
          v = make_float2(x.x - m.x, x.y - m.y);
          if (pe != 0.0f && k > 0) {  // separately rounded multiply and subtract, as the reference's array expression
            const float2 xp = bufA[(size_t)pr * NP + skew(k - 1)];
            v.x = __fsub_rn(v.x, __fmul_rn(pe, xp.x - m.x));
            v.y = __fsub_rn(v.y, __fmul_rn(pe, xp.y - m.y));
          }
          const float w = p.window[k];
          v.x *= w;
          v.y *= w;
        }
        bufB[(size_t)pr * NP + skew(k)] = v;
      }
      __syncthreads();
      Z = run_fft(p.fft, bufB, bufA, tw_s, PAIRS);
    } else {
      if (fft_first_pass_fusable(p.fft)) {
        fft_first_pass_from_frames(p.fft, FrameSrc{xs, p.window, hop, nf}, bufA, PAIRS);
        Z = run_fft(p.fft, bufA, bufB, tw_s, PAIRS, true);
      } else {
      // a thread owns sample positions k and walks the tile's frame pairs: one window tap and one skewed offset per k
      for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const float w = __ldg(p.window + k);
        const float* x0 = xs + k;
        float2* d = bufA + skew(k);
        const int full = nf >> 1;  // pairs whose two frames both exist
#pragma unroll 4
        for (int pr = 0; pr < full; ++pr, x0 += 2 * hop, d += NP) *d = make_float2(x0[0] * w, x0[hop] * w);
        for (int pr = full; pr < PAIRS; ++pr, x0 += 2 * hop, d += NP) *d = make_float2(2 * pr < nf ? x0[0] * w : 0.0f, 0.0f);
      }
      __syncthreads();
      Z = run_fft(p.fft, bufA, bufB, tw_s, PAIRS);
      }
    }
    float2* other = (Z == bufA) ? bufB : bufA;

    // ---- Hermitian separation ------------------------------------------------------------------------
    // A thread owns BINS and walks the tile's frame pairs: Z[k] and Z[N-k] are read once for both frames of a pair, the skewed
    // offsets are computed once per bin, and no element index is split by a division.
    if (p.n_mels == 0 && p.spec_kind == B2A_SPEC_COMPLEX) {
      float2* o = reinterpret_cast<float2*>(p.out) + (int64_t)clip_i * p.out_clip_stride + lt0 * F;
      for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float2* zk_p = Z + skew(k);
        const float2* zm_p = Z + (k == 0 ? 0 : skew(N - k));
        for (int pr = 0; pr < PAIRS; ++pr) {
          const int fa = 2 * pr, fb = 2 * pr + 1;
          if (fa >= nf) break;
          const float2 zk = zk_p[(size_t)pr * NP], zm = zm_p[(size_t)pr * NP];
          o[(int64_t)fa * F + k] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
          if (fb < nf) o[(int64_t)fb * F + k] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
        }
      }
      __syncthreads();
      continue;
    }
    float* P = reinterpret_cast<float*>(other);  // [FT][F]
    // magnitude: ONE MUFU instruction (sqrt.approx, <= 2 ulp, as the fused kernels) instead of the IEEE sqrtf sequence with its
    // denormal slow path behind a branch per bin
    const float spec_add = p.spec_kind == B2A_SPEC_MAGNITUDE ? 0.0f : p.spec_eps;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      const float2* zk_p = Z + skew(k);
      const float2* zm_p = Z + (k == 0 ? 0 : skew(N - k));
      float* pk = P + k;
      for (int pr = 0; pr < PAIRS; ++pr) {
        const int fa = 2 * pr, fb = 2 * pr + 1;
        const float2 zk = zk_p[(size_t)pr * NP], zm = zm_p[(size_t)pr * NP];
        float va = 0.0f, vb = 0.0f;
        if (fa < nf) {
          const float re = 0.5f * (zk.x + zm.x), im = 0.5f * (zk.y - zm.y);
          const float pw = re * re + im * im;
          va = p.spec_kind == B2A_SPEC_POWER ? pw : sqrt_approx(pw + spec_add);
        }
        if (fb < nf) {
          const float re = 0.5f * (zk.y + zm.y), im = 0.5f * (zm.x - zk.x);
          const float pw = re * re + im * im;
          vb = p.spec_kind == B2A_SPEC_POWER ? pw : sqrt_approx(pw + spec_add);
        }
        pk[(size_t)fa * F] = va;
        pk[(size_t)fb * F] = vb;
      }
    }
    __syncthreads();

    // ---- mel projection (banded), guard, log, affine; staged in smem ---------------------------------
    const int M = p.n_mels > 0 ? p.n_mels : F;
    float* Y = reinterpret_cast<float*>(Z);  // [FT][M]   (M <= F <= N)
    float lmax = -INFINITY, lmin = INFINITY;
    // One item = (mel row, group of 8 frames), owned by a QUAD of lanes: the four lanes take every fourth tap, each tap is
    // loaded once and used for the 8 frames' accumulators (9 loads per 8 multiply-adds instead of 16), the quad folds with two
    // shuffles per frame and each lane finishes two of the frames (guard, logarithm, affine).  What matters at 8 frames per tile
    // (the 1920-point shapes): every thread gets the same mix of short and long rows — with one thread per (row, frame) item the
    // 640 items of an S3Gen tile left three quarters of the CTA idle while 128 threads walked the 60-tap rows after their short
    // ones — and the projection costs ~2 instructions per multiply-add instead of ~6 (ncu source page: the phase was 12 % of the
    // kernel's time plus most of the 9 % spent at the barrier behind it).
    const int G8 = (FT + 7) >> 3, items = M * G8, items_pad = (items + 7) & ~7;  // warp-uniform trip count: the shuffles need every lane
    const int quad = threadIdx.x >> 2, sub = threadIdx.x & 3;
    for (int i = quad; i < items_pad; i += blockDim.x >> 2) {
      const bool on = i < items;
      const int g = on ? (M > 1 ? fast_div(i, M, mg_M) : i) : 0, m = on ? i - g * M : 0;
      const int f0 = 8 * g;
      float a8[8];
#pragma unroll
      for (int f = 0; f < 8; ++f) a8[f] = 0.0f;
      if (on) {
        int fo[8];  // row offsets of the group's frames; frames past the tile repeat its last row (never stored)
#pragma unroll
        for (int f = 0; f < 8; ++f) fo[f] = min(f0 + f, FT - 1) * F;
        if (p.n_mels > 0) {
          int start, len, off;
          const float* w;
          if (p.mel_smem) {
            start = melm_s[3 * m], len = melm_s[3 * m + 1], off = melm_s[3 * m + 2];
            w = melw_s + off;
          } else {
            start = __ldg(p.mel_start + m), len = __ldg(p.mel_len + m), off = __ldg(p.mel_off + m);
            w = p.mel_w + off;
          }
          const float* row = P + start;
          for (int j = sub; j < len; j += 4) {
            const float wv = w[j];
#pragma unroll
            for (int f = 0; f < 8; ++f) a8[f] = fmaf(row[fo[f] + j], wv, a8[f]);
          }
        } else if (sub == 0) {
#pragma unroll
          for (int f = 0; f < 8; ++f) a8[f] = P[fo[f] + m];
        }
      }
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        a8[f] += __shfl_xor_sync(0xffffffffu, a8[f], 1);
        a8[f] += __shfl_xor_sync(0xffffffffu, a8[f], 2);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // lane `sub` finishes frames f0 + 2 sub + h
        const int f = f0 + 2 * sub + h;
        float acc = sub == 0 ? a8[h] : sub == 1 ? a8[2 + h] : sub == 2 ? a8[4 + h] : a8[6 + h];
        if (!on || f >= nf) continue;
        if (p.guard_kind == B2A_GUARD_MAX) acc = fmaxf(acc, p.guard_eps);
        else if (p.guard_kind == B2A_GUARD_ADD) acc = acc + p.guard_eps;
        if (p.log_kind == B2A_LOG_LOG10) acc = log10f(acc);
        else if (p.log_kind == B2A_LOG_LN) acc = logf(acc);
        lmax = fmaxf(lmax, acc);
        lmin = fminf(lmin, acc);
        if (p.apply_affine) acc = (acc + p.affine_add) / p.affine_div;
        Y[f * M + m] = acc;
      }
    }
    if (p.clip_max) {
      for (int o = 16; o > 0; o >>= 1) {
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      }
      if ((threadIdx.x & 31) == 0) { red_max[threadIdx.x >> 5] = lmax; red_min[threadIdx.x >> 5] = lmin; }
    }
    __syncthreads();
    if (p.clip_max && threadIdx.x == 0) {
      float a = red_max[0], b = red_min[0];
      for (int w = 1; w < kThreads / 32; ++w) { a = fmaxf(a, red_max[w]); b = fminf(b, red_min[w]); }
      atomic_max_float(p.clip_max + clip_i, a);
      p.tile_min[(int64_t)clip_i * p.tiles_per_clip + tile_i] = b;
    }
    if (p.feat_sums) {
      for (int m = threadIdx.x; m < M; m += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int f = 0; f < nf; ++f) { const double y = Y[f * M + m]; s1 += y; s2 += y * y; }
        atomicAdd(p.feat_sums + ((int64_t)clip_i * M + m) * 2 + 0, s1);
        atomicAdd(p.feat_sums + ((int64_t)clip_i * M + m) * 2 + 1, s2);
      }
    }
    // ---- coalesced write-out ---------------------------------------------------------------------------
    float* o = reinterpret_cast<float*>(p.out) + (int64_t)clip_i * p.out_clip_stride;
    if (p.out_layout == B2A_LAYOUT_TM) {
      float* ot = o + lt0 * M;
      for (int i = threadIdx.x; i < nf * M; i += blockDim.x) ot[i] = Y[i];
    } else {
      // (M, T): a warp takes 32 / fp feature rows at a time, fp = the tile's frame count rounded up to a power of two —
      // no per-element division by the run-time frame count (a 40-instruction sequence per stored value)
      const int fp = nf <= 1 ? 1 : nf <= 2 ? 2 : nf <= 4 ? 4 : nf <= 8 ? 8 : nf <= 16 ? 16 : 32;
      const int sh = __ffs(fp) - 1, lane = threadIdx.x & 31, rpw = 32 >> sh;
      const int fl = lane & (fp - 1), rl = lane >> sh;
      for (int m = (threadIdx.x >> 5) * rpw + rl; m < M; m += (blockDim.x >> 5) * rpw)
        for (int f = fl; f < nf; f += fp) o[(int64_t)m * p.frame_count + lt0 + f] = Y[f * M + m];
    }
    __syncthreads();
  }
}

// ---- finalize: clamp / normalise in place --------------------------------------------------------------
struct FinParams {
  float* out;
  int64_t out_clip_stride;
  int batch, n_mels, out_layout;
  int64_t frames;         // local frames held in out
  int64_t global_frames;  // frames the statistics cover
  int clamp_kind;
  float clamp_value;
  int apply_affine;
  float affine_add, affine_div;
  int norm_kind, norm_ddof;
  float norm_eps;
  float* clip_max;
  const float* tile_min;
  int tile_frames, tiles_per_clip;
  const double* feat_sums;
  int stats_affine;  // 1: clip_max / tile_min were recorded AFTER the affine map (fast kernels)
  int tile_min_pitch;  // tiles per clip in the tile_min table (tiles_per_clip may be cut short: constant padding rows)
  int out_dtype;     // B2A_DTYPE_*: 16-bit features (fast 400/160 kernels, (T, M) layout): the floor is cast the same way
  int fill_unwritten;  // 1: tiles with tile_min == -inf were NOT stored by the fused kernel (all-silent: every value is the
  EpilogueConsts ec;   //    epilogue's constant for zero power) — the fix-up writes max(c, floor) there
};

// B2A_CLAMP_BATCH_MAX: one max over the whole batch (s3tokenizer/utils.py:131) -> broadcast into clip_max
__global__ void batch_max_kernel(float* clip_max, int batch) {
  __shared__ float red[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < batch; i += blockDim.x) m = fmaxf(m, clip_max[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    red[0] = m;
  }
  __syncthreads();
  m = red[0];
  for (int i = threadIdx.x; i < batch; i += blockDim.x) clip_max[i] = m;
}

// Clamp fix-up: one WARP per tile; a tile whose recorded minimum is not below the floor is already final
// (the affine map is monotone increasing, so it commutes with max()) and costs one 4-byte read.
__global__ void __launch_bounds__(256) clamp_fixup_kernel(const FinParams p) {
  const int clip_i = blockIdx.y;
  const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= p.tiles_per_clip) return;
  float floor_out, floor_cmp;
  if (p.clamp_kind == B2A_CLAMP_FIXED) {
    floor_out = p.apply_affine ? (p.clamp_value + p.affine_add) / p.affine_div : p.clamp_value;
    floor_cmp = p.stats_affine ? floor_out : p.clamp_value;
  } else {
    const float mx = p.clip_max[clip_i];
    if (p.stats_affine) {
      floor_out = p.apply_affine ? mx - p.clamp_value / p.affine_div : mx - p.clamp_value;
      floor_cmp = floor_out;
    } else {
      floor_cmp = mx - p.clamp_value;
      floor_out = p.apply_affine ? (floor_cmp + p.affine_add) / p.affine_div : floor_cmp;
    }
  }
  const float tmin = p.tile_min[(int64_t)clip_i * p.tile_min_pitch + tile];
  const bool unwritten = p.fill_unwritten && tmin == -INFINITY;
  if (!unwritten && !(tmin < floor_cmp)) return;
  float* o = p.out + (int64_t)clip_i * p.out_clip_stride;
  const int M = p.n_mels;
  const int64_t f0 = (int64_t)tile * p.tile_frames;
  const int nf = (int)min((int64_t)p.tile_frames, p.frames - f0);
  if (p.out_dtype == B2A_DTYPE_F32 && p.out_layout == B2A_LAYOUT_TM && (reinterpret_cast<uintptr_t>(o + f0 * M) & 15) == 0 && (M & 3) == 0) {
    // float32 (T, M) rows of a tile are one contiguous, 16-byte aligned block: float4 sweeps
    float4* t4 = reinterpret_cast<float4*>(o + f0 * M);
    const int n4 = nf * M / 4;
    if (unwritten) {  // the tile was never stored: every element is max(c, floor) — one write pass, no read
      const float v = fmaxf(epilogue_of_zero(p.ec), floor_out);
      const float4 v4 = make_float4(v, v, v, v);
      for (int i = lane; i < n4; i += 32) t4[i] = v4;
    } else {
      for (int i = lane; i < n4; i += 32) {
        float4 v = t4[i];
        if (v.x < floor_out || v.y < floor_out || v.z < floor_out || v.w < floor_out) {  // (a NaN stays a NaN)
          v.x = v.x < floor_out ? floor_out : v.x;
          v.y = v.y < floor_out ? floor_out : v.y;
          v.z = v.z < floor_out ? floor_out : v.z;
          v.w = v.w < floor_out ? floor_out : v.w;
          t4[i] = v;
        }
      }
    }
    return;
  }
  if (p.out_dtype == B2A_DTYPE_F16) {  // cast(max(y, floor)) == max(cast(y), cast(floor)): the cast is monotone
    __half* t = reinterpret_cast<__half*>(p.out) + (int64_t)clip_i * p.out_clip_stride + f0 * M;
    const __half fl = __float2half_rn(floor_out);
    for (int i = lane; i < nf * M; i += 32)
      if (__hlt(t[i], fl)) t[i] = fl;
  } else if (p.out_dtype == B2A_DTYPE_BF16) {
    __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)clip_i * p.out_clip_stride + f0 * M;
    const __nv_bfloat16 fl = __float2bfloat16_rn(floor_out);
    for (int i = lane; i < nf * M; i += 32)
      if (__hlt(t[i], fl)) t[i] = fl;
  } else if (p.out_layout == B2A_LAYOUT_TM) {
    float* t = o + f0 * M;
    const float cfill = unwritten ? fmaxf(epilogue_of_zero(p.ec), floor_out) : 0.0f;
    for (int i = lane; i < nf * M; i += 32) {
      const float v = unwritten ? cfill : t[i];
      if (unwritten || v < floor_out) t[i] = unwritten ? cfill : floor_out;
    }
  } else {
    for (int i = lane; i < nf * M; i += 32) {
      const int m = i / nf, f = i - m * nf;
      float* q = o + (int64_t)m * p.frames + f0 + f;
      if (*q < floor_out) *q = floor_out;
    }
  }
}

// (x - mean) / (std + eps), in place.  The per-mel (or global) mean and denominator are derived ONCE per block
// from the float64 sums into shared memory; the sweep itself is float4 loads / stores (HBM bound: it re-reads
// and re-writes the features, the only part of the path that touches the output twice).
__global__ void __launch_bounds__(256) normalise_kernel(const FinParams p) {
  __shared__ float s_mean[256], s_den[256];
  const int clip_i = blockIdx.y;
  const int M = p.n_mels;
  float* o = p.out + (int64_t)clip_i * p.out_clip_stride;
  const int64_t total = p.frames * M;
  const double n = (double)p.global_frames;
  const double* S = p.feat_sums + (int64_t)clip_i * M * 2;
  if (p.norm_kind == B2A_NORM_GLOBAL) {
    if (threadIdx.x == 0) {
      double s1 = 0, s2 = 0;
      for (int m = 0; m < M; ++m) { s1 += S[2 * m]; s2 += S[2 * m + 1]; }
      const double cnt = n * M, mean = s1 / cnt;
      double var = (s2 - cnt * mean * mean) / (cnt - p.norm_ddof);
      if (var < 0) var = 0;
      s_mean[0] = (float)mean;
      s_den[0] = (float)sqrt(var) + p.norm_eps;
    }
    __syncthreads();
    const float g_mean = s_mean[0], g_den = s_den[0];
    __syncthreads();
    for (int m = threadIdx.x; m < M && m < 256; m += blockDim.x) {
      s_mean[m] = g_mean;
      s_den[m] = g_den;
    }
  } else {
    for (int m = threadIdx.x; m < M && m < 256; m += blockDim.x) {
      const double mu = S[2 * m] / n;
      double var = (S[2 * m + 1] - n * mu * mu) / (n - p.norm_ddof);
      if (var < 0) var = 0;
      s_mean[m] = (float)mu;
      s_den[m] = (float)sqrt(var) + p.norm_eps;
    }
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = M <= 256 && (M & 3) == 0 && p.out_layout == B2A_LAYOUT_TM && (reinterpret_cast<uintptr_t>(o) & 15) == 0;
  if (vec) {
    float4* o4 = reinterpret_cast<float4*>(o);
    const int M4 = M >> 2;
    if (stride % M4 == 0) {
      // the launch makes the grid's stride a multiple of the row length (in quads): a thread meets the SAME four features in
      // every iteration, so its means and reciprocal denominators live in registers and the sweep is load, 4 FADD, 4 FMUL,
      // store.  (Round 1: a 64-bit modulo, eight 4-way bank-conflicted scalar LDS and four IEEE divisions per quad — 70 % of the
      // HBM peak with 68 % of its shared-memory wavefronts conflicts, profiles/r02_normalise_ncu_full.json.)
      const int m = (int)(first % M4) << 2;
      const float4 mu = make_float4(s_mean[m], s_mean[m + 1], s_mean[m + 2], s_mean[m + 3]);
      const float4 rd = make_float4(__frcp_rn(s_den[m]), __frcp_rn(s_den[m + 1]), __frcp_rn(s_den[m + 2]), __frcp_rn(s_den[m + 3]));
      for (int64_t i = first; i < (total >> 2); i += stride) {
        float4 v = o4[i];
        v.x = (v.x - mu.x) * rd.x;
        v.y = (v.y - mu.y) * rd.y;
        v.z = (v.z - mu.z) * rd.z;
        v.w = (v.w - mu.w) * rd.w;
        o4[i] = v;
      }
      return;
    }
    for (int64_t i = first; i < (total >> 2); i += stride) {
      const int m = (int)(i % M4) << 2;
      float4 v = o4[i];
      v.x = __fdiv_rn(v.x - s_mean[m], s_den[m]);
      v.y = __fdiv_rn(v.y - s_mean[m + 1], s_den[m + 1]);
      v.z = __fdiv_rn(v.z - s_mean[m + 2], s_den[m + 2]);
      v.w = __fdiv_rn(v.w - s_mean[m + 3], s_den[m + 3]);
      o4[i] = v;
    }
    return;
  }
  for (int64_t i = first; i < total; i += stride) {
    const int m = p.out_layout == B2A_LAYOUT_TM ? (int)(i % M) : (int)(i / p.frames);
    float mean, den;
    if (m < 256) {
      mean = s_mean[m];
      den = s_den[m];
    } else {  // very wide outputs (n_mels == 0: normalised spectra) — derive per element
      const double mu = p.norm_kind == B2A_NORM_GLOBAL ? (double)s_mean[0] : S[2 * m] / n;
      double var = (S[2 * m + 1] - n * mu * mu) / (n - p.norm_ddof);
      if (var < 0) var = 0;
      mean = (float)mu;
      den = p.norm_kind == B2A_NORM_GLOBAL ? s_den[0] : (float)sqrt(var) + p.norm_eps;
    }
    o[i] = (o[i] - mean) / den;
  }
}

// compute_deltas_kaldi (dsp.py:439-483): a (2n+1)-tap antisymmetric FIR along time, edge samples replicated (or zeros)
__global__ void __launch_bounds__(256) deltas_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows,
                                                     int64_t cols, int n, float denom, int edge) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, t = i - r * cols;
    const float* row = x + r * cols;
    float acc = 0.0f;
    for (int k = -n; k <= n; ++k) {  // ascending tap order, as mx.sum(window * kernel_weights) over the window axis
      int64_t j = t + k;
      float v;
      if (j < 0) v = edge ? row[0] : 0.0f;
      else if (j >= cols) v = edge ? row[cols - 1] : 0.0f;
      else v = row[j];
      acc += v * (float)k;
    }
    out[i] = acc / denom;
  }
}

__global__ void init_stats_kernel(float* clip_max, double* feat_sums, int batch, int n_sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) {
    if (clip_max) clip_max[i] = -INFINITY;
  }
  if (feat_sums && i < n_sums) feat_sums[i] = 0.0;
}

// ---- inverse ---------------------------------------------------------------------------------------------
struct InvParams {
  const float2* spec;     // interleaved, or nullptr
  const float* spec_re;   // planar form
  const float* spec_im;
  int64_t clip_stride, T;
  int batch;
  int n_fft, hop, n_freqs;
  int norm_sq, div_clamp;
  int64_t out_start, out_len, out_clip_stride;
  float* out;
  const float2* tw;
  const float* window;
  FftDesc fft;
  int frames_adv;       // frames advanced per tile (tile covers frames_adv*hop output samples)
  int frames_cap;       // max frames resident per tile
  int pairs_chunk;      // pairs transformed per FFT round
  int planes_vec2;      // both planes start on an 8-byte boundary: pairs of frames may be read with one 8-byte load
  int tiles_per_clip;
  PolarSpec polar;
  float div_eps;
};

// A CTA owns a contiguous range of S = frames_adv * hop OUTPUT samples and keeps their overlap-add sums in shared memory.
// It inverse-transforms the frames that touch the range, a few pairs per round (two Hermitian spectra per complex FFT), in
// ASCENDING frame order, and after every round each thread adds the round's windowed samples into the sums it owns — frame by
// frame, so every output sample accumulates its frames in exactly the order of the reference's scatter-add (dsp.py:193-204).
// Round 1 kept every windowed frame of the tile in shared memory instead (N floats per frame): with N = 1920 that left room for
// 4 new frames per tile next to the 5 it had to recompute from its neighbours; the sums need hop floats per frame.
__global__ void __launch_bounds__(kThreads, 1) istft_generic_kernel(const InvParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft, hop = p.hop, F = p.n_freqs;
  const int NP = seq_pitch(N);  // skewed sequence pitch of the FFT buffers
  float2* bufA = reinterpret_cast<float2*>(smem_raw);
  float2* bufB = bufA + (size_t)p.pairs_chunk * NP;
  float2* tw_s = bufB + (size_t)p.pairs_chunk * NP;         // [N] twiddles
  float* win_s = reinterpret_cast<float*>(tw_s + N);        // [N] window
  float* acc = win_s + N;                                    // [frames_adv * hop] overlap-add sums of the tile
  const float invN = 1.0f / (float)N;
  const int S = p.frames_adv * hop;
  for (int i = threadIdx.x; i < N; i += blockDim.x) win_s[i] = p.window[i];
  stage_twiddles(tw_s, p.tw, N);

  const int64_t total_tiles = (int64_t)p.batch * p.tiles_per_clip;
  // tile-relative sample positions stay below 2^20 and hop below 2^12 for every shape the host hands over with fdiv_ok set:
  // their quotients by the hop are then one multiply-high (fast_div) instead of a 64-bit division sequence per output sample
  const bool fdiv_ok = hop < 4096 && (int64_t)S + N + (int64_t)(2 * p.pairs_chunk + 2) * hop < (1 << 20);
  const unsigned mg_hop = div_magic(hop);
  auto div_hop = [&](int x) -> int { return fdiv_ok ? fast_div(x, hop, mg_hop) : x / hop; };
  // frames touching a tile's OLA range [n0, n1): t*hop <= n1-1  and  t*hop + N - 1 >= n0
  auto tile_frames = [&](int64_t tl, int& clip_o, int64_t& tlo_o, int& nfr_o) {
    clip_o = (int)(tl / p.tiles_per_clip);
    const int ti = (int)(tl - (int64_t)clip_o * p.tiles_per_clip);
    const int64_t a0 = (int64_t)ti * S, a1 = min(a0 + (int64_t)S, p.out_len);
    const int64_t m0 = p.out_start + a0, m1 = p.out_start + a1;
    tlo_o = m0 - N + 1 <= 0 ? 0 : (m0 - N + 1 + hop - 1) / hop;
    int64_t th = (m1 - 1) / hop;
    if (th > p.T - 1) th = p.T - 1;
    nfr_o = (int)(th - tlo_o + 1);
  };
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int clip_i = (int)(tile / p.tiles_per_clip);
    const int tile_i = (int)(tile - (int64_t)clip_i * p.tiles_per_clip);
    const int64_t j0 = (int64_t)tile_i * S;                   // first output index of this tile
    const int64_t j1 = min(j0 + (int64_t)S, p.out_len);
    const int s_tile = (int)(j1 - j0);
    const int64_t n0 = p.out_start + j0, n1 = p.out_start + j1;  // OLA coordinates [n0, n1)
    int64_t t_lo = n0 - N + 1 <= 0 ? 0 : (n0 - N + 1 + hop - 1) / hop;
    int64_t t_hi = (n1 - 1) / hop;
    if (t_hi > p.T - 1) t_hi = p.T - 1;
    const int nfr = (int)(t_hi - t_lo + 1);
    // the first round of the NEXT tile, for the L2 prefetch issued during this tile's last round
    int nx_clip = 0, nx_nfr = 0;
    int64_t nx_tlo = 0;
    if (tile + gridDim.x < total_tiles) tile_frames(tile + gridDim.x, nx_clip, nx_tlo, nx_nfr);
    for (int i = threadIdx.x; i < s_tile; i += blockDim.x) acc[i] = 0.0f;
    // (the first round's barrier below orders these stores before the first accumulation)

    for (int c0 = 0; c0 < nfr; c0 += 2 * p.pairs_chunk) {
      const int cf = min(2 * p.pairs_chunk, nfr - c0);  // frames in this round
      const int cp = (cf + 1) / 2;
      // load + pack: Z[k] = conj(Xa[k] + i Xb[k]) over the full Hermitian-extended spectrum, so that
      // FFT(Z) = conj(N * (xa + i xb)).  A thread owns a BIN and walks the round's frame pairs: the frames of a bin are
      // adjacent in the (B, F, T) layout, so its loads fall into one or two 32-byte sectors (with bin-major threads every
      // 4- or 8-byte load opened a sector of its own and the load phase, latency-bound, was most of the kernel), and each
      // spectrum element is read once — the mirrored half of the extended spectrum is written from the same registers.
      // what the NEXT round (of this tile, or the first of the CTA's next tile) will read goes to L2 now: the load phase is
      // latency-bound (every warp of the SM waits on it) and nothing else keeps the DRAM busy during the transform
      int64_t pf_base = -1;
      int pf_n = 0;
      if (c0 + 2 * p.pairs_chunk < nfr) {
        pf_base = (int64_t)clip_i * p.clip_stride + t_lo + c0 + 2 * p.pairs_chunk;
        pf_n = min(2 * p.pairs_chunk, nfr - c0 - 2 * p.pairs_chunk);
      } else if (nx_nfr > 0) {
        pf_base = (int64_t)nx_clip * p.clip_stride + nx_tlo;
        pf_n = min(2 * p.pairs_chunk, nx_nfr);
      }
      if (pf_n > 0) {
        for (int kk = threadIdx.x; kk < F; kk += blockDim.x) {
          const int64_t b0 = pf_base + (int64_t)kk * p.T, b1 = b0 + pf_n - 1;
          if (p.spec) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec + b0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec + b1));
          } else {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_re + b0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_re + b1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_im + b0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_im + b1));
          }
        }
      }
      // One item = (bin, frame PAIR), pair fastest: the lanes that share a bin read ADJACENT frame pairs, so a warp's load covers
      // 32 / cp bins with cp * 8 (planes) or cp * 16 (complex) contiguous bytes each — whole sectors when four pairs make a round.
      // With one thread per bin every load instruction opened 32 sectors and used 4-8 bytes of each (the load / store queue was
      // the second stall reason of this phase, which was half of the MossFormer2 inverse).  An item is exactly one packed FFT input:
      // z = xa + i xb of frames 2 pr, 2 pr + 1.  KB items per thread are loaded before anything is converted or stored.
      constexpr int KB = 4;
      const int n_items = F * cp;
      const unsigned mg_cp = div_magic(cp);
      for (int i0 = threadIdx.x; i0 < n_items; i0 += KB * blockDim.x) {
        float2 xa[KB], xb[KB];
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          const int i = i0 + b * blockDim.x;
          xa[b] = make_float2(0.f, 0.f);
          xb[b] = make_float2(0.f, 0.f);
          if (i >= n_items) continue;
          const int kk = cp > 1 ? fast_div(i, cp, mg_cp) : i, pr = i - kk * cp;
          const bool has_b = 2 * pr + 1 < cf;
          const int64_t base = (int64_t)clip_i * p.clip_stride + (int64_t)kk * p.T + t_lo + c0 + 2 * pr;
          if (p.spec) {
            xa[b] = p.spec[base];
            if (has_b) xb[b] = p.spec[base + 1];
          } else if (has_b && p.planes_vec2 && (base & 1) == 0) {
            const float2 r = *reinterpret_cast<const float2*>(p.spec_re + base), im = *reinterpret_cast<const float2*>(p.spec_im + base);
            xa[b] = make_float2(r.x, im.x);
            xb[b] = make_float2(r.y, im.y);
          } else {
            xa[b] = make_float2(p.spec_re[base], p.spec_im[base]);
            if (has_b) xb[b] = make_float2(p.spec_re[base + 1], p.spec_im[base + 1]);
          }
        }
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          const int i = i0 + b * blockDim.x;
          if (i >= n_items) break;
          const int kk = cp > 1 ? fast_div(i, cp, mg_cp) : i, pr = i - kk * cp;
          const bool real_bin = kk == 0 || 2 * kk == N;  // irfft ignores Im(DC) and, for even N, Im(Nyquist)
          const int km = real_bin ? -1 : N - kk;         // mirrored position (bins > N/2), none for DC / Nyquist
          float2 a2 = xa[b], b2 = xb[b];
          if (!p.spec && p.polar.polar) {
            a2 = polar_to_complex(p.polar, a2);
            if (2 * pr + 1 < cf) b2 = polar_to_complex(p.polar, b2);
          }
          if (real_bin) { a2.y = 0.f; b2.y = 0.f; }
          // z = xa + i xb ; store conj(z).  Mirrored bin: conj(xa) + i conj(xb)
          bufA[(size_t)pr * NP + skew(kk)] = make_float2(a2.x - b2.y, -(a2.y + b2.x));
          if (km >= 0) bufA[(size_t)pr * NP + skew(km)] = make_float2(a2.x + b2.y, -(b2.x - a2.y));
        }
      }
      __syncthreads();
      const float2* Z = run_fft(p.fft, bufA, bufB, tw_s, cp);
      // accumulate: a thread owns output samples; for each it walks the round's frames in ascending order
      const int64_t first = (t_lo + c0) * hop - n0;  // tile-relative position of sample 0 of the round's first frame
      int lo = first < 0 ? 0 : (int)first;
      int64_t hi64 = first + (int64_t)(cf - 1) * hop + N;
      const int hi = hi64 > s_tile ? s_tile : (int)hi64;
      for (int n = lo + threadIdx.x; n < hi; n += blockDim.x) {
        float sum = acc[n];
        const int k0 = n - (int)first;  // sample index within the round's frame 0 (>= 0); decreases by hop per frame
        // the frames that hold this sample: 0 <= k0 - f*hop < N — walked without per-frame range tests
        const int f_lo = k0 - N + 1 <= 0 ? 0 : div_hop(k0 - N + hop);
        int f_hi = div_hop(k0);
        f_hi = f_hi > cf - 1 ? cf - 1 : f_hi;
        int k = k0 - f_lo * hop;
        for (int f = f_lo; f <= f_hi; ++f, k -= hop) {
          const float2 z = Z[(size_t)(f >> 1) * NP + skew(k)];
          sum += (((f & 1) == 0 ? z.x : -z.y) * invN) * win_s[k];
        }
        acc[n] = sum;
      }
      __syncthreads();
    }
    // envelope (position only, ascending frame order) and the division
    float* o = p.out + (int64_t)clip_i * p.out_clip_stride;
    const int64_t qn = n0 / hop, cn = n0 - N + hop, qc = cn >= 0 ? cn / hop : 0;
    const int rn = (int)(n0 - qn * hop), rc = cn >= 0 ? (int)(cn - qc * hop) : 0;
    const bool fast_idx = fdiv_ok && cn >= 0;  // every sample of the tile has n >= N: both quotients are tile-relative
    for (int i = threadIdx.x; i < s_tile; i += blockDim.x) {
      const int64_t n = n0 + i;
      int64_t ta, tb;
      if (fast_idx) {
        ta = qc + fast_div(rc + i, hop, mg_hop);
        tb = qn + fast_div(rn + i, hop, mg_hop);
      } else {
        ta = n - N + 1 <= 0 ? 0 : (n - N + 1 + hop - 1) / hop;
        tb = n / hop;
      }
      if (tb > p.T - 1) tb = p.T - 1;
      float den = 0.f;
      int k = (int)(n - ta * hop);
      for (int64_t t = ta; t <= tb; ++t, k -= hop) {
        const float w = win_s[k];
        den += p.norm_sq ? w * w : w;
      }
      const float num = acc[i];
      float r;
      if (p.div_clamp) r = num / fmaxf(den, p.div_eps);
      else r = den > p.div_eps ? num / den : num;
      o[j0 + i] = r;
    }
    __syncthreads();
  }
}

static FftDesc make_fft_desc(const b2a_plan* plan) {
  FftDesc d;
  d.n = plan->n_fft;
  d.nstages = plan->nstages;
  for (int i = 0; i < kMaxStages; ++i) d.radix[i] = plan->radix[i];
  d.tw_plain = plan->d_twiddle;
  return d;
}

}  // namespace

// Tile budget of the generic kernels: dynamic shared memory that still fits the SM's 196 KB shared-memory configuration (1 KB
// reserved per CTA, ~1 KB static), which leaves 60 KB of L1.  One step further — the 228 KB configuration, 28 KB of L1 — the
// spectrum loads of the inverse (half a 32-byte sector per bin and round; the next round finds the other half in L1) and the
// window / filterbank reads start missing: measured 15-25 % slower on every shape (MossFormer2 inverse 7.3 -> 9.1 ms).
constexpr size_t kTileBudget = 194 * 1024;

size_t generic_smem_limit(const b2a_plan* plan) {
  (void)plan;
  return 200 * 1024;
}

// dynamic shared memory of frontend_generic_kernel: sample span (padded to 16 bytes), two FFT buffers, twiddle table
static size_t fwd_smem_bytes(int ft, int N, int hop) {
  const size_t span = (size_t)(ft - 1) * hop + N;
  return (size_t)4 * ((span + 3) & ~(size_t)3) + (size_t)16 * (ft / 2) * seq_pitch(N) + (size_t)8 * N;
}

static int choose_frames_per_tile(int N, int hop, size_t budget, int max_ft) {
  int best = 0;
  for (int ft = 2; ft <= max_ft; ft += 2)
    if (fwd_smem_bytes(ft, N, hop) <= budget) best = ft;
  return best;
}

int generic_tile_frames(const b2a_plan* plan, const b2a_forward_args* a) {
  const b2a_frontend_desc& d = plan->fd;
  // (the CSR filterbank tables ride in the same budget: at most ~4 * n_fft + 12 * n_mels bytes)
  const size_t mel_room = d.n_mels > 0 ? (size_t)4 * (plan->mel.nnz + 4) + (size_t)12 * d.n_mels : 0;
  int ft = choose_frames_per_tile(d.n_fft, d.hop, kTileBudget > mel_room + 32 * 1024 ? kTileBudget - mel_room : kTileBudget, 64);
  if (ft == 0) ft = choose_frames_per_tile(d.n_fft, d.hop, generic_smem_limit(plan), 2);
  if (ft == 0) return 0;
  // small inputs: shrink tiles so that the grid still covers the SMs
  while (ft > 2 && (int64_t)a->batch * ((a->frame_count + ft - 1) / ft) < 2 * plan->sm_count) ft -= 2;
  return ft;
}

int generic_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min,
                             double* feat_sums, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  FwdParams p;
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
  p.n_fft = d.n_fft;
  p.hop = d.hop;
  p.n_freqs = plan->n_freqs;
  p.pad_mode = d.pad_mode;
  p.preemph = d.preemph;
  p.spec_kind = d.spec_kind;
  p.spec_eps = d.spec_eps;
  p.n_mels = d.n_mels;
  p.log_kind = d.log_kind;
  p.guard_kind = d.guard_kind;
  p.guard_eps = d.guard_eps;
  p.apply_affine = d.affine_div != 0.0f;
  p.affine_add = d.affine_add;
  p.affine_div = d.affine_div;
  p.out_layout = d.out_layout;
  p.out = a->out;
  const int M = d.n_mels > 0 ? d.n_mels : plan->n_freqs;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * M;
  p.clip_max = clip_max;
  p.tile_min = tile_min;
  p.feat_sums = feat_sums;
  p.tw = plan->d_twiddle_passes;
  p.window = plan->d_window;
  p.mel_start = plan->mel.d_start;
  p.mel_len = plan->mel.d_len;
  p.mel_off = plan->mel.d_off;
  p.mel_w = plan->mel.d_w;
  p.fft = make_fft_desc(plan);
  p.frame_len = d.frame_len;
  p.frame_dc = d.frame_dc;
  p.frame_preemph = d.frame_preemph;
  p.dither = d.dither;
  p.seed = a->seed;
  const int ft = generic_tile_frames(plan, a);
  if (ft == 0) {
    set_error("n_fft=%d too large for the generic kernel", d.n_fft);
    return B2A_ERR_UNSUPPORTED;
  }
  p.frames_per_tile = ft;
  p.tiles_per_clip = (int)((a->frame_count + ft - 1) / ft);
  size_t smem = fwd_smem_bytes(ft, d.n_fft, d.hop);
  // the CSR filterbank rides behind the twiddles when it fits the slack between the tile budget (200 KB) and the 227 KB a CTA may own
  p.mel_nnz = d.n_mels > 0 ? plan->mel.nnz : 0;
  const size_t mel_bytes = d.n_mels > 0 ? (size_t)4 * ((p.mel_nnz + 3) & ~3) + (size_t)12 * d.n_mels : 0;
  p.mel_smem = d.n_mels > 0 && smem + mel_bytes <= 226 * 1024;
  if (p.mel_smem) smem += mel_bytes;
  B2A_CUDA(cudaFuncSetAttribute(frontend_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (int64_t)a->batch * p.tiles_per_clip;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  frontend_generic_kernel<<<grid, kThreads, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

int dump_frames(b2a_plan* plan, const b2a_forward_args* a, int apply_window, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  FwdParams p;
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
  p.n_fft = d.n_fft;
  p.hop = d.hop;
  p.n_freqs = plan->n_freqs;
  p.pad_mode = d.pad_mode;
  p.preemph = d.preemph;
  p.out = a->out;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * d.n_fft;
  p.window = plan->d_window;
  p.dump_frames = 1;
  p.dump_windowed = apply_window;
  int ft = choose_frames_per_tile(d.n_fft, d.hop, 96 * 1024, 16);
  if (ft == 0) ft = 2;
  p.frames_per_tile = ft;
  p.tiles_per_clip = (int)((a->frame_count + ft - 1) / ft);
  const size_t smem = fwd_smem_bytes(ft, d.n_fft, d.hop);
  B2A_CUDA(cudaFuncSetAttribute(frontend_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (int64_t)a->batch * p.tiles_per_clip;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * 2);
  if (grid < 1) grid = 1;
  frontend_generic_kernel<<<grid, kThreads, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

int deltas(const float* x, float* out, int64_t rows, int64_t cols, int win_length, int edge, cudaStream_t st) {
  const int n = (win_length - 1) / 2;
  const float denom = (float)((double)n * (n + 1) * (2 * n + 1) / 3.0);
  const int64_t total = rows * cols;
  if (total <= 0) return B2A_OK;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  deltas_kernel<<<grid, 256, 0, st>>>(x, out, rows, cols, n, denom, edge);
  B2A_LAUNCHED();
  return B2A_OK;
}

int init_stats(float* clip_max, double* feat_sums, int batch, int n_mels, cudaStream_t st) {
  const int n_sums = feat_sums ? batch * n_mels * 2 : 0;
  const int n = batch > n_sums ? batch : n_sums;
  if (n <= 0) return B2A_OK;
  init_stats_kernel<<<(n + 255) / 256, 256, 0, st>>>(clip_max, feat_sums, batch, n_sums);
  B2A_LAUNCHED();
  return B2A_OK;
}

int frontend_finalize(b2a_plan* plan, const b2a_forward_args* a, int64_t global_frames, float* clip_max,
                      const float* tile_min, int tile_frames, const double* feat_sums, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  if (d.clamp_kind == B2A_CLAMP_NONE && d.norm_kind == B2A_NORM_NONE) return B2A_OK;
  FinParams p;
  memset(&p, 0, sizeof(p));
  const int M = d.n_mels > 0 ? d.n_mels : plan->n_freqs;
  p.out = reinterpret_cast<float*>(a->out);
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * M;
  p.batch = a->batch;
  p.n_mels = M;
  p.out_layout = d.out_layout;
  p.frames = a->frame_count;
  p.global_frames = global_frames;
  p.clamp_kind = d.clamp_kind;
  p.clamp_value = d.clamp_value;
  p.apply_affine = d.affine_div != 0.0f;
  p.affine_add = d.affine_add;
  p.affine_div = d.affine_div;
  p.norm_kind = d.norm_kind;
  p.norm_ddof = d.norm_ddof;
  p.norm_eps = d.norm_eps;
  p.clip_max = clip_max;
  p.tile_min = tile_min;
  p.tile_frames = tile_frames;
  p.tiles_per_clip = (int)((a->frame_count + tile_frames - 1) / tile_frames);
  p.tile_min_pitch = p.tiles_per_clip;
  // trailing all-padding frames (fast family): constant rows, written below with the clamp applied — not fixed up
  const int64_t const_row0 = fast_const_row0(plan, a);
  if (const_row0 >= 0) p.tiles_per_clip = (int)(const_row0 / tile_frames);
  p.feat_sums = feat_sums;
  p.stats_affine = plan->family == KF_FAST ? 1 : 0;
  p.out_dtype = d.out_dtype;
  p.fill_unwritten = fast_skip_floor_tiles(plan) ? 1 : 0;
  p.ec = epilogue_consts(d);
  if (d.clamp_kind != B2A_CLAMP_NONE) {
    if (d.clamp_kind == B2A_CLAMP_BATCH_MAX) {
      batch_max_kernel<<<1, 256, 0, st>>>(clip_max, a->batch);
      B2A_LAUNCHED();
    }
    dim3 grid((p.tiles_per_clip + 7) / 8, a->batch);
    clamp_fixup_kernel<<<grid, 256, 0, st>>>(p);
    B2A_LAUNCHED();
    if (const_row0 >= 0) return fast_const_rows_finalize(plan, a, const_row0, clip_max, st);
    return B2A_OK;
  }
  const int64_t total = a->frame_count * M;
  int gx = (int)std::min<int64_t>((total + 256 * 16 - 1) / (256 * 16), 8 * plan->sm_count);
  if (gx < 1) gx = 1;
  if ((M & 3) == 0 && M <= 256) {  // grid stride (in quads) a multiple of the row length: per-thread features stay fixed
    const int M4 = M >> 2;
    int g = M4, b = 256;
    while (b) { const int t = g % b; g = b; b = t; }  // gcd(M4, 256)
    const int unit = M4 / g;
    gx = (gx + unit - 1) / unit * unit;
  }
  dim3 grid(gx, a->batch);
  normalise_kernel<<<grid, 256, 0, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

int generic_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st) {
  const b2a_istft_desc& d = plan->id;
  InvParams p;
  memset(&p, 0, sizeof(p));
  const int N = d.n_fft, hop = d.hop, F = plan->n_freqs;
  if (a->spec_imag) {
    p.spec = nullptr;
    p.spec_re = reinterpret_cast<const float*>(a->spec);
    p.spec_im = reinterpret_cast<const float*>(a->spec_imag);
  } else {
    p.spec = reinterpret_cast<const float2*>(a->spec);
  }
  p.T = a->num_frames;
  p.clip_stride = a->clip_stride ? a->clip_stride : (int64_t)F * a->num_frames;
  p.batch = a->batch;
  p.planes_vec2 = a->spec_imag && reinterpret_cast<uintptr_t>(a->spec) % 8 == 0 && reinterpret_cast<uintptr_t>(a->spec_imag) % 8 == 0;
  p.n_fft = N;
  p.hop = hop;
  p.n_freqs = F;
  p.norm_sq = d.norm_kind == B2A_ISTFT_NORM_WINDOW_SQ;
  p.div_clamp = d.div_kind == B2A_ISTFT_DIV_CLAMP;
  p.polar = make_polar_spec(d);
  p.div_eps = istft_div_eps(d);
  int64_t ola, start, len;
  int64_t eff_len = a->length;
  b2a_istft_geometry(a->num_frames, N, hop, d.center, d.trim_tail ? eff_len : -1, &ola, &start, &len);
  if (!d.trim_tail) {  // ISTFTCache: strip only the front, then [:audio_length]
    start = d.center ? N / 2 : 0;
    len = ola - start;
    if (len < 0) len = 0;
    if (a->length >= 0 && a->length < len) len = a->length;
  }
  p.out_start = start;
  p.out_len = len;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : len;
  p.out = a->out;
  p.tw = plan->d_twiddle_passes;
  p.window = plan->d_window;
  p.fft = make_fft_desc(plan);
  if (len <= 0) return B2A_OK;
  const int ov = (N + hop - 1) / hop;  // frames overlapping one sample (upper bound)
  // pairs per FFT round: enough butterflies for the CTA (a radix-16 pass over two 2048-point pairs is 256 butterflies for 512
  // threads; four pairs also make a bin's frames of one round a whole 32-byte sector), as long as the tile still advances by
  // >= 4 * ov frames — every tile re-transforms the ov - 1 frames it shares with its left neighbour.  The tile stays within
  // kTileBudget (the 196 KB shared-memory configuration).
  auto smem_for = [&](int pc, int adv) { return (size_t)16 * pc * seq_pitch(N) + (size_t)8 * N + (size_t)4 * N + (size_t)4 * adv * hop + 16; };
  auto adv_for = [&](int pc) {
    int best = 0;
    for (int cand = 1; cand <= 1024; cand *= 2)
      if (smem_for(pc, cand) <= kTileBudget && (int64_t)cand * hop <= (1 << 20)) best = cand;
    return best;
  };
  static const int pc_override = getenv("B2A_X_PC") ? atoi(getenv("B2A_X_PC")) : 0;  // development: pairs per FFT round
  int pairs_chunk = 8192 / N;
  pairs_chunk = pairs_chunk < 1 ? 1 : (pairs_chunk > 16 ? 16 : pairs_chunk);
  if (adv_for(pairs_chunk) < 4 * ov) pairs_chunk = pairs_chunk / 2 < 1 ? 1 : pairs_chunk / 2;
  if (pc_override > 0) pairs_chunk = pc_override;
  int adv = adv_for(pairs_chunk);
  if (adv < 4 * ov) {  // large transforms: one CTA per SM, all of its shared memory
    for (int cand = adv > 0 ? adv : 1; cand <= 1024; cand *= 2)
      if (smem_for(pairs_chunk, cand) <= generic_smem_limit(plan)) adv = cand;
  }
  if (adv == 0) {
    pairs_chunk = 1;
    adv = 1;
    if (smem_for(1, 1) > generic_smem_limit(plan)) {
      set_error("n_fft=%d / hop=%d too large for the generic iSTFT kernel", N, hop);
      return B2A_ERR_UNSUPPORTED;
    }
  }
  while (adv > 1 && (int64_t)a->batch * ((len + (int64_t)adv * hop - 1) / ((int64_t)adv * hop)) < 2 * plan->sm_count)
    adv /= 2;
  p.frames_adv = adv;
  p.frames_cap = adv + ov;
  p.pairs_chunk = pairs_chunk;
  const int64_t S = (int64_t)adv * hop;
  p.tiles_per_clip = (int)((len + S - 1) / S);
  const size_t smem = smem_for(pairs_chunk, adv);
  B2A_CUDA(cudaFuncSetAttribute(istft_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (int64_t)a->batch * p.tiles_per_clip;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  istft_generic_kernel<<<grid, kThreads, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

}  // namespace b2a
