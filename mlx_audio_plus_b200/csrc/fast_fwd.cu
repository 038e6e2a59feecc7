// b200audio — specialised fused log-mel front-end kernels ("fast" family).
//
// Design: one CTA owns a tile of 32 consecutive frames of one clip; LANE == FRAME,
// WARP == COLUMN ROLE.  A real frame of n_fft = 2*Nc samples is treated as Nc complex samples
// z[m] = x[2m] + i x[2m+1] (half-size complex FFT + Hermitian post-twiddle); the Nc-point FFT is split
// Nc = N1 x N2 and BOTH stages run entirely in registers with compile-time twiddles (fft_regs.cuh):
//
//   fill     the tile's contiguous sample span is copied ONCE global->shared with cp.async (8 B / thread,
//            fully coalesced), in rows of `hop` samples with a padded pitch so that frame-strided reads are
//            bank-conflict free; the next tile's span is prefetched while stage 2 / mel of this tile run.
//   stage 1  warp = column n2 (N2 of them): each lane loads its frame's N1 strided complex samples, applies
//            the window (warp-uniform, broadcast from smem), DFT-N1 in registers, multiplies the inter-stage
//            twiddle W_Nc^(n2*k1) (warp-uniform) and stores to the exchange buffer E[frame][slot(k1)][n2].
//   stage 2  warp = column pair (k1, N1-k1): two DFT-N2 in registers give Z[k] and Z[Nc-k] in the SAME
//            thread, so the real-FFT post-twiddle X[k] = E + W_N^k O, the power / magnitude and the store to
//            P[frame][k] need no further exchange.
//   mel      warp = subset of mel rows, lane = frame: the filterbank is a banded CSR (<= 2 non-zeros per
//            bin); start/len/weights are warp-uniform smem broadcasts; guard, MUFU log2, fused
//            scale+affine; values staged in smem, per-tile max/min reduced by shuffles, per-mel sums in fp64.
//   store    coalesced 128 B rows to HBM.
// Every shared-memory access pattern is lane-strided by an ODD pitch (conflict free) or a broadcast.
// HBM traffic is the compulsory input-once + output-once.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "fft_regs.cuh"

namespace b2a {

namespace {

using regs::Dft;
using regs::static_for;

template <int N1_, int N2_, int HOP_>
struct Cfg {
  static constexpr int N1 = N1_, N2 = N2_, HOP = HOP_;
  static constexpr int NC = N1 * N2, N = 2 * NC, F = NC + 1;
  static constexpr int WARPS = N1 / 2;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int RPW = N2 / WARPS;  // stage-1 roles per warp
  static constexpr int FT = 32;           // frames per tile == warp width
  static constexpr int P = HOP + (((HOP / 2) % 2 == 0) ? 2 : 0);  // row pitch (floats); P/2 odd
  static constexpr int ROWS = FT - 1 + (N + HOP - 1) / HOP;
  static constexpr int SPAN = (FT - 1) * HOP + N;
  static constexpr int XS_FLOATS = ROWS * P;
  static constexpr int EP = NC + 1;          // exchange pitch per frame (float2), odd
  static constexpr int PP = (F % 2) ? F : F + 1;  // power pitch per frame (floats), odd
  static constexpr int K = HOP / (2 * N2);   // taps pairs per row per role step
  static_assert(N1 % 2 == 0 && N2 % WARPS == 0, "role split");
  static_assert(HOP % (2 * N2) == 0, "hop must be a multiple of 2*N2");
  static_assert(NC % 2 == 0, "Nc even");
};

struct FastParams {
  const float* audio;
  int64_t clip_stride, valid_length, sample_offset, frame_begin, frame_count;
  float pad_value;
  int batch;
  Geometry geo;
  int pad_mode;
  float preemph;
  int fast_fill_ok;  // alignment preconditions for the cp.async path
  int spec_kind;
  float spec_eps;
  int n_mels, guard_kind;
  float guard_eps;
  float log_scale;   // 0 = no log; else y = log2(x) * log_scale  (ln2 or log10(2)), then affine
  float aff_mul, aff_add;  // y' = y * aff_mul + aff_add   (aff_mul = 1/div, aff_add = add/div)
  int out_layout;
  float* out;
  int64_t out_clip_stride;
  float *clip_max, *clip_min;  // affine-domain statistics
  double* feat_sums;
  const float2* win2;   // [NC] (w[2m], w[2m+1]) * 0.5
  const float2* tw1;    // [N2][N1]  W_Nc^(n2*k1)
  const float2* twp;    // [NC/2+1.. NC] W_N^k, k = 0..NC
  const int *mel_start, *mel_len, *mel_off;
  const float* mel_w;
  int mel_nnz;
  int tiles_per_clip;
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (v > __int_as_float(old)) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (v < __int_as_float(old)) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

__device__ __forceinline__ float fetch_sample_f(const FastParams& p, const float* clip, int64_t s) {
  float x = s < p.valid_length ? __ldg(clip + (s - p.sample_offset)) : p.pad_value;
  if (p.preemph != 0.0f && s > 0) {
    const int64_t sm = s - 1;
    const float xm = sm < p.valid_length ? __ldg(clip + (sm - p.sample_offset)) : p.pad_value;
    x = __fsub_rn(x, __fmul_rn(p.preemph, xm));
  }
  return x;
}

template <class C>
__device__ __forceinline__ void fill_tile(const FastParams& p, float* xs, int64_t tile) {
  const int clip_i = (int)(tile / p.tiles_per_clip);
  const int tile_i = (int)(tile - (int64_t)clip_i * p.tiles_per_clip);
  const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
  const int64_t lt0 = (int64_t)tile_i * C::FT;
  const int64_t q0 = (p.frame_begin + lt0) * C::HOP;  // padded coordinate of the tile's first sample
  const int64_t s0 = q0 - p.geo.pad_left;             // source coordinate
  const bool interior = p.fast_fill_ok && s0 >= 0 && (s0 + C::SPAN) <= p.valid_length && s0 >= p.sample_offset;
  if (interior) {
    const float* src = clip + (s0 - p.sample_offset);
    for (int j = threadIdx.x; j < C::SPAN / 2; j += C::THREADS) {
      const int s = 2 * j;
      const int row = s / C::HOP, col = s - row * C::HOP;
      cp_async8(xs + row * C::P + col, src + s);
    }
  } else {
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);
    const int need = (nf - 1) * C::HOP + C::N;
    for (int i = threadIdx.x; i < C::SPAN; i += C::THREADS) {
      float v = 0.0f;
      if (i < need) {
        const int64_t s = source_index(p.geo, p.pad_mode, q0 + i);
        if (s >= 0) v = fetch_sample_f(p, clip, s);
      }
      const int row = i / C::HOP, col = i - row * C::HOP;
      xs[row * C::P + col] = v;
    }
  }
  cp_async_commit();
}

// real-FFT post-twiddle for one bin pair (k, Nc-k); Zk = Z[k], Zm = Z[Nc-k]; w = W_N^k; all scaled by the
// 0.5 folded into the window.  Returns |X[k]|^2 and |X[Nc-k]|^2.
__device__ __forceinline__ void post_pair(float2 zk, float2 zm, float2 w, float& pk, float& pm) {
  const float ex = zk.x + zm.x, ey = zk.y - zm.y;
  const float ox = zk.y + zm.y, oy = zm.x - zk.x;
  const float tx = w.x * ox - w.y * oy, ty = w.x * oy + w.y * ox;
  const float ar = ex + tx, ai = ey + ty;
  const float br = ex - tx, bi = ey - ty;
  pk = ar * ar + ai * ai;
  pm = br * br + bi * bi;
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 2) fast_logmel_kernel(const FastParams p) {
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve
  float2* s_win2 = reinterpret_cast<float2*>(smem_raw);             // [NC]
  float2* s_tw1 = s_win2 + NC;                                       // [N2*N1]
  float2* s_twp = s_tw1 + NC;                                        // [NC+1] (+1 pad)
  float2* E = s_twp + (NC + 2);                                      // [FT*EP]
  float* Pw = reinterpret_cast<float*>(E + C::FT * C::EP);           // [FT*PP]
  float* xs = Pw + C::FT * C::PP + 1;                                // [XS_FLOATS]  (8 B aligned below)
  xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(xs) + 7) & ~(uintptr_t)7);
  int* s_mel = reinterpret_cast<int*>(xs + C::XS_FLOATS);            // start[M], len[M], off[M]
  const int M = p.n_mels;
  float* s_melw = reinterpret_cast<float*>(s_mel + 3 * M);           // [nnz]
  double* s_sums = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_melw + p.mel_nnz) + 7) & ~(uintptr_t)7);  // [2*M]
  __shared__ float red_max[C::WARPS], red_min[C::WARPS];
  __shared__ int s_cur_clip;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int YP = M | 1;
  float* Y = reinterpret_cast<float*>(E);  // [FT*YP], aliases the exchange buffer

  for (int i = threadIdx.x; i < NC; i += C::THREADS) {
    s_win2[i] = p.win2[i];
    s_tw1[i] = p.tw1[i];
  }
  for (int i = threadIdx.x; i <= NC; i += C::THREADS) s_twp[i] = p.twp[i];
  for (int i = threadIdx.x; i < M; i += C::THREADS) {
    s_mel[i] = p.mel_start[i];
    s_mel[M + i] = p.mel_len[i];
    s_mel[2 * M + i] = p.mel_off[i];
  }
  for (int i = threadIdx.x; i < p.mel_nnz; i += C::THREADS) s_melw[i] = p.mel_w[i];
  if (p.feat_sums)
    for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) s_sums[i] = 0.0;
  if (threadIdx.x == 0) s_cur_clip = -1;

  const int64_t total_tiles = (int64_t)p.batch * p.tiles_per_clip;
  int64_t tile = blockIdx.x;
  if (tile < total_tiles) fill_tile<C>(p, xs, tile);

  for (; tile < total_tiles; tile += gridDim.x) {
    const int clip_i = (int)(tile / p.tiles_per_clip);
    const int tile_i = (int)(tile - (int64_t)clip_i * p.tiles_per_clip);
    const int64_t lt0 = (int64_t)tile_i * C::FT;
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);

    cp_async_wait_all();
    __syncthreads();  // xs ready; previous tile's Y fully written out

    // per-CTA running per-mel sums: flush when the clip changes
    if (p.feat_sums && s_cur_clip != clip_i) {
      const int prev = s_cur_clip;
      __syncthreads();
      if (prev >= 0)
        for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) {
          atomicAdd(p.feat_sums + (int64_t)prev * 2 * M + i, s_sums[i]);
          s_sums[i] = 0.0;
        }
      if (threadIdx.x == 0) s_cur_clip = clip_i;
      __syncthreads();
    }

    // ---- stage 1 ----------------------------------------------------------------------------------------
#pragma unroll 1
    for (int rr = 0; rr < C::RPW; ++rr) {
      const int n2 = warp * C::RPW + rr;
      const float* xb = xs + lane * C::P + 2 * n2;
      const float2* wb = s_win2 + n2;
      float2 v[N1];
      static_for<0, N1>([&](auto I_) {
        constexpr int n1 = decltype(I_)::value;
        constexpr int off = (n1 / C::K) * C::P + (n1 % C::K) * 2 * N2;
        const float2 x = *reinterpret_cast<const float2*>(xb + off);
        const float2 w = wb[N2 * n1];
        v[n1] = make_float2(x.x * w.x, x.y * w.y);
      });
      Dft<N1>::run(v);
      const float2* tb = s_tw1 + n2 * N1;
      float2* eb = E + lane * C::EP + n2;
      static_for<0, N1>([&](auto I_) {
        constexpr int k1 = decltype(I_)::value;
        constexpr int slot = (k1 <= N1 / 2) ? k1 : (3 * N1 / 2 - k1);
        float2 y = v[k1];
        if constexpr (k1 > 0) y = regs::cmul(y, tb[k1]);
        eb[slot * N2] = y;
      });
    }
    __syncthreads();  // E complete, xs free

    // prefetch the next tile's samples while stage 2 / mel run
    {
      const int64_t next = tile + gridDim.x;
      if (next < total_tiles) fill_tile<C>(p, xs, next);
    }

    // ---- stage 2 ----------------------------------------------------------------------------------------
    {
      const int u = warp;  // column pair (u, N1-u); u == 0 owns columns 0 and N1/2
      float2 A[N2], B[N2];
      const float2* ea = E + lane * C::EP + u * N2;
      const float2* eb = E + lane * C::EP + (N1 / 2 + u) * N2;
      static_for<0, N2>([&](auto I_) {
        constexpr int j = decltype(I_)::value;
        A[j] = ea[j];
        B[j] = eb[j];
      });
      Dft<N2>::run(A);
      Dft<N2>::run(B);
      float* pr = Pw + lane * C::PP;
      const bool pw_only = p.spec_kind == B2A_SPEC_POWER;
      const float eps = p.spec_eps;
      auto emit = [&](int k, float v) { pr[k] = pw_only ? v : sqrtf(v + eps); };
      if (u != 0) {
        const float2* tw = s_twp + u;
        static_for<0, N2>([&](auto I_) {
          constexpr int k2 = decltype(I_)::value;
          float pk, pm;
          post_pair(A[k2], B[N2 - 1 - k2], tw[N1 * k2], pk, pm);
          const int k = u + N1 * k2;
          emit(k, pk);
          emit(NC - k, pm);
        });
      } else {
        static_for<0, N2 / 2 + 1>([&](auto I_) {  // column 0: k = N1*k2 <-> Nc - k = N1*(N2-k2)
          constexpr int k2 = decltype(I_)::value;
          float pk, pm;
          post_pair(A[k2], A[(N2 - k2) % N2], s_twp[N1 * k2], pk, pm);
          emit(N1 * k2, pk);
          emit(NC - N1 * k2, pm);
        });
        static_for<0, N2 / 2>([&](auto I_) {  // column N1/2: k = N1/2 + N1*k2 <-> N1/2 + N1*(N2-1-k2)
          constexpr int k2 = decltype(I_)::value;
          float pk, pm;
          post_pair(B[k2], B[N2 - 1 - k2], s_twp[N1 / 2 + N1 * k2], pk, pm);
          emit(N1 / 2 + N1 * k2, pk);
          emit(NC - (N1 / 2 + N1 * k2), pm);
        });
      }
    }
    __syncthreads();  // Pw complete, E free (Y aliases E)

    // ---- mel projection + log + affine; lane = frame, warp = mel subset ------------------------------------
    {
      const float* pr = Pw + lane * C::PP;
      float* yr = Y + lane * YP;
      float lmax = -INFINITY, lmin = INFINITY;
      const bool valid = lane < nf;
      for (int m = warp; m < M; m += C::WARPS) {
        const int start = s_mel[m], len = s_mel[M + m];
        const float* wt = s_melw + s_mel[2 * M + m];
        const float* pp = pr + start;
        float acc = 0.0f;
        for (int j = 0; j < len; ++j) acc = fmaf(pp[j], wt[j], acc);
        if (p.guard_kind == B2A_GUARD_MAX) acc = fmaxf(acc, p.guard_eps);
        else if (p.guard_kind == B2A_GUARD_ADD) acc = acc + p.guard_eps;
        float y = acc;
        if (p.log_scale != 0.0f) y = __log2f(acc) * p.log_scale;
        y = fmaf(y, p.aff_mul, p.aff_add);
        if (valid) {
          lmax = fmaxf(lmax, y);
          lmin = fminf(lmin, y);
        }
        yr[m] = y;
        if (p.feat_sums) {
          double d1 = valid ? (double)y : 0.0, d2 = d1 * d1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
          }
          if (lane == 0) {
            s_sums[2 * m] += d1;
            s_sums[2 * m + 1] += d2;
          }
        }
      }
      if (p.clip_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
          lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        }
        if (lane == 0) {
          red_max[warp] = lmax;
          red_min[warp] = lmin;
        }
      }
    }
    __syncthreads();
    if (p.clip_max && threadIdx.x == 0) {
      float a = red_max[0], b = red_min[0];
#pragma unroll
      for (int w = 1; w < C::WARPS; ++w) {
        a = fmaxf(a, red_max[w]);
        b = fminf(b, red_min[w]);
      }
      atomic_max_f(p.clip_max + clip_i, a);
      atomic_min_f(p.clip_min + clip_i, b);
    }
    // ---- coalesced store ----------------------------------------------------------------------------------
    {
      float* o = p.out + (int64_t)clip_i * p.out_clip_stride;
      if (p.out_layout == B2A_LAYOUT_TM) {
        float* ot = o + lt0 * M;
        int f = 0, m = threadIdx.x;
        while (m >= M) { m -= M; ++f; }
        while (f < nf) {
          ot[f * M + m] = Y[f * YP + m];
          m += C::THREADS;
          while (m >= M) { m -= M; ++f; }
        }
      } else {
        for (int i = threadIdx.x; i < M * 32; i += C::THREADS) {
          const int m = i >> 5, f = i & 31;
          if (f < nf) o[(int64_t)m * p.frame_count + lt0 + f] = Y[f * YP + m];
        }
      }
    }
  }
  cp_async_wait_all();
  if (p.feat_sums) {
    __syncthreads();
    const int prev = s_cur_clip;
    if (prev >= 0)
      for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) atomicAdd(p.feat_sums + (int64_t)prev * 2 * M + i, s_sums[i]);
  }
}

template <class C>
size_t smem_bytes(int M, int nnz) {
  size_t b = 0;
  b += sizeof(float2) * (C::NC + C::NC + C::NC + 2);
  b += sizeof(float2) * C::FT * C::EP;
  b += sizeof(float) * (C::FT * C::PP + 1) + 8;
  b += sizeof(float) * C::XS_FLOATS;
  b += sizeof(int) * 3 * M + sizeof(float) * nnz + 8;
  b += sizeof(double) * 2 * M;
  return b + 16;
}

struct FastState {
  float2* d_win2 = nullptr;
  float2* d_tw1 = nullptr;
  float2* d_twp = nullptr;
  int variant = 0;  // 1: 400/160, 2: 512/160, 3: 1024/256
};

template <class C>
int launch(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) {
  const size_t smem = smem_bytes<C>(p.n_mels, p.mel_nnz);
  static size_t attr_smem = 0;
  if (smem > attr_smem && smem <= 226 * 1024) {
    B2A_CUDA(cudaFuncSetAttribute(fast_logmel_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  if (smem > 226 * 1024) {
    set_error("fast kernel: %zu bytes of shared memory needed", smem);
    return B2A_ERR_UNSUPPORTED;
  }
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, 2));
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  fast_logmel_kernel<C><<<grid, C::THREADS, smem, st>>>(p);
  B2A_CUDA(cudaGetLastError());
  return B2A_OK;
}

using Cfg400 = Cfg<20, 10, 160>;
using Cfg512 = Cfg<16, 16, 160>;

}  // namespace

bool fast_frontend_supported(const b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  if (d.n_mels <= 0 || d.spec_kind == B2A_SPEC_COMPLEX) return false;
  if (d.affine_div < 0.0f) return false;
  const bool v400 = d.n_fft == 400 && d.hop == 160;
  const bool v512 = d.n_fft == 512 && d.hop == 160;
  if (!(v400 || v512)) return false;
  if (d.n_mels > 256) return false;
  return true;
}

int fast_frontend_init(b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = new FastState();
  plan->fast = fs;
  int N1, N2;
  if (d.n_fft == 400) { fs->variant = 1; N1 = 20; N2 = 10; }
  else { fs->variant = 2; N1 = 16; N2 = 16; }
  const int NC = N1 * N2, N = 2 * NC;
  std::vector<float2> win2(NC), tw1(NC), twp(NC + 2);
  for (int m = 0; m < NC; ++m)  // 0.5 of the real-FFT post-processing is folded into the window
    win2[m] = make_float2(0.5f * plan->h_window[2 * m], 0.5f * plan->h_window[2 * m + 1]);
  for (int n2 = 0; n2 < N2; ++n2)
    for (int k1 = 0; k1 < N1; ++k1) {
      const double a = -2.0 * M_PI * (double)((n2 * k1) % NC) / (double)NC;
      tw1[n2 * N1 + k1] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k = 0; k <= NC; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)N;
    twp[k] = make_float2((float)cos(a), (float)sin(a));
  }
  twp[NC + 1] = make_float2(0.f, 0.f);
  B2A_CUDA(cudaMalloc(&fs->d_win2, sizeof(float2) * NC));
  B2A_CUDA(cudaMalloc(&fs->d_tw1, sizeof(float2) * NC));
  B2A_CUDA(cudaMalloc(&fs->d_twp, sizeof(float2) * (NC + 2)));
  B2A_CUDA(cudaMemcpy(fs->d_win2, win2.data(), sizeof(float2) * NC, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_tw1, tw1.data(), sizeof(float2) * NC, cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_twp, twp.data(), sizeof(float2) * (NC + 2), cudaMemcpyHostToDevice));
  plan->kernel_name = fs->variant == 1 ? "fast_logmel_400x160" : "fast_logmel_512x160";
  return B2A_OK;
}

void fast_frontend_destroy(b2a_plan* plan) {
  FastState* fs = reinterpret_cast<FastState*>(plan->fast);
  if (!fs) return;
  cudaFree(fs->d_win2);
  cudaFree(fs->d_tw1);
  cudaFree(fs->d_twp);
  delete fs;
  plan->fast = nullptr;
}

int fast_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* clip_min,
                          double* feat_sums, cudaStream_t st) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = reinterpret_cast<FastState*>(plan->fast);
  FastParams p;
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
  p.fast_fill_ok = d.preemph == 0.0f && (reinterpret_cast<uintptr_t>(a->audio) % 8 == 0) && (a->clip_stride % 2 == 0) &&
                   (p.geo.pad_left % 2 == 0) && (a->sample_offset % 2 == 0);
  p.spec_kind = d.spec_kind;
  p.spec_eps = d.spec_kind == B2A_SPEC_SQRT_POWER_EPS ? d.spec_eps : 0.0f;
  p.n_mels = d.n_mels;
  p.guard_kind = d.guard_kind;
  p.guard_eps = d.guard_eps;
  p.log_scale = d.log_kind == B2A_LOG_LOG10 ? 0.30102999566398119521f : (d.log_kind == B2A_LOG_LN ? 0.69314718055994530942f : 0.0f);
  if (d.affine_div != 0.0f) {
    p.aff_mul = 1.0f / d.affine_div;
    p.aff_add = d.affine_add / d.affine_div;
  } else {
    p.aff_mul = 1.0f;
    p.aff_add = 0.0f;
  }
  p.out_layout = d.out_layout;
  p.out = reinterpret_cast<float*>(a->out);
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * d.n_mels;
  p.clip_max = clip_max;
  p.clip_min = clip_min;
  p.feat_sums = feat_sums;
  p.win2 = fs->d_win2;
  p.tw1 = fs->d_tw1;
  p.twp = fs->d_twp;
  p.mel_start = plan->mel.d_start;
  p.mel_len = plan->mel.d_len;
  p.mel_off = plan->mel.d_off;
  p.mel_w = plan->mel.d_w;
  p.mel_nnz = plan->mel.nnz;
  p.tiles_per_clip = (int)((a->frame_count + 31) / 32);
  if (fs->variant == 1) return launch<Cfg400>(plan, fs, p, st);
  return launch<Cfg512>(plan, fs, p, st);
}

}  // namespace b2a
