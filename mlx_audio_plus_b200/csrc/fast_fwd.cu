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

template <int N1_, int N2_, int HOP_, bool ALIAS_, int MIN_BLOCKS_>
struct Cfg {
  static constexpr int N1 = N1_, N2 = N2_, HOP = HOP_;
  // ALIAS: the power tile P reuses the sample tile's shared memory (no prefetch of the next tile) so that the
  // CTA fits the occupancy target; otherwise the next tile's samples are prefetched during stage 2 / mel.
  static constexpr bool ALIAS = ALIAS_;
  static constexpr int MIN_BLOCKS = MIN_BLOCKS_;
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
  // power pitch per frame (floats): odd (stage-2 lane==frame stores are conflict free) and == 9 (mod 32) so that
  // the mel phase's (4 frames x 8 mel rows) gathers land in distinct banks
  static constexpr int PP = F + ((9 - F % 32 + 32) % 32);
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
  int debug_skip;    // profiling aid (B2A_SKIP bitmask): 1 stage1, 2 stage2, 4 mel, 16 output stores only
  int spec_kind;
  float spec_eps;
  int n_mels;
  float guard_add, guard_floor;  // a = max(a + guard_add, guard_floor)   (ADD: (eps, -inf); MAX: (0, eps))
  int use_log;                   // y = log2(a) if use_log else a
  float y_mul, y_add;            // y' = y * y_mul + y_add   (log base change and the affine map folded together)
  int out_layout;
  float* out;
  int64_t out_clip_stride;
  float *clip_max, *tile_min;  // affine-domain statistics (per clip max, per tile min)
  double* feat_sums;
  const float2* win2;   // [N2][N1] (w[2m], w[2m+1]) * 0.5 with m = N2*n1 + n2
  const float2* tw1;    // [N2][N1]  W_Nc^(n2*k1) as (wr, wi); two per LDS.128 broadcast
  const float2* twp;    // [N1/2][2*N2] post-twiddles W_N^k in the order stage 2 consumes them
  // mel filterbank, lane == mel layout: G = ceil(M/32) groups of 32 consecutive mel rows
  const int* mel_start;   // [G*32] first bin of each row (0 for rows >= M)
  const int* mel_ginfo;   // [2*G]  (group max length, offset of the group's weights in floats)
  const float* mel_wg;    // [sum_g glen[g]*32]  W[g][j][lane], zero padded
  int mel_groups, mel_wg_count;
  int tiles_per_clip;
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Fire-and-forget float max (no read-back, so the issuing warp never waits on an HBM round trip):
// non-negative floats order like signed ints, negative floats order inversely as unsigned ints.
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
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
__device__ __forceinline__ void fill_tile(const FastParams& p, float* xs, unsigned tile) {
  const int clip_i = (int)(tile / (unsigned)p.tiles_per_clip);
  const int tile_i = (int)(tile - (unsigned)clip_i * (unsigned)p.tiles_per_clip);
  const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
  const int64_t lt0 = (int64_t)tile_i * C::FT;
  const int64_t q0 = (p.frame_begin + lt0) * C::HOP;  // padded coordinate of the tile's first sample
  const int64_t s0 = q0 - p.geo.pad_left;             // source coordinate
  const bool interior = p.fast_fill_ok && s0 >= 0 && (s0 + C::SPAN) <= p.valid_length && s0 >= p.sample_offset;
  if (interior && p.preemph != 0.0f && s0 > p.sample_offset) {
    // pre-emphasis on the way in: y[n] = x[n] - a*x[n-1] with separately rounded multiply and subtract
    // (bit-exact vs the reference's `x[1:] - a*x[:-1]`); vectorised, coalesced direct loads
    const float* src = clip + (s0 - p.sample_offset);
    const float a = p.preemph;
    for (int j = threadIdx.x; j < C::SPAN / 2; j += C::THREADS) {
      const int s = 2 * j;
      const float2 x = __ldg(reinterpret_cast<const float2*>(src + s));
      const float xm = __ldg(src + s - 1);
      const int row = s / C::HOP, col = s - row * C::HOP;
      *reinterpret_cast<float2*>(xs + row * C::P + col) =
          make_float2(__fsub_rn(x.x, __fmul_rn(a, xm)), __fsub_rn(x.y, __fmul_rn(a, x.x)));
    }
  } else if (interior && p.preemph == 0.0f) {
    // thread t copies the 8-byte pair (row r0 + RPI*i, column 2*c): both addresses are linear in i
    constexpr int PPR = C::HOP / 2;             // pairs per row
    constexpr int RPI = C::THREADS / PPR;       // rows per iteration
    constexpr int TOTAL_ROWS = (C::SPAN + C::HOP - 1) / C::HOP;
    constexpr int TAIL = C::SPAN - (TOTAL_ROWS - 1) * C::HOP;  // samples in the last (partial) row
    const int r0 = threadIdx.x / PPR, c = threadIdx.x - r0 * PPR;
    if (r0 < RPI) {
      const float* src = clip + (s0 - p.sample_offset) + r0 * C::HOP + 2 * c;
      float* dst = xs + r0 * C::P + 2 * c;
#pragma unroll
      for (int i = 0; i < (TOTAL_ROWS + RPI - 1) / RPI; ++i) {
        const int row = r0 + RPI * i;
        if (row < TOTAL_ROWS - 1 || (row == TOTAL_ROWS - 1 && 2 * c < TAIL))
          cp_async8(dst + i * RPI * C::P, src + i * RPI * C::HOP);
      }
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
  using namespace regs;
  const float2 e = pfma(zm, make_float2(1.0f, -1.0f), zk);                 // (zk.x + zm.x, zk.y - zm.y)
  const float2 o = pfma(pswap(zk), make_float2(1.0f, -1.0f), pswap(zm));   // (zk.y + zm.y, zm.x - zk.x)
  const float2 t = cmul(o, w);
  const float2 a = padd(e, t), b = psub(e, t);
  pk = fmaf(a.y, a.y, a.x * a.x);
  pm = fmaf(b.y, b.y, b.x * b.x);
}

// L taps of one mel group for NFW frames with every load issued before the first FMA (one shared-memory
// latency per <= 8 taps instead of one per tap: the mel phase is latency-, not throughput-bound)
template <int L, int NFW, int ROWSTEP, int WSTR>
__device__ __forceinline__ void mel_group_taps(const float* wp, const float* pq, float (&acc)[NFW]) {
  static_for<0, (L + 7) / 8>([&](auto C_) {
    constexpr int c0 = decltype(C_)::value * 8;
    constexpr int CL = (L - c0) < 8 ? (L - c0) : 8;
    float w[CL], pv[NFW][CL];
#pragma unroll
    for (int j = 0; j < CL; ++j) w[j] = wp[(c0 + j) * WSTR];
#pragma unroll
    for (int i = 0; i < NFW; ++i)
#pragma unroll
      for (int j = 0; j < CL; ++j) pv[i][j] = pq[i * ROWSTEP + c0 + j];
#pragma unroll
    for (int j = 0; j < CL; ++j)
#pragma unroll
      for (int i = 0; i < NFW; ++i) acc[i] = fmaf(pv[i][j], w[j], acc[i]);
  });
}

template <class C>
struct Smem {  // section offsets in float4 units from the 16-byte aligned dynamic smem base
  static constexpr int cdiv4(int bytes) { return (bytes + 15) / 16; }
  static constexpr int WIN = 0;
  static constexpr int TW1 = WIN + cdiv4(8 * C::NC);
  static constexpr int TWP = TW1 + cdiv4(8 * C::NC);
  static constexpr int EX = TWP + cdiv4(8 * C::NC);
  static constexpr int PW = EX + cdiv4(8 * C::FT * C::EP);
  static constexpr int XS = C::ALIAS ? PW : PW + cdiv4(4 * C::FT * C::PP);
  static constexpr int PX_END = C::ALIAS ? PW + (cdiv4(4 * C::FT * C::PP) > cdiv4(4 * C::XS_FLOATS) ? cdiv4(4 * C::FT * C::PP)
                                                                                                  : cdiv4(4 * C::XS_FLOATS))
                                         : XS + cdiv4(4 * C::XS_FLOATS);
  static constexpr int DYN = PX_END;  // then: sums (double), mel weights, starts, group info
};

template <class C, bool LAYOUT_TM, bool WANT_SUMS>
__global__ void __launch_bounds__(C::THREADS, C::MIN_BLOCKS) fast_logmel_kernel(const FastParams p) {
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC;
  using S = Smem<C>;
  extern __shared__ float4 smem4[];
  float2* const s_win2 = reinterpret_cast<float2*>(smem4 + S::WIN);  // [N2][N1]
  float2* const s_tw1 = reinterpret_cast<float2*>(smem4 + S::TW1);   // [N2][N1]
  float2* const s_twp = reinterpret_cast<float2*>(smem4 + S::TWP);   // [N1/2][2*N2]
  float2* const E = reinterpret_cast<float2*>(smem4 + S::EX);        // [FT][EP]
  float* const Pw = reinterpret_cast<float*>(smem4 + S::PW);         // [FT][PP]
  float* const xs = reinterpret_cast<float*>(smem4 + S::XS);         // [ROWS][P]
  const int M = p.n_mels;
  const int G = p.mel_groups;
  double* const s_sums = reinterpret_cast<double*>(smem4 + S::DYN);   // [2*G*8]
  float* const s_wg = reinterpret_cast<float*>(s_sums + 2 * G * 8);   // [mel_wg_count]
  int* const s_start = reinterpret_cast<int*>(s_wg + p.mel_wg_count); // [G*8]
  int* const s_ginfo = s_start + G * 8 + ((G * 8) & 1);               // [2*G], 8-byte aligned
  __shared__ float red_max[C::WARPS], red_min[C::WARPS];
  __shared__ int s_cur_clip;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* const Y = reinterpret_cast<float*>(E);  // [M][33] staging for the (M, T) layout; aliases the exchange buffer

  for (int i = threadIdx.x; i < NC; i += C::THREADS) {
    s_win2[i] = p.win2[i];
    s_tw1[i] = p.tw1[i];
    s_twp[i] = p.twp[i];
  }
  for (int i = threadIdx.x; i < G * 8; i += C::THREADS) s_start[i] = p.mel_start[i];
  for (int i = threadIdx.x; i < 2 * G; i += C::THREADS) s_ginfo[i] = p.mel_ginfo[i];
  for (int i = threadIdx.x; i < p.mel_wg_count; i += C::THREADS) s_wg[i] = p.mel_wg[i];
  constexpr bool want_sums = WANT_SUMS;
  const bool want_max = p.clip_max != nullptr;
  if (want_sums)
    for (int i = threadIdx.x; i < 2 * G * 8; i += C::THREADS) s_sums[i] = 0.0;
  if (threadIdx.x == 0) s_cur_clip = -1;

  const unsigned total_tiles = (unsigned)p.batch * (unsigned)p.tiles_per_clip;
  const unsigned tpc = (unsigned)p.tiles_per_clip;
  unsigned tile = blockIdx.x;
  if (!C::ALIAS && tile < total_tiles) fill_tile<C>(p, xs, tile);

  const float guard_add = p.guard_add, guard_floor = p.guard_floor, y_mul = p.y_mul, y_add = p.y_add;
  const bool use_log = p.use_log != 0;
  const bool pw_only = p.spec_kind == B2A_SPEC_POWER;
  const float spec_eps = p.spec_eps;

  for (; tile < total_tiles; tile += gridDim.x) {
    const int clip_i = (int)(tile / tpc);
    const int tile_i = (int)(tile - (unsigned)clip_i * tpc);
    const int64_t lt0 = (int64_t)tile_i * C::FT;
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);

    if (C::ALIAS) {
      __syncthreads();  // previous tile's mel phase has finished reading P (which shares xs' memory)
      fill_tile<C>(p, xs, tile);
    }
    cp_async_wait_all();
    __syncthreads();  // xs ready; previous tile's Y fully written out

    // per-CTA running per-mel sums: flush when the clip changes
    if (want_sums && s_cur_clip != clip_i) {
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
    const int dbg = p.debug_skip;
#pragma unroll 1
    for (int rr = 0; rr < ((dbg & 1) ? 0 : C::RPW); ++rr) {
      const int n2 = warp * C::RPW + rr;
      const float* xb = xs + lane * C::P + 2 * n2;
      const float4* wb4 = reinterpret_cast<const float4*>(s_win2 + n2 * N1);
      float2 v[N1];
      static_for<0, N1 / 2>([&](auto I_) {
        constexpr int n1 = 2 * decltype(I_)::value;
        constexpr int off0 = (n1 / C::K) * C::P + (n1 % C::K) * 2 * N2;
        constexpr int off1 = ((n1 + 1) / C::K) * C::P + ((n1 + 1) % C::K) * 2 * N2;
        const float2 x0 = *reinterpret_cast<const float2*>(xb + off0);
        const float2 x1 = *reinterpret_cast<const float2*>(xb + off1);
        const float4 w = wb4[n1 / 2];
        v[n1] = regs::pmul(x0, make_float2(w.x, w.y));
        v[n1 + 1] = regs::pmul(x1, make_float2(w.z, w.w));
      });
      Dft<N1>::run(v);
      const float4* tb4 = reinterpret_cast<const float4*>(s_tw1 + n2 * N1);
      float2* eb = E + lane * C::EP + n2;
      static_for<0, N1 / 2>([&](auto I_) {
        constexpr int k1 = 2 * decltype(I_)::value;
        constexpr int slot0 = (k1 <= N1 / 2) ? k1 : (3 * N1 / 2 - k1);
        constexpr int slot1 = (k1 + 1 <= N1 / 2) ? (k1 + 1) : (3 * N1 / 2 - (k1 + 1));
        const float4 t = tb4[k1 / 2];
        float2 y0 = v[k1];
        if constexpr (k1 > 0) y0 = regs::cmul(y0, make_float2(t.x, t.y));
        const float2 y1 = regs::cmul(v[k1 + 1], make_float2(t.z, t.w));
        eb[slot0 * N2] = y0;
        eb[slot1 * N2] = y1;
      });
    }
    __syncthreads();  // E complete, xs free

    // prefetch the next tile's samples while stage 2 / mel run
    {
      const unsigned next = tile + gridDim.x;
      if (!C::ALIAS && next < total_tiles) fill_tile<C>(p, xs, next);
    }

    // ---- stage 2 ----------------------------------------------------------------------------------------
    if (!(dbg & 2)) {
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
      auto emit = [&](int k, float v) { pr[k] = pw_only ? v : sqrtf(v + spec_eps); };
      const float2* tw = s_twp + u * 2 * N2;
      if (u != 0) {
        const float4* tw4 = reinterpret_cast<const float4*>(tw);
        static_for<0, N2 / 2>([&](auto I_) {
          constexpr int k2 = 2 * decltype(I_)::value;
          const float4 t = tw4[k2 / 2];
          float pk, pm;
          post_pair(A[k2], B[N2 - 1 - k2], make_float2(t.x, t.y), pk, pm);
          emit(u + N1 * k2, pk);
          emit(NC - (u + N1 * k2), pm);
          post_pair(A[k2 + 1], B[N2 - 2 - k2], make_float2(t.z, t.w), pk, pm);
          emit(u + N1 * (k2 + 1), pk);
          emit(NC - (u + N1 * (k2 + 1)), pm);
        });
      } else {  // unit 0: [0..N2/2] column 0, [N2 .. N2+N2/2) column N1/2
        static_for<0, N2 / 2 + 1>([&](auto I_) {  // column 0: k = N1*k2 <-> Nc - k = N1*(N2-k2)
          constexpr int k2 = decltype(I_)::value;
          float pk, pm;
          post_pair(A[k2], A[(N2 - k2) % N2], tw[k2], pk, pm);
          emit(N1 * k2, pk);
          emit(NC - N1 * k2, pm);
        });
        static_for<0, N2 / 2>([&](auto I_) {  // column N1/2: k = N1/2 + N1*k2 <-> N1/2 + N1*(N2-1-k2)
          constexpr int k2 = decltype(I_)::value;
          float pk, pm;
          post_pair(B[k2], B[N2 - 1 - k2], tw[N2 + k2], pk, pm);
          emit(N1 / 2 + N1 * k2, pk);
          emit(NC - (N1 / 2 + N1 * k2), pm);
        });
      }
    }
    __syncthreads();  // Pw complete, E free (Y aliases E)

    if (!(dbg & 4))
    // ---- mel projection + log + affine: LANE = (4 frames) x (8 consecutive mel rows) -----------------------------
    // The filterbank is banded (<= 2 non-zeros per bin): a mel row is a short run of taps.  A warp-instruction
    // covers 8 consecutive rows for 4 frames, so (a) the P gathers touch ~32 distinct banks (row pitch == 9 mod 32,
    // neighbouring rows start a few bins apart), (b) rows are zero-padded only to the longest of 8 neighbours,
    // (c) each store instruction writes four fully used 32-byte sectors of the (T, M) output.  A work item is
    // (octet of rows, half of the tile's frames): 4 independent accumulators per lane share one weight load.
    {
      constexpr int NQ = 4;                       // frame quads per item
      constexpr int ROWSTEP = 4 * C::PP;          // P rows of consecutive quads
      const int ms = lane & 7, fs = lane >> 3;
      float lmax = -INFINITY, lmin = INFINITY;
      float* const o = p.out + (int64_t)clip_i * p.out_clip_stride;
      const int2* ginfo2 = reinterpret_cast<const int2*>(s_ginfo);
      // output addressing: 32-bit element offsets from a per-tile base; (T, M): (f0 + 4q)*M + m, (M, T) staging:
      // m*33 + f0 + 4q
      float* const obase = LAYOUT_TM ? (o + lt0 * M) : Y;
      const int qstep = LAYOUT_TM ? 4 * M : 4;
      const int fstep = LAYOUT_TM ? M : 1, mstep = LAYOUT_TM ? 1 : 33;
      const bool full = nf == C::FT && (M & 7) == 0 && !(dbg & 16);  // every slot valid: no per-output predicates
      auto item = [&](auto FULL_, int it) {
        constexpr bool FULL = decltype(FULL_)::value;
        const int oct = it >> 1, half = it & 1;
        const int m = oct * 8 + ms;
        const int f0 = half * 16 + fs;            // this lane's frames: f0 + 4q
        const int2 gi = ginfo2[oct];              // (octet length, weight offset)
        const float* wp = s_wg + gi.y + ms;
        const float* pq = Pw + f0 * C::PP + s_start[m];
        float acc[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) acc[i] = 0.0f;
        switch (gi.x) {  // one dispatch per item, taps fully unrolled with all loads issued up front
          case 0: break;
#define B2A_MEL_CASE(LL) case LL: mel_group_taps<LL, NQ, ROWSTEP, 8>(wp, pq, acc); break;
          B2A_MEL_CASE(1) B2A_MEL_CASE(2) B2A_MEL_CASE(3) B2A_MEL_CASE(4) B2A_MEL_CASE(5) B2A_MEL_CASE(6)
          B2A_MEL_CASE(7) B2A_MEL_CASE(8) B2A_MEL_CASE(9) B2A_MEL_CASE(10) B2A_MEL_CASE(11) B2A_MEL_CASE(12)
          B2A_MEL_CASE(13) B2A_MEL_CASE(14) B2A_MEL_CASE(15) B2A_MEL_CASE(16)
#undef B2A_MEL_CASE
          default:
#pragma unroll 1
            for (int j = 0; j < gi.x; ++j) {
              const float w = wp[j * 8];
#pragma unroll
              for (int i = 0; i < NQ; ++i) acc[i] = fmaf(pq[i * ROWSTEP + j], w, acc[i]);
            }
        }
        float* const op = obase + (f0 * fstep + m * mstep);
        const bool mok = m < M;
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float a = fmaxf(acc[i] + guard_add, guard_floor);
          float y = use_log ? lg2_approx(a) : a;
          y = fmaf(y, y_mul, y_add);
          if (FULL) {
            lmax = fmaxf(lmax, y);
            lmin = fminf(lmin, y);
            op[i * qstep] = y;
            if (WANT_SUMS) {
              d1 += (double)y;
              d2 += (double)y * (double)y;
            }
          } else {
            const bool ok = mok && (f0 + 4 * i < nf);
            const float yv = ok ? y : __int_as_float(0x7fc00000);  // NaN is ignored by fmaxf / fminf
            lmax = fmaxf(lmax, yv);
            lmin = fminf(lmin, yv);
            if (ok && (!(dbg & 16) || y == 1234.5678f)) {
              op[i * qstep] = y;
              if (WANT_SUMS) {
                d1 += (double)y;
                d2 += (double)y * (double)y;
              }
            }
          }
        }
        if (WANT_SUMS) {  // fold the 4 frame-sub lanes of each mel row, then one shared-memory atomic per row
          d1 += __shfl_xor_sync(0xffffffffu, d1, 8);
          d2 += __shfl_xor_sync(0xffffffffu, d2, 8);
          d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
          d2 += __shfl_xor_sync(0xffffffffu, d2, 16);
          if (fs == 0 && mok) {
            atomicAdd(&s_sums[2 * m], d1);
            atomicAdd(&s_sums[2 * m + 1], d2);
          }
        }
      };
      if (full) {
#pragma unroll 1
        for (int it = warp; it < 2 * G; it += C::WARPS) item(std::true_type{}, it);
      } else {
#pragma unroll 1
        for (int it = warp; it < 2 * G; it += C::WARPS) item(std::false_type{}, it);
      }
      if (want_max) {
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
          lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o2));
          lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o2));
        }
        if (lane == 0) {
          red_max[warp] = lmax;
          red_min[warp] = lmin;
        }
      }
      if (want_max || !LAYOUT_TM) __syncthreads();
      if (want_max && threadIdx.x == 0) {
        float a = red_max[0], b = red_min[0];
#pragma unroll
        for (int w = 1; w < C::WARPS; ++w) {
          a = fmaxf(a, red_max[w]);
          b = fminf(b, red_min[w]);
        }
        atomic_max_f(p.clip_max + clip_i, a);
        p.tile_min[(int64_t)clip_i * tpc + tile_i] = b;
      }
      if (!LAYOUT_TM && lane < nf)
        for (int m = warp; m < M; m += C::WARPS) o[(int64_t)m * p.frame_count + lt0 + lane] = Y[m * 33 + lane];
    }
  }
  cp_async_wait_all();
  if (want_sums) {
    __syncthreads();
    const int prev = s_cur_clip;
    if (prev >= 0)
      for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) atomicAdd(p.feat_sums + (int64_t)prev * 2 * M + i, s_sums[i]);
  }
}

template <class C>
size_t smem_bytes(int G, int wg_count) {
  return (size_t)16 * Smem<C>::DYN + sizeof(double) * 2 * G * 8 + sizeof(float) * (wg_count + (wg_count & 1)) + sizeof(int) * (G * 8 + 2 * G + 2) + 16;
}

struct FastState {
  float2* d_win2 = nullptr;
  float2* d_tw1 = nullptr;
  float2* d_twp = nullptr;
  int* d_start = nullptr;
  int* d_ginfo = nullptr;
  float* d_wg = nullptr;
  int groups = 0, wg_count = 0;
  int variant = 0;  // 1: 400/160, 2: 512/160, 3: 1024/256
};

template <class C, bool TM, bool SUMS>
int launch_variant(b2a_plan* plan, FastParams& p, size_t smem, int grid, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    B2A_CUDA(cudaFuncSetAttribute(fast_logmel_kernel<C, TM, SUMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  fast_logmel_kernel<C, TM, SUMS><<<grid, C::THREADS, smem, st>>>(p);
  B2A_CUDA(cudaGetLastError());
  return B2A_OK;
}

template <class C>
int launch(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) {
  const size_t smem = smem_bytes<C>(p.mel_groups, p.mel_wg_count);
  if (smem > 226 * 1024) {
    set_error("fast kernel: %zu bytes of shared memory needed", smem);
    return B2A_ERR_UNSUPPORTED;
  }
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  if (tiles >= (int64_t)1 << 31) {
    set_error("fast kernel: %lld tiles in one launch (split the batch)", (long long)tiles);
    return B2A_ERR_UNSUPPORTED;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, C::MIN_BLOCKS));
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  const bool tm = p.out_layout == B2A_LAYOUT_TM, sums = p.feat_sums != nullptr;
  if (tm && !sums) return launch_variant<C, true, false>(plan, p, smem, grid, st);
  if (tm && sums) return launch_variant<C, true, true>(plan, p, smem, grid, st);
  if (!tm && !sums) return launch_variant<C, false, false>(plan, p, smem, grid, st);
  return launch_variant<C, false, true>(plan, p, smem, grid, st);
}

using Cfg400 = Cfg<20, 10, 160, false, 2>;   // Whisper / Voxtral-RT / S3Tokenizer: 320 threads, 2 CTAs / SM
using Cfg512 = Cfg<16, 16, 160, true, 2>;    // Parakeet / Sortformer: 256 threads, 2 CTAs / SM
using Cfg1024 = Cfg<32, 16, 256, true, 1>;   // Vocos / Qwen3-TTS mel: 512 threads, 1 CTA / SM

}  // namespace

bool fast_frontend_supported(const b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  if (d.n_mels <= 0 || d.spec_kind == B2A_SPEC_COMPLEX) return false;
  if (d.affine_div < 0.0f) return false;
  const bool v400 = d.n_fft == 400 && d.hop == 160;
  const bool v512 = d.n_fft == 512 && d.hop == 160;
  const bool v1024 = d.n_fft == 1024 && d.hop == 256;
  if (!(v400 || v512 || v1024)) return false;
  if (d.n_mels > 256) return false;
  return true;
}

int fast_frontend_init(b2a_plan* plan) {
  const b2a_frontend_desc& d = plan->fd;
  FastState* fs = new FastState();
  plan->fast = fs;
  int N1, N2;
  if (d.n_fft == 400) { fs->variant = 1; N1 = 20; N2 = 10; }
  else if (d.n_fft == 512) { fs->variant = 2; N1 = 16; N2 = 16; }
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
  for (int k2 = 0; k2 <= N2 / 2; ++k2) twp[k2] = wn(N1 * k2);                    // unit 0, column 0
  for (int k2 = 0; k2 < N2 / 2; ++k2) twp[N2 + k2] = wn(N1 / 2 + N1 * k2);        // unit 0, column N1/2
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
  plan->kernel_name = fs->variant == 1 ? "fast_logmel_400x160" : (fs->variant == 2 ? "fast_logmel_512x160" : "fast_logmel_1024x256");
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

int fast_frontend_partial(b2a_plan* plan, const b2a_forward_args* a, float* clip_max, float* tile_min,
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
  p.fast_fill_ok = (reinterpret_cast<uintptr_t>(a->audio) % 8 == 0) && (a->clip_stride % 2 == 0) &&
                   (p.geo.pad_left % 2 == 0) && (a->sample_offset % 2 == 0);
  p.debug_skip = getenv("B2A_SKIP") ? atoi(getenv("B2A_SKIP")) : 0;
  p.spec_kind = d.spec_kind;
  p.spec_eps = d.spec_kind == B2A_SPEC_SQRT_POWER_EPS ? d.spec_eps : 0.0f;
  p.n_mels = d.n_mels;
  p.guard_add = d.guard_kind == B2A_GUARD_ADD ? d.guard_eps : 0.0f;
  p.guard_floor = d.guard_kind == B2A_GUARD_MAX ? d.guard_eps : -INFINITY;
  p.use_log = d.log_kind != B2A_LOG_NONE;
  const double lscale = d.log_kind == B2A_LOG_LOG10 ? 0.30102999566398119521 : (d.log_kind == B2A_LOG_LN ? 0.69314718055994530942 : 1.0);
  if (d.affine_div != 0.0f) {  // ((log2(a) * lscale) + add) / div
    p.y_mul = (float)(lscale / (double)d.affine_div);
    p.y_add = (float)((double)d.affine_add / (double)d.affine_div);
  } else {
    p.y_mul = (float)lscale;
    p.y_add = 0.0f;
  }
  p.out_layout = d.out_layout;
  p.out = reinterpret_cast<float*>(a->out);
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : a->frame_count * d.n_mels;
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
  if (fs->variant == 1) return launch<Cfg400>(plan, fs, p, st);
  if (fs->variant == 2) return launch<Cfg512>(plan, fs, p, st);
  return launch<Cfg1024>(plan, fs, p, st);
}

}  // namespace b2a
