// b200audio — specialised fused log-mel front-end kernels ("fast" family): shared kernel template, included by the
// per-variant translation units fast_400.cu / fast_512.cu / fast_1024.cu (compiled in parallel by csrc/build.py).
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
#pragma once
#include <algorithm>
#include <stdlib.h>

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cooperative_groups.h>
#include "common.cuh"
#include "fft_regs.cuh"
#include "mel_gen.cuh"

#ifndef B2A_X_TWPF2
#define B2A_X_TWPF2 2  // stage 2: post-twiddle loads, same idea
#endif
#ifndef B2A_X_TWPF
#define B2A_X_TWPF 3  // stage 1: inter-stage twiddle loads issued this many pairs ahead of their use (0 = compiler's order)
#endif

namespace b2a {

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
  int n_mels;
  float guard_add, guard_floor;  // a = max(a + guard_add, guard_floor)   (ADD: (eps, -inf); MAX: (0, eps))
  int use_log;                   // y = log2(a) if use_log else a
  float y_mul, y_add;            // y' = y * y_mul + y_add   (log base change and the affine map folded together)
  int out_layout;
  int out_dtype;  // B2A_DTYPE_*: float16 / bfloat16 features straight from phase B (generated-mel 400/160 kernels)
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
  int tile_min_pitch;  // tiles per clip in the tile_min table (>= tiles_per_clip when trailing all-padding tiles are skipped)
  // single-launch forward (fast_logmel_tma_kernel<..., FUSED>, cooperative launch): per-tile maxima instead of atomics on an
  // initialised clip_max, then a grid-wide barrier and the clamp fix-up by the same CTAs (one CTA per clip)
  float* tile_max;       // [batch][tile_min_pitch]; non-null only in the fused launch
  float clamp_delta;     // floor = clip max - clamp_delta (affine domain)
  int skip_floor_tiles;  // 1: a tile whose every value is the guard-floor constant (digital silence) is NOT stored; its
                         // tile_min entry is -inf and the clamp fix-up writes max(c, floor) there (fast_logmel_tma_kernel)
  long long* dbg_clk;  // profiling aid (B2A_CLOCKS=file): per CTA, cycles accumulated per phase [8] (thread 0's view)
};

struct FastState {
  float2* d_win2 = nullptr;
  float2* d_tw1 = nullptr;
  float2* d_twp = nullptr;
  int* d_start = nullptr;
  int* d_ginfo = nullptr;
  float* d_wg = nullptr;
  int groups = 0, wg_count = 0;
  int variant = 0;  // 1: 400/160, 2: 512/160, 3: 1024/256, 4: 800/200, 5: 1024/320
  int spec = 0;     // index into the variant's generated mel specs (0: run-time tables)
  const char* spec_name = nullptr;
};

// per-variant translation units (fast_400.cu, fast_512.cu, fast_1024.cu): filterbank matching and launch
int fast_match_400(const b2a_plan* plan, const char** name);
int fast_match_512(const b2a_plan* plan, const char** name);
int fast_match_1024(const b2a_plan* plan, const char** name);
int fast_launch_400(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st);
int fast_launch_512(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st);
int fast_launch_1024(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st);
int fast_match_800(const b2a_plan* plan, const char** name);
int fast_launch_800(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st);
int fast_match_1024h320(const b2a_plan* plan, const char** name);
int fast_launch_1024h320(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st);

namespace {

using regs::Dft;
using regs::static_for;

struct NoSpec {  // run-time mel tables (any filterbank); melgen::MelSpec_* bake a named filterbank into code
  static constexpr int M = 0, NW = 0, F = 0;
  template <class C, class Emit>
  static __device__ __forceinline__ void run(int, const float*, Emit&&) {}
};

// CP_: pitch (float2 slots) of one column of the exchange buffer.  N2 for the log-mel kernels; N2 + 1 (odd) for the
// complex-spectrum kernel, whose copy-out reads consecutive BINS = consecutive columns: an odd pitch keeps those LDS.64
// bank-conflict free (with pitch N2 = 10 or 16 they were 2-way / 8-way conflicts).
template <int N1_, int N2_, int HOP_, bool ALIAS_, int MIN_BLOCKS_, int CP_ = N2_>
struct Cfg {
  using Padded = Cfg<N1_, N2_, HOP_, ALIAS_, MIN_BLOCKS_, N2_ + 1>;
  static constexpr int N1 = N1_, N2 = N2_, HOP = HOP_, CP = CP_;
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
  static constexpr int XS_HEAD = 4;  // floats in front of the sample tile: [-4] raw-samples flag, [-3] sample before the span
  static constexpr int XS_FLOATS = ROWS * P + XS_HEAD;
  static constexpr int EP = N1 * CP + 1;     // exchange pitch per frame (float2), odd
  // power pitch per frame (floats): odd (stage-2 lane==frame stores are conflict free) and == 9 (mod 32) so that
  // the mel phase's (4 frames x 8 mel rows) gathers land in distinct banks
  static constexpr int PP = F + ((9 - F % 32 + 32) % 32);
  static constexpr int K = HOP / (2 * N2);   // taps pairs per row per role step
  // Slot (float2 index within a frame's exchange row) of power bin k when the power tile is written IN PLACE over the
  // exchange buffer (generated-mel kernels): a stage-2 unit only overwrites slots of the two columns it has just
  // consumed — unit u: bin u + N1*s -> slot u*N2 + s, its mirror Nc - k -> slot (N1/2 + u)*N2 + s; unit 0 keeps its
  // slot order (see stage 2), parks DC in the duplicate slot of its self-paired bin Nc/2 and Nyquist in the pad slot.
  __host__ __device__ static constexpr int sig(int k) {
    if (k == 0) return (N1 / 2) * CP + N2 / 2 - 1;
    if (k == NC) return N1 * CP;
    const int r = k % N1, q = k / N1;
    if (r == 0) return q <= N2 / 2 ? (q - 1) : (N1 / 2) * CP + (N2 - 1 - q);
    if (r == N1 / 2) return q < N2 / 2 ? (N2 / 2 + q) : (N1 / 2) * CP + (N2 / 2 + (N2 - 1 - q));
    if (r < N1 / 2) return r * CP + q;
    const int u = N1 - r;
    return (N1 / 2 + u) * CP + (NC - k - u) / N1;
  }
  // stage-1 constant tables in tensor memory (window + inter-stage twiddles, 4 * N1 columns per role): each warp owns
  // RPW roles; warps w, w + 4, ... share a lane quadrant and take successive column blocks.  The allocation is a power of two
  // and every resident CTA needs its own.
  static constexpr int TM_PER_WARP = 4 * N1 * RPW;
  static constexpr int TM_NEED = TM_PER_WARP * ((WARPS + 3) / 4);
  static constexpr int TM_COLS = TM_NEED <= 32 ? 32 : TM_NEED <= 64 ? 64 : TM_NEED <= 128 ? 128 : TM_NEED <= 256 ? 256 : 512;
  static constexpr bool TM_OK = N1 % 4 == 0 && TM_NEED <= 512 && TM_COLS * MIN_BLOCKS <= 512;
  static_assert(N1 % 2 == 0 && N2 % WARPS == 0, "role split");
  static_assert(EP % 2 == 1 && CP >= N2, "odd frame pitch; columns do not overlap");
  static_assert(HOP % (2 * N2) == 0, "hop must be a multiple of 2*N2");
  static_assert(NC % 2 == 0, "Nc even");
};


__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {  // FMNMX3 (sm_100+)
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// order-preserving float <-> signed-int key (an involution): lets REDUX.MIN/MAX.S32 reduce floats across a warp
__device__ __forceinline__ int float_key(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT-class approximation, flush-to-zero
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int BYTE_OFF>
__device__ __forceinline__ float4 lds128_at(unsigned base) {  // pinned (volatile) broadcast load: keeps its place in program order
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base), "n"(BYTE_OFF));
  return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- warp-uniform constant tables in TENSOR MEMORY ------------------------------------------------------------------
// Window taps, inter-stage twiddles and post-twiddles are the same for every lane of a warp (lane == frame), and a warp
// keeps its role for the whole kernel: each warp parks its tables once in its own TMEM columns (replicated over the 32
// lanes of its quadrant) and reads them back with tcgen05.ld (SASS LDTM) — a datapath of its own, so the ~520 broadcast
// LDS wavefronts per tile (20 % of the kernel's shared-memory traffic, the pipe that limits it) disappear from the
// shared-memory pipe (microbenchmark scratch/r2/ubench_tmem.cu: 2 x LDS.64 + LDS.128 broadcast 6.0 cycles per group and
// SM, 2 x LDS.64 + LDTM.x4 4.0).  tcgen05.wait::ld waits for ALL of the thread's outstanding loads: the loaded registers
// are passed through the wait statement ("+f") so that no use can be scheduled in front of it.
template <int N>
__device__ __forceinline__ void tm_ld(uint32_t taddr, float* r) {
  static_assert(N == 4 || N == 8, "chunk");
  if constexpr (N == 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "r"(taddr));
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void tm_wait(float* r) {
  if constexpr (N == 8)
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));
  else
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]));
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d));
}
// allocation: one warp allocates NCOLS columns (power of two >= 32) and publishes the base address through shared memory
template <int NCOLS>
__device__ __forceinline__ uint32_t tm_alloc(uint32_t* s_slot, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  return *reinterpret_cast<volatile uint32_t*>(s_slot);
}
template <int NCOLS>
__device__ __forceinline__ void tm_free(uint32_t base, int warp) {  // every warp of the CTA has finished its TMEM reads
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS));
}
// this warp's table columns: lanes of quadrant warp % 4, column block warp / 4 of `cols_per_warp` columns
__device__ __forceinline__ uint32_t tm_warp_base(uint32_t base, int warp, int cols_per_warp) {
  return base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cols_per_warp);
}
// parks n float2 constants (same values in every lane) at column `taddr`; n even
__device__ __forceinline__ void tm_store_table(uint32_t taddr, const float2* __restrict__ src, int n) {
  for (int j = 0; j < n; j += 2) {
    const float2 a = __ldg(src + j), b = __ldg(src + j + 1);
    tm_st4(taddr + 2 * j, a.x, a.y, b.x, b.y);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

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

// Per-thread constants of the interior fill path, computed once per kernel: thread t copies the 8-byte pair
// (row r0 + RPI*i, column 2*c) for i = 0..ITERS-1; both addresses are linear in i, so every copy is one LDGSTS
// with immediate offsets and only the last (partial) iteration is predicated.
template <class C>
struct FillCtx {
  static constexpr int PPR = C::HOP / 2;             // pairs per row
  static constexpr int RPI = C::THREADS / PPR;       // rows per iteration
  static constexpr int TOTAL_ROWS = (C::SPAN + C::HOP - 1) / C::HOP;
  static constexpr int TAIL = C::SPAN - (TOTAL_ROWS - 1) * C::HOP;  // samples in the last (partial) row
  static constexpr int ITERS = (TOTAL_ROWS + RPI - 1) / RPI;
  static_assert((TOTAL_ROWS - 1) / RPI == ITERS - 1, "exactly the last iteration is partial");
  const float* src0;  // this thread's first source pair of tile 0 of clip 0 (valid for interior tiles only)
  unsigned dst;       // shared-window byte address
  int lo, hi;         // tile indices [lo, hi] whose whole span is interior AND copyable raw (same for every clip)
  bool active, last_ok, raw_pre;
  static __device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {  // b > 0
    const int64_t q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
  }
  template <int PREK>
  __device__ __forceinline__ void init(const FastParams& p, unsigned xs_sa) {
    const int r0 = threadIdx.x / PPR, c = threadIdx.x - r0 * PPR;
    dst = xs_sa + 4u * (unsigned)(r0 * C::P + 2 * c);
    active = r0 < RPI;
    const int row = r0 + RPI * (ITERS - 1);
    last_ok = active && (row < TOTAL_ROWS - 1 || (row == TOTAL_ROWS - 1 && 2 * c < TAIL));
    // source coordinate of tile t's first sample: s0(t) = base + t * FT * HOP.  Interior: s0 >= max(0, sample_offset)
    // and s0 + SPAN <= valid_length; with pre-emphasis the raw copy also needs the sample in front of the span
    // (s0 > sample_offset) and the pitch HOP + 2 (see fill_tile).
    const bool pre = p.preemph != 0.0f;
    raw_pre = PREK != 0 && pre && C::P == C::HOP + 2;
    const int64_t base = p.frame_begin * C::HOP - p.geo.pad_left, T = (int64_t)C::FT * C::HOP;
    int64_t smin = p.sample_offset + (pre ? 1 : 0);
    if (smin < 0) smin = 0;
    int64_t l = -floor_div(-(smin - base), T);                    // ceil((smin - base) / T)
    int64_t h = floor_div(p.valid_length - C::SPAN - base, T);
    if (l < 0) l = 0;
    if (h > 0x3fffffff) h = 0x3fffffff;
    if (!p.fast_fill_ok || (pre && !raw_pre)) h = l - 1;
    if (h < l) { l = 1; h = 0; }
    lo = (int)l;
    hi = (int)h;
    src0 = p.audio + (base - p.sample_offset) + (r0 * C::HOP + 2 * c);
  }
};

// Copies the tile's sample span into shared memory (rows of HOP samples at pitch P).
template <class C, int PREK>
__device__ __forceinline__ void fill_tile(const FastParams& p, float* xs, const FillCtx<C>& fc, int clip_i, int tile_i,
                                          unsigned dst_off = 0) {  // dst_off: byte offset of `xs` from the buffer fc was set up for
  // Interior tiles are copied RAW with cp.async (asynchronous: the copy of the next tile overlaps stage 2 / mel).  With
  // pre-emphasis the filter y[n] = x[n] - a*x[n-1] is then applied by stage 1 as it reads (separately rounded multiply
  // and subtract, bit-exact vs the reference's `x[1:] - a*x[:-1]`); the one sample in front of the span goes to
  // xs[-3] — where "the last sample of the previous row" lives for every other row start (pitch = HOP + 2).
  // Whether a tile is interior depends on its index only (every clip has the same geometry): FillCtx::lo / hi.
  if (tile_i >= fc.lo && tile_i <= fc.hi) {
    const bool raw_pre = PREK != 0 && fc.raw_pre;
    const float* src = fc.src0 + ((int64_t)clip_i * p.clip_stride + (int64_t)tile_i * (C::FT * C::HOP));
    if (fc.active) {
#pragma unroll
      for (int i = 0; i < FillCtx<C>::ITERS - 1; ++i)
        cp_async8(fc.dst + dst_off + 4u * (unsigned)(i * FillCtx<C>::RPI * C::P), src + i * FillCtx<C>::RPI * C::HOP);
    }
    if (fc.last_ok) {
      constexpr int i = FillCtx<C>::ITERS - 1;
      cp_async8(fc.dst + dst_off + 4u * (unsigned)(i * FillCtx<C>::RPI * C::P), src + i * FillCtx<C>::RPI * C::HOP);
    }
    if (threadIdx.x == 0) {
      reinterpret_cast<int*>(xs)[-4] = raw_pre ? 1 : 0;
      // thread 0's first pair is the span's first sample: the one in front of it is src[-1]
      if (raw_pre) cp_async4((unsigned)__cvta_generic_to_shared(xs - 3), src - 1);
    }
  } else {
    const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
    const int64_t lt0 = (int64_t)tile_i * C::FT;
    const int64_t q0 = (p.frame_begin + lt0) * C::HOP;  // padded coordinate of the tile's first sample
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);
    const int need = (nf - 1) * C::HOP + C::N;
    if (threadIdx.x == 0) reinterpret_cast<int*>(xs)[-4] = 0;  // this path stores finished (pre-emphasised) samples
    // Edge tiles (clip start / end: 2 per clip): sample PAIRS that lie inside the signal still go through cp.async —
    // asynchronous like an interior tile — and only the padded / reflected / virtual part (a few hundred samples) is
    // computed with ordinary loads.  With pre-emphasis every sample has to be computed (the padding reflects the
    // FILTERED signal), so the whole tile takes the scalar path.
    const bool pairs_ok = p.fast_fill_ok && p.preemph == 0.0f;
    const int64_t s0 = q0 - p.geo.pad_left;
    const int64_t smin = p.sample_offset > 0 ? p.sample_offset : 0;
    const unsigned xs_sa = (unsigned)__cvta_generic_to_shared(xs);
    auto sample = [&](int j) {
      float v = 0.0f;
      if (j < need) {
        const int64_t s = source_index(p.geo, p.pad_mode, q0 + j);
        if (s >= 0) v = fetch_sample_f(p, clip, s);
      }
      return v;
    };
    static_assert(C::SPAN % 2 == 0 && C::HOP % 2 == 0 && C::P % 2 == 0, "pairs stay inside a row");
    for (int i = threadIdx.x; i < C::SPAN / 2; i += C::THREADS) {
      const int e = 2 * i, row = e / C::HOP, col = e - row * C::HOP;
      const int64_t s = s0 + e;
      if (pairs_ok && e + 1 < need && s >= smin && s + 1 < p.valid_length) {
        cp_async8(xs_sa + 4u * (unsigned)(row * C::P + col), clip + (s - p.sample_offset));
      } else {
        *reinterpret_cast<float2*>(xs + row * C::P + col) = make_float2(sample(e), sample(e + 1));
      }
    }
  }
  cp_async_commit();
}

// Pre-emphasis tiles are filled with ordinary loads (the copy has to compute), so their HBM latency is exposed.  Pull
// the span of the tile AFTER the next one into L2 early: one prefetch per 128-byte line.
template <class C>
__device__ __forceinline__ void prefetch_span_l2(const FastParams& p, int clip_i, int tile_i) {
  if (p.preemph == 0.0f || clip_i >= p.batch) return;
  const float* clip = p.audio + (int64_t)clip_i * p.clip_stride;
  const int64_t s0 = (p.frame_begin + (int64_t)tile_i * C::FT) * C::HOP - p.geo.pad_left;
  const int64_t i = s0 + 32 * (int64_t)threadIdx.x;  // 32 floats = 128 bytes
  if (32 * (int)threadIdx.x < C::SPAN && i >= p.sample_offset && i < p.valid_length)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(clip + (i - p.sample_offset)));
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

// same, returning the spectrum itself: X[k] and X[Nc-k] = conj(E - W O)
__device__ __forceinline__ void post_pair_c(float2 zk, float2 zm, float2 w, float2& xk, float2& xm) {
  using namespace regs;
  const float2 e = pfma(zm, make_float2(1.0f, -1.0f), zk);
  const float2 o = pfma(pswap(zk), make_float2(1.0f, -1.0f), pswap(zm));
  const float2 t = cmul(o, w);
  xk = padd(e, t);
  xm = pmul(psub(e, t), make_float2(1.0f, -1.0f));
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

// YP: pitch (floats) of the generated-mel staging tile Y[frame][YP]; 0 = run-time-table kernel.
// Generated-mel kernels (YP > 0): the power tile lives IN PLACE in the exchange buffer (Cfg::sig), Y and the sample
// tile have their own space, and the next tile's samples are always prefetched.  Run-time-table kernels keep a
// separate power tile (natural bin order), alias Y with the exchange buffer and, for ALIAS configs, the power tile
// with the sample tile (no prefetch).
template <class C, int YP>
struct Smem {  // section offsets in float4 units from the 16-byte aligned dynamic smem base
  static constexpr int cdiv4(int bytes) { return (bytes + 15) / 16; }
  static constexpr bool PIE = YP > 0;
  static constexpr bool PREFETCH = PIE || !C::ALIAS;
  static constexpr int WIN = 0;
  static constexpr int TW1 = WIN + cdiv4(8 * C::NC);
  static constexpr int TWP = TW1 + cdiv4(8 * C::NC);
  static constexpr int EX = TWP + cdiv4(8 * C::NC);
  static constexpr int PW = EX + cdiv4(8 * C::FT * C::EP);   // PIE: the staging tile Y starts here
  static constexpr int PW_SIZE = PIE ? cdiv4(4 * C::FT * YP) : cdiv4(4 * C::FT * C::PP);
  static constexpr int XS = PREFETCH ? PW + PW_SIZE : PW;
  static constexpr int PX_END = PREFETCH ? XS + cdiv4(4 * C::XS_FLOATS)
                                         : PW + (PW_SIZE > cdiv4(4 * C::XS_FLOATS) ? PW_SIZE : cdiv4(4 * C::XS_FLOATS));
  static constexpr int DYN = PX_END;  // then: sums (double), mel weights, starts, group info
};

// ---- stage 1 of one tile: warp = residue n2 (RPW of them per warp), lane = frame -----------------------------------
// `preemph` != 0 with a RAW sample tile (flag at xs[-4]): y[n] = x[n] - a*x[n-1] is applied here, on the way in.
// PREK: 0 = the kernel instance never pre-emphasises (no code for it), 1 / -1 = decided at run time.
#ifndef B2A_X_CONTIG
#define B2A_X_CONTIG 0  // experiment: contiguous tile runs for every fast_logmel instance (default: only with per-feature sums)
#endif
#ifndef B2A_X_TMC
#define B2A_X_TMC 1  // stage-1 window / twiddle tables from tensor memory (0: shared-memory broadcasts) in fast_logmel_kernel
#endif
#ifndef B2A_X_CTAB
#define B2A_X_CTAB 0  // experiment: stage-1 window / inter-stage twiddle broadcasts from the constant bank (LDC) instead of shared memory
#endif
#if B2A_X_CTAB
__constant__ float2 c_win2[1024];
__constant__ float2 c_tw1[1024];
#endif

// TMC: window taps and inter-stage twiddles come from this warp's tensor-memory columns `tmc` — per role 2 * N1 window
// columns then 2 * N1 twiddle columns — instead of shared-memory broadcasts (s_win2 / s_tw1 unused).
template <class C, int PREK, bool TMC = false>
__device__ __forceinline__ void stage1_tile(const float* xs, float2* E, const float2* s_win2, const float2* s_tw1, int warp,
                                            int lane, float preemph, uint32_t tmc = 0) {
  constexpr int N1 = C::N1, N2 = C::N2;
  static_assert(!TMC || N1 % 4 == 0, "tensor-memory tables are read in chunks of four constants");
  const bool pre = PREK != 0 && preemph != 0.0f && reinterpret_cast<const int*>(xs)[-4] != 0;
#pragma unroll 1
  for (int rr = 0; rr < C::RPW; ++rr) {
    const int n2 = warp * C::RPW + rr;
    const float* xb = xs + lane * C::P + 2 * n2;
    const float4* wb4 = reinterpret_cast<const float4*>(s_win2 + n2 * N1);
    float2 v[N1];
    // the sample in front of a pair sits one float back — or, for the first pair of a row (n2 == 0 and column 0),
    // behind the row padding: P - HOP + 1 floats back
    const int back0 = n2 == 0 ? C::P - C::HOP + 1 : 1;
    [[maybe_unused]] float tmw[2][8];  // TMC: two chunks of four constants in flight
    [[maybe_unused]] const uint32_t tm_role = tmc + (uint32_t)(rr * 4 * N1);
    if constexpr (TMC) tm_ld<8>(tm_role, tmw[0]);
    static_for<0, N1 / 2>([&](auto I_) {
      constexpr int n1 = 2 * decltype(I_)::value;
      if constexpr (TMC && n1 % 4 == 0) {  // chunk n1 / 4 has landed; start the next one (after the last window chunk: twiddle chunk 0)
        constexpr int c = n1 / 4;
        tm_wait<8>(tmw[c & 1]);
        tm_ld<8>(tm_role + 8 * (c + 1), tmw[(c + 1) & 1]);
      }
      constexpr int off0 = (n1 / C::K) * C::P + (n1 % C::K) * 2 * N2;
      constexpr int off1 = ((n1 + 1) / C::K) * C::P + ((n1 + 1) % C::K) * 2 * N2;
      float2 x0 = *reinterpret_cast<const float2*>(xb + off0);
      float2 x1 = *reinterpret_cast<const float2*>(xb + off1);
      if (PREK != 0 && pre) {
        const float q0 = xb[off0 - ((n1 % C::K) == 0 ? back0 : 1)];
        const float q1 = xb[off1 - (((n1 + 1) % C::K) == 0 ? back0 : 1)];
        x0 = regs::psub(x0, regs::pmul(make_float2(q0, x0.x), make_float2(preemph, preemph)));
        x1 = regs::psub(x1, regs::pmul(make_float2(q1, x1.x), make_float2(preemph, preemph)));
      }
#if B2A_X_CTAB
      const float2 wa = c_win2[n2 * N1 + n1], wc = c_win2[n2 * N1 + n1 + 1];
      const float4 w = make_float4(wa.x, wa.y, wc.x, wc.y);
#else
      float4 w;
      if constexpr (TMC) {
        const float* t = tmw[(n1 / 4) & 1] + (n1 % 4) * 2;
        w = make_float4(t[0], t[1], t[2], t[3]);
      } else {
        w = wb4[n1 / 2];
      }
#endif
      v[n1] = regs::pmul(x0, make_float2(w.x, w.y));
      v[n1 + 1] = regs::pmul(x1, make_float2(w.z, w.w));
    });
    Dft<N1>::run(v);
    const float4* tb4 = reinterpret_cast<const float4*>(s_tw1 + n2 * N1);
    float2* eb = E + lane * C::EP + n2;
    if constexpr (TMC) {  // twiddle chunk c sits in tmw[(N1 / 4 + c) & 1] (chunk 0 was requested with the last window chunk)
      static_for<0, N1 / 2>([&](auto I_) {
        constexpr int k1 = 2 * decltype(I_)::value;
        constexpr int slot0 = (k1 <= N1 / 2) ? k1 : (3 * N1 / 2 - k1);
        constexpr int slot1 = (k1 + 1 <= N1 / 2) ? (k1 + 1) : (3 * N1 / 2 - (k1 + 1));
        constexpr int c = N1 / 4 + k1 / 4;
        if constexpr (k1 % 4 == 0) {
          tm_wait<8>(tmw[c & 1]);
          if constexpr (k1 + 4 < N1) tm_ld<8>(tm_role + 8 * (c + 1), tmw[(c + 1) & 1]);
        }
        const float* t = tmw[c & 1] + (k1 % 4) * 2;
        float2 y0 = v[k1];
        if constexpr (k1 > 0) y0 = regs::cmul(y0, make_float2(t[0], t[1]));
        const float2 y1 = regs::cmul(v[k1 + 1], make_float2(t[2], t[3]));
        eb[slot0 * C::CP] = y0;
        eb[slot1 * C::CP] = y1;
      });
      continue;
    }
#if B2A_X_TWPF > 0 && !B2A_X_CTAB
    // inter-stage twiddles: software-pipelined broadcast loads, TWD pairs ahead of their use (pinned with volatile asm:
    // left alone, ptxas issues each LDS.128 right in front of its FMUL2 and every pair waits out a shared-memory latency;
    // measured -1 %)
    constexpr int TWD = B2A_X_TWPF;
    const unsigned tb_sa = (unsigned)__cvta_generic_to_shared(tb4);
    float4 tq[N1 / 2];
    static_for<0, (TWD < N1 / 2 ? TWD : N1 / 2)>([&](auto I_) { tq[decltype(I_)::value] = lds128_at<16 * decltype(I_)::value>(tb_sa); });
#endif
    static_for<0, N1 / 2>([&](auto I_) {
      constexpr int k1 = 2 * decltype(I_)::value;
      constexpr int slot0 = (k1 <= N1 / 2) ? k1 : (3 * N1 / 2 - k1);
      constexpr int slot1 = (k1 + 1 <= N1 / 2) ? (k1 + 1) : (3 * N1 / 2 - (k1 + 1));
#if B2A_X_CTAB
      const float2 ta = c_tw1[n2 * N1 + k1], tc = c_tw1[n2 * N1 + k1 + 1];
      const float4 t = make_float4(ta.x, ta.y, tc.x, tc.y);
#elif B2A_X_TWPF > 0
      if constexpr (k1 / 2 + TWD < N1 / 2) tq[k1 / 2 + TWD] = lds128_at<16 * (k1 / 2 + TWD)>(tb_sa);
      const float4 t = tq[k1 / 2];
#else
      const float4 t = tb4[k1 / 2];
#endif
      float2 y0 = v[k1];
      if constexpr (k1 > 0) y0 = regs::cmul(y0, make_float2(t.x, t.y));
      const float2 y1 = regs::cmul(v[k1 + 1], make_float2(t.z, t.w));
      eb[slot0 * C::CP] = y0;
      eb[slot1 * C::CP] = y1;
    });
  }
}

// ---- stage 2 of one tile: warp = unit u, lane = frame.  Unit u owns the column pair (u, N1-u); unit 0 owns columns 0
// and N1/2, whose bins pair up WITHIN a column: it runs the same post-processing code after a register permutation and
// only its output bins differ (slots [0, N2/2) -> N1*(s+1), slots [N2/2, N2) -> N1/2 + N1*(s - N2/2)).
// INPLACE: the power tile overwrites this frame's exchange row (Cfg::sig); otherwise it goes to Pw in natural bin order.
// CPLX (with INPLACE): the spectrum itself, one complex value per slot (fast_stft_kernel).
// TMC: the unit's N2 post-twiddles come from tensor-memory columns `tmc` (2 * N2 of them) instead of shared memory.
template <class C, bool INPLACE, bool CPLX = false, bool TMC = false>
__device__ __forceinline__ void stage2_tile(float2* E, float* Pw, const float2* s_twp, int warp, int lane, bool pw_only,
                                            float spec_eps, uint32_t tmc = 0) {
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC;
  constexpr bool SPEC = INPLACE;
  constexpr int NQ2 = N2 / 2;  // post-twiddle quads (two twiddles each); TMC reads them in chunks of two quads
  [[maybe_unused]] float tmp[2][8];
  [[maybe_unused]] auto tm_chunk_ld = [&](auto C_) {
    constexpr int c = decltype(C_)::value;
    if constexpr (2 * c + 1 < NQ2) tm_ld<8>(tmc + 8 * c, tmp[c & 1]);
    else if constexpr (2 * c < NQ2) tm_ld<4>(tmc + 8 * c, tmp[c & 1]);
  };
  [[maybe_unused]] auto tm_chunk_wait = [&](auto C_) {
    constexpr int c = decltype(C_)::value;
    if constexpr (2 * c + 1 < NQ2) tm_wait<8>(tmp[c & 1]);
    else tm_wait<4>(tmp[c & 1]);
  };
  if constexpr (TMC) tm_chunk_ld(std::integral_constant<int, 0>{});
#ifndef B2A_NO_UREMAP
  // unit 0 (the longest: DC / Nyquist + register permutation) goes to the highest warp id — the scheduler favours
  // high warp ids, so the longest unit is not also the last one served (measured: -0.3 %)
  const int u = (C::WARPS - 1) - warp;
#else
  const int u = warp;
#endif
  const int kb_lo = u != 0 ? u : N1;                      // bin of slot s (< N2/2): kb_lo + N1*s
  const int kb_hi = u != 0 ? u : N1 / 2 - (N2 / 2) * N1;  // bin of slot s (>= N2/2): kb_hi + N1*s
  float2 A[N2], B[N2];
  const float2* ea = E + lane * C::EP + u * C::CP;
  const float2* eb = E + lane * C::EP + (N1 / 2 + u) * C::CP;
  static_for<0, N2>([&](auto I_) {
    constexpr int j = decltype(I_)::value;
    A[j] = ea[j];
    B[j] = eb[j];
  });
  Dft<N2>::run(A);
  Dft<N2>::run(B);
  // power tile: separate, natural bin order (run-time-table kernels) — or in place over this frame's exchange row,
  // one value per float2 slot, in the half selected by lane >> 4 (the row pitch is 2 * odd floats, so lanes l and
  // l + 16 would otherwise share a bank)
  float* pr = SPEC ? reinterpret_cast<float*>(E + lane * C::EP) + (lane >> 4) : Pw + lane * C::PP;
  // magnitude: ONE MUFU instruction (sqrt.approx, <= 2 ulp) — sqrtf() is the IEEE sequence (MUFU.RSQ, two Newton FFMAs, a
  // denormal slow path behind BSSY / BRA / BSYNC): ~8 instructions and a branch per bin, 513 bins per frame in the 1024 / 256
  // instance = 19 % of its instructions (ncu source page, profiles/r02_k1_1024_ncu_full.json).  The reference's abs() is a float32
  // sqrt of float32 squares; the features that follow are held to 1e-4 after a logarithm.
  auto emit = [&](float* q, float v) { *q = pw_only ? v : sqrt_approx(v + spec_eps); };
  float dc_k = 0.0f, dc_m = 0.0f;
  float2 dc_kc = make_float2(0.f, 0.f), dc_mc = make_float2(0.f, 0.f);
  if (CPLX && u == 0) post_pair_c(A[0], A[0], make_float2(1.0f, 0.0f), dc_kc, dc_mc);
  if (u == 0) {
    // unit 0: the DC / Nyquist pair comes from A[0] alone; then permute so that the shared post-processing
    // below pairs column 0 with itself (slots 0..N2/2-1: A[s+1] with A[N2-1-s]) and column N1/2 with itself
    // (slots N2/2..N2-1: B[s-N2/2] with B[3N2/2-1-s]).  One code path for every warp keeps the loop body
    // inside the 32 KB instruction cache.
    if (!CPLX) post_pair(A[0], A[0], make_float2(1.0f, 0.0f), dc_k, dc_m);
    float2 T[N2 / 2];
    static_for<0, N2 / 2>([&](auto I_) {  // newB[j] = B[N2/2+j] (j < N2/2); stash B's lower half
      constexpr int j = decltype(I_)::value;
      T[j] = B[j];
      B[j] = B[N2 / 2 + j];
    });
    static_for<0, N2 / 2>([&](auto I_) {  // newB[j] = A[j] (j >= N2/2)
      constexpr int j = decltype(I_)::value;
      B[N2 / 2 + j] = A[N2 / 2 + j];
    });
    static_for<0, N2 / 2>([&](auto I_) {  // newA[s] = A[s+1] (s < N2/2); A[N2/2] is still intact in newB
      constexpr int s = decltype(I_)::value;
      A[s] = (s + 1 < N2 / 2) ? A[s + 1] : B[N2 / 2];
    });
    static_for<0, N2 / 2>([&](auto I_) {  // newA[s] = old B[s-N2/2] (s >= N2/2)
      constexpr int j = decltype(I_)::value;
      A[N2 / 2 + j] = T[j];
    });
  }
  // post-twiddles: broadcast loads issued TWD2 pairs ahead of their use (pinned), like the inter-stage twiddles of stage 1
  constexpr int TWD2 = TMC ? 0 : (B2A_X_TWPF2 < N2 / 2 ? B2A_X_TWPF2 : N2 / 2);
  const unsigned tw_sa = (unsigned)__cvta_generic_to_shared(s_twp + u * 2 * N2);
  float4 tw4[N2 / 2];
  static_for<0, TWD2>([&](auto I_) { tw4[decltype(I_)::value] = lds128_at<16 * decltype(I_)::value>(tw_sa); });
  // quad q of the post-twiddles: from the pipelined shared-memory loads, or from the tensor-memory chunk q / 2
  auto twq = [&](auto Q_) -> float4 {
    constexpr int q = decltype(Q_)::value;
    if constexpr (TMC) {
      if constexpr (q % 2 == 0) {
        tm_chunk_wait(std::integral_constant<int, q / 2>{});
        tm_chunk_ld(std::integral_constant<int, q / 2 + 1>{});
      }
      const float* t = tmp[(q / 2) & 1] + (q % 2) * 4;
      return make_float4(t[0], t[1], t[2], t[3]);
    } else {
      if constexpr (q + TWD2 < N2 / 2) tw4[q + TWD2] = lds128_at<16 * (q + TWD2)>(tw_sa);
      return tw4[q];
    }
  };
  // slot s holds the bin pair (k, Nc - k): natural layout -> k = kb + N1*s; in-place layout -> Cfg::sig
  float* const plo = SPEC ? pr + 2 * (u * C::CP) : pr + kb_lo;
  float* const mlo = SPEC ? pr + 2 * ((N1 / 2 + u) * C::CP) : pr + (NC - kb_lo);
  float* const phi = SPEC ? plo : pr + kb_hi;
  float* const mhi = SPEC ? mlo : pr + (NC - kb_hi);
  constexpr int SK = SPEC ? 2 : N1, SM = SPEC ? 2 : -N1;
  if constexpr (CPLX) {
    float2* const ck = E + lane * C::EP + u * C::CP;             // slot of bin pair s: (u*CP + s, (N1/2 + u)*CP + s)
    float2* const cm = E + lane * C::EP + (N1 / 2 + u) * C::CP;
    static_for<0, N2 / 2>([&](auto I_) {
      constexpr int k2 = 2 * decltype(I_)::value;
      const float4 t = twq(std::integral_constant<int, k2 / 2>{});
      float2 xk, xm;
      post_pair_c(A[k2], B[N2 - 1 - k2], make_float2(t.x, t.y), xk, xm);
      ck[k2] = xk;
      cm[k2] = xm;
      post_pair_c(A[k2 + 1], B[N2 - 2 - k2], make_float2(t.z, t.w), xk, xm);
      ck[k2 + 1] = xk;
      cm[k2 + 1] = xm;
    });
    if (u == 0) {
      E[lane * C::EP + C::sig(0)] = dc_kc;
      E[lane * C::EP + C::sig(NC)] = dc_mc;
    }
    return;
  }
  static_for<0, N2 / 2>([&](auto I_) {
    constexpr int k2 = 2 * decltype(I_)::value;
    const float4 t = twq(std::integral_constant<int, k2 / 2>{});
    float pk, pm;
    post_pair(A[k2], B[N2 - 1 - k2], make_float2(t.x, t.y), pk, pm);
    emit((k2 < N2 / 2 ? plo : phi) + SK * k2, pk);
    emit((k2 < N2 / 2 ? mlo : mhi) + SM * k2, pm);
    post_pair(A[k2 + 1], B[N2 - 2 - k2], make_float2(t.z, t.w), pk, pm);
    emit((k2 + 1 < N2 / 2 ? plo : phi) + SK * (k2 + 1), pk);
    emit((k2 + 1 < N2 / 2 ? mlo : mhi) + SM * (k2 + 1), pm);
  });
  if (u == 0) {  // after the loop: in the in-place layout DC reuses the duplicate slot of the self-paired bin Nc/2
    emit(pr + (SPEC ? 2 * C::sig(0) : 0), dc_k);
    emit(pr + (SPEC ? 2 * C::sig(NC) : NC), dc_m);
  }
}

// ---- run-time-table mel phase (any filterbank, both layouts): LANE = (4 frames) x (8 consecutive mel rows) ------
// The filterbank is banded (<= 2 non-zeros per bin): a mel row is a short run of taps.  A warp-instruction
// covers 8 consecutive rows for 4 frames, so (a) the P gathers touch ~32 distinct banks (row pitch == 9 mod 32,
// neighbouring rows start a few bins apart), (b) rows are zero-padded only to the longest of 8 neighbours,
// (c) each store instruction writes four fully used 32-byte sectors of the (T, M) output.  A work item is
// (octet of rows, half of the tile's frames): 4 independent accumulators per lane share one weight load.
template <class C, bool LAYOUT_TM, bool WANT_SUMS>
__device__ __forceinline__ void mel_runtime_tables(const FastParams& p, const float* Pw, float* Y, const float* s_wg,
                                                   const int* s_start, const int* s_ginfo, double* s_sums, float* o,
                                                   int64_t lt0, int nf, int warp, int lane, float& lmax, float& lmin) {
  constexpr int NQ = 4;                       // frame quads per item
  constexpr int ROWSTEP = 4 * C::PP;          // P rows of consecutive quads
  const int M = p.n_mels, G = p.mel_groups;
  const float guard_add = p.guard_add, guard_floor = p.guard_floor, y_mul = p.y_mul, y_add = p.y_add;
  const bool use_log = p.use_log != 0;
  const int ms = lane & 7, fs = lane >> 3;
  const int2* ginfo2 = reinterpret_cast<const int2*>(s_ginfo);
  // output addressing: 32-bit element offsets from a per-tile base; (T, M): (f0 + 4q)*M + m, (M, T) staging:
  // m*33 + f0 + 4q
  float* const obase = LAYOUT_TM ? (o + lt0 * M) : Y;
  const int qstep = LAYOUT_TM ? 4 * M : 4;
  const int fstep = LAYOUT_TM ? M : 1, mstep = LAYOUT_TM ? 1 : 33;
  const bool full = nf == C::FT && (M & 7) == 0;  // every slot valid: no per-output predicates
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
        if (ok) {
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
}

// SPECK: spectrum kind fixed at compile time (B2A_SPEC_POWER / MAGNITUDE / SQRT_POWER_EPS), or -1 = run-time
// PREK: pre-emphasis support compiled in (0 = no, -1 / 1 = run-time switch); the generated-mel instances fix it per spec
// four finished values -> one store in the output element type (ODT: B2A_DTYPE_*); round-to-nearest-even like astype()
template <int ODT>
__device__ __forceinline__ void store_quad(void* base, int64_t quad_index, const float4& v) {
  if constexpr (ODT == B2A_DTYPE_F32) {
    reinterpret_cast<float4*>(base)[quad_index] = v;
  } else if constexpr (ODT == B2A_DTYPE_F16) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&a);
    u.y = *reinterpret_cast<const unsigned*>(&b);
    reinterpret_cast<uint2*>(base)[quad_index] = u;
  } else {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&a);
    u.y = *reinterpret_cast<const unsigned*>(&b);
    reinterpret_cast<uint2*>(base)[quad_index] = u;
  }
}

template <class C, bool LAYOUT_TM, bool WANT_SUMS, class MS, int SPECK, int PREK, int ODT = B2A_DTYPE_F32>
__global__ void __launch_bounds__(C::THREADS, C::MIN_BLOCKS) fast_logmel_kernel(const FastParams p) {
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC;
  constexpr bool SPEC = MS::M > 0;  // mel structure baked into code (mel_gen.cuh)
  static_assert(!SPEC || (MS::NW == C::WARPS && MS::F == C::F && MS::M % 4 == 0), "mel spec / kernel variant mismatch");
  // (M, T) output of a generated-mel kernel: float32, no per-feature sums (lane == frame in its write-out phase)
  static_assert(!SPEC || LAYOUT_TM || (!WANT_SUMS && ODT == B2A_DTYPE_F32), "generated-mel (M, T) variant: float32, no sums");
  // SPEC path staging tile Y[frame][YP]: YP/4 odd -> the STS.128 of phase A and the LDS.128 of phase B are
  // both bank-conflict free
  constexpr int YP = SPEC ? (((MS::M / 4) & 1) ? MS::M : MS::M + 4) : 4;
  constexpr int QL = SPEC ? MS::M / 4 : 1;  // lanes that carry a row quad in phase B
  using S = Smem<C, SPEC ? YP : 0>;
  constexpr bool PREFETCH = S::PREFETCH;
  extern __shared__ float4 smem4[];
  float2* const s_win2 = reinterpret_cast<float2*>(smem4 + S::WIN);  // [N2][N1]
  float2* const s_tw1 = reinterpret_cast<float2*>(smem4 + S::TW1);   // [N2][N1]
  float2* const s_twp = reinterpret_cast<float2*>(smem4 + S::TWP);   // [N1/2][2*N2]
  float2* const E = reinterpret_cast<float2*>(smem4 + S::EX);        // [FT][EP]
  float* const Pw = reinterpret_cast<float*>(smem4 + S::PW);         // [FT][PP]
  float* const xs = reinterpret_cast<float*>(smem4 + S::XS) + C::XS_HEAD;  // [ROWS][P] (+ XS_HEAD floats in front)
  const int M = SPEC ? MS::M : p.n_mels;
  const int G = p.mel_groups;
  double* const s_sums = reinterpret_cast<double*>(smem4 + S::DYN);   // [2*G*8]
  float* const s_wg = reinterpret_cast<float*>(s_sums + 2 * G * 8);   // [mel_wg_count]
  int* const s_start = reinterpret_cast<int*>(s_wg + p.mel_wg_count); // [G*8]
  int* const s_ginfo = s_start + G * 8 + ((G * 8) & 1);               // [2*G], 8-byte aligned
  __shared__ float red_max[C::WARPS], red_min[C::WARPS];
  __shared__ int s_cur_clip;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // staging tile: own space in the generated-mel kernels (the exchange buffer then holds the power tile), else it
  // aliases the exchange buffer
  float* const Y = SPEC ? Pw : reinterpret_cast<float*>(E);
  FillCtx<C> fc;
  fc.template init<PREK>(p, (unsigned)__cvta_generic_to_shared(xs));

  for (int i = threadIdx.x; i < NC; i += C::THREADS) {
    s_win2[i] = p.win2[i];
    s_tw1[i] = p.tw1[i];
    s_twp[i] = p.twp[i];
  }
  constexpr bool TMC = B2A_X_TMC && C::TM_OK;
  __shared__ uint32_t s_tm_slot;
  uint32_t tm_base = 0, tmc = 0;
  if constexpr (TMC) {
    tm_base = tm_alloc<C::TM_COLS>(&s_tm_slot, warp);
    tmc = tm_warp_base(tm_base, warp, C::TM_PER_WARP);
    for (int rr = 0; rr < C::RPW; ++rr) {
      const int n2 = warp * C::RPW + rr;
      tm_store_table(tmc + rr * 4 * N1, p.win2 + n2 * N1, N1);
      tm_store_table(tmc + rr * 4 * N1 + 2 * N1, p.tw1 + n2 * N1, N1);
    }
  }
  if (!SPEC) {
    for (int i = threadIdx.x; i < G * 8; i += C::THREADS) s_start[i] = p.mel_start[i];
    for (int i = threadIdx.x; i < 2 * G; i += C::THREADS) s_ginfo[i] = p.mel_ginfo[i];
    for (int i = threadIdx.x; i < p.mel_wg_count; i += C::THREADS) s_wg[i] = p.mel_wg[i];
  }
  constexpr bool want_sums = WANT_SUMS;
  const bool want_max = p.clip_max != nullptr;
  if (want_sums)
    for (int i = threadIdx.x; i < 2 * G * 8; i += C::THREADS) s_sums[i] = 0.0;
  if (threadIdx.x == 0) s_cur_clip = -1;

  // tile walk: tile = clip_i * tiles_per_clip + tile_i advances by gridDim.x without a division per tile
  // With per-feature sums a CTA walks a CONTIGUOUS run of tiles instead (same static balance): its float64 sums are flushed
  // when it changes clip, and the strided walk changes clip at every tile once a batch has more clips than CTAs
  // (1024 x 30 s of Parakeet-80 features: 3.16 -> 2.37 ms; LFM2-128: 4.79 -> 3.11 ms).
  const int tpc = p.tiles_per_clip;
  int step_c = (int)(gridDim.x / (unsigned)tpc), step_t = (int)(gridDim.x - (unsigned)step_c * (unsigned)tpc);
  int clip_i = (int)(blockIdx.x / (unsigned)tpc), tile_i = (int)(blockIdx.x - (unsigned)clip_i * (unsigned)tpc);
  int64_t run_left = 0;  // contiguous walk: tiles this CTA still owns (including the current one)
  if constexpr (WANT_SUMS || B2A_X_CONTIG) {
    const int64_t total = (int64_t)p.batch * tpc, q = total / gridDim.x, r = total % gridDim.x;
    const int64_t g0 = (int64_t)blockIdx.x * q + ((int64_t)blockIdx.x < r ? (int64_t)blockIdx.x : r);
    run_left = q + ((int64_t)blockIdx.x < r ? 1 : 0);
    clip_i = run_left > 0 ? (int)(g0 / tpc) : p.batch;
    tile_i = run_left > 0 ? (int)(g0 - (int64_t)clip_i * tpc) : 0;
    step_c = 0;
    step_t = 1;
  }
  if (PREFETCH && clip_i < p.batch) fill_tile<C, PREK>(p, xs, fc, clip_i, tile_i);

  const float guard_add = p.guard_add, guard_floor = p.guard_floor, y_mul = p.y_mul, y_add = p.y_add;
  const bool use_log = p.use_log != 0;
  const bool pw_only = SPECK >= 0 ? SPECK == B2A_SPEC_POWER : p.spec_kind == B2A_SPEC_POWER;
  const float spec_eps = p.spec_eps;

  // SPEC path: statistics live in registers across tiles (lane == row quad in the write-out phase)
  constexpr int NS = (SPEC && WANT_SUMS) ? 4 : 1;
  double d1[NS], d2[NS];
  int sum_clip = -1;
#pragma unroll
  for (int j = 0; j < NS; ++j) d1[j] = d2[j] = 0.0;
  int red_clip = -1, red_tile = 0;  // tile whose per-warp max / min wait in red_max / red_min (thread 0 folds them)
  auto fold_red = [&]() {
    if (SPEC && want_max && threadIdx.x == 0 && red_clip >= 0) {
      float a = red_max[0], b = red_min[0];
#pragma unroll
      for (int w = 1; w < C::WARPS; ++w) {
        a = fmaxf(a, red_max[w]);
        b = fminf(b, red_min[w]);
      }
      atomic_max_f(p.clip_max + red_clip, a);
      p.tile_min[(int64_t)red_clip * p.tile_min_pitch + red_tile] = b;
    }
  };
  auto flush_sums = [&]() {  // SPEC path; called by every thread of the CTA (the clip change is CTA-uniform)
    if (SPEC && WANT_SUMS && sum_clip >= 0) {
      if (lane < QL) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          atomicAdd(&s_sums[2 * (4 * lane + j)], d1[j]);
          atomicAdd(&s_sums[2 * (4 * lane + j) + 1], d2[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) d1[j] = d2[j] = 0.0;
      __syncthreads();
      for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) {
        atomicAdd(p.feat_sums + (int64_t)sum_clip * 2 * M + i, s_sums[i]);
        s_sums[i] = 0.0;
      }
      __syncthreads();
    }
  };

  // ---- mel, phase B of the generated-mel kernels: LANE = ROW QUAD (code shared by every warp).  Per frame: one
  // LDS.128 from the staging tile, guard + MUFU log2 + folded base-change / affine FFMA on four values, one coalesced
  // STG.128 into the (T, M) output, FMNMX3 for the tile max / min; per-feature sums stay in registers across tiles.
  // It runs one barrier interval LATE — together with stage 1 of the NEXT tile (Y has its own space): one barrier
  // fewer per tile, and its load -> store latency chains overlap the butterflies.
  auto phase_b = [&](int pclip, int ptile, int pnf) {
    if (SPEC && want_sums && sum_clip != pclip) {
      flush_sums();
      sum_clip = pclip;
    }
    float lmax = -INFINITY, lmin = INFINITY;
    // element offset of the tile's first row; quads are indexed from there (the element size is ODT's)
    const int64_t obase = (int64_t)pclip * p.out_clip_stride + (int64_t)ptile * C::FT * MS::M;
    char* const orow = reinterpret_cast<char*>(p.out) + obase * (ODT == B2A_DTYPE_F32 ? 4 : 2);
    const float4* const yb = reinterpret_cast<const float4*>(Y) + lane;
    if constexpr (SPEC && !LAYOUT_TM) {  // lane == frame, warp == rows m, m + WARPS, ...: one 128-byte column per row
      if (lane < pnf) {
        float* const ob = p.out + (int64_t)pclip * p.out_clip_stride + (int64_t)ptile * C::FT + lane;
#pragma unroll 2
        for (int m = warp; m < MS::M; m += C::WARPS) {
          const float a = fmaxf(Y[m * 32 + lane] + guard_add, guard_floor);
          const float y = use_log ? lg2_approx(a) : a;
          const float v = fmaf(y, y_mul, y_add);
          ob[(int64_t)m * p.frame_count] = v;
          lmax = fmaxf(lmax, v);
          lmin = fminf(lmin, v);
        }
      }
    } else if (lane < QL) {
      // per-feature sums: float32 partials over this warp's (at most four) frames of the tile, folded into the float64
      // accumulators once per tile — the float64 pipe is the slow one here (per value it doubled the kernel's time)
      float s1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, s2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
      for (int f = warp; f < pnf; f += C::WARPS) {
        float4 v = yb[f * (YP / 4)];
        float* e = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float a = fmaxf(e[c] + guard_add, guard_floor);
          const float y = use_log ? lg2_approx(a) : a;
          e[c] = fmaf(y, y_mul, y_add);
          if (WANT_SUMS) {
            s1[c] += e[c];
            s2[c] = fmaf(e[c], e[c], s2[c]);
          }
        }
        store_quad<ODT>(orow, f * (MS::M / 4) + lane, v);
        lmax = fmax3(lmax, v.x, v.y);
        lmin = fmin3(lmin, v.x, v.y);
        lmax = fmax3(lmax, v.z, v.w);
        lmin = fmin3(lmin, v.z, v.w);
      }
      if (WANT_SUMS) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          d1[c % NS] += (double)s1[c];
          d2[c % NS] += (double)s2[c];
        }
      }
    }
    if (want_max) {  // one REDUX each on order-preserving integer keys instead of ten dependent shuffles
      const int kmax = __reduce_max_sync(0xffffffffu, float_key(lmax));
      const int kmin = __reduce_min_sync(0xffffffffu, float_key(lmin));
      if (lane == 0) {
        red_max[warp] = key_float(kmax);
        red_min[warp] = key_float(kmin);
      }
      red_clip = pclip;   // folded by thread 0 after the next barrier
      red_tile = ptile;
    }
  };
  int prev_clip = -1, prev_tile = 0, prev_nf = 0;  // tile whose staged rows wait in Y

#ifdef B2A_PHASE_CLOCKS  // development builds: per-phase cycle counters (thread 0's view), dumped by B2A_CLOCKS=file
  long long clk_prev = 0;
  long long clk_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool clk_on = p.dbg_clk != nullptr && threadIdx.x == 0;
  auto tick = [&](int slot) {
    if (clk_on) {
      const long long t = clock64();
      clk_acc[slot] += t - clk_prev;
      clk_prev = t;
    }
  };
  if (clk_on) clk_prev = clock64();
#else
  auto tick = [](int) {};
#endif
#pragma unroll 1
  while (clip_i < p.batch) {
    const int64_t lt0 = (int64_t)tile_i * C::FT;
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);
    int nclip = clip_i + step_c, ntile = tile_i + step_t;  // the tile this CTA processes next
    if (ntile >= tpc) {
      ntile -= tpc;
      ++nclip;
    }
    if constexpr (WANT_SUMS || B2A_X_CONTIG) {
      if (--run_left <= 0) nclip = p.batch;  // end of this CTA's run
    }

    if (!PREFETCH) {
      __syncthreads();  // previous tile's mel phase has finished reading P (which shares xs' memory)
      fill_tile<C, PREK>(p, xs, fc, clip_i, tile_i);
    }
    cp_async_wait_all();
    __syncthreads();  // xs ready; run-time-table kernels: previous tile's Y written out; generated: Y(prev) complete
    tick(0);
    if (SPEC && prev_clip >= 0) phase_b(prev_clip, prev_tile, prev_nf);

    // per-CTA running per-mel sums: flush when the clip changes
    if (!SPEC && want_sums && s_cur_clip != clip_i) {
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
    stage1_tile<C, PREK, TMC>(xs, E, s_win2, s_tw1, warp, lane, p.preemph, tmc);
    __syncthreads();  // E complete, xs free
    tick(1);
    fold_red();  // the per-warp max / min of the rows phase B has just written

    // prefetch the next tile's samples while stage 2 / mel run
    if (PREFETCH && nclip < p.batch) fill_tile<C, PREK>(p, xs, fc, nclip, ntile);
    if (PREFETCH && PREK != 0) {  // pre-emphasis configs: the tile after the next one goes to L2 now
      int c2 = nclip + step_c, t2 = ntile + step_t;
      if (t2 >= tpc) {
        t2 -= tpc;
        ++c2;
      }
      if ((WANT_SUMS || B2A_X_CONTIG) && run_left <= 1) c2 = p.batch;  // contiguous walk: nothing of this CTA's beyond the next tile
      prefetch_span_l2<C>(p, c2, t2);
    }

    // ---- stage 2 ----------------------------------------------------------------------------------------
    stage2_tile<C, SPEC>(E, Pw, s_twp, warp, lane, pw_only, spec_eps);
    __syncthreads();  // Pw complete, E free (Y aliases E)
    tick(2);

    float* const o = p.out + (int64_t)clip_i * p.out_clip_stride;
    if constexpr (SPEC) {
      // ---- mel, phase A: LANE = FRAME.  The named filterbank is straight-line code (mel_gen.cuh): one
      // conflict-free LDS per bin (odd row pitch) shared by the two rows it feeds, one FFMA per tap with the weight
      // as an immediate; four finished rows are parked in Y[frame][m..m+3] with one STS.128.
      {
        const float* pr = reinterpret_cast<const float*>(E + lane * C::EP) + (lane >> 4);  // in-place power tile
        // one STS.128 per finished row quad, as inline PTX with the row offset as an immediate: written as a C++ store,
        // the last quad of every warp's branch is tail-merged into a register-addressed store of unknown alignment and
        // split into four conflicting scalar STS (+120 shared-memory wavefronts per tile)
        // (M, T) output: the staging tile is transposed, Y[m][frame] at pitch 32 — four scalar stores per quad (the same
        // four wavefronts as the STS.128), and the write-out phase reads and stores whole 128-byte columns
        const unsigned yl_sa = (unsigned)__cvta_generic_to_shared(LAYOUT_TM ? Y + lane * YP : Y + lane);
        MS::template run<C>(warp, pr, [&](auto M_, float a0, float a1, float a2, float a3) {
          constexpr int m = decltype(M_)::value;
          if constexpr (LAYOUT_TM) {
            asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(yl_sa), "n"(4 * (m / 4) * 4), "f"(a0), "f"(a1), "f"(a2),
                         "f"(a3));
          } else {
            asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(yl_sa), "n"(128 * m), "f"(a0));
            asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(yl_sa), "n"(128 * (m + 1)), "f"(a1));
            asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(yl_sa), "n"(128 * (m + 2)), "f"(a2));
            asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(yl_sa), "n"(128 * (m + 3)), "f"(a3));
          }
        });
      }
      prev_clip = clip_i;  // phase B of this tile runs after the next barrier, next to stage 1 of the next tile
      prev_tile = tile_i;
      prev_nf = nf;
      tick(3);
    } else {
      float lmax = -INFINITY, lmin = INFINITY;
      mel_runtime_tables<C, LAYOUT_TM, WANT_SUMS>(p, Pw, Y, s_wg, s_start, s_ginfo, s_sums, o, lt0, nf, warp, lane, lmax, lmin);
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
        p.tile_min[(int64_t)clip_i * p.tile_min_pitch + tile_i] = b;
      }
      if (!LAYOUT_TM && lane < nf)
        for (int m = warp; m < M; m += C::WARPS) o[(int64_t)m * p.frame_count + lt0 + lane] = Y[m * 33 + lane];
    }
    clip_i = nclip;
    tile_i = ntile;
    tick(4);
#ifdef B2A_PHASE_CLOCKS
    if (clk_on) ++clk_acc[7];
#endif
  }
#ifdef B2A_PHASE_CLOCKS
  if (clk_on)
    for (int i = 0; i < 8; ++i) p.dbg_clk[blockIdx.x * 8 + i] = clk_acc[i];
#endif
  cp_async_wait_all();
  if (SPEC) {
    __syncthreads();
    if (prev_clip >= 0) phase_b(prev_clip, prev_tile, prev_nf);
    __syncthreads();
    fold_red();
    if (want_sums) flush_sums();
  } else if (want_sums) {
    __syncthreads();
    const int prev = s_cur_clip;
    if (prev >= 0)
      for (int i = threadIdx.x; i < 2 * M; i += C::THREADS) atomicAdd(p.feat_sums + (int64_t)prev * 2 * M + i, s_sums[i]);
  }
  if constexpr (TMC) tm_free<C::TM_COLS>(tm_base, warp);
}

#include "fast_ws.cuh"  // warp-specialised two-tile pipeline kernels (fast_logmel_ws_kernel, fast_logmel_ws_tma_kernel)

// ---- fast_stft_kernel: the complex spectrum itself (dsp.stft, dsp.py:92-141) through the same fill / stage 1 / stage 2.
// Stage 2 leaves X[frame][sig(k)] in place in the exchange buffer; the copy-out makes every row of the (T, F) complex64
// output one contiguous run: warp = frame, lanes sweep the bins (coalesced 256-byte stores).
template <class C>
struct SmemStft {
  static constexpr int cdiv4(int bytes) { return (bytes + 15) / 16; }
  static constexpr int WIN = 0;
  static constexpr int TW1 = WIN + cdiv4(8 * C::NC);
  static constexpr int TWP = TW1 + cdiv4(8 * C::NC);
  static constexpr int EX = TWP + cdiv4(8 * C::NC);
  static constexpr int XS = EX + cdiv4(8 * C::FT * C::EP);
  static constexpr int SIG = XS + cdiv4(4 * C::XS_FLOATS);
  static constexpr int END = SIG + cdiv4(4 * C::F);
};

template <class C, int PREK>
__global__ void __launch_bounds__(C::THREADS, C::MIN_BLOCKS) fast_stft_kernel(const FastParams p) {
  constexpr int NC = C::NC, F = C::F;
  using S = SmemStft<C>;
  extern __shared__ float4 smem4[];
  float2* const s_win2 = reinterpret_cast<float2*>(smem4 + S::WIN);
  float2* const s_tw1 = reinterpret_cast<float2*>(smem4 + S::TW1);
  float2* const s_twp = reinterpret_cast<float2*>(smem4 + S::TWP);
  float2* const E = reinterpret_cast<float2*>(smem4 + S::EX);
  float* const xs = reinterpret_cast<float*>(smem4 + S::XS) + C::XS_HEAD;
  int* const s_sig = reinterpret_cast<int*>(smem4 + S::SIG);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FillCtx<C> fc;
  fc.template init<PREK>(p, (unsigned)__cvta_generic_to_shared(xs));
  for (int i = threadIdx.x; i < NC; i += C::THREADS) {
    s_win2[i] = p.win2[i];
    s_tw1[i] = p.tw1[i];
    s_twp[i] = p.twp[i];
  }
  for (int k = threadIdx.x; k < F; k += C::THREADS) s_sig[k] = C::sig(k);

  const int tpc = p.tiles_per_clip;
  const int step_c = (int)(gridDim.x / (unsigned)tpc), step_t = (int)(gridDim.x - (unsigned)step_c * (unsigned)tpc);
  int clip_i = (int)(blockIdx.x / (unsigned)tpc), tile_i = (int)(blockIdx.x - (unsigned)clip_i * (unsigned)tpc);
  if (clip_i < p.batch) fill_tile<C, PREK>(p, xs, fc, clip_i, tile_i);
  constexpr int J = (F + 31) / 32;  // bins per lane
#pragma unroll 1
  while (clip_i < p.batch) {
    const int64_t lt0 = (int64_t)tile_i * C::FT;
    const int64_t frames_left = p.frame_count - lt0;
    const int nf = (int)(frames_left < C::FT ? frames_left : C::FT);
    int nclip = clip_i + step_c, ntile = tile_i + step_t;
    if (ntile >= tpc) {
      ntile -= tpc;
      ++nclip;
    }
    cp_async_wait_all();
    __syncthreads();  // xs ready, E free (the previous tile's rows are on their way out)
    stage1_tile<C, PREK>(xs, E, s_win2, s_tw1, warp, lane, p.preemph);
    __syncthreads();  // E complete, xs free
    if (nclip < p.batch) fill_tile<C, PREK>(p, xs, fc, nclip, ntile);
    if (PREK != 0) {
      int c2 = nclip + step_c, t2 = ntile + step_t;
      if (t2 >= tpc) {
        t2 -= tpc;
        ++c2;
      }
      prefetch_span_l2<C>(p, c2, t2);
    }
    stage2_tile<C, true, true>(E, nullptr, s_twp, warp, lane, true, 0.0f);
    __syncthreads();  // X complete (in place)
    {
      int sl[J];
#pragma unroll
      for (int j = 0; j < J; ++j) sl[j] = s_sig[min(lane + 32 * j, F - 1)];
      float2* const o = reinterpret_cast<float2*>(p.out) + (int64_t)clip_i * p.out_clip_stride + lt0 * F + lane;
#pragma unroll 1
      for (int f = warp; f < nf; f += C::WARPS) {
        const float2* er = E + f * C::EP;
        float2* orow = o + f * F;
        float2 v[J];
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] = er[sl[j]];
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (j < J - 1 || lane + 32 * j < F) orow[32 * j] = v[j];
      }
    }
    clip_i = nclip;
    tile_i = ntile;
  }
  cp_async_wait_all();
}

template <class C, int PREK>
int launch_stft_variant(b2a_plan* plan, FastParams& p, cudaStream_t st) {
  const size_t smem = (size_t)16 * SmemStft<C>::END + 16;
  if (smem > 226 * 1024) {
    set_error("fast stft kernel: %zu bytes of shared memory needed", smem);
    return B2A_ERR_UNSUPPORTED;
  }
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, C::MIN_BLOCKS));
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  static SmemAttrOnce attr;
  if (attr.need(plan->device, smem))
    B2A_CUDA(cudaFuncSetAttribute(fast_stft_kernel<C, PREK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#if B2A_X_CTAB
  B2A_CUDA(cudaMemcpyToSymbolAsync(c_win2, p.win2, sizeof(float2) * C::NC, 0, cudaMemcpyDeviceToDevice, st));
  B2A_CUDA(cudaMemcpyToSymbolAsync(c_tw1, p.tw1, sizeof(float2) * C::NC, 0, cudaMemcpyDeviceToDevice, st));
#endif
  fast_stft_kernel<C, PREK><<<grid, C::THREADS, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

template <class C, int YP>
size_t smem_bytes(int G, int wg_count) {
  return (size_t)16 * Smem<C, YP>::DYN + sizeof(double) * 2 * G * 8 + sizeof(float) * (wg_count + (wg_count & 1)) + sizeof(int) * (G * 8 + 2 * G + 2) + 16;
}


// staging pitch of a generated-mel kernel (0 for the run-time-table kernels); must match the kernel's own YP
template <class MS>
constexpr int spec_yp() {
  return MS::M > 0 ? ((((MS::M / 4) & 1) ? MS::M : MS::M + 4)) : 0;
}

template <class C, bool TM, bool SUMS, class MS, int SPECK = -1, int PREK = -1, int ODT = B2A_DTYPE_F32>
int launch_variant(b2a_plan* plan, FastParams& p, cudaStream_t st) {
  size_t smem = smem_bytes<C, spec_yp<MS>()>(p.mel_groups, p.mel_wg_count);
  static const size_t smem_pad = getenv("B2A_SMEM_PAD") ? (size_t)atoi(getenv("B2A_SMEM_PAD")) : 0;  // profiling aid (read once): lowers the CTAs / SM
  smem += smem_pad;
  if (smem > 226 * 1024) {
    set_error("fast kernel: %zu bytes of shared memory needed", smem);
    return B2A_ERR_UNSUPPORTED;
  }
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, C::MIN_BLOCKS));
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm);
  if (grid < 1) grid = 1;
  static SmemAttrOnce attr;
  if (attr.need(plan->device, smem))
    B2A_CUDA(cudaFuncSetAttribute(fast_logmel_kernel<C, TM, SUMS, MS, SPECK, PREK, ODT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#if B2A_X_CTAB
  B2A_CUDA(cudaMemcpyToSymbolAsync(c_win2, p.win2, sizeof(float2) * C::NC, 0, cudaMemcpyDeviceToDevice, st));
  B2A_CUDA(cudaMemcpyToSymbolAsync(c_tw1, p.tw1, sizeof(float2) * C::NC, 0, cudaMemcpyDeviceToDevice, st));
#endif
  fast_logmel_kernel<C, TM, SUMS, MS, SPECK, PREK, ODT><<<grid, C::THREADS, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}

// B2A_WS (A/B runs; read once): 0 = write-out-phase kernel (fast_logmel_kernel), 1 = warp-specialised TMA kernel,
// 2 (default) = single-group TMA kernel
inline int ws_mode() {
  static const int v = [] {
    const char* e = getenv("B2A_WS");
    return e ? atoi(e) : 2;
  }();
  return v;
}

// cuTensorMapEncodeTiled through the runtime (the library links only the static CUDA runtime, not libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// B2A_TMA_OUT=0: warps write the rows themselves (phase B) instead of the TMA (A/B runs); read once
inline bool tma_out_enabled() {
  static const int v = [] {
    const char* e = getenv("B2A_TMA_OUT");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

template <class C, class MS, int SPECK, int ODT>
int launch_ws(b2a_plan* plan, FastParams& p, cudaStream_t st) {
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)plan->sm_count);
  if (grid < 1) grid = 1;
  if constexpr (ODT == B2A_DTYPE_F32 && MS::M % 32 == 0) {
    EncodeTiledFn enc = tma_out_enabled() ? encode_tiled_fn() : nullptr;
    // (instantiated for the max-type guard of the Whisper family; an additive guard takes the kernel below)
    if (enc && p.guard_add == 0.0f && p.frame_count < ((int64_t)1 << 31) && reinterpret_cast<uintptr_t>(p.out) % 16 == 0 && p.out_clip_stride % 4 == 0) {
      // output as a 4-D tensor (feature within a box, frame, box, clip): one box of the map is the whole staging tile —
      // n_mels / 32 sub-tiles of 32 frames x 32 features (128 bytes, the swizzle span), sub-tile after sub-tile in shared memory
      CUtensorMap map;
      const cuuint64_t gdim[4] = {32, (cuuint64_t)p.frame_count, (cuuint64_t)(MS::M / 32), (cuuint64_t)p.batch};
      const cuuint64_t gstride[3] = {(cuuint64_t)MS::M * 4, 128, (cuuint64_t)p.out_clip_stride * 4};
      const cuuint32_t box[4] = {32, 32, (cuuint32_t)(MS::M / 32), 1}, estr[4] = {1, 1, 1, 1};
      if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        if (p.tile_max) {  // single launch: cooperative grid (every CTA resident), grid-wide barrier, in-kernel fix-up
          if (ws_mode() == 1) return 1;
          using S1 = SmemT<C, MS::M>;
          constexpr size_t smem1 = (size_t)16 * S1::END + 16;
          auto kern = fast_logmel_tma_kernel<C, MS, SPECK, true, true>;
          static SmemAttrOnce attrf;
          if (attrf.need(plan->device, smem1)) {
            B2A_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            // a cooperative grid is validated against the occupancy calculator, which assumes the function's preferred carveout:
            // ask for the largest shared-memory partition (two 90 KB CTAs per SM; the default reported one)
            B2A_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          }
          static int resident = -1;  // CTAs of this kernel one SM holds (registers, shared memory, tensor memory)
          if (resident < 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, C::THREADS, smem1) != cudaSuccess) n = 0;
            resident = std::min(n, C::MIN_BLOCKS);
          }
          // Measured on B200 / CUDA 12.9: the calculator (and the driver's cooperative-launch check: a 296-CTA grid is refused as
          // too large) grants this kernel ONE co-resident CTA per SM although ordinary launches run two (ncu: limits 2 / 2) — at
          // one CTA per SM the kernel is 40 % slower (512 clips: 0.90 vs 0.64 ms).  The single launch is therefore taken only
          // where every tile has a CTA of its own anyway: small batches, where the two saved launches are what matters.
          if (resident < 1 || tiles > (int64_t)plan->sm_count * resident) return 1;
          const int gridf = (int)tiles;
          void* kargs[2] = {(void*)&p, (void*)&map};
          if (cudaLaunchCooperativeKernel((const void*)kern, dim3(gridf), dim3(C::THREADS), kargs, smem1, st) != cudaSuccess) {
            cudaGetLastError();  // not launchable as a cooperative grid here: the caller takes the three-launch path
            return 1;
          }
          note_launch();
          return B2A_OK;
        }
        if (ws_mode() != 1) {  // single-group kernel with the TMA write-out (the default)
          using S1 = SmemT<C, MS::M>;
          constexpr size_t smem1 = (size_t)16 * S1::END + 16;
          int per_sm = (int)((227 * 1024) / (smem1 + 1024));
          per_sm = std::max(1, std::min(per_sm, C::MIN_BLOCKS));
          const int grid1 = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)plan->sm_count * per_sm));
          static SmemAttrOnce attr1;
          if (attr1.need(plan->device, smem1))
            B2A_CUDA(cudaFuncSetAttribute(fast_logmel_tma_kernel<C, MS, SPECK, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            if (getenv("B2A_X_CARVE_K1"))  // development: shared-memory configuration as a percentage of the largest one
              B2A_CUDA(cudaFuncSetAttribute(fast_logmel_tma_kernel<C, MS, SPECK, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("B2A_X_CARVE_K1"))));
          fast_logmel_tma_kernel<C, MS, SPECK, true, false><<<grid1, C::THREADS, smem1, st>>>(p, map);
          B2A_LAUNCHED();
          return B2A_OK;
        }
        using S = SmemWST<C, MS::M>;
        constexpr size_t smem = (size_t)16 * S::END + 16;
        static_assert(smem <= 226 * 1024, "warp-specialised TMA kernel: shared memory");
        static SmemAttrOnce attr;
        if (attr.need(plan->device, smem))
          B2A_CUDA(cudaFuncSetAttribute(fast_logmel_ws_tma_kernel<C, MS, SPECK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fast_logmel_ws_tma_kernel<C, MS, SPECK, true><<<grid, 2 * C::THREADS, smem, st>>>(p, map);
        B2A_LAUNCHED();
        return B2A_OK;
      }
    }
  }
  return 1;  // no TMA instance for this call: the caller takes the write-out-phase kernel
}

// does the plan's filterbank equal the generated spec bit for bit?
template <class MS>
bool spec_matches(const b2a_plan* plan) {
  if (plan->fd.n_mels != MS::M || plan->n_freqs != MS::F) return false;
  int t = 0;
  for (int m = 0; m < MS::M; ++m) {
    const float* row = plan->h_fb.data() + (size_t)m * MS::F;
    for (int f = 0; f < MS::F; ++f) {
      const bool inside = f >= MS::kStart[m] && f < MS::kStart[m] + MS::kLen[m];
      unsigned bits;
      memcpy(&bits, row + f, 4);
      if (inside ? bits != MS::kWBits[t + f - MS::kStart[m]] : (row[f] != 0.0f)) return false;
    }
    t += MS::kLen[m];
  }
  return true;
}

// the specs each kernel variant is instantiated for: X(index, spec type, also with per-feature sums, spectrum kind
// of the wrapper that uses it, pre-emphasis compiled in — any other combination takes the run-time-table kernel)
#define B2A_SPECS_400(X)                                                                                    \
  X(1, melgen::MelSpec_whisper80, false, B2A_SPEC_POWER, 0) X(2, melgen::MelSpec_whisper128, false, B2A_SPEC_POWER, 0) \
  X(3, melgen::MelSpec_funasr80, false, B2A_SPEC_POWER, 0) X(4, melgen::MelSpec_hf_whisper80, false, B2A_SPEC_POWER, 0) \
  X(5, melgen::MelSpec_hf_whisper128, false, B2A_SPEC_POWER, 0)
#define B2A_SPECS_512(X)                                                                                     \
  X(1, melgen::MelSpec_parakeet80, true, B2A_SPEC_POWER, 1) X(2, melgen::MelSpec_parakeet128, true, B2A_SPEC_POWER, 1) \
  X(3, melgen::MelSpec_nemo_slaney80, true, B2A_SPEC_POWER, 1) X(4, melgen::MelSpec_nemo_slaney128, true, B2A_SPEC_POWER, 1)
#define B2A_SPECS_1024(X) \
  X(1, melgen::MelSpec_vocos100, false, B2A_SPEC_MAGNITUDE, 0) X(2, melgen::MelSpec_qwen3tts128, false, B2A_SPEC_SQRT_POWER_EPS, 0)

template <class C>
struct SpecList;

#define B2A_MATCH(IDX, MS, SUMS_OK, SPECK, PREK) if (spec_matches<MS>(plan)) { *name = MS::kName; return IDX; }
#define B2A_LAUNCH(IDX, MS, SUMS_OK, SPECK, PREK)                                                        \
  case IDX:                                                                                              \
    if (p.spec_kind != SPECK || (PREK == 0 && p.preemph != 0.0f)) break;                                 \
    if (p.out_layout != B2A_LAYOUT_TM) { /* (M, T) rows: S3Tokenizer, Voxtral-RT — the 400/160 family */ \
      if constexpr (C::N == 400) {                                                                       \
        if (!sums && p.out_dtype == B2A_DTYPE_F32)                                                       \
          return launch_variant<C, false, false, MS, SPECK, PREK>(plan, p, st);                         \
      }                                                                                                  \
      break;                                                                                             \
    }                                                                                                    \
    if constexpr (C::N == 400 && PREK == 0 && MS::M % 32 == 0) { /* epilogue in the mel phase + TMA write-out (fast_ws.cuh) */ \
      if (!sums && p.out_dtype == B2A_DTYPE_F32 && ws_mode() != 0) {                                     \
        const int rc = launch_ws<C, MS, SPECK, B2A_DTYPE_F32>(plan, p, st);                              \
        if (rc != 1) return rc;                                                                          \
      }                                                                                                  \
    }                                                                                                    \
    if (p.tile_max) return 1; /* a fused (single-launch) request is served by the TMA kernel or not at all */ \
    if constexpr (C::N == 400) { /* 16-bit feature output: the encoder-facing 400/160 family */          \
      if (!sums && p.out_dtype == B2A_DTYPE_F16)                                                         \
        return launch_variant<C, true, false, MS, SPECK, PREK, B2A_DTYPE_F16>(plan, p, st);             \
      if (!sums && p.out_dtype == B2A_DTYPE_BF16)                                                        \
        return launch_variant<C, true, false, MS, SPECK, PREK, B2A_DTYPE_BF16>(plan, p, st);            \
    }                                                                                                    \
    if (p.out_dtype != B2A_DTYPE_F32) return B2A_ERR_UNSUPPORTED;                                        \
    if (!sums) return launch_variant<C, true, false, MS, SPECK, PREK>(plan, p, st);                     \
    if constexpr (SUMS_OK) return launch_variant<C, true, true, MS, SPECK, PREK>(plan, p, st);          \
    break;
#define B2A_SPECLIST(CFG, LIST)                                                                             \
  template <>                                                                                               \
  struct SpecList<CFG> {                                                                                    \
    using C = CFG;                                                                                          \
    static int match(const b2a_plan* plan, const char** name) { LIST(B2A_MATCH) return 0; }                 \
    /* returns B2A_OK after a launch, or 1 when (spec, sums) has no instance and the caller falls back */   \
    static int launch(int spec, bool sums, b2a_plan* plan, FastParams& p, cudaStream_t st) {                \
      switch (spec) { LIST(B2A_LAUNCH) default: break; }                                                    \
      return 1;                                                                                             \
    }                                                                                                       \
  };

template <class C>
int launch(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) {
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  if (tiles >= (int64_t)1 << 31) {
    set_error("fast kernel: %lld tiles in one launch (split the batch)", (long long)tiles);
    return B2A_ERR_UNSUPPORTED;
  }
  if (p.spec_kind == B2A_SPEC_COMPLEX) {
    using CS = typename C::Padded;
    return p.preemph != 0.0f ? launch_stft_variant<CS, 1>(plan, p, st) : launch_stft_variant<CS, 0>(plan, p, st);
  }
  const bool tm = p.out_layout == B2A_LAYOUT_TM, sums = p.feat_sums != nullptr;
  const bool vec_ok = reinterpret_cast<uintptr_t>(p.out) % 16 == 0 && p.out_clip_stride % 4 == 0;  // STG.128 rows
  if (p.out_dtype != B2A_DTYPE_F32 && !(tm && vec_ok && fs->spec > 0 && !sums && C::N == 400)) {
    set_error("16-bit feature output needs a 400/160 generated-mel kernel, (T, M) layout, 16-byte aligned rows, no normalisation");
    return B2A_ERR_UNSUPPORTED;
  }
  static const bool no_melspec = getenv("B2A_NO_MELSPEC") != nullptr;  // development toggle, read once
  if (p.tile_max) {  // fused single-launch request: generated-mel TMA kernel only
    if (fs->spec > 0 && !no_melspec && tm && vec_ok && !sums) return SpecList<C>::launch(fs->spec, sums, plan, p, st);
    return 1;
  }
  if (fs->spec > 0 && !no_melspec && ((tm && vec_ok) || (!tm && !sums && p.out_dtype == B2A_DTYPE_F32 && C::N == 400))) {  // named filterbank: mel structure compiled into the kernel
    const int rc = SpecList<C>::launch(fs->spec, sums, plan, p, st);
    if (rc != 1) return rc;
  }
  if (tm && !sums) return launch_variant<C, true, false, NoSpec>(plan, p, st);
  if (tm && sums) return launch_variant<C, true, true, NoSpec>(plan, p, st);
  if (!tm && !sums) return launch_variant<C, false, false, NoSpec>(plan, p, st);
  return launch_variant<C, false, true, NoSpec>(plan, p, st);
}

}  // namespace
}  // namespace b2a
