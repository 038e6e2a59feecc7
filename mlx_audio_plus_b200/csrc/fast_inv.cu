// b200audio — fused iSTFT ("fast inverse") for n_fft = 4*hop vocoder heads: Vocos 1024/256, Vocos-Encodec 1280/320
// (dsp.py:144-217 istft, dsp.py:350-417 ISTFTCache.istft).
//
// Mirror image of the forward fast kernel: LANE == FRAME, WARP == COLUMN ROLE, everything between the (B, F, T)
// spectrum in HBM and the waveform in HBM lives in registers and ONE shared-memory buffer.
//
//   tile     32 consecutive frames (lane l <-> frame t0 - 4 + l); the tile owns the FA = 28 * hop output samples all
//            of whose (up to 4) contributing frames are inside it; the 3-frame halo is recomputed by the neighbour
//            (12.5 % redundant work, but no atomics, no cross-CTA ordering, loads start on a 32-byte sector).
//   step 1   warp = unit u (bin columns u and N1-u): each lane reads its frame's X[k], X[Nc-k] (the (B, F, T) layout
//            makes every load a coalesced 256-byte row across the warp), undoes the Hermitian packing
//            (Z = E + iO: the half-size complex spectrum of z[m] = x[2m] + i x[2m+1]; Im(DC), Im(Nyquist) ignored as
//            irfft does), inverse DFT-N2 in registers, inter-stage twiddle, store to E[frame][k1][n2].
//            The inverse transform runs in the "swapped domain" (re <-> im): it is then the FORWARD codelets with
//            the FORWARD twiddles.
//   step 2   warp = residue n2: DFT-N1 in registers, window (1/N folded in), and the time samples go back IN PLACE:
//            row `frame` of E becomes the windowed frame in natural sample order.
//   step 3   gather-form overlap-add: a thread owns two adjacent output samples, sums its <= 4 frames in ascending
//            frame order (the order of the reference's sequential scatter-add), divides by the window envelope
//            (period-hop table when all four frames exist, summed on the fly at the clip's edges) and stores
//            coalesced float2.  No T x n_fft intermediate, no index arrays, no atomics in HBM.
// Algorithmic bytes: F*8 in + hop*4 out per frame (5128 B for Vocos).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "fft_regs.cuh"

#ifndef B2A_X_PF
#define B2A_X_PF 1
#endif
// Spectrum loads bypass L1 (ld.global.cg): every sector is read by exactly one warp, and an L1 miss has to hold an L1 line until its
// data returns — with most of the SM's unified memory carved out as shared memory (28 KB of L1 next to two 89 KB CTAs) the few
// lines left cap the loads in flight, and the kernels are load-latency bound
#ifndef B2A_X_LDCG
#define B2A_X_LDCG 0
#endif
#if B2A_X_LDCG
#define B2A_LD(p) __ldcg(p)
#else
#define B2A_LD(p) __ldg(p)
#endif

namespace b2a {

namespace {

using regs::Dft;
using regs::static_for;

template <int N1_, int N2_>
struct ICfg {
  static constexpr int N1 = N1_, N2 = N2_;
  static constexpr int NC = N1 * N2, N = 2 * NC, F = NC + 1, HOP = N / 4;
  static constexpr int WARPS = N1 / 2, THREADS = WARPS * 32;
  static constexpr int FT = 32, HALO = 4, FA = FT - HALO;   // frames per tile, halo lanes (3 needed, 4 keeps alignment)
  static constexpr int EP = NC + 1;                         // row pitch in float2 (odd: lane-strided access conflict free)
  static constexpr int S = FA * HOP;                        // output samples owned by a tile
  static_assert(N2 == WARPS, "one step-2 residue per warp");
  static_assert(N2 % 2 == 0 && N1 % 2 == 0, "even factors");
  static_assert((S / 2) % THREADS == 0, "whole sample pairs per thread");
};

struct InvFastParams {
  const float2* spec;    // interleaved (B, F, T), or nullptr
  const float* spec_re;  // planar
  const float* spec_im;
  int64_t clip_stride, T;
  int batch, tiles_per_clip;
  int norm_sq, div_clamp, vec_ok;
  int64_t out_start, out_len, out_clip_stride;
  float* out;
  const float2* twp;   // [WARPS][N2]   pre-twiddles conj(W_N^k) per unit slot
  const float2* tw1;   // [WARPS][2][N2] inter-stage twiddles W_Nc^(n2*k1) for the unit's two columns
  const float2* win2;  // [N2][N1]      (w[2m], w[2m+1]) / N with m = N2*n1 + n2
  const float* wenv;   // [N]           w or w^2 (envelope taps)
  const float* den;    // [HOP]         1 / (full-overlap envelope summed in ascending frame order)
  int rden_ok;         // every entry of the full-overlap envelope is above the division guard
  PolarSpec polar;
  float div_eps;
};

struct InvFastState {
  float2 *d_twp = nullptr, *d_tw1 = nullptr, *d_win2 = nullptr;
  float *d_wenv = nullptr, *d_den = nullptr;
  int variant = 0, rden_ok = 1;
};

// POLAR (magnitude / phase planes) is a compile-time variant: the accurate sincosf is ~100 instructions per call
// site and would otherwise sit, unused, in the instruction stream of the complex-input kernel (32 KB I-cache).
template <int BYTE_OFF>
__device__ __forceinline__ float4 lds128_pinned(unsigned base) {  // volatile: keeps its place in program order
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base), "n"(BYTE_OFF));
  return v;
}

// PLANAR (two float32 planes: real / imaginary, or magnitude / phase with POLAR) is a compile-time variant as well: as a run-time
// test every one of the 34 loads per thread carried its own branch chain (frame alive? planes or interleaved?) — 547 BRA,
// 1 313 IADD3 and 1 066 LEA next to 516 LDG per tile on the ncu source page (round 2).
template <class C, bool POLAR, bool PLANAR>
__global__ void __launch_bounds__(C::THREADS, 1) fast_istft_kernel(const InvFastParams p) {
  static_assert(PLANAR || !POLAR, "magnitude / phase input comes as two planes");
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC, HOP = C::HOP;
  extern __shared__ float4 smem4[];
  float2* const E = reinterpret_cast<float2*>(smem4);                   // [FT][EP]
  float2* const s_twp = E + C::FT * C::EP;                              // [WARPS][N2] (FT*EP is even: 16-byte aligned)
  float2* const s_tw1 = s_twp + C::WARPS * N2;                          // [WARPS][2][N2]
  float2* const s_win2 = s_tw1 + C::WARPS * 2 * N2;                     // [N2][N1]
  float* const s_wenv = reinterpret_cast<float*>(s_win2 + N2 * N1);     // [N]
  float* const s_rden = s_wenv + C::N;                                  // [HOP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C::WARPS * N2; i += C::THREADS) s_twp[i] = p.twp[i];
  for (int i = threadIdx.x; i < C::WARPS * 2 * N2; i += C::THREADS) s_tw1[i] = p.tw1[i];
  for (int i = threadIdx.x; i < NC; i += C::THREADS) s_win2[i] = p.win2[i];
  for (int i = threadIdx.x; i < C::N; i += C::THREADS) s_wenv[i] = p.wenv[i];
  for (int i = threadIdx.x; i < HOP; i += C::THREADS) s_rden[i] = p.den[i];

  const int u = warp;
  const int kb_lo = u != 0 ? u : N1;                      // bin of slot s (< N2/2): kb_lo + N1*s
  const int kb_hi = u != 0 ? u : N1 / 2 - (N2 / 2) * N1;  // bin of slot s (>= N2/2): kb_hi + N1*s
  const int rowA = u, rowB = u != 0 ? N1 - u : N1 / 2;    // E rows (k1) of the unit's two columns
  const int tpc = p.tiles_per_clip;
  const int64_t T = p.T;

#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < (int64_t)p.batch * tpc; tile += gridDim.x) {
    const int clip_i = (int)(tile / tpc);
    const int tile_i = (int)(tile - (int64_t)clip_i * tpc);
    const int64_t t_first = (int64_t)tile_i * C::FA - C::HALO;  // frame of lane 0
    const int64_t t = t_first + lane;
    const bool live = t >= 0 && t < T;
    __syncthreads();  // previous tile's overlap-add has finished reading E

    // ---- step 1 ---------------------------------------------------------------------------------------
    {
      float2 A[N2], B[N2];  // columns (u, N1-u) in k2 order; swapped domain (x = Im, y = Re)
      // Running POINTERS: bin k of slot s is kb + N1*s, its partner Nc - k; consecutive slots are N1*T elements apart, so every
      // address is one 64-bit add away from the previous one; a dead lane (frame outside the clip) loads nothing (predicated)
      // (magnitude / phase planes are converted after ALL loads of the unit are in flight: a conversion per load site put
      // its range-reduction branch between consecutive loads and serialised them)
      const int64_t stepT = (int64_t)N1 * T, jumpT = (int64_t)(kb_hi - kb_lo) * T;
      const int64_t oa0 = (int64_t)clip_i * p.clip_stride + (live ? t : 0) + (int64_t)kb_lo * T;
      const int64_t ob0 = (int64_t)clip_i * p.clip_stride + (live ? t : 0) + (int64_t)(NC - kb_lo) * T;
      const float2* qa = p.spec + oa0;   // interleaved
      const float2* qb = p.spec + ob0;
      const float *ra = p.spec_re + oa0, *ia = p.spec_im + oa0;  // planes
      const float *rb = p.spec_re + ob0, *ib = p.spec_im + ob0;
      const float2 zero2 = make_float2(0.0f, 0.0f);
      float2 xa[N2], xb[N2];
      static_for<0, N2>([&](auto S_) {
        constexpr int s = decltype(S_)::value;
        if (s == N2 / 2) {  // unit 0 switches from column 0 to column N1/2 here (jumpT == 0 for the other units)
          if constexpr (PLANAR) {
            ra += jumpT; ia += jumpT; rb -= jumpT; ib -= jumpT;
          } else {
            qa += jumpT; qb -= jumpT;
          }
        }
        if constexpr (PLANAR) {
          xa[s] = live ? make_float2(B2A_LD(ra), B2A_LD(ia)) : zero2;
          xb[s] = live ? make_float2(B2A_LD(rb), B2A_LD(ib)) : zero2;
          ra += stepT; ia += stepT; rb -= stepT; ib -= stepT;
        } else {
          xa[s] = live ? B2A_LD(qa) : zero2;
          xb[s] = live ? B2A_LD(qb) : zero2;
          qa += stepT; qb -= stepT;
        }
      });
      auto load_at = [&](int64_t i) -> float2 {
        if (!live) return zero2;
        if constexpr (PLANAR) return make_float2(B2A_LD(p.spec_re + i), B2A_LD(p.spec_im + i));
        else return B2A_LD(p.spec + i);
      };
      float2 dc = make_float2(0.0f, 0.0f);
      if (u == 0) {  // Im(DC), Im(Nyquist) are ignored (irfft)
        const int64_t o0 = (int64_t)clip_i * p.clip_stride + t;
        float2 x0 = load_at(o0), xn = load_at(o0 + (int64_t)NC * T);
        if (POLAR && live) {
          x0 = polar_to_complex(p.polar, x0);
          xn = polar_to_complex(p.polar, xn);
        }
        dc = make_float2(x0.x, xn.x);
      }
      if (POLAR && live) {
#pragma unroll
        for (int s = 0; s < N2; ++s) {
          xa[s] = polar_to_complex(p.polar, xa[s]);
          xb[s] = polar_to_complex(p.polar, xb[s]);
        }
      }
      const float4* tp4 = reinterpret_cast<const float4*>(s_twp + u * N2);
      // Hermitian unpacking of one slot; the two results land in DIFFERENT registers for unit 0 (whose columns 0 and
      // N1/2 pair up within themselves), so the unpacking is instantiated twice behind a warp-uniform branch rather
      // than followed by ~50 predicated register moves in every warp:
      //   general: A[s] = Z'[k], B[N2-1-s] = Z'[Nc-k]
      //   unit 0 : slots s < N2/2 -> col0[s+1], col0[N2-1-s]; slots s >= N2/2 -> colH[s-N2/2], colH[3N2/2-1-s]
      auto unpack = [&](auto U0_) {
        constexpr bool U0 = decltype(U0_)::value;
        static_for<0, N2 / 2>([&](auto S_) {
          constexpr int s0 = 2 * decltype(S_)::value;
          const float4 w2 = tp4[s0 / 2];  // (c, s) of slots s0, s0+1: conj(W_N^k) = (cos, sin)(2 pi k / N)
          static_for<0, 2>([&](auto J_) {
            constexpr int s = s0 + decltype(J_)::value;
            const float2 w = decltype(J_)::value == 0 ? make_float2(w2.x, w2.y) : make_float2(w2.z, w2.w);
            const float2 a = xa[s], b = xb[s];
            const float2 e2 = regs::pfma(b, make_float2(1.0f, -1.0f), a);   // a + conj(b)
            const float2 d = regs::pfma(b, make_float2(-1.0f, 1.0f), a);    // a - conj(b)
            const float2 o2 = regs::cmul(d, w);                             // (a - conj b) * conj(W_N^k)
            // Z[k] = E + iO, Z[Nc-k] = conj(E) + i conj(O); stored swapped (im, re)
            const float2 zk = regs::pfma(o2, make_float2(1.0f, -1.0f), regs::pswap(e2));   // (e_i + o_r, e_r - o_i)
            const float2 zm = regs::pfma(regs::pswap(e2), make_float2(-1.0f, 1.0f), o2);   // (o_r - e_i, o_i + e_r)
            if constexpr (!U0) {
              A[s] = zk;
              B[N2 - 1 - s] = zm;
            } else if constexpr (s < N2 / 2) {
              if constexpr (s + 1 < N2 / 2) A[s + 1] = zk;   // the self-paired bin Nc/2 (s + 1 == N2/2) arrives as zm
              A[N2 - 1 - s] = zm;
            } else {
              B[s - N2 / 2] = zk;
              B[3 * N2 / 2 - 1 - s] = zm;
            }
          });
        });
      };
      if (u == 0) {
        unpack(std::true_type{});
        A[0] = make_float2(dc.x - dc.y, dc.x + dc.y);  // Z[0] = (X0 + XN) + i (X0 - XN), swapped
      } else {
        unpack(std::false_type{});
      }
      Dft<N2>::run(A);
      Dft<N2>::run(B);
      const float4* t4a = reinterpret_cast<const float4*>(s_tw1 + (u * 2 + 0) * N2);
      const float4* t4b = reinterpret_cast<const float4*>(s_tw1 + (u * 2 + 1) * N2);
      float2* ea = E + lane * C::EP + rowA * N2;
      float2* eb = E + lane * C::EP + rowB * N2;
      static_for<0, N2 / 2>([&](auto I_) {
        constexpr int n2 = 2 * decltype(I_)::value;
        const float4 ta = t4a[n2 / 2], tb = t4b[n2 / 2];
        ea[n2] = n2 == 0 ? A[0] : regs::cmul(A[n2], make_float2(ta.x, ta.y));
        ea[n2 + 1] = regs::cmul(A[n2 + 1], make_float2(ta.z, ta.w));
        eb[n2] = n2 == 0 ? B[0] : regs::cmul(B[n2], make_float2(tb.x, tb.y));
        eb[n2 + 1] = regs::cmul(B[n2 + 1], make_float2(tb.z, tb.w));
      });
    }
    __syncthreads();  // E[frame][k1][n2] complete
#if B2A_X_PF
    // The NEXT tile's spectrum rows are pulled into L2 while this tile is transformed: the kernel holds ONE tile per SM (131 KB
    // exchange buffer, 126 registers), so nothing else keeps the DRAM busy during steps 2 and 3, and the loads at the top of the
    // next tile otherwise wait out a full DRAM round trip with every warp of the SM stalled on them.
    {
      const int64_t ntile = tile + gridDim.x;
      if (ntile < (int64_t)p.batch * tpc) {
        const int nclip = (int)(ntile / tpc);
        const int nti = (int)(ntile - (int64_t)nclip * tpc);
        int64_t tf = (int64_t)nti * C::FA - C::HALO, tl = tf + C::FT - 1;
        tf = tf < 0 ? 0 : tf;
        tl = tl > T - 1 ? T - 1 : tl;
        if (tl >= tf) {
          const int64_t nb = (int64_t)nclip * p.clip_stride;
          for (int k = threadIdx.x; k <= NC; k += C::THREADS) {
            const int64_t o0 = nb + (int64_t)k * T + tf, o1 = nb + (int64_t)k * T + tl;
            if constexpr (PLANAR) {
              for (uintptr_t q = reinterpret_cast<uintptr_t>(p.spec_re + o0) & ~(uintptr_t)127; q <= reinterpret_cast<uintptr_t>(p.spec_re + o1); q += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
              for (uintptr_t q = reinterpret_cast<uintptr_t>(p.spec_im + o0) & ~(uintptr_t)127; q <= reinterpret_cast<uintptr_t>(p.spec_im + o1); q += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            } else {
              for (uintptr_t q = reinterpret_cast<uintptr_t>(p.spec + o0) & ~(uintptr_t)127; q <= reinterpret_cast<uintptr_t>(p.spec + o1); q += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
          }
        }
      }
    }
#endif

    // ---- step 2 ---------------------------------------------------------------------------------------
    {
      const int n2 = warp;
      float2 v[N1];
      float2* er = E + lane * C::EP + n2;
      static_for<0, N1>([&](auto K_) {
        constexpr int k1 = decltype(K_)::value;
        v[k1] = er[k1 * N2];
      });
      Dft<N1>::run(v);
      // window pairs: pinned broadcast loads issued WPD pairs ahead of their use (as the twiddles of the forward kernel)
      constexpr int WPD = 2;
      const unsigned wb_sa = (unsigned)__cvta_generic_to_shared(s_win2 + n2 * N1);
      float4 wq[N1 / 2];
      static_for<0, WPD>([&](auto I_) { wq[decltype(I_)::value] = lds128_pinned<16 * decltype(I_)::value>(wb_sa); });
      static_for<0, N1 / 2>([&](auto I_) {
        constexpr int n1 = 2 * decltype(I_)::value;
        if constexpr (n1 / 2 + WPD < N1 / 2) wq[n1 / 2 + WPD] = lds128_pinned<16 * (n1 / 2 + WPD)>(wb_sa);
        const float4 w = wq[n1 / 2];
        // swapped domain: v = (Im z, Re z); sample pair (x[2m], x[2m+1]) = (Re z, Im z) * (w[2m], w[2m+1]) / N
        er[n1 * N2] = regs::pmul(regs::pswap(v[n1]), make_float2(w.x, w.y));
        er[(n1 + 1) * N2] = regs::pmul(regs::pswap(v[n1 + 1]), make_float2(w.z, w.w));
      });
    }
    __syncthreads();  // rows of E are now windowed frames in natural sample order

    // ---- step 3: overlap-add -------------------------------------------------------------------------
    {
      // THREADS * 2 consecutive samples per sweep: a thread keeps its hop-relative offset r and walks 4 hops per
      // sweep, so every address below advances by a constant
      constexpr int HPS = 2 * C::THREADS / HOP;            // hops per sweep
      static_assert((2 * C::THREADS) % HOP == 0, "whole hops per sweep");
      const float* Yf = reinterpret_cast<const float*>(E);
      float* const o = p.out + (int64_t)clip_i * p.out_clip_stride;
      const int q0 = (2 * threadIdx.x) / HOP, r = 2 * threadIdx.x - q0 * HOP;
      const int t_q0 = tile_i * C::FA;  // frame index of q = 0
      const int Ti = (int)T;
      const float* y = Yf + (q0 + 1) * (2 * C::EP) + r + 3 * HOP;
      int64_t j0 = (int64_t)tile_i * C::S + 2 * threadIdx.x - p.out_start;
      const float2 rden = *reinterpret_cast<const float2*>(s_rden + r);
      // interior tile (CTA-uniform): every frame the tile's samples touch exists, the reciprocal envelope table applies, every
      // output pair is inside the clip's range and 8-byte aligned — the sweep is then loads, adds, one multiply and one store
      // (the general loop below spends more instructions on its range / edge tests than on the overlap-add: ISETP 1 076, IADD3 821,
      // BRA 380 against 474 LDS per tile, ncu source page)
      const int64_t jt0 = (int64_t)tile_i * C::S - p.out_start;
      const bool interior = t_q0 - 3 >= 0 && t_q0 + C::FA - 1 < Ti && p.rden_ok && p.vec_ok && jt0 >= 0 && jt0 + C::S <= p.out_len;
      if (interior) {
        float* oj = o + j0;
#pragma unroll
        for (int it = 0; it < (C::S / 2) / C::THREADS; ++it) {
          const float* yi = y + it * (HPS * 2 * C::EP);
          float2 num = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 yv = *reinterpret_cast<const float2*>(yi + j * (2 * C::EP - HOP));
            num.x += yv.x;
            num.y += yv.y;
          }
          *reinterpret_cast<float2*>(oj + it * (2 * C::THREADS)) = make_float2(num.x * rden.x, num.y * rden.y);
        }
      } else
#pragma unroll 1
      for (int it = 0; it < (C::S / 2) / C::THREADS; ++it, y += HPS * 2 * C::EP, j0 += 2 * C::THREADS) {
        const int q = q0 + HPS * it;
        // contributing frames: lanes q+1 .. q+4 (frames t_q0 + q - 3 .. t_q0 + q), sample offsets r + 3*HOP .. r
        float2 num = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 yv = *reinterpret_cast<const float2*>(y + j * (2 * C::EP - HOP));
          num.x += yv.x;
          num.y += yv.y;
        }
        const int ta = t_q0 + q - 3, tb = t_q0 + q;
        float2 res;
        if (ta >= 0 && tb < Ti && p.rden_ok) {
          // all four frames exist: the envelope is the period-hop table; its reciprocal (rounded once from double)
          // replaces the division (<= 1 ulp from the quotient)
          res = make_float2(num.x * rden.x, num.y * rden.y);
        } else {
          float2 den = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tt = ta + j;
            if (tt >= 0 && tt < Ti) {
              const float2 wv = *reinterpret_cast<const float2*>(s_wenv + r + (3 - j) * HOP);
              den.x += wv.x;
              den.y += wv.y;
            }
          }
          if (p.div_clamp) {
            res.x = __fdiv_rn(num.x, fmaxf(den.x, p.div_eps));
            res.y = __fdiv_rn(num.y, fmaxf(den.y, p.div_eps));
          } else {
            res.x = den.x > p.div_eps ? __fdiv_rn(num.x, den.x) : num.x;
            res.y = den.y > p.div_eps ? __fdiv_rn(num.y, den.y) : num.y;
          }
        }
        if (p.vec_ok && j0 >= 0 && j0 + 1 < p.out_len) {
          *reinterpret_cast<float2*>(o + j0) = res;
        } else {
          if (j0 >= 0 && j0 < p.out_len) o[j0] = res.x;
          if (j0 + 1 >= 0 && j0 + 1 < p.out_len) o[j0 + 1] = res.y;
        }
      }
    }
  }
}

template <class C>
size_t inv_smem_bytes() {
  return sizeof(float2) * ((size_t)C::FT * C::EP + C::WARPS * C::N2 * 3 + C::NC) + sizeof(float) * (C::N + C::HOP) + 16;
}

template <class C, bool POLAR, bool PLANAR>
int launch_inv(b2a_plan* plan, InvFastParams& p, cudaStream_t st) {
  const size_t smem = inv_smem_bytes<C>();
  static SmemAttrOnce attr;
  if (attr.need(plan->device, smem))
    B2A_CUDA(cudaFuncSetAttribute(fast_istft_kernel<C, POLAR, PLANAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_clip;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, plan->sm_count));
  fast_istft_kernel<C, POLAR, PLANAR><<<grid, C::THREADS, smem, st>>>(p);
  B2A_LAUNCHED();
  return B2A_OK;
}


// ---- fast_istft16_kernel: 16-frame tiles, two CTAs per SM, overlap carried from tile to tile -------------------------------------
// The 32-frame kernel above holds ONE tile per SM (131 KB exchange buffer, 126 registers x 512 threads): its three phases — spectrum
// loads, register transforms, overlap-add — run one after the other and nothing overlaps them (issue slots 36 %, DRAM 38 % busy,
// profiles/r02_k3_1024_ncu_full.json), and every tile re-transforms the 4 halo frames it shares with its neighbour (12.5 %).
// Here a half-warp is a group of 16 FRAMES and the two half-warps of a warp play two different column roles (units in step 1,
// residues in step 2), so a tile is 16 frames, its exchange buffer 64 KB and a CTA 8 warps: TWO CTAs per SM whose phases interleave.
// A CTA walks a RUN of consecutive tiles of one clip and carries the partial overlap-add sums of the next three hops from tile to
// tile (3 x hop floats, double buffered), so only the first tile of a run re-transforms a 3-frame lead-in.  The summation order per
// output sample is unchanged: frames in ascending order, starting from 0 (the reference's sequential scatter-add, dsp.py:193-204).
template <int N1_, int N2_>
struct ICfg16 {
  static constexpr int N1 = N1_, N2 = N2_;
  static constexpr int NC = N1 * N2, N = 2 * NC, F = NC + 1, HOP = N / 4;
  static constexpr int ROLES = N1 / 2, WARPS = ROLES / 2, THREADS = WARPS * 32;
  static constexpr int FT = 16, LEAD = 3;     // frames per tile; frames a run's first tile transforms ahead of its first hop
  static constexpr int EP = NC + 1;           // row pitch in float2 (odd)
  static_assert(N2 == ROLES, "one step-2 residue per role");
  static_assert(N2 % 2 == 0 && N1 % 4 == 0, "even factors, two roles per warp");
  static_assert(HOP % 2 == 0 && (HOP / 2) * 2 == THREADS, "a hop is THREADS / 2 sample pairs: two hops per overlap-add sweep");
};

struct InvRunParams {
  int hop_first, hop_end;   // hops [hop_first, hop_end) intersect the output range
  int run_hops;             // hops a run produces (16 * tiles - 3)
  int runs_per_clip;
};

template <class C>
size_t inv16_smem_bytes() {
  return sizeof(float2) * ((size_t)C::FT * C::EP + C::ROLES * C::N2 * 3 + C::NC) + sizeof(float) * (C::N + C::HOP + 2 * 3 * C::HOP) + 16;
}

template <class C, bool POLAR, bool PLANAR>
__global__ void __launch_bounds__(C::THREADS, 2) fast_istft16_kernel(const InvFastParams p, const InvRunParams rp) {
  static_assert(PLANAR || !POLAR, "magnitude / phase input comes as two planes");
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC, HOP = C::HOP, FT = C::FT;
  extern __shared__ float4 smem4[];
  float2* const E = reinterpret_cast<float2*>(smem4);                   // [FT][EP]
  float2* const s_twp_al = E + FT * C::EP;                              // [ROLES][N2] (FT * EP is even: 16-byte aligned)
  float2* const s_tw1 = s_twp_al + C::ROLES * N2;                       // [ROLES][2][N2]
  float2* const s_win2 = s_tw1 + C::ROLES * 2 * N2;                     // [N2][N1]
  float* const s_wenv = reinterpret_cast<float*>(s_win2 + N2 * N1);     // [N]
  float* const s_rden = s_wenv + C::N;                                  // [HOP]
  float* const s_carry = s_rden + HOP;                                  // [2][3][HOP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fr = lane & 15, role = 2 * warp + (lane >> 4);
  for (int i = threadIdx.x; i < C::ROLES * N2; i += C::THREADS) s_twp_al[i] = p.twp[i];
  for (int i = threadIdx.x; i < C::ROLES * 2 * N2; i += C::THREADS) s_tw1[i] = p.tw1[i];
  for (int i = threadIdx.x; i < NC; i += C::THREADS) s_win2[i] = p.win2[i];
  for (int i = threadIdx.x; i < C::N; i += C::THREADS) s_wenv[i] = p.wenv[i];
  for (int i = threadIdx.x; i < HOP; i += C::THREADS) s_rden[i] = p.den[i];

  const int u = role;
  const int kb_lo = u != 0 ? u : N1;                      // bin of slot s (< N2/2): kb_lo + N1*s
  const int kb_hi = u != 0 ? u : N1 / 2 - (N2 / 2) * N1;  // bin of slot s (>= N2/2): kb_hi + N1*s
  const int rowA = u, rowB = u != 0 ? N1 - u : N1 / 2;    // E rows (k1) of the unit's two columns
  const int64_t T = p.T;
  const int Ti = (int)T;
  const int total_runs = p.batch * rp.runs_per_clip;

#pragma unroll 1
  for (int run = blockIdx.x; run < total_runs; run += gridDim.x) {
    const int clip_i = run / rp.runs_per_clip;
    const int h0 = rp.hop_first + (run - clip_i * rp.runs_per_clip) * rp.run_hops;
    const int h1 = min(h0 + rp.run_hops, rp.hop_end);
    const int nt = (h1 - h0 + C::LEAD + FT - 1) / FT;
    __syncthreads();  // the previous run's last overlap-add has finished (E, carry)
    for (int i = threadIdx.x; i < 3 * HOP; i += C::THREADS) s_carry[i] = 0.0f;  // buffer 0: nothing carried into the run
#pragma unroll 1
    for (int ti = 0; ti < nt; ++ti) {
      const int s0 = h0 - C::LEAD + FT * ti;  // frame of lane-group position 0 == hop of tile-relative q = 0
      const int t = s0 + fr;
      const bool live = t >= 0 && t < Ti;
      if (ti > 0) __syncthreads();  // previous tile's overlap-add has finished reading E
      // Every other tile the spectrum rows of the run's NEXT TWO tiles go to L2: 32 frames = 256 contiguous bytes per bin (128 per
      // plane), the same DRAM access size as the 32-frame kernel — a 16-frame row fetched on its own is a 128-byte access whose
      // neighbour (the next tile's) reaches the DRAM a whole tile later, after the page has been closed.
      if ((ti & 1) == 0 && ti + 1 < nt) {
        int64_t tf = (int64_t)s0 + FT, tl = tf + (ti + 2 < nt ? 2 * FT : FT) - 1;
        tf = tf < 0 ? 0 : tf;
        tl = tl > T - 1 ? T - 1 : tl;
        if (tl >= tf) {
          const int64_t nb = (int64_t)clip_i * p.clip_stride;
          for (int k = threadIdx.x; k <= NC; k += C::THREADS) {
            const int64_t o0 = nb + (int64_t)k * T + tf, o1 = nb + (int64_t)k * T + tl;
            if constexpr (PLANAR) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_re + o0));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_re + o1));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_im + o0));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec_im + o1));
            } else {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec + o0));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec + (o0 + o1) / 2));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.spec + o1));
            }
          }
        }
      }

      // ---- step 1: role = unit u (bin columns u and N1 - u) ----------------------------------------------
      {
        float2 A[N2], B[N2];
        const int64_t stepT = (int64_t)N1 * T, jumpT = (int64_t)(kb_hi - kb_lo) * T;
        const int64_t oa0 = (int64_t)clip_i * p.clip_stride + (live ? t : 0) + (int64_t)kb_lo * T;
        const int64_t ob0 = (int64_t)clip_i * p.clip_stride + (live ? t : 0) + (int64_t)(NC - kb_lo) * T;
        const float2* qa = p.spec + oa0;
        const float2* qb = p.spec + ob0;
        const float *ra = p.spec_re + oa0, *ia = p.spec_im + oa0;
        const float *rb = p.spec_re + ob0, *ib = p.spec_im + ob0;
        const float2 zero2 = make_float2(0.0f, 0.0f);
        float2 xa[N2], xb[N2];
        static_for<0, N2>([&](auto S_) {
          constexpr int s = decltype(S_)::value;
          if (s == N2 / 2) {  // unit 0 switches from column 0 to column N1/2 here (jumpT == 0 for the other units)
            if constexpr (PLANAR) {
              ra += jumpT; ia += jumpT; rb -= jumpT; ib -= jumpT;
            } else {
              qa += jumpT; qb -= jumpT;
            }
          }
          if constexpr (PLANAR) {
            xa[s] = live ? make_float2(B2A_LD(ra), B2A_LD(ia)) : zero2;
            xb[s] = live ? make_float2(B2A_LD(rb), B2A_LD(ib)) : zero2;
            ra += stepT; ia += stepT; rb -= stepT; ib -= stepT;
          } else {
            xa[s] = live ? B2A_LD(qa) : zero2;
            xb[s] = live ? B2A_LD(qb) : zero2;
            qa += stepT; qb -= stepT;
          }
        });
        auto load_at = [&](int64_t i) -> float2 {
          if (!live) return zero2;
          if constexpr (PLANAR) return make_float2(B2A_LD(p.spec_re + i), B2A_LD(p.spec_im + i));
          else return B2A_LD(p.spec + i);
        };
        float2 dc = make_float2(0.0f, 0.0f);
        if (u == 0) {  // Im(DC), Im(Nyquist) are ignored (irfft)
          const int64_t o0 = (int64_t)clip_i * p.clip_stride + t;
          float2 x0 = load_at(o0), xn = load_at(o0 + (int64_t)NC * T);
          if (POLAR && live) {
            x0 = polar_to_complex(p.polar, x0);
            xn = polar_to_complex(p.polar, xn);
          }
          dc = make_float2(x0.x, xn.x);
        }
        if (POLAR && live) {
#pragma unroll
          for (int s = 0; s < N2; ++s) {
            xa[s] = polar_to_complex(p.polar, xa[s]);
            xb[s] = polar_to_complex(p.polar, xb[s]);
          }
        }
        const float4* tp4 = reinterpret_cast<const float4*>(s_twp_al + u * N2);
        auto unpack = [&](auto U0_) {
          constexpr bool U0 = decltype(U0_)::value;
          static_for<0, N2 / 2>([&](auto S_) {
            constexpr int sl0 = 2 * decltype(S_)::value;
            const float4 w2 = tp4[sl0 / 2];  // (c, s) of slots sl0, sl0 + 1: conj(W_N^k) = (cos, sin)(2 pi k / N)
            static_for<0, 2>([&](auto J_) {
              constexpr int s = sl0 + decltype(J_)::value;
              const float2 w = decltype(J_)::value == 0 ? make_float2(w2.x, w2.y) : make_float2(w2.z, w2.w);
              const float2 a = xa[s], b = xb[s];
              const float2 e2 = regs::pfma(b, make_float2(1.0f, -1.0f), a);   // a + conj(b)
              const float2 d = regs::pfma(b, make_float2(-1.0f, 1.0f), a);    // a - conj(b)
              const float2 o2 = regs::cmul(d, w);                             // (a - conj b) * conj(W_N^k)
              const float2 zk = regs::pfma(o2, make_float2(1.0f, -1.0f), regs::pswap(e2));
              const float2 zm = regs::pfma(regs::pswap(e2), make_float2(-1.0f, 1.0f), o2);
              if constexpr (!U0) {
                A[s] = zk;
                B[N2 - 1 - s] = zm;
              } else if constexpr (s < N2 / 2) {
                if constexpr (s + 1 < N2 / 2) A[s + 1] = zk;
                A[N2 - 1 - s] = zm;
              } else {
                B[s - N2 / 2] = zk;
                B[3 * N2 / 2 - 1 - s] = zm;
              }
            });
          });
        };
        if (u == 0) {  // (half-warp divergent in warp 0: its two roles are units 0 and 1)
          unpack(std::true_type{});
          A[0] = make_float2(dc.x - dc.y, dc.x + dc.y);  // Z[0] = (X0 + XN) + i (X0 - XN), swapped
        } else {
          unpack(std::false_type{});
        }
        Dft<N2>::run(A);
        Dft<N2>::run(B);
        const float4* t4a = reinterpret_cast<const float4*>(s_tw1 + (u * 2 + 0) * N2);
        const float4* t4b = reinterpret_cast<const float4*>(s_tw1 + (u * 2 + 1) * N2);
        float2* ea = E + fr * C::EP + rowA * N2;
        float2* eb = E + fr * C::EP + rowB * N2;
        static_for<0, N2 / 2>([&](auto I_) {
          constexpr int n2 = 2 * decltype(I_)::value;
          const float4 ta = t4a[n2 / 2], tb = t4b[n2 / 2];
          ea[n2] = n2 == 0 ? A[0] : regs::cmul(A[n2], make_float2(ta.x, ta.y));
          ea[n2 + 1] = regs::cmul(A[n2 + 1], make_float2(ta.z, ta.w));
          eb[n2] = n2 == 0 ? B[0] : regs::cmul(B[n2], make_float2(tb.x, tb.y));
          eb[n2 + 1] = regs::cmul(B[n2 + 1], make_float2(tb.z, tb.w));
        });
      }
      __syncthreads();  // E[frame][k1][n2] complete
      // ---- step 2: role = residue n2 -------------------------------------------------------------------------
      {
        const int n2 = role;
        float2 v[N1];
        float2* er = E + fr * C::EP + n2;
        static_for<0, N1>([&](auto K_) {
          constexpr int k1 = decltype(K_)::value;
          v[k1] = er[k1 * N2];
        });
        Dft<N1>::run(v);
        constexpr int WPD = 2;
        const unsigned wb_sa = (unsigned)__cvta_generic_to_shared(s_win2 + n2 * N1);
        float4 wq[N1 / 2];
        static_for<0, WPD>([&](auto I_) { wq[decltype(I_)::value] = lds128_pinned<16 * decltype(I_)::value>(wb_sa); });
        static_for<0, N1 / 2>([&](auto I_) {
          constexpr int n1 = 2 * decltype(I_)::value;
          if constexpr (n1 / 2 + WPD < N1 / 2) wq[n1 / 2 + WPD] = lds128_pinned<16 * (n1 / 2 + WPD)>(wb_sa);
          const float4 w = wq[n1 / 2];
          er[n1 * N2] = regs::pmul(regs::pswap(v[n1]), make_float2(w.x, w.y));
          er[(n1 + 1) * N2] = regs::pmul(regs::pswap(v[n1 + 1]), make_float2(w.z, w.w));
        });
      }
      __syncthreads();  // rows of E are now windowed frames in natural sample order

      // ---- step 3: overlap-add; hop q of the tile (q = 0 .. 15) is the sum of the carried partial sum (q < 3) and the quarters
      // 3 - j of frames q - j, j = 3 .. 0 (ascending frame order); two hops per sweep ---------------------------------------
      {
        const float* Yf = reinterpret_cast<const float*>(E);
        const float* cin = s_carry + (ti & 1) * 3 * HOP;
        float* cout = s_carry + ((ti & 1) ^ 1) * 3 * HOP;
        const int hsel = threadIdx.x / (HOP / 2), r = 2 * (threadIdx.x - hsel * (HOP / 2));
        float* const o = p.out + (int64_t)clip_i * p.out_clip_stride;
        const float2 rden = *reinterpret_cast<const float2*>(s_rden + r);
        // interior tile (CTA-uniform): every hop has its four frames, lies inside the run and the output range, the reciprocal
        // envelope table applies and the stores are 8-byte aligned — then hops 4 .. 15 are four loads, three adds, one multiply
        // and one store each (the general loop below spends more instructions on range tests than on the overlap-add)
        const bool interior = ti > 0 && s0 >= 3 && s0 + FT - 1 < Ti && s0 + FT <= h1 && p.rden_ok && p.vec_ok &&
                              (int64_t)s0 * HOP - p.out_start >= 0 && (int64_t)(s0 + FT) * HOP - p.out_start <= p.out_len;
        if (interior) {
          constexpr int D = 2 * C::EP - HOP;  // from quarter j of frame q - j to quarter j + 1 of frame q - j - 1
          float* oj = o + ((int64_t)(s0 + hsel) * HOP + r - p.out_start);
          const float* y = Yf + hsel * (2 * C::EP) + r;
#pragma unroll
          for (int it = 0; it < 2; ++it) {  // hops 0 .. 3: the carried partial sums, fewer in-tile frames
            const int q = 2 * it + hsel;
            float2 num = q < 3 ? *reinterpret_cast<const float2*>(cin + q * HOP + r) : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 3; j >= 0; --j) {
              if (q - j >= 0) {
                const float2 yv = *reinterpret_cast<const float2*>(Yf + (q - j) * (2 * C::EP) + j * HOP + r);
                num.x += yv.x;
                num.y += yv.y;
              }
            }
            *reinterpret_cast<float2*>(oj + 2 * it * HOP) = make_float2(num.x * rden.x, num.y * rden.y);
          }
#pragma unroll
          for (int it = 2; it < FT / 2; ++it) {
            const float* yi = y + 2 * it * (2 * C::EP);
            float2 num = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 3; j >= 0; --j) {
              const float2 yv = *reinterpret_cast<const float2*>(yi - j * D);
              num.x += yv.x;
              num.y += yv.y;
            }
            *reinterpret_cast<float2*>(oj + 2 * it * HOP) = make_float2(num.x * rden.x, num.y * rden.y);
          }
        } else
#pragma unroll 2
        for (int it = 0; it < FT / 2; ++it) {
          const int q = 2 * it + hsel, hh = s0 + q;
          float2 num = q < 3 ? *reinterpret_cast<const float2*>(cin + q * HOP + r) : make_float2(0.0f, 0.0f);
#pragma unroll
          for (int j = 3; j >= 0; --j) {
            if (q - j >= 0) {
              const float2 yv = *reinterpret_cast<const float2*>(Yf + (q - j) * (2 * C::EP) + j * HOP + r);
              num.x += yv.x;
              num.y += yv.y;
            }
          }
          if (hh < h0 || hh >= h1) continue;  // lead-in hops belong to the previous run, trailing ones to the next
          float2 res;
          if (hh - 3 >= 0 && hh < Ti && p.rden_ok) {
            res = make_float2(num.x * rden.x, num.y * rden.y);
          } else {
            float2 den = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 3; j >= 0; --j) {
              const int tt = hh - j;
              if (tt >= 0 && tt < Ti) {
                const float2 wv = *reinterpret_cast<const float2*>(s_wenv + r + j * HOP);
                den.x += wv.x;
                den.y += wv.y;
              }
            }
            if (p.div_clamp) {
              res.x = __fdiv_rn(num.x, fmaxf(den.x, p.div_eps));
              res.y = __fdiv_rn(num.y, fmaxf(den.y, p.div_eps));
            } else {
              res.x = den.x > p.div_eps ? __fdiv_rn(num.x, den.x) : num.x;
              res.y = den.y > p.div_eps ? __fdiv_rn(num.y, den.y) : num.y;
            }
          }
          const int64_t j0 = (int64_t)hh * HOP + r - p.out_start;
          if (p.vec_ok && j0 >= 0 && j0 + 1 < p.out_len) {
            *reinterpret_cast<float2*>(o + j0) = res;
          } else {
            if (j0 >= 0 && j0 < p.out_len) o[j0] = res.x;
            if (j0 + 1 >= 0 && j0 + 1 < p.out_len) o[j0 + 1] = res.y;
          }
        }
        // partial sums of the next tile's hops 0, 1, 2: the quarters of this tile's last three frames that reach past it
        for (int i = threadIdx.x; i < 3 * (HOP / 2); i += C::THREADS) {
          const int qn = i / (HOP / 2), rr = 2 * (i - qn * (HOP / 2));
          float2 sum = make_float2(0.0f, 0.0f);
          for (int j = 3; j > qn; --j) {
            const float2 yv = *reinterpret_cast<const float2*>(Yf + (FT + qn - j) * (2 * C::EP) + j * HOP + rr);
            sum.x += yv.x;
            sum.y += yv.y;
          }
          *reinterpret_cast<float2*>(cout + qn * HOP + rr) = sum;
        }
      }
    }
  }
}

template <class C, bool POLAR, bool PLANAR>
int launch_inv16(b2a_plan* plan, InvFastParams& p, const InvRunParams& rp, cudaStream_t st) {
  const size_t smem = inv16_smem_bytes<C>();
  static SmemAttrOnce attr;
  if (attr.need(plan->device, smem)) {
    B2A_CUDA(cudaFuncSetAttribute(fast_istft16_kernel<C, POLAR, PLANAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need 2 x 89 KB: ask for the shared-memory configuration that holds them (the loads are streaming, L1 is idle)
    static const int carve = getenv("B2A_X_CARVE") ? atoi(getenv("B2A_X_CARVE")) : 80;  // percent of 228 KB: rounds up to the 196 KB configuration (two CTAs)
    if (carve >= 0)
      B2A_CUDA(cudaFuncSetAttribute(fast_istft16_kernel<C, POLAR, PLANAR>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  }
  const int64_t runs = (int64_t)p.batch * rp.runs_per_clip;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(runs, 2 * (int64_t)plan->sm_count));
  fast_istft16_kernel<C, POLAR, PLANAR><<<grid, C::THREADS, smem, st>>>(p, rp);
  B2A_LAUNCHED();
  return B2A_OK;
}

using ICfg1024x16 = ICfg16<32, 16>;  // the same transform as ICfg1024, 16-frame tiles: 256 threads, 2 CTAs / SM

using ICfg1024 = ICfg<32, 16>;  // Vocos / Vocos-mel heads: n_fft 1024, hop 256; 512 threads, 1 CTA / SM

}  // namespace

bool fast_istft_supported(const b2a_plan* plan) {
  const b2a_istft_desc& d = plan->id;
  if (getenv("B2A_FORCE_GENERIC")) return false;
  return d.n_fft == 1024 && d.hop == 256;
}

int fast_istft_init(b2a_plan* plan) {
  using C = ICfg1024;
  const b2a_istft_desc& d = plan->id;
  InvFastState* fs = new InvFastState();
  plan->fast = fs;
  fs->variant = 1;
  constexpr int N1 = C::N1, N2 = C::N2, NC = C::NC, N = C::N, HOP = C::HOP;
  auto expand = [](double a) { return make_float2((float)cos(a), (float)sin(a)); };
  std::vector<float2> twp(C::WARPS * N2), tw1(C::WARPS * 2 * N2), win2(NC);
  std::vector<float> wenv(N), den(HOP);
  for (int u = 0; u < C::WARPS; ++u) {
    const int kb_lo = u != 0 ? u : N1, kb_hi = u != 0 ? u : N1 / 2 - (N2 / 2) * N1;
    for (int s = 0; s < N2; ++s) {
      const int k = (s < N2 / 2 ? kb_lo : kb_hi) + N1 * s;
      twp[u * N2 + s] = expand(2.0 * M_PI * (double)k / (double)N);  // conj(W_N^k)
    }
    const int rowA = u, rowB = u != 0 ? N1 - u : N1 / 2;
    for (int n2 = 0; n2 < N2; ++n2) {  // forward twiddles W_Nc^(n2*k1) (the inverse runs in the swapped domain)
      tw1[(u * 2 + 0) * N2 + n2] = expand(-2.0 * M_PI * (double)((n2 * rowA) % NC) / (double)NC);
      tw1[(u * 2 + 1) * N2 + n2] = expand(-2.0 * M_PI * (double)((n2 * rowB) % NC) / (double)NC);
    }
  }
  const float inv_n = 1.0f / (float)N;  // a power of two for n_fft = 1024: exact, commutes with the window product
  for (int n2 = 0; n2 < N2; ++n2)
    for (int n1 = 0; n1 < N1; ++n1) {
      const int m = N2 * n1 + n2;
      win2[n2 * N1 + n1] = make_float2(plan->h_window[2 * m] * inv_n, plan->h_window[2 * m + 1] * inv_n);
    }
  const bool sq = d.norm_kind == B2A_ISTFT_NORM_WINDOW_SQ;
  for (int n = 0; n < N; ++n) wenv[n] = sq ? plan->h_window[n] * plan->h_window[n] : plan->h_window[n];
  for (int r = 0; r < HOP; ++r) {  // ascending frame order: the oldest frame contributes tap r + 3*hop
    float s = 0.0f;
    for (int j = 0; j < 4; ++j) s += wenv[r + (3 - j) * HOP];
    if (!(s > (d.div_eps > 0.0f ? d.div_eps : 1e-10f))) fs->rden_ok = 0;  // guard active: keep the exact division path
    den[r] = (float)(1.0 / (double)s);
  }
  B2A_CUDA(cudaMalloc(&fs->d_twp, sizeof(float2) * twp.size()));
  B2A_CUDA(cudaMalloc(&fs->d_tw1, sizeof(float2) * tw1.size()));
  B2A_CUDA(cudaMalloc(&fs->d_win2, sizeof(float2) * win2.size()));
  B2A_CUDA(cudaMalloc(&fs->d_wenv, sizeof(float) * wenv.size()));
  B2A_CUDA(cudaMalloc(&fs->d_den, sizeof(float) * den.size()));
  B2A_CUDA(cudaMemcpy(fs->d_twp, twp.data(), sizeof(float2) * twp.size(), cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_tw1, tw1.data(), sizeof(float2) * tw1.size(), cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_win2, win2.data(), sizeof(float2) * win2.size(), cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_wenv, wenv.data(), sizeof(float) * wenv.size(), cudaMemcpyHostToDevice));
  B2A_CUDA(cudaMemcpy(fs->d_den, den.data(), sizeof(float) * den.size(), cudaMemcpyHostToDevice));
  plan->kernel_name = "fast_istft_1024x256";
  return B2A_OK;
}

void fast_istft_destroy(b2a_plan* plan) {
  InvFastState* fs = reinterpret_cast<InvFastState*>(plan->fast);
  if (!fs) return;
  cudaFree(fs->d_twp);
  cudaFree(fs->d_tw1);
  cudaFree(fs->d_win2);
  cudaFree(fs->d_wenv);
  cudaFree(fs->d_den);
  delete fs;
  plan->fast = nullptr;
}

int fast_istft(b2a_plan* plan, const b2a_inverse_args* a, cudaStream_t st) {
  using C = ICfg1024;
  const b2a_istft_desc& d = plan->id;
  InvFastState* fs = reinterpret_cast<InvFastState*>(plan->fast);
  InvFastParams p;
  memset(&p, 0, sizeof(p));
  const int N = d.n_fft, hop = d.hop, F = plan->n_freqs;
  if (a->spec_imag) {
    p.spec_re = reinterpret_cast<const float*>(a->spec);
    p.spec_im = reinterpret_cast<const float*>(a->spec_imag);
  } else {
    p.spec = reinterpret_cast<const float2*>(a->spec);
  }
  p.T = a->num_frames;
  p.clip_stride = a->clip_stride ? a->clip_stride : (int64_t)F * a->num_frames;
  p.batch = a->batch;
  p.norm_sq = d.norm_kind == B2A_ISTFT_NORM_WINDOW_SQ;
  p.div_clamp = d.div_kind == B2A_ISTFT_DIV_CLAMP;
  p.polar = make_polar_spec(d);
  p.div_eps = istft_div_eps(d);
  int64_t ola, start, len;
  b2a_istft_geometry(a->num_frames, N, hop, d.center, d.trim_tail ? a->length : -1, &ola, &start, &len);
  if (!d.trim_tail) {  // ISTFTCache: strip only the front, then [:audio_length]
    start = d.center ? N / 2 : 0;
    len = ola - start;
    if (len < 0) len = 0;
    if (a->length >= 0 && a->length < len) len = a->length;
  }
  if (len <= 0) return B2A_OK;
  p.out_start = start;
  p.out_len = len;
  p.out_clip_stride = a->out_clip_stride ? a->out_clip_stride : len;
  p.out = a->out;
  p.vec_ok = (reinterpret_cast<uintptr_t>(a->out) % 8 == 0) && (p.out_clip_stride % 2 == 0) && (start % 2 == 0);
  p.twp = fs->d_twp;
  p.tw1 = fs->d_tw1;
  p.win2 = fs->d_win2;
  p.wenv = fs->d_wenv;
  p.den = fs->d_den;
  p.rden_ok = fs->rden_ok;
  static const bool use16 = getenv("B2A_X_INV16") != nullptr;  // development: the two-CTA-per-SM kernel with 16-frame tiles (not faster, see above)
  if (use16 && (int64_t)a->batch * (a->num_frames + 3) < (int64_t)1 << 30) {
    using C16 = ICfg1024x16;
    InvRunParams rp;
    rp.hop_first = (int)(start / hop);
    rp.hop_end = (int)((start + len + hop - 1) / hop);  // hops cover OLA coordinates [start, start + len)
    const int hops = rp.hop_end - rp.hop_first;
    // runs: enough of them to balance two CTAs per SM (>= 4 per CTA where the batch allows), each a whole number of tiles
    // after its 3-frame lead-in; at least one tile (13 hops)
    const int64_t want = 4 * 2 * (int64_t)plan->sm_count;
    int rpc = (int)std::max<int64_t>(1, (want + a->batch - 1) / a->batch);
    int tiles = ((hops + rpc - 1) / rpc + C16::LEAD + C16::FT - 1) / C16::FT;
    tiles = tiles < 1 ? 1 : (tiles > 64 ? 64 : tiles);
    rp.run_hops = C16::FT * tiles - C16::LEAD;
    rp.runs_per_clip = (hops + rp.run_hops - 1) / rp.run_hops;
    if (p.polar.polar) return launch_inv16<C16, true, true>(plan, p, rp, st);
    return p.spec == nullptr ? launch_inv16<C16, false, true>(plan, p, rp, st) : launch_inv16<C16, false, false>(plan, p, rp, st);
  }
  // tiles cover OLA coordinates [0, start + len)
  p.tiles_per_clip = (int)((start + len + C::S - 1) / C::S);
  if (p.polar.polar) return launch_inv<C, true, true>(plan, p, st);
  return p.spec == nullptr ? launch_inv<C, false, true>(plan, p, st) : launch_inv<C, false, false>(plan, p, st);
}

}  // namespace b2a
