// b200audio — in-register DFT codelets with compile-time twiddles (host+device so they can be unit-tested
// on the CPU: tests/test_codelets.py compiles csrc/codelet_check.cu for the host).
//
// Dft<R>::run(v): in-place forward DFT (W = exp(-2*pi*i/R)) of R complex values held in registers,
// natural order in and out.  R in {2,3,4,5} are hand-written butterflies; composite sizes are assembled at
// compile time by Cooley-Tukey (twiddles folded to immediates) or, for coprime factors, by the Good-Thomas
// prime-factor map (no twiddles at all): 8=2x4, 16=4x4, 32=4x8, 10=2x5 (PFA), 20=4x5 (PFA), 25=5x5.
//
// On the device every complex value is one packed f32x2 register pair and the butterflies are written with
// Blackwell's packed FP32 instructions (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, sm_100+):
// a complex add is ONE instruction, multiplication by +-i is folded into an FFMA2 with a (+-1, -+1) constant on
// the half-swapped operand (the swap is an operand swizzle, not a MOV).  DFT-20 = 112 instructions instead of
// 224 scalar ones.  The host path (unit test) uses the same formulas in plain C++.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>
#include <utility>

#ifndef B2A_HD
#define B2A_HD __host__ __device__ __forceinline__
#endif

namespace b2a {
namespace regs {

// ---- compile-time trigonometry (double, exact octant reduction + Taylor on |x| <= pi/4) -----------------
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double sin_small(double x) {
  const double x2 = x * x;
  double term = x, sum = x;
  for (int i = 1; i <= 10; ++i) {
    term *= -x2 / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double cos_small(double x) {
  const double x2 = x * x;
  double term = 1.0, sum = 1.0;
  for (int i = 1; i <= 10; ++i) {
    term *= -x2 / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
struct cplx_d {
  double re, im;
};
// exp(+2*pi*i*num/den)
__host__ __device__ constexpr cplx_d unit_root(long long num, long long den) {
  num %= den;
  if (num < 0) num += den;
  const long long q = (4 * num) / den;        // quadrant
  const long long r = 4 * num - q * den;      // angle within quadrant = (pi/2) * r/den
  double c0 = 0, s0 = 0;
  if (2 * r <= den) {
    const double th = (kPi / 2) * (double)r / (double)den;
    c0 = cos_small(th);
    s0 = sin_small(th);
  } else {
    const double th = (kPi / 2) * (double)(den - r) / (double)den;
    c0 = sin_small(th);
    s0 = cos_small(th);
  }
  if (r == 0) { c0 = 1.0; s0 = 0.0; }
  switch (q) {
    case 0: return {c0, s0};
    case 1: return {-s0, c0};
    case 2: return {-c0, -s0};
    default: return {s0, -c0};
  }
}

template <int I, int N, class F>
B2A_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- packed complex primitives -------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pk2(float2 a) {
  u64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 up2(u64_t r) {
  float2 a;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 padd(float2 a, float2 b) {
  u64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(r);
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
  u64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(r);
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
  u64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(r);
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
  u64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return up2(r);
}
#else
inline float2 padd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 psub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 pmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 pfma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
B2A_HD float2 pswap(float2 a) { return make_float2(a.y, a.x); }
B2A_HD float2 pbc(float c) { return make_float2(c, c); }

B2A_HD float2 cadd(float2 a, float2 b) { return padd(a, b); }
B2A_HD float2 csub(float2 a, float2 b) { return psub(a, b); }
// a * b with b given as (re, im): 2 packed instructions
B2A_HD float2 cmul(float2 a, float2 b) { return pfma(pswap(a), make_float2(-b.y, b.y), pmul(a, pbc(b.x))); }
// a * w with the twiddle pre-expanded as (wr, wr, -wi, wi): no negation / broadcast needed at run time
B2A_HD float2 cmul_x(float2 a, float4 w) { return pfma(pswap(a), make_float2(w.z, w.w), pmul(a, make_float2(w.x, w.y))); }
// a * (-i) and a * (+i)
B2A_HD float2 mul_mi(float2 a) { return pmul(pswap(a), make_float2(1.0f, -1.0f)); }
B2A_HD float2 mul_pi(float2 a) { return pmul(pswap(a), make_float2(-1.0f, 1.0f)); }

// v * exp(-2*pi*i*NUM/DEN) with the trivial rotations resolved at compile time
template <int NUM, int DEN>
B2A_HD float2 twiddle(float2 v) {
  constexpr int n = ((NUM % DEN) + DEN) % DEN;
  if constexpr (n == 0) {
    return v;
  } else if constexpr (4 * n == DEN) {  // -i
    return mul_mi(v);
  } else if constexpr (2 * n == DEN) {  // -1
    return pmul(v, pbc(-1.0f));
  } else if constexpr (4 * n == 3 * DEN) {  // +i
    return mul_pi(v);
  } else {
    constexpr cplx_d w = unit_root(-n, DEN);
    constexpr float wr = (float)w.re, wi = (float)w.im;
    return pfma(pswap(v), make_float2(-wi, wi), pmul(v, make_float2(wr, wr)));
  }
}

template <int R>
struct Dft;

template <>
struct Dft<1> {
  static B2A_HD void run(float2 (&)[1]) {}
};
template <>
struct Dft<2> {
  static B2A_HD void run(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = padd(a, b);
    v[1] = psub(a, b);
  }
};
template <>
struct Dft<3> {
  static B2A_HD void run(float2 (&v)[3]) {
    constexpr float s = 0.86602540378443864676f;
    const float2 t1 = padd(v[1], v[2]);
    const float2 t2 = pfma(t1, pbc(-0.5f), v[0]);
    const float2 ds = pswap(psub(v[1], v[2]));
    const float2 t3 = pmul(ds, make_float2(s, -s));  // -i * s * d
    v[0] = padd(v[0], t1);
    v[1] = padd(t2, t3);
    v[2] = psub(t2, t3);
  }
};
template <>
struct Dft<4> {
  static B2A_HD void run(float2 (&v)[4]) {
    const float2 a = padd(v[0], v[2]), b = psub(v[0], v[2]);
    const float2 c = padd(v[1], v[3]), ds = pswap(psub(v[1], v[3]));
    v[0] = padd(a, c);
    v[2] = psub(a, c);
    v[1] = pfma(ds, make_float2(1.0f, -1.0f), b);   // b - i*d
    v[3] = pfma(ds, make_float2(-1.0f, 1.0f), b);   // b + i*d
  }
};
template <>
struct Dft<5> {
  static B2A_HD void run(float2 (&v)[5]) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 a1 = padd(v[1], v[4]), b1s = pswap(psub(v[1], v[4]));
    const float2 a2 = padd(v[2], v[3]), b2s = pswap(psub(v[2], v[3]));
    const float2 x0 = v[0];
    v[0] = padd(padd(x0, a1), a2);
    const float2 p1 = pfma(a2, pbc(c2), pfma(a1, pbc(c1), x0));
    const float2 p2 = pfma(a2, pbc(c1), pfma(a1, pbc(c2), x0));
    // q1 = -i*(s1 b1 + s2 b2), q2 = -i*(s2 b1 - s1 b2) on the half-swapped differences
    const float2 q1 = pfma(b2s, make_float2(s2, -s2), pmul(b1s, make_float2(s1, -s1)));
    const float2 q2 = pfma(b2s, make_float2(-s1, s1), pmul(b1s, make_float2(s2, -s2)));
    v[1] = padd(p1, q1);
    v[4] = psub(p1, q1);
    v[2] = padd(p2, q2);
    v[3] = psub(p2, q2);
  }
};

// Cooley-Tukey in registers: R = RA*RB, n = RB*a + b, k = c + RA*d
template <int RA, int RB>
struct DftCT {
  static constexpr int R = RA * RB;
  static B2A_HD void run(float2 (&v)[R]) {
    float2 t[R];
    static_for<0, RB>([&](auto B_) {
      constexpr int b = decltype(B_)::value;
      float2 u[RA];
      static_for<0, RA>([&](auto A_) {
        constexpr int a = decltype(A_)::value;
        u[a] = v[RB * a + b];
      });
      Dft<RA>::run(u);
      static_for<0, RA>([&](auto C_) {
        constexpr int c = decltype(C_)::value;
        t[b * RA + c] = twiddle<b * c, R>(u[c]);
      });
    });
    static_for<0, RA>([&](auto C_) {
      constexpr int c = decltype(C_)::value;
      float2 s[RB];
      static_for<0, RB>([&](auto B_) {
        constexpr int b = decltype(B_)::value;
        s[b] = t[b * RA + c];
      });
      Dft<RB>::run(s);
      static_for<0, RB>([&](auto D_) {
        constexpr int d = decltype(D_)::value;
        v[c + RA * d] = s[d];
      });
    });
  }
};

__host__ __device__ constexpr int mod_inverse(int a, int m) {
  a %= m;
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 1;
}

// Good-Thomas prime-factor algorithm (gcd(RA,RB)=1): n = (RB*a + RA*b) mod R, k = (RB*ia*c + RA*ib*d) mod R
template <int RA, int RB>
struct DftPFA {
  static constexpr int R = RA * RB;
  static constexpr int IA = mod_inverse(RB, RA), IB = mod_inverse(RA, RB);
  static B2A_HD void run(float2 (&v)[R]) {
    float2 t[R], o[R];
    static_for<0, RB>([&](auto B_) {
      constexpr int b = decltype(B_)::value;
      float2 u[RA];
      static_for<0, RA>([&](auto A_) {
        constexpr int a = decltype(A_)::value;
        u[a] = v[(RB * a + RA * b) % R];
      });
      Dft<RA>::run(u);
      static_for<0, RA>([&](auto C_) {
        constexpr int c = decltype(C_)::value;
        t[b * RA + c] = u[c];
      });
    });
    static_for<0, RA>([&](auto C_) {
      constexpr int c = decltype(C_)::value;
      float2 s[RB];
      static_for<0, RB>([&](auto B_) {
        constexpr int b = decltype(B_)::value;
        s[b] = t[b * RA + c];
      });
      Dft<RB>::run(s);
      static_for<0, RB>([&](auto D_) {
        constexpr int d = decltype(D_)::value;
        o[(RB * IA * c + RA * IB * d) % R] = s[d];
      });
    });
    static_for<0, R>([&](auto K_) {
      constexpr int k = decltype(K_)::value;
      v[k] = o[k];
    });
  }
};

template <> struct Dft<8> { static B2A_HD void run(float2 (&v)[8]) { DftCT<2, 4>::run(v); } };
template <> struct Dft<16> { static B2A_HD void run(float2 (&v)[16]) { DftCT<4, 4>::run(v); } };
template <> struct Dft<32> { static B2A_HD void run(float2 (&v)[32]) { DftCT<4, 8>::run(v); } };
template <> struct Dft<10> { static B2A_HD void run(float2 (&v)[10]) { DftPFA<2, 5>::run(v); } };
template <> struct Dft<20> { static B2A_HD void run(float2 (&v)[20]) { DftPFA<4, 5>::run(v); } };
template <> struct Dft<25> { static B2A_HD void run(float2 (&v)[25]) { DftCT<5, 5>::run(v); } };
template <> struct Dft<6> { static B2A_HD void run(float2 (&v)[6]) { DftPFA<2, 3>::run(v); } };
template <> struct Dft<12> { static B2A_HD void run(float2 (&v)[12]) { DftPFA<4, 3>::run(v); } };
template <> struct Dft<15> { static B2A_HD void run(float2 (&v)[15]) { DftPFA<3, 5>::run(v); } };

}  // namespace regs
}  // namespace b2a
