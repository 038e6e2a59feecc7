// b200audio — in-register DFT codelets with compile-time twiddles (host+device so they can be unit-tested
// on the CPU: tests/test_codelets.py compiles csrc/codelet_check.cu for the host).
//
// Dft<R>::run(v): in-place forward DFT (W = exp(-2*pi*i/R)) of R complex values held in registers,
// natural order in and out.  R in {2,3,4,5} are hand-written butterflies; composite sizes are assembled at
// compile time by Cooley-Tukey (twiddles folded to immediates) or, for coprime factors, by the Good-Thomas
// prime-factor map (no twiddles at all): 8=2x4, 16=4x4, 32=4x8, 10=2x5 (PFA), 20=4x5 (PFA), 25=5x5.
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

B2A_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B2A_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B2A_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// v * exp(-2*pi*i*NUM/DEN) with the trivial rotations resolved at compile time
template <int NUM, int DEN>
B2A_HD float2 twiddle(float2 v) {
  constexpr int n = ((NUM % DEN) + DEN) % DEN;
  if constexpr (n == 0) {
    return v;
  } else if constexpr (4 * n == DEN) {  // -i
    return make_float2(v.y, -v.x);
  } else if constexpr (2 * n == DEN) {  // -1
    return make_float2(-v.x, -v.y);
  } else if constexpr (4 * n == 3 * DEN) {  // +i
    return make_float2(-v.y, v.x);
  } else {
    constexpr cplx_d w = unit_root(-n, DEN);
    constexpr float wr = (float)w.re, wi = (float)w.im;
    return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
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
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <>
struct Dft<3> {
  static B2A_HD void run(float2 (&v)[3]) {
    constexpr float s = 0.86602540378443864676f;
    const float2 t1 = cadd(v[1], v[2]);
    const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    const float2 d = csub(v[1], v[2]);
    const float2 t3 = make_float2(s * d.y, -s * d.x);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
  }
};
template <>
struct Dft<4> {
  static B2A_HD void run(float2 (&v)[4]) {
    const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    const float2 md = make_float2(d.y, -d.x);  // -i*d
    v[0] = cadd(a, c);
    v[2] = csub(a, c);
    v[1] = cadd(b, md);
    v[3] = csub(b, md);
  }
};
template <>
struct Dft<5> {
  static B2A_HD void run(float2 (&v)[5]) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
    const float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
    const float2 x0 = v[0];
    v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
    const float2 p1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
    const float2 p2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
    const float2 q1 = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
    const float2 q2 = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
    v[1] = cadd(p1, q1);
    v[4] = csub(p1, q1);
    v[2] = cadd(p2, q2);
    v[3] = csub(p2, q2);
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
