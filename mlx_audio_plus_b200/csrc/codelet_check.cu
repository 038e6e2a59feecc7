// Host-side check of the in-register DFT codelets (compiled for the CPU by tests/test_codelets.py):
// prints the max relative error of Dft<R> against a float64 direct DFT for every supported R.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fft_regs.cuh"

template <int R>
double check() {
  float2 v[R];
  double re[R], im[R];
  srand(R * 7919);
  for (int i = 0; i < R; ++i) {
    re[i] = rand() / (double)RAND_MAX - 0.5;
    im[i] = rand() / (double)RAND_MAX - 0.5;
    v[i] = make_float2((float)re[i], (float)im[i]);
    re[i] = v[i].x;
    im[i] = v[i].y;
  }
  b2a::regs::Dft<R>::run(v);
  double maxerr = 0, maxmag = 0;
  for (int k = 0; k < R; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < R; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % R) / R;
      sr += re[n] * cos(a) - im[n] * sin(a);
      si += re[n] * sin(a) + im[n] * cos(a);
    }
    maxerr = fmax(maxerr, hypot(v[k].x - sr, v[k].y - si));
    maxmag = fmax(maxmag, hypot(sr, si));
  }
  return maxerr / maxmag;
}

int main() {
  printf("2 %.3e\n3 %.3e\n4 %.3e\n5 %.3e\n6 %.3e\n8 %.3e\n10 %.3e\n12 %.3e\n15 %.3e\n16 %.3e\n20 %.3e\n25 %.3e\n32 %.3e\n",
         check<2>(), check<3>(), check<4>(), check<5>(), check<6>(), check<8>(), check<10>(), check<12>(), check<15>(),
         check<16>(), check<20>(), check<25>(), check<32>());
  return 0;
}
