// Host check of rlao_b200/csrc/fft_codelets.cuh against a naive float64 DFT:  nvcc -o /tmp/chk tools/check_fft_codelets.cu && /tmp/chk
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../rlao_b200/csrc/fft_codelets.cuh"
using namespace aoenv::fftc;

template <int N>
double check() {
  cpx x[N];
  double xr[N], xi[N];
  for (int n = 0; n < N; ++n) { xr[n] = drand48() - 0.5; xi[n] = drand48() - 0.5; x[n] = mk((float)xr[n], (float)xi[n]); }
  Dft<N>::run(x);
  double err = 0, mag = 0;
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      const double a = -2.0 * M_PI * n * k / N;
      sr += xr[n] * cos(a) - xi[n] * sin(a);
      si += xr[n] * sin(a) + xi[n] * cos(a);
    }
    err = fmax(err, fmax(fabs(sr - x[k].x), fabs(si - x[k].y)));
    mag = fmax(mag, fmax(fabs(sr), fabs(si)));
  }
  return err / mag;
}

int main() {
  double e4 = check<4>(), e8 = check<8>(), e9 = check<9>(), e16 = check<16>(), e18 = check<18>();
  printf("rel err: dft4 %.2e dft8 %.2e dft9 %.2e dft16 %.2e dft18 %.2e\n", e4, e8, e9, e16, e18);
  return (e4 < 1e-6 && e8 < 1e-6 && e9 < 1e-6 && e16 < 1e-6 && e18 < 1e-6) ? 0 : 1;
}
