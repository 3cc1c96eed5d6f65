// Host check of csrc/fft400.cuh: the 20x20 two-pass factorisation + two-real-frames split, against a
// naive fp64 DFT.  Prints "max_rel_err <value>"; tests/test_fft400_host.py builds and runs it with g++.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft400.cuh"
using namespace ttasr;

int main() {
  const double PI = 3.14159265358979323846;
  std::vector<float> a(400), b(400), win(400);
  srand(7);
  for (int n = 0; n < 400; ++n) {
    win[n] = (float)(0.5 - 0.5 * cos(2 * PI * n / 400.0));
    a[n] = win[n] * (float)((rand() / (double)RAND_MAX - 0.5) + 0.3 * sin(2 * PI * 37.3 * n / 400.0));
    b[n] = win[n] * (float)((rand() / (double)RAND_MAX - 0.5) * 0.01 + sin(2 * PI * 50 * n / 400.0));
  }
  // pass 1: for each n2: DFT over n1 of z[20 n1 + n2], then twiddle W400^(n2 k1)
  static float Tr[20][20], Ti[20][20];  // [k1][n2]
  for (int n2 = 0; n2 < 20; ++n2) {
    float xr[20], xi[20], yr[20], yi[20];
    for (int n1 = 0; n1 < 20; ++n1) { xr[n1] = a[20 * n1 + n2]; xi[n1] = b[20 * n1 + n2]; }
    dft20(xr, xi, yr, yi);
    for (int k1 = 0; k1 < 20; ++k1) {
      float c = (float)cos(2 * PI * (n2 * k1) / 400.0), s = (float)-sin(2 * PI * (n2 * k1) / 400.0);
      Tr[k1][n2] = yr[k1] * c - yi[k1] * s;
      Ti[k1][n2] = yr[k1] * s + yi[k1] * c;
    }
  }
  // pass 2: for each k1: DFT over n2 -> Z[k1 + 20 k2]
  std::vector<float> Zr(400), Zi(400);
  for (int k1 = 0; k1 < 20; ++k1) {
    float xr[20], xi[20], yr[20], yi[20];
    for (int n2 = 0; n2 < 20; ++n2) { xr[n2] = Tr[k1][n2]; xi[n2] = Ti[k1][n2]; }
    dft20(xr, xi, yr, yi);
    for (int k2 = 0; k2 < 20; ++k2) { Zr[k1 + 20 * k2] = yr[k2]; Zi[k1 + 20 * k2] = yi[k2]; }
  }
  double max_rel = 0, pmax = 0;
  std::vector<double> pa_ref(201), pb_ref(201);
  for (int k = 0; k <= 200; ++k) {
    std::complex<double> A = 0, B = 0;
    for (int n = 0; n < 400; ++n) {
      std::complex<double> w = std::polar(1.0, -2 * PI * k * n / 400.0);
      A += (double)a[n] * w; B += (double)b[n] * w;
    }
    pa_ref[k] = std::norm(A); pb_ref[k] = std::norm(B);
    pmax = std::max(pmax, std::max(pa_ref[k], pb_ref[k]));
  }
  for (int k = 0; k <= 200; ++k) {
    float pa, pb; int kk = (400 - k) % 400;
    split_power(Zr[k], Zi[k], Zr[kk], Zi[kk], pa, pb);
    // error relative to the frame's peak power: what matters after the max-8 clamp
    max_rel = std::max(max_rel, std::fabs(pa - pa_ref[k]) / std::max(pa_ref[k], 1e-8 * pmax));
    max_rel = std::max(max_rel, std::fabs(pb - pb_ref[k]) / std::max(pb_ref[k], 1e-8 * pmax));
  }
  printf("max_rel_err %.3e\n", max_rel);
  return max_rel < 2e-3 ? 0 : 1;
}
