// 400-point complex FFT building blocks for the log-mel front end, written so that the same code compiles for
// the device (registers only, fully unrolled) and for the host (tests/test_fft400_host.py builds a g++ harness).
//
// Factorisation: 400 = 20 x 20 (two passes with one shared-memory transpose between them), and each 20-point DFT
// is a twiddle-free Good-Thomas 4 x 5 prime-factor transform (gcd(4,5)=1): five radix-4 and four radix-5
// butterflies.  Two real frames ride one complex transform (frame a in Re, frame b in Im) and are separated
// afterwards from the Hermitian halves.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TT_HD __host__ __device__ __forceinline__
#else
#define TT_HD inline
#endif

namespace ttasr {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kNfreq = 201;
constexpr int kRadix = 20;

// forward 5-point DFT (e^{-2 pi i jk/5}), in place
TT_HD void dft5(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3, float& r4,
                float& i4) {
  const float c1 = 0.30901699437494742f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357f;   // sin(2pi/5)
  const float s2 = 0.58778525229247313f;   // sin(4pi/5)
  const float t1r = r1 + r4, t1i = i1 + i4;
  const float t2r = r2 + r3, t2i = i2 + i3;
  const float t3r = r1 - r4, t3i = i1 - i4;
  const float t4r = r2 - r3, t4i = i2 - i3;
  const float m1r = r0 + c1 * t1r + c2 * t2r, m1i = i0 + c1 * t1i + c2 * t2i;
  const float m2r = r0 + c2 * t1r + c1 * t2r, m2i = i0 + c2 * t1i + c1 * t2i;
  const float n1r = s1 * t3r + s2 * t4r, n1i = s1 * t3i + s2 * t4i;
  const float n2r = s2 * t3r - s1 * t4r, n2i = s2 * t3i - s1 * t4i;
  r0 = r0 + t1r + t2r;
  i0 = i0 + t1i + t2i;
  // y1 = m1 - i n1, y4 = m1 + i n1, y2 = m2 - i n2, y3 = m2 + i n2   (-i(a+ib) = b - ia)
  r1 = m1r + n1i; i1 = m1i - n1r;
  r4 = m1r - n1i; i4 = m1i + n1r;
  r2 = m2r + n2i; i2 = m2i - n2r;
  r3 = m2r - n2i; i3 = m2i + n2r;
}

// forward 4-point DFT, in place
TT_HD void dft4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
  const float ar = r0 + r2, ai = i0 + i2;
  const float br = r0 - r2, bi = i0 - i2;
  const float cr = r1 + r3, ci = i1 + i3;
  const float dr = r1 - r3, di = i1 - i3;
  r0 = ar + cr; i0 = ai + ci;
  r2 = ar - cr; i2 = ai - ci;
  // y1 = b - i d, y3 = b + i d
  r1 = br + di; i1 = bi - dr;
  r3 = br - di; i3 = bi + dr;
}

// forward 20-point DFT, natural order in -> natural order out.
// input index n = (5a + 4b) mod 20, output index k = (5 ka + 16 kb) mod 20  (CRT maps; no twiddles).
TT_HD void dft20(const float (&xr)[20], const float (&xi)[20], float (&yr)[20], float (&yi)[20]) {
  float ur[4][5], ui[4][5];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      ur[a][b] = xr[(5 * a + 4 * b) % 20];
      ui[a][b] = xi[(5 * a + 4 * b) % 20];
    }
    dft5(ur[a][0], ui[a][0], ur[a][1], ui[a][1], ur[a][2], ui[a][2], ur[a][3], ui[a][3], ur[a][4], ui[a][4]);
  }
#pragma unroll
  for (int kb = 0; kb < 5; ++kb) {
    dft4(ur[0][kb], ui[0][kb], ur[1][kb], ui[1][kb], ur[2][kb], ui[2][kb], ur[3][kb], ui[3][kb]);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
      yr[(5 * ka + 16 * kb) % 20] = ur[ka][kb];
      yi[(5 * ka + 16 * kb) % 20] = ui[ka][kb];
    }
  }
}

#if defined(__CUDACC__)
// ---- the same transforms on packed (re, im) pairs: one FADD2 / FFMA2 / FMUL2 issue slot (sm_100 f32x2) does the real
// and the imaginary lane of a butterfly.  138 issue slots per 20-point DFT instead of 224 (26 of them the half-swaps
// that multiplication by -i costs in this representation).  Same arithmetic per lane as the scalar versions above.
typedef unsigned long long c32_t;  // {re (low 32 bits), im (high)}
__device__ __forceinline__ c32_t c_pack(float re, float im) { c32_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ void c_unpack(c32_t v, float& re, float& im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v)); }
__device__ __forceinline__ c32_t c_add(c32_t a, c32_t b) { c32_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c32_t c_sub(c32_t a, c32_t b) { c32_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c32_t c_mul(c32_t a, c32_t b) { c32_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c32_t c_fma(c32_t a, c32_t b, c32_t c) { c32_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ c32_t c_swap(c32_t a) { float re, im; c_unpack(a, re, im); return c_pack(im, re); }

__device__ __forceinline__ void dft5_packed(c32_t& z0, c32_t& z1, c32_t& z2, c32_t& z3, c32_t& z4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const c32_t C1 = c_pack(c1, c1), C2 = c_pack(c2, c2);
  const c32_t S1N = c_pack(s1, -s1), S2N = c_pack(s2, -s2), S1M = c_pack(-s1, s1);
  const c32_t t1 = c_add(z1, z4), t2 = c_add(z2, z3);
  const c32_t t3s = c_swap(c_sub(z1, z4)), t4s = c_swap(c_sub(z2, z3));   // (im, re)
  const c32_t m1 = c_fma(C2, t2, c_fma(C1, t1, z0));
  const c32_t m2 = c_fma(C1, t2, c_fma(C2, t1, z0));
  // -i n1 = (n1.im, -n1.re) with n1 = s1 t3 + s2 t4;  -i n2 with n2 = s2 t3 - s1 t4
  const c32_t n1 = c_fma(t4s, S2N, c_mul(t3s, S1N));
  const c32_t n2 = c_fma(t4s, S1M, c_mul(t3s, S2N));
  z0 = c_add(c_add(z0, t1), t2);
  z1 = c_add(m1, n1);
  z4 = c_sub(m1, n1);
  z2 = c_add(m2, n2);
  z3 = c_sub(m2, n2);
}

__device__ __forceinline__ void dft4_packed(c32_t& z0, c32_t& z1, c32_t& z2, c32_t& z3) {
  const c32_t a = c_add(z0, z2), b = c_sub(z0, z2), c = c_add(z1, z3);
  const c32_t ds = c_swap(c_sub(z1, z3));   // (d.im, d.re)
  z0 = c_add(a, c);
  z2 = c_sub(a, c);
  z1 = c_fma(ds, c_pack(1.0f, -1.0f), b);   // b - i d = (b.re + d.im, b.im - d.re)
  z3 = c_fma(ds, c_pack(-1.0f, 1.0f), b);   // b + i d
}

// forward 20-point DFT on packed pairs, natural order in -> natural order out (index maps as dft20)
__device__ __forceinline__ void dft20_packed(const c32_t (&x)[20], c32_t (&y)[20]) {
  c32_t u[4][5];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 5; ++b) u[a][b] = x[(5 * a + 4 * b) % 20];
    dft5_packed(u[a][0], u[a][1], u[a][2], u[a][3], u[a][4]);
  }
#pragma unroll
  for (int kb = 0; kb < 5; ++kb) {
    dft4_packed(u[0][kb], u[1][kb], u[2][kb], u[3][kb]);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) y[(5 * ka + 16 * kb) % 20] = u[ka][kb];
  }
}
#endif  // __CUDACC__

// power spectra of the two real frames packed as z = a + i b, from Z[k] and Z[(400-k) mod 400]
TT_HD void split_power(float zr, float zi, float wr, float wi, float& pa, float& pb) {
  const float ar = zr + wr, ai = zi - wi;  // 2 * A[k]
  const float br = zi + wi, bi = zr - wr;  // 2 * B[k] (up to a unit factor)
  pa = 0.25f * (ar * ar + ai * ai);
  pb = 0.25f * (br * br + bi * bi);
}

}  // namespace ttasr
