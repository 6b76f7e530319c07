// Small in-register DFTs (2, 3, 4, 8, 9, 16, 18 points) used by the Pyramid transform kernels (pyr.cu): fully unrolled,
// literal twiddles.  Forward transform X[k] = sum_n x[n] exp(-2 pi i n k / N), natural order in and out.  The inverse is
// taken as conj(DFT(conj(x))) by the callers.  __host__ __device__ so that the arithmetic can be checked on the host
// (tools/check_fft_codelets.cu).
#pragma once
#include <cuda_runtime.h>

namespace aoenv {
namespace fftc {

struct __align__(8) cpx { float x, y; };
__host__ __device__ __forceinline__ cpx mk(float a, float b) { cpx c; c.x = a; c.y = b; return c; }
// On the device a complex number lives in one 64-bit register pair and additions / subtractions / real scalings are single
// packed FP32 instructions (Blackwell add.f32x2 / fma.rn.f32x2: two IEEE operations per issue slot).
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long as_u64(cpx a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ cpx as_cpx(unsigned long long v) { return *reinterpret_cast<cpx*>(&v); }
__device__ __forceinline__ cpx operator+(cpx a, cpx b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_cpx(d);
}
__device__ __forceinline__ cpx operator-(cpx a, cpx b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_cpx(d);
}
// a + s * b, s real
__device__ __forceinline__ cpx axpy(float s, cpx b, cpx a) {
  unsigned long long d;
  const cpx ss = mk(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(as_u64(ss)), "l"(as_u64(b)), "l"(as_u64(a)));
  return as_cpx(d);
}
// a + (sx * b.x, sy * b.y)
__device__ __forceinline__ cpx axpy2(float sx, float sy, cpx b, cpx a) {
  unsigned long long d;
  const cpx ss = mk(sx, sy);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(as_u64(ss)), "l"(as_u64(b)), "l"(as_u64(a)));
  return as_cpx(d);
}
#else
__host__ __device__ __forceinline__ cpx operator+(cpx a, cpx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cpx operator-(cpx a, cpx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cpx axpy(float s, cpx b, cpx a) { return mk(a.x + s * b.x, a.y + s * b.y); }
__host__ __device__ __forceinline__ cpx axpy2(float sx, float sy, cpx b, cpx a) { return mk(a.x + sx * b.x, a.y + sy * b.y); }
#endif
__host__ __device__ __forceinline__ cpx swap(cpx a) { return mk(a.y, a.x); }
// a * (wr + i wi) = wr * a + wi * (-a.y, a.x)
__host__ __device__ __forceinline__ cpx cmul(cpx a, float wr, float wi) {
  return axpy2(-wi, wi, swap(a), axpy(wr, a, mk(0.f, 0.f)));
}
__host__ __device__ __forceinline__ cpx mul_neg_i(cpx a) { return mk(a.y, -a.x); }     // a * (-i)
__host__ __device__ __forceinline__ cpx conj(cpx a) { return mk(a.x, -a.y); }

__host__ __device__ __forceinline__ void dft2(cpx& a, cpx& b) {
  const cpx t = a - b;
  a = a + b;
  b = t;
}

// 3-point: y0 = a + b + c; y1,2 = a - (b + c)/2 -+ i (sqrt3/2) (b - c)
__host__ __device__ __forceinline__ void dft3(cpx& a, cpx& b, cpx& c) {
  constexpr float s = 0.8660254037844386f;
  const cpx t = b + c, d = swap(b - c);
  const cpx m = axpy(-0.5f, t, a);
  a = a + t;
  b = axpy2(s, -s, d, m);                       // m - i s (b - c)
  c = axpy2(-s, s, d, m);                       // m + i s (b - c)
}

__host__ __device__ __forceinline__ void dft4(cpx (&x)[4]) {
  cpx a = x[0], b = x[2], c = x[1], d = x[3];
  dft2(a, b);                                   // a = x0 + x2, b = x0 - x2
  dft2(c, d);                                   // c = x1 + x3, d = x1 - x3
  const cpx e = mul_neg_i(d);
  x[0] = a + c; x[2] = a - c; x[1] = b + e; x[3] = b - e;
}

__host__ __device__ __forceinline__ void dft8(cpx (&x)[8]) {
  constexpr float r = 0.7071067811865476f;
  cpx e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
  dft4(e);
  dft4(o);
  o[1] = cmul(o[1], r, -r);
  o[2] = mul_neg_i(o[2]);
  o[3] = cmul(o[3], -r, -r);
#pragma unroll
  for (int k = 0; k < 4; ++k) { x[k] = e[k] + o[k]; x[k + 4] = e[k] - o[k]; }
}

__host__ __device__ __forceinline__ void dft16(cpx (&x)[16]) {
  constexpr float c1 = 0.9238795325112867f, s1 = 0.3826834323650898f, r = 0.7071067811865476f;
  cpx e[8], o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
  dft8(e);
  dft8(o);
  o[1] = cmul(o[1], c1, -s1);
  o[2] = cmul(o[2], r, -r);
  o[3] = cmul(o[3], s1, -c1);
  o[4] = mul_neg_i(o[4]);
  o[5] = cmul(o[5], -s1, -c1);
  o[6] = cmul(o[6], -r, -r);
  o[7] = cmul(o[7], -c1, -s1);
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = e[k] + o[k]; x[k + 8] = e[k] - o[k]; }
}

// 9 = 3 x 3 (Cooley-Tukey): n = 3 n1 + n2, k = k1 + 3 k2
__host__ __device__ __forceinline__ void dft9(cpx (&x)[9]) {
  // twiddles W9^(n2 k1), W9 = exp(-2 pi i / 9)
  constexpr float c1 = 0.766044443118978f, s1 = 0.6427876096865393f;      // cos, sin of 2 pi / 9
  constexpr float c2 = 0.17364817766693041f, s2 = 0.984807753012208f;     // 4 pi / 9
  constexpr float c4 = -0.9396926207859083f, s4 = 0.3420201433256689f;    // 8 pi / 9
  cpx a[3][3];                                  // a[n2][k1] = DFT3 over n1 of x[3 n1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 3; ++n2) {
    cpx p = x[n2], q = x[3 + n2], r = x[6 + n2];
    dft3(p, q, r);
    a[n2][0] = p; a[n2][1] = q; a[n2][2] = r;
  }
  a[1][1] = cmul(a[1][1], c1, -s1);
  a[1][2] = cmul(a[1][2], c2, -s2);
  a[2][1] = cmul(a[2][1], c2, -s2);
  a[2][2] = cmul(a[2][2], c4, -s4);
#pragma unroll
  for (int k1 = 0; k1 < 3; ++k1) {
    cpx p = a[0][k1], q = a[1][k1], r = a[2][k1];
    dft3(p, q, r);                              // over n2 -> k2
    x[k1] = p; x[k1 + 3] = q; x[k1 + 6] = r;
  }
}

// 18 = 2 x 9: even / odd samples, X[k] = E[k mod 9] + W18^k O[k mod 9]
__host__ __device__ __forceinline__ void dft18(cpx (&x)[18]) {
  constexpr float c[9] = {1.0f, 0.9396926207859084f, 0.766044443118978f, 0.5f, 0.17364817766693041f, -0.17364817766693036f,
                          -0.5f, -0.7660444431189779f, -0.9396926207859083f};
  constexpr float s[9] = {0.0f, 0.3420201433256687f, 0.6427876096865393f, 0.8660254037844386f, 0.984807753012208f,
                          0.984807753012208f, 0.8660254037844387f, 0.6427876096865395f, 0.3420201433256689f};
  cpx e[9], o[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
  dft9(e);
  dft9(o);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const cpx t = cmul(o[k], c[k], -s[k]);
    x[k] = e[k] + t;
    x[k + 9] = e[k] - t;
  }
}

template <int N> struct Dft;
template <> struct Dft<4> { __host__ __device__ static __forceinline__ void run(cpx (&x)[4]) { dft4(x); } };
template <> struct Dft<8> { __host__ __device__ static __forceinline__ void run(cpx (&x)[8]) { dft8(x); } };
template <> struct Dft<9> { __host__ __device__ static __forceinline__ void run(cpx (&x)[9]) { dft9(x); } };
template <> struct Dft<16> { __host__ __device__ static __forceinline__ void run(cpx (&x)[16]) { dft16(x); } };
template <> struct Dft<18> { __host__ __device__ static __forceinline__ void run(cpx (&x)[18]) { dft18(x); } };

}  // namespace fftc
}  // namespace aoenv
