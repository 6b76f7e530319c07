// Deformable-mirror surface for the reference's default geometry (OOPAO/DeformableMirror.py:286-305,494-514): actuators
// on a Cartesian grid, axis-aligned Gaussian influence functions exp(-a (X-x0)^2 - c (Y-y0)^2).  Then
//     OPD[y][x] = sum_i gy[i][y] * ( sum_j C[i][j] * gx[j][x] ),        C = command image (0 at invalid actuators),
// i.e. modes @ coefs factorises into two small banded products.  Terms beyond `cut` pixels from the actuator are
// dropped only where the Gaussian is below 2^-30 of its peak (6.5 sigma), i.e. under float32 resolution of the sum, so
// the result equals the dense product to rounding.  One CTA per environment; the intermediate T = C gx lives in shared
// memory.  DMs with rotation / anamorphosis / custom modes use the dense tensor-core GEMM (gemm_tc.cu) instead.
#include "common.cuh"

namespace aoenv {

__global__ void __launch_bounds__(256)
dm_separable_kernel(const float* __restrict__ coefs, int ldc, const int32_t* __restrict__ act_pos, int nA, int nAct,
                    const float* __restrict__ gx, const float* __restrict__ gy, const int2* __restrict__ band_x,
                    const int2* __restrict__ band_y, int R, int xw, float* __restrict__ opd) {
  const int x0 = blockIdx.x * xw;                 // column slab of this CTA
  extern __shared__ __align__(16) float sm[];
  float* sC = sm;                       // [nAct][nAct]
  float* sT = sm + ((nAct * nAct + 3) & ~3);   // [nAct][xw]   (xw = columns handled by this CTA, multiple of 4)
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < nAct * nAct; k += blockDim.x) sC[k] = 0.f;
  __syncthreads();
  for (int k = threadIdx.x; k < nA; k += blockDim.x) sC[__ldg(&act_pos[k])] = __ldg(&coefs[(size_t)b * ldc + k]);
  __syncthreads();
  // stage 1: T[i][x] = sum_{j in band(x)} C[i][j] gx[j][x]
  for (int idx = threadIdx.x; idx < nAct * xw; idx += blockDim.x) {
    const int i = idx / xw, xl = idx - i * xw;
    const int x = x0 + xl;
    float t = 0.f;
    if (x < R) {
      const int2 bd = __ldg(&band_x[x]);
      for (int j = bd.x; j <= bd.y; ++j) t = fmaf(sC[i * nAct + j], __ldg(&gx[(size_t)j * R + x]), t);
    }
    sT[i * xw + xl] = t;
  }
  __syncthreads();
  // stage 2: OPD[y][x..x+3] = sum_{i in band(y)} gy[i][y] T[i][x..x+3]
  const int nq = xw >> 2;
  float* __restrict__ out = opd + (size_t)b * R * R;
  for (int idx = threadIdx.x; idx < R * nq; idx += blockDim.x) {
    const int y = idx / nq, q = idx - y * nq;
    const int2 bd = __ldg(&band_y[y]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = bd.x; i <= bd.y; ++i) {
      const float g = __ldg(&gy[(size_t)i * R + y]);
      const float4 t = *reinterpret_cast<const float4*>(&sT[i * xw + 4 * q]);
      acc.x = fmaf(g, t.x, acc.x); acc.y = fmaf(g, t.y, acc.y); acc.z = fmaf(g, t.z, acc.z); acc.w = fmaf(g, t.w, acc.w);
    }
    const int x = x0 + 4 * q;
    float* __restrict__ o = out + (size_t)y * R + x;
    if (x + 3 < R && (R & 3) == 0) {
      *reinterpret_cast<float4*>(o) = acc;
    } else {
      if (x < R) o[0] = acc.x;
      if (x + 1 < R) o[1] = acc.y;
      if (x + 2 < R) o[2] = acc.z;
      if (x + 3 < R) o[3] = acc.w;
    }
  }
}

// Banded variant: per pixel column x the (at most W) contributing actuator columns start at j0x[x] with weights
// wx[x][0..W); per pair of pixel rows (2k, 2k+1) the contributing actuator rows start at i0y[k] with weights
// wyp[k][0..1][0..W).  Fixed trip counts -> fully unrolled, every load independent; each thread produces a 2 x 4
// block of the surface, so one 128-bit shared-memory load of T feeds 8 FMAs.
template <int W>
__global__ void __launch_bounds__(256, 3)
dm_separable_banded_kernel(const float* __restrict__ coefs, int ldc, const int32_t* __restrict__ act_pos, int nA, int nAct,
                           const float* __restrict__ wx, const int32_t* __restrict__ j0x, const float* __restrict__ wyp,
                           const int32_t* __restrict__ i0y, int R, int xw, float* __restrict__ opd) {
  const int x0 = blockIdx.x * xw;                 // column slab of this CTA
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  float* sC = sm;
  float* sT = sm + ((nAct * nAct + 3) & ~3);
  float* sW = sT + nAct * xw;                     // [R/2][2][W] row weights of stage 2
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < nAct * nAct; k += blockDim.x) sC[k] = 0.f;
  for (int k = threadIdx.x; k < (R >> 1) * 2 * W / 4; k += blockDim.x)
    reinterpret_cast<float4*>(sW)[k] = __ldg(reinterpret_cast<const float4*>(wyp) + k);
  __syncthreads();
  for (int k = threadIdx.x; k < nA; k += blockDim.x) sC[__ldg(&act_pos[k])] = __ldg(&coefs[(size_t)b * ldc + k]);
  __syncthreads();
  // stage 1: lanes along x (coalesced weight loads), loop over actuator rows i
  for (int xl = threadIdx.x; xl < xw; xl += blockDim.x) {
    const int x = x0 + xl;
    float w[W];
    int j0 = 0;
    if (x < R) {
      j0 = __ldg(&j0x[x]);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wx + (size_t)x * W) + q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < W; ++q) w[q] = 0.f;
    }
    for (int i = 0; i < nAct; ++i) {
      const float* __restrict__ c = sC + i * nAct;
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < W; ++q) t = fmaf(c[min(j0 + q, nAct - 1)], w[q], t);
      sT[i * xw + xl] = t;
    }
  }
  __syncthreads();
  // stage 2: thread (q, chunk) owns four neighbouring columns and walks down a chunk of row pairs.  The W rows of T its
  // band touches stay in registers and are re-read from shared memory only when the band start moves (once per actuator
  // pitch), not once per output pair.  Lanes of a warp share k, so the weight loads are warp-uniform.  Packed FP32: each FFMA2 updates two neighbouring pixels with the (broadcast) row weight.
  const int nq = xw >> 2, npair = R >> 1;
  const int chunks = max(1, (int)blockDim.x / nq);
  const int per = (npair + chunks - 1) / chunks;
  float* __restrict__ out = opd + (size_t)b * R * R;
  for (int item = threadIdx.x; item < nq * chunks; item += blockDim.x) {
    const int ch = item / nq, q = item - ch * nq;
    const int k_end = min(npair, (ch + 1) * per);
    float4 win[W];
    int base = -(1 << 20);
    for (int k = ch * per; k < k_end; ++k) {
      const int i0 = __ldg(&i0y[k]);
      if (i0 != base) {                     // the band start moves once per actuator pitch (every few row pairs)
#pragma unroll
        for (int t = 0; t < W; ++t) win[t] = *reinterpret_cast<const float4*>(&sT[min(i0 + t, nAct - 1) * xw + 4 * q]);
        base = i0;
      }
      float2 a0 = make_float2(0.f, 0.f), b0 = a0, a1 = a0, b1 = a0;
#pragma unroll
      for (int t4 = 0; t4 < W / 4; ++t4) {
        const float4 u0 = reinterpret_cast<const float4*>(sW + (size_t)k * 2 * W)[t4];
        const float4 u1 = reinterpret_cast<const float4*>(sW + (size_t)k * 2 * W + W)[t4];
        const float w0[4] = {u0.x, u0.y, u0.z, u0.w}, w1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = win[4 * t4 + j];
          const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
          a0 = fma2(dup2(w0[j]), lo, a0); b0 = fma2(dup2(w0[j]), hi, b0);
          a1 = fma2(dup2(w1[j]), lo, a1); b1 = fma2(dup2(w1[j]), hi, b1);
        }
      }
      const float4 r0 = make_float4(a0.x, a0.y, b0.x, b0.y), r1 = make_float4(a1.x, a1.y, b1.x, b1.y);
      const int x = x0 + 4 * q;
      float* __restrict__ o0 = out + (size_t)(2 * k) * R + x;
      float* __restrict__ o1 = o0 + R;
      if (x + 3 < R && (R & 3) == 0) {
        *reinterpret_cast<float4*>(o0) = r0;
        *reinterpret_cast<float4*>(o1) = r1;
      } else {
        const float c0[4] = {r0.x, r0.y, r0.z, r0.w}, c1[4] = {r1.x, r1.y, r1.z, r1.w};
        for (int c = 0; c < 4; ++c)
          if (x + c < R) { o0[c] = c0[c]; o1[c] = c1[c]; }
      }
    }
  }
}

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_dm_surface_separable(const float* coefs, int ldc, const int32_t* act_pos, int nA, int nAct,
                                          const float* gx, const float* gy, const int32_t* band_x, const int32_t* band_y,
                                          const float* wx, const int32_t* j0x, const float* wyp, const int32_t* i0y, int W,
                                          int B, int R, float* opd, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && nAct > 0 && nA > 0 && nA <= nAct * nAct && ldc >= nA, "dm_surface_separable: bad shape");
  const bool banded = (W == 12 || W == 16) && (R % 2 == 0) && wx && j0x && wyp && i0y;
  // split the columns across CTAs so that C and T fit in shared memory (and small batches still fill the GPU)
  int parts = 1;
  auto smem_for = [&](int p) {
    const int xw = (((R + p - 1) / p) + 3) / 4 * 4;
    return (size_t)(((nAct * nAct + 3) & ~3) + nAct * xw + (banded ? (R / 2) * 2 * W : 0)) * sizeof(float);
  };
  while (smem_for(parts) > 96 * 1024 || (B * parts < 2 * kNumSMs && parts < 8)) ++parts;
  const int xw = (((R + parts - 1) / parts) + 3) / 4 * 4;
  const size_t smem = smem_for(parts);
  AOENV_CHECK_ARG(smem <= 200 * 1024, "dm_surface_separable: %d actuators across do not fit in shared memory", nAct);
  const int which = !banded ? 0 : (W == 12 ? 1 : 2);
  if (smem > 48 * 1024) {       // per device and size: cheap enough to repeat at every launch
    cudaError_t e = which == 0 ? cudaFuncSetAttribute(dm_separable_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : which == 1 ? cudaFuncSetAttribute(dm_separable_banded_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                               : cudaFuncSetAttribute(dm_separable_banded_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "dm_surface_separable smem attribute: %s", cudaGetErrorString(e));
  }
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid(parts, B);
  if (which == 1)
    AOENV_LAUNCH(dm_separable_banded_kernel<12>, grid, 256, smem, s, coefs, ldc, act_pos, nA, nAct, wx, j0x, wyp, i0y, R, xw, opd);
  else if (which == 2)
    AOENV_LAUNCH(dm_separable_banded_kernel<16>, grid, 256, smem, s, coefs, ldc, act_pos, nA, nAct, wx, j0x, wyp, i0y, R, xw, opd);
  else
    dm_separable_kernel<<<grid, 256, smem, s>>>(coefs, ldc, act_pos, nA, nAct, gx, gy, (const int2*)band_x,
                                                (const int2*)band_y, R, xw, opd);
  AOENV_LAUNCH_CHECK("dm_separable");
  return 0;
}
