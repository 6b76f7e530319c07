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
                    const int2* __restrict__ band_y, int R, int x0, int xw, float* __restrict__ opd) {
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

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_dm_surface_separable(const float* coefs, int ldc, const int32_t* act_pos, int nA, int nAct,
                                          const float* gx, const float* gy, const int32_t* band_x, const int32_t* band_y,
                                          int B, int R, float* opd, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && nAct > 0 && nA > 0 && nA <= nAct * nAct && ldc >= nA, "dm_surface_separable: bad shape");
  // split the columns across CTAs so that C and T fit in shared memory (and small batches still fill the GPU)
  int parts = 1;
  auto smem_for = [&](int p) {
    const int xw = (((R + p - 1) / p) + 3) / 4 * 4;
    return (size_t)(((nAct * nAct + 3) & ~3) + nAct * xw) * sizeof(float);
  };
  while (smem_for(parts) > 96 * 1024 || (B * parts < 2 * kNumSMs && parts < 8)) ++parts;
  const int xw = (((R + parts - 1) / parts) + 3) / 4 * 4;
  const size_t smem = smem_for(parts);
  AOENV_CHECK_ARG(smem <= 200 * 1024, "dm_surface_separable: %d actuators across do not fit in shared memory", nAct);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(dm_separable_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "dm_surface_separable smem attribute: %s", cudaGetErrorString(e));
    attr = smem;
  }
  for (int p = 0; p < parts; ++p) {
    dm_separable_kernel<<<dim3(1, B), 256, smem, (cudaStream_t)stream>>>(coefs, ldc, act_pos, nA, nAct, gx, gy,
                                                                         (const int2*)band_x, (const int2*)band_y, R, p * xw,
                                                                         xw, opd);
    AOENV_LAUNCH_CHECK("dm_separable");
  }
  return 0;
}
