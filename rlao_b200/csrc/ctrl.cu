// Command update, observation assembly and scalar diagnostics of the env step
// (MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:479,484,492-514,604-605,621-641).
#include "common.cuh"

namespace aoenv {

__global__ void __launch_bounds__(256)
command_update_kernel(const float* __restrict__ action, const int32_t* __restrict__ act_idx, int nA, int nAct2,
                      float leak, float* __restrict__ coefs, float* __restrict__ dm_prev, int ldc) {
  pdl_enter();
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nA) return;
  const float act = __ldg(&action[(size_t)b * nAct2 + __ldg(&act_idx[a])]) * 1e-6f;
  const float c = dm_prev[(size_t)b * ldc + a] * leak + act;
  coefs[(size_t)b * ldc + a] = c;
  dm_prev[(size_t)b * ldc + a] = c;
}

// one block per environment
__global__ void __launch_bounds__(256)
observe_kernel(const float* __restrict__ rec, int ldr, const int32_t* __restrict__ act_idx, int nA, int nAct2,
               const double* __restrict__ stats, double n_pupil, float phase_scale, float* __restrict__ obs,
               float* __restrict__ reward, float* __restrict__ strehl, float* __restrict__ total,
               float* __restrict__ residual) {
  pdl_enter();
  const int b = blockIdx.x;
  float* __restrict__ o = obs + (size_t)b * nAct2;
  for (int i = threadIdx.x; i < nAct2; i += blockDim.x) o[i] = 0.f;
  __syncthreads();
  float ss = 0.f;
  for (int a = threadIdx.x; a < nA; a += blockDim.x) {
    const float v = -__ldg(&rec[(size_t)b * ldr + a]) * 1e6f;
    o[__ldg(&act_idx[a])] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    reward[b] = -sqrtf(t);
    if (stats != nullptr) {
      const double* s = stats + (size_t)b * 4;
      const double ma = s[0] / n_pupil, va = fmax(s[1] / n_pupil - ma * ma, 0.0);
      const double mt = s[2] / n_pupil, vt = fmax(s[3] / n_pupil - mt * mt, 0.0);
      total[b] = (float)(sqrt(va) * 1e9);
      residual[b] = (float)(sqrt(vt) * 1e9);
      strehl[b] = (float)exp(-vt * (double)phase_scale * (double)phase_scale);
    }
  }
}

// Exploration noise, first half (MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:616-619: np.random.normal(0, sigma, nValidAct)):
// z[r][c] = sigma * N(0, 1) from Philox4x32-10 keyed by `seed`, counter (c / 4, r, call counter): four normals per block
// (two Box-Muller pairs).  Written as float32 and, optionally, as split-bf16 operand planes for aoenv_gemm_tn_tc.
__global__ void __launch_bounds__(256)
normal_fill_kernel(unsigned long long seed, unsigned long long counter, int rows, int cols, int ld, float sigma,
                   float* __restrict__ out, __nv_bfloat16* __restrict__ planes, int parts) {
  const int r = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;            // group of four columns
  if (4 * q >= ld) return;
  Philox rng(seed);
  const uint4 w = rng((uint32_t)q, (uint32_t)r, (uint32_t)counter, (uint32_t)(counter >> 32));
  const float2 n0 = box_muller(w.x, w.y), n1 = box_muller(w.z, w.w);
  const float v[4] = {n0.x, n0.y, n1.x, n1.y};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = 4 * q + k;
    if (c >= ld) break;
    const float x = c < cols ? v[k] * sigma : 0.f;                // padding columns stay zero
    out[(size_t)r * ld + c] = x;
    if (planes != nullptr) store_bf16_planes(planes, (size_t)rows * ld, (size_t)r * ld + c, parts, x);
  }
}

// vec_to_img (OOPAOEnvRazor.py:621-630): img[b][act_idx[a]] = vec[b][a] * scale, zero elsewhere
__global__ void __launch_bounds__(256)
vec_to_img_kernel(const float* __restrict__ vec, int ldv, const int32_t* __restrict__ act_idx, int nA, int nAct2, float scale,
                  float* __restrict__ img) {
  const int b = blockIdx.x;
  float* __restrict__ o = img + (size_t)b * nAct2;
  for (int i = threadIdx.x; i < nAct2; i += blockDim.x) o[i] = 0.f;
  __syncthreads();
  for (int a = threadIdx.x; a < nA; a += blockDim.x) o[__ldg(&act_idx[a])] = __ldg(&vec[(size_t)b * ldv + a]) * scale;
}

}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_command_update(const float* action, const int32_t* act_idx, int B, int nA, int nAct2, float leak,
                         float* coefs, float* dm_prev, int ldc, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && nA > 0 && ldc >= nA, "command_update: bad shape");
  dim3 grid((nA + 255) / 256, B);
  AOENV_LAUNCH(command_update_kernel, grid, 256, 0, (cudaStream_t)stream, action, act_idx, nA, nAct2, leak, coefs, dm_prev, ldc);
  AOENV_LAUNCH_CHECK("command_update");
  return 0;
}

int aoenv_normal_fill(uint64_t seed, uint64_t counter, int rows, int cols, int ld, float sigma, float* out, void* planes,
                      int parts, void* stream) {
  AOENV_CHECK_ARG(rows > 0 && rows <= 65535 && cols > 0 && ld >= cols, "normal_fill: bad shape rows=%d cols=%d ld=%d", rows, cols, ld);
  AOENV_CHECK_ARG(planes == nullptr || parts == 2 || parts == 3, "normal_fill: parts must be 2 or 3");
  dim3 grid(((ld + 3) / 4 + 255) / 256, rows);
  normal_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, counter, rows, cols, ld, sigma, out, (__nv_bfloat16*)planes, parts);
  AOENV_LAUNCH_CHECK("normal_fill");
  return 0;
}

int aoenv_vec_to_img(const float* vec, int ldv, const int32_t* act_idx, int B, int nA, int nAct2, float scale, float* img,
                     void* stream) {
  AOENV_CHECK_ARG(B > 0 && nA > 0 && ldv >= nA && nAct2 >= nA, "vec_to_img: bad shape");
  vec_to_img_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(vec, ldv, act_idx, nA, nAct2, scale, img);
  AOENV_LAUNCH_CHECK("vec_to_img");
  return 0;
}

int aoenv_observe(const float* rec, int ldr, const int32_t* act_idx, int B, int nA, int nAct2, const double* stats,
                  double n_pupil, float phase_scale, float* obs, float* reward, float* strehl, float* total,
                  float* residual, void* stream) {
  AOENV_CHECK_ARG(B > 0 && nA > 0 && ldr >= nA, "observe: bad shape");
  AOENV_LAUNCH(observe_kernel, dim3(B), 256, 0, (cudaStream_t)stream, rec, ldr, act_idx, nA, nAct2, stats, n_pupil, phase_scale, obs,
                                                      reward, strehl, total, residual);
  AOENV_LAUNCH_CHECK("observe");
  return 0;
}

}  // extern "C"
