// Command update, observation assembly and scalar diagnostics of the env step
// (MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:479,484,492-514,604-605,621-641).
#include "common.cuh"

namespace aoenv {

__global__ void __launch_bounds__(256)
command_update_kernel(const float* __restrict__ action, const int32_t* __restrict__ act_idx, int nA, int nAct2,
                      float leak, float* __restrict__ coefs, float* __restrict__ dm_prev, int ldc) {
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

}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_command_update(const float* action, const int32_t* act_idx, int B, int nA, int nAct2, float leak,
                         float* coefs, float* dm_prev, int ldc, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && nA > 0 && ldc >= nA, "command_update: bad shape");
  dim3 grid((nA + 255) / 256, B);
  command_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(action, act_idx, nA, nAct2, leak, coefs, dm_prev, ldc);
  AOENV_LAUNCH_CHECK("command_update");
  return 0;
}

int aoenv_observe(const float* rec, int ldr, const int32_t* act_idx, int B, int nA, int nAct2, const double* stats,
                  double n_pupil, float phase_scale, float* obs, float* reward, float* strehl, float* total,
                  float* residual, void* stream) {
  AOENV_CHECK_ARG(B > 0 && nA > 0 && ldr >= nA, "observe: bad shape");
  observe_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(rec, ldr, act_idx, nA, nAct2, stats, n_pupil, phase_scale, obs,
                                                      reward, strehl, total, residual);
  AOENV_LAUNCH_CHECK("observe");
  return 0;
}

}  // extern "C"
