// Science-path PSF peak (OOPAO/Telescope.py:260-360: computePSF -> PropagateField; callers take PSF.max()).
// The reference zero-pads the R x R pupil field to N x N, multiplies by the half-pixel phasor exp(-i pi (x+y)/N),
// takes the centred FFT / N, |.|^2 and bins os x os.  Only the neighbourhood of the core is needed for a Strehl
// ratio, so this is a pruned DFT onto the Wu = os*win wanted output rows / columns:
//
//   stage 0  psf_field_kernel   E[b][y][x] = amp * exp(i phase) written TRANSPOSED and split in bf16 parts:
//                               X[p][(b, x)][c*R + y]  (c = 0 real, 1 imaginary)  — the operand layout of gemm_tc.cu
//   stage 1  tcgen05 GEMM       T[(b, x)][c'*Wu + u] = sum_{c,y} X[(b, x)][c*R + y] * W1[c'*Wu + u][c*R + y]
//                               with W1 = [[Wr, -Wi], [Wi, Wr]], w(u, y) = exp(-i pi (pad + y)(2 d_u + 1) / N),
//                               d_u = s0 + u - N/2: the R input rows -> Wu output rows as ONE real GEMM
//                               (M = B*R, N = 2*Wu, K = 2*R; FP32-grade through the split-bf16 scheme)
//   stage 2  psf_window_kernel  F[b][u][v] = sum_x T[b][u][x] * w(v, x); |F|^2 / N^2, os x os binning, window max
//
// (a factor (-1)^d is dropped from w: it does not change |F|^2).  The twiddle operands come from the host in
// float64-exact form: w1 as split-bf16 planes, g2[x][v] = w(v, x) as float2.
#include "common.cuh"

namespace aoenv {

// ---- stage 0 ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
psf_field_kernel(const float* __restrict__ opd_a, const float* __restrict__ opd_b, const float* __restrict__ pupil,
                 const float* __restrict__ amp, int R, float phase_turns, __nv_bfloat16* __restrict__ planes, int ldk,
                 size_t plane_stride) {
  __shared__ float sr[32][33], si[32][33];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const size_t img = (size_t)b * R * R;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = y0 + ty + 8 * j, x = x0 + tx;
    float er = 0.f, ei = 0.f;
    if (y < R && x < R) {
      const size_t o = (size_t)y * R + x;
      float t = __ldg(opd_a + img + o);
      if (opd_b) t += __ldg(opd_b + img + o);
      const float turns = t * __ldg(pupil + o) * phase_turns;
      const float ang = (turns - rintf(turns)) * 6.283185307179586f;
      const float am = __ldg(amp + o);
      er = am * __cosf(ang);
      ei = am * __sinf(ang);
    }
    sr[ty + 8 * j][tx] = er;
    si[ty + 8 * j][tx] = ei;
  }
  __syncthreads();
  const int y = y0 + tx;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + ty + 8 * j;
    if (y < R && x < R) {
      const size_t row = ((size_t)b * R + x) * ldk;
      float vr = sr[tx][ty + 8 * j], vi = si[tx][ty + 8 * j];
      const __nv_bfloat16 hr = __float2bfloat16_rn(vr), hi = __float2bfloat16_rn(vi);
      planes[row + y] = hr;
      planes[row + R + y] = hi;
      vr -= __bfloat162float(hr);
      vi -= __bfloat162float(hi);
      planes[plane_stride + row + y] = __float2bfloat16_rn(vr);
      planes[plane_stride + row + R + y] = __float2bfloat16_rn(vi);
    }
  }
}

// ---- stage 2 ----------------------------------------------------------------------------------------------
constexpr int kRows = 8;     // un-binned output rows per block

__global__ void __launch_bounds__(256)
psf_window_kernel(const float* __restrict__ T, int ldt, const float2* __restrict__ g2, int R, int Wu, int os, int win,
                  float inv_n2, float* __restrict__ psf_win, int* __restrict__ psf_max_bits) {
  extern __shared__ float2 sT[];                 // [kRows][R]
  __shared__ float sI[kRows * 256];              // [kRows][Wu], Wu <= 256
  const int b = blockIdx.y, u0 = blockIdx.x * kRows;
  for (int i = threadIdx.x; i < R * kRows; i += blockDim.x) {
    const int x = i >> 3, j = i & 7;
    const float* __restrict__ row = T + ((size_t)b * R + x) * ldt + u0 + j;
    sT[j * R + x] = make_float2(__ldg(row), __ldg(row + Wu));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kRows * Wu; o += blockDim.x) {
    const int r = o / Wu, v = o - r * Wu;
    const float2* __restrict__ t = sT + r * R;
    float fr = 0.f, fi = 0.f;
#pragma unroll 4
    for (int x = 0; x < R; ++x) {
      const float2 g = __ldg(&g2[(size_t)x * Wu + v]);
      const float2 e = t[x];
      fr = fmaf(e.x, g.x, fmaf(-e.y, g.y, fr));
      fi = fmaf(e.x, g.y, fmaf(e.y, g.x, fi));
    }
    sI[o] = (fr * fr + fi * fi) * inv_n2;
  }
  __syncthreads();
  const int rows_b = kRows / os;
  for (int o = threadIdx.x; o < rows_b * win; o += blockDim.x) {
    const int rb = o / win, cb = o - rb * win;
    float v = 0.f;
    for (int i = 0; i < os; ++i)
      for (int j = 0; j < os; ++j) v += sI[(rb * os + i) * Wu + cb * os + j];
    const int yb = u0 / os + rb;
    if (psf_win) psf_win[((size_t)b * win + yb) * win + cb] = v;
    atomicMax(&psf_max_bits[b], __float_as_int(v));   // v >= 0: int order == float order
  }
}

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_psf_peak(const float* opd_a, const float* opd_b, const float* pupil, const float* amp,
                              const void* w1_planes, const float* g2, int B, int R, int N, int os, int win,
                              float phase_scale, void* field_planes, int ldk, float* scratch, float* psf_win,
                              float* psf_max, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && N >= R && (N - R) % 2 == 0, "psf_peak: bad shape B=%d R=%d N=%d", B, R, N);
  AOENV_CHECK_ARG((os == 1 || os == 2) && N % os == 0, "psf_peak: oversampling %d unsupported", os);
  const int Wu = os * win;
  AOENV_CHECK_ARG(win > 0 && Wu <= 256 && Wu <= N && Wu % kRows == 0, "psf_peak: window of %d binned pixels unsupported", win);
  AOENV_CHECK_ARG(ldk >= 2 * R && ldk % 8 == 0, "psf_peak: ldk=%d must be a multiple of 8 and >= 2R", ldk);
  AOENV_CHECK_ARG((long long)B * R < (1LL << 31), "psf_peak: B*R too large");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(psf_max, 0, sizeof(float) * (size_t)B, s);
  if (e != cudaSuccess) return fail(-3, "psf_peak memset: %s", cudaGetErrorString(e));
  const int MX = B * R;
  dim3 g0((R + 31) / 32, (R + 31) / 32, B);
  psf_field_kernel<<<g0, 256, 0, s>>>(opd_a, opd_b, pupil, amp, R, phase_scale * 0.15915494309189535f,
                                      (__nv_bfloat16*)field_planes, ldk, (size_t)MX * ldk);
  AOENV_LAUNCH_CHECK("psf_field");
  int rc = aoenv_gemm_tn_tc(field_planes, w1_planes, ldk, 2, scratch, 2 * Wu, MX, 2 * Wu, 2 * R, 1.0f, stream);
  if (rc) return rc;
  dim3 g2d(Wu / kRows, B);
  const size_t smem = sizeof(float2) * (size_t)kRows * R;
  if (smem > 48 * 1024) {      // per device; cheap enough to repeat
    e = cudaFuncSetAttribute(psf_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "psf_window smem attribute: %s", cudaGetErrorString(e));
  }
  psf_window_kernel<<<g2d, 256, smem, s>>>(scratch, 2 * Wu, (const float2*)g2, R, Wu, os, win,
                                           1.0f / ((float)N * (float)N), psf_win, (int*)psf_max);
  AOENV_LAUNCH_CHECK("psf_window");
  return 0;
}
