// Science-path PSF peak (OOPAO/Telescope.py:260-360: computePSF -> PropagateField; callers take PSF.max()).
// The reference zero-pads the R x R pupil field to N x N, multiplies by the half-pixel phasor exp(-i pi (x+y)/N),
// takes the centred FFT / N, |.|^2 and bins os x os.  Only the neighbourhood of the core is needed for a Strehl
// ratio, so this is a pruned DFT: stage 1 transforms the R input rows to the Wu = os*win wanted output rows,
// stage 2 the R input columns to the Wu wanted output columns; twiddles come from one exact table
// tw[m] = exp(-i pi m / N), m in [0, 2N).  Row/column kernel: exp(-i pi (pad + y)(2 d + 1) / N), d = s - N/2
// (a factor (-1)^d is dropped: it does not change |F|^2).
#include "common.cuh"

namespace aoenv {

constexpr int kS = 16;   // output rows per thread in stage 1

__global__ void __launch_bounds__(128)
psf_stage1_kernel(const float* __restrict__ opd_a, const float* __restrict__ opd_b, const float* __restrict__ pupil,
                  const float* __restrict__ amp, const float2* __restrict__ tw, int R, int N, int pad, int s0, int Wu,
                  float phase_scale, float2* __restrict__ T) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int sb = blockIdx.y * kS;
  const bool ok = x < R;
  float ar[kS], ai[kS];
  long mult[kS];
#pragma unroll
  for (int k = 0; k < kS; ++k) {
    ar[k] = 0.f; ai[k] = 0.f;
    const long d = (long)(s0 + sb + k) - N / 2;
    mult[k] = (((2 * d + 1) % (2L * N)) + 2L * N) % (2L * N);
  }
  const size_t img = (size_t)b * R * R;
  for (int y = 0; y < R; ++y) {
    float er = 0.f, ei = 0.f;
    if (ok) {
      const size_t o = (size_t)y * R + x;
      float t = __ldg(opd_a + img + o);
      if (opd_b) t += __ldg(opd_b + img + o);
      float sn, cs;
      sincosf(t * __ldg(pupil + o) * phase_scale, &sn, &cs);
      const float am = __ldg(amp + o);
      er = am * cs; ei = am * sn;
    }
    const long py = pad + y;
#pragma unroll
    for (int k = 0; k < kS; ++k) {
      const float2 g = __ldg(&tw[(py * mult[k]) % (2L * N)]);
      ar[k] = fmaf(er, g.x, fmaf(-ei, g.y, ar[k]));
      ai[k] = fmaf(er, g.y, fmaf(ei, g.x, ai[k]));
    }
  }
  if (ok) {
#pragma unroll
    for (int k = 0; k < kS; ++k)
      if (sb + k < Wu) T[((size_t)b * Wu + sb + k) * R + x] = make_float2(ar[k], ai[k]);
  }
}

// one block per (binned output row, environment); thread t = un-binned output column
__global__ void __launch_bounds__(256)
psf_stage2_kernel(const float2* __restrict__ T, const float2* __restrict__ tw, int R, int N, int pad, int s0, int Wu,
                  int os, int win, float* __restrict__ psf_win, int* __restrict__ psf_max_bits) {
  extern __shared__ float2 sT[];            // [os][R]
  __shared__ float sI[256];
  const int b = blockIdx.y, yb = blockIdx.x;
  for (int i = threadIdx.x; i < os * R; i += blockDim.x)
    sT[i] = T[((size_t)b * Wu + (size_t)yb * os) * R + i];
  __syncthreads();
  const int t = threadIdx.x;
  float inten = 0.f;
  if (t < Wu) {
    const long d = (long)(s0 + t) - N / 2;
    const long mult = (((2 * d + 1) % (2L * N)) + 2L * N) % (2L * N);
    for (int r = 0; r < os; ++r) {
      float fr = 0.f, fi = 0.f;
      for (int x = 0; x < R; ++x) {
        const float2 g = __ldg(&tw[((long)(pad + x) * mult) % (2L * N)]);
        const float2 v = sT[r * R + x];
        fr = fmaf(v.x, g.x, fmaf(-v.y, g.y, fr));
        fi = fmaf(v.x, g.y, fmaf(v.y, g.x, fi));
      }
      inten += fr * fr + fi * fi;
    }
    inten /= (float)N * (float)N;
  }
  sI[t] = inten;
  __syncthreads();
  if (t < win) {
    float v = 0.f;
    for (int r = 0; r < os; ++r) v += sI[t * os + r];
    if (psf_win) psf_win[((size_t)b * win + yb) * win + t] = v;
    atomicMax(&psf_max_bits[b], __float_as_int(v));   // v >= 0: int order == float order
  }
}

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_psf_peak(const float* opd_a, const float* opd_b, const float* pupil, const float* amp,
                              const float* tw, int B, int R, int N, int os, int win, float phase_scale, float* scratch,
                              float* psf_win, float* psf_max, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && N >= R && (N - R) % 2 == 0, "psf_peak: bad shape B=%d R=%d N=%d", B, R, N);
  AOENV_CHECK_ARG((os == 1 || os == 2) && N % os == 0, "psf_peak: oversampling %d unsupported", os);
  const int Wu = os * win;
  AOENV_CHECK_ARG(win > 0 && Wu <= 256 && Wu <= N && Wu % kS == 0, "psf_peak: window of %d binned pixels unsupported", win);
  cudaStream_t s = (cudaStream_t)stream;
  const int pad = (N - R) / 2;
  const int s0 = os * ((N / os) / 2 - win / 2);
  cudaError_t e = cudaMemsetAsync(psf_max, 0, sizeof(float) * (size_t)B, s);
  if (e != cudaSuccess) return fail(-3, "psf_peak memset: %s", cudaGetErrorString(e));
  dim3 g1((R + 127) / 128, Wu / kS, B);
  psf_stage1_kernel<<<g1, 128, 0, s>>>(opd_a, opd_b, pupil, amp, (const float2*)tw, R, N, pad, s0, Wu, phase_scale,
                                       (float2*)scratch);
  AOENV_LAUNCH_CHECK("psf_stage1");
  dim3 g2(win, B);
  psf_stage2_kernel<<<g2, 256, sizeof(float2) * os * R, s>>>((const float2*)scratch, (const float2*)tw, R, N, pad, s0, Wu,
                                                            os, win, psf_win, (int*)psf_max);
  AOENV_LAUNCH_CHECK("psf_stage2");
  return 0;
}
