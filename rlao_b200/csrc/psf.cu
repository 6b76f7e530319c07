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
// 64 (y) x 32 (x) tiles through shared memory: coalesced float reads along x, bf16x2 stores along y (128 B per warp).
__global__ void __launch_bounds__(256)
psf_field_kernel(const float* __restrict__ opd_a, const float* __restrict__ opd_b, const float* __restrict__ pupil,
                 const float* __restrict__ amp, int R, float phase_turns, __nv_bfloat16* __restrict__ planes, int ldk,
                 size_t plane_stride) {
  __shared__ float sr[64][33], si[64][33];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const size_t img = (size_t)b * R * R;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
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
  const int yl = 2 * tx, y = y0 + yl;             // this thread's pair of rows (R is even: ldk >= 2R, both multiples of 2)
  if (y >= R) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xl = ty + 8 * j, x = x0 + xl;
    if (x < R) {
      const size_t row = ((size_t)b * R + x) * ldk;
      const float r0 = sr[yl][xl], r1 = sr[yl + 1][xl], i0 = si[yl][xl], i1 = si[yl + 1][xl];
      const __nv_bfloat162 hr = __floats2bfloat162_rn(r0, r1), hi = __floats2bfloat162_rn(i0, i1);
      const float2 fr = __bfloat1622float2(hr), fi = __bfloat1622float2(hi);
      *reinterpret_cast<__nv_bfloat162*>(planes + row + y) = hr;
      *reinterpret_cast<__nv_bfloat162*>(planes + row + R + y) = hi;
      *reinterpret_cast<__nv_bfloat162*>(planes + plane_stride + row + y) = __floats2bfloat162_rn(r0 - fr.x, r1 - fr.y);
      *reinterpret_cast<__nv_bfloat162*>(planes + plane_stride + row + R + y) = __floats2bfloat162_rn(i0 - fi.x, i1 - fi.y);
    }
  }
}

// ---- stage 2 ----------------------------------------------------------------------------------------------
constexpr int kRows = 8;     // un-binned output rows per block

// One block per (group of kRows output rows, environment).  Thread (v, slice) accumulates all kRows rows of output
// column v over its slice of the R input columns: one coalesced twiddle load feeds 4*kRows FMAs, the T values are
// shared-memory broadcasts.  The slices are then summed through shared memory.
__global__ void __launch_bounds__(256)
psf_window_kernel(const float* __restrict__ T, int ldt, const float2* __restrict__ g2, int R, int Wu, int os, int win,
                  int slices, float inv_n2, float* __restrict__ psf_win, int* __restrict__ psf_max_bits) {
  extern __shared__ float2 sT[];                 // [kRows][R], then reused as [slices][kRows][Wu] partial sums
  __shared__ float sI[kRows * 256];              // [kRows][Wu], Wu <= 256
  const int b = blockIdx.y, u0 = blockIdx.x * kRows;
  for (int i = threadIdx.x; i < R * kRows; i += blockDim.x) {
    const int x = i >> 3, j = i & 7;
    const float* __restrict__ row = T + ((size_t)b * R + x) * ldt + u0 + j;
    sT[j * R + x] = make_float2(__ldg(row), __ldg(row + Wu));
  }
  __syncthreads();
  const int v = threadIdx.x % Wu, sl = threadIdx.x / Wu;
  float fr[kRows], fi[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) { fr[r] = 0.f; fi[r] = 0.f; }
  if (sl < slices) {
    const int per = (R + slices - 1) / slices;
    const int x0 = sl * per, x1 = min(R, x0 + per);
#pragma unroll 2
    for (int x = x0; x < x1; ++x) {
      const float2 g = __ldg(&g2[(size_t)x * Wu + v]);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const float2 e = sT[r * R + x];
        fr[r] = fmaf(e.x, g.x, fmaf(-e.y, g.y, fr[r]));
        fi[r] = fmaf(e.x, g.y, fmaf(e.y, g.x, fi[r]));
      }
    }
  }
  __syncthreads();                               // everyone is done reading sT
  if (sl < slices) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) sT[(sl * kRows + r) * Wu + v] = make_float2(fr[r], fi[r]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kRows * Wu; o += blockDim.x) {
    float ar = 0.f, ai = 0.f;
    for (int q = 0; q < slices; ++q) {
      const float2 p = sT[q * kRows * Wu + o];
      ar += p.x;
      ai += p.y;
    }
    sI[o] = (ar * ar + ai * ai) * inv_n2;
  }
  __syncthreads();
  const int rows_b = kRows / os;
  for (int o = threadIdx.x; o < rows_b * win; o += blockDim.x) {
    const int rb = o / win, cb = o - rb * win;
    float val = 0.f;
    for (int i = 0; i < os; ++i)
      for (int j = 0; j < os; ++j) val += sI[(rb * os + i) * Wu + cb * os + j];
    const int yb = u0 / os + rb;
    if (psf_win) psf_win[((size_t)b * win + yb) * win + cb] = val;
    atomicMax(&psf_max_bits[b], __float_as_int(val));   // val >= 0: int order == float order
  }
}

// ---- full image: both transforms on the tensor cores -----------------------------------------------------------------------
// planes[(b, u)][c * R + x] = T[(b, x)][c * Wu + u]: the [x][u] -> [u][x] transposition between the row and the column
// transform, through a 32 x 32 shared-memory tile, written in split-bf16 operand form (2 parts)
__global__ void __launch_bounds__(256)
psf_transpose_kernel(const float* __restrict__ T, int ldt, int R, int Wu, __nv_bfloat16* __restrict__ planes, int ldk,
                     size_t plane_stride) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z >> 1, c = blockIdx.z & 1;
  const int x0 = blockIdx.y * 32, u0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int x = x0 + r, u = u0 + tx;
    tile[r][tx] = (x < R && u < Wu) ? __ldg(&T[((size_t)b * R + x) * ldt + c * Wu + u]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int u = u0 + r, x = x0 + tx;
    if (u < Wu && x < R) {
      const float v = tile[tx][r];
      const size_t o = ((size_t)b * Wu + u) * ldk + c * R + x;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      planes[o] = h;
      planes[plane_stride + o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// |F|^2 / N^2 binned os x os: F[(b, u)][c * Wu + v] -> psf[b][u / os][v / os], and the maximum of every image
__global__ void __launch_bounds__(256)
psf_intensity_kernel(const float* __restrict__ F, int ldf, int Wu, int os, int win, float inv_n2, float* __restrict__ psf,
                     int* __restrict__ psf_max_bits) {
  const int b = blockIdx.z, yb = blockIdx.y;
  const int xb = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (xb < win) {
    for (int i = 0; i < os; ++i) {
      const float* __restrict__ row = F + ((size_t)b * Wu + yb * os + i) * ldf;
      for (int j = 0; j < os; ++j) {
        const float re = __ldg(row + xb * os + j), im = __ldg(row + Wu + xb * os + j);
        val = fmaf(re, re, fmaf(im, im, val));
      }
    }
    val *= inv_n2;
    psf[((size_t)b * win + yb) * win + xb] = val;
  }
  val = warp_max(val);
  if ((threadIdx.x & 31) == 0) atomicMax(&psf_max_bits[b], __float_as_int(val));      // val >= 0: int order == float order
}

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_psf_image(const float* opd_a, const float* opd_b, const float* pupil, const float* amp,
                               const void* w_planes, int B, int R, int N, int os, int win, float phase_scale,
                               void* field_planes, int ldk, float* work_t, void* planes_u, float* work_f, float* psf,
                               float* psf_max, void* stream) {
  AOENV_CHECK_ARG(B > 0 && 2 * B <= 65535 && R > 0 && N >= R && (N - R) % 2 == 0, "psf_image: bad shape B=%d R=%d N=%d", B, R, N);
  AOENV_CHECK_ARG((os == 1 || os == 2) && N % os == 0, "psf_image: oversampling %d unsupported", os);
  const int Wu = os * win;
  AOENV_CHECK_ARG(win > 0 && Wu <= N && win <= 65535, "psf_image: image of %d pixels unsupported", win);
  AOENV_CHECK_ARG(ldk >= 2 * R && ldk % 8 == 0 && R % 2 == 0, "psf_image: ldk=%d must be a multiple of 8 and >= 2R, R even", ldk);
  AOENV_CHECK_ARG((long long)B * Wu < (1LL << 31) && (long long)B * R < (1LL << 31), "psf_image: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(psf_max, 0, sizeof(float) * (size_t)B, s);
  if (e != cudaSuccess) return fail(-3, "psf_image memset: %s", cudaGetErrorString(e));
  psf_field_kernel<<<dim3((R + 31) / 32, (R + 63) / 64, B), 256, 0, s>>>(opd_a, opd_b, pupil, amp, R, phase_scale * 0.15915494309189535f,
                                                                         (__nv_bfloat16*)field_planes, ldk, (size_t)B * R * ldk);
  AOENV_LAUNCH_CHECK("psf_field");
  // rows: T[(b, x)][(c', u)]
  int rc = aoenv_gemm_tn_tc(field_planes, w_planes, ldk, 2, work_t, 2 * Wu, B * R, 2 * Wu, 2 * R, 1.0f, stream);
  if (rc) return rc;
  psf_transpose_kernel<<<dim3((Wu + 31) / 32, (R + 31) / 32, 2 * B), 256, 0, s>>>(work_t, 2 * Wu, R, Wu, (__nv_bfloat16*)planes_u, ldk,
                                                                               (size_t)B * Wu * ldk);
  AOENV_LAUNCH_CHECK("psf_transpose");
  // columns: F[(b, u)][(c'', v)] — the same operator (the kernel is symmetric in the two axes)
  rc = aoenv_gemm_tn_tc(planes_u, w_planes, ldk, 2, work_f, 2 * Wu, B * Wu, 2 * Wu, 2 * R, 1.0f, stream);
  if (rc) return rc;
  psf_intensity_kernel<<<dim3((win + 255) / 256, win, B), 256, 0, s>>>(work_f, 2 * Wu, Wu, os, win, 1.0f / ((float)N * (float)N), psf,
                                                                      (int*)psf_max);
  AOENV_LAUNCH_CHECK("psf_intensity");
  return 0;
}

extern "C" int aoenv_psf_peak(const float* opd_a, const float* opd_b, const float* pupil, const float* amp,
                              const void* w1_planes, const float* g2, int B, int R, int N, int os, int win,
                              float phase_scale, void* field_planes, int ldk, float* scratch, float* psf_win,
                              float* psf_max, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && N >= R && (N - R) % 2 == 0, "psf_peak: bad shape B=%d R=%d N=%d", B, R, N);
  AOENV_CHECK_ARG((os == 1 || os == 2) && N % os == 0, "psf_peak: oversampling %d unsupported", os);
  const int Wu = os * win;
  AOENV_CHECK_ARG(win > 0 && Wu <= 256 && Wu <= N && Wu % kRows == 0, "psf_peak: window of %d binned pixels unsupported", win);
  AOENV_CHECK_ARG(ldk >= 2 * R && ldk % 8 == 0, "psf_peak: ldk=%d must be a multiple of 8 and >= 2R", ldk);
  AOENV_CHECK_ARG(R % 2 == 0, "psf_peak: odd pupil size R=%d", R);
  AOENV_CHECK_ARG((long long)B * R < (1LL << 31), "psf_peak: B*R too large");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(psf_max, 0, sizeof(float) * (size_t)B, s);
  if (e != cudaSuccess) return fail(-3, "psf_peak memset: %s", cudaGetErrorString(e));
  const int MX = B * R;
  dim3 g0((R + 31) / 32, (R + 63) / 64, B);
  psf_field_kernel<<<g0, 256, 0, s>>>(opd_a, opd_b, pupil, amp, R, phase_scale * 0.15915494309189535f,
                                      (__nv_bfloat16*)field_planes, ldk, (size_t)MX * ldk);
  AOENV_LAUNCH_CHECK("psf_field");
  int rc = aoenv_gemm_tn_tc(field_planes, w1_planes, ldk, 2, scratch, 2 * Wu, MX, 2 * Wu, 2 * R, 1.0f, stream);
  if (rc) return rc;
  dim3 g2d(Wu / kRows, B);
  const int slices = 256 / Wu > 0 ? 256 / Wu : 1;
  size_t smem = sizeof(float2) * (size_t)kRows * R;
  const size_t partial = sizeof(float2) * (size_t)slices * kRows * Wu;
  if (partial > smem) smem = partial;
  if (smem > 48 * 1024) {      // per device; cheap enough to repeat
    e = cudaFuncSetAttribute(psf_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "psf_window smem attribute: %s", cudaGetErrorString(e));
  }
  psf_window_kernel<<<g2d, 256, smem, s>>>(scratch, 2 * Wu, (const float2*)g2, R, Wu, os, win, slices,
                                           1.0f / ((float)N * (float)N), psf_win, (int*)psf_max);
  AOENV_LAUNCH_CHECK("psf_window");
  return 0;
}
