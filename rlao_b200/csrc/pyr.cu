// Pyramid wavefront sensor (OOPAO/Pyramid.py:469-504 pyramid_transform, :581-603 modulation loop, :987-1002 detector binning):
// per environment and modulation point the padded pupil field is Fourier transformed, multiplied by the pyramid's phase mask
// and transformed back; the intensities of all modulation points are summed and binned to the detector.
//
// The N x N transforms (N = 288 for the 20 x 20 system: 2^5 3^2; 128 for the 12 x 12 test system) are hand-written here:
// N = N1 * N2 Cooley-Tukey with in-register codelets (fft_codelets.cuh), THIRTY-TWO independent vectors per CTA — one per
// lane — in a shared-memory tile with an odd row stride (conflict-free), warp w = sub-transform w.  A forward transform
// leaves its output in digit-scrambled order (position j1 N2 + j2 holds frequency j1 + N1 j2), the inverse consumes that
// order and returns natural order, so no reordering pass exists; the mask is stored scrambled by the host.
//   K1 pyr_cols     per (env, theta, 32 pupil columns): field amp exp(i (phase + modulation tilt)) x half-pixel phasor
//                   (exact integer phase reduction), zero padded along y, forward transform over y -> X1[u'][x]
//                   (only the R non-zero columns are ever touched: the pruned half of the 2-D transform)
//   K2 pyr_rows     per (env, theta, 32 rows u'): zero padded along x, forward transform over x, x mask, inverse transform
//                   over the same axis (the middle stages fuse in registers), stored transposed -> Yt[q][u']
//   K3 pyr_image    per (env, 32 columns q): for every theta the inverse transform over u', |.|^2 accumulated in registers
//                   over theta -> intensity[p][q]
//   K4 pyr_bin      b x b binning to the detector pixels
// HBM traffic per environment and frame: nTheta x (X1 + Yt) written and read once (37 MB at N = 288, 20 points) — the
// previous cuFFT path moved 53 MB through ~8 passes per modulation point.
#include "common.cuh"
#include "fft_codelets.cuh"

namespace aoenv {
namespace pyr {

using fftc::cpx;
using fftc::mk;

// twiddles exp(-2 pi i n2 j1 / N), [n2][j1], one table per compiled factorisation (uploaded once per device)
__constant__ float2 c_tw_16x18[18 * 16];
__constant__ float2 c_tw_16x8[8 * 16];
template <int N1, int N2> __device__ __forceinline__ float2 twiddle(int i);
template <> __device__ __forceinline__ float2 twiddle<16, 18>(int i) { return c_tw_16x18[i]; }
template <> __device__ __forceinline__ float2 twiddle<16, 8>(int i) { return c_tw_16x8[i]; }

template <int N1, int N2>
struct Tile {
  static constexpr int N = N1 * N2;
  static constexpr int LD = N + 1;               // float2 per row: odd -> lanes (rows) fall in different banks
  static constexpr int kWarps = N1 > N2 ? N1 : N2;
  static constexpr int kThreads = kWarps * 32;

  // forward, stage 1: task n2 = warp, vector = lane
  static __device__ __forceinline__ void fwd1(float2* __restrict__ row, int n2) {
    cpx v[N1];
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) { const float2 t = row[n2 + N2 * k1]; v[k1] = mk(t.x, t.y); }
    fftc::Dft<N1>::run(v);
#pragma unroll
    for (int j1 = 0; j1 < N1; ++j1) {
      const float2 w = twiddle<N1, N2>(n2 * N1 + j1);
      const cpx t = fftc::cmul(v[j1], w.x, w.y);
      row[n2 + N2 * j1] = make_float2(t.x, t.y);
    }
  }
  // forward, stage 2 into registers: v[j2] = X[j1 + N1 j2]
  static __device__ __forceinline__ void fwd2_load(const float2* __restrict__ row, int j1, cpx (&v)[N2]) {
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) { const float2 t = row[N2 * j1 + n2]; v[n2] = mk(t.x, t.y); }
    fftc::Dft<N2>::run(v);
  }
  // inverse, first stage from registers (v[j2] = X[j1 + N1 j2]) -> row[N2 j1 + n2], conjugate twiddle applied
  static __device__ __forceinline__ void inv2_store(float2* __restrict__ row, int j1, cpx (&v)[N2]) {
#pragma unroll
    for (int j2 = 0; j2 < N2; ++j2) v[j2] = fftc::conj(v[j2]);
    fftc::Dft<N2>::run(v);                        // conj(DFT(conj x)) = N2 * inverse DFT
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) {
      const float2 w = twiddle<N1, N2>(n2 * N1 + j1);
      const cpx t = fftc::cmul(fftc::conj(v[n2]), w.x, -w.y);
      row[N2 * j1 + n2] = make_float2(t.x, t.y);
    }
  }
  static __device__ __forceinline__ void inv2(float2* __restrict__ row, int j1) {
    cpx v[N2];
#pragma unroll
    for (int j2 = 0; j2 < N2; ++j2) { const float2 t = row[N2 * j1 + j2]; v[j2] = mk(t.x, t.y); }
    inv2_store(row, j1, v);
  }
  // inverse, last stage into registers: v[k1] = x[N2 k1 + n2] (natural order), unnormalised (N x the inverse DFT)
  static __device__ __forceinline__ void inv1_load(const float2* __restrict__ row, int n2, cpx (&v)[N1]) {
#pragma unroll
    for (int j1 = 0; j1 < N1; ++j1) { const float2 t = row[N2 * j1 + n2]; v[j1] = mk(t.x, -t.y); }
    fftc::Dft<N1>::run(v);
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) v[k1] = fftc::conj(v[k1]);
  }
};

// ---- K1 ---------------------------------------------------------------------------------------------------
struct ColsArgs {
  const float* opd_a;          // [B][R][R]
  const float* opd_b;          // nullable
  const float* pupil;          // [R][R]
  const float* amp;            // [R][R] sqrt(flux / nTheta) * reflectivity
  const float* lin;            // [R] linspace(-pi, pi, R)
  const float2* mod;           // [nTheta] (px, py) modulation path, lambda/D
  float2* X1;                  // [B][nTheta][N][R]
  float phase_scale;
  int R, nTheta, lo;
};

template <int N1, int N2>
__global__ void __launch_bounds__(Tile<N1, N2>::kThreads, 2)
pyr_cols_kernel(const __grid_constant__ ColsArgs p) {
  using T = Tile<N1, N2>;
  constexpr int N = T::N;
  extern __shared__ __align__(16) float2 tile[];               // [32][LD]
  const int R = p.R, lo = p.lo;
  const int x0 = blockIdx.x * 32, th = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = x0 + lane;
  const bool col_ok = x < R;
  float2* __restrict__ row = tile + (size_t)lane * T::LD;
  // zero padding, then the field of this column: one y per warp iteration, lanes = columns (coalesced)
  for (int i = warp; i < N; i += T::kWarps) row[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const float2 m = __ldg(&p.mod[th]);
  const float phase_turns = p.phase_scale * 0.15915494309189535f;
  const float linx = col_ok ? __ldg(&p.lin[x]) : 0.f;
  for (int y = warp; y < R; y += T::kWarps) {
    if (col_ok) {
      const size_t o = (size_t)y * R + x;
      float t = __ldg(p.opd_a + (size_t)b * R * R + o);
      if (p.opd_b) t += __ldg(p.opd_b + (size_t)b * R * R + o);
      const float pu = __ldg(p.pupil + o);
      // modulation tilt in float32 like the reference (Pyramid.py:960-961), in radians -> turns
      const float pm = (m.x * linx + m.y * __ldg(&p.lin[y])) * pu;
      // half-pixel phasor exp(-i pi (N+1)(Y+X)/N), exact: integer reduction modulo 2N
      const int red = ((N + 1) * (2 * lo + y + x)) % (2 * N);
      const float turns = fmaf(t * pu, phase_turns, pm * 0.15915494309189535f) - (float)red / (float)(2 * N);
      const float ang = (turns - rintf(turns)) * 6.283185307179586f;
      const float a = __ldg(p.amp + o);
      row[lo + y] = make_float2(a * __cosf(ang), a * __sinf(ang));
    }
  }
  __syncthreads();
  if (warp < N2) T::fwd1(row, warp);
  __syncthreads();
  if (warp < N1) {
    cpx v[N2];
    T::fwd2_load(row, warp, v);
#pragma unroll
    for (int j2 = 0; j2 < N2; ++j2) row[N2 * warp + j2] = make_float2(v[j2].x, v[j2].y);
  }
  __syncthreads();
  // X1[b][th][pos][x]: lanes = columns, contiguous
  if (col_ok) {
    float2* __restrict__ out = p.X1 + (((size_t)b * p.nTheta + th) * N) * R + x;
    for (int pos = warp; pos < N; pos += T::kWarps) out[(size_t)pos * R] = row[pos];
  }
}

// ---- K2 ---------------------------------------------------------------------------------------------------
template <int N1, int N2>
__global__ void __launch_bounds__(Tile<N1, N2>::kThreads, 2)
pyr_rows_kernel(const float2* __restrict__ X1, const float2* __restrict__ mask_s, int R, int lo, int nTheta,
                float2* __restrict__ Yt) {
  using T = Tile<N1, N2>;
  constexpr int N = T::N;
  extern __shared__ __align__(16) float2 tile[];
  const int u0 = blockIdx.x * 32, th = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t img = (size_t)b * nTheta + th;
  // load 32 rows of X1 (R complex each, contiguous) into columns [lo, lo + R) of the tile, zero elsewhere
  for (int r = warp; r < 32; r += T::kWarps) {
    float2* __restrict__ dst = tile + (size_t)r * T::LD;
    const float2* __restrict__ src = X1 + (img * N + u0 + r) * R;
    if (((lo | R) & 1) == 0) {                      // two complex numbers per 128-bit load (lo, R even; rows 16-byte aligned)
      // all loads of the row first, then the stores: one load in flight per warp made this loop the top stall of the
      // kernel (ncu: 15 % of the samples on the first shared store, waiting for its load)
      constexpr int kIt = (N + 63) / 64;
      float4 v[kIt];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int xx = 2 * lane + 64 * it - lo;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (2 * lane + 64 * it < N && xx >= 0 && xx < R) v[it] = __ldg(reinterpret_cast<const float4*>(src + xx));
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = 2 * lane + 64 * it;
        if (i < N) {
          dst[i] = make_float2(v[it].x, v[it].y);
          dst[i + 1] = make_float2(v[it].z, v[it].w);
        }
      }
    } else {
      for (int i = lane; i < N; i += 32) {
        const int xx = i - lo;
        dst[i] = (xx >= 0 && xx < R) ? __ldg(src + xx) : make_float2(0.f, 0.f);
      }
    }
  }
  __syncthreads();
  float2* __restrict__ row = tile + (size_t)lane * T::LD;
  if (warp < N2) T::fwd1(row, warp);
  __syncthreads();
  if (warp < N1) {
    // forward stage 2 -> mask (stored in the same scrambled order) -> inverse stage 2, all in registers
    // the mask values of this lane's row segment (N2 consecutive complex numbers, 16-byte aligned: N2 is even) are requested
    // before the transform that precedes their use — issued right before the multiplication they were a quarter of the
    // kernel's samples (L2 latency, one consumer after the other)
    const float4* __restrict__ mk4 = reinterpret_cast<const float4*>(mask_s + (size_t)(u0 + lane) * N + N2 * warp);
    float4 m4[N2 / 2];
#pragma unroll
    for (int j = 0; j < N2 / 2; ++j) m4[j] = __ldg(mk4 + j);
    cpx v[N2];
    T::fwd2_load(row, warp, v);
#pragma unroll
    for (int j2 = 0; j2 < N2; ++j2) {
      const float wx = (j2 & 1) ? m4[j2 >> 1].z : m4[j2 >> 1].x, wy = (j2 & 1) ? m4[j2 >> 1].w : m4[j2 >> 1].y;
      v[j2] = fftc::cmul(v[j2], wx, wy);
    }
    T::inv2_store(row, warp, v);
  }
  __syncthreads();
  if (warp < N2) {
    cpx v[N1];
    T::inv1_load(row, warp, v);                   // reads and rewrites this warp's own positions {N2 k + warp}: in place
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) row[N2 * k1 + warp] = make_float2(v[k1].x, v[k1].y);
  }
  __syncthreads();
  // transposed store: Yt[b][th][q][u0 + lane], lanes = consecutive u'
  float2* __restrict__ out = Yt + (img * N) * N + u0 + lane;
#pragma unroll 4
  for (int q = warp; q < N; q += T::kWarps) out[(size_t)q * N] = row[q];
}

// ---- K3 ---------------------------------------------------------------------------------------------------
template <int N1, int N2>
__global__ void __launch_bounds__(Tile<N1, N2>::kThreads, 2)
pyr_image_kernel(const float2* __restrict__ Yt, int nTheta, float scale, float* __restrict__ intensity) {
  using T = Tile<N1, N2>;
  constexpr int N = T::N;
  extern __shared__ __align__(16) float2 tile[];
  const int q0 = blockIdx.x * 32, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[N1];
#pragma unroll
  for (int k1 = 0; k1 < N1; ++k1) acc[k1] = 0.f;
  float2* __restrict__ row = tile + (size_t)lane * T::LD;
  for (int th = 0; th < nTheta; ++th) {
    const size_t img = (size_t)b * nTheta + th;
    // all loads of a row first, then its stores (N is even and the rows are 16-byte aligned): with one load in flight per
    // warp this loop held 47 % of the kernel's samples
    constexpr int kIt = (N + 63) / 64;
    for (int r = warp; r < 32; r += T::kWarps) {
      float2* __restrict__ dst = tile + (size_t)r * T::LD;
      const float2* __restrict__ src = Yt + (img * N + q0 + r) * N;
      float4 v[kIt];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = 2 * lane + 64 * it;
        if (i < N) v[it] = __ldg(reinterpret_cast<const float4*>(src + i));
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = 2 * lane + 64 * it;
        if (i < N) {
          dst[i] = make_float2(v[it].x, v[it].y);
          dst[i + 1] = make_float2(v[it].z, v[it].w);
        }
      }
    }
    __syncthreads();
    if (warp < N1) T::inv2(row, warp);
    __syncthreads();
    if (warp < N2) {
      cpx v[N1];
      T::inv1_load(row, warp, v);
#pragma unroll
      for (int k1 = 0; k1 < N1; ++k1) acc[k1] = fmaf(v[k1].x, v[k1].x, fmaf(v[k1].y, v[k1].y, acc[k1]));
    }
    __syncthreads();
  }
  if (warp < N2) {
    float* __restrict__ out = intensity + (size_t)b * N * N + q0 + lane;
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) out[(size_t)(N2 * k1 + warp) * N] = acc[k1] * scale;
  }
}

// ---- K4 ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pyr_bin_kernel(const float* __restrict__ intensity, int N, int bin, float* __restrict__ frame) {
  const int nc = N / bin, b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nc * nc) return;
  const int yc = o / nc, xc = o - yc * nc;
  const float* __restrict__ src = intensity + (size_t)b * N * N + (size_t)(yc * bin) * N + xc * bin;
  float s = 0.f;
  for (int i = 0; i < bin; ++i)
    for (int j = 0; j < bin; ++j) s += __ldg(src + (size_t)i * N + j);
  frame[(size_t)b * nc * nc + o] = s;
}

static int g_tw_done[64][2] = {{0}};
template <int N1, int N2>
static int ensure_tables() {
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int N = N1 * N2, which = N2 == 18 ? 0 : 1;
  if (dev < 64 && g_tw_done[dev][which]) return 0;
  float2 h[N1 * N2];
  for (int n2 = 0; n2 < N2; ++n2)
    for (int j1 = 0; j1 < N1; ++j1) {
      const double a = -2.0 * M_PI * (double)((n2 * j1) % N) / (double)N;
      h[n2 * N1 + j1] = make_float2((float)cos(a), (float)sin(a));
    }
  // synchronous (`h` is a stack buffer); each factorisation has its own table, so an upload never races a running kernel
  cudaError_t e = which == 0 ? cudaMemcpyToSymbol(c_tw_16x18, h, sizeof(float2) * N1 * N2)
                             : cudaMemcpyToSymbol(c_tw_16x8, h, sizeof(float2) * N1 * N2);
  if (e != cudaSuccess) return fail(-3, "pyramid twiddle upload: %s", cudaGetErrorString(e));
  if (dev < 64) g_tw_done[dev][which] = 1;
  return 0;
}

template <int N1, int N2>
static int run(const ColsArgs& ca, const float2* mask_s, int B, int bin, float2* Yt, float* intensity, float* frame, cudaStream_t s) {
  using T = Tile<N1, N2>;
  constexpr int N = T::N;
  const size_t smem = sizeof(float2) * 32 * T::LD;
  cudaError_t e = cudaFuncSetAttribute(pyr_cols_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pyr_rows_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pyr_image_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(-3, "pyramid smem attribute: %s", cudaGetErrorString(e));
  int rc = ensure_tables<N1, N2>();
  if (rc) return rc;
  pyr_cols_kernel<N1, N2><<<dim3((ca.R + 31) / 32, ca.nTheta, B), T::kThreads, smem, s>>>(ca);
  AOENV_LAUNCH_CHECK("pyr_cols");
  pyr_rows_kernel<N1, N2><<<dim3(N / 32, ca.nTheta, B), T::kThreads, smem, s>>>(ca.X1, mask_s, ca.R, ca.lo, ca.nTheta, Yt);
  AOENV_LAUNCH_CHECK("pyr_rows");
  const float scale = 1.0f / ((float)N * (float)N * (float)N * (float)N);       // ifft2 normalisation, squared
  pyr_image_kernel<N1, N2><<<dim3(N / 32, B), T::kThreads, smem, s>>>(Yt, ca.nTheta, scale, intensity);
  AOENV_LAUNCH_CHECK("pyr_image");
  const int nc = N / bin;
  pyr_bin_kernel<<<dim3((nc * nc + 255) / 256, B), 256, 0, s>>>(intensity, N, bin, frame);
  AOENV_LAUNCH_CHECK("pyr_bin");
  return 0;
}

}  // namespace pyr
}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_pyramid_supported(int N) { return N == 288 || N == 128; }

int aoenv_pyramid_frames(const float* opd_a, const float* opd_b, const float* pupil, const float* amp, const float* lin,
                         const float* mod, const float* mask_s, int B, int R, int N, int nTheta, int bin, float phase_scale,
                         float* work_x1, float* work_yt, float* intensity, float* frame, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && R <= N && nTheta > 0 && nTheta <= 65535, "pyramid_frames: bad shape B=%d R=%d N=%d nTheta=%d", B, R, N, nTheta);
  AOENV_CHECK_ARG(aoenv_pyramid_supported(N), "pyramid_frames: N=%d is not a compiled transform size (128, 288)", N);
  AOENV_CHECK_ARG(bin > 0 && N % bin == 0 && (N - R) % 2 == 0, "pyramid_frames: bad binning %d / padding", bin);
  pyr::ColsArgs ca;
  ca.opd_a = opd_a; ca.opd_b = opd_b; ca.pupil = pupil; ca.amp = amp; ca.lin = lin;
  ca.mod = reinterpret_cast<const float2*>(mod);
  ca.X1 = reinterpret_cast<float2*>(work_x1);
  ca.phase_scale = phase_scale; ca.R = R; ca.nTheta = nTheta; ca.lo = (N - R) / 2;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 288)
    return pyr::run<16, 18>(ca, reinterpret_cast<const float2*>(mask_s), B, bin, reinterpret_cast<float2*>(work_yt), intensity, frame, s);
  return pyr::run<16, 8>(ca, reinterpret_cast<const float2*>(mask_s), B, bin, reinterpret_cast<float2*>(work_yt), intensity, frame, s);
}

}  // extern "C"
