// FP32 SIMT "TN" GEMM: D[m][n] = alpha * sum_k X[m][k] * W[n][k], both operands K-contiguous.
// Serves the three dense contractions of the step (DM surface, add_row predictor, reconstructor) in exact FP32;
// the tensor-core (tcgen05, split-bf16) path lives in gemm_tc.cu.
#include "common.cuh"

namespace aoenv {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int TM = 8, TN = 8;        // per-thread micro-tile
constexpr int kThreads = 256;        // (BM/TM) * (BN/TN)
constexpr int PAD = 4;

// Loads a [rows x BK] tile (K-contiguous in global memory) and stores it transposed: s[k][row].
__device__ __forceinline__ void load_tile(const float* __restrict__ g, int ld, int row0, int nrows, int k0,
                                          float4 (&reg)[2]) {
  // 128 rows x 16 k = 512 float4; 256 threads x 2
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int idx = threadIdx.x + t * kThreads;
    const int row = idx >> 2;          // 0..127
    const int kq = (idx & 3) * 4;      // 0,4,8,12
    const int gr = row0 + row;
    reg[t] = gr < nrows ? __ldg(reinterpret_cast<const float4*>(g + (size_t)gr * ld + k0 + kq))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void store_tile(float (*s)[BM + PAD], const float4 (&reg)[2]) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int idx = threadIdx.x + t * kThreads;
    const int row = idx >> 2;
    const int kq = (idx & 3) * 4;
    s[kq + 0][row] = reg[t].x;
    s[kq + 1][row] = reg[t].y;
    s[kq + 2][row] = reg[t].z;
    s[kq + 3][row] = reg[t].w;
  }
}

__global__ void __launch_bounds__(kThreads)
gemm_tn_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, float* __restrict__ D,
               int ldd, int M, int N, int K, float alpha) {
  __shared__ __align__(16) float sX[2][BK][BM + PAD];
  __shared__ __align__(16) float sW[2][BK][BN + PAD];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tm = (threadIdx.x / (BN / TN)) * TM;   // 0..120
  const int tn = (threadIdx.x % (BN / TN)) * TN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 rx[2], rw[2];
  load_tile(X, ldx, m0, M, 0, rx);
  load_tile(W, ldw, n0, N, 0, rw);
  store_tile(sX[0], rx);
  store_tile(sW[0], rw);
  __syncthreads();

  const int nk = K / BK;
  for (int kb = 0; kb < nk; ++kb) {
    const int cur = kb & 1;
    if (kb + 1 < nk) {
      load_tile(X, ldx, m0, M, (kb + 1) * BK, rx);
      load_tile(W, ldw, n0, N, (kb + 1) * BK, rw);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&sX[cur][k][tm]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sX[cur][k][tm + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sW[cur][k][tn]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sW[cur][k][tn + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tile(sX[cur ^ 1], rx);
      store_tile(sW[cur ^ 1], rw);
    }
    __syncthreads();
  }

  const bool vec_ok = (ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm + i;
    if (m >= M) continue;
    float* __restrict__ drow = D + (size_t)m * ldd + n0 + tn;
    if (vec_ok && n0 + tn + TN <= N) {
      *reinterpret_cast<float4*>(drow) = make_float4(alpha * acc[i][0], alpha * acc[i][1], alpha * acc[i][2], alpha * acc[i][3]);
      *reinterpret_cast<float4*>(drow + 4) = make_float4(alpha * acc[i][4], alpha * acc[i][5], alpha * acc[i][6], alpha * acc[i][7]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n0 + tn + j < N) drow[j] = alpha * acc[i][j];
    }
  }
}

// A handful of X rows (one or a few environments): the tiled kernel above would run three CTAs through ~100 barriers, and
// the tensor-core kernel spends longer setting up than computing.  One warp per output column streams its W row once
// (128-bit loads) against up to MR rows of X, then reduces across lanes.
template <int MR>
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, float* __restrict__ D, int ldd,
                   int M, int N, int K, float alpha) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int m0 = blockIdx.y * MR;
  if (n >= N) return;
  float acc[MR];
#pragma unroll
  for (int r = 0; r < MR; ++r) acc[r] = 0.f;
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(W + (size_t)n * ldw);
  for (int k4 = lane; k4 < K / 4; k4 += 32) {
    const float4 w = __ldg(w4 + k4);
#pragma unroll
    for (int r = 0; r < MR; ++r)
      if (m0 + r < M) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * ldx) + k4);
        acc[r] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[r]))));
      }
  }
#pragma unroll
  for (int r = 0; r < MR; ++r) {
    float v = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && m0 + r < M) D[(size_t)(m0 + r) * ldd + n] = alpha * v;
  }
}

}  // namespace aoenv

using namespace aoenv;

extern "C" int aoenv_gemm_tn(const float* X, int ldx, const float* W, int ldw, float* D, int ldd, int M, int N, int K,
                             float alpha, void* stream) {
  AOENV_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tn: empty problem M=%d N=%d K=%d", M, N, K);
  AOENV_CHECK_ARG(K % BK == 0, "gemm_tn: K=%d must be a multiple of %d (zero-pad the operands)", K, BK);
  AOENV_CHECK_ARG(ldx % 4 == 0 && ldw % 4 == 0 && ldx >= K && ldw >= K, "gemm_tn: ldx=%d ldw=%d must be >= K and multiples of 4", ldx, ldw);
  AOENV_CHECK_ARG(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) == 0, "gemm_tn: operands must be 16-byte aligned");
  AOENV_CHECK_ARG(ldd >= N, "gemm_tn: ldd=%d < N=%d", ldd, N);
  if (M <= AOENV_SKINNY_MAX_ROWS) {
    AOENV_LAUNCH(gemm_skinny_kernel<8>, dim3((N + 7) / 8, (M + 7) / 8), 256, 0, (cudaStream_t)stream, X, ldx, W, ldw, D, ldd, M, N,
                 K, alpha);
    AOENV_LAUNCH_CHECK("gemm_tn(skinny)");
    return 0;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  AOENV_CHECK_ARG(grid.y <= 65535, "gemm_tn: M too large");
  gemm_tn_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(X, ldx, W, ldw, D, ldd, M, N, K, alpha);
  AOENV_LAUNCH_CHECK("gemm_tn");
  return 0;
}
