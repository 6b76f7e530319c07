// Microbenchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, int iters) {
  float2 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f);
  const float2 m = make_float2(s, s * 0.999f), a = make_float2(1e-3f, 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {
        acc[j].x = fmaf(acc[j].x, m.x, a.x);
        acc[j].y = fmaf(acc[j].y, m.y, a.y);
      } else {
        acc[j] = ffma2(acc[j], m, a);
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) r += acc[j].x + acc[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out;
  const int blocks = 148 * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, 256>>>(out, 0.9999f, iters); else k<1><<<blocks, 256>>>(out, 0.9999f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 16 * (double)iters * blocks * 256;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s FP32\n", mode ? "FFMA2 (f32x2)" : "FFMA (scalar)", ms, flops / ms / 1e9);
    }
  }
  return 0;
}
