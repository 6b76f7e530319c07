// Shared device/host helpers for libaoenv_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/aoenv.h"

namespace aoenv {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char* fmt, ...);

#define AOENV_CHECK_ARG(cond, ...)                         \
  do {                                                     \
    if (!(cond)) return ::aoenv::fail(-2, __VA_ARGS__);    \
  } while (0)

#define AOENV_LAUNCH_CHECK(name)                                                        \
  do {                                                                                  \
    ::aoenv::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return ::aoenv::fail(-3, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// The step is a chain of ~13 dependent kernels on one stream; launched the ordinary way each of them starts only after
// the previous grid has drained completely and its completion has travelled back to the front end.  With the
// programmatic-stream-serialization attribute the next grid is scheduled as soon as every CTA of the current one has
// been started (pdl_enter() issues griddepcontrol.launch_dependents first thing), takes the SM slots that free up while
// the last wave drains, and parks at griddepcontrol.wait until the predecessor's memory is complete and visible.
// EVERY thread calls pdl_enter() before anything else — in particular before any early return: a grid that could finish
// without waiting would let ITS successor run ahead of the predecessor.
extern int g_pdl;                       // 1 unless AOENV_PDL=0 (api.cu)
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// setup that touches no global memory may run between the two halves (gemm_tc: barriers, TMEM allocation)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct PdlLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
};
// AOENV_LAUNCH(kernel, grid, block, smem, stream, args...): kernels launched this way must start with pdl_enter().
#define AOENV_LAUNCH(kernel, grid, block, smem, stream, ...)                  \
  do {                                                                        \
    ::aoenv::PdlLaunch l__(grid, block, smem, stream);                        \
    cudaLaunchKernelEx(&l__.cfg, kernel, __VA_ARGS__);                        \
  } while (0)

// ---- monotone float <-> int encoding so atomicMax/atomicMin on int32 order floats ----------------------
__device__ __forceinline__ int32_t float_to_ordered(float f) {
  int32_t i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int32_t i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: one issue slot, two IEEE-exact FP32 operations) --------------
// A float2 built as make_float2(s, s) from a scalar (register, uniform register, immediate) is encoded by ptxas as a
// broadcast operand, so "pair x scalar" costs no extra instruction.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 dup2(float s) { return make_float2(s, s); }

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011) ------------------------------------------
struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ Philox(uint64_t seed) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
  }
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// uniform in (0, 1]
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = u32_to_unit(a);
  const float u2 = u32_to_unit(b);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// x = x_0 + x_1 (+ x_2), x_0 = bf16(x), x_1 = bf16(x - x_0), ...: the operand format of the split-bf16 tensor-core GEMM
// (gemm_tc.cu).  Producers of a GEMM operand call this to write the planes next to (or instead of) the float32 value.
__device__ __forceinline__ void store_bf16_planes(__nv_bfloat16* __restrict__ planes, size_t plane_stride, size_t idx,
                                                  int parts, float x) {
  for (int p = 0; p < parts; ++p) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    planes[(size_t)p * plane_stride + idx] = h;
    x -= __bfloat162float(h);
  }
}

}  // namespace aoenv
