// Tensor-core "TN" GEMM for sm_100a with FP32-grade accuracy from split-bf16 operands.
//
//   D[x][w] = alpha * sum_k X[x][k] * W[w][k]        X: [MX][K] (per-environment vectors: DM commands, [Z|xi],
//                                                       slopes), W: [NW][K] (static operators: influence functions,
//                                                       [A|B], reconstructor); D row-major [MX][ldd].
//
// Each float32 operand is stored as `parts` bf16 planes (x = x_0 + x_1 (+ x_2), x_0 = bf16(x), x_1 = bf16(x - x_0), ..),
// and the product keeps every cross term down to 2^-16 (parts = 2: 3 MMAs per k-step, error ~2^-17 relative to
// sum |x||w|) or 2^-24 (parts = 3: 6 MMAs).  The tensor cores accumulate in FP32 in tensor memory.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the W and X part tiles (128B-swizzled, K-major) into a
//               ring of shared-memory stages, completion on mbarriers;
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (cta_group::1, M = 128 rows of W in the TMEM
//               lanes, N = BLOCK_N rows of X in the TMEM columns, K = 16 per instruction), tcgen05.commit releases
//               the stage and, after the last k-block, publishes the accumulator;
//   warps 2-5   epilogue: tcgen05.ld the accumulator (each warp its own 32-lane quarter), scale, store — lanes are
//               consecutive W rows (= consecutive addresses of D's contiguous axis), so every store instruction
//               writes one full 128-byte line;
//   two accumulator buffers in TMEM let the epilogue of tile t overlap the MMAs of tile t+1.
#include <cuda.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tma.cuh"

namespace aoenv {
namespace tc {

constexpr int BLOCK_M = 128;     // W rows per tile (TMEM lanes)
constexpr int BLOCK_K = 64;      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kMaxParts = 3;

struct TensorMaps {
  CUtensorMap w[kMaxParts];
  CUtensorMap x[kMaxParts];
};

// ---- PTX wrappers (TMA / mbarrier ones live in tma.cuh) -----------------------------------------------------------
using tma::smem_u32;
using tma::mbar_init;
using tma::mbar_expect_tx;
using tma::mbar_arrive;
using tma::mbar_wait;
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  tma::load_2d(map, bar, smem, c0, c1);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// K-major, 128B-swizzled operand tile whose rows are 128 bytes: 8-row atoms of 1024 bytes (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)0 << 16;                          // leading byte offset (unused: one swizzle atom along K)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, shape M x N.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Work list of one CTA.  A segment is (output tile mn, k-blocks [kb0, kb1)).
//   * tile mode: segments t = blockIdx.x, + gridDim.x, ... of num_mn * split_k; segment t covers slice t / num_mn of K;
//   * stream-K mode (more tiles than SMs, not a multiple of them): the num_mn * num_kb units (tile-major) are cut into
//     gridDim.x equal contiguous ranges, so every SM does the same number of k-blocks instead of one SM wave running
//     nearly empty.  A range is at least num_kb long, hence a tile is shared by at most two CTAs; both add their partial
//     sum to a zeroed D with float atomics, and two addends commute — the result stays bit-reproducible.
struct Seg { int mn, kb0, kb1; };
struct SegIter {
  int stream, num_mn, num_kb, kb_per, t, t_end, stride;
  __device__ SegIter(int stream_k, int num_mn_, int num_kb_, int split_k)
      : stream(stream_k), num_mn(num_mn_), num_kb(num_kb_), kb_per((num_kb_ + split_k - 1) / split_k) {
    if (stream) {
      const long long units = (long long)num_mn * num_kb;
      t = (int)(units * blockIdx.x / gridDim.x);
      t_end = (int)(units * (blockIdx.x + 1) / gridDim.x);
      stride = 0;
    } else {
      t = blockIdx.x;
      t_end = num_mn * split_k;
      stride = gridDim.x;
    }
  }
  __device__ __forceinline__ bool next(Seg& s) {
    if (t >= t_end) return false;
    if (stream) {
      s.mn = t / num_kb;
      s.kb0 = t - s.mn * num_kb;
      const int len = min(num_kb - s.kb0, t_end - t);
      s.kb1 = s.kb0 + len;
      t += len;
    } else {
      s.mn = t % num_mn;
      s.kb0 = (t / num_mn) * kb_per;
      s.kb1 = min(num_kb, s.kb0 + kb_per);
      t += stride;
    }
    return true;
  }
};

template <int kParts, int BLOCK_N>
struct Cfg {
  static constexpr int kStages = 2;
  static constexpr int kWBytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kXBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kParts * (kWBytes + kXBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  // parts == 3 keeps the leading product x_0 w_0 and the five cross terms in separate accumulators (summed in FP32 by
  // the epilogue): the tensor core adds into the running sum with truncation, so the error grows with the number of
  // MMAs that touch a LARGE accumulator — this keeps it to K/16 instead of 6K/16.
  static constexpr int kAccPerBuf = (kParts == 3) ? 2 : 1;
  static constexpr int kTmemCols = 2 * kAccPerBuf * BLOCK_N;   // two accumulator buffers
  static_assert(kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int kParts, int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ TensorMaps maps, float* __restrict__ D, int ldd, int MX, int NW, int K, float alpha,
               int split_k, int stream_k) {
  using C = Cfg<kParts, BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + C::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * C::kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (NW + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (MX + BLOCK_N - 1) / BLOCK_N;
  const int num_mn = num_m * num_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  // The three roles walk the same list of segments (output tile, k-block range) — see SegIter.
  const SegIter seg_begin(stream_k, num_mn, num_kb, split_k);

  pdl_trigger();                                   // the successor may be scheduled while this grid's last wave runs
  if (warp == 0 && lane == 0) {
    for (int p = 0; p < kParts; ++p) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[p])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.x[p])) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);      // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                      // barriers, TMEM and tensor-map prefetch are set up; now the operands

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it = seg_begin;
      Seg sg;
      while (it.next(sg)) {
        const int m_blk = sg.mn / num_n, n_blk = sg.mn % num_n;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::kStageBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
          for (int p = 0; p < kParts; ++p) {
            tma_load_2d(&maps.w[p], &full_bar[stage], st + p * C::kWBytes, kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(&maps.x[p], &full_bar[stage], st + kParts * C::kWBytes + p * C::kXBytes, kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SegIter it = seg_begin;
      Seg sg;
      while (it.next(sg)) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * C::kAccPerBuf * BLOCK_N;
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
          uint32_t first = (kb == kb0) ? 1u : 0u;      // first MMA into the main accumulator
          uint32_t first_x = (kb == kb0) ? 1u : 0u;    // first MMA into the cross-term accumulator
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // cross terms in increasing order of magnitude last: (pw, px) with pw + px < kParts
#pragma unroll
            for (int pw = kParts - 1; pw >= 0; --pw) {
#pragma unroll
              for (int px = kParts - 1; px >= 0; --px) {
                if (pw + px >= kParts) continue;
                const uint64_t da = smem_desc_sw128(st + pw * C::kWBytes + k * UMMA_K * 2);
                const uint64_t db = smem_desc_sw128(st + kParts * C::kWBytes + px * C::kXBytes + k * UMMA_K * 2);
                if (C::kAccPerBuf == 2 && (pw + px) > 0) {
                  umma_bf16(tmem_d + BLOCK_N, da, db, idesc, first_x ? 0u : 1u);
                  first_x = 0;
                } else {
                  umma_bf16(tmem_d, da, db, idesc, first ? 0u : 1u);
                  first = 0;
                }
              }
            }
          }
          umma_commit(&empty_bar[stage]);                     // frees the stage when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);     // accumulator complete
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue (warps 2..5) ==============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    SegIter it = seg_begin;
    Seg sg;
    while (it.next(sg)) {
      const int m_blk = sg.mn / num_n, n_blk = sg.mn % num_n;
      const bool partial = sg.kb1 - sg.kb0 < num_kb;     // this CTA holds only part of the sum over K
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int w_row = m_blk * BLOCK_M + q * 32 + lane;
      const bool w_ok = w_row < NW;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::kAccPerBuf * BLOCK_N + c0);
        tmem_ld32(taddr, v);
        if (C::kAccPerBuf == 2) {
          uint32_t u[32];
          tmem_ld32(taddr + BLOCK_N, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        const int x0 = n_blk * BLOCK_N + c0;
        if (w_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (x0 + j < MX) {
              float* dst = &D[(size_t)(x0 + j) * ldd + w_row];
              const float val = alpha * __uint_as_float(v[j]);
              if (partial) atomicAdd(dst, val);           // D was zeroed; two partials add commutatively -> deterministic
              else *dst = val;
            }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

// ---- operand splitting ---------------------------------------------------------------------------------------
template <int kParts>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, int lds, int rows, int K, __nv_bfloat16* __restrict__ dst, int ldk,
                  size_t plane) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ldk) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    float x = (k < K) ? __ldg(&src[(size_t)r * lds + k]) : 0.f;
#pragma unroll
    for (int p = 0; p < kParts; ++p) {
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      dst[(size_t)p * plane + (size_t)r * ldk + k] = h;
      x -= __bfloat162float(h);
    }
  }
}

// ---- host side: tensor maps ----------------------------------------------------------------------------------
using tma::EncodeTiledFn;
static EncodeTiledFn get_encode() { return tma::get_encode(); }

static int make_map(CUtensorMap* m, const void* base, int rows, int K, int ldk, int box_rows) {
  static thread_local tma::MapCache<32> cache;
  const tma::MapKey key{base, ((unsigned long long)(unsigned)rows << 32) | (unsigned)K,
                        ((unsigned long long)(unsigned)ldk << 32) | (unsigned)box_rows};
  return cache.get(key, m, [&](CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(-4, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldk * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-4, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ldk=%d", (int)r, rows, K, ldk);
    return 0;
  });
}

template <int kParts, int BLOCK_N>
static int launch(const __nv_bfloat16* Xs, const __nv_bfloat16* Ws, int ldk, float* D, int ldd, int MX, int NW, int K,
                  float alpha, cudaStream_t s) {
  using C = Cfg<kParts, BLOCK_N>;
  TensorMaps maps;
  for (int p = 0; p < kParts; ++p) {
    int rc = make_map(&maps.w[p], Ws + (size_t)p * NW * ldk, NW, K, ldk, BLOCK_M);
    if (rc) return rc;
    rc = make_map(&maps.x[p], Xs + (size_t)p * MX * ldk, MX, K, ldk, BLOCK_N);
    if (rc) return rc;
  }
  // function attributes are per device: remember which devices have been configured (one process may drive several)
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((configured.load(std::memory_order_relaxed) >> dev) & 1)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<kParts, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return fail(-3, "gemm_tc smem attribute: %s", cudaGetErrorString(e));
    if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  const int num_mn = ((NW + BLOCK_M - 1) / BLOCK_M) * ((MX + BLOCK_N - 1) / BLOCK_N);
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  // small outputs leave most SMs idle: split K in two (two partial sums added with float atomics commute, so the
  // result stays bit-reproducible; more than two would not)
  const int split_k = (2 * num_mn <= kNumSMs && num_kb >= 8) ? 2 : 1;
  // more tiles than SMs and not a multiple of them: equal k-block ranges per SM (see SegIter)
  const int stream_k = (num_mn > kNumSMs && num_mn % kNumSMs != 0) ? 1 : 0;
  if (split_k > 1 || stream_k) {
    cudaError_t e = cudaMemset2DAsync(D, sizeof(float) * (size_t)ldd, 0, sizeof(float) * (size_t)NW, (size_t)MX, s);
    if (e != cudaSuccess) return fail(-3, "gemm_tc memset: %s", cudaGetErrorString(e));
  }
  const int num_tiles = num_mn * split_k;
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  AOENV_LAUNCH((gemm_tc_kernel<kParts, BLOCK_N>), dim3(grid), kThreads, C::kSmemBytes, s, maps, D, ldd, MX, NW, K, alpha, split_k,
               stream_k);
  AOENV_LAUNCH_CHECK("gemm_tc");
  return 0;
}

}  // namespace tc
}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_split_bf16(const float* src, int lds, int rows, int K, int parts, void* dst, int ldk, void* stream) {
  AOENV_CHECK_ARG(rows > 0 && K > 0 && lds >= K && ldk >= K && ldk % 8 == 0, "split_bf16: bad shape rows=%d K=%d lds=%d ldk=%d", rows, K, lds, ldk);
  AOENV_CHECK_ARG(parts == 2 || parts == 3, "split_bf16: parts must be 2 or 3");
  dim3 grid((ldk + 255) / 256, rows < 65535 ? rows : 65535);
  const size_t plane = (size_t)rows * ldk;
  if (parts == 2)
    tc::split_bf16_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, K, (__nv_bfloat16*)dst, ldk, plane);
  else
    tc::split_bf16_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, K, (__nv_bfloat16*)dst, ldk, plane);
  AOENV_LAUNCH_CHECK("split_bf16");
  return 0;
}

int aoenv_gemm_tn_tc(const void* Xs, const void* Ws, int ldk, int parts, float* D, int ldd, int MX, int NW, int K,
                     float alpha, void* stream) {
  AOENV_CHECK_ARG(MX > 0 && NW > 0 && K > 0 && ldk >= K && ldk % 8 == 0, "gemm_tn_tc: bad shape MX=%d NW=%d K=%d ldk=%d", MX, NW, K, ldk);
  AOENV_CHECK_ARG(ldd >= NW, "gemm_tn_tc: ldd=%d < NW=%d", ldd, NW);
  AOENV_CHECK_ARG(((reinterpret_cast<uintptr_t>(Xs) | reinterpret_cast<uintptr_t>(Ws)) & 15) == 0, "gemm_tn_tc: operands must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (parts == 2) return tc::launch<2, 256>((const __nv_bfloat16*)Xs, (const __nv_bfloat16*)Ws, ldk, D, ldd, MX, NW, K, alpha, s);
  if (parts == 3) return tc::launch<3, 128>((const __nv_bfloat16*)Xs, (const __nv_bfloat16*)Ws, ldk, D, ldd, MX, NW, K, alpha, s);
  return fail(-2, "gemm_tn_tc: parts must be 2 or 3");
}

}  // extern "C"
