// Atmosphere kernels: von Karman infinite phase screens (Assemat et al. 2006) as OOPAO/Atmosphere.py runs them,
// batched over environments.  All are HBM-bound streaming kernels; the dense part of add_row (X = A Z + B xi)
// is a GEMM in gemm.cu.
#include "common.cuh"

namespace aoenv {

// ---------------------------------------------------------------------------------------------------------
// add_row step 1: gather the two inner rings of the one-pixel-shifted map + the innovation vector
// (OOPAO/Atmosphere.py:303-308).  One thread per entry of zx[b][:].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
atm_gather_kernel(const float* __restrict__ map, int M, int pitch, int sx, int sy,
                  const int2* __restrict__ inner_rc, int nI, int nO, const float* __restrict__ xi,
                  uint64_t seed, uint64_t stream_id, float* __restrict__ zx, int ldz) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ldz) return;
  float v = 0.f;
  if (k < nI) {
    const int2 rc = __ldg(&inner_rc[k]);
    v = __ldg(&map[((size_t)b * M + (rc.x - sy)) * pitch + (rc.y - sx)]);
  } else if (k < nI + nO) {
    const int j = k - nI;
    if (xi != nullptr) {
      v = __ldg(&xi[(size_t)b * nO + j]);
    } else {
      Philox rng(seed);
      const uint4 r = rng((uint32_t)j, (uint32_t)b, (uint32_t)stream_id, (uint32_t)(stream_id >> 32));
      v = box_muller(r.x, r.y).x;
    }
  }
  zx[(size_t)b * ldz + k] = v;
}

// ---------------------------------------------------------------------------------------------------------
// add_row step 3: write the shifted interior and the freshly extruded outer ring into the other map buffer
// (OOPAO/Atmosphere.py:309-310) and reduce the new map's extrema.  The ring index follows numpy boolean-mask
// order of `outerMask` (:263-264): row 0, then (r,0),(r,M-1) pairs, then row M-1.
// ---------------------------------------------------------------------------------------------------------
__global__ void minmax_init_kernel(int32_t* __restrict__ minmax, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    minmax[2 * b + 0] = float_to_ordered(INFINITY);
    minmax[2 * b + 1] = float_to_ordered(-INFINITY);
  }
}

__device__ __forceinline__ void block_minmax_commit(float lo, float hi, int32_t* __restrict__ mm) {
  __shared__ float s_lo[32], s_hi[32];
  lo = warp_min(lo);
  hi = warp_max(hi);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    lo = l < nw ? s_lo[l] : INFINITY;
    hi = l < nw ? s_hi[l] : -INFINITY;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (l == 0) {
      atomicMin(&mm[0], float_to_ordered(lo));
      atomicMax(&mm[1], float_to_ordered(hi));
    }
  }
}

// One thread produces 4 consecutive columns of one row (one aligned 128-bit store; pitch % 4 == 0); the shifted
// source row is read with 4 scalar loads that coalesce across the warp whatever the shift.
__global__ void __launch_bounds__(256)
atm_scatter_kernel(const float* __restrict__ map_in, float* __restrict__ map_out, int M, int pitch, int sx, int sy,
                   const float* __restrict__ X, int ldx, int32_t* __restrict__ minmax, int rows_per_block) {
  const int b = blockIdx.y;
  const size_t base = (size_t)b * M * pitch;
  const float* __restrict__ xb = X + (size_t)b * ldx;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(M, r_begin + rows_per_block);
  const int nvec = pitch >> 2;
  const int rows_in_flight = blockDim.x / 64;          // 64 threads (256 columns) per row
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  float lo = INFINITY, hi = -INFINITY;
  const int last_vec = (M - 1) >> 2;                   // vector holding column M-1
  auto slow_vec = [&](int r, int v4) {                 // vectors that touch the ring or the padding
    const int c = v4 << 2;
    const bool ring_row = (r == 0) || (r == M - 1);
    const float* __restrict__ src = map_in + base + (size_t)(r - sy) * pitch - sx;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cc = c + k;
      float x = 0.f;
      if (cc < M) {
        if (ring_row) x = __ldg(&xb[(r == 0 ? 0 : M + 2 * (M - 2)) + cc]);
        else if (cc == 0) x = __ldg(&xb[M + 2 * (r - 1)]);
        else if (cc == M - 1) x = __ldg(&xb[M + 2 * (r - 1) + 1]);
        else x = __ldg(&src[cc]);
        lo = fminf(lo, x);
        hi = fmaxf(hi, x);
      }
      v[k] = x;
    }
    *reinterpret_cast<float4*>(map_out + base + (size_t)r * pitch + c) = make_float4(v[0], v[1], v[2], v[3]);
  };
  for (int r = r_begin + ty; r < r_end; r += 2 * rows_in_flight) {
    const int r2 = r + rows_in_flight;
    const bool two = r2 < r_end;
    const bool fast1 = r > 0 && r < M - 1;
    const bool fast2 = two && r2 > 0 && r2 < M - 1;
    for (int v4 = tx; v4 < nvec; v4 += 64) {
      const bool interior_vec = v4 > 0 && v4 < last_vec;
      if (interior_vec && fast1 && fast2) {
        // both rows: plain shifted copy, 8 independent loads in flight
        const int c = v4 << 2;
        const float* __restrict__ s1 = map_in + base + (size_t)(r - sy) * pitch - sx + c;
        const float* __restrict__ s2 = map_in + base + (size_t)(r2 - sy) * pitch - sx + c;
        const float a0 = __ldg(s1), a1 = __ldg(s1 + 1), a2 = __ldg(s1 + 2), a3 = __ldg(s1 + 3);
        const float b0 = __ldg(s2), b1 = __ldg(s2 + 1), b2 = __ldg(s2 + 2), b3 = __ldg(s2 + 3);
        *reinterpret_cast<float4*>(map_out + base + (size_t)r * pitch + c) = make_float4(a0, a1, a2, a3);
        *reinterpret_cast<float4*>(map_out + base + (size_t)r2 * pitch + c) = make_float4(b0, b1, b2, b3);
        lo = fminf(fminf(fminf(lo, a0), fminf(a1, a2)), fminf(fminf(a3, b0), fminf(b1, fminf(b2, b3))));
        hi = fmaxf(fmaxf(fmaxf(hi, a0), fmaxf(a1, a2)), fmaxf(fmaxf(a3, b0), fmaxf(b1, fmaxf(b2, b3))));
      } else {
        slow_vec(r, v4);
        if (two) slow_vec(r2, v4);
      }
    }
  }
  block_minmax_commit(lo, hi, &minmax[2 * b]);
}

__global__ void __launch_bounds__(256)
map_minmax_kernel(const float* __restrict__ map, int M, int pitch, int32_t* __restrict__ minmax, int rows_per_block) {
  const int b = blockIdx.y;
  const size_t base = (size_t)b * M * pitch;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(M, r_begin + rows_per_block);
  float lo = INFINITY, hi = -INFINITY;
  for (int r = r_begin; r < r_end; ++r)
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
      const float v = __ldg(&map[base + (size_t)r * pitch + c]);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
  block_minmax_commit(lo, hi, &minmax[2 * b]);
}

// ---------------------------------------------------------------------------------------------------------
// Sub-pixel shift + clip + footprint crop + Cn2-weighted layer sum (OOPAO/Atmosphere.py:406-407,439-450,474-478;
// interpolation: scikit-image 0.18.3 bicubic, see oracle/warp018.py).  Each thread produces a 1 x ROWS column
// strip: the horizontal 4-tap pass is done once per input row and reused by the vertical 4-tap pass.
// ---------------------------------------------------------------------------------------------------------
struct AtmPhaseParams {
  const float* map[AOENV_MAX_LAYERS];
  const int32_t* minmax[AOENV_MAX_LAYERS];
  int row_off[AOENV_MAX_LAYERS];
  int col_off[AOENV_MAX_LAYERS];
  float wrow[AOENV_MAX_LAYERS][4];
  float wcol[AOENV_MAX_LAYERS][4];
  float weight[AOENV_MAX_LAYERS];
  int nLayer;
};

// Tile = 128 x 32 output pixels per CTA of 128 threads; each thread owns a 4 (columns) x 8 (rows) register block.
// Per layer the (32+3) x (128+3) input window is staged in shared memory with coalesced loads (its global alignment
// depends on the layer's tap offset, the shared-memory copy is 16-byte aligned for every thread), then the separable
// interpolation runs out of shared memory with two 128-bit loads per input row.
constexpr int kPhTileW = 128, kPhTileH = 32, kPhThreadsX = 32, kPhThreadsY = 4, kPhRows = 8;
constexpr int kPhSmemW = kPhTileW + 8;    // 3 extra taps, padded to a multiple of 4

__global__ void __launch_bounds__(kPhThreadsX * kPhThreadsY)
atm_phase_kernel(const __grid_constant__ AtmPhaseParams p, int R, int M, int pitch, int fp_off, float opd_scale,
                 float* __restrict__ opd_out) {
  __shared__ __align__(16) float tile[kPhTileH + 3][kPhSmemW];
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * kPhTileW, i0 = blockIdx.y * kPhTileH;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * kPhThreadsX + tx;
  float acc[kPhRows][4];
#pragma unroll
  for (int r = 0; r < kPhRows; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  for (int l = 0; l < p.nLayer; ++l) {
    const float* __restrict__ m = p.map[l] + (size_t)b * M * pitch;
    const int cbase = j0 + fp_off + p.col_off[l];
    const int rbase = i0 + fp_off + p.row_off[l];
    __syncthreads();
    // asynchronous 4-byte copies (LDGSTS): every load of the window is in flight before the first one lands
    for (int rr = ty; rr < kPhTileH + 3; rr += kPhThreadsY) {           // one warp per input row, lanes along columns
      const float* __restrict__ grow = m + (size_t)min(rbase + rr, M - 1) * pitch;   // clamps only touch unused outputs
      const uint32_t srow = (uint32_t)__cvta_generic_to_shared(&tile[rr][0]);
#pragma unroll
      for (int cc = tx; cc < kPhTileW + 3; cc += kPhThreadsX)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(srow + 4u * cc), "l"(grow + min(cbase + cc, M - 1)) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float lo = ordered_to_float(__ldg(&p.minmax[l][2 * b + 0]));
    const float hi = ordered_to_float(__ldg(&p.minmax[l][2 * b + 1]));
    const float wc0 = p.wcol[l][0], wc1 = p.wcol[l][1], wc2 = p.wcol[l][2], wc3 = p.wcol[l][3];
    const float wr0 = p.wrow[l][0], wr1 = p.wrow[l][1], wr2 = p.wrow[l][2], wr3 = p.wrow[l][3];
    const float w = p.weight[l];
    float h[kPhRows + 3][4];
#pragma unroll
    for (int t = 0; t < kPhRows + 3; ++t) {
      const float4 a = *reinterpret_cast<const float4*>(&tile[ty * kPhRows + t][tx * 4]);
      const float4 c = *reinterpret_cast<const float4*>(&tile[ty * kPhRows + t][tx * 4 + 4]);
      h[t][0] = wc0 * a.x + wc1 * a.y + wc2 * a.z + wc3 * a.w;
      h[t][1] = wc0 * a.y + wc1 * a.z + wc2 * a.w + wc3 * c.x;
      h[t][2] = wc0 * a.z + wc1 * a.w + wc2 * c.x + wc3 * c.y;
      h[t][3] = wc0 * a.w + wc1 * c.x + wc2 * c.y + wc3 * c.z;
    }
#pragma unroll
    for (int t = 0; t < kPhRows; ++t)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = wr0 * h[t][c] + wr1 * h[t + 1][c] + wr2 * h[t + 2][c] + wr3 * h[t + 3][c];
        v = fminf(fmaxf(v, lo), hi);
        acc[t][c] += w * v;
      }
  }
  float* __restrict__ out = opd_out + (size_t)b * R * R;
  const int jc = j0 + tx * 4;
#pragma unroll
  for (int t = 0; t < kPhRows; ++t) {
    const int i = i0 + ty * kPhRows + t;
    if (i >= R || jc >= R) continue;
    float* __restrict__ o = out + (size_t)i * R + jc;
    if (jc + 3 < R && ((R & 3) == 0)) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[t][0] * opd_scale, acc[t][1] * opd_scale, acc[t][2] * opd_scale, acc[t][3] * opd_scale);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (jc + c < R) o[c] = acc[t][c] * opd_scale;
    }
  }
}

}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_atm_gather(const float* map, int B, int M, int pitch, int sx, int sy, const int32_t* inner_rc, int nI,
                     int nO, const float* xi, uint64_t seed, uint64_t stream_id, float* zx, int ldz, void* stream) {
  AOENV_CHECK_ARG(B > 0 && M > 6 && pitch >= M, "atm_gather: bad shape B=%d M=%d pitch=%d", B, M, pitch);
  AOENV_CHECK_ARG(sx >= -1 && sx <= 1 && sy >= -1 && sy <= 1, "atm_gather: shift must be in {-1,0,1}");
  AOENV_CHECK_ARG(ldz >= nI + nO, "atm_gather: ldz=%d < nI+nO=%d", ldz, nI + nO);
  dim3 grid((ldz + 255) / 256, B);
  atm_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(map, M, pitch, sx, sy, (const int2*)inner_rc, nI, nO, xi,
                                                            seed, stream_id, zx, ldz);
  AOENV_LAUNCH_CHECK("atm_gather");
  return 0;
}

static int rows_per_block_for(int B, int M) {
  // enough blocks to fill 148 SMs a few times over even for a handful of environments
  int chunks = (4 * kNumSMs + B - 1) / B;
  if (chunks < 1) chunks = 1;
  if (chunks > M) chunks = M;
  return (M + chunks - 1) / chunks;
}

int aoenv_atm_scatter(const float* map_in, float* map_out, int B, int M, int pitch, int sx, int sy, int nO,
                      const float* X, int ldx, int32_t* minmax, void* stream) {
  AOENV_CHECK_ARG(B > 0 && M > 6 && pitch >= M && pitch % 4 == 0, "atm_scatter: bad shape (pitch must be a multiple of 4)");
  AOENV_CHECK_ARG(nO == 4 * M - 4 && ldx >= nO, "atm_scatter: ring has %d pixels, got nO=%d ldx=%d", 4 * M - 4, nO, ldx);
  AOENV_CHECK_ARG(map_in != map_out, "atm_scatter: in-place shift is not supported");
  cudaStream_t s = (cudaStream_t)stream;
  minmax_init_kernel<<<(B + 255) / 256, 256, 0, s>>>(minmax, B);
  AOENV_LAUNCH_CHECK("minmax_init");
  const int rpb = rows_per_block_for(B, M);
  dim3 grid((M + rpb - 1) / rpb, B);
  atm_scatter_kernel<<<grid, 256, 0, s>>>(map_in, map_out, M, pitch, sx, sy, X, ldx, minmax, rpb);
  AOENV_LAUNCH_CHECK("atm_scatter");
  return 0;
}

int aoenv_map_minmax(const float* map, int B, int M, int pitch, int32_t* minmax, void* stream) {
  AOENV_CHECK_ARG(B > 0 && M > 0 && pitch >= M, "map_minmax: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  minmax_init_kernel<<<(B + 255) / 256, 256, 0, s>>>(minmax, B);
  AOENV_LAUNCH_CHECK("minmax_init");
  const int rpb = rows_per_block_for(B, M);
  dim3 grid((M + rpb - 1) / rpb, B);
  map_minmax_kernel<<<grid, 256, 0, s>>>(map, M, pitch, minmax, rpb);
  AOENV_LAUNCH_CHECK("map_minmax");
  return 0;
}

int aoenv_atm_phase(const float* const* h_map, const int32_t* const* h_minmax, int nLayer, int B, int R, int M,
                    int pitch, int fp_off, const int32_t* h_row_off, const int32_t* h_col_off, const float* h_wrow,
                    const float* h_wcol, const float* h_weight, float opd_scale, float* opd_out, void* stream) {
  AOENV_CHECK_ARG(nLayer >= 1 && nLayer <= AOENV_MAX_LAYERS, "atm_phase: nLayer=%d out of range", nLayer);
  AOENV_CHECK_ARG(B > 0 && B <= 65535, "atm_phase: B=%d out of range (1..65535)", B);
  AtmPhaseParams p;
  p.nLayer = nLayer;
  for (int l = 0; l < nLayer; ++l) {
    p.map[l] = h_map[l];
    p.minmax[l] = h_minmax[l];
    p.row_off[l] = h_row_off[l];
    p.col_off[l] = h_col_off[l];
    for (int k = 0; k < 4; ++k) {
      p.wrow[l][k] = h_wrow[4 * l + k];
      p.wcol[l][k] = h_wcol[4 * l + k];
    }
    p.weight[l] = h_weight[l];
    // every tap of every footprint pixel must lie inside the map
    const int lo_r = fp_off + h_row_off[l], hi_r = fp_off + R - 1 + h_row_off[l] + 3;
    const int lo_c = fp_off + h_col_off[l], hi_c = fp_off + R - 1 + h_col_off[l] + 3;
    AOENV_CHECK_ARG(lo_r >= 0 && lo_c >= 0 && hi_r < M && hi_c < M, "atm_phase: taps of layer %d leave the map", l);
  }
  dim3 block(kPhThreadsX, kPhThreadsY);
  dim3 grid((R + kPhTileW - 1) / kPhTileW, (R + kPhTileH - 1) / kPhTileH, B);
  atm_phase_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(p, R, M, pitch, fp_off, opd_scale, opd_out);
  AOENV_LAUNCH_CHECK("atm_phase");
  return 0;
}

}  // extern "C"
