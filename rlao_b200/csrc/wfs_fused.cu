// Fused wavefront-sensing kernel of the environment step: DM surface (separable Gaussian influence functions,
// OOPAO/DeformableMirror.py:534-570) + Shack-Hartmann lenslet fields, pruned 2-D DFTs, 2x2 binning
// (OOPAO/ShackHartmann.py:340-353,529-577) + thresholded centre of gravity and slopes (:314-324,580-601) + the pupil
// statistics behind env.total / env.residual / get_strehl (MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:484,502,604-605).
//
// One thread-block CLUSTER per environment; CTA r of the cluster owns a strip of lenslet rows [row_start[r], row_start[r+1])
// (strips are cut so that the work — rows + lit lenslets — is even: the pupil edge has few lit lenslets per row).
//   phase 0  the strip of the atmosphere OPD and the rows of T = C gx (aoenv_dm_rows) the strip's influence bands touch —
//            both contiguous in HBM — arrive in shared memory by TMA bulk copies (cp.async.bulk, one mbarrier);
//   phase D  a thread owns four columns of one lenslet row: DM surface = gy^T T from a register window of T (FFMA2), pupil
//            statistics of the atmosphere and of the residual (locally centred float32 partial sums, promoted to
//            float64), OPD <- atmosphere + DM in place;
//   phase F  warp groups of n/2 warps take 32 lenslets at a time: thread (lenslet, t) turns tile rows 2t, 2t+1 into the
//            complex field (SFU sine / cosine) and leaves it in the group's shared-memory buffer;
//   phase T  thread (lenslet, p) transforms: pass 1 for the DFT rows 2p, 2p+1 (and, by the radix-2 symmetry, n + those)
//            over all columns, pass 2 over all output columns, |.|^2, 2x2 binning -> binned rows p and p + n/2 of the spot,
//            written over the lenslet's own OPD tile (the strip buffer becomes the camera frame);
//   phase S  block maximum -> pushed into every peer's shared memory (DSMEM stores) -> one cluster barrier -> centre of
//            gravity of every lit lenslet from the shared-memory spots -> slopes (+ their split-bf16 planes for the
//            reconstruction GEMM); the strip of the camera frame leaves by TMA bulk stores when the caller asked for it.
// Neither the DM surface nor atmosphere + DM nor (optionally) the camera frame ever touch HBM.
#include <cooperative_groups.h>

#include "common.cuh"
#include "tma.cuh"
#include "wfs_twiddles.inc"

namespace cg = cooperative_groups;

namespace aoenv {

constexpr int kMaxCluster = 16;
// row weights of the DM surface: WL window rows in two halves, each padded to a multiple of 4 floats (128-bit loads)
__host__ __device__ constexpr int wl_half(int WL) { return (WL + 1) / 2; }
__host__ __device__ constexpr int wl_half_pad(int WL) { return (wl_half(WL) + 3) & ~3; }
__host__ __device__ constexpr int wl_stride(int WL) { return 2 * wl_half_pad(WL); }

struct FusedArgs {
  const float* opd_a;
  const float* opd_b;
  aoenv_dm_sep_t dm;               // dm.rows == nullptr: no separable DM (opd_b, possibly null, is the second term)
  const uint8_t* pupil8;
  const int32_t* order;
  const int32_t* nlit;
  const int32_t* slot_of;
  const float* ref_xy;
  float amp0, phase_scale, inv_units, threshold_cog;
  float* frame;
  float* slopes;
  __nv_bfloat16* planes;
  int32_t* envmax;
  double* stats;
  int lds, parts, nV, nS, rows_max, do_slopes;
  int row_start[kMaxCluster + 1];
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tma::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_load_chunks(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  for (uint32_t o = 0; o < bytes; o += 32768u)
    bulk_load(reinterpret_cast<char*>(smem_dst) + o, reinterpret_cast<const char*>(gsrc) + o, min(32768u, bytes - o), bar);
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(tma::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// index of the float2 holding (E[a][2k], E[a][2k+1]) real (reim = 0) or imaginary (1) parts in a group's field buffer,
// lane-minor: consecutive lanes (lenslets) are consecutive 8-byte words -> conflict-free
template <int n>
__device__ __forceinline__ constexpr int fidx(int a, int k, int reim) { return ((a * (n / 2) + k) * 2 + reim) * 32; }

// Binned spot rows PP and PP + n/2 of one lenslet from its field (see shwfs_frame_kernel in wfs.cu for the algebra:
// radix-2 split of both pruned DFT passes, packed FP32 on (row u, row u + n) pairs, literal twiddles).
template <int n, int PP>
__device__ __forceinline__ void spot_rows(const float2* __restrict__ F, float2 (&acc)[n]) {
  constexpr int h = n / 2;
  float2 pr[2][h], pi[2][h], qr[2][h], qi[2][h];
#pragma unroll
  for (int du = 0; du < 2; ++du)
#pragma unroll
    for (int k = 0; k < h; ++k) pr[du][k] = pi[du][k] = qr[du][k] = qi[du][k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int aa = 0; aa < n; ++aa) {
#pragma unroll
    for (int k = 0; k < h; ++k) {
      const float2 er = F[fidx<n>(aa, k, 0)], ei = F[fidx<n>(aa, k, 1)];
#pragma unroll
      for (int du = 0; du < 2; ++du) {
        const float gx = WfsTw<n>::re((2 * PP + du) * n + aa), gy = WfsTw<n>::im((2 * PP + du) * n + aa);
        if (((aa + h) & 1) == 0) {
          pr[du][k] = fma2(dup2(gx), er, fma2(dup2(-gy), ei, pr[du][k]));
          pi[du][k] = fma2(dup2(gx), ei, fma2(dup2(gy), er, pi[du][k]));
        } else {
          qr[du][k] = fma2(dup2(gx), er, fma2(dup2(-gy), ei, qr[du][k]));
          qi[du][k] = fma2(dup2(gx), ei, fma2(dup2(gy), er, qi[du][k]));
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < n; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
  for (int du = 0; du < 2; ++du) {
    float2 yr[n], yi[n];                     // (row u: P + Q, row u + n: P - Q)
#pragma unroll
    for (int bb = 0; bb < n; ++bb) {
      const float p_r = (bb & 1) ? pr[du][bb >> 1].y : pr[du][bb >> 1].x, q_r = (bb & 1) ? qr[du][bb >> 1].y : qr[du][bb >> 1].x;
      const float p_i = (bb & 1) ? pi[du][bb >> 1].y : pi[du][bb >> 1].x, q_i = (bb & 1) ? qi[du][bb >> 1].y : qi[du][bb >> 1].x;
      yr[bb] = make_float2(p_r + q_r, p_r - q_r);
      yi[bb] = make_float2(p_i + q_i, p_i - q_i);
    }
#pragma unroll
    for (int v = 0; v < n; ++v) {
      float2 er_ = make_float2(0.f, 0.f), ei_ = er_, or_ = er_, oi_ = er_;
#pragma unroll
      for (int bb = 0; bb < n; ++bb) {
        const float gx = WfsTw<n>::re(v * n + bb), gy = WfsTw<n>::im(v * n + bb);
        if (((bb + h) & 1) == 0) {
          er_ = fma2(yr[bb], dup2(gx), fma2(yi[bb], dup2(-gy), er_));
          ei_ = fma2(yr[bb], dup2(gy), fma2(yi[bb], dup2(gx), ei_));
        } else {
          or_ = fma2(yr[bb], dup2(gx), fma2(yi[bb], dup2(-gy), or_));
          oi_ = fma2(yr[bb], dup2(gy), fma2(yi[bb], dup2(gx), oi_));
        }
      }
      const float2 f0r = add2(er_, or_), f0i = add2(ei_, oi_);     // F[row][v]
      const float2 f1r = sub2(er_, or_), f1i = sub2(ei_, oi_);     // F[row][v + n]
      acc[v >> 1] = add2(acc[v >> 1], fma2(f0r, f0r, mul2(f0i, f0i)));
      acc[(v >> 1) + h] = add2(acc[(v >> 1) + h], fma2(f1r, f1r, mul2(f1i, f1i)));
    }
  }
}

template <int n>
__device__ __forceinline__ void spot_rows_dispatch(int pp, const float2* __restrict__ F, float2 (&acc)[n]) {
  if (pp == 0) spot_rows<n, 0>(F, acc);
  else if (pp == 1) spot_rows<n, 1>(F, acc);
  else if (n >= 6 && pp == 2) spot_rows<n, (n >= 6 ? 2 : 0)>(F, acc);
  else if (n >= 8) spot_rows<n, (n >= 8 ? 3 : 0)>(F, acc);
}

// shared-memory carve-up, identical on the host (size) and on the device (pointers)
struct FusedSmem {
  int off_S, off_mask, off_F, off_T, off_W, off_misc, total;
};
__host__ __device__ inline FusedSmem fused_smem_layout(int nS, int n, int rows_max, int groups, int t_rows, int WL, bool sep) {
  const int R = nS * n, rows_px = rows_max * n;
  auto up = [](int v) { return (v + 127) & ~127; };
  FusedSmem s;
  int o = 0;
  s.off_S = o;      o = up(o + rows_px * R * 4);
  s.off_mask = o;   o = up(o + rows_px * R);
  s.off_F = o;      o = up(o + groups * n * n * 256);            // n * (n/2) * 2 float2 x 32 lanes per group
  s.off_T = o;      o = up(o + (sep ? t_rows * R * 4 : 0));
  s.off_W = o;      o = up(o + (sep ? rows_px * wl_stride(WL) * 4 : 0));
  s.off_misc = o;   o = up(o + 2048);                            // FusedMisc
  s.total = o;
  return s;
}

struct FusedMisc {
  unsigned long long bar;                     // mbarrier of the strip loads
  double peer_stats[kMaxCluster][4];          // pupil sums pushed by every rank (read by rank 0)
  float peer_max[kMaxCluster];                // spot maxima pushed by every rank
  float warp_max[32];
  double warp_stats[32][4];
  float env_max;
};
static_assert(sizeof(FusedMisc) <= 2048, "misc block must fit the bytes fused_smem_layout reserves for it");

// pupil statistics of one group of four pixels: sums of (x - c) and (x - c)^2 over the pixels inside the pupil
struct Moments {
  float c, s1, s2;
  __device__ __forceinline__ void add(const float4& v, const float4& in) {
    const float d0 = (v.x - c) * in.x, d1 = (v.y - c) * in.y, d2 = (v.z - c) * in.z, d3 = (v.w - c) * in.w;
    s1 += (d0 + d1) + (d2 + d3);
    s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
  }
  __device__ __forceinline__ void add_all(const float4& v) {        // all four pixels inside the pupil
    const float2 nc = dup2(-c);
    const float2 a = add2(make_float2(v.x, v.y), nc), b = add2(make_float2(v.z, v.w), nc);
    const float2 s = add2(a, b), q = fma2(a, a, mul2(b, b));
    s1 += s.x + s.y;
    s2 += q.x + q.y;
  }
};

template <int n, int NG, int WL>
__global__ void __launch_bounds__(NG * (n / 2) * 32, (NG * (n / 2) * 32 <= 384) ? 2 : 1)
shwfs_fused_kernel(const __grid_constant__ FusedArgs p) {
  constexpr int T = n / 2, h = n / 2, N = 2 * n;
  constexpr int kThreads = NG * T * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nS = p.nS, R = nS * n, HP = R / 2, NQ = R / 4;
  const bool sep = p.dm.rows != nullptr;
  const FusedSmem L = fused_smem_layout(nS, n, p.rows_max, NG, p.dm.t_rows, WL, sep);
  float* const S = reinterpret_cast<float*>(smem_raw + L.off_S);
  float2* const S2 = reinterpret_cast<float2*>(S);
  float4* const S4 = reinterpret_cast<float4*>(S);
  const uint16_t* const mask16 = reinterpret_cast<const uint16_t*>(smem_raw + L.off_mask);
  const uint32_t* const mask32 = reinterpret_cast<const uint32_t*>(smem_raw + L.off_mask);
  float2* const Fall = reinterpret_cast<float2*>(smem_raw + L.off_F);
  const float4* const sT4 = reinterpret_cast<const float4*>(smem_raw + L.off_T);
  float* const sW = reinterpret_cast<float*>(smem_raw + L.off_W);
  FusedMisc* const misc = reinterpret_cast<FusedMisc*>(smem_raw + L.off_misc);

  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)gridDim.x;                       // CTAs per environment = cluster size
  const int rank = (int)blockIdx.x, b = (int)blockIdx.y;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr0 = p.row_start[rank], rows = p.row_start[rank + 1] - lr0;      // this CTA's lenslet rows
  const int rows_px = rows * n, row0_px = lr0 * n;
  const size_t env_off = (size_t)b * R * R;
  const uint32_t strip_bytes = (uint32_t)rows_px * R * 4u;

  // ---- phase 0: strip + T rows by TMA bulk copies; pupil mask and row weights by plain loads ---------------------------
  uint64_t* bar = reinterpret_cast<uint64_t*>(&misc->bar);
  int tBase = 0;
  if (sep) tBase = __ldg(&p.dm.ilr[lr0]);
  if (tid == 0) {
    tma::mbar_init(bar, 1);
    tma::mbar_fence_init();
    uint32_t t_bytes = 0;
    if (sep) t_bytes = (uint32_t)(__ldg(&p.dm.ilr[lr0 + rows - 1]) + WL - tBase) * R * 4u;
    tma::mbar_expect_tx(bar, strip_bytes + t_bytes);
    bulk_load_chunks(S, p.opd_a + env_off + (size_t)row0_px * R, strip_bytes, bar);
    if (sep) bulk_load_chunks(smem_raw + L.off_T, p.dm.rows + ((size_t)b * p.dm.nActP + tBase) * R, t_bytes, bar);
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.pupil8 + (size_t)row0_px * R);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw + L.off_mask);
    for (int i = tid; i < rows_px * R / 4; i += kThreads) dst[i] = __ldg(src + i);
    if (sep) {
      const float4* wsrc = reinterpret_cast<const float4*>(p.dm.wlr + (size_t)row0_px * wl_stride(WL));
      for (int i = tid; i < rows_px * wl_stride(WL) / 4; i += kThreads) reinterpret_cast<float4*>(sW)[i] = __ldg(wsrc + i);
    }
  }
  __syncthreads();                 // mask, weights and the mbarrier initialisation are visible
  tma::mbar_wait(bar, 0);

  // ---- phase D: DM surface + statistics, OPD <- atmosphere + DM in place ----------------------------------------------
  double A1 = 0.0, A2 = 0.0, T1 = 0.0, T2 = 0.0;
  for (int item = tid; item < NQ * rows; item += kThreads) {
    const int lr = item / NQ, q = item - lr * NQ;
    const int y0 = lr * n;
    float4 dmv[n];
#pragma unroll
    for (int r = 0; r < n; ++r) dmv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sep) {
      const float4* __restrict__ tw = sT4 + (size_t)(__ldg(&p.dm.ilr[lr0 + lr]) - tBase) * NQ + q;
      // the WL rows of T this lenslet row's bands touch, in two register windows
      constexpr int kHalf = wl_half(WL), kHalfPad = wl_half_pad(WL);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 win[kHalf];
#pragma unroll
        for (int t = 0; t < kHalf; ++t)
          if (half * kHalf + t < WL) win[t] = tw[(size_t)(half * kHalf + t) * NQ];
#pragma unroll
        for (int r = 0; r < n; ++r) {
          const float4* __restrict__ w4 = reinterpret_cast<const float4*>(sW + (y0 + r) * wl_stride(WL) + half * kHalfPad);
          float w[kHalfPad];
#pragma unroll
          for (int j = 0; j < kHalfPad / 4; ++j) {
            const float4 v = w4[j];             // warp-uniform address: broadcast
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
          }
          float2 lo = make_float2(dmv[r].x, dmv[r].y), hi = make_float2(dmv[r].z, dmv[r].w);
#pragma unroll
          for (int t = 0; t < kHalf; ++t)
            if (half * kHalf + t < WL) {
              const float2 ww = dup2(w[t]);
              lo = fma2(ww, make_float2(win[t].x, win[t].y), lo);
              hi = fma2(ww, make_float2(win[t].z, win[t].w), hi);
            }
          dmv[r] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
      }
    } else if (p.opd_b != nullptr) {
      const float4* __restrict__ gb = reinterpret_cast<const float4*>(p.opd_b + env_off + (size_t)(row0_px + y0) * R) + q;
#pragma unroll
      for (int r = 0; r < n; ++r) dmv[r] = __ldg(gb + (size_t)r * NQ);
    }
    uint32_t m[n], all_in = 0x01010101u;
#pragma unroll
    for (int r = 0; r < n; ++r) {
      m[r] = mask32[(y0 + r) * NQ + q];
      all_in &= m[r];
    }
    Moments ma{0.f, 0.f, 0.f}, mt{0.f, 0.f, 0.f};
    float cnt;
    if (all_in == 0x01010101u) {
      cnt = (float)(4 * n);
#pragma unroll
      for (int r = 0; r < n; ++r) {
        const int idx = (y0 + r) * NQ + q;
        const float4 a = S4[idx];
        const float4 t = make_float4(a.x + dmv[r].x, a.y + dmv[r].y, a.z + dmv[r].z, a.w + dmv[r].w);
        S4[idx] = t;
        if (r == 0) { ma.c = a.x; mt.c = t.x; }
        ma.add_all(a);
        mt.add_all(t);
      }
    } else {
      cnt = 0.f;
#pragma unroll
      for (int r = 0; r < n; ++r) {
        const int idx = (y0 + r) * NQ + q;
        const float4 a = S4[idx];
        const float4 t = make_float4(a.x + dmv[r].x, a.y + dmv[r].y, a.z + dmv[r].z, a.w + dmv[r].w);
        S4[idx] = t;
        if (r == 0) { ma.c = a.x; mt.c = t.x; }
        const float4 in = make_float4((m[r] & 0xffu) ? 1.f : 0.f, (m[r] & 0xff00u) ? 1.f : 0.f, (m[r] & 0xff0000u) ? 1.f : 0.f,
                                      (m[r] & 0xff000000u) ? 1.f : 0.f);
        ma.add(a, in);
        mt.add(t, in);
        cnt += (in.x + in.y) + (in.z + in.w);
      }
    }
    if (p.stats != nullptr) {       // sum (x - c) -> sum x, sum (x - c)^2 -> sum x^2, in float64
      const double dca = (double)ma.c, dct = (double)mt.c, dn = (double)cnt;
      A1 += (double)ma.s1 + dn * dca;
      A2 += (double)ma.s2 + 2.0 * dca * (double)ma.s1 + dn * dca * dca;
      T1 += (double)mt.s1 + dn * dct;
      T2 += (double)mt.s2 + 2.0 * dct * (double)mt.s1 + dn * dct * dct;
    }
  }
  if (p.stats != nullptr) {
    A1 = warp_sum(A1); A2 = warp_sum(A2); T1 = warp_sum(T1); T2 = warp_sum(T2);
    if (lane == 0) { misc->warp_stats[warp][0] = A1; misc->warp_stats[warp][1] = A2; misc->warp_stats[warp][2] = T1; misc->warp_stats[warp][3] = T2; }
  }
  __syncthreads();                 // the whole strip now holds atmosphere + DM

  // ---- phases F / T: 32 lenslets at a time per warp group ---------------------------------------------------------
  const int g = warp / T, part = warp - g * T;
  const int LPC = rows * nS;                          // lenslets of this CTA
  const int nlit = __ldg(&p.nlit[rank]);
  const int32_t* __restrict__ order = p.order + (size_t)rank * p.rows_max * nS;
  float2* const F = Fall + (size_t)g * (n * n * 32) + lane;
  const float phase_turns = p.phase_scale * 0.15915494309189535f;
  const float norm = 1.0f / (float)(N * N);
  float vmax = -INFINITY;
  for (int c0 = g * 32; c0 < LPC; c0 += NG * 32) {
    const int idx = c0 + lane;
    const bool lit = idx < nlit;
    const int lens = idx < LPC ? __ldg(&order[idx]) : 0;
    const int lr = lens / nS, l = lens - lr * nS;
    const int cbase = (lr * n) * HP + l * h;          // float2 index of the tile's first row
    const bool any_lit = c0 < nlit;                   // warp-uniform (the list is sorted lit-first)
    if (any_lit) {
      if (lit) {
        float2 rowv[2][h];
        uint32_t rowm[2][h];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int a2 = 0; a2 < h; ++a2) {
            rowv[e][a2] = S2[cbase + (2 * part + e) * HP + a2];
            rowm[e][a2] = mask16[cbase + (2 * part + e) * HP + a2];
          }
#pragma unroll
        for (int aa = 0; aa < n; ++aa) {
          float re[2], im[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float t = (aa & 1) ? rowv[e][aa >> 1].y : rowv[e][aa >> 1].x;
            const bool in = ((rowm[e][aa >> 1] >> ((aa & 1) * 8)) & 0xffu) != 0;
            const float turns = t * phase_turns;
            // nearest integer by the 1.5 * 2^23 trick; then the SFU sine / cosine on [-pi, pi]
            const float ang = (turns - ((turns + 12582912.0f) - 12582912.0f)) * 6.283185307179586f;
            const float am = in ? p.amp0 : 0.f;
            re[e] = am * __cosf(ang);
            im[e] = am * __sinf(ang);
          }
          F[fidx<n>(aa, part, 0)] = make_float2(re[0], re[1]);
          F[fidx<n>(aa, part, 1)] = make_float2(im[0], im[1]);
        }
      }
      group_bar(1 + g, T * 32);
    }
    if (idx < LPC) {
      float2 acc[n];
      if (lit) {
        spot_rows_dispatch<n>(part, F, acc);
      } else {
#pragma unroll
        for (int q = 0; q < n; ++q) acc[q] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int q2 = 0; q2 < h; ++q2) {
        const float2 lo = make_float2(acc[2 * q2].x * norm, acc[2 * q2 + 1].x * norm);
        const float2 hi = make_float2(acc[2 * q2].y * norm, acc[2 * q2 + 1].y * norm);
        S2[cbase + part * HP + q2] = lo;
        S2[cbase + (part + h) * HP + q2] = hi;
        if (lit) vmax = fmaxf(vmax, fmaxf(fmaxf(lo.x, lo.y), fmaxf(hi.x, hi.y)));
      }
    }
    if (any_lit) group_bar(1 + g, T * 32);            // the field buffer is free again
  }

  // ---- phase S: maximum over the environment, centre of gravity, slopes ------------------------------------------
  vmax = warp_max(vmax);
  if (lane == 0) misc->warp_max[warp] = vmax;
  if (p.frame != nullptr) {
    // generic-proxy writes of S must be visible to the async proxy before the bulk store reads them
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();                 // also: every spot of the strip is in S
  // every CTA pushes its maximum into every peer's slot array (and its pupil sums into rank 0's): after ONE cluster
  // barrier each CTA only reads its own shared memory, so nobody has to wait for anybody at exit
  if (tid < C) {
    float m = -INFINITY;
    for (int w = 0; w < kThreads / 32; ++w) m = fmaxf(m, misc->warp_max[w]);
    *cluster.map_shared_rank(&misc->peer_max[rank], tid) = m;
  } else if (p.stats != nullptr && tid >= 32 && tid < 36) {
    const int k = tid - 32;
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += misc->warp_stats[w][k];
    *cluster.map_shared_rank(&misc->peer_stats[rank][k], 0) = s;
  }
  if (p.frame != nullptr && tid == 64) {
    char* dst = reinterpret_cast<char*>(p.frame + env_off + (size_t)row0_px * R);
    for (uint32_t o = 0; o < strip_bytes; o += 32768u)
      bulk_store(dst + o, reinterpret_cast<const char*>(S) + o, min(32768u, strip_bytes - o));
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  cluster.sync();
  if (tid == 0) {
    float m = -INFINITY;
    for (int r = 0; r < C; ++r) m = fmaxf(m, misc->peer_max[r]);
    misc->env_max = m;
    if (rank == 0 && p.envmax != nullptr) p.envmax[b] = float_to_ordered(p.do_slopes ? m : -INFINITY);
  }
  if (p.stats != nullptr && rank == 0 && tid >= 32 && tid < 36) {
    double s = 0.0;
    for (int r = 0; r < C; ++r) s += misc->peer_stats[r][tid - 32];
    p.stats[(size_t)b * 4 + (tid - 32)] = s;
  }
  __syncthreads();
  if (p.do_slopes) {
    const float thr = p.threshold_cog * misc->env_max;
    const size_t plane_stride = (size_t)gridDim.y * p.lds;
    for (int idx = tid; idx < nlit; idx += kThreads) {
      const int lens = __ldg(&order[idx]);
      const int lr = lens / nS, l = lens - lr * nS;
      const int cbase = (lr * n) * HP + l * h;
      float s = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
      for (int pr_ = 0; pr_ < n; ++pr_)
#pragma unroll
        for (int q2 = 0; q2 < h; ++q2) {
          const float2 v = S2[cbase + pr_ * HP + q2];
          const float x0 = v.x < thr ? 0.f : v.x, x1 = v.y < thr ? 0.f : v.y;
          s += x0; sx = fmaf(x0, (float)pr_, sx); sy = fmaf(x0, (float)(2 * q2), sy);
          s += x1; sx = fmaf(x1, (float)pr_, sx); sy = fmaf(x1, (float)(2 * q2 + 1), sy);
        }
      float cx = sx / s, cy = sy / s;
      if (!isfinite(cx)) cx = 0.f;      // ShackHartmann.py:583-593
      if (!isfinite(cy)) cy = 0.f;
      const int t = __ldg(&p.slot_of[(lr0 + lr) * nS + l]);
      const float sx_ = (cx - __ldg(&p.ref_xy[t])) * p.inv_units, sy_ = (cy - __ldg(&p.ref_xy[p.nV + t])) * p.inv_units;
      p.slopes[(size_t)b * p.lds + t] = sx_;
      p.slopes[(size_t)b * p.lds + p.nV + t] = sy_;
      if (p.planes != nullptr) {
        store_bf16_planes(p.planes, plane_stride, (size_t)b * p.lds + t, p.parts, sx_);
        store_bf16_planes(p.planes, plane_stride, (size_t)b * p.lds + p.nV + t, p.parts, sy_);
      }
    }
  }
  if (p.frame != nullptr && tid == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// T[b][i][x] = sum_q C[b][i][j0x[x] + q] wx[x][q]: the column half of the separable DM surface, once per environment and
// command (one CTA per environment; the command image lives in shared memory, zero padded to the right)
template <int W>
__global__ void __launch_bounds__(256)
dm_rows_kernel(const float* __restrict__ coefs, int ldc, const int32_t* __restrict__ act_pos, int nA, int nAct, int nActP,
               const float* __restrict__ wx, const int32_t* __restrict__ j0x, int R, float* __restrict__ rows) {
  pdl_enter();
  extern __shared__ __align__(16) float sC[];          // [nAct][nAct + W]
  const int b = blockIdx.x, ldC = nAct + W;
  for (int k = threadIdx.x; k < nAct * ldC; k += blockDim.x) sC[k] = 0.f;
  __syncthreads();
  for (int k = threadIdx.x; k < nA; k += blockDim.x) {
    const int pos = __ldg(&act_pos[k]);
    const int r = pos / nAct;
    sC[r * ldC + (pos - r * nAct)] = __ldg(&coefs[(size_t)b * ldc + k]);
  }
  __syncthreads();
  float* __restrict__ out = rows + (size_t)b * nActP * R;
  for (int x = threadIdx.x; x < R; x += blockDim.x) {
    float w[W];
    const int j0 = __ldg(&j0x[x]);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(wx + (size_t)x * W) + q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    // gridDim.y CTAs share the rows of one environment (the whole command image is cheap to stage, the rows are the work)
    const int per = (nAct + gridDim.y - 1) / gridDim.y;
    const int i_lo = blockIdx.y * per, i_hi = min(nAct, i_lo + per);
#pragma unroll 4
    for (int i = i_lo; i < i_hi; ++i) {                  // rows are independent: four FMA chains in flight
      const float* __restrict__ c = sC + i * ldC + j0;
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < W; ++q) t = fmaf(c[q], w[q], t);
      out[(size_t)i * R + x] = t;
    }
  }
}

template <int n, int NG, int WL>
static int launch_fused(const FusedArgs& a, int B, int C, size_t smem, cudaStream_t s) {
  auto kern = shwfs_fused_kernel<n, NG, WL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && C > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) return fail(-3, "shwfs_fused attributes: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(NG * (n / 2) * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) return fail(-3, "shwfs_fused launch (cluster %d, %zu B shared): %s", C, smem, cudaGetErrorString(e));
  return 0;
}

}  // namespace aoenv

using namespace aoenv;

extern "C" {

int aoenv_shwfs_fused_smem(int nS, int n, int rows_max, int groups, int t_rows, int WL) {
  if (nS <= 0 || rows_max <= 0 || rows_max > nS || (n != 4 && n != 6 && n != 8)) return -1;
  return fused_smem_layout(nS, n, rows_max, groups, t_rows, WL > 0 ? WL : 14, t_rows > 0).total;
}

int aoenv_dm_rows(const float* coefs, int ldc, const int32_t* act_pos, int nA, int nAct, int nActP, const float* wx,
                  const int32_t* j0x, int W, int B, int R, float* rows, void* stream) {
  AOENV_CHECK_ARG(B > 0 && R > 0 && nAct > 0 && nA > 0 && nA <= nAct * nAct && ldc >= nA && nActP >= nAct, "dm_rows: bad shape");
  AOENV_CHECK_ARG(W == 12 || W == 16, "dm_rows: band width %d (12 or 16)", W);
  const size_t smem = (size_t)nAct * (nAct + W) * sizeof(float);
  AOENV_CHECK_ARG(smem <= 200 * 1024, "dm_rows: %d actuators across do not fit in shared memory", nAct);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = W == 12 ? cudaFuncSetAttribute(dm_rows_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(dm_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(-3, "dm_rows smem attribute: %s", cudaGetErrorString(e));
  const dim3 grid(B, 1);        // (four CTAs per environment measured slower, 35.6 vs 27 us at cfg3: the staging is repeated)
  if (W == 12) AOENV_LAUNCH(dm_rows_kernel<12>, grid, 256, smem, s, coefs, ldc, act_pos, nA, nAct, nActP, wx, j0x, R, rows);
  else AOENV_LAUNCH(dm_rows_kernel<16>, grid, 256, smem, s, coefs, ldc, act_pos, nA, nAct, nActP, wx, j0x, R, rows);
  AOENV_LAUNCH_CHECK("dm_rows");
  return 0;
}

int aoenv_shwfs_fused(const float* opd_a, const float* opd_b, const aoenv_dm_sep_t* dm, const uint8_t* pupil8, float amp0,
                      const int32_t* h_row_start, const int32_t* order, const int32_t* nlit, const int32_t* slot_of, int B,
                      int nS, int n, int cluster, int groups, float phase_scale, const float* ref_xy, int nV, float inv_units,
                      float threshold_cog, float* frame, float* slopes, int lds, void* slope_planes, int parts,
                      int32_t* envmax, double* stats, void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && nS > 0, "shwfs_fused: bad shape B=%d nS=%d", B, nS);
  AOENV_CHECK_ARG(n == 4 || n == 6 || n == 8, "shwfs_fused: %d pixels per lenslet is not a compiled size (4, 6, 8)", n);
  AOENV_CHECK_ARG((nS * n) % 4 == 0, "shwfs_fused: the pupil must be a multiple of 4 pixels across (got %d)", nS * n);
  AOENV_CHECK_ARG(cluster >= 1 && cluster <= kMaxCluster && cluster <= nS, "shwfs_fused: cluster=%d out of range (1..%d)", cluster, kMaxCluster);
  AOENV_CHECK_ARG(groups >= 1 && groups <= 8, "shwfs_fused: groups=%d warp groups per CTA", groups);
  AOENV_CHECK_ARG(slopes == nullptr || (lds >= 2 * nV && nV > 0 && ref_xy != nullptr && slot_of != nullptr), "shwfs_fused: bad slopes arguments");
  AOENV_CHECK_ARG(slope_planes == nullptr || parts == 2 || parts == 3, "shwfs_fused: parts must be 2 or 3");
  AOENV_CHECK_ARG(slopes != nullptr || frame != nullptr, "shwfs_fused: neither slopes nor frame requested");
  AOENV_CHECK_ARG(h_row_start != nullptr && h_row_start[0] == 0 && h_row_start[cluster] == nS, "shwfs_fused: row_start must run from 0 to nS");
  const bool sep = dm != nullptr && dm->rows != nullptr;
  int WL = 14;
  if (sep) {
    AOENV_CHECK_ARG(dm->WL == 14 || dm->WL == 18, "shwfs_fused: DM window of %d actuator rows (14 or 18)", dm->WL);
    AOENV_CHECK_ARG(opd_b == nullptr, "shwfs_fused: give either the separable DM or an explicit second OPD term");
    AOENV_CHECK_ARG(dm->t_rows >= dm->WL && dm->nActP > 0 && dm->wlr != nullptr && dm->ilr != nullptr, "shwfs_fused: bad DM tables");
    WL = dm->WL;
  }
  FusedArgs a{};
  a.opd_a = opd_a; a.opd_b = opd_b;
  if (sep) a.dm = *dm;
  a.pupil8 = pupil8; a.order = order; a.nlit = nlit; a.slot_of = slot_of; a.ref_xy = ref_xy;
  a.amp0 = amp0; a.phase_scale = phase_scale; a.inv_units = inv_units; a.threshold_cog = threshold_cog;
  a.frame = frame; a.slopes = slopes; a.planes = (__nv_bfloat16*)slope_planes; a.envmax = envmax; a.stats = stats;
  a.lds = lds; a.parts = parts; a.nV = nV; a.nS = nS; a.do_slopes = slopes != nullptr;
  int rows_max = 0;
  for (int r = 0; r < cluster; ++r) {
    const int rows = h_row_start[r + 1] - h_row_start[r];
    AOENV_CHECK_ARG(rows >= 1, "shwfs_fused: strip %d is empty", r);
    rows_max = rows > rows_max ? rows : rows_max;
    a.row_start[r] = h_row_start[r];
  }
  a.row_start[cluster] = nS;
  a.rows_max = rows_max;
  const FusedSmem L = fused_smem_layout(nS, n, rows_max, groups, sep ? dm->t_rows : 0, WL, sep);
  AOENV_CHECK_ARG(L.total <= 227 * 1024, "shwfs_fused: %d bytes of shared memory per CTA (cluster %d): use a larger cluster", L.total, cluster);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = -2;
#define AOENV_FUSED_CASE(NN, GG, WW) \
  if (n == NN && groups == GG && WL == WW) rc = launch_fused<NN, GG, WW>(a, B, cluster, (size_t)L.total, s);
  AOENV_FUSED_CASE(4, 4, 14) AOENV_FUSED_CASE(4, 4, 18)
  AOENV_FUSED_CASE(6, 2, 14) AOENV_FUSED_CASE(6, 4, 14) AOENV_FUSED_CASE(6, 4, 18) AOENV_FUSED_CASE(6, 6, 14)
  AOENV_FUSED_CASE(8, 4, 14) AOENV_FUSED_CASE(8, 4, 18)
  if (rc == -2) return fail(-2, "shwfs_fused: no kernel compiled for n=%d groups=%d WL=%d", n, groups, WL);
#undef AOENV_FUSED_CASE
  if (rc != 0) return rc;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
