// Fused wavefront-sensing kernel of the environment step: DM surface (separable Gaussian influence functions,
// OOPAO/DeformableMirror.py:534-570) + Shack-Hartmann lenslet fields, pruned 2-D DFTs, 2x2 binning
// (OOPAO/ShackHartmann.py:340-353,529-577) + thresholded centre of gravity and slopes (:314-324,580-601) + the pupil
// statistics behind env.total / env.residual / get_strehl (MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:484,502,604-605).
//
// One thread-block CLUSTER per environment; CTA r of the cluster owns a strip of nS/C lenslet rows.
//   phase 0  the strip of the atmosphere OPD (contiguous in HBM) arrives in shared memory by one TMA bulk copy
//            (cp.async.bulk, mbarrier completion) while the CTA builds T = C gx for the actuator rows its strip sees;
//   phase D  every thread owns a column pair of one lenslet row: DM surface = gy^T T for its 2 x n pixels (FFMA2), pupil
//            statistics of the atmosphere and of the residual (locally centred float32 partial sums, promoted to
//            float64), OPD <- atmosphere + DM in place;
//   phase F  warp groups of n/2 warps take 32 lenslets at a time: thread (lenslet, t) turns tile rows 2t, 2t+1 into the
//            complex field (SFU sine / cosine) and leaves it in the group's shared-memory buffer;
//   phase T  thread (lenslet, p) transforms: pass 1 for the DFT rows 2p, 2p+1 (and, by the radix-2 symmetry, n + those)
//            over all columns, pass 2 over all output columns, |.|^2, 2x2 binning -> binned rows p and p + n/2 of the spot,
//            written over the lenslet's own OPD tile (the strip buffer becomes the camera frame);
//   phase S  block maximum -> cluster maximum through distributed shared memory -> centre of gravity of every lit
//            lenslet from the shared-memory spots -> slopes (+ their split-bf16 planes for the reconstruction GEMM);
//            the strip of the camera frame leaves by one TMA bulk store when the caller asked for it.
// Neither the DM surface nor atmosphere + DM nor (optionally) the camera frame ever touch HBM.
#include <cooperative_groups.h>

#include "common.cuh"
#include "tma.cuh"
#include "wfs_twiddles.inc"

namespace cg = cooperative_groups;

namespace aoenv {

struct FusedArgs {
  const float* opd_a;
  const float* opd_b;
  aoenv_dm_sep_t dm;               // dm.coefs == nullptr: no separable DM (opd_b, possibly null, is the second term)
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
  int lds, parts, nV, nS, rows_per_cta, do_slopes;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tma::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tma::smem_u32(bar))
               : "memory");
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
  int off_S, off_mask, off_F, off_T, off_C, off_Wy, off_I0, off_misc, total;
};
__host__ __device__ inline FusedSmem fused_smem_layout(int nS, int n, int rows_per_cta, int groups, int t_rows, int nAct, int W,
                                                        bool sep) {
  const int R = nS * n, rows_px = rows_per_cta * n;
  auto up = [](int v) { return (v + 127) & ~127; };
  FusedSmem s;
  int o = 0;
  s.off_S = o;      o = up(o + rows_px * R * 4);
  s.off_mask = o;   o = up(o + rows_px * R);
  s.off_F = o;      o = up(o + groups * n * n * 256);            // n * (n/2) * 2 float2 x 32 lanes per group
  s.off_T = o;      o = up(o + (sep ? t_rows * R * 4 : 0));
  s.off_C = o;      o = up(o + (sep ? t_rows * nAct * 4 : 0));
  s.off_Wy = o;     o = up(o + (sep ? (rows_px / 2) * 2 * W * 4 : 0));
  s.off_I0 = o;     o = up(o + (sep ? (rows_px / 2) * 4 : 0));
  s.off_misc = o;   o = up(o + 1536);                          // FusedMisc
  s.total = o;
  return s;
}

struct FusedMisc {
  unsigned long long bar;         // mbarrier of the strip load
  double cta_stats[4];            // this CTA's pupil sums, read by rank 0 through DSMEM
  float cta_max;                  // this CTA's spot maximum, read by every rank through DSMEM
  float env_max;
  float warp_max[32];
  double warp_stats[32][4];
};
static_assert(sizeof(FusedMisc) <= 1536, "misc block must fit the bytes fused_smem_layout reserves for it");

template <int n, int NG, int W>
__global__ void __launch_bounds__(NG * (n / 2) * 32, (NG * (n / 2) * 32 <= 384) ? 2 : 1)
shwfs_fused_kernel(const __grid_constant__ FusedArgs p) {
  constexpr int T = n / 2, h = n / 2, N = 2 * n;
  constexpr int kThreads = NG * T * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nS = p.nS, R = nS * n, HP = R / 2;
  const int rows_px = p.rows_per_cta * n;
  const bool sep = p.dm.coefs != nullptr;
  const FusedSmem L = fused_smem_layout(nS, n, p.rows_per_cta, NG, p.dm.t_rows, p.dm.nAct, W, sep);
  float* const S = reinterpret_cast<float*>(smem_raw + L.off_S);
  float2* const S2 = reinterpret_cast<float2*>(S);
  const uint16_t* const mask16 = reinterpret_cast<const uint16_t*>(smem_raw + L.off_mask);
  float2* const Fall = reinterpret_cast<float2*>(smem_raw + L.off_F);
  float* const sT = reinterpret_cast<float*>(smem_raw + L.off_T);
  float* const sC = reinterpret_cast<float*>(smem_raw + L.off_C);
  float* const sWy = reinterpret_cast<float*>(smem_raw + L.off_Wy);
  int* const sI0 = reinterpret_cast<int*>(smem_raw + L.off_I0);
  FusedMisc* const misc = reinterpret_cast<FusedMisc*>(smem_raw + L.off_misc);

  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)gridDim.x;                       // CTAs per environment = cluster size
  const int rank = (int)blockIdx.x, b = (int)blockIdx.y;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0_px = rank * rows_px;                 // first pixel row of this CTA's strip
  const size_t env_off = (size_t)b * R * R;
  const uint32_t strip_bytes = (uint32_t)rows_px * R * 4u;

  // ---- phase 0: strip load (TMA bulk) + DM stage 1 -------------------------------------------------------------
  uint64_t* bar = reinterpret_cast<uint64_t*>(&misc->bar);
  if (tid == 0) {
    tma::mbar_init(bar, 1);
    tma::mbar_fence_init();
    tma::mbar_expect_tx(bar, strip_bytes);
    const char* src = reinterpret_cast<const char*>(p.opd_a + env_off + (size_t)row0_px * R);
    for (uint32_t o = 0; o < strip_bytes; o += 32768u) {
      const uint32_t len = min(32768u, strip_bytes - o);
      bulk_load(reinterpret_cast<char*>(S) + o, src + o, len, bar);
    }
  }
  {
    // pupil mask bytes of the strip (contiguous; a multiple of 4 bytes because n is even)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.pupil8 + (size_t)row0_px * R);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw + L.off_mask);
    for (int i = tid; i < rows_px * R / 4; i += kThreads) dst[i] = __ldg(src + i);
  }
  int tBase = 0;
  if (sep) {
    const int nAct = p.dm.nAct, t_rows = p.dm.t_rows;
    const int pair0 = row0_px >> 1, npair = rows_px >> 1;
    tBase = __ldg(&p.dm.i0y[pair0]);
    for (int i = tid; i < t_rows * nAct; i += kThreads) sC[i] = 0.f;
    for (int i = tid; i < npair * 2 * W / 4; i += kThreads)
      reinterpret_cast<float4*>(sWy)[i] = __ldg(reinterpret_cast<const float4*>(p.dm.wyp + (size_t)pair0 * 2 * W) + i);
    for (int i = tid; i < npair; i += kThreads) sI0[i] = __ldg(&p.dm.i0y[pair0 + i]);
    __syncthreads();
    // commands of the actuator rows [tBase, tBase + t_rows): act_row_start[r] = first valid-actuator index of row r
    const int r_end = min(nAct, tBase + t_rows);
    const int k0 = __ldg(&p.dm.act_row_start[tBase]), k1 = __ldg(&p.dm.act_row_start[r_end]);
    for (int k = k0 + tid; k < k1; k += kThreads)
      sC[__ldg(&p.dm.act_pos[k]) - tBase * nAct] = __ldg(&p.dm.coefs[(size_t)b * p.dm.ldc + k]);
    __syncthreads();
    // T[i][x] = sum_q C[i][j0(x) + q] wx[x][q]: a thread keeps the W weights of its column and walks a slice of the rows
    const int slices = max(1, kThreads / R);
    const int per = (t_rows + slices - 1) / slices;
    for (int item = tid; item < R * slices; item += kThreads) {
      const int sl = item / R, x = item - sl * R;
      float w[W];
      const int j0 = __ldg(&p.dm.j0x[x]);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.dm.wx + (size_t)x * W) + q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      const int i_end = min(t_rows, (sl + 1) * per);
      for (int i = sl * per; i < i_end; ++i) {
        const float* __restrict__ c = sC + i * nAct;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < W; ++q) t = fmaf(c[min(j0 + q, nAct - 1)], w[q], t);
        sT[i * R + x] = t;
      }
    }
  }
  __syncthreads();                 // T, weights, mask and the mbarrier initialisation are visible
  tma::mbar_wait(bar, 0);

  // ---- phase D: DM surface + statistics, OPD <- atmosphere + DM in place ----------------------------------------------
  double A1 = 0.0, A2 = 0.0, T1 = 0.0, T2 = 0.0;
  for (int item = tid; item < HP * p.rows_per_cta; item += kThreads) {
    const int lr = item / HP, j = item - lr * HP;
    const int y0 = lr * n;
    float2 dmv[n];
    if (sep) {
      const float2* __restrict__ sT2 = reinterpret_cast<const float2*>(sT);
      const int nAct = p.dm.nAct;
#pragma unroll
      for (int k = 0; k < h; ++k) {
        const int pk = (y0 >> 1) + k;
        const int i0 = sI0[pk];
        float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
        for (int t4 = 0; t4 < W / 4; ++t4) {
          const float4 u0 = reinterpret_cast<const float4*>(sWy + (size_t)pk * 2 * W)[t4];
          const float4 u1 = reinterpret_cast<const float4*>(sWy + (size_t)pk * 2 * W + W)[t4];
          const float w0[4] = {u0.x, u0.y, u0.z, u0.w}, w1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 v = sT2[(min(i0 + 4 * t4 + q, nAct - 1) - tBase) * HP + j];
            a0 = fma2(dup2(w0[q]), v, a0);
            a1 = fma2(dup2(w1[q]), v, a1);
          }
        }
        dmv[2 * k] = a0;
        dmv[2 * k + 1] = a1;
      }
    } else if (p.opd_b != nullptr) {
      const float2* __restrict__ gb = reinterpret_cast<const float2*>(p.opd_b + env_off + (size_t)(row0_px + y0) * R) + j;
#pragma unroll
      for (int r = 0; r < n; ++r) dmv[r] = __ldg(gb + (size_t)r * HP);
    } else {
#pragma unroll
      for (int r = 0; r < n; ++r) dmv[r] = make_float2(0.f, 0.f);
    }
    float ca = 0.f, ct = 0.f, sa1 = 0.f, sa2 = 0.f, st1 = 0.f, st2 = 0.f, cnt = 0.f;
#pragma unroll
    for (int r = 0; r < n; ++r) {
      const int idx = (y0 + r) * HP + j;
      const float2 a = S2[idx];
      const float2 t = make_float2(a.x + dmv[r].x, a.y + dmv[r].y);
      S2[idx] = t;
      if (r == 0) { ca = a.x; ct = t.x; }
      const uint32_t m = mask16[idx];
      const float in0 = (m & 0xffu) ? 1.f : 0.f, in1 = (m >> 8) ? 1.f : 0.f;
      const float da0 = (a.x - ca) * in0, da1 = (a.y - ca) * in1, dt0 = (t.x - ct) * in0, dt1 = (t.y - ct) * in1;
      sa1 += da0 + da1; sa2 = fmaf(da0, da0, fmaf(da1, da1, sa2));
      st1 += dt0 + dt1; st2 = fmaf(dt0, dt0, fmaf(dt1, dt1, st2));
      cnt += in0 + in1;
    }
    if (p.stats != nullptr) {       // sum (x - c) -> sum x, sum (x - c)^2 -> sum x^2, in float64
      const double dca = (double)ca, dct = (double)ct, dn = (double)cnt;
      A1 += (double)sa1 + dn * dca;
      A2 += (double)sa2 + 2.0 * dca * (double)sa1 + dn * dca * dca;
      T1 += (double)st1 + dn * dct;
      T2 += (double)st2 + 2.0 * dct * (double)st1 + dn * dct * dct;
    }
  }
  if (p.stats != nullptr) {
    A1 = warp_sum(A1); A2 = warp_sum(A2); T1 = warp_sum(T1); T2 = warp_sum(T2);
    if (lane == 0) { misc->warp_stats[warp][0] = A1; misc->warp_stats[warp][1] = A2; misc->warp_stats[warp][2] = T1; misc->warp_stats[warp][3] = T2; }
  }
  __syncthreads();                 // the whole strip now holds atmosphere + DM
  if (p.stats != nullptr && tid < 4) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += misc->warp_stats[w][tid];
    misc->cta_stats[tid] = s;
  }

  // ---- phases F / T: 32 lenslets at a time per warp group ---------------------------------------------------------
  const int g = warp / T, part = warp - g * T;
  const int LPC = p.rows_per_cta * nS;                // lenslets of this CTA
  const int nlit = __ldg(&p.nlit[rank]);
  const int32_t* __restrict__ order = p.order + (size_t)rank * LPC;
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
  __syncthreads();                 // also: every spot of the strip is in S
  if (tid == 0) {
    float m = -INFINITY;
    for (int w = 0; w < kThreads / 32; ++w) m = fmaxf(m, misc->warp_max[w]);
    misc->cta_max = m;
  }
  if (p.frame != nullptr) {
    // generic-proxy writes of S must be visible to the async proxy before the bulk store reads them
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (C > 1) cluster.sync(); else __syncthreads();
  if (p.frame != nullptr && tid == 0) {
    char* dst = reinterpret_cast<char*>(p.frame + env_off + (size_t)row0_px * R);
    for (uint32_t o = 0; o < strip_bytes; o += 32768u)
      bulk_store(dst + o, reinterpret_cast<const char*>(S) + o, min(32768u, strip_bytes - o));
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (tid == 0) {
    float m = misc->cta_max;
    for (int r = 0; r < C; ++r)
      if (r != rank) m = fmaxf(m, *cluster.map_shared_rank(&misc->cta_max, r));
    misc->env_max = m;
    if (rank == 0 && p.envmax != nullptr) p.envmax[b] = float_to_ordered(p.do_slopes ? m : -INFINITY);
  }
  if (p.stats != nullptr && rank == 0 && tid < 4) {
    double s = misc->cta_stats[tid];
    for (int r = 1; r < C; ++r) s += *cluster.map_shared_rank(&misc->cta_stats[tid], r);
    p.stats[(size_t)b * 4 + tid] = s;
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
      const int t = __ldg(&p.slot_of[(rank * p.rows_per_cta + lr) * nS + l]);
      const float sx_ = (cx - __ldg(&p.ref_xy[t])) * p.inv_units, sy_ = (cy - __ldg(&p.ref_xy[p.nV + t])) * p.inv_units;
      p.slopes[(size_t)b * p.lds + t] = sx_;
      p.slopes[(size_t)b * p.lds + p.nV + t] = sy_;
      if (p.planes != nullptr) {
        store_bf16_planes(p.planes, plane_stride, (size_t)b * p.lds + t, p.parts, sx_);
        store_bf16_planes(p.planes, plane_stride, (size_t)b * p.lds + p.nV + t, p.parts, sy_);
      }
    }
  }
  if (p.frame != nullptr && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (C > 1) cluster.sync();       // nobody leaves while a peer may still read its shared memory
}

template <int n, int NG, int W>
static int launch_fused(const FusedArgs& a, int B, int C, size_t smem, cudaStream_t s) {
  auto kern = shwfs_fused_kernel<n, NG, W>;
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

int aoenv_shwfs_fused_smem(int nS, int n, int cluster, int groups, int t_rows, int nAct, int W) {
  if (nS <= 0 || cluster <= 0 || nS % cluster != 0 || (n != 4 && n != 6 && n != 8)) return -1;
  return fused_smem_layout(nS, n, nS / cluster, groups, t_rows, nAct, W, t_rows > 0).total;
}

int aoenv_shwfs_fused(const float* opd_a, const float* opd_b, const aoenv_dm_sep_t* dm, const uint8_t* pupil8, float amp0,
                      const int32_t* order, const int32_t* nlit, const int32_t* slot_of, int B, int nS, int n, int cluster,
                      int groups, float phase_scale, const float* ref_xy, int nV, float inv_units, float threshold_cog,
                      float* frame, float* slopes, int lds, void* slope_planes, int parts, int32_t* envmax, double* stats,
                      void* stream) {
  AOENV_CHECK_ARG(B > 0 && B <= 65535 && nS > 0, "shwfs_fused: bad shape B=%d nS=%d", B, nS);
  AOENV_CHECK_ARG(n == 4 || n == 6 || n == 8, "shwfs_fused: %d pixels per lenslet is not a compiled size (4, 6, 8)", n);
  AOENV_CHECK_ARG(cluster >= 1 && cluster <= 16 && nS % cluster == 0, "shwfs_fused: cluster=%d must divide nS=%d (1..16)", cluster, nS);
  AOENV_CHECK_ARG(groups >= 1 && groups <= 8, "shwfs_fused: groups=%d warp groups per CTA", groups);
  AOENV_CHECK_ARG(slopes == nullptr || (lds >= 2 * nV && nV > 0 && ref_xy != nullptr && slot_of != nullptr), "shwfs_fused: bad slopes arguments");
  AOENV_CHECK_ARG(slope_planes == nullptr || parts == 2 || parts == 3, "shwfs_fused: parts must be 2 or 3");
  AOENV_CHECK_ARG(slopes != nullptr || frame != nullptr, "shwfs_fused: neither slopes nor frame requested");
  const bool sep = dm != nullptr && dm->coefs != nullptr;
  int W = 12;
  if (sep) {
    AOENV_CHECK_ARG(dm->W == 12 || dm->W == 16, "shwfs_fused: DM band width %d (12 or 16)", dm->W);
    AOENV_CHECK_ARG(opd_b == nullptr, "shwfs_fused: give either the separable DM or an explicit second OPD term");
    AOENV_CHECK_ARG(dm->t_rows > 0 && dm->t_rows <= dm->nAct && dm->ldc >= dm->nA && dm->act_row_start != nullptr, "shwfs_fused: bad DM tables");
    W = dm->W;
  }
  FusedArgs a{};
  a.opd_a = opd_a; a.opd_b = opd_b;
  if (sep) a.dm = *dm;
  a.pupil8 = pupil8; a.order = order; a.nlit = nlit; a.slot_of = slot_of; a.ref_xy = ref_xy;
  a.amp0 = amp0; a.phase_scale = phase_scale; a.inv_units = inv_units; a.threshold_cog = threshold_cog;
  a.frame = frame; a.slopes = slopes; a.planes = (__nv_bfloat16*)slope_planes; a.envmax = envmax; a.stats = stats;
  a.lds = lds; a.parts = parts; a.nV = nV; a.nS = nS; a.rows_per_cta = nS / cluster; a.do_slopes = slopes != nullptr;
  const FusedSmem L = fused_smem_layout(nS, n, nS / cluster, groups, sep ? dm->t_rows : 0, sep ? dm->nAct : 0, W, sep);
  AOENV_CHECK_ARG(L.total <= 227 * 1024, "shwfs_fused: %d bytes of shared memory per CTA (cluster %d): use a larger cluster", L.total, cluster);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = -2;
#define AOENV_FUSED_CASE(NN, GG, WW) \
  if (n == NN && groups == GG && W == WW) rc = launch_fused<NN, GG, WW>(a, B, cluster, (size_t)L.total, s);
  AOENV_FUSED_CASE(4, 4, 12) AOENV_FUSED_CASE(4, 4, 16)
  AOENV_FUSED_CASE(6, 2, 12) AOENV_FUSED_CASE(6, 4, 12) AOENV_FUSED_CASE(6, 4, 16) AOENV_FUSED_CASE(6, 6, 12)
  AOENV_FUSED_CASE(8, 4, 12) AOENV_FUSED_CASE(8, 4, 16)
  if (rc == -2) return fail(-2, "shwfs_fused: no kernel compiled for n=%d groups=%d W=%d", n, groups, W);
#undef AOENV_FUSED_CASE
  if (rc != 0) return rc;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
