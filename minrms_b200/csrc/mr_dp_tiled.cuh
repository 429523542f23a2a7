// mr_dp_tiled.cuh -- register-tiled 3-D DP kernel (K1+K3+K4 for problems whose (layer,row)
// cross-section fits one CTA's registers).
//
// Recurrence (Dyn3d::Solve_IncrN, lib/dyn/dyn3d.cc:183-387; layer 1: InitTable_IncrN 135-179):
//     A(n,i,j) = min{ A(n-1,i-1,j-1) + C(i,j),  A(n,i-1,j),  A(n,i,j-1) }
// with the reference's tie order (match, then "<=" skip-i, then "<=" skip-j) recorded as two
// bits per cell.  Written over the whole (n,i,j) box with the boundary values
//     A(0,.,.) = 0,   A(n>=1, 0, j) = A(n>=1, i, 0) = +inf
// it reproduces the reference's special-cased first row / first column of every layer
// exactly: cells with i < n or j < n stay +inf and never win a min, live cells only ever
// read live cells or those boundaries.
//
// Mapping (one CTA = one orientation):
//   * the sweep axis is j (columns of molecule 2), one column per step;
//   * a thread owns a K x R register tile: K consecutive layers x R consecutive rows;
//     lane = layer chunk (32 lanes x K layers >= n_max + 1; layer 0 is the constant-0 base
//     layer so the lowest real layer needs no special case), warp = row strip;
//   * A(n-1,i-1,j-1) for the lowest layer of a lane comes from the neighbouring lane's
//     previous-column registers by warp shuffle (the previous column is complete, so no
//     skew is needed between lanes); inside the tile it is a register of the same thread;
//   * A(n,i-1,j) across strips goes through a double-buffered shared-memory row of
//     boundary values; strip w runs one column behind strip w-1 (wavefront over warps),
//     one __syncthreads per step;
//   * C(i,j) is built once per orientation into shared memory, transposed ([j][i]) and
//     padded with +inf rows/columns so pipeline fill/drain steps leave the state untouched;
//   * the two tie-break predicates are taken from the sign bits of (d - up) and
//     (best - left) (exact for IEEE floats) and funnel-shifted into a per-layer bit
//     accumulator: 3 FADD (FMA pipe) + 1 FMNMX3 + 2 SHF (ALU pipe) per cell.
#pragma once
#include <algorithm>
#include "mr_device.cuh"

// Traceback storage of the tiled kernel.  One 32-bit word per (strip, column pair, layer):
// 2 columns x R=8 rows x 2 bits, appended MSB-first in processing order.
// bit pair of cell (row r of the strip, column half h): s1 at bit 31-(16h+2r), s2 one below.
//   s1 = 1  <=>  NOT (up <= d)        s2 = 1  <=>  NOT (left <= min(d,up))
struct TbTiled {
  int K, R, W, JP, pad, NP;  // layers/lane, rows/strip, strips, column pairs, top padding rows
  __host__ __device__ size_t words(int /*nB*/) const { return (size_t)W * JP * 32 * K; }
  __device__ __forceinline__ uint32_t code(const uint32_t* __restrict__ tbo, int n, int i, int j) const {
    const int rho = i - 1 + pad;
    const int w = rho / R, r = rho - w * R;
    const int c = n / K, k = n - c * K;
    const int jp = (j - 1) >> 1, h = (j - 1) & 1;
    const uint32_t word = tbo[((size_t)w * JP + jp) * (32 * K) + k * 32 + c];
    const int a = h * 2 * R + 2 * r;
    const uint32_t s1 = (word >> (31 - a)) & 1u, s2 = (word >> (30 - a)) & 1u;
    return s2 ? (s1 ? 0u : 1u) : 2u;
  }
};

struct MrTiledPlan {
  bool ok = false;
  const char* why = "not planned";
  int K = 0, R = 8, W = 0;
  int nBe = 0;            // columns rounded up to even
  int threads = 0;
  size_t smem = 0;
  int ctas_per_sm = 1;
  int grid_cap = 0;
  size_t tb_words = 0;
  TbTiled layout;
};

template <int K, int R, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
k_dp_tiled(MrDevParams P, TbTiled L, int nBe, const float* __restrict__ A, const float* __restrict__ B,
           const float* __restrict__ orient, const int* __restrict__ surv,
           const int* __restrict__ surv_count, uint32_t* __restrict__ tb, float* __restrict__ cost_out) {
  extern __shared__ float smem[];
  const int m = P.m, nB = P.nB, nb = P.nb;
  const int W = L.W, WR = W * R, pad = L.pad, JP = L.JP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;  // warp = strip, lane = layer chunk
  const int LT = 32 * K;
  float* sB = smem;                                   // rotated molecule 2: nB*nb*3
  float* Ct = sB + ((nB * nb * 3 + 3) & ~3);          // (nBe+2) x WR, column jc in 0..nBe+1
  float* ring = Ct + (size_t)(nBe + 2) * WR;          // 2 x W x LT boundary rows
  const float INF = __int_as_float(0x7f800000);
  const int nsurv = *surv_count;

  // per-layer constants of this lane
  float initk[K];
  int jlo[K], jhi[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int n = lane * K + k;
    initk[k] = (n == 0) ? 0.0f : INF;
    // columns where layer n has live cells in this strip (rows i_lo..i_hi of the strip)
    const int i_lo = w * R + 1 - pad, i_hi = w * R + R - pad;
    const int imax = min(m, n + P.wA - 1);
    const bool live = n >= 1 && n <= P.n_max && n <= i_hi && i_lo <= imax;
    jlo[k] = live ? n : 0x3fffffff;
    jhi[k] = live ? min(nB, n + P.wB - 1) : -1;
  }

  for (int sv = blockIdx.x; sv < nsurv; sv += gridDim.x) {
    const float* T = orient + 12 * (size_t)surv[sv];
    __syncthreads();
    for (int k = threadIdx.x; k < nB * nb; k += blockDim.x) {
      float x, y, z;
      mr_xform(T, B[3 * k], B[3 * k + 1], B[3 * k + 2], x, y, z);
      sB[3 * k] = x; sB[3 * k + 1] = y; sB[3 * k + 2] = z;
    }
    for (int k = threadIdx.x; k < 2 * W * LT; k += blockDim.x) ring[k] = (k % LT == 0) ? 0.0f : INF;  // layer 0 (lane 0, k 0) is the constant-0 base layer
    __syncthreads();
    // distance matrix, transposed and padded (Dyn3d::FillLookupTable, dyn3d.cc:664-701)
    for (int idx = threadIdx.x; idx < (nBe + 2) * WR; idx += blockDim.x) {
      const int jc = idx / WR, rho = idx - jc * WR;
      const int i = rho - pad + 1;
      float c = INF;
      if (jc >= 1 && jc <= nB && i >= 1) {
        c = 0.0f;
        for (int q = 0; q < nb; q++) {
          const float* a = A + 3 * ((i - 1) * nb + q);
          const float* b = sB + 3 * ((jc - 1) * nb + q);
          c += mr_d2(a[0], a[1], a[2], b[0], b[1], b[2]);
        }
        if (c > P.dp_cutoff) c = MR_UB;
      }
      Ct[idx] = c;
    }
    __syncthreads();

    float st[K][R], bprev[K];
    uint32_t acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      bprev[k] = initk[k];
      acc[k] = 0;
#pragma unroll
      for (int r = 0; r < R; r++) st[k][r] = initk[k];
    }
    uint32_t* tbo = tb + (size_t)sv * L.words(nB);
    const int S = nBe + W - 1;
    for (int s = 1; s <= S; s++) {
      const int j = s - w;
      const int jc = min(max(j, 0), nBe + 1);
      // C(i, j) for the R rows of this strip
      float c[R];
      {
        const float4* cp = reinterpret_cast<const float4*>(Ct + (size_t)jc * WR + w * R);
#pragma unroll
        for (int q = 0; q < R / 4; q++) {
          const float4 v = cp[q];
          c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
        }
      }
      // A(n, i_lo - 1, j) for my layers: boundary row written by the strip above last step
      float bin[K];
      if (w == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) bin[k] = initk[k];
      } else {
        const float* rp = ring + ((size_t)((s - 1) & 1) * W + (w - 1)) * LT;
#pragma unroll
        for (int k = 0; k < K; k++) bin[k] = rp[k * 32 + lane];
      }
      // A(n0-1, i-1, j-1) for my lowest layer: neighbour lane's highest layer, previous column
      float dsh[R];
      dsh[0] = __shfl_up_sync(0xffffffffu, bprev[K - 1], 1);
#pragma unroll
      for (int r = 1; r < R; r++) dsh[r] = __shfl_up_sync(0xffffffffu, st[K - 1][r - 1], 1);

#pragma unroll
      for (int k = K - 1; k >= 0; k--) {
        float up = bin[k];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float dg = (k > 0) ? ((r > 0) ? st[k - 1][r - 1] : bprev[k - 1]) : dsh[r];
          const float d = dg + c[r];
          const float left = st[k][r];
          const float best = fminf(fminf(d, up), left);
          const float s1 = __fsub_rn(d, up);       // sign set <=> NOT (up <= d)
          const float s2 = __fsub_rn(best, left);  // sign set <=> NOT (left <= min(d, up))
          acc[k] = __funnelshift_l(__float_as_uint(s1), acc[k], 1);
          acc[k] = __funnelshift_l(__float_as_uint(s2), acc[k], 1);
          st[k][r] = best;
          up = best;
        }
      }
      // hand the bottom row of this strip to the strip below; remember this column's top boundary
      {
        float* wp = ring + ((size_t)(s & 1) * W + w) * LT;
#pragma unroll
        for (int k = 0; k < K; k++) {
          wp[k * 32 + lane] = st[k][R - 1];
          bprev[k] = bin[k];
        }
      }
      // traceback words: one per layer every second column
      if (j >= 1 && j <= nBe && (j & 1) == 0) {
        uint32_t* tw = tbo + ((size_t)w * JP + ((j - 1) >> 1)) * LT + lane;
#pragma unroll
        for (int k = 0; k < K; k++)
          if (j >= jlo[k] && j - 1 <= jhi[k]) tw[k * 32] = acc[k];
      }
      // cost(n) = A(n, m, nB) for n >= n_min (row m is the last row of the last strip)
      if (j == nB && w == W - 1) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          const int n = lane * K + k;
          if (n >= P.n_min && n <= P.n_max) cost_out[(size_t)sv * L.NP + (n - 1)] = st[k][R - 1];
        }
      }
      // n < n_min: the reference reports the corner of the layer's window,
      // A(n, n+wA-1, n+wB-1) (dyn3d.cc:385, dyn3d.h:407-408): one layer per column
      if (P.n_min > 1) {
        const int nq = j - P.wB + 1;
        if (nq >= 1 && nq < P.n_min && nq <= P.n_max && nq / K == lane) {
          const int rho = nq + P.wA - 2 + pad;
          if (rho / R == w) {
            const int kq = nq - lane * K, rq = rho - w * R;
            float v = 0.0f;
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
              for (int r = 0; r < R; r++)
                if (k == kq && r == rq) v = st[k][r];
            cost_out[(size_t)sv * L.NP + (nq - 1)] = v;
          }
        }
      }
      __syncthreads();
    }
  }
}

static inline void mr_plan_dp_tiled(MrTiledPlan* p, const MrDevParams& P, const cudaDeviceProp& prop, int NP) {
  p->ok = false;
  const int R = 8;
  const int K = (P.n_max + 1 + 31) / 32;
  if (K > 6) { p->why = "n_max + 1 > 192 layers"; return; }
  const int W = (P.m + R - 1) / R;
  if (W > 32) { p->why = "more than 256 rows"; return; }
  const int nBe = (P.nB + 1) & ~1;
  const size_t smem = 4 * (size_t)(((P.nB * P.nb * 3 + 3) & ~3) + (size_t)(nBe + 2) * W * R + 2 * (size_t)W * 32 * K);
  if (smem > prop.sharedMemPerBlockOptin) { p->why = "distance matrix does not fit shared memory"; return; }
  p->K = K; p->R = R; p->W = W; p->nBe = nBe;
  p->threads = 32 * W;
  p->smem = smem;
  p->ctas_per_sm = 1;
  p->grid_cap = prop.multiProcessorCount * p->ctas_per_sm;
  p->layout.K = K; p->layout.R = R; p->layout.W = W; p->layout.JP = nBe / 2;
  p->layout.pad = W * R - P.m; p->layout.NP = NP;
  p->tb_words = p->layout.words(P.nB);
  p->ok = true;
  p->why = "";
}

static inline void mr_free_dp_tiled(MrTiledPlan*) {}

template <int K, int MAXW>
static inline int mr_launch_dp_tiled_kw(const MrTiledPlan& p, const MrDevParams& P, const float* A,
                                       const float* B, const float* orient, const int* surv,
                                       const int* surv_count, uint32_t* tb, float* cost, int count,
                                       cudaStream_t stream) {
  auto kern = k_dp_tiled<K, 8, MAXW>;
  static size_t attr_set = 0;
  if (attr_set < p.smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) return 1;
    attr_set = p.smem;
  }
  const int grid = std::min(count, p.grid_cap);
  kern<<<grid, p.threads, p.smem, stream>>>(P, p.layout, p.nBe, A, B, orient, surv, surv_count, tb, cost);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <int K>
static inline int mr_launch_dp_tiled_k(const MrTiledPlan& p, const MrDevParams& P, const float* A,
                                       const float* B, const float* orient, const int* surv,
                                       const int* surv_count, uint32_t* tb, float* cost, int count,
                                       cudaStream_t stream) {
  if (p.W <= 12) return mr_launch_dp_tiled_kw<K, 12>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
  if (p.W <= 18) return mr_launch_dp_tiled_kw<K, 18>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
  if (p.W <= 24) return mr_launch_dp_tiled_kw<K, 24>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
  return mr_launch_dp_tiled_kw<K, 32>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
}

static inline int mr_launch_dp_tiled(const MrTiledPlan& p, const MrDevParams& P, const float* A,
                                     const float* B, const float* orient, const int* surv,
                                     const int* surv_count, uint32_t* tb, float* cost, int count,
                                     cudaStream_t stream) {
  switch (p.K) {
    case 1: return mr_launch_dp_tiled_k<1>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
    case 2: return mr_launch_dp_tiled_k<2>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
    case 3: return mr_launch_dp_tiled_k<3>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
    case 4: return mr_launch_dp_tiled_k<4>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
    case 5: return mr_launch_dp_tiled_k<5>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
    case 6: return mr_launch_dp_tiled_k<6>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream);
  }
  return 1;
}
