// mr_kernels.cuh -- the sm_100a kernels of the per-orientation search.
//
//   k_fragment_fit     K0  orientation generation (lib/or_gen/or_gen.cc:507-544)
//   k_nw_filter        K1+K2 rotate + banded NW filter (apply_trans.cc:72-108,
//                      orientation_filter.cc:88-130, dynNW.cc:243-531)
//   k_dp_columns       K1+K3+K4 general 3-D DP: rotate, distance build, all layers,
//                      2-bit traceback (dyn3d.cc:135-387, 664-701)
//   k_traceback_kabsch K5+K6 per-(orientation,size) traceback + superposition
//                      (dyn3d.cc:391-431, pair_alignment.cc:714-820, superimpose.cc:17-287)
//   k_select           K7 early-exit prefix + per-size arg-min (search_or_dyn3d.cc:1714-1835)
#pragma once
#include "mr_device.cuh"

// ---------------------------------------------------------------------------
// Traceback storage of the general kernel: 2 bits per cell {0 match, 1 skip-i, 2 skip-j}
// (codes of lib/dyn/dyn3d.cc:53-58), 16 consecutive rows i per 32-bit word,
// word index = ((j-1) * IW + ((i-1) >> 4)) * NP + (n-1): consecutive layers n are
// consecutive words, so the layer-per-thread kernel writes coalesced.
// ---------------------------------------------------------------------------
struct TbColumns {
  int IW, NP;
  __host__ __device__ size_t words(int nB) const { return (size_t)nB * IW * NP; }
  __device__ __forceinline__ size_t word(int n, int i, int j) const {
    return ((size_t)(j - 1) * IW + ((i - 1) >> 4)) * NP + (n - 1);
  }
  __device__ __forceinline__ int shift(int i) const { return 2 * ((i - 1) & 15); }
  __host__ __device__ int table_ints() const { return 0; }
  __device__ __forceinline__ void load_table(int*, int, int) const {}
  // Dyn3d::ExportSolution (dyn3d.cc:391-431): from (n, m, nB) follow the codes; emit(k, i0, j0)
  // for the k-th (0-based) matched pair with 0-based residue indices
  template <bool GROUP, class Emit>
  __device__ __forceinline__ void walk(const int*, const uint32_t* __restrict__ tbo, int n, int m, int nB, Emit emit) const {
    int cn = n, ci = m, cj = nB;
    while (cn > 0) {
      MR_ASSERT(ci >= cn && cj >= cn && ci <= m && cj <= nB);  // a path never leaves the live cells
      const uint32_t code = (tbo[word(cn, ci, cj)] >> shift(ci)) & 3u;  // 0 match, 1 skip-i, 2 skip-j
      if (code == 0) emit(cn - 1, ci - 1, cj - 1);
      cn -= (code == 0);
      ci -= (code != 2);
      cj -= (code != 1);
    }
  }
  __device__ __forceinline__ size_t addr(const int*, int n, int i, int j) const { return word(n, i, j); }
  __device__ __forceinline__ uint32_t decode(uint32_t w, int i, int /*j*/) const { return (w >> shift(i)) & 3u; }
};

// ---------------------------------------------------------------------------
// K0: one thread per fragment pair (I,J): least-squares fit of L*nb points of B onto A.
// ---------------------------------------------------------------------------
__global__ void k_fragment_fit(const float* __restrict__ A, const float* __restrict__ B, int nb,
                               int L, const int2* __restrict__ ij, int64_t count,
                               float* __restrict__ out12, float* __restrict__ out_rmsd) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const float* f = A + 3 * (size_t)ij[t].x * nb;
  const float* mv = B + 3 * (size_t)ij[t].y * nb;
  float T[12];
  float r = mr_superimpose(L * nb, [&](int a, float* pf, float* pm) {
    pf[0] = f[3 * a]; pf[1] = f[3 * a + 1]; pf[2] = f[3 * a + 2];
    pm[0] = mv[3 * a]; pm[1] = mv[3 * a + 1]; pm[2] = mv[3 * a + 2];
  }, T);
#pragma unroll
  for (int k = 0; k < 12; k++) out12[12 * t + k] = T[k];
  out_rmsd[t] = r;
}

// generic batched superposition (mr_superimpose_batch)
__global__ void k_superimpose_batch(int batch, int np, const float* __restrict__ fixed,
                                    const float* __restrict__ moveable, float* __restrict__ Tout,
                                    float* __restrict__ rmsd) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch) return;
  const float* f = fixed + (size_t)t * np * 3;
  const float* mv = moveable + (size_t)t * np * 3;
  float T[12];
  float r = mr_superimpose(np, [&](int a, float* pf, float* pm) {
    pf[0] = f[3 * a]; pf[1] = f[3 * a + 1]; pf[2] = f[3 * a + 2];
    pm[0] = mv[3 * a]; pm[1] = mv[3 * a + 1]; pm[2] = mv[3 * a + 2];
  }, T);
#pragma unroll
  for (int k = 0; k < 12; k++) Tout[12 * (size_t)t + k] = T[k];
  rmsd[t] = r;
}

// ---------------------------------------------------------------------------
// K1+K2: NW orientation filter, one thread per orientation.  All threads of a warp walk
// the same (i,j) sequence, so A_i / B_j loads are warp-uniform broadcasts; the one-row
// table lives in a [j][thread] scratch array (coalesced).  The number of matches on the
// traceback is carried forward as a counter (border cells reset it: dynNW.h GetPrevCell
// PREV_PTR_UNINITIALIZED jumps to (0,0)), so no traceback table is stored.
// Writes nw_matches[o] and appends survivors (matches >= of_min) to surv[] (unordered).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_nw_filter(MrDevParams P, const float* __restrict__ A, const float* __restrict__ B,
            const float* __restrict__ orient, int count, float* __restrict__ row_cost,
            unsigned short* __restrict__ row_cnt, int* __restrict__ nw_matches,
            int* __restrict__ surv, int* __restrict__ surv_count, int o_base) {
  // orientations o_base .. o_base + count - 1 of the list (the row scratch holds one batch)
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  if (tid >= count) return;
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = orient[12 * (size_t)(o_base + tid) + k];
  const int m = P.m, nB = P.nB, nb = P.nb, w0 = P.w0, w1 = P.w1;
  const float G = P.nw_gap, cutoff = P.nw_cutoff;
  float left = 0.f;
  int leftc = 0;
  for (int i = 1; i <= m; i++) {
    const int jlo = (i <= w0) ? 1 : i - w0 + 1;
    const int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
    // band of the previous row (row 0 has none)
    const int pjlo = (i - 1 <= w0) ? 1 : i - 1 - w0 + 1;
    const int pjhi = (i - 1 + w1 <= nB) ? i - 1 + w1 - 1 : nB;
    const bool prow = (i - 1) >= 1;
    left = (float)(i + jlo - 1) * G;  // border cell (i, jlo-1), dynNW.cc:200-216
    leftc = 0;
    float diag;
    int diagc;
    {
      const int j = jlo - 1;
      if (prow && j >= pjlo && j <= pjhi) {
        diag = row_cost[(size_t)j * nthreads + tid];
        diagc = row_cnt[(size_t)j * nthreads + tid];
      } else {
        diag = (float)(i - 1 + j) * G;
        diagc = 0;
      }
    }
    for (int j = jlo; j <= jhi; j++) {
      float up;
      int upc;
      if (prow && j >= pjlo && j <= pjhi) {
        up = row_cost[(size_t)j * nthreads + tid];
        upc = row_cnt[(size_t)j * nthreads + tid];
      } else {
        up = (float)(i - 1 + j) * G;  // border cell, dynNW.cc:218-234
        upc = 0;
      }
      float c = 0.0f;
      for (int q = 0; q < nb; q++) {
        const float* a = A + 3 * ((size_t)(i - 1) * nb + q);
        const float* b = B + 3 * ((size_t)(j - 1) * nb + q);
        float bx, by, bz;
        mr_xform(T, b[0], b[1], b[2], bx, by, bz);
        c += mr_d2(a[0], a[1], a[2], bx, by, bz);
      }
      if (c > cutoff) c = MR_UB;
      const float c_match = diag + c, c_i = up + G, c_j = left + G;
      float best = c_i;  // tie order of dynNW.cc:350-363: skip-i, then skip-j (<), then match (<)
      int bc = upc;
      if (c_j < best) { best = c_j; bc = leftc; }
      if (c_match < best) { best = c_match; bc = diagc + 1; }
      row_cost[(size_t)j * nthreads + tid] = best;
      row_cnt[(size_t)j * nthreads + tid] = (unsigned short)bc;
      diag = up; diagc = upc;
      left = best; leftc = bc;
    }
  }
  nw_matches[o_base + tid] = leftc;
  if (leftc >= P.of_min) {
    int slot = atomicAdd(surv_count, 1);
    surv[slot] = o_base + tid;
  }
}

// ---------------------------------------------------------------------------
// K1+K2, warp-per-orientation variant (nB <= 32*CPL): lane l owns columns l*CPL+1..(l+1)*CPL
// of the NW table and walks the rows as a wavefront (lane l is at row t-l at time t); the
// value to its left arrives by shuffle from lane l-1, everything else stays in registers.
// The rotated coordinates of the lane's columns are computed once.  Cells outside the band
// take the border formula (i+j)*G with count 0 -- exactly the border cells the reference
// initialises (dynNW.cc:200-234) are the out-of-band cells that ever get read.
// ---------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(128)
k_nw_filter_warp(MrDevParams P, const float* __restrict__ A, const float* __restrict__ B,
                 const float* __restrict__ orient, int count, int* __restrict__ nw_matches,
                 int* __restrict__ surv, int* __restrict__ surv_count) {
  extern __shared__ float nw_sA[];  // m*3 (one atom per residue)
  const int m = P.m, nB = P.nB, w0 = P.w0, w1 = P.w1;
  for (int k = threadIdx.x; k < m * 3; k += blockDim.x) nw_sA[k] = A[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= count) return;
  const float G = P.nw_gap, cutoff = P.nw_cutoff;
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = orient[12 * (size_t)o + k];
  const int j0 = lane * CPL;  // my columns are j0+1 .. j0+CPL
  float bx[CPL], by[CPL], bz[CPL], up[CPL];
  int upc[CPL];
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    const int j = j0 + 1 + c;
    bx[c] = by[c] = bz[c] = 0.0f;
    if (j <= nB) mr_xform(T, B[3 * (j - 1)], B[3 * (j - 1) + 1], B[3 * (j - 1) + 2], bx[c], by[c], bz[c]);
    up[c] = (float)j * G;  // row 0
    upc[c] = 0;
  }
  float last = 0.0f, diag_in = (float)j0 * G;  // my last column at my current row; A[i-1][j0]
  int lastc = 0, diag_inc = 0;
  for (int t = 1; t <= m + 31; t++) {
    // value / count to my left at the row I am about to process: lane-1 finished it last step
    float left = __shfl_up_sync(0xffffffffu, last, 1);
    int leftc = __shfl_up_sync(0xffffffffu, lastc, 1);
    const int i = t - lane;
    if (lane == 0) { left = (float)i * G; leftc = 0; }  // column 0
    if (i >= 1 && i <= m) {
      const int jlo = (i <= w0) ? 1 : i - w0 + 1;
      const int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
      const float ax = nw_sA[3 * (i - 1)], ay = nw_sA[3 * (i - 1) + 1], az = nw_sA[3 * (i - 1) + 2];
      float diag = diag_in;
      int diagc = diag_inc;
      diag_in = left;  // A[i][j0] is the diagonal input of my first column at the next row
      diag_inc = leftc;
      // rows whose band covers all of my columns skip the per-cell border test
      const bool inside = (j0 + 1 >= jlo) && (j0 + CPL <= jhi);
#pragma unroll
      for (int c = 0; c < CPL; c++) {
        const int j = j0 + 1 + c;
        float cc = mr_d2(ax, ay, az, bx[c], by[c], bz[c]);
        if (cc > cutoff) cc = MR_UB;
        const float c_match = diag + cc, c_i = up[c] + G, c_j = left + G;
        float best = c_i;  // tie order of dynNW.cc:350-363: skip-i, then skip-j (<), then match (<)
        int bc = upc[c];
        if (c_j < best) { best = c_j; bc = leftc; }
        if (c_match < best) { best = c_match; bc = diagc + 1; }
        if (!inside && (j < jlo || j > jhi)) { best = (float)(i + j) * G; bc = 0; }
        diag = up[c]; diagc = upc[c];
        up[c] = best; upc[c] = bc;
        left = best; leftc = bc;
      }
      last = left;
      lastc = leftc;
    }
  }
  // the lane that owns column nB holds the count at (m, nB): its row-m value of that column
  const int owner = (nB - 1) / CPL;
  int result = 0;
#pragma unroll
  for (int c = 0; c < CPL; c++)
    if (c == (nB - 1) - owner * CPL) result = upc[c];
  result = __shfl_sync(0xffffffffu, result, owner);
  if (lane == 0) {
    nw_matches[o] = result;
    if (result >= P.of_min) surv[atomicAdd(surv_count, 1)] = o;
  }
}

// ---------------------------------------------------------------------------
// K1+K2, CTA-per-orientation variant for long chains (nB > 256): the warps of the CTA split
// the columns (warp v owns columns 32*CPL*v+1 ..), each warp runs the register wavefront of
// k_nw_filter_warp, and the last lane of warp v-1 hands A(i, last column) and its match
// count to lane 0 of warp v through a 128-row ring in shared memory.  Time is cut into
// blocks of 32 steps with one __syncthreads per block; warp v runs two blocks behind warp
// v-1 (row i leaves warp v-1 at its step i+31 and enters warp v at step i), so every value
// is written at least one barrier before it is read and read before its slot is reused.
// ---------------------------------------------------------------------------
template <int CPL, int MAXT>
__global__ void __launch_bounds__(MAXT)
k_nw_filter_cta(MrDevParams P, const float* __restrict__ A, const float* __restrict__ B,
                const float* __restrict__ orient, int count, int* __restrict__ nw_matches,
                int* __restrict__ surv, int* __restrict__ surv_count) {
  extern __shared__ float nw_sA[];  // m*3 floats, then the rings
  const int m = P.m, nB = P.nB, w0 = P.w0, w1 = P.w1;
  float2* ring = reinterpret_cast<float2*>(nw_sA + ((m * 3 + 3) & ~3));  // [warps-1][128] {cost, count}
  for (int k = threadIdx.x; k < m * 3; k += blockDim.x) nw_sA[k] = A[k];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float G = P.nw_gap, cutoff = P.nw_cutoff;
  const int nblk_local = (m + 31 + 31) / 32;
  const int nblk = nblk_local + 2 * (nwarps - 1);
  const int j0 = (warp * 32 + lane) * CPL;  // my columns are j0+1 .. j0+CPL
  for (int o = blockIdx.x; o < count; o += gridDim.x) {
    __syncthreads();
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = orient[12 * (size_t)o + k];
    float bx[CPL], by[CPL], bz[CPL], up[CPL];
    int upc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      const int j = j0 + 1 + c;
      bx[c] = by[c] = bz[c] = 0.0f;
      if (j <= nB) mr_xform(T, B[3 * (j - 1)], B[3 * (j - 1) + 1], B[3 * (j - 1) + 2], bx[c], by[c], bz[c]);
      up[c] = (float)j * G;  // row 0
      upc[c] = 0;
    }
    float last = 0.0f, diag_in = (float)j0 * G;
    int lastc = 0, diag_inc = 0;
    for (int g = 0; g < nblk; g++) {
      const int beta = g - 2 * warp;
      if (beta >= 0 && beta < nblk_local) {
        for (int t = 32 * beta + 1; t <= 32 * beta + 32; t++) {
          float left = __shfl_up_sync(0xffffffffu, last, 1);
          int leftc = __shfl_up_sync(0xffffffffu, lastc, 1);
          const int i = t - lane;
          const bool live = i >= 1 && i <= m;
          if (lane == 0) {
            if (warp == 0) { left = (float)i * G; leftc = 0; }  // column 0
            else if (live) {
              const float2 v = ring[(warp - 1) * 128 + (i & 127)];
              left = v.x; leftc = __float_as_int(v.y);
            }
          }
          if (live) {
            const int jlo = (i <= w0) ? 1 : i - w0 + 1;
            const int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
            const float ax = nw_sA[3 * (i - 1)], ay = nw_sA[3 * (i - 1) + 1], az = nw_sA[3 * (i - 1) + 2];
            float diag = diag_in;
            int diagc = diag_inc;
            diag_in = left;
            diag_inc = leftc;
            const bool inside = (j0 + 1 >= jlo) && (j0 + CPL <= jhi);
#pragma unroll
            for (int c = 0; c < CPL; c++) {
              const int j = j0 + 1 + c;
              float cc = mr_d2(ax, ay, az, bx[c], by[c], bz[c]);
              if (cc > cutoff) cc = MR_UB;
              const float c_match = diag + cc, c_i = up[c] + G, c_j = left + G;
              float best = c_i;  // tie order of dynNW.cc:350-363: skip-i, then skip-j (<), then match (<)
              int bc = upc[c];
              if (c_j < best) { best = c_j; bc = leftc; }
              if (c_match < best) { best = c_match; bc = diagc + 1; }
              if (!inside && (j < jlo || j > jhi)) { best = (float)(i + j) * G; bc = 0; }
              diag = up[c]; diagc = upc[c];
              up[c] = best; upc[c] = bc;
              left = best; leftc = bc;
            }
            last = left;
            lastc = leftc;
            if (lane == 31 && warp + 1 < nwarps) ring[warp * 128 + (i & 127)] = make_float2(last, __int_as_float(lastc));
          }
        }
      }
      __syncthreads();
    }
    // the thread that owns column nB holds the count at (m, nB): its row-m value of that column
    if (warp * 32 + lane == (nB - 1) / CPL) {
      int result = 0;
#pragma unroll
      for (int c = 0; c < CPL; c++)
        if (c == (nB - 1) - ((nB - 1) / CPL) * CPL) result = upc[c];
      nw_matches[o] = result;
      if (result >= P.of_min) surv[atomicAdd(surv_count, 1)] = o;
    }
  }
}

// ---------------------------------------------------------------------------
// K1+K2, decision-only form of the NW filter (FirstPass needs "NumMatches >= of_min", not the
// number itself, orientation_filter.cc:129).  The optimal cost decides it exactly:
//   along the traceback from (m, nB) to the border cell (ib, jb) it starts from
//   (value fl((ib+jb) G), count 0; dynNW.cc:200-234, 484-498) there are k match steps and
//   g = m + nB - ib - jb - 2k gap steps, every matched C is <= cutoff (a clipped 1e26 never wins
//   against the all-gap path), so in exact arithmetic
//        (m + nB - 2k) G  <=  cost  <=  (m + nB - 2k) G + k cutoff
//   and the float table differs from that by at most E = sum over the <= m+nB+1 roundings on the
//   path of half an ulp of the running value (<= s G at i + j = s).  With
//   2G - of_min cutoff - 2E > 0 the cost intervals of "k >= of_min" and "k <= of_min - 1" do not
//   touch, and one comparison with a threshold between them reproduces the reference's decision
//   (the threshold and the margin are computed in mr_create; without a margin the counting
//   kernels above run).  The cost table is the reference's, bit for bit (same float operations;
//   a min does not depend on the tie order), but nothing else is carried: no match counts, no
//   compare-and-select pairs -- 13 instead of ~35 instructions per cell.
//   (best + G is formed once per cell for both of its uses.)
// Same thread mapping as k_nw_filter_warp / k_nw_filter_cta.
// ---------------------------------------------------------------------------
// one row of a lane's CPL columns.  BAND: with the per-cell band test (cells outside the band take
// the border value (i + j) G); without it when every lane of the warp sits fully inside the band
// at its row (a warp-uniform decision: __all_sync in the callers) -- half the instructions.
template <int CPL, bool BAND>
__device__ __forceinline__ float mr_nw_cost_row(float (&up)[CPL], float (&upg)[CPL], const float (&bx)[CPL],
                                                const float (&by)[CPL], const float (&bz)[CPL], float ax, float ay, float az,
                                                float left, float diag, float G, float cutoff, int i, int j0, int jlo, int jhi) {
  float leftg = left + G;
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    const int j = j0 + 1 + c;
    float cc = mr_d2(ax, ay, az, bx[c], by[c], bz[c]);
    if (cc > cutoff) cc = MR_UB;
    float best = fminf(fminf(upg[c], leftg), diag + cc);
    if (BAND && (j < jlo || j > jhi)) best = (float)(i + j) * G;
    diag = up[c];
    up[c] = best;
    leftg = best + G;
    upg[c] = leftg;
    left = best;
  }
  return left;
}

// SW lanes per orientation (32, 16 or 8): short chains put 2 or 4 orientations into a warp, each
// lane then owns up to 16 columns -- the per-row work of a lane (shuffle, band bounds, the row's
// coordinates) is spread over more cells and the wavefront fills and drains in SW - 1 steps
// instead of 31.
template <int CPL, int SW = 32>
__global__ void __launch_bounds__(128)
k_nw_cost_warp(MrDevParams P, float thr, const float* __restrict__ A, const float* __restrict__ B,
               const float* __restrict__ orient, int count, int* __restrict__ surv,
               int* __restrict__ surv_count) {
  extern __shared__ float nw_sA[];  // m*3 (one atom per residue)
  const int m = P.m, nB = P.nB, w0 = P.w0, w1 = P.w1;
  for (int k = threadIdx.x; k < m * 3; k += blockDim.x) nw_sA[k] = A[k];
  __syncthreads();
  static_assert(SW == 32 || SW == 16 || SW == 8, "lanes per orientation");
  const int lane = threadIdx.x & (SW - 1);            // lane within the orientation's group
  const int sub = (threadIdx.x & 31) / SW;
  const int o_first = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / SW);
  if (o_first >= count) return;                        // the whole warp is beyond the list
  const bool valid = o_first + sub < count;            // a group beyond the list runs along (shuffles, votes) on the last orientation
  const int o = valid ? o_first + sub : count - 1;
  const float G = P.nw_gap, cutoff = P.nw_cutoff;
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = orient[12 * (size_t)o + k];
  const int j0 = lane * CPL;  // my columns are j0+1 .. j0+CPL
  float bx[CPL], by[CPL], bz[CPL], up[CPL], upg[CPL];  // up = A(i-1, j), upg = up + G
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    const int j = j0 + 1 + c;
    bx[c] = by[c] = bz[c] = 0.0f;
    if (j <= nB) mr_xform(T, B[3 * (j - 1)], B[3 * (j - 1) + 1], B[3 * (j - 1) + 2], bx[c], by[c], bz[c]);
    up[c] = (float)j * G;  // row 0
    upg[c] = up[c] + G;
  }
  float last = 0.0f, diag_in = (float)j0 * G;  // my last column at my current row; A[i-1][j0]
  for (int t = 1; t <= m + SW - 1; t++) {
    float left = __shfl_up_sync(0xffffffffu, last, 1, SW);
    const int i = t - lane;
    if (lane == 0) left = (float)i * G;  // column 0
    const bool live = i >= 1 && i <= m;
    const int jlo = (i <= w0) ? 1 : i - w0 + 1;
    const int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
    // fully inside the band (columns beyond nB do not count: nothing in the band ever reads them)
    const bool inside = (j0 + 1 >= jlo) && (min(j0 + CPL, nB) <= jhi);
    const bool no_edge = __all_sync(0xffffffffu, !live || inside || j0 + 1 > nB);
    if (live) {
      const float ax = nw_sA[3 * (i - 1)], ay = nw_sA[3 * (i - 1) + 1], az = nw_sA[3 * (i - 1) + 2];
      const float diag = diag_in;
      diag_in = left;  // A[i][j0] is the diagonal input of my first column at the next row
      if (no_edge) last = mr_nw_cost_row<CPL, false>(up, upg, bx, by, bz, ax, ay, az, left, diag, G, cutoff, i, j0, jlo, jhi);
      else last = mr_nw_cost_row<CPL, true>(up, upg, bx, by, bz, ax, ay, az, left, diag, G, cutoff, i, j0, jlo, jhi);
    }
  }
  // the lane that owns column nB holds A(m, nB): its row-m value of that column
  const int owner = (nB - 1) / CPL;
  float result = 0.0f;
#pragma unroll
  for (int c = 0; c < CPL; c++)
    if (c == (nB - 1) - owner * CPL) result = up[c];
  result = __shfl_sync(0xffffffffu, result, owner, SW);
  if (lane == 0 && valid && result <= thr) surv[atomicAdd(surv_count, 1)] = o;
}

template <int CPL, int MAXT>
__global__ void __launch_bounds__(MAXT)
k_nw_cost_cta(MrDevParams P, float thr, const float* __restrict__ A, const float* __restrict__ B,
              const float* __restrict__ orient, int count, int* __restrict__ surv,
              int* __restrict__ surv_count) {
  extern __shared__ float nw_sA[];  // m*3 floats, then the rings
  const int m = P.m, nB = P.nB, w0 = P.w0, w1 = P.w1;
  float* ring = nw_sA + ((m * 3 + 3) & ~3);  // [warps-1][128] cost of the warp's last column
  for (int k = threadIdx.x; k < m * 3; k += blockDim.x) nw_sA[k] = A[k];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float G = P.nw_gap, cutoff = P.nw_cutoff;
  const int nblk_local = (m + 31 + 31) / 32;
  const int nblk = nblk_local + 2 * (nwarps - 1);
  const int j0 = (warp * 32 + lane) * CPL;  // my columns are j0+1 .. j0+CPL
  for (int o = blockIdx.x; o < count; o += gridDim.x) {
    __syncthreads();
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = orient[12 * (size_t)o + k];
    float bx[CPL], by[CPL], bz[CPL], up[CPL], upg[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      const int j = j0 + 1 + c;
      bx[c] = by[c] = bz[c] = 0.0f;
      if (j <= nB) mr_xform(T, B[3 * (j - 1)], B[3 * (j - 1) + 1], B[3 * (j - 1) + 2], bx[c], by[c], bz[c]);
      up[c] = (float)j * G;  // row 0
      upg[c] = up[c] + G;
    }
    float last = 0.0f, diag_in = (float)j0 * G;
    for (int g = 0; g < nblk; g++) {
      const int beta = g - 2 * warp;
      if (beta >= 0 && beta < nblk_local) {
        for (int t = 32 * beta + 1; t <= 32 * beta + 32; t++) {
          float left = __shfl_up_sync(0xffffffffu, last, 1);
          const int i = t - lane;
          const bool live = i >= 1 && i <= m;
          if (lane == 0) {
            if (warp == 0) left = (float)i * G;  // column 0
            else if (live) left = ring[(warp - 1) * 128 + (i & 127)];
          }
          const int jlo = (i <= w0) ? 1 : i - w0 + 1;
          const int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
          const bool inside = (j0 + 1 >= jlo) && (min(j0 + CPL, nB) <= jhi);
          const bool no_edge = __all_sync(0xffffffffu, !live || inside || j0 + 1 > nB);
          if (live) {
            const float ax = nw_sA[3 * (i - 1)], ay = nw_sA[3 * (i - 1) + 1], az = nw_sA[3 * (i - 1) + 2];
            const float diag = diag_in;
            diag_in = left;
            if (no_edge) last = mr_nw_cost_row<CPL, false>(up, upg, bx, by, bz, ax, ay, az, left, diag, G, cutoff, i, j0, jlo, jhi);
            else last = mr_nw_cost_row<CPL, true>(up, upg, bx, by, bz, ax, ay, az, left, diag, G, cutoff, i, j0, jlo, jhi);
            if (lane == 31 && warp + 1 < nwarps) ring[warp * 128 + (i & 127)] = last;
          }
        }
      }
      __syncthreads();
    }
    // the thread that owns column nB holds A(m, nB): its row-m value of that column
    if (warp * 32 + lane == (nB - 1) / CPL) {
      float result = 0.0f;
#pragma unroll
      for (int c = 0; c < CPL; c++)
        if (c == (nB - 1) - ((nB - 1) / CPL) * CPL) result = up[c];
      if (result <= thr) surv[atomicAdd(surv_count, 1)] = o;
    }
  }
}

__global__ void k_set_int(int* __restrict__ p, int v) { *p = v; }
__global__ void k_fill_f32(float* __restrict__ p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Refine batch set-up in one launch: identity survivor list of `count` items and zeroed
// per-size outputs (what k_iota + four memsets did)
__global__ void k_refine_prep(int* __restrict__ surv, int* __restrict__ surv_count, int count, int NP,
                              float* __restrict__ cost, float* __restrict__ ssd,
                              unsigned char* __restrict__ valid, int* __restrict__ n_last) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) *surv_count = count;
  if (t < count) { surv[t] = t; n_last[t] = 0; }
  if (t < count * NP) { cost[t] = 0.0f; ssd[t] = 0.0f; valid[t] = 0; }
}

// identity survivor list when the filter is off
__global__ void k_iota(int* __restrict__ surv, int* __restrict__ surv_count, int count) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) surv[t] = t;
  if (t == 0) *surv_count = count;
}

// ---------------------------------------------------------------------------
// K1+K3+K4, general kernel: one CTA per orientation, one thread per layer n, sweeping
// the columns j of molecule 2.  Within a column a thread walks its rows i = n..i_max(n)
// keeping A(n,i-1,j) in a register; A(n-1,i-1,j-1) and A(n,i,j-1) come from the previous
// column's buffer colA[i][n-1] (shared memory when it fits, else a per-CTA global work
// space), the results go to colB; one __syncthreads per column.
// Recurrence and tie order of lib/dyn/dyn3d.cc:283-382: start with match, "<=" skip-i,
// "<=" skip-j.  C(i,j) is rebuilt on the fly from A_i (shared) and the rotated B'_j.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_dp_columns(MrDevParams P, TbColumns L, const float* __restrict__ A, const float* __restrict__ B,
             const float* __restrict__ orient, const int* __restrict__ surv,
             const int* __restrict__ surv_count, uint32_t* __restrict__ tb,
             float* __restrict__ cost_out, float* __restrict__ gcol, int use_gcol) {
  extern __shared__ float smem[];
  const int m = P.m, nB = P.nB, nb = P.nb, NP = L.NP;
  const int n0 = threadIdx.x + 1;
  float* sA = smem;                       // m*nb*3
  float* sB = sA + (size_t)m * nb * 3;     // nB*nb*3 rotated
  float* col0;
  float* col1;
  const size_t colsz = (size_t)(m + 1) * NP;
  if (use_gcol) {
    col0 = gcol + (size_t)blockIdx.x * 2 * colsz;
    col1 = col0 + colsz;
  } else {
    col0 = sB + (size_t)nB * nb * 3;
    col1 = col0 + colsz;
  }
  for (int k = threadIdx.x; k < m * nb * 3; k += blockDim.x) sA[k] = A[k];
  const int nsurv = *surv_count;
  const float INF = __int_as_float(0x7f800000);
  for (int s = blockIdx.x; s < nsurv; s += gridDim.x) {
    const int o = surv[s];
    const float* T = orient + 12 * (size_t)o;
    __syncthreads();
    for (int k = threadIdx.x; k < nB * nb; k += blockDim.x) {
      float x, y, z;
      mr_xform(T, B[3 * k], B[3 * k + 1], B[3 * k + 2], x, y, z);
      sB[3 * k] = x; sB[3 * k + 1] = y; sB[3 * k + 2] = z;
    }
    __syncthreads();
    uint32_t* tbo = tb + (size_t)s * L.words(nB);
    float* colA = col0;
    float* colB = col1;
    for (int j = 1; j <= nB; j++) {
      // one layer per thread; more than blockDim layers (n_max > 1024): a thread takes
      // layers n0, n0 + blockDim, ... (all state of a layer lives in the column buffers)
      for (int n = n0; n <= P.n_max; n += blockDim.x) {
        const int imax = min(m, n + P.wA - 1), jmax = min(nB, n + P.wB - 1);
        if (j < n || j > jmax) continue;
        float up = INF;
        uint32_t acc = 0;
        float res = 0.f;
        for (int i = n; i <= imax; i++) {
          float c = 0.0f;
          for (int q = 0; q < nb; q++) {
            const float* a = sA + 3 * ((i - 1) * nb + q);
            const float* b = sB + 3 * ((j - 1) * nb + q);
            c += mr_d2(a[0], a[1], a[2], b[0], b[1], b[2]);
          }
          if (c > P.dp_cutoff) c = MR_UB;
          const float prev = (n > 1) ? colA[(size_t)(i - 1) * NP + (n - 2)] : 0.0f;
          const float d = (n > 1) ? prev + c : c;
          const float left = (j > n) ? colA[(size_t)i * NP + (n - 1)] : INF;
          float best = d;
          uint32_t code = 0;
          if (up <= best) { best = up; code = 1; }
          if (left <= best) { best = left; code = 2; }
          colB[(size_t)i * NP + (n - 1)] = best;
          up = best;
          res = best;
          acc |= code << L.shift(i);
          if (((i - 1) & 15) == 15 || i == imax) {
            tbo[L.word(n, i, j)] = acc;
            acc = 0;
          }
        }
        if (j == jmax) cost_out[(size_t)s * NP + (n - 1)] = res;
      }
      __syncthreads();
      float* t = colA; colA = colB; colB = t;
    }
  }
}

// ---------------------------------------------------------------------------
// K5+K6: one thread per (survivor s, size n): walk the traceback from (n, m, nB)
// (Dyn3d::ExportSolution, dyn3d.cc:391-431), then superimpose the n matched residue pairs
// taken from the ORIGINAL coordinates (pair_alignment.cc:754-804) and form
// ssd = SQR(rmsd)*n*nb (search_or_dyn3d.cc:587-588).  The pair list is staged in a
// per-thread scratch slice so the sums run in ascending pair order like the reference.
// The walk is branch-free (one dependent table word per step); the layout's lookup table
// and both coordinate sets sit in shared memory.
//   items == NULL: dense grid over (s, n in [n_lo, n_max]);  else explicit (s,n) list.
//   EMIT: also write the transform and the pair list.
#ifndef MR_TB_MINB
#define MR_TB_MINB 12
#endif
// The emit variant runs a few hundred threads whose result the host waits for (winners'
// payload, Refine batches): no occupancy target, so no register cap and no spills.
#ifndef MR_TB_EMIT_MINB
#define MR_TB_EMIT_MINB 4
#endif
//   EXCL (-alt-r, SearchOrDyn3d::MultiPass, search_or_dyn3d.cc:2086-2107): also count the
//         pairs this alignment shares with the stored alignment of the same size of each of
//         excl_ranks lower ranks (NumMatchesInBothAssumingMonotonicity,
//         lib/pair_alignment/pair_alignment.cc:2529-2555: both lists are strictly increasing in
//         i, so the count is the number of pairs k with partner[i_k] == j_k) and flag the
//         alignment when any count exceeds excl_max_common.  excl_partner: [rank][R][m] shorts,
//         j matched to residue i or -1; excl_out: [survivor][NP] bytes.
// ---------------------------------------------------------------------------
template <class Layout, bool EMIT, bool EXCL = false>
__global__ void __launch_bounds__(128, EMIT ? MR_TB_EMIT_MINB : MR_TB_MINB)
k_traceback_kabsch(MrDevParams P, Layout L, const float* __restrict__ A, const float* __restrict__ B,
                   const float* __restrict__ orient, const int* __restrict__ surv,
                   const int* __restrict__ surv_count, const uint32_t* __restrict__ tb,
                   const float* __restrict__ cost, const int2* __restrict__ items, int n_items,
                   int n_lo, short2* __restrict__ scratch, float* __restrict__ ssd_out,
                   float* __restrict__ T_out, short* __restrict__ pairs_out, int pairs_in_smem,
                   const short* __restrict__ excl_partner = nullptr, int excl_ranks = 0,
                   int excl_max_common = 0, unsigned char* __restrict__ excl_out = nullptr) {
  extern __shared__ int tb_smem[];
  const int nb = P.nb;
  int* tab = tb_smem;                                              // layout table
  float* sA = reinterpret_cast<float*>(tab + L.table_ints());      // m*nb*3
  float* sB = sA + P.m * nb * 3;                                   // nB*nb*3
  // emit mode, when it fits: a second copy of the pair lists in shared memory, [pair][thread].
  // A few hundred threads run alone here, and the two Kabsch passes would otherwise wait for
  // an L2 round trip per pair (stores do not allocate in L1): measured 0.14 -> see DESIGN.md
  short2* sp = (EMIT && pairs_in_smem) ? reinterpret_cast<short2*>(sB + P.nB * nb * 3) + threadIdx.x : nullptr;
  L.load_table(tab, threadIdx.x, blockDim.x);
  for (int k = threadIdx.x; k < P.m * nb * 3; k += blockDim.x) sA[k] = A[k];
  for (int k = threadIdx.x; k < P.nB * nb * 3; k += blockDim.x) sB[k] = B[k];
  __syncthreads();
  const int NP = L.NP;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int s, n;
  if (items) {
    if (tid >= n_items) return;
    s = items[tid].x;
    n = items[tid].y;
  } else {
    const int per = P.n_max - n_lo + 1;
    if (per <= 0) return;
    s = tid / per;
    n = n_lo + tid % per;
    if (s >= *surv_count) return;
  }
  const size_t slot = (size_t)s * NP + (n - 1);
  // warp-interleaved slice: entry k of lane l sits at [warp][k][l] (coalesced per warp)
  short2* my = scratch + (size_t)(tid >> 5) * 32 * P.N + (tid & 31);
  MR_ASSERT(n >= 1 && n <= P.N && s >= 0);
#define MY(k) my[(size_t)(k) * 32]
  const uint32_t* tbo = tb + (size_t)s * L.words(P.nB);
  L.template walk<EMIT>(tab, tbo, n, P.m, P.nB, [&](int k, int i0, int j0) {
    const short2 v = make_short2((short)i0, (short)j0);
    MY(k) = v;
    if (EMIT && sp) sp[(size_t)k * 128] = v;
  });
#define MYR(k) ((EMIT && sp) ? sp[(size_t)(k) * 128] : MY(k))
  if (EXCL) {
    bool too_similar = false;
    if (n >= P.n_min) {
      const int R = P.n_max - P.n_min + 1;
      for (int q = 0; q < excl_ranks; q++) {
        const short* partner = excl_partner + ((size_t)q * R + (n - P.n_min)) * P.m;
        int common = 0;
        for (int k = 0; k < n; k++) {
          const short2 pr = MYR(k);
          common += (partner[pr.x] == pr.y) ? 1 : 0;
        }
        too_similar = too_similar || common > excl_max_common;
      }
    }
    excl_out[slot] = too_similar ? 1 : 0;
  }
  float T[12];
  float ssd;
  if (P.refine_while_searching) {
    float rmsd;
    if (nb == 1) {
      rmsd = mr_superimpose(n, [&](int a, float* pf, float* pm) {
        const short2 pr = MYR(a);
        const float* f = sA + 3 * pr.x;
        const float* mv = sB + 3 * pr.y;
        pf[0] = f[0]; pf[1] = f[1]; pf[2] = f[2];
        pm[0] = mv[0]; pm[1] = mv[1]; pm[2] = mv[2];
      }, T);
    } else {
      rmsd = mr_superimpose(n * nb, [&](int a, float* pf, float* pm) {
        const int k = a / nb, q = a - k * nb;
        const short2 pr = MYR(k);
        const float* f = sA + 3 * (pr.x * nb + q);
        const float* mv = sB + 3 * (pr.y * nb + q);
        pf[0] = f[0]; pf[1] = f[1]; pf[2] = f[2];
        pm[0] = mv[0]; pm[1] = mv[1]; pm[2] = mv[2];
      }, T);
    }
    ssd = ((rmsd * rmsd) * (float)n) * (float)nb;
  } else {
    ssd = cost[slot];
    if (EMIT) {
      const float* To = orient + 12 * (size_t)surv[s];
#pragma unroll
      for (int k = 0; k < 12; k++) T[k] = To[k];
    }
  }
  ssd_out[slot] = ssd;
  if (EMIT) {
    const size_t e = items ? (size_t)tid : slot;
#pragma unroll
    for (int k = 0; k < 12; k++) T_out[12 * e + k] = T[k];
    // pairs_out == NULL: the caller reads the pair lists from the scratch slices themselves
    // (Refine batches: only a few of the items are ever consumed)
    if (pairs_out) {
      short* po = pairs_out + e * 2 * P.N;
      for (int k = 0; k < n; k++) {
        const short2 pr = MYR(k);
        po[2 * k] = pr.x;
        po[2 * k + 1] = pr.y;
      }
    }
  }
}
#undef MY
#undef MYR

// transforms of the winning orientations: out[k] = orient[idx[k]]
__global__ void k_gather_orient(const float* __restrict__ orient, const int* __restrict__ idx, int n,
                                float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * 12) out[t] = orient[12 * (size_t)idx[t / 12] + t % 12];
}

// Refine: orientation that optimally superimposes an alignment
// (PairwiseAlignment::CalcMinRMSD, pair_alignment.cc:714-820), one problem per block: block b
// superimposes the n[b] pairs at pairs[b * N ...].  The host waits for this kernel, so the
// block works as a team: every thread gathers / centres / multiplies for its points, and each
// of the 6 centroid sums, 9 covariance sums and E0 is added up by ONE thread in point order
// (the reference's order; float addition is not associative) from shared memory -- 16 short
// dependent chains side by side instead of one thread walking all of them twice.
__global__ void k_superimpose_pairs_batch(MrDevParams P, const float* __restrict__ A,
                                          const float* __restrict__ B, const short2* __restrict__ pairs,
                                          const int* __restrict__ n_of, int N, float* __restrict__ T_out) {
  extern __shared__ float sp_smem[];
  const int nb = P.nb, b = blockIdx.x, n = n_of[b], np = n * nb;
  const int npp = N * nb;                 // pitch of the per-point arrays
  float* val = sp_smem;                   // [12][npp]: f0 f1 f2 m0 m1 m2, then reused for the products
  float* sq = val + 12 * (size_t)npp;     // [3][npp]: squared differences, point-major order below
  __shared__ float cen[6];
  __shared__ float sums[10];
  const short2* mine = pairs + (size_t)b * N;
  for (int a = threadIdx.x; a < np; a += blockDim.x) {
    const int k = a / nb, q = a - k * nb;
    const short2 pr = mine[k];
    const float* f = A + 3 * ((size_t)pr.x * nb + q);
    const float* mv = B + 3 * ((size_t)pr.y * nb + q);
    val[a] = f[0]; val[npp + a] = f[1]; val[2 * npp + a] = f[2];
    val[3 * npp + a] = mv[0]; val[4 * npp + a] = mv[1]; val[5 * npp + a] = mv[2];
  }
  __syncthreads();
  if (threadIdx.x < 6) {                  // centroids: cf0 cf1 cf2 cm0 cm1 cm2
    const float* v = val + (size_t)threadIdx.x * npp;
    float acc = 0.f;
    for (int a = 0; a < np; a++) acc += v[a];
    cen[threadIdx.x] = acc * (1.0f / (float)np);
  }
  __syncthreads();
  for (int a = threadIdx.x; a < np; a += blockDim.x) {
    float f[3], mv[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { f[i] = val[i * npp + a] - cen[i]; mv[i] = val[(3 + i) * npp + a] - cen[3 + i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float df = f[i] - mv[i];
      sq[i * npp + a] = df * df;
    }
    // M[i][j] += mv[i] * f[j]: the 9 products of this point (overwrites the gathered values,
    // which only this thread reads)
    float pr9[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) pr9[3 * i + j] = mv[i] * f[j];
#pragma unroll
    for (int q = 0; q < 9; q++) val[q * npp + a] = pr9[q];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    const float* v = val + (size_t)threadIdx.x * npp;
    float acc = 0.f;
    for (int a = 0; a < np; a++) acc += v[a];
    sums[threadIdx.x] = acc;
  } else if (threadIdx.x == 32) {         // E0 += df0^2; += df1^2; += df2^2 per point, in point order
    float acc = 0.f;
    for (int a = 0; a < np; a++) { acc += sq[a]; acc += sq[npp + a]; acc += sq[2 * npp + a]; }
    sums[9] = acc;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float M[3][3], T[12];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) M[i][j] = sums[3 * i + j];
  mr_superimpose_finish(np, cen[0], cen[1], cen[2], cen[3], cen[4], cen[5], M, sums[9], T);
  for (int k = 0; k < 12; k++) T_out[12 * (size_t)b + k] = T[k];
}

// The same with one thread per problem: fallback when the team version's work space (15 floats
// per point) exceeds the shared memory of a block.
// Refine: orientation that optimally superimposes an alignment
// (PairwiseAlignment::CalcMinRMSD, pair_alignment.cc:714-820), one single-threaded problem per
// block: block b superimposes the n[b] pairs at pairs[b * N ...]
__global__ void k_superimpose_pairs_batch_seq(MrDevParams P, const float* __restrict__ A,
                                          const float* __restrict__ B, const short2* __restrict__ pairs,
                                          const int* __restrict__ n_of, int N, float* __restrict__ T_out) {
  // the block stages both molecules and the pair list in shared memory (coalesced); thread 0
  // then runs the sequential sums from there instead of chasing ~2n dependent global loads
  extern __shared__ float sp_smem[];
  const int nb = P.nb, b = blockIdx.x, n = n_of[b];
  float* sA = sp_smem;                    // m*nb*3
  float* sB = sA + P.m * nb * 3;          // nB*nb*3
  short2* sp = reinterpret_cast<short2*>(sB + P.nB * nb * 3);  // n pairs
  for (int k = threadIdx.x; k < P.m * nb * 3; k += blockDim.x) sA[k] = A[k];
  for (int k = threadIdx.x; k < P.nB * nb * 3; k += blockDim.x) sB[k] = B[k];
  const short2* mine = pairs + (size_t)b * N;
  for (int k = threadIdx.x; k < n; k += blockDim.x) sp[k] = mine[k];
  __syncthreads();
  if (threadIdx.x != 0) return;
  float T[12];
  mr_superimpose(n * nb, [&](int a, float* pf, float* pm) {
    const int k = a / nb, q = a - k * nb;
    const short2 pr = sp[k];
    const float* f = sA + 3 * (pr.x * nb + q);
    const float* mv = sB + 3 * (pr.y * nb + q);
    pf[0] = f[0]; pf[1] = f[1]; pf[2] = f[2];
    pm[0] = mv[0]; pm[1] = mv[1]; pm[2] = mv[2];
  }, T);
  for (int k = 0; k < 12; k++) T_out[12 * (size_t)b + k] = T[k];
}

// K7 for a handful of survivors whose result the host waits for (Refine batches, mr_eval_one):
// one WARP per survivor.  The sequential n-loop of k_select is a prefix -- everything before
// the first size at which the reference breaks -- so the lanes test 32 sizes at once and a
// ballot finds the break: one memory latency per 32 sizes instead of one per 8.
__global__ void k_select_warp(MrDevParams P, int NP, const int* __restrict__ surv,
                              const int* __restrict__ surv_count, const float* __restrict__ cost,
                              const float* __restrict__ ssd, long long id_base,
                              unsigned long long* __restrict__ keys, unsigned char* __restrict__ valid,
                              int* __restrict__ n_last_out) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= *surv_count) return;
  const unsigned long long oid = (unsigned long long)(id_base + surv[s]);
  int n_last = min(1, P.n_max);
  for (int n0 = 2; n0 <= P.n_max; n0 += 32) {
    const int n = n0 + lane;
    const bool in = n <= P.n_max;
    const size_t slot = (size_t)s * NP + (size_t)(min(n, NP) - 1);
    const float cv = cost[slot], v = ssd[slot];
    const bool brk = in && (cv >= MR_UB || (n >= P.n_min && P.max_rmsd >= 0.0f &&
                                           v > (float)(n * P.nb) * (P.max_rmsd * P.max_rmsd)));
    const unsigned m = __ballot_sync(0xffffffffu, brk);
    const int n_break = m ? n0 + (__ffs(m) - 1) : 0x7fffffff;  // the loop stops AT this size
    if (in && n < n_break && n >= P.n_min) {
      if (valid) valid[slot] = 1;
      if (keys && v < MR_UB) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | oid;
        if (key < *(volatile unsigned long long*)&keys[n - P.n_min]) atomicMin(&keys[n - P.n_min], key);
      }
    }
    if (m) { n_last = n_break; break; }
    n_last = min(n0 + 31, P.n_max);
  }
  if (n_last_out && lane == 0) n_last_out[s] = n_last;
}

// ---------------------------------------------------------------------------
// K7: one thread per survivor: replay the reference's sequential n-loop early exits
// (search_or_dyn3d.cc:1714-1721 max-dist, 1810-1811 max-rmsd) as a prefix over n, then
// atomicMin the 64-bit key (float_bits(ssd) << 32 | o_id) per size: strict "<" with the
// lowest orientation id winning ties is exactly search_or_dyn3d.cc:1825-1835 evaluated
// in id order.  Optionally writes valid[] / n_last for mr_eval_one, and every compared score
// (scores, mr_search_scores) for the alternate-solutions bookkeeping; excl (mr_search_dissimilar)
// marks the (survivor, size) alignments that may not be recorded.
// ---------------------------------------------------------------------------
__global__ void k_select(MrDevParams P, int NP, const int* __restrict__ surv,
                         const int* __restrict__ surv_count, const float* __restrict__ cost,
                         const float* __restrict__ ssd, long long id_base,
                         unsigned long long* __restrict__ keys, unsigned char* __restrict__ valid,
                         int* __restrict__ n_last_out, float* __restrict__ scores = nullptr,
                         const unsigned char* __restrict__ excl = nullptr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= *surv_count) return;
  const unsigned long long oid = (unsigned long long)(id_base + surv[s]);
  int n_last = min(1, P.n_max);
  // the loop exits on loaded values, so its loads are fetched eight sizes ahead (rows are padded
  // to NP, a multiple of 32): one memory latency per eight sizes instead of one per size --
  // what a Refine batch (a few dozen threads, the host waiting) is bound by
  bool done = false;
  for (int n0 = 2; n0 <= P.n_max && !done; n0 += 8) {
    float cv[8], sv8[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int n = min(n0 + q, NP);
      cv[q] = cost[(size_t)s * NP + (n - 1)];
      sv8[q] = ssd[(size_t)s * NP + (n - 1)];
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int n = n0 + q;
      if (done || n > P.n_max) continue;
      const size_t slot = (size_t)s * NP + (n - 1);
      n_last = n;
      if (cv[q] >= MR_UB) { done = true; continue; }
      if (n >= P.n_min) {
        const float v = sv8[q];
        if (P.max_rmsd >= 0.0f && v > (float)(n * P.nb) * (P.max_rmsd * P.max_rmsd)) { done = true; continue; }
        if (valid) valid[slot] = 1;
        // -alt-o: the score KeepTrackOfAltSolutions sees for (orientation, size)
        // (search_or_dyn3d.cc:1897-1904), row = position of the orientation in the searched list
        if (scores && v < MR_UB) scores[(size_t)surv[s] * (P.n_max - P.n_min + 1) + (n - P.n_min)] = v;
        // -alt-r: an alignment too similar to a lower rank's still takes part in the early exits
        // above, it just cannot be recorded (search_or_dyn3d.cc:2080-2107)
        if (keys && v < MR_UB && !(excl && excl[slot])) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | oid;
          // most candidates lose: a plain (possibly stale, only ever too large) read first
          // keeps the 64-bit atomics on the ~140 hot addresses rare
          if (key < *(volatile unsigned long long*)&keys[n - P.n_min]) atomicMin(&keys[n - P.n_min], key);
        }
      }
    }
  }
  if (n_last_out) n_last_out[s] = n_last;
}

// ---------------------------------------------------------------------------
// K8: per-size statistics of align_chimera.plot (SearchOrDyn3d::WritePlotFile,
// bin/minrms/search_or_dyn3d.cc:2703-2760): one warp per size r of a best table.
//   furthest[r] = FindFurthestResPair (lib/pair_alignment/pair_alignment.cc:2351-2379,
//                 84-105): sqrt of the largest per-pair sum over the nb atoms of squared
//                 distances between A_i and T.B_j -- a max, so lanes take pairs k, k+32, ...
//   sstr[r]     = LevittGerstein98::Sstr_Score (pair_alignment.cc:2690-2790): sum over the
//                 pairs, in order, of 20 / (1 + d_CA^2 / 25) minus 10 per gap.  The terms
//                 (transform, distance, division) are produced by all lanes into shared
//                 memory; lane 0 then adds them in the reference's order (float addition
//                 is not associative).
// Transform = ApplyTransformBackbone (lib/biopolymer/apply_trans.cc:72-108) of the two
// atoms involved only.  Sizes without a solution are skipped (has[r] = 0).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_plot_stats(int nb, int ca, int n_min, int N, const float* __restrict__ A, const float* __restrict__ B,
             const float* __restrict__ ssd, const float* __restrict__ T12, const short* __restrict__ pairs,
             float* __restrict__ furthest, float* __restrict__ sstr, int* __restrict__ has) {
  extern __shared__ float plot_terms[];  // [N] S_str terms, then [N] gap flags (as float 0 / 1)
  const int r = blockIdx.x, lane = threadIdx.x, n = n_min + r;
  if (!(ssd[r] < MR_UB)) {
    if (lane == 0) { has[r] = 0; furthest[r] = 0.0f; sstr[r] = 0.0f; }
    return;
  }
  const float* T = T12 + 12 * (size_t)r;
  const short* pr = pairs + (size_t)r * 2 * N;
  float* term = plot_terms;
  float* gap = plot_terms + N;
  float far = 0.0f;
  for (int k = lane; k < n; k += 32) {
    const int i = pr[2 * k], j = pr[2 * k + 1];
    float c = 0.0f, dca = 0.0f;
    for (int q = 0; q < nb; q++) {
      const float* a = A + 3 * ((size_t)i * nb + q);
      const float* b = B + 3 * ((size_t)j * nb + q);
      float x, y, z;
      mr_xform(T, b[0], b[1], b[2], x, y, z);
      const float d = mr_d2(a[0], a[1], a[2], x, y, z);
      c += d;
      if (q == ca) dca = d;
    }
    far = fmaxf(far, c);
    term[k] = 20.0f / (1 + (dca / 25.0f));
    const int pi = (k > 0) ? pr[2 * k - 2] : -1, pj = (k > 0) ? pr[2 * k - 1] : -1;
    gap[k] = (k > 0 && (i - pi - 1) + (j - pj - 1) > 0) ? 1.0f : 0.0f;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) far = fmaxf(far, __shfl_xor_sync(0xffffffffu, far, off));
  __syncwarp();
  if (lane == 0) {
    float total = 0.0f;
    for (int k = 0; k < n; k++) {
      total += term[k];
      if (gap[k] != 0.0f) total -= 10.0f;
    }
    has[r] = 1;
    furthest[r] = sqrtf(far);
    sstr[r] = total;
  }
}
