// mr_dp_tiled.cuh -- register-tiled 3-D DP kernel (K1+K3+K4 for problems whose (layer,row)
// cross-section fits one CTA's registers).
//
// Recurrence (Dyn3d::Solve_IncrN, lib/dyn/dyn3d.cc:183-387; layer 1: InitTable_IncrN 135-179):
//     A(n,i,j) = min{ A(n-1,i-1,j-1) + C(i,j),  A(n,i-1,j),  A(n,i,j-1) }
// with the reference's tie order (match, then "<=" skip-i, then "<=" skip-j) recorded as two
// bits per cell.  Written over the whole (n,i,j) box with the boundary values
//     A(0,.,.) = 0,   A(n>=1, 0, j) = A(n>=1, i, 0) = +inf
// it reproduces the reference's special-cased first row / first column of every layer
// exactly: cells with i < n or j < n stay +inf and never win a min, and live cells only ever
// read live cells or those boundaries (layer 0 is carried as a real, constant-0 layer).
//
// Mapping (one CTA = one orientation, a small dataflow machine):
//   * the sweep axis is j (columns of molecule 2), one column per step; the (layer,row)
//     cross-section is cut into K x R tiles (K consecutive layers x R consecutive rows) and
//     every LIVE tile is owned by one thread, which keeps it in registers for the whole
//     sweep; the tiles of row strip w run w columns behind strip 0 (wavefront);
//   * per step a thread reads, from shared memory, the K boundary values A(n, i_lo-1, j) its
//     upper neighbour published one step earlier and the R values A(n_lo-1, i-1, j-1) its
//     lower-layer neighbour published one step earlier, updates its K*R cells in registers
//     and publishes its own bottom row / top layer (double-buffered by step parity, one
//     __syncthreads per step).  No warp shuffles: neighbours may sit anywhere, so
//   * tiles are packed into warps sorted by the step at which they become live (cells with
//     j < n are dead): a warp skips the steps outside the window of its lanes, which removes
//     most of the triangular waste of the (n <= i, n <= j) pyramid;
//   * C(i,j) = |A_i - B'_j|^2 is produced 8 columns at a time by all threads of the CTA
//     into a 32-column shared-memory ring (+inf outside the matrix, so pipeline fill and
//     drain steps leave the state untouched);
//   * the two tie-break predicates NOT(up <= d) and NOT(left <= min(d, up)) are the sign bits
//     of float differences (exact for IEEE floats), funnel-shifted into a per-layer
//     accumulator.  Two forms of the cell update (template parameter PP):
//       ping-pong (CTAs of 320+ threads, ~150 registers, 1 CTA per SM): the state is kept
//         twice, read from one copy and written to the other, so that
//         best = min3(d, up, left) is one FMNMX3 and s2 = best - left still sees the old
//         value: 3 FADD (FMA pipes) + FMNMX3 + 2 SHF (ALU pipe) per cell;
//       two-min (smaller CTAs, 88 registers, 2-10 CTAs per SM): t = min(d, up),
//         s2 = t - left, best = min(t, left) updated in place: one more FMNMX.
#pragma once
#include <algorithm>
#include <type_traits>
#include <vector>
#include "mr_device.cuh"
#include "mr_tile_step.cuh"
#ifndef MR_DP_TILED_DEVICE_ONLY
#include "mr_vp_module.h"
#endif

// Traceback storage of the tiled kernel.  Each tile (thread t of the CTA) writes, every
// second column, one 32-bit word per layer of the tile: 2 columns x R=8 rows x 2 bits,
// appended MSB-first in processing order; the K words of a tile are one 16-byte store and
// the stores of a warp are contiguous: word index = ((jp * T) + t) * K + k.
// bit pair of cell (row r of the strip, column half h): s1 at bit 31-(16h+2r), s2 one below.
//   s1 = 1  <=>  NOT (up <= d)        s2 = 1  <=>  NOT (left <= min(d,up))
// tile_of[w * CN + c] maps (strip, chunk) to the owning thread.
struct TbTiled {
  int K, R, W, JP, pad, NP, CN, T;  // layers/tile, rows/tile, strips, column pairs, top pad rows, cost pitch, chunks, threads
  const int* tile_of;
  // the (strip, chunk) -> tile table is staged in shared memory by the traceback kernel when
  // it is small (<= 8 KB); larger tables are read through L1/L2
  __host__ __device__ size_t words(int /*nB*/) const { return (size_t)JP * T * K; }
  __host__ __device__ bool table_in_smem() const { return W * CN <= 2048; }
  __host__ __device__ int table_ints() const { return table_in_smem() ? ((W * CN + 3) & ~3) : 0; }
  __device__ __forceinline__ void load_table(int* tab, int tid, int nthreads) const {
    if (!table_in_smem()) return;
    for (int k = tid; k < W * CN; k += nthreads) tab[k] = tile_of[k];
  }
  // K = 4 layers x R = 8 rows per tile (the only shape built): shifts instead of divisions
  // branch-free walk with 32-bit index math on (n, rho = i-1+pad, jm = j-1)
  // GROUP: keep the last 16-byte group (the 4 layers of a chunk at one column pair) in
  // registers.  It removes ~40 % of the dependent loads: no gain when millions of walks hide
  // each other's latency (FirstPass), a shorter critical path when only a few hundred run
  // (Refine batches, winners' payload).
  template <bool GROUP, class Emit>
  __device__ __forceinline__ void walk(const int* __restrict__ tab, const uint32_t* __restrict__ tbo, int n,
                                       int m, int nB, Emit emit) const {
    if (table_in_smem()) walk_t<GROUP>(tab, tbo, n, m, nB, emit);
    else walk_t<GROUP>(tile_of, tbo, n, m, nB, emit);
  }
  template <bool GROUP, class Emit>
  __device__ __forceinline__ void walk_t(const int* __restrict__ tab, const uint32_t* __restrict__ tbo, int n,
                                         int m, int nB, Emit emit) const {
    int cn = n, rho = m - 1 + pad, jm = nB - 1;
    const int T4 = 4 * T;
    const uint4* tb4 = reinterpret_cast<const uint4*>(tbo);
    uint4 grp = make_uint4(0, 0, 0, 0);
    int have = -1;
    while (cn > 0) {
      MR_ASSERT(rho >= pad && jm >= 0 && (rho >> 3) < W && (cn >> 2) < CN);  // a path never leaves the live cells
      const int t = tab[(rho >> 3) * CN + (cn >> 2)];
      MR_ASSERT(t >= 0 && t < T && (jm >> 1) < JP);
      uint32_t word;
      if (GROUP) {
        const int gi = (jm >> 1) * T + t;
        if (gi != have) { grp = tb4[gi]; have = gi; }
        const int kq = cn & 3;
        word = (kq & 2) ? ((kq & 1) ? grp.w : grp.z) : ((kq & 1) ? grp.y : grp.x);
      } else {
        word = tbo[(jm >> 1) * T4 + t * 4 + (cn & 3)];
      }
      const int a = (jm & 1) * 16 + 2 * (rho & 7);
      const uint32_t bits = (word >> (30 - a)) & 3u;      // (s1 << 1) | s2
      // s2 = 0: skip-j;  s2 = 1: s1 ? match : skip-i  (decode() below spells the codes out)
      if (bits == 3u) { emit(cn - 1, rho - pad, jm); cn--; }
      rho -= (int)(bits & 1u);    // match or skip-i
      jm -= (bits != 1u);         // match or skip-j
    }
  }
  __device__ __forceinline__ size_t addr(const int* __restrict__ tab, int n, int i, int j) const {
    const int t = tab[((i - 1 + pad) >> 3) * CN + (n >> 2)];
    return ((size_t)((j - 1) >> 1) * T + t) * 4 + (n & 3);
  }
  __device__ __forceinline__ uint32_t decode(uint32_t word, int i, int j) const {
    const int a = ((j - 1) & 1) * 16 + 2 * ((i - 1 + pad) & 7);
    const uint32_t bits = (word >> (30 - a)) & 3u;  // (s1 << 1) | s2
    return (0x0212u >> (4 * bits)) & 3u;            // s2 = 0 -> 2 (skip-j); s2 = 1: s1 ? 0 (match) : 1 (skip-i)
  }
};

// one entry per thread of the CTA (built on the host by mr_plan_dp_tiled)
struct MrTile {
  int w, c;              // row strip, layer chunk (-1: idle thread)
  int s_lo, s_hi;        // first / last step at which this tile has live cells
  int up_slot, dg_slot;  // exchange slots of the tiles (w-1,c) and (w,c-1) (or an INIT slot)
  int warp_lo, warp_hi;  // step window of the whole warp
};

struct MrTiledPlan {
  bool ok = false;
  const char* why = "not planned";
  int K = 0, R = 8, W = 0, CN = 0;
  int nBe = 0;  // columns rounded up to even
  bool pp = false;  // ping-pong form of the tile update (even step windows)
  int threads = 0, nslots = 0, steps = 0;
  size_t smem = 0;
  int ctas_per_sm = 1;
  int grid_cap = 0;
  int grid_cap_costs = 0;  // costs-only kernel (fewer registers: two CTAs per SM up to 384 threads)
  int vp_bucket = 0;       // CTA-size bucket of the min-with-predicate kernel (mr_dp_vp.cu), 0: none
  int grid_cap_vp = 0;
  mutable void* vp_fn = nullptr;       // CUfunction, looked up at the first launch (per plan: one context, one device)
  size_t tb_words = 0;
  double lane_efficiency = 0.0;
  TbTiled layout;
  MrTile* d_tiles = nullptr;
  int* d_tile_of = nullptr;
};

#define MR_CRING 32  // columns in the C ring (>= strips + MR_CBATCH)
#define MR_CBATCH 8  // columns of C produced at a time
__host__ __device__ inline int mr_ring_stride(int WR) { return WR + ((36 - (WR & 31)) & 31); }

// Distance columns [j0, j0 + MR_CBATCH) into the ring (j0 = 1 mod 8).
// One atom per residue (NB1): molecule 2 is kept as x[], y[], z[] arrays so that a work item
// = one row x 4 consecutive columns reads A_i once and B' as three float4; otherwise the
// generic per-element loop.  Ring rows are WRp floats apart (WRp = 4 mod 32: the float4
// reads of tiles in different strips / columns fall into different banks).
template <bool NB1>
__device__ __forceinline__ void mr_produce_c(float* __restrict__ Cr, const float* __restrict__ sA,
                                             const float* __restrict__ sB, int j0, int WR, int WRp, int pad,
                                             int nB, int nBp, int nb, float cutoff, int tid, int nthreads,
                                             int rho0 = 0) {
  const float INF = __int_as_float(0x7f800000);
  if (NB1) {
    for (int item = tid; item < (MR_CBATCH / 4) * WR; item += nthreads) {
      static_assert(MR_CBATCH == 8, "two groups of 4 columns per batch");
      const int q4 = item >= WR ? 1 : 0, rho = item - q4 * WR;  // item < 2 * WR: no division
      const int jn = j0 + 4 * q4, i = rho0 + rho - pad + 1;
      float cc[4] = {INF, INF, INF, INF};
      if (i >= 1 && jn <= nB) {
        const float ax = sA[3 * (i - 1)], ay = sA[3 * (i - 1) + 1], az = sA[3 * (i - 1) + 2];
        const float4 bx = *reinterpret_cast<const float4*>(sB + (jn - 1));
        const float4 by = *reinterpret_cast<const float4*>(sB + nBp + (jn - 1));
        const float4 bz = *reinterpret_cast<const float4*>(sB + 2 * nBp + (jn - 1));
        cc[0] = mr_d2(ax, ay, az, bx.x, by.x, bz.x);
        cc[1] = mr_d2(ax, ay, az, bx.y, by.y, bz.y);
        cc[2] = mr_d2(ax, ay, az, bx.z, by.z, bz.z);
        cc[3] = mr_d2(ax, ay, az, bx.w, by.w, bz.w);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (cc[q] > cutoff) cc[q] = MR_UB;
          if (jn + q > nB) cc[q] = INF;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) Cr[((jn + q) & (MR_CRING - 1)) * WRp + rho] = cc[q];
    }
  } else {
    // flat loop over (column, row) items; the indices advance without a division per item
    int dj = tid / WR, rho = tid - dj * WR;
    while (dj < MR_CBATCH) {
      const int jn = j0 + dj, i = rho0 + rho - pad + 1;
      float cc = INF;
      if (i >= 1 && jn <= nB) {
        cc = 0.0f;
        for (int q = 0; q < nb; q++) {
          const float* a = sA + 3 * ((i - 1) * nb + q);
          const float* b = sB + 3 * ((jn - 1) * nb + q);
          cc += mr_d2(a[0], a[1], a[2], b[0], b[1], b[2]);
        }
        if (cc > cutoff) cc = MR_UB;
      }
      Cr[(jn & (MR_CRING - 1)) * WRp + rho] = cc;
      rho += nthreads;
      while (rho >= WR) { rho -= WR; dj++; }
    }
  }
}

// the kernel body (a device function so that the min-with-predicate form can be instantiated
// under extern "C" names in its own translation unit, mr_dp_vp.cu)
template <int K, int R, bool NB1, bool WINDOWED, bool PP, bool COSTS, bool VP>
__device__ __forceinline__ void
mr_dp_tiles_body(const MrDevParams& P, const TbTiled& L, int nslots, int nsteps, const MrTile* __restrict__ tiles,
                 const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ orient,
                 const int* __restrict__ surv, const int* __restrict__ surv_count,
                 uint32_t* __restrict__ tb, float* __restrict__ cost_out) {
  static_assert(R == MR_TR && K == MR_TK, "traceback word packing assumes 4 x 8 tiles");
  extern __shared__ float4 smem4[];
  const int m = P.m, nB = P.nB, nb = P.nb;
  const int W = L.W, WR = W * R, pad = L.pad;
  const int WRp = mr_ring_stride(WR);   // ring row stride (floats), = 4 mod 32
  const int nBp = (nB * nb + 11) & ~3;  // padded length of one coordinate array of molecule 2
  const int tid = threadIdx.x, nthreads = blockDim.x;
  // shared memory carve-up (all 16-byte aligned).  Exchange slot (parity p, slot q) is the
  // 48 bytes at exch[(p*nslots + q)*3]: {top layer: upper boundary, rows 0..2}, {rows 3..6},
  // {bottom row of the K = 4 layers}.  Consecutive slots are 48 B apart: 16-byte reads of
  // neighbouring slots by a quarter warp fall into distinct banks.
  float4* exch = smem4;
  float* Cr = reinterpret_cast<float*>(exch + 6 * nslots);  // [MR_CRING][WRp] distance columns
  float* sB = Cr + MR_CRING * WRp;                           // rotated molecule 2: x[], y[], z[] (NB1) or xyz triples
  float* sA = sB + 3 * nBp;                                  // molecule 1: m*nb*3
  const float INF = __int_as_float(0x7f800000);
  const int nsurv = *surv_count;

  const MrTile me = tiles[tid];
  MR_ASSERT(me.up_slot >= 0 && me.up_slot < nslots && me.dg_slot >= 0 && me.dg_slot < nslots && (int)blockDim.x + 2 == nslots);
  const bool has_tile = me.w >= 0;
  const int w = has_tile ? me.w : 0, c = has_tile ? me.c : 0;
  const int warp_lo = me.warp_lo, warp_hi = min(me.warp_hi, nsteps);
  // steps at which this tile stores traceback words / reports costs
  const int tb_hi = has_tile ? me.s_hi : -1;
  const int tb_first = has_tile ? me.s_lo + ((me.s_lo - w) & 1) : 0x3fffffff;  // first step with an even column
  const int cost_step = (has_tile && w == W - 1) ? nB + w : -1;
  // 32-bit shared addresses: parity set 0; set 1 is `par_bytes` further
  const uint32_t exch0 = (uint32_t)__cvta_generic_to_shared(exch);
  const uint32_t par_bytes = (uint32_t)nslots * 48u;
  const uint32_t a_up = exch0 + (uint32_t)me.up_slot * 48u + 32u;  // bottom row of the tile above
  const uint32_t a_dg = exch0 + (uint32_t)me.dg_slot * 48u;        // top layer of the tile one chunk down
  const uint32_t a_me = exch0 + (uint32_t)tid * 48u;
  const uint32_t a_c = (uint32_t)__cvta_generic_to_shared(Cr) + (uint32_t)(w * R) * 4u;
  const uint32_t c_row = (uint32_t)WRp * 4u;
  for (int k = tid; k < m * nb * 3; k += nthreads) sA[k] = A[k];

  for (int sv = blockIdx.x; sv < nsurv; sv += gridDim.x) {
    const float* T = orient + 12 * (size_t)surv[sv];
    __syncthreads();
    for (int k = tid; k < nBp; k += nthreads) {
      float x = 0.f, y = 0.f, z = 0.f;
      if (k < nB * nb) mr_xform(T, B[3 * k], B[3 * k + 1], B[3 * k + 2], x, y, z);
      if (NB1) {
        sB[k] = x; sB[nBp + k] = y; sB[2 * nBp + k] = z;
      } else if (k < nB * nb) {
        sB[3 * k] = x; sB[3 * k + 1] = y; sB[3 * k + 2] = z;
      }
    }
    // exchange slots start at the boundary values
    // (slot nslots-2: {0, inf, ..} above strip 0 for chunk 0; slot nslots-1: all +inf)
    for (int k = tid; k < 2 * nslots; k += nthreads) {
      const int sl = k % nslots;
      const bool zero = sl == nslots - 2 || (sl < nslots - 2 && tiles[sl].c == 0);
      exch[3 * k] = make_float4(INF, INF, INF, INF);
      exch[3 * k + 1] = make_float4(INF, INF, INF, INF);
      exch[3 * k + 2] = make_float4(zero ? 0.0f : INF, INF, INF, INF);
    }
    for (int k = tid; k < MR_CRING * WRp; k += nthreads) Cr[k] = INF;
    __syncthreads();
    mr_produce_c<NB1>(Cr, sA, sB, 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads);

    // tile state: the ping-pong form keeps two copies (read one, write the other), the in-place
    // form one; bprev = upper boundary A(n, i_lo-1, .) of the previous column (and, ping-pong, this one)
    static_assert(!(COSTS && PP) && !(VP && (PP || COSTS)), "the costs-only and min-with-predicate updates are in place");
    constexpr int NC = PP ? 2 : 1;
    float st[NC][K][R], bprev[NC][K];
    uint32_t acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const float v = (c * K + k == 0) ? 0.0f : INF;
      acc[k] = 0;
#pragma unroll
      for (int q = 0; q < NC; q++) {
        bprev[q][k] = v;
#pragma unroll
        for (int r = 0; r < R; r++) st[q][k][r] = v;
      }
    }
    uint4* tbp = reinterpret_cast<uint4*>(tb + (size_t)sv * L.words(nB)) + tid +
                 (size_t)max(0, (tb_first - w - 1) >> 1) * nthreads;
    int tb_next = tb_first;
    float* const cost_row = cost_out + (size_t)sv * L.NP;
    __syncthreads();

    // steps before this warp has live cells: keep pace with the CTA (barrier + C production)
    int s = 1;
    for (; s < warp_lo && s <= nsteps; s++) {
      if ((s & (MR_CBATCH - 1)) == 0)
        mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads);
      __syncthreads();
    }
    // one column: read the neighbours' slots of parity rd_off, write mine of parity wr_off
    auto step = [&](auto src_tag, const int s, const uint32_t rd_off, const uint32_t wr_off) {
      constexpr int S = decltype(src_tag)::value, D = PP ? 1 - S : 0;
      float cv[R], dsh[R];
      mr_tile_load8(cv, a_c + (uint32_t)((s - w) & (MR_CRING - 1)) * c_row);  // ring column j & 31 (+inf outside the matrix)
      const float4 ub = mr_lds128(a_up + rd_off);                             // A(n, i_lo-1, j): bottom row of the tile above
      mr_tile_load8(dsh, a_dg + rd_off);
      if constexpr (PP) {
        bprev[D][0] = ub.x; bprev[D][1] = ub.y; bprev[D][2] = ub.z; bprev[D][3] = ub.w;
        mr_tile_update_pp<S>(st, bprev, acc, cv, dsh);
      } else {
        const float bin[K] = {ub.x, ub.y, ub.z, ub.w};
        if constexpr (COSTS) mr_tile_update_costs(st[0], bprev[0], bin, cv, dsh);
        else if constexpr (VP) mr_tile_update_vp(st[0], bprev[0], bin, acc, cv, dsh);
        else mr_tile_update_ip(st[0], bprev[0], bin, acc, cv, dsh);
      }
      mr_tile_publish(a_me + wr_off, bprev[D][K - 1], st[D][K - 1], st[D][0][R - 1], st[D][1][R - 1], st[D][2][R - 1],
                      st[D][3][R - 1]);
      // traceback: the K words of this tile every second column (j = s - w even)
      if (!COSTS && s == tb_next) {
        MR_ASSERT(tbp >= reinterpret_cast<uint4*>(tb + (size_t)sv * L.words(nB)) &&
                  tbp + 1 <= reinterpret_cast<uint4*>(tb + (size_t)(sv + 1) * L.words(nB)));
        *tbp = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        tbp += nthreads;
        tb_next = (s + 2 <= tb_hi) ? s + 2 : 0x3fffffff;
      }
      // cost(n) = A(n, m, nB) for n >= n_min (row m is the last row of the last strip)
      if (s == cost_step) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          const int n = c * K + k;
          if (n >= P.n_min && n <= P.n_max) cost_row[n - 1] = st[D][k][R - 1];
        }
      }
      if (WINDOWED && has_tile) mr_tile_window_cost(P, st[D], s - w, w, c, pad, cost_row);
      // every MR_CBATCH steps all threads produce the next MR_CBATCH columns of C
      if ((s & (MR_CBATCH - 1)) == 0)
        mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads);
      __syncthreads();
    };
    // parity of the exchange set written at step s: s & 1
    const uint32_t off_a = (s & 1) ? 0u : par_bytes, off_b = par_bytes - off_a;
    if constexpr (PP) {
      // the plan makes every warp window an even number of steps
      for (; s < warp_hi; s += 2) {
        step(std::integral_constant<int, 0>(), s, off_a, off_b);
        step(std::integral_constant<int, 1>(), s + 1, off_b, off_a);
      }
    } else {
      uint32_t rd_off = off_a, wr_off = off_b;
      for (; s <= warp_hi; s++) {
        step(std::integral_constant<int, 0>(), s, rd_off, wr_off);
        const uint32_t t = rd_off; rd_off = wr_off; wr_off = t;
      }
    }
    // steps after this warp's window
    for (; s <= nsteps; s++) {
      if ((s & (MR_CBATCH - 1)) == 0)
        mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads);
      __syncthreads();
    }
  }
}

#ifndef MR_DP_TILED_DEVICE_ONLY   // mr_dp_vp.cu takes the kernel body only
template <int K, int R, int MAXT, int MINB, bool NB1, bool WINDOWED, bool PP, bool COSTS = false>
__global__ void __launch_bounds__(MAXT, MINB)
k_dp_tiles(MrDevParams P, TbTiled L, int nslots, int nsteps, const MrTile* __restrict__ tiles,
           const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ orient,
           const int* __restrict__ surv, const int* __restrict__ surv_count,
           uint32_t* __restrict__ tb, float* __restrict__ cost_out) {
  mr_dp_tiles_body<K, R, NB1, WINDOWED, PP, COSTS, false>(P, L, nslots, nsteps, tiles, A, B, orient, surv, surv_count, tb, cost_out);
}

// ---------------------------------------------------------------------------------- host
static inline int mr_dp_tiles_minb(int threads);
static inline int mr_dp_vp_bucket(int threads);
static inline int mr_dp_vp_minb(int bucket);
static inline int mr_dp_tiles_minb_costs(int threads);
static inline bool mr_dp_tiles_pp(int threads);
static inline void mr_free_dp_tiled(MrTiledPlan* p) {
  if (p->d_tiles) cudaFree(p->d_tiles);
  if (p->d_tile_of) cudaFree(p->d_tile_of);
  p->d_tiles = nullptr;
  p->d_tile_of = nullptr;
}

static inline void mr_plan_dp_tiled(MrTiledPlan* p, const MrDevParams& P, const cudaDeviceProp& prop, int NP) {
  p->ok = false;
  const int R = 8, K = 4;
  const int m = P.m, nB = P.nB;
  const int W = (m + R - 1) / R, pad = W * R - m;
  const int CN = (P.n_max + 1 + K - 1) / K;  // chunks incl. the constant layer 0
  const int nBe = (nB + 1) & ~1;
  if (W + MR_CBATCH > MR_CRING) { p->why = "more than 192 rows"; return; }
  // live tiles and their step windows
  struct T { int w, c, lo, hi; };
  std::vector<T> ts;
  std::vector<int> index((size_t)W * CN, -1);
  long long live_cells = 0;
  for (int n = 1; n <= P.n_max; n++)
    live_cells += (long long)(std::min(m, n + P.wA - 1) - n + 1) * (std::min(nB, n + P.wB - 1) - n + 1);
  for (int w = 0; w < W; w++) {
    const int i_lo = w * R + 1 - pad, i_hi = w * R + R - pad;
    for (int c = 0; c < CN; c++) {
      int jlo = 1 << 30, jhi = -1;
      for (int n = std::max(1, c * K); n <= std::min(P.n_max, c * K + K - 1); n++) {
        const int imax = std::min(m, n + P.wA - 1);
        if (std::max(n, i_lo) > std::min(imax, i_hi)) continue;
        jlo = std::min(jlo, n);
        jhi = std::max(jhi, std::min(nB, n + P.wB - 1));
      }
      if (c == 0) { jlo = std::min(jlo, 1); jhi = std::max(jhi, 1); }  // the base-layer chain always exists
      if (jhi < 0) continue;
      jhi = (jhi + 1) & ~1;  // finish the traceback word of the last column pair
      ts.push_back({w, c, jlo + w, jhi + w});
    }
  }
  // closure: the lowest layer of tile (w,c) reads A(n_lo-1, i_lo-1, j-1) from what tile
  // (w,c-1) relays of ITS upper boundary, so (w,c-1) must run even where all of its own
  // cells are outside the band (-minN > 1): add it with the window of (w,c), one step early
  {
    std::vector<int> have((size_t)W * CN, -1);
    for (size_t t = 0; t < ts.size(); t++) have[(size_t)ts[t].w * CN + ts[t].c] = (int)t;
    const size_t n0 = ts.size();
    for (size_t t = 0; t < n0; t++) {
      const T x = ts[t];
      if (x.c >= 1 && have[(size_t)x.w * CN + x.c - 1] < 0) {
        have[(size_t)x.w * CN + x.c - 1] = (int)ts.size();
        ts.push_back({x.w, x.c - 1, std::max(1, x.lo - 1), x.hi});
      }
    }
  }
  std::stable_sort(ts.begin(), ts.end(), [](const T& a, const T& b) { return a.lo != b.lo ? a.lo < b.lo : a.hi < b.hi; });
  const int ntiles = (int)ts.size();
  const int threads = ((ntiles + 31) / 32) * 32;
  if (threads > 1024) { p->why = "more than 1024 live tiles"; return; }
  for (int t = 0; t < ntiles; t++) index[(size_t)ts[t].w * CN + ts[t].c] = t;
  const int nslots = threads + 2;
  const bool pp = mr_dp_tiles_pp(threads);
  std::vector<MrTile> tab((size_t)threads);
  int steps = 0;
  long long warp_steps = 0;
  for (int wp = 0; wp < threads / 32; wp++) {
    int lo = 1 << 30, hi = -1;
    for (int t = wp * 32; t < std::min(ntiles, wp * 32 + 32); t++) { lo = std::min(lo, ts[t].lo); hi = std::max(hi, ts[t].hi); }
    if (pp && hi >= lo && ((hi - lo + 1) & 1)) { if (lo > 1) lo--; else hi++; }  // steps come in pairs
    warp_steps += hi - lo + 1;
    steps = std::max(steps, hi);
    for (int t = wp * 32; t < wp * 32 + 32; t++) {
      MrTile& e = tab[t];
      e.warp_lo = lo; e.warp_hi = hi;
      if (t >= ntiles) { e.w = -1; e.c = -1; e.s_lo = 1 << 30; e.s_hi = -1; e.up_slot = nslots - 1; e.dg_slot = nslots - 1; continue; }
      e.w = ts[t].w; e.c = ts[t].c; e.s_lo = ts[t].lo; e.s_hi = ts[t].hi;
      const int up = (e.w > 0) ? index[(size_t)(e.w - 1) * CN + e.c] : -1;
      const int dg = (e.c > 0) ? index[(size_t)e.w * CN + e.c - 1] : -1;
      e.up_slot = up >= 0 ? up : ((e.c == 0) ? nslots - 2 : nslots - 1);
      e.dg_slot = dg >= 0 ? dg : nslots - 1;
    }
  }
  const size_t smem = 16 * (size_t)(6 * nslots) + 4 * (size_t)(MR_CRING * mr_ring_stride(W * R) + 3 * ((nB * P.nb + 11) & ~3) + ((m * P.nb * 3 + 3) & ~3));
  if (smem > prop.sharedMemPerBlockOptin) { p->why = "work space does not fit shared memory"; return; }
  mr_free_dp_tiled(p);
  if (cudaMalloc(&p->d_tiles, sizeof(MrTile) * tab.size()) != cudaSuccess) { p->why = "cudaMalloc"; return; }
  if (cudaMemcpy(p->d_tiles, tab.data(), sizeof(MrTile) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) { p->why = "cudaMemcpy"; return; }
  for (auto& v : index) if (v < 0) v = 0;
  if (cudaMalloc(&p->d_tile_of, sizeof(int) * index.size()) != cudaSuccess) { p->why = "cudaMalloc"; return; }
  if (cudaMemcpy(p->d_tile_of, index.data(), sizeof(int) * index.size(), cudaMemcpyHostToDevice) != cudaSuccess) { p->why = "cudaMemcpy"; return; }
  p->K = K; p->R = R; p->W = W; p->CN = CN; p->nBe = nBe;
  p->threads = threads; p->nslots = nslots; p->steps = steps; p->pp = pp;
  p->smem = smem;
  const int by_smem = (int)(prop.sharedMemPerMultiprocessor / (smem + 1024));
  const int want = mr_dp_tiles_minb(threads);  // CTAs per SM the launch-bound bucket of this CTA size allows
  p->ctas_per_sm = std::max(1, std::min(by_smem, want));
  p->grid_cap = prop.multiProcessorCount * p->ctas_per_sm;
  p->grid_cap_costs = prop.multiProcessorCount * std::max(1, std::min(by_smem, mr_dp_tiles_minb_costs(threads)));
  p->vp_bucket = mr_dp_vp_bucket(threads);
  p->vp_fn = nullptr;
  p->grid_cap_vp = prop.multiProcessorCount * std::max(1, std::min(by_smem, mr_dp_vp_minb(p->vp_bucket)));
  p->layout.K = K; p->layout.R = R; p->layout.W = W; p->layout.JP = nBe / 2;
  p->layout.pad = pad; p->layout.NP = NP; p->layout.CN = CN; p->layout.T = threads;
  p->layout.tile_of = p->d_tile_of;
  p->tb_words = p->layout.words(nB);
  p->lane_efficiency = (double)live_cells / ((double)warp_steps * 32 * K * R);
  p->ok = true;
  p->why = "";
}

template <int MAXT, int MINB, bool PP>
static inline int mr_launch_dp_tiles_t(const MrTiledPlan& p, const MrDevParams& P, const float* A,
                                       const float* B, const float* orient, const int* surv,
                                       const int* surv_count, uint32_t* tb, float* cost, int count,
                                       cudaStream_t stream, bool costs_only) {
  typedef void (*kern_t)(MrDevParams, TbTiled, int, int, const MrTile*, const float*, const float*, const float*,
                         const int*, const int*, uint32_t*, float*);
  static const kern_t kerns[4] = {k_dp_tiles<4, 8, MAXT, MINB, false, false, PP>, k_dp_tiles<4, 8, MAXT, MINB, false, true, PP>,
                                  k_dp_tiles<4, 8, MAXT, MINB, true, false, PP>, k_dp_tiles<4, 8, MAXT, MINB, true, true, PP>};
  constexpr int MINB_C = MAXT <= 384 ? (MINB > 2 ? MINB : 2) : 1;
  static const kern_t kerns_c[4] = {k_dp_tiles<4, 8, MAXT, MINB_C, false, false, false, true>, k_dp_tiles<4, 8, MAXT, MINB_C, false, true, false, true>,
                                    k_dp_tiles<4, 8, MAXT, MINB_C, true, false, false, true>, k_dp_tiles<4, 8, MAXT, MINB_C, true, true, false, true>};
  const kern_t kern = (costs_only ? kerns_c : kerns)[(P.nb == 1 ? 2 : 0) + (P.n_min > 1 ? 1 : 0)];
  if (mr_ensure_smem(kern, p.smem)) return 1;
  const int grid = std::min(count, costs_only ? p.grid_cap_costs : p.grid_cap);
  kern<<<grid, p.threads, p.smem, stream>>>(P, p.layout, p.nslots, p.steps, p.d_tiles, A, B, orient,
                                           surv, surv_count, tb, cost);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Launch-bound buckets (CTA size, CTAs per SM, ping-pong form): the register budget is
// 65536 / (MAXT * MINB).  The two-min form of the 4x8 tile update needs ~88 registers, the
// ping-pong form (one FMNMX3 per cell) ~150.  Small pairs have few tiles: several CTAs
// (orientations) per SM keep the pipes fed; from 320 threads on, one CTA per SM in the
// ping-pong form is faster than two in the two-min form (config 2: 2.07 vs 2.01 TCUPS).
#define MR_DP_BIG_BUCKETS(X) X(320, 1, true) X(352, 1, true) X(384, 1, true)
#define MR_DP_BUCKETS(X) \
  X(64, 10, false) X(128, 5, false) X(192, 3, false) X(256, 2, false) MR_DP_BIG_BUCKETS(X) X(512, 1, true)
static inline int mr_dp_tiles_minb(int threads) {
#define MR_X(MAXT, MINB, PP) if (threads <= MAXT) return MINB;
  MR_DP_BUCKETS(MR_X)
#undef MR_X
  return 1;
}
static inline int mr_dp_tiles_minb_costs(int threads) {
  const int b = mr_dp_tiles_minb(threads);
  return threads <= 384 ? std::max(b, 2) : 1;
}
static inline bool mr_dp_tiles_pp(int threads) {
#define MR_X(MAXT, MINB, PP) if (threads <= MAXT) return PP;
  MR_DP_BUCKETS(MR_X)
#undef MR_X
  return false;
}

// Buckets of the min-with-predicate kernels: the (MAXT, MINB) pairs MR_VP_BUCKET instantiates in
// mr_dp_vp.cu, by name k_dp_vp_<MAXT>_{g,gw,c,cw} (general / one atom per residue, windowed).
// (80 registers per thread, allocated per warp in units of 512: 2560 per warp, 25 warps per SM)
#define MR_VP_BUCKETS(X) X(64, 12) X(128, 6) X(192, 4) X(256, 3) X(320, 2) X(352, 2) X(384, 2)
static inline int mr_dp_vp_bucket(int threads) {
#if MR_HAVE_VP_CUBIN
#define MR_X(MAXT, MINB) if (threads <= MAXT) return MAXT;
  MR_VP_BUCKETS(MR_X)
#undef MR_X
#endif
  (void)threads;
  return 0;
}
static inline int mr_dp_vp_minb(int bucket) {
#define MR_X(MAXT, MINB) if (bucket == MAXT) return MINB;
  MR_VP_BUCKETS(MR_X)
#undef MR_X
  return 1;
}
// 0 launched, 1 failed (*why says what)
static inline int mr_launch_dp_tiles_vp(const MrTiledPlan& p, const MrDevParams& P, const float* A, const float* B,
                                        const float* orient, const int* surv, const int* surv_count, uint32_t* tb,
                                        float* cost, int count, cudaStream_t stream, int device, std::string* why) {
  char name[64];
  snprintf(name, sizeof name, "k_dp_vp_%d_%s%s", p.vp_bucket, P.nb == 1 ? "c" : "g", P.n_min > 1 ? "w" : "");
  if (!p.vp_fn) p.vp_fn = mr_vp_function(device, name, p.smem, why);
  CUfunction f = static_cast<CUfunction>(p.vp_fn);
  if (!f) return 1;
  MrDevParams Pv = P;
  TbTiled Lv = p.layout;
  int nslots = p.nslots, nsteps = p.steps;
  const MrTile* tiles = p.d_tiles;
  void* args[] = {&Pv, &Lv, &nslots, &nsteps, &tiles, &A, &B, &orient, &surv, &surv_count, &tb, &cost};
  const int grid = std::min(count, p.grid_cap_vp);
  const CUresult r = mr_vp_driver().LaunchKernel(f, grid, 1, 1, p.threads, 1, 1, (unsigned)p.smem, (CUstream)stream, args, nullptr);
  if (r != CUDA_SUCCESS) { *why = std::string("cuLaunchKernel(") + name + "): " + mr_vp_errstr(r); return 1; }
  return 0;
}

static inline int mr_launch_dp_tiled(const MrTiledPlan& p, const MrDevParams& P, const float* A,
                                     const float* B, const float* orient, const int* surv,
                                     const int* surv_count, uint32_t* tb, float* cost, int count,
                                     cudaStream_t stream, bool costs_only = false) {
#define MR_X(MAXT, MINB, PP) \
  if (p.threads <= MAXT) return mr_launch_dp_tiles_t<MAXT, MINB, PP>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream, costs_only);
  MR_DP_BUCKETS(MR_X)
#undef MR_X
  return mr_launch_dp_tiles_t<1024, 1, false>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream, costs_only);
}
#endif  // MR_DP_TILED_DEVICE_ONLY
