// mr_dp_tiled_mp.cuh -- multi-pass form of the register-tiled 3-D DP kernel (mr_dp_tiled.cuh)
// for problems whose (layer,row) cross-section has more live 4x8 tiles than one CTA has
// threads (roughly: chains longer than ~190 residues).
//
// Same recurrence, tie order, tile shape, exchange-slot protocol and traceback word format
// as k_dp_tiles (Dyn3d::Solve_IncrN, lib/dyn/dyn3d.cc:183-387).  The row strips are cut into
// PASSES of consecutive strips [w0, w1); a CTA still owns one orientation and sweeps the
// columns once per pass, top strips first.  Between passes only the bottom row of the pass's
// last strip, A(n, i_hi, j) for every layer and column, has to survive: it is written to a
// per-CTA carry buffer in global memory (carry[j * CN + chunk] = 4 layers as one float4,
// L2-resident) and read back, one column ahead, as the upper boundary of the first strip of
// the next pass.  The buffer is updated in place: the reader of column j runs at step j-1,
// the writer of column j at step j + (w1-1-w0), separated by at least one barrier.
// (A 4-deep cp.async ring instead of the one-column register prefetch was measured 3-6 %
// slower at 300..2000 residues: the carry latency is not what bounds the kernel.)
// A reader takes carry values only inside its own column window (outside it substitutes the
// boundary constants, like an exchange slot that was never written), and the window of a
// writer is widened to cover its reader's, so every value read was written for this
// orientation.
//
// Global tile ids (pass offset + thread) index the same traceback layout as the single-pass
// kernel (TbTiled with T = tiles of all passes), so K5 (k_traceback_kabsch) is unchanged.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "mr_dp_tiled.cuh"

struct MrPass {
  int w0, nstrips;   // first strip and number of strips of the pass
  int nsteps;        // columns (even) + strips - 1
  int gid_off;       // global id of thread 0's tile
  int lag;           // split mode: step s of this pass runs at tick lag + s
  int ticks;         // split mode: ticks of the whole orientation (max over passes of lag + nsteps)
};

// Split mode (template parameter SPLIT): the passes of ONE orientation run concurrently on the
// CTAs of a thread-block cluster (CTA rank = pass), pipelined through the carry buffer: pass p
// runs lag_p = sum over the passes above it of (strips + MR_SPLIT_GROUP) steps behind pass 0, and
// every MR_SPLIT_GROUP-th per-step barrier is a cluster barrier (release / acquire, which also
// orders the L2 carry traffic): a carry column is written at least one cluster barrier before
// the pass below prefetches it.  The sweep of
// an orientation then takes columns + all strips (+ passes) ticks instead of passes x (columns
// + strips): the latency of a single evaluation, which is what Refine batches wait for.
// Throughput per SM is lower (cluster barriers, idle head and tail ticks): FirstPass keeps one
// CTA per orientation.
#define MR_SPLIT_GROUP 8  // ticks between cluster barriers (a cluster barrier costs ~0.8 us, a CTA barrier ~0.1)
__device__ __forceinline__ void mr_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned mr_cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned mr_cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

struct MrTileMp {
  int w, c;              // global row strip, layer chunk (-1: idle thread)
  int s_lo, s_hi;        // first / last step of the pass at which this tile runs
  int up_slot, dg_slot;  // exchange slots (up_slot = -1: upper boundary comes from the carry buffer)
  int warp_lo, warp_hi;  // step window of the whole warp
  int out;               // 1: publishes its bottom row to the carry buffer
  int pad_;
};

// (The min-with-predicate tile update of the single-pass kernel was tried here at 384 x 2 and
// 320 x 2 CTAs per SM and lost against 512 x 1 ping-pong: more, thinner passes, and spills in the
// column loop at 80 registers -- 0.9-1.4 against 2.0 TCUPS, profiles/README.md.)
template <int K, int R, int MAXT, int MINB, bool NB1, bool WINDOWED, bool PP, bool SPLIT = false, bool COSTS = false>
__global__ void __launch_bounds__(MAXT, MINB)
k_dp_tiles_mp(MrDevParams P, TbTiled L, int nslots, int npasses, int wl_max, const MrPass* __restrict__ passes,
              const MrTileMp* __restrict__ tiles, const float* __restrict__ A, const float* __restrict__ B,
              const float* __restrict__ orient, const int* __restrict__ surv,
              const int* __restrict__ surv_count, float4* carry_all, size_t carry_pitch,
              uint32_t* __restrict__ tb, float* __restrict__ cost_out) {
  static_assert(R == MR_TR && K == MR_TK, "traceback word packing assumes 4 x 8 tiles");
  extern __shared__ float4 smem4[];
  const int m = P.m, nB = P.nB, nb = P.nb;
  const int W = L.W, pad = L.pad, CN = L.CN;
  const int nBe = 2 * L.JP;
  const int WRp = mr_ring_stride(wl_max * R);  // ring row stride (floats), = 4 mod 32
  const int nBp = (nB * nb + 11) & ~3;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  float4* exch = smem4;
  float* Cr = reinterpret_cast<float*>(exch + 6 * nslots);
  float* sB = Cr + MR_CRING * WRp;
  float* sA = sB + 3 * nBp;
  const float INF = __int_as_float(0x7f800000);
  const int nsurv = *surv_count;
  // split mode: the cluster (min(passes, 16) CTAs) owns the orientation slot, the CTA rank is the pass
  // (more passes than CTAs in the cluster: CTA r takes passes r, r + cluster size, ... as a ring)
  const int csize = SPLIT ? (int)mr_cluster_nctarank() : 1;
  const int slot = (int)blockIdx.x / csize;
  const int nslots_grid = (int)gridDim.x / csize;
  const int ps_first = SPLIT ? (int)mr_cluster_ctarank() : 0;
  float4* carry = carry_all + (size_t)slot * carry_pitch;

  const uint32_t exch0 = (uint32_t)__cvta_generic_to_shared(exch);
  const uint32_t par_bytes = (uint32_t)nslots * 48u;
  const uint32_t a_me = exch0 + (uint32_t)tid * 48u;
  const uint32_t c_row = (uint32_t)WRp * 4u;
  const uint32_t cr0 = (uint32_t)__cvta_generic_to_shared(Cr);
  for (int k = tid; k < m * nb * 3; k += nthreads) sA[k] = A[k];

// split mode: a cluster barrier every MR_SPLIT_GROUP ticks (tick = lag + step), CTA barriers in between
#define MR_MP_SYNC(S) do { if (SPLIT && ((tick0 + (S)) & (MR_SPLIT_GROUP - 1)) == 0) mr_cluster_sync(); else __syncthreads(); } while (0)
  for (int sv = slot; sv < nsurv; sv += nslots_grid) {
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
    uint4* const tb_o = reinterpret_cast<uint4*>(tb + (size_t)sv * L.words(nB));

    int cur_tick = 0;  // split mode: ticks this CTA has been through for this orientation
    for (int ps = ps_first; ps < npasses; ps += csize) {
      const MrPass pass = passes[ps];
      const MrTileMp* ptiles = tiles + (size_t)ps * nthreads;
      const MrTileMp me = ptiles[tid];
      MR_ASSERT(me.up_slot < nslots && me.dg_slot >= 0 && me.dg_slot < nslots && (size_t)(nBe + 1) * CN + (CN - 1) < carry_pitch);
      const bool has_tile = me.w >= 0;
      const int w = has_tile ? me.w : pass.w0, c = has_tile ? me.c : 0;
      const int wl = w - pass.w0;  // strip within the pass: this tile runs wl columns behind
      const int nsteps = pass.nsteps;
      const int WR = pass.nstrips * R, rho0 = pass.w0 * R;
      const int warp_lo = me.warp_lo, warp_hi = min(me.warp_hi, nsteps);
      const int tb_hi = has_tile ? me.s_hi : -1;
      const int tb_first = has_tile ? me.s_lo + ((me.s_lo - wl) & 1) : 0x3fffffff;
      const int cost_step = (has_tile && w == W - 1) ? nB + wl : -1;
      const bool from_carry = has_tile && me.up_slot < 0;
      const int out_lo = (has_tile && me.out) ? me.s_lo : 0x3fffffff;
      const unsigned own_span = has_tile ? (unsigned)(me.s_hi - me.s_lo) : 0u;
      const uint32_t a_up = exch0 + (uint32_t)max(me.up_slot, 0) * 48u + 32u;
      const uint32_t a_dg = exch0 + (uint32_t)me.dg_slot * 48u;
      const uint32_t a_c = cr0 + (uint32_t)(wl * R) * 4u;
      float4* const cme = carry + c;

      __syncthreads();  // previous pass (or the rotation) is complete
      for (int k = tid; k < 2 * nslots; k += nthreads) {
        const int sl = k % nslots;
        const bool zero = sl == nslots - 2 || (sl < nslots - 2 && ptiles[sl].c == 0);
        exch[3 * k] = make_float4(INF, INF, INF, INF);
        exch[3 * k + 1] = make_float4(INF, INF, INF, INF);
        exch[3 * k + 2] = make_float4(zero ? 0.0f : INF, INF, INF, INF);
      }
      for (int k = tid; k < MR_CRING * WRp; k += nthreads) Cr[k] = INF;
      __syncthreads();
      mr_produce_c<NB1>(Cr, sA, sB, 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads, rho0);
      const int tick0 = pass.lag;
      if constexpr (SPLIT) {
        __syncthreads();
        for (int t = cur_tick + 1; t <= pass.lag; t++)  // the passes above are that far ahead
          if ((t & (MR_SPLIT_GROUP - 1)) == 0) mr_cluster_sync();
      }

      // tile state and step: as in k_dp_tiles (mr_tile_step.cuh); the upper boundary of a pass's
      // first strip comes from the carry buffer (fetched one column ahead into a register), and
      // the bottom row of its last strip goes there
      static_assert(!(COSTS && PP), "the costs-only update is in place");
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
      uint4* tbp = tb_o + (pass.gid_off + tid) + (size_t)max(0, (tb_first - wl - 1) >> 1) * L.T;
      int tb_next = tb_first;
      float* const cost_row = cost_out + (size_t)sv * L.NP;
      __syncthreads();
      int s = 1;
      for (; s < warp_lo && s <= nsteps; s++) {
        if ((s & (MR_CBATCH - 1)) == 0)
          mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads, rho0);
        MR_MP_SYNC(s);
      }
      float4 gnext = make_float4(INF, INF, INF, INF);
      if (from_carry) gnext = __ldcg(cme + (size_t)min(s, nBe + 1) * CN);
      auto step = [&](auto src_tag, const int s, const uint32_t rd_off, const uint32_t wr_off) {
        constexpr int S = decltype(src_tag)::value, D = PP ? 1 - S : 0;
        float cv[R], dsh[R];
        mr_tile_load8(cv, a_c + (uint32_t)((s - wl) & (MR_CRING - 1)) * c_row);
        float4 ub = mr_lds128(a_up + rd_off);
        if (from_carry) {
          ub = ((unsigned)(s - me.s_lo) <= own_span) ? gnext : make_float4(c == 0 ? 0.0f : INF, INF, INF, INF);
          gnext = __ldcg(cme + (size_t)min(s + 1, nBe + 1) * CN);
        }
        mr_tile_load8(dsh, a_dg + rd_off);
        if constexpr (PP) {
          bprev[D][0] = ub.x; bprev[D][1] = ub.y; bprev[D][2] = ub.z; bprev[D][3] = ub.w;
          mr_tile_update_pp<S>(st, bprev, acc, cv, dsh);
        } else {
          const float bin[K] = {ub.x, ub.y, ub.z, ub.w};
          if constexpr (COSTS) mr_tile_update_costs(st[0], bprev[0], bin, cv, dsh);
          else mr_tile_update_ip(st[0], bprev[0], bin, acc, cv, dsh);
        }
        mr_tile_publish(a_me + wr_off, bprev[D][K - 1], st[D][K - 1], st[D][0][R - 1], st[D][1][R - 1], st[D][2][R - 1],
                        st[D][3][R - 1]);
        // bottom row of the pass -> carry[column j = s - wl]
        MR_ASSERT(!((unsigned)(s - out_lo) <= own_span) || (s - wl >= 0 && (size_t)(s - wl) * CN + c < carry_pitch));
        if ((unsigned)(s - out_lo) <= own_span)
          __stcg(cme + (size_t)(s - wl) * CN, make_float4(st[D][0][R - 1], st[D][1][R - 1], st[D][2][R - 1], st[D][3][R - 1]));
        if (!COSTS && s == tb_next) {
          MR_ASSERT(tbp >= tb_o && tbp + 1 <= tb_o + L.words(nB) / 4);
          *tbp = make_uint4(acc[0], acc[1], acc[2], acc[3]);
          tbp += L.T;
          tb_next = (s + 2 <= tb_hi) ? s + 2 : 0x3fffffff;
        }
        if (s == cost_step) {
#pragma unroll
          for (int k = 0; k < K; k++) {
            const int n = c * K + k;
            if (n >= P.n_min && n <= P.n_max) cost_row[n - 1] = st[D][k][R - 1];
          }
        }
        if (WINDOWED && has_tile) mr_tile_window_cost(P, st[D], s - wl, w, c, pad, cost_row);
        if ((s & (MR_CBATCH - 1)) == 0)
          mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads, rho0);
        MR_MP_SYNC(s);
      };
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
      for (; s <= nsteps; s++) {
        if ((s & (MR_CBATCH - 1)) == 0)
          mr_produce_c<NB1>(Cr, sA, sB, s + 1, WR, WRp, pad, nB, nBp, nb, P.dp_cutoff, tid, nthreads, rho0);
        MR_MP_SYNC(s);
      }
      cur_tick = pass.lag + nsteps;
      if constexpr (SPLIT)
        if (ps + csize >= npasses)  // my last pass: the passes below finish
          for (int t = cur_tick + 1; t <= pass.ticks; t++)
            if ((t & (MR_SPLIT_GROUP - 1)) == 0) mr_cluster_sync();
    }
  }
#undef MR_MP_SYNC
}

// ---------------------------------------------------------------------------------- host
struct MrTiledMpPlan {
  bool ok = false;
  const char* why = "not planned";
  int W = 0, CN = 0, nBe = 0;
  int threads = 0, minb = 1, nslots = 0, npasses = 0, wl_max = 0;
  bool pp = false;
  size_t smem = 0;
  int grid_cap = 0;
  size_t tb_words = 0, carry_pitch = 0;
  double lane_efficiency = 0.0;
  int split_ticks = 0;   // split mode: ticks per orientation
  int split_csize = 0;   // split mode: CTAs per cluster = min(passes, 16)
  int split_cap = 0;     // split mode: orientations (clusters) resident at once; 0 = not available
  TbTiled layout;
  MrPass* d_passes = nullptr;
  MrTileMp* d_tiles = nullptr;
  int* d_tile_of = nullptr;
  float4* d_carry = nullptr;
};

static inline void mr_free_dp_tiled_mp(MrTiledMpPlan* p) {
  if (p->d_passes) cudaFree(p->d_passes);
  if (p->d_tiles) cudaFree(p->d_tiles);
  if (p->d_tile_of) cudaFree(p->d_tile_of);
  if (p->d_carry) cudaFree(p->d_carry);
  p->d_passes = nullptr; p->d_tiles = nullptr; p->d_tile_of = nullptr; p->d_carry = nullptr;
}

static inline void mr_plan_dp_tiled_mp(MrTiledMpPlan* p, const MrDevParams& P, const cudaDeviceProp& prop, int NP,
                                       int max_threads = 0) {
  p->ok = false;
  const int R = 8, K = 4;
  const int m = P.m, nB = P.nB;
  const int W = (m + R - 1) / R, pad = W * R - m;
  const int CN = (P.n_max + 1 + K - 1) / K;
  const int nBe = (nB + 1) & ~1;
  const int WL_MAX = MR_CRING - MR_CBATCH;  // strips per pass the C ring can hold
  struct T { int w, c, jlo, jhi; };         // window in columns
  std::vector<T> all;
  std::vector<int> have((size_t)W * CN, -1);
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
      if (c == 0) { jlo = std::min(jlo, 1); jhi = std::max(jhi, 1); }
      if (jhi < 0) continue;
      jhi = (jhi + 1) & ~1;
      have[(size_t)w * CN + c] = (int)all.size();
      all.push_back({w, c, jlo, jhi});
    }
  }
  // closure (see mr_plan_dp_tiled): tile (w,c-1) relays the corner value tile (w,c) needs
  {
    const size_t n0 = all.size();
    for (size_t t = 0; t < n0; t++) {
      const T x = all[t];
      if (x.c >= 1 && have[(size_t)x.w * CN + x.c - 1] < 0) {
        have[(size_t)x.w * CN + x.c - 1] = (int)all.size();
        all.push_back({x.w, x.c - 1, std::max(1, x.jlo - 1), x.jhi});
      }
    }
  }
  std::vector<int> per_strip(W, 0);
  for (const T& x : all) per_strip[x.w]++;
  const int widest = *std::max_element(per_strip.begin(), per_strip.end());
  // 512 threads, one CTA per SM (128 registers per thread: the tile code does not spill);
  // measured on B200 against 320 x 2, 352 x 2, 640 and 768 at 300..2000 residues
  int threads = 512, minb = 1;
  if (widest > 512) threads = 640;
  if (widest > 640) threads = 768;
  if (widest > 768) threads = 1024;
  if (max_threads > 0 && widest <= max_threads) { threads = max_threads; minb = threads <= 384 ? 2 : 1; }
  if (widest > 1024) { p->why = "more than 1024 live tiles in one row strip"; return; }
  // passes: consecutive strips while the tiles fit one CTA
  std::vector<MrPass> passes;
  // as few passes as the CTA size allows, with the tiles spread evenly over them (a last pass
  // with a handful of warps runs as long as a full one: it is latency-bound)
  int remaining = 0;
  for (int w = 0; w < W; w++) remaining += per_strip[w];
  for (int w0 = 0; w0 < W;) {
    const int passes_left = (remaining + threads - 1) / threads;
    const int target = (remaining + passes_left - 1) / passes_left;
    int w1 = w0, cnt = 0;
    while (w1 < W && w1 - w0 < WL_MAX && cnt + per_strip[w1] <= threads &&
           (w1 == w0 || cnt + per_strip[w1] / 2 <= target))
      cnt += per_strip[w1++];
    remaining -= cnt;
    MrPass ps;
    ps.w0 = w0; ps.nstrips = w1 - w0; ps.nsteps = nBe + (w1 - w0) - 1; ps.gid_off = 0; ps.lag = 0; ps.ticks = 0;
    passes.push_back(ps);
    w0 = w1;
  }
  const int npasses = (int)passes.size();
  const bool pp = threads == 512;  // ping-pong tile update: fits 128 registers only in the 512-thread bucket
  // a writer (last strip of a pass) covers the window of its reader (first strip of the next
  // pass, same chunk); bottom-up, so chains of one-strip passes are covered too
  for (int q = npasses - 2; q >= 0; q--) {
    const int wlast = passes[q].w0 + passes[q].nstrips - 1;
    for (int c = 0; c < CN; c++) {
      const int a = have[(size_t)wlast * CN + c], b = have[(size_t)(wlast + 1) * CN + c];
      if (a < 0 || b < 0) continue;
      all[a].jlo = std::min(all[a].jlo, all[b].jlo);
      all[a].jhi = std::max(all[a].jhi, all[b].jhi);
    }
  }
  const int nslots = threads + 2;
  std::vector<MrTileMp> tab((size_t)npasses * threads);
  std::vector<int> index((size_t)W * CN, 0);
  int gid = 0, wl_max = 0;
  long long warp_steps = 0;
  for (int q = 0; q < npasses; q++) {
    MrPass& ps = passes[q];
    const int w0 = ps.w0, w1 = ps.w0 + ps.nstrips;
    wl_max = std::max(wl_max, ps.nstrips);
    struct S { int w, c, lo, hi, out; };
    std::vector<S> ts;
    for (const T& x : all) {
      if (x.w < w0 || x.w >= w1) continue;
      const int wl = x.w - w0;
      const bool out = (x.w == w1 - 1) && (q + 1 < npasses) && have[(size_t)(x.w + 1) * CN + x.c] >= 0;
      ts.push_back({x.w, x.c, x.jlo + wl, x.jhi + wl, out ? 1 : 0});
    }
    std::stable_sort(ts.begin(), ts.end(), [](const S& a, const S& b) { return a.lo != b.lo ? a.lo < b.lo : a.hi < b.hi; });
    const int ntiles = (int)ts.size();
    std::vector<int> local((size_t)W * CN, -1);
    for (int t = 0; t < ntiles; t++) {
      local[(size_t)ts[t].w * CN + ts[t].c] = t;
      index[(size_t)ts[t].w * CN + ts[t].c] = gid + t;
    }
    ps.gid_off = gid;
    const int used = ((ntiles + 31) / 32) * 32;
    gid += used;
    for (int wp = 0; wp < threads / 32; wp++) {
      int lo = 1 << 30, hi = -1;
      for (int t = wp * 32; t < std::min(ntiles, wp * 32 + 32); t++) { lo = std::min(lo, ts[t].lo); hi = std::max(hi, ts[t].hi); }
      if (pp && hi >= lo && ((hi - lo + 1) & 1)) { if (lo > 1) lo--; else hi++; }  // steps come in pairs
      if (pp) ps.nsteps = std::max(ps.nsteps, hi);
      if (hi >= lo) warp_steps += hi - lo + 1;
      for (int t = wp * 32; t < wp * 32 + 32; t++) {
        MrTileMp& e = tab[(size_t)q * threads + t];
        e.warp_lo = lo; e.warp_hi = hi; e.out = 0; e.pad_ = 0;
        if (t >= ntiles) { e.w = -1; e.c = -1; e.s_lo = 1 << 30; e.s_hi = -1; e.up_slot = nslots - 1; e.dg_slot = nslots - 1; continue; }
        e.w = ts[t].w; e.c = ts[t].c; e.s_lo = ts[t].lo; e.s_hi = ts[t].hi; e.out = ts[t].out;
        int up;
        if (e.w > w0) up = local[(size_t)(e.w - 1) * CN + e.c];
        else up = (e.w > 0 && have[(size_t)(e.w - 1) * CN + e.c] >= 0) ? -2 : -1;  // -2: carry
        const int dg = (e.c > 0) ? local[(size_t)e.w * CN + e.c - 1] : -1;
        e.up_slot = up >= 0 ? up : (up == -2 ? -1 : ((e.c == 0) ? nslots - 2 : nslots - 1));
        e.dg_slot = dg >= 0 ? dg : nslots - 1;
      }
    }
  }
  const int T_total = gid;
  // split mode: pass q runs lag_q steps behind pass 0: per pass above it, its strips (a column
  // leaves the pass that many steps after it entered) + the ticks between two cluster barriers
  // With more than 16 passes the cluster (16 CTAs) works as a ring: CTA r takes passes r, r + 16,
  // ..., and a pass also waits until its CTA has finished the previous one.
  const int split_csize = std::min(npasses, 16);
  int split_ticks = 0;
  {
    int lag = 0;
    std::vector<int> free_at(split_csize, 0);  // tick after which CTA r is idle
    for (int q = 0; q < npasses; q++) {
      const int r = q % split_csize;
      lag = std::max(lag, free_at[r] + (q >= split_csize ? MR_SPLIT_GROUP : 0));
      passes[q].lag = lag;
      free_at[r] = lag + passes[q].nsteps;
      split_ticks = std::max(split_ticks, lag + passes[q].nsteps);
      lag += passes[q].nstrips + MR_SPLIT_GROUP;
    }
    split_ticks = (split_ticks + MR_SPLIT_GROUP - 1) & ~(MR_SPLIT_GROUP - 1);  // ends on a cluster barrier
    for (int q = 0; q < npasses; q++) passes[q].ticks = split_ticks;
  }
  const size_t smem = 16 * (size_t)(6 * nslots) +
                      4 * (size_t)(MR_CRING * mr_ring_stride(wl_max * R) + 3 * ((nB * P.nb + 11) & ~3) + ((m * P.nb * 3 + 3) & ~3));
  if (smem > prop.sharedMemPerBlockOptin) { p->why = "work space does not fit shared memory"; return; }
  const int by_smem = (int)(prop.sharedMemPerMultiprocessor / (smem + 1024));
  const int ctas_per_sm = std::max(1, std::min(by_smem, minb));
  const int grid_cap = prop.multiProcessorCount * ctas_per_sm;
  const size_t carry_pitch = (size_t)(nBe + 2) * CN;
  mr_free_dp_tiled_mp(p);
  auto fail = [&](const char* why) { mr_free_dp_tiled_mp(p); p->why = why; };
  if (cudaMalloc(&p->d_passes, sizeof(MrPass) * passes.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(p->d_passes, passes.data(), sizeof(MrPass) * passes.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy");
  if (cudaMalloc(&p->d_tiles, sizeof(MrTileMp) * tab.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(p->d_tiles, tab.data(), sizeof(MrTileMp) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy");
  if (cudaMalloc(&p->d_tile_of, sizeof(int) * index.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(p->d_tile_of, index.data(), sizeof(int) * index.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy");
  if (cudaMalloc(&p->d_carry, sizeof(float4) * carry_pitch * grid_cap) != cudaSuccess) { cudaGetLastError(); return fail("cudaMalloc (carry buffer)"); }
  if (cudaMemset(p->d_carry, 0x7f, sizeof(float4) * carry_pitch * grid_cap) != cudaSuccess) return fail("cudaMemset");
  p->W = W; p->CN = CN; p->nBe = nBe;
  p->pp = pp;
  p->threads = threads; p->minb = minb; p->nslots = nslots; p->npasses = npasses; p->wl_max = wl_max;
  p->smem = smem; p->grid_cap = grid_cap; p->carry_pitch = carry_pitch;
  p->split_ticks = split_ticks;
  // clusters of up to 8 CTAs are portable; one CTA per SM; GPCs of a B200 hold 16-20 SMs, so a
  // conservative count of co-resident clusters is enough (more would only queue)
  // (9..16 passes: non-portable cluster size, opt-in per kernel; fewer such clusters fit a GPC)
  p->split_csize = split_csize;
  p->split_cap = (threads == 512 && npasses >= 2)
                     ? std::max(1, (split_csize <= 8 ? grid_cap * 3 / 4 : grid_cap / 2) / split_csize) : 0;
  p->layout.K = K; p->layout.R = R; p->layout.W = W; p->layout.JP = nBe / 2;
  p->layout.pad = pad; p->layout.NP = NP; p->layout.CN = CN; p->layout.T = T_total;
  p->layout.tile_of = p->d_tile_of;
  p->tb_words = p->layout.words(nB);
  p->lane_efficiency = (double)live_cells / ((double)warp_steps * 32 * K * R);
  p->ok = true;
  p->why = "";
  if (getenv("MR_DEBUG_PLAN")) {
    fprintf(stderr, "[mr] multi-pass plan: %d x %d, layers %d: %d threads x %d CTAs/SM, %d passes (", m, nB, P.n_max,
            threads, ctas_per_sm, npasses);
    for (int q = 0; q < npasses; q++) fprintf(stderr, "%s%d", q ? "," : "", passes[q].nstrips);
    fprintf(stderr, " strips), %d tiles, lane efficiency %.3f, smem %zu B, carry %.1f MB/CTA, traceback %.1f MB/orientation\n",
            T_total, p->lane_efficiency, smem, carry_pitch * 16.0 / 1048576.0, p->tb_words * 4.0 / 1048576.0);
  }
}

template <int MAXT, int MINB, bool PP>
static inline int mr_launch_dp_tiles_mp_t(const MrTiledMpPlan& p, const MrDevParams& P, const float* A,
                                          const float* B, const float* orient, const int* surv,
                                          const int* surv_count, uint32_t* tb, float* cost, int count,
                                          cudaStream_t stream, bool split = false, bool costs_only = false) {
  typedef void (*kern_t)(MrDevParams, TbTiled, int, int, int, const MrPass*, const MrTileMp*, const float*,
                         const float*, const float*, const int*, const int*, float4*, size_t, uint32_t*, float*);
  if constexpr (MAXT == 512 && PP) {
    if (split && p.split_cap > 0) {
      // one cluster of min(passes, 16) CTAs per orientation (see the comment on split mode above)
      static const kern_t skerns[4] = {k_dp_tiles_mp<4, 8, MAXT, MINB, false, false, PP, true>, k_dp_tiles_mp<4, 8, MAXT, MINB, false, true, PP, true>,
                                       k_dp_tiles_mp<4, 8, MAXT, MINB, true, false, PP, true>, k_dp_tiles_mp<4, 8, MAXT, MINB, true, true, PP, true>};
      const kern_t kern = skerns[(P.nb == 1 ? 2 : 0) + (P.n_min > 1 ? 1 : 0)];
      if (mr_ensure_smem(kern, p.smem)) return 1;
      if (p.split_csize > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(std::min(count, p.split_cap) * p.split_csize));
      cfg.blockDim = dim3((unsigned)p.threads);
      cfg.dynamicSmemBytes = p.smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)p.split_csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P, p.layout, p.nslots, p.npasses, p.wl_max, (const MrPass*)p.d_passes,
                                               (const MrTileMp*)p.d_tiles, A, B, orient, surv, surv_count, p.d_carry,
                                               p.carry_pitch, tb, cost);
      return e == cudaSuccess ? 0 : 1;
    }
  }
  static const kern_t kerns[4] = {k_dp_tiles_mp<4, 8, MAXT, MINB, false, false, PP>, k_dp_tiles_mp<4, 8, MAXT, MINB, false, true, PP>,
                                  k_dp_tiles_mp<4, 8, MAXT, MINB, true, false, PP>, k_dp_tiles_mp<4, 8, MAXT, MINB, true, true, PP>};
  static const kern_t kerns_c[4] = {k_dp_tiles_mp<4, 8, MAXT, MINB, false, false, false, false, true>, k_dp_tiles_mp<4, 8, MAXT, MINB, false, true, false, false, true>,
                                    k_dp_tiles_mp<4, 8, MAXT, MINB, true, false, false, false, true>, k_dp_tiles_mp<4, 8, MAXT, MINB, true, true, false, false, true>};
  const kern_t kern = (costs_only ? kerns_c : kerns)[(P.nb == 1 ? 2 : 0) + (P.n_min > 1 ? 1 : 0)];
  if (mr_ensure_smem(kern, p.smem)) return 1;
  const int grid = std::min(count, p.grid_cap);
  kern<<<grid, p.threads, p.smem, stream>>>(P, p.layout, p.nslots, p.npasses, p.wl_max, p.d_passes, p.d_tiles, A, B,
                                           orient, surv, surv_count, p.d_carry, p.carry_pitch, tb, cost);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

static inline int mr_launch_dp_tiled_mp(const MrTiledMpPlan& p, const MrDevParams& P, const float* A,
                                        const float* B, const float* orient, const int* surv,
                                        const int* surv_count, uint32_t* tb, float* cost, int count,
                                        cudaStream_t stream, bool split = false, bool costs_only = false) {
#define MR_TRY(MAXT, MINB, PP) \
  if (p.threads == MAXT) return mr_launch_dp_tiles_mp_t<MAXT, MINB, PP>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream, split && !costs_only, costs_only)
  MR_TRY(512, 1, true);
  MR_TRY(640, 1, false);
  MR_TRY(768, 1, false);
#undef MR_TRY
  return mr_launch_dp_tiles_mp_t<1024, 1, false>(p, P, A, B, orient, surv, surv_count, tb, cost, count, stream, false, costs_only);
}
