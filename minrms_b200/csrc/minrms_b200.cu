// minrms_b200.cu -- C ABI (include/minrms_b200.h) over the sm_100a kernels.
// Host side of the per-orientation search: chunking, streams, events, result tables.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//             -Xcompiler -fPIC -shared  (see __graft_entry__.build()).
#include "../../include/minrms_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <mutex>
#include <vector>

#include "mr_kernels.cuh"
#include "mr_dp_tiled.cuh"
#include "mr_dp_tiled_mp.cuh"

namespace {

std::string g_create_error;

struct EventPair {
  cudaEvent_t a, b;
  int kind;  // 0 dp, 1 nw, 2 tb
};

}  // namespace

struct Arena { char* base = nullptr; size_t size = 0; };

struct mr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop;
  mr_params hp;
  MrDevParams P;
  int m = 0, nB = 0, N = 0, R = 0, NP = 0;
  TbColumns Lc;
  float *dA = nullptr, *dB = nullptr;
  // resident orientation list
  float* d_orient = nullptr;
  int64_t orient_cap = 0, orient_count = 0;
  // per-chunk work space
  int chunk_cap = 0;
  size_t tb_words = 0;  // per orientation (largest layout in use)
  uint32_t* d_tb = nullptr;
  float *d_cost = nullptr, *d_ssd = nullptr;
  int *d_surv = nullptr, *d_surv_count = nullptr, *d_nw = nullptr, *d_nlast = nullptr;
  int* surv = nullptr;           // survivor list the DP / traceback / select kernels read (d_surv or a window of d_surv_all)
  int *d_surv_all = nullptr, *d_nw_all = nullptr;  // filter results over a whole orientation list
  int64_t filter_cap = 0;
  struct Arena arena;            // chunk work space (see arena_reserve)
  size_t gcol_bytes = 0;         // column work space of the general DP kernel (lazy)
  unsigned char* d_valid = nullptr;
  float* d_rowcost = nullptr;
  unsigned short* d_rowcnt = nullptr;
  int nw_threads_cap = 0;
  short2* d_scratch = nullptr;
  size_t scratch_threads = 0;
  float* d_gcol = nullptr;
  int dp_grid_cap = 0;
  int use_gcol = 0;
  size_t dp_smem = 0;
  unsigned long long* d_keys = nullptr;
  // mr_search_scores: every compared score of the running search, [count][R], grow-only
  float* d_scores = nullptr;
  int64_t scores_cap = 0;       // floats
  float* scores_out = nullptr;  // non-null while mr_search_scores runs: search_device records into it
  // mr_search_dissimilar: alignments of the lower ranks as residue -> partner tables
  short* d_excl_partner = nullptr;    // [ranks][R][m]
  size_t excl_partner_cap = 0;        // shorts
  unsigned char* d_excl = nullptr;    // [chunk][NP] flags written by the traceback kernel
  size_t excl_cap = 0;
  int excl_ranks = 0, excl_max_common = 0;
  bool excl_on = false;               // true while mr_search_dissimilar runs
  // multi-GPU reduction on the device (mr_keys_snapshot / mr_fetch_rows_device)
  unsigned long long* d_keys_local = nullptr;  // this context's keys before the cross-rank MIN
  int32_t* d_rows = nullptr;                   // [R][2 + 12 + N] int32: key, transform, pairs of the sizes this context owns
  int* d_item_row = nullptr;                   // [R] table row of each emitted item
  // grow-only device scratch of the small entry points (mr_superimpose_batch, mr_plot_stats,
  // mr_gen_orientations): they synchronise before they return, so one buffer serves them all
  char* d_misc = nullptr;
  size_t misc_bytes = 0;
  // winner / eval payload
  float* d_orient_w = nullptr;  // [max(R,1)] transforms
  int* d_widx = nullptr;
  int2* d_items = nullptr;
  float* d_Tout = nullptr;      // [max(R,NP)*12]
  short* d_pairs_out = nullptr; // [max(R,NP)*2N]
  // Refine batches (allocated on first use)
  int ref_kcap = 0;
  short2* d_ref_pairs_in = nullptr;  // [kcap][N]
  int* d_ref_n = nullptr;            // [kcap]
  float* d_ref_orient = nullptr;     // [kcap][12]
  float* d_ref_T = nullptr;          // [kcap][NP][12]
  int ref_slice_warps = 0;           // warps of traceback threads that can cover one item's sizes
  char* h_ref = nullptr;             // pinned staging
  size_t ref_off_n = 0, ref_off_ssd = 0, ref_off_valid = 0, ref_off_nlast = 0, ref_off_T = 0, ref_off_pairs = 0;
  int64_t opt_refine_batch = 0;      // 0: all flagged sizes per batch; k > 0: at most k (1 = sequential)
  int64_t opt_refine_seq_superpose = 0;  // tests: the set-up superposition of a Refine batch by the sequential kernel
  int64_t opt_refine_split = 1;      // Refine batches of the multi-pass DP: passes of an orientation on a cluster of CTAs
  int ref_host_item = -1;            // batch item whose transforms / alignments sit in the pinned buffer
  int* h_pinned = nullptr;      // surv counts per chunk
  int h_pinned_cap = 0;
  long long id_offset = 0;
  bool id_offset_set = false;
  int64_t opt_dp_kernel = 0;    // 0 auto, 1 general, 2 tiled (single pass), 3 tiled (multi-pass)
  int64_t opt_budget_mb = 32768;  // chunk work space; raised automatically (see ensure_chunk) unless set by the caller
  bool budget_is_default = true;
  int64_t opt_chunk = 0;
  int64_t opt_nw_general = 0;   // force the thread-per-orientation NW filter
  // decision-only NW filter (k_nw_cost_*): the optimal cost decides NumMatches >= of_min when the
  // cost intervals of of_min and of_min - 1 matches do not touch (see mr_kernels.cuh)
  bool nw_cost_ok = false;
  float nw_cost_thr = 0.0f;     // pass <=> A(m, nB) <= thr
  double nw_cost_margin = 0.0;
  int64_t opt_nw_cost = 1;      // 0: FirstPass filters with the counting kernels
  int64_t opt_dp_form = 0;      // 0: min-with-predicate kernels where a CTA bucket exists (single-pass problems), 1: two-min / ping-pong forms only
  int64_t opt_dp_only = 0;      // experiments: skip traceback/select (results invalid)
  int64_t opt_costs_only_off = 0;  // tests / measurements: -IR FirstPass through the full kernel
  mr_stats st;
  std::vector<EventPair> events;
  size_t events_used = 0;
  std::vector<std::pair<int, int>> pending_counts;  // (pinned slot, chunk count)
  std::string err;
  MrTiledPlan tiled;
  MrTiledMpPlan tiled_mp;
  int64_t opt_mp_threads = 0;   // tests: threads per CTA of the multi-pass kernel (forces more passes)
  cudaEvent_t user_ev[16] = {};
};

namespace {

#define MR_FAIL(ctx, ...)                               \
  do {                                                  \
    char _b[512];                                       \
    snprintf(_b, sizeof _b, __VA_ARGS__);               \
    (ctx)->err = _b;                                    \
    return 1;                                           \
  } while (0)

#define CU(ctx, expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      MR_FAIL(ctx, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
              cudaGetErrorString(_e));                                                      \
    }                                                                                       \
  } while (0)

const unsigned long long KEY_NONE = ((unsigned long long)0x6aa56fa6u /* bits(1e26f) */ << 32) | 0xffffffffull;

float metric_cutoff(float d, int nb) {  // lib/dyn/distance_metric.cc:76-87
  if (d < 0.0f) return MR_UPPER_BOUND;
  return (d * d) * (float)nb;
}

int64_t dp_cells(int m, int nB, int n_min, int n_max) {
  int wA = m - n_min + 1, wB = nB - n_min + 1;
  int64_t c = 0;
  for (int n = 1; n <= n_max; n++)
    c += (int64_t)(std::min(m, n + wA - 1) - n + 1) * (int64_t)(std::min(nB, n + wB - 1) - n + 1);
  return c;
}

int64_t nw_cells(int m, int nB, int of_min) {
  int w0 = m - of_min + 1, w1 = nB - of_min + 1;
  int64_t c = 0;
  for (int i = 1; i <= m; i++) {
    int jlo = (i <= w0) ? 1 : i - w0 + 1;
    int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
    c += jhi - jlo + 1;
  }
  return c;
}

// at least `bytes` of device scratch, 256-byte aligned, owned by the context (freed in mr_destroy)
int misc_reserve(mr_ctx* c, size_t bytes, char** out) {
  if (bytes > c->misc_bytes) {
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->d_misc) cudaFree(c->d_misc);
    c->d_misc = nullptr;
    c->misc_bytes = 0;
    const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    if (cudaMalloc(&c->d_misc, want) != cudaSuccess) {
      cudaGetLastError();
      MR_FAIL(c, "out of device memory: %zu bytes of scratch", want);
    }
    c->misc_bytes = want;
  }
  *out = c->d_misc;
  return 0;
}

EventPair* next_events(mr_ctx* c, int kind) {
  if (c->events_used == c->events.size()) {
    EventPair e;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return nullptr;
    c->events.push_back(e);
  }
  EventPair* e = &c->events[c->events_used++];
  e->kind = kind;
  return e;
}

// fold finished event pairs and pinned survivor counts into the stats (stream must be idle)
void harvest(mr_ctx* c) {
  for (size_t k = 0; k < c->events_used; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->events[k].a, c->events[k].b) == cudaSuccess) {
      if (c->events[k].kind == 0) { c->st.dp_kernel_ms += ms; c->st.dp_kernel_launches++; }
      else if (c->events[k].kind == 1) c->st.nw_kernel_ms += ms;
      else c->st.tb_kernel_ms += ms;
    }
  }
  c->events_used = 0;
  const int64_t per = dp_cells(c->m, c->nB, c->hp.n_min, c->hp.n_max);
  for (auto& pc : c->pending_counts) {
    c->st.orientations_dp += c->h_pinned[pc.first];
    c->st.dp_cells += per * (int64_t)c->h_pinned[pc.first];
  }
  c->pending_counts.clear();
}

int ensure_orient_cap(mr_ctx* c, int64_t count) {
  if (count <= c->orient_cap) return 0;
  if (c->d_orient) cudaFree(c->d_orient);
  c->d_orient = nullptr;
  c->orient_cap = 0;
  CU(c, cudaMalloc(&c->d_orient, sizeof(float) * 12 * (size_t)count));
  c->orient_cap = count;
  return 0;
}

// DP kernel in use: 0 general (k_dp_columns), 1 tiled single-pass, 2 tiled multi-pass
int dp_path(const mr_ctx* c) {
  if (c->opt_dp_kernel == 1) return 0;
  if (c->opt_dp_kernel == 3) return c->tiled_mp.ok ? 2 : 0;
  if (c->opt_dp_kernel == 2) return c->tiled.ok ? 1 : 0;
  // more than 512 live tiles: the single-pass kernel would run 1024-thread CTAs at 64
  // registers (spills); two passes of 512 threads are faster (190x190: 1.6 vs 2.0 TCUPS)
  if (c->tiled.ok && !(c->tiled.threads > 512 && c->tiled_mp.ok)) return 1;
  return c->tiled_mp.ok ? 2 : (c->tiled.ok ? 1 : 0);
}
bool use_tiled(const mr_ctx* c) { return dp_path(c) != 0; }
const TbTiled& tiled_layout(const mr_ctx* c) { return dp_path(c) == 2 ? c->tiled_mp.layout : c->tiled.layout; }
static void plan_mp(mr_ctx* c, int max_threads = 0) { mr_plan_dp_tiled_mp(&c->tiled_mp, c->P, c->prop, c->NP, max_threads); }

// The per-chunk work space (tens of GB; ~100 GB for long chains) is ONE device allocation, the
// arena, carved up by ensure_chunk.  A destroyed context parks its arena in a per-device cache
// and the next context on that device picks it up: in an all-vs-all run cudaMalloc +
// cudaFree of 32 GB cost ~0.5 s per pair, as much as the search itself.
static std::mutex g_arena_mu;
constexpr int kArenaSlots = 4;  // per device: contexts of concurrent host threads (all-vs-all workers) each park one
static Arena g_arena_cache[64][kArenaSlots];

static const bool g_arena_dbg = getenv("MR_DEBUG_PLAN") != nullptr;
#define MR_ARENA_DBG(...) do { if (g_arena_dbg) fprintf(stderr, __VA_ARGS__); } while (0)
static void arena_release(mr_ctx* c) {  // stream must be idle
  if (!c->arena.base) return;
  MR_ARENA_DBG("[mr] arena_release: %zu MB\n", c->arena.size >> 20);
  std::lock_guard<std::mutex> lk(g_arena_mu);
  Arena* slots = g_arena_cache[c->device & 63];
  Arena* smallest = &slots[0];
  for (int k = 1; k < kArenaSlots; k++)
    if (slots[k].size < smallest->size) smallest = &slots[k];
  if (smallest->size < c->arena.size) std::swap(*smallest, c->arena);
  if (c->arena.base) { MR_ARENA_DBG("[mr] arena_release: cudaFree %zu MB\n", c->arena.size >> 20); cudaFree(c->arena.base); }
  c->arena = Arena();
}
static size_t arena_cached_bytes(const mr_ctx* c) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  size_t sum = 0;  // all of it is given back when a larger arena has to be made
  for (int k = 0; k < kArenaSlots; k++) sum += g_arena_cache[c->device & 63][k].size;
  return sum;
}
// make c->arena at least `bytes` large (contents are lost)
static int arena_reserve(mr_ctx* c, size_t bytes) {
  if (c->arena.size >= bytes) return 0;
  cudaStreamSynchronize(c->stream);
  {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    Arena* slots = g_arena_cache[c->device & 63];
    Arena* fit = nullptr;  // smallest parked arena that is large enough
    for (int k = 0; k < kArenaSlots; k++)
      if (slots[k].size >= bytes && (!fit || slots[k].size < fit->size)) fit = &slots[k];
    if (fit) {
      std::swap(*fit, c->arena);  // ours (smaller or none) is parked instead
      MR_ARENA_DBG("[mr] arena_reserve %zu MB: took a parked arena of %zu MB%s\n", bytes >> 20, c->arena.size >> 20, fit->base ? ", cudaFree of ours" : "");
      if (fit->base) { cudaFree(fit->base); *fit = Arena(); }
      return 0;
    }
    // nothing fits: parked arenas are only in the way of the allocation below
    for (int k = 0; k < kArenaSlots; k++)
      if (slots[k].base) { cudaFree(slots[k].base); slots[k] = Arena(); }
  }
  if (c->arena.base) cudaFree(c->arena.base);
  c->arena = Arena();
  bytes = (bytes + ((size_t)256 << 20) - 1) & ~(((size_t)256 << 20) - 1);
  MR_ARENA_DBG("[mr] arena_reserve: cudaMalloc %zu MB\n", bytes >> 20);
  if (cudaMalloc(&c->arena.base, bytes) != cudaSuccess) {
    cudaGetLastError();
    MR_FAIL(c, "out of device memory: %zu MB of chunk work space", bytes >> 20);
  }
  c->arena.size = bytes;
  return 0;
}

// (re)allocate the per-chunk work space for chunks of up to `want` orientations
int ensure_chunk(mr_ctx* c, int64_t want) {
  const size_t words_c = c->Lc.words(c->nB);
  const int path = dp_path(c);
  const size_t words = path == 2 ? c->tiled_mp.tb_words : path == 1 ? c->tiled.tb_words : words_c;
  const size_t per_thread_scr = (size_t)c->N * sizeof(short2);
  const int per = std::max(1, c->hp.n_max - std::max(2, c->hp.n_min) + 1);
  // bytes per orientation in a chunk
  const size_t per_o = words * 4 + (size_t)c->NP * 4 * 2 + (size_t)per * per_thread_scr + 64;
  size_t budget = (size_t)c->opt_budget_mb * 1024 * 1024;
  // one wave of the DP kernel (one orientation per resident CTA) should fit a chunk: long
  // chains need hundreds of MB of traceback words per orientation, and a chunk smaller than
  // the wave leaves SMs idle.  Use up to 80 % of the free HBM for that.
  const int64_t wave = path == 2 ? c->tiled_mp.grid_cap : path == 1 ? c->tiled.grid_cap : c->dp_grid_cap;
  if (c->budget_is_default && budget < per_o * (size_t)wave && want > (int64_t)(budget / per_o)) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      // our arena and the parked ones are given back before a larger one is made
      const size_t held = c->arena.size + arena_cached_bytes(c);
      budget = std::max(budget, std::min(per_o * (size_t)wave + ((size_t)64 << 20), (size_t)(0.8 * (double)(free_b + held))));
    }
  }
  // (64 MB of the budget cover alignment and rounding of the sub-buffers, so that a chunk at
  // the cap always fits an arena of exactly `budget` bytes)
  int64_t cap = (int64_t)((budget - std::min<size_t>(budget / 8, (size_t)64 << 20)) / per_o);
  cap = std::max<int64_t>(1, std::min<int64_t>(cap, 1 << 18));
  if (cap > wave && want > cap) cap -= cap % wave;  // several chunks: whole waves
  if (c->opt_chunk > 0) cap = std::min<int64_t>(cap, c->opt_chunk);
  int64_t need = std::min<int64_t>(want, cap);
  if (need <= c->chunk_cap && words == c->tb_words) return 0;
  cudaStreamSynchronize(c->stream);
  harvest(c);
  c->d_tb = nullptr; c->d_cost = c->d_ssd = nullptr; c->d_surv = c->d_nw = c->d_nlast = nullptr;
  c->d_valid = nullptr; c->d_scratch = nullptr; c->d_rowcost = nullptr; c->d_rowcnt = nullptr;
  c->chunk_cap = 0;
  c->tb_words = words;
  c->scratch_threads = (((size_t)need * per + 127) / 128) * 128;  // one slice of N pairs per (orientation,size) thread
  c->nw_threads_cap = c->hp.use_filter ? (int)(((need + 127) / 128) * 128) : 0;
  // carve the arena (256-byte aligned sub-buffers)
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_tb = take(words * 4 * (size_t)need), o_cost = take((size_t)need * c->NP * 4),
               o_ssd = take((size_t)need * c->NP * 4), o_valid = take((size_t)need * c->NP),
               o_surv = take((size_t)need * 4), o_nw = take((size_t)need * 4), o_nlast = take((size_t)need * 4),
               o_scr = take(c->scratch_threads * per_thread_scr),
               o_rc = take((size_t)(c->nB + 1) * c->nw_threads_cap * 4),
               o_rn = take((size_t)(c->nB + 1) * c->nw_threads_cap * 2);
  // a chunk at the budget keeps the arena the same size from pair to pair (reusable as is)
  if (arena_reserve(c, std::max(off, need == cap && want > cap ? budget : (size_t)0))) return 1;
  char* base = c->arena.base;
  c->d_tb = reinterpret_cast<uint32_t*>(base + o_tb);
  c->d_cost = reinterpret_cast<float*>(base + o_cost);
  c->d_ssd = reinterpret_cast<float*>(base + o_ssd);
  c->d_valid = reinterpret_cast<unsigned char*>(base + o_valid);
  c->d_surv = reinterpret_cast<int*>(base + o_surv);
  c->d_nw = reinterpret_cast<int*>(base + o_nw);
  c->d_nlast = reinterpret_cast<int*>(base + o_nlast);
  c->d_scratch = reinterpret_cast<short2*>(base + o_scr);
  if (c->hp.use_filter) {
    c->d_rowcost = reinterpret_cast<float*>(base + o_rc);
    c->d_rowcnt = reinterpret_cast<unsigned short*>(base + o_rn);
  }
  c->chunk_cap = (int)need;
  return 0;
}

// launch the 3-D DP for the survivors of the current chunk
// latency: a few orientations whose result the host waits for (Refine batches): the multi-pass
// kernel then spreads the passes of an orientation over a cluster of CTAs (split mode)
// costs_only: no tie predicates and no traceback table (tiled kernels only; the general kernel
// always writes its table) -- FirstPass under -IR, see search_device
int launch_dp(mr_ctx* c, const float* d_orient_chunk, int count, bool latency = false, bool costs_only = false) {
  EventPair* ev = next_events(c, 0);
  if (!ev) MR_FAIL(c, "cudaEventCreate failed");
  CU(c, cudaEventRecord(ev->a, c->stream));
  const int path = dp_path(c);
  if (path == 1 && !costs_only && c->tiled.vp_bucket && c->opt_dp_form != 1) {
    // min-with-predicate form, from the patched cubin (mr_dp_vp.cu); there is no silent fallback:
    // a library that carries the cubin and cannot launch it reports the error
    std::string why;
    if (mr_launch_dp_tiles_vp(c->tiled, c->P, c->dA, c->dB, d_orient_chunk, c->surv, c->d_surv_count, c->d_tb,
                              c->d_cost, count, c->stream, c->device, &why) != 0)
      MR_FAIL(c, "tiled DP launch (min-with-predicate kernel) failed: %s", why.c_str());
  } else if (path == 1) {
    if (mr_launch_dp_tiled(c->tiled, c->P, c->dA, c->dB, d_orient_chunk, c->surv, c->d_surv_count,
                           c->d_tb, c->d_cost, count, c->stream, costs_only) != 0)
      MR_FAIL(c, "tiled DP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  } else if (path == 2) {
    if (mr_launch_dp_tiled_mp(c->tiled_mp, c->P, c->dA, c->dB, d_orient_chunk, c->surv, c->d_surv_count,
                              c->d_tb, c->d_cost, count, c->stream,
                              latency && c->opt_refine_split && count <= c->tiled_mp.split_cap, costs_only) != 0)
      MR_FAIL(c, "multi-pass tiled DP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  } else {
    if (c->use_gcol && !c->d_gcol) CU(c, cudaMalloc(&c->d_gcol, c->gcol_bytes));
    int grid = std::min(count, c->dp_grid_cap);
    k_dp_columns<<<grid, std::min(c->NP, 1024), c->dp_smem, c->stream>>>(c->P, c->Lc, c->dA, c->dB, d_orient_chunk,
                                                        c->surv, c->d_surv_count, c->d_tb,
                                                        c->d_cost, c->d_gcol, c->use_gcol);
  }
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(ev->b, c->stream));
  c->st.kernel_launches++;
  return 0;
}

template <bool EMIT, bool EXCL = false>
int launch_tb(mr_ctx* c, const float* d_orient_chunk, int count, const int2* items, int n_items,
              float* d_T = nullptr, short* d_pairs = nullptr, bool pairs_in_scratch = false) {
  if (!d_T) d_T = c->d_Tout;
  if (!d_pairs) d_pairs = c->d_pairs_out;
  if (pairs_in_scratch) d_pairs = nullptr;  // the pair lists stay in the per-thread scratch slices
  const int n_lo = std::max(2, c->hp.n_min);
  const int per = c->hp.n_max - n_lo + 1;
  const int64_t threads = items ? n_items : (int64_t)count * std::max(per, 0);
  if (threads <= 0) return 0;
  EventPair* ev = next_events(c, 2);
  if (!ev) MR_FAIL(c, "cudaEventCreate failed");
  CU(c, cudaEventRecord(ev->a, c->stream));
  const int grid = (int)((threads + 127) / 128);
  const size_t coords = 4 * 3 * (size_t)(c->m + c->nB) * c->hp.nb;
  const size_t base = coords + (use_tiled(c) ? 4 * (size_t)tiled_layout(c).table_ints() : 0);
  // emit mode: the pair lists also in shared memory (N pairs x 128 threads) when that fits
  size_t pairs_smem = 0;
  if (EMIT) {
    const size_t want = sizeof(short2) * 128 * (size_t)c->N;
    if (base + want <= (size_t)c->prop.sharedMemPerBlockOptin) pairs_smem = want;
  }
  if (use_tiled(c)) {
    auto kern = k_traceback_kabsch<TbTiled, EMIT, EXCL>;
    if (mr_ensure_smem(kern, base + pairs_smem)) MR_FAIL(c, "traceback kernel: shared memory attribute failed");
    kern<<<grid, 128, base + pairs_smem, c->stream>>>(
        c->P, tiled_layout(c), c->dA, c->dB, d_orient_chunk, c->surv, c->d_surv_count, c->d_tb,
        c->d_cost, items, n_items, n_lo, c->d_scratch, c->d_ssd, d_T, d_pairs, pairs_smem ? 1 : 0,
        EXCL ? c->d_excl_partner : nullptr, c->excl_ranks, c->excl_max_common, EXCL ? c->d_excl : nullptr);
  } else {
    auto kern = k_traceback_kabsch<TbColumns, EMIT, EXCL>;
    if (mr_ensure_smem(kern, base + pairs_smem)) MR_FAIL(c, "traceback kernel: shared memory attribute failed");
    kern<<<grid, 128, base + pairs_smem, c->stream>>>(
        c->P, c->Lc, c->dA, c->dB, d_orient_chunk, c->surv, c->d_surv_count, c->d_tb, c->d_cost,
        items, n_items, n_lo, c->d_scratch, c->d_ssd, d_T, d_pairs, pairs_smem ? 1 : 0,
        EXCL ? c->d_excl_partner : nullptr, c->excl_ranks, c->excl_max_common, EXCL ? c->d_excl : nullptr);
  }
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(ev->b, c->stream));
  c->st.kernel_launches++;
  return 0;
}

// filter (or identity) -> survivors of the chunk
// decision_only: the caller needs the survivors, not the match counts (FirstPass)
int launch_filter(mr_ctx* c, const float* d_orient_chunk, int count, bool apply_filter,
                  int* surv_out = nullptr, int* nw_out = nullptr, bool decision_only = false) {
  int* const surv = surv_out ? surv_out : c->d_surv;
  int* const nwm = nw_out ? nw_out : c->d_nw;
  c->surv = surv;
  if (apply_filter && c->hp.use_filter) {
    CU(c, cudaMemsetAsync(c->d_surv_count, 0, 4, c->stream));
    EventPair* ev = next_events(c, 1);
    if (!ev) MR_FAIL(c, "cudaEventCreate failed");
    CU(c, cudaEventRecord(ev->a, c->stream));
    const int cpl = (c->nB + 31) / 32;
    const bool by_cost = decision_only && c->nw_cost_ok && c->opt_nw_cost && c->hp.nb == 1 && !c->opt_nw_general;
    if (by_cost && cpl <= 16) {
      // up to 512 columns (k_nw_cost_warp, up to 16 columns per lane): 8 lanes per orientation up to 128
      // columns, 16 up to 256, a whole warp beyond
      const int sw = c->nB <= 128 ? 8 : c->nB <= 256 ? 16 : 32;
      const int cp = (c->nB + sw - 1) / sw;   // 1..16 (9..16 for 16 and 32 lanes)
      const int grid = (count + 4 * (32 / sw) - 1) / (4 * (32 / sw));
      const size_t sm = 4 * 3 * (size_t)c->m;
      typedef void (*nwk_t)(MrDevParams, float, const float*, const float*, const float*, int, int*, int*);
#define MR_NWC8(CPL) k_nw_cost_warp<CPL, 8>
#define MR_NWC16(CPL) k_nw_cost_warp<CPL, 16>
      static const nwk_t k8[16] = {MR_NWC8(1), MR_NWC8(2), MR_NWC8(3), MR_NWC8(4), MR_NWC8(5), MR_NWC8(6), MR_NWC8(7), MR_NWC8(8),
                                   MR_NWC8(9), MR_NWC8(10), MR_NWC8(11), MR_NWC8(12), MR_NWC8(13), MR_NWC8(14), MR_NWC8(15), MR_NWC8(16)};
      static const nwk_t k16[8] = {MR_NWC16(9), MR_NWC16(10), MR_NWC16(11), MR_NWC16(12), MR_NWC16(13), MR_NWC16(14), MR_NWC16(15), MR_NWC16(16)};
      static const nwk_t k32[8] = {k_nw_cost_warp<9, 32>, k_nw_cost_warp<10, 32>, k_nw_cost_warp<11, 32>, k_nw_cost_warp<12, 32>,
                                   k_nw_cost_warp<13, 32>, k_nw_cost_warp<14, 32>, k_nw_cost_warp<15, 32>, k_nw_cost_warp<16, 32>};
#undef MR_NWC8
#undef MR_NWC16
      const nwk_t k = sw == 8 ? k8[cp - 1] : sw == 16 ? k16[cp - 9] : k32[cp - 9];
      if (mr_ensure_smem(k, sm)) MR_FAIL(c, "NW filter: shared memory attribute failed");
      k<<<grid, 128, sm, c->stream>>>(c->P, c->nw_cost_thr, c->dA, c->dB, d_orient_chunk, count, surv, c->d_surv_count);
    } else if (by_cost && c->nB <= 8192) {
      // columns per lane: up to 16 (a warp spans 512 columns) beyond 512 columns -- the per-row work of
      // a lane (shuffle, band bounds, coordinates of the row) is spread over more cells: 600x600 801 ->
      // 1000 GCUPS, 1000x900 909 -> 1183, 2000x1500 1104 -> 1227 -- up to 8 below (two warps at least)
      const int wcols = c->nB > 512 ? 512 : 256;
      const int nw = (c->nB + wcols - 1) / wcols, cp = (c->nB + 32 * nw - 1) / (32 * nw);
      const size_t sm = 4 * (size_t)((3 * c->m + 3) & ~3) + 4 * 128 * (size_t)(nw - 1);
      const int grid = std::min(count, c->prop.multiProcessorCount * 16);
      typedef void (*nwk_t)(MrDevParams, float, const float*, const float*, const float*, int, int*, int*);
      static const nwk_t ks[2][12] = {{k_nw_cost_cta<5, 256>, k_nw_cost_cta<6, 256>, k_nw_cost_cta<7, 256>, k_nw_cost_cta<8, 256>,
                                       k_nw_cost_cta<9, 256>, k_nw_cost_cta<10, 256>, k_nw_cost_cta<11, 256>, k_nw_cost_cta<12, 256>,
                                       k_nw_cost_cta<13, 256>, k_nw_cost_cta<14, 256>, k_nw_cost_cta<15, 256>, k_nw_cost_cta<16, 256>},
                                      {k_nw_cost_cta<5, 1024>, k_nw_cost_cta<6, 1024>, k_nw_cost_cta<7, 1024>, k_nw_cost_cta<8, 1024>,
                                       k_nw_cost_cta<9, 1024>, k_nw_cost_cta<10, 1024>, k_nw_cost_cta<11, 1024>, k_nw_cost_cta<12, 1024>,
                                       k_nw_cost_cta<13, 1024>, k_nw_cost_cta<14, 1024>, k_nw_cost_cta<15, 1024>, k_nw_cost_cta<16, 1024>}};
      const nwk_t k = ks[nw > 8 ? 1 : 0][std::max(cp, 5) - 5];
      if (mr_ensure_smem(k, sm)) MR_FAIL(c, "NW filter: shared memory attribute failed");
      k<<<grid, 32 * nw, sm, c->stream>>>(c->P, c->nw_cost_thr, c->dA, c->dB, d_orient_chunk, count, surv, c->d_surv_count);
    } else if (c->hp.nb == 1 && cpl <= 8 && !c->opt_nw_general) {
      // warp per orientation, wavefront in registers
      const int grid = (count + 3) / 4;
      const size_t sm = 4 * 3 * (size_t)c->m;
#define MR_NW(CPL) if (mr_ensure_smem(k_nw_filter_warp<CPL>, sm)) MR_FAIL(c, "NW filter: shared memory attribute failed"); \
      k_nw_filter_warp<CPL><<<grid, 128, sm, c->stream>>>(c->P, c->dA, c->dB, d_orient_chunk, count, nwm, surv, c->d_surv_count)
      switch (cpl) {
        case 1: MR_NW(1); break; case 2: MR_NW(2); break; case 3: MR_NW(3); break; case 4: MR_NW(4); break;
        case 5: MR_NW(5); break; case 6: MR_NW(6); break; case 7: MR_NW(7); break; default: MR_NW(8); break;
      }
#undef MR_NW
    } else if (c->hp.nb == 1 && c->nB <= 8192 && !c->opt_nw_general) {
      // CTA per orientation: warps split the columns (256 per warp at most)
      const int nw = (c->nB + 255) / 256, cp = (c->nB + 32 * nw - 1) / (32 * nw);
      const size_t sm = 4 * (size_t)((3 * c->m + 3) & ~3) + 8 * 128 * (size_t)(nw - 1);
      const int grid = std::min(count, c->prop.multiProcessorCount * 16);
      typedef void (*nwk_t)(MrDevParams, const float*, const float*, const float*, int, int*, int*, int*);
      static const nwk_t ks[2][4] = {{k_nw_filter_cta<5, 256>, k_nw_filter_cta<6, 256>, k_nw_filter_cta<7, 256>, k_nw_filter_cta<8, 256>},
                                     {k_nw_filter_cta<5, 1024>, k_nw_filter_cta<6, 1024>, k_nw_filter_cta<7, 1024>, k_nw_filter_cta<8, 1024>}};
      const nwk_t k = ks[nw > 8 ? 1 : 0][std::max(cp, 5) - 5];
      if (mr_ensure_smem(k, sm)) MR_FAIL(c, "NW filter: shared memory attribute failed");
      k<<<grid, 32 * nw, sm, c->stream>>>(c->P, c->dA, c->dB, d_orient_chunk, count, nwm, surv, c->d_surv_count);
    } else {
      // thread per orientation; the row scratch holds nw_threads_cap orientations at a time
      for (int base = 0; base < count; base += c->nw_threads_cap) {
        const int sub = std::min(c->nw_threads_cap, count - base);
        k_nw_filter<<<(sub + 127) / 128, 128, 0, c->stream>>>(c->P, c->dA, c->dB, d_orient_chunk, sub, c->d_rowcost,
                                                             c->d_rowcnt, nwm, surv, c->d_surv_count, base);
      }
    }
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(ev->b, c->stream));
    c->st.nw_cells += nw_cells(c->m, c->nB, c->hp.of_min) * (int64_t)count;
  } else {
    k_iota<<<(count + 255) / 256, 256, 0, c->stream>>>(surv, c->d_surv_count, count);
    CU(c, cudaGetLastError());
  }
  c->st.kernel_launches++;
  return 0;
}

// flags of mr_search_dissimilar, one byte per (orientation of a chunk, size)
static int ensure_excl(mr_ctx* c) {
  if (!c->excl_on) return 0;
  const size_t need = (size_t)c->chunk_cap * c->NP;
  if (need <= c->excl_cap) return 0;
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_excl);
  c->d_excl = nullptr;
  c->excl_cap = 0;
  CU(c, cudaMalloc(&c->d_excl, need));
  c->excl_cap = need;
  return 0;
}

// FirstPass over a device-resident list
int search_device(mr_ctx* c, const float* d_orient, int64_t count, int64_t first_id) {
  if (count <= 0) return 0;
  // -IR (refine_while_searching = 0): the score of an orientation is cost(n) itself
  // (search_or_dyn3d.cc:576-589), so FirstPass needs no alignment: the DP runs without tie
  // predicates / traceback words and the select kernel reads the costs.  The <= R winners get
  // their paths in fetch_payload, which re-evaluates them with the full kernel.
  // (not with mr_search_dissimilar: the exclusion test needs every alignment)
  const bool costs_only = !c->hp.refine_while_searching && use_tiled(c) && !c->opt_costs_only_off && !c->excl_on;
  if (c->hp.use_filter) {
    // Phase 1: filter the whole list (it needs nothing per orientation but two ints), so that
    // phase 2 runs the DP over full chunks of survivors whatever the pass rate -- at config-4
    // size one orientation in 25 000 passes and a DP launch per list chunk would idle the GPU.
    if (count > 0x7fffffff) MR_FAIL(c, "orientation list too long");
    if (count > c->filter_cap) {
      CU(c, cudaStreamSynchronize(c->stream));
      cudaFree(c->d_surv_all); cudaFree(c->d_nw_all);
      c->d_surv_all = c->d_nw_all = nullptr;
      c->filter_cap = 0;
      CU(c, cudaMalloc(&c->d_surv_all, 4 * (size_t)count));
      CU(c, cudaMalloc(&c->d_nw_all, 4 * (size_t)count));
      c->filter_cap = count;
    }
    if (c->chunk_cap == 0 && ensure_chunk(c, std::min<int64_t>(count, 4096))) return 1;  // row scratch of the fallback filter
    if (launch_filter(c, d_orient, (int)count, true, c->d_surv_all, c->d_nw_all, true)) return 1;
    int* h_total = &c->h_pinned[c->h_pinned_cap - 1];
    CU(c, cudaMemcpyAsync(h_total, c->d_surv_count, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const int64_t total = *h_total;
    c->st.orientations += count;
    c->st.orientations_dp += total;
    c->st.dp_cells += dp_cells(c->m, c->nB, c->hp.n_min, c->hp.n_max) * total;
    c->st.d2h_bytes += 4;
    if (total > 0 && ensure_chunk(c, total)) return 1;
    if (total > 0 && ensure_excl(c)) return 1;
    for (int64_t s0 = 0; s0 < total; s0 += c->chunk_cap) {
      const int cnt = (int)std::min<int64_t>(c->chunk_cap, total - s0);
      k_set_int<<<1, 1, 0, c->stream>>>(c->d_surv_count, cnt);
      CU(c, cudaGetLastError());
      c->surv = c->d_surv_all + s0;  // indices into the whole list
      if (launch_dp(c, d_orient, cnt, false, costs_only)) return 1;
      if (c->opt_dp_only) continue;
      if (!costs_only && (c->excl_on ? launch_tb<false, true>(c, d_orient, cnt, nullptr, 0)
                                      : launch_tb<false>(c, d_orient, cnt, nullptr, 0))) return 1;
      k_select<<<(cnt + 127) / 128, 128, 0, c->stream>>>(c->P, c->NP, c->surv, c->d_surv_count, c->d_cost,
                                                        costs_only ? c->d_cost : c->d_ssd, (long long)first_id, c->d_keys,
                                                        nullptr, nullptr, c->scores_out, c->excl_on ? c->d_excl : nullptr);
      CU(c, cudaGetLastError());
      c->st.kernel_launches += 2;
    }
    c->surv = c->d_surv;
    return 0;
  }
  if (ensure_chunk(c, count)) return 1;
  if (ensure_excl(c)) return 1;
  for (int64_t start = 0; start < count; start += c->chunk_cap) {
    const int cnt = (int)std::min<int64_t>(c->chunk_cap, count - start);
    const float* oc = d_orient + 12 * (size_t)start;
    if (launch_filter(c, oc, cnt, false)) return 1;
    // survivor count -> pinned host (for the stats only; nothing waits on it)
    if ((int)c->pending_counts.size() >= c->h_pinned_cap - 1) {
      CU(c, cudaStreamSynchronize(c->stream));
      harvest(c);
    }
    const int slot = (int)c->pending_counts.size();
    CU(c, cudaMemcpyAsync(&c->h_pinned[slot], c->d_surv_count, 4, cudaMemcpyDeviceToHost, c->stream));
    c->pending_counts.push_back({slot, cnt});
    if (launch_dp(c, oc, cnt, false, costs_only)) return 1;
    if (c->opt_dp_only) { c->st.orientations += cnt; continue; }
    if (!costs_only && (c->excl_on ? launch_tb<false, true>(c, oc, cnt, nullptr, 0)
                                    : launch_tb<false>(c, oc, cnt, nullptr, 0))) return 1;
    k_select<<<(cnt + 127) / 128, 128, 0, c->stream>>>(c->P, c->NP, c->surv, c->d_surv_count,
                                                      c->d_cost, costs_only ? c->d_cost : c->d_ssd,
                                                      (long long)(first_id + start), c->d_keys, nullptr, nullptr,
                                                      c->scores_out ? c->scores_out + (size_t)start * c->R : nullptr,
                                                      c->excl_on ? c->d_excl : nullptr);
    CU(c, cudaGetLastError());
    c->st.kernel_launches++;
    c->st.orientations += cnt;
  }
  return 0;
}

int reset_keys(mr_ctx* c) {
  std::vector<unsigned long long> k((size_t)std::max(c->R, 1), KEY_NONE);
  CU(c, cudaMemcpyAsync(c->d_keys, k.data(), sizeof(unsigned long long) * k.size(),
                        cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->id_offset_set = false;
  return 0;
}

void best_init_host(const mr_ctx* c, mr_best* b) {  // search_or_dyn3d.cc:165-184
  for (int r = 0; r < c->R; r++) {
    b->ssd[r] = MR_UPPER_BOUND;
    for (int k = 0; k < 12; k++) b->T[12 * r + k] = (k % 5 == 0) ? 1.0f : 0.0f;
    if (b->o_id) b->o_id[r] = -1;
  }
}

// rows of the device-side payload table: {key lo, key hi, 12 transform floats, N x (i | j << 16)}
static inline int row_ints(const mr_ctx* c) { return 2 + 12 + c->N; }

__global__ void k_rows_init(int32_t* __restrict__ rows, int R, int C) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * C) return;
  const int k = t % C;
  rows[t] = k == 0 ? (int32_t)0xffffffffu : k == 1 ? (int32_t)0x6aa56fa6 : 0;  // KEY_NONE, then zeros
}
// one block per emitted item: row[item_row[k]] = {key, T_out[k], pairs_out[k]}
__global__ void k_rows_pack(const int* __restrict__ item_row, const unsigned long long* __restrict__ keys,
                            const float* __restrict__ T_out, const short* __restrict__ pairs_out, int N, int C,
                            int32_t* __restrict__ rows) {
  const int k = blockIdx.x, r = item_row[k];
  int32_t* row = rows + (size_t)r * C;
  const int32_t* pr = reinterpret_cast<const int32_t*>(pairs_out + (size_t)k * 2 * N);
  for (int q = threadIdx.x; q < C; q += blockDim.x) {
    int32_t v;
    if (q == 0) v = (int32_t)(keys[r] & 0xffffffffull);
    else if (q == 1) v = (int32_t)(keys[r] >> 32);
    else if (q < 14) v = __float_as_int(T_out[12 * (size_t)k + (q - 2)]);
    else v = pr[q - 14];
    row[q] = v;
  }
}

// Turn the device key table into a full best table: re-run the DP for the (few) winning
// orientations and emit transform + alignment for exactly the winning (orientation,size)
// pairs.  d_orient / id_offset locate a winner's transform: index = o_id - id_offset.
//   out != NULL: the host table is filled (mr_fetch_best, mr_search).
//   out == NULL: multi-GPU reduction: d_keys holds the cross-rank minima, d_keys_local this
//     context's own keys; only the sizes this context owns (equal keys) are emitted, as rows
//     of c->d_rows in device memory (nothing but the two key vectors is copied to the host).
int fetch_payload(mr_ctx* c, const float* d_orient, long long id_offset, mr_best* out) {
  const bool to_rows = out == nullptr;
  std::vector<unsigned long long> keys((size_t)c->R), local;
  CU(c, cudaMemcpyAsync(keys.data(), c->d_keys, 8 * (size_t)c->R, cudaMemcpyDeviceToHost, c->stream));
  if (to_rows) {
    local.resize((size_t)c->R);
    CU(c, cudaMemcpyAsync(local.data(), c->d_keys_local, 8 * (size_t)c->R, cudaMemcpyDeviceToHost, c->stream));
    const int C = row_ints(c), total = c->R * C;
    k_rows_init<<<(total + 255) / 256, 256, 0, c->stream>>>(c->d_rows, c->R, C);
    CU(c, cudaGetLastError());
    c->st.kernel_launches++;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  c->st.d2h_bytes += 8 * (int64_t)c->R * (to_rows ? 2 : 1);
  if (out) best_init_host(c, out);
  std::map<long long, int> widx;  // o_id -> winner slot
  std::vector<int> idx;
  std::vector<int2> items;
  std::vector<int> item_r;
  for (int r = 0; r < c->R; r++) {
    if (keys[r] == KEY_NONE) continue;
    if (to_rows && local[r] != keys[r]) continue;  // another rank's orientation won this size
    const long long oid = (long long)(keys[r] & 0xffffffffull);
    auto it = widx.find(oid);
    int w;
    if (it == widx.end()) {
      w = (int)idx.size();
      widx[oid] = w;
      idx.push_back((int)(oid - id_offset));
    } else {
      w = it->second;
    }
    items.push_back(make_int2(w, c->hp.n_min + r));
    item_r.push_back(r);
  }
  if (items.empty()) return 0;
  if (ensure_chunk(c, (int64_t)idx.size())) return 1;
  // the winners are re-evaluated in groups that fit the chunk work space
  // (large problems hold only a few traceback tables at a time)
  for (size_t g0 = 0; g0 < idx.size(); g0 += (size_t)c->chunk_cap) {
    const int nw = (int)std::min<size_t>((size_t)c->chunk_cap, idx.size() - g0);
    std::vector<int2> sub;
    std::vector<int> sub_r;
    for (size_t k = 0; k < items.size(); k++)
      if (items[k].x >= (int)g0 && items[k].x < (int)g0 + nw) {
        sub.push_back(make_int2(items[k].x - (int)g0, items[k].y));
        sub_r.push_back(item_r[k]);
      }
    const int ni = (int)sub.size();
    CU(c, cudaMemcpyAsync(c->d_widx, idx.data() + g0, 4 * (size_t)nw, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_items, sub.data(), sizeof(int2) * (size_t)ni, cudaMemcpyHostToDevice, c->stream));
    if (to_rows) CU(c, cudaMemcpyAsync(c->d_item_row, sub_r.data(), 4 * (size_t)ni, cudaMemcpyHostToDevice, c->stream));
    k_gather_orient<<<(nw * 12 + 127) / 128, 128, 0, c->stream>>>(d_orient, c->d_widx, nw, c->d_orient_w);
    CU(c, cudaGetLastError());
    c->st.kernel_launches++;
    if (launch_filter(c, c->d_orient_w, nw, false)) return 1;
    if (launch_dp(c, c->d_orient_w, nw)) return 1;
    if (launch_tb<true>(c, c->d_orient_w, nw, c->d_items, ni)) return 1;
    if (to_rows) {
      k_rows_pack<<<ni, 128, 0, c->stream>>>(c->d_item_row, c->d_keys, c->d_Tout, c->d_pairs_out, c->N, row_ints(c), c->d_rows);
      CU(c, cudaGetLastError());
      c->st.kernel_launches++;
      CU(c, cudaStreamSynchronize(c->stream));  // the host vectors above go out of scope
      continue;
    }
    std::vector<float> T((size_t)ni * 12);
    std::vector<short> pairs((size_t)ni * 2 * c->N);
    CU(c, cudaMemcpyAsync(T.data(), c->d_Tout, 4 * T.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(pairs.data(), c->d_pairs_out, 2 * pairs.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->st.d2h_bytes += (int64_t)(4 * T.size() + 2 * pairs.size());
    for (int k = 0; k < ni; k++) {
      const int r = sub_r[k], n = c->hp.n_min + r;
      unsigned int bits = (unsigned int)(keys[r] >> 32);
      memcpy(&out->ssd[r], &bits, 4);
      memcpy(out->T + 12 * r, &T[12 * (size_t)k], 48);
      memcpy(out->pairs + (size_t)r * 2 * c->N, &pairs[(size_t)k * 2 * c->N], sizeof(short) * 2 * (size_t)n);
      if (out->o_id) out->o_id[r] = (long long)(keys[r] & 0xffffffffull);
    }
  }
  return 0;
}

}  // namespace

// --------------------------------------------------------------------------- C ABI
extern "C" {

const char* mr_last_error(const mr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t mr_dp_cells_per_orientation(int32_t m, int32_t nB, int32_t n_min, int32_t n_max) {
  return dp_cells(m, nB, n_min, n_max);
}

int mr_create(mr_ctx** out, const float* A, int32_t m, const float* B, int32_t nB,
              const mr_params* p, int32_t device) {
  *out = nullptr;
  auto fail = [&](const std::string& s) { g_create_error = s; return 1; };
  if (!A || !B || !p) return fail("mr_create: null argument");
  const int N = std::min(m, nB);
  if (m < 1 || nB < 1 || p->nb < 1) return fail("mr_create: empty molecule");
  if (p->n_min < 1 || p->n_min > N || p->n_max < 1 || p->n_max > N || p->n_max < p->n_min)
    return fail("Minimum and/or maximum number of matches allowed does not make sense "
                "(lib/dyn/dyn3d.cc:459-467)");
  if ((long long)N * p->nb > 32767) return fail("n*nb exceeds the reference's short limit (32767)");
  if (p->use_filter && (p->of_min < 1 || p->of_min > N)) return fail("mr_create: of_min out of range");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(std::string("mr_create: no CUDA device (this library has no CPU fallback): ") +
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("mr_create: bad device index");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(cudaGetErrorString(e));
  mr_ctx* c = new mr_ctx();
  c->device = device;
  const bool dbg = getenv("MR_DEBUG_PLAN") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    if (!dbg) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mr] create: %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  {  // cudaGetDeviceProperties costs 2-11 ms: once per device and process
    static cudaDeviceProp props[64];
    static bool have[64] = {false};
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (!have[device & 63]) { cudaGetDeviceProperties(&props[device & 63], device); have[device & 63] = true; }
    c->prop = props[device & 63];
  }
  tick("device properties");
  c->hp = *p;
  // MINRMS_B200_DP_FORM=1: the two-min / ping-pong tile updates for every context of the process
  // (field debugging and A/B runs without touching the caller; mr_set_option("dp_form") overrides)
  if (const char* e = getenv("MINRMS_B200_DP_FORM")) c->opt_dp_form = atoi(e) == 1 ? 1 : 0;
  c->m = m; c->nB = nB; c->N = N; c->R = p->n_max - p->n_min + 1;
  c->NP = ((p->n_max + 31) / 32) * 32;
  memset(&c->st, 0, sizeof c->st);
  MrDevParams& P = c->P;
  P.m = m; P.nB = nB; P.nb = p->nb; P.N = N;
  P.n_min = p->n_min; P.n_max = p->n_max;
  P.wA = m - p->n_min + 1; P.wB = nB - p->n_min + 1;
  P.dp_cutoff = metric_cutoff(p->max_dist, p->nb);
  P.max_rmsd = p->max_rmsd;
  P.use_filter = p->use_filter;
  P.nw_cutoff = metric_cutoff(p->of_dist, p->nb);
  // orientation_filter.cc:77-79 (double expression rounded to float)
  P.nw_gap = (float)(0.5 * (double)(p->of_dist * p->of_dist) * (double)N * (double)p->nb);
  P.of_min = p->of_min;
  P.w0 = m - p->of_min + 1; P.w1 = nB - p->of_min + 1;
  P.refine_while_searching = p->refine_while_searching;
  if (p->use_filter) {
    // decision-only NW filter: cost intervals of "k >= of_min" and "k <= of_min - 1" matches
    // (derivation at k_nw_cost_warp); E = half an ulp of the running value per rounding on a path
    const double G = (double)P.nw_gap, cut = (double)P.nw_cutoff;
    double E = 0.0;
    auto half_ulp = [](double x) { int e; frexp(x, &e); return ldexp(1.0, e - 1 - 23) * 0.5; };  // float ulp / 2
    for (int sgap = 1; sgap <= m + nB; sgap++) E += half_ulp(1.001 * sgap * G);
    E += half_ulp(1.001 * (m + nB) * G);
    const double U = (double)(m + nB - 2 * p->of_min) * G + (double)p->of_min * cut + E;   // most a cost with k >= of_min can be
    const double L = (double)(m + nB - 2 * (p->of_min - 1)) * G - E;                       // least a cost with k <= of_min - 1 can be
    const float thr = (float)(0.5 * (U + L));
    c->nw_cost_margin = L - U;
    c->nw_cost_ok = U < (double)thr && (double)thr < L;
    c->nw_cost_thr = thr;
  }
  c->Lc.IW = (m + 15) / 16;
  c->Lc.NP = c->NP;
  auto bail = [&](const char* what) {
    g_create_error = std::string("mr_create: ") + what + ": " + cudaGetErrorString(cudaGetLastError());
    mr_destroy(c);
    return 1;
  };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
  const size_t na = (size_t)m * p->nb * 3, nbb = (size_t)nB * p->nb * 3;
  if (cudaMalloc(&c->dA, na * 4) != cudaSuccess || cudaMalloc(&c->dB, nbb * 4) != cudaSuccess) return bail("malloc");
  if (cudaMemcpy(c->dA, A, na * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail("memcpy");
  if (cudaMemcpy(c->dB, B, nbb * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail("memcpy");
  tick("stream + coordinates");
  // general DP kernel: column buffers in shared memory when they fit
  const size_t colbytes = 2 * (size_t)(m + 1) * c->NP * 4;
  const size_t coordbytes = (na + nbb) * 4;
  const size_t smem_max = c->prop.sharedMemPerBlockOptin;
  if (coordbytes > smem_max) { g_create_error = "mr_create: molecules too large for shared memory"; mr_destroy(c); return 1; }
  if (coordbytes + colbytes <= smem_max) {
    c->use_gcol = 0;
    c->dp_smem = coordbytes + colbytes;
    c->dp_grid_cap = c->prop.multiProcessorCount * 4;
  } else {
    c->use_gcol = 1;
    c->dp_smem = coordbytes;
    c->dp_grid_cap = c->prop.multiProcessorCount * 2;
    c->gcol_bytes = colbytes * c->dp_grid_cap;  // allocated when the general kernel first runs
  }
  if (mr_ensure_smem(k_dp_columns, c->dp_smem)) return bail("smem attribute");
  tick("general kernel setup");
  mr_plan_dp_tiled(&c->tiled, P, c->prop, c->NP);
  if (!c->tiled.ok || c->tiled.threads > 512) plan_mp(c);
  tick("tiled plan");
  const int RR = std::max(c->R, 1), slots = std::max(RR, c->NP);
  if (cudaMalloc(&c->d_keys, 8 * (size_t)RR) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_surv_count, 4) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_orient_w, 48 * (size_t)RR) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_widx, 4 * (size_t)RR) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_items, sizeof(int2) * (size_t)RR) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_Tout, 48 * (size_t)slots) != cudaSuccess) return bail("malloc");
  if (cudaMalloc(&c->d_pairs_out, sizeof(short) * 2 * (size_t)N * slots) != cudaSuccess) return bail("malloc");
  c->h_pinned_cap = 4096;
  if (cudaMallocHost(&c->h_pinned, 4 * (size_t)c->h_pinned_cap) != cudaSuccess) return bail("pinned");
  if (reset_keys(c)) { g_create_error = c->err; mr_destroy(c); return 1; }
  tick("small buffers");
  *out = c;
  return 0;
}

void mr_destroy(mr_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& e : c->user_ev) if (e) cudaEventDestroy(e);
  arena_release(c);
  cudaFree(c->dA); cudaFree(c->dB); cudaFree(c->d_orient); cudaFree(c->d_surv_count);
  cudaFree(c->d_surv_all); cudaFree(c->d_nw_all);
  cudaFree(c->d_gcol); cudaFree(c->d_keys); cudaFree(c->d_orient_w); cudaFree(c->d_scores);
  cudaFree(c->d_excl_partner); cudaFree(c->d_excl);
  cudaFree(c->d_keys_local); cudaFree(c->d_rows); cudaFree(c->d_item_row); cudaFree(c->d_misc);
  cudaFree(c->d_widx); cudaFree(c->d_items); cudaFree(c->d_Tout); cudaFree(c->d_pairs_out);
  cudaFree(c->d_ref_pairs_in); cudaFree(c->d_ref_n); cudaFree(c->d_ref_orient); cudaFree(c->d_ref_T);
  if (c->h_ref) cudaFreeHost(c->h_ref);
  mr_free_dp_tiled(&c->tiled);
  mr_free_dp_tiled_mp(&c->tiled_mp);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int mr_best_init(const mr_ctx* c, mr_best* best) {
  if (!c || !best || !best->ssd || !best->T || !best->pairs) return 1;
  best_init_host(c, best);
  return 0;
}

// read-only view of what a context will run (tests, logs): see include/minrms_b200.h
int mr_get_info(mr_ctx* c, const char* name, int64_t* value) {
  if (!c || !name || !value) return 1;
  const int path = dp_path(c);
  if (!strcmp(name, "dp_path")) *value = path;   // 0 general, 1 tiled single-pass, 2 tiled multi-pass
  else if (!strcmp(name, "dp_threads")) *value = path == 1 ? c->tiled.threads : path == 2 ? c->tiled_mp.threads : std::min(c->NP, 1024);
  else if (!strcmp(name, "dp_vp_bucket")) *value = (path == 1 && c->opt_dp_form != 1) ? c->tiled.vp_bucket : 0;
  else if (!strcmp(name, "dp_ctas_per_sm"))
    *value = path == 1 ? ((c->tiled.vp_bucket && c->opt_dp_form != 1) ? c->tiled.grid_cap_vp : c->tiled.grid_cap) / c->prop.multiProcessorCount
                       : path == 2 ? c->tiled_mp.grid_cap / c->prop.multiProcessorCount : c->dp_grid_cap / c->prop.multiProcessorCount;
  else if (!strcmp(name, "dp_passes")) *value = path == 2 ? c->tiled_mp.npasses : 1;
  else if (!strcmp(name, "nw_decided_by_cost")) *value = (c->hp.use_filter && c->nw_cost_ok && c->opt_nw_cost && c->hp.nb == 1 && !c->opt_nw_general) ? 1 : 0;
  else MR_FAIL(c, "mr_get_info: unknown name %s", name);
  return 0;
}

int mr_set_option(mr_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return 1;
  if (!strcmp(name, "dp_kernel")) c->opt_dp_kernel = value;
  else if (!strcmp(name, "budget_mb")) { c->opt_budget_mb = std::max<int64_t>(64, value); c->budget_is_default = false; }
  else if (!strcmp(name, "chunk")) c->opt_chunk = value;
  else if (!strcmp(name, "dp_form")) {
    if (value < 0 || value > 1) MR_FAIL(c, "mr_set_option: dp_form is 0 or 1");
    c->opt_dp_form = value;
  }
  else if (!strcmp(name, "exp_dp_only")) c->opt_dp_only = value;
  else if (!strcmp(name, "ir_costs_only")) c->opt_costs_only_off = value == 0;
  else if (!strcmp(name, "nw_kernel")) c->opt_nw_general = (value == 1);
  else if (!strcmp(name, "nw_cost_filter")) c->opt_nw_cost = value != 0;
  else if (!strcmp(name, "refine_batch")) c->opt_refine_batch = std::max<int64_t>(0, value);
  else if (!strcmp(name, "refine_split")) c->opt_refine_split = value != 0;
  else if (!strcmp(name, "refine_seq_superpose")) c->opt_refine_seq_superpose = value != 0;
  else if (!strcmp(name, "mp_threads")) {
    if (value != 0 && (value < 32 || value > 1024 || value % 32)) MR_FAIL(c, "mr_set_option: mp_threads must be a multiple of 32 in [32, 1024]");
    c->opt_mp_threads = value;
    if (cudaSetDevice(c->device) != cudaSuccess) MR_FAIL(c, "cudaSetDevice failed");
    cudaStreamSynchronize(c->stream);
    if (c->tiled_mp.ok || c->opt_dp_kernel == 3) plan_mp(c, (int)c->opt_mp_threads);
  }
  else MR_FAIL(c, "mr_set_option: unknown option %s", name);
  if (c->opt_dp_kernel == 2 && !c->tiled.ok) MR_FAIL(c, "tiled DP kernel not available for this problem: %s", c->tiled.why);
  if (c->opt_dp_kernel == 3) {
    if (cudaSetDevice(c->device) != cudaSuccess) MR_FAIL(c, "cudaSetDevice failed");
    if (!c->tiled_mp.ok) plan_mp(c, (int)c->opt_mp_threads);
    if (!c->tiled_mp.ok) MR_FAIL(c, "multi-pass tiled DP kernel not available for this problem: %s", c->tiled_mp.why);
  }
  return 0;
}

int mr_upload_orientations(mr_ctx* c, const float* orient12, int64_t count) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  if (count < 0 || (count > 0 && !orient12)) MR_FAIL(c, "mr_upload_orientations: bad arguments");
  if (ensure_orient_cap(c, std::max<int64_t>(count, 1))) return 1;
  CU(c, cudaMemcpyAsync(c->d_orient, orient12, 48 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->orient_count = count;
  c->st.h2d_bytes += 48 * count;
  return 0;
}

int mr_reset_best(mr_ctx* c) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  return reset_keys(c);
}

int mr_search_resident(mr_ctx* c, int64_t first, int64_t count, int64_t first_id) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  if (first < 0 || count < 0 || first + count > c->orient_count) MR_FAIL(c, "mr_search_resident: range outside the uploaded list");
  if (first_id < 0 || first_id + count > 0xfffffffell) MR_FAIL(c, "orientation ids must fit 32 bits");
  const long long off = first_id - first;
  if (c->id_offset_set && off != c->id_offset) MR_FAIL(c, "mr_search_resident: id mapping changed since mr_reset_best");
  c->id_offset = off;
  c->id_offset_set = true;
  if (search_device(c, c->d_orient + 12 * (size_t)first, count, first_id)) return 1;
  CU(c, cudaStreamSynchronize(c->stream));
  harvest(c);
  return 0;
}

int mr_fetch_keys(mr_ctx* c, uint64_t* keys) {
  if (!c || !keys) return 1;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(keys, c->d_keys, 8 * (size_t)c->R, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->st.d2h_bytes += 8 * (int64_t)c->R;
  return 0;
}

int mr_fetch_best(mr_ctx* c, mr_best* best) {
  if (!c || !best) return 1;
  CU(c, cudaSetDevice(c->device));
  if (fetch_payload(c, c->d_orient, c->id_offset, best)) return 1;
  harvest(c);
  return 0;
}

int mr_device_keys(mr_ctx* c, uint64_t** d_keys, int32_t* count) {
  if (!c || !d_keys) return 1;
  CU(c, cudaSetDevice(c->device));
  const size_t RR = (size_t)std::max(c->R, 1);
  if (!c->d_keys_local) {
    CU(c, cudaMalloc(&c->d_keys_local, 8 * RR));
    CU(c, cudaMalloc(&c->d_rows, 4 * RR * (size_t)row_ints(c)));
    CU(c, cudaMalloc(&c->d_item_row, 4 * RR));
  }
  // snapshot of this context's own keys: after the caller's MIN all-reduce over d_keys, a size
  // whose reduced key still equals the snapshot is owned here
  CU(c, cudaMemcpyAsync(c->d_keys_local, c->d_keys, 8 * (size_t)c->R, cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *d_keys = reinterpret_cast<uint64_t*>(c->d_keys);
  if (count) *count = c->R;
  return 0;
}

int mr_fetch_rows_device(mr_ctx* c, int32_t** d_rows, int32_t* ints_per_row) {
  if (!c || !d_rows) return 1;
  CU(c, cudaSetDevice(c->device));
  if (!c->d_keys_local) MR_FAIL(c, "mr_fetch_rows_device: call mr_device_keys first");
  if (fetch_payload(c, c->d_orient, c->id_offset, nullptr)) return 1;
  CU(c, cudaStreamSynchronize(c->stream));
  harvest(c);
  *d_rows = c->d_rows;
  if (ints_per_row) *ints_per_row = row_ints(c);
  return 0;
}

int mr_search(mr_ctx* c, const float* orient12, int64_t count, int64_t first_id, mr_best* best) {
  if (!c || !best) return 1;
  CU(c, cudaSetDevice(c->device));
  if (count < 0 || (count > 0 && !orient12)) MR_FAIL(c, "mr_search: bad arguments");
  if (first_id < 0 || first_id + count > 0xfffffffell) MR_FAIL(c, "orientation ids must fit 32 bits");
  if (count == 0) return 0;
  if (mr_upload_orientations(c, orient12, count)) return 1;
  if (reset_keys(c)) return 1;
  c->id_offset = first_id;
  c->id_offset_set = true;
  if (search_device(c, c->d_orient, count, first_id)) return 1;
  std::vector<float> ssd((size_t)c->R), T((size_t)c->R * 12);
  std::vector<int16_t> pairs((size_t)c->R * 2 * c->N);
  std::vector<int64_t> oid((size_t)c->R);
  mr_best mine = {ssd.data(), T.data(), pairs.data(), oid.data()};
  if (fetch_payload(c, c->d_orient, first_id, &mine)) return 1;
  harvest(c);
  // merge into the caller's table: strictly lower score wins; on equal score the lower
  // orientation id wins (what sequential evaluation in id order gives, :1825-1835)
  for (int r = 0; r < c->R; r++) {
    if (!(mine.ssd[r] < MR_UPPER_BOUND)) continue;
    bool take = mine.ssd[r] < best->ssd[r];
    if (!take && best->o_id && mine.ssd[r] == best->ssd[r] && best->o_id[r] >= 0 && mine.o_id[r] < best->o_id[r])
      take = true;
    if (!take) continue;
    const int n = c->hp.n_min + r;
    best->ssd[r] = mine.ssd[r];
    memcpy(best->T + 12 * r, mine.T + 12 * r, 48);
    memcpy(best->pairs + (size_t)r * 2 * c->N, mine.pairs + (size_t)r * 2 * c->N, sizeof(int16_t) * 2 * (size_t)n);
    if (best->o_id) best->o_id[r] = mine.o_id[r];
  }
  return 0;
}

// evaluate the transform sitting in d_orient_w[0]; copy every per-size result back
static int eval_w0(mr_ctx* c, bool apply_filter, int32_t* n_last, int32_t* nw_matches, float* cost,
                   uint8_t* valid, float* ssd, float* T, int16_t* pairs) {
  const int N = c->N, NP = c->NP;
  if (ensure_chunk(c, 1)) return 1;
  if (launch_filter(c, c->d_orient_w, 1, apply_filter)) return 1;
  int h[2] = {0, -1};
  CU(c, cudaMemcpyAsync(&c->h_pinned[c->h_pinned_cap - 1], c->d_surv_count, 4, cudaMemcpyDeviceToHost, c->stream));
  if (apply_filter && c->hp.use_filter) {
    CU(c, cudaMemcpyAsync(&h[1], c->d_nw, 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(c, cudaMemsetAsync(c->d_valid, 0, (size_t)NP, c->stream));
  CU(c, cudaMemsetAsync(c->d_cost, 0, (size_t)NP * 4, c->stream));
  CU(c, cudaMemsetAsync(c->d_ssd, 0, (size_t)NP * 4, c->stream));
  CU(c, cudaMemsetAsync(c->d_nlast, 0, 4, c->stream));
  if (launch_dp(c, c->d_orient_w, 1)) return 1;
  if (launch_tb<true>(c, c->d_orient_w, 1, nullptr, 0)) return 1;
  k_select<<<1, 32, 0, c->stream>>>(c->P, NP, c->d_surv, c->d_surv_count, c->d_cost, c->d_ssd, 0ll,
                                   nullptr, c->d_valid, c->d_nlast);
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  std::vector<float> hc((size_t)NP), hs((size_t)NP), hT((size_t)NP * 12);
  std::vector<uint8_t> hv((size_t)NP);
  std::vector<int16_t> hp((size_t)NP * 2 * N);
  CU(c, cudaMemcpyAsync(&h[0], c->d_nlast, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hc.data(), c->d_cost, 4 * (size_t)NP, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hs.data(), c->d_ssd, 4 * (size_t)NP, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hv.data(), c->d_valid, (size_t)NP, cudaMemcpyDeviceToHost, c->stream));
  if (T) CU(c, cudaMemcpyAsync(hT.data(), c->d_Tout, 48 * (size_t)NP, cudaMemcpyDeviceToHost, c->stream));
  if (pairs) CU(c, cudaMemcpyAsync(hp.data(), c->d_pairs_out, 2 * hp.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->st.d2h_bytes += (int64_t)NP * (9 + (T ? 48 : 0) + (pairs ? 4 * N : 0));
  const bool passed = c->h_pinned[c->h_pinned_cap - 1] > 0;
  c->events_used = 0;  // single evaluations are not part of the throughput stats
  if (n_last) *n_last = passed ? h[0] : 0;
  if (nw_matches) *nw_matches = h[1];
  const int nl = passed ? h[0] : 0;
  for (int n = 1; n <= N; n++) {
    const bool v = passed && n <= c->hp.n_max && hv[n - 1];
    if (cost) cost[n - 1] = (n <= nl) ? hc[n - 1] : 0.0f;
    if (valid) valid[n - 1] = v;
    if (ssd) ssd[n - 1] = v ? hs[n - 1] : 0.0f;
    if (T) for (int k = 0; k < 12; k++) T[12 * (n - 1) + k] = v ? hT[12 * (size_t)(n - 1) + k] : 0.0f;
    if (pairs)
      for (int k = 0; k < 2 * N; k++)
        pairs[(size_t)(n - 1) * 2 * N + k] = (v && k < 2 * n) ? hp[(size_t)(n - 1) * 2 * N + k] : 0;
  }
  return 0;
}

// FirstPass that also returns what KeepTrackOfAltSolutions is handed for every (orientation,
// size) it is called on (search_or_dyn3d.cc:1795-1811, 1897-1904): the compared score, or
// MR_UPPER_BOUND where the reference never gets there (filtered orientation, early exit).
int mr_search_scores(mr_ctx* c, const float* orient12, int64_t count, int64_t first_id, mr_best* best,
                     float* scores) {
  if (!c || !best || !scores) return 1;
  CU(c, cudaSetDevice(c->device));
  if (count <= 0) return mr_search(c, orient12, count, first_id, best);
  if (c->opt_dp_only) MR_FAIL(c, "mr_search_scores: dp_only is set");
  const int64_t need = count * (int64_t)c->R;
  if (need > ((int64_t)1 << 34)) MR_FAIL(c, "mr_search_scores: orientations x sizes exceeds 2^34 scores; search the list in slices");
  if (need > c->scores_cap) {
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_scores);
    c->d_scores = nullptr;
    c->scores_cap = 0;
    CU(c, cudaMalloc(&c->d_scores, 4 * (size_t)need));
    c->scores_cap = need;
  }
  k_fill_f32<<<(int)std::min<int64_t>((need + 255) / 256, 148 * 16), 256, 0, c->stream>>>(c->d_scores, (size_t)need, MR_UPPER_BOUND);
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  c->scores_out = c->d_scores;
  const int rc = mr_search(c, orient12, count, first_id, best);
  c->scores_out = nullptr;
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(scores, c->d_scores, 4 * (size_t)need, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->st.d2h_bytes += 4 * need;
  return 0;
}

// One pass of SearchOrDyn3d::MultiPass (search_or_dyn3d.cc:1975-2250): FirstPass in which an
// alignment may be recorded only if it shares at most max_pairs_common pairs with the stored
// alignment of the same size of EVERY table in lower[] (:2086-2107).  The early exits of the
// n-loop do not depend on that test, so the only change on the device is a flag per
// (orientation, size) from the traceback kernel that the select kernel honours.
int mr_search_dissimilar(mr_ctx* c, const float* orient12, int64_t count, int64_t first_id, mr_best* best,
                         int32_t n_lower, const mr_best* lower, int32_t max_pairs_common) {
  if (!c || !best) return 1;
  CU(c, cudaSetDevice(c->device));
  if (n_lower < 0 || (n_lower > 0 && !lower) || max_pairs_common < 0) MR_FAIL(c, "mr_search_dissimilar: bad arguments");
  if (n_lower == 0 || count <= 0) return mr_search(c, orient12, count, first_id, best);
  if (c->opt_dp_only) MR_FAIL(c, "mr_search_dissimilar: dp_only is set");
  const size_t need = (size_t)n_lower * c->R * c->m;
  std::vector<short> partner(need, (short)-1);
  for (int q = 0; q < n_lower; q++)
    for (int r = 0; r < c->R; r++) {
      if (!lower[q].ssd || !lower[q].pairs) MR_FAIL(c, "mr_search_dissimilar: lower table without ssd / pairs");
      if (!(lower[q].ssd[r] < MR_UPPER_BOUND)) continue;   // no stored alignment: nothing to be similar to
      const int n = c->hp.n_min + r;
      const int16_t* pr = lower[q].pairs + (size_t)r * 2 * c->N;
      for (int k = 0; k < n; k++) {
        if (pr[2 * k] < 0 || pr[2 * k] >= c->m || pr[2 * k + 1] < 0 || pr[2 * k + 1] >= c->nB)
          MR_FAIL(c, "mr_search_dissimilar: pair outside the molecules in a lower table");
        partner[((size_t)q * c->R + r) * c->m + pr[2 * k]] = pr[2 * k + 1];
      }
    }
  if (need > c->excl_partner_cap) {
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_excl_partner);
    c->d_excl_partner = nullptr;
    c->excl_partner_cap = 0;
    CU(c, cudaMalloc(&c->d_excl_partner, 2 * need));
    c->excl_partner_cap = need;
  }
  CU(c, cudaMemcpyAsync(c->d_excl_partner, partner.data(), 2 * need, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->st.h2d_bytes += 2 * (int64_t)need;
  c->excl_ranks = n_lower;
  c->excl_max_common = max_pairs_common;
  c->excl_on = true;
  const int rc = mr_search(c, orient12, count, first_id, best);
  c->excl_on = false;
  return rc;
}

int mr_eval_one(mr_ctx* c, const float* T12, int32_t apply_filter, int32_t* n_last,
                int32_t* nw_matches, float* cost, uint8_t* valid, float* ssd, float* T, int16_t* pairs) {
  if (!c || !T12) return 1;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(c->d_orient_w, T12, 48, cudaMemcpyHostToDevice, c->stream));
  c->st.h2d_bytes += 48;
  return eval_w0(c, apply_filter != 0, n_last, nw_matches, cost, valid, ssd, T, pairs);
}

// lib/global_utils/simple_numeric_utils.h:350-376
static int find_median(const std::vector<char>& in_set, int n_min, int n_max) {
  int num = 0;
  for (int n = n_min; n <= n_max; n++) num += in_set[n - n_min] ? 1 : 0;
  if (num == 0) return n_min - 1;
  const int half = (num + 1) / 2;
  int n = n_min - 1, count = 0;
  do {
    ++n;
    if (in_set[n - n_min]) ++count;
  } while (count < half);
  return n;
}

// ---- Refine (SearchOrDyn3d::Refine, search_or_dyn3d.cc:2318-2638) -------------------------
// The reference loop is sequential: pick the median flagged size n, superimpose its current
// alignment, run the whole DP at that orientation, take every strictly better size, flag the
// sizes that changed.  One evaluation depends only on the alignment it starts from, so the
// evaluations of ALL currently flagged sizes are launched as one batch (one orientation per
// CTA: a batch costs about the latency of a single evaluation) and the loop is replayed on
// the host in the reference's order.  A batch result is used only while the alignment it was
// computed from is still the stored one; otherwise a new batch is launched.  Same
// evaluations, same order, same comparisons as the sequential loop.
static int refine_alloc(mr_ctx* c) {
  if (c->ref_kcap > 0) return 0;
  const size_t N = (size_t)c->N, NP = (size_t)c->NP;
  const size_t per_item = NP * 48 + 4 * N;
  int kcap = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->R, ((size_t)1 << 30) / per_item));
  CU(c, cudaMalloc(&c->d_ref_pairs_in, sizeof(short2) * N * kcap));
  CU(c, cudaMalloc(&c->d_ref_n, sizeof(int) * kcap));
  CU(c, cudaMalloc(&c->d_ref_orient, 48 * (size_t)kcap));
  CU(c, cudaMalloc(&c->d_ref_T, 48 * NP * kcap));
  // The alignments of a batch are NOT written out per (item, size): the traceback threads'
  // scratch slices (warp-interleaved, [warp][pair][lane]) hold them already, and only the few
  // items the replay consumes are fetched -- as the raw slices of the warps that cover the
  // item's sizes -- and de-interleaved on the host.
  const int per = std::max(1, c->hp.n_max - std::max(2, c->hp.n_min) + 1);
  c->ref_slice_warps = (per + 31) / 32 + 1;
  // pinned staging: pairs_in | n | ssd | valid | n_last | T (one item) | raw pair slices (one item)
  c->ref_off_n = sizeof(short2) * N * kcap;
  c->ref_off_ssd = c->ref_off_n + sizeof(int) * kcap;
  c->ref_off_valid = c->ref_off_ssd + 4 * NP * kcap;
  c->ref_off_nlast = (c->ref_off_valid + NP * kcap + 15) & ~(size_t)15;
  c->ref_off_T = c->ref_off_nlast + sizeof(int) * kcap;
  c->ref_off_pairs = (c->ref_off_T + 48 * NP + 15) & ~(size_t)15;
  CU(c, cudaMallocHost(&c->h_ref, c->ref_off_pairs + sizeof(short2) * 32 * N * (size_t)c->ref_slice_warps));
  c->ref_kcap = kcap;
  return 0;
}

// Refine batch item b = traceback threads [b * per, (b + 1) * per): copy the scratch slices of
// the warps that cover them to the pinned buffer (asynchronous).
static int refine_fetch_slices(mr_ctx* c, int b) {
  const int per = std::max(1, c->hp.n_max - std::max(2, c->hp.n_min) + 1);
  const size_t w_lo = ((size_t)b * per) >> 5, w_hi = ((size_t)(b + 1) * per - 1) >> 5;
  const size_t bytes = sizeof(short2) * 32 * (size_t)c->N * (w_hi - w_lo + 1);
  CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_pairs, c->d_scratch + w_lo * 32 * (size_t)c->N, bytes,
                        cudaMemcpyDeviceToHost, c->stream));
  c->st.d2h_bytes += (int64_t)bytes;
  return 0;
}
// pair list of (item b, size m2) from the fetched slices -> dst[2 * m2]
static void refine_unpack_pairs(const mr_ctx* c, int b, int m2, int16_t* dst) {
  const int n_lo = std::max(2, c->hp.n_min);
  const int per = std::max(1, c->hp.n_max - n_lo + 1);
  const size_t tid = (size_t)b * per + (size_t)(m2 - n_lo);
  const size_t w_lo = ((size_t)b * per) >> 5;
  const short2* raw = reinterpret_cast<const short2*>(c->h_ref + c->ref_off_pairs) +
                      ((tid >> 5) - w_lo) * 32 * (size_t)c->N + (tid & 31);
  for (int k = 0; k < m2; k++) {
    dst[2 * k] = raw[(size_t)k * 32].x;
    dst[2 * k + 1] = raw[(size_t)k * 32].y;
  }
}

// evaluate the stored alignments of the sizes n_min + r, r in `rs`, as one batch
static int refine_launch(mr_ctx* c, const std::vector<int>& rs, const mr_best* best, int first_item) {
  const int N = c->N, NP = c->NP, n_min = c->hp.n_min, k = (int)rs.size();
  short2* hp = reinterpret_cast<short2*>(c->h_ref);
  int* hn = reinterpret_cast<int*>(c->h_ref + c->ref_off_n);
  for (int b = 0; b < k; b++) {
    const int n = n_min + rs[b];
    hn[b] = n;
    memcpy(hp + (size_t)b * N, best->pairs + (size_t)rs[b] * 2 * N, sizeof(short2) * (size_t)n);
  }
  static const bool dbg = getenv("MR_DEBUG_REFINE") != nullptr;
  static cudaEvent_t e0 = nullptr, e1 = nullptr;
  const auto w0 = std::chrono::steady_clock::now();
  if (dbg) { if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); } cudaEventRecord(e0, c->stream); }
  CU(c, cudaMemcpyAsync(c->d_ref_pairs_in, hp, sizeof(short2) * (size_t)N * k, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_ref_n, hn, sizeof(int) * (size_t)k, cudaMemcpyHostToDevice, c->stream));
  c->st.h2d_bytes += (int64_t)k * (4 * N + 4);
  // orientation that optimally superimposes each alignment (:2401-2407)
  {
    const size_t team = 4 * 15 * (size_t)N * c->hp.nb;
    if (team <= (size_t)c->prop.sharedMemPerBlockOptin && !c->opt_refine_seq_superpose) {
      if (mr_ensure_smem(k_superimpose_pairs_batch, team)) MR_FAIL(c, "superposition kernel: shared memory attribute failed");
      k_superimpose_pairs_batch<<<k, 128, team, c->stream>>>(c->P, c->dA, c->dB, c->d_ref_pairs_in, c->d_ref_n, N,
                                                            c->d_ref_orient);
    } else {
      // chains too long for the team kernel's 60 bytes per point (N * nb > ~3 800)
      const size_t seq = 4 * 3 * (size_t)(c->m + c->nB) * c->hp.nb + 4 * (size_t)N;
      if (seq > (size_t)c->prop.sharedMemPerBlockOptin)
        MR_FAIL(c, "Refine: the coordinates of both molecules (%zu bytes) do not fit shared memory", seq);
      if (mr_ensure_smem(k_superimpose_pairs_batch_seq, seq)) MR_FAIL(c, "superposition kernel: shared memory attribute failed");
      k_superimpose_pairs_batch_seq<<<k, 128, seq, c->stream>>>(c->P, c->dA, c->dB, c->d_ref_pairs_in, c->d_ref_n, N,
                                                              c->d_ref_orient);
    }
  }
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  c->surv = c->d_surv;
  k_refine_prep<<<(k * NP + 255) / 256, 256, 0, c->stream>>>(c->d_surv, c->d_surv_count, k, NP, c->d_cost, c->d_ssd,
                                                            c->d_valid, c->d_nlast);
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  if (launch_dp(c, c->d_ref_orient, k, true)) return 1;
  if (launch_tb<true>(c, c->d_ref_orient, k, nullptr, 0, c->d_ref_T, nullptr, true)) return 1;
  k_select_warp<<<(k + 3) / 4, 128, 0, c->stream>>>(c->P, NP, c->d_surv, c->d_surv_count, c->d_cost, c->d_ssd, 0ll,
                                                  nullptr, c->d_valid, c->d_nlast);
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_ssd, c->d_ssd, 4 * (size_t)NP * k, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_valid, c->d_valid, (size_t)NP * k, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_nlast, c->d_nlast, 4 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  // the item consumed first (the median that triggered the batch) usually improves something:
  // its transforms and alignments ride along instead of costing a round trip later
  CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_T, c->d_ref_T + 12 * (size_t)NP * first_item, 48 * (size_t)NP,
                        cudaMemcpyDeviceToHost, c->stream));
  if (refine_fetch_slices(c, first_item)) return 1;
  c->ref_host_item = first_item;
  c->st.d2h_bytes += (int64_t)NP * 48;
  if (dbg) cudaEventRecord(e1, c->stream);
  CU(c, cudaStreamSynchronize(c->stream));
  if (dbg) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    float dp = 0.f, tb = 0.f;
    for (size_t q = 0; q < c->events_used; q++) {
      float t = 0.f;
      cudaEventElapsedTime(&t, c->events[q].a, c->events[q].b);
      (c->events[q].kind == 0 ? dp : tb) += t;
    }
    fprintf(stderr, "[mr] refine batch: %3d items, device %.3f ms (dp %.3f, traceback %.3f), wall %.3f ms\n", k, ms, dp, tb,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count());
  }
  c->st.d2h_bytes += (int64_t)k * (5 * NP + 4);
  c->events_used = 0;  // Refine evaluations are not part of the search throughput stats
  c->st.refine_batches++;
  c->st.refine_evals_launched += k;
  return 0;
}

int mr_refine(mr_ctx* c, mr_best* best, int32_t max_iters, int32_t* n_evals) {
  if (!c || !best) return 1;
  CU(c, cudaSetDevice(c->device));
  const int N = c->N, NP = c->NP, n_min = c->hp.n_min, n_max = c->hp.n_max, R = c->R;
  std::vector<char> need((size_t)R);
  for (int r = 0; r < R; r++) need[r] = best->ssd[r] < MR_UPPER_BOUND;  // :2348-2358
  if (refine_alloc(c)) return 1;
  std::vector<int> item_of((size_t)R, -1);  // batch item evaluated from the CURRENT alignment of a size
  std::vector<int> rs;
  std::vector<int16_t> old((size_t)2 * N);
  const float* h_ssd = reinterpret_cast<const float*>(c->h_ref + c->ref_off_ssd);
  const uint8_t* h_valid = reinterpret_cast<const uint8_t*>(c->h_ref + c->ref_off_valid);
  const int* h_nlast = reinterpret_cast<const int*>(c->h_ref + c->ref_off_nlast);
  const float* h_T = reinterpret_cast<const float*>(c->h_ref + c->ref_off_T);
  int evals = 0;
  for (int it = 1; it <= max_iters || max_iters < 0; it++) {
    const int n = find_median(need, n_min, n_max);
    if (n == n_min - 1) break;
    const int rn = n - n_min;
    if (item_of[rn] < 0) {
      // new batch: every flagged size (the window of flagged sizes around n if they do not fit)
      if (ensure_chunk(c, std::min(R, c->ref_kcap))) return 1;
      // at most one wave of the DP kernel: up to that size a batch costs the time of one evaluation
      const int path = dp_path(c);
      const int wave = path == 2 ? c->tiled_mp.grid_cap : path == 1 ? c->tiled.grid_cap : c->dp_grid_cap;
      int kmax = std::min(std::min(c->ref_kcap, c->chunk_cap), std::max(wave, 1));
      // split mode of the multi-pass kernel: one cluster of CTAs per item, fewer items per wave
      if (path == 2 && c->opt_refine_split && c->tiled_mp.split_cap > 0) kmax = std::min(kmax, c->tiled_mp.split_cap);
      if (c->opt_refine_batch > 0) kmax = std::min<int>(kmax, (int)c->opt_refine_batch);
      std::vector<int> flagged;
      int pos = 0;
      for (int r = 0; r < R; r++)
        if (need[r]) { if (r == rn) pos = (int)flagged.size(); flagged.push_back(r); }
      int lo = std::max(0, pos - kmax / 2);
      lo = std::min(lo, std::max(0, (int)flagged.size() - kmax));
      rs.assign(flagged.begin() + lo, flagged.begin() + std::min<int>((int)flagged.size(), lo + kmax));
      if (refine_launch(c, rs, best, pos - lo)) return 1;
      std::fill(item_of.begin(), item_of.end(), -1);
      for (int b = 0; b < (int)rs.size(); b++) item_of[rs[b]] = b;
    }
    const int b = item_of[rn];
    item_of[rn] = -1;
    evals++;
    int16_t* cur = best->pairs + (size_t)rn * 2 * N;
    memcpy(old.data(), cur, sizeof(int16_t) * 2 * (size_t)n);
    const int nl = h_nlast[b];
    bool fetched = false;
    for (int m2 = std::max(2, n_min); m2 <= nl; m2++) {
      const size_t slot = (size_t)b * NP + (m2 - 1);
      if (m2 > n_max || !h_valid[slot] || !(h_ssd[slot] < best->ssd[m2 - n_min])) continue;  // strict <, :2515
      if (!fetched) {  // transforms and alignments of this item, once
        if (c->ref_host_item != b) {
          CU(c, cudaMemcpyAsync(c->h_ref + c->ref_off_T, c->d_ref_T + 12 * (size_t)NP * b, 48 * (size_t)NP,
                                cudaMemcpyDeviceToHost, c->stream));
          if (refine_fetch_slices(c, b)) return 1;
          CU(c, cudaStreamSynchronize(c->stream));
          c->st.d2h_bytes += (int64_t)NP * 48;
          c->ref_host_item = b;
        }
        fetched = true;
      }
      const int r = m2 - n_min;
      best->ssd[r] = h_ssd[slot];
      memcpy(best->T + 12 * r, h_T + 12 * (size_t)(m2 - 1), 48);
      refine_unpack_pairs(c, b, m2, best->pairs + (size_t)r * 2 * N);
      if (best->o_id) best->o_id[r] = -2;
      if (m2 != n) need[r] = 1;  // :2546
      item_of[r] = -1;           // its alignment changed: a result computed from the old one is void
    }
    need[rn] = memcmp(old.data(), cur, sizeof(int16_t) * 2 * (size_t)n) != 0;  // :2619-2629
  }
  if (n_evals) *n_evals = evals;
  return 0;
}

int mr_superimpose_batch(mr_ctx* c, int32_t batch, int32_t np, const float* fixed,
                         const float* moveable, float* T, float* rmsd) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  if (batch <= 0 || np <= 0) return 0;
  const size_t pts = ((size_t)batch * np * 3 * 4 + 255) & ~(size_t)255, bT = ((size_t)batch * 48 + 255) & ~(size_t)255;
  char* base = nullptr;
  if (misc_reserve(c, 2 * pts + bT + (size_t)batch * 4, &base)) return 1;
  float *df = reinterpret_cast<float*>(base), *dm = reinterpret_cast<float*>(base + pts),
        *dT = reinterpret_cast<float*>(base + 2 * pts), *dr = reinterpret_cast<float*>(base + 2 * pts + bT);
  CU(c, cudaMemcpyAsync(df, fixed, (size_t)batch * np * 12, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(dm, moveable, (size_t)batch * np * 12, cudaMemcpyHostToDevice, c->stream));
  k_superimpose_batch<<<(batch + 63) / 64, 64, 0, c->stream>>>(batch, np, df, dm, dT, dr);
  CU(c, cudaGetLastError());
  c->st.kernel_launches++;
  CU(c, cudaMemcpyAsync(T, dT, (size_t)batch * 48, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(rmsd, dr, (size_t)batch * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// The probability of LevittGerstein98::Sstr_Probability (lib/pair_alignment/
// pair_alignment.cc:2798-2925) from the score the GPU produced: a closed form per size in the
// reference's long double arithmetic (its constants are double literals; log of a float and
// of doubles are the float / double functions, exp of a long double is expl).
static void sstr_probability(float S_str, int n, float out4[4]) {
  const long double c = 18.411424, d = -4.501719, e = 2.637163, f = 21.351857, g = -37.521269;
  const long double ln120 = log(120.0);
  const long double a = 2.0 * ln120 * c + d;
  const long double b = ln120 * ln120 * c + ln120 * d + e - a * ln120;
  const long double lnN = logf((float)n);
  const long double S = S_str;
  long double mu, delta;
  if (n < 120) {
    mu = c * (lnN * lnN);
    mu += d * lnN;
    mu += e;
    delta = f * lnN;
    delta += g;
  } else {
    mu = a * lnN;
    mu += b;
    delta = f * ln120 + g;
  }
  const long double Z = (S - mu) / delta;
  const long double result = (Z < log(10000.0 / 2)) ? 1.0f - expl(-expl(-Z)) : expl(-Z);
  out4[0] = (float)mu;
  out4[1] = (float)delta;
  out4[2] = (float)Z;
  out4[3] = (float)result;
}

int mr_plot_stats(mr_ctx* c, const mr_best* best, int32_t ca_index, double* rows) {
  if (!c) return 1;
  if (!best || !best->ssd || !best->T || !best->pairs || !rows) MR_FAIL(c, "mr_plot_stats: null argument");
  const int nb = c->hp.nb, R = c->R, N = c->N;
  if (ca_index < 0 || ca_index >= nb)
    MR_FAIL(c, "mr_plot_stats: the C-alpha atom must be one of the %d atoms used per residue (index %d)", nb, ca_index);
  CU(c, cudaSetDevice(c->device));
  for (int r = 0; r < R; r++) {
    if (!(best->ssd[r] < MR_UPPER_BOUND)) continue;
    int pi = -1, pj = -1;
    for (int k = 0; k < c->hp.n_min + r; k++) {
      const int i = best->pairs[(size_t)r * 2 * N + 2 * k], j = best->pairs[(size_t)r * 2 * N + 2 * k + 1];
      if (i <= pi || j <= pj || i >= c->m || j >= c->nB)
        MR_FAIL(c, "mr_plot_stats: size %d: pair %d = (%d,%d) is not an ascending alignment", c->hp.n_min + r, k, i, j);
      pi = i; pj = j;
    }
  }
  const size_t b_ssd = (size_t)R * 4, b_T = (size_t)R * 48, b_pairs = (size_t)R * 2 * N * 2;
  char* d = nullptr;
  const size_t off_T = (b_ssd + 15) & ~(size_t)15, off_pairs = off_T + b_T, off_far = (off_pairs + b_pairs + 15) & ~(size_t)15,
               off_s = off_far + b_ssd, off_has = off_s + b_ssd, total = off_has + b_ssd;
  if (misc_reserve(c, total, &d)) return 1;
  if (mr_ensure_smem(k_plot_stats, 2 * (size_t)N * sizeof(float))) MR_FAIL(c, "plot statistics kernel: shared memory attribute failed");
  std::vector<float> far(R), sstr(R);
  std::vector<int> has(R);
  cudaError_t e = cudaMemcpyAsync(d, best->ssd, b_ssd, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_T, best->T, b_T, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_pairs, best->pairs, b_pairs, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_plot_stats<<<R, 32, 2 * (size_t)N * sizeof(float), c->stream>>>(
        nb, ca_index, c->hp.n_min, N, c->dA, c->dB, (const float*)d, (const float*)(d + off_T),
        (const short*)(d + off_pairs), (float*)(d + off_far), (float*)(d + off_s), (int*)(d + off_has));
    e = cudaGetLastError();
    c->st.kernel_launches++;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(far.data(), d + off_far, b_ssd, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sstr.data(), d + off_s, b_ssd, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(has.data(), d + off_has, b_ssd, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  CU(c, e);
  c->st.h2d_bytes += (int64_t)(b_ssd + b_T + b_pairs);
  c->st.d2h_bytes += (int64_t)(3 * b_ssd);
  for (int r = 0; r < R; r++) {
    double* o = rows + 9 * (size_t)r;
    for (int k = 0; k < 9; k++) o[k] = 0.0;
    if (!has[r]) continue;
    const int n = c->hp.n_min + r;
    float s4[4];
    sstr_probability(sstr[r], n, s4);
    o[0] = 1.0;
    o[1] = sqrtf(best->ssd[r] / (n * nb));  // search_or_dyn3d.cc:2719-2721
    o[2] = far[r];
    o[3] = sstr[r];
    o[4] = s4[0]; o[5] = s4[1]; o[6] = s4[2]; o[7] = s4[3];
    o[8] = (s4[3] == 0.0) ? 0.0 : -log10f(s4[3]);  // :2755
  }
  return 0;
}

// glibc rand() in its never-seeded state (TYPE_3 additive feedback generator, seed 1):
// std::random_shuffle in lib/or_gen/or_gen.cc:179 is driven by it.
namespace {
struct GlibcRand {
  int32_t r[34];
  int f = 3, b = 0;
  GlibcRand() {
    int32_t word = 1;
    r[0] = word;
    for (int i = 1; i < 31; i++) {
      long hi = word / 127773, lo = word % 127773;
      long w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      word = (int32_t)w;
      r[i] = word;
    }
    for (int i = 0; i < 310; i++) next();
  }
  int next() {
    uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
    r[f] = (int32_t)v;
    f = (f + 1) % 31;
    b = (b + 1) % 31;
    return (int)(v >> 1);
  }
};
}  // namespace

int mr_gen_orientations(mr_ctx* c, int32_t L, float cutoff_rmsd, const int32_t* intervals,
                        int32_t n_intervals, int32_t shuffle, float* out12, int64_t cap, int64_t* count) {
  if (!c || !count) return 1;
  CU(c, cudaSetDevice(c->device));
  const int m = c->m, nB = c->nB, nb = c->hp.nb;
  if (L < 1 || L > c->N) MR_FAIL(c, "fragment size must be in [1, min(m,n)]");
  const int w1 = m - c->hp.n_min + 1, w2 = nB - c->hp.n_min + 1;  // or_gen.cc:253-254
  std::vector<char> mask((size_t)m * nB, 0);
  int32_t whole[4] = {0, m - 1, 0, nB - 1};
  if (!intervals || n_intervals == 0) { intervals = whole; n_intervals = 1; }
  for (int p = 0; p < n_intervals; p++) {  // or_gen.cc:345-493
    const int s1 = intervals[4 * p], e1 = intervals[4 * p + 1], s2 = intervals[4 * p + 2], e2 = intervals[4 * p + 3];
    if (s1 < 0 || s2 < 0 || e1 >= m || e2 >= nB || e1 < s1 || e2 < s2) MR_FAIL(c, "bad fragment-matching interval");
    if (s1 > e1 - L + 1 || s2 > e2 - L + 1) continue;
    for (int I = s1; I <= e1 - L + 1; I++) {
      const int jlo = std::max(s2, I - (w1 - 1)), jhi = std::min(I + (w2 - 1), e2 - L + 1);
      for (int J = jlo; J <= jhi; J++) mask[(size_t)I * nB + J] = 1;
    }
  }
  std::vector<int2> ij;
  for (int I = 0; I < m; I++)
    for (int J = 0; J < nB; J++)
      if (mask[(size_t)I * nB + J]) ij.push_back(make_int2(I, J));
  const int64_t cand = (int64_t)ij.size();
  std::vector<float> T((size_t)cand * 12), rm((size_t)cand);
  if (cand > 0) {
    const size_t b_ij = (sizeof(int2) * (size_t)cand + 255) & ~(size_t)255, b_T = (48 * (size_t)cand + 255) & ~(size_t)255;
    char* base = nullptr;
    if (misc_reserve(c, b_ij + b_T + 4 * (size_t)cand, &base)) return 1;
    int2* dij = reinterpret_cast<int2*>(base);
    float *dT = reinterpret_cast<float*>(base + b_ij), *dr = reinterpret_cast<float*>(base + b_ij + b_T);
    CU(c, cudaMemcpyAsync(dij, ij.data(), sizeof(int2) * (size_t)cand, cudaMemcpyHostToDevice, c->stream));
    k_fragment_fit<<<(unsigned)((cand + 63) / 64), 64, 0, c->stream>>>(c->dA, c->dB, nb, L, dij, cand, dT, dr);
    CU(c, cudaGetLastError());
    c->st.kernel_launches++;
    CU(c, cudaMemcpyAsync(T.data(), dT, 48 * (size_t)cand, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(rm.data(), dr, 4 * (size_t)cand, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  std::vector<int64_t> keep;
  for (int64_t k = 0; k < cand; k++)
    if (cutoff_rmsd < 0.0f || rm[k] <= cutoff_rmsd) keep.push_back(k);  // or_gen.cc:537-541
  const int64_t n = (int64_t)keep.size();
  *count = n;
  if (!out12) return 0;
  if (cap < n) MR_FAIL(c, "mr_gen_orientations: output capacity %lld < %lld", (long long)cap, (long long)n);
  if (shuffle) {  // libstdc++ std::random_shuffle(first,last): j = rand() % (i+1); swap
    GlibcRand g;
    for (int64_t i = 1; i < n; i++) {
      const int64_t j = (int64_t)g.next() % (i + 1);
      if (i != j) std::swap(keep[i], keep[j]);
    }
  }
  for (int64_t k = 0; k < n; k++) memcpy(out12 + 12 * k, &T[12 * (size_t)keep[k]], 48);
  return 0;
}

int mr_event_record(mr_ctx* c, int32_t slot) {
  if (!c || slot < 0 || slot >= 16) return 1;
  CU(c, cudaSetDevice(c->device));
  if (!c->user_ev[slot]) CU(c, cudaEventCreate(&c->user_ev[slot]));
  CU(c, cudaEventRecord(c->user_ev[slot], c->stream));
  return 0;
}

int mr_event_elapsed_ms(mr_ctx* c, int32_t a, int32_t b, float* ms) {
  if (!c || !ms || a < 0 || a >= 16 || b < 0 || b >= 16 || !c->user_ev[a] || !c->user_ev[b]) return 1;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(c->user_ev[b]));
  CU(c, cudaEventElapsedTime(ms, c->user_ev[a], c->user_ev[b]));
  return 0;
}

int mr_get_stats(mr_ctx* c, mr_stats* out) {
  if (!c || !out) return 1;
  *out = c->st;
  return 0;
}

int mr_reset_stats(mr_ctx* c) {
  if (!c) return 1;
  memset(&c->st, 0, sizeof c->st);
  return 0;
}

}  // extern "C"
