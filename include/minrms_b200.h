/*
 * minrms_b200.h -- C ABI of the B200-native MINRMS per-orientation search.
 *
 * Drop-in boundary: the reference has no FFI; its seam is the abstract C++ class
 * minrms::SearchOr (bin/minrms/search_or.h:57-130: Search(og,start,stop), Refine(),
 * WriteMSF(), WriteChimeraInfo()).  A `SearchOrCuda : SearchOr` (see INTEGRATION.md and
 * minrms_b200/host/search_or_cuda.cc) implements Search/Refine by calling the entry
 * points below; everything else of the reference (CLI, parser, PDB I/O, MSF writers)
 * is unchanged host code.
 *
 * Conventions
 *  - plain pointers and sizes only; caller owns every host buffer; the context owns
 *    device memory, streams and events.
 *  - every function returns 0 on success, non-zero on error; mr_last_error() returns
 *    the message (the reference prints to cerr and exit(-1), lib/global_utils/ERR.h:20-21;
 *    the C++ wrapper turns a non-zero return into ERR(...)).
 *  - coordinates are float[npoints][3] with npoints = residues * nb, exactly the arrays
 *    the reference flattens at bin/minrms/search_or_dyn3d.cc:186-224; A = molecule 1
 *    (fixed), B = molecule 2 (moved).
 *  - a transform is 12 floats, row-major Matrix3x4 (lib/vect_3d/vect_3d.h:66-92,
 *    lib/or_gen/or_gen.h:73-78): x' = M[r][0..2].x + M[r][3].
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MINRMS_B200_H
#define MINRMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MR_UPPER_BOUND 1.0e26f /* DistanceMetric::UPPER_BOUND, lib/dyn/distance_metric.cc:49 */

/* The scalars SearchOrDyn3d::Settings carries to the hot path
 * (bin/minrms/search_or_dyn3d.h:159-215, bin/minrms/search_or.h:149-308). */
typedef struct mr_params {
  int32_t nb;          /* atoms per residue, Biopolymer::Residue::NumBackboneAtoms() */
  int32_t n_min;       /* settings.n_min  (-minN) */
  int32_t n_max;       /* settings.n_max  (-maxN) */
  float   max_dist;    /* settings.dyn3d_max_dist, < 0 = none (-max-dist) */
  float   max_rmsd;    /* settings.dyn3d_max_rmsd, < 0 = none (-max-rmsd) */
  int32_t use_filter;  /* settings.orientation_filter_nw (-of / -OF) */
  float   of_dist;     /* settings.orientation_filter_nw_max_dist */
  int32_t of_min;      /* settings.orientation_filter_nw_min_num_matches */
  int32_t refine_while_searching; /* settings.refine_while_searching (-ir / -IR) */
} mr_params;

typedef struct mr_ctx mr_ctx;

/* Replaces the SearchOrDyn3d constructor (bin/minrms/search_or_dyn3d.cc:109-234):
 * uploads both coordinate arrays and sizes the device work space on CUDA device `device`. */
int mr_create(mr_ctx **out, const float *A, int32_t m, const float *B, int32_t nB,
              const mr_params *params, int32_t device);
void mr_destroy(mr_ctx *ctx);
/* message of the last failure on ctx (or of the last failed mr_create when ctx == NULL) */
const char *mr_last_error(const mr_ctx *ctx);

/* The per-size best table, exactly SearchOrDyn3d::aaSumSqdDists / aaTransforms /
 * aaAlignments[0][n - n_min] (bin/minrms/search_or_dyn3d.h:293-310).  R = n_max-n_min+1,
 * N = min(m, nB).  Slot r holds size n = n_min + r.
 *   ssd[r]          selection score (MR_UPPER_BOUND = no solution)
 *   T[12 r ..]      stored transform
 *   pairs[2 N r ..] n pairs (i,j), 0-based indices into A/B residues, ascending
 *   o_id[r]         id of the winning orientation (-1 none, -2 produced by mr_refine)
 */
typedef struct mr_best {
  float   *ssd;
  float   *T;
  int16_t *pairs;
  int64_t *o_id;
} mr_best;

/* Initialise a caller-allocated table (search_or_dyn3d.cc:165-184). */
int mr_best_init(const mr_ctx *ctx, mr_best *best);

/* Replaces SearchOrDyn3d::Search -> FirstPass over orientations
 * (bin/minrms/search_or_dyn3d.cc:283-299, 1643-1946) for the `count` transforms in
 * orient12 (host memory), whose ids are first_id, first_id+1, ...  The table is updated
 * with the reference's rule "strictly lower score wins, lowest id wins ties"
 * (search_or_dyn3d.cc:1825-1835), so calls over disjoint id ranges -- on one GPU or on
 * several, in any order -- merge to the sequential result.  Host->device copy of the
 * orientations and device->host copy of the table are inside the call. */
int mr_search(mr_ctx *ctx, const float *orient12, int64_t count, int64_t first_id,
              mr_best *best);

/* mr_search that also returns every score FirstPass compares: the input of the reference's
 * alternate-solutions bookkeeping (-alt-o / -alt-o-orig).  scores[(o - first_id) * R + r] is the
 * sum_sqd_dist handed to SearchOrDyn3d::KeepTrackOfAltSolutions for orientation o and size
 * n_min + r (bin/minrms/search_or_dyn3d.cc:477-542, called at :1897-1904), MR_UPPER_BOUND
 * where the reference does not get there (orientation rejected by the -of filter, n-loop left
 * early at :1714-1721 / :1810-1811).  The temp files the reference writes (one record per
 * kept score, :523-540) follow from this array and the running minimum over o; the
 * 3x4 matrix of a record is mr_eval_one's T for that orientation and size.  Host memory,
 * count * R floats, copied inside the call. */
int mr_search_scores(mr_ctx *ctx, const float *orient12, int64_t count, int64_t first_id,
                     mr_best *best, float *scores);

/* Replaces one pass of SearchOrDyn3d::MultiPass (bin/minrms/search_or_dyn3d.cc:1975-2250), the
 * search for the alternates of -alt-r (ALT_METHOD_PAIRS): mr_search in which an alignment of
 * size n is recorded only if, for every table q < n_lower, it shares at most
 * max_pairs_common residue pairs with lower[q]'s alignment of size n
 * (NotTooManyPairsInCommon, :352-365; NumMatchesInBothAssumingMonotonicity,
 * lib/pair_alignment/pair_alignment.cc:2529-2555).  Sizes a lower table has no solution for
 * constrain nothing.  The early exits of the n-loop (:2040, :2078) are those of mr_search.
 * lower[q].ssd and lower[q].pairs are read (host memory); T and o_id may be NULL. */
int mr_search_dissimilar(mr_ctx *ctx, const float *orient12, int64_t count, int64_t first_id,
                         mr_best *best, int32_t n_lower, const mr_best *lower,
                         int32_t max_pairs_common);

/* Same search with the orientation list already resident in device memory
 * (uploaded once with mr_upload_orientations); nothing is copied to the host until
 * mr_fetch_best.  Used to time the kernels alone. */
int mr_upload_orientations(mr_ctx *ctx, const float *orient12, int64_t count);
int mr_search_resident(mr_ctx *ctx, int64_t first, int64_t count, int64_t first_id);
int mr_reset_best(mr_ctx *ctx);
int mr_fetch_best(mr_ctx *ctx, mr_best *best);

/* Per-size sort keys of the device-side table: key[r] = (float_bits(ssd) << 32) | o_id.
 * Non-negative IEEE floats order like unsigned integers, so an element-wise integer MIN
 * over ranks (one small NCCL all-reduce) reproduces the reference's selection rule; the
 * rank whose key equals the reduced key owns the payload for that size. */
int mr_fetch_keys(mr_ctx *ctx, uint64_t *keys /* [R] */);

/* The same reduction without staging through the host (the reference's rule is
 * bin/minrms/search_or_dyn3d.cc:1825-1835).  Orientations are partitioned over ranks, one
 * context per GPU:
 *   1. mr_device_keys: device address of the R keys of this context (it also snapshots them);
 *      the caller reduces the buffer IN PLACE over the ranks with an element-wise MIN on
 *      64-bit integers (ncclAllReduce(ncclMin, ncclUint64) / torch.distributed.all_reduce on a
 *      tensor that aliases the buffer) and waits for it;
 *   2. mr_fetch_rows_device: the sizes whose reduced key equals the snapshot were won by an
 *      orientation of THIS context: only those are re-evaluated for their payload, written
 *      as rows of *ints_per_row = 2 + 12 + N int32 in device memory, row r (size n_min + r) =
 *      {key lo, key hi, transform (12 float bit patterns), N pairs as i | j << 16}; the rows
 *      of sizes won elsewhere carry the key of "no solution" (0x6aa56fa6ffffffff).
 *   3. one all-gather of the R rows and, per size, the row whose key is the minimum, give
 *      the merged table on every rank. */
int mr_device_keys(mr_ctx *ctx, uint64_t **d_keys, int32_t *count);
int mr_fetch_rows_device(mr_ctx *ctx, int32_t **d_rows, int32_t *ints_per_row);

/* One orientation, the body of the FirstPass / Refine n-loop, with every per-size
 * result returned (used by Refine and by the parity tests).  Arrays are indexed n-1:
 *   cost[N], valid[N] (1 = the reference would have compared this size), ssd[N],
 *   T[12 N], pairs[2 N N]; any pointer may be NULL.  *n_last = last layer the reference
 *   computes (0 = rejected by the -of filter), *nw_matches = NW filter count or -1. */
int mr_eval_one(mr_ctx *ctx, const float *T12, int32_t apply_filter, int32_t *n_last,
                int32_t *nw_matches, float *cost, uint8_t *valid, float *ssd, float *T,
                int16_t *pairs);

/* Replaces SearchOrDyn3d::Refine (bin/minrms/search_or_dyn3d.cc:2261-2658, REFINE_BEST,
 * one solution per size).  max_iters < 0 = until convergence.  *n_evals = number of
 * single-orientation evaluations performed. */
int mr_refine(mr_ctx *ctx, mr_best *best, int32_t max_iters, int32_t *n_evals);

/* Replaces OrientationGenerator::GenerateOrientationsUsingFragmentMatching + the
 * std::random_shuffle that follows (lib/or_gen/or_gen.cc:237-587, 179).
 * intervals: n_intervals x {startA, endA, startB, endB} inclusive 0-based residue
 * indices (NULL/0 = whole-vs-whole, i.e. -HS).  cutoff_rmsd < 0 = none.
 * Pass out12 = NULL to query the count.  Order = the reference's shuffled order when
 * shuffle != 0, row-major (I,J) otherwise. */
int mr_gen_orientations(mr_ctx *ctx, int32_t frag_len, float cutoff_rmsd,
                        const int32_t *intervals, int32_t n_intervals, int32_t shuffle,
                        float *out12, int64_t cap, int64_t *count);

/* Least-squares superposition of np point pairs (Superimpose::FindTransformMinRMSD,
 * lib/superimpose/superimpose.cc:17-287) for `batch` independent problems on the GPU.
 * fixed/moveable: batch x np x 3 floats; T: batch x 12; rmsd: batch. */
int mr_superimpose_batch(mr_ctx *ctx, int32_t batch, int32_t np, const float *fixed,
                         const float *moveable, float *T, float *rmsd);

/* Replaces the per-size loop of SearchOrDyn3d::WritePlotFile
 * (bin/minrms/search_or_dyn3d.cc:2703-2760), i.e. the rows of align_chimera.plot: for every
 * size of `best`, PairwiseAlignment::FindFurthestResPair and LevittGerstein98::Sstr_Score
 * (lib/pair_alignment/pair_alignment.cc:2351-2379, 2690-2790) run on the GPU at the stored
 * transform; the closed form of Sstr_Probability (2798-2925) is evaluated per size on the host
 * in the reference's long double arithmetic.  ca_index = position of the " CA" atom among
 * the nb atoms used per residue (Biopolymer::Residue::LookupBackboneAtomSymbol).
 * rows: R x 9 doubles {has solution, rmsd, longest distance, S_str, mu_str, delta_str, Z,
 * probability, -log10(probability)}; the reference prints columns n, 1, 2 and 8. */
int mr_plot_stats(mr_ctx *ctx, const mr_best *best, int32_t ca_index, double *rows);

/* Counters for the benchmark (reset by mr_reset_stats). */
typedef struct mr_stats {
  int64_t kernel_launches;   /* kernels of this library launched */
  int64_t orientations;      /* orientations submitted */
  int64_t orientations_dp;   /* orientations that reached the 3-D DP (passed -of) */
  int64_t dp_cells;          /* Dyn3d cell updates executed on the GPU */
  int64_t nw_cells;          /* NW filter cell updates executed on the GPU */
  double  dp_kernel_ms;      /* device time inside the 3-D DP kernel (CUDA events) */
  int64_t dp_kernel_launches;
  double  nw_kernel_ms;
  double  tb_kernel_ms;      /* traceback + Kabsch kernel */
  int64_t h2d_bytes, d2h_bytes;
  int64_t refine_batches;         /* batched launches made by mr_refine */
  int64_t refine_evals_launched;  /* evaluations in those batches (>= evaluations consumed) */
} mr_stats;
int mr_get_stats(mr_ctx *ctx, mr_stats *out);
int mr_reset_stats(mr_ctx *ctx);

/* CUDA events on the context's own stream (slots 0..15), so a caller can time a region
 * on the device: torch.cuda.Event only sees torch's stream, not this library's. */
int mr_event_record(mr_ctx *ctx, int32_t slot);
int mr_event_elapsed_ms(mr_ctx *ctx, int32_t slot_a, int32_t slot_b, float *ms);

/* Tuning knobs (by name): "dp_kernel" = 0 auto / 1 general / 2 tiled single-pass /
 * 3 tiled multi-pass; "budget_mb", "chunk" (work space per chunk of orientations);
 * "refine_batch" = 0 all flagged sizes per Refine batch / k at most k (1 = sequential);
 * "refine_split" = 0 keeps one CTA per orientation in the DP of Refine batches (default 1: the
 * passes of the multi-pass kernel run on a cluster of CTAs); "nw_kernel" = 1
 * thread-per-orientation filter; "mp_threads" (tests); "dp_form" = 1 runs the two-min / ping-pong
 * tile updates instead of the min-with-predicate kernels of the patched cubin (default 0);
 * "nw_cost_filter" = 0 keeps the counting NW kernels where the optimal cost would decide the
 * filter (default 1); "ir_costs_only" = 0 sends -IR searches through the full DP kernel. */
int mr_set_option(mr_ctx *ctx, const char *name, int64_t value);

/* What the context will run, by name: "dp_path" (0 general kernel, 1 tiled single-pass, 2 tiled
 * multi-pass), "dp_threads" (CTA size), "dp_vp_bucket" (CTA bucket of the min-with-predicate
 * kernel that FirstPass will launch, 0: the two-min / ping-pong forms), "dp_ctas_per_sm",
 * "dp_passes", "nw_decided_by_cost" (1: FirstPass filters with the decision-only kernels). */
int mr_get_info(mr_ctx *ctx, const char *name, int64_t *value);

/* number of Dyn3d cells per orientation for all layers 1..n_max */
int64_t mr_dp_cells_per_orientation(int32_t m, int32_t nB, int32_t n_min, int32_t n_max);

#ifdef __cplusplus
}
#endif
#endif
