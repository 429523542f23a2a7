/*
 * minrms_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, strict IEEE float: build with -O2 -ffp-contract=off,
 * never -ffast-math) of MINRMS's per-orientation search.  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path.  Nothing under minrms_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  Checked bit-for-bit against the unmodified reference built
 * with strict flags (oracle/_ref/strict, recipe oracle/build_ref.sh) through
 * oracle/ref_dump.cc on the reference's own test pair (configs 1 and 2), and against
 * the committed golden tables in tests/golden/ (see tests/test_oracle_golden.py).  The
 * align_chimera.plot statistics (orc_plot_stats) are pinned the same way: oracle/ref_plot.cc
 * evaluates the reference's own FindFurthestResPair / LevittGerstein98 functions on the
 * reference's best tables of five configurations (tests/golden/ref_plot.npz,
 * tests/test_plot_stats.py: all nine values per size bit-identical).
 *
 * Every function cites the reference file:line (under /root/reference) it follows.
 */
#ifndef MINRMS_ORACLE_H
#define MINRMS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_UPPER_BOUND 1.0e26f /* lib/dyn/distance_metric.cc:49 */

typedef struct orc_params {
  int32_t nb;          /* atoms per residue (Biopolymer::Residue::NumBackboneAtoms) */
  int32_t n_min;       /* -minN  (bin/minrms/minrms_parser.cc:329-346) */
  int32_t n_max;       /* -maxN  (minrms_parser.cc:1017-1036) */
  float   max_dist;    /* -max-dist, <0 = off (minrms_parser.cc:1045-1052) */
  float   max_rmsd;    /* -max-rmsd, <0 = off (minrms_parser.cc:1037-1044) */
  int32_t use_filter;  /* -of / -OF (minrms_parser.cc:1056-1103) */
  float   of_dist;
  int32_t of_min;
  int32_t refine_while_searching; /* -ir / -IR (minrms_parser.cc:748-752) */
} orc_params;

/* lib/superimpose/superimpose.cc:17-287 (+ lib/eigen/jacobi.c:7-85,
 * lib/vect_3d/vect_3d.cc:61-117).  fixed/moveable: np x 3 floats.  Returns rmsd. */
float orc_superimpose(int np, const float *fixed, const float *moveable, float T[12]);

/* lib/biopolymer/apply_trans.cc:72-108, lib/vect_3d/vect_3d.h:145-155 */
void orc_apply_transform(const float T[12], const float *src, int npoints, float *dst);

/* lib/or_gen/or_gen.cc:237-587 with one whole-vs-whole interval pair (-HS) or the
 * given interval pairs (each: startA,endA,startB,endB inclusive, 0-based residues).
 * Writes up to cap transforms (12 floats each) in the reference's row-major (I,J)
 * order; returns the count (which may exceed cap: nothing past cap is written). */
int64_t orc_gen_orientations(const float *A, int m, const float *B, int nB, int nb,
                             int frag_len, float cutoff_rmsd /* <0 = none */, int n_min,
                             const int32_t *intervals, int n_intervals,
                             float *out12, int64_t cap);

/* lib/or_gen/or_gen.cc:179: std::random_shuffle (libstdc++, driven by glibc rand()
 * in its never-seeded state).  Permutes `count` records of 12 floats in place. */
void orc_shuffle_orientations(float *orient12, int64_t count);
/* the permutation itself: perm[k] = original index that ends up at position k */
void orc_shuffle_perm(int64_t *perm, int64_t count);

/* lib/or_gen/orientation_filter.cc:42-130 + lib/dyn/dynNW.cc:147-531.
 * Bp = already-rotated molecule 2.  Returns the number of matches on the NW traceback. */
int orc_nw_num_matches(const float *A, int m, const float *Bp, int nB, int nb,
                       float of_dist, int of_min);

/* Per-orientation result of the FirstPass loop body
 * (bin/minrms/search_or_dyn3d.cc:1691-1835).  Arrays are indexed by n (1..N), slot n-1.
 *   n_last     : last layer computed by the n-loop (n_max if it ran to the end, the
 *                breaking layer otherwise, 0 if the NW filter rejected the orientation)
 *   cost[n-1]  : Dyn3d cost(n) for n <= n_last (layer 1 included)
 *   valid[n-1] : 1 iff (n >= n_min, n >= 2) and the reference would have compared this
 *                n against its best table (i.e. no early exit at or before n)
 *   ssd[n-1]   : the selection score (refined SSD if -ir else cost) when valid
 *   T[12*(n-1)]: stored transform when valid
 *   pairs[(n-1)*2*N + 2k + {0,1}] : alignment (i,j) 0-based, k < n, when valid
 */
typedef struct orc_eval {
  int32_t  n_last;
  int32_t  nw_matches; /* -1 if the filter is off */
  float   *cost;
  uint8_t *valid;
  float   *ssd;
  float   *T;
  int16_t *pairs;
} orc_eval;

/* work space for one pair (m, nB) */
typedef struct orc_ws orc_ws;
orc_ws *orc_ws_create(int m, int nB, const orc_params *p);
void    orc_ws_destroy(orc_ws *);

/* one orientation, exactly the body of the FirstPass loop (filter optional) */
void orc_eval_orientation(orc_ws *ws, const float *A, const float *B,
                          const float T12[12], int apply_filter, orc_eval *out);

/* Best table: slot r = n - n_min, r in [0, n_max-n_min] */
typedef struct orc_best {
  float   *ssd;   /* [R], ORC_UPPER_BOUND = none */
  float   *T;     /* [R*12] */
  int16_t *pairs; /* [R*2*N] */
  int64_t *o_id;  /* [R], -1 = none (Refine results get -2) */
} orc_best;

/* FirstPass over orientations [0,count) with ids first_id.. (search_or_dyn3d.cc:1643-1946).
 * best must be initialised by orc_best_init.  Returns the number of DP cells the
 * reference semantics execute (layers actually computed). */
void    orc_best_init(const orc_params *p, int N, orc_best *best);
int64_t orc_search(orc_ws *ws, const float *A, const float *B, const float *orient12,
                   int64_t count, int64_t first_id, orc_best *best);

/* bin/minrms/search_or_dyn3d.cc:2261-2658 (REFINE_BEST, num_solutions = 1).
 * max_iters < 0 = iterate until convergence.  Returns number of evaluations done. */
int orc_refine(orc_ws *ws, const float *A, const float *B, orc_best *best, int max_iters);

/* Statistics of align_chimera.plot for every size of a best table
 * (bin/minrms/search_or_dyn3d.cc:2703-2760; lib/pair_alignment/pair_alignment.cc:84-105,
 * 2351-2379, 2690-2925).  out: R x 9 doubles {has, rmsd, furthest distance, S_str, mu_str,
 * delta_str, Z, probability, -log10(probability)}; ca = index of " CA" among the nb atoms. */
void orc_plot_stats(const float *A, int m, const float *B, int nB, int nb, int ca, int n_min,
                    int R, int N, const orc_best *best, double *out);
void orc_sstr_probability(float S_str, int n, float out4[4] /* mu, delta, Z, probability */);

/* closed-form / counted number of Dyn3d cells for all layers 1..n_max */
int64_t orc_dp_cells(int m, int nB, int n_min, int n_max);

#ifdef __cplusplus
}
#endif
#endif
