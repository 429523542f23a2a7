/*
 * mr_oracle_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The entry points of include/minrms_b200.h that minrms_b200/host/search_or_cuda.cc calls,
 * answered by the CPU oracle (oracle/minrms_oracle.c).  It exists so that the HOST logic of
 * SearchOrCuda -- table merging, the alternate-solutions bookkeeping (-alt-o), the file
 * writers -- can be checked against the unmodified reference binary in the CPU test suite
 * (tests/test_host_alternates_cpu.py), where no GPU exists.  Nothing under minrms_b200/
 * links it; the shipped minrms_cuda binary links libminrms_b200.so (CUDA, no CPU fallback).
 */
#include <stdlib.h>
#include <string.h>

#include "minrms_b200.h"
#include "minrms_oracle.h"

struct mr_ctx {
  orc_ws *ws;
  orc_params p;
  float *A, *B;
  int m, nB, N, R;
};

static orc_params to_orc(const mr_params *q)
{
  orc_params p;
  p.nb = q->nb; p.n_min = q->n_min; p.n_max = q->n_max; p.max_dist = q->max_dist;
  p.max_rmsd = q->max_rmsd; p.use_filter = q->use_filter; p.of_dist = q->of_dist;
  p.of_min = q->of_min; p.refine_while_searching = q->refine_while_searching;
  return p;
}

int mr_create(mr_ctx **out, const float *A, int32_t m, const float *B, int32_t nB,
              const mr_params *params, int32_t device)
{
  mr_ctx *c = (mr_ctx *)calloc(1, sizeof(mr_ctx));
  (void)device;
  c->p = to_orc(params);
  c->m = m; c->nB = nB; c->N = m < nB ? m : nB; c->R = params->n_max - params->n_min + 1;
  c->A = (float *)malloc(sizeof(float) * 3 * m * params->nb);
  c->B = (float *)malloc(sizeof(float) * 3 * nB * params->nb);
  memcpy(c->A, A, sizeof(float) * 3 * m * params->nb);
  memcpy(c->B, B, sizeof(float) * 3 * nB * params->nb);
  c->ws = orc_ws_create(m, nB, &c->p);
  *out = c;
  return 0;
}

void mr_destroy(mr_ctx *c)
{
  if (!c) return;
  orc_ws_destroy(c->ws);
  free(c->A); free(c->B); free(c);
}

const char *mr_last_error(const mr_ctx *c) { (void)c; return "oracle shim"; }

int mr_best_init(const mr_ctx *c, mr_best *best)
{
  orc_best b = {best->ssd, best->T, best->pairs, best->o_id};
  orc_best_init(&c->p, c->N, &b);
  return 0;
}

int mr_search(mr_ctx *c, const float *orient12, int64_t count, int64_t first_id, mr_best *best)
{
  orc_best b = {best->ssd, best->T, best->pairs, best->o_id};
  orc_search(c->ws, c->A, c->B, orient12, count, first_id, &b);
  return 0;
}

/* the same scores through the oracle's per-orientation body (search_or_dyn3d.cc:1691-1835) */
int mr_search_scores(mr_ctx *c, const float *orient12, int64_t count, int64_t first_id,
                     mr_best *best, float *scores)
{
  const int N = c->N, R = c->R;
  orc_eval e;
  int64_t o;
  int r;
  e.cost = (float *)malloc(sizeof(float) * N);
  e.valid = (uint8_t *)malloc(N);
  e.ssd = (float *)malloc(sizeof(float) * N);
  e.T = (float *)malloc(sizeof(float) * 12 * N);
  e.pairs = (int16_t *)malloc(sizeof(int16_t) * 2 * N * N);
  for (o = 0; o < count; o++) {
    orc_eval_orientation(c->ws, c->A, c->B, orient12 + 12 * o, 1, &e);
    for (r = 0; r < R; r++) {
      const int n = c->p.n_min + r;
      const int ok = e.n_last > 0 && n <= N && e.valid[n - 1] && e.ssd[n - 1] < ORC_UPPER_BOUND;
      scores[o * R + r] = ok ? e.ssd[n - 1] : ORC_UPPER_BOUND;
    }
  }
  free(e.cost); free(e.valid); free(e.ssd); free(e.T); free(e.pairs);
  return mr_search(c, orient12, count, first_id, best);
}

int mr_eval_one(mr_ctx *c, const float *T12, int32_t apply_filter, int32_t *n_last,
                int32_t *nw_matches, float *cost, uint8_t *valid, float *ssd, float *T,
                int16_t *pairs)
{
  const int N = c->N;
  orc_eval e;
  int n, k;
  e.cost = (float *)malloc(sizeof(float) * N);
  e.valid = (uint8_t *)malloc(N);
  e.ssd = (float *)malloc(sizeof(float) * N);
  e.T = (float *)malloc(sizeof(float) * 12 * N);
  e.pairs = (int16_t *)malloc(sizeof(int16_t) * 2 * N * N);
  orc_eval_orientation(c->ws, c->A, c->B, T12, apply_filter, &e);
  if (n_last) *n_last = e.n_last;
  if (nw_matches) *nw_matches = e.nw_matches;
  for (n = 1; n <= N; n++) {
    const int v = e.n_last > 0 && e.valid[n - 1];
    if (cost) cost[n - 1] = n <= e.n_last ? e.cost[n - 1] : 0.0f;
    if (valid) valid[n - 1] = (uint8_t)v;
    if (ssd) ssd[n - 1] = v ? e.ssd[n - 1] : 0.0f;
    if (T) for (k = 0; k < 12; k++) T[12 * (n - 1) + k] = v ? e.T[12 * (n - 1) + k] : 0.0f;
    if (pairs)
      for (k = 0; k < 2 * N; k++)
        pairs[(size_t)(n - 1) * 2 * N + k] = (v && k < 2 * n) ? e.pairs[(size_t)(n - 1) * 2 * N + k] : 0;
  }
  free(e.cost); free(e.valid); free(e.ssd); free(e.T); free(e.pairs);
  return 0;
}

int mr_refine(mr_ctx *c, mr_best *best, int32_t max_iters, int32_t *n_evals)
{
  orc_best b = {best->ssd, best->T, best->pairs, best->o_id};
  const int k = orc_refine(c->ws, c->A, c->B, &b, max_iters);
  if (n_evals) *n_evals = k;
  return 0;
}

int mr_plot_stats(mr_ctx *c, const mr_best *best, int32_t ca_index, double *rows)
{
  orc_best b = {best->ssd, best->T, best->pairs, best->o_id};
  orc_plot_stats(c->A, c->m, c->B, c->nB, c->p.nb, ca_index, c->p.n_min, c->R, c->N, &b, rows);
  return 0;
}

int mr_superimpose_batch(mr_ctx *c, int32_t batch, int32_t np, const float *fixed,
                         const float *moveable, float *T, float *rmsd)
{
  int b;
  (void)c;
  for (b = 0; b < batch; b++)
    rmsd[b] = orc_superimpose(np, fixed + (size_t)3 * np * b, moveable + (size_t)3 * np * b, T + 12 * b);
  return 0;
}

/* SearchOrDyn3d::MultiPass (search_or_dyn3d.cc:1975-2250) through the oracle's per-orientation
 * body: same early exits, a size is recorded only if its alignment shares at most
 * max_pairs_common pairs with every lower table's alignment of that size. */
int mr_search_dissimilar(mr_ctx *c, const float *orient12, int64_t count, int64_t first_id,
                         mr_best *best, int32_t n_lower, const mr_best *lower,
                         int32_t max_pairs_common)
{
  const int N = c->N, R = c->R;
  orc_eval e;
  int64_t o;
  int r, q, k, k2;
  e.cost = (float *)malloc(sizeof(float) * N);
  e.valid = (uint8_t *)malloc(N);
  e.ssd = (float *)malloc(sizeof(float) * N);
  e.T = (float *)malloc(sizeof(float) * 12 * N);
  e.pairs = (int16_t *)malloc(sizeof(int16_t) * 2 * N * N);
  for (o = 0; o < count; o++) {
    orc_eval_orientation(c->ws, c->A, c->B, orient12 + 12 * o, 1, &e);
    if (e.n_last <= 0) continue;
    for (r = 0; r < R; r++) {
      const int n = c->p.n_min + r;
      const int16_t *mine = e.pairs + (size_t)(n - 1) * 2 * N;
      int ok = 1;
      if (n > N || !e.valid[n - 1] || !(e.ssd[n - 1] < best->ssd[r])) continue;
      for (q = 0; q < n_lower && ok; q++) {
        const int16_t *theirs = lower[q].pairs + (size_t)r * 2 * N;
        int common = 0;
        if (!(lower[q].ssd[r] < ORC_UPPER_BOUND)) continue;
        for (k = 0; k < n; k++)
          for (k2 = 0; k2 < n; k2++)
            if (mine[2 * k] == theirs[2 * k2] && mine[2 * k + 1] == theirs[2 * k2 + 1]) common++;
        if (common > max_pairs_common) ok = 0;
      }
      if (!ok) continue;
      best->ssd[r] = e.ssd[n - 1];
      memcpy(best->T + 12 * r, e.T + 12 * (n - 1), 12 * sizeof(float));
      memcpy(best->pairs + (size_t)r * 2 * N, mine, sizeof(int16_t) * 2 * n);
      if (best->o_id) best->o_id[r] = first_id + o;
    }
  }
  free(e.cost); free(e.valid); free(e.ssd); free(e.T); free(e.pairs);
  return 0;
}

int64_t mr_dp_cells_per_orientation(int32_t m, int32_t nB, int32_t n_min, int32_t n_max)
{
  return orc_dp_cells(m, nB, n_min, n_max);
}

int mr_set_option(mr_ctx *c, const char *name, int64_t value)
{
  (void)c; (void)name; (void)value;   /* device work-space and kernel switches: nothing to do on the CPU */
  return 0;
}
