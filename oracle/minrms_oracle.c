/*
 * minrms_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see minrms_oracle.h).
 *
 * Plain-C restatement of the MINRMS per-orientation search, written from the
 * behaviour of the reference (file:line cited per function, paths under
 * /root/reference).  Strict IEEE-754 binary32 arithmetic in the reference's source
 * order; build with:  gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared
 *
 * Parity status: PINNED against oracle/_ref/strict (the unmodified reference built
 * with -O2 -ffp-contract=off) via oracle/ref_dump.cc, and against tests/golden/.
 */
#include "minrms_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MINI(a, b) ((a) < (b) ? (a) : (b))
#define MAXI(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------- */
/* 4x4 cyclic Jacobi eigen-solver.
 * Follows the arithmetic of lib/eigen/jacobi.c:7-85 as compiled for C (float
 * variables, double literals => those sub-expressions are evaluated in double and
 * rounded to float on assignment; products of two floats stay float).
 * a: symmetric 4x4 (upper triangle is used and destroyed), d: eigenvalues,
 * v: eigenvectors in columns.  0-based here (the reference is 1-based).          */
static void jacobi4(float a[4][4], float d[4], float v[4][4])
{
  float b[4], z[4];
  int ip, iq, j, sweep;
  for (ip = 0; ip < 4; ip++) {
    for (iq = 0; iq < 4; iq++) v[ip][iq] = 0.0f;
    v[ip][ip] = 1.0f;
  }
  for (ip = 0; ip < 4; ip++) {
    b[ip] = d[ip] = a[ip][ip];
    z[ip] = 0.0f;
  }
  for (sweep = 1; sweep <= 50; sweep++) {
    float sm = 0.0f, tresh;
    for (ip = 0; ip < 3; ip++)
      for (iq = ip + 1; iq < 4; iq++)
        sm = (float)((double)sm + fabs((double)a[ip][iq])); /* jacobi.c:26-29 */
    if ((double)sm == 0.0) return;                           /* jacobi.c:30-34 */
    if (sweep < 4)
      tresh = (float)(0.2 * (double)sm / (double)(4 * 4));   /* jacobi.c:35-36 */
    else
      tresh = 0.0f;
    for (ip = 0; ip < 3; ip++) {
      for (iq = ip + 1; iq < 4; iq++) {
        float g = (float)(100.0 * fabs((double)a[ip][iq])); /* jacobi.c:41 */
        if (sweep > 4 &&
            (float)(fabs((double)d[ip]) + (double)g) == (float)fabs((double)d[ip]) &&
            (float)(fabs((double)d[iq]) + (double)g) == (float)fabs((double)d[iq])) {
          a[ip][iq] = 0.0f;                                  /* jacobi.c:42-44 */
        } else if (fabs((double)a[ip][iq]) > (double)tresh) {
          float h = d[iq] - d[ip];                           /* jacobi.c:46 */
          float t, theta, c, s, tau;
          if ((float)(fabs((double)h) + (double)g) == (float)fabs((double)h)) {
            t = a[ip][iq] / h;                               /* jacobi.c:48 */
          } else {
            float th2;
            theta = (float)(0.5 * (double)h / (double)a[ip][iq]);      /* :50 */
            th2 = theta * theta;                             /* float product */
            t = (float)(1.0 / (fabs((double)theta) + sqrt(1.0 + (double)th2)));
            if ((double)theta < 0.0) t = -t;                 /* jacobi.c:51-52 */
          }
          {
            float tt = t * t;
            float one_tt = 1.0f + tt;      /* "1+t*t": int 1 + float => float */
            c = (float)(1.0 / sqrt((double)one_tt));         /* jacobi.c:54 */
          }
          s = t * c;
          tau = (float)((double)s / (1.0 + (double)c));      /* jacobi.c:56 */
          h = t * a[ip][iq];
          z[ip] -= h;
          z[iq] += h;
          d[ip] -= h;
          d[iq] += h;
          a[ip][iq] = 0.0f;
#define ORC_ROTATE(M, i1, j1, k1, l1)            \
  {                                              \
    float gg = M[i1][j1], hh = M[k1][l1];        \
    M[i1][j1] = gg - s * (hh + gg * tau);        \
    M[k1][l1] = hh + s * (gg - hh * tau);        \
  }
          for (j = 0; j <= ip - 1; j++) ORC_ROTATE(a, j, ip, j, iq)
          for (j = ip + 1; j <= iq - 1; j++) ORC_ROTATE(a, ip, j, j, iq)
          for (j = iq + 1; j < 4; j++) ORC_ROTATE(a, ip, j, iq, j)
          for (j = 0; j < 4; j++) ORC_ROTATE(v, j, ip, j, iq)
#undef ORC_ROTATE
        }
      }
    }
    for (ip = 0; ip < 4; ip++) {
      b[ip] += z[ip];
      d[ip] = b[ip];
      z[ip] = 0.0f;
    }
  }
  /* the reference aborts here ("Too many iterations"); keep going with what we have */
}

/* lib/superimpose/superimpose.cc:17-287 (Diamond 1988 quaternion least squares). */
float orc_superimpose(int np, const float *fixed, const float *moveable, float T[12])
{
  float cf[3] = {0.0f, 0.0f, 0.0f}, cm[3] = {0.0f, 0.0f, 0.0f};
  float M[3][3] = {{0}}, Q[3][3], V[3], P[4][4], d[4], ev[4][4], p[4], R[3][3];
  float one_over_np = 1.0f / (float)np; /* vect_3d.cc:73-94 FindCentroid */
  float traceM, E0 = 0.0f, lam, len, inv, min_ssd, tr[3];
  int a, i, j, which;

  for (a = 0; a < np; a++) {
    cf[0] += fixed[3 * a + 0];
    cf[1] += fixed[3 * a + 1];
    cf[2] += fixed[3 * a + 2];
  }
  for (i = 0; i < 3; i++) cf[i] *= one_over_np;
  for (a = 0; a < np; a++) {
    cm[0] += moveable[3 * a + 0];
    cm[1] += moveable[3 * a + 1];
    cm[2] += moveable[3 * a + 2];
  }
  for (i = 0; i < 3; i++) cm[i] *= one_over_np;

  /* superimpose.cc:49-68: centred copies, M[i][j] += moveable'[a][i]*fixed'[a][j];
   * E0 (superimpose.cc:258-263) uses the same centred values, so do it in one pass. */
  for (a = 0; a < np; a++) {
    float f[3], mv[3];
    for (i = 0; i < 3; i++) {
      f[i] = fixed[3 * a + i] - cf[i];
      mv[i] = moveable[3 * a + i] - cm[i];
    }
    for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) M[i][j] += mv[i] * f[j];
    for (i = 0; i < 3; i++) {
      float df = f[i] - mv[i];
      E0 += df * df;
    }
  }
  traceM = M[0][0] + M[1][1] + M[2][2]; /* vect_3d.h:353-360 */
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++)
      Q[i][j] = M[i][j] + M[j][i] - ((i == j) ? (2 * traceM) : 0.0f); /* :75-82 */
  V[0] = M[1][2] - M[2][1];
  V[1] = M[2][0] - M[0][2];
  V[2] = M[0][1] - M[1][0];
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) P[i][j] = Q[i][j];
  P[0][3] = P[3][0] = V[0];
  P[1][3] = P[3][1] = V[1];
  P[2][3] = P[3][2] = V[2];
  P[3][3] = 0.0f;

  jacobi4(P, d, ev); /* superimpose.cc:177 */

  which = 0;
  lam = d[0];
  for (i = 1; i < 4; i++) /* superimpose.cc:187-196, strict > */
    if (d[i] > lam) {
      lam = d[i];
      which = i;
    }
  for (i = 0; i < 4; i++) p[i] = ev[i][which];
  /* vect_3d.cc:61-71 Normalize4 */
  len = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  inv = (float)(1.0 / (double)len);
  for (i = 0; i < 4; i++) p[i] *= inv;
  /* vect_3d.cc:97-117 ConvertQuaternionToMatrix3x3 */
  R[0][0] = +p[0] * p[0] - p[1] * p[1] - p[2] * p[2] + p[3] * p[3];
  R[1][1] = -(p[0] * p[0]) + p[1] * p[1] - p[2] * p[2] + p[3] * p[3];
  R[2][2] = -(p[0] * p[0]) - p[1] * p[1] + p[2] * p[2] + p[3] * p[3];
  R[0][1] = 2 * (p[0] * p[1] - p[2] * p[3]);
  R[1][0] = 2 * (p[0] * p[1] + p[2] * p[3]);
  R[1][2] = 2 * (p[1] * p[2] - p[0] * p[3]);
  R[2][1] = 2 * (p[1] * p[2] + p[0] * p[3]);
  R[0][2] = 2 * (p[0] * p[2] + p[1] * p[3]);
  R[2][0] = 2 * (p[0] * p[2] - p[1] * p[3]);
  /* superimpose.cc:245-256: t = c_fixed - R*c_moveable */
  for (i = 0; i < 3; i++) tr[i] = R[i][0] * cm[0] + R[i][1] * cm[1] + R[i][2] * cm[2];
  for (i = 0; i < 3; i++) {
    T[4 * i + 0] = R[i][0];
    T[4 * i + 1] = R[i][1];
    T[4 * i + 2] = R[i][2];
    T[4 * i + 3] = cf[i] - tr[i];
  }
  min_ssd = E0 - 2 * lam; /* superimpose.cc:266 */
  if (min_ssd < 0.0f) min_ssd = 0.0f;
  return sqrtf(min_ssd / (float)np); /* superimpose.cc:285 */
}

/* lib/vect_3d/vect_3d.h:145-155 / lib/biopolymer/apply_trans.cc:72-108 */
void orc_apply_transform(const float T[12], const float *src, int npoints, float *dst)
{
  int a;
  for (a = 0; a < npoints; a++) {
    float x = src[3 * a], y = src[3 * a + 1], z = src[3 * a + 2];
    float t0 = T[0] * x + T[1] * y + T[2] * z + T[3];
    float t1 = T[4] * x + T[5] * y + T[6] * z + T[7];
    float t2 = T[8] * x + T[9] * y + T[10] * z + T[11];
    dst[3 * a] = t0;
    dst[3 * a + 1] = t1;
    dst[3 * a + 2] = t2;
  }
}

/* lib/dyn/distance_metric.h:104-130 */
static inline float sq_dist(const float *A, int i, const float *B, int j, int nb, float cutoff)
{
  float out = 0.0f;
  int q;
  for (q = 0; q < nb; q++) {
    const float *a = A + 3 * (i * nb + q), *b = B + 3 * (j * nb + q);
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    out += dx * dx + dy * dy + dz * dz; /* vect_3d.h:268-276 */
  }
  if (out > cutoff) return ORC_UPPER_BOUND;
  return out;
}

/* lib/dyn/distance_metric.cc:76-87 */
static float metric_cutoff(float max_phys_dist, int nb)
{
  if (max_phys_dist == -1.0f) return ORC_UPPER_BOUND;
  return (max_phys_dist * max_phys_dist) * (float)nb;
}

/* ------------------------------------------------------------------------- */
/* glibc rand() in its default (never seeded => seed 1) TYPE_3 state, restated so the
 * shuffle does not depend on who else called rand() in this process.            */
typedef struct {
  int32_t r[34];
  int f, b; /* front / rear indices */
} glibc_rand;

static void glibc_rand_init(glibc_rand *g, unsigned seed)
{
  int i;
  int32_t word = (int32_t)seed;
  if (word == 0) word = 1;
  g->r[0] = word;
  for (i = 1; i < 31; i++) {
    long hi = word / 127773, lo = word % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    word = (int32_t)w;
    g->r[i] = word;
  }
  g->f = 3;
  g->b = 0;
  for (i = 0; i < 310; i++) {
    g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
    g->f = (g->f + 1) % 31;
    g->b = (g->b + 1) % 31;
  }
}

static int glibc_rand_next(glibc_rand *g)
{
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  g->f = (g->f + 1) % 31;
  g->b = (g->b + 1) % 31;
  return (int)(v >> 1);
}

/* lib/or_gen/or_gen.cc:179 -> libstdc++ std::random_shuffle(first,last):
 *   for i in [1,n): j = rand() % (i+1); if (i != j) swap(x[i], x[j])            */
void orc_shuffle_perm(int64_t *perm, int64_t count)
{
  glibc_rand g;
  int64_t i;
  glibc_rand_init(&g, 1u);
  for (i = 0; i < count; i++) perm[i] = i;
  for (i = 1; i < count; i++) {
    int64_t j = (int64_t)glibc_rand_next(&g) % (i + 1);
    if (i != j) {
      int64_t t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
    }
  }
}

void orc_shuffle_orientations(float *o, int64_t count)
{
  int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * (size_t)(count > 0 ? count : 1));
  float *tmp = (float *)malloc(sizeof(float) * 12 * (size_t)(count > 0 ? count : 1));
  int64_t k;
  orc_shuffle_perm(perm, count);
  memcpy(tmp, o, sizeof(float) * 12 * (size_t)count);
  for (k = 0; k < count; k++) memcpy(o + 12 * k, tmp + 12 * perm[k], sizeof(float) * 12);
  free(tmp);
  free(perm);
}

/* lib/or_gen/or_gen.cc:237-587 */
int64_t orc_gen_orientations(const float *A, int m, const float *B, int nB, int nb,
                             int L, float cutoff_rmsd, int n_min,
                             const int32_t *intervals, int n_intervals,
                             float *out12, int64_t cap)
{
  int w1 = m - n_min + 1, w2 = nB - n_min + 1; /* or_gen.cc:253-254 */
  unsigned char *mask = (unsigned char *)calloc((size_t)m * (size_t)nB, 1);
  int32_t whole[4];
  int p, I, J;
  int64_t count = 0;
  if (n_intervals == 0 || !intervals) { /* minrms_parser.cc:677-689 (-HS default) */
    whole[0] = 0; whole[1] = m - 1; whole[2] = 0; whole[3] = nB - 1;
    intervals = whole;
    n_intervals = 1;
  }
  for (p = 0; p < n_intervals; p++) { /* or_gen.cc:345-493 */
    int s1 = intervals[4 * p], e1 = intervals[4 * p + 1];
    int s2 = intervals[4 * p + 2], e2 = intervals[4 * p + 3];
    if (s1 > e1 - L + 1 || s2 > e2 - L + 1) continue; /* too small: skipped */
    for (I = s1; I <= e1 - L + 1; I++) {
      int jlo = MAXI(s2, I - (w1 - 1));
      int jhi = MINI(I + (w2 - 1), e2 - L + 1);
      for (J = jlo; J <= jhi; J++) mask[(size_t)I * nB + J] = 1;
    }
  }
  for (I = 0; I < m; I++)   /* or_gen.cc:507-544 */
    for (J = 0; J < nB; J++)
      if (mask[(size_t)I * nB + J]) {
        float T[12];
        float r = orc_superimpose(L * nb, A + 3 * I * nb, B + 3 * J * nb, T);
        if (cutoff_rmsd < 0.0f || r <= cutoff_rmsd) {
          if (count < cap) memcpy(out12 + 12 * count, T, sizeof T);
          count++;
        }
      }
  free(mask);
  return count;
}

/* ------------------------------------------------------------------------- */
/* Needleman-Wunsch orientation filter.
 * lib/or_gen/orientation_filter.cc:42-130, lib/dyn/dynNW.cc:67-135 (band),
 * 147-236 (borders), 243-395 (recurrence, tie order), 401-475 (costs), 480-531.  */
int orc_nw_num_matches(const float *A, int m, const float *Bp, int nB, int nb,
                       float of_dist, int of_min)
{
  /* orientation_filter.cc:77-79: 0.5 (double) * SQR(float) * int * int -> float */
  float G = (float)(0.5 * (double)(of_dist * of_dist) * (double)MINI(m, nB) * (double)nb);
  float cutoff = metric_cutoff(of_dist, nb); /* orientation_filter.cc:96-97 */
  int w0 = m - of_min + 1, w1 = nB - of_min + 1; /* dynNW.cc:107-108 */
  int ts = nB + 1;
  float *Aij = (float *)malloc(sizeof(float) * (size_t)(m + 1) * ts);
  unsigned char *prev = (unsigned char *)malloc((size_t)(m + 1) * ts);
  int i, j, ci, cj, nmatch = 0;
  enum { C_DIAG = 0, C_I1 = 1, C_J1 = 2, C_UNINIT = 3 };
  for (i = 0; i < (m + 1) * ts; i++) {
    Aij[i] = -1.0f;
    prev[i] = C_UNINIT;
  }
  Aij[0] = 0.0f;
  for (i = 1; i <= m; i++) { /* dynNW.cc:200-216 */
    j = i - w0;
    if (j < 0) j = 0;
    Aij[i * ts + j] = (float)(i + j) * G;
    prev[i * ts + j] = C_UNINIT;
  }
  for (j = 1; j <= nB; j++) { /* dynNW.cc:218-234 */
    i = j - w1;
    if (i < 0) i = 0;
    Aij[i * ts + j] = (float)(i + j) * G;
    prev[i * ts + j] = C_UNINIT;
  }
  for (i = 1; i <= m; i++) { /* dynNW.cc:290-389 */
    int jlo = (i <= w0) ? 1 : i - w0 + 1;
    int jhi = (i + w1 <= nB) ? i + w1 - 1 : nB;
    for (j = jlo; j <= jhi; j++) {
      float Cij = sq_dist(A, i - 1, Bp, j - 1, nb, cutoff);
      float c_match = Aij[(i - 1) * ts + (j - 1)] + Cij;
      float c_i = Aij[(i - 1) * ts + j] + G;
      float c_j = Aij[i * ts + (j - 1)] + G;
      float best = c_i;
      unsigned char code = C_I1;
      if (c_j < best) {
        best = c_j;
        code = C_J1;
      }
      if (c_match < best) {
        best = c_match;
        code = C_DIAG;
      }
      Aij[i * ts + j] = best;
      prev[i * ts + j] = code;
    }
  }
  ci = m;
  cj = nB;
  while (ci > 0 && cj > 0) { /* dynNW.cc:484-498 */
    unsigned char c = prev[ci * ts + cj];
    if (c == C_DIAG) {
      nmatch++;
      ci--;
      cj--;
    } else if (c == C_I1) {
      ci--;
    } else if (c == C_J1) {
      cj--;
    } else {
      ci = 0;
      cj = 0;
    }
  }
  free(Aij);
  free(prev);
  return nmatch;
}

/* ------------------------------------------------------------------------- */
/* Dyn3d work space: lookup table C, two cost layers, jagged traceback table.
 * lib/dyn/dyn3d.cc:443-512 (Alloc), 554-624 (NLayerTable::Alloc).            */
struct orc_ws {
  int m, nB, N;
  orc_params p;
  int wA, wB;
  float *Bp;      /* rotated molecule 2 */
  float *C;       /* m x nB */
  float *cost[2]; /* (m+1) x (nB+1), layer n&1 */
  size_t *tb_off; /* [n_max+1] offsets into tb */
  unsigned char *tb;
  float *gf, *gm; /* gather buffers for Kabsch */
  /* refine scratch */
  orc_eval ev;
};

static int layer_imax(const orc_ws *w, int n) { return MINI(w->m, n + w->wA - 1); }
static int layer_jmax(const orc_ws *w, int n) { return MINI(w->nB, n + w->wB - 1); }

static unsigned char *tb_at(const orc_ws *w, int n, int i, int j)
{
  int jw = layer_jmax(w, n) - n + 1;
  return w->tb + w->tb_off[n] + (size_t)(i - n) * jw + (j - n);
}

static void eval_alloc(orc_eval *e, int N)
{
  e->cost = (float *)calloc((size_t)N, sizeof(float));
  e->valid = (uint8_t *)calloc((size_t)N, 1);
  e->ssd = (float *)calloc((size_t)N, sizeof(float));
  e->T = (float *)calloc((size_t)N * 12, sizeof(float));
  e->pairs = (int16_t *)calloc((size_t)N * 2 * N, sizeof(int16_t));
}
static void eval_free(orc_eval *e)
{
  free(e->cost); free(e->valid); free(e->ssd); free(e->T); free(e->pairs);
}

orc_ws *orc_ws_create(int m, int nB, const orc_params *p)
{
  orc_ws *w = (orc_ws *)calloc(1, sizeof(orc_ws));
  int n;
  size_t off = 0;
  w->m = m;
  w->nB = nB;
  w->N = MINI(m, nB);
  w->p = *p;
  w->wA = m - p->n_min + 1; /* dyn3d.cc:472-473 */
  w->wB = nB - p->n_min + 1;
  w->Bp = (float *)malloc(sizeof(float) * 3 * (size_t)nB * p->nb);
  w->C = (float *)malloc(sizeof(float) * (size_t)m * nB);
  w->cost[0] = (float *)calloc((size_t)(m + 1) * (nB + 1), sizeof(float));
  w->cost[1] = (float *)calloc((size_t)(m + 1) * (nB + 1), sizeof(float));
  w->tb_off = (size_t *)calloc((size_t)p->n_max + 2, sizeof(size_t));
  for (n = 1; n <= p->n_max; n++) {
    w->tb_off[n] = off;
    off += (size_t)(layer_imax(w, n) - n + 1) * (size_t)(layer_jmax(w, n) - n + 1);
  }
  w->tb = (unsigned char *)malloc(off ? off : 1);
  w->gf = (float *)malloc(sizeof(float) * 3 * (size_t)w->N * p->nb);
  w->gm = (float *)malloc(sizeof(float) * 3 * (size_t)w->N * p->nb);
  eval_alloc(&w->ev, w->N);
  return w;
}

void orc_ws_destroy(orc_ws *w)
{
  if (!w) return;
  free(w->Bp); free(w->C); free(w->cost[0]); free(w->cost[1]);
  free(w->tb_off); free(w->tb); free(w->gf); free(w->gm);
  eval_free(&w->ev);
  free(w);
}

int64_t orc_dp_cells(int m, int nB, int n_min, int n_max)
{
  int wA = m - n_min + 1, wB = nB - n_min + 1, n;
  int64_t c = 0;
  for (n = 1; n <= n_max; n++)
    c += (int64_t)(MINI(m, n + wA - 1) - n + 1) * (int64_t)(MINI(nB, n + wB - 1) - n + 1);
  return c;
}

#define COST(w, n, i, j) ((w)->cost[(n) & 1][(size_t)(i) * ((w)->nB + 1) + (j)])

/* lib/dyn/dyn3d.cc:664-701 */
static void fill_lookup(orc_ws *w, const float *A)
{
  float cutoff = metric_cutoff(w->p.max_dist, w->p.nb); /* search_or_dyn3d.cc:124 */
  int i, j;
  for (i = 0; i < w->m; i++) {
    int jlo = (i < w->wA) ? 0 : i - w->wA + 1;
    int jhi = (i + w->wB < w->nB) ? i + w->wB : w->nB;
    for (j = jlo; j < jhi; j++)
      w->C[(size_t)i * w->nB + j] = sq_dist(A, i, w->Bp, j, w->p.nb, cutoff);
  }
}

/* lib/dyn/dyn3d.cc:135-179.  Returns cost(1). */
static float init_layer1(orc_ws *w)
{
  int i, j;
  for (i = 1; i <= w->wA; i++)
    for (j = 1; j <= w->wB; j++) {
      unsigned char code = 0;
      float best = w->C[(size_t)(i - 1) * w->nB + (j - 1)];
      if (i > 1 && COST(w, 1, i - 1, j) <= best) {
        best = COST(w, 1, i - 1, j);
        code = 1;
      }
      if (j > 1 && COST(w, 1, i, j - 1) <= best) {
        best = COST(w, 1, i, j - 1);
        code = 2;
      }
      COST(w, 1, i, j) = best;
      *tb_at(w, 1, i, j) = code;
    }
  return COST(w, 1, w->wA, w->wB);
}

/* lib/dyn/dyn3d.cc:183-387.  Returns cost(n) = A(n, i_max, j_max). */
static float solve_layer(orc_ws *w, int n)
{
  int imax = layer_imax(w, n), jmax = layer_jmax(w, n), i, j;
  for (j = n; j <= jmax; j++) { /* row i = n: match or skip-j (dyn3d.cc:197-220) */
    unsigned char code = 0;
    float best = COST(w, n - 1, n - 1, j - 1) + w->C[(size_t)(n - 1) * w->nB + (j - 1)];
    if (j > n && COST(w, n, n, j - 1) <= best) {
      best = COST(w, n, n, j - 1);
      code = 2;
    }
    COST(w, n, n, j) = best;
    *tb_at(w, n, n, j) = code;
  }
  for (i = n + 1; i <= imax; i++) {
    { /* column j = n: match or skip-i (dyn3d.cc:224-234) */
      unsigned char code = 0;
      float best = COST(w, n - 1, i - 1, n - 1) + w->C[(size_t)(i - 1) * w->nB + (n - 1)];
      if (COST(w, n, i - 1, n) <= best) {
        best = COST(w, n, i - 1, n);
        code = 1;
      }
      COST(w, n, i, n) = best;
      *tb_at(w, n, i, n) = code;
    }
    for (j = n + 1; j <= jmax; j++) { /* dyn3d.cc:283-382, tie order match<skip-i<skip-j */
      unsigned char code = 0;
      float best = COST(w, n - 1, i - 1, j - 1) + w->C[(size_t)(i - 1) * w->nB + (j - 1)];
      float up = COST(w, n, i - 1, j), left = COST(w, n, i, j - 1);
      if (up <= best) {
        best = up;
        code = 1;
      }
      if (left <= best) {
        best = left;
        code = 2;
      }
      COST(w, n, i, j) = best;
      *tb_at(w, n, i, j) = code;
    }
  }
  return COST(w, n, imax, jmax);
}

/* lib/dyn/dyn3d.cc:391-431 */
static void export_solution(const orc_ws *w, int n, int16_t *pairs)
{
  int cn = n, ci = w->m, cj = w->nB;
  while (cn > 0) {
    unsigned char c = *tb_at(w, cn, ci, cj);
    if (c == 0) {
      pairs[2 * (cn - 1)] = (int16_t)(ci - 1);
      pairs[2 * (cn - 1) + 1] = (int16_t)(cj - 1);
      cn--;
      ci--;
      cj--;
    } else if (c == 1)
      ci--;
    else
      cj--;
  }
}

/* lib/pair_alignment/pair_alignment.cc:714-820: gather from the ORIGINAL coordinates */
static float calc_min_rmsd(orc_ws *w, const float *A, const float *B, const int16_t *pairs,
                           int n, float T[12])
{
  int nb = w->p.nb, k, q, c, a = 0;
  for (k = 0; k < n; k++)
    for (q = 0; q < nb; q++, a++)
      for (c = 0; c < 3; c++) {
        w->gf[3 * a + c] = A[3 * (pairs[2 * k] * nb + q) + c];
        w->gm[3 * a + c] = B[3 * (pairs[2 * k + 1] * nb + q) + c];
      }
  return orc_superimpose(n * nb, w->gf, w->gm, T);
}

/* bin/minrms/search_or_dyn3d.cc:463-473 */
static int max_rmsd_exceeded(const orc_params *p, float ssd, int n)
{
  if (p->max_rmsd < 0.0f) return 0;
  return ssd > (float)(n * p->nb) * (p->max_rmsd * p->max_rmsd);
}

/* The n-loop shared by FirstPass (search_or_dyn3d.cc:1703-1835) and Refine
 * (search_or_dyn3d.cc:2411-2617); Bp must already hold the rotated molecule 2. */
static void eval_rotated(orc_ws *w, const float *A, const float *B, const float *T_orig,
                         orc_eval *out)
{
  const orc_params *p = &w->p;
  int n;
  memset(out->valid, 0, (size_t)w->N);
  fill_lookup(w, A);
  out->cost[0] = init_layer1(w);
  out->n_last = 1;
  for (n = 2; n <= p->n_max; n++) {
    float c = solve_layer(w, n);
    out->cost[n - 1] = c;
    out->n_last = n;
    if (c >= ORC_UPPER_BOUND) break; /* dyn3d.h:407-408, search_or_dyn3d.cc:1720 */
    if (n >= p->n_min) {
      int16_t *pairs = out->pairs + (size_t)(n - 1) * 2 * w->N;
      float *T = out->T + 12 * (n - 1);
      float ssd;
      export_solution(w, n, pairs);
      if (p->refine_while_searching) { /* search_or_dyn3d.cc:576-589 */
        float rmsd = calc_min_rmsd(w, A, B, pairs, n, T);
        ssd = ((rmsd * rmsd) * (float)n) * (float)p->nb;
      } else {
        ssd = c;
        memcpy(T, T_orig, 12 * sizeof(float));
      }
      if (max_rmsd_exceeded(p, ssd, n)) break; /* search_or_dyn3d.cc:1810-1811 */
      out->ssd[n - 1] = ssd;
      out->valid[n - 1] = 1;
    }
  }
}

void orc_eval_orientation(orc_ws *w, const float *A, const float *B, const float T12[12],
                          int apply_filter, orc_eval *out)
{
  const orc_params *p = &w->p;
  orc_apply_transform(T12, B, w->nB * p->nb, w->Bp); /* search_or_dyn3d.cc:1691-1693 */
  out->nw_matches = -1;
  if (apply_filter && p->use_filter) {
    out->nw_matches = orc_nw_num_matches(A, w->m, w->Bp, w->nB, p->nb, p->of_dist, p->of_min);
    if (out->nw_matches < p->of_min) { /* orientation_filter.cc:129 */
      out->n_last = 0;
      memset(out->valid, 0, (size_t)w->N);
      return;
    }
  }
  eval_rotated(w, A, B, T12, out);
}

void orc_best_init(const orc_params *p, int N, orc_best *best)
{
  int R = p->n_max - p->n_min + 1, r, k;
  (void)N;
  for (r = 0; r < R; r++) { /* search_or_dyn3d.cc:165-184 */
    best->ssd[r] = ORC_UPPER_BOUND;
    for (k = 0; k < 12; k++) best->T[12 * r + k] = (k % 5 == 0) ? 1.0f : 0.0f;
    best->o_id[r] = -1;
  }
}

static void record(orc_ws *w, orc_best *best, const orc_eval *ev, int n, int64_t o_id)
{
  int r = n - w->p.n_min; /* search_or_dyn3d.cc:597-606 */
  best->ssd[r] = ev->ssd[n - 1];
  memcpy(best->T + 12 * r, ev->T + 12 * (n - 1), 12 * sizeof(float));
  memcpy(best->pairs + (size_t)r * 2 * w->N, ev->pairs + (size_t)(n - 1) * 2 * w->N,
         (size_t)2 * n * sizeof(int16_t));
  best->o_id[r] = o_id;
}

int64_t orc_search(orc_ws *w, const float *A, const float *B, const float *orient12,
                   int64_t count, int64_t first_id, orc_best *best)
{
  int64_t o, cells = 0;
  int n;
  for (o = 0; o < count; o++) {
    orc_eval *ev = &w->ev;
    orc_eval_orientation(w, A, B, orient12 + 12 * o, 1, ev);
    for (n = 1; n <= ev->n_last; n++)
      cells += (int64_t)(layer_imax(w, n) - n + 1) * (int64_t)(layer_jmax(w, n) - n + 1);
    for (n = MAXI(2, w->p.n_min); n <= ev->n_last; n++)
      if (ev->valid[n - 1] && ev->ssd[n - 1] < best->ssd[n - w->p.n_min]) /* :1825 strict < */
        record(w, best, ev, n, first_id + o);
  }
  return cells;
}

/* lib/global_utils/simple_numeric_utils.h:350-376 */
static int find_median(const unsigned char *in_set, int n_min, int n_max)
{
  int n, num = 0, half, count = 0;
  for (n = n_min; n <= n_max; n++)
    if (in_set[n - n_min]) num++;
  if (num == 0) return n_min - 1;
  half = (num + 1) / 2;
  n = n_min - 1;
  do {
    ++n;
    if (in_set[n - n_min]) ++count;
  } while (count < half);
  return n;
}

/* bin/minrms/search_or_dyn3d.cc:2261-2658 (r = 0 only) */
int orc_refine(orc_ws *w, const float *A, const float *B, orc_best *best, int max_iters)
{
  const orc_params *p = &w->p;
  int R = p->n_max - p->n_min + 1, n, m2, iters, evals = 0;
  unsigned char *need = (unsigned char *)calloc((size_t)R, 1);
  int16_t *old = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)w->N);
  for (n = p->n_min; n <= p->n_max; n++)
    need[n - p->n_min] = best->ssd[n - p->n_min] < ORC_UPPER_BOUND;
  for (iters = 1; iters <= max_iters || max_iters < 0; iters++) {
    float T[12];
    orc_eval *ev = &w->ev;
    int16_t *cur;
    n = find_median(need, p->n_min, p->n_max);
    if (n == p->n_min - 1) break;
    cur = best->pairs + (size_t)(n - p->n_min) * 2 * w->N;
    calc_min_rmsd(w, A, B, cur, n, T); /* :2401-2407 */
    orc_apply_transform(T, B, w->nB * p->nb, w->Bp);
    memcpy(old, cur, sizeof(int16_t) * 2 * (size_t)n);
    ev->nw_matches = -1;
    eval_rotated(w, A, B, T, ev);
    evals++;
    for (m2 = MAXI(2, p->n_min); m2 <= ev->n_last; m2++)
      if (ev->valid[m2 - 1] && ev->ssd[m2 - 1] < best->ssd[m2 - p->n_min]) {
        record(w, best, ev, m2, -2);
        if (m2 != n) need[m2 - p->n_min] = 1; /* :2546 */
      }
    need[n - p->n_min] = memcmp(old, cur, sizeof(int16_t) * 2 * (size_t)n) != 0; /* :2619-2629 */
  }
  free(need);
  free(old);
  return evals;
}

/* ------------------------------------------------------------------------------------
 * Per-size statistics of align_chimera.plot.
 * bin/minrms/search_or_dyn3d.cc:2703-2760 (WritePlotFile) */

/* lib/pair_alignment/pair_alignment.cc:84-105 (SumSqdDistBetweenResidues) with
 * lib/vect_3d/vect_3d.h:268-276: every atom's dot product is summed left to right and
 * accumulated into a float that starts at 0. */
static float sum_sqd_dist_residues(const float *A, int i, const float *Bp, int j, int nb)
{
  float out = 0.0f;
  int q;
  for (q = 0; q < nb; q++) {
    const float *a = A + 3 * ((size_t)i * nb + q), *b = Bp + 3 * ((size_t)j * nb + q);
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    out += dx * dx + dy * dy + dz * dz;
  }
  return out;
}

/* lib/pair_alignment/pair_alignment.cc:2351-2379 (FindFurthestResPair) */
static float furthest_res_pair(const float *A, const float *Bp, int nb, const int16_t *pairs, int n)
{
  float furthest = 0.0f;
  int k;
  for (k = 0; k < n; k++) {
    float c = sum_sqd_dist_residues(A, pairs[2 * k], Bp, pairs[2 * k + 1], nb);
    if (c > furthest || k == 0) furthest = c;
  }
  return sqrtf(furthest);
}

/* lib/pair_alignment/pair_alignment.cc:2690-2790 (LevittGerstein98::Sstr_Score):
 * M = 20, d_o = 5, gap opening 10, gap extension 0; C-alpha atoms only. */
static float sstr_score(const float *A, const float *Bp, int nb, int ca, const int16_t *pairs, int n)
{
  const float M = 20.0f, d_o_sqrd = 5.0f * 5.0f;
  float total = 0.0f;
  int prev_i = -1, prev_j = -1, k;
  for (k = 0; k < n; k++) {
    const int i = pairs[2 * k], j = pairs[2 * k + 1];
    const float *a = A + 3 * ((size_t)i * nb + ca), *b = Bp + 3 * ((size_t)j * nb + ca);
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float d2 = dx * dx + dy * dy + dz * dz;
    int gap;
    total += M / (1 + (d2 / d_o_sqrd));
    gap = (i - prev_i - 1) + (j - prev_j - 1);
    if (gap > 0 && prev_i != -1) total -= (10.0f + (gap - 1) * 0.0f);
    prev_i = i;
    prev_j = j;
  }
  return total;
}

/* lib/pair_alignment/pair_alignment.cc:2798-2925 (LevittGerstein98::Sstr_Probability):
 * long double arithmetic on constants that are double literals; log() of a float and of
 * doubles resolve to the float / double overloads, exp() of a long double to expl. */
void orc_sstr_probability(float S_str_f, int n, float out4[4] /* mu, delta, Z, probability */)
{
  const long double c = 18.411424, d = -4.501719, e = 2.637163, f = 21.351857, g = -37.521269;
  const long double ln120 = log(120.0);
  const long double a = 2.0 * ln120 * c + d;
  const long double b = ln120 * ln120 * c + ln120 * d + e - a * ln120;
  long double mu, delta, Z, result;
  const long double lnN = logf((float)n);
  const long double S = S_str_f;
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
  Z = (S - mu) / delta;
  if (Z < log(10000.0 / 2))
    result = 1.0f - expl(-expl(-Z));
  else
    result = expl(-Z);
  out4[0] = (float)mu;
  out4[1] = (float)delta;
  out4[2] = (float)Z;
  out4[3] = (float)result;
}

/* One row of the plot per size of a best table: out[9 r ..] = {has solution, rmsd,
 * furthest distance, S_str, mu_str, delta_str, Z, probability, -log10(probability)} as
 * WritePlotFile evaluates them (search_or_dyn3d.cc:2703-2760).  ca = index of the " CA"
 * atom among the nb atoms used per residue. */
void orc_plot_stats(const float *A, int m, const float *B, int nB, int nb, int ca, int n_min,
                    int R, int N, const orc_best *best, double *out)
{
  float *Bp = (float *)malloc(sizeof(float) * 3 * (size_t)nB * nb);
  int r;
  (void)m;
  for (r = 0; r < R; r++) {
    const int n = n_min + r;
    double *o = out + 9 * (size_t)r;
    const int16_t *pairs = best->pairs + (size_t)r * 2 * N;
    float s4[4], S;
    memset(o, 0, 9 * sizeof(double));
    if (!(best->ssd[r] < ORC_UPPER_BOUND)) continue;
    o[0] = 1.0;
    o[1] = sqrtf(best->ssd[r] / (n * nb));
    orc_apply_transform(best->T + 12 * (size_t)r, B, nB * nb, Bp);
    o[2] = furthest_res_pair(A, Bp, nb, pairs, n);
    S = sstr_score(A, Bp, nb, ca, pairs, n);
    orc_sstr_probability(S, n, s4);
    o[3] = S; o[4] = s4[0]; o[5] = s4[1]; o[6] = s4[2]; o[7] = s4[3];
    o[8] = (s4[3] == 0.0) ? 0.0 : -log10f(s4[3]);
  }
  free(Bp);
}
