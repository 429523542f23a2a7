// mr_device.cuh -- device-side arithmetic shared by every kernel.
//
// All float arithmetic is IEEE-754 binary32, round-to-nearest, in the reference's
// source order; the translation unit is compiled with -fmad=false (no FMA
// contraction), default -prec-div=true -prec-sqrt=true and no -ftz, so results are
// bit-identical to the reference built without -ffast-math.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <mutex>
#include <utility>

// Opt-in dynamic shared memory of a kernel (cudaFuncAttributeMaxDynamicSharedMemorySize), raised
// monotonically and recorded per (kernel, device).  Function attributes belong to the process:
// contexts driven by several host threads (all-vs-all workers, one thread per GPU) must never
// lower the limit another thread is about to launch with.  Returns 0 on success.
template <class Kern>
static inline int mr_ensure_smem(Kern kern, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;  // within the default limit
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> have;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  std::lock_guard<std::mutex> lk(mu);
  size_t& h = have[std::make_pair(reinterpret_cast<const void*>(kern), dev)];
  if (h >= bytes) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return 1;
  h = bytes;
  return 0;
}

// -DMR_CHECK: device-side bounds checks on every table / scratch / carry index the kernels
// form (printf + trap); a checking build of the library is run over scratch/sanitize_case.py and
// the parity tests (profiles/r02_safety.md) -- compute-sanitizer is closed on this GPU pool.
#ifdef MR_CHECK
#include <cstdio>
#define MR_ASSERT(cond)                                                                                   \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      printf("MR_ASSERT failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, blockIdx.x, \
             threadIdx.x);                                                                                \
      __trap();                                                                                           \
    }                                                                                                     \
  } while (0)
#else
#define MR_ASSERT(cond) ((void)0)
#endif

#define MR_UB 1.0e26f  // DistanceMetric::UPPER_BOUND (lib/dyn/distance_metric.cc:49)

struct MrDevParams {
  int m, nB, nb, N;          // residues of A / B, atoms per residue, min(m, nB)
  int n_min, n_max;
  int wA, wB;                // Dyn3d window sizes (lib/dyn/dyn3d.cc:472-473)
  float dp_cutoff;           // metric cutoff of the 3-D DP (nb * max_dist^2 or 1e26)
  float max_rmsd;            // < 0 = none
  int use_filter;
  float nw_cutoff;           // nb * of_dist^2
  float nw_gap;              // G (lib/or_gen/orientation_filter.cc:77-79)
  int of_min;
  int w0, w1;                // DynNW window sizes (lib/dyn/dynNW.cc:107-108)
  int refine_while_searching;
};

// ---------------------------------------------------------------------------
// 4x4 cyclic Jacobi eigen-solver (classical Jacobi with the threshold strategy of
// Rutishauser; same operation order and float/double mix as the solver the reference
// links, lib/eigen/jacobi.c:7-85, so eigenvalues/vectors agree to the last bit).
// Fully unrolled over (ip,iq) so a, v, d, b, z live in registers.
// ---------------------------------------------------------------------------
#define MR_ROT(M, i1, j1, k1, l1)             \
  {                                           \
    float gg = M[i1][j1], hh = M[k1][l1];     \
    M[i1][j1] = gg - s * (hh + gg * tau);     \
    M[k1][l1] = hh + s * (gg - hh * tau);     \
  }

__device__ __forceinline__ void mr_jacobi4(float (&a)[4][4], float (&d)[4], float (&v)[4][4]) {
  float b[4], z[4];
#pragma unroll
  for (int ip = 0; ip < 4; ip++) {
#pragma unroll
    for (int iq = 0; iq < 4; iq++) v[ip][iq] = (ip == iq) ? 1.0f : 0.0f;
    b[ip] = d[ip] = a[ip][ip];
    z[ip] = 0.0f;
  }
  for (int sweep = 1; sweep <= 50; sweep++) {
    float sm = 0.0f;
#pragma unroll
    for (int ip = 0; ip < 3; ip++)
#pragma unroll
      for (int iq = ip + 1; iq < 4; iq++) sm = (float)((double)sm + fabs((double)a[ip][iq]));
    if (sm == 0.0f) return;
    float tresh = (sweep < 4) ? (float)(0.2 * (double)sm / 16.0) : 0.0f;
#pragma unroll
    for (int ip = 0; ip < 3; ip++) {
#pragma unroll
      for (int iq = ip + 1; iq < 4; iq++) {
        float apq = a[ip][iq];
        float g = (float)(100.0 * fabs((double)apq));
        float adp = fabsf(d[ip]), adq = fabsf(d[iq]);
        if (sweep > 4 && (float)((double)adp + (double)g) == adp &&
            (float)((double)adq + (double)g) == adq) {
          a[ip][iq] = 0.0f;
        } else if (fabsf(apq) > tresh) {
          float h = d[iq] - d[ip];
          float ah = fabsf(h);
          float t;
          if ((float)((double)ah + (double)g) == ah) {
            t = apq / h;
          } else {
            float theta = (float)(0.5 * (double)h / (double)apq);
            float th2 = theta * theta;
            t = (float)(1.0 / (fabs((double)theta) + sqrt(1.0 + (double)th2)));
            if (theta < 0.0f) t = -t;
          }
          float one_tt = 1.0f + t * t;
          float c = (float)(1.0 / sqrt((double)one_tt));
          float s = t * c;
          float tau = (float)((double)s / (1.0 + (double)c));
          h = t * apq;
          z[ip] -= h;
          z[iq] += h;
          d[ip] -= h;
          d[iq] += h;
          a[ip][iq] = 0.0f;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (j < ip) MR_ROT(a, j, ip, j, iq)
            else if (j > ip && j < iq) MR_ROT(a, ip, j, j, iq)
            else if (j > iq) MR_ROT(a, ip, j, iq, j)
          }
#pragma unroll
          for (int j = 0; j < 4; j++) MR_ROT(v, j, ip, j, iq)
        }
      }
    }
#pragma unroll
    for (int ip = 0; ip < 4; ip++) {
      b[ip] += z[ip];
      d[ip] = b[ip];
      z[ip] = 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------
// Diamond (1988) quaternion least-squares superposition of np point pairs
// (Superimpose::FindTransformMinRMSD, lib/superimpose/superimpose.cc:17-287;
// FindCentroid / Normalize4 / ConvertQuaternionToMatrix3x3 lib/vect_3d/vect_3d.cc:61-117).
// Fetch(a, f, mv) must load the a-th fixed and moveable points, a = 0..np-1; sums are
// accumulated sequentially in that order (the reference's order).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float mr_superimpose_finish(int np, float cf0, float cf1, float cf2, float cm0, float cm1,
                                                       float cm2, const float (&M)[3][3], float E0, float (&T)[12]);

template <class Fetch>
__device__ __forceinline__ float mr_superimpose(int np, Fetch fetch, float (&T)[12]) {
  float cf0 = 0.f, cf1 = 0.f, cf2 = 0.f, cm0 = 0.f, cm1 = 0.f, cm2 = 0.f;
  for (int a = 0; a < np; a++) {
    float f[3], mv[3];
    fetch(a, f, mv);
    cf0 += f[0]; cf1 += f[1]; cf2 += f[2];
    cm0 += mv[0]; cm1 += mv[1]; cm2 += mv[2];
  }
  const float inv_np = 1.0f / (float)np;
  cf0 *= inv_np; cf1 *= inv_np; cf2 *= inv_np;
  cm0 *= inv_np; cm1 *= inv_np; cm2 *= inv_np;
  float M[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float E0 = 0.f;
  for (int a = 0; a < np; a++) {
    float f[3], mv[3];
    fetch(a, f, mv);
    f[0] -= cf0; f[1] -= cf1; f[2] -= cf2;
    mv[0] -= cm0; mv[1] -= cm1; mv[2] -= cm2;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) M[i][j] += mv[i] * f[j];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float df = f[i] - mv[i];
      E0 += df * df;
    }
  }
  return mr_superimpose_finish(np, cf0, cf1, cf2, cm0, cm1, cm2, M, E0, T);
}

// second half of mr_superimpose: from the centroids, the 3x3 covariance sums and E0 to the
// transform and the RMSD (superimpose.cc:120-287)
__device__ __forceinline__ float mr_superimpose_finish(int np, float cf0, float cf1, float cf2, float cm0, float cm1,
                                                       float cm2, const float (&M)[3][3], float E0, float (&T)[12]) {
  const float tr = M[0][0] + M[1][1] + M[2][2];
  float P[4][4], d[4], ev[4][4];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) P[i][j] = M[i][j] + M[j][i] - ((i == j) ? (2.0f * tr) : 0.0f);
  P[0][3] = P[3][0] = M[1][2] - M[2][1];
  P[1][3] = P[3][1] = M[2][0] - M[0][2];
  P[2][3] = P[3][2] = M[0][1] - M[1][0];
  P[3][3] = 0.0f;
  mr_jacobi4(P, d, ev);
  float lam = d[0];
  float p0 = ev[0][0], p1 = ev[1][0], p2 = ev[2][0], p3 = ev[3][0];
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (d[i] > lam) {
      lam = d[i];
      p0 = ev[0][i]; p1 = ev[1][i]; p2 = ev[2][i]; p3 = ev[3][i];
    }
  const float len = sqrtf(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
  const float inv = (float)(1.0 / (double)len);
  p0 *= inv; p1 *= inv; p2 *= inv; p3 *= inv;
  float R[3][3];
  R[0][0] = p0 * p0 - p1 * p1 - p2 * p2 + p3 * p3;
  R[1][1] = -(p0 * p0) + p1 * p1 - p2 * p2 + p3 * p3;
  R[2][2] = -(p0 * p0) - p1 * p1 + p2 * p2 + p3 * p3;
  R[0][1] = 2.0f * (p0 * p1 - p2 * p3);
  R[1][0] = 2.0f * (p0 * p1 + p2 * p3);
  R[1][2] = 2.0f * (p1 * p2 - p0 * p3);
  R[2][1] = 2.0f * (p1 * p2 + p0 * p3);
  R[0][2] = 2.0f * (p0 * p2 + p1 * p3);
  R[2][0] = 2.0f * (p0 * p2 - p1 * p3);
  const float cf[3] = {cf0, cf1, cf2};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float t = R[i][0] * cm0 + R[i][1] * cm1 + R[i][2] * cm2;
    T[4 * i + 0] = R[i][0];
    T[4 * i + 1] = R[i][1];
    T[4 * i + 2] = R[i][2];
    T[4 * i + 3] = cf[i] - t;
  }
  float min_ssd = E0 - 2.0f * lam;
  if (min_ssd < 0.0f) min_ssd = 0.0f;
  return sqrtf(min_ssd / (float)np);
}

// B' = R.B + t  (Mult_Mat3x4_by_Vect3, lib/vect_3d/vect_3d.h:145-155)
__device__ __forceinline__ void mr_xform(const float* __restrict__ T, float x, float y, float z,
                                         float& ox, float& oy, float& oz) {
  ox = T[0] * x + T[1] * y + T[2] * z + T[3];
  oy = T[4] * x + T[5] * y + T[6] * z + T[7];
  oz = T[8] * x + T[9] * y + T[10] * z + T[11];
}

// one atom's contribution to the squared-distance metric (lib/vect_3d/vect_3d.h:268-276)
__device__ __forceinline__ float mr_d2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return dx * dx + dy * dy + dz * dz;
}
