// Micro-benchmark of the 3-D DP cell update (4 layers x 8 rows per thread, column sweep) in
// isolation: which instruction mix / occupancy the SM pipes actually sustain on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o cell_ubench cell_ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define K 4
#define R 8

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// VAR: 0 ping-pong scalar (3 FADD, FMNMX3, 2 SHF)     1 two-min in place (3 FADD, 2 FMNMX, 2 SHF)
//      2 ping-pong, FADD2 over layer pairs             3 = 2 + s2 through FMUL.SAT / FFMA (float accumulators)
//      4 costs only (FADD + FMNMX3)                    5 = 2 + s2 through LEA.HI-free IMAD pair
// MEM: 0 registers only   1 LDS/STS exchange as in the real kernel (no barrier)   2 = 1 + __syncthreads per step
template <int VAR, int MEM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cell(int steps, int nslots, const float* __restrict__ seed, uint32_t* __restrict__ out) {
  extern __shared__ float4 smem4[];
  const int tid = threadIdx.x;
  float* sm = reinterpret_cast<float*>(smem4);
  const int ring = 32 * 64;  // 32 columns x 64 rows of C
  for (int k = tid; k < ring + 24 * nslots; k += blockDim.x) sm[k] = seed[k & 1023];
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t a_c = base + (uint32_t)((tid & 7) * 8) * 4u;
  const uint32_t exch0 = base + ring * 4u;
  const uint32_t par = (uint32_t)nslots * 48u;
  const uint32_t a_me = exch0 + (uint32_t)tid * 48u;
  const uint32_t a_up = exch0 + (uint32_t)((tid + 1) % nslots) * 48u + 32u;
  const uint32_t a_dg = exch0 + (uint32_t)((tid + 37) % nslots) * 48u;
  float st[2][K][R], bprev[2][K];
  uint32_t acc[K];
  float accf[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    bprev[0][k] = bprev[1][k] = seed[tid + k];
    acc[k] = 0; accf[k] = 0.f;
#pragma unroll
    for (int r = 0; r < R; r++) st[0][k][r] = st[1][k][r] = seed[(tid * 32 + k * 8 + r) & 1023] + 50.0f;
  }
  uint32_t sink = 0;
  auto step = [&](auto src_tag, const int s, const uint32_t rd_off, const uint32_t wr_off) {
    constexpr int S = (VAR == 1) ? 0 : decltype(src_tag)::value, D = (VAR == 1) ? 0 : 1 - S;
    float cv[R], dsh[R], dexp[R];
    if (MEM) {
      const uint32_t ca = a_c + (uint32_t)(s & 31) * 256u;
      const float4 v0 = lds128(ca), v1 = lds128(ca + 16u);
      cv[0] = v0.x; cv[1] = v0.y; cv[2] = v0.z; cv[3] = v0.w; cv[4] = v1.x; cv[5] = v1.y; cv[6] = v1.z; cv[7] = v1.w;
      const float4 v = lds128(a_up + rd_off);
      bprev[D][0] = v.x; bprev[D][1] = v.y; bprev[D][2] = v.z; bprev[D][3] = v.w;
      const float4 a4 = lds128(a_dg + rd_off), b4 = lds128(a_dg + rd_off + 16u);
      dsh[0] = a4.x; dsh[1] = a4.y; dsh[2] = a4.z; dsh[3] = a4.w; dsh[4] = b4.x; dsh[5] = b4.y; dsh[6] = b4.z; dsh[7] = b4.w;
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) { cv[r] = st[S][0][r] * 0.5f; dsh[r] = st[S][K - 1][r]; }
#pragma unroll
      for (int k = 0; k < K; k++) bprev[D][k] = st[S][k][R - 1];
    }
    if constexpr (VAR == 0 || VAR == 1 || VAR == 4) {
#pragma unroll
      for (int k = K - 1; k >= 0; k--) {
        float up = bprev[D][k];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float dg = (k > 0) ? ((r > 0) ? st[S][k - 1][r - 1] : bprev[S][k - 1]) : dsh[r];
          const float d = dg + cv[r];
          const float left = st[S][k][r];
          float best;
          if (VAR == 1) {
            const float t = fminf(d, up);
            const float s1 = __fsub_rn(d, up), s2 = __fsub_rn(t, left);
            best = fminf(t, left);
            acc[k] = __funnelshift_l(__float_as_uint(s1), acc[k], 1);
            acc[k] = __funnelshift_l(__float_as_uint(s2), acc[k], 1);
          } else {
            best = fminf(fminf(d, up), left);
            if (VAR == 0) {
              const float s1 = __fsub_rn(d, up), s2 = __fsub_rn(best, left);
              acc[k] = __funnelshift_l(__float_as_uint(s1), acc[k], 1);
              acc[k] = __funnelshift_l(__float_as_uint(s2), acc[k], 1);
            }
          }
          st[D][k][r] = best;
          up = best;
        }
      }



    } else if constexpr (VAR == 9) {
      // hand-ordered (volatile asm keeps the order): per row, the 4 layers side by side so that the
      // shared operands sit in the same slot of consecutive instructions (operand reuse cache)
      float up[K];
#pragma unroll
      for (int k = 0; k < K; k++) up[k] = bprev[D][k];
#pragma unroll
      for (int r = 0; r < R; r++) {
        float d[K], best[K], z[K], s2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
          const float dg = (k > 0) ? ((r > 0) ? st[S][k - 1][r - 1] : bprev[S][k - 1]) : dsh[r];
          asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d[k]) : "f"(dg), "f"(cv[r]));
        }
#pragma unroll
        for (int k = 0; k < K; k++)
          asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(best[k]) : "f"(d[k]), "f"(up[k]), "f"(st[S][k][r]));
#pragma unroll
        for (int k = 0; k < K; k++) {
          asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(z[k]) : "f"(best[k]), "f"(up[k]));
          asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(s2[k]) : "f"(best[k]), "f"(st[S][k][r]));
        }
#pragma unroll
        for (int k = 0; k < K; k++) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(acc[k]) : "r"(__float_as_uint(z[k])));
#pragma unroll
        for (int k = 0; k < K; k++) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(acc[k]) : "r"(__float_as_uint(s2[k])));
#pragma unroll
        for (int k = 0; k < K; k++) { st[D][k][r] = best[k]; up[k] = best[k]; }
      }
    } else if constexpr (VAR == 8) {
      // row-major order: the 4 layers of a row side by side (independent chains, shared c operand)
      float up[K];
#pragma unroll
      for (int k = 0; k < K; k++) up[k] = bprev[D][k];
#pragma unroll
      for (int r = 0; r < R; r++) {
        float d[K], best[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
          const float dg = (k > 0) ? ((r > 0) ? st[S][k - 1][r - 1] : bprev[S][k - 1]) : dsh[r];
          d[k] = dg + cv[r];
        }
#pragma unroll
        for (int k = 0; k < K; k++) best[k] = fminf(fminf(d[k], up[k]), st[S][k][r]);
#pragma unroll
        for (int k = 0; k < K; k++) {
          const float z = __fsub_rn(best[k], up[k]), s2 = __fsub_rn(best[k], st[S][k][r]);
          acc[k] = __funnelshift_l(__float_as_uint(z), acc[k], 1);
          acc[k] = __funnelshift_l(__float_as_uint(s2), acc[k], 1);
          st[D][k][r] = best[k];
          up[k] = best[k];
        }
      }
    } else if constexpr (VAR == 6 || VAR == 7) {
      // aligned layer pairs: left/best pairs (0,1),(2,3); d pairs (1,2),(3,4) computed from the left
      // pairs of row r-1; predicates z = best - up (pairs) and s2 = best - left (pairs): no copies.
      // d(0,r) arrives pre-added from the lower chunk (dsh); d(4,r) is exported instead of the top layer.
      float2 upA = make_float2(bprev[D][0], bprev[D][1]), upB = make_float2(bprev[D][2], bprev[D][3]);
      float2 dgA = make_float2(bprev[S][0], bprev[S][1]), dgB = make_float2(bprev[S][2], bprev[S][3]);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float2 c2 = make_float2(cv[r], cv[r]);
        const float2 d12 = __fadd2_rn(dgA, c2), d34 = __fadd2_rn(dgB, c2);
        const float2 lA = make_float2(st[S][0][r], st[S][1][r]), lB = make_float2(st[S][2][r], st[S][3][r]);
        float2 bA, bB;
        bA.x = fminf(fminf(dsh[r], upA.x), lA.x);
        bA.y = fminf(fminf(d12.x, upA.y), lA.y);
        bB.x = fminf(fminf(d12.y, upB.x), lB.x);
        bB.y = fminf(fminf(d34.x, upB.y), lB.y);
        dexp[r] = d34.y;
        const float2 zA = __fadd2_rn(bA, make_float2(-upA.x, -upA.y)), zB = __fadd2_rn(bB, make_float2(-upB.x, -upB.y));
        const float2 sA = __fadd2_rn(bA, make_float2(-lA.x, -lA.y)), sB = __fadd2_rn(bB, make_float2(-lB.x, -lB.y));
        acc[0] = __funnelshift_l(__float_as_uint(zA.x), acc[0], 1);
        acc[1] = __funnelshift_l(__float_as_uint(zA.y), acc[1], 1);
        acc[2] = __funnelshift_l(__float_as_uint(zB.x), acc[2], 1);
        acc[3] = __funnelshift_l(__float_as_uint(zB.y), acc[3], 1);
        if (VAR == 6) {
          acc[0] = __funnelshift_l(__float_as_uint(sA.x), acc[0], 1);
          acc[1] = __funnelshift_l(__float_as_uint(sA.y), acc[1], 1);
          acc[2] = __funnelshift_l(__float_as_uint(sB.x), acc[2], 1);
          acc[3] = __funnelshift_l(__float_as_uint(sB.y), acc[3], 1);
        } else {
          const float2 fA = make_float2(__saturatef(sA.x * -1.7014118e38f), __saturatef(sA.y * -1.7014118e38f));
          const float2 fB = make_float2(__saturatef(sB.x * -1.7014118e38f), __saturatef(sB.y * -1.7014118e38f));
          const float2 aA = __ffma2_rn(make_float2(accf[0], accf[1]), make_float2(2.f, 2.f), fA);
          const float2 aB = __ffma2_rn(make_float2(accf[2], accf[3]), make_float2(2.f, 2.f), fB);
          accf[0] = aA.x; accf[1] = aA.y; accf[2] = aB.x; accf[3] = aB.y;
        }
        dgA = lA; dgB = lB;
        st[D][0][r] = bA.x; st[D][1][r] = bA.y; st[D][2][r] = bB.x; st[D][3][r] = bB.y;
        upA = bA; upB = bB;
      }
    } else {
      // layer pairs (k, k+1) as float2 lanes: FADD2 for d, s1, s2
#pragma unroll
      for (int kp = K / 2 - 1; kp >= 0; kp--) {
        const int k0 = 2 * kp, k1 = 2 * kp + 1;
        float2 up = make_float2(bprev[D][k0], bprev[D][k1]);
#pragma unroll
        for (int r = 0; r < R; r++) {
          float2 dg;
          dg.x = (k0 > 0) ? ((r > 0) ? st[S][k0 - 1][r - 1] : bprev[S][k0 - 1]) : dsh[r];
          dg.y = (r > 0) ? st[S][k1 - 1][r - 1] : bprev[S][k1 - 1];
          const float2 c2 = make_float2(cv[r], cv[r]);
          const float2 d = __fadd2_rn(dg, c2);
          const float2 left = make_float2(st[S][k0][r], st[S][k1][r]);
          float2 best;
          best.x = fminf(fminf(d.x, up.x), left.x);
          best.y = fminf(fminf(d.y, up.y), left.y);
          const float2 s1 = __fadd2_rn(d, make_float2(-up.x, -up.y));
          const float2 s2 = __fadd2_rn(best, make_float2(-left.x, -left.y));
          acc[k0] = __funnelshift_l(__float_as_uint(s1.x), acc[k0], 1);
          acc[k1] = __funnelshift_l(__float_as_uint(s1.y), acc[k1], 1);
          if (VAR == 2) {
            acc[k0] = __funnelshift_l(__float_as_uint(s2.x), acc[k0], 1);
            acc[k1] = __funnelshift_l(__float_as_uint(s2.y), acc[k1], 1);
          } else if (VAR == 3) {
            const float b0 = __saturatef(s2.x * -1.7014118e38f), b1 = __saturatef(s2.y * -1.7014118e38f);
            const float2 a2 = __ffma2_rn(make_float2(accf[k0], accf[k1]), make_float2(2.f, 2.f), make_float2(b0, b1));
            accf[k0] = a2.x; accf[k1] = a2.y;
          } else {
            accf[k0] = __uint_as_float(__umulhi(__float_as_uint(s2.x), 2u) + 2u * __float_as_uint(accf[k0]));
            accf[k1] = __uint_as_float(__umulhi(__float_as_uint(s2.y), 2u) + 2u * __float_as_uint(accf[k1]));
          }
          st[D][k0][r] = best.x; st[D][k1][r] = best.y;
          up = best;
        }
      }
    }
    if (!MEM && (VAR == 6 || VAR == 7)) { sink ^= __float_as_uint(dexp[s & 7]); }
    if (MEM) {
      if (VAR == 6 || VAR == 7) {
        sts128(a_me + wr_off, dexp[0], dexp[1], dexp[2], dexp[3]);
        sts128(a_me + wr_off + 16u, dexp[4], dexp[5], dexp[6], dexp[7]);
      } else {
        sts128(a_me + wr_off, bprev[D][K - 1], st[D][K - 1][0], st[D][K - 1][1], st[D][K - 1][2]);
        sts128(a_me + wr_off + 16u, st[D][K - 1][3], st[D][K - 1][4], st[D][K - 1][5], st[D][K - 1][6]);
      }
      sts128(a_me + wr_off + 32u, st[D][0][R - 1], st[D][1][R - 1], st[D][2][R - 1], st[D][3][R - 1]);
    }
    if ((s & 2) == 0 && (s & 1)) {  // every second pair of columns: fold the words (stands for the store)
#pragma unroll
      for (int k = 0; k < K; k++) { sink ^= acc[k] + __float_as_uint(accf[k]); if (VAR == 3 || VAR == 7) accf[k] = 0.f; }
    }
    if (MEM == 2) __syncthreads();
  };
  for (int s = 0; s < steps; s += 2) {
    step(std::integral_constant<int, 0>(), s, 0u, par);
    step(std::integral_constant<int, 1>(), s + 1, par, 0u);
  }
  float fs = 0.f;
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int r = 0; r < R; r++) fs += st[0][k][r] + st[1][k][r];
  out[blockIdx.x * blockDim.x + tid] = sink ^ __float_as_uint(fs);
}

template <int VAR, int MEM, int MAXT, int MINB>
static void run(const char* name, int threads, int ctas_per_sm, int steps, const float* d_seed, uint32_t* d_out) {
  const int nslots = threads;
  const size_t smem = (32 * 64 + 24 * nslots) * 4;
  cudaFuncSetAttribute(k_cell<VAR, MEM, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cell<VAR, MEM, MAXT, MINB>, threads, smem);
  if (occ < ctas_per_sm) { printf("%-28s thr %4d x%d: occupancy %d, skipped\n", name, threads, ctas_per_sm, occ); return; }
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_cell<VAR, MEM, MAXT, MINB><<<grid, threads, smem>>>(64, nslots, d_seed, d_out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(e0);
    k_cell<VAR, MEM, MAXT, MINB><<<grid, threads, smem>>>(steps, nslots, d_seed, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  const double cells = (double)grid * threads * 32.0 * steps;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_cell<VAR, MEM, MAXT, MINB>);
  printf("%-28s thr %4d x%d regs %3d: %7.3f ms  %6.3f TCUPS  (%s)\n", name, threads, ctas_per_sm, fa.numRegs, best,
         cells / best * 1e-9, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 4000;
  std::vector<float> h(1024);
  srand(7);
  for (auto& v : h) v = (float)(rand() % 1000) * 0.37f + 0.1f;
  float* d_seed; uint32_t* d_out;
  cudaMalloc(&d_seed, 4096); cudaMalloc(&d_out, 148 * 2 * 1024 * 4);
  cudaMemcpy(d_seed, h.data(), 4096, cudaMemcpyHostToDevice);
  const int T[] = {128, 256, 352, 384};
  for (int t : T) {
#define RUNV(V, M, NAME) run<V, M, 384, 1>(NAME, t, 1, steps, d_seed, d_out);
    RUNV(4, 0, "costs-only reg")
    RUNV(0, 0, "pp scalar reg")
    RUNV(1, 0, "two-min reg")
    RUNV(2, 0, "pp fadd2 reg")
    RUNV(3, 0, "pp fadd2 + s2 fmul.sat reg")
    RUNV(5, 0, "pp fadd2 + s2 imad reg")
    RUNV(6, 0, "aligned pairs reg")
    RUNV(8, 0, "pp scalar z-form row-major reg")
    RUNV(9, 0, "hand-ordered asm reg")
    RUNV(7, 0, "aligned pairs + s2 fmul.sat reg")
    RUNV(4, 1, "costs-only smem")
    RUNV(0, 1, "pp scalar smem")
    RUNV(1, 1, "two-min smem")
    RUNV(2, 1, "pp fadd2 smem")
    RUNV(3, 1, "pp fadd2 + s2 fmul.sat smem")
    RUNV(6, 1, "aligned pairs smem")
    RUNV(8, 1, "pp scalar z-form row-major smem")
    RUNV(9, 1, "hand-ordered asm smem")
    RUNV(9, 2, "hand-ordered asm smem+bar")
    RUNV(8, 2, "pp scalar z-form row-major smem+bar")
    RUNV(7, 1, "aligned pairs + s2 fmul.sat smem")
    RUNV(0, 2, "pp scalar smem+bar")
    RUNV(6, 2, "aligned pairs smem+bar")
    RUNV(7, 2, "aligned pairs + s2 fmul.sat smem+bar")
    RUNV(2, 2, "pp fadd2 smem+bar")
    RUNV(3, 2, "pp fadd2 + s2 fmul.sat smem+bar")
  }
  // two CTAs per SM where registers allow
  run<1, 1, 352, 2>("two-min smem", 352, 2, steps, d_seed, d_out);
  run<1, 2, 352, 2>("two-min smem+bar", 352, 2, steps, d_seed, d_out);
  run<4, 1, 352, 2>("costs-only smem", 352, 2, steps, d_seed, d_out);
  return 0;
}
