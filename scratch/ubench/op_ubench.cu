// Reciprocal throughput of single SASS ops on sm_100a (cycles per warp instruction per SM sub-partition).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o op_ubench op_ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define NCH 8
template <int OP>
__global__ void __launch_bounds__(1024, 1) k_op(int iters, const float* __restrict__ seed, float* __restrict__ out, long long* clk) {
  float a[NCH], b[NCH], c[NCH];
  unsigned u[NCH];
  unsigned long long q[NCH], q2[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) { a[i] = seed[threadIdx.x + i]; b[i] = seed[threadIdx.x + 64 + i]; c[i] = seed[threadIdx.x + 128 + i]; u[i] = __float_as_uint(seed[threadIdx.x + 192 + i]); q[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | u[i]; q2[i] = ((unsigned long long)__float_as_uint(b[i]) << 32) | __float_as_uint(c[i]); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        if (OP == 1) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
        if (OP == 2) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        if (OP == 3) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(u[i]) : "r"(__float_as_uint(b[i])));
        if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(q2[i]));
        if (OP == 5) asm volatile("mul.rn.sat.f32 %0, %0, 0fFF000000;" : "+f"(a[i]));
        if (OP == 6) asm volatile("fma.rn.f32 %0, %0, 0f40000000, %1;" : "+f"(a[i]) : "f"(b[i]));
        if (OP == 7) asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(u[i]) : "r"(__float_as_uint(b[i])));
        if (OP == 8) asm volatile("mad.hi.u32 %0, %0, 2, %1;" : "+r"(u[i]) : "r"(__float_as_uint(b[i])));
        if (OP == 9) asm volatile("add.rn.f32 %0, %0, 0f3FC00000;" : "+f"(a[i]));
        if (OP == 10) { u[i] = (unsigned)__vimin3_s32((int)u[i], __float_as_int(b[i]) + rep, __float_as_int(c[i])); }
        if (OP == 11) asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(__float_as_uint(b[i])));
        if (OP == 12) asm volatile("lop3.b32 %0, %0, %1, 0x80000000, 0xE4;" : "+r"(u[i]) : "r"(__float_as_uint(b[i])));
        if (OP == 13) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(q2[i]), "l"(q2[(i + 1) % NCH]));
        if (OP == 14) asm volatile("add.rn.f32 %0, %0, %1; min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
        if (OP == 15) asm volatile("add.rn.f32 %0, %0, %1; shf.l.wrap.b32 %2, %0, %2, 1;" : "+f"(a[i]), "+r"(u[i]) : "f"(b[i]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += a[i] + c[i] + __uint_as_float(u[i]) + __uint_as_float((unsigned)(q[i] >> 32)) + __uint_as_float((unsigned)q[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP>
static void run(const char* name, const float* d_seed, float* d_out, long long* d_clk) {
  const int iters = 2000;
  const int W[] = {4, 8, 12, 16, 32};
  printf("%-14s", name);
  for (int w : W) {
    k_op<OP><<<1, 32 * w>>>(iters, d_seed, d_out, d_clk);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, d_clk, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)iters * 4 * NCH * (w / 4.0));
    printf("  w%-2d %5.2f", w, per);
  }
  printf("   (clk per warp-instr per SMSP)  %s\n", cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* d_seed; float* d_out; long long* d_clk;
  cudaMalloc(&d_seed, 8192); cudaMalloc(&d_out, 1024 * 4); cudaMalloc(&d_clk, 8);
  float h[2048];
  for (int i = 0; i < 2048; i++) h[i] = (float)(rand() % 1000) * 0.37f + 0.1f;
  cudaMemcpy(d_seed, h, 8192, cudaMemcpyHostToDevice);
  run<0>("FADD", d_seed, d_out, d_clk);
  run<9>("FADD imm", d_seed, d_out, d_clk);
  run<1>("FMNMX3", d_seed, d_out, d_clk);
  run<2>("FMNMX", d_seed, d_out, d_clk);
  run<3>("SHF", d_seed, d_out, d_clk);
  run<4>("FADD2", d_seed, d_out, d_clk);
  run<13>("FFMA2", d_seed, d_out, d_clk);
  run<5>("FMUL.SAT imm", d_seed, d_out, d_clk);
  run<6>("FFMA imm", d_seed, d_out, d_clk);
  run<7>("IMAD", d_seed, d_out, d_clk);
  run<8>("IMAD.HI", d_seed, d_out, d_clk);
  run<10>("VIMNMX3", d_seed, d_out, d_clk);
  run<11>("VIMNMX", d_seed, d_out, d_clk);
  run<12>("LOP3", d_seed, d_out, d_clk);
  run<14>("FADD+FMNMX3 (x2)", d_seed, d_out, d_clk);
  run<15>("FADD+SHF (x2)", d_seed, d_out, d_clk);
  return 0;
}
