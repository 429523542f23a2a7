// mr_dp_tiled.cuh -- register-tiled 3-D DP kernel (placeholder until the tiled kernel lands).
#pragma once
#include "mr_device.cuh"

struct TbTiled {
  int NP;
  __host__ __device__ size_t words(int) const { return 0; }
  __device__ __forceinline__ size_t word(int, int, int) const { return 0; }
  __device__ __forceinline__ int shift(int) const { return 0; }
};

struct MrTiledPlan {
  bool ok = false;
  const char* why = "not built";
  size_t tb_words = 0;
  TbTiled layout;
};

static inline void mr_plan_dp_tiled(MrTiledPlan* p, const MrDevParams&, const cudaDeviceProp&) { p->ok = false; }
static inline void mr_free_dp_tiled(MrTiledPlan*) {}
static inline int mr_launch_dp_tiled(const MrTiledPlan&, const MrDevParams&, const float*, const float*,
                                     const float*, const int*, const int*, uint32_t*, float*, int,
                                     cudaStream_t) { return 1; }
