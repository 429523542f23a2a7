// mr_dp_vp.cu -- the tiled 3-D DP kernel in its min-with-predicate form (mr_tile_update_vp,
// mr_tile_step.cuh), one extern "C" kernel per (CTA size, atoms per residue, windowed) bucket.
// Compiled to a cubin, patched by minrms_b200/sass_patch.py (VIMNMX.S16x2 -> VIMNMX.U32, the
// predicate destinations stay) and embedded in the library, which loads it through the driver
// API (mr_vp_module.h).  Nothing in here is ever launched unpatched.
#define MR_DP_TILED_DEVICE_ONLY
#include "mr_dp_tiled.cuh"

// launch bounds: registers are allocated per warp in units of 512, so two CTAs of 352 threads
// (11 warps) get 80 registers per thread (88 measured: one CTA per SM)
#define MR_VP_ATTR(MAXT, MINB) __launch_bounds__(MAXT, MINB)
#define MR_VP_KERNEL(NAME, MAXT, MINB, NB1, WINDOWED)                                                               \
  extern "C" __global__ void MR_VP_ATTR(MAXT, MINB)                                                            \
  NAME(MrDevParams P, TbTiled L, int nslots, int nsteps, const MrTile* __restrict__ tiles,                            \
       const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ orient,                    \
       const int* __restrict__ surv, const int* __restrict__ surv_count, uint32_t* __restrict__ tb,                   \
       float* __restrict__ cost_out) {                                                                                \
    mr_dp_tiles_body<4, 8, NB1, WINDOWED, false, false, true>(P, L, nslots, nsteps, tiles, A, B, orient, surv,        \
                                                              surv_count, tb, cost_out);                              \
  }
#define MR_VP_BUCKET(MAXT, MINB)                                   \
  MR_VP_KERNEL(k_dp_vp_##MAXT##_g, MAXT, MINB, false, false)       \
  MR_VP_KERNEL(k_dp_vp_##MAXT##_gw, MAXT, MINB, false, true)       \
  MR_VP_KERNEL(k_dp_vp_##MAXT##_c, MAXT, MINB, true, false)        \
  MR_VP_KERNEL(k_dp_vp_##MAXT##_cw, MAXT, MINB, true, true)
MR_VP_BUCKET(64, 12)
MR_VP_BUCKET(128, 6)
MR_VP_BUCKET(192, 4)
MR_VP_BUCKET(256, 3)
MR_VP_BUCKET(320, 2)
MR_VP_BUCKET(352, 2)
MR_VP_BUCKET(384, 2)

