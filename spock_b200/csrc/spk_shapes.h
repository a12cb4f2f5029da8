// spk_shapes.h -- table entry of one shape-specialised set of fused kernels (see spk_shapes.def)
#pragma once
#include "spk_fused.cuh"
#include "spk_mm.cuh"

typedef void (*spk_fused_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs);
typedef void (*spk_dual2_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, int, long, long);
typedef void (*spk_dual3_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, long, long);
typedef void (*spk_top_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, SpkTopPlan);
typedef void (*spk_mm_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, const double *);
typedef void (*spk_mm_top_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, const double *, const double *, int);
struct SpkShapeEntry {
  int nx, nu, d, Tb, Tf;
  spk_fused_kernel_t bwd, fwd, dual;
  spk_dual2_kernel_t dual2;
  spk_dual3_kernel_t dual3[2][2][2];  // [children are leaves][uniform cost diagonals][with <r,Pr>]
  spk_dual3_kernel_t dual3_flat[2][2];  // non-leaf children, flat owned-family indices (subtree partition): [uniform cost][with <r,Pr>]
  spk_mm_kernel_t bwd_mm[2], fwd_mm;  // per-stage factors on the fp64 tensor pipe; bwd_mm[children are leaves]
  spk_mm_top_kernel_t top_mm;         // levels with <= 64 families, both sweeps, one CTA
  spk_top_kernel_t bwd_top, fwd_top;  // all narrow levels of a sweep in one launch (grid barrier between levels)
};
