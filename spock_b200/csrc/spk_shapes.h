// spk_shapes.h -- table entry of one shape-specialised set of fused kernels (see spk_shapes.def)
#pragma once
#include "spk_fused.cuh"

typedef void (*spk_fused_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs);
typedef void (*spk_dual2_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, int, long, long);
typedef void (*spk_dual3_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, long, long);
struct SpkShapeEntry {
  int nx, nu, d, Tb, Tf;
  spk_fused_kernel_t bwd, fwd, dual;
  spk_dual2_kernel_t dual2;
  spk_dual3_kernel_t dual3[2][2];  // [children are leaves][uniform cost diagonals]
};
