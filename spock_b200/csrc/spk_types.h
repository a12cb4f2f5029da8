// spk_types.h -- problem dimensions, device pointer tables and index arithmetic shared by all kernels.
//
// Indexing contract (bit-exact with the reference, src/scenario_tree.jl:61-109): nodes are numbered
// breadth-first; kernels use 0-based node numbers i0 = i - 1:
//   children(i0)   = d*i0 + 1 + c,  c = 0..d-1          (Julia: d(i-1)+2 .. d(i-1)+d+1)
//   parent(j0)     = (j0 - 1) / d                        (Julia: (j-2) div d + 1)
//   mode(j0)       = (j0 + 1) % d   (0-based)            (Julia: w = (j % d) + 1)
//   first0(level t)= (d^t - 1)/(d - 1),  count = d^t     (Julia: min_index_per_timestep)
// Vector layouts are the reference's (implicit_l.jl:5-44,106-158; cpock.jl:51-60), see include/spock_b200.h.
#pragma once
#include "spk_runtime.h"

struct SpkDims {
  int N, d, nx, nu, ny, nls, n_cones;
  long n, nnl, nl, nz, nv;
  long zo[5];   // x, u, s, tau, y
  long vo[11];  // v1, v2, v3, v4, v5, v6, v7, v11, v12, v13, v14
  int cone_type[SPK_MAX_CONES], cone_dim[SPK_MAX_CONES];
  double x_min, x_max, u_min, u_max;
  int cost_dense, n_sqrtQ, n_sqrtQN;
  int rms_per_node;        // 1: b (ny) and the kernel projector ls (nls*nls) differ per non-leaf node (get_nonuniform_rms_avar_v2)
  int per_node;            // factor mode: 1 = one factor slot per node, 0 = per (stage, mode)
  long abk_stride;         // doubles per ABK slot (nx*nx rounded up to an even count: 16-byte aligned slots)
  long k_stride, r_stride; // doubles per K/Kt/G slot and per Rinv slot
};

struct SpkPtrs {
  const double *A, *B, *Bt;              // d modes: A nx*nx, B nx*nu, Bt = B' (nu*nx), column-major
  const double *sqrtQ, *sqrtR, *sqrtQN;  // diagonals or dense matrices
  const double *b, *ls;                  // risk vector b (ny), kernel projector (nls*nls)
  const double *ABK, *ABKt;              // closed-loop matrices A_w + B_w K and their transposes
  const double *K, *Kt, *Rinv, *G;       // feedback gain (nu*nx), its transpose, Rtilde^{-1}, sum ABK' P B
  const double *x0;
  const int *w;                          // 0-based dynamics mode of every node, or NULL: the generator's map (j % d) + 1 (scenario_tree.jl:77)
};

SPK_HD long spk_level_first0(int d, int t) {
  long p = 1;
  for (int k = 0; k < t; ++k) p *= d;
  return (p - 1) / (d - 1);
}
SPK_HD long spk_level_count(int d, int t) {
  long p = 1;
  for (int k = 0; k < t; ++k) p *= d;
  return p;
}

// row k of (S x) for a sqrt-cost matrix S that is Diagonal (stored as its diagonal) or dense symmetric
SPK_HD double spk_cost_row(const double *__restrict__ S, int dense, int dim, int k, const double *__restrict__ x) {
  if (!dense) return S[k] * x[k];
  double acc = 0.0;
  for (int m = 0; m < dim; ++m) acc += S[k + (long)m * dim] * x[m];
  return acc;
}

SPK_HD const double *spk_sqrtQ(const SpkDims &D, const SpkPtrs &P, long j0) {  // Julia sqrtQ[j-1]
  const long id = (D.n_sqrtQ == 1) ? 0 : (j0 - 1);
  return P.sqrtQ + id * (D.cost_dense ? (long)D.nx * D.nx : (long)D.nx);
}
SPK_HD const double *spk_sqrtR(const SpkDims &D, const SpkPtrs &P, long j0) {
  const long id = (D.n_sqrtQ == 1) ? 0 : (j0 - 1);
  return P.sqrtR + id * (D.cost_dense ? (long)D.nu * D.nu : (long)D.nu);
}
SPK_HD const double *spk_sqrtQN(const SpkDims &D, const SpkPtrs &P, long leaf) {
  const long id = (D.n_sqrtQN == 1) ? 0 : leaf;
  return P.sqrtQN + id * (D.cost_dense ? (long)D.nx * D.nx : (long)D.nx);
}

// Maximum of NON-NEGATIVE doubles (|.| of residual entries) that PROPAGATES NaN like LinearAlgebra.norm(., Inf) in
// should_terminate! (cp.jl:77-78, sp.jl:294-295): the IEEE bit patterns of non-negative doubles are monotone as
// integers and every sign-cleared NaN pattern lies above +Inf, so an integer maximum of the bit patterns is exact and
// keeps a NaN once it has been seen (fmax and `a > b ? a : b` both drop it: a diverged iterate would look converged).
SPK_HD long long spk_bits_nn(double a) {
#if !SPK_EMU && defined(__CUDA_ARCH__)
  return __double_as_longlong(a);
#else
  long long i;
  memcpy(&i, &a, sizeof(i));
  return i;
#endif
}
SPK_HD double spk_max_nn(double a, double b) { return spk_bits_nn(a) >= spk_bits_nn(b) ? a : b; }

// dynamics mode (0-based) of node j0: node_info[j].w - 1 (scenario_tree.jl:77), or a caller-supplied map
SPK_HD int spk_mode(const SpkPtrs &P, int d, long j0) { return P.w ? P.w[j0] : (int)((j0 + 1) % d); }

// risk vector b_i and kernel projector ls_matrix[i] of non-leaf node i0 (uniform risk measures: one copy for all nodes)
SPK_HD const double *spk_b(const SpkDims &D, const SpkPtrs &P, long i0) { return D.rms_per_node ? P.b + i0 * D.ny : P.b; }
SPK_HD const double *spk_ls(const SpkDims &D, const SpkPtrs &P, long i0) {
  return D.rms_per_node ? P.ls + i0 * (long)D.nls * D.nls : P.ls;
}

SPK_HD double spk_clip(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }  // constraints.jl:28-30
