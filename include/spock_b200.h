/*
 * spock_b200.h -- C ABI of libspock_b200.so: the Chambolle-Pock / SuperMann hot path of
 * kul-optec/spock.jl on one NVIDIA B200 (sm_100a) per process.
 *
 * This is what the reference's Julia layer `ccall`s (see INTEGRATION.md and julia/SpockB200.jl):
 * `build_model` -> spk_create, `solve_model!` -> spk_solve, and the inner operator interface that
 * src/model_algorithms/{cp,sp}.jl dispatch on (`L!`, `L_transpose!`, `prox_f!`, `prox_h_conj!`,
 * `spock_dot`) -> spk_apply / spk_dot for unit parity.  Plain pointers and sizes only; all host
 * vectors use the REFERENCE layout (z = [x|u|s|tau|y], v = [v1|v2|v3|v4|v5|v6|v7|v11|v12|v13|v14],
 * src/model_dynamics/implicit_l.jl:5-44,106-158; src/models/cpock.jl:51-60) so the Julia side keeps
 * `model.solver_state.z`, `.v` and the `*_inds` index vectors unchanged.
 *
 * Every function returns 0 (SPK_OK) or a spk_status code and never throws; spk_last_error() returns a
 * message for the calling thread's last failure.  One handle = one owner thread; calls are synchronous.
 * Multi-GPU: either one process per GPU (spk_problem.rank / world_size / comm_id; torchrun, MPI) or all GPUs of
 * the node behind one handle in one process (spk_problem.n_devices / devices).
 * There is no CPU fallback: without a CUDA device spk_create fails with SPK_ERR_CUDA.
 *
 * All matrices are column-major (Julia layout), all reals are IEEE binary64, node numbers are the
 * reference's 1-based BFS numbers (src/scenario_tree.jl:61-109).
 */
#ifndef SPOCK_B200_H
#define SPOCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spk_handle spk_handle;

typedef enum {
  SPK_OK = 0,
  SPK_ERR_INVALID = 1, /* bad argument: wrong length, d <= 1, unsupported option (implicit_l.jl:184, scenario_tree.jl:62, model.jl:295) */
  SPK_ERR_CUDA = 2,    /* CUDA runtime failure, or no device */
  SPK_ERR_ALLOC = 3,
  SPK_ERR_COMM = 4,    /* multi-GPU exchange failure */
  SPK_ERR_STATE = 5
} spk_status;

/* AlgorithmOptions, src/model.jl:61-64 */
typedef enum { SPK_CP = 1, SPK_SP = 2 } spk_algorithm;

/* ConvexBaseCone, src/risk_measures.jl:5 */
typedef enum { SPK_CONE_ZEROS = 0, SPK_CONE_NONNEG = 1, SPK_CONE_NONPOS = 2, SPK_CONE_REALS = 3, SPK_CONE_SOC = 4 } spk_cone;

/* Where the Riccati factors live:
 * PER_NODE  -- one copy per tree node like the reference (model.jl:99-102); HBM-bound, the graded mode.
 * PER_STAGE -- one copy per (stage, mode); valid for every tree the reference can generate. */
typedef enum { SPK_FACTORS_PER_NODE = 0, SPK_FACTORS_PER_STAGE = 1 } spk_factor_mode;

/* operators of the inner interface, src/model_dynamics/implicit_l.jl */
typedef enum {
  SPK_OP_L = 0,          /* out(nv) = L in(nz)                         L! :177            */
  SPK_OP_LT = 1,         /* out(nz) = L' in(nv)                        L_transpose! :320  */
  SPK_OP_PROX_F = 2,     /* out(nz) = prox_f^scalar(in)                prox_f! :729       */
  SPK_OP_PROX_HCONJ = 3, /* out(nv) = prox_{h*}^scalar(in)             prox_h_conj! :911  */
  SPK_OP_PROJ_S1 = 4,    /* out(n*nx+nnl*nu) = projection on dynamics  projection_S1! :559 */
  SPK_OP_PROJ_S2 = 5     /* out(nz-s_off-1) = kernel projection        projection_S2! :688 */
} spk_op;

/* Inputs of build_model (src/model.jl:273, src/models/cpock.jl:1-60) as flat arrays. */
typedef struct {
  int32_t N;  /* stages including the root: n = (d^N - 1)/(d - 1) nodes (scenario_tree.jl:61-71) */
  int32_t d;  /* branching factor = number of dynamics modes */
  int32_t nx, nu;
  const double *A; /* d matrices nx*nx */
  const double *B; /* d matrices nx*nu */
  /* sqrt of the cost matrices (cpock.jl:119-121 `map(sqrt, cost.Q)` stays on the caller's side) */
  int32_t cost_dense; /* 0: Diagonal matrices, pass the diagonals; 1: dense symmetric matrices */
  int32_t n_sqrtQ;    /* 1 (same for all nodes) or n-1 (entry j-2 belongs to node j, Julia sqrtQ[j-1]) */
  const double *sqrtQ, *sqrtR;
  int32_t n_sqrtQN;   /* 1 or n_leaf */
  const double *sqrtQN;
  /* uniform conic risk measure (risk_measures.jl:79-84): E ny*d, F ny*d, b ny, K = product of sub-cones */
  int32_t ny;
  const double *E, *F, *b;
  int32_t n_cones;
  const int32_t *cone_type; /* spk_cone */
  const int32_t *cone_dim;
  /* UniformRectangle, constraints.jl:11-22 */
  double x_min, x_max, u_min, u_max;
  int32_t algorithm;   /* spk_algorithm */
  double L_norm;       /* <= 0: the reference's constants 10.0 (cpock.jl:63) / 3.28 (spock.jl:64) */
  int32_t factor_mode; /* spk_factor_mode */
  int32_t device;      /* CUDA device ordinal, -1 = current device */
  /* multi-GPU: this process owns the subtrees rank*m .. (rank+1)*m-1 below the cut stage; world_size 1 = whole tree */
  int32_t rank, world_size;
  const uint8_t *comm_id; /* 128 bytes from spk_comm_unique_id (rank 0), broadcast by the host; NULL if world_size == 1 */
  /* multi-GPU inside ONE process (the reference's host is one Julia process, src/models/cpock.jl:169): n_devices > 1
   * partitions the tree over these CUDA ordinals behind a single handle -- one host thread, stream and NCCL rank per
   * device, bootstrapped internally; rank / world_size / comm_id must then be 0 / 1 / NULL.  Every call on the handle
   * fans out to all devices; spk_get_iterate assembles the complete (z, v) in the caller's host vectors.  0 or 1: `device`. */
  int32_t n_devices;
  const int32_t *devices;
  /* non-uniform risk measures (get_nonuniform_rms_avar_v2, risk_measures.jl:127-136; rms[i].E / .F / .b of implicit_l.jl:219,427
   * and the per-node ls_matrix of cpock.jl:71-81): 0 or 1 = one (E, F, b) for all nodes; n_non_leaf = E, F, b hold one entry
   * per non-leaf node, back to back (node i at offsets (i-1)*ny*d and (i-1)*ny).  ny and the sub-cone list stay uniform,
   * as in the reference (implicit_l.jl:14,19). */
  int32_t n_rms;
  /* caller-supplied mode map: mode_of_node[j-1] = node_info[j].w of the reference (scenario_tree.jl:77), the 1-based dynamics
   * mode that leads to node j (entry 0, the root, is ignored); NULL = the generator's map w(j) = (j % d) + 1.  With a map
   * the Riccati factors differ from node to node inside a stage, so `ricatti_offline` (implicit_l.jl:46-97) runs per node --
   * batched on the device, one CTA per node, level by level -- and factor_mode must be SPK_FACTORS_PER_NODE (single GPU). */
  const int32_t *mode_of_node;
} spk_problem;

typedef struct {
  int32_t passes;       /* loop bodies executed */
  int32_t printed_iter; /* the number the reference prints: "Finished in N iterations." (cp.jl:230, sp.jl:468) */
  int32_t converged;
  double xi1, xi2, xi;  /* last residuals (xi only with verbose_log) */
  int32_t k1_count, k2_count, fallback_count; /* SuperMann branch counts (sp.jl:411-446) */
  double solve_ms;      /* device time of the loop, CUDA events on the solver stream */
} spk_result;

const char *spk_last_error(void);
int spk_version(void);

int spk_create(const spk_problem *problem, spk_handle **out);
int spk_destroy(spk_handle *h);

/* sizes[0..4] = n, n_non_leaf, n_leaf, nz, nv  (get_nz / get_nv, implicit_l.jl:5-44) */
int spk_sizes(const spk_handle *h, int64_t sizes[5]);
/* 0-based offsets of x,u,s,tau,y in z and of v1,v2,v3,v4,v5,v6,v7,v11,v12,v13,v14 in v */
int spk_offsets(const spk_handle *h, int64_t z_off[5], int64_t v_off[11]);
/* Tree index maps with the reference's 1-based values (scenario_tree.jl:73-107); any pointer may be NULL.
 * child_first[i-1] (i <= n_non_leaf): children are child_first .. child_first+d-1;
 * anc[j-1] (0 for the root); w[j-1] mode (0 for the root); min_index_per_timestep[0..N-1]. */
int spk_tree_maps(int32_t N, int32_t d, int64_t *child_first, int64_t *anc, int32_t *w,
                  int64_t *min_index_per_timestep);

/* problem_definition.x0 (cpock.jl:185) */
int spk_set_x0(spk_handle *h, const double *x0);
/* solver_state.z / .v, host vectors in the reference layout; either pointer may be NULL */
int spk_set_iterate(spk_handle *h, const double *z, const double *v);
int spk_get_iterate(spk_handle *h, double *z, double *v);

/* One operator application on host vectors (lengths as listed at spk_op).  `scalar` is gamma / sigma. */
int spk_apply(spk_handle *h, int32_t op, const double *in, double *out, double scalar);
/* spock_mul!(model, C, trans, B, alpha, beta), implicit_l.jl:451-481: C <- alpha * op(L) B + beta * C on host vectors;
 * trans != 0: op(L) = L' (B has nv entries, C nz), else L (B nz, C nv). */
int spk_mul(spk_handle *h, int32_t trans, const double *B, double alpha, double beta, double *C);
/* spock_dot(model, a_z, a_v, b_z, b_v, gamma, sigma), implicit_l.jl:485 */
int spk_dot(spk_handle *h, const double *a_z, const double *a_v, const double *b_z, const double *b_v, double gamma,
            double sigma, double *out);

/* Offline factors of node i (1-based) as the reference stores them (implicit_l.jl:46-97): P nx*nx,
 * ABK nx*nx (i >= 2), K nu*nx and Rtilde^{-1} nu*nu (i <= n_non_leaf); plus the kernel projector
 * ls_matrix (ny+2d)^2 (cpock.jl:125-128).  Any pointer may be NULL. */
int spk_get_factors(const spk_handle *h, int64_t i, double *P, double *ABK, double *K, double *Rinv, double *ls_matrix);

/* Exactly K passes of the CP / SuperMann loop on the device-resident iterate, termination flag ignored
 * (parity and benchmarking).  gamma, sigma <= 0: 0.99/sqrt(L_norm) (cp.jl:198-201).
 * ms (may be NULL) receives the device time of the K passes. */
int spk_step(spk_handle *h, int32_t K, double gamma, double sigma, float *ms);

/* solve_model!(model, x0; tol, verbose, z0, v0, gamma, sigma)  (cpock.jl:169-189, spock.jl:231-251).
 * z0/v0: both or neither (the reference's guard); NULL keeps the iterate of the previous solve (warm start).
 * max_iter <= 0: 5000 (CP) / 1000 (SP).  verbose_log != 0 reproduces verbose=LOG: termination on xi <= tol
 * and per-pass logs (each log_* may be NULL, capacity log_cap entries). */
int spk_solve(spk_handle *h, const double *x0, double tol, int32_t max_iter, double gamma, double sigma,
              int32_t verbose_log, const double *z0, const double *v0, spk_result *result, double *log_xi,
              double *log_xi1, double *log_xi2, int32_t *log_backtrack, int32_t log_cap);

/* MPC serving (examples/server_heat/mpc_simulation.jl:38-183).
 * spk_shift_iterate: shift-initialisation of the next receding-horizon solve -- after the first input has been applied and
 * the disturbance observed (`child` = 0-based ordinal of the realised child of the root), every node of the new problem
 * takes the warm-start values of the corresponding node one stage deeper in that child's subtree; the last stage keeps
 * its values.  (The reference re-solves from the unshifted iterate; this is an optional better initial guess.)
 * spk_solve_batch: solve_model! (SILENT, default step sizes, warm start) for n independent models that share the GPU --
 * the device-resident loops of all handles run concurrently on their own streams.  x0s = n vectors of nx, back to back. */
int spk_shift_iterate(spk_handle *h, int32_t child);
int spk_solve_batch(spk_handle **handles, int32_t n, const double *x0s, double tol, int32_t max_iter, spk_result *results);

/* Solver memory that persists ACROSS solves besides (z, v): checkpoint / resume and bit-exact warm starts (MPC).
 * CP (model.jl:80-125): [z_old | v_old]                                                          -> nz+nv doubles.
 * SP (model.jl:137-226, spock.jl:160-202): [z_old|v_old], [rz_old|rv_old], [delta_z|delta_v], the Anderson columns
 * MP(:,1..3), MR(:,1..3), aa_Q(:,1..3) in the reference's column order, aa_R (3x3 column-major)     -> 12(nz+nv)+9 doubles.
 * Together with spk_get_iterate / spk_set_iterate this is everything solve_model! reads from the previous solve. */
int spk_state_size(const spk_handle *h, int64_t *doubles);
int spk_get_state(spk_handle *h, double *state);
int spk_set_state(spk_handle *h, const double *state);

/* ||L||^2 by power iteration on L'L with the device L / L' kernels (iters <= 0: 50).  The reference hard-codes
 * L_norm = 10 (cpock.jl:62-63) / 3.28 (spock.jl:63-64), which is only valid for its server_heat data;
 * spk_set_L_norm replaces the value behind the default step sizes 0.99/sqrt(L_norm) (cp.jl:198-201). */
int spk_estimate_L_norm(spk_handle *h, int32_t iters, double *L_norm);
int spk_set_L_norm(spk_handle *h, double L_norm);

/* SuperMann variants that the reference's source contains but does not execute.  Default 0 =
 * the reference as shipped (parity mode).  SPK_SP_LINESEARCH: tau <- beta tau after a failed K1 / K2 test -- sp.jl:439 discards
 * the result of perform_linesearch!, so the reference re-evaluates the same candidate nine times.  SPK_SP_K0: the blind update
 * K0 (sp.jl:72-106), whose test ends in "&& false" (sp.jl:80).  With an option set the loop is driven from the host. */
#define SPK_SP_LINESEARCH 1
#define SPK_SP_K0 2
int spk_set_sp_options(spk_handle *h, int32_t options);

/* Residuals of the last pass without a full D2H of the iterate: out[0] = xi1, out[1] = xi2 */
int spk_get_residuals(spk_handle *h, double out[2]);

/* Per-kernel-class device time accumulated since the last reset (CUDA events on the solver stream).
 * class ids: 0 = S1 backward sweep, 1 = S1 forward sweep, 2 = L/L', 3 = prox_h*, 4 = S2, 5 = vector ops,
 * 6 = fused backward family kernel, 7 = fused forward family kernel.  Profiling adds event overhead;
 * enable only for measurement runs. */
int spk_profile(spk_handle *h, int32_t enable, int32_t reset);
int spk_profile_get(spk_handle *h, int32_t class_id, double *ms, int64_t *launches);
/* total kernel launches issued by this handle since creation */
int spk_launch_count(const spk_handle *h, int64_t *launches);

/* Bytes of device memory owned by the handle */
int spk_device_bytes(const spk_handle *h, int64_t *bytes);

/* multi-GPU bootstrap: rank 0 fills 128 bytes, the host broadcasts them to every rank's spk_problem.comm_id */
int spk_comm_unique_id(uint8_t id[128]);
/* The non-leaf nodes whose families this rank computes, as contiguous ranges of 1-based node numbers (at most 64
 * ranges; the replicated levels above the cut stage come first).  With world_size > 1 spk_get_iterate returns the
 * rank's local view: entries of owned families (and of the replicated top) are valid, the rest is stale. */
int spk_owned_ranges(const spk_handle *h, int32_t *n_ranges, int64_t *start, int64_t *count);

/* Collective over all ranks (no-op with world_size 1): all-gathers the entries every rank's families own, so that
 * afterwards the device-resident iterate -- and hence spk_get_iterate -- is the complete (z, v) on every rank
 * (model.solver_state.z / .v of the reference).  Every level below the cut is a rank-major contiguous slice of every
 * segment, so this is one in-place all-gather per (segment, level). */
int spk_gather_iterate(spk_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SPOCK_B200_H */
