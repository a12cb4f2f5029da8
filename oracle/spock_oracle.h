/*
 * spock_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, single-threaded, loop-faithful restatement of the reference solver
 * kul-optec/spock.jl (pure Julia) for the Chambolle-Pock / SuperMann hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (spock_b200/) never does.
 *
 * Parity status: PINNED.  This restatement reproduces the reference's committed traces
 * examples/server_heat/output/residuals.tikz (CP: 3159 iterations; SP: first iterations and the
 * K1/K2/fallback pattern) -- see tests/test_oracle_golden.py.  Julia itself cannot run in this image.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Indices: the reference is 1-based; this file stores 0-based arrays and documents the mapping.
 */
#ifndef SPOCK_ORACLE_H
#define SPOCK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* cone types of risk_measures.jl:5 (ConvexBaseCone) */
enum { SPO_CONE_ZEROS = 0, SPO_CONE_NONNEG = 1, SPO_CONE_NONPOS = 2, SPO_CONE_REALS = 3, SPO_CONE_SOC = 4 };
enum { SPO_ALG_CP = 1, SPO_ALG_SP = 2 }; /* model.jl:61-64 */

typedef struct {
  int N, d, nx, nu, ny;      /* N = number of stages incl. root (scenario_tree.jl:61) */
  const double *A, *B;       /* d modes; A[w]: nx*nx, B[w]: nx*nu, column-major (Julia layout) */
  int cost_dense;            /* 0: sqrt-cost matrices are Diagonal (store the diagonal), 1: dense column-major */
  int n_sqrtQ;               /* 1 = the same matrix for every node, or n-1 (Julia sqrtQ[j-1], j = 2..n) */
  const double *sqrtQ, *sqrtR;
  int n_sqrtQN;              /* 1 or n_leaf */
  const double *sqrtQN;
  const double *E, *F, *b;   /* uniform risk measure: E ny*d, F ny*d column-major, b ny (risk_measures.jl:79-84) */
  int n_cones;
  const int *cone_type, *cone_dim;
  const double *ls_matrix;   /* (ny+2d)^2 column-major orthogonal projector on ker[E' -I -I; F' 0 0] (cpock.jl:71-81,125-128) */
  double x_min, x_max, u_min, u_max; /* constraints.jl:11-22 */
  int algorithm;             /* SPO_ALG_CP / SPO_ALG_SP */
  double L_norm;             /* <=0: reference constants 10.0 (cpock.jl:63) / 3.28 (spock.jl:64) */
  const int *w;              /* node_info[j].w supplied by the caller (1-based mode, n entries, entry 0 ignored), or NULL: the
                                generator's (j % d) + 1 (scenario_tree.jl:77) */
  int n_rms;                 /* 0/1: uniform risk measures; n_non_leaf: b (ny each) and ls_matrix ((ny+2d)^2 each) hold one entry
                                per non-leaf node, back to back (rms[i].b implicit_l.jl:219,427; ls_matrix[i] cpock.jl:125-128) */
} spo_problem;

typedef struct spo_model spo_model;

spo_model *spo_create(const spo_problem *p, int n_threads);
void spo_destroy(spo_model *m);

long spo_nz(const spo_model *m);
long spo_nv(const spo_model *m);
long spo_n(const spo_model *m);
/* tree maps, written as the reference's 1-based values (scenario_tree.jl:61-109).
 * child_first[i-1] = first child of non-leaf i (children are child_first .. child_first+d-1)
 * anc[j-1] = ancestor of j (0 for the root), w[j-1] = mode (0 for the root), min_index[t] */
void spo_tree_maps(const spo_model *m, long *child_first, long *anc, int *w, long *min_index_per_timestep);
/* 0-based offsets of the z segments x,u,s,tau,y and v segments v1,v2,v3,v4,v5,v6,v7,v11,v12,v13,v14 */
void spo_offsets(const spo_model *m, long *z_off5, long *v_off11);

/* pointers to the live state (length nz / nv) */
double *spo_z(spo_model *m);
double *spo_v(spo_model *m);
void spo_set_x0(spo_model *m, const double *x0);

/* operators, all in place on caller vectors like the reference */
void spo_L(spo_model *m, const double *z, double *v);                 /* implicit_l.jl:177 */
void spo_Lt(spo_model *m, double *z, const double *v);                /* implicit_l.jl:320 */
void spo_mul(spo_model *m, double *C, int trans, const double *B, double alpha, double beta); /* 451 */
double spo_dot(spo_model *m, const double *a_z, const double *a_v, const double *b_z, const double *b_v,
               double gamma, double sigma);                           /* implicit_l.jl:485 */
void spo_projection_S1(spo_model *m, double *z1);                     /* implicit_l.jl:559 */
void spo_projection_S2(spo_model *m, double *z2);                     /* implicit_l.jl:688 */
void spo_prox_f(spo_model *m, double *arg, double gamma);             /* implicit_l.jl:729 */
void spo_prox_h_conj(spo_model *m, double *arg, double sigma);        /* implicit_l.jl:911 */
void spo_project_soc(double *vec, long len);                          /* risk_measures.jl:40 */

/* Riccati factors of node i (1-based) for inspection: P nx*nx, ABK nx*nx, K nu*nx, Rchol lower nu*nu */
void spo_factors(const spo_model *m, long i, double *P, double *ABK, double *K, double *Rchol_lower);

/* solve_model! (cpock.jl:169 / spock.jl:231): persistent state, optional z0/v0 (both or none), returns
 * the number of loop passes executed (the reference prints passes-1 when it breaks, or MAX_ITER).
 * verbose_log != 0 reproduces verbose=LOG termination (xi <= tol) and fills the trace buffers
 * (xi, xi1, xi2, backtrack_count per pass; each of capacity log_cap, may be NULL). */
typedef struct {
  int passes;           /* loop bodies executed */
  int printed_iter;     /* what the reference's println shows */
  int converged;
  double xi1, xi2, xi;
  int k1_count, k2_count, fallback_count;
} spo_result;

int spo_solve(spo_model *m, const double *x0, double tol, int verbose_log, const double *z0, const double *v0,
              double gamma, double sigma, int max_iter, spo_result *res, double *log_xi, double *log_xi1,
              double *log_xi2, int *log_bc, int log_cap);

/* SuperMann variants the reference's source contains but does not execute; 0 = as shipped.
 * SPO_SP_LINESEARCH: tau <- beta*tau after a failed K1/K2 test (sp.jl:439 discards the result of perform_linesearch!).
 * SPO_SP_K0: the blind update K0, whose test ends in "&& false" (sp.jl:80). */
#define SPO_SP_LINESEARCH 1
#define SPO_SP_K0 2
void spo_set_sp_options(spo_model *m, int options);

/* exactly K passes without looking at the termination flag (parity helper; still evaluates should_terminate!
 * because it has side effects on delta_z/delta_v/res_0) */
int spo_step(spo_model *m, int K, double gamma, double sigma);

#ifdef __cplusplus
}
#endif
#endif
