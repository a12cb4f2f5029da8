/*
 * spock_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See spock_oracle.h.
 *
 * Loop-faithful C restatement of kul-optec/spock.jl's CP / SuperMann hot path.  Single-threaded by
 * default like the reference; n_threads > 1 only parallelises loops over independent nodes
 * (used for the "all host cores" CPU baseline), never changing any per-node arithmetic.
 *
 * Reference citations are file:line relative to /root/reference.
 */
#include "spock_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AA_M 3 /* ANDERSON_BUFFER_SIZE, anderson.jl:6, spock.jl:108 */

struct spo_model {
  int sp_options; /* SPO_SP_LINESEARCH | SPO_SP_K0; 0 = the reference as shipped */
  int custom_w;   /* node modes supplied by the caller */
  long n_rms;     /* 1 = uniform risk measures; nnl = per-node b and ls_matrix (get_nonuniform_rms_avar_v2) */
  int N, d, nx, nu, ny, nt, alg;
  long n, nnl, nl, nz, nv;
  long zo[5];  /* x,u,s,tau,y   (implicit_l.jl:106-158) */
  long vo[11]; /* v1,v2,v3,v4,v5,v6,v7,v11,v12,v13,v14 (cpock.jl:51-60) */
  long *child_first, *anc, *min_idx;
  int *w;
  double *A, *B;
  int cost_dense, n_sqrtQ, n_sqrtQN;
  double *sqrtQ, *sqrtR, *sqrtQN;
  double *E, *F, *b, *ls;
  int n_cones, *cone_type, *cone_dim;
  double x_min, x_max, u_min, u_max, L_norm;
  double *x0;
  /* Riccati (implicit_l.jl:46-97), per node like the reference */
  double *P, *K, *Rc, *ABK;
  /* GENERIC_SOLVER_STATE (model.jl:23-44) */
  double *z, *v, *zbar, *vbar, *z_workspace, *rz, *rv, *xi_1, *xi_2, *xi, *delta_z, *delta_v, *z_old, *v_old,
      *delta_rz, *delta_rv;
  double res_0[2];
  /* internal workspaces */
  double *mul_buf_z, *mul_buf_v, *prox_v_ws, *ric_q, *ric_d;
  /* SuperMann / Anderson (model.jl:137-226) */
  double *ws_rho_z, *ws_rho_v, *dz, *dv, *sw, *su, *wbar, *ubar, *rw, *ru, *w_workspace, *ws_Lz, *ws_Lv, *rz_old,
      *rv_old, *MP, *MR, *aaQ, *aa_wsp;
  double aaR[AA_M * AA_M]; /* column-major upper triangular */
  double aa_gamma[AA_M];
};

/* ------------------------------------------------------------------------------------------ helpers */

static double *dalloc(long k) {
  double *p = (double *)calloc((size_t)(k > 0 ? k : 1), sizeof(double));
  if (!p) {
    fprintf(stderr, "spock_oracle: out of memory (%ld doubles)\n", k);
    abort();
  }
  return p;
}
static double *dcopy(const double *src, long k) {
  double *p = dalloc(k);
  memcpy(p, src, (size_t)k * sizeof(double));
  return p;
}
static void fill(double *p, long k, double val) {
  for (long i = 0; i < k; ++i) p[i] = val;
}

/* y = M x, M rows*cols column-major (LA.mul!(y, M, x)) */
static void gemv_n(const double *M, int rows, int cols, const double *x, double *y) {
  for (int r = 0; r < rows; ++r) y[r] = 0.0;
  for (int c = 0; c < cols; ++c) {
    const double xc = x[c];
    const double *col = M + (long)c * rows;
    for (int r = 0; r < rows; ++r) y[r] += col[r] * xc;
  }
}
/* y = M' x (LA.mul!(y, M', x)) */
static void gemv_t(const double *M, int rows, int cols, const double *x, double *y) {
  for (int c = 0; c < cols; ++c) {
    const double *col = M + (long)c * rows;
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) acc += col[r] * x[r];
    y[c] = acc;
  }
}
/* C = A*B, A m*k, B k*n, all column-major */
static void gemm_nn(const double *A, const double *B, double *C, int m, int k, int n) {
  for (int j = 0; j < n; ++j) gemv_n(A, m, k, B + (long)j * k, C + (long)j * m);
}
/* C = A'*B, A k*m, B k*n */
static void gemm_tn(const double *A, const double *B, double *C, int m, int k, int n) {
  for (int j = 0; j < n; ++j) gemv_t(A, k, m, B + (long)j * k, C + (long)j * m);
}
/* in-place lower Cholesky of the symmetric n*n matrix S (column-major); returns 0 on success */
static int chol_lower(double *S, int n) {
  for (int j = 0; j < n; ++j) {
    double djj = S[j + (long)j * n];
    for (int k = 0; k < j; ++k) djj -= S[j + (long)k * n] * S[j + (long)k * n];
    if (!(djj > 0.0)) return 1;
    djj = sqrt(djj);
    S[j + (long)j * n] = djj;
    for (int i = j + 1; i < n; ++i) {
      double s = S[i + (long)j * n];
      for (int k = 0; k < j; ++k) s -= S[i + (long)k * n] * S[j + (long)k * n];
      S[i + (long)j * n] = s / djj;
    }
    for (int i = 0; i < j; ++i) S[i + (long)j * n] = 0.0;
  }
  return 0;
}
/* x <- (L L')^{-1} x  (LA.ldiv!(::Cholesky, x)) */
static void chol_solve(const double *Lf, int n, double *x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= Lf[i + (long)k * n] * x[k];
    x[i] = s / Lf[i + (long)i * n];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= Lf[k + (long)i * n] * x[k];
    x[i] = s / Lf[i + (long)i * n];
  }
}

static inline const double *sqrtQ_of(const spo_model *m, long j_ind /* Julia j-1, 1-based */) {
  long id = (m->n_sqrtQ == 1) ? 0 : (j_ind - 1);
  return m->sqrtQ + id * (m->cost_dense ? (long)m->nx * m->nx : (long)m->nx);
}
static inline const double *sqrtR_of(const spo_model *m, long j_ind) {
  long id = (m->n_sqrtQ == 1) ? 0 : (j_ind - 1);
  return m->sqrtR + id * (m->cost_dense ? (long)m->nu * m->nu : (long)m->nu);
}
static inline const double *sqrtQN_of(const spo_model *m, long leaf_ord /* 1-based */) {
  long id = (m->n_sqrtQN == 1) ? 0 : (leaf_ord - 1);
  return m->sqrtQN + id * (m->cost_dense ? (long)m->nx * m->nx : (long)m->nx);
}
/* y = S x for a sqrt-cost matrix that is either Diagonal or dense symmetric */
static void cost_mul(const spo_model *m, const double *S, int dim, const double *x, double *y) {
  if (m->cost_dense) {
    gemv_n(S, dim, dim, x, y);
  } else {
    for (int k = 0; k < dim; ++k) y[k] = S[k] * x[k];
  }
}

/* ------------------------------------------------------------------------------------------ build */

/* scenario_tree.jl:61-109 */
static void build_tree(spo_model *m, const int *w_override) {
  const long d = m->d;
  long dN = 1;
  for (int k = 0; k < m->N; ++k) dN *= d; /* d^N */
  m->n = (dN - 1) / (d - 1);                /* :67 */
  m->nl = dN / d;                           /* :69 d^(N-1) */
  m->nnl = (dN / d - 1) / (d - 1);          /* :71 */
  m->child_first = (long *)calloc((size_t)(m->nnl > 0 ? m->nnl : 1), sizeof(long));
  m->anc = (long *)calloc((size_t)m->n, sizeof(long));
  m->w = (int *)calloc((size_t)m->n, sizeof(int));
  m->min_idx = (long *)calloc((size_t)m->N, sizeof(long));
  for (long i = 1; i <= m->n; ++i) m->w[i - 1] = (i > 1) ? (int)(i % d) + 1 : 0; /* :77 */
  if (w_override) { /* the caller's node_info[j].w */
    for (long i = 2; i <= m->n; ++i) m->w[i - 1] = w_override[i - 1];
    m->custom_w = 1;
  }
  long child_index = 2;                                                           /* :82-88 */
  for (long i = 1; i <= m->nnl; ++i) {
    m->child_first[i - 1] = child_index;
    for (long c = child_index; c < child_index + d; ++c) m->anc[c - 1] = i; /* :90-95 */
    child_index += d;
  }
  m->min_idx[0] = 1; /* :107 */
  long dt = 1;
  for (int t = 1; t <= m->N - 1; ++t) {
    dt *= d;
    m->min_idx[t] = 1 + (dt - 1) / (d - 1);
  }
}

/* implicit_l.jl:46-97.  The reference loops over every non-leaf node; all nodes of one stage see
 * identical inputs, so the factors are computed once per stage with the reference's formulas and
 * replicated to the stage's other nodes (bit-identical; tests/test_oracle_units.py checks against the
 * node-by-node loop enabled with SPO_FAITHFUL_OFFLINE=1). */
static void riccati_node(spo_model *m, long i, double *wk) {
  const int nx = m->nx, nu = m->nu, d = m->d;
  double *sum_r = wk, *sum_k = sum_r + nu * nu, *sum_p = sum_k + nu * nx, *PB = sum_p + nx * nx,
         *PA = PB + nx * nu, *T1 = PA + nx * nx, *T2 = T1 + nx * nx;
  memset(sum_r, 0, sizeof(double) * nu * nu);
  memset(sum_k, 0, sizeof(double) * nu * nx);
  const long c0 = m->child_first[i - 1];
  for (long j = c0; j < c0 + d; ++j) { /* :76-79 */
    const double *Aw = m->A + (long)(m->w[j - 1] - 1) * nx * nx;
    const double *Bw = m->B + (long)(m->w[j - 1] - 1) * nx * nu;
    const double *Pj = m->P + (j - 1) * (long)nx * nx;
    gemm_nn(Pj, Bw, PB, nx, nx, nu);
    gemm_nn(Pj, Aw, PA, nx, nx, nx);
    gemm_tn(Bw, PB, T1, nu, nx, nu); /* B' P B */
    for (int k = 0; k < nu * nu; ++k) sum_r[k] += T1[k];
    gemm_tn(Bw, PA, T1, nu, nx, nx); /* B' P A */
    for (int k = 0; k < nu * nx; ++k) sum_k[k] += T1[k];
  }
  double *Rc = m->Rc + (i - 1) * (long)nu * nu;
  for (int c = 0; c < nu; ++c) /* :82-84 symmetrise, add I */
    for (int r = 0; r < nu; ++r)
      Rc[r + c * nu] = ((r == c) ? 1.0 : 0.0) + 0.5 * (sum_r[r + c * nu] + sum_r[c + r * nu]);
  if (chol_lower(Rc, nu)) {
    fprintf(stderr, "spock_oracle: Cholesky failed at node %ld\n", i);
    abort();
  }
  double *Ki = m->K + (i - 1) * (long)nu * nx; /* :86 K = chol \ (-sum_for_k) */
  for (int c = 0; c < nx; ++c) {
    for (int r = 0; r < nu; ++r) Ki[r + c * nu] = -sum_k[r + c * nu];
    chol_solve(Rc, nu, Ki + (long)c * nu);
  }
  memset(sum_p, 0, sizeof(double) * nx * nx);
  for (long j = c0; j < c0 + d; ++j) { /* :89-92 */
    const double *Aw = m->A + (long)(m->w[j - 1] - 1) * nx * nx;
    const double *Bw = m->B + (long)(m->w[j - 1] - 1) * nx * nu;
    const double *Pj = m->P + (j - 1) * (long)nx * nx;
    double *ABKj = m->ABK + (j - 1) * (long)nx * nx;
    gemm_nn(Bw, Ki, ABKj, nx, nu, nx);
    for (int k = 0; k < nx * nx; ++k) ABKj[k] = Aw[k] + ABKj[k];
    gemm_nn(Pj, ABKj, T1, nx, nx, nx);
    gemm_tn(ABKj, T1, T2, nx, nx, nx);
    for (int k = 0; k < nx * nx; ++k) sum_p[k] += T2[k];
  }
  double *Pi = m->P + (i - 1) * (long)nx * nx; /* :93 */
  gemm_tn(Ki, Ki, T1, nx, nu, nx);
  for (int c = 0; c < nx; ++c)
    for (int r = 0; r < nx; ++r) Pi[r + c * nx] = ((r == c) ? 1.0 : 0.0) + T1[r + c * nx] + sum_p[r + c * nx];
}

static void riccati_offline(spo_model *m) {
  const int nx = m->nx, nu = m->nu, d = m->d;
  m->P = dalloc(m->n * (long)nx * nx);
  m->ABK = dalloc(m->n * (long)nx * nx);
  m->K = dalloc(m->nnl * (long)nu * nx);
  m->Rc = dalloc(m->nnl * (long)nu * nu);
  for (long i = m->nnl + 1; i <= m->n; ++i) /* :67-69 */
    for (int k = 0; k < nx; ++k) m->P[(i - 1) * (long)nx * nx + k + (long)k * nx] = 1.0;
  double *wk = dalloc((long)nu * nu + 2L * nu * nx + 4L * nx * nx + 8);
  const char *env = getenv("SPO_FAITHFUL_OFFLINE");
  if ((env && env[0] == '1') || m->custom_w) { /* with a caller-supplied mode map the nodes of a stage differ: node by node */
    for (long i = m->nnl; i >= 1; --i) riccati_node(m, i, wk); /* :71 */
  } else {
    for (int t = m->N - 2; t >= 0; --t) {
      const long i0 = m->min_idx[t], i1 = m->min_idx[t + 1]; /* nodes i0 .. i1-1 */
      riccati_node(m, i0, wk);
      const long c0 = m->child_first[i0 - 1];
#pragma omp parallel for schedule(static) if (m->nt > 1) num_threads(m->nt)
      for (long i = i0 + 1; i < i1; ++i) {
        memcpy(m->P + (i - 1) * (long)nx * nx, m->P + (i0 - 1) * (long)nx * nx, sizeof(double) * nx * nx);
        memcpy(m->K + (i - 1) * (long)nu * nx, m->K + (i0 - 1) * (long)nu * nx, sizeof(double) * nu * nx);
        memcpy(m->Rc + (i - 1) * (long)nu * nu, m->Rc + (i0 - 1) * (long)nu * nu, sizeof(double) * nu * nu);
        const long ci = m->child_first[i - 1];
        for (int c = 0; c < d; ++c)
          memcpy(m->ABK + (ci + c - 1) * (long)nx * nx, m->ABK + (c0 + c - 1) * (long)nx * nx,
                 sizeof(double) * nx * nx);
      }
    }
  }
  free(wk);
}

spo_model *spo_create(const spo_problem *p, int n_threads) {
  if (p->d <= 1) { /* scenario_tree.jl:62-64 */
    fprintf(stderr, "Branching factor d must be larger than 1, but is %d.\n", p->d);
    return NULL;
  }
  spo_model *m = (spo_model *)calloc(1, sizeof(spo_model));
  m->N = p->N; m->d = p->d; m->nx = p->nx; m->nu = p->nu; m->ny = p->ny;
  m->nt = n_threads > 1 ? n_threads : 1;
  m->alg = p->algorithm;
  build_tree(m, p->w);
  const long n = m->n, nnl = m->nnl, nl = m->nl;
  const int nx = m->nx, nu = m->nu, ny = m->ny, d = m->d;
  /* get_nz implicit_l.jl:5-15 and z_to_* :106-158 */
  m->zo[0] = 0;
  m->zo[1] = n * nx;
  m->zo[2] = m->zo[1] + nnl * nu;
  m->zo[3] = m->zo[2] + n;
  m->zo[4] = m->zo[3] + (n - 1);
  m->nz = m->zo[4] + nnl * ny;
  /* get_nv implicit_l.jl:17-44, offsets cpock.jl:51-60 */
  m->vo[0] = 0;
  m->vo[1] = nnl * ny;
  m->vo[2] = m->vo[1] + nnl;
  m->vo[3] = m->vo[2] + (long)nx * (n - 1);
  m->vo[4] = m->vo[3] + (long)nu * (n - 1);
  m->vo[5] = m->vo[4] + (n - 1);
  m->vo[6] = m->vo[5] + (n - 1);
  m->vo[7] = m->vo[6] + nnl * (nx + nu);
  m->vo[8] = m->vo[7] + nl * nx;
  m->vo[9] = m->vo[8] + nl;
  m->vo[10] = m->vo[9] + nl;
  m->nv = m->vo[10] + nl * nx;

  m->A = dcopy(p->A, (long)d * nx * nx);
  m->B = dcopy(p->B, (long)d * nx * nu);
  m->cost_dense = p->cost_dense;
  m->n_sqrtQ = p->n_sqrtQ;
  m->n_sqrtQN = p->n_sqrtQN;
  m->sqrtQ = dcopy(p->sqrtQ, (long)p->n_sqrtQ * (p->cost_dense ? nx * nx : nx));
  m->sqrtR = dcopy(p->sqrtR, (long)p->n_sqrtQ * (p->cost_dense ? nu * nu : nu));
  m->sqrtQN = dcopy(p->sqrtQN, (long)p->n_sqrtQN * (p->cost_dense ? nx * nx : nx));
  m->E = dcopy(p->E, (long)ny * d);
  m->F = dcopy(p->F, (long)ny * d);
  m->n_rms = p->n_rms > 1 ? p->n_rms : 1; /* rms[i].b (implicit_l.jl:219,427), ls_matrix[i] (cpock.jl:125-128) */
  m->b = dcopy(p->b, (long)m->n_rms * ny);
  m->ls = dcopy(p->ls_matrix, (long)m->n_rms * (ny + 2 * d) * (ny + 2 * d));
  m->n_cones = p->n_cones;
  m->cone_type = (int *)calloc((size_t)p->n_cones, sizeof(int));
  m->cone_dim = (int *)calloc((size_t)p->n_cones, sizeof(int));
  memcpy(m->cone_type, p->cone_type, sizeof(int) * p->n_cones);
  memcpy(m->cone_dim, p->cone_dim, sizeof(int) * p->n_cones);
  m->x_min = p->x_min; m->x_max = p->x_max; m->u_min = p->u_min; m->u_max = p->u_max;
  m->L_norm = p->L_norm > 0 ? p->L_norm : (p->algorithm == SPO_ALG_SP ? 3.28 : 10.0); /* spock.jl:64 / cpock.jl:63 */
  m->x0 = dalloc(nx);

  riccati_offline(m);

  const long nz = m->nz, nv = m->nv;
  m->z = dalloc(nz); m->v = dalloc(nv); m->zbar = dalloc(nz); m->vbar = dalloc(nv);
  m->z_workspace = dalloc(nz); m->rz = dalloc(nz); m->rv = dalloc(nv);
  m->xi_1 = dalloc(nz); m->xi_2 = dalloc(nv); m->xi = dalloc(nz);
  m->delta_z = dalloc(nz); m->delta_v = dalloc(nv);
  m->z_old = dalloc(nz); m->v_old = dalloc(nv);
  fill(m->z_old, nz, 1.0); fill(m->v_old, nv, 1.0); /* cpock.jl:98-99 ones(nz), ones(nv) */
  m->delta_rz = dalloc(nz); m->delta_rv = dalloc(nv);
  m->res_0[0] = -INFINITY; m->res_0[1] = -INFINITY;
  m->mul_buf_z = dalloc(nz); m->mul_buf_v = dalloc(nv); m->prox_v_ws = dalloc(nv);
  m->ric_q = dalloc(n * nx); m->ric_d = dalloc(n * nu);
  if (m->alg == SPO_ALG_SP) { /* spock.jl:160-202 */
    m->ws_rho_z = dalloc(nz); m->ws_rho_v = dalloc(nv); m->dz = dalloc(nz); m->dv = dalloc(nv);
    m->sw = dalloc(nz); m->su = dalloc(nv); m->wbar = dalloc(nz); m->ubar = dalloc(nv);
    m->rw = dalloc(nz); m->ru = dalloc(nv); m->w_workspace = dalloc(nz);
    m->ws_Lz = dalloc(nz); m->ws_Lv = dalloc(nv);
    m->rz_old = dalloc(nz); m->rv_old = dalloc(nv);
    fill(m->rz_old, nz, 1.0); fill(m->rv_old, nv, 1.0); /* spock.jl:187-188 */
    m->MP = dalloc((nz + nv) * AA_M); m->MR = dalloc((nz + nv) * AA_M); m->aaQ = dalloc((nz + nv) * AA_M);
    m->aa_wsp = dalloc(nz + nv);
    for (int k = 0; k < AA_M; ++k) m->aaR[k + k * AA_M] = 1.0; /* spock.jl:200 */
  } else {
    m->ws_Lz = dalloc(nz); m->ws_Lv = dalloc(nv); /* used by spo_dot for unit tests only */
  }
  return m;
}

void spo_destroy(spo_model *m) {
  if (!m) return;
  double *ptrs[] = {m->A, m->B, m->sqrtQ, m->sqrtR, m->sqrtQN, m->E, m->F, m->b, m->ls, m->x0, m->P, m->K, m->Rc,
                    m->ABK, m->z, m->v, m->zbar, m->vbar, m->z_workspace, m->rz, m->rv, m->xi_1, m->xi_2, m->xi,
                    m->delta_z, m->delta_v, m->z_old, m->v_old, m->delta_rz, m->delta_rv, m->mul_buf_z,
                    m->mul_buf_v, m->prox_v_ws, m->ric_q, m->ric_d, m->ws_rho_z, m->ws_rho_v, m->dz, m->dv, m->sw,
                    m->su, m->wbar, m->ubar, m->rw, m->ru, m->w_workspace, m->ws_Lz, m->ws_Lv, m->rz_old,
                    m->rv_old, m->MP, m->MR, m->aaQ, m->aa_wsp};
  for (size_t k = 0; k < sizeof(ptrs) / sizeof(ptrs[0]); ++k) free(ptrs[k]);
  free(m->child_first); free(m->anc); free(m->w); free(m->min_idx); free(m->cone_type); free(m->cone_dim);
  free(m);
}

long spo_nz(const spo_model *m) { return m->nz; }
long spo_nv(const spo_model *m) { return m->nv; }
long spo_n(const spo_model *m) { return m->n; }
double *spo_z(spo_model *m) { return m->z; }
double *spo_v(spo_model *m) { return m->v; }
void spo_set_x0(spo_model *m, const double *x0) { memcpy(m->x0, x0, sizeof(double) * m->nx); }

void spo_tree_maps(const spo_model *m, long *child_first, long *anc, int *w, long *min_index) {
  memcpy(child_first, m->child_first, sizeof(long) * m->nnl);
  memcpy(anc, m->anc, sizeof(long) * m->n);
  memcpy(w, m->w, sizeof(int) * m->n);
  memcpy(min_index, m->min_idx, sizeof(long) * m->N);
}
void spo_offsets(const spo_model *m, long *z_off5, long *v_off11) {
  memcpy(z_off5, m->zo, sizeof(m->zo));
  memcpy(v_off11, m->vo, sizeof(m->vo));
}
void spo_factors(const spo_model *m, long i, double *P, double *ABK, double *K, double *Rc) {
  const int nx = m->nx, nu = m->nu;
  memcpy(P, m->P + (i - 1) * (long)nx * nx, sizeof(double) * nx * nx);
  memcpy(ABK, m->ABK + (i - 1) * (long)nx * nx, sizeof(double) * nx * nx);
  if (i <= m->nnl) {
    memcpy(K, m->K + (i - 1) * (long)nu * nx, sizeof(double) * nu * nx);
    memcpy(Rc, m->Rc + (i - 1) * (long)nu * nu, sizeof(double) * nu * nu);
  }
}

/* ------------------------------------------------------------------------------------------ L, L' */

#define PAR_FOR _Pragma("omp parallel for schedule(static) if (m->nt > 1) num_threads(m->nt)")

/* implicit_l.jl:177-318 */
void spo_L(spo_model *m, const double *z, double *v) {
  const int nx = m->nx, nu = m->nu, ny = m->ny, d = m->d;
  const long n = m->n, nnl = m->nnl, nl = m->nl;
  const long *zo = m->zo, *vo = m->vo;
  const long leaf_offset = nnl; /* leaf_node_min_index - 1 */
  (void)n;
  /* v1 :207-210 */
  PAR_FOR
  for (long k = 0; k < nnl * ny; ++k) v[k] = z[zo[4] + k];
  /* v2 :214-222 */
  PAR_FOR
  for (long i = 1; i <= nnl; ++i) {
    double dot_p = 0.0;
    for (int k = 0; k < ny; ++k) dot_p += m->b[(m->n_rms > 1 ? (i - 1) * ny : 0) + k] * z[zo[4] + (i - 1) * ny + k];
    v[vo[1] + i - 1] = z[zo[2] + i - 1] - dot_p;
  }
  /* v3 :226-235, v4 :239-248, v5 :253-260, v6 :265-272 */
  PAR_FOR
  for (long i = 1; i <= nnl; ++i) {
    const long c0 = m->child_first[i - 1];
    for (long j = c0; j < c0 + d; ++j) {
      const long j_ind = j - 1;
      cost_mul(m, sqrtQ_of(m, j_ind), nx, z + zo[0] + (i - 1) * nx, v + vo[2] + (j_ind - 1) * nx);
      cost_mul(m, sqrtR_of(m, j_ind), nu, z + zo[1] + (i - 1) * nu, v + vo[3] + (j_ind - 1) * nu);
      v[vo[4] + j_ind - 1] = 0.5 * z[zo[3] + j_ind - 1];
      v[vo[5] + j_ind - 1] = 0.5 * z[zo[3] + j_ind - 1];
    }
  }
  /* v7 :275-282, constraints.jl:111-127 */
  PAR_FOR
  for (long i = 1; i <= nnl; ++i) {
    double *dst = v + vo[6] + (i - 1) * (nx + nu);
    for (int k = 0; k < nx; ++k) dst[k] = z[zo[0] + (i - 1) * nx + k];
    for (int k = 0; k < nu; ++k) dst[nx + k] = z[zo[1] + (i - 1) * nu + k];
  }
  /* v11 :288-294, v12 :298-301, v13 :305-308, v14 :311-317 */
  PAR_FOR
  for (long i = 1; i <= nl; ++i) {
    const long j = i + leaf_offset;
    cost_mul(m, sqrtQN_of(m, i), nx, z + zo[0] + (j - 1) * nx, v + vo[7] + (i - 1) * nx);
    v[vo[8] + i - 1] = 0.5 * z[zo[2] + j - 1];
    v[vo[9] + i - 1] = 0.5 * z[zo[2] + j - 1];
    for (int k = 0; k < nx; ++k) v[vo[10] + (i - 1) * nx + k] = z[zo[0] + (j - 1) * nx + k];
  }
}

/* implicit_l.jl:320-449 */
void spo_Lt(spo_model *m, double *z, const double *v) {
  const int nx = m->nx, nu = m->nu, ny = m->ny, d = m->d;
  const long nnl = m->nnl, nl = m->nl;
  const long *zo = m->zo, *vo = m->vo;
  const long leaf_offset = nnl;
  PAR_FOR
  for (long i = 1; i <= nnl; ++i) {
    double tmpx[nx > nu ? nx : nu];
    const double *src = v + vo[6] + (i - 1) * (nx + nu);
    /* x = Gamma' v7 :351-356 then += sqrtQ v3 :365-373 */
    double *xi = z + zo[0] + (i - 1) * nx;
    double *ui = z + zo[1] + (i - 1) * nu;
    for (int k = 0; k < nx; ++k) xi[k] = src[k];
    const long c0 = m->child_first[i - 1];
    for (long j = c0; j < c0 + d; ++j) {
      const long j_ind = j - 1;
      cost_mul(m, sqrtQ_of(m, j_ind), nx, v + vo[2] + (j_ind - 1) * nx, tmpx);
      for (int k = 0; k < nx; ++k) xi[k] += tmpx[k];
    }
    /* u = Gamma' v7 :399-404 then += sqrtR v4 :412-420 */
    for (int k = 0; k < nu; ++k) ui[k] = src[nx + k];
    for (long j = c0; j < c0 + d; ++j) {
      const long j_ind = j - 1;
      cost_mul(m, sqrtR_of(m, j_ind), nu, v + vo[3] + (j_ind - 1) * nu, tmpx);
      for (int k = 0; k < nu; ++k) ui[k] += tmpx[k];
    }
    /* y :423-429 */
    for (int k = 0; k < ny; ++k) z[zo[4] + (i - 1) * ny + k] = v[(i - 1) * ny + k] - m->b[(m->n_rms > 1 ? (i - 1) * ny : 0) + k] * v[vo[1] + i - 1];
    /* tau :432-437 */
    for (long j = c0; j < c0 + d; ++j) {
      const long j_ind = j - 1;
      z[zo[3] + j_ind - 1] = 0.5 * v[vo[4] + j_ind - 1] + 0.5 * v[vo[5] + j_ind - 1];
    }
    /* s non-leaf :440-442 */
    z[zo[2] + i - 1] = v[vo[1] + i - 1];
  }
  PAR_FOR
  for (long i = 1; i <= nl; ++i) {
    double tmpx[nx];
    const long j = leaf_offset + i;
    /* x leaf = v14 :376-381 then += sqrtQN v11 :390-396 */
    double *xj = z + zo[0] + (j - 1) * nx;
    for (int k = 0; k < nx; ++k) xj[k] = v[vo[10] + (i - 1) * nx + k];
    cost_mul(m, sqrtQN_of(m, i), nx, v + vo[7] + (i - 1) * nx, tmpx);
    for (int k = 0; k < nx; ++k) xj[k] += tmpx[k];
    /* s leaf :445-448 */
    z[zo[2] + j - 1] = 0.5 * v[vo[8] + i - 1] + 0.5 * v[vo[9] + i - 1];
  }
}

/* implicit_l.jl:451-481 */
void spo_mul(spo_model *m, double *C, int trans, const double *B, double alpha, double beta) {
  if (!trans) {
    memcpy(m->mul_buf_v, C, sizeof(double) * m->nv);
    spo_L(m, B, C);
    PAR_FOR
    for (long i = 0; i < m->nv; ++i) C[i] = alpha * C[i] + beta * m->mul_buf_v[i];
  } else {
    memcpy(m->mul_buf_z, C, sizeof(double) * m->nz);
    spo_Lt(m, C, B);
    PAR_FOR
    for (long i = 0; i < m->nz; ++i) C[i] = alpha * C[i] + beta * m->mul_buf_z[i];
  }
}

/* implicit_l.jl:485-516 */
double spo_dot(spo_model *m, const double *a_z, const double *a_v, const double *b_z, const double *b_v, double gamma,
               double sigma) {
  spo_L(m, b_z, m->ws_Lv);
  spo_Lt(m, m->ws_Lz, b_v);
  for (long k = 0; k < m->nz; ++k) m->ws_Lz[k] = b_z[k] - gamma * m->ws_Lz[k];
  for (long k = 0; k < m->nv; ++k) m->ws_Lv[k] = -sigma * m->ws_Lv[k] + b_v[k];
  double res = 0.0;
  for (long k = 0; k < m->nz; ++k) res += a_z[k] * m->ws_Lz[k];
  for (long k = 0; k < m->nv; ++k) res += a_v[k] * m->ws_Lv[k];
  return res;
}

/* ------------------------------------------------------------------------------------------ prox_f */

/* one backward node step, implicit_l.jl:584-637 body */
static void s1_backward_node(spo_model *m, double *z1, long i) {
  const int nx = m->nx, nu = m->nu, d = m->d;
  const long u_offset = m->n * (long)nx;
  double *q = m->ric_q, *dd = m->ric_d;
  double sum_for_d[nu], wu[nu], wx[nx], wx2[nx];
  const long c0 = m->child_first[i - 1];
  for (int k = 0; k < nu; ++k) sum_for_d[k] = 0.0;
  for (long j = c0; j < c0 + d; ++j) { /* :590-596 */
    const double *Bw = m->B + (long)(m->w[j - 1] - 1) * nx * nu;
    gemv_t(Bw, nx, nu, q + (j - 1) * nx, wu);
    for (int k = 0; k < nu; ++k) sum_for_d[k] += wu[k];
  }
  for (int k = 0; k < nu; ++k) wu[k] = z1[u_offset + (i - 1) * nu + k] - sum_for_d[k]; /* :598-600 */
  chol_solve(m->Rc + (i - 1) * (long)nu * nu, nu, wu);                                 /* :601 */
  for (int k = 0; k < nu; ++k) dd[(i - 1) * nu + k] = wu[k];
  for (int k = 0; k < nx; ++k) q[(i - 1) * nx + k] = 0.0; /* :607-609 */
  for (long j = c0; j < c0 + d; ++j) {                    /* :610-624 */
    const double *Bw = m->B + (long)(m->w[j - 1] - 1) * nx * nu;
    gemv_n(Bw, nx, nu, dd + (i - 1) * nu, wx);
    gemv_n(m->P + (j - 1) * (long)nx * nx, nx, nx, wx, wx2);
    for (int k = 0; k < nx; ++k) wx2[k] += q[(j - 1) * nx + k];
    gemv_t(m->ABK + (j - 1) * (long)nx * nx, nx, nx, wx2, wx);
    for (int k = 0; k < nx; ++k) q[(i - 1) * nx + k] += wx[k];
  }
  for (int k = 0; k < nu; ++k) wu[k] = dd[(i - 1) * nu + k] - z1[u_offset + (i - 1) * nu + k]; /* :626-628 */
  gemv_t(m->K + (i - 1) * (long)nu * nx, nu, nx, wu, wx);                                       /* :629 K' */
  for (int k = 0; k < nx; ++k) q[(i - 1) * nx + k] += wx[k];
  for (int k = 0; k < nx; ++k) q[(i - 1) * nx + k] -= z1[(i - 1) * nx + k]; /* :634-636 */
}

/* one forward node step, implicit_l.jl:669-681 body */
static void s1_forward_node(spo_model *m, double *z1, long i) {
  const int nx = m->nx, nu = m->nu, d = m->d;
  const long u_offset = m->n * (long)nx;
  double *dd = m->ric_d;
  double wx[nx];
  gemv_n(m->K + (i - 1) * (long)nu * nx, nu, nx, z1 + (i - 1) * nx, z1 + u_offset + (i - 1) * nu);
  for (int k = 0; k < nu; ++k) z1[u_offset + (i - 1) * nu + k] += dd[(i - 1) * nu + k];
  const long c0 = m->child_first[i - 1];
  for (long j = c0; j < c0 + d; ++j) {
    const double *Bw = m->B + (long)(m->w[j - 1] - 1) * nx * nu;
    gemv_n(m->ABK + (j - 1) * (long)nx * nx, nx, nx, z1 + (i - 1) * nx, z1 + (j - 1) * nx);
    gemv_n(Bw, nx, nu, dd + (i - 1) * nu, wx);
    for (int k = 0; k < nx; ++k) z1[(j - 1) * nx + k] += wx[k];
  }
}

/* implicit_l.jl:559-686.  The reference loops i = nnl..1 and 1..nnl; the per-stage loops below visit the
 * nodes in exactly that order when single-threaded and only reorder independent nodes otherwise. */
void spo_projection_S1(spo_model *m, double *z1) {
  const int nx = m->nx;
  double *q = m->ric_q;
  PAR_FOR
  for (long i = m->nnl + 1; i <= m->n; ++i) /* :578-582 */
    for (int k = 0; k < nx; ++k) q[(i - 1) * nx + k] = -z1[(i - 1) * nx + k];
  for (int t = m->N - 2; t >= 0; --t) {
    const long i0 = m->min_idx[t], i1 = m->min_idx[t + 1];
    PAR_FOR
    for (long i = i1 - 1; i >= i0; --i) s1_backward_node(m, z1, i);
  }
  for (int k = 0; k < nx; ++k) z1[k] = m->x0[k]; /* :667 */
  for (int t = 0; t <= m->N - 2; ++t) {
    const long i0 = m->min_idx[t], i1 = m->min_idx[t + 1];
    PAR_FOR
    for (long i = i0; i < i1; ++i) s1_forward_node(m, z1, i);
  }
}

/* implicit_l.jl:688-727; z2 = view(arg, s_inds[1]+1 : nz) */
void spo_projection_S2(spo_model *m, double *z2) {
  const int ny = m->ny, d = m->d, nls = ny + 2 * d;
  const long n = m->n;
  const long tau_offset = n - 1, y_offset = 2 * (n - 1);
  PAR_FOR
  for (long i = 1; i <= m->nnl; ++i) {
    double ls_b[nls], ls_b2[nls];
    const long ch1 = m->child_first[i - 1];
    /* Julia copyto!(dest, do, src, so, N) with 1-based so: src[so .. so+N-1] */
    for (int k = 0; k < ny; ++k) ls_b[k] = z2[y_offset + (i - 1) * ny + k];          /* :712 */
    for (int k = 0; k < d; ++k) ls_b[ny + k] = z2[(ch1 - 1) - 1 + k];                /* :713 */
    for (int k = 0; k < d; ++k) ls_b[ny + d + k] = z2[tau_offset + (ch1 - 1) - 1 + k]; /* :714 */
    gemv_n(m->ls + (m->n_rms > 1 ? (i - 1) * (long)nls * nls : 0), nls, nls, ls_b, ls_b2);                                            /* :717 */
    for (int k = 0; k < ny; ++k) z2[y_offset + (i - 1) * ny + k] = ls_b2[k];         /* :721 */
    for (int k = 0; k < d; ++k) z2[(ch1 - 1) - 1 + k] = ls_b2[ny + k];               /* :723 */
    for (int k = 0; k < d; ++k) z2[tau_offset + (ch1 - 1) - 1 + k] = ls_b2[ny + d + k]; /* :725 */
  }
}

/* implicit_l.jl:729-748 */
void spo_prox_f(spo_model *m, double *arg, double gamma) {
  arg[m->zo[2]] -= gamma;                   /* :741 s_inds[1] */
  spo_projection_S1(m, arg);                /* :744 view(arg, 1 : s_inds[1]-1) */
  spo_projection_S2(m, arg + m->zo[2] + 1); /* :747 view(arg, s_inds[1]+1 : nz) */
}

/* ------------------------------------------------------------------------------------------ prox_h* */

/* risk_measures.jl:40-71, MOI order (t, x) */
void spo_project_soc(double *vec, long len) {
  const double t = vec[0];
  double x_norm = 0.0;
  for (long i = 1; i < len; ++i) x_norm += vec[i] * vec[i];
  x_norm = sqrt(x_norm);
  if (x_norm <= t) {
  } else if (x_norm <= -t) {
    for (long k = 0; k < len; ++k) vec[k] = 0.0;
  } else {
    vec[0] = (t + x_norm) / 2.0;
    for (long k = 1; k < len; ++k) vec[k] = vec[0] * vec[k] / x_norm;
  }
}

/* risk_measures.jl:14-38 applied to MOI.dual_set(cone) as in implicit_l.jl:841 */
static void project_dual_cone(double *vec, int type, int dim) {
  switch (type) {
    case SPO_CONE_ZEROS: /* dual = Reals: nothing */
      break;
    case SPO_CONE_NONNEG: /* self-dual */
      for (int k = 0; k < dim; ++k)
        if (vec[k] < 0.0) vec[k] = 0.0;
      break;
    case SPO_CONE_NONPOS:
      for (int k = 0; k < dim; ++k)
        if (vec[k] > 0.0) vec[k] = 0.0;
      break;
    case SPO_CONE_REALS: /* dual = Zeros */
      for (int k = 0; k < dim; ++k) vec[k] = 0.0;
      break;
    case SPO_CONE_SOC:
      spo_project_soc(vec, dim);
      break;
  }
}

static inline double clip(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); } /* constraints.jl:28-30 */

/* implicit_l.jl:752-810 */
static void project_on_leaf_constraints(spo_model *m, double *arg) {
  const int nx = m->nx;
  const long *vo = m->vo;
  PAR_FOR
  for (long i = 1; i <= m->nl; ++i) {
    double ws[nx + 2];
    ws[0] = arg[vo[9] + i - 1]; /* v13 */
    for (int j = 0; j < nx; ++j) ws[j + 1] = arg[vo[7] + (i - 1) * nx + j];
    ws[nx + 1] = arg[vo[8] + i - 1]; /* v12 */
    spo_project_soc(ws, nx + 2);
    arg[vo[9] + i - 1] = ws[0];
    for (int j = 0; j < nx; ++j) arg[vo[7] + (i - 1) * nx + j] = ws[j + 1];
    arg[vo[8] + i - 1] = ws[nx + 1];
    for (int j = 0; j < nx; ++j) /* v14 box :804-809, constraints.jl:24-33 */
      arg[vo[10] + (i - 1) * nx + j] = clip(arg[vo[10] + (i - 1) * nx + j], m->x_min, m->x_max);
  }
}

/* implicit_l.jl:812-909 */
static void project_on_nonleaf_constraints(spo_model *m, double *arg) {
  const int nx = m->nx, nu = m->nu, ny = m->ny, d = m->d;
  const long *vo = m->vo;
  PAR_FOR
  for (long i = 1; i <= m->nnl; ++i) {
    double ws[nx + nu + 2];
    long offset = (i - 1) * ny; /* v1 :833-845 */
    for (int k = 0; k < m->n_cones; ++k) {
      project_dual_cone(arg + offset, m->cone_type[k], m->cone_dim[k]);
      offset += m->cone_dim[k];
    }
    arg[vo[1] + i - 1] = arg[vo[1] + i - 1] >= 0 ? arg[vo[1] + i - 1] : 0.0; /* v2 :849-851 */
    const long c0 = m->child_first[i - 1];
    for (long j = c0; j < c0 + d; ++j) { /* v3-v6 :854-900 */
      const long j_ind = j - 1;
      ws[0] = arg[vo[5] + j_ind - 1];
      for (int k = 0; k < nx; ++k) ws[k + 1] = arg[vo[2] + (j_ind - 1) * nx + k];
      for (int k = 0; k < nu; ++k) ws[1 + nx + k] = arg[vo[3] + (j_ind - 1) * nu + k];
      ws[1 + nx + nu] = arg[vo[4] + j_ind - 1];
      spo_project_soc(ws, nx + nu + 2);
      arg[vo[5] + j_ind - 1] = ws[0];
      for (int k = 0; k < nx; ++k) arg[vo[2] + (j_ind - 1) * nx + k] = ws[k + 1];
      for (int k = 0; k < nu; ++k) arg[vo[3] + (j_ind - 1) * nu + k] = ws[1 + nx + k];
      arg[vo[4] + j_ind - 1] = ws[1 + nx + nu];
    }
    double *xu = arg + vo[6] + (i - 1) * (nx + nu); /* v7 :903-908, constraints.jl:38-55 */
    for (int k = 0; k < nx; ++k) xu[k] = clip(xu[k], m->x_min, m->x_max);
    for (int k = 0; k < nu; ++k) xu[nx + k] = clip(xu[nx + k], m->u_min, m->u_max);
  }
}

/* implicit_l.jl:911-951 */
void spo_prox_h_conj(spo_model *m, double *arg, double sigma) {
  const long *vo = m->vo;
  const long n = m->n, nl = m->nl;
  PAR_FOR
  for (long i = 0; i < m->nv; ++i) arg[i] = arg[i] / sigma;
  for (long i = 0; i < n - 1; ++i) arg[vo[4] + i] -= 0.5; /* v5 */
  for (long i = 0; i < n - 1; ++i) arg[vo[5] + i] += 0.5; /* v6 */
  for (long i = 0; i < nl; ++i) arg[vo[8] + i] -= 0.5;    /* v12 */
  for (long i = 0; i < nl; ++i) arg[vo[9] + i] += 0.5;    /* v13 */
  memcpy(m->prox_v_ws, arg, sizeof(double) * m->nv);
  project_on_leaf_constraints(m, arg);
  project_on_nonleaf_constraints(m, arg);
  PAR_FOR
  for (long i = 0; i < m->nv; ++i) arg[i] = sigma * (m->prox_v_ws[i] - arg[i]);
}

/* ------------------------------------------------------------------------------------------ CP / SP */

static double norm_inf(const double *x, long k) {
  double r = 0.0;
  for (long i = 0; i < k; ++i) {
    double a = fabs(x[i]);
    if (a > r || a != a) r = a; /* LinearAlgebra.norm(x, Inf) propagates NaN (r stays NaN once set) */
  }
  return r;
}
static double norm2(const double *x, long k) {
  double r = 0.0;
  for (long i = 0; i < k; ++i) r += x[i] * x[i];
  return sqrt(r);
}

/* T: (z,v) -> (zbar,vbar); cp.jl:5-32 == sp.jl:5-32 == sp.jl:133-153 */
static void operator_T(spo_model *m, const double *z, const double *v, double *zbar, double *vbar, double *zws,
                       double gamma, double sigma) {
  memcpy(zbar, z, sizeof(double) * m->nz);
  spo_mul(m, zbar, 1, v, -gamma, 1.0);
  spo_prox_f(m, zbar, gamma);
  memcpy(zws, z, sizeof(double) * m->nz);
  PAR_FOR
  for (long i = 0; i < m->nz; ++i) zws[i] = 2.0 * zbar[i] + (-1.0) * zws[i]; /* axpby! */
  memcpy(vbar, v, sizeof(double) * m->nv);
  spo_mul(m, vbar, 0, zws, sigma, 1.0);
  spo_prox_h_conj(m, vbar, sigma);
}

/* cp.jl:54-123 / sp.jl:270-344; returns 1 when the loop must break */
static int should_terminate(spo_model *m, double a1, double a2, double tol, int verbose_log, double *xi_out,
                            double *xi1_out, double *xi2_out) {
  const long nz = m->nz, nv = m->nv;
  PAR_FOR
  for (long i = 0; i < nz; ++i) m->delta_z[i] = m->z[i] - m->z_old[i];
  PAR_FOR
  for (long i = 0; i < nv; ++i) m->delta_v[i] = m->v[i] - m->v_old[i];
  memcpy(m->xi_1, m->delta_z, sizeof(double) * nz);
  memcpy(m->xi_2, m->delta_v, sizeof(double) * nv);
  spo_mul(m, m->xi_1, 1, m->delta_v, 1.0, -1.0 / a1);
  spo_mul(m, m->xi_2, 0, m->delta_z, 1.0, -1.0 / a2);
  const double xi_1 = norm_inf(m->xi_1, nz), xi_2 = norm_inf(m->xi_2, nv);
  double xi = NAN;
  int res = 0;
  if (verbose_log) {
    if (m->alg == SPO_ALG_CP) { /* cp.jl:82-86 */
      memcpy(m->xi, m->xi_1, sizeof(double) * nz);
      spo_mul(m, m->xi, 1, m->xi_2, 1.0, 1.0);
      xi = norm_inf(m->xi, nz);
    } else { /* sp.jl:304 */
      xi = xi_1 > xi_2 ? xi_1 : xi_2;
    }
    if (xi <= tol) res = 1;
  } else {
    const double t1 = tol * m->res_0[0] > tol ? tol * m->res_0[0] : tol;
    const double t2 = tol * m->res_0[1] > tol ? tol * m->res_0[1] : tol;
    if (xi_1 <= t1 && xi_2 <= t2) res = 1;
  }
  if (m->res_0[0] == -INFINITY) m->res_0[0] = xi_1;
  if (m->res_0[1] == -INFINITY) m->res_0[1] = xi_2;
  *xi_out = xi; *xi1_out = xi_1; *xi2_out = xi_2;
  return res;
}

/* anderson.jl:1-123 */
static void anderson(spo_model *m, int iter) {
  const long nz = m->nz, nv = m->nv, nr = nz + nv;
  double *MP = m->MP, *MR = m->MR, *Q = m->aaQ, *R = m->aaR, *g = m->aa_gamma;
  for (int j = AA_M - 1; j >= 1; --j) /* :13-24 shift MP, MR right */
    memcpy(MP + (long)j * nr, MP + (long)(j - 1) * nr, sizeof(double) * nr);
  for (int j = AA_M - 1; j >= 1; --j) memcpy(MR + (long)j * nr, MR + (long)(j - 1) * nr, sizeof(double) * nr);
  for (long i = 0; i < nz; ++i) { /* :26-33 */
    MR[i] = m->delta_rz[i];
    MP[i] = m->delta_z[i] - MR[i];
  }
  for (long i = 0; i < nv; ++i) {
    MR[nz + i] = m->delta_rv[i];
    MP[nz + i] = m->delta_v[i] - MR[nz + i];
  }
  memcpy(m->aa_wsp, m->rz, sizeof(double) * nz); /* :35-40 */
  memcpy(m->aa_wsp + nz, m->rv, sizeof(double) * nv);
  /* :46-50 diagonal shift of the UpperTriangular R; writes below the diagonal are structural zeros */
  for (int j = AA_M - 1; j >= 1; --j)
    for (int i = AA_M - 1; i >= 1; --i)
      if (i <= j) R[i + j * AA_M] = R[(i - 1) + (j - 1) * AA_M];
  for (int j = AA_M - 1; j >= 1; --j) memcpy(Q + (long)j * nr, Q + (long)(j - 1) * nr, sizeof(double) * nr); /* :52-56 */
  memcpy(Q, MR, sizeof(double) * nr); /* :63-65 */
  if (iter == 1) {                    /* :68-76 */
    double r = norm2(Q, nr);
    for (long i = 0; i < nr; ++i) Q[i] /= r;
    R[0] = r;
  } else {
    const int jmax = iter < AA_M ? iter : AA_M;
    for (int j = 2; j <= jmax; ++j) { /* :78-91 */
      const double *Qj = Q + (long)(j - 1) * nr;
      double r = 0.0;
      for (long i = 0; i < nr; ++i) r += Qj[i] * Q[i];
      r /= norm2(Qj, nr);
      R[0 + (j - 1) * AA_M] = r;
      for (long i = 0; i < nr; ++i) Q[i] = Q[i] - r * Qj[i];
    }
    double r = norm2(Q, nr); /* :93-98 */
    for (long i = 0; i < nr; ++i) Q[i] /= r;
    R[0] = r;
  }
  for (int j = 0; j < AA_M; ++j) { /* :111 gamma = Q' wsp */
    const double *Qj = Q + (long)j * nr;
    double acc = 0.0;
    for (long i = 0; i < nr; ++i) acc += Qj[i] * m->aa_wsp[i];
    g[j] = acc;
  }
  for (int j = AA_M - 1; j >= 0; --j) { /* :113 ldiv!(UpperTriangular R, gamma), column-oriented back substitution */
    g[j] = g[j] / R[j + j * AA_M];
    for (int i = 0; i < j; ++i) g[i] -= R[i + j * AA_M] * g[j];
  }
  for (long i = 0; i < nr; ++i) m->aa_wsp[i] = 0.0; /* :116 wsp = MP * gamma */
  for (int j = 0; j < AA_M; ++j) {
    const double *MPj = MP + (long)j * nr;
    const double gj = g[j];
    for (long i = 0; i < nr; ++i) m->aa_wsp[i] += MPj[i] * gj;
  }
  for (long i = 0; i < nz; ++i) m->dz[i] = -m->rz[i] - m->aa_wsp[i]; /* :117-122 */
  for (long i = 0; i < nv; ++i) m->dv[i] = -m->rv[i] - m->aa_wsp[nz + i];
}

typedef struct {
  double r_safe;
  int iter, broyden_k;
  int k1, k2, fb;
  double eta; /* sp.jl:381 eta = r_safe */
  int k0;
} sp_locals;

/* one pass of cp.jl:205-228; returns should_terminate! */
static int cp_pass(spo_model *m, double gamma, double sigma, double tol, int verbose_log, double *xi, double *xi1,
                   double *xi2) {
  memcpy(m->z_old, m->z, sizeof(double) * m->nz); /* :207-208 */
  memcpy(m->v_old, m->v, sizeof(double) * m->nv);
  operator_T(m, m->z, m->v, m->zbar, m->vbar, m->z_workspace, gamma, sigma); /* :210-211 */
  /* update_z!/update_v! with lambda = 1 (:213-214): lambda*zbar + (1-lambda)*z */
  for (long i = 0; i < m->nz; ++i) m->z[i] = 1.0 * m->zbar[i] + (1.0 - 1.0) * m->z[i];
  for (long i = 0; i < m->nv; ++i) m->v[i] = 1.0 * m->vbar[i] + (1.0 - 1.0) * m->v[i];
  return should_terminate(m, gamma, sigma, tol, verbose_log, xi, xi1, xi2);
}

/* one pass of sp.jl:388-463; returns should_terminate!, *bc = backtrack_count */
static int sp_pass(spo_model *m, sp_locals *L, double gamma, double sigma, double tol, int verbose_log, double *xi,
                   double *xi1, double *xi2, int *bc) {
  const long nz = m->nz, nv = m->nv;
  const double c1 = 0.99, q = 0.99, sigma_k2 = 0.1, lambda_sp = 1.0; /* sp.jl:366-372 */
  const int MAX_BACKTRACK_COUNT = 8;
  operator_T(m, m->z, m->v, m->zbar, m->vbar, m->z_workspace, gamma, sigma); /* :389-390 */
  for (long i = 0; i < nz; ++i) m->rz[i] = m->z[i] - m->zbar[i];             /* :34-51 */
  for (long i = 0; i < nv; ++i) m->rv[i] = m->v[i] - m->vbar[i];
  const double rnorm = sqrt(spo_dot(m, m->rz, m->rv, m->rz, m->rv, gamma, sigma));
  for (long i = 0; i < nz; ++i) m->delta_rz[i] = m->rz[i] - m->rz_old[i]; /* :257-268 */
  for (long i = 0; i < nv; ++i) m->delta_rv[i] = m->rv[i] - m->rv_old[i];
  memcpy(m->z_old, m->z, sizeof(double) * nz); /* :232-255 */
  memcpy(m->v_old, m->v, sizeof(double) * nv);
  memcpy(m->rz_old, m->rz, sizeof(double) * nz);
  memcpy(m->rv_old, m->rv, sizeof(double) * nv);
  L->broyden_k += 1; /* spock.jl:212-224 */
  anderson(m, L->broyden_k);
  double tau = 1.0, rtilde_norm = INFINITY;
  int backtrack_count = 0, loop = 1;
  const double c0 = 0.99, beta = 0.5; /* sp.jl:364,369 */
  /* K0 is disabled in the reference: should_perform_k0 ends in "&& false" (sp.jl:80).  SPO_SP_K0 restores the test
   * as written before the "&& false": the blind update of perform_k0! (:83-106). */
  if ((m->sp_options & SPO_SP_K0) && rnorm <= c0 * L->eta) {
    L->eta = rnorm;
    for (long i = 0; i < nz; ++i) m->z[i] += m->dz[i];
    for (long i = 0; i < nv; ++i) m->v[i] += m->dv[i];
    loop = 0; /* (update_candidate_residual! :406 only computes rtilde_norm, which nobody reads afterwards) */
    L->k0++;
  }
  while (loop && backtrack_count <= MAX_BACKTRACK_COUNT) { /* :411-441 */
    for (long i = 0; i < nz; ++i) m->sw[i] = m->z[i] + tau * m->dz[i];
    for (long i = 0; i < nv; ++i) m->su[i] = m->v[i] + tau * m->dv[i];
    operator_T(m, m->sw, m->su, m->wbar, m->ubar, m->w_workspace, gamma, sigma); /* :133-153 */
    for (long i = 0; i < nz; ++i) m->rw[i] = m->sw[i] - m->wbar[i];
    for (long i = 0; i < nv; ++i) m->ru[i] = m->su[i] - m->ubar[i];
    rtilde_norm = sqrt(spo_dot(m, m->rw, m->ru, m->rw, m->ru, gamma, sigma));
    if (rnorm <= L->r_safe && rtilde_norm <= c1 * rnorm) { /* K1 :166-191 */
      memcpy(m->z, m->sw, sizeof(double) * nz);
      memcpy(m->v, m->su, sizeof(double) * nv);
      L->r_safe = rtilde_norm + pow(q, (double)L->iter);
      loop = 0;
      L->k1++;
      continue;
    }
    for (long k = 0; k < nz; ++k) m->ws_rho_z[k] = m->rw[k] - tau * m->dz[k]; /* :419-424 */
    for (long k = 0; k < nv; ++k) m->ws_rho_v[k] = m->ru[k] - tau * m->dv[k];
    double rho = spo_dot(m, m->rw, m->ru, m->ws_rho_z, m->ws_rho_v, gamma, sigma);
    if (rho >= sigma_k2 * rnorm * rtilde_norm) { /* K2 :193-222 */
      rho /= rtilde_norm * rtilde_norm;
      for (long i = 0; i < nz; ++i) m->z[i] -= lambda_sp * rho * m->rw[i];
      for (long i = 0; i < nv; ++i) m->v[i] -= lambda_sp * rho * m->ru[i];
      loop = 0;
      L->k2++;
      continue;
    }
    /* perform_linesearch!(tau, beta) returns tau*beta but the result is discarded (sp.jl:439);
     * SPO_SP_LINESEARCH assigns it (the evident intent: tau = perform_linesearch!(tau, beta)) */
    if (m->sp_options & SPO_SP_LINESEARCH) tau = tau * beta;
    backtrack_count += 1;
  }
  if (loop) { /* :443-446 plain CP update */
    for (long i = 0; i < nz; ++i) m->z[i] = 1.0 * m->zbar[i] + (1.0 - 1.0) * m->z[i];
    for (long i = 0; i < nv; ++i) m->v[i] = 1.0 * m->vbar[i] + (1.0 - 1.0) * m->v[i];
    L->fb++;
  }
  /* update_delta_old! (:346-356) only copies into buffers nobody reads */
  *bc = backtrack_count;
  return should_terminate(m, gamma, sigma, tol, verbose_log, xi, xi1, xi2);
}

/* cpock.jl:169-189 + cp.jl:188-232 / spock.jl:231-251 + sp.jl:358-469 */
int spo_solve(spo_model *m, const double *x0, double tol, int verbose_log, const double *z0, const double *v0,
              double gamma, double sigma, int max_iter, spo_result *res, double *log_xi, double *log_xi1,
              double *log_xi2, int *log_bc, int log_cap) {
  if (z0 != NULL || v0 != NULL) { /* cpock.jl:180-183: logical-or guard, both are then required */
    if (z0 == NULL || v0 == NULL) return 1;
    memcpy(m->z, z0, sizeof(double) * m->nz);
    memcpy(m->v, v0, sizeof(double) * m->nv);
  }
  memcpy(m->x0, x0, sizeof(double) * m->nx);
  m->res_0[0] = -INFINITY; m->res_0[1] = -INFINITY;
  if (!(sigma > 0) || !(gamma > 0)) { /* cp.jl:198-201 */
    sigma = 0.99 / sqrt(m->L_norm);
    gamma = sigma;
  }
  if (max_iter <= 0) max_iter = (m->alg == SPO_ALG_SP) ? 1000 : 5000; /* sp.jl:360 / cp.jl:190 */
  sp_locals L = {INFINITY, 0, 0, 0, 0, 0, INFINITY, 0};
  int iter = 0, passes = 0, converged = 0;
  double xi = NAN, xi1 = NAN, xi2 = NAN;
  while (iter < max_iter) {
    int bc = 0, stop;
    L.iter = iter;
    if (m->alg == SPO_ALG_SP)
      stop = sp_pass(m, &L, gamma, sigma, tol, verbose_log, &xi, &xi1, &xi2, &bc);
    else
      stop = cp_pass(m, gamma, sigma, tol, verbose_log, &xi, &xi1, &xi2);
    if (passes < log_cap) {
      if (log_xi) log_xi[passes] = xi;
      if (log_xi1) log_xi1[passes] = xi1;
      if (log_xi2) log_xi2[passes] = xi2;
      if (log_bc) log_bc[passes] = bc;
    }
    passes++;
    if (stop) {
      converged = 1;
      break;
    }
    iter += 1;
  }
  if (res) {
    res->passes = passes; res->printed_iter = iter; res->converged = converged;
    res->xi = xi; res->xi1 = xi1; res->xi2 = xi2;
    res->k1_count = L.k1; res->k2_count = L.k2; res->fallback_count = L.fb;
  }
  return 0;
}

void spo_set_sp_options(spo_model *m, int options) { m->sp_options = options; }

int spo_step(spo_model *m, int K, double gamma, double sigma) {
  if (!(sigma > 0) || !(gamma > 0)) {
    sigma = 0.99 / sqrt(m->L_norm);
    gamma = sigma;
  }
  static sp_locals L; /* parity helper: SP locals restart like a fresh run_sp! on every call */
  L = (sp_locals){INFINITY, 0, 0, 0, 0, 0, INFINITY, 0};
  double xi, xi1, xi2;
  int bc;
  for (int k = 0; k < K; ++k) {
    L.iter = k;
    if (m->alg == SPO_ALG_SP)
      sp_pass(m, &L, gamma, sigma, 0.0, 0, &xi, &xi1, &xi2, &bc);
    else
      cp_pass(m, gamma, sigma, 0.0, 0, &xi, &xi1, &xi2);
  }
  return 0;
}
