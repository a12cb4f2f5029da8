// spk_sp.cuh -- device-resident control of the Chambolle-Pock / SuperMann loops.
//
// The reference's loops (cp.jl:188-232, sp.jl:358-469, anderson.jl:1-123) interleave vector sweeps with scalar
// decisions: termination test, Anderson's Gram-Schmidt coefficients and 3x3 triangular solve, the K1/K2/fallback branch
// of SuperMann.  Driving those from the host costs one stream synchronisation per decision (5 per SuperMann pass),
// which dominates the pass time of every small tree.  Here all scalars live in one control block in device memory
// (SpkCtl); each vector sweep is ONE kernel that also folds the dot products it can and whose last block finishes the
// scalar step ("last block done" pattern: fixed-order folds, so results are deterministic); branch decisions are
// uniform reads of the control block.  A whole pass is a fixed launch sequence (captured as a CUDA graph) and the host
// only polls the `done` flag every few passes; once it is set every kernel returns immediately, so the iterate is
// exactly the one of the pass that met the termination test, as in the reference.
#pragma once
#include "spk_kernels.cuh"

struct SpkCtl {
  // should_terminate! (cp.jl:54-123 / sp.jl:270-344)
  double tol, res0[2], xi1, xi2;
  int done, passes, ignore_done, verbose_log, log_cap, pad0;
  // SuperMann / Anderson (sp.jl:358-469, anderson.jl)
  double R[9], g[3], rho2, rho3, r11inv, rnorm, rtnorm, r_safe, coef;
  int k, it, branch, need_k2, rot, k1, k2, fb;
};

#define SPK_SP_THREADS 256
#if !SPK_EMU
#define SPK_SP_KERNEL __global__ void __launch_bounds__(SPK_SP_THREADS)
#define SPK_LDCG(p) __ldcg(p)
#else  // host emulation (tests only): blocks run one after the other with one thread each
#define SPK_SP_KERNEL static void
#define SPK_LDCG(p) (*(p))
#endif

// Folds NS running sums of every thread of the grid.  Returns true in thread 0 of the LAST block to arrive, with the
// grid totals in tot[]; every other thread gets false.  partial: gridDim.x * NS doubles; ticket: self-resetting counter.
template <int NS>
SPK_D bool spk_grid_fold(double (&acc)[NS], double *partial, unsigned *ticket, double (&tot)[NS]) {
#if SPK_EMU
  (void)ticket;
  for (int j = 0; j < NS; ++j) partial[(size_t)blockIdx.x * NS + j] = acc[j];
  if (blockIdx.x != gridDim.x - 1) return false;  // blocks are emulated in order: the last one folds
  for (int j = 0; j < NS; ++j) {
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += partial[(size_t)b * NS + j];
    tot[j] = v;
  }
  return true;
#else
  __shared__ double s_w[SPK_SP_THREADS / 32][NS];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_w[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double v = 0.0;
      for (int w = 0; w < nw; ++w) v += s_w[w][j];
      partial[(size_t)blockIdx.x * NS + j] = v;
    }
    __threadfence();
    s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // last block: fixed-order fold of the partials (thread-strided, then the same warp / block tree)
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    double v = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) v += SPK_LDCG(partial + (size_t)b * NS + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_w[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return false;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    double v = 0.0;
    for (int w = 0; w < nw; ++w) v += s_w[w][j];
    tot[j] = v;
  }
  return true;
#endif
}

struct SpkSpVecs {
  double *zv, *zvbar, *zv_old, *dzv, *rzv, *rzv_old, *dir, *wu, *wubar, *rwu, *ws_rho;
  double *MP[3], *MR[3], *AQ[3];  // PHYSICAL column slots; logical column j is slot (j + 2 rot) % 3
};
SPK_HD int spk_sp_slot(int j, int rot) { return (j + 2 * rot) % 3; }

// ---- KA (after T(z,v)): update_residual!, update_delta_r!, update_sy! (sp.jl:34-51, :232-268), the new Anderson
//      columns (anderson.jl:13-65, rotating slots instead of shifting) and the first Gram-Schmidt dots (:78-91)
SPK_SP_KERNEL spk_k_sp_a(long nr, SpkSpVecs V, SpkCtl *c, const double *scal, double *partial,
                                                               unsigned *ticket) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int rot = c->rot + 1;  // this pass's rotation; stored by the epilogue after every block has read the old value
  double *mr0 = V.MR[spk_sp_slot(0, rot)], *mp0 = V.MP[spk_sp_slot(0, rot)], *q0 = V.AQ[spk_sp_slot(0, rot)];
  const double *q1 = V.AQ[spk_sp_slot(1, rot)];
  double acc[2] = {0.0, 0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    const double z = V.zv[i], r = z - V.zvbar[i], dr = r - V.rzv_old[i];
    V.rzv[i] = r;
    V.rzv_old[i] = r;
    V.zv_old[i] = z;
    mr0[i] = dr;
    q0[i] = dr;
    mp0[i] = V.dzv[i] - dr;
    const double q = q1[i];
    acc[0] += q * dr;
    acc[1] += q * q;
  }
  double tot[2];
  if (spk_grid_fold<2>(acc, partial, ticket, tot)) {
    c->rot = rot;
    const int k = ++c->k;  // spock.jl:212-224
    c->rnorm = sqrt(scal[4]);
    double *R = c->R;
    for (int j = 2; j >= 1; --j)  // anderson.jl:46-50
      for (int i = 2; i >= 1; --i)
        if (i <= j) R[i + j * 3] = R[(i - 1) + (j - 1) * 3];
    double rho2 = 0.0;
    if (k >= 2) {
      rho2 = tot[0] / sqrt(tot[1]);
      R[0 + 1 * 3] = rho2;
    }
    c->rho2 = rho2;
  }
}
// ---- KB: Q1 -= rho2 Q2 (anderson.jl:78-91, j = 2) and the dots for j = 3
SPK_SP_KERNEL spk_k_sp_b(long nr, SpkSpVecs V, SpkCtl *c, double *partial, unsigned *ticket) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int rot = c->rot;
  double *q0 = V.AQ[spk_sp_slot(0, rot)];
  const double *q1 = V.AQ[spk_sp_slot(1, rot)], *q2 = V.AQ[spk_sp_slot(2, rot)];
  const double rho2 = c->rho2;
  const bool upd = c->k >= 2;
  double acc[2] = {0.0, 0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    double a = q0[i];
    if (upd) {
      a = a + (-rho2) * q1[i];
      q0[i] = a;
    }
    const double q = q2[i];
    acc[0] += q * a;
    acc[1] += q * q;
  }
  double tot[2];
  if (spk_grid_fold<2>(acc, partial, ticket, tot)) {
    double rho3 = 0.0;
    if (c->k >= 3) {
      rho3 = tot[0] / sqrt(tot[1]);
      c->R[0 + 2 * 3] = rho3;
    }
    c->rho3 = rho3;
  }
}
// ---- KC: Q1 -= rho3 Q3, ||Q1|| (anderson.jl:68-76 / :93-98)
SPK_SP_KERNEL spk_k_sp_c(long nr, SpkSpVecs V, SpkCtl *c, double *partial, unsigned *ticket) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int rot = c->rot;
  double *q0 = V.AQ[spk_sp_slot(0, rot)];
  const double *q2 = V.AQ[spk_sp_slot(2, rot)];
  const double rho3 = c->rho3;
  const bool upd = c->k >= 3;
  double acc[1] = {0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    double a = q0[i];
    if (upd) {
      a = a + (-rho3) * q2[i];
      q0[i] = a;
    }
    acc[0] += a * a;
  }
  double tot[1];
  if (spk_grid_fold<1>(acc, partial, ticket, tot)) {
    const double r11 = sqrt(tot[0]);
    c->R[0] = r11;
    c->r11inv = 1.0 / r11;
  }
}
// ---- KD: Q1 /= R11, Q'r (anderson.jl:111), then gamma = R^{-1} Q'r (:113)
SPK_SP_KERNEL spk_k_sp_d(long nr, SpkSpVecs V, SpkCtl *c, double *partial, unsigned *ticket) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int rot = c->rot;
  double *q0 = V.AQ[spk_sp_slot(0, rot)];
  const double *q1 = V.AQ[spk_sp_slot(1, rot)], *q2 = V.AQ[spk_sp_slot(2, rot)];
  const double inv = c->r11inv;
  double acc[3] = {0.0, 0.0, 0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    const double a = inv * q0[i], r = V.rzv[i];
    q0[i] = a;
    acc[0] += a * r;
    acc[1] += q1[i] * r;
    acc[2] += q2[i] * r;
  }
  double tot[3];
  if (spk_grid_fold<3>(acc, partial, ticket, tot)) {
    const double *R = c->R;
    double g[3] = {tot[0], tot[1], tot[2]};
    for (int j = 2; j >= 0; --j) {  // ldiv!(UpperTriangular(R), .)
      g[j] = g[j] / R[j + j * 3];
      for (int i = 0; i < j; ++i) g[i] -= R[i + j * 3] * g[j];
    }
    c->g[0] = g[0]; c->g[1] = g[1]; c->g[2] = g[2];
  }
}
// ---- KE: d = -r - MP gamma (anderson.jl:116-122), candidate (w,u) = (z,v) + d (sp.jl:118-131, tau = 1)
SPK_SP_KERNEL spk_k_sp_e(long nr, SpkSpVecs V, const SpkCtl *c) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int rot = c->rot;
  const double *p0 = V.MP[spk_sp_slot(0, rot)], *p1 = V.MP[spk_sp_slot(1, rot)], *p2 = V.MP[spk_sp_slot(2, rot)];
  const double g0 = c->g[0], g1 = c->g[1], g2 = c->g[2];
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    double d = -1.0 * V.rzv[i] + (-g0) * p0[i];
    d = d + (-g1) * p1[i];
    d = d + (-g2) * p2[i];
    V.dir[i] = d;
    V.wu[i] = V.zv[i] + d;
  }
}
// ---- SF (one thread): the K1 test (sp.jl:166-191)
SPK_KERNEL void spk_k_sp_f_decide(SpkCtl *c, const double *scal) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const double c1 = 0.99, q = 0.99;  // sp.jl:366-372
  const double rt = sqrt(scal[5]);
  c->rtnorm = rt;
  if (c->rnorm <= c->r_safe && rt <= c1 * c->rnorm) {
    c->branch = 1;
    c->need_k2 = 0;
    c->r_safe = rt + pow(q, (double)c->it);
    c->k1++;
  } else {
    c->branch = 0;
    c->need_k2 = 1;
  }
}
// ---- KF: candidate residual (sp.jl:133-164) and, unless K1 holds, r~ - d for the K2 inner product (:419-424)
SPK_SP_KERNEL spk_k_sp_f(long nr, SpkSpVecs V, const SpkCtl *c) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const bool k2 = c->need_k2 != 0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    const double r = V.wu[i] - V.wubar[i];
    V.rwu[i] = r;
    if (k2) V.ws_rho[i] = r - V.dir[i];
  }
}
// ---- SG (one thread): K2 test or fallback (sp.jl:193-222, :443-446)
SPK_KERNEL void spk_k_sp_g_decide(SpkCtl *c, double *scal) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const double sigma_k2 = 0.1, lambda_sp = 1.0;
  if (c->need_k2) {
    const double rho = scal[6], rt = c->rtnorm;
    if (rho >= sigma_k2 * c->rnorm * rt) {
      c->branch = 2;
      c->coef = lambda_sp * (rho / (rt * rt));
      c->k2++;
    } else {
      c->branch = 3;
      c->fb++;
    }
  }
  scal[10] = 0.0;  // the residual maxima of this pass accumulate here
  scal[11] = 0.0;
}
// ---- KG: the accepted update and delta = (z,v) - (z,v)_old for should_terminate! and the next Anderson step
SPK_SP_KERNEL spk_k_sp_g(long nr, SpkSpVecs V, const SpkCtl *c) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const int branch = c->branch;
  const double coef = c->coef;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    double z;
    if (branch == 1) z = V.wu[i];
    else if (branch == 2) z = V.zv[i] + (-coef) * V.rwu[i];
    else z = V.zvbar[i];
    V.zv[i] = z;
    V.dzv[i] = z - V.zv_old[i];
  }
}
// ---- L / L' with device-side gating and fused reductions (spock_mul! :451-481, spock_dot :485-516, should_terminate!):
//   out != NULL: out[e] = alpha op(.)[e] + beta in[e]
//   dot_with != NULL: acc += dot_with[e] * (that value)       (P-inner product of the K2 test; grid fold into scal[slot])
//   absmax != NULL: *absmax = max |that value|                (residual maxima; exact, order-free integer atomicMax)
//   run_if != NULL: the kernel only runs when *run_if != 0
template <bool TRANS>
SPK_SP_KERNEL spk_k_Lop_ctl(SpkDims D, SpkPtrs P, const double *__restrict__ x, double alpha, double beta,
                                                                  const double *__restrict__ in, double *__restrict__ out,
                                                                  const double *__restrict__ dot_with, double *dot_out, int dot_accumulate,
                                                                  double *absmax, const SpkCtl *c, const int *run_if, double *partial,
                                                                  unsigned *ticket) {
  spk_pdl_trigger();
  spk_pdl_wait();
  if (c->done || (run_if && !*run_if)) return;
  const long n = TRANS ? D.nz : D.nv;
  double acc[1] = {0.0}, mx = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const double op = TRANS ? spk_Lt_entry(D, P, x, e) : spk_L_entry(D, P, x, e);
    const double val = (beta == 0.0) ? alpha * op : alpha * op + beta * in[e];
    if (out) out[e] = val;
    if (dot_with) acc[0] += dot_with[e] * val;
    mx = spk_max_nn(mx, fabs(val));
  }
  if (absmax) {
#if SPK_EMU
    *absmax = spk_max_nn(*absmax, mx);
#else
    __shared__ double s_m[SPK_SP_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = spk_max_nn(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = spk_max_nn(mx, s_m[w]);
      if (spk_bits_nn(mx) > spk_bits_nn(__ldcg(absmax))) atomicMax(reinterpret_cast<unsigned long long *>(absmax), (unsigned long long)__double_as_longlong(mx));
    }
#endif
  }
  if (dot_with) {
    double tot[1];
    if (spk_grid_fold<1>(acc, partial, ticket, tot)) *dot_out = dot_accumulate ? *dot_out + tot[0] : tot[0];
  }
}

// ---- SH (one thread): should_terminate! (sp.jl:270-344 / cp.jl:54-123), per-pass log, counters
//      xi_slot: where the pass left (xi1, xi2); bc_from_branch: SuperMann's backtrack_count (9 on fallback)
SPK_KERNEL void spk_k_terminate(SpkCtl *c, const double *scal, int xi_slot, int is_sp, double *log) {
  spk_pdl_trigger();
  spk_pdl_wait();  // chained to the previous launch of the pass by programmatic dependent launch
  if (c->done) return;
  const double xi1 = scal[xi_slot], xi2 = scal[xi_slot + 1];
  bool stop;
  if (c->verbose_log) {  // sp.jl:304: xi = max(xi1, xi2) <= tol
    stop = (xi1 <= c->tol) && (xi2 <= c->tol);  // false for a NaN residual, like Julia's max(NaN, x) = NaN
  } else {
    const double t1 = fmax(c->tol * c->res0[0], c->tol), t2 = fmax(c->tol * c->res0[1], c->tol);
    stop = (xi1 <= t1) && (xi2 <= t2);
  }
  if (c->res0[0] == -INFINITY) c->res0[0] = xi1;
  if (c->res0[1] == -INFINITY) c->res0[1] = xi2;
  c->xi1 = xi1;
  c->xi2 = xi2;
  if (log && c->passes < c->log_cap) {
    log[3 * (size_t)c->passes + 0] = xi1;
    log[3 * (size_t)c->passes + 1] = xi2;
    log[3 * (size_t)c->passes + 2] = (is_sp && c->branch == 3) ? 9.0 : 0.0;
  }
  c->passes++;
  c->it++;
  if (stop && !c->ignore_done) c->done = 1;
}
