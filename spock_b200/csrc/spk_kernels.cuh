// spk_kernels.cuh -- operator kernels of the CP / SuperMann hot path (fp64, reference vector layout).
//
// Each kernel names the reference function it replaces (file:line under /root/reference/src).
// Work decomposition:
//   * element kernels (L, L', box/cone parts of prox_h*, vector ops): one thread per output element,
//     index decoded from the reference layout, neighbouring threads touch neighbouring addresses;
//   * group kernels (S1 sweeps, SOC projections): LANES (1..32, power of two) lanes cooperate on one tree
//     node; lanes stride over the rows of column-major factor matrices so every load is coalesced and no
//     cross-lane reduction is needed for the mat-vecs (transposed copies of the factors are stored for the
//     backward sweep); the only reductions are the SOC norms (warp shuffles).
// All kernels are correct for any power-of-two block size and for inactive tail groups.
#pragma once
#include "spk_types.h"

// ------------------------------------------------------------------------------------------------ group helpers

template <int LANES>
SPK_D double spk_group_sum(double v) {
#if !SPK_EMU
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
#endif
  return v;
}

// MOI SecondOrderCone (t, x) projection, risk_measures.jl:40-71, on an in-thread vector
SPK_HD void spk_project_soc_serial(double *vec, int len) {
  const double t = vec[0];
  double nrm = 0.0;
  for (int i = 1; i < len; ++i) nrm += vec[i] * vec[i];
  nrm = sqrt(nrm);
  if (nrm <= t) {
  } else if (nrm <= -t) {
    for (int k = 0; k < len; ++k) vec[k] = 0.0;
  } else {
    vec[0] = (t + nrm) / 2.0;
    for (int k = 1; k < len; ++k) vec[k] = vec[0] * vec[k] / nrm;
  }
}

// ------------------------------------------------------------------------------------------------ L and L'

// out[e] = alpha * (L' v)[e] + beta * in[e]     L_transpose! implicit_l.jl:320-449 fused with spock_mul! :451-481
SPK_D double spk_Lt_entry(const SpkDims &D, const SpkPtrs &P, const double *__restrict__ v, long e) {
  const int nx = D.nx, nu = D.nu, d = D.d;
  double val;
  if (e < D.zo[1]) {  // x
    const long i0 = e / nx;
    const int k = (int)(e - i0 * nx);
    if (i0 < D.nnl) {  // :351-373
      val = v[D.vo[6] + i0 * (nx + nu) + k];
      for (int c = 0; c < d; ++c) {
        const long j0 = (long)d * i0 + 1 + c;
        val += spk_cost_row(spk_sqrtQ(D, P, j0), D.cost_dense, nx, k, v + D.vo[2] + (j0 - 1) * nx);
      }
    } else {  // :376-396
      const long l = i0 - D.nnl;
      val = v[D.vo[10] + l * nx + k];
      val += spk_cost_row(spk_sqrtQN(D, P, l), D.cost_dense, nx, k, v + D.vo[7] + l * nx);
    }
  } else if (e < D.zo[2]) {  // u :399-420
    const long idx = e - D.zo[1];
    const long i0 = idx / nu;
    const int k = (int)(idx - i0 * nu);
    val = v[D.vo[6] + i0 * (nx + nu) + nx + k];
    for (int c = 0; c < d; ++c) {
      const long j0 = (long)d * i0 + 1 + c;
      val += spk_cost_row(spk_sqrtR(D, P, j0), D.cost_dense, nu, k, v + D.vo[3] + (j0 - 1) * nu);
    }
  } else if (e < D.zo[3]) {  // s :440-448
    const long i0 = e - D.zo[2];
    if (i0 < D.nnl) {
      val = v[D.vo[1] + i0];
    } else {
      const long l = i0 - D.nnl;
      val = 0.5 * v[D.vo[8] + l] + 0.5 * v[D.vo[9] + l];
    }
  } else if (e < D.zo[4]) {  // tau :432-437
    const long jm = e - D.zo[3];
    val = 0.5 * v[D.vo[4] + jm] + 0.5 * v[D.vo[5] + jm];
  } else {  // y :423-429
    const long idx = e - D.zo[4];
    const long i0 = idx / D.ny;
    const int k = (int)(idx - i0 * D.ny);
    val = v[idx] - spk_b(D, P, i0)[k] * v[D.vo[1] + i0];
  }
  return val;
}
SPK_KERNEL void spk_k_Lt(SpkDims D, SpkPtrs P, const double *__restrict__ v, double alpha, double beta,
                         const double *__restrict__ in, double *__restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D.nz) return;
  const double val = spk_Lt_entry(D, P, v, e);
  out[e] = (beta == 0.0) ? alpha * val : alpha * val + beta * in[e];
}

// out[e] = alpha * (L z)[e] + beta * in[e]      L! implicit_l.jl:177-318 fused with spock_mul!
SPK_D double spk_L_entry(const SpkDims &D, const SpkPtrs &P, const double *__restrict__ z, long e) {
  const int nx = D.nx, nu = D.nu, d = D.d;
  double val;
  if (e < D.vo[1]) {  // v1 :207-210
    val = z[D.zo[4] + e];
  } else if (e < D.vo[2]) {  // v2 :214-222
    const long i0 = e - D.vo[1];
    double dot_p = 0.0;
    for (int k = 0; k < D.ny; ++k) dot_p += spk_b(D, P, i0)[k] * z[D.zo[4] + i0 * D.ny + k];
    val = z[D.zo[2] + i0] - dot_p;
  } else if (e < D.vo[3]) {  // v3 :226-235
    const long idx = e - D.vo[2];
    const long jm = idx / nx;
    const int k = (int)(idx - jm * nx);
    val = spk_cost_row(spk_sqrtQ(D, P, jm + 1), D.cost_dense, nx, k, z + (jm / d) * nx);
  } else if (e < D.vo[4]) {  // v4 :239-248
    const long idx = e - D.vo[3];
    const long jm = idx / nu;
    const int k = (int)(idx - jm * nu);
    val = spk_cost_row(spk_sqrtR(D, P, jm + 1), D.cost_dense, nu, k, z + D.zo[1] + (jm / d) * nu);
  } else if (e < D.vo[5]) {  // v5 :253-260
    val = 0.5 * z[D.zo[3] + (e - D.vo[4])];
  } else if (e < D.vo[6]) {  // v6 :265-272
    val = 0.5 * z[D.zo[3] + (e - D.vo[5])];
  } else if (e < D.vo[7]) {  // v7 :275-282, constraints.jl:111-127
    const long idx = e - D.vo[6];
    const long i0 = idx / (nx + nu);
    const int k = (int)(idx - i0 * (nx + nu));
    val = (k < nx) ? z[i0 * nx + k] : z[D.zo[1] + i0 * nu + (k - nx)];
  } else if (e < D.vo[8]) {  // v11 :288-294
    const long idx = e - D.vo[7];
    const long l = idx / nx;
    const int k = (int)(idx - l * nx);
    val = spk_cost_row(spk_sqrtQN(D, P, l), D.cost_dense, nx, k, z + (D.nnl + l) * nx);
  } else if (e < D.vo[9]) {  // v12 :298-301
    val = 0.5 * z[D.zo[2] + D.nnl + (e - D.vo[8])];
  } else if (e < D.vo[10]) {  // v13 :305-308
    val = 0.5 * z[D.zo[2] + D.nnl + (e - D.vo[9])];
  } else {  // v14 :311-317
    val = z[D.nnl * nx + (e - D.vo[10])];
  }
  return val;
}
SPK_KERNEL void spk_k_L(SpkDims D, SpkPtrs P, const double *__restrict__ z, double alpha, double beta,
                        const double *__restrict__ in, double *__restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D.nv) return;
  const double val = spk_L_entry(D, P, z, e);
  out[e] = (beta == 0.0) ? alpha * val : alpha * val + beta * in[e];
}

// ------------------------------------------------------------------------------------------------ prox_f: S1 sweeps

// Backward step of projection_S1! (implicit_l.jl:584-637) for every node of level t (one launch per level).
// arg = (xbar | ubar) in/out vector; q (n*nx), dd (nnl*nu) scratch vectors of the sweep.
// d_i = Rinv_i (ubar_i - sum_j B_w' q_j);  q_i = sum_j ABK_j' q_j + G_i d_i + K_i'(d_i - ubar_i) - xbar_i,
// with G_i = sum_j ABK_j' P_j B_w folded offline (same value as :610-624 up to rounding).
// Shared memory per group: d*nx + 3*nu doubles.
template <int LANES>
SPK_KERNEL void spk_k_s1_bwd(SpkDims D, SpkPtrs P, int t, int children_are_leaves, const double *__restrict__ arg,
                             double *__restrict__ q, double *__restrict__ dd) {
  SPK_SMEM_DECL(double, smem);
  const int nx = D.nx, nu = D.nu, d = D.d;
  const int gpb = blockDim.x / LANES, g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long count = spk_level_count(d, t), first0 = spk_level_first0(d, t);
  const long p_raw = (long)blockIdx.x * gpb + g;
  const bool active = p_raw < count;
  const long p = active ? p_raw : 0;
  const long i0 = first0 + p;
  double *qc = smem + (long)g * (d * nx + 3 * nu), *tu = qc + d * nx, *rhs = tu + nu, *dv = rhs + nu;
  const long cbase = ((long)d * i0 + 1) * nx;  // children x-entries are contiguous
  for (int idx = lane; idx < d * nx; idx += LANES) qc[idx] = children_are_leaves ? -arg[cbase + idx] : q[cbase + idx];
  for (int k = lane; k < nu; k += LANES) tu[k] = arg[D.zo[1] + i0 * nu + k];
  SPK_SYNCWARP();
  // sum_for_d (:590-596) and the right-hand side (:598-600)
  for (int k = lane; k < nu; k += LANES) {
    double s = 0.0;
    for (int c = 0; c < d; ++c) {
      const double *Bt = P.Bt + (long)spk_mode(P, d, (long)d * i0 + 1 + c) * nx * nu;
      const double *qj = qc + c * nx;
      double acc = 0.0;
      for (int r = 0; r < nx; ++r) acc += Bt[k + (long)r * nu] * qj[r];
      s += acc;
    }
    rhs[k] = tu[k] - s;
  }
  SPK_SYNCWARP();
  const long kslot = D.per_node ? i0 : t;
  const double *Rinv = P.Rinv + kslot * D.r_stride;
  for (int k = lane; k < nu; k += LANES) {  // :601 ldiv!(R_chol, .) as a product with the explicit inverse
    double acc = 0.0;
    for (int l = 0; l < nu; ++l) acc += Rinv[k + (long)l * nu] * rhs[l];
    dv[k] = acc;
    if (active) dd[i0 * nu + k] = acc;
  }
  SPK_SYNCWARP();
  const double *G = P.G + kslot * D.k_stride, *Kt = P.Kt + kslot * D.k_stride;
  for (int r = lane; r < nx; r += LANES) {
    double acc = 0.0;
    for (int c = 0; c < d; ++c) {
      const long aslot = D.per_node ? ((long)d * i0 + c) : ((long)t * d + c);
      const double *At = P.ABKt + aslot * D.abk_stride;
      const double *qj = qc + c * nx;
      double a2 = 0.0;
      for (int m = 0; m < nx; ++m) a2 += At[r + (long)m * nx] * qj[m];
      acc += a2;
    }
    double a3 = 0.0;
    for (int k = 0; k < nu; ++k) a3 += G[r + (long)k * nx] * dv[k];
    acc += a3;
    double a4 = 0.0;
    for (int k = 0; k < nu; ++k) a4 += Kt[r + (long)k * nx] * (dv[k] - tu[k]);  // :626-632
    acc += a4;
    acc -= arg[i0 * nx + r];  // :634-636
    if (active) q[i0 * nx + r] = acc;
  }
}

// Forward step of projection_S1! (implicit_l.jl:667-681) for level t: u_i = K_i x_i + d_i,
// x_j = ABK_j x_i + B_w d_i.  Shared memory per group: nx + nu doubles.
template <int LANES>
SPK_KERNEL void spk_k_s1_fwd(SpkDims D, SpkPtrs P, int t, double *__restrict__ arg, const double *__restrict__ dd) {
  SPK_SMEM_DECL(double, smem);
  const int nx = D.nx, nu = D.nu, d = D.d;
  const int gpb = blockDim.x / LANES, g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long count = spk_level_count(d, t), first0 = spk_level_first0(d, t);
  const long p_raw = (long)blockIdx.x * gpb + g;
  const bool active = p_raw < count;
  const long p = active ? p_raw : 0;
  const long i0 = first0 + p;
  double *xs = smem + (long)g * (nx + nu), *ds = xs + nx;
  for (int r = lane; r < nx; r += LANES) {
    const double xv = (t == 0) ? P.x0[r] : arg[i0 * nx + r];  // :667
    xs[r] = xv;
    if (t == 0 && active) arg[r] = xv;
  }
  for (int k = lane; k < nu; k += LANES) ds[k] = dd[i0 * nu + k];
  SPK_SYNCWARP();
  const long kslot = D.per_node ? i0 : t;
  const double *K = P.K + kslot * D.k_stride;
  for (int k = lane; k < nu; k += LANES) {  // :670-673
    double acc = 0.0;
    for (int m = 0; m < nx; ++m) acc += K[k + (long)m * nu] * xs[m];
    acc += ds[k];
    if (active) arg[D.zo[1] + i0 * nu + k] = acc;
  }
  for (int c = 0; c < d; ++c) {  // :674-680
    const long j0 = (long)d * i0 + 1 + c;
    const long aslot = D.per_node ? (j0 - 1) : ((long)t * d + c);
    const double *Am = P.ABK + aslot * D.abk_stride;
    const double *Bw = P.B + (long)spk_mode(P, d, j0) * nx * nu;
    for (int r = lane; r < nx; r += LANES) {
      double acc = 0.0;
      for (int m = 0; m < nx; ++m) acc += Am[r + (long)m * nx] * xs[m];
      double a2 = 0.0;
      for (int k = 0; k < nu; ++k) a2 += Bw[r + (long)k * nx] * ds[k];
      acc += a2;
      if (active) arg[j0 * nx + r] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ prox_f: S2

// projection_S2! (implicit_l.jl:688-727): one thread per non-leaf node multiplies (y_i, s_children, tau_children)
// by the kernel projector in place.  arg is the full z-sized vector.
SPK_KERNEL void spk_k_s2(SpkDims D, SpkPtrs P, double *__restrict__ arg) {
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= D.nnl) return;
  const int ny = D.ny, d = D.d, nls = D.nls;
  double in[SPK_MAX_LS], outv[SPK_MAX_LS];
  for (int k = 0; k < ny; ++k) in[k] = arg[D.zo[4] + i0 * ny + k];                    // :712
  for (int c = 0; c < d; ++c) in[ny + c] = arg[D.zo[2] + (long)d * i0 + 1 + c];       // :713 s of the children
  for (int c = 0; c < d; ++c) in[ny + d + c] = arg[D.zo[3] + (long)d * i0 + c];       // :714 tau of the children
  for (int k = 0; k < nls; ++k) outv[k] = 0.0;
  for (int m = 0; m < nls; ++m) {  // :717
    const double xm = in[m];
    for (int k = 0; k < nls; ++k) outv[k] += spk_ls(D, P, i0)[k + (long)m * nls] * xm;
  }
  for (int k = 0; k < ny; ++k) arg[D.zo[4] + i0 * ny + k] = outv[k];                  // :721
  for (int c = 0; c < d; ++c) arg[D.zo[2] + (long)d * i0 + 1 + c] = outv[ny + c];     // :723
  for (int c = 0; c < d; ++c) arg[D.zo[3] + (long)d * i0 + c] = outv[ny + d + c];     // :725
}

// ------------------------------------------------------------------------------------------------ prox_h*
// prox_h_conj! (implicit_l.jl:911-951): w = arg/sigma (+-1/2 on v5,v6,v12,v13), out = sigma (w - Proj_S3(w)).

// v1 (dual-cone projections, :833-845) and v2 (:849-851): one thread per non-leaf node
SPK_KERNEL void spk_k_prox_own(SpkDims D, double sigma, const double *in, double *out) {
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= D.nnl) return;
  const int ny = D.ny;
  double w[SPK_MAX_LS], pr[SPK_MAX_LS];
  for (int k = 0; k < ny; ++k) {
    w[k] = in[i0 * ny + k] / sigma;
    pr[k] = w[k];
  }
  int off = 0;
  for (int c = 0; c < D.n_cones; ++c) {
    const int dim = D.cone_dim[c];
    switch (D.cone_type[c]) {  // projection on MOI.dual_set(cone), risk_measures.jl:14-71
      case 0: break;                                                                  // Zeros -> Reals
      case 1: for (int k = 0; k < dim; ++k) if (pr[off + k] < 0.0) pr[off + k] = 0.0; break;  // Nonnegatives
      case 2: for (int k = 0; k < dim; ++k) if (pr[off + k] > 0.0) pr[off + k] = 0.0; break;  // Nonpositives
      case 3: for (int k = 0; k < dim; ++k) pr[off + k] = 0.0; break;                 // Reals -> Zeros
      case 4: spk_project_soc_serial(pr + off, dim); break;                           // SOC self-dual
    }
    off += dim;
  }
  for (int k = 0; k < ny; ++k) out[i0 * ny + k] = sigma * (w[k] - pr[k]);
  const double w2 = in[D.vo[1] + i0] / sigma;
  const double p2 = w2 >= 0 ? w2 : 0.0;
  out[D.vo[1] + i0] = sigma * (w2 - p2);
}

// v7 box (:903-908, constraints.jl:38-55) and v14 box (:804-809, constraints.jl:24-33): one thread per element
SPK_KERNEL void spk_k_prox_box(SpkDims D, double sigma, const double *in, double *out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n7 = D.nnl * (long)(D.nx + D.nu), n14 = D.nl * (long)D.nx;
  if (e >= n7 + n14) return;
  long idx;
  double lo, hi;
  if (e < n7) {
    idx = D.vo[6] + e;
    const int k = (int)(e % (D.nx + D.nu));
    lo = k < D.nx ? D.x_min : D.u_min;
    hi = k < D.nx ? D.x_max : D.u_max;
  } else {
    idx = D.vo[10] + (e - n7);
    lo = D.x_min;
    hi = D.x_max;
  }
  const double w = in[idx] / sigma;
  out[idx] = sigma * (w - spk_clip(w, lo, hi));
}

// SOC(nx+nu+2) on (t = v6_j; v3_j, v4_j, v5_j) for every non-root node j (:854-900): one group per node
template <int LANES>
SPK_KERNEL void spk_k_prox_soc_child(SpkDims D, double sigma, const double *in, double *out) {
  const int nx = D.nx, nu = D.nu;
  const int gpb = blockDim.x / LANES, g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long jm_raw = (long)blockIdx.x * gpb + g;  // jm = j0 - 1
  const bool active = jm_raw < D.n - 1;
  const long jm = active ? jm_raw : 0;
  const double *v3 = in + D.vo[2] + jm * nx, *v4 = in + D.vo[3] + jm * nu;
  double part = 0.0;
  for (int k = lane; k < nx; k += LANES) {
    const double w = v3[k] / sigma;
    part += w * w;
  }
  for (int k = lane; k < nu; k += LANES) {
    const double w = v4[k] / sigma;
    part += w * w;
  }
  const double w5 = in[D.vo[4] + jm] / sigma - 0.5, t = in[D.vo[5] + jm] / sigma + 0.5;
  if (lane == 0) part += w5 * w5;
  const double nrm = sqrt(spk_group_sum<LANES>(part));
  // projection (risk_measures.jl:57-70): keep | zero | scale; out = sigma (w - proj)
  int mode;
  double tp = 0.0;
  if (nrm <= t) mode = 0; else if (nrm <= -t) mode = 1; else { mode = 2; tp = (t + nrm) / 2.0; }
  if (!active) return;
  double *o3 = out + D.vo[2] + jm * nx, *o4 = out + D.vo[3] + jm * nu;
  for (int k = lane; k < nx; k += LANES) {
    const double w = v3[k] / sigma;
    const double pr = mode == 0 ? w : (mode == 1 ? 0.0 : tp * w / nrm);
    o3[k] = sigma * (w - pr);
  }
  for (int k = lane; k < nu; k += LANES) {
    const double w = v4[k] / sigma;
    const double pr = mode == 0 ? w : (mode == 1 ? 0.0 : tp * w / nrm);
    o4[k] = sigma * (w - pr);
  }
  if (lane == 0) {
    const double p5 = mode == 0 ? w5 : (mode == 1 ? 0.0 : tp * w5 / nrm);
    const double pt = mode == 0 ? t : (mode == 1 ? 0.0 : tp);
    out[D.vo[4] + jm] = sigma * (w5 - p5);
    out[D.vo[5] + jm] = sigma * (t - pt);
  }
}

// SOC(nx+2) on (t = v13; v11, v12) for every leaf (:768-801): one group per leaf
template <int LANES>
SPK_KERNEL void spk_k_prox_soc_leaf(SpkDims D, double sigma, const double *in, double *out) {
  const int nx = D.nx;
  const int gpb = blockDim.x / LANES, g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long l_raw = (long)blockIdx.x * gpb + g;
  const bool active = l_raw < D.nl;
  const long l = active ? l_raw : 0;
  const double *v11 = in + D.vo[7] + l * nx;
  double part = 0.0;
  for (int k = lane; k < nx; k += LANES) {
    const double w = v11[k] / sigma;
    part += w * w;
  }
  const double w12 = in[D.vo[8] + l] / sigma - 0.5, t = in[D.vo[9] + l] / sigma + 0.5;
  if (lane == 0) part += w12 * w12;
  const double nrm = sqrt(spk_group_sum<LANES>(part));
  int mode;
  double tp = 0.0;
  if (nrm <= t) mode = 0; else if (nrm <= -t) mode = 1; else { mode = 2; tp = (t + nrm) / 2.0; }
  if (!active) return;
  double *o11 = out + D.vo[7] + l * nx;
  for (int k = lane; k < nx; k += LANES) {
    const double w = v11[k] / sigma;
    const double pr = mode == 0 ? w : (mode == 1 ? 0.0 : tp * w / nrm);
    o11[k] = sigma * (w - pr);
  }
  if (lane == 0) {
    const double p12 = mode == 0 ? w12 : (mode == 1 ? 0.0 : tp * w12 / nrm);
    const double pt = mode == 0 ? t : (mode == 1 ? 0.0 : tp);
    out[D.vo[8] + l] = sigma * (w12 - p12);
    out[D.vo[9] + l] = sigma * (t - pt);
  }
}

// ------------------------------------------------------------------------------------------------ vector ops

// out = a*x + b*y (y may be NULL when b == 0)
SPK_KERNEL void spk_k_axpby(long n, double a, const double *x, double b, const double *y, double *out) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];
}

// Shift-initialisation of a receding-horizon warm start (the reference's closed loop,
// examples/server_heat/mpc_simulation.jl:38-54, re-solves from the previous (z, v) unshifted): after the first input has
// been applied and the disturbance `w` (ordinal of the realised child of the root) observed, the new problem's node at
// (stage t, position p) is the old problem's node at (stage t+1, position w d^t + p).  Every entry of (z | v) keyed by a node
// takes the old value of its source node where that source has such an entry; the last stage keeps its old values.
SPK_KERNEL void spk_k_shift(SpkDims D, int w, const double *__restrict__ oldzv, double *__restrict__ zv) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D.n) return;
  int t = 0;
  long first = 0, cnt = 1;  // first0 and node count of level t
  while (j >= first + cnt) { first += cnt; cnt *= D.d; ++t; }
  if (t > D.N - 2) return;                          // leaves have no deeper source
  const long s = first + cnt + (long)w * cnt + (j - first);  // source node (0-based) on level t+1
  const double *oz = oldzv, *ov = oldzv + D.nz;
  double *z = zv, *v = zv + D.nz;
  for (int k = 0; k < D.nx; ++k) z[j * D.nx + k] = oz[s * D.nx + k];
  z[D.zo[2] + j] = oz[D.zo[2] + s];
  if (j >= 1) {
    z[D.zo[3] + j - 1] = oz[D.zo[3] + s - 1];
    for (int k = 0; k < D.nx; ++k) v[D.vo[2] + (j - 1) * D.nx + k] = ov[D.vo[2] + (s - 1) * D.nx + k];
    for (int k = 0; k < D.nu; ++k) v[D.vo[3] + (j - 1) * D.nu + k] = ov[D.vo[3] + (s - 1) * D.nu + k];
    v[D.vo[4] + j - 1] = ov[D.vo[4] + s - 1];
    v[D.vo[5] + j - 1] = ov[D.vo[5] + s - 1];
  }
  if (s < D.nnl) {                                   // entries keyed by a non-leaf node (j is non-leaf since t <= N-2)
    for (int k = 0; k < D.nu; ++k) z[D.zo[1] + j * D.nu + k] = oz[D.zo[1] + s * D.nu + k];
    for (int k = 0; k < D.ny; ++k) {
      z[D.zo[4] + j * D.ny + k] = oz[D.zo[4] + s * D.ny + k];
      v[D.vo[0] + j * D.ny + k] = ov[D.vo[0] + s * D.ny + k];
    }
    v[D.vo[1] + j] = ov[D.vo[1] + s];
    for (int k = 0; k < D.nx + D.nu; ++k) v[D.vo[6] + j * (D.nx + D.nu) + k] = ov[D.vo[6] + s * (D.nx + D.nu) + k];
  }
}

// x[idx] += a
SPK_KERNEL void spk_k_add_scalar(double *x, long idx, double a) {
  if (blockIdx.x == 0 && threadIdx.x == 0) x[idx] += a;
}

// Deterministic two-stage reductions: stage 1 writes one partial per block (fixed grid, fixed tree),
// stage 2 (one block) folds the partials in a fixed order.
// mode 0: sum x[i]*y[i]; mode 1: max |x[i]|; mode 2: sum x[i]
// mask (may be NULL): only entries with mask[i] != 0 take part -- the entries this rank counts under the subtree partition
SPK_KERNEL void spk_k_reduce_partial(long n, int mode, const double *__restrict__ x, const double *__restrict__ y,
                                     double *__restrict__ partial, const unsigned char *__restrict__ mask) {
  SPK_SMEM_DECL(double, sm);
  const long stride = (long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (mask && !mask[i]) continue;
    if (mode == 0) acc += x[i] * y[i];
    else if (mode == 2) acc += x[i];
    else acc = spk_max_nn(acc, fabs(x[i]));
  }
  sm[threadIdx.x] = acc;
  SPK_SYNCTHREADS();
  for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double o = sm[threadIdx.x + s];
      sm[threadIdx.x] = (mode != 1) ? sm[threadIdx.x] + o : spk_max_nn(sm[threadIdx.x], o);
    }
    SPK_SYNCTHREADS();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
// out[slot] (+)= fold(partial[0..np)); accumulate != 0 adds to the existing value (sum mode only)
SPK_KERNEL void spk_k_reduce_final(int np, int mode, const double *__restrict__ partial, double *__restrict__ out,
                                   int slot, int accumulate) {
  SPK_SMEM_DECL(double, sm);
  double acc = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const double o = partial[i];
    acc = (mode != 1) ? acc + o : spk_max_nn(acc, o);
  }
  sm[threadIdx.x] = acc;
  SPK_SYNCTHREADS();
  for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double o = sm[threadIdx.x + s];
      sm[threadIdx.x] = (mode != 1) ? sm[threadIdx.x] + o : spk_max_nn(sm[threadIdx.x], o);
    }
    SPK_SYNCTHREADS();
  }
  if (threadIdx.x == 0) out[slot] = (accumulate && mode != 1) ? out[slot] + sm[0] : sm[0];
}

// replicate table slots into the per-node slots of one tree level: dst[s*stride + e] = src[id(s)*stride + e]
// kind 0 (per child node; slot s = d*family + c): id = t*d + c;  kind 1 (per non-leaf node): id = t
SPK_KERNEL void spk_k_replicate_level(int d, int kind, int t, long stride, long n_slots, const double *__restrict__ src,
                                      double *__restrict__ dst) {
  const long total = n_slots * stride;
  const long gs = (long)gridDim.x * blockDim.x;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gs) {
    const long s = idx / stride, e = idx - s * stride;
    const long id = (kind == 0) ? (long)t * d + (s % d) : t;
    dst[idx] = src[id * stride + e];
  }
}
