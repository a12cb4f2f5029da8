// spk_offline_dev.cuh -- the offline phase of build_model on the device:
//   * spk_k_riccati: `ricatti_offline` (src/model_dynamics/implicit_l.jl:46-97) per stage -- on every tree the reference can
//     generate all nodes of a stage see the same child modes, so the recursion runs over the N-1 parent stages (O(N d nx^3))
//     and writes the per-(stage, child) tables that the sweeps read (or that are replicated into per-node slots);
//   * spk_k_projector: the kernel projector of cpock.jl:71-81,125-128, one per risk measure -- one thread per non-leaf node
//     with non-uniform risk measures (get_nonuniform_rms_avar_v2: hundreds of thousands of small projectors).
// Same formulas and the same order of operations as the host versions in spk_offline.cpp, which remain as the cross-check
// (SPK_HOST_OFFLINE=1; tests compare the two to 1e-12).  Both kernels also run in the host-emulation build.
#pragma once
#include "spk_types.h"

struct SpkRiccatiArgs {
  int N, d, nx, nu;
  long abk_stride, k_stride, r_stride;
  const double *A, *B;                          // d modes, column-major
  double *P, *ABK, *ABKt, *K, *Kt, *G, *Rinv;   // tables (SpkFactorTable layout)
  double *ws;                                   // scratch, spk_riccati_ws_doubles
  int *fail;                                    // set to 1 + stage when I + sum B'PB is not positive definite
};
SPK_HD long spk_riccati_ws_doubles(int d, int nx, int nu) {
  // PB (d nx nu), PA / PABK (d nx nx), ABK (d nx nx), sum_r, Rt, L (nu nu each), sum_k (nu nx), Pt (nx nx)
  return (long)d * nx * nu + 2L * d * nx * nx + 3L * nu * nu + (long)nu * nx + (long)nx * nx;
}

// C (m x n) = [C +] op(A) * B with op(A) = A (m x k, ld m) or A' (A is k x m); one output entry per thread, serial dot
SPK_D void spk_cta_matmul(double *C, const double *A, bool a_trans, const double *B, int m, int k, int n, bool accumulate, int tid, int T) {
  for (int idx = tid; idx < m * n; idx += T) {
    const int i = idx % m, j = idx / m;
    double s = 0.0;
    if (a_trans)
      for (int q = 0; q < k; ++q) s += A[q + (long)i * k] * B[q + (long)j * k];
    else
      for (int q = 0; q < k; ++q) s += A[i + (long)q * m] * B[q + (long)j * k];
    C[idx] = accumulate ? C[idx] + s : s;
  }
}
SPK_D int spk_child_mode_off(int d, int c) { return (c + 2) % d; }  // Julia w(j) = (j % d) + 1 with j = 2 + c (mod d)

SPK_KERNEL void spk_k_riccati(SpkRiccatiArgs R) {
  const int nx = R.nx, nu = R.nu, d = R.d, S = R.N - 1;
  const int tid = threadIdx.x, T = blockDim.x;
  double *PB = R.ws, *PA = PB + (long)d * nx * nu, *ABKs = PA + (long)d * nx * nx, *sum_r = ABKs + (long)d * nx * nx, *Rt = sum_r + nu * nu,
         *L = Rt + nu * nu, *sum_k = L + nu * nu, *Pt = sum_k + nu * nx;
  // leaves: P = I (implicit_l.jl:67-69)
  double *Pleaf = R.P + (long)(R.N - 1) * nx * nx;
  for (int idx = tid; idx < nx * nx; idx += T) Pleaf[idx] = (idx % nx == idx / nx) ? 1.0 : 0.0;
  SPK_SYNCTHREADS();
  for (int t = S - 1; t >= 0; --t) {
    const double *Pn = R.P + (long)(t + 1) * nx * nx;
    for (int c = 0; c < d; ++c) {  // :76-79
      const int w = spk_child_mode_off(d, c);
      spk_cta_matmul(PB + (long)c * nx * nu, Pn, false, R.B + (long)w * nx * nu, nx, nx, nu, false, tid, T);
      spk_cta_matmul(PA + (long)c * nx * nx, Pn, false, R.A + (long)w * nx * nx, nx, nx, nx, false, tid, T);
    }
    SPK_SYNCTHREADS();
    for (int c = 0; c < d; ++c) {
      const int w = spk_child_mode_off(d, c);
      spk_cta_matmul(sum_r, R.B + (long)w * nx * nu, true, PB + (long)c * nx * nu, nu, nx, nu, c > 0, tid, T);
      spk_cta_matmul(sum_k, R.B + (long)w * nx * nu, true, PA + (long)c * nx * nx, nu, nx, nx, c > 0, tid, T);
    }
    SPK_SYNCTHREADS();
    for (int idx = tid; idx < nu * nu; idx += T) {  // :82-84 symmetrise, add I
      const int i = idx % nu, j = idx / nu;
      Rt[idx] = (i == j ? 1.0 : 0.0) + 0.5 * (sum_r[i + j * nu] + sum_r[j + i * nu]);
      L[idx] = 0.0;
    }
    SPK_SYNCTHREADS();
    for (int j = 0; j < nu; ++j) {  // lower Cholesky factor, column by column
      if (tid == 0) {
        double djj = Rt[j + j * nu];
        for (int k = 0; k < j; ++k) djj -= L[j + k * nu] * L[j + k * nu];
        if (!(djj > 0.0)) {
          *R.fail = 1 + t;
          djj = 1.0;
        }
        L[j + j * nu] = sqrt(djj);
      }
      SPK_SYNCTHREADS();
      const double djj = L[j + j * nu];
      for (int i = j + 1 + tid; i < nu; i += T) {
        double s = Rt[i + j * nu];
        for (int k = 0; k < j; ++k) s -= L[i + k * nu] * L[j + k * nu];
        L[i + j * nu] = s / djj;
      }
      SPK_SYNCTHREADS();
    }
    // K = -(L L')^{-1} sum_k (:86) and Rinv = (L L')^{-1}: one column per thread
    double *Kt_ = R.K + (long)t * R.k_stride, *Ri = R.Rinv + (long)t * R.r_stride;
    for (int col = tid; col < nx + nu; col += T) {
      double *X = col < nx ? Kt_ + (long)col * nu : Ri + (long)(col - nx) * nu;
      for (int i = 0; i < nu; ++i) X[i] = col < nx ? -sum_k[i + (long)col * nu] : (i == col - nx ? 1.0 : 0.0);
      for (int i = 0; i < nu; ++i) {
        double s = X[i];
        for (int k = 0; k < i; ++k) s -= L[i + k * nu] * X[k];
        X[i] = s / L[i + i * nu];
      }
      for (int i = nu - 1; i >= 0; --i) {
        double s = X[i];
        for (int k = i + 1; k < nu; ++k) s -= L[k + i * nu] * X[k];
        X[i] = s / L[i + i * nu];
      }
    }
    SPK_SYNCTHREADS();
    for (int idx = tid; idx < nu * nu; idx += T) {  // exact symmetry of Rinv (upper triangle <- mean)
      const int i = idx % nu, j = idx / nu;
      if (i > j) {
        const double m = 0.5 * (Ri[i + j * nu] + Ri[j + i * nu]);
        Ri[i + j * nu] = m;
        Ri[j + i * nu] = m;
      }
    }
    for (int idx = tid; idx < nu * nx; idx += T) {  // K' (nx x nu)
      const int i = idx % nu, j = idx / nu;
      R.Kt[(long)t * R.k_stride + j + (long)i * nx] = Kt_[idx];
    }
    // P_t = I + K'K + sum_c ABK_c' P ABK_c (:89-93); G = sum_c ABK_c' P B_w
    spk_cta_matmul(Pt, Kt_, true, Kt_, nx, nu, nx, false, tid, T);
    for (int c = 0; c < d; ++c) {  // ABK_c = A_w + B_w K
      const int w = spk_child_mode_off(d, c);
      spk_cta_matmul(ABKs + (long)c * nx * nx, R.B + (long)w * nx * nu, false, Kt_, nx, nu, nx, false, tid, T);
    }
    SPK_SYNCTHREADS();
    for (int idx = tid; idx < nx * nx; idx += T) Pt[idx] += (idx % nx == idx / nx) ? 1.0 : 0.0;
    for (int c = 0; c < d; ++c) {
      const int w = spk_child_mode_off(d, c);
      double *ABKc = ABKs + (long)c * nx * nx;
      for (int idx = tid; idx < nx * nx; idx += T) ABKc[idx] += R.A[(long)w * nx * nx + idx];
    }
    SPK_SYNCTHREADS();
    for (int c = 0; c < d; ++c)  // PA is free: reuse it for P ABK_c
      spk_cta_matmul(PA + (long)c * nx * nx, Pn, false, ABKs + (long)c * nx * nx, nx, nx, nx, false, tid, T);
    SPK_SYNCTHREADS();
    double *Gt = R.G + (long)t * R.k_stride;
    for (int c = 0; c < d; ++c) {
      const double *ABKc = ABKs + (long)c * nx * nx;
      spk_cta_matmul(Pt, ABKc, true, PA + (long)c * nx * nx, nx, nx, nx, true, tid, T);
      spk_cta_matmul(Gt, ABKc, true, PB + (long)c * nx * nu, nx, nx, nu, c > 0, tid, T);
      double *o = R.ABK + ((long)t * d + c) * R.abk_stride, *ot = R.ABKt + ((long)t * d + c) * R.abk_stride;
      for (int idx = tid; idx < nx * nx; idx += T) {
        const int i = idx % nx, j = idx / nx;
        o[idx] = ABKc[idx];
        ot[j + (long)i * nx] = ABKc[idx];
      }
    }
    SPK_SYNCTHREADS();
    double *Pout = R.P + (long)t * nx * nx;
    for (int idx = tid; idx < nx * nx; idx += T) Pout[idx] = Pt[idx];
    SPK_SYNCTHREADS();
  }
}

// ---- kernel projector (cpock.jl:71-81,125-128): I - Q Q' with Q an orthonormal basis of the row space of
//      M = [E' -I -I; F' 0 0] ((2d) x (ny+2d)), by pivoted Gram-Schmidt with re-orthogonalisation (rank-revealing: the
//      all-zero rows of F' are dropped) -- the algorithm of spk_kernel_projector (spk_offline.cpp), one thread per risk measure.
//      rows: scratch of n_rms * 2d * nls doubles (the accepted basis vectors overwrite the rows they came from).
SPK_KERNEL void spk_k_projector(int ny, int d, long n_rms, const double *__restrict__ E, const double *__restrict__ F,
                                double *__restrict__ rows_all, double *__restrict__ ls_all, int *fail) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rms) return;
  const int nls = ny + 2 * d, mr = 2 * d;
  const double *Er = E + r * (long)ny * d, *Fr = F + r * (long)ny * d;
  double *rows = rows_all + r * (long)mr * nls, *ls = ls_all + r * (long)nls * nls;
  for (int k = 0; k < mr * nls; ++k) rows[k] = 0.0;
  for (int c = 0; c < d; ++c) {
    for (int k = 0; k < ny; ++k) rows[c * nls + k] = Er[k + (long)c * ny];  // (E')[c,k] = E[k,c]
    rows[c * nls + ny + c] = -1.0;
    rows[c * nls + ny + d + c] = -1.0;
    for (int k = 0; k < ny; ++k) rows[(d + c) * nls + k] = Fr[k + (long)c * ny];
  }
  double scale = 0.0;
  for (int i = 0; i < mr; ++i) {
    double s = 0.0;
    for (int k = 0; k < nls; ++k) s += rows[i * nls + k] * rows[i * nls + k];
    s = sqrt(s);
    if (s > scale) scale = s;
  }
  if (!(scale > 0.0)) {
    *fail = 1;
    return;
  }
  unsigned long long used = 0ull, basis = 0ull;  // rows consumed / rows that now hold an accepted basis vector (2d <= 62)
  for (int it = 0; it < mr; ++it) {
    int best = -1;
    double bn = 0.0;
    for (int i = 0; i < mr; ++i)
      if (!((used >> i) & 1ull)) {
        double s = 0.0;
        for (int k = 0; k < nls; ++k) s += rows[i * nls + k] * rows[i * nls + k];
        s = sqrt(s);
        if (s > bn) { bn = s; best = i; }
      }
    if (best < 0 || bn <= 1e-12 * scale) break;
    used |= 1ull << best;
    double *q = rows + best * nls;
    for (int pass = 0; pass < 2; ++pass)  // re-orthogonalise against the accepted basis (in acceptance order = any order)
      for (int b = 0; b < mr; ++b)
        if ((basis >> b) & 1ull) {
          const double *bv = rows + b * nls;
          double s = 0.0;
          for (int k = 0; k < nls; ++k) s += bv[k] * q[k];
          for (int k = 0; k < nls; ++k) q[k] -= s * bv[k];
        }
    double nq = 0.0;
    for (int k = 0; k < nls; ++k) nq += q[k] * q[k];
    nq = sqrt(nq);
    if (nq <= 1e-12 * scale) continue;
    for (int k = 0; k < nls; ++k) q[k] /= nq;
    for (int i = 0; i < mr; ++i)  // deflate the remaining rows
      if (!((used >> i) & 1ull)) {
        double s = 0.0;
        for (int k = 0; k < nls; ++k) s += q[k] * rows[i * nls + k];
        for (int k = 0; k < nls; ++k) rows[i * nls + k] -= s * q[k];
      }
    basis |= 1ull << best;
  }
  for (int j = 0; j < nls; ++j)
    for (int i = 0; i < nls; ++i) {
      double v = (i == j) ? 1.0 : 0.0;
      for (int b = 0; b < mr; ++b)
        if ((basis >> b) & 1ull) v -= rows[b * nls + i] * rows[b * nls + j];
      ls[i + (long)j * nls] = v;
    }
}
