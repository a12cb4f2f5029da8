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
#define SPK_RICCATI_MAX_D 32
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

// One step of the recursion for ONE family (a non-leaf node and its d children), by all threads of the CTA:
// Pc[c] = P of child c (NULL: the child is a leaf, P = I, implicit_l.jl:67-69), mode[c] = its dynamics mode (0-based).
// Writes P of the node (oP, may be NULL) and the factor slots oABK / oABKt (d slots of abk_stride), oK, oKt, oG, oRinv.
SPK_D void spk_riccati_family(int nx, int nu, int d, long abk_stride, const double *A, const double *B, const double *const *Pc, const int *mode,
                              double *ws, double *oP, double *oABK, double *oABKt, double *oK, double *oKt, double *oG, double *oRinv,
                              int *fail, int fail_code, int tid, int T) {
  double *PB = ws, *PA = PB + (long)d * nx * nu, *ABKs = PA + (long)d * nx * nx, *sum_r = ABKs + (long)d * nx * nx, *Rt = sum_r + nu * nu,
         *L = Rt + nu * nu, *sum_k = L + nu * nu, *Pt = sum_k + nu * nx;
  for (int c = 0; c < d; ++c) {  // :76-79
    const double *Bw = B + (long)mode[c] * nx * nu, *Aw = A + (long)mode[c] * nx * nx;
    if (Pc[c]) {
      spk_cta_matmul(PB + (long)c * nx * nu, Pc[c], false, Bw, nx, nx, nu, false, tid, T);
      spk_cta_matmul(PA + (long)c * nx * nx, Pc[c], false, Aw, nx, nx, nx, false, tid, T);
    } else {  // leaf child: P = I
      for (int idx = tid; idx < nx * nu; idx += T) PB[(long)c * nx * nu + idx] = Bw[idx];
      for (int idx = tid; idx < nx * nx; idx += T) PA[(long)c * nx * nx + idx] = Aw[idx];
    }
  }
  SPK_SYNCTHREADS();
  for (int c = 0; c < d; ++c) {
    const double *Bw = B + (long)mode[c] * nx * nu;
    spk_cta_matmul(sum_r, Bw, true, PB + (long)c * nx * nu, nu, nx, nu, c > 0, tid, T);
    spk_cta_matmul(sum_k, Bw, true, PA + (long)c * nx * nx, nu, nx, nx, c > 0, tid, T);
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
        *fail = fail_code;
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
  for (int col = tid; col < nx + nu; col += T) {
    double *X = col < nx ? oK + (long)col * nu : oRinv + (long)(col - nx) * nu;
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
  for (int idx = tid; idx < nu * nu; idx += T) {  // exact symmetry of Rinv
    const int i = idx % nu, j = idx / nu;
    if (i > j) {
      const double m = 0.5 * (oRinv[i + j * nu] + oRinv[j + i * nu]);
      oRinv[i + j * nu] = m;
      oRinv[j + i * nu] = m;
    }
  }
  for (int idx = tid; idx < nu * nx; idx += T) {  // K' (nx x nu)
    const int i = idx % nu, j = idx / nu;
    oKt[j + (long)i * nx] = oK[idx];
  }
  // P = I + K'K + sum_c ABK_c' P_c ABK_c (:89-93); G = sum_c ABK_c' P_c B_w
  spk_cta_matmul(Pt, oK, true, oK, nx, nu, nx, false, tid, T);
  for (int c = 0; c < d; ++c)  // ABK_c = A_w + B_w K
    spk_cta_matmul(ABKs + (long)c * nx * nx, B + (long)mode[c] * nx * nu, false, oK, nx, nu, nx, false, tid, T);
  SPK_SYNCTHREADS();
  for (int idx = tid; idx < nx * nx; idx += T) Pt[idx] += (idx % nx == idx / nx) ? 1.0 : 0.0;
  for (int c = 0; c < d; ++c) {
    double *ABKc = ABKs + (long)c * nx * nx;
    const double *Aw = A + (long)mode[c] * nx * nx;
    for (int idx = tid; idx < nx * nx; idx += T) ABKc[idx] += Aw[idx];
  }
  SPK_SYNCTHREADS();
  for (int c = 0; c < d; ++c) {  // PA is free: reuse it for P_c ABK_c
    if (Pc[c]) spk_cta_matmul(PA + (long)c * nx * nx, Pc[c], false, ABKs + (long)c * nx * nx, nx, nx, nx, false, tid, T);
    else
      for (int idx = tid; idx < nx * nx; idx += T) PA[(long)c * nx * nx + idx] = ABKs[(long)c * nx * nx + idx];
  }
  SPK_SYNCTHREADS();
  for (int c = 0; c < d; ++c) {
    const double *ABKc = ABKs + (long)c * nx * nx;
    spk_cta_matmul(Pt, ABKc, true, PA + (long)c * nx * nx, nx, nx, nx, true, tid, T);
    spk_cta_matmul(oG, ABKc, true, PB + (long)c * nx * nu, nx, nx, nu, c > 0, tid, T);
    double *o = oABK + (long)c * abk_stride, *ot = oABKt + (long)c * abk_stride;
    for (int idx = tid; idx < nx * nx; idx += T) {
      const int i = idx % nx, j = idx / nx;
      o[idx] = ABKc[idx];
      ot[j + (long)i * nx] = ABKc[idx];
    }
  }
  SPK_SYNCTHREADS();
  if (oP)
    for (int idx = tid; idx < nx * nx; idx += T) oP[idx] = Pt[idx];
  SPK_SYNCTHREADS();
}

// per stage: one CTA walks the N-1 parent stages (all nodes of a stage share their factors with the generator's mode map)
SPK_KERNEL void spk_k_riccati(SpkRiccatiArgs R) {
  const int nx = R.nx, d = R.d, S = R.N - 1;
  const int tid = threadIdx.x, T = blockDim.x;
  double *Pleaf = R.P + (long)(R.N - 1) * nx * nx;  // leaves: P = I (implicit_l.jl:67-69); kept in the table for spk_get_factors
  for (int idx = tid; idx < nx * nx; idx += T) Pleaf[idx] = (idx % nx == idx / nx) ? 1.0 : 0.0;
  SPK_SYNCTHREADS();
  for (int t = S - 1; t >= 0; --t) {
    const double *Pc[SPK_RICCATI_MAX_D];
    int mode[SPK_RICCATI_MAX_D];
    for (int c = 0; c < d; ++c) {
      Pc[c] = R.P + (long)(t + 1) * nx * nx;
      mode[c] = spk_child_mode_off(d, c);
    }
    spk_riccati_family(nx, R.nu, d, R.abk_stride, R.A, R.B, Pc, mode, R.ws, R.P + (long)t * nx * nx, R.ABK + (long)t * d * R.abk_stride,
                       R.ABKt + (long)t * d * R.abk_stride, R.K + (long)t * R.k_stride, R.Kt + (long)t * R.k_stride, R.G + (long)t * R.k_stride,
                       R.Rinv + (long)t * R.r_stride, R.fail, 1 + t, tid, T);
  }
}

// per node (a caller-supplied mode map makes the factors of the nodes of a stage differ: `ricatti_offline` as the reference
// runs it, implicit_l.jl:71-95, one node at a time -- here one CTA per node of a level, level by level from the leaves up).
// Pchild: P of the nodes of level t+1 in level order (NULL: they are leaves), Pout: P of the nodes of level t.
// Factor slots: node i0 -> slot i0 (single GPU).  ws: gridDim.x scratch areas of ws_stride doubles.
SPK_KERNEL void spk_k_riccati_level(SpkRiccatiArgs R, int t, long first0, long count, const double *Pchild, double *Pout, const int *w,
                                    long ws_stride) {
  const int nx = R.nx, d = R.d;
  const int tid = threadIdx.x, T = blockDim.x;
  double *ws = R.ws + (long)blockIdx.x * ws_stride;
  for (long p = blockIdx.x; p < count; p += gridDim.x) {
    const long i0 = first0 + p, c0 = (long)d * i0 + 1;
    const double *Pc[SPK_RICCATI_MAX_D];
    int mode[SPK_RICCATI_MAX_D];
    for (int c = 0; c < d; ++c) {
      Pc[c] = Pchild ? Pchild + ((long)d * p + c) * nx * nx : nullptr;
      mode[c] = w[c0 + c];
    }
    spk_riccati_family(nx, R.nu, d, R.abk_stride, R.A, R.B, Pc, mode, ws, Pout + p * (long)nx * nx, R.ABK + (long)d * i0 * R.abk_stride,
                       R.ABKt + (long)d * i0 * R.abk_stride, R.K + i0 * R.k_stride, R.Kt + i0 * R.k_stride, R.G + i0 * R.k_stride,
                       R.Rinv + i0 * R.r_stride, R.fail, 1 + t, tid, T);
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
