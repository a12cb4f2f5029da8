// spk_offline.cpp -- see spk_offline.h
#include "spk_offline.h"

#include <algorithm>
#include <cmath>

namespace {

// small dense column-major matrix
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
  double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * r]; }
  static Mat from(const double *p, int r_, int c_) {
    Mat m(r_, c_);
    std::copy(p, p + (size_t)r_ * c_, m.a.begin());
    return m;
  }
  static Mat eye(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
};

Mat mul(const Mat &A, const Mat &B) {  // A*B
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int k = 0; k < A.c; ++k) {
      const double b = B(k, j);
      for (int i = 0; i < A.r; ++i) C(i, j) += A(i, k) * b;
    }
  return C;
}
Mat mul_tn(const Mat &A, const Mat &B) {  // A'*B
  Mat C(A.c, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int i = 0; i < A.c; ++i) {
      double s = 0.0;
      for (int k = 0; k < A.r; ++k) s += A(k, i) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
Mat transpose(const Mat &A) {
  Mat T(A.c, A.r);
  for (int j = 0; j < A.c; ++j)
    for (int i = 0; i < A.r; ++i) T(j, i) = A(i, j);
  return T;
}
void add_into(Mat &A, const Mat &B) {
  for (size_t k = 0; k < A.a.size(); ++k) A.a[k] += B.a[k];
}

// lower Cholesky factor of a symmetric positive definite matrix; false if not SPD
bool cholesky(const Mat &S, Mat &L) {
  const int n = S.r;
  L = Mat(n, n);
  for (int j = 0; j < n; ++j) {
    double djj = S(j, j);
    for (int k = 0; k < j; ++k) djj -= L(j, k) * L(j, k);
    if (!(djj > 0.0)) return false;
    djj = std::sqrt(djj);
    L(j, j) = djj;
    for (int i = j + 1; i < n; ++i) {
      double s = S(i, j);
      for (int k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
      L(i, j) = s / djj;
    }
  }
  return true;
}
// X = (L L')^{-1} Bm
Mat chol_solve(const Mat &L, const Mat &Bm) {
  const int n = L.r;
  Mat X = Bm;
  for (int c = 0; c < X.c; ++c) {
    for (int i = 0; i < n; ++i) {
      double s = X(i, c);
      for (int k = 0; k < i; ++k) s -= L(i, k) * X(k, c);
      X(i, c) = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = X(i, c);
      for (int k = i + 1; k < n; ++k) s -= L(k, i) * X(k, c);
      X(i, c) = s / L(i, i);
    }
  }
  return X;
}

long even_up(long k) { return (k + 1) & ~1L; }

}  // namespace

void spk_table_init(int N, int d, int nx, int nu, SpkFactorTable &T) {
  T.N = N; T.d = d; T.nx = nx; T.nu = nu;
  T.abk_stride = even_up((long)nx * nx);
  T.k_stride = even_up((long)nx * nu);
  T.r_stride = even_up((long)nu * nu);
  const int S = N - 1;  // number of parent stages
  T.P.assign((size_t)N * nx * nx, 0.0);
  T.ABK.assign((size_t)std::max(S, 0) * d * T.abk_stride, 0.0);
  T.ABKt = T.ABK;
  T.K.assign((size_t)std::max(S, 0) * T.k_stride, 0.0);
  T.Kt = T.K;
  T.G = T.K;
  T.Rinv.assign((size_t)std::max(S, 0) * T.r_stride, 0.0);
}

bool spk_offline_factors(int N, int d, int nx, int nu, const double *A, const double *B, SpkFactorTable &T,
                         std::string &err) {
  spk_table_init(N, d, nx, nu, T);
  const int S = N - 1;
  std::vector<Mat> Am(d), Bm(d);
  for (int w = 0; w < d; ++w) {
    Am[w] = Mat::from(A + (size_t)w * nx * nx, nx, nx);
    Bm[w] = Mat::from(B + (size_t)w * nx * nu, nx, nu);
  }
  Mat Pnext = Mat::eye(nx);  // leaves: P = I (implicit_l.jl:67-69)
  std::copy(Pnext.a.begin(), Pnext.a.end(), T.P.begin() + (size_t)(N - 1) * nx * nx);
  for (int t = S - 1; t >= 0; --t) {
    Mat sum_r(nu, nu), sum_k(nu, nx);
    for (int c = 0; c < d; ++c) {  // :76-79
      const int w = spk_child_mode(d, c);
      const Mat PB = mul(Pnext, Bm[w]);
      const Mat PA = mul(Pnext, Am[w]);
      add_into(sum_r, mul_tn(Bm[w], PB));
      add_into(sum_k, mul_tn(Bm[w], PA));
    }
    Mat Rt(nu, nu);  // :82-84
    for (int j = 0; j < nu; ++j)
      for (int i = 0; i < nu; ++i) Rt(i, j) = (i == j ? 1.0 : 0.0) + 0.5 * (sum_r(i, j) + sum_r(j, i));
    Mat L;
    if (!cholesky(Rt, L)) {
      err = "Riccati recursion: I + sum B'PB is not positive definite at stage " + std::to_string(t);
      return false;
    }
    Mat negk = sum_k;
    for (double &x : negk.a) x = -x;
    const Mat K = chol_solve(L, negk);  // :86
    Mat Rinv = chol_solve(L, Mat::eye(nu));
    for (int j = 0; j < nu; ++j)  // exact symmetry
      for (int i = j + 1; i < nu; ++i) {
        const double m = 0.5 * (Rinv(i, j) + Rinv(j, i));
        Rinv(i, j) = m;
        Rinv(j, i) = m;
      }
    Mat Pt = Mat::eye(nx);
    add_into(Pt, mul_tn(K, K));
    Mat G(nx, nu);
    for (int c = 0; c < d; ++c) {  // :89-92
      const int w = spk_child_mode(d, c);
      Mat ABK = mul(Bm[w], K);
      add_into(ABK, Am[w]);
      const Mat PABK = mul(Pnext, ABK);
      add_into(Pt, mul_tn(ABK, PABK));
      add_into(G, mul_tn(ABK, mul(Pnext, Bm[w])));
      const Mat ABKt = transpose(ABK);
      std::copy(ABK.a.begin(), ABK.a.end(), T.ABK.begin() + ((size_t)t * d + c) * T.abk_stride);
      std::copy(ABKt.a.begin(), ABKt.a.end(), T.ABKt.begin() + ((size_t)t * d + c) * T.abk_stride);
    }
    const Mat Kt = transpose(K);
    std::copy(K.a.begin(), K.a.end(), T.K.begin() + (size_t)t * T.k_stride);
    std::copy(Kt.a.begin(), Kt.a.end(), T.Kt.begin() + (size_t)t * T.k_stride);
    std::copy(G.a.begin(), G.a.end(), T.G.begin() + (size_t)t * T.k_stride);
    std::copy(Rinv.a.begin(), Rinv.a.end(), T.Rinv.begin() + (size_t)t * T.r_stride);
    std::copy(Pt.a.begin(), Pt.a.end(), T.P.begin() + (size_t)t * nx * nx);
    Pnext = Pt;
  }
  return true;
}

bool spk_kernel_projector(int ny, int d, const double *E, const double *F, std::vector<double> &ls, std::string &err) {
  // M = [E' -I -I; F' 0 0] is (2d) x (ny+2d) (cpock.jl:71-81).  The projector on ker M is I - Q Q' with Q an
  // orthonormal basis of the row space of M, built by pivoted Gram-Schmidt with re-orthogonalisation
  // (rank-revealing: all-zero rows of F' are dropped).
  const int nls = ny + 2 * d, mr = 2 * d;
  std::vector<std::vector<double>> rows(mr, std::vector<double>(nls, 0.0));
  for (int c = 0; c < d; ++c) {
    for (int k = 0; k < ny; ++k) rows[c][k] = E[k + (size_t)c * ny];  // (E')[c,k] = E[k,c]
    rows[c][ny + c] = -1.0;
    rows[c][ny + d + c] = -1.0;
    for (int k = 0; k < ny; ++k) rows[d + c][k] = F[k + (size_t)c * ny];
  }
  auto nrm2 = [&](const std::vector<double> &v) {
    double s = 0.0;
    for (double x : v) s += x * x;
    return std::sqrt(s);
  };
  double scale = 0.0;
  for (auto &r : rows) scale = std::max(scale, nrm2(r));
  if (!(scale > 0.0)) {
    err = "kernel projector: the risk measure has E = F = 0";
    return false;
  }
  std::vector<std::vector<double>> Q;
  std::vector<bool> used(mr, false);
  for (int it = 0; it < mr; ++it) {
    int best = -1;
    double bn = 0.0;
    for (int i = 0; i < mr; ++i)
      if (!used[i]) {
        const double ni = nrm2(rows[i]);
        if (ni > bn) { bn = ni; best = i; }
      }
    if (best < 0 || bn <= 1e-12 * scale) break;
    used[best] = true;
    std::vector<double> q = rows[best];
    for (int pass = 0; pass < 2; ++pass)  // re-orthogonalise against the accepted basis
      for (auto &b : Q) {
        double s = 0.0;
        for (int k = 0; k < nls; ++k) s += b[k] * q[k];
        for (int k = 0; k < nls; ++k) q[k] -= s * b[k];
      }
    const double nq = nrm2(q);
    if (nq <= 1e-12 * scale) continue;
    for (double &x : q) x /= nq;
    for (int i = 0; i < mr; ++i)  // deflate the remaining rows
      if (!used[i]) {
        double s = 0.0;
        for (int k = 0; k < nls; ++k) s += q[k] * rows[i][k];
        for (int k = 0; k < nls; ++k) rows[i][k] -= s * q[k];
      }
    Q.push_back(q);
  }
  ls.assign((size_t)nls * nls, 0.0);
  for (int i = 0; i < nls; ++i) ls[i + (size_t)i * nls] = 1.0;
  for (auto &q : Q)
    for (int j = 0; j < nls; ++j)
      for (int i = 0; i < nls; ++i) ls[i + (size_t)j * nls] -= q[i] * q[j];
  return true;
}

// ------------------------------------------------------------------------------------------------ tensor-pipe fragment tables
#include "spk_mm_layout.h"

namespace {
// W(n, k) accessor of a column-major rows x cols matrix; zero outside
struct WView {
  const double *p;
  int rows, cols, ld;
  bool transposed;  // W(n, k) = p[k + n * ld] instead of p[n + k * ld]
  double at(int n, int k) const {
    if (n >= rows || k >= cols) return 0.0;
    return transposed ? p[k + (long)n * ld] : p[n + (long)k * ld];
  }
};
void put_natural(double *frag, const WView &W, int s, int nb, int steps) {
  for (int T = 0; T < 32; ++T) frag[T] = W.at(8 * nb + T / 4, spk_mm_kidx(s, T % 4, steps));
}
void put_permuted(double *frag, const WView &W, int ks, int nb) {
  for (int T = 0; T < 32; ++T) frag[T] = W.at(8 * nb + T / 4, 8 * (ks / 2) + 2 * (T % 4) + (ks % 2));
}
}  // namespace

void spk_mm_fragment_tables(const SpkFactorTable &T, const double *B, std::vector<double> &bwd, std::vector<double> &fwd) {
  const int N = T.N, d = T.d, nx = T.nx, nu = T.nu;
  const SpkMmLayout L = spk_mm_layout(nx, nu, d);
  bwd.assign((size_t)(N - 1) * L.bwd_frags * 32, 0.0);
  fwd.assign((size_t)(N - 1) * L.fwd_frags * 32, 0.0);
  for (int t = 0; t < N - 1; ++t) {
    double *fb = bwd.data() + (size_t)t * L.bwd_frags * 32, *ff = fwd.data() + (size_t)t * L.fwd_frags * 32;
    const WView Rinv{T.Rinv.data() + (size_t)t * T.r_stride, nu, nu, nu, false};
    const WView G{T.G.data() + (size_t)t * T.k_stride, nx, nu, nx, false};
    const WView Kt{T.Kt.data() + (size_t)t * T.k_stride, nx, nu, nx, false};
    const WView K{T.K.data() + (size_t)t * T.k_stride, nu, nx, nu, false};
    for (int c = 0; c < d; ++c) {
      const int w = spk_child_mode(d, c);
      const double *Bw = B + (size_t)w * nx * nu;                       // nx x nu column-major
      const WView Bt{Bw, nu, nx, nx, true};                             // Bt(n, k) = B(k, n)
      const WView Bn{Bw, nx, nu, nx, false};
      const WView ABKt{T.ABKt.data() + (size_t)(t * d + c) * T.abk_stride, nx, nx, nx, false};
      const WView ABK{T.ABK.data() + (size_t)(t * d + c) * T.abk_stride, nx, nx, nx, false};
      for (int s = 0; s < L.KX; ++s) {
        double *base = fb + (size_t)(L.bwd_sweep + (c * L.KX + s) * (L.NBU + L.NBX)) * 32;
        for (int nb = 0; nb < L.NBU; ++nb) put_natural(base + (size_t)nb * 32, Bt, s, nb, L.KX);
        for (int nb = 0; nb < L.NBX; ++nb) put_natural(base + (size_t)(L.NBU + nb) * 32, ABKt, s, nb, L.KX);
        for (int nb = 0; nb < L.NBX; ++nb) put_natural(ff + (size_t)(L.fwd_abk + (c * L.KX + s) * L.NBX + nb) * 32, ABK, s, nb, L.KX);
      }
      for (int s = 0; s < L.KU; ++s)
        for (int nb = 0; nb < L.NBX; ++nb) put_natural(ff + (size_t)(L.fwd_b + (c * L.KU + s) * L.NBX + nb) * 32, Bn, s, nb, L.KU);
    }
    for (int ks = 0; ks < L.KUP; ++ks) {
      for (int nb = 0; nb < L.NBU; ++nb) put_permuted(fb + (size_t)(L.bwd_s2 + ks * L.NBU + nb) * 32, Rinv, ks, nb);
      for (int nb = 0; nb < L.NBX; ++nb) {
        put_permuted(fb + (size_t)(L.bwd_s3b + ks * L.NBX + nb) * 32, G, ks, nb);
        put_permuted(fb + (size_t)(L.bwd_s3c + ks * L.NBX + nb) * 32, Kt, ks, nb);
      }
    }
    for (int s = 0; s < L.KX; ++s)
      for (int nb = 0; nb < L.NBU; ++nb) put_natural(ff + (size_t)(L.fwd_k + s * L.NBU + nb) * 32, K, s, nb, L.KX);
  }
}
