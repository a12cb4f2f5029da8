// spk_offline.h -- offline factors of the hot path on the HOST in fp64: the cross-check (SPK_HOST_OFFLINE=1) of the device
// versions in spk_offline_dev.cuh, which spk_create uses by default.
//
// Replaces `ricatti_offline` (src/model_dynamics/implicit_l.jl:46-97) and the kernel-projector construction
// (src/models/cpock.jl:71-81,125-128).  On the trees the reference can generate every node of a stage has the
// same children modes, so the factors depend on (stage, child ordinal) only: they are computed per stage
// (O(N nx^3)) and, in SPK_FACTORS_PER_NODE mode, replicated to per-node slots on the device.
#pragma once
#include <string>
#include <vector>

struct SpkFactorTable {
  int N = 0, d = 0, nx = 0, nu = 0;
  long abk_stride = 0, k_stride = 0, r_stride = 0;  // padded slot sizes in doubles (even => 16-byte aligned)
  std::vector<double> P;     // N slots of nx*nx (unpadded): P of a node at level t (level N-1: identity)
  std::vector<double> ABK;   // (N-1)*d slots, slot t*d + c: child ordinal c of a parent at level t
  std::vector<double> ABKt;  // transposes
  std::vector<double> K;     // N-1 slots, nu*nx column-major
  std::vector<double> Kt;    // nx*nu column-major
  std::vector<double> G;     // nx*nu column-major: sum_c ABK_c' P_{t+1} B_w(c)
  std::vector<double> Rinv;  // nu*nu: (I + sym(sum_c B_w' P_{t+1} B_w))^{-1}
};

// mode (0-based) of the child with ordinal c: Julia w(j) = (j % d) + 1 and j = 2 + c (mod d)
inline int spk_child_mode(int d, int c) { return (c + 2) % d; }

// dimensions, padded strides and zero-filled vectors of the table (the device version of the recursion fills them by D2H)
void spk_table_init(int N, int d, int nx, int nu, SpkFactorTable &out);
bool spk_offline_factors(int N, int d, int nx, int nu, const double *A, const double *B, SpkFactorTable &out,
                         std::string &err);

// Orthogonal projector on ker [E' -I -I; F' 0 0], (ny+2d)^2 column-major
bool spk_kernel_projector(int ny, int d, const double *E, const double *F, std::vector<double> &ls, std::string &err);

// Fragment-ordered weight tables of the tensor-pipe S1 kernels (spk_mm_layout.h): level t occupies
// bwd[t * bwd_frags * 32 ...] and fwd[t * fwd_frags * 32 ...].
void spk_mm_fragment_tables(const SpkFactorTable &T, const double *B, std::vector<double> &bwd, std::vector<double> &fwd);
