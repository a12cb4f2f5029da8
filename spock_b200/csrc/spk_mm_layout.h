// spk_mm_layout.h -- fragment tables of the per-stage ("factor table") S1 sweeps on the fp64 tensor pipe.
//
// With SPK_FACTORS_PER_STAGE every family of a tree level shares the same Riccati factors, so the backward / forward
// step of projection_S1! (implicit_l.jl:584-637, :667-681) over a level is a chain of small dense mat-mats
//   OUT^T (families x rows) = IN^T (families x inner) * W^T
// that spk_mm.cuh runs with mma.sync.m8n8k4.f64 (DMMA): M = 8 families, N = 8 output rows, K = 4 inner entries.
// The weights are stored per level in FRAGMENT ORDER -- 32 consecutive doubles per (k-step, n-block), entry T being
// the B-operand register of lane T -- so that a CTA fetches the whole level with one bulk copy and every operand read
// is a conflict-free 256-byte shared-memory access.
//   "natural" k-step s: lane T holds W[8 nb + T/4][spk_mm_kidx(s, T%4, steps)] -- steps are PAIRED: steps 2S and 2S+1
//                       hold the inner entries 8S + 2(T%4) and 8S + 2(T%4) + 1, so that a lane fetches the A operands of
//                       both steps with one 16-byte load of two consecutive vector entries (an odd last step holds
//                       4 s + T%4)
//   permuted k-step ks: lane T holds W[8 nb + T/4][8 (ks/2) + 2 (T%4) + (ks%2)]   -- the inner index order in which
//                       the accumulator registers of a previous DMMA (rows 2(T%4), 2(T%4)+1 of n-block ks/2) can be fed
//                       back as the A operand without any shuffle.
// Rows / columns beyond the matrix are zero.
#pragma once
#if defined(__CUDACC__)
#define SPK_MM_HD __host__ __device__ constexpr
#else
#define SPK_MM_HD constexpr
#endif

// inner index held by lane position kk (= T%4) at k-step s of a matrix with `steps` k-steps
SPK_MM_HD int spk_mm_kidx(int s, int kk, int steps) {
  return ((steps & 1) && s == steps - 1) ? 4 * s + kk : 8 * (s >> 1) + 2 * kk + (s & 1);
}

struct SpkMmLayout {
  int KX, KU, KUP, NBX, NBU;          // k-steps over nx (natural), nu (natural), nu (permuted); n-blocks of nx / nu rows
  int bwd_sweep, bwd_s2, bwd_s3b, bwd_s3c, bwd_frags;  // fragment offsets inside a level's backward table
  int fwd_k, fwd_abk, fwd_b, fwd_frags;                // ... forward table
};
SPK_MM_HD SpkMmLayout spk_mm_layout(int nx, int nu, int d) {
  SpkMmLayout L = {};
  L.KX = (nx + 3) / 4; L.KU = (nu + 3) / 4; L.NBX = (nx + 7) / 8; L.NBU = (nu + 7) / 8; L.KUP = 2 * L.NBU;
  // backward: sweep over (c, s): [Bt_c -> NBU frags | ABKt_c -> NBX frags]; then Rinv, G, Kt with permuted k
  L.bwd_sweep = 0;
  L.bwd_s2 = d * L.KX * (L.NBU + L.NBX);
  L.bwd_s3b = L.bwd_s2 + L.KUP * L.NBU;
  L.bwd_s3c = L.bwd_s3b + L.KUP * L.NBX;
  L.bwd_frags = L.bwd_s3c + L.KUP * L.NBX;
  // forward: K (s, nb<NBU); ABK_c (c, s, nb<NBX); B_c (c, su, nb<NBX)
  L.fwd_k = 0;
  L.fwd_abk = L.KX * L.NBU;
  L.fwd_b = L.fwd_abk + d * L.KX * L.NBX;
  L.fwd_frags = L.fwd_b + d * L.KU * L.NBX;
  return L;
}
