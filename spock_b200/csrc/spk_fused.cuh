// spk_fused.cuh -- the fused hot path: one operator evaluation T(z,v) = (zbar, vbar)
//
//   zbar = prox_f(z - gamma L'v),  vbar = prox_{sigma h*}(v + sigma L(2 zbar - z))      (cp.jl:5-32, sp.jl:5-32)
//
// Unit of work = one FAMILY: a non-leaf node i, its d children and the dual blocks that only they touch.
//   1. spk_k_bwd, one launch per level N-2 .. 0: forms (z - gamma L'v) for x_i, u_i on the fly (L_transpose! :320-449) and
//      does the backward step of projection_S1! (:584-637)                                   -> q, d
//   2. spk_k_fwd_s1, one launch per level 0 .. N-2: u_i, x_children (:667-681)                -> zbar (x, u)
//   3. spk_k_s2t, one launch: (z - gamma L'v) on (y_i, s_children, tau_children) and projection_S2! (:688-727),
//      prox_f!'s s_1 -= gamma (:741)                                                          -> zbar (s, tau, y)
//   4. the dual kernel, one launch for the families with non-leaf children and one for those with leaf children:
//      L(2 zbar - z) (:177-318) and prox_h_conj! (:911-951) for v1_i, v2_i, v7_i, (v3..v6)_children and, for leaf
//      children, (v11..v14); plus the residuals of should_terminate! (cp.jl:54-80) xi1 = ||L'dv - dz/gamma||_inf,
//      xi2 = ||L dz - dv/sigma||_inf and <r, P r> (spock_dot :485-516), which are all family-local
//                                                                                             -> vbar, xi1, xi2, pnorm
//      Three implementations of the same arithmetic: spk_k_dual3 (register-resident, one lane per column; specialised
//      shapes with Diagonal costs -- the production kernel), spk_k_dual2 (batched through shared memory; dense costs and
//      unspecialised shapes), spk_k_dual (one warp per family; fallback when a batch does not fit in shared memory).
// z and v are read twice per pass, zbar and vbar written once, q and d written and read once.
//
// Kernels 1 and 2 stream the per-node Riccati factors: one CTA per family at a time, the family's factor slots are
// contiguous in HBM and are fetched with TMA bulk copies (cp.async.bulk -> UBLKCP) into shared memory, completing on
// mbarriers (large shapes: in chunks that are re-fetched for the next family as soon as they are consumed), while the
// family's vector entries are staged one family ahead with cp.async; the mat-vecs run out of shared memory with the inner
// dimension split over the CTA and only 3-6 CTA barriers per family, so that the time a CTA spends NOT waiting for HBM is
// short (tools/microbench/stream_bench.cu: single-shot bulk copies of 20-60 KB reach the HBM roofline as long as the
// consumer is quick).  With per-stage factors the matrices are fetched once per CTA; the wide levels then run the
// tensor-pipe kernels of spk_mm.cuh instead.
// Every kernel returns immediately when the loop's device-side `done` flag (SpkFusedArgs.gate, spk_sp.cuh) is set.
#pragma once
#include "spk_types.h"

#define SPK_FUSED_MAX_D 8
#ifndef SPK_DUAL_MINB
#define SPK_DUAL_MINB 4
#endif

// ------------------------------------------------------------------------------------------------ TMA / mbarrier

#if !SPK_EMU
SPK_D uint32_t spk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SPK_D void spk_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spk_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
SPK_D void spk_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spk_smem_u32(bar)), "r"(bytes) : "memory");
}
SPK_D void spk_bulk_g2s(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   spk_smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(spk_smem_u32(bar))
               : "memory");
}
SPK_D void spk_mbar_wait(uint64_t *bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SPK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SPK_DONE_%=;\n"
      "bra SPK_WAIT_%=;\n"
      "SPK_DONE_%=:\n"
      "}\n" ::"r"(spk_smem_u32(bar)),
      "r"(phase)
      : "memory");
}
#else
SPK_D void spk_mbar_init(uint64_t *, uint32_t) {}
SPK_D void spk_mbar_expect_tx(uint64_t *, uint32_t) {}
SPK_D void spk_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
SPK_D void spk_mbar_wait(uint64_t *, uint32_t) {}
#endif

SPK_D double spk_warp_sum(double v) {
#if !SPK_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
SPK_D double spk_warp_max(double v) {
#if !SPK_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = spk_max_nn(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
#endif
  return v;
}
// Programmatic dependent launch (sm_90+): consecutive level launches of a sweep are chained so that the NEXT level's CTAs
// may become resident, set up their barriers and issue the TMA fetch of their first family's factors (read-only data) while
// the current level is still finishing; everything that depends on the previous level comes after spk_pdl_wait().
// Both are no-ops for a kernel launched without the programmatic-serialization attribute.
#if !SPK_EMU
SPK_D void spk_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SPK_D void spk_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
SPK_D void spk_pdl_trigger() {}
SPK_D void spk_pdl_wait() {}
#endif

// Top-of-tree kernels (spk_k_bwd_top / spk_k_fwd_top): all levels of a sweep that are narrower than one resident wave run
// in ONE launch, a software grid barrier between levels instead of a kernel boundary (a level of a small tree costs a
// few microseconds of work; 2(N-1) launches per evaluation dominated every latency-bound configuration).
struct SpkTopPlan {
  int n;                                  // level steps of this launch, in execution order
  int t[SPK_MAX_LEVELS], leaves[SPK_MAX_LEVELS];
  long p_lo[SPK_MAX_LEVELS], p_cnt[SPK_MAX_LEVELS], kshift[SPK_MAX_LEVELS];
  unsigned *bar;                          // arrival counter, zeroed by the evaluation's first kernel (spk_k_s2t)
};
#if !SPK_EMU
SPK_D unsigned spk_ld_acquire_gpu_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all CTAs of the grid are co-resident (grid <= one resident wave); target counts arrivals since the counter was zeroed
SPK_D void spk_grid_barrier(unsigned *ctr, unsigned &target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (spk_ld_acquire_gpu_u32(ctr) < target) {}
  }
  __syncthreads();
}
SPK_D void spk_mbar_inval(uint64_t *bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(spk_smem_u32(bar)) : "memory"); }
SPK_D double spk_ldcg(const double *p) { return __ldcg(p); }
#else
SPK_D void spk_grid_barrier(unsigned *, unsigned &) {}
SPK_D void spk_mbar_inval(uint64_t *) {}
SPK_D double spk_ldcg(const double *p) { return *p; }
#endif

// 8-byte asynchronous global->shared copies (LDGSTS); the reference layout guarantees no more than 8-byte alignment
#if !SPK_EMU
SPK_D void spk_cp_async8(double *dst_smem, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(spk_smem_u32(dst_smem)), "l"(src) : "memory");
}
SPK_D void spk_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
SPK_D void spk_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SPK_D void spk_cp_async_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }  // all but the newest group
#else
SPK_D void spk_cp_async8(double *dst, const double *src) { *dst = *src; }
SPK_D void spk_cp_async_wait_all() {}
SPK_D void spk_cp_async_commit() {}
SPK_D void spk_cp_async_wait_prev() {}
#endif

// max of non-negative doubles: the IEEE bit pattern is monotone, so an integer atomicMax is exact and order-free
SPK_D void spk_atomic_max_nonneg(double *addr, double val) {
#if !SPK_EMU
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(val));
#else
  *addr = spk_max_nn(*addr, val);
#endif
}

// ------------------------------------------------------------------------------------------------ launch parameters

struct SpkFusedArgs {
  int t;                   // level of the parents handled by this launch (kernels 1, 2)
  int children_are_leaves; // t == N-2
  double gamma, sigma;
  double inv_gamma, inv_sigma;  // reciprocals, computed once on the host
  const double *z, *v;     // input iterate (reference layout)
  double *zbar, *vbar;     // output iterate
  double *q, *dd;          // sweep scratch: q (n*nx), d (nnl*nu)
  double *xi;              // xi[0] = xi1, xi[1] = xi2 (atomic max; reset by the first backward launch of the pass)
  double *pnorm;           // per-family partial of <r, P r> (nnl doubles), or NULL
  int reset_xi;            // this launch zeroes xi[0..1] (block 0)
  int vec2;                // tensor-pipe kernels: every pair of consecutive entries they touch is 16-byte aligned
  unsigned *topbar;        // 3 grid-barrier counters of the top-of-tree kernels (zeroed by spk_k_s2t), or NULL
  const int *gate;         // device-side "done" flag of the loop (spk_sp.cuh): every kernel returns at once when it is set
  // ownership (multi-GPU subtree partition; a single GPU owns everything):
  long p_lo, p_cnt;        // kernels 1, 2: positions [p_lo, p_lo + p_cnt) of level t are processed by this rank
  long kshift;             // per-node factors: family i0 of this launch reads factor slot i0 - kshift (a rank stores only its
                           // own families' factors, packed level by level; 0 on a single GPU)
  int n_ranges;            // kernels 3, 4: the owned families are n_ranges contiguous node ranges
  long r_start[SPK_MAX_LEVELS], r_prefix[SPK_MAX_LEVELS + 1];  // range k = nodes r_start[k] .. , flat offsets r_prefix
};

// flat owned-family index -> node number
SPK_HD long spk_owned_family(const SpkFusedArgs &A, long f) {
  int k = 0;
  while (k + 1 < A.n_ranges && f >= A.r_prefix[k + 1]) ++k;
  return A.r_start[k] + (f - A.r_prefix[k]);
}

SPK_HD int spk_slices(int T, int rows) { return T / rows > 1 ? T / rows : 1; }

// sum_{m in [m0, m1)} Mr[m * ld] * x[m]  -- Mr points at the row's first element; 4-way unrolled, two accumulators
SPK_D double spk_row_dot(const double *__restrict__ Mr, int ld, const double *__restrict__ x, int m0, int m1) {
  double a0 = 0.0, a1 = 0.0;
  const double *Mp = Mr + (long)m0 * ld;
  int m = m0;
  for (; m + 4 <= m1; m += 4) {
    a0 += Mp[0] * x[m];
    a1 += Mp[ld] * x[m + 1];
    a0 += Mp[2 * ld] * x[m + 2];
    a1 += Mp[3 * ld] * x[m + 3];
    Mp += 4 * ld;
  }
  for (; m < m1; ++m) {
    a0 += Mp[0] * x[m];
    Mp += ld;
  }
  return a0 + a1;
}
// chunk [m0, m1) of slice sl out of S over n columns
SPK_D void spk_chunk(int n, int S, int sl, int &m0, int &m1) {
  const int per = (n + S - 1) / S;
  m0 = sl * per;
  m1 = m0 + per < n ? m0 + per : n;
  if (m0 > n) m0 = n;
}


// compile-time-length variant: PER columns starting at m0, clipped at n; fully unrolled, offsets become immediates
template <int PER>
SPK_D double spk_dot_fixed(const double *__restrict__ Mr, int ld, const double *__restrict__ x, int m0, int n) {
  double a0 = 0.0, a1 = 0.0;
  const double *Mp = Mr + (long)m0 * ld;
  const double *xp = x + m0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (m0 + j < n) {
      if (j & 1) a1 += Mp[j * ld] * xp[j];
      else a0 += Mp[j * ld] * xp[j];
    }
  }
  return a0 + a1;
}
// PER > 0: specialised shape (slice length known at compile time), PER == 0: runtime sizes
template <int PER>
SPK_D double spk_dot_any(const double *Mr, int ld, const double *x, int m0, int m1, int n) {
  if constexpr (PER > 0) return spk_dot_fixed<PER>(Mr, ld, x, m0, n);
  else return spk_row_dot(Mr, ld, x, m0, m1);
}
SPK_CONSTEXPR_HD int spk_cslices(int T, int rows) { return (T > 0 && rows > 0) ? (T / rows > 1 ? T / rows : 1) : 0; }
SPK_CONSTEXPR_HD int spk_cper(int n, int S) { return S > 0 ? (n + S - 1) / S : 0; }

// ------------------------------------------------------------------------------------------------ 1. backward S1

// Work item of a CTA-wide mat-vec: output row r, slice sl of the inner dimension, i.e. columns [m0, m1) of an nx-wide
// and [k0, k1) of an nu-wide operand.  The items of a thread do not depend on the family, so they are computed once
// per kernel (runtime integer divisions are the most expensive instructions of these kernels).
struct SpkItem {
  int r, m0, m1, k0, k1;
};
SPK_D SpkItem spk_make_item(int o, int rows, int S, int nx, int nu) {
  SpkItem it;
  it.r = o % rows;
  const int sl = o / rows;
  spk_chunk(nx, S, sl, it.m0, it.m1);
  spk_chunk(nu, S, sl, it.k0, it.k1);
  return it;
}
SPK_D int spk_child_mode_dev(int d, int c) {  // mode of child ordinal c: (j0 + 1) % d with j0 = d*i0 + 1 + c
  int m = c + 2;
  while (m >= d) m -= d;
  return m;
}
// W: a caller-supplied mode map (spk_problem.mode_of_node) replaces the generator's; child c of the family whose first
// child is cfirst.  A template parameter and not a run-time test of P.w: the instantiations for the generator's map stay
// exactly the kernels that were tuned and verified, and the map costs them nothing (with a run-time branch here the
// specialised spk_k_bwd<10,10,4,64> faulted on the GENERATOR's map -- invalid address space, every d = 4 case -- while the
// runtime-size instantiation and the mode-map cases ran).  Only the runtime-size kernels are instantiated with W.
template <bool W>
SPK_D int spk_mode_c(const SpkPtrs &P, int d, long cfirst, int c) {
  if constexpr (W) return P.w[cfirst + c];
  else return spk_child_mode_dev(d, c);
}

SPK_HD long spk_bwd_smem_doubles(const SpkDims &D, int T) {
  const long SA = spk_slices(T, D.nx), SB = spk_slices(T, D.nu), nw = (T + 31) / 32;
  const long raw = 2L * D.nu + 2L * D.nx + (long)D.d * D.nu + 4L * D.d * D.nx;  // zu, zx, v7, v4c, v3c, qraw, v14c, v11c
  // (partial sums of the ABK' mat-vec: up to 2 SA + 1 slices -- the row-pair mapping of the large shapes slices finer)
  return 2 + (long)D.d * D.abk_stride + 2 * D.k_stride + D.r_stride + (long)D.d * D.nx + D.nu + D.nx + (2 * SA + 1) * D.nx + SB * D.nu +
         nw * 3L * D.nu + raw + 4;
}

// NX_, NU_, D_, T_ > 0: sizes fixed at compile time (every loop unrolls, index arithmetic folds); 0: runtime sizes
// TOP: one level step inside a top-of-tree kernel -- the dependency on the previous step is a grid barrier instead of the
// kernel boundary, and what other CTAs wrote in the previous step is read past the (incoherent) L1
template <int NX_, int NU_, int D_, int T_, bool TOP, bool W>
SPK_D void spk_bwd_body(const SpkDims &D, const SpkPtrs &P, const SpkFusedArgs &A, double *smem, unsigned *gbar, unsigned &gtarget,
                        bool first_step) {
  const int nx = NX_ > 0 ? NX_ : D.nx, nu = NX_ > 0 ? NU_ : D.nu, d = NX_ > 0 ? D_ : D.d, T = NX_ > 0 ? T_ : (int)blockDim.x;
  const int tid = threadIdx.x, t = A.t;
  const int WL = T < 32 ? T : 32, warp = tid / WL, lane = tid % WL;
  const int SA = spk_slices(T, nx), SB = spk_slices(T, nu);
  constexpr int cSA = spk_cslices(T_, NX_), cSB = spk_cslices(T_, NU_);
  constexpr int PER_AX = spk_cper(NX_, cSA), PER_AU = spk_cper(NU_, cSA), PER_BX = spk_cper(NX_, cSB);
  const long count = A.p_cnt, first0 = spk_level_first0(d, t) + A.p_lo;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  double *mABKt = smem + 2, *mG = mABKt + (long)d * D.abk_stride, *mKt = mG + D.k_stride, *mRinv = mKt + D.k_stride;
  double *qc = mRinv + D.r_stride, *tu = qc + (long)d * nx, *tx = tu + nu, *partA = tx + nx, *partB = partA + (long)(2 * SA + 1) * nx,
         *wrhs = partB + (long)SB * nu + (long)warp * 3 * nu, *wdv = wrhs + nu, *wdmt = wdv + nu;
  // raw staging of the next family's vector entries (asynchronous copies)
  double *rzu = partB + (long)SB * nu + (long)((T + WL - 1) / WL) * 3 * nu, *rzx = rzu + nu, *rv7 = rzx + nx, *rv4 = rv7 + nx + nu,
         *rv3 = rv4 + (long)d * nu, *rq = rv3 + (long)d * nx, *rv14 = rq + (long)d * nx, *rv11 = rv14 + (long)d * nx;
  const uint32_t bytes_abk = (uint32_t)(d * D.abk_stride * sizeof(double)), bytes_k = (uint32_t)(D.k_stride * sizeof(double)),
                 bytes_r = (uint32_t)(D.r_stride * sizeof(double));
  const bool leaves = A.children_are_leaves != 0;
  const double gamma = A.gamma;
  if (tid == 0) spk_mbar_init(bar, 1);
  // family-independent work items of this thread
  const SpkItem itA = spk_make_item(tid, nx, SA, nx, nu), itB = spk_make_item(tid, nu, SB, nx, nu);
  const int qc_c = tid / nx, qc_k = tid - qc_c * nx;  // (child, component) of this thread's q entry
  SPK_SYNCTHREADS();
  uint32_t phase = 0;
  auto fetch = [&](long i0) {  // TMA: the family's factor slots are contiguous
    const long aslot = D.per_node ? (long)d * (i0 - A.kshift) : (long)t * d, kslot = D.per_node ? i0 - A.kshift : t;
    spk_mbar_expect_tx(bar, bytes_abk + 2 * bytes_k + bytes_r);
    spk_bulk_g2s(mABKt, P.ABKt + aslot * D.abk_stride, bytes_abk, bar);
    spk_bulk_g2s(mG, P.G + kslot * D.k_stride, bytes_k, bar);
    spk_bulk_g2s(mKt, P.Kt + kslot * D.k_stride, bytes_k, bar);
    spk_bulk_g2s(mRinv, P.Rinv + kslot * D.r_stride, bytes_r, bar);
  };
  // asynchronous copies of one family's vector entries into the raw staging buffer
  auto stage_raw = [&](long i0) {
    const long cf = (long)d * i0 + 1;
    for (int k = tid; k < nu; k += T) spk_cp_async8(rzu + k, A.z + D.zo[1] + i0 * nu + k);
    for (int k = tid; k < nx; k += T) spk_cp_async8(rzx + k, A.z + i0 * nx + k);
    for (int k = tid; k < nx + nu; k += T) spk_cp_async8(rv7 + k, A.v + D.vo[6] + i0 * (nx + nu) + k);
    for (int k = tid; k < d * nu; k += T) spk_cp_async8(rv4 + k, A.v + D.vo[3] + (cf - 1) * nu + k);
    for (int k = tid; k < d * nx; k += T) {
      spk_cp_async8(rv3 + k, A.v + D.vo[2] + (cf - 1) * nx + k);
      if (leaves) {
        const long lf = cf - D.nnl;
        spk_cp_async8(rq + k, A.z + cf * nx + k);
        spk_cp_async8(rv14 + k, A.v + D.vo[10] + lf * nx + k);
        spk_cp_async8(rv11 + k, A.v + D.vo[7] + lf * nx + k);
      } else if constexpr (TOP) {
        rq[k] = spk_ldcg(A.q + cf * nx + k);  // written by another CTA in the previous level step
      } else {
        spk_cp_async8(rq + k, A.q + cf * nx + k);
      }
    }
  };
  // SPLIT (large specialised shapes): the factors arrive in D_+1 chunks on their own mbarriers -- ABK_c' per child, then
  // (G, K', Rinv) -- and the ABK' mat-vec runs child by child, so that a child's matrix is re-fetched for the next family as
  // soon as it is consumed instead of after the whole family (see spk_k_fwd_s1).
  constexpr bool SPLIT = !SPK_EMU && NX_ > 0 && D_ <= 4 && NX_ * NX_ * 8 >= 16 * 1024 && NX_ * spk_cslices(T_, NX_) <= T_;
  uint64_t *cbar = reinterpret_cast<uint64_t *>(rv11 + (long)d * nx);  // SPLIT: D_ barriers in the 4 spare doubles behind the staging
  // PAIR: row-pair mapping of the ABK' mat-vec (even nx: rows 2p, 2p+1 of a column are one aligned 16-byte word)
#ifndef SPK_NO_PAIR
  constexpr bool PAIR = SPLIT && NX_ % 2 == 0;
#else
  constexpr bool PAIR = false;
#endif
  constexpr int NP = NX_ > 0 ? NX_ / 2 : 1;
  constexpr int SA2 = PAIR ? (T_ / NP < NX_ ? T_ / NP : NX_) : 1, PER2 = PAIR ? (NX_ + SA2 - 1) / SA2 : 1;
  const int pair_p = tid % NP, pair_sl = tid / NP, pair_m0 = pair_sl * PER2;
  const bool pair_on = PAIR && tid < NP * SA2;
  if constexpr (SPLIT) {
    if (tid == 0)
      for (int c = 0; c < D_; ++c) spk_mbar_init(cbar + c, 1);
    SPK_SYNCTHREADS();
  }
  auto fetch_child = [&](long i0, int c) {
    const long aslot = D.per_node ? (long)d * (i0 - A.kshift) : (long)t * d;
    spk_mbar_expect_tx(cbar + c, (uint32_t)(D.abk_stride * sizeof(double)));
    spk_bulk_g2s(mABKt + (long)c * D.abk_stride, P.ABKt + (aslot + c) * D.abk_stride, (uint32_t)(D.abk_stride * sizeof(double)), cbar + c);
  };
  auto fetch_rest = [&](long i0) {
    const long kslot = D.per_node ? i0 - A.kshift : t;
    spk_mbar_expect_tx(bar, 2 * bytes_k + bytes_r);
    spk_bulk_g2s(mG, P.G + kslot * D.k_stride, bytes_k, bar);
    spk_bulk_g2s(mKt, P.Kt + kslot * D.k_stride, bytes_k, bar);
    spk_bulk_g2s(mRinv, P.Rinv + kslot * D.r_stride, bytes_r, bar);
  };
  if ((long)blockIdx.x < count && tid == 0) {  // the factors do not depend on the previous level: fetch before the wait
    if constexpr (SPLIT) {
      for (int c = 0; c < D_; ++c) fetch_child(first0 + blockIdx.x, c);
      fetch_rest(first0 + blockIdx.x);
    } else {
      fetch(first0 + blockIdx.x);
    }
  }
  if constexpr (TOP) {
    if (!first_step) spk_grid_barrier(gbar, gtarget);  // the level below (the children's q) is complete
  } else {
    spk_pdl_wait();  // the previous launch (the level below: the children's q) has completed and its writes are visible
  }
  if (A.reset_xi && blockIdx.x == 0 && tid == 0) {
    A.xi[0] = 0.0;
    A.xi[1] = 0.0;
  }
  if ((long)blockIdx.x < count) stage_raw(first0 + blockIdx.x);
  uint32_t cph = 0;     // SPLIT: bit c = phase of child barrier c
  bool pending = true;  // a fetch is in flight / not yet consumed
  for (long p = blockIdx.x; p < count; p += gridDim.x) {
    const long i0 = first0 + p;
    const long cfirst = (long)d * i0 + 1;  // first child (0-based node), children are consecutive
    const bool fresh = D.per_node || p == (long)blockIdx.x;   // this family's factors are (still) in flight
    const bool more = D.per_node && p + gridDim.x < count;    // the next family needs its own factors
    spk_cp_async_wait_all();  // this thread's raw copies have landed
    SPK_SYNCTHREADS();
    // ---- P0: (z - gamma L'v) for u_i, x_i (L_transpose! :351-373, :399-420) and the children's q, from the raw staging
    for (int k = tid; k < nu; k += T) {
      double val = rv7[nx + k];
      for (int c = 0; c < d; ++c) val += spk_cost_row(spk_sqrtR(D, P, cfirst + c), D.cost_dense, nu, k, rv4 + c * nu);
      tu[k] = rzu[k] - gamma * val;
    }
    for (int k = tid; k < nx; k += T) {
      double val = rv7[k];
      for (int c = 0; c < d; ++c) val += spk_cost_row(spk_sqrtQ(D, P, cfirst + c), D.cost_dense, nx, k, rv3 + c * nx);
      tx[k] = rzx[k] - gamma * val;
    }
    for (int idx = tid; idx < d * nx; idx += T) {
      if (leaves) {  // q_leaf = -(x - gamma L'v)_leaf, implicit_l.jl:578-582 with :376-396
        const int c = (idx == tid) ? qc_c : idx / nx, k = (idx == tid) ? qc_k : idx - (idx / nx) * nx;
        const double val = rv14[idx] + spk_cost_row(spk_sqrtQN(D, P, cfirst + c - D.nnl), D.cost_dense, nx, k, rv11 + c * nx);
        qc[idx] = -(rq[idx] - gamma * val);
      } else {
        qc[idx] = rq[idx];
      }
    }
    SPK_SYNCTHREADS();
    if (p + gridDim.x < count) stage_raw(i0 + gridDim.x);  // the staging buffer is free: prefetch the next family's entries
    // ---- P1: partial sums of sum_c B_w' q_c (:590-596; B' is a small table that stays in L1) and of sum_c ABK_c' q_c
    for (int o = tid; o < nu * SB; o += T) {
      const SpkItem it = (o == tid) ? itB : spk_make_item(o, nu, SB, nx, nu);
      double acc = 0.0;
      for (int c = 0; c < d; ++c)
        acc += spk_dot_any<PER_BX>(P.Bt + (long)spk_mode_c<W>(P, d, cfirst, c) * nx * nu + it.r, nu, qc + c * nx, it.m0, it.m1, nx);
      partB[o] = acc;
    }
    if constexpr (PAIR) {
      // child by child, TWO output rows per thread: one 16-byte shared-memory read feeds two FMAs and the q entry is read
      // once for both (half the LDS instructions of the one-row mapping, two independent accumulator chains)
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int c = 0; c < D_; ++c) {
        if (fresh) {
          spk_mbar_wait(cbar + c, (cph >> c) & 1u);
          cph ^= 1u << c;
        }
        if (pair_on) {
          const double *Mp = mABKt + (long)c * D.abk_stride + (long)pair_m0 * NX_ + 2 * pair_p;
          const double *xq = qc + c * NX_ + pair_m0;
#pragma unroll
          for (int j = 0; j < PER2; ++j) {
            if (pair_m0 + j < NX_) {
              const double2 mm = *reinterpret_cast<const double2 *>(Mp + j * NX_);
              const double xv = xq[j];
              a0 += mm.x * xv;
              a1 += mm.y * xv;
            }
          }
          if (c == D_ - 1) {
            partA[pair_sl * NX_ + 2 * pair_p] = a0;
            partA[pair_sl * NX_ + 2 * pair_p + 1] = a1;
          }
        }
        SPK_SYNCTHREADS();  // ABK_c' is consumed (after the last child: partA, partB complete)
        if (more && tid == 0) fetch_child(i0 + gridDim.x, c);
      }
      if (fresh) {
        spk_mbar_wait(bar, phase);
        phase ^= 1;
      }
    } else if constexpr (SPLIT) {  // child by child; one work item per thread (nx * SA <= T)
      const bool activeA = tid < NX_ * cSA;
      double accA = 0.0;
#pragma unroll
      for (int c = 0; c < D_; ++c) {
        if (fresh) {
          spk_mbar_wait(cbar + c, (cph >> c) & 1u);
          cph ^= 1u << c;
        }
        if (activeA) accA += spk_dot_fixed<PER_AX>(mABKt + (long)c * D.abk_stride + itA.r, NX_, qc + c * NX_, itA.m0, NX_);
        if (c == D_ - 1 && activeA) partA[tid] = accA;
        SPK_SYNCTHREADS();  // ABK_c' is consumed (after the last child: partA, partB complete)
        if (more && tid == 0) fetch_child(i0 + gridDim.x, c);
      }
      if (fresh) {
        spk_mbar_wait(bar, phase);
        phase ^= 1;
      }
    } else {
      if (pending && fresh) {
        spk_mbar_wait(bar, phase);
        phase ^= 1;
      }
      pending = false;
      for (int o = tid; o < nx * SA; o += T) {
        const SpkItem it = (o == tid) ? itA : spk_make_item(o, nx, SA, nx, nu);
        double acc = 0.0;
        for (int c = 0; c < d; ++c)
          acc += spk_dot_any<PER_AX>(mABKt + (long)c * D.abk_stride + it.r, nx, qc + c * nx, it.m0, it.m1, nx);
        partA[o] = acc;
      }
      SPK_SYNCTHREADS();
    }
    // ---- P2+P3, redundantly in every warp (no CTA barrier): rhs (:598-600), d_i = Rtilde^{-1} rhs (:601)
    for (int k = lane; k < nu; k += WL) {
      double s2 = 0.0;
      for (int sl = 0; sl < SB; ++sl) s2 += partB[sl * nu + k];
      wrhs[k] = tu[k] - s2;
    }
    SPK_SYNCWARP();
    for (int k = lane; k < nu; k += WL) {
      const double acc = spk_dot_any<NU_>(mRinv + k, nu, wrhs, 0, nu, nu);
      wdv[k] = acc;
      wdmt[k] = acc - tu[k];
      if (warp == 0) A.dd[i0 * nu + k] = acc;
    }
    SPK_SYNCWARP();
    // ---- P4: + G d + K'(d - ubar) (:610-632)
    if constexpr (PAIR) {  // the same row-pair mapping: slice pair_sl of the nu columns of G and K'
      constexpr int PERU2 = (NU_ + SA2 - 1) / SA2;
      if (pair_on) {
        const int k0 = pair_sl * PERU2;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < PERU2; ++j) {
          if (k0 + j < NU_) {
            const double2 g2 = *reinterpret_cast<const double2 *>(mG + (long)(k0 + j) * NX_ + 2 * pair_p);
            const double2 k2 = *reinterpret_cast<const double2 *>(mKt + (long)(k0 + j) * NX_ + 2 * pair_p);
            const double dv_ = wdv[k0 + j], dm_ = wdmt[k0 + j];
            a0 += g2.x * dv_ + k2.x * dm_;
            a1 += g2.y * dv_ + k2.y * dm_;
          }
        }
        partA[pair_sl * NX_ + 2 * pair_p] += a0;
        partA[pair_sl * NX_ + 2 * pair_p + 1] += a1;
      }
    } else {
      for (int o = tid; o < nx * SA; o += T) {
        const SpkItem it = (o == tid) ? itA : spk_make_item(o, nx, SA, nx, nu);
        partA[o] += spk_dot_any<PER_AU>(mG + it.r, nx, wdv, it.k0, it.k1, nu) + spk_dot_any<PER_AU>(mKt + it.r, nx, wdmt, it.k0, it.k1, nu);
      }
    }
    SPK_SYNCTHREADS();
    // the matrices are consumed: start the next family's transfer before finishing this one
    if (more) {
      if (tid == 0) {
        if constexpr (SPLIT) fetch_rest(i0 + gridDim.x);
        else fetch(i0 + gridDim.x);
      }
      pending = true;
    }
    // ---- P5: q_i (:634-636)
    for (int r = tid; r < nx; r += T) {
      double acc = 0.0;
      for (int sl = 0; sl < (PAIR ? SA2 : SA); ++sl) acc += partA[sl * nx + r];
      A.q[i0 * nx + r] = acc - tx[r];
    }
    SPK_SYNCTHREADS();  // work buffers are reused by the next family
  }
  if constexpr (TOP) {  // the next level step re-initialises the barriers
    if (tid == 0) {
      spk_mbar_inval(bar);
      if constexpr (SPLIT)
        for (int c = 0; c < D_; ++c) spk_mbar_inval(cbar + c);
    }
    SPK_SYNCTHREADS();
  }
}

template <int NX_, int NU_, int D_, int T_, bool W = false>
SPK_KERNEL void spk_k_bwd(SpkDims D, SpkPtrs P, SpkFusedArgs A) {
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  spk_pdl_trigger();
  unsigned dummy = 0;
  spk_bwd_body<NX_, NU_, D_, T_, false, W>(D, P, A, smem, nullptr, dummy, true);
}

// all narrow levels of the backward sweep (plan.t descending) in one launch
template <int NX_, int NU_, int D_, int T_, bool W = false>
SPK_KERNEL void spk_k_bwd_top(SpkDims D, SpkPtrs P, SpkFusedArgs A, SpkTopPlan plan) {
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  spk_pdl_wait();
  unsigned target = 0;
  for (int s = 0; s < plan.n; ++s) {
    SpkFusedArgs B = A;
    B.t = plan.t[s]; B.children_are_leaves = plan.leaves[s]; B.p_lo = plan.p_lo[s]; B.p_cnt = plan.p_cnt[s]; B.kshift = plan.kshift[s];
    B.reset_xi = (s == 0) ? A.reset_xi : 0;
    spk_bwd_body<NX_, NU_, D_, T_, true, W>(D, P, B, smem, plan.bar, target, s == 0);
  }
}

// ------------------------------------------------------------------------------------------------ 2. forward S1

SPK_HD long spk_fwd_smem_doubles(const SpkDims &D, int T) {
  const long R = D.nu + (long)D.d * D.nx, S = spk_slices(T, (int)R);
  return 2 + (long)D.d * D.abk_stride + D.k_stride + 2L * (D.nx + D.nu) + R * S + 4;
}

template <int NX_, int NU_, int D_, int T_, bool TOP, bool W>
SPK_D void spk_fwd_body(const SpkDims &D, const SpkPtrs &P, const SpkFusedArgs &A, double *smem, unsigned *gbar, unsigned &gtarget,
                        bool first_step) {
  const int nx = NX_ > 0 ? NX_ : D.nx, nu = NX_ > 0 ? NU_ : D.nu, d = NX_ > 0 ? D_ : D.d, T = NX_ > 0 ? T_ : (int)blockDim.x;
  const int tid = threadIdx.x, t = A.t;
  const int R = nu + d * nx, S = spk_slices(T, R);
  constexpr int cS = spk_cslices(T_, NU_ + D_ * NX_);
  constexpr int PER_X = spk_cper(NX_, cS), PER_U = spk_cper(NU_, cS);
  const long count = A.p_cnt, first0 = spk_level_first0(d, t) + A.p_lo;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  double *mABK = smem + 2, *mK = mABK + (long)d * D.abk_stride, *raw0 = mK + D.k_stride, *part = raw0 + 2L * (nx + nu);
  const uint32_t bytes_abk = (uint32_t)(d * D.abk_stride * sizeof(double)), bytes_k = (uint32_t)(D.k_stride * sizeof(double));
  if (tid == 0) spk_mbar_init(bar, 1);
  // family-independent work item of this thread: row r of [K; ABK_0; ..; ABK_{d-1}] and an inner-dimension slice
  const SpkItem it0 = spk_make_item(tid, R, S, nx, nu);
  const int c0 = it0.r < nu ? 0 : (it0.r - nu) / nx, rr0 = it0.r < nu ? it0.r : (it0.r - nu) - c0 * nx;
  // SPLIT (specialised shapes with two inner slices): the factors arrive as two column halves on two mbarriers and the
  // mat-vec runs half by half, so that a half is re-fetched for the next family while the other half is still being
  // consumed -- ncu showed 57 % of the stall samples of the single-buffer version on the wait for the whole fetch.
  constexpr int HC = spk_cper(NX_, 2);  // columns of the first half
  constexpr bool SPLIT = !SPK_EMU && NX_ > 0 && cS == 2 && (HC * NX_) % 2 == 0 && (HC * NU_) % 2 == 0 &&
                         ((NX_ - HC) * NX_) % 2 == 0 && ((NX_ - HC) * NU_) % 2 == 0 &&
                         NX_ * NX_ * 8 >= 16 * 1024;  // measured: no gain (slightly slower) with a few KB per family
  // (a row-pair / 16-byte-read mapping like the backward kernel's was measured neutral here: this kernel is DRAM-bound)
  if (SPLIT && tid == 0) spk_mbar_init(bar + 1, 1);
  SPK_SYNCTHREADS();
  uint32_t phase = 0;
  auto fetch = [&](long i0) {
    const long aslot = D.per_node ? (long)d * (i0 - A.kshift) : (long)t * d, kslot = D.per_node ? i0 - A.kshift : t;
    spk_mbar_expect_tx(bar, bytes_abk + bytes_k);
    spk_bulk_g2s(mABK, P.ABK + aslot * D.abk_stride, bytes_abk, bar);
    spk_bulk_g2s(mK, P.K + kslot * D.k_stride, bytes_k, bar);
  };
  // x_i (:667: x0 at the root) and d_i of a family, asynchronously into one of the two small staging buffers
  auto stage_raw = [&](long i0, int buf) {
    double *dst = raw0 + (long)buf * (nx + nu);
    if constexpr (TOP) {  // x_i was written by another CTA in the previous level step: read it past L1
      for (int k = tid; k < nx; k += T) dst[k] = (t == 0) ? P.x0[k] : spk_ldcg(A.zbar + i0 * nx + k);
    } else {
      for (int k = tid; k < nx; k += T) spk_cp_async8(dst + k, (t == 0) ? P.x0 + k : A.zbar + i0 * nx + k);
    }
    for (int k = tid; k < nu; k += T) spk_cp_async8(dst + nx + k, A.dd + i0 * nu + k);
  };
  auto fetch_half = [&](long i0, int hf) {  // SPLIT: columns [0, HC) or [HC, nx) of ABK_0..ABK_{d-1} and K
    const long aslot = D.per_node ? (long)d * (i0 - A.kshift) : (long)t * d, kslot = D.per_node ? i0 - A.kshift : t;
    const int c0h = hf ? HC : 0, nc = hf ? NX_ - HC : HC;
    uint64_t *br = bar + hf;
    spk_mbar_expect_tx(br, (uint32_t)((D_ * NX_ + NU_) * nc * sizeof(double)));
    for (int c = 0; c < D_; ++c)
      spk_bulk_g2s(mABK + (long)c * D.abk_stride + c0h * NX_, P.ABK + (aslot + c) * D.abk_stride + c0h * NX_,
                   (uint32_t)(nc * NX_ * sizeof(double)), br);
    spk_bulk_g2s(mK + c0h * NU_, P.K + kslot * D.k_stride + c0h * NU_, (uint32_t)(nc * NU_ * sizeof(double)), br);
  };
  if ((long)blockIdx.x < count && tid == 0) {  // the factors do not depend on the previous level: fetch before the wait
    if constexpr (SPLIT) {
      fetch_half(first0 + blockIdx.x, 0);
      fetch_half(first0 + blockIdx.x, 1);
    } else {
      fetch(first0 + blockIdx.x);
    }
  }
  if constexpr (TOP) {
    if (!first_step) spk_grid_barrier(gbar, gtarget);  // the level above (x_i) is complete
  } else {
    spk_pdl_wait();  // the previous launch (the level above: x_i; the backward sweep: d_i) has completed
  }
  if ((long)blockIdx.x < count) stage_raw(first0 + blockIdx.x, 0);
  spk_cp_async_commit();
  uint32_t phase1 = 0;  // SPLIT: phase of the second half's barrier
  bool pending = true;
  int buf = 0;
  for (long p = blockIdx.x; p < count; p += gridDim.x, buf ^= 1) {
    const long i0 = first0 + p;
    const long cfirst = (long)d * i0 + 1;
    const bool fresh = D.per_node || p == (long)blockIdx.x;   // this family's factors are (still) in flight
    const bool more = D.per_node && p + gridDim.x < count;    // the next family needs its own factors
    if (p + gridDim.x < count) stage_raw(i0 + gridDim.x, buf ^ 1);  // one family ahead
    spk_cp_async_commit();
    spk_cp_async_wait_prev();  // this family's entries have landed
    SPK_SYNCTHREADS();
    const double *xb = raw0 + (long)buf * (nx + nu), *ds = xb + nx;
    if (t == 0)
      for (int k = tid; k < nx; k += T) A.zbar[k] = xb[k];
    // rows 0..nu-1: K x_i (:670); rows nu + c*nx + rr: ABK_c x_i + B_w d_i (:675-679)
    if constexpr (SPLIT) {  // half by half; one work item per thread: row r, sub-slice ss of both column halves
      constexpr int PA = spk_cper(HC, 2), PB = spk_cper(NX_ - HC, 2);
      const int ss = tid / R, r = tid - ss * R;
      const bool active = tid < 2 * R;
      const int c = r < NU_ ? 0 : (r - NU_) / NX_, rr = r < NU_ ? r : (r - NU_) - c * NX_;
      if (fresh) { spk_mbar_wait(bar, phase); phase ^= 1; }
      double acc = 0.0;
      if (active) {
        if (r < NU_) acc = spk_dot_fixed<PA>(mK + rr, NU_, xb, ss * PA, HC);
        else acc = spk_dot_fixed<PA>(mABK + (long)c * D.abk_stride + rr, NX_, xb, ss * PA, HC) +
                   spk_dot_fixed<PER_U>(P.B + (long)spk_mode_c<W>(P, d, cfirst, c) * nx * nu + rr, NX_, ds, ss * PER_U, NU_);
      }
      SPK_SYNCTHREADS();  // the first half is consumed: refill it while the second half is processed
      if (more && tid == 0) fetch_half(i0 + gridDim.x, 0);
      if (fresh) { spk_mbar_wait(bar + 1, phase1); phase1 ^= 1; }
      if (active) {
        if (r < NU_) acc += spk_dot_fixed<PB>(mK + rr, NU_, xb, HC + ss * PB, NX_);
        else acc += spk_dot_fixed<PB>(mABK + (long)c * D.abk_stride + rr, NX_, xb, HC + ss * PB, NX_);
        part[tid] = acc;
      }
    } else {
      if (pending && fresh) {
        spk_mbar_wait(bar, phase);
        phase ^= 1;
      }
      pending = false;
      for (int o = tid; o < R * S; o += T) {
        SpkItem it = it0;
        int c = c0, rr = rr0;
        if (o != tid) {
          it = spk_make_item(o, R, S, nx, nu);
          c = it.r < nu ? 0 : (it.r - nu) / nx;
          rr = it.r < nu ? it.r : (it.r - nu) - c * nx;
        }
        double acc;
        if (it.r < nu) {
          acc = spk_dot_any<PER_X>(mK + rr, nu, xb, it.m0, it.m1, nx);
        } else {
          acc = spk_dot_any<PER_X>(mABK + (long)c * D.abk_stride + rr, nx, xb, it.m0, it.m1, nx) +
                spk_dot_any<PER_U>(P.B + (long)spk_mode_c<W>(P, d, cfirst, c) * nx * nu + rr, nx, ds, it.k0, it.k1, nu);
        }
        part[o] = acc;
      }
    }
    SPK_SYNCTHREADS();
    if (more) {  // matrices (SPLIT: the second half) consumed: start the next family's transfer
      if (tid == 0) {
        if constexpr (SPLIT) fetch_half(i0 + gridDim.x, 1);
        else fetch(i0 + gridDim.x);
      }
      pending = true;
    }
    for (int r = tid; r < R; r += T) {
      double acc = 0.0;
      for (int sl = 0; sl < S; ++sl) acc += part[sl * R + r];
      if (r < nu) A.zbar[D.zo[1] + i0 * nu + r] = acc + ds[r];  // :671-673
      else A.zbar[cfirst * nx + (r - nu)] = acc;
    }
    SPK_SYNCTHREADS();
  }
  if constexpr (TOP) {
    if (tid == 0) {
      spk_mbar_inval(bar);
      if (SPLIT) spk_mbar_inval(bar + 1);
    }
    SPK_SYNCTHREADS();
  }
}

template <int NX_, int NU_, int D_, int T_, bool W = false>
SPK_KERNEL void spk_k_fwd_s1(SpkDims D, SpkPtrs P, SpkFusedArgs A) {
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  spk_pdl_trigger();
  unsigned dummy = 0;
  spk_fwd_body<NX_, NU_, D_, T_, false, W>(D, P, A, smem, nullptr, dummy, true);
}

// all narrow levels of the forward sweep (plan.t ascending) in one launch
template <int NX_, int NU_, int D_, int T_, bool W = false>
SPK_KERNEL void spk_k_fwd_top(SpkDims D, SpkPtrs P, SpkFusedArgs A, SpkTopPlan plan) {
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  spk_pdl_wait();
  unsigned target = 0;
  for (int s = 0; s < plan.n; ++s) {
    SpkFusedArgs B = A;
    B.t = plan.t[s]; B.children_are_leaves = plan.leaves[s]; B.p_lo = plan.p_lo[s]; B.p_cnt = plan.p_cnt[s]; B.kshift = plan.kshift[s];
    spk_fwd_body<NX_, NU_, D_, T_, true, W>(D, P, B, smem, plan.bar, target, s == 0);
  }
}

// ------------------------------------------------------------------------------------------------ 3. S2 on (z - gamma L'v)

// One thread per non-leaf node: gathers (z - gamma L'v) for (y_i, s_children, tau_children) (L_transpose! :423-448),
// multiplies by the kernel projector (projection_S2! :688-727) and writes zbar; thread 0 also does prox_f!'s
// s_1 -= gamma (:741), the only s entry outside every S2 group.
// NY_, DD_ > 0: risk-vector length and branching fixed at compile time, so that the gathered vector stays in registers
// (with runtime sizes it is a dynamically indexed local array: 1.4 ms per pass at config 5); 0: runtime sizes.
template <int NY_, int DD_>
SPK_KERNEL void spk_k_s2t(SpkDims D, SpkPtrs P, SpkFusedArgs A) {
  if (A.gate && *A.gate) return;
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0 && A.topbar) {  // arrival counters of this evaluation's top-of-tree kernels
    A.topbar[0] = 0u; A.topbar[1] = 0u; A.topbar[2] = 0u;
  }
  if (f >= A.r_prefix[A.n_ranges]) return;
  const long i0 = spk_owned_family(A, f);
  const int ny = NY_ > 0 ? NY_ : D.ny, d = NY_ > 0 ? DD_ : D.d, nls = ny + 2 * d;
  constexpr int CAP = NY_ > 0 ? NY_ + 2 * DD_ : SPK_MAX_LS;
  const double gamma = A.gamma;
  const long cfirst = (long)d * i0 + 1;
  double in[CAP];
  const double v2i = A.v[D.vo[1] + i0];
  const double *__restrict__ bi = spk_b(D, P, i0), *__restrict__ lsi = spk_ls(D, P, i0);
  _Pragma("unroll") for (int k = 0; k < (NY_ > 0 ? NY_ : ny); ++k)
    in[k] = A.z[D.zo[4] + i0 * ny + k] - gamma * (A.v[i0 * ny + k] - bi[k] * v2i);
  _Pragma("unroll") for (int c = 0; c < (NY_ > 0 ? DD_ : d); ++c) {
    const long j0 = cfirst + c;
    const double lt = (j0 < D.nnl) ? A.v[D.vo[1] + j0] : 0.5 * A.v[D.vo[8] + (j0 - D.nnl)] + 0.5 * A.v[D.vo[9] + (j0 - D.nnl)];
    in[ny + c] = A.z[D.zo[2] + j0] - gamma * lt;
    in[ny + d + c] = A.z[D.zo[3] + j0 - 1] - gamma * (0.5 * A.v[D.vo[4] + j0 - 1] + 0.5 * A.v[D.vo[5] + j0 - 1]);
  }
  _Pragma("unroll") for (int k = 0; k < (NY_ > 0 ? CAP : nls); ++k) {
    double acc = 0.0;
    _Pragma("unroll") for (int m = 0; m < (NY_ > 0 ? CAP : nls); ++m) acc += lsi[k + (long)m * nls] * in[m];
    if (k < ny) A.zbar[D.zo[4] + i0 * ny + k] = acc;
    else if (k < ny + d) A.zbar[D.zo[2] + cfirst + (k - ny)] = acc;
    else A.zbar[D.zo[3] + cfirst + (k - ny - d) - 1] = acc;
  }
  if (i0 == 0) A.zbar[D.zo[2]] = A.z[D.zo[2]] - gamma * v2i - gamma;
}

// 3b. The same for ONE contiguous node range and one risk measure for all nodes (every single-GPU reference workload),
// staged through shared memory.  With a thread per node the direct version's gathers and scatters are strided (ny or d
// doubles per thread: a warp-level access touches 8-20 cache lines) and the projector is read entry by entry from L1:
// ncu counts ~1200 L1 wavefronts per warp, which is what bounds it (0.82 ms per pass at config 5, 22 % of the HBM rate
// for its bytes).  Here the CTA's 128 nodes are loaded and stored with coalesced accesses (their entries are contiguous
// in every segment), the combination z - gamma L'v happens on the way into shared memory (rows padded to an odd length:
// conflict-free row reads), and the projector arrives as a kernel parameter, so its entries are constant-bank operands
// of the FMAs instead of loads.  Arithmetic and its order are those of spk_k_s2t.
template <int NY_, int DD_>
struct SpkLsMat {
  double m[(NY_ + 2 * DD_) * (NY_ + 2 * DD_)];  // column-major like P.ls
};
#if !SPK_EMU
template <int NY_, int DD_>
__launch_bounds__(128)
__global__ void spk_k_s2t_st(SpkDims D, SpkPtrs P, SpkFusedArgs A, const __grid_constant__ SpkLsMat<NY_, DD_> M) {
  if (A.gate && *A.gate) return;
  constexpr int T = 128, NLS = NY_ + 2 * DD_, NYP = NY_ | 1, DP = DD_ | 1;
  __shared__ double sy[T * NYP], ss[T * DP], st[T * DP];
  const int tid = threadIdx.x;
  const long total = A.r_prefix[1], off = (long)blockIdx.x * T;
  if (off == 0 && tid == 0 && A.topbar) {  // arrival counters of this evaluation's top-of-tree kernels
    A.topbar[0] = 0u; A.topbar[1] = 0u; A.topbar[2] = 0u;
  }
  if (off >= total) return;
  const long b0 = A.r_start[0] + off, c0 = (long)DD_ * b0 + 1;  // the CTA's first node and that node's first child
  const int cnt = (int)(total - off < T ? total - off : T);
  const double gamma = A.gamma;
  const double *__restrict__ z = A.z, *__restrict__ v = A.v;
  for (int j = tid; j < cnt * NY_; j += T) {  // (z - gamma L'v) on y_i (L_transpose! :423-431)
    const int n = j / NY_, k = j - n * NY_;
    sy[n * NYP + k] = z[D.zo[4] + b0 * NY_ + j] - gamma * (v[b0 * NY_ + j] - P.b[k] * v[D.vo[1] + b0 + n]);
  }
  for (int j = tid; j < cnt * DD_; j += T) {  // ... on s_c and tau_c of the children (:433-448)
    const int n = j / DD_, c = j - n * DD_;
    const long j0 = c0 + j;
    const double lt = (j0 < D.nnl) ? v[D.vo[1] + j0] : 0.5 * v[D.vo[8] + (j0 - D.nnl)] + 0.5 * v[D.vo[9] + (j0 - D.nnl)];
    ss[n * DP + c] = z[D.zo[2] + j0] - gamma * lt;
    st[n * DP + c] = z[D.zo[3] + j0 - 1] - gamma * (0.5 * v[D.vo[4] + j0 - 1] + 0.5 * v[D.vo[5] + j0 - 1]);
  }
  __syncthreads();
  if (tid < cnt) {  // projection_S2! :688-727 for the thread's node, results back into its own rows
    double in[NLS];
#pragma unroll
    for (int k = 0; k < NY_; ++k) in[k] = sy[tid * NYP + k];
#pragma unroll
    for (int c = 0; c < DD_; ++c) {
      in[NY_ + c] = ss[tid * DP + c];
      in[NY_ + DD_ + c] = st[tid * DP + c];
    }
#pragma unroll
    for (int k = 0; k < NLS; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NLS; ++m) acc += M.m[k + m * NLS] * in[m];
      if (k < NY_) sy[tid * NYP + k] = acc;
      else if (k < NY_ + DD_) ss[tid * DP + (k - NY_)] = acc;
      else st[tid * DP + (k - NY_ - DD_)] = acc;
    }
  }
  __syncthreads();
  for (int j = tid; j < cnt * NY_; j += T) {
    const int n = j / NY_, k = j - n * NY_;
    A.zbar[D.zo[4] + b0 * NY_ + j] = sy[n * NYP + k];
  }
  for (int j = tid; j < cnt * DD_; j += T) {
    const int n = j / DD_, c = j - n * DD_;
    A.zbar[D.zo[2] + c0 + j] = ss[n * DP + c];
    A.zbar[D.zo[3] + c0 + j - 1] = st[n * DP + c];
  }
  if (b0 == 0 && tid == 0) A.zbar[D.zo[2]] = z[D.zo[2]] - gamma * v[D.vo[1]] - gamma;  // prox_f!'s s_1 -= gamma (:741)
}
#endif

// ------------------------------------------------------------------------------------------------ 4. dual blocks

SPK_D void spk_soc_coeffs(double t, double nrm, int &mode, double &tp) {  // risk_measures.jl:57-70
  if (nrm <= t) { mode = 0; tp = 0.0; }
  else if (nrm <= -t) { mode = 1; tp = 0.0; }
  else { mode = 2; tp = (t + nrm) / 2.0; }
}
SPK_D double spk_soc_apply(int mode, double tp, double nrm, double w) {
  return mode == 0 ? w : (mode == 1 ? 0.0 : tp * w / nrm);
}

// per-warp shared scratch (doubles): wx, dx (nx each), wu, du (nu), wy, dy, dv1 (ny), dv7 (nx+nu), per child:
// dv3 (nx), dv4 (nu), wxc, dxc, dv11, dv14 (nx each)
SPK_HD long spk_dual_warp_doubles(const SpkDims &D) {
  return 3L * D.nx + 3L * D.nu + 3L * D.ny + (long)D.d * (5L * D.nx + D.nu) + 8;
}

template <int NX_, int NU_, int D_>
#if !SPK_EMU
__launch_bounds__(128, SPK_DUAL_MINB)
#endif
SPK_KERNEL void spk_k_dual(SpkDims D, SpkPtrs P, SpkFusedArgs A) {
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  const int nx = NX_ > 0 ? NX_ : D.nx, nu = NX_ > 0 ? NU_ : D.nu, d = NX_ > 0 ? D_ : D.d, ny = D.ny;
  const int T = blockDim.x, tid = threadIdx.x;
  const int WL = T < 32 ? T : 32;  // lanes per warp (1 in the host-emulation test build)
  const int wpb = T / WL, warp = tid / WL, lane = tid % WL;
  const double gamma = A.gamma, sigma = A.sigma, inv_sigma = 1.0 / sigma, inv_gamma = 1.0 / gamma;
  double *W = smem + (long)warp * spk_dual_warp_doubles(D);
  double *wx = W, *dx = wx + nx, *wu = dx + nx, *du = wu + nu, *wy = du + nu, *dy = wy + ny, *dv1 = dy + ny,
         *dv7 = dv1 + ny, *dv3 = dv7 + nx + nu, *dv4 = dv3 + (long)d * nx, *wxc = dv4 + (long)d * nu,
         *dxc = wxc + (long)d * nx, *dv11 = dxc + (long)d * nx, *dv14 = dv11 + (long)d * nx;
  double my_xi1 = 0.0, my_xi2 = 0.0;
  const long nfam_stride = (long)gridDim.x * wpb;
  for (long f = (long)blockIdx.x * wpb + warp; f < A.r_prefix[A.n_ranges]; f += nfam_stride) {
    const long i0 = spk_owned_family(A, f);
    const long cfirst = (long)d * i0 + 1;
    const bool leaves = cfirst >= D.nnl;
    const long lfirst = cfirst - D.nnl;
    double my_pn = 0.0;
    // ---- w = 2 zbar - z and delta = zbar - z of the family's primal entries
    _Pragma("unroll") for (int k = lane; k < nx; k += WL) {
      const double nb = A.zbar[i0 * nx + k], old = A.z[i0 * nx + k];
      wx[k] = 2.0 * nb - old;
      dx[k] = nb - old;
    }
    _Pragma("unroll") for (int k = lane; k < nu; k += WL) {
      const double nb = A.zbar[D.zo[1] + i0 * nu + k], old = A.z[D.zo[1] + i0 * nu + k];
      wu[k] = 2.0 * nb - old;
      du[k] = nb - old;
    }
    for (int k = lane; k < ny; k += WL) {
      const double nb = A.zbar[D.zo[4] + i0 * ny + k], old = A.z[D.zo[4] + i0 * ny + k];
      wy[k] = 2.0 * nb - old;
      dy[k] = nb - old;
    }
    if (leaves)
      for (int k = lane; k < d * nx; k += WL) {
        const double nb = A.zbar[cfirst * nx + k], old = A.z[cfirst * nx + k];
        wxc[k] = 2.0 * nb - old;
        dxc[k] = nb - old;
      }
    const double sb_i = A.zbar[D.zo[2] + i0], so_i = A.z[D.zo[2] + i0];
    const double ws_i = 2.0 * sb_i - so_i, ds_i = sb_i - so_i;
    SPK_SYNCWARP();
#define SPK_FIN(w_, vold_, Ld_, pr_, gidx_, dslot_)                     \
  {                                                                     \
    const double vnew_ = sigma * ((w_) - (pr_));                        \
    A.vbar[gidx_] = vnew_;                                              \
    const double dl_ = vnew_ - (vold_);                                 \
    dslot_ = dl_;                                                       \
    my_xi2 = spk_max_nn(my_xi2, fabs((Ld_) - dl_ * inv_sigma));               \
    my_pn += dl_ * (dl_ - sigma * (Ld_));                               \
  }
    // ---- own blocks: v1_i (dual-cone projection :833-845), v2_i (:849-851), v7_i box (:903-908)
    double dv2 = 0.0;
    {
      int off = 0;
      for (int c = 0; c < D.n_cones; ++c) {
        const int dim = D.cone_dim[c], type = D.cone_type[c];
        double snrm = 0.0, tp = 0.0;
        int mode = 0;
        if (type == 4) {  // SecondOrderCone (t, x)
          double part = 0.0;
          for (int k = 1 + lane; k < dim; k += WL) {
            const double w = (A.v[i0 * ny + off + k] + sigma * wy[off + k]) / sigma;
            part += w * w;
          }
          snrm = sqrt(spk_warp_sum(part));
          spk_soc_coeffs((A.v[i0 * ny + off] + sigma * wy[off]) / sigma, snrm, mode, tp);
        }
        for (int k = lane; k < dim; k += WL) {
          const double vold = A.v[i0 * ny + off + k];
          const double w = (vold + sigma * wy[off + k]) / sigma;  // L! :207-210
          double pr;
          if (type == 0) pr = w;                         // Zeros -> dual Reals
          else if (type == 1) pr = w < 0.0 ? 0.0 : w;     // Nonnegatives
          else if (type == 2) pr = w > 0.0 ? 0.0 : w;     // Nonpositives
          else if (type == 3) pr = 0.0;                  // Reals -> dual Zeros
          else pr = (k == 0) ? (mode == 0 ? w : (mode == 1 ? 0.0 : tp)) : spk_soc_apply(mode, tp, snrm, w);
          SPK_FIN(w, vold, dy[off + k], pr, i0 * ny + off + k, dv1[off + k]);
        }
        off += dim;
      }
      double bwy = 0.0, bdy = 0.0;  // every lane computes b'wy (ny is tiny) so that no broadcast is needed
      for (int k = 0; k < ny; ++k) {
        bwy += spk_b(D, P, i0)[k] * wy[k];
        bdy += spk_b(D, P, i0)[k] * dy[k];
      }
      const double v2old = A.v[D.vo[1] + i0];
      const double w2 = (v2old + sigma * (ws_i - bwy)) / sigma;  // L! :214-222
      const double v2new = sigma * (w2 - (w2 >= 0 ? w2 : 0.0));
      dv2 = v2new - v2old;
      if (lane == 0) {
        A.vbar[D.vo[1] + i0] = v2new;
        const double l2 = ds_i - bdy;
        my_xi2 = spk_max_nn(my_xi2, fabs(l2 - dv2 * inv_sigma));
        my_pn += dv2 * (dv2 - sigma * l2);
      }
    }
    _Pragma("unroll") for (int k = lane; k < nx + nu; k += WL) {
      const long e = D.vo[6] + i0 * (nx + nu) + k;
      const double vold = A.v[e];
      const double w = (vold + sigma * (k < nx ? wx[k] : wu[k - nx])) / sigma;  // L! :275-282
      const double pr = k < nx ? spk_clip(w, D.x_min, D.x_max) : spk_clip(w, D.u_min, D.u_max);
      SPK_FIN(w, vold, (k < nx ? dx[k] : du[k - nx]), pr, e, dv7[k]);
    }
    // ---- children: SOC(nx+nu+2) on (t = v6; v3, v4, v5) (:854-900); leaf children: SOC(nx+2) on (v13; v11, v12), v14 box
    double xi1_tau = 0.0;
    _Pragma("unroll") for (int c = 0; c < d; ++c) {
      const long j0 = cfirst + c, jm = j0 - 1;
      const double *sQ = spk_sqrtQ(D, P, j0), *sR = spk_sqrtR(D, P, j0);
      const double tb = A.zbar[D.zo[3] + jm], to = A.z[D.zo[3] + jm];
      const double wtau = 2.0 * tb - to, dtau = tb - to;
      double part = 0.0;
      _Pragma("unroll") for (int k = lane; k < nx; k += WL) {  // L! :226-235
        const double w = (A.v[D.vo[2] + jm * nx + k] + sigma * spk_cost_row(sQ, D.cost_dense, nx, k, wx)) / sigma;
        dv3[c * nx + k] = w;
        part += w * w;
      }
      _Pragma("unroll") for (int k = lane; k < nu; k += WL) {  // :239-248
        const double w = (A.v[D.vo[3] + jm * nu + k] + sigma * spk_cost_row(sR, D.cost_dense, nu, k, wu)) / sigma;
        dv4[c * nu + k] = w;
        part += w * w;
      }
      const double v5old = A.v[D.vo[4] + jm], v6old = A.v[D.vo[5] + jm];
      const double w5 = (v5old + sigma * (0.5 * wtau)) / sigma - 0.5, tt = (v6old + sigma * (0.5 * wtau)) / sigma + 0.5;
      if (lane == 0) part += w5 * w5;
      const double nrm = sqrt(spk_warp_sum(part));
      int mode;
      double tp;
      spk_soc_coeffs(tt, nrm, mode, tp);
      _Pragma("unroll") for (int k = lane; k < nx; k += WL) {
        const double w = dv3[c * nx + k];
        SPK_FIN(w, A.v[D.vo[2] + jm * nx + k], spk_cost_row(sQ, D.cost_dense, nx, k, dx), spk_soc_apply(mode, tp, nrm, w),
                D.vo[2] + jm * nx + k, dv3[c * nx + k]);
      }
      _Pragma("unroll") for (int k = lane; k < nu; k += WL) {
        const double w = dv4[c * nu + k];
        SPK_FIN(w, A.v[D.vo[3] + jm * nu + k], spk_cost_row(sR, D.cost_dense, nu, k, du), spk_soc_apply(mode, tp, nrm, w),
                D.vo[3] + jm * nu + k, dv4[c * nu + k]);
      }
      {  // v5, v6 and the tau entry of xi1 (L_transpose! :432-437); computed by every lane, counted by lane 0
        const double v5new = sigma * (w5 - spk_soc_apply(mode, tp, nrm, w5));
        const double v6new = sigma * (tt - (mode == 0 ? tt : (mode == 1 ? 0.0 : tp)));
        const double d5 = v5new - v5old, d6 = v6new - v6old, ld = 0.5 * dtau;
        if (lane == 0) {
          A.vbar[D.vo[4] + jm] = v5new;
          A.vbar[D.vo[5] + jm] = v6new;
          my_xi2 = spk_max_nn(my_xi2, spk_max_nn(fabs(ld - d5 * inv_sigma), fabs(ld - d6 * inv_sigma)));
          my_pn += d5 * (d5 - sigma * ld) + d6 * (d6 - sigma * ld);
          const double lt = 0.5 * d5 + 0.5 * d6;
          xi1_tau = spk_max_nn(xi1_tau, fabs(lt - dtau * inv_gamma));
          my_pn += dtau * (dtau - gamma * lt);
        }
      }
      if (leaves) {
        const long l = lfirst + c;
        const double *sQN = spk_sqrtQN(D, P, l);
        const double *wxl = wxc + (long)c * nx, *dxl = dxc + (long)c * nx;
        const double sb = A.zbar[D.zo[2] + j0], so = A.z[D.zo[2] + j0];
        const double wsl = 2.0 * sb - so, dsl = sb - so;
        double pl = 0.0;
        _Pragma("unroll") for (int k = lane; k < nx; k += WL) {  // L! :288-294
          const double w = (A.v[D.vo[7] + l * nx + k] + sigma * spk_cost_row(sQN, D.cost_dense, nx, k, wxl)) / sigma;
          dv11[c * nx + k] = w;
          pl += w * w;
        }
        const double v12old = A.v[D.vo[8] + l], v13old = A.v[D.vo[9] + l];
        const double w12 = (v12old + sigma * (0.5 * wsl)) / sigma - 0.5, tl = (v13old + sigma * (0.5 * wsl)) / sigma + 0.5;
        if (lane == 0) pl += w12 * w12;
        const double nrl = sqrt(spk_warp_sum(pl));
        int ml;
        double tpl;
        spk_soc_coeffs(tl, nrl, ml, tpl);
        _Pragma("unroll") for (int k = lane; k < nx; k += WL) {
          const double w = dv11[c * nx + k];
          SPK_FIN(w, A.v[D.vo[7] + l * nx + k], spk_cost_row(sQN, D.cost_dense, nx, k, dxl), spk_soc_apply(ml, tpl, nrl, w),
                  D.vo[7] + l * nx + k, dv11[c * nx + k]);
          const double v14old = A.v[D.vo[10] + l * nx + k];  // v14 box :804-809, L! :311-317
          const double wb = (v14old + sigma * wxl[k]) / sigma;
          SPK_FIN(wb, v14old, dxl[k], spk_clip(wb, D.x_min, D.x_max), D.vo[10] + l * nx + k, dv14[c * nx + k]);
        }
        const double v12new = sigma * (w12 - spk_soc_apply(ml, tpl, nrl, w12));
        const double v13new = sigma * (tl - (ml == 0 ? tl : (ml == 1 ? 0.0 : tpl)));
        const double d12 = v12new - v12old, d13 = v13new - v13old, lds = 0.5 * dsl;
        if (lane == 0) {
          A.vbar[D.vo[8] + l] = v12new;
          A.vbar[D.vo[9] + l] = v13new;
          my_xi2 = spk_max_nn(my_xi2, spk_max_nn(fabs(lds - d12 * inv_sigma), fabs(lds - d13 * inv_sigma)));
          my_pn += d12 * (d12 - sigma * lds) + d13 * (d13 - sigma * lds);
          const double lt = 0.5 * d12 + 0.5 * d13;  // s_leaf :445-448
          xi1_tau = spk_max_nn(xi1_tau, fabs(lt - dsl * inv_gamma));
          my_pn += dsl * (dsl - gamma * lt);
        }
      }
    }
#undef SPK_FIN
    my_xi1 = spk_max_nn(my_xi1, xi1_tau);
    SPK_SYNCWARP();
    // ---- xi1 = || L'dv - dz/gamma ||_inf and the z-part of <r, P r> for the z entries owned by this family
#define SPK_Z(dz_, lt_)                                        \
  {                                                            \
    const double dzv_ = (dz_), ltv_ = (lt_);                   \
    my_xi1 = spk_max_nn(my_xi1, fabs(ltv_ - dzv_ * inv_gamma));      \
    my_pn += dzv_ * (dzv_ - gamma * ltv_);                     \
  }
    _Pragma("unroll") for (int k = lane; k < nx; k += WL) {  // x_i :351-373
      double lt = dv7[k];
      for (int c = 0; c < d; ++c) lt += spk_cost_row(spk_sqrtQ(D, P, cfirst + c), D.cost_dense, nx, k, dv3 + c * nx);
      SPK_Z(dx[k], lt);
    }
    _Pragma("unroll") for (int k = lane; k < nu; k += WL) {  // u_i :399-420
      double lt = dv7[nx + k];
      for (int c = 0; c < d; ++c) lt += spk_cost_row(spk_sqrtR(D, P, cfirst + c), D.cost_dense, nu, k, dv4 + c * nu);
      SPK_Z(du[k], lt);
    }
    for (int k = lane; k < ny; k += WL) SPK_Z(dy[k], dv1[k] - spk_b(D, P, i0)[k] * dv2);  // y_i :423-429
    if (lane == 0) SPK_Z(ds_i, dv2);                                          // s_i :440-442
    if (leaves)
      for (int idx = lane; idx < d * nx; idx += WL) {  // x_leaf :376-396
        const int c = idx / nx, k = idx - c * nx;
        const double lt = dv14[idx] + spk_cost_row(spk_sqrtQN(D, P, lfirst + c), D.cost_dense, nx, k, dv11 + c * nx);
        SPK_Z(dxc[idx], lt);
      }
#undef SPK_Z
    if (A.pnorm) {  // deterministic: fixed butterfly over the lanes
      const double ws = spk_warp_sum(my_pn);
      if (lane == 0) A.pnorm[i0] = ws;
    }
    SPK_SYNCWARP();  // the warp's scratch is reused by its next family
  }
  const double m1 = spk_warp_max(my_xi1), m2 = spk_warp_max(my_xi2);
  if (lane == 0) {
    spk_atomic_max_nonneg(A.xi + 0, m1);
    spk_atomic_max_nonneg(A.xi + 1, m2);
  }
}

// ------------------------------------------------------------------------------------------------ 4b. batched dual kernel
//
// Same arithmetic as spk_k_dual, organised for memory-level parallelism: a CTA takes F consecutive families (all their
// segments are contiguous in the reference layout), stages every input with one burst of coalesced loads while forming
// w = 2 zbar - z and delta = zbar - z, and then runs element-parallel phases out of shared memory:
//   P2  W = (v + sigma L w)/sigma (+-1/2)      P3  SOC norms, one warp per block
//   P4  projection, vbar store, delta v, xi2   P5  xi1 and the z-part of <r, P r>
// A launch covers families whose children are all non-leaf or all leaf (A.children_are_leaves).
struct SpkDualMap {
  long wx, dx, wu, du, wy, dy, ws, ds, wsc, dsc, wtau, dtau, wxc, dxc;          // primal: w and delta
  long v1, v2, v7, v3, v4, v5, v6, v11, v12, v13, v14;                          // old dual values
  long W1, W2, W7, W3, W4, W5, W6, W11, W12, W13, W14;                          // W, later delta v
  long nrm, nrm1, red, total;
};
SPK_HD SpkDualMap spk_dual_map(int nx, int nu, int d, int ny, int F, bool leaves) {
  SpkDualMap m;
  long o = 0;
  auto take = [&](long cnt) { const long r = o; o += cnt; return r; };
  const long fx = (long)F * nx, fu = (long)F * nu, fy = (long)F * ny, fd = (long)F * d, fdx = fd * nx, fdu = fd * nu;
  m.wx = take(fx); m.dx = take(fx); m.wu = take(fu); m.du = take(fu); m.wy = take(fy); m.dy = take(fy);
  m.ws = take(F); m.ds = take(F); m.wsc = take(fd); m.dsc = take(fd); m.wtau = take(fd); m.dtau = take(fd);
  m.wxc = take(leaves ? fdx : 0); m.dxc = take(leaves ? fdx : 0);
  m.v1 = take(fy); m.v2 = take(F); m.v7 = take(fx + fu); m.v3 = take(fdx); m.v4 = take(fdu); m.v5 = take(fd); m.v6 = take(fd);
  m.v11 = take(leaves ? fdx : 0); m.v12 = take(leaves ? fd : 0); m.v13 = take(leaves ? fd : 0); m.v14 = take(leaves ? fdx : 0);
  m.W1 = take(fy); m.W2 = take(F); m.W7 = take(fx + fu); m.W3 = take(fdx); m.W4 = take(fdu); m.W5 = take(fd); m.W6 = take(fd);
  m.W11 = take(leaves ? fdx : 0); m.W12 = take(leaves ? fd : 0); m.W13 = take(leaves ? fd : 0); m.W14 = take(leaves ? fdx : 0);
  m.nrm = take(2 * fd); m.nrm1 = take(2L * F * SPK_MAX_CONES); m.red = take(3 * 32);
  m.total = o + 2;
  return m;
}

template <int NX_, int NU_, int D_>
#if !SPK_EMU
__launch_bounds__(256, 4)
#endif
SPK_KERNEL void spk_k_dual2(SpkDims D, SpkPtrs P, SpkFusedArgs A, int F, long fam_lo, long fam_cnt) {
  SPK_SMEM_DECL(double, S);
  if (A.gate && *A.gate) return;
  const int nx = NX_ > 0 ? NX_ : D.nx, nu = NX_ > 0 ? NU_ : D.nu, d = NX_ > 0 ? D_ : D.d, ny = D.ny;
  const int T = blockDim.x, tid = threadIdx.x;
  const int WL = T < 32 ? T : 32, nwarps = (T + WL - 1) / WL, warp = tid / WL, lane = tid % WL;
  const double gamma = A.gamma, sigma = A.sigma, inv_sigma = 1.0 / sigma, inv_gamma = 1.0 / gamma;
  const bool leaves = A.children_are_leaves != 0;
  const SpkDualMap M = spk_dual_map(nx, nu, d, ny, F, leaves);
  const int nxu = nx + nu;
  double my_xi1 = 0.0, my_xi2 = 0.0, my_pn = 0.0;
  const long nbatch = (fam_cnt + F - 1) / F;
  for (long b = blockIdx.x; b < nbatch; b += gridDim.x) {
    const long i0 = fam_lo + b * F;                                  // first family of the batch
    const int f = (int)((fam_cnt - b * F) < F ? (fam_cnt - b * F) : F);  // families in this batch
    const long cfirst = (long)d * i0 + 1, lfirst = cfirst - D.nnl;
    const int fd = f * d;
    // ---- P1: stage every input with asynchronous copies (all in flight at once), then turn the primal pairs
    //      (z, zbar) into w = 2 zbar - z and delta = zbar - z in place
#define SPK_LOADWD(cnt_, zoff_, wslot_, dslot_)                     \
  for (int e = tid; e < (cnt_); e += T) {                           \
    spk_cp_async8(S + (wslot_) + e, A.z + (zoff_) + e);             \
    spk_cp_async8(S + (dslot_) + e, A.zbar + (zoff_) + e);          \
  }
#define SPK_LOADV(cnt_, voff_, slot_) \
  for (int e = tid; e < (cnt_); e += T) spk_cp_async8(S + (slot_) + e, A.v + (voff_) + e);
    SPK_LOADWD(f * nx, i0 * nx, M.wx, M.dx)
    SPK_LOADWD(f * nu, D.zo[1] + i0 * nu, M.wu, M.du)
    SPK_LOADWD(f * ny, D.zo[4] + i0 * ny, M.wy, M.dy)
    SPK_LOADWD(f, D.zo[2] + i0, M.ws, M.ds)
    SPK_LOADWD(fd, D.zo[2] + cfirst, M.wsc, M.dsc)
    SPK_LOADWD(fd, D.zo[3] + cfirst - 1, M.wtau, M.dtau)
    SPK_LOADV(f * ny, i0 * ny, M.v1)
    SPK_LOADV(f, D.vo[1] + i0, M.v2)
    SPK_LOADV(f * nxu, D.vo[6] + i0 * nxu, M.v7)
    SPK_LOADV(fd * nx, D.vo[2] + (cfirst - 1) * nx, M.v3)
    SPK_LOADV(fd * nu, D.vo[3] + (cfirst - 1) * nu, M.v4)
    SPK_LOADV(fd, D.vo[4] + cfirst - 1, M.v5)
    SPK_LOADV(fd, D.vo[5] + cfirst - 1, M.v6)
    if (leaves) {
      SPK_LOADWD(fd * nx, cfirst * nx, M.wxc, M.dxc)
      SPK_LOADV(fd * nx, D.vo[7] + lfirst * nx, M.v11)
      SPK_LOADV(fd, D.vo[8] + lfirst, M.v12)
      SPK_LOADV(fd, D.vo[9] + lfirst, M.v13)
      SPK_LOADV(fd * nx, D.vo[10] + lfirst * nx, M.v14)
    }
#undef SPK_LOADWD
#undef SPK_LOADV
    spk_cp_async_wait_all();
    SPK_SYNCTHREADS();
    // the primal slots are laid out as consecutive (w, delta) pairs of equal length: wx,dx | wu,du | wy,dy | ws,ds | ...
#define SPK_WD(cnt_, wslot_, dslot_)                                \
  for (int e = tid; e < (cnt_); e += T) {                           \
    const double old_ = S[(wslot_) + e], nb_ = S[(dslot_) + e];     \
    S[(wslot_) + e] = 2.0 * nb_ - old_;                             \
    S[(dslot_) + e] = nb_ - old_;                                   \
  }
    SPK_WD(f * nx, M.wx, M.dx)
    SPK_WD(f * nu, M.wu, M.du)
    SPK_WD(f * ny, M.wy, M.dy)
    SPK_WD(f, M.ws, M.ds)
    SPK_WD(fd, M.wsc, M.dsc)
    SPK_WD(fd, M.wtau, M.dtau)
    if (leaves) SPK_WD(fd * nx, M.wxc, M.dxc)
#undef SPK_WD
    SPK_SYNCTHREADS();
    // ---- P2: W = (v + sigma L(2zbar - z)) / sigma with the +-1/2 shifts (prox_h_conj! :923-939, L! :177-318)
    for (int e = tid; e < f * ny; e += T) S[M.W1 + e] = (S[M.v1 + e] + sigma * S[M.wy + e]) / sigma;
    for (int ff = tid; ff < f; ff += T) {
      double bwy = 0.0;
      for (int k = 0; k < ny; ++k) bwy += spk_b(D, P, i0 + ff)[k] * S[M.wy + ff * ny + k];
      S[M.W2 + ff] = (S[M.v2 + ff] + sigma * (S[M.ws + ff] - bwy)) / sigma;
    }
    for (int e = tid; e < f * nxu; e += T) {
      const int ff = e / nxu, k = e - ff * nxu;
      S[M.W7 + e] = (S[M.v7 + e] + sigma * (k < nx ? S[M.wx + ff * nx + k] : S[M.wu + ff * nu + k - nx])) / sigma;
    }
    for (int e = tid; e < fd * nx; e += T) {
      const int fc = e / nx, k = e - fc * nx, ff = fc / d;
      S[M.W3 + e] = (S[M.v3 + e] + sigma * spk_cost_row(spk_sqrtQ(D, P, cfirst + fc), D.cost_dense, nx, k, S + M.wx + ff * nx)) / sigma;
      if (leaves) {
        const double *wxl = S + M.wxc + fc * nx;
        S[M.W11 + e] = (S[M.v11 + e] + sigma * spk_cost_row(spk_sqrtQN(D, P, lfirst + fc), D.cost_dense, nx, k, wxl)) / sigma;
        S[M.W14 + e] = (S[M.v14 + e] + sigma * wxl[k]) / sigma;
      }
    }
    for (int e = tid; e < fd * nu; e += T) {
      const int fc = e / nu, k = e - fc * nu, ff = fc / d;
      S[M.W4 + e] = (S[M.v4 + e] + sigma * spk_cost_row(spk_sqrtR(D, P, cfirst + fc), D.cost_dense, nu, k, S + M.wu + ff * nu)) / sigma;
    }
    for (int fc = tid; fc < fd; fc += T) {
      S[M.W5 + fc] = (S[M.v5 + fc] + sigma * (0.5 * S[M.wtau + fc])) / sigma - 0.5;
      S[M.W6 + fc] = (S[M.v6 + fc] + sigma * (0.5 * S[M.wtau + fc])) / sigma + 0.5;
      if (leaves) {
        S[M.W12 + fc] = (S[M.v12 + fc] + sigma * (0.5 * S[M.wsc + fc])) / sigma - 0.5;
        S[M.W13 + fc] = (S[M.v13 + fc] + sigma * (0.5 * S[M.wsc + fc])) / sigma + 0.5;
      }
    }
    SPK_SYNCTHREADS();
    // ---- P3: SOC norms, one warp per block: [0, fd) children (v3, v4, v5), [fd, 2fd) leaf children (v11, v12)
    for (int bb = warp; bb < (leaves ? 2 * fd : fd); bb += nwarps) {
      double part = 0.0;
      if (bb < fd) {
        for (int k = lane; k < nx; k += WL) { const double w = S[M.W3 + bb * nx + k]; part += w * w; }
        for (int k = lane; k < nu; k += WL) { const double w = S[M.W4 + bb * nu + k]; part += w * w; }
        if (lane == 0) part += S[M.W5 + bb] * S[M.W5 + bb];
      } else {
        const int fc = bb - fd;
        for (int k = lane; k < nx; k += WL) { const double w = S[M.W11 + fc * nx + k]; part += w * w; }
        if (lane == 0) part += S[M.W12 + fc] * S[M.W12 + fc];
      }
      part = spk_warp_sum(part);
      if (lane == 0) S[M.nrm + bb] = sqrt(part);
    }
    for (int idx = tid; idx < f * D.n_cones; idx += T) {  // SecondOrderCone sub-cones of the risk cone: (t, ||x||) per family
      const int ff = idx / D.n_cones, c = idx - ff * D.n_cones;
      if (D.cone_type[c] == 4) {
        int off = 0;
        for (int c2 = 0; c2 < c; ++c2) off += D.cone_dim[c2];
        double nn = 0.0;
        for (int k = 1; k < D.cone_dim[c]; ++k) nn += S[M.W1 + ff * ny + off + k] * S[M.W1 + ff * ny + off + k];
        S[M.nrm1 + 2 * idx] = S[M.W1 + ff * ny + off];
        S[M.nrm1 + 2 * idx + 1] = sqrt(nn);
      }
    }
    SPK_SYNCTHREADS();
    // ---- P4: projections (:752-909), vbar = sigma (W - Proj W), delta v (into the W slots), xi2, v-part of <r, P r>
#define SPK_FIN(Wslot, vold_, Ld_, pr_, gidx_)                          \
  {                                                                     \
    const double vnew_ = sigma * (S[Wslot] - (pr_));                    \
    A.vbar[gidx_] = vnew_;                                              \
    const double dl_ = vnew_ - (vold_);                                 \
    S[Wslot] = dl_;                                                     \
    my_xi2 = spk_max_nn(my_xi2, fabs((Ld_) - dl_ * inv_sigma));               \
    my_pn += dl_ * (dl_ - sigma * (Ld_));                               \
  }
    for (int e = tid; e < f * ny; e += T) {  // v1: projection on the dual risk cone (:833-845), one thread per entry
      const int ff = e / ny, k = e - ff * ny;
      int off = 0, type = 0, c = 0;
      for (; c < D.n_cones; ++c) {
        if (k < off + D.cone_dim[c]) { type = D.cone_type[c]; break; }
        off += D.cone_dim[c];
      }
      const double w = S[M.W1 + e];
      double pr;
      if (type == 0) pr = w;                          // Zeros -> dual Reals
      else if (type == 1) pr = w < 0.0 ? 0.0 : w;      // Nonnegatives
      else if (type == 2) pr = w > 0.0 ? 0.0 : w;      // Nonpositives
      else if (type == 3) pr = 0.0;                   // Reals -> dual Zeros
      else {                                          // SecondOrderCone (t, x), self-dual
        int mode;
        double tp;
        const double tt = S[M.nrm1 + 2 * (ff * D.n_cones + c)], nn = S[M.nrm1 + 2 * (ff * D.n_cones + c) + 1];
        spk_soc_coeffs(tt, nn, mode, tp);
        pr = (k == off) ? (mode == 0 ? w : (mode == 1 ? 0.0 : tp)) : spk_soc_apply(mode, tp, nn, w);
      }
      SPK_FIN(M.W1 + e, S[M.v1 + e], S[M.dy + e], pr, (i0 + ff) * ny + k);
    }
    for (int ff = tid; ff < f; ff += T) {  // v2 (:849-851)
      double bdy = 0.0;
      for (int k = 0; k < ny; ++k) bdy += spk_b(D, P, i0 + ff)[k] * S[M.dy + ff * ny + k];
      const double w2 = S[M.W2 + ff];
      SPK_FIN(M.W2 + ff, S[M.v2 + ff], S[M.ds + ff] - bdy, (w2 >= 0 ? w2 : 0.0), D.vo[1] + i0 + ff);
    }
    for (int e = tid; e < f * nxu; e += T) {  // v7 box (:903-908)
      const int ff = e / nxu, k = e - ff * nxu;
      const double w = S[M.W7 + e];
      const double pr = k < nx ? spk_clip(w, D.x_min, D.x_max) : spk_clip(w, D.u_min, D.u_max);
      SPK_FIN(M.W7 + e, S[M.v7 + e], (k < nx ? S[M.dx + ff * nx + k] : S[M.du + ff * nu + k - nx]), pr,
              D.vo[6] + i0 * nxu + e);
    }
    for (int e = tid; e < fd * nx; e += T) {  // v3 (:854-900) and, for leaf children, v11, v14 (:768-809)
      const int fc = e / nx, k = e - fc * nx, ff = fc / d;
      int mode;
      double tp;
      const double nrm = S[M.nrm + fc];
      spk_soc_coeffs(S[M.W6 + fc], nrm, mode, tp);
      const double ld = spk_cost_row(spk_sqrtQ(D, P, cfirst + fc), D.cost_dense, nx, k, S + M.dx + ff * nx);
      SPK_FIN(M.W3 + e, S[M.v3 + e], ld, spk_soc_apply(mode, tp, nrm, S[M.W3 + e]), D.vo[2] + (cfirst - 1) * nx + e);
      if (leaves) {
        const double nrl = S[M.nrm + fd + fc];
        spk_soc_coeffs(S[M.W13 + fc], nrl, mode, tp);
        const double *dxl = S + M.dxc + fc * nx;
        const double ld11 = spk_cost_row(spk_sqrtQN(D, P, lfirst + fc), D.cost_dense, nx, k, dxl);
        SPK_FIN(M.W11 + e, S[M.v11 + e], ld11, spk_soc_apply(mode, tp, nrl, S[M.W11 + e]), D.vo[7] + lfirst * nx + e);
        SPK_FIN(M.W14 + e, S[M.v14 + e], dxl[k], spk_clip(S[M.W14 + e], D.x_min, D.x_max), D.vo[10] + lfirst * nx + e);
      }
    }
    for (int e = tid; e < fd * nu; e += T) {  // v4
      const int fc = e / nu, k = e - fc * nu, ff = fc / d;
      int mode;
      double tp;
      const double nrm = S[M.nrm + fc];
      spk_soc_coeffs(S[M.W6 + fc], nrm, mode, tp);
      const double ld = spk_cost_row(spk_sqrtR(D, P, cfirst + fc), D.cost_dense, nu, k, S + M.du + ff * nu);
      SPK_FIN(M.W4 + e, S[M.v4 + e], ld, spk_soc_apply(mode, tp, nrm, S[M.W4 + e]), D.vo[3] + (cfirst - 1) * nu + e);
    }
    SPK_SYNCTHREADS();  // every reader of W6 / W13 (the t components) is done before they are overwritten
    for (int fc = tid; fc < fd; fc += T) {  // v5, v6; v12, v13 of leaf children
      int mode;
      double tp;
      const double nrm = S[M.nrm + fc], tt = S[M.W6 + fc];
      spk_soc_coeffs(tt, nrm, mode, tp);
      const double ld = 0.5 * S[M.dtau + fc];
      SPK_FIN(M.W5 + fc, S[M.v5 + fc], ld, spk_soc_apply(mode, tp, nrm, S[M.W5 + fc]), D.vo[4] + cfirst - 1 + fc);
      SPK_FIN(M.W6 + fc, S[M.v6 + fc], ld, (mode == 0 ? tt : (mode == 1 ? 0.0 : tp)), D.vo[5] + cfirst - 1 + fc);
      if (leaves) {
        const double nrl = S[M.nrm + fd + fc], tl = S[M.W13 + fc];
        spk_soc_coeffs(tl, nrl, mode, tp);
        const double lds = 0.5 * S[M.dsc + fc];
        SPK_FIN(M.W12 + fc, S[M.v12 + fc], lds, spk_soc_apply(mode, tp, nrl, S[M.W12 + fc]), D.vo[8] + lfirst + fc);
        SPK_FIN(M.W13 + fc, S[M.v13 + fc], lds, (mode == 0 ? tl : (mode == 1 ? 0.0 : tp)), D.vo[9] + lfirst + fc);
      }
    }
#undef SPK_FIN
    SPK_SYNCTHREADS();
    // ---- P5: xi1 = || L'dv - dz/gamma ||_inf and the z-part of <r, P r> (L_transpose! :320-449 on delta v)
#define SPK_Z(dz_, lt_)                                        \
  {                                                            \
    const double dzv_ = (dz_), ltv_ = (lt_);                   \
    my_xi1 = spk_max_nn(my_xi1, fabs(ltv_ - dzv_ * inv_gamma));      \
    my_pn += dzv_ * (dzv_ - gamma * ltv_);                     \
  }
    for (int e = tid; e < f * nx; e += T) {  // x_i :351-373
      const int ff = e / nx, k = e - ff * nx;
      double lt = S[M.W7 + ff * nxu + k];
      for (int c = 0; c < d; ++c)
        lt += spk_cost_row(spk_sqrtQ(D, P, cfirst + ff * d + c), D.cost_dense, nx, k, S + M.W3 + (ff * d + c) * nx);
      SPK_Z(S[M.dx + e], lt);
    }
    for (int e = tid; e < f * nu; e += T) {  // u_i :399-420
      const int ff = e / nu, k = e - ff * nu;
      double lt = S[M.W7 + ff * nxu + nx + k];
      for (int c = 0; c < d; ++c)
        lt += spk_cost_row(spk_sqrtR(D, P, cfirst + ff * d + c), D.cost_dense, nu, k, S + M.W4 + (ff * d + c) * nu);
      SPK_Z(S[M.du + e], lt);
    }
    for (int e = tid; e < f * ny; e += T) {  // y_i :423-429
      const int ff = e / ny, k = e - ff * ny;
      SPK_Z(S[M.dy + e], S[M.W1 + e] - spk_b(D, P, i0 + ff)[k] * S[M.W2 + ff]);
    }
    for (int ff = tid; ff < f; ff += T) SPK_Z(S[M.ds + ff], S[M.W2 + ff]);  // s_i :440-442
    for (int fc = tid; fc < fd; fc += T) {
      SPK_Z(S[M.dtau + fc], 0.5 * S[M.W5 + fc] + 0.5 * S[M.W6 + fc]);                 // tau_c :432-437
      if (leaves) SPK_Z(S[M.dsc + fc], 0.5 * S[M.W12 + fc] + 0.5 * S[M.W13 + fc]);    // s_leaf :445-448
    }
    if (leaves)
      for (int e = tid; e < fd * nx; e += T) {  // x_leaf :376-396
        const int fc = e / nx, k = e - fc * nx;
        const double lt = S[M.W14 + e] + spk_cost_row(spk_sqrtQN(D, P, lfirst + fc), D.cost_dense, nx, k, S + M.W11 + fc * nx);
        SPK_Z(S[M.dxc + e], lt);
      }
#undef SPK_Z
    SPK_SYNCTHREADS();  // the staging buffers are reused by the next batch
  }
  // ---- CTA-level folds: maxima by atomicMax; <r, P r> as one deterministic partial per CTA
  const double m1 = spk_warp_max(my_xi1), m2 = spk_warp_max(my_xi2), ps = spk_warp_sum(my_pn);
  if (lane == 0) {
    S[M.red + warp] = m1;
    S[M.red + 32 + warp] = m2;
    S[M.red + 64 + warp] = ps;
  }
  SPK_SYNCTHREADS();
  if (tid == 0) {
    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int w2 = 0; w2 < nwarps; ++w2) {
      a1 = spk_max_nn(a1, S[M.red + w2]);
      a2 = spk_max_nn(a2, S[M.red + 32 + w2]);
      a3 += S[M.red + 64 + w2];
    }
    spk_atomic_max_nonneg(A.xi + 0, a1);
    spk_atomic_max_nonneg(A.xi + 1, a2);
    if (A.pnorm) A.pnorm[blockIdx.x] = a3;
  }
}

// ------------------------------------------------------------------------------------------------ 4c. register-resident dual kernel
//
// Same arithmetic as spk_k_dual, organised as a pure streaming kernel for the specialised shapes with Diagonal
// sqrt-costs (every reference workload).  A group of G lanes (G = 8, 16 or 32, the smallest power of two covering nx and
// nu) owns one family and lane k of the group owns COLUMN k of the family -- x_i[k], v7_i[k], v3_c[k] and, for leaf
// children, x_c[k], v11_c[k], v14_c[k] (likewise the u-columns with v4_c) -- because L and L' with diagonal costs only
// couple entries of the same column.  Every input of the family is loaded into registers up front (~40 independent
// 8-byte loads per lane in flight), the only cross-lane work is the SOC norms (butterflies inside the group), and there
// is no shared-memory staging, no barrier in the main body and no integer division.  One family per group and no
// grid-stride loop: address registers die as soon as the loads are issued, so more warps are resident.
// UC: the sqrt-cost diagonals are uniform over the nodes instead of per node.
#if !SPK_EMU
#ifndef SPK_DUAL3_MINB
#define SPK_DUAL3_MINB 3
#endif
SPK_CONSTEXPR_HD int spk_dual3_group(int nx, int nu, int d) {
  int m = nx > nu ? nx : nu;
  if (m > 32) return 32;
  // nx, nu in 9..10: THREE families per warp on lanes 0 .. 3m-1 (the other lanes retire at once) instead of two in
  // groups of 16 with 6 idle lanes each -- the kernel is bound by instruction issue at these sizes
  if (m > 8 && m <= 10 && d <= m) return m;
  int g = 8;
  while (g < m || g < d) g <<= 1;
  return g;
}
// the residual maxima only ever grow: skip the atomic when the value cannot raise the maximum (a stale read only
// causes a redundant atomic, never a missed one)
SPK_D void spk_atomic_max_lazy(double *addr, double val) {
  if (spk_bits_nn(val) > spk_bits_nn(__ldcg(addr))) spk_atomic_max_nonneg(addr, val);
}

// exact maximum of non-negative doubles over the warp with two integer REDUX operations (high words, then the low
// words of the lanes that hold the maximal high word)
SPK_D double spk_warp_max_nonneg(double v, unsigned mask = 0xffffffffu) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(mask, hi);
  const unsigned ml = __reduce_max_sync(mask, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// PN: also accumulate <r, P r> (SuperMann only; Chambolle-Pock needs just the residual maxima)
// FLAT (subtree partition only): the launch covers flat owned-family indices over the rank's owned ranges instead of one
// contiguous node range.  A template parameter, not a run-time branch: merely having the range-table lookup in the kernel
// cost the single-GPU instantiation 9 % (0.951 -> 1.038 ms per pass at config 3, same-box A/B).
template <int NX, int NU, int DD, bool LEAVES, bool UC, bool PN, bool FLAT = false>
__launch_bounds__(128, SPK_DUAL3_MINB)
__global__ void spk_k_dual3(SpkDims D, SpkPtrs P, SpkFusedArgs A, long fam_lo, long fam_cnt) {
  if (A.gate && *A.gate) return;
  // current residual maxima, read now so that the round trip overlaps the family's loads (a stale value only causes a
  // redundant atomic at the end, never a missed one)
  const double seen_xi1 = __ldcg(A.xi + 0), seen_xi2 = __ldcg(A.xi + 1);
  constexpr int G = spk_dual3_group(NX, NU, DD), FPW = 32 / G;  // lanes per family, families per warp
  constexpr bool POW2 = (G & (G - 1)) == 0;                     // (else: groups of G consecutive lanes, 32 - FPW G lanes unused)
  constexpr unsigned MASK = POW2 ? 0xffffffffu : ((1u << (FPW * G)) - 1u);
  constexpr int GP2 = G > 16 ? 16 : (G > 8 ? 8 : (G > 4 ? 4 : 2));  // largest power of two below a non-power-of-two G
  constexpr int RX = (NX + G - 1) / G, RU = (NU + G - 1) / G, CD = UC ? 1 : DD, LD = LEAVES ? DD : 1;
  __shared__ double s_red[3 * 4];
  const int wlane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  if (!POW2 && wlane >= FPW * G) return;  // (exited lanes count as arrived at the barrier of the <r,Pr> fold)
  const int lane = POW2 ? (wlane & (G - 1)) : (wlane % G);  // lane within the family's group
  const int gbase = wlane - lane;
  // sum over the group on every lane of it: butterflies, or (G not a power of two) a fold towards lane 0 and a broadcast
#define SPK_GSUM_STEP(x_, o_)                                                \
  {                                                                          \
    if constexpr (POW2) (x_) += __shfl_xor_sync(0xffffffffu, (x_), (o_));    \
    else {                                                                   \
      const double t_ = __shfl_down_sync(MASK, (x_), (o_));                  \
      if (lane + (o_) < G) (x_) += t_;                                       \
    }                                                                        \
  }
#define SPK_GSUM_DONE(x_) \
  { if constexpr (!POW2) (x_) = __shfl_sync(MASK, (x_), gbase); }
#define SPK_GBCAST(x_, src_) (POW2 ? __shfl_sync(0xffffffffu, (x_), (src_), G) : __shfl_sync(MASK, (x_), gbase + (src_)))
  const int ny = D.ny;
  const double gamma = A.gamma, sigma = A.sigma, inv_sigma = A.inv_sigma, inv_gamma = A.inv_gamma;
  const double *__restrict__ z = A.z, *__restrict__ zb = A.zbar, *__restrict__ v = A.v;
  double *__restrict__ vb = A.vbar;
  const double x_min = D.x_min, x_max = D.x_max, u_min = D.u_min, u_max = D.u_max;
  // trailing groups of the last warp recompute the last family (identical stores; only their <r,Pr> share is dropped),
  // so that every shuffle below is executed by full warps
  long f = ((long)blockIdx.x * wpb + warp) * FPW + wlane / G;
  const bool valid = f < fam_cnt;
  if (!valid) f = fam_cnt - 1;
  // FLAT: fam_lo = first flat owned-family index of the launch
  const long i0 = FLAT ? spk_owned_family(A, fam_lo + f) : fam_lo + f;
  const long cfirst = (long)DD * i0 + 1, lfirst = cfirst - D.nnl;
  double my_xi1 = 0.0, my_xi2 = 0.0, my_pn = 0.0;
#define SPK_XON(r) ((NX % G == 0) || (r) < RX - 1 || lane + G * (r) < NX)
#define SPK_UON(r) ((NU % G == 0) || (r) < RU - 1 || lane + G * (r) < NU)
#define SPK_Z3(dz_, lt_)                                       \
  {                                                            \
    const double dzv_ = (dz_), ltv_ = (lt_);                   \
    my_xi1 = spk_max_nn(my_xi1, fabs(ltv_ - dzv_ * inv_gamma));      \
    if (PN) my_pn += dzv_ * (dzv_ - gamma * ltv_);             \
  }
  // vbar = sigma (W - Proj W), delta v, xi2 and the v-part of <r, P r> of one dual entry
#define SPK_FIN3(ptr_, w_, vold_, Ld_, pr_, dl_)                        \
  {                                                                     \
    const double vnew_ = sigma * ((w_) - (pr_));                        \
    *(ptr_) = vnew_;                                                    \
    dl_ = vnew_ - (vold_);                                              \
    my_xi2 = spk_max_nn(my_xi2, fabs((Ld_) - dl_ * inv_sigma));               \
    if (PN) my_pn += dl_ * (dl_ - sigma * (Ld_));                       \
  }
  // ---------------------------------------------------------------- loads (all independent)
  double wx[RX], dx[RX], v7x[RX], v3[DD][RX], sQ[CD][RX], sQN[CD][RX];
  double wu[RU], du[RU], v7u[RU], v4[DD][RU], sR[CD][RU];
  double wxc[LD][RX], dxc[LD][RX], v11[LD][RX], v14[LD][RX];
  {
    const double *zx_p = z + i0 * NX, *bx_p = zb + i0 * NX, *v7_p = v + D.vo[6] + i0 * (NX + NU);
    const double *v3_p = v + D.vo[2] + (cfirst - 1) * NX;
#pragma unroll
    for (int r = 0; r < RX; ++r) {
      const int k = lane + G * r;
      const bool on = SPK_XON(r);
      wx[r] = on ? zx_p[k] : 0.0;
      dx[r] = on ? bx_p[k] : 0.0;
      v7x[r] = on ? v7_p[k] : 0.0;
#pragma unroll
      for (int c = 0; c < DD; ++c) v3[c][r] = on ? v3_p[c * NX + k] : 0.0;
#pragma unroll
      for (int c = 0; c < CD; ++c) sQ[c][r] = on ? P.sqrtQ[(UC ? 0 : (cfirst - 1 + c) * NX) + k] : 0.0;
      if (LEAVES) {
        const double *zc_p = z + cfirst * NX, *bc_p = zb + cfirst * NX;
        const double *v11_p = v + D.vo[7] + lfirst * NX, *v14_p = v + D.vo[10] + lfirst * NX;
#pragma unroll
        for (int c = 0; c < LD; ++c) {
          wxc[c][r] = on ? zc_p[c * NX + k] : 0.0;
          dxc[c][r] = on ? bc_p[c * NX + k] : 0.0;
          v11[c][r] = on ? v11_p[c * NX + k] : 0.0;
          v14[c][r] = on ? v14_p[c * NX + k] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < CD; ++c) sQN[c][r] = on ? P.sqrtQN[(UC ? 0 : (lfirst + c) * NX) + k] : 0.0;
      }
    }
    const double *zu_p = z + D.zo[1] + i0 * NU, *bu_p = zb + D.zo[1] + i0 * NU;
    const double *v4_p = v + D.vo[3] + (cfirst - 1) * NU;
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      const int k = lane + G * r;
      const bool on = SPK_UON(r);
      wu[r] = on ? zu_p[k] : 0.0;
      du[r] = on ? bu_p[k] : 0.0;
      v7u[r] = on ? v7_p[NX + k] : 0.0;
#pragma unroll
      for (int c = 0; c < DD; ++c) v4[c][r] = on ? v4_p[c * NU + k] : 0.0;
#pragma unroll
      for (int c = 0; c < CD; ++c) sR[c][r] = on ? P.sqrtR[(UC ? 0 : (cfirst - 1 + c) * NU) + k] : 0.0;
    }
  }
  // y_i, v1_i on lanes < ny; the per-child scalars (tau_c, v5_c, v6_c; leaf: s_c, v12_c, v13_c) on lane c
  double wy = 0.0, dy = 0.0, v1o = 0.0;
  if (lane < ny) {
    wy = z[D.zo[4] + i0 * ny + lane];
    dy = zb[D.zo[4] + i0 * ny + lane];
    v1o = v[i0 * ny + lane];
  }
  double wtau = 0.0, dtau = 0.0, v5o = 0.0, v6o = 0.0, wsl = 0.0, dsl = 0.0, v12o = 0.0, v13o = 0.0;
  if (lane < DD) {
    wtau = z[D.zo[3] + cfirst - 1 + lane];
    dtau = zb[D.zo[3] + cfirst - 1 + lane];
    v5o = v[D.vo[4] + cfirst - 1 + lane];
    v6o = v[D.vo[5] + cfirst - 1 + lane];
    if (LEAVES) {
      wsl = z[D.zo[2] + cfirst + lane];
      dsl = zb[D.zo[2] + cfirst + lane];
      v12o = v[D.vo[8] + lfirst + lane];
      v13o = v[D.vo[9] + lfirst + lane];
    }
  }
  double ws_i = z[D.zo[2] + i0], ds_i = zb[D.zo[2] + i0];
  const double v2o = v[D.vo[1] + i0];
  // lane -> risk sub-cone of its y entry (lanes >= ny are idle in the y phases)
  int y_type = 3, y_off = 0, y_cone = 0;
  bool has_soc = false;
  {
    int off = 0;
    for (int c = 0; c < D.n_cones; ++c) {
      if (lane >= off && lane < off + D.cone_dim[c]) { y_type = D.cone_type[c]; y_off = off; y_cone = c; }
      has_soc |= (D.cone_type[c] == 4);
      off += D.cone_dim[c];
    }
  }
  const double b_k = lane < ny ? spk_b(D, P, i0)[lane] : 0.0;
  // ---------------------------------------------------------------- w = 2 zbar - z, delta = zbar - z (in place)
#define SPK_WD3(w_, d_) { const double old_ = (w_), nb_ = (d_); (w_) = 2.0 * nb_ - old_; (d_) = nb_ - old_; }
#pragma unroll
  for (int r = 0; r < RX; ++r) {
    SPK_WD3(wx[r], dx[r]);
    if (LEAVES) {
#pragma unroll
      for (int c = 0; c < LD; ++c) SPK_WD3(wxc[c][r], dxc[c][r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RU; ++r) SPK_WD3(wu[r], du[r]);
  SPK_WD3(wy, dy);
  SPK_WD3(wtau, dtau);
  SPK_WD3(wsl, dsl);
  SPK_WD3(ws_i, ds_i);
#undef SPK_WD3
  // ---------------------------------------------------------------- SOC norms: children (v3, v4, v5), leaf children (v11, v12)
  // W = (v + sigma L(2zbar - z))/sigma with the +-1/2 shifts (prox_h_conj! :923-939, L! :226-260, :288-301)
  const double w5 = v5o * inv_sigma + 0.5 * wtau - 0.5, tt = v6o * inv_sigma + 0.5 * wtau + 0.5;
  const double w12 = v12o * inv_sigma + 0.5 * wsl - 0.5, tl = v13o * inv_sigma + 0.5 * wsl + 0.5;
  double part[DD], partl[LD];
#pragma unroll
  for (int c = 0; c < DD; ++c) {
    double acc = (lane == c) ? w5 * w5 : 0.0;
#pragma unroll
    for (int r = 0; r < RX; ++r) {
      const double w = v3[c][r] * inv_sigma + sQ[UC ? 0 : c][r] * wx[r];
      acc += w * w;
    }
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      const double w = v4[c][r] * inv_sigma + sR[UC ? 0 : c][r] * wu[r];
      acc += w * w;
    }
    part[c] = acc;
    if (LEAVES) {
      double al = (lane == c) ? w12 * w12 : 0.0;
#pragma unroll
      for (int r = 0; r < RX; ++r) {
        const double w = v11[c][r] * inv_sigma + sQN[UC ? 0 : c][r] * wxc[c][r];
        al += w * w;
      }
      partl[c] = al;
    }
  }
#pragma unroll
  for (int o = POW2 ? G / 2 : GP2; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      SPK_GSUM_STEP(part[c], o);
      if (LEAVES) SPK_GSUM_STEP(partl[c], o);
    }
  }
#pragma unroll
  for (int c = 0; c < DD; ++c) {
    SPK_GSUM_DONE(part[c]);
    if (LEAVES) SPK_GSUM_DONE(partl[c]);
  }
  // lane c turns (t_c, ||.||_c) into the SOC projection coefficients of child c (risk_measures.jl:57-70):
  // Proj(t, x) = (pt, a x) with a = 1 | 0 | t'/||x||, pt = t | 0 | t'
  double my_a = 1.0, my_pt = 0.0, my_al = 1.0, my_ptl = 0.0;
  {
    double my_n = 0.0, my_nl = 0.0;
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      if (lane == c) {
        my_n = part[c];
        if (LEAVES) my_nl = partl[c];
      }
    }
    if (lane < DD) {
      int mode;
      double tp;
      const double nrm = sqrt(my_n);
      spk_soc_coeffs(tt, nrm, mode, tp);
      my_a = mode == 0 ? 1.0 : (mode == 1 ? 0.0 : tp / nrm);
      my_pt = mode == 0 ? tt : (mode == 1 ? 0.0 : tp);
      if (LEAVES) {
        const double nrl = sqrt(my_nl);
        spk_soc_coeffs(tl, nrl, mode, tp);
        my_al = mode == 0 ? 1.0 : (mode == 1 ? 0.0 : tp / nrl);
        my_ptl = mode == 0 ? tl : (mode == 1 ? 0.0 : tp);
      }
    }
  }
  // ---------------------------------------------------------------- own blocks v1_i (:833-845), v2_i (:849-851)
  double dv1 = 0.0, dv2;
  {
    const double W1 = v1o * inv_sigma + wy;  // L! :207-210
    double pr;
    if (y_type == 0) pr = W1;                            // Zeros -> dual Reals
    else if (y_type == 1) pr = W1 < 0.0 ? 0.0 : W1;       // Nonnegatives
    else if (y_type == 2) pr = W1 > 0.0 ? 0.0 : W1;       // Nonpositives
    else pr = 0.0;                                       // Reals -> dual Zeros (SecondOrderCone below)
    if (has_soc) {
      for (int c = 0; c < D.n_cones; ++c) {
        if (D.cone_type[c] != 4) continue;
        int off = 0;
        for (int c2 = 0; c2 < c; ++c2) off += D.cone_dim[c2];
        double sq = (lane > off && lane < off + D.cone_dim[c]) ? W1 * W1 : 0.0;
#pragma unroll
        for (int o = POW2 ? G / 2 : GP2; o > 0; o >>= 1) SPK_GSUM_STEP(sq, o);
        SPK_GSUM_DONE(sq);
        const double snrm = sqrt(sq);
        const double st = SPK_GBCAST(W1, off);
        int mode;
        double tp;
        spk_soc_coeffs(st, snrm, mode, tp);
        if (y_type == 4 && y_cone == c)
          pr = (lane == y_off) ? (mode == 0 ? W1 : (mode == 1 ? 0.0 : tp)) : spk_soc_apply(mode, tp, snrm, W1);
      }
    }
    if (lane < ny) SPK_FIN3(vb + i0 * ny + lane, W1, v1o, dy, pr, dv1);
    double bwy = b_k * wy, bdy = b_k * dy;
#pragma unroll
    for (int o = POW2 ? G / 2 : GP2; o > 0; o >>= 1) {
      SPK_GSUM_STEP(bwy, o);
      SPK_GSUM_STEP(bdy, o);
    }
    SPK_GSUM_DONE(bwy);
    SPK_GSUM_DONE(bdy);
    const double w2 = v2o * inv_sigma + (ws_i - bwy);  // L! :214-222
    const double v2new = sigma * (w2 - (w2 >= 0 ? w2 : 0.0));
    dv2 = v2new - v2o;
    if (lane == 0) {
      vb[D.vo[1] + i0] = v2new;
      const double l2 = ds_i - bdy;
      my_xi2 = spk_max_nn(my_xi2, fabs(l2 - dv2 * inv_sigma));
      if (PN) my_pn += dv2 * (dv2 - sigma * l2);
    }
  }
  // ---------------------------------------------------------------- v7_i box (:903-908); starts the L'dv accumulators
  double ltx[RX], ltu[RU];
  {
    double *vb7 = vb + D.vo[6] + i0 * (NX + NU);
#pragma unroll
    for (int r = 0; r < RX; ++r) {
      ltx[r] = 0.0;
      if (SPK_XON(r)) {
        const double w = v7x[r] * inv_sigma + wx[r];  // L! :275-282
        SPK_FIN3(vb7 + lane + G * r, w, v7x[r], dx[r], spk_clip(w, x_min, x_max), ltx[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      ltu[r] = 0.0;
      if (SPK_UON(r)) {
        const double w = v7u[r] * inv_sigma + wu[r];
        SPK_FIN3(vb7 + NX + lane + G * r, w, v7u[r], du[r], spk_clip(w, u_min, u_max), ltu[r]);
      }
    }
  }
  // ---------------------------------------------------------------- children: SOC(nx+nu+2) on (t = v6; v3, v4, v5) (:854-900)
#pragma unroll
  for (int c = 0; c < DD; ++c) {
    const long jm = cfirst - 1 + c;
    const double a = SPK_GBCAST(my_a, c);
    double *vb3 = vb + D.vo[2] + jm * NX, *vb4 = vb + D.vo[3] + jm * NU;
#pragma unroll
    for (int r = 0; r < RX; ++r) {
      if (SPK_XON(r)) {
        const double q = sQ[UC ? 0 : c][r], w = v3[c][r] * inv_sigma + q * wx[r];
        double dl;
        SPK_FIN3(vb3 + lane + G * r, w, v3[c][r], q * dx[r], a * w, dl);
        ltx[r] += q * dl;
      }
    }
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      if (SPK_UON(r)) {
        const double q = sR[UC ? 0 : c][r], w = v4[c][r] * inv_sigma + q * wu[r];
        double dl;
        SPK_FIN3(vb4 + lane + G * r, w, v4[c][r], q * du[r], a * w, dl);
        ltu[r] += q * dl;
      }
    }
    if (LEAVES) {  // leaf children: SOC(nx+2) on (v13; v11, v12) (:768-800), v14 box (:804-809)
      const long l = lfirst + c;
      const double al = SPK_GBCAST(my_al, c);
      double *vb11 = vb + D.vo[7] + l * NX, *vb14 = vb + D.vo[10] + l * NX;
#pragma unroll
      for (int r = 0; r < RX; ++r) {
        if (SPK_XON(r)) {
          const double q = sQN[UC ? 0 : c][r], w = v11[c][r] * inv_sigma + q * wxc[c][r];  // L! :288-294
          double dl, d14;
          SPK_FIN3(vb11 + lane + G * r, w, v11[c][r], q * dxc[c][r], al * w, dl);
          const double wb = v14[c][r] * inv_sigma + wxc[c][r];  // L! :311-317
          SPK_FIN3(vb14 + lane + G * r, wb, v14[c][r], dxc[c][r], spk_clip(wb, x_min, x_max), d14);
          SPK_Z3(dxc[c][r], d14 + q * dl);  // x_leaf :376-396
        }
      }
    }
  }
  if (lane < DD) {  // v5, v6 (t) of child `lane`, tau entry of xi1 (L_transpose! :432-437); leaf: v12, v13, s_leaf (:445-448)
    const long jm = cfirst - 1 + lane;
    const double ld = 0.5 * dtau;
    double d5, d6;
    SPK_FIN3(vb + D.vo[4] + jm, w5, v5o, ld, my_a * w5, d5);
    SPK_FIN3(vb + D.vo[5] + jm, tt, v6o, ld, my_pt, d6);
    SPK_Z3(dtau, 0.5 * d5 + 0.5 * d6);
    if (LEAVES) {
      const long l = lfirst + lane;
      const double lds = 0.5 * dsl;
      double d12, d13;
      SPK_FIN3(vb + D.vo[8] + l, w12, v12o, lds, my_al * w12, d12);
      SPK_FIN3(vb + D.vo[9] + l, tl, v13o, lds, my_ptl, d13);
      SPK_Z3(dsl, 0.5 * d12 + 0.5 * d13);
    }
  }
  // ---------------------------------------------------------------- xi1 and the z-part of <r, P r> for x_i, u_i, y_i, s_i
#pragma unroll
  for (int r = 0; r < RX; ++r)
    if (SPK_XON(r)) SPK_Z3(dx[r], ltx[r]);  // :351-373
#pragma unroll
  for (int r = 0; r < RU; ++r)
    if (SPK_UON(r)) SPK_Z3(du[r], ltu[r]);  // :399-420
  if (lane < ny) SPK_Z3(dy, dv1 - b_k * dv2);  // :423-429
  if (lane == 0) SPK_Z3(ds_i, dv2);            // :440-442
#undef SPK_FIN3
#undef SPK_Z3
#undef SPK_XON
#undef SPK_UON
#undef SPK_GSUM_STEP
#undef SPK_GSUM_DONE
#undef SPK_GBCAST
  // ---- folds: maxima by (lazy) atomicMax; <r, P r> as one deterministic partial per CTA
  if (!valid) my_pn = 0.0;
  const double m1 = spk_warp_max_nonneg(my_xi1, MASK), m2 = spk_warp_max_nonneg(my_xi2, MASK);
  if (wlane == 0) {
    if (spk_bits_nn(m1) > spk_bits_nn(seen_xi1)) spk_atomic_max_nonneg(A.xi + 0, m1);
    if (spk_bits_nn(m2) > spk_bits_nn(seen_xi2)) spk_atomic_max_nonneg(A.xi + 1, m2);
  }
  if (PN) {
    double ps;
    if constexpr (POW2) {
      ps = spk_warp_sum(my_pn);
    } else {  // fold towards lane 0 over the FPW G live lanes
      ps = my_pn;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_down_sync(MASK, ps, o);
        if (wlane + o < FPW * G) ps += t;
      }
    }
    if (wlane == 0) s_red[warp] = ps;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a3 = 0.0;
      for (int w2 = 0; w2 < wpb; ++w2) a3 += s_red[w2];
      A.pnorm[blockIdx.x] = a3;
    }
  }
}
#endif
