// spk_mm.cuh -- per-stage ("factor table") S1 sweeps on the fp64 tensor pipe (DMMA, mma.sync.m8n8k4.f64).
//
// With SPK_FACTORS_PER_STAGE all families of a level share their Riccati factors, so a level of the backward / forward
// sweep of projection_S1! (implicit_l.jl:584-637, :667-681) is a chain of small mat-mats over the families of the level
// (tools/microbench/dmma_bench.cu: DMMA reaches the same 37 TFLOP/s as DFMA on B200, but takes its weight operand from a
// single conflict-free 8-byte shared-memory read per lane and needs no per-thread index arithmetic).  One warp owns a
// tile of 8 consecutive families:
//   * the vectors of the tile (children's q, x_i, d_i) are read from HBM DIRECTLY in A-fragment order (lane (m, kk)
//     reads entry 4s+kk of family m: 32 contiguous bytes per family and k-step), the right-hand sides / outputs in
//     C-fragment order (two consecutive entries per lane) -- there is no shared-memory staging of vectors;
//   * the level's weights sit in shared memory in B-fragment order (spk_mm_layout.h), fetched once per CTA with TMA bulk
//     copies from the (L2-resident) per-level table;
//   * intermediate results (rhs -> d_i -> G d_i) are fed back from accumulator registers as the next A operand through
//     the permuted inner ordering, without shuffles.
// Arithmetic per family is that of spk_k_bwd / spk_k_fwd_s1 (same folded factors), summed in DMMA order.
#pragma once
#include "spk_fused.cuh"
#include "spk_mm_layout.h"

#if !SPK_EMU
SPK_D void spk_dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
SPK_CONSTEXPR_HD int spk_mm_minb(int frags) { return frags * 256 > 110 * 1024 ? 1 : 2; }
SPK_CONSTEXPR_HD size_t spk_mm_smem_bytes(int frags) { return 16 + (size_t)frags * 256; }

// Two consecutive entries p[n0], p[n0 + 1] (zero beyond nlim).  vec2: every such pair of this launch is 16-byte aligned
// (host-checked: even sizes and segment offsets), so it is ONE 16-byte load -- half the L1 wavefronts of the 8-byte form.
SPK_D void spk_ld2(const double *__restrict__ p, int n0, int nlim, bool vec2, double &a, double &b) {
  if (vec2) {
    if (n0 < nlim) {
      const double2 t = *reinterpret_cast<const double2 *>(p + n0);
      a = t.x;
      b = t.y;
    } else {
      a = 0.0;
      b = 0.0;
    }
  } else {
    a = n0 < nlim ? p[n0] : 0.0;
    b = n0 + 1 < nlim ? p[n0 + 1] : 0.0;
  }
}
SPK_D void spk_st2(double *p, int n0, int nlim, bool vec2, double a, double b) {
  if (vec2) {
    if (n0 < nlim) *reinterpret_cast<double2 *>(p + n0) = make_double2(a, b);
  } else {
    if (n0 < nlim) p[n0] = a;
    if (n0 + 1 < nlim) p[n0 + 1] = b;
  }
}
// A operands of all k-steps of an `n`-vector at p (inner order spk_mm_kidx): steps in pairs, an odd last step alone
template <int STEPS>
SPK_D void spk_ld_afrag(const double *__restrict__ p, int n, int kk, bool vec2, double (&a)[STEPS]) {
#pragma unroll
  for (int S = 0; S < STEPS / 2; ++S) spk_ld2(p, 8 * S + 2 * kk, n, vec2, a[2 * S], a[2 * S + 1]);
  if (STEPS & 1) {
    const int k = 4 * (STEPS - 1) + kk;
    a[STEPS - 1] = k < n ? p[k] : 0.0;
  }
}

// thread 0: fetch `frags` fragments of the level's table into shared memory (completion on `bar`)
SPK_D void spk_mm_fetch(double *smem, const double *frag_g, int frags, int tid) {
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  if (tid == 0) {
    spk_mbar_init(bar, 1);
    const uint32_t total = (uint32_t)frags * 256u;
    spk_mbar_expect_tx(bar, total);
    for (uint32_t off = 0; off < total; off += 32768u) {
      const uint32_t sz = total - off < 32768u ? total - off : 32768u;
      spk_bulk_g2s(reinterpret_cast<char *>(smem + 2) + off, reinterpret_cast<const char *>(frag_g) + off, sz, bar);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ backward level
// One tile (8 consecutive families starting at position 8*tile of the level whose first node is first0) of the
// backward step.  fr: the level's fragments in shared memory, complete once `bar` has flipped past `phase`.
template <int NX, int NU, int DD, bool LEAVES>
__device__ __forceinline__ void spk_bwd_mm_tile(const SpkDims &D, const SpkPtrs &P, const SpkFusedArgs &A, const double *fr,
                                                uint64_t *bar, uint32_t phase, bool &waited, long first0, long count, long tile,
                                                int lane) {
  constexpr SpkMmLayout L = spk_mm_layout(NX, NU, DD);
  constexpr int KX = L.KX, NBX = L.NBX, NBU = L.NBU, NBS = L.NBU + L.NBX;
  const int m = lane >> 2, kk = lane & 3;
  const double gamma = A.gamma;
  const double *__restrict__ z = A.z, *__restrict__ v = A.v;
  const bool ucq = D.n_sqrtQ == 1, ucn = D.n_sqrtQN == 1;
  {
    long fm = tile * 8 + m;
    const bool valid = fm < count;
    if (!valid) fm = count - 1;  // trailing rows of the last tile recompute the last family; their stores are dropped
    const long i0 = first0 + fm, cfirst = (long)DD * i0 + 1, lfirst = cfirst - D.nnl;
    // ---- tu = (z - gamma L'v) on u_i (L_transpose! :399-420) in C-fragment order: entries 8 nb + 2 kk + {0, 1}
    const bool vec2 = A.vec2 != 0;
    double tu[NBU][2];
    {
      const double *v7u = v + D.vo[6] + i0 * (NX + NU) + NX, *zu = z + D.zo[1] + i0 * NU;
#pragma unroll
      for (int nb = 0; nb < NBU; ++nb) {
        const int n0 = 8 * nb + 2 * kk;
        double l0, l1, z0, z1;
        spk_ld2(v7u, n0, NU, vec2, l0, l1);
#pragma unroll
        for (int c = 0; c < DD; ++c) {
          double s0, s1, w0, w1;
          spk_ld2(P.sqrtR + (ucq ? 0 : (cfirst - 1 + c) * NU), n0, NU, vec2, s0, s1);
          spk_ld2(v + D.vo[3] + (cfirst - 1 + c) * NU, n0, NU, vec2, w0, w1);
          l0 += s0 * w0;
          l1 += s1 * w1;
        }
        spk_ld2(zu, n0, NU, vec2, z0, z1);
        tu[nb][0] = z0 - gamma * l0;
        tu[nb][1] = z1 - gamma * l1;
      }
    }
    // ---- sweep over the children's q (A fragments straight from HBM): acc1 = sum_c B_c' q_c (:590-596),
    //      acc3 = sum_c ABK_c' q_c (:610-632)
    double acc1[NBU][2], acc3[NBX][2];
#pragma unroll
    for (int nb = 0; nb < NBU; ++nb) acc1[nb][0] = acc1[nb][1] = 0.0;
    // acc3 starts at -(z - gamma L'v) on x_i (:634-636 with L_transpose! :351-373): these loads are issued with the tile's
    // first loads instead of sitting exposed behind the DMMA chain
    {
      const double *v7x = v + D.vo[6] + i0 * (NX + NU), *zx = z + i0 * NX;
#pragma unroll
      for (int nb = 0; nb < NBX; ++nb) {
        const int n0 = 8 * nb + 2 * kk;
        double l0, l1, z0, z1;
        spk_ld2(v7x, n0, NX, vec2, l0, l1);
#pragma unroll
        for (int c = 0; c < DD; ++c) {
          double s0, s1, w0, w1;
          spk_ld2(P.sqrtQ + (ucq ? 0 : (cfirst - 1 + c) * NX), n0, NX, vec2, s0, s1);
          spk_ld2(v + D.vo[2] + (cfirst - 1 + c) * NX, n0, NX, vec2, w0, w1);
          l0 += s0 * w0;
          l1 += s1 * w1;
        }
        spk_ld2(zx, n0, NX, vec2, z0, z1);
        acc3[nb][0] = gamma * l0 - z0;
        acc3[nb][1] = gamma * l1 - z1;
      }
    }
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      double qa[KX];
      if (LEAVES) {  // q_leaf = -(x - gamma L'v)_leaf, implicit_l.jl:578-582 with :376-396
        const long l = lfirst + c;
        const double *pz = z + (cfirst + c) * NX, *p14 = v + D.vo[10] + l * NX, *psq = P.sqrtQN + (ucn ? 0 : l * NX),
                     *p11 = v + D.vo[7] + l * NX;
#pragma unroll
        for (int S = 0; S < KX / 2; ++S) {
          const int n0 = 8 * S + 2 * kk;
          double z0, z1, a0, a1, q0, q1, b0, b1;
          spk_ld2(pz, n0, NX, vec2, z0, z1);
          spk_ld2(p14, n0, NX, vec2, a0, a1);
          spk_ld2(psq, n0, NX, vec2, q0, q1);
          spk_ld2(p11, n0, NX, vec2, b0, b1);
          qa[2 * S] = -(z0 - gamma * (a0 + q0 * b0));
          qa[2 * S + 1] = -(z1 - gamma * (a1 + q1 * b1));
        }
        if (KX & 1) {
          const int k = 4 * (KX - 1) + kk;
          qa[KX - 1] = k < NX ? -(pz[k] - gamma * (p14[k] + psq[k] * p11[k])) : 0.0;
        }
      } else {
        spk_ld_afrag<KX>(A.q + (cfirst + c) * NX, NX, kk, vec2, qa);
      }
      if (!waited) {
        spk_mbar_wait(bar, phase);
        waited = true;
      }
#pragma unroll
      for (int s = 0; s < KX; ++s) {
        const double *base = fr + (size_t)((c * KX + s) * NBS) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < NBU; ++nb) spk_dmma(acc1[nb][0], acc1[nb][1], qa[s], base[nb * 32]);
#pragma unroll
        for (int nb = 0; nb < NBX; ++nb) spk_dmma(acc3[nb][0], acc3[nb][1], qa[s], base[(NBU + nb) * 32]);
      }
    }
    // ---- d_i = Rtilde^{-1} (tu - sum_c B_c' q_c) (:598-601): the accumulators are the next A operand (permuted k)
    double dv[NBU][2];
#pragma unroll
    for (int nb = 0; nb < NBU; ++nb) dv[nb][0] = dv[nb][1] = 0.0;
#pragma unroll
    for (int kb = 0; kb < NBU; ++kb)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double a = tu[kb][i] - acc1[kb][i];
        const double *base = fr + (size_t)(L.bwd_s2 + (2 * kb + i) * NBU) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < NBU; ++nb) spk_dmma(dv[nb][0], dv[nb][1], a, base[nb * 32]);
      }
    if (valid) {
#pragma unroll
      for (int nb = 0; nb < NBU; ++nb) spk_st2(A.dd + i0 * NU, 8 * nb + 2 * kk, NU, vec2, dv[nb][0], dv[nb][1]);
    }
    // ---- + G d + K'(d - ubar) (:610-632)
#pragma unroll
    for (int kb = 0; kb < NBU; ++kb)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double a1 = dv[kb][i], a2 = dv[kb][i] - tu[kb][i];
        const double *bg = fr + (size_t)(L.bwd_s3b + (2 * kb + i) * NBX) * 32 + lane;
        const double *bk = fr + (size_t)(L.bwd_s3c + (2 * kb + i) * NBX) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < NBX; ++nb) {
          spk_dmma(acc3[nb][0], acc3[nb][1], a1, bg[nb * 32]);
          spk_dmma(acc3[nb][0], acc3[nb][1], a2, bk[nb * 32]);
        }
      }
    // ---- q_i (:634-636)
    if (valid) {
#pragma unroll
      for (int nb = 0; nb < NBX; ++nb) spk_st2(A.q + i0 * NX, 8 * nb + 2 * kk, NX, vec2, acc3[nb][0], acc3[nb][1]);
    }
  }
}

template <int NX, int NU, int DD, bool LEAVES>
__launch_bounds__(256, spk_mm_minb(spk_mm_layout(NX, NU, DD).bwd_frags))
__global__ void spk_k_bwd_mm(SpkDims D, SpkPtrs P, SpkFusedArgs A, const double *__restrict__ frag_g) {
  constexpr SpkMmLayout L = spk_mm_layout(NX, NU, DD);
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpb = blockDim.x >> 5;
  const long count = A.p_cnt, first0 = spk_level_first0(DD, A.t) + A.p_lo, ntiles = (count + 7) / 8;
  if (A.reset_xi && blockIdx.x == 0 && tid == 0) {
    A.xi[0] = 0.0;
    A.xi[1] = 0.0;
  }
  spk_mm_fetch(smem, frag_g, L.bwd_frags, tid);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  bool waited = false;
  // (an L2 prefetch of the next tile's operand lines was measured slower -- 1.04 -> 1.19 ms per pass at config 3 -- and dropped)
  for (long tile = (long)blockIdx.x * wpb + warp; tile < ntiles; tile += (long)gridDim.x * wpb)
    spk_bwd_mm_tile<NX, NU, DD, LEAVES>(D, P, A, smem + 2, bar, 0, waited, first0, count, tile, lane);
  if (!waited) spk_mbar_wait(bar, 0);  // never exit with a bulk copy in flight
}

// ------------------------------------------------------------------------------------------------ forward level
template <int NX, int NU, int DD>
__device__ __forceinline__ void spk_fwd_mm_tile(const SpkDims &D, const SpkPtrs &P, const SpkFusedArgs &A, const double *fr,
                                                uint64_t *bar, uint32_t phase, bool &waited, long first0, long count, long tile,
                                                int lane, bool root) {
  constexpr SpkMmLayout L = spk_mm_layout(NX, NU, DD);
  constexpr int KX = L.KX, KU = L.KU, NBX = L.NBX, NBU = L.NBU;
  const int m = lane >> 2, kk = lane & 3;
  {
    long fm = tile * 8 + m;
    const bool valid = fm < count;
    if (!valid) fm = count - 1;
    const long i0 = first0 + fm, cfirst = (long)DD * i0 + 1;
    // ---- x_i (:667: x0 at the root) and d_i as A fragments; d_i again in C-fragment order as the start value of u_i
    const bool vec2 = A.vec2 != 0;
    double xa[KX], da[KU], accu[NBU][2];
    spk_ld_afrag<KX>(root ? P.x0 : A.zbar + i0 * NX, NX, kk, vec2, xa);
    spk_ld_afrag<KU>(A.dd + i0 * NU, NU, kk, vec2, da);
#pragma unroll
    for (int nb = 0; nb < NBU; ++nb) spk_ld2(A.dd + i0 * NU, 8 * nb + 2 * kk, NU, vec2, accu[nb][0], accu[nb][1]);
    if (root && valid) {
#pragma unroll
      for (int s = 0; s < KX; ++s) {
        const int k = spk_mm_kidx(s, kk, KX);
        if (k < NX) A.zbar[k] = xa[s];
      }
    }
    if (!waited) {
      spk_mbar_wait(bar, phase);
      waited = true;
    }
    // ---- u_i = K x_i + d_i (:670-673)
#pragma unroll
    for (int s = 0; s < KX; ++s) {
      const double *base = fr + (size_t)(L.fwd_k + s * NBU) * 32 + lane;
#pragma unroll
      for (int nb = 0; nb < NBU; ++nb) spk_dmma(accu[nb][0], accu[nb][1], xa[s], base[nb * 32]);
    }
    if (valid) {
#pragma unroll
      for (int nb = 0; nb < NBU; ++nb) spk_st2(A.zbar + D.zo[1] + i0 * NU, 8 * nb + 2 * kk, NU, vec2, accu[nb][0], accu[nb][1]);
    }
    // ---- x_c = ABK_c x_i + B_c d_i (:675-679)
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      double accx[NBX][2];
#pragma unroll
      for (int nb = 0; nb < NBX; ++nb) accx[nb][0] = accx[nb][1] = 0.0;
#pragma unroll
      for (int s = 0; s < KX; ++s) {
        const double *base = fr + (size_t)(L.fwd_abk + (c * KX + s) * NBX) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < NBX; ++nb) spk_dmma(accx[nb][0], accx[nb][1], xa[s], base[nb * 32]);
      }
#pragma unroll
      for (int s = 0; s < KU; ++s) {
        const double *base = fr + (size_t)(L.fwd_b + (c * KU + s) * NBX) * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < NBX; ++nb) spk_dmma(accx[nb][0], accx[nb][1], da[s], base[nb * 32]);
      }
      if (valid) {
#pragma unroll
        for (int nb = 0; nb < NBX; ++nb) spk_st2(A.zbar + (cfirst + c) * NX, 8 * nb + 2 * kk, NX, vec2, accx[nb][0], accx[nb][1]);
      }
    }
  }
}

template <int NX, int NU, int DD>
__launch_bounds__(256, spk_mm_minb(spk_mm_layout(NX, NU, DD).fwd_frags))
__global__ void spk_k_fwd_mm(SpkDims D, SpkPtrs P, SpkFusedArgs A, const double *__restrict__ frag_g) {
  constexpr SpkMmLayout L = spk_mm_layout(NX, NU, DD);
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpb = blockDim.x >> 5;
  const long count = A.p_cnt, first0 = spk_level_first0(DD, A.t) + A.p_lo, ntiles = (count + 7) / 8;
  spk_mm_fetch(smem, frag_g, L.fwd_frags, tid);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  bool waited = false;
  for (long tile = (long)blockIdx.x * wpb + warp; tile < ntiles; tile += (long)gridDim.x * wpb)
    spk_fwd_mm_tile<NX, NU, DD>(D, P, A, smem + 2, bar, 0, waited, first0, count, tile, lane, A.t == 0);
  if (!waited) spk_mbar_wait(bar, 0);
}

// ------------------------------------------------------------------------------------------------ top of the tree
// Levels t_top .. 0 of the backward sweep and 0 .. t_top of the forward sweep (at most 64 families each, one tile per
// warp) in ONE launch of a single CTA: a level of a small tree is far below the fixed cost of a kernel (launch, cold
// instruction cache of the unrolled DMMA chain, fragment fetch), which is what bounds every latency-bound configuration
// (MPC, scaling sweep).  Between levels: __syncthreads (the level's q / x are visible to the CTA) and a TMA
// refill of the fragment buffer with the next level's weights.
SPK_CONSTEXPR_HD int spk_mm_top_frags(int nx, int nu, int d) {
  return spk_mm_layout(nx, nu, d).bwd_frags > spk_mm_layout(nx, nu, d).fwd_frags ? spk_mm_layout(nx, nu, d).bwd_frags
                                                                                 : spk_mm_layout(nx, nu, d).fwd_frags;
}
template <int NX, int NU, int DD>
__launch_bounds__(256, 1)
__global__ void spk_k_top_mm(SpkDims D, SpkPtrs P, SpkFusedArgs A, const double *__restrict__ frag_bwd, const double *__restrict__ frag_fwd,
                             int t_top) {
  constexpr SpkMmLayout L = spk_mm_layout(NX, NU, DD);
  SPK_SMEM_DECL(double, smem);
  if (A.gate && *A.gate) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  if (tid == 0) spk_mbar_init(bar, 1);
  uint32_t phase = 0;
  for (int pass = 0; pass < 2; ++pass) {  // 0: backward levels t_top .. 0, 1: forward levels 0 .. t_top
    for (int step = 0; step <= t_top; ++step) {
      const int t = pass == 0 ? t_top - step : step;
      const int frags = pass == 0 ? L.bwd_frags : L.fwd_frags;
      const double *src = (pass == 0 ? frag_bwd : frag_fwd) + (size_t)t * frags * 32;
      __syncthreads();  // the previous level's results are visible; nobody reads the fragment buffer any more
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t total = (uint32_t)frags * 256u;
        spk_mbar_expect_tx(bar, total);
        for (uint32_t off = 0; off < total; off += 32768u) {
          const uint32_t sz = total - off < 32768u ? total - off : 32768u;
          spk_bulk_g2s(reinterpret_cast<char *>(smem + 2) + off, reinterpret_cast<const char *>(src) + off, sz, bar);
        }
      }
      const long count = spk_level_count(DD, t), first0 = spk_level_first0(DD, t);
      bool waited = false;
      if ((long)warp * 8 < count) {
        if (pass == 0) spk_bwd_mm_tile<NX, NU, DD, false>(D, P, A, smem + 2, bar, phase, waited, first0, count, warp, lane);
        else spk_fwd_mm_tile<NX, NU, DD>(D, P, A, smem + 2, bar, phase, waited, first0, count, warp, lane, t == 0);
      }
      if (!waited) spk_mbar_wait(bar, phase);
      phase ^= 1;
    }
  }
}
#endif
