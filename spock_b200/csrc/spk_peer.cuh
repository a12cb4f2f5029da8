// spk_peer.cuh -- the cut-level exchanges of the subtree-partitioned operator evaluation as ONE small kernel each, over
// peer memory (NVLink 5 / NVSwitch P2P stores) instead of host-issued NCCL collectives.
//
// Why: the messages are tiny (<= 3.2 kB: the subtree roots' q_j at the turning point of the backward sweep; the roots'
// vbar2_j and the residual pair after the dual kernel), so a collective costs its launch latency plus a gap in the stream
// -- and, being issued by the host between two CUDA graphs, it splits every evaluation into three pieces.  Here every rank
// STORES its slice straight into a mailbox in every peer's memory, publishes an epoch number with a system-scope release,
// and spins (system-scope acquire) until every peer's epoch has arrived in its own mailbox.  The kernel is an ordinary
// graph node, so a whole evaluation -- and a whole Chambolle-Pock pass including the max of (xi1, xi2) over the ranks and
// the termination test -- is one graph replay with no host involvement and no NCCL call.
//
// Mailbox layout per rank (doubles), all of it in one allocation that the peers map (cudaIpc handles between processes,
// plain peer access inside a device group):
//   [A | B | flags]   A = 2 parities x world slots x lenA   (lenA = m_sub * nx:  q of the rank's subtree roots)
//                     B = 2 parities x world slots x lenB   (lenB = m_sub + 2:   vbar2 of the roots, xi1, xi2)
//                     flags = 2 kinds x world epochs (uint64), then 2 local epoch counters, then 1 error word
// Epoch e of a kind uses parity e & 1: a rank can run at most one exchange ahead of a peer (it needs the peer's epoch e
// to finish exchange e), so the slot written for e + 1 is never the one a slow peer is still reading for e.
#pragma once
#include "spk_fused.cuh"

#define SPK_PEER_MAX 64

struct SpkPeerArgs {
  int rank, world, kind;         // kind 0 = A (q at the cut level), 1 = B (vbar2 + residual maxima)
  long len;                      // payload doubles per rank (lenA or lenB)
  long off_payload;              // offset of this kind's area in the mailbox (doubles)
  long off_flags;                // offset of the flag words (doubles == uint64 words)
  double *mine;                  // this rank's mailbox
  double *peer[SPK_PEER_MAX];    // every rank's mailbox as mapped here (peer[rank] == mine)
  const int *gate;               // the loop's `done` flag (identical on all ranks)
  // kind A: src/dst = q of the cut-level nodes (world * len doubles, rank-major); kind B: vbar2 of the cut-level nodes
  double *vec;                   // base of the rank-major array: slice r = vec + r * vlen
  long vlen;                     // m_sub * nx (A) or m_sub (B)
  double *xi;                    // kind B: d_scal[0..1] -- in: this rank's maxima, out: the maxima over all ranks
};

SPK_HD long spk_peer_mailbox_doubles(int world, long lenA, long lenB) {
  return 2L * world * lenA + 2L * world * lenB + 2L * world + 2 + 2;
}

#if !SPK_EMU
SPK_D void spk_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SPK_D unsigned long long spk_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128) spk_k_peer_exchange(SpkPeerArgs X) {
  if (X.gate && *X.gate) return;
  __shared__ unsigned long long s_epoch;
  __shared__ int s_fail;
  unsigned long long *flags = reinterpret_cast<unsigned long long *>(X.mine + X.off_flags);
  unsigned long long *epochs = flags + 2L * X.world;
  const int tid = threadIdx.x, T = blockDim.x;
  if (tid == 0) {
    s_epoch = epochs[X.kind] + 1;
    epochs[X.kind] = s_epoch;
    s_fail = 0;
  }
  __syncthreads();
  const unsigned long long e = s_epoch;
  const long slot = ((long)(e & 1ull) * X.world + X.rank) * X.len;  // where THIS rank's payload goes in every mailbox
  const double *src = X.vec + (long)X.rank * X.vlen;
  // 1. push: my slice (and, kind B, my residual maxima) into every peer's mailbox
  for (int r = 0; r < X.world; ++r) {
    if (r == X.rank) continue;
    double *dst = X.peer[r] + X.off_payload + slot;
    for (long k = tid; k < X.vlen; k += T) dst[k] = src[k];
    if (X.kind == 1 && tid < 2) dst[X.vlen + tid] = X.xi[tid];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish my epoch in every peer's flag word, then wait for every peer's epoch in mine
  if (tid < X.world && tid != X.rank) {
    unsigned long long *pf = reinterpret_cast<unsigned long long *>(X.peer[tid] + X.off_flags) + (long)X.kind * X.world + X.rank;
    spk_st_release_sys(pf, e);
    const unsigned long long *mf = flags + (long)X.kind * X.world + tid;
    const long long t0 = clock64();
    while (spk_ld_acquire_sys(mf) < e) {
      if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer died; fail loudly instead of hanging the GPU
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) flags[2L * X.world + 2] = 1ull;  // error word, read by the host when it polls the loop
    return;
  }
  // 3. collect: the peers' slices out of my mailbox (L2 reads: the lines were written remotely)
  double m0 = 0.0, m1 = 0.0;
  for (int r = 0; r < X.world; ++r) {
    if (r == X.rank) continue;
    const double *box = X.mine + X.off_payload + ((long)(e & 1ull) * X.world + r) * X.len;
    double *dst = X.vec + (long)r * X.vlen;
    for (long k = tid; k < X.vlen; k += T) dst[k] = __ldcg(box + k);
    if (X.kind == 1 && tid == 0) {
      m0 = spk_max_nn(m0, __ldcg(box + X.vlen));
      m1 = spk_max_nn(m1, __ldcg(box + X.vlen + 1));
    }
  }
  if (X.kind == 1 && tid == 0) {  // residual maxima over all ranks: every rank ends with the same pair
    X.xi[0] = spk_max_nn(X.xi[0], m0);
    X.xi[1] = spk_max_nn(X.xi[1], m1);
  }
}
#endif
