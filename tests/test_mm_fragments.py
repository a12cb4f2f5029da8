"""CPU check of the host-side logic behind the tensor-pipe S1 kernels (spk_mm.cuh): the fragment-ordered weight tables
(spk_mm_fragment_tables / spk_mm_layout.h: natural-paired and permuted inner orderings, zero padding) are pulled out of the
host-emulation build and a numpy model of mma.sync.m8n8k4.f64 replays the kernels' DMMA chains for one tile of 8 families;
the results must equal the dense formulas of projection_S1! (implicit_l.jl:584-637, :667-681) built from the factors the
library reports.  The CUDA kernels themselves are covered by the -m gpu parity tests."""
import ctypes as C

import numpy as np
import pytest

from spock_b200.examples import get_random_system_specs
from spock_b200.handle import Handle
from spock_b200.problem import flatten_problem


def kidx(s, kk, steps):  # spk_mm_kidx
    return 4 * s + kk if (steps & 1) and s == steps - 1 else 8 * (s >> 1) + 2 * kk + (s & 1)


def dmma(acc, a, frag):
    """acc (8 families x 8 rows) += A (8 x 4) @ B (4 x 8); lane T = 4 m + kk holds a[T] = A[m, kk] and frag[T] = B[kk, n = T // 4]."""
    return acc + a.reshape(8, 4) @ frag.reshape(8, 4).T


def afrag(vecs, s, steps, n):
    """A operand of k-step s for 8 vectors (8 x n): lane (m, kk) holds entry kidx(s, kk) of vector m (0 beyond n)."""
    out = np.zeros(32)
    for m in range(8):
        for kk in range(4):
            k = kidx(s, kk, steps)
            out[4 * m + kk] = vecs[m, k] if k < n else 0.0
    return out


def afrag_from_acc(acc_block, i):
    """accumulator registers (rows 2 kk + i of an n-block) fed back as the A operand of permuted k-step 2 nb + i"""
    out = np.zeros(32)
    for m in range(8):
        for kk in range(4):
            out[4 * m + kk] = acc_block[m, 2 * kk + i]
    return out


@pytest.mark.parametrize("nx,nu,d", [(50, 20, 2), (10, 10, 4), (5, 5, 2), (20, 20, 3), (7, 3, 2)])
def test_fragment_tables_reproduce_the_s1_steps(emu_lib, nx, nu, d):
    N = 4
    flat = flatten_problem(*get_random_system_specs(N, d, nx, nu))
    h = Handle(flat, 1, factor_mode=1, lib=emu_lib)
    emu_lib.spk_emu_mm_tables.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    cnt = (C.c_int64 * 16)()
    assert emu_lib.spk_emu_mm_tables(h._h, None, None, cnt) == 0
    (bwd_frags, fwd_frags, KX, KU, KUP, NBX, NBU, o_sweep, o_s2, o_s3b, o_s3c, o_fk, o_fabk, o_fb, NN, _) = [int(c) for c in cnt]
    assert NN == N and KX == -(-nx // 4) and NBX == -(-nx // 8) and NBU == -(-nu // 8) and KUP == 2 * NBU
    # every inner index is covered exactly once by the paired ordering
    for steps, n in ((KX, nx), (KU, nu)):
        seen = sorted(kidx(s, kk, steps) for s in range(steps) for kk in range(4))
        assert seen == list(range(4 * steps)) and 4 * steps >= n
    fb, ff = np.zeros((N - 1) * bwd_frags * 32), np.zeros((N - 1) * fwd_frags * 32)
    dp = C.POINTER(C.c_double)
    assert emu_lib.spk_emu_mm_tables(h._h, fb.ctypes.data_as(dp), ff.ctypes.data_as(dp), cnt) == 0
    fb, ff = fb.reshape(N - 1, bwd_frags, 32), ff.reshape(N - 1, fwd_frags, 32)
    B = [flat.B[k * nx * nu:(k + 1) * nx * nu].reshape(nu, nx).T for k in range(d)]
    rng = np.random.default_rng(nx * 100 + nu)
    for t in range(N - 1):
        i = (d**t - 1) // (d - 1) + 1  # first node (1-based) of level t; factors depend on the level only
        _, _, K, Rinv, _ = h.factors(i)
        ch = [d * (i - 1) + 2 + c for c in range(d)]
        Pc = [h.factors(j)[0] for j in ch]
        ABK = [h.factors(j)[1] for j in ch]
        Bc = [B[(c + 2) % d] for c in range(d)]  # mode of child ordinal c
        G = sum(ABK[c].T @ Pc[c] @ Bc[c] for c in range(d))
        # ---- backward tile: rhs -> d -> q (spk_bwd_mm_tile)
        qc = rng.standard_normal((d, 8, nx))
        tu, tx = rng.standard_normal((8, nu)), rng.standard_normal((8, nx))
        acc1 = [np.zeros((8, 8)) for _ in range(NBU)]
        acc3 = [np.zeros((8, 8)) for _ in range(NBX)]
        for nb in range(NBX):
            for m in range(8):
                for n in range(8):
                    if 8 * nb + n < nx:
                        acc3[nb][m, n] = -tx[m, 8 * nb + n]
        for c in range(d):
            for s in range(KX):
                a = afrag(qc[c], s, KX, nx)
                base = o_sweep + (c * KX + s) * (NBU + NBX)
                for nb in range(NBU):
                    acc1[nb] = dmma(acc1[nb], a, fb[t, base + nb])
                for nb in range(NBX):
                    acc3[nb] = dmma(acc3[nb], a, fb[t, base + NBU + nb])
        tu_pad = np.zeros((8, 8 * NBU))
        tu_pad[:, :nu] = tu
        rhs = [tu_pad[:, 8 * nb:8 * nb + 8] - acc1[nb] for nb in range(NBU)]
        dv = [np.zeros((8, 8)) for _ in range(NBU)]
        for kb in range(NBU):
            for i2 in range(2):
                a = afrag_from_acc(rhs[kb], i2)
                for nb in range(NBU):
                    dv[nb] = dmma(dv[nb], a, fb[t, o_s2 + (2 * kb + i2) * NBU + nb])
        for kb in range(NBU):
            for i2 in range(2):
                a1 = afrag_from_acc(dv[kb], i2)
                a2 = afrag_from_acc(dv[kb] - tu_pad[:, 8 * kb:8 * kb + 8], i2)
                for nb in range(NBX):
                    acc3[nb] = dmma(acc3[nb], a1, fb[t, o_s3b + (2 * kb + i2) * NBX + nb])
                    acc3[nb] = dmma(acc3[nb], a2, fb[t, o_s3c + (2 * kb + i2) * NBX + nb])
        d_mm = np.concatenate(dv, axis=1)[:, :nu]
        q_mm = np.concatenate(acc3, axis=1)[:, :nx]
        rhs_ref = tu - sum(qc[c] @ Bc[c] for c in range(d))             # rows: families;  B_c' q_c
        d_ref = rhs_ref @ Rinv.T
        q_ref = sum(qc[c] @ ABK[c] for c in range(d)) + d_ref @ G.T + (d_ref - tu) @ K - tx
        assert np.max(np.abs(d_mm - d_ref)) < 1e-11 * max(1.0, np.max(np.abs(d_ref)))
        assert np.max(np.abs(q_mm - q_ref)) < 1e-11 * max(1.0, np.max(np.abs(q_ref)))
        assert np.max(np.abs(np.concatenate(dv, axis=1)[:, nu:])) == 0.0  # padding rows stay exactly zero
        # ---- forward tile: u = K x + d, x_c = ABK_c x + B_c d (spk_fwd_mm_tile)
        x, dd = rng.standard_normal((8, nx)), rng.standard_normal((8, nu))
        dd_pad = np.zeros((8, 8 * NBU))
        dd_pad[:, :nu] = dd
        accu = [dd_pad[:, 8 * nb:8 * nb + 8].copy() for nb in range(NBU)]
        for s in range(KX):
            a = afrag(x, s, KX, nx)
            for nb in range(NBU):
                accu[nb] = dmma(accu[nb], a, ff[t, o_fk + s * NBU + nb])
        u_mm = np.concatenate(accu, axis=1)[:, :nu]
        assert np.max(np.abs(u_mm - (x @ K.T + dd))) < 1e-11 * max(1.0, np.max(np.abs(u_mm)))
        for c in range(d):
            accx = [np.zeros((8, 8)) for _ in range(NBX)]
            for s in range(KX):
                a = afrag(x, s, KX, nx)
                for nb in range(NBX):
                    accx[nb] = dmma(accx[nb], a, ff[t, o_fabk + (c * KX + s) * NBX + nb])
            for s in range(KU):
                a = afrag(dd, s, KU, nu)
                for nb in range(NBX):
                    accx[nb] = dmma(accx[nb], a, ff[t, o_fb + (c * KU + s) * NBX + nb])
            xc_mm = np.concatenate(accx, axis=1)[:, :nx]
            xc_ref = x @ ABK[c].T + dd @ Bc[c].T
            assert np.max(np.abs(xc_mm - xc_ref)) < 1e-11 * max(1.0, np.max(np.abs(xc_ref)))
    h.close()
