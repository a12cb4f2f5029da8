"""Helpers for the subtree-partition tests: which entries of (z, v) a set of families writes, and assembling
the ranks' local views into one iterate."""
import numpy as np


def family_masks(flat, z_off, v_off, ranges):
    """Boolean masks over z and v of the entries written by the families of the non-leaf nodes in `ranges`
    ([(first 1-based node, count), ...])."""
    nx, nu, ny, d = flat.nx, flat.nu, flat.ny, flat.d
    n, nnl = flat.n, flat.n_non_leaf
    mz, mv = np.zeros(flat.nz, bool), np.zeros(flat.nv, bool)
    for first, count in ranges:
        for i0 in range(first - 1, first - 1 + count):
            c0 = d * i0 + 1
            mz[z_off[1] + i0 * nu: z_off[1] + (i0 + 1) * nu] = True          # u_i
            mz[z_off[4] + i0 * ny: z_off[4] + (i0 + 1) * ny] = True          # y_i
            mz[c0 * nx: (c0 + d) * nx] = True                                # x_children
            mz[z_off[2] + c0: z_off[2] + c0 + d] = True                      # s_children
            mz[z_off[3] + c0 - 1: z_off[3] + c0 - 1 + d] = True              # tau_children
            if i0 == 0:
                mz[0:nx] = True
                mz[z_off[2]] = True
            mv[i0 * ny: (i0 + 1) * ny] = True                                # v1_i
            mv[v_off[1] + i0] = True                                         # v2_i
            mv[v_off[6] + i0 * (nx + nu): v_off[6] + (i0 + 1) * (nx + nu)] = True
            mv[v_off[2] + (c0 - 1) * nx: v_off[2] + (c0 - 1 + d) * nx] = True
            mv[v_off[3] + (c0 - 1) * nu: v_off[3] + (c0 - 1 + d) * nu] = True
            mv[v_off[4] + c0 - 1: v_off[4] + c0 - 1 + d] = True
            mv[v_off[5] + c0 - 1: v_off[5] + c0 - 1 + d] = True
            if c0 >= nnl:                                                    # leaf children
                l0 = c0 - nnl
                mv[v_off[7] + l0 * nx: v_off[7] + (l0 + d) * nx] = True
                mv[v_off[8] + l0: v_off[8] + l0 + d] = True
                mv[v_off[9] + l0: v_off[9] + l0 + d] = True
                mv[v_off[10] + l0 * nx: v_off[10] + (l0 + d) * nx] = True
    return mz, mv


def assemble(flat, z_off, v_off, views):
    """views: [(ranges, z, v)] per rank -> (z, v, max disagreement on entries several ranks hold)."""
    z, v = np.full(flat.nz, np.nan), np.full(flat.nv, np.nan)
    worst = 0.0
    for ranges, zg, vg in views:
        mz, mv = family_masks(flat, z_off, v_off, ranges)
        for full, part, m in ((z, zg, mz), (v, vg, mv)):
            both = m & ~np.isnan(full)
            if both.any():
                worst = max(worst, float(np.max(np.abs(full[both] - part[both]))))
            full[m] = part[m]
    assert not np.isnan(z).any() and not np.isnan(v).any(), "the ranks' families do not cover the tree"
    return z, v, worst
