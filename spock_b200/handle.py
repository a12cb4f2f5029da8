"""Thin object wrapper over the C ABI (one `spk_handle`).  Host vectors are numpy float64 arrays in the
reference layout; all arithmetic happens in libspock_b200.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (CP, FACTORS_PER_NODE, FACTORS_PER_STAGE, OP_L, OP_LT, OP_PROJ_S1, OP_PROJ_S2, OP_PROX_F,
                   OP_PROX_HCONJ, SP, SpkResult, check, dptr)

PROFILE_CLASSES = ["s1_backward", "s1_forward", "L_Lt", "prox_hconj", "s2", "vector_ops", "fused_backward",
                   "fused_forward"]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    def __init__(self, flat, algorithm: int = CP, L_norm: float = 0.0, factor_mode: int = FACTORS_PER_NODE,
                 device: int = -1, rank: int = 0, world_size: int = 1, comm_id=None, lib=None, devices=None):
        """devices: list of CUDA ordinals -> the tree is partitioned over them inside this one handle / process."""
        self.lib = lib if lib is not None else _lib.load()
        self.flat = flat
        self.algorithm = algorithm
        prob, self._keep = _lib.make_problem(flat, algorithm, L_norm, factor_mode, device, rank, world_size, comm_id, devices)
        self._h = C.c_void_p()
        check(self.lib, self.lib.spk_create(C.byref(prob), C.byref(self._h)))
        s = (C.c_int64 * 5)()
        check(self.lib, self.lib.spk_sizes(self._h, s))
        self.n, self.n_non_leaf, self.n_leaf, self.nz, self.nv = (int(x) for x in s)
        zo, vo = (C.c_int64 * 5)(), (C.c_int64 * 11)()
        check(self.lib, self.lib.spk_offsets(self._h, zo, vo))
        self.z_off, self.v_off = [int(x) for x in zo], [int(x) for x in vo]
        self.L_norm = L_norm if L_norm > 0 else (3.28 if algorithm == SP else 10.0)
        self.rank, self.world_size = rank, max(1, world_size)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.spk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def set_x0(self, x0):
        x0 = _f64(x0)
        if x0.shape != (self.flat.nx,):
            raise ValueError("x0 has the wrong length")
        check(self.lib, self.lib.spk_set_x0(self._h, dptr(x0)))

    def set_iterate(self, z=None, v=None):
        z = None if z is None else _f64(z)
        v = None if v is None else _f64(v)
        if (z is not None and z.shape != (self.nz,)) or (v is not None and v.shape != (self.nv,)):
            raise ValueError("z or v has the wrong length")  # implicit_l.jl:184-186
        check(self.lib, self.lib.spk_set_iterate(self._h, dptr(z), dptr(v)))

    def get_iterate(self, z=None, v=None):
        """(z, v) in the reference layout.  Under the subtree partition (world_size > 1) only the entries this rank's
        families own are copied -- the rest keeps the caller's values (zeros in fresh arrays) -- unless gather_iterate()
        was called since the last pass."""
        alloc = np.empty if self.world_size == 1 else np.zeros
        z = alloc(self.nz) if z is None else z
        v = alloc(self.nv) if v is None else v
        check(self.lib, self.lib.spk_get_iterate(self._h, dptr(z), dptr(v)))
        return z, v

    def gather_iterate(self):
        """Collective over all ranks: afterwards get_iterate() returns the complete (z, v) on every rank."""
        check(self.lib, self.lib.spk_gather_iterate(self._h))

    def get_state(self):
        """Solver memory that persists across solves besides (z, v) (checkpoint / resume; layout: include/spock_b200.h)."""
        n = C.c_int64()
        check(self.lib, self.lib.spk_state_size(self._h, C.byref(n)))
        st = np.empty(int(n.value))
        check(self.lib, self.lib.spk_get_state(self._h, dptr(st)))
        return st

    def set_state(self, state):
        n = C.c_int64()
        check(self.lib, self.lib.spk_state_size(self._h, C.byref(n)))
        state = _f64(state)
        if state.shape != (int(n.value),):
            raise ValueError("state has the wrong length")
        check(self.lib, self.lib.spk_set_state(self._h, dptr(state)))

    def set_sp_options(self, linesearch: bool = False, k0: bool = False):
        """SuperMann variants present but not executed in the reference (sp.jl:439, sp.jl:80); default = as shipped."""
        check(self.lib, self.lib.spk_set_sp_options(self._h, (1 if linesearch else 0) | (2 if k0 else 0)))

    def shift_iterate(self, child: int):
        """Shift-initialisation of the next receding-horizon solve (`child`: 0-based ordinal of the realised child of the root)."""
        check(self.lib, self.lib.spk_shift_iterate(self._h, int(child)))

    def estimate_L_norm(self, iters: int = 0, apply: bool = False) -> float:
        """||L||^2 by power iteration on the device; apply=True makes it the value behind the default step sizes."""
        out = C.c_double()
        check(self.lib, self.lib.spk_estimate_L_norm(self._h, int(iters), C.byref(out)))
        if apply:
            check(self.lib, self.lib.spk_set_L_norm(self._h, out.value))
            self.L_norm = out.value
        return out.value

    # ------------------------------------------------------------------ operators
    _IN_LEN = {OP_L: "nz", OP_LT: "nv", OP_PROX_F: "nz", OP_PROX_HCONJ: "nv"}
    _OUT_LEN = {OP_L: "nv", OP_LT: "nz", OP_PROX_F: "nz", OP_PROX_HCONJ: "nv"}

    def apply(self, op: int, vec, scalar: float = 0.0):
        vec = _f64(vec)
        if op == OP_PROJ_S1:
            n_in = n_out = self.z_off[2]
        elif op == OP_PROJ_S2:
            n_in = n_out = self.nz - self.z_off[2] - 1
        else:
            n_in, n_out = getattr(self, self._IN_LEN[op]), getattr(self, self._OUT_LEN[op])
        if vec.shape != (n_in,):
            raise ValueError("z or v has the wrong length")  # implicit_l.jl:184-186, 327-329
        out = np.empty(n_out)
        check(self.lib, self.lib.spk_apply(self._h, op, dptr(vec), dptr(out), float(scalar)))
        return out

    def L(self, z):
        return self.apply(OP_L, z)

    def Lt(self, v):
        return self.apply(OP_LT, v)

    def mul(self, C_vec, trans: bool, B, alpha: float, beta: float):
        """spock_mul!: C <- alpha * op(L) B + beta * C, in place on the numpy vector C_vec."""
        B = _f64(B)
        if B.shape != ((self.nv,) if trans else (self.nz,)) or C_vec.shape != ((self.nz,) if trans else (self.nv,)):
            raise ValueError("z or v has the wrong length")
        check(self.lib, self.lib.spk_mul(self._h, int(bool(trans)), dptr(B), float(alpha), float(beta), dptr(C_vec)))

    def prox_f(self, arg, gamma):
        return self.apply(OP_PROX_F, arg, gamma)

    def prox_h_conj(self, arg, sigma):
        return self.apply(OP_PROX_HCONJ, arg, sigma)

    def projection_S1(self, xu):
        return self.apply(OP_PROJ_S1, xu)

    def projection_S2(self, z2):
        return self.apply(OP_PROJ_S2, z2)

    def dot(self, a_z, a_v, b_z, b_v, gamma, sigma) -> float:
        arrs = [_f64(a) for a in (a_z, a_v, b_z, b_v)]
        out = C.c_double()
        check(self.lib, self.lib.spk_dot(self._h, *[dptr(a) for a in arrs], gamma, sigma, C.byref(out)))
        return out.value

    def factors(self, i: int):
        """(P, ABK, K, Rinv, ls_matrix) of node i (1-based), as 2-D arrays."""
        nx, nu, nls = self.flat.nx, self.flat.nu, self.flat.ny + 2 * self.flat.d
        P, ABK, K, Ri, ls = np.zeros(nx * nx), np.zeros(nx * nx), np.zeros(nu * nx), np.zeros(nu * nu), np.zeros(nls * nls)
        check(self.lib, self.lib.spk_get_factors(self._h, i, dptr(P), dptr(ABK), dptr(K), dptr(Ri), dptr(ls)))
        return (P.reshape(nx, nx).T, ABK.reshape(nx, nx).T, K.reshape(nx, nu).T, Ri.reshape(nu, nu).T,
                ls.reshape(nls, nls).T)

    # ------------------------------------------------------------------ loops
    def step(self, K: int, gamma: float = 0.0, sigma: float = 0.0) -> float:
        """Exactly K passes; returns the device time in ms."""
        ms = C.c_float()
        check(self.lib, self.lib.spk_step(self._h, int(K), gamma, sigma, C.byref(ms)))
        return ms.value

    def residuals(self):
        out = np.zeros(2)
        check(self.lib, self.lib.spk_get_residuals(self._h, dptr(out)))
        return float(out[0]), float(out[1])

    def solve(self, x0, tol=1e-3, max_iter=0, gamma=0.0, sigma=0.0, verbose_log=False, z0=None, v0=None):
        x0 = _f64(x0)
        if x0.shape != (self.flat.nx,):
            raise ValueError("x0 has the wrong length")
        cap = max_iter if max_iter > 0 else (1000 if self.algorithm == SP else 5000)
        xi, xi1, xi2 = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
        bc = np.zeros(cap, np.int32)
        z0 = None if z0 is None else _f64(z0)
        v0 = None if v0 is None else _f64(v0)
        res = SpkResult()
        check(self.lib, self.lib.spk_solve(self._h, dptr(x0), tol, max_iter, gamma, sigma, int(bool(verbose_log)),
                                           dptr(z0), dptr(v0), C.byref(res), dptr(xi), dptr(xi1), dptr(xi2),
                                           bc.ctypes.data_as(C.POINTER(C.c_int32)), cap))
        k = res.passes
        out = {f: getattr(res, f) for f, _ in SpkResult._fields_}
        return out, {"xi": xi[:k], "xi1": xi1[:k], "xi2": xi2[:k], "backtrack_count": bc[:k]}

    # ------------------------------------------------------------------ instrumentation
    def profile(self, enable: bool, reset: bool = True):
        check(self.lib, self.lib.spk_profile(self._h, int(enable), int(reset)))

    def profile_report(self):
        out = {}
        for cid, name in enumerate(PROFILE_CLASSES):
            ms, n = C.c_double(), C.c_int64()
            check(self.lib, self.lib.spk_profile_get(self._h, cid, C.byref(ms), C.byref(n)))
            out[name] = {"ms": ms.value, "launches": int(n.value)}
        return out

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self.lib, self.lib.spk_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def owned_ranges(self):
        """[(first node (1-based), count), ...] of the non-leaf nodes whose families this rank computes."""
        n = C.c_int32()
        st, ct = (C.c_int64 * 64)(), (C.c_int64 * 64)()
        check(self.lib, self.lib.spk_owned_ranges(self._h, C.byref(n), st, ct))
        return [(int(st[k]), int(ct[k])) for k in range(n.value)]

    def device_bytes(self) -> int:
        n = C.c_int64()
        check(self.lib, self.lib.spk_device_bytes(self._h, C.byref(n)))
        return int(n.value)


def solve_batch(handles, x0s, tol=1e-3, max_iter=0):
    """solve_model! for several independent models at once (their device-resident loops overlap on the GPU).
    x0s: one initial state per handle.  Returns the list of result dicts."""
    lib = handles[0].lib
    n = len(handles)
    arr = (C.c_void_p * n)(*[h._h for h in handles])
    x = np.ascontiguousarray(np.concatenate([_f64(x0) for x0 in x0s]))
    res = (SpkResult * n)()
    check(lib, lib.spk_solve_batch(arr, n, dptr(x), float(tol), int(max_iter), res))
    return [{f: getattr(res[k], f) for f, _ in SpkResult._fields_} for k in range(n)]


def tree_maps(N: int, d: int, lib=None):
    """(child_first, anc, w, min_index_per_timestep) from the library, reference 1-based values."""
    lib = lib if lib is not None else _lib.load()
    n = (d**N - 1) // (d - 1)
    nnl = (d ** (N - 1) - 1) // (d - 1)
    cf = np.zeros(max(nnl, 1), np.int64)
    anc = np.zeros(n, np.int64)
    w = np.zeros(n, np.int32)
    mi = np.zeros(N, np.int64)
    i64, i32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    check(lib, lib.spk_tree_maps(N, d, cf.ctypes.data_as(i64), anc.ctypes.data_as(i64), w.ctypes.data_as(i32),
                                 mi.ctypes.data_as(i64)))
    return cf[:nnl], anc, w, mi
