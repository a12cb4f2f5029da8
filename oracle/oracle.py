"""ctypes wrapper around the C CPU oracle (oracle/spock_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (spock_b200/) must never import this module.

The only arithmetic done here (in numpy) is the offline kernel projector of cpock.jl:71-81,125-128,
restated as `I - pinv(M) M` (the orthogonal projector on ker M, identical to the reference's
`svd(nullspace(M)).U * pinv(svd(nullspace(M)).U)`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

CP, SP = 1, 2


class _Problem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("d", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("ny", C.c_int),
        ("A", C.POINTER(C.c_double)), ("B", C.POINTER(C.c_double)),
        ("cost_dense", C.c_int), ("n_sqrtQ", C.c_int),
        ("sqrtQ", C.POINTER(C.c_double)), ("sqrtR", C.POINTER(C.c_double)),
        ("n_sqrtQN", C.c_int), ("sqrtQN", C.POINTER(C.c_double)),
        ("E", C.POINTER(C.c_double)), ("F", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)),
        ("n_cones", C.c_int), ("cone_type", C.POINTER(C.c_int)), ("cone_dim", C.POINTER(C.c_int)),
        ("ls_matrix", C.POINTER(C.c_double)),
        ("x_min", C.c_double), ("x_max", C.c_double), ("u_min", C.c_double), ("u_max", C.c_double),
        ("algorithm", C.c_int), ("L_norm", C.c_double), ("w", C.POINTER(C.c_int)), ("n_rms", C.c_int),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("passes", C.c_int), ("printed_iter", C.c_int), ("converged", C.c_int),
        ("xi1", C.c_double), ("xi2", C.c_double), ("xi", C.c_double),
        ("k1_count", C.c_int), ("k2_count", C.c_int), ("fallback_count", C.c_int),
    ]


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (gcc only)."""
    so = os.path.join(_HERE, "_build", "libspock_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "spock_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def _load(omp: bool):
    key = "omp" if omp else "st"
    if key not in _LIBS:
        build()
        name = "libspock_oracle_omp.so" if omp else "libspock_oracle.so"
        lib = C.CDLL(os.path.join(_HERE, "_build", name))
        dp, lp, ip = C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_int)
        lib.spo_create.restype = C.c_void_p
        lib.spo_create.argtypes = [C.POINTER(_Problem), C.c_int]
        lib.spo_destroy.argtypes = [C.c_void_p]
        for f in ("spo_nz", "spo_nv", "spo_n"):
            getattr(lib, f).restype = C.c_long
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.spo_tree_maps.argtypes = [C.c_void_p, lp, lp, ip, lp]
        lib.spo_offsets.argtypes = [C.c_void_p, lp, lp]
        lib.spo_z.restype = dp
        lib.spo_z.argtypes = [C.c_void_p]
        lib.spo_v.restype = dp
        lib.spo_v.argtypes = [C.c_void_p]
        lib.spo_set_x0.argtypes = [C.c_void_p, dp]
        lib.spo_L.argtypes = [C.c_void_p, dp, dp]
        lib.spo_Lt.argtypes = [C.c_void_p, dp, dp]
        lib.spo_mul.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_double, C.c_double]
        lib.spo_dot.restype = C.c_double
        lib.spo_dot.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_double]
        lib.spo_projection_S1.argtypes = [C.c_void_p, dp]
        lib.spo_projection_S2.argtypes = [C.c_void_p, dp]
        lib.spo_prox_f.argtypes = [C.c_void_p, dp, C.c_double]
        lib.spo_prox_h_conj.argtypes = [C.c_void_p, dp, C.c_double]
        lib.spo_project_soc.argtypes = [dp, C.c_long]
        lib.spo_factors.argtypes = [C.c_void_p, C.c_long, dp, dp, dp, dp]
        lib.spo_solve.restype = C.c_int
        lib.spo_solve.argtypes = [C.c_void_p, dp, C.c_double, C.c_int, dp, dp, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(_Result), dp, dp, dp, ip, C.c_int]
        lib.spo_set_sp_options.argtypes = [C.c_void_p, C.c_int]
        lib.spo_step.restype = C.c_int
        lib.spo_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        _LIBS[key] = lib
    return _LIBS[key]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def kernel_projector(E2d: np.ndarray, F2d: np.ndarray) -> np.ndarray:
    """cpock.jl:71-81 (M) and :125-128 (projector on ker M)."""
    ny, d = E2d.shape
    M = np.vstack([
        np.hstack([E2d.T, -np.eye(d), -np.eye(d)]),
        np.hstack([F2d.T, np.zeros((F2d.shape[1], d)), np.zeros((F2d.shape[1], d))]),
    ])
    return np.eye(ny + 2 * d) - np.linalg.pinv(M) @ M


class Oracle:
    """One reference model (`build_model` result) living in the C oracle."""

    def __init__(self, flat, algorithm: int = CP, L_norm: float = 0.0, n_threads: int = 1):
        self.lib = _load(n_threads > 1)
        self.flat = flat
        self.algorithm = algorithm
        n_rms = int(getattr(flat, "n_rms", 1))
        if n_rms > 1:  # one projector per non-leaf node (cpock.jl:71-81,125-128 loops over the nodes)
            ls = np.concatenate([np.ascontiguousarray(kernel_projector(np.asarray(r.E, float), np.asarray(r.F, float)).T).reshape(-1)
                                 for r in flat.rms_list])
        else:
            ls = np.ascontiguousarray(kernel_projector(flat.E2d, flat.F2d).T).reshape(-1)
        self._keep = [ls, flat.cone_type.astype(np.int32), flat.cone_dim.astype(np.int32)]
        ip = C.POINTER(C.c_int)
        p = _Problem(
            flat.N, flat.d, flat.nx, flat.nu, flat.ny, _dp(flat.A), _dp(flat.B), flat.cost_dense, flat.n_sqrtQ,
            _dp(flat.sqrtQ), _dp(flat.sqrtR), flat.n_sqrtQN, _dp(flat.sqrtQN), _dp(flat.E), _dp(flat.F), _dp(flat.b),
            len(flat.cone_type), self._keep[1].ctypes.data_as(ip), self._keep[2].ctypes.data_as(ip), _dp(ls),
            flat.x_min, flat.x_max, flat.u_min, flat.u_max, algorithm, float(L_norm), None, n_rms,
        )
        if getattr(flat, "w", None) is not None:
            wv = np.ascontiguousarray(flat.w, dtype=np.int32)
            self._keep.append(wv)
            p.w = wv.ctypes.data_as(ip)
        self.h = self.lib.spo_create(C.byref(p), int(n_threads))
        if not self.h:
            raise ValueError("spo_create failed")
        self.nz, self.nv, self.n = self.lib.spo_nz(self.h), self.lib.spo_nv(self.h), self.lib.spo_n(self.h)
        self.L_norm = L_norm if L_norm > 0 else (3.28 if algorithm == SP else 10.0)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.spo_destroy(self.h)
            self.h = None

    # -- state views (live memory of the C model)
    @property
    def z(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.spo_z(self.h), shape=(self.nz,))

    @property
    def v(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.spo_v(self.h), shape=(self.nv,))

    def default_step(self) -> float:
        return 0.99 / np.sqrt(self.L_norm)

    def tree_maps(self):
        nnl = (self.flat.d ** (self.flat.N - 1) - 1) // (self.flat.d - 1)
        cf = np.zeros(max(nnl, 1), np.int64)
        anc = np.zeros(self.n, np.int64)
        w = np.zeros(self.n, np.int32)
        mi = np.zeros(self.flat.N, np.int64)
        lp, ip = C.POINTER(C.c_long), C.POINTER(C.c_int)
        self.lib.spo_tree_maps(self.h, cf.ctypes.data_as(lp), anc.ctypes.data_as(lp), w.ctypes.data_as(ip),
                               mi.ctypes.data_as(lp))
        return cf[:nnl], anc, w, mi

    def offsets(self):
        zo, vo = np.zeros(5, np.int64), np.zeros(11, np.int64)
        lp = C.POINTER(C.c_long)
        self.lib.spo_offsets(self.h, zo.ctypes.data_as(lp), vo.ctypes.data_as(lp))
        return zo, vo

    def set_x0(self, x0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.lib.spo_set_x0(self.h, _dp(x0))

    # -- operators (return new arrays; the C side works in place like the reference)
    def L(self, z):
        z = np.ascontiguousarray(z, np.float64)
        v = np.zeros(self.nv)
        self.lib.spo_L(self.h, _dp(z), _dp(v))
        return v

    def Lt(self, v):
        v = np.ascontiguousarray(v, np.float64)
        z = np.zeros(self.nz)
        self.lib.spo_Lt(self.h, _dp(z), _dp(v))
        return z

    def dot(self, a_z, a_v, b_z, b_v, gamma, sigma):
        arrs = [np.ascontiguousarray(a, np.float64) for a in (a_z, a_v, b_z, b_v)]
        return self.lib.spo_dot(self.h, *[_dp(a) for a in arrs], gamma, sigma)

    def projection_S1(self, xu):
        out = np.array(xu, dtype=np.float64, copy=True)
        self.lib.spo_projection_S1(self.h, _dp(out))
        return out

    def projection_S2(self, z2):
        out = np.array(z2, dtype=np.float64, copy=True)
        self.lib.spo_projection_S2(self.h, _dp(out))
        return out

    def prox_f(self, arg, gamma):
        out = np.array(arg, dtype=np.float64, copy=True)
        self.lib.spo_prox_f(self.h, _dp(out), gamma)
        return out

    def prox_h_conj(self, arg, sigma):
        out = np.array(arg, dtype=np.float64, copy=True)
        self.lib.spo_prox_h_conj(self.h, _dp(out), sigma)
        return out

    def project_soc(self, vec):
        out = np.array(vec, dtype=np.float64, copy=True)
        self.lib.spo_project_soc(_dp(out), len(out))
        return out

    def factors(self, i):
        nx, nu = self.flat.nx, self.flat.nu
        P, ABK, K, Rc = np.zeros(nx * nx), np.zeros(nx * nx), np.zeros(nu * nx), np.zeros(nu * nu)
        self.lib.spo_factors(self.h, i, _dp(P), _dp(ABK), _dp(K), _dp(Rc))
        return (P.reshape(nx, nx).T, ABK.reshape(nx, nx).T, K.reshape(nx, nu).T, Rc.reshape(nu, nu).T)

    def solve(self, x0, tol=1e-3, verbose_log=False, z0=None, v0=None, gamma=0.0, sigma=0.0, max_iter=0):
        """solve_model! -- returns (result dict, log dict)."""
        x0 = np.ascontiguousarray(x0, np.float64)
        cap = max_iter if max_iter > 0 else (1000 if self.algorithm == SP else 5000)
        xi, xi1, xi2 = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
        bc = np.zeros(cap, np.int32)
        res = _Result()
        z0 = None if z0 is None else np.ascontiguousarray(z0, np.float64)
        v0 = None if v0 is None else np.ascontiguousarray(v0, np.float64)
        rc = self.lib.spo_solve(self.h, _dp(x0), tol, int(verbose_log), _dp(z0), _dp(v0), gamma, sigma, max_iter,
                                C.byref(res), _dp(xi), _dp(xi1), _dp(xi2), bc.ctypes.data_as(C.POINTER(C.c_int)), cap)
        if rc != 0:
            raise ValueError("z0 and v0 must be given together (cpock.jl:180-183)")
        k = res.passes
        out = {f: getattr(res, f) for f, _ in _Result._fields_}
        return out, {"xi": xi[:k], "xi1": xi1[:k], "xi2": xi2[:k], "backtrack_count": bc[:k]}

    def set_sp_options(self, linesearch: bool = False, k0: bool = False):
        """SuperMann variants present but not executed in the reference (sp.jl:439, sp.jl:80); default = as shipped."""
        self.lib.spo_set_sp_options(self.h, (1 if linesearch else 0) | (2 if k0 else 0))

    def step(self, K, gamma=0.0, sigma=0.0):
        self.lib.spo_step(self.h, int(K), gamma, sigma)
