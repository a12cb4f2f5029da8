"""ctypes binding of the C ABI in include/spock_b200.h.

`load()` returns the CUDA library (built in-tree by spock_b200.build) and raises if it is missing --
there is no CPU fallback.  `bind(path)` binds any shared object exporting the same ABI (the test-suite
uses it for the host-emulation build of the kernels under tests/hostemu).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# (SPOCK_B200_LIB: development builds of the same library under csrc/_variants/, used for A/B timing of kernel variants)
LIB_PATH = os.environ.get("SPOCK_B200_LIB") or os.path.join(HERE, "csrc", "libspock_b200.so")

SPK_OK = 0
CP, SP = 1, 2
FACTORS_PER_NODE, FACTORS_PER_STAGE = 0, 1
OP_L, OP_LT, OP_PROX_F, OP_PROX_HCONJ, OP_PROJ_S1, OP_PROJ_S2 = range(6)

EXPORTS = [
    "spk_last_error", "spk_version", "spk_create", "spk_destroy", "spk_sizes", "spk_offsets", "spk_tree_maps",
    "spk_set_x0", "spk_set_iterate", "spk_get_iterate", "spk_apply", "spk_dot", "spk_get_factors", "spk_step",
    "spk_solve", "spk_get_residuals", "spk_profile", "spk_profile_get", "spk_launch_count", "spk_device_bytes",
    "spk_comm_unique_id", "spk_owned_ranges", "spk_state_size", "spk_get_state", "spk_set_state",
    "spk_estimate_L_norm", "spk_set_L_norm", "spk_gather_iterate", "spk_set_sp_options", "spk_shift_iterate", "spk_solve_batch", "spk_mul",
]

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


class SpkProblem(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("d", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
        ("A", _dp), ("B", _dp),
        ("cost_dense", C.c_int32), ("n_sqrtQ", C.c_int32), ("sqrtQ", _dp), ("sqrtR", _dp),
        ("n_sqrtQN", C.c_int32), ("sqrtQN", _dp),
        ("ny", C.c_int32), ("E", _dp), ("F", _dp), ("b", _dp),
        ("n_cones", C.c_int32), ("cone_type", _i32p), ("cone_dim", _i32p),
        ("x_min", C.c_double), ("x_max", C.c_double), ("u_min", C.c_double), ("u_max", C.c_double),
        ("algorithm", C.c_int32), ("L_norm", C.c_double), ("factor_mode", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("comm_id", C.POINTER(C.c_uint8)),
        ("n_devices", C.c_int32), ("devices", _i32p),
        ("n_rms", C.c_int32),
        ("mode_of_node", _i32p),
    ]


class SpkResult(C.Structure):
    _fields_ = [
        ("passes", C.c_int32), ("printed_iter", C.c_int32), ("converged", C.c_int32),
        ("xi1", C.c_double), ("xi2", C.c_double), ("xi", C.c_double),
        ("k1_count", C.c_int32), ("k2_count", C.c_int32), ("fallback_count", C.c_int32),
        ("solve_ms", C.c_double),
    ]


class SpockLibraryError(RuntimeError):
    pass


def bind(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.spk_last_error.restype = C.c_char_p
    lib.spk_last_error.argtypes = []
    lib.spk_version.restype = C.c_int
    lib.spk_create.argtypes = [C.POINTER(SpkProblem), C.POINTER(vp)]
    lib.spk_destroy.argtypes = [vp]
    lib.spk_sizes.argtypes = [vp, _i64p]
    lib.spk_offsets.argtypes = [vp, _i64p, _i64p]
    lib.spk_tree_maps.argtypes = [C.c_int32, C.c_int32, _i64p, _i64p, _i32p, _i64p]
    lib.spk_set_x0.argtypes = [vp, _dp]
    lib.spk_set_iterate.argtypes = [vp, _dp, _dp]
    lib.spk_get_iterate.argtypes = [vp, _dp, _dp]
    lib.spk_apply.argtypes = [vp, C.c_int32, _dp, _dp, C.c_double]
    lib.spk_mul.argtypes = [vp, C.c_int32, _dp, C.c_double, C.c_double, _dp]
    lib.spk_dot.argtypes = [vp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.spk_get_factors.argtypes = [vp, C.c_int64, _dp, _dp, _dp, _dp, _dp]
    lib.spk_step.argtypes = [vp, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_float)]
    lib.spk_solve.argtypes = [vp, _dp, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32, _dp, _dp,
                              C.POINTER(SpkResult), _dp, _dp, _dp, _i32p, C.c_int32]
    lib.spk_get_residuals.argtypes = [vp, _dp]
    lib.spk_profile.argtypes = [vp, C.c_int32, C.c_int32]
    lib.spk_profile_get.argtypes = [vp, C.c_int32, _dp, _i64p]
    lib.spk_launch_count.argtypes = [vp, _i64p]
    lib.spk_device_bytes.argtypes = [vp, _i64p]
    lib.spk_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.spk_owned_ranges.argtypes = [vp, _i32p, _i64p, _i64p]
    lib.spk_state_size.argtypes = [vp, _i64p]
    lib.spk_get_state.argtypes = [vp, _dp]
    lib.spk_set_state.argtypes = [vp, _dp]
    lib.spk_estimate_L_norm.argtypes = [vp, C.c_int32, _dp]
    lib.spk_set_L_norm.argtypes = [vp, C.c_double]
    lib.spk_gather_iterate.argtypes = [vp]
    lib.spk_set_sp_options.argtypes = [vp, C.c_int32]
    lib.spk_shift_iterate.argtypes = [vp, C.c_int32]
    lib.spk_solve_batch.argtypes = [C.POINTER(vp), C.c_int32, _dp, C.c_double, C.c_int32, C.POINTER(SpkResult)]
    for name in EXPORTS:
        if name not in ("spk_last_error",):
            getattr(lib, name).restype = C.c_int
    lib.spk_last_error.restype = C.c_char_p
    return lib


_LIB = None


def load() -> C.CDLL:
    """The CUDA library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SpockLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m spock_b200.build` (nvcc, sm_100a). "
                "spock_b200 has no CPU fallback."
            )
        _LIB = bind(LIB_PATH)
    return _LIB


def check(lib: C.CDLL, rc: int) -> None:
    if rc != SPK_OK:
        msg = lib.spk_last_error()
        raise SpockLibraryError(f"libspock_b200 error {rc}: {msg.decode() if msg else '?'}")


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def make_problem(flat, algorithm: int, L_norm: float = 0.0, factor_mode: int = FACTORS_PER_NODE, device: int = -1,
                 rank: int = 0, world_size: int = 1, comm_id=None, devices=None):
    """FlatProblem -> (SpkProblem, keepalive list)."""
    ct = np.ascontiguousarray(flat.cone_type, dtype=np.int32)
    cd = np.ascontiguousarray(flat.cone_dim, dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
            (flat.A, flat.B, flat.sqrtQ, flat.sqrtR, flat.sqrtQN, flat.E, flat.F, flat.b)]
    A, B, sQ, sR, sQN, E, F, b = arrs
    cid = None
    keep = arrs + [ct, cd]
    if comm_id is not None:
        cid_arr = np.ascontiguousarray(comm_id, dtype=np.uint8)
        keep.append(cid_arr)
        cid = cid_arr.ctypes.data_as(C.POINTER(C.c_uint8))
    p = SpkProblem(
        flat.N, flat.d, flat.nx, flat.nu, dptr(A), dptr(B), flat.cost_dense, flat.n_sqrtQ, dptr(sQ), dptr(sR),
        flat.n_sqrtQN, dptr(sQN), flat.ny, dptr(E), dptr(F), dptr(b), len(ct), ct.ctypes.data_as(_i32p),
        cd.ctypes.data_as(_i32p), flat.x_min, flat.x_max, flat.u_min, flat.u_max, int(algorithm), float(L_norm),
        int(factor_mode), int(device), int(rank), int(world_size), cid, 0, None, int(getattr(flat, "n_rms", 1)), None,
    )
    if getattr(flat, "w", None) is not None:
        wv = np.ascontiguousarray(flat.w, dtype=np.int32)
        keep.append(wv)
        p.mode_of_node = wv.ctypes.data_as(_i32p)
    if devices is not None and len(devices) > 0:
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        keep.append(dv)
        p.n_devices, p.devices = len(dv), dv.ctypes.data_as(_i32p)
    return p, keep
