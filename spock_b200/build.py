"""Builds libspock_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m spock_b200.build [--force] [-v]

The solver (spk_solver.cu), the offline phase (spk_offline.cpp) and one translation unit per shape-specialised kernel
set (spk_shape.cu x the lines of spk_shapes.def) are compiled in parallel into csrc/_obj/ and linked.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "libspock_b200.so")
def _headers():
    """every header / table of csrc plus the public C header: any change rebuilds every translation unit"""
    hs = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".def")))
    return hs + [os.path.join("..", "..", "include", "spock_b200.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libspock_b200.so cannot be built (there is no CPU fallback)")


def shapes():
    """(nx, nu, d, threads_bwd, threads_fwd) per line of csrc/spk_shapes.def"""
    txt = open(os.path.join(CSRC, "spk_shapes.def")).read()
    out = [tuple(int(x) for x in m) for m in
           re.findall(r"^SPK_SHAPE_X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", txt, flags=re.M)]
    # SPK_SHAPE_THREADS="20,20,2:64,64;10,10,4:32,32": thread counts of development variants (tools/ab_variants.sh)
    for item in filter(None, os.environ.get("SPK_SHAPE_THREADS", "").split(";")):
        key, val = item.split(":")
        key, val = tuple(int(x) for x in key.split(",")), tuple(int(x) for x in val.split(","))
        out = [s[:3] + val if s[:3] == key else s for s in out]
    return out


def _units():
    """(object name, source, extra flags)"""
    units = [("spk_solver.o", "spk_solver.cu", []), ("spk_offline.o", "spk_offline.cpp", [])]
    for nx, nu, d, tb, tf in shapes():
        units.append((f"spk_shape_{nx}_{nu}_{d}.o", "spk_shape.cu",
                      [f"-DSPK_SHAPE_NX={nx}", f"-DSPK_SHAPE_NU={nu}", f"-DSPK_SHAPE_D={d}", f"-DSPK_SHAPE_TB={tb}",
                       f"-DSPK_SHAPE_TF={tf}"]))
    return units


def _dep_stamp(src: str, extra) -> str:
    h = hashlib.sha1()
    for d in [src] + _headers():
        p = os.path.join(CSRC, d)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    for obj, src, extra in _units():
        stamp = os.path.join(OBJ, obj + ".stamp")
        if not os.path.exists(stamp) or open(stamp).read() != _dep_stamp(src, extra):
            return True
    return False


def build(force: bool = False, verbose: bool = False, variant: str = "", extra_flags=()) -> str:
    """variant/extra_flags: development builds with extra nvcc flags into csrc/_variants/<variant>.so (tuning sweeps)."""
    global OBJ, LIB, NVCC_FLAGS
    if variant:
        default_obj = OBJ
        OBJ = os.path.join(CSRC, "_variants", variant + "_obj")
        LIB = os.path.join(CSRC, "_variants", variant + ".so")
        if not os.path.exists(OBJ) and os.path.exists(default_obj):
            shutil.copytree(default_obj, OBJ)  # units whose flags and sources are unchanged are not recompiled
        NVCC_FLAGS = NVCC_FLAGS + list(extra_flags)
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a CC wrapper that nvcc must not pick up
    logs = []

    def compile_one(unit):
        obj, src, extra = unit
        out, stamp_p = os.path.join(OBJ, obj), os.path.join(OBJ, obj + ".stamp")
        stamp = _dep_stamp(src, extra)
        if not force and os.path.exists(out) and os.path.exists(stamp_p) and open(stamp_p).read() == stamp:
            return
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + list(extra) + \
            ["-c", os.path.join(CSRC, src), "-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        logs.append(res.stderr)
        open(stamp_p, "w").write(stamp)

    units = _units()
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(units), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, units))
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + [os.path.join(OBJ, u[0]) for u in units]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    flags = sys.argv[sys.argv.index("--flags") + 1].split() if "--flags" in sys.argv else []
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var, extra_flags=flags))
