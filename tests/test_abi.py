"""The C-ABI library loads, exports every symbol include/spock_b200.h declares, rejects bad arguments with the
reference's error behaviour and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from spock_b200 import _lib, build
from tests.helpers import flat_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_exports_match_header(lib):
    header = open(os.path.join(ROOT, "include", "spock_b200.h")).read()
    declared = set(re.findall(r"\b(spk_[A-Za-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.spk_version() == 100


def test_struct_layout_matches_header():
    # field order of spk_problem / spk_result in the header == ctypes mirror
    header = open(os.path.join(ROOT, "include", "spock_b200.h")).read()
    body = header[header.index("typedef struct {", header.index("Inputs of build_model")):header.index("} spk_problem;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        names += [n.strip().lstrip("*") for n in re.sub(r"^(const\s+)?\w+\s+", "", decl).split(",")]
    assert names == [f for f, _ in _lib.SpkProblem._fields_]


def test_argument_errors(lib):
    flat = flat_of("fixture_N3_d2")
    h = C.c_void_p()
    p, keep = _lib.make_problem(flat, _lib.CP)
    p.d = 1  # scenario_tree.jl:62-64
    assert lib.spk_create(C.byref(p), C.byref(h)) == 1
    assert b"Branching factor d must be larger than 1, but is 1." in lib.spk_last_error()
    p, keep = _lib.make_problem(flat, 7)  # model.jl:295
    assert lib.spk_create(C.byref(p), C.byref(h)) == 1
    assert b"not supported" in lib.spk_last_error()
    p, keep = _lib.make_problem(flat, _lib.CP)
    p.cone_dim[0] = 3  # 3 + 1 != ny
    assert lib.spk_create(C.byref(p), C.byref(h)) == 1
    assert lib.spk_set_x0(None, None) == 1 and lib.spk_step(None, 1, 0.0, 0.0, None) == 1
    assert lib.spk_tree_maps(3, 1, None, None, None, None) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="this checks the no-GPU behaviour")
def test_no_cpu_fallback_without_device(lib):
    flat = flat_of("fixture_N3_d2")
    p, keep = _lib.make_problem(flat, _lib.CP)
    h = C.c_void_p()
    rc = lib.spk_create(C.byref(p), C.byref(h))
    assert rc == 2 and not h.value
    assert b"no CPU fallback" in lib.spk_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spock_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
                assert "hostemu" not in src or f in ("_lib.py", "spk_runtime.h"), f


def test_shape_table_and_thread_variants(monkeypatch):
    """spk_shapes.def: threads per family are whole warps; SPK_SHAPE_THREADS (development variants for same-box A/B
    timing, tools/ab_variants.sh) overrides exactly the named shape."""
    from spock_b200 import build
    monkeypatch.delenv("SPK_SHAPE_THREADS", raising=False)
    base = build.shapes()
    assert (50, 20, 2) in [s[:3] for s in base] and all(s[3] % 32 == 0 and s[4] % 32 == 0 for s in base)
    monkeypatch.setenv("SPK_SHAPE_THREADS", "20,20,2:64,96;10,10,4:32,32")
    var = build.shapes()
    assert [s for s in var if s[:3] == (20, 20, 2)] == [(20, 20, 2, 64, 96)]
    assert [s for s in var if s[:3] == (10, 10, 4)] == [(10, 10, 4, 32, 32)]
    assert [s for s in var if s[:3] not in ((20, 20, 2), (10, 10, 4))] == [s for s in base if s[:3] not in ((20, 20, 2), (10, 10, 4))]
