"""The subtree partition on CPU: G ranks of the host-emulation build exchange the cut-level
entries through the emulated communicator and must reproduce the single-rank iterate.  Ranks are threads of one
process in the first test and world_size-2 `gloo` processes (torch.distributed) in the second."""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest

from oracle.oracle import CP, SP
from spock_b200 import _lib
from spock_b200.handle import Handle
from tests.helpers import flat_of, relerr
from tests.multirank_helpers import assemble


def _comm_id(lib):
    raw = (C.c_uint8 * 128)()
    _lib.check(lib, lib.spk_comm_unique_id(raw))
    return np.frombuffer(bytes(raw), dtype=np.uint8).copy()


def _run_rank(lib, flat, fm, rank, world, cid, K, out, alg=CP):
    try:
        h = Handle(flat, alg, factor_mode=fm, rank=rank, world_size=world, comm_id=cid, lib=lib)
        h.set_x0(0.1 * np.ones(flat.nx))
        h.step(K)
        z, v = h.get_iterate()
        xi = h.residuals()
        if alg == CP:  # spk_gather_iterate: the complete iterate on every rank; the passes that follow are unaffected
            h.gather_iterate()
            zfull, vfull = h.get_iterate()
            h.step(3)
            h.gather_iterate()
            out[("full", rank)] = (zfull, vfull) + h.get_iterate()
        extra = None
        if alg == SP:  # then solve_model! to tolerance from this point: every rank must take the same decisions
            res, log = h.solve(0.1 * np.ones(flat.nx), tol=1e-4)
            extra = (res["passes"], res["converged"], res["k1_count"], res["k2_count"], res["fallback_count"], log["xi1"].copy())
        out[rank] = (h.owned_ranges(), z, v, xi, h.z_off, h.v_off, extra)
        h.close()
    except Exception as e:  # surface the failure in the main thread
        out[rank] = e


@pytest.mark.parametrize("name,world", [("heat_N7_d2_nx5", 2), ("heat_N7_d2_nx5", 4), ("heat_N4_d3_nx3", 3),
                                        ("random_N5_d2_nx6_nu3", 8), ("l2ball_N4_d3_nx3", 3)])
def test_partition_matches_single_rank_threads(emu_lib, name, world):
    flat = flat_of(name)
    K = 25
    ref = Handle(flat, CP, lib=emu_lib)
    ref.set_x0(0.1 * np.ones(flat.nx))
    ref.step(K)
    zr, vr = ref.get_iterate()
    xr = ref.residuals()
    cid = _comm_id(emu_lib)
    out = {}
    threads = [threading.Thread(target=_run_rank, args=(emu_lib, flat, g % 2, g, world, cid, K, out)) for g in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    for g in range(world):
        assert g in out and not isinstance(out[g], Exception), out.get(g)
    z, v, worst = assemble(flat, out[0][4], out[0][5], [(out[g][0], out[g][1], out[g][2]) for g in range(world)])
    assert worst < 1e-12                      # replicated levels agree across ranks
    assert relerr(z, zr) < 1e-12 and relerr(v, vr) < 1e-12
    for g in range(world):                   # xi1, xi2 are global maxima on every rank
        assert abs(out[g][3][0] - xr[0]) < 1e-12 * xr[0] and abs(out[g][3][1] - xr[1]) < 1e-12 * xr[1]
    # spk_gather_iterate: every rank holds the complete single-rank iterate, also after further passes
    ref.step(3)
    zr3, vr3 = ref.get_iterate()
    for g in range(world):
        zf, vf, zf3, vf3 = out[("full", g)]
        assert relerr(zf, zr) < 1e-12 and relerr(vf, vr) < 1e-12 and relerr(zf3, zr3) < 1e-12 and relerr(vf3, vr3) < 1e-12
    # ownership: every non-leaf node of the levels below the cut belongs to exactly one rank
    owned = np.zeros(flat.n_non_leaf, int)
    for g in range(world):
        for first, count in out[g][0]:
            owned[first - 1: first - 1 + count] += 1
    assert owned.min() >= 1 and (owned[owned > 1] == world).all()


@pytest.mark.parametrize("name,world", [("heat_N7_d2_nx5", 2), ("heat_N7_d2_nx5", 4), ("l2ball_N4_d3_nx3", 3),
                                        ("random_N5_d2_nx6_nu3", 4)])
def test_supermann_over_the_partition_threads(emu_lib, name, world):
    """SuperMann over the subtree partition: the global sums (P-norms, Anderson's Gram-Schmidt dots, the K2 inner product)
    and maxima are reduced over each rank's counted entries and all-reduced; the assembled iterate after K passes equals
    the single-rank one up to the rounding of the re-associated sums."""
    flat = flat_of(name)
    K = 12
    ref = Handle(flat, SP, lib=emu_lib)
    ref.set_x0(0.1 * np.ones(flat.nx))
    ref.step(K)
    zr, vr = ref.get_iterate()
    xr = ref.residuals()
    cid = _comm_id(emu_lib)
    out = {}
    threads = [threading.Thread(target=_run_rank, args=(emu_lib, flat, g % 2, g, world, cid, K, out, SP)) for g in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(180)
    for g in range(world):
        assert g in out and not isinstance(out[g], Exception), out.get(g)
    z, v, worst = assemble(flat, out[0][4], out[0][5], [(out[g][0], out[g][1], out[g][2]) for g in range(world)])
    assert worst < 1e-9
    assert relerr(z, zr) < 1e-9 and relerr(v, vr) < 1e-9
    for g in range(world):
        assert abs(out[g][3][0] - xr[0]) < 1e-7 * xr[0] and abs(out[g][3][1] - xr[1]) < 1e-7 * xr[1]
    for g in range(1, world):  # identical stopping pass, branch counts and residual trace on every rank
        assert out[g][6][:5] == out[0][6][:5] and out[0][6][1] == 1
        assert np.array_equal(out[g][6][5], out[0][6][5])


def test_uneven_split_is_rejected(emu_lib):
    flat = flat_of("heat_N4_d3_nx3")  # d = 3 never splits over 2 ranks
    with pytest.raises(_lib.SpockLibraryError):
        Handle(flat, CP, rank=0, world_size=2, comm_id=_comm_id(emu_lib), lib=emu_lib)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.hostemu.build_hostemu import build

    lib = _lib.bind(build())
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.from_numpy(_comm_id(lib)))
    dist.broadcast(buf, 0)
    flat = flat_of("heat_N7_d2_nx5")
    h = Handle(flat, CP, rank=rank, world_size=world, comm_id=buf.numpy(), lib=lib)
    res, _ = h.solve(0.1 * np.ones(5), tol=1e-3)
    z, v = h.get_iterate()
    q.put((rank, h.owned_ranges(), z, v, res["passes"], res["xi1"], h.z_off, h.v_off))
    h.close()
    dist.barrier()
    dist.destroy_process_group()


def test_partition_world2_gloo_processes(emu_lib):
    import torch.multiprocessing as mp

    flat = flat_of("heat_N7_d2_nx5")
    ref = Handle(flat, CP, lib=emu_lib)
    r0, _ = ref.solve(0.1 * np.ones(5), tol=1e-3)
    zr, vr = ref.get_iterate()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(g, 2, port, q)) for g in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
    got.sort(key=lambda t: t[0])
    assert got[0][4] == got[1][4] == r0["passes"]          # same stopping pass on every rank as on one rank
    assert abs(got[0][5] - r0["xi1"]) < 1e-12 * r0["xi1"]
    z, v, worst = assemble(flat, got[0][6], got[0][7], [(g[1], g[2], g[3]) for g in got])
    assert worst < 1e-12 and relerr(z, zr) < 1e-12 and relerr(v, vr) < 1e-12


@pytest.mark.parametrize("name,G,alg", [("heat_N7_d2_nx5", 2, CP), ("heat_N7_d2_nx5", 4, CP), ("heat_N4_d3_nx3", 3, CP),
                                        ("heat_N7_d2_nx5", 2, SP), ("l2ball_N4_d3_nx3", 3, SP)])
def test_single_process_device_group(emu_lib, name, G, alg):
    """spk_problem.n_devices / devices: ONE handle in ONE process drives G ranks of the subtree partition (one worker
    thread per device; the Julia host of the reference is a single process, SURVEY 8(b) "Threading").  Same iterate,
    residuals, stopping pass and branch counts as a single-device handle; spk_get_iterate assembles the complete (z, v)."""
    flat = flat_of(name)
    x0 = 0.1 * np.ones(flat.nx)
    K = 20 if alg == CP else 8
    ref = Handle(flat, alg, lib=emu_lib)
    grp = Handle(flat, alg, lib=emu_lib, devices=list(range(G)))
    assert (grp.nz, grp.nv, grp.n) == (ref.nz, ref.nv, ref.n) and grp.z_off == ref.z_off
    for h in (ref, grp):
        h.set_x0(x0)
        h.step(K)
    (zr, vr), (zg, vg) = ref.get_iterate(), grp.get_iterate()
    assert relerr(zg, zr) < 1e-11 and relerr(vg, vr) < 1e-11
    xr, xg = ref.residuals(), grp.residuals()
    assert abs(xr[0] - xg[0]) <= 1e-10 * xr[0] and abs(xr[1] - xg[1]) <= 1e-10 * xr[1]
    rr, _ = ref.solve(x0, tol=1e-3)
    rg, lg = grp.solve(x0, tol=1e-3)
    assert rg["converged"] == 1 and len(lg["xi1"]) == rg["passes"]
    if alg == CP:
        assert rg["passes"] == rr["passes"]
    (zr, vr), (zg, vg) = ref.get_iterate(), grp.get_iterate()
    if alg == CP:  # (SuperMann to 1e-3: different reduction orders end at different points of the tolerance ball)
        assert relerr(zg, zr) < 1e-10 and relerr(vg, vr) < 1e-10
    # inner operator interface and the P-inner product through the group handle
    rng = np.random.default_rng(4)
    z, v = rng.standard_normal(ref.nz), rng.standard_normal(ref.nv)
    assert relerr(grp.L(z), ref.L(z)) < 1e-13 and relerr(grp.prox_h_conj(v, 0.3), ref.prox_h_conj(v, 0.3)) < 1e-13
    assert abs(grp.dot(z, v, z, v, 0.2, 0.2) - ref.dot(z, v, z, v, 0.2, 0.2)) < 1e-10 * abs(ref.dot(z, v, z, v, 0.2, 0.2))
    assert grp.device_bytes() > ref.device_bytes() and grp.owned_ranges() == [(1, flat.n_non_leaf)]
    # argument errors surface from the worker threads
    with pytest.raises(Exception, match="splits evenly|exclusive|devices"):
        Handle(flat_of("heat_N4_d3_nx3"), CP, lib=emu_lib, devices=[0, 1])
    grp.close()
    ref.close()
