"""torchrun --nproc-per-node G tools/multigpu_check.py : the subtree partition over G GPUs (NCCL) reproduces the
single-GPU iterate.  Prints one 'MULTIGPU_OK ...' line on rank 0."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spock_b200 import _lib  # noqa: E402
from spock_b200.examples import get_random_system_specs, get_server_heat_specs  # noqa: E402
from spock_b200.handle import Handle  # noqa: E402
from spock_b200.problem import flatten_problem  # noqa: E402
from tests.multirank_helpers import assemble  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
worst_all = 0.0
for name, spec, K in (("heat N=8 nx=5", get_server_heat_specs(8, 5, 2), 40), ("random N=10 nx=20 nu=8", get_random_system_specs(10, 2, 20, 8), 20)):
    flat = flatten_problem(*spec)
    for fm in (0, 1):
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")  # one unique id per communicator
        if rank == 0:
            raw = (C.c_uint8 * 128)()
            _lib.check(lib, lib.spk_comm_unique_id(raw))
            buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        h = Handle(flat, 1, factor_mode=fm, device=local, rank=rank, world_size=world, comm_id=buf.cpu().numpy())
        h.set_x0(0.1 * np.ones(flat.nx))
        h.step(K)
        z, v = h.get_iterate()
        xi = h.residuals()
        views = [None] * world
        dist.all_gather_object(views, (h.owned_ranges(), z, v, xi))
        if rank == 0:
            ref = Handle(flat, 1, factor_mode=fm, device=local)
            ref.set_x0(0.1 * np.ones(flat.nx))
            ref.step(K)
            zr, vr = ref.get_iterate()
            za, va, worst = assemble(flat, h.z_off, h.v_off, [(w[0], w[1], w[2]) for w in views])
            ez = np.max(np.abs(za - zr)) / np.max(np.abs(zr))
            ev = np.max(np.abs(va - vr)) / np.max(np.abs(vr))
            xr = ref.residuals()
            assert worst < 1e-11 and ez < 1e-11 and ev < 1e-11, (name, fm, worst, ez, ev)
            for w in views:
                assert abs(w[3][0] - xr[0]) < 1e-10 * xr[0] and abs(w[3][1] - xr[1]) < 1e-10 * xr[1], (w[3], xr)
            worst_all = max(worst_all, ez, ev)
            ref.close()
        # solve_model! to tolerance with the device-resident termination test (xi all-reduced over the ranks every pass):
        # every rank stops on the same pass as the single-GPU solve
        x0 = 0.1 * np.ones(flat.nx)
        res, _ = h.solve(x0, tol=1e-3, z0=np.zeros(h.nz), v0=np.zeros(h.nv))
        allres = [None] * world
        dist.all_gather_object(allres, (res["passes"], res["converged"], res["xi1"], res["xi2"]))
        if rank == 0:
            ref = Handle(flat, 1, factor_mode=fm, device=local)
            rr, _ = ref.solve(x0, tol=1e-3, z0=np.zeros(ref.nz), v0=np.zeros(ref.nv))
            for a in allres:
                assert a[0] == rr["passes"] and a[1] == rr["converged"] == 1, (name, fm, a, rr["passes"])
                assert abs(a[2] - rr["xi1"]) < 1e-9 * rr["xi1"] and abs(a[3] - rr["xi2"]) < 1e-9 * rr["xi2"]
            ref.close()
        h.close()
        dist.barrier()
# SuperMann over the partition (host-driven loop; global sums / maxima all-reduced over each rank's counted entries)
sp_err = 0.0
for name, spec, K in (("heat N=8 nx=5", get_server_heat_specs(8, 5, 2), 12), ("random N=10 nx=20 nu=8", get_random_system_specs(10, 2, 20, 8), 10)):
    flat = flatten_problem(*spec)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_uint8 * 128)()
        _lib.check(lib, lib.spk_comm_unique_id(raw))
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    h = Handle(flat, 2, factor_mode=1, device=local, rank=rank, world_size=world, comm_id=buf.cpu().numpy())
    h.set_x0(0.1 * np.ones(flat.nx))
    h.step(K)
    z, v = h.get_iterate()
    res, _ = h.solve(0.1 * np.ones(flat.nx), tol=1e-3)
    views = [None] * world
    dist.all_gather_object(views, (h.owned_ranges(), z, v, res["passes"], res["converged"]))
    if rank == 0:
        ref = Handle(flat, 2, factor_mode=1, device=local)
        ref.set_x0(0.1 * np.ones(flat.nx))
        ref.step(K)
        zr, vr = ref.get_iterate()
        za, va, worst = assemble(flat, h.z_off, h.v_off, [(w[0], w[1], w[2]) for w in views])
        ez = np.max(np.abs(za - zr)) / np.max(np.abs(zr))
        ev = np.max(np.abs(va - vr)) / np.max(np.abs(vr))
        assert worst < 1e-9 and ez < 1e-9 and ev < 1e-9, (name, worst, ez, ev)
        assert all(w[3] == views[0][3] and w[4] == 1 for w in views), [(w[3], w[4]) for w in views]
        sp_err = max(sp_err, ez, ev)
        ref.close()
    h.close()
    dist.barrier()
# the headline shape on wide levels (2048-4096 families per level and rank: the persistent family kernels loop), over NCCL
# incl. spk_gather_iterate: the complete iterate on every rank equals the single-GPU one
from spock_b200.distributed import partition_parity  # noqa: E402

wide = []
flat = flatten_problem(*get_random_system_specs(14, 2, 50, 20))
for alg, K, fm in ((1, 10, 0), (1, 10, 1), (2, 5, 1), (2, 5, 0)):
    r = partition_parity(lib, flat, alg, K, fm, rank, world, local)
    assert r["relerr_z"] < 1e-9 and r["relerr_v"] < 1e-9 and r["gathered_iterate_identical_on_all_ranks"], r
    wide.append(r)
if rank == 0:
    import json
    print("MULTIGPU_WIDE " + json.dumps(wide), flush=True)
if rank == 0:
    print(f"MULTIGPU_OK world={world} max_rel_err={worst_all:.2e} supermann_max_rel_err={sp_err:.2e}", flush=True)
dist.destroy_process_group()
