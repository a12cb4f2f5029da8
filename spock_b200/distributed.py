"""One-process-per-GPU plumbing over torch.distributed (NCCL / gloo): bootstrap of the library's communicator and the
1-GPU-vs-G-GPU equality check of the subtree partition.

The data path itself never goes through torch: the two cut-level all-gathers and the scalar all-reduces of an operator
evaluation are issued by libspock_b200.so on its own stream with its own NCCL communicator.  torch.distributed only
carries the 128-byte ncclUniqueId to the ranks and, here, the verdict of the check."""
from __future__ import annotations

import ctypes as C
import zlib

import numpy as np

from . import _lib
from .handle import Handle


def broadcast_comm_id(lib, rank: int, device=None) -> np.ndarray:
    """spk_comm_unique_id on rank 0, broadcast to every rank (one id per communicator, i.e. per handle)."""
    import torch
    import torch.distributed as dist

    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = (C.c_uint8 * 128)()
        _lib.check(lib, lib.spk_comm_unique_id(raw))
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return buf.cpu().numpy()


def relerr(a, b) -> float:
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def partition_parity(lib, flat, algorithm: int, K: int, factor_mode: int, rank: int, world: int, device: int = -1):
    """K passes on the tree partitioned over `world` ranks, spk_gather_iterate, then on rank 0 the same K passes on one
    GPU: relative error of the gathered (z, v) and of the residual pair, and whether every rank holds bitwise the same
    gathered iterate (CRC of the bytes).  Collective: every rank must call it; every rank returns the same dict."""
    import torch.distributed as dist

    x0 = 0.1 * np.ones(flat.nx)
    cid = broadcast_comm_id(lib, rank)
    h = Handle(flat, algorithm, factor_mode=factor_mode, device=device, rank=rank, world_size=world, comm_id=cid, lib=lib)
    h.set_x0(x0)
    h.step(K)
    xi = h.residuals()
    h.gather_iterate()
    z, v = h.get_iterate()
    dbytes = h.device_bytes()
    h.close()
    crc = zlib.crc32(v.tobytes(), zlib.crc32(z.tobytes()))
    crcs = [None] * world
    dist.all_gather_object(crcs, (crc, xi))
    out = [None]
    if rank == 0:
        ref = Handle(flat, algorithm, factor_mode=factor_mode, device=device, lib=lib)
        ref.set_x0(x0)
        ref.step(K)
        zr, vr = ref.get_iterate()
        xr = ref.residuals()
        ref.close()
        out[0] = {
            "nodes": int(flat.n), "passes": K, "algorithm": "CP" if algorithm == _lib.CP else "SP", "world": world,
            "factor_mode": "per_node" if factor_mode == 0 else "per_stage",
            "relerr_z": relerr(z, zr), "relerr_v": relerr(v, vr),
            "relerr_xi": max(abs(a[1][0] - xr[0]) / xr[0] for a in crcs) if xr[0] > 0 else 0.0,
            "gathered_iterate_identical_on_all_ranks": len({a[0] for a in crcs}) == 1,
            "device_bytes_per_rank": int(dbytes),
        }
    dist.broadcast_object_list(out, 0)
    return out[0]
