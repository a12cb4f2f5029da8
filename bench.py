#!/usr/bin/env python3
"""bench.py -- CP iterations/s of the Chambolle-Pock hot path over a scenario tree (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one CP pass (cp.jl:205-228 including the termination residuals) over the whole tree.
Default workload = BASELINE.json configs[2] / SURVEY.md section 8(d) config 3: random linear system, N=20, d=2,
nx=50, nu=20 (1 048 575 nodes), AV@R + box constraints, per-node Riccati factors (the graded, HBM-bound mode).
For N > 1 the tree is partitioned by subtrees (one process per GPU, launched by torchrun) -- strong scaling.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator, N, d, nx, nu, description)
    "config3": ("random", 20, 2, 50, 20, "SURVEY 8(d) config 3: random linear system N=20 d=2 nx=50 nu=20, 1048575 nodes, AV@R+box"),
    "config3_half": ("random", 19, 2, 50, 20, "config 3 with N=19 (524287 nodes)"),
    "heat15": ("heat", 15, 2, 50, 50, "reference scaling.jl largest case: server_heat N=15 d=2 nx=nu=50, 32767 nodes"),
    "config5_slice": ("heat_l2", 11, 4, 10, 10, "config 5 weak-scaling slice: server_heat N=11 d=4 nx=nu=10, SOC risk, 1398101 nodes"),
    "config5": ("heat_l2", 13, 4, 10, 10, "SURVEY 8(d) config 5: server_heat N=13 d=4 nx=nu=10, l2-ball (SOC) risk, 22369621 nodes"),
    "config5_n12": ("heat_l2", 12, 4, 10, 10, "config 5 strong-scaling slice: N=12, 5592405 nodes"),
    "small": ("random", 12, 2, 50, 20, "development size: random N=12 d=2 nx=50 nu=20, 4095 nodes"),
}
CPU_SAMPLE_N = {"config3": 14, "config3_half": 14, "heat15": 12, "config5_slice": 8, "config5": 8, "config5_n12": 8, "small": 10}
# DRAM traffic of the dominant kernel from `ncu --set full` (profiles/r01f_bwd_config3_ncu.txt): deepest-level launch
# (262144 of the 524287 families) read 16.749 GB + wrote 0.152 GB; scaled to all level launches of one pass
NCU_TRAFFIC = {("config3", "per_node"): (16.749308e9 + 0.152153344e9) * 524287 / 262144}


DMMA_PEAK_TFLOPS = 37.1  # mma.sync.m8n8k4.f64, measured (tools/microbench/dmma_bench.cu)


def make_spec(name, N_override=None):
    from spock_b200.examples import get_random_system_specs, get_server_heat_specs

    gen, N, d, nx, nu, _ = WORKLOADS[name]
    N = N_override or N
    if gen == "random":
        return get_random_system_specs(N, d, nx, nu)
    if gen == "heat":
        return get_server_heat_specs(N, nx, d)
    return get_server_heat_specs(N, nx, d, risk="l2ball")


def algorithmic_bytes(flat, factor_mode):
    """B_vec = 8[3(nz+nv) + 2(n nx + nnl nu)], B_fac per-node / per-stage."""
    n, nnl, nl, nx, nu, N, d = flat.n, flat.n_non_leaf, flat.n_leaf, flat.nx, flat.nu, flat.N, flat.d
    b_vec = 8 * (3 * (flat.nz + flat.nv) + 2 * (n * nx + nnl * nu))
    if factor_mode == 0:
        b_fac = 8 * (2 * nx * nx * (n - 1) + (3 * nu * nx + nu * nu) * nnl)
    else:
        b_fac = 8 * (2 * nx * nx * (N - 1) * d + (3 * nu * nx + nu * nu) * (N - 1))
    # the dominant kernel (backward S1 sweep, all its level launches of one pass)
    fac_bwd = 8 * ((nx * nx * (n - 1) + (2 * nu * nx + nu * nu) * nnl) if factor_mode == 0
                   else (nx * nx * (N - 1) * d + (2 * nu * nx + nu * nu) * (N - 1)))
    vec_bwd = 8 * (nnl * (2 * nx + 2 * nu) + (n - 1) * (nx + nu) + nl * 3 * nx + max(nnl - 1, 0) * nx + nnl * (nx + nu))
    return b_vec, b_fac, fac_bwd + vec_bwd


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (nvidia-ml-py)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_baseline(name, n_threads, iters=None):
    """The oracle (CPU port of the reference) on a bounded sample of the workload: the same system with a shallower
    tree, extrapolated linearly in the node count (the per-node work of every reference loop is constant)."""
    from oracle.oracle import CP, Oracle
    from spock_b200.problem import flatten_problem

    gen, N_full, d, nx, nu, _ = WORKLOADS[name]
    Ns = min(CPU_SAMPLE_N[name], N_full)
    flat_s = flatten_problem(*make_spec(name, Ns))
    full_nodes = (d**N_full - 1) // (d - 1)
    out = {}
    for label, nt in (("1thread", 1), ("threads", n_threads)):
        if label == "threads" and n_threads <= 1:
            continue
        o = Oracle(flat_s, CP, n_threads=nt)
        o.set_x0(0.1 * np.ones(nx))
        o.step(2)
        k = iters or 5
        t0 = time.perf_counter()
        o.step(k)
        dt = (time.perf_counter() - t0) / k
        if dt * k < 2.0 and iters is None:  # aim for >= 2 s of work
            k = int(min(200, max(5, 2.0 / dt)))
            t0 = time.perf_counter()
            o.step(k)
            dt = (time.perf_counter() - t0) / k
        out[label] = {"s_per_iter_sample": dt, "iters": k}
        del o
    scale = full_nodes / flat_s.n
    best = "threads" if "threads" in out else "1thread"
    return {
        "value": 1.0 / (out[best]["s_per_iter_sample"] * scale),
        "unit": "iterations/s",
        "cores": n_threads if best == "threads" else 1,
        "kind": "port",
        "sample": f"same system, N={Ns} ({flat_s.n} nodes), {out[best]['iters']} CP passes, scaled x{scale:.1f} by node count to N={N_full}",
        "value_1thread": 1.0 / (out["1thread"]["s_per_iter_sample"] * scale),
        "note": "C port of the single-threaded Julia reference (Julia is not installed here; Julia thread count = 1); "
                "'threads' = the same loops with OpenMP over the nodes of a stage",
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port, all host threads) on this arm's config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    gen, N_full, d, nx, nu, desc = WORKLOADS[args.workload]
    per_step_iters = 1
    # one "step" = one CP pass on the bounded sample; steps/warmup honoured on the sample
    from oracle.oracle import CP, Oracle
    from spock_b200.problem import flatten_problem

    Ns = min(CPU_SAMPLE_N[args.workload], N_full)
    flat_s = flatten_problem(*make_spec(args.workload, Ns))
    o = Oracle(flat_s, CP, n_threads=nthreads)
    o.set_x0(0.1 * np.ones(nx))
    o.step(max(args.warmup, 1))
    t0 = time.perf_counter()
    o.step(args.steps * per_step_iters)
    dt = (time.perf_counter() - t0) / args.steps
    scale = ((d**N_full - 1) // (d - 1)) / flat_s.n
    value = 1.0 / (dt * scale)
    line = {
        "impl": "reference", "metric": "cp_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt * scale,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "factor_mode": "per_node"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": nthreads, "kind": "port",
                         "sample": f"same system, N={Ns} ({flat_s.n} nodes), scaled x{scale:.1f} by node count to N={N_full}"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="config3", choices=list(WORKLOADS))
    ap.add_argument("--factor-mode", default="per_node", choices=["per_node", "per_stage"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline, per-stage line and time-to-tolerance")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from spock_b200 import _lib, build
    from spock_b200.handle import Handle
    from spock_b200.problem import flatten_problem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build.build()
    lib = _lib.load()

    gen, N, d, nx, nu, desc = WORKLOADS[args.workload]
    flat = flatten_problem(*make_spec(args.workload))
    fm = 0 if args.factor_mode == "per_node" else 1
    comm_id = None
    if world > 1:
        import ctypes as C

        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_uint8 * 128)()
            _lib.check(lib, lib.spk_comm_unique_id(raw))
            buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        comm_id = buf.cpu().numpy()
    h = Handle(flat, 1, factor_mode=fm, device=local_rank, rank=rank, world_size=world, comm_id=comm_id)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    x0_pinned = torch.full((nx,), 0.1, dtype=torch.float64).pin_memory()
    x0 = x0_pinned.numpy()
    h.set_x0(x0)
    h.step(args.warmup)  # W untimed warm-up passes (also pages the factors in)

    # ---- timed region: exactly K passes, device time from CUDA events on the solver stream, max over ranks
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = h.launch_count()
    barrier()
    ms = h.step(args.steps)
    barrier()
    launches = h.launch_count() - l0
    clocks = sampler.result()
    # the same K passes once more with one CUDA-event pair per kernel launch on the solver stream: per-kernel times for
    # the roofline (kept out of the timed region above so that `value` carries no instrumentation overhead)
    h.profile(True)
    barrier()
    ms_prof = h.step(args.steps)
    barrier()
    prof = h.profile_report()
    h.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- e2e through the C ABI with host buffers: every step copies x0 in (pinned) and the residual pair out,
    #      and the run ends with one D2H of the iterate (amortised over the K steps)
    z_host = torch.empty(h.nz, dtype=torch.float64).pin_memory().numpy()
    v_host = torch.empty(h.nv, dtype=torch.float64).pin_memory().numpy()
    import ctypes as C

    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.set_x0(x0)
        h.step(1)
        h.residuals()
    _lib.check(lib, lib.spk_get_iterate(h._h, _lib.dptr(z_host), _lib.dptr(v_host)))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 8 * nx,
           "d2h_bytes_per_step": 16 + 8 * (h.nz + h.nv) / args.steps,
           "note": "per step: spk_set_x0 (pinned H2D) + spk_step(1) + spk_get_residuals; one spk_get_iterate (D2H of z,v) "
                   "per run, amortised"}

    if rank == 0:
        b_vec, b_fac, b_bwd = algorithmic_bytes(flat, fm)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bwd = prof["fused_backward"]
        bwd_ms_per_pass = bwd["ms"] / args.steps if bwd["launches"] else None
        achieved = (b_bwd / world) / 1e9 / (bwd_ms_per_pass / 1e3) if bwd_ms_per_pass else None
        roofline = {
            "bound": "hbm", "kernel": "spk_k_bwd (backward S1 sweep: all N-1 level launches of one pass)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": (NCU_TRAFFIC.get((args.workload, args.factor_mode)) / world) if NCU_TRAFFIC.get((args.workload, args.factor_mode)) else None,
            "traffic_source": "ncu --set full, deepest-level launch scaled to one pass (profiles/r01f_bwd_config3_ncu.txt)"
            if NCU_TRAFFIC.get((args.workload, args.factor_mode)) else None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_pass": b_bwd / world, "avg_launch_ms": bwd["ms"] / max(bwd["launches"], 1),
            "share_of_step": (bwd["ms"] / ms_prof) if ms_prof else None,
            "measured_over": f"a repeat of the same {args.steps} passes with per-launch CUDA events ({ms_prof / args.steps:.3f} ms/pass "
                             f"instrumented vs {ms_per_step:.3f} ms/pass in the timed region)",
            "whole_pass": {"algorithmic_bytes": (b_vec + b_fac) / world, "achieved": (b_vec + b_fac) / world / 1e9 / (ms_per_step / 1e3),
                           "frac": (b_vec + b_fac) / world / 1e9 / (ms_per_step / 1e3) / peak},
            "per_kernel_ms_per_pass": {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]},
        }
        line = {
            "metric": "cp_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "nodes": flat.n, "nz": flat.nz, "nv": flat.nv,
                       "factor_mode": args.factor_mode, "l2_flush": "inputs larger than L2 (iterate + factors per GPU >> 126 MB)"
                       if (b_vec + b_fac) / world > 4e8 else "working set fits in L2: latency-bound configuration",
                       "partition": f"subtrees over {world} GPUs" if world > 1 else "single GPU",
                       "device_bytes": h.device_bytes()},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        }
        if not args.no_extras and world == 1:
            # time to xi <= 1e-3 from a cold start (solve_model! semantics)
            res, _ = h.solve(x0, tol=1e-3, max_iter=1500, z0=np.zeros(h.nz), v0=np.zeros(h.nv))
            line["time_to_tol"] = {"tol": 1e-3, "seconds": res["solve_ms"] / 1e3, "iterations": res["passes"],
                                   "converged": bool(res["converged"]), "xi1": res["xi1"], "xi2": res["xi2"]}
        h.close()
        if not args.no_extras and world == 1:
            # SuperMann (the reference's SP algorithm) to the same tolerance, cold start
            try:
                hs = Handle(flat, 2, factor_mode=fm, device=local_rank)
                rs, _ = hs.solve(x0, tol=1e-3, max_iter=1000)
                line["time_to_tol_sp"] = {"tol": 1e-3, "seconds": rs["solve_ms"] / 1e3, "iterations": rs["passes"],
                                          "converged": bool(rs["converged"]), "k1": rs["k1_count"], "k2": rs["k2_count"],
                                          "fallback": rs["fallback_count"], "operator_evaluations": 2 * rs["passes"]}
                hs.close()
            except Exception as e:  # e.g. out of device memory for the 17 extra iterate-sized vectors
                line["time_to_tol_sp"] = {"error": str(e)[:200]}
            # the same workload with de-duplicated (per-stage) factors, reported separately
            if fm == 0:
                h2 = Handle(flat, 1, factor_mode=1, device=local_rank)
                h2.set_x0(x0)
                h2.step(args.warmup)
                ms2 = h2.step(args.steps) / args.steps
                h2.profile(True)
                h2.step(args.steps)
                prof2 = {k: v["ms"] / args.steps for k, v in h2.profile_report().items() if v["launches"]}
                h2.profile(False)
                bv, bf, _ = algorithmic_bytes(flat, 1)
                # S1 sweeps on the fp64 tensor pipe (spk_mm.cuh): useful flops of SURVEY 8(d) against the measured DMMA peak
                fl_bwd = 2.0 * ((nx * nx + nx * nu) * (flat.n - 1) + (nu * nu + 2 * nx * nu) * flat.n_non_leaf)
                fl_fwd = 2.0 * ((nx * nx + nx * nu) * (flat.n - 1) + nx * nu * flat.n_non_leaf)
                r2, _ = h2.solve(x0, tol=1e-3, max_iter=1500, z0=np.zeros(h2.nz), v0=np.zeros(h2.nv))
                line["per_stage_factors"] = {
                    "value": 1000.0 / ms2, "unit": "iterations/s", "ms_per_step": ms2,
                    "algorithmic_bytes": bv + bf, "achieved_GBs": (bv + bf) / 1e9 / (ms2 / 1e3),
                    "frac_of_hbm_peak": (bv + bf) / 1e9 / (ms2 / 1e3) / peak,
                    "per_kernel_ms_per_pass": prof2,
                    "tensor_pipe": {"bound": "fp64 DMMA", "peak_TFLOPs": DMMA_PEAK_TFLOPS,
                                    "peak_source": "tools/microbench/dmma_bench.cu on this pool (profiles/r01d_dmma_microbench.txt)",
                                    "backward_TFLOPs": fl_bwd / 1e12 / (prof2["fused_backward"] / 1e3),
                                    "forward_TFLOPs": fl_fwd / 1e12 / (prof2["fused_forward"] / 1e3)},
                    "time_to_tol": {"tol": 1e-3, "seconds": r2["solve_ms"] / 1e3, "iterations": r2["passes"],
                                    "converged": bool(r2["converged"])},
                }
                h2.close()
            line["cpu_baseline"] = cpu_baseline(args.workload, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    else:
        h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
