"""Latency-bound configurations: microseconds per CP / SuperMann pass and kernel launches per
pass on small trees, device-resident loop (default) vs host-driven loop (SPK_HOST_LOOP=1).
usage: [SPK_LAT_FACTOR_MODE=0|1] python tools/small_tree_latency.py > latency.jsonl   (0: per-node factors, 1: per-stage tables)"""
import json
import os
import subprocess
import sys

sys.path.insert(0, ".")

CASES = [(3, 50, 2), (5, 50, 2), (7, 5, 2), (8, 50, 2), (10, 20, 2), (10, 50, 2), (12, 50, 2), (6, 20, 3)]

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np

    from spock_b200.examples import get_server_heat_specs
    from spock_b200.handle import Handle
    from spock_b200.problem import flatten_problem

    for N, nx, d in CASES:
        flat = flatten_problem(*get_server_heat_specs(N, nx, d))
        row = {"N": N, "nx": nx, "d": d, "nodes": flat.n, "host_loop": os.environ.get("SPK_HOST_LOOP", "0"),
               "factor_mode": int(os.environ.get("SPK_LAT_FACTOR_MODE", "1"))}
        for alg, name in ((1, "cp"), (2, "sp")):
            h = Handle(flat, alg, factor_mode=int(os.environ.get("SPK_LAT_FACTOR_MODE", "1")))
            x0 = 0.1 * np.ones(nx)
            h.solve(x0, tol=1e-3)  # warm-up: graph capture, instruction caches
            l0 = h.launch_count()
            r, _ = h.solve(x0, tol=1e-3, z0=np.zeros(h.nz), v0=np.zeros(h.nv))
            row[name] = {"passes": r["passes"], "solve_ms": r["solve_ms"], "us_per_pass": 1e3 * r["solve_ms"] / r["passes"],
                         "launches_per_pass": (h.launch_count() - l0) / r["passes"], "converged": bool(r["converged"])}
            h.close()
        print(json.dumps(row), flush=True)
else:
    for hl in os.environ.get("SPK_LAT_LOOPS", "0,1").split(","):
        env = dict(os.environ, SPK_HOST_LOOP=hl)
        subprocess.run([sys.executable, __file__, "--child"], env=env, check=True)
