"""BASELINE metric as a curve: CP iterations/s and time to xi <= 1e-3 versus tree nodes on one B200 (random linear system of
config 3, nx=50, nu=20, d=2, N = 8..20), per-node and per-stage factors, next to the CPU port of the reference (1 thread,
the reference's execution model) for the sizes it finishes in seconds.
usage: python tools/nodes_sweep.py [Nmin=8] [Nmax=20] [cpu_Nmax=13] > nodes_sweep.jsonl"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle.oracle import CP as OCP, Oracle  # noqa: E402  (CPU baseline leg only)
from spock_b200.examples import get_random_system_specs  # noqa: E402
from spock_b200.handle import Handle  # noqa: E402
from spock_b200.problem import flatten_problem  # noqa: E402

Nmin = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cpu_Nmax = int(sys.argv[3]) if len(sys.argv) > 3 else 13
nx, nu, d = 50, 20, 2
x0 = 0.1 * np.ones(nx)
for N in range(Nmin, Nmax + 1, 2):
    flat = flatten_problem(*get_random_system_specs(N, d, nx, nu))
    row = {"N": N, "nodes": flat.n, "nz": flat.nz, "nv": flat.nv}
    for fm, name in ((0, "per_node"), (1, "per_stage")):
        h = Handle(flat, 1, factor_mode=fm)
        h.set_x0(x0)
        h.step(5)
        K = 50 if N <= 16 else 20
        ms = h.step(K) / K
        res, _ = h.solve(x0, tol=1e-3, max_iter=2000, z0=np.zeros(h.nz), v0=np.zeros(h.nv))
        row[name] = {"it_per_s": 1000.0 / ms, "ms_per_pass": ms, "passes_to_tol": res["passes"], "s_to_tol": res["solve_ms"] / 1e3,
                     "converged": bool(res["converged"])}
        h.close()
    if N <= cpu_Nmax:
        o = Oracle(flat, OCP)
        o.set_x0(x0)
        o.step(2)
        K = max(2, int(2e5 / flat.n))
        t0 = time.perf_counter()
        o.step(K)
        dt = (time.perf_counter() - t0) / K
        row["cpu_port_1thread"] = {"it_per_s": 1.0 / dt, "s_to_tol_est": dt * row["per_node"]["passes_to_tol"]}
    print(json.dumps(row), flush=True)
