"""The reference's scaling experiment (examples/server_heat/scaling.jl): server_heat, d=2, nx=nu=50, cold start,
tol 1e-3, N = 3..Nmax, solved with SP (the reference's choice) and CP on one B200, next to the published Julia
SPOCK times (examples/server_heat/output/scaling.tikz:29-41, unknown CPU, 1 thread).
usage: python tools/scaling_sweep.py [Nmax=15] [nx=50] [d=2] > sweep.jsonl"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from spock_b200 import SP, CP, SolverOptions, build_model, get_server_heat_specs, solve_model_  # noqa: E402

REF_SP_S = {3: 0.0299, 4: 0.0609, 5: 0.1206, 6: 0.2853, 7: 0.5575, 8: 1.541, 9: 2.428, 10: 4.801, 11: 14.62, 12: 22.88,
            13: 45.66, 14: 78.27, 15: 151.5}
Nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 15
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x0 = 0.1 * np.ones(nx)
for N in range(3, Nmax + 1):
    spec = get_server_heat_specs(N, nx, d)
    row = {"N": N, "d": d, "nx": nx, "nodes": spec[0].n, "ref_julia_sp_s": REF_SP_S.get(N) if (nx, d) == (50, 2) else None}
    for alg, name in ((SP, "sp"), (CP, "cp")):
        for fm, fname in ((1, "per_stage"), (0, "per_node")):
            t0 = time.perf_counter()
            model = build_model(*spec, SolverOptions(alg, None, factor_mode=fm))
            t_build = time.perf_counter() - t0
            t0 = time.perf_counter()
            solve_model_(model, x0, tol=1e-3, quiet=True)
            wall = time.perf_counter() - t0
            r = model.last_result
            row[f"{name}_{fname}"] = {"solve_s": r["solve_ms"] / 1e3, "wall_s": wall, "build_s": t_build, "passes": r["passes"],
                                      "converged": bool(r["converged"]), "xi1": r["xi1"], "xi2": r["xi2"]}
            model.handle.close()
    print(json.dumps(row), flush=True)
