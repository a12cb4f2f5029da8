"""The reference's closed-loop experiment (examples/server_heat/mpc_simulation.jl:11-54,160-183): server_heat N=10, d=2,
nx=nu=20 (1023 nodes), 20 MPC steps, SP, tol 1e-3, warm start by state persistence; disturbance w_t from
numpy.default_rng(1).integers(1, d+1) (the reference draws rand(1:d) after Random.seed!(1)).  Prints the per-step solve
latency next to the published trace (examples/server_heat/output/mpc_simulation.tikz:29-48, mean of 15 runs)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from spock_b200 import SP, SolverOptions, build_model, get_server_heat_specs, solve_model_  # noqa: E402

REF = [3.216, 1.020, 0.675, 0.418, 0.193, 0.108, 0.120, 0.060, 0.053, 0.056, 0.047, 0.063, 0.037, 0.042, 0.054, 0.056,
       0.036, 0.066, 0.040, 0.034]
N, d, nx = 10, 2, 20
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec = get_server_heat_specs(N, nx, d)
dyn = spec[2]
lat = np.zeros((runs, 20))
passes = np.zeros((runs, 20), int)
for m in range(runs):
    model = build_model(*spec, SolverOptions(SP, None, factor_mode=1))
    rng = np.random.default_rng(1)
    x0 = 0.1 * np.ones(nx)
    for t in range(20):
        t0 = time.perf_counter()
        solve_model_(model, x0, tol=1e-3, quiet=True)
        lat[m, t] = time.perf_counter() - t0
        passes[m, t] = model.last_result["passes"]
        u = model.solver_state.z[model.solver_state_internal.u_inds[:nx]]
        w = int(rng.integers(1, d + 1)) - 1
        x0 = dyn.A[w] @ x0 + dyn.B[w] @ u
    model.handle.close()
print(json.dumps({"config": "server_heat N=10 d=2 nx=nu=20, SP, tol 1e-3, warm start, 20 closed-loop steps",
                  "runs": runs, "latency_s_mean": lat.mean(0).round(5).tolist(), "passes_first_run": passes[0].tolist(),
                  "ref_julia_latency_s": REF, "x_norm_final": float(np.linalg.norm(x0))}))
