"""Development probe: time K passes of the CP/SP loop at a given size and print the per-kernel-class breakdown.
usage: python tools/quick_time.py N d nx nu [factor_mode=0] [K=20] [alg=1] [gen=random|heat]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from spock_b200.examples import get_random_system_specs, get_server_heat_specs  # noqa: E402
from spock_b200.handle import Handle  # noqa: E402
from spock_b200.problem import flatten_problem  # noqa: E402

N, d, nx, nu = (int(a) for a in sys.argv[1:5])
fm = int(sys.argv[5]) if len(sys.argv) > 5 else 0
K = int(sys.argv[6]) if len(sys.argv) > 6 else 20
alg = int(sys.argv[7]) if len(sys.argv) > 7 else 1
gen = sys.argv[8] if len(sys.argv) > 8 else "random"
spec = get_random_system_specs(N, d, nx, nu) if gen == "random" else get_server_heat_specs(N, nx, d)
flat = flatten_problem(*spec)
t0 = time.time()
h = Handle(flat, alg, factor_mode=fm)
t_create = time.time() - t0
h.set_x0(0.1 * np.ones(nx))
h.step(3)
ms = h.step(K)
h.profile(True)
ms_p = h.step(K)
rep = h.profile_report()
h.profile(False)
n, nnl = flat.n, flat.n_non_leaf
b_vec = 8 * (3 * (flat.nz + flat.nv) + 2 * (n * nx + nnl * nu))
b_fac = 8 * (2 * nx * nx * (n - 1) + (3 * nu * nx + nu * nu) * nnl) if fm == 0 else \
    8 * (2 * nx * nx * (N - 1) * d + (3 * nu * nx + nu * nu) * (N - 1))
out = {
    "N": N, "d": d, "nx": nx, "nu": nu, "factor_mode": fm, "alg": alg, "nodes": n, "create_s": round(t_create, 2),
    "ms_per_pass": ms / K, "passes_per_s": 1000.0 * K / ms, "ms_per_pass_profiled": ms_p / K,
    "B_iter_GB": (b_vec + b_fac) / 1e9, "achieved_GBs": (b_vec + b_fac) / 1e9 / (ms / K / 1000.0),
    "device_GB": h.device_bytes() / 1e9, "xi": h.residuals(),
    "profile_ms_per_pass": {k: round(v["ms"] / K, 4) for k, v in rep.items() if v["launches"]},
    "launches_per_pass": sum(v["launches"] for v in rep.values()) / K,
}
print(json.dumps(out))
