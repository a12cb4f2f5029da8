"""python tools/device_group_check.py [G]: ONE process, G GPUs behind one handle (spk_problem.n_devices / devices) against a
single-GPU handle -- Chambolle-Pock and SuperMann, K passes and solve_model!, on the headline shape with wide levels.
Prints one 'DEVICE_GROUP_OK ...' line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spock_b200.examples import get_random_system_specs, get_server_heat_specs  # noqa: E402
from spock_b200.handle import Handle  # noqa: E402
from spock_b200.problem import flatten_problem  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
assert G >= 2, "needs at least 2 GPUs"


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


out = []
for name, spec in (("random N=14 nx=50 nu=20", get_random_system_specs(14, 2, 50, 20)), ("heat N=9 nx=10", get_server_heat_specs(9, 10, 2))):
    flat = flatten_problem(*spec)
    x0 = 0.1 * np.ones(flat.nx)
    for alg, K, fm in ((1, 10, 0), (1, 10, 1), (2, 5, 1)):
        ref = Handle(flat, alg, factor_mode=fm, device=0)
        grp = Handle(flat, alg, factor_mode=fm, devices=list(range(G)))
        for h in (ref, grp):
            h.set_x0(x0)
            h.step(K)
        (zr, vr), (zg, vg) = ref.get_iterate(), grp.get_iterate()
        ez, ev = relerr(zg, zr), relerr(vg, vr)
        assert ez < 1e-9 and ev < 1e-9, (name, alg, fm, ez, ev)
        rr, _ = ref.solve(x0, tol=1e-3, z0=np.zeros(ref.nz), v0=np.zeros(ref.nv))
        rg, _ = grp.solve(x0, tol=1e-3, z0=np.zeros(ref.nz), v0=np.zeros(ref.nv))
        assert rg["converged"] == 1 and (alg == 2 or rg["passes"] == rr["passes"]), (rr, rg)
        out.append({"problem": name, "alg": "CP" if alg == 1 else "SP", "factor_mode": fm, "relerr_z": ez, "relerr_v": ev,
                    "passes_single": rr["passes"], "passes_group": rg["passes"], "solve_ms_single": rr["solve_ms"],
                    "solve_ms_group": rg["solve_ms"]})
        ref.close()
        grp.close()
print("DEVICE_GROUP_OK devices=%d " % G + json.dumps(out), flush=True)
