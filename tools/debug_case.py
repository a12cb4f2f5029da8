"""One parity case in a fresh process (a faulting kernel poisons the CUDA context): prints OK + errors or the library error.
usage: python tools/debug_case.py N nx d risk alg K factor_mode"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle.oracle import Oracle  # noqa: E402  (checker)
from spock_b200.examples import get_server_heat_specs  # noqa: E402
from spock_b200.handle import Handle  # noqa: E402
from spock_b200.problem import flatten_problem  # noqa: E402

N, nx, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
risk, alg, K, fm = sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
flat = flatten_problem(*get_server_heat_specs(N, nx, d, risk=risk))
x0 = 0.1 * np.ones(flat.nx)
try:
    h = Handle(flat, alg, factor_mode=fm)
    h.set_x0(x0)
    h.step(K)
    z, v = h.get_iterate()
    o = Oracle(flat, alg)
    o.set_x0(x0)
    o.step(K)
    print("OK", float(np.abs(z - o.z).max() / np.abs(o.z).max()), float(np.abs(v - o.v).max() / max(np.abs(o.v).max(), 1e-300)))
except Exception as e:  # noqa: BLE001
    print("FAIL", str(e)[:200])
