"""SASS evidence without a GPU: per-kernel mnemonic histogram of one shape-specialised translation unit
(cuobjdump -sass of spock_b200/csrc/_obj/spk_shape_<nx>_<nu>_<d>.o), restricted to the mnemonics that prove what the
kernels use -- TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), cp.async (LDGSTS), fp64 FMA / tensor pipe (DFMA, DMMA),
16-byte shared loads (LDS.128), system-scope release / acquire of the peer exchange, programmatic dependent launch.
    python tools/sass_histogram.py 50 20 2 > profiles/r02_sass_50_20_2.txt"""
import re
import subprocess
import sys
from collections import Counter, defaultdict

nx, nu, d = sys.argv[1:4]
objs = [f"spock_b200/csrc/_obj/spk_shape_{nx}_{nu}_{d}.o", "spock_b200/csrc/_obj/spk_solver.o"]
KEYS = ["UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMMA", "DADD", "DMUL", "LDS.128", "LDS.64", "LDS", "LDG.E.128", "LDG.E.64", "STG", "REDUX",
        "ATOMG", "ACQBULK", "PREEXIT", "ACQ", "RELEASE", "SHFL", "BAR.SYNC", "MEMBAR", "UTCHMMA", "UTMALDG", "LDTM", "HMMA"]
for obj in objs:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, hist = None, defaultdict(Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            hist[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k in ("ACQ", "RELEASE") and k in op):
                    hist[cur][k] += 1
                    break
    print(f"## {obj}")
    for fn, h in sorted(hist.items()):
        if not any(s in fn for s in ("spk_k_bwd<", "spk_k_fwd_s1<", "spk_k_dual3<" + f"{nx}, {nu}, {d}, true, true, false", "spk_k_bwd_mm", "spk_k_fwd_mm",
                                     "spk_k_s2t<5", "spk_k_peer_exchange", "spk_k_bwd_top<" + nx)):
            continue
        body = ", ".join(f"{k} {v}" for k, v in h.most_common() if k != "_total")
        print(f"{fn}: {h['_total']} instructions; {body}")
