"""Builds tests/hostemu/_build/libspock_b200_hostemu.so: the product's kernels and solver orchestration
compiled as SERIAL HOST CODE (-DSPK_HOST_EMU, g++).  Test infrastructure only: it lets the CPU test-suite
check index arithmetic and the CP/SuperMann orchestration against the oracle in a container without a GPU.
It is never loaded by the product (spock_b200._lib.load only ever opens csrc/libspock_b200.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spock_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libspock_b200_hostemu.so")


def build(force=False):
    srcs = [os.path.join(CSRC, "spk_solver.cu"), os.path.join(CSRC, "spk_offline.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared",
           "-DSPK_HOST_EMU", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
           "-x", "c++", srcs[0], srcs[1], "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host-emulation build failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
