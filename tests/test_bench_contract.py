"""bench.py's JSON contract on CPU: the reference arm (the one arm that runs without a GPU) prints one line with the
contract's keys, the machine-readable sample description, and the SAME `config` dict the GPU arm prints for the same
command line (the driver compares the two arms)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_and_config_identity():
    line = _run("--impl", "reference", "--workload", "small", "--steps", "5", "--warmup", "3")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "cp_iterations_per_sec" and line["dtype"] == "f64"
    assert line["vs_baseline"] is None and line["higher_is_better"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["extrapolated"] is False and cb["sample_nodes"] == line["config"]["nodes"] == 4095
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # ms_per_step x steps is the time really spent in the timed passes
    assert abs(line["ms_per_step"] * line["steps"] / 1e3 - cb["details"]["seconds"]) < 1e-6 * max(1.0, cb["details"]["seconds"])
    # the GPU arm builds its config with the same function from the same arguments
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    args = argparse.Namespace(workload="small", factor_mode="per_node", gpus=1)
    assert bench.config_dict(args, 1) == line["config"]
    n, nnl, nl, nz, nv = bench.tree_sizes("config3")
    assert (n, nz, nv) == (1048575, 67633074, 168820516)


def test_latency_workload_reference_arm():
    line = _run("--impl", "reference", "--workload", "config2", "--steps", "1", "--warmup", "1")
    assert line["metric"] == "sp_solves_per_sec" and line["unit"] == "solves/s" and line["value"] > 0
    assert "step" in line["config"]
