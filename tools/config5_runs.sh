#!/usr/bin/env bash
# config 5 (d = 4, nx = nu = 10, SOC risk) strong / weak scaling points: gpurun --gpus N -- bash tools/config5_runs.sh N "<workloads>" <tag>
N=$1; WL=$2; TAG=${3:-r02h}; OUT=gpurun_out
for w in $WL; do
  if [ "$N" = "1" ]; then
    python bench.py --workload $w --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_${w}_g1.json 2> $OUT/${TAG}_${w}_g1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --workload $w --gpus $N --steps 20 --warmup 5 --no-extras \
      > $OUT/${TAG}_bench_${w}_g$N.json 2> $OUT/${TAG}_${w}_g$N.err
  fi
  python - "$OUT/${TAG}_bench_${w}_g$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 2), "it/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["whole_pass"]["frac"], 3),
          "bytes", d.get("device_bytes"), "mg_parity", (d.get("multi_gpu_parity") or {}).get("ok"), d["roofline"]["per_kernel_ms_per_pass"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
