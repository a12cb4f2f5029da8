#!/usr/bin/env bash
# Final measurements of the round for one value of N (GPUs of the box): gpurun --gpus N -- bash tools/final_runs.sh N <tag>
N=$1; TAG=${2:-r02g}; OUT=gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_config3.json 2> $OUT/${TAG}_bench_config3.err
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_ref_config3.json 2> $OUT/${TAG}_ref_config3.err
  python bench.py --workload config4 --steps 3 --warmup 3 > $OUT/${TAG}_bench_config4.json 2> $OUT/${TAG}_c4.err
  python bench.py --workload config2 --steps 3 --warmup 3 > $OUT/${TAG}_bench_config2.json 2> $OUT/${TAG}_c2.err
  python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/${TAG}_gpu_tests.log
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 \
      > $OUT/${TAG}_bench_config3_g$N.json 2> $OUT/${TAG}_g$N.err
  timeout 600 python bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_config3_group$N.json 2> $OUT/${TAG}_group$N.err
  if [ "$N" = "2" ]; then python -m pytest tests -m gpu -q -k "partition_over_gpus or device_group" 2>&1 | tail -3 > $OUT/${TAG}_gpu_tests_g2.log; fi
fi
for f in $OUT/${TAG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("impl", "b200"), d["metric"], round(d["value"], 3), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 2),
          "mode", d.get("launch_mode"), "bytes", d.get("device_bytes"), "mg_parity", (d.get("multi_gpu_parity") or {}).get("ok"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -2 $OUT/${TAG}_*.log 2>/dev/null
