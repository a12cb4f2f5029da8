#!/usr/bin/env bash
# How the files under profiles/ were produced (one B200; run through `gpurun -- bash tools/collect_profiles.sh <tag>`).
# A number printed by a run under ncu is never a bench value: bench lines come from the plain runs only.
set -x
TAG=${1:-r01x}
OUT=gpurun_out
# 1. bench lines (config 3 default = per-node factors, with the per_stage_factors block; config 5)
python bench.py > $OUT/${TAG}_bench_config3.json 2> $OUT/${TAG}_bench_config3.err
python bench.py --workload config5 --steps 10 > $OUT/${TAG}_bench_config5.json 2> $OUT/${TAG}_bench_config5.err
# 2. the reference's own experiments and the latency-bound configurations
python tools/scaling_sweep.py 15 > $OUT/${TAG}_scaling_sweep.jsonl
python tools/mpc_simulation.py 3 > $OUT/${TAG}_mpc_simulation.json
python tools/small_tree_latency.py > $OUT/${TAG}_small_tree_latency.jsonl
python tools/nodes_sweep.py 8 20 14 > $OUT/${TAG}_nodes_sweep.jsonl
# 3. ncu launch lists of the bench command (kernel shares of a pass), per-node and per-stage factors
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_config3.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_config3_per_stage.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --factor-mode per_stage > $OUT/${TAG}_ncu_l2.log 2>&1
# 4. ncu --set full of the deepest-level launch of every hot kernel (launch-skip picks that launch of the first timed pass)
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 19 -c 1 -o $OUT/${TAG}_bwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_s1 --launch-skip 18 -c 1 -o $OUT/${TAG}_fwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o $OUT/${TAG}_dual3 \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd_mm --launch-skip 10 -c 1 -o $OUT/${TAG}_bwd_mm \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_mm --launch-skip 9 -c 1 -o $OUT/${TAG}_fwd_mm \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_f5.log 2>&1
# 5. summaries for profiles/ (run these here, after the .ncu-rep files have been merged back):
#    python tools/ncu_summary.py launches $OUT/${TAG}_launches_config3.csv > profiles/${TAG}_launches_config3.txt
#    (python tools/ncu_summary.py report $OUT/${TAG}_bwd.ncu-rep; python tools/ncu_lines.py $OUT/${TAG}_bwd.ncu-rep 25 samples) > profiles/${TAG}_bwd_config3_ncu.txt
