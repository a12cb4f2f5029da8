#!/usr/bin/env bash
# Round-2 evidence (one B200): gpurun -- bash tools/collect_profiles_r02.sh <tag>
# A number printed by a run under ncu is never a bench value: bench lines come from the plain runs only.
TAG=${1:-r02e}
OUT=gpurun_out
# 1. plain bench run first (must exit 0 without ncu)
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_config3.json 2> $OUT/${TAG}_bench_config3.err || exit 1
# 2. launch list of the same command (kernel shares of a pass)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_config3.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_ncu_l1.log 2>&1
# 3. ncu --set full of the deepest-level launch of the hot kernels (quick_time: 3 warm-up passes, then K = 1)
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 19 -c 1 -o $OUT/${TAG}_bwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_s1 --launch-skip 18 -c 1 -o $OUT/${TAG}_fwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o $OUT/${TAG}_dual3 \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f3.log 2>&1
# config 5 (d = 4, nx = 10, SOC risk): the dual kernel on the leaf families and the backward kernel (no ncu existed)
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o $OUT/${TAG}_dual3_config5 \
    python tools/quick_time.py 12 4 10 10 0 1 1 heat > $OUT/${TAG}_ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 11 -c 1 -o $OUT/${TAG}_bwd_config5 \
    python tools/quick_time.py 12 4 10 10 0 1 1 heat > $OUT/${TAG}_ncu_f5.log 2>&1
ls -la $OUT/${TAG}_*
