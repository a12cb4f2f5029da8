# end-of-round evidence on one B200 (plain runs first; a number printed under ncu is never a bench value)
OUT=gpurun_out; TAG=${1:-r02z}
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_config3.json 2> $OUT/${TAG}_bench_config3.err
timeout 300 python bench.py --workload config4 --steps 3 --warmup 3 > $OUT/${TAG}_bench_config4.json 2> $OUT/${TAG}_c4.err
timeout 300 python bench.py --workload config2 --steps 3 --warmup 3 > $OUT/${TAG}_bench_config2.json 2> $OUT/${TAG}_c2.err
SPK_LAT_FACTOR_MODE=0 SPK_LAT_LOOPS=0 timeout 300 python tools/small_tree_latency.py > $OUT/${TAG}_small_tree_latency.jsonl 2> $OUT/${TAG}_lat.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_config3.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 19 -c 1 -o $OUT/${TAG}_bwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_s1 --launch-skip 9 -c 1 -o $OUT/${TAG}_fwd_20_20_2 \
    python tools/quick_time.py 18 2 20 20 0 1 1 heat > $OUT/${TAG}_ncu_f2.log 2>&1
ls -la $OUT/${TAG}_*
