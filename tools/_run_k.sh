OUT=gpurun_out; TAG=r02k
timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests.log
timeout 600 python bench.py --workload config5 --steps 20 --warmup 5 > $OUT/${TAG}_bench_config5.json 2> $OUT/${TAG}_bench_config5.err
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o $OUT/${TAG}_dual3_config5 \
    python tools/quick_time.py 12 4 10 10 0 1 1 heat > $OUT/${TAG}_ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 11 -c 1 -o $OUT/${TAG}_bwd_config5 \
    python tools/quick_time.py 12 4 10 10 0 1 1 heat > $OUT/${TAG}_ncu_f5.log 2>&1
