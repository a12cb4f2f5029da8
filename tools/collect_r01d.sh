# round-1 (d) measurement set: bench lines, ncu launch lists, ncu --set full of the new kernels
set -x
python bench.py > gpurun_out/r01d_bench_config3.json 2> gpurun_out/r01d_bench_config3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r01d_ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_config3_per_stage.csv python bench.py --steps 2 --warmup 3 --no-extras --factor-mode per_stage > gpurun_out/r01d_ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o gpurun_out/r01d_dual3 python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01d_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd_mm --launch-skip 19 -c 1 -o gpurun_out/r01d_bwd_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01d_ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_mm --launch-skip 18 -c 1 -o gpurun_out/r01d_fwd_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01d_ncu_f3.log 2>&1
./tools/microbench/dmma_bench > gpurun_out/r01d_dmma_microbench.txt 2>&1
tail -c 1500 gpurun_out/r01d_bench_config3.json
