# round-1 (e) measurement set: bench lines, reference experiments, ncu launch lists, ncu --set full of the new kernels
set -x
python bench.py > gpurun_out/r01e_bench_config3.json 2> gpurun_out/r01e_bench_config3.err
python bench.py --workload config5 --steps 10 > gpurun_out/r01e_bench_config5.json 2> gpurun_out/r01e_bench_config5.err
python tools/scaling_sweep.py 15 > gpurun_out/r01e_scaling_sweep.jsonl 2> gpurun_out/r01e_sweep.err
python tools/mpc_simulation.py 3 > gpurun_out/r01e_mpc_simulation.json 2> gpurun_out/r01e_mpc.err
python tools/small_tree_latency.py > gpurun_out/r01e_small_tree_latency.jsonl 2> gpurun_out/r01e_lat.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r01e_ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches_config3_per_stage.csv python bench.py --steps 2 --warmup 3 --no-extras --factor-mode per_stage > gpurun_out/r01e_ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o gpurun_out/r01e_dual3 python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01e_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd_mm --launch-skip 12 -c 1 -o gpurun_out/r01e_bwd_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01e_ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_mm --launch-skip 11 -c 1 -o gpurun_out/r01e_fwd_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01e_ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_top_mm --launch-skip 1 -c 1 -o gpurun_out/r01e_top_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01e_ncu_f4.log 2>&1
tail -c 600 gpurun_out/r01e_bench_config3.json
