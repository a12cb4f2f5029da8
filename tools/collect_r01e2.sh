set -x
python bench.py > gpurun_out/r01e_bench_config3.json 2> gpurun_out/r01e_bench_config3.err
python bench.py --workload config5 --steps 10 > gpurun_out/r01e_bench_config5.json 2> gpurun_out/r01e_bench_config5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches_config3_per_stage.csv python bench.py --steps 2 --warmup 3 --no-extras --factor-mode per_stage > gpurun_out/r01e_ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd_mm --launch-skip 12 -c 1 -o gpurun_out/r01e_bwd_mm python tools/quick_time.py 20 2 50 20 1 1 > gpurun_out/r01e_ncu_f2.log 2>&1
python tools/quick_time.py 20 2 50 20 1 20
