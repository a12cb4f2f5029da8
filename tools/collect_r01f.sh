set -x
python bench.py > gpurun_out/r01f_bench_config3.json 2> gpurun_out/r01f_bench_config3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r01f_ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_s1 --launch-skip 18 -c 1 -o gpurun_out/r01f_fwd python tools/quick_time.py 20 2 50 20 0 1 > gpurun_out/r01f_ncu_f6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd --launch-skip 19 -c 1 -o gpurun_out/r01f_bwd python tools/quick_time.py 20 2 50 20 0 1 > gpurun_out/r01f_ncu_f5.log 2>&1
