OUT=gpurun_out; TAG=r02u
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_config5.csv \
    python bench.py --workload config5 --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_ncu_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_dual3 --launch-skip 1 -c 1 -o $OUT/${TAG}_dual3 \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_s1 --launch-skip 18 -c 1 -o $OUT/${TAG}_fwd \
    python tools/quick_time.py 20 2 50 20 0 1 > $OUT/${TAG}_ncu_3.log 2>&1
