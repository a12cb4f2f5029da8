OUT=gpurun_out; TAG=r02s
# per-stage factors at config 3: the leaf-level and the deepest non-leaf-level launch of the tensor-pipe backward sweep, the deepest forward launch
ncu --set full --clock-control none --import-source on -k regex:spk_k_bwd_mm --launch-skip 19 -c 2 -o $OUT/${TAG}_bwd_mm \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spk_k_fwd_mm --launch-skip 18 -c 1 -o $OUT/${TAG}_fwd_mm \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_per_stage.csv \
    python tools/quick_time.py 20 2 50 20 1 1 > $OUT/${TAG}_ncu_3.log 2>&1
