OUT=gpurun_out; TAG=r02w
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_config3_noextras.json 2> $OUT/${TAG}_bench_config3.err
