OUT=gpurun_out; TAG=r02v
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests.log
qt() { python tools/quick_time.py "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/pass %.4f'%d['ms_per_pass'], d['profile_ms_per_pass'], 'xi', d['xi'])"; }
for rep in 1 2; do
  echo -n "staged  config5-slice " >> $OUT/${TAG}_s2_ab.txt; qt 11 4 10 10 0 30 1 heat >> $OUT/${TAG}_s2_ab.txt 2>&1
  echo -n "direct  config5-slice " >> $OUT/${TAG}_s2_ab.txt; SPK_NO_S2_STAGED=1 qt 11 4 10 10 0 30 1 heat >> $OUT/${TAG}_s2_ab.txt 2>&1
done
echo -n "staged  config3 " >> $OUT/${TAG}_s2_ab.txt; qt 20 2 50 20 0 20 >> $OUT/${TAG}_s2_ab.txt 2>&1
echo -n "direct  config3 " >> $OUT/${TAG}_s2_ab.txt; SPK_NO_S2_STAGED=1 qt 20 2 50 20 0 20 >> $OUT/${TAG}_s2_ab.txt 2>&1
echo -n "staged  N=10 nx=20 " >> $OUT/${TAG}_s2_ab.txt; qt 10 2 20 20 0 300 1 heat >> $OUT/${TAG}_s2_ab.txt 2>&1
echo -n "direct  N=10 nx=20 " >> $OUT/${TAG}_s2_ab.txt; SPK_NO_S2_STAGED=1 qt 10 2 20 20 0 300 1 heat >> $OUT/${TAG}_s2_ab.txt 2>&1
timeout 600 python bench.py --workload config5 --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_config5.json 2> $OUT/${TAG}_bench_config5.err
