# A/B of the top-of-tree kernels (SPK_TOPK=1; default: one launch per level) on latency-bound trees and on config 3
for nt in 0 1; do
  echo "SPK_TOPK=$nt"
  for args in "10 2 20 20 0 300 1 heat" "10 2 20 20 1 300 1 heat" "10 2 20 20 1 100 2 heat" "7 2 5 5 0 500 1 heat" "13 2 50 20 0 200" "17 2 50 20 0 50" "20 2 50 20 0 30" "11 4 10 10 0 50 1 heat"; do
    SPK_TOPK=$nt timeout 300 python tools/quick_time.py $args | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['N'],d['d'],d['nx'],'fm',d['factor_mode'],'alg',d['alg'],'us/pass %.1f'%(1000*d['ms_per_pass']),'launches/pass',d['launches_per_pass'], d['xi'])"
  done
done
