# A/B of several kernel variants inside ONE gpurun (same box), alternating runs:
#   tools/ab_variants.sh "<variant> <variant> ..." [quick_time args ...]      (variants built with spock_b200.build --variant)
VS=$1; shift
ARGS=${@:-20 2 50 20 0 30}
for rep in 1 2; do
  for lib in default $VS; do
    if [ $lib = default ]; then unset SPOCK_B200_LIB; else export SPOCK_B200_LIB=$PWD/spock_b200/csrc/_variants/$lib.so; fi
    python tools/quick_time.py $ARGS | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'ms/pass %.4f'%d['ms_per_pass'], d['profile_ms_per_pass'], 'xi', d['xi'])"
  done
done
