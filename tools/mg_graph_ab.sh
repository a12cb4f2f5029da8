# A/B of the cut-level exchange: peer-memory kernels (default) vs host-issued NCCL all-gathers (SPK_NO_PEER=1)
G=${1:-2}
if [ -z "$SKIP_CHECK" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29520 tools/multigpu_check.py 2>&1 | tail -2 | cut -c1-400; fi
for np in 1 0; do
  echo "SPK_NO_PEER=$np"
  SPK_NO_PEER=$np timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$np bench.py --gpus $G --steps 50 --warmup 5 --no-extras --workload ${2:-config3_half} 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read().strip(); 
try:
    d=json.loads(l); print('it/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('multi_gpu_parity',{}).get('ok'), 'median', d['protocol']['value_median'], 'bytes', d['device_bytes'], 'launches', d['gpu_launches'], 'exch_ms', d.get('exchange',{}).get('ms_per_pass_instrumented'))
except Exception as e: print('ERR', l[-800:])"
done
if [ -z "$SKIP_CHECK" ]; then timeout 300 python tools/device_group_check.py $G 2>&1 | tail -1 | cut -c1-300; fi
