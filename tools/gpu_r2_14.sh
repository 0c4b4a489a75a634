#!/bin/bash
# round 2, call 14 (2 GPUs): new CLI tests; phase timelines of cfg3 at N=1 and N=2; cfg4 (row-split giant table) at N=2,
# partial-sum forward and remote-row-read forward
T=r2_14
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_facade.py -q -m gpu -x -k "cli or checkpoint" > gpurun_out/${T}_pytest_cli.log 2>&1; tail -3 gpurun_out/${T}_pytest_cli.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { # name, then bench args
  n=$1; shift
  (timeout 900 "$@" 2>gpurun_out/${T}_$n.err | grep '^{' | tail -1) > gpurun_out/${T}_$n.json
  tail -2 gpurun_out/${T}_$n.err | cut -c1-300
}
run cfg3_n1 python bench.py --steps 40 --warmup 6 --no-check --no-cpu --phases 10
run cfg3_n2 $TR --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 6 --no-check --phases 10
run cfg4_n2_partial $TR --master-port 29613 bench.py --gpus 2 --workload cfg4 --steps 30 --warmup 5 --no-check --phases 5
run cfg4_n2_remote $TR --master-port 29615 bench.py --gpus 2 --workload cfg4 --steps 30 --warmup 5 --no-check --phases 5 --split-forward remote
for f in gpurun_out/${T}_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])
    print(d.get('placement')); print(d.get('nvlink')); print(d.get('roofline') and d['roofline']['avg_launch_us'], d.get('roofline_update') and d['roofline_update']['avg_launch_us'])
    for m in (d.get('phases') or {}).get('marks', []): print('   ', m)
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
