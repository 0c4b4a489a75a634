#!/bin/bash
# round 2, 8-GPU call: sharded parity vs live-reference goldens, CLI under torchrun, cfg3 / cfg2 bench at N=2
T=r2_n8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -3
TR0="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 600 $TR0 --master-port 29599 tests/dist_check.py 2>&1 | grep -E "PASS|FAIL|Error" | head -12) > gpurun_out/${T}_dist_check.log; cat gpurun_out/${T}_dist_check.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 900 $TR --master-port 29601 bench.py --gpus 8 --steps 60 --warmup 8 2>gpurun_out/${T}_bench_cfg3.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3.json
tail -3 gpurun_out/${T}_bench_cfg3.err
(timeout 900 $TR --master-port 29603 bench.py --gpus 8 --workload cfg2 --steps 200 --warmup 20 2>gpurun_out/${T}_bench_cfg2.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg2.json
tail -3 gpurun_out/${T}_bench_cfg2.err
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    print(d.get('placement')); print(d.get('parity_check'))
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
