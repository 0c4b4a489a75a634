#!/bin/bash
# round 2, 8-GPU call: cfg3 (MLPerf-DLRM synthetic, the default bench workload) at 8 ranks with its phase timeline, then
# the sharded parity cases against the live-reference goldens at 8 ranks
T=r2_n8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29601 bench.py --gpus 8 --steps 40 --warmup 6 --phases 5 2>gpurun_out/${T}_bench_cfg3.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3.json
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${T}_bench_cfg3.err | tail -4 | cut -c1-1500
(timeout 200 $TR --master-port 29599 tests/dist_check.py 2>&1 | grep -E "PASS|FAIL|Error" | head -12) > gpurun_out/${T}_dist_check.log; cat gpurun_out/${T}_dist_check.log
python -c "
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_cfg3.json').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    print(d.get('placement')); print(d.get('parity_check')); print(d.get('nvlink')); print(d.get('roofline')); print(d.get('roofline_update'))
except Exception as e: print('ERR', repr(e))
"
