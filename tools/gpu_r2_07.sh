#!/bin/bash
# round 2, call 7: chain kernel v3 (m-major order, 8 epilogue warps), shard tests, cfg2 A/B bench with the new bench.py
T=r2_07
mkdir -p gpurun_out
timeout 600 python tests/gemm_tc_check.py gpurun_out/${T}_gemm_tc_check.txt | tail -1
grep "^FAIL" gpurun_out/${T}_gemm_tc_check.txt | head -20
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head -30
B="python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu"
(timeout 300 $B 2>gpurun_out/${T}_bench_chain.err | tail -1) > gpurun_out/${T}_bench_chain.json
(DLRM_CHAIN=0 timeout 300 $B 2>/dev/null | tail -1) > gpurun_out/${T}_bench_nochain.json
(DLRM_TUNE=chain_order=1 timeout 300 $B 2>/dev/null | tail -1) > gpurun_out/${T}_bench_chain_layermajor.json
(DLRM_CHAIN_TILE_N=128 timeout 300 $B 2>/dev/null | tail -1) > gpurun_out/${T}_bench_chain_tn128.json
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d.get('parity_check'))
except Exception as e: print('ERR', open('$f').read()[-300:])
"; done
tail -5 gpurun_out/${T}_bench_chain.err
timeout 300 python tools/chain_timeline.py ${T} 2>&1 | tail -24
