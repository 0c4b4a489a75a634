#!/bin/bash
# round 2, call 6: straight-line epilogue (hoisted args): GEMM check, chain tests, suite, A/B bench, timeline
T=r2_06
mkdir -p gpurun_out
timeout 600 python tests/gemm_tc_check.py gpurun_out/${T}_gemm_tc_check.txt | tail -1
grep "^FAIL" gpurun_out/${T}_gemm_tc_check.txt | head -20
timeout 900 python -X faulthandler -m pytest tests/test_gpu_chain.py -q -m gpu > gpurun_out/${T}_chain_full.log 2>&1
grep -n "max|diff|\|losses \|passed\|failed" gpurun_out/${T}_chain_full.log | head -40
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head -30
B="python bench.py --steps 200 --warmup 20 --no-cpu"
(timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain.json
(DLRM_CHAIN=0 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_nochain.json
(DLRM_CHAIN_TILE_N=64 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain_tn64.json
(DLRM_CHAIN_TILE_N=128 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain_tn128.json
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as e: print('ERR', open('$f').read()[-300:])
"; done
timeout 300 python tools/chain_timeline.py ${T} 2>&1 | tail -40
DLRM_CHAIN_TILE_N=128 timeout 300 python tools/chain_timeline.py ${T}_tn128 2>&1 | tail -40
