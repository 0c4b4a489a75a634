#!/bin/bash
# round 2, call 9: lean update kernel: parity tests, cfg3 + cfg2 bench A/B (lean vs general update kernel)
T=r2_09
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shards.py tests/test_gpu_dist.py tests/test_gpu_engine_tc.py -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head -30
for mode in lean general; do
  TUNE=""; [ $mode = general ] && TUNE="upd_lean=2"
  (DLRM_TUNE=$TUNE timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu --no-check 2>/dev/null | tail -1) > gpurun_out/${T}_bench_cfg3_${mode}.json
  (DLRM_TUNE=$TUNE timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu --no-check 2>/dev/null | tail -1) > gpurun_out/${T}_bench_cfg2_${mode}.json
done
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','avg_launch_us')})
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline_update') or {}).items() if k in ('achieved','frac','avg_launch_us','train_gather_plus_link_us')})
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
