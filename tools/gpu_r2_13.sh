#!/bin/bash
# round 2, call 13: row layout with accumulator + list head inside the row (default): full GPU suite, cfg3 / cfg2 bench
T=r2_13
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head -20
(timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu 2>gpurun_out/${T}_cfg3.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3.json
(timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu --no-check 2>/dev/null | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg2.json
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('parity_check'))
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','avg_launch_us')})
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline_update') or {}).items() if k in ('achieved','frac','avg_launch_us','train_gather_plus_link_us','train_gather')})
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
tail -3 gpurun_out/${T}_cfg3.err | cut -c1-300
