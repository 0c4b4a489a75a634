#!/bin/bash
# round 2, call 8: first cfg3 (MLPerf-DLRM synthetic) numbers at N=1; tests; cfg2 continuity; launch list
T=r2_08
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head -30
grep -E "PASS|FAIL" gpurun_out/${T}_pytest.log | head -12
(timeout 900 python bench.py --steps 100 --warmup 10 2>gpurun_out/${T}_bench_cfg3.err | tail -1) > gpurun_out/${T}_bench_cfg3.json
tail -5 gpurun_out/${T}_bench_cfg3.err
(timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu 2>gpurun_out/${T}_bench_cfg2.err | tail -1) > gpurun_out/${T}_bench_cfg2.json
(DLRM_CHAIN=1 timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu --no-check 2>/dev/null | tail -1) > gpurun_out/${T}_bench_cfg2_chain.json
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d.get('parity_check'))
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','avg_launch_us','algorithmic_bytes_per_launch')})
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline_update') or {}).items() if k in ('achieved','frac','avg_launch_us','train_gather_plus_link_us','train_gather')})
    print((d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'))
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${T}_launches_cfg3.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu_bench.log 2>&1
tail -2 gpurun_out/${T}_ncu_bench.log | cut -c1-300
