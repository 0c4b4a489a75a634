#!/bin/bash
# round 2, call 10: cfg3 bench with the lean update kernel + ncu --set full of the update / gather kernels (source view)
T=r2_10
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu --no-check 2>gpurun_out/${T}_bench_cfg3.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3_lean.json
(DLRM_TUNE=upd_lean=2 timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu --no-check 2>/dev/null | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3_general.json
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','avg_launch_us')})
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline_update') or {}).items() if k in ('achieved','frac','avg_launch_us','train_gather_plus_link_us')})
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'emb_update_lean|emb_fwd_vec' -s 6 -c 4 \
  -o gpurun_out/${T}_emb python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu --no-check --no-graph > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log | cut -c1-200
ls -la gpurun_out/${T}_emb.ncu-rep
