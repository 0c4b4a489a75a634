#!/bin/bash
# round 2, call 11: launch lists of one cfg3 / cfg2 step (our kernels only), shard tests incl. remote reads
T=r2_11
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.log | head; grep -E "PASS|FAIL" gpurun_out/${T}_pytest.log | head -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dlrm' -s 150 -c 80 --csv --log-file gpurun_out/${T}_launches_cfg3.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu_cfg3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dlrm' -s 150 -c 80 --csv --log-file gpurun_out/${T}_launches_cfg2.csv \
  python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu_cfg2.log 2>&1
python - <<'PY'
import csv
for w in ("cfg3","cfg2"):
    with open("gpurun_out/r2_11_launches_%s.csv"%w) as f:
        lines=[l for l in f if not l.startswith("==")]
    rows=[(x["Kernel Name"][:52], float(x["Metric Value"].replace(",",""))/1e3, x.get("Grid Size")) for x in csv.DictReader(lines)]
    names=[r[0] for r in rows]
    st=[i for i,n in enumerate(names) if "split_bf16" in n]
    if len(st)>=2:
        seg=rows[st[-2]:st[-1]]
    else: seg=rows[-40:]
    print("==", w, "one step:", len(seg), "launches, sum %.1f us"%sum(r[1] for r in seg))
    for r in seg: print("  %-54s %9.1f us %s"%r)
PY
for tune in "" "upd_lean=3" "upd_lean=2"; do
  (DLRM_TUNE=$tune timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-check 2>/dev/null | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3_${tune:-default}.json
done
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline_update') or {}).items() if k in ('achieved','frac','avg_launch_us','train_gather_plus_link_us')})
except Exception as e: print('ERR', repr(e), open('$f').read()[-300:])
"; done
