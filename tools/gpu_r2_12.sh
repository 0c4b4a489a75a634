#!/bin/bash
# round 2, call 12: what bounds the embedding update?  (timing-only variants that skip one memory stream each) + launch lists
T=r2_12
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 3 12 15; do
  (DLRM_TUNE=upd_debug=$dbg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-check 2>/dev/null | grep '^{' | tail -1) > gpurun_out/${T}_upd_debug_$dbg.json
  python -c "
import json
d=json.loads(open('gpurun_out/${T}_upd_debug_$dbg.json').read()); u=d['roofline_update']
print('upd_debug=$dbg update %.1f us  gather+link %.1f us  step %.3f ms' % (u['avg_launch_us'], u['train_gather_plus_link_us'], d['ms_per_step']))"
done
K="regex:emb_|gemm_|interact|head_kernel|dense_update|split_bf16|block_copy|reduce_partials|p2p_"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 150 -c 80 --csv --log-file gpurun_out/${T}_launches_cfg3.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu_cfg3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 150 -c 80 --csv --log-file gpurun_out/${T}_launches_cfg2.csv \
  python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu_cfg2.log 2>&1
python - <<'PY'
import csv
for w in ("cfg3","cfg2"):
    with open("gpurun_out/r2_12_launches_%s.csv"%w) as f:
        lines=[l for l in f if not l.startswith("==")]
    hi=[i for i,l in enumerate(lines) if l.startswith('"ID"')]
    if not hi: print(w, "no kernels"); continue
    rows=[(x["Kernel Name"][:52], float(x["Metric Value"].replace(",",""))/1e3, x.get("Grid Size")) for x in csv.DictReader(lines[hi[0]:])]
    names=[r[0] for r in rows]
    st=[i for i,n in enumerate(names) if "split_bf16" in n]
    seg=rows[st[-2]:st[-1]] if len(st)>=2 else rows[-40:]
    print("==", w, "one step:", len(seg), "launches, sum %.1f us"%sum(r[1] for r in seg))
    for r in seg: print("  %-54s %9.1f us %s"%r)
PY
