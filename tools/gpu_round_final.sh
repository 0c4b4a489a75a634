#!/bin/bash
# 1-GPU validation of the tree as committed: GPU parity tests, smoke(), default bench line, launch list.
mkdir -p gpurun_out
T=${1:-r18}
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8) > gpurun_out/${T}_smoke.log
(timeout 400 python bench.py 2>&1 | tail -2) > gpurun_out/${T}_bench_default.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1)
tail -6 gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_smoke.log
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_default.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","steps","warmup") if k in d}, "e2e", d.get("e2e",{}).get("value"), "roof", d.get("roofline"), "upd", (d.get("roofline_update") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_default.log").read()[-1500:])
PY
wc -l gpurun_out/${T}_launches.csv
