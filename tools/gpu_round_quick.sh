#!/bin/bash
mkdir -p gpurun_out
T=${1:-r10}
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
for kp in 4 2 1; do
 (timeout 300 python bench.py --steps 200 --warmup 20 --gemm tc --no-cpu --pipeline $kp 2>&1 | tail -2) > gpurun_out/${T}_bench_p$kp.log
done
(timeout 300 python bench.py --steps 200 --warmup 20 --gemm tc --no-cpu --no-graph 2>&1 | tail -2) > gpurun_out/${T}_bench_nograph.log
(timeout 300 python bench.py --steps 50 --warmup 5 --cpu-budget 8 2>&1 | tail -2) > gpurun_out/${T}_bench_driverlike.log
tail -8 gpurun_out/${T}_pytest.log
for f in p4 p2 p1 nograph driverlike; do echo "== $f"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_$f.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","steps_per_graph","cuda_graph") if k in d}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), "upd", (d.get("roofline_update") or {}).get("frac"), (d.get("roofline_update") or {}).get("avg_launch_us"), (d.get("roofline_update") or {}).get("train_gather_plus_link_us"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_$f.log").read()[-1500:])
PY
done
