#!/bin/bash
# One 1-GPU call: all GPU tests, smoke, sweep, benches, ncu captures.  Outputs in gpurun_out/.
mkdir -p gpurun_out
T=${1:-r8}
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -60) > gpurun_out/${T}_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5) > gpurun_out/${T}_smoke.log
(timeout 300 python tools/sweep_gather.py gpurun_out/${T}_sweep.json 2>&1 | tail -8) > gpurun_out/${T}_sweep.log
(timeout 400 python bench.py --steps 200 --warmup 20 --gemm tc --cpu-budget 10 2>&1 | tail -2) > gpurun_out/${T}_bench_tc.log
(timeout 400 python bench.py --steps 200 --warmup 20 --gemm tc --no-graph --no-cpu 2>&1 | tail -2) > gpurun_out/${T}_bench_tc_nograph.log
(timeout 400 python bench.py --steps 200 --warmup 20 --workload cfg1 --no-cpu 2>&1 | tail -2) > gpurun_out/${T}_bench_cfg1.log
(timeout 400 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -2) > gpurun_out/${T}_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches_tc.csv python bench.py --steps 3 --warmup 3 --gemm tc --no-cpu --no-graph > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"emb_fwd_vec|emb_update|emb_classify|emb_link" -s 8 -c 8 -o gpurun_out/${T}_emb -f python bench.py --steps 3 --warmup 3 --gemm tc --no-cpu --no-graph > gpurun_out/${T}_ncu_emb.log 2>&1
tail -15 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_smoke.log; tail -6 gpurun_out/${T}_sweep.log
for f in tc tc_nograph cfg1 reference; do echo "== $f"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_$f.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("avg_launch_us"), "upd", (d.get("roofline_update") or {}).get("frac"), (d.get("roofline_update") or {}).get("avg_launch_us"), (d.get("roofline_update") or {}).get("train_gather_plus_link_us"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_$f.log").read()[-1500:])
PY
done
