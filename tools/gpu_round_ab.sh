#!/bin/bash
# 1-GPU A/B of the GEMM operand-ring budget (1 vs 2 CTAs/SM) + parity of the tree as committed.
mkdir -p gpurun_out
T=${1:-r19}
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15) > gpurun_out/${T}_pytest.log
(DLRM_TC_BWD_SMEM_KB=100 DLRM_TC_FWD_SMEM_KB=100 timeout 300 python -m pytest tests/test_gpu_engine_tc.py tests/test_gpu_gemm_tc.py -q -m gpu -x 2>&1 | tail -8) > gpurun_out/${T}_pytest_smem100.log
B="python bench.py --steps 200 --warmup 20 --no-cpu"
(timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_a_default.log
(DLRM_TC_BWD_SMEM_KB=100 timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_b_bwd100.log
(DLRM_TC_BWD_SMEM_KB=100 DLRM_TC_FWD_SMEM_KB=100 timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_c_all100.log
(DLRM_TUNE="head_rows=32,interact_bwd_cols=1" timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_d_oldkernels.log
(DLRM_TC_BWD_SMEM_KB=100 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/${T}_launches_bwd100.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1)
tail -4 gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest_smem100.log
for f in a_default b_bwd100 c_all100 d_oldkernels; do echo "== $f"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_$f.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), "upd", (d.get("roofline_update") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_$f.log").read()[-1200:])
PY
done
