#!/bin/bash
# First call of the next round: where does the step's time go (timeline), do the written-but-unmeasured
# kernels work (DLRM_EXPERIMENTAL tests), and what do they buy (A/B benches).
mkdir -p gpurun_out
T=${1:-n1}
(timeout 300 python tools/step_timeline.py --steps 20 2>&1 | tail -60) > gpurun_out/${T}_timeline_pdl1.txt
(timeout 300 python tools/step_timeline.py --steps 20 --pdl 0 2>&1 | tail -60) > gpurun_out/${T}_timeline_pdl0.txt
(DLRM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -x 2>&1 | tail -15) > gpurun_out/${T}_pytest_experimental.log
B="python bench.py --steps 200 --warmup 20 --no-cpu"
(timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_default.log
(DLRM_GROUP_WGRAD=1 timeout 200 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_group_wgrad.log
head -45 gpurun_out/${T}_timeline_pdl1.txt; tail -5 gpurun_out/${T}_pytest_experimental.log
for f in default group_wgrad; do echo "== $f"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_$f.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, "e2e", d.get("e2e",{}).get("value"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_$f.log").read()[-1500:])
PY
done
