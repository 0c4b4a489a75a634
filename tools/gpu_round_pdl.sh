#!/bin/bash
# 1-GPU A/B of programmatic dependent launch on the dense chain + full parity run with it enabled.
mkdir -p gpurun_out
T=${1:-r20}
B="python bench.py --steps 200 --warmup 20 --no-cpu"
(DLRM_TUNE="pdl=2" timeout 150 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_pdl_off.log
(DLRM_TUNE="pdl=1" timeout 150 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_pdl_on.log
(DLRM_TUNE="pdl=2" timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${T}_smoke_pdl_off.log
(DLRM_TUNE="pdl=1" timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -12) > gpurun_out/${T}_pytest_pdl_on.log
tail -3 gpurun_out/${T}_smoke_pdl_off.log; tail -4 gpurun_out/${T}_pytest_pdl_on.log
for f in pdl_off pdl_on; do echo "== $f"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/${T}_bench_$f.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${T}_bench_$f.log").read()[-1500:])
PY
done
