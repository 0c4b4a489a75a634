#!/bin/bash
# Multi-GPU call: GPU tests (incl. facade/CLI/dist), N=1 and N=2.. benches.  Outputs in gpurun_out/.
mkdir -p gpurun_out
T=${1:-r6}
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG" > gpurun_out/${T}_env.txt
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40) > gpurun_out/${T}_pytest.log
(timeout 300 python tools/sweep_gather.py gpurun_out/${T}_sweep.json 2>&1 | tail -8) > gpurun_out/${T}_sweep.log
(timeout 400 python bench.py --steps 200 --warmup 20 --gemm tc --cpu-budget 10 2>&1 | tail -2) > gpurun_out/${T}_bench_n1.log
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    for ex in p2p nccl; do
      (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 100 --warmup 10 --gemm tc --exchange $ex 2>&1 | tail -4) > gpurun_out/${T}_bench_n${n}_$ex.log
    done
  fi
done
tail -15 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_sweep.log
for f in gpurun_out/${T}_bench_n*.log; do echo "== $f"; python - "$f" <<'PY'
import json,sys
txt=open(sys.argv[1]).read()
try:
    l=[x for x in txt.splitlines() if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches","exchange")}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), "upd", (d.get("roofline_update") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(txt[-2500:])
PY
done
