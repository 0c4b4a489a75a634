#!/bin/bash
mkdir -p gpurun_out
T=${1:-r7}
export NCCL_DEBUG=WARN
(DLRM_EXCHANGE=p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py 2>&1 | grep -v "^$" | tail -60) > gpurun_out/${T}_p2p_check.log
(timeout 400 python bench.py --steps 200 --warmup 20 --gemm tc --no-cpu 2>&1 | tail -2) > gpurun_out/${T}_bench_n1.log
for ex in nccl p2p; do
 (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 100 --warmup 10 --gemm tc --exchange $ex 2>&1 | grep -v "^$" | tail -25) > gpurun_out/${T}_bench_n2_$ex.log
 (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 100 --warmup 10 --gemm tc --exchange $ex --no-graph 2>&1 | grep -v "^$" | tail -25) > gpurun_out/${T}_bench_n2_${ex}_nograph.log
done
cat gpurun_out/${T}_p2p_check.log | cut -c1-300 | tail -40
for f in gpurun_out/${T}_bench_n*.log; do echo "== $f"; python - "$f" <<'PY'
import json,sys
txt=open(sys.argv[1]).read()
try:
    l=[x for x in txt.splitlines() if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches","exchange","cuda_graph")}, "e2e", d.get("e2e",{}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("avg_launch_us"), "upd", (d.get("roofline_update") or {}).get("frac"), (d.get("roofline_update") or {}).get("avg_launch_us"))
except Exception as e:
    print("ERR", e); print(txt[-1800:])
PY
done
