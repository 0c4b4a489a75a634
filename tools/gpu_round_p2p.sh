#!/bin/bash
# 2-GPU call: IPC probe, sharded parity (p2p exchange, own barrier/all-reduce), N=2 benches.
mkdir -p gpurun_out
T=${1:-r9}
export NCCL_DEBUG=WARN
(CUDA_LAUNCH_BLOCKING=1 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/p2p_probe.py 2>&1 | grep -v "^$\|frame #" | head -80) > gpurun_out/${T}_probe.log
(DLRM_EXCHANGE=p2p CUDA_LAUNCH_BLOCKING=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tests/dist_check.py 2>&1 | grep -v "^$\|frame #" | head -60) > gpurun_out/${T}_p2p_check_blocking.log
if grep -q PASS gpurun_out/${T}_p2p_check_blocking.log; then
(DLRM_EXCHANGE=p2p timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/dist_check.py 2>&1 | grep -v "^$\|frame #" | head -40) > gpurun_out/${T}_p2p_check.log
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 100 --warmup 10 --gemm tc --exchange p2p 2>&1 | grep -v "^$\|frame #" | tail -12) > gpurun_out/${T}_bench_n2_p2p.log
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 2 --steps 100 --warmup 10 --gemm tc --exchange p2p --no-graph 2>&1 | grep -v "^$\|frame #" | tail -12) > gpurun_out/${T}_bench_n2_p2p_nograph.log
fi
echo "== probe"; cut -c1-260 gpurun_out/${T}_probe.log | head -40
echo "== check blocking"; cut -c1-260 gpurun_out/${T}_p2p_check_blocking.log | head -30
echo "== check"; cut -c1-260 gpurun_out/${T}_p2p_check.log | tail -6; 
for f in gpurun_out/${T}_bench_n*.log; do echo "== $f"; python - "$f" <<'PY'
import json,sys
txt=open(sys.argv[1]).read()
try:
    l=[x for x in txt.splitlines() if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches","exchange","cuda_graph")}, "e2e", d.get("e2e",{}).get("value"))
except Exception as e:
    print("ERR", e); print(txt[-1200:])
PY
done
