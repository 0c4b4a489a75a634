#!/bin/bash
# N-GPU call (N = $2, default 8): sharded parity with the peer-memory exchange, then the N-GPU bench.
mkdir -p gpurun_out
T=${1:-r16}; N=${2:-8}
export NCCL_DEBUG=WARN
(DLRM_EXCHANGE=p2p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/dist_check.py 2>&1 | grep -v "^$\|frame #" | head -60) > gpurun_out/${T}_n${N}_p2p_check.log
if [ "$(grep -c PASS gpurun_out/${T}_n${N}_p2p_check.log)" -ge 1 ] && ! grep -q FAIL gpurun_out/${T}_n${N}_p2p_check.log; then
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 100 --warmup 10 2>&1 | grep -v "^$\|frame #" | tail -12) > gpurun_out/${T}_bench_n${N}_p2p.log
fi
echo "== check"; cut -c1-400 gpurun_out/${T}_n${N}_p2p_check.log | tail -12
for f in gpurun_out/${T}_bench_n*.log; do echo "== $f"; python - "$f" <<'PY'
import json,sys
txt=open(sys.argv[1]).read()
try:
    l=[x for x in txt.splitlines() if x.startswith("{")][-1]
    d=json.loads(l)
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches","exchange","cuda_graph","clocks")}, "e2e", d.get("e2e",{}).get("value"))
except Exception as e:
    print("ERR", e); print(txt[-1500:])
PY
done
