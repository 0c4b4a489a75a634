#!/bin/bash
# round 2, call 1: persistent MLP-chain kernel -- correctness vs per-layer launches, full GPU suite, A/B bench, launch list
T=r2_01
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt
(timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -30) > gpurun_out/${T}_chain.log
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 20 --no-cpu"
(timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain.json
(DLRM_CHAIN=0 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_nochain.json
(DLRM_CHAIN_TILE_N=64 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain_tn64.json
(DLRM_CHAIN_TILE_N=128 timeout 300 $B 2>&1 | tail -1) > gpurun_out/${T}_bench_chain_tn128.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/${T}_ncu_bench.log 2>&1
tail -3 gpurun_out/${T}_chain.log gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_bench_*.json | cut -c1-400
