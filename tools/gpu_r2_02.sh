#!/bin/bash
# round 2, call 2: diagnose the chain kernel (ragged case mismatch, graph-mode crash)
T=r2_02
mkdir -p gpurun_out
(timeout 900 python -X faulthandler -m pytest tests/test_gpu_chain.py -q -m gpu 2>&1 | tail -80) > gpurun_out/${T}_chain.log
(timeout 300 python -X faulthandler bench.py --steps 40 --warmup 5 --no-cpu 2>&1 | tail -40) > gpurun_out/${T}_bench_chain.log
(CUDA_LAUNCH_BLOCKING=1 timeout 300 python -X faulthandler bench.py --steps 40 --warmup 5 --no-cpu --no-graph 2>&1 | tail -5 | cut -c1-300) > gpurun_out/${T}_bench_chain_nograph.log
cat gpurun_out/${T}_chain.log | tail -60
cat gpurun_out/${T}_bench_chain.log | cut -c1-600
