#!/bin/bash
# round 2, call 5: source-level profile of the chain kernel (why is the epilogue slow?)
T=r2_05
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_chain_kernel -s 9 -c 2 \
  -o gpurun_out/${T}_chain python tools/chain_timeline.py ${T} > gpurun_out/${T}_ncu.log 2>&1
tail -5 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_chain.ncu-rep
