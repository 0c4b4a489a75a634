#!/bin/bash
# round 2, call 16 (1 GPU): software-pipelined update kernel: full GPU suite, then every variant timed in one process
T=r2_16
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
grep "cfg2 full size" gpurun_out/${T}_pytest.log
timeout 400 python tools/upd_variants.py cfg3 5,1,4,6,7,3,1 2>gpurun_out/${T}_var_cfg3.err | grep upd_lean | tee gpurun_out/${T}_variants_cfg3.txt
timeout 300 python tools/upd_variants.py cfg2 5,1,4,6,7,3,1 2>gpurun_out/${T}_var_cfg2.err | grep upd_lean | tee gpurun_out/${T}_variants_cfg2.txt
tail -3 gpurun_out/${T}_var_cfg3.err | cut -c1-300
