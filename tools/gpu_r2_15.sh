#!/bin/bash
# round 2, call 15 (1 GPU): full GPU suite with the new tests; cfg3 with the tiny-table update on a side stream, row pad 4
# vs 16 floats; ncu --set full (+ source counters) of the cfg3 training gather and embedding update
T=r2_15
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
grep "cfg2 full size" gpurun_out/${T}_pytest.log
for pad in 4 16; do
(DLRM_ROW_PAD=$pad timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-check 2>gpurun_out/${T}_cfg3_pad$pad.err | grep '^{' | tail -1) > gpurun_out/${T}_bench_cfg3_pad$pad.json
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_cfg3_pad$pad.json').read()); u=d['roofline_update']; r=d['roofline']
print('pad=$pad step %.4f ms e2e %.4f  gather %.1f us (%.3f)  gather+link %.1f  update %.1f us (%.3f)' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['avg_launch_us'], r['frac'], u['train_gather_plus_link_us'], u['avg_launch_us'], u['frac']))"
done
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:emb_update_lean|emb_fwd_vec" -s 8 -c 4 \
  -o gpurun_out/${T}_cfg3_emb -f python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-check > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log | cut -c1-200
ls -la gpurun_out/${T}_cfg3_emb.ncu-rep
