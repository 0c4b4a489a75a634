"""bench.py host logic that runs without a GPU: the stage watchdog and the ncu-summary reader."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys, time
sys.path.insert(0, %r)
import bench
p = bench.Progress(0)
p.stage("setup")
if %d:
    p.line = {"metric": "m", "value": 123.0, "e2e": None}
    p.extra["parity_check"] = {"ok": True}
p.stage("a stage that never ends")
time.sleep(30)
print("NOT REACHED")
"""


def _run(with_value):
    env = dict(os.environ, DLRM_BENCH_STAGE_LIMIT="1")
    return subprocess.run([sys.executable, "-c", _SCRIPT % (ROOT, with_value)], capture_output=True, text=True, timeout=60,
                          env=env)


def test_watchdog_prints_the_partial_line_and_leaves():
    r = _run(1)
    assert r.returncode == 0 and "NOT REACHED" not in r.stdout
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 123.0 and line["parity_check"] == {"ok": True}
    assert "stalled in stage 'a stage that never ends'" in line["error"]
    assert "WATCHDOG" in r.stderr and "[bench r0" in r.stderr


def test_watchdog_without_a_measurement_exits_nonzero_with_an_error_line():
    r = _run(0)
    assert r.returncode == 3
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] is None and "stalled" in line["error"]


def test_ncu_summary_reader():
    sys.path.insert(0, ROOT)
    import bench

    g = bench.ncu_traffic("emb_fwd_vec_kernel")
    u = bench.ncu_traffic("emb_update_lean_kernel")
    assert g is not None and 0.9e9 < g < 1.3e9          # training gather, cfg3: ~1.08 GB of DRAM traffic per launch
    assert u is not None and 1.5e9 < u < 2.2e9
    assert bench.ncu_traffic("no_such_kernel") is None
