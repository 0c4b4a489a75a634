"""The committed recipe for tests/golden/*.npz must keep running: regenerate two fixtures from the LIVE
reference (build container only: /root/reference is absent on the GPU box) and compare them bit for bit
with the committed files."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DLRM_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "dlrm_s_pytorch.py")), reason="live reference not present")
def test_make_goldens_regenerates_committed_fixtures(tmp_path):
    env = dict(os.environ, DLRM_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_goldens.py"), "tiny_default", "cfg0_cat"],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("tiny_default", "cfg0_cat"):
        new = np.load(os.path.join(str(tmp_path), name + ".npz"))
        old = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            assert np.array_equal(new[k], old[k]), (name, k)
