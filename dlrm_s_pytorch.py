#!/usr/bin/env python
"""Drop-in entry point: `python dlrm_s_pytorch.py <reference flags>` (what bench/dlrm_s_benchmark.sh
execs from its cwd) runs the dlrm_b200 engine.  See dlrm_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dlrm_b200.cli import run  # noqa: E402

if __name__ == "__main__":
    run()
