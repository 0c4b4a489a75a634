"""Regenerate tests/golden/cli_cfg0_{A,B,C}.txt (and cli_cfg0_C_ref.pt) by running the UNMODIFIED reference command
line (/root/reference/dlrm_s_pytorch.py, CPU) with the flags the GPU tests use (tests/test_gpu_facade.py,
tests/test_gpu_dist.py): the common flags below + tests/golden/cli_cfg0_<tag>.flags.  Test infrastructure; needs the
reference checkout, so it runs in the build container only.

    python oracle/make_cli_goldens.py [--out DIR]        # default: tests/golden

The recorded lines are the ones the tests compare: 'Finished training it ...', 'Testing at - ...', ' accuracy ...',
'Saving model ...' (the checkpoint path inside that line is not compared)."""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DLRM_REFERENCE", "/root/reference")
COMMON = ["--arch-sparse-feature-size=16", "--arch-embedding-size=1000-1000-1000", "--arch-mlp-bot=13-512-256-64-16",
          "--arch-mlp-top=512-256-1", "--mini-batch-size=128", "--data-generation=random", "--num-batches=6",
          "--print-freq=1", "--learning-rate=0.1", "--numpy-rand-seed=727"]
KEEP = re.compile(r"Finished training|Testing at|^ accuracy|Saving model")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--tags", default="A,B,C")
    args = ap.parse_args()
    for tag in args.tags.split(","):
        flags = open(os.path.join(ROOT, "tests", "golden", "cli_cfg0_%s.flags" % tag)).read().split()
        with tempfile.TemporaryDirectory() as tmp:        # the reference writes its TensorBoard run into the cwd
            extra = []
            ck = os.path.join(tmp, "ref.pt")
            if any(f.startswith("--test-freq") for f in flags):
                extra = ["--save-model=" + ck]
            r = subprocess.run([sys.executable, os.path.join(REF, "dlrm_s_pytorch.py")] + COMMON + flags + extra,
                               cwd=tmp, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("reference CLI failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
            lines = [ln for ln in r.stdout.splitlines() if KEEP.search(ln)]
            with open(os.path.join(args.out, "cli_cfg0_%s.txt" % tag), "w") as fh:
                fh.write("\n".join(lines) + "\n")
            if extra:
                os.replace(ck, os.path.join(args.out, "cli_cfg0_%s_ref.pt" % tag))
        print("tag %s: %d lines" % (tag, len(lines)))


if __name__ == "__main__":
    main()
