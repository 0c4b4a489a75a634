"""The UNMODIFIED reference, importable where /root/reference does not exist (the GPU box).  TEST / BASELINE
INFRASTRUCTURE ONLY -- never imported from dlrm_b200/.

`vendor()` (called by __graft_entry__.build() in the build container) copies the reference's Python files for
this path into oracle/_ref/ -- git-ignored, so no reference source enters the history, but shipped to the GPU
box with the snapshot like the built .so.  `load()` imports dlrm_s_pytorch from there (or from /root/reference
when present) and returns the module; `build_model()` builds a stock `DLRM_Net` through the reference's own
constructor and swaps in quickly-initialised stock `nn.EmbeddingBag` tables (the reference itself reassigns
`emb_l`, dlrm_s_pytorch.py:1319-1326; its numpy init takes minutes at these sizes)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DLRM_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["dlrm_s_pytorch.py", "dlrm_data_pytorch.py", "extend_distributed.py", "mlperf_logger.py", "data_utils.py",
         "data_loader_terabyte.py", "optim/rwsadagrad.py", "tricks/md_embedding_bag.py", "tricks/qr_embedding_bag.py"]


def vendor():
    """Returns the directory holding the reference files, or None when there is no reference to copy."""
    if not os.path.exists(os.path.join(REF_SRC, "dlrm_s_pytorch.py")):
        return REF_DST if os.path.exists(os.path.join(REF_DST, "dlrm_s_pytorch.py")) else None
    for f in FILES:
        dst = os.path.join(REF_DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, f), dst)
    return REF_DST


def location():
    for d in (REF_SRC, REF_DST):
        if os.path.exists(os.path.join(d, "dlrm_s_pytorch.py")):
            return d
    return None


def load():
    """(module dlrm_s_pytorch, directory) or (None, None)."""
    d = location()
    if d is None:
        return None, None
    sys.path.insert(0, d)
    for m in ("dlrm_s_pytorch", "dlrm_data_pytorch", "extend_distributed", "mlperf_logger"):
        sys.modules.pop(m, None)      # never the repo-root drop-in shim of the same name
    import builtins

    import contextlib

    keep_print = builtins.print       # extend_distributed overrides builtins.print on import
    with contextlib.redirect_stdout(sys.stderr):     # "Unable to import onnx" etc. must not land on stdout
        import dlrm_s_pytorch as R

    builtins.print = keep_print
    assert os.path.realpath(os.path.dirname(R.__file__)) == os.path.realpath(d)
    return R, d


def build_model(R, m_spa, ln_emb, ln_bot, ln_top, loss="bce"):
    import numpy as np
    import torch
    import torch.nn as nn

    T = len(ln_emb)
    net = R.DLRM_Net(m_spa, np.array([2] * T), np.array(ln_bot), np.array(ln_top), arch_interaction_op="dot",
                     sigmoid_bot=-1, sigmoid_top=len(ln_top) - 2, ndevices=-1, loss_function=loss)
    emb = nn.ModuleList()
    for n in ln_emb:
        n = int(n)
        a = float(np.sqrt(1.0 / n))
        W = torch.empty(n, m_spa).uniform_(-a, a)
        emb.append(nn.EmbeddingBag(n, m_spa, mode="sum", sparse=True, _weight=W))
    net.emb_l, net.v_W_l = emb, [None] * T
    opt = R.RowWiseSparseAdagrad.RWSAdagrad(net.parameters(), lr=0.01)
    return net, opt
