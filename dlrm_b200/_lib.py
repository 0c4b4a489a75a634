"""ctypes binding of include/dlrm_b200.h -- the one and only way Python reaches the kernels.

There is NO fallback: if the shared library is missing (and cannot be built) or a call fails,
a RuntimeError is raised.  Nothing here imports oracle/ or computes on the CPU.
"""
import ctypes as C
import os

from . import _build

MAX_TABLES = 64
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
LOSS_MSE, LOSS_BCE, LOSS_WBCE = 0, 1, 2
OPT_SGD, OPT_RWSADAGRAD = 0, 1
GEMM_SIMT_FP32, GEMM_TC_BF16X3, GEMM_TC_BF16 = 0, 1, 2
TUNE = dict(emb_bags_per_group=0, emb_unroll=1, emb_block=2, upd_block=3, gemm_splitk=4)


class EmbFwdTable(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("indices", C.c_void_p), ("offsets", C.c_void_p),
                ("row_weights", C.c_void_p), ("nnz", C.c_int64), ("rows", C.c_int64)]


class EmbBwdTable(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("momentum", C.c_void_p), ("head", C.c_void_p),
                ("indices", C.c_void_p), ("offsets", C.c_void_p), ("nnz", C.c_int64),
                ("rows", C.c_int64), ("pair_base", C.c_int64)]


_lib = None

# every symbol include/dlrm_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "dlrm_b200_abi_version", "dlrm_b200_last_error", "dlrm_b200_device_info",
    "dlrm_b200_emb_bag_fwd", "dlrm_b200_emb_bwd_link", "dlrm_b200_emb_bwd_update",
    "dlrm_b200_linear_fwd", "dlrm_b200_linear_dgrad", "dlrm_b200_linear_wgrad",
    "dlrm_b200_interact_fwd", "dlrm_b200_interact_bwd", "dlrm_b200_loss_fwd_bwd",
    "dlrm_b200_dense_update",
]


def _declare(lib):
    i64, i32, f32, vp = C.c_int64, C.c_int, C.c_float, C.c_void_p
    lib.dlrm_b200_abi_version.restype = i32
    lib.dlrm_b200_last_error.restype = C.c_char_p
    lib.dlrm_b200_set_tunable.argtypes = [i32, i32]
    lib.dlrm_b200_device_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.dlrm_b200_emb_bag_fwd.argtypes = [C.POINTER(EmbFwdTable), i32, i32, i64, i32, i32, vp, i64, i64, vp]
    lib.dlrm_b200_emb_bwd_link.argtypes = [C.POINTER(EmbBwdTable), i32, i64, i32, i32, vp, vp]
    lib.dlrm_b200_emb_bwd_update.argtypes = [C.POINTER(EmbBwdTable), i32, i32, i64, i32, i32, vp, vp,
                                             i64, i64, i32, f32, f32, vp]
    lib.dlrm_b200_linear_fwd.argtypes = [vp, i64, vp, i64, vp, vp, i64, i64, i64, i64, i32, i32, vp]
    lib.dlrm_b200_linear_dgrad.argtypes = [vp, i64, vp, i64, vp, i64, i32, vp, i64, i64, i64, i64, i32, vp]
    lib.dlrm_b200_linear_wgrad.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, i32, vp]
    lib.dlrm_b200_interact_fwd.argtypes = [vp, i64, vp, i64, i64, i32, i32, i32, vp]
    lib.dlrm_b200_interact_bwd.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, i32, vp]
    lib.dlrm_b200_loss_fwd_bwd.argtypes = [vp, vp, vp, i64, i32, f32, i32, vp, vp, vp, vp]
    lib.dlrm_b200_dense_update.argtypes = [vp, vp, vp, i64, i32, f32, f32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name != "dlrm_b200_last_error":
            fn.restype = i32


def lib():
    """Load (building first if the in-tree .so is stale or absent).  Raises on failure."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not _build.is_fresh():
            if _build.nvcc_path() is not None:
                path = _build.build()
            elif not os.path.exists(path):
                raise RuntimeError(
                    "libdlrm_b200.so is missing and nvcc is unavailable: run "
                    "`python -c 'import __graft_entry__ as g; g.build()'` first. "
                    "dlrm_b200 has no CPU or PyTorch fallback.")
        _lib = C.CDLL(path)
        _declare(_lib)
        if _lib.dlrm_b200_abi_version() != 1:
            raise RuntimeError("libdlrm_b200.so ABI mismatch")
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libdlrm_b200 {what}: " + lib().dlrm_b200_last_error().decode())


def set_tunable(name, value):
    check(lib().dlrm_b200_set_tunable(TUNE[name], int(value)), "set_tunable")


def device_info(device=0):
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(lib().dlrm_b200_device_info(device, C.byref(sm), C.byref(ma), C.byref(mi)), "device_info")
    return sm.value, ma.value, mi.value
