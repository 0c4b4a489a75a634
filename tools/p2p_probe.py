"""Diagnostic: does CUDA-IPC peer mapping work between two torchrun ranks on this box, with torch's
own copy kernels and with our P2P gather?  torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import ctypes as C
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def log(*a):
    print("[rank %s]" % os.environ.get("RANK"), *a, flush=True)


def main():
    from dlrm_b200.dist import init_distributed

    rank, world = init_distributed("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    log("device", dev, torch.cuda.get_device_name(dev), "can_access_peer",
        [torch.cuda.can_device_access_peer(dev.index, j) for j in range(torch.cuda.device_count()) if j != dev.index])
    from torch.multiprocessing.reductions import reduce_tensor

    t = torch.zeros(4096, device=dev)
    torch.cuda.synchronize()
    allh = [None] * world
    dist.all_gather_object(allh, reduce_tensor(t))
    peers = []
    for r in range(world):
        if r == rank:
            peers.append(t)
        else:
            fn, args = allh[r]
            log("rebuilding peer", r, "args device", args[6] if len(args) > 6 else None)
            peers.append(fn(*args))
    for r in range(world):
        log("peer", r, "tensor device", peers[r].device, "ptr", hex(peers[r].data_ptr()))
    dist.barrier()
    try:
        for r in range(world):
            if r != rank:
                peers[r][rank * 8:(rank + 1) * 8].fill_(float(rank + 1))   # torch kernel writing peer memory
        torch.cuda.synchronize()
        dist.barrier()
        log("after torch peer writes, my buffer head:", t[:16].tolist())
    except Exception:
        log("torch peer write FAILED"); traceback.print_exc()
    # raw kernel from our library writing through a mapping opened by US with our device current
    try:
        from dlrm_b200 import _lib

        lib = _lib.lib()
        D, B, R = 128, 64, 1000
        W = torch.randn(R, D, device=dev)
        idx = torch.randint(0, R, (B * world * 3,), device=dev)
        off = torch.arange(0, B * world * 3, 3, device=dev)
        backing = torch.zeros(8 << 20, device=dev)      # 32 MB: a cudaMalloc segment of its own
        out = backing[4096:4096 + B * 2 * D].view(B, 2, D)
        torch.cuda.synchronize()
        hb, ob = C.create_string_buffer(64), C.c_int64()
        _lib.check(lib.dlrm_b200_ipc_export(out.data_ptr(), hb, C.byref(ob)), "ipc_export")
        mine = (hb.raw, int(ob.value))
        allo = [None] * world
        dist.all_gather_object(allo, mine)
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(out.data_ptr() + rank * D * 4)
            else:
                base = C.c_void_p()
                _lib.check(lib.dlrm_b200_ipc_open(allo[r][0], dev.index, C.byref(base)), "ipc_open")
                log("opened peer", r, "base", hex(base.value), "offset", allo[r][1])
                ptrs.append(base.value + allo[r][1] + rank * D * 4)
        arr = (C.c_void_p * world)(*ptrs)
        desc = (_lib.EmbFwdTable * 1)()
        desc[0].weight, desc[0].indices, desc[0].offsets = W.data_ptr(), idx.data_ptr(), off.data_ptr()
        desc[0].row_weights, desc[0].nnz, desc[0].rows = None, idx.numel(), R
        dist.barrier()
        rc = lib.dlrm_b200_emb_bag_fwd_p2p(desc, None, 1, D, B * world, 8, 0, None, arr, world, B, 2 * D, D, None,
                                           torch.cuda.current_stream().cuda_stream)
        log("emb_bag_fwd_p2p rc", rc, lib.dlrm_b200_last_error())
        torch.cuda.synchronize()
        dist.barrier()
        log("my out |sum| per feature slot (both must be non-zero):",
            [float(out[:, s, :].abs().sum().item()) for s in range(2)])
    except Exception:
        log("p2p kernel FAILED"); traceback.print_exc()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
