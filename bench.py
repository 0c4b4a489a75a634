#!/usr/bin/env python
"""bench.py -- samples/s of the DLRM fwd+bwd+optimizer hot path on N B200s (contract in the task
statement).  One "step" = one pass of the hot path (forward, loss, backward, fused row-wise-Adagrad
embedding update, dense update) over one synthetic batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1]

Workload at N=1: BASELINE.json configs[2] ("cfg2": 26 tables x 1e6 rows x dim 128, bot 13-512-256-128,
top 479-1024-512-256-1, batch 2048, --data-generation=random distribution, fwd+bwd+RWSAdagrad);
`--workload cfg1` times the forward only (configs[1]).  N>1: the same model with tables sharded
table-wise, per-GPU batch fixed at 2048 (weak scaling).

Timing hygiene: a ring of >= 16 distinct pre-generated batches (their touched rows, 138 MB per batch,
exceed the 126 MB L2 many times over) -> "inputs larger than L2"; CUDA events on the launching
stream; max over ranks; nvidia-smi clocks sampled during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(m_spa=128, rows=1_000_000, T=26, ln_bot=[13, 512, 256, 128], top_tail=[1024, 512, 256, 1],
           B=2048, lmax=10)


def model_dims(T=CFG["T"]):
    D = CFG["m_spa"]
    ln_emb = [CFG["rows"]] * T
    ln_top = [D + (T + 1) * T // 2] + CFG["top_tail"]
    return D, ln_emb, CFG["ln_bot"], ln_top


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.thr, self.gpu = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx_ = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mx_
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample
            sm = [float(l.split(",")[1]) for _, l in self.rows if len(l.split(",")) >= 9] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(train=True, budget_s=20.0, threads=None):
    """The torch-CPU port of the oracle (same ATen ops as the reference's CPU path), timed on the
    host cores on a bounded sample of the same workload."""
    from oracle.torch_cpu_port import CpuDLRM, RowWiseAdagradCPU, time_cpu_steps
    from dlrm_b200.data import make_batch

    if threads:
        torch.set_num_threads(threads)
    D, ln_emb, ln_bot, ln_top = model_dims()
    rows = CFG["rows"]
    note = ""
    try:
        model = CpuDLRM(D, ln_emb, ln_bot, ln_top, loss="bce")
    except (RuntimeError, MemoryError):  # host RAM too small for 13.3 GB of tables
        rows = 100_000
        ln_emb = [rows] * CFG["T"]
        model = CpuDLRM(D, ln_emb, ln_bot, ln_top, loss="bce")
        note = " (rows capped at 1e5: host RAM)"
    opt = RowWiseAdagradCPU(model.parameters(), lr=0.01)
    rng = np.random.default_rng(99)
    batches = [make_batch(rng, ln_emb, CFG["B"], 13, CFG["lmax"], pin=False).reference_format()
               for _ in range(4)]
    batches = [(X, [o for o in lS_o], lS_i, T) for X, lS_o, lS_i, T in batches]
    time_cpu_steps(model, opt, batches, 2, train)  # warm-up
    # "all the host threads it can use": torch's default is one thread per core, which is NOT the
    # fastest setting for this sparse, sync-heavy path -- probe a few counts and keep the best.
    ncpu = os.cpu_count() or 1
    best = (None, float("inf"))
    if not threads:
        for th in sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu}):
            torch.set_num_threads(th)
            time_cpu_steps(model, opt, batches, 1, train)
            tt = time_cpu_steps(model, opt, batches, 2, train) / 2
            if tt < best[1]:
                best = (th, tt)
        torch.set_num_threads(best[0])
    cores = torch.get_num_threads()
    t1 = time_cpu_steps(model, opt, batches, 3, train) / 3
    n = int(max(5, min(200, budget_s / max(t1, 1e-4))))
    dt = time_cpu_steps(model, opt, batches, n, train)
    sps = n * CFG["B"] / dt
    return {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d steps of batch %d, %s, 26x%dx128 tables%s, torch %s CPU, %d threads "
                      "(best of 8/16/32/64/all on a %d-cpu host)" % (
                n, CFG["B"], "fwd+bwd+RWSAdagrad" if train else "fwd only", rows, note,
                torch.__version__, cores, ncpu),
            "ms_per_step": 1e3 * dt / n}


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path (its torch-CPU port: the
    Python reference cannot travel to the GPU box), all host threads, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    train = args.workload != "cfg1"
    per_step_budget = 1.0
    cb = cpu_baseline(train, budget_s=max(5.0, per_step_budget * (args.steps + args.warmup)))
    line = {
        "impl": "reference", "metric": metric_name(train), "value": cb["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic", "config": config_dict(args, 1),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def metric_name(train):
    return "samples/sec (fwd+bwd) MLPerf-DLRM synthetic" if train else "samples/sec (fwd) MLPerf-DLRM synthetic"


def config_dict(args, n):
    return {"workload": "cfg2: 26x1e6x128 tables, bot 13-512-256-128, top 479-1024-512-256-1, "
                        "batch 2048/GPU, random data Lmax=10, fwd+bwd+RWSAdagrad" if args.workload != "cfg1"
            else "cfg1: same model, forward only",
            "global_batch": CFG["B"] * n, "parallelism": "table-wise x%d + dp%d" % (n, n) if n > 1 else "single",
            "l2_policy": "ring of %d distinct batches (inputs larger than L2)" % args.ring,
            "gemm": args.gemm}


# ------------------------------------------------------------------------------ our arm
def bytes_fwd_gather(nnz, T, B, D):
    """rows read + indices + offsets + pooled output written."""
    return nnz * D * 4 + nnz * 8 + T * B * 8 + T * B * D * 4


def ours(args):
    from dlrm_b200.data import DeviceBatch, make_batch
    from dlrm_b200.engine import Engine, GraphedTrainStep

    n = args.gpus
    if n > 1:
        from dlrm_b200 import dist as ddist

        return ddist.bench_main(args, CFG, metric_name, config_dict, ClockSampler)
    torch.cuda.set_device(0)
    dev = "cuda:0"
    train = args.workload != "cfg1"
    D, ln_emb, ln_bot, ln_top = model_dims()
    B, T = CFG["B"], CFG["T"]
    eng = Engine(D, ln_emb, ln_bot, ln_top, loss="bce", sigmoid_top=len(ln_top) - 2, device=dev,
                 max_batch=B, gemm=args.gemm)
    eng.init_params(0)
    eng.ensure_optimizer_state("rwsadagrad")
    rng = np.random.default_rng(1234)
    host = [make_batch(rng, ln_emb, B, 13, CFG["lmax"]) for _ in range(args.ring)]
    devb = []
    for hb in host:
        db = DeviceBatch(hb.layout, dev)
        db.load(hb, non_blocking=False)
        devb.append(db)
    torch.cuda.synchronize()
    lr = 0.01
    # K consecutive steps per CUDA graph (the embedding update of step j overlaps the bottom MLP of step
    # j+1 on a side stream); two sets of K static staging buffers so the H2D of the next K batches can
    # run while the current graph executes.
    Kp = 1
    if train and not args.no_pipeline:
        if args.pipeline >= 1 and args.steps % args.pipeline == 0:
            Kp = args.pipeline
    sets = [[DeviceBatch(host[0].layout, dev) for _ in range(Kp)] for _ in range(2)]
    for st_set in sets:
        for st in st_set:
            st.load(host[0], non_blocking=False)
    use_graph = not args.no_graph
    from dlrm_b200.engine import GraphedTrainSteps

    graphs = None
    if use_graph:
        if train:
            graphs = [GraphedTrainSteps(eng, st_set, lr, "rwsadagrad") for st_set in sets]
        else:
            graphs = [GraphedTrainStep(eng, st_set[0], lr, "rwsadagrad", train=False) for st_set in sets]

    def run_set(si):
        """Kp steps on the batches currently in staging set si; returns a tensor holding a loss / output."""
        if graphs is not None:
            return graphs[si].replay()
        out = None
        for j, st in enumerate(sets[si]):
            if train:
                out = eng.train_step(st.X, st.sparse, st.target, lr, "rwsadagrad", join_update=(j == Kp - 1))
            else:
                out = eng.forward(st.X, st.sparse)
        return out

    def resident_round(i):
        # inputs already resident in HBM: device-to-device copies of the packed batches into the
        # graph's static buffers (2.7 MB each), then Kp steps
        for j, st in enumerate(sets[0]):
            src = devb[(i * Kp + j) % args.ring]
            nbytes = src.layout.used(src.nnz)
            st.buf[:nbytes].copy_(src.buf[:nbytes], non_blocking=True)
        run_set(0)

    rounds, wrounds = args.steps // Kp, max((args.warmup + Kp - 1) // Kp, 1)
    for w in range(wrounds):
        resident_round(w)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    launches0 = eng.n_launch
    ev0.record()
    for r in range(rounds):
        resident_round(wrounds + r)
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    launches = eng.n_launch - launches0
    clocks = sampler.stop(t0, t1)
    ms = ev0.elapsed_time(ev1) / args.steps
    value = B / (ms * 1e-3)

    # ---- e2e: host buffers; H2D of the packed batches + D2H of the loss inside the timed region
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(1).pin_memory()
    main = torch.cuda.current_stream()
    h2d = 0

    def load_set(si, base):
        nonlocal h2d
        for j, st in enumerate(sets[si]):
            h2d += st.load(host[(base + j) % args.ring])

    def e2e_loop(nrounds, base):
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        for f in freed:
            f.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[0])
            load_set(0, base)
            ready[0].record(copy_stream)
        for r in range(nrounds):
            cur, nxt = r & 1, (r + 1) & 1
            if r + 1 < nrounds:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[nxt])
                    load_set(nxt, base + (r + 1) * Kp)
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            out = run_set(cur)
            loss_host.copy_(out.view(-1)[-1:], non_blocking=True)
            freed[cur].record(main)
        main.synchronize()

    e2e_loop(wrounds, 0)
    h2d = 0
    torch.cuda.synchronize()
    ev0.record()
    e2e_loop(rounds, wrounds * Kp)
    ev1.record()
    torch.cuda.synchronize()
    ms_e2e = ev0.elapsed_time(ev1) / args.steps
    e2e = {"value": B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d / args.steps),
           "d2h_bytes_per_step": 4.0 / Kp, "ms_per_step": ms_e2e,
           "note": "packed pinned batches -> one cudaMemcpyAsync each on a copy stream (double-buffered sets of "
                   "%d) -> one CUDA-graph launch per %d steps; loss read back after every launch" % (Kp, Kp)}

    # ---- rooflines of the HBM-bound kernels, timed with CUDA events on the launching stream
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roof = measure_gather_alone(eng, devb, args, hbm_peak, peak_src, T, B, D)
    roof_upd = measure_update_alone(eng, devb, args, hbm_peak, T, B, D) if train else None

    cb = cpu_baseline(train, budget_s=args.cpu_budget) if not args.no_cpu else None
    line = {
        "metric": metric_name(train), "value": value, "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"simt": "fp32", "tc": "fp32 (bf16x3 split on tcgen05, fp32 accumulate)", "tc_bf16": "bf16"}[args.gemm],
        "data": "synthetic", "config": config_dict(args, 1),
        "roofline": roof, "roofline_update": roof_upd, "cpu_baseline": cb, "e2e": e2e,
        "gpu_launches": int(launches), "cuda_graph": bool(use_graph), "steps_per_graph": Kp, "clocks": clocks,
    }
    print(json.dumps(line))


def _time_loop(fn, n, ring):
    """n back-to-back calls between ONE pair of CUDA events on the launching stream (the launch queue
    stays full, so host launch gaps are not counted)."""
    for i in range(3):
        fn(ring[i % len(ring)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(ring[i % len(ring)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n


def measure_update_alone(eng, devb, args, hbm_peak, T, B, D):
    """Training-mode gather (+link) and the fused coalesce + row-wise Adagrad update, timed with CUDA
    events between the launches of back-to-back (gather+link, update) pairs: each kernel is long enough
    (>= 35 us) to hide the host's launch latency of the next one, so the intervals hold no idle gaps."""
    FD = eng.F * eng.D
    out = eng.Tbuf.view(-1)[eng.D:]
    eng.dT.normal_()
    eng.head.zero_()
    n = max(args.steps, 32)
    evs = []
    for i in range(n + 3):
        db = devb[i % len(devb)]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.emb_forward(db.sparse, out, FD, eng.D, link=True)
        e1.record()
        eng.emb_update(db.sparse, eng.dT.view(-1)[eng.D:], FD, eng.D, "rwsadagrad", 1e-6)
        e2.record()
        if i >= 3:
            evs.append((e0, e1, e2))
    torch.cuda.synchronize()
    t_gl = float(np.mean([a.elapsed_time(b) for a, b, _ in evs])) * 1e-3
    tu = float(np.mean([b.elapsed_time(c) for _, b, c in evs])) * 1e-3
    nnz = float(np.mean([d.nnz for d in devb]))
    by_u = nnz * (D * 4 * 2 + 8) + T * B * D * 4 + nnz * 8   # SURVEY 8(d) bytes_bwd
    ach = by_u / tu / 1e9
    return {"kernel": "emb_update_kernel (coalesce + row-wise Adagrad, in place)", "bound": "hbm",
            "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": 337.2e6,
            "traffic_source": "ncu --set full r4: dram__bytes_read.sum 228.7 MB + dram__bytes_write.sum 108.5 MB "
                              "per launch (profiles/r1_ncu_full_emb_interact_head.csv)",
            "avg_launch_us": tu * 1e6, "algorithmic_bytes_per_launch": by_u,
            "train_gather_plus_link_us": t_gl * 1e6,
            "how": "CUDA events between the launches of back-to-back (gather+link, update) pairs"}


def measure_gather_alone(eng, devb, args, hbm_peak, peak_src, T, B, D):
    FD = eng.F * eng.D
    out = eng.Tbuf.view(-1)[eng.D:]
    n = max(args.steps, 32)
    tg = _time_loop(lambda db: eng.emb_forward(db.sparse, out, FD, eng.D), n, devb)
    nnz = float(np.mean([d.nnz for d in devb]))
    by = bytes_fwd_gather(nnz, T, B, D)
    ach = by / tg / 1e9
    return {"kernel": "emb_fwd_vec_kernel (multi-table EmbeddingBag gather)", "bound": "hbm", "achieved": ach,
            "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": 146.2e6,
            "traffic_source": "ncu --set full r2: dram__bytes_read.sum 139.6 MB + dram__bytes_write.sum 6.6 MB per launch "
                              "(profiles/r1_ncu_full_emb_interact_head.csv)",
            "peak_source": peak_src, "avg_launch_us": tg * 1e6, "algorithmic_bytes_per_launch": by,
            "how": "back-to-back launches over the batch ring, one CUDA-event pair on the launching stream"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg1"])
    ap.add_argument("--ring", type=int, default=16)
    ap.add_argument("--gemm", default="tc", choices=["tc", "tc_bf16", "simt"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pipeline", type=int, default=1,
                    help="training steps per CUDA graph (cross-step overlap of the embedding update; measured "
                         "neutral at CFG2 -- the step is bound by the sum of kernel work, not by its critical path)")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1: pooled-vector / gradient exchange fused into the kernels over NVLink peer memory (p2p), or "
                         "NCCL all-to-all; auto = p2p when every GPU pair has peer access, else nccl")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dlrm_b200 has no CPU path (use --impl reference for the CPU arm)")
    ours(args)


if __name__ == "__main__":
    main()
