#!/usr/bin/env python
"""Generate tests/golden/datagen_*.npz by running the LIVE reference generators
(/root/reference/dlrm_data_pytorch.py, imported in place; nothing is copied) under fixed numpy
seeds.  TEST INFRASTRUCTURE -- run in the build container only:

    python oracle/make_goldens_datagen.py

Each fixture stores the arguments, the flattened outputs (X, per-table offsets / indices, targets)
and three doubles drawn from numpy's global generator right AFTER the call, which pin the position
the call leaves the random stream in.  The stack-distance profiles used by the 'synthetic' cases are
written by this script (tests/golden/datagen_dist_emb_{0,1}.log), not taken from the reference.
"""
import os
import sys

import numpy as np

REF = os.environ.get("DLRM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, REF)
_print = print

import dlrm_data_pytorch as dp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def flat(lS_o, lS_i):
    return (np.concatenate([o.numpy().astype(np.int64) for o in lS_o]),
            np.concatenate([i.numpy().astype(np.int64) for i in lS_i]),
            np.asarray([int(i.numel()) for i in lS_i], dtype=np.int64))


def write_profiles():
    """Two small stack-distance profiles: (unique lines, distances, cumulative probabilities)."""
    paths = []
    specs = [(list(range(3, 43)), [0, 1, 2, 3, 5, 8, 13, 21, 34], [0.30, 0.42, 0.55, 0.63, 0.74, 0.83, 0.9, 0.96, 1.0]),
             (list(range(0, 12)), [0, 1, 2, 4, 7, 11], [0.5, 0.6, 0.7, 0.85, 0.95, 1.0])]
    for t, (uniq, sd, cum) in enumerate(specs):
        p = os.path.join(OUT, "datagen_dist_emb_%d.log" % t)
        dp.write_dist_to_file(p, uniq, sd, cum)
        paths.append(p)
    return os.path.join(OUT, "datagen_dist_emb_j.log")


def case_dist(name, seed, m_den, ln_emb, n, L, fixed, dist, lo, hi, mu, sigma):
    np.random.seed(seed)
    X, lS_o, lS_i = dp.generate_dist_input_batch(m_den, np.asarray(ln_emb), n, L, fixed, dist, lo, hi, mu, sigma)
    T = dp.generate_random_output_batch(n, 1, True)
    tail = np.random.random(3)
    o, i, cnt = flat(lS_o, lS_i)
    np.savez_compressed(os.path.join(OUT, name), kind="dist", seed=seed, m_den=m_den, ln_emb=np.asarray(ln_emb),
                        n=n, L=L, fixed=int(fixed), dist=dist, lo=lo, hi=hi, mu=mu, sigma=sigma,
                        X=X.numpy(), offsets=o, indices=i, counts=cnt, T=T.numpy(), tail=tail)
    _print(name, "nnz", int(cnt.sum()))


def case_trace(name, seed, m_den, ln_emb, n, L, fixed, trace_file, padding):
    np.random.seed(seed)
    X, lS_o, lS_i = dp.generate_synthetic_input_batch(m_den, np.asarray(ln_emb), n, L, fixed, trace_file, padding)
    tail = np.random.random(3)
    o, i, cnt = flat(lS_o, lS_i)
    np.savez_compressed(os.path.join(OUT, name), kind="trace", seed=seed, m_den=m_den, ln_emb=np.asarray(ln_emb),
                        n=n, L=L, fixed=int(fixed), padding=int(padding), X=X.numpy(), offsets=o, indices=i,
                        counts=cnt, tail=tail)
    _print(name, "nnz", int(cnt.sum()))


def case_dataset(name, seed, m_den, ln_emb, data_size, num_batches, mb, L, fixed, epochs=2):
    ds = dp.RandomDataset(m_den, np.asarray(ln_emb), data_size, num_batches, mb, L, fixed, 1, True, "random", "",
                          False, reset_seed_on_access=True, rand_data_dist="uniform", rand_data_min=0,
                          rand_data_max=1, rand_data_mu=-1, rand_data_sigma=1, rand_seed=seed)
    np.random.seed(999)  # must be overridden by the reset on batch 0
    rec = {}
    for e in range(epochs):
        for j in range(len(ds)):
            X, lS_o, lS_i, T = dp.collate_wrapper_random_offset([ds[j]])
            rec["e%d_b%d_X" % (e, j)] = X.numpy()
            rec["e%d_b%d_o" % (e, j)] = lS_o.numpy()
            rec["e%d_b%d_i" % (e, j)] = np.concatenate([i.numpy() for i in lS_i])
            rec["e%d_b%d_c" % (e, j)] = np.asarray([int(i.numel()) for i in lS_i], dtype=np.int64)
            rec["e%d_b%d_T" % (e, j)] = T.numpy()
    np.savez_compressed(os.path.join(OUT, name), kind="dataset", seed=seed, m_den=m_den, ln_emb=np.asarray(ln_emb),
                        data_size=data_size, num_batches=num_batches, mb=mb, L=L, fixed=int(fixed), epochs=epochs,
                        nb=len(ds), **rec)
    _print(name, "batches", len(ds))


def case_profile(name, seed):
    rng = np.random.RandomState(seed)
    trace = [np.uint64(x) for x in rng.zipf(1.6, 300) % 37]
    stack, sds, firsts = dp.trace_profile(trace, False)
    np.savez_compressed(os.path.join(OUT, name), kind="profile", trace=np.asarray(trace, dtype=np.uint64),
                        stack=np.asarray(list(stack), dtype=np.uint64), sds=np.asarray(list(sds), dtype=np.int64),
                        firsts=np.asarray(list(firsts), dtype=np.uint64))
    _print(name, "distinct", len(firsts))


def main():
    os.makedirs(OUT, exist_ok=True)
    case_dist("datagen_uniform_var.npz", 123, 13, [1000, 37, 5, 1, 100000], 64, 10, False, "uniform", 0, 1, -1, 1)
    case_dist("datagen_uniform_fixed.npz", 7, 4, [50, 3, 2000], 33, 4, True, "uniform", 0, 1, -1, 1)
    case_dist("datagen_uniform_L1.npz", 11, 2, [9, 9], 17, 1, False, "uniform", 0, 1, -1, 1)
    case_dist("datagen_gauss_var.npz", 5, 3, [200, 64], 40, 8, False, "gaussian", 0, 60, -1, 7.5)
    case_dist("datagen_gauss_fixed.npz", 6, 3, [200], 21, 5, True, "gaussian", 2, 90, 40, 20)
    tf = write_profiles()
    case_trace("datagen_trace_var.npz", 123, 5, [64, 16], 24, 6, False, tf, False)
    case_trace("datagen_trace_fixed_pad.npz", 321, 5, [64, 16], 12, 9, True, tf, True)
    case_dataset("datagen_dataset.npz", 727, 13, [1000, 1000, 1000], 300, 0, 128, 10, False)
    case_dataset("datagen_dataset_nb.npz", 4, 6, [30, 7], 1000, 3, 16, 3, True)
    case_profile("datagen_profile.npz", 3)


if __name__ == "__main__":
    main()
