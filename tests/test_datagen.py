"""Input pipeline (dlrm_b200/datagen.py) against vectors recorded from the live reference
(oracle/make_goldens_datagen.py) and, when /root/reference is present, against the reference
itself on more seeds.  CPU only; bit-exact (integers) / exact float32 equality."""
import os
import sys

import numpy as np
import pytest
import torch

from dlrm_b200 import datagen as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = os.environ.get("DLRM_REFERENCE", "/root/reference")


def _flat(lS_o, lS_i):
    return (np.concatenate([np.asarray(o).astype(np.int64) for o in lS_o]),
            np.concatenate([np.asarray(i).astype(np.int64) for i in lS_i]),
            np.asarray([int(len(i)) for i in lS_i], dtype=np.int64))


def _same_batch(g, X, lS_o, lS_i):
    o, i, c = _flat([t.numpy() for t in lS_o], [t.numpy() for t in lS_i])
    assert np.array_equal(X.numpy(), g["X"]) and X.dtype == torch.float32
    assert np.array_equal(c, g["counts"])
    assert np.array_equal(o, g["offsets"])
    assert np.array_equal(i, g["indices"])
    assert all(t.dtype == torch.int64 for t in lS_o) and all(t.dtype == torch.int64 for t in lS_i)


@pytest.mark.parametrize("name", ["datagen_uniform_var", "datagen_uniform_fixed", "datagen_uniform_L1",
                                  "datagen_gauss_var", "datagen_gauss_fixed"])
def test_generate_dist_input_batch_matches_reference_vectors(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    np.random.seed(int(g["seed"]))
    X, lS_o, lS_i = G.generate_dist_input_batch(int(g["m_den"]), g["ln_emb"], int(g["n"]), int(g["L"]),
                                                bool(g["fixed"]), str(g["dist"]), float(g["lo"]), float(g["hi"]),
                                                float(g["mu"]), float(g["sigma"]))
    T = G.generate_random_output_batch(int(g["n"]), 1, True)
    tail = np.random.random(3)
    _same_batch(g, X, lS_o, lS_i)
    assert np.array_equal(T.numpy(), g["T"])
    assert np.array_equal(tail, g["tail"])      # the global random stream is left where the reference leaves it


@pytest.mark.parametrize("name", ["datagen_trace_var", "datagen_trace_fixed_pad"])
def test_generate_synthetic_input_batch_matches_reference_vectors(name, monkeypatch):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    monkeypatch.chdir(GOLD)                     # the 'j' -> table number rule applies to the whole path
    np.random.seed(int(g["seed"]))
    X, lS_o, lS_i = G.generate_synthetic_input_batch(int(g["m_den"]), g["ln_emb"], int(g["n"]), int(g["L"]),
                                                     bool(g["fixed"]), "datagen_dist_emb_j.log", bool(g["padding"]))
    tail = np.random.random(3)
    _same_batch(g, X, lS_o, lS_i)
    assert np.array_equal(tail, g["tail"])


@pytest.mark.parametrize("name", ["datagen_dataset", "datagen_dataset_nb"])
def test_random_dataset_epochs_and_packed_fill(name):
    from dlrm_b200.data import HostBatch, PackedLayout

    g = np.load(os.path.join(GOLD, name + ".npz"))
    ln_emb = g["ln_emb"]
    ds = G.RandomDataset(int(g["m_den"]), ln_emb, int(g["data_size"]), int(g["num_batches"]), int(g["mb"]),
                         int(g["L"]), bool(g["fixed"]), 1, True, "random", "", False, reset_seed_on_access=True,
                         rand_data_dist="uniform", rand_data_min=0, rand_data_max=1, rand_data_mu=-1,
                         rand_data_sigma=1, rand_seed=int(g["seed"]))
    assert len(ds) == int(g["nb"])
    np.random.seed(999)
    for e in range(int(g["epochs"])):
        for j in range(len(ds)):
            X, lS_o, lS_i, T = G.collate_wrapper_random_offset([ds[j]])
            k = "e%d_b%d_" % (e, j)
            assert np.array_equal(X.numpy(), g[k + "X"])
            assert np.array_equal(lS_o.numpy(), g[k + "o"]) and lS_o.dtype == torch.int64
            assert np.array_equal(np.concatenate([i.numpy() for i in lS_i]), g[k + "i"])
            assert np.array_equal(T.numpy(), g[k + "T"])
    # product path: the same batches written straight into the packed (pinnable) buffer
    for j in range(len(ds)):
        n = g["e0_b%d_X" % j].shape[0]
        cap = int(n * sum(min(int(r), int(g["L"])) for r in ln_emb))
        hb = HostBatch(PackedLayout(n, len(ln_emb), int(g["m_den"]), cap), pin=False)
        ds.fill(j, hb)
        X, lS_o, lS_i, T = hb.reference_format()
        k = "e0_b%d_" % j
        assert np.array_equal(X.numpy(), g[k + "X"]) and np.array_equal(T.numpy(), g[k + "T"])
        assert np.array_equal(lS_o.numpy(), g[k + "o"])
        assert np.array_equal(np.concatenate([i.numpy() for i in lS_i]), g[k + "i"])
        assert hb.nnz == int(g[k + "c"].sum())
    # length view (Caffe2-style input)
    X, lS_l, lS_i, T = G.collate_wrapper_random_length([ds[0]])
    assert lS_l.dtype == torch.int32
    assert np.array_equal(lS_l.sum(dim=1).numpy(), g["e0_b0_c"])


def test_trace_profile_and_dist_file_roundtrip(tmp_path):
    g = np.load(os.path.join(GOLD, "datagen_profile.npz"))
    stack, sds, firsts = G.trace_profile([np.uint64(x) for x in g["trace"]], False)
    assert np.array_equal(np.asarray(list(stack), dtype=np.uint64), g["stack"])
    assert np.array_equal(np.asarray(list(sds), dtype=np.int64), g["sds"])
    assert np.array_equal(np.asarray(list(firsts), dtype=np.uint64), g["firsts"])
    p = str(tmp_path / "d.log")
    G.write_dist_to_file(p, [4, 5, 6], [0, 1, 2], [0.5, 0.75, 1.0])
    assert G.read_dist_from_file(p) == ([4, 5, 6], [0, 1, 2], [0.5, 0.75, 1.0])
    q = str(tmp_path / "t.log")
    G.write_trace_to_file(q, [3, 1, 2])
    assert [int(x) for x in G.read_trace_from_file(q)] == [3, 1, 2]


@pytest.mark.parametrize("tag", ["A", "B"])
def test_cli_batches_are_the_dataset_batches(tag):
    """With the flags of the pinned CLI loss curves (tests/golden/cli_cfg0_*.flags, test_gpu_facade.py), the
    loader the CLI builds yields exactly the batches of the sample-by-sample restatement
    (cli.reference_order_batch) those curves were recorded with."""
    from dlrm_b200.cli import build_parser, reference_order_batch

    flags = open(os.path.join(GOLD, "cli_cfg0_%s.flags" % tag)).read().split()
    args = build_parser().parse_args(
        ["--arch-sparse-feature-size=16", "--arch-embedding-size=1000-1000-1000",
         "--arch-mlp-bot=13-512-256-64-16", "--arch-mlp-top=512-256-1", "--mini-batch-size=128",
         "--data-generation=random", "--num-batches=6", "--print-freq=1", "--learning-rate=0.1",
         "--numpy-rand-seed=727", "--use-gpu"] + flags)
    ln_emb = np.fromstring(args.arch_embedding_size, dtype=int, sep="-")
    train_data, train_loader, _, _ = G.make_random_data_and_loader(args, ln_emb, 13)
    assert len(train_data) == 6
    np.random.seed(1)                               # model initialisation would have moved the stream
    got = [G.collate_wrapper_random_offset([train_data[j]]) for j in range(6)]
    np.random.seed(args.numpy_rand_seed)            # what the CLI did at batch 0 of the epoch
    want = [reference_order_batch(13, ln_emb, 128, args.num_indices_per_lookup,
                                  args.num_indices_per_lookup_fixed, args.round_targets) for _ in range(6)]
    for a, b in zip(got, want):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[3], b[3])
        assert len(a[2]) == len(b[2]) and all(torch.equal(x, y) for x, y in zip(a[2], b[2]))
    first = next(iter(train_loader))                # the DataLoader path re-seeds at batch 0 as well
    assert torch.equal(first[0], want[0][0]) and torch.equal(first[3], want[0][3])


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present (GPU box)")
def test_against_live_reference_more_seeds(monkeypatch):
    sys.path.insert(0, REF)
    import builtins

    keep_print = builtins.print
    try:
        import dlrm_data_pytorch as dp
    finally:
        builtins.print = keep_print
    rng = np.random.RandomState(2024)
    for trial in range(12):
        T = int(rng.randint(1, 5))
        ln_emb = np.asarray([int(rng.choice([1, 2, 7, 100, 5000, 1000000])) for _ in range(T)])
        n, L = int(rng.randint(1, 40)), int(rng.randint(1, 12))
        fixed = bool(rng.randint(0, 2))
        dist = "uniform" if trial % 3 else "gaussian"
        args = (int(rng.randint(1, 6)), ln_emb, n, L, fixed, dist, 0, 80, -1 if trial % 2 else 33, 9.0)
        np.random.seed(trial)
        a = dp.generate_dist_input_batch(*args)
        ta = np.random.random(2)
        np.random.seed(trial)
        b = G.generate_dist_input_batch(*args)
        tb = np.random.random(2)
        assert torch.equal(a[0], b[0]) and np.array_equal(ta, tb)
        for x, y in zip(a[1] + a[2], b[1] + b[2]):
            assert torch.equal(x, y)
    # trace generators, both replay policies
    monkeypatch.chdir(GOLD)
    for seed in range(4):
        for fn in ("trace_generate_lru", "trace_generate_rand"):
            la, sd, cum = G.read_dist_from_file("datagen_dist_emb_0.log")
            np.random.seed(seed)
            a = list(getattr(dp, fn)(list(la), sd, cum, 25, bool(seed % 2)))
            np.random.seed(seed)
            b = list(getattr(G, fn)(list(la), sd, cum, 25, bool(seed % 2)))
            assert [int(x) for x in a] == [int(x) for x in b]
