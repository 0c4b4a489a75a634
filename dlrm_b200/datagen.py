"""Input pipeline at the boundary of the hot path: the synthetic
batch sources of the reference, restated for the packed-batch format of dlrm_b200/data.py.

Mirrors, by name and argument meaning, the callers the training loop sees in the reference:

    RandomDataset, collate_wrapper_random_offset/_length, make_random_data_and_loader
        dlrm_data_pytorch.py:571-765 (one __getitem__ = one whole mini-batch, reset of numpy's
        GLOBAL seed on access to batch 0)
    generate_dist_input_batch      :899-958   (uniform | gaussian index distribution)
    generate_synthetic_input_batch :962-1034  (indices replayed from a stack-distance profile)
    generate_random_output_batch   :838-845
    generate_stack_distance, trace_generate_lru/_rand, trace_profile, read/write_dist/trace
        :1037-1208

Every generator draws from numpy's global RandomState in the reference's order, so the same
--numpy-rand-seed yields the same batches bit for bit (tests/test_datagen.py pins that against the
live reference and against committed vectors).  Unlike the reference, the uniform generator does not
loop over samples in Python: it draws one block of doubles per table, walks it once to find each
bag's span, builds all bags of the table with array operations, and then rewinds the global
generator to exactly where the reference's draws would have left it.

`RandomDataset.fill(index, host_batch)` writes a batch straight into the pinned packed buffer that
goes to the GPU with one copy (data.py); `__getitem__` returns the reference's
(X, lS_o, lS_i, T) tuple built from the same arrays.
"""
from __future__ import annotations

import bisect
import sys
from collections import deque
from typing import List, Tuple

import numpy as np
import torch

ra = np.random  # the reference's generators all share numpy's global RandomState

cache_line_size = 1  # dlrm_data_pytorch.py:1055 (granularity of a trace reference)


# --------------------------------------------------------------------------- per-table index draws
def _bag_lengths_fixed(n: int, k: int) -> np.ndarray:
    return np.full(n, int(k), dtype=np.int64)


def _table_uniform(size: int, n: int, lmax: int, fixed: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Bags of one table, uniform ids (generate_dist_input_batch 'uniform' branch, :919-943).

    Reference order of draws per sample: [random(1) unless fixed], random(k);
    ids = unique(round(r * (size - 1))).  Returns (bag sizes after dedup [n], ids flat, bag-major)."""
    size = int(size)
    span = min(size, int(lmax))
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    if fixed:
        k = int(lmax)
        r = ra.random(n * k) if k > 0 else np.zeros(0)
        ks = _bag_lengths_fixed(n, k)
        vals = r
    else:
        # one block covers the worst case; the walk finds how much of it the reference would have used
        state = ra.get_state()
        block = ra.random(n * (1 + max(span, 1)))
        ks = np.empty(n, dtype=np.int64)
        starts = np.empty(n, dtype=np.int64)
        p = 0
        scale = float(span)
        for s in range(n):
            k = int(np.rint(max(1.0, block[p] * scale)))
            ks[s] = k
            starts[s] = p + 1
            p += 1 + k
        ra.set_state(state)
        ra.random(p)                      # leave the global stream where sample-by-sample draws end
        tot = int(ks.sum())
        take = np.repeat(starts - np.concatenate(([0], np.cumsum(ks)[:-1])), ks) + np.arange(tot)
        vals = block[take]
    ids = np.round(vals * (size - 1)).astype(np.int64)
    bag = np.repeat(np.arange(n, dtype=np.int64), ks)
    order = np.lexsort((ids, bag))        # np.unique per bag == sort + drop equal neighbours
    ids, bag = ids[order], bag[order]
    keep = np.ones(ids.size, dtype=bool)
    keep[1:] = (ids[1:] != ids[:-1]) | (bag[1:] != bag[:-1])
    ids, bag = ids[keep], bag[keep]
    return np.bincount(bag, minlength=n).astype(np.int64), ids


def _table_gaussian(size: int, n: int, lmax: int, fixed: bool, lo, hi, mu, sigma) -> Tuple[np.ndarray, np.ndarray]:
    """'gaussian' branch (:931-936): normal(mu, sigma, k) clipped to [lo, hi], unique on the FLOATS,
    then truncated to int64 (so equal ids can repeat inside a bag, and ids ignore the table size --
    both are the reference's behaviour).  normal() keeps a cached second variate, so draws stay
    sample by sample."""
    span = min(int(size), int(lmax))
    sizes = np.empty(n, dtype=np.int64)
    parts: List[np.ndarray] = []
    for s in range(n):
        if fixed:
            k = int(lmax)
        else:
            k = int(np.rint(max(1.0, ra.random(1)[0] * span)))
        r = ra.normal(mu, sigma, k)
        grp = np.unique(np.clip(r, lo, hi)).astype(np.int64)
        sizes[s] = grp.size
        parts.append(grp)
    return sizes, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64))


def _table_trace(table: int, size: int, n: int, lmax: int, fixed: bool, trace_file: str,
                 enable_padding: bool) -> Tuple[np.ndarray, np.ndarray]:
    """generate_synthetic_input_batch (:962-1034): ids replayed from the stack-distance profile in
    `trace_file` with every 'j' replaced by the table number (the reference's file naming)."""
    span = min(int(size), int(lmax))
    sizes = np.empty(n, dtype=np.int64)
    parts: List[np.ndarray] = []
    path = trace_file.replace("j", str(table))
    for s in range(n):
        if fixed:
            k = int(lmax)
        else:
            k = int(max(1, np.round(ra.random(1) * span)[0]))
        line_accesses, list_sd, cumm_sd = read_dist_from_file(path)   # fresh LRU stack per bag (reference)
        refs = trace_generate_lru(line_accesses, list_sd, cumm_sd, k, enable_padding)
        grp = np.unique(np.asarray(list(refs), dtype=np.uint64)).astype(np.int64)
        if grp.min() < 0 or size <= grp.max():
            print("WARNING: distribution is inconsistent with embedding "
                  + "table size (using mod to recover and continue)")
            grp = np.mod(grp, size).astype(np.int64)
        sizes[s] = grp.size
        parts.append(grp)
    return sizes, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64))


def _as_reference_lists(per_table):
    """[(sizes, ids)] -> (lS_emb_offsets, lS_emb_indices): start offset of every bag, ids, int64."""
    lS_o, lS_i = [], []
    for sizes, ids in per_table:
        off = np.zeros(sizes.size, dtype=np.int64)
        if sizes.size:
            np.cumsum(sizes[:-1], out=off[1:])
        lS_o.append(torch.from_numpy(off))
        lS_i.append(torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)))
    return lS_o, lS_i


def _draw_sparse(ln_emb, n, num_indices_per_lookup, fixed, data_generation="random", rand_data_dist="uniform",
                 rand_data_min=1, rand_data_max=1, rand_data_mu=-1, rand_data_sigma=1, trace_file="",
                 enable_padding=False):
    out = []
    for t, size in enumerate(ln_emb):
        if data_generation == "synthetic":
            out.append(_table_trace(t, int(size), n, num_indices_per_lookup, fixed, trace_file, enable_padding))
        elif rand_data_dist == "uniform":
            out.append(_table_uniform(int(size), n, num_indices_per_lookup, fixed))
        elif rand_data_dist == "gaussian":
            if rand_data_mu == -1:        # sticks for the remaining tables, like the reference's local
                rand_data_mu = (rand_data_max + rand_data_min) / 2.0
            out.append(_table_gaussian(int(size), n, num_indices_per_lookup, fixed, rand_data_min,
                                       rand_data_max, rand_data_mu, rand_data_sigma))
        else:
            raise ValueError(str(rand_data_dist) + " distribution is not supported. "
                             "please select uniform or gaussian")
    return out


# --------------------------------------------------------------------------- reference-shaped generators
def generate_dist_input_batch(m_den, ln_emb, n, num_indices_per_lookup, num_indices_per_lookup_fixed,
                              rand_data_dist, rand_data_min, rand_data_max, rand_data_mu, rand_data_sigma):
    """(Xt [n, m_den] fp32, list of T offset tensors [n], list of T index tensors) -- :899-958."""
    Xt = torch.tensor(ra.rand(n, m_den).astype(np.float32))
    per_table = _draw_sparse(ln_emb, n, num_indices_per_lookup, num_indices_per_lookup_fixed, "random",
                             rand_data_dist, rand_data_min, rand_data_max, rand_data_mu, rand_data_sigma)
    lS_o, lS_i = _as_reference_lists(per_table)
    return Xt, lS_o, lS_i


def generate_synthetic_input_batch(m_den, ln_emb, n, num_indices_per_lookup, num_indices_per_lookup_fixed,
                                   trace_file, enable_padding=False):
    """Same tuple, ids from the trace profile (:962-1034)."""
    Xt = torch.tensor(ra.rand(n, m_den).astype(np.float32))
    per_table = _draw_sparse(ln_emb, n, num_indices_per_lookup, num_indices_per_lookup_fixed, "synthetic",
                             trace_file=trace_file, enable_padding=enable_padding)
    lS_o, lS_i = _as_reference_lists(per_table)
    return Xt, lS_o, lS_i


def generate_random_output_batch(n, num_targets, round_targets=False):
    """Click targets in [0, 1), optionally rounded to {0, 1} (:838-845)."""
    P = ra.rand(n, num_targets).astype(np.float32)
    if round_targets:
        P = np.round(P).astype(np.float32)
    return torch.tensor(P)


# --------------------------------------------------------------------------- dataset / loader surface
class RandomDataset(torch.utils.data.Dataset):
    """One item = one mini-batch (dlrm_data_pytorch.py:571-686).  `fill()` is the product path."""

    def __init__(self, m_den, ln_emb, data_size, num_batches, mini_batch_size, num_indices_per_lookup,
                 num_indices_per_lookup_fixed, num_targets=1, round_targets=False, data_generation="random",
                 trace_file="", enable_padding=False, reset_seed_on_access=False, rand_data_dist="uniform",
                 rand_data_min=1, rand_data_max=1, rand_data_mu=-1, rand_data_sigma=1, rand_seed=0):
        nbatches = int(np.ceil((data_size * 1.0) / mini_batch_size))
        if num_batches != 0:
            nbatches = num_batches
            data_size = nbatches * mini_batch_size
        self.m_den, self.ln_emb = m_den, ln_emb
        self.data_size, self.num_batches, self.mini_batch_size = data_size, nbatches, mini_batch_size
        self.num_indices_per_lookup = num_indices_per_lookup
        self.num_indices_per_lookup_fixed = num_indices_per_lookup_fixed
        self.num_targets, self.round_targets = num_targets, round_targets
        self.data_generation, self.trace_file, self.enable_padding = data_generation, trace_file, enable_padding
        self.reset_seed_on_access, self.rand_seed = reset_seed_on_access, rand_seed
        self.rand_data_dist = rand_data_dist
        self.rand_data_min, self.rand_data_max = rand_data_min, rand_data_max
        self.rand_data_mu, self.rand_data_sigma = rand_data_mu, rand_data_sigma
        if data_generation not in ("random", "synthetic"):
            sys.exit("ERROR: --data-generation=" + str(data_generation) + " is not supported")

    def reset_numpy_seed(self, numpy_rand_seed):
        np.random.seed(numpy_rand_seed)

    def __len__(self):
        return self.num_batches      # batches, not samples: __getitem__ yields a whole batch

    def _draw(self, index):
        """(X [n, m_den] f32, [(sizes, ids)] per table, T [n, num_targets] f32) as numpy arrays, drawn
        in the reference's order: dense, tables in order, targets."""
        if self.reset_seed_on_access and index == 0:
            self.reset_numpy_seed(self.rand_seed)
        n = min(self.mini_batch_size, self.data_size - (index * self.mini_batch_size))
        X = ra.rand(n, self.m_den).astype(np.float32)
        per_table = _draw_sparse(self.ln_emb, n, self.num_indices_per_lookup, self.num_indices_per_lookup_fixed,
                                 self.data_generation, self.rand_data_dist, self.rand_data_min,
                                 self.rand_data_max, self.rand_data_mu, self.rand_data_sigma, self.trace_file,
                                 self.enable_padding)
        P = ra.rand(n, self.num_targets).astype(np.float32)
        if self.round_targets:
            P = np.round(P).astype(np.float32)
        return X, per_table, P

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self[i] for i in range(index.start or 0, index.stop or len(self), index.step or 1)]
        X, per_table, P = self._draw(index)
        lS_o, lS_i = _as_reference_lists(per_table)
        return torch.from_numpy(X), lS_o, lS_i, torch.from_numpy(P)

    def fill(self, index, hb):
        """Draw batch `index` straight into a packed pinned HostBatch (data.py): X | target |
        offsets [T, B+1] as global positions | indices, ready for ONE host-to-device copy."""
        X, per_table, P = self._draw(index)
        L = hb.layout
        n = X.shape[0]
        if n != L.B or len(per_table) != L.T or self.num_targets != 1:
            raise ValueError("batch (%d samples, %d tables) does not match the packed layout (%d, %d)"
                             % (n, len(per_table), L.B, L.T))
        total = sum(int(ids.size) for _, ids in per_table)
        if total > L.cap_nnz:
            raise RuntimeError("packed batch capacity %d exceeded (%d indices)" % (L.cap_nnz, total))
        hb.X.numpy()[...] = X
        hb.target.numpy()[...] = P
        idx_all = hb.indices_t.numpy()
        pos = 0
        for k, (sizes, ids) in enumerate(per_table):
            hb.offsets[k, 0] = pos
            hb.offsets[k, 1:] = pos + np.cumsum(sizes)
            idx_all[pos:pos + ids.size] = ids
            pos += int(ids.size)
        hb.nnz = pos
        return hb


def collate_wrapper_random_offset(list_of_tuples):
    """DataLoader(batch_size=1) hands over a 1-element list; offsets become one [T, B] tensor (:689-692)."""
    (X, lS_o, lS_i, T) = list_of_tuples[0]
    return (X, torch.stack(lS_o), lS_i, T)


def offset_to_length_converter(lS_o, lS_i):
    """[T, B] bag lengths from start offsets: the last bag of table k ends at len(lS_i[k])
    (Caffe2-style input; dlrm_data_pytorch.py:384-393)."""
    rows = []
    for k in range(len(lS_o)):
        ends = torch.cat((lS_o[k][1:], torch.tensor([lS_i[k].shape[0]], dtype=lS_o[k].dtype)))
        rows.append((ends - lS_o[k]).int())
    return torch.stack(rows)


def collate_wrapper_random_length(list_of_tuples):
    (X, lS_o, lS_i, T) = list_of_tuples[0]
    return (X, offset_to_length_converter(torch.stack(lS_o), lS_i), lS_i, T)


def make_random_data_and_loader(args, ln_emb, m_den, offset_to_length_converter=False):
    """(train_data, train_loader, test_data, test_loader) from the CLI namespace (:695-765): both
    datasets re-seed numpy with --numpy-rand-seed when batch 0 is read."""
    def dataset():
        return RandomDataset(m_den, ln_emb, args.data_size, args.num_batches, args.mini_batch_size,
                             args.num_indices_per_lookup, args.num_indices_per_lookup_fixed, 1,
                             args.round_targets, args.data_generation, args.data_trace_file,
                             args.data_trace_enable_padding, reset_seed_on_access=True,
                             rand_data_dist=args.rand_data_dist, rand_data_min=args.rand_data_min,
                             rand_data_max=args.rand_data_max, rand_data_mu=args.rand_data_mu,
                             rand_data_sigma=args.rand_data_sigma, rand_seed=args.numpy_rand_seed)

    collate = collate_wrapper_random_length if offset_to_length_converter else collate_wrapper_random_offset
    train_data, test_data = dataset(), dataset()

    def loader(ds):
        return torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=args.num_workers,
                                           collate_fn=collate, pin_memory=False, drop_last=False)

    return train_data, loader(train_data), test_data, loader(test_data)


# --------------------------------------------------------------------------- stack-distance traces
def generate_stack_distance(cumm_val, cumm_dist, max_i, i, enable_padding=False):
    """One stack distance from the cumulative profile (:1037-1051).  While fewer than max_i distinct
    lines have been seen, the support is shrunk to distances that already exist."""
    u = ra.rand(1)
    if i < max_i:
        j = bisect.bisect(cumm_val, i) - 1
        u = u * cumm_dist[j]
    elif enable_padding:
        f0 = cumm_dist[0]
        u = (1.0 - f0) * u + f0
    j = int(np.searchsorted(np.asarray(cumm_dist, dtype=np.float64), float(u[0]), side="left"))
    return cumm_val[j] if j < len(cumm_val) else None


def trace_generate_lru(line_accesses, list_sd, cumm_sd, out_trace_len, enable_padding=False):
    """Replay `out_trace_len` references against an LRU stack (:1058-1083): distance 0 brings the
    oldest line in as a new reference, distance d re-touches the d-th most recent line."""
    max_sd = list_sd[-1]
    depth = len(line_accesses)
    seen = 0
    out = deque()
    for _ in range(out_trace_len):
        sd = generate_stack_distance(list_sd, cumm_sd, max_sd, seen, enable_padding)
        if sd == 0:
            line = line_accesses.pop(0)
            seen += 1
        else:
            line = line_accesses.pop(depth - sd)
        line_accesses.append(line)
        out.append(np.uint64(line * cache_line_size))
    return out


def trace_generate_rand(line_accesses, list_sd, cumm_sd, out_trace_len, enable_padding=False):
    """Variant that does not re-order on a hit (:1086-1106)."""
    max_sd = list_sd[-1]
    depth = len(line_accesses)
    seen = 0
    out = []
    for _ in range(out_trace_len):
        sd = generate_stack_distance(list_sd, cumm_sd, max_sd, seen, enable_padding)
        if sd == 0:
            line = line_accesses.pop(0)
            line_accesses.append(line)
            seen += 1
        else:
            line = line_accesses[depth - sd]
        out.append(np.uint64(line * cache_line_size))
    return out


def trace_profile(trace, enable_padding=False):
    """(LRU stack, stack distances newest-first, first-touch lines newest-first) of a trace (:1109-1158)."""
    stack: deque = deque()
    distances: deque = deque()
    firsts: deque = deque()
    for x in trace:
        line = np.uint64(x / cache_line_size)
        depth = len(stack)
        try:
            at = stack.index(line)
            distances.appendleft(depth - at)
            del stack[at]
        except ValueError:
            distances.appendleft(0)
            firsts.appendleft(line)
        stack.append(line)
    if enable_padding:
        pad = int(np.ceil(len(distances) / max(distances)))
        distances = distances + [0] * pad      # the reference's expression (deque + list), kept as is
    return stack, distances, firsts


def read_dist_from_file(file_path):
    """Three comma-separated lines: unique lines, stack distances, cumulative probabilities (:1186-1198)."""
    with open(file_path, "r") as f:
        lines = f.read().splitlines()
    unique_accesses = [int(el) for el in lines[0].split(", ")]
    list_sd = [int(el) for el in lines[1].split(", ")]
    cumm_sd = [float(el) for el in lines[2].split(", ")]
    return unique_accesses, list_sd, cumm_sd


def write_dist_to_file(file_path, unique_accesses, list_sd, cumm_sd):
    with open(file_path, "w") as f:
        for seq in (list(unique_accesses), list(list_sd), list(cumm_sd)):
            s = str(seq)
            f.write(s[1:len(s) - 1] + "\n")


def read_trace_from_file(file_path, binary=False):
    with open(file_path) as f:
        if binary:
            return np.fromfile(f, dtype=np.uint64).astype(np.uint64).tolist()
        return [np.uint64(x) for x in f.readline().split(", ")]


def write_trace_to_file(file_path, trace, binary=False):
    if binary:
        with open(file_path, "wb+") as f:
            np.array(trace).astype(np.uint64).tofile(f)
    else:
        with open(file_path, "w+") as f:
            s = str(list(trace))
            f.write(s[1:len(s) - 1])
