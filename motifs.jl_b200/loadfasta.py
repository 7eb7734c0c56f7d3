"""Host-side mirror of src/loadfasta: FASTA → filtered, equally long, upper-case reads → 90/10
train/test split → k-mer-shuffled backgrounds → four 2-bit packed sequence sets resident in HBM.

  reading / read_fasta            loadfasta/helpers.jl:83-108   (drop reads with N/n, cap 100 000, keep len == first)
  get_train_test_inds             loadfasta/helpers.jl:141-159  (test size = floor((1-0.9) n): 99 of 1 000)
  get_data_matrices, FASTA_DNA    loadfasta/helpers.jl:206-245, fasta.jl:6-101
  get_data_bg                     MOTIFs.jl:35-39

The reference builds four one-hot Float32 arrays on the host (16 B/bp each); here the reads are packed to 2 bit/base
on the GPU (csrc/seqs.cu) and only ASCII stays on the host.  SeqShuffle 0.2.2 (`seq_shuffle(k)`) is not vendored in the
reference tree; the shuffles here are k=1: a permutation of the bases, k=2: a permutation of consecutive 2-mers
[inferred behaviour] — backgrounds are RNG-dependent in the reference too, so parity tests feed the same background to
the oracle and to the kernels.
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import Context, Sequences

max_num_read_fasta = 100000


def reading(filepath: str, max_entries=max_num_read_fasta):
    reads = []
    with open(filepath) as fh:
        txt = fh.read()
    for rec in txt.split(">"):
        if not rec:
            continue
        lines = rec.split("\n")
        this_read = "".join(lines[1:])
        if "N" not in this_read and "n" not in this_read:
            reads.append(this_read)
    if len(reads) > max_entries:
        reads = reads[:max_entries]
    return [s for s in reads if len(s) == len(reads[0])] if reads else reads


def read_fasta(filepath: str, max_entries=max_num_read_fasta):
    return [s.upper() for s in reading(filepath, max_entries)]


def get_train_test_inds(n, train_test_split_ratio, shuffle, rng):
    n_test = int(math.floor((1 - train_test_split_ratio) * n))
    shuffled = rng.permutation(n)
    test = rng.choice(shuffled, n_test, replace=False) if shuffle else np.arange(n - n_test, n)
    tset = set(int(i) for i in test)
    train = np.array([i for i in shuffled if int(i) not in tset], np.int64)
    return train, np.asarray(test, np.int64)


def seq_shuffle(row: np.ndarray, k: int, rng) -> np.ndarray:
    if k <= 1:
        return rng.permutation(row)
    n = len(row) // k
    body = row[: n * k].reshape(n, k)
    return np.concatenate([body[rng.permutation(n)].reshape(-1), row[n * k:]])


def _ascii_matrix(reads):
    return np.frombuffer("".join(reads).encode(), np.uint8).reshape(len(reads), -1).copy() if reads else np.zeros((0, 0), np.uint8)


class FASTA_DNA:
    """FASTA_DNA{Float32}(path) (fasta.jl:61-101).  Attributes follow the reference (N, L, N_test, raw_data, raw_data_test,
    acgt_freq, markov_bg_mat); the four data matrices are `seqs`, `seqs_bg`, `seqs_test`, `seqs_bg_test` (packed, on the GPU)
    with their ASCII host copies in `ascii`, `ascii_bg`, `ascii_test`, `ascii_bg_test`."""

    def __init__(self, source, ctx: Context, k_train=1, k_test=2, train_test_split_ratio=0.9, shuffle=True,
                 max_entries=max_num_read_fasta, rng: np.random.Generator | None = None):
        rng = rng or np.random.default_rng()
        reads = read_fasta(source, max_entries) if isinstance(source, str) else [s.upper() for s in source]
        assert len(reads) != 0, "There aren't DNA strings found in the input"
        train_idx, test_idx = get_train_test_inds(len(reads), train_test_split_ratio, shuffle, rng)
        allrows = _ascii_matrix(reads)
        self.ascii, self.ascii_test = allrows[train_idx], allrows[test_idx]
        self.ascii_bg = np.stack([seq_shuffle(r, k_train, rng) for r in self.ascii]) if len(self.ascii) else self.ascii
        self.ascii_bg_test = np.stack([seq_shuffle(r, k_test, rng) for r in self.ascii_test]) if len(self.ascii_test) else self.ascii_test
        self.N, self.L, self.N_test = len(train_idx), allrows.shape[1], len(test_idx)
        self.raw_data = [reads[i] for i in train_idx]
        self.raw_data_test = [reads[i] for i in test_idx]
        self.acgt_freq, self.markov_bg_mat = est_1st_order_markov_bg(self.ascii_bg)
        self.ctx = ctx
        self.seqs = ctx.seqs_from_ascii(self.ascii)
        self.seqs_bg = ctx.seqs_from_ascii(self.ascii_bg)
        self.seqs_test = ctx.seqs_from_ascii(self.ascii_test) if self.N_test else None
        self.seqs_bg_test = ctx.seqs_from_ascii(self.ascii_bg_test) if self.N_test else None

    def free(self):
        for s in (self.seqs, self.seqs_bg, self.seqs_test, self.seqs_bg_test):
            if s is not None:
                s.free()


_CODE = np.full(256, 255, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def est_1st_order_markov_bg(ascii_rows):
    """ACGT frequencies and the first-order transition matrix of the (shuffled) reads, Float32 (helpers.jl:225-226)."""
    codes = _CODE[ascii_rows]
    freq = np.bincount(codes.ravel(), minlength=4)[:4].astype(np.float64)
    freq = (freq / max(freq.sum(), 1)).astype(np.float32)
    trans = np.zeros((4, 4), np.float64)
    if codes.shape[1] > 1:
        np.add.at(trans, (codes[:, :-1].ravel(), codes[:, 1:].ravel()), 1)
    rs = trans.sum(axis=1, keepdims=True)
    return freq, (trans / np.where(rs == 0, 1, rs)).astype(np.float32)


def get_data_bg(data: FASTA_DNA):
    """this_bg = ACGT frequencies of the training reads, Float32 (MOTIFs.jl:35-39)."""
    codes = _CODE[data.ascii]
    cnt = np.bincount(codes.ravel(), minlength=4)[:4].astype(np.float64)
    return (cnt / cnt.sum()).astype(np.float32)


def write_fasta(path: str, ascii_rows: np.ndarray):
    with open(path, "w") as fh:
        for i, r in enumerate(ascii_rows):
            fh.write(f">seq{i}\n{bytes(r).decode()}\n")
