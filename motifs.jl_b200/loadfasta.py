"""Host-side mirror of src/loadfasta — a thin caller of the library:

  reading / read_fasta            loadfasta/helpers.jl:83-108   -> mb200_fasta_read   (drop reads with N/n, cap 100 000, keep len == first)
  get_train_test_inds             loadfasta/helpers.jl:141-159  -> mb200_fasta_split  (test size = floor((1-0.9) n): 99 of 1 000)
  get_data_matrices, FASTA_DNA    loadfasta/helpers.jl:206-245, fasta.jl:6-101
                                  -> one upload (mb200_seqs_from_ascii), device gather of the train / test rows (mb200_seqs_gather),
                                     device k-mer-preserving shuffles from a host seed (mb200_seqs_shuffle), device base / transition
                                     counts (mb200_seqs_base_counts)
  get_data_bg                     MOTIFs.jl:35-39               -> mb200_seqs_base_counts

The reference builds four one-hot Float32 arrays on the host (16 B/bp each); here only the ASCII reads cross PCIe once and
everything else happens on the packed 2-bit store in HBM.  SeqShuffle 0.2.2 (`seq_shuffle(k)`) is not vendored in the reference
tree [inferred]: it is taken to preserve the k-mer counts of every read exactly (csrc/fasta.cu); backgrounds are RNG-dependent
in the reference too, so parity tests feed the same background to the oracle and to the kernels.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import Context, Sequences  # noqa: F401

max_num_read_fasta = 100000


def read_fasta(filepath: str, max_entries=max_num_read_fasta):
    """-> list of upper-case reads (helpers.jl:101-108); parsing and filtering run in the library."""
    return [bytes(r).decode() for r in _lib.fasta_read(filepath, max_entries)]


reading = read_fasta          # `reading` differs only by the missing uppercase, which the library always applies


def get_train_test_inds(n, train_test_split_ratio, shuffle, seed):
    """0-based (train, test) index arrays (helpers.jl:141-159), drawn by the library from `seed`."""
    return _lib.fasta_split(n, train_test_split_ratio, shuffle, seed)


def _ascii_matrix(reads):
    return np.frombuffer("".join(reads).encode(), np.uint8).reshape(len(reads), -1).copy() if reads else np.zeros((0, 0), np.uint8)


class FASTA_DNA:
    """FASTA_DNA{Float32}(path) (fasta.jl:61-101).  Attributes follow the reference (N, L, N_test, raw_data, raw_data_test,
    acgt_freq, markov_bg_mat); the four data matrices are `seqs`, `seqs_bg`, `seqs_test`, `seqs_bg_test` (packed, on the GPU).
    Their ASCII host copies `ascii`, `ascii_bg`, `ascii_test`, `ascii_bg_test` are read back lazily (tests, oracle)."""

    def __init__(self, source, ctx: Context, k_train=1, k_test=2, train_test_split_ratio=0.9, shuffle=True,
                 max_entries=max_num_read_fasta, rng: np.random.Generator | None = None):
        rng = rng or np.random.default_rng()
        if isinstance(source, str):
            allrows = _lib.fasta_read(source, max_entries)
        else:
            allrows = _ascii_matrix(list(source))
            allrows = np.where((allrows >= 97) & (allrows <= 122), allrows - 32, allrows).astype(np.uint8)      # uppercase (helpers.jl:107)
        assert len(allrows) != 0, "There aren't DNA strings found in the input"
        seed = int(rng.integers(0, 2 ** 62))                             # the host hands the library ONE seed
        train_idx, test_idx = get_train_test_inds(len(allrows), train_test_split_ratio, shuffle, seed)
        self.ctx = ctx
        everything = ctx.seqs_from_ascii(allrows)                        # the only upload: 1 B/bp
        self.N, self.L, self.N_test = len(train_idx), allrows.shape[1], len(test_idx)
        self.seqs = everything.gather(train_idx)
        self.seqs_bg = self.seqs.shuffle(k_train, seed + 1)
        self.seqs_test = everything.gather(test_idx) if self.N_test else None
        self.seqs_bg_test = self.seqs_test.shuffle(k_test, seed + 2) if self.N_test else None
        everything.free()
        self.train_idx, self.test_idx = train_idx, test_idx
        self._ascii = {"ascii": allrows[train_idx], "ascii_test": allrows[test_idx]}
        cnt, trans = self.seqs_bg.base_counts()                          # est_1st_order_markov_bg(shuffled_dna_read_train), helpers.jl:225
        self.acgt_freq = (cnt / max(cnt.sum(), 1)).astype(np.float32)
        rs = trans.sum(axis=1, keepdims=True)
        self.markov_bg_mat = (trans / np.where(rs == 0, 1, rs)).astype(np.float32)

    @property
    def raw_data(self):
        return [bytes(r).decode() for r in self.ascii]

    @property
    def raw_data_test(self):
        return [bytes(r).decode() for r in self.ascii_test]

    def __getattr__(self, name):
        if name in ("ascii", "ascii_test"):
            return self.__dict__["_ascii"][name]
        if name in ("ascii_bg", "ascii_bg_test"):
            cache = self.__dict__["_ascii"]
            if name not in cache:
                src = self.__dict__["seqs_bg"] if name == "ascii_bg" else self.__dict__["seqs_bg_test"]
                cache[name] = src.to_ascii() if src is not None else np.zeros((0, self.__dict__["L"]), np.uint8)
            return cache[name]
        raise AttributeError(name)

    def free(self):
        for s in (self.seqs, self.seqs_bg, self.seqs_test, self.seqs_bg_test):
            if s is not None:
                s.free()


def est_1st_order_markov_bg(seqs: Sequences):
    """ACGT frequencies and the first-order transition matrix of a (shuffled) read set, Float32 (helpers.jl:225-226)."""
    cnt, trans = seqs.base_counts()
    rs = trans.sum(axis=1, keepdims=True)
    return (cnt / max(cnt.sum(), 1)).astype(np.float32), (trans / np.where(rs == 0, 1, rs)).astype(np.float32)


def get_data_bg(data: FASTA_DNA):
    """this_bg = ACGT frequencies of the training reads, Float32 (MOTIFs.jl:35-39)."""
    cnt, _ = data.seqs.base_counts()
    return (cnt.astype(np.float64) / cnt.sum()).astype(np.float32)


def write_fasta(path: str, ascii_rows: np.ndarray):
    with open(path, "w") as fh:
        for i, r in enumerate(ascii_rows):
            fh.write(f">seq{i}\n{bytes(r).decode()}\n")
