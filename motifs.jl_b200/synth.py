"""Synthetic inputs of the shapes BASELINE.json names.  numpy Generator(PCG64(seed)) only.

Used by bench.py, the tests and the golden-vector generator; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

from .inference import Motifs, countmat2pfm, freq2pwm, get_high_ic_segments, cmat2ic

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def random_ascii(N, Lb, seed):
    """iid uniform bases, (N, Lb) uint8 ASCII."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return _ACGT[rng.integers(0, 4, size=(N, Lb), dtype=np.uint8)]


def planted_gapped(N, Lb, seed, frac=0.8, half1="TGACGT", half2="ACGTCA", mut=0.05):
    """config 1/2: a gapped motif (two half-sites, spacer U{4,5,6}) planted in `frac` of the sequences at a
    uniform position on a random strand; every half-site base mutates with probability `mut`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seqs = _ACGT[rng.integers(0, 4, size=(N, Lb), dtype=np.uint8)].copy()
    for n in range(N):
        if rng.random() >= frac:
            continue
        sp = int(rng.integers(4, 7))
        spacer = "".join("ACGT"[i] for i in rng.integers(0, 4, size=sp))
        site = list(half1 + spacer + half2)
        for i in list(range(len(half1))) + list(range(len(half1) + sp, len(site))):
            if rng.random() < mut:
                site[i] = "ACGT"[int(rng.integers(0, 4))]
        s = "".join(site)
        if rng.random() < 0.5:
            s = "".join(_COMP[c] for c in reversed(s))
        p = int(rng.integers(0, Lb - len(s) + 1))
        seqs[n, p:p + len(s)] = np.frombuffer(s.encode(), np.uint8)
    return seqs


def shuffle_rows(ascii_rows, seed):
    """1-mer-preserving shuffle of every row (the train background, loadfasta/helpers.jl:217 k=1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.permuted(ascii_rows, axis=1)


def random_count_matrices(K, min_len, max_len, seed, total=1000.0, conc=0.3):
    """config 4: len_k ~ U{min_len..max_len}; every column ~ total * Dirichlet(conc * 1_4), rounded to Float16."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(min_len, max_len + 1, size=K)
    return [(total * rng.dirichlet(np.full(4, conc), size=int(l)).T).astype(np.float16) for l in lens]


def motifs_from_count_matrices(cmats, bg=(0.25, 0.25, 0.25, 0.25)) -> Motifs:
    """count matrix -> pfm -> pwm exactly as the reference (B0), keeping every matrix (no IC filter)."""
    pfms = [countmat2pfm(c) for c in cmats]
    pwms = [freq2pwm(p, np.asarray(bg, np.float32)) for p in pfms]
    segs = [get_high_ic_segments(cmat2ic(c)) for c in cmats]
    return Motifs(pwms=pwms, lens=np.array([p.shape[1] for p in pwms], np.int64), cmats=list(cmats), pfms=pfms,
                  effective_segments=segs,
                  max_effective_lens=np.array([max([len(r) for r in s], default=0) for s in segs], np.int64))


def count_matrix_from_sites(sites):
    """column counts of equally long strings -> (4, len) Float16."""
    L = len(sites[0])
    cm = np.zeros((4, L), np.float32)
    for s in sites:
        for j, c in enumerate(s):
            cm["ACGT".index(c), j] += 1
    return cm.astype(np.float16)


def stated_thresholds(ms: Motifs, frac=0.7):
    """config 4 thresholds: frac * best possible score (Float16), stated in SURVEY §8d."""
    return np.array([np.float16(frac * float(np.asarray(p, np.float32).max(axis=0).sum())) for p in ms.pwms], np.float16)
