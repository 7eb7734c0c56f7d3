"""Oracle for the steps between code retrieval and the scans: literal pure-Python restatement of
inference/_2_enumerate.jl:25-65, _3_make_pfms.jl:3-46 and _h6_positions2countmat.jl:39-54.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.  Records are 1-based like the reference's."""
from __future__ import annotations

import numpy as np


def scanning_ranges(seq1):
    """get_scanning_range_of_filtered_code_components (_2_enumerate.jl:25-35); returns 1-based inclusive (start, stop)."""
    cur_seq, cur_start, ranges = 1, 1, []
    for i in range(1, len(seq1) + 1):
        if seq1[i - 1] != cur_seq:
            ranges.append((cur_start, i - 1))
            cur_start = i
            cur_seq += 1
    return ranges


def enumerate_triplets(position1, fil1, ranges, h):
    """enumerate_triplets (_2_enumerate.jl:50-65): H[key] = [(seq_num = range index, pos = position of the first word)]."""
    H = {}
    for ind, (a, b) in enumerate(ranges, start=1):
        items = [(position1[i - 1], fil1[i - 1]) for i in range(a, b + 1)]
        items.sort(key=lambda x: x[0])                                   # stable, by position
        n = len(items)
        for i in range(n - 2):
            for j in range(i + 1, n - 1):
                for k in range(j + 1, n):
                    d12, d13 = items[j][0] - items[i][0], items[k][0] - items[i][0]
                    key = (items[i][1], items[j][1], items[k][1], d12, d13, d13 + h)
                    H.setdefault(key, []).append((ind, items[i][0], False))
    return H


def get_enriched_keys(H, max_word_combinations=500, dec=-5, count_from=200, count_to=10, num_pfms2process=500):
    enriched = None
    for count in range(count_from, count_to - 1, dec):
        enriched = [k for k in H if len(H[k]) > count]
        if len(enriched) > max_word_combinations:
            order = sorted(range(len(enriched)), key=lambda i: -len(H[enriched[i]]))
            return [enriched[i] for i in order[:num_pfms2process]]
    return enriched


def count_matrix(codes, seq1, pos1, comp, length):
    """msa_add! / obtain_count_matrices: add the (4, len) one-hot window, reversed in both dims when comp."""
    cm = np.zeros((4, length), np.float32)
    for s, p, c in zip(seq1, pos1, comp):
        win = codes[s - 1, p - 1: p - 1 + length]
        oh = np.zeros((4, length), np.float32)
        oh[win, np.arange(length)] = 1
        cm += oh[::-1, ::-1] if c else oh
    return cm


def posdicts2countmats(codes, positions, use_comp, lens):
    """posdicts2countmats(ms, data_matrix) (_h6_positions2countmat.jl:26-37): per motif msa_add!(...; return_count_mat=true) =
    Float32 window sums .+ 0.01 (a Float64 literal), converted to Float16 (float_type_retrieval).  positions[m]: {seq1: [pos1...]}."""
    out = []
    for m, L in enumerate(lens):
        seq1 = [n for n, ps in positions[m].items() for _ in ps]
        pos1 = [p for ps in positions[m].values() for p in ps]
        comp = [c for cs in use_comp[m].values() for c in cs]
        out.append((count_matrix(codes, seq1, pos1, comp, int(L)).astype(np.float64) + 0.01).astype(np.float16))
    return out
