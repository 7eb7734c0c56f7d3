"""The oracle itself: literal greedy_search! form vs table form vs the independent numpy twin, the union_ranges
restatement, and the frozen golden vectors (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from motifs_jl_b200 import synth
from oracle import scan_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_literal_form_equals_table_form():
    a = synth.random_ascii(40, 70, 5)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(6, 8, 30, 6))
    ms.pwms[1][2, 3] = np.float16(-np.inf)       # Inf*0 = NaN in the literal sum
    ms.pwms[2][0, 1] = np.float16(np.nan)
    ms.pwms[3][3, 2] = np.float16(np.inf)
    codes = so.ascii_to_codes(a)
    pw, lens = so.pack_pwms(ms.pwms)
    for rc in (0, 1):
        dense = so.pos_scores(pw, lens, codes, rc).view(np.uint16)           # literal: 4 products per column
        for k in range(6):
            for n in range(40):
                for l in range(70 - int(lens[k]) + 1):
                    t = so.lib().oracle_score_tab(so._p(pw.view(np.uint16)), so._p(lens), 6, k, rc, so._p(codes[n]), l)
                    assert t == dense[k, n, l]
        # findall(pos_scores .> 0) == oracle_scan hits of that strand
        hits, _ = so.scan(pw, lens, codes, strands=1 << rc)
        kk, nn, ll = np.nonzero(dense.view(np.float16) > 0)
        assert len(hits) == len(kk)
        order = np.lexsort((ll, kk, nn))
        assert np.array_equal(hits["seq"], nn[order]) and np.array_equal(hits["motif"], kk[order]) and np.array_equal(hits["pos"], ll[order])
        assert np.array_equal(hits["score_f16"], dense[kk[order], nn[order], ll[order]])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_equals_numpy_twin(seed):
    a = synth.random_ascii(60, 90, seed)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(7, 8, 40, seed + 10))
    codes = so.ascii_to_codes(a)
    pw, lens = so.pack_pwms(ms.pwms)
    for thr in (None, synth.stated_thresholds(ms, 0.4)):
        h, c = so.scan(pw, lens, codes, thr)
        h2, c2 = so.scan_numpy(ms.pwms, codes, thr)
        assert np.array_equal(h, h2) and np.array_equal(c, c2)


def test_union_ranges_quirk():
    # starts {3,10}, len 8 -> 8 (the last interval is never added), not 15
    assert so.union_ranges_total([3, 10], 8) == 8
    assert so.union_ranges_total([3], 8) == 8
    assert so.union_ranges_total([3, 3], 8) == 8            # same start on both strands: dropping one copy changes nothing
    assert so.union_ranges_total([3, 10, 10], 8) == 15
    assert so.union_ranges_total([1, 20, 40], 8) == 16
    assert so.union_ranges_total([40, 1, 20], 8) == 16      # sorted first
    assert so.union_ranges_total([], 8) == 0


def test_golden_config1():
    g = np.load(os.path.join(GOLD, "scan_config1.npz"))
    codes, codes_bg = so.ascii_to_codes(g["train"]), so.ascii_to_codes(g["bg"])
    pw, lens = g["pwms"].view(np.float16), g["lens"]
    h, c = so.scan(pw, lens, codes)
    assert np.array_equal(h, g["hits"]) and np.array_equal(c, g["counts"])
    hb, cb = so.scan(pw, lens, codes_bg)
    assert np.array_equal(hb, g["hits_bg"]) and np.array_equal(cb, g["counts_bg"])
    thr = g["thresh"].view(np.float16)
    fh, fc = so.scan(pw, lens, codes, thr)
    assert np.array_equal(fh, g["filt_hits"]) and np.array_equal(fc, g["filt_counts"])


def test_golden_edge():
    g = np.load(os.path.join(GOLD, "scan_edge.npz"))
    h, c = so.scan(g["pwms"].view(np.float16), g["lens"], so.ascii_to_codes(g["seqs"]))
    assert np.array_equal(h, g["hits"]) and np.array_equal(c, g["counts"])
    assert c[9, 0] == 0                                      # motif longer than the sequence: no position at all


def test_pack_layout():
    a = synth.random_ascii(3, 37, 9)
    w = so.pack_codes(so.ascii_to_codes(a))
    assert w.shape == (3, 3)
    codes = so.ascii_to_codes(a)
    for n in range(3):
        for p in range(37):
            assert (w[n, p // 16] >> (2 * (p % 16))) & 3 == codes[n, p]
