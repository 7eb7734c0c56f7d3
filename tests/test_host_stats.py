"""Host-side statistics of the product (motifs.jl_b200/inference.py) against the oracle's literal restatement
(oracle/stats_oracle.py) and scipy.  Float64 results: agreement stated per test.  CPU only."""
import numpy as np
import pytest
from scipy.stats import hypergeom

from motifs_jl_b200 import inference as inf, synth
from oracle import stats_oracle as st


def test_countmat2pfm_and_pwm_bitwise():
    for seed in range(5):
        for cm in synth.random_count_matrices(4, 8, 20, seed):
            assert np.array_equal(inf.countmat2pfm(cm).view(np.uint16), st.countmat2pfm(cm).view(np.uint16))
            bg = np.array([0.3, 0.2, 0.2, 0.3], np.float32)
            a, b = inf.freq2pwm(inf.countmat2pfm(cm), bg), st.freq2pwm(st.countmat2pfm(cm), bg)
            assert np.array_equal(a.view(np.uint16), b.view(np.uint16))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pvalue2score_bitwise(seed):
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(3, 8, 15, seed))
    for p in ms.pwms:
        for pval in (3e-4, 1e-4, 1e-2):
            assert inf.pvalue2score(p, pval) == st.pvalue2score(p, pval)       # same Float64 operations in the same order
            bg = np.array([0.31, 0.19, 0.21, 0.29], np.float32)
            assert inf.pvalue2score(p, pval, bg=bg) == st.pvalue2score(p, pval, bg=bg)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_library_pvalue2score_bitwise(seed):
    """mb200_pvalue2score (csrc/touzet.cu, host arithmetic, no device) against the oracle: identical Float64 bits, incl. -Inf
    entries, a non-uniform background, and the `nothing` case."""
    from motifs_jl_b200 import _lib
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(4, 2, 15, seed))
    bgs = (np.full(4, 0.25, np.float32), np.array([0.31, 0.19, 0.21, 0.29], np.float32))
    for p in ms.pwms:
        for pval in (3e-4, 1e-4, 1e-2, 0.0, 1.0):
            for bg in bgs:
                assert _lib.pvalue2score(None, p, pval, 1e-1, bg) == st.pvalue2score(p, pval, bg=bg)
    p = np.array(ms.pwms[0], np.float16).copy()
    p[2, 1] = -np.inf
    assert _lib.pvalue2score(None, p, 3e-4, 1e-1, bgs[0]) == st.pvalue2score(p, 3e-4, bg=bgs[0])
    assert _lib.pvalue2score(None, p, 0.5, 1e-2, bgs[1]) == st.pvalue2score(p, 0.5, eps=1e-2, bg=bgs[1])
    with pytest.raises(ValueError):
        _lib.pvalue2score(None, p, 1.5, 1e-1, bgs[0])


def test_fisher_right_vs_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        S = int(rng.integers(1000, 2_000_000))
        a = int(rng.integers(0, min(S, 5000)))
        b = int(rng.integers(0, min(S, 5000)))
        p = inf.fisher_right(a, S - a, b, S - b)
        q = float(hypergeom.sf(a - 1, 2 * S, S, a + b))
        assert p == pytest.approx(q, rel=1e-9, abs=1e-300)
    assert inf.fisher_right(0, 10, 0, 10) == 1.0


def test_get_best_thresh_sweep_and_touzet():
    rng = np.random.default_rng(3)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(2, 16, 20, 5))
    pwm = ms.pwms[0]
    fg = {1: (rng.normal(6, 3, 300)).astype(np.float16), 2: (rng.normal(7, 2, 50)).astype(np.float16)}
    bg = {1: (rng.normal(3, 3, 300)).astype(np.float16)}
    fg = {k: v[v > 0] for k, v in fg.items()}
    bg = {k: v[v > 0] for k, v in bg.items()}
    mx, mn = inf.get_max_score(fg, bg), inf.get_min_score(fg, bg)
    # all effective segments >= 15 -> Fisher sweep
    t = inf.get_best_thresh(fg, bg, mx, mn, 16, [range(1, 17)], pwm, 100 * 1000, np.full(4, .25, np.float32))
    t0 = st.get_best_thresh(np.concatenate(list(fg.values())), np.concatenate(list(bg.values())), [(0, 16)], pwm, 100 * 1000, np.full(4, .25, np.float32))
    assert np.float16(t) == np.float16(t0)
    # a short segment -> Touzet sum over segments with 1 < len <= 15
    segs = [range(1, 10), range(12, 13), range(14, 20)]
    t = inf.get_best_thresh(fg, bg, mx, mn, 9, segs, pwm, 100 * 1000, np.full(4, .25, np.float32))
    t0 = st.get_best_thresh(None, None, [(0, 9), (11, 12), (13, 19)], pwm, 100 * 1000, np.full(4, .25, np.float32))
    assert t == t0


def test_union_ranges_mirror():
    from oracle import scan_oracle as so
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(0, 8))
        pos = rng.integers(1, 60, n).tolist()
        ln = int(rng.integers(1, 20))
        got = sum(e - s + 1 for s, e in inf.union_pos(pos, ln))
        assert got == so.union_ranges_total(pos, ln)
