"""The tensor-core scan (csrc/scan_tc.cuh) is a FILTER: a one-hot GEMM with FP32 accumulation proposes candidates, the sequential
Float16 sum of greedy_search! (_h3_1_alignment.jl:18-36) decides.  Its one obligation is that no true hit is filtered out:

    Float16 running sum of a window > t   ==>   GEMM value D of that window > 0,

where D = (entries of column 0 minus t', rounded UP to Float16) + the remaining entries.  The library exposes the numbers it uses
(mb200_scan_prefilter_bound, host arithmetic only), so the implication can be checked on the CPU: exhaustively over all 4^len windows
for short motifs, and on windows biased towards the threshold and towards large partial sums for long ones.  D is evaluated exactly (float64 holds these sums exactly) and every true hit must clear zero by more than eps32, the slack
the library reserves for the tensor core's FP32 accumulation."""
import itertools

import numpy as np
import pytest

from motifs_jl_b200 import _lib, synth

f16 = np.float16


def seq_f16(vals):
    """left-to-right Float16 running sum of the rows of vals (n_windows, len)"""
    s = np.zeros(vals.shape[0], f16)
    for j in range(vals.shape[1]):
        s = (s.astype(np.float32) + vals[:, j].astype(np.float32)).astype(f16)      # one correctly rounded Float16 add (float32 sum of two halves is exact)
    return s


def check(cols, thr, idx):
    """cols (len, 4) float16, idx (n_windows, len) base indices"""
    E, tp, col0, possible = _lib.scan_prefilter_bound(cols, thr)
    n = cols.shape[0]
    vals = cols[np.arange(n)[None, :], idx]
    s16 = seq_f16(vals)
    t = max(float(f16(thr)), 0.0)
    hit = s16.astype(np.float64) > t
    if not possible:
        assert not hit.any(), "slot disabled although a window exceeds the threshold"
        return 0, 0
    S = vals.astype(np.float64).sum(axis=1)
    # the bound itself
    if hit.any():
        assert np.abs(s16.astype(np.float64) - S)[hit].max() <= E
    # the GEMM value with the threshold folded into column 0 (rounded up), exact arithmetic
    D = col0.astype(np.float64)[idx[:, 0]] + vals[:, 1:].astype(np.float64).sum(axis=1)
    assert (col0.astype(np.float64) >= cols[0].astype(np.float64) - tp).all()          # rounded UP
    slack = (tp - (t - E))                                                             # = -eps32 < 0: what FP32 accumulation may lose
    assert slack < 0
    assert (D[hit] > -slack).all(), "a true hit is closer to the pre-filter threshold than the FP32 slack"
    return int(hit.sum()), int((D > 0).sum())


@pytest.mark.parametrize("seed", range(6))
def test_exhaustive_short_motifs(seed):
    rng = np.random.default_rng(seed)
    for length in (1, 2, 3, 5, 8, 9):
        cm = synth.random_count_matrices(1, length, length, 100 * seed + length)[0]
        pwm = synth.motifs_from_count_matrices([cm]).pwms[0]                       # (4, len) float16
        if seed % 2:
            pwm = (pwm.astype(np.float32) * rng.choice([0.03, 7.0, 40.0])).astype(f16)
        cols = np.ascontiguousarray(pwm.T)
        idx = np.array(list(itertools.product(range(4), repeat=length)), np.int64)
        best = float(cols.astype(np.float32).max(axis=1).sum())
        nh = nc = 0
        for frac in (0.0, 0.3, 0.7, 0.95, 1.0, 1.2):
            h, c = check(cols, f16(frac * best), idx)
            nh += h; nc += c
        assert nc >= nh


@pytest.mark.parametrize("seed", range(4))
def test_sampled_long_motifs_near_the_threshold(seed):
    rng = np.random.default_rng(10 + seed)
    for length in (16, 33, 40, 64):
        cm = synth.random_count_matrices(1, length, length, 7 * seed + length)[0]
        pwm = synth.motifs_from_count_matrices([cm]).pwms[0]
        if seed == 3:
            pwm = (pwm.astype(np.float32) * 16.0).astype(f16)                      # partial sums in the thousands: Float16 spacing 1-4
        cols = np.ascontiguousarray(pwm.T)
        w = cols.astype(np.float64)
        # windows biased to high scores (so that many land around the threshold) mixed with uniform ones
        pr = np.exp2(np.clip(w - w.max(axis=1, keepdims=True), -30, 0) * rng.choice([0.5, 1.0, 2.0]))
        pr /= pr.sum(axis=1, keepdims=True)
        mix = 0.7 * pr + 0.3 * 0.25
        idx = np.stack([rng.choice(4, size=200_000, p=mix[j]) for j in range(length)], axis=1)
        best = float(w.max(axis=1).sum())
        total = 0
        for frac in (0.2, 0.5, 0.7, 0.9):
            h, c = check(cols, f16(frac * best), idx)
            total += h
        assert total > 0


def test_disabled_and_rejected_inputs():
    cols = np.full((5, 4), -1.0, f16)
    E, tp, col0, possible = _lib.scan_prefilter_bound(cols, f16(0.5))
    assert not possible                                                            # best window is negative
    cols[2, 1] = -np.inf
    with pytest.raises(ValueError):
        _lib.scan_prefilter_bound(cols, f16(0.5))                                   # non-finite entries never reach the tensor-core path
