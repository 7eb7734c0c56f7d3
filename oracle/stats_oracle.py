"""Oracle for the host-side statistics around the scan: count matrix -> PWM (B0), Touzet p-value ->
score (B4), threshold choice (B4), threshold filter (B5), Fisher exact test (B8).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.  Paths are under the reference's src/.

Third-party arithmetic restated here (sources not in /root/reference, versions from Manifest.toml):
  HypothesisTests 0.10.13 FisherExactTest/pvalue(tail=:right) -> Distributions Hypergeometric ccdf:
      P[X >= a], X ~ Hypergeometric(successes a+b, failures c+d, draws a+c)  (published definition);
      restated with scipy.stats.hypergeom.sf (agreement ~1e-12 relative, not bitwise).
  DataStructures 0.18.15 SortedDict: an ordered map; restated as dict + sorted key iteration.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.stats import hypergeom

f16 = np.float16
f32 = np.float32


# ---- B0: inference/_s1_make_motifs.jl:68-76 -----------------------------------------------------
def countmat2pfm(count_matrix):
    """countmat2pfm: Float16((cnt + ps) / (colsum + 4ps)), ps = Float16(0.01); count matrices reaching
    countmats2motifs are Float16 (_h6_positions2countmat.jl:24,36,54), so every op is a Float16 op."""
    cm = np.asarray(count_matrix, f16)
    ps = f16(0.01)
    colsum = np.zeros((1, cm.shape[1]), f16)
    for a in range(4):                       # sum(count_matrix, dims=1): sequential Float16 adds down the column
        colsum = (colsum + cm[a:a + 1]).astype(f16)
    four_ps = f16(f16(4) * ps)
    with np.errstate(all="ignore"):
        return ((cm + ps).astype(f16) / (colsum + four_ps).astype(f16)).astype(f16)


def freq2pwm(pfm, bg):
    """freq2pwm(pfm, bg) = log2.(pfm ./ bg) with bg Float32 (MOTIFs.jl:35-39) -> Float32, then stored into
    Matrix{Float16} by the motifs constructor (_s1_make_motifs.jl:107-125)."""
    with np.errstate(all="ignore"):
        q = np.asarray(pfm, f16).astype(f32) / np.asarray(bg, f32).reshape(4, 1)
        return np.log2(q.astype(f32)).astype(f32).astype(f16)


# ---- Touzet: inference/_h2_Touzet.jl ------------------------------------------------------------
def _best(pwm):
    return float(sum(max(pwm[:, i]) for i in range(pwm.shape[1])))


def _worst(pwm):
    return float(sum(min(pwm[:, i]) for i in range(pwm.shape[1])))


def pvalue2score(pwm, pval, eps=1e-1, bg=(0.25, 0.25, 0.25, 0.25)):
    """pvalue2score (_h2_Touzet.jl:170-187): Float64 throughout."""
    pwm64 = np.asarray(pwm, np.float64)
    bg64 = [float(np.float64(b)) for b in bg]
    m = pwm64.shape[1]
    # min_score_range (:35-37): columns by best-worst, descending (stable)
    delta = np.array([max(pwm64[:, i]) - min(pwm64[:, i]) for i in range(m)])
    perm = np.argsort(-delta, kind="stable")
    mp = pwm64[:, perm]
    pe = np.floor(mp / eps) * eps                                   # round_pwm (:46)
    alpha, beta = _worst(pe), math.inf
    Q = [dict() for _ in range(m + 1)]
    Q[0][0.0] = 1.0                                                 # create_Q (:69-70)
    for i in range(1, m + 1):                                       # score_distribution (:105-115)
        rest = pe[:, i:m]
        bs = 0.0 if i + 1 > m else _best(rest)
        ws = 0.0 if i + 1 > m else _worst(rest)
        for score in sorted(Q[i - 1].keys()):                       # SortedDict iteration order
            for j in range(4):
                t = score + pe[j, i - 1]
                if alpha - bs <= t <= beta - ws:
                    Q[i][t] = Q[i].get(t, 0.0) + Q[i - 1][score] * bg64[j]
    Qm = Q[m]
    q_sum = 0.0
    for k in sorted(Qm.keys()):                                     # Q_sum: sum(values) in key order
        q_sum += Qm[k]
    largest = None
    for k in sorted(Qm.keys()):                                     # find_largest_alpha (:120-132)
        if q_sum >= pval:
            largest = k
        else:
            return k
        q_sum -= Qm[k]
    return largest


def get_pvalue(pwm):
    """_s2_filter_pos_w_scores.jl:38-43 with _0_const.jl:63-66."""
    n = pwm.shape[1]
    if 9 < n <= 11:
        return 0.0001
    if n <= 9:
        return 0.0003
    return 0.0001


# ---- Fisher: inference/_h7_fisher.jl:21-36 ------------------------------------------------------
def fisher_right(a, c, b, d):
    """pvalue(FisherExactTest(a, c, b, d), tail=:right) with the reference's argument order (a, c, b, d)."""
    a, c, b, d = int(a), int(c), int(b), int(d)
    return float(hypergeom.sf(a - 1, a + c + b + d, a + c, a + b))


def fisher_pvec(activate_counts, activate_counts_bg, N, L):
    out = []
    S = N * L
    for a, b in zip(activate_counts, activate_counts_bg):
        if a == 0 and b == 0:
            out.append(1.0)
        else:
            out.append(fisher_right(a, S - a, b, S - b))
    return np.array(out, np.float64)


# ---- thresholds: inference/_s2_filter_pos_w_scores.jl:11-36, 90-114 ------------------------------
def get_best_thresh(scores, bg_scores, eff_pos, pwm, asum, bg):
    """scores / bg_scores: flat float16 arrays of all (unfiltered) hit scores of this motif.
    eff_pos: list of 0-based (start, stop_exclusive) effective segments.  Returns a Python float
    (the caller stores Float16(result), :132-134)."""
    if any((e - s) < 15 for s, e in eff_pos):
        best = f32(0)
        best = float(best)
        for s, e in eff_pos:
            if (e - s) > 15 or (e - s) <= 1:
                continue
            sub = np.asarray(pwm, f16)[:, s:e]
            best += pvalue2score(sub, get_pvalue(sub), bg=bg)
        return best
    sc = np.asarray(scores, f16)
    bsc = np.asarray(bg_scores, f16)
    both = np.concatenate([sc, bsc])
    min_score = f16(np.inf) if both.size == 0 else both.min()       # get_min_score (:24-36)
    max_score = f16(-np.inf) if both.size == 0 else both.max()      # get_max_score (:11-23)
    best_thresh = min_score
    t = min_score
    best_p = f32(1)
    while t < max_score:
        a = int((sc > t).sum())
        b = int((bsc > t).sum())
        p = fisher_right(a, asum - a, b, asum - b)
        if p < best_p:
            best_p = p
            best_thresh = t
        t = f16(t + f16(0.5))                                        # score_thresh_increment (_0_const.jl:29)
    return float(best_thresh)
