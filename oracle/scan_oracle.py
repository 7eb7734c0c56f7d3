"""Oracle for the PWM scan path (B0-B8 of SURVEY.md §8a): ctypes front end of oracle/scan_oracle.c plus
an independent numpy twin.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

All paths cited are under the reference's src/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HIT_DTYPE = np.dtype([("seq", "<u4"), ("pos", "<u4"), ("motif", "<u2"), ("score_f16", "<u2"),
                      ("comp", "u1"), ("_pad", "u1", (3,))])
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make oracle`")
        L = C.CDLL(path)
        p, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.oracle_pos_scores.restype = None
        L.oracle_pos_scores.argtypes = [p, p, i32, i32, p, i64, i64, i32, p]
        L.oracle_scan.restype = i64
        L.oracle_scan.argtypes = [p, p, i32, i32, p, p, i64, i64, i32, p, i64, p]
        for f in (L.oracle_score_tab, L.oracle_score_literal):
            f.restype = C.c_uint16
            f.argtypes = [p, p, i32, i32, i32, p, i64]
        L.oracle_set_threads.restype = None
        L.oracle_set_threads.argtypes = [i32]
        L.oracle_max_threads.restype = i32
        L.oracle_max_threads.argtypes = []
        _lib = L
    return _lib


def set_threads(n: int) -> int:
    """OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1); returns what OpenMP will use."""
    lib().oracle_set_threads(int(n))
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_CODE = np.full(256, 255, np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def ascii_to_codes(ascii_rows: np.ndarray) -> np.ndarray:
    """A,C,G,T -> 0..3 (loadfasta/helpers.jl:125-128 row order)."""
    codes = _CODE[np.asarray(ascii_rows, np.uint8)]
    if (codes == 255).any():
        raise ValueError("non-ACGT symbol")
    return np.ascontiguousarray(codes)


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """2 bit/base, 16 bases per little-endian uint32, rows padded to whole words (the library's layout)."""
    N, Lb = codes.shape
    W = (Lb + 15) // 16
    pad = np.zeros((N, W * 16), np.uint32)
    pad[:, :Lb] = codes
    sh = (2 * np.arange(16, dtype=np.uint32))[None, None, :]
    return (pad.reshape(N, W, 16) << sh).sum(axis=2, dtype=np.uint64).astype(np.uint32)


def pack_pwms(pwm_list):
    """get_pos_scores_arr (_h3_1_alignment.jl:65-69, rc=false): pwms[k, :, 1:len_k] = ms.pwms[k], zero padded.
    Returned in Julia memory order (K fastest) as a numpy (maxlen, 4, K) float16 array, plus lens."""
    K = len(pwm_list)
    lens = np.array([p.shape[1] for p in pwm_list], np.int64)
    maxlen = int(lens.max())
    out = np.zeros((maxlen, 4, K), np.float16)
    for k, p in enumerate(pwm_list):
        out[: p.shape[1], :, k] = np.asarray(p, np.float16).T
    return out, lens


def pos_scores(pwms, lens, codes, rc):
    """Dense greedy_search! output [K, N, Lb] Float16 (literal 4-products-per-column form)."""
    pw = np.ascontiguousarray(pwms).view(np.uint16)
    maxlen, _, K = pw.shape
    N, Lb = codes.shape
    out = np.zeros((K, N, Lb), np.uint16)
    lib().oracle_pos_scores(_p(pw), _p(np.ascontiguousarray(lens, np.int64)), K, maxlen, _p(codes), N, Lb, int(rc), _p(out))
    return out.view(np.float16)


def scan(pwms, lens, codes, thresh=None, strands=3, want_hits=True):
    """gpu_scan (+ filter_position_by_best_thresh! when thresh is given) + counts.  Returns (hits, counts)."""
    pw = np.ascontiguousarray(pwms).view(np.uint16)
    maxlen, _, K = pw.shape
    ln = np.ascontiguousarray(lens, np.int64)
    N, Lb = codes.shape
    th = None if thresh is None else np.ascontiguousarray(thresh, np.float16).view(np.uint16)
    counts = np.zeros((K, 4), np.int64)
    if not want_hits:
        lib().oracle_scan(_p(pw), _p(ln), K, maxlen, _p(th), _p(codes), N, Lb, strands, None, 0, _p(counts))
        return None, counts
    n = lib().oracle_scan(_p(pw), _p(ln), K, maxlen, _p(th), _p(codes), N, Lb, strands, None, 0, None)
    hits = np.zeros(n, HIT_DTYPE)
    n2 = lib().oracle_scan(_p(pw), _p(ln), K, maxlen, _p(th), _p(codes), N, Lb, strands, _p(hits), n, _p(counts))
    assert n2 == n
    return hits, counts


# --------------------------------------------------------------------------------------------------
# numpy twin (independent of the C file): same arithmetic through numpy's float16.
# --------------------------------------------------------------------------------------------------
def _eff_table(pwm_f16: np.ndarray) -> np.ndarray:
    """(4, len) -> table whose entry [b, j] is what column j contributes when base b is selected:
    pwm[b, j], or NaN if another row of column j is non-finite (Inf*0 = NaN in greedy_search!)."""
    p = np.asarray(pwm_f16, np.float16)
    nonfin = ~np.isfinite(p)
    others = nonfin.sum(axis=0, keepdims=True) - nonfin
    return np.where(others > 0, np.float16(np.nan), p).astype(np.float16)


def scan_numpy(pwm_list, codes, thresh=None, strands=3):
    """Returns hits sorted by (seq, motif, comp, pos) and counts [K,4], computed with numpy float16 adds."""
    N, Lb = codes.shape
    recs = []
    K = len(pwm_list)
    counts = np.zeros((K, 4), np.int64)
    per = {}
    for k, pwm in enumerate(pwm_list):
        pwm = np.asarray(pwm, np.float16)
        ln = pwm.shape[1]
        npos = Lb - ln + 1
        if npos <= 0:
            continue
        for rc in (0, 1):
            if not (strands >> rc) & 1:
                continue
            tab = _eff_table(pwm[::-1, ::-1] if rc else pwm)     # reverse(pwm): both dims (_h3_1:68-69)
            s = np.zeros((N, npos), np.float16)
            for j in range(ln):
                with np.errstate(invalid="ignore", over="ignore"):
                    s = (s + tab[codes[:, j:j + npos], j]).astype(np.float16)   # one rounded Float16 add per column
            hit = s > np.float16(0)
            if thresh is not None:
                hit &= s > np.float16(thresh[k])
            nn, pp = np.nonzero(hit)
            for n, p_ in zip(nn, pp):
                recs.append((n, k, rc, p_, s[n, p_]))
                per.setdefault((k, int(n)), []).append(int(p_))
    recs.sort(key=lambda r: (r[0], r[1], r[2], r[3]))
    hits = np.zeros(len(recs), HIT_DTYPE)
    for i, (n, k, rc, p_, sc) in enumerate(recs):
        hits[i] = (n, p_, k, np.float16(sc).view(np.uint16), rc, (0, 0, 0))
    for (k, n), pos in per.items():
        ln = np.asarray(pwm_list[k]).shape[1]
        counts[k, 0] += len(pos)
        counts[k, 1] += len(set(pos))                              # get_uniq_pos (_h4_overlap_ratio.jl:5-11)
        counts[k, 2] += union_ranges_total(pos, ln)
        counts[k, 3] += len({q for p_ in pos for q in range(p_, p_ + ln)})
    return hits, counts


def union_ranges_total(positions, ln):
    """union_pos + union_ranges + get_total_occupied_positions (_h4_overlap_ratio.jl:40-79), literal,
    including `for i in eachindex(@view ranges[2:end])` indexing the unsliced array (1-based i = 1..n-1)."""
    ranges = [(p, p + ln - 1) for p in positions]
    if not ranges:
        return 0
    ranges = sorted(ranges, key=lambda r: r[0])
    out = [ranges[0]]
    for i in range(1, len(ranges)):          # i = 1..n-1 (1-based) -> ranges[i] is 0-based ranges[i-1]
        r = ranges[i - 1]
        if out[-1][1] >= r[0]:
            out[-1] = (out[-1][0], r[1])
        else:
            out.append(r)
    return sum(e - s + 1 for s, e in out)
