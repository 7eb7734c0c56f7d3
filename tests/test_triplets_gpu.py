"""§8f-3 on the GPU (csrc/triplets.cu through the C ABI): ranges, key counts, first-insertion order, enriched keys and dictionary
values against the host numpy form and the literal oracle (oracle/extract_oracle.py).  Exact (integers)."""
import numpy as np
import pytest

from motifs_jl_b200 import _lib, extract, model as mdl
from oracle import extract_oracle as eo
from test_extract import _fake_codes

pytestmark = pytest.mark.gpu


def _check(ctx, cf, hp, count_to, count_from, dec, mws):
    t = extract.enumerate_triplets_gpu(ctx, cf)
    ranges = extract.get_scanning_range_of_filtered_code_components(cf)
    a, b = t.ranges()
    assert list(zip(a.tolist(), b.tolist())) == ranges
    H = extract.enumerate_triplets(cf, ranges, hp)
    assert t.n_triplets == sum(len(v) for v in H.values())
    # every key with more than `count_to` values: same keys, counts and Dictionary order
    cand = t.frequent(count_to)
    exp = [(k, len(v)) for k, v in H.items() if len(v) > count_to]
    assert list(zip(cand["key"].tolist(), cand["count"].tolist())) == exp
    for mw in mws:
        got = extract.get_enriched_keys_gpu(t, max_word_combinations=mw, count_from=count_from, count_to=count_to, dec=dec)
        ref = extract.get_enriched_keys(H, max_word_combinations=mw, count_from=count_from, count_to=count_to, dec=dec)
        assert got["key"].tolist() == ref
        vals = t.values(got["key"], total=int(got["count"].sum()))
        for k, v in zip(ref, vals):
            assert np.array_equal(v, H[k]), extract.unpack_key(k, hp.h)
    t.free()
    return H


def test_device_dictionary_matches_host_and_literal_oracle(ctx):
    hp = mdl.Hyperparam()
    for seed in (1, 2, 3):
        codes = _fake_codes(seed)
        cf = extract.filter_code_components_using_quantile(codes, 0.25)
        H = _check(ctx, cf, hp, count_to=1, count_from=4, dec=-1, mws=(500, 3))
        oranges = eo.scanning_ranges(cf["seq"].astype(np.int64) + 1)
        oH = eo.enumerate_triplets(cf["position"].astype(np.int64) + 1, cf["fil"].astype(np.int64) + 1, oranges, hp.h)
        assert len(H) == len(oH)


def test_device_dictionary_larger_case_with_repeated_words(ctx):
    """3000 sequences, 12-20 components each, a planted word (three filters at fixed spacing) in a third of them: the planted
    keys pass the reference's default thresholds (count > 10 .. 200) and the table holds ~2 million distinct keys."""
    hp = mdl.Hyperparam()
    rng = np.random.default_rng(5)
    rec = []
    for s in range(3000):
        n = int(rng.integers(12, 21))
        pos = rng.integers(0, 82, n); fil = rng.integers(0, 24, n)
        if s % 3 == 0:
            p0 = int(rng.integers(0, 50))
            pos[:3] = (p0, p0 + 7, p0 + 19); fil[:3] = (2, 11, 5)
        for p, f in zip(pos, fil):
            rec.append((p, f, s, np.float16(rng.random()).view(np.uint16), 0))
    a = np.array(rec, _lib.CODE_DTYPE)
    codes = a[np.lexsort((a["position"], a["fil"], a["seq"]))]
    H = _check(ctx, codes, hp, count_to=extract.cover_at_least, count_from=extract.cover_more_than, dec=-5, mws=(1000, 2))
    planted = extract.pack_key(np.array([3]), np.array([12]), np.array([6]), np.array([7]), np.array([19]))[0]
    assert len(H[int(planted)]) >= 1000


def test_triplets_edge_cases(ctx):
    hp = mdl.Hyperparam()
    empty = np.zeros(0, _lib.CODE_DTYPE)
    t = extract.enumerate_triplets_gpu(ctx, empty)
    assert t.n_ranges == 0 and t.n_triplets == 0 and len(t.frequent(0)) == 0
    t.free()
    # a single sequence: its range is never closed (the reference loses the last sequence)
    one = np.zeros(5, _lib.CODE_DTYPE); one["position"] = [1, 5, 9, 20, 30]; one["fil"] = [0, 1, 2, 3, 4]
    t = extract.enumerate_triplets_gpu(ctx, one)
    assert t.n_ranges == 0 and t.n_triplets == 0
    t.free()
    # ranges with fewer than three components contribute nothing but still count as range indices
    c = np.zeros(9, _lib.CODE_DTYPE)
    c["seq"] = [0, 0, 1, 1, 1, 1, 2, 2, 3]; c["position"] = [3, 4, 9, 2, 2, 7, 1, 2, 3]; c["fil"] = [0, 1, 5, 4, 3, 2, 1, 1, 1]
    _check(ctx, c, hp, count_to=0, count_from=2, dec=-1, mws=(500,))
    # asking for values of a key that is not frequent is an error, not a silent empty answer
    t = extract.enumerate_triplets_gpu(ctx, c)
    t.frequent(0)
    with pytest.raises(_lib.MB200Error):
        t.values(np.array([12345], np.uint64))
    t.free()
    # too many components in one range
    big = np.zeros(300, _lib.CODE_DTYPE); big["seq"][200:] = 1; big["position"] = np.arange(300) % 80
    with pytest.raises(_lib.MB200Error):
        extract.enumerate_triplets_gpu(ctx, big)
