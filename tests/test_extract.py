"""§8f-3 (triplet enumeration, enriched keys) on the host against the literal oracle; CPU only."""
import numpy as np

from motifs_jl_b200 import _lib, extract, model as mdl
from oracle import extract_oracle as eo


def _fake_codes(seed, nseq=60, per_seq=9, l=82, K=24):
    rng = np.random.default_rng(seed)
    rec = []
    for s in range(nseq):
        if s == 7:
            continue                                   # a sequence without codes: exercises the range quirk
        n = int(rng.integers(3, per_seq))
        e = np.sort(rng.choice(l * 4, n, replace=False))
        for x in e:
            rec.append((x % l if s % 3 else (x % 6) * 5, int(rng.integers(0, 4 if s % 2 else K)), s, np.float16(rng.random()).view(np.uint16), 0))
    a = np.array(rec, _lib.CODE_DTYPE)
    return a[np.lexsort((a["position"], a["fil"], a["seq"]))]       # code_retrieval order: seq, fil, position


def test_triplets_and_enriched_keys_match_literal():
    hp = mdl.Hyperparam()
    for seed in (1, 2):
        codes = _fake_codes(seed)
        cf = extract.filter_code_components_using_quantile(codes, 0.25)
        mags = codes["mag_f16"].view(np.float16).astype(np.float64)
        assert len(cf) == (mags > np.quantile(mags, 0.25)).sum()
        ranges = extract.get_scanning_range_of_filtered_code_components(cf)
        oranges = eo.scanning_ranges(cf["seq"].astype(np.int64) + 1)
        assert [(a + 1, b) for a, b in ranges] == oranges
        H = extract.enumerate_triplets(cf, ranges, hp)
        oH = eo.enumerate_triplets(cf["position"].astype(np.int64) + 1, cf["fil"].astype(np.int64) + 1, oranges, hp.h)
        assert len(H) == len(oH)
        for (pk, v), (ok, ov) in zip(H.items(), oH.items()):          # same keys in the same (insertion) order, same value lists
            u = extract.unpack_key(pk, hp.h)
            assert (u["f1"], u["f2"], u["f3"], u["d12"], u["d13"], u["len"]) == ok
            assert [tuple(r) for r in v.tolist()] == [(a, b) for a, b, _ in ov]
        for mw in (500, 3):
            got = extract.get_enriched_keys(H, max_word_combinations=mw, count_from=4, count_to=1, dec=-1)
            exp = eo.get_enriched_keys(oH, max_word_combinations=mw, count_from=4, count_to=1, dec=-1)
            assert [tuple(extract.unpack_key(k, hp.h)[f] for f in ("f1", "f2", "f3", "d12", "d13", "len")) for k in got] == exp
