"""Host-side mirror of what sits between code retrieval and the scans (SURVEY §8f-3, §8f-4 — the "next" rows):

  filter_code_components_using_quantile!, get_scanning_range_of_filtered_code_components,
  enumerate_triplets                          inference/_2_enumerate.jl:5-65
  get_enriched_keys, get_words                inference/_3_make_pfms.jl:3-26
  obtain_count_matrices                       inference/_3_make_pfms.jl:28-46      -> GPU (mb200_count_matrices)
  enriched_keys2motifs                        inference/_s1_make_motifs.jl:234-259 -> GPU counts + countmats2motifs
  posdicts2countmats                          inference/_h6_positions2countmat.jl:26-37 -> GPU
  run_thru (the in-scope stages)              inference/_g1_obtain_coutmats.jl:131-173

The triplet enumeration and key counting run on the GPU (csrc/triplets.cu: hash table of packed 64-bit keys, mb200_triplets_*);
run_thru uses that path.  `enumerate_triplets` / `get_enriched_keys` below are the same stage as plain numpy on explicit
dictionaries: the form the tests compare the GPU path and the literal oracle with, not what run_thru executes.  Count-matrix
accumulation runs on the GPU.  Everything is 1-based like the reference at this level.
OUT OF SCOPE and therefore NOT applied by run_thru here: merge_H / trim_H, expansions_ms!,
alignment_merge!, merge_to_remove_redundancy!, trim_cmats, merge_count_matrices — small sequential CPU heuristics.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np

from . import _lib
from .inference import Motifs, countmats2motifs

# inference/_0_const.jl
num_pfms2process = 500
cover_more_than, cover_at_least = 200, 10

_TRIPLES = {}


def _triples(n):
    if n not in _TRIPLES:
        _TRIPLES[n] = np.array(list(combinations(range(n), 3)), np.int64).reshape(-1, 3)
    return _TRIPLES[n]


def filter_code_components_using_quantile(codes, p):
    """keep components with mag > quantile(mags, p) (Float64 quantile of the Float16 magnitudes, _2_enumerate.jl:10-13)."""
    mags = codes["mag_f16"].view(np.float16).astype(np.float64)
    if len(mags) == 0:
        return codes
    return codes[mags > np.quantile(mags, p)]


def get_scanning_range_of_filtered_code_components(codes):
    """_2_enumerate.jl:25-35, literal (1-based sequence ids): a new range is closed every time `seq != cur_seq`, cur_seq only
    counts up by one, and the range of the last sequence is never closed.  Returns 0-based half-open (start, stop) pairs."""
    ranges, cur_seq, start = [], 1, 0
    seq1 = codes["seq"].astype(np.int64) + 1
    for i in range(len(seq1)):
        if seq1[i] != cur_seq:
            ranges.append((start, i))
            start = i
            cur_seq += 1
    return ranges


def pack_key(f1, f2, f3, d12, d13):
    return (f1.astype(np.uint64) | (f2.astype(np.uint64) << np.uint64(8)) | (f3.astype(np.uint64) << np.uint64(16))
            | (d12.astype(np.uint64) << np.uint64(24)) | (d13.astype(np.uint64) << np.uint64(40)))


def unpack_key(key, h):
    key = int(key)
    f1, f2, f3 = key & 0xFF, (key >> 8) & 0xFF, (key >> 16) & 0xFF
    d12, d13 = (key >> 24) & 0xFFFF, (key >> 40) & 0xFFFF
    return dict(f1=f1, f2=f2, f3=f3, d12=d12, d13=d13, len=d13 + h)


def enumerate_triplets(codes, seq_ranges, hp):
    """_2_enumerate.jl:50-65.  Returns H as {packed key: (n,2) int64 array of (seq_num, pos)} in first-insertion order; filter ids
    and positions are 1-based like the reference's records, seq_num is the RANGE index (1-based), as in insert_H!."""
    keys, vals = [], []
    for ind, (a, b) in enumerate(seq_ranges, start=1):
        w = codes[a:b]
        if len(w) < 3:
            continue
        order = np.argsort(w["position"], kind="stable")          # sort(by = x -> x[1]) on (position, fil, seq, mag)
        pos = w["position"][order].astype(np.int64) + 1
        fil = w["fil"][order].astype(np.int64) + 1
        t = _triples(len(w))
        i, j, k = t[:, 0], t[:, 1], t[:, 2]
        keys.append(pack_key(fil[i], fil[j], fil[k], pos[j] - pos[i], pos[k] - pos[i]))
        vals.append(np.stack([np.full(len(t), ind, np.int64), pos[i]], axis=1))
    if not keys:
        return {}
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(inv, kind="stable")                        # values grouped by key, insertion order kept inside a key
    counts = np.bincount(inv, minlength=len(uniq))
    starts = np.concatenate([[0], np.cumsum(counts)])
    H = {}
    for u in np.argsort(first, kind="stable"):                    # Dictionary keeps insertion order
        H[int(uniq[u])] = vals[order[starts[u]:starts[u + 1]]]
    return H


def get_enriched_keys(H, max_word_combinations=num_pfms2process, dec=-5, count_from=cover_more_than, count_to=cover_at_least):
    """_3_make_pfms.jl:13-26 (get_words :3-11 keeps the `num_pfms2process` most frequent keys, not max_word_combinations)."""
    enriched = None
    for count in range(count_from, count_to - 1, dec):
        enriched = [k for k, v in H.items() if len(v) > count]
        if len(enriched) > max_word_combinations:
            top = sorted(enriched, key=lambda k: -len(H[k]))       # sort(length.(q), rev=true): stable
            return top[:num_pfms2process]
    return enriched or []


def enumerate_triplets_gpu(ctx, codes):
    """enumerate_triplets on the device (mb200_triplets_create): ranges, per-range stable sort by position, one hash-table count
    per triplet key.  Returns the device-resident dictionary (`_lib.Triplets`); .ranges() gives the reference's ranges."""
    return _lib.Triplets(ctx, codes)


def get_enriched_keys_gpu(t, max_word_combinations=num_pfms2process, dec=-5, count_from=cover_more_than, count_to=cover_at_least):
    """get_enriched_keys (_3_make_pfms.jl:13-26) on the device dictionary: every candidate list of the reference's loop is a subset
    of {count > count_to}, so one compaction of those keys (with their first-insertion rank = Dictionary order) serves the whole
    loop.  Returns a structured array (key, count, first) in the order the reference returns its keys."""
    cand = t.frequent(count_to)
    enriched = cand[:0]
    for count in range(count_from, count_to - 1, dec):
        enriched = cand[cand["count"] > count]
        if len(enriched) > max_word_combinations:
            top = enriched[np.argsort(-enriched["count"].astype(np.int64), kind="stable")]     # sort(length.(q), rev=true): stable
            return top[:num_pfms2process]
    return enriched


def sites_from_H(H, keys, hp, range_to_seq=None):
    """(motif, seq, pos, comp) records of the enriched keys, 0-based, for the GPU count-matrix kernel."""
    rec = []
    lens = []
    for m, k in enumerate(keys):
        v = H[k]
        lens.append(unpack_key(k, hp.h)["len"])
        seq = v[:, 0] - 1
        s = np.zeros(len(v), _lib.SITE_DTYPE)
        s["motif"], s["seq"], s["pos"] = m, seq, v[:, 1] - 1
        if v.shape[1] > 2:
            s["comp"] = v[:, 2]
        rec.append(s)
    return (np.concatenate(rec) if rec else np.zeros(0, _lib.SITE_DTYPE)), np.array(lens, np.int64)


def obtain_count_matrices(data, H, keys, hp):
    """_3_make_pfms.jl:28-46 on the GPU; Float32 matrices like the reference's."""
    sites, lens = sites_from_H(H, keys, hp)
    sites = sites[(sites["pos"].astype(np.int64) + lens[sites["motif"]]) <= data.L]
    return [c.astype(np.float32) for c in _lib.count_matrices(data.ctx, data.seqs, sites, lens)], sites, lens


def posdicts2countmats(ms: Motifs, data):
    """_h6_positions2countmat.jl:26-37: count matrices of ms.positions / ms.use_comp, Float16 like the reference."""
    rec = []
    for m in range(ms.num_motifs):
        for n, pos in ms.positions[m].items():
            s = np.zeros(len(pos), _lib.SITE_DTYPE)
            s["motif"], s["seq"], s["pos"], s["comp"] = m, n - 1, np.asarray(pos) - 1, np.asarray(ms.use_comp[m][n])
            rec.append(s)
    sites = np.concatenate(rec) if rec else np.zeros(0, _lib.SITE_DTYPE)
    # msa_add!(...; ps=0.01, return_count_mat=true) returns `msa .+ ps` (Float32 counts + the Float64 literal 0.01), then Float16
    return [(c.astype(np.float32).astype(np.float64) + 0.01).astype(np.float16) for c in _lib.count_matrices(data.ctx, data.seqs, sites, ms.lens)]


def run_thru(data, cdl, hp, ln, projs, this_bg, quantiles=(0.75, 0.65, 0.5, 0.45, 0.35, 0.25, 0.15, 0.05), codes=None):
    """run_thru (_g1_obtain_coutmats.jl:131-173), in-scope stages: code retrieval (GPU) -> per quantile: filter, ranges, triplet
    enumeration and key counting (GPU, csrc/triplets.cu), enriched keys -> merged dictionary (earlier quantiles win, like Dictionaries.merge's left-to-right update) ->
    count matrices (GPU) -> countmats2motifs.  Returns None when no key is enriched (:162-163)."""
    from .model import code_retrieval
    codes = code_retrieval(data, cdl, hp) if codes is None else codes
    merged = {}
    for q in quantiles:
        cf = filter_code_components_using_quantile(codes, q)
        t = enumerate_triplets_gpu(data.ctx, cf)
        ek = get_enriched_keys_gpu(t, max_word_combinations=1000)
        if len(ek):
            for k, v in zip(ek["key"].tolist(), t.values(ek["key"], total=int(ek["count"].sum()))):
                merged[k] = v                                      # merge(a, b): b's value replaces a's for equal keys
        t.free()
    if not merged:
        return None
    keys = list(merged.keys())
    cmats, sites, lens = obtain_count_matrices(data, merged, keys, hp)
    ms = countmats2motifs(cmats, this_bg)
    return ms if ms.num_motifs else None
