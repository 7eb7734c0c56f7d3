"""Parity of the CUDA scan path (through the C ABI) against the CPU oracle.  Bit-exact: positions, strands,
Float16 score bits, counts.  Reference: inference/_h3_1_alignment.jl, _s2_filter_pos_w_scores.jl:116-125,
_h4_overlap_ratio.jl:5-79."""
import numpy as np
import pytest

import motifs_jl_b200 as mb
from motifs_jl_b200 import inference, synth, _lib
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu


def _check(ctx, ascii_rows, ms, thresh=None, strands=3):
    codes = so.ascii_to_codes(ascii_rows)
    pw, lens = so.pack_pwms(ms.pwms)
    seqs = ctx.seqs_from_ascii(ascii_rows)
    hits, counts = ctx.scan(seqs, pw, lens, thresh, fwd=bool(strands & 1), rc=bool(strands & 2))
    ohits, ocounts = so.scan(pw, lens, codes, thresh, strands)
    assert len(hits) == len(ohits)
    for f in ("seq", "pos", "motif", "score_f16", "comp"):
        assert np.array_equal(hits[f], ohits[f]), f
    assert np.array_equal(counts, ocounts)
    # counts-only call must agree with the hits call
    _, c2 = ctx.scan(seqs, pw, lens, thresh, fwd=bool(strands & 1), rc=bool(strands & 2), want_hits=False)
    assert np.array_equal(c2, ocounts)
    if thresh is not None:
        # the call above took the tensor-core pre-filter when eligible; the SIMT kernel must give the same hits and counts
        h3, c3 = ctx.scan(seqs, pw, lens, thresh, fwd=bool(strands & 1), rc=bool(strands & 2), tensor=False)
        assert _lib.scan_last_path(ctx) == 0
        assert np.array_equal(c3, ocounts)
        assert h3.tobytes() == hits.tobytes()
    seqs.free()
    return hits, counts


def _tc_path(ctx, ascii_rows, ms, thresh, strands=3):
    """path taken by a thresholded counts-only scan (1 = tensor-core pre-filter)"""
    pw, lens = so.pack_pwms(ms.pwms)
    seqs = ctx.seqs_from_ascii(ascii_rows)
    ctx.scan(seqs, pw, lens, thresh, fwd=bool(strands & 1), rc=bool(strands & 2), want_hits=False)
    seqs.free()
    return _lib.scan_last_path(ctx)


def test_pack_matches_layout(ctx):
    a = synth.random_ascii(37, 101, 7)
    s = ctx.seqs_from_ascii(a)
    assert np.array_equal(s.download(), so.pack_codes(so.ascii_to_codes(a)))
    s.free()
    # lower case and one-hot input give the same words
    low = np.char.lower(a.view("S1")).view(np.uint8).reshape(a.shape)
    s2 = ctx.seqs_from_ascii(low)
    assert np.array_equal(s2.download(), so.pack_codes(so.ascii_to_codes(a)))
    s2.free()
    codes = so.ascii_to_codes(a)
    onehot = np.zeros((37, 101, 4), np.float32)
    np.put_along_axis(onehot, codes[:, :, None].astype(np.int64), 1.0, axis=2)
    s3 = ctx.seqs_from_onehot(onehot)
    assert np.array_equal(s3.download(), so.pack_codes(codes))
    s3.free()


def test_bad_symbol_rejected(ctx):
    a = synth.random_ascii(4, 50, 1)
    a[2, 17] = ord("N")
    with pytest.raises(mb.MB200Error) as e:
        ctx.seqs_from_ascii(a)
    assert e.value.code == mb._lib.E_BAD_SEQUENCE


def test_config1_planted_motif_scan(ctx):
    """config 1: 1,000 x 100 bp with a planted gapped motif; K=3 PWMs (half sites and the full site)."""
    a = synth.planted_gapped(1000, 100, 1)
    sites_full = ["TGACGT" + sp + "ACGTCA" for sp in ("AAAAA", "CCCCC", "GGGGG", "TTTTT", "ACGTA")]
    cms = [synth.count_matrix_from_sites(["TGACGT"] * 40), synth.count_matrix_from_sites(["ACGTCA"] * 40),
           synth.count_matrix_from_sites(sites_full * 8)]
    ms = synth.motifs_from_count_matrices(cms)
    hits, counts = _check(ctx, a, ms)
    assert counts[:, 0].min() > 0
    thr = synth.stated_thresholds(ms, 0.7)
    _check(ctx, a, ms, thr)
    _check(ctx, a, ms, thr, strands=1)
    _check(ctx, a, ms, thr, strands=2)
    bg = synth.shuffle_rows(a, 11)
    _check(ctx, bg, ms, thr)


@pytest.mark.parametrize("K,Lb,N,seed", [(1, 100, 300, 3), (7, 64, 257, 4), (8, 200, 100, 5), (9, 33, 64, 6),
                                         (40, 200, 500, 8), (130, 100, 200, 9)])
def test_random_pwms(ctx, K, Lb, N, seed):
    a = synth.random_ascii(N, Lb, seed)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(K, 8, min(40, Lb), seed))
    _check(ctx, a, ms)                                   # reference scan semantics: score > 0
    _check(ctx, a, ms, synth.stated_thresholds(ms, 0.5))


def test_long_and_short_edges(ctx):
    # motif as long as the sequence, longer than the sequence, length 1 and the maximum supported length
    a = synth.random_ascii(50, 40, 21)
    cms = synth.random_count_matrices(3, 40, 40, 22) + synth.random_count_matrices(2, 41, 64, 23) + \
        synth.random_count_matrices(2, 1, 1, 24)
    ms = synth.motifs_from_count_matrices(cms)
    _check(ctx, a, ms)
    b = synth.random_ascii(20, 300, 25)
    ms2 = synth.motifs_from_count_matrices(synth.random_count_matrices(5, 60, 64, 26))
    _check(ctx, b, ms2)
    _check(ctx, b, ms2, np.full(5, -3.0, np.float16))    # negative threshold: still needs score > 0
    # longer than MB200_MAX_MOTIF_LEN: slow path, mixed with short motifs, and a motif as long as the sequence
    ms3 = synth.motifs_from_count_matrices(synth.random_count_matrices(3, 65, 120, 27) + synth.random_count_matrices(20, 8, 30, 28)
                                           + synth.random_count_matrices(1, 300, 300, 29))
    for p in ms3.pwms[:3] + ms3.pwms[-1:]:
        p[:] = (p / np.float16(8)).astype(np.float16)       # keep long sums finite and some of them positive
        p[:, ::3] = np.abs(p[:, ::3])
    h3, c3 = _check(ctx, b, ms3)
    assert c3[:3, 0].sum() > 0


def test_nonfinite_and_negative_pwms(ctx):
    a = synth.random_ascii(64, 80, 31)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(6, 8, 12, 32))
    ms.pwms[0][:] = np.float16(-1.0)                     # all negative: no hit anywhere
    ms.pwms[1][2, 3] = np.float16(-np.inf)               # -Inf entry: Inf*0 = NaN poisons every other base of that column
    ms.pwms[2][1, 0] = np.float16(np.nan)
    ms.pwms[3][0, 5] = np.float16(np.inf)
    hits, counts = _check(ctx, a, ms)
    assert counts[0, 0] == 0


def test_palindrome_and_dense_hits(ctx):
    # a palindromic site hits on both strands at the same start: unique starts < hits; union_ranges quirk exercised
    a = np.frombuffer((b"ACGTACGT" * 13)[:100] * 16, np.uint8).reshape(16, 100).copy()
    ms = synth.motifs_from_count_matrices([synth.count_matrix_from_sites(["ACGTACGT"] * 30)])
    hits, counts = _check(ctx, a, ms)
    assert counts[0, 1] < counts[0, 0]
    assert counts[0, 2] <= counts[0, 3]


def test_all_positive_pwm_every_position_hits(ctx):
    a = synth.random_ascii(33, 130, 41)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(2, 8, 9, 42))
    for p in ms.pwms:
        p[:] = np.abs(p) + np.float16(0.01)
    hits, counts = _check(ctx, a, ms)
    assert counts[0, 0] == 2 * 33 * (130 - ms.lens[0] + 1)


def test_single_long_sequence(ctx):
    a = synth.random_ascii(1, 50000, 51)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(12, 8, 40, 52))
    _check(ctx, a, ms, synth.stated_thresholds(ms, 0.6))


def test_hits_overflow_reports_required_size(ctx):
    a = synth.random_ascii(100, 100, 61)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(4, 8, 12, 62))
    pw, lens = so.pack_pwms(ms.pwms)
    seqs = ctx.seqs_from_ascii(a)
    with pytest.raises(mb.MB200Error) as e:
        ctx.scan(seqs, pw, lens, hits_cap=3)
    assert e.value.code == mb._lib.E_HITS_OVERFLOW
    seqs.free()


def test_reference_entry_points(ctx):
    """gpu_scan / scan_w_gpu! / filter / counts / Fisher through the host mirror == oracle-side restatement."""
    from types import SimpleNamespace
    from oracle import stats_oracle as st
    a = synth.planted_gapped(600, 100, 2)
    bg = synth.shuffle_rows(a, 3)
    cms = [synth.count_matrix_from_sites(["TGACGT"] * 40), synth.count_matrix_from_sites(["TGACGTAAAAAACGTCA"] * 9 + ["TGACGTCCCCCACGTCA"] * 9)]
    ms = synth.motifs_from_count_matrices(cms)
    data = SimpleNamespace(N=600, L=100, seqs=ctx.seqs_from_ascii(a), seqs_bg=ctx.seqs_from_ascii(bg))
    inference.scan_w_gpu_(ms, data)
    inference.scan_w_gpu_(ms, data, bg=True)
    codes, codes_bg = so.ascii_to_codes(a), so.ascii_to_codes(bg)
    pw, lens = so.pack_pwms(ms.pwms)
    oh, _ = so.scan(pw, lens, codes)
    ohb, _ = so.scan(pw, lens, codes_bg)
    # dict contents: forward hits ascending then rc hits ascending, 1-based
    for m in range(2):
        sel = oh[oh["motif"] == m]
        for n in np.unique(sel["seq"]):
            r = sel[sel["seq"] == n]
            assert np.array_equal(ms.positions[m][int(n) + 1], r["pos"].astype(np.int64) + 1)
            assert np.array_equal(ms.scores[m][int(n) + 1].view(np.uint16), r["score_f16"])
            assert np.array_equal(ms.use_comp[m][int(n) + 1], r["comp"].astype(bool))
        assert len(ms.positions[m]) == len(np.unique(sel["seq"]))
    bgfreq = np.array([0.25, 0.25, 0.25, 0.25], np.float32)
    inference.filter_positions_scores_usecomp_(ms, data, bgfreq)
    for m in range(2):
        segs0 = [(r.start - 1, r.stop - 1) for r in ms.effective_segments[m]]
        t = st.get_best_thresh(oh[oh["motif"] == m]["score_f16"].view(np.float16), ohb[ohb["motif"] == m]["score_f16"].view(np.float16),
                               segs0, ms.pwms[m], 600 * 100, bgfreq)
        assert np.float16(t) == ms.score_thresh[m]
    # fused counts with the derived thresholds == dict-based counts of the mirror == oracle
    c = inference.scan_counts(ms, data, thresh=ms.score_thresh)
    cb = inference.scan_counts(ms, data, bg=True, thresh=ms.score_thresh)
    _, oc = so.scan(pw, lens, codes, ms.score_thresh, want_hits=False)
    _, ocb = so.scan(pw, lens, codes_bg, ms.score_thresh, want_hits=False)
    assert np.array_equal(c, oc) and np.array_equal(cb, ocb)
    uq, uqb = inference.get_uniq_counts(ms)
    assert np.array_equal(uq, c[:, 1]) and np.array_equal(uqb, cb[:, 1])
    p = inference.get_fisher_p_values(ms, data)
    assert np.allclose(p, st.fisher_pvec(oc[:, 2], ocb[:, 2], 600, 100), rtol=1e-9, atol=0)


def test_chromosome_scale_counts(ctx):
    """config 5 shape at test size: one long sequence, tiled scan with motif-length halos, blocked parallel counting (W > 4096
    mask words) — counts equal the oracle's on the untiled sequence; the 1-mer shuffled background gives the Fisher inputs."""
    from motifs_jl_b200 import inference
    from types import SimpleNamespace
    Lb = 400_000
    a = synth.random_ascii(1, Lb, 81)
    site = np.frombuffer(b"TGACGTCATTGACGTCA", np.uint8)
    rng = np.random.default_rng(82)
    for p in rng.integers(0, Lb - 20, 300):
        a[0, p:p + len(site)] = site                     # planted family
    a[0, Lb - 17:] = site                                # a hit ending exactly at the end of the sequence
    a[0, 131072 - 8:131072 - 8 + 17] = site              # a hit straddling a counting-block boundary (512 words = 16384 positions)
    bg = synth.shuffle_rows(a, 83)
    cms = [synth.count_matrix_from_sites(["TGACGTCATTGACGTCA"] * 30), synth.count_matrix_from_sites(["TGACGTCA"] * 30)] + \
        synth.random_count_matrices(10, 8, 40, 84)
    ms = synth.motifs_from_count_matrices(cms)
    pw, lens = so.pack_pwms(ms.pwms)
    thr = synth.stated_thresholds(ms, 0.6)
    res = []
    for rows in (a, bg):
        seqs = ctx.seqs_from_ascii(rows)
        _, c = ctx.scan(seqs, pw, lens, thr, want_hits=False)
        _, oc = so.scan(pw, lens, so.ascii_to_codes(rows), thr, want_hits=False)
        assert np.array_equal(c, oc)
        _, c0 = ctx.scan(seqs, pw, lens, None, want_hits=False)          # dense regime (score > 0): exercises the look-back
        _, oc0 = so.scan(pw, lens, so.ascii_to_codes(rows), None, want_hits=False)
        assert np.array_equal(c0, oc0)
        res.append(c)
        seqs.free()
    p = inference.fisher_pvec(res[0][:, 2], res[1][:, 2], SimpleNamespace(N=1, L=Lb, N_test=0))
    assert p[0] < 1e-10 and p[1] < 1e-10                                # the planted family is enriched over the shuffle


def test_score_histogram_and_fused_thresholds(ctx):
    """mb200_scan_hist == histogram of the oracle's hit scores; thresholds derived from histograms == thresholds derived from
    the hit dictionaries (filter_positions_scores_usecomp!), for the Touzet branch and for the Fisher-sweep branch."""
    from types import SimpleNamespace
    import copy
    a = synth.planted_gapped(500, 100, 12)
    bg = synth.shuffle_rows(a, 13)
    cms = [synth.count_matrix_from_sites(["TGACGT"] * 40),                                   # short segments: Touzet branch
           synth.count_matrix_from_sites(["TGACGTAAAAAACGTCA"] * 9 + ["TGACGTCCCCCACGTCA"] * 9),
           synth.count_matrix_from_sites(["ACGTTGCAACGTTGCAAC"] * 25)]                      # one long high-IC segment: sweep branch
    ms = synth.motifs_from_count_matrices(cms)
    assert any(all(len(r) >= 15 for r in segs) for segs in ms.effective_segments)
    data = SimpleNamespace(N=500, L=100, seqs=ctx.seqs_from_ascii(a), seqs_bg=ctx.seqs_from_ascii(bg))
    pw, lens = so.pack_pwms(ms.pwms)
    h = inference.scan_hist(ms, data)
    oh, _ = so.scan(pw, lens, so.ascii_to_codes(a))
    exp = np.zeros_like(h)
    np.add.at(exp, (oh["motif"].astype(np.int64), oh["score_f16"].astype(np.int64)), 1)
    assert np.array_equal(h, exp)
    bgfreq = np.full(4, 0.25, np.float32)
    ms_dict = copy.deepcopy(ms)
    inference.scan_w_gpu_(ms_dict, data)
    inference.scan_w_gpu_(ms_dict, data, bg=True)
    inference.filter_positions_scores_usecomp_(ms_dict, data, bgfreq)
    c, cb = inference.filter_positions_scores_usecomp_fused_(ms, data, bgfreq)
    assert np.array_equal(ms.score_thresh.view(np.uint16), ms_dict.score_thresh.view(np.uint16))
    uq, uqb = inference.get_uniq_counts(ms_dict)
    assert np.array_equal(uq, c[:, 1]) and np.array_equal(uqb, cb[:, 1])
    tot = [inference.get_total_occupied_positions(inference.get_union_ranges(p, l)) for p, l in zip(ms_dict.positions, ms_dict.lens)]
    assert np.array_equal(np.array(tot), c[:, 2])


def test_empty_and_degenerate_inputs(ctx):
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(3, 8, 12, 91))
    pw, lens = so.pack_pwms(ms.pwms)
    # no sequences at all
    s0 = ctx.seqs_from_ascii(np.zeros((0, 50), np.uint8))
    hits, counts = ctx.scan(s0, pw, lens)
    assert len(hits) == 0 and counts.sum() == 0
    s0.free()
    # every motif longer than the sequences: nothing can be scored
    s1 = ctx.seqs_from_ascii(synth.random_ascii(9, 7, 92))
    hits, counts = ctx.scan(s1, pw, lens)
    assert len(hits) == 0 and counts.sum() == 0
    s1.free()
    # a single motif, a single sequence, a single valid position
    one = synth.motifs_from_count_matrices([synth.count_matrix_from_sites(["ACGTACGT"] * 20)])
    _check(ctx, np.frombuffer(b"ACGTACGT", np.uint8).reshape(1, 8).copy(), one)
    # argument errors
    s2 = ctx.seqs_from_ascii(synth.random_ascii(4, 30, 93))
    with pytest.raises(mb.MB200Error):
        ctx.scan(s2, pw, np.array([8, 0, 9]))                # zero-length motif
    with pytest.raises(mb.MB200Error):
        ctx.scan(s2, pw, lens, fwd=False, rc=False)          # no strand requested
    s2.free()
    onehot = np.zeros((3, 20, 4), np.float32); onehot[:, :, 1] = 1; onehot[1, 4, 2] = 1     # two ones in one column
    with pytest.raises(mb.MB200Error) as e:
        ctx.seqs_from_onehot(onehot)
    assert e.value.code == mb._lib.E_BAD_SEQUENCE


def test_asynchronous_upload_overlapping_the_scan(ctx):
    """mb200_seqs_from_ascii_async: the scan may start on the first chunks while later ones are still being copied; results are
    those of the blocking upload (1.2 M x 200 bp = 240 MB = four 64 MB chunks, several scan batches), bad symbols are reported by
    the first consumer."""
    N, Lb = 1_200_000, 200
    a = synth.random_ascii(N, Lb, 11)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(40, 8, 30, 3))
    pw, lens = so.pack_pwms(ms.pwms)
    thr = synth.stated_thresholds(ms, 0.6)
    s1 = ctx.seqs_from_ascii(a)
    _, c1 = ctx.scan(s1, pw, lens, thr, want_hits=False, want_counts=True)
    s2 = ctx.seqs_from_host_ptr(a.ctypes.data, N, Lb, wait=False)
    _, c2 = ctx.scan(s2, pw, lens, thr, want_hits=False, want_counts=True)
    assert np.array_equal(c1, c2)
    assert np.array_equal(s1.download(), s2.download())
    s3 = ctx.seqs_from_host_ptr(a.ctypes.data, N, Lb, wait=False)           # no scan: wait() completes it
    s3.wait()
    assert np.array_equal(s3.download(), s1.download())
    for s in (s1, s2, s3):
        s.free()
    b = a[:5000].copy()
    b[4321, 77] = ord("N")
    s4 = ctx.seqs_from_host_ptr(b.ctypes.data, 5000, Lb, wait=False)
    with pytest.raises(mb.MB200Error) as e:
        ctx.scan(s4, pw, lens, thr, want_hits=False, want_counts=True)
    assert e.value.code == mb._lib.E_BAD_SEQUENCE
    s4.free()
    s5 = ctx.seqs_from_host_ptr(b.ctypes.data, 5000, Lb, wait=False)
    s5.free()                                                               # freeing a pending upload is safe


# ---------------------------------------------------------------------------------------------------------------------
# Tensor-core pre-filter (csrc/scan_tc.cuh): every thresholded _check above already runs it next to the SIMT kernel and the
# oracle.  The cases below aim at what could break a filter: scores that land exactly on the threshold, partial sums whose
# Float16 rounding error is large, windows that straddle sequences and tiles, disabled slots, and the overflow fallback.
# ---------------------------------------------------------------------------------------------------------------------
def test_tensor_path_is_taken_and_exact_at_the_threshold(ctx):
    a = synth.random_ascii(3000, 100, 71)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(150, 8, 40, 72))
    thr = synth.stated_thresholds(ms, 0.5)
    assert _tc_path(ctx, a, ms, thr) == 1
    # thresholds equal to scores that occur: "score > thresh" must drop exactly those hits
    pw, lens = so.pack_pwms(ms.pwms)
    oh, _ = so.scan(pw, lens, so.ascii_to_codes(a), thr, 3)
    thr2 = thr.copy()
    for m in range(len(ms.pwms)):
        sc = np.sort(oh[oh["motif"] == m]["score_f16"].view(np.float16))
        if len(sc):
            thr2[m] = sc[len(sc) // 2]
    h, c = _check(ctx, a, ms, thr2)
    assert 0 < len(h) < len(oh)


@pytest.mark.parametrize("scale,seed", [(1.0, 81), (16.0, 82), (0.01, 83)])
def test_tensor_path_large_and_tiny_magnitudes(ctx, scale, seed):
    # entries up to +-250 (rounding error of a partial sum ~0.1 per add) and entries around 1e-2 (Float16 spacing effects differ)
    a = synth.random_ascii(2000, 77, seed)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(70, 8, 40, seed + 100))
    for p in ms.pwms:
        p[:] = (p.astype(np.float32) * scale).astype(np.float16)
    thr = synth.stated_thresholds(ms, 0.35)
    assert _tc_path(ctx, a, ms, thr) == 1
    _check(ctx, a, ms, thr)
    _check(ctx, a, ms, thr, strands=1)
    _check(ctx, a, ms, thr, strands=2)


def test_tensor_path_short_sequences_and_tile_edges(ctx):
    # sequences shorter than a 256-position tile, lengths that are not multiples of anything, motifs as long as the sequence
    for Lb, N, seed in [(9, 5000, 91), (33, 999, 92), (64, 513, 93), (255, 40, 94), (257, 40, 95), (1000, 7, 96)]:
        a = synth.random_ascii(N, Lb, seed)
        ms = synth.motifs_from_count_matrices(synth.random_count_matrices(20, 1, min(64, Lb + 3), seed + 50))
        thr = synth.stated_thresholds(ms, 0.4)
        _check(ctx, a, ms, thr)


def test_tensor_path_zero_threshold_and_disabled_slots(ctx):
    a = synth.random_ascii(400, 120, 97)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(12, 8, 30, 98))
    ms.pwms[1][2, 3] = np.float16(-np.inf)               # -Inf entry: the motif never hits
    ms.pwms[4][:] = np.float16(-1.0)                     # best window below any threshold
    thr = synth.stated_thresholds(ms, 0.5)
    thr[2] = np.float16(np.nan)                          # NaN threshold: never a hit
    thr[3] = np.float16(np.inf)
    thr[5] = np.float16(0.0)
    thr[6] = np.float16(-2.0)
    thr[1] = np.float16(1.0)
    thr[4] = np.float16(1.0)
    assert _tc_path(ctx, a, ms, thr) == 1
    h, c = _check(ctx, a, ms, thr)
    assert c[1, 0] == 0 and c[2, 0] == 0 and c[3, 0] == 0 and c[4, 0] == 0 and c[5, 0] > 0
    # +Inf / NaN entries are left to the SIMT kernel
    ms.pwms[7][0, 5] = np.float16(np.inf)
    assert _tc_path(ctx, a, ms, thr) == 0
    _check(ctx, a, ms, thr)


def test_tensor_path_candidate_overflow_falls_back(ctx):
    # thresholds of zero on all-positive PWMs: every cell is a candidate; 33.5 M records overflow the list and the scan must
    # finish on the SIMT kernel with the same counts
    a = synth.random_ascii(20000, 200, 99)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(6, 8, 12, 100))
    for p in ms.pwms:
        p[:] = np.abs(p) + np.float16(0.01)
    thr = np.zeros(6, np.float16)
    pw, lens = so.pack_pwms(ms.pwms)
    seqs = ctx.seqs_from_ascii(a)
    _, c = ctx.scan(seqs, pw, lens, thr, want_hits=False)
    assert _lib.scan_last_path(ctx) == 2
    _, c0 = ctx.scan(seqs, pw, lens, thr, want_hits=False, tensor=False)
    assert np.array_equal(c, c0)
    assert c[0, 0] == 2 * 20000 * (200 - ms.lens[0] + 1)
    seqs.free()


def test_tensor_path_multi_batch_equals_simt_at_scale(ctx):
    """BASELINE config 4 shape at a size that takes three batches of the pipelined tensor-core path (two alternating buffer sets, CTA
    re-balancing between batches, mask words cleared by the listed count): counts must equal the SIMT kernel's on the same sequences,
    be additive over a split of the sequences, and equal the oracle's on a sample."""
    N, Lb, K = 700_000, 200, 500
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(K, 8, 40, 4))          # bench.py's motif set
    thr = synth.stated_thresholds(ms, 0.7)
    pw, lens = so.pack_pwms(ms.pwms)
    a = synth.random_ascii(N, Lb, 123)
    seqs = ctx.seqs_from_ascii(a)
    _, c_tc = ctx.scan(seqs, pw, lens, thr, want_hits=False)
    assert _lib.scan_last_path(ctx) == 1
    _, c_tc2 = ctx.scan(seqs, pw, lens, thr, want_hits=False)                  # second call: cached CTA split, mask buffer reused as is
    assert np.array_equal(c_tc, c_tc2)
    _, c_simt = ctx.scan(seqs, pw, lens, thr, want_hits=False, tensor=False)
    assert _lib.scan_last_path(ctx) == 0
    assert np.array_equal(c_tc, c_simt)
    _, c_tc3 = ctx.scan(seqs, pw, lens, thr, want_hits=False)                  # after the SIMT kernel dirtied the mask buffer
    assert np.array_equal(c_tc, c_tc3)
    # sorted hit lists over three batches: listed-unit counting, emit, clearing of the listed units between batches
    h_tc, ch_tc = ctx.scan(seqs, pw, lens, thr, hits_cap=8_000_000)
    assert _lib.scan_last_path(ctx) == 1 and np.array_equal(ch_tc, c_tc) and len(h_tc) == int(c_tc[:, 0].sum())
    _, c_tc4 = ctx.scan(seqs, pw, lens, thr, want_hits=False)                  # the hit-list scan must leave the mask buffer clean
    assert np.array_equal(c_tc, c_tc4)
    h_simt, _ = ctx.scan(seqs, pw, lens, thr, hits_cap=8_000_000, tensor=False)
    assert h_tc.tobytes() == h_simt.tobytes()
    seqs.free()
    h = N // 2 + 12345
    parts = []
    for rows in (a[:h], a[h:]):
        s = ctx.seqs_from_ascii(rows)
        parts.append(ctx.scan(s, pw, lens, thr, want_hits=False)[1])
        s.free()
    assert np.array_equal(parts[0] + parts[1], c_tc)
    sub = a[300_000:303_000]
    s = ctx.seqs_from_ascii(sub)
    _, c_sub = ctx.scan(s, pw, lens, thr, want_hits=False)
    s.free()
    _, oc = so.scan(pw, lens, so.ascii_to_codes(sub), thr, want_hits=False)
    assert np.array_equal(c_sub, oc)
    assert c_tc[:, 0].sum() > 0


def test_tensor_path_many_motifs(ctx):
    # 1100 motifs = 9 slot blocks (pairs of neighbouring blocks + one single entry); 2300 motifs exceed the 16 work entries a launch
    # carries in its parameters and stay on the SIMT kernel
    a = synth.random_ascii(300, 120, 201)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(1100, 6, 40, 202))
    thr = synth.stated_thresholds(ms, 0.55)
    assert _tc_path(ctx, a, ms, thr) == 1
    _check(ctx, a, ms, thr)
    ms2 = synth.motifs_from_count_matrices(synth.random_count_matrices(2300, 6, 30, 203))
    thr2 = synth.stated_thresholds(ms2, 0.6)
    assert _tc_path(ctx, a[:100], ms2, thr2) == 0
    _check(ctx, a[:100], ms2, thr2)


def test_config4_sample_of_100k_sequences_against_the_oracle(ctx):
    """counts of a 100 000-sequence sample and the hit lists of 1 % of it must equal the oracle's (50 of the
    bench's 500 PWMs, so that the CPU side stays within seconds); both kernels."""
    N, Lb = 100_000, 200
    cms = synth.random_count_matrices(500, 8, 40, 4)                                           # bench.py's motif set
    ms = synth.motifs_from_count_matrices([cms[i] for i in range(0, 500, 10)])
    thr = synth.stated_thresholds(ms, 0.7)
    pw, lens = so.pack_pwms(ms.pwms)
    a = synth.random_ascii(N, Lb, 2024)
    codes = so.ascii_to_codes(a)
    seqs = ctx.seqs_from_ascii(a)
    _, c_tc = ctx.scan(seqs, pw, lens, thr, want_hits=False)
    assert _lib.scan_last_path(ctx) == 1
    seqs.free()
    s20 = ctx.seqs_from_ascii(a[:20_000])
    _, c_gt0 = ctx.scan(s20, pw, lens, None, want_hits=False)                   # the reference's own semantics: hit = score > 0
    s20.free()
    _, oc = so.scan(pw, lens, codes, thr, want_hits=False)
    assert np.array_equal(c_tc, oc)
    _, oc0 = so.scan(pw, lens, codes[:20_000], None, want_hits=False)
    assert np.array_equal(c_gt0, oc0)
    sub = a[37_000:38_000]
    s = ctx.seqs_from_ascii(sub)
    for t in (thr, None):
        hits, _ = ctx.scan(s, pw, lens, t, hits_cap=4_000_000)
        ohits, _ = so.scan(pw, lens, codes[37_000:38_000], t)
        assert len(hits) == len(ohits) and len(hits) > 0
        for f in ("seq", "pos", "motif", "score_f16", "comp"):
            assert np.array_equal(hits[f], ohits[f]), f
    s.free()
