"""Round-2 additions through the C ABI: device split / k-mer-preserving shuffles / base counts, the held hit list of an
overflowing scan, the in-library communicator on one rank, sharded code retrieval and posdicts2countmats."""
from collections import Counter

import numpy as np
import pytest

import motifs_jl_b200 as mb
from motifs_jl_b200 import _lib, extract, inference, model as mdl, synth
from oracle import csc_oracle as co, extract_oracle as eo, scan_oracle as so

pytestmark = pytest.mark.gpu


def _kmers(row, k):
    return Counter(bytes(row[i:i + k]) for i in range(len(row) - k + 1))


def test_gather_to_ascii_and_base_counts(ctx):
    a = synth.random_ascii(300, 77, 11)
    seqs = ctx.seqs_from_ascii(a)
    assert np.array_equal(seqs.to_ascii(), a)
    idx = np.random.default_rng(0).permutation(300)[:123]
    sub = seqs.gather(idx)
    assert np.array_equal(sub.to_ascii(), a[idx])
    cnt, trans = sub.base_counts()
    codes = so.ascii_to_codes(a[idx])
    assert np.array_equal(cnt, np.bincount(codes.ravel(), minlength=4))
    t = np.zeros((4, 4), np.int64)
    np.add.at(t, (codes[:, :-1].ravel(), codes[:, 1:].ravel()), 1)
    assert np.array_equal(trans, t)
    sub.free(); seqs.free()


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_shuffle_preserves_kmer_counts(ctx, k):
    a = synth.planted_gapped(200, 100, 3)
    a[5] = ord("A")                                                  # degenerate reads: one letter, two letters alternating
    a[6, ::2], a[6, 1::2] = ord("C"), ord("G")
    seqs = ctx.seqs_from_ascii(a)
    sh = seqs.shuffle(k, seed=42)
    b = sh.to_ascii()
    assert b.shape == a.shape
    moved = 0
    for r in range(len(a)):
        assert _kmers(a[r], k) == _kmers(b[r], k), (k, r)             # seq_shuffle(s; k) keeps every k-mer count of every read
        if k > 1:
            assert bytes(a[r, :k - 1]) == bytes(b[r, :k - 1]) and bytes(a[r, -(k - 1):]) == bytes(b[r, -(k - 1):])
        moved += not np.array_equal(a[r], b[r])
    assert moved >= 190                                              # ...and really shuffles
    # reproducible from the seed; a different seed gives a different background
    assert np.array_equal(seqs.shuffle(k, seed=42).to_ascii(), b)
    assert not np.array_equal(seqs.shuffle(k, seed=43).to_ascii(), b)
    # sequence i draws from stream first_stream + i: a shard shuffled on its own equals the rows of the whole
    part = seqs.gather(np.arange(50, 120))
    assert np.array_equal(part.shuffle(k, seed=42, first_stream=50).to_ascii(), b[50:120])
    for s in (seqs, sh, part):
        s.free()


def test_shuffle_long_sequence_is_a_permutation(ctx):
    a = synth.random_ascii(1, 300_000, 5)                             # > 65 536 bp: sort-by-random-key path (BASELINE config 5's background)
    seqs = ctx.seqs_from_ascii(a)
    b = seqs.shuffle(1, seed=9).to_ascii()
    assert np.array_equal(np.bincount(a[0], minlength=256), np.bincount(b[0], minlength=256)) and not np.array_equal(a, b)
    with pytest.raises(mb._lib.MB200Error):
        seqs.shuffle(2, seed=9)
    seqs.free()


def test_overflowing_hit_list_is_held_not_rescanned(ctx):
    a = synth.planted_gapped(4000, 100, 2)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(40, 8, 20, 1))
    pw, lens = inference.pack_pwms(ms), ms.lens
    seqs = ctx.seqs_from_ascii(a)
    hits, counts = ctx.scan(seqs, pw, lens)                             # default capacity 65 536 < number of hits: one scan + take_hits
    assert len(hits) > (1 << 16)
    ohits, ocounts = so.scan(pw, lens, so.ascii_to_codes(a), None)
    assert len(hits) == len(ohits) and all(np.array_equal(hits[f], ohits[f]) for f in ("seq", "pos", "motif", "score_f16", "comp"))
    assert np.array_equal(counts, ocounts)
    big, _ = ctx.scan(seqs, pw, lens, hits_cap=len(ohits))
    assert np.array_equal(big, hits)
    with pytest.raises(mb._lib.MB200Error):                              # explicit capacity: the overflow is reported, the list can still be taken
        ctx.scan(seqs, pw, lens, hits_cap=1000)
    seqs.free()


def test_communicator_single_rank_and_sharded_codes(ctx):
    """world = 1 exercises the NCCL binding (dlopen, init, collectives as identities); explicit (rank, world) shards decode disjoint
    batch ranges whose concatenation is the single-call result (mb200_csc_codes_sharded, _1_code_retrieval.jl:38-50)."""
    c2 = mb.Context(0)
    c2.comm_init(mb.Context.comm_unique_id(), 0, 1)
    rank, world, ver = c2.comm_info()
    assert (rank, world) == (0, 1) and ver >= 20000
    x = np.arange(5, dtype=np.int64)
    assert np.array_equal(c2.comm_allreduce_i64(x.copy()), x) and np.array_equal(c2.comm_allgather(x)[0], x)
    hp = mdl.Hyperparam()
    a = synth.planted_gapped(61, 100, 8)
    seqs = c2.seqs_from_ascii(a)
    flat = co.init_params(co.Hyperparam(), 3)
    m = mb._lib.CscModel(c2, hp, 100, n_groups=4, forward_only=True)
    m.set_params(flat)
    m.broadcast_params(0)
    whole = m.codes(seqs)
    assert len(whole) > 0 and whole["seq"].max() < 60
    assert np.array_equal(m.codes(seqs, shard="comm"), whole)
    for w in (2, 3, 8, 16):                                               # 10 batches over up to 16 ranks: some ranks get none
        parts = [m.codes(seqs, shard=(r, w)) for r in range(w)]
        assert np.array_equal(np.concatenate(parts), whole), w
        assert sum(len(p) > 0 for p in parts) <= 10
    # counts reduce on one rank is the identity
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(5, 8, 20, 1))
    pw = inference.pack_pwms(ms)
    thr = synth.stated_thresholds(ms, 0.5)
    _, c_plain = c2.scan(seqs, pw, ms.lens, thr, want_hits=False)
    _, c_red = c2.scan(seqs, pw, ms.lens, thr, want_hits=False, reduce=True)
    assert np.array_equal(c_plain, c_red)
    m.free(); seqs.free(); c2.comm_destroy(); c2.close()


def test_posdicts2countmats_matches_literal(ctx):
    a = synth.planted_gapped(120, 100, 6)
    codes = so.ascii_to_codes(a)

    class D:
        pass
    data = D()
    data.ctx, data.seqs = ctx, ctx.seqs_from_ascii(a)
    ms = synth.motifs_from_count_matrices([synth.count_matrix_from_sites(["TGACGT"] * 40), synth.count_matrix_from_sites(["ACGTCAGG"] * 40)])
    hits, _ = ctx.scan(data.seqs, inference.pack_pwms(ms), ms.lens, synth.stated_thresholds(ms, 0.6))
    ms.positions, ms.scores, ms.use_comp = inference._hits_to_dicts(hits, ms.num_motifs)
    got = extract.posdicts2countmats(ms, data)
    exp = eo.posdicts2countmats(codes, ms.positions, ms.use_comp, ms.lens)
    for g, e in zip(got, exp):
        assert g.dtype == np.float16 and np.array_equal(g, e)
        assert g.min() > 0                                            # the 0.01 pseudo-count of msa_add!(...; return_count_mat=true)
    data.seqs.free()
