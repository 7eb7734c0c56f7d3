"""FASTA loader mirror (loadfasta/helpers.jl, fasta.jl): filtering, split sizes, shuffles, packed upload; then the whole
mirror pipeline on a small file: load -> train a few steps -> code retrieval -> scan -> thresholds -> Fisher."""
import os

import numpy as np
import pytest

from motifs_jl_b200 import inference, loadfasta, model as mdl, synth
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu


def test_fasta_loader_follows_reference_rules(ctx, tmp_path):
    a = synth.planted_gapped(1000, 100, 1)
    path = os.path.join(tmp_path, "reads.fa")
    with open(path, "w") as fh:
        for i, r in enumerate(a):
            s = bytes(r).decode()
            if i == 3:
                s = s[:50] + "N" + s[51:]          # dropped: contains N (helpers.jl:92)
            if i == 5:
                s = s[:90]                         # dropped: length differs from the first read (helpers.jl:98)
            if i == 7:
                s = s.lower()                      # kept, upper-cased (helpers.jl:107)
            fh.write(f">seq{i}\n{s[:60]}\n{s[60:]}\n")   # multi-line records are joined (helpers.jl:90)
    data = loadfasta.FASTA_DNA(path, ctx, rng=np.random.default_rng(0))
    assert data.N + data.N_test == 998 and data.N_test == int(np.floor((1 - 0.9) * 998)) and data.L == 100
    assert np.array_equal(data.seqs.download(), so.pack_codes(so.ascii_to_codes(data.ascii)))
    # k=1 background keeps every read's base composition; k=2 keeps its 2-mer (non-overlapping) multiset
    assert np.array_equal(np.sort(data.ascii, axis=1), np.sort(data.ascii_bg, axis=1))
    assert np.array_equal(np.sort(data.ascii_test, axis=1), np.sort(data.ascii_bg_test, axis=1))
    bg = loadfasta.get_data_bg(data)
    assert bg.dtype == np.float32 and abs(float(bg.sum()) - 1) < 1e-6
    data.free()


def test_mirror_pipeline_end_to_end(ctx):
    a = synth.planted_gapped(600, 100, 4)
    data = loadfasta.FASTA_DNA([bytes(r).decode() for r in a], ctx, rng=np.random.default_rng(1))
    cdl, hp, ln, _, m = mdl.train_ucdl(data, num_epochs=1, rng=np.random.default_rng(2), verbose=False, max_steps=40)
    codes = mdl.code_retrieval(data, cdl, hp)
    assert len(codes) > 0 and codes["seq"].max() < data.N - data.N % 6 and codes["fil"].max() < hp.K
    ms = synth.motifs_from_count_matrices([synth.count_matrix_from_sites(["TGACGT"] * 40), synth.count_matrix_from_sites(["ACGTCA"] * 40)],
                                          loadfasta.get_data_bg(data))
    inference.scan_w_gpu_(ms, data)
    inference.scan_w_gpu_(ms, data, bg=True)
    inference.filter_positions_scores_usecomp_(ms, data, loadfasta.get_data_bg(data))
    pvec, uniq_test = inference.pvec_from_test_data(ms, data)
    assert np.all(pvec < 1e-5) and np.all(uniq_test > 0)          # the planted half sites are enriched over the shuffled background
    m.free(); data.free()


def test_count_matrices_kernel(ctx):
    from motifs_jl_b200 import _lib
    from oracle import extract_oracle as eo
    a = synth.random_ascii(40, 70, 3)
    codes = so.ascii_to_codes(a)
    seqs = ctx.seqs_from_ascii(a)
    rng = np.random.default_rng(5)
    lens = np.array([8, 21, 13], np.int64)
    sites = np.zeros(300, _lib.SITE_DTYPE)
    sites["motif"] = rng.integers(0, 3, 300)
    sites["seq"] = rng.integers(0, 40, 300)
    sites["pos"] = [int(rng.integers(0, 70 - lens[m] + 1)) for m in sites["motif"]]
    sites["comp"] = rng.integers(0, 2, 300)
    got = _lib.count_matrices(ctx, seqs, sites, lens)
    for m in range(3):
        sel = sites[sites["motif"] == m]
        exp = eo.count_matrix(codes, sel["seq"].astype(int) + 1, sel["pos"].astype(int) + 1, sel["comp"].astype(bool), int(lens[m]))
        assert np.array_equal(got[m].astype(np.float32), exp)
    seqs.free()


def test_discover_motifs_writes_reference_layout(ctx, tmp_path):
    """discover_motifs(fasta, outdir; num_epochs) end to end on a planted motif; checks the output tree of render/const.jl."""
    import motifs_jl_b200 as mb
    a = synth.planted_gapped(1500, 100, 8)
    fa = os.path.join(tmp_path, "reads.fa")
    loadfasta.write_fasta(fa, a)
    out = os.path.join(tmp_path, "out")
    ms, res = mb.discover_motifs(fa, out, num_epochs=2, rng=np.random.default_rng(0))
    assert os.path.isfile(os.path.join(out, "summary.html"))
    for d in ("logos_olap", "logos_no_olap", "pics_olap", "pics_no_olap"):
        assert os.path.isdir(os.path.join(out, d))
    assert ms.num_motifs >= 1
    for i in range(1, ms.num_motifs + 1):
        for f in (f"d{i}.transfac", f"d{i}_c.transfac", f"d{i}.meme"):
            assert os.path.isfile(os.path.join(out, "logos_olap", f))
    assert (res["pvec"] < 1e-5).any()                       # the planted motif is found and enriched on the held-out split
