"""CUDA path (through the C ABI) against the committed golden fixtures (tests/golden/, produced by make_golden.py)."""
import os

import numpy as np
import pytest

import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def test_scan_config1_golden(ctx):
    g = np.load(os.path.join(GOLD, "scan_config1.npz"))
    pw, lens, thr = g["pwms"], g["lens"], g["thresh"]
    for rows, hk, ck, fhk, fck in (("train", "hits", "counts", "filt_hits", "filt_counts"), ("bg", "hits_bg", "counts_bg", "filt_hits_bg", "filt_counts_bg")):
        seqs = ctx.seqs_from_ascii(g[rows])
        hits, counts = ctx.scan(seqs, pw, lens)
        assert np.array_equal(hits, g[hk]) and np.array_equal(counts, g[ck])
        fh, fc = ctx.scan(seqs, pw, lens, thr)
        assert np.array_equal(fh, g[fhk]) and np.array_equal(fc, g[fck])
        seqs.free()


def test_scan_edge_golden(ctx):
    g = np.load(os.path.join(GOLD, "scan_edge.npz"))
    seqs = ctx.seqs_from_ascii(g["seqs"])
    hits, counts = ctx.scan(seqs, g["pwms"], g["lens"])
    assert np.array_equal(hits, g["hits"]) and np.array_equal(counts, g["counts"])
    seqs.free()


def test_csc_golden(ctx):
    g = np.load(os.path.join(GOLD, "csc_config1.npz"))
    hp = mdl.Hyperparam()
    seqs = ctx.seqs_from_ascii(_ACGT[g["codes"]])
    m = mb._lib.CscModel(ctx, hp, 100)
    m.set_params(g["flat"])
    loss, grad = m.loss_grad(seqs, np.arange(6))
    assert loss[0, 0] == pytest.approx(float(g["loss"]), rel=1e-5)
    assert np.abs(grad - g["grad"][: m.n_trainable]).max() <= 2e-4 * np.abs(g["grad"]).max()
    x = m.get_buffer("x", 6 * 82 * 24).reshape(6, 82, 24)
    assert np.array_equal(x != 0, g["x"] != 0)
    m.free(); seqs.free()
    # code records of the first two batches
    seqs = ctx.seqs_from_ascii(_ACGT[g["code_input"]])
    mf = mb._lib.CscModel(ctx, hp, 100, n_groups=2, forward_only=True)
    mf.set_params(g["flat"])
    rec = mf.codes(seqs)
    exp = g["code_records"]
    assert len(rec) == len(exp)
    for f in ("position", "fil", "seq"):
        assert np.array_equal(rec[f], exp[f])
    mf.free(); seqs.free()
