"""Generates tests/golden/*.npz from the CPU oracle (the reference has no fixtures and cannot run here —
PARITY UNPINNED; these vectors pin OUR oracle so that a later change to it is noticed, and give the GPU tests a
second, frozen comparison point).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from motifs_jl_b200 import synth  # noqa: E402
from oracle import scan_oracle as so, stats_oracle as st  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def config1_inputs():
    """SURVEY §8d config 1 (scan part): seed 1, 1000 x 100 bp planted gapped motif, split 901/99, K=3 PWMs."""
    a = synth.planted_gapped(1000, 100, 1)
    train = a[:901]
    bg = synth.shuffle_rows(train, 101)
    full = ["TGACGT" + sp + "ACGTCA" for sp in ("AAAAA", "CCCCC", "GGGGG", "TTTTT", "ACGTA")]
    cms = [synth.count_matrix_from_sites(["TGACGT"] * 40), synth.count_matrix_from_sites(["ACGTCA"] * 40),
           synth.count_matrix_from_sites(full * 8)]
    # data ACGT frequencies (MOTIFs.jl:35-39)
    codes = so.ascii_to_codes(train)
    bgfreq = (np.bincount(codes.ravel(), minlength=4) / codes.size).astype(np.float32)
    ms = synth.motifs_from_count_matrices(cms, bgfreq)
    return train, bg, ms, bgfreq


def main():
    train, bg, ms, bgfreq = config1_inputs()
    codes, codes_bg = so.ascii_to_codes(train), so.ascii_to_codes(bg)
    pw, lens = so.pack_pwms(ms.pwms)
    hits, counts = so.scan(pw, lens, codes)
    hits_bg, counts_bg = so.scan(pw, lens, codes_bg)
    thr = []
    for m in range(len(lens)):
        segs0 = [(r.start - 1, r.stop - 1) for r in ms.effective_segments[m]]
        thr.append(np.float16(st.get_best_thresh(hits[hits["motif"] == m]["score_f16"].view(np.float16),
                                                 hits_bg[hits_bg["motif"] == m]["score_f16"].view(np.float16),
                                                 segs0, ms.pwms[m], 901 * 100, bgfreq)))
    thr = np.array(thr, np.float16)
    fh, fc = so.scan(pw, lens, codes, thr)
    fhb, fcb = so.scan(pw, lens, codes_bg, thr)
    pvec = st.fisher_pvec(fc[:, 2], fcb[:, 2], 901, 100)
    np.savez_compressed(os.path.join(HERE, "scan_config1.npz"), train=train, bg=bg, pwms=pw.view(np.uint16), lens=lens,
                        bgfreq=bgfreq, hits=hits, counts=counts, hits_bg=hits_bg, counts_bg=counts_bg, thresh=thr.view(np.uint16),
                        filt_hits=fh, filt_counts=fc, filt_hits_bg=fhb, filt_counts_bg=fcb, pvec=pvec)
    # small adversarial case: non-finite entries, palindromes, motif longer than the sequence
    a = synth.random_ascii(48, 60, 71)
    a[:8] = np.frombuffer((b"ACGTACGT" * 8)[:60], np.uint8)
    ms2 = synth.motifs_from_count_matrices(synth.random_count_matrices(9, 8, 30, 72) + synth.random_count_matrices(1, 61, 62, 73)
                                           + [synth.count_matrix_from_sites(["ACGTACGT"] * 30)])
    ms2.pwms[1][2, 3] = np.float16(-np.inf)
    ms2.pwms[2][0, 0] = np.float16(np.nan)
    pw2, lens2 = so.pack_pwms(ms2.pwms)
    h2, c2 = so.scan(pw2, lens2, so.ascii_to_codes(a))
    np.savez_compressed(os.path.join(HERE, "scan_edge.npz"), seqs=a, pwms=pw2.view(np.uint16), lens=lens2, hits=h2, counts=c2)
    print("config1:", len(hits), "hits;", counts.tolist(), "thresholds", thr, "pvec", pvec)
    print("edge:", len(h2), "hits;", c2[:, 0].tolist())


if __name__ == "__main__":
    main()


def csc_golden():
    """config 1, first batch of 6: loss, gradient and final codes of the default network (position-space oracle, fp32)."""
    import torch
    from oracle import csc_oracle as co
    hp = co.Hyperparam()
    flat = co.init_params(hp, 1)
    train, _, _, _ = config1_inputs()
    codes = so.ascii_to_codes(train[:6])
    loss, grad, aux = co.loss_and_grad(codes, flat, hp, "pos", torch.float32)
    recs = co.code_retrieval(so.ascii_to_codes(train[:12]), flat, hp)
    np.savez_compressed(os.path.join(HERE, "csc_config1.npz"), codes=codes, flat=flat, loss=np.float32(loss), grad=grad,
                        x=aux["x"].detach().numpy(), z=aux["z"].detach().numpy().astype(np.float32), D=aux["D"].detach().numpy(),
                        code_records=recs, code_input=so.ascii_to_codes(train[:12]))
    print("csc: loss", loss, "grad max", np.abs(grad).max(), "codes", len(recs))


if __name__ == "__main__":
    csc_golden()
