"""discover_motifs(datapath, save_path; num_epochs) — the reference's only exported symbol (src/wrap.jl:1-11) — and
render_result! (src/render/render.jl:61-98) with the reference's output layout (render/const.jl, render/helpers.jl:27-88):

  <outdir>/summary.html
  <outdir>/logos_olap/d{i}.transfac, d{i}_c.transfac, d{i}.meme      (d{i}.png / d{i}_c.png need the external `weblogo`)
  <outdir>/logos_no_olap/, pics_olap/, pics_no_olap/                 (created empty, as the reference does)

Hot path on the GPU: training, code retrieval, the 8 scans, threshold filtering, counts.  The reference's CPU motif
post-processing between code retrieval and the scans is out of scope (see extract.run_thru) and the display ordering is the
reference's fallback (significant motifs first).
"""
from __future__ import annotations

import logging
import os
import shutil
import subprocess

import numpy as np

from . import extract, inference, loadfasta, model
from ._lib import Context

log = logging.getLogger("motifs_b200")
logo_olap_folder_name, logo_no_olap_folder_name = "logos_olap", "logos_no_olap"
pics_olap_folder_name, pics_no_olap_folder_name = "pics_olap", "pics_no_olap"


def get_rounded_pval(pval, low_pval):
    """render/helpers.jl:15-25."""
    s = repr(float(pval))
    if "e-" not in s:
        s = f"{float(f'{pval:.3g}')}"
    else:
        a, b = s.split("e-")
        s = a[:min(len(a), 4)] + "e-" + b.lstrip("0")
    return s if low_pval else f'<p style="color:grey">{s}</p>'


def save_pfms_as_transfac(logo_folder, cmats, sort_perm, numbers, data_name="unspecified"):
    """render/helpers.jl:27-88 (file contents identical; weblogo is invoked only when it is installed)."""
    pfms = [inference.countmat2pfm(c) for c in cmats]
    counts_each = [int(np.floor(float(np.asarray(c, np.float32)[:, 0].sum()))) for c in cmats]
    have_weblogo = shutil.which("weblogo") is not None
    for i, ind in zip(numbers, sort_perm):
        pfm, nsites = np.asarray(pfms[ind], np.float32), counts_each[ind]
        for suffix, mat in (("", pfm), ("_c", pfm[::-1, ::-1])):
            q = np.floor(mat * nsites).astype(np.int64)
            with open(os.path.join(logo_folder, f"d{i}{suffix}.transfac"), "w") as io:
                io.write("ID\t\nXX\t\nBF\t\nXX\t\nP0\tA\tC\tG\tT\n")
                for j in range(1, pfm.shape[1] + 1):
                    row = f"0{j}" if j < 10 else str(j)
                    io.write(f"{row}\t{q[0, j - 1]}\t{q[1, j - 1]}\t{q[2, j - 1]}\t{q[3, j - 1]}\n")
                io.write("XX\t\n")
            if have_weblogo:
                subprocess.run(["weblogo", "-D", "transfac", "-f", os.path.join(logo_folder, f"d{i}{suffix}.transfac"), "-n", "150",
                                "--number-fontsize", "17", "--errorbars", "NO", "-F", "png", "--fineprint", " ", "--resolution", "96",
                                "-s", "medium", "--fontsize", "24", "--small-fontsize", "18", "--color-scheme", "classic",
                                "-o", os.path.join(logo_folder, f"d{i}{suffix}.png")], check=False)
        with open(os.path.join(logo_folder, f"d{i}.meme"), "w") as io:
            io.write("MEME version 4\n\nALPHABET= ACGT\n\nstrands: + -\n\nBackground letter frequencies\nA 0.25 C 0.25 G 0.25 T 0.25\n\n")
            io.write(f"MOTIF {i} {data_name} \nletter-probability matrix: alength= 4 w= {pfm.shape[1]} nsites= {nsites} E= 0\n")
            for col in range(pfm.shape[1]):
                io.write(" " + "".join(f"{np.float16(pfm[a, col])} " for a in range(4)) + "\n")


def render_result_(target_folder, ms, data, bg, alpha_fisher=1e-5):
    """render_result! (render/render.jl:61-98): scan fg + bg, thresholds, filtered counts, test-set Fisher p-values, files."""
    folders = [target_folder] + [os.path.join(target_folder, f) for f in
                                 (logo_olap_folder_name, logo_no_olap_folder_name, pics_olap_folder_name, pics_no_olap_folder_name)]
    for f in folders:
        os.makedirs(f, exist_ok=True)
    # scan_w_gpu! x2, filter_positions_scores_usecomp! and get_uniq_counts (render.jl:70-76) as four fused device scans: score
    # histograms of foreground and background -> the same thresholds -> filtered counts.  The positions / scores / use_comp
    # dictionaries of the reference are never built (inference.scan_w_gpu_ / filter_positions_scores_usecomp_ still offer them):
    # with ~900 PWMs x 18 000 sequences they are 12 million Python objects and took 48 of this function's 50 seconds.
    log.info("Scanning the foreground and the shuffled background...")
    counts_fg, _counts_bg = inference.filter_positions_scores_usecomp_fused_(ms, data, bg)
    active_counts = counts_fg[:, 1].astype(np.float64)          # unique start positions per motif = get_uniq_counts(ms)[0]
    log.info("Calculating p-values...")
    pvec, uniq_test = inference.pvec_from_test_data(ms, data)
    order = list(range(ms.num_motifs))                          # obtain_groupings_for_display1 is cosmetic (out of scope)
    sig = [i for i in order if pvec[i] < alpha_fisher]
    insig = [i for i in order if pvec[i] >= alpha_fisher]
    display = sig + insig
    pvalues = [get_rounded_pval(pvec[i], pvec[i] < alpha_fisher) for i in display]
    log.info("save the PWMs...")
    save_pfms_as_transfac(folders[1], ms.cmats, display, list(range(1, ms.num_motifs + 1)))
    totals = (active_counts + uniq_test).astype(np.int64)[display]
    rows = "".join(f"<tr><td>D{j + 1}</td><td>{p}</td><td>{c}</td><td><img src=\"{logo_olap_folder_name}/d{j + 1}.png\"></td></tr>\n"
                   for j, (p, c) in enumerate(zip(pvalues, totals)))
    with open(os.path.join(target_folder, "summary.html"), "w") as io:
        io.write("<html><body><p>Number of sequences: " + str(data.N + data.N_test) + "</p>\n<table>\n"
                 "<tr><th>Label</th><th>P-value</th><th># instances</th><th>Logo</th></tr>\n" + rows + "</table></body></html>\n")
    return {"pvec": pvec, "order": display, "counts": totals, "score_thresh": ms.score_thresh}


def discover_motifs(datapath, save_path, num_epochs=None, device=0, rng=None, verbose=False):
    """discover_motifs(datapath, save_path; num_epochs=nothing) (wrap.jl:1-11)."""
    rng = rng or np.random.default_rng()
    ctx = Context(device)
    log.info("load data")
    data = loadfasta.FASTA_DNA(datapath, ctx, rng=rng)
    this_bg = loadfasta.get_data_bg(data)
    log.info("training...")
    cdl, hp, ln, projs, m = model.train_ucdl(data, num_epochs=num_epochs, rng=rng, verbose=verbose)
    log.info("extract motifs...")
    ms = extract.run_thru(data, cdl, hp, ln, projs, this_bg)
    m.free()
    if ms is None:
        data.free()
        raise RuntimeError("no enriched word combination survived (run_thru returned nothing, _g1_obtain_coutmats.jl:162-163)")
    out = render_result_(save_path, ms, data, this_bg)
    data.free()
    return ms, out
