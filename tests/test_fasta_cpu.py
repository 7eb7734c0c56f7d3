"""FASTA parsing and the train/test split run in the library on the host (csrc/fasta.cu): no GPU needed.
Rules checked are the reference's (loadfasta/helpers.jl:83-108, 141-159)."""
import os

import numpy as np

from motifs_jl_b200 import _lib, synth


def test_fasta_read_follows_reference_rules(tmp_path):
    a = synth.random_ascii(50, 40, 1)
    path = os.path.join(tmp_path, "r.fa")
    kept = []
    with open(path, "w") as fh:
        for i, r in enumerate(a):
            s = bytes(r).decode()
            if i == 3:
                s = s[:20] + "N" + s[21:]                 # dropped: contains N (helpers.jl:92)
            elif i == 4:
                s = s[:20] + "n" + s[21:]                 # dropped: contains n
            elif i == 5:
                s = s[:30]                                # dropped: not as long as the first read (helpers.jl:98)
            elif i == 7:
                kept.append(s); s = s.lower()             # kept, upper-cased (helpers.jl:107)
            else:
                kept.append(s)
            fh.write(f">seq{i} some header text\n{s[:25]}\n{s[25:]}\n")      # multi-line records are joined (helpers.jl:90-91)
    rows = _lib.fasta_read(path)
    assert rows.shape == (47, 40)
    assert [bytes(r).decode() for r in rows] == kept
    capped = _lib.fasta_read(path, max_entries=10)        # the cap applies BEFORE the equal-length filter (helpers.jl:95,98)
    assert [bytes(r).decode() for r in capped] == kept[:9]        # records 0..9 without 3 and 4, then record 5 falls to the length rule


def test_fasta_split_sizes_and_partition():
    for n in (0, 1, 10, 1000, 20000):
        tr, te = _lib.fasta_split(n, 0.9, True, 7)
        assert len(te) == int(np.floor((1 - 0.9) * n)) and len(tr) + len(te) == n      # 99 of 1 000, 1 999 of 20 000 (helpers.jl:144)
        assert np.array_equal(np.sort(np.concatenate([tr, te])), np.arange(n))
    tr, te = _lib.fasta_split(1000, 0.9, True, 7)
    tr2, te2 = _lib.fasta_split(1000, 0.9, True, 7)
    assert np.array_equal(tr, tr2) and np.array_equal(te, te2)                       # reproducible from the seed
    tr3, _ = _lib.fasta_split(1000, 0.9, True, 8)
    assert not np.array_equal(tr, tr3)
    assert not np.array_equal(tr, np.sort(tr))                                         # train keeps randperm order (setdiff), not sorted
    tr, te = _lib.fasta_split(1000, 0.9, False, 7)
    assert np.array_equal(te, np.arange(901, 1000))                                    # shuffle=false: the last n_test reads
    assert np.array_equal(np.sort(tr), np.arange(901))
