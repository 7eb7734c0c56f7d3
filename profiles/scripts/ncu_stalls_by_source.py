"""Attributes the warp-stall samples of an `ncu --set full --import-source on` capture to source lines and device functions.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; cuobjdump -xelf all build/csc.o; nvdisasm --print-line-info csc.sm_100a.cubin > csc.dis
       python profiles/scripts/ncu_stalls_by_source.py src.csv csc.dis k_csc_fused_fwd motifs.jl_b200/csrc/csc_fused.cuh
(the object must be the build that was profiled: SASS offsets are joined with nvdisasm's line table)"""
import collections
import csv
import re
import sys

src_csv, dis, kern, cu = sys.argv[1:5]
fn = cur = None
off2line = {}
for ln in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and fn and kern in fn:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ia = hdr.index("Address"); isamp = hdr.index("# Samples")
base = int(rows[2][ia], 16)
src = open(cu).read().split("\n")
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template.*)?__device__.*?\b(\w+)\(|__global__.*?\b(k_\w+)\(", l)
    if m:
        starts.append((i, m.group(1) or m.group(2)))


def owner(line):
    o = None
    for st, name in starts:
        if st <= line:
            o = name
    return o


byf, byl, byop, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
cufile = cu.split("/")[-1]
for r in rows[2:]:
    a = int(r[ia], 16) - base; n = int(r[isamp] or 0); tot += n
    fl = off2line.get(a)
    if fl is None:
        byf["?"] += n; continue
    f, l = fl
    byf[owner(l) if f == cufile else f] += n
    byl[(f, l)] += n
    w = r[1].split()
    op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "")).split(".")[0]
    byop[op] += n
print(f"{kern}: {tot} warp-stall samples")
print("by device function / header:")
for k, v in byf.most_common(16):
    print(f"{v:7d} {100 * v / tot:5.1f}%  {k}")
print("by source line:")
for (f, l), v in byl.most_common(24):
    print(f"{v:7d} {100 * v / tot:5.1f}%  {f}:{l}  {src[l - 1].strip()[:100] if f == cufile else ''}")
print("by SASS opcode:", ", ".join(f"{k} {v}" for k, v in byop.most_common(10)))
