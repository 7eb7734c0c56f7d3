# ncu launch list of a many-group training step (the data-parallel config-3 path): bash profiles/scripts/r02_groups_launches.sh [Lb] [groups]
Lb=${1:-200}; G=${2:-64}
mkdir -p gpurun_out
timeout 200 python profiles/scripts/prof_csc_groups.py $Lb $G > gpurun_out/plain_groups.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 800 --csv --log-file gpurun_out/r02_groups_launches.csv python profiles/scripts/prof_csc_groups.py $Lb $G > gpurun_out/ncu_groups.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r02_groups_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki].split('(')[0]][0]+=1; agg[r[ki].split('(')[0]][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:30]: print(f"{k[:40]:40s} {v[0]:4d} {v[1]/1e3:9.1f} us avg {v[1]/v[0]/1e3:8.2f} {v[1]/tot:.3f}")
print(len(rows)-1, 'launches', tot/1e3, 'us')
PY
