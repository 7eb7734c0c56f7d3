# per-kernel durations of batched code retrieval (500 groups per launch); arg "tc" selects the tcgen05 path
mkdir -p gpurun_out
timeout 200 python profiles/scripts/prof_codes.py $1 > gpurun_out/plain_codes.log 2>&1 || { tail -5 gpurun_out/plain_codes.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/codes_launches.csv python profiles/scripts/prof_codes.py $1 > gpurun_out/ncu_codes.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/codes_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki].split('(')[0]][0]+=1; agg[r[ki].split('(')[0]][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print(f"{k[:36]:36s} {v[0]:4d} {v[1]/1e3:8.1f} us avg {v[1]/v[0]/1e3:7.2f} {v[1]/tot:.3f}")
print(len(rows)-1, 'launches', tot/1e3, 'us')
PY
