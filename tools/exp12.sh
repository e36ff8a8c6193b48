# per-pass kernel times of the multi-pass sizes (ncu launch list of tools/prof_case.py)
for c in n16 n18 n20 n21 n22 n23 n24; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t61_$c.csv python tools/prof_case.py $c 1 > /dev/null 2>&1
python - <<PY
import csv,re
rows=list(csv.reader(open('gpurun_out/t61_$c.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
out=[]
for r in rows[hi+1:]:
    if len(r)>vi and 'fft_unit' in r[ki]:
        out.append((re.sub(r'\(tfft::UnitPlan.*','',r[ki]).replace('void tfft::',''), float(r[vi].replace(',',''))/1000))
print('$c', ' | '.join(f"{n} {v:.0f} us" for n,v in out[-3 if '$c'=='n24' else -2:]))
PY
done
