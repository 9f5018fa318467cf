OUT=gpurun_out
ncu --clock-control none --set full --import-source on -k regex:k_iter_tmast -s 20 -c 2 -f -o $OUT/prof_mix python tools/perf_probe.py --only 180,240,32,64,12,5,1,0 > $OUT/ncu_mix.log 2>&1
ncu -i $OUT/prof_mix.ncu-rep --page source --csv > $OUT/prof_mix_source.csv 2>/dev/null
python tools/ncu_source_mix.py $OUT/prof_mix_source.csv 2 > $OUT/r2_k_iter_tmast_source_mix.txt 2>&1
head -c 3000 $OUT/prof_mix_source.csv > $OUT/prof_mix_source_head.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_mix_source.csv')))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr)]
ia,ie,iss=hdr.index('Source'),hdr.index('Instructions Executed'),hdr.index('Warp Stall Sampling (All Samples)')
with open('gpurun_out/r2_k_iter_tmast_sass_counts.csv','w') as f:
    w=csv.writer(f); w.writerow(['sass','instructions_executed','stall_samples'])
    for r in data: w.writerow([r[ia].strip(), r[ie], r[iss]])
PY
rm -f $OUT/prof_mix.ncu-rep $OUT/prof_mix_source.csv
ls -la $OUT
