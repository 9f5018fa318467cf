#!/bin/bash
# Final round-2 ncu evidence for the build with CTA widths / L2 hints / new persistent-kernel band height
# (run on the GPU box under gpurun; every command has run without ncu first: bench.py and tools/probe_r2*.sh).
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $OUT/launches_r2b.csv \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --frames-per-step 128 --no-configs > $OUT/ncu_launches_r2b.log 2>&1
full() {   # name, kernel regex, skip, probe args, extra ncu flags
    $NCU --set full $5 -k regex:$2 -s $3 -c 2 -f -o $OUT/prof_r2b_$1 python tools/perf_probe.py --only $4 > $OUT/ncu_r2b_$1.log 2>&1
    if [ -n "$5" ]; then
        ncu -i $OUT/prof_r2b_$1.ncu-rep --page source --csv > $OUT/prof_r2b_$1_source.csv 2>/dev/null
        python tools/ncu_source_mix.py $OUT/prof_r2b_$1_source.csv 2 > $OUT/prof_r2b_$1_source_mix.txt 2>&1
        rm -f $OUT/prof_r2b_$1_source.csv
    fi
    python tools/ncu_summary.py $OUT/prof_r2b_$1.ncu-rep > $OUT/prof_r2b_$1.txt 2>&1
    rm -f $OUT/prof_r2b_$1.ncu-rep
}
full c2_d32_batch64   k_iter_tmast 20 180,240,32,64,12,5,1,0 "--import-source on"
full c3_d64_batch4    k_iter_tmast 20 480,640,64,4,12,5,1,0 ""
full c3_d64_single    k_iter_flow  1  480,640,64,1,12,5,1,0 ""
full narrow_80x60_d22 k_iter_tmast 8  60,80,22,256,10,5,1,0 ""
ls -la $OUT/ | tail -20
