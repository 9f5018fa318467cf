#!/bin/bash
# Round-2 ncu evidence (run on the GPU box under gpurun; outputs land in gpurun_out/, summaries are made by
# tools/make_profiles_r2.py in the build container).  Every command has run without ncu first (bench / probes).
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the bench command (serialised, cold cache: compare SHARES, not absolute times)
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $OUT/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --frames-per-step 128 --no-configs > $OUT/ncu_launches_r2.log 2>&1
# 2. full captures, two launches each, of the message kernel at the shapes DESIGN.md tabulates
# (the .ncu-rep files are 20-40 MB each and gpurun brings back at most 64 MiB: summarise on the box, drop the reports)
summarise() {   # name
    python tools/ncu_summary.py $OUT/prof_r2_$1.ncu-rep > $OUT/prof_r2_$1.txt 2>&1
    rm -f $OUT/prof_r2_$1.ncu-rep
}
full() {   # name, kernel regex, skip, probe args, extra ncu flags
    $NCU --set full $5 -k regex:$2 -s $3 -c 2 -f -o $OUT/prof_r2_$1 python tools/perf_probe.py --only $4 > $OUT/ncu_$1.log 2>&1
    if [ -n "$5" ]; then
        ncu -i $OUT/prof_r2_$1.ncu-rep --page source --csv > $OUT/prof_r2_$1_source.csv 2>/dev/null
        python tools/ncu_source_mix.py $OUT/prof_r2_$1_source.csv 2 > $OUT/prof_r2_$1_source_mix.txt 2>&1
        rm -f $OUT/prof_r2_$1_source.csv
    fi
    summarise $1
}
full c2_d32_batch64   k_iter_tmast 20 180,240,32,64,12,5,1,0 "--import-source on"
full c3_d64_batch4    k_iter_tmast 20 480,640,64,4,12,5,1,0 ""
full c3_d64_single    k_iter_flow  1  480,640,64,1,12,5,1,0 ""
full c5_d256_single   k_iter_flow  1  1080,1920,256,1,4,5,1,0 ""
full c5_d256_tma      k_iter_tma   4  1080,1920,256,1,4,2,1,0 ""
full d20_batch64      k_iter_tmast 20 180,240,20,64,12,5,1,0 ""
full d22_batch64      k_iter_tmast 20 180,240,22,64,12,5,1,0 ""
full d50_batch64      k_iter_tmast 20 180,240,50,64,12,5,1,0 ""
# 3. the stand-alone readout kernel (used after the persistent dataflow kernel; batches fuse it into the last iteration)
$NCU --set full -k regex:k_readout -s 2 -c 1 -f -o $OUT/prof_r2_readout \
    python tools/latency_probe.py > $OUT/ncu_readout.log 2>&1
summarise readout
ls -la $OUT/
