#!/bin/bash
# round-2 probes behind DESIGN.md 4.1's later paragraphs (run on the GPU box; same-box A/B, compare rows of one file):
#   narrow   band heights / kernel choice for batches of the shipped 80x60 / 70x60 frames  -> profiles/r2_probe_narrow_bands.jsonl
#   hints    L2 eviction hints of the default kernel family (HSM_L2_HINTS bit mask), twice -> profiles/r2_probe_l2_hints.txt
#   flow     band heights of the persistent dataflow kernel for one frame                  -> profiles/r2_probe_flow_bands.txt
P="python tools/perf_probe.py --only"
case "${1:-all}" in
narrow|all)
  for shape in 60,80,22 60,70,27 60,80,20; do for algo in 5 6; do for br in 0 15 20 30 60; do
    $P $shape,256,10,$algo,1,$br
  done; done; done ;;&
hints|all)
  for rep in 1 2; do for h in 0 5 1 4; do
    for cfg in 180,240,32,128,50,5,1,0 480,640,64,4,50,5,1,0 180,240,50,64,50,5,1,0 180,240,20,64,50,5,1,0 \
               60,80,22,256,10,5,1,0 270,480,128,8,50,5,1,0 180,240,64,32,50,5,1,0; do
      echo -n "rep=$rep hints=$h "; HSM_L2_HINTS=$h $P $cfg
    done
  done; done ;;&
flow|all)
  for cfg in 270,480,128,1,50,5,1,0 270,480,128,1,50,5,1,12 270,480,128,1,50,5,1,15 270,480,128,1,50,5,1,18 \
             540,960,64,1,50,5,1,0 540,960,64,1,50,5,1,16 540,960,64,1,50,5,1,24 540,960,64,1,50,5,1,32 \
             480,640,64,1,50,5,1,0 480,640,64,1,50,5,1,12 480,640,64,1,50,5,1,16 480,640,64,1,50,5,1,18 \
             480,640,64,1,50,5,1,22 480,640,64,1,50,5,1,24 480,640,64,1,50,5,1,32; do
    $P $cfg
  done ;;
esac
