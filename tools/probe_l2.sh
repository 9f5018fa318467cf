#!/bin/bash
# Do small frame groups that stay resident in the 126 MB L2 beat the HBM roofline? (240x180xD32: 33 MB of W/N/E per frame)
P="python tools/perf_probe.py --only"
for b in 1 2 3 4 6 8; do
  for alg in 5 6; do
    HSM_FRAME_GROUPS=0 $P 180,240,32,$b,50,$alg,1,0
  done
  HSM_FRAME_GROUPS=1 $P 180,240,32,$b,50,5,1,0
done
for b in 2 3 4; do for br in 2 3 6 8 12; do HSM_FRAME_GROUPS=1 $P 180,240,32,$b,50,5,1,$br; done; done
