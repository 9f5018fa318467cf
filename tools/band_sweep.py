"""Band-height sweep over several shapes (input for the heuristic in choose_band_rows).
python tools/band_sweep.py > gpurun_out/band_sweep.jsonl"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
SWEEP = [
    ("180,240,32,64,50", (10, 15, 18, 20, 23, 26, 30, 36, 45, 60, 90)),
    ("180,240,64,32,50", (12, 18, 23, 26, 30, 36, 45)),
    ("480,640,64,4,20", (8, 10, 12, 15, 18, 20, 24, 30, 40, 48, 60)),
    ("270,480,128,8,20", (9, 12, 15, 18, 23, 27, 34, 45, 68)),
    ("180,240,16,128,50", (12, 15, 20, 30, 45, 60, 90)),
    ("1080,1920,256,1,20", (36, 45, 54, 64, 72, 90, 108, 135)),
    ("360,480,32,16,50", (10, 15, 20, 30, 45, 60, 90)),
    ("270,1920,256,1,20", (15, 18, 23, 27, 34, 45, 68)),
    ("135,1920,256,1,20", (15, 18, 23, 27, 34, 45, 68)),
]
SWEEP += [("480,640,64,1,50", (6, 8, 10, 12, 14, 16, 20)), ("180,240,32,1,50", (2, 3, 4, 5, 6)),
          ("60,80,22,1,10", (2, 3, 4, 6))]
if len(sys.argv) > 1:                      # python tools/band_sweep.py 3,4  -> only those entries
    SWEEP = [SWEEP[int(i)] for i in sys.argv[1].split(",")]
for shape, bands in SWEEP:
    for br in (0,) + tuple(bands):
        out = subprocess.run([sys.executable, os.path.join(HERE, "perf_probe.py"), "--only", "%s,%s,1,%d,0" % (shape, os.environ.get("HSM_SWEEP_ALGO", "5"), br)],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True).stdout
        print(out.strip().splitlines()[-1], flush=True)
