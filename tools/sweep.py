import subprocess, sys, json, os
cfgs = []
for br in (15, 30, 45, 60, 90):
    for st in (2, 3):
        cfgs.append("180,240,32,64,50,2,1,%d,%d" % (br, st))
cfgs += ["180,240,32,128,50,2,1,%d,3" % br for br in (30, 45, 60)]
cfgs += ["180,240,32,64,50,2,0,30,3", "180,240,32,64,50,2,1,30,3,7"]
cfgs += ["480,640,64,4,20,2,1,0,0,%d" % l for l in (0, 3, 9)] + ["480,640,64,4,20,2,1,%d,0,0" % br for br in (24, 48, 96)]
cfgs += ["270,480,128,8,20,2,1,0,0,%d" % l for l in (0, 9)]
cfgs += ["1080,1920,256,1,6,2,1,%d,0" % br for br in (0, 30, 60, 120)]
for c in cfgs:
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "perf_probe.py"), "--only", c],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True).stdout.strip().split("\n")[-1]
    try:
        d = json.loads(out)
        print("%-34s band %3d st %d lane %d dt %d  %.4f ms/iter  frac %.3f" % (d["shape"] + " b%d" % d["batch"], d["band_rows"], d["stages"], d["lane"], d["data_term"], d["ms_per_iter"], d["frac"]), flush=True)
    except Exception:
        print(c, "->", out[:200], flush=True)
