import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import hybrid_stereo_matching_b200 as hsm
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    rows, cols, levels, prior, dt = [int(v) for v in sys.argv[1].split(",")]
    l, r, d = synth.textured_pair(rows, cols, levels, 3)
    p = synth.sparse_prior(d, levels, 3) if prior else None
    m = hsm.StereoMRF((rows, cols), levels)
    m.set_option(capi.OPT_DATA_TERM, dt)
    m.set_option(capi.OPT_ALGORITHM, int(os.environ.get("HSM_ALGO", "2")))
    n_it = int(os.environ.get("HSM_NITER", "2"))
    out = m.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_it)
    m2 = hsm.StereoMRF((rows, cols), levels)
    m2.set_option(capi.OPT_ALGORITHM, 0)
    ref = m2.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_it)
    import numpy as np
    pl, pr = m._message_field, m2._message_field
    bad = {k: int((pl[k] != pr[k]).sum()) for k in pl}
    print(sys.argv[1], "ok, labels equal:", bool((out == ref).all()), "plane mismatches:", bad)
else:
    for cfg in ["180,240,32,0,0", "180,240,32,0,1", "180,240,32,1,1", "20,40,8,0,1", "20,40,32,0,1", "30,300,64,0,1", "16,400,200,0,1"]:
        res = subprocess.run([sys.executable, __file__, cfg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        print(cfg, "->", res.stdout.strip().split("\n")[-1][:300])
