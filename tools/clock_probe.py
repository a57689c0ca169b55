"""Does a short pass run at full SM clock?  1/8 shard of the N = 1e6 workload timed (a) cold: after 0.5 s of idle,
(b) after 100 ms of back-to-back FFMA kernels, (c) after 300 launches of itself."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gplum_b200 import disk, functors as F, tree
from gplum_b200.walks import Walks
n = 1000000
d = disk.make_disk(n)
ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
F.init(0); F.set_params(0.0, True, 0)
for k in (8, 1):
    m = w.n_walk // k
    sub = Walks(w.epi, w.epi_off[:m], w.ni[:m], w.adr_epj, w.epj_disp[:m], w.n_epj[:m], w.adr_spj, w.spj_disp[:m], w.n_spj[:m], w.epj_all, w.spj_all)
    F.walks_upload(sub); F.walks_run(repack=False)
    for label, prep in (("idle 0.5 s", lambda: time.sleep(0.5)), ("after 100 ms of FFMA", lambda: F.fp32_peak(300)),
                        ("after 300 launches", lambda: F.walks_time(300, repack=False)), ("idle 0.5 s again", lambda: time.sleep(0.5))):
        prep()
        a = F.walks_time(5, repack=False); b = F.walks_time(30, repack=False); c = F.walks_time(200, repack=False)
        print("1/%d %-22s  5 launches %.4f  30 launches %.4f  200 launches %.4f ms per pass" % (k, label, a, b, c))
