"""Per-rank pass of an N-way sharded run, emulated on one GPU: the first 1/k of the walks of the N=1e6
workload (Morton-contiguous, like gplum_b200/shard.py's interior set), timed alone.  Usage:
  GPLUM_B200_SPLIT_M={0,1,2,3,4} python tools/shard_probe.py [k ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gplum_b200 import disk, functors as F, tree
from gplum_b200.walks import Walks

n = 1000000
d = disk.make_disk(n)
ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
F.init(0); F.set_params(0.0, True, 0)
full = None
for k in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    m = w.n_walk // k
    sub = Walks(w.epi, w.epi_off[:m], w.ni[:m], w.adr_epj, w.epj_disp[:m], w.n_epj[:m], w.adr_spj, w.spj_disp[:m],
                w.n_spj[:m], w.epj_all, w.spj_all)
    F.walks_upload(sub)
    F.walks_run(repack=False)
    ms = F.walks_time(30, repack=False)
    if k == 1:
        full = ms
    print("split_m=%s 1/%d of the walks: %d walks, %.4f ms per pass%s" % (
        os.environ.get("GPLUM_B200_SPLIT_M", "2"), k, m, ms, "" if full is None else "  (x%d = %.3f of the full pass)" % (k, ms * k / full)))
