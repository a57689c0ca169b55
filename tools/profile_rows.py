"""One pass over rows f1-f3 at N=1e6 (BASELINE configs[2]) for ncu: resident state -> GPU list build ->
force pass with capture -> changeover correction -> kick -> Kepler drift -> pull.  Run under
`ncu --set full -k regex:...`; prints nothing that is a bench value."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gplum_b200 import disk, functors as F, state as ST, structs as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = disk.make_disk(n)
ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
F.init(0); F.set_params(0.0, True, 0)
prm_c, prm_i = S.corr_params(), ST.iso_params()
dt = float(prm_i["dt_tree"][0])
ST.upload(ST.make_epj(d["pos"], d["vel"], d["mass"], ro, rs), np.zeros(n), np.zeros(n))
F.soft_corr_enable(True)
for k in range(2):
    sz = ST.tree_build(n_group_limit=512)
    F.walks_run(repack=False)
    F.correct_long_run(prm_c)
    ST.kick(dt)
    ST.drift(prm_i, k * dt, (k + 1) * dt)
    rec, idx = ST.pull_unhandled(n)
    ST.push(rec, idx)
F.soft_corr_enable(False)
print("rows pass done:", sz.tolist(), len(idx))
