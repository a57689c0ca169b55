"""Forces of a pass over GPU-built lists against the oracle at a bench configuration, with the margins spelled out:
python tools/parity_probe.py <n> <a_in> <a_out> [stride]   (test infrastructure: reads oracle/)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as O
from gplum_b200 import disk, functors as F, tree

n, a_in, a_out = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 50
d = disk.make_disk(n, a_in=a_in, a_out=a_out)
ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
w, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
F.init(0); F.set_params(0.0, True, 0)
idx = np.arange(0, w.n_walk, stride)
s = O.Walks(w.epi, w.epi_off[idx], w.ni[idx], w.adr_epj, w.epj_disp[idx], w.n_epj[idx], w.adr_spj, w.spj_disp[idx], w.n_spj[idx], w.epj_all, w.spj_all)
want, _ = O.calc_walks(s, 0.0, n_threads=0)
sa, sp = O.calc_walks_abs(s, 0.0)
sel = np.concatenate([np.arange(w.epi_off[k], w.epi_off[k] + w.ni[k]) for k in idx])
for name in ("host lists", "GPU lists"):
    if name == "host lists":
        F.walks_upload(w); F.walks_run(repack=True)
    else:
        sz = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs, n_group_limit=512)
        assert (int(sz[6]), int(sz[7])) == w.n_interactions()
        F.walks_run(repack=False)
    got = F.walks_download(n)[sel]
    an = np.linalg.norm(want["acc"][sel].astype(np.float64), axis=1)
    da = np.linalg.norm(got["acc"].astype(np.float64) - want["acc"][sel], axis=1)
    unit = 2.0 ** -24 * sa[sel]
    tol = np.maximum(1e-4 * an, 8 * unit)
    dp = np.abs(got["phi"].astype(np.float64) - want["phi"][sel]); ps = np.abs(want["phi"][sel].astype(np.float64))
    tolp = np.maximum(1e-4 * ps, 8 * 2.0 ** -24 * sp[sel])
    ints = all(np.array_equal(got[k], want[k][sel]) for k in ("number", "id_max", "id_min"))
    print("%s: %d particles of %d walks; acc rel err max %.3e, 99.99%% %.3e, above 1e-4: %d; err in units of 2^-24 sum|f|: max %.2f, 99.99%% %.2f; "
          "err/tol max %.3f; phi err/tol max %.3f; neighbour ints equal: %s"
          % (name, len(sel), len(idx), (da / an).max(), np.quantile(da / an, 0.9999), int((da / an > 1e-4).sum()),
             (da / unit).max(), np.quantile(da / unit, 0.9999), (da / tol).max(), (dp / tolp).max(), ints))
