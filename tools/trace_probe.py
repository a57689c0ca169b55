"""Where and when does every work item of a (sub-wave) force pass run?  gplum_b200_debug_trace records per item
{start, end, SM, hardware warp slot, times run}; this prints, for the first 1/k of the walks of the N=1e6 workload, the per-SM
and per-scheduler (warp slot mod 4) sums of the items' model costs and busy times, and saves the raw trace.
Usage: python tools/trace_probe.py [k ...]"""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gplum_b200 import disk, functors as F, tree
from gplum_b200._lib import check, lib
from gplum_b200.walks import Walks

n = 1000000
d = disk.make_disk(n)
ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
F.init(0); F.set_params(0.0, True, 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for k in [int(a) for a in sys.argv[1:]] or [8]:
    m = w.n_walk // k
    sub = Walks(w.epi, w.epi_off[:m], w.ni[:m], w.adr_epj, w.epj_disp[:m], w.n_epj[:m], w.adr_spj, w.spj_disp[:m],
                w.n_spj[:m], w.epj_all, w.spj_all)
    F.walks_upload(sub)
    for _ in range(5):
        F.walks_run(repack=False)
    ms = F.walks_time(30, repack=False)
    check(lib().gplum_b200_debug_trace(1, None, 0, None))
    for _ in range(3):
        F.walks_run(repack=False)
    cap = 1 << 20
    tr = np.zeros((cap, 4), dtype=np.uint64)
    cnt = C.c_int(0)
    check(lib().gplum_b200_debug_trace(0, tr.ctypes.data_as(C.c_void_p), cap, C.byref(cnt)))
    tr = tr[:cnt.value]
    assert (tr[:, 3] == 1).all(), "every item runs exactly once"
    np.save(os.path.join(ROOT, "gpurun_out", "trace_k%d.npy" % k), tr)
    t0 = tr[:, 0].astype(np.int64); t1 = tr[:, 1].astype(np.int64)
    smid = (tr[:, 2] & np.uint64(0xffffffff)).astype(np.int64); wslot = (tr[:, 2] >> np.uint64(32)).astype(np.int64)
    base = t0.min()
    dur = (t1 - t0) * 1e-3
    print("1/%d: %d walks, %d items, %.4f ms per pass untraced; traced pass spans %.1f us; item duration us min/med/max %.1f/%.1f/%.1f"
          % (k, m, len(tr), ms, (t1.max() - base) * 1e-3, dur.min(), np.median(dur), dur.max()))
    print("   SM ids: %d distinct, max %d; items per SM min/max %d/%d; warp slots seen: %s"
          % (len(np.unique(smid)), smid.max(), np.bincount(smid).min(), np.bincount(smid).max(), np.unique(wslot)[:24]))
    end_sm = np.zeros(smid.max() + 1); busy_sm = np.zeros(smid.max() + 1)
    np.maximum.at(end_sm, smid, (t1 - base) * 1e-3); np.add.at(busy_sm, smid, dur)
    print("   per-SM end time us: min %.1f mean %.1f max %.1f; per-SM sum of item durations: min %.0f mean %.0f max %.0f"
          % (end_sm[end_sm > 0].min(), end_sm[end_sm > 0].mean(), end_sm.max(), busy_sm[busy_sm > 0].min(), busy_sm[busy_sm > 0].mean(), busy_sm.max()))
    print("   start time of items us: min %.1f med %.1f max %.1f" % ((t0 - base).min() * 1e-3, np.median(t0 - base) * 1e-3, (t0 - base).max() * 1e-3))
    # is CTA c on SM c mod 148 ?
    cta_sm = smid[::4][:600]
    print("   SM of the first CTAs:", cta_sm[:20].tolist(), "...", cta_sm[148:158].tolist())
