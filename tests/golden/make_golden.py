"""Generate the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref, built from
/root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures (inputs + the reference's own outputs; the reference is the unmodified scalar build,
i.e. gravity_kernel.hpp's non-PIKG branch incl. its `tr = xx+yy+xx`):
  init3000_g64.npz   config 1: /root/reference/sample/INIT3000.dat at t=0, interaction lists
                     from the reference's FDPS tree (theta=0.5, n_leaf_limit=8, n_group_limit=64,
                     sample/parameter.dat), force = reference functors on every walk.
  disk2k_g256.npz    2000-particle annulus, n_group_limit=256 (GPU-sized groups).
  corr_long.npz      changeover correction: a crowded 1500-particle annulus (most particles have
                     > 2 candidates), lists + tree force from the reference's tree, and the results
                     of the reference's correctForceLong AND correctForceLongInitial
                     (src/gravity_soft.h:245-528) per particle in original order + neighbour lists.
  groups.npz         single functor calls incl. edge cases (ni=1, nj=0, multi-rank, eps2>0,
                     pre-loaded force for accumulate semantics).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_api as O  # noqa: E402
import synth  # noqa: E402
from gplum_b200 import disk, structs as S  # noqa: E402

REF_SAMPLE = "/root/reference/sample/INIT3000.dat"


def walks_fixture(path, pos, vel, mass, n_group_limit, eps2=0.0):
    r_out, r_search = disk.cutoff_radii(pos, vel, mass)
    w, f_tree = O.ref_tree_walks(pos, mass, r_out, r_search, theta=0.5, n_leaf_limit=8,
                                 n_group_limit=n_group_limit, eps2=eps2, vel=vel, with_force=True)
    f_ref, _ = O.calc_walks(w, eps2, lib="scalar")
    assert f_ref.tobytes() == f_tree.tobytes()
    z = {k: getattr(w, k) for k in ("epi", "epi_off", "ni", "adr_epj", "epj_disp", "n_epj", "adr_spj",
                                    "spj_disp", "n_spj", "epj_all", "spj_all")}
    np.savez_compressed(path, force_ref=f_ref, eps2=np.float32(eps2), **z)
    print(path, "walks", w.n_walk, "interactions", w.n_interactions(),
          "with candidates", int((f_ref["number"] > 0).sum()), "%.0f kB" % (os.path.getsize(path) / 1e3))


def corr_fixture(path, n=1500, seed=2):
    d = disk.make_disk(n, a_in=0.995, a_out=1.005, seed=seed)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    ro, rs = ro * 5.0, rs * 6.0
    rng = np.random.default_rng(seed + 100)
    acc_d = rng.normal(size=(n, 3)) * 1e-3
    ids = rng.permutation(n).astype(np.int64) * 3 + 11
    z = {}
    for initial in (0, 1):
        prm = S.corr_params(initial=bool(initial))
        w, f_tree, ref, lists = O.ref_correct_long(d["pos"], d["vel"], acc_d, d["mass"], ro, rs, ids, prm, n_group_limit=32)
        tag = "init_" if initial else "long_"
        for k, v in ref.items():
            z[tag + k] = v
        z[tag + "ngb"] = np.concatenate(lists) if len(lists) else np.zeros((0, 3), np.int64)
        z[tag + "prm"] = prm
    for k in ("epi", "epi_off", "ni", "adr_epj", "epj_disp", "n_epj", "adr_spj", "spj_disp", "n_spj", "epj_all", "spj_all"):
        z[k] = getattr(w, k)
    np.savez_compressed(path, force_ref=f_tree, **z)
    print(path, "walks", w.n_walk, "neighbours", int(ref["number"].sum()), "with >2 candidates",
          int((f_tree["number"] > 2).sum()), "%.0f kB" % (os.path.getsize(path) / 1e3))


def iso_fixture(path):
    """velKick + the isolated-particle Kepler drift by the reference's own functions (ref_shim)."""
    import iso_cases
    c = iso_cases.make_case()
    prm = O.iso_params()
    v1 = O.vel_kick(c["vel"], c["acc"], float(prm["dt_tree"][0]), lib="scalar")
    pos, vel, time, dt, star, handled = O.kepler_isolated(c["pos"], v1, c["time"], c["dt"], c["acc0"], c["isolated"],
                                                          c["t0"], c["t1"], prm, lib="scalar")
    np.savez_compressed(path, prm=prm, vel_kicked=v1, pos=pos, vel=vel, time=time, dt=dt, star=star, handled=handled,
                        **{"in_" + k: np.asarray(v) for k, v in c.items()})
    print(path, "handled", int(handled.sum()), "of", len(handled), "%.0f kB" % (os.path.getsize(path) / 1e3))


def snapshot_fixture(path, keep=96):
    """Files written by the reference program itself (oracle/_ref/gplum_ref.out on config 1, 4 steps):
    the binary restart file and the ASCII snapshot of the same instant, truncated to `keep` particles."""
    import shutil
    import tempfile
    import gplum_run
    d = tempfile.mkdtemp(prefix="gplum_snap_")
    try:
        gplum_run.run("gplum_ref.out", d, t_end="2^-4", dt_snap="2^-5", threads=4)
        raw = open(os.path.join(d, "TEST", "snap_tmp.dat"), "rb").read()
        txt = open(os.path.join(d, "TEST", "snap000002.dat"), "rb").read().split(b"\n")
    finally:
        shutil.rmtree(d, ignore_errors=True)
    np.savez_compressed(path, binary=np.frombuffer(raw[:128 + 344 * keep], dtype=np.uint8),
                        ascii=np.frombuffer(b"\n".join(txt[:1 + keep]) + b"\n", dtype=np.uint8), keep=np.int32(keep),
                        n_body=np.int32((len(raw) - 128) // 344))
    print(path, "%.0f kB" % (os.path.getsize(path) / 1e3))


def main():
    assert O.have_ref("scalar"), "build oracle/_ref first: make -C oracle ref"
    d = np.loadtxt(REF_SAMPLE, skiprows=1)
    assert d.shape == (3000, 12)
    walks_fixture(os.path.join(HERE, "init3000_g64.npz"), d[:, 4:7].copy(), d[:, 7:10].copy(), d[:, 1].copy(), 64)
    dk = disk.make_disk(2000, a_in=0.98, a_out=1.02, seed=7)
    walks_fixture(os.path.join(HERE, "disk2k_g256.npz"), dk["pos"], dk["vel"], dk["mass"], 256)

    corr_fixture(os.path.join(HERE, "corr_long.npz"))
    iso_fixture(os.path.join(HERE, "iso_step.npz"))
    snapshot_fixture(os.path.join(HERE, "snapshot_ref.npz"))

    cases = {}
    specs = [(1, 1, 1, 0, 0.0, 1), (24, 157, 166, 1, 0.0, 1), (64, 301, 200, 2, 0.0, 2),
             (31, 123, 60, 3, 1e-8, 3), (403, 739, 228, 4, 0.0, 1), (5, 0, 0, 5, 0.0, 1),
             (17, 40, 0, 6, 0.0, 1), (1, 513, 7, 7, 0.0, 1), (130, 33, 1, 8, 0.0, 4)]
    for k, (ni, nj, ns, seed, eps2, n_rank) in enumerate(specs):
        epi, epj, spj = synth.make_group(ni, nj, ns, seed=seed, n_rank=n_rank, dup_self=nj >= ni)
        f0 = S.cleared_force(ni)
        if k % 3 == 2:   # accumulate semantics: non-cleared input force
            f0["acc"] = 0.25; f0["phi"] = -1.0; f0["number"] = 2; f0["id_max"] = 77; f0["id_min"] = 3
        f1 = O.epep(epi, epj, eps2, force=f0, lib="scalar")
        f2 = O.epsp(epi, spj, eps2, force=f1, lib="scalar")
        for nm, a in (("epi", epi), ("epj", epj), ("spj", spj), ("f0", f0), ("f_epep", f1), ("f_both", f2)):
            cases["c%d_%s" % (k, nm)] = a
        cases["c%d_eps2" % k] = np.float32(eps2)
    cases["n_cases"] = np.int32(len(specs))
    p = os.path.join(HERE, "groups.npz")
    np.savez_compressed(p, **cases)
    print(p, "%.0f kB" % (os.path.getsize(p) / 1e3))


if __name__ == "__main__":
    main()
