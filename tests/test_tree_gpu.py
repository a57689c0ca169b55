"""GPU interaction-list builder (csrc/dev_tree.cu) against the host builder
(csrc/let_tree.cpp, itself checked against the reference's FDPS tree in test_tree.py /
test_oracle_vs_ref.py) and, end to end, against the oracle's forces.

Bar (integer / index work): groups, EP lists and SP lists equal as sets; tree order, EPJ and SPJ
records (FP64 moments) equal bit for bit -- the device code keeps the host's evaluation order and is
compiled without FMA contraction.  Forces from the GPU-built lists: 1e-4 on acc/phi, neighbour
info exact (the list order inside a walk differs, so sums differ in the last bits)."""
import numpy as np
import pytest

import oracle_api as O
import synth
from gplum_b200 import disk, functors as F, structs as S, tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    F.init(0)
    F.set_params(0.0, True, 0)
    F.walks_select(0)
    yield
    F.set_params(0.0, True, 0)


def _disk(n, seed=0, a_in=0.95, a_out=1.05):
    d = disk.make_disk(n, a_in=a_in, a_out=a_out, seed=seed)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    return d, ro, rs


def assert_same_walks(g, h, order_g, order_h):
    assert g.n_walk == h.n_walk
    assert np.array_equal(order_g, order_h)
    for k in ("epi_off", "ni", "n_epj", "n_spj", "epj_disp", "spj_disp"):
        assert np.array_equal(getattr(g, k), getattr(h, k)), k
    assert g.epi.tobytes() == h.epi.tobytes()
    assert g.epj_all.tobytes() == h.epj_all.tobytes()
    assert len(g.spj_all) == len(h.spj_all)
    for f in g.spj_all.dtype.names:
        assert np.array_equal(g.spj_all[f], h.spj_all[f]), "spj." + f      # -0.0 == 0.0 allowed
    for w in range(g.n_walk):
        a = np.sort(g.adr_epj[g.epj_disp[w]:g.epj_disp[w] + g.n_epj[w]])
        b = np.sort(h.adr_epj[h.epj_disp[w]:h.epj_disp[w] + h.n_epj[w]])
        assert np.array_equal(a, b), ("EP list", w)
        a = np.sort(g.adr_spj[g.spj_disp[w]:g.spj_disp[w] + g.n_spj[w]])
        b = np.sort(h.adr_spj[h.spj_disp[w]:h.spj_disp[w] + h.n_spj[w]])
        assert np.array_equal(a, b), ("SP list", w)


@pytest.mark.parametrize("n,group,leaf,theta,rs_scale,seed", [
    (3000, 64, 8, 0.5, 1.0, 0),          # config 1's size and parameters
    (20000, 512, 8, 0.5, 1.0, 1),
    (20000, 64, 8, 0.5, 3.0, 2),         # large search radii: box-overlap openings matter
    (7000, 16, 4, 0.3, 1.0, 3),
    (7000, 8, 16, 0.8, 1.0, 4),          # n_group_limit < n_leaf_limit: leaves become groups
    (5000, 100000, 8, 0.5, 1.0, 5),      # one group: the root
    (7, 64, 8, 0.5, 1.0, 6),             # the root is a leaf
    (1, 64, 8, 0.5, 1.0, 7),
    (9, 4, 2, 0.5, 1.0, 8),
])
def test_lists_equal_host_builder(n, group, leaf, theta, rs_scale, seed):
    d, ro, rs = _disk(n, seed=seed)
    rng = np.random.default_rng(seed)
    mass = d["mass"] * (0.5 + rng.random(n))
    h, oh = tree.build_walks(d["pos"], mass, ro, rs * rs_scale, theta=theta, n_leaf_limit=leaf, n_group_limit=group)
    sz = tree.build_walks_gpu(d["pos"], mass, ro, rs * rs_scale, theta=theta, n_leaf_limit=leaf, n_group_limit=group)
    g, og = tree.copy_walks_gpu(sz)
    assert (int(sz[6]), int(sz[7])) == h.n_interactions() == g.n_interactions()
    assert_same_walks(g, h, og, oh)


def test_coincident_particles_reach_the_deepest_level():
    """More than n_leaf_limit particles at one point: the cell chain runs to level 42 (FDPS's TREE_LEVEL_LIMIT) and ends in a big leaf."""
    d, ro, rs = _disk(2000, seed=11)
    pos = d["pos"].copy()
    pos[100:140] = pos[100]
    h, oh = tree.build_walks(pos, d["mass"], ro, rs, n_group_limit=32)
    sz = tree.build_walks_gpu(pos, d["mass"], ro, rs, n_group_limit=32)
    g, og = tree.copy_walks_gpu(sz)
    assert_same_walks(g, h, og, oh)


def test_particles_closer_than_the_21_level_grid():
    """Clumps of distinct particles closer than 2^-21 of the root edge share the sorted 63-bit key word: their order
    comes from the lower 21 levels of FDPS's 128-bit key (tie_fix_kernel) and their cells from digits re-derived
    from the positions, down to level 42."""
    d, ro, rs = _disk(3000, seed=13)
    rng = np.random.default_rng(13)
    pos = d["pos"].copy()
    pos[100:130] = pos[100] + (rng.random((30, 3)) - 0.5) * 2e-9          # root edge ~2 AU: 2^-21 of it is 1e-6
    pos[500:512] = pos[500] + (rng.random((12, 3)) - 0.5) * 1e-11
    pos[700:703] = pos[700]                                                # and exactly coincident ones
    for group in (32, 4):
        h, oh = tree.build_walks(pos, d["mass"], ro, rs, n_group_limit=group)
        sz = tree.build_walks_gpu(pos, d["mass"], ro, rs, n_group_limit=group)
        g, og = tree.copy_walks_gpu(sz)
        assert_same_walks(g, h, og, oh)
    assert len(h.spj_all) > 3000 * 0.7 + 8 * 20            # the clumps really made deep cell chains


def test_monopole_spj_records():
    d, ro, rs = _disk(4000, seed=12)
    h, oh = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=64, quad=False)
    F.set_params(0.0, False, 0)
    try:
        sz = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs, n_group_limit=64)
        g, og = tree.copy_walks_gpu(sz, quad=False)
        assert_same_walks(g, h, og, oh)
        F.walks_run(repack=False)
        got = F.walks_download(4000)
    finally:
        F.set_params(0.0, True, 0)
    want, _ = O.calc_walks(h, 0.0)
    synth.assert_force_close(got, want, 1e-4, "monopole, GPU lists")


@pytest.mark.parametrize("n,group,rs_scale", [(3000, 64, 1.0), (30000, 512, 2.0), (30000, 64, 1.0)])
def test_force_from_gpu_lists_matches_oracle(n, group, rs_scale):
    """build on the GPU -> force pass -> download, nothing but particles and forces cross PCIe."""
    d, ro, rs = _disk(n, seed=20, a_in=0.98, a_out=1.02)
    h, oh = tree.build_walks(d["pos"], d["mass"], ro, rs * rs_scale, n_group_limit=group)
    want, n_int = O.calc_walks(h, 0.0)
    F.counters(reset=True)
    sz = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs * rs_scale, n_group_limit=group)
    F.walks_run(repack=False)
    got = F.walks_download(n)
    synth.assert_force_close(got, want, 1e-4, "GPU lists")
    launches, n_ee, n_es = F.counters()
    assert n_ee + n_es == n_int and launches >= 10
    t = tree.gpu_build_times()
    assert all(v >= 0 for v in t.values()) and sum(t.values()) > 0
    # a second pass over the resident set gives the same bits (deterministic list order)
    F.walks_run(repack=True)
    again = F.walks_download(n)
    assert got.tobytes() == again.tobytes()
    # and a rebuild reproduces the lists exactly, order included
    g1, _ = tree.copy_walks_gpu(sz)
    sz2 = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs * rs_scale, n_group_limit=group)
    g2, _ = tree.copy_walks_gpu(sz2)
    assert np.array_equal(sz, sz2) and np.array_equal(g1.adr_epj, g2.adr_epj) and np.array_equal(g1.adr_spj, g2.adr_spj)


@pytest.mark.parametrize("pinned", [False, True])
def test_column_form_with_velocities_and_compact_download(pinned):
    """gplum_b200_tree_build_gpu_vel (what include/gravity_tree_b200.hpp calls: positions first, the other columns
    while the GPU sorts) builds the same tree-order records as the EPJGrav form, from pageable and from pinned
    columns; gplum_b200_tree_download_compact returns what gplum_b200_tree_download_original returns."""
    n = 40000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=31)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    rs = rs * 2.0
    raw = np.zeros(n, dtype=S.EPJ)
    raw["id_local"] = np.arange(n); raw["myrank"] = 0; raw["pos"] = d["pos"]; raw["r_out"] = ro; raw["r_search"] = rs
    raw["id"] = np.arange(n); raw["mass"] = d["mass"]; raw["vel"] = d["vel"]
    acc_d = np.random.default_rng(3).normal(size=(n, 3)) * 1e-3
    sz0 = tree.build_walks_gpu_epj(raw, n_group_limit=128)
    g0, o0 = tree.copy_walks_gpu(sz0)
    cols = {"pos": d["pos"], "vel": d["vel"], "mass": d["mass"], "r_out": ro, "r_search": rs}
    keep = []
    if pinned:
        import torch
        for k, v in cols.items():
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).pin_memory()
            keep.append(t); cols[k] = t.numpy()
    for rep_ in range(2):                                   # the second build overwrites columns a build has just read
        sz = tree.build_walks_gpu(cols["pos"], cols["mass"], cols["r_out"], cols["r_search"], n_group_limit=128, vel=cols["vel"])
    g, o = tree.copy_walks_gpu(sz)
    assert np.array_equal(sz, sz0) and np.array_equal(o, o0)
    assert g.epj_all.tobytes() == g0.epj_all.tobytes() and g.epi.tobytes() == g0.epi.tobytes()
    assert np.array_equal(g.adr_epj, g0.adr_epj) and np.array_equal(g.adr_spj, g0.adr_spj)
    F.walks_run(repack=False)
    full = tree.download_original(n)
    comp, n_nb = tree.download_compact(n)
    assert full.tobytes() == comp.tobytes()
    assert n_nb == int((full["number"] > 0).sum()) and 0 < n_nb < n
    # without velocities: the same tree, vel = 0 in the records
    sz1 = tree.build_walks_gpu(cols["pos"], cols["mass"], cols["r_out"], cols["r_search"], n_group_limit=128)
    g1, o1 = tree.copy_walks_gpu(sz1)
    assert np.array_equal(o1, o0) and not g1.epj_all["vel"].any() and np.array_equal(g1.epj_all["pos"], g0.epj_all["pos"])
    # ... and tree_set_motion completes the records for the correction: the EPJGrav form with vel and acc_d
    import ctypes as C
    from gplum_b200._lib import check, lib
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().gplum_b200_tree_set_motion(n, vp(np.ascontiguousarray(d["vel"])), vp(acc_d)))
    g2, _ = tree.copy_walks_gpu(sz1)
    raw["acc_d"] = acc_d
    tree.build_walks_gpu_epj(raw, n_group_limit=128)
    g3, _ = tree.copy_walks_gpu(sz1)
    assert g2.epj_all.tobytes() == g3.epj_all.tobytes() and g3.epj_all["acc_d"].any()


def test_epj_form_and_changeover_correction():
    """EPJGrav records in arbitrary order in, tree force + changeover correction out: the soft-force
    evaluation of one step (calcForceAllAndWriteBack + correctForceLong) without host-side lists."""
    n = 6000
    d = disk.make_disk(n, a_in=0.99, a_out=1.01, seed=5)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    ro, rs = ro * 2.0, rs * 3.0
    rng = np.random.default_rng(5)
    acc_d = rng.normal(size=(n, 3)) * 1e-3
    ids = rng.permutation(n).astype(np.int64) * 5 + 3
    # three absorbed particles of mergers: same position and same id as their targets until MergeParticle removes
    # them (src/collisionA.h:267-277, src/func.h:160-205); the post-pass must take them for the particle itself
    for k in range(3):
        a, b = 10 + k, n - 1 - k
        for key in ("pos", "vel"):
            d[key][b] = d[key][a]
        ro[b], rs[b], acc_d[b], ids[b] = ro[a], rs[a], acc_d[a], ids[a]
    h, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=64)
    h.epj_all["vel"] = d["vel"][order]; h.epj_all["acc_d"] = acc_d[order]; h.epj_all["id"] = ids[order]
    # the unsorted records FDPS would hand over (epj_org_): particle k at slot k
    raw = np.zeros(n, dtype=S.EPJ)
    raw["id_local"] = np.arange(n); raw["myrank"] = 0; raw["pos"] = d["pos"]; raw["r_out"] = ro; raw["r_search"] = rs
    raw["id"] = ids; raw["mass"] = d["mass"]; raw["vel"] = d["vel"]; raw["acc_d"] = acc_d
    prm = S.corr_params()
    want_f, _ = O.calc_walks(h, 0.0)
    oc, oi, on = O.correct_long(h, prm, force=want_f)
    F.soft_corr_enable(True)
    try:
        sz = tree.build_walks_gpu_epj(raw, n_group_limit=64)
        g, og = tree.copy_walks_gpu(sz)
        assert g.epj_all.tobytes() == h.epj_all.tobytes()
        F.walks_run(repack=False)
        got_f = F.walks_download(n)
        F.correct_long_run(prm)
        corr, init, ngb = F.correct_long_download(n)
    finally:
        F.soft_corr_enable(False)
    synth.assert_force_close(got_f, want_f, 1e-4, "epj form")
    from test_soft_corr_gpu import assert_corr_equal
    assert_corr_equal(corr, oc, want_f, None, None, ngb, on)
    assert corr["number"].sum() > 0
    # The same stage the way include/gravity_tree_b200.hpp drives it: columns up (48 B), compact forces down, then
    # velocity, direct acceleration and id of the LISTED particles only, then the correction.
    import ctypes as C
    from gplum_b200._lib import check, lib
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    oc2, on2 = oc, on
    F.soft_corr_enable(True)
    try:
        tree.build_walks_gpu(d["pos"], d["mass"], ro, rs, n_group_limit=64)
        F.walks_run(repack=False)
        f_org, n_listed = tree.download_compact(n)
        idx = np.zeros(n, np.int32); nbw = np.zeros((n, 4), np.int32); acc4 = np.zeros((n, 4), np.float32); cnt = C.c_int(0)
        check(lib().gplum_b200_tree_download_compact(vp(acc4), vp(idx), vp(nbw), n, C.byref(cnt)))
        listed = idx[:cnt.value]
        assert cnt.value == n_listed and set(np.nonzero(f_org["number"] > 0)[0].tolist()) <= set(listed.tolist())
        assert len(listed) < n // 2                      # a fraction of the disk, not all of it
        check(lib().gplum_b200_tree_set_motion_sparse(len(listed), vp(listed), vp(np.ascontiguousarray(d["vel"][listed])),
                                                      vp(np.ascontiguousarray(acc_d[listed])), vp(np.ascontiguousarray(ids[listed]))))
        F.correct_long_run(prm)
        corr2, _, ngb2 = F.correct_long_download(n)
        # the same from whole columns, gathered by the library
        check(lib().gplum_b200_tree_set_motion(n, None, None))
        check(lib().gplum_b200_tree_set_motion_gather(len(listed), vp(listed), vp(np.ascontiguousarray(d["vel"])), vp(acc_d), vp(ids)))
        F.correct_long_run(prm)
        corr3, _, ngb3 = F.correct_long_download(n)
    finally:
        F.soft_corr_enable(False)
    assert f_org[order].tobytes() == got_f.tobytes()
    # particles that are not listed keep the tree's numbering (id = index): their id_cluster, which no pair touches,
    # is their index; everything else -- and every field of the listed ones -- equals the oracle's with the real ids
    unlisted = ~np.isin(order, listed)
    assert (corr2["id_cluster"][unlisted] == order[unlisted]).all() and (corr2["number"][unlisted] == 0).all()
    for c in (corr2, corr3):
        c["id_cluster"][unlisted] = oc2["id_cluster"][unlisted]
    assert_corr_equal(corr2, oc2, want_f, None, None, ngb2, on2)
    assert corr3.tobytes() == corr2.tobytes() and ngb3.tobytes() == ngb2.tobytes()


def test_full_size_properties():
    """BASELINE configs[2] size (N = 1e6, n_group_limit = 512): properties that need no oracle run --
    every walk's EP + SP lists carry the whole disk's mass exactly once, groups partition the particles
    in tree order, counters equal the host builder's."""
    n = 1000000
    d = disk.make_disk(n)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    sz = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs, n_group_limit=512)
    g, og = tree.copy_walks_gpu(sz)
    assert np.array_equal(np.sort(og), np.arange(n))
    assert g.ni.sum() == n and np.array_equal(g.epi_off, np.concatenate([[0], np.cumsum(g.ni)[:-1]]))
    assert g.ni.max() <= 512
    mtot = d["mass"].sum()
    me = np.add.reduceat(g.epj_all["mass"][g.adr_epj], g.epj_disp)
    ms = np.add.reduceat(g.spj_all["mass"][g.adr_spj], g.spj_disp)
    assert np.abs(me + ms - mtot).max() < 1e-11 * mtot
    for w in range(0, g.n_walk, 97):
        e = g.adr_epj[g.epj_disp[w]:g.epj_disp[w] + g.n_epj[w]]
        assert len(np.unique(e)) == len(e)
        # the group's own particles are in its EP list (self-interaction is part of the reference's sum)
        own = np.arange(g.epi_off[w], g.epi_off[w] + g.ni[w])
        assert np.isin(own, e).all()
    h, oh = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
    assert (int(sz[6]), int(sz[7])) == h.n_interactions()
    assert np.array_equal(og, oh) and np.array_equal(g.n_epj, h.n_epj) and np.array_equal(g.n_spj, h.n_spj)
    assert g.spj_all.tobytes() == h.spj_all.tobytes() or all(np.array_equal(g.spj_all[f], h.spj_all[f]) for f in g.spj_all.dtype.names)
