"""Host-side interaction-list builder (csrc/let_tree.cpp, libgplum_lists.so): the guarantees the force pass
relies on and identity, list for list, with the reference's FDPS tree.  CPU only."""
import numpy as np
import pytest

import oracle_api as O
from gplum_b200 import disk, structs as S, tree


def _disk(n, seed=0, a_in=0.95, a_out=1.05):
    d = disk.make_disk(n, a_in=a_in, a_out=a_out, seed=seed)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    return d, ro, rs


def test_lists_cover_all_mass_exactly_once():
    d, ro, rs = _disk(4000, seed=2)
    w, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=64)
    assert sorted(order.tolist()) == list(range(4000))
    assert w.ni.sum() == 4000 and (w.epi_off == np.concatenate([[0], np.cumsum(w.ni)[:-1]])).all()
    mtot = d["mass"].sum()
    for k in range(0, w.n_walk, 5):
        me = w.epj_all["mass"][w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]]].sum()
        ms = w.spj_all["mass"][w.adr_spj[w.spj_disp[k]:w.spj_disp[k] + w.n_spj[k]]].sum()
        assert abs(me + ms - mtot) < 1e-12 * mtot
        assert len(set(w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]].tolist())) == w.n_epj[k]


def test_symmetric_search_superset_property():
    """Every j within 1.1*max(r_search_i, r_search_j) of an i of the group is in its EP list."""
    d, ro, rs = _disk(3000, seed=3, a_in=0.99, a_out=1.01)
    rs = rs * 3.0          # make the search radius matter
    w, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=32)
    pj = w.epj_all["pos"]; rj = w.epj_all["r_search"]
    for k in range(w.n_walk):
        sl = slice(w.epi_off[k], w.epi_off[k] + w.ni[k])
        pi = w.epi["pos"][sl]; ri = w.epi["r_search"][sl]
        d2 = ((pi[:, None, :] - pj[None, :, :]) ** 2).sum(-1)
        lim = 1.1 * np.maximum(ri[:, None], rj[None, :])
        need = np.nonzero((d2 <= lim * lim).any(0))[0]
        have = set(w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]].tolist())
        assert set(need.tolist()) <= have, k


def test_tree_force_close_to_direct_sum():
    d, ro, rs = _disk(2000, seed=4)
    w, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=64)
    f_tree, _ = O.calc_walks(w, 0.0)
    n = 2000
    direct = O.Walks(w.epi, [0], [n], np.arange(n), [0], [n], np.zeros(0, np.int32), [0], [0], w.epj_all, w.spj_all)
    f_dir, _ = O.calc_walks(direct, 0.0)
    da = np.linalg.norm(f_tree["acc"] - f_dir["acc"], axis=1) / np.linalg.norm(f_dir["acc"], axis=1)
    assert np.median(da) < 2e-3 and da.max() < 0.05
    assert np.abs(f_tree["phi"] / f_dir["phi"] - 1).max() < 1e-3
    for k in ("number", "id_max", "id_min"):
        assert (f_tree[k] == f_dir[k]).all()


def test_quadrupole_moments_of_cells():
    d, ro, rs = _disk(500, seed=5)
    w, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=8)
    root = w.spj_all[0]
    m = d["mass"]; x = d["pos"]
    com = (m[:, None] * x).sum(0) / m.sum()
    dx = x - com
    q = np.array([(m * dx[:, a] * dx[:, b]).sum() for a, b in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))])
    assert abs(root["mass"] - m.sum()) < 1e-15 and np.allclose(root["pos"], com, rtol=0, atol=1e-12)
    assert np.allclose(root["quad"], q, rtol=1e-9, atol=1e-22)


def _init3000():
    """config 1's particles from the committed golden fixture (written by the compiled reference from
    sample/INIT3000.dat): positions, masses and radii of the 3000 planetesimals, in FDPS's tree order."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init3000_g64.npz"))
    e = z["epj_all"]
    o = np.argsort(e["id_local"])
    return e["pos"][o], e["mass"][o], e["r_out"][o], e["r_search"][o], z


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["init3000_g64", "disk20k_g64", "disk20k_g512", "disk20k_wide_rs", "disk5k_leaf4_theta03",
                                  "disk3k_clumps_below_level_21"])
def test_lists_identical_to_reference_fdps_tree(case):
    """The host builder IS FDPS's tree, list for list: same particle order, same i-groups, the same EP and SP
    index lists in the same order, the same cell numbering, and the same EPI / SPJ records bit for bit as the
    reference's own TreeForForce (oracle/ref_shim.cpp: ref_tree_build drives calcForceAllAndWriteBackMultiWalkIndex
    of the unmodified FDPS and records what it hands to the dispatch functor)."""
    kw = dict(theta=0.5, n_leaf_limit=8, n_group_limit=64)
    if case == "init3000_g64":
        pos, mass, ro, rs, _ = _init3000()
        vel = np.zeros_like(pos)
    else:
        n = 5000 if case.startswith("disk5k") else 20000
        d, ro, rs = _disk(n, seed=6, a_in=0.9, a_out=1.1)
        pos, mass, vel = d["pos"], d["mass"], d["vel"]
        if case == "disk20k_g512":
            kw["n_group_limit"] = 512
        if case == "disk20k_wide_rs":
            rs = rs * 3.0
        if case == "disk5k_leaf4_theta03":
            kw.update(theta=0.3, n_leaf_limit=4, n_group_limit=16)
        if case == "disk3k_clumps_below_level_21":
            # distinct particles closer than 2^-21 of the root edge: ordered and split by the LOWER word of FDPS's key
            rng = np.random.default_rng(13)
            pos, mass, vel, ro, rs = pos[:3000].copy(), mass[:3000], vel[:3000], ro[:3000], rs[:3000]
            pos[100:130] = pos[100] + (rng.random((30, 3)) - 0.5) * 2e-9
            pos[500:512] = pos[500] + (rng.random((12, 3)) - 0.5) * 1e-11
            kw["n_group_limit"] = 32
    w, order = tree.build_walks(pos, mass, ro, rs, **kw)
    wr = O.ref_tree_walks(pos, mass, ro, rs, vel=vel, **kw)
    assert w.n_walk == wr.n_walk and w.n_interactions() == wr.n_interactions()
    assert np.array_equal(order, wr.epi["id_local"])
    for k in ("epi_off", "ni", "n_epj", "n_spj", "epj_disp", "spj_disp", "adr_epj", "adr_spj"):
        assert np.array_equal(getattr(w, k), getattr(wr, k)), k
    assert w.epi.tobytes() == wr.epi.tobytes()
    assert w.spj_all.tobytes() == wr.spj_all.tobytes()
    for f in ("id_local", "myrank", "pos", "r_out", "r_search", "mass"):
        assert np.array_equal(w.epj_all[f], wr.epj_all[f]), f
    # hence the oracle's forces on the two list sets are the same bits
    f1, _ = O.calc_walks(w, 0.0); f2, _ = O.calc_walks(wr, 0.0)
    assert f1.tobytes() == f2.tobytes()


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
def test_golden_init3000_lists_are_reproduced():
    """The committed fixture of config 1 (lists recorded from the reference's tree) is what the host builder gives."""
    pos, mass, ro, rs, z = _init3000()
    w, _ = tree.build_walks(pos, mass, ro, rs, theta=0.5, n_leaf_limit=8, n_group_limit=64)
    for k in ("epi_off", "ni", "n_epj", "n_spj", "adr_epj", "adr_spj"):
        assert np.array_equal(getattr(w, k), z[k]), k
    assert w.spj_all.tobytes() == z["spj_all"].tobytes()


def test_lists_library_is_not_the_product_library():
    """The host builder lives in libgplum_lists.so; libgplum_b200.so neither exports nor needs it."""
    import ctypes as C
    from gplum_b200 import _lib
    assert tree.LISTS_PATH != _lib.LIB_PATH
    h = C.CDLL(_lib.LIB_PATH)
    for name in ("gplum_b200_tree_build", "gplum_b200_tree_copy", "gplum_b200_tree_free"):
        assert hasattr(tree.lists_lib(), name) and not hasattr(h, name), name
