"""Host-side interaction-list builder (csrc/let_tree.cpp): the guarantees the force pass relies
on and agreement with the reference's FDPS tree.  CPU only."""
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


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
def test_statistics_match_reference_fdps_tree():
    """Same group / list statistics as the reference's tree on the same disk (SURVEY 8a numbers)."""
    d, ro, rs = _disk(20000, seed=6, a_in=0.9, a_out=1.1)
    for g in (64, 512):
        w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=g)
        wr = O.ref_tree_walks(d["pos"], d["mass"], ro, rs, n_group_limit=g, vel=d["vel"])
        assert abs(w.n_walk - wr.n_walk) <= 0.03 * wr.n_walk
        a, b = w.n_interactions(), wr.n_interactions()
        assert abs(a[0] - b[0]) <= 0.03 * b[0] and abs(a[1] - b[1]) <= 0.03 * b[1]
        # and the forces the two list sets produce agree at tree-approximation level
        f1, _ = O.calc_walks(w, 0.0); f2, _ = O.calc_walks(wr, 0.0)
        a1 = np.zeros((20000, 3)); a2 = np.zeros((20000, 3))
        a1[w.epi["id_local"]] = f1["acc"]; a2[wr.epi["id_local"]] = f2["acc"]
        rel = np.linalg.norm(a1 - a2, axis=1) / np.linalg.norm(a2, axis=1)
        assert np.median(rel) < 1e-3
