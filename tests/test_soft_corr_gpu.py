"""GPU parity of the changeover correction (libgplum_b200: soft_corr.cu through the C ABI) against
the oracle restatement and the golden fixture produced by the reference's own
correctForceLong / correctForceLongInitial (src/gravity_soft.h:245-528).

Tolerances.  The correction is FP64 and compiled without FMA contraction, so its terms are the
reference's; sums run over the same neighbours in ascending EP-index order instead of the
reference's tree-search order.  Bar: |d corr| <= 1e-12 * |tree force + corr| per particle (far
inside north_star's 1e-6 for FP64 quantities); neighbour lists, counts, cluster ids: exact as sets."""
import os

import numpy as np
import pytest

import oracle_api as O
from test_oracle_golden import check_corr_against_fixture, load_corr_fixture
from gplum_b200 import disk, functors as F, structs as S, tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    F.init(0)
    F.set_params(0.0, True, 0)
    yield
    F.soft_corr_enable(False)


def assert_corr_equal(got, want, f_tree, got_init=None, want_init=None, got_ngb=None, want_ngb=None, rtol=1e-12):
    gc, wc = got, want
    tot = np.linalg.norm(f_tree["acc"].astype(np.float64) + wc["acc"], axis=1)[:, None]
    assert (np.abs(gc["acc"] - wc["acc"]) <= rtol * tot).all(), np.abs(gc["acc"] - wc["acc"]).max()
    ptot = np.abs(f_tree["phi"].astype(np.float64) + wc["phi"])
    assert (np.abs(gc["phi"] - wc["phi"]) <= rtol * ptot).all()
    assert np.allclose(gc["acc0"], wc["acc0"], rtol=rtol, atol=0)
    for k in ("number", "id_cluster", "in_domain", "id_local"):
        assert (gc[k] == wc[k]).all(), k
    if want_init is not None:
        for key in ("acc_d", "jerk_d"):
            s = np.maximum(np.abs(want_init[key]).max(axis=1), 1e-300)[:, None]
            assert (np.abs(got_init[key] - want_init[key]) <= 1e-10 * s).all(), key
        assert np.allclose(got_init["phi_d"], want_init["phi_d"], rtol=1e-10, atol=0)
    if want_ngb is not None:
        for k in range(len(gc)):
            a = got_ngb[gc["ngb_off"][k]:gc["ngb_off"][k] + gc["number"][k]]
            b = want_ngb[wc["ngb_off"][k]:wc["ngb_off"][k] + wc["number"][k]]
            assert sorted(a.tolist()) == sorted(b.tolist()), k


@pytest.mark.parametrize("tag", ["long_", "init_"])
def test_correction_vs_reference_fixture(tag):
    """config: crowded annulus, most particles with > 2 candidates (the reference's tree-search branch)."""
    w, z = load_corr_fixture()
    prm = z[tag + "prm"]
    initial = tag == "init_"
    F.set_params(0.0, True, F.TRACE_AS_SHIPPED)      # the fixture's tree force is the as-shipped reference's
    try:
        f, corr, init, ngb = F.correctForceLong(w, prm, initial=initial)
    finally:
        F.set_params(0.0, True, 0)
    f_ref = z["force_ref"]
    for k in ("number", "id_max", "id_min"):
        assert (f[k] == f_ref[k]).all()
    # the correction itself against the reference's totals, using the reference's own FP32 tree force
    check_corr_against_fixture(w, z, tag, f_ref, corr, init, ngb, 1e-12)
    # and against the oracle restatement, term by term
    oc, oi, on = O.correct_long(w, prm, force=f_ref)
    assert_corr_equal(corr, oc, f_ref, init, oi, ngb, on)


@pytest.mark.parametrize("group,rs_scale,ro_scale,seed", [(64, 1.0, 1.0, 5), (512, 3.0, 2.0, 6), (16, 8.0, 8.0, 7)])
def test_correction_vs_oracle_on_own_lists(group, rs_scale, ro_scale, seed):
    """Lists from the library's own host builder at several group sizes / crowding levels."""
    n = 6000
    d = disk.make_disk(n, a_in=0.99, a_out=1.01, seed=seed)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, order = tree.build_walks(d["pos"], d["mass"], ro * ro_scale, rs * rs_scale, n_group_limit=group)
    rng = np.random.default_rng(seed)
    w.epj_all["vel"] = d["vel"][order]
    w.epj_all["acc_d"] = rng.normal(size=(n, 3)) * 1e-3
    w.epj_all["id"] = rng.permutation(n).astype(np.int64) * 5 + 3
    for initial in (False, True):
        prm = S.corr_params(initial=initial)
        f, corr, init, ngb = F.correctForceLong(w, prm, initial=initial)
        want_f, _ = O.calc_walks(w, 0.0)
        assert (f["number"] == want_f["number"]).all()
        oc, oi, on = O.correct_long(w, prm, force=want_f)
        assert_corr_equal(corr, oc, want_f, init, oi, ngb, on)
    assert corr["number"].sum() > 0


def test_no_candidates_gives_self_term_only():
    d = disk.make_disk(2000, a_in=0.5, a_out=3.0, seed=9)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, order = tree.build_walks(d["pos"], d["mass"], ro * 0.01, rs * 0.01, n_group_limit=64)
    w.epj_all["id"] = np.arange(2000)
    prm = S.corr_params()
    f, corr, init, ngb = F.correctForceLong(w, prm)
    assert f["number"].sum() == 0 and corr["number"].sum() == 0 and len(ngb) == 0
    i_self = np.empty(2000, dtype=np.int64)
    lut = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(w.epj_all["id_local"], w.epj_all["myrank"]))}
    for k in range(2000):
        i_self[k] = lut[(int(w.epi["id_local"][k]), int(w.epi["myrank"][k]))]
    assert np.array_equal(corr["phi"], w.epj_all["mass"][i_self] * (1.0 / w.epj_all["r_out"][i_self]))
    assert (corr["acc"] == 0).all() and (corr["acc0"] == 0).all()
    assert (corr["id_cluster"] == w.epj_all["id"][i_self]).all() and (corr["in_domain"] == 1).all()


def test_pair_buffer_overflow_is_reported():
    w, z = load_corr_fixture()
    F.soft_corr_enable(True, pair_cap=16)
    try:
        F.calc_walks(w)
        F.correct_long_run(z["long_prm"])
        with pytest.raises(Exception, match="pair"):
            F.correct_long_download(len(w.epi))
    finally:
        F.soft_corr_enable(False)


def test_run_without_capture_fails_loudly():
    w, z = load_corr_fixture()
    F.soft_corr_enable(False)
    F.calc_walks(w)
    with pytest.raises(Exception, match="captured"):
        F.correct_long_run(z["long_prm"])


def test_compact_download_equals_the_filtered_full_download():
    """Only particles with neighbours cross PCIe; the others carry the self term alone (checked here too)."""
    n = 8000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=3)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, order = tree.build_walks(d["pos"], d["mass"], ro, rs * 1.5, n_group_limit=64)
    w.epj_all["id"] = np.arange(n) * 2 + 1
    prm = S.corr_params()
    F.soft_corr_enable(True)
    try:
        F.calc_walks(w)
        F.correct_long_run(prm)
        full, _, ngb = F.correct_long_download(n)
        comp, ngb2 = F.correct_long_download_compact(n)
    finally:
        F.soft_corr_enable(False)
    has = full["number"] > 0
    assert 0 < has.sum() < n
    assert comp.tobytes() == full[has].tobytes() and ngb2.tobytes() == ngb.tobytes()
    rest = full[~has]
    assert (rest["acc"] == 0).all() and (rest["acc0"] == 0).all()
    lut = np.empty(n, np.int64); lut[w.epj_all["id_local"]] = np.arange(n)
    k = lut[rest["id_local"]]
    assert np.array_equal(rest["phi"], w.epj_all["mass"][k] * (1.0 / w.epj_all["r_out"][k]))
    assert np.array_equal(rest["id_cluster"], w.epj_all["id"][k])
    with pytest.raises(Exception, match="records"):
        F.correct_long_download_compact(n, corr_cap=3)
