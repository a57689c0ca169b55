"""Pins oracle/soft_corr_oracle.c (the restatement of correctForceLong / correctForceLongInitial,
src/gravity_soft.h:76-372,375-528 + src/cutfunc.h) against the reference's own functions compiled
from /root/reference (oracle/_ref/libgplum_ref_scalar.so: ref_correct_long).  CPU only."""
import numpy as np
import pytest

import oracle_api as O
from gplum_b200 import disk, structs as S

pytestmark = pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")


def crowded_disk(n, seed, rs_scale=1.0, ro_scale=1.0, n_twin=0):
    """A narrow annulus so that many particles have neighbours (some more than two).  n_twin: absorbed particles of
    mergers -- position, velocity and ID of their targets (src/collisionA.h:267-277), as they exist between the hard
    part and MergeParticle."""
    d = disk.make_disk(n, a_in=0.995, a_out=1.005, seed=seed)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    rng = np.random.default_rng(seed + 100)
    acc_d = rng.normal(size=(n, 3)) * 1e-3
    ids = rng.permutation(n).astype(np.int64) * 3 + 11          # id != id_local
    ro, rs = ro * ro_scale, rs * rs_scale
    for k in range(n_twin):
        a, b = 7 * k + 3, n - 1 - k
        for key in ("pos", "vel"):
            d[key][b] = d[key][a]
        ro[b], rs[b], acc_d[b], ids[b] = ro[a], rs[a], acc_d[a], ids[a]
    return d, ro, rs, acc_d, ids


def compare(n, seed, initial, rs_scale, ro_scale, group, n_twin=0):
    d, ro, rs, acc_d, ids = crowded_disk(n, seed, rs_scale, ro_scale, n_twin)
    prm = S.corr_params(initial=initial)
    w, f_tree, ref, ref_lists = O.ref_correct_long(d["pos"], d["vel"], acc_d, d["mass"], ro, rs, ids, prm,
                                                   n_group_limit=group)
    corr, init, ngb = O.correct_long(w, prm, force=f_tree)
    il = corr["id_local"]
    assert sorted(il.tolist()) == list(range(n))
    # tree force + correction, in the reference's write-back order
    acc = f_tree["acc"].astype(np.float64) + corr["acc"]
    phi = f_tree["phi"].astype(np.float64) + corr["phi"]
    assert (ref["acc_before"][il] == f_tree["acc"].astype(np.float64)).all()
    scale = np.linalg.norm(ref["acc"][il], axis=1)[:, None]
    assert (np.abs(acc - ref["acc"][il]) <= 1e-13 * scale).all()
    assert np.allclose(phi, ref["phi"][il], rtol=1e-13, atol=0)
    assert np.allclose(corr["acc0"], ref["acc0"][il], rtol=1e-13, atol=0)
    assert (corr["number"] == ref["number"][il]).all()
    assert (corr["id_cluster"] == ref["id_cluster"][il]).all()
    assert (corr["in_domain"] == ref["in_domain"][il]).all()
    for k in range(len(corr)):
        mine = ngb[corr["ngb_off"][k]:corr["ngb_off"][k] + corr["number"][k]]
        theirs = ref_lists[il[k]]
        assert sorted(zip(mine["id"].tolist(), mine["rank"].tolist(), mine["id_local"].tolist())) == \
            sorted(map(tuple, theirs.tolist()))
    if initial:
        for key in ("acc_d", "jerk_d"):
            s = np.maximum(np.abs(ref[key][il]).max(axis=1), 1e-300)[:, None]
            assert (np.abs(init[key] - ref[key][il]) <= 1e-12 * s).all(), key
        assert np.allclose(init["phi_d"], ref["phi_d"][il], rtol=1e-12, atol=0)
    return corr, f_tree


@pytest.mark.parametrize("initial", [False, True])
def test_oracle_matches_reference_correct_force_long(initial):
    corr, f = compare(3000, 1, initial, rs_scale=1.0, ro_scale=1.0, group=64)
    assert (corr["number"] > 0).sum() > 10           # the case exercises real neighbours


@pytest.mark.parametrize("initial", [False, True])
def test_oracle_matches_reference_many_neighbours(initial):
    """Large radii: most particles have > 2 candidates, i.e. the reference takes its tree-search branch
    (src/gravity_soft.h:295-303) and the changeover region r < r_out is populated."""
    corr, f = compare(1500, 2, initial, rs_scale=6.0, ro_scale=5.0, group=32)
    assert (f["number"] > 2).sum() > 100
    assert (np.abs(corr["acc"]).sum(axis=1) > 0).sum() > 50


@pytest.mark.parametrize("rs_scale,ro_scale,group", [(1.0, 1.0, 64), (6.0, 5.0, 32)])
def test_absorbed_particles_of_mergers(rs_scale, ro_scale, group):
    """20 twins.  Both of the reference's branches treat an entry with the particle's own id as the particle itself,
    but only the at-most-two-candidates branch adds the twin's m / r_out to phi (src/gravity_soft.h:295-317,105-108);
    the second case puts most twins on the tree-search branch."""
    n = 3000 if rs_scale == 1.0 else 1500
    corr, f = compare(n, 1 if rs_scale == 1.0 else 2, False, rs_scale=rs_scale, ro_scale=ro_scale, group=group, n_twin=20)
    assert (f["number"] > 0).sum() >= 40 and np.isfinite(corr["acc"]).all() and np.isfinite(corr["phi"]).all()
