"""GPU parity of the isolated-particle half of a soft step (csrc/iso_step.cu through the C ABI) against
the oracle restatement (oracle/iso_step_oracle.c, bit-equal to the compiled reference) and the fixture
the reference's own velKick / timeIntegrateKepler_isolated produced.

Bar.  The kick is two FP64 operations per component in the reference's order: bit-exact.  The drift is
FP64 with the reference's evaluation order (-fmad=false); only sin / cos / atan2 come from CUDA's libm
instead of glibc (<= 2 ulp), so pos / vel / phi_s / acc_s / jerk_s are held to 1e-13 relative to the
vector norm -- seven orders inside north_star's 1e-6 for FP64 quantities.  Which particles take the
Kepler branch, `time`, and the power-of-two `dt` must be identical."""
import os

import numpy as np
import pytest

import iso_cases
import oracle_api as O
from gplum_b200 import disk, functors as F, state as ST, structs as S, tree

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iso_step.npz")


@pytest.fixture(scope="module", autouse=True)
def _init():
    F.init(0)
    F.set_params(0.0, True, 0)
    F.walks_select(0)
    yield
    F.soft_corr_enable(False)


def vec_close(got, want, rtol, what):
    s = np.maximum(np.linalg.norm(want, axis=-1, keepdims=True), 1e-300)
    err = (np.abs(got - want) / s).max()
    assert err <= rtol, "%s rel err %.3e" % (what, err)


def run_drift(c, prm, vel):
    n = len(c["pos"])
    epj = ST.make_epj(c["pos"], vel, np.full(n, 1e-10), np.full(n, 1e-3), np.full(n, 1.2e-3))
    ST.upload(epj, c["time"], c["dt"])
    ST.drift(ST.iso_params(), c["t0"], c["t1"], isolated=c["isolated"], acc0=c["acc0"])
    return ST.download(n)


def check_drift(got, want_pos, want_vel, want_time, want_dt, want_star, want_handled, c):
    epj, time, dt, star, handled = got
    assert np.array_equal(handled, want_handled)
    h = handled == 1
    vec_close(epj["pos"][h], want_pos[h], 1e-13, "pos")
    vec_close(epj["vel"][h], want_vel[h], 1e-13, "vel")
    assert np.array_equal(time, want_time)
    assert np.array_equal(dt, want_dt)
    assert np.allclose(star["phi_s"][h], want_star["phi_s"][h], rtol=1e-13, atol=0)
    vec_close(star["acc_s"][h], want_star["acc_s"][h], 1e-13, "acc_s")
    vec_close(star["jerk_s"][h], want_star["jerk_s"][h], 1e-12, "jerk_s")
    assert np.array_equal(star["dt"][h], want_dt[h])
    # untouched: everything the reference integrates elsewhere
    assert np.array_equal(epj["pos"][~h], c["pos"][~h]) and (epj["acc_d"][h] == 0).all()


def test_drift_vs_reference_fixture():
    z = np.load(GOLD)
    c = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    c["t0"], c["t1"] = float(c["t0"]), float(c["t1"])
    got = run_drift(c, z["prm"], z["vel_kicked"])
    check_drift(got, z["pos"], z["vel"], z["time"], z["dt"], z["star"], z["handled"], c)
    assert got[0]["vel"][z["handled"] == 0].tobytes() == z["vel_kicked"][z["handled"] == 0].tobytes()


@pytest.mark.parametrize("n,seed,t0", [(3000, 1, 0.0), (50000, 2, 1.0), (1, 3, 5.5), (257, 4, 123.984375)])
def test_drift_vs_oracle(n, seed, t0):
    c = iso_cases.make_case(n=max(n, 40), seed=seed, t0=t0)
    c = {k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in c.items()}
    prm = O.iso_params()
    want = O.kepler_isolated(c["pos"], c["vel"], c["time"], c["dt"], c["acc0"], c["isolated"], c["t0"], c["t1"], prm)
    got = run_drift(c, prm, c["vel"])
    check_drift(got, *want, c)


def test_edge_orbits():
    """Exactly circular (the ecc == 0 branch), free fall (ecc = 1), unbound, a particle with neighbours, and a
    softened Sun (nobody takes the Kepler branch)."""
    pos = np.array([[1.0, 0, 0], [0, 2.0, 0], [1.0, 0, 0], [0.7, 0.1, 0.01], [1.0, 0.5, 0.0]])
    vel = np.array([[0, 1.0, 0], [-np.sqrt(0.5), 0, 0], [0, 0, 0], [0, 2.5, 0], [0.1, 0.9, 0.02]])
    n = len(pos)
    c = {"pos": pos, "vel": vel, "time": np.zeros(n), "dt": np.array([0.0, 2.0 ** -9, 0.0, 0.0, 2.0 ** -12]),
         "acc0": np.full(n, 1e-5), "isolated": np.array([1, 1, 1, 1, 0], np.int32), "t0": 0.0, "t1": 2.0 ** -6}
    for eps2_sun in (0.0, 1e-8):
        want = O.kepler_isolated(pos, vel, c["time"], c["dt"], c["acc0"], c["isolated"], c["t0"], c["t1"],
                                 O.iso_params(eps2_sun=eps2_sun))
        epj = ST.make_epj(pos, vel, np.full(n, 1e-10), np.full(n, 1e-3), np.full(n, 1.2e-3))
        ST.upload(epj, c["time"], c["dt"])
        ST.drift(ST.iso_params(eps2_sun=eps2_sun), c["t0"], c["t1"], isolated=c["isolated"], acc0=c["acc0"])
        got = ST.download(n)
        assert got[4].tolist() == ([1, 1, 0, 0, 0] if eps2_sun == 0.0 else [0] * 5)
        if eps2_sun == 0.0:
            check_drift(got, *want, c)
        else:
            assert np.array_equal(got[0]["pos"], pos) and np.array_equal(got[0]["vel"], vel)


def test_whole_resident_step_matches_the_host_sequence():
    """kick - drift - tree - force - correction - kick on the resident state against the same sequence
    assembled from parts that are pinned individually: oracle kick / drift on the host, the GPU force +
    correction of the drifted particles (themselves checked in test_tree_gpu / test_soft_corr_gpu)."""
    n = 20000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=8)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    ro, rs = ro * 2.0, rs * 3.0
    ids = np.random.default_rng(8).permutation(n).astype(np.int64) + 7
    prm_c, prm_i = S.corr_params(), ST.iso_params()
    dt_tree = float(prm_i["dt_tree"][0])
    epj0 = ST.make_epj(d["pos"], d["vel"], d["mass"], ro, rs, ids=ids)
    F.soft_corr_enable(True)
    try:
        def force_and_corr():
            sz = ST.tree_build(n_group_limit=64)
            F.walks_run(repack=False)
            F.correct_long_run(prm_c)
            f = F.walks_download(n)
            corr, _, ngb = F.correct_long_download(n)
            return f, corr
        ST.upload(epj0, np.zeros(n), np.zeros(n))
        f0, c0 = force_and_corr()
        # ---- the step on the device
        ST.kick(dt_tree)
        ST.drift(prm_i, 0.0, dt_tree)
        f1, c1 = force_and_corr()
        ST.kick(dt_tree)
        epj, time, dt, star, handled = ST.download(n)
        rec, idx = ST.pull_unhandled(n)
    finally:
        F.soft_corr_enable(False)
    # ---- the same step from its parts (particle order through corr.id_local)
    def acc_of(f, c):
        a = np.zeros((n, 3)); a[c["id_local"]] = f["acc"].astype(np.float64) + c["acc"]
        return a
    iso = np.zeros(n, np.int32); iso[c0["id_local"]] = c0["number"] == 0
    a0 = np.zeros(n); a0[c0["id_local"]] = c0["acc0"]
    v = O.vel_kick(d["vel"], acc_of(f0, c0), dt_tree)
    pos, v, t_, dt_, star_, handled_ = O.kepler_isolated(d["pos"], v, np.zeros(n), np.zeros(n), a0, iso, 0.0, dt_tree, O.iso_params())
    assert np.array_equal(handled, handled_) and 0 < handled.sum() < n
    h = handled == 1
    vec_close(epj["pos"][h], pos[h], 1e-13, "pos after drift")
    assert np.array_equal(epj["pos"][~h], d["pos"][~h])
    # second kick: exact given the device's own second force (bit-exact FP64 update)
    v2 = O.vel_kick(v, acc_of(f1, c1), dt_tree)
    vec_close(epj["vel"], v2, 1e-13, "vel after the step")
    nh = ~h
    assert epj["vel"][nh].tobytes() == O.vel_kick(O.vel_kick(d["vel"], acc_of(f0, c0), dt_tree)[nh], acc_of(f1, c1)[nh], dt_tree).tobytes()
    # the host's share: exactly the particles the drift left alone, records = current state
    assert sorted(idx.tolist()) == np.nonzero(nh)[0].tolist()
    assert rec.tobytes() == epj[idx].tobytes()
    # push them back changed and see the state change
    rec2 = rec.copy(); rec2["pos"] += 1.0
    ST.push(rec2, idx)
    epj3, *_ = ST.download(n)
    assert np.array_equal(epj3["pos"][idx], rec2["pos"]) and np.array_equal(epj3["pos"][h], epj["pos"][h])
