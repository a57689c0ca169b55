"""Pins oracle/iso_step_oracle.c (velKick, Kepler drift of isolated particles, calcStarGravity,
calcDeltatInitial: src/particle.h:878-914, src/kepler.h, src/hermite.h:787-816, src/gravity_hard.h:5-39,
src/hard.h:793-817) against the reference's own functions (oracle/_ref: ref_vel_kick,
ref_kepler_isolated) and against the committed fixture those functions produced.  CPU only.
Same libm, same evaluation order, no FMA contraction: the bar is bit equality."""
import os

import numpy as np
import pytest

import iso_cases
import oracle_api as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iso_step.npz")


def run_oracle(c, prm):
    v1 = O.vel_kick(c["vel"], c["acc"], float(prm["dt_tree"][0]))
    return (v1,) + O.kepler_isolated(c["pos"], v1, c["time"], c["dt"], c["acc0"], c["isolated"], c["t0"], c["t1"], prm)


def test_oracle_reproduces_the_reference_fixture():
    z = np.load(GOLD)
    c = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    c["t0"], c["t1"] = float(c["t0"]), float(c["t1"])
    v1, pos, vel, time, dt, star, handled = run_oracle(c, z["prm"])
    assert np.array_equal(handled, z["handled"])
    assert 0 < handled.sum() < len(handled)
    # branch coverage of the fixture: neighbours, e >= 0.8, unbound, circular, both dt branches
    ecc_hi = (c["isolated"] == 1) & (handled == 0)
    assert ecc_hi.sum() > 5 and (c["isolated"] == 0).sum() > 100
    for name, got in (("vel_kicked", v1), ("pos", pos), ("vel", vel), ("time", time), ("dt", dt)):
        assert got.tobytes() == z[name].tobytes(), name
    assert star.tobytes() == z["star"].tobytes()


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,t0", [(1, 0.0), (2, 1.0), (5, 37.015625)])
def test_oracle_equals_compiled_reference(seed, t0):
    c = iso_cases.make_case(n=3000, seed=seed, t0=t0)
    prm = O.iso_params()
    v1, pos, vel, time, dt, star, handled = run_oracle(c, prm)
    r_v1 = O.vel_kick(c["vel"], c["acc"], float(prm["dt_tree"][0]), lib="scalar")
    r = O.kepler_isolated(c["pos"], r_v1, c["time"], c["dt"], c["acc0"], c["isolated"], c["t0"], c["t1"], prm, lib="scalar")
    assert v1.tobytes() == r_v1.tobytes()
    for got, want, name in zip((pos, vel, time, dt, star, handled), r, ("pos", "vel", "time", "dt", "star", "handled")):
        assert got.tobytes() == want.tobytes(), name


def test_kepler_drift_conserves_the_orbit():
    """Size-independent properties: energy and angular momentum of the two-body orbit are kept to
    round-off, and a full period returns the particle to where it was."""
    c = iso_cases.make_case(n=2000, seed=9)
    prm = O.iso_params()
    iso = np.ones(2000, np.int32)
    pos, vel, _, _, _, handled = O.kepler_isolated(c["pos"], c["vel"], c["time"], c["dt"], c["acc0"], iso, c["t0"], c["t1"], prm)
    h = handled == 1
    e0 = 0.5 * (c["vel"] ** 2).sum(1) - 1.0 / np.sqrt((c["pos"] ** 2).sum(1))
    e1 = 0.5 * (vel ** 2).sum(1) - 1.0 / np.sqrt((pos ** 2).sum(1))
    assert np.abs(e1 - e0)[h].max() < 1e-13 * np.abs(e0[h]).max()
    l0, l1 = np.cross(c["pos"], c["vel"]), np.cross(pos, vel)
    assert np.abs(l1 - l0)[h].max() < 1e-13
    ax = -0.5 / e0
    k = np.nonzero(h & (ax > 0))[0][:50]
    for i in k:
        period = 2 * np.pi * ax[i] ** 1.5
        p, v, *_ = O.kepler_isolated(c["pos"][i:i + 1], c["vel"][i:i + 1], [0.0], [0.0], [0.0], [1], 0.0, period, prm)
        assert np.abs(p - c["pos"][i]).max() < 1e-11 and np.abs(v - c["vel"][i]).max() < 1e-11


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
def test_edge_orbits_equal_compiled_reference():
    """Exactly circular (ecc == 0 takes the u = 0 branch of posVel2OrbitalElement), exactly at rest in the
    frame (falls to the Sun: ecc = 1), unbound (ax < 0), softened Sun (eps2_sun != 0: nobody drifts on a
    Kepler orbit), and a particle with neighbours."""
    pos = np.array([[1.0, 0, 0], [0, 2.0, 0], [1.0, 0, 0], [0.7, 0.1, 0.01], [1.0, 0.5, 0.0]])
    vel = np.array([[0, 1.0, 0], [-np.sqrt(0.5), 0, 0], [0, 0, 0], [0, 2.5, 0], [0.1, 0.9, 0.02]])
    n = len(pos)
    args = (np.zeros(n), np.array([0.0, 2.0 ** -9, 0.0, 0.0, 2.0 ** -12]), np.full(n, 1e-5), np.array([1, 1, 1, 1, 0]), 0.0, 2.0 ** -6)
    for eps2_sun in (0.0, 1e-8):
        prm = O.iso_params(eps2_sun=eps2_sun)
        a = O.kepler_isolated(pos, vel, *args, prm)
        b = O.kepler_isolated(pos, vel, *args, prm, lib="scalar")
        for x, y, name in zip(a, b, ("pos", "vel", "time", "dt", "star", "handled")):
            assert x.tobytes() == y.tobytes(), (name, eps2_sun)
        if eps2_sun == 0.0:
            assert a[5].tolist() == [1, 1, 0, 0, 0]
            assert np.allclose(np.sqrt((a[0][0] ** 2).sum()), 1.0, rtol=1e-15)        # stays on the unit circle
        else:
            assert a[5].sum() == 0 and np.array_equal(a[0], pos)
