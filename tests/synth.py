"""Seeded synthetic inputs for single functor calls (one i-group against j-lists)."""
import numpy as np

from gplum_b200 import structs as S


def make_group(ni, nj, ns, seed=0, box=0.02, r_out=2.0e-3, spread_sp=0.5, n_rank=1, dup_self=True):
    """ni i-particles (also present in the j-list, as FDPS lists contain the group itself),
    nj EP j-particles in a box around (1,0,0) AU, ns quadrupole superparticles further out."""
    rng = np.random.default_rng(seed)
    c = np.array([1.0, 0.02, 0.001])
    epj = np.zeros(nj, dtype=S.EPJ)
    epj["pos"] = c + (rng.random((nj, 3)) - 0.5) * box * np.array([1.0, 1.0, 0.1])
    epj["mass"] = 1e-10 * (0.5 + rng.random(nj))
    epj["r_out"] = r_out * (0.5 + rng.random(nj))
    epj["r_search"] = epj["r_out"] * 1.1 + 1e-4 * rng.random(nj)
    epj["id_local"] = rng.permutation(nj).astype(np.int32) + 5
    epj["myrank"] = rng.integers(0, n_rank, nj).astype(np.int32)
    epj["id"] = np.arange(nj) + 1000
    epj["vel"] = rng.normal(size=(nj, 3))
    epj["acc_d"] = rng.normal(size=(nj, 3))
    epi = np.zeros(ni, dtype=S.EPI)
    if dup_self and nj >= ni:
        sel = rng.choice(nj, ni, replace=False)
        for k in ("pos", "r_out", "r_search", "id_local", "myrank"):
            epi[k] = epj[k][sel]
    else:
        epi["pos"] = c + (rng.random((ni, 3)) - 0.5) * box * np.array([1.0, 1.0, 0.1])
        epi["r_out"] = r_out * (0.5 + rng.random(ni))
        epi["r_search"] = epi["r_out"] * 1.1
        epi["id_local"] = np.arange(ni) + 100000
    spj = np.zeros(ns, dtype=S.SPJ_QUAD)
    d = rng.normal(size=(ns, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    spj["pos"] = c + d * (box + spread_sp * rng.random((ns, 1))) * np.array([1.0, 1.0, 0.05])
    spj["mass"] = 1e-8 * (0.5 + rng.random(ns))
    q = rng.normal(size=(ns, 3, 3)) * 1e-3
    q = np.einsum("nij,nkj->nik", q, q) * spj["mass"][:, None, None]
    spj["quad"] = np.stack([q[:, 0, 0], q[:, 1, 1], q[:, 2, 2], q[:, 0, 1], q[:, 0, 2], q[:, 1, 2]], axis=1)
    return epi, epj, spj


COND_K = 8.0      # floor of the per-particle tolerance in units of 2^-24 * sum_j |f_ij| (see assert_force_close)


def assert_force_close(got, want, rtol=1e-4, what="", cond=None):
    """acc / phi within rtol of |acc| and |phi| per particle; neighbour info bit-exact
    (rank compared as rank==0, the only way the reference reads it: src/particle.h:67).

    cond = (sum_j |f_ij|, sum_j |phi_ij|) from oracle_api.calc_walks_abs switches on the conditioning floor the
    large-N tests need: tolerance_i = max(rtol * |acc_i|, COND_K * 2^-24 * sum_j |f_ij|).  Every FP32 evaluation
    of the reference's kernel rounds each pair term to ~2^-24 of its size; the reference's own two CPU variants (DSL
    order vs fallback order of dx^2+dy^2+dz^2) differ by up to 4.5 * 2^-24 * sum|f| per particle = 0.95e-4 of |acc|
    on a 3e5-particle disk (tests/test_oracle_vs_ref.py::test_fp32_noise_floor_of_the_reference_variants).  The
    floor only exceeds rtol * |acc| for particles whose force sum cancels below 1/210 of its terms."""
    import numpy as np
    assert len(got) == len(want)
    an = np.linalg.norm(want["acc"].astype(np.float64), axis=1)
    da = np.linalg.norm(got["acc"].astype(np.float64) - want["acc"].astype(np.float64), axis=1)
    scale = np.maximum(an, 1e-30)
    tol = rtol * scale
    if cond is not None:
        tol = np.maximum(tol, COND_K * 2.0 ** -24 * cond[0])
    assert (da <= tol).all(), "%s acc err / tolerance max %.3f (rel err max %.3e)" % (what, (da / tol).max(), (da / scale).max())
    dp = np.abs(got["phi"].astype(np.float64) - want["phi"].astype(np.float64))
    ps = np.maximum(np.abs(want["phi"].astype(np.float64)), 1e-30)
    tolp = rtol * ps
    if cond is not None:
        tolp = np.maximum(tolp, COND_K * 2.0 ** -24 * cond[1])
    assert (dp <= tolp).all(), "%s phi rel err max %.3e" % (what, (dp / ps).max())
    for k in ("number", "id_max", "id_min"):
        bad = np.nonzero(got[k] != want[k])[0]
        assert len(bad) == 0, "%s %s differs at %s: got %s want %s" % (what, k, bad[:5], got[k][bad[:5]], want[k][bad[:5]])
    assert ((got["rank"] == 0) == (want["rank"] == 0)).all(), what + " rank==0 flag differs"
