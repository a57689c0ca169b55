"""Synthetic Kokubo-Ida-style planetesimal disks: the bench / test input generator.

Own numpy implementation of the recipe in the reference's IC generator and cut-off radius
set-up (nothing is imported from the reference):
  surface density, particle mass, a-distribution : /root/reference/src/disk.h:61-87,90-171
  e,i ~ |N(0, sigma)|, sigma = {ecc,inc}_hill * h   : src/disk.h:147-148, src/mathfunc.h:25-49
  Kepler solve + orbital elements -> pos/vel        : src/kepler.h:5-38,80-93
  velocity dispersion in 32 radial bins             : src/particle.h:1320-1437 (calcRandomVel)
  r_out, r_search                                   : src/particle.h:713-733 (setROutRSearch)
Defaults are sample/parameter.dat's (theta etc. live with the tree builder).
Units: G = M_sun = AU = 1.
"""
import numpy as np

L_CGS = 14959787070000.0
M_CGS = 1.9884e33


def _dust_mass(a0, a1, p, f_dust, eta_ice, inside_ice):
    if a1 < a0:
        return 0.0
    coef = 10.0 * f_dust * (1.0 if inside_ice else eta_ice) / M_CGS * L_CGS * L_CGS
    return 2.0 * np.pi * coef / (2.0 - p) * (a1 ** (2.0 - p) - a0 ** (2.0 - p))


def _semimajor(rng, n, a0, a1, p):
    r = rng.random(n)
    if p != 2:
        return ((a1 ** (2.0 - p) - a0 ** (2.0 - p)) * r + a0 ** (2.0 - p)) ** (1.0 / (2.0 - p))
    return np.exp((np.log(a1) - np.log(a0)) * r + np.log(a0))


def _solve_kepler(l, ecc):
    u = l + ecc * np.sin(l)
    for _ in range(12):
        u = u - (u - ecc * np.sin(u) - l) / (1.0 - ecc * np.cos(u))
    return u


def make_disk(n, a_in=0.9, a_out=1.1, seed=0, p=1.5, f_dust=0.71, eta_ice=30.0 / 7.1, a_ice=2.0,
              ecc_hill=2.0, inc_hill=1.0, m_sun=1.0, m_init=0.0):
    """Returns dict(pos[n,3], vel[n,3], mass[n]) for an equal-mass planetesimal disk."""
    rng = np.random.default_rng(seed)
    if a_out < a_ice:
        m_in, m_out = _dust_mass(a_in, a_out, p, f_dust, eta_ice, True), 0.0
    elif a_ice < a_in:
        m_in, m_out = 0.0, _dust_mass(a_in, a_out, p, f_dust, eta_ice, False)
    else:
        m_in = _dust_mass(a_in, a_ice, p, f_dust, eta_ice, True)
        m_out = _dust_mass(a_ice, a_out, p, f_dust, eta_ice, False)
    m = (m_in + m_out) / n if m_init == 0.0 else m_init
    n_in = int(round(m_in / (m_in + m_out) * n))
    if a_out < a_ice or a_ice < a_in:
        ax = _semimajor(rng, n, a_in, a_out, p)
    else:
        ax = np.concatenate([_semimajor(rng, n_in, a_in, a_ice, p), _semimajor(rng, n - n_in, a_ice, a_out, p)])
    h = (m / (3.0 * m_sun)) ** (1.0 / 3.0)
    ecc = np.abs(rng.normal(0.0, ecc_hill * h, n))
    inc = np.abs(rng.normal(0.0, inc_hill * h, n))
    l = 2 * np.pi * rng.random(n)
    u = _solve_kepler(l, ecc)
    omg = 2 * np.pi * rng.random(n)
    OMG = 2 * np.pi * rng.random(n)
    nn = np.sqrt(m_sun / ax ** 3)
    co, so, cO, sO, ci, si = np.cos(omg), np.sin(omg), np.cos(OMG), np.sin(OMG), np.cos(inc), np.sin(inc)
    P = np.stack([co * cO - so * sO * ci, co * sO + so * cO * ci, so * si], axis=1)
    Q = np.stack([-so * cO - co * sO * ci, -so * sO + co * cO * ci, co * si], axis=1)
    cu, su = np.cos(u), np.sin(u)
    esq = np.sqrt(1.0 - ecc * ecc)
    pos = ax[:, None] * ((cu - ecc)[:, None] * P + (esq * su)[:, None] * Q)
    rinv = 1.0 / np.sqrt((pos * pos).sum(1))
    vel = (ax * ax * nn * rinv)[:, None] * (-su[:, None] * P + (esq * cu)[:, None] * Q)
    return {"pos": pos, "vel": vel, "mass": np.full(n, m), "m_sun": m_sun}


def velocity_dispersion(pos, vel, m_sun=1.0, nbin=32):
    """Per-particle v_disp: rms random velocity of the particle's radial bin (calcRandomVel)."""
    r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2)
    r_max, r_min = r.max() * 1.01, r.min() * 0.99
    dr = (r_max - r_min) / nbin
    j = np.minimum(((r - r_min) / dr).astype(np.int64), nbin - 1)
    vk = np.sqrt(m_sun / r)
    v_kep = np.stack([-pos[:, 1] / r * vk, pos[:, 0] / r * vk, np.zeros_like(r)], axis=1)
    v_ran = vel - v_kep
    s = np.bincount(j, weights=(v_ran * v_ran).sum(1), minlength=nbin)
    c = np.bincount(j, minlength=nbin)
    vd = np.where(c > 0, np.sqrt(s / np.maximum(c, 1)), 0.0)
    return vd[j]


def cutoff_radii(pos, vel, mass, m_sun=1.0, dt_tree=2.0 ** -6, R_cut0=3.0, R_cut1=8.0, R_search0=1.1,
                 R_search1=6.0, p_cut=0.0, r_cut_min=0.0, r_cut_max=0.0):
    """(r_out, r_search) per particle, individual cut-off (setROutRSearch)."""
    r = np.sqrt((pos * pos).sum(1))
    v2 = (vel * vel).sum(1)
    rv = (pos * vel).sum(1)
    ax = 1.0 / (2.0 / r - v2 / m_sun)
    ecc = np.sqrt((1.0 - r / ax) ** 2 + rv * rv / (m_sun * ax))
    ax2 = np.where(ecc < 0.6, ax, r)
    r_hill = (mass / (3.0 * m_sun)) ** (1.0 / 3.0) * ax2
    v_disp = velocity_dispersion(pos, vel, m_sun)
    r_out = np.maximum(R_cut0 * ax2 ** (-p_cut) * r_hill, R_cut1 * v_disp * dt_tree)
    r_out = np.maximum(r_out, r_cut_min)
    if r_cut_max > 0.0:
        r_out = np.minimum(r_out, r_cut_max)
    r_search = R_search0 * r_out + R_search1 * v_disp * dt_tree
    return r_out, r_search
