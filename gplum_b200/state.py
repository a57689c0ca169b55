"""Device-resident particle state and the isolated-particle half of a soft step (csrc/iso_step.cu,
SURVEY 8 f3): FPGrav::velKick (src/particle.h:878-884) and the Kepler drift of particles without
neighbours (src/hard.h:793-817, src/hermite.h:787-816, src/kepler.h) over the C ABI."""
import ctypes as C

import numpy as np

from . import structs as S
from ._lib import check, lib

ISO_PARAMS = np.dtype([("m_sun", "<f8"), ("dt_tree", "<f8"), ("eta_0", "<f8"), ("eta_sun0", "<f8"),
                       ("alpha2", "<f8"), ("dt_min", "<f8"), ("eps2_sun", "<f8")])
STAR = np.dtype([("phi_s", "<f8"), ("acc_s", "<f8", (3,)), ("jerk_s", "<f8", (3,)), ("dt", "<f8")])
assert ISO_PARAMS.itemsize == 56 and STAR.itemsize == 64


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def iso_params(m_sun=1.0, dt_tree=2.0 ** -6, eta_0=0.002, eta_sun0=0.002, alpha=1.0, dt_min=2.0 ** -30, eps2_sun=0.0):
    """Defaults are sample/parameter.dat's (lines 52-60)."""
    p = np.zeros(1, dtype=ISO_PARAMS)
    p["m_sun"], p["dt_tree"], p["eta_0"], p["eta_sun0"] = m_sun, dt_tree, eta_0, eta_sun0
    p["alpha2"], p["dt_min"], p["eps2_sun"] = alpha * alpha, dt_min, eps2_sun
    return p


def make_epj(pos, vel, mass, r_out, r_search, ids=None, acc_d=None, rank=0):
    """EPJGrav[n] with particle k at slot k, as FDPS's epj_org_ holds them."""
    n = len(pos)
    e = np.zeros(n, dtype=S.EPJ)
    e["id_local"] = np.arange(n); e["myrank"] = rank
    e["pos"], e["vel"], e["mass"], e["r_out"], e["r_search"] = pos, vel, mass, r_out, r_search
    e["id"] = np.arange(n) if ids is None else ids
    if acc_d is not None:
        e["acc_d"] = acc_d
    return e


def upload(epj, time=None, dt=None):
    epj = np.ascontiguousarray(epj, dtype=S.EPJ)
    t = None if time is None else np.ascontiguousarray(time, dtype=np.float64)
    d = None if dt is None else np.ascontiguousarray(dt, dtype=np.float64)
    check(lib().gplum_b200_state_upload(len(epj), _p(epj), _p(t), _p(d)))


def download(n):
    epj = np.zeros(n, dtype=S.EPJ)
    time, dt = np.zeros(n), np.zeros(n)
    star = np.zeros(n, dtype=STAR)
    handled = np.zeros(n, dtype=np.int32)
    check(lib().gplum_b200_state_download(_p(epj), _p(time), _p(dt), _p(star), _p(handled)))
    return epj, time, dt, star, handled


def tree_build(theta=0.5, n_leaf_limit=8, n_group_limit=64):
    sz = np.zeros(8, dtype=np.int64)
    check(lib().gplum_b200_state_tree_build(float(theta), int(n_leaf_limit), int(n_group_limit), _p(sz)))
    return sz


def kick(dt_tree, slot=0, use_corr=True):
    check(lib().gplum_b200_state_kick(int(slot), int(bool(use_corr)), float(dt_tree)))


def drift(prm, t0, t1, slot=0, isolated=None, acc0=None):
    assert prm.dtype == ISO_PARAMS
    iso = None if isolated is None else np.ascontiguousarray(isolated, dtype=np.int32)
    a0 = None if acc0 is None else np.ascontiguousarray(acc0, dtype=np.float64)
    check(lib().gplum_b200_state_drift(_p(prm), float(t0), float(t1), int(slot), _p(iso), _p(a0)))


def pull_unhandled(cap):
    rec = np.zeros(cap, dtype=S.EPJ)
    idx = np.zeros(cap, dtype=np.int32)
    n = C.c_int(0)
    check(lib().gplum_b200_state_pull_unhandled(_p(rec), _p(idx), int(cap), C.byref(n)))
    return rec[:n.value], idx[:n.value]


def push(rec, idx):
    rec = np.ascontiguousarray(rec, dtype=S.EPJ)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    check(lib().gplum_b200_state_push(_p(rec), _p(idx), len(rec)))
