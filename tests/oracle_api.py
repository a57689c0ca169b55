"""ctypes access to oracle/ (test infrastructure; never imported by the product package).

liboracle.so        : the C restatement (oracle/pikg_oracle.c), built on demand with gcc.
_ref/libgplum_ref_* : the reference's own functors compiled from /root/reference (prebuilt;
                      rebuilt here only when the reference tree is present).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gplum_b200 import structs as S

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(HERE), "oracle")

ORDER_DSL, RANK_SQ, TRACE_FALLBACK = 1, 2, 4
CANONICAL = 0            # ORDER_FALLBACK | RANK_ABS | TRACE_DSL
AS_SHIPPED = TRACE_FALLBACK   # what the unmodified compiled reference computes

_vp, _i, _f, _ip, _lp = C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_longlong)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "pikg_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_epep.argtypes = [_vp, _i, _vp, _i, _vp, _f, _i]
        lib.oracle_epsp_quad.argtypes = [_vp, _i, _vp, _i, _vp, _f, _i]
        lib.oracle_epsp_mono.argtypes = [_vp, _i, _vp, _i, _vp, _f, _i]
        lib.oracle_force_clear.argtypes = [_vp, _i]
        lib.oracle_calc_walks.restype = C.c_longlong
        lib.oracle_calc_walks.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _f, _i, _i, _i, _i]
        _oracle = lib
    return _oracle


def have_ref(kind="scalar"):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libgplum_ref_%s.so" % kind))


_ref = {}


def ref(kind="scalar"):
    if kind not in _ref:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libgplum_ref_%s.so" % kind))
        lib.ref_epep.argtypes = [_vp, _i, _vp, _i, _vp, _f]
        lib.ref_epsp.argtypes = [_vp, _i, _vp, _i, _vp, _f]
        lib.ref_force_clear.argtypes = [_vp, _i]
        lib.ref_layout.argtypes = [_vp]
        lib.ref_calc_walks.restype = C.c_longlong
        lib.ref_calc_walks.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _f, _i, _i]
        lib.ref_tree_build.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, C.c_double, _i, _i, _i, _f]
        lib.ref_tree_sizes.argtypes = [_vp]
        lib.ref_tree_copy.argtypes = [_vp] * 12
        _ref[kind] = lib
    return _ref[kind]


# ------------------------------------------------------------------ single functor calls
def epep(epi, epj, eps2, flags=CANONICAL, force=None, lib="oracle"):
    f = S.cleared_force(len(epi)) if force is None else force.copy()
    epi = np.ascontiguousarray(epi); epj = np.ascontiguousarray(epj)
    if lib == "oracle":
        oracle().oracle_epep(_ptr(epi), len(epi), _ptr(epj), len(epj), _ptr(f), eps2, flags)
    else:
        ref(lib).ref_epep(_ptr(epi), len(epi), _ptr(epj), len(epj), _ptr(f), eps2)
    return f


def epsp(epi, spj, eps2, flags=CANONICAL, force=None, lib="oracle"):
    f = S.cleared_force(len(epi)) if force is None else force.copy()
    epi = np.ascontiguousarray(epi); spj = np.ascontiguousarray(spj)
    quad = spj.dtype.itemsize == 80
    if lib == "oracle":
        fn = oracle().oracle_epsp_quad if quad else oracle().oracle_epsp_mono
        fn(_ptr(epi), len(epi), _ptr(spj), len(spj), _ptr(f), eps2, flags)
    else:
        assert quad
        ref(lib).ref_epsp(_ptr(epi), len(epi), _ptr(spj), len(spj), _ptr(f), eps2)
    return f


# ------------------------------------------------------------------ batched walks
from gplum_b200.walks import Walks  # noqa: E402,F401


def calc_walks(w, eps2, flags=CANONICAL, lib="oracle", n_threads=0, clear=True, force=None):
    f = S.cleared_force(len(w.epi)) if force is None else force.copy()
    args = [w.n_walk, _ptr(w.epi), _ptr(w.epi_off), _ptr(w.ni), _ptr(w.adr_epj), _ptr(w.epj_disp),
            _ptr(w.n_epj), _ptr(w.adr_spj), _ptr(w.spj_disp), _ptr(w.n_spj), _ptr(w.epj_all),
            _ptr(w.spj_all), _ptr(f), eps2]
    if lib == "oracle":
        n = oracle().oracle_calc_walks(*args, int(w.quad), flags, int(clear), n_threads)
    else:
        assert w.quad
        n = ref(lib).ref_calc_walks(*args, int(clear), n_threads)
    return f, n


def calc_walks_abs(w, eps2=0.0, n_threads=0):
    """Per i-particle sums of the pair terms' MAGNITUDES (FP64): (sum_j |f_ij|, sum_j |phi_ij|) -- the
    conditioning of the force sums, used as the floor of the per-particle tolerance (synth.assert_force_close)."""
    lib = oracle()
    lib.oracle_calc_walks_abs.restype = None
    lib.oracle_calc_walks_abs.argtypes = [_i] + [_vp] * 13 + [C.c_double, _i, _i]
    n = len(w.epi)
    sa, sp = np.zeros(n), np.zeros(n)
    lib.oracle_calc_walks_abs(w.n_walk, _ptr(w.epi), _ptr(w.epi_off), _ptr(w.ni), _ptr(w.adr_epj), _ptr(w.epj_disp),
                              _ptr(w.n_epj), _ptr(w.adr_spj), _ptr(w.spj_disp), _ptr(w.n_spj), _ptr(w.epj_all),
                              _ptr(w.spj_all), _ptr(sa), _ptr(sp), float(eps2), int(w.quad), n_threads)
    return sa, sp


def ref_tree_walks(pos, mass, r_out, r_search, theta=0.5, n_leaf_limit=8, n_group_limit=64,
                   n_walk_limit=200, eps2=0.0, vel=None, kind="scalar", with_force=False):
    """Interaction lists produced by the reference's own FDPS tree (multi-walk-index interface)."""
    lib = ref(kind)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = len(pos)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    r_out = np.ascontiguousarray(r_out, dtype=np.float64)
    r_search = np.ascontiguousarray(r_search, dtype=np.float64)
    velp = None if vel is None else _ptr(np.ascontiguousarray(vel, dtype=np.float64))
    nw = lib.ref_tree_build(n, _ptr(pos), velp, _ptr(mass), _ptr(r_out), _ptr(r_search), theta,
                            n_leaf_limit, n_group_limit, n_walk_limit, eps2)
    sz = np.zeros(8, dtype=np.int64)
    lib.ref_tree_sizes(_ptr(sz))
    assert sz[0] == nw
    epi = np.zeros(sz[1], dtype=S.EPI)
    epi_off = np.zeros(nw, dtype=np.int32); ni = np.zeros(nw, dtype=np.int32)
    adr_epj = np.zeros(sz[2], dtype=np.int32); epj_disp = np.zeros(nw, dtype=np.int64)
    n_epj = np.zeros(nw, dtype=np.int32)
    adr_spj = np.zeros(sz[3], dtype=np.int32); spj_disp = np.zeros(nw, dtype=np.int64)
    n_spj = np.zeros(nw, dtype=np.int32)
    epj_all = np.zeros(sz[4], dtype=S.EPJ); spj_all = np.zeros(sz[5], dtype=S.SPJ_QUAD)
    force = np.zeros(sz[1], dtype=S.FORCE)
    lib.ref_tree_copy(_ptr(epi), _ptr(epi_off), _ptr(ni), _ptr(adr_epj), _ptr(epj_disp), _ptr(n_epj),
                      _ptr(adr_spj), _ptr(spj_disp), _ptr(n_spj), _ptr(epj_all), _ptr(spj_all), _ptr(force))
    w = Walks(epi, epi_off, ni, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, epj_all, spj_all)
    return (w, force) if with_force else w


# ------------------------------------------------------------------ changeover correction
def correct_long(w, prm, force=None):
    """oracle/soft_corr_oracle.c on the walks `w`: (corr[n_epi], init[n_epi] or None, ngb[n])."""
    lib = oracle()
    lib.oracle_correct_long.restype = C.c_longlong
    lib.oracle_correct_long.argtypes = [_i] + [_vp] * 12 + [C.c_longlong]
    n = len(w.epi)
    out = np.zeros(n, dtype=S.CORR)
    init = np.zeros(n, dtype=S.CORR_INIT) if int(prm["initial"][0]) else None
    cap = max(1024, 64 * n)
    ngb = np.zeros(cap, dtype=S.NGB)
    r = lib.oracle_correct_long(w.n_walk, _ptr(w.epi), _ptr(w.epi_off), _ptr(w.ni), _ptr(w.adr_epj),
                                _ptr(w.epj_disp), _ptr(w.n_epj), _ptr(w.epj_all),
                                None if force is None else _ptr(force), _ptr(prm), _ptr(out),
                                None if init is None else _ptr(init), _ptr(ngb), cap)
    assert r >= 0, "oracle_correct_long failed: %d" % r
    return out, init, ngb[:r]


def ref_correct_long(pos, vel, acc_d, mass, r_out, r_search, ids, prm, theta=0.5, n_leaf_limit=8,
                     n_group_limit=64, n_walk_limit=200, kind="scalar"):
    """The reference's own tree force + correctForceLong{,Initial} on one rank (oracle/ref_shim.cpp).
    Returns (walks recorded from that tree, tree force in walk order, dict of per-particle results
    in ORIGINAL order, list of neighbour arrays)."""
    lib = ref(kind)
    lib.ref_correct_long.restype = C.c_longlong
    lib.ref_correct_long.argtypes = [_i] + [_vp] * 7 + [C.c_double, _i, _i, _i] + [C.c_double] * 5 + [_i, _vp, _vp, _vp, C.c_longlong]
    n = len(pos)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pos, vel, acc_d, mass, r_out, r_search = map(f64, (pos, vel, acc_d, mass, r_out, r_search))
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    of = np.zeros((n, 16)); oi = np.zeros((n, 4), dtype=np.int64)
    cap = max(1024, 64 * n)
    ngb = np.zeros((cap, 3), dtype=np.int64)
    p = prm[0]
    r = lib.ref_correct_long(n, _ptr(pos), _ptr(vel), _ptr(acc_d), _ptr(mass), _ptr(r_out), _ptr(r_search),
                             _ptr(ids), theta, n_leaf_limit, n_group_limit, n_walk_limit, float(p["eps2"]),
                             float(p["dt_tree"]), float(p["gamma"]), float(p["R_search2"]), float(p["R_search3"]),
                             int(p["initial"]), _ptr(of), _ptr(oi), _ptr(ngb), cap)
    assert r >= 0, "ref_correct_long failed: %d" % r
    sz = np.zeros(8, dtype=np.int64)
    lib.ref_tree_sizes(_ptr(sz))
    nw = int(sz[0])
    epi = np.zeros(sz[1], dtype=S.EPI)
    epi_off = np.zeros(nw, dtype=np.int32); ni = np.zeros(nw, dtype=np.int32)
    adr_epj = np.zeros(sz[2], dtype=np.int32); epj_disp = np.zeros(nw, dtype=np.int64)
    n_epj = np.zeros(nw, dtype=np.int32)
    adr_spj = np.zeros(sz[3], dtype=np.int32); spj_disp = np.zeros(nw, dtype=np.int64)
    n_spj = np.zeros(nw, dtype=np.int32)
    epj_all = np.zeros(sz[4], dtype=S.EPJ); spj_all = np.zeros(sz[5], dtype=S.SPJ_QUAD)
    force = np.zeros(sz[1], dtype=S.FORCE)
    lib.ref_tree_copy(_ptr(epi), _ptr(epi_off), _ptr(ni), _ptr(adr_epj), _ptr(epj_disp), _ptr(n_epj),
                      _ptr(adr_spj), _ptr(spj_disp), _ptr(n_spj), _ptr(epj_all), _ptr(spj_all), _ptr(force))
    w = Walks(epi, epi_off, ni, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, epj_all, spj_all)
    res = {"acc": of[:, 0:3], "phi": of[:, 3], "acc0": of[:, 4], "acc_d": of[:, 5:8], "phi_d": of[:, 8],
           "jerk_d": of[:, 9:12], "acc_before": of[:, 12:15], "id_cluster": oi[:, 0], "number": oi[:, 1],
           "in_domain": oi[:, 2]}
    lists = [ngb[oi[i, 3]:oi[i, 3] + oi[i, 1]] for i in range(n)]
    return w, force, res, lists


def ref_stage_time(pos, vel, mass, r_out, r_search, prm, theta=0.5, n_leaf_limit=8, n_group_limit=64, reps=1, kind="simd"):
    """Seconds of the reference's own soft-force stage on this host (oracle/ref_shim.cpp: ref_stage_time):
    (calcForceAllAndWriteBack, correctForceLong, neighbours found).  The OpenMP build of FDPS keeps per-level
    arrays on the stack: the caller must have raised RLIMIT_STACK (bench.py does) for N >= 2e5."""
    lib = ref(kind)
    lib.ref_stage_time.restype = C.c_longlong
    lib.ref_stage_time.argtypes = [_i] + [_vp] * 5 + [C.c_double, _i, _i] + [C.c_double] * 5 + [_i, _vp]
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pos, vel, mass, r_out, r_search = map(f64, (pos, vel, mass, r_out, r_search))
    sec = np.zeros(2)
    p = prm[0]
    n_ngb = lib.ref_stage_time(len(pos), _ptr(pos), _ptr(vel), _ptr(mass), _ptr(r_out), _ptr(r_search), theta, n_leaf_limit,
                               n_group_limit, float(p["eps2"]), float(p["dt_tree"]), float(p["gamma"]), float(p["R_search2"]),
                               float(p["R_search3"]), int(reps), _ptr(sec))
    return float(sec[0]), float(sec[1]), int(n_ngb)


# ------------------------------------------------------------------ isolated-particle step
ISO_PARAMS = np.dtype([("m_sun", "<f8"), ("dt_tree", "<f8"), ("eta_0", "<f8"), ("eta_sun0", "<f8"),
                       ("alpha2", "<f8"), ("dt_min", "<f8"), ("eps2_sun", "<f8")])
ISO_STAR = np.dtype([("phi_s", "<f8"), ("acc_s", "<f8", (3,)), ("jerk_s", "<f8", (3,)), ("dt", "<f8")])


def iso_params(m_sun=1.0, dt_tree=2.0 ** -6, eta_0=0.002, eta_sun0=0.002, alpha=1.0, dt_min=2.0 ** -30, eps2_sun=0.0):
    """sample/parameter.dat's values (lines 52-60)."""
    p = np.zeros(1, dtype=ISO_PARAMS)
    p["m_sun"], p["dt_tree"], p["eta_0"], p["eta_sun0"] = m_sun, dt_tree, eta_0, eta_sun0
    p["alpha2"], p["dt_min"], p["eps2_sun"] = alpha * alpha, dt_min, eps2_sun
    return p


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt).copy()


def vel_kick(vel, acc, dt_tree, lib="oracle"):
    v, a = _c(vel), _c(acc)
    L = oracle() if lib == "oracle" else ref(lib)
    fn = L.oracle_vel_kick if lib == "oracle" else L.ref_vel_kick
    fn.argtypes = [_i, _vp, _vp, C.c_double]
    fn(len(v), _ptr(v), _ptr(a), float(dt_tree))
    return v


def kepler_isolated(pos, vel, time, dt, acc0, isolated, t0, t1, prm, lib="oracle"):
    """The loop of src/hard.h:793-817 on arrays: returns pos, vel, time, dt, star, handled."""
    n = len(pos)
    pos, vel, time, dt, acc0 = _c(pos), _c(vel), _c(time), _c(dt), _c(acc0)
    iso = _c(isolated, np.int32)
    star = np.zeros(n, dtype=ISO_STAR)
    handled = np.zeros(n, dtype=np.int32)
    if lib == "oracle":
        fn = oracle().oracle_kepler_isolated
        fn.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp]
        fn.restype = _i
        cnt = fn(n, _ptr(pos), _ptr(vel), _ptr(time), _ptr(dt), _ptr(acc0), _ptr(iso), float(t0), float(t1), _ptr(prm),
                 _ptr(star), _ptr(handled))
    else:
        fn = ref(lib).ref_kepler_isolated
        fn.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp]
        fn.restype = _i
        pv = np.array([prm[k][0] for k in ISO_PARAMS.names], dtype=np.float64)
        cnt = fn(n, _ptr(pos), _ptr(vel), _ptr(time), _ptr(dt), _ptr(acc0), _ptr(iso), float(t0), float(t1), _ptr(pv),
                 _ptr(star), _ptr(handled))
    assert cnt == int(handled.sum()) and cnt >= 0
    return pos, vel, time, dt, star, handled
