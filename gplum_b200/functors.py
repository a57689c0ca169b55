"""Host-side mirror of the reference's interaction interface, over the C ABI.

  calcForceEPEPWithSearch / calcForceEPSP : src/gravity_kernel.hpp:8-23,125-136 -- same call
      shape `f(epi, ni, epj, nj, force)`, same accumulate semantics, FP_t::eps2 passed at
      construction like the PIKG-generated kernel's constructor (gravity_kernel.hpp:17-21).
  dispatch / retrieve : FDPS multi-walk-index accelerator functors
      (FDPS/src/tree_for_force_impl_force.hpp:78-83,232-242).
  calc_walks : one whole calcForce pass over flat arrays (impl_force.hpp:1404-1564).
  correctForceLong : the changeover correction + final neighbour lists on the walks of the last
      pass (src/gravity_soft.h:245-372; `initial=True` = correctForceLongInitial, :375-528).

Arrays are numpy structured arrays with the reference layouts (gplum_b200.structs).
"""
import ctypes as C

import numpy as np

from . import structs as S
from ._lib import check, lib

TRACE_AS_SHIPPED, RANK_SQUARED, NO_ACCUMULATE = 1, 2, 4


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def init(device=0, max_i=0, max_j=0):
    check(lib().gplum_b200_init(device, max_i, max_j))


def set_params(eps2=0.0, quad=True, flags=0):
    check(lib().gplum_b200_set_params(float(eps2), int(bool(quad)), int(flags)))


class calcForceEPEPWithSearch:
    def __init__(self, eps2=0.0):
        self.eps2 = float(eps2)

    def __call__(self, epi, ni, epj, nj, force):
        assert epi.dtype == S.EPI and epj.dtype == S.EPJ and force.dtype == S.FORCE
        assert epi.flags.c_contiguous and epj.flags.c_contiguous and force.flags.c_contiguous
        assert len(epi) >= ni and len(epj) >= nj and len(force) >= ni
        check(lib().gplum_b200_epep(_p(epi), ni, _p(epj), nj, _p(force), self.eps2))


class calcForceEPSP:
    def __init__(self, eps2=0.0):
        self.eps2 = float(eps2)

    def __call__(self, epi, ni, spj, nj, force):
        assert epi.dtype == S.EPI and spj.dtype in (S.SPJ_QUAD, S.SPJ_MONO) and force.dtype == S.FORCE
        assert epi.flags.c_contiguous and spj.flags.c_contiguous and force.flags.c_contiguous
        assert len(epi) >= ni and len(spj) >= nj and len(force) >= ni
        check(lib().gplum_b200_epsp(_p(epi), ni, _p(spj), nj, _p(force), self.eps2,
                                    int(spj.dtype == S.SPJ_QUAD)))


def _ptr_array(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def dispatch(tag, epi_list, adr_epj_list, adr_spj_list, epj_all, spj_all, send_all=False):
    """FDPS dispatch functor.  send_all=True ships epj_all/spj_all; otherwise the three lists
    hold one array per walk (views into FDPS's sorted arrays in the reference)."""
    L = lib()
    if send_all:
        check(L.gplum_b200_dispatch(tag, 0, None, None, None, None, None, None, _p(epj_all), len(epj_all),
                                    _p(spj_all), len(spj_all), 1))
        return
    nw = len(epi_list)
    ni = np.array([len(a) for a in epi_list], dtype=np.int32)
    ne = np.array([len(a) for a in adr_epj_list], dtype=np.int32)
    ns = np.array([len(a) for a in adr_spj_list], dtype=np.int32)
    check(L.gplum_b200_dispatch(tag, nw, _ptr_array(epi_list), _p(ni), _ptr_array(adr_epj_list), _p(ne),
                                _ptr_array(adr_spj_list), _p(ns), _p(epj_all), len(epj_all),
                                _p(spj_all), len(spj_all), 0))


def retrieve(tag, force_list):
    ni = np.array([len(a) for a in force_list], dtype=np.int32)
    check(lib().gplum_b200_retrieve(tag, len(force_list), _p(ni), _ptr_array(force_list)))


def _walk_args(w):
    return [w.n_walk, _p(w.epi), _p(w.epi_off), _p(w.ni), _p(w.adr_epj), _p(w.epj_disp), _p(w.n_epj),
            _p(w.adr_spj), _p(w.spj_disp), _p(w.n_spj), _p(w.epj_all), len(w.epj_all), _p(w.spj_all),
            len(w.spj_all)]


def calc_walks(w, force=None, clear=True):
    """One force pass through host buffers (H2D, kernels, D2H inside the call)."""
    f = S.cleared_force(len(w.epi)) if force is None else force
    check(lib().gplum_b200_calc_walks(*_walk_args(w), _p(f), int(clear)))
    return f


def walks_select(slot):
    check(lib().gplum_b200_walks_select(int(slot)))


def walks_upload(w, with_j=True):
    """with_j=False: upload only the walks (lists + i-particles); the j-set stays as it is."""
    a = _walk_args(w)
    if not with_j:
        a[10], a[11], a[12], a[13] = None, 0, None, 0
    check(lib().gplum_b200_walks_upload(*a))


def walks_run(repack=True):
    check(lib().gplum_b200_walks_run(int(repack)))


def walks_download(n_epi):
    f = np.zeros(n_epi, dtype=S.FORCE)
    check(lib().gplum_b200_walks_download(_p(f)))
    return f


def walks_time(iters, repack=True):
    ms = C.c_float(0)
    check(lib().gplum_b200_walks_time(iters, int(repack), C.byref(ms)))
    return ms.value


def fp32_peak(iters=20):
    t, ms = C.c_float(0), C.c_float(0)
    check(lib().gplum_b200_fp32_peak(iters, C.byref(t), C.byref(ms)))
    return t.value, ms.value


def counters(reset=False):
    a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
    lib().gplum_b200_counters(C.byref(a), C.byref(b), C.byref(c), int(reset))
    return a.value, b.value, c.value


# ------------------------------------------------------------------ changeover correction
def soft_corr_enable(on=True, pair_cap=0):
    """Make the following force passes record their neighbour-candidate pairs."""
    check(lib().gplum_b200_soft_corr_enable(int(bool(on)), int(pair_cap)))


def correct_long_run(prm, initial=False, slot=0):
    assert prm.dtype == S.CORR_PARAMS
    check(lib().gplum_b200_correct_long_run(int(slot), _p(prm), int(bool(initial))))


def correct_long_download(n_epi, initial=False, slot=0, ngb_cap=None):
    """(corr[n_epi], init[n_epi] or None, ngb) -- particle k's neighbours are
    ngb[corr[k].ngb_off : corr[k].ngb_off + corr[k].number]."""
    corr = np.zeros(n_epi, dtype=S.CORR)
    init = np.zeros(n_epi, dtype=S.CORR_INIT) if initial else None
    cap = int(ngb_cap) if ngb_cap is not None else 4 * n_epi + (1 << 20)
    ngb = np.zeros(cap, dtype=S.NGB)
    n_slots, n_pairs = C.c_longlong(0), C.c_longlong(0)
    check(lib().gplum_b200_correct_long_download(int(slot), _p(corr), None if init is None else _p(init), _p(ngb),
                                                 cap, C.byref(n_slots), C.byref(n_pairs)))
    return corr, init, ngb[:n_slots.value]


def correct_long_download_compact(n_epi, slot=0, corr_cap=None, ngb_cap=None):
    """(corr[m], ngb): only the particles with neighbours (walk order); the rest carry the self term alone."""
    cc = int(corr_cap) if corr_cap is not None else n_epi
    corr = np.zeros(cc, dtype=S.CORR)
    cap = int(ngb_cap) if ngb_cap is not None else 4 * n_epi + (1 << 20)
    ngb = np.zeros(cap, dtype=S.NGB)
    m, n_slots, n_pairs = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
    check(lib().gplum_b200_correct_long_download_compact(int(slot), _p(corr), cc, C.byref(m), _p(ngb), cap,
                                                         C.byref(n_slots), C.byref(n_pairs)))
    return corr[:m.value], ngb[:n_slots.value]


def correctForceLong(w, prm, initial=False):
    """Tree force + changeover correction of the walks `w` through host buffers: returns
    (force, corr, init, ngb).  The reference: calcForceAllAndWriteBack followed by
    correctForceLong{,Initial} (src/main_p3t.cpp:583-593, 351-361)."""
    soft_corr_enable(True)
    try:
        f = calc_walks(w)
        correct_long_run(prm, initial)
        corr, init, ngb = correct_long_download(len(w.epi), initial)
    finally:
        soft_corr_enable(False)
    return f, corr, init, ngb


def correct_long_time(prm, iters, initial=False, slot=0):
    ms = C.c_float(0)
    check(lib().gplum_b200_correct_long_time(int(slot), _p(prm), int(bool(initial)), int(iters), C.byref(ms)))
    return ms.value
