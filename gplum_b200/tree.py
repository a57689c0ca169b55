"""Interaction lists: from the host-side builder (libgplum_lists.so, csrc/let_tree.cpp -- workload tooling, FDPS's
tree list for list) and from the GPU builder inside libgplum_b200.so (csrc/dev_tree.cu)."""
import ctypes as C
import os

import numpy as np

from . import structs as S
from ._lib import check, lib
from .walks import Walks

LISTS_PATH = os.environ.get("GPLUM_B200_LISTS_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgplum_lists.so"))
_lists = None


def lists_lib():
    """libgplum_lists.so (include/gplum_b200_lists.h): host code only, loads without a GPU."""
    global _lists
    if _lists is None:
        if not os.path.exists(LISTS_PATH):
            raise RuntimeError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LISTS_PATH)
        h = C.CDLL(LISTS_PATH)
        vp, i = C.c_void_p, C.c_int
        h.gplum_b200_tree_build.restype = i
        h.gplum_b200_tree_build.argtypes = [i, vp, vp, vp, vp, C.c_double, i, i, vp]
        h.gplum_b200_tree_copy.restype = i
        h.gplum_b200_tree_copy.argtypes = [vp] * 11 + [i, i, vp]
        h.gplum_b200_tree_free.restype = None
        _lists = h
    return _lists


def _check_lists(rc):
    if rc != 0:
        raise RuntimeError("libgplum_lists error %d" % rc)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_walks(pos, mass, r_out, r_search, theta=0.5, n_leaf_limit=8, n_group_limit=64, quad=True, rank=0):
    """Returns (Walks, sorted_to_original): the i-groups and index lists of one force pass.
    Defaults are sample/parameter.dat's (theta=0.5, n_leaf_limit=8, n_group_limit=64)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = len(pos)
    mass = np.ascontiguousarray(np.broadcast_to(mass, (n,)), dtype=np.float64)
    r_out = np.ascontiguousarray(np.broadcast_to(r_out, (n,)), dtype=np.float64)
    r_search = np.ascontiguousarray(np.broadcast_to(r_search, (n,)), dtype=np.float64)
    sz = np.zeros(8, dtype=np.int64)
    L = lists_lib()
    _check_lists(L.gplum_b200_tree_build(n, _p(pos), _p(mass), _p(r_out), _p(r_search), float(theta),
                                         int(n_leaf_limit), int(n_group_limit), _p(sz)))
    nw = int(sz[0])
    epi = np.zeros(n, dtype=S.EPI)
    epj = np.zeros(n, dtype=S.EPJ)
    spj = np.zeros(int(sz[5]), dtype=S.SPJ_QUAD if quad else S.SPJ_MONO)
    epi_off = np.zeros(nw, np.int32); ni = np.zeros(nw, np.int32)
    adr_e = np.zeros(int(sz[2]), np.int32); adr_s = np.zeros(int(sz[3]), np.int32)
    ed = np.zeros(nw, np.int64); sd = np.zeros(nw, np.int64)
    ne = np.zeros(nw, np.int32); ns = np.zeros(nw, np.int32)
    order = np.zeros(n, np.int32)
    _check_lists(L.gplum_b200_tree_copy(_p(epi), _p(epi_off), _p(ni), _p(adr_e), _p(ed), _p(ne), _p(adr_s), _p(sd),
                                        _p(ns), _p(epj), _p(spj), int(quad), int(rank), _p(order)))
    L.gplum_b200_tree_free()
    w = Walks(epi, epi_off, ni, adr_e, ed, ne, adr_s, sd, ns, epj, spj)
    assert w.n_interactions() == (int(sz[6]), int(sz[7]))
    return w, order


# ------------------------------------------------------------------ on the GPU (csrc/dev_tree.cu)
def build_walks_gpu(pos, mass, r_out, r_search, theta=0.5, n_leaf_limit=8, n_group_limit=64, rank=0, vel=None):
    """Builds tree, i-groups and lists on the device from SoA host arrays; the result becomes the
    selected resident walk set + j-set (nothing is copied back).  Returns sizes[8] like the host
    builder: n_walk, n_epi, n_adr_epj, n_adr_spj, n_epj_all, n_spj_all, n_int_epep, n_int_epsp."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = len(pos)
    mass = np.ascontiguousarray(np.broadcast_to(mass, (n,)), dtype=np.float64)
    r_out = np.ascontiguousarray(np.broadcast_to(r_out, (n,)), dtype=np.float64)
    r_search = np.ascontiguousarray(np.broadcast_to(r_search, (n,)), dtype=np.float64)
    sz = np.zeros(8, dtype=np.int64)
    if vel is not None:                 # the records the changeover correction reads (gplum_b200_tree_build_gpu_vel)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        assert vel.shape == (n, 3)
        check(lib().gplum_b200_tree_build_gpu_vel(n, _p(pos), _p(vel), _p(mass), _p(r_out), _p(r_search), float(theta),
                                                  int(n_leaf_limit), int(n_group_limit), int(rank), _p(sz)))
        return sz
    check(lib().gplum_b200_tree_build_gpu(n, _p(pos), _p(mass), _p(r_out), _p(r_search), float(theta),
                                          int(n_leaf_limit), int(n_group_limit), int(rank), _p(sz)))
    return sz


def download_original(n):
    """ForceGrav[n] of the last GPU-built pass in the order the particles were handed in."""
    f = np.zeros(n, dtype=S.FORCE)
    check(lib().gplum_b200_tree_download_original(_p(f)))
    return f


def download_compact(n):
    """The same through gplum_b200_tree_download_compact, expanded on the host: {acc, phi} of every particle, the
    neighbour words of those that have candidates, ForceGrav::clear()'s values for the rest."""
    accphi = np.zeros((n, 4), dtype=np.float32)
    idx = np.zeros(n, dtype=np.int32); nb = np.zeros((n, 4), dtype=np.int32)
    cnt = C.c_int(0)
    check(lib().gplum_b200_tree_download_compact(_p(accphi), _p(idx), _p(nb), n, C.byref(cnt)))
    f = S.cleared_force(n)
    f["acc"] = accphi[:, :3]; f["phi"] = accphi[:, 3]
    k = idx[:cnt.value]
    f["number"][k] = nb[:cnt.value, 0]; f["rank"][k] = nb[:cnt.value, 1]
    f["id_max"][k] = nb[:cnt.value, 2]; f["id_min"][k] = nb[:cnt.value, 3]
    return f, cnt.value


def build_walks_gpu_epj(epj, theta=0.5, n_leaf_limit=8, n_group_limit=64):
    """Same from EPJGrav records in any order (FDPS's epj_org_), host memory."""
    epj = np.ascontiguousarray(epj, dtype=S.EPJ)
    sz = np.zeros(8, dtype=np.int64)
    check(lib().gplum_b200_tree_build_gpu_epj(len(epj), _p(epj), 0, float(theta), int(n_leaf_limit),
                                              int(n_group_limit), _p(sz)))
    return sz


def copy_walks_gpu(sz, quad=True):
    """(Walks, sorted_to_original) of the last GPU build, copied to the host (tests)."""
    nw, n = int(sz[0]), int(sz[1])
    epi = np.zeros(n, dtype=S.EPI)
    epj = np.zeros(int(sz[4]), dtype=S.EPJ)
    spj = np.zeros(int(sz[5]), dtype=S.SPJ_QUAD if quad else S.SPJ_MONO)
    epi_off = np.zeros(nw, np.int32); ni = np.zeros(nw, np.int32)
    adr_e = np.zeros(int(sz[2]), np.int32); adr_s = np.zeros(int(sz[3]), np.int32)
    ed = np.zeros(nw, np.int64); sd = np.zeros(nw, np.int64)
    ne = np.zeros(nw, np.int32); ns = np.zeros(nw, np.int32)
    order = np.zeros(n, np.int32)
    check(lib().gplum_b200_tree_copy_gpu(_p(epi), _p(epi_off), _p(ni), _p(adr_e), _p(ed), _p(ne), _p(adr_s),
                                         _p(sd), _p(ns), _p(epj), _p(spj), _p(order)))
    return Walks(epi, epi_off, ni, adr_e, ed, ne, adr_s, sd, ns, epj, spj), order


def gpu_build_times():
    """Device milliseconds of the last GPU build, by phase."""
    ms = (C.c_float * 6)()
    check(lib().gplum_b200_tree_gpu_times(ms))
    return dict(zip(("sort_gather", "cells_moments", "groups_sync", "count_walk", "fill_walk", "items_spj"), [float(x) for x in ms]))


def gpu_build_stamps():
    """Microseconds between the level boundaries inside the cooperative cells+moments kernel."""
    a = np.zeros(96, dtype=np.uint64)
    n = lib().gplum_b200_tree_gpu_stamps(_p(a), 96)
    return (np.diff(a[:n].astype(np.int64)) / 1e3).round(1).tolist()
