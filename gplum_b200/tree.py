"""Interaction lists from the host-side builder in libgplum_b200.so (csrc/let_tree.cpp): the
single-rank caller side of the force pass (FDPS semantics, own implementation)."""
import ctypes as C

import numpy as np

from . import structs as S
from ._lib import check, lib
from .walks import Walks


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
    check(lib().gplum_b200_tree_build(n, _p(pos), _p(mass), _p(r_out), _p(r_search), float(theta),
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
    check(lib().gplum_b200_tree_copy(_p(epi), _p(epi_off), _p(ni), _p(adr_e), _p(ed), _p(ne), _p(adr_s), _p(sd),
                                     _p(ns), _p(epj), _p(spj), int(quad), int(rank), _p(order)))
    lib().gplum_b200_tree_free()
    w = Walks(epi, epi_off, ni, adr_e, ed, ne, adr_s, sd, ns, epj, spj)
    assert w.n_interactions() == (int(sz[6]), int(sz[7]))
    return w, order
