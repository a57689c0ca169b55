"""Host work-list builder of the force pass (csrc/gplum_b200.cu: build_items, csrc/items.h): every i-particle of
every walk is covered exactly once per list kind, tile shapes fit their i-counts, the list is sorted longest first,
and the EP/SP split is applied exactly to passes with less than one wave.  CPU only (no device call)."""
import ctypes as C

import numpy as np
import pytest

from gplum_b200._lib import check, lib

SHAPE = {0: 32, 1: 64, 9: 16, 10: 8, 11: 4}


def build(ni, ne, ns, warp_slots=3552, tile_cap=0, jsplit=1, epsp_split=-1):
    ni, ne, ns = (np.ascontiguousarray(a, dtype=np.int32) for a in (ni, ne, ns))
    cap = int(2 * ((ni + 3) // 4 + 1).sum() + 8)
    out = np.zeros((cap, 4), dtype=np.int32)
    n, hs = C.c_int(0), C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().gplum_b200_debug_build_items(len(ni), p(ni), p(ne), p(ns), warp_slots, tile_cap, jsplit, epsp_split,
                                             p(out), cap, C.byref(n), C.byref(hs)))
    return out[:n.value], bool(hs.value)


def check_cover(items, ni, ne, ns):
    ep = [np.zeros(k, np.int32) for k in ni]
    sp = [np.zeros(k, np.int32) for k in ni]
    for w, i0, n, cfg in items:
        k, part = cfg & 15, cfg >> 4
        assert k in SHAPE and 0 < n <= SHAPE[k] and i0 >= 0 and i0 + n <= ni[w]
        assert part in (0, 1, 2) and (part == 0 or k in (0, 1))
        if part in (0, 1):
            ep[w][i0:i0 + n] += 1
        if part in (0, 2):
            sp[w][i0:i0 + n] += 1
        if part:
            assert ne[w] > 0 and ns[w] > 0
    for w in range(len(ni)):
        assert (ep[w] == 1).all() and (sp[w] == 1).all(), w


@pytest.mark.parametrize("seed,n_walk,max_ni", [(0, 1, 64), (1, 40, 64), (2, 600, 512), (3, 5000, 512), (4, 300, 9)])
@pytest.mark.parametrize("jsplit,epsp", [(1, -1), (0, -1), (1, 0), (1, 1)])
def test_items_cover_every_i_particle_once(seed, n_walk, max_ni, jsplit, epsp):
    rng = np.random.default_rng(seed)
    ni = rng.integers(0, max_ni + 1, n_walk)
    ne = rng.integers(0, 900, n_walk) * (rng.random(n_walk) > 0.05)
    ns = rng.integers(0, 400, n_walk) * (rng.random(n_walk) > 0.05)
    items, has_split = build(ni, ne, ns, jsplit=jsplit, epsp_split=epsp)
    check_cover(items, ni, ne, ns)
    assert has_split == bool(((items[:, 3] >> 4) != 0).any())
    if epsp == 0:
        assert not has_split
    if epsp == 1 and ((ne > 0) & (ns > 0) & (ni > 0)).any() and (items[:, 3] & 15 <= 1).any():
        assert has_split
    if not jsplit:
        assert ((items[:, 3] & 15) <= 1).all()


def test_split_only_below_one_wave():
    ni = np.full(2000, 256); ne = np.full(2000, 700); ns = np.full(2000, 300)
    items, hs = build(ni, ne, ns)                       # 8000 tiles > 3552 warp slots
    assert not hs and len(items) == 8000 and (items[:, 3] == 1).all()
    items, hs = build(ni[:500], ne[:500], ns[:500])     # 2000 tiles < 3552: every tile twice
    assert hs and len(items) == 4000
    assert sorted(np.unique(items[:, 3]).tolist()) == [1 | 16, 1 | 32]
    # SP halves are the longer ones here (37 x 300 > 18.5 x 700 is false -> EP first): longest first either way
    cost = np.where(items[:, 3] & 16, 18.5 * 700, 37.0 * 300)
    assert (np.diff(cost) <= 0).all()


def test_small_pass_uses_smaller_tiles_and_tile_cap_is_honoured():
    ni = np.full(20, 64); ne = np.full(20, 500); ns = np.full(20, 200)
    items, _ = build(ni, ne, ns, epsp_split=0)          # 20 tiles of 64 would leave the GPU empty: j-split shapes
    assert ((items[:, 3] & 15) >= 9).all() and len(items) >= 888 // 4
    items, _ = build(np.full(3000, 200), np.full(3000, 500), np.full(3000, 200), tile_cap=32, epsp_split=0)
    assert (items[:, 2] <= 32).all()
