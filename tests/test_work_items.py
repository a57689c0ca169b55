"""Host work-list builder of the force pass (csrc/gplum_b200.cu: build_items, csrc/items.h): every i-particle of
every walk meets every j-tile of its walk exactly once, tile shapes fit their i-counts, the list is sorted longest
base tile first, the parts of a cut tile are consecutive and share scratch slots / an arrival counter, and a pass
with less than a quarter wave of items is laid out as one wave of equal-cost segments.  CPU only (no device call)."""
import ctypes as C

import numpy as np
import pytest

from gplum_b200._lib import check, lib

SHAPE = {0: 32, 1: 64, 9: 16, 10: 8, 11: 4}


def build(ni, ne, ns, warp_slots=3552, tile_cap=0, jsplit=1, split_m=2):
    ni, ne, ns = (np.ascontiguousarray(a, dtype=np.int32) for a in (ni, ne, ns))
    cap = int(((ni + 3) // 4 + 1).sum() + warp_slots + 16)
    out = np.zeros((cap, 8), dtype=np.int32)
    seg = np.zeros(warp_slots + 1, dtype=np.int32)
    n, n_slots, n_groups, n_seg = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().gplum_b200_debug_build_items(len(ni), p(ni), p(ne), p(ns), warp_slots, tile_cap, jsplit, split_m,
                                             p(out), cap, C.byref(n), C.byref(n_slots), C.byref(n_groups),
                                             p(seg), len(seg), C.byref(n_seg)))
    return out[:n.value], n_slots.value, n_groups.value, (seg[:n_seg.value + 1] if n_seg.value else None)


def n_tiles(ne, ns):
    return (ne + 63) // 64 + (ns + 63) // 64


def check_cover(items, ni, ne, ns, n_slots, n_groups):
    cover = [np.zeros((k, n_tiles(e, s)), np.int32) for k, e, s in zip(ni, ne, ns)]
    empty = [np.zeros(k, np.int32) for k in ni]             # walks without any j-tile still write their (cleared) force
    slots_seen, groups_seen = set(), set()
    k = 0
    while k < len(items):
        w, i0, n, cfg, t0, t1, slot0, group = items[k]
        shape, K = cfg & 15, (cfg >> 8) & 0xff
        assert shape in SHAPE and 0 < n <= SHAPE[shape] and i0 >= 0 and i0 + n <= ni[w]
        nt = n_tiles(ne[w], ns[w])
        if K <= 1:
            assert (t0, t1) == (0, -1)
            cover[w][i0:i0 + n, :] += 1
            empty[w][i0:i0 + n] += 1
            k += 1
            continue
        assert 2 <= K <= 250 and K <= nt
        assert group not in groups_seen and 0 <= group < n_groups
        groups_seen.add(group)
        prev = 0
        for q in range(K):                                   # the K parts are consecutive, in part order
            w2, i02, n2, cfg2, a, b, s2, g2 = items[k + q]
            assert (w2, i02, n2, s2, g2) == (w, i0, n, slot0, group)
            assert cfg2 == (shape | (K << 8) | (q << 16))
            assert a == prev and a < b <= nt
            prev = b
            cover[w][i0:i0 + n, a:b] += 1
            assert slot0 + q not in slots_seen and slot0 + q < n_slots
            slots_seen.add(slot0 + q)
        assert prev == nt
        empty[w][i0:i0 + n] += 1
        k += K
    for w in range(len(ni)):
        assert (cover[w] == 1).all() and (empty[w] == 1).all(), w
    assert len(slots_seen) == n_slots and len(groups_seen) == n_groups


def item_costs(items, ne, ns):
    """issue slots per lane of every work item (items.h cost model: 15.5 per EP pair, 34 per SP pair, 90 per j-tile)"""
    out = np.zeros(len(items))
    for k, (w, i0, n, cfg, t0, t1, _, _) in enumerate(items):
        shape = SHAPE[cfg & 15]
        ep_t = (ne[w] + 63) // 64
        nt = n_tiles(ne[w], ns[w])
        a, b = (0, nt) if t1 < 0 else (t0, t1)
        je = min(min(b, ep_t) * 64, ne[w]) - min(min(a, ep_t) * 64, ne[w])
        js = min(max(b - ep_t, 0) * 64, ns[w]) - min(max(a - ep_t, 0) * 64, ns[w])
        out[k] = (15.5 * je + 34.0 * js) * shape / 32.0 + 90.0 * (b - a) + (4000.0 if a == 0 else 0.0)
    return out


@pytest.mark.parametrize("seed,n_walk,max_ni", [(0, 1, 64), (1, 40, 64), (2, 600, 512), (3, 5000, 512), (4, 300, 9)])
@pytest.mark.parametrize("jsplit,split_m", [(1, 2), (0, 2), (1, 0)])
def test_items_cover_every_pair_once(seed, n_walk, max_ni, jsplit, split_m):
    rng = np.random.default_rng(seed)
    ni = rng.integers(0, max_ni + 1, n_walk)
    ne = rng.integers(0, 900, n_walk) * (rng.random(n_walk) > 0.05)
    ns = rng.integers(0, 400, n_walk) * (rng.random(n_walk) > 0.05)
    items, n_slots, n_groups, seg = build(ni, ne, ns, jsplit=jsplit, split_m=split_m)
    check_cover(items, ni, ne, ns, n_slots, n_groups)
    if split_m == 0:
        assert n_slots == 0 and ((items[:, 3] >> 8) == 0).all() and seg is None
    if not jsplit:
        assert ((items[:, 3] & 15) <= 1).all()
    if seg is not None:                                      # every item belongs to exactly one warp's segment
        assert len(seg) == 3552 + 1 and seg[0] == 0 and seg[-1] == len(items) and (np.diff(seg) >= 0).all()


def test_segments_only_for_small_passes_and_they_carry_equal_work():
    ni = np.full(2000, 256); ne = np.full(2000, 700); ns = np.full(2000, 300)
    items, n_slots, _, seg = build(ni, ne, ns)          # 8000 tiles: whole tiles, one per warp
    assert seg is None and n_slots == 0 and len(items) == 8000 and (items[:, 3] == 1).all()
    rng = np.random.default_rng(7)
    m = 550                                              # one rank's share of the N = 1e6 disk on 8 GPUs: ~2400 tiles,
    ni = rng.integers(100, 513, m); ne = rng.integers(300, 900, m); ns = rng.integers(200, 450, m)
    items, n_slots, n_groups, seg = build(ni, ne, ns)   # 0.66 waves: still whole tiles (measured faster), longest first
    assert seg is None and n_slots == 0
    check_cover(items, ni, ne, ns, n_slots, n_groups)
    c = item_costs(items, ne, ns)
    assert (np.diff(c) <= 1e-6 * c[:-1]).all()
    # such a pass is PLACED (items.h: place_item): entry k of bin b; the 592 schedulers' sums come out even
    assert len(c) <= 3552
    rounds = (len(c) + 591) // 592
    load = np.zeros(592); seen = np.zeros(len(c), int)
    for b in range(592):
        for k in range(rounds):
            idx = k * 592 + (591 - b if k & 1 else b)
            if idx < len(c):
                load[b] += c[idx]; seen[idx] += 1
    assert (seen == 1).all()
    plain = np.zeros(592)
    for idx in range(len(c)):
        plain[idx % 592] += c[idx]
    print(load.max() / load.mean(), plain.max() / plain.mean())
    assert load.max() / load.mean() < 1.08 < plain.max() / plain.mean()
    # below 0.25 waves (split_m = 2): one wave of equal segments
    m = 140
    items, n_slots, n_groups, seg = build(ni[:m], ne[:m], ns[:m])
    assert seg is not None and n_groups > 0
    check_cover(items, ni[:m], ne[:m], ns[:m], n_slots, n_groups)
    c = item_costs(items, ne, ns)
    per_seg = np.add.reduceat(np.concatenate([c, [0.0]]), np.minimum(seg[:-1], len(c)))
    per_seg[np.diff(seg) == 0] = 0.0
    mean = c.sum() / 3552
    # cuts fall on j-tile boundaries (one tile of 64 SP j against 64 i = 4.8e3 slots)
    assert per_seg.max() < 1.8 * mean and np.percentile(per_seg, 10) > 0.4 * mean
    items, _, _, seg = build(ni, ne, ns, split_m=16)    # the limit scales with split_m
    assert seg is not None


def test_small_pass_uses_smaller_tiles_and_tile_cap_is_honoured():
    ni = np.full(20, 64); ne = np.full(20, 500); ns = np.full(20, 200)
    items, _, _, _ = build(ni, ne, ns, split_m=0)       # 20 tiles of 64 would leave the GPU empty: lane-split shapes
    assert ((items[:, 3] & 15) >= 9).all() and len(items) >= 888 // 4
    items, _, _, seg = build(ni, ne, ns)                # with segments the tiles stay 64 wide and are cut along j instead
    assert seg is not None and ((items[:, 3] & 15) == 1).all() and len(items) > 150
    items, _, _, _ = build(np.full(3000, 200), np.full(3000, 500), np.full(3000, 200), tile_cap=32, split_m=0)
    assert (items[:, 2] <= 32).all()
