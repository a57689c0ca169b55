"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.
acc/phi: 1e-4 relative per particle (the reference kernel is FP32 -- BASELINE.json north_star);
neighbour info (number, id_max, id_min, rank==0): bit-exact."""
import os

import numpy as np
import pytest

import oracle_api as O
import synth
from gplum_b200 import functors as F, structs as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _init():
    F.init(0)
    F.set_params(0.0, True, 0)
    yield


def _load_walks(name):
    z = np.load(os.path.join(GOLD, name))
    w = O.Walks(*[z[k] for k in ("epi", "epi_off", "ni", "adr_epj", "epj_disp", "n_epj", "adr_spj",
                                 "spj_disp", "n_spj", "epj_all", "spj_all")])
    return w, z["force_ref"], float(z["eps2"])


@pytest.mark.parametrize("ni,nj,ns,seed,eps2,n_rank", [
    (1, 1, 1, 0, 0.0, 1), (24, 157, 166, 1, 0.0, 1), (64, 301, 200, 2, 0.0, 2), (31, 123, 60, 3, 1e-8, 3),
    (403, 739, 228, 4, 0.0, 1), (5, 0, 0, 5, 0.0, 1), (17, 40, 0, 6, 0.0, 1), (1, 513, 7, 7, 0.0, 1),
    (130, 33, 1, 8, 0.0, 4), (1000, 1479, 300, 9, 0.0, 1), (33, 257, 255, 10, 0.0, 1), (600, 64, 700, 11, 0.0, 2),
])
def test_functor_calls_vs_oracle(ni, nj, ns, seed, eps2, n_rank):
    epi, epj, spj = synth.make_group(ni, nj, ns, seed=seed, n_rank=n_rank, dup_self=nj >= ni)
    want1 = O.epep(epi, epj, eps2)
    got1 = S.cleared_force(ni)
    F.calcForceEPEPWithSearch(eps2)(epi, ni, epj, nj, got1)
    synth.assert_force_close(got1, want1, RTOL, "epep")
    want2 = O.epsp(epi, spj, eps2, force=want1)
    got2 = got1.copy()
    F.calcForceEPSP(eps2)(epi, ni, spj, ns, got2)
    synth.assert_force_close(got2, want2, RTOL, "epep+epsp")


@pytest.mark.parametrize("ni", [1, 2, 3, 4, 5, 8, 9, 15, 16, 17, 31, 32, 33, 36, 40, 47, 48, 49, 63, 64, 65, 68, 73, 80, 97, 129])
def test_every_tile_shape_incl_j_split_tails(ni):
    """i-tiles of 64 (two per lane), 32, and short tails whose j-lists are split over 2/4/8 lane
    groups (gplum_b200.cu build_items): forces and neighbour info must not depend on the shape.
    Dense field so that candidates fall into every lane group."""
    nj, ns = 333, 77
    epi, epj, spj = synth.make_group(ni, nj, ns, seed=100 + ni, box=0.02, r_out=2.0e-3, n_rank=2, dup_self=True)
    want = O.epsp(epi, spj, 0.0, force=O.epep(epi, epj, 0.0))
    assert want["number"].sum() > 0
    got = S.cleared_force(ni)
    F.calcForceEPEPWithSearch(0.0)(epi, ni, epj, nj, got)
    F.calcForceEPSP(0.0)(epi, ni, spj, ns, got)
    synth.assert_force_close(got, want, RTOL, "tile shapes ni=%d" % ni)
    assert (got["rank"] == want["rank"]).all()


def test_functor_accumulates_like_reference():
    epi, epj, spj = synth.make_group(20, 90, 30, seed=11)
    f0 = S.cleared_force(20)
    f0["acc"] = 1.5; f0["phi"] = -2.0; f0["number"] = 3; f0["id_max"] = 10 ** 6; f0["id_min"] = 2
    want = O.epep(epi, epj, 0.0, force=f0)
    got = f0.copy()
    F.calcForceEPEPWithSearch(0.0)(epi, 20, epj, 90, got)
    synth.assert_force_close(got, want, RTOL, "accumulate")
    assert (got["number"] == want["number"]).all() and (got["rank"] == want["rank"]).all()


def test_monopole_spj():
    epi, epj, spj = synth.make_group(50, 80, 120, seed=5)
    mono = np.zeros(len(spj), dtype=S.SPJ_MONO)
    mono["mass"] = spj["mass"]; mono["pos"] = spj["pos"]
    want = O.epsp(epi, mono, 0.0)
    got = S.cleared_force(50)
    F.calcForceEPSP(0.0)(epi, 50, mono, len(mono), got)
    synth.assert_force_close(got, want, RTOL, "mono")


def test_golden_groups_as_shipped_mode():
    """Against the compiled reference's outputs (fixtures), in its own arithmetic (tr = xx+yy+xx)."""
    z = np.load(os.path.join(GOLD, "groups.npz"))
    F.set_params(0.0, True, F.TRACE_AS_SHIPPED)
    try:
        for k in range(int(z["n_cases"])):
            g = lambda nm: z["c%d_%s" % (k, nm)]
            eps2 = float(g("eps2"))
            f = g("f0").copy()
            F.calcForceEPEPWithSearch(eps2)(g("epi"), len(g("epi")), g("epj"), len(g("epj")), f)
            synth.assert_force_close(f, g("f_epep"), RTOL, "golden epep %d" % k)
            assert (f["rank"] == g("f_epep")["rank"]).all()
            F.calcForceEPSP(eps2)(g("epi"), len(g("epi")), g("spj"), len(g("spj")), f)
            synth.assert_force_close(f, g("f_both"), RTOL, "golden both %d" % k)
    finally:
        F.set_params(0.0, True, 0)


@pytest.mark.parametrize("name", ["init3000_g64.npz", "disk2k_g256.npz"])
def test_golden_walks_flat_pass(name):
    w, f_ref, eps2 = _load_walks(name)
    F.set_params(eps2, True, F.TRACE_AS_SHIPPED)
    try:
        got = F.calc_walks(w)
        synth.assert_force_close(got, f_ref, RTOL, name + " as-shipped")
        assert (got["rank"] == f_ref["rank"]).all()
    finally:
        F.set_params(0.0, True, 0)
    F.set_params(eps2, True, 0)
    got = F.calc_walks(w)
    want, _ = O.calc_walks(w, eps2, flags=O.CANONICAL)
    synth.assert_force_close(got, want, RTOL, name + " canonical")


def test_dispatch_retrieve_matches_flat_pass():
    """FDPS multi-walk protocol: send-all, then batches of walks, retrieve accumulates."""
    w, f_ref, eps2 = _load_walks("init3000_g64.npz")
    F.set_params(eps2, True, 0)
    want, _ = O.calc_walks(w, eps2)
    force = S.cleared_force(len(w.epi))
    F.dispatch(0, None, None, None, w.epj_all, w.spj_all, send_all=True)
    batch = 40
    for b0 in range(0, w.n_walk, batch):
        ws = range(b0, min(b0 + batch, w.n_walk))
        epi_l = [w.epi[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws]
        ae_l = [w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]] for k in ws]
        as_l = [w.adr_spj[w.spj_disp[k]:w.spj_disp[k] + w.n_spj[k]] for k in ws]
        f_l = [force[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws]
        F.dispatch(0, epi_l, ae_l, as_l, w.epj_all, w.spj_all)
        F.retrieve(0, f_l)
    synth.assert_force_close(force, want, RTOL, "dispatch/retrieve")


def test_pipelined_dispatch_many_sub_batches_and_tags():
    """A dispatch large enough to be cut into several PCIe sub-batches (copy-in / compute / copy-out streams),
    two tags in flight, accumulate semantics.  Checked against the ORACLE on the same lists: what retrieve added to
    the caller's arrays is the oracle's force to the path's FP32 bar and its neighbour info exactly.  The start
    values have the force's own magnitude, so the relative bar bites on the sum."""
    from gplum_b200 import disk, tree
    n = 300000
    d = disk.make_disk(n, a_in=0.9, a_out=1.1, seed=4)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=256)
    F.set_params(0.0, True, 0)
    want, _ = O.calc_walks(w, 0.0, n_threads=0)
    cond = O.calc_walks_abs(w, 0.0)          # 3e5 particles: some force sums cancel to 1e-3 of their terms
    half = w.n_walk // 2
    force = S.cleared_force(n)
    force["acc"] = -0.5 * want["acc"]; force["phi"] = 0.25 * want["phi"]
    force["number"] = 1; force["rank"] = 0; force["id_max"] = 5; force["id_min"] = 2
    start = force.copy()
    F.dispatch(0, None, None, None, w.epj_all, w.spj_all, send_all=True)
    lists = []
    for tag, ws in ((0, range(0, half)), (1, range(half, w.n_walk))):
        epi_l = [w.epi[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws]
        ae_l = [w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]] for k in ws]
        as_l = [w.adr_spj[w.spj_disp[k]:w.spj_disp[k] + w.n_spj[k]] for k in ws]
        lists.append([force[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws])
        F.dispatch(tag, epi_l, ae_l, as_l, w.epj_all, w.spj_all)        # both tags queued before any retrieve
    F.retrieve(1, lists[1])
    F.retrieve(0, lists[0])
    # retrieve ACCUMULATES (PIKG/src/CUDA.rb:488-494): += on acc/phi/number/rank, max/min on the ids.
    # added = what arrived on top of the start values (FP64 difference of two FP32 numbers of the force's size:
    # the FP32 rounding of the sum is <= 1.2e-7 of it, three orders inside the bar)
    added = S.cleared_force(n)
    added["acc"] = (force["acc"].astype(np.float64) - start["acc"]).astype(np.float32)
    added["phi"] = (force["phi"].astype(np.float64) - start["phi"]).astype(np.float32)
    added["number"] = force["number"] - start["number"]
    added["rank"] = force["rank"] - start["rank"]
    added["id_max"] = want["id_max"]; added["id_min"] = want["id_min"]
    synth.assert_force_close(added, want, RTOL, "pipelined dispatch, accumulated part vs oracle", cond=cond)
    assert np.array_equal(force["id_max"], np.maximum(start["id_max"], want["id_max"]))
    assert np.array_equal(force["id_min"], np.minimum(start["id_min"], want["id_min"]))
    # and the overwrite mode FDPS's clear=true corresponds to
    F.set_params(0.0, True, F.NO_ACCUMULATE)
    try:
        F.dispatch(0, None, None, None, w.epj_all, w.spj_all, send_all=True)
        ws = range(w.n_walk)
        F.dispatch(2, [w.epi[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws],
                   [w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]] for k in ws],
                   [w.adr_spj[w.spj_disp[k]:w.spj_disp[k] + w.n_spj[k]] for k in ws], w.epj_all, w.spj_all)
        out = S.cleared_force(n)
        F.retrieve(2, [out[w.epi_off[k]:w.epi_off[k] + w.ni[k]] for k in ws])
    finally:
        F.set_params(0.0, True, 0)
    synth.assert_force_close(out, want, RTOL, "pipelined dispatch (overwrite) vs oracle", cond=cond)
    assert np.array_equal(out["rank"] == 0, want["rank"] == 0)
    # the flat pass of the same lists: same bar against the oracle
    synth.assert_force_close(F.calc_walks(w), want, RTOL, "flat pass vs oracle", cond=cond)


def test_force_parity_at_the_benchmark_configuration():
    """BASELINE configs[2] as bench.py runs it (N = 1e6, 0.9-1.1 AU, theta = 0.5, n_leaf_limit = 8, n_group_limit =
    512): the resident pass of ALL walks against the oracle on the same lists (about 2 s of CPU on 8 threads) --
    acc / phi to 1e-4 per particle, neighbour info exactly -- and the split work lists of a sub-wave shard
    (every 8th walk, as one rank of an 8-GPU run holds) to the same bar."""
    from gplum_b200 import disk, tree
    from gplum_b200.walks import Walks
    n = 1000000
    d = disk.make_disk(n)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, theta=0.5, n_leaf_limit=8, n_group_limit=512)
    want, n_int = O.calc_walks(w, 0.0, n_threads=0)
    cond = O.calc_walks_abs(w, 0.0)
    F.set_params(0.0, True, 0)
    F.walks_upload(w)
    F.walks_run(repack=True)
    got = F.walks_download(n)
    assert n_int == sum(w.n_interactions())
    synth.assert_force_close(got, want, RTOL, "N=1e6 g=512 resident pass", cond=cond)
    # how much of the disk needs the conditioning floor at all, and how far the plain bar is missed
    an = np.linalg.norm(want["acc"].astype(np.float64), axis=1)
    rel = np.linalg.norm(got["acc"].astype(np.float64) - want["acc"], axis=1) / an
    print("N=1e6 g=512: rel err of acc vs oracle: max %.3e, 99.99%% %.3e, particles above 1e-4: %d of %d"
          % (rel.max(), np.quantile(rel, 0.9999), int((rel > RTOL).sum()), n))
    assert (rel > RTOL).sum() <= 1e-4 * n and rel.max() < 1e-3
    assert want["number"].sum() > 50000          # the disk has real neighbour pairs: the exact path is exercised
    # one rank's share of an 8-way run: a pass with less than one wave of items (split work list)
    m = w.n_walk // 8
    sub = Walks(w.epi, w.epi_off[:m], w.ni[:m], w.adr_epj, w.epj_disp[:m], w.n_epj[:m], w.adr_spj, w.spj_disp[:m],
                w.n_spj[:m], w.epj_all, w.spj_all)
    F.walks_upload(sub)
    F.walks_run(repack=False)
    n_sub = int((sub.epi_off + sub.ni).max())
    got8 = F.walks_download(n_sub)
    synth.assert_force_close(got8, want[:n_sub], RTOL, "N=1e6 g=512, 1/8 shard (placed pass)", cond=(cond[0][:n_sub], cond[1][:n_sub]))
    # such a pass is placed (items.h: place_item): every warp claims its item by the SM and scheduler it runs on.
    # The trace shows that every item ran exactly once and that the (SM, scheduler) bins hold `rounds` items each
    # (one less where the last round is incomplete) -- unless the hardware spread the CTAs unevenly, which the
    # stealing path absorbs: then a few bins differ
    import ctypes as C
    from gplum_b200._lib import check, lib
    check(lib().gplum_b200_debug_trace(1, None, 0, None))
    F.walks_run(repack=False)
    tr = np.zeros((1 << 16, 4), dtype=np.uint64); cnt = C.c_int(0)
    check(lib().gplum_b200_debug_trace(0, tr.ctypes.data_as(C.c_void_p), len(tr), C.byref(cnt)))
    tr = tr[:cnt.value]
    assert 592 < len(tr) <= 3552 and (tr[:, 3] == 1).all()
    bins = (tr[:, 2] & np.uint64(0xffffffff)).astype(np.int64) * 4 + ((tr[:, 2] >> np.uint64(32)).astype(np.int64) & 3)
    per_bin = np.bincount(bins, minlength=592)
    rounds = (len(tr) + 591) // 592
    assert len(per_bin) == 592 and ((per_bin == rounds) | (per_bin == rounds - 1)).mean() > 0.9, np.bincount(per_bin)
    got8b = F.walks_download(n_sub)
    assert got8b.tobytes() == got8.tobytes()          # same layout, same bits


def test_device_resident_pass_and_counters():
    w, _, eps2 = _load_walks("disk2k_g256.npz")
    F.set_params(eps2, True, 0)
    F.counters(reset=True)
    F.walks_upload(w)
    F.walks_run(repack=True)
    got = F.walks_download(len(w.epi))
    want, n_int = O.calc_walks(w, eps2)
    synth.assert_force_close(got, want, RTOL, "resident")
    launches, n_ee, n_es = F.counters()
    assert (n_ee, n_es) == w.n_interactions() and n_ee + n_es == n_int
    assert launches >= 3          # 2 pack kernels (upload) + 2 (repack) + force
    ms = F.walks_time(3)
    assert ms > 0


def test_empty_and_ragged_walks():
    w, _, eps2 = _load_walks("init3000_g64.npz")
    # zero-length lists for some walks; a walk with ni == 0
    n_epj = w.n_epj.copy(); n_spj = w.n_spj.copy(); ni = w.ni.copy()
    n_epj[3] = 0; n_spj[5] = 0; n_epj[7] = 0; n_spj[7] = 0; ni[9] = 0
    w2 = O.Walks(w.epi, w.epi_off, ni, w.adr_epj, w.epj_disp, n_epj, w.adr_spj, w.spj_disp, n_spj,
                 w.epj_all, w.spj_all)
    F.set_params(eps2, True, 0)
    base = S.cleared_force(len(w.epi)); base["phi"] = 7.0
    got = F.calc_walks(w2, force=base.copy())
    want, _ = O.calc_walks(w2, eps2, force=base.copy())
    synth.assert_force_close(got, want, RTOL, "ragged")
    k = 9
    assert (got["phi"][w.epi_off[k]:w.epi_off[k] + w.ni[k]] == 7.0).all()     # untouched
    k = 7
    sl = slice(w.epi_off[k], w.epi_off[k] + w.ni[k])
    assert got[sl].tobytes() == S.cleared_force(w.ni[k]).tobytes()             # cleared, nothing added


def test_neighbour_flags_dense_field():
    """Many candidates per particle (r_search comparable to the spacing), several ranks."""
    epi, epj, spj = synth.make_group(200, 900, 0, seed=21, box=0.01, r_out=3.0e-3, n_rank=3)
    want = O.epep(epi, epj, 0.0)
    assert want["number"].mean() > 20
    got = S.cleared_force(200)
    F.calcForceEPEPWithSearch(0.0)(epi, 200, epj, 900, got)
    synth.assert_force_close(got, want, RTOL, "dense")
    assert (got["rank"] == want["rank"]).all()


def test_two_resident_walk_sets_share_one_j_set():
    """Interior/boundary split of the multi-GPU path on one GPU: two walk sets, one j upload."""
    w, _, eps2 = _load_walks("init3000_g64.npz")
    F.set_params(eps2, True, 0)
    want, _ = O.calc_walks(w, eps2)
    idx_a = np.arange(0, w.n_walk, 2); idx_b = np.arange(1, w.n_walk, 2)
    sub = lambda idx: O.Walks(w.epi, w.epi_off[idx], w.ni[idx], w.adr_epj, w.epj_disp[idx], w.n_epj[idx],
                              w.adr_spj, w.spj_disp[idx], w.n_spj[idx], w.epj_all, w.spj_all)
    wa, wb = sub(idx_a), sub(idx_b)
    try:
        F.walks_select(0); F.walks_upload(wa)
        F.walks_select(1); F.walks_upload(wb, with_j=False)
        F.walks_select(0); F.walks_run(repack=True); fa = F.walks_download(int((wa.epi_off + wa.ni).max()))
        F.walks_select(1); F.walks_run(repack=False); fb = F.walks_download(int((wb.epi_off + wb.ni).max()))
    finally:
        F.walks_select(0)
    got = S.cleared_force(len(w.epi))
    for ws, f in ((wa, fa), (wb, fb)):
        for k in range(ws.n_walk):
            sl = slice(ws.epi_off[k], ws.epi_off[k] + ws.ni[k])
            got[sl] = f[sl]
    synth.assert_force_close(got, want, RTOL, "two sets")
