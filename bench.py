#!/usr/bin/env python
"""bench.py -- soft-force interactions/s of the P3T force pass on B200 (BASELINE.json metric).

A "step" is one force pass (TreeForForce::calcForce: j-pack + every walk's EP-EP + EP-SP +
neighbour-candidate detection + force write) over the interaction lists of a synthetic
Kokubo-Ida disk.  Default workload: BASELINE configs[2], N=1e6 planetesimals in a 0.9-1.1 AU annulus,
sample/parameter.dat tree parameters, n_group_limit per --group (--n / --a-in / --a-out: configs[1], [4]).
The lists are FDPS's own, list for list (tests/test_tree.py pins the host builder to the compiled reference).

  value     interactions/s with raw particles + lists resident in HBM (CUDA events, max over ranks)
  e2e       N=1: the whole soft-force evaluation a caller of calcForceAllAndWriteBack sees, through the C ABI with
            HOST buffers: particle columns (48 B each) up, tree + groups + lists built on the GPU, force pass, {acc, phi}
            back in particle order (16 B each) + the neighbour words of the particles that have candidates --
            gplum_b200_tree_build_gpu + walks_run + tree_download_compact (what include/gravity_tree_b200.hpp calls).
            N>1: the same on N GPUs -- every rank ships its n/N particles (48 B records), NCCL all-gather over NVLink, the
            same tree on every GPU, every rank its share of the walks, ForceGrav of the share back (multi_tree_e2e_leg)
  e2e_multiwalk  the FDPS multi-walk-index functors (dispatch/retrieve) with host-built lists shipped every pass; at N>1
            every rank ships what its rank of an MPI-FDPS run holds (own walks, local + LET particles, its superparticles)
  parity_check   after the timed region: this run's forces of every 50th walk against the oracle
            (acc/phi 1e-4 with the conditioning floor of tests/synth.py, neighbour ints exact); non-zero exit on failure
  roofline  dominant kernel (force_pass_kernel): algorithmic flop (30/EP-EP pair, 59/EP-SP pair,
            SURVEY 8d) / its CUDA-event time, against the FP32 FFMA peak measured in the same run
  cpu_baseline  the reference's own functors (oracle/_ref, AVX2+OpenMP build) on a bounded sample;
  cpu_baseline_stage  the reference's whole stage (FDPS tree + walk + functors + correctForceLong) once

`--impl reference` times only the reference CPU path (all host threads) on the same config, with the lists the
reference's own FDPS tree records (oracle/_ref; nothing of the product library is loaded).
N>1 (torchrun): walks are sharded by Morton-contiguous domains; per step every rank packs its own j-particles and
signals its peers in one launch, and one force launch evaluates all its walks -- the walks that read other ranks'
particles gather them from the owners' HBM over NVLink inside the kernel, after an in-kernel wait for the peers'
flags (gplum_b200/multigpu.py; `--exchange halo|allgather` use NCCL collectives instead).
Fixed total work => strong scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the OpenMP build of FDPS (the CPU arms) keeps per-level arrays on the stack: 8 MB overflows at N >= 2e5
try:
    import resource
    _h = resource.getrlimit(resource.RLIMIT_STACK)[1]
    resource.setrlimit(resource.RLIMIT_STACK, (_h if _h != resource.RLIM_INFINITY and _h < (1 << 30) else (1 << 30), _h))
except Exception:
    pass
# torchrun defaults OMP_NUM_THREADS to 1 per rank; the plugin's host side (list flattening, force accumulation in
# dispatch/retrieve) is OpenMP code like the reference's: give every rank its share of the host cores
if os.environ.get("LOCAL_WORLD_SIZE") and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ["LOCAL_WORLD_SIZE"])))
# stdout carries ONE JSON line: keep NCCL's banner / debug output off it (must be set before NCCL initialises)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

FLOP_EPEP, FLOP_EPSP = 30.0, 59.0
def metric_name(n):
    return "soft-force interactions/s (EP-EP+EP-SP), N=%s disk" % ("1e%d" % round(np.log10(n)) if 10 ** round(np.log10(n)) == n else str(n))


def make_workload(n, group, seed=0, a_in=0.9, a_out=1.1):
    from gplum_b200 import disk, tree
    t0 = time.time()
    d = disk.make_disk(n, a_in=a_in, a_out=a_out, seed=seed)
    r_out, r_search = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, order = tree.build_walks(d["pos"], d["mass"], r_out, r_search, theta=0.5, n_leaf_limit=8,
                                n_group_limit=group)
    # fields the force kernels never read but the changeover correction does (EPJGrav::vel, id)
    w.epj_all["vel"] = d["vel"][order]
    w.epj_all["id"] = order
    t_build = time.time() - t0
    w.raw = {"pos": d["pos"], "mass": d["mass"], "r_out": r_out, "r_search": r_search}     # the particles themselves
    w.raw_vel = d["vel"]
    t0 = time.time()
    tree.build_walks(d["pos"], d["mass"], r_out, r_search, theta=0.5, n_leaf_limit=8, n_group_limit=group)
    w.t_host_lists = time.time() - t0            # host builder alone (all host threads), lists copied out
    return w, t_build


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        busy = [s for s in sm if s > 0.5 * (mx[0] if mx else 1)]
        return {"sm_mhz": float(np.median(busy)) if busy else (float(np.median(sm)) if sm else None),
                "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(w, eps2, budget_s, threads):
    """The reference's functors on a bounded sample of the walks (every k-th walk)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    kind = "reference" if O.have_ref("simd") else "port"
    lib = "simd" if kind == "reference" else "oracle"
    # probe rate on a small sample, then size the real sample to the budget
    def sample(stride):
        idx = np.arange(0, w.n_walk, stride)
        return O.Walks(w.epi, w.epi_off[idx], w.ni[idx], w.adr_epj, w.epj_disp[idx], w.n_epj[idx],
                       w.adr_spj, w.spj_disp[idx], w.n_spj[idx], w.epj_all, w.spj_all)
    tot = sum(w.n_interactions())
    probe = sample(max(1, w.n_walk // 64))
    t0 = time.time(); _, n_int = O.calc_walks(probe, eps2, lib=lib, n_threads=threads); dt = time.time() - t0
    rate = n_int / max(dt, 1e-6)
    stride = max(1, int(np.ceil(tot / (rate * budget_s))))
    s = sample(stride)
    # a whole pass may take well under the budget on a many-core host: repeat it
    reps = max(1, int(budget_s * rate / max(1, sum(s.n_interactions()))))
    return kind, s, O, lib, reps


def reference_lists(args):
    """The lists of the reference arm come from the reference itself: its FDPS tree records what it hands to the
    dispatch functor (oracle/ref_shim.cpp: ref_tree_build).  Same lists as make_workload's, bit for bit
    (tests/test_tree.py); falls back to the host builder where oracle/_ref does not exist."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    from gplum_b200 import disk
    if not O.have_ref("scalar"):
        return make_workload(args.n, args.group, a_in=args.a_in, a_out=args.a_out)[0]
    d = disk.make_disk(args.n, a_in=args.a_in, a_out=args.a_out, seed=0)
    r_out, r_search = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    return O.ref_tree_walks(d["pos"], d["mass"], r_out, r_search, theta=0.5, n_leaf_limit=8, n_group_limit=args.group,
                            n_walk_limit=512, vel=d["vel"])


def run_reference(args, w):
    threads = os.cpu_count()
    kind, s, O, lib, reps = cpu_reference_run(w, 0.0, args.cpu_seconds / max(1, args.steps + args.warmup), threads)
    for _ in range(args.warmup):
        O.calc_walks(s, 0.0, lib=lib, n_threads=threads)
    t0 = time.time()
    n_int = 0
    for _ in range(args.steps):
        for _ in range(reps):
            _, n = O.calc_walks(s, 0.0, lib=lib, n_threads=threads)
            n_int += n
    dt = time.time() - t0
    val = n_int / dt
    sample_desc = "%d of %d walks (every %d-th) of the same lists, %d pass(es) per step, %.3g interactions/step" % (
        s.n_walk, w.n_walk, max(1, w.n_walk // max(1, s.n_walk)), reps, n_int / args.steps)
    return {"metric": metric_name(args.n), "value": val, "unit": "interactions/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, w),
            "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": threads, "kind": kind, "sample": sample_desc},
            "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_config(args, w):
    ee, es = w.n_interactions()
    cfg = {(1000000, 0.9, 1.1): "configs[2]", (100000, 0.9, 1.1): "configs[1]", (10000000, 0.5, 10.0): "configs[4]"}.get(
        (args.n, args.a_in, args.a_out), "custom")
    return {"workload": "Kokubo-Ida disk N=%d a=[%g,%g]AU (BASELINE %s); theta=0.5 n_leaf_limit=8 "
                        "n_group_limit=%d; individual cutoff, quadrupole SPJ" % (args.n, args.a_in, args.a_out, cfg, args.group),
            "n_particles": args.n, "n_group_limit": args.group, "n_walks": int(w.n_walk),
            "interactions_epep": ee, "interactions_epsp": es,
            "l2": getattr(args, "l2_note", "n/a"),
            "parallelism": "i-groups sharded over %d GPU(s), EPJ exchange: %s" % (args.gpus, getattr(args, "exchange", "halo"))}


def parity_check(w, walk_range, force_local, e0, stride=50):
    """Forces this run produced (force_local[k] belongs to i-particle e0 + k of the global walk set w) against the
    oracle on every stride-th walk of [walk_range): acc/phi within 1e-4 (tests/synth.py: with the conditioning floor
    the large-N tests use), neighbour ints exact.  Returns a dict; "ok" False on any violation."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    import synth
    idx = np.arange(walk_range[0], walk_range[1], stride)
    if len(idx) == 0:
        return {"walks": 0, "particles": 0, "max_rel": 0.0, "ok": True}
    s = O.Walks(w.epi, w.epi_off[idx], w.ni[idx], w.adr_epj, w.epj_disp[idx], w.n_epj[idx], w.adr_spj, w.spj_disp[idx],
                w.n_spj[idx], w.epj_all, w.spj_all)
    want, _ = O.calc_walks(s, 0.0, n_threads=0)
    sa, sp = O.calc_walks_abs(s, 0.0)
    sel = np.concatenate([np.arange(w.epi_off[k], w.epi_off[k] + w.ni[k]) for k in idx])
    got = force_local[sel - e0]
    an = np.linalg.norm(want["acc"][sel].astype(np.float64), axis=1)
    rel = np.linalg.norm(got["acc"].astype(np.float64) - want["acc"][sel], axis=1) / np.maximum(an, 1e-300)
    ok, msg = True, ""
    try:
        synth.assert_force_close(got, want[sel], 1e-4, "bench parity", cond=(sa[sel], sp[sel]))
    except AssertionError as e:
        ok, msg = False, str(e)[:200]
    out = {"walks": int(len(idx)), "particles": int(len(sel)), "max_rel": float(rel.max()),
           "above_1e-4": int((rel > 1e-4).sum()), "neighbour_ints_exact": bool(ok or "acc" in msg or "phi" in msg), "ok": ok}
    if msg:
        out["error"] = msg
    return out


def stage_baseline(args, w):
    """The reference's whole stage on the host cores: tree_grav.calcForceAllAndWriteBack(...) + correctForceLong(...)
    exactly as src/main_p3t.cpp:583-593 calls them, all OpenMP threads, AVX2 build (oracle/ref_shim.cpp: ref_stage_time)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    from gplum_b200 import structs as S
    if not O.have_ref("simd"):
        return None
    raw = w.raw
    t_force, t_corr, n_ngb = O.ref_stage_time(raw["pos"], w.raw_vel, raw["mass"], raw["r_out"], raw["r_search"], S.corr_params(),
                                              theta=0.5, n_leaf_limit=8, n_group_limit=args.group, reps=2, kind="simd")
    dt = t_force + t_corr
    return {"ms_per_step": dt * 1e3, "calc_force_all_ms": t_force * 1e3, "correct_force_long_ms": t_corr * 1e3,
            "interactions_per_s": sum(w.n_interactions()) / dt, "neighbours": n_ngb, "cores": os.cpu_count(), "kind": "reference",
            "sample": "best of 2 evaluations of calcForceAllAndWriteBack + correctForceLong (src/main_p3t.cpp:583-593), "
                      "n_group_limit=%d" % args.group}


def captured_traffic(args, world):
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r2_force_pass_traffic.json")))
        if world == 1 and (args.n, args.group, args.a_in, args.a_out) == (j["n"], j["group"], j["a_in"], j["a_out"]):
            return int(j["dram_bytes"])
    except Exception:
        pass
    return None


def pinned_like(a):
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    b = t.numpy().view(a.dtype).reshape(a.shape)
    b[...] = a
    return b, t


def soft_step_leg(args, w, F, S, L, check):
    """Next rows of the path (SURVEY 8f-1 + 8f-2): one whole soft-force evaluation without host-side
    lists, driven the way include/gravity_tree_b200.hpp drives it.  Per step, inside the timed region: particle
    columns (pinned host memory, 48 B each) -> device, tree + i-groups + interaction lists built on the GPU
    (dev_tree.cu), force pass with candidate capture, {acc, phi} of every particle and the neighbour words of the
    listed ones back; velocities of the listed particles up, changeover correction, the corrections of the
    particles that have neighbours + neighbour lists back to pinned host memory.
    The reference: setParticleLocalTree .. calcForceAllAndWriteBack + correctForceLong
    (src/main_p3t.cpp:583-593)."""
    import ctypes as C
    import torch
    from gplum_b200 import tree
    n = args.n
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    raw = {k: pin(np.ascontiguousarray(v, dtype=np.float64)) for k, v in w.raw.items()}
    p_acc = pin(np.zeros((n, 4), dtype=np.float32)); p_idx = pin(np.zeros(n, dtype=np.int32)); p_nb = pin(np.zeros((n, 4), dtype=np.int32))
    vel_all = np.ascontiguousarray(w.raw_vel, dtype=np.float64)
    n_listed = C.c_int(0)
    p_corr = pin(np.zeros(n, dtype=S.CORR))
    ngb_cap = 4 * n + (1 << 20)
    p_ngb = pin(np.zeros(ngb_cap, dtype=S.NGB))
    prm = S.corr_params()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_slots, n_pairs, n_corr = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
    sizes = [None]

    def one():
        sizes[0] = tree.build_walks_gpu(raw["pos"], raw["mass"], raw["r_out"], raw["r_search"], theta=0.5,
                                        n_leaf_limit=8, n_group_limit=args.group)
        F.walks_run(repack=False)
        check(L.gplum_b200_tree_download_compact(vp(p_acc), vp(p_idx), vp(p_nb), n, C.byref(n_listed)))
        check(L.gplum_b200_tree_set_motion_gather(n_listed.value, vp(p_idx), vp(vel_all), None, None))   # velocities of the listed particles
        F.correct_long_run(prm)
        check(L.gplum_b200_correct_long_download_compact(0, vp(p_corr), n, C.byref(n_corr), vp(p_ngb), ngb_cap,
                                                         C.byref(n_slots), C.byref(n_pairs)))

    F.soft_corr_enable(True)
    try:
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        reps = max(3, args.steps // 2)
        t0 = time.perf_counter()
        for _ in range(reps):
            one()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        phases = tree.gpu_build_times()
        # the list build alone (device time of its kernels; two host syncs inside)
        t0 = time.perf_counter()
        for _ in range(reps):
            tree.build_walks_gpu(raw["pos"], raw["mass"], raw["r_out"], raw["r_search"], theta=0.5, n_leaf_limit=8,
                                 n_group_limit=args.group)
        torch.cuda.synchronize()
        dt_build = (time.perf_counter() - t0) / reps
        k_ms = F.walks_time(max(3, args.steps), repack=False)
    finally:
        F.soft_corr_enable(False)
    sz = sizes[0]
    n_int = int(sz[6] + sz[7])
    assert (int(sz[6]), int(sz[7])) == w.n_interactions(), "GPU-built lists differ from the host builder's"
    return {"ms_per_step": dt * 1e3, "interactions_per_s": n_int / dt,
            "h2d_bytes_per_step": int(48 * n + 28 * n_listed.value),
            "d2h_bytes_per_step": int(16 * n + 20 * n_listed.value + 64 * n_corr.value + 16 * n_slots.value),
            "particles_with_neighbours": int(n_corr.value), "particles_listed": int(n_listed.value),
            "list_build_ms_wall": dt_build * 1e3, "list_build_gpu_phases_ms": {k: round(v, 4) for k, v in phases.items()},
            "list_build_ms_host_builder": w.t_host_lists * 1e3,
            "force_pass_ms_on_gpu_lists": k_ms, "n_walks": int(sz[0]), "n_cells": int(sz[5]),
            "neighbour_pairs": int(n_pairs.value),
            "api": "gplum_b200_tree_build_gpu + walks_run + tree_download_compact + tree_set_motion_gather + correct_long_run + "
                   "correct_long_download_compact, pinned host buffers (include/gravity_tree_b200.hpp: "
                   "calcForceAllAndWriteBack + correctForceLong)"}


def tree_e2e_leg(args, w, F, S, L, check, reps):
    """e2e at N = 1: one soft-force evaluation as the caller of calcForceAllAndWriteBack sees it
    (FDPS/src/tree_for_force.hpp:1239-1253), through the C ABI with host buffers.  Inside the timed region, every
    step: raw particles (pinned host SoA, 48 B each) -> device, Morton sort + tree + moments + i-groups + lists on
    the GPU, force pass, results scattered back to particle order and copied to pinned host memory: {acc, phi} of
    every particle (16 B) + the neighbour words of the particles that have candidates (the others hold
    ForceGrav::clear()'s values).  Also timed with PAGEABLE caller buffers (FDPS's arrays are pageable)."""
    import ctypes as C
    import torch
    n = args.n
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    sz = np.zeros(8, dtype=np.int64)

    n_nb = C.c_int(0)

    def run(raw, out):
        acc4, idx, nbw = out
        def one():
            check(L.gplum_b200_tree_build_gpu(n, vp(raw["pos"]), vp(raw["mass"]), vp(raw["r_out"]), vp(raw["r_search"]),
                                              0.5, 8, args.group, 0, vp(sz)))
            F.walks_run(repack=False)
            check(L.gplum_b200_tree_download_compact(vp(acc4), vp(idx), vp(nbw), n, C.byref(n_nb)))
        for _ in range(2):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            one()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    raw_pin = {k: pin(np.ascontiguousarray(v, dtype=np.float64)) for k, v in w.raw.items()}
    outs = (pin(np.zeros((n, 4), dtype=np.float32)), pin(np.zeros(n, dtype=np.int32)), pin(np.zeros((n, 4), dtype=np.int32)))
    dt = run(raw_pin, outs)
    # the forces that came back are the pass's forces, in particle order
    out_pin = S.cleared_force(n)
    out_pin["acc"] = outs[0][:, :3]; out_pin["phi"] = outs[0][:, 3]
    kk = outs[1][:n_nb.value]
    for c, name in enumerate(("number", "rank", "id_max", "id_min")):
        out_pin[name][kk] = outs[2][:n_nb.value, c]
    order = w.epi["id_local"]
    F.walks_upload(w); F.walks_run(repack=True)
    ref = F.walks_download(n)
    da = np.linalg.norm(out_pin["acc"][order].astype(np.float64) - ref["acc"], axis=1) / np.linalg.norm(ref["acc"].astype(np.float64), axis=1)
    same = bool(np.array_equal(out_pin["number"][order], ref["number"]) and np.array_equal(out_pin["id_max"][order], ref["id_max"])
                and np.quantile(da, 0.9999) < 1e-4 and da.max() < 2e-3)        # list order inside a walk differs: last bits
    raw_page = {k: np.ascontiguousarray(v, dtype=np.float64).copy() for k, v in w.raw.items()}
    dt_page = run(raw_page, (np.zeros((n, 4), dtype=np.float32), np.zeros(n, dtype=np.int32), np.zeros((n, 4), dtype=np.int32)))
    n_int = int(sz[6] + sz[7])
    assert (int(sz[6]), int(sz[7])) == w.n_interactions(), "GPU-built lists differ from FDPS's"
    return {"value": n_int / dt, "unit": "interactions/s", "h2d_bytes_per_step": int(48 * n), "d2h_bytes_per_step": int(16 * n + 20 * n_nb.value + 4),
            "particles_with_neighbour_candidates": int(n_nb.value),
            "ms_per_step": dt * 1e3, "ms_per_step_pageable_buffers": dt_page * 1e3, "forces_match_resident_pass": same,
            "api": "gplum_b200_tree_build_gpu + gplum_b200_walks_run + gplum_b200_tree_download_compact, pinned host buffers "
                   "(include/gravity_tree_b200.hpp binds these behind calcForceAllAndWriteBack)"}


def multi_soft_step_leg(args, w, F, S, L, check, world, rank, dist, torch):
    """The soft-force evaluation of an N-GPU run without host-side lists (SURVEY 8e with 8f-1/2): every rank ships
    ITS n / world particles (pinned host EPJGrav records) to its GPU, the ranks all-gather the records over NVLink
    (NCCL), every GPU builds the same tree and evaluates its Morton-contiguous share of the walks -- forces, changeover
    correction, neighbour lists --, and every rank copies back the forces of its share and the corrections of its
    particles that have neighbours.  Wall clock per step, max over ranks."""
    import ctypes as C
    from gplum_b200 import state as ST
    from gplum_b200.multigpu import MultiGpuSoftStep
    n = args.n
    if n % world:
        return None
    m = n // world
    epj = ST.make_epj(w.raw["pos"][rank * m:(rank + 1) * m], w.raw_vel[rank * m:(rank + 1) * m], w.raw["mass"][rank * m:(rank + 1) * m],
                      w.raw["r_out"][rank * m:(rank + 1) * m], w.raw["r_search"][rank * m:(rank + 1) * m],
                      ids=np.arange(rank * m, (rank + 1) * m))
    epj["id_local"] = np.arange(rank * m, (rank + 1) * m)
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    p_epj = pin(epj)
    p_force = pin(np.zeros(n, dtype=S.FORCE))
    ms = MultiGpuSoftStep(epj, n, world, rank, theta=0.5, n_leaf_limit=8, n_group_limit=args.group)
    prm = S.corr_params()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    p_corr = pin(np.zeros(n, dtype=S.CORR))
    ngb_cap = 4 * n + (1 << 20)
    p_ngb = pin(np.zeros(ngb_cap, dtype=S.NGB))
    n_slots, n_pairs, n_corr = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)

    def one():
        ms.upload_local(p_epj)
        ms.step(prm)
        ms.forces(out=p_force)
        check(L.gplum_b200_correct_long_download_compact(0, vp(p_corr), n, C.byref(n_corr), vp(p_ngb), ngb_cap,
                                                         C.byref(n_slots), C.byref(n_pairs)))

    F.walks_select(0)
    F.soft_corr_enable(True)
    try:
        for _ in range(3):
            one()
        torch.cuda.synchronize(); dist.barrier()
        reps = max(3, args.steps // 2)
        t0 = time.perf_counter()
        for _ in range(reps):
            one()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        phases = {}
        from gplum_b200 import tree
        phases = tree.gpu_build_times()
    finally:
        F.soft_corr_enable(False)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    w0, w1, e0, e1 = ms.share()
    mine = torch.tensor([float(ms.sizes[6] + ms.sizes[7]), float(e1 - e0), float(n_corr.value)], dtype=torch.float64, device="cuda")
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    return {"ms_per_step": dt.item() * 1e3, "interactions_per_s": mine[0].item() / dt.item(),
            "interactions": int(mine[0].item()), "particles_covered": int(mine[1].item()), "particles_with_neighbours": int(mine[2].item()),
            "h2d_bytes_per_rank": int(112 * m), "allgather_bytes_per_rank": int(112 * n),
            "rank0_share": {"walks": [w0, w1], "particles": [e0, e1]},
            "rank0_list_build_gpu_phases_ms": {k: round(v, 4) for k, v in phases.items()},
            "api": "all_gather_into_tensor (NCCL) + gplum_b200_tree_build_gpu_part + walks_run + correct_long_run + "
                   "walks_download_range + correct_long_download_compact"}


def multi_tree_e2e_leg(args, w, F, S, L, check, world, rank, dist, torch, reps):
    """e2e at N > 1: the evaluation a caller of calcForceAllAndWriteBack sees on an N-GPU run, no host-side lists.
    Every step, inside the timed region: every rank copies ITS n / world particles (pinned host memory, 48 B records
    {pos, mass, r_out, r_search}) to its GPU, the ranks all-gather the records over NVLink (NCCL), every GPU builds
    the same tree and evaluates its Morton-contiguous share of the walks, and copies the ForceGrav records of its
    share back to pinned host memory (32 B each).  Wall clock per step, max over ranks."""
    from gplum_b200.multigpu import MultiGpuSoftStep
    n = args.n
    if n % world:
        return None
    m = n // world
    sl = slice(rank * m, (rank + 1) * m)
    rec = np.empty((m, 6), dtype=np.float64)
    rec[:, :3] = w.raw["pos"][sl]; rec[:, 3] = w.raw["mass"][sl]; rec[:, 4] = w.raw["r_out"][sl]; rec[:, 5] = w.raw["r_search"][sl]
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    p_rec = pin(rec)
    p_force = pin(np.zeros(n, dtype=S.FORCE))
    ms = MultiGpuSoftStep(rec, n, world, rank, theta=0.5, n_leaf_limit=8, n_group_limit=args.group)

    def one():
        ms.upload_local(p_rec)
        ms.step(None)
        ms.forces(out=p_force)

    F.walks_select(0)
    for _ in range(3):
        one()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    w0, w1, e0, e1 = ms.share()
    # this rank's forces against the single-rank resident pass of the same lists (tree order is the same tree's)
    mine = torch.tensor([float(ms.sizes[6] + ms.sizes[7]), float(e1 - e0)], dtype=torch.float64, device="cuda")
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    # the forces that came back, against the oracle on every 50th walk of this rank's share (the GPU-built tree is the
    # host builder's tree: same walks, same order)
    par = parity_check(w, (w0, w1), p_force[:e1 - e0], e0) if w1 > w0 else {"ok": True, "walks": 0}
    ok = torch.tensor([1.0 if par["ok"] else 0.0, float(par["walks"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.SUM)
    errs = [None] * world
    dist.all_gather_object(errs, par.get("error"))
    return {"value": mine[0].item() / dt.item(), "unit": "interactions/s", "h2d_bytes_per_step": int(48 * m),
            "d2h_bytes_per_step": int(32 * (e1 - e0)), "ms_per_step": dt.item() * 1e3,
            "allgather_bytes_per_rank": int(48 * n), "interactions": int(mine[0].item()), "particles_covered": int(mine[1].item()),
            "api": "per rank: 48 B records up + all_gather_into_tensor (NCCL) + gplum_b200_tree_build_gpu_part_rec48 + "
                   "gplum_b200_walks_run + gplum_b200_walks_download_range, pinned host buffers",
            "forces_match_oracle": bool(ok[0].item() == world), "walks_checked": int(ok[1].item()),
            **({"parity_errors": [e for e in errs if e]} if any(errs) else {})}


def resident_step_leg(args, w, F, S, L, check):
    """SURVEY 8f-3 on top of f1 + f2: the particles stay in HBM across steps.  One step, all inside the timed
    region: velKick, Kepler drift of the isolated particles, the particles that need the host's hard part
    (neighbours, eccentric orbits) pulled to pinned host memory and pushed back (stand-in for the Hermite
    integration the reference does there), tree + lists on the GPU, force pass with capture, changeover
    correction, second velKick.  The reference: one iteration of src/main_p3t.cpp:423-708 without the hard part."""
    import ctypes as C
    import torch
    from gplum_b200 import state as ST
    n = args.n
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    epj = ST.make_epj(w.raw["pos"], w.raw_vel, w.raw["mass"], w.raw["r_out"], w.raw["r_search"])
    prm_c, prm_i = S.corr_params(), ST.iso_params()
    dt_tree = float(prm_i["dt_tree"][0])
    cap = n
    p_rec, p_idx = pin(np.zeros(cap, dtype=S.EPJ)), pin(np.zeros(cap, dtype=np.int32))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_host = C.c_int(0)
    sz = np.zeros(8, dtype=np.int64)

    def force():
        check(L.gplum_b200_state_tree_build(0.5, 8, args.group, vp(sz)))
        F.walks_run(repack=False)
        F.correct_long_run(prm_c)

    def one(k):
        ST.kick(dt_tree)
        ST.drift(prm_i, k * dt_tree, (k + 1) * dt_tree)
        check(L.gplum_b200_state_pull_unhandled(vp(p_rec), vp(p_idx), cap, C.byref(n_host)))
        check(L.gplum_b200_state_push(vp(p_rec), vp(p_idx), n_host.value))
        force()
        ST.kick(dt_tree)

    F.soft_corr_enable(True)
    try:
        ST.upload(epj, np.zeros(n), np.zeros(n))
        force()
        for k in range(3):
            one(k)
        torch.cuda.synchronize()
        reps = max(3, args.steps // 2)
        t0 = time.perf_counter()
        for k in range(3, 3 + reps):
            one(k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    finally:
        F.soft_corr_enable(False)
    n_int = int(sz[6] + sz[7])
    return {"ms_per_step": dt * 1e3, "interactions_per_s": n_int / dt, "interactions_last_step": n_int,
            "particles_to_host_per_step": int(n_host.value),
            "h2d_bytes_per_step": int(n_host.value) * 116, "d2h_bytes_per_step": int(n_host.value) * 116 + 4,
            "api": "state_kick + state_drift + state_pull_unhandled/_push + state_tree_build + walks_run + "
                   "correct_long_run + state_kick; particles resident in HBM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1000000, help="particles (under torchrun use --particles: its own parser claims --n)")
    ap.add_argument("--group", type=int, default=512)
    ap.add_argument("--a-in", type=float, default=0.9)
    ap.add_argument("--a-out", type=float, default=1.1)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-stage-baseline", action="store_true", help="skip the one evaluation of the reference's whole stage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--boundary-cap", type=int, default=32, help="N>1: i-particles per work item of the boundary walks")
    ap.add_argument("--exchange", default="peer", choices=["peer", "halo", "allgather"],
                    help="N>1: boundary walks read peer HBM over NVLink inside the kernel (CUDA IPC), or a trimmed "
                         "halo all-to-all (the reference's LET idea), or a full all-gather of the packed EPJ")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(run_reference(args, reference_lists(args))))
        return

    import torch
    import torch.distributed as dist
    from gplum_b200 import functors as F, structs as S
    from gplum_b200._lib import lib, check
    from gplum_b200.multigpu import MultiGpuPass
    import ctypes as C

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, t_build = make_workload(args.n, args.group, a_in=args.a_in, a_out=args.a_out)
    ee, es = w.n_interactions()
    F.init(local_rank)
    F.set_params(0.0, True, 0)
    L = lib()
    stream = torch.cuda.Stream()          # a real (non-default) stream: library kernels, NCCL and the
    torch.cuda.set_stream(stream)         # timing events all go through it
    check(L.gplum_b200_set_stream(C.c_void_p(stream.cuda_stream)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ------------------------------------------------------------------ device-resident passes
    if world == 1:
        F.walks_upload(w)
        my_int = ee + es
        def step():
            F.walks_run(repack=True)
    else:
        mg = MultiGpuPass(w, world, rank, stream, exchange=args.exchange, boundary_cap=args.boundary_cap)
        sh, lw = mg.sh, mg.lw
        wi, wb = sh.walks_interior, sh.walks_boundary
        my_int = sum(lw.n_interactions())
        step, exchange, exch_bytes = mg.step, mg.exchange, mg.exchange_bytes

    # bytes one step of this rank reads: when they fit the 126 MB L2 a step would find the previous step's data there,
    # so a buffer larger than L2 is written between the timed steps (each step timed by its own pair of events)
    mine = w if world == 1 else lw
    step_bytes = (48 + 32) * len(mine.epi) + 4 * (len(mine.adr_epj) + len(mine.adr_spj)) + 160 * len(mine.epj_all) + 144 * len(mine.spj_all)
    l2_flush = step_bytes < 2 * 126e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if l2_flush else None
    for _ in range(args.warmup):
        step()
    barrier()
    F.counters(reset=True)
    if l2_flush:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush_buf.zero_()
            a.record(stream)
            step()
            b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
    launches, c_ee, c_es = F.counters()
    args.l2_note = ("this rank's per-step inputs (%.0f MB) fit the 126 MB L2: a 256 MB buffer is written between timed "
                    "steps, each step timed by its own CUDA events" % (step_bytes / 1e6)) if l2_flush else (
                    "per-step inputs (%.0f MB of lists + particles) exceed the 126 MB L2; no flush" % (step_bytes / 1e6))
    # ---- parity of THIS run's forces (every 50th walk of this rank) against the oracle
    parity = parity_check(w, (0, w.n_walk) if world == 1 else sh.walk_range,
                          F.walks_download(len(w.epi)) if world == 1 else mg.forces(), 0 if world == 1 else sh.epi_range[0])
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    n_int = torch.tensor([float(my_int)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_int, op=dist.ReduceOp.SUM)
    ms_per_step = t_ms.item() / args.steps
    value = n_int.item() / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ dominant kernel alone
    barrier()
    if world == 1:
        k_ms = F.walks_time(max(3, args.steps), repack=False)      # CUDA events on the launching stream
    else:
        F.walks_select(0); k_int = F.walks_time(max(3, args.steps), repack=False)
        if args.exchange == "peer":
            k_bnd = 0.0                     # one work list, one launch: interior and boundary items together
        else:
            F.walks_select(1); k_bnd = F.walks_time(max(3, args.steps), repack=False)
        k_ms = k_int + k_bnd
        # phase timings (this rank), for the scaling analysis: the all-gather alone, the two kernels alone
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ea.record(stream)
        for _ in range(10):
            h = exchange()
            if h is not None:
                h.wait()
        eb.record(stream)
        torch.cuda.synchronize()
        my_step_ms = ms / args.steps
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        host_ms = (time.perf_counter() - t0) / 10 * 1e3          # host enqueue time per step (no sync inside)
        barrier()
        tr = {}
        mg.step(trace=tr)
        timeline = mg.trace_ms(tr)
        phases = {"rank": rank, "step_ms": my_step_ms, "host_enqueue_ms": host_ms, "timeline_ms": timeline, "exchange": args.exchange, "exchange_ms": ea.elapsed_time(eb) / 10, "interior_kernel_ms": k_int, "boundary_kernel_ms": k_bnd,
                  "interior_walks": int(wi.n_walk), "boundary_walks": int(wb.n_walk),
                  "exchange_bytes": int(exch_bytes)}
    soft_corr = None
    if world == 1:
        # next row of the path: the pass with candidate capture on + the changeover
        # correction kernels (correctForceLong, src/gravity_soft.h:245-372) on its pairs
        prm = S.corr_params()
        F.soft_corr_enable(True)
        F.walks_run(repack=False)
        k_cap_ms = F.walks_time(max(3, args.steps), repack=False)
        F.walks_run(repack=False)
        c_ms = F.correct_long_time(prm, max(3, args.steps))
        corr, _, ngb = F.correct_long_download(len(w.epi))
        F.soft_corr_enable(False)
        soft_corr = {"force_pass_with_capture_ms": k_cap_ms, "correction_ms": c_ms,
                     "candidate_pairs": int(len(ngb)), "neighbours": int(corr["number"].sum()),
                     "particles_with_neighbours": int((corr["number"] > 0).sum())}
    soft_step = None
    if world == 1:
        soft_step = soft_step_leg(args, w, F, S, L, check)
        soft_step["resident"] = resident_step_leg(args, w, F, S, L, check)
        F.walks_upload(w)                        # back to the host-built set for the legs below
    peak_tf, _ = F.fp32_peak(10)
    my_ee, my_es = (ee, es) if world == 1 else sh.local.n_interactions()
    flop = FLOP_EPEP * my_ee + FLOP_EPSP * my_es
    achieved = flop / (k_ms * 1e-3) / 1e12
    alg_bytes = (48 + 32) * (len(w.epi) if world == 1 else len(sh.local.epi)) + \
        4 * ((len(w.adr_epj) + len(w.adr_spj)) if world == 1 else (len(sh.local.adr_epj) + len(sh.local.adr_spj)))
    roofline = {"bound": "fp32", "kernel": "force_pass_kernel", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf,
                # DRAM bytes of one launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed
                # `ncu --set full` capture of this workload by the same kernel source (profiles/r2_force_pass_traffic.json,
                # written by tools/ncu_traffic.py); null for workloads without a capture
                "traffic": captured_traffic(args, world),
                "traffic_unit": "bytes/launch",
                "peak_nominal": 74.4,
                "peak_source": "FFMA microbenchmark in this run (MEASURED_PEAKS.json has no CUDA-core figure); "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "frac_of_nominal": achieved / 74.4,
                "kernel_ms": k_ms, "flop_per_launch": flop,
                # `achieved` counts the DSL's algorithm (30 / 59 flop per pair, its 5-flop Newton step
                # included).  The shipped kernel takes MUFU.RSQ's 2^-22.9 as the refined value and does not run that
                # step (kernels.cuh GB_NEWTON, profiles/r2_newton_step.txt): the flop it executes are 25 / 54 per pair
                "executed_flop_per_launch": (FLOP_EPEP - 5.0) * my_ee + (FLOP_EPSP - 5.0) * my_es,
                "executed_frac": ((FLOP_EPEP - 5.0) * my_ee + (FLOP_EPSP - 5.0) * my_es) / (k_ms * 1e-3) / 1e12 / peak_tf,
                "interactions_per_s_kernel_only": (my_ee + my_es) / (k_ms * 1e-3),
                "gflops_38flop_convention": 38.0 * (my_ee + my_es) / (k_ms * 1e-3) / 1e9,
                "list_bytes_per_launch": alg_bytes}

    # ------------------------------------------------------------------ end to end via the C ABI
    if world > 1:
        mg.close()                               # peer mode off: the legs below use plain indices
        soft_step = multi_soft_step_leg(args, w, F, S, L, check, world, rank, dist, torch)
        if soft_step is not None and soft_step["interactions"] != ee + es:
            soft_step["error"] = "the ranks' lists do not add up to the single-rank pass (%d vs %d)" % (soft_step["interactions"], ee + es)
    F.walks_select(0)
    if world == 1:
        lw = w
    else:
        # each rank ships what its rank of an MPI-FDPS run holds: its own walks, epj_sorted_ = local particles + LET
        # imports, spj_sorted_ = the superparticles its walks reference (gplum_b200/shard.py: rank_inputs)
        from gplum_b200.shard import HaloShard
        lw = HaloShard(w, world, rank).rank_inputs(w)
    keep = []
    def pin(a):
        b, t = pinned_like(a); keep.append(t); return b
    p_epi, p_epj, p_spj = pin(lw.epi), pin(lw.epj_all), pin(lw.spj_all)
    p_ae, p_as = pin(lw.adr_epj), pin(lw.adr_spj)
    p_force = pin(S.cleared_force(len(lw.epi)))
    epi_l = [p_epi[lw.epi_off[k]:lw.epi_off[k] + lw.ni[k]] for k in range(lw.n_walk)]
    ae_l = [p_ae[lw.epj_disp[k]:lw.epj_disp[k] + lw.n_epj[k]] for k in range(lw.n_walk)]
    as_l = [p_as[lw.spj_disp[k]:lw.spj_disp[k] + lw.n_spj[k]] for k in range(lw.n_walk)]
    f_l = [p_force[lw.epi_off[k]:lw.epi_off[k] + lw.ni[k]] for k in range(lw.n_walk)]
    nw = lw.n_walk
    c_epi = (C.c_void_p * nw)(*[a.ctypes.data for a in epi_l])
    c_ae = (C.c_void_p * nw)(*[a.ctypes.data for a in ae_l])
    c_as = (C.c_void_p * nw)(*[a.ctypes.data for a in as_l])
    c_f = (C.c_void_p * nw)(*[a.ctypes.data for a in f_l])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    F.set_params(0.0, True, F.NO_ACCUMULATE)     # FDPS clear=true: forces are overwritten

    def e2e_step():
        check(L.gplum_b200_dispatch(0, 0, None, None, None, None, None, None, vp(p_epj), len(p_epj), vp(p_spj), len(p_spj), 1))
        check(L.gplum_b200_dispatch(0, nw, c_epi, vp(lw.ni), c_ae, vp(lw.n_epj), c_as, vp(lw.n_spj),
                                    vp(p_epj), len(p_epj), vp(p_spj), len(p_spj), 0))
        check(L.gplum_b200_retrieve(0, nw, vp(lw.ni), c_f))

    for _ in range(2):
        e2e_step()
    barrier()
    n_e2e = max(3, args.steps // 4)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    h2d = p_epi.nbytes + p_epj.nbytes + p_spj.nbytes + p_ae.nbytes + p_as.nbytes + nw * 28
    d2h = p_force.nbytes
    e2e = {"value": n_int.item() / t_e2e.item(), "unit": "interactions/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e.item() * 1e3,
           "api": "gplum_b200_dispatch(send_all) + gplum_b200_dispatch(walks) + gplum_b200_retrieve, pinned host buffers"}
    F.set_params(0.0, True, 0)
    e2e_multiwalk = None
    if world > 1:
        # the headline e2e at N > 1 is the same evaluation as at N = 1 -- particles in host memory, no host-side lists --
        # on N GPUs; the multi-walk-index functors above ship host-built lists every pass: reported as e2e_multiwalk
        res = multi_tree_e2e_leg(args, w, F, S, L, check, world, rank, dist, torch, n_e2e)
        if res is not None:
            e2e_multiwalk = e2e
            e2e = res
    if world == 1:
        # the headline e2e at N = 1 is the whole evaluation a caller of calcForceAllAndWriteBack sees: raw particles in
        # host memory -> tree, groups and lists built on the GPU -> force pass -> forces back in particle order.
        # (the multi-walk-index functors above ship host-built lists every pass: reported as e2e_multiwalk)
        e2e_multiwalk = e2e
        e2e = tree_e2e_leg(args, w, F, S, L, check, n_e2e)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    all_phases = None
    if world > 1:
        all_phases = [None] * world
        dist.all_gather_object(all_phases, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in phases.items()
                                            if k in ("rank", "step_ms", "host_enqueue_ms", "timeline_ms", "exchange_ms", "interior_kernel_ms", "boundary_kernel_ms",
                                                     "interior_walks", "boundary_walks")})
    par_all = [parity]
    if world > 1:
        par_all = [None] * world
        dist.all_gather_object(par_all, parity)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sys.exit(0 if parity["ok"] else 3)
    parity_out = {"walks": sum(p["walks"] for p in par_all), "particles": sum(p["particles"] for p in par_all),
                  "max_rel": max(p["max_rel"] for p in par_all), "above_1e-4": sum(p["above_1e-4"] for p in par_all),
                  "ok": all(p["ok"] for p in par_all), "ranks_checked": len(par_all),
                  "what": "forces of the timed run, every 50th walk of every rank, against the oracle (acc/phi 1e-4 "
                          "with the conditioning floor of tests/synth.py; number/id_max/id_min/rank==0 exact)"}
    errs = [p["error"] for p in par_all if "error" in p]
    if errs:
        parity_out["errors"] = errs
    out = {"metric": metric_name(args.n), "value": value, "unit": "interactions/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, w),
           "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
           "parity_check": parity_out, "list_build_s_host": t_build}
    if e2e_multiwalk is not None:
        out["e2e_multiwalk"] = e2e_multiwalk
    if soft_corr is not None:
        out["soft_corr"] = soft_corr
    if soft_step is not None:
        out["soft_step"] = soft_step
    if world > 1:
        out["phases_rank0"] = phases
        out["phases_all"] = all_phases
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count()
        kind, s, O, libname, reps = cpu_reference_run(w, 0.0, args.cpu_seconds, threads)
        t0 = time.time()
        n = 0
        for _ in range(reps):
            n += O.calc_walks(s, 0.0, lib=libname, n_threads=threads)[1]
        dt = time.time() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": "interactions/s", "cores": threads, "kind": kind,
                               "sample": "%d of %d walks of the same lists x %d passes (%.3g interactions, %.1f s)" % (
                                   s.n_walk, w.n_walk, reps, n, dt)}
    if world == 1 and not args.no_cpu_baseline and not args.no_stage_baseline:
        st = stage_baseline(args, w)
        if st is not None:
            out["cpu_baseline_stage"] = st
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if not parity_out["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
