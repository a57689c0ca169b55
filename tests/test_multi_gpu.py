"""Multi-GPU force pass on real GPUs (needs >= 2 devices; skipped on a 1-GPU box -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`, or at 8 ranks with GPLUM_TEST_WORLD=8 under
`gpurun --gpus 8`; logs of both: profiles/r2_pytest_multi_w2.log, r2_pytest_multi_w8.log).  NCCL ranks shard the walks
of one disk, exchange EPJ (halo all-to-all and full all-gather variants) and evaluate their share
with the CUDA kernels; the union must match the single-rank oracle pass: acc/phi 1e-4, neighbour
info bit-exact."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
WORLD = int(os.environ.get("GPLUM_TEST_WORLD", "2"))          # gpurun --gpus 8: GPLUM_TEST_WORLD=8


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _workload(n=20000, group=128):
    from gplum_b200 import disk, tree
    d = disk.make_disk(n, a_in=0.97, a_out=1.03, seed=5)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=group)
    return w


def _worker(rank, world, port, out_dir, exchange, n=20000, group=128):
    import torch
    import torch.distributed as dist
    from gplum_b200 import functors as F
    from gplum_b200.multigpu import MultiGpuPass
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    w = _workload(n, group)
    if exchange == "peer" and n > 20000:
        os.environ["GPLUM_B200_FUSE_PACK"] = "1"       # read at init: the step's pack runs in the force launch's prologue
    F.init(rank)
    F.set_params(0.0, True, 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    import ctypes as C
    from gplum_b200._lib import check, lib
    check(lib().gplum_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    mg = MultiGpuPass(w, world, rank, stream, exchange=exchange)
    for _ in range(3):                      # repeated steps reuse the buffers: results must not drift
        F.counters(reset=True)
        mg.step()
        launches = F.counters()[0]
    np.save(os.path.join(out_dir, "l%d.npy" % rank), np.array([launches, mg.sh.walks_interior.n_walk + mg.sh.walks_boundary.n_walk]))
    f = mg.forces()
    np.save(os.path.join(out_dir, "f%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(list(mg.sh.epi_range) + [mg.n_boundary]))
    mg.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < WORLD, reason="needs >= %d GPUs" % WORLD)
@pytest.mark.parametrize("exchange", ["peer", "halo", "allgather"])
def test_two_ranks_match_single_rank_oracle(exchange, tmp_path):
    import torch.multiprocessing as mp
    import oracle_api as O
    import synth
    from gplum_b200 import structs as S
    world = WORLD
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), exchange), nprocs=world, join=True)
    w = _workload()
    want, _ = O.calc_walks(w, 0.0)
    got = S.cleared_force(len(w.epi)); covered = 0; n_bnd = 0
    for r in range(world):
        f = np.load(tmp_path / ("f%d.npy" % r)); e0, e1, nb = np.load(tmp_path / ("r%d.npy" % r))
        got[e0:e1] = f; covered += e1 - e0; n_bnd += nb
    assert covered == len(w.epi) and n_bnd > 0
    synth.assert_force_close(got, want, 1e-4, "%d ranks, " % world + exchange)


@pytest.mark.skipif(_ngpu() < WORLD, reason="needs >= %d GPUs" % WORLD)
def test_peer_step_of_a_placed_pass_is_one_launch(tmp_path):
    """A rank's share with at most one wave of work items (what 8 GPUs hold of the N = 1e6 disk) is a placed pass, and
    with GPLUM_B200_FUSE_PACK=1 the step's pack in peer mode -- EPJ slab, superparticles, flags to the peers -- runs in
    the prologue of that one cooperative launch (kernels.cuh: fp_*; off by default, it measured slower than the
    separate pack launch).  Forces against the single-rank oracle; one kernel launch per step."""
    import torch.multiprocessing as mp
    import oracle_api as O
    import synth
    from gplum_b200 import structs as S
    world = WORLD
    n, group = 80000 * world, 512               # ~1900 items per rank
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), "peer", n, group), nprocs=world, join=True)
    w = _workload(n, group)
    want, _ = O.calc_walks(w, 0.0)
    cond = O.calc_walks_abs(w, 0.0)
    got = S.cleared_force(len(w.epi)); covered = 0
    for r in range(world):
        f = np.load(tmp_path / ("f%d.npy" % r)); e0, e1, nb = np.load(tmp_path / ("r%d.npy" % r))
        got[e0:e1] = f; covered += e1 - e0
        launches, n_walks = np.load(tmp_path / ("l%d.npy" % r))
        assert launches == 1, (r, launches, n_walks)
    assert covered == len(w.epi)
    synth.assert_force_close(got, want, 1e-4, "%d ranks, peer, placed + fused pack" % world, cond=cond)


def _soft_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from gplum_b200 import disk, functors as F, state as ST, structs as S, tree
    from gplum_b200._lib import check, lib
    from gplum_b200.multigpu import MultiGpuSoftStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = 24000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=9)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    ro, rs = ro * 2.0, rs * 3.0                                 # neighbours across the rank boundary
    epj = ST.make_epj(d["pos"], d["vel"], d["mass"], ro, rs)
    F.init(rank)
    F.set_params(0.0, True, 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    check(lib().gplum_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    m = n // world
    ms = MultiGpuSoftStep(epj[rank * m:(rank + 1) * m], n, world, rank, n_group_limit=128)
    prm = S.corr_params()
    F.soft_corr_enable(True)
    try:
        for _ in range(2):
            sz = ms.step(prm)
        f = ms.forces()
        corr, ngb = F.correct_long_download_compact(n)
        g, order = tree.copy_walks_gpu(sz[:8])
    finally:
        F.soft_corr_enable(False)
    np.save(os.path.join(out_dir, "sf%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "sc%d.npy" % rank), corr)
    np.save(os.path.join(out_dir, "sn%d.npy" % rank), ngb)
    np.save(os.path.join(out_dir, "ss%d.npy" % rank), np.array(ms.share()))
    np.save(os.path.join(out_dir, "so%d.npy" % rank), order)
    dist.barrier()
    dist.destroy_process_group()


def _rec48_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from gplum_b200 import disk, functors as F
    from gplum_b200._lib import check, lib
    from gplum_b200.multigpu import MultiGpuSoftStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = 24000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=9)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    F.init(rank)
    F.set_params(0.0, True, 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    check(lib().gplum_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    m = n // world
    sl = slice(rank * m, (rank + 1) * m)
    rec = np.empty((m, 6))
    rec[:, :3] = d["pos"][sl]; rec[:, 3] = d["mass"][sl]; rec[:, 4] = ro[sl]; rec[:, 5] = rs[sl] * 2.0
    ms = MultiGpuSoftStep(rec, n, world, rank, n_group_limit=128)
    for _ in range(2):
        ms.step(None)
    np.save(os.path.join(out_dir, "rf%d.npy" % rank), ms.forces())
    np.save(os.path.join(out_dir, "rs%d.npy" % rank), np.array(ms.share()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < WORLD, reason="needs >= %d GPUs" % WORLD)
def test_multi_rank_force_pass_from_48_byte_records(tmp_path):
    """the e2e form of bench.py at N > 1: all-gather of {pos, mass, r_out, r_search} records, the same tree on every
    GPU, every rank its share of the walks; the union equals the single-rank oracle, the shares partition the disk."""
    import torch.multiprocessing as mp
    import oracle_api as O
    import synth
    from gplum_b200 import disk, structs as S, tree
    world = WORLD
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_rec48_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 24000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=9)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    h, _ = tree.build_walks(d["pos"], d["mass"], ro, rs * 2.0, n_group_limit=128)
    want, _ = O.calc_walks(h, 0.0)
    got = S.cleared_force(n); prev = 0
    for r in range(world):
        f = np.load(tmp_path / ("rf%d.npy" % r)); w0, w1, e0, e1 = np.load(tmp_path / ("rs%d.npy" % r))
        assert e0 == prev and len(f) == e1 - e0
        got[e0:e1] = f; prev = e1
    assert prev == n and want["number"].sum() > 100
    synth.assert_force_close(got, want, 1e-4, "multi-rank pass from 48 B records")


@pytest.mark.skipif(_ngpu() < WORLD, reason="needs >= %d GPUs" % WORLD)
def test_multi_rank_soft_step_without_host_lists(tmp_path):
    """all-gather of raw particles over NCCL -> the same tree on every GPU -> every rank its share of walks, forces,
    corrections and neighbour lists: the union equals the single-rank oracle (forces 1e-4, corrections 1e-12,
    neighbour lists as sets), and the shares partition the particles."""
    import torch.multiprocessing as mp
    import oracle_api as O
    import synth
    from gplum_b200 import disk, structs as S, tree
    world = WORLD
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_soft_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 24000
    d = disk.make_disk(n, a_in=0.98, a_out=1.02, seed=9)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    ro, rs = ro * 2.0, rs * 3.0
    h, order = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=128)
    h.epj_all["vel"] = d["vel"][order]
    want, _ = O.calc_walks(h, 0.0)
    prm = S.corr_params()
    oc, _, on = O.correct_long(h, prm, force=want)
    got = S.cleared_force(n); covered = 0; prev = 0
    n_ngb = 0
    for r in range(world):
        f = np.load(tmp_path / ("sf%d.npy" % r)); w0, w1, e0, e1 = np.load(tmp_path / ("ss%d.npy" % r))
        assert np.array_equal(np.load(tmp_path / ("so%d.npy" % r)), order)        # same tree on every rank
        assert e0 == prev and len(f) == e1 - e0
        got[e0:e1] = f; covered += e1 - e0; prev = e1
        corr = np.load(tmp_path / ("sc%d.npy" % r)); ngb = np.load(tmp_path / ("sn%d.npy" % r))
        # compact records = this rank's particles that have neighbours
        sel = np.nonzero(oc["number"][e0:e1] > 0)[0] + e0
        assert len(corr) == len(sel) and np.array_equal(corr["id_local"], oc["id_local"][sel])
        assert np.array_equal(corr["number"], oc["number"][sel])
        tot = np.linalg.norm(want["acc"][sel].astype(np.float64) + oc["acc"][sel], axis=1)[:, None]
        assert (np.abs(corr["acc"] - oc["acc"][sel]) <= 1e-12 * tot).all()
        for c, k in zip(corr, sel):
            a = set(ngb["id"][c["ngb_off"]:c["ngb_off"] + c["number"]].tolist())
            b = set(on["id"][oc["ngb_off"][k]:oc["ngb_off"][k] + oc["number"][k]].tolist())
            assert a == b
        n_ngb += int(corr["number"].sum())
    assert covered == n and n_ngb == int(oc["number"].sum()) and n_ngb > 100
    synth.assert_force_close(got, want, 1e-4, "multi-rank soft step")
