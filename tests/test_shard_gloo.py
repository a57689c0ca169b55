"""N>1 host logic on CPU (gloo, world_size 2 and 3): walk sharding by domain, slab-padded
all-gather of j-data, index remap.  Each rank evaluates only its own walks (with the oracle
standing in for the kernel) on the GATHERED j-arrays; the union must equal the single-rank
pass bit-for-bit -- i.e. the partition changes nothing but who computes what."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_api as O
from gplum_b200 import disk, structs as S, tree
from gplum_b200.shard import HaloShard, Shard, split_walks


def _workload():
    d = disk.make_disk(3000, a_in=0.97, a_out=1.03, seed=12)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=64)
    return w


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _workload()
    sh = Shard(w, world, rank)
    lw = sh.local
    # "pack" = this rank's own j-records, padded to the slab size; all-gather the slabs
    def gather(local, cap, dtype):
        send = np.zeros(cap, dtype=dtype); send[:len(local)] = local
        t_send = torch.from_numpy(send.view(np.uint8).copy())
        t_all = torch.zeros(world * t_send.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(t_all, t_send)
        return t_all.numpy().view(dtype)
    epj_g = gather(lw.epj_all, sh.epj_cap, S.EPJ)
    # remapped lists address the same particles as the global lists
    a0, a1 = sh.adr_epj_range
    assert (epj_g[lw.adr_epj].tobytes() == w.epj_all[w.adr_epj[a0:a1]].tobytes())
    # interior walks need nothing from other ranks: evaluate them on the rank's OWN slab only
    own = np.zeros_like(epj_g); lo = rank * sh.epj_cap
    own[lo:lo + len(lw.epj_all)] = lw.epj_all
    f = S.cleared_force(len(lw.epi)); n_int = 0
    for part, jarr in ((sh.walks_interior, own), (sh.walks_boundary, epj_g)):
        pw = O.Walks(part.epi, part.epi_off, part.ni, part.adr_epj, part.epj_disp, part.n_epj, part.adr_spj,
                     part.spj_disp, part.n_spj, jarr, lw.spj_all)
        f, n = O.calc_walks(pw, 0.0, force=f)
        n_int += n
    assert sh.interior.sum() + (~sh.interior).sum() == lw.n_walk and (world == 1 or (~sh.interior).any())
    np.save(os.path.join(out_dir, "f%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(list(sh.epi_range) + [n_int]))
    dist.barrier()
    dist.destroy_process_group()


def _worker_halo(rank, world, port, out_dir):
    """Trimmed exchange: own particles + one all-to-all of the records other ranks' walks need."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _workload()
    sh = HaloShard(w, world, rank)
    lw = sh.local
    nb = S.EPJ.itemsize
    send = torch.from_numpy(lw.epj_all[sh.send_idx].view(np.uint8).reshape(-1, nb).copy())
    recv = torch.zeros((sh.n_halo, nb), dtype=torch.uint8)
    dist.all_to_all_single(recv, send, sh.recv_counts, sh.send_counts)
    jl = np.concatenate([lw.epj_all, recv.numpy().reshape(-1).view(S.EPJ)])
    a0, a1 = sh.adr_epj_range
    assert jl[lw.adr_epj].tobytes() == w.epj_all[w.adr_epj[a0:a1]].tobytes()
    assert sum(sh.recv_counts) == sh.n_halo and sh.n_halo < len(w.epj_all) - sh.n_own   # trimmed, not everything
    own_only = jl.copy(); own_only[sh.n_own:] = 0                 # interior walks must not touch the halo
    f = S.cleared_force(len(lw.epi)); n_int = 0
    for part, jarr in ((sh.walks_interior, own_only), (sh.walks_boundary, jl)):
        pw = O.Walks(part.epi, part.epi_off, part.ni, part.adr_epj, part.epj_disp, part.n_epj, part.adr_spj,
                     part.spj_disp, part.n_spj, jarr, lw.spj_all)
        f, n = O.calc_walks(pw, 0.0, force=f)
        n_int += n
    np.save(os.path.join(out_dir, "f%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(list(sh.epi_range) + [n_int]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,worker", [(2, _worker), (3, _worker), (2, _worker_halo), (3, _worker_halo)])
def test_sharded_pass_equals_single_rank(world, worker, tmp_path):
    mp.spawn(worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    w = _workload()
    want, n_tot = O.calc_walks(w, 0.0)
    got = S.cleared_force(len(w.epi)); covered = 0; n_sum = 0
    for r in range(world):
        f = np.load(tmp_path / ("f%d.npy" % r)); e0, e1, n_int = np.load(tmp_path / ("r%d.npy" % r))
        assert len(f) == e1 - e0
        got[e0:e1] = f; covered += e1 - e0; n_sum += n_int
    assert covered == len(w.epi) and n_sum == n_tot
    assert got.tobytes() == want.tobytes()


def test_split_is_contiguous_and_balanced():
    w = _workload()
    for world in (1, 2, 4, 8):
        r = split_walks(w, world)
        assert r[0][0] == 0 and r[-1][1] == w.n_walk and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        cost = w.ni.astype(np.int64) * (w.n_epj * 20 + w.n_spj * 38)
        per = [cost[a:b].sum() for a, b in r]
        assert max(per) <= 1.35 * (sum(per) / world) + cost.max()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_rank_inputs_reproduce_the_single_rank_pass(world):
    """The per-rank dispatch inputs of bench.py's multi-GPU e2e leg (own walks, local + LET particles, trimmed
    SPJ): evaluated rank by rank with the oracle they give the single-rank forces bit for bit."""
    import oracle_api as O
    from gplum_b200 import disk, tree
    from gplum_b200.shard import HaloShard
    d = disk.make_disk(6000, a_in=0.97, a_out=1.03, seed=11)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs * 2.0, n_group_limit=64)
    want, n_int = O.calc_walks(w, 0.0)
    seen, tot = 0, 0
    for r in range(world):
        sh = HaloShard(w, world, r)
        lw = sh.rank_inputs(w)
        assert len(lw.epj_all) == sh.n_own + sh.n_halo < len(w.epj_all) or world == 1
        assert lw.adr_epj.max(initial=0) < len(lw.epj_all) and lw.adr_spj.max(initial=0) < len(lw.spj_all)
        got, n = O.calc_walks(lw, 0.0)
        e0, e1 = sh.epi_range
        assert got.tobytes() == want[e0:e1].tobytes(), r
        seen += e1 - e0; tot += n
    assert seen == len(w.epi) and tot == n_int
