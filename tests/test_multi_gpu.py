"""Multi-GPU force pass on real GPUs (needs >= 2 devices; skipped on a 1-GPU box -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).  Two NCCL ranks shard the walks
of one disk, exchange EPJ (halo all-to-all and full all-gather variants) and evaluate their share
with the CUDA kernels; the union must match the single-rank oracle pass: acc/phi 1e-4, neighbour
info bit-exact."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _workload():
    from gplum_b200 import disk, tree
    d = disk.make_disk(20000, a_in=0.97, a_out=1.03, seed=5)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=128)
    return w


def _worker(rank, world, port, out_dir, exchange):
    import torch
    import torch.distributed as dist
    from gplum_b200 import functors as F
    from gplum_b200.multigpu import MultiGpuPass
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    w = _workload()
    F.init(rank)
    F.set_params(0.0, True, 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    import ctypes as C
    from gplum_b200._lib import check, lib
    check(lib().gplum_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    mg = MultiGpuPass(w, world, rank, stream, exchange=exchange)
    for _ in range(3):                      # repeated steps reuse the buffers: results must not drift
        mg.step()
    f = mg.forces()
    np.save(os.path.join(out_dir, "f%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(list(mg.sh.epi_range) + [mg.n_boundary]))
    mg.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["peer", "halo", "allgather"])
def test_two_ranks_match_single_rank_oracle(exchange, tmp_path):
    import torch.multiprocessing as mp
    import oracle_api as O
    import synth
    from gplum_b200 import structs as S
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), exchange), nprocs=world, join=True)
    w = _workload()
    want, _ = O.calc_walks(w, 0.0)
    got = S.cleared_force(len(w.epi)); covered = 0; n_bnd = 0
    for r in range(world):
        f = np.load(tmp_path / ("f%d.npy" % r)); e0, e1, nb = np.load(tmp_path / ("r%d.npy" % r))
        got[e0:e1] = f; covered += e1 - e0; n_bnd += nb
    assert covered == len(w.epi) and n_bnd > 0
    synth.assert_force_close(got, want, 1e-4, "2 ranks, " + exchange)
