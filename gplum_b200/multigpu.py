"""One rank's share of a multi-GPU force pass: i-groups sharded by domain, EPJ exchanged
over NCCL, interior walks overlapped with the exchange, boundary walks on a second stream.

Replaces FDPS's LET exchange + per-rank calcForce (FDPS/src/tree_for_force_impl_exlet.hpp:343-403,
tree_for_force_impl_force.hpp:1404-1564).  torch / torch.distributed are plumbing (streams, NCCL);
every kernel on the path is libgplum_b200's.  Used by bench.py and tests/test_multi_gpu.py.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import functors as F, structs as S
from ._lib import check, lib
from .shard import HaloShard, Shard


class _PeerFlags:
    """What exchange() returns in peer mode: wait() enqueues the flag-wait kernel on the library stream."""

    def __init__(self, L):
        self.L = L

    def wait(self):
        check(self.L.gplum_b200_peer_wait())


class MultiGpuPass:
    """`exchange` =
    "peer"      no data moves ahead of time and no collective runs per step: every rank packs into
                its own slab, the slabs are mapped into every process with CUDA IPC and boundary
                walks gather the records they need straight from the owner's HBM over NVLink, tile
                by tile, inside the force kernel.  A step is TWO launches: the pack kernel (EPJ slab,
                superparticles, and "packed" flag words stored into every rank's memory over NVLink)
                and one force launch over all walks, whose boundary items spin on those flags inside
                the kernel while the interior items already run (slabs are double-buffered, so this
                one barrier per step also covers the write-after-read hazard);
    "halo"      one all-to-all of only the records other ranks' boundary walks read;
    "allgather" in-place all-gather of every rank's packed slab."""

    def __init__(self, w, world, rank, stream, exchange="peer", boundary_cap=32):
        L = lib()
        self.L, self.stream, self.exchange_kind = L, stream, exchange
        halo, peer = exchange == "halo", exchange == "peer"
        self.sh = sh = HaloShard(w, world, rank) if halo else Shard(w, world, rank, pow2_cap=peer)
        self.lw = lw = sh.local
        eb, sb = C.c_int(0), C.c_int(0)
        L.gplum_b200_packed_sizes(C.byref(eb), C.byref(sb))
        EB = eb.value
        # raw AoS particles of this rank's own domain, resident in HBM (inputs of the step)
        self.d_epj_raw = torch.from_numpy(lw.epj_all.view(np.uint8).copy()).cuda()
        wi, wb = sh.walks_interior, sh.walks_boundary
        self.n_interior, self.n_boundary = wi.n_walk, wb.n_walk
        vp = lambda t: C.c_void_p(t.data_ptr())
        if peer:
            # the handles first: a walk set uploaded while peer mode is open flags the walks that read other ranks
            h = (C.c_char * 128)()
            check(L.gplum_b200_peer_setup(world, rank, sh.shift, h))
            mine = torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).cuda()
            every = torch.zeros(world * 128, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(every, mine)
            self._handles = every.cpu().numpy().tobytes()
            check(L.gplum_b200_peer_open(self._handles))
            self.exchange_bytes = 4 * world
            # ONE walk set (interior and boundary walks in one work list, one launch per step) + the superparticles,
            # which every rank holds itself
            F.walks_select(0)
            F.walks_upload(type(lw)(lw.epi, lw.epi_off, lw.ni, lw.adr_epj, lw.epj_disp, lw.n_epj, lw.adr_spj,
                                    lw.spj_disp, lw.n_spj, np.zeros(0, S.EPJ), lw.spj_all))
            dist.barrier()
        else:
            # slot 0: interior walks (+ the superparticles, which every rank holds itself); slot 1: boundary
            F.walks_select(0)
            F.walks_upload(type(lw)(wi.epi, wi.epi_off, wi.ni, wi.adr_epj, wi.epj_disp, wi.n_epj, wi.adr_spj,
                                    wi.spj_disp, wi.n_spj, np.zeros(0, S.EPJ), lw.spj_all))
            F.walks_select(1)
            # the boundary set is small, but it co-runs with the interior kernel on a busy GPU: throughput,
            # not one item's latency, is what counts -> full-width tiles instead of the per-pass choice
            check(L.gplum_b200_set_tile_cap(boundary_cap))
            F.walks_upload(wb, with_j=False)
            check(L.gplum_b200_set_tile_cap(0))
        vp = lambda t: C.c_void_p(t.data_ptr())
        if halo:
            # local j-array = [own | halo]: own particles are packed in place, the halo region is the
            # receive buffer of one all-to-all of the records other ranks' boundary walks need (LET)
            n_own, n_halo, n_send = sh.n_own, sh.n_halo, len(sh.send_idx)
            self.jbuf = torch.zeros((n_own + n_halo) * EB + 16, dtype=torch.uint8, device="cuda")
            self.my_slab = self.jbuf[:n_own * EB]
            self.halo_rows = self.jbuf[n_own * EB:(n_own + n_halo) * EB].view(n_halo, EB)
            self.send_rows = torch.zeros((max(n_send, 1), EB), dtype=torch.uint8, device="cuda")[:n_send]
            self.d_send_idx = torch.from_numpy(np.ascontiguousarray(sh.send_idx)).cuda()
            check(L.gplum_b200_walks_set_packed_dev(vp(self.jbuf), n_own + n_halo, None, 0))
            self.exchange_bytes = (n_send + n_halo) * EB
        elif peer:
            pass
        else:
            # the gather buffer: every rank's packed slab; this rank packs straight into its own slab
            self.jbuf = torch.zeros(world * sh.epj_cap * EB, dtype=torch.uint8, device="cuda")
            self.my_slab = self.jbuf[rank * sh.epj_cap * EB:(rank + 1) * sh.epj_cap * EB]
            check(L.gplum_b200_walks_set_packed_dev(vp(self.jbuf), world * sh.epj_cap, None, 0))
            self.exchange_bytes = int(self.jbuf.numel())
        # boundary walks: start when the exchange lands; HIGH priority so that their CTAs are scheduled
        # ahead of the interior kernel's not-yet-resident CTAs and the two kernels really co-run
        self.side = torch.cuda.Stream(priority=-1)
        self.ev_side, self.ev_pack = torch.cuda.Event(), torch.cuda.Event()

    def _use(self, st):
        check(self.L.gplum_b200_set_stream(C.c_void_p(st.cuda_stream)))

    def exchange(self):
        """Pack own EPJ, start the NCCL exchange; returns the async work handle."""
        L, vp = self.L, (lambda t: C.c_void_p(t.data_ptr()))
        if self.exchange_kind == "peer":
            check(L.gplum_b200_peer_pack(vp(self.d_epj_raw), len(self.lw.epj_all)))    # pack + signal the peers
            return _PeerFlags(L)
        check(L.gplum_b200_pack_epj_dev(vp(self.d_epj_raw), len(self.lw.epj_all), vp(self.my_slab)))
        if self.exchange_kind == "halo":
            check(L.gplum_b200_gather_epj_packed_dev(vp(self.my_slab), vp(self.d_send_idx), len(self.sh.send_idx),
                                                     vp(self.send_rows)))
            return dist.all_to_all_single(self.halo_rows, self.send_rows, self.sh.recv_counts, self.sh.send_counts,
                                          async_op=True)
        return dist.all_gather_into_tensor(self.jbuf, self.my_slab, async_op=True)

    def step(self, trace=None):
        """One force pass of this rank.  trace: optional dict that receives timing events."""
        stream, side = self.stream, self.side
        if self.exchange_kind == "peer":
            # pack (EPJ slab + superparticles + flags to the peers, one launch) and ONE force launch: interior items
            # start at once, boundary items wait for the peers' flags inside the kernel
            if trace is not None:
                ev = {k: torch.cuda.Event(enable_timing=True) for k in ("start", "packed", "int_end", "side_ready", "bnd_end")}
                ev["start"].record(stream)
            self.exchange()
            if trace is not None:
                ev["packed"].record(stream)
            F.walks_select(0)
            F.walks_run(repack=False)
            if trace is not None:
                for k in ("int_end", "side_ready", "bnd_end"):
                    ev[k].record(stream)
                trace["events"] = ev
            return
        if trace is not None:
            ev = {k: torch.cuda.Event(enable_timing=True) for k in ("start", "packed", "int_end", "side_ready", "bnd_end")}
            ev["start"].record(stream)
        work = self.exchange()
        check(self.L.gplum_b200_walks_pack())      # SPJ pack (every rank holds the cells itself)
        self.ev_pack.record(stream)
        if trace is not None:
            ev["packed"].record(stream)
        F.walks_select(0)
        F.walks_run(repack=False)          # interior walks: overlap the exchange
        if trace is not None:
            ev["int_end"].record(stream)
        with torch.cuda.stream(side):
            side.wait_event(self.ev_pack)  # the side stream waits for the packed SPJ ...
            self._use(side)
            work.wait()                    # ... and for the exchange (NCCL work)
            if trace is not None:
                ev["side_ready"].record(side)
            F.walks_select(1)
            F.walks_run(repack=False)      # boundary walks read the other ranks' particles
            self.ev_side.record(side)
            if trace is not None:
                ev["bnd_end"].record(side)
        self._use(stream)
        stream.wait_event(self.ev_side)
        if trace is not None:
            trace["events"] = ev

    @staticmethod
    def trace_ms(trace):
        torch.cuda.synchronize()
        ev = trace["events"]
        return {k: round(ev["start"].elapsed_time(ev[k]), 4) for k in ("packed", "int_end", "side_ready", "bnd_end")}

    def forces(self):
        """This rank's forces (host, ForceGrav[ len(local.epi) ]) after step()."""
        torch.cuda.synchronize()
        n = len(self.lw.epi)
        if self.exchange_kind == "peer":
            F.walks_select(0)
            return F.walks_download(n)
        out = S.cleared_force(n)
        for slot, ws in ((0, self.sh.walks_interior), (1, self.sh.walks_boundary)):
            if ws.n_walk == 0:
                continue
            F.walks_select(slot)
            f = F.walks_download(int((ws.epi_off + ws.ni).max()))
            for k in range(ws.n_walk):
                sl = slice(int(ws.epi_off[k]), int(ws.epi_off[k] + ws.ni[k]))
                out[sl] = f[sl]
        F.walks_select(0)
        return out

    def close(self):
        torch.cuda.synchronize()
        if self.exchange_kind == "peer":
            dist.barrier()                         # nobody reads a peer slab any more
            check(self.L.gplum_b200_peer_close())
            dist.barrier()                         # every mapping is closed before a slab is freed
            check(self.L.gplum_b200_peer_free())
        check(self.L.gplum_b200_walks_set_packed_dev(None, 0, None, 0))
        F.walks_select(0)


class MultiGpuSoftStep:
    """One soft-force evaluation of a multi-GPU run with NO host-side lists and no rank touching the global particle
    set on the host (SURVEY 8e + 8f-1/2): every rank holds n / world particles in HBM; per step

        NCCL all-gather of the raw EPJGrav records over NVLink   (FDPS's LET exchange, exlet.hpp:343-403: here every
                                                                  rank gets every particle -- 112 MB at N = 1e6)
        gplum_b200_tree_build_gpu_part                            the same tree on every GPU, lists for this rank's walks
        walks_run + correct_long_run                              forces, changeover correction, neighbour lists: own share

    The share of rank r is the Morton-contiguous range of walks whose first particle lies in [n r / W, n (r+1) / W) of
    the tree order -- a compact spatial domain, like an FDPS domain."""

    def __init__(self, epj_local, n_total, world, rank, theta=0.5, n_leaf_limit=8, n_group_limit=64):
        assert n_total % world == 0 and len(epj_local) == n_total // world, "equal shares (all_gather_into_tensor)"
        self.L = lib()
        self.world, self.rank, self.n = world, rank, n_total
        self.theta, self.leaf, self.group = theta, n_leaf_limit, n_group_limit
        # EPJGrav records (112 B: the correction reads vel / acc_d / id), or -- for the force pass alone -- 48 B records
        # {pos[3], mass, r_out, r_search}: an [m, 6] float64 array
        local = np.ascontiguousarray(epj_local)
        self.rec48 = local.dtype == np.float64
        assert self.rec48 and local.shape[1:] == (6,) or local.dtype == S.EPJ
        self.rec_bytes = 48 if self.rec48 else S.EPJ.itemsize
        self.d_local = torch.from_numpy(local.view(np.uint8).reshape(-1).copy()).cuda()
        self.d_all = torch.empty(n_total * self.rec_bytes, dtype=torch.uint8, device="cuda")
        self.sizes = np.zeros(12, dtype=np.int64)

    def upload_local(self, epj_local_pinned):
        """host -> device of this rank's records (a step's H2D when the particles live on the host)"""
        self.d_local.copy_(torch.from_numpy(epj_local_pinned.view(np.uint8).reshape(-1)), non_blocking=True)

    def step(self, prm=None):
        dist.all_gather_into_tensor(self.d_all, self.d_local)
        build = self.L.gplum_b200_tree_build_gpu_part_rec48 if self.rec48 else self.L.gplum_b200_tree_build_gpu_part
        check(build(self.n, C.c_void_p(self.d_all.data_ptr()), float(self.theta), int(self.leaf), int(self.group), self.rank, self.world,
                    self.sizes.ctypes.data_as(C.c_void_p)))
        F.walks_run(repack=False)
        if prm is not None:
            F.correct_long_run(prm)
        return self.sizes

    def share(self):
        """(w0, w1, e0, e1): this rank's walks and i-particles (tree order) of the last step"""
        return tuple(int(x) for x in self.sizes[8:12])

    def forces(self, out=None):
        """ForceGrav[e1 - e0] of this rank's particles (tree order)."""
        _, _, e0, e1 = self.share()
        f = np.zeros(e1 - e0, dtype=S.FORCE) if out is None else out
        check(self.L.gplum_b200_walks_download_range(f.ctypes.data_as(C.c_void_p), e0, e1 - e0))
        return f[:e1 - e0]
