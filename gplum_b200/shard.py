"""Multi-GPU partition of one force pass: i-groups sharded by spatial domain, j-data exchanged by
one all-gather (the B200 replacement of FDPS's LET exchange,
FDPS/src/tree_for_force_impl_exlet.hpp:343-403: allGather + 2x allToAllV of EPJ/SPJ).

Walks are in Morton order, so a contiguous range of walks is a compact spatial domain (what
`dinfo.decomposeDomainAll` gives each MPI rank in the reference).  Rank r owns
  * the i-particles / forces of walks [w0_r, w1_r)          -> no collective on the i side
  * the j-particles of the same Morton range (its "domain") -> packed locally, all-gathered
Superparticles are derived data (moments of the tree over the gathered particles): every rank
holds / rebuilds them itself, so only EPJ cross NVLink.
The all-gather buffer is slab-padded (every rank contributes `cap` records) so one in-place
`all_gather_into_tensor` moves everything; EP list indices are remapped once to the padded layout.
Walks whose EP list lies entirely inside the rank's own slab ("interior") can be evaluated while
the all-gather is in flight; the remaining "boundary" walks wait for it.
"""
import numpy as np

from .walks import Walks


def split_walks(w, world):
    """Contiguous walk ranges with ~equal interaction counts.  Returns [(w0, w1)] * world."""
    cost = w.ni.astype(np.int64) * (w.n_epj.astype(np.int64) * 20 + w.n_spj.astype(np.int64) * 38)
    c = np.cumsum(cost)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(c, c[-1] * r / world)))
    bounds.append(w.n_walk)
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def trim_spj(w, s0, s1):
    """A rank holds only the superparticles its own walks reference (in the reference each rank's
    spj_sorted_ is its own LET's, FDPS/src/tree_for_force_impl_exlet.hpp): returns the remapped
    index list, the compact SPJ array and the global indices it was taken from."""
    adr = w.adr_spj[s0:s1]
    used = np.unique(adr)
    return np.searchsorted(used, adr).astype(np.int32), w.spj_all[used], used


class Shard:
    """Everything rank `rank` needs: its walks (re-indexed to the padded gather layout), the
    range of j-particles / cells it packs, and the padded slab sizes."""

    def __init__(self, w, world, rank, pow2_cap=False):
        self.world, self.rank = world, rank
        ranges = split_walks(w, world)
        # particle (EPJ) ownership = Morton range of the rank's i-particles
        p_lo = [int(w.epi_off[a]) if a < w.n_walk else len(w.epi) for a, _ in ranges]
        p_lo[0] = 0
        p_hi = p_lo[1:] + [len(w.epj_all)]
        self.epj_ranges = list(zip(p_lo, p_hi))
        self.epj_cap = max(b - a for a, b in self.epj_ranges)
        if pow2_cap:            # peer mode: index = (owner << shift) | local
            self.shift = max(1, int(self.epj_cap - 1).bit_length())
            self.epj_cap = 1 << self.shift
        self.walk_range = ranges[rank]
        w0, w1 = ranges[rank]
        e0 = int(w.epi_off[w0]) if w0 < w.n_walk else len(w.epi)
        e1 = int(w.epi_off[w1 - 1] + w.ni[w1 - 1]) if w1 > w0 else e0
        self.epi_range = (e0, e1)
        a0 = int(w.epj_disp[w0]) if w1 > w0 else 0
        a1 = int(w.epj_disp[w1 - 1] + w.n_epj[w1 - 1]) if w1 > w0 else 0
        s0 = int(w.spj_disp[w0]) if w1 > w0 else 0
        s1 = int(w.spj_disp[w1 - 1] + w.n_spj[w1 - 1]) if w1 > w0 else 0
        self.adr_epj_range, self.adr_spj_range = (a0, a1), (s0, s1)
        adr_e = self.remap(w.adr_epj[a0:a1], self.epj_ranges, self.epj_cap)
        adr_s, spj_loc, self.spj_used = trim_spj(w, s0, s1)
        self.local = Walks(w.epi[e0:e1], w.epi_off[w0:w1] - e0, w.ni[w0:w1],
                           adr_e, w.epj_disp[w0:w1] - a0, w.n_epj[w0:w1],
                           adr_s, w.spj_disp[w0:w1] - s0, w.n_spj[w0:w1],
                           w.epj_all[self.epj_ranges[rank][0]:self.epj_ranges[rank][1]], spj_loc)
        # interior walks: every EP index inside this rank's own slab of the gather buffer
        lo, hi = rank * self.epj_cap, rank * self.epj_cap + (self.epj_ranges[rank][1] - self.epj_ranges[rank][0])
        lw = self.local
        inside = ((adr_e >= lo) & (adr_e < hi)).astype(np.int64)
        c = np.concatenate([[0], np.cumsum(inside)])
        d0 = lw.epj_disp
        self.interior = (c[d0 + lw.n_epj] - c[d0]) == lw.n_epj
        self.walks_interior = self.subset(np.nonzero(self.interior)[0])
        self.walks_boundary = self.subset(np.nonzero(~self.interior)[0])

    def subset(self, idx):
        """The walks `idx` of the local set, sharing its epi / list arrays (force indexing unchanged)."""
        lw = self.local
        return Walks(lw.epi, lw.epi_off[idx], lw.ni[idx], lw.adr_epj, lw.epj_disp[idx], lw.n_epj[idx],
                     lw.adr_spj, lw.spj_disp[idx], lw.n_spj[idx], lw.epj_all, lw.spj_all)

    @staticmethod
    def remap(adr, ranges, cap):
        """global index -> index in the slab-padded all-gather buffer (rank r's slab at r*cap)."""
        lo = np.array([a for a, _ in ranges], dtype=np.int64)
        owner = np.searchsorted(lo, adr, side="right") - 1
        return (adr - lo[owner] + owner * cap).astype(np.int32)


class HaloShard:
    """Same partition as `Shard`, but the exchange is trimmed to what is needed -- the LET idea of
    the reference (FDPS/src/tree_for_force_impl_exlet.hpp:343-403: every rank sends each other
    rank only the EPJ that rank's walks can touch) instead of a full all-gather.

    Local j-array of rank r:  [ own particles (n_own) | halo from rank 0 | halo from rank 1 | ... ]
    where "halo from s" are the particles owned by s that appear in r's EP lists, in ascending
    global order.  Per step: pack own particles in place, gather `send_idx` into a send buffer,
    one all-to-all (`send_counts` / `recv_counts` records), then the boundary walks.  Interior
    walks (all EP indices own) run while the exchange is in flight.
    """

    def __init__(self, w, world, rank):
        self.world, self.rank = world, rank
        ranges = split_walks(w, world)
        p_lo = [int(w.epi_off[a]) if a < w.n_walk else len(w.epi) for a, _ in ranges]
        p_lo[0] = 0
        p_hi = p_lo[1:] + [len(w.epj_all)]
        self.epj_ranges = list(zip(p_lo, p_hi))
        lo = np.array(p_lo, dtype=np.int64)

        def lists_of(r):
            w0, w1 = ranges[r]
            if w1 <= w0:
                return 0, 0, np.zeros(0, np.int32)
            a0 = int(w.epj_disp[w0]); a1 = int(w.epj_disp[w1 - 1] + w.n_epj[w1 - 1])
            return a0, a1, w.adr_epj[a0:a1]

        def needs_of(r):
            """Sorted global indices rank r reads from other ranks."""
            _, _, adr = lists_of(r)
            q0, q1 = self.epj_ranges[r]
            return np.unique(adr[(adr < q0) | (adr >= q1)])

        p0, p1 = self.epj_ranges[rank]
        self.n_own = p1 - p0
        need = needs_of(rank)
        owner = np.searchsorted(lo, need, side="right") - 1
        self.recv_counts = [int((owner == s).sum()) for s in range(world)]
        self.n_halo = len(need)
        send_idx, self.send_counts = [], []
        for d in range(world):
            if d == rank:
                self.send_counts.append(0)
                continue
            nd = needs_of(d)
            mine = nd[(nd >= p0) & (nd < p1)] - p0
            send_idx.append(mine.astype(np.int32))
            self.send_counts.append(len(mine))
        self.send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, np.int32)

        w0, w1 = ranges[rank]
        self.walk_range = (w0, w1)
        e0 = int(w.epi_off[w0]) if w0 < w.n_walk else len(w.epi)
        e1 = int(w.epi_off[w1 - 1] + w.ni[w1 - 1]) if w1 > w0 else e0
        self.epi_range = (e0, e1)
        a0, a1, adr = lists_of(rank)
        s0 = int(w.spj_disp[w0]) if w1 > w0 else 0
        s1 = int(w.spj_disp[w1 - 1] + w.n_spj[w1 - 1]) if w1 > w0 else 0
        self.adr_epj_range, self.adr_spj_range = (a0, a1), (s0, s1)
        own = (adr >= p0) & (adr < p1)
        adr_l = np.where(own, adr - p0, self.n_own + np.searchsorted(need, adr)).astype(np.int32)
        self.need = need
        adr_s, spj_loc, self.spj_used = trim_spj(w, s0, s1)
        self.local = Walks(w.epi[e0:e1], w.epi_off[w0:w1] - e0, w.ni[w0:w1],
                           adr_l, w.epj_disp[w0:w1] - a0, w.n_epj[w0:w1],
                           adr_s, w.spj_disp[w0:w1] - s0, w.n_spj[w0:w1],
                           w.epj_all[p0:p1], spj_loc)
        lw = self.local
        c = np.concatenate([[0], np.cumsum(own.astype(np.int64))])
        d0 = lw.epj_disp
        self.interior = (c[d0 + lw.n_epj] - c[d0]) == lw.n_epj
        self.walks_interior = self.subset(np.nonzero(self.interior)[0])
        self.walks_boundary = self.subset(np.nonzero(~self.interior)[0])

    subset = Shard.subset

    def rank_inputs(self, w):
        """What this rank of an MPI-FDPS run would hand to the dispatch functor: its own walks, with
        `epj_sorted_` = its local particles followed by its LET imports (the halo records, in the order the
        lists index them) and `spj_sorted_` = the superparticles its walks reference
        (FDPS/src/tree_for_force_impl_exlet.hpp:343-403).  Same lists, same records, so the forces of these
        walks equal the single-rank pass bit for bit."""
        lw = self.local
        epj = np.concatenate([lw.epj_all, w.epj_all[self.need]])
        return Walks(lw.epi, lw.epi_off, lw.ni, lw.adr_epj, lw.epj_disp, lw.n_epj, lw.adr_spj, lw.spj_disp, lw.n_spj,
                     epj, lw.spj_all)
