// items.h -- how one walk (i-group) is cut into i-tiles ("work items") of the force pass.
// Shared by the host work-list builder (gplum_b200.cu: build_items) and the device one
// (dev_tree.cu: emit_items_kernel / split_items_kernel) so both produce the same items and the same cost keys.
#pragma once

#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

namespace gb {

// kernel configuration code of a tile shape (kernels.cuh: force_pass_kernel decodes it)
GB_HD int tile_cfg_of(int c) { return c == 64 ? 1 : c == 32 ? 0 : c == 16 ? 9 : c == 8 ? 10 : 11; }

// Next tile of a walk with `rem` i-particles left (rmax <= 2 builds): n = i-particles it takes,
// shape = lanes x registers capacity of the tile (64 / 32 / 16 / 8 / 4).
// Finer decompositions of a remainder (e.g. 20 -> 16 + 4) were measured slower at
// n_group_limit = 64: every extra item pays its own staging.
GB_HD void tile_next(int rem, int cap, bool split, int &n, int &shape)
{
    if (rem >= cap) { n = cap; shape = cap; return; }
    if (!split) { n = rem; shape = rem > 32 ? 64 : 32; return; }
    if (rem > 32 && rem <= 48) { n = 32; shape = 32; return; }      // 32 + a j-split tail beats a half-empty 64
    n = rem;
    shape = rem > 32 ? 64 : rem > 16 ? 32 : rem > 8 ? 16 : rem > 4 ? 8 : 4;
}

// cost model in issue slots per lane (EP-EP 15.5, EP-SP 34 per pair; 3 more each with the Newton step compiled in,
// kernels.cuh: GB_NEWTON) + per-tile staging overhead
#ifndef GB_NEWTON
#define GB_NEWTON 0
#endif
constexpr double COST_EP = GB_NEWTON ? 18.5 : 15.5, COST_SP = GB_NEWTON ? 37.0 : 34.0, COST_TILE = 90.0;
#ifndef GB_COST_ITEM
#define GB_COST_ITEM 4000.0
#endif
// per item: its prologue is a chain of dependent global loads (item -> walk record -> index list -> particles),
// several microseconds in which the warp issues nothing; 4000 slots measured best for the segmented layout
constexpr double COST_ITEM = GB_COST_ITEM;
GB_HD double tile_cost(int n_epj, int n_spj, int shape)
{
    const double cost_j = COST_EP * n_epj + COST_SP * n_spj;
    const double cost_tiles = COST_TILE * ((n_epj + 63) / 64 + (n_spj + 63) / 64) + COST_ITEM;
    return cost_j * shape / 32.0 + cost_tiles;
}

// ---- j-split of full-width tiles: a pass with few items (a rank's share of a multi-GPU run, a small dispatch)
// cannot fill the GPU with whole tiles -- 2300 items on 3552 warp slots leave a third of the issue slots empty and
// the pass ends when the longest serial chains do.  Such a pass cuts every full-width tile (64 or 32 i-particles)
// into K parts along its j-tile sequence (EP tiles first, then SP tiles); every part is its own work item on its
// own warp, writes its partial sums to a scratch record, and the part that finishes LAST adds the K partial sums
// in part order (kernels.cuh, warp_force) -- a fixed order, so the result does not depend on which warp finishes
// when.  Cfg word of a final item: bits 0-3 tile shape, bit 6 waits for the peers' flags (multi-GPU peer mode),
// bits 8-15 K, bits 16-23 part index.
constexpr int ITEM_PEER_WAIT = 64;
constexpr int SPLIT_K_MAX = 250;       // parts of one tile (the cfg word holds 8 bits of it)
GB_HD int item_parts(int cfg) { return (cfg >> 8) & 0xff; }           // 0 or 1: unsplit
GB_HD int item_part_index(int cfg) { return (cfg >> 16) & 0xff; }

// number of j-tiles of a walk and the cost of the first t of them for a tile shape (EP tiles, then SP tiles)
GB_HD int walk_tiles(int n_epj, int n_spj) { return (n_epj + 63) / 64 + (n_spj + 63) / 64; }
GB_HD double tiles_cost(int n_epj, int n_spj, int shape, int t)
{
    const int nt_ep = (n_epj + 63) / 64;
    const int te = t < nt_ep ? t : nt_ep, ts = t - te;
    const int je = te * 64 < n_epj ? te * 64 : n_epj, js = ts * 64 < n_spj ? ts * 64 : n_spj;
    return (COST_EP * je + COST_SP * js) * shape / 32.0 + COST_TILE * t;
}

// Segments: a pass with less than two waves of base items is laid out as ONE WAVE of warps with equal work.  The
// base items in list order form a line of cost (integer units of the cost model); the line is cut into n_seg equal
// segments, one per resident warp slot, at j-tile boundaries: a segment is a run of consecutive work items (the tail
// part of one tile, whole tiles, the head part of another) that one warp executes back to back.  All warps start
// together and finish together -- no scheduling tail, every issue port has its full set of warps to the end.
// (Equal PARTS of every tile were measured first: 2-4 parts per tile did not beat whole tiles on a 1/8 shard,
// because the pass still ended on a tail of single warps; profiles/r2_split_probe.txt.)
// Measured on 1/4, 1/8 and 1/16 shards of the N = 1e6 disk (profiles/r2_split_probe.txt): with 0.3 waves of tiles or
// more, whole tiles under the hardware's dynamic CTA scheduling are faster (short items' warps leave early and the
// rest speed up); below that one wave of segments wins (0.081 against 0.095 ms at 1/16).  split_m scales the limit.
GB_HD bool split_active(long long n_base, long long warp_slots, int split_m) { return split_m > 0 && n_base > 0 && n_base * 8 < warp_slots * split_m; }
GB_HD long long item_cost_units(double cost) { return (long long)(cost + 0.5); }
// segment of cost position x on a line of total cost W cut into n_seg pieces; boundary b sits at ceil(b W / n_seg)
GB_HD long long seg_of(long long x, long long W, long long n_seg) { const long long s = (x * n_seg) / W; return s < n_seg ? s : n_seg - 1; }   // W n_seg < 2^63: passes that split are small
GB_HD long long seg_boundary(long long b, long long W, long long n_seg) { return (b * W + n_seg - 1) / n_seg; }
// j-tile boundary nearest to the point where a tile of shape `shape` has spent `offset` cost units
GB_HD int split_tile_at(int n_epj, int n_spj, int shape, long long offset)
{
    const int nt = walk_tiles(n_epj, n_spj);
    int lo = 0, hi = nt;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (tiles_cost(n_epj, n_spj, shape, mid) + COST_ITEM < (double)offset) lo = mid + 1; else hi = mid; }
    // the nearer of the two tile boundaries around the cut
    if (lo > 0 && (double)offset - (tiles_cost(n_epj, n_spj, shape, lo - 1) + COST_ITEM) < (tiles_cost(n_epj, n_spj, shape, lo) + COST_ITEM) - (double)offset) lo--;
    return lo;
}
// The parts of one base item on the segmented line, generated one after the other (host and device builders run
// the same code): C = cost units before the item, c = its own.  Lane-split tails, one-tile walks and tiles that lie
// inside one segment stay whole; every tile shape of the RMAX = 2 build can be cut, lane-split tails included.
struct SegCut {
    long long W, n_seg, C, s1, sb;
    int n_epj, n_spj, shape, nt, t_prev, emitted;
    bool whole, done;
};
GB_HD void seg_cut_begin(SegCut &q, long long W, long long n_seg, long long C, long long c, int cfg, int n_epj, int n_spj)
{
    q.W = W; q.n_seg = n_seg; q.C = C; q.n_epj = n_epj; q.n_spj = n_spj;
    const int kc = cfg & 15;
    q.shape = kc == 1 ? 64 : kc == 0 ? 32 : kc == 9 ? 16 : kc == 10 ? 8 : 4;
    q.nt = walk_tiles(n_epj, n_spj);
    const long long s0 = seg_of(C, W, n_seg);
    q.s1 = c > 0 ? seg_of(C + c - 1, W, n_seg) : s0;
    q.sb = s0 + 1; q.t_prev = 0; q.emitted = 0; q.done = false;
    q.whole = (kc > 1 && kc < 9) || q.s1 == s0 || q.nt <= 1;      // RMAX = 4 shapes are not cut
}
// next part [t0, t1) and its segment; false when the item is exhausted.  A whole item reports t0 = 0, t1 = nt.
GB_HD bool seg_cut_next(SegCut &q, int &t0, int &t1, long long &seg)
{
    if (q.done) return false;
    if (q.whole) { t0 = 0; t1 = q.nt; seg = seg_of(q.C, q.W, q.n_seg); q.done = true; q.emitted = 1; return true; }
    while (q.sb <= q.s1 && q.emitted < SPLIT_K_MAX - 1) {
        const long long sb = q.sb++;
        int t_cut = split_tile_at(q.n_epj, q.n_spj, q.shape, seg_boundary(sb, q.W, q.n_seg) - q.C);
        if (t_cut > q.nt) t_cut = q.nt;
        if (t_cut > q.t_prev) { t0 = q.t_prev; t1 = t_cut; seg = sb - 1; q.t_prev = t_cut; q.emitted++; return true; }
    }
    q.done = true;
    if (q.nt > q.t_prev || q.emitted == 0) { t0 = q.t_prev; t1 = q.nt; seg = q.s1; q.emitted++; return true; }
    return false;
}
// upper bound of the split list's length: every base item once, plus one extra part per segment boundary
GB_HD long long split_items_bound(long long n_base, long long n_seg) { return n_base + n_seg + 8; }

// Placement of a pass whose warps are all resident at once (at most warp_slots items): the hardware puts CTA c on an
// SM of its own choosing (measured on B200: neither c mod n_sm nor balanced -- profiles/r2_placement.txt) and nothing
// rebalances afterwards, so the pass lasts as long as the scheduler with the largest sum of item costs (the model
// cost predicts a scheduler's finishing time to 4 %).  A placed pass launches `rounds` CTAs per SM; every warp looks
// up where it runs -- bin = SM x 4 + scheduler -- and is the k-th warp of its bin to ask: it takes entry (k, bin) of
// the list, which is sorted longest first and dealt to the bins in boustrophedon order, round k left to right for
// even k and right to left for odd k.  No table; host- and device-built lists need nothing but their order.
GB_HD int place_item(int k, int bin, int n_bins) { return (k & 1) ? k * n_bins + (n_bins - 1 - bin) : k * n_bins + bin; }
GB_HD int place_rounds(long long n_items, int n_bins) { return (int)((n_items + n_bins - 1) / n_bins); }

// Tile capacity of a pass: 64 i-particles per warp is the most efficient shape (staging is amortised
// over the most pairs), and measured on 1/4- and 1/8-size shards it stays the fastest even at 0.6 waves.
// Only a pass that cannot give every fourth warp slot an item (per-call functor form, a small boundary
// set) is latency-bound on one item's serial chain: there use 32, or j-split tiles.
// n_items_at[k] = sum over walks of ceil(ni / (64 >> k)), k = 0..4.
GB_HD int tile_cap_choose(const long long n_items_at[5], long long warp_slots, int tile_cap, bool split, int rmax)
{
    int cap = rmax >= 2 ? 64 : 32;
    if (tile_cap > 0) return cap < tile_cap ? cap : tile_cap;
    if (split && rmax <= 2) {
        const long long target = warp_slots / 4;
        for (; cap > 4; cap >>= 1) {
            int k = 0;
            for (int c = 64; c > cap; c >>= 1) k++;
            if (n_items_at[k] >= target) break;
        }
    }
    return cap;
}

}  // namespace gb
