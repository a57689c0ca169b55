// gplum_b200.cu -- host side of libgplum_b200.so: the C ABI declared in include/gplum_b200.h.
//
// Mirrors what FDPS does around the interaction functors (FDPS/src/tree_for_force_impl_force.hpp
// :63-266 multi-walk-index, :1404-1589 calcForce/calcForceOnly) with device-resident buffers:
// j-particles are shipped and packed once per force pass ("send all"), walks arrive as index
// lists, one kernel launch evaluates every walk of a dispatch, forces come back in one copy.
// There is no CPU implementation in this library: without a CUDA device every call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/gplum_b200.h"
#include "kernels.cuh"
#include "items.h"
#include "soft_corr.h"
#include "dev_tree.h"
#include "iso_step.h"

namespace {

using namespace gb;

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(GPLUM_B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                        cudaGetErrorString(e_));                                                   \
    } while (0)

// grow-only device / pinned buffers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) CU(cudaFree(p));
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) CU(cudaFreeHost(p));
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        CU(cudaMallocHost(&p, want));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One set of walks resident on the device (a dispatch, a flat pass, or a per-call functor call)
struct WalkSet {
    DevBuf epi, epi_off, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, items, force;
    DevBuf scratch, arrive;      // partial sums and arrival counters of j-split tiles (items.h)
    DevBuf need;                 // tree_download_compact: particles that occur in captured candidate pairs
    DevBuf org_index;            // tree_download_compact: particle indices of the compact neighbour records
    DevBuf place;                // claim counters of placed passes (kernels.cuh: PassParams::place), zero between launches
    DevBuf force_org;            // forces in the caller's particle order (tree_download_original)
    DevBuf seg_off;              // segments of a one-wave pass: warp s runs items [seg_off[s], seg_off[s+1]) (items.h)
    int n_seg = 0;               // 0: one item per warp
    int part_e0 = 0, part_e1 = 0;  // multi-GPU tree build: this rank's i-particles [e0, e1) of the set (0, 0: all)
    PinBuf h_force, h_stage;     // pinned: results, flattened inputs of dispatch()
    int n_walk = 0, n_items = 0;
    long long n_epi = 0, n_adr_epj = 0, n_adr_spj = 0;
    long long n_int_epep = 0, n_int_epsp = 0;
    std::vector<int> ni_host;                    // for retrieve()
    std::vector<long long> epi_off_host;
    bool pending = false;
    cudaEvent_t done = nullptr;                  // recorded after the D2H of a dispatch: retrieve(tag) waits on it only
    // dispatch() pipeline: the walks of one dispatch are cut into sub-batches; sub-batch b's lists go up on the
    // copy-in stream while b-1 computes and b-2's forces come down on the copy-out stream
    std::vector<int> sub_w0;                     // first walk of each sub-batch (+ n_walk at the end)
    std::vector<cudaEvent_t> ev_in, ev_k, ev_out;
    // changeover correction (soft_corr.cu): candidate capture of the last pass + work/result buffers
    DevBuf self_adr, pairs, corr_meta, cnt, off, cursor, csr, corr_out, corr_init, ngb, scan_temp, corr_compact;
    unsigned int pair_cap = 0;
    bool captured = false, corrected = false, corrected_initial = false;
    bool corr_checked = false;                   // the status words of the last correction have been read and were clean
    void release()
    {
        if (done) { cudaEventDestroy(done); done = nullptr; }
        for (auto *v : {&ev_in, &ev_k, &ev_out}) { for (cudaEvent_t e : *v) cudaEventDestroy(e); v->clear(); }
        for (DevBuf *b : {&epi, &epi_off, &adr_epj, &epj_disp, &n_epj, &adr_spj, &spj_disp, &n_spj, &items, &force, &scratch, &arrive, &place, &org_index, &need, &seg_off, &force_org,
                          &self_adr, &pairs, &corr_meta, &cnt, &off, &cursor, &csr, &corr_out, &corr_init, &ngb, &scan_temp, &corr_compact})
            b->release();
        h_force.release(); h_stage.release();
    }
};

// j-particles of the current force pass
struct JSet {
    DevBuf epj_aos, spj_aos, epj_packed, spj_packed;
    const void *ext_epj = nullptr, *ext_spj = nullptr;   // caller-owned packed arrays (all-gather output)
    const void *spj_src = nullptr;                        // the SPJ records live elsewhere on the device (a GPU-built tree's cell
                                                          // moments, dev_tree.cu) instead of in spj_aos
    bool epj_packed_fresh = false;                        // the list builder's gather wrote epj_packed itself
    int n_epj = 0, n_spj = 0;
    const void *spj_records() const { return spj_src ? spj_src : spj_aos.p; }
    const EpjPacked *epj() const { return ext_epj ? (const EpjPacked *)ext_epj : (const EpjPacked *)epj_packed.p; }
    const SpjPacked *spj() const { return ext_spj ? (const SpjPacked *)ext_spj : (const SpjPacked *)spj_packed.p; }
    void release() { epj_aos.release(); spj_aos.release(); epj_packed.release(); spj_packed.release(); }
};

constexpr int N_TAG = 4;
constexpr int MAX_PEERS = 16;

// Multi-GPU peer mode: every rank owns two slabs (double buffer) of packed EP records; the other
// ranks' slabs are mapped with CUDA IPC, so boundary walks read them straight over NVLink.
struct Peer {
    bool on = false;
    int world = 0, rank = 0, shift = 0, parity = 0;
    int epoch = 0;                                      // packs so far; flags[q] of every rank reach it once q has packed
    void *slab[2] = {nullptr, nullptr};               // this rank's own slabs (cudaMalloc)
    void *mapped[2][MAX_PEERS] = {};                    // every rank's slabs in this process' address space
    DevBuf table[2];                                    // device copies of mapped[b][0..world)
    DevBuf done;                                        // block counter of peer_pack_kernel
    // a pack requested by gplum_b200_peer_pack and not launched yet: the next force launch does it in its prologue if
    // it is a placed pass (kernels.cuh: fp_*), else it launches peer_pack_kernel first
    bool pending = false; const void *pending_epj = nullptr; int pending_n = 0;
    int fuse = 0;                                       // GPLUM_B200_FUSE_PACK=1: pack in the prologue of placed passes; measured 4 % slower
                                                        // than the separate launch at 8 GPUs (profiles/r2_fused_pack.txt), so off by default
};

struct Ctx {
    bool ready = false;
    int device = -1;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // dispatch()/retrieve() pipeline: H2D and D2H engines
    cudaEvent_t ev_compute = nullptr, ev_j[8] = {};        // compute-stream marker; j-set upload chunks
    float eps2 = 0.0f;
    int quad = 1, flags = 0;
    JSet jset;
    Peer peer;
    WalkSet slots[N_TAG];
    std::atomic<long long> launches{0}, n_epep{0}, n_epsp{0};
    std::mutex mu;
    int smem_bytes = 0;
    int rmax = 2;               // i-particles per lane (GPLUM_B200_RMAX = 2 or 4)
    int cur = 0;                // resident walk set used by the walks_* calls
    long long warp_slots = 148 * 24;   // resident warps of the force kernel on this device
    int tile_cap = 0;           // 0 = choose per pass (build_items), else the i-tile capacity to use
    int jsplit = 1;             // short i-tiles split their j-lists over lane groups (GPLUM_B200_JSPLIT=0 disables)
    int bulk = 0;               // GPLUM_B200_BULK=1: EP tiles staged by bulk copies (cp.async.bulk + mbarrier) of index runs;
                                // default per-record cp.async, measured 2 % faster (profiles/r2_bulk_copy_probe.txt)
    int place = 1;              // passes of at most `place` waves of items are placed by SM and scheduler (items.h: place_item;
                                // GPLUM_B200_PLACE=0: item s on warp s of the grid, wherever the hardware puts it)
    int split_m = 2;            // j-split of full-width tiles in passes with less than two waves of items: pieces per warp
                                // slot (items.h); 0 = never (GPLUM_B200_SPLIT_M)
    bool corr_on = false;       // the force pass records candidate pairs for the changeover correction
    long long corr_cap = 0;     // pair-buffer capacity (0 = 4 x n_epi + 2^20)
    DevBuf tree_in, tree_raw;   // GPU list builder: SoA inputs, unsorted EPJGrav
    DevBuf tree_motion;         // tree_set_motion: vel / acc_d columns
    DevBuf tree_inv; bool tree_inv_valid = false;   // particle -> tree-order index of the last GPU build (built on demand)
    PinBuf tree_pin;            // pinned staging of pageable host inputs (upload_host)
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    PinBuf motion_pin;                                     // tree_set_motion_gather: staging of the listed particles' columns
    PinBuf h_small; cudaEvent_t ev_small = nullptr;        // a few pinned words for counts that come back mid-call
    cudaEvent_t ev_cols = nullptr, ev_cols0 = nullptr;     // tree_build_columns: columns on the device / columns free to overwrite
    bool stage_used[2] = {false, false};
    bool tree_built = false;    // the selected slot + j-set hold a GPU-built tree
    bool trace_on = false;      // gplum_b200_debug_trace: the force kernel records where and when every item ran
    DevBuf trace; int trace_items = 0;
    // device-resident particle state (iso_step.cu): EPJGrav[n] with particle k at slot k, + time, dt, acc0, flags
    DevBuf st_epj, st_time, st_dt, st_acc0, st_iso, st_star, st_handled, st_rec, st_idx, st_cnt;
    PinBuf st_pin;
    int st_n = 0;
};
Ctx g;

int ensure_init()
{
    if (g.ready) return 0;
    const char *env = getenv("GPLUM_B200_DEVICE");
    return gplum_b200_init(env ? atoi(env) : 0, 0, 0);
}

// ---- work list: split every walk into i-tiles, choose a tile shape, longest first; a pass with few items also
// cuts its full-width tiles along j (items.h).  n_slots / n_groups receive the scratch slots and arrival counters
// the split tiles of this list need (numbered from slot_base / group_base).  peer_walk (optional, per walk): the
// walk reads other ranks' particles -> its items wait for the peers' flags.
struct ItemList {
    std::vector<WorkItem> items;
    std::vector<int> seg_off;          // empty: one item per warp; else warp s runs items [seg_off[s], seg_off[s+1])
    int n_slots = 0, n_groups = 0;
};
void build_items(int n_walk, const int *ni, const int *n_epj, const int *n_spj, ItemList &out, int walk_base = 0,
                 bool allow_split = true, int slot_base = 0, int group_base = 0, const unsigned char *peer_walk = nullptr)
{
    out.items.clear(); out.seg_off.clear(); out.n_slots = 0; out.n_groups = 0;
    std::vector<std::pair<double, BaseItem>> tmp;
    tmp.reserve((size_t)n_walk * 2);
    const bool split = g.jsplit != 0;
    long long n_items_at[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < n_walk; w++)
        for (int k = 0; k < 5; k++) n_items_at[k] += (ni[w] + (64 >> k) - 1) / (64 >> k);
    int cap = tile_cap_choose(n_items_at, g.warp_slots, g.tile_cap, split, g.rmax);
    // a pass that may be laid out in segments keeps the most efficient tile shape: the segments fill the GPU
    if (allow_split && g.split_m > 0 && g.rmax == 2 && g.tile_cap == 0) cap = 64;
    for (int w = 0; w < n_walk; w++) {
        int rem = ni[w], i0 = 0;
        while (rem > 0) {
            if (g.rmax >= 3 && rem > 64) {                   // RMAX = 4 build: up to 128 i-particles per warp
                const int n = std::min(rem, 128);
                const int cfg = (n + 31) / 32 - 1;
                tmp.push_back({tile_cost(n_epj[w], n_spj[w], 32 * (cfg + 1)), BaseItem{w, i0, n, cfg}});
                rem -= n; i0 += n;
                continue;
            }
            int n, shape;
            tile_next(rem, cap, split, n, shape);
            tmp.push_back({tile_cost(n_epj[w], n_spj[w], shape), BaseItem{w, i0, n, tile_cfg_of(shape)}});
            rem -= n; i0 += n;
        }
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    const bool do_split = allow_split && g.rmax <= 2 && split_active((long long)tmp.size(), g.warp_slots, g.split_m);
    if (!do_split) {
        out.items.reserve(tmp.size());
        for (auto &t : tmp) {
            const BaseItem &b = t.second;
            const int wait = (peer_walk && peer_walk[b.walk]) ? ITEM_PEER_WAIT : 0;
            out.items.push_back(WorkItem{b.walk + walk_base, b.i0, b.ni, b.cfg | wait, 0, -1, 0, 0});
        }
        return;
    }
    // one wave of equal segments (items.h): cut the line of costs at every n_seg-th of its length, at j-tile boundaries
    long long W = 0;
    for (auto &t : tmp) W += item_cost_units(t.first);
    const long long n_seg = g.warp_slots;
    out.items.reserve(tmp.size() + (size_t)n_seg);
    out.seg_off.assign((size_t)n_seg + 1, 0);
    std::vector<int> seg_of_item;
    seg_of_item.reserve(tmp.size() + (size_t)n_seg);
    long long C = 0;
    for (auto &t : tmp) {
        const BaseItem &b = t.second;
        const int w = b.walk;
        const int wait = (peer_walk && peer_walk[w]) ? ITEM_PEER_WAIT : 0;
        const long long c = item_cost_units(t.first);
        struct Part { int t0, t1; long long seg; };
        Part parts[SPLIT_K_MAX + 1];
        int K = 0;
        SegCut q;
        seg_cut_begin(q, W, n_seg, C, c, b.cfg, n_epj[w], n_spj[w]);
        for (int t0, t1; K <= SPLIT_K_MAX; K++) {
            long long sg;
            if (!seg_cut_next(q, t0, t1, sg)) break;
            parts[K] = Part{t0, t1, sg};
        }
        for (int k = 0; k < K; k++) {
            const bool whole = K == 1;
            out.items.push_back(WorkItem{w + walk_base, b.i0, b.ni, b.cfg | wait | (whole ? 0 : (K << 8) | (k << 16)),
                                         whole ? 0 : parts[k].t0, whole ? -1 : parts[k].t1,
                                         whole ? 0 : slot_base + out.n_slots, whole ? 0 : group_base + out.n_groups});
            seg_of_item.push_back((int)parts[k].seg);
        }
        if (K > 1) { out.n_slots += K; out.n_groups += 1; }
        C += c;
    }
    // seg_off[s] = first item of segment s (segments are non-decreasing along the list)
    {
        size_t k = 0;
        for (long long sgm = 0; sgm <= n_seg; sgm++) {
            while (k < seg_of_item.size() && seg_of_item[k] < sgm) k++;
            out.seg_off[(size_t)sgm] = (int)k;
        }
        out.seg_off[(size_t)n_seg] = (int)seg_of_item.size();
    }
}

// scratch records and arrival counters of a walk set's split tiles (grow-only; counters start at zero and every
// pass leaves them at zero)
int reserve_split(WalkSet &ws, int n_slots, int n_groups, cudaStream_t st)
{
    if (int r = ws.scratch.reserve((size_t)std::max(n_slots, 1) * 64 * sizeof(ForceAos))) return r;
    const size_t cap0 = ws.arrive.cap;
    if (int r = ws.arrive.reserve((size_t)std::max(n_groups, 1) * 4)) return r;
    if (ws.arrive.cap != cap0) CU(cudaMemsetAsync(ws.arrive.p, 0, ws.arrive.cap, st));
    return 0;
}

// The pack of a multi-GPU peer step as a launch of its own (kernels.cuh: peer_pack_kernel)
int launch_peer_pack(cudaStream_t st)
{
    Peer &pe = g.peer;
    JSet &j = g.jset;
    const int n = pe.pending_n;
    const int n_spj = (j.n_spj > 0 && !j.ext_spj) ? j.n_spj : 0;
    const int nb_e = std::max(1, (n + PACK_BLOCK - 1) / PACK_BLOCK), nb_s = (n_spj + PACK_BLOCK - 1) / PACK_BLOCK;
    peer_pack_kernel<<<nb_e + nb_s, PACK_BLOCK, 0, st>>>((const EpjAos *)pe.pending_epj, n, (EpjPacked *)pe.slab[pe.parity],
                                                         j.spj_records(), n_spj, (SpjPacked *)j.spj_packed.p, g.quad,
                                                         (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0, g.eps2, nb_e,
                                                         (unsigned int *)pe.done.p, (void *const *)pe.table[0].p,
                                                         ((size_t)1 << pe.shift) * sizeof(EpjPacked), pe.rank, pe.world, pe.epoch);
    CU(cudaGetLastError());
    g.launches++;
    pe.pending = false;
    return 0;
}

// Launches the force kernel on items [item0, item0 + n_items) of the set (default: all).  `first` resets the
// candidate capture and counts the pass's interactions; sub-batch launches of one pass pass first = false.
int launch_pass(WalkSet &ws, cudaStream_t st, float eps2, int item0 = 0, int n_items = -1, bool first = true,
                int seg0 = 0, int n_seg = -1)
{
    if (n_seg < 0) n_seg = ws.n_seg;
    if (n_items < 0) n_items = ws.n_items;
    // the bookkeeping of a pass's first launch happens even when that launch is empty (a dispatch whose first
    // sub-batch holds no i-particles): later sub-batches append to THIS pass's capture, not to the previous one's
    if (first) {
        ws.captured = false; ws.corrected = false; ws.corr_checked = false;
        g.n_epep += ws.n_int_epep; g.n_epsp += ws.n_int_epsp;
        if (g.corr_on && ws.n_epi > 0) {
            const long long cap = g.corr_cap > 0 ? g.corr_cap : 4 * ws.n_epi + (1 << 20);
            if (cap > 0x7fffffffLL) return fail(GPLUM_B200_ERR_ARG, "pair capacity %lld exceeds 2^31", cap);
            if (int r = ws.self_adr.reserve((size_t)ws.n_epi * 4)) return r;
            if (int r = ws.pairs.reserve((size_t)cap * sizeof(int2))) return r;
            if (int r = ws.corr_meta.reserve(16)) return r;
            ws.pair_cap = (unsigned int)cap;
            CU(cudaMemsetAsync(ws.self_adr.p, 0xff, (size_t)ws.n_epi * 4, st));
            CU(cudaMemsetAsync(ws.corr_meta.p, 0, 16, st));
            // i-particles no walk covers keep number = 0 (and trip the "not in its own EP list" status) instead of
            // feeding uninitialised counts into the correction's scan
            CU(cudaMemsetAsync(ws.force.p, 0, (size_t)ws.n_epi * sizeof(ForceAos), st));
            ws.captured = true;
        }
    }
    if (n_items == 0) return 0;
    PassParams p;
    p.epi = (const EpiAos *)ws.epi.p;
    p.epi_off = (const int *)ws.epi_off.p;
    p.adr_epj = (const int *)ws.adr_epj.p; p.epj_disp = (const long long *)ws.epj_disp.p; p.n_epj = (const int *)ws.n_epj.p;
    p.adr_spj = (const int *)ws.adr_spj.p; p.spj_disp = (const long long *)ws.spj_disp.p; p.n_spj = (const int *)ws.n_spj.p;
    p.epj = g.jset.epj(); p.spj = g.jset.spj();
    p.peer_epj = g.peer.on ? (const EpjPacked *const *)g.peer.table[g.peer.parity].p : nullptr;
    p.peer_shift = g.peer.shift;
    p.force = (ForceAos *)ws.force.p;
    p.items = (const WorkItem *)ws.items.p + item0;
    p.eps2 = eps2;
    p.rank_squared = (g.flags & GPLUM_B200_RANK_SQUARED) ? 1 : 0;
    p.self_adr = nullptr; p.pairs = nullptr; p.pair_count = nullptr; p.pair_cap = 0;
    p.scratch = (ForceAos *)ws.scratch.p; p.arrive = (int *)ws.arrive.p;
    p.seg_off = n_seg > 0 ? (const int *)ws.seg_off.p + seg0 : nullptr; p.n_seg = n_seg;
    p.peer_flags = nullptr; p.peer_world = 0; p.peer_epoch = 0;
    if (g.peer.on) {
        p.peer_flags = reinterpret_cast<const int *>(static_cast<const char *>(g.peer.slab[0]) + ((size_t)1 << g.peer.shift) * sizeof(EpjPacked));
        p.peer_world = g.peer.world; p.peer_epoch = g.peer.epoch;
    }
    p.trace = nullptr;
    if (g.trace_on) {
        if (int r = g.trace.reserve((size_t)(item0 + n_items) * 32)) return r;
        if (item0 == 0) CU(cudaMemsetAsync(g.trace.p, 0, (size_t)n_items * 32, st));
        p.trace = (unsigned long long *)g.trace.p + 4 * (size_t)item0;
        g.trace_items = item0 + n_items;
    }
    if (g.corr_on && ws.captured) {
        p.self_adr = (int *)ws.self_adr.p;
        p.pairs = (int2 *)ws.pairs.p;
        p.pair_count = (unsigned int *)ws.corr_meta.p;
        p.pair_cap = ws.pair_cap;
    }
    int n_warps = n_seg > 0 ? n_seg : n_items;
    // a whole pass of at most one wave of items, more than one per scheduler: placed (items.h)
    const int n_bins = (int)(g.warp_slots / 24) * 4;
    p.place = nullptr; p.place_bins = 0; p.place_rounds = 0;
    if (g.place && g.rmax <= 2 && n_seg == 0 && item0 == 0 && n_items == ws.n_items && n_items > n_bins && n_items <= g.warp_slots * g.place) {
        const size_t cap0 = ws.place.cap;
        if (int r = ws.place.reserve((size_t)(n_bins + 3) * 4)) return r;
        if (ws.place.cap != cap0) CU(cudaMemsetAsync(ws.place.p, 0, ws.place.cap, st));
        p.place = (int *)ws.place.p; p.place_bins = n_bins; p.place_rounds = place_rounds(n_items, n_bins);
        n_warps = n_bins * std::min(p.place_rounds, (int)(g.warp_slots / n_bins));      // more rounds than fit: the resident warps loop
    }
    // bulk-copy staging of the EP tiles (kernels.cuh) or per-record cp.async; peer slabs are always gathered by cp.async
    const bool bulk = g.bulk && !g.peer.on;
    const dim3 grid((n_warps + WPB - 1) / WPB), block(WPB * 32);
    // multi-GPU peer mode: the step's pack, fused into a placed pass (cooperative launch: the prologue ends in a grid
    // barrier) or as a launch of its own
    p.fp_slab = nullptr;
    if (g.peer.on && g.peer.pending) {
        Peer &pe = g.peer;
        bool fused = false;
        if (pe.fuse && p.place && g.rmax <= 2 && !bulk) {
            JSet &j = g.jset;
            p.fp_epj_in = (const EpjAos *)pe.pending_epj; p.fp_n_epj = pe.pending_n; p.fp_slab = (EpjPacked *)pe.slab[pe.parity];
            p.fp_n_spj = (j.n_spj > 0 && !j.ext_spj) ? j.n_spj : 0;
            p.fp_spj_in = j.spj_records(); p.fp_spj_out = (SpjPacked *)j.spj_packed.p;
            p.fp_quad = g.quad; p.fp_trace = (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0;
            p.fp_slab0_of = (void *const *)pe.table[0].p; p.fp_flag_off = ((size_t)1 << pe.shift) * sizeof(EpjPacked); p.fp_rank = pe.rank;
            int n_it = n_items;
            void *args[] = {(void *)&p, (void *)&n_it};
            const cudaError_t ce = cudaLaunchCooperativeKernel((const void *)force_pass_kernel<2, false>, grid, block, args, (size_t)g.smem_bytes, st);
            if (ce == cudaSuccess) {
                fused = true;
            } else {                                    // not co-resident here (another context on the GPU): separate launches
                cudaGetLastError();
                p.fp_slab = nullptr;
            }
        }
        pe.pending = false;
        if (fused) { g.launches++; return 0; }
        if (int r = launch_peer_pack(st)) return r;
    }
    if (g.rmax > 2) force_pass_kernel<4, false><<<grid, block, g.smem_bytes, st>>>(p, n_items);
    else if (bulk) force_pass_kernel<2, true><<<grid, block, g.smem_bytes, st>>>(p, n_items);
    else force_pass_kernel<2, false><<<grid, block, g.smem_bytes, st>>>(p, n_items);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

int pack_j(cudaStream_t st, float eps2)
{
    JSet &j = g.jset;
    if (j.n_epj > 0 && !j.ext_epj && !j.epj_packed_fresh) {
        pack_epj_kernel<<<(j.n_epj + 255) / 256, 256, 0, st>>>((const EpjAos *)j.epj_aos.p, j.n_epj, (EpjPacked *)j.epj_packed.p);
        CU(cudaGetLastError());
        g.launches++;
    }
    j.epj_packed_fresh = false;
    if (j.n_spj > 0 && !j.ext_spj) {
        pack_spj_kernel<<<(j.n_spj + 255) / 256, 256, 0, st>>>(j.spj_records(), j.n_spj, (SpjPacked *)j.spj_packed.p, g.quad,
                                                               (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0, eps2);
        CU(cudaGetLastError());
        g.launches++;
    }
    return 0;
}

int upload_j(const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all, cudaStream_t st)
{
    JSet &j = g.jset;
    const size_t ssz = g.quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos);
    j.ext_epj = j.ext_spj = nullptr; j.spj_src = nullptr; j.epj_packed_fresh = false;
    j.n_epj = n_epj_all; j.n_spj = n_spj_all;
    if (int r = j.epj_aos.reserve((size_t)n_epj_all * sizeof(EpjAos))) return r;
    if (int r = j.epj_packed.reserve((size_t)n_epj_all * sizeof(EpjPacked))) return r;
    if (int r = j.spj_aos.reserve((size_t)n_spj_all * ssz)) return r;
    if (int r = j.spj_packed.reserve((size_t)n_spj_all * sizeof(SpjPacked))) return r;
    if (n_epj_all) CU(cudaMemcpyAsync(j.epj_aos.p, epj_all, (size_t)n_epj_all * sizeof(EpjAos), cudaMemcpyHostToDevice, st));
    if (n_spj_all) CU(cudaMemcpyAsync(j.spj_aos.p, spj_all, (size_t)n_spj_all * ssz, cudaMemcpyHostToDevice, st));
    return 0;
}

// Copy flat walk arrays to the device and build the work list.
int upload_walks(WalkSet &ws, int n_walk, const void *epi_all, const int *epi_off, const int *ni,
                 const int *adr_epj, const long long *epj_disp, const int *n_epj,
                 const int *adr_spj, const long long *spj_disp, const int *n_spj, cudaStream_t st)
{
    long long n_epi = 0, n_ae = 0, n_as = 0, i_ee = 0, i_es = 0;
    for (int w = 0; w < n_walk; w++) {
        if (ni[w] < 0 || n_epj[w] < 0 || n_spj[w] < 0) return fail(GPLUM_B200_ERR_ARG, "negative count in walk %d", w);
        n_epi = std::max(n_epi, (long long)epi_off[w] + ni[w]);
        n_ae = std::max(n_ae, epj_disp[w] + n_epj[w]);
        n_as = std::max(n_as, spj_disp[w] + n_spj[w]);
        i_ee += (long long)ni[w] * n_epj[w];
        i_es += (long long)ni[w] * n_spj[w];
    }
    ws.n_walk = n_walk; ws.n_epi = n_epi; ws.n_adr_epj = n_ae; ws.n_adr_spj = n_as;
    ws.n_int_epep = i_ee; ws.n_int_epsp = i_es;
    ws.part_e0 = ws.part_e1 = 0;
    g.tree_built = false;
    // multi-GPU peer mode: a walk whose EP list names a particle of another rank (index = owner << shift | local index)
    // must not start before that rank has packed: its items wait for the peers' flags inside the kernel, and one empty
    // barrier item ends the pass only after every peer has packed (a slab is reused two epochs later; kernels.cuh)
    std::vector<unsigned char> peer_walk;
    if (g.peer.on) {
        peer_walk.assign((size_t)n_walk, 0);
#pragma omp parallel for schedule(dynamic, 16)
        for (int w = 0; w < n_walk; w++) {
            const int *a = adr_epj + epj_disp[w];
            for (int j = 0; j < n_epj[w]; j++)
                if ((a[j] >> g.peer.shift) != g.peer.rank) { peer_walk[w] = 1; break; }
        }
    }
    ItemList il;
    build_items(n_walk, ni, n_epj, n_spj, il, 0, true, 0, 0, g.peer.on ? peer_walk.data() : nullptr);
    if (g.peer.on) {
        il.items.push_back(WorkItem{0, 0, 0, ITEM_PEER_WAIT, 0, 0, 0, 0});
        if (!il.seg_off.empty()) il.seg_off.back() = (int)il.items.size();        // the last segment's warp also holds the barrier
    }
    std::vector<WorkItem> &items = il.items;
    ws.n_items = (int)items.size();
    ws.n_seg = il.seg_off.empty() ? 0 : (int)il.seg_off.size() - 1;
    if (ws.n_seg > 0) {
        if (int r = ws.seg_off.reserve(il.seg_off.size() * 4)) return r;
        CU(cudaMemcpyAsync(ws.seg_off.p, il.seg_off.data(), il.seg_off.size() * 4, cudaMemcpyHostToDevice, st));   // pageable: staged on return
    }
    if (int r = reserve_split(ws, il.n_slots, il.n_groups, st)) return r;
    if (int r = ws.epi.reserve((size_t)n_epi * sizeof(EpiAos))) return r;
    if (int r = ws.force.reserve((size_t)n_epi * sizeof(ForceAos))) return r;
    if (int r = ws.epi_off.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.n_epj.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.n_spj.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.epj_disp.reserve((size_t)n_walk * 8)) return r;
    if (int r = ws.spj_disp.reserve((size_t)n_walk * 8)) return r;
    if (int r = ws.adr_epj.reserve((size_t)n_ae * 4)) return r;
    if (int r = ws.adr_spj.reserve((size_t)n_as * 4)) return r;
    if (int r = ws.items.reserve(items.size() * sizeof(WorkItem))) return r;
    if (n_walk == 0) return 0;
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
    if (n_epi) CU(cudaMemcpyAsync(ws.epi.p, epi_all, (size_t)n_epi * sizeof(EpiAos), H2D, st));
    CU(cudaMemcpyAsync(ws.epi_off.p, epi_off, (size_t)n_walk * 4, H2D, st));
    CU(cudaMemcpyAsync(ws.n_epj.p, n_epj, (size_t)n_walk * 4, H2D, st));
    CU(cudaMemcpyAsync(ws.n_spj.p, n_spj, (size_t)n_walk * 4, H2D, st));
    CU(cudaMemcpyAsync(ws.epj_disp.p, epj_disp, (size_t)n_walk * 8, H2D, st));
    CU(cudaMemcpyAsync(ws.spj_disp.p, spj_disp, (size_t)n_walk * 8, H2D, st));
    if (n_ae) CU(cudaMemcpyAsync(ws.adr_epj.p, adr_epj, (size_t)n_ae * 4, H2D, st));
    if (n_as) CU(cudaMemcpyAsync(ws.adr_spj.p, adr_spj, (size_t)n_as * 4, H2D, st));
    if (!items.empty()) {
        // items lives on the host stack frame: the copy below is from pageable memory and therefore
        // complete (staged) when cudaMemcpyAsync returns.
        CU(cudaMemcpyAsync(ws.items.p, items.data(), items.size() * sizeof(WorkItem), H2D, st));
    }
    return 0;
}

inline void accumulate_force(ForceAos *dst, const ForceAos *src, long long n, bool overwrite)
{
    if (overwrite) { memcpy(dst, src, (size_t)n * sizeof(ForceAos)); return; }
    for (long long i = 0; i < n; i++) {   // "+=" / max / min: src/gravity_kernel.hpp:115-120
        dst[i].acc[0] += src[i].acc[0]; dst[i].acc[1] += src[i].acc[1]; dst[i].acc[2] += src[i].acc[2];
        dst[i].phi += src[i].phi;
        dst[i].number += src[i].number; dst[i].rank += src[i].rank;
        dst[i].id_max = std::max(dst[i].id_max, src[i].id_max);
        dst[i].id_min = std::min(dst[i].id_min, src[i].id_min);
    }
}

// per-thread state of the per-call functor form
struct CallSlot {
    cudaStream_t st = nullptr;
    DevBuf epi, jaos, jpacked, iota, force, meta;
    PinBuf h_force;
    int iota_n = 0;
    ~CallSlot() {}
};
thread_local CallSlot t_slot;

// ---- peer-mode barrier without a collective: every rank owns a small flag array behind its first slab;
// after packing epoch e a rank stores e into ITS entry of EVERY rank's array (over NVLink for the others),
// and a rank's boundary walks start once all entries of its own array have reached e ----
__global__ void peer_signal_kernel(void *const *slab0_of, size_t flag_off, int rank, int world, int epoch)
{
    const int q = threadIdx.x;
    if (q >= world) return;
    int *f = reinterpret_cast<int *>(static_cast<char *>(slab0_of[q]) + flag_off) + rank;
    // the pack kernel before this one (same stream) has completed: its records are in this GPU's L2/HBM
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__global__ void peer_wait_kernel(const int *flags, int world, int epoch)
{
    const int q = threadIdx.x;
    if (q >= world) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (;;) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
        if (v - epoch >= 0) break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 > 10000000000ull) __trap();           // 10 s: a peer died; fail loudly instead of hanging
    }
}

// out[idx[k]] = in[k] for 32 B ForceGrav records (two 16 B halves per record)
__global__ void unsort_force_kernel(int n, const uint4 *__restrict__ in, const int *__restrict__ idx, uint4 *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    out[2 * (size_t)idx[t >> 1] + (t & 1)] = in[t];
}

// Particle order, compact: accphi[i] for every particle; the neighbour words only of particles with candidates
// (all others hold ForceGrav::clear()'s values, src/particle.h:81-85).  Warp-aggregated append.
// need[k] = 1 for every particle (tree order) that occurs in a captured candidate pair, as i or as j
__global__ void mark_pairs_kernel(const int2 *__restrict__ pairs, const unsigned int *__restrict__ pair_count, unsigned int cap,
                                  unsigned char *__restrict__ need)
{
    const unsigned int np = min(*pair_count, cap);
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < np; t += gridDim.x * blockDim.x) {
        const int2 p = pairs[t];
        need[p.x] = 1; need[p.y] = 1;
    }
}
__global__ void unsort_compact_kernel(int n, const uint4 *__restrict__ in, const int *__restrict__ idx, uint4 *__restrict__ accphi,
                                      int *__restrict__ count, int *__restrict__ nb_index, uint4 *__restrict__ nb, int cap,
                                      const unsigned char *__restrict__ need)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    int i = 0;
    uint4 b = make_uint4(0, 0, 0, 0);
    if (k < n) {
        i = idx[k];
        accphi[i] = in[2 * (size_t)k];
        b = in[2 * (size_t)k + 1];
        has = (int)b.x > 0 || (need && need[k]);
    }
    const unsigned int m = __ballot_sync(0xffffffffu, has);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (has) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < cap) { nb_index[slot] = i; nb[slot] = b; }
    }
}

// velocities / direct accelerations (columns in particle order) into the tree-order EPJGrav records
__global__ void invert_order_kernel(int n, const int *__restrict__ idx, int *__restrict__ inv)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) inv[idx[k]] = k;
}
// the same for m listed particles: index[t] = particle, columns [m][3]
__global__ void set_motion_sparse_kernel(int m, int n, const int *__restrict__ index, const int *__restrict__ inv, const double *__restrict__ vel,
                                         const double *__restrict__ acc_d, const long long *__restrict__ id, EpjAos *__restrict__ epj)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    if ((unsigned int)index[t] >= (unsigned int)n) return;          // not a particle of this tree: ignored
    const int k = inv[index[t]];
    for (int d = 0; d < 3; d++) {
        epj[k].vel[d] = vel ? vel[3 * (size_t)t + d] : 0.0;
        epj[k].acc_d[d] = acc_d ? acc_d[3 * (size_t)t + d] : 0.0;
    }
    if (id) epj[k].id = id[t];
}
__global__ void set_motion_kernel(int n, const int *__restrict__ idx, const double *__restrict__ vel, const double *__restrict__ acc_d,
                                  EpjAos *__restrict__ epj)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t s = 3 * (size_t)idx[k];
    for (int d = 0; d < 3; d++) {
        epj[k].vel[d] = vel ? vel[s + d] : 0.0;
        epj[k].acc_d[d] = acc_d ? acc_d[s + d] : 0.0;
    }
}

__global__ void iota_kernel(int *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// One functor call: which = 0 EP-EP, 1 EP-SP.
int single_call(int which, const void *epi, int ni, const void *jp, int nj, void *force, float eps2, int quad)
{
    if (int r = ensure_init()) return r;
    if (ni < 0 || nj < 0) return fail(GPLUM_B200_ERR_ARG, "negative ni/nj");
    if (ni == 0) return 0;
    if (!epi || !force || (nj > 0 && !jp)) return fail(GPLUM_B200_ERR_ARG, "null pointer");
    CU(cudaSetDevice(g.device));
    CallSlot &s = t_slot;
    if (!s.st) CU(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    const size_t jsz = which == 0 ? sizeof(EpjAos) : (quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos));
    const size_t psz = which == 0 ? sizeof(EpjPacked) : sizeof(SpjPacked);
    if (int r = s.epi.reserve((size_t)ni * sizeof(EpiAos))) return r;
    if (int r = s.force.reserve((size_t)ni * sizeof(ForceAos))) return r;
    if (int r = s.h_force.reserve((size_t)ni * sizeof(ForceAos))) return r;
    if (int r = s.jaos.reserve((size_t)std::max(nj, 1) * jsz)) return r;
    if (int r = s.jpacked.reserve((size_t)std::max(nj, 1) * psz)) return r;
    if (nj > s.iota_n) {
        if (int r = s.iota.reserve((size_t)nj * 4)) return r;
        s.iota_n = (int)(s.iota.cap / 4);
        iota_kernel<<<(s.iota_n + 255) / 256, 256, 0, s.st>>>((int *)s.iota.p, s.iota_n);
        CU(cudaGetLastError());
    }
    // meta: [epi_off(int) | n_epj | n_spj | pad | epj_disp(ll) | spj_disp(ll) | items...]
    ItemList il;
    const int zero = 0;
    const int ne = which == 0 ? nj : 0, ns = which == 1 ? nj : 0;
    build_items(1, &ni, &ne, &ns, il, 0, false);
    std::vector<WorkItem> &items = il.items;
    struct Meta { int epi_off, n_epj, n_spj, pad; long long epj_disp, spj_disp; } meta = {zero, ne, ns, 0, 0, 0};
    const size_t meta_bytes = sizeof(Meta) + items.size() * sizeof(WorkItem);
    if (int r = s.meta.reserve(meta_bytes)) return r;
    std::vector<unsigned char> hm(meta_bytes);
    memcpy(hm.data(), &meta, sizeof(Meta));
    memcpy(hm.data() + sizeof(Meta), items.data(), items.size() * sizeof(WorkItem));
    CU(cudaMemcpyAsync(s.meta.p, hm.data(), meta_bytes, cudaMemcpyHostToDevice, s.st));
    CU(cudaMemcpyAsync(s.epi.p, epi, (size_t)ni * sizeof(EpiAos), cudaMemcpyHostToDevice, s.st));
    if (nj) CU(cudaMemcpyAsync(s.jaos.p, jp, (size_t)nj * jsz, cudaMemcpyHostToDevice, s.st));
    if (nj && which == 0) {
        pack_epj_kernel<<<(nj + 255) / 256, 256, 0, s.st>>>((const EpjAos *)s.jaos.p, nj, (EpjPacked *)s.jpacked.p);
        g.launches++;
    } else if (nj) {
        pack_spj_kernel<<<(nj + 255) / 256, 256, 0, s.st>>>(s.jaos.p, nj, (SpjPacked *)s.jpacked.p, quad,
                                                           (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0, eps2);
        g.launches++;
    }
    CU(cudaGetLastError());
    PassParams p;
    const unsigned char *dm = (const unsigned char *)s.meta.p;
    p.epi = (const EpiAos *)s.epi.p;
    p.epi_off = (const int *)(dm + offsetof(Meta, epi_off));
    p.n_epj = (const int *)(dm + offsetof(Meta, n_epj));
    p.n_spj = (const int *)(dm + offsetof(Meta, n_spj));
    p.epj_disp = (const long long *)(dm + offsetof(Meta, epj_disp));
    p.spj_disp = (const long long *)(dm + offsetof(Meta, spj_disp));
    p.adr_epj = (const int *)s.iota.p; p.adr_spj = (const int *)s.iota.p;
    p.epj = (const EpjPacked *)s.jpacked.p; p.spj = (const SpjPacked *)s.jpacked.p;
    p.peer_epj = nullptr; p.peer_shift = 0;
    p.force = (ForceAos *)s.force.p;
    p.items = (const WorkItem *)(dm + sizeof(Meta));
    p.eps2 = eps2;
    p.rank_squared = (g.flags & GPLUM_B200_RANK_SQUARED) ? 1 : 0;
    p.self_adr = nullptr; p.pairs = nullptr; p.pair_count = nullptr; p.pair_cap = 0;
    p.scratch = nullptr; p.arrive = nullptr; p.peer_flags = nullptr; p.peer_world = 0; p.peer_epoch = 0;
    p.seg_off = nullptr; p.n_seg = 0; p.trace = nullptr; p.place = nullptr; p.place_bins = 0; p.place_rounds = 0; p.fp_slab = nullptr;
    if (g.rmax > 2) force_pass_kernel<4, false><<<((int)items.size() + WPB - 1) / WPB, WPB * 32, g.smem_bytes, s.st>>>(p, (int)items.size());
    else if (g.bulk) force_pass_kernel<2, true><<<((int)items.size() + WPB - 1) / WPB, WPB * 32, g.smem_bytes, s.st>>>(p, (int)items.size());
    else force_pass_kernel<2, false><<<((int)items.size() + WPB - 1) / WPB, WPB * 32, g.smem_bytes, s.st>>>(p, (int)items.size());
    CU(cudaGetLastError());
    g.launches++;
    if (which == 0) g.n_epep += (long long)ni * nj; else g.n_epsp += (long long)ni * nj;
    CU(cudaMemcpyAsync(s.h_force.p, s.force.p, (size_t)ni * sizeof(ForceAos), cudaMemcpyDeviceToHost, s.st));
    CU(cudaStreamSynchronize(s.st));
    accumulate_force((ForceAos *)force, (const ForceAos *)s.h_force.p, ni, (g.flags & GPLUM_B200_NO_ACCUMULATE) != 0);
    return 0;
}

}  // namespace

// ============================================================================================
extern "C" {

int gplum_b200_abi_version(void) { return GPLUM_B200_ABI_VERSION; }
const char *gplum_b200_last_error(void) { return g_err; }

int gplum_b200_init(int device, size_t max_i, size_t max_j)
{
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.ready && g.device == device) return 0;
    if (g.ready) return fail(GPLUM_B200_ERR_STATE, "already initialised on device %d", g.device);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(GPLUM_B200_ERR_NO_DEVICE, "no CUDA device (%s); libgplum_b200 has no CPU path",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(GPLUM_B200_ERR_ARG, "device %d out of range (%d devices)", device, n_dev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GPLUM_B200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CU(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    if (const char *e = getenv("GPLUM_B200_RMAX")) g.rmax = atoi(e) >= 3 ? 4 : (atoi(e) <= 1 ? 1 : 2);
    if (const char *e = getenv("GPLUM_B200_FLAGS")) g.flags = atoi(e);
    if (const char *e = getenv("GPLUM_B200_JSPLIT")) g.jsplit = atoi(e);
    if (const char *e = getenv("GPLUM_B200_SPLIT_M")) g.split_m = std::max(0, atoi(e));
    if (const char *e = getenv("GPLUM_B200_PLACE")) g.place = atoi(e);
    if (const char *e = getenv("GPLUM_B200_FUSE_PACK")) g.peer.fuse = atoi(e);
    if (const char *e = getenv("GPLUM_B200_BULK")) g.bulk = atoi(e) ? 1 : 0;
    g.smem_bytes = (int)(g.rmax <= 2 ? sizeof(WarpSmem<64>) : sizeof(WarpSmem<128>)) * WPB;
    CU(cudaFuncSetAttribute(force_pass_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem<64>) * WPB));
    CU(cudaFuncSetAttribute(force_pass_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem<64>) * WPB));
    CU(cudaFuncSetAttribute(force_pass_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem<128>) * WPB));
    g.warp_slots = (long long)prop.multiProcessorCount * 24;      // 6 CTAs x 4 warps per SM (80 registers)
    g.device = device;
    if (max_i) {
        if (int r = g.slots[0].epi.reserve(max_i * sizeof(EpiAos))) return r;
        if (int r = g.slots[0].force.reserve(max_i * sizeof(ForceAos))) return r;
    }
    if (max_j) {
        if (int r = g.slots[0].adr_epj.reserve(max_j * 4)) return r;
    }
    g.ready = true;
    return 0;
}

int gplum_b200_finalize(void)
{
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.ready) return 0;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    gplum_b200_peer_close();
    gplum_b200_peer_free();
    g.jset.release();
    g.tree_in.release(); g.tree_raw.release(); g.tree_pin.release(); g.h_small.release(); g.motion_pin.release(); g.tree_motion.release(); g.tree_inv.release(); g.tree_inv_valid = false;
    for (DevBuf *b : {&g.st_epj, &g.st_time, &g.st_dt, &g.st_acc0, &g.st_iso, &g.st_star, &g.st_handled, &g.st_rec, &g.st_idx, &g.st_cnt}) b->release();
    g.st_pin.release(); g.st_n = 0;
    gbt::tree_release();
    g.tree_built = false;
    for (auto &s : g.slots) s.release();
    if (g.own_stream) cudaStreamDestroy(g.own_stream);
    if (g.copy_in) { cudaStreamDestroy(g.copy_in); g.copy_in = nullptr; }
    if (g.copy_out) { cudaStreamDestroy(g.copy_out); g.copy_out = nullptr; }
    if (g.ev_compute) { cudaEventDestroy(g.ev_compute); g.ev_compute = nullptr; }
    for (auto &e : g.ev_j) if (e) { cudaEventDestroy(e); e = nullptr; }
    g.own_stream = g.stream = nullptr;
    g.ready = false;
    g.device = -1;
    return 0;
}

int gplum_b200_set_params(float eps2, int quad, int flags)
{
    g.eps2 = eps2; g.quad = quad ? 1 : 0;
    if (flags >= 0) g.flags = flags;      // flags < 0: keep the current ones (e.g. GPLUM_B200_FLAGS)
    return 0;
}

int gplum_b200_set_tile_cap(int cap)
{
    if (cap != 0 && cap != 64 && cap != 32 && cap != 16 && cap != 8 && cap != 4)
        return fail(GPLUM_B200_ERR_ARG, "tile cap %d not in {0, 64, 32, 16, 8, 4}", cap);
    g.tile_cap = cap;
    return 0;
}

int gplum_b200_set_stream(void *cuda_stream)
{
    if (int r = ensure_init()) return r;
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own_stream;
    return 0;
}

int gplum_b200_synchronize(void)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    CU(cudaStreamSynchronize(g.stream));
    if (g.copy_in) CU(cudaStreamSynchronize(g.copy_in));
    if (g.copy_out) CU(cudaStreamSynchronize(g.copy_out));
    return 0;
}

void gplum_b200_counters(long long *kernel_launches, long long *n_epep, long long *n_epsp, int reset)
{
    if (kernel_launches) *kernel_launches = g.launches.load();
    if (n_epep) *n_epep = g.n_epep.load();
    if (n_epsp) *n_epsp = g.n_epsp.load();
    if (reset) { g.launches = 0; g.n_epep = 0; g.n_epsp = 0; }
}

void gplum_b200_packed_sizes(int *epj_packed_bytes, int *spj_packed_bytes)
{
    if (epj_packed_bytes) *epj_packed_bytes = (int)sizeof(EpjPacked);
    if (spj_packed_bytes) *spj_packed_bytes = (int)sizeof(SpjPacked);
}

// ---- per-call functor form ----
int gplum_b200_epep(const void *epi, int ni, const void *epj, int nj, void *force, float eps2)
{
    return single_call(0, epi, ni, epj, nj, force, eps2, 1);
}

int gplum_b200_epsp(const void *epi, int ni, const void *spj, int ns, void *force, float eps2, int quad)
{
    return single_call(1, epi, ni, spj, ns, force, eps2, quad);
}

// ---- flat form ----
int gplum_b200_calc_walks(int n_walk, const void *epi_all, const int *epi_off, const int *ni,
                          const int *adr_epj, const long long *epj_disp, const int *n_epj,
                          const int *adr_spj, const long long *spj_disp, const int *n_spj,
                          const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all,
                          void *force_all, int clear)
{
    if (int r = ensure_init()) return r;
    if (n_walk < 0) return fail(GPLUM_B200_ERR_ARG, "n_walk < 0");
    if (n_walk == 0) return 0;
    CU(cudaSetDevice(g.device));
    WalkSet &ws = g.slots[0];
    cudaStream_t st = g.stream;
    if (int r = upload_j(epj_all, n_epj_all, spj_all, n_spj_all, st)) return r;
    if (int r = pack_j(st, g.eps2)) return r;
    if (int r = upload_walks(ws, n_walk, epi_all, epi_off, ni, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, st)) return r;
    if (int r = launch_pass(ws, st, g.eps2)) return r;
    if (ws.n_epi == 0) { CU(cudaStreamSynchronize(st)); return 0; }
    if (clear) {
        // walks tile [0, n_epi) in FDPS; entries not covered by any walk keep the caller's values
        if (int r = ws.h_force.reserve((size_t)ws.n_epi * sizeof(ForceAos))) return r;
        CU(cudaMemcpyAsync(ws.h_force.p, ws.force.p, (size_t)ws.n_epi * sizeof(ForceAos), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int w = 0; w < n_walk; w++)
            memcpy((ForceAos *)force_all + epi_off[w], (const ForceAos *)ws.h_force.p + epi_off[w], (size_t)ni[w] * sizeof(ForceAos));
    } else {
        if (int r = ws.h_force.reserve((size_t)ws.n_epi * sizeof(ForceAos))) return r;
        CU(cudaMemcpyAsync(ws.h_force.p, ws.force.p, (size_t)ws.n_epi * sizeof(ForceAos), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int w = 0; w < n_walk; w++)
            accumulate_force((ForceAos *)force_all + epi_off[w], (const ForceAos *)ws.h_force.p + epi_off[w], ni[w], false);
    }
    return 0;
}

// ---- FDPS multi-walk-index form ----
// dispatch()/retrieve() run as a three-stage pipeline over PCIe: j-particles and walk sub-batches go up on a
// copy-in stream, each sub-batch's force kernel runs on the compute stream as soon as its lists have landed,
// its forces come down on a copy-out stream, and retrieve() hands sub-batches back as they arrive.
namespace {
int ensure_pipe()
{
    if (!g.copy_in) CU(cudaStreamCreateWithFlags(&g.copy_in, cudaStreamNonBlocking));
    if (!g.copy_out) CU(cudaStreamCreateWithFlags(&g.copy_out, cudaStreamNonBlocking));
    if (!g.ev_compute) CU(cudaEventCreateWithFlags(&g.ev_compute, cudaEventDisableTiming));
    for (auto &e : g.ev_j) if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}

// "send all": upload the j-particles in chunks on the copy-in stream; each chunk is packed on the compute
// stream as soon as it has landed.
int send_all_pipelined(const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all)
{
    if (int r = ensure_pipe()) return r;
    JSet &j = g.jset;
    cudaStream_t st = g.stream, ci = g.copy_in;
    const size_t ssz = g.quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos);
    j.ext_epj = j.ext_spj = nullptr; j.spj_src = nullptr; j.epj_packed_fresh = false;
    j.n_epj = n_epj_all; j.n_spj = n_spj_all;
    if (int r = j.epj_aos.reserve((size_t)n_epj_all * sizeof(EpjAos))) return r;
    if (int r = j.epj_packed.reserve((size_t)n_epj_all * sizeof(EpjPacked))) return r;
    if (int r = j.spj_aos.reserve((size_t)n_spj_all * ssz)) return r;
    if (int r = j.spj_packed.reserve((size_t)n_spj_all * sizeof(SpjPacked))) return r;
    // the j arrays may still be read by kernels already queued on the compute stream
    CU(cudaEventRecord(g.ev_compute, st));
    CU(cudaStreamWaitEvent(ci, g.ev_compute, 0));
    constexpr int NCH = 4;                       // chunks per array: g.ev_j[0..3] EP, [4..7] SP
    const int trace = (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0;
    for (int c = 0; c < NCH; c++) {
        // chunk boundaries on multiples of the pack kernels' block (256 records)
        const long long per = ((long long)(n_epj_all + NCH - 1) / NCH + PACK_BLOCK - 1) / PACK_BLOCK * PACK_BLOCK;
        const long long i0 = std::min<long long>(n_epj_all, c * per), i1 = std::min<long long>(n_epj_all, i0 + per);
        if (i1 <= i0) continue;
        CU(cudaMemcpyAsync((EpjAos *)j.epj_aos.p + i0, (const EpjAos *)epj_all + i0, (size_t)(i1 - i0) * sizeof(EpjAos), cudaMemcpyHostToDevice, ci));
        CU(cudaEventRecord(g.ev_j[c], ci));
        CU(cudaStreamWaitEvent(st, g.ev_j[c], 0));
        const int n = (int)(i1 - i0);
        pack_epj_kernel<<<(n + PACK_BLOCK - 1) / PACK_BLOCK, PACK_BLOCK, 0, st>>>((const EpjAos *)j.epj_aos.p + i0, n, (EpjPacked *)j.epj_packed.p + i0);
        CU(cudaGetLastError());
        g.launches++;
    }
    for (int c = 0; c < NCH; c++) {
        const long long per = ((long long)(n_spj_all + NCH - 1) / NCH + PACK_BLOCK - 1) / PACK_BLOCK * PACK_BLOCK;
        const long long i0 = std::min<long long>(n_spj_all, c * per), i1 = std::min<long long>(n_spj_all, i0 + per);
        if (i1 <= i0) continue;
        CU(cudaMemcpyAsync((char *)j.spj_aos.p + i0 * ssz, (const char *)spj_all + i0 * ssz, (size_t)(i1 - i0) * ssz, cudaMemcpyHostToDevice, ci));
        CU(cudaEventRecord(g.ev_j[4 + c], ci));
        CU(cudaStreamWaitEvent(st, g.ev_j[4 + c], 0));
        const int n = (int)(i1 - i0);
        pack_spj_kernel<<<(n + PACK_BLOCK - 1) / PACK_BLOCK, PACK_BLOCK, 0, st>>>((const char *)j.spj_aos.p + i0 * ssz, n, (SpjPacked *)j.spj_packed.p + i0,
                                                                                  g.quad, trace, g.eps2);
        CU(cudaGetLastError());
        g.launches++;
    }
    return 0;
}
}  // namespace

int gplum_b200_dispatch(int tag, int n_walk, const void *const *epi, const int *ni,
                        const int *const *adr_epj, const int *n_epj,
                        const int *const *adr_spj, const int *n_spj,
                        const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all,
                        int send_all)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    if (send_all) return send_all_pipelined(epj_all, n_epj_all, spj_all, n_spj_all);
    if (int r = ensure_pipe()) return r;
    cudaStream_t st = g.stream, ci = g.copy_in, co = g.copy_out;
    if (tag < 0 || tag >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "tag %d out of range [0,%d)", tag, N_TAG);
    if (n_walk < 0) return fail(GPLUM_B200_ERR_ARG, "n_walk < 0");
    WalkSet &ws = g.slots[tag];
    if (ws.pending) return fail(GPLUM_B200_ERR_STATE, "dispatch(tag=%d) while a previous dispatch is not retrieved", tag);
    // flatten the per-walk pointers into one pinned staging block
    long long n_epi = 0, n_ae = 0, n_as = 0, i_ee = 0, i_es = 0, n_it_max = 0;
    for (int w = 0; w < n_walk; w++) {
        if (ni[w] < 0 || n_epj[w] < 0 || n_spj[w] < 0) return fail(GPLUM_B200_ERR_ARG, "negative count in walk %d", w);
        n_epi += ni[w]; n_ae += n_epj[w]; n_as += n_spj[w];
        i_ee += (long long)ni[w] * n_epj[w]; i_es += (long long)ni[w] * n_spj[w];
        n_it_max += (long long)SPLIT_K_MAX * ((ni[w] + 31) / 32) + (ni[w] + 3) / 4 + 1;    // j-split parts of full-width tiles + short tails
    }
    const size_t b_epi = ((size_t)n_epi * sizeof(EpiAos) + 15) & ~(size_t)15, b_ae = (size_t)n_ae * 4, b_as = (size_t)n_as * 4;
    const size_t b_meta = (size_t)n_walk * (4 * 3 + 8 * 2);
    const size_t b_lists = (b_meta + b_ae + b_as + 15) & ~(size_t)15;
    if (int r = ws.h_stage.reserve(b_epi + b_lists + (size_t)n_it_max * sizeof(WorkItem) + 64)) return r;
    unsigned char *h = (unsigned char *)ws.h_stage.p;
    EpiAos *h_epi = (EpiAos *)h;
    long long *h_edisp = (long long *)(h + b_epi);
    long long *h_sdisp = h_edisp + n_walk;
    int *h_off = (int *)(h_sdisp + n_walk);
    int *h_ne = h_off + n_walk, *h_ns = h_ne + n_walk;
    int *h_ae = h_ns + n_walk, *h_as = h_ae + n_ae;
    WorkItem *h_items = (WorkItem *)(h + b_epi + b_lists);
    ws.ni_host.assign(ni, ni + n_walk);
    ws.epi_off_host.resize(n_walk);
    long long oi = 0, oe = 0, os = 0;
    for (int w = 0; w < n_walk; w++) {
        h_off[w] = (int)oi; h_edisp[w] = oe; h_sdisp[w] = os; h_ne[w] = n_epj[w]; h_ns[w] = n_spj[w];
        ws.epi_off_host[w] = oi;
        oi += ni[w]; oe += n_epj[w]; os += n_spj[w];
    }
    ws.n_walk = n_walk; ws.n_epi = n_epi; ws.n_adr_epj = n_ae; ws.n_adr_spj = n_as;
    ws.n_int_epep = i_ee; ws.n_int_epsp = i_es; ws.n_items = 0;
    ws.part_e0 = ws.part_e1 = 0;
    g.tree_built = false;
    if (int r = ws.epi.reserve((size_t)n_epi * sizeof(EpiAos))) return r;
    if (int r = ws.force.reserve((size_t)n_epi * sizeof(ForceAos))) return r;
    if (int r = ws.epi_off.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.n_epj.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.n_spj.reserve((size_t)n_walk * 4)) return r;
    if (int r = ws.epj_disp.reserve((size_t)n_walk * 8)) return r;
    if (int r = ws.spj_disp.reserve((size_t)n_walk * 8)) return r;
    if (int r = ws.adr_epj.reserve((size_t)n_ae * 4)) return r;
    if (int r = ws.adr_spj.reserve((size_t)n_as * 4)) return r;
    if (int r = ws.items.reserve((size_t)n_it_max * sizeof(WorkItem))) return r;
    if (int r = ws.h_force.reserve((size_t)std::max<long long>(n_epi, 1) * sizeof(ForceAos))) return r;
    // sub-batches of about equal PCIe volume; small dispatches stay whole
    const size_t bytes = (size_t)n_epi * (sizeof(EpiAos) + sizeof(ForceAos)) + b_ae + b_as;
    const int n_sub = n_walk == 0 ? 0 : (int)std::max<size_t>(1, std::min<size_t>({(size_t)8, bytes / (4u << 20), (size_t)n_walk}));
    ws.sub_w0.assign(1, 0);
    {
        size_t acc = 0; int b = 1;
        for (int w = 0; w < n_walk && b < n_sub; w++) {
            acc += (size_t)ni[w] * (sizeof(EpiAos) + sizeof(ForceAos)) + 4 * ((size_t)n_epj[w] + n_spj[w]);
            if (acc * n_sub >= bytes * b) { ws.sub_w0.push_back(w + 1); b++; }
        }
        if (n_walk > 0) { if (ws.sub_w0.back() != n_walk) ws.sub_w0.push_back(n_walk); }
    }
    const int nsb = (int)ws.sub_w0.size() - 1;
    for (auto *v : {&ws.ev_in, &ws.ev_k, &ws.ev_out})
        while ((int)v->size() < nsb) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); v->push_back(e); }
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
    // buffers of this tag may still be read by a kernel queued earlier on the compute stream (flat / resident forms)
    CU(cudaEventRecord(g.ev_compute, st));
    CU(cudaStreamWaitEvent(ci, g.ev_compute, 0));
    if (n_walk > 0) {
        CU(cudaMemcpyAsync(ws.epi_off.p, h_off, (size_t)n_walk * 4, H2D, ci));
        CU(cudaMemcpyAsync(ws.n_epj.p, h_ne, (size_t)n_walk * 4, H2D, ci));
        CU(cudaMemcpyAsync(ws.n_spj.p, h_ns, (size_t)n_walk * 4, H2D, ci));
        CU(cudaMemcpyAsync(ws.epj_disp.p, h_edisp, (size_t)n_walk * 8, H2D, ci));
        CU(cudaMemcpyAsync(ws.spj_disp.p, h_sdisp, (size_t)n_walk * 8, H2D, ci));
    }
    ItemList il;
    std::vector<WorkItem> &items = il.items;
    int item0 = 0, slot0 = 0, group0 = 0, seg0 = 0;
    ws.n_seg = 0;
    if (int r = ws.seg_off.reserve((size_t)std::max(nsb, 1) * (size_t)(g.warp_slots + 1) * 4)) return r;
    {   // scratch for the j-split tiles of small sub-batches: at most SPLIT_K_MAX parts per 32 i-particles
        long long tiles = 0;
        for (int w = 0; w < n_walk; w++) tiles += (ni[w] + 31) / 32;
        const bool may_split = g.split_m > 0 && tiles < 2 * g.warp_slots * (long long)std::max(nsb, 1);
        if (int r = reserve_split(ws, may_split ? (int)std::min<long long>(tiles * SPLIT_K_MAX, 1 << 22) : 1, may_split ? (int)tiles : 1, ci)) return r;
    }
    for (int b = 0; b < nsb; b++) {
        const int w0 = ws.sub_w0[b], w1 = ws.sub_w0[b + 1];
#pragma omp parallel for schedule(static)
        for (int w = w0; w < w1; w++) {
            memcpy(h_epi + h_off[w], epi[w], (size_t)ni[w] * sizeof(EpiAos));
            memcpy(h_ae + h_edisp[w], adr_epj[w], (size_t)n_epj[w] * 4);
            memcpy(h_as + h_sdisp[w], adr_spj[w], (size_t)n_spj[w] * 4);
        }
        build_items(w1 - w0, ni + w0, n_epj + w0, n_spj + w0, il, w0, true, slot0, group0);
        slot0 += il.n_slots; group0 += il.n_groups;
        if ((size_t)slot0 * 64 * sizeof(ForceAos) > ws.scratch.cap || (size_t)group0 * 4 > ws.arrive.cap)
            return fail(GPLUM_B200_ERR_STATE, "split-tile scratch estimate too small");
        const int n_it = (int)items.size();
        if (item0 + n_it > n_it_max) return fail(GPLUM_B200_ERR_STATE, "work-item estimate too small");
        if (n_it) memcpy(h_items + item0, items.data(), (size_t)n_it * sizeof(WorkItem));
        const long long e0 = h_off[w0], e1 = (w1 < n_walk) ? h_off[w1] : n_epi;
        const long long a0 = h_edisp[w0], a1 = (w1 < n_walk) ? h_edisp[w1] : n_ae;
        const long long s0 = h_sdisp[w0], s1 = (w1 < n_walk) ? h_sdisp[w1] : n_as;
        if (e1 > e0) CU(cudaMemcpyAsync((EpiAos *)ws.epi.p + e0, h_epi + e0, (size_t)(e1 - e0) * sizeof(EpiAos), H2D, ci));
        if (a1 > a0) CU(cudaMemcpyAsync((int *)ws.adr_epj.p + a0, h_ae + a0, (size_t)(a1 - a0) * 4, H2D, ci));
        if (s1 > s0) CU(cudaMemcpyAsync((int *)ws.adr_spj.p + s0, h_as + s0, (size_t)(s1 - s0) * 4, H2D, ci));
        if (n_it) CU(cudaMemcpyAsync((WorkItem *)ws.items.p + item0, h_items + item0, (size_t)n_it * sizeof(WorkItem), H2D, ci));
        ws.n_items = item0 + n_it;
        const int n_seg_b = il.seg_off.empty() ? 0 : (int)il.seg_off.size() - 1;
        if (n_seg_b > 0) CU(cudaMemcpyAsync((int *)ws.seg_off.p + seg0, il.seg_off.data(), il.seg_off.size() * 4, H2D, ci));   // pageable: staged on return
        CU(cudaEventRecord(ws.ev_in[b], ci));
        CU(cudaStreamWaitEvent(st, ws.ev_in[b], 0));
        if (int r = launch_pass(ws, st, g.eps2, item0, n_it, b == 0, seg0, n_seg_b)) return r;
        seg0 += n_seg_b > 0 ? n_seg_b + 1 : 0;
        CU(cudaEventRecord(ws.ev_k[b], st));
        CU(cudaStreamWaitEvent(co, ws.ev_k[b], 0));
        if (e1 > e0) CU(cudaMemcpyAsync((ForceAos *)ws.h_force.p + e0, (const ForceAos *)ws.force.p + e0, (size_t)(e1 - e0) * sizeof(ForceAos), cudaMemcpyDeviceToHost, co));
        CU(cudaEventRecord(ws.ev_out[b], co));
        item0 += n_it;
    }
    if (!ws.done) CU(cudaEventCreateWithFlags(&ws.done, cudaEventDisableTiming));
    // `done` on the compute stream as well: work queued there after this dispatch orders behind its D2H
    if (nsb > 0) CU(cudaStreamWaitEvent(st, ws.ev_out[nsb - 1], 0));
    CU(cudaEventRecord(ws.done, st));
    ws.pending = true;
    return 0;
}

int gplum_b200_retrieve(int tag, int n_walk, const int *ni, void *const *force)
{
    if (int r = ensure_init()) return r;
    if (tag < 0 || tag >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "tag %d out of range", tag);
    WalkSet &ws = g.slots[tag];
    if (!ws.pending) return fail(GPLUM_B200_ERR_STATE, "retrieve(tag=%d) without dispatch", tag);
    if (n_walk != ws.n_walk) return fail(GPLUM_B200_ERR_ARG, "retrieve n_walk %d != dispatched %d", n_walk, ws.n_walk);
    CU(cudaSetDevice(g.device));
    const bool overwrite = (g.flags & GPLUM_B200_NO_ACCUMULATE) != 0;
    const ForceAos *src = (const ForceAos *)ws.h_force.p;
    const int nsb = (int)ws.sub_w0.size() - 1;
    for (int b = 0; b < nsb; b++) {
        CU(cudaEventSynchronize(ws.ev_out[b]));      // later sub-batches and other tags keep running
        const int w0 = ws.sub_w0[b], w1 = ws.sub_w0[b + 1];
#pragma omp parallel for schedule(static)
        for (int w = w0; w < w1; w++) {
            if (ni[w] != ws.ni_host[w]) continue;
            accumulate_force((ForceAos *)force[w], src + ws.epi_off_host[w], ni[w], overwrite);
        }
    }
    CU(cudaEventSynchronize(ws.done));
    ws.pending = false;
    for (int w = 0; w < n_walk; w++)
        if (ni[w] != ws.ni_host[w]) return fail(GPLUM_B200_ERR_ARG, "retrieve ni[%d]=%d != dispatched %d", w, ni[w], ws.ni_host[w]);
    return 0;
}

// ---- device-resident form ----
int gplum_b200_walks_select(int slot)
{
    if (slot < 0 || slot >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "slot %d out of range [0,%d)", slot, N_TAG);
    g.cur = slot;
    return 0;
}

int gplum_b200_walks_upload(int n_walk, const void *epi_all, const int *epi_off, const int *ni,
                            const int *adr_epj, const long long *epj_disp, const int *n_epj,
                            const int *adr_spj, const long long *spj_disp, const int *n_spj,
                            const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    if (epj_all || spj_all) {          // NULL, NULL: keep the current j-set (another slot / external arrays)
        if (int r = upload_j(epj_all, n_epj_all, spj_all, n_spj_all, st)) return r;
        if (int r = pack_j(st, g.eps2)) return r;
    }
    if (int r = upload_walks(g.slots[g.cur], n_walk, epi_all, epi_off, ni, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, st)) return r;
    CU(cudaStreamSynchronize(st));
    return 0;
}

int gplum_b200_walks_run(int repack)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    if (repack) if (int r = pack_j(g.stream, g.eps2)) return r;
    return launch_pass(g.slots[g.cur], g.stream, g.eps2);
}

int gplum_b200_walks_pack(void)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    return pack_j(g.stream, g.eps2);
}

int gplum_b200_walks_download(void *force_all)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    WalkSet &ws = g.slots[g.cur];
    CU(cudaStreamSynchronize(g.stream));
    if (ws.n_epi) CU(cudaMemcpy(force_all, ws.force.p, (size_t)ws.n_epi * sizeof(ForceAos), cudaMemcpyDeviceToHost));
    return 0;
}

int gplum_b200_walks_time(int iters, int repack, float *ms_per_pass)
{
    if (int r = ensure_init()) return r;
    if (iters <= 0) return fail(GPLUM_B200_ERR_ARG, "iters <= 0");
    CU(cudaSetDevice(g.device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaStreamSynchronize(g.stream));
    CU(cudaEventRecord(e0, g.stream));
    for (int i = 0; i < iters; i++) {
        if (repack) if (int r = pack_j(g.stream, g.eps2)) return r;
        if (int r = launch_pass(g.slots[g.cur], g.stream, g.eps2)) return r;
    }
    CU(cudaEventRecord(e1, g.stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_per_pass) *ms_per_pass = ms / iters;
    return 0;
}

int gplum_b200_walks_set_packed_dev(const void *epj_packed_dev, int n_epj_all, const void *spj_packed_dev, int n_spj_all)
{
    if (int r = ensure_init()) return r;
    g.jset.ext_epj = epj_packed_dev; g.jset.ext_spj = spj_packed_dev;
    if (epj_packed_dev) g.jset.n_epj = n_epj_all;
    if (spj_packed_dev) g.jset.n_spj = n_spj_all;
    return 0;
}

int gplum_b200_pack_epj_dev(const void *epj_aos_dev, int n, void *epj_packed_dev)
{
    if (int r = ensure_init()) return r;
    if (n <= 0) return 0;
    CU(cudaSetDevice(g.device));
    pack_epj_kernel<<<(n + 255) / 256, 256, 0, g.stream>>>((const EpjAos *)epj_aos_dev, n, (EpjPacked *)epj_packed_dev);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

int gplum_b200_gather_epj_packed_dev(const void *src_packed_dev, const int *idx_dev, int n, void *dst_packed_dev)
{
    if (int r = ensure_init()) return r;
    if (n <= 0) return 0;
    CU(cudaSetDevice(g.device));
    gather_epj_packed_kernel<<<(3 * n + 255) / 256, 256, 0, g.stream>>>((const uint4 *)src_packed_dev, idx_dev, n, (uint4 *)dst_packed_dev);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

// ---- multi-GPU peer mode ----
int gplum_b200_peer_setup(int world, int rank, int shift, void *handles_out)
{
    if (int r = ensure_init()) return r;
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || shift < 1 || shift > 30 || !handles_out)
        return fail(GPLUM_B200_ERR_ARG, "peer_setup(world=%d, rank=%d, shift=%d)", world, rank, shift);
    if ((long long)world << shift > 0x7fffffffLL) return fail(GPLUM_B200_ERR_ARG, "world << shift overflows the int index space");
    CU(cudaSetDevice(g.device));
    Peer &pe = g.peer;
    if (pe.slab[0]) return fail(GPLUM_B200_ERR_STATE, "peer mode already set up");
    pe.world = world; pe.rank = rank; pe.shift = shift; pe.parity = 0;
    pe.epoch = 0;
    for (int b = 0; b < 2; b++) {
        // slab 0 carries this rank's flag array (MAX_PEERS ints) behind the records
        const size_t bytes = ((size_t)1 << shift) * sizeof(EpjPacked) + (b == 0 ? 256 : 0);
        CU(cudaMalloc(&pe.slab[b], bytes));
        CU(cudaMemset(pe.slab[b], 0, bytes));
        CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handles_out + b, pe.slab[b]));
    }
    CU(cudaDeviceSynchronize());
    return 0;
}

int gplum_b200_peer_open(const void *all_handles)
{
    if (int r = ensure_init()) return r;
    Peer &pe = g.peer;
    if (!pe.slab[0] || !all_handles) return fail(GPLUM_B200_ERR_STATE, "peer_open before peer_setup");
    CU(cudaSetDevice(g.device));
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
    for (int b = 0; b < 2; b++) {
        for (int q = 0; q < pe.world; q++) {
            if (q == pe.rank) { pe.mapped[b][q] = pe.slab[b]; continue; }
            CU(cudaIpcOpenMemHandle(&pe.mapped[b][q], h[2 * q + b], cudaIpcMemLazyEnablePeerAccess));
        }
        if (int r = pe.table[b].reserve(sizeof(void *) * MAX_PEERS)) return r;
        CU(cudaMemcpy(pe.table[b].p, pe.mapped[b], sizeof(void *) * pe.world, cudaMemcpyHostToDevice));
    }
    if (int r = pe.done.reserve(16)) return r;
    CU(cudaMemset(pe.done.p, 0, 16));
    pe.on = true;
    return 0;
}

int gplum_b200_peer_pack(const void *epj_aos_dev, int n)
{
    if (int r = ensure_init()) return r;
    Peer &pe = g.peer;
    if (!pe.on) return fail(GPLUM_B200_ERR_STATE, "peer_pack without peer_open");
    if (n < 0 || n > (1 << pe.shift)) return fail(GPLUM_B200_ERR_ARG, "peer_pack n=%d exceeds the slab (%d)", n, 1 << pe.shift);
    CU(cudaSetDevice(g.device));
    if (pe.pending) if (int r = launch_peer_pack(g.stream)) return r;       // two packs in a row: the first one goes out now
    pe.parity ^= 1;                  // peers may still be reading the slab of the previous step
    pe.epoch++;
    // the pack itself -- this rank's EPJ into its slab, its superparticles (the current j-set's, unless they are
    // external) and the flag stores that tell every rank "packed" -- goes out with the next force launch
    // (launch_pass): in that kernel's prologue when the pass is placed, else as peer_pack_kernel right before it
    pe.pending = true; pe.pending_epj = epj_aos_dev; pe.pending_n = n;
    return 0;
}

int gplum_b200_peer_wait(void)
{
    if (int r = ensure_init()) return r;
    Peer &pe = g.peer;
    if (!pe.on) return fail(GPLUM_B200_ERR_STATE, "peer_wait without peer_open");
    CU(cudaSetDevice(g.device));
    if (pe.pending) if (int r = launch_peer_pack(g.stream)) return r;
    const int *flags = reinterpret_cast<const int *>(static_cast<const char *>(pe.slab[0]) + ((size_t)1 << pe.shift) * sizeof(EpjPacked));
    peer_wait_kernel<<<1, 32, 0, g.stream>>>(flags, pe.world, pe.epoch);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

int gplum_b200_peer_close(void)
{
    Peer &pe = g.peer;
    if (!pe.slab[0]) return 0;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    pe.pending = false;
    for (int b = 0; b < 2; b++) {
        for (int q = 0; q < pe.world; q++)
            if (pe.on && q != pe.rank && pe.mapped[b][q]) cudaIpcCloseMemHandle(pe.mapped[b][q]);
        pe.table[b].release();
    }
    pe.done.release();
    // the slabs themselves are freed by the caller's barrier-then-free protocol: peers must have
    // closed their mappings first (see gplum_b200/multigpu.py)
    pe.on = false;
    return 0;
}

int gplum_b200_peer_free(void)
{
    Peer &pe = g.peer;
    cudaSetDevice(g.device);
    for (int b = 0; b < 2; b++) { if (pe.slab[b]) cudaFree(pe.slab[b]); pe.slab[b] = nullptr; }
    pe = Peer();
    return 0;
}

int gplum_b200_pack_spj_dev(const void *spj_aos_dev, int n, void *spj_packed_dev)
{
    if (int r = ensure_init()) return r;
    if (n <= 0) return 0;
    CU(cudaSetDevice(g.device));
    pack_spj_kernel<<<(n + 255) / 256, 256, 0, g.stream>>>(spj_aos_dev, n, (SpjPacked *)spj_packed_dev, g.quad,
                                                          (g.flags & GPLUM_B200_TRACE_AS_SHIPPED) ? 1 : 0, g.eps2);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

// ---- changeover correction (soft_corr.cu) ----
namespace {
// Reads the correction's status words (16 B, one stream sync; once per correction): dropped pairs and i-particles
// missing from their own EP list make every consumer of corr_out fail, not only the download calls.
int corr_status_check(WalkSet &ws, unsigned int *meta_out = nullptr, int *total_out = nullptr)
{
    unsigned int meta[4] = {0, 0, 0, 0};
    int total = 0;
    const int n = (int)ws.n_epi;
    if (n > 0) {
        CU(cudaMemcpyAsync(meta, ws.corr_meta.p, 16, cudaMemcpyDeviceToHost, g.stream));
        CU(cudaMemcpyAsync(&total, (const int *)ws.off.p + n, 4, cudaMemcpyDeviceToHost, g.stream));
        CU(cudaStreamSynchronize(g.stream));
    }
    if (meta_out) memcpy(meta_out, meta, 16);
    if (total_out) *total_out = total;
    if (meta[1] > 0)
        return fail(GPLUM_B200_ERR_OVERFLOW, "%u candidate pairs dropped: pair buffer holds %u, pass produced %u", meta[1], ws.pair_cap, meta[0]);
    if (meta[2] > 0)
        return fail(GPLUM_B200_ERR_STATE, "%u i-particles are not in their own EP list (not an FDPS interaction list)", meta[2]);
    if ((long long)total != (long long)meta[0])
        return fail(GPLUM_B200_ERR_STATE, "candidate counts (%d) and captured pairs (%u) disagree", total, meta[0]);
    ws.corr_checked = true;
    return 0;
}
}  // namespace

int gplum_b200_soft_corr_enable(int on, long long pair_cap)
{
    if (pair_cap < 0) return fail(GPLUM_B200_ERR_ARG, "pair_cap < 0");
    g.corr_on = on != 0;
    g.corr_cap = pair_cap;
    return 0;
}

int gplum_b200_correct_long_run(int slot, const gplum_b200_corr_params *prm, int initial)
{
    if (int r = ensure_init()) return r;
    if (slot < 0 || slot >= N_TAG || !prm) return fail(GPLUM_B200_ERR_ARG, "correct_long_run(slot=%d)", slot);
    WalkSet &ws = g.slots[slot];
    if (!ws.captured) return fail(GPLUM_B200_ERR_STATE, "no captured pass in slot %d: call soft_corr_enable(1, cap) before the pass", slot);
    if (g.jset.ext_epj || g.peer.on || !g.jset.epj_aos.p)
        return fail(GPLUM_B200_ERR_STATE, "the correction needs the raw EPJGrav array of the pass on this device");
    CU(cudaSetDevice(g.device));
    const int n = (int)ws.n_epi;
    if (n == 0) { ws.corrected = true; ws.corr_checked = true; return 0; }
    if (int r = ws.cnt.reserve((size_t)(n + 1) * 4)) return r;
    if (int r = ws.off.reserve((size_t)(n + 1) * 4)) return r;
    if (int r = ws.cursor.reserve((size_t)(n + 1) * 4)) return r;
    if (int r = ws.csr.reserve((size_t)ws.pair_cap * 4)) return r;
    if (int r = ws.corr_out.reserve((size_t)n * sizeof(SoftCorr))) return r;
    if (initial) if (int r = ws.corr_init.reserve((size_t)n * sizeof(SoftCorrInit))) return r;
    if (int r = ws.ngb.reserve((size_t)ws.pair_cap * sizeof(SoftNgb))) return r;
    const size_t tb = soft_corr_scan_temp_bytes(n);
    if (int r = ws.scan_temp.reserve(tb)) return r;
    SoftCorrArgs a;
    a.n_epi = n; a.i0 = ws.part_e1 > ws.part_e0 ? ws.part_e0 : 0; a.i1 = ws.part_e1 > ws.part_e0 ? ws.part_e1 : n; a.epi = ws.epi.p; a.force = ws.force.p; a.epj_aos = g.jset.epj_aos.p;
    a.self_adr = (const int *)ws.self_adr.p;
    a.pairs = (const int2 *)ws.pairs.p; a.pair_count = (const unsigned int *)ws.corr_meta.p; a.pair_cap = ws.pair_cap;
    a.cnt = (int *)ws.cnt.p; a.off = (int *)ws.off.p; a.cursor = (int *)ws.cursor.p; a.csr = (int *)ws.csr.p;
    a.out = (SoftCorr *)ws.corr_out.p; a.init_out = initial ? (SoftCorrInit *)ws.corr_init.p : nullptr;
    a.ngb = (SoftNgb *)ws.ngb.p;
    a.status = (unsigned int *)ws.corr_meta.p + 1;
    a.prm.eps2 = prm->eps2; a.prm.dt_tree = prm->dt_tree; a.prm.gamma = prm->gamma;
    a.prm.R_search2 = prm->R_search2; a.prm.R_search3 = prm->R_search3;
    a.prm.re_search = prm->re_search; a.prm.initial = initial ? 1 : 0;
    CU(cudaMemsetAsync((unsigned int *)ws.corr_meta.p + 1, 0, 12, g.stream));
    int launched = 0;
    const int e = soft_corr_launch(a, ws.scan_temp.p, ws.scan_temp.cap, g.stream, &launched);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "soft_corr_launch -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches += launched;
    ws.corrected = true; ws.corrected_initial = initial != 0; ws.corr_checked = false;
    return 0;
}

int gplum_b200_correct_long_download(int slot, void *corr_out, void *init_out, void *ngb_out,
                                     long long ngb_cap, long long *n_ngb_slots, long long *n_pairs)
{
    if (int r = ensure_init()) return r;
    if (slot < 0 || slot >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "slot %d out of range", slot);
    WalkSet &ws = g.slots[slot];
    if (!ws.corrected) return fail(GPLUM_B200_ERR_STATE, "correct_long_download before correct_long_run");
    if (init_out && !ws.corrected_initial) return fail(GPLUM_B200_ERR_STATE, "init_out requested but the run was not `initial`");
    CU(cudaSetDevice(g.device));
    const int n = (int)ws.n_epi;
    if (n_ngb_slots) *n_ngb_slots = 0;
    if (n_pairs) *n_pairs = 0;
    if (n == 0) return 0;
    unsigned int meta[4] = {0, 0, 0, 0};
    int total = 0;
    if (corr_out) CU(cudaMemcpyAsync(corr_out, ws.corr_out.p, (size_t)n * sizeof(SoftCorr), cudaMemcpyDeviceToHost, g.stream));
    if (init_out) CU(cudaMemcpyAsync(init_out, ws.corr_init.p, (size_t)n * sizeof(SoftCorrInit), cudaMemcpyDeviceToHost, g.stream));
    const int rs = corr_status_check(ws, meta, &total);
    if (n_pairs) *n_pairs = meta[0];
    if (rs) return rs;
    if (n_ngb_slots) *n_ngb_slots = total;
    if (ngb_out && total > 0) {
        if (total > ngb_cap) return fail(GPLUM_B200_ERR_ARG, "ngb_out holds %lld entries, %d needed", ngb_cap, total);
        CU(cudaMemcpy(ngb_out, ws.ngb.p, (size_t)total * sizeof(SoftNgb), cudaMemcpyDeviceToHost));
    }
    return 0;
}

// Like correct_long_download, but only the particles that have neighbours cross PCIe (stable order);
// corr_out holds room for corr_cap records, *n_corr receives how many were written.
int gplum_b200_correct_long_download_compact(int slot, void *corr_out, long long corr_cap, long long *n_corr,
                                             void *ngb_out, long long ngb_cap, long long *n_ngb_slots, long long *n_pairs)
{
    if (int r = ensure_init()) return r;
    if (slot < 0 || slot >= N_TAG || !corr_out || !n_corr) return fail(GPLUM_B200_ERR_ARG, "correct_long_download_compact: bad argument");
    WalkSet &ws = g.slots[slot];
    if (!ws.corrected) return fail(GPLUM_B200_ERR_STATE, "correct_long_download_compact before correct_long_run");
    CU(cudaSetDevice(g.device));
    const int n = (int)ws.n_epi;
    *n_corr = 0;
    if (n_ngb_slots) *n_ngb_slots = 0;
    if (n_pairs) *n_pairs = 0;
    if (n == 0) return 0;
    cudaStream_t st = g.stream;
    const size_t tb = soft_corr_compact_temp_bytes(n);
    if (int r = ws.scan_temp.reserve(tb)) return r;
    if (int r = ws.corr_compact.reserve((size_t)n * sizeof(SoftCorr) + 16)) return r;
    int *d_cnt = (int *)((char *)ws.corr_compact.p + (size_t)n * sizeof(SoftCorr));
    const int e = soft_corr_compact(n, (const SoftCorr *)ws.corr_out.p, (SoftCorr *)ws.corr_compact.p, d_cnt, ws.scan_temp.p, ws.scan_temp.cap, st);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "soft_corr_compact -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches++;
    unsigned int meta[4] = {0, 0, 0, 0};
    int total = 0, m = 0;
    CU(cudaMemcpyAsync(&m, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    const int rs = corr_status_check(ws, meta, &total);
    if (n_pairs) *n_pairs = meta[0];
    if (rs) return rs;
    if (m > corr_cap) return fail(GPLUM_B200_ERR_ARG, "corr_out holds %lld records, %d needed", corr_cap, m);
    *n_corr = m;
    if (n_ngb_slots) *n_ngb_slots = total;
    if (m > 0) CU(cudaMemcpyAsync(corr_out, ws.corr_compact.p, (size_t)m * sizeof(SoftCorr), cudaMemcpyDeviceToHost, st));
    if (ngb_out && total > 0) {
        if (total > ngb_cap) return fail(GPLUM_B200_ERR_ARG, "ngb_out holds %lld entries, %d needed", ngb_cap, total);
        CU(cudaMemcpyAsync(ngb_out, ws.ngb.p, (size_t)total * sizeof(SoftNgb), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return 0;
}

int gplum_b200_correct_long_time(int slot, const gplum_b200_corr_params *prm, int initial, int iters, float *ms_out)
{
    if (int r = ensure_init()) return r;
    if (iters <= 0) return fail(GPLUM_B200_ERR_ARG, "iters <= 0");
    CU(cudaSetDevice(g.device));
    if (int r = gplum_b200_correct_long_run(slot, prm, initial)) return r;     // sizes the buffers
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaStreamSynchronize(g.stream));
    CU(cudaEventRecord(e0, g.stream));
    for (int i = 0; i < iters; i++)
        if (int r = gplum_b200_correct_long_run(slot, prm, initial)) return r;
    CU(cudaEventRecord(e1, g.stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms / iters;
    return 0;
}

int gplum_b200_fp32_peak(int iters, float *tflops, float *ms_out)
{
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, g.device));
    float *d = nullptr;
    CU(cudaMalloc(&d, 4));
    const int blocks = prop.multiProcessorCount * 8, inner = 4096;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    fp32_peak_kernel<<<blocks, 256, 0, g.stream>>>(d, inner, 1.0000001f, 1e-9f);
    CU(cudaStreamSynchronize(g.stream));
    CU(cudaEventRecord(e0, g.stream));
    for (int i = 0; i < iters; i++) fp32_peak_kernel<<<blocks, 256, 0, g.stream>>>(d, inner, 1.0000001f, 1e-9f);
    CU(cudaEventRecord(e1, g.stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    const double flop = 2.0 * 16.0 * inner * 256.0 * blocks * iters;
    if (tflops) *tflops = (float)(flop / (ms * 1e-3) / 1e12);
    if (ms_out) *ms_out = ms / iters;
    return 0;
}

}  // extern "C"

// ---- interaction lists built on the GPU (dev_tree.cu) ----
namespace {
// Host -> device copy of a caller's array that may be PAGEABLE (FDPS's ReallocatableArrays, std::vectors): a plain
// cudaMemcpyAsync from pageable memory is staged by the driver at a fraction of the PCIe rate.  Here the array is cut
// into chunks that OpenMP threads copy into two pinned staging buffers while the previous chunk is on the wire.
// Pinned arrays go straight through.  The copy is complete on the stream, not on return.
constexpr size_t STAGE_CHUNK = 8u << 20;
int upload_host(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return 0;
    cudaPointerAttributes at;
    const cudaError_t pe = cudaPointerGetAttributes(&at, src);
    if (pe != cudaSuccess) cudaGetLastError();
    if (pe == cudaSuccess && at.type != cudaMemoryTypeUnregistered) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    if (int r = g.tree_pin.reserve(2 * STAGE_CHUNK)) return r;
    if (!g.ev_stage[0]) for (auto &e : g.ev_stage) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    char *pin[2] = {(char *)g.tree_pin.p, (char *)g.tree_pin.p + STAGE_CHUNK};
    int k = 0;
    for (size_t off = 0; off < bytes; off += STAGE_CHUNK, k ^= 1) {
        const size_t nb = std::min(STAGE_CHUNK, bytes - off);
        if (g.stage_used[k]) CU(cudaEventSynchronize(g.ev_stage[k]));          // the copy that last used this buffer is done
        const char *s0 = (const char *)src + off;
        const long long n_part = (long long)((nb + (1u << 20) - 1) >> 20);
#pragma omp parallel for schedule(static)
        for (long long q = 0; q < n_part; q++) {
            const size_t a = (size_t)q << 20, b = std::min(nb, a + (1u << 20));
            memcpy(pin[k] + a, s0 + a, b - a);
        }
        CU(cudaMemcpyAsync((char *)dst + off, pin[k], nb, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(g.ev_stage[k], st));
        g.stage_used[k] = true;
    }
    return 0;
}

// Device -> host copy into a caller's array that may be pageable: chunks land in the pinned staging buffers and
// OpenMP threads copy chunk k out while chunk k+1 is on the wire.  Complete on return.
int download_host(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return 0;
    cudaPointerAttributes at;
    const cudaError_t pe = cudaPointerGetAttributes(&at, dst);
    if (pe != cudaSuccess) cudaGetLastError();
    if (pe == cudaSuccess && at.type != cudaMemoryTypeUnregistered) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return 0;
    }
    if (int r = g.tree_pin.reserve(2 * STAGE_CHUNK)) return r;
    if (!g.ev_stage[0]) for (auto &e : g.ev_stage) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) if (g.stage_used[k]) { CU(cudaEventSynchronize(g.ev_stage[k])); g.stage_used[k] = false; }
    char *pin[2] = {(char *)g.tree_pin.p, (char *)g.tree_pin.p + STAGE_CHUNK};
    const size_t n_chunk = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    for (size_t c = 0; c <= n_chunk; c++) {
        if (c < n_chunk) {
            const size_t off = c * STAGE_CHUNK, nb = std::min(STAGE_CHUNK, bytes - off);
            CU(cudaMemcpyAsync(pin[c & 1], (const char *)src + off, nb, cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(g.ev_stage[c & 1], st));
        }
        if (c > 0) {
            const size_t off = (c - 1) * STAGE_CHUNK, nb = std::min(STAGE_CHUNK, bytes - off);
            CU(cudaEventSynchronize(g.ev_stage[(c - 1) & 1]));
            const long long n_part = (long long)((nb + (1u << 20) - 1) >> 20);
            const char *p0 = pin[(c - 1) & 1];
#pragma omp parallel for schedule(static)
            for (long long q = 0; q < n_part; q++) {
                const size_t a = (size_t)q << 20, b = std::min(nb, a + (1u << 20));
                memcpy((char *)dst + off + a, p0 + a, b - a);
            }
        }
    }
    return 0;
}

int tree_build_common(int n, const gbt::TreeSrc &src, double theta, int n_leaf_limit, int n_group_limit, long long *sizes,
                      int part_rank = 0, int part_world = 1)
{
    if (g.rmax > 2) return fail(GPLUM_B200_ERR_STATE, "the GPU list builder needs the RMAX <= 2 kernel");
    cudaStream_t st = g.stream;
    WalkSet &ws = g.slots[g.cur];
    JSet &j = g.jset;
    g.tree_built = false; g.tree_inv_valid = false;
    if (int r = ws.epi.reserve((size_t)n * sizeof(EpiAos))) return r;
    if (int r = ws.force.reserve((size_t)n * sizeof(ForceAos))) return r;
    if (int r = j.epj_aos.reserve((size_t)n * sizeof(EpjAos))) return r;
    if (int r = j.epj_packed.reserve((size_t)n * sizeof(EpjPacked))) return r;
    gbt::TreeCfg cfg;
    cfg.n = n; cfg.theta = theta; cfg.n_leaf = n_leaf_limit; cfg.n_group = n_group_limit; cfg.quad = g.quad;
    cfg.warp_slots = g.warp_slots; cfg.tile_cap = g.tile_cap; cfg.jsplit = g.jsplit; cfg.rmax = g.rmax; cfg.split_m = g.split_m;
    cfg.part_rank = part_rank; cfg.part_world = part_world;
    gbt::TreeCounts c;
    memset(&c, 0, sizeof(c));
    int launches = 0;
    int e = gbt::tree_phase1(cfg, src, j.epj_aos.p, ws.epi.p, j.epj_packed.p, &c, st, &launches);
    g.launches += launches;
    if (e > 0) return fail(GPLUM_B200_ERR_CUDA, "GPU list builder, phase 1: %s", cudaGetErrorString((cudaError_t)e));
    if (e < 0) return fail(GPLUM_B200_ERR_OVERFLOW, "GPU list builder: %s overflow", c.overflow == 1 ? "cell capacity" : "walk stack");
    const size_t ssz = g.quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos);
    const size_t nw = (size_t)c.n_walk;
    if (int r = ws.epi_off.reserve(nw * 4)) return r;
    if (int r = ws.n_epj.reserve(nw * 4)) return r;
    if (int r = ws.n_spj.reserve(nw * 4)) return r;
    if (int r = ws.epj_disp.reserve(nw * 8)) return r;
    if (int r = ws.spj_disp.reserve(nw * 8)) return r;
    if (int r = ws.adr_epj.reserve((size_t)c.n_adr_epj * 4)) return r;
    if (int r = ws.adr_spj.reserve((size_t)c.n_adr_spj * 4)) return r;
    // a pass with few items cuts its tiles along j on the device too (dev_tree.cu: item_split_kernel); the list is
    // then sized by its upper bound and padded with empty items, so that no further host sync is needed
    const bool dev_split = split_active(c.n_items, g.warp_slots, g.split_m);
    const int n_items_out = dev_split ? (int)split_items_bound(c.n_items, g.warp_slots) : c.n_items;
    if (int r = ws.items.reserve((size_t)n_items_out * sizeof(WorkItem))) return r;
    if (int r = reserve_split(ws, dev_split ? n_items_out : 1, dev_split ? c.n_items : 1, st)) return r;
    if (int r = ws.seg_off.reserve((size_t)(g.warp_slots + 1) * 4)) return r;
    ws.n_seg = dev_split ? (int)g.warp_slots : 0;
    if (!g.quad) if (int r = j.spj_aos.reserve((size_t)c.n_cells * ssz)) return r;
    if (int r = j.spj_packed.reserve((size_t)c.n_cells * sizeof(SpjPacked))) return r;
    gbt::TreeOut o;
    o.epi_off = (int *)ws.epi_off.p; o.ni = nullptr; o.n_epj = (int *)ws.n_epj.p; o.n_spj = (int *)ws.n_spj.p;
    o.epj_disp = (long long *)ws.epj_disp.p; o.spj_disp = (long long *)ws.spj_disp.p;
    o.adr_epj = (int *)ws.adr_epj.p; o.adr_spj = (int *)ws.adr_spj.p;
    o.items = ws.items.p; o.n_items_out = n_items_out; o.seg_off = (int *)ws.seg_off.p;
    o.spj_aos = g.quad ? nullptr : j.spj_aos.p;           // quadrupole records ARE the cells' moment records: read in place
    launches = 0;
    e = gbt::tree_phase2(cfg, o, st, &launches);
    g.launches += launches;
    if (e) return fail(GPLUM_B200_ERR_CUDA, "GPU list builder, phase 2: %s", cudaGetErrorString((cudaError_t)e));
    ws.n_walk = c.n_walk; ws.n_items = n_items_out; ws.n_epi = n;
    ws.part_e0 = part_world > 1 ? c.e0 : 0; ws.part_e1 = part_world > 1 ? c.e1 : 0;
    ws.n_adr_epj = c.n_adr_epj; ws.n_adr_spj = c.n_adr_spj;
    ws.n_int_epep = c.n_int_epep; ws.n_int_epsp = c.n_int_epsp;
    ws.ni_host.clear(); ws.epi_off_host.clear();
    ws.pending = false; ws.captured = false; ws.corrected = false;
    j.ext_epj = j.ext_spj = nullptr;
    j.n_epj = n; j.n_spj = c.n_cells;
    j.spj_src = g.quad ? gbt::tree_cell_moments() : nullptr;
    j.epj_packed_fresh = true;                           // written by the gather (dev_tree.cu)
    if (int r = pack_j(st, g.eps2)) return r;
    g.tree_built = true;
    if (sizes) {
        sizes[0] = c.n_walk; sizes[1] = n; sizes[2] = c.n_adr_epj; sizes[3] = c.n_adr_spj; sizes[4] = n;
        sizes[5] = c.n_cells; sizes[6] = c.n_int_epep; sizes[7] = c.n_int_epsp;
        if (part_world > 1) { sizes[8] = c.w0; sizes[9] = c.w1; sizes[10] = c.e0; sizes[11] = c.e1; }
    }
    return 0;
}
// Particles as host columns (pos [n][3], mass, r_out, r_search, optional vel [n][3]).  Keys and sort need the positions
// only, so those go up first and the other columns follow on the copy engine while the GPU sorts (dev_tree.h:
// TreeSrc::before_gather); pageable columns are staged in that same window.
struct ColumnUpload {
    const double *cols[4]; size_t bytes[4]; double *dst[4]; int n_cols;
};
int upload_rest_columns(void *arg)
{
    ColumnUpload &u = *static_cast<ColumnUpload *>(arg);
    for (int k = 0; k < u.n_cols; k++)
        if (upload_host(u.dst[k], u.cols[k], u.bytes[k], g.copy_in)) return (int)cudaErrorUnknown;     // last_error holds the text
    if (cudaEventRecord(g.ev_cols, g.copy_in) != cudaSuccess) return (int)cudaGetLastError();
    if (cudaStreamWaitEvent(g.stream, g.ev_cols, 0) != cudaSuccess) return (int)cudaGetLastError();
    return 0;
}
int tree_build_columns(int n, const double *pos, const double *mass, const double *r_out, const double *r_search, const double *vel,
                       double theta, int n_leaf_limit, int n_group_limit, int rank, long long *sizes)
{
    cudaStream_t st = g.stream;
    const size_t N = (size_t)n;
    if (!g.copy_in) CU(cudaStreamCreateWithFlags(&g.copy_in, cudaStreamNonBlocking));
    if (!g.ev_cols) { CU(cudaEventCreateWithFlags(&g.ev_cols, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&g.ev_cols0, cudaEventDisableTiming)); }
    if (g.tree_in.cap < N * 72) {
        CU(cudaStreamSynchronize(st));                      // a build in flight may still read the old columns
        if (int r = g.tree_in.reserve(N * 72)) return r;
    }
    double *d = (double *)g.tree_in.p;
    // the copy engine starts once everything enqueued so far (the previous build's gather) is done with the columns
    CU(cudaEventRecord(g.ev_cols0, st));
    CU(cudaStreamWaitEvent(g.copy_in, g.ev_cols0, 0));
    if (int r = upload_host(d, pos, N * 24, g.copy_in)) return r;
    CU(cudaEventRecord(g.ev_cols, g.copy_in));
    CU(cudaStreamWaitEvent(st, g.ev_cols, 0));
    ColumnUpload u;
    u.n_cols = 0;
    auto add = [&](const double *c, size_t bytes, double *dst) { u.cols[u.n_cols] = c; u.bytes[u.n_cols] = bytes; u.dst[u.n_cols] = dst; u.n_cols++; };
    add(mass, N * 8, d + 3 * N); add(r_out, N * 8, d + 4 * N); add(r_search, N * 8, d + 5 * N);
    if (vel) add(vel, N * 24, d + 6 * N);
    gbt::TreeSrc ts;
    ts.pos = d; ts.mass = d + 3 * N; ts.r_out = d + 4 * N; ts.r_search = d + 5 * N; ts.vel = vel ? d + 6 * N : nullptr;
    ts.rank = rank;
    ts.before_gather = upload_rest_columns; ts.before_gather_arg = &u;
    return tree_build_common(n, ts, theta, n_leaf_limit, n_group_limit, sizes);
}
}  // namespace

extern "C" {

int gplum_b200_tree_build_gpu(int n, const double *pos, const double *mass, const double *r_out,
                              const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                              int rank, long long *sizes)
{
    if (n <= 0 || !pos || !mass || !r_out || !r_search || theta <= 0.0) return fail(GPLUM_B200_ERR_ARG, "tree_build_gpu: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    return tree_build_columns(n, pos, mass, r_out, r_search, nullptr, theta, n_leaf_limit, n_group_limit, rank, sizes);
}

/* The same with the velocities (the changeover correction's neighbour re-search and its `initial` form read them,
 * src/gravity_soft.h:300-346): vel is [n][3] or NULL. */
int gplum_b200_tree_build_gpu_vel(int n, const double *pos, const double *vel, const double *mass, const double *r_out,
                                  const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                                  int rank, long long *sizes)
{
    if (n <= 0 || !pos || !mass || !r_out || !r_search || theta <= 0.0) return fail(GPLUM_B200_ERR_ARG, "tree_build_gpu_vel: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    return tree_build_columns(n, pos, mass, r_out, r_search, vel, theta, n_leaf_limit, n_group_limit, rank, sizes);
}

int gplum_b200_tree_build_gpu_epj(int n, const void *epj, int on_device, double theta, int n_leaf_limit,
                                  int n_group_limit, long long *sizes)
{
    if (n <= 0 || !epj || theta <= 0.0) return fail(GPLUM_B200_ERR_ARG, "tree_build_gpu_epj: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    const void *src = epj;
    if (!on_device) {
        if (int r = g.tree_raw.reserve((size_t)n * sizeof(EpjAos))) return r;
        if (int r = upload_host(g.tree_raw.p, epj, (size_t)n * sizeof(EpjAos), g.stream)) return r;
        src = g.tree_raw.p;
    }
    gbt::TreeSrc ts;
    ts.epj = src;
    return tree_build_common(n, ts, theta, n_leaf_limit, n_group_limit, sizes);
}

// Multi-GPU form: every rank hands over the SAME n particles (device pointer: the all-gathered EPJGrav
// records of all ranks) and builds the same tree, but walks, lists, work items -- and hence forces and corrections --
// only for its share: the walks whose first particle in tree order lies in [n r / W, n (r+1) / W).
// sizes[12]: [0..7] as tree_build_gpu (counts of THIS rank's lists), [8], [9] = its walks [w0, w1), [10], [11] = its
// i-particles [e0, e1) in tree order.
int gplum_b200_tree_build_gpu_part(int n, const void *epj_dev, double theta, int n_leaf_limit, int n_group_limit,
                                   int part_rank, int part_world, long long *sizes)
{
    if (n <= 0 || !epj_dev || theta <= 0.0 || part_world < 1 || part_rank < 0 || part_rank >= part_world)
        return fail(GPLUM_B200_ERR_ARG, "tree_build_gpu_part: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    gbt::TreeSrc ts;
    ts.epj = epj_dev;
    return tree_build_common(n, ts, theta, n_leaf_limit, n_group_limit, sizes, part_rank, part_world);
}

// The same from 48 B records {pos[3], mass, r_out, r_search} (device pointer, the all-gather of every rank's records):
// what the interaction kernels read of EPJGrav and nothing else -- 48 instead of 112 B per particle on NVLink.
// id_local = id = index in the gathered array, myrank = 0, vel = acc_d = 0.
int gplum_b200_tree_build_gpu_part_rec48(int n, const double *rec_dev, double theta, int n_leaf_limit, int n_group_limit,
                                         int part_rank, int part_world, long long *sizes)
{
    if (n <= 0 || !rec_dev || theta <= 0.0 || part_world < 1 || part_rank < 0 || part_rank >= part_world)
        return fail(GPLUM_B200_ERR_ARG, "tree_build_gpu_part_rec48: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    gbt::TreeSrc ts;
    ts.pos = rec_dev; ts.mass = rec_dev + 3; ts.r_out = rec_dev + 4; ts.r_search = rec_dev + 5;
    ts.pos_stride = 6; ts.col_stride = 6;
    return tree_build_common(n, ts, theta, n_leaf_limit, n_group_limit, sizes, part_rank, part_world);
}

// ForceGrav[count] of i-particles [first, first + count) of the selected walk set (tree order), host pointer
int gplum_b200_walks_download_range(void *force_out, long long first, long long count)
{
    if (int r = ensure_init()) return r;
    WalkSet &ws = g.slots[g.cur];
    if (first < 0 || count < 0 || first + count > ws.n_epi || (count > 0 && !force_out)) return fail(GPLUM_B200_ERR_ARG, "walks_download_range: bad range");
    CU(cudaSetDevice(g.device));
    if (count == 0) { CU(cudaStreamSynchronize(g.stream)); return 0; }
    return download_host(force_out, (const ForceAos *)ws.force.p + first, (size_t)count * sizeof(ForceAos), g.stream);
}

int gplum_b200_tree_copy_gpu(void *epi, int *epi_off, int *ni, int *adr_epj, long long *epj_disp, int *n_epj,
                             int *adr_spj, long long *spj_disp, int *n_spj, void *epj_all, void *spj_all,
                             int *sorted_to_original)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_copy_gpu: no GPU-built tree in the selected slot");
    CU(cudaSetDevice(g.device));
    WalkSet &ws = g.slots[g.cur];
    JSet &j = g.jset;
    CU(cudaStreamSynchronize(g.stream));
    const cudaMemcpyKind D2H = cudaMemcpyDeviceToHost;
    const size_t nw = (size_t)ws.n_walk, ssz = g.quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos);
    if (epi) CU(cudaMemcpy(epi, ws.epi.p, (size_t)ws.n_epi * sizeof(EpiAos), D2H));
    if (epj_all) CU(cudaMemcpy(epj_all, j.epj_aos.p, (size_t)j.n_epj * sizeof(EpjAos), D2H));
    if (spj_all && j.n_spj) CU(cudaMemcpy(spj_all, j.spj_records(), (size_t)j.n_spj * ssz, D2H));
    if (sorted_to_original) CU(cudaMemcpy(sorted_to_original, gbt::tree_sorted_to_original(), (size_t)ws.n_epi * 4, D2H));
    if (nw) {
        if (epi_off) CU(cudaMemcpy(epi_off, ws.epi_off.p, nw * 4, D2H));
        if (ni) CU(cudaMemcpy(ni, gbt::tree_walk_ni(), nw * 4, D2H));
        if (n_epj) CU(cudaMemcpy(n_epj, ws.n_epj.p, nw * 4, D2H));
        if (n_spj) CU(cudaMemcpy(n_spj, ws.n_spj.p, nw * 4, D2H));
        if (epj_disp) CU(cudaMemcpy(epj_disp, ws.epj_disp.p, nw * 8, D2H));
        if (spj_disp) CU(cudaMemcpy(spj_disp, ws.spj_disp.p, nw * 8, D2H));
    }
    if (adr_epj && ws.n_adr_epj) CU(cudaMemcpy(adr_epj, ws.adr_epj.p, (size_t)ws.n_adr_epj * 4, D2H));
    if (adr_spj && ws.n_adr_spj) CU(cudaMemcpy(adr_spj, ws.adr_spj.p, (size_t)ws.n_adr_spj * 4, D2H));
    return 0;
}

// forces of the GPU-built tree's pass in the order of the particles as they were handed in (FDPS's
// copyForceOriginalOrder + writeBack, FDPS/src/tree_for_force_impl.hpp:873-883): scattered on the device, one D2H
int gplum_b200_tree_download_original(void *force_out)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_download_original: no GPU-built tree in the selected slot");
    if (!force_out) return fail(GPLUM_B200_ERR_ARG, "tree_download_original: NULL");
    CU(cudaSetDevice(g.device));
    WalkSet &ws = g.slots[g.cur];
    const int n = (int)ws.n_epi;
    if (int r = ws.force_org.reserve((size_t)n * sizeof(ForceAos))) return r;
    unsort_force_kernel<<<(2 * n + 255) / 256, 256, 0, g.stream>>>(n, (const uint4 *)ws.force.p, gbt::tree_sorted_to_original(), (uint4 *)ws.force_org.p);
    CU(cudaGetLastError());
    g.launches++;
    return download_host(force_out, ws.force_org.p, (size_t)n * sizeof(ForceAos), g.stream);
}

// What the changeover correction reads besides the positions (src/gravity_soft.h:76-242: relative velocity, and the
// difference of the direct accelerations in the neighbour re-search): columns in particle order, written into the
// resident tree-order records of the last GPU build.  Either may be NULL (zeros).
int gplum_b200_tree_set_motion(int n, const double *vel, const double *acc_d)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_set_motion: no GPU-built tree in the selected slot");
    WalkSet &ws = g.slots[g.cur];
    if (n != (int)ws.n_epi || n != g.jset.n_epj) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion: n = %d, the tree holds %lld particles", n, (long long)ws.n_epi);
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    const size_t N = (size_t)n;
    if (g.tree_motion.cap < N * 48) {
        CU(cudaStreamSynchronize(st));
        if (int r = g.tree_motion.reserve(N * 48)) return r;
    }
    double *d = (double *)g.tree_motion.p;
    if (vel) if (int r = upload_host(d, vel, N * 24, st)) return r;
    if (acc_d) if (int r = upload_host(d + 3 * N, acc_d, N * 24, st)) return r;
    set_motion_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, gbt::tree_sorted_to_original(), vel ? d : nullptr, acc_d ? d + 3 * N : nullptr,
                                                       (EpjAos *)g.jset.epj_aos.p);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

// The same for m listed particles (index[t] = particle, columns [m][3]): the post-pass reads the motion only of
// particles that occur in candidate pairs -- the ones tree_download_compact lists while the capture is on.
int gplum_b200_tree_set_motion_sparse(int m, const int *index, const double *vel, const double *acc_d, const long long *id)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_set_motion_sparse: no GPU-built tree in the selected slot");
    if (m < 0 || (m > 0 && !index)) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion_sparse: bad argument");
    if (m == 0) return 0;
    WalkSet &ws = g.slots[g.cur];
    const int n = (int)ws.n_epi;
    if (m > n) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion_sparse: m = %d, the tree holds %d particles", m, n);
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    const size_t M = (size_t)m;
    if (g.tree_motion.cap < M * 60 + 16) {
        CU(cudaStreamSynchronize(st));
        if (int r = g.tree_motion.reserve(M * 60 + 16)) return r;
    }
    if (!g.tree_inv_valid) {
        if (int r = g.tree_inv.reserve((size_t)n * 4)) return r;
        invert_order_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, gbt::tree_sorted_to_original(), (int *)g.tree_inv.p);
        CU(cudaGetLastError());
        g.launches++;
        g.tree_inv_valid = true;
    }
    double *d = (double *)g.tree_motion.p;
    long long *d_id = (long long *)(d + 6 * M);
    int *d_index = (int *)(d + 7 * M);
    if (int r = upload_host(d_index, index, M * 4, st)) return r;
    if (vel) if (int r = upload_host(d, vel, M * 24, st)) return r;
    if (acc_d) if (int r = upload_host(d + 3 * M, acc_d, M * 24, st)) return r;
    if (id) if (int r = upload_host(d_id, id, M * 8, st)) return r;
    set_motion_sparse_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, n, d_index, (const int *)g.tree_inv.p, vel ? d : nullptr, acc_d ? d + 3 * M : nullptr,
                                                              id ? d_id : nullptr, (EpjAos *)g.jset.epj_aos.p);
    CU(cudaGetLastError());
    g.launches++;
    return 0;
}

// The same for a caller that holds whole columns: vel_all / acc_d_all are [n][3] in particle order; the library
// gathers the m listed particles (OpenMP) into pinned staging and sends only those.
int gplum_b200_tree_set_motion_gather(int m, const int *index, const double *vel_all, const double *acc_d_all, const long long *id_all)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_set_motion_gather: no GPU-built tree in the selected slot");
    if (m < 0 || (m > 0 && !index)) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion_gather: bad argument");
    if (m == 0) return 0;
    const int n = (int)g.slots[g.cur].n_epi;
    if (m > n) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion_gather: m = %d, the tree holds %d particles", m, n);
    const size_t M = (size_t)m;
    CU(cudaSetDevice(g.device));
    CU(cudaStreamSynchronize(g.stream));                      // the staging buffer of the previous call is free
    if (int r = g.motion_pin.reserve(M * 56)) return r;
    double *hv = (double *)g.motion_pin.p, *ha = hv + 3 * M;
    long long *hi = (long long *)(hv + 6 * M);
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (long long t = 0; t < (long long)M; t++) {
        const int i = index[t];
        if (i < 0 || i >= n) { bad = true; continue; }
        for (int d = 0; d < 3; d++) {
            if (vel_all) hv[3 * t + d] = vel_all[3 * (size_t)i + d];
            if (acc_d_all) ha[3 * t + d] = acc_d_all[3 * (size_t)i + d];
        }
        if (id_all) hi[t] = id_all[i];
    }
    if (bad) return fail(GPLUM_B200_ERR_ARG, "tree_set_motion_gather: an index is outside [0, %d)", n);
    return gplum_b200_tree_set_motion_sparse(m, index, vel_all ? hv : nullptr, acc_d_all ? ha : nullptr, id_all ? hi : nullptr);
}

// Page-locked host memory for a caller's staging columns (include/gravity_tree_b200.hpp keeps its columns in it, so
// that uploads and downloads go straight over PCIe instead of through the library's pinned chunks)
void *gplum_b200_pinned_alloc(size_t bytes)
{
    if (ensure_init()) return nullptr;
    void *p = nullptr;
    if (cudaSetDevice(g.device) != cudaSuccess || cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(GPLUM_B200_ERR_CUDA, "pinned_alloc of %zu bytes failed", bytes);
        return nullptr;
    }
    return p;
}
void gplum_b200_pinned_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// The same results with half the bytes on the wire: {acc, phi} of every particle (16 B) and the neighbour words
// {number, rank, id_max, id_min} only of the particles that have candidates (9 % of the N = 1e6 disk); the caller
// fills in ForceGrav::clear()'s values (what gplum_b200_force_clear writes) for the rest.
int gplum_b200_tree_download_compact(float *accphi_out, int *nb_index_out, int *nb_out, int cap, int *n_nb_out)
{
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_download_compact: no GPU-built tree in the selected slot");
    if (!accphi_out || !nb_index_out || !nb_out || !n_nb_out || cap < 0) return fail(GPLUM_B200_ERR_ARG, "tree_download_compact: bad argument");
    CU(cudaSetDevice(g.device));
    WalkSet &ws = g.slots[g.cur];
    cudaStream_t st = g.stream;
    const int n = (int)ws.n_epi;
    if (int r = ws.force_org.reserve((size_t)n * sizeof(ForceAos))) return r;          // [0, 16 n): accphi; [16 n, 32 n): neighbour words
    if (int r = ws.org_index.reserve((size_t)n * 4 + 16)) return r;                    // particle indices, then the counter
    if (int r = g.h_small.reserve(64)) return r;
    uint4 *d_acc = (uint4 *)ws.force_org.p, *d_nb = d_acc + n;
    int *d_idx = (int *)ws.org_index.p, *d_cnt = d_idx + n;
    CU(cudaMemsetAsync(d_cnt, 0, 4, st));
    // with the candidate capture on, the list also names every particle that occurs in a captured pair (a pair that
    // passes the FP32 test from one side only leaves its other end at number = 0): these are the particles whose
    // velocities the post-pass needs (tree_set_motion_sparse)
    unsigned char *need = nullptr;
    if (g.corr_on && ws.captured) {
        if (int r = ws.need.reserve((size_t)n)) return r;
        need = (unsigned char *)ws.need.p;
        CU(cudaMemsetAsync(need, 0, (size_t)n, st));
        mark_pairs_kernel<<<148, 256, 0, st>>>((const int2 *)ws.pairs.p, (const unsigned int *)ws.corr_meta.p, ws.pair_cap, need);
        g.launches++;
    }
    unsort_compact_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const uint4 *)ws.force.p, gbt::tree_sorted_to_original(), d_acc, d_cnt, d_idx, d_nb, n, need);
    CU(cudaGetLastError());
    g.launches++;
    if (!g.ev_small) CU(cudaEventCreateWithFlags(&g.ev_small, cudaEventDisableTiming));
    int *h_cnt = (int *)g.h_small.p;
    CU(cudaMemcpyAsync(h_cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(g.ev_small, st));
    cudaPointerAttributes at;
    const cudaError_t pe = cudaPointerGetAttributes(&at, accphi_out);
    if (pe != cudaSuccess) cudaGetLastError();
    const bool pinned = pe == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
    if (pinned) CU(cudaMemcpyAsync(accphi_out, d_acc, (size_t)n * 16, cudaMemcpyDeviceToHost, st));     // the count arrives while this is on the wire
    else if (int r = download_host(accphi_out, d_acc, (size_t)n * 16, st)) return r;
    CU(cudaEventSynchronize(g.ev_small));
    const int cnt = *h_cnt;
    *n_nb_out = cnt;
    if (cnt > cap) return fail(GPLUM_B200_ERR_OVERFLOW, "tree_download_compact: %d particles have neighbour candidates, room for %d", cnt, cap);
    if (int r = download_host(nb_index_out, d_idx, (size_t)cnt * 4, st)) return r;
    if (int r = download_host(nb_out, d_nb, (size_t)cnt * 16, st)) return r;
    CU(cudaStreamSynchronize(st));
    return 0;
}

int gplum_b200_tree_gpu_stamps(unsigned long long *ns_out, int cap)
{
    if (!ns_out || cap <= 0 || !g.ready || !g.tree_built) return 0;
    return gbt::tree_stamps(ns_out, cap);
}

int gplum_b200_tree_gpu_times(float *ms6)
{
    if (!ms6) return fail(GPLUM_B200_ERR_ARG, "tree_gpu_times: NULL");
    if (!g.ready || !g.tree_built) return fail(GPLUM_B200_ERR_STATE, "tree_gpu_times: no GPU-built tree");
    CU(cudaSetDevice(g.device));
    CU(cudaStreamSynchronize(g.stream));
    gbt::tree_phase_ms(ms6);
    return 0;
}

}  // extern "C"

// ---- device-resident particle state: kick + Kepler drift of isolated particles (iso_step.cu) ----
extern "C" {

int gplum_b200_state_upload(int n, const void *epj, const double *time, const double *dt)
{
    if (n <= 0 || !epj) return fail(GPLUM_B200_ERR_ARG, "state_upload: bad argument");
    if (int r = ensure_init()) return r;
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    const size_t N = (size_t)n;
    if (int r = g.st_epj.reserve(N * sizeof(EpjAos))) return r;
    for (DevBuf *b : {&g.st_time, &g.st_dt, &g.st_acc0}) if (int r = b->reserve(N * 8)) return r;
    for (DevBuf *b : {&g.st_iso, &g.st_handled}) if (int r = b->reserve(N * 4)) return r;
    if (int r = g.st_star.reserve(N * sizeof(gplum_b200_star))) return r;
    if (int r = g.st_cnt.reserve(16)) return r;
    g.st_n = n;
    CU(cudaMemcpyAsync(g.st_epj.p, epj, N * sizeof(EpjAos), cudaMemcpyHostToDevice, st));
    if (time) CU(cudaMemcpyAsync(g.st_time.p, time, N * 8, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(g.st_time.p, 0, N * 8, st));
    if (dt) CU(cudaMemcpyAsync(g.st_dt.p, dt, N * 8, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(g.st_dt.p, 0, N * 8, st));
    CU(cudaMemsetAsync(g.st_acc0.p, 0, N * 8, st));
    CU(cudaMemsetAsync(g.st_iso.p, 0, N * 4, st));
    CU(cudaMemsetAsync(g.st_handled.p, 0, N * 4, st));
    CU(cudaMemsetAsync(g.st_star.p, 0, N * sizeof(gplum_b200_star), st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int gplum_b200_state_download(void *epj_out, double *time, double *dt, void *star_out, int *handled_out)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_download: no resident state");
    CU(cudaSetDevice(g.device));
    CU(cudaStreamSynchronize(g.stream));
    const size_t N = (size_t)g.st_n;
    const cudaMemcpyKind D2H = cudaMemcpyDeviceToHost;
    if (epj_out) CU(cudaMemcpy(epj_out, g.st_epj.p, N * sizeof(EpjAos), D2H));
    if (time) CU(cudaMemcpy(time, g.st_time.p, N * 8, D2H));
    if (dt) CU(cudaMemcpy(dt, g.st_dt.p, N * 8, D2H));
    if (star_out) CU(cudaMemcpy(star_out, g.st_star.p, N * sizeof(gplum_b200_star), D2H));
    if (handled_out) CU(cudaMemcpy(handled_out, g.st_handled.p, N * 4, D2H));
    return 0;
}

int gplum_b200_state_tree_build(double theta, int n_leaf_limit, int n_group_limit, long long *sizes)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_tree_build: no resident state");
    return gplum_b200_tree_build_gpu_epj(g.st_n, g.st_epj.p, 1, theta, n_leaf_limit, n_group_limit, sizes);
}

int gplum_b200_state_kick(int slot, int use_corr, double dt_tree)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_kick: no resident state");
    if (slot < 0 || slot >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "state_kick(slot=%d)", slot);
    WalkSet &ws = g.slots[slot];
    if (ws.n_epi != g.st_n) return fail(GPLUM_B200_ERR_STATE, "state_kick: walk set holds %lld i-particles, the state %d", ws.n_epi, g.st_n);
    if (use_corr && !ws.corrected) return fail(GPLUM_B200_ERR_STATE, "state_kick: no correction in slot %d", slot);
    CU(cudaSetDevice(g.device));
    if (use_corr && !ws.corr_checked) if (int r = corr_status_check(ws)) return r;    // incomplete corrections must not be kicked in
    const int e = gbi::iso_kick(g.st_n, g.st_epj.p, ws.epi.p, ws.force.p, use_corr ? ws.corr_out.p : nullptr, 0.5 * dt_tree, g.stream);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "iso_kick -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches++;
    return 0;
}

int gplum_b200_state_drift(const gplum_b200_iso_params *prm, double t0, double t1, int slot,
                           const int *isolated, const double *acc0)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_drift: no resident state");
    if (!prm) return fail(GPLUM_B200_ERR_ARG, "state_drift: NULL parameters");
    // the reference's step-halving loop (src/hermite.h, `while (fmod(time, dt) != 0) dt *= 0.5`) never ends on
    // these inputs; one host thread would hang there, here the whole stream would
    if (!(prm->dt_tree > 0.0) || !std::isfinite(prm->dt_tree) || !std::isfinite(t0) || !std::isfinite(t1))
        return fail(GPLUM_B200_ERR_ARG, "state_drift: dt_tree = %g, t0 = %g, t1 = %g (need dt_tree > 0 and finite times)", prm->dt_tree, t0, t1);
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    const size_t N = (size_t)g.st_n;
    if (isolated) {
        CU(cudaMemcpyAsync(g.st_iso.p, isolated, N * 4, cudaMemcpyHostToDevice, st));
        if (acc0) CU(cudaMemcpyAsync(g.st_acc0.p, acc0, N * 8, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));           // the caller's arrays may be pageable and short-lived
    } else {
        if (slot < 0 || slot >= N_TAG) return fail(GPLUM_B200_ERR_ARG, "state_drift(slot=%d)", slot);
        WalkSet &ws = g.slots[slot];
        if (!ws.corrected || ws.n_epi != g.st_n) return fail(GPLUM_B200_ERR_STATE, "state_drift: slot %d holds no correction of this state", slot);
        if (!ws.corr_checked) if (int r = corr_status_check(ws)) return r;
        const int e = gbi::iso_flags_from_corr(g.st_n, ws.corr_out.p, (int *)g.st_iso.p, (double *)g.st_acc0.p, st);
        if (e) return fail(GPLUM_B200_ERR_CUDA, "iso_flags -> %s", cudaGetErrorString((cudaError_t)e));
        g.launches++;
    }
    const int e = gbi::iso_drift(g.st_n, g.st_epj.p, (double *)g.st_time.p, (double *)g.st_dt.p, (const double *)g.st_acc0.p,
                                 (const int *)g.st_iso.p, t0, t1, *prm, g.st_star.p, (int *)g.st_handled.p, st);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "iso_drift -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches++;
    return 0;
}

int gplum_b200_state_pull_unhandled(void *rec_out, int *idx_out, int cap, int *n_out)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_pull: no resident state");
    if (!rec_out || !idx_out || !n_out || cap < 0) return fail(GPLUM_B200_ERR_ARG, "state_pull: bad argument");
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    if (int r = g.st_rec.reserve((size_t)std::max(cap, 1) * sizeof(EpjAos))) return r;
    if (int r = g.st_idx.reserve((size_t)std::max(cap, 1) * 4)) return r;
    if (int r = g.st_pin.reserve(16)) return r;
    CU(cudaMemsetAsync(g.st_cnt.p, 0, 4, st));
    const int e = gbi::iso_pull_unhandled(g.st_n, g.st_epj.p, (const int *)g.st_handled.p, g.st_rec.p, (int *)g.st_idx.p, (int *)g.st_cnt.p, cap, st);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "iso_pull -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches++;
    int *h_cnt = (int *)g.st_pin.p;
    CU(cudaMemcpyAsync(h_cnt, g.st_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = *h_cnt;
    if (*h_cnt > cap) return fail(GPLUM_B200_ERR_OVERFLOW, "state_pull: %d particles need the host, buffers hold %d", *h_cnt, cap);
    if (*h_cnt > 0) {
        CU(cudaMemcpyAsync(rec_out, g.st_rec.p, (size_t)*h_cnt * sizeof(EpjAos), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(idx_out, g.st_idx.p, (size_t)*h_cnt * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

int gplum_b200_state_push(const void *rec, const int *idx, int n_rec)
{
    if (!g.ready || g.st_n <= 0) return fail(GPLUM_B200_ERR_STATE, "state_push: no resident state");
    if (n_rec < 0 || (n_rec > 0 && (!rec || !idx))) return fail(GPLUM_B200_ERR_ARG, "state_push: bad argument");
    if (n_rec == 0) return 0;
    CU(cudaSetDevice(g.device));
    cudaStream_t st = g.stream;
    if (int r = g.st_rec.reserve((size_t)n_rec * sizeof(EpjAos))) return r;
    if (int r = g.st_idx.reserve((size_t)n_rec * 4)) return r;
    CU(cudaMemcpyAsync(g.st_rec.p, rec, (size_t)n_rec * sizeof(EpjAos), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g.st_idx.p, idx, (size_t)n_rec * 4, cudaMemcpyHostToDevice, st));
    const int e = gbi::iso_push(n_rec, g.st_rec.p, (const int *)g.st_idx.p, g.st_epj.p, st);
    if (e) return fail(GPLUM_B200_ERR_CUDA, "iso_push -> %s", cudaGetErrorString((cudaError_t)e));
    g.launches++;
    return 0;
}

}  // extern "C"

// ---- host work-list builder, exposed for tests (no device needed) ----
extern "C" int gplum_b200_debug_trace(int on, unsigned long long *out, int cap_items, int *n_items_out)
{
    if (int r = ensure_init()) return r;
    std::lock_guard<std::mutex> lk(g.mu);
    if (out) {
        if (!n_items_out) return fail(GPLUM_B200_ERR_ARG, "debug_trace: n_items_out is null");
        if (g.trace_items > cap_items) return fail(GPLUM_B200_ERR_ARG, "debug_trace: %d items, room for %d", g.trace_items, cap_items);
        CU(cudaStreamSynchronize(g.stream));
        if (g.trace_items > 0) CU(cudaMemcpy(out, g.trace.p, (size_t)g.trace_items * 32, cudaMemcpyDeviceToHost));
        *n_items_out = g.trace_items;
    }
    g.trace_on = on != 0;
    return 0;
}

extern "C" int gplum_b200_debug_build_items(int n_walk, const int *ni, const int *n_epj, const int *n_spj,
                                            long long warp_slots, int tile_cap, int jsplit, int split_m,
                                            int *items_out, int cap_items, int *n_items_out, int *n_slots_out, int *n_groups_out,
                                            int *seg_off_out, int cap_seg, int *n_seg_out)
{
    if (n_walk < 0 || !ni || !n_epj || !n_spj || !n_items_out) return fail(GPLUM_B200_ERR_ARG, "debug_build_items: bad argument");
    const long long ws0 = g.warp_slots; const int tc0 = g.tile_cap, js0 = g.jsplit, sm0 = g.split_m;
    g.warp_slots = warp_slots; g.tile_cap = tile_cap; g.jsplit = jsplit; g.split_m = split_m;
    ItemList il;
    build_items(n_walk, ni, n_epj, n_spj, il);
    g.warp_slots = ws0; g.tile_cap = tc0; g.jsplit = js0; g.split_m = sm0;
    *n_items_out = (int)il.items.size();
    if (n_slots_out) *n_slots_out = il.n_slots;
    if (n_groups_out) *n_groups_out = il.n_groups;
    const int n_seg = il.seg_off.empty() ? 0 : (int)il.seg_off.size() - 1;
    if (n_seg_out) *n_seg_out = n_seg;
    if (items_out) {
        if ((int)il.items.size() > cap_items) return fail(GPLUM_B200_ERR_ARG, "debug_build_items: %zu items, room for %d", il.items.size(), cap_items);
        memcpy(items_out, il.items.data(), il.items.size() * sizeof(WorkItem));
    }
    if (seg_off_out && n_seg > 0) {
        if (n_seg + 1 > cap_seg) return fail(GPLUM_B200_ERR_ARG, "debug_build_items: %d segments, room for %d", n_seg, cap_seg - 1);
        memcpy(seg_off_out, il.seg_off.data(), il.seg_off.size() * sizeof(int));
    }
    return 0;
}
