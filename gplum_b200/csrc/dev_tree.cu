// dev_tree.cu -- interaction-list construction on the GPU: what FDPS does between
// setParticleLocalTree and calcForce, single rank, open boundary, SEARCH_MODE_LONG_SYMMETRY.
//
//   Morton keys + sort      FDPS/src/tree_for_force_impl.hpp:280-330 (setParticleLocalTree, mortonSortLocalTreeOnly)
//   cells (n_leaf_limit)    FDPS/src/tree_for_force_utils.hpp:289 (LinkCell)
//   moments, in/out boxes   FDPS/src/tree_for_force_utils_moment.hpp:6,189; tree.hpp:576-641,1186-1205
//   i-groups (n_group_limit) tree_for_force_utils.hpp:619-650 (MakeIPGroup)
//   per-group walk          FDPS/src/tree_walk.hpp:545-583,706-785 (symmetric search: a cell is opened if the
//                           group's inner box overlaps its outer box, or the group's outer box overlaps its
//                           inner box, or dist^2(group inner box, cell com) <= (size/theta)^2)
//
// Same semantics, same cell numbering and bit-identical FP64 results as the host builder
// (let_tree.cpp): this file is compiled with -fmad=false and every FP64 expression keeps the
// host's evaluation order, so the lists the two produce are equal as sets and the SPJ records are
// equal bit for bit (tests/test_tree_gpu.py).  Within a list the order differs (the host walks
// depth-first with a scalar stack, a warp here expands four cells = 32 children per step).
//
// Layout in HBM: cells are SoA-of-records -- int4 {first, n, child, level}, moments as
// MySPJQuadrupole-shaped 80 B records {mass, com[3], quad[6]} (so the SPJ array of the force pass is
// a plain copy), boxes as 12 doubles {in.lo, in.hi, out.lo, out.hi}.  Children of a cell are 8
// consecutive records; cells of one level are contiguous (breadth-first numbering), which is what
// makes the level-by-level build and the bottom-up moment sweep coalesced.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/block/block_reduce.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <stdint.h>

#include <algorithm>
#include <utility>

#include "dev_tree.h"
#include "items.h"
#include "records.h"

namespace gbt {

using gb::EpiAos;
using gb::EpjAos;
using gb::SpjMonoAos;
using gb::SpjQuadAos;
using gb::EpjPacked;
using gb::BaseItem;
using gb::WorkItem;

namespace {

constexpr int MAX_LEVEL = 42;              // TREE_LEVEL_LIMIT of FDPS's 128-bit key (FDPS/src/ps_defs.hpp:206)
constexpr int LEVEL_HI = 21;               // levels 1..21 live in the sorted 63-bit key word, deeper ones are re-derived
constexpr int N_LVL = MAX_LEVEL + 2;       // lvl_start[l] .. lvl_start[l+1] = cells of level l, l = 0..42
constexpr int NB = 444;                    // blocks of the cooperative tree kernel (3 per SM on 148 SMs)
constexpr int TPB = 256;
constexpr int WALK_WPB = 8;                // warps per block of the walk kernels
constexpr int STK = 1024;                  // per-warp stack of open cells (depth-first in steps of 4: <= ~28 x levels)
constexpr unsigned FULL = 0xffffffffu;

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// device-side scalars of a build
struct Meta {
    double cen[3], hlen, nf, len;          // root cube: centre, half edge, (1 / edge) * 2^42, edge (FDPS/src/key.hpp:160-165)
    int lvl_start[N_LVL + 1];
    int fcount[N_LVL];                     // cells of level l that split
    int overflow;
    int n_walk;
    int w0, w1, e0, e1;                    // this rank's share: walks [w0, w1) = particles [e0, e1) in tree order
    int cap, n_items;
    long long n_adr_epj, n_adr_spj, n_int_epep, n_int_epsp;
    unsigned long long stamp[96];          // globaltimer at the phase boundaries of tree_coop_kernel (block 0)
    int n_stamp;
};

struct State {
    Buf keys_a, keys_b, idx_a, idx_b, cub_temp;
    Buf bbox_part, meta, blk_cnt;
    Buf c_meta, c_mom, c_box, fr_a, fr_b, grp_at, walk_cell;
    Buf w_epi_off, w_ni, w_ne, w_ns, w_ed, w_sd, w_nitems, w_ioff;
    Buf item_key_a, item_key_b, item_a, item_b, item_cum, item_seg;
    Meta *h_meta = nullptr;                // pinned
    Meta *h_meta_stamps = nullptr;         // = h_meta once a build has completed
    int coop_blocks = 0, cell_cap = 0, n = 0, n_walk = 0, n_cells = 0, n_levels = 0, lvl_start[N_LVL + 1] = {};
    cudaEvent_t ev[7] = {};
    bool ev_ok = false, timed = false;
} S;

#define CK(call)                                  \
    do {                                          \
        cudaError_t e_ = (call);                  \
        if (e_ != cudaSuccess) return (int)e_;    \
    } while (0)

struct KP {                                // pointers every tree kernel sees
    int n, n_leaf, n_group, cell_cap;
    double theta;
    const EpjAos *epj;                     // sorted
    const uint64_t *key;                   // sorted
    int4 *c_meta; double *c_mom; double *c_box;
    int *grp_at;
    Meta *meta;
};

__device__ __forceinline__ uint64_t spread3(uint64_t x)
{
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

// 42-bit grid coordinates of a position (FDPS/src/key.hpp:166-181): (U64)((pos - centre + half) * nfactor), clamped
__device__ __forceinline__ void grid_coords(const double *pos, const Meta *m, uint64_t c[3])
{
    const uint64_t nmax = (1ULL << MAX_LEVEL) - 1;
    for (int k = 0; k < 3; k++) {
        const double f = (pos[k] - m->cen[k] + m->hlen) * m->nf;
        uint64_t v = f < 0.0 ? 0 : (f >= 9.2e18 ? nmax : (uint64_t)f);
        c[k] = v > nmax ? nmax : v;
    }
}
// the 21 levels below the sorted key word (FDPS's KeyT::lo_)
__device__ __forceinline__ uint64_t key_lo(const uint64_t c[3])
{
    return spread3(c[0] & 0x1fffff) << 2 | spread3(c[1] & 0x1fffff) << 1 | spread3(c[2] & 0x1fffff);
}

// ---- root cube: bbox of the POSITIONS (FDPS/src/tree_for_force_impl.hpp:770-868; GetMyRSearch yields 0 for EPJGrav) ----
// positions of the unsorted particles: EPJGrav records (stride 14 doubles, pos at +1) or a packed [n][3] column
struct PosView {
    const double *base; int stride;
    __device__ __forceinline__ const double *at(int i) const { return base + (size_t)i * stride; }
};
__global__ void __launch_bounds__(TPB) bbox_kernel(PosView p, int n, double *__restrict__ part)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        const double *x = p.at(i);
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], x[k]);
            hi[k] = fmax(hi[k], x[k]);
        }
    }
    __shared__ double sm[TPB / 32][6];
    for (int k = 0; k < 3; k++)
        for (int d = 16; d > 0; d >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(FULL, lo[k], d));
            hi[k] = fmax(hi[k], __shfl_xor_sync(FULL, hi[k], d));
        }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; k++) { sm[threadIdx.x >> 5][k] = lo[k]; sm[threadIdx.x >> 5][3 + k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < TPB / 32; w++) v = threadIdx.x < 3 ? fmin(v, sm[w][threadIdx.x]) : fmax(v, sm[w][threadIdx.x]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void bbox_final_kernel(const double *__restrict__ part, int n_part, KP P)
{
    __shared__ double sm[6];
    {   // one warp: lanes stride over the blocks' partial boxes, then a shuffle reduction (min / max: any order)
        double v[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
        for (int b = threadIdx.x; b < n_part; b += 32)
            for (int k = 0; k < 6; k++) v[k] = k < 3 ? fmin(v[k], part[b * 6 + k]) : fmax(v[k], part[b * 6 + k]);
        for (int k = 0; k < 6; k++)
            for (int d = 16; d > 0; d >>= 1) {
                const double o = __shfl_xor_sync(FULL, v[k], d);
                v[k] = k < 3 ? fmin(v[k], o) : fmax(v[k], o);
            }
        if (threadIdx.x == 0) for (int k = 0; k < 6; k++) sm[k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *lo = sm, *hi = sm + 3;
        double length = 0, cen[3], len_dim[3];
        for (int k = 0; k < 3; k++) { cen[k] = (hi[k] + lo[k]) * 0.5; len_dim[k] = hi[k] - lo[k]; length = fmax(length, len_dim[k]); }
        // a dimension much thinner than the cube (a disk's z) is pushed wholly into one half
        for (int k = 0; k < 3; k++) if (len_dim[k] < 0.1 * length) cen[k] -= len_dim[k] * 0.51;
        length *= 1.000001;
        if (!(length > 0)) length = 1.0;
        const double hlen = length * 0.5;
        Meta *m = P.meta;
        for (int k = 0; k < 3; k++) m->cen[k] = cen[k];
        m->hlen = hlen;
        m->len = hlen * 2.0;
        m->nf = (1.0 / (hlen * 2.0)) * (double)(1ULL << MAX_LEVEL);
        // the root cell, FDPS's seven unused cells behind it (LinkCell: tc_array[1..7]) and the level table
        for (int l = 0; l <= N_LVL; l++) m->lvl_start[l] = l == 0 ? 0 : 8;
        for (int c = 1; c < 8; c++) P.c_meta[c] = make_int4(0, 0, -1, 0);
        for (int l = 0; l < N_LVL; l++) m->fcount[l] = 0;
        const bool root_splits = P.n > P.n_leaf;
        m->fcount[0] = root_splits ? 1 : 0;
        m->overflow = 0;
        P.c_meta[0] = make_int4(0, P.n, -1, 0);
        if (P.n <= P.n_group || !root_splits) P.grp_at[0] = 0;
    }
}

__global__ void __launch_bounds__(TPB) key_kernel(PosView p, int n, const Meta *__restrict__ m,
                                                  uint64_t *__restrict__ key, int *__restrict__ idx)
{
    const int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    uint64_t c[3];
    grid_coords(p.at(i), m, c);
    key[i] = spread3(c[0] >> LEVEL_HI) << 2 | spread3(c[1] >> LEVEL_HI) << 1 | spread3(c[2] >> LEVEL_HI);   // KeyT::hi_
    idx[i] = i;
}

// FDPS sorts by (hi, lo); the radix sort above orders by hi only (stable: ties in particle order).  Particles that
// share all 21 upper levels are closer than 2^-21 of the root edge -- rare --, so the runs of equal hi are put
// into (lo, particle index) order here, one thread per run.
__global__ void __launch_bounds__(TPB) tie_fix_kernel(PosView raw, int n, const Meta *__restrict__ m,
                                                      const uint64_t *__restrict__ key, int *__restrict__ idx)
{
    const int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n - 1) return;
    const uint64_t k = key[i];
    if (key[i + 1] != k || (i > 0 && key[i - 1] == k)) return;      // not the first element of a run of >= 2
    int e = i + 2;
    while (e < n && key[e] == k) e++;
    for (int a = i + 1; a < e; a++) {                                // insertion sort of idx[i..e) by (lo, index)
        const int va = idx[a];
        uint64_t c[3];
        grid_coords(raw.at(va), m, c);
        const uint64_t la = key_lo(c);
        int b = a - 1;
        while (b >= i) {
            const int vb = idx[b];
            grid_coords(raw.at(vb), m, c);
            const uint64_t lb = key_lo(c);
            if (lb < la || (lb == la && vb < va)) break;
            idx[b + 1] = vb;
            b--;
        }
        idx[b + 1] = va;
    }
}

// sorted EPJ (112 B records moved as 7 x 16 B) and the EPI of the same particle
__global__ void __launch_bounds__(TPB) gather_kernel(const uint4 *__restrict__ in, const int *__restrict__ idx, int n,
                                                     uint4 *__restrict__ epj, EpiAos *__restrict__ epi, EpjPacked *__restrict__ packed)
{
    const int t = blockIdx.x * TPB + threadIdx.x;
    const int i = t >> 3, q = t & 7;              // 8 lanes per record, 7 of them move 16 B
    if (i >= n) return;
    const int src = idx[i];
    if (q == 7) {                                 // the eighth writes the force kernel's packed record (kernels.cuh)
        if (packed) {
            const EpjAos *r = reinterpret_cast<const EpjAos *>(in) + src;
            packed[i] = gb::epj_pack(r->pos, r->mass, r->r_out, r->r_search, r->id_local, r->myrank);
        }
        return;
    }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < 7) {
        v = __ldg(in + (size_t)src * 7 + q);
        epj[(size_t)i * 7 + q] = v;
    }
    // EPIGrav = the first 48 B of EPJGrav (src/particle.h:93-109,149-156)
    if (q < 3) reinterpret_cast<uint4 *>(epi)[(size_t)i * 3 + q] = v;
}

// sorted EPJ / EPI from columns: 8 lanes per particle, lane q < 7 writes bytes [16 q, 16 q + 16) of the EPJGrav record
// (src/particle.h:149-156: id_local, myrank | pos | r_out, r_search | id, mass | vel | acc_d)
__global__ void __launch_bounds__(TPB) gather_soa_kernel(const double *__restrict__ pos, const double *__restrict__ mass,
                                                         const double *__restrict__ r_out, const double *__restrict__ r_search,
                                                         const double *__restrict__ vel, int pos_stride, int col_stride, int rank,
                                                         const int *__restrict__ idx, int n, uint4 *__restrict__ epj, EpiAos *__restrict__ epi,
                                                         EpjPacked *__restrict__ packed)
{
    const int t = blockIdx.x * TPB + threadIdx.x;
    const int i = t >> 3, q = t & 7;
    if (i >= n) return;
    const int src = idx[i];
    const size_t s3 = 3 * (size_t)src;
    const size_t sp = (size_t)pos_stride * src, sc = (size_t)col_stride * src;
    if (q == 7) {                                 // the force kernel's packed record
        if (packed) {
            const double x[3] = {pos[sp], pos[sp + 1], pos[sp + 2]};
            packed[i] = gb::epj_pack(x, mass[sc], r_out[sc], r_search[sc], src, rank);
        }
        return;
    }
    union { uint4 v; double d[2]; int w[4]; long long l[2]; } u;
    u.v = make_uint4(0, 0, 0, 0);
    switch (q) {
        case 0: u.w[0] = src; u.w[1] = rank; u.d[1] = pos[sp]; break;
        case 1: u.d[0] = pos[sp + 1]; u.d[1] = pos[sp + 2]; break;
        case 2: u.d[0] = r_out[sc]; u.d[1] = r_search[sc]; break;
        case 3: u.l[0] = src; u.d[1] = mass[sc]; break;
        case 4: if (vel) { u.d[0] = vel[s3]; u.d[1] = vel[s3 + 1]; } break;
        case 5: if (vel) u.d[0] = vel[s3 + 2]; break;
        default: break;
    }
    epj[(size_t)i * 7 + q] = u.v;
    if (q < 3) reinterpret_cast<uint4 *>(epi)[(size_t)i * 3 + q] = u.v;
}

// ---- cells, one level per launch triple ----
// The frontier of level L = its cells that split (n > n_leaf), in breadth-first order; the children of
// frontier entry f are the 8 cells lvl_start[L+1] + 8 f + o.  8 lanes work on one cell: lane o finds
// the end of octant o by binary search on the key digit of that level.
// (Cells and frontier entries are written by other blocks of the same launch one grid barrier earlier:
// they are read with ld.global.cg, never through L1 or the read-only path.)
__device__ __forceinline__ void child_range(const KP &P, const int *fr, int f, int f1, int o, int L,
                                            int &cell, int &pn, int &lb, int &ub)
{
    int first = 0, n = 0;
    cell = -1;
    if (f < f1) { cell = __ldcg(fr + f); const int4 m = __ldcg(P.c_meta + cell); first = m.x; n = m.y; }
    pn = n;
    const int end = first + n;
    ub = end;
    if (o < 7 && n > 0) {
        int lo = first, hi = end;
        if (L < LEVEL_HI) {                                  // child level L+1 <= 21: a digit of the sorted key word
            const int shift = 3 * (LEVEL_HI - 1 - L);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)((P.key[mid] >> shift) & 7) <= o) lo = mid + 1; else hi = mid;
            }
        } else {                                             // deeper (particles closer than 2^-21 of the root edge):
            const int bit = MAX_LEVEL - 1 - L;               // the digit is re-derived from the position
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                uint64_t c[3];
                grid_coords(P.epj[mid].pos, P.meta, c);
                const int dg = (int)(((c[0] >> bit) & 1) << 2 | ((c[1] >> bit) & 1) << 1 | ((c[2] >> bit) & 1));
                if (dg <= o) lo = mid + 1; else hi = mid;
            }
        }
        ub = lo;
    }
    lb = __shfl_up_sync(FULL, ub, 1, 8);
    if (o == 0) lb = first;
}

__device__ __forceinline__ void chunk_of(int F, int &f0, int &f1)
{
    const int chunk = ((F + (int)gridDim.x - 1) / (int)gridDim.x + 31) & ~31;
    f0 = min(F, (int)blockIdx.x * chunk);
    f1 = min(F, f0 + chunk);
}

// One thread per cell.  A cell's latency is what a level costs (every level ends in a grid barrier), so the
// loads of a cell are issued in batches that do not depend on each other: leaf particles four at a time
// (index clamped, accumulation predicated -- the order of the sums stays the host's), the 8 children of an
// inner cell unconditionally (an empty child holds mass = com = quad = 0 and boxes at +-1e300: adding it is
// exact, so skipping it like the host does and not skipping it give the same bits).
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STAMP() do { if (b == 0 && threadIdx.x == 0 && ns < 96) meta->stamp[ns++] = gtimer(); } while (0)

__device__ __forceinline__ void moment_cell(const KP &P, int c)
{
    const int4 m = __ldcg(P.c_meta + c);
    double mass = 0, com[3] = {0, 0, 0}, q[6] = {0, 0, 0, 0, 0, 0};
    double ilo[3] = {1e300, 1e300, 1e300}, ihi[3] = {-1e300, -1e300, -1e300};
    double olo[3] = {1e300, 1e300, 1e300}, ohi[3] = {-1e300, -1e300, -1e300};
    if (m.y > 0) {
        if (m.z < 0) {
            const int end = m.x + m.y;
            for (int i0 = m.x; i0 < end; i0 += 4) {
                double mi[4], rs[4], x[4][3];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const EpjAos &p = P.epj[min(i0 + u, end - 1)];
                    mi[u] = p.mass; rs[u] = p.r_search;
                    x[u][0] = p.pos[0]; x[u][1] = p.pos[1]; x[u][2] = p.pos[2];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + u < end) {
                        const double r = 1.1 * rs[u];
                        mass += mi[u];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            com[k] += mi[u] * x[u][k];
                            ilo[k] = fmin(ilo[k], x[u][k] - 0.0); ihi[k] = fmax(ihi[k], x[u][k] + 0.0);
                            olo[k] = fmin(olo[k], x[u][k] - r); ohi[k] = fmax(ohi[k], x[u][k] + r);
                        }
                    }
                }
            }
            { const double inv_m = 1.0 / mass;       // PS::F64vec / F64 multiplies by the reciprocal (FDPS/src/vector3.hpp:217-220)
              for (int k = 0; k < 3; k++) com[k] = (mass != 0.0) ? com[k] * inv_m : 0.0; }
            for (int i0 = m.x; i0 < end; i0 += 4) {
                double mi[4], x[4][3];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const EpjAos &p = P.epj[min(i0 + u, end - 1)];
                    mi[u] = p.mass;
                    x[u][0] = p.pos[0]; x[u][1] = p.pos[1]; x[u][2] = p.pos[2];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + u < end) {
                        const double d0 = x[u][0] - com[0], d1 = x[u][1] - com[1], d2 = x[u][2] - com[2];
                        q[0] += mi[u] * d0 * d0; q[1] += mi[u] * d1 * d1; q[2] += mi[u] * d2 * d2;
                        q[3] += mi[u] * d0 * d1; q[4] += mi[u] * d0 * d2; q[5] += mi[u] * d1 * d2;
                    }
                }
            }
        } else {
            // children are 8 consecutive 80 B moment records and 8 consecutive 96 B boxes: 16 B loads
            const double2 *cm0 = reinterpret_cast<const double2 *>(P.c_mom + (size_t)m.z * 10);
            const double2 *cb0 = reinterpret_cast<const double2 *>(P.c_box + (size_t)m.z * 12);
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const double2 a = __ldcg(cm0 + o * 5), b2 = __ldcg(cm0 + o * 5 + 1);       // mass, cx | cy, cz
                const double2 b0 = __ldcg(cb0 + o * 6), b1 = __ldcg(cb0 + o * 6 + 1), b2_ = __ldcg(cb0 + o * 6 + 2),
                              b3 = __ldcg(cb0 + o * 6 + 3), b4 = __ldcg(cb0 + o * 6 + 4), b5 = __ldcg(cb0 + o * 6 + 5);
                const double cmass = a.x;
                mass += cmass;
                com[0] += cmass * a.y; com[1] += cmass * b2.x; com[2] += cmass * b2.y;
                ilo[0] = fmin(ilo[0], b0.x); ilo[1] = fmin(ilo[1], b0.y); ilo[2] = fmin(ilo[2], b1.x);
                ihi[0] = fmax(ihi[0], b1.y); ihi[1] = fmax(ihi[1], b2_.x); ihi[2] = fmax(ihi[2], b2_.y);
                olo[0] = fmin(olo[0], b3.x); olo[1] = fmin(olo[1], b3.y); olo[2] = fmin(olo[2], b4.x);
                ohi[0] = fmax(ohi[0], b4.y); ohi[1] = fmax(ohi[1], b5.x); ohi[2] = fmax(ohi[2], b5.y);
            }
            { const double inv_m = 1.0 / mass;       // PS::F64vec / F64 multiplies by the reciprocal (FDPS/src/vector3.hpp:217-220)
              for (int k = 0; k < 3; k++) com[k] = (mass != 0.0) ? com[k] * inv_m : 0.0; }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const double2 a = __ldcg(cm0 + o * 5), b2 = __ldcg(cm0 + o * 5 + 1), q01 = __ldcg(cm0 + o * 5 + 2),
                              q23 = __ldcg(cm0 + o * 5 + 3), q45 = __ldcg(cm0 + o * 5 + 4);
                const double mi = a.x;
                const double d0 = a.y - com[0], d1 = b2.x - com[1], d2 = b2.y - com[2];
                q[0] += mi * d0 * d0 + q01.x; q[1] += mi * d1 * d1 + q01.y; q[2] += mi * d2 * d2 + q23.x;
                q[3] += mi * d0 * d1 + q23.y; q[4] += mi * d0 * d2 + q45.x; q[5] += mi * d1 * d2 + q45.y;
            }
        }
    }
    double2 *om = reinterpret_cast<double2 *>(P.c_mom + (size_t)c * 10), *ob = reinterpret_cast<double2 *>(P.c_box + (size_t)c * 12);
    om[0] = make_double2(mass, com[0]); om[1] = make_double2(com[1], com[2]);
    om[2] = make_double2(q[0], q[1]); om[3] = make_double2(q[2], q[3]); om[4] = make_double2(q[4], q[5]);
    ob[0] = make_double2(ilo[0], ilo[1]); ob[1] = make_double2(ilo[2], ihi[0]); ob[2] = make_double2(ihi[1], ihi[2]);
    ob[3] = make_double2(olo[0], olo[1]); ob[4] = make_double2(olo[2], ohi[0]); ob[5] = make_double2(ohi[1], ohi[2]);
}

// ---- cells (top-down, one level per pair of grid barriers) and moments + boxes (bottom-up, one level per
// barrier) in ONE cooperative launch: the level loops are latency chains of tiny kernels otherwise (a 1e6
// disk has 12 levels; 3 launches per level and 21 possible levels cost more than the work) ----
__global__ void __launch_bounds__(TPB, 3) tree_coop_kernel(KP P, int *fr_a, int *fr_b, int *blk_cnt)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ int wcnt[TPB / 32];
    __shared__ int s_red[2][TPB / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, o = threadIdx.x & 7;
    const int nb = gridDim.x, b = blockIdx.x;
    Meta *meta = P.meta;
    int *fr = fr_a, *fr_next = fr_b;
    int L = 0, ns = 0;
    STAMP();
    for (; L < MAX_LEVEL; L++) {
        const int F = __ldcg(&meta->fcount[L]);
        if (F == 0) break;                                   // same value in every block
        int f0, f1;
        chunk_of(F, f0, f1);
        // phase A: how many children of my chunk of the frontier split again
        int cnt = 0;
        int cell0 = -1, pn0 = 0, lb0 = 0, ub0 = 0;           // first round's searches, reused in phase B
        for (int base = f0; base < f1; base += TPB / 8) {
            int cell, pn, lb, ub;
            child_range(P, fr, base + (threadIdx.x >> 3), f1, o, L, cell, pn, lb, ub);
            if (base == f0) { cell0 = cell; pn0 = pn; lb0 = lb; ub0 = ub; }
            cnt += (ub - lb > P.n_leaf && L + 1 < MAX_LEVEL) ? 1 : 0;
        }
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        if (lane == 0) wcnt[w] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int k = 0; k < TPB / 32; k++) s += wcnt[k];
            blk_cnt[b] = s;
        }
        grid.sync();
        STAMP();
        // phase B: my offset into the next frontier = counts of the blocks before me
        int before_me = 0, total = 0;
        for (int k = threadIdx.x; k < nb; k += TPB) { const int v = __ldcg(blk_cnt + k); total += v; if (k < b) before_me += v; }
        for (int d = 16; d > 0; d >>= 1) { before_me += __shfl_xor_sync(FULL, before_me, d); total += __shfl_xor_sync(FULL, total, d); }
        if (lane == 0) { s_red[0][w] = before_me; s_red[1][w] = total; }
        __syncthreads();
        before_me = 0; total = 0;
        for (int k = 0; k < TPB / 32; k++) { before_me += s_red[0][k]; total += s_red[1][k]; }
        const int cbase = __ldcg(&meta->lvl_start[L + 1]);
        if (b == 0 && threadIdx.x == 0) {
            const int next_start = cbase + 8 * F;            // first cell of level L+2
            int t = total;
            // this level's children were checked one level up; make sure the next level's children fit
            if ((long long)next_start + 8LL * t > (long long)P.cell_cap) { meta->overflow = 1; t = 0; }
            for (int l = L + 2; l <= N_LVL; l++) meta->lvl_start[l] = next_start;
            meta->fcount[L + 1] = t;
        }
        int running = before_me;
        for (int base = f0; base < f1; base += TPB / 8) {
            const int f = base + (threadIdx.x >> 3);
            int cell = cell0, pn = pn0, lb = lb0, ub = ub0;
            if (base != f0) child_range(P, fr, f, f1, o, L, cell, pn, lb, ub);      // block-uniform branch
            const int cn = ub - lb;
            const bool will = cell >= 0 && cn > P.n_leaf && L + 1 < MAX_LEVEL;
            const unsigned bal = __ballot_sync(FULL, will);
            __syncthreads();                                 // wcnt of the previous round has been read
            if (lane == 0) wcnt[w] = __popc(bal);
            __syncthreads();
            int before = 0, tot = 0;
            for (int k = 0; k < TPB / 32; k++) { const int c = wcnt[k]; if (k < w) before += c; tot += c; }
            if (cell >= 0) {
                const int cidx = cbase + 8 * f + o;
                P.c_meta[cidx] = make_int4(lb, cn, -1, L + 1);
                if (o == 0) reinterpret_cast<int *>(&P.c_meta[cell])[2] = cbase + 8 * f;
                // i-groups: the shallowest cells with <= n_group particles (or leaves)
                if (pn > P.n_group && cn > 0 && (cn <= P.n_group || !will)) P.grp_at[lb] = cidx;
                if (will) fr_next[running + before + __popc(bal & ((1u << lane) - 1))] = cidx;
            }
            running += tot;
        }
        grid.sync();
        STAMP();
        int *t = fr; fr = fr_next; fr_next = t;
    }
    // cells exist on levels 0..L; the barrier that ended the last level also published its cells
    // warps, not threads, are dealt round-robin over the blocks: a small level still uses every SM's
    // load pipe (uncoalesced record loads cost one L1 tag cycle per sector), and a warp's 32 cells stay
    // consecutive (children of consecutive cells are consecutive in memory)
    const int gthreads = nb * TPB;
    const int gtid = (w * nb + b) * 32 + lane;
    for (int l = L; l >= 0; l--) {
        const int c0 = __ldcg(&meta->lvl_start[l]), c1 = __ldcg(&meta->lvl_start[l + 1]);
        for (int c = c0 + gtid; c < c1; c += gthreads) moment_cell(P, c);
        if (l > 0) grid.sync();
        STAMP();
    }
    if (b == 0 && threadIdx.x == 0) meta->n_stamp = ns;
}

// ---- multi-GPU: this rank's share of the walks.  Every rank builds the same tree over all particles; rank r of W
// evaluates the walks whose first particle (tree order) lies in [n r / W, n (r+1) / W) -- contiguous in Morton order,
// i.e. a compact spatial domain, what dinfo.decomposeDomainAll gives an MPI rank of the reference. ----
__global__ void walk_range_kernel(KP P, const int *__restrict__ walk_cell, int n_walk, int part_rank, int part_world)
{
    if (threadIdx.x != 0) return;
    Meta *m = P.meta;
    if (part_world <= 1) { m->w0 = 0; m->w1 = n_walk; m->e0 = n_walk > 0 ? 0 : P.n; m->e1 = P.n; return; }   // no searches (20 us of dependent loads)
    const long long p0 = (long long)P.n * part_rank / part_world, p1 = (long long)P.n * (part_rank + 1) / part_world;
    int b[2];
    for (int q = 0; q < 2; q++) {                   // number of walks whose first particle is below the bound
        const long long bound = q ? p1 : p0;
        int lo = 0, hi = n_walk;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)P.c_meta[walk_cell[mid]].x < bound) lo = mid + 1; else hi = mid; }
        b[q] = lo;
    }
    m->w0 = b[0]; m->w1 = b[1];
    m->e0 = b[0] < n_walk ? P.c_meta[walk_cell[b[0]]].x : P.n;
    m->e1 = b[1] < n_walk ? P.c_meta[walk_cell[b[1]]].x : P.n;
}

// ---- per-group walk: one warp per group, depth-first in steps of 4 cells x 8 children ----
struct WalkP {
    KP P;
    const int *walk_cell; int n_walk;
    int *epi_off, *ni, *n_epj, *n_spj;                     // per walk
    const long long *epj_disp, *spj_disp;                  // FILL
    int *adr_epj, *adr_spj;                                // FILL
};

template <bool FILL>
__global__ void __launch_bounds__(WALK_WPB * 32) walk_kernel(WalkP A)
{
    __shared__ int stk_all[WALK_WPB][STK];
    const KP &P = A.P;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int *stk = stk_all[wib];
    const double inv_theta2 = 1.0 / (P.theta * P.theta);
    const double len = P.meta->len;
    const unsigned lt = (1u << lane) - 1;
    const int w0 = P.meta->w0, w1 = P.meta->w1;
    if (!FILL)                                              // walks of other ranks: empty lists, no work items
        for (int g = blockIdx.x * WALK_WPB * 32 + threadIdx.x; g < A.n_walk; g += gridDim.x * WALK_WPB * 32)
            if (g < w0 || g >= w1) { const int4 gm = P.c_meta[A.walk_cell[g]]; A.epi_off[g] = gm.x; A.ni[g] = gm.y; A.n_epj[g] = 0; A.n_spj[g] = 0; }
    for (int g = w0 + blockIdx.x * WALK_WPB + wib; g < w1; g += gridDim.x * WALK_WPB) {
        const int gc = A.walk_cell[g];
        const int4 gm = P.c_meta[gc];
        // group boxes: lanes 0..11 load one double each, everybody gets all 12
        double bv = lane < 12 ? P.c_box[(size_t)gc * 12 + lane] : 0.0;
        double gb[12];
#pragma unroll
        for (int k = 0; k < 12; k++) gb[k] = __shfl_sync(FULL, bv, k);
        long long ed = 0, sd = 0;
        if (FILL) { ed = A.epj_disp[g]; sd = A.spj_disp[g]; }
        int ne = 0, ns = 0, top = 0;
        const int4 root = P.c_meta[0];
        if (root.z < 0) {                                   // the root is a leaf: every particle
            if (FILL) for (int k = lane; k < root.y; k += 32) A.adr_epj[ed + k] = root.x + k;
            ne = root.y;
        } else {
            if (lane == 0) stk[0] = 0;
            top = 1;
        }
        __syncwarp();
        while (top > 0) {
            const int take = min(top, 4);
            top -= take;
            const int sub = lane >> 3, o = lane & 7;
            int ci = -1;
            int4 cm = make_int4(0, 0, -1, 0);
            if (sub < take) { ci = P.c_meta[stk[top + sub]].z + o; cm = P.c_meta[ci]; }
            __syncwarp();                                   // stack entries read before they are overwritten
            const bool valid = cm.y > 0;
            bool open = false;
            if (valid) {
                const double *mo = P.c_mom + (size_t)ci * 10;
                const double cx[3] = {mo[1], mo[2], mo[3]};
                const double size = ldexp(len, -cm.w);
                double d2 = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) { const double d = fmax(0.0, fmax(gb[k] - cx[k], cx[k] - gb[3 + k])); d2 += d * d; }
                open = d2 <= size * size * inv_theta2;
                if (!open) {
                    const double *cb = P.c_box + (size_t)ci * 12;
                    bool a = true, b = true;              // a: group.in overlaps cell.out; b: group.out overlaps cell.in
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const double cilo = cb[k], cihi = cb[3 + k], colo = cb[6 + k], cohi = cb[9 + k];
                        if (gb[3 + k] < colo || cohi < gb[k]) a = false;
                        if (gb[9 + k] < cilo || cihi < gb[6 + k]) b = false;
                    }
                    open = a || b;
                }
            }
            const bool leaf = cm.z < 0;
            const bool is_sp = valid && !open, is_push = valid && open && !leaf, is_ep = valid && open && leaf;
            const unsigned m_sp = __ballot_sync(FULL, is_sp), m_push = __ballot_sync(FULL, is_push);
            if (FILL && is_sp) A.adr_spj[sd + ns + __popc(m_sp & lt)] = ci;
            ns += __popc(m_sp);
            const int n_push = __popc(m_push);
            if (top + n_push > STK) {                       // cannot happen for sane trees; flagged, never silent
                if (lane == 0) P.meta->overflow = 2;
                top = 0;
                break;
            }
            if (is_push) stk[top + __popc(m_push & lt)] = ci;
            top += n_push;
            const int cnt = is_ep ? cm.y : 0;
            int inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += u; }
            if (FILL && is_ep) {
                int *dst = A.adr_epj + ed + ne + (inc - cnt);
                for (int k = 0; k < cnt; k++) dst[k] = cm.x + k;
            }
            ne += __shfl_sync(FULL, inc, 31);
            __syncwarp();
        }
        if (!FILL && lane == 0) {
            A.epi_off[g] = gm.x; A.ni[g] = gm.y; A.n_epj[g] = ne; A.n_spj[g] = ns;
        }
    }
}

// ---- totals, tile capacity, per-walk item counts ----
struct AsLL { __host__ __device__ long long operator()(const int &v) const { return (long long)v; } };
struct NonNeg { __host__ __device__ bool operator()(const int &v) const { return v >= 0; } };

__global__ void __launch_bounds__(1024) totals_kernel(int n_walk, const int *__restrict__ ni, const int *__restrict__ ne,
                                                      const int *__restrict__ ns, Meta *m, long long warp_slots,
                                                      int tile_cap, int jsplit, int rmax, int split_m)
{
    long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // adr_epj, adr_spj, int_ee, int_es, items at cap 64..4
    for (int w = threadIdx.x + m->w0; w < m->w1; w += 1024) {
        const long long i = ni[w], e = ne[w], s = ns[w];
        v[0] += e; v[1] += s; v[2] += i * e; v[3] += i * s;
        for (int k = 0; k < 5; k++) v[4 + k] += (i + (64 >> k) - 1) / (64 >> k);
    }
    __shared__ long long sm[32][9];
    for (int k = 0; k < 9; k++)
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], d);
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 9; k++) sm[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t[9];
        for (int k = 0; k < 9; k++) { t[k] = 0; for (int w = 0; w < 32; w++) t[k] += sm[w][k]; }
        m->n_adr_epj = t[0]; m->n_adr_spj = t[1]; m->n_int_epep = t[2]; m->n_int_epsp = t[3];
        m->cap = gb::tile_cap_choose(t + 4, warp_slots, tile_cap, jsplit != 0, rmax);
        if (split_m > 0 && rmax == 2 && tile_cap == 0) m->cap = 64;      // a small pass is laid out in segments (items.h)
        m->n_walk = n_walk;
    }
}

__global__ void __launch_bounds__(TPB) item_count_kernel(int n_walk, const int *__restrict__ ni, const Meta *__restrict__ m,
                                                         int jsplit, int *__restrict__ n_items)
{
    const int w = blockIdx.x * TPB + threadIdx.x;
    if (w >= n_walk) return;
    const int cap = m->cap;
    int rem = (w >= m->w0 && w < m->w1) ? ni[w] : 0, cnt = 0;
    while (rem > 0) { int n, shape; gb::tile_next(rem, cap, jsplit != 0, n, shape); rem -= n; cnt++; }
    n_items[w] = cnt;
}

__global__ void item_total_kernel(int n_walk, const int *__restrict__ n_items, const int *__restrict__ ioff, Meta *m)
{
    m->n_items = n_walk > 0 ? ioff[n_walk - 1] + n_items[n_walk - 1] : 0;
}

__device__ __forceinline__ uint64_t cost_key(double c) { return (uint64_t)__double_as_longlong(c); }   // c > 0

__global__ void __launch_bounds__(TPB) item_emit_kernel(int n_walk, const int *__restrict__ ni, const int *__restrict__ ne,
                                                        const int *__restrict__ ns, const int *__restrict__ ioff,
                                                        const Meta *__restrict__ m, int jsplit,
                                                        BaseItem *__restrict__ items, uint64_t *__restrict__ keys)
{
    const int w = blockIdx.x * TPB + threadIdx.x;
    if (w >= n_walk) return;
    const int cap = m->cap;
    int rem = (w >= m->w0 && w < m->w1) ? ni[w] : 0, i0 = 0, k = ioff[w];
    while (rem > 0) {
        int n, shape;
        gb::tile_next(rem, cap, jsplit != 0, n, shape);
        items[k] = BaseItem{w, i0, n, gb::tile_cfg_of(shape)};
        keys[k] = cost_key(gb::tile_cost(ne[w], ns[w], shape));
        rem -= n; i0 += n; k++;
    }
}

// ---- base items (sorted, longest first) -> work items.  A pass with many items: one work item per base item. ----
__global__ void __launch_bounds__(TPB) item_widen_kernel(int n, const BaseItem *__restrict__ base, WorkItem *__restrict__ out)
{
    const int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= n) return;
    const BaseItem b = base[k];
    out[k] = WorkItem{b.walk, b.i0, b.ni, b.cfg, 0, -1, 0, 0};
}

// A pass with less than two waves of items is laid out as one wave of equal segments (items.h: seg_cut_begin /
// seg_cut_next, the same code the host work-list builder runs -> the same work list).  One block: the list is short
// by definition.  out holds n_out_cap items; the ones behind the last part are empty (ni = 0).
constexpr int SPLIT_TPB = 1024;
__global__ void __launch_bounds__(SPLIT_TPB) item_split_kernel(int n, const BaseItem *__restrict__ base, const uint64_t *__restrict__ keys,
                                                               const int *__restrict__ ne, const int *__restrict__ ns,
                                                               long long n_seg, long long *__restrict__ cum, int *__restrict__ seg_of_item,
                                                               WorkItem *__restrict__ out, int n_out_cap, int *__restrict__ seg_off)
{
    typedef cub::BlockScan<long long, SPLIT_TPB> ScanLL;
    typedef cub::BlockScan<int, SPLIT_TPB> Scan;
    __shared__ union { typename ScanLL::TempStorage l; typename Scan::TempStorage s; } tmp;
    __shared__ long long s_runll;
    __shared__ int s_run[3];
    if (threadIdx.x == 0) { s_runll = 0; s_run[0] = s_run[1] = s_run[2] = 0; }
    __syncthreads();
    // pass 1: cost units before every base item, and their total
    for (int b0 = 0; b0 < n; b0 += SPLIT_TPB) {
        const int k = b0 + threadIdx.x;
        const long long c = k < n ? gb::item_cost_units(__longlong_as_double((long long)keys[k])) : 0;
        long long o, t;
        ScanLL(tmp.l).ExclusiveSum(c, o, t);
        if (k < n) cum[k] = o + s_runll;
        __syncthreads();
        if (threadIdx.x == 0) s_runll += t;
        __syncthreads();
    }
    const long long W = s_runll;
    // pass 2: count the parts of every item, scan, emit
    for (int b0 = 0; b0 < n; b0 += SPLIT_TPB) {
        const int k = b0 + threadIdx.x;
        BaseItem b = BaseItem{0, 0, 0, 0};
        int K = 0, e = 0, sp = 0;
        long long C = 0, c = 0;
        if (k < n) {
            b = base[k];
            e = ne[b.walk]; sp = ns[b.walk];
            C = cum[k]; c = gb::item_cost_units(__longlong_as_double((long long)keys[k]));
            gb::SegCut q;
            gb::seg_cut_begin(q, W, n_seg, C, c, b.cfg, e, sp);
            int t0, t1; long long sg;
            while (gb::seg_cut_next(q, t0, t1, sg)) K++;
        }
        int o_item, o_slot, o_group, t_item, t_slot, t_group;
        Scan(tmp.s).ExclusiveSum(K, o_item, t_item);
        __syncthreads();
        Scan(tmp.s).ExclusiveSum(K > 1 ? K : 0, o_slot, t_slot);
        __syncthreads();
        Scan(tmp.s).ExclusiveSum(K > 1 ? 1 : 0, o_group, t_group);
        __syncthreads();
        o_item += s_run[0]; o_slot += s_run[1]; o_group += s_run[2];
        if (k < n) {
            gb::SegCut q;
            gb::seg_cut_begin(q, W, n_seg, C, c, b.cfg, e, sp);
            int t0, t1; long long sg;
            for (int p = 0; gb::seg_cut_next(q, t0, t1, sg); p++) {
                if (o_item + p >= n_out_cap) break;
                const bool whole = K == 1;
                out[o_item + p] = WorkItem{b.walk, b.i0, b.ni, b.cfg | (whole ? 0 : (K << 8) | (p << 16)), whole ? 0 : t0, whole ? -1 : t1,
                                           whole ? 0 : o_slot, whole ? 0 : o_group};
                seg_of_item[o_item + p] = (int)sg;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_run[0] += t_item; s_run[1] += t_slot; s_run[2] += t_group; }
        __syncthreads();
    }
    const int n_parts = min(s_run[0], n_out_cap);
    for (int k = n_parts + threadIdx.x; k < n_out_cap; k += SPLIT_TPB) out[k] = WorkItem{0, 0, 0, 0, 0, 0, 0, 0};
    // seg_off[s] = first part of segment s (segments are non-decreasing along the list)
    for (long long sgm = threadIdx.x; sgm <= n_seg; sgm += SPLIT_TPB) {
        int lo = 0, hi = n_parts;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg_of_item[mid] < (int)sgm) lo = mid + 1; else hi = mid; }
        seg_off[sgm] = sgm == n_seg ? n_parts : lo;
    }
}

// ---- SPJ records of the force pass: the cells' moments ----
__global__ void __launch_bounds__(TPB) spj_mono_kernel(int n_cells, const double *__restrict__ mom, SpjMonoAos *__restrict__ out)
{
    const int c = blockIdx.x * TPB + threadIdx.x;
    if (c >= n_cells) return;
    SpjMonoAos s;
    s.mass = mom[(size_t)c * 10];
    for (int k = 0; k < 3; k++) s.pos[k] = mom[(size_t)c * 10 + 1 + k];
    out[c] = s;
}

inline int nblk(long long n, int per) { return (int)((n + per - 1) / per); }

}  // namespace

const void *tree_cell_moments() { return S.c_mom.p; }
const int *tree_sorted_to_original() { return (const int *)S.idx_b.p; }
const int *tree_walk_ni() { return (const int *)S.w_ni.p; }

void tree_phase_ms(float ms[6])
{
    for (int k = 0; k < 6; k++) {
        ms[k] = 0.f;
        if (S.timed) cudaEventElapsedTime(&ms[k], S.ev[k], S.ev[k + 1]);
    }
}

int tree_stamps(unsigned long long *out, int cap)
{
    if (!S.h_meta_stamps) return 0;
    const Meta *m = S.h_meta_stamps;
    int n = std::min(std::min(m->n_stamp, 96), cap);
    for (int k = 0; k < n; k++) out[k] = m->stamp[k];
    return n;
}

void tree_release()
{
    for (Buf *b : {&S.keys_a, &S.keys_b, &S.idx_a, &S.idx_b, &S.cub_temp, &S.bbox_part, &S.meta, &S.blk_cnt,
                   &S.c_meta, &S.c_mom, &S.c_box, &S.fr_a, &S.fr_b, &S.grp_at, &S.walk_cell, &S.w_epi_off, &S.w_ni, &S.w_ne,
                   &S.w_ns, &S.w_ed, &S.w_sd, &S.w_nitems, &S.w_ioff, &S.item_key_a, &S.item_key_b, &S.item_a, &S.item_b, &S.item_cum, &S.item_seg})
        b->release();
    if (S.h_meta) { cudaFreeHost(S.h_meta); S.h_meta = nullptr; S.h_meta_stamps = nullptr; }
    if (S.ev_ok) { for (auto &e : S.ev) cudaEventDestroy(e); S.ev_ok = false; }
    S.timed = false; S.cell_cap = 0; S.n = 0; S.coop_blocks = 0;
}

static KP make_kp(const TreeCfg &cfg, const void *epj_sorted)
{
    KP P;
    P.n = cfg.n; P.n_leaf = cfg.n_leaf; P.n_group = cfg.n_group; P.cell_cap = S.cell_cap; P.theta = cfg.theta;
    P.epj = (const EpjAos *)epj_sorted; P.key = (const uint64_t *)S.keys_b.p;
    P.c_meta = (int4 *)S.c_meta.p; P.c_mom = (double *)S.c_mom.p; P.c_box = (double *)S.c_box.p;
    P.grp_at = (int *)S.grp_at.p; P.meta = (Meta *)S.meta.p;
    return P;
}

int tree_phase1(const TreeCfg &cfg, const TreeSrc &src, void *epj_sorted, void *epi, void *epj_packed, TreeCounts *counts,
                cudaStream_t st, int *launches)
{
    const int n = cfg.n;
    S.timed = false;
    if (!S.ev_ok) { for (auto &e : S.ev) CK(cudaEventCreate(&e)); S.ev_ok = true; }
    if (!S.h_meta) CK(cudaMallocHost((void **)&S.h_meta, sizeof(Meta)));
    if (S.coop_blocks == 0) {
        int dev = 0, sms = 0, occ = 0, coop = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tree_coop_kernel, TPB, 0));
        if (!coop || occ < 1) return (int)cudaErrorCooperativeLaunchTooLarge;
        S.coop_blocks = std::min(NB, occ * sms);          // every block must be resident: grid barriers
    }
    // the host builder's tree of the N = 1e6 disk has 0.74 n cells; a rejected capacity doubles and retries
    if (S.n != n || S.cell_cap == 0) S.cell_cap = 2 * n + 4096;
    S.n = n;
    for (int attempt = 0;; attempt++) {
        const size_t C = (size_t)S.cell_cap;
        CK(S.keys_a.reserve((size_t)n * 8)); CK(S.keys_b.reserve((size_t)n * 8));
        CK(S.idx_a.reserve((size_t)n * 4)); CK(S.idx_b.reserve((size_t)n * 4));
        CK(S.bbox_part.reserve((size_t)NB * 6 * 8)); CK(S.meta.reserve(sizeof(Meta)));
        CK(S.blk_cnt.reserve(NB * 4));
        CK(S.c_meta.reserve(C * sizeof(int4))); CK(S.c_mom.reserve(C * 80)); CK(S.c_box.reserve(C * 96));
        CK(S.fr_a.reserve(C * 4 + 64)); CK(S.fr_b.reserve(C * 4 + 64));     // a frontier <= the cells of its level
        CK(S.grp_at.reserve((size_t)n * 4)); CK(S.walk_cell.reserve((size_t)n * 4 + 16));
        KP P = make_kp(cfg, epj_sorted);
        const PosView raw = src.pos ? PosView{src.pos, src.pos_stride} : PosView{reinterpret_cast<const double *>(src.epj) + 1, (int)(sizeof(EpjAos) / 8)};

        CK(cudaEventRecord(S.ev[0], st));
        CK(cudaMemsetAsync(S.grp_at.p, 0xff, (size_t)n * 4, st));
        const int nbb = std::min(NB, nblk(n, TPB));
        bbox_kernel<<<nbb, TPB, 0, st>>>(raw, n, (double *)S.bbox_part.p);
        bbox_final_kernel<<<1, 32, 0, st>>>((const double *)S.bbox_part.p, nbb, P);
        key_kernel<<<nblk(n, TPB), TPB, 0, st>>>(raw, n, P.meta, (uint64_t *)S.keys_a.p, (int *)S.idx_a.p);
        CK(cudaGetLastError());
        *launches += 3;
        size_t tb = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint64_t *)S.keys_a.p, (uint64_t *)S.keys_b.p,
                                           (const int *)S.idx_a.p, (int *)S.idx_b.p, n, 0, 3 * LEVEL_HI, st));
        CK(S.cub_temp.reserve(tb));
        CK(cub::DeviceRadixSort::SortPairs(S.cub_temp.p, tb, (const uint64_t *)S.keys_a.p, (uint64_t *)S.keys_b.p,
                                           (const int *)S.idx_a.p, (int *)S.idx_b.p, n, 0, 3 * LEVEL_HI, st));
        tie_fix_kernel<<<nblk(n, TPB), TPB, 0, st>>>(raw, n, P.meta, (const uint64_t *)S.keys_b.p, (int *)S.idx_b.p);
        if (attempt == 0 && src.before_gather) {
            if (int e = src.before_gather(src.before_gather_arg)) return e;
        }
        if (src.pos)
            gather_soa_kernel<<<nblk((long long)n * 8, TPB), TPB, 0, st>>>(src.pos, src.mass, src.r_out, src.r_search, src.vel, src.pos_stride, src.col_stride, src.rank,
                                                                           (const int *)S.idx_b.p, n, (uint4 *)epj_sorted, (EpiAos *)epi,
                                                                           (EpjPacked *)epj_packed);
        else
            gather_kernel<<<nblk((long long)n * 8, TPB), TPB, 0, st>>>((const uint4 *)src.epj, (const int *)S.idx_b.p, n,
                                                                       (uint4 *)epj_sorted, (EpiAos *)epi, (EpjPacked *)epj_packed);
        CK(cudaGetLastError());
        *launches += 2;
        CK(cudaEventRecord(S.ev[1], st));

        // cells + moments: one cooperative launch; frontier arrays ping-pong, level 0's frontier is the root
        CK(cudaMemsetAsync(S.fr_a.p, 0, 4, st));
        {
            int *fr_a = (int *)S.fr_a.p, *fr_b = (int *)S.fr_b.p, *bc = (int *)S.blk_cnt.p;
            void *args[] = {&P, &fr_a, &fr_b, &bc};
            CK(cudaLaunchCooperativeKernel((const void *)tree_coop_kernel, dim3(S.coop_blocks), dim3(TPB), args, 0, st));
            *launches += 1;
        }
        CK(cudaEventRecord(S.ev[2], st));
        // i-groups in Morton order: grp_at[first particle] = cell, compacted
        int *d_nwalk = &((Meta *)S.meta.p)->n_walk;
        CK(cub::DeviceSelect::If(nullptr, tb, (const int *)S.grp_at.p, (int *)S.walk_cell.p, d_nwalk, n, NonNeg(), st));
        CK(S.cub_temp.reserve(tb));
        CK(cub::DeviceSelect::If(S.cub_temp.p, tb, (const int *)S.grp_at.p, (int *)S.walk_cell.p, d_nwalk, n, NonNeg(), st));
        CK(cudaMemcpyAsync(S.h_meta, S.meta.p, sizeof(Meta), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(S.ev[3], st));
        CK(cudaStreamSynchronize(st));
        if (S.h_meta->overflow) {
            if (attempt >= 4) { counts->overflow = 1; return -1; }
            S.cell_cap *= 2;
            continue;
        }
        break;
    }
    KP P = make_kp(cfg, epj_sorted);
    for (int l = 0; l <= N_LVL; l++) S.lvl_start[l] = S.h_meta->lvl_start[l];
    S.n_cells = S.lvl_start[N_LVL];
    S.n_walk = S.h_meta->n_walk;
    walk_range_kernel<<<1, 32, 0, st>>>(P, (const int *)S.walk_cell.p, S.n_walk, cfg.part_rank, cfg.part_world > 0 ? cfg.part_world : 1);
    CK(cudaGetLastError());
    ++*launches;
    S.n_levels = 0;
    for (int l = 0; l < N_LVL; l++) if (S.lvl_start[l + 1] > S.lvl_start[l]) S.n_levels = l + 1;
    const int nw = S.n_walk;
    for (Buf *b : {&S.w_epi_off, &S.w_ni, &S.w_ne, &S.w_ns, &S.w_nitems, &S.w_ioff}) CK(b->reserve((size_t)nw * 4 + 16));
    CK(S.w_ed.reserve((size_t)nw * 8 + 16)); CK(S.w_sd.reserve((size_t)nw * 8 + 16));
    WalkP A;
    A.P = P; A.walk_cell = (const int *)S.walk_cell.p; A.n_walk = nw;
    A.epi_off = (int *)S.w_epi_off.p; A.ni = (int *)S.w_ni.p; A.n_epj = (int *)S.w_ne.p; A.n_spj = (int *)S.w_ns.p;
    A.epj_disp = nullptr; A.spj_disp = nullptr; A.adr_epj = nullptr; A.adr_spj = nullptr;
    if (nw > 0) {
        walk_kernel<false><<<std::min(nblk(nw, WALK_WPB), 148 * 8), WALK_WPB * 32, 0, st>>>(A);
        CK(cudaGetLastError());
        ++*launches;
        size_t tb = 0, tb2 = 0;
        cub::TransformInputIterator<long long, AsLL, const int *> ne_ll((const int *)S.w_ne.p, AsLL()), ns_ll((const int *)S.w_ns.p, AsLL());
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ne_ll, (long long *)S.w_ed.p, nw, st));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, (const int *)S.w_nitems.p, (int *)S.w_ioff.p, nw, st));
        CK(S.cub_temp.reserve(std::max(tb, tb2)));
        CK(cub::DeviceScan::ExclusiveSum(S.cub_temp.p, tb, ne_ll, (long long *)S.w_ed.p, nw, st));
        CK(cub::DeviceScan::ExclusiveSum(S.cub_temp.p, tb, ns_ll, (long long *)S.w_sd.p, nw, st));
        totals_kernel<<<1, 1024, 0, st>>>(nw, (const int *)S.w_ni.p, (const int *)S.w_ne.p, (const int *)S.w_ns.p, P.meta,
                                          cfg.warp_slots, cfg.tile_cap, cfg.jsplit, cfg.rmax, cfg.split_m);
        item_count_kernel<<<nblk(nw, TPB), TPB, 0, st>>>(nw, (const int *)S.w_ni.p, P.meta, cfg.jsplit, (int *)S.w_nitems.p);
        CK(cub::DeviceScan::ExclusiveSum(S.cub_temp.p, tb2, (const int *)S.w_nitems.p, (int *)S.w_ioff.p, nw, st));
        item_total_kernel<<<1, 1, 0, st>>>(nw, (const int *)S.w_nitems.p, (const int *)S.w_ioff.p, P.meta);
        CK(cudaGetLastError());
        *launches += 3;
    }
    CK(cudaMemcpyAsync(S.h_meta, S.meta.p, sizeof(Meta), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(S.ev[4], st));
    CK(cudaStreamSynchronize(st));
    S.h_meta_stamps = S.h_meta;
    counts->n_cells = S.n_cells; counts->n_walk = nw; counts->n_levels = S.n_levels;
    counts->w0 = S.h_meta->w0; counts->w1 = S.h_meta->w1; counts->e0 = S.h_meta->e0; counts->e1 = S.h_meta->e1;
    counts->overflow = S.h_meta->overflow;
    counts->n_items = nw > 0 ? S.h_meta->n_items : 0; counts->cap = nw > 0 ? S.h_meta->cap : 0;
    counts->n_adr_epj = nw > 0 ? S.h_meta->n_adr_epj : 0; counts->n_adr_spj = nw > 0 ? S.h_meta->n_adr_spj : 0;
    counts->n_int_epep = nw > 0 ? S.h_meta->n_int_epep : 0; counts->n_int_epsp = nw > 0 ? S.h_meta->n_int_epsp : 0;
    return counts->overflow ? -1 : 0;
}

int tree_phase2(const TreeCfg &cfg, const TreeOut &out, cudaStream_t st, int *launches)
{
    const int nw = S.n_walk;
    const cudaMemcpyKind D2D = cudaMemcpyDeviceToDevice;
    if (nw > 0) {
        WalkP A;
        A.P = make_kp(cfg, nullptr); A.walk_cell = (const int *)S.walk_cell.p; A.n_walk = nw;
        A.epi_off = nullptr; A.ni = nullptr; A.n_epj = nullptr; A.n_spj = nullptr;
        A.epj_disp = (const long long *)S.w_ed.p; A.spj_disp = (const long long *)S.w_sd.p;
        A.adr_epj = out.adr_epj; A.adr_spj = out.adr_spj;
        walk_kernel<true><<<std::min(nblk(nw, WALK_WPB), 148 * 8), WALK_WPB * 32, 0, st>>>(A);
        CK(cudaGetLastError());
        ++*launches;
    }
    CK(cudaEventRecord(S.ev[5], st));
    if (nw > 0) {
        CK(cudaMemcpyAsync(out.epi_off, S.w_epi_off.p, (size_t)nw * 4, D2D, st));
        if (out.ni) CK(cudaMemcpyAsync(out.ni, S.w_ni.p, (size_t)nw * 4, D2D, st));
        CK(cudaMemcpyAsync(out.n_epj, S.w_ne.p, (size_t)nw * 4, D2D, st));
        CK(cudaMemcpyAsync(out.n_spj, S.w_ns.p, (size_t)nw * 4, D2D, st));
        CK(cudaMemcpyAsync(out.epj_disp, S.w_ed.p, (size_t)nw * 8, D2D, st));
        CK(cudaMemcpyAsync(out.spj_disp, S.w_sd.p, (size_t)nw * 8, D2D, st));
        // work items, longest first (stable: equal costs keep walk order, as the host list does)
        const int nit = S.h_meta->n_items;
        CK(S.item_key_a.reserve((size_t)nit * 8 + 16)); CK(S.item_key_b.reserve((size_t)nit * 8 + 16));
        CK(S.item_a.reserve((size_t)nit * sizeof(BaseItem) + 16)); CK(S.item_b.reserve((size_t)nit * sizeof(BaseItem) + 16));
        item_emit_kernel<<<nblk(nw, TPB), TPB, 0, st>>>(nw, (const int *)S.w_ni.p, (const int *)S.w_ne.p, (const int *)S.w_ns.p,
                                                        (const int *)S.w_ioff.p, (const Meta *)S.meta.p, cfg.jsplit,
                                                        (BaseItem *)S.item_a.p, (uint64_t *)S.item_key_a.p);
        CK(cudaGetLastError());
        ++*launches;
        if (nit > 0) {
            size_t tb = 0;
            CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const uint64_t *)S.item_key_a.p, (uint64_t *)S.item_key_b.p,
                                                         (const int4 *)S.item_a.p, (int4 *)S.item_b.p, nit, 0, 64, st));
            CK(S.cub_temp.reserve(tb));
            CK(cub::DeviceRadixSort::SortPairsDescending(S.cub_temp.p, tb, (const uint64_t *)S.item_key_a.p, (uint64_t *)S.item_key_b.p,
                                                         (const int4 *)S.item_a.p, (int4 *)S.item_b.p, nit, 0, 64, st));
            if (out.n_items_out > nit) {      // the caller sized `items` for a segmented list (items.h: split_items_bound)
                CK(S.item_cum.reserve((size_t)nit * 8 + 16)); CK(S.item_seg.reserve((size_t)out.n_items_out * 4 + 16));
                item_split_kernel<<<1, SPLIT_TPB, 0, st>>>(nit, (const BaseItem *)S.item_b.p, (const uint64_t *)S.item_key_b.p,
                                                           (const int *)S.w_ne.p, (const int *)S.w_ns.p, cfg.warp_slots,
                                                           (long long *)S.item_cum.p, (int *)S.item_seg.p,
                                                           (WorkItem *)out.items, out.n_items_out, out.seg_off);
            } else
                item_widen_kernel<<<nblk(nit, TPB), TPB, 0, st>>>(nit, (const BaseItem *)S.item_b.p, (WorkItem *)out.items);
            CK(cudaGetLastError());
            ++*launches;
        }
    }
    if (S.n_cells > 0) {
        if (cfg.quad) { if (out.spj_aos) CK(cudaMemcpyAsync(out.spj_aos, S.c_mom.p, (size_t)S.n_cells * 80, D2D, st)); }
        else {
            spj_mono_kernel<<<nblk(S.n_cells, TPB), TPB, 0, st>>>(S.n_cells, (const double *)S.c_mom.p, (SpjMonoAos *)out.spj_aos);
            CK(cudaGetLastError());
            ++*launches;
        }
    }
    CK(cudaEventRecord(S.ev[6], st));
    S.timed = true;
    return 0;
}

}  // namespace gbt
