// dev_tree.h -- launch interface of dev_tree.cu: the interaction-list builder on the GPU
// (SURVEY 8 f1).  Device pointers only; gplum_b200.cu owns the walk-set / j-set buffers and
// reserves them between the two phases.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gbt {

struct TreeCfg {
    int n;                       // particles
    double theta;
    int n_leaf, n_group;
    int quad;                    // SPJ layout: 1 = MySPJQuadrupole (80 B), 0 = MySPJMonopole (32 B)
    // work-list policy (items.h)
    long long warp_slots; int tile_cap, jsplit, rmax, split_m;
    int part_rank = 0, part_world = 1;   // multi-GPU: lists and work items only for this rank's share of the walks
};

// totals of a build, valid on the host after phase1 returned
struct TreeCounts {
    int n_cells, n_walk, n_items, cap;
    long long n_adr_epj, n_adr_spj, n_int_epep, n_int_epsp;
    int n_levels, overflow;      // overflow: 1 = cell capacity, 2 = walk stack
    int w0, w1, e0, e1;          // this rank's walks [w0, w1) = particles [e0, e1) in tree order
};

struct TreeOut {                 // destination buffers of phase 2 (sizes from TreeCounts)
    int *epi_off, *ni, *n_epj, *n_spj;           // n_walk (ni may be NULL)
    long long *epj_disp, *spj_disp;              // n_walk
    int *adr_epj, *adr_spj;                      // n_adr_epj, n_adr_spj
    void *items;                                 // n_items_out x WorkItem (32 B)
    int n_items_out;                             // = TreeCounts::n_items (one work item per tile), or tree_items_bound() of a
                                                 //   pass that splits its tiles along j (trailing entries are empty items)
    int *seg_off;                                // warp_slots + 1 entries, written when the pass is laid out in segments
    void *spj_aos;                               // n_cells x (80 | 32) B; NULL (quadrupole only): read tree_cell_moments() in place
};

// Where phase 1 finds the particles.  Records: epj = EPJGrav[n] on the device, any order.  Columns (pos != NULL): the
// caller's SoA arrays on the device, pos is [n][3]; id_local = id = index, myrank = rank, vel = vel[i] or 0, acc_d = 0.
// Keys and sort need the positions only: `before_gather` (optional) is called once after they are enqueued -- the
// caller uploads the remaining columns there and makes the stream wait for them, so that the copy of 24 of the 48
// bytes per particle overlaps the sort.
struct TreeSrc {
    const void *epj = nullptr;
    const double *pos = nullptr, *mass = nullptr, *r_out = nullptr, *r_search = nullptr, *vel = nullptr;
    int pos_stride = 3, col_stride = 1;      // doubles between particles: packed columns, or one interleaved record per
                                             // particle ({pos[3], mass, r_out, r_search}: strides 6 and 6, columns offset into it)
    int rank = 0;
    int (*before_gather)(void *) = nullptr; void *before_gather_arg = nullptr;
};

// Phase 1: Morton keys, radix sort, gather (writes epj_sorted = EPJGrav[n] and epi = EPIGrav[n] in tree
// order and, if epj_packed != NULL, the force kernel's packed 48 B records of the same particles), cells, moments, i-groups, counting walk, scans.  Synchronises the stream twice (cell / group
// counts, list totals).  epj_unsorted: EPJGrav[n] on the device, any order.  Returns cudaError_t (0 = ok)
// or -1 with counts->overflow set.
int tree_phase1(const TreeCfg &cfg, const TreeSrc &src, void *epj_sorted, void *epi, void *epj_packed,
                TreeCounts *counts, cudaStream_t st, int *launches);
// Phase 2: filling walk, work items (sorted longest first), SPJ records.
int tree_phase2(const TreeCfg &cfg, const TreeOut &out, cudaStream_t st, int *launches);
const int *tree_sorted_to_original();            // device pointer, n entries, valid after phase 1
const void *tree_cell_moments();                 // device pointer, n_cells x 80 B {mass, pos, quad}: MySPJQuadrupole records
const int *tree_walk_ni();                       // device pointer, n_walk entries: i-particles per walk
// milliseconds between the phase marks of the last build: [0] keys+sort+gather, [1] cells+moments (one
// cooperative kernel), [2] i-group compaction + first host sync, [3] counting walk + scans + second sync,
// [4] filling walk, [5] items + SPJ
void tree_phase_ms(float ms[6]);
// globaltimer (ns) at the level boundaries inside the cooperative cells+moments kernel of the last build
int tree_stamps(unsigned long long *out, int cap);
void tree_release();

}  // namespace gbt
