// records.h -- the reference's particle / force records as the device code sees them, and the
// work item of the force pass.  Shared by kernels.cuh, soft_corr.cu and dev_tree.cu.
#pragma once

namespace gb {

// ---- the reference's AoS structs as seen by the device (default macro set) ----
struct EpiAos { int id_local, myrank; double pos[3]; double r_out, r_search; };                    // 48
struct EpjAos { int id_local, myrank; double pos[3]; double r_out, r_search; long long id;
                double mass; double vel[3]; double acc_d[3]; };                                    // 112
struct SpjQuadAos { double mass; double pos[3]; double quad[6]; };                                 // 80
struct SpjMonoAos { double mass; double pos[3]; };                                                 // 32
struct __align__(16) ForceAos { float acc[3]; float phi; int number, rank, id_max, id_min; };      // 32
static_assert(sizeof(EpiAos) == 48 && sizeof(EpjAos) == 112 && sizeof(SpjQuadAos) == 80 &&
              sizeof(SpjMonoAos) == 32 && sizeof(ForceAos) == 32, "reference layout");

// base item = one i-tile of one walk against its whole lists (what the cost sort orders); work item = the part of
// a base item one warp executes: j-tiles [t0, t1) of the walk's tile sequence (EP tiles, then SP tiles; t1 < 0 = all).
// Parts of a split tile share `group` (arrival counter) and write their partial sums to scratch slot slot0 + part
// index (items.h: cfg bits).
struct BaseItem { int walk, i0, ni, cfg; };
struct __align__(16) WorkItem { int walk, i0, ni, cfg; int t0, t1, slot0, group; };
static_assert(sizeof(BaseItem) == 16 && sizeof(WorkItem) == 32, "item layout");

}  // namespace gb
