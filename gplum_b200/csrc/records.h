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

struct WorkItem { int walk, i0, ni, cfg; };

}  // namespace gb
