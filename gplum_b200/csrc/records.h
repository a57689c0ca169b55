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

// ---- packed j-records in HBM (written by the pack kernels, read by vectorised 16 B loads) ----
struct __align__(16) EpjPacked {   // 48 B
    double x, y;
    double z; float m, rout2;
    float rs2; int id; int rank; int pad;
};
struct __align__(16) SpjPacked {   // 64 B; Q = (3q - tr*I)/RS^2, mtr = -(eps2*tr)/RS^2 hoisted (j-only terms;
                                   // the exact power-of-two scale pairs with a = RS*y in the pair loop)
    double x, y;
    double z; float m, qxx;
    float qyy, qzz, qxy, qyz;
    float qzx, mtr, pad0, pad1;
};
static_assert(sizeof(EpjPacked) == 48 && sizeof(SpjPacked) == 64, "packed layout");

// The packed EP record of an EPJGrav's fields (FP32 roundings as src/gravity_kernel_epep.pikg:53-68 makes them: the
// squares are taken in FP32, the search radius carries the kernel's 1.0201 safety factor)
#if defined(__CUDACC__)
__device__ __forceinline__ EpjPacked epj_pack(const double pos[3], double mass, double r_out, double r_search, int id_local, int myrank)
{
    EpjPacked o;
    o.x = pos[0]; o.y = pos[1]; o.z = pos[2];
    o.m = (float)mass;
    const float ro = (float)r_out, rs = (float)r_search;
    o.rout2 = __fmul_rn(ro, ro);
    o.rs2 = __fmul_rn(__fmul_rn(rs, rs), 1.0201f);
    o.id = id_local; o.rank = myrank; o.pad = 0;
    return o;
}
#endif

// base item = one i-tile of one walk against its whole lists (what the cost sort orders); work item = the part of
// a base item one warp executes: j-tiles [t0, t1) of the walk's tile sequence (EP tiles, then SP tiles; t1 < 0 = all).
// Parts of a split tile share `group` (arrival counter) and write their partial sums to scratch slot slot0 + part
// index (items.h: cfg bits).
struct BaseItem { int walk, i0, ni, cfg; };
struct __align__(16) WorkItem { int walk, i0, ni, cfg; int t0, t1, slot0, group; };
static_assert(sizeof(BaseItem) == 16 && sizeof(WorkItem) == 32, "item layout");

}  // namespace gb
