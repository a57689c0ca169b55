// kernels.cuh -- sm_100a device code of the P3T soft-force pass.
//
// One launch evaluates a whole FDPS force pass (TreeForForce::calcForce,
// FDPS/src/tree_for_force_impl_force.hpp:1404-1564): every work item is one i-tile of one
// i-group ("walk") against that group's EP and SP interaction lists, given as indices into
// packed j-arrays resident in HBM.  Arithmetic follows src/gravity_kernel_epep.pikg:53-97 and
// src/gravity_kernel_epsp.pikg:47-100: FP64 shift by the group's first i-particle, then FP32.
//
// Per pair, EP-EP (hot loop, 15.5 issue slots for the DSL's 30 flop + rsqrt):
//   d = xj-xi (3 FADD); r2 = d.d+eps2 (3 FFMA); rmin = min3(rmin, r2[i0], r2[i1]) (FMNMX3 per 2 pairs);
//   r2c = max3(r2, rout2_i, rout2_j) (FMNMX3; max(a,b)^2 == max(a^2,b^2) exactly in FP);
//   a = MUFU.RSQ(r2c) (GB_NEWTON below: the DSL's Newton step is what PIKG adds to a 12-bit estimate; MUFU.RSQ has
//   the refined accuracy already.  With GB_NEWTON = 1 the step runs as a = y*(3 - r2c*y*y) = 2*y', 3 more slots, and
//   its *0.5, a power of two, is carried exactly in the accumulators' scale and undone at the write);
//   t = m*a; v = (a*a)*t (3); acc += v*d (3 FFMA); phi -= t (FADD).
// Neighbour candidates: the hot loop only evaluates a conservative filter (min r2 < T with
// T >= rsearch2 of every pair of the lane and tile, widened by 2^-16); j-groups that pass
// (rare: self + true candidates) are re-tested in the reference's exact, non-fused evaluation
// order (src/gravity_kernel.hpp:94,104-111) so number/rank/id_max/id_min are bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "records.h"
#include "items.h"

namespace gb {

#ifndef GB_UNROLL
#define GB_UNROLL 4
#endif
#ifndef GB_MINB2
#define GB_MINB2 6
#endif
constexpr int UNROLL = GB_UNROLL;
// Reciprocal square root.  The DSL writes rsqrt(r2) and PIKG expands it for the target: on AVX2 a 12-bit estimate
// + one Newton step (rel. error ~2e-7), on AVX-512 a 14-bit one.  MUFU.RSQ (rsqrt.approx.f32) is specified to 2^-22.9
// = 1.3e-7 over the whole range, i.e. it already has the accuracy the reference's estimate reaches AFTER its Newton
// step, so the shipped build uses it as is (GB_NEWTON = 0).  GB_NEWTON = 1 compiles the step in (3 more issue slots per
// pair, ~10 % of a pass); measured difference between the two on the N = 1e6 pass: profiles/r2_newton_step.txt.
#ifndef GB_NEWTON
#define GB_NEWTON 0
#endif
// rinv_scaled() returns RS * rsqrt(x); the power-of-two RS is carried exactly in the accumulators' scale
constexpr float RS = GB_NEWTON ? 2.0f : 1.0f;
constexpr float RS2_INV = 1.0f / (RS * RS), RS_INV = 1.0f / RS, RS3_INV = 1.0f / (RS * RS * RS);

struct PassParams {
    const EpiAos *epi;            // concatenated i-particles
    const int *epi_off;           // per walk
    const int *adr_epj; const long long *epj_disp; const int *n_epj;
    const int *adr_spj; const long long *spj_disp; const int *n_spj;
    const EpjPacked *epj; const SpjPacked *spj;
    // multi-GPU peer mode: EP index = (owner rank << peer_shift) | index in the owner's slab; slabs of
    // the other ranks are their HBM, mapped through CUDA IPC and read over NVLink by the same cp.async
    const EpjPacked *const *peer_epj; int peer_shift;
    ForceAos *force;
    const WorkItem *items;
    float eps2;
    int rank_squared;
    // candidate capture for the changeover correction (soft_corr.cu; src/gravity_soft.h:245-372): every
    // (i, j) that passes the exact candidate test is appended to `pairs` (i = index into epi/force,
    // j = EP index as listed), and the j that IS i (same id_local and rank) is noted in self_adr[i].
    // All three are nullptr unless capture is enabled.
    int *self_adr;
    int2 *pairs; unsigned int *pair_count; unsigned int pair_cap;
    // j-split tiles (items.h): the K parts of a tile leave their partial sums in scratch[(slot0 + part) * 64 + i]
    // and count their arrival in arrive[group]; the last one adds the partial sums in part order and writes ForceGrav
    ForceAos *scratch; int *arrive;
    // one-wave passes (items.h, segments): warp s executes items [seg_off[s], seg_off[s + 1]); nullptr: item s
    const int *seg_off; int n_seg;
    // multi-GPU peer mode: items flagged ITEM_PEER_WAIT start once every rank's flag word has reached the epoch,
    // i.e. once all slabs of this step are packed (gplum_b200.cu: peer_pack)
    const int *peer_flags; int peer_world, peer_epoch;
    // Placed passes (items.h: place_item): a pass whose warps are all resident at once has no dynamic balancing; its
    // warps claim their item by WHERE they run -- bin = SM x scheduler, k = how many warps of that bin came before.
    // place[0 .. place_bins) = the bins' claim counters, [place_bins] = items claimed so far, [place_bins + 1] = warps
    // that have exited, [place_bins + 2] = CTAs at the fused pack's barrier; all zero before and after a launch.  nullptr: warp s runs item s (or segment s).
    int *place; int place_bins, place_rounds;
    // Multi-GPU peer mode, fused pack (placed passes only: every CTA is resident, the launch is cooperative): before
    // their first item the CTAs pack this rank's EPJ into its slab and its SPJ, meet at a grid barrier
    // (place[place_bins + 2]), and the last one to arrive stores the epoch into every rank's flag array -- what
    // peer_pack_kernel does as a launch of its own.  fp_slab == nullptr: off.
    const EpjAos *fp_epj_in; int fp_n_epj; EpjPacked *fp_slab;
    const void *fp_spj_in; int fp_n_spj; SpjPacked *fp_spj_out; int fp_quad, fp_trace;
    void *const *fp_slab0_of; size_t fp_flag_off; int fp_rank;
    // placement trace (gplum_b200_debug_trace): per item {start, end (globaltimer ns), %smid | %warpid << 32, times run};
    // nullptr = off
    unsigned long long *trace;
};

// spins until flags[0 .. world) >= epoch (the peers store their epoch over NVLink after packing, st.release.sys)
__device__ __forceinline__ void peer_flags_wait(const int *flags, int world, int epoch)
{
    const int q = threadIdx.x & 31;
    if (q < world) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
            if (v - epoch >= 0) break;
            __nanosleep(100);
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > 10000000000ull) __trap();           // 10 s: a peer died; fail loudly instead of hanging
        }
    }
    __syncwarp();
}

__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rinv_scaled(float x)
{
    const float y = rsqrt_approx(x);
#if GB_NEWTON
    return y * fmaf(-x, y * y, 3.0f);      // the DSL's y*(3 - x*y*y)*0.5 without the 0.5
#else
    return y;
#endif
}
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ double2 ldg_d2(const void *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_f4(const void *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------
// pack kernels: AoS (as FDPS holds epj_sorted_/spj_sorted_) -> packed records
// ------------------------------------------------------------------------------------------
// A block's 256 AoS records are first copied to shared memory with coalesced 16 B loads (a thread
// reading its own 112 B / 80 B record straight from HBM touches 4-5 sectors per load instruction);
// each thread then picks its fields from shared memory.  `in` must be 16 B aligned.
constexpr int PACK_BLOCK = 256;
template <int REC>
__device__ __forceinline__ void stage_records(const void *__restrict__ in, int n, unsigned char *sm, int blk)
{
    static_assert(REC % 16 == 0, "record size");
    const int first = blk * PACK_BLOCK;
    const int cnt = min(PACK_BLOCK, n - first);
    const int n16 = cnt * (REC / 16);
    const uint4 *src = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(in) + (size_t)first * REC);
    for (int k = threadIdx.x; k < n16; k += PACK_BLOCK) reinterpret_cast<uint4 *>(sm)[k] = __ldg(src + k);
    __syncthreads();
}

// one block's 256 EPJGrav records -> packed records (sm: PACK_BLOCK * 112 B)
__device__ __forceinline__ void pack_epj_block(const EpjAos *__restrict__ in, int n, EpjPacked *__restrict__ out, int blk, unsigned char *sm)
{
    stage_records<sizeof(EpjAos)>(in, n, sm, blk);
    const int i = blk * PACK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const EpjAos &a = reinterpret_cast<const EpjAos *>(sm)[threadIdx.x];
    out[i] = epj_pack(a.pos, a.mass, a.r_out, a.r_search, a.id_local, a.myrank);
}

__global__ void __launch_bounds__(PACK_BLOCK) pack_epj_kernel(const EpjAos *__restrict__ in, int n, EpjPacked *__restrict__ out)
{
    __shared__ __align__(16) unsigned char sm[PACK_BLOCK * sizeof(EpjAos)];
    pack_epj_block(in, n, out, blockIdx.x, sm);
}

// Halo send staging of the multi-GPU step: dst[k] = src[idx[k]] for 48 B packed EP records,
// one 16 B chunk per thread (the trimmed LET exchange, FDPS/src/tree_for_force_impl_exlet.hpp:343-403).
__global__ void gather_epj_packed_kernel(const uint4 *__restrict__ src, const int *__restrict__ idx, int n, uint4 *__restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int k = t / 3, c = t - 3 * k;
    dst[3 * (size_t)k + c] = src[3 * (size_t)idx[k] + c];
}

// one packed superparticle from its AoS record (quad: MySPJQuadrupole, else MySPJMonopole)
__device__ __forceinline__ SpjPacked spj_pack(const void *rec, int quad, int trace_as_shipped, float eps2)
{
    SpjPacked o;
    if (quad) {
        const SpjQuadAos &a = *reinterpret_cast<const SpjQuadAos *>(rec);
        o.x = a.pos[0]; o.y = a.pos[1]; o.z = a.pos[2]; o.m = (float)a.mass;
        const float qxx = (float)a.quad[0], qyy = (float)a.quad[1], qzz = (float)a.quad[2];
        const float qxy = (float)a.quad[3], qzx = (float)a.quad[4], qyz = (float)a.quad[5];
        const float tr = trace_as_shipped ? (float)(a.quad[0] + a.quad[1] + a.quad[0])
                                          : __fadd_rn(__fadd_rn(qxx, qyy), qzz);
        o.qxx = RS2_INV * __fsub_rn(__fmul_rn(3.0f, qxx), tr);
        o.qyy = RS2_INV * __fsub_rn(__fmul_rn(3.0f, qyy), tr);
        o.qzz = RS2_INV * __fsub_rn(__fmul_rn(3.0f, qzz), tr);
        o.qxy = RS2_INV * __fmul_rn(3.0f, qxy); o.qyz = RS2_INV * __fmul_rn(3.0f, qyz); o.qzx = RS2_INV * __fmul_rn(3.0f, qzx);
        o.mtr = -RS2_INV * __fmul_rn(eps2, tr);
    } else {
        const SpjMonoAos &a = *reinterpret_cast<const SpjMonoAos *>(rec);
        o.x = a.pos[0]; o.y = a.pos[1]; o.z = a.pos[2]; o.m = (float)a.mass;
        o.qxx = o.qyy = o.qzz = o.qxy = o.qyz = o.qzx = 0.0f; o.mtr = 0.0f;
    }
    o.pad0 = o.pad1 = 0.0f;
    return o;
}

// quad: 1 = MySPJQuadrupole (80 B), 0 = MySPJMonopole (32 B).  trace_as_shipped reproduces
// src/gravity_kernel.hpp:177 (F32 <- qxx+qyy+qxx summed in F64).  sm: PACK_BLOCK * 80 B.
__device__ __forceinline__ void pack_spj_block(const void *__restrict__ in, int n, SpjPacked *__restrict__ out,
                                               int quad, int trace_as_shipped, float eps2, int blk, unsigned char *sm)
{
    if (quad) stage_records<sizeof(SpjQuadAos)>(in, n, sm, blk);
    else stage_records<sizeof(SpjMonoAos)>(in, n, sm, blk);
    const int i = blk * PACK_BLOCK + threadIdx.x;
    if (i >= n) return;
    out[i] = spj_pack(sm + (size_t)threadIdx.x * (quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos)), quad, trace_as_shipped, eps2);
}

__global__ void __launch_bounds__(PACK_BLOCK) pack_spj_kernel(const void *__restrict__ in, int n, SpjPacked *__restrict__ out,
                                int quad, int trace_as_shipped, float eps2)
{
    __shared__ __align__(16) unsigned char sm[PACK_BLOCK * sizeof(SpjQuadAos)];
    pack_spj_block(in, n, out, quad, trace_as_shipped, eps2, blockIdx.x, sm);
}

// Multi-GPU peer mode, one launch per step: blocks [0, nb_e) pack this rank's EPJ into its slab, the others pack its
// SPJ.  The EPJ block that finishes last stores the new epoch into this rank's entry of EVERY rank's flag array
// (the others' over NVLink): "my slab of this step is complete" (gplum_b200.cu: peer mode).
__global__ void __launch_bounds__(PACK_BLOCK) peer_pack_kernel(const EpjAos *__restrict__ epj_in, int n_epj, EpjPacked *__restrict__ slab,
                                                               const void *__restrict__ spj_in, int n_spj, SpjPacked *__restrict__ spj_out,
                                                               int quad, int trace_as_shipped, float eps2, int nb_e,
                                                               unsigned int *done, void *const *slab0_of, size_t flag_off,
                                                               int rank, int world, int epoch)
{
    __shared__ __align__(16) unsigned char sm[PACK_BLOCK * sizeof(EpjAos)];
    __shared__ int s_last;
    if ((int)blockIdx.x >= nb_e) {
        pack_spj_block(spj_in, n_spj, spj_out, quad, trace_as_shipped, eps2, blockIdx.x - nb_e, sm);
        return;
    }
    if (n_epj > 0) pack_epj_block(epj_in, n_epj, slab, blockIdx.x, sm);
    __threadfence();                                    // records visible in this GPU's L2, where the peers' loads arrive
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == (unsigned int)(nb_e - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) *done = 0;                    // ready for the next step
    if ((int)threadIdx.x < world) {
        int *f = reinterpret_cast<int *>(static_cast<char *>(slab0_of[threadIdx.x]) + flag_off) + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// force pass.  Every WARP is independent: it owns up to 32*R i-particles of one walk (one per
// lane and register slot), streams that walk's EP list and then its SP list through a private
// shared-memory tile of JW j-particles, and writes its forces.  No block-level barrier exists
// in the kernel; a CTA is only a container of WPB warps.
//   staging pipeline per j-tile t (EP tiles first, then SP tiles, one sequence):
//     cp.async(packed records of tile t+1 -> raw)  |  compute tile t  |  wait, convert raw -> tile
//   convert = FP64 subtract of the group origin, narrow to FP32 (gravity_kernel_epep.pikg:53-64)
// ------------------------------------------------------------------------------------------
constexpr int WPB = 4;         // warps per CTA
constexpr int JW = 64;         // j-particles per warp tile (2 per lane)
template <int IW>              // IW = max i-particles per warp (32 * RMAX)
struct WarpSmem {
    float4 raw[JW * 4];        // cp.async landing zone: 64 B per j (EP records use 48)
    float4 j4[JW];             // dx,dy,dz,m
    union {
        struct { float rout2[JW]; float rs2[JW]; int id[JW]; int rank[JW]; int adr[JW]; } ep;
        struct { float4 q0[JW]; float4 q1[JW]; } sp;   // Qxx,Qyy,Qzz,Qxy | Qyz,Qzx,mtr,-
    };
    unsigned long long mbar;   // completion barrier of the bulk copies of the tile in flight (one phase per tile)
    unsigned long long pad_;
    float i_rs2[IW];
    int i_id[IW], i_rank[IW];
    int nb_number[IW], nb_rank[IW], nb_idmax[IW], nb_idmin[IW];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- bulk asynchronous copies (the TMA unit's 1-D form, SASS UBLKCP) with mbarrier completion ----
// In tree order an interaction list is mostly runs of consecutive indices (a leaf's particles, the leaves of an
// opened subtree, sibling cells): one bulk copy moves a whole run of packed records into the warp's landing zone,
// where the per-lane cp.async form needs three or four 16 B copies per record.
__device__ __forceinline__ void mbar_init(unsigned mbar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// R = i-particles per lane; G = j-split: for short i-tiles (ni <= 32/G, R == 1) the warp is cut into
// G lane groups that hold the SAME i-particles and each take every G-th block of UNROLL j's of a
// tile; their partial sums are combined by shuffles before the write.
template <int R, int G, bool BULK, class Smem>
__device__ __forceinline__ void warp_force(const PassParams &p, const WorkItem it, Smem &s)
{
    const int lane = threadIdx.x & 31;
    const int w = it.walk;
    const int ibase = p.epi_off[w] + it.i0;
    const EpiAos *epi0 = p.epi + p.epi_off[w];
    // FP64 origin of the group: epi[0].pos (gravity_kernel_epep.pikg:53 "xi - xi[0]")
    const double ox = epi0->pos[0], oy = epi0->pos[1], oz = epi0->pos[2];
    const float eps2 = p.eps2;
    const int nj_ep = p.n_epj[w], nj_sp = p.n_spj[w];
    const int *adr_ep = p.adr_epj + p.epj_disp[w];
    const int *adr_sp = p.adr_spj + p.spj_disp[w];
    const int nt_ep = (nj_ep + JW - 1) / JW, nt_sp = (nj_sp + JW - 1) / JW;
    const int nt = nt_ep + nt_sp;
    // this item's share of the walk's tile sequence (items.h: a split tile's parts cover [0, nt) between them)
    const int t_begin = it.t0;
    const int t_end = it.t1 < 0 ? nt : min(it.t1, nt);
    if (it.cfg & ITEM_PEER_WAIT) peer_flags_wait(p.peer_flags, p.peer_world, p.peer_epoch);

    // list index of slot (lane + 32k) of tile t, or -1 for padding
    auto slot_index = [&](int t, int k) -> int {
        if (t >= t_end) return -1;
        if (t < nt_ep) { const int j = t * JW + lane + 32 * k; return j < nj_ep ? adr_ep[j] : -1; }
        if (t < nt) { const int j = (t - nt_ep) * JW + lane + 32 * k; return j < nj_sp ? adr_sp[j] : -1; }
        return -1;
    };
    constexpr bool bulk = BULK;               // compile-time: the cp.async build carries none of the bulk-copy state
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s.mbar);
    const unsigned raw_u32 = (unsigned)__cvta_generic_to_shared(&s.raw[0]);
    unsigned mphase = 0;                     // parity of the mbarrier phase the next wait completes
    if (bulk) {
        if (lane == 0) { mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
    }
    // stage tile t: slot (lane + 32 k) of the landing zone receives record idx_k (records are 48 B for EP, 64 B for SP;
    // bulk mode packs them at their own size, cp.async mode at a 64 B stride).  Returns true if anything is in flight.
    auto issue_tile = [&](int t, int i0, int i1) -> bool {
        const bool ep = t < nt_ep;
        // EP lists are runs of consecutive indices in tree order (measured on the N = 1e6 disk: 3.9 runs per 64-slot
        // tile) -> bulk copies; SP lists name scattered cells (57 runs per tile) -> per-record cp.async
        if (!bulk || !ep) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int idx = k ? i1 : i0;
                if (idx < 0) continue;
                float4 *dst = &s.raw[(lane + 32 * k) * 4];
                if (ep) {
                    const EpjPacked *rec = p.peer_epj ? p.peer_epj[idx >> p.peer_shift] + (idx & ((1 << p.peer_shift) - 1)) : p.epj + idx;
                    const char *q = reinterpret_cast<const char *>(rec);
                    cp_async16(dst + 0, q); cp_async16(dst + 1, q + 16); cp_async16(dst + 2, q + 32);
                } else {
                    const char *q = reinterpret_cast<const char *>(p.spj + idx);
                    cp_async16(dst + 0, q); cp_async16(dst + 1, q + 16); cp_async16(dst + 2, q + 32); cp_async16(dst + 3, q + 48);
                }
            }
            return true;
        }
        const unsigned rec = (unsigned)sizeof(EpjPacked);
        const char *base = reinterpret_cast<const char *>(p.epj);
        const unsigned v0 = __ballot_sync(0xffffffffu, i0 >= 0), v1 = __ballot_sync(0xffffffffu, i1 >= 0);
        const unsigned n_valid = __popc(v0) + __popc(v1);
        if (n_valid == 0) return false;
        if (lane == 0) {
            fence_proxy_async_smem();        // the landing zone was last read through the generic proxy (convert)
            mbar_expect_tx(mbar, n_valid * rec);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int idx = k ? i1 : i0;
            const int prev = __shfl_up_sync(0xffffffffu, idx, 1);
            const bool start = idx >= 0 && (lane == 0 || idx != prev + 1);
            const unsigned sm = __ballot_sync(0xffffffffu, start);
            if (start) {                                     // this lane's record opens a run of consecutive indices
                const unsigned after = lane == 31 ? 0u : (sm >> (lane + 1));
                const int next = after ? lane + __ffs(after) : 32;
                const int len = min(next, __popc(k ? v1 : v0)) - lane;       // valid slots are a prefix of each half
                bulk_g2s(raw_u32 + (unsigned)(lane + 32 * k) * rec, base + (size_t)idx * rec, (unsigned)len * rec, mbar);
            }
        }
        return true;
    };
    auto wait_tile = [&](int t, bool in_flight) {         // t = the tile that was staged
        if (!bulk || t >= nt_ep) { cp_async_commit_wait_all(); return; }
        if (in_flight) { mbar_wait(mbar, mphase); mphase ^= 1u; }
    };
    // raw record -> FP32 tile entry; returns rs2 of the entry (EP) for the tile's candidate threshold
    auto convert = [&](int t, int k, int idx) -> float {
        const int sl = lane + 32 * k;
        const bool ep = t < nt_ep;
        const float4 *src = (bulk && ep) ? reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(&s.raw[0]) + sl * 48) : &s.raw[sl * 4];
        if (idx < 0) {                        // padding: massless, far away, never a candidate
            s.j4[sl] = make_float4(1.0e10f, 1.0e10f, 1.0e10f, 0.0f);
            if (ep) { s.ep.rout2[sl] = 0.0f; s.ep.rs2[sl] = 0.0f; s.ep.id[sl] = -1; s.ep.rank[sl] = 0; s.ep.adr[sl] = -1; }
            else { s.sp.q0[sl] = make_float4(0.f, 0.f, 0.f, 0.f); s.sp.q1[sl] = make_float4(0.f, 0.f, 0.f, 0.f); }
            return 0.0f;
        }
        const double2 a = *reinterpret_cast<const double2 *>(src);
        const float4 bq = src[1];
        const float4 c = src[2];
        const double z = __hiloint2double(__float_as_int(bq.y), __float_as_int(bq.x));
        s.j4[sl] = make_float4((float)(a.x - ox), (float)(a.y - oy), (float)(z - oz), bq.z);
        if (ep) {
            s.ep.rout2[sl] = bq.w; s.ep.rs2[sl] = c.x;
            s.ep.id[sl] = __float_as_int(c.y); s.ep.rank[sl] = __float_as_int(c.z);
            s.ep.adr[sl] = idx;
            return c.x;
        }
        const float4 d = src[3];
        s.sp.q0[sl] = make_float4(bq.w, c.x, c.y, c.z);   // Qxx Qyy Qzz Qxy
        s.sp.q1[sl] = make_float4(c.w, d.x, d.y, 0.0f);   // Qyz Qzx mtr
        return 0.0f;
    };

    // ---- prologue: start tile 0, load this warp's i-particles meanwhile ----
    int idx0 = slot_index(t_begin, 0), idx1 = slot_index(t_begin, 1);
    bool in_flight = issue_tile(t_begin, idx0, idx1);
    int nidx0 = slot_index(t_begin + 1, 0), nidx1 = slot_index(t_begin + 1, 1);

    float xi[R], yi[R], zi[R], ro2i[R], rs2i[R];
    float ax[R], ay[R], az[R], ph[R];
    static_assert(G == 1 || R == 1, "j-split tiles hold one i-particle per lane");
    constexpr int W = 32 / G;                 // i-particles per lane group
    const int phase = lane / W;               // which j-blocks this lane takes (0 when G == 1)
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int il = lane + 32 * r;         // per-lane slot in the warp's shared arrays
        const int i = (G == 1) ? il : lane % W;
        if (i < it.ni) {
            const EpiAos &e = p.epi[ibase + i];
            const float ro = (float)e.r_out, rs = (float)e.r_search;
            xi[r] = (float)(e.pos[0] - ox); yi[r] = (float)(e.pos[1] - oy); zi[r] = (float)(e.pos[2] - oz);
            ro2i[r] = __fmul_rn(ro, ro);
            rs2i[r] = __fmul_rn(__fmul_rn(rs, rs), 1.0201f);
            s.i_id[il] = e.id_local; s.i_rank[il] = e.myrank;
        } else {
            xi[r] = yi[r] = zi[r] = -1.0e10f; ro2i[r] = 0.0f; rs2i[r] = -1.0f;   // far from everything: never a candidate
            s.i_id[il] = 0; s.i_rank[il] = 0;
        }
        s.i_rs2[il] = rs2i[r];
        s.nb_number[il] = 0; s.nb_rank[il] = 0; s.nb_idmax[il] = -1; s.nb_idmin[il] = 2147483647;
        ax[r] = ay[r] = az[r] = ph[r] = 0.0f;
    }
    float tmax = 0.0f;
    wait_tile(t_begin, in_flight);
    if (t_end > t_begin) tmax = fmaxf(convert(t_begin, 0, idx0), convert(t_begin, 1, idx1));
    __syncwarp();

    for (int t = t_begin; t < t_end; t++) {
        // stage tile t+1 while computing tile t
        idx0 = nidx0; idx1 = nidx1;
        in_flight = issue_tile(t + 1, idx0, idx1);
        nidx0 = slot_index(t + 2, 0); nidx1 = slot_index(t + 2, 1);
        const bool is_ep = t < nt_ep;
        const int n_t = is_ep ? min(JW, nj_ep - t * JW) : min(JW, nj_sp - (t - nt_ep) * JW);
        const int n_pad = (n_t + UNROLL * G - 1) / (UNROLL * G) * (UNROLL * G);   // every slot of the tile is initialised
        const int j_first = phase * UNROLL;
        if (is_ep) {
            // =============================== EP-EP ===============================
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            tmax *= 1.0000153f;           // conservative candidate threshold (exact test in the rare path)
            float T = -1.0f;                       // one threshold per lane: max over its i-slots and the tile's j
#pragma unroll
            for (int r = 0; r < R; r++) T = fmaxf(T, (rs2i[r] < 0.0f) ? -1.0f : fmaxf(rs2i[r] * 1.0000153f, tmax));
            unsigned int pend = 0u;                // blocks of UNROLL j of this tile in which this lane saw a candidate
#pragma unroll 1
            for (int jj = j_first; jj < n_pad; jj += UNROLL * G) {
                float rmin = 3.0e38f;
                float ro2a[UNROLL];
#pragma unroll
                for (int q = 0; q < UNROLL / 4; q++) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(&s.ep.rout2[jj + 4 * q]);
                    ro2a[4 * q] = v4.x; ro2a[4 * q + 1] = v4.y; ro2a[4 * q + 2] = v4.z; ro2a[4 * q + 3] = v4.w;
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const float4 pj = s.j4[jj + u];
                    const float ro2 = ro2a[u];
                    float r2[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float dx = pj.x - xi[r], dy = pj.y - yi[r], dz = pj.z - zi[r];
                        r2[r] = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, eps2)));
                        const float r2c = fmax3(r2[r], ro2i[r], ro2);
                        const float a = rinv_scaled(r2c);                  // RS*y
                        const float t = pj.w * a;                          // RS*m*y
                        const float v = (a * a) * t;                       // RS^3*m*y^3
                        ax[r] = fmaf(v, dx, ax[r]);
                        ay[r] = fmaf(v, dy, ay[r]);
                        az[r] = fmaf(v, dz, az[r]);
                        ph[r] -= t;
                    }
                    if (R == 1) rmin = fminf(rmin, r2[0]);
                    else if (R == 2) rmin = fmin3(rmin, r2[0], r2[1]);
                    else {
#pragma unroll
                        for (int r = 0; r < R; r++) rmin = fminf(rmin, r2[r]);
                    }
                }
                // candidate blocks are only noted here (one predicated OR); they are re-tested after the tile's
                // pair loop, every lane on ITS OWN blocks.  A warp-wide re-test at this point would run for every
                // block that holds one of the tile's own i-particles (each i meets itself at r2 == eps2): 16 warp-wide
                // re-tests per 64-wide item instead of the 2-3 per lane the deferred form needs.
                pend |= (rmin < T) ? (1u << (jj / UNROLL)) : 0u;
            }
            while (__any_sync(0xffffffffu, pend != 0u)) {
                if (pend != 0u) {
                    const int j0 = (__ffs(pend) - 1) * UNROLL;
                    pend &= pend - 1u;
                    // exact re-test, reference evaluation order, no FMA contraction
#pragma unroll 1
                    for (int u = 0; u < UNROLL; u++) {
                        const int j = j0 + u;
                        const float4 pj = s.j4[j];
                        const float rs2j = s.ep.rs2[j];
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const int il = lane + 32 * r;
                            const float dx = xi[r] - pj.x, dy = yi[r] - pj.y, dz = zi[r] - pj.z;
                            const float r2e = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)),
                                                                  __fmul_rn(dz, dz)), eps2);
                            const float rs2 = fmaxf(rs2i[r], rs2j);
                            if (rs2i[r] >= 0.0f && r2e < rs2) {
                                const int idj = s.ep.id[j], rkj = s.ep.rank[j];
                                const int idi = s.i_id[il], rki = s.i_rank[il];
                                const int gi = ibase + ((G == 1) ? il : lane % W);
                                if (idi != idj || rki != rkj) {     // only this lane touches entry il
                                    const int dr = rki - rkj;
                                    s.nb_number[il] += 1;
                                    s.nb_rank[il] += p.rank_squared ? dr * dr : abs(dr);
                                    s.nb_idmax[il] = max(s.nb_idmax[il], idj);
                                    s.nb_idmin[il] = min(s.nb_idmin[il], idj);
                                    if (p.pairs) {
                                        const unsigned int k = atomicAdd(p.pair_count, 1u);
                                        if (k < p.pair_cap) p.pairs[k] = make_int2(gi, s.ep.adr[j]);
                                    }
                                } else if (p.self_adr) {
                                    p.self_adr[gi] = s.ep.adr[j];
                                }
                            }
                        }
                    }
                }
            }
        } else {
            // =============================== EP-SP ===============================
#pragma unroll 1
            for (int jj = j_first; jj < n_pad; jj += UNROLL * G) {
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const float4 pj = s.j4[jj + u];
                    const float4 qa = s.sp.q0[jj + u];
                    const float4 qb = s.sp.q1[jj + u];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float dx = pj.x - xi[r], dy = pj.y - yi[r], dz = pj.z - zi[r];
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, eps2)));
                        const float a = rinv_scaled(r2);                               // RS*y
                        const float a2 = a * a, a3 = a2 * a, a4 = a2 * a2, a5 = a2 * a3;
                        const float qrx = fmaf(qb.y, dz, fmaf(qa.w, dy, qa.x * dx));   // (Qxx dx + Qxy dy + Qzx dz)/4
                        const float qry = fmaf(qa.w, dx, fmaf(qb.x, dz, qa.y * dy));   // (Qyy dy + Qyz dz + Qxy dx)/4
                        const float qrz = fmaf(qb.x, dy, fmaf(qb.y, dx, qa.z * dz));   // (Qzz dz + Qzx dx + Qyz dy)/4
                        const float rqr = fmaf(qrz, dz, fmaf(qry, dy, fmaf(qrx, dx, qb.z)));
                        const float wq = rqr * a4;                                      // RS^2 * rqr*y^4
                        const float meff = fmaf(0.5f * RS2_INV, wq, pj.w);              // m + 0.5 rqr y^4
                        const float meff3 = fmaf(2.5f * RS2_INV, wq, pj.w) * a3;        // RS^3 (m + 2.5 rqr y^4) y^3
                        ph[r] = fmaf(-meff, a, ph[r]);
                        ax[r] = fmaf(meff3, dx, fmaf(-a5, qrx, ax[r]));
                        ay[r] = fmaf(meff3, dy, fmaf(-a5, qry, ay[r]));
                        az[r] = fmaf(meff3, dz, fmaf(-a5, qrz, az[r]));
                    }
                }
            }
        }
        // tile t+1 has landed in raw (it had the whole compute phase to do so): convert it in place
        wait_tile(t + 1, in_flight);
        __syncwarp();                     // every lane is done reading tile t
        tmax = 0.0f;
        if (t + 1 < t_end) tmax = fmaxf(convert(t + 1, 0, idx0), convert(t + 1, 1, idx1));
        __syncwarp();
    }

    // ---- write-back: ForceGrav::clear + this pass's sums (a4 + a7 fused) ----
    __syncwarp();
    float4 fo[R]; int4 no[R]; bool ok[R];           // this warp's result per i-slot (lane + 32 r), unscaled
    if (G > 1) {
        // combine the lane groups' partial results (same i in lanes l, l+W, l+2W, ...)
        int nn = s.nb_number[lane], nr = s.nb_rank[lane], nmax = s.nb_idmax[lane], nmin = s.nb_idmin[lane];
#pragma unroll
        for (int o = W; o < 32; o <<= 1) {
            ax[0] += __shfl_xor_sync(0xffffffffu, ax[0], o); ay[0] += __shfl_xor_sync(0xffffffffu, ay[0], o);
            az[0] += __shfl_xor_sync(0xffffffffu, az[0], o); ph[0] += __shfl_xor_sync(0xffffffffu, ph[0], o);
            nn += __shfl_xor_sync(0xffffffffu, nn, o); nr += __shfl_xor_sync(0xffffffffu, nr, o);
            nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o)); nmin = min(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
        }
        ok[0] = lane < W && lane < it.ni;
        fo[0] = make_float4(RS3_INV * ax[0], RS3_INV * ay[0], RS3_INV * az[0], RS_INV * ph[0]);   // undo the exact scales
        no[0] = make_int4(nn, nr, nmax, nmin);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            ok[r] = i < it.ni;
            fo[r] = make_float4(RS3_INV * ax[r], RS3_INV * ay[r], RS3_INV * az[r], RS_INV * ph[r]);
            no[r] = make_int4(s.nb_number[i], s.nb_rank[i], s.nb_idmax[i], s.nb_idmin[i]);
        }
    }
    const int K = item_parts(it.cfg);
    if (K <= 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (ok[r]) {
                float4 *out = reinterpret_cast<float4 *>(p.force + ibase + lane + 32 * r);
                out[0] = fo[r];
                reinterpret_cast<int4 *>(out)[1] = no[r];
            }
        }
    } else {
        // one of K parts of a tile: leave the partial sums in this part's scratch slot, count the arrival, and if
        // this warp is the last of the K, add the partial sums in part order (a fixed order: the result does not
        // depend on which warp finishes when) and write ForceGrav
        ForceAos *slot = p.scratch + (size_t)(it.slot0 + item_part_index(it.cfg)) * 64;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (ok[r]) {
                float4 *out = reinterpret_cast<float4 *>(slot + lane + 32 * r);
                __stcg(out, fo[r]);
                __stcg(reinterpret_cast<int4 *>(out) + 1, no[r]);
            }
        }
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(p.arrive + it.group, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == K - 1) {
            __threadfence();
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (ok[r]) {
                    const int i = lane + 32 * r;
                    const ForceAos *q = p.scratch + (size_t)it.slot0 * 64 + i;
                    float4 f = __ldcg(reinterpret_cast<const float4 *>(q));
                    int4 nb = __ldcg(reinterpret_cast<const int4 *>(q) + 1);
                    for (int k = 1; k < K; k++) {
                        q += 64;
                        const float4 g = __ldcg(reinterpret_cast<const float4 *>(q));
                        const int4 mb = __ldcg(reinterpret_cast<const int4 *>(q) + 1);
                        f.x += g.x; f.y += g.y; f.z += g.z; f.w += g.w;
                        nb.x += mb.x; nb.y += mb.y; nb.z = max(nb.z, mb.z); nb.w = min(nb.w, mb.w);
                    }
                    float4 *out = reinterpret_cast<float4 *>(p.force + ibase + i);
                    out[0] = f;
                    reinterpret_cast<int4 *>(out)[1] = nb;
                }
            }
            if (lane == 0) p.arrive[it.group] = 0;          // ready for the next pass over this work list
        }
    }
}

// work item = up to 32*RMAX i-particles of one walk, handled by one warp with R = cfg+1 register
// slots per lane.  RMAX = 2: 80 registers, 24 warps/SM.  RMAX = 4: fewer shared-memory reads and
// less loop overhead per pair at 16 warps/SM.
template <int RMAX, bool BULK>
__global__ void __launch_bounds__(WPB * 32, RMAX <= 2 ? GB_MINB2 : 4) force_pass_kernel(const PassParams p, int n_items)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps &smem[warp] in a uniform
    // register instead of re-deriving it from the thread id inside the pair loops (9 of 167 instructions of the
    // EP-EP hot path at the 80-register cap; profiles/r1_force_pass_analysis.md)
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    using Smem = WarpSmem<32 * RMAX>;
    Smem &s = reinterpret_cast<Smem *>(smem_raw)[wid];
    const int slot = blockIdx.x * WPB + wid;
    const int lane = threadIdx.x & 31;
    int item = slot, item_end = min(slot + 1, n_items);
    if (p.seg_off) {
        if (slot >= p.n_seg) return;
        item = p.seg_off[slot]; item_end = p.seg_off[slot + 1];
    }
    if (p.fp_slab) {
        const int nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
        for (int i = tid; i < p.fp_n_epj; i += nthreads) {
            const EpjAos &a = p.fp_epj_in[i];
            p.fp_slab[i] = epj_pack(a.pos, a.mass, a.r_out, a.r_search, a.id_local, a.myrank);
        }
        const size_t ssz = p.fp_quad ? sizeof(SpjQuadAos) : sizeof(SpjMonoAos);
        for (int i = tid; i < p.fp_n_spj; i += nthreads)
            p.fp_spj_out[i] = spj_pack(static_cast<const char *>(p.fp_spj_in) + (size_t)i * ssz, p.fp_quad, p.fp_trace, p.eps2);
        __threadfence();                                   // records visible in this GPU's L2, where the peers' loads arrive
        __syncthreads();
        if (threadIdx.x == 0) {
            int *bar = p.place + p.place_bins + 2;
            const int old = atomicAdd(bar, 1);
            if (old == (int)gridDim.x - 1) {               // the slab is complete: tell every rank
                __threadfence_system();
                for (int q = 0; q < p.peer_world; q++) {
                    int *f = reinterpret_cast<int *>(static_cast<char *>(p.fp_slab0_of[q]) + p.fp_flag_off) + p.fp_rank;
                    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(p.peer_epoch) : "memory");
                }
            }
            int seen;
            do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); } while (seen < (int)gridDim.x);
        }
        __syncthreads();
    }
    // placed pass: this warp's bin, and where its search for unclaimed entries stands (-1: own bin not yet tried)
    int my_bin = 0, scan = -1;
    bool first_claim = true;
    if (p.place) {
        unsigned int smid, warpid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %warpid;" : "=r"(warpid));
        my_bin = (int)((smid * 4u + (warpid & 3u)) % (unsigned int)p.place_bins);
        item = 0; item_end = 0;
    }
    for (;;) {
        if (p.place) {
            // next unclaimed entry: first this warp's own bin, then -- only if items are still unclaimed once this warp
            // has nothing left, i.e. the hardware did not spread the CTAs evenly -- any bin, 32 counters at a time
            int found = -1;
            if (scan < 0) {
                scan = 0;
                int k = 0;
                if (lane == 0) k = atomicAdd(p.place + my_bin, 1);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k < p.place_rounds && place_item(k, my_bin, p.place_bins) < n_items) {
                    found = place_item(k, my_bin, p.place_bins);
                    scan = -1;                      // more rounds than resident warps: come back to this bin afterwards
                } else if (first_claim) {
                    __nanosleep(2000);              // no entry here: let the rightful owners claim theirs before looking around
                }
                first_claim = false;
            }
            while (found < 0) {
                int claimed = 0;
                if (lane == 0) claimed = *reinterpret_cast<volatile int *>(p.place + p.place_bins);
                claimed = __shfl_sync(0xffffffffu, claimed, 0);
                if (claimed >= n_items || scan >= p.place_bins) break;
                const int b = (my_bin + 1 + scan + lane) % p.place_bins;
                int k = p.place_rounds;
                if (scan + lane < p.place_bins) k = *reinterpret_cast<volatile int *>(p.place + b);
                const bool open = k < p.place_rounds && place_item(k, b, p.place_bins) < n_items;
                const unsigned int m = __ballot_sync(0xffffffffu, open);
                if (m == 0u) { scan += 32; continue; }
                const int src = __ffs(m) - 1;
                const int bb = __shfl_sync(0xffffffffu, b, src);
                int kk = 0;
                if (lane == 0) kk = atomicAdd(p.place + bb, 1);
                kk = __shfl_sync(0xffffffffu, kk, 0);
                if (kk < p.place_rounds && place_item(kk, bb, p.place_bins) < n_items) found = place_item(kk, bb, p.place_bins);
                // else: someone else took it; look at the same 32 bins again
            }
            if (found < 0) break;
            if (lane == 0) atomicAdd(p.place + p.place_bins, 1);
            item = found;
        } else {
            if (item >= item_end) break;
        }
        const WorkItem it = p.items[item];
        unsigned long long tr0 = 0;
        if (p.trace) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr0));
        if (it.ni == 0) {           // barrier item of the multi-GPU peer mode: the pass ends after every peer has packed
            if (it.cfg & ITEM_PEER_WAIT) peer_flags_wait(p.peer_flags, p.peer_world, p.peer_epoch);
            item++;
            continue;
        }
        // cfg & 15: 0..3 = R-1 register slots per lane (full-width tiles); 8+k = lane-split tile, G = 2^k lane groups;
        // bit 6, bits 8-23: items.h (peer wait, j-split part k of K)
        const int kcfg = it.cfg & 15;
        if (kcfg >= 8) {
            switch (kcfg) {
                case 9: warp_force<1, 2, BULK>(p, it, s); break;
                case 10: warp_force<1, 4, BULK>(p, it, s); break;
                default: warp_force<1, 8, BULK>(p, it, s); break;
            }
        } else if (RMAX <= 2) {
            if (kcfg == 0) warp_force<1, 1, BULK>(p, it, s);
            else warp_force<2, 1, BULK>(p, it, s);
        } else {
            switch (kcfg) {
                case 0: warp_force<1, 1, BULK>(p, it, s); break;
                case 1: warp_force<2, 1, BULK>(p, it, s); break;
                case 2: warp_force<3, 1, BULK>(p, it, s); break;
                default: warp_force<4, 1, BULK>(p, it, s); break;
            }
        }
        __syncwarp();               // the next item reuses this warp's shared-memory arrays
        if (p.trace && (threadIdx.x & 31) == 0) {
            unsigned long long tr1; unsigned int smid, warpid;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr1));
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            asm volatile("mov.u32 %0, %warpid;" : "=r"(warpid));
            p.trace[4 * (size_t)item] = tr0; p.trace[4 * (size_t)item + 1] = tr1;
            p.trace[4 * (size_t)item + 2] = (unsigned long long)smid | ((unsigned long long)warpid << 32);
            atomicAdd(p.trace + 4 * (size_t)item + 3, 1ull);
        }
        item++;
    }
    if (p.place) {
        // the last warp to leave zeroes the counters for the next launch (every other warp's last access is its
        // arrival here)
        int old = 0;
        if (lane == 0) old = atomicAdd(p.place + p.place_bins + 1, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == (int)(gridDim.x * WPB) - 1) {
            __threadfence();
            for (int k = lane; k < p.place_bins + 3; k += 32) p.place[k] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// FP32 FMA issue-rate microbenchmark: the roofline denominator, measured on the same clocks.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b)
{
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = fmaf(v[k], a, b);
    }
    float sacc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) sacc += v[k];
    if (sacc == 123.456f) out[0] = sacc;
}

}  // namespace gb
