// soft_corr.cu -- the changeover (cutoff) correction of the soft force and the final neighbour
// lists on the device: GPLUM's correctForceLong / correctForceLongInitial
// (src/gravity_soft.h:76-153,155-242,245-372,375-528) with the cutoff functions of src/cutfunc.h.
//
// The reference, per particle with neighbour candidates, asks the tree for every EPJ inside the
// search ball (or takes NeighborInfo::id_min/id_max) and runs correctForceBetween2Particles on
// each.  Here the force pass itself has already written every (i, j) that passed its candidate
// test into a pair buffer (kernels.cuh, rare path), so no second search exists:
//   count  : cnt[i] = ForceGrav::number  -> exclusive scan -> off[]        (segments per i)
//   scatter: pair k = (i, j)             -> csr[off[i] + cursor[i]++] = j
//   apply  : one thread per i: sort its (few) j ascending -- the pair buffer's order is not
//            deterministic, the sums must be --, then the reference's FP64 arithmetic per j.
// Compiled with -fmad=false: the reference build has no FMA contraction (x86-64 -O2), and several
// terms cancel (rinv*W - r_min, r3inv*K - r_min^3).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <stdint.h>

#include "soft_corr.h"

namespace gb {

namespace {

struct Cut { double g, g_1_inv, g_1_inv7, w_y, f1; };

// std::max / std::min of the reference (argument order matters for NaN)
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }

// src/cutfunc.h:4-14
__device__ double cutoff_f(double y, const Cut &c)
{
    const double g = c.g, g2 = g * g;
    return (((((((-10. / 3. * y + 14. * (g + 1.)) * y - 21. * ((g + 3.) * g + 1.)) * y
                + 35. / 3. * (((g + 9.) * g + 9.) * g + 1.)) * y
               - 70. * ((g + 3.) * g + 1.) * g) * y
              + 210. * (g + 1.) * g2) * y - 140. * g2 * g * log(y)) * y
            + (((g - 7.) * g + 21.) * g - 35.) * g2 * g2) * c.g_1_inv7;
}
// src/cutfunc.h:17-29
__device__ double cutoff_W(double rij, double r_out_inv, const Cut &c)
{
    const double y = rij * r_out_inv;
    if (1.0 <= y) return 1.0;
    if (y <= c.g) return y * c.w_y;
    return cutoff_f(y, c) + y * (1. - c.f1);
}
// src/cutfunc.h:31-40,59-62
__device__ double cutoff_K(double rij, double r_out_inv, const Cut &c)
{
    const double x = (c.g - rij * r_out_inv) * c.g_1_inv;
    if (x < 0.) return 0.;
    if (x >= 1.) return 1.;
    const double x2 = x * x;
    return (((-20. * x + 70.) * x - 84.) * x + 35.) * x2 * x2;
}
// src/cutfunc.h:41-46,63-66
__device__ double cutoff_dKdt(double rij, double r_out_inv, double alpha, const Cut &c)
{
    const double x = (c.g - rij * r_out_inv) * c.g_1_inv;
    const double x_1 = x - 1.;
    const double dKdr = (x < 0. || x >= 1.) ? 0. : (140. * x * x * x * x_1 * x_1 * x_1 * r_out_inv * c.g_1_inv);
    return alpha * rij * dKdr;
}

struct EpjAos { int id_local, myrank; double pos[3]; double r_out, r_search; long long id;
                double mass; double vel[3]; double acc_d[3]; };                                    // 112
struct EpiAos { int id_local, myrank; double pos[3]; double r_out, r_search; };                    // 48
struct ForceAos { float acc[3]; float phi; int number, rank, id_max, id_min; };                    // 32
static_assert(sizeof(EpjAos) == 112 && sizeof(EpiAos) == 48 && sizeof(ForceAos) == 32, "reference layout");

__global__ void corr_count_kernel(const ForceAos *__restrict__ force, int n, int i0, int i1, int *__restrict__ cnt, int *__restrict__ cursor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) cnt[i] = (i >= i0 && i < i1) ? force[i].number : 0;
    if (i < n) cursor[i] = 0;
}

// status[0] = pairs dropped because the buffer was too small, status[1] = i-particles that are not
// in their own EP list, status[2] = total neighbours written (for the caller's D2H size)
__global__ void corr_scatter_kernel(const int2 *__restrict__ pairs, const unsigned int *__restrict__ pair_count,
                                    unsigned int pair_cap, const int *__restrict__ off, int *__restrict__ cursor,
                                    int *__restrict__ csr, unsigned int *__restrict__ status)
{
    const unsigned int total = *pair_count;
    // Overflow: off[] comes from ForceGrav::number (the true pair count) while csr / ngb hold pair_cap entries,
    // so nothing may be scattered or applied; the host reports GPLUM_B200_ERR_OVERFLOW from status[0] and the
    // caller enlarges the buffer and repeats the pass.
    if (total > pair_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = total - pair_cap;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int2 pr = pairs[k];
        const int slot = atomicAdd(&cursor[pr.x], 1);
        const long long at = (long long)off[pr.x] + slot;
        if (at >= 0 && at < (long long)pair_cap) csr[at] = pr.y;     // counts and pairs always agree for FDPS lists
        else atomicAdd(&status[0], 1u);
    }
}

__global__ void __launch_bounds__(128) corr_apply_kernel(SoftCorrArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_epi) return;
    Cut cut;
    {   // FPGrav::setGamma, src/particle.h:619-633
        const double g = a.prm.gamma;
        cut.g = g;
        cut.g_1_inv = 1. / (g - 1.);
        const double g2 = g * g;
        const double g_1_inv3 = cut.g_1_inv * cut.g_1_inv * cut.g_1_inv;
        cut.g_1_inv7 = g_1_inv3 * g_1_inv3 * cut.g_1_inv;
        cut.w_y = 7. / 3. * ((((((g - 9.) * g + 45.) * g - 60. * log(g)) * g - 45.) * g + 9.) * g - 1.) * cut.g_1_inv7;
        cut.f1 = (-10. / 3. + 14. * (g + 1.) - 21. * ((g + 3.) * g + 1.)
                  + 35. / 3. * (((g + 9.) * g + 9.) * g + 1.)
                  - 70. * ((g + 3.) * g + 1.) * g
                  + 210. * (g + 1.) * g2
                  + (((g - 7.) * g + 21.) * g - 35.) * g2 * g2) * cut.g_1_inv7;
    }
    const EpjAos *epj = (const EpjAos *)a.epj_aos;
    const int base = a.off[i];
    const bool overflow = *a.pair_count > a.pair_cap || a.off[a.n_epi] < 0 || (unsigned int)a.off[a.n_epi] > a.pair_cap;
    const int n_cand = overflow ? 0 : min(a.cursor[i], a.off[i + 1] - base);   // == ForceGrav::number
    SoftCorr out;
    out.id_local = ((const EpiAos *)a.epi)[i].id_local;
    out.ngb_off = overflow ? 0 : base;
    const bool mine = i >= a.i0 && i < a.i1;
    const int sa = mine ? a.self_adr[i] : -1;
    if (sa < 0 || overflow) {   // sa < 0 cannot happen with FDPS lists (a group's own particles are in its EP list);
                                // overflow: neutral record, nothing outside the buffers is touched (scatter kernel)
        if (!overflow && mine) atomicAdd(&a.status[1], 1u);
        out.acc[0] = out.acc[1] = out.acc[2] = 0.; out.phi = 0.; out.acc0 = 0.;
        out.id_cluster = -1; out.number = 0; out.in_domain = 1;
        a.out[i] = out;
        if (a.init_out) { SoftCorrInit ci = {}; a.init_out[i] = ci; }
        return;
    }
    const EpjAos self = epj[sa];
    double acci[3] = {0., 0., 0.}, phii = 0., acc0i = 0.;
    double acc_di[3] = {0., 0., 0.}, jerki[3] = {0., 0., 0.}, phi_di = 0.;
    long long id_cluster = self.id;
    int number = 0, in_domain = 1;
    const double eps2 = a.prm.eps2;
    const double r_out_inv_i = 1. / self.r_out;               // src/particle.h:728
    phii += self.mass * r_out_inv_i;                          // src/gravity_soft.h:281,293
    // deterministic order: ascending EP index (insertion sort; segments are a handful of entries)
    int *seg = a.csr + base;
    for (int k = 1; k < n_cand; k++) {
        const int v = seg[k];
        int m = k - 1;
        while (m >= 0 && seg[m] > v) { seg[m + 1] = seg[m]; m--; }
        seg[m + 1] = v;
    }
    // An entry with the particle's own id is the absorbed twin of a merger (src/collisionA.h:267-277).  The reference
    // adds massj * r_out_inv to phi for it only on its branch for at most two candidates, all on this rank
    // (src/gravity_soft.h:295-317, 105-108); its tree-search branch skips such entries.
    const bool twin_phi = n_cand <= 2 && ((const ForceAos *)a.force)[i].rank == 0;
    for (int k = 0; k < n_cand; k++) {
        const EpjAos q = epj[seg[k]];
        // ---- correctForceBetween2Particles{,Initial} (src/gravity_soft.h:76-153,155-242) ----
        const double massj = q.mass;
        const double r_out = std_max(self.r_out, q.r_out);
        const double r_out_inv = std_min(r_out_inv_i, 1. / q.r_out);
        const double r_search = std_max(self.r_search, q.r_search);
        if (q.id == self.id) { if (twin_phi) phii += massj * r_out_inv; continue; }
        const double dr[3] = {q.pos[0] - self.pos[0], q.pos[1] - self.pos[1], q.pos[2] - self.pos[2]};
        double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        dr2 += eps2;
        const double rij = sqrt(dr2);
        const double dv[3] = {q.vel[0] - self.vel[0], q.vel[1] - self.vel[1], q.vel[2] - self.vel[2]};
        const double drdv = dr[0] * dv[0] + dr[1] * dv[1] + dr[2] * dv[2];
        bool pass = true;
        if (a.prm.re_search) {                                // USE_RE_SEARCH_NEIGHBOR, src/main_p3t.cpp:15
            const double da[3] = {q.acc_d[0] - self.acc_d[0], q.acc_d[1] - self.acc_d[1], q.acc_d[2] - self.acc_d[2]};
            const double dv2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
            const double da2 = da[0] * da[0] + da[1] * da[1] + da[2] * da[2];
            const double t_min = std_min(std_max(-drdv / sqrt(dv2), 0.), a.prm.dt_tree);
            const double dr2_min = std_min(dr2, dr2 + 2. * drdv * t_min + dv2 * t_min * t_min);
            const double r_crit = a.prm.R_search2 * r_out;
            const double v_crit_a = a.prm.R_search3 * 0.5 * a.prm.dt_tree;
            pass = (dr2_min < r_crit * r_crit) || (dv2 < v_crit_a * v_crit_a * da2) || (a.prm.initial && da2 == 0.);
        }
        if (pass && rij < r_search) {                         // NeighborList::addNeighbor, src/neighbor.h:635-664
            SoftNgb nb; nb.id = q.id; nb.rank = q.myrank; nb.id_local = q.id_local;
            a.ngb[base + number] = nb;
            number++;
            if (q.id < id_cluster) id_cluster = q.id;
            if (q.myrank != self.myrank) in_domain = 0;
            acc0i += r_out * r_out / massj;
        }
        if (rij < r_out) {
            const double rinv = 1. / rij, r2inv = rinv * rinv, r3inv = rinv * r2inv;
            const double W = cutoff_W(rij, r_out_inv, cut);
            const double K = cutoff_K(rij, r_out_inv, cut);
            const double r_min = std_min(rinv, r_out_inv);
            phii -= massj * (rinv * W - r_min);
            const double ca = massj * (r3inv * K - r_min * r_min * r_min);
            acci[0] += ca * dr[0]; acci[1] += ca * dr[1]; acci[2] += ca * dr[2];
            if (a.prm.initial) {
                const double alpha = drdv * r2inv;
                const double dKdt = cutoff_dKdt(rij, r_out_inv, alpha, cut);
                const double alpha_c = alpha * (1. - K);
                phi_di -= massj * rinv * (1. - W);
                const double cd = massj * r3inv * (1. - K);
                const double cj = massj * r3inv;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    acc_di[d] += cd * dr[d];
                    jerki[d] += cj * ((1. - K) * dv[d] - (3. * alpha_c + dKdt) * dr[d]);
                }
            }
        }
    }
    out.acc[0] = acci[0]; out.acc[1] = acci[1]; out.acc[2] = acci[2];
    out.phi = phii;
    out.acc0 = (acc0i > 0.) ? number / acc0i : 0.;            // src/gravity_soft.h:368
    out.id_cluster = id_cluster; out.number = number; out.in_domain = in_domain;
    a.out[i] = out;
    if (a.init_out) {
        SoftCorrInit ci;
#pragma unroll
        for (int d = 0; d < 3; d++) { ci.acc_d[d] = acc_di[d]; ci.jerk_d[d] = jerki[d]; }
        ci.phi_d = phi_di; ci.pad = 0.;
        a.init_out[i] = ci;
    }
}

}  // namespace

size_t soft_corr_scan_temp_bytes(int n_epi)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, n_epi + 1);
    return bytes;
}

// Enqueues the three steps on `st`; returns the CUDA error of the launches (0 = ok) and adds the
// number of kernels launched to *launches.
int soft_corr_launch(const SoftCorrArgs &a, void *scan_temp, size_t scan_temp_bytes, cudaStream_t st, int *launches)
{
    if (a.n_epi <= 0) return 0;
    const int n1 = a.n_epi + 1;
    corr_count_kernel<<<(n1 + 255) / 256, 256, 0, st>>>((const ForceAos *)a.force, a.n_epi, a.i0, a.i1, a.cnt, a.cursor);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(scan_temp, scan_temp_bytes, (const int *)a.cnt, a.off, n1, st);
    if (e != cudaSuccess) return (int)e;
    corr_scatter_kernel<<<148 * 4, 256, 0, st>>>(a.pairs, a.pair_count, a.pair_cap, a.off, a.cursor, a.csr, a.status);
    corr_apply_kernel<<<(a.n_epi + 127) / 128, 128, 0, st>>>(a);
    if (launches) *launches += 4;      // count, scan (CUB, >= 1 kernel), scatter, apply
    return (int)cudaGetLastError();
}

// ---- compact result: only the particles that have neighbours (the others carry the self term alone:
// acc = 0, phi = m / r_out, acc0 = 0, id_cluster = id, number = 0), in walk order (stable select) ----
struct HasNeighbour { __host__ __device__ bool operator()(const SoftCorr &c) const { return c.number > 0; } };

size_t soft_corr_compact_temp_bytes(int n_epi)
{
    size_t bytes = 0;
    cub::DeviceSelect::If(nullptr, bytes, (const SoftCorr *)nullptr, (SoftCorr *)nullptr, (int *)nullptr, n_epi, HasNeighbour());
    return bytes;
}

int soft_corr_compact(int n_epi, const SoftCorr *corr, SoftCorr *out, int *n_out_dev, void *temp, size_t temp_bytes, cudaStream_t st)
{
    if (n_epi <= 0) return 0;
    return (int)cub::DeviceSelect::If(temp, temp_bytes, corr, out, n_out_dev, n_epi, HasNeighbour(), st);
}

}  // namespace gb
