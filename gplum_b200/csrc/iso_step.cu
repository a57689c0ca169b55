// iso_step.cu -- the isolated-particle half of a GPLUM soft step on a device-resident particle state
// (SURVEY 8 f3): velocity kick and Kepler drift.  FP64 throughout, compiled with -fmad=false so that
// every product/sum rounds as in the reference's x86-64 build; only sin/cos/atan2 come from a
// different libm (CUDA's, <= 2 ulp), which bounds the difference to the reference at ~1e-15 relative.
//
//   FPGrav::velKick                      src/particle.h:878-884     vel += 0.5*dt_tree*acc
//   acc = (F64)ForceGrav::acc + acci     src/particle.h:761-766, src/gravity_soft.h:366-367
//   which particles drift on a Kepler orbit   src/hard.h:793-803    (!neighbor.number && ecc < 0.8 && eps2_sun == 0)
//   FPGrav::getEccentricity              src/particle.h:668-685
//   timeIntegrateKepler_isolated         src/hermite.h:787-816
//   KeplerEq / solveKeplerEq / posVel2OrbitalElement / orbitalElement2PosVel   src/kepler.h:3-97
//   calcStarGravity                      src/gravity_hard.h:5-39
//   calcDt2nd, FPGrav::calcDeltatInitial src/particle.h:345-355,886-914
// HBM-bound elementwise kernels: one thread per particle, 112 B state record read + 48 B written.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "iso_step.h"
#include "records.h"

namespace gbi {

using gb::EpiAos;
using gb::EpjAos;
using gb::ForceAos;
typedef gplum_b200_corr SoftCorr;
typedef gplum_b200_star Star;
static_assert(sizeof(SoftCorr) == 64 && sizeof(Star) == 64, "layout");

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ double dot3(const double *a, const double *b) { return (a[0] * b[0]) + (a[1] * b[1]) + (a[2] * b[2]); }

__global__ void __launch_bounds__(TPB) kick_kernel(int n, EpjAos *__restrict__ state, const EpiAos *__restrict__ epi,
                                                   const ForceAos *__restrict__ force, const SoftCorr *__restrict__ corr,
                                                   double half_dt)
{
    const int t = blockIdx.x * TPB + threadIdx.x;
    if (t >= n) return;
    const ForceAos f = force[t];
    double a[3] = {(double)f.acc[0], (double)f.acc[1], (double)f.acc[2]};
    int k;
    if (corr) { const SoftCorr c = corr[t]; k = c.id_local; a[0] += c.acc[0]; a[1] += c.acc[1]; a[2] += c.acc[2]; }
    else k = epi[t].id_local;
    EpjAos &p = state[k];
    for (int d = 0; d < 3; d++) p.vel[d] += half_dt * a[d];
}

__global__ void __launch_bounds__(TPB) flags_kernel(int n, const SoftCorr *__restrict__ corr, int *__restrict__ isolated,
                                                    double *__restrict__ acc0)
{
    const int t = blockIdx.x * TPB + threadIdx.x;
    if (t >= n) return;
    const SoftCorr c = corr[t];
    isolated[c.id_local] = c.number == 0 ? 1 : 0;
    acc0[c.id_local] = c.acc0;
}

__device__ __forceinline__ double kepler_eq(double u, double ecc) { return u - ecc * sin(u); }

__device__ double solve_kepler_eq(double l, double ecc)
{
    double u;
    const double ecc2 = ecc * ecc, ecc3 = ecc2 * ecc, ecc4 = ecc2 * ecc2, ecc5 = ecc3 * ecc2, ecc6 = ecc3 * ecc3;
    u = l
        + (ecc - ecc3 / 8. + ecc5 / 192.) * sin(l)
        + (ecc2 / 2. - ecc4 / 6. + ecc6 / 48.) * sin(2. * l)
        + (3. * ecc3 / 8. - 27. * ecc5 / 128.) * sin(3. * l)
        + (ecc4 / 3. - 4. * ecc6 / 15.) * sin(4. * l)
        + 125. * ecc5 / 384. * sin(5. * l)
        + 27. * ecc6 / 80. * sin(6. * l);
    if (fabs(kepler_eq(u, ecc) - l) > 1.e-15) {
        double u0;
        int loop = 0;
        do {
            u0 = u;
            double sinu0, cosu0;
            sincos(u0, &sinu0, &cosu0);
            u = u0 - ((u0 - ecc * sinu0 - l) / (1. - ecc * cosu0));
            loop++;
        } while (fabs(u - u0) > 1.e-15 && loop < 10);
    }
    return u;
}

__device__ __forceinline__ double calc_dt2nd(double eta, double alpha2, double acc0, const double *acc, const double *jerk)
{
    const double Acc2 = dot3(acc, acc) + alpha2 * acc0 * acc0;
    const double Jerk2 = dot3(jerk, jerk);
    return (Jerk2 > 0.) ? eta * sqrt(Acc2 / Jerk2) : DBL_MAX;
}

__global__ void __launch_bounds__(TPB) drift_kernel(int n, EpjAos *__restrict__ state, double *__restrict__ time_,
                                                    double *__restrict__ dt_, const double *__restrict__ acc0_,
                                                    const int *__restrict__ isolated, double t0, double t1,
                                                    gplum_b200_iso_params prm, Star *__restrict__ star, int *__restrict__ handled)
{
    const int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= n) return;
    int h = 0;
    if (isolated[k]) {
        EpjAos &p = state[k];
        double pos[3] = {p.pos[0], p.pos[1], p.pos[2]}, vel[3] = {p.vel[0], p.vel[1], p.vel[2]};
        const double mu = prm.m_sun;
        // FPGrav::getEccentricity
        const double r2 = dot3(pos, pos), r = sqrt(r2), v2 = dot3(vel, vel), rv = dot3(pos, vel);
        double ecc_test;
        {
            const double ax = 1.0 / (2.0 / r - v2 / mu);
            const double ecccosu = 1. - r / ax;
            const double eccsinu2 = rv * rv / (mu * ax);
            ecc_test = sqrt(ecccosu * ecccosu + eccsinu2);
        }
        if (ecc_test < 0.8 && prm.eps2_sun == 0.) {
            h = 1;
            // posVel2OrbitalElement
            const double rinv = 1. / r;
            const double ax = 1.0 / (2.0 * rinv - v2 / mu);
            const double ecccosu = 1. - r / ax;
            const double eccsinu = rv / sqrt(mu * ax);
            const double ecc = sqrt(ecccosu * ecccosu + eccsinu * eccsinu);
            const double nn = sqrt(mu / (ax * ax * ax));
            double u, cosu, sinu;
            if (ecc != 0) { u = atan2(eccsinu, ecccosu); cosu = ecccosu / ecc; sinu = eccsinu / ecc; }
            else { u = 0.; cosu = 1.; sinu = 0.; }
            const double aninv = sqrt(ax / mu);
            const double ecc_sq = sqrt(1. - ecc * ecc);
            double P[3], Q[3];
            {
                const double a = rinv * cosu, b = aninv * sinu, c = rinv * sinu, d = aninv * (cosu - ecc), inv = 1.0 / ecc_sq;
                for (int q = 0; q < 3; q++) { P[q] = pos[q] * a - vel[q] * b; Q[q] = (pos[q] * c + vel[q] * d) * inv; }
            }
            // advance the mean anomaly, solve Kepler's equation
            double l = kepler_eq(u, ecc);
            l += nn * (t1 - t0);
            u = solve_kepler_eq(l, ecc);
            // orbitalElement2PosVel
            {
                double cu, su;
                sincos(u, &su, &cu);
                const double a = cu - ecc, b = ecc_sq * su;
                for (int q = 0; q < 3; q++) pos[q] = (P[q] * a + Q[q] * b) * ax;
                const double rinv2 = sqrt(1. / dot3(pos, pos));
                const double s = ax * ax * nn * rinv2, c = -su, d = ecc_sq * cu;
                for (int q = 0; q < 3; q++) vel[q] = (P[q] * c + Q[q] * d) * s;
            }
            double tm = time_[k];
            tm += (t1 - t0);
            // calcStarGravity
            Star st;
            double dr[3], dv[3];
            for (int q = 0; q < 3; q++) { dr[q] = -pos[q]; dv[q] = -vel[q]; }
            const double r2inv = 1. / (dot3(dr, dr) + prm.eps2_sun);
            const double rinv3 = sqrt(r2inv), r3inv = rinv3 * r2inv;
            const double mj = mu * r3inv;
            const double alpha = dot3(dr, dv) * r2inv;
            st.phi_s = -mu * rinv3;
            for (int q = 0; q < 3; q++) { st.acc_s[q] = dr[q] * mj; st.jerk_s[q] = (dv[q] - dr[q] * (3. * alpha)) * mj; }
            // calcDeltatInitial with acc_d = jerk_d = 0
            const double zero[3] = {0., 0., 0.};
            double dt_next = 0.5 * prm.dt_tree;
            const double d1a = calc_dt2nd(prm.eta_0, prm.alpha2, acc0_[k], zero, zero);
            const double d1b = calc_dt2nd(prm.eta_sun0, prm.alpha2, 0., st.acc_s, st.jerk_s);
            const double dt_1 = (d1b < d1a) ? d1b : d1a;
            double rem = fmod(tm, dt_next);
            // the reference's loop has no bound: a NaN/Inf time (caller-supplied state) never satisfies it and
            // would hang the stream.  2^-1100 underflows to zero, so a finite time leaves within 1100 halvings;
            // anything else stays unhandled and goes to the host like a clustered particle.
            int halvings = 0;
            while (rem != 0.0 && halvings < 1100) { dt_next *= 0.5; rem = fmod(tm, dt_next); halvings++; }
            if (rem != 0.0) h = 0;
            else {
                const double dt_old = dt_[k];
                if (dt_old > 0.) while (2. * dt_old < dt_next) dt_next *= 0.5;
                while (dt_1 < dt_next) dt_next *= 0.5;
                if (dt_next < 2. * prm.dt_min) dt_next = prm.dt_min;
                st.dt = dt_next;
                for (int q = 0; q < 3; q++) { p.pos[q] = pos[q]; p.vel[q] = vel[q]; p.acc_d[q] = 0.; }
                time_[k] = tm; dt_[k] = dt_next;
                star[k] = st;
            }
        }
    }
    handled[k] = h;
}

__global__ void __launch_bounds__(TPB) pull_kernel(int n, const uint4 *__restrict__ state, const int *__restrict__ handled,
                                                   uint4 *__restrict__ rec, int *__restrict__ idx, int *count, int cap)
{
    const int k = blockIdx.x * TPB + threadIdx.x;
    const bool want = k < n && !handled[k];
    // one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!want) return;
    const int slot = base + __popc(m & ((1u << lane) - 1));
    if (slot >= cap) return;
    idx[slot] = k;
    for (int q = 0; q < 7; q++) rec[(size_t)slot * 7 + q] = state[(size_t)k * 7 + q];
}

__global__ void __launch_bounds__(TPB) push_kernel(int n_rec, const uint4 *__restrict__ rec, const int *__restrict__ idx, uint4 *__restrict__ state)
{
    const int t = blockIdx.x * TPB + threadIdx.x;
    const int s = t >> 3, q = t & 7;
    if (s >= n_rec || q >= 7) return;
    state[(size_t)idx[s] * 7 + q] = rec[(size_t)s * 7 + q];
}

inline int nblk(long long n) { return (int)((n + TPB - 1) / TPB); }

}  // namespace

#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

int iso_kick(int n, void *state, const void *epi, const void *force, const void *corr, double half_dt, cudaStream_t st)
{
    if (n <= 0) return 0;
    kick_kernel<<<nblk(n), TPB, 0, st>>>(n, (EpjAos *)state, (const EpiAos *)epi, (const ForceAos *)force, (const SoftCorr *)corr, half_dt);
    CKL();
    return 0;
}

int iso_flags_from_corr(int n, const void *corr, int *isolated, double *acc0, cudaStream_t st)
{
    if (n <= 0) return 0;
    flags_kernel<<<nblk(n), TPB, 0, st>>>(n, (const SoftCorr *)corr, isolated, acc0);
    CKL();
    return 0;
}

int iso_drift(int n, void *state, double *time, double *dt, const double *acc0, const int *isolated,
              double t0, double t1, const gplum_b200_iso_params &prm, void *star, int *handled, cudaStream_t st)
{
    if (n <= 0) return 0;
    drift_kernel<<<nblk(n), TPB, 0, st>>>(n, (EpjAos *)state, time, dt, acc0, isolated, t0, t1, prm, (Star *)star, handled);
    CKL();
    return 0;
}

int iso_pull_unhandled(int n, const void *state, const int *handled, void *rec_out, int *idx_out, int *count, int cap, cudaStream_t st)
{
    if (n <= 0) return 0;
    pull_kernel<<<nblk(n), TPB, 0, st>>>(n, (const uint4 *)state, handled, (uint4 *)rec_out, idx_out, count, cap);
    CKL();
    return 0;
}

int iso_push(int n_rec, const void *rec, const int *idx, void *state, cudaStream_t st)
{
    if (n_rec <= 0) return 0;
    push_kernel<<<nblk((long long)n_rec * 8), TPB, 0, st>>>(n_rec, (const uint4 *)rec, idx, (uint4 *)state);
    CKL();
    return 0;
}

}  // namespace gbi
