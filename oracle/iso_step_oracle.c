/*
 * iso_step_oracle.c -- CPU restatement of the isolated-particle half of a GPLUM soft step
 * (SURVEY 8 f3): the velocity kick and the Kepler drift of particles without neighbours.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product links or calls this file; it is the checker of
 * gplum_b200/csrc/iso_step.cu in tests/ and is itself pinned against the reference's own functions
 * compiled from /root/reference (oracle/ref_shim.cpp: ref_vel_kick, ref_kepler_isolated) by
 * tests/test_iso_step_oracle.py and against the committed fixture tests/golden/iso_step.npz.
 *
 * Follows, line by line and in the reference's evaluation order (compiled with -ffp-contract=off):
 *   FPGrav::velKick                      src/particle.h:878-884     vel += 0.5*dt_tree*acc
 *   FPGrav::getEccentricity              src/particle.h:668-685
 *   KeplerEq, solveKeplerEq              src/kepler.h:3-38
 *   posVel2OrbitalElement                src/kepler.h:39-79
 *   orbitalElement2PosVel                src/kepler.h:80-97
 *   timeIntegrateKepler_isolated         src/hermite.h:787-816      (INTEGRATE_6TH_SUN off)
 *   calcStarGravity                      src/gravity_hard.h:5-39
 *   calcDt2nd, FPGrav::calcDeltatInitial src/particle.h:345-355,886-914
 *   which particles take this branch     src/hard.h:797-803         (!neighbor.number && ecc < 0.8 && eps2_sun == 0)
 * PS::F64vec arithmetic is FDPS/src/vector3.hpp: dot = (x*x)+(y*y)+(z*z); vec/scalar multiplies by 1.0/s.
 */
#include <float.h>
#include <math.h>

typedef struct { double m_sun, dt_tree, eta_0, eta_sun0, alpha2, dt_min, eps2_sun; } iso_params;
/* what timeIntegrateKepler_isolated leaves in FPGrav besides pos/vel: 64 B */
typedef struct { double phi_s, acc_s[3], jerk_s[3], dt; } iso_star;

static double dot3(const double *a, const double *b) { return (a[0] * b[0]) + (a[1] * b[1]) + (a[2] * b[2]); }

void oracle_vel_kick(int n, double *vel, const double *acc, double dt_tree)
{
    for (int i = 0; i < 3 * n; i++) vel[i] += 0.5 * dt_tree * acc[i];
}

double oracle_eccentricity(const double *pos, const double *vel, double m_sun)
{
    const double r = sqrt(dot3(pos, pos));
    const double rv = dot3(pos, vel);
    const double ax = 1.0 / (2.0 / r - dot3(vel, vel) / m_sun);
    const double ecccosu = 1. - r / ax;
    const double eccsinu2 = rv * rv / (m_sun * ax);
    return sqrt(ecccosu * ecccosu + eccsinu2);
}

static double kepler_eq(double u, double ecc) { return u - ecc * sin(u); }

static double solve_kepler_eq(double l, double ecc)
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
            const double sinu0 = sin(u0), cosu0 = cos(u0);
            u = u0 - ((u0 - ecc * sinu0 - l) / (1. - ecc * cosu0));
            loop++;
        } while (fabs(u - u0) > 1.e-15 && loop < 10);
    }
    return u;
}

static void posvel2elem(const double *pos, const double *vel, double mu, double *ax, double *ecc, double *n, double *u,
                        double *P, double *Q)
{
    const double r2 = dot3(pos, pos), r = sqrt(r2), rinv = 1. / r;
    const double v2 = dot3(vel, vel), rv = dot3(pos, vel);
    *ax = 1.0 / (2.0 * rinv - v2 / mu);
    const double ecccosu = 1. - r / *ax;
    const double eccsinu = rv / sqrt(mu * *ax);
    *ecc = sqrt(ecccosu * ecccosu + eccsinu * eccsinu);
    *n = sqrt(mu / (*ax * *ax * *ax));
    double cosu, sinu;
    if (*ecc != 0) { *u = atan2(eccsinu, ecccosu); cosu = ecccosu / *ecc; sinu = eccsinu / *ecc; }
    else { *u = 0.; cosu = 1.; sinu = 0.; }
    const double aninv = sqrt(*ax / mu);
    const double ecc_sq = sqrt(1. - *ecc * *ecc);
    const double a = rinv * cosu, b = aninv * sinu, c = rinv * sinu, d = aninv * (cosu - *ecc), inv = 1.0 / ecc_sq;
    for (int k = 0; k < 3; k++) {
        P[k] = pos[k] * a - vel[k] * b;
        Q[k] = (pos[k] * c + vel[k] * d) * inv;
    }
}

static void elem2posvel(double *pos, double *vel, double ax, double ecc, double n, double u, const double *P, const double *Q)
{
    const double cosu = cos(u), sinu = sin(u), ecc_sq = sqrt(1. - ecc * ecc);
    const double a = cosu - ecc, b = ecc_sq * sinu;
    for (int k = 0; k < 3; k++) pos[k] = (P[k] * a + Q[k] * b) * ax;
    const double rinv = sqrt(1. / dot3(pos, pos));
    const double s = ax * ax * n * rinv, c = -sinu, d = ecc_sq * cosu;
    for (int k = 0; k < 3; k++) vel[k] = (P[k] * c + Q[k] * d) * s;
}

static double calc_dt2nd(double eta, double alpha2, double acc0, const double *acc, const double *jerk)
{
    const double Acc2 = dot3(acc, acc) + alpha2 * acc0 * acc0;
    const double Jerk2 = dot3(jerk, jerk);
    return (Jerk2 > 0.) ? eta * sqrt(Acc2 / Jerk2) : DBL_MAX;
}

/* One particle through timeIntegrateKepler_isolated(pi, t0, t1); time/dt are FPGrav::time/dt, acc0 is
 * FPGrav::acc0 (set by correctForceLong).  acc_d = jerk_d = 0 and phi_d = 0 on return, as in the reference. */
void oracle_kepler_isolated_one(double *pos, double *vel, double *time, double *dt, double acc0, double t0, double t1,
                                const iso_params *p, iso_star *st)
{
    double ax, ecc, n, u, l, P[3], Q[3];
    posvel2elem(pos, vel, p->m_sun, &ax, &ecc, &n, &u, P, Q);
    l = kepler_eq(u, ecc);
    l += n * (t1 - t0);
    u = solve_kepler_eq(l, ecc);
    elem2posvel(pos, vel, ax, ecc, n, u, P, Q);
    *time += (t1 - t0);
    /* calcStarGravity */
    double dr[3], dv[3];
    for (int k = 0; k < 3; k++) { dr[k] = -pos[k]; dv[k] = -vel[k]; }
    const double r2inv = 1. / (dot3(dr, dr) + p->eps2_sun);
    const double rinv = sqrt(r2inv), r3inv = rinv * r2inv;
    const double mj_rij3 = p->m_sun * r3inv;
    const double alpha = dot3(dr, dv) * r2inv;
    st->phi_s = -p->m_sun * rinv;
    for (int k = 0; k < 3; k++) {
        st->acc_s[k] = dr[k] * mj_rij3;
        st->jerk_s[k] = (dv[k] - dr[k] * (3. * alpha)) * mj_rij3;
    }
    /* calcDeltatInitial with acc_d = jerk_d = 0 */
    const double zero[3] = {0., 0., 0.};
    double dt_next = 0.5 * p->dt_tree;
    const double d1a = calc_dt2nd(p->eta_0, p->alpha2, acc0, zero, zero);
    const double d1b = calc_dt2nd(p->eta_sun0, p->alpha2, 0., st->acc_s, st->jerk_s);
    const double dt_1 = (d1b < d1a) ? d1b : d1a;          /* std::min(a, b) */
    double rem = fmod(*time, dt_next);
    while (rem != 0.0) { dt_next *= 0.5; rem = fmod(*time, dt_next); }
    if (*dt > 0.) while (2. * *dt < dt_next) dt_next *= 0.5;
    while (dt_1 < dt_next) dt_next *= 0.5;
    if (dt_next < 2. * p->dt_min) dt_next = p->dt_min;
    *dt = dt_next;
    st->dt = dt_next;
}

/* The loop of src/hard.h:793-817 over particle arrays: isolated[i] != 0 <=> neighbor.number == 0.
 * Returns how many particles took the Kepler branch; handled[i] = 1 for those, 0 for the ones the
 * reference integrates otherwise (neighbours -> hard clusters, ecc >= 0.8 -> Hermite). */
int oracle_kepler_isolated(int n, double *pos, double *vel, double *time, double *dt, const double *acc0,
                           const int *isolated, double t0, double t1, const iso_params *p, iso_star *star, int *handled)
{
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        handled[i] = 0;
        if (!isolated[i]) continue;
        if (!(oracle_eccentricity(pos + 3 * i, vel + 3 * i, p->m_sun) < 0.8 && p->eps2_sun == 0.)) continue;
        oracle_kepler_isolated_one(pos + 3 * i, vel + 3 * i, time + i, dt + i, acc0[i], t0, t1, p, star + i);
        handled[i] = 1;
        cnt++;
    }
    return cnt;
}
