/*
 * soft_corr_oracle.c -- CPU restatement of GPLUM's changeover correction of the soft force
 * (correctForceLong / correctForceLongInitial) on top of the tree force's neighbour candidates.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as pikg_oracle.c): only tests/, __graft_entry__.smoke()
 * and the CPU legs of bench.py may use it.
 *
 * Parity status: PINNED against the reference's own correctForceLong{,Initial} compiled from
 * /root/reference (oracle/ref_shim.cpp: ref_correct_long) by tests/test_soft_corr_oracle.py.
 *
 * What is restated (line numbers relative to /root/reference):
 *   cutoff_f / cutoff_W / cutoff_K / cutoff_dKdr / cutoff_dKdt   src/cutfunc.h:4-33,41-72
 *   FPGrav::setGamma (g_1_inv, g_1_inv7, w_y, f1)                src/particle.h:619-633
 *   correctForceBetween2Particles                                src/gravity_soft.h:76-153
 *   correctForceBetween2ParticlesInitial                         src/gravity_soft.h:155-242
 *   correctForceLong / ...Initial (per-particle driver)          src/gravity_soft.h:245-372,375-528
 *   NeighborList::addNeighbor (number, id_cluster, inDomain)     src/neighbor.h:635-664
 *   candidate test (which j reach the correction)                src/gravity_kernel.hpp:88-112
 *
 * The reference reaches the candidates either through NeighborInfo::id_min/id_max (<= 2
 * candidates, all on this rank) or through a tree neighbour search (gravity_soft.h:295-311); both
 * hand correctForceBetween2Particles a superset of the candidate set whose extra members fail
 * `rij < r_search` and `rij < r_out`, i.e. contribute nothing.  Here the
 * candidates are enumerated straight from the walk's EP list with the force kernel's FP32 test,
 * in list order.  All correction arithmetic is FP64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t id_local, myrank; double pos[3]; double r_out, r_search; } epi_t;           /* 48  */
typedef struct { int32_t id_local, myrank; double pos[3]; double r_out, r_search;
                 int64_t id; double mass; double vel[3]; double acc_d[3]; } epj_t;                   /* 112 */
typedef struct { float acc[3]; float phi; int32_t number, rank, id_max, id_min; } force_t;           /* 32  */

/* per i-particle result, in the walk-concatenated order of epi_all / force_all */
typedef struct {
    double acc[3];        /* acci  : to be added to (double)force.acc     gravity_soft.h:366 */
    double phi;           /* phii  : to be added to (double)force.phi     :367 */
    double acc0;          /* number / sum(r_out^2/m_j), 0 without neighbours  :368 */
    int64_t id_cluster;   /* min(id_i, ids of the neighbours)             neighbor.h:647 */
    int32_t number;       /* final neighbour count                        neighbor.h:645 */
    int32_t id_local;     /* pp index of this particle (EPI.id_local)     */
    int32_t ngb_off;      /* first entry of this particle in ngb[]        */
    int32_t in_domain;    /* 0 if a neighbour lives on another rank       neighbor.h:660 */
} corr_t;                                                                                            /* 64 */
typedef struct { double acc_d[3]; double jerk_d[3]; double phi_d; double pad; } corr_init_t;         /* 64 */
typedef struct { int64_t id; int32_t rank, id_local; } ngb_t;                                        /* 16 */
typedef struct { double eps2, dt_tree, gamma, R_search2, R_search3; int32_t re_search, initial; } corr_params_t;

_Static_assert(sizeof(corr_t) == 64 && sizeof(corr_init_t) == 64 && sizeof(ngb_t) == 16 && sizeof(corr_params_t) == 48, "layout");

typedef struct { double g, g_1_inv, g_1_inv7, w_y, f1; } cut_t;

/* std::max / std::min as the reference uses them (NaN handling differs from fmax/fmin) */
#define STDMAX(a, b) (((a) < (b)) ? (b) : (a))
#define STDMIN(a, b) (((b) < (a)) ? (b) : (a))

/* src/particle.h:619-633 */
static cut_t set_gamma(double g)
{
    cut_t c;
    c.g = g;
    c.g_1_inv = 1. / (g - 1.);
    const double g2 = g * g;
    const double g_1_inv3 = c.g_1_inv * c.g_1_inv * c.g_1_inv;
    c.g_1_inv7 = g_1_inv3 * g_1_inv3 * c.g_1_inv;
    c.w_y = 7. / 3. * ((((((g - 9.) * g + 45.) * g - 60. * log(g)) * g - 45.) * g + 9.) * g - 1.) * c.g_1_inv7;
    c.f1 = (-10. / 3. + 14. * (g + 1.) - 21. * ((g + 3.) * g + 1.)
            + 35. / 3. * (((g + 9.) * g + 9.) * g + 1.)
            - 70. * ((g + 3.) * g + 1.) * g
            + 210. * (g + 1.) * g2
            + (((g - 7.) * g + 21.) * g - 35.) * g2 * g2) * c.g_1_inv7;
    return c;
}

/* src/cutfunc.h:4-14 */
static double cutoff_f(double y, const cut_t *c)
{
    const double g = c->g, g2 = g * g;
    return (((((((-10. / 3. * y + 14. * (g + 1.)) * y - 21. * ((g + 3.) * g + 1.)) * y
                + 35. / 3. * (((g + 9.) * g + 9.) * g + 1.)) * y
               - 70. * ((g + 3.) * g + 1.) * g) * y
              + 210. * (g + 1.) * g2) * y - 140. * g2 * g * log(y)) * y
            + (((g - 7.) * g + 21.) * g - 35.) * g2 * g2) * c->g_1_inv7;
}
/* src/cutfunc.h:17-29 */
static double cutoff_W(double rij, double r_out_inv, const cut_t *c)
{
    const double y = rij * r_out_inv;
    if (1.0 <= y) return 1.0;
    if (y <= c->g) return y * c->w_y;
    return cutoff_f(y, c) + y * (1. - c->f1);
}
/* src/cutfunc.h:31-40,59-62 */
static double cutoff_K(double rij, double r_out_inv, const cut_t *c)
{
    const double x = (c->g - rij * r_out_inv) * c->g_1_inv;
    if (x < 0.) return 0.;
    if (x >= 1.) return 1.;
    const double x2 = x * x;
    return (((-20. * x + 70.) * x - 84.) * x + 35.) * x2 * x2;
}
/* src/cutfunc.h:41-46,63-66 */
static double cutoff_dKdt(double rij, double r_out_inv, double alpha, const cut_t *c)
{
    const double x = (c->g - rij * r_out_inv) * c->g_1_inv;
    const double x_1 = x - 1.;
    const double dKdr = (x < 0. || x >= 1.) ? 0. : (140. * x * x * x * x_1 * x_1 * x_1 * r_out_inv * c->g_1_inv);
    return alpha * rij * dKdr;
}

/* the force kernel's candidate predicate, src/gravity_kernel.hpp:88-106 (FP32, fallback order) */
static int is_candidate(const epi_t *ei, const epj_t *ej, const double o[3], float eps2)
{
    const float xi = (float)(ei->pos[0] - o[0]), yi = (float)(ei->pos[1] - o[1]), zi = (float)(ei->pos[2] - o[2]);
    const float xj = (float)(ej->pos[0] - o[0]), yj = (float)(ej->pos[1] - o[1]), zj = (float)(ej->pos[2] - o[2]);
    const float rs = fmaxf((float)ei->r_search, (float)ej->r_search);
    const float rs2 = rs * rs * 1.0201f;
    const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
    const float r2_real = dx * dx + dy * dy + dz * dz + eps2;
    return r2_real < rs2;
}

/*
 * Walk arguments as oracle_calc_walks.  force_all may be NULL; when given, every particle's
 * candidate count is checked against force_all[].number (returns -2 on a mismatch).
 * init_out may be NULL unless prm->initial.  Returns the total number of neighbours written to
 * ngb[] (entries of particle i: ngb[out[i].ngb_off .. +out[i].number)), -1 if ngb_cap is too small,
 * -3 if a particle is missing from its own EP list.
 */
long long oracle_correct_long(int n_walk, const epi_t *epi_all, const int *epi_off, const int *ni,
                              const int *adr_epj, const long long *epj_disp, const int *n_epj,
                              const epj_t *epj_all, const force_t *force_all, const corr_params_t *prm,
                              corr_t *out, corr_init_t *init_out, ngb_t *ngb, long long ngb_cap)
{
    const cut_t cut = set_gamma(prm->gamma);
    const double eps2 = prm->eps2;
    long long n_ngb = 0;
    for (int w = 0; w < n_walk; w++) {
        const epi_t *e = epi_all + epi_off[w];
        const int *ae = adr_epj + epj_disp[w];
        const double o[3] = {e[0].pos[0], e[0].pos[1], e[0].pos[2]};
        for (int i = 0; i < ni[w]; i++) {
            const int gi = epi_off[w] + i;
            const epj_t *self = NULL;
            for (int j = 0; j < n_epj[w]; j++) {
                const epj_t *q = epj_all + ae[j];
                if (q->id_local == e[i].id_local && q->myrank == e[i].myrank) { self = q; break; }
            }
            if (!self) return -3;
            double acci[3] = {0., 0., 0.}, phii = 0., acc0i = 0.;
            double acc_di[3] = {0., 0., 0.}, jerki[3] = {0., 0., 0.}, phi_di = 0.;
            int64_t id_cluster = self->id;
            int number = 0, in_domain = 1, n_cand = 0;
            const double r_out_inv_i = 1. / self->r_out;          /* particle.h:728 */
            phii += self->mass * r_out_inv_i;                     /* gravity_soft.h:281,293 */
            const long long off0 = n_ngb;
            /* The reference has two branches (gravity_soft.h:295-317): with more than two candidates, or a candidate
             * on another rank, it searches the tree and SKIPS every entry that carries the particle's own id -- the
             * particle itself and, after a merger, the absorbed twin (collisionA.h:267-277); otherwise it walks
             * neighbor.getId() and a twin reaches correctForceBetween2Particles, which adds massj * r_out_inv to phi
             * for an entry with the particle's own id (gravity_soft.h:105-108). */
            int n_all = 0, other_rank = 0;
            for (int j = 0; j < n_epj[w]; j++) {
                const epj_t *q = epj_all + ae[j];
                if (q == self || !is_candidate(&e[i], q, o, (float)eps2)) continue;
                n_all++;
                if (q->myrank != self->myrank) other_rank = 1;
            }
            const int twin_phi = n_all <= 2 && !other_rank;
            for (int j = 0; j < n_epj[w]; j++) {
                const epj_t *q = epj_all + ae[j];
                if (q == self) continue;
                if (!is_candidate(&e[i], q, o, (float)eps2)) continue;
                n_cand++;
                /* ---- correctForceBetween2Particles{,Initial} ---- */
                const double massj = q->mass;
                const double r_out = STDMAX(self->r_out, q->r_out);
                const double r_out_inv = STDMIN(r_out_inv_i, 1. / q->r_out);
                const double r_search = STDMAX(self->r_search, q->r_search);
                if (q->id == self->id) { if (twin_phi) phii += massj * r_out_inv; continue; }
                const double dr[3] = {q->pos[0] - self->pos[0], q->pos[1] - self->pos[1], q->pos[2] - self->pos[2]};
                double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
                dr2 += eps2;
                const double rij = sqrt(dr2);
                const double dv[3] = {q->vel[0] - self->vel[0], q->vel[1] - self->vel[1], q->vel[2] - self->vel[2]};
                const double drdv = dr[0] * dv[0] + dr[1] * dv[1] + dr[2] * dv[2];
                int pass = 1;
                if (prm->re_search) {
                    const double da[3] = {q->acc_d[0] - self->acc_d[0], q->acc_d[1] - self->acc_d[1], q->acc_d[2] - self->acc_d[2]};
                    const double dv2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
                    const double da2 = da[0] * da[0] + da[1] * da[1] + da[2] * da[2];
                    const double t_neg = -drdv / sqrt(dv2);
                    const double t_pos = STDMAX(t_neg, 0.);
                    const double t_min = STDMIN(t_pos, prm->dt_tree);
                    const double dr2_alt = dr2 + 2. * drdv * t_min + dv2 * t_min * t_min;
                    const double dr2_min = STDMIN(dr2, dr2_alt);
                    const double r_crit = prm->R_search2 * r_out;
                    const double v_crit_a = prm->R_search3 * 0.5 * prm->dt_tree;
                    pass = (dr2_min < r_crit * r_crit) || (dv2 < v_crit_a * v_crit_a * da2) || (prm->initial && da2 == 0.);
                }
                if (pass && rij < r_search) {
                    if (n_ngb >= ngb_cap) return -1;
                    ngb[n_ngb].id = q->id; ngb[n_ngb].rank = q->myrank; ngb[n_ngb].id_local = q->id_local;
                    n_ngb++; number++;
                    if (q->id < id_cluster) id_cluster = q->id;
                    if (q->myrank != self->myrank) in_domain = 0;
                    acc0i += r_out * r_out / massj;
                }
                if (rij < r_out) {
                    const double rinv = 1. / rij, r2inv = rinv * rinv, r3inv = rinv * r2inv;
                    const double W = cutoff_W(rij, r_out_inv, &cut);
                    const double K = cutoff_K(rij, r_out_inv, &cut);
                    const double r_min = STDMIN(rinv, r_out_inv);
                    phii -= massj * (rinv * W - r_min);
                    const double ca = massj * (r3inv * K - r_min * r_min * r_min);
                    acci[0] += ca * dr[0]; acci[1] += ca * dr[1]; acci[2] += ca * dr[2];
                    if (prm->initial) {
                        const double alpha = drdv * r2inv;
                        const double dKdt = cutoff_dKdt(rij, r_out_inv, alpha, &cut);
                        const double alpha_c = alpha * (1. - K);
                        phi_di -= massj * rinv * (1. - W);
                        const double cd = massj * r3inv * (1. - K);
                        const double cj = massj * r3inv;
                        for (int k = 0; k < 3; k++) {
                            acc_di[k] += cd * dr[k];
                            jerki[k] += cj * ((1. - K) * dv[k] - (3. * alpha_c + dKdt) * dr[k]);
                        }
                    }
                }
            }
            if (force_all && force_all[gi].number != n_cand) return -2;
            corr_t *c = out + gi;
            c->acc[0] = acci[0]; c->acc[1] = acci[1]; c->acc[2] = acci[2];
            c->phi = phii;
            c->acc0 = (acc0i > 0.) ? number / acc0i : 0.;
            c->id_cluster = id_cluster; c->number = number; c->id_local = e[i].id_local;
            c->ngb_off = (int32_t)off0; c->in_domain = in_domain;
            if (init_out) {
                corr_init_t *ci = init_out + gi;
                for (int k = 0; k < 3; k++) { ci->acc_d[k] = acc_di[k]; ci->jerk_d[k] = jerki[k]; }
                ci->phi_d = phi_di; ci->pad = 0.;
            }
        }
    }
    return n_ngb;
}
