// ref_shim.cpp -- C-ABI window onto the UNMODIFIED reference, compiled from /root/reference.
//
// TEST INFRASTRUCTURE ONLY (see oracle/pikg_oracle.c header).  This translation unit contains
// no reference source: it #includes the reference's single TU (src/main_p3t.cpp, with `main`
// renamed) from where it lies, so every type and functor below is the reference's own:
//   calcForceEPEPWithSearch / calcForceEPSP     src/gravity_kernel.hpp:8-122,125-239 (non-PIKG branch)
//   EPIGrav / EPJGrav / ForceGrav / FPGrav      src/particle.h:70-156,404-
//   Tree_t (PS::TreeForForce<LONG_SYMMETRY,..>) src/main_p3t.cpp:65-74
// Built by oracle/Makefile into oracle/_ref/ (git-ignored, travels to the GPU box).
//
// Exports
//   ref_layout        sizeof/offsetof table used to pin include/gplum_b200.h and the oracle structs
//   ref_epep/ref_epsp one functor call, exactly as FDPS calcForceOnly would issue it
//   ref_calc_walks    FDPS calcForce loop (gather by index -> clear -> EP-EP -> EP-SP), OpenMP
//   ref_tree_*        build the reference's FDPS tree on given particles and record the
//                     interaction lists it hands to the multi-walk-index accelerator interface
//                     (FDPS/src/tree_for_force_impl_force.hpp:63-266)
//   ref_vel_kick / ref_kepler_isolated   FPGrav::velKick and the isolated-particle Kepler drift of the hard
//                     part (src/hard.h:793-817, src/hermite.h:787-816) on particle arrays
//   ref_main          the reference program itself (argc/argv), for energy-history runs
#define main gplum_reference_main
#include "main_p3t.cpp"
#undef main

#include <cstddef>
#include <cstring>
#include <vector>

namespace {
struct Recorded {
    std::vector<EPI_t> epi;        // concatenated per-walk i-particles (sorted order)
    std::vector<int> epi_off, ni;  // per walk
    std::vector<int> adr_epj, adr_spj;
    std::vector<long long> epj_disp, spj_disp;
    std::vector<int> n_epj, n_spj;
    std::vector<EPJ_t> epj_all;
    std::vector<SPJ_t> spj_all;
    std::vector<Force_t> force;    // reference force in the same (walk-concatenated) order
    std::vector<Force_t *> force_ptr;
    long long n_int_epep = 0, n_int_epsp = 0;
    void clear() { *this = Recorded(); }
} g_rec;

bool g_ps_initialized = false;

PS::S32 RecDispatch(const PS::S32 tag, const PS::S32 n_walk, const EPI_t **epi, const PS::S32 *n_epi,
                    const PS::S32 **id_epj, const PS::S32 *n_epj, const PS::S32 **id_spj,
                    const PS::S32 *n_spj, const EPJ_t *epj, const PS::S32 n_epj_tot,
                    const SPJ_t *spj, const PS::S32 n_spj_tot, const bool send_flag)
{
    (void)tag;
    if (send_flag) {
        g_rec.epj_all.assign(epj, epj + n_epj_tot);
        g_rec.spj_all.assign(spj, spj + n_spj_tot);
        return 0;
    }
    for (int w = 0; w < n_walk; w++) {
        g_rec.epi_off.push_back((int)g_rec.epi.size());
        g_rec.ni.push_back(n_epi[w]);
        g_rec.epi.insert(g_rec.epi.end(), epi[w], epi[w] + n_epi[w]);
        g_rec.epj_disp.push_back((long long)g_rec.adr_epj.size());
        g_rec.spj_disp.push_back((long long)g_rec.adr_spj.size());
        g_rec.n_epj.push_back(n_epj[w]);
        g_rec.n_spj.push_back(n_spj[w]);
        g_rec.adr_epj.insert(g_rec.adr_epj.end(), id_epj[w], id_epj[w] + n_epj[w]);
        g_rec.adr_spj.insert(g_rec.adr_spj.end(), id_spj[w], id_spj[w] + n_spj[w]);
        g_rec.n_int_epep += (long long)n_epi[w] * n_epj[w];
        g_rec.n_int_epsp += (long long)n_epi[w] * n_spj[w];
    }
    return 0;
}

// Retrieve: evaluate the walks just recorded with the reference functors themselves.
PS::S32 RecRetrieve(const PS::S32 tag, const PS::S32 n_walk, const PS::S32 *ni, Force_t **force)
{
    (void)tag;
    const int w0 = (int)g_rec.force_ptr.size();
    std::vector<EPJ_t> ebuf;
    std::vector<SPJ_t> sbuf;
    for (int w = 0; w < n_walk; w++) {
        const int gw = w0 + w;
        ebuf.resize(g_rec.n_epj[gw]);
        sbuf.resize(g_rec.n_spj[gw]);
        for (int j = 0; j < g_rec.n_epj[gw]; j++) ebuf[j] = g_rec.epj_all[g_rec.adr_epj[g_rec.epj_disp[gw] + j]];
        for (int j = 0; j < g_rec.n_spj[gw]; j++) sbuf[j] = g_rec.spj_all[g_rec.adr_spj[g_rec.spj_disp[gw] + j]];
        for (int i = 0; i < ni[w]; i++) force[w][i].clear();
        calcForceEPEPWithSearch()(&g_rec.epi[g_rec.epi_off[gw]], ni[w], ebuf.data(), (int)ebuf.size(), force[w]);
        calcForceEPSP()(&g_rec.epi[g_rec.epi_off[gw]], ni[w], sbuf.data(), (int)sbuf.size(), force[w]);
        g_rec.force_ptr.push_back(force[w]);
        g_rec.force.insert(g_rec.force.end(), force[w], force[w] + ni[w]);
    }
    return 0;
}
}  // namespace

extern "C" {

int ref_abi_version() { return 1; }

// out[] = sizeof(EPI), sizeof(EPJ), sizeof(SPJ), sizeof(Force), sizeof(FPGrav), then offsets:
// EPI{id_local,myrank,pos,r_out,r_search} EPJ{id,mass,vel,acc_d} SPJ{mass,pos,quad}
// Force{acc,phi,neighbor} NeighborInfo{number,rank,id_max,id_min}; returns the count written.
int ref_layout(int *out)
{
    int k = 0;
    out[k++] = (int)sizeof(EPI_t); out[k++] = (int)sizeof(EPJ_t); out[k++] = (int)sizeof(SPJ_t);
    out[k++] = (int)sizeof(Force_t); out[k++] = (int)sizeof(FP_t);
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
    out[k++] = (int)offsetof(EPI_t, id_local); out[k++] = (int)offsetof(EPI_t, myrank);
    out[k++] = (int)offsetof(EPI_t, pos); out[k++] = (int)offsetof(EPI_t, r_out);
    out[k++] = (int)offsetof(EPI_t, r_search);
    out[k++] = (int)offsetof(EPJ_t, id); out[k++] = (int)offsetof(EPJ_t, mass);
    out[k++] = (int)offsetof(EPJ_t, vel); out[k++] = (int)offsetof(EPJ_t, acc_d);
    out[k++] = (int)offsetof(SPJ_t, mass); out[k++] = (int)offsetof(SPJ_t, pos);
    out[k++] = (int)offsetof(SPJ_t, quad);
    out[k++] = (int)offsetof(Force_t, acc); out[k++] = (int)offsetof(Force_t, phi);
    out[k++] = (int)offsetof(Force_t, neighbor);
    out[k++] = (int)offsetof(NeighborInfo, number); out[k++] = (int)offsetof(NeighborInfo, rank);
    out[k++] = (int)offsetof(NeighborInfo, id_max); out[k++] = (int)offsetof(NeighborInfo, id_min);
    out[k++] = (int)sizeof(MySPJMonopole);
#pragma GCC diagnostic pop
    return k;
}

void ref_force_clear(void *force, int n)
{
    Force_t *f = (Force_t *)force;
    for (int i = 0; i < n; i++) f[i].clear();
}

void ref_epep(const void *epi, int ni, const void *epj, int nj, void *force, float eps2)
{
    FP_t::eps2 = eps2;
    calcForceEPEPWithSearch()((const EPI_t *)epi, ni, (const EPJ_t *)epj, nj, (Force_t *)force);
}

void ref_epsp(const void *epi, int ni, const void *spj, int nj, void *force, float eps2)
{
    FP_t::eps2 = eps2;
    calcForceEPSP()((const EPI_t *)epi, ni, (const SPJ_t *)spj, nj, (Force_t *)force);
}

// Same contract as oracle_calc_walks (oracle/pikg_oracle.c), executed by the reference's functors
// in the reference's loop structure (FDPS/src/tree_for_force_impl_force.hpp:1515-1535).
long long ref_calc_walks(int n_walk, const void *epi_all_, const int *epi_off, const int *ni,
                         const int *adr_epj, const long long *epj_disp, const int *n_epj,
                         const int *adr_spj, const long long *spj_disp, const int *n_spj,
                         const void *epj_all_, const void *spj_all_, void *force_all_,
                         float eps2, int clear, int n_threads)
{
    const EPI_t *epi_all = (const EPI_t *)epi_all_;
    const EPJ_t *epj_all = (const EPJ_t *)epj_all_;
    const SPJ_t *spj_all = (const SPJ_t *)spj_all_;
    Force_t *force_all = (Force_t *)force_all_;
    FP_t::eps2 = eps2;
    long long n_int = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : n_int)
    {
        std::vector<EPJ_t> ebuf;
        std::vector<SPJ_t> sbuf;
#pragma omp for schedule(guided)
        for (int w = 0; w < n_walk; w++) {
            ebuf.resize(n_epj[w]);
            sbuf.resize(n_spj[w]);
            const int *ae = adr_epj + epj_disp[w];
            const int *as = adr_spj + spj_disp[w];
            for (int j = 0; j < n_epj[w]; j++) ebuf[j] = epj_all[ae[j]];
            for (int j = 0; j < n_spj[w]; j++) sbuf[j] = spj_all[as[j]];
            Force_t *f = force_all + epi_off[w];
            if (clear) for (int i = 0; i < ni[w]; i++) f[i].clear();
            calcForceEPEPWithSearch()(epi_all + epi_off[w], ni[w], ebuf.data(), n_epj[w], f);
            calcForceEPSP()(epi_all + epi_off[w], ni[w], sbuf.data(), n_spj[w], f);
            n_int += (long long)ni[w] * (n_epj[w] + n_spj[w]);
        }
    }
    return n_int;
}

// Build the reference's tree on n particles and record what FDPS's multi-walk-index interface
// dispatches.  id_local = array index, myrank = 0 (src/func.h:135-147 on one rank).
// Returns the number of walks (i-groups).
int ref_tree_build(int n, const double *pos, const double *vel, const double *mass, const double *r_out,
                   const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                   int n_walk_limit, float eps2)
{
    if (!g_ps_initialized) {
        int argc = 1;
        char arg0[] = "ref_shim";
        char *argv_[] = {arg0, nullptr};
        char **argv = argv_;
        PS::Initialize(argc, argv);
        g_ps_initialized = true;
    }
    g_rec.clear();
    FP_t::eps2 = eps2;
    PS::ParticleSystem<FP_t> psys;
    psys.initialize();
    psys.setNumberOfParticleLocal(n);
    for (int i = 0; i < n; i++) {
        psys[i].id = i;
        psys[i].id_local = i;
        psys[i].myrank = 0;
        psys[i].pos = PS::F64vec(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        psys[i].vel = vel ? PS::F64vec(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : PS::F64vec(0.0);
        psys[i].acc_d = 0.0;
        psys[i].mass = mass[i];
        psys[i].r_out = r_out[i];
        psys[i].r_search = r_search[i];
    }
    PS::DomainInfo dinfo;
    dinfo.initialize(0.3);
    dinfo.setNumberOfDomainMultiDimension(1, 1, 1);
    dinfo.setBoundaryCondition(PS::BOUNDARY_CONDITION_OPEN);
    dinfo.collectSampleParticle(psys, true);
    dinfo.decomposeDomain();
    Tree_t tree;
    tree.initialize(n, theta, n_leaf_limit, n_group_limit);
    tree.calcForceAllAndWriteBackMultiWalkIndex(RecDispatch, RecRetrieve, 1, psys, dinfo, n_walk_limit, true);
    return (int)g_rec.ni.size();
}

// sizes[] = n_walk, n_epi_total, n_adr_epj, n_adr_spj, n_epj_all, n_spj_all
void ref_tree_sizes(long long *sizes)
{
    sizes[0] = (long long)g_rec.ni.size();
    sizes[1] = (long long)g_rec.epi.size();
    sizes[2] = (long long)g_rec.adr_epj.size();
    sizes[3] = (long long)g_rec.adr_spj.size();
    sizes[4] = (long long)g_rec.epj_all.size();
    sizes[5] = (long long)g_rec.spj_all.size();
    sizes[6] = g_rec.n_int_epep;
    sizes[7] = g_rec.n_int_epsp;
}

void ref_tree_copy(void *epi, int *epi_off, int *ni, int *adr_epj, long long *epj_disp, int *n_epj,
                   int *adr_spj, long long *spj_disp, int *n_spj, void *epj_all, void *spj_all, void *force)
{
    const size_t nw = g_rec.ni.size();
    std::memcpy(epi, g_rec.epi.data(), g_rec.epi.size() * sizeof(EPI_t));
    std::memcpy(epi_off, g_rec.epi_off.data(), nw * sizeof(int));
    std::memcpy(ni, g_rec.ni.data(), nw * sizeof(int));
    std::memcpy(adr_epj, g_rec.adr_epj.data(), g_rec.adr_epj.size() * sizeof(int));
    std::memcpy(epj_disp, g_rec.epj_disp.data(), nw * sizeof(long long));
    std::memcpy(n_epj, g_rec.n_epj.data(), nw * sizeof(int));
    std::memcpy(adr_spj, g_rec.adr_spj.data(), g_rec.adr_spj.size() * sizeof(int));
    std::memcpy(spj_disp, g_rec.spj_disp.data(), nw * sizeof(long long));
    std::memcpy(n_spj, g_rec.n_spj.data(), nw * sizeof(int));
    std::memcpy(epj_all, g_rec.epj_all.data(), g_rec.epj_all.size() * sizeof(EPJ_t));
    std::memcpy(spj_all, g_rec.spj_all.data(), g_rec.spj_all.size() * sizeof(SPJ_t));
    if (force) std::memcpy(force, g_rec.force.data(), g_rec.force.size() * sizeof(Force_t));
}

// ---- the changeover correction: the reference's own correctForceLong / correctForceLongInitial ----
// Runs what src/main_p3t.cpp:582-593 (or :351-361 with initial != 0) runs on one rank: id_local /
// myrank assignment (src/func.h:135-147), the tree force through the multi-walk-index interface
// (so the interaction lists of exactly this tree are recorded for ref_tree_copy), then
// correctForceLong (src/gravity_soft.h:245-372) or correctForceLongInitial (:375-528) with the
// reference's NeighborList.  Per particle (original order) out_f64[i*16 ..] =
// {acc xyz, phi, acc0, acc_d xyz, phi_d, jerk_d xyz, acc_before xyz (F32 tree force widened)},
// out_i64[i*4 ..] = {id_cluster, neighbor.number, inDomain, offset into ngb}; ngb holds
// {id, rank, id_local} triples (3 x int64 per neighbour).  Returns the total neighbour count, or
// -1 when ngb_cap is too small.
long long ref_correct_long(int n, const double *pos, const double *vel, const double *acc_d, const double *mass,
                           const double *r_out, const double *r_search, const long long *id,
                           double theta, int n_leaf_limit, int n_group_limit, int n_walk_limit,
                           double eps2, double dt_tree, double gamma, double R_search2, double R_search3,
                           int initial, double *out_f64, long long *out_i64, long long *ngb, long long ngb_cap)
{
    if (!g_ps_initialized) {
        int argc = 1;
        char arg0[] = "ref_shim";
        char *argv_[] = {arg0, nullptr};
        char **argv = argv_;
        PS::Initialize(argc, argv);
        g_ps_initialized = true;
    }
    g_rec.clear();
    FP_t::eps2 = eps2;
    FP_t::dt_tree = dt_tree;
    FP_t::R_search2 = R_search2;
    FP_t::R_search3 = R_search3;
    FP_t::setGamma(gamma);
    PS::ParticleSystem<FP_t> psys;
    psys.initialize();
    psys.setNumberOfParticleLocal(n);
    for (int i = 0; i < n; i++) {
        psys[i].id = id ? id[i] : i;
        psys[i].pos = PS::F64vec(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        psys[i].vel = PS::F64vec(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        psys[i].acc_d = PS::F64vec(acc_d[3 * i], acc_d[3 * i + 1], acc_d[3 * i + 2]);
        psys[i].mass = mass[i];
        psys[i].r_out = r_out[i];
        psys[i].r_out_inv = 1. / r_out[i];          // src/particle.h:728
        psys[i].r_search = r_search[i];
    }
    NeighborList NList;
    setIDLocalAndMyrank(psys, NList);
    PS::DomainInfo dinfo;
    dinfo.initialize(0.3);
    dinfo.setNumberOfDomainMultiDimension(1, 1, 1);
    dinfo.setBoundaryCondition(PS::BOUNDARY_CONDITION_OPEN);
    dinfo.collectSampleParticle(psys, true);
    dinfo.decomposeDomain();
    Tree_t tree;
    tree.initialize(n, theta, n_leaf_limit, n_group_limit);
    tree.calcForceAllAndWriteBackMultiWalkIndex(RecDispatch, RecRetrieve, 1, psys, dinfo, n_walk_limit, true);
    std::vector<PS::F64vec> before(n);
    for (int i = 0; i < n; i++) before[i] = psys[i].acc;
    PS::S32 n_ngb_tot = 0, n_with_ngb = 0;
    if (initial) correctForceLongInitial(psys, tree, NList, n_ngb_tot, n_with_ngb);
    else correctForceLong(psys, tree, NList, n_ngb_tot, n_with_ngb);
    long long off = 0;
    for (int i = 0; i < n; i++) {
        double *o = out_f64 + 16 * (size_t)i;
        o[0] = psys[i].acc.x; o[1] = psys[i].acc.y; o[2] = psys[i].acc.z; o[3] = psys[i].phi; o[4] = psys[i].acc0;
        o[5] = psys[i].acc_d.x; o[6] = psys[i].acc_d.y; o[7] = psys[i].acc_d.z; o[8] = psys[i].phi_d;
        o[9] = psys[i].jerk_d.x; o[10] = psys[i].jerk_d.y; o[11] = psys[i].jerk_d.z;
        o[12] = before[i].x; o[13] = before[i].y; o[14] = before[i].z; o[15] = 0.0;
        long long *q = out_i64 + 4 * (size_t)i;
        q[0] = psys[i].id_cluster; q[1] = psys[i].neighbor.number; q[2] = psys[i].inDomain ? 1 : 0; q[3] = off;
        const std::vector<NeighborId> &l = NList.n_list[i];
        if ((long long)l.size() != psys[i].neighbor.number) return -2;
        for (size_t k = 0; k < l.size(); k++) {
            if (off >= ngb_cap) return -1;
            ngb[3 * off] = l[k].id; ngb[3 * off + 1] = l[k].rank; ngb[3 * off + 2] = l[k].id_local;
            off++;
        }
    }
    if (off != n_ngb_tot) return -3;
    return off;
}

// ---- the reference's whole soft-force stage, timed (bench.py: cpu_baseline_stage) ----
// Exactly the calls of src/main_p3t.cpp:583-593: tree_grav.calcForceAllAndWriteBack(calcForceEPEPWithSearch(),
// calcForceEPSP(), system_grav, dinfo, true, MAKE_LIST, false) and correctForceLong(...), on one rank with all
// OpenMP threads.  seconds_out[0] = force stage (tree build + walks + functors + write-back), [1] = correctForceLong,
// both the minimum over `reps` repetitions; returns the number of neighbours found (>= 0).
long long ref_stage_time(int n, const double *pos, const double *vel, const double *mass, const double *r_out,
                         const double *r_search, double theta, int n_leaf_limit, int n_group_limit, double eps2,
                         double dt_tree, double gamma, double R_search2, double R_search3, int reps, double *seconds_out)
{
    if (!g_ps_initialized) {
        int argc = 1;
        char arg0[] = "ref_shim";
        char *argv_[] = {arg0, nullptr};
        char **argv = argv_;
        PS::Initialize(argc, argv);
        g_ps_initialized = true;
    }
    FP_t::eps2 = eps2;
    FP_t::dt_tree = dt_tree;
    FP_t::R_search2 = R_search2;
    FP_t::R_search3 = R_search3;
    FP_t::setGamma(gamma);
    PS::ParticleSystem<FP_t> psys;
    psys.initialize();
    psys.setNumberOfParticleLocal(n);
    for (int i = 0; i < n; i++) {
        psys[i].id = i;
        psys[i].pos = PS::F64vec(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        psys[i].vel = PS::F64vec(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        psys[i].acc_d = 0.0;
        psys[i].mass = mass[i];
        psys[i].r_out = r_out[i];
        psys[i].r_out_inv = 1. / r_out[i];
        psys[i].r_search = r_search[i];
    }
    NeighborList NList;
    setIDLocalAndMyrank(psys, NList);
    PS::DomainInfo dinfo;
    dinfo.initialize(0.3);
    dinfo.setNumberOfDomainMultiDimension(1, 1, 1);
    dinfo.setBoundaryCondition(PS::BOUNDARY_CONDITION_OPEN);
    dinfo.collectSampleParticle(psys, true);
    dinfo.decomposeDomain();
    Tree_t tree;
    tree.initialize(n, theta, n_leaf_limit, n_group_limit);
    PS::S32 n_ngb_tot = 0, n_with_ngb = 0;
    double best_f = 1e300, best_c = 1e300;
    for (int r = 0; r < reps; r++) {
        const double t0 = PS::GetWtime();
        tree.calcForceAllAndWriteBack(calcForceEPEPWithSearch(), calcForceEPSP(), psys, dinfo, true, PS::MAKE_LIST, false);
        const double t1 = PS::GetWtime();
        correctForceLong(psys, tree, NList, n_ngb_tot, n_with_ngb);
        const double t2 = PS::GetWtime();
        best_f = std::min(best_f, t1 - t0);
        best_c = std::min(best_c, t2 - t1);
    }
    seconds_out[0] = best_f; seconds_out[1] = best_c;
    return n_ngb_tot;
}

// ---- snapshot wire format: layout of the raw records the reference dumps ----
// snap_tmp.dat = FileHeader (src/energy.h:70-127, fwrite of the struct) + n_body x FPGrav (src/particle.h:860-876).
// out[] = sizeof(FileHeader), sizeof(Energy), offsetof(FileHeader: n_body, id_next, time, e_init, e_now),
//         sizeof(FPGrav), then offsetof(FPGrav: id_local, myrank, pos, r_out, r_search, id, mass, vel, acc_d,
//         acc, acc_s, jerk_d, jerk_s, acc_gd, phi, phi_d, phi_s, v_disp, r_out_inv, time, dt, acc0, r_planet, f,
//         neighbor, id_cluster, n_cluster, inDomain, isSent, isDead, isMerged); returns the count written.
int ref_snapshot_layout(int *out)
{
    int k = 0;
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
    out[k++] = (int)sizeof(FileHeader); out[k++] = (int)sizeof(Energy);
    out[k++] = (int)offsetof(FileHeader, n_body); out[k++] = (int)offsetof(FileHeader, id_next);
    out[k++] = (int)offsetof(FileHeader, time); out[k++] = (int)offsetof(FileHeader, e_init);
    out[k++] = (int)offsetof(FileHeader, e_now);
    out[k++] = (int)sizeof(FP_t);
    out[k++] = (int)offsetof(FP_t, id_local); out[k++] = (int)offsetof(FP_t, myrank); out[k++] = (int)offsetof(FP_t, pos);
    out[k++] = (int)offsetof(FP_t, r_out); out[k++] = (int)offsetof(FP_t, r_search); out[k++] = (int)offsetof(FP_t, id);
    out[k++] = (int)offsetof(FP_t, mass); out[k++] = (int)offsetof(FP_t, vel); out[k++] = (int)offsetof(FP_t, acc_d);
    out[k++] = (int)offsetof(FP_t, acc); out[k++] = (int)offsetof(FP_t, acc_s); out[k++] = (int)offsetof(FP_t, jerk_d);
    out[k++] = (int)offsetof(FP_t, jerk_s); out[k++] = (int)offsetof(FP_t, acc_gd); out[k++] = (int)offsetof(FP_t, phi);
    out[k++] = (int)offsetof(FP_t, phi_d); out[k++] = (int)offsetof(FP_t, phi_s); out[k++] = (int)offsetof(FP_t, v_disp);
    out[k++] = (int)offsetof(FP_t, r_out_inv); out[k++] = (int)offsetof(FP_t, time); out[k++] = (int)offsetof(FP_t, dt);
    out[k++] = (int)offsetof(FP_t, acc0); out[k++] = (int)offsetof(FP_t, r_planet); out[k++] = (int)offsetof(FP_t, f);
    out[k++] = (int)offsetof(FP_t, neighbor); out[k++] = (int)offsetof(FP_t, id_cluster); out[k++] = (int)offsetof(FP_t, n_cluster);
    out[k++] = (int)offsetof(FP_t, inDomain); out[k++] = (int)offsetof(FP_t, isSent); out[k++] = (int)offsetof(FP_t, isDead);
    out[k++] = (int)offsetof(FP_t, isMerged);
#pragma GCC diagnostic pop
    return k;
}

// ---- the isolated-particle half of a soft step: the reference's own velKick and Kepler drift ----
// ref_vel_kick: FPGrav::velKick (src/particle.h:878-884) on n particles.
void ref_vel_kick(int n, double *vel, const double *acc, double dt_tree)
{
    FP_t::dt_tree = dt_tree;
    for (int i = 0; i < n; i++) {
        FP_t p;
        p.vel = PS::F64vec(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        p.acc = PS::F64vec(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]);
        p.velKick();
        vel[3 * i] = p.vel.x; vel[3 * i + 1] = p.vel.y; vel[3 * i + 2] = p.vel.z;
    }
}

// ref_kepler_isolated: the loop of src/hard.h:793-817 -- particles with neighbor.number == 0 and
// getEccentricity() < 0.8 (and eps2_sun == 0) go through timeIntegrateKepler_isolated
// (src/hermite.h:787-816).  prm = {m_sun, dt_tree, eta_0, eta_sun0, alpha2, dt_min, eps2_sun};
// star[i*8 ..] = {phi_s, acc_s xyz, jerk_s xyz, dt}; handled[i] = 1 where the branch was taken.
int ref_kepler_isolated(int n, double *pos, double *vel, double *time, double *dt, const double *acc0,
                        const int *isolated, double t0, double t1, const double *prm, double *star, int *handled)
{
    FP_t::m_sun = prm[0]; FP_t::dt_tree = prm[1]; FP_t::eta_0 = prm[2]; FP_t::eta_sun0 = prm[3];
    FP_t::alpha2 = prm[4]; FP_t::dt_min = prm[5]; FP_t::eps2_sun = prm[6];
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        handled[i] = 0;
        FP_t p;
        p.pos = PS::F64vec(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        p.vel = PS::F64vec(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        p.time = time[i]; p.dt = dt[i]; p.acc0 = acc0[i];
        p.neighbor.number = isolated[i] ? 0 : 1;
        if (p.neighbor.number) continue;
        if (!(p.getEccentricity() < 0.8 && FP_t::eps2_sun == 0.)) continue;
        timeIntegrateKepler_isolated(p, t0, t1);
        pos[3 * i] = p.pos.x; pos[3 * i + 1] = p.pos.y; pos[3 * i + 2] = p.pos.z;
        vel[3 * i] = p.vel.x; vel[3 * i + 1] = p.vel.y; vel[3 * i + 2] = p.vel.z;
        time[i] = p.time; dt[i] = p.dt;
        double *o = star + 8 * (size_t)i;
        o[0] = p.phi_s; o[1] = p.acc_s.x; o[2] = p.acc_s.y; o[3] = p.acc_s.z;
        o[4] = p.jerk_s.x; o[5] = p.jerk_s.y; o[6] = p.jerk_s.z; o[7] = p.dt;
        if (p.phi_d != 0. || p.acc_d.x != 0. || p.jerk_d.x != 0.) return -1;
        handled[i] = 1;
        cnt++;
    }
    return cnt;
}

// The reference program, unmodified (src/main_p3t.cpp:83).  Runs in the current directory.
int ref_main(int argc, char **argv) { return gplum_reference_main(argc, argv); }

}  // extern "C"
