// gravity_tree_b200.hpp -- the third form of the drop-in: the WHOLE soft-force stage on the GPU.
//
// Replaces FDPS's TreeForForce at GPLUM's three call sites
//     tree_grav.calcForceAllAndWriteBack(calcForceEPEPWithSearch(), calcForceEPSP(), system_grav, dinfo,
//                                        true, MY_INTERACTION_LIST_MODE, false);          src/main_p3t.cpp:351,583,672
// and the two post-passes that follow them
//     correctForceLong / correctForceLongInitial(system_grav, tree_grav, NList, n_ngb_tot, n_with_ngb);
//                                                                                          src/main_p3t.cpp:358-360,593,684
// by ONE type alias:   using Tree_t = gplum_b200::TreeB200;        (instead of PS::TreeForForce<...>, src/main_p3t.cpp:71-73)
// plus `#include "gravity_tree_b200.hpp"` after `#include "gravity_kernel.hpp"` (src/main_p3t.cpp:63).  The call
// sites themselves stay as they are: TreeB200 has the members they use, and the overloads of correctForceLong{,Initial}
// below are more specialised than the reference's templates (src/gravity_soft.h:245,375), so they are the ones chosen.
//
// calcForceAllAndWriteBack: 48 B per particle go up (position, mass, r_out, r_search -- what the interaction kernels
// read of EPJGrav) and 16 B per particle come back ({acc, phi}, in particle order) plus the four neighbour words of
// the particles that have candidates; Morton sort, tree, moments, i-groups, interaction lists, both interaction
// kernels and the neighbour candidates stay on the device (gplum_b200/csrc/dev_tree.cu, kernels.cuh).  The post-pass
// sends the fields only IT reads (velocity and direct acceleration, 48 B) of the particles that occur in candidate
// pairs (9 % of an N = 1e6 disk), and downloads what soft_corr.cu
// computed from the pass's candidate pairs -- the FP64 changeover correction and the final neighbour lists -- and
// enters it into FPGrav / NeighborList exactly where the reference does (src/gravity_soft.h:349-368,506-521,
// src/neighbor.h:636-668); star gravity and the first time step of the Initial form stay the reference's own host
// functions.  Single rank (open boundary), default macro set.
#pragma once
#include <cstdio>
#include <vector>

#include "gplum_b200.h"

namespace gplum_b200 {

inline void tree_check(int rc, const char *what)
{
    if (rc != 0) {
        std::fprintf(stderr, "libgplum_b200: %s failed (%d): %s\n", what, rc, gplum_b200_last_error());
        PS::Abort(-1);
    }
}

// a growable array in page-locked host memory (gplum_b200_pinned_alloc): what crosses PCIe directly
template <class T>
class PinnedArray {
    T *p_ = nullptr;
    size_t cap_ = 0;
public:
    PinnedArray() = default;
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    ~PinnedArray() { gplum_b200_pinned_free(p_); }
    void resize(size_t n)
    {
        if (n <= cap_) return;
        gplum_b200_pinned_free(p_);
        cap_ = n + n / 4 + 64;
        p_ = static_cast<T *>(gplum_b200_pinned_alloc(cap_ * sizeof(T)));
        if (!p_) tree_check(-1, "gplum_b200_pinned_alloc");
    }
    T *data() { return p_; }
    T &operator[](size_t i) { return p_[i]; }
};

class TreeB200 {
public:
    PS::F64 theta_ = 0.5;
    PS::S32 n_leaf_limit_ = 8, n_group_limit_ = 64;
    PinnedArray<double> pos_, mass_, r_out_, r_search_, vel_, acc_d_;   // columns, particle k at slot k (FDPS's epj_org_ order)
    PinnedArray<long long> id_;
    PinnedArray<float> accphi_;
    PinnedArray<int> nb_index_, nb_;
    PinnedArray<gplum_b200_corr> corr_;
    PinnedArray<gplum_b200_corr_init> init_;
    PinnedArray<gplum_b200_ngb> ngb_;
    long long sizes_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PS::S64 n_walk_ = 0, n_int_epep_ = 0, n_int_epsp_ = 0;
    PS::S32 n_listed_ = 0;                       // particles the last pass's download listed (candidate pairs)

    void initialize(const PS::U64 n_glb_tot, const PS::F64 theta = 0.7, const PS::U32 n_leaf_limit = 8, const PS::U32 n_group_limit = 64)
    {
        static_assert(sizeof(EPJ_t) == 112 && sizeof(Force_t) == 32, "EPJGrav / ForceGrav layout (default macro set)");
        theta_ = theta; n_leaf_limit_ = (PS::S32)n_leaf_limit; n_group_limit_ = (PS::S32)n_group_limit;
        tree_check(gplum_b200_init(0, (size_t)n_glb_tot, 0), "gplum_b200_init");
    }
    template <class T> void setExchangeLETMode(const T) {}
    void clearNumberOfInteraction() { n_walk_ = n_int_epep_ = n_int_epsp_ = 0; }
    PS::S64 getNumberOfWalkGlobal() const { return n_walk_; }
    PS::S64 getNumberOfInteractionEPEPGlobal() const { return n_int_epep_; }
    PS::S64 getNumberOfInteractionEPSPGlobal() const { return n_int_epsp_; }
    // Wtime::showTime (src/time.h:66-73) prints FDPS's phase times: the stage's wall time goes to calc_force
    PS::TimeProfile time_profile_;
    PS::TimeProfile getTimeProfile() const { return time_profile_; }
    void clearTimeProfile() { time_profile_.clear(); }

    // FDPS/src/tree_for_force.hpp:1239-1253.  The functor arguments name the kernels; the library runs its own.
    template <class Tfunc_ep_ep, class Tfunc_ep_sp, class Tpsys, class Tdinfo>
    void calcForceAllAndWriteBack(Tfunc_ep_ep, Tfunc_ep_sp, Tpsys &psys, Tdinfo &, const bool clear = true,
                                  const PS::INTERACTION_LIST_MODE = PS::MAKE_LIST, const bool = false)
    {
        const PS::F64 t0 = PS::GetWtime();
        const PS::S32 n = psys.getNumberOfParticleLocal();
        const size_t N = (size_t)n;
        pos_.resize(3 * N); mass_.resize(N); r_out_.resize(N); r_search_.resize(N);
        accphi_.resize(4 * N); nb_index_.resize(N); nb_.resize(4 * N);
#pragma omp parallel for
        for (PS::S32 i = 0; i < n; i++) {            // EPJGrav::copyFromFP (src/particle.h:131-169), the fields the kernels read
            EPJ_t e;
            e.copyFromFP(psys[i]);
            pos_[3 * (size_t)i] = e.pos.x; pos_[3 * (size_t)i + 1] = e.pos.y; pos_[3 * (size_t)i + 2] = e.pos.z;
            mass_[i] = e.mass; r_out_[i] = e.r_out; r_search_[i] = e.r_search;
        }
#ifdef USE_QUAD
        const int quad = 1;
#else
        const int quad = 0;
#endif
        tree_check(gplum_b200_set_params((float)FP_t::eps2, quad, -1), "gplum_b200_set_params");
        tree_check(gplum_b200_soft_corr_enable(1, 0), "gplum_b200_soft_corr_enable");       // the post-pass needs the pairs
        tree_check(gplum_b200_walks_select(0), "gplum_b200_walks_select");
        tree_check(gplum_b200_tree_build_gpu(n, pos_.data(), mass_.data(), r_out_.data(), r_search_.data(), theta_, n_leaf_limit_,
                                             n_group_limit_, PS::Comm::getRank(), sizes_), "gplum_b200_tree_build_gpu");
        tree_check(gplum_b200_walks_run(0), "gplum_b200_walks_run");
        int n_nb = 0;
        tree_check(gplum_b200_tree_download_compact(accphi_.data(), nb_index_.data(), nb_.data(), n, &n_nb), "gplum_b200_tree_download_compact");
        n_listed_ = n_nb;
        n_walk_ += sizes_[0]; n_int_epep_ += sizes_[6]; n_int_epsp_ += sizes_[7];
        (void)clear;                                                     // the pass overwrites: ForceGrav::clear is fused
#pragma omp parallel for
        for (PS::S32 i = 0; i < n; i++) {
            Force_t f;
            f.clear();                                                   // neighbour words of a particle without candidates
            f.acc.x = accphi_[4 * (size_t)i]; f.acc.y = accphi_[4 * (size_t)i + 1]; f.acc.z = accphi_[4 * (size_t)i + 2];
            f.phi = accphi_[4 * (size_t)i + 3];
            psys[i].copyFromForce(f);
        }
#pragma omp parallel for
        for (PS::S32 k = 0; k < n_nb; k++) {
            NeighborInfo &nb = psys[nb_index_[k]].neighbor;
            nb.number = nb_[4 * (size_t)k]; nb.rank = nb_[4 * (size_t)k + 1]; nb.id_max = nb_[4 * (size_t)k + 2]; nb.id_min = nb_[4 * (size_t)k + 3];
        }
        time_profile_.calc_force += PS::GetWtime() - t0;
    }

    // the post-pass: src/gravity_soft.h:245-372 (initial = false), :375-528 (initial = true)
    template <class Tpsys>
    void correct(Tpsys &pp, NeighborList &NList, PS::S32 &n_ngb_tot, PS::S32 &n_with_ngb, const bool initial)
    {
        const PS::S32 n = pp.getNumberOfParticleLocal();
        gplum_b200_corr_params prm;
        prm.eps2 = FP_t::eps2; prm.dt_tree = FP_t::dt_tree; prm.gamma = FP_t::gamma;
        prm.R_search2 = FP_t::R_search2; prm.R_search3 = FP_t::R_search3;
#ifdef USE_RE_SEARCH_NEIGHBOR
        prm.re_search = 1;
#else
        prm.re_search = 0;
#endif
        prm.reserved = 0;
        // the fields only the post-pass reads (src/gravity_soft.h:76-242: velocity, direct acceleration), of the
        // particles that occur in candidate pairs -- the ones the pass's download listed
        const PS::S32 m = n_listed_;
        vel_.resize(3 * (size_t)m); acc_d_.resize(3 * (size_t)m); id_.resize((size_t)m);
#pragma omp parallel for
        for (PS::S32 t = 0; t < m; t++) {
            EPJ_t e;
            e.copyFromFP(pp[nb_index_[t]]);
            vel_[3 * (size_t)t] = e.vel.x; vel_[3 * (size_t)t + 1] = e.vel.y; vel_[3 * (size_t)t + 2] = e.vel.z;
            acc_d_[3 * (size_t)t] = e.acc_d.x; acc_d_[3 * (size_t)t + 1] = e.acc_d.y; acc_d_[3 * (size_t)t + 2] = e.acc_d.z;
            id_[t] = e.id;       // the absorbed particle of a merger carries its target's id and position (src/collisionA.h:267-277)
        }
        tree_check(gplum_b200_tree_set_motion_sparse(m, nb_index_.data(), vel_.data(), acc_d_.data(), id_.data()), "gplum_b200_tree_set_motion_sparse");
        tree_check(gplum_b200_correct_long_run(0, &prm, initial ? 1 : 0), "gplum_b200_correct_long_run");
        corr_.resize(n);
        if (initial) init_.resize(n);
        const long long ngb_cap = 4LL * n + (1 << 20);
        ngb_.resize((size_t)ngb_cap);
        long long n_slots = 0, n_pairs = 0;
        tree_check(gplum_b200_correct_long_download(0, corr_.data(), initial ? init_.data() : nullptr, ngb_.data(), ngb_cap, &n_slots, &n_pairs),
                   "gplum_b200_correct_long_download");
        NList.initializeList(pp);
        n_ngb_tot = 0; n_with_ngb = 0;
        // records come in tree order; corr.id_local and ngb.id_local are particle indices (setIDLocalAndMyrank,
        // src/func.h:135-143), ngb.id the ids sent above.
        // Serial: NeighborList::addNeighbor appends to shared lists (the reference guards them with omp critical).
        for (PS::S32 k = 0; k < n; k++) {
            const gplum_b200_corr &c = corr_[k];
            const PS::S32 i = c.id_local;
            if (initial) {
#ifndef INTEGRATE_6TH_SUN
                calcStarGravity(pp[i]);
#else
                calcStarAccJerk(pp[i]);
#endif
            }
            pp[i].neighbor.number = 0;
            pp[i].id_cluster = pp[i].id;
            for (PS::S32 q = 0; q < c.number; q++) {
                const gplum_b200_ngb &b = ngb_[(size_t)c.ngb_off + q];
                NList.addNeighbor(pp, i, b.id, b.rank, b.id_local);      // number++, id_cluster = min, pair / exchange lists
            }
            if (pp[i].neighbor.number) {
                NList.with_neighbor_list.push_back(i);
                n_ngb_tot += pp[i].neighbor.number;
                n_with_ngb++;
            }
            pp[i].acc += PS::F64vec(c.acc[0], c.acc[1], c.acc[2]);
            pp[i].phi += c.phi;
            pp[i].acc0 = c.acc0;
            if (initial) {
                const gplum_b200_corr_init &d = init_[k];
                pp[i].acc_d = PS::F64vec(d.acc_d[0], d.acc_d[1], d.acc_d[2]);
                pp[i].phi_d = d.phi_d;
                pp[i].jerk_d = PS::F64vec(d.jerk_d[0], d.jerk_d[1], d.jerk_d[2]);
#ifndef INTEGRATE_6TH_SUN
                pp[i].calcDeltatInitial();
#endif
            }
        }
#ifdef INTEGRATE_6TH_SUN
        if (initial) {
#pragma omp parallel for
            for (PS::S32 i = 0; i < n; i++) { pp[i].setAcc_(); calcStarSnap(pp[i]); pp[i].calcDeltatInitial(); }
        }
#endif
    }
};

}  // namespace gplum_b200

// more specialised than the reference's templates over the tree type (src/gravity_soft.h:245-250,375-379)
template <class Tpsys>
void correctForceLong(Tpsys &pp, gplum_b200::TreeB200 &tree_grav, NeighborList &NList, PS::S32 &n_ngb_tot, PS::S32 &n_with_ngb)
{
    tree_grav.correct(pp, NList, n_ngb_tot, n_with_ngb, false);
}
template <class Tpsys>
void correctForceLongInitial(Tpsys &pp, gplum_b200::TreeB200 &tree_grav, NeighborList &NList, PS::S32 &n_ngb_tot, PS::S32 &n_with_ngb)
{
    tree_grav.correct(pp, NList, n_ngb_tot, n_with_ngb, true);
}
