// gravity_kernel_b200.hpp -- the batched form of the drop-in: accelerator functors for FDPS's
// multi-walk-index interface (FDPS/src/tree_for_force.hpp:1528-1560,
// tree_for_force_impl_force.hpp:63-266), the same pair PIKG's CUDA back-end generates as
// Dispatch<K>/Retrieve<K> (PIKG/src/CUDA.rb:336-349,466-494).
//
// Use (the one change to src/main_p3t.cpp:351,583,672):
//   tree_grav.calcForceAllAndWriteBackMultiWalkIndex(DispatchKernelB200(), RetrieveKernelB200(),
//           1, system_grav, dinfo, GPLUM_B200_N_WALK_LIMIT, true, MY_INTERACTION_LIST_MODE);
// FDPS calls dispatch once with send_flag=true (all of epj_sorted_/spj_sorted_), then per batch of
// walks dispatch(...,false) followed one batch later by retrieve, which accumulates into
// force_sorted_ (cleared by FDPS when clear=true).
#pragma once
#include "gplum_b200.h"

#ifndef GPLUM_B200_N_WALK_LIMIT
#define GPLUM_B200_N_WALK_LIMIT 4096
#endif

struct DispatchKernelB200 {
    PS::S32 operator()(const PS::S32 tag, const PS::S32 n_walk, const EPI_t **epi, const PS::S32 *n_epi,
                       const PS::S32 **id_epj, const PS::S32 *n_epj, const PS::S32 **id_spj, const PS::S32 *n_spj,
                       const EPJ_t *epj, const PS::S32 n_epj_tot, const SPJ_t *spj, const PS::S32 n_spj_tot,
                       const bool send_flag) const
    {
#ifdef USE_QUAD
        const int quad = 1;
#else
        const int quad = 0;
#endif
        if (send_flag) gplum_b200_set_params((float)FP_t::eps2, quad, -1);
        const int rc = gplum_b200_dispatch(tag, n_walk, (const void *const *)epi, n_epi, (const int *const *)id_epj, n_epj,
                                           (const int *const *)id_spj, n_spj, epj, n_epj_tot, spj, n_spj_tot, send_flag ? 1 : 0);
        if (rc != 0) {
            std::fprintf(stderr, "libgplum_b200: dispatch failed (%d): %s\n", rc, gplum_b200_last_error());
            PS::Abort(-1);
        }
        return 0;
    }
};

struct RetrieveKernelB200 {
    PS::S32 operator()(const PS::S32 tag, const PS::S32 n_walk, const PS::S32 *ni, Force_t **force) const
    {
        const int rc = gplum_b200_retrieve(tag, n_walk, ni, (void *const *)force);
        if (rc != 0) {
            std::fprintf(stderr, "libgplum_b200: retrieve failed (%d): %s\n", rc, gplum_b200_last_error());
            PS::Abort(-1);
        }
        return 0;
    }
};
