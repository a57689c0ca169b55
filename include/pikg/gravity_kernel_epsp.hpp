// gravity_kernel_epsp.hpp -- drop-in for the PIKG-generated EP-SP kernel header
// (src/gravity_kernel.hpp:3-6,134-135; see gravity_kernel_epep.hpp in this directory).
#pragma once
#include "gravity_kernel_epep.hpp"

struct CalcForceLongEPSP {
    float eps2;
    explicit CalcForceLongEPSP(float eps2_) : eps2(eps2_) {}
    void operator()(const EPI_t *__restrict__ epi, const int ni, const SPJ_t *__restrict__ spj, const int nj,
                    Force_t *__restrict__ force, const int kernel_select = 1)
    {
        (void)kernel_select;
#ifdef USE_QUAD
        static_assert(sizeof(SPJ_t) == 80, "MySPJQuadrupole layout");
        const int quad = 1;
#else
        static_assert(sizeof(SPJ_t) == 32, "MySPJMonopole layout");
        const int quad = 0;
#endif
        gplum_b200_detail::check(gplum_b200_epsp(epi, ni, spj, nj, force, eps2, quad), "gplum_b200_epsp");
    }
};
