// gravity_kernel_epep.hpp -- drop-in for the PIKG-generated EP-EP kernel header.
//
// GPLUM's src/gravity_kernel.hpp:3-6 includes "gravity_kernel_epep.hpp" when built with
// -DUSE_PIKG (src/Makefile:155-158 generates it with the Ruby PIKG compiler).  Put THIS
// directory on the quote-include path instead (-iquote <repo>/include/pikg -DUSE_PIKG) and the
// unmodified GPLUM sources compile against libgplum_b200.so: the struct below has the contract
// of the generated kernel (PIKG/src/parserdriver.rb:447-517): constructor taking the DSL's free
// scalars in declaration order (F32 eps2; src/gravity_kernel.hpp:17), and
// operator()(const EPI*, int ni, const EPJ*, int nj, FORCE*, int kernel_select = 1) that
// ACCUMULATES into force[0..ni).
//
// EPI_t / EPJ_t / SPJ_t / Force_t are the aliases src/main_p3t.cpp:38-48 defines before this
// header is reached.  Individual cut-off + cartesian coordinates only (the default macro set).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "../gplum_b200.h"

#if !defined(USE_INDIVIDUAL_CUTOFF) || defined(USE_POLAR_COORDINATE)
#error "libgplum_b200 implements the default macro set: USE_INDIVIDUAL_CUTOFF, cartesian coordinates"
#endif

namespace gplum_b200_detail {
inline void check(int rc, const char *what)
{
    if (rc != 0) {
        std::fprintf(stderr, "libgplum_b200: %s failed (%d): %s\n", what, rc, gplum_b200_last_error());
        PS::Abort(-1);
        std::abort();
    }
}
}  // namespace gplum_b200_detail

struct CalcForceLongEPEP {
    float eps2;
    explicit CalcForceLongEPEP(float eps2_) : eps2(eps2_) {}
    void operator()(const EPI_t *__restrict__ epi, const int ni, const EPJ_t *__restrict__ epj, const int nj,
                    Force_t *__restrict__ force, const int kernel_select = 1)
    {
        (void)kernel_select;
        static_assert(sizeof(EPI_t) == 48 && sizeof(EPJ_t) == 112 && sizeof(Force_t) == 32,
                      "particle.h layout differs from the one libgplum_b200 was built for");
        gplum_b200_detail::check(gplum_b200_epep(epi, ni, epj, nj, force, eps2), "gplum_b200_epep");
    }
};

#include "../gravity_kernel_b200.hpp"
