/* gplum_b200_lists.h -- C ABI of libgplum_lists.so: the HOST-side interaction-list builder
 * (gplum_b200/csrc/let_tree.cpp).  Workload tooling and the host mirror of the GPU list builder; not part of
 * libgplum_b200.so (the product library never builds lists on the host).
 *
 * The caller side of the path on one rank: Morton sort, octree (n_leaf_limit), moments, i-groups (n_group_limit)
 * and the symmetric-search tree walk exactly as FDPS does them -- root cell FDPS/src/tree_for_force_impl.hpp:770-868,
 * keys FDPS/src/key.hpp:118-226, cells tree_for_force_utils.hpp:289-420, groups :619-650, walk
 * FDPS/src/tree_walk.hpp:545-583,706-785 -- pinned list for list against the compiled reference (tests/test_tree.py).
 * pos is [n][3].  sizes[8] = n_walk, n_epi, n_adr_epj, n_adr_spj, n_epj_all, n_spj_all, n_int_epep, n_int_epsp.
 * tree_copy writes the reference's AoS layouts; any output pointer may be NULL.  Returns 0, or 2 (bad argument) /
 * 3 (no tree built). */
#ifndef GPLUM_B200_LISTS_H
#define GPLUM_B200_LISTS_H
#ifdef __cplusplus
extern "C" {
#endif
int gplum_b200_tree_build(int n, const double *pos, const double *mass, const double *r_out,
                          const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                          long long *sizes);
int gplum_b200_tree_copy(void *epi, int *epi_off, int *ni, int *adr_epj, long long *epj_disp, int *n_epj,
                         int *adr_spj, long long *spj_disp, int *n_spj, void *epj_all, void *spj_all,
                         int quad, int rank, int *sorted_to_original);
void gplum_b200_tree_free(void);
#ifdef __cplusplus
}
#endif
#endif
