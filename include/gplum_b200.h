/*
 * gplum_b200.h -- C ABI of libgplum_b200.so, the B200 (sm_100a) replacement for GPLUM's P3T
 * soft-force interaction stage.  Plain pointers and sizes only; no C++/torch types.
 *
 * What each entry point replaces in the reference (paths relative to the GPLUM tree):
 *
 *   gplum_b200_epep      calcForceEPEPWithSearch::operator()(epi,ni,epj,nj,force)
 *                        src/gravity_kernel.hpp:12-23 (PIKG: CalcForceLongEPEP, generated from
 *                        src/gravity_kernel_epep.pikg), incl. neighbour-candidate detection.
 *   gplum_b200_epsp      calcForceEPSP::operator()(epi,ni,spj,nj,force)
 *                        src/gravity_kernel.hpp:128-136 (src/gravity_kernel_epsp.pikg).
 *   gplum_b200_dispatch  the accelerator "dispatch" functor of FDPS's multi-walk-index interface,
 *   gplum_b200_retrieve  FDPS/src/tree_for_force_impl_force.hpp:78-83 (send all), :232-238
 *                        (dispatch), :230,242 (retrieve); what PIKG's CUDA back-end generates as
 *                        Dispatch<K>/Retrieve<K> (PIKG/src/CUDA.rb:336-349,466-494).
 *   gplum_b200_calc_walks  one whole TreeForForce::calcForce pass
 *                        (FDPS/src/tree_for_force_impl_force.hpp:1404-1564) over flat arrays:
 *                        gather by index (CopyPjForForceST, tree_for_force_impl.hpp:652-693),
 *                        ForceGrav::clear (src/particle.h:81-85), EP-EP, EP-SP, write-back.
 *
 * Struct layouts are the reference's, default macro set (USE_INDIVIDUAL_CUTOFF, USE_QUAD,
 * cartesian): EPIGrav 48 B, EPJGrav 112 B, MySPJQuadrupole 80 B, MySPJMonopole 32 B,
 * ForceGrav 32 B (src/particle.h:16-24,70-86,93-109,149-156; FDPS/src/tree.hpp:883-967).
 *
 * All pointers are HOST pointers owned by the caller unless a name says "dev".  The library
 * owns device buffers and pinned staging and grows them on demand.  Every function returns 0 on
 * success or a non-zero code (gplum_b200_last_error() describes it); nothing throws.  There is
 * no CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef GPLUM_B200_H
#define GPLUM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPLUM_B200_ABI_VERSION 3

/* mode flags (gplum_b200_set_params) */
#define GPLUM_B200_TRACE_AS_SHIPPED 1 /* tr = qxx+qyy+qxx, the non-PIKG branch's arithmetic
                                         (src/gravity_kernel.hpp:177); default is the DSL's
                                         qxx+qyy+qzz (src/gravity_kernel_epsp.pikg:75) */
#define GPLUM_B200_RANK_SQUARED     2 /* rank += (ranki-rankj)^2 as in the DSL (.pikg:77-83);
                                         default |ranki-rankj| (gravity_kernel.hpp:108) */
#define GPLUM_B200_NO_ACCUMULATE    4 /* retrieve/epep/epsp overwrite instead of accumulating */

/* error codes */
#define GPLUM_B200_OK            0
#define GPLUM_B200_ERR_CUDA      1
#define GPLUM_B200_ERR_NO_DEVICE 2
#define GPLUM_B200_ERR_ARG       3
#define GPLUM_B200_ERR_STATE     4

int gplum_b200_abi_version(void);
const char *gplum_b200_last_error(void);

/* Select the device, create streams, pre-size buffers for max_i i-particles and max_j list
 * entries per dispatch (hints; buffers grow).  Idempotent for the same device. */
int gplum_b200_init(int device, size_t max_i, size_t max_j);
int gplum_b200_finalize(void);

/* eps2: FP_t::eps2 narrowed to F32 (src/gravity_kernel.hpp:17,31); quad: 1 = MySPJQuadrupole
 * (USE_QUAD), 0 = MySPJMonopole; flags: GPLUM_B200_* above, or < 0 to keep the current flags
 * (initialised from the environment variable GPLUM_B200_FLAGS, default 0). */
int gplum_b200_set_params(float eps2, int quad, int flags);

/* ---- per-call functor form (re-entrant; one stream + staging slot per calling thread) ---- */
int gplum_b200_epep(const void *epi, int ni, const void *epj, int nj, void *force, float eps2);
int gplum_b200_epsp(const void *epi, int ni, const void *spj, int ns, void *force, float eps2, int quad);

/* ---- FDPS multi-walk-index form (called from the master thread) ----
 * send_all != 0: upload epj_all / spj_all (n_walk is 0, list arguments ignored).
 * send_all == 0: enqueue n_walk walks; returns as soon as the work is queued on the GPU.
 * retrieve(tag, ...) waits for the dispatch with the same tag and accumulates into force[w]. */
int gplum_b200_dispatch(int tag, int n_walk, const void *const *epi, const int *ni,
                        const int *const *adr_epj, const int *n_epj,
                        const int *const *adr_spj, const int *n_spj,
                        const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all,
                        int send_all);
int gplum_b200_retrieve(int tag, int n_walk, const int *ni, void *const *force);

/* ---- flat form: one whole force pass, host buffers in, host forces out (synchronous) ----
 * epi_off[w], ni[w]: slice of epi_all / force_all of walk w; adr_epj[epj_disp[w] .. +n_epj[w])
 * index epj_all; same for spj.  clear != 0 overwrites force (FDPS clear=true), else accumulates. */
int gplum_b200_calc_walks(int n_walk, const void *epi_all, const int *epi_off, const int *ni,
                          const int *adr_epj, const long long *epj_disp, const int *n_epj,
                          const int *adr_spj, const long long *spj_disp, const int *n_spj,
                          const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all,
                          void *force_all, int clear);

/* ---- device-resident form (bench / multi-GPU): upload once, run many times ----
 * walks_upload copies lists + particles to HBM and builds the work list; walks_run launches the
 * j-pack + force kernels on the library stream (asynchronous); walks_download synchronises
 * and copies the n_epi_total forces back.  walks_time runs `iters` passes bracketed by CUDA
 * events on the launching stream and returns the mean milliseconds per pass. */
/* Up to 4 resident walk sets (e.g. interior / boundary walks of a domain); walks_upload, _run,
 * _download and _time act on the selected one.  walks_upload with epj_all == spj_all == NULL
 * keeps the current j-set. */
int gplum_b200_walks_select(int slot);
/* i-particles per work item for the walk sets uploaded / dispatched next: 64 (two per lane, the most
 * efficient), 32, or 16 / 8 / 4 (j-lists split over lane groups); 0 = chosen per pass: 64 unless the
 * pass has too few items to occupy a quarter of the GPU's warp slots (latency-bound small passes). */
int gplum_b200_set_tile_cap(int cap);
int gplum_b200_walks_upload(int n_walk, const void *epi_all, const int *epi_off, const int *ni,
                            const int *adr_epj, const long long *epj_disp, const int *n_epj,
                            const int *adr_spj, const long long *spj_disp, const int *n_spj,
                            const void *epj_all, int n_epj_all, const void *spj_all, int n_spj_all);
int gplum_b200_walks_run(int repack);
int gplum_b200_walks_pack(void);          /* only the j-pack kernels of walks_run(repack=1) */
int gplum_b200_walks_download(void *force_all);
int gplum_b200_walks_time(int iters, int repack, float *ms_per_pass);
/* Replace the packed j-arrays by caller-owned device buffers (e.g. the output of an NCCL
 * all-gather); layouts are gplum_b200_packed_sizes().  Pass NULL to return to library buffers. */
int gplum_b200_walks_set_packed_dev(const void *epj_packed_dev, int n_epj_all,
                                    const void *spj_packed_dev, int n_spj_all);
/* Pack n raw AoS j-particles already in HBM into the packed layout (device pointers; runs on
 * the library stream): the per-rank step before the all-gather. */
int gplum_b200_pack_epj_dev(const void *epj_aos_dev, int n, void *epj_packed_dev);
int gplum_b200_pack_spj_dev(const void *spj_aos_dev, int n, void *spj_packed_dev);
/* dst[k] = src[idx[k]] on packed EP records (device pointers): stages the particles another rank's
 * boundary walks need (the trimmed LET exchange) before an all-to-all. */
int gplum_b200_gather_epj_packed_dev(const void *src_packed_dev, const int *idx_dev, int n, void *dst_packed_dev);
void gplum_b200_packed_sizes(int *epj_packed_bytes, int *spj_packed_bytes);

/* ---- multi-GPU peer mode: boundary walks read the other ranks' packed EPJ straight over NVLink ----
 * Every rank owns two slabs (double buffer) of 2^shift packed records; EP list indices are
 * (owner_rank << shift) | index_in_owner_slab.  peer_setup allocates the slabs and writes this
 * rank's two CUDA IPC handles (2 x 64 bytes) to handles_out; the caller all-gathers the handles
 * (rank-major) and passes them to peer_open.  peer_pack flips the buffer and, in ONE launch on the library
 * stream, packs n AoS records (device pointer) into this rank's slab, packs the current j-set's
 * superparticles, and stores the new epoch into this rank's entry of EVERY rank's flag array (kept
 * behind slab 0; remote ones over NVLink).  Walk sets uploaded while peer mode is open know which of
 * their walks name another rank's particles: those walks' work items spin INSIDE the force kernel until
 * every rank's flag has reached the epoch of the last peer_pack, the others start at once, and one empty
 * barrier item keeps the pass from ending before that -- pack + one force launch is a whole step, with
 * no collective library call and no second stream on the path.  Because a rank only packs epoch e+1
 * after its own pass of epoch e, having seen all flags at e+1 also means every peer is done reading this
 * rank's slab of epoch e -- the double buffer needs no second barrier.  peer_wait enqueues the same wait
 * as a one-warp kernel (for callers that launch their own kernels on peer data).
 * Every rank must call peer_pack the same number of times.  Replaces the EPJ part of FDPS's LET
 * exchange (FDPS/src/tree_for_force_impl_exlet.hpp:343-403) without moving data ahead of time.
 * peer_close unmaps, peer_free (after a barrier) releases the slabs. */
int gplum_b200_peer_setup(int world, int rank, int shift, void *handles_out);
int gplum_b200_peer_open(const void *all_handles);
int gplum_b200_peer_pack(const void *epj_aos_dev, int n);
int gplum_b200_peer_wait(void);
int gplum_b200_peer_close(void);
int gplum_b200_peer_free(void);

/* ---- changeover correction + final neighbour lists (next row of the path: SURVEY 8f-2) ----
 * Replaces correctForceLong / correctForceLongInitial (src/gravity_soft.h:245-372,375-528), i.e.
 * correctForceBetween2Particles{,Initial} (:76-153,155-242) with cutoff_W/K/dKdt (src/cutfunc.h),
 * for the i-particles of a resident walk set.  The reference re-searches the tree for every
 * particle with candidates (getNeighborListOneParticle, :295-297); here the force pass records
 * the candidate pairs itself while it runs, so the correction is three small kernels after it.
 *   1. gplum_b200_soft_corr_enable(1, cap) BEFORE the pass is launched (cap = pair-buffer capacity,
 *      0 = 4 x n_epi + 2^20); capture costs nothing in the pair loop, only the rare path writes.
 *   2. run the pass (walks_run / calc_walks / dispatch).
 *   3. gplum_b200_correct_long_run(slot, prm, initial) launches the correction on the library
 *      stream; gplum_b200_correct_long_download copies the results to host buffers.
 * Results are per i-particle in the walk-concatenated order of epi/force:
 *   gplum_b200_corr.acc/phi are the FP64 sums `acci`, `phii` the reference ADDS to the widened tree
 *   force (:366-367); acc0, number, id_cluster, in_domain as src/gravity_soft.h:368 and
 *   NeighborList::addNeighbor (src/neighbor.h:635-664); the particle's neighbours are
 *   ngb[ngb_off .. ngb_off+number) = NeighborId{id, rank, id_local} (src/neighbor.h:205-245), in
 *   ascending EP-index order; id_local is the pp index FDPS's write-back targets.
 * Needs the raw EPJGrav array of the pass on the device (every form that takes epj_all keeps it). */
typedef struct { double eps2, dt_tree, gamma, R_search2, R_search3; int re_search, reserved; } gplum_b200_corr_params;
typedef struct { double acc[3]; double phi; double acc0; long long id_cluster;
                 int number, id_local, ngb_off, in_domain; } gplum_b200_corr;              /* 64 B */
typedef struct { double acc_d[3]; double jerk_d[3]; double phi_d; double pad; } gplum_b200_corr_init; /* 64 B */
typedef struct { long long id; int rank, id_local; } gplum_b200_ngb;                      /* 16 B */
#define GPLUM_B200_ERR_OVERFLOW 5  /* pair buffer too small: enlarge (soft_corr_enable) and redo the pass */
int gplum_b200_soft_corr_enable(int on, long long pair_cap);
int gplum_b200_correct_long_run(int slot, const gplum_b200_corr_params *prm, int initial);
/* n_ngb_slots: entries of ngb to copy = ngb_off + number of the last particle with candidates (the
 * list is segmented by candidate counts); n_ngb_total: neighbours actually present.  init_out may be
 * NULL.  ngb_cap = capacity of ngb_out in entries. */
int gplum_b200_correct_long_download(int slot, void *corr_out, void *init_out, void *ngb_out,
                                     long long ngb_cap, long long *n_ngb_slots, long long *n_pairs);
/* Same, but only the records of particles that HAVE neighbours are copied (in walk order); every other particle
 * carries the self term alone (acc = 0, phi = mass / r_out, acc0 = 0, id_cluster = id, number = 0,
 * src/gravity_soft.h:280,368), which the caller applies itself.  *n_corr = records written (<= corr_cap). */
int gplum_b200_correct_long_download_compact(int slot, void *corr_out, long long corr_cap, long long *n_corr,
                                             void *ngb_out, long long ngb_cap, long long *n_ngb_slots, long long *n_pairs);
/* mean milliseconds of `iters` correction launches (CUDA events on the library stream) */
int gplum_b200_correct_long_time(int slot, const gplum_b200_corr_params *prm, int initial, int iters, float *ms);

/* Use an existing CUDA stream (cudaStream_t as void*) for the batched / device-resident
 * forms, so that a caller's events on that stream bracket the kernels.  NULL = own stream. */
int gplum_b200_set_stream(void *cuda_stream);
int gplum_b200_synchronize(void);

/* Counters since the last reset: kernels launched by this library, interactions evaluated
 * (sum ni*(n_epj+n_spj), FDPS's n_interaction_ep_ep_local_ + n_interaction_ep_sp_local_). */
void gplum_b200_counters(long long *kernel_launches, long long *n_epep, long long *n_epsp, int reset);

/* ---- interaction-list builder on the GPU (gplum_b200/csrc/dev_tree.cu) ----
 * Same semantics, cell numbering and FP64 results as the host builder (include/gplum_b200_lists.h), but every stage runs on the
 * device: Morton keys, radix sort, cells level by level (FDPS LinkCell, tree_for_force_utils.hpp:289),
 * moments + in/out boxes bottom-up (utils_moment.hpp:6,189), i-groups (MakeIPGroup, utils.hpp:619-650),
 * one warp per group for the symmetric-search walk (tree_walk.hpp:545-583,706-785), work items.  The
 * result is not copied anywhere: it becomes the selected resident walk set and the current j-set (EPJ in
 * tree order, SPJ = cell moments, both packed), ready for gplum_b200_walks_run / _download and the
 * changeover correction.  Lists equal the host builder's as sets; within a list the order differs.
 *   tree_build_gpu      host columns (pos is [n][3]), 48 B per particle over PCIe; EPJ records get
 *                       id_local = id = input index, myrank = rank, vel = acc_d = 0.  The positions go up first;
 *                       the other columns follow on the copy engine while the GPU computes keys and sorts.
 *   tree_build_gpu_vel  the same plus the velocities ([n][3], 72 B per particle): what the changeover correction's
 *                       neighbour re-search reads (src/gravity_soft.h:300-346).  This is the form
 *                       include/gravity_tree_b200.hpp binds behind calcForceAllAndWriteBack.
 *   tree_build_gpu_epj  EPJGrav[n] in any order (FDPS's epj_org_), host or device memory (16 B aligned);
 *                       records are carried whole, so the correction sees vel / acc_d / id.
 *   tree_copy_gpu       copies lists / particles back to host arrays (tests; any pointer may be NULL).
 *   tree_gpu_times      device milliseconds of the last build: keys+sort+gather, cells+moments, i-group
 *                       compaction, counting walk+scans, filling walk, items+SPJ. */
int gplum_b200_tree_build_gpu(int n, const double *pos, const double *mass, const double *r_out,
                              const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                              int rank, long long *sizes);
int gplum_b200_tree_build_gpu_vel(int n, const double *pos, const double *vel, const double *mass, const double *r_out,
                                  const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                                  int rank, long long *sizes);
int gplum_b200_tree_build_gpu_epj(int n, const void *epj, int on_device, double theta, int n_leaf_limit,
                                  int n_group_limit, long long *sizes);
/* Multi-GPU form: every rank passes the SAME n EPJGrav records (device pointer -- the all-gather of all
 * ranks' particles, FDPS's LET exchange FDPS/src/tree_for_force_impl_exlet.hpp:343-403 turned into one NCCL
 * all-gather over NVLink) and builds the same tree, but walks, lists and work items only for its share: the walks
 * whose first particle in tree order lies in [n r / W, n (r+1) / W) -- a Morton-contiguous spatial domain.  Forces
 * (walks_run) and corrections (correct_long_run) then exist for that share only.  sizes[12]: [0..7] as
 * tree_build_gpu for THIS rank's lists, [8], [9] = its walks [w0, w1), [10], [11] = its i-particles [e0, e1). */
int gplum_b200_tree_build_gpu_part(int n, const void *epj_dev, double theta, int n_leaf_limit, int n_group_limit,
                                   int part_rank, int part_world, long long *sizes);
/* The same from 48 B records {pos[3], mass, r_out, r_search} (device pointer: the all-gather of all ranks' records) --
 * the fields the interaction kernels read, 48 instead of 112 B per particle over NVLink.  id_local = id = index in the
 * gathered array, myrank = 0, vel = acc_d = 0 (a following correction needs gplum_b200_tree_set_motion*). */
int gplum_b200_tree_build_gpu_part_rec48(int n, const double *rec_dev, double theta, int n_leaf_limit, int n_group_limit,
                                         int part_rank, int part_world, long long *sizes);
/* ForceGrav[count] of i-particles [first, first + count) of the selected walk set (tree order); host pointer. */
int gplum_b200_walks_download_range(void *force_out, long long first, long long count);
int gplum_b200_tree_copy_gpu(void *epi, int *epi_off, int *ni, int *adr_epj, long long *epj_disp, int *n_epj,
                             int *adr_spj, long long *spj_disp, int *n_spj, void *epj_all, void *spj_all,
                             int *sorted_to_original);
int gplum_b200_tree_gpu_times(float *ms6);
/* ForceGrav[n] of the last pass over the GPU-built tree, in the order the particles were handed to
 * tree_build_gpu / tree_build_gpu_epj (FDPS's copyForceOriginalOrder + writeBack,
 * FDPS/src/tree_for_force_impl.hpp:873-883, tree_for_force.hpp:148-153); host pointer, synchronises. */
int gplum_b200_tree_download_original(void *force_out);
/* Velocities and direct accelerations of the particles of the last GPU build -- columns [n][3] in the order the
 * particles were handed in, either may be NULL (zeros) -- written into the resident tree-order EPJGrav records: what
 * gplum_b200_correct_long_run reads besides the positions (relative velocity and acc_d difference of the neighbour
 * re-search, jerk of the `initial` form; src/gravity_soft.h:76-242).  tree_build_gpu + tree_set_motion carry the same
 * fields as tree_build_gpu_epj, split the way the reference splits calcForceAllAndWriteBack and correctForceLong. */
int gplum_b200_tree_set_motion(int n, const double *vel, const double *acc_d);
/* The same for m listed particles: index[t] = particle (as handed in), vel / acc_d are [m][3].  The post-pass reads
 * the motion only of particles that occur in candidate pairs; gplum_b200_tree_download_compact lists exactly those
 * while the candidate capture (gplum_b200_soft_corr_enable) is on.  id (optional, [m]): the particles' ids.  The tree
 * built from columns numbers particles by index; the post-pass treats a candidate with the particle's own id as the
 * particle itself (src/gravity_soft.h:105-108) -- after a merging collision the absorbed particle carries the id and the
 * position of its target until MergeParticle removes it (src/collisionA.h:267-277), so a caller whose ids can repeat
 * must pass them. */
int gplum_b200_tree_set_motion_sparse(int m, const int *index, const double *vel, const double *acc_d, const long long *id);
/* The same for a caller that holds whole columns (vel_all / acc_d_all are [n][3] in particle order): the library
 * gathers the m listed particles with OpenMP into pinned staging and sends only those. */
int gplum_b200_tree_set_motion_gather(int m, const int *index, const double *vel_all, const double *acc_d_all, const long long *id_all);

/* Page-locked host memory for a caller's staging arrays (NULL on failure, gplum_b200_last_error has the reason):
 * arrays handed to the tree_build_gpu / tree_download / tree_set_motion calls from such memory cross PCIe directly;
 * pageable ones are staged through the library's own pinned chunks (correct, 1.4x slower end to end at N = 1e6). */
void *gplum_b200_pinned_alloc(size_t bytes);
void gplum_b200_pinned_free(void *p);

/* The same results with about half the bytes over PCIe.  accphi_out[4 i .. 4 i + 3] = {acc, phi} of particle i (in
 * the order the particles were handed in); the neighbour words of ForceGrav come back only for the *n_nb_out
 * particles that have candidates (number > 0): nb_index_out[k] = particle, nb_out[4 k ..] = {number, rank, id_max,
 * id_min}.  Every other particle has exactly what gplum_b200_force_clear / ForceGrav::clear() writes
 * (src/particle.h:81-85).  While the candidate capture is on, the list also names every particle that occurs in a
 * captured pair from one side only (FP32 borderline; its words are the cleared ones): the set whose motion the
 * post-pass needs.  cap = room in nb_index_out / nb_out (entries); n is always enough. */
int gplum_b200_tree_download_compact(float *accphi_out, int *nb_index_out, int *nb_out, int cap, int *n_nb_out);
/* diagnostics: %globaltimer (ns) at the level boundaries inside the cells+moments kernel of the last build;
 * returns the number of stamps written (<= cap) */
int gplum_b200_tree_gpu_stamps(unsigned long long *ns_out, int cap);

/* ---- device-resident particle state: the isolated-particle half of a soft step ----
 * The state is EPJGrav[n] with particle k at slot k (id_local == k, myrank = this rank) plus FPGrav::time
 * and FPGrav::dt per particle.  With it a whole soft step runs without the particles leaving HBM:
 *   state_kick   FPGrav::velKick (src/particle.h:878-884): vel += 0.5*dt_tree*acc with
 *                acc = (F64)ForceGrav::acc [+ the changeover correction] of the walk set in `slot`
 *                (src/particle.h:761-766, src/gravity_soft.h:366-367);
 *   state_drift  the isolated-particle loop of the hard part (src/hard.h:793-817): particles without
 *                neighbours and with eccentricity < 0.8 move along their Kepler orbit
 *                (timeIntegrateKepler_isolated, src/hermite.h:787-816; src/kepler.h), get phi_s/acc_s/jerk_s
 *                (calcStarGravity, src/gravity_hard.h:5-39) and their next hard time step (calcDeltatInitial,
 *                src/particle.h:886-914).  isolated / acc0 come from the correction of `slot`
 *                (number == 0, acc0), or from host arrays (particle order) when `isolated` != NULL;
 *   state_pull_unhandled / state_push   the particles the drift did not take (neighbours -> hard clusters,
 *                eccentric orbits -> Hermite) as EPJGrav records + indices, for the host's hard part, and back;
 *   state_tree_build = gplum_b200_tree_build_gpu_epj on the resident state. */
typedef struct { double m_sun, dt_tree, eta_0, eta_sun0, alpha2, dt_min, eps2_sun; } gplum_b200_iso_params;
typedef struct { double phi_s; double acc_s[3]; double jerk_s[3]; double dt; } gplum_b200_star;        /* 64 B */
int gplum_b200_state_upload(int n, const void *epj, const double *time, const double *dt);
int gplum_b200_state_download(void *epj_out, double *time, double *dt, void *star_out, int *handled_out);
int gplum_b200_state_tree_build(double theta, int n_leaf_limit, int n_group_limit, long long *sizes);
int gplum_b200_state_kick(int slot, int use_corr, double dt_tree);
int gplum_b200_state_drift(const gplum_b200_iso_params *prm, double t0, double t1, int slot,
                           const int *isolated, const double *acc0);
int gplum_b200_state_pull_unhandled(void *rec_out, int *idx_out, int cap, int *n_out);
int gplum_b200_state_push(const void *rec, const int *idx, int n_rec);

/* The host work-list builder of the force pass, for tests (needs no device): cuts every walk into i-tiles and, in
 * a pass with less than two waves of them (warp_slots), lays the tiles out as ONE wave of warp_slots equal-cost
 * segments, cutting full-width tiles along j where a segment boundary falls.  One item = 8 ints
 * {walk, i0, ni, cfg, t0, t1, slot0, group}: cfg & 15: 0 = 32 i, 1 = 64 i, 9/10/11 = 16/8/4 i with the j-list split
 * over 2/4/8 lane groups; cfg bits 8-15 = K parts of this tile (0: whole tile), bits 16-23 = part index; [t0, t1) =
 * the part's j-tiles of the walk's sequence (EP tiles of 64, then SP tiles; t1 < 0: all); slot0 + part index = its
 * scratch record, group = its tile's arrival counter.  Longest base tile first.  seg_off_out (n_seg + 1 entries, n_seg
 * = 0 when the pass is not segmented): warp s executes items [seg_off[s], seg_off[s+1]).  split_m = 0: never segment. */
int gplum_b200_debug_build_items(int n_walk, const int *ni, const int *n_epj, const int *n_spj,
                                 long long warp_slots, int tile_cap, int jsplit, int split_m,
                                 int *items_out, int cap_items, int *n_items_out, int *n_slots_out, int *n_groups_out,
                                 int *seg_off_out, int cap_seg, int *n_seg_out);

/* Test / profiling support: with `on`, every following force pass records for each work item (in list order) when and
 * where it ran: {start, end} in globaltimer nanoseconds, %smid | %warpid << 32, and how many times it was run (4 x 64 bit
 * per item; barrier items of the peer mode leave zeros).  `out` != NULL first copies the last pass's records
 * (cap_items = room in items). */
int gplum_b200_debug_trace(int on, unsigned long long *out, int cap_items, int *n_items_out);

/* FP32 FMA issue-rate microbenchmark (the roofline denominator measured in the same run):
 * returns achieved FFMA TFLOP/s (2 flop per FFMA) over `iters` launches. */
int gplum_b200_fp32_peak(int iters, float *tflops, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* GPLUM_B200_H */
