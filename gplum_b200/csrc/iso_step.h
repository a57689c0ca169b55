// iso_step.h -- launch interface of iso_step.cu (device pointers only): the isolated-particle half of a
// soft step on a device-resident particle state.
#pragma once
#include <cuda_runtime.h>

#include "../../include/gplum_b200.h"

namespace gbi {

// state = EPJGrav[n], particle k at slot k.  Tree-order inputs (epi / force / corr of a walk set) are
// scattered through id_local.
// vel[k] += half_dt * ((double)force[t].acc + corr[t].acc)      FPGrav::velKick, src/particle.h:878-884
int iso_kick(int n, void *state, const void *epi, const void *force, const void *corr, double half_dt, cudaStream_t st);
// isolated[k] / acc0[k] from the correction of a walk set (number == 0, acc0), tree order -> particle order
int iso_flags_from_corr(int n, const void *corr, int *isolated, double *acc0, cudaStream_t st);
// the loop of src/hard.h:793-817: Kepler drift of particles with isolated[k] != 0 and ecc < 0.8
int iso_drift(int n, void *state, double *time, double *dt, const double *acc0, const int *isolated,
              double t0, double t1, const gplum_b200_iso_params &prm, void *star, int *handled, cudaStream_t st);
// records of the particles the drift did not handle (neighbours, ecc >= 0.8): compacted for the host's hard part
int iso_pull_unhandled(int n, const void *state, const int *handled, void *rec_out, int *idx_out, int *count, int cap, cudaStream_t st);
int iso_push(int n_rec, const void *rec, const int *idx, void *state, cudaStream_t st);

}  // namespace gbi
