// soft_corr.h -- launch interface of soft_corr.cu (device pointers only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/gplum_b200.h"

namespace gb {

typedef gplum_b200_corr SoftCorr;
typedef gplum_b200_corr_init SoftCorrInit;
typedef gplum_b200_ngb SoftNgb;
struct SoftCorrPrm { double eps2, dt_tree, gamma, R_search2, R_search3; int re_search, initial; };
static_assert(sizeof(SoftCorr) == 64 && sizeof(SoftCorrInit) == 64 && sizeof(SoftNgb) == 16, "result layout");

struct SoftCorrArgs {
    int n_epi;
    int i0, i1;                 // only i-particles [i0, i1) are corrected (multi-GPU: this rank's share); the rest get neutral records
    const void *epi;            // EPIGrav[n_epi], walk-concatenated
    const void *force;          // ForceGrav[n_epi] of the pass (number = candidate count)
    const void *epj_aos;        // EPJGrav[] of the pass, as FDPS holds epj_sorted_
    const int *self_adr;        // per i: EP index of the particle itself
    const int2 *pairs; const unsigned int *pair_count; unsigned int pair_cap;
    int *cnt, *off, *cursor;    // n_epi + 1 each
    int *csr;                   // pair_cap
    SoftCorr *out; SoftCorrInit *init_out; SoftNgb *ngb;
    unsigned int *status;       // [0] dropped pairs, [1] particles without self entry
    SoftCorrPrm prm;
};

size_t soft_corr_scan_temp_bytes(int n_epi);
size_t soft_corr_compact_temp_bytes(int n_epi);
// out[0 .. *n_out_dev) = the records of corr with number > 0, order kept; returns the CUDA error (0 = ok)
int soft_corr_compact(int n_epi, const SoftCorr *corr, SoftCorr *out, int *n_out_dev, void *temp, size_t temp_bytes, cudaStream_t st);
int soft_corr_launch(const SoftCorrArgs &a, void *scan_temp, size_t scan_temp_bytes, cudaStream_t st, int *launches);

}  // namespace gb
