// setup.cpp — host-side mirror of the set-up the reference performs before the hot path
// (BLAST_MainSetUp, LookupTableWrapInit, BLAST_GapAlignSetUp, BlastInitialWordParametersNew).
// PLACEHOLDER: filled in after the first GPU bring-up.
#include "../../include/gblastn_b200.h"

extern "C" {
int bn_setup_create(const BnSetupOptions *, int32_t, const uint8_t *, const int32_t *, const int32_t *,
                    const int32_t *, BnSetup **) { return BN_ERR_UNSUPPORTED; }
const BnQueryBatch *bn_setup_batch(const BnSetup *) { return nullptr; }
const double *bn_setup_kbp_std(const BnSetup *) { return nullptr; }
const double *bn_setup_kbp_gap(const BnSetup *) { return nullptr; }
int32_t bn_setup_gap_x_dropoff_final(const BnSetup *) { return 0; }
int32_t bn_setup_longest_chain(const BnSetup *) { return 0; }
void bn_setup_free(BnSetup *) {}
}
