/* blastn_port.h — CPU restatement ("oracle port") of the blastn preliminary-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under gblastn_b200/ may include, link or load this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * Plain sequential C restating the reference algorithms (file:line cited at each function in
 * blastn_port.c).  It consumes the same BnQueryBatch the product's C ABI consumes and emits the
 * same record types, so a parity test is "same inputs, compare arrays".
 *
 * Pinning: this port is checked tap-by-tap against the reference engine itself
 * (oracle/_ref/libblastref.so, built from /root/reference by oracle/Makefile) in
 * tests/test_oracle_vs_reference.py, and against the committed golden vectors in tests/golden/.
 *
 * Known, deliberate simplification: the gapped-stage containment test uses a linear scan over
 * saved HSPs instead of the reference's interval tree (core/blast_itree.c).  The tree is an
 * exact index for containment, but its common-endpoint pruning on insertion
 * (core/blast_itree.c:330-470) has tree-shape-dependent corner cases; the port mirrors the
 * pruning rule on a flat list.  The product re-implements the tree itself.
 */
#ifndef GBLASTN_B200_ORACLE_PORT_H
#define GBLASTN_B200_ORACLE_PORT_H

#include "../../include/gblastn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PortResults {
    BnHSP     *hsps;    int64_t n_hsps;
    BnInitHit *init;    int64_t n_init;
    BnHSP     *gapped;  int64_t n_gapped;
    BnOffsetPair *scan; int32_t *scan_oid; int32_t *scan_chunk; int64_t n_scan;
    BnStats    stats;
} PortResults;

#define PORT_TAP_SCAN   1
#define PORT_TAP_INIT   2
#define PORT_TAP_GAPPED 4

int port_prelim_search(const BnQueryBatch *b, const uint8_t *packed, const int64_t *seq_byte_off,
                       const int32_t *seq_len, int32_t n_seq, int taps, PortResults *out);
/* same with database masks (blastn -db_soft_mask / -db_hard_mask): smask_type 1 soft / 2 hard, smask_n[i] masked
 * [begin, end) intervals of subject i, flat pairs in smask_iv (core/blast_engine.c:136-301, core/masksubj.inl) */
int port_prelim_search_masked(const BnQueryBatch *b, const uint8_t *packed, const int64_t *seq_byte_off,
                              const int32_t *seq_len, int32_t n_seq, int taps, int32_t smask_type,
                              const int32_t *smask_n, const int32_t *smask_iv, PortResults *out);
void port_results_free(PortResults *r);

/* unit-level entry points used by focused tests */
int port_greedy_align(const uint8_t *query, int32_t qlen, const uint8_t *subject_packed,
                      int32_t slen, int32_t q_off, int32_t s_off, int32_t reward, int32_t penalty,
                      int32_t xdrop, int32_t out[7]);
int port_dp_align(const uint8_t *query, int32_t qlen, const uint8_t *subject_packed, int32_t slen,
                  int32_t q_off, int32_t s_off, const int32_t *matrix16, int32_t gap_open,
                  int32_t gap_extend, int32_t xdrop, int32_t out[5]);

#ifdef __cplusplus
}
#endif
#endif
