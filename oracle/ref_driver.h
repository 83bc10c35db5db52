/* ref_driver.h — C interface of the reference-engine driver (TEST INFRASTRUCTURE).
 *
 * The driver (ref_driver.c) is OUR code; it links against the reference's own,
 * unmodified C sources compiled in place from /root/reference (see
 * oracle/Makefile) and exposes the taps SURVEY.md §7 step 0 asks for.
 * Nothing under gblastn_b200/ may include, link or load this.
 */
#ifndef GBLASTN_B200_ORACLE_REF_DRIVER_H
#define GBLASTN_B200_ORACLE_REF_DRIVER_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefConfig {
    int32_t task;            /* 0 = megablast, 1 = blastn (ws 11 DP) */
    int32_t word_size;       /* 0 => task default (28 / 11) */
    int32_t reward;          /* 0 => task default (1 / 2) */
    int32_t penalty;         /* 0 => task default (-2 / -3) */
    int32_t gap_open;        /* -1 => task default (0 / 5) */
    int32_t gap_extend;      /* -1 => task default (0 / 2) */
    int32_t greedy;          /* -1 => task default (1 / 0) */
    int32_t window_size;     /* two-hit window (default 0) */
    int32_t scan_range;      /* off-diagonal range (default 0) */
    int32_t min_diag_separation; /* -1 => task default (6 / 50) */
    int32_t hitlist_size;    /* 0 => 500 */
    int32_t mask_at_hash;    /* 1 => masks only affect lookup table (CLI default) */
    double  xdrop_ungap;     /* 0 => 20 bits */
    double  xdrop_gap;       /* 0 => 25 / 30 bits */
    double  xdrop_gap_final; /* 0 => 100 bits */
    double  evalue;          /* 0 => 10 */
    double  low_score_perc;  /* <0 => 0.15 */
    int64_t db_length;       /* 0 => real */
    int32_t db_num_seqs;     /* 0 => real */
    int32_t num_threads;     /* <=1 => single thread, taps allowed */
    int32_t taps;            /* bit0: scan pairs, bit1: init hits, bit2: gapped lists, bit3: lookup table dump,
                              * bit4: calls of the gapped-alignment-with-traceback routines (needs prelim_only 0) */
    int32_t prelim_only;     /* 1: preliminary stage only; 0: Blast_RunTracebackSearch follows (single thread) */
    /* database masks (blastn -db_soft_mask / -db_hard_mask): per subject smask_n[i] masked intervals,
     * flat pairs [begin, end) in smask_iv, ascending and disjoint; smask_type 0 none, 1 soft, 2 hard */
    int32_t smask_type;
    const int32_t *smask_n;
    const int32_t *smask_iv;
    int32_t hsp_num_max;     /* hit_options->hsp_num_max (0 = unlimited); ignored by gapped searches (BlastHspNumMax,
                              * core/blast_hits.c:169-191) */
    /* ambiguity data of the database (what CSeqDBVol::x_GetAmbigSeq lays over the 2-bit bases when a subject is
     * fetched in blastna for the traceback stage): amb_first[n_subjects + 1] indexes flat triples
     * {first base, bases, blastna code} in amb_runs, applied in order; NULL = none */
    const int64_t *amb_first;
    const int32_t *amb_runs;
    int32_t seam;            /* 0: the reference's own word finder and gapped stage; 1: the B200 engine behind the
                              * same two seams through oracle/shim (only in oracle/_ref/libblastshim.so) */
    double  percent_identity; /* hit_options->percent_identity (blastn -perc_identity), 0 = off */
    int32_t min_hit_length;  /* hit_options->min_hit_length, 0 = off */
} RefConfig;

/* Flat growable int32 table: rows x ncol */
typedef struct RefTable {
    int32_t *data;
    int64_t  rows;
    int64_t  cap;
    int32_t  ncol;
} RefTable;

typedef struct RefResult {
    /* taps (single-thread mode only) */
    RefTable scan;    /* oid, chunk_off, q_off, s_off                                   */
    RefTable init;    /* oid, chunk_off, q_off, s_off, q_start, s_start, length, score  */
    RefTable gapped;  /* oid, chunk_off, context, q_off, q_end, s_off, s_end, score, q_gapped_start, s_gapped_start */
    /* final per-subject lists as written to the HSP stream (absolute subject coords) */
    RefTable final_;  /* oid, context, q_off, q_end, s_off, s_end, score, q_gapped_start, s_gapped_start, evalue_lo, evalue_hi */
    /* parameters */
    int32_t  num_contexts;
    int32_t *ctx_query_offset, *ctx_query_length, *ctx_length_adjustment;
    int64_t *ctx_eff_searchsp;
    int32_t *ctx_x_dropoff, *ctx_cutoff_score, *ctx_reduced_cutoff;   /* ungapped (word_params) */
    int32_t *ctx_gapped_cutoff;                                          /* hit_params->cutoffs[].cutoff_score */
    double  *ctx_kbp_std;  /* 4 per context: Lambda K logK H */
    double  *ctx_kbp_gap;  /* 4 per context */
    int32_t  gap_x_dropoff, gap_x_dropoff_final;
    int32_t  container_type;   /* 0 = diag array, 1 = diag hash */
    int32_t  round_down;       /* sbp->round_down */
    int32_t  nucl_score_table[256];
    int32_t  matrix[16 * 16];
    /* lookup table */
    int32_t  lut_type;         /* 0 MB, 1 SmallNa, 2 Na */
    int32_t  lut_word_length, word_length, scan_step, longest_chain, pv_array_bts;
    int64_t  hashsize, next_pos_len, pv_len, overflow_len;
    int32_t *hashtable, *next_pos;      /* MB (tap bit3) */
    uint32_t *pv_array;
    int16_t *backbone, *overflow;       /* SmallNa (tap bit3) */
    int32_t  n_masked_locations;
    int32_t *masked_locations;          /* pairs left,right */
    int32_t  concat_len;                /* query->length */
    uint8_t *concat_query;              /* sequence_start: concat_len + 2 bytes */
    /* diagnostics */
    int64_t  lookup_hits, init_extends, good_init_extends, gap_extensions, good_extensions;
    double   seconds_prelim;            /* wall time of the preliminary search alone */
    int32_t  status;
    /* eNaLookupTable (lut_type 2, tap bit3): thick_backbone as 4 ints per cell {num_used, entries[3] | overflow_cursor}, overflow */
    int32_t *na_backbone, *na_overflow;
    int64_t  na_overflow_len;
    /* traceback stage (prelim_only 0).  Edit scripts live in tb_ops as (op, num) rows, op = EGapAlignOpType
     * (0 deletion = gap in query, 3 substitution, 6 insertion = gap in subject); esp_off / esp_n index it. */
    RefTable tb_calls;  /* kind (0 BLAST_GappedAlignmentWithTraceback, 1 BLAST_GreedyGappedAlignment with traceback),
                         * oid, context, s_shift (start_shift of AdjustSubjectRange), q_start, s_start, q_len, s_len,
                         * -> score, query_start, query_stop, subject_start, subject_stop, esp_off, esp_n */
    RefTable tb_ops;    /* op, num */
    RefTable tb_final;  /* query_index, oid, context, q_off, q_end, s_off, s_end, score, num_ident,
                         * evalue_lo, evalue_hi, bits_lo, bits_hi, esp_off, esp_n */
    double   seconds_traceback; /* wall time of Blast_RunTracebackSearch */
    /* the per-query hit lists the HSP stream holds when the preliminary stage ends (single thread, prelim_only 1):
     * what survives prelim_hitlist_size (core/hspfilter_collector.c:328-342, Blast_HitListUpdate) and reaches the
     * traceback stage.  Array order (a heap once a list overflowed): compare as sets. */
    RefTable kept;      /* query_index, oid, best score, number of HSPs */
} RefResult;

/* queries: blastna bytes (0..3 ACGT, 4..14 ambiguity) concatenated, lengths in qlens.
 * qmask: optional masked intervals in plus-strand query coordinates
 *        (qmask_n[i] intervals for query i; flat pairs [left,right] inclusive in qmask_iv).
 * db: ncbi2na packed bytes; sequence i starts at byte sbyteoff[i], has slen[i] bases. */
int ref_search(const RefConfig *cfg,
               int32_t n_queries, const uint8_t *qseq, const int32_t *qlens,
               const int32_t *qmask_n, const int32_t *qmask_iv,
               int32_t n_subjects, const uint8_t *packed, const int64_t *sbyteoff,
               const int32_t *slen,
               RefResult *res);

/* Direct calls of the reference's alignment-with-traceback routine (chosen by the configuration's traceback
 * algorithm) on arbitrary start points; items: 6 ints per call {oid, context, s_shift, s_length, q_start, s_start}.
 * Fills the parameter block, tb_calls and tb_ops of res. */
int ref_traceback_calls(const RefConfig *cfg,
                        int32_t n_queries, const uint8_t *qseq, const int32_t *qlens,
                        const int32_t *qmask_n, const int32_t *qmask_iv,
                        int32_t n_subjects, const uint8_t *packed, const int64_t *sbyteoff, const int32_t *slen,
                        int32_t n_items, const int32_t *items, RefResult *res);

void ref_free_result(RefResult *res);

#ifdef __cplusplus
}
#endif
#endif
