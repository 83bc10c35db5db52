/* gblastn_b200.h — C ABI of the B200-native blastn preliminary-search hot path.
 *
 * Drop-in boundary.  Plain C: POD structs, pointers and sizes only.
 * Every entry point names the reference interface it replaces; paths are relative to the
 * reference tree (core/ = c++/src/algo/blast/core, inc-core/ = c++/include/algo/blast/core,
 * gpu/ = c++/src/algo/blast/gpu_blast, inc-gpu/ = c++/include/algo/blast/gpu_blast).
 *
 * Layering, mirroring where the reference draws its own lines:
 *
 *   bn_init / bn_release            <->  Blast_gpu_Init / Blast_gpu_Release
 *                                        (inc-gpu/gpu_blastn.h:50-51, app/blast/blastn_app.cpp:462,491)
 *   bn_db_load / bn_db_free         <->  the per-GPU subject cache G-BLASTN fills on first sight
 *                                        of an oid (gpu/gpu_blastn_MB_and_smallNa.cu:1462-1468);
 *                                        input is what BlastSeqSrcGetSequence hands the engine
 *                                        (core/blast_engine.c:1203; packed ncbi2na, inc-core/blast_def.h:242)
 *   bn_db_set_masks                 <->  BlastSeqBlkSetSeqRanges (core/blast_util.c:186-223) as called by the seqsrc
 *                                        for -db_soft_mask / -db_hard_mask
 *   bn_db_load_files                <->  CSeqDBVol reading .nin/.nsq (objtools/blast/seqdb_reader/seqdbvol.cpp)
 *   bn_query_load / bn_query_free   <->  GpuLookUpSetUp / gpu_InitQueryMemory
 *                                        (gpu/gpu_blastn_na_ungapped_v3.cpp:595-696): the arrays of
 *                                        LookupTableWrap (inc-core/blast_nalookup.h:60,236), the query
 *                                        BLAST_SequenceBlk, BlastQueryInfo contexts, and the derived
 *                                        BlastInitialWordParameters / BlastExtensionParameters /
 *                                        BlastHitSavingParameters values (inc-core/blast_parameters.h)
 *   bn_prelim_search                <->  BLAST_PreliminarySearchEngine's subject loop
 *                                        (core/blast_engine.c:1187-1330): per subject chunk
 *                                        BlastNaWordFinder (core/na_ungapped.c:1559) +
 *                                        BLAST_GetGappedScore (core/blast_gapalign.c:3233) +
 *                                        the post-processing of s_BlastSearchEngineOneContext
 *                                        (core/blast_engine.c:503-540) and E-values (:788-806)
 *   bn_word_finder                  <->  BlastWordFinderType  (inc-core/blast_engine.h:227-238)
 *   bn_get_gapped_score             <->  BlastGetGappedScoreType (inc-core/blast_engine.h:212-224)
 *   bn_prelim_search_batches        <->  blastn's query-batch loop (app/blast/blastn_app.cpp:574-640) / G-BLASTN's
 *                                        work_thread pipeline (gpu/work_thread.cpp:16-438)
 *   bn_scan_subject                 <->  TNaScanSubjectFunction (inc-core/blast_nascan.h:43-47), whole-
 *                                        subject form (max_hits batching is invisible to results)
 *   bn_setup_*                      <->  host-side set-up the reference does before the path:
 *                                        BLAST_MainSetUp / LookupTableWrapInit / BLAST_GapAlignSetUp /
 *                                        BlastInitialWordParametersNew (see each function)
 *
 * Error convention (core/blast_engine.c:1237-1243): 0 = success, non-zero aborts; no exceptions,
 * no exit().  bn_last_error() returns a thread-local message.  There is NO CPU fallback: when no
 * CUDA device is usable every compute entry point returns BN_ERR_NO_DEVICE.
 */
#ifndef GBLASTN_B200_H
#define GBLASTN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_OK               0
#define BN_ERR_INVALID      1
#define BN_ERR_MEMORY       2   /* BLASTERR_MEMORY analogue */
#define BN_ERR_NO_DEVICE    3
#define BN_ERR_CUDA         4
#define BN_ERR_UNSUPPORTED  5
#define BN_ERR_OVERFLOW     6

#define BN_LUT_MB        0      /* eMBLookupTable      (inc-core/blast_options.h:164) */
#define BN_LUT_SMALL_NA  1      /* eSmallNaLookupTable */
#define BN_LUT_NA        2      /* eNaLookupTable: small word sizes with a query batch beyond the 15-bit offsets of the small table */

#define BN_DIAG_ARRAY    0      /* eDiagArray (inc-core/blast_parameters.h) */
#define BN_DIAG_HASH     1      /* eDiagHash  */

#define BN_GAP_DP        0      /* eDynProgScoreOnly  -> s_BlastDynProgNtGappedAlignment */
#define BN_GAP_GREEDY    1      /* eGreedyScoreOnly   -> BLAST_GreedyGappedAlignment    */

#define BN_MAX_DBSEQ_LEN        200000000  /* MAX_DBSEQ_LEN, inc-core/blast_gapalign.h:54-55 (G-BLASTN value) */
#define BN_DBSEQ_CHUNK_OVERLAP  100        /* DBSEQ_CHUNK_OVERLAP, inc-core/blast_hits.h:169 */

/* One query context = one strand of one query (BlastContextInfo, inc-core/blast_query_info.h:46-58)
 * plus the per-context cutoffs the reference keeps in BlastUngappedCutoffs /
 * BlastGappedCutoffs (inc-core/blast_parameters.h) and the gapped Karlin block. */
typedef struct BnContext {
    int32_t query_offset;       /* offset in the concatenated query */
    int32_t query_length;
    int32_t query_index;
    int32_t frame;              /* +1 / -1 */
    int32_t is_valid;
    int32_t length_adjustment;
    int64_t eff_searchsp;
    int32_t x_dropoff;          /* ungapped X (raw, positive)              */
    int32_t cutoff_score;       /* ungapped cutoff (gap trigger)           */
    int32_t reduced_cutoff;     /* reduced_nucl_cutoff_score               */
    int32_t gapped_cutoff;      /* hit_params->cutoffs[ctx].cutoff_score   */
    double  gap_lambda;         /* sbp->kbp_gap[ctx]->Lambda               */
    double  gap_logK;           /* sbp->kbp_gap[ctx]->logK                 */
} BnContext;

/* Everything the path needs from one query batch.  All pointers are HOST pointers owned by
 * the caller; bn_query_load copies what it needs to the device(s). */
typedef struct BnQueryBatch {
    /* query BLAST_SequenceBlk: sequence_start (leading sentinel) .. trailing sentinel */
    const uint8_t *query_start;      /* concat_len + 2 bytes, blastna, sentinel 15 */
    int32_t        concat_len;       /* query->length */
    int32_t        num_contexts;
    const BnContext *contexts;
    int32_t        num_queries;

    /* lookup table arrays (layout is the reference's: SURVEY.md A.3) */
    int32_t        lut_type;         /* BN_LUT_*            */
    int32_t        word_length;      /* full word size       */
    int32_t        lut_word_length;
    int32_t        scan_step;
    int64_t        hashsize;         /* MB: 4^lut; SmallNa: backbone_size */
    const int32_t *hashtable;        /* MB; NULL (with next_pos NULL) => built on the device from lookup_segments */
    const int32_t *next_pos;         /* MB, concat_len + 1 */
    const uint32_t *pv_array;        /* MB, optional (NULL => not used; the device builds its own) */
    int32_t        pv_array_bts;
    const int16_t *backbone;         /* SmallNa final_backbone */
    const int16_t *overflow;         /* SmallNa */
    int64_t        overflow_len;
    const int32_t *masked_locations; /* pairs [left,right]; NULL when lut->masked_locations == NULL */
    int32_t        n_masked_locations;

    /* BlastInitialWordParameters / options */
    int32_t        container_type;   /* BN_DIAG_*           */
    int32_t        window_size;      /* two-hit window (0 = one-hit) */
    int32_t        scan_range;
    int32_t        nucl_score_table[256];
    int32_t        matrix[256];      /* 16 x 16, row-major  */

    /* BlastScoringParameters / BlastExtensionParameters / BlastHitSavingOptions */
    int32_t        gap_algo;         /* BN_GAP_*            */
    int32_t        reward, penalty, gap_open, gap_extend;
    int32_t        gap_x_dropoff;    /* raw */
    int32_t        min_diag_separation;
    int32_t        round_down;       /* sbp->round_down (odd-score rounding) */
    int32_t        hsp_num_max;      /* 0 => unlimited; no effect on gapped searches (BlastHspNumMax, core/blast_hits.c:169-191) */
    int32_t        hitlist_size;     /* for the low_score rule */
    double         evalue_cutoff;    /* hit_options->expect_value */
    double         low_score_perc;   /* 0 disables the rule */

    /* lookup_segments of LookupTableWrapInit (core/lookup_wrap.c:49-122): unmasked [left, right]
     * intervals of the concatenated query, ascending.  When an MB batch arrives WITHOUT
     * hashtable/next_pos, bn_query_load runs the fill of s_FillContigMBTable
     * (core/blast_nalookup.c:832-937) on the device from these (same table, bit for bit) and the
     * 4^lut-entry table never crosses PCIe. */
    const int32_t *lookup_segments;
    int32_t        n_lookup_segments;

    /* BlastNaLookupTable (inc-core/blast_nalookup.h:111-160), lut_type BN_LUT_NA: thick_backbone as it lies
     * in memory (hashsize cells of 4 ints: num_used, then entries[3] or overflow_cursor) and overflow. */
    const int32_t *na_backbone;
    const int32_t *na_overflow;
    int64_t        na_overflow_len;

    /* BlastHitSavingOptions of the traceback stage (Blast_HSPTest, core/blast_hits.c:864-871): an HSP whose
     * num_ident * 100 < align_length * percent_identity, or whose align_length < min_hit_length, is dropped
     * (blastn -perc_identity; 0 / 0 = off).  The preliminary gapped stage does not look at them. */
    double         percent_identity;
    int32_t        min_hit_length;
    int32_t        reserved0;
} BnQueryBatch;

/* BlastOffsetPair (inc-core/blast_def.h:141) tagged with its subject. */
typedef struct BnOffsetPair { uint32_t q_off, s_off; } BnOffsetPair;

/* BlastInitHSP + BlastUngappedData (inc-core/blast_extend.h:141-163). */
typedef struct BnInitHit {
    int32_t oid, chunk_off;
    int32_t q_off, s_off;                    /* seed (word start) */
    int32_t q_start, s_start, length, score; /* ungapped_data    */
} BnInitHit;

/* BlastHSP after the preliminary stage (inc-core/blast_hits.h:107). */
typedef struct BnHSP {
    int32_t oid, context;
    int32_t q_off, q_end, s_off, s_end;
    int32_t score;
    int32_t q_gapped_start, s_gapped_start;
    int32_t chunk_off;          /* subject chunk the HSP came from (diagnostic) */
    double  evalue;
} BnHSP;

/* BlastUngappedStats / BlastGappedStats (inc-core/blast_diagnostics.h) + timing. */
typedef struct BnStats {
    int64_t lookup_hits, init_extends, good_init_extends, gap_extensions, good_extensions;
    int64_t subject_bases_scanned;
    double  ms_scan, ms_extend, ms_gapped, ms_host, ms_total;   /* CUDA-event / host timers */
    int64_t kernel_launches;
} BnStats;

typedef struct BnResults {
    BnHSP    *hsps;      int64_t n_hsps;       /* per subject, oid ascending, list order */
    BnInitHit *init;     int64_t n_init;       /* only when BN_TAP_INIT */
    BnHSP    *gapped;    int64_t n_gapped;     /* only when BN_TAP_GAPPED: per-chunk lists before post-processing */
    BnStats   stats;
} BnResults;

#define BN_TAP_INIT    2
#define BN_TAP_GAPPED  4

/* ---- lifecycle --------------------------------------------------------------------------- */
int  bn_init(int n_gpu, const int *device_ids);     /* n_gpu <= 0: all visible devices */
void bn_release(void);
int  bn_device_count(void);
const char *bn_last_error(void);
const char *bn_version(void);

/* ---- database residency: one volume per handle, bound to one device ------------------------
 * packed: ncbi2na bytes; sequence i = bytes [seq_byte_off[i], ...), seq_len[i] bases.
 * >= 16 readable bytes must follow the last sequence. */
int  bn_db_load(int device, const uint8_t *packed, int64_t packed_bytes,
                const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq,
                int *vol_handle);
int  bn_db_free(int vol_handle);
/* Database masks (blastn -db_soft_mask / -db_hard_mask; BLAST_SequenceBlk::seq_ranges + mask_type,
 * inc-core/blast_def.h:242; BlastSeqBlkSetSeqRanges core/blast_util.c:186-223): sequence i carries mask_n[i]
 * masked [begin, end) intervals, ascending and disjoint, flat pairs in mask_iv.  Soft masks keep seeds out of
 * the masked ranges (extensions may run through them); hard masks also split the subject into chunks at the
 * masks (core/blast_engine.c:220-301).  BN_MASK_NONE removes them. */
#define BN_MASK_NONE 0
#define BN_MASK_SOFT 1
#define BN_MASK_HARD 2
int  bn_db_set_masks(int vol_handle, int mask_type, const int32_t *mask_n, const int32_t *mask_iv);

/* BLAST database volume files (version 4 .nin index + .nsq packed sequences, the files
 * `makeblastdb -dbtype nucl` writes and CSeqDBVol reads: objtools/blast/seqdb_reader/seqdbfile.cpp:195-250,
 * seqdbvol.cpp:263-285,1734-1815).  bn_dbfile_index is host-only (no device needed): it fills `info`
 * and, when non-NULL, the per-sequence byte offsets into the .nsq and lengths (info->n_seq entries).
 * bn_db_load_files maps the .nsq and sends it to HBM unchanged; the handle is a volume like any other.
 * bn_dbfile_write stores an in-memory volume in the same format (no deflines, no ambiguity data). */
typedef struct BnDbFileInfo {
    int32_t n_seq, max_len;
    int64_t total_bases, nsq_bytes;
    char    title[256];
} BnDbFileInfo;
int  bn_dbfile_index(const char *nin_path, const char *nsq_path, BnDbFileInfo *info,
                     int64_t *seq_byte_off, int32_t *seq_len);
int  bn_db_load_files(int device, const char *nin_path, const char *nsq_path, int *vol_handle);
/* Ambiguity data of a volume's sequences (host only): what CSeqDBVol::x_GetAmbigSeq overlays on the 2-bit bases when the
 * traceback stage fetches a subject in blastna (objtools/blast/seqdb_reader/seqdbvol.cpp:832-870, 1565-1640).
 * first: n_seq + 1 entries, runs of sequence i are first[i] .. first[i+1]; *runs: malloc'ed flat triples
 * {first base, number of bases, blastna code} in file order (later runs overwrite earlier ones); free with bn_free. */
/* Ambiguity runs of a resident volume (bn_db_load_files installs the ones of its .nsq by itself; this is for volumes
 * loaded from memory): same layout as bn_dbfile_ambiguity returns.  The preliminary stage reads the 2-bit bases as they
 * are, exactly like the reference (api/seqsrc_seqdb.cpp:283-388 hands out ncbi2na there); the traceback stage
 * (bn_gapped_traceback, bn_traceback_hsps, bn_traceback_search) lays the runs over them, as the reference's blastna
 * subject fetch does.  first == NULL removes the runs. */
int  bn_db_set_ambiguity(int vol_handle, const int64_t *first, const int32_t *runs);
int  bn_dbfile_ambiguity(const char *nin_path, const char *nsq_path, int64_t *first, int32_t **runs, int64_t *n_runs);
int  bn_dbfile_write(const char *nin_path, const char *nsq_path, const char *title, const uint8_t *packed,
                     const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq);

/* ---- query batch: replicated to every device in use ------------------------------------------ */
int  bn_query_load(const BnQueryBatch *batch, int *query_handle);
int  bn_query_free(int query_handle);

/* ---- the path ------------------------------------------------------------------------------- */
/* Whole preliminary stage for oids [oid_begin, oid_end) of a resident volume. */
int  bn_prelim_search(int vol_handle, int query_handle, int32_t oid_begin, int32_t oid_end,
                      int taps, BnResults *out);
/* HOST-buffer convenience used by the reference-facing plugin path and bench e2e:
 * H2D of the volume + search + D2H inside one call. */
int  bn_prelim_search_host(int device, const BnQueryBatch *batch,
                           const uint8_t *packed, int64_t packed_bytes,
                           const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq,
                           int taps, BnResults *out);
/* A stream of query batches against one resident volume — blastn's batch loop (app/blast/blastn_app.cpp:574-640)
 * and G-BLASTN's Prepare -> Prelim -> ... thread pipeline (gpu/work_thread.cpp:16-438) — as a two-stage pipeline:
 * while batch k is on the GPU (table upload/fill, scan ... gapped) a host thread finishes batch k-1
 * (containment replay, list post-processing, E-values).  results[k] receives what bn_prelim_search would
 * return for batches[k]; every batch sees the whole volume. */
int  bn_prelim_search_batches(int vol_handle, int32_t n_batches, const BnQueryBatch *const *batches, int taps,
                              BnResults *results);
/* Database volumes sharded over the GPUs: the preliminary stage for n_volumes resident volumes
 * (any devices; one host thread per volume drives its device) and ONE query batch, gathered on the host.  OIDs of the
 * result are those of the concatenated database: volume v's sequences follow those of volumes 0 .. v-1.  The host
 * replay walks the volumes in that order with a single set of per-query hit lists, so hit_params->low_score
 * (core/blast_engine.c:1313-1320) evolves as in one pass of the reference over the whole database, and with
 * prune_hitlists != 0 only the subject lists that the HSP stream still holds at the end of the stage are returned:
 * at most prelim_hitlist_size per query (BlastHSPCollectorParamsNew core/hspfilter_collector.c:328-342,
 * Blast_HitListUpdate core/blast_hits.c:2924-2981) - the rule is applied once, after the gather, never per volume. */
int  bn_prelim_search_volumes(int32_t n_volumes, const int *vol_handles, int query_handle, int taps,
                              int prune_hitlists, BnResults *out);
void bn_results_free(BnResults *r);

/* Stage-level entry points (parity taps; same semantics as the reference callbacks).
 * bn_scan_subject: chunk_len > 0 selects the subject chunk [chunk_off, chunk_off + chunk_len) of the reference's split
 * (BN_ERR_INVALID if the subject has no such chunk); chunk_len == 0 (with chunk_off 0) returns every chunk of the
 * subject in order.  Subject offsets are chunk-relative either way. */
int  bn_scan_subject(int vol_handle, int query_handle, int32_t oid, int32_t chunk_off,
                     int32_t chunk_len, BnOffsetPair **pairs, int64_t *n_pairs);
int  bn_word_finder(int vol_handle, int query_handle, int32_t oid_begin, int32_t oid_end,
                    BnInitHit **init, int64_t *n_init);
/* BlastGetGappedScoreType (inc-core/blast_engine.h:212-224), i.e. BLAST_GetGappedScore
 * (core/blast_gapalign.c:3233-3559) for ONE subject chunk: `init` is the BlastInitHitList the word
 * finder produced for the chunk that starts at base chunk_off of sequence oid (any order; it is
 * sorted like Blast_InitHitListSortByScore, ties keep the given order), low_score is
 * hit_params->low_score (per query, may be NULL).  Returns the HSP list as it stands when the
 * reference function returns (before purge / sort / E-values), subject offsets chunk-relative.  Every init hit is
 * checked (seed and ungapped segment inside one query context and inside the chunk): BN_ERR_INVALID otherwise. */
int  bn_get_gapped_score(int vol_handle, int query_handle, int32_t oid, int32_t chunk_off,
                         const BnInitHit *init, int64_t n_init, const int32_t *low_score,
                         BnHSP **hsps, int64_t *n_hsps);
void bn_free(void *p);

/* ---- traceback stage, first row ----------------------------------------------
 * BLAST_GappedAlignmentWithTraceback (core/blast_gapalign.c:3994-4155; Blast_SemiGappedAlign with traceback ->
 * ALIGN_EX :350-709) for a batch of start points, as Blast_TracebackFromHSPList makes the call for blastn with
 * eDynProgTbck (core/blast_traceback.c:565-571): query = the context's strand of the query block, subject =
 * sequence `oid` + s_shift with s_length bases (the window AdjustSubjectRange leaves, core/blast_traceback.c:513-520),
 * start point (q_start, s_start) relative to those, X-drop = gap_x_dropoff_final (gap_align->gap_x_dropoff is set to
 * it at core/blast_traceback.c:1403), costs and matrix from the query batch.  The subject is the resident packed
 * volume with its ambiguity runs, if it has any, laid over the 2-bit bases (the blastna subject the reference fetches).
 * Results mirror BlastGapAlignStruct after the call: score, query_start/stop, subject_start/stop (relative to the
 * window) and gap_align->edit_script as ops[esp_off .. esp_off + esp_n) with op_type = EGapAlignOpType
 * (0 eGapAlignDel, 3 eGapAlignSub, 6 eGapAlignIns; inc-core/gapinfo.h:44-54).  When the batch's gap_algo is
 * BN_GAP_GREEDY the routine is BLAST_GreedyGappedAlignment with do_traceback (core/blast_gapalign.c:2620-2751:
 * BLAST_GreedyAlign, or BLAST_AffineGreedyAlign's own body when gap costs are given, then s_ReduceGaps), as the
 * reference calls it with eGreedyTbck (core/blast_traceback.c:559-563).  Free both arrays with bn_free. */
typedef struct BnTracebackItem {
    int32_t oid, context;
    int32_t s_shift, s_length;
    int32_t q_start, s_start;
} BnTracebackItem;
typedef struct BnTracebackResult {
    int32_t score, query_start, query_stop, subject_start, subject_stop;
    int32_t esp_n;
    int64_t esp_off;
    int32_t status, pad;
} BnTracebackResult;
typedef struct BnEditOp { int32_t op_type, num; } BnEditOp;
int  bn_gapped_traceback(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                         const BnTracebackItem *items, int64_t n_items,
                         BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops);
/* The per-HSP body of Blast_TracebackFromHSPList's loop (core/blast_traceback.c:490-571) for a list of preliminary
 * HSPs (what bn_prelim_search returns: absolute subject coordinates, gapped start points): on the device, the start
 * point — BLAST_CheckStartForGappedAlignment (:97-153), else BlastGetOffsetsForGappedAlignment
 * (core/blast_gapalign.c:3059-3131); a good stored start is moved into the longest run of identities by
 * BlastGetStartForGappedAlignmentNucl (:3134-3182) — and AdjustSubjectRange (:3608-3636), then the alignment with
 * traceback as in bn_gapped_traceback.  items[i] is the call the reference would make for hsps[i] (oid = -1 when no
 * start point exists and the reference drops the HSP; results[i].status = -1 then).  Every HSP is extended: the
 * caller replays the containment test of :449 over the results in score order.  Free the three arrays with bn_free. */
int  bn_traceback_hsps(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                       const BnHSP *hsps, int64_t n_hsps, BnTracebackItem **items,
                       BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops);
/* The traceback stage of a blastn / megablast database search: BLAST_ComputeTraceback (core/blast_traceback.c:1375-1640)
 * -> Blast_TracebackFromHSPList (:336-790) for every (query, subject) list -> s_HSPListPostTracebackUpdate (:278-335)
 * -> Blast_HSPResultsSortByEvalue / s_BlastPruneExtraHits.  `hsps` = the preliminary lists as bn_prelim_search returns
 * them (per subject, sorted by score).  All HSPs are aligned on the device at once (start point, alignment with
 * traceback), the reference's sequential decisions — containment of an HSP in a better one's alignment, the
 * common-endpoint pass with its edit-script trimming (Blast_HSPListPurgeHSPsWithCommonEndpoints with purge = FALSE), the
 * second containment pass, odd-score rounding, E-values (BLAST_KarlinStoE_simple), reap, bit scores — are replayed on
 * the host over those results, and Blast_HSPReevaluateWithAmbiguitiesGapped + the identity count run on the device
 * in between.  Output: per query (ascending), the subject lists in s_EvalueCompareHSPLists order, at most hitlist_size
 * of them, HSPs in list order; edit scripts in ops.  hit_options->percent_identity / min_hit_length (Blast_HSPTest,
 * core/blast_hits.c:864-871) come with the batch: DP tracebacks are tested right after the alignment, before the HSP
 * enters the containment tree (core/blast_traceback.c:658-669), greedy ones and trimmed HSPs after the re-evaluation
 * (:727-735).  Limits: the identity count reads the query block the batch carries (`sequence`), which equals
 * `sequence_nomask` unless the query was hard-masked; the per-query pruning of the preliminary hit lists
 * (prelim_hitlist_size) is the caller's. */
typedef struct BnTracebackHSP {
    int32_t query_index, oid, context;
    int32_t q_off, q_end, s_off, s_end;
    int32_t score, num_ident, esp_n;
    int64_t esp_off;
    double  evalue, bit_score;
} BnTracebackHSP;
int  bn_traceback_search(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                         const BnHSP *hsps, int64_t n_hsps,
                         BnTracebackHSP **out, int64_t *n_out, BnEditOp **ops, int64_t *n_ops);

/* A stream of searches on one device as a software pipeline — G-BLASTN's Prepare -> Prelim -> Traceback -> Print worker
 * threads (gpu/work_thread.cpp:16-438, app/blast/blastn_app.cpp:886-989) without the Print stage.  A job is one
 * (volume, query batch) pair; either side may be resident (a handle) or come from the caller's HOST buffers for this
 * job only (handle -1: the upload is part of the job).  Per job:
 *   prepare    on the caller's thread: uploads + device-side table fill of a host-side batch, upload of a host-side
 *              volume (copy stream), chunk table, and ALL kernels of the preliminary stage queued on the lane's stream;
 *              job k+1 is prepared before the call waits for job k, so the device never idles between jobs and the
 *              upload of job k+1's volume overlaps the kernels of job k
 *   complete   wait for the job's kernels, second-tier gapped extensions if any were needed
 *   host       on a worker thread: containment replay, list post-processing, E-values (results[k])
 *   traceback  (tb != NULL) on a second worker with a lane of its own: bn_traceback_search on results[k].hsps
 * results[k] (and tb[k]) are exactly what bn_prelim_search (bn_traceback_search) returns for the pair; free them with
 * bn_results_free / bn_free.  All host arrays the jobs point to must stay valid until the call returns. */
typedef struct BnJob {
    int vol_handle;                 /* resident volume, or -1 */
    int query_handle;               /* resident batch, or -1 */
    const BnQueryBatch *batch;      /* query_handle == -1 */
    const uint8_t *packed;          /* vol_handle == -1: as for bn_db_load */
    int64_t        packed_bytes;
    const int64_t *seq_byte_off;
    const int32_t *seq_len;
    int32_t        n_seq;
    int32_t        gap_x_dropoff_final;   /* traceback stage (tb != NULL) */
} BnJob;
typedef struct BnTracebackOut {
    BnTracebackHSP *hsps; int64_t n_hsps;
    BnEditOp *ops;        int64_t n_ops;
} BnTracebackOut;
int  bn_prelim_search_jobs(int device, int32_t n_jobs, const BnJob *jobs, int taps, BnResults *results,
                           BnTracebackOut *tb);

/* Host-only self-test: the containment replay (BLAST_GetGappedScore's interval-tree filter, core/blast_itree.c)
 * runs with one tree per query strand; this compares it with the reference's one-tree-per-subject layout on
 * n_cases seeded random init-HSP sets and reports how many differ (0 expected).  Needs no device. */
int  bn_selftest_replay(uint64_t seed, int32_t n_cases, int64_t *n_mismatch);

/* Device self-test of the path's own sort and prefix sum (csrc/radix_sort.cu: the seed hits of the general word-finder
 * path, the (word, position) pairs of the device-side table fill): n seeded random pairs with key_bits-bit keys are
 * sorted on `device` and compared with a stable host sort, the inclusive prefix sum of n counts with a host loop;
 * *n_mismatch = positions that differ (0 expected). */
int  bn_selftest_sort(int device, int64_t n, int key_bits, uint64_t seed, int64_t *n_mismatch);

/* Parity tap for the device-side table fill: reconstructs hashtable[hashsize] and
 * next_pos[concat_len + 1] of an MB batch from the arrays resident on `device`. */
int  bn_query_download_lookup(int query_handle, int device, int32_t *hashtable, int32_t *next_pos);

/* Kernel-only timing hook for bench.py / ncu: runs the scan(+mini-extension) kernel `iters`
 * times over the resident volume and returns the average device time per launch. */
int  bn_bench_scan(int vol_handle, int query_handle, int iters, double *ms_per_launch,
                   int64_t *bases_per_launch, int64_t *hits);

/* ---- host-side set-up mirror (what the reference computes before the path) ------------------- */
typedef struct BnSetupOptions {
    int32_t task;               /* 0 megablast, 1 blastn */
    int32_t word_size;          /* 0 => 28 / 11 */
    int32_t reward, penalty;    /* 0 => 1/-2 or 2/-3 */
    int32_t gap_open, gap_extend; /* -1 => 0/0 or 5/2 */
    int32_t greedy;             /* -1 => task default */
    int32_t window_size, scan_range;
    int32_t min_diag_separation; /* -1 => 6 / 50 */
    int32_t hitlist_size;       /* 0 => 500 */
    int32_t mask_at_hash;
    double  xdrop_ungap, xdrop_gap, xdrop_gap_final, evalue, low_score_perc;
    int64_t db_length;          /* total bases of the database (all volumes) */
    int32_t db_num_seqs;
    int32_t avg_subject_length; /* BlastSeqSrcGetAvgSeqLen */
    int32_t device_lookup;      /* 1: leave the megablast table fill to bn_query_load (device); the batch
                                   then carries lookup_segments and NULL hashtable/next_pos */
    int32_t hsp_num_max;        /* hit_options->hsp_num_max; carried into the batch, and — like the reference, whose
                                   BlastHspNumMax returns INT4_MAX for gapped searches (core/blast_hits.c:169-191) —
                                   without effect on this (always gapped) path */
    double  percent_identity;   /* hit_options->percent_identity (blastn -perc_identity), 0 = off */
    int32_t min_hit_length;     /* hit_options->min_hit_length, 0 = off */
} BnSetupOptions;

typedef struct BnSetup BnSetup;   /* opaque; owns the arrays a BnQueryBatch points to */

/* queries: blastna bytes concatenated; masks: optional plus-strand inclusive intervals. */
int  bn_setup_create(const BnSetupOptions *opt, int32_t n_queries, const uint8_t *qseq,
                     const int32_t *qlens, const int32_t *qmask_n, const int32_t *qmask_iv,
                     BnSetup **out);
const BnQueryBatch *bn_setup_batch(const BnSetup *s);
/* Karlin-Altschul blocks computed by the set-up: 4 doubles (Lambda, K, logK, H) per context. */
const double *bn_setup_kbp_std(const BnSetup *s);
const double *bn_setup_kbp_gap(const BnSetup *s);
int32_t bn_setup_gap_x_dropoff_final(const BnSetup *s);
int32_t bn_setup_longest_chain(const BnSetup *s);
void bn_setup_free(BnSetup *s);

/* ---- query low-complexity masking ------------------------------------------------
 * Symmetric DUST as blastn applies it to its queries (`-dust yes`, task default level 20 / window 64 / linker 1):
 * CSymDustMasker (c++/src/algo/dustmask/symdust.cpp:213-319) called by Blast_FindDustFilterLoc
 * (c++/src/algo/blast/api/dust_filter.cpp:65-151).  seq: blastna bytes of ONE query (codes >= 4 read as A, like every
 * non-ACGT IUPAC letter does in the reference); out: n_intervals inclusive [from, to] pairs, ascending, already merged
 * by `linker`; free with bn_free.  Host only (no device needed): the masks are the qmask_iv input of bn_setup_create,
 * i.e. they only keep words out of the lookup table (mask-at-hash) and switch on s_TypeOfWord's re-probing. */
int  bn_dust_mask(const uint8_t *seq, int32_t len, int32_t level, int32_t window, int32_t linker,
                  int32_t **intervals, int32_t *n_intervals);
/* The same for a whole query batch ON THE DEVICE (csrc/dust_kernel.cu: one thread per query runs the window scan with
 * its state in local memory; identical intervals).  seqs: the queries' blastna bytes back to back, lens[n_queries].
 * Output in the form bn_setup_create takes: mask_n[i] intervals for query i, flat inclusive [from, to] pairs in
 * mask_iv (query coordinates); free both with bn_free.  A query whose window ever holds more perfect intervals than the
 * kernel's list (2048) is redone by bn_dust_mask. */
int  bn_dust_mask_batch(int device, const uint8_t *seqs, const int32_t *lens, int32_t n_queries,
                        int32_t level, int32_t window, int32_t linker, int32_t **mask_n, int32_t **mask_iv,
                        int64_t *n_intervals);

#ifdef __cplusplus
}
#endif
#endif /* GBLASTN_B200_H */
