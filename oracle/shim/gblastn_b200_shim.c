/* gblastn_b200_shim.c — the reference-side binding of libgblastn_b200.so, compiled against the reference's
 * own headers.  It is what a maintainer of the reference adds to route the two seams of the preliminary
 * search engine into the B200 engine and leave everything else — the subject loop, chunking, list
 * post-processing, E-values, the HSP stream — to the reference's unmodified code:
 *
 *   aux_struct->WordFinder     = BlastNaWordFinder     (core/blast_engine.c:926)  ->  bnshim_word_finder
 *   aux_struct->GetGappedScore = BLAST_GetGappedScore  (core/blast_engine.c:940)  ->  bnshim_get_gapped_score
 *   BLAST_GapAlignSetUp (core/blast_setup.c; called by Blast_RunPreliminarySearchWithInterrupt right before
 *       BLAST_PreliminarySearchEngine, core/blast_engine.c:1407-1420)             ->  bnshim_prelim_begin
 *   Blast_RunPreliminarySearchWithInterrupt (the call G-BLASTN itself replaces, api/prelim_search_runner.hpp:96,
 *       inc-gpu/gpu_blastn.h:31-48)                            ->  bnshim_RunPreliminarySearchWithInterrupt
 *
 * Link recipe for a `blastn` built from the reference tree:
 *   -Wl,--wrap=BlastNaWordFinder -Wl,--wrap=BLAST_GetGappedScore -Wl,--wrap=BLAST_GapAlignSetUp
 *   gblastn_b200_shim.o (compiled with -DBNSHIM_DEFINE_WRAPS) -lgblastn_b200
 * (`ld --wrap` only redirects references between object files, which is why the seams are the three exported
 * functions above and not BLAST_PreliminarySearchEngine, whose caller lives in the same object.)
 * With BNSHIM_DEFINE_WRAPS this file defines the three __wrap_ symbols itself.  In this repository the shim is
 * linked into oracle/_ref/libblastshim.so next to the test driver (oracle/ref_driver.c), whose own --wrap taps
 * call the bnshim_* functions when RefConfig.seam == 1, so the hybrid (reference engine + B200 seams) is
 * tapped exactly like the pure reference and compared with it (tests/test_shim_hybrid.py).
 *
 * TEST / INTEGRATION INFRASTRUCTURE: lives under oracle/, compiled only where /root/reference exists, never
 * loaded by the product package.
 *
 * Ownership: init_hitlist is filled through the reference's own BLAST_SaveInitialHit with
 * libc-malloc'ed BlastUngappedData (BlastInitHitListReset frees each with sfree, core/blast_extend.c:229-236);
 * HSP lists are built with Blast_HSPInit / Blast_HSPListSaveHSP and freed by the stream.
 * Threading: all state is thread-local; the engine itself is re-entrant (one lane per concurrent caller).
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include <algo/blast/core/blast_def.h>
#include <algo/blast/core/blast_options.h>
#include <algo/blast/core/blast_engine.h>
#include <algo/blast/core/blast_util.h>
#include <algo/blast/core/blast_hits.h>
#include <algo/blast/core/blast_extend.h>
#include <algo/blast/core/blast_nalookup.h>
#include <algo/blast/core/na_ungapped.h>
#include <algo/blast/core/blast_gapalign.h>
#include <algo/blast/core/blast_parameters.h>
#include <algo/blast/core/lookup_wrap.h>

#include "../../include/gblastn_b200.h"
#include "gblastn_b200_shim.h"

typedef struct ShimState {
    /* the volume the seqsrc hands sequences out of (host pointers) and its resident copy */
    const uint8_t *host_base;
    const int64_t *seq_byte_off;
    int32_t n_seq;
    int vol_handle;
    int owns_volume;
    /* the search in progress (BLAST_PreliminarySearchEngine's arguments) */
    const BlastScoringParameters *score_params;
    const BlastExtensionParameters *ext_params;
    const BlastHitSavingParameters *hit_params;
    const BlastGapAlignStruct *gap_align;
    int active;
    int query_handle;                    /* -1 until the first word-finder call of the search */
    const LookupTableWrap *loaded_for;
    /* init hits of the subject being searched (all chunks), from one bn_word_finder call */
    int32_t cached_oid;
    BnInitHit *init;
    int64_t n_init;
    char err[256];
} ShimState;

static __thread ShimState g_shim = { NULL, NULL, 0, -1, 0, NULL, NULL, NULL, NULL, 0, -1, NULL, -1, NULL, 0, {0} };

const char *bnshim_last_error(void) { return g_shim.err; }

static int shim_fail(const char *what)
{
    snprintf(g_shim.err, sizeof g_shim.err, "%s: %s", what, bn_last_error());
    return -1;
}

int bnshim_attach_volume(const uint8_t *packed, int64_t packed_bytes, const int64_t *seq_byte_off,
                         const int32_t *seq_len, int32_t n_seq, int device)
{
    ShimState *S = &g_shim;
    bnshim_detach_volume();
    if (bn_db_load(device, packed, packed_bytes, seq_byte_off, seq_len, n_seq, &S->vol_handle) != BN_OK)
        return shim_fail("bn_db_load");
    S->host_base = packed; S->seq_byte_off = seq_byte_off; S->n_seq = n_seq; S->owns_volume = 1;
    return 0;
}

int bnshim_attach_resident_volume(int vol_handle, const uint8_t *host_base, const int64_t *seq_byte_off, int32_t n_seq)
{
    ShimState *S = &g_shim;
    bnshim_detach_volume();
    S->vol_handle = vol_handle; S->host_base = host_base; S->seq_byte_off = seq_byte_off; S->n_seq = n_seq;
    S->owns_volume = 0;
    return 0;
}

void bnshim_detach_volume(void)
{
    ShimState *S = &g_shim;
    if (S->owns_volume && S->vol_handle >= 0) bn_db_free(S->vol_handle);
    S->vol_handle = -1; S->owns_volume = 0; S->host_base = NULL; S->seq_byte_off = NULL; S->n_seq = 0;
}

void bnshim_prelim_begin(const BlastScoringParameters *score_params, const BlastExtensionParameters *ext_params,
                         const BlastHitSavingParameters *hit_params, const BlastGapAlignStruct *gap_align)
{
    ShimState *S = &g_shim;
    if (S->active) bnshim_prelim_end();          /* a search that never reached its end (error paths) */
    S->err[0] = 0;
    S->score_params = score_params; S->ext_params = ext_params; S->hit_params = hit_params; S->gap_align = gap_align;
    S->active = 1; S->query_handle = -1; S->loaded_for = NULL; S->cached_oid = -1; S->init = NULL; S->n_init = 0;
}

void bnshim_prelim_end(void)
{
    ShimState *S = &g_shim;
    if (S->query_handle >= 0) bn_query_free(S->query_handle);
    bn_free(S->init);
    S->query_handle = -1; S->loaded_for = NULL; S->init = NULL; S->n_init = 0; S->cached_oid = -1; S->active = 0;
}

/* BlastSeqLoc list -> flat [left, right] pairs */
static int32_t *flatten_locs(const BlastSeqLoc *loc, int32_t *n_out)
{
    const BlastSeqLoc *l;
    int32_t n = 0, *out;
    for (l = loc; l; l = l->next) n++;
    out = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n ? n : 1));
    for (n = 0, l = loc; l; l = l->next, n++) { out[2 * n] = l->ssr->left; out[2 * n + 1] = l->ssr->right; }
    *n_out = n;
    return out;
}

/* GpuLookUpSetUp's counterpart (gpu/gpu_blastn_na_ungapped_v3.cpp:595-696): everything the engine needs from the
 * reference's structures, by pointer; bn_query_load copies it to the device(s). */
static int load_query(BLAST_SequenceBlk *query, BlastQueryInfo *query_info, LookupTableWrap *lookup_wrap,
                      Int4 **matrix, const BlastInitialWordParameters *word_params)
{
    ShimState *S = &g_shim;
    const BlastScoreBlk *sbp = S->gap_align->sbp;
    const BlastHitSavingOptions *hopt = S->hit_params->options;
    const int n_ctx = query_info->last_context + 1;
    BnQueryBatch b;
    BnContext *ctx;
    int32_t *masked = NULL;
    int c, i, j, rc;
    const BlastSeqLoc *masked_locations = NULL;

    memset(&b, 0, sizeof b);
    ctx = (BnContext *)calloc((size_t)n_ctx, sizeof(BnContext));
    for (c = 0; c < n_ctx; c++) {
        const BlastContextInfo *ci = &query_info->contexts[c];
        const Blast_KarlinBlk *kbp = sbp->kbp_gap ? sbp->kbp_gap[c] : NULL;
        ctx[c].query_offset = ci->query_offset; ctx[c].query_length = ci->query_length;
        ctx[c].query_index = ci->query_index; ctx[c].frame = ci->frame; ctx[c].is_valid = ci->is_valid;
        ctx[c].length_adjustment = ci->length_adjustment; ctx[c].eff_searchsp = ci->eff_searchsp;
        ctx[c].x_dropoff = word_params->cutoffs[c].x_dropoff;
        ctx[c].cutoff_score = word_params->cutoffs[c].cutoff_score;
        ctx[c].reduced_cutoff = word_params->cutoffs[c].reduced_nucl_cutoff_score;
        ctx[c].gapped_cutoff = S->hit_params->cutoffs[c].cutoff_score;
        ctx[c].gap_lambda = kbp ? kbp->Lambda : -1.0;
        ctx[c].gap_logK = kbp ? kbp->logK : -1.0;
    }
    b.query_start = query->sequence_start; b.concat_len = query->length;
    b.num_contexts = n_ctx; b.contexts = ctx; b.num_queries = query_info->num_queries;
    switch (lookup_wrap->lut_type) {
    case eMBLookupTable: {
        const BlastMBLookupTable *l = (const BlastMBLookupTable *)lookup_wrap->lut;
        b.lut_type = BN_LUT_MB; b.word_length = l->word_length; b.lut_word_length = l->lut_word_length;
        b.scan_step = l->scan_step; b.hashsize = l->hashsize; b.hashtable = l->hashtable; b.next_pos = l->next_pos;
        b.pv_array = l->pv_array; b.pv_array_bts = l->pv_array_bts;
        masked_locations = l->masked_locations;
        break;
    }
    case eSmallNaLookupTable: {
        const BlastSmallNaLookupTable *l = (const BlastSmallNaLookupTable *)lookup_wrap->lut;
        b.lut_type = BN_LUT_SMALL_NA; b.word_length = l->word_length; b.lut_word_length = l->lut_word_length;
        b.scan_step = l->scan_step; b.hashsize = l->backbone_size; b.backbone = l->final_backbone;
        b.overflow = l->overflow; b.overflow_len = l->overflow_size;
        masked_locations = l->masked_locations;
        break;
    }
    case eNaLookupTable: {
        const BlastNaLookupTable *l = (const BlastNaLookupTable *)lookup_wrap->lut;
        b.lut_type = BN_LUT_NA; b.word_length = l->word_length; b.lut_word_length = l->lut_word_length;
        b.scan_step = l->scan_step; b.hashsize = l->backbone_size;
        b.na_backbone = (const int32_t *)l->thick_backbone; b.na_overflow = l->overflow;
        b.na_overflow_len = l->overflow_size;
        masked_locations = l->masked_locations;
        break;
    }
    default:
        free(ctx);
        snprintf(S->err, sizeof S->err, "lookup table type %d is not a blastn table", (int)lookup_wrap->lut_type);
        return -1;
    }
    if (masked_locations) { masked = flatten_locs(masked_locations, &b.n_masked_locations); b.masked_locations = masked; }
    b.container_type = word_params->container_type == eDiagHash ? BN_DIAG_HASH : BN_DIAG_ARRAY;
    b.window_size = word_params->options->window_size; b.scan_range = word_params->options->scan_range;
    memcpy(b.nucl_score_table, word_params->nucl_score_table, sizeof b.nucl_score_table);
    for (i = 0; i < 16; i++) for (j = 0; j < 16; j++) b.matrix[16 * i + j] = matrix[i][j];
    b.gap_algo = S->ext_params->options->ePrelimGapExt == eGreedyScoreOnly ? BN_GAP_GREEDY : BN_GAP_DP;
    b.reward = S->score_params->reward; b.penalty = S->score_params->penalty;
    b.gap_open = S->score_params->gap_open; b.gap_extend = S->score_params->gap_extend;
    b.gap_x_dropoff = S->ext_params->gap_x_dropoff;
    b.min_diag_separation = hopt->min_diag_separation; b.round_down = sbp->round_down ? 1 : 0;
    b.hsp_num_max = hopt->hsp_num_max; b.hitlist_size = hopt->hitlist_size;
    b.evalue_cutoff = hopt->expect_value; b.low_score_perc = hopt->low_score_perc;
    b.percent_identity = hopt->percent_identity; b.min_hit_length = hopt->min_hit_length;
    rc = bn_query_load(&b, &S->query_handle);
    free(ctx); free(masked);
    if (rc != BN_OK) { S->query_handle = -1; return shim_fail("bn_query_load"); }
    S->loaded_for = lookup_wrap;
    return 0;
}

static int32_t chunk_offset_of(const BLAST_SequenceBlk *subject)
{
    const ShimState *S = &g_shim;
    /* the engine advances subject->sequence to the chunk's first byte (s_GetNextSubjectChunk,
     * core/blast_engine.c:234-236); the seqsrc handed out a pointer into the attached volume */
    return (int32_t)((subject->sequence - (S->host_base + S->seq_byte_off[subject->oid])) * 4);
}

/* BlastWordFinderType (inc-core/blast_engine.h:227-238) */
Int2 bnshim_word_finder(BLAST_SequenceBlk *subject, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                        LookupTableWrap *lookup_wrap, Int4 **matrix, const BlastInitialWordParameters *word_params,
                        Blast_ExtendWord *ewp, BlastOffsetPair *offset_pairs, Int4 max_hits,
                        BlastInitHitList *init_hitlist, BlastUngappedStats *ungapped_stats)
{
    ShimState *S = &g_shim;
    int32_t chunk_off;
    int64_t i;
    (void)ewp; (void)offset_pairs; (void)max_hits; (void)ungapped_stats;
    if (!S->active || S->vol_handle < 0 || subject->oid < 0 || subject->oid >= S->n_seq) {
        snprintf(S->err, sizeof S->err, "bnshim_word_finder: no search in progress or no volume attached");
        return -1;
    }
    if (S->query_handle < 0 || S->loaded_for != lookup_wrap)
        if (load_query(query, query_info, lookup_wrap, matrix, word_params)) return -1;
    chunk_off = chunk_offset_of(subject);
    if (S->cached_oid != subject->oid) {
        /* one device pass per subject: the init hits of all its chunks */
        bn_free(S->init); S->init = NULL; S->n_init = 0;
        if (bn_word_finder(S->vol_handle, S->query_handle, subject->oid, subject->oid + 1, &S->init, &S->n_init) != BN_OK)
            return (Int2)shim_fail("bn_word_finder");
        S->cached_oid = subject->oid;
    }
    for (i = 0; i < S->n_init; i++) {
        const BnInitHit *h = &S->init[i];
        BlastUngappedData *u;
        if (h->chunk_off != chunk_off) continue;
        u = (BlastUngappedData *)malloc(sizeof(BlastUngappedData));
        u->q_start = h->q_start; u->s_start = h->s_start; u->length = h->length; u->score = h->score;
        BLAST_SaveInitialHit(init_hitlist, h->q_off, h->s_off, u);
    }
    /* BlastNaWordFinder leaves the list sorted (core/na_ungapped.c:1650-1655); BLAST_GetGappedScore asserts it */
    Blast_InitHitListSortByScore(init_hitlist);
    return 0;
}

/* BlastGetGappedScoreType (inc-core/blast_engine.h:212-224) */
Int2 bnshim_get_gapped_score(EBlastProgramType program_number, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                             BLAST_SequenceBlk *subject, BlastGapAlignStruct *gap_align,
                             const BlastScoringParameters *score_params, const BlastExtensionParameters *ext_params,
                             const BlastHitSavingParameters *hit_params, BlastInitHitList *init_hitlist,
                             BlastHSPList **hsp_list_ptr, BlastGappedStats *gapped_stats, Boolean *fence_hit)
{
    ShimState *S = &g_shim;
    const int32_t chunk_off = chunk_offset_of(subject);
    BnInitHit *in;
    BnHSP *hsps = NULL;
    int64_t n_hsps = 0, i;
    BlastHSPList *list;
    (void)program_number; (void)query; (void)gap_align; (void)score_params; (void)ext_params; (void)fence_hit;
    if (!S->active || S->query_handle < 0) {
        snprintf(S->err, sizeof S->err, "bnshim_get_gapped_score: no query batch loaded");
        return -1;
    }
    in = (BnInitHit *)malloc(sizeof(BnInitHit) * (size_t)(init_hitlist->total ? init_hitlist->total : 1));
    for (i = 0; i < init_hitlist->total; i++) {
        const BlastInitHSP *h = &init_hitlist->init_hsp_array[i];
        in[i].oid = subject->oid; in[i].chunk_off = chunk_off;
        in[i].q_off = (int32_t)h->offsets.qs_offsets.q_off; in[i].s_off = (int32_t)h->offsets.qs_offsets.s_off;
        in[i].q_start = h->ungapped_data->q_start; in[i].s_start = h->ungapped_data->s_start;
        in[i].length = h->ungapped_data->length; in[i].score = h->ungapped_data->score;
    }
    if (bn_get_gapped_score(S->vol_handle, S->query_handle, subject->oid, chunk_off, in, init_hitlist->total,
                            hit_params->low_score, &hsps, &n_hsps) != BN_OK) {
        free(in);
        return (Int2)shim_fail("bn_get_gapped_score");
    }
    free(in);
    list = *hsp_list_ptr;
    if (!list) *hsp_list_ptr = list = Blast_HSPListNew(BlastHspNumMax(TRUE, hit_params->options));
    for (i = 0; i < n_hsps; i++) {
        const BnHSP *h = &hsps[i];
        BlastHSP *hsp = NULL;
        Blast_HSPInit(h->q_off, h->q_end, h->s_off, h->s_end, h->q_gapped_start, h->s_gapped_start, h->context,
                      query_info->contexts[h->context].frame, subject->frame, h->score, NULL, &hsp);
        Blast_HSPListSaveHSP(list, hsp);
    }
    (void)gapped_stats;          /* diagnostics of the replaced stage are the engine's (BnStats), not re-derived here */
    bn_free(hsps);
    return 0;
}

#ifdef BNSHIM_DEFINE_WRAPS
/* the symbols `ld --wrap` redirects the engine's references to */
Int2 __real_BLAST_GapAlignSetUp(EBlastProgramType, const BlastSeqSrc *, const BlastScoringOptions *,
                                const BlastEffectiveLengthsOptions *, const BlastExtensionOptions *,
                                const BlastHitSavingOptions *, BlastQueryInfo *, BlastScoreBlk *,
                                BlastScoringParameters **, BlastExtensionParameters **, BlastHitSavingParameters **,
                                BlastEffectiveLengthsParameters **, BlastGapAlignStruct **);
static __thread int g_in_prelim = 0;
Int2 __wrap_BLAST_GapAlignSetUp(EBlastProgramType program_number, const BlastSeqSrc *seq_src,
                                const BlastScoringOptions *scoring_options,
                                const BlastEffectiveLengthsOptions *eff_len_options,
                                const BlastExtensionOptions *ext_options, const BlastHitSavingOptions *hit_options,
                                BlastQueryInfo *query_info, BlastScoreBlk *sbp, BlastScoringParameters **score_params,
                                BlastExtensionParameters **ext_params, BlastHitSavingParameters **hit_params,
                                BlastEffectiveLengthsParameters **eff_len_params, BlastGapAlignStruct **gap_align)
{
    const Int2 st = __real_BLAST_GapAlignSetUp(program_number, seq_src, scoring_options, eff_len_options, ext_options,
                                               hit_options, query_info, sbp, score_params, ext_params, hit_params,
                                               eff_len_params, gap_align);
    if (g_in_prelim && st == 0) bnshim_prelim_begin(*score_params, *ext_params, *hit_params, *gap_align);
    return st;
}
/* drop-in for the call at api/prelim_search_runner.hpp:96 (where G-BLASTN calls Blast_gpu_RunPreliminarySearchWithInterrupt) */
Int4 bnshim_RunPreliminarySearchWithInterrupt(EBlastProgramType program, BLAST_SequenceBlk *query,
                                              BlastQueryInfo *query_info, const BlastSeqSrc *seq_src,
                                              const BlastScoringOptions *score_options, BlastScoreBlk *sbp,
                                              LookupTableWrap *lookup_wrap, const BlastInitialWordOptions *word_options,
                                              const BlastExtensionOptions *ext_options,
                                              const BlastHitSavingOptions *hit_options,
                                              const BlastEffectiveLengthsOptions *eff_len_options,
                                              const PSIBlastOptions *psi_options, const BlastDatabaseOptions *db_options,
                                              BlastHSPStream *hsp_stream, BlastDiagnostics *diagnostics,
                                              TInterruptFnPtr interrupt_search, SBlastProgress *progress_info)
{
    Int4 st;
    g_in_prelim = 1;
    st = Blast_RunPreliminarySearchWithInterrupt(program, query, query_info, seq_src, score_options, sbp, lookup_wrap,
                                                 word_options, ext_options, hit_options, eff_len_options, psi_options,
                                                 db_options, hsp_stream, diagnostics, interrupt_search, progress_info);
    g_in_prelim = 0;
    bnshim_prelim_end();
    if (st == 0 && g_shim.err[0]) st = -1;      /* the engine ignores the word finder's status (core/blast_engine.c:481) */
    return st;
}
Int2 __wrap_BlastNaWordFinder(BLAST_SequenceBlk *subject, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                              LookupTableWrap *lookup_wrap, Int4 **matrix, const BlastInitialWordParameters *word_params,
                              Blast_ExtendWord *ewp, BlastOffsetPair *offset_pairs, Int4 max_hits,
                              BlastInitHitList *init_hitlist, BlastUngappedStats *ungapped_stats)
{
    return bnshim_word_finder(subject, query, query_info, lookup_wrap, matrix, word_params, ewp, offset_pairs, max_hits,
                              init_hitlist, ungapped_stats);
}
Int2 __wrap_BLAST_GetGappedScore(EBlastProgramType program_number, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                                 BLAST_SequenceBlk *subject, BlastGapAlignStruct *gap_align,
                                 const BlastScoringParameters *score_params, const BlastExtensionParameters *ext_params,
                                 const BlastHitSavingParameters *hit_params, BlastInitHitList *init_hitlist,
                                 BlastHSPList **hsp_list_ptr, BlastGappedStats *gapped_stats, Boolean *fence_hit)
{
    return bnshim_get_gapped_score(program_number, query, query_info, subject, gap_align, score_params, ext_params,
                                   hit_params, init_hitlist, hsp_list_ptr, gapped_stats, fence_hit);
}
#endif
