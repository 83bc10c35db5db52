/* ref_driver.c — drives the UNMODIFIED reference engine (NCBI-BLAST 2.2.28+ core as
 * shipped in OpenHero/gblastn) over in-memory synthetic inputs and records taps.
 *
 * TEST INFRASTRUCTURE ONLY. This file is our own code. It is compiled together with the
 * reference's C sources *where they lie* under /root/reference (oracle/Makefile) into
 * oracle/_ref/libblastref.so. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Taps:
 *   scan    every BlastOffsetPair the scansub callback emits, in emission order
 *           (re-running the scanner loop of BlastNaWordFinder, core/na_ungapped.c:1635-1646)
 *   init    BlastInitHitList after BlastNaWordFinder              (core/na_ungapped.c:1559)
 *   gapped  BlastHSPList after BLAST_GetGappedScore                (core/blast_gapalign.c:3233)
 *   final   BlastHSPList as written to the stream by the engine    (core/blast_engine.c:1309)
 * The first three are captured by ld --wrap on the exported symbols the engine calls
 * through (core/blast_engine.c:926,940), the last by --wrap=BlastHSPStreamWrite.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <pthread.h>

#include <algo/blast/core/blast_def.h>
#include <algo/blast/core/blast_options.h>
#include <algo/blast/core/blast_setup.h>
#include <algo/blast/core/blast_engine.h>
#include <algo/blast/core/blast_filter.h>
#include <algo/blast/core/blast_util.h>
#include <algo/blast/core/blast_hits.h>
#include <algo/blast/core/blast_hspstream.h>
#include <algo/blast/core/blast_seqsrc.h>
#include <algo/blast/core/blast_seqsrc_impl.h>
#include <algo/blast/core/blast_nalookup.h>
#include <algo/blast/core/blast_nascan.h>
#include <algo/blast/core/na_ungapped.h>
#include <algo/blast/core/blast_gapalign.h>
#include <algo/blast/core/blast_parameters.h>
#include <algo/blast/core/blast_diagnostics.h>
#include <algo/blast/core/hspfilter_collector.h>
#include <algo/blast/core/lookup_wrap.h>
#include <algo/blast/core/blast_traceback.h>
#include <algo/blast/core/gapinfo.h>

#include "ref_driver.h"

/* ------------------------------------------------------------------ tables */
static void tab_init(RefTable *t, int ncol) { memset(t, 0, sizeof *t); t->ncol = ncol; }
static int32_t *tab_row(RefTable *t)
{
    if (t->rows == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->data = (int32_t *)realloc(t->data, (size_t)t->cap * t->ncol * sizeof(int32_t));
    }
    return t->data + (t->rows++) * t->ncol;
}

/* ----------------------------------------------------------- in-memory DB */
typedef struct MemDb {
    int32_t n;
    const uint8_t *packed;
    const int64_t *byteoff;
    const int32_t *len;
    int64_t total;
    int32_t maxlen;
    int32_t oid_begin, oid_end;   /* iteration range of this copy */
    int32_t smask_type;           /* 0 none, 1 soft, 2 hard */
    const int32_t *smask_n;       /* masked intervals per subject */
    const int32_t *smask_iv;      /* flat [begin, end) pairs */
    const int64_t *smask_first;   /* index of each subject's first interval (prefix sum) */
    const int64_t *amb_first;     /* ambiguity runs per subject (RefConfig), NULL = none */
    const int32_t *amb_runs;
} MemDb;

/* s_SeqDBRebuildDNA_NA8 + s_SeqDBMapNcbiNA8ToBlastNA8 (objtools/blast/seqdb_reader/seqdbvol.cpp:832-870, 597-633):
 * the runs, in order, written over the blastna bytes of subject oid (buf[0] = base 0) */
static void mdb_overlay_ambiguity(const MemDb *d, Int4 oid, Uint1 *buf, Int4 len)
{
    int64_t k;
    if (!d->amb_first || !d->amb_runs) return;
    for (k = d->amb_first[oid]; k < d->amb_first[oid + 1]; k++) {
        const int32_t a = d->amb_runs[3 * k], n = d->amb_runs[3 * k + 1], code = d->amb_runs[3 * k + 2];
        int32_t j;
        for (j = 0; j < n; j++) if (a + j >= 0 && a + j < len) buf[a + j] = (Uint1)code;
    }
}

/* --------------------------------------------------------------- tap state */
typedef struct TapCtx {
    RefResult *res;
    int taps;
    const MemDb *db;
    Int4 cur_chunk_off;   /* set by the word-finder wrapper, reused by the gapped wrapper */
    /* traceback stage */
    const Uint1 *tb_seq;  /* blastna subject handed out last (base 0) */
    Int4 tb_oid;
    const Uint1 *q_base;  /* query->sequence */
    const BlastQueryInfo *qinfo;
} TapCtx;
static __thread TapCtx *g_tap = NULL;

/* Seam selection (RefConfig.seam): 0 = the reference's own BlastNaWordFinder / BLAST_GetGappedScore; 1 = the
 * B200 engine through oracle/shim/gblastn_b200_shim.c.  The shim is only linked into oracle/_ref/libblastshim.so;
 * in libblastref.so these weak symbols are NULL and seam 1 is refused. */
#include "shim/gblastn_b200_shim.h"
#pragma weak bnshim_word_finder
#pragma weak bnshim_get_gapped_score
#pragma weak bnshim_prelim_begin
#pragma weak bnshim_prelim_end
#pragma weak bnshim_attach_volume
#pragma weak bnshim_detach_volume
#pragma weak bnshim_last_error
static __thread int g_seam = 0;

static Int4 mdb_num_seqs(void *h, void *a) { (void)a; return ((MemDb *)h)->n; }
static Int4 mdb_max_len(void *h, void *a) { (void)a; return ((MemDb *)h)->maxlen; }
static Int4 mdb_min_len(void *h, void *a) { (void)a; (void)h; return 1; }
static Int4 mdb_avg_len(void *h, void *a)
{
    MemDb *d = (MemDb *)h; (void)a;
    return d->n ? (Int4)(d->total / d->n) : 0;
}
static Int8 mdb_tot_len(void *h, void *a) { (void)a; return ((MemDb *)h)->total; }
static Int8 mdb_zero8(void *h, void *a) { (void)h; (void)a; return 0; }
static Int4 mdb_zero4(void *h, void *a) { (void)h; (void)a; return 0; }
static const char *mdb_name(void *h, void *a) { (void)h; (void)a; return "synthetic"; }
static Boolean mdb_is_prot(void *h, void *a) { (void)h; (void)a; return FALSE; }
static Boolean mdb_partial(void *h, void *a) { (void)h; (void)a; return FALSE; }
static void mdb_set_threads(void *h, int n) { (void)h; (void)n; }
static void mdb_reset_iter(void *h) { (void)h; }
static Int4 mdb_seq_len(void *h, void *a) { return ((MemDb *)h)->len[*(Int4 *)a]; }

static Int2 mdb_get_seq(void *h, BlastSeqSrcGetSeqArg *args)
{
    MemDb *d = (MemDb *)h;
    Int4 oid = args->oid;
    if (oid < 0 || oid >= d->n) return BLAST_SEQSRC_ERROR;
    if (args->seq) BlastSequenceBlkClean(args->seq);
    if (args->encoding == eBlastEncodingNucleotide) {
        /* traceback stage: blastna, one base per byte, sentinel bytes on both sides
         * (what s_SeqDbGetSequence hands out for this encoding, api/seqsrc_seqdb.cpp:283-388) */
        const Int4 len = d->len[oid];
        const uint8_t *pk = d->packed + d->byteoff[oid];
        Uint1 *buf = (Uint1 *)malloc((size_t)len + 2);
        Int4 k;
        if (!buf) return BLAST_SEQSRC_ERROR;
        buf[0] = buf[len + 1] = 15;
        for (k = 0; k < len; k++) buf[k + 1] = (pk[k >> 2] >> (6 - 2 * (k & 3))) & 3;
        mdb_overlay_ambiguity(d, oid, buf + 1, len);
        BlastSetUp_SeqBlkNew(buf, len, &args->seq, TRUE);
        args->seq->oid = oid;
        if (g_tap) { g_tap->tb_seq = args->seq->sequence; g_tap->tb_oid = oid; }
        return BLAST_SEQSRC_SUCCESS;
    }
    BlastSetUp_SeqBlkNew(d->packed + d->byteoff[oid], d->len[oid], &args->seq, FALSE);
    args->seq->oid = oid;
    if (d->smask_type && d->smask_n) {      /* every sequence of a masked database carries ranges (one when it has no mask) */
        /* what s_SeqDbGetSequence does with CSeqDB's mask list (api/seqsrc_seqdb.cpp:283-388): the unmasked
         * ranges (0, m0.begin), (m0.end, m1.begin) ... (m_last.end, length) */
        const int32_t n = d->smask_n[oid];
        const int32_t *iv = d->smask_iv + 2 * d->smask_first[oid];
        SSeqRange *r = (SSeqRange *)calloc((size_t)n + 1, sizeof(SSeqRange));
        int32_t k;
        for (k = 0; k < n; k++) { r[k].right = iv[2 * k]; r[k + 1].left = iv[2 * k + 1]; }
        BlastSeqBlkSetSeqRanges(args->seq, r, (Uint4)n + 1, TRUE,
                                d->smask_type == 2 ? eHardSubjMasking : eSoftSubjMasking);
        free(r);
    }
    return BLAST_SEQSRC_SUCCESS;
}
static void mdb_release_seq(void *h, BlastSeqSrcGetSeqArg *args) { (void)h; (void)args; }

static Int4 mdb_iter_next(void *h, BlastSeqSrcIterator *itr)
{
    MemDb *d = (MemDb *)h;
    if (itr->current_pos == UINT4_MAX) itr->current_pos = (unsigned)d->oid_begin;
    if ((Int4)itr->current_pos >= d->oid_end) return BLAST_SEQSRC_EOF;
    return (Int4)itr->current_pos++;
}
static BlastSeqSrc *mdb_free(BlastSeqSrc *s)
{
    if (s) free(_BlastSeqSrcImpl_GetDataStructure(s));
    return NULL;
}
static BlastSeqSrc *mdb_copy(BlastSeqSrc *s)
{
    MemDb *d = (MemDb *)malloc(sizeof(MemDb));
    *d = *(MemDb *)_BlastSeqSrcImpl_GetDataStructure(s);
    _BlastSeqSrcImpl_SetDataStructure(s, d);
    return s;
}
static BlastSeqSrc *mdb_new(BlastSeqSrc *r, void *arg)
{
    MemDb *d = (MemDb *)malloc(sizeof(MemDb));
    *d = *(MemDb *)arg;
    _BlastSeqSrcImpl_SetDeleteFnPtr(r, &mdb_free);
    _BlastSeqSrcImpl_SetCopyFnPtr(r, &mdb_copy);
    _BlastSeqSrcImpl_SetDataStructure(r, d);
    _BlastSeqSrcImpl_SetGetNumSeqs(r, &mdb_num_seqs);
    _BlastSeqSrcImpl_SetGetNumSeqsStats(r, &mdb_zero4);
    _BlastSeqSrcImpl_SetGetMaxSeqLen(r, &mdb_max_len);
    _BlastSeqSrcImpl_SetGetMinSeqLen(r, &mdb_min_len);
    _BlastSeqSrcImpl_SetGetAvgSeqLen(r, &mdb_avg_len);
    _BlastSeqSrcImpl_SetGetTotLen(r, &mdb_tot_len);
    _BlastSeqSrcImpl_SetGetTotLenStats(r, &mdb_zero8);
    _BlastSeqSrcImpl_SetGetName(r, &mdb_name);
    _BlastSeqSrcImpl_SetGetIsProt(r, &mdb_is_prot);
    _BlastSeqSrcImpl_SetGetSupportsPartialFetching(r, &mdb_partial);
    _BlastSeqSrcImpl_SetGetSequence(r, &mdb_get_seq);
    _BlastSeqSrcImpl_SetGetSeqLen(r, &mdb_seq_len);
    _BlastSeqSrcImpl_SetIterNext(r, &mdb_iter_next);
    _BlastSeqSrcImpl_SetResetChunkIterator(r, &mdb_reset_iter);
    _BlastSeqSrcImpl_SetReleaseSequence(r, &mdb_release_seq);
    _BlastSeqSrcImpl_SetSetNumberOfThreads(r, &mdb_set_threads);
    return r;
}


Int2 __real_BlastNaWordFinder(BLAST_SequenceBlk *subject, BLAST_SequenceBlk *query,
                              BlastQueryInfo *query_info, LookupTableWrap *lookup_wrap,
                              Int4 **matrix, const BlastInitialWordParameters *word_params,
                              Blast_ExtendWord *ewp, BlastOffsetPair *offset_pairs,
                              Int4 max_hits, BlastInitHitList *init_hitlist,
                              BlastUngappedStats *ungapped_stats);

Int2 __wrap_BlastNaWordFinder(BLAST_SequenceBlk *subject, BLAST_SequenceBlk *query,
                              BlastQueryInfo *query_info, LookupTableWrap *lookup_wrap,
                              Int4 **matrix, const BlastInitialWordParameters *word_params,
                              Blast_ExtendWord *ewp, BlastOffsetPair *offset_pairs,
                              Int4 max_hits, BlastInitHitList *init_hitlist,
                              BlastUngappedStats *ungapped_stats)
{
    TapCtx *t = g_tap;
    Int4 chunk_off = 0;
    Int2 st;
    if (t) {
        const uint8_t *base = t->db->packed + t->db->byteoff[subject->oid];
        chunk_off = (Int4)((subject->sequence - base) * 4);
        t->cur_chunk_off = chunk_off;
    }
    if (t && (t->taps & 1) && subject->mask_type == eNoSubjMasking) {
        /* replay of the scanner loop only (core/na_ungapped.c:1609-1611,1635-1637) */
        TNaScanSubjectFunction scansub = NULL;
        Int4 lut_word_length = 0;
        Int4 scan_range[3];
        if (lookup_wrap->lut_type == eMBLookupTable) {
            BlastMBLookupTable *l = (BlastMBLookupTable *)lookup_wrap->lut;
            scansub = (TNaScanSubjectFunction)l->scansub_callback;
            lut_word_length = l->lut_word_length;
        } else if (lookup_wrap->lut_type == eSmallNaLookupTable) {
            BlastSmallNaLookupTable *l = (BlastSmallNaLookupTable *)lookup_wrap->lut;
            scansub = (TNaScanSubjectFunction)l->scansub_callback;
            lut_word_length = l->lut_word_length;
        } else {
            BlastNaLookupTable *l = (BlastNaLookupTable *)lookup_wrap->lut;
            scansub = (TNaScanSubjectFunction)l->scansub_callback;
            lut_word_length = l->lut_word_length;
        }
        scan_range[0] = 0;
        scan_range[1] = 0;
        scan_range[2] = subject->length - lut_word_length;
        while (scan_range[1] <= scan_range[2]) {
            Int4 i, n = scansub(lookup_wrap, subject, offset_pairs, max_hits, &scan_range[1]);
            for (i = 0; i < n; i++) {
                int32_t *r = tab_row(&t->res->scan);
                r[0] = subject->oid; r[1] = chunk_off;
                r[2] = (int32_t)offset_pairs[i].qs_offsets.q_off;
                r[3] = (int32_t)offset_pairs[i].qs_offsets.s_off;
            }
        }
    }
    if (g_seam)
        st = bnshim_word_finder(subject, query, query_info, lookup_wrap, matrix, word_params,
                                ewp, offset_pairs, max_hits, init_hitlist, ungapped_stats);
    else
        st = __real_BlastNaWordFinder(subject, query, query_info, lookup_wrap, matrix, word_params,
                                      ewp, offset_pairs, max_hits, init_hitlist, ungapped_stats);
    if (st) return st;
    if (t && (t->taps & 2)) {
        Int4 i;
        for (i = 0; i < init_hitlist->total; i++) {
            BlastInitHSP *h = &init_hitlist->init_hsp_array[i];
            int32_t *r = tab_row(&t->res->init);
            r[0] = subject->oid; r[1] = chunk_off;
            r[2] = (int32_t)h->offsets.qs_offsets.q_off;
            r[3] = (int32_t)h->offsets.qs_offsets.s_off;
            if (h->ungapped_data) {
                r[4] = h->ungapped_data->q_start; r[5] = h->ungapped_data->s_start;
                r[6] = h->ungapped_data->length;  r[7] = h->ungapped_data->score;
            } else { r[4] = r[5] = r[6] = r[7] = -1; }
        }
    }
    return st;
}

Int2 __real_BLAST_GetGappedScore(EBlastProgramType program_number, BLAST_SequenceBlk *query,
                                 BlastQueryInfo *query_info, BLAST_SequenceBlk *subject,
                                 BlastGapAlignStruct *gap_align,
                                 const BlastScoringParameters *score_params,
                                 const BlastExtensionParameters *ext_params,
                                 const BlastHitSavingParameters *hit_params,
                                 BlastInitHitList *init_hitlist, BlastHSPList **hsp_list_ptr,
                                 BlastGappedStats *gapped_stats, Boolean *fence_hit);

Int2 __wrap_BLAST_GetGappedScore(EBlastProgramType program_number, BLAST_SequenceBlk *query,
                                 BlastQueryInfo *query_info, BLAST_SequenceBlk *subject,
                                 BlastGapAlignStruct *gap_align,
                                 const BlastScoringParameters *score_params,
                                 const BlastExtensionParameters *ext_params,
                                 const BlastHitSavingParameters *hit_params,
                                 BlastInitHitList *init_hitlist, BlastHSPList **hsp_list_ptr,
                                 BlastGappedStats *gapped_stats, Boolean *fence_hit)
{
    TapCtx *t = g_tap;
    Int2 st;
    if (g_seam)
        st = bnshim_get_gapped_score(program_number, query, query_info, subject, gap_align, score_params,
                                     ext_params, hit_params, init_hitlist, hsp_list_ptr, gapped_stats, fence_hit);
    else
        st = __real_BLAST_GetGappedScore(program_number, query, query_info, subject, gap_align,
                                         score_params, ext_params, hit_params, init_hitlist,
                                         hsp_list_ptr, gapped_stats, fence_hit);
    if (t && (t->taps & 4) && hsp_list_ptr && *hsp_list_ptr) {
        BlastHSPList *l = *hsp_list_ptr;
        Int4 i;
        for (i = 0; i < l->hspcnt; i++) {
            BlastHSP *h = l->hsp_array[i];
            int32_t *r = tab_row(&t->res->gapped);
            r[0] = subject->oid; r[1] = t->cur_chunk_off; r[2] = h->context;
            r[3] = h->query.offset; r[4] = h->query.end;
            r[5] = h->subject.offset; r[6] = h->subject.end; r[7] = h->score;
            r[8] = h->query.gapped_start; r[9] = h->subject.gapped_start;
        }
    }
    return st;
}

/* BLAST_GapAlignSetUp (core/blast_setup.c) is what Blast_RunPreliminarySearchWithInterrupt calls to create the
 * parameter blocks right before BLAST_PreliminarySearchEngine (core/blast_engine.c:1407-1420): the seams do not
 * receive them all, so the binding picks them up here. */
Int2 __real_BLAST_GapAlignSetUp(EBlastProgramType program_number, const BlastSeqSrc *seq_src,
                                const BlastScoringOptions *scoring_options,
                                const BlastEffectiveLengthsOptions *eff_len_options,
                                const BlastExtensionOptions *ext_options, const BlastHitSavingOptions *hit_options,
                                BlastQueryInfo *query_info, BlastScoreBlk *sbp, BlastScoringParameters **score_params,
                                BlastExtensionParameters **ext_params, BlastHitSavingParameters **hit_params,
                                BlastEffectiveLengthsParameters **eff_len_params, BlastGapAlignStruct **gap_align);
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
    if (g_seam && st == 0) bnshim_prelim_begin(*score_params, *ext_params, *hit_params, *gap_align);
    return st;
}

int __real_BlastHSPStreamWrite(BlastHSPStream *hsp_stream, BlastHSPList **hsp_list);
int __wrap_BlastHSPStreamWrite(BlastHSPStream *hsp_stream, BlastHSPList **hsp_list)
{
    TapCtx *t = g_tap;
    if (t && hsp_list && *hsp_list) {
        BlastHSPList *l = *hsp_list;
        Int4 i;
        for (i = 0; i < l->hspcnt; i++) {
            BlastHSP *h = l->hsp_array[i];
            int32_t *r = tab_row(&t->res->final_);
            uint64_t eb;
            memcpy(&eb, &h->evalue, 8);
            r[0] = l->oid; r[1] = h->context;
            r[2] = h->query.offset; r[3] = h->query.end;
            r[4] = h->subject.offset; r[5] = h->subject.end; r[6] = h->score;
            r[7] = h->query.gapped_start; r[8] = h->subject.gapped_start;
            r[9] = (int32_t)(uint32_t)(eb & 0xffffffffu);
            r[10] = (int32_t)(uint32_t)(eb >> 32);
        }
    }
    return __real_BlastHSPStreamWrite(hsp_stream, hsp_list);
}

/* ---- traceback stage: the reference's own calls of its alignment-with-traceback routines ---------- */
static void tap_tb_call(TapCtx *t, int kind, const Uint1 *query, const Uint1 *subject, Int4 q_start, Int4 s_start,
                        Int4 q_len, Int4 s_len, const BlastGapAlignStruct *ga)
{
    int32_t *r = tab_row(&t->res->tb_calls);
    const Int4 qoff = (Int4)(query - t->q_base);
    Int4 ctx = -1, c, i;
    for (c = t->qinfo->first_context; c <= t->qinfo->last_context; c++)
        if (t->qinfo->contexts[c].is_valid && t->qinfo->contexts[c].query_offset == qoff) { ctx = c; break; }
    r[0] = kind; r[1] = t->tb_oid; r[2] = ctx; r[3] = (int32_t)(subject - t->tb_seq);
    r[4] = q_start; r[5] = s_start; r[6] = q_len; r[7] = s_len;
    r[8] = ga->score; r[9] = ga->query_start; r[10] = ga->query_stop;
    r[11] = ga->subject_start; r[12] = ga->subject_stop;
    r[13] = (int32_t)t->res->tb_ops.rows; r[14] = 0;
    if (ga->edit_script) {
        r = NULL;   /* tab_row may move the table */
        for (i = 0; i < ga->edit_script->size; i++) {
            int32_t *o = tab_row(&t->res->tb_ops);
            o[0] = (int32_t)ga->edit_script->op_type[i]; o[1] = ga->edit_script->num[i];
        }
        t->res->tb_calls.data[(t->res->tb_calls.rows - 1) * 15 + 14] = ga->edit_script->size;
    }
}

Int2 __real_BLAST_GappedAlignmentWithTraceback(EBlastProgramType program, const Uint1 *query, const Uint1 *subject,
                                               BlastGapAlignStruct *gap_align, const BlastScoringParameters *score_params,
                                               Int4 q_start, Int4 s_start, Int4 query_length, Int4 subject_length,
                                               Boolean *fence_hit);
Int2 __wrap_BLAST_GappedAlignmentWithTraceback(EBlastProgramType program, const Uint1 *query, const Uint1 *subject,
                                               BlastGapAlignStruct *gap_align, const BlastScoringParameters *score_params,
                                               Int4 q_start, Int4 s_start, Int4 query_length, Int4 subject_length,
                                               Boolean *fence_hit)
{
    TapCtx *t = g_tap;
    Int2 st = __real_BLAST_GappedAlignmentWithTraceback(program, query, subject, gap_align, score_params, q_start,
                                                        s_start, query_length, subject_length, fence_hit);
    if (t && (t->taps & 16) && t->tb_seq)
        tap_tb_call(t, 0, query, subject, q_start, s_start, query_length, subject_length, gap_align);
    return st;
}

Int2 __real_BLAST_GreedyGappedAlignment(const Uint1 *query, const Uint1 *subject, Int4 query_length, Int4 subject_length,
                                        BlastGapAlignStruct *gap_align, const BlastScoringParameters *score_params,
                                        Int4 q_off, Int4 s_off, Boolean compressed_subject, Boolean do_traceback,
                                        Boolean *fence_hit);
Int2 __wrap_BLAST_GreedyGappedAlignment(const Uint1 *query, const Uint1 *subject, Int4 query_length, Int4 subject_length,
                                        BlastGapAlignStruct *gap_align, const BlastScoringParameters *score_params,
                                        Int4 q_off, Int4 s_off, Boolean compressed_subject, Boolean do_traceback,
                                        Boolean *fence_hit)
{
    TapCtx *t = g_tap;
    Int2 st = __real_BLAST_GreedyGappedAlignment(query, subject, query_length, subject_length, gap_align, score_params,
                                                 q_off, s_off, compressed_subject, do_traceback, fence_hit);
    if (t && (t->taps & 16) && do_traceback && !compressed_subject && t->tb_seq)
        tap_tb_call(t, 1, query, subject, q_off, s_off, query_length, subject_length, gap_align);
    return st;
}

/* final results of the traceback stage: every HSP of every hit list, in the order of BlastHSPResults */
static void dump_tb_results(RefResult *res, const BlastHSPResults *results)
{
    Int4 qi, li, hi, k;
    if (!results) return;
    for (qi = 0; qi < results->num_queries; qi++) {
        const BlastHitList *hl = results->hitlist_array[qi];
        if (!hl) continue;
        for (li = 0; li < hl->hsplist_count; li++) {
            const BlastHSPList *l = hl->hsplist_array[li];
            if (!l) continue;
            for (hi = 0; hi < l->hspcnt; hi++) {
                const BlastHSP *h = l->hsp_array[hi];
                int32_t *r;
                uint64_t eb, bb;
                const int32_t off = (int32_t)res->tb_ops.rows;
                int32_t n = 0;
                if (h->gap_info) {
                    n = h->gap_info->size;
                    for (k = 0; k < n; k++) {
                        int32_t *o = tab_row(&res->tb_ops);
                        o[0] = (int32_t)h->gap_info->op_type[k]; o[1] = h->gap_info->num[k];
                    }
                }
                r = tab_row(&res->tb_final);
                memcpy(&eb, &h->evalue, 8); memcpy(&bb, &h->bit_score, 8);
                r[0] = qi; r[1] = l->oid; r[2] = h->context;
                r[3] = h->query.offset; r[4] = h->query.end; r[5] = h->subject.offset; r[6] = h->subject.end;
                r[7] = h->score; r[8] = h->num_ident;
                r[9] = (int32_t)(uint32_t)(eb & 0xffffffffu); r[10] = (int32_t)(uint32_t)(eb >> 32);
                r[11] = (int32_t)(uint32_t)(bb & 0xffffffffu); r[12] = (int32_t)(uint32_t)(bb >> 32);
                r[13] = off; r[14] = n;
            }
        }
    }
}

/* ------------------------------------------------------------ query set-up */
static const uint8_t kBlastnaComplement[16] = {
    /* A C G T  R Y M K  W S  B  D  H  V  N gap */
    3, 2, 1, 0, 5, 4, 7, 6, 8, 9, 13, 12, 11, 10, 14, 15
};

typedef struct Setup {
    LookupTableOptions *lookup_options;
    QuerySetUpOptions *query_options;
    BlastInitialWordOptions *word_options;
    BlastExtensionOptions *ext_options;
    BlastHitSavingOptions *hit_options;
    BlastScoringOptions *score_options;
    BlastEffectiveLengthsOptions *eff_len_options;
    PSIBlastOptions *psi_options;
    BlastDatabaseOptions *db_options;
    BLAST_SequenceBlk *query;
    BlastQueryInfo *query_info;
    BlastScoreBlk *sbp;
    BlastSeqLoc *lookup_segments;
    LookupTableWrap *lookup_wrap;
} Setup;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int build_setup(const RefConfig *cfg, int32_t nq, const uint8_t *qseq, const int32_t *qlens,
                       const int32_t *qmask_n, const int32_t *qmask_iv, Setup *S)
{
    const EBlastProgramType prog = eBlastTypeBlastn;
    const int mb = (cfg->task == 0);
    Blast_Message *msg = NULL;
    Int2 st;
    int32_t i, c;
    int64_t total = 0;
    uint8_t *buf;
    const uint8_t *src = qseq;
    const int32_t *iv = qmask_iv;

    memset(S, 0, sizeof *S);
    st = BLAST_InitDefaultOptions(prog, &S->lookup_options, &S->query_options, &S->word_options,
                                  &S->ext_options, &S->hit_options, &S->score_options,
                                  &S->eff_len_options, &S->psi_options, &S->db_options);
    if (st) return 100 + st;

    /* options as the blastn CLI sets them (api/blast_nucl_options.cpp:95-240) */
    S->lookup_options->lut_type = mb ? eMBLookupTable : eNaLookupTable;
    S->lookup_options->word_size = cfg->word_size ? cfg->word_size : (mb ? 28 : 11);
    S->lookup_options->threshold = 0;
    S->word_options->window_size = cfg->window_size;
    S->word_options->scan_range = cfg->scan_range;
    S->word_options->x_dropoff = cfg->xdrop_ungap > 0 ? cfg->xdrop_ungap : BLAST_UNGAPPED_X_DROPOFF_NUCL;
    S->word_options->gap_trigger = BLAST_GAP_TRIGGER_NUCL;
    S->score_options->reward = (Int2)(cfg->reward ? cfg->reward : (mb ? 1 : 2));
    S->score_options->penalty = (Int2)(cfg->penalty ? cfg->penalty : (mb ? -2 : -3));
    S->score_options->gap_open = cfg->gap_open >= 0 ? cfg->gap_open : (mb ? 0 : 5);
    S->score_options->gap_extend = cfg->gap_extend >= 0 ? cfg->gap_extend : (mb ? 0 : 2);
    S->score_options->gapped_calculation = TRUE;
    {
        int greedy = cfg->greedy >= 0 ? cfg->greedy : mb;
        S->ext_options->ePrelimGapExt = greedy ? eGreedyScoreOnly : eDynProgScoreOnly;
        S->ext_options->eTbackExt = greedy ? eGreedyTbck : eDynProgTbck;
        S->ext_options->gap_x_dropoff = cfg->xdrop_gap > 0 ? cfg->xdrop_gap
                                        : (greedy ? BLAST_GAP_X_DROPOFF_GREEDY : BLAST_GAP_X_DROPOFF_NUCL);
        S->ext_options->gap_x_dropoff_final = cfg->xdrop_gap_final > 0 ? cfg->xdrop_gap_final
                                              : BLAST_GAP_X_DROPOFF_FINAL_NUCL;
    }
    S->hit_options->hitlist_size = cfg->hitlist_size ? cfg->hitlist_size : 500;
    S->hit_options->expect_value = cfg->evalue > 0 ? cfg->evalue : BLAST_EXPECT_VALUE;
    S->hit_options->min_diag_separation =
        cfg->min_diag_separation >= 0 ? cfg->min_diag_separation : (mb ? 6 : 50);
    S->hit_options->mask_level = 101;
    S->hit_options->low_score_perc = cfg->low_score_perc >= 0 ? cfg->low_score_perc : 0.15;
    S->hit_options->hsp_num_max = cfg->hsp_num_max;
    S->hit_options->percent_identity = cfg->percent_identity;
    S->hit_options->min_hit_length = cfg->min_hit_length;
    S->eff_len_options->db_length = cfg->db_length;
    S->eff_len_options->dbseq_num = cfg->db_num_seqs;
    S->query_options->strand_option = 3;
    SBlastFilterOptionsNew(&S->query_options->filtering_options, eEmpty);
    S->query_options->filtering_options->mask_at_hash = cfg->mask_at_hash ? TRUE : FALSE;

    /* concatenated query: [sentinel] q0+ [sentinel] q0- [sentinel] q1+ ...    */
    for (i = 0; i < nq; i++) total += 2 * (int64_t)qlens[i] + 2;
    buf = (uint8_t *)malloc((size_t)total + 2);
    S->query_info = BlastQueryInfoNew(prog, nq);
    {
        int64_t pos = 0;
        Uint4 maxlen = 0;
        buf[pos++] = 15; /* kNuclSentinel */
        for (i = 0; i < nq; i++) {
            int32_t L = qlens[i], k;
            for (c = 0; c < 2; c++) {
                BlastContextInfo *ci = &S->query_info->contexts[2 * i + c];
                ci->query_offset = (Int4)(pos - 1);
                ci->query_length = L;
                ci->query_index = i;
                ci->frame = (Int1)(c == 0 ? 1 : -1);
                ci->is_valid = TRUE;
                if (c == 0) for (k = 0; k < L; k++) buf[pos++] = src[k];
                else        for (k = 0; k < L; k++) buf[pos++] = kBlastnaComplement[src[L - 1 - k] & 15];
                buf[pos++] = 15;
            }
            if ((Uint4)L > maxlen) maxlen = (Uint4)L;
            src += L;
        }
        S->query_info->max_length = maxlen;
        BlastSeqBlkNew(&S->query);
        BlastSeqBlkSetSequence(S->query, buf, (Int4)(pos - 2));
        S->query->sequence_start_allocated = TRUE;
    }
    if (qmask_n) {
        BlastMaskLoc *ml = BlastMaskLocNew(2 * nq);
        for (i = 0; i < nq; i++) {
            int32_t k;
            for (c = 0; c < 2; c++) {
                BlastSeqLoc *tail = NULL;
                for (k = 0; k < qmask_n[i]; k++)
                    tail = BlastSeqLocNew(tail ? &tail : &ml->seqloc_array[2 * i + c],
                                          iv[2 * k], iv[2 * k + 1]);
            }
            iv += 2 * qmask_n[i];
        }
        S->query->lcase_mask = ml;
        S->query->lcase_mask_allocated = TRUE;
    }

    st = BLAST_MainSetUp(prog, S->query_options, S->score_options, S->query, S->query_info, 1.0,
                         &S->lookup_segments, NULL, &S->sbp, &msg, NULL);
    if (st) { if (msg) fprintf(stderr, "ref_driver: MainSetUp: %s\n", msg->message); return 200 + st; }
    msg = Blast_MessageFree(msg);
    st = LookupTableWrapInit(S->query, S->lookup_options, S->query_options, S->lookup_segments,
                             S->sbp, &S->lookup_wrap, NULL, &msg);
    if (st) { if (msg) fprintf(stderr, "ref_driver: LookupTableWrapInit: %s\n", msg->message); return 300 + st; }
    msg = Blast_MessageFree(msg);
    return 0;
}

static void free_setup(Setup *S)
{
    S->lookup_wrap = LookupTableWrapFree(S->lookup_wrap);
    S->lookup_segments = BlastSeqLocFree(S->lookup_segments);
    S->sbp = BlastScoreBlkFree(S->sbp);
    S->query = BlastSequenceBlkFree(S->query);
    S->query_info = BlastQueryInfoFree(S->query_info);
    LookupTableOptionsFree(S->lookup_options);
    BlastQuerySetUpOptionsFree(S->query_options);
    BlastInitialWordOptionsFree(S->word_options);
    BlastExtensionOptionsFree(S->ext_options);
    BlastHitSavingOptionsFree(S->hit_options);
    BlastScoringOptionsFree(S->score_options);
    BlastEffectiveLengthsOptionsFree(S->eff_len_options);
    PSIBlastOptionsFree(S->psi_options);
    BlastDatabaseOptionsFree(S->db_options);
}

static void kbp4(double *dst, const Blast_KarlinBlk *k)
{
    if (!k) { dst[0] = dst[1] = dst[2] = dst[3] = -1; return; }
    dst[0] = k->Lambda; dst[1] = k->K; dst[2] = k->logK; dst[3] = k->H;
}

/* Parameters exactly as the engine derives them (core/blast_engine.c:1379-1440, 1114-1150) */
static int dump_params(const Setup *S, const BlastSeqSrc *seq_src, RefResult *res, int dump_lut)
{
    const EBlastProgramType prog = eBlastTypeBlastn;
    BlastScoringParameters *score_params = NULL;
    BlastExtensionParameters *ext_params = NULL;
    BlastHitSavingParameters *hit_params = NULL;
    BlastEffectiveLengthsParameters *eff_len_params = NULL;
    BlastGapAlignStruct *gap_align = NULL;
    BlastInitialWordParameters *word_params = NULL;
    int n = S->query_info->last_context + 1, c, i, j;
    Int2 st = BLAST_GapAlignSetUp(prog, seq_src, S->score_options, S->eff_len_options,
                                  S->ext_options, S->hit_options, S->query_info, S->sbp,
                                  &score_params, &ext_params, &hit_params, &eff_len_params,
                                  &gap_align);
    if (st) return 400 + st;
    BlastInitialWordParametersNew(prog, S->word_options, hit_params, S->lookup_wrap, S->sbp,
                                  S->query_info, BlastSeqSrcGetAvgSeqLen(seq_src), &word_params);
    res->num_contexts = n;
    res->ctx_query_offset = (int32_t *)calloc(n, 4);
    res->ctx_query_length = (int32_t *)calloc(n, 4);
    res->ctx_length_adjustment = (int32_t *)calloc(n, 4);
    res->ctx_eff_searchsp = (int64_t *)calloc(n, 8);
    res->ctx_x_dropoff = (int32_t *)calloc(n, 4);
    res->ctx_cutoff_score = (int32_t *)calloc(n, 4);
    res->ctx_reduced_cutoff = (int32_t *)calloc(n, 4);
    res->ctx_gapped_cutoff = (int32_t *)calloc(n, 4);
    res->ctx_kbp_std = (double *)calloc(4 * n, 8);
    res->ctx_kbp_gap = (double *)calloc(4 * n, 8);
    for (c = 0; c < n; c++) {
        const BlastContextInfo *ci = &S->query_info->contexts[c];
        res->ctx_query_offset[c] = ci->query_offset;
        res->ctx_query_length[c] = ci->query_length;
        res->ctx_length_adjustment[c] = ci->length_adjustment;
        res->ctx_eff_searchsp[c] = ci->eff_searchsp;
        res->ctx_x_dropoff[c] = word_params->cutoffs[c].x_dropoff;
        res->ctx_cutoff_score[c] = word_params->cutoffs[c].cutoff_score;
        res->ctx_reduced_cutoff[c] = word_params->cutoffs[c].reduced_nucl_cutoff_score;
        res->ctx_gapped_cutoff[c] = hit_params->cutoffs[c].cutoff_score;
        kbp4(res->ctx_kbp_std + 4 * c, S->sbp->kbp_std ? S->sbp->kbp_std[c] : NULL);
        kbp4(res->ctx_kbp_gap + 4 * c, S->sbp->kbp_gap ? S->sbp->kbp_gap[c] : NULL);
    }
    res->gap_x_dropoff = ext_params->gap_x_dropoff;
    res->gap_x_dropoff_final = ext_params->gap_x_dropoff_final;
    res->container_type = (word_params->container_type == eDiagHash);
    res->round_down = S->sbp->round_down ? 1 : 0;
    memcpy(res->nucl_score_table, word_params->nucl_score_table, sizeof res->nucl_score_table);
    for (i = 0; i < 16; i++)
        for (j = 0; j < 16; j++) res->matrix[16 * i + j] = S->sbp->matrix->data[i][j];

    res->concat_len = S->query->length;
    res->concat_query = (uint8_t *)malloc((size_t)S->query->length + 2);
    memcpy(res->concat_query, S->query->sequence_start, (size_t)S->query->length + 2);

    if (S->lookup_wrap->lut_type == eMBLookupTable) {
        BlastMBLookupTable *l = (BlastMBLookupTable *)S->lookup_wrap->lut;
        BlastSeqLoc *loc;
        int cnt = 0;
        res->lut_type = 0;
        res->lut_word_length = l->lut_word_length; res->word_length = l->word_length;
        res->scan_step = l->scan_step; res->longest_chain = l->longest_chain;
        res->pv_array_bts = l->pv_array_bts; res->hashsize = l->hashsize;
        res->next_pos_len = (int64_t)S->query->length + 1;
        res->pv_len = l->hashsize >> l->pv_array_bts;
        for (loc = l->masked_locations; loc; loc = loc->next) cnt++;
        res->n_masked_locations = l->masked_locations ? cnt : -1;
        if (cnt) {
            res->masked_locations = (int32_t *)malloc(8 * (size_t)cnt);
            for (cnt = 0, loc = l->masked_locations; loc; loc = loc->next, cnt++) {
                res->masked_locations[2 * cnt] = loc->ssr->left;
                res->masked_locations[2 * cnt + 1] = loc->ssr->right;
            }
        }
        if (dump_lut) {
            res->hashtable = (int32_t *)malloc(4 * (size_t)res->hashsize);
            memcpy(res->hashtable, l->hashtable, 4 * (size_t)res->hashsize);
            res->next_pos = (int32_t *)malloc(4 * (size_t)res->next_pos_len);
            memcpy(res->next_pos, l->next_pos, 4 * (size_t)res->next_pos_len);
            res->pv_array = (uint32_t *)malloc(4 * (size_t)res->pv_len);
            memcpy(res->pv_array, l->pv_array, 4 * (size_t)res->pv_len);
        }
    } else if (S->lookup_wrap->lut_type == eSmallNaLookupTable) {
        BlastSmallNaLookupTable *l = (BlastSmallNaLookupTable *)S->lookup_wrap->lut;
        BlastSeqLoc *loc;
        int cnt = 0;
        res->lut_type = 1;
        res->lut_word_length = l->lut_word_length; res->word_length = l->word_length;
        res->scan_step = l->scan_step; res->longest_chain = l->longest_chain;
        res->hashsize = l->backbone_size; res->overflow_len = l->overflow_size;
        for (loc = l->masked_locations; loc; loc = loc->next) cnt++;
        res->n_masked_locations = l->masked_locations ? cnt : -1;
        if (cnt) {
            res->masked_locations = (int32_t *)malloc(8 * (size_t)cnt);
            for (cnt = 0, loc = l->masked_locations; loc; loc = loc->next, cnt++) {
                res->masked_locations[2 * cnt] = loc->ssr->left;
                res->masked_locations[2 * cnt + 1] = loc->ssr->right;
            }
        }
        if (dump_lut) {
            res->backbone = (int16_t *)malloc(2 * (size_t)res->hashsize);
            memcpy(res->backbone, l->final_backbone, 2 * (size_t)res->hashsize);
            if (res->overflow_len > 0) {
                res->overflow = (int16_t *)malloc(2 * (size_t)res->overflow_len);
                memcpy(res->overflow, l->overflow, 2 * (size_t)res->overflow_len);
            }
        }
    } else {
        BlastNaLookupTable *l = (BlastNaLookupTable *)S->lookup_wrap->lut;
        res->lut_type = 2;
        res->lut_word_length = l->lut_word_length; res->word_length = l->word_length;
        res->scan_step = l->scan_step; res->longest_chain = l->longest_chain;
        res->hashsize = l->backbone_size; res->na_overflow_len = l->overflow_size;
        res->n_masked_locations = -1;
        if (l->masked_locations) {
            BlastSeqLoc *m; int32_t k = 0;
            for (m = l->masked_locations; m; m = m->next) ++k;
            res->n_masked_locations = k;
            res->masked_locations = (int32_t *)malloc(8 * (size_t)(k ? k : 1));
            for (m = l->masked_locations, k = 0; m; m = m->next, ++k) {
                res->masked_locations[2 * k] = m->ssr->left; res->masked_locations[2 * k + 1] = m->ssr->right;
            }
        }
        if (dump_lut) {
            res->na_backbone = (int32_t *)malloc(16 * (size_t)res->hashsize);
            memcpy(res->na_backbone, l->thick_backbone, 16 * (size_t)res->hashsize);
            if (res->na_overflow_len > 0) {
                res->na_overflow = (int32_t *)malloc(4 * (size_t)res->na_overflow_len);
                memcpy(res->na_overflow, l->overflow, 4 * (size_t)res->na_overflow_len);
            }
        }
    }

    word_params = BlastInitialWordParametersFree(word_params);
    gap_align->sbp = NULL;
    BLAST_GapAlignStructFree(gap_align);
    BlastScoringParametersFree(score_params);
    BlastHitSavingParametersFree(hit_params);
    BlastExtensionParametersFree(ext_params);
    BlastEffectiveLengthsParametersFree(eff_len_params);
    return 0;
}

typedef struct Worker {
    const Setup *S;
    MemDb db;
    RefResult *res;     /* thread-private result for final_ rows */
    int taps;
    int traceback;
    int seam;
    int status;
    BlastDiagnostics *diag;
    pthread_t th;
} Worker;

static void *worker_main(void *arg)
{
    Worker *w = (Worker *)arg;
    const Setup *S = w->S;
    const EBlastProgramType prog = eBlastTypeBlastn;
    BlastSeqSrcNewInfo info;
    BlastSeqSrc *seq_src;
    BlastHSPStream *stream;
    BlastHSPWriterInfo *winfo;
    BlastHSPWriter *writer;
    TapCtx tap;

    info.constructor = &mdb_new;
    info.ctor_argument = &w->db;
    seq_src = BlastSeqSrcNew(&info);

    winfo = BlastHSPCollectorInfoNew(
        BlastHSPCollectorParamsNew(S->hit_options, S->ext_options->compositionBasedStats,
                                   S->score_options->gapped_calculation));
    writer = BlastHSPWriterNew(&winfo, S->query_info);
    stream = BlastHSPStreamNew(prog, S->ext_options, TRUE, S->query_info->num_queries, writer);

    tap.res = w->res; tap.taps = w->taps; tap.db = &w->db; tap.cur_chunk_off = 0;
    tap.tb_seq = NULL; tap.tb_oid = -1; tap.q_base = S->query->sequence; tap.qinfo = S->query_info;
    g_tap = &tap;
    g_seam = w->seam;
    if (g_seam) {
        /* the volume this thread's seqsrc hands out: last sequence's bytes + the 16 pad bytes every volume carries */
        const int32_t last = w->db.n - 1;
        const int64_t bytes = last >= 0 ? w->db.byteoff[last] + w->db.len[last] / 4 + 1 + 16 : 16;
        if (bnshim_attach_volume(w->db.packed, bytes, w->db.byteoff, w->db.len, w->db.n, 0)) {
            fprintf(stderr, "ref_driver: %s\n", bnshim_last_error());
            w->status = 90; g_tap = NULL; g_seam = 0;
            BlastHSPStreamFree(stream); BlastSeqSrcFree(seq_src);
            return NULL;
        }
    }
    w->diag = Blast_DiagnosticsInit();
    w->status = Blast_RunPreliminarySearch(prog, S->query, S->query_info, seq_src,
                                           S->score_options, S->sbp, S->lookup_wrap,
                                           S->word_options, S->ext_options, S->hit_options,
                                           S->eff_len_options, S->psi_options, S->db_options,
                                           stream, w->diag);
    if (g_seam) {
        bnshim_prelim_end();
        if (w->status == 0 && bnshim_last_error()[0]) {      /* the engine ignores the word finder's status */
            fprintf(stderr, "ref_driver: %s\n", bnshim_last_error());
            w->status = 92;
        }
    }
    if (w->status == 0 && w->traceback) {
        BlastHSPResults *results = NULL;
        const double tb0 = now_s();
        w->status = Blast_RunTracebackSearch(prog, S->query, S->query_info, seq_src, S->score_options,
                                             S->ext_options, S->hit_options, S->eff_len_options, S->db_options,
                                             S->psi_options, S->sbp, stream, NULL, NULL, &results);
        w->res->seconds_traceback = now_s() - tb0;
        if (w->status == 0) dump_tb_results(w->res, results);
        Blast_HSPResultsFree(results);
    }
    g_tap = NULL;
    if (g_seam) { bnshim_detach_volume(); g_seam = 0; }
    if (w->status == 0 && !w->traceback && w->res->kept.ncol && stream->results) {
        Int4 qi, li;
        for (qi = 0; qi < stream->results->num_queries; qi++) {
            const BlastHitList *hl = stream->results->hitlist_array[qi];
            if (!hl) continue;
            for (li = 0; li < hl->hsplist_count; li++) {
                const BlastHSPList *l = hl->hsplist_array[li];
                int32_t *r;
                if (!l || l->hspcnt == 0) continue;
                r = tab_row(&w->res->kept);
                r[0] = qi; r[1] = l->oid; r[2] = l->hsp_array[0]->score; r[3] = l->hspcnt;
            }
        }
    }
    BlastHSPStreamFree(stream);
    BlastSeqSrcFree(seq_src);
    return NULL;
}

int ref_search(const RefConfig *cfg,
               int32_t nq, const uint8_t *qseq, const int32_t *qlens,
               const int32_t *qmask_n, const int32_t *qmask_iv,
               int32_t ns, const uint8_t *packed, const int64_t *sbyteoff, const int32_t *slen,
               RefResult *res)
{
    Setup S;
    MemDb db;
    int st, i, nth;
    Worker *ws;
    double t0;

    memset(res, 0, sizeof *res);
    tab_init(&res->scan, 4);
    tab_init(&res->init, 8);
    tab_init(&res->gapped, 10);
    tab_init(&res->final_, 11);
    tab_init(&res->tb_calls, 15);
    tab_init(&res->tb_ops, 2);
    tab_init(&res->tb_final, 15);
    tab_init(&res->kept, 4);

    if (cfg->seam && (!bnshim_word_finder || cfg->smask_n)) {
        fprintf(stderr, "ref_driver: seam 1 needs libblastshim.so (and no database masks)\n");
        res->status = 91; return 91;
    }
    st = build_setup(cfg, nq, qseq, qlens, qmask_n, qmask_iv, &S);
    if (st) { res->status = st; return st; }

    db.n = ns; db.packed = packed; db.byteoff = sbyteoff; db.len = slen;
    db.total = 0; db.maxlen = 0; db.oid_begin = 0; db.oid_end = ns;
    for (i = 0; i < ns; i++) { db.total += slen[i]; if (slen[i] > db.maxlen) db.maxlen = slen[i]; }
    db.smask_type = cfg->smask_n ? cfg->smask_type : 0;
    db.smask_n = cfg->smask_n; db.smask_iv = cfg->smask_iv; db.smask_first = NULL;
    db.amb_first = cfg->amb_first; db.amb_runs = cfg->amb_runs;
    if (db.smask_type) {
        int64_t *first = (int64_t *)calloc((size_t)ns + 1, sizeof(int64_t));
        for (i = 0; i < ns; i++) first[i + 1] = first[i] + cfg->smask_n[i];
        db.smask_first = first;
    }

    /* the engine's callback choice mutates the table: do it once, before threads start */
    BlastChooseNucleotideScanSubject(S.lookup_wrap);
    BlastChooseNaExtend(S.lookup_wrap);

    {
        BlastSeqSrcNewInfo info;
        BlastSeqSrc *seq_src;
        info.constructor = &mdb_new; info.ctor_argument = &db;
        seq_src = BlastSeqSrcNew(&info);
        st = dump_params(&S, seq_src, res, (cfg->taps & 8) != 0);
        BlastSeqSrcFree(seq_src);
        if (st) { res->status = st; free_setup(&S); return st; }
    }

    nth = cfg->num_threads > 1 ? cfg->num_threads : 1;
    if (nth > ns) nth = ns > 0 ? ns : 1;
    ws = (Worker *)calloc(nth, sizeof(Worker));
    for (i = 0; i < nth; i++) {
        ws[i].S = &S; ws[i].db = db; ws[i].seam = cfg->seam;
        ws[i].db.oid_begin = (int32_t)((int64_t)ns * i / nth);
        ws[i].db.oid_end = (int32_t)((int64_t)ns * (i + 1) / nth);
        if (nth == 1) { ws[i].res = res; ws[i].taps = cfg->taps; ws[i].traceback = (cfg->prelim_only == 0); }
        else {
            ws[i].res = (RefResult *)calloc(1, sizeof(RefResult));
            tab_init(&ws[i].res->final_, 11);
            ws[i].taps = 0;
        }
    }
    t0 = now_s();
    if (nth == 1) worker_main(&ws[0]);
    else {
        for (i = 0; i < nth; i++) pthread_create(&ws[i].th, NULL, worker_main, &ws[i]);
        for (i = 0; i < nth; i++) pthread_join(ws[i].th, NULL);
    }
    res->seconds_prelim = now_s() - t0 - res->seconds_traceback;

    st = 0;
    for (i = 0; i < nth; i++) {
        if (ws[i].status) st = 500 + ws[i].status;
        if (ws[i].diag) {
            if (ws[i].diag->ungapped_stat) {
                res->lookup_hits += ws[i].diag->ungapped_stat->lookup_hits;
                res->init_extends += ws[i].diag->ungapped_stat->init_extends;
                res->good_init_extends += ws[i].diag->ungapped_stat->good_init_extends;
            }
            if (ws[i].diag->gapped_stat) {
                res->gap_extensions += ws[i].diag->gapped_stat->extensions;
                res->good_extensions += ws[i].diag->gapped_stat->good_extensions;
            }
            Blast_DiagnosticsFree(ws[i].diag);
        }
        if (nth > 1) {
            int64_t r;
            for (r = 0; r < ws[i].res->final_.rows; r++)
                memcpy(tab_row(&res->final_), ws[i].res->final_.data + r * 11, 44);
            free(ws[i].res->final_.data);
            free(ws[i].res);
        }
    }
    free(ws);
    free_setup(&S);
    res->status = st;
    return st;
}

/* Direct differential oracle for the traceback-stage alignment routines: runs the reference's
 * BLAST_GappedAlignmentWithTraceback (eDynProgTbck) or BLAST_GreedyGappedAlignment with traceback (eGreedyTbck)
 * on arbitrary start points.  items: 6 ints per call {oid, context, s_shift, s_length, q_start, s_start};
 * the calls are recorded in res->tb_calls / tb_ops through the same wrappers as the taps. */
int ref_traceback_calls(const RefConfig *cfg,
                        int32_t nq, const uint8_t *qseq, const int32_t *qlens,
                        const int32_t *qmask_n, const int32_t *qmask_iv,
                        int32_t ns, const uint8_t *packed, const int64_t *sbyteoff, const int32_t *slen,
                        int32_t n_items, const int32_t *items, RefResult *res)
{
    const EBlastProgramType prog = eBlastTypeBlastn;
    Setup S;
    MemDb db;
    BlastSeqSrcNewInfo info;
    BlastSeqSrc *seq_src;
    BlastScoringParameters *score_params = NULL;
    BlastExtensionParameters *ext_params = NULL;
    BlastHitSavingParameters *hit_params = NULL;
    BlastEffectiveLengthsParameters *eff_len_params = NULL;
    BlastGapAlignStruct *gap_align = NULL;
    TapCtx tap;
    Uint1 *buf = NULL;
    Int4 cur_oid = -1;
    int st, i;

    memset(res, 0, sizeof *res);
    tab_init(&res->scan, 4); tab_init(&res->init, 8); tab_init(&res->gapped, 10); tab_init(&res->final_, 11);
    tab_init(&res->tb_calls, 15); tab_init(&res->tb_ops, 2); tab_init(&res->tb_final, 15);
    st = build_setup(cfg, nq, qseq, qlens, qmask_n, qmask_iv, &S);
    if (st) { res->status = st; return st; }
    db.n = ns; db.packed = packed; db.byteoff = sbyteoff; db.len = slen;
    db.total = 0; db.maxlen = 0; db.oid_begin = 0; db.oid_end = ns;
    for (i = 0; i < ns; i++) { db.total += slen[i]; if (slen[i] > db.maxlen) db.maxlen = slen[i]; }
    db.smask_type = 0; db.smask_n = NULL; db.smask_iv = NULL; db.smask_first = NULL;
    db.amb_first = cfg->amb_first; db.amb_runs = cfg->amb_runs;
    BlastChooseNucleotideScanSubject(S.lookup_wrap);
    BlastChooseNaExtend(S.lookup_wrap);
    info.constructor = &mdb_new; info.ctor_argument = &db;
    seq_src = BlastSeqSrcNew(&info);
    st = dump_params(&S, seq_src, res, (cfg->taps & 8) != 0);
    if (!st) {
        st = BLAST_GapAlignSetUp(prog, seq_src, S.score_options, S.eff_len_options, S.ext_options, S.hit_options,
                                 S.query_info, S.sbp, &score_params, &ext_params, &hit_params, &eff_len_params,
                                 &gap_align);
        if (st) st += 400;
    }
    if (!st) {
        gap_align->gap_x_dropoff = ext_params->gap_x_dropoff_final;      /* core/blast_traceback.c:1403 */
        tap.res = res; tap.taps = 16; tap.db = &db; tap.cur_chunk_off = 0;
        tap.tb_seq = NULL; tap.tb_oid = -1; tap.q_base = S.query->sequence; tap.qinfo = S.query_info;
        g_tap = &tap;
        for (i = 0; i < n_items && !st; i++) {
            const int32_t *it = items + 6 * i;
            const Int4 oid = it[0], ctx = it[1], s_shift = it[2], s_len = it[3], q_start = it[4], s_start = it[5];
            const Uint1 *query = S.query->sequence + S.query_info->contexts[ctx].query_offset;
            const Int4 q_len = S.query_info->contexts[ctx].query_length;
            if (oid != cur_oid) {
                const Int4 len = slen[oid];
                const uint8_t *pk = packed + sbyteoff[oid];
                Int4 k;
                free(buf);
                buf = (Uint1 *)malloc((size_t)len + 2);
                buf[0] = buf[len + 1] = 15;
                for (k = 0; k < len; k++) buf[k + 1] = (pk[k >> 2] >> (6 - 2 * (k & 3))) & 3;
                mdb_overlay_ambiguity(&db, oid, buf + 1, len);
                cur_oid = oid;
            }
            tap.tb_seq = buf + 1; tap.tb_oid = oid;
            if (S.ext_options->eTbackExt == eGreedyTbck)
                BLAST_GreedyGappedAlignment(query, buf + 1 + s_shift, q_len, s_len, gap_align, score_params,
                                            q_start, s_start, FALSE, TRUE, NULL);
            else
                BLAST_GappedAlignmentWithTraceback(prog, query, buf + 1 + s_shift, gap_align, score_params,
                                                   q_start, s_start, q_len, s_len, NULL);
            gap_align->edit_script = GapEditScriptDelete(gap_align->edit_script);
        }
        g_tap = NULL;
    }
    free(buf);
    BLAST_GapAlignStructFree(gap_align);
    BlastScoringParametersFree(score_params);
    BlastExtensionParametersFree(ext_params);
    BlastHitSavingParametersFree(hit_params);
    BlastEffectiveLengthsParametersFree(eff_len_params);
    BlastSeqSrcFree(seq_src);
    free_setup(&S);
    res->status = st;
    return st;
}

void ref_free_result(RefResult *res)
{
    free(res->scan.data); free(res->init.data); free(res->gapped.data); free(res->final_.data);
    free(res->tb_calls.data); free(res->tb_ops.data); free(res->tb_final.data); free(res->kept.data);
    free(res->ctx_query_offset); free(res->ctx_query_length); free(res->ctx_length_adjustment);
    free(res->ctx_eff_searchsp); free(res->ctx_x_dropoff); free(res->ctx_cutoff_score);
    free(res->ctx_reduced_cutoff); free(res->ctx_gapped_cutoff);
    free(res->ctx_kbp_std); free(res->ctx_kbp_gap);
    free(res->hashtable); free(res->next_pos); free(res->pv_array);
    free(res->backbone); free(res->overflow); free(res->masked_locations); free(res->na_backbone); free(res->na_overflow);
    free(res->concat_query);
    memset(res, 0, sizeof *res);
}
