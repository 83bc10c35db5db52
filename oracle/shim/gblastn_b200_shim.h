/* gblastn_b200_shim.h — entry points of the reference-side binding (oracle/shim/gblastn_b200_shim.c). */
#ifndef GBLASTN_B200_SHIM_H
#define GBLASTN_B200_SHIM_H
#include <stdint.h>
#include <algo/blast/core/blast_engine.h>
#include <algo/blast/core/blast_gapalign.h>
#include <algo/blast/core/na_ungapped.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The volume the calling thread's BlastSeqSrc hands sequences out of: `packed` is the host memory the subject
 * pointers point into (a mapped .nsq, or an in-memory volume), sequence i at packed + seq_byte_off[i].  The first
 * form uploads it (bn_db_load) and frees it on detach; the second binds an already resident volume. */
int  bnshim_attach_volume(const uint8_t *packed, int64_t packed_bytes, const int64_t *seq_byte_off,
                          const int32_t *seq_len, int32_t n_seq, int device);
int  bnshim_attach_resident_volume(int vol_handle, const uint8_t *host_base, const int64_t *seq_byte_off, int32_t n_seq);
void bnshim_detach_volume(void);

/* brackets of one preliminary search: begin = right after BLAST_GapAlignSetUp created the parameter blocks the seams
 * do not receive, end = after Blast_RunPreliminarySearchWithInterrupt returned */
void bnshim_prelim_begin(const BlastScoringParameters *score_params, const BlastExtensionParameters *ext_params,
                         const BlastHitSavingParameters *hit_params, const BlastGapAlignStruct *gap_align);
void bnshim_prelim_end(void);

/* BlastWordFinderType / BlastGetGappedScoreType (inc-core/blast_engine.h:212-238) */
Int2 bnshim_word_finder(BLAST_SequenceBlk *subject, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                        LookupTableWrap *lookup_wrap, Int4 **matrix, const BlastInitialWordParameters *word_params,
                        Blast_ExtendWord *ewp, BlastOffsetPair *offset_pairs, Int4 max_hits,
                        BlastInitHitList *init_hitlist, BlastUngappedStats *ungapped_stats);
Int2 bnshim_get_gapped_score(EBlastProgramType program_number, BLAST_SequenceBlk *query, BlastQueryInfo *query_info,
                             BLAST_SequenceBlk *subject, BlastGapAlignStruct *gap_align,
                             const BlastScoringParameters *score_params, const BlastExtensionParameters *ext_params,
                             const BlastHitSavingParameters *hit_params, BlastInitHitList *init_hitlist,
                             BlastHSPList **hsp_list_ptr, BlastGappedStats *gapped_stats, Boolean *fence_hit);
const char *bnshim_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
