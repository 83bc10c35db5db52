/* blastn_port.c — CPU restatement of the blastn preliminary-search hot path (TEST INFRASTRUCTURE).
 *
 * See blastn_port.h for the role of this file.  Reference paths: core/ = c++/src/algo/blast/core.
 * Each function cites the reference lines it restates.  Sequential, single-threaded, written for
 * clarity; it is the checker, never the thing measured as the product.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include "blastn_port.h"

#define PMIN(a, b) ((a) < (b) ? (a) : (b))
#define PMAX(a, b) ((a) > (b) ? (a) : (b))
#define MININT (INT_MIN / 2)

/* ------------------------------------------------------------------ small utilities */
typedef struct Vec { void *p; int64_t n, cap; size_t el; } Vec;
static void vec_init(Vec *v, size_t el) { v->p = NULL; v->n = v->cap = 0; v->el = el; }
static void *vec_push(Vec *v)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? 2 * v->cap : 256;
        v->p = realloc(v->p, (size_t)v->cap * v->el);
    }
    return (char *)v->p + (size_t)(v->n++) * v->el;
}

typedef struct Subject {
    const uint8_t *seq;   /* packed, chunk start (byte aligned) */
    int32_t len;          /* chunk length in bases */
    int32_t oid, chunk_off;
    /* subject->seq_ranges of a masked database (unmasked [left, right) intervals, chunk-relative);
     * masked == 0: one implicit range covering the chunk, scanned from offset 0 */
    int32_t masked, n_ranges;
    const int32_t *ranges;
} Subject;

/* NCBI2NA_UNPACK_BASE, inc-core/blast_util.h:52-55 */
static inline int sbase(const uint8_t *s, int32_t pos)
{
    return (s[pos >> 2] >> (6 - 2 * (pos & 3))) & 3;
}

/* BSearchContextInfo, core/blast_query_info.c:220-236 */
static int32_t ctx_search(const BnQueryBatch *b, int32_t n)
{
    int32_t lo = 0, hi = b->num_contexts, m;
    while (lo < hi - 1) {
        m = (lo + hi) / 2;
        if (b->contexts[m].query_offset > n) hi = m; else lo = m;
    }
    return lo;
}

/* ------------------------------------------------------------------ lookup probes */
/* s_MBLookup core/na_ungapped.c:52-75, s_SmallNaLookup :83-106 */
static int lut_contains(const BnQueryBatch *b, uint32_t index, int32_t q_pos)
{
    if (b->lut_type == BN_LUT_MB) {
        int32_t q;
        index &= (uint32_t)(b->hashsize - 1);
        q = b->hashtable[index];
        ++q_pos;
        while (q) {
            if (q == q_pos) return 1;
            q = b->next_pos[q];
        }
        return 0;
    } else if (b->lut_type == BN_LUT_NA) {
        /* s_NaLookup core/na_ungapped.c:112-138 */
        const int32_t *cell = b->na_backbone + 4 * (size_t)(index & (uint32_t)(b->hashsize - 1));
        const int32_t n = cell[0];
        const int32_t *pos = (n <= 3) ? cell + 1 : b->na_overflow + cell[1];
        int32_t i;
        for (i = 0; i < n; i++) if (pos[i] == q_pos) return 1;
        return 0;
    } else {
        int32_t v = b->backbone[index & (uint32_t)(b->hashsize - 1)], src;
        if (v == q_pos) return 1;
        if (v == -1 || v >= 0) return 0;
        src = -v;
        v = b->overflow[src++];
        do {
            if (v == q_pos) return 1;
            v = b->overflow[src++];
        } while (v >= 0);
        return 0;
    }
}

/* s_IsSeedMasked core/na_ungapped.c:459-471 */
static int seed_masked(const BnQueryBatch *b, const Subject *S, int32_t s_off, int32_t lut,
                       int32_t q_pos)
{
    const uint8_t *s = S->seq + s_off / 4;
    int shift = 2 * (16 - s_off % 4 - lut);
    uint32_t w = ((uint32_t)s[0] << 24) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 8) | s[3];
    return !lut_contains(b, w >> shift, q_pos);
}

/* ------------------------------------------------------------------ ungapped extension */
typedef struct Ungapped { int32_t q_start, s_start, length, score; } Ungapped;

/* s_NuclUngappedExtendExact core/na_ungapped.c:153-245 */
static void ungapped_exact(const BnQueryBatch *b, const Subject *S, int32_t q_off, int32_t s_off,
                           int32_t X, Ungapped *u)
{
    const uint8_t *query = b->query_start + 1;
    const int32_t *M = b->matrix;
    int32_t sum = 0, score = 0;
    int32_t q = q_off, q_beg = q_off, q_end = q_off;
    int32_t sp, s_lo, s_hi;
    int32_t q_avail = b->concat_len - q_off, s_avail = S->len - s_off;

    s_lo = (q_off < s_off) ? s_off - q_off : 0;  /* first subject base reachable on the left */
    sp = s_off;
    while (sp > s_lo) {
        sp--; q--;
        sum += M[16 * query[q] + sbase(S->seq, sp)];
        if (sum > 0) { q_beg = q; score += sum; sum = 0; }
        else if (sum < X) break;
    }
    u->q_start = q_beg;
    u->s_start = s_off - (q_off - q_beg);

    s_hi = (q_avail < s_avail) ? s_off + q_avail : S->len;
    q = q_off; sp = s_off; sum = 0;
    while (sp < s_hi) {
        sum += M[16 * query[q] + sbase(S->seq, sp)];
        q++; sp++;
        if (sum > 0) { q_end = q; score += sum; sum = 0; }
        else if (sum < X) break;
    }
    u->length = q_end - q_beg;
    u->score = score;
}

/* s_NuclUngappedExtend core/na_ungapped.c:263-350 */
static void ungapped_extend(const BnQueryBatch *b, const Subject *S, int32_t q_off,
                            int32_t s_match_end, int32_t s_off, int32_t X, Ungapped *u,
                            int32_t reduced_cutoff)
{
    const uint8_t *query = b->query_start + 1;
    const int32_t *tab = b->nucl_score_table;
    int32_t len = (4 - (s_off % 4)) % 4;
    int32_t q_ext = q_off + len, s_ext = s_off + len;
    int32_t q = q_ext, sb = s_ext / 4;
    int32_t sum = 0, score = 0, new_q = q_ext, i;

    len = PMIN(q_ext, s_ext) / 4;
    for (i = 0; i < len; sb--, q -= 4, i++) {
        uint8_t s_byte = S->seq[sb - 1];
        uint8_t q_byte = (uint8_t)((query[q - 4] << 6) | (query[q - 3] << 4) |
                                   (query[q - 2] << 2) | query[q - 1]);
        sum += tab[q_byte ^ s_byte];
        if (sum > 0) { new_q = q - 4; score += sum; sum = 0; }
        if (sum < X) break;
    }
    u->q_start = new_q;
    u->s_start = s_ext - (q_ext - new_q);

    q = q_ext; sb = s_ext / 4;
    len = PMIN(b->concat_len - q_ext, S->len - s_ext) / 4;
    sum = 0; new_q = q;
    for (i = 0; i < len; sb++, q += 4, i++) {
        uint8_t s_byte = S->seq[sb];
        uint8_t q_byte = (uint8_t)((query[q] << 6) | (query[q + 1] << 4) |
                                   (query[q + 2] << 2) | query[q + 3]);
        sum += tab[q_byte ^ s_byte];
        if (sum > 0) { new_q = q + 3; score += sum; sum = 0; }
        if (sum < X) break;
    }
    if (score >= reduced_cutoff) {
        ungapped_exact(b, S, q_off, s_off, X, u);
    } else {
        u->score = score;
        u->length = PMAX(s_match_end - u->s_start, new_q - u->q_start + 1);
    }
}

/* ------------------------------------------------------------------ s_TypeOfWord */
/* core/na_ungapped.c:489-588.  Returns 0 / 1 / 2; may advance q_off,s_off; sets *extended. */
static int type_of_word(const BnQueryBatch *b, const Subject *S, int32_t *q_off, int32_t *s_off,
                        int has_locations, uint32_t s_range, uint32_t word_length,
                        uint32_t lut_word_length, int check_double, int32_t *extended)
{
    int32_t context, q_range, ext_to, ext_max;
    int32_t q_end = *q_off + (int32_t)word_length, s_end = *s_off + (int32_t)word_length;
    int32_t s_pos, q_pos;
    const int32_t lut = (int32_t)lut_word_length;

    *extended = 0;
    if (word_length == lut_word_length) return 1;

    context = ctx_search(b, q_end);
    q_range = b->contexts[context].query_offset + b->contexts[context].query_length;

    if (has_locations) {
        if (seed_masked(b, S, s_end - lut, lut, q_end - lut)) return 0;
        for (;; ++(*s_off), ++(*q_off))
            if (!seed_masked(b, S, *s_off, lut, *q_off)) break;
    }
    ext_to = (int32_t)word_length - (q_end - *q_off);
    {   /* MIN() of an Int4 and a Uint4 compares as unsigned in the reference (:534) */
        uint32_t a = (uint32_t)(q_range - q_end), c = s_range - (uint32_t)s_end;
        ext_max = (int32_t)(a > c ? c : a);
    }

    if (ext_to || has_locations) {
        if (ext_to > ext_max) return 0;
        q_end += ext_to; s_end += ext_to;
        for (s_pos = s_end - lut, q_pos = q_end - lut; s_pos > *s_off; s_pos -= lut, q_pos -= lut)
            if (seed_masked(b, S, s_pos, lut, q_pos)) return 0;
        *extended = ext_to;
    }
    if (!check_double) return 1;

    ext_to += (int32_t)word_length;
    ext_max = PMIN(ext_max, ext_to);
    for (s_pos = s_end, q_pos = q_end; (uint32_t)*extended + lut_word_length <= (uint32_t)ext_max;
         s_pos += lut, q_pos += lut, *extended += lut)
        if (seed_masked(b, S, s_pos, lut, q_pos)) break;
    s_pos -= (lut - 1); q_pos -= (lut - 1);
    while (*extended < ext_max) {
        if (seed_masked(b, S, s_pos, lut, q_pos)) return 1;
        (*extended)++; ++s_pos; ++q_pos;
    }
    return (ext_max == ext_to) ? 2 : 1;
}

/* ------------------------------------------------------------------ diagonal containers */
typedef struct HashCell { int32_t diag, level, hit_len; uint32_t hit_saved, next; } HashCell;
typedef struct DiagState {
    int is_hash;
    /* hash: BLAST_DiagHash inc-core/blast_extend.h:98; constants :51,:54 */
    uint32_t backbone[512];
    HashCell *chain; uint32_t occupancy, capacity;
    /* array: BLAST_DiagTable :77 */
    int32_t *last_hit; uint8_t *flag; uint8_t *hit_len_arr;
    int32_t diag_array_length, diag_mask;
    int32_t offset, window;
} DiagState;

/* BlastExtendWordNew core/blast_extend.c:117-162, s_BlastDiagTableNew :46-72 */
static void diag_new(DiagState *d, const BnQueryBatch *b)
{
    memset(d, 0, sizeof *d);
    d->is_hash = (b->container_type == BN_DIAG_HASH);
    d->window = b->window_size;
    d->offset = b->window_size;
    if (d->is_hash) {
        d->capacity = 256; d->occupancy = 1;
        d->chain = (HashCell *)calloc(d->capacity, sizeof(HashCell));
    } else {
        int32_t n = 1;
        while (n < b->concat_len + b->window_size) n <<= 1;
        d->diag_array_length = n; d->diag_mask = n - 1;
        d->last_hit = (int32_t *)calloc((size_t)n, 4);
        d->flag = (uint8_t *)calloc((size_t)n, 1);
        d->hit_len_arr = (uint8_t *)calloc((size_t)n, 1);
    }
}
static void diag_free(DiagState *d) { free(d->chain); free(d->last_hit); free(d->flag); free(d->hit_len_arr); }

/* Blast_ExtendWordExit core/blast_extend.c:164-186 (+ s_BlastDiagClear :85-105) */
static void diag_exit(DiagState *d, int32_t subject_length)
{
    if (d->offset >= INT_MAX / 4) {
        d->offset = d->window;
        if (d->is_hash) { d->occupancy = 1; memset(d->backbone, 0, sizeof d->backbone); }
        else {
            int32_t i;
            for (i = 0; i < d->diag_array_length; i++) {
                d->flag[i] = 0; d->last_hit[i] = -d->window; d->hit_len_arr[i] = 0;
            }
        }
    } else d->offset += subject_length + d->window;
}

/* s_BlastDiagHashRetrieve core/na_ungapped.c:361-380 */
static int hash_get(const DiagState *d, int32_t diag, int32_t *level, int32_t *hit_len, int32_t *saved)
{
    uint32_t idx = d->backbone[((uint32_t)diag * 0x9E370001u) % 512u];
    while (idx) {
        if (d->chain[idx].diag == diag) {
            *level = d->chain[idx].level; *hit_len = d->chain[idx].hit_len;
            *saved = (int32_t)d->chain[idx].hit_saved;
            return 1;
        }
        idx = d->chain[idx].next;
    }
    return 0;
}
/* s_BlastDiagHashInsert core/na_ungapped.c:395-450 */
static void hash_put(DiagState *d, int32_t diag, int32_t level, int32_t len, int32_t saved,
                     int32_t s_off, int32_t window_size)
{
    uint32_t bucket = ((uint32_t)diag * 0x9E370001u) % 512u;
    uint32_t idx = d->backbone[bucket];
    HashCell *c;
    while (idx) {
        c = &d->chain[idx];
        if (c->diag == diag || s_off - c->level > window_size) {
            c->diag = diag; c->level = level; c->hit_len = len; c->hit_saved = (uint32_t)saved;
            return;
        }
        idx = c->next;
    }
    if (d->occupancy == d->capacity) {
        d->capacity *= 2;
        d->chain = (HashCell *)realloc(d->chain, d->capacity * sizeof(HashCell));
    }
    c = &d->chain[d->occupancy];
    c->diag = diag; c->level = level; c->hit_len = len; c->hit_saved = (uint32_t)saved;
    c->next = d->backbone[bucket];
    d->backbone[bucket] = d->occupancy++;
}

/* ------------------------------------------------------------------ one word hit */
typedef struct WordCtx {
    const BnQueryBatch *b;
    const Subject *S;
    DiagState *diag;
    Vec *init;       /* BnInitHit */
    int64_t n_extended;
} WordCtx;

static void save_init(WordCtx *w, int32_t q_off, int32_t s_off, const Ungapped *u)
{
    BnInitHit *h = (BnInitHit *)vec_push(w->init);
    h->oid = w->S->oid; h->chunk_off = w->S->chunk_off;
    h->q_off = q_off; h->s_off = s_off;
    h->q_start = u->q_start; h->s_start = u->s_start; h->length = u->length; h->score = u->score;
}

/* s_BlastnDiagHashExtendInitialHit core/na_ungapped.c:779-922 and
 * s_BlastnDiagTableExtendInitialHit :612-757 (they differ in the container and in the
 * `word_length < 11` exact-extension branch, :720-721). */
static int diag_extend_hit(WordCtx *w, int32_t q_off, int32_t s_off, int has_locations,
                           int32_t s_range, int32_t word_length, int32_t lut_word_length)
{
    const BnQueryBatch *b = w->b;
    DiagState *d = w->diag;
    const int32_t window_size = b->window_size;
    const int two_hits = window_size > 0;
    int32_t Delta = PMIN(b->scan_range, window_size - word_length);
    int32_t diag, real_diag = 0, last_hit, hit_saved = 0, s_l = 0;
    int32_t s_end = s_off + word_length;
    int32_t s_off_pos = s_off + d->offset, s_end_pos = s_end + d->offset;
    int32_t extended = 0, word_type, hit_ready = 1, off_found = 0;
    Ungapped u;

    if (d->is_hash) {
        diag = s_off - q_off;
        if (!hash_get(d, diag, &last_hit, &s_l, &hit_saved)) last_hit = 0;
    } else {
        diag = s_off + d->diag_array_length - q_off;
        real_diag = diag & d->diag_mask;
        last_hit = d->last_hit[real_diag];
        hit_saved = d->flag[real_diag];
    }
    if (s_off_pos < last_hit) return 0;

    if (two_hits && (hit_saved || s_end_pos > last_hit + window_size)) {
        word_type = type_of_word(b, w->S, &q_off, &s_off, has_locations, (uint32_t)s_range,
                                 (uint32_t)word_length, (uint32_t)lut_word_length, 1, &extended);
        if (!word_type) return 0;
        s_end += extended; s_end_pos += extended;
        if (word_type == 1) {
            int32_t s_a = s_off_pos + word_length - window_size;
            int32_t s_b = s_end_pos - 2 * word_length;
            int32_t delta;
            if (Delta < 0) Delta = 0;
            for (delta = 1; delta <= Delta; ++delta) {
                int32_t e = 0, l = 0, sv = 0;
                if (d->is_hash) {
                    if (hash_get(d, diag + delta, &e, &l, &sv) && l && e - delta >= s_a && e - l <= s_b) { off_found = 1; break; }
                    e = l = 0;
                    if (hash_get(d, diag - delta, &e, &l, &sv) && l && e >= s_a && e - l + delta <= s_b) { off_found = 1; break; }
                } else {
                    int32_t orig = real_diag + d->diag_array_length;
                    int32_t od = (orig + delta) & d->diag_mask;
                    e = d->last_hit[od]; l = d->hit_len_arr[od];
                    if (l && e - delta >= s_a && e - l <= s_b) { off_found = 1; break; }
                    od = (orig - delta) & d->diag_mask;
                    e = d->last_hit[od]; l = d->hit_len_arr[od];
                    if (l && e >= s_a && e - l + delta <= s_b) { off_found = 1; break; }
                }
            }
            if (!off_found) hit_ready = 0;
        }
    } else {
        if (!type_of_word(b, w->S, &q_off, &s_off, has_locations, (uint32_t)s_range,
                          (uint32_t)word_length, (uint32_t)lut_word_length, 0, &extended))
            return 0;
        s_end += extended; s_end_pos += extended;
    }

    if (hit_ready) {
        int32_t context = ctx_search(b, q_off);
        const BnContext *c = &b->contexts[context];
        if (!d->is_hash && word_length < 11)
            ungapped_exact(b, w->S, q_off, s_off, -c->x_dropoff, &u);
        else
            ungapped_extend(b, w->S, q_off, s_end, s_off, -c->x_dropoff, &u, c->reduced_cutoff);
        if (off_found || u.score >= c->cutoff_score) {
            save_init(w, q_off, s_off, &u);
            s_end_pos = u.length + u.s_start + d->offset;
        } else hit_ready = 0;
    }
    if (d->is_hash) {
        hash_put(d, diag, s_end_pos, hit_ready ? 0 : s_end_pos - s_off_pos, hit_ready, s_off_pos,
                 window_size + Delta + 1);
    } else {
        d->last_hit[real_diag] = s_end_pos;
        d->flag[real_diag] = (uint8_t)hit_ready;
        if (two_hits) d->hit_len_arr[real_diag] = (uint8_t)(hit_ready ? 0 : s_end_pos - s_off_pos);
    }
    return hit_ready;
}

/* ------------------------------------------------------------------ mini-extension */
/* s_BlastNaExtend core/na_ungapped.c:1026-1148 (s_BlastNaExtendAligned :1166-1290 is the same
 * function specialised for 4-aligned hits); s_BlastNaExtendDirect :942-1005. */
static void extend_mb_hit(WordCtx *w, int32_t q_offset, int32_t s_offset, int32_t s_range)
{
    const BnQueryBatch *b = w->b;
    const uint8_t *query = b->query_start + 1;
    const int32_t word = b->word_length, lut = b->lut_word_length, ext_to = word - lut;
    const int has_loc = b->masked_locations != NULL;
    int32_t ext_left = 0, s_off = s_offset, q = q_offset, lim;

    if (ext_to == 0) {
        w->n_extended += diag_extend_hit(w, q_offset, s_offset, 0, s_range, word, word);
        return;
    }
    lim = PMIN(ext_to, s_offset);
    for (; ext_left < lim; ++ext_left) {
        s_off--; q--;
        if (sbase(w->S->seq, s_off) != query[q]) break;
    }
    if (ext_left < ext_to) {
        int32_t ext_right = 0;
        s_off = s_offset + lut;
        if (s_off + ext_to - ext_left > s_range) return;
        q = q_offset + lut;
        for (; ext_right < ext_to - ext_left; ++ext_right) {
            if (sbase(w->S->seq, s_off) != query[q]) break;
            s_off++; q++;
        }
        if (ext_left + ext_right < ext_to) return;
    }
    w->n_extended += diag_extend_hit(w, q_offset - ext_left, s_offset - ext_left, has_loc, s_range,
                                     word, lut);
}

/* compressed_nuc_seq[i]: bases i..i+3 packed, built by BlastCompressBlastnaSequence
 * core/blast_util.c:459-501 (values & 3, zero padding beyond both ends). */
static inline uint8_t cq(const BnQueryBatch *b, int32_t i)
{
    const uint8_t *query = b->query_start + 1;
    uint8_t v = 0;
    int k;
    for (k = 0; k < 4; k++) {
        int32_t p = i + k;
        v = (uint8_t)(v << 2);
        if (p >= 0 && p < b->concat_len) v |= (query[p] & 3);
    }
    return v;
}
static int match_left(uint8_t x)  { int n = 0; while (n < 4 && ((x >> (2 * n)) & 3) == 0) n++; return n; }
static int match_right(uint8_t x) { int n = 0; while (n < 4 && ((x >> (6 - 2 * n)) & 3) == 0) n++; return n; }

/* s_BlastSmallNaExtendAlignedOneByte core/na_ungapped.c:1347-1427,
 * s_BlastSmallNaExtend :1450-1555 */
static void extend_small_hit(WordCtx *w, int32_t q_offset, int32_t s_offset, int32_t s_range)
{
    const BnQueryBatch *b = w->b;
    const int32_t word = b->word_length, lut = b->lut_word_length, ext_to = word - lut;
    const int has_loc = b->masked_locations != NULL;
    const uint8_t *s = w->S->seq;
    int32_t context, q_start, q_range, ext_left = 0, ext_right = 0;

    if (ext_to == 0) {
        w->n_extended += diag_extend_hit(w, q_offset, s_offset, 0, s_range, word, word);
        return;
    }
    context = ctx_search(b, q_offset);
    q_start = b->contexts[context].query_offset;
    q_range = q_start + b->contexts[context].query_length;

    if (lut % 4 == 0 && b->scan_step % 4 == 0 && ext_to <= 4) {
        if (s_offset > 0 && q_offset > 0) {
            ext_left = match_left(cq(b, q_offset - 4) ^ s[s_offset / 4 - 1]);
            ext_left = PMIN(PMIN(ext_left, ext_to), q_offset - q_start);
        }
        if (ext_left < ext_to && (q_offset + lut) < b->concat_len) {
            ext_right = match_right(cq(b, q_offset + lut) ^ s[(s_offset + lut) / 4]);
            ext_right = PMIN(PMIN(ext_right, s_range - (s_offset + lut)), q_range - (q_offset + lut));
            if (ext_left + ext_right < ext_to) return;
        }
    } else {
        int32_t ext_max = PMIN(PMIN(ext_to, s_offset), q_offset - q_start);
        int32_t rsdl = 4 - (s_offset % 4), s_off, q_off;
        s_offset += rsdl; q_offset += rsdl; ext_max += rsdl;
        s_off = s_offset; q_off = q_offset;
        while (ext_left < ext_max) {
            int bases = match_left(cq(b, q_off - 4) ^ s[s_off / 4 - 1]);
            ext_left += bases;
            if (bases < 4) break;
            q_off -= 4; s_off -= 4;
        }
        ext_left = PMIN(ext_left, ext_max);
        s_off = s_offset; q_off = q_offset;
        ext_max = PMIN(PMIN(word - ext_left, s_range - s_off), q_range - q_off);
        while (ext_right < ext_max) {
            int bases = match_right(cq(b, q_off) ^ s[s_off / 4]);
            ext_right += bases;
            if (bases < 4) break;
            q_off += 4; s_off += 4;
        }
        ext_right = PMIN(ext_right, ext_max);
        if (ext_left + ext_right < word) return;
    }
    w->n_extended += diag_extend_hit(w, q_offset - ext_left, s_offset - ext_left, has_loc, s_range,
                                     word, lut);
}

/* ------------------------------------------------------------------ word finder */
/* BlastNaWordFinder core/na_ungapped.c:1559-1657 for an unmasked subject: scan positions
 * 0, step, 2*step ... <= len - lut (scanners core/blast_nascan.c:1489-1591, 445-560; chain
 * expansion :1413-1427 / :312-335), mini-extension, diagonal logic, then
 * Blast_InitHitListSortByScore (core/blast_extend.c:274-310; glibc qsort is a stable merge sort,
 * so full ties keep emission order). */
static int init_cmp(const void *a, const void *c)
{
    const BnInitHit *h1 = (const BnInitHit *)a, *h2 = (const BnInitHit *)c;
    if (h1->score != h2->score) return h2->score > h1->score ? 1 : -1;
    if (h1->s_start != h2->s_start) return h1->s_start > h2->s_start ? 1 : -1;
    if (h1->length != h2->length) return h2->length > h1->length ? 1 : -1;
    if (h1->q_start != h2->q_start) return h1->q_start > h2->q_start ? 1 : -1;
    return 0;
}
static void stable_sort(void *base, size_t n, size_t el, int (*cmp)(const void *, const void *))
{
    /* bottom-up merge sort: stable, like glibc's qsort when memory allows */
    char *a = (char *)base, *t;
    size_t w, i;
    if (n < 2) return;
    t = (char *)malloc(n * el);
    for (w = 1; w < n; w *= 2) {
        for (i = 0; i < n; i += 2 * w) {
            size_t l = i, m = PMIN(i + w, n), r = PMIN(i + 2 * w, n), p = l, q = m, o = l;
            while (p < m && q < r) {
                if (cmp(a + q * el, a + p * el) < 0) memcpy(t + (o++) * el, a + (q++) * el, el);
                else memcpy(t + (o++) * el, a + (p++) * el, el);
            }
            while (p < m) memcpy(t + (o++) * el, a + (p++) * el, el);
            while (q < r) memcpy(t + (o++) * el, a + (q++) * el, el);
        }
        memcpy(a, t, n * el);
    }
    free(t);
}

static void word_finder(const BnQueryBatch *b, const Subject *S, DiagState *diag, Vec *init,
                        int64_t first_init, PortResults *out, int taps, Vec *scanv, Vec *scan_oid,
                        Vec *scan_chunk)
{
    WordCtx w;
    const int32_t lut = b->lut_word_length, step = b->scan_step;
    int32_t p, ri;
    const int32_t whole[2] = {0, S->len};
    const int32_t nr = S->masked ? S->n_ranges : 1;
    w.b = b; w.S = S; w.diag = diag; w.init = init; w.n_extended = 0;

    /* BlastNaWordFinder core/na_ungapped.c:1610-1645 + s_DetermineScanningOffsets core/masksubj.inl:43-59:
     * every unmasked range is scanned from left + (word - lut) (0 for an unmasked subject) to right - lut,
     * and its right end bounds the mini-extension and s_TypeOfWord */
    for (ri = 0; ri < nr; ri++) {
    const int32_t *rg = S->masked ? S->ranges + 2 * ri : whole;
    const int32_t first = S->masked ? rg[0] + b->word_length - lut : 0;
    const int32_t last = rg[1] - lut, s_range = rg[1];
    for (p = first; p <= last; p += step) {
        const uint8_t *s = S->seq + p / 4;
        uint32_t word = ((uint32_t)s[0] << 24) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 8) | s[3];
        uint32_t idx = (word >> (2 * (16 - (p % 4 + lut)))) & (uint32_t)(b->hashsize - 1);
        if (b->lut_type == BN_LUT_MB) {
            int32_t q;
            if (b->pv_array &&
                !(b->pv_array[idx >> b->pv_array_bts] & (1u << (idx & 31)))) continue;
            q = b->hashtable[idx];
            while (q) {
                out->stats.lookup_hits++;
                if (taps & PORT_TAP_SCAN) {
                    BnOffsetPair *o = (BnOffsetPair *)vec_push(scanv);
                    o->q_off = (uint32_t)(q - 1); o->s_off = (uint32_t)p;
                    *(int32_t *)vec_push(scan_oid) = S->oid;
                    *(int32_t *)vec_push(scan_chunk) = S->chunk_off;
                }
                extend_mb_hit(&w, q - 1, p, s_range);
                q = b->next_pos[q];
            }
        } else if (b->lut_type == BN_LUT_NA) {
            /* s_BlastNaScanSubject_8_4 / _Any + s_BlastLookupRetrieve core/blast_nascan.c:63-290; the
             * extension is s_BlastNaExtend(Direct/Aligned) as for the megablast table (BlastChooseNaExtend
             * core/na_ungapped.c:1780-1792) */
            const int32_t *cell = b->na_backbone + 4 * (size_t)idx;
            const int32_t n = cell[0];
            const int32_t *pos = (n <= 3) ? cell + 1 : b->na_overflow + cell[1];
            int32_t i;
            for (i = 0; i < n; i++) {
                out->stats.lookup_hits++;
                if (taps & PORT_TAP_SCAN) {
                    BnOffsetPair *o = (BnOffsetPair *)vec_push(scanv);
                    o->q_off = (uint32_t)pos[i]; o->s_off = (uint32_t)p;
                    *(int32_t *)vec_push(scan_oid) = S->oid;
                    *(int32_t *)vec_push(scan_chunk) = S->chunk_off;
                }
                extend_mb_hit(&w, pos[i], p, s_range);
            }
        } else {
            int32_t v = b->backbone[idx];
            if (v == -1) continue;
            if (v >= 0) {
                out->stats.lookup_hits++;
                if (taps & PORT_TAP_SCAN) {
                    BnOffsetPair *o = (BnOffsetPair *)vec_push(scanv);
                    o->q_off = (uint32_t)v; o->s_off = (uint32_t)p;
                    *(int32_t *)vec_push(scan_oid) = S->oid;
                    *(int32_t *)vec_push(scan_chunk) = S->chunk_off;
                }
                extend_small_hit(&w, v, p, s_range);
            } else {
                int32_t src = -v;
                v = b->overflow[src++];
                do {
                    out->stats.lookup_hits++;
                    if (taps & PORT_TAP_SCAN) {
                        BnOffsetPair *o = (BnOffsetPair *)vec_push(scanv);
                        o->q_off = (uint32_t)v; o->s_off = (uint32_t)p;
                        *(int32_t *)vec_push(scan_oid) = S->oid;
                        *(int32_t *)vec_push(scan_chunk) = S->chunk_off;
                    }
                    extend_small_hit(&w, v, p, s_range);
                    v = b->overflow[src++];
                } while (v >= 0);
            }
        }
    }
    }
    diag_exit(diag, S->len);
    out->stats.init_extends += w.n_extended;
    out->stats.good_init_extends += init->n - first_init;
    stable_sort((BnInitHit *)init->p + first_init, (size_t)(init->n - first_init),
                sizeof(BnInitHit), init_cmp);
}

/* ------------------------------------------------------------------ greedy aligner */
/* s_FindFirstMismatch core/greedy_align.c:318-381 (compressed seq2 only; rem < 4) */
static int32_t first_mismatch(const uint8_t *seq1, const uint8_t *seq2, int32_t len1, int32_t len2,
                              int32_t i1, int32_t i2, int reverse, int rem)
{
    int32_t start = i1;
    if (reverse) {
        while (i1 < len1 && i2 < len2 &&
               seq1[len1 - 1 - i1] == sbase(seq2, len2 - 1 - i2)) { ++i1; ++i2; }
    } else {
        while (i1 < len1 && i2 < len2 &&
               seq1[i1] == sbase(seq2, i2 + rem)) { ++i1; ++i2; }
    }
    return i1 - start;
}

typedef struct GreedyMem { int32_t *row[2]; int32_t *max_score; int32_t max_d; } GreedyMem;
typedef struct GreedySeed { int32_t start_q, start_s, match_length; } GreedySeed;
#define GREEDY_MAX_COST 10000
#define GREEDY_INVALID (-2)    /* kInvalidOffset, core/greedy_align.c */

/* BLAST_GreedyAlign core/greedy_align.c:385-681 (score only: edit_block == NULL) */
static int32_t greedy_align(const uint8_t *seq1, int32_t len1, const uint8_t *seq2, int32_t len2,
                            int reverse, int32_t xdrop_threshold, int32_t match_cost,
                            int32_t mismatch_cost, int32_t *seq1_len, int32_t *seq2_len,
                            GreedyMem *mem, int rem, GreedySeed *seed)
{
    int32_t seq1_index, seq2_index, index, d, k, diag_lower, diag_upper, max_dist, diag_origin;
    int32_t best_dist = 0, best_diag = 0, xdrop_offset, longest_match_run;
    int32_t *max_score;
    int32_t *row[3];
    int end1_reached = 0, end2_reached = 0;

    max_dist = PMIN(GREEDY_MAX_COST, len2 / 2 + 1);
    diag_origin = max_dist + 2;
    xdrop_offset = (xdrop_threshold + match_cost / 2) / (match_cost + mismatch_cost) + 1;

    index = first_mismatch(seq1, seq2, len1, len2, 0, 0, reverse, rem);
    *seq1_len = index; *seq2_len = index;
    seq1_index = index;
    seed->start_q = 0; seed->start_s = 0;
    seed->match_length = longest_match_run = index;
    if (index == len1 || index == len2) return 0;

    max_score = mem->max_score + xdrop_offset;
    for (index = 0; index < xdrop_offset; index++) mem->max_score[index] = 0;

    /* rolling rows: last_seq2_off[d+1] = last_seq2_off[d-1] (:661) */
    row[0] = mem->row[0]; row[1] = mem->row[1];
    row[0][diag_origin] = seq1_index;
    max_score[0] = seq1_index * match_cost;
    diag_lower = diag_origin - 1;
    diag_upper = diag_origin + 1;

    for (d = 1; d <= max_dist; d++) {
        int32_t xdrop_score, curr_score, curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        int32_t tmp_lower = diag_lower, tmp_upper = diag_upper;
        int32_t *prev = row[(d - 1) & 1], *cur = row[d & 1];

        prev[diag_lower - 1] = GREEDY_INVALID;
        prev[diag_lower] = GREEDY_INVALID;
        prev[diag_upper] = GREEDY_INVALID;
        prev[diag_upper + 1] = GREEDY_INVALID;

        xdrop_score = max_score[d - xdrop_offset] + (match_cost + mismatch_cost) * d - xdrop_threshold;
        xdrop_score = (int32_t)ceil((double)xdrop_score / (match_cost / 2));

        for (k = tmp_lower; k <= tmp_upper; k++) {
            seq2_index = PMAX(prev[k + 1], prev[k]) + 1;
            seq2_index = PMAX(seq2_index, prev[k - 1]);
            seq1_index = seq2_index + k - diag_origin;
            if (seq2_index < 0 || seq1_index + seq2_index < xdrop_score) {
                if (k == diag_lower) diag_lower++;
                else cur[k] = GREEDY_INVALID;
                continue;
            }
            diag_upper = k;
            index = first_mismatch(seq1, seq2, len1, len2, seq1_index, seq2_index, reverse, rem);
            if (index > longest_match_run) {
                seed->start_q = seq1_index; seed->start_s = seq2_index;
                seed->match_length = longest_match_run = index;
            }
            seq1_index += index; seq2_index += index;
            cur[k] = seq2_index;
            if (seq1_index + seq2_index > curr_extent) {
                curr_extent = seq1_index + seq2_index;
                curr_seq2_index = seq2_index;
                curr_diag = k;
            }
            if (seq2_index == len2) { diag_lower = k + 1; end2_reached = 1; }
            if (seq1_index == len1) { diag_upper = k - 1; end1_reached = 1; }
        }
        curr_score = curr_extent * (match_cost / 2) - d * (match_cost + mismatch_cost);
        if (curr_score > max_score[d - 1]) {
            max_score[d] = curr_score;
            best_dist = d; best_diag = curr_diag;
            *seq2_len = curr_seq2_index;
            *seq1_len = curr_seq2_index + best_diag - diag_origin;
        } else max_score[d] = max_score[d - 1];
        if (diag_lower > diag_upper) break;
        if (!end2_reached) diag_lower--;
        if (!end1_reached) diag_upper++;
    }
    return best_dist;
}

/* BLAST_Gcd / BLAST_Gdb3 core/ncbi_math.c:405-440 */
static int32_t gcd_i(int32_t a, int32_t b)
{
    int32_t c;
    b = abs(b);
    if (b > a) { c = a; a = b; b = c; }
    while (b != 0) { c = a % b; a = b; b = c; }
    return a;
}
static int32_t gdb3(int32_t *a, int32_t *b, int32_t *c)
{
    int32_t g = (*b == 0) ? gcd_i(*a, *c) : gcd_i(*a, gcd_i(*b, *c));
    if (g > 1) { *a /= g; *b /= g; *c /= g; }
    return g;
}

/* BLAST_AffineGreedyAlign core/greedy_align.c:755-1237, affine body, score only (edit_block == NULL:
 * rows of last_seq2_off are recycled every max_penalty + 1 distances, :1165-1172).  Scores arrive
 * already doubled when the reward is odd (:792-798, done by the caller here).  Returns the SCORE
 * (max_score[best_dist], or index * match_score on the early exit), not a distance. */
typedef struct AffOff { int32_t insert_off, match_off, delete_off; } AffOff;
static int32_t greedy_align_affine(const uint8_t *seq1, int32_t len1, const uint8_t *seq2, int32_t len2,
                                   int reverse, int32_t xdrop_threshold, int32_t match_score,
                                   int32_t mismatch_score, int32_t in_gap_open, int32_t in_gap_extend,
                                   int32_t *seq1_len, int32_t *seq2_len, int rem, GreedySeed *seed)
{
    const int32_t kInvalidDiag = 100000000;
    int32_t seq1_index, seq2_index, index, d, k, max_dist, scaled_max_dist, diag_origin;
    int32_t best_dist = 0, best_diag = 0, longest_match_run, xdrop_offset;
    int32_t end1_diag = 0, end2_diag = 0, curr_diag_lower, curr_diag_upper, num_nonempty_dist;
    int32_t match_score_half = match_score / 2;
    int32_t op_cost = match_score + mismatch_score;
    int32_t gap_open = in_gap_open, gap_extend = in_gap_extend + match_score_half;
    int32_t score_common_factor = gdb3(&op_cost, &gap_open, &gap_extend);
    int32_t gap_open_extend = gap_open + gap_extend;
    int32_t max_penalty = PMAX(op_cost, gap_open_extend);
    int32_t width, nrows, result;
    AffOff *store, **rows;
    int32_t *max_score_base, *max_score, *bounds, *diag_lower, *diag_upper;

    max_dist = PMIN(GREEDY_MAX_COST, len2 / 2 + 1);
    scaled_max_dist = max_dist * gap_extend;
    diag_origin = max_dist + 2;
    xdrop_offset = (xdrop_threshold + match_score_half) / score_common_factor + 1;

    index = first_mismatch(seq1, seq2, len1, len2, 0, 0, reverse, rem);
    *seq1_len = index; *seq2_len = index;
    seq1_index = index;
    seed->start_q = 0; seed->start_s = 0;
    seed->match_length = longest_match_run = index;
    if (index == len1 || index == len2) return index * match_score;

    width = 2 * max_dist + 6;
    nrows = max_penalty + 1;
    store = (AffOff *)calloc((size_t)width * (size_t)nrows, sizeof(AffOff));
    rows = (AffOff **)malloc(((size_t)scaled_max_dist + 2) * sizeof(AffOff *));
    max_score_base = (int32_t *)calloc((size_t)scaled_max_dist + 2 + (size_t)xdrop_offset, 4);
    bounds = (int32_t *)calloc(2 * ((size_t)scaled_max_dist + 1 + (size_t)max_penalty), 4);
    for (index = 0; index <= max_penalty && index <= scaled_max_dist; index++) rows[index] = store + (size_t)index * width;

    max_score = max_score_base + xdrop_offset;
    diag_lower = bounds;
    diag_upper = bounds + scaled_max_dist + 1 + max_penalty;
    for (index = 0; index < max_penalty; index++) { diag_lower[index] = kInvalidDiag; diag_upper[index] = -kInvalidDiag; }
    diag_lower += max_penalty; diag_upper += max_penalty;

    rows[0][diag_origin].match_off = seq1_index;
    rows[0][diag_origin].insert_off = GREEDY_INVALID;
    rows[0][diag_origin].delete_off = GREEDY_INVALID;
    max_score[0] = seq1_index * match_score;
    diag_lower[0] = diag_origin; diag_upper[0] = diag_origin;
    curr_diag_lower = diag_origin - 1; curr_diag_upper = diag_origin + 1;
    num_nonempty_dist = 1;
    d = 1;

    while (d <= scaled_max_dist) {
        int32_t xdrop_score, curr_score, curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = curr_diag_lower, tmp_upper = curr_diag_upper;
        AffOff *cur = rows[d];

        xdrop_score = max_score[d - xdrop_offset] + score_common_factor * d - xdrop_threshold;
        xdrop_score = (int32_t)ceil((double)xdrop_score / match_score_half);
        if (xdrop_score < 0) xdrop_score = 0;

        for (k = tmp_lower; k <= tmp_upper; k++) {
            seq2_index = GREEDY_INVALID;
            if (k + 1 <= diag_upper[d - gap_open_extend] && k + 1 >= diag_lower[d - gap_open_extend])
                seq2_index = rows[d - gap_open_extend][k + 1].match_off;
            if (k + 1 <= diag_upper[d - gap_extend] && k + 1 >= diag_lower[d - gap_extend] &&
                seq2_index < rows[d - gap_extend][k + 1].delete_off)
                seq2_index = rows[d - gap_extend][k + 1].delete_off;
            cur[k].delete_off = (seq2_index == GREEDY_INVALID) ? GREEDY_INVALID : seq2_index + 1;

            seq2_index = GREEDY_INVALID;
            if (k - 1 <= diag_upper[d - gap_open_extend] && k - 1 >= diag_lower[d - gap_open_extend])
                seq2_index = rows[d - gap_open_extend][k - 1].match_off;
            if (k - 1 <= diag_upper[d - gap_extend] && k - 1 >= diag_lower[d - gap_extend] &&
                seq2_index < rows[d - gap_extend][k - 1].insert_off)
                seq2_index = rows[d - gap_extend][k - 1].insert_off;
            cur[k].insert_off = seq2_index;

            seq2_index = PMAX(cur[k].insert_off, cur[k].delete_off);
            if (k <= diag_upper[d - op_cost] && k >= diag_lower[d - op_cost])
                seq2_index = PMAX(seq2_index, rows[d - op_cost][k].match_off + 1);
            seq1_index = seq2_index + k - diag_origin;

            if (seq2_index < 0 || seq1_index + seq2_index < xdrop_score) {
                if (k == curr_diag_lower) curr_diag_lower++;
                else cur[k].match_off = GREEDY_INVALID;
                continue;
            }
            curr_diag_upper = k;
            index = first_mismatch(seq1, seq2, len1, len2, seq1_index, seq2_index, reverse, rem);
            if (index > longest_match_run) {
                seed->start_q = seq1_index; seed->start_s = seq2_index;
                seed->match_length = longest_match_run = index;
            }
            seq1_index += index; seq2_index += index;
            cur[k].match_off = seq2_index;
            if (seq1_index + seq2_index > curr_extent) {
                curr_extent = seq1_index + seq2_index;
                curr_seq2_index = seq2_index;
                curr_diag = k;
            }
            if (seq1_index == len1) { curr_diag_upper = k; end1_diag = k - 1; }
            if (seq2_index == len2) { curr_diag_lower = k; end2_diag = k + 1; }
        }

        curr_score = curr_extent * match_score_half - d * score_common_factor;
        if (curr_score > max_score[d - 1]) {
            max_score[d] = curr_score;
            best_dist = d; best_diag = curr_diag;
            *seq2_len = curr_seq2_index;
            *seq1_len = curr_seq2_index + best_diag - diag_origin;
        } else max_score[d] = max_score[d - 1];

        if (curr_diag_lower <= curr_diag_upper) {
            num_nonempty_dist++;
            diag_lower[d] = curr_diag_lower; diag_upper[d] = curr_diag_upper;
        } else { diag_lower[d] = kInvalidDiag; diag_upper[d] = -kInvalidDiag; }
        if (diag_lower[d - max_penalty] <= diag_upper[d - max_penalty]) num_nonempty_dist--;
        if (num_nonempty_dist == 0) break;

        d++;
        curr_diag_lower = PMIN(diag_lower[d - gap_open_extend], diag_lower[d - gap_extend]) - 1;
        curr_diag_lower = PMIN(curr_diag_lower, diag_lower[d - op_cost]);
        if (end2_diag > 0) curr_diag_lower = PMAX(curr_diag_lower, end2_diag);
        curr_diag_upper = PMAX(diag_upper[d - gap_open_extend], diag_upper[d - gap_extend]) + 1;
        curr_diag_upper = PMAX(curr_diag_upper, diag_upper[d - op_cost]);
        if (end1_diag > 0) curr_diag_upper = PMIN(curr_diag_upper, end1_diag);
        if (d > max_penalty && d <= scaled_max_dist) rows[d] = rows[d - max_penalty - 1];
    }
    result = max_score[best_dist];
    free(store); free(rows); free(max_score_base); free(bounds);
    return result;
}

typedef struct GapResult {
    int32_t q_start, q_stop, s_start, s_stop, score, q_seed, s_seed;
} GapResult;

/* BLAST_GreedyGappedAlignment core/blast_gapalign.c:2620-2751 with gap costs 0/0
 * (BLAST_AffineGreedyAlign's dispatch, core/greedy_align.c:801-815: odd reward doubles
 * match/mismatch/xdrop). */
static void greedy_gapped(const uint8_t *query, const uint8_t *subject, int32_t qlen, int32_t slen,
                          int32_t q_off, int32_t s_off, int32_t reward, int32_t penalty,
                          int32_t gap_open, int32_t gap_extend, int32_t X, GreedyMem *mem, GapResult *g)
{
    int32_t q_ext_l, q_ext_r, s_ext_l, s_ext_r, score;
    int32_t match = reward, mismatch = -penalty, xd = X, go = gap_open, ge = gap_extend;
    GreedySeed fwd, rev;
    if (match % 2 == 1) { match *= 2; mismatch *= 2; xd *= 2; go *= 2; ge *= 2; }

    if (go == 0 && ge == 0) {
        score = greedy_align(query + q_off, qlen - q_off, subject + s_off / 4, slen - s_off, 0, xd,
                             match, mismatch, &q_ext_r, &s_ext_r, mem, s_off % 4, &fwd);
        score += greedy_align(query, q_off, subject, s_off, 1, xd, match, mismatch, &q_ext_l, &s_ext_l,
                              mem, 0, &rev);
        score = (q_ext_r + s_ext_r + q_ext_l + s_ext_l) * reward / 2 - score * (reward - penalty);
    } else {
        score = greedy_align_affine(query + q_off, qlen - q_off, subject + s_off / 4, slen - s_off, 0, xd,
                                    match, mismatch, go, ge, &q_ext_r, &s_ext_r, s_off % 4, &fwd);
        score += greedy_align_affine(query, q_off, subject, s_off, 1, xd, match, mismatch, go, ge,
                                     &q_ext_l, &s_ext_l, 0, &rev);
        if (reward % 2 == 1) score /= 2;
    }
    {
        int32_t q_box_l = q_off - q_ext_l, s_box_l = s_off - s_ext_l;
        int32_t q_box_r = q_off + q_ext_r, s_box_r = s_off + s_ext_r;
        int32_t q_seed_l = q_off - rev.start_q, s_seed_l = s_off - rev.start_s;
        int32_t q_seed_r = q_off + fwd.start_q, s_seed_r = s_off + fwd.start_s;
        int32_t vl = 0, vr = 0;
        if (q_seed_r < q_box_r && s_seed_r < s_box_r) {
            vr = PMIN(q_box_r - q_seed_r, s_box_r - s_seed_r);
            vr = PMIN(vr, fwd.match_length) / 2;
        } else { q_seed_r = q_off; s_seed_r = s_off; }
        if (q_seed_l > q_box_l && s_seed_l > s_box_l) {
            vl = PMIN(q_seed_l - q_box_l, s_seed_l - s_box_l);
            vl = PMIN(vl, rev.match_length) / 2;
        } else { q_seed_l = q_off; s_seed_l = s_off; }
        if (vr > vl) { g->q_seed = q_seed_r + vr; g->s_seed = s_seed_r + vr; }
        else { g->q_seed = q_seed_l - vl; g->s_seed = s_seed_l - vl; }
        g->q_start = q_box_l; g->s_start = s_box_l; g->q_stop = q_box_r; g->s_stop = s_box_r;
    }
    g->score = score;
}

/* ------------------------------------------------------------------ packed DP */
typedef struct DpMem { int32_t *best, *best_gap; int32_t alloc; } DpMem;
static void dp_reserve(DpMem *m, int32_t need)
{
    if (need > m->alloc) {
        m->alloc = PMAX(need + 100, 2 * m->alloc);
        m->best = (int32_t *)realloc(m->best, (size_t)m->alloc * 4);
        m->best_gap = (int32_t *)realloc(m->best_gap, (size_t)m->alloc * 4);
    }
}

/* s_BlastAlignPackedNucl core/blast_gapalign.c:2843-3056.
 * B = query bytes, A = packed subject; forward: B[1..N], bases of A from A[1] on;
 * reverse: B[N-1..0], subject bases M-1..0 of A. */
static int32_t dp_packed(const uint8_t *B, const uint8_t *A, int32_t N, int32_t M,
                         int32_t *b_offset, int32_t *a_offset, const int32_t *matrix,
                         int32_t gap_open, int32_t gap_extend, int32_t x_dropoff, int reverse,
                         DpMem *mem)
{
    int32_t i, a_index, b_index, b_size, first_b_index, last_b_index, b_inc;
    int32_t gap_open_extend = gap_open + gap_extend, num_extra_cells;
    int32_t score, score_gap_row, score_gap_col, next_score, best_score;
    *a_offset = 0; *b_offset = 0;
    if (x_dropoff < gap_open_extend) x_dropoff = gap_open_extend;
    if (N <= 0 || M <= 0) return 0;

    num_extra_cells = gap_extend > 0 ? x_dropoff / gap_extend + 3 : N + 3;
    dp_reserve(mem, num_extra_cells);

    score = -gap_open_extend;
    mem->best[0] = 0; mem->best_gap[0] = -gap_open_extend;
    for (i = 1; i <= N; i++) {
        if (score < -x_dropoff) break;
        dp_reserve(mem, i + 1);
        mem->best[i] = score; mem->best_gap[i] = score - gap_open_extend;
        score -= gap_extend;
    }
    b_size = i; best_score = 0; first_b_index = 0;
    b_inc = reverse ? -1 : 1;

    for (a_index = 1; a_index <= M; a_index++) {
        const int32_t *mrow;
        const uint8_t *b_ptr;
        int a_bp;
        if (reverse) a_bp = (A[(M - a_index) / 4] >> (2 * ((a_index - 1) % 4))) & 3;
        else a_bp = (A[1 + (a_index - 1) / 4] >> (2 * (3 - (a_index - 1) % 4))) & 3;
        mrow = matrix + 16 * a_bp;
        b_ptr = reverse ? &B[N - first_b_index] : &B[first_b_index];
        score = MININT; score_gap_row = MININT; last_b_index = first_b_index;

        for (b_index = first_b_index; b_index < b_size; b_index++) {
            b_ptr += b_inc;
            score_gap_col = mem->best_gap[b_index];
            next_score = mem->best[b_index] + mrow[*b_ptr];
            if (score < score_gap_col) score = score_gap_col;
            if (score < score_gap_row) score = score_gap_row;
            if (best_score - score > x_dropoff) {
                if (b_index == first_b_index) first_b_index++;
                else mem->best[b_index] = MININT;
            } else {
                last_b_index = b_index;
                if (score > best_score) { best_score = score; *a_offset = a_index; *b_offset = b_index; }
                score_gap_row -= gap_extend;
                score_gap_col -= gap_extend;
                mem->best_gap[b_index] = PMAX(score - gap_open_extend, score_gap_col);
                score_gap_row = PMAX(score - gap_open_extend, score_gap_row);
                mem->best[b_index] = score;
            }
            score = next_score;
        }
        if (first_b_index == b_size) break;
        dp_reserve(mem, last_b_index + num_extra_cells + 4);
        if (last_b_index < b_size - 1) b_size = last_b_index + 1;
        else {
            while (score_gap_row >= (best_score - x_dropoff) && b_size <= N) {
                mem->best[b_size] = score_gap_row;
                mem->best_gap[b_size] = score_gap_row - gap_open_extend;
                score_gap_row -= gap_extend;
                b_size++;
            }
        }
        if (b_size <= N) { mem->best[b_size] = MININT; mem->best_gap[b_size] = MININT; b_size++; }
    }
    return best_score;
}

/* s_BlastDynProgNtGappedAlignment core/blast_gapalign.c:2763-2825 */
static void dp_gapped(const uint8_t *query, const uint8_t *subject, int32_t qlen, int32_t slen,
                      int32_t q_off, int32_t s_off, const int32_t *matrix, int32_t gap_open,
                      int32_t gap_extend, int32_t X, DpMem *mem, GapResult *g)
{
    int32_t adj = 4 - (s_off % 4);
    int32_t q_length = q_off + adj, s_length = s_off + adj;
    int32_t pq, ps, left, right = 0;
    if (q_length > qlen || s_length > slen) { q_length -= 4; s_length -= 4; }
    left = dp_packed(query, subject, q_length, s_length, &pq, &ps, matrix, gap_open, gap_extend, X, 1, mem);
    g->q_start = q_length - pq; g->s_start = s_length - ps;
    if (q_length < qlen && s_length < slen) {
        right = dp_packed(query + q_length - 1, subject + (s_length + 3) / 4 - 1, qlen - q_length,
                          slen - s_length, &g->q_stop, &g->s_stop, matrix, gap_open, gap_extend, X, 0, mem);
        g->q_stop += q_length; g->s_stop += s_length;
    } else { g->q_stop = q_length; g->s_stop = s_length; }
    g->score = left + right;
}

/* ------------------------------------------------------------------ gapped stage */
typedef struct TreeHsp { int32_t q_strand_start, q_off, q_end, s_off, s_end, score, alive; } TreeHsp;

/* s_GetQueryStrandOffset core/blast_itree.c:219-234 (blastn: frames +1/-1 alternate, so the
 * strand offset is the context's own query_offset) */
static int32_t strand_offset(const BnQueryBatch *b, int32_t context)
{
    int32_t c = context;
    while (c) {
        int f = b->contexts[c].frame, pf = b->contexts[c - 1].frame;
        if (f == 0 || (f > 0) != (pf > 0) || (f < 0) != (pf < 0)) break;
        c--;
    }
    return b->contexts[c].query_offset;
}

/* s_HSPIsContained core/blast_itree.c:815-853 */
static int hsp_contained(const TreeHsp *in, const TreeHsp *t, int32_t min_diag_sep)
{
    if (in->q_strand_start != t->q_strand_start) return 0;
    if (in->score <= t->score &&
        t->q_off <= in->q_off && t->q_end >= in->q_off && t->s_off <= in->s_off && t->s_end >= in->s_off &&
        t->q_off <= in->q_end && t->q_end >= in->q_end && t->s_off <= in->s_end && t->s_end >= in->s_end) {
        if (min_diag_sep == 0) return 1;
        if (abs((t->q_off - t->s_off) - (in->q_off - in->s_off)) < min_diag_sep ||
            abs((t->q_end - t->s_end) - (in->q_end - in->s_end)) < min_diag_sep) return 1;
    }
    return 0;
}

/* s_HSPsHaveCommonEndpoint core/blast_itree.c:251-306: 0 none, 1 keep tree hsp, 2 keep new */
static int common_endpoint(const TreeHsp *in, const TreeHsp *t, int right)
{
    int match;
    if (in->q_strand_start != t->q_strand_start) return 0;
    match = right ? (in->q_end == t->q_end && in->s_end == t->s_end)
                  : (in->q_off == t->q_off && in->s_off == t->s_off);
    if (!match) return 0;
    if (in->score > t->score) return 2;
    if (in->score < t->score) return 1;
    if (in->q_end - in->q_off > t->q_end - t->q_off) return 1;
    if (in->q_end - in->q_off < t->q_end - t->q_off) return 2;
    if (in->s_end - in->s_off > t->s_end - t->s_off) return 1;
    if (in->s_end - in->s_off < t->s_end - t->s_off) return 2;
    return 1;
}

/* BlastIntervalTreeAddHSP (eQueryAndSubject) core/blast_itree.c:545-590, flat-list model */
static void tree_add(Vec *tree, const TreeHsp *in)
{
    int pass;
    int64_t i;
    for (pass = 0; pass < 2; pass++) {
        for (i = 0; i < tree->n; i++) {
            TreeHsp *t = (TreeHsp *)tree->p + i;
            int r;
            if (!t->alive) continue;
            r = common_endpoint(in, t, pass);
            if (r == 1) return;
            if (r == 2) t->alive = 0;
        }
    }
    *(TreeHsp *)vec_push(tree) = *in;
}

/* BLAST_GetGappedScore core/blast_gapalign.c:3233-3559 (blastn branches only) */
static void gapped_stage(const BnQueryBatch *b, const Subject *S, const BnInitHit *init, int64_t n_init,
                         const int32_t *low_score, GreedyMem *gm, DpMem *dm, Vec *hsps,
                         PortResults *out)
{
    const uint8_t *query = b->query_start + 1;
    Vec tree;
    int64_t i;
    int32_t *found_high = (int32_t *)calloc((size_t)b->num_queries, 4);
    vec_init(&tree, sizeof(TreeHsp));

    if (low_score) {
        for (i = 0; i < n_init; i++) {
            int32_t qi = b->contexts[ctx_search(b, init[i].q_off)].query_index;
            if (init[i].score > low_score[qi]) found_high[qi] = 1;
        }
    }
    for (i = 0; i < n_init; i++) {
        int32_t context = ctx_search(b, init[i].q_off);
        const BnContext *c = &b->contexts[context];
        int32_t qstart0 = c->query_offset;
        int32_t q_off = init[i].q_off - qstart0, s_off = init[i].s_off;
        int64_t k;
        int contained = 0;
        TreeHsp t;
        GapResult g;

        if (low_score && !found_high[c->query_index]) continue;

        t.q_strand_start = strand_offset(b, context);
        t.q_off = init[i].q_start - qstart0; t.q_end = t.q_off + init[i].length;
        t.s_off = init[i].s_start; t.s_end = t.s_off + init[i].length;
        t.score = init[i].score; t.alive = 1;
        for (k = 0; k < tree.n && !contained; k++) {
            const TreeHsp *h = (const TreeHsp *)tree.p + k;
            if (h->alive && hsp_contained(&t, h, b->min_diag_separation)) contained = 1;
        }
        if (contained) continue;
        out->stats.gap_extensions++;

        if (b->gap_algo == BN_GAP_GREEDY) {
            q_off = t.q_off + init[i].length / 2;
            s_off = init[i].s_start + init[i].length / 2;
            greedy_gapped(query + qstart0, S->seq, c->query_length, S->len, q_off, s_off,
                          b->reward, b->penalty, b->gap_open, b->gap_extend, b->gap_x_dropoff, gm, &g);
        } else {
            if (t.s_end >= s_off + 8) { s_off += 3; q_off += 3; }
            dp_gapped(query + qstart0, S->seq, c->query_length, S->len, q_off, s_off, b->matrix,
                      b->gap_open, b->gap_extend, b->gap_x_dropoff, dm, &g);
            g.q_seed = q_off; g.s_seed = s_off;
        }
        if (g.score >= c->gapped_cutoff) {
            BnHSP *h = (BnHSP *)vec_push(hsps);
            TreeHsp nt;
            h->oid = S->oid; h->context = context; h->chunk_off = S->chunk_off;
            h->q_off = g.q_start; h->q_end = g.q_stop; h->s_off = g.s_start; h->s_end = g.s_stop;
            h->score = g.score; h->q_gapped_start = g.q_seed; h->s_gapped_start = g.s_seed;
            h->evalue = 0;
            nt.q_strand_start = t.q_strand_start; nt.q_off = g.q_start; nt.q_end = g.q_stop;
            nt.s_off = g.s_start; nt.s_end = g.s_stop; nt.score = g.score; nt.alive = 1;
            tree_add(&tree, &nt);
        }
    }
    free(tree.p);
    free(found_high);
}

/* ------------------------------------------------------------------ HSP list post-processing */
static int cmp_qoff(const void *a, const void *c)
{   /* s_QueryOffsetCompareHSPs core/blast_hits.c:2035-2090 */
    const BnHSP *h1 = (const BnHSP *)a, *h2 = (const BnHSP *)c;
    if (h1->context != h2->context) return h1->context < h2->context ? -1 : 1;
    if (h1->q_off != h2->q_off) return h1->q_off < h2->q_off ? -1 : 1;
    if (h1->s_off != h2->s_off) return h1->s_off < h2->s_off ? -1 : 1;
    if (h1->score != h2->score) return h1->score < h2->score ? 1 : -1;
    if (h1->q_end != h2->q_end) return h1->q_end < h2->q_end ? 1 : -1;
    if (h1->s_end != h2->s_end) return h1->s_end < h2->s_end ? 1 : -1;
    return 0;
}
static int cmp_qend(const void *a, const void *c)
{   /* s_QueryEndCompareHSPs core/blast_hits.c:2103-2156 */
    const BnHSP *h1 = (const BnHSP *)a, *h2 = (const BnHSP *)c;
    if (h1->context != h2->context) return h1->context < h2->context ? -1 : 1;
    if (h1->q_end != h2->q_end) return h1->q_end < h2->q_end ? -1 : 1;
    if (h1->s_end != h2->s_end) return h1->s_end < h2->s_end ? -1 : 1;
    if (h1->score != h2->score) return h1->score < h2->score ? 1 : -1;
    if (h1->q_off != h2->q_off) return h1->q_off < h2->q_off ? 1 : -1;
    if (h1->s_off != h2->s_off) return h1->s_off < h2->s_off ? 1 : -1;
    return 0;
}
static int cmp_score(const void *a, const void *c)
{   /* ScoreCompareHSPs core/blast_hits.c:1182-1210 */
    const BnHSP *h1 = (const BnHSP *)a, *h2 = (const BnHSP *)c;
    if (h1->score != h2->score) return h2->score > h1->score ? 1 : -1;
    if (h1->s_off != h2->s_off) return h1->s_off > h2->s_off ? 1 : -1;
    if (h1->s_end != h2->s_end) return h2->s_end > h1->s_end ? 1 : -1;
    if (h1->q_off != h2->q_off) return h1->q_off > h2->q_off ? 1 : -1;
    if (h1->q_end != h2->q_end) return h2->q_end > h1->q_end ? 1 : -1;
    return 0;
}

/* Blast_HSPListPurgeHSPsWithCommonEndpoints(purge=TRUE) core/blast_hits.c:2224-2300 */
static int64_t purge_common(BnHSP *h, int64_t n)
{
    int64_t i, o;
    stable_sort(h, (size_t)n, sizeof *h, cmp_qoff);
    for (i = 0, o = 0; i < n; i++) {
        if (o > 0 && h[o - 1].context == h[i].context && h[o - 1].q_off == h[i].q_off &&
            h[o - 1].s_off == h[i].s_off) continue;
        h[o++] = h[i];
    }
    n = o;
    stable_sort(h, (size_t)n, sizeof *h, cmp_qend);
    for (i = 0, o = 0; i < n; i++) {
        if (o > 0 && h[o - 1].context == h[i].context && h[o - 1].q_end == h[i].q_end &&
            h[o - 1].s_end == h[i].s_end) continue;
        h[o++] = h[i];
    }
    return o;
}

/* Blast_HSPListsMerge for a split subject core/blast_hits.c:2545-2716
 * (s_BlastMergeTwoHSPs :1337-1375, OVERLAP_DIAG_CLOSE 10) */
static void merge_chunks(Vec *comb, BnHSP *nw, int64_t n_new, int32_t split_offset, int32_t overlap)
{
    BnHSP *c = (BnHSP *)comb->p;
    int64_t n1 = 0, n2 = 0, i, j, o;
    if (n_new == 0) return;
    if (comb->n == 0) {
        for (i = 0; i < n_new; i++) *(BnHSP *)vec_push(comb) = nw[i];
        return;
    }
    for (i = 0; i < comb->n; i++)
        if (c[i].s_end > split_offset) { BnHSP t = c[n1]; c[n1] = c[i]; c[i] = t; n1++; }
    for (i = 0; i < n_new; i++)
        if (nw[i].s_off < split_offset + overlap) { BnHSP t = nw[n2]; nw[n2] = nw[i]; nw[i] = t; n2++; }
    if (n1 > 0 && n2 > 0) {
        for (i = 0; i < n1; i++) {
            BnHSP *h1 = &c[i];
            for (j = 0; j < n2; j++) {
                BnHSP *h2 = &nw[j];
                if (h2->oid < 0 || h1->context != h2->context) continue;
                if (abs((h1->q_end - h1->s_end) - (h2->q_off - h2->s_off)) < 10) {
                    if ((h1->q_off <= h2->q_off && h1->q_end >= h2->q_off &&
                         h1->s_off <= h2->s_off && h1->s_end >= h2->s_off) ||
                        (h1->q_off <= h2->q_end && h1->q_end >= h2->q_end &&
                         h1->s_off <= h2->s_end && h1->s_end >= h2->s_end)) {
                        h1->q_off = PMIN(h1->q_off, h2->q_off); h1->s_off = PMIN(h1->s_off, h2->s_off);
                        h1->q_end = PMAX(h1->q_end, h2->q_end); h1->s_end = PMAX(h1->s_end, h2->s_end);
                        if (h2->score > h1->score) {
                            h1->q_gapped_start = h2->q_gapped_start;
                            h1->s_gapped_start = h2->s_gapped_start;
                            h1->score = h2->score;
                        }
                        h2->oid = -1;   /* freed */
                    }
                }
            }
        }
        for (i = 0, o = 0; i < n_new; i++) if (nw[i].oid >= 0) nw[o++] = nw[i];
        n_new = o;
    }
    for (i = 0; i < n_new; i++) *(BnHSP *)vec_push(comb) = nw[i];
    stable_sort(comb->p, (size_t)comb->n, sizeof(BnHSP), cmp_score);
}

/* ------------------------------------------------------------------ whole preliminary stage */
/* ------------------------------------------------------------------ hit lists behind hit_params->low_score
 * The collector splits a subject's list per query (core/hspfilter_collector.c:104-150) and files each part with
 * Blast_HitListUpdate (core/blast_hits.c:2924-2981) into a hit list of prelim_hitlist_size entries
 * (BlastHSPCollectorParamsNew, core/hspfilter_collector.c:335-342); once a list is full and a better subject
 * arrives it becomes a heap with the worst subject at the root (s_CreateHeap / s_Heapify :1470-1521 under
 * s_EvalueCompareHSPLists :2759-2788).  The engine then raises low_score[query] to low_score_perc x the root's
 * best score (core/blast_engine.c:1313-1320). */
typedef struct HlKey { double best_evalue; int32_t best_score, oid; } HlKey;
typedef struct HitList { HlKey *a; int32_t count, heapified, low_score; double worst_evalue; } HitList;

static int hl_fuzzy(double e1, double e2)          /* s_FuzzyEvalueComp :2742-2753 */
{
    if (e1 < (1 - 1e-6) * e2) return -1;
    if (e1 > (1 + 1e-6) * e2) return 1;
    return 0;
}
static int hl_cmp(const HlKey *x, const HlKey *y)
{
    int r = hl_fuzzy(x->best_evalue, y->best_evalue);
    if (r) return r;
    if (x->best_score > y->best_score) return -1;
    if (x->best_score < y->best_score) return 1;
    return y->oid > x->oid ? 1 : (y->oid < x->oid ? -1 : 0);
}
static void hl_heapify(HlKey *a, int64_t base, int64_t lim, int64_t last)
{
    int64_t left = 2 * base + 1;
    while (base <= lim) {
        int64_t large = (left == last) ? left : (hl_cmp(&a[left], &a[left + 1]) >= 0 ? left : left + 1);
        if (hl_cmp(&a[base], &a[large]) < 0) {
            HlKey t = a[base]; a[base] = a[large]; a[large] = t;
            base = large; left = 2 * base + 1;
        } else break;
    }
}
static void hl_update(HitList *L, int32_t max, HlKey k)
{
    if (L->count < max) {
        if (!L->a) { L->a = (HlKey *)malloc((size_t)max * sizeof(HlKey)); L->low_score = INT_MAX; }
        L->a[L->count++] = k;
        if (k.best_evalue > L->worst_evalue) L->worst_evalue = k.best_evalue;
        if (k.best_score < L->low_score) L->low_score = k.best_score;
        return;
    }
    {
        const int order = hl_fuzzy(k.best_evalue, L->worst_evalue);
        if (order > 0 || (order == 0 && k.best_score < L->low_score)) return;
    }
    if (!L->heapified) {
        const int64_t n = L->count;
        if (n >= 2) {
            int64_t i;
            for (i = n / 2; i > 0; i--) hl_heapify(L->a, i - 1, (n - 2) / 2, n - 1);
        }
        L->heapified = 1;
    }
    L->a[0] = k;
    if (L->count >= 2) hl_heapify(L->a, 0, L->count / 2 - 1, L->count - 1);
    L->worst_evalue = L->a[0].best_evalue;
    L->low_score = L->a[0].best_score;
}
/* one subject's final list (sorted by score) -> hit lists -> low_score[] */
static void hl_subject_done(const BnQueryBatch *b, HitList *lists, int32_t max, const BnHSP *h, int64_t n,
                            int32_t *low_score)
{
    int64_t i, j;
    for (i = 0; i < n; i++) {
        const int32_t qi = b->contexts[h[i].context].query_index;
        HlKey k;
        int seen = 0;
        for (j = 0; j < i && !seen; j++) seen = b->contexts[h[j].context].query_index == qi;
        if (seen) continue;
        k.best_evalue = h[i].evalue; k.best_score = h[i].score; k.oid = h[i].oid;
        for (j = i + 1; j < n; j++)
            if (b->contexts[h[j].context].query_index == qi && h[j].evalue < k.best_evalue) k.best_evalue = h[j].evalue;
        hl_update(&lists[qi], max, k);
    }
    for (i = 0; i < b->num_queries; i++)
        if (lists[i].heapified) {
            const double v = b->low_score_perc * (double)lists[i].low_score;
            if ((double)low_score[i] < v) low_score[i] = (int32_t)v;
        }
}

int port_prelim_search_masked(const BnQueryBatch *b, const uint8_t *packed, const int64_t *seq_byte_off,
                              const int32_t *seq_len, int32_t n_seq, int taps, int32_t smask_type,
                              const int32_t *smask_n, const int32_t *smask_iv, PortResults *out);

int port_prelim_search(const BnQueryBatch *b, const uint8_t *packed, const int64_t *seq_byte_off,
                       const int32_t *seq_len, int32_t n_seq, int taps, PortResults *out)
{
    return port_prelim_search_masked(b, packed, seq_byte_off, seq_len, n_seq, taps, 0, NULL, NULL, out);
}

/* smask_*: database masks (smask_type 1 soft / 2 hard): smask_n[i] masked [begin, end) intervals of subject i,
 * flat pairs in smask_iv */
int port_prelim_search_masked(const BnQueryBatch *b, const uint8_t *packed, const int64_t *seq_byte_off,
                              const int32_t *seq_len, int32_t n_seq, int taps, int32_t smask_type,
                              const int32_t *smask_n, const int32_t *smask_iv, PortResults *out)
{
    int64_t *smask_first = NULL;
    DiagState diag;
    Vec init, gapped_tap, final_, scanv, scan_oid, scan_chunk;
    GreedyMem gm;
    DpMem dm;
    int32_t oid, max_len = 0;
    int32_t *low_score = NULL;
    HitList *hitlists = NULL; int32_t hitlist_max = 0;       /* per query: hit-list model */

    memset(out, 0, sizeof *out);
    vec_init(&init, sizeof(BnInitHit)); vec_init(&gapped_tap, sizeof(BnHSP));
    vec_init(&final_, sizeof(BnHSP)); vec_init(&scanv, sizeof(BnOffsetPair));
    vec_init(&scan_oid, 4); vec_init(&scan_chunk, 4);
    for (oid = 0; oid < n_seq; oid++) if (seq_len[oid] > max_len) max_len = seq_len[oid];
    gm.max_d = PMIN(GREEDY_MAX_COST, max_len / 2 + 1);
    gm.row[0] = (int32_t *)calloc((size_t)(2 * gm.max_d + 6) * 2, 4);
    gm.row[1] = gm.row[0] + 2 * gm.max_d + 6;
    gm.max_score = (int32_t *)calloc((size_t)gm.max_d + 1 + 4096, 4);
    memset(&dm, 0, sizeof dm);
    diag_new(&diag, b);
    if (b->low_score_perc > 0.00001) {
        int32_t hs = b->hitlist_size > 0 ? b->hitlist_size : 500;
        low_score = (int32_t *)calloc((size_t)b->num_queries, 4);
        hitlists = (HitList *)calloc((size_t)b->num_queries, sizeof(HitList));
        hs = PMIN(2 * hs, hs + 50);
        hitlist_max = PMAX(hs, 10);
    }
    if (smask_type && smask_n) {
        smask_first = (int64_t *)calloc((size_t)n_seq + 1, 8);
        for (oid = 0; oid < n_seq; oid++) smask_first[oid + 1] = smask_first[oid] + smask_n[oid];
    }

    for (oid = 0; oid < n_seq; oid++) {
        /* s_BlastSearchEngineOneContext chunk loop core/blast_engine.c:459-541,
         * s_GetNextSubjectChunk :220-301 (no masking: residual 0) */
        const uint8_t *base = packed + seq_byte_off[oid];
        const int32_t full = seq_len[oid];
        int32_t next = 0;
        Vec comb;
        /* database masks: unmasked ranges as BlastSeqBlkSetSeqRanges leaves them (core/blast_util.c:186-223) */
        const int32_t mt = (smask_type && smask_n) ? smask_type : 0;
        int32_t n_r = 1, hm = 0, n_hard = 1, n_soft = 1;
        int32_t *R = NULL, *chunk_r = NULL;
        int32_t full_range[2];
        const int32_t *hard, *soft;
        full_range[0] = 0; full_range[1] = full;
        if (mt) {
            const int32_t nm = smask_n[oid];
            const int32_t *iv = smask_iv + 2 * smask_first[oid];
            int32_t k;
            n_r = nm + 1;
            R = (int32_t *)calloc((size_t)n_r * 2, 4);
            for (k = 0; k < nm; k++) { R[2 * k + 1] = iv[2 * k]; R[2 * k + 2] = iv[2 * k + 1]; }
            R[0] = 0; R[2 * (n_r - 1) + 1] = full;
            chunk_r = (int32_t *)calloc((size_t)n_r * 2, 4);
        }
        hard = (mt == 2) ? R : full_range; n_hard = (mt == 2) ? n_r : 1;
        soft = (mt == 1) ? R : full_range; n_soft = (mt == 1) ? n_r : 1;
        next = hard[0];
        vec_init(&comb, sizeof(BnHSP));
        while (next < full) {
            Subject S;
            const int32_t residual = next % 4;
            int32_t offset = next - residual;
            int64_t first_init = init.n, first_hsp, n_h;
            Vec hs;
            S.seq = base + offset / 4; S.oid = oid; S.chunk_off = offset;
            if ((int64_t)offset + BN_MAX_DBSEQ_LEN < hard[2 * hm + 1]) {
                S.len = BN_MAX_DBSEQ_LEN;
                next = offset + BN_MAX_DBSEQ_LEN - BN_DBSEQ_CHUNK_OVERLAP;
            } else {
                S.len = hard[2 * hm + 1] - offset;
                hm++;
                next = hm < n_hard ? hard[2 * hm] : full;
            }
            S.masked = mt != 0; S.n_ranges = 1; S.ranges = full_range;
            if (mt) {
                if (offset == 0 && residual == 0 && next == full) { S.ranges = soft; S.n_ranges = n_soft; }
                else if (mt != 1) { chunk_r[0] = residual; chunk_r[1] = S.len; S.ranges = chunk_r; S.n_ranges = 1; }
                else {
                    int32_t i = 0, st, cnt, k;
                    const int32_t end = offset + S.len;
                    while (soft[2 * i + 1] < offset) ++i;
                    st = i;
                    while (i < n_soft && soft[2 * i] < end) ++i;
                    cnt = i - st;
                    if (cnt == 0) continue;                       /* SUBJECT_SPLIT_NO_RANGE */
                    for (k = 0; k < cnt; k++) { chunk_r[2 * k] = soft[2 * (st + k)] - offset; chunk_r[2 * k + 1] = soft[2 * (st + k) + 1] - offset; }
                    if (chunk_r[0] < 0) chunk_r[0] = 0;
                    if (chunk_r[2 * (cnt - 1) + 1] > S.len) chunk_r[2 * (cnt - 1) + 1] = S.len;
                    S.ranges = chunk_r; S.n_ranges = cnt;
                }
            }
            out->stats.subject_bases_scanned += S.len;

            word_finder(b, &S, &diag, &init, first_init, out, taps, &scanv, &scan_oid, &scan_chunk);
            if (init.n == first_init) continue;

            vec_init(&hs, sizeof(BnHSP));
            gapped_stage(b, &S, (BnInitHit *)init.p + first_init, init.n - first_init, low_score,
                         &gm, &dm, &hs, out);
            if (!(taps & PORT_TAP_INIT)) init.n = first_init;
            if (taps & PORT_TAP_GAPPED) {
                int64_t k;
                for (k = 0; k < hs.n; k++) *(BnHSP *)vec_push(&gapped_tap) = ((BnHSP *)hs.p)[k];
            }
            first_hsp = 0; (void)first_hsp;
            n_h = purge_common((BnHSP *)hs.p, hs.n);
            if (b->round_down) { int64_t k; for (k = 0; k < n_h; k++) ((BnHSP *)hs.p)[k].score &= ~1; }
            stable_sort(hs.p, (size_t)n_h, sizeof(BnHSP), cmp_score);
            if (n_h > 0) {
                int64_t k;
                for (k = 0; k < n_h; k++) {
                    BnHSP *h = (BnHSP *)hs.p + k;
                    h->s_off += offset; h->s_end += offset; h->s_gapped_start += offset;
                }
                merge_chunks(&comb, (BnHSP *)hs.p, n_h, offset, offset == 0 ? 0 : BN_DBSEQ_CHUNK_OVERLAP);
            }
            free(hs.p);
        }
        /* E-values and reap: s_BlastSearchEngineCore core/blast_engine.c:788-806,
         * BLAST_KarlinStoE_simple core/blast_stat.c:4111-4125 */
        {
            int64_t k, kept = 0;
            const int64_t first_final = final_.n;
            for (k = 0; k < comb.n; k++) {
                BnHSP *h = (BnHSP *)comb.p + k;
                const BnContext *c = &b->contexts[h->context];
                h->evalue = (double)c->eff_searchsp * exp((double)(-c->gap_lambda * h->score) + c->gap_logK);
                if (h->evalue > b->evalue_cutoff) continue;
                *(BnHSP *)vec_push(&final_) = *h;
                kept++;
            }
            if (kept) out->stats.good_extensions += kept;
            if (kept && hitlists)
                hl_subject_done(b, hitlists, hitlist_max, (const BnHSP *)final_.p + first_final, kept, low_score);
        }
        free(comb.p);
        free(R); free(chunk_r);
    }
    diag_free(&diag);
    free(smask_first);
    free(gm.row[0]); free(gm.max_score); free(dm.best); free(dm.best_gap); free(low_score);
    if (hitlists) { int32_t q; for (q = 0; q < b->num_queries; q++) free(hitlists[q].a); free(hitlists); }
    out->hsps = (BnHSP *)final_.p; out->n_hsps = final_.n;
    out->init = (BnInitHit *)init.p; out->n_init = init.n;
    out->gapped = (BnHSP *)gapped_tap.p; out->n_gapped = gapped_tap.n;
    out->scan = (BnOffsetPair *)scanv.p; out->n_scan = scanv.n;
    out->scan_oid = (int32_t *)scan_oid.p; out->scan_chunk = (int32_t *)scan_chunk.p;
    return BN_OK;
}

void port_results_free(PortResults *r)
{
    free(r->hsps); free(r->init); free(r->gapped); free(r->scan); free(r->scan_oid); free(r->scan_chunk);
    memset(r, 0, sizeof *r);
}

int port_greedy_align(const uint8_t *query, int32_t qlen, const uint8_t *subject_packed,
                      int32_t slen, int32_t q_off, int32_t s_off, int32_t reward, int32_t penalty,
                      int32_t xdrop, int32_t o[7])
{
    GreedyMem gm;
    GapResult g;
    gm.max_d = PMIN(GREEDY_MAX_COST, slen / 2 + 1);
    gm.row[0] = (int32_t *)calloc((size_t)(2 * gm.max_d + 6) * 2, 4);
    gm.row[1] = gm.row[0] + 2 * gm.max_d + 6;
    gm.max_score = (int32_t *)calloc((size_t)gm.max_d + 1 + 4096, 4);
    greedy_gapped(query, subject_packed, qlen, slen, q_off, s_off, reward, penalty, 0, 0, xdrop, &gm, &g);
    o[0] = g.q_start; o[1] = g.q_stop; o[2] = g.s_start; o[3] = g.s_stop; o[4] = g.score;
    o[5] = g.q_seed; o[6] = g.s_seed;
    free(gm.row[0]); free(gm.max_score);
    return 0;
}

int port_dp_align(const uint8_t *query, int32_t qlen, const uint8_t *subject_packed, int32_t slen,
                  int32_t q_off, int32_t s_off, const int32_t *matrix16, int32_t gap_open,
                  int32_t gap_extend, int32_t xdrop, int32_t o[5])
{
    DpMem dm;
    GapResult g;
    memset(&dm, 0, sizeof dm);
    dp_gapped(query, subject_packed, qlen, slen, q_off, s_off, matrix16, gap_open, gap_extend, xdrop, &dm, &g);
    o[0] = g.q_start; o[1] = g.q_stop; o[2] = g.s_start; o[3] = g.s_stop; o[4] = g.score;
    free(dm.best); free(dm.best_gap);
    return 0;
}
