/*
 * oracle/filter_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of AlignmentFilters::filter (src/util/oarfish_types.rs:955-1130), the per-read filter and
 * score -> probability step that builds the store the EM runs on, followed by
 * add_filtered_group (:718-738: groups without a retained alignment are dropped).  Records come as columns, one
 * entry per alignment record, exactly the fields filter() reads through AlnRecordLike (:186-202).
 * Parity is unpinned by the reference (no test calls filter()).
 *
 * discard[10] follows DiscardTable's declaration order (:812-826).
 */
#include <math.h>
#include <stdint.h>

enum { D5P = 0, D3P, DSCORE, DFRAC, DLEN, DORI, DSUPP, NOMAP, NOVALID, VALIDBEST };

typedef struct {
    int32_t which_strand;          /* 0 unknown, 1 forward, 2 reverse */
    uint32_t min_aligned_len;
    int64_t three_prime_clip;
    uint32_t five_prime_clip;
    float min_aligned_fraction, score_threshold, score_prob_denom;
} filter_opts;

/* Returns the number of retained alignments; out_row_ptr (u64, capacity n_groups + 1) / out_txp / out_prob / out_src
 * (capacity n_records) / out_group (capacity n_groups) receive the store; *out_rows the number of retained reads. */
uint64_t oracle_filter(const uint64_t *group_ptr, const uint32_t *ref_id, const uint32_t *aln_start, const uint32_t *aln_end,
                       const uint32_t *aln_span, const int32_t *score, const uint8_t *flags, const uint32_t *seq_len_rec,
                       uint64_t n_groups, const uint32_t *txp_len, const filter_opts *f, uint64_t *out_row_ptr,
                       uint32_t *out_txp, float *out_prob, uint32_t *out_src, uint32_t *out_group, uint64_t *out_rows,
                       uint64_t *discard)
{
    uint64_t nnz = 0, rows = 0;
    for (int k = 0; k < 10; ++k) discard[k] = 0;
    out_row_ptr[0] = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
        const uint64_t b = group_ptr[g], e = group_ptr[g + 1];
        int32_t best = INT32_MIN;                       /* :975 */
        float frac_at_best = 0.0f;                      /* :978 */
        uint32_t len_at_best = 0;                       /* :981 */
        uint64_t n_mapped = 0;
        uint32_t seq_len = 0;
        for (uint64_t j = b; j < e; ++j) {
            if (!(flags[j] & 1)) ++n_mapped;            /* :986 */
            if (seq_len == 0 && seq_len_rec[j] != 0) seq_len = seq_len_rec[j];   /* :991-994 */
        }
        /* ag.retain(...) (:997-1072): mark survivors */
        uint64_t first = nnz, kept = 0;
        for (uint64_t j = b; j < e; ++j) {
            if (flags[j] & 1) continue;                                          /* unmapped */
            const int rc = (flags[j] & 2) != 0;
            if ((f->which_strand == 1 && rc) || (f->which_strand == 2 && !rc)) { discard[DORI]++; continue; }
            if (flags[j] & 4) { discard[DSUPP]++; continue; }
            if (aln_span[j] < f->min_aligned_len) { discard[DLEN]++; continue; }
            if ((int64_t)aln_end[j] <= (int64_t)txp_len[ref_id[j]] - f->three_prime_clip) { discard[D3P]++; continue; }
            if (aln_start[j] >= f->five_prime_clip) { discard[D5P]++; continue; }
            if (score[j] > best) {
                best = score[j];
                len_at_best = aln_span[j];
                frac_at_best = seq_len > 0 ? (float)aln_span[j] / (float)seq_len : 0.0f;
            }
            out_src[first + kept++] = (uint32_t)j;                               /* survivors, provisional */
        }
        if (kept == 0 || len_at_best == 0 || best <= 0) {                       /* :1074-1086 */
            if (n_mapped == 0) discard[NOMAP]++; else discard[NOVALID]++;
            continue;
        }
        if (frac_at_best < f->min_aligned_fraction) { discard[DFRAC]++; continue; }   /* :1087-1092 */
        discard[VALIDBEST]++;
        const float mscore = (float)best, inv_max = 1.0f / mscore;              /* :1098-1099 */
        uint64_t w = 0;
        for (uint64_t k = 0; k < kept; ++k) {
            const uint32_t j = out_src[first + k];
            const float fs = (float)score[j];
            if (fs * inv_max >= f->score_threshold) {                            /* :1110-1114 */
                out_src[first + w] = j;
                out_txp[first + w] = ref_id[j];
                out_prob[first + w] = expf((fs - mscore) / f->score_prob_denom);
                ++w;
            } else discard[DSCORE]++;
        }
        if (w == 0) continue;                                                    /* add_filtered_group drops empty groups */
        nnz += w;
        out_group[rows] = (uint32_t)g;
        out_row_ptr[++rows] = nnz;
    }
    *out_rows = rows;
    return nnz;
}
