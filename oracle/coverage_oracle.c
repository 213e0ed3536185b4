/*
 * oracle/coverage_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of oarfish's bulk coverage model (--model-coverage), the pre-EM stage that
 * produces InMemoryAlignmentStore.coverage_probabilities:
 *
 *   add_interval                      src/util/oarfish_types.rs:496-537 (called per alignment at
 *                                     ingest, :724-727, weight 1.0)
 *   get_normalized_counts_and_lengths src/util/oarfish_types.rs:471-493 (f64 bins -> f32 counts)
 *   logistic / logstic_function       src/util/logistic_probability.rs:7-38
 *   logistic_prob                     src/util/logistic_probability.rs:40-79 (min_cov = total_weight/100)
 *   normalize_read_probs              src/util/normalize_probability.rs:5-74
 *
 * Parity is unpinned by the reference (no tests touch these functions).  Quirks are kept as they are:
 * add_interval and normalize_read_probs iterate the half-open bin range start_bin..end_bin, so the bin
 * holding the alignment's end is never visited unless it is also the start bin.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t n_bins_of(uint32_t len, uint32_t bin_width)
{   /* TranscriptInfo::with_len_and_bin_width, oarfish_types.rs:460-468 */
    return (uint32_t)ceil((double)len / (double)bin_width);
}

/* bin_off: M+1 prefix of the per-transcript bin counts (returned total = bin_off[M]). */
uint64_t oracle_cov_bin_offsets(const uint32_t *txp_len, uint32_t n_txps, uint32_t bin_width, uint64_t *bin_off)
{
    bin_off[0] = 0;
    for (uint32_t t = 0; t < n_txps; ++t) bin_off[t + 1] = bin_off[t] + n_bins_of(txp_len[t], bin_width);
    return bin_off[n_txps];
}

/* add_interval for every alignment (oarfish_types.rs:496-537); bins and total_weight must be zeroed. */
void oracle_cov_add_intervals(const uint32_t *txp, const uint32_t *start, const uint32_t *end, uint64_t nnz,
                              const uint32_t *txp_len, const uint64_t *bin_off, double *bins, double *total_weight)
{
    for (uint64_t j = 0; j < nnz; ++j) {
        const uint32_t t = txp[j];
        const uint32_t nI = (uint32_t)(bin_off[t + 1] - bin_off[t]);
        const double nIf = (double)nI, tlen = (double)txp_len[t];
        const double bw = round(tlen / nIf);                            /* :501 */
        const uint32_t st = start[j] < end[j] ? start[j] : end[j];      /* :502 */
        const uint32_t sp = st > end[j] ? st : end[j];                  /* :503 */
        const uint32_t sb = (uint32_t)floor(((double)st / tlen) * nIf); /* :504 */
        const uint32_t eb = (uint32_t)floor(((double)sp / tlen) * nIf); /* :505 */
        double *b = bins + bin_off[t];
        for (uint32_t i = sb; i < eb && i < nI; ++i) {                  /* coverage_bins[start_bin..end_bin] */
            const double bf = (double)i;
            const uint32_t cbs = (uint32_t)(bf * bw);
            double ce = (bf + 1.0) * bw; if (ce > tlen) ce = tlen;
            const uint32_t cbe = (uint32_t)ce;
            uint32_t olap = 0;
            if (st <= cbe) olap = (sp < cbe ? sp : cbe) - (st > cbs ? st : cbs);   /* :507-513 */
            b[i] += (double)olap / (double)(cbe - cbs);                 /* :523-525 */
        }
        total_weight[t] += 1.0;                                         /* :536 */
    }
}

/* logistic_prob (logistic_probability.rs:40-79): bins += total_weight/100, then the clamped logistic of the
 * relative deficit of each bin's f32 count.  cov_prob has the same layout as bins. */
void oracle_cov_logistic(uint32_t n_txps, const uint64_t *bin_off, double *bins, const double *total_weight,
                         double growth_rate, double *cov_prob)
{
    for (uint32_t t = 0; t < n_txps; ++t) {
        const uint64_t o = bin_off[t];
        const uint32_t n = (uint32_t)(bin_off[t + 1] - o);
        const double min_cov = total_weight[t] / 100.0;                 /* :51 */
        double count_sum = 0.0;
        for (uint32_t i = 0; i < n; ++i) { bins[o + i] += min_cov; count_sum += (double)(float)bins[o + i]; }   /* :52, :19 */
        if (count_sum <= 1e-8) { for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] = 0.0; continue; }          /* :21-23 */
        const double expected = count_sum / (double)n;                  /* :27 */
        for (uint32_t i = 0; i < n; ++i) {
            const double diff = (expected - (double)(float)bins[o + i]) / expected;   /* :32 */
            double r = 1.0 / (1.0 + exp(-growth_rate * diff));          /* :8 */
            if (r < 1e-8) r = 1e-8; else if (r > 0.99999) r = 0.99999;  /* :9 */
            cov_prob[o + i] = r;
        }
    }
}

/* normalize_read_probs (normalize_probability.rs:5-74) -> coverage_probabilities (nnz f64). */
void oracle_cov_normalize(const uint64_t *row_ptr, const uint32_t *txp, const uint32_t *start, const uint32_t *end,
                          uint64_t n_reads, const uint32_t *txp_len, const uint64_t *bin_off, const double *cov_prob,
                          uint32_t bin_width, double *out)
{
    const double bl = (double)bin_width;
    for (uint64_t r = 0; r < n_reads; ++r) {
        double nsum = 0.0;
        for (uint64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) {
            const uint32_t t = txp[j];
            const double sa = (double)start[j], ea = (double)end[j], tlen = (double)txp_len[t];
            const double *cp = cov_prob + bin_off[t];
            const uint64_t nb = bin_off[t + 1] - bin_off[t];
            const uint64_t sb = (uint64_t)(sa / bl);
            uint64_t eb = (uint64_t)(ea / bl); if (eb > nb - 1) eb = nb - 1;      /* :27-28 */
            double tw = 0.0, cpv = 0.0;
            if (sb == eb) {                                                        /* :34-36 */
                const double w = (ea - sa) / bl; tw = w; cpv = w * cp[sb];
            } else {
                for (uint64_t i = sb; i < eb; ++i) {                               /* :38-48, (start_bin..end_bin) */
                    double w;
                    if (i == sb) { double be = bl * (double)sb + bl; if (be > tlen) be = tlen; w = (be - sa) / bl; }
                    else w = 1.0;                                                  /* i == end_bin is unreachable */
                    tw += w; cpv += w * cp[i];
                }
            }
            const double e = cpv / tw;                                             /* :60 */
            nsum += e; out[j] = e;
        }
        const double d = nsum > 0.0 ? nsum : 1.0;                                  /* :64 */
        for (uint64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) out[j] /= d;
    }
}

/* The whole stage, as bulk.rs:103-108 runs it. */
void oracle_coverage_model(const uint64_t *row_ptr, const uint32_t *txp, const uint32_t *start, const uint32_t *end,
                           uint64_t n_reads, uint64_t nnz, const uint32_t *txp_len, uint32_t n_txps, uint32_t bin_width,
                           double growth_rate, double *out)
{
    uint64_t *bin_off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_txps + 1));
    const uint64_t nb = oracle_cov_bin_offsets(txp_len, n_txps, bin_width, bin_off);
    double *bins = (double *)calloc(nb ? nb : 1, sizeof(double));
    double *tw = (double *)calloc(n_txps ? n_txps : 1, sizeof(double));
    double *cp = (double *)calloc(nb ? nb : 1, sizeof(double));
    oracle_cov_add_intervals(txp, start, end, nnz, txp_len, bin_off, bins, tw);
    oracle_cov_logistic(n_txps, bin_off, bins, tw, growth_rate, cp);
    oracle_cov_normalize(row_ptr, txp, start, end, n_reads, txp_len, bin_off, cp, bin_width, out);
    free(bin_off); free(bins); free(tw); free(cp);
}
