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
 *   binomial_continuous_prob          src/util/binomial_probability.rs:180-224 (the single-cell driver's model)
 *   binomial_probability              src/util/binomial_probability.rs:7-178 (ln_gamma: statrs 0.18, restated)
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

/* statrs 0.18 (Cargo.toml:46, not vendored under /root/reference) statrs::function::gamma::ln_gamma: the Lanczos
 * approximation with g = 10.900511 and the 11 coefficients below (Godfrey), branch x >= 0.5 -- the only one the
 * coverage model reaches (arguments are counts + 1).  Restated from the crate's published source; pinned in
 * tests/test_oracle_kat.py against libm's lgamma (agreement to ~1e-14 relative). */
double oracle_statrs_ln_gamma(double x)
{
    static const double dk[11] = {
        2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
        -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
        -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9 };
    const double gamma_r = 10.900511, ln_2_sqrt_e_over_pi = 0.6207822376352452223455184457816472122518527279025978;
    double sum = dk[0];
    for (int k = 1; k < 11; ++k) sum += dk[k] / (x + (double)k - 1.0);
    return log(sum) + ln_2_sqrt_e_over_pi + (x - 0.5) * log((x - 0.5 + gamma_r) / 2.718281828459045235360287471352662497757);
}

/* binomial_continuous_prob + binomial_probability (src/util/binomial_probability.rs:180-224, :7-178), the coverage
 * model of the single-cell driver (single_cell.rs:132-137): bins += total_weight/100, f32 counts and f32 bin
 * lengths (get_normalized_counts_and_lengths, oarfish_types.rs:471-493), counts rescaled so that the fullest bin is
 * 709, binomial pmf of every bin against the rest in log space, normalised over the transcript's bins.  Sums that
 * the reference takes in f32 (count_sum :14, sum_vec :74) are f32 here, in the same order.  The reference panics on
 * NaN / infinite intermediates (:88-127); this restatement lets them through. */
void oracle_cov_binomial(uint32_t n_txps, const uint32_t *txp_len, const uint64_t *bin_off, double *bins,
                         const double *total_weight, double *cov_prob)
{
    const double zero_thresh = 1e-20, max_scale = 709.0;
    for (uint32_t t = 0; t < n_txps; ++t) {
        const uint64_t o = bin_off[t];
        const uint32_t n = (uint32_t)(bin_off[t + 1] - o);
        if (n == 0) continue;
        const double min_cov = total_weight[t] / 100.0;                                   /* :191 */
        const float bwf = (float)round((double)txp_len[t] / (double)n);                   /* oarfish_types.rs:475 */
        float count_sum = 0.0f, max_count = NAN;
        double rate = 0.0;
        for (uint32_t i = 0; i < n; ++i) {
            bins[o + i] += min_cov;                                                       /* :192 */
            const float c = (float)bins[o + i];
            const float bs = (float)i * bwf, be = fminf(((float)i + 1.0f) * bwf, (float)(double)txp_len[t]);
            const float len = be - bs;                                                    /* oarfish_types.rs:480-483 */
            rate += (double)c / (double)len;                                              /* :195-199 */
            count_sum += c;                                                               /* :14 */
            max_count = fmaxf(max_count, c);                                              /* :50 */
        }
        if (count_sum == 0.0f || rate == 0.0) { for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] = 0.0; continue; }   /* :19-25 */
        float sum_vec = 0.0f;
        for (uint32_t i = 0; i < n; ++i) {
            const float c = (float)bins[o + i];
            const float m = c == max_count ? (float)max_scale : (float)(((double)c * max_scale) / (double)max_count);   /* :62-72 */
            sum_vec += m;                                                                 /* :74 */
        }
        const double ln1 = oracle_statrs_ln_gamma((double)sum_vec + 1.0);                 /* :77 */
        double total = 0.0;
        for (uint32_t i = 0; i < n; ++i) {
            const float c = (float)bins[o + i];
            const float bs = (float)i * bwf, be = fminf(((float)i + 1.0f) * bwf, (float)(double)txp_len[t]);
            const float len = be - bs;
            const double prob = (c == 0.0f || len == 0.0f) ? 0.0 : (double)c / ((double)len * rate);   /* :27-43 */
            const float m = c == max_count ? (float)max_scale : (float)(((double)c * max_scale) / (double)max_count);
            const float rest = sum_vec - m;
            const double den = oracle_statrs_ln_gamma((double)m + 1.0) + oracle_statrs_ln_gamma((double)rest + 1.0);   /* :78-81 */
            const double num2 = (prob > zero_thresh ? log(prob) : log(zero_thresh)) * (double)m;                 /* :84 */
            const double num3 = ((1.0 - prob) > zero_thresh ? log(1.0 - prob) : log(zero_thresh)) * (double)rest; /* :91 */
            const double res = exp(ln1 - den + num2 + num3);                              /* :103 */
            cov_prob[o + i] = res; total += res;
        }
        for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] /= total;                        /* :120-134 */
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

/* The whole stage with the binomial model, as single_cell.rs:132-137 runs it per cell. */
void oracle_coverage_model_binomial(const uint64_t *row_ptr, const uint32_t *txp, const uint32_t *start, const uint32_t *end,
                                    uint64_t n_reads, uint64_t nnz, const uint32_t *txp_len, uint32_t n_txps, uint32_t bin_width,
                                    double *out)
{
    uint64_t *bin_off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_txps + 1));
    const uint64_t nb = oracle_cov_bin_offsets(txp_len, n_txps, bin_width, bin_off);
    double *bins = (double *)calloc(nb ? nb : 1, sizeof(double));
    double *tw = (double *)calloc(n_txps ? n_txps : 1, sizeof(double));
    double *cp = (double *)calloc(nb ? nb : 1, sizeof(double));
    oracle_cov_add_intervals(txp, start, end, nnz, txp_len, bin_off, bins, tw);
    oracle_cov_binomial(n_txps, txp_len, bin_off, bins, tw, cp);
    oracle_cov_normalize(row_ptr, txp, start, end, n_reads, txp_len, bin_off, cp, bin_width, out);
    free(bin_off); free(bins); free(tw); free(cp);
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
