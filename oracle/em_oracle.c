/*
 * oracle/em_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Sequential, bit-reproducible CPU restatement of the oarfish EM hot path
 * (reference: /root/reference @ v0.10.3).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the reference ships no test, golden vector or fixture that
 * touches src/em.rs or src/bootstrap.rs, and the Rust
 * toolchain is absent so the reference itself cannot be run here.  What pins
 * this restatement instead: analytic known-answer tests and an independent
 * pure-Python restatement (tests/test_oracle_kat.py, tests/pyref.py).
 *
 * Functions restated (file:line under /root/reference/):
 *   oracle_m_step          src/em.rs:87-133      (m_step)
 *   oracle_do_em           src/em.rs:144-255     (do_em; min_iter=50)
 *                          src/em.rs:320-447     (em_par stop rule; min_iter=1)
 *   oracle_get_sample_inds src/bootstrap.rs:7-16 (get_sample_inds)
 *   row iteration          src/util/oarfish_types.rs:571-669 (iter / random_sampling_iter)
 *   constants              src/util/constants.rs:1-2
 *
 * The KDE density factor (em.rs:173-178) comes from the un-vendored `kders`
 * crate; it is iteration-invariant and is folded by the caller into `cov`
 * (the per-alignment f64 factor), so dens_prob == 1.0 here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MIN_READ_THRESH 1e-5  /* constants.rs:1 */
#define EM_DENOM_THRESH 1e-30 /* constants.rs:2 */

/* One E+M sweep over `n_iter_rows` rows.  If `inds` is NULL the rows are
 * 0..n_iter_rows-1 (store.iter(), oarfish_types.rs:602-633); otherwise row k is
 * inds[k], repeats included (random_sampling_iter, oarfish_types.rs:571-598).
 * `cov` may be NULL => model_coverage == false => factor 1.0 (em.rs:108).
 * `wts` (may be NULL) is NOT in the reference: an integer multiplicity per row,
 * applied by visiting the row wts[r] times -- it exists to prove that the
 * weight formulation used on the GPU equals the index-list formulation. */
void oracle_m_step(const uint64_t *row_ptr, const uint32_t *txp, const float *prob,
                   const double *cov, const uint64_t *inds, uint64_t n_iter_rows,
                   const uint32_t *wts, const double *prev, double *curr)
{
    for (uint64_t k = 0; k < n_iter_rows; ++k) {
        uint64_t r = inds ? inds[k] : k;
        uint32_t reps = wts ? wts[r] : 1u;
        uint64_t s = row_ptr[r], e = row_ptr[r + 1];
        for (uint32_t rep = 0; rep < reps; ++rep) {
            double denom = 0.0; /* em.rs:98 */
            for (uint64_t j = s; j < e; ++j) {
                double p = (double)prob[j];              /* em.rs:107 */
                double cp = cov ? cov[j] : 1.0;          /* em.rs:108 */
                double dens = 1.0;                       /* em.rs:109 */
                denom += prev[txp[j]] * p * cp * dens;   /* em.rs:111 */
            }
            if (denom > EM_DENOM_THRESH) {               /* em.rs:115 */
                for (uint64_t j = s; j < e; ++j) {
                    double p = (double)prob[j];
                    double cp = cov ? cov[j] : 1.0;
                    double dens = 1.0;
                    double inc = (prev[txp[j]] * p * cp * dens) / denom; /* em.rs:128 */
                    curr[txp[j]] += inc;                 /* em.rs:129 */
                }
            }
        }
    }
}

/* do_em (em.rs:144-255).  `min_iter` = 50 reproduces do_em's stop rule
 * (em.rs:212); `min_iter` = 1 reproduces em_par's (em.rs:399).  `n_reads` is
 * eq_map.num_aligned_reads() (em.rs:154), used for the uniform init even when
 * an index list is given.  Returns the number of m_step calls made inside the
 * loop (i.e. excluding the final extra one); *out_niter is the loop variable
 * `niter` at exit, *out_rel_diff the last rel_diff evaluated. */
uint32_t oracle_do_em(const uint64_t *row_ptr, const uint32_t *txp, const float *prob,
                      const double *cov, uint64_t n_reads, uint32_t n_txps,
                      const double *init, uint32_t max_iter, double conv_thresh,
                      uint32_t min_iter, const uint64_t *inds, uint64_t n_inds,
                      const uint32_t *wts, double *out_counts, uint32_t *out_niter,
                      double *out_rel_diff)
{
    uint64_t n_iter_rows = inds ? n_inds : n_reads;
    double total_weight = (double)n_reads;                     /* em.rs:154 */
    double *prev = (double *)malloc(sizeof(double) * (n_txps ? n_txps : 1));
    double *curr = (double *)calloc(n_txps ? n_txps : 1, sizeof(double)); /* em.rs:158 */
    if (init) {
        memcpy(prev, init, sizeof(double) * n_txps);           /* em.rs:162 */
    } else {
        double avg = total_weight / (double)n_txps;            /* em.rs:165 */
        for (uint32_t i = 0; i < n_txps; ++i) prev[i] = avg;
    }
    double rel_diff = 0.0;
    uint32_t niter = 0, sweeps = 0;
    while (niter < max_iter) {                                 /* em.rs:181 */
        oracle_m_step(row_ptr, txp, prob, cov, inds, n_iter_rows, wts, prev, curr);
        ++sweeps;
        for (uint32_t i = 0; i < n_txps; ++i) {                /* em.rs:194-201 */
            if (prev[i] > MIN_READ_THRESH) {
                double cc = curr[i], pc = prev[i];
                double rd = (cc - pc) / pc;                    /* signed */
                /* f64::max: NaN-ignoring; rd is never NaN for finite inputs */
                rel_diff = (rd > rel_diff) ? rd : rel_diff;
            }
        }
        double *t = prev; prev = curr; curr = t;               /* em.rs:204 */
        memset(curr, 0, sizeof(double) * n_txps);              /* em.rs:207 */
        if (rel_diff < conv_thresh && niter > min_iter) break; /* em.rs:212 / :399 */
        niter += 1;                                            /* em.rs:218 */
        if (niter < max_iter) rel_diff = 0.0;                  /* em.rs:234 (kept for reporting on the last pass) */
    }
    for (uint32_t i = 0; i < n_txps; ++i)                      /* em.rs:238-242 */
        if (prev[i] < MIN_READ_THRESH) prev[i] = 0.0;
    oracle_m_step(row_ptr, txp, prob, cov, inds, n_iter_rows, wts, prev, curr); /* em.rs:245 */
    memcpy(out_counts, curr, sizeof(double) * n_txps);
    if (out_niter) *out_niter = niter;
    if (out_rel_diff) *out_rel_diff = rel_diff;
    free(prev); free(curr);
    return sweeps;
}

/* splitmix64: the oracle's own seeded generator.  The reference draws from the
 * unseeded thread RNG (em.rs:274), so no seed-level parity exists by design. */
static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* get_sample_inds (bootstrap.rs:7-16): n uniform draws from [0,n) with
 * replacement, returned sorted. */
void oracle_get_sample_inds(uint64_t n, uint64_t seed, uint64_t *out)
{
    uint64_t s = seed;
    for (uint64_t i = 0; i < n; ++i) {
        /* unbiased enough for a test oracle: 128-bit multiply-shift */
        unsigned __int128 m = (unsigned __int128)splitmix64(&s) * (unsigned __int128)n;
        out[i] = (uint64_t)(m >> 64);
    }
    qsort(out, n, sizeof(uint64_t), cmp_u64); /* bootstrap.rs:14 sort_unstable */
}

/* Histogram an index list into per-row multiplicities (test helper). */
void oracle_inds_to_weights(const uint64_t *inds, uint64_t n_inds, uint64_t n_rows, uint32_t *wts)
{
    memset(wts, 0, sizeof(uint32_t) * n_rows);
    for (uint64_t i = 0; i < n_inds; ++i) wts[inds[i]] += 1u;
}

/* Read-level assignment probabilities: the inner loop of write_out_prob
 * (src/util/write_function.rs:283-332).  For every read: denom over its
 * alignments with the FINAL counts (:286-291; note the reference leaves the KDE
 * factor out here), nprob = clamp(c*p*cov/denom, 0, 1) (:307), kept iff
 * nprob >= display_thresh (:309), kept values renormalised by their sum
 * (:316-318).  out[j] = renormalised probability, 0 for dropped alignments;
 * kept[r] = number of alignments kept for read r. */
void oracle_posteriors(const uint64_t *row_ptr, const uint32_t *txp, const float *prob,
                       const double *cov, uint64_t n_reads, const double *counts,
                       double display_thresh, double *out, uint32_t *kept)
{
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t s = row_ptr[r], e = row_ptr[r + 1];
        double denom = 0.0;
        for (uint64_t j = s; j < e; ++j)
            denom += counts[txp[j]] * (double)prob[j] * (cov ? cov[j] : 1.0);
        double denom2 = 0.0;
        uint32_t k = 0;
        for (uint64_t j = s; j < e; ++j) {
            double np = (counts[txp[j]] * (double)prob[j] * (cov ? cov[j] : 1.0)) / denom;
            /* f64::clamp propagates NaN; NaN >= thresh is false */
            if (np < 0.0) np = 0.0; else if (np > 1.0) np = 1.0;
            if (np >= display_thresh) { out[j] = np; denom2 += np; ++k; } else out[j] = 0.0;
        }
        for (uint64_t j = s; j < e; ++j) if (out[j] != 0.0 || (display_thresh <= 0.0 && k)) out[j] /= denom2;
        if (kept) kept[r] = k;
    }
}

/* get_aux_counts (src/util/aux_counts.rs:23-50): per transcript, the number of
 * alignments (total) and the number of those from single-alignment reads. */
void oracle_aux_counts(const uint64_t *row_ptr, const uint32_t *txp, uint64_t n_reads,
                       uint32_t n_txps, uint32_t *unique, uint32_t *total)
{
    memset(unique, 0, sizeof(uint32_t) * n_txps);
    memset(total, 0, sizeof(uint32_t) * n_txps);
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t s = row_ptr[r], e = row_ptr[r + 1];
        int is_unique = (e - s) == 1;
        for (uint64_t j = s; j < e; ++j) {
            if (txp[j] < n_txps) { total[txp[j]] += 1; if (is_unique) unique[txp[j]] += 1; }
        }
    }
}
