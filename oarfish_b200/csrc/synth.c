/*
 * synth.c -- seeded, deterministic generator of synthetic alignment stores
 * shaped like what oarfish's ingest (src/alignment_parser.rs ->
 * InMemoryAlignmentStore, src/util/oarfish_types.rs:547-558) hands to the EM.
 * Shared by the oracle-side and GPU-side runs so both see identical inputs
 * (SURVEY.md section 8d):
 *
 *   - transcripts are grouped into "genes" of size 1+Geom(mean 7); isoform
 *     ambiguity is local to a gene (contiguous transcript ids);
 *   - abundances ~ lognormal(0, 2), normalised;
 *   - per read: true transcript ~ abundance; k = min(1+Poisson(avg-1), 100)
 *     (--best-n default, prog_opts.rs:428); targets = true transcript + k-1
 *     distinct transcripts of the same gene, topped up from neighbouring genes
 *     when the gene is small; order shuffled;
 *   - prob = expf(-d/5) as f32 with d = 0 for the true hit and
 *     d ~ UniformInt[0,60] otherwise -- mirrors exp((score-best)/5) with the
 *     0.95 score threshold (oarfish_types.rs:1107-1118);
 *   - reads are stored in generation order (no sorting).
 *
 * Every read draws from its own counter-based stream hash(seed, read index),
 * so the output does not depend on the number of threads.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_BEST_N 100
#define MAX_WINDOW 2048

static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static inline double rng_unif(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n)
{ return (uint32_t)(((unsigned __int128)rng_next(r) * (unsigned __int128)n) >> 64); }

typedef struct {
    uint32_t n_txps, n_genes;
    uint32_t *gene_start;   /* n_genes + 1 */
    uint32_t *gene_of;      /* n_txps */
    double *cdf;            /* n_txps, inclusive cumulative abundance */
    double *abund;          /* n_txps, normalised */
} model_t;

static void model_free(model_t *m)
{ free(m->gene_start); free(m->gene_of); free(m->cdf); free(m->abund); }

static int model_build(model_t *m, uint32_t n_txps, uint64_t seed)
{
    memset(m, 0, sizeof(*m));
    m->n_txps = n_txps;
    m->gene_start = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_txps + 2));
    m->gene_of = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_txps + 1));
    m->cdf = (double *)malloc(sizeof(double) * ((size_t)n_txps + 1));
    m->abund = (double *)malloc(sizeof(double) * ((size_t)n_txps + 1));
    if (!m->gene_start || !m->gene_of || !m->cdf || !m->abund) return -1;
    rng_t r = { mix64(seed ^ 0xA5A5A5A55A5A5A5Aull) };
    const double log1mp = log(1.0 - 1.0 / 8.0);  /* Geom on {0,1,..} with mean 7 */
    uint32_t t = 0, g = 0;
    while (t < n_txps) {
        double u = rng_unif(&r); if (u <= 0.0) u = 1e-300;
        uint32_t size = 1u + (uint32_t)floor(log(u) / log1mp);
        if (size > n_txps - t) size = n_txps - t;
        m->gene_start[g] = t;
        for (uint32_t i = 0; i < size; ++i) m->gene_of[t + i] = g;
        t += size; ++g;
    }
    m->gene_start[g] = n_txps; m->n_genes = g;
    double tot = 0.0;
    for (uint32_t i = 0; i < n_txps; ++i) {      /* lognormal(0, 2) via Box-Muller */
        double u1 = rng_unif(&r), u2 = rng_unif(&r); if (u1 <= 0.0) u1 = 1e-300;
        double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        m->abund[i] = exp(2.0 * z); tot += m->abund[i];
    }
    double acc = 0.0;
    for (uint32_t i = 0; i < n_txps; ++i) { m->abund[i] /= tot; acc += m->abund[i]; m->cdf[i] = acc; }
    m->cdf[n_txps - 1] = 1.0;
    return 0;
}

static inline uint32_t sample_txp(const model_t *m, double u)
{
    uint32_t lo = 0, hi = m->n_txps - 1;
    while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (m->cdf[mid] < u) lo = mid + 1; else hi = mid; }
    return lo;
}

static inline uint32_t sample_poisson(rng_t *r, double lambda)
{   /* Knuth; lambda is small (avg-1 <= ~20) */
    if (lambda <= 0.0) return 0;
    double L = exp(-lambda), p = 1.0; uint32_t k = 0;
    do { ++k; p *= rng_unif(r); } while (p > L && k < 1000);
    return k - 1;
}

/* draws the read's true transcript and its number of alignments */
static inline void read_head(const model_t *m, rng_t *r, double avg_aln, uint32_t *true_t, uint32_t *k)
{
    *true_t = sample_txp(m, rng_unif(r));
    uint32_t kk = 1u + sample_poisson(r, avg_aln - 1.0);
    if (kk > MAX_BEST_N) kk = MAX_BEST_N;
    if (kk > m->n_txps) kk = m->n_txps;
    *k = kk;
}

static inline rng_t read_rng(uint64_t seed, uint64_t read)
{ rng_t r = { mix64(mix64(seed) + read * 0xD6E8FEB86659FD93ull) }; return r; }

/* Pass 1: row_ptr[0..n_reads] (exclusive prefix of per-read alignment counts).
 * Returns 0 on success. */
int oar_synth_plan(uint64_t n_reads, uint32_t n_txps, double avg_aln, uint64_t seed, uint64_t *row_ptr)
{
    if (n_txps == 0) return -1;
    model_t m; if (model_build(&m, n_txps, seed)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n_reads; ++i) {
        rng_t r = read_rng(seed, (uint64_t)i);
        uint32_t t, k; read_head(&m, &r, avg_aln, &t, &k);
        row_ptr[i + 1] = k;
    }
    row_ptr[0] = 0;
    for (uint64_t i = 0; i < n_reads; ++i) row_ptr[i + 1] += row_ptr[i];
    model_free(&m);
    return 0;
}

/* Pass 2: fill txp/prob given row_ptr from oar_synth_plan.  `true_txp` (n_reads)
 * and `abund` (n_txps) are optional outputs. */
int oar_synth_fill(uint64_t n_reads, uint32_t n_txps, double avg_aln, uint64_t seed,
                   const uint64_t *row_ptr, uint32_t *txp, float *prob,
                   uint32_t *true_txp, double *abund)
{
    if (n_txps == 0) return -1;
    model_t m; if (model_build(&m, n_txps, seed)) return -1;
    float ptab[61];
    for (int d = 0; d <= 60; ++d) ptab[d] = expf(-(float)d / 5.0f);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < (int64_t)n_reads; ++i) {
        rng_t r = read_rng(seed, (uint64_t)i);
        uint32_t t, k; read_head(&m, &r, avg_aln, &t, &k);
        uint64_t base = row_ptr[i];
        if (row_ptr[i + 1] - base != k) { bad |= 1; continue; }
        if (true_txp) true_txp[i] = t;
        /* candidate window: the gene, extended over following (then preceding)
         * genes until it holds at least k transcripts */
        uint32_t g = m.gene_of[t];
        uint32_t ws = m.gene_start[g], we = m.gene_start[g + 1];
        uint32_t gl = g, gr = g + 1;
        while (we - ws < k) {
            if (gr < m.n_genes) { ++gr; we = m.gene_start[gr]; }
            else if (gl > 0) { --gl; ws = m.gene_start[gl]; }
            else break;
        }
        uint32_t W = we - ws; if (W > MAX_WINDOW) { W = MAX_WINDOW; if (t >= ws + W) ws = t - W + 1; }
        uint32_t cand[MAX_WINDOW];
        for (uint32_t c = 0; c < W; ++c) cand[c] = ws + c;
        /* move the true transcript to slot 0, then partial Fisher-Yates for k-1 more */
        { uint32_t pos = t - ws; uint32_t tmp = cand[0]; cand[0] = cand[pos]; cand[pos] = tmp; }
        for (uint32_t c = 1; c < k; ++c) {
            uint32_t j = c + rng_below(&r, W - c);
            uint32_t tmp = cand[c]; cand[c] = cand[j]; cand[j] = tmp;
        }
        float pr[MAX_BEST_N];
        pr[0] = ptab[0];
        for (uint32_t c = 1; c < k; ++c) pr[c] = ptab[rng_below(&r, 61)];
        /* shuffle the order of the k alignments */
        for (uint32_t c = k; c > 1; --c) {
            uint32_t j = rng_below(&r, c);
            uint32_t tt = cand[c - 1]; cand[c - 1] = cand[j]; cand[j] = tt;
            float tp = pr[c - 1]; pr[c - 1] = pr[j]; pr[j] = tp;
        }
        for (uint32_t c = 0; c < k; ++c) { txp[base + c] = cand[c]; prob[base + c] = pr[c]; }
    }
    if (abund) memcpy(abund, m.abund, sizeof(double) * n_txps);
    model_free(&m);
    return bad ? -2 : 0;
}
