/*
 * oracle/em_par_port.c -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Multi-threaded CPU port of the reference's *timed* code paths, laid out in
 * memory the way the Rust code lays them out, so that bench.py can report a
 * CPU baseline ("restated reference (C), not the Rust binary") next to the GPU
 * numbers.  Only bench.py (cpu_baseline / --impl reference) and tests/ load it.
 *
 *   port_em_par      src/em.rs:320-447 (em_par) + :22-79 (m_step_par):
 *                    AoS 24-byte AlnInfo (oarfish_types.rs:330-337) + f32 prob +
 *                    f64 coverage arrays, a per-read slice-triple array
 *                    (em.rs:335, 48 B/read), CAS-loop f64 atomic adds standing
 *                    in for atomic_float::AtomicF64::fetch_add (em.rs:74),
 *                    SERIAL rel-diff scan (em.rs:379-386), parallel zeroing
 *                    (em.rs:392-394), stop rule niter > 1 (em.rs:399).
 *                    OpenMP dynamic chunks stand in for rayon's work stealing.
 *   port_bootstrap   src/em.rs:292-314 (bootstrap) + :273-290 (do_bootstrap):
 *                    T concurrent *sequential* do_em runs (em.rs:144-255, stop
 *                    rule niter > 50), each preceded by N uniform draws + sort
 *                    (bootstrap.rs:7-16) and iterating rows through the index
 *                    list (oarfish_types.rs:571-598).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MIN_READ_THRESH 1e-5
#define EM_DENOM_THRESH 1e-30

typedef struct {           /* AlnInfo, oarfish_types.rs:330-337 (24 bytes) */
    double prob;           /* always 0.0 (oarfish_types.rs:352), unused by EM */
    uint32_t ref_id;
    uint32_t start;
    uint32_t end;
    uint8_t strand;
} aln_info_t;

typedef struct {           /* EqIterateT, em.rs:14: three fat slices = 48 B */
    const aln_info_t *alns; uint64_t n_alns;
    const float *probs;     uint64_t n_probs;
    const double *covs;     uint64_t n_covs;
} eq_iterate_t;

typedef struct {
    uint64_t n_reads, nnz;
    uint32_t n_txps;
    aln_info_t *alns;
    float *probs;
    double *covs;
    uint64_t *boundaries;
    eq_iterate_t *iterates;
} port_store_t;

int port_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py sets the pool size explicitly
   (the rayon pool of em_par / bootstrap is sized by the caller as well: em.rs:324-327, :299-302). */
void port_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

port_store_t *port_store_create(const uint64_t *row_ptr, const uint32_t *txp, const float *prob,
                                const double *cov_or_null, uint64_t n_reads, uint64_t nnz,
                                uint32_t n_txps)
{
    port_store_t *s = (port_store_t *)calloc(1, sizeof(port_store_t));
    s->n_reads = n_reads; s->nnz = nnz; s->n_txps = n_txps;
    s->alns = (aln_info_t *)malloc(sizeof(aln_info_t) * (nnz ? nnz : 1));
    s->probs = (float *)malloc(sizeof(float) * (nnz ? nnz : 1));
    s->covs = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    s->boundaries = (uint64_t *)malloc(sizeof(uint64_t) * (n_reads + 1));
    s->iterates = (eq_iterate_t *)malloc(sizeof(eq_iterate_t) * (n_reads ? n_reads : 1));
    memcpy(s->boundaries, row_ptr, sizeof(uint64_t) * (n_reads + 1));
#pragma omp parallel for schedule(static)
    for (uint64_t j = 0; j < nnz; ++j) {
        s->alns[j].prob = 0.0; s->alns[j].ref_id = txp[j];
        s->alns[j].start = 0; s->alns[j].end = 1000; s->alns[j].strand = 0;
        s->probs[j] = prob[j];
        s->covs[j] = cov_or_null ? cov_or_null[j] : 0.0; /* zeros unless --model-coverage, oarfish_types.rs:731 */
    }
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < n_reads; ++r) {              /* em.rs:335 collect() */
        uint64_t b = row_ptr[r], n = row_ptr[r + 1] - b;
        s->iterates[r].alns = s->alns + b;   s->iterates[r].n_alns = n;
        s->iterates[r].probs = s->probs + b; s->iterates[r].n_probs = n;
        s->iterates[r].covs = s->covs + b;   s->iterates[r].n_covs = n;
    }
    return s;
}

void port_store_destroy(port_store_t *s)
{
    if (!s) return;
    free(s->alns); free(s->probs); free(s->covs); free(s->boundaries); free(s->iterates); free(s);
}

static inline void atomic_f64_add(double *addr, double v)
{   /* AtomicF64::fetch_add == compare_exchange loop on the u64 bits */
    uint64_t *p = (uint64_t *)addr;
    uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED), neu;
    do {
        double d; memcpy(&d, &old, 8); d += v; memcpy(&neu, &d, 8);
    } while (!__atomic_compare_exchange_n(p, &old, neu, 1, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED));
}

static void m_step_par(const port_store_t *s, int model_coverage, const double *prev, double *curr)
{   /* em.rs:22-79 */
    const eq_iterate_t *it = s->iterates;
    const int64_t n = (int64_t)s->n_reads;
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t r = 0; r < n; ++r) {
        const aln_info_t *a = it[r].alns; const float *p = it[r].probs; const double *c = it[r].covs;
        uint64_t m = it[r].n_alns;
        double denom = 0.0;
        for (uint64_t j = 0; j < m; ++j) {
            double cp = model_coverage ? c[j] : 1.0;
            denom += prev[a[j].ref_id] * (double)p[j] * cp * 1.0; /* Relaxed load, em.rs:51 */
        }
        if (denom > EM_DENOM_THRESH) {
            for (uint64_t j = 0; j < m; ++j) {
                double cp = model_coverage ? c[j] : 1.0;
                double inc = (prev[a[j].ref_id] * (double)p[j] * cp * 1.0) / denom;
                atomic_f64_add(&curr[a[j].ref_id], inc);
            }
        }
    }
}

/* em_par, em.rs:320-447.  `max_sweeps_budget` (0 = unlimited) lets bench.py
 * time a bounded number of iterations of the same loop. Returns loop sweeps. */
uint32_t port_em_par(const port_store_t *s, int model_coverage, const double *init,
                     uint32_t max_iter, double conv_thresh, double *out_counts,
                     uint32_t *out_niter, double *out_rel_diff)
{
    uint32_t M = s->n_txps;
    double *prev = (double *)malloc(sizeof(double) * (M ? M : 1));
    double *curr = (double *)calloc(M ? M : 1, sizeof(double));
    if (init) memcpy(prev, init, sizeof(double) * M);
    else { double avg = (double)s->n_reads / (double)M; for (uint32_t i = 0; i < M; ++i) prev[i] = avg; }
    double rel_diff = 0.0; uint32_t niter = 0, sweeps = 0;
    while (niter < max_iter) {
        m_step_par(s, model_coverage, prev, curr);
        ++sweeps;
        for (uint32_t i = 0; i < M; ++i) {               /* serial scan, em.rs:379-386 */
            if (prev[i] > MIN_READ_THRESH) {
                double rd = (curr[i] - prev[i]) / prev[i];
                rel_diff = rd > rel_diff ? rd : rel_diff;
            }
        }
        double *t = prev; prev = curr; curr = t;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)M; ++i) curr[i] = 0.0;  /* em.rs:392-394 */
        if (rel_diff < conv_thresh && niter > 1) break;  /* em.rs:399 */
        niter += 1;
        if (niter < max_iter) rel_diff = 0.0;
    }
    for (uint32_t i = 0; i < M; ++i) if (prev[i] < MIN_READ_THRESH) prev[i] = 0.0; /* em.rs:424-429 */
    m_step_par(s, model_coverage, prev, curr);           /* em.rs:433-440 */
    memcpy(out_counts, curr, sizeof(double) * M);
    if (out_niter) *out_niter = niter;
    if (out_rel_diff) *out_rel_diff = rel_diff;
    free(prev); free(curr);
    return sweeps;
}

/* ---- bootstrap ---------------------------------------------------------- */

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* LSD radix sort (stands in for sort_unstable / pdqsort, bootstrap.rs:14). */
static void radix_sort_u64(uint64_t *a, uint64_t *tmp, uint64_t n, uint64_t max_val)
{
    int passes = 0; while (passes < 8 && (max_val >> (8 * passes)) != 0) ++passes;
    if (passes == 0) passes = 1;
    uint64_t *src = a, *dst = tmp;
    for (int p = 0; p < passes; ++p) {
        uint64_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        int sh = 8 * p;
        for (uint64_t i = 0; i < n; ++i) cnt[((src[i] >> sh) & 0xFF) + 1]++;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & 0xFF]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, sizeof(uint64_t) * n);
}

static void m_step_seq_inds(const port_store_t *s, int model_coverage, const uint64_t *inds,
                            uint64_t n_inds, const double *prev, double *curr)
{   /* em.rs:87-133 over random_sampling_iter (oarfish_types.rs:571-598) */
    for (uint64_t k = 0; k < n_inds; ++k) {
        uint64_t r = inds[k];
        uint64_t b = s->boundaries[r], e = s->boundaries[r + 1];
        const aln_info_t *a = s->alns + b; const float *p = s->probs + b; const double *c = s->covs + b;
        uint64_t m = e - b;
        double denom = 0.0;
        for (uint64_t j = 0; j < m; ++j) {
            double cp = model_coverage ? c[j] : 1.0;
            denom += prev[a[j].ref_id] * (double)p[j] * cp * 1.0;
        }
        if (denom > EM_DENOM_THRESH) {
            for (uint64_t j = 0; j < m; ++j) {
                double cp = model_coverage ? c[j] : 1.0;
                curr[a[j].ref_id] += (prev[a[j].ref_id] * (double)p[j] * cp * 1.0) / denom;
            }
        }
    }
}

static uint32_t do_em_inds(const port_store_t *s, int model_coverage, const uint64_t *inds,
                           uint64_t n_inds, uint32_t max_iter, double conv_thresh,
                           double *out_counts)
{   /* em.rs:144-255 */
    uint32_t M = s->n_txps;
    double *prev = (double *)malloc(sizeof(double) * (M ? M : 1));
    double *curr = (double *)calloc(M ? M : 1, sizeof(double));
    double avg = (double)s->n_reads / (double)M;
    for (uint32_t i = 0; i < M; ++i) prev[i] = avg;
    double rel_diff = 0.0; uint32_t niter = 0;
    while (niter < max_iter) {
        m_step_seq_inds(s, model_coverage, inds, n_inds, prev, curr);
        for (uint32_t i = 0; i < M; ++i)
            if (prev[i] > MIN_READ_THRESH) {
                double rd = (curr[i] - prev[i]) / prev[i];
                rel_diff = rd > rel_diff ? rd : rel_diff;
            }
        double *t = prev; prev = curr; curr = t;
        memset(curr, 0, sizeof(double) * M);
        if (rel_diff < conv_thresh && niter > 50) break;
        niter += 1; rel_diff = 0.0;
    }
    for (uint32_t i = 0; i < M; ++i) if (prev[i] < MIN_READ_THRESH) prev[i] = 0.0;
    m_step_seq_inds(s, model_coverage, inds, n_inds, prev, curr);
    memcpy(out_counts, curr, sizeof(double) * M);
    free(prev); free(curr);
    return niter;
}

/* bootstrap, em.rs:292-314: num_boot replicates on a pool of nthreads. */
static void port_bootstrap_impl(const port_store_t *s, int model_coverage, uint32_t num_boot, uint64_t seed,
                                uint32_t max_iter, double conv_thresh, int nthreads,
                                double *out, uint32_t *out_niter, double *out_em_seconds);

void port_bootstrap(const port_store_t *s, int model_coverage, uint32_t num_boot, uint64_t seed,
                    uint32_t max_iter, double conv_thresh, int nthreads,
                    double *out /* num_boot x M */, uint32_t *out_niter /* num_boot or NULL */)
{ port_bootstrap_impl(s, model_coverage, num_boot, seed, max_iter, conv_thresh, nthreads, out, out_niter, NULL); }

/* Same, and reports per replicate the seconds spent in do_em alone (without drawing and sorting the sample): a bounded
   benchmark sample caps max_iter, which would otherwise overweight the per-replicate set-up a full run amortises
   over hundreds of sweeps. */
void port_bootstrap_timed(const port_store_t *s, int model_coverage, uint32_t num_boot, uint64_t seed,
                          uint32_t max_iter, double conv_thresh, int nthreads,
                          double *out, uint32_t *out_niter, double *out_em_seconds /* num_boot */)
{ port_bootstrap_impl(s, model_coverage, num_boot, seed, max_iter, conv_thresh, nthreads, out, out_niter, out_em_seconds); }

static void port_bootstrap_impl(const port_store_t *s, int model_coverage, uint32_t num_boot, uint64_t seed,
                                uint32_t max_iter, double conv_thresh, int nthreads,
                                double *out, uint32_t *out_niter, double *out_em_seconds)
{
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t b = 0; b < (int64_t)num_boot; ++b) {
        uint64_t n = s->n_reads;
        uint64_t *inds = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
        uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
        uint64_t st = seed * 0x9E3779B97F4A7C15ull + (uint64_t)b * 0xD1342543DE82EF95ull + 1;
        for (uint64_t i = 0; i < n; ++i) {               /* bootstrap.rs:8-13 */
            unsigned __int128 m = (unsigned __int128)splitmix64(&st) * (unsigned __int128)n;
            inds[i] = (uint64_t)(m >> 64);
        }
        radix_sort_u64(inds, tmp, n, n ? n - 1 : 0);     /* bootstrap.rs:14 */
        const double t0 = omp_get_wtime();
        uint32_t it = do_em_inds(s, model_coverage, inds, n, max_iter, conv_thresh,
                                 out + (uint64_t)b * s->n_txps);
        if (out_em_seconds) out_em_seconds[b] = omp_get_wtime() - t0;
        if (out_niter) out_niter[b] = it;
        free(inds); free(tmp);
    }
}
