/*
 * oarfish_em.h -- C ABI of the B200-native EM / bootstrap engine.
 *
 * This is the drop-in boundary for oarfish's inference stage.  The reference
 * (100 % Rust, /root/reference @ v0.10.3) has no FFI of its own; the boundary is
 * the three Rust functions the drivers call, and each entry point below names
 * the one it replaces:
 *
 *   oar_store_create      <- the data EMInfo.eq_map points at:
 *                            InMemoryAlignmentStore (src/util/oarfish_types.rs:547-558),
 *                            read through iter() (:602-656)
 *   oar_em                <- em::em      (src/em.rs:262, called bulk.rs:157-158, single_cell.rs:150)
 *                            em::em_par  (src/em.rs:320, called bulk.rs:155-156)
 *   oar_bootstrap         <- em::bootstrap (src/em.rs:292, called bulk.rs:179)
 *                            + bootstrap::get_sample_inds (src/bootstrap.rs:7)
 *   oar_bootstrap_weights <- em::do_bootstrap (src/em.rs:273) with the resampling
 *                            made explicit (test hook: exact per-weight-vector parity)
 *   oar_em_batched        <- the per-cell em::em(&emi, 1) calls of
 *                            single_cell.rs:150 batched into one call
 *   oar_multi_*           <- the internal fan-out of em::bootstrap / the single-cell driver over workers
 *   oar_store_create_filtered <- AlignmentFilters::filter (src/util/oarfish_types.rs:955-1130)
 *   oar_store_coverage_model[_binomial] <- logistic_prob / binomial_continuous_prob + normalize_read_probs
 *
 * Conventions
 *   - every function returns 0 on success or a negative oar_status; nothing
 *     throws or aborts across the ABI; oar_last_error() returns a thread-local
 *     message for the last failing call on this thread.
 *   - buffers passed in are BORROWED for the duration of the call only.
 *     Pointers documented "host or device" are resolved with CUDA unified
 *     addressing (cudaMemcpyDefault).
 *   - a handle is used by one host thread at a time; distinct handles are
 *     independent and may be created, used and destroyed concurrently from
 *     different host threads (the single-cell driver calls em::em from a pool
 *     of workers, single_cell.rs:91-193).  Device memory is owned by the handle.
 *   - transcript indexing is bit-exact: out[i] is the count of header reference
 *     id i, exactly as em::em returns Vec<f64> indexed by ref_id.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with OAR_ERR_CUDA.
 */
#ifndef OARFISH_EM_H
#define OARFISH_EM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    OAR_OK = 0,
    OAR_ERR_INVALID = -1,   /* bad argument (null pointer, inconsistent sizes, id out of range) */
    OAR_ERR_CUDA = -2,      /* CUDA runtime error or no device */
    OAR_ERR_OOM = -3,       /* host or device allocation failed */
    OAR_ERR_UNSUPPORTED = -4
} oar_status;

typedef struct oar_store oar_store; /* opaque; one per alignment store per device */

/* Sweep-kernel selection (OAR_KERNEL_AUTO picks the fastest valid one). */
typedef enum {
    OAR_KERNEL_AUTO = 0,
    OAR_KERNEL_ROWGROUP = 1,  /* 8-lane group per read row, global f64 reductions */
    OAR_KERNEL_TILED = 2      /* locality-sorted tiles of warp-chunks (4 slots per lane), in-tile aggregation
                                 (default) */
} oar_kernel;

/* ABI version: major*1000 + minor. */
int oar_version(void);

/* Number of visible CUDA devices, or a negative oar_status. */
int oar_device_count(void);

const char *oar_last_error(void);

/*
 * Upload an alignment store to `device` as CSR.
 *   row_ptr  N+1 u64  == InMemoryAlignmentStore.boundaries (private Vec<usize>, oarfish_types.rs:555)
 *   txp_id   nnz u32  == AlnInfo.ref_id de-interleaved (oarfish_types.rs:331)
 *   prob     nnz f32  == as_probabilities (oarfish_types.rs:551)
 *   aux      nnz f64 or NULL == coverage_probabilities * density factor when
 *            --model-coverage / --use-kde is on (em.rs:108-111); NULL means 1.0
 * All four may be host or device pointers.  Validates row_ptr monotonicity,
 * row_ptr[N] == nnz and txp_id < n_txps (OAR_ERR_INVALID otherwise).
 */
int oar_store_create(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob,
                     const double *aux_or_null, uint64_t n_reads, uint64_t nnz,
                     uint32_t n_txps, int device, oar_store **out);

void oar_store_destroy(oar_store *store);

/* Store geometry. */
int oar_store_info(const oar_store *store, uint64_t *n_reads, uint64_t *nnz, uint32_t *n_txps,
                   int *device);

/* Select the sweep kernel for subsequent calls on this store. */
int oar_store_set_kernel(oar_store *store, int kernel);

/*
 * One EM run to convergence.  Replaces em::em (min_iter = 50, stop rule
 * em.rs:212) and em::em_par (min_iter = 1, stop rule em.rs:399).
 *   init_or_null  M f64 (EMInfo.init_abundances, em.rs:160-167) or NULL for the
 *                 uniform N/M start
 *   out_counts    M f64, host or device, caller-allocated
 *   out_niter     loop counter `niter` at exit (em.rs:170)
 *   out_rel_diff  last relative difference evaluated (em.rs:194-201)
 */
int oar_em(oar_store *store, const double *init_or_null, uint32_t max_iter, double conv_thresh,
           uint32_t min_iter, double *out_counts, uint32_t *out_niter, double *out_rel_diff);

/*
 * Progress of a running EM, the counterpart of the reference's log lines ("iteration {niter}; rel diff {rel_diff}",
 * em.rs:219-233: every 10 iterations up to 100, then every 100).  The convergence test runs on the device and the
 * host looks at it once per batch of 16 iterations, so `fn` is called from the calling thread of oar_em /
 * oar_bootstrap with the state after every such batch: niter (the loop counter) and the last rel_diff evaluated.
 * NULL removes the callback.
 */
typedef void (*oar_progress_fn)(uint32_t niter, double rel_diff, void *user);
int oar_store_set_progress(oar_store *store, oar_progress_fn fn, void *user);

/*
 * Bootstrap replicates (em::bootstrap, em.rs:292-314).  Replicate b of this
 * call is global replicate  g = first_replicate + b * replicate_stride ; its
 * resampling weights are a pure function of (seed, g), so sharding replicates
 * over devices or processes (rank r of G: first = r, stride = G) gives results
 * independent of G.  Each replicate: multinomial(N; 1/N..) read weights
 * (== the histogram of get_sample_inds, bootstrap.rs:7-16), then do_em with
 * min_iter = 50 (em.rs:287-289).
 *   out        num_boot x M f64 row-major, host or device
 *   out_niter  num_boot u32 (host) or NULL
 */
int oar_bootstrap(oar_store *store, uint32_t num_boot, uint64_t seed, uint32_t first_replicate,
                  uint32_t replicate_stride, uint32_t max_iter, double conv_thresh,
                  double *out, uint32_t *out_niter);

/* Same, with caller-supplied integer read weights (R x N u32, host or device):
 * weights[r*N + i] = number of times read i appears in replicate r's sample.  The tiled sweep carries
 * a weight as 16 bits: a weight above 65535 fails with OAR_ERR_UNSUPPORTED (OAR_KERNEL_ROWGROUP has
 * no such limit). */
int oar_bootstrap_weights(oar_store *store, const uint32_t *weights, uint32_t n_replicates,
                          uint32_t max_iter, double conv_thresh, uint32_t min_iter,
                          double *out, uint32_t *out_niter);

/* The weights oar_bootstrap uses for global replicate g (N u32, host or device). */
int oar_bootstrap_sample_weights(oar_store *store, uint64_t seed, uint32_t replicate,
                                 uint32_t *out_weights);

/*
 * Batched independent EMs over contiguous groups of reads of one store
 * (single-cell mode, single_cell.rs:91-193: one em::em(&emi,1) per cell, i.e.
 * do_em with the full transcriptome as parameter space and the N_cell/M start).
 *   cell_row_ptr  C+1 u64 over reads: cell c owns reads [cell_row_ptr[c], cell_row_ptr[c+1])
 * Output is sparse, CSR over cells (single_cell.rs:155-160 keeps only v > 0):
 *   out_cell_ptr  C+1 u64, host or device
 *   out_txp       transcript ids of cell c at [out_cell_ptr[c], out_cell_ptr[c+1]), ascending:
 *                 every transcript with an alignment in the cell (all others are exactly 0)
 *   out_val       their counts (f64; may be 0)
 *   capacity      elements available in out_txp / out_val; the store's nnz is always enough.
 *                 If too small: OAR_ERR_INVALID, *out_nnz = required size, out_cell_ptr filled.
 *   out_niter     C u32 (host or device) or NULL
 */
int oar_em_batched(oar_store *store, const uint64_t *cell_row_ptr, uint32_t n_cells,
                   uint32_t max_iter, double conv_thresh, uint32_t min_iter,
                   uint64_t *out_cell_ptr, uint32_t *out_txp, double *out_val, uint64_t capacity,
                   uint64_t *out_nnz, uint32_t *out_niter);

/*
 * ---- store construction from alignment records: the filters and the score -> probability formula ----------------
 * AlignmentFilters::filter (src/util/oarfish_types.rs:955-1130) applied to every read's group of records on the
 * device; BAM decoding and grouping by read name (alignment_parser.rs:301-437) stay with the caller, who passes the
 * fields filter() reads through AlnRecordLike (oarfish_types.rs:186-202) as columns, one entry per record:
 *   group_ptr  G+1 u64   records of read g are [group_ptr[g], group_ptr[g+1])
 *   ref_id, aln_start, aln_end, aln_span   u32 (aln_span: reference span from the CIGAR)
 *   score      i32       the AS tag (a missing tag is i32::MIN, :993)
 *   flags      u8        OAR_REC_* bits
 *   seq_len    u32       length of the record's sequence, 0 if the record carries none (:981-984 takes the first non-zero)
 *   txp_len    M u32     TranscriptInfo.len
 * Groups without a retained alignment are dropped (add_filtered_group, :718-738).  The result is a store exactly as
 * oar_store_create would build it from the reference's own InMemoryAlignmentStore.
 *   out_discard        10 u64 or NULL: DiscardTable in declaration order (:812-826): discard_5p, discard_3p, discard_score,
 *                      discard_aln_frac, discard_aln_len, discard_ori, discard_supp, no_mapping, no_valid_aln, valid_best_aln
 *   out_src_or_null    per retained alignment the index of its record (capacity n_records u32, host or device)
 *   out_group_or_null  per retained read the index of its group (capacity n_groups u32)
 * All inputs may be host or device pointers.
 */
enum { OAR_REC_UNMAPPED = 1, OAR_REC_REVERSE = 2, OAR_REC_SUPPLEMENTARY = 4 };
enum { OAR_STRAND_UNKNOWN = 0, OAR_STRAND_FORWARD = 1, OAR_STRAND_REVERSE = 2 };
typedef struct {
    int32_t which_strand;          /* OAR_STRAND_*: keep both / forward-only / reverse-only alignments (:1003-1022) */
    uint32_t min_aligned_len;      /* :1033 */
    int64_t three_prime_clip;      /* :1040 */
    uint32_t five_prime_clip;      /* :1047 */
    float min_aligned_fraction;    /* :1077 */
    float score_threshold;         /* :1100 */
    float score_prob_denom;        /* :1102, --score-prob-denom, default 5.0 */
} oar_filter_opts;
int oar_store_create_filtered(const uint64_t *group_ptr, const uint32_t *ref_id, const uint32_t *aln_start,
                              const uint32_t *aln_end, const uint32_t *aln_span, const int32_t *score,
                              const uint8_t *flags, const uint32_t *seq_len, uint64_t n_groups, uint64_t n_records,
                              const uint32_t *txp_len, uint32_t n_txps, const oar_filter_opts *opts, int device,
                              oar_store **out, uint64_t out_discard[10], uint32_t *out_src_or_null,
                              uint32_t *out_group_or_null);
/* The CSR a store holds (row_ptr as u64 like `boundaries`, txp_id, prob); each output may be NULL. */
int oar_store_export(oar_store *store, uint64_t *out_row_ptr, uint32_t *out_txp_id, float *out_prob);

/*
 * ---- all GPUs of the box behind one call from one host thread ---------------------------------
 * em::bootstrap(em_info, num_boot, nthreads) (src/em.rs:292-314) builds its own pool and fans the
 * replicates out internally; so does the single-cell driver with cells (src/single_cell.rs:91-193).
 * The entry points below give a caller the same shape: the library runs one host thread per
 * device.  All pointers are HOST pointers here.
 */
typedef struct oar_multi oar_multi; /* opaque: one resident copy of a store on each of `devices` */

/*
 * Upload the store once to devices[0] (as oar_store_create), copy the validated CSR device-to-device
 * (NVLink peer copies) to the other devices; every device builds its own layout.  No collective on
 * the data path afterwards.
 */
int oar_multi_create(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob,
                     const double *aux_or_null, uint64_t n_reads, uint64_t nnz, uint32_t n_txps,
                     const int *devices, int n_devices, oar_multi **out);
void oar_multi_destroy(oar_multi *m);
/* The store on devices[i] (for oar_em, oar_posteriors, ... on one device); owned by the handle. */
oar_store *oar_multi_store(oar_multi *m, int i);
/* n_devices; out_ms: [0] upload + layout on devices[0], [1] peer copies + layouts on the others (wall);
 * out_last_per_device (n_devices u32): units each device ran in the last oar_multi_bootstrap.  Any may be NULL. */
int oar_multi_info(const oar_multi *m, int *n_devices, double out_ms[2], uint32_t *out_last_per_device);

/*
 * em::bootstrap (src/em.rs:292-314) over all devices of the handle: replicate g in [0, num_boot) is
 * the same (seed, g) replicate oar_bootstrap computes, so out (num_boot x M, row g = replicate g) does
 * not depend on the number of devices.  Devices pull replicate ids from a shared counter.
 */
int oar_multi_bootstrap(oar_multi *m, uint32_t num_boot, uint64_t seed, uint32_t max_iter,
                        double conv_thresh, double *out, uint32_t *out_niter);

/*
 * oar_em_batched over several devices (single_cell.rs:91-193): cells are split into contiguous ranges
 * balanced on alignments, every device receives only its cells' rows, results come back as one CSR over
 * all cells (same layout and meaning as oar_em_batched).  out_cells_per_device: n_devices u32 or NULL.
 */
int oar_em_batched_multi(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob,
                         const double *aux_or_null, uint64_t n_reads, uint64_t nnz, uint32_t n_txps,
                         const uint64_t *cell_row_ptr, uint32_t n_cells, const int *devices, int n_devices,
                         uint32_t max_iter, double conv_thresh, uint32_t min_iter,
                         uint64_t *out_cell_ptr, uint32_t *out_txp, double *out_val, uint64_t capacity,
                         uint64_t *out_nnz, uint32_t *out_niter, uint32_t *out_cells_per_device);

/*
 * The bulk coverage model (--model-coverage; bulk.rs:103-108) computed on the device from the resident
 * store: add_interval histograms (src/util/oarfish_types.rs:496-537), logistic_prob
 * (src/util/logistic_probability.rs:40-79) and normalize_read_probs
 * (src/util/normalize_probability.rs:5-74).  The result (== InMemoryAlignmentStore.coverage_probabilities)
 * becomes the store's aux factor, so a following oar_em / oar_bootstrap uses it (em.rs:108).
 *   aln_start, aln_end  nnz u32 == AlnInfo.start / .end, host or device
 *   txp_len             M u32 == TranscriptInfo.len
 *   bin_width           --bin-width (default 100); growth_rate: --growth-rate (default 2.0)
 *   out_aux_or_null     nnz f64 (host or device): the coverage probabilities, for the caller's records
 */
int oar_store_coverage_model(oar_store *store, const uint32_t *aln_start, const uint32_t *aln_end,
                             const uint32_t *txp_len, uint32_t bin_width, double growth_rate,
                             double *out_aux_or_null);

/*
 * The same stage with the binomial bin model the single-cell driver uses per cell (single_cell.rs:132-137):
 * binomial_continuous_prob / binomial_probability (src/util/binomial_probability.rs:180-224, :7-178; ln_gamma as in
 * statrs 0.18) in place of logistic_prob; histograms and normalize_read_probs as above.
 */
int oar_store_coverage_model_binomial(oar_store *store, const uint32_t *aln_start, const uint32_t *aln_end,
                                      const uint32_t *txp_len, uint32_t bin_width, double *out_aux_or_null);

/*
 * Read-level assignment probabilities with the final counts: the inner loop of
 * write_out_prob (src/util/write_function.rs:283-332).  out_prob[j] (nnz f64, host
 * or device) = the alignment's probability, clamped to [0,1], dropped (0) below
 * display_thresh and renormalised over the kept ones; out_kept_or_null[r] (N u32) =
 * alignments kept for read r.  Uses prob and the store's aux factor as the
 * coverage probability (the reference leaves the KDE factor out of this step).
 */
int oar_posteriors(oar_store *store, const double *counts, double display_thresh, double *out_prob,
                   uint32_t *out_kept_or_null);

/* aux_counts::get_aux_counts (src/util/aux_counts.rs:23-50): per transcript (M u32 each,
 * host or device) the number of alignments and the number from single-alignment reads. */
int oar_aux_counts(oar_store *store, uint32_t *out_unique, uint32_t *out_total);

/* Layout of the store in HBM: [0] tiled layout built, [1] tiles, [2] alignment
 * slots in tiles, [3] rows swept from the CSR instead (too long / did not fit),
 * [4] sum of per-tile distinct transcripts, [5] sum of per-tile M-step items,
 * [6] tile span, [7] active kernel (oar_kernel). */
int oar_store_layout_info(const oar_store *store, uint64_t out[8]);

/* Inspection (tools/layout_model.py): the per-slot words of tiles [first_tile, first_tile + n_tiles) of the
 * tiled layout, 1024 u32 per tile to host memory: byte offset of the slot's transcript in the tile's prev[] table
 * (low 16 bits) | byte offset of its x position (high 16 bits); optionally (n_tiles u32) the byte offset of each
 * tile's trash slots, i.e. the end of its items: positions at or above it belong to padding and to alignments whose
 * transcript is not aggregated in the tile.  No counterpart in the reference. */
int oar_store_layout_lpos(oar_store *store, uint32_t first_tile, uint32_t n_tiles, uint32_t *out, uint32_t *out_trash_or_null);

/* Timings of the last compute call on this store, milliseconds (CUDA events):
 * [0] upload+layout, [1] EM loop (device), [2] result download, [3] weight generation. */
int oar_store_timings(const oar_store *store, double out_ms[4]);

/* Counters of the last compute call: [0] kernels launched, [1] sweeps executed. */
int oar_store_counters(const oar_store *store, uint64_t out[2]);

/*
 * Raw single E+M sweep for measurement and tests: curr (M f64, device) is
 * zeroed then accumulated from prev (M f64, device).  weights_or_null: N u32
 * device.  Runs on the store's stream; *not* synchronised unless sync != 0.
 */
int oar_sweep(oar_store *store, const double *prev_dev, double *curr_dev,
              const uint32_t *weights_or_null, int sync);

/* `reps` back-to-back sweeps bracketed by CUDA events recorded on the store's
 * stream; *out_ms_total is the elapsed device time of all of them (curr is
 * zeroed once and accumulates).  For roofline measurement. */
int oar_sweep_timed(oar_store *store, const double *prev_dev, double *curr_dev,
                    const uint32_t *weights_or_null, int reps, float *out_ms_total);

/* The CUDA stream (cudaStream_t) the store launches on, for event timing. */
void *oar_store_stream(oar_store *store);

#ifdef __cplusplus
}
#endif
#endif /* OARFISH_EM_H */
