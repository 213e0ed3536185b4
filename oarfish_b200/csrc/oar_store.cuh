// oar_store.cuh -- the store handle behind the opaque oar_store of the C ABI.
#pragma once
#include "oar_common.cuh"
#include "oar_ctx.cuh"
#ifndef OAR_TILE_WARPS
#define OAR_TILE_WARPS 8
#endif
#define OAR_TILE_WARPS_DEFAULT OAR_TILE_WARPS

namespace oar {

// Locality-tiled copy of the store (oar_tiled.cuh): tiles of warp-chunks.
struct TiledLayout {
    bool ready = false;
    uint32_t n_tiles = 0, n_tiled_rows = 0, n_fallback = 0, span = 0;
    uint64_t sum_d = 0, sum_u = 0;
    float *prob = nullptr;
    uint32_t *lpos = nullptr;
    double *aux = nullptr;
    uint2 *rec = nullptr;          // per tile: {record offset (16 B granules), record bytes}
    uint4 *records = nullptr;      // per-tile records (lane descriptors, chunk info, table, units)
    uint64_t record_bytes = 0;
    uint32_t max_rec = 0, max_d = 0, max_u = 0;   // per-tile maxima (size the sweep's shared memory)
    uint32_t *trow = nullptr;      // tile-order row -> original row (n_tiled_rows)
    uint32_t *fallback = nullptr;  // original row ids swept from the CSR
    uint32_t *wperm = nullptr;     // bootstrap weights in tile order
    uint16_t *wlane = nullptr;     // bootstrap weight per lane of every tile (tiled::lane_weights); allocated with the first replicate
};

struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int kernel = 0;
    bool fused = false;   // built with the convergence bookkeeping inside the sweep
};

}  // namespace oar

struct oar_store {
    int device = 0;
    int sm_count = 148;
    oar::DeviceCtx *ctx = nullptr;  // per-device pools the stream, events and pinned state below come from
    cudaStream_t stream = nullptr;
    uint64_t n_reads = 0, nnz = 0;
    uint32_t n_txps = 0;
    int kernel = OAR_KERNEL_ROWGROUP;
    bool allow_fused = true;   // convergence bookkeeping inside the sweep's head (OAR_FUSED_UPDATE=0 turns it off)
    bool borrowed = false;  // sub-store of another store: row_ptr / prob / aux / stream / events are not owned
    int ctas_per_sm = 0;  // persistent CTAs of the tiled sweep per SM: 0 = the instantiation's own register budget (tiled::sweep_ctas), else a cap (OAR_CTAS_PER_SM); shared memory may allow fewer

    // CSR in HBM (original read order)
    uint32_t *d_row_ptr = nullptr;  // N+1
    uint32_t *d_txp = nullptr;      // nnz
    float *d_prob = nullptr;        // nnz
    double *d_aux = nullptr;        // nnz or null

    oar::TiledLayout tl;

    // EM work buffers
    double *d_counts[3] = {nullptr, nullptr, nullptr};   // three rotate: prev, curr, and the one being zeroed for the sweep after
    OarEmState *d_state = nullptr;
    OarEmState *h_state = nullptr;  // pinned, oar::kHostStateSlots slots (from the context's pool)
    uint32_t *d_weights = nullptr;  // N, bootstrap weights of the current replicate (read order)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t slot_ev[2] = {nullptr, nullptr};

    oar::GraphSlot graphs[2];  // [0] unweighted, [1] weighted

    oar_progress_fn progress = nullptr;   // called after every polled batch of iterations (em.rs:219-233's log lines)
    void *progress_user = nullptr;

    double timings[4] = {0, 0, 0, 0};
    uint64_t counters[2] = {0, 0};
};

namespace oar {
// Build s->tl from the CSR arrays already resident on s->device (enqueued on
// s->stream, synchronises).  Returns an oar_status.
int build_tiled_layout(oar_store *s, uint32_t span);
// A store over the parent's reads with different transcript ids (takes ownership of d_txp): used by the
// batched per-cell EM, where ids are (cell, transcript) pairs.  Shares the parent's stream and CSR arrays.
int substore_create(oar_store *parent, uint32_t *d_txp, uint32_t n_txps, oar_store **out);
// One fused E+M sweep prev -> curr on the store's stream (curr must be zero); see oar_em.cu.
cudaError_t sweep_enqueue(oar_store *s, const double *prev, double *curr, const OarEmState *state, int check_done);
// The same over the tiles in tile_list[0 .. *n_active) only (both in device memory; the batched per-cell EM drops the
// tiles of converged cells).
cudaError_t sweep_enqueue_list(oar_store *s, const double *prev, double *curr, const OarEmState *state, int check_done,
                               const uint32_t *tile_list, const uint32_t *n_active);
// d_out[tile] = {smallest, largest} group among the tile's rows, groups being the contiguous row ranges
// d_group_rows[0 .. n_groups] (cells of the batched per-cell EM)
cudaError_t tile_group_ranges_enqueue(oar_store *s, const uint64_t *d_group_rows, uint32_t n_groups, uint2 *d_out);
void free_tiled_layout(oar_store *s);
// Handle + stream + events + pinned state + EM work buffers for a store whose CSR arrays the caller fills in on
// s->stream (d_row_ptr u32 N+1, d_txp, d_prob[, d_aux]; n_reads / nnz may be set afterwards); finish_store() then
// builds the tiled layout and reads the tuning environment.
int new_store(int device, uint64_t n_reads, uint64_t nnz, uint32_t n_txps, const char *who, oar_store **out);
int finish_store(oar_store *s);
// oar_store_create for a slice of a larger store (row_ptr[0] == row_base; txp_id / prob / aux start at the slice's first alignment)
int store_create_slice(const uint64_t *row_ptr, uint64_t row_base, const uint32_t *txp_id, const float *prob,
                       const double *aux_or_null, uint64_t n_reads, uint64_t nnz, uint32_t n_txps, int device, oar_store **out);
// a copy of `src` on another device (peer copy of the validated CSR + that device's own layout)
int store_clone(const oar_store *src, int device, oar_store **out);
}  // namespace oar
