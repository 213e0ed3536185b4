// oar_layout.cu -- builds the locality-tiled layout (oar_tiled.cuh) on the device.
//
// One-time work per store: a key-value radix sort of the rows by their smallest
// transcript id (CUB DeviceRadixSort; library code, not on the EM hot path), a
// scan of the sorted row lengths, and one CTA per tile that packs rows into
// warp-chunks, sorts the tile's alignments by transcript and emits the
// per-alignment (table index, position) stream.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "oar_store.cuh"
#include "oar_tiled.cuh"

namespace oar {

void free_tiled_layout(oar_store *s)
{
    TiledLayout &t = s->tl;
    cudaStream_t st = s->stream;
    dfree(t.prob, st); dfree(t.lpos, st); dfree(t.aux, st); dfree(t.rec, st); dfree(t.records, st); dfree(t.trow, st);
    dfree(t.fallback, st); dfree(t.wperm, st); dfree(t.wlane, st);
    t = TiledLayout();
}

namespace {
struct Scratch {
    cudaStream_t st = nullptr;
    std::vector<void *> ptrs;
    ~Scratch() { for (void *p : ptrs) dfree(p, st); }
    template <typename T> cudaError_t alloc(T **p, size_t n)
    {
        cudaError_t e = dmalloc(p, sizeof(T) * std::max<size_t>(n, 1), st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};
}  // namespace

static int build_chunk_layout(oar_store *s, uint32_t span)
{
    using namespace tiled;
    free_tiled_layout(s);
    TiledLayout &t = s->tl;
    const uint32_t N = (uint32_t)s->n_reads;
    if (span == 0 || span > (uint32_t)kTile) span = (uint32_t)kTile - 5u * kWarps;
    t.span = span;
    if (N == 0 || s->nnz == 0) { t.ready = true; return OAR_OK; }
    if (s->n_txps >= kMaxTxps) return fail(OAR_ERR_UNSUPPORTED, "tiled layout needs n_txps < 2^32 - 1");
    cudaStream_t st = s->stream;
    Scratch sc; sc.st = st;
    uint32_t *key = nullptr, *idx = nullptr, *key_s = nullptr, *srow = nullptr, *slen = nullptr, *soff = nullptr;
    uint32_t *counters = nullptr;
    OAR_CUDA(sc.alloc(&key, N)); OAR_CUDA(sc.alloc(&idx, N));
    OAR_CUDA(sc.alloc(&key_s, N)); OAR_CUDA(sc.alloc(&srow, N));
    OAR_CUDA(sc.alloc(&counters, 12));
    OAR_CUDA(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 12, st));
    const int threads = 256;
    const int gridN = (int)std::min<uint64_t>((N + threads - 1) / threads, (uint64_t)s->sm_count * 32);
    // transcript ids that are not gene-local: key the rows by a co-occurrence numbering instead (see oar_tiled.cuh)
    uint32_t *vid = nullptr;
    {
        const char *cl = getenv("OAR_CLUSTER_IDS");   // 0 = never, 1 = always, default: when > 2 % of the rows have widely spread ids
        const int mode = cl ? atoi(cl) : -1;
        bool cluster = mode == 1;
        if (mode < 0) {
            row_id_spread<<<gridN, threads, 0, st>>>(s->d_row_ptr, s->d_txp, N, counters + 11);
            OAR_CUDA(cudaGetLastError());
            uint32_t h_spread = 0;
            OAR_CUDA(cudaMemcpyAsync(&h_spread, counters + 11, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            OAR_CUDA(cudaStreamSynchronize(st));
            cluster = (uint64_t)h_spread * 50u > (uint64_t)N;
        }
        if (cluster) {
            const uint32_t M = s->n_txps;
            uint32_t *label = nullptr, *tidx = nullptr, *tsorted = nullptr;
            uint64_t *lkey = nullptr, *lkey_s = nullptr;
            OAR_CUDA(sc.alloc(&label, M)); OAR_CUDA(sc.alloc(&vid, M)); OAR_CUDA(sc.alloc(&tidx, M)); OAR_CUDA(sc.alloc(&tsorted, M));
            OAR_CUDA(sc.alloc(&lkey, M)); OAR_CUDA(sc.alloc(&lkey_s, M));
            const int gridM = (int)std::min<uint64_t>((M + threads - 1) / threads, (uint64_t)s->sm_count * 8);
            label_init<<<gridM, threads, 0, st>>>(label, M);
            for (int round = 0; round < 4; ++round) {
                label_rows<<<gridN, threads, 0, st>>>(s->d_row_ptr, s->d_txp, N, label);
                label_jump<<<gridM, threads, 0, st>>>(label, M);
            }
            label_keys<<<gridM, threads, 0, st>>>(label, M, lkey, tidx);
            OAR_CUDA(cudaGetLastError());
            size_t tmp_bytes = 0;
            OAR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, lkey, lkey_s, tidx, tsorted, (int)M, 0, 64, st));
            void *tmp = nullptr;
            OAR_CUDA(sc.alloc((char **)&tmp, tmp_bytes));
            OAR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, lkey, lkey_s, tidx, tsorted, (int)M, 0, 64, st));
            label_number<<<gridM, threads, 0, st>>>(tsorted, M, vid);
            OAR_CUDA(cudaGetLastError());
        }
    }
    row_keys<<<gridN, threads, 0, st>>>(s->d_row_ptr, s->d_txp, N, vid, key, idx, counters);
    OAR_CUDA(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        OAR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key_s, idx, srow, (int)N, 0, 32, st));
        void *tmp = nullptr;
        OAR_CUDA(sc.alloc((char **)&tmp, tmp_bytes));
        OAR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key_s, idx, srow, (int)N, 0, 32, st));
    }
    uint32_t h_counters[12];
    OAR_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_tiled = N - h_counters[0];
    const uint32_t n_long = h_counters[1];
    t.n_tiled_rows = n_tiled;

    uint32_t n_tiles = 0;
    uint32_t *tile_row = nullptr;
    uint64_t total = 0;
    if (n_tiled > 0) {
        OAR_CUDA(sc.alloc(&slen, n_tiled + 1)); OAR_CUDA(sc.alloc(&soff, n_tiled + 1));
        const int gridT = (int)std::min<uint64_t>((n_tiled + threads - 1) / threads, (uint64_t)s->sm_count * 32);
        sorted_lens<<<gridT, threads, 0, st>>>(s->d_row_ptr, srow, n_tiled, slen);
        OAR_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        OAR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, slen, soff, (int)n_tiled + 1, st));
        void *tmp = nullptr;
        OAR_CUDA(sc.alloc((char **)&tmp, tmp_bytes));
        OAR_CUDA(cudaMemsetAsync(slen + n_tiled, 0, sizeof(uint32_t), st));
        OAR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, slen, soff, (int)n_tiled + 1, st));
        uint32_t h_total = 0;
        OAR_CUDA(cudaMemcpyAsync(&h_total, soff + n_tiled, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        OAR_CUDA(cudaStreamSynchronize(st));
        total = h_total;
        n_tiles = (uint32_t)((total + span - 1) / span);
        OAR_CUDA(sc.alloc(&tile_row, n_tiles + 1));
        tile_row_starts<<<(n_tiles + 1 + threads - 1) / threads, threads, 0, st>>>(soff, n_tiled, span, n_tiles, tile_row);
        OAR_CUDA(cudaGetLastError());
    }
    t.n_tiles = n_tiles;

    // outputs (the variable-length records first at their worst-case size, compacted below)
    const size_t slots = (size_t)n_tiles * kTile;
    uint4 *records_tmp = nullptr;
    OAR_CUDA(dmalloc(&t.fallback, sizeof(uint32_t) * std::max<uint32_t>(N, 1), st));
    OAR_CUDA(dmalloc(&t.trow, sizeof(uint32_t) * std::max<uint32_t>(n_tiled, 1), st));
    OAR_CUDA(dmalloc(&t.wperm, sizeof(uint32_t) * ((size_t)n_tiled + kChunk + 257), st));
    OAR_CUDA(cudaMemsetAsync(t.wperm, 0, sizeof(uint32_t) * ((size_t)n_tiled + kChunk + 257), st));
    if (n_tiles > 0) {
        OAR_CUDA(dmalloc(&t.prob, sizeof(float) * slots, st));
        OAR_CUDA(dmalloc(&t.lpos, sizeof(uint32_t) * slots, st));
        if (s->d_aux) OAR_CUDA(dmalloc(&t.aux, sizeof(double) * slots, st));
        OAR_CUDA(dmalloc(&t.rec, sizeof(uint2) * n_tiles, st));
        // a record holds at most (slots of the tile) table entries: bound the total by the real slot count
        const size_t worst = (size_t)n_tiles * (kRecTable + 16 + kRecAlign) + 4 * (size_t)total + 4 * ((size_t)total / kAggMin) + 64;
        OAR_CUDA(sc.alloc((char **)&records_tmp, worst));
        BuildArgs a;
        a.row_ptr = s->d_row_ptr; a.txp = s->d_txp; a.prob = s->d_prob; a.aux = s->d_aux;
        a.srow = srow; a.tile_row = tile_row;
        a.o_prob = t.prob; a.o_lpos = t.lpos; a.o_aux = t.aux; a.o_rec = t.rec; a.o_records = records_tmp;
        a.o_trow = t.trow; a.fallback = t.fallback; a.cursors = counters + 4;
        const bool trace = getenv("OAR_TRACE") != nullptr;   // development: time of the tile builder on stderr
        if (trace) OAR_CUDA(cudaEventRecord(s->ev[2], st));
        build_tiles<<<n_tiles, kThreads, 0, st>>>(a);
        OAR_CUDA(cudaGetLastError());
        if (trace) {
            OAR_CUDA(cudaEventRecord(s->ev[3], st));
            OAR_CUDA(cudaEventSynchronize(s->ev[3]));
            float ms_pre = 0.f, ms_tiles = 0.f;
            cudaEventElapsedTime(&ms_pre, s->ev[0], s->ev[2]);
            cudaEventElapsedTime(&ms_tiles, s->ev[2], s->ev[3]);
            fprintf(stderr, "[oar] layout: upload + keys + sort %.2f ms, build_tiles %.2f ms (%u tiles)\n", ms_pre, ms_tiles, n_tiles);
        }
    }
    if (n_long > 0) {
        const int g = (int)std::min<uint64_t>((N - n_tiled + threads - 1) / threads, (uint64_t)s->sm_count * 8);
        collect_long_rows<<<g, threads, 0, st>>>(s->d_row_ptr, srow, n_tiled, N, t.fallback, counters + 4);
        OAR_CUDA(cudaGetLastError());
    }
    OAR_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    t.n_fallback = h_counters[4];
    t.sum_d = h_counters[5];
    t.sum_u = h_counters[6];
    t.record_bytes = (uint64_t)h_counters[7] * 16u;
    t.max_rec = h_counters[8]; t.max_d = h_counters[9]; t.max_u = h_counters[10];
    if (n_tiles > 0) {
        OAR_CUDA(dmalloc(&t.records, std::max<uint64_t>(t.record_bytes, 16), st));
        OAR_CUDA(cudaMemcpyAsync(t.records, records_tmp, t.record_bytes, cudaMemcpyDeviceToDevice, st));
        OAR_CUDA(cudaStreamSynchronize(st));
    }
    t.ready = true;
    return OAR_OK;
}

namespace {
// per tile: the smallest and the largest group (cell) among its rows; groups are contiguous row ranges given by
// group_rows[0 .. n_groups] (original row numbers).  One warp per tile.
__global__ void tile_group_ranges(const uint2 *__restrict__ rec, const uint4 *__restrict__ records, const uint32_t *__restrict__ trow,
                                  uint32_t n_tiles, uint32_t n_tiled_rows, const uint64_t *__restrict__ group_rows, uint32_t n_groups,
                                  uint2 *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += warps) {
        auto first_row = [&](uint32_t t) {   // tile-order index of the tile's first row: chunk_row[0] of its record
            return reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(records + rec[t].x) + tiled::kRecRow)[0];
        };
        const uint32_t b = first_row(tile), e = tile + 1 < n_tiles ? first_row(tile + 1) : n_tiled_rows;
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
        for (uint32_t k = b + lane; k < e; k += 32u) {
            const uint64_t row = trow[k];
            uint32_t l = 0, h = n_groups;   // last group whose first row is <= row
            while (h - l > 1) { const uint32_t m = l + (h - l) / 2; if (group_rows[m] <= row) l = m; else h = m; }
            lo = min(lo, l); hi = max(hi, l);
        }
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) out[tile] = make_uint2(lo, hi);
    }
}
}  // namespace

cudaError_t tile_group_ranges_enqueue(oar_store *s, const uint64_t *d_group_rows, uint32_t n_groups, uint2 *d_out)
{
    const TiledLayout &t = s->tl;
    if (!t.ready || t.n_tiles == 0 || n_groups == 0) return cudaSuccess;
    const int blocks = (int)std::min<uint32_t>((t.n_tiles + 7) / 8, (uint32_t)s->sm_count * 8);
    tile_group_ranges<<<blocks, 256, 0, s->stream>>>(t.rec, t.records, t.trow, t.n_tiles, t.n_tiled_rows, d_group_rows, n_groups, d_out);
    return cudaGetLastError();
}

// Rebuilds (coverage model) and sub-stores go through the same entry point.
int build_tiled_layout(oar_store *s, uint32_t span)
{
    return build_chunk_layout(s, span);
}

}  // namespace oar
