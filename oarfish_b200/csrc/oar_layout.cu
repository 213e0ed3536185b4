// oar_layout.cu -- builds the locality-tiled layout (oar_tiled.cuh) on the device.
//
// One-time work per store: a key-value radix sort of the rows by their smallest
// transcript id (CUB DeviceRadixSort; library code, not on the EM hot path), a
// scan of the sorted row lengths, and one CTA per tile that packs rows into
// warp-chunks, sorts the tile's alignments by transcript and emits the
// per-alignment (table index, position) stream.
#include <algorithm>
#include <vector>

#include "oar_store.cuh"
#include "oar_tiled.cuh"
#include "oar_lane.cuh"

namespace oar {

void free_tiled_layout(oar_store *s)
{
    TiledLayout &t = s->tl;
    cudaStream_t st = s->stream;
    dfree(t.prob, st); dfree(t.lpos, st); dfree(t.aux, st); dfree(t.rec, st); dfree(t.records, st); dfree(t.trow, st);
    dfree(t.fallback, st); dfree(t.wperm, st); dfree(t.blobs, st); dfree(t.groups, st);
    const int kind = t.kind;
    t = TiledLayout();
    t.kind = kind;
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
    s->tl.kind = 0;
    TiledLayout &t = s->tl;
    const uint32_t N = (uint32_t)s->n_reads;
    if (span == 0 || span > (uint32_t)kTile) span = (uint32_t)kTile - 5u * kWarps;
    t.span = span;
    if (N == 0 || s->nnz == 0) { t.ready = true; return OAR_OK; }
    if (s->n_txps >= kMaxTxps) return fail(OAR_ERR_UNSUPPORTED, "tiled layout needs n_txps < 2^28");
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
    row_keys<<<gridN, threads, 0, st>>>(s->d_row_ptr, s->d_txp, N, key, idx, counters);
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
        const size_t worst = (size_t)n_tiles * (kRecTable + 16) + 4 * (size_t)total + 4 * ((size_t)total / kAggMin) + 64;
        OAR_CUDA(sc.alloc((char **)&records_tmp, worst));
        BuildArgs a;
        a.row_ptr = s->d_row_ptr; a.txp = s->d_txp; a.prob = s->d_prob; a.aux = s->d_aux;
        a.srow = srow; a.tile_row = tile_row;
        a.o_prob = t.prob; a.o_lpos = t.lpos; a.o_aux = t.aux; a.o_rec = t.rec; a.o_records = records_tmp;
        a.o_trow = t.trow; a.fallback = t.fallback; a.cursors = counters + 4;
        if (s->prob_ready) OAR_CUDA(cudaStreamWaitEvent(st, s->prob_ready, 0));   // prob / aux uploaded on a second stream
        build_tiles<<<n_tiles, kThreads, 0, st>>>(a);
        OAR_CUDA(cudaGetLastError());
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

// Row-per-lane layout (oar_lane.cuh): same row order and tile windows as the chunk layout, but a
// tile's rows are re-sorted by length and stored one read per lane; no padding slots in HBM.
static int build_lane_layout(oar_store *s, uint32_t span)
{
    using namespace lane;
    free_tiled_layout(s);
    TiledLayout &t = s->tl;
    t.kind = 1;
    const uint32_t N = (uint32_t)s->n_reads;
    if (span < 256u || span > (uint32_t)kSpanMax) span = (uint32_t)kSpanDefault;
    t.span = span;
    if (N == 0 || s->nnz == 0) { t.ready = true; return OAR_OK; }
    cudaStream_t st = s->stream;
    Scratch sc; sc.st = st;
    uint32_t *key = nullptr, *idx = nullptr, *key_s = nullptr, *srow = nullptr, *slen = nullptr, *soff = nullptr;
    uint32_t *counters = nullptr;
    OAR_CUDA(sc.alloc(&key, N)); OAR_CUDA(sc.alloc(&idx, N));
    OAR_CUDA(sc.alloc(&key_s, N)); OAR_CUDA(sc.alloc(&srow, N));
    OAR_CUDA(sc.alloc(&counters, 16));
    OAR_CUDA(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 16, st));
    const int threads = 256;
    const int gridN = (int)std::min<uint64_t>((N + threads - 1) / threads, (uint64_t)s->sm_count * 32);
    tiled::row_keys<<<gridN, threads, 0, st>>>(s->d_row_ptr, s->d_txp, N, key, idx, counters);
    OAR_CUDA(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        OAR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key_s, idx, srow, (int)N, 0, 32, st));
        void *tmp = nullptr;
        OAR_CUDA(sc.alloc((char **)&tmp, tmp_bytes));
        OAR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key_s, idx, srow, (int)N, 0, 32, st));
    }
    uint32_t h_counters[16];
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
        tiled::sorted_lens<<<gridT, threads, 0, st>>>(s->d_row_ptr, srow, n_tiled, slen);
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
        tiled::tile_row_starts<<<(n_tiles + 1 + threads - 1) / threads, threads, 0, st>>>(soff, n_tiled, span, n_tiles, tile_row);
        OAR_CUDA(cudaGetLastError());
    }
    t.n_tiles = n_tiles;

    OAR_CUDA(dmalloc(&t.fallback, sizeof(uint32_t) * std::max<uint32_t>(n_long, 1), st));
    OAR_CUDA(dmalloc(&t.trow, sizeof(uint32_t) * std::max<uint32_t>(n_tiled, 1), st));
    OAR_CUDA(dmalloc(&t.wperm, sizeof(uint32_t) * ((size_t)n_tiled + 64), st));
    OAR_CUDA(cudaMemsetAsync(t.wperm, 0, sizeof(uint32_t) * ((size_t)n_tiled + 64), st));
    uint4 *blobs_tmp = nullptr;
    if (n_tiles > 0) {
        // a group is closed by its 32nd row, by the alignment cap, or by the end of its tile
        const size_t max_groups = (size_t)n_tiled / 32 + (size_t)(total / (uint64_t)(kGroupCap - kRowCap)) + n_tiles + 1;
        const size_t max_pairs = (size_t)total + max_groups + 4;   // one pad pair per odd group
        if (s->d_aux) OAR_CUDA(dmalloc(&t.aux, sizeof(double) * max_pairs, st));
        OAR_CUDA(dmalloc(&t.groups, sizeof(uint2) * max_groups, st));
        // blob bytes: 8 per pair; per group: header + row lengths + section roundings (80 B); per (group,
        // transcript): a table entry and at most one item more than its alignments / 16 (<= 9 per alignment)
        const size_t worst = 8 * max_pairs + 80 * max_groups + 9 * (size_t)total + 64;
        OAR_CUDA(sc.alloc((char **)&blobs_tmp, worst));
        static bool attr_set[64] = {false};
        if (!attr_set[s->device & 63]) {
            OAR_CUDA(cudaFuncSetAttribute(build_lane_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BuildSmem)));
            attr_set[s->device & 63] = true;
        }
        BuildArgs a;
        a.row_ptr = s->d_row_ptr; a.txp = s->d_txp; a.prob = s->d_prob; a.aux = s->d_aux;
        a.srow = srow; a.soff = soff; a.tile_row = tile_row;
        a.o_aux = t.aux; a.o_groups = t.groups; a.o_blobs = blobs_tmp; a.o_trow = t.trow;
        a.cursors = counters + 4;
        if (s->prob_ready) OAR_CUDA(cudaStreamWaitEvent(st, s->prob_ready, 0));
        build_lane_tiles<<<n_tiles, kBuildThreads, sizeof(BuildSmem), st>>>(a);
        OAR_CUDA(cudaGetLastError());
    }
    if (n_long > 0) {
        const int g = (int)std::min<uint64_t>((N - n_tiled + threads - 1) / threads, (uint64_t)s->sm_count * 8);
        tiled::collect_long_rows<<<g, threads, 0, st>>>(s->d_row_ptr, srow, n_tiled, N, t.fallback, counters + 2);
        OAR_CUDA(cudaGetLastError());
    }
    OAR_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    t.n_fallback = h_counters[2];
    t.record_bytes = (uint64_t)h_counters[4] * 16u;
    t.n_groups = h_counters[5]; t.n_pairs = h_counters[6];
    t.sum_d = h_counters[7]; t.sum_u = h_counters[8];
    t.max_rec = h_counters[9]; t.max_d = h_counters[10]; t.max_xs = h_counters[11]; t.max_nnz = h_counters[12];
    if (n_tiles > 0) {
        // compact copy of the blobs (the worst-case scratch goes back to the pool)
        OAR_CUDA(dmalloc(&t.blobs, std::max<uint64_t>(t.record_bytes, 16) + 64, st));
        OAR_CUDA(cudaMemcpyAsync(t.blobs, blobs_tmp, t.record_bytes, cudaMemcpyDeviceToDevice, st));
        OAR_CUDA(cudaStreamSynchronize(st));
        const Geometry g = make_geometry(t.max_rec, t.max_d, t.max_xs);
        if ((size_t)g.warp_bytes * kWarps > 227u * 1024u) {
            free_tiled_layout(s);
            return fail(OAR_ERR_UNSUPPORTED, "lane layout: a group does not fit shared memory");
        }
    }
    t.ready = true;
    return OAR_OK;
}

// s->tl.kind picks the layout: 0 = warp-chunk tiles (default), 1 = row-per-lane groups (OAR_LAYOUT=lane at
// store creation).  Rebuilds (coverage model) and sub-stores keep the store's kind.
int build_tiled_layout(oar_store *s, uint32_t span)
{
    return s->tl.kind == 0 ? build_chunk_layout(s, span) : build_lane_layout(s, span);
}

}  // namespace oar
