// oar_tiled.cuh -- the locality-tiled store layout and its fused E+M sweep.
//
// Why: the plain CSR sweep (em_sweep_rowgroup) issues one random 8-byte gather
// and one f64 RED per alignment.  On B200 both serialise at ~1 lane-address per
// cycle per SM (L1 wavefronts / REDG issue), 3-10x over the HBM budget of
// ~0.37 cycles per alignment per SM.  The tiled layout removes every
// per-alignment global access:
//
//   * reads (rows) are ordered by their smallest transcript id, so the rows of
//     one tile (<= 1024 alignment slots) touch a handful of transcripts;
//   * a tile carries its own table of distinct transcripts; prev[] for the
//     table is gathered once into shared memory, alignments store a 16-bit
//     table index instead of the 32-bit id;
//   * the M-step scatter goes through shared memory: every alignment also
//     stores `pos`, its position in the tile's transcript-sorted order, x_j =
//     w_j/denom is written to xs[pos] and 8-slot units of xs are summed
//     contiguously, then combined across a warp and flushed with ONE f64 RED
//     per (warp, transcript) instead of one per alignment.  No shared-memory
//     atomics (f64 smem atomics are CAS loops on sm_100a).
//   * rows never straddle a 128-slot warp-chunk, so the per-row denominator
//     (em.rs:98-112) is a segmented warp scan in registers.
//
// Per alignment the HBM stream is 4 B (prob f32) + 4 B (table index u16 | pos
// u16): the same 8 B as CSR's txp_id + prob; row boundaries are a 1-bit head
// mask instead of a 4-byte row_ptr entry.
//
// Rows longer than a warp-chunk, or that do not fit their tile, are listed in
// `fallback_rows` and swept by em_sweep_rowgroup from the original CSR.
#pragma once
#include <cub/cub.cuh>

#include "oar_common.cuh"

namespace oar {
namespace tiled {

constexpr int kWarps = 8;                 // warp-chunks per tile
constexpr int kChunk = 128;               // alignment slots per warp-chunk (4 per lane)
constexpr int kChunkCap = kChunk - 1;     // rows use at most 127 slots: a padding pseudo-row always closes the chunk
constexpr int kTile = kWarps * kChunk;    // 1024 slots
constexpr int kThreads = kWarps * 32;     // 256
constexpr int kAggMin = 4;                // transcripts with >= kAggMin alignments in a tile are aggregated in smem
constexpr int kMaxUnits = kTile / 4;      // sum ceil(cnt/8) over cnt >= 4  <=  kTile/4
constexpr int kTrash = kMaxUnits * 9;     // xs slot for padding alignments (never summed)
constexpr uint32_t kInfoStray = 8u;       // chunk_info bit 3: chunk holds alignments that RED straight to global
constexpr uint32_t kInfoMulti = 16u;      // chunk_info bit 4: some lane holds >= 2 row heads (general path)
constexpr uint32_t kNoTxp = 0xFFFFFFFFu;
static_assert(kMaxUnits == kThreads, "one unit per thread in phase 2");

struct View {
    uint32_t n_tiles;
    const float *prob;         // n_tiles * kTile
    const uint32_t *lpos;      // n_tiles * kTile : (table index * 8) | (pos * 8) << 16  (smem byte offsets)
    const double *aux;         // n_tiles * kTile or null
    const uint4 *heads;        // n_tiles * kWarps : 128-bit row-head mask per warp-chunk
    const uint32_t *chunk_row; // n_tiles * kWarps : tile-order index of the chunk's first row
    const uint32_t *chunk_info;// n_tiles * kWarps : scan steps (bits 0-2) | kInfoStray | kInfoMulti
    const uint4 *meta;         // n_tiles : {table_off, unit_off, D | U << 16, first tile-order row}
    const uint32_t *table;     // sum D : distinct transcript ids per tile
    const uint32_t *unit_txp;  // sum U : transcript id of each 8-slot unit
    const uint8_t *unit_cnt;   // sum U : valid slots (1..8) of each unit
};

// ---------------------------------------------------------------------------
// layout construction
// ---------------------------------------------------------------------------

// key = smallest transcript id of the row (locality key); rows that cannot be
// tiled (empty, or longer than a warp-chunk) get kNoTxp and sort to the end.
static __global__ void row_keys(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp, uint64_t n_rows,
                         uint32_t *__restrict__ key, uint32_t *__restrict__ idx, uint32_t *__restrict__ counters)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t n_long = 0, n_skip = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        uint32_t k = kNoTxp;
        if (e > s && e - s <= (uint32_t)kChunkCap) {
            for (uint32_t j = s; j < e; ++j) k = min(k, txp[j]);
        } else {
            ++n_skip;
            if (e > s) ++n_long;
        }
        key[r] = k;
        idx[r] = (uint32_t)r;
    }
    if (n_skip) atomicAdd(counters + 0, n_skip);   // rows not tiled
    if (n_long) atomicAdd(counters + 1, n_long);   // of which: too long (go to fallback)
}

// lengths of the tiled rows in sorted order
static __global__ void sorted_lens(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ srow, uint32_t n,
                            uint32_t *__restrict__ slen)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t r = srow[k];
        slen[k] = row_ptr[r + 1] - row_ptr[r];
    }
}

// rows that were not tiled because they are longer than a warp-chunk
static __global__ void collect_long_rows(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ srow,
                                  uint32_t first, uint32_t n_rows, uint32_t *__restrict__ fallback,
                                  uint32_t *__restrict__ cursor)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = first + blockIdx.x * blockDim.x + threadIdx.x; k < n_rows; k += stride) {
        const uint32_t r = srow[k];
        if (row_ptr[r + 1] > row_ptr[r]) fallback[atomicAdd(cursor, 1u)] = r;
    }
}

// tile t owns the sorted rows whose first alignment offset lies in [t*span, (t+1)*span)
static __global__ void tile_row_starts(const uint32_t *__restrict__ soff, uint32_t n_rows, uint32_t span, uint32_t n_tiles,
                                uint32_t *__restrict__ tile_row)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { tile_row[t] = n_rows; return; }
    const uint64_t target = (uint64_t)t * span;
    uint32_t lo = 0, hi = n_rows;
    while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if ((uint64_t)soff[mid] < target) lo = mid + 1; else hi = mid; }
    tile_row[t] = lo;
}

struct BuildArgs {
    const uint32_t *row_ptr; const uint32_t *txp; const float *prob; const double *aux;
    const uint32_t *srow;      // sorted position -> original row
    const uint32_t *tile_row;  // n_tiles + 1
    float *o_prob; uint32_t *o_lpos; double *o_aux; uint4 *o_heads; uint32_t *o_chunk_row; uint32_t *o_chunk_info; uint4 *o_meta;
    uint32_t *o_table; uint32_t *o_unit_txp; uint8_t *o_unit_cnt;
    uint32_t *o_trow;          // tile-order row -> original row
    uint32_t *fallback; uint32_t *cursors;  // [0] fallback rows, [1] table entries, [2] units
};

// One CTA lays out one tile.
static __global__ void __launch_bounds__(kThreads) build_tiles(BuildArgs a)
{
    using Sort = cub::BlockRadixSort<uint32_t, kThreads, 4, uint32_t>;
    using Scan = cub::BlockScan<uint32_t, kThreads>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ uint32_t s_txp[kTile];      // slot -> transcript; later reused as sorted keys
    __shared__ uint32_t s_src[kTile];      // slot -> source alignment index in the CSR
    __shared__ uint32_t s_lpos[kTile];
    __shared__ uint32_t s_seg[kTile + 1];  // segment -> first sorted rank; later unit base
    __shared__ uint16_t s_rlen[kTile], s_rslot[kTile], s_rnew[kTile];
    __shared__ uint32_t s_heads[kWarps * 4];
    __shared__ uint32_t s_used[kWarps], s_nrow[kWarps], s_info[kWarps];
    __shared__ uint32_t s_misc[4];

    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    const uint32_t nrows = min(r1 - r0, (uint32_t)kTile);  // every row has >= 1 alignment and span <= kTile

    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) { s_txp[i] = kNoTxp; s_lpos[i] = 0; s_src[i] = kNoTxp; }
    if (tid < kWarps * 4) s_heads[tid] = 0;
    for (uint32_t i = tid; i < nrows; i += kThreads) {
        const uint32_t r = a.srow[r0 + i];
        s_rlen[i] = (uint16_t)(a.row_ptr[r + 1] - a.row_ptr[r]);
    }
    __syncthreads();

    // first-fit packing of rows into warp-chunks (rows never straddle a chunk)
    if (tid == 0) {
        uint32_t used[kWarps], cnt[kWarps], span[kWarps];
#pragma unroll
        for (int c = 0; c < kWarps; ++c) { used[c] = 0; cnt[c] = 0; span[c] = 0; }
        for (uint32_t i = 0; i < nrows; ++i) {
            const uint32_t len = s_rlen[i];
            int pick = -1;
#pragma unroll
            for (int c = 0; c < kWarps; ++c) if (pick < 0 && used[c] + len <= (uint32_t)kChunkCap) pick = c;
            if (pick < 0) { s_rslot[i] = 0xFFFF; continue; }
#pragma unroll
            for (int c = 0; c < kWarps; ++c) if (c == pick) {
                s_rslot[i] = (uint16_t)(c * kChunk + used[c]);
                s_rnew[i] = (uint16_t)cnt[c];  // order inside the chunk
                const uint32_t lanes = ((used[c] + len - 1) >> 2) - (used[c] >> 2);  // lanes the row's tail must travel
                span[c] = max(span[c], lanes);
                used[c] += len; cnt[c] += 1;
            }
        }
#pragma unroll
        for (int c = 0; c < kWarps; ++c) {
            s_used[c] = used[c]; s_nrow[c] = cnt[c];
            uint32_t steps = 0;
            while ((1u << steps) <= span[c]) ++steps;  // Hillis-Steele steps 1,2,..,2^(steps-1) cover `span` lanes
            s_info[c] = steps;
        }
    }
    __syncthreads();
    // chunk row bases (tile order = chunk by chunk), overflow rows go last and to the fallback list
    uint32_t chunk_base[kWarps];
    {
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < kWarps; ++c) { chunk_base[c] = acc; acc += s_nrow[c]; }
        if (tid == 0) s_misc[0] = acc;  // rows placed
        if (tid < kWarps) {
            a.o_chunk_row[tile * kWarps + tid] = r0 + chunk_base[tid];
            // the padding slots form a pseudo row, so every real row ends at a head inside the chunk
            const uint32_t u = s_used[tid];
            atomicOr(&s_heads[tid * 4 + (u >> 5)], 1u << (u & 31));
        }
    }
    __syncthreads();
    const uint32_t placed = s_misc[0];
    for (uint32_t i = tid; i < nrows; i += kThreads) {
        const uint32_t r = a.srow[r0 + i];
        const uint32_t slot = s_rslot[i];
        if (slot == 0xFFFFu) continue;  // did not fit: handled below
        const uint32_t c = slot / kChunk;
        a.o_trow[r0 + chunk_base[c] + s_rnew[i]] = r;
        const uint32_t s = a.row_ptr[r], len = s_rlen[i];
        atomicOr(&s_heads[slot >> 5], 1u << (slot & 31));
        // the order of a row's alignments is free: sorting them by transcript makes neighbouring rows
        // hit neighbouring smem banks in the M-step scatter (insertion sort; rows are short)
        for (uint32_t j = 0; j < len; ++j) {
            const uint32_t t = a.txp[s + j];
            uint32_t k = j;
            while (k > 0 && s_txp[slot + k - 1] > t) {
                s_txp[slot + k] = s_txp[slot + k - 1]; s_src[slot + k] = s_src[slot + k - 1]; --k;
            }
            s_txp[slot + k] = t; s_src[slot + k] = s + j;
        }
    }
    // overflow rows: serial append by thread 0 (rare)
    if (tid == 0 && placed < nrows) {
        uint32_t k = placed;
        for (uint32_t i = 0; i < nrows; ++i) if (s_rslot[i] == 0xFFFFu) {
            const uint32_t r = a.srow[r0 + i];
            a.fallback[atomicAdd(a.cursors + 0, 1u)] = r;
            a.o_trow[r0 + k++] = r;
        }
    }
    __syncthreads();
    // lanes holding two or more row heads need the general segmented-sum path
    {
        const uint32_t nib = (s_heads[tid >> 3] >> ((tid & 7u) * 4u)) & 0xFu;  // thread t <-> lane t of the tile
        if (__popc(nib) >= 2) atomicOr(&s_info[tid >> 5], kInfoMulti);
    }
    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) {
        const uint32_t src = s_src[i];
        a.o_prob[(size_t)tile * kTile + i] = src != kNoTxp ? a.prob[src] : 0.f;   // padding: prob 0 (aux 1)
        if (a.aux) a.o_aux[(size_t)tile * kTile + i] = src != kNoTxp ? a.aux[src] : 1.0;
    }

    // sort slots by transcript
    uint32_t keys[4], vals[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { keys[i] = s_txp[tid * 4 + i]; vals[i] = tid * 4 + i; }
    __syncthreads();
    Sort(tmp.sort).Sort(keys, vals);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) s_txp[tid * 4 + i] = keys[i];
    __syncthreads();
    // segments of equal transcript
    uint32_t hf[4], seg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = tid * 4 + i;
        hf[i] = (keys[i] != kNoTxp && (r == 0 || s_txp[r - 1] != keys[i])) ? 1u : 0u;
    }
    uint32_t D = 0;
    Scan(tmp.scan).InclusiveSum(hf, seg, D);
    __syncthreads();
    uint32_t nvalid_local = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (hf[i]) s_seg[seg[i] - 1] = tid * 4 + i;
        if (keys[i] != kNoTxp) ++nvalid_local;
    }
    uint32_t nvalid = 0;
    {
        uint32_t dummy;
        Scan(tmp.scan).ExclusiveSum(nvalid_local, dummy, nvalid);
    }
    __syncthreads();
    if (tid == 0) s_seg[D] = nvalid;
    __syncthreads();
    // units: aggregated segments get ceil(cnt/8) 8-slot units
    uint32_t nun[4], ubase[4], U = 0, cntd[4], startd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        nun[i] = 0; cntd[i] = 0; startd[i] = 0;
        if (d < D) {
            startd[i] = s_seg[d];
            cntd[i] = s_seg[d + 1] - startd[i];
            nun[i] = cntd[i] >= (uint32_t)kAggMin ? (cntd[i] + 7) >> 3 : 0;
        }
    }
    Scan(tmp.scan).ExclusiveSum(nun, ubase, U);
    __syncthreads();
    if (tid == 0) {
        s_misc[2] = atomicAdd(a.cursors + 1, D);
        s_misc[3] = atomicAdd(a.cursors + 2, U);
    }
    __syncthreads();
    const uint32_t table_off = s_misc[2], unit_off = s_misc[3];
    // per-segment outputs; stash (start, unit base) for the per-slot pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        if (d < D) {
            const uint32_t key = s_txp[startd[i]];
            a.o_table[table_off + d] = key;
            for (uint32_t v = 0; v < nun[i]; ++v) {
                a.o_unit_txp[unit_off + ubase[i] + v] = key;
                a.o_unit_cnt[unit_off + ubase[i] + v] = (uint8_t)min(8u, cntd[i] - 8u * v);
            }
        }
    }
    __syncthreads();  // everyone has read s_seg[d+1]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        if (d < D) s_seg[d] = startd[i] | ((nun[i] ? ubase[i] : 0x1FFu) << 11);  // start (11 bits) | unit base (9 bits, 0x1FF = stray)
    }
    __syncthreads();
    // per sorted element: table index and position, as shared-memory byte offsets
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = tid * 4 + i;
        if (keys[i] != kNoTxp) {
            const uint32_t d = seg[i] - 1;
            const uint32_t pk = s_seg[d];
            const uint32_t start = pk & 0x7FFu, ub = pk >> 11;
            uint32_t pos = kTrash;
            if (ub != 0x1FFu) { const uint32_t p = ub * 8 + (r - start); pos = p + (p >> 3); }
            else atomicOr(&s_info[vals[i] / kChunk], kInfoStray);
            s_lpos[vals[i]] = (d * 8u) | ((pos * 8u) << 16);
        } else {
            s_lpos[vals[i]] = 0u | (((uint32_t)kTrash * 8u) << 16);
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) a.o_lpos[(size_t)tile * kTile + i] = s_lpos[i];
    if (tid < kWarps) {
        a.o_heads[tile * kWarps + tid] = make_uint4(s_heads[tid * 4], s_heads[tid * 4 + 1], s_heads[tid * 4 + 2], s_heads[tid * 4 + 3]);
        a.o_chunk_info[tile * kWarps + tid] = s_info[tid];
    }
    if (tid == 0) a.o_meta[tile] = make_uint4(table_off, unit_off, D | (U << 16), r0);
}

// bootstrap weights from read order into tile order
static __global__ void permute_weights(const uint32_t *__restrict__ w, const uint32_t *__restrict__ trow, uint32_t n,
                                uint32_t *__restrict__ wperm)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) wperm[k] = w[trow[k]];
}

// ---------------------------------------------------------------------------
// the sweep
// ---------------------------------------------------------------------------

__device__ __forceinline__ double fast_rcp(double d)
{   // MUFU.RCP64H seed (~20 bits) + two Newton steps: <= 1-2 ulp for normal d
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__device__ __forceinline__ float4 ld_stream_f4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Everything one lane holds of its warp-chunk: 4 alignment slots.
template <bool HAS_AUX>
struct ChunkRegs {
    float4 p4;        // conditional probabilities
    uint4 lp4;        // (table index * 8) | (pos * 8) << 16
    uint4 hm;         // 128-bit row-head mask of the chunk
    uint32_t info;    // scan steps | kInfoStray | kInfoMulti
    uint32_t row_base;
    double a0, a1, a2, a3;
};

template <bool HAS_AUX, bool HAS_WTS>
__device__ __forceinline__ void load_chunk(const View &v, uint32_t tile, uint32_t warp, uint32_t lane,
                                           ChunkRegs<HAS_AUX> &c)
{
    const size_t base = (size_t)tile * kTile + warp * kChunk + lane * 4;
    c.p4 = ld_stream_f4(v.prob + base);
    c.lp4 = ld_stream_u4(v.lpos + base);
    c.hm = v.heads[tile * kWarps + warp];
    c.info = v.chunk_info[tile * kWarps + warp];
    c.row_base = 0;
    if (HAS_WTS) c.row_base = v.chunk_row[tile * kWarps + warp];
    if (HAS_AUX) {
        const double2 q0 = *reinterpret_cast<const double2 *>(v.aux + base);
        const double2 q1 = *reinterpret_cast<const double2 *>(v.aux + base + 2);
        c.a0 = q0.x; c.a1 = q0.y; c.a2 = q1.x; c.a3 = q1.y;
    }
}

// Phase 1 for one warp-chunk: E-step in registers (per-row denominators by a
// segmented warp scan), then the M-step scatter into the transcript-sorted
// shared-memory order.
template <bool HAS_AUX, bool HAS_WTS>
__device__ __forceinline__ void chunk_phase1(const ChunkRegs<HAS_AUX> &c, const View &v, uint32_t table_off,
                                             const double *s_prev, double *xs, double *__restrict__ curr,
                                             const uint32_t *__restrict__ wperm, uint32_t lane)
{
    const uint4 lp4 = c.lp4;
    const uint4 hm = c.hm;
    const uint32_t info = c.info;
    const char *sp = reinterpret_cast<const char *>(s_prev);
    double w0 = *reinterpret_cast<const double *>(sp + (lp4.x & 0xFFFFu)) * (double)c.p4.x;
    double w1 = *reinterpret_cast<const double *>(sp + (lp4.y & 0xFFFFu)) * (double)c.p4.y;
    double w2 = *reinterpret_cast<const double *>(sp + (lp4.z & 0xFFFFu)) * (double)c.p4.z;
    double w3 = *reinterpret_cast<const double *>(sp + (lp4.w & 0xFFFFu)) * (double)c.p4.w;
    if (HAS_AUX) { w0 *= c.a0; w1 *= c.a1; w2 *= c.a2; w3 *= c.a3; }

    const uint32_t wq = lane >> 3;
    const uint32_t hword = wq == 0 ? hm.x : wq == 1 ? hm.y : wq == 2 ? hm.z : hm.w;
    const uint32_t hb = (hword >> ((lane & 7u) * 4u)) & 0xFu;
    const unsigned full = 0xffffffffu;
    const unsigned lanes_h = __ballot_sync(full, hb != 0u);          // lane 0 always has a head
    const int P = 31 - __clz(lanes_h & (full >> (31u - lane)));       // nearest lane <= me holding a head
    const unsigned later = lanes_h & ~(full >> (31u - lane));        // lanes after me holding a head
    const int E = later ? (__ffs(later) - 1) : (int)lane;
    const int nsteps = (int)(info & 7u);

    double x0, x1, x2, x3;
    if (!(info & kInfoMulti)) {
        // fast path: no lane holds more than one row head.  a = my slots before the head (they
        // close the row entering this lane), z = my slots from the head on (they open a row).
        const bool b0 = !(hb & 1u), b1 = !(hb & 3u), b2 = !(hb & 7u), b3 = !(hb & 15u);
        double a = b0 ? w0 : 0.0, z = b0 ? 0.0 : w0;
        a += b1 ? w1 : 0.0; z += b1 ? 0.0 : w1;
        a += b2 ? w2 : 0.0; z += b2 ? 0.0 : w2;
        a += b3 ? w3 : 0.0; z += b3 ? 0.0 : w3;
        double incl = hb ? z : a;                                     // what this lane adds to the open row
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i >= nsteps) break;                                   // warp-uniform
            const int d = 1 << i;
            const double t = __shfl_up_sync(full, incl, d);
            if ((int)lane - d >= P) incl += t;
        }
        double carry = __shfl_up_sync(full, incl, 1);                 // sum of the row entering this lane
        if (lane == 0) carry = 0.0;
        const double t_in = carry + a;                                // its total, if it ends here (hb != 0)
        // reads whose denominator is <= 1e-30 contribute nothing (em.rs:115)
        double inv_in = t_in > OAR_EM_DENOM_THRESH ? fast_rcp(t_in) : 0.0;
        double inv_out = __shfl_sync(full, inv_in, E);                // the row leaving this lane ends in lane E
        if (!later) inv_out = 0.0;                                    // only padding lies beyond the last head
        if (!hb) inv_in = inv_out;                                    // a lane without a head is inside one row
        x0 = w0 * (b0 ? inv_in : inv_out);
        x1 = w1 * (b1 ? inv_in : inv_out);
        x2 = w2 * (b2 ? inv_in : inv_out);
        x3 = w3 * (b3 ? inv_in : inv_out);
    } else {
        // general path: rows may start and end inside one lane
        const double s0 = w0;
        const double s1 = (hb & 2u) ? w1 : s0 + w1;
        const double s2 = (hb & 4u) ? w2 : s1 + w2;
        const double s3 = (hb & 8u) ? w3 : s2 + w3;
        double incl = s3;                                             // segmented inclusive scan of lane tails
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i >= nsteps) break;
            const int d = 1 << i;
            const double t = __shfl_up_sync(full, incl, d);
            if ((int)lane - d >= P) incl += t;
        }
        double carry = __shfl_up_sync(full, incl, 1);
        if (lane == 0) carry = 0.0;
        const double sA = (hb & 1u) ? 0.0 : (hb & 2u) ? s0 : (hb & 4u) ? s1 : (hb & 8u) ? s2 : s3;
        const double t_in = carry + sA;
        const double t_in_e = __shfl_sync(full, t_in, E);
        const double t_out = later ? t_in_e : 0.0;
        const double pre0 = (hb & 1u) ? s0 : carry + s0;
        const double pre1 = (hb & 3u) ? s1 : carry + s1;
        const double pre2 = (hb & 7u) ? s2 : carry + s2;
        const double tot3 = t_out;
        const double tot2 = (hb & 8u) ? pre2 : tot3;
        const double tot1 = (hb & 4u) ? pre1 : tot2;
        const double tot0 = (hb & 2u) ? pre0 : tot1;
        x0 = tot0 > OAR_EM_DENOM_THRESH ? w0 * fast_rcp(tot0) : 0.0;
        x1 = tot1 > OAR_EM_DENOM_THRESH ? w1 * fast_rcp(tot1) : 0.0;
        x2 = tot2 > OAR_EM_DENOM_THRESH ? w2 * fast_rcp(tot2) : 0.0;
        x3 = tot3 > OAR_EM_DENOM_THRESH ? w3 * fast_rcp(tot3) : 0.0;
    }

    if (HAS_WTS) {
        // row index inside the chunk = (number of heads at or before the slot) - 1
        uint32_t before = 0;
        if (wq > 0) before += __popc(hm.x);
        if (wq > 1) before += __popc(hm.y);
        if (wq > 2) before += __popc(hm.z);
        const uint32_t sh = (lane & 7u) * 4u;
        const uint32_t r0 = before + __popc(hword & (full >> (31u - sh))) - 1u;
        const uint32_t r1 = r0 + ((hb >> 1) & 1u), r2 = r1 + ((hb >> 2) & 1u), r3 = r2 + ((hb >> 3) & 1u);
        x0 *= (double)wperm[c.row_base + r0];
        x1 *= (double)wperm[c.row_base + r1];
        x2 *= (double)wperm[c.row_base + r2];
        x3 *= (double)wperm[c.row_base + r3];
    }

    char *xp = reinterpret_cast<char *>(xs);
    *reinterpret_cast<double *>(xp + (lp4.x >> 16)) = x0;
    *reinterpret_cast<double *>(xp + (lp4.y >> 16)) = x1;
    *reinterpret_cast<double *>(xp + (lp4.z >> 16)) = x2;
    *reinterpret_cast<double *>(xp + (lp4.w >> 16)) = x3;
    if (info & kInfoStray) {
        // transcripts with fewer than kAggMin alignments in this tile: straight to global
        const uint32_t trash = (uint32_t)kTrash * 8u;
        if ((lp4.x >> 16) == trash && x0 != 0.0) atomicAdd(curr + v.table[table_off + ((lp4.x & 0xFFFFu) >> 3)], x0);
        if ((lp4.y >> 16) == trash && x1 != 0.0) atomicAdd(curr + v.table[table_off + ((lp4.y & 0xFFFFu) >> 3)], x1);
        if ((lp4.z >> 16) == trash && x2 != 0.0) atomicAdd(curr + v.table[table_off + ((lp4.z & 0xFFFFu) >> 3)], x2);
        if ((lp4.w >> 16) == trash && x3 != 0.0) atomicAdd(curr + v.table[table_off + ((lp4.w & 0xFFFFu) >> 3)], x3);
    }
}

// Phase 2 for one warp: sum 8-slot units of the transcript-sorted order, combine
// equal transcripts across the warp, one f64 RED per (warp, transcript).
__device__ __forceinline__ void units_phase2(const double *xs, uint32_t tid, uint32_t lane, uint32_t u_txp,
                                             uint32_t u_cnt, double *__restrict__ curr)
{
    const unsigned full = 0xffffffffu;
    double acc = 0.0;
    const double *b = xs + tid * 9;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) if (k < u_cnt) acc += b[k];
    const uint32_t up = __shfl_up_sync(full, u_txp, 1);
    const uint32_t dn = __shfl_down_sync(full, u_txp, 1);
    const bool head = (lane == 0) || (up != u_txp);
    const bool tail = (lane == 31) || (dn != u_txp);
    const unsigned hmask = __ballot_sync(full, head);
    const int P2 = 31 - __clz(hmask & (full >> (31u - lane)));
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(full, acc, d);
        if ((int)lane - d >= P2) acc += t;
    }
    if (tail && u_txp != kNoTxp) atomicAdd(curr + u_txp, acc);
}

// m_step (em.rs:87-133) over one tile per CTA.
template <bool HAS_AUX, bool HAS_WTS>
__global__ void __launch_bounds__(kThreads) em_sweep_tiled(View v, const double *__restrict__ prev,
                                                           double *__restrict__ curr,
                                                           const uint32_t *__restrict__ wperm,
                                                           const OarEmState *__restrict__ st, int check_done)
{
    __shared__ __align__(16) double xs[kTrash + 1];
    __shared__ __align__(16) double s_prev[kTile];

    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t done = 0;
    if (check_done) done = st->done;   // tested after the loads below are in flight
    const uint4 meta = v.meta[tile];
    ChunkRegs<HAS_AUX> c;
    load_chunk<HAS_AUX, HAS_WTS>(v, tile, warp, lane, c);   // streaming loads first: they overlap the table gather
    const uint32_t D = meta.z & 0xFFFFu, U = meta.z >> 16;
    uint32_t u_txp = kNoTxp, u_cnt = 0;
    if (tid < U) { u_txp = v.unit_txp[meta.y + tid]; u_cnt = v.unit_cnt[meta.y + tid]; }
    if (done) return;
    for (uint32_t d = tid; d < D; d += kThreads) s_prev[d] = prev[v.table[meta.x + d]];
    __syncthreads();
    chunk_phase1<HAS_AUX, HAS_WTS>(c, v, meta.x, s_prev, xs, curr, wperm, lane);
    __syncthreads();
    if (warp * 32u < U) units_phase2(xs, tid, lane, u_txp, u_cnt, curr);
}

}  // namespace tiled
}  // namespace oar
