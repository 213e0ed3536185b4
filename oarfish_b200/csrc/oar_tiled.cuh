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
//     table offset instead of the 32-bit id;
//   * the M-step scatter goes through shared memory: every alignment also
//     stores `pos`, one of its transcript's x positions in the tile (chosen at
//     layout time so that the 16 lanes of a half-warp store into 16 different
//     banks), x_j = w_j/denom is written to xs[pos]; a transcript's positions
//     form ITEMS of <= 16 consecutive slots, one thread sums one item and
//     flushes it with ONE f64 RED instead of one per alignment.  No shared-
//     memory atomics (f64 smem atomics are CAS loops on sm_100a).
//   * rows never straddle a 128-slot warp-chunk, so the per-row denominator
//     (em.rs:98-112) is a segmented warp scan in registers, driven by a
//     precomputed 16-bit descriptor per lane.
//
// The sweep is a persistent kernel: CTAs walk tiles blockIdx.x, +gridDim.x, ...
// and a two-stage ring of shared-memory buffers is filled by TMA bulk copies
// (cp.async.bulk, completion on an mbarrier), issued two tiles ahead by one
// thread, so HBM latency never sits on the compute path and no registers are
// spent on prefetching.  One warp alone waits on the mbarriers and gathers
// prev[]; the CTA barrier hands the stage on to the others (two CTA barriers per
// tile; the single-barrier and streaming variants tried in rounds 1 and 2 were
// slower, profiles/experiments/).
//
// Per alignment the HBM stream is 4 B (prob f32) + 4 B (table offset u16 | pos
// u16), the same 8 B as CSR's txp_id + prob; row structure costs 2 B per lane
// (4 slots) instead of a 4-byte row_ptr entry per row.
//
// Rows longer than a warp-chunk, or that do not fit their tile, are listed in
// `fallback_rows` and swept from the original CSR, every CTA of the sweep taking
// its share (a very long list gets its own em_sweep_rowgroup launch).
#pragma once
#include <cub/cub.cuh>

#include "oar_common.cuh"
#include "oar_kernels.cuh"

namespace oar {
namespace tiled {

#ifndef OAR_TILE_WARPS
#define OAR_TILE_WARPS 8
#endif
constexpr int kWarps = OAR_TILE_WARPS;    // warp-chunks per tile
constexpr int kChunk = 128;               // alignment slots per warp-chunk (4 per lane)
constexpr int kChunkCap = kChunk - 1;     // rows use at most 127 slots: a padding pseudo-row always closes the chunk
constexpr int kTile = kWarps * kChunk;    // 1024 slots
constexpr int kThreads = kWarps * 32;     // 256
constexpr int kAggMin = 4;                // transcripts with >= kAggMin alignments in a tile are aggregated in smem
#ifndef OAR_ITEM_MAX
#define OAR_ITEM_MAX 16
#endif
#ifndef OAR_TILED_MIN_CTAS
#define OAR_TILED_MIN_CTAS 5    // register budget of the sweep with an aux factor or a tile list: 48 registers per thread, 5 CTAs per SM ...
#endif
#ifndef OAR_TILED_MIN_CTAS_PLAIN
#define OAR_TILED_MIN_CTAS_PLAIN 6   // ... and 40 registers, 6 CTAs per SM without them: the plain sweep fits with 0-12 bytes of spills (C3: 170-172.7
#endif                               // vs 173.2 us), the bootstrap-weighted one with 4-32 bytes since its lane weights sit in front of the record (173.5 vs 175.5 us)
#ifndef OAR_SCATTER_GREEDY
#define OAR_SCATTER_GREEDY 1    // layout: x positions chosen so that the M-step scatter spreads over the banks
#endif
#ifndef OAR_GREEDY_SCARCE
#define OAR_GREEDY_SCARCE 1     // layout: in the x position greedy the lane whose transcript offers the fewest residues wins a contested bank
#endif
#ifndef OAR_GREEDY_SUPPLY
#define OAR_GREEDY_SUPPLY 1     // layout: a lane takes the candidate residue its transcript has most slots of left
#endif
#ifndef OAR_SCAN_COND
#define OAR_SCAN_COND 1         // sweep: the 4-lane step of the segmented scan only in chunks that hold a row spanning more than 4 lanes
#endif
#ifndef OAR_SHORT_ROW_DEFER
#define OAR_SHORT_ROW_DEFER 1   // layout: keep rows shorter than a lane from starting and ending inside one lane
#endif
__host__ __device__ constexpr int sweep_ctas(bool aux, bool /*wts*/, bool list) { return (!aux && !list) ? OAR_TILED_MIN_CTAS_PLAIN : OAR_TILED_MIN_CTAS; }
constexpr int kItemMax = OAR_ITEM_MAX;    // x slots of the largest item size class; the classes are kItemMax, /2, /4
constexpr int kMaxItems = kTile / 4;      // every item belongs to a transcript with >= 4 alignments and holds >= 4 of them
// An aggregated transcript with cnt alignments in the tile owns cnt / kItemMax items of kItemMax consecutive x slots
// and one item for the remainder, of the smallest size class (kItemMax, /2, /4 slots) that holds it.  Items are
// ordered by class (largest first); a class-c item sits c + 2 doubles behind its predecessor, so the 8 lanes of an
// LDS.128 phase hit 8 different 16-byte banks.  One thread sums one item and issues one RED: no cross-lane scan, no
// padding to clear.  (Measured on C3: 16-slot items 228 us, 32-slot items 232 us, 8-slot units + scan 244 us.)
// padding and non-aggregated alignments write to the trash slots right after the last item of the tile (kTrashSlots)
// prev[] of a tile's transcripts sits in shared memory as SPLIT 32-bit words: blocks of 32 table entries, 128 bytes
// of high words followed by 128 bytes of low words.  A warp-wide LDS.32 is one wavefront whenever its lanes hit 32
// different banks or the same word, so the E-step's gather is conflict-free for tiles with up to 32 distinct
// transcripts (C3: 24 on average); the former LDS.64 from an array of doubles conflicted from 17 entries on
// (98 wavefronts per tile against 64).  An alignment stores the byte offset of its entry's high word:
#ifndef OAR_PREV_SPLIT
#define OAR_PREV_SPLIT 1        // 0: prev[] as an array of doubles read with LDS.64 (A/B timing only)
#endif
#if OAR_PREV_SPLIT
__host__ __device__ constexpr uint32_t table_off(uint32_t d) { return ((d >> 5) << 8) | ((d & 31u) << 2); }
__host__ __device__ constexpr uint32_t table_index(uint32_t off) { return ((off >> 8) << 5) | ((off & 127u) >> 2); }
__host__ __device__ constexpr uint32_t table_bytes(uint32_t max_d) { return 256u * ((max_d + 31u) / 32u); }
#else
__host__ __device__ constexpr uint32_t table_off(uint32_t d) { return 8u * d; }
__host__ __device__ constexpr uint32_t table_index(uint32_t off) { return off >> 3; }
__host__ __device__ constexpr uint32_t table_bytes(uint32_t max_d) { return 8u * ((max_d + 1u) & ~1u); }
#endif
constexpr uint32_t kPrevLo = 128u;        // low word = high word + 128 bytes
constexpr uint32_t kInfoStray = 8u;       // chunk_info bit 3: chunk holds alignments that RED straight to global
constexpr uint32_t kInfoMulti = 16u;      // chunk_info bit 4: some lane holds >= 2 row heads (general path)
constexpr uint32_t kNoTxp = 0xFFFFFFFFu;
constexpr uint32_t kMaxTxps = 0xFFFFFFFFu; // kNoTxp is reserved
constexpr uint32_t kDescRare = 1u << 9, kDescMid = 1u << 10;
constexpr uint32_t kTrashSlots = 16;      // x slots behind the tile's items that take the stores of padding and stray alignments: per half-warp store the one in a bank the other lanes leave free
static_assert(kMaxItems == kThreads, "one item per thread in phase 2");
static_assert(!OAR_SCATTER_GREEDY || kItemMax == 16, "the bank-aware x positions assume 16-slot items at a stride of 18 doubles: one slot per 8-byte bank residue");

// Per-tile record (variable length, 16-byte granules), one TMA bulk copy:
//   [0,512)    lane descriptors u16[8][32]: hb(4) | dist(5) << 4 | rare << 9 | mid << 10 | E(5) << 11
//                hb    row-head bits of the lane's 4 slots
//                dist  lanes back to the nearest lane (<= this one) holding a head
//                rare  (same in all lanes of a chunk) the chunk needs the general copy of phase 1: a lane with two
//                      heads, stray alignments or a row spanning more than 8 lanes
//                mid   (same in all lanes of a chunk) some row spans more than 4 lanes: third scan step needed
//                E     first later lane holding a head (where the row leaving this lane ends);
//                      the lane itself if there is none (then only padding follows).  Top bits: `desc >> 11` is
//                      the shuffle's source lane without a mask
//              The sweep tests the fields in place (`desc & 0xE`, `desc & 0x1E0`, ...: one LOP3 with predicate
//              output each) instead of extracting them first.
//   [512,544)  chunk_info u32[8]: scan steps (bits 0-2) | kInfoStray | kInfoMulti
//   [544,576)  chunk_row  u32[8]: tile-order index of the chunk's first row (bootstrap weights)
//   [576,592)  D, byte offset of the table in the record, items, trash x offset in bytes
//   [592, ..)  items u32[roundup4(items)]  : 4 * table index | (x offset in doubles) << 12 | slots << 24   (0 = no item)
//              table u32[roundup4(D)]      : distinct transcript ids
//              (items first: thread t's item sits at a fixed offset, 592 + 4 t)
constexpr int kRecDesc = 0, kRecInfo = 64 * kWarps, kRecRow = kRecInfo + 4 * kWarps, kRecDU = kRecRow + 4 * kWarps,
              kRecTable = kRecDU + 16;
constexpr int kRecMax = kRecTable + 4 * kTile + 4 * kMaxItems;   // 5712
#ifndef OAR_REC_ALIGN
#define OAR_REC_ALIGN 128
#endif
constexpr uint32_t kRecAlign = OAR_REC_ALIGN;   // alignment of a record in global memory (the TMA copy's source)
constexpr int kStages = 2;

// Shared-memory geometry of the sweep, sized for the store at hand (largest record, table and
// unit count over all tiles) so that as many CTAs as possible fit on an SM.
struct Geometry {
    uint32_t stage_bytes;   // prob (4 KB) | lpos (4 KB) | [lane weights u16[256] (bootstrap) |] record (max over tiles, rounded up to 128 B)
    uint32_t xs_base;       // shared-space address of the dynamic window = of the x array (transcript-sorted x values in
                            // items + trash), which sits at its start.  Filled in by the launcher (1 KB reserved by the
                            // system + the kernel's static shared memory) and handed over as a kernel PARAMETER so that it
                            // is warp-uniform by construction: the scatter and the item reads are `[offset + UR]`.  Left to
                            // the compiler (a cvta on the extern array) the base was rematerialised per tile either with
                            // S2UR + ULEA or -- after unrelated changes to the parameter list -- with S2R + LEA and four
                            // vector adds per thread: 185 vs 191 us.  The kernel traps if the address is not the real one.
    uint32_t stage_off;     // the two stages
    uint32_t prev_off;      // prev[] of the tile's transcripts
    uint32_t bar_off;       // two mbarriers
    uint32_t total;         // dynamic shared memory per CTA
    uint32_t xs_doubles;    // doubles to clear at start
};
inline Geometry make_geometry(uint32_t max_rec_bytes, uint32_t max_d, uint32_t max_x_doubles, bool weighted)
{
    Geometry g;
#ifndef OAR_STAGE_ALIGN
#define OAR_STAGE_ALIGN 128u    // every TMA destination of a stage starts on a 128-byte line
#endif
    const uint32_t rec = (max_rec_bytes + OAR_STAGE_ALIGN - 1u) & ~(OAR_STAGE_ALIGN - 1u);
    g.stage_bytes = 8u * kTile + (weighted ? 2u * kThreads : 0u) + rec;   // the weights sit right in front of the record: one base register serves both
    g.xs_base = 0;   // set by the launcher
    // the items of the fullest tile, then the trash slots; even count
    g.xs_doubles = (max_x_doubles + kTrashSlots + 1u) & ~1u;
    g.stage_off = (8u * g.xs_doubles + 127u) & ~127u;
    g.prev_off = g.stage_off + kStages * g.stage_bytes;
    g.bar_off = g.prev_off + table_bytes(max_d);   // split high / low words, see table_off()
    g.total = g.bar_off + 16u;
    return g;
}

struct View {
    uint32_t n_tiles;
    const float *prob;         // n_tiles * kTile
    const uint32_t *lpos;      // n_tiles * kTile : table_off(table index) | (pos * 8) << 16  (smem byte offsets)
    const double *aux;         // n_tiles * kTile or null
    const uint2 *rec;          // n_tiles : {record offset in 16-byte granules, record bytes}
    const uint4 *records;      // all records
    const uint16_t *wlane;     // n_tiles * kThreads: bootstrap weight of the row that ends in each lane (lane_weights) or null
    const uint32_t *tile_list; const uint32_t *n_active;   // LIST sweeps: the tiles to walk and how many (device memory)
    // rows that are not tiled (longer than a chunk, or did not fit): swept from the CSR, spread over all CTAs
    const uint32_t *fb_rows; uint32_t n_fb;
    const uint32_t *csr_row_ptr; const uint32_t *csr_txp; const float *csr_prob; const double *csr_aux;
    const uint32_t *csr_wts;   // bootstrap weights in read order (fallback rows are not in tile order)
};

// ---------------------------------------------------------------------------
// layout construction
// ---------------------------------------------------------------------------

// key = smallest transcript id of the row (locality key); rows that cannot be
// tiled (empty, or longer than a warp-chunk) get kNoTxp and sort to the end.
// `vid` (optional): a locality numbering of the transcripts (cluster_ids below) used INSTEAD of the ids for the key,
// for stores whose transcript ids are not gene-local.
static __global__ void row_keys(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp, uint64_t n_rows,
                                const uint32_t *__restrict__ vid, uint32_t *__restrict__ key, uint32_t *__restrict__ idx,
                                uint32_t *__restrict__ counters)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t n_long = 0, n_skip = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        uint32_t k = kNoTxp;
        if (e > s && e - s <= (uint32_t)kChunkCap) {
            if (vid) { for (uint32_t j = s; j < e; ++j) k = min(k, vid[txp[j]]); }
            else     { for (uint32_t j = s; j < e; ++j) k = min(k, txp[j]); }
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

// ---- transcript ids that are not gene-local --------------------------------------------------------------
// The row order above clusters the reads of a gene only if its isoforms have neighbouring ids.  When they do not (a
// reference not grouped by gene; tools/bench_robust.py "permuted"), the layout first clusters the transcripts by
// co-occurrence: min-label propagation over the reads (every read pulls its transcripts to the smallest label among
// them) with pointer jumping, a few rounds -- genes are small, dense components -- then transcripts are numbered by
// (label, id) and the rows are keyed by that number.  For gene-local ids the numbering is the identity, and the
// whole step is skipped when few rows have widely spread ids (row_id_spread).

// counters[0] += rows whose ids span more than 16 * (alignments) + 64: not what reads of one gene look like
static __global__ void row_id_spread(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp, uint64_t n_rows,
                                     uint32_t *__restrict__ counters)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t n = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        if (e - s < 2u) continue;
        uint32_t lo = kNoTxp, hi = 0;
        for (uint32_t j = s; j < e; ++j) { const uint32_t t = txp[j]; lo = min(lo, t); hi = max(hi, t); }
        if (hi - lo > 16u * (e - s) + 64u) ++n;
    }
    if (n) atomicAdd(counters, n);
}

static __global__ void label_init(uint32_t *__restrict__ label, uint32_t M)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < M; t += stride) label[t] = t;
}

// one round: every row pulls the labels of its transcripts (and of their current labels: hooking) down to the row's minimum
static __global__ void label_rows(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp, uint64_t n_rows,
                                  uint32_t *__restrict__ label)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        if (e - s < 2u) continue;
        uint32_t m = kNoTxp;
        for (uint32_t j = s; j < e; ++j) m = min(m, label[txp[j]]);
        for (uint32_t j = s; j < e; ++j) {
            const uint32_t t = txp[j], l = label[t];
            if (l > m) { atomicMin(label + t, m); atomicMin(label + l, m); }
        }
    }
}

static __global__ void label_jump(uint32_t *__restrict__ label, uint32_t M)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < M; t += stride) {
        uint32_t l = label[t], ll = label[l];
        while (ll != l) { l = ll; ll = label[l]; }   // labels only decrease and label[x] <= x: the chain ends at a root
        label[t] = l;
    }
}

static __global__ void label_keys(const uint32_t *__restrict__ label, uint32_t M, uint64_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < M; t += stride) { key[t] = ((uint64_t)label[t] << 32) | t; idx[t] = t; }
}

static __global__ void label_number(const uint32_t *__restrict__ sorted_t, uint32_t M, uint32_t *__restrict__ vid)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) vid[sorted_t[i]] = i;
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
static __global__ void tile_row_starts(const uint32_t *__restrict__ soff, uint32_t n_rows, uint32_t span,
                                       uint32_t n_tiles, uint32_t *__restrict__ tile_row)
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
    float *o_prob; uint32_t *o_lpos; double *o_aux; uint2 *o_rec; uint4 *o_records;
    uint32_t *o_trow;          // tile-order row -> original row
    uint32_t *fallback; uint32_t *cursors;  // [0] fallback rows, [1] sum D, [2] sum items, [3] record granules, [4..6] max record bytes / D / x doubles
};

// One CTA lays out one tile.
static __global__ void __launch_bounds__(kThreads, 4) build_tiles(BuildArgs a)   // 64 registers: four CTAs per SM (42 KB of shared memory each)
{
    using Sort = cub::BlockRadixSort<uint32_t, kThreads, 4, uint32_t>;
    using Scan = cub::BlockScan<uint32_t, kThreads>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ uint32_t s_txp[kTile];      // slot -> transcript; later reused as sorted keys
    __shared__ uint32_t s_src[kTile];      // slot -> source alignment index in the CSR
    __shared__ uint32_t s_lpos[kTile];
    __shared__ uint32_t s_seg[kTile + 1];  // segment -> first sorted rank; later start | first 32-slot item
    __shared__ uint32_t s_seg2[kTile];     // segment -> x offset of its remainder item | n32 | stray flag
    __shared__ uint16_t s_rlen[kTile], s_rslot[kTile], s_rnew[kTile];
    __shared__ uint32_t s_heads[kWarps * 4];
    __shared__ uint32_t s_used[kWarps], s_nrow[kWarps], s_info[kWarps];
    __shared__ uint32_t s_misc[4];
    __shared__ __align__(16) uint32_t s_rec[kRecMax / 4];

    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    const uint32_t nrows = min(r1 - r0, (uint32_t)kTile);  // every row has >= 1 alignment and span <= kTile

    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) { s_txp[i] = kNoTxp; s_lpos[i] = 0; s_src[i] = kNoTxp; }
    for (uint32_t i = tid; i < (uint32_t)kRecMax / 4; i += kThreads) s_rec[i] = 0;
    if (tid < kWarps * 4) s_heads[tid] = 0;
    for (uint32_t i = tid; i < nrows; i += kThreads) {
        const uint32_t r = a.srow[r0 + i];
        s_rlen[i] = (uint16_t)(a.row_ptr[r + 1] - a.row_ptr[r]);
    }
    __syncthreads();

    // first-fit packing of rows into warp-chunks (rows never straddle a chunk): rows in order, one warp, lane c keeps
    // the fill state of chunk c and a ballot finds the first chunk that fits
    __shared__ uint16_t s_pend[12];
    if (tid < 32u) {
        const unsigned full = 0xffffffffu;
        uint32_t used = 0, cnt = 0, span = 0;   // of chunk `tid` (lanes 0 .. kWarps-1)
        // Rows shorter than a lane (4 slots) must not start and end inside one lane: two row heads in a lane
        // send the whole chunk down the general segmented-sum path of the sweep.  The order of the rows inside a
        // tile is free, so a short row waits until a chunk's fill position lets it cross a lane boundary
        // ((used & 3) + len >= 4); what is still waiting at the end is placed wherever it fits.
        auto place = [&](uint32_t i, int pick) {
            if ((int)tid == pick) {
                const uint32_t len = s_rlen[i];
                s_rslot[i] = (uint16_t)(tid * kChunk + used);
                s_rnew[i] = (uint16_t)cnt;  // order inside the chunk
                const uint32_t lanes = ((used + len - 1) >> 2) - (used >> 2);  // lanes the row's tail must travel
                span = max(span, lanes);
                used += len; cnt += 1;
            }
        };
        auto first_fit = [&](uint32_t len, bool clean) -> int {
            const bool fits = tid < (uint32_t)kWarps && used + len <= (uint32_t)kChunkCap && (!clean || (used & 3u) + len >= 4u);
            const unsigned m = __ballot_sync(full, fits);
            return m ? __ffs((int)m) - 1 : -1;
        };
        constexpr uint32_t kPend = 12;
        uint32_t np = 0;
        for (uint32_t i = 0; i < nrows; ++i) {
            const uint32_t len = s_rlen[i];
            if (OAR_SHORT_ROW_DEFER && len < 4u && np < kPend) {
                const int pc = first_fit(len, true);
                if (pc >= 0) place(i, pc); else { if (tid == 0) s_pend[np] = (uint16_t)i; ++np; __syncwarp(); }
                continue;
            }
            const int pick = first_fit(len, false);
            if (pick < 0) { if (tid == 0) s_rslot[i] = 0xFFFF; continue; }
            place(i, pick);
            for (uint32_t k = 0; k < np;) {       // waiting short rows: does one fit cleanly now?
                const uint32_t pi = s_pend[k];
                const int pc = first_fit(s_rlen[pi], true);
                if (pc < 0) { ++k; continue; }
                place(pi, pc);
                __syncwarp();
                if (tid == 0) for (uint32_t m = k + 1; m < np; ++m) s_pend[m - 1] = s_pend[m];
                --np;
                __syncwarp();
            }
        }
        for (uint32_t k = 0; k < np; ++k) {
            const uint32_t i = s_pend[k];
            int pick = first_fit(s_rlen[i], true);
            if (pick < 0) pick = first_fit(s_rlen[i], false);
            if (pick < 0) { if (tid == 0) s_rslot[i] = 0xFFFF; } else place(i, pick);
        }
        if (tid < (uint32_t)kWarps) {
            s_used[tid] = used; s_nrow[tid] = cnt;
            uint32_t steps = 0;
            while ((1u << steps) <= span) ++steps;  // Hillis-Steele steps 1,2,..,2^(steps-1) cover `span` lanes
            s_info[tid] = steps;
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
            s_rec[kRecRow / 4 + tid] = r0 + chunk_base[tid];
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
    // lane descriptors: thread t <-> lane (t & 31) of chunk (t >> 5)
    {
        const uint32_t lane = tid & 31u;
        const unsigned full = 0xffffffffu;
        const uint32_t hb = (s_heads[tid >> 3] >> ((tid & 7u) * 4u)) & 0xFu;
        if (__popc(hb) >= 2) atomicOr(&s_info[tid >> 5], kInfoMulti);  // general segmented-sum path needed
        const unsigned lanes_h = __ballot_sync(full, hb != 0u);       // lane 0 always has a head
        const int P = 31 - __clz(lanes_h & (full >> (31u - lane)));
        const unsigned later = lanes_h & ~(full >> (31u - lane));
        const uint32_t E = later ? (uint32_t)(__ffs(later) - 1) : lane;
        const uint32_t desc = hb | ((lane - (uint32_t)P) << 4) | (E << 11);   // the chunk's rare / mid bits are added at the end
        reinterpret_cast<uint16_t *>(s_rec)[kRecDesc / 2 + tid] = (uint16_t)desc;
    }
    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) {
        const uint32_t src = s_src[i];
        a.o_prob[(size_t)tile * kTile + i] = src != kNoTxp ? a.prob[src] : 0.f;   // padding: prob 0 (aux 1)
        if (a.aux) a.o_aux[(size_t)tile * kTile + i] = src != kNoTxp ? a.aux[src] : 1.0;
    }

    // sort slots by transcript
    uint32_t keys[4], vals[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // feed the (stable) sort in (chunk, slot-in-lane k, lane) order: the alignments one STS of the
        // sweep scatters for one transcript then get consecutive ranks, i.e. consecutive smem banks
        const uint32_t j = tid * 4 + i;
        const uint32_t slot = (j & ~127u) | ((j & 31u) << 2) | ((j >> 5) & 3u);
        keys[i] = s_txp[slot]; vals[i] = slot;
    }
    // the ids of a tile span a narrow range (rows are ordered by their smallest id): sort the keys relative to the
    // tile's smallest id, on as many bits as the range needs (2-3 radix passes instead of 8); padding sorts last
    {
        uint32_t lo = kNoTxp, hi = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (keys[i] != kNoTxp) { lo = min(lo, keys[i]); hi = max(hi, keys[i]); }
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if (tid == 0) { s_misc[1] = kNoTxp; s_misc[3] = 0; }
        __syncthreads();
        if ((tid & 31u) == 0u && lo != kNoTxp) { atomicMin(&s_misc[1], lo); atomicMax(&s_misc[3], hi); }
        __syncthreads();
    }
    const uint32_t key_lo = s_misc[1], key_pad = s_misc[1] == kNoTxp ? 0u : s_misc[3] - s_misc[1] + 1u;   // relative key of the padding
#pragma unroll
    for (int i = 0; i < 4; ++i) keys[i] = keys[i] != kNoTxp ? keys[i] - key_lo : key_pad;
    Sort(tmp.sort).Sort(keys, vals, 0, 32 - __clz((int)key_pad | 1));
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) keys[i] = keys[i] != key_pad ? keys[i] + key_lo : kNoTxp;
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
    // items of the aggregated segments: packed per-class counts n32 | n16 << 10 | n8 << 20 (each total <= 256)
    uint32_t pk[4], pbase[4], cntd[4], startd[4], PT = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        pk[i] = 0; cntd[i] = 0; startd[i] = 0;
        if (d < D) {
            startd[i] = s_seg[d];
            cntd[i] = s_seg[d + 1] - startd[i];
            if (cntd[i] >= (uint32_t)kAggMin) {
                const uint32_t rem = cntd[i] % (uint32_t)kItemMax;
                pk[i] = cntd[i] / (uint32_t)kItemMax + (rem > (uint32_t)kItemMax / 2u ? 1u : 0u) +
                        ((rem > (uint32_t)kItemMax / 4u && rem <= (uint32_t)kItemMax / 2u) ? (1u << 10) : 0u) +
                        ((rem >= 1u && rem <= (uint32_t)kItemMax / 4u) ? (1u << 20) : 0u);
            }
        }
    }
    Scan(tmp.scan).ExclusiveSum(pk, pbase, PT);
    __syncthreads();
    const uint32_t N32 = PT & 0x3FFu, N16 = (PT >> 10) & 0x3FFu, N8 = PT >> 20;
    const uint32_t U = N32 + N16 + N8;                          // items of the tile
    constexpr uint32_t kS0 = kItemMax + 2, kS1 = kItemMax / 2 + 2, kS2 = kItemMax / 4 + 2, kI = kItemMax;   // strides of the three classes
    const uint32_t XD = kS0 * N32 + kS1 * N16 + kS2 * N8;       // x doubles of the tile; the trash slot sits right behind
    const uint32_t D4 = (D + 3u) & ~3u, U4 = (U + 3u) & ~3u;
    const uint32_t rec_bytes = kRecTable + 4u * D4 + 4u * U4;
    if (tid == 0) {
        atomicAdd(a.cursors + 1, D);
        atomicAdd(a.cursors + 2, U);
        s_misc[2] = atomicAdd(a.cursors + 3, ((rec_bytes + kRecAlign - 1u) / kRecAlign) * (kRecAlign / 16u));   // records start on kRecAlign-byte boundaries
        atomicMax(a.cursors + 4, rec_bytes);
        atomicMax(a.cursors + 5, D);
        atomicMax(a.cursors + 6, XD);
        s_rec[kRecDU / 4 + 0] = D;
        s_rec[kRecDU / 4 + 1] = kRecTable + 4u * U4;
        s_rec[kRecDU / 4 + 2] = U;
        s_rec[kRecDU / 4 + 3] = XD * 8u;
    }
    // table and item descriptors; per segment: where its alignments go
    uint32_t sega[4], segb[4];   // start | first 32-slot item << 11 ;  x offset (doubles) of the remainder item | n32 << 12 | stray << 31
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        sega[i] = segb[i] = 0;
        if (d < D) {
            s_rec[kRecTable / 4 + U4 + d] = s_txp[startd[i]];
            uint32_t *items = s_rec + kRecTable / 4;
            const uint32_t i32 = pbase[i] & 0x3FFu, i16 = (pbase[i] >> 10) & 0x3FFu, i8 = pbase[i] >> 20;
            const uint32_t n32 = pk[i] & 0x3FFu, rem = cntd[i] % kI;
            if (pk[i] == 0u) { segb[i] = 1u << 31; sega[i] = startd[i]; continue; }
            // item = 4 * table index | x offset (doubles) << 12 | valid slots << 24
            for (uint32_t v = 0; v < n32; ++v) items[i32 + v] = (d << 2) | ((kS0 * (i32 + v)) << 12) | (min(kI, cntd[i] - kI * v) << 24);
            uint32_t rem_x = 0;
            if (rem > kI / 4u && rem <= kI / 2u) { rem_x = kS0 * N32 + kS1 * i16; items[N32 + i16] = (d << 2) | (rem_x << 12) | (rem << 24); }
            else if (rem >= 1u && rem <= kI / 4u) { rem_x = kS0 * N32 + kS1 * N16 + kS2 * i8; items[N32 + N16 + i8] = (d << 2) | (rem_x << 12) | (rem << 24); }
            sega[i] = startd[i] | (i32 << 11);
            segb[i] = rem_x | (n32 << 12);
        }
    }
    __syncthreads();  // everyone has read s_seg[d+1]; s_misc[2] visible
    for (uint32_t u = U + tid; u < U4; u += kThreads) s_rec[kRecTable / 4 + u] = 0u;   // no item
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        if (d < D) { s_seg[d] = sega[i]; s_seg2[d] = segb[i]; }
    }
    __syncthreads();
#if OAR_SCATTER_GREEDY
    // Which of its transcript's x positions an alignment gets is free.  The sweep scatters the x values of slot k
    // of the 32 lanes of a warp with one STS.64; the 16 lanes of a half-warp go through in one wavefront iff their
    // positions fall into 16 different 8-byte banks.  A transcript's 16-slot items offer every bank residue once
    // each, its remainder item a few of them; a serial greedy pass over the 64 half-warp groups of the tile hands
    // every alignment a residue that is still unused in its group (starting the search at its lane, which spreads
    // the demand evenly), if its transcript still has one to offer.  (With ranks in sort order, as before, a
    // half-warp store took 3.1 wavefronts on C3 -- the same as positions drawn at random.)
    uint32_t *s_state = s_txp;          // per transcript: residues on offer (bits 0-15) | free remainder slots (16-31); s_txp is dead
    uint16_t *s_ci = s_rnew;            // per transcript: compact index of its full-item counters | partial << 15; dead since the rows were placed
    uint32_t *s_best = reinterpret_cast<uint32_t *>(&tmp);   // per transcript: the residues it has the MOST slots of left (the sort / scan scratch is dead)
    static_assert(sizeof(tmp) >= 4 * kTile, "s_best is carved out of the sort scratch");
    __shared__ __align__(16) uint8_t s_fulluse[kTile / kItemMax][16];                 // [compact transcript][residue] -> full items used
    if (tid == 0) s_misc[3] = 0;
    for (uint32_t i = tid; i < (uint32_t)(kTile / kItemMax) * 16u; i += kThreads) (&s_fulluse[0][0])[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s_src[vals[i]] = keys[i] != kNoTxp ? seg[i] - 1u : kNoTxp;   // slot -> segment
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = tid * 4 + i;
        if (d < D) {
            uint32_t avail = 0, remfree = 0, ci = 0, best = 0;
            if ((segb[i] >> 31) == 0u) {
                // q full 16-slot items, then `rem` valid slots in the remainder item -- a smaller-class item at
                // x offset rem_x, or (rem > 8) one more 16-slot item right behind the full ones
                const uint32_t q = cntd[i] / kI, rem = cntd[i] % kI, partial = rem > kI / 2u ? 1u : 0u;
                const uint32_t br = partial ? (2u * ((pbase[i] & 0x3FFu) + q)) & 15u : segb[i] & 15u;
                remfree = (1u << rem) - 1u;
                avail = ((remfree << br) | (remfree << br >> 16)) & 0xFFFFu;   // bank residues of the remainder's slots
                best = avail;   // residues the transcript has the most slots of: those of the remainder item, if it has one
                if (q) { avail = 0xFFFFu; ci = atomicAdd(&s_misc[3], 1u); }
                if (best == 0u) best = avail;
                ci |= partial << 15;
            }
            s_state[d] = avail | (remfree << 16); s_ci[d] = (uint16_t)ci; s_best[d] = best;
        }
    }
    __syncthreads();
    // One warp walks the 32 scatter instructions of the tile (chunk c, slot-in-lane k); its lanes are the lanes of
    // the sweep.  Per round every lane that still needs a position proposes the first residue at or after its lane
    // number that its transcript offers and its half-warp has not used; the lowest lane of every (half, residue)
    // and (transcript, residue) group goes ahead and claims it.  Deterministic: no atomics decide anything.
    if (tid < 32u) {
        const unsigned full = 0xffffffffu;
        const uint32_t lane = tid, half = lane >> 4, l16 = lane & 15u;
        for (uint32_t ck = 0; ck < (uint32_t)kWarps * 4u; ++ck) {
            const uint32_t slot = (ck >> 2) * kChunk + 4u * lane + (ck & 3u);
            const uint32_t d = s_src[slot];
            bool todo = d != kNoTxp && (s_state[d] & 0xFFFFu) != 0u;   // stray transcripts have no x positions
            uint32_t i32 = 0, q = 0, br = 0, ci = 0;
            if (todo) {
                const uint32_t pb = s_seg2[d], pa = s_seg[d], cw = s_ci[d];
                i32 = pa >> 11; q = ((pb >> 12) & 0x3FFu) - (cw >> 15); ci = cw & 0x7FFFu;
                br = (cw >> 15) ? (2u * (i32 + q)) & 15u : pb & 15u;
            }
            const bool nopos = !todo;   // padding or stray alignment: stores to a trash slot, chosen below
            uint32_t G = 0;   // residues taken in this lane's half-warp
            // Scarce first (model, tools/layout_model.py: 89 -> 80 scatter wavefronts per tile): lanes whose transcript
            // offers at most kScarce residues (remainder items of 4 or 8 slots) choose in a first phase, the others
            // after them.  (Ranking the lanes with __reduce_min_sync over the match groups did the same in one phase,
            // but a reduction whose mask differs between the lanes is serialised group by group: build_tiles 17 -> 45 ms.)
            constexpr uint32_t kScarce = 8;
#pragma unroll 1
            for (uint32_t phase = OAR_GREEDY_SCARCE ? 0u : 1u; phase < 2u; ++phase)
            for (;;) {
                uint32_t rho = 0, st = 0;
                bool prop = false;                         // does this lane propose in this round?
                if (todo) {
                    st = s_state[d];
                    prop = phase == 1u || (uint32_t)__popc(st & 0xFFFFu) <= kScarce;
                }
                if (!__any_sync(full, prop)) break;
                if (prop) {
                    const uint32_t av = st & 0xFFFFu;      // never 0 here: the transcript still owes this lane a position
                    uint32_t cand = av & ~G;
                    if (cand == 0u) cand = av;             // no unused residue on offer: accept a bank conflict
#if OAR_GREEDY_SUPPLY
                    // Of the candidates prefer the residues the transcript has MOST slots of left (unused full items plus
                    // the remainder item's slot): that keeps a transcript's supply level across the residues, so the later
                    // instructions of the tile still find every bank on offer.  s_best[d] holds that set; a lane that finds
                    // it exhausted recomputes it (all lanes of the transcript write the same word).  Real C3 tiles: 99 -> 75
                    // scatter wavefronts per tile, 64 is the floor; the exact "largest supply first" rule gives the same.
                    uint32_t best = s_best[d] & av;
                    if (best == 0u) {
                        const uint32_t remfree = st >> 16, remrot = ((remfree << br) | (remfree << br >> 16)) & 0xFFFFu;
                        uint32_t top = 0;
#pragma unroll 1
                        for (uint32_t r = 0; r < 16u; ++r) {   // rare (once per ~16 positions of a transcript): kept small
                            const uint32_t sup = q - (q ? (uint32_t)s_fulluse[ci][r] : 0u) + ((remrot >> r) & 1u);
                            if (sup > top) { top = sup; best = 0u; }
                            if (sup == top) best |= 1u << r;
                        }
                        best &= av;   // (top >= 1: av is not empty)
                        atomicExch(&s_best[d], best);   // (every lane of the transcript writes the same word)
                    }
                    if (cand & best) cand &= best;
#endif
                    const uint32_t rot = ((cand >> l16) | (cand << (16u - l16))) & 0xFFFFu;
                    rho = (l16 + (uint32_t)__ffs((int)rot) - 1u) & 15u;
                }
                const uint32_t idle = 0x80000000u | lane;
                const unsigned m1 = __match_any_sync(full, prop ? ((d << 4) | rho) : idle);
                const unsigned m2 = __match_any_sync(full, prop ? ((half << 4) | rho) : idle);
                // one winner per (transcript, residue) and per (half-warp, residue): the lowest lane of each group; the
                // lowest proposing lane of the warp wins both of its groups, so every round places someone
                const bool go = prop && (uint32_t)(__ffs((int)m1) - 1) == lane && (uint32_t)(__ffs((int)m2) - 1) == lane;
                uint32_t took = 0;
                if (go) {
                    uint32_t m = q ? s_fulluse[ci][rho] : 0u;
                    const uint32_t orem = (rho - br) & 15u;
                    uint32_t clear = 0;
                    if (m < q) {                               // item m of the transcript's full items
                        const uint32_t o = (rho - 2u * (i32 + m)) & 15u;   // item base = 18 * (i32 + m) doubles
                        s_src[slot] = d | ((kI * m + o) << 16);
                        s_fulluse[ci][rho] = (uint8_t)(++m);
                    } else {                                   // the remainder item (its slot of this residue is free)
                        s_src[slot] = d | ((kI * q + orem) << 16);
                        clear = 1u << (16u + orem);
                        st &= ~clear;
                    }
                    if (m >= q && !((st >> (16u + orem)) & 1u)) clear |= 1u << rho;   // residue no longer on offer
                    if (clear) atomicAnd(&s_state[d], ~clear);   // lanes of other residues may update the same word
#if OAR_GREEDY_SUPPLY
                    atomicAnd(&s_best[d], ~(1u << rho));
#endif
                    todo = false;
                    took = 1u << (rho + 16u * half);
                }
                __syncwarp();
                G |= (__reduce_or_sync(full, took) >> (16u * half)) & 0xFFFFu;
            }
            // the lanes without a position all store to ONE trash slot per half-warp, in a bank the others left free
            // (a store to the same address from several lanes is a single wavefront)
            if (nopos) {
                const uint32_t freeb = ~G & 0xFFFFu, r = freeb ? (uint32_t)__ffs((int)freeb) - 1u : 0u;
                s_lpos[slot] = (r - XD) & 15u;   // trash slot index: (XD + index) & 15 == r
            }
        }
    }
    __syncthreads();
    // per slot: table index and position, as shared-memory byte offsets
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t slot = tid * 4 + i;
        const uint32_t dj = s_src[slot];
        if (dj != kNoTxp) {
            const uint32_t d = dj & 0xFFFFu;
            const uint32_t pa = s_seg[d], pb = s_seg2[d];
            uint32_t pos = XD + s_lpos[slot];   // the trash slot picked for this lane's half-warp
            if ((pb >> 31) == 0u) {
                const uint32_t j = dj >> 16, n32 = (pb >> 12) & 0x3FFu;
                pos = j < kI * n32 ? kS0 * ((pa >> 11) + j / kI) + j % kI : (pb & 0xFFFu) + (j - kI * n32);
            } else atomicOr(&s_info[slot / kChunk], kInfoStray);
            s_lpos[slot] = table_off(d) | ((pos * 8u) << 16);
        } else {
            s_lpos[slot] = 0u | (((XD + s_lpos[slot]) * 8u) << 16);
        }
    }
#else
    // per sorted element: table index and position, as shared-memory byte offsets
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = tid * 4 + i;
        if (keys[i] != kNoTxp) {
            const uint32_t d = seg[i] - 1;
            const uint32_t pa = s_seg[d], pb = s_seg2[d];
            uint32_t pos = XD + ((vals[i] >> 2) & 15u);   // trash slot of this lane
            if ((pb >> 31) == 0u) {
                const uint32_t rr = r - (pa & 0x7FFu), n32 = (pb >> 12) & 0x3FFu;
                pos = rr < kI * n32 ? kS0 * ((pa >> 11) + rr / kI) + rr % kI : (pb & 0xFFFu) + (rr - kI * n32);
            } else atomicOr(&s_info[vals[i] / kChunk], kInfoStray);
            s_lpos[vals[i]] = table_off(d) | ((pos * 8u) << 16);
        } else {
            s_lpos[vals[i]] = 0u | (((XD + ((vals[i] >> 2) & 15u)) * 8u) << 16);
        }
    }
#endif
    __syncthreads();
    for (uint32_t i = tid; i < (uint32_t)kTile; i += kThreads) a.o_lpos[(size_t)tile * kTile + i] = s_lpos[i];
    if (tid < kWarps) s_rec[kRecInfo / 4 + tid] = s_info[tid];
    {
        const uint32_t info = s_info[tid >> 5];
        uint16_t *dp = reinterpret_cast<uint16_t *>(s_rec) + kRecDesc / 2 + tid;
        *dp = (uint16_t)(*dp | ((info & (kInfoMulti | kInfoStray | 4u)) ? kDescRare : 0u) | ((info & 7u) >= 3u ? kDescMid : 0u));
    }
    __syncthreads();
    const uint32_t rec_off = s_misc[2];
    uint4 *dst = a.o_records + rec_off;
    const uint4 *srcv = reinterpret_cast<const uint4 *>(s_rec);
    for (uint32_t i = tid; i < rec_bytes / 16u; i += kThreads) dst[i] = srcv[i];
    if (tid == 0) a.o_rec[tile] = make_uint2(rec_off, rec_bytes);
}

// bootstrap weights from read order into tile order
static __global__ void permute_weights(const uint32_t *__restrict__ w, const uint32_t *__restrict__ trow, uint32_t n,
                                       uint32_t *__restrict__ wperm)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) wperm[k] = w[trow[k]];
}

// Bootstrap weights per LANE of every tile: the weight of the row that ends in the lane (fast path of the sweep: at
// most one row head per lane, so that row is number (lanes with a head before this one) - 1 of its chunk).  The sweep
// gets them with the tile's TMA copies and reads one u16 per thread instead of deriving the row index (ballot, popc,
// chunk base) and loading the weight from global memory inside the E-step.  Chunks on the general path (two heads in
// a lane) read wperm as before.  Weights above 65535 do not fit: flag[0] is set and the caller reports an error.
static __global__ void __launch_bounds__(kThreads) lane_weights(const uint2 *__restrict__ rec, const uint4 *__restrict__ records,
                                                                const uint32_t *__restrict__ wperm, uint32_t n_tiles,
                                                                uint16_t *__restrict__ wlane, uint32_t *__restrict__ flag)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    bool over = false;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned char *r = reinterpret_cast<const unsigned char *>(records + rec[tile].x);
        const uint32_t desc = reinterpret_cast<const uint16_t *>(r + kRecDesc)[tid];
        const uint32_t row_base = reinterpret_cast<const uint32_t *>(r + kRecRow)[warp];
        const unsigned lanes_h = __ballot_sync(0xffffffffu, (desc & 15u) != 0u);
        const uint32_t before = __popc(lanes_h & ((1u << lane) - 1u));
        uint32_t w = before ? wperm[row_base + before - 1u] : 0u;
        if (w > 0xFFFFu) { over = true; w = 0xFFFFu; }
        wlane[(size_t)tile * kThreads + tid] = (uint16_t)w;
    }
    if (over) atomicOr(flag, 1u);
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

// 1.0 or 0.0 from a condition, as a bit pattern: lets "add if" be a single DFMA (ptxas turns
// predicated f64 adds into DADD + two FSELs)
__device__ __forceinline__ double mask01(bool c) { return __hiloint2double(c ? 0x3FF00000 : 0, 0); }

// --- mbarrier / TMA bulk copy (sm_90+; SASS: SYNCS, UBLKCP) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// --- explicit shared-space accesses ------------------------------------------
// The sweep addresses shared memory with 32-bit shared-space addresses: with generic pointers the compiler
// re-derives the shared window base (S2UR SR_CgaCtaId + ULEA) every iteration and cannot fold the window
// offsets into the LDS/STS address operands.
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t r; asm volatile("{.reg .u16 h; ld.shared.u16 h, [%1]; cvt.u32.u16 %0, h;}" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{ uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{ uint2 r; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a)); return r; }
__device__ __forceinline__ float4 lds_v4f(uint32_t a)
{ float4 r; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// prev[] entry from its split words (see table_off): two LDS.32, the second at a constant offset
__device__ __forceinline__ double lds_prev(uint32_t a)
{
#if OAR_PREV_SPLIT
    uint32_t hi, lo;
    asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+128];" : "=r"(hi), "=r"(lo) : "r"(a));
    return __hiloint2double((int)hi, (int)lo);
#else
    double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a)); return r;
#endif
}
__device__ __forceinline__ void sts_prev(uint32_t a, double v)
{
#if OAR_PREV_SPLIT
    sts_u32(a, (uint32_t)__double2hiint(v));
    sts_u32(a + kPrevLo, (uint32_t)__double2loint(v));
#else
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
#endif
}
__device__ __forceinline__ void sts_f64_if(uint32_t a, double v, bool on)
{ asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f64 [%0], %1;}" ::"r"(a), "d"(v), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ double lds_f64(uint32_t a) { double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a)); return r; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// 0.0 unless `on`
__device__ __forceinline__ double lds_f64_if(uint32_t a, bool on)
{
    double r;
    asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; mov.f64 %0, 0d0000000000000000; @p ld.shared.f64 %0, [%1];}"
                 : "=d"(r) : "r"(a), "r"((uint32_t)on));
    return r;
}
// {0.0, 0.0} unless `on`
__device__ __forceinline__ double2 lds_v2f64_if(uint32_t a, bool on)
{
    double2 r;
    asm volatile("{.reg .pred p; setp.ne.u32 p, %3, 0; mov.f64 %0, 0d0000000000000000; mov.f64 %1, 0d0000000000000000;"
                 " @p ld.shared.v2.f64 {%0,%1}, [%2];}" : "=d"(r.x), "=d"(r.y) : "r"(a), "r"((uint32_t)on));
    return r;
}
// warp vote on bits of a word that is the same in every lane (LOP3 with predicate output + VOTE)
__device__ __forceinline__ bool any_bits(uint32_t v, uint32_t mask)
{
    uint32_t r;
    asm volatile("{.reg .pred p, q; .reg .b32 t; and.b32 t, %1, %2; setp.ne.u32 p, t, 0; vote.sync.any.pred q, p, 0xffffffff;"
                 " selp.u32 %0, 1, 0, q;}" : "=r"(r) : "r"(v), "r"(mask));
    return r != 0u;
}

// ---- phase 1 of a tile: E-step in registers + M-step scatter into the transcript-sorted x array -------------
//   bulk, rec: shared-space addresses of the tile's prob | lpos block and of its record; sp_a: prev[] of the tile's
//   transcripts; xs_a: the x array.  Returns the thread's item and the record's DU word for phase 2.
//   The thread's four slots (prob, lpos) come in registers: p4 / lp4 = slots warp * kChunk + lane * 4 ...
#ifndef OAR_COMMON_PATH
#define OAR_COMMON_PATH 1       // sweep: one vote sends chunks without a rare feature down a copy of phase 1 that has none of them
#endif
// COMMON: the chunk has no lane with two row heads, no row spanning more than 8 lanes and no stray alignments (90 % of
// the chunks on C3): those three tests are compile-time false and cost neither votes nor branches.
template <bool HAS_AUX, bool HAS_WTS, bool COMMON = false>
__device__ __forceinline__ void tile_phase1_core(const View &v, uint32_t tile, const float4 p4, const uint4 lp4, uint32_t desc, uint32_t rec, uint32_t sp_a,
                                                 uint32_t xs_a, uint32_t w_in, uint32_t tid, uint32_t lane, uint32_t warp,
                                                 double *__restrict__ curr, const uint32_t *__restrict__ wperm)
{
    const unsigned full = 0xffffffffu;
    const uint32_t info = COMMON ? 0u : lds_u32(rec + kRecInfo + 4u * warp);

    double w0 = lds_prev(sp_a + (lp4.x & 0xFFFFu)) * (double)p4.x;
    double w1 = lds_prev(sp_a + (lp4.y & 0xFFFFu)) * (double)p4.y;
    double w2 = lds_prev(sp_a + (lp4.z & 0xFFFFu)) * (double)p4.z;
    double w3 = lds_prev(sp_a + (lp4.w & 0xFFFFu)) * (double)p4.w;
    if (HAS_AUX) {
        const double *ax = v.aux + (size_t)tile * kTile + 4u * tid;
        const double2 q0 = *reinterpret_cast<const double2 *>(ax);
        const double2 q1 = *reinterpret_cast<const double2 *>(ax + 2);
        w0 *= q0.x; w1 *= q0.y; w2 *= q1.x; w3 *= q1.y;
    }

    // the descriptor's fields are tested in place: hb = desc & 15, dist >= 2^i <=> desc & (0x1F0 & ~((2^i - 1) << 4)), and the
    // shuffle takes its source lane modulo 32, so E = desc >> 11 needs no mask
    const uint32_t hb = desc & 15u, E = desc >> 11;
    auto dist_ge = [&](uint32_t d) -> bool { return (desc & (0x1F0u & ~((d - 1u) << 4))) != 0u; };   // d a power of two
    // bootstrap: w_in = the resampling weight of the row that ends in this lane (fast path; staged with the tile,
    // see lane_weights)
    // chunk_info is the same word for the whole warp; votes make that visible to the compiler
    const bool multi = !COMMON && any_bits(info, kInfoMulti);
    const bool strays = !COMMON && any_bits(info, kInfoStray);
    const bool long_rows = !COMMON && any_bits(info, 4u);   // scan steps > 3: rows spanning more than 8 lanes
    const bool mid_rows = !OAR_SCAN_COND || any_bits(desc, kDescMid);   // scan steps > 2
    double x0, x1, x2, x3;
    if (!multi) {
        // fast path: no lane holds more than one row head.  Slot i lies before that head (it
        // closes the row entering the lane) iff hb >> (i+1) != 0; slot 3 never does.
        //   a = slots before the head, z = slots from the head on (all four if there is none)
        const bool c0 = (desc & 0xEu) != 0u, c1 = (desc & 0xCu) != 0u, c2 = (desc & 0x8u) != 0u;
        double a = w0 * mask01(c0);
        a = fma(w1, mask01(c1), a);
        a = fma(w2, mask01(c2), a);
        double z = fma(w0, mask01(!c0), w3);
        z = fma(w1, mask01(!c1), z);
        z = fma(w2, mask01(!c2), z);
        // segmented inclusive scan over lanes: what each lane adds to the row open at its end
        double incl = z;
        incl = fma(__shfl_up_sync(full, incl, 1), mask01(dist_ge(1u)), incl);
        incl = fma(__shfl_up_sync(full, incl, 2), mask01(dist_ge(2u)), incl);
        if (mid_rows) incl = fma(__shfl_up_sync(full, incl, 4), mask01(dist_ge(4u)), incl);   // rows spanning more than 4 lanes (a third of the chunks on C3)
        if (long_rows) {   // rows spanning more than 8 lanes (rare)
            incl = fma(__shfl_up_sync(full, incl, 8), mask01(dist_ge(8u)), incl);
            incl = fma(__shfl_up_sync(full, incl, 16), mask01(dist_ge(16u)), incl);
        }
        const double carry = __shfl_up_sync(full, incl, 1);       // sum of the row entering this lane
        const double t_in = carry + a;                            // its total, if it ends here
        // reads whose denominator is <= 1e-30 contribute nothing (em.rs:115)
        double inv_in = t_in > OAR_EM_DENOM_THRESH ? fast_rcp(t_in) : 0.0;
        if (HAS_WTS) inv_in *= (double)w_in;   // fold the row's resampling weight into its inverse denominator
        // the row leaving this lane ends in lane E (E == lane: only padding follows, w == 0)
        const double inv_out = __shfl_sync(full, inv_in, E);
        x0 = w0 * (c0 ? inv_in : inv_out);
        x1 = w1 * (c1 ? inv_in : inv_out);
        x2 = w2 * (c2 ? inv_in : inv_out);
        x3 = w3 * inv_out;
    } else {
        // general path: rows may start and end inside one lane
        const uint32_t nsteps = info & 7u;
        const double s0 = w0;
        const double s1 = (hb & 2u) ? w1 : s0 + w1;
        const double s2 = (hb & 4u) ? w2 : s1 + w2;
        const double s3 = (hb & 8u) ? w3 : s2 + w3;
        double incl = s3;                                         // segmented inclusive scan of lane tails
#pragma unroll
        for (uint32_t i = 0; i < 5; ++i) {
            if (i >= nsteps) break;
            const uint32_t d = 1u << i;
            incl = fma(__shfl_up_sync(full, incl, d), mask01(dist_ge(d)), incl);
        }
        double carry = __shfl_up_sync(full, incl, 1);
        if (lane == 0) carry = 0.0;
        const double sA = (hb & 1u) ? 0.0 : (hb & 2u) ? s0 : (hb & 4u) ? s1 : (hb & 8u) ? s2 : s3;
        const double t_in = carry + sA;
        const double t_out = __shfl_sync(full, t_in, E);
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

    if (HAS_WTS && multi) {
        // general path: row index of a slot inside the chunk = (number of heads at or before it) - 1
        uint32_t incl_h = __popc(hb);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(full, incl_h, d);
            if ((int)lane >= d) incl_h += t;
        }
        const uint32_t row_base = lds_u32(rec + kRecRow + 4u * warp);
        const uint32_t r0 = row_base + incl_h - __popc(hb) + (hb & 1u) - 1u;
        const uint32_t r1 = r0 + ((hb >> 1) & 1u), r2 = r1 + ((hb >> 2) & 1u), r3 = r2 + ((hb >> 3) & 1u);
        x0 *= (double)wperm[r0];
        x1 *= (double)wperm[r1];
        x2 *= (double)wperm[r2];
        x3 *= (double)wperm[r3];
    }

    // ---- M-step scatter into the transcript-sorted smem order -----------------------------
    // (padding and non-aggregated alignments carry the address of a trash slot in a bank their half-warp leaves free, so the
    // stores are unconditional.  What lands in a trash slot is always +0.0: padding has w = 0, and a chunk with stray
    // alignments takes the predicated stores below)
    const uint32_t q0 = lp4.x >> 16, q1 = lp4.y >> 16, q2 = lp4.z >> 16, q3 = lp4.w >> 16;
    if (!strays) {
        sts_f64(xs_a + q0, x0);
        sts_f64(xs_a + q1, x1);
        sts_f64(xs_a + q2, x2);
        sts_f64(xs_a + q3, x3);
    } else {
        // transcripts with fewer than kAggMin alignments in this tile: straight to global
        const uint4 du = lds_v4(rec + kRecDU);   // D, table offset, items, trash offset
        const uint32_t trash = du.w, table_a = rec + du.y;
        sts_f64_if(xs_a + q0, x0, q0 < trash);
        sts_f64_if(xs_a + q1, x1, q1 < trash);
        sts_f64_if(xs_a + q2, x2, q2 < trash);
        sts_f64_if(xs_a + q3, x3, q3 < trash);
        if (q0 >= trash && x0 != 0.0) atomicAdd(curr + lds_u32(table_a + 4u * table_index(lp4.x & 0xFFFFu)), x0);
        if (q1 >= trash && x1 != 0.0) atomicAdd(curr + lds_u32(table_a + 4u * table_index(lp4.y & 0xFFFFu)), x1);
        if (q2 >= trash && x2 != 0.0) atomicAdd(curr + lds_u32(table_a + 4u * table_index(lp4.z & 0xFFFFu)), x2);
        if (q3 >= trash && x3 != 0.0) atomicAdd(curr + lds_u32(table_a + 4u * table_index(lp4.w & 0xFFFFu)), x3);
    }
}

// phase 1 with the tile's prob | lpos block staged in shared memory (`bulk`); also returns the thread's item and the
// record's DU word for phase 2 (read now: the record's stage is refilled before phase 2).
template <bool HAS_AUX, bool HAS_WTS, bool COMMON_OK>
__device__ __forceinline__ void tile_phase1(const View &v, uint32_t tile, uint32_t bulk, uint32_t rec, uint32_t sp_a, uint32_t xs_a,
                                            uint32_t tid, uint32_t lane, uint32_t warp, double *__restrict__ curr,
                                            const uint32_t *__restrict__ wperm, uint32_t &item, uint32_t &item_txp, uint32_t &U)
{
    const float4 p4 = lds_v4f(bulk + 16u * tid);
    const uint4 lp4 = lds_v4(bulk + 4u * kTile + 16u * tid);
    const uint32_t desc = lds_u16(rec + kRecDesc + 2u * tid);
    const uint4 du = lds_v4(rec + kRecDU);   // D, table offset, items, trash offset
    U = du.z;
    item = 0; item_txp = 0;
    if (tid < U) {
        item = lds_u32(rec + kRecTable + 4u * tid);
        item_txp = lds_u32(rec + du.y + (item & 0xFFCu));
    }
    uint32_t w_in = 0;
    if (HAS_WTS) w_in = lds_u16(rec - 2u * kThreads + 2u * tid);   // the lane weights sit right in front of the record
#if OAR_COMMON_PATH
    // the rare bit is the same in every lane of the warp: one vote decides between the two copies of phase 1
    if (COMMON_OK && !any_bits(desc, kDescRare))
        tile_phase1_core<HAS_AUX, HAS_WTS, true>(v, tile, p4, lp4, desc, rec, sp_a, xs_a, w_in, tid, lane, warp, curr, wperm);
    else
#endif
    tile_phase1_core<HAS_AUX, HAS_WTS, false>(v, tile, p4, lp4, desc, rec, sp_a, xs_a, w_in, tid, lane, warp, curr, wperm);
}

// ---- phase 2 of a tile: one thread sums one item (<= 16 consecutive x slots of one transcript), one RED -------
__device__ __forceinline__ void tile_phase2(uint32_t xs_a, uint32_t item, uint32_t item_txp, uint32_t U, uint32_t warp,
                                            double *__restrict__ curr)
{
    const unsigned full = 0xffffffffu;
    if (__any_sync(full, warp * 32u < U)) {
        const uint32_t slots = item >> 24, npair = slots >> 1;   // 0 slots: no item
        constexpr uint32_t kP = kItemMax / 2;
        const uint32_t b = xs_a + ((item >> 9) & 0x7FF8u);       // x offset in doubles at bit 12 -> bytes
        double a0 = lds_f64_if(b + 8u * (slots - 1u), (slots & 1u) != 0u), a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (uint32_t k = 0; k < kP / 4u; k += 2) {
            const double2 u = lds_v2f64_if(b + 16u * k, k < npair), w = lds_v2f64_if(b + 16u * (k + 1u), k + 1u < npair);
            a0 += u.x; a1 += u.y; a2 += w.x; a3 += w.y;
        }
        if (__any_sync(full, npair > kP / 4u)) {        // items of the two larger classes (they come first in the item order)
#pragma unroll
            for (uint32_t k = kP / 4u; k < kP / 2u; k += 2) {
                const double2 u = lds_v2f64_if(b + 16u * k, k < npair), w = lds_v2f64_if(b + 16u * (k + 1u), k + 1u < npair);
                a0 += u.x; a1 += u.y; a2 += w.x; a3 += w.y;
            }
            if (__any_sync(full, npair > kP / 2u)) {    // items of the largest class
#pragma unroll
                for (uint32_t k = kP / 2u; k < kP; k += 2) {
                    const double2 u = lds_v2f64_if(b + 16u * k, k < npair), w = lds_v2f64_if(b + 16u * (k + 1u), k + 1u < npair);
                    a0 += u.x; a1 += u.y; a2 += w.x; a3 += w.y;
                }
            }
        }
        const double acc = (a0 + a2) + (a1 + a3);
        if (acc != 0.0) atomicAdd(curr + item_txp, acc);   // (no item: nothing was loaded, acc == 0)
    }

}

// m_step (em.rs:87-133), persistent and TMA-fed.
// LIST: the sweep walks v.tile_list[0 .. *v.n_active) instead of all tiles (batched per-cell EM: tiles whose cells
// have all converged are dropped from the list between graph launches, oar_cells.cu).
// FUSED: the head of the sweep carries the convergence bookkeeping of the previous iteration (kern::em_update_slice).  A
// separate instantiation, because the mere presence of that code changes the register allocation of the tile loop
// (+3.5 us per sweep on C3): stores whose sweep is long enough not to care about one launch keep the lean kernel.
template <bool HAS_AUX, bool HAS_WTS, bool LIST = false, bool FUSED = false>
__global__ void __launch_bounds__(kThreads, (sweep_ctas(HAS_AUX, HAS_WTS, LIST) * 8) / kWarps) em_sweep_tiled(View v, Geometry g, const double *__restrict__ prev,
                                                              double *__restrict__ curr,
                                                              const uint32_t *__restrict__ wperm,
                                                              const OarEmState *__restrict__ st, int check_done)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // [xs][stage 0][stage 1][s_prev][mbarriers]; a stage = prob | lpos | record

    if (check_done && st->done) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tiles = LIST ? *v.n_active : v.n_tiles, stride = gridDim.x;
    const uint32_t tile0 = blockIdx.x;   // position in the walk; phys() is the tile it stands for
    auto phys = [&](uint32_t i) -> uint32_t { return LIST ? v.tile_list[i] : i; };
    // rows that are not tiled: every CTA takes its share of the list (8 lanes per row, straight from the CSR)
    auto fallback_rows = [&]() {
        if (v.n_fb)
            kern::rowgroup_rows<HAS_AUX, HAS_WTS>(v.csr_row_ptr, v.csr_txp, v.csr_prob, v.csr_aux, v.csr_wts, v.fb_rows, prev, curr,
                                                  (uint64_t)blockIdx.x * (kThreads >> 3) + (threadIdx.x >> 3),
                                                  (uint64_t)gridDim.x * (kThreads >> 3), v.n_fb);
    };
    if (tile0 >= n_tiles) { fallback_rows(); return; }
    const uint32_t kStageBytes = g.stage_bytes;
    constexpr uint32_t kRecOff = 8u * kTile + (HAS_WTS ? 2u * kThreads : 0u);   // record's offset in a stage
    uint32_t sm0;   // shared-space address of the window, computed once (a plain cvta is rematerialised every iteration)
    asm volatile("{.reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t;}" : "=r"(sm0) : "l"(smem));
    const uint32_t stage0 = sm0 + g.stage_off, sp_a = sm0 + g.prev_off, bar0 = sm0 + g.bar_off;
    // xs sits at the start of the window; its address comes as a parameter (see Geometry::xs_base)
#ifndef OAR_XS_PARAM
#define OAR_XS_PARAM 1          // 0: the x-array base as a cvta on the extern array (A/B timing)
#endif
    // Which instantiation gets the parameter base and the common-path copy of phase 1 was settled by A/B timing on C3
    // (profiles/experiments/r2_kernel_experiments.md): ptxas moves 1-3 % either way with each of them.
#ifndef OAR_XS_PARAM_WTS
#define OAR_XS_PARAM_WTS 1      // lean weighted: parameter base + common path since the descriptor carries the chunk's rare / mid bits
#endif                          // (179.6 us; common path alone 181.4, parameter base alone 186.7, neither 187.7-188.4)
#ifndef OAR_COMMON_PATH_WTS
#define OAR_COMMON_PATH_WTS 1
#endif
#ifndef OAR_XS_PARAM_FUSED_WTS
#define OAR_XS_PARAM_FUSED_WTS 1
#endif
#ifndef OAR_COMMON_PATH_FUSED_WTS
#define OAR_COMMON_PATH_FUSED_WTS 1
#endif
    constexpr bool kXsParam = OAR_XS_PARAM && (!HAS_WTS || (FUSED ? OAR_XS_PARAM_FUSED_WTS : OAR_XS_PARAM_WTS));
    constexpr bool kCommonOk = !HAS_WTS || (FUSED ? OAR_COMMON_PATH_FUSED_WTS : OAR_COMMON_PATH_WTS);
    uint32_t xs_a;
    if (kXsParam) { xs_a = g.xs_base; if (sm0 != xs_a) __trap(); }
    else xs_a = smem_u32(smem);

    // Work between the two CTA barriers of a tile is spread over the warps: the first warps sum the items
    // (phase 2), lane 0 of the last-but-one warp issues the TMA copies, the last warp gathers prev[].
    const bool is_tma = tid == 32u * (kWarps - 2);
    auto issue = [&](uint32_t tile, uint32_t s, uint2 r) {   // the TMA thread only
        const uint32_t bar = bar0 + 8u * s, dst = stage0 + s * kStageBytes;
        mbar_expect_tx(bar, 8u * kTile + r.y + (HAS_WTS ? 2u * kThreads : 0u));
        bulk_g2s(dst, v.prob + (size_t)tile * kTile, 4u * kTile, bar);
        bulk_g2s(dst + 4u * kTile, v.lpos + (size_t)tile * kTile, 4u * kTile, bar);
        bulk_g2s(dst + kRecOff, v.records + r.x, r.y, bar);
        if (HAS_WTS) bulk_g2s(dst + 8u * kTile, v.wlane + (size_t)tile * kThreads, 2u * kThreads, bar);
    };
    // prev[] of a tile's transcripts into s_prev, by one warp (two gathers in flight per lane)
    auto gather_prev = [&](uint32_t rec_a) {
        const uint2 du = lds_v2(rec_a + kRecDU);   // D, table offset
        const uint32_t Dn = du.x, table_a = rec_a + du.y;
        for (uint32_t d = lane; d < Dn; d += 64u) {
            const uint32_t d2 = d + 32u;
            const double p0 = prev[lds_u32(table_a + 4u * d)];
            double p1 = 0.0;
            if (d2 < Dn) p1 = prev[lds_u32(table_a + 4u * d2)];
            sts_prev(sp_a + table_off(d), p0);
            if (d2 < Dn) sts_prev(sp_a + table_off(d2), p1);
        }
    };

    uint2 r_pending = make_uint2(0, 0);   // record locator of the tile two ahead (TMA thread)
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t t_pending = 0;               // and that tile itself
    if (is_tma) {
        const uint32_t p0 = phys(tile0);
        issue(p0, 0, v.rec[p0]);
        if (tile0 + stride < n_tiles) { const uint32_t p1 = phys(tile0 + stride); issue(p1, 1, v.rec[p1]); }
        if (tile0 + 2 * stride < n_tiles) { t_pending = phys(tile0 + 2 * stride); r_pending = v.rec[t_pending]; }
    }
    // the first copies are in flight: judge the previous iteration on this CTA's slice of the count vectors
    if (FUSED && !LIST) {
        // the three count buffers rotate: the one that is neither this sweep's prev nor its curr was the previous sweep's prev.
        // (Everything the bookkeeping needs comes from the EM state in device memory: growing the kernel's parameter list
        // for it changed how ptxas treats the shared window base in the tile loop, 185 -> 191 us.)
        OarEmState *stw = const_cast<OarEmState *>(st);
        double *old = stw->bufs[0] != prev && stw->bufs[0] != curr ? stw->bufs[0] : (stw->bufs[1] != prev && stw->bufs[1] != curr ? stw->bufs[1] : stw->bufs[2]);
        kern::em_update_slice(old, prev, stw->n_txps, stw, reinterpret_cast<double *>(smem));   // scratch: the x array, not in use yet
    }
    // Only the last warp waits on the stage mbarriers and gathers prev[]; the CTA barrier that follows hands the
    // TMA-written stage on to the other warps (mbarrier completion observed by one thread + bar.sync is cumulative).
    if (warp == kWarps - 1) {
        mbar_wait(bar0, 0);
        gather_prev(stage0 + kRecOff);
    }

    uint32_t tile = tile0;
    for (uint32_t it = 0;; ++it) {
        const uint32_t s = it & 1u;
        const uint32_t stg = stage0 + s * kStageBytes;
        const uint32_t rec = stg + kRecOff;
        __syncthreads();   // stage s and s_prev of this tile are in place; phase 2 of the previous tile has left xs

        uint32_t item, item_txp, U;
        tile_phase1<HAS_AUX, HAS_WTS, kCommonOk>(v, HAS_AUX ? phys(tile) : 0u, stg, rec, sp_a, xs_a, tid, lane, warp, curr, wperm, item, item_txp, U);
        __syncthreads();   // xs complete; stage s and s_prev are free again

        // ---- refill stage s two tiles ahead; the last warp fetches prev[] of the next tile ----
        const uint32_t next = tile + stride;
        const bool has_next = next < n_tiles;
        if (warp >= kWarps - 2) {   // the two service warps
            if (is_tma && next + stride < n_tiles) {
                issue(LIST ? t_pending : next + stride, s, r_pending);
                if (next + 2 * stride < n_tiles) { t_pending = phys(next + 2 * stride); r_pending = v.rec[t_pending]; }
            }
            if (warp == kWarps - 1 && has_next) {
                mbar_wait(bar0 + 8u * (s ^ 1u), ((it + 1u) >> 1) & 1u);
                gather_prev(stage0 + (s ^ 1u) * kStageBytes + kRecOff);
            }
        }
        __syncwarp();

        tile_phase2(xs_a, item, item_txp, U, warp, curr);

        if (!has_next) break;
        tile = next;
    }
    fallback_rows();
}


}  // namespace tiled
}  // namespace oar
