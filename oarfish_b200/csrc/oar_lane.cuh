// oar_lane.cuh -- the row-per-lane layout and its fused E+M sweep (default kernel).
//
// Third generation of the locality-tiled design (first: oar_tiled.cuh).  Shared idea: rows (reads)
// are ordered by their smallest transcript id, so a few hundred consecutive rows touch a few dozen
// transcripts; prev[] of those is gathered once into shared memory, the M-step is aggregated in
// shared memory per transcript and flushed with few f64 REDs instead of one per alignment.
//
// What is different here:
//   * ONE READ PER LANE.  A build tile (~2048 alignments, ~256 rows) sorts its rows by length and
//     cuts them into GROUPS of <= 32 rows / <= kGroupCap alignments.  Lane l of the warp that owns a
//     group walks row l serially: the per-read denominator (em.rs:98-112) is a plain sum in
//     registers -- no shuffles, no segment masks.  The j-th alignments of a group's rows are stored
//     contiguously (the lanes with len > j are a prefix), so every access is a coalesced,
//     conflict-free LDS.64 and the HBM stream has no padding.
//   * WARPS ARE INDEPENDENT.  A group is a complete unit of work: its own transcript table, its own
//     x array, its own record.  Every warp runs a private two-stage TMA pipeline (cp.async.bulk +
//     mbarrier) over groups gw, gw + W, gw + 2W, ... and synchronises only with itself
//     (__syncwarp): there is no CTA barrier in the sweep, so a slow group never stalls seven others
//     (the CTA-tile version of this kernel spent 43 % of its warp time in BAR.SYNC).
//   * CHEAP M-STEP.  x = w/denom is scattered to the group's transcript-sorted x array; a transcript
//     with cnt alignments owns ceil(cnt/32) ITEMS of <= 32 consecutive slots (34-double stride), one
//     lane sums one item with LDS.128 and issues ONE RED: no cross-lane scan, no padding units.
//
// Per alignment the HBM stream is one 8-byte pair {prob f32, lpos u32}; lpos = (byte offset of the
// transcript's prev[] copy in the warp's smem) | (byte offset of the alignment's x slot) << 16.
#pragma once
#include <cub/cub.cuh>

#include "oar_common.cuh"
#include "oar_kernels.cuh"
#include "oar_tiled.cuh"

namespace oar {
namespace lane {

#ifndef OAR_LANE_WARPS
#define OAR_LANE_WARPS 2
#endif
#ifndef OAR_LANE_IPT
#define OAR_LANE_IPT 9
#endif
#ifndef OAR_LANE_MIN_CTAS
#define OAR_LANE_MIN_CTAS 10
#endif
#ifndef OAR_LANE_REG_ROWS
#define OAR_LANE_REG_ROWS 16
#endif
#ifndef OAR_LANE_GROUP_CAP
#define OAR_LANE_GROUP_CAP 384
#endif
constexpr int kWarps = OAR_LANE_WARPS;          // warps per sweep CTA (each one independent)
constexpr int kThreads = kWarps * 32;
constexpr int kMinCtas = OAR_LANE_MIN_CTAS;     // register budget: CTAs per SM the sweep is compiled for
constexpr int kBuildThreads = 256;              // layout construction CTA
constexpr int kIpt = OAR_LANE_IPT;              // slots per build thread
constexpr int kSlotsMax = kBuildThreads * kIpt; // alignment slots a build tile can hold (2304)
constexpr int kRowCap = tiled::kChunkCap;       // longer rows are swept from the CSR (fallback list)
constexpr int kGroupCap = OAR_LANE_GROUP_CAP;   // alignments per group
constexpr int kGroupsMax = kSlotsMax / 32 + kSlotsMax / (kGroupCap - kRowCap) + 2;
constexpr int kSpanMax = kSlotsMax - kRowCap - kGroupsMax;   // tile alignments <= span + kRowCap - 1, + one pad slot per group
constexpr int kSpanDefault = kSpanMax < 2048 ? kSpanMax : 2048;
constexpr int kRegRows = OAR_LANE_REG_ROWS;     // alignments per read the E-step keeps in registers (8, 12 or 16)
constexpr int kItem = 16;                       // x slots one lane sums in the M-step
constexpr int kItemStride = 18;                 // doubles between consecutive items of a transcript (LDS.128 bank skew)
constexpr uint32_t kNoTxp = 0xFFFFFFFFu;
static_assert(kRegRows == 12 || kRegRows == 16, "register rows: 12 or 16");
static_assert(kGroupCap >= 2 * kRowCap && kGroupCap <= 1024, "group capacity");
static_assert(kSlotsMax <= 4096 && kGroupsMax <= 128, "sort keys pack (group, transcript index) into 7 + 12 bits");

// Per-group record (16-byte granules, one TMA bulk copy):
//   header u32[4]: D | items << 16,  nnz | rows << 16,  row_base,  first pair
//   rlen   u8[32]      row lengths in lane order (0 = no row), non-increasing
//   table  u32[D]      distinct transcript ids (padded to 4)
//   items  u32[items]  x offset in doubles (12 bits) | (slots - 1) << 12 | table index << 17 (padded to 4)
constexpr uint32_t kRecRlen = 16, kRecTable = 48;
__host__ __device__ inline uint32_t r4(uint32_t x) { return (x + 3u) & ~3u; }
__host__ __device__ inline uint32_t rec_bytes_of(uint32_t D, uint32_t items) { return kRecTable + 4u * r4(D) + 4u * r4(items); }
template <typename T> __device__ __forceinline__ T lds(const unsigned char *smem, uint32_t off)
{ return *reinterpret_cast<const T *>(smem + off); }

// Shared memory of ONE WARP of the sweep, sized for the store at hand (maxima over all groups):
// two stages (pairs | record), the x array, prev[] of the group's transcripts, two mbarriers.
struct Geometry {
    uint32_t pair_bytes, stage_bytes, xs_off, prev_off, bar_off, warp_bytes;
};
inline Geometry make_geometry(uint32_t max_nnz, uint32_t max_rec_bytes, uint32_t max_d, uint32_t max_xs_doubles)
{
    Geometry g;
    g.pair_bytes = ((max_nnz + 1u) & ~1u) * 8u;
    g.stage_bytes = g.pair_bytes + ((max_rec_bytes + 15u) & ~15u);
    g.xs_off = 2u * g.stage_bytes;
    g.prev_off = g.xs_off + 8u * ((max_xs_doubles + 3u) & ~1u);
    g.bar_off = g.prev_off + 8u * ((max_d + 1u) & ~1u);
    g.warp_bytes = (g.bar_off + 16u + 127u) & ~127u;
    return g;
}

struct View {
    uint32_t n_groups;
    const uint2 *pairs;        // {prob bits, lpos}; group k starts at groups[k].z (even)
    const double *aux;         // same indexing, or null
    const uint4 *groups;       // {record offset (16 B granules), record bytes, first pair, nnz}
    const uint4 *records;
    const uint32_t *fb_rows; uint32_t n_fb;
    const uint32_t *csr_row_ptr; const uint32_t *csr_txp; const float *csr_prob; const double *csr_aux;
    const uint32_t *csr_wts;
};

// ---------------------------------------------------------------------------
// layout construction: one CTA per build tile
// ---------------------------------------------------------------------------

struct BuildArgs {
    const uint32_t *row_ptr; const uint32_t *txp; const float *prob; const double *aux;
    const uint32_t *srow;      // sorted position -> original row
    const uint32_t *soff;      // exclusive prefix of the sorted rows' lengths (n_tiled + 1)
    const uint32_t *tile_row;  // n_tiles + 1
    uint2 *o_pairs; double *o_aux; uint4 *o_groups; uint4 *o_records;
    uint32_t *o_trow;          // length-sorted tile order row -> original row
    // [0] record granules, [1] groups, [2] pairs, [3] sum D, [4] sum items, [5] max record bytes, [6] max D,
    // [7] max x doubles, [8] max nnz
    uint32_t *cursors;
};

using BSortK = cub::BlockRadixSort<uint32_t, kBuildThreads, kIpt>;
using BSortKV = cub::BlockRadixSort<uint32_t, kBuildThreads, kIpt, uint32_t>;
using BScan = cub::BlockScan<uint32_t, kBuildThreads>;
union BuildTemp { typename BSortK::TempStorage sortk; typename BSortKV::TempStorage sortkv; typename BScan::TempStorage scan; };

struct BuildSmem {
    uint32_t a1[kSlotsMax + 4];   // row-contiguous transcript ids -> tile-level transcript table
    uint32_t a2[kSlotsMax + 4];   // row-contiguous source indices -> segment starts -> segment info
    uint32_t b1[kSlotsMax];       // slot -> transcript -> sorted keys -> lpos
    uint32_t b2[kSlotsMax];       // slot -> source alignment index in the CSR
    uint16_t dslot[kSlotsMax];    // slot -> index into the tile-level transcript table
    uint16_t rord[kSlotsMax];     // length-sorted position -> row of the tile
    uint16_t coff[kSlotsMax];     // row of the tile -> offset of its alignments in a1/a2
    uint8_t plen[kSlotsMax + 32]; // length-sorted position -> row length
    uint8_t gslot[kSlotsMax];     // slot -> group
    uint16_t grow[kGroupsMax + 2];    // group -> first length-sorted row
    uint32_t goff[kGroupsMax + 2];    // group -> first slot (even)
    uint32_t gfs[kGroupsMax + 2];     // group -> first (group, transcript) segment
    uint32_t git0[kGroupsMax + 2];    // group -> items before it
    uint32_t gxs0[kGroupsMax + 2];    // group -> x doubles before it
    uint32_t grec[kGroupsMax + 2];    // group -> record offset (bytes) inside the tile's record block
    uint32_t misc[8];
    BuildTemp tmp;
};

static __global__ void __launch_bounds__(kBuildThreads) build_lane_tiles(BuildArgs a)
{
    extern __shared__ __align__(16) unsigned char bsm_raw[];
    BuildSmem &sm = *reinterpret_cast<BuildSmem *>(bsm_raw);
    const unsigned full = 0xffffffffu;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    const uint32_t nrows = r1 - r0;
    const uint32_t base_off = a.soff[r0];
    const uint32_t nnz = a.soff[r1] - base_off;
    if (nrows == 0) return;

    // A. rows sorted by length, longest first (ties keep the locality order)
    uint32_t rk[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t i = tid * kIpt + k;
        if (i < nrows) {
            const uint32_t o0 = a.soff[r0 + i] - base_off, len = a.soff[r0 + i + 1] - base_off - o0;
            sm.coff[i] = (uint16_t)o0;
            rk[k] = (((uint32_t)kRowCap - len) << 13) | i;
        } else {
            rk[k] = (((uint32_t)kRowCap + 1u) << 13) | i;
        }
        sm.b1[i] = kNoTxp; sm.b2[i] = kNoTxp;
    }
    BSortK(sm.tmp.sortk).Sort(rk, 0, 21);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t p = tid * kIpt + k;
        if (p < nrows) { sm.rord[p] = (uint16_t)(rk[k] & 0x1FFFu); sm.plen[p] = (uint8_t)((uint32_t)kRowCap - (rk[k] >> 13)); }
        else sm.plen[p] = 0;
    }
    if (tid < 32) sm.plen[kSlotsMax + tid] = 0;
    __syncthreads();

    // groups: up to 32 consecutive rows of the length order, at most kGroupCap alignments
    if (tid == 0) {
        uint32_t g = 0, rows_in = 0, slots_in = 0;
        sm.grow[0] = 0;
        for (uint32_t p = 0; p < nrows; ++p) {
            const uint32_t len = sm.plen[p];
            if (rows_in == 32u || slots_in + len > (uint32_t)kGroupCap) { ++g; sm.grow[g] = (uint16_t)p; rows_in = 0; slots_in = 0; }
            ++rows_in; slots_in += len;
        }
        sm.grow[g + 1] = (uint16_t)nrows;
        sm.misc[1] = g + 1;
    }

    // B. each row's alignments sorted by transcript (the order inside a row is free; insertion sort, rows are short)
    for (uint32_t i = tid; i < nrows; i += kBuildThreads) {
        const uint32_t r = a.srow[r0 + i];
        const uint32_t s = a.row_ptr[r], len = a.row_ptr[r + 1] - s, o = sm.coff[i];
        for (uint32_t j = 0; j < len; ++j) {
            const uint32_t t = a.txp[s + j];
            uint32_t k = j;
            while (k > 0 && sm.a1[o + k - 1] > t) { sm.a1[o + k] = sm.a1[o + k - 1]; sm.a2[o + k] = sm.a2[o + k - 1]; --k; }
            sm.a1[o + k] = t; sm.a2[o + k] = s + j;
        }
    }
    __syncthreads();
    const uint32_t G = sm.misc[1];

    // C. group slot ranges (even starts: TMA needs 16-byte alignment), then the (group, j, lane) slot order
    for (uint32_t g = warp; g < G; g += kBuildThreads / 32) {
        const uint32_t p = sm.grow[g] + lane;
        uint32_t sum = p < sm.grow[g + 1] ? sm.plen[p] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
        if (lane == 0) sm.goff[g] = sum;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t g = 0; g < G; ++g) { const uint32_t t = sm.goff[g]; sm.goff[g] = acc; acc += (t + 1u) & ~1u; }
        sm.goff[G] = acc;
    }
    __syncthreads();
    const uint32_t T = sm.goff[G];   // slots incl. the pad slot of odd groups
    for (uint32_t g = warp; g < G; g += kBuildThreads / 32) {
        const uint32_t p = sm.grow[g] + lane;
        const bool has = p < sm.grow[g + 1];
        const uint32_t len = has ? sm.plen[p] : 0u;
        const uint32_t i = has ? sm.rord[p] : 0u;
        const uint32_t o = sm.coff[i];
        if (has) a.o_trow[r0 + p] = a.srow[r0 + i];
        const uint32_t L = __shfl_sync(full, len, 0);
        uint32_t off = sm.goff[g] + lane;
        for (uint32_t j = 0; j < L; ++j) {
            const bool act = j < len;
            if (act) { sm.b1[off] = sm.a1[o + j]; sm.b2[off] = sm.a2[o + j]; sm.gslot[off] = (uint8_t)g; }
            off += __popc(__ballot_sync(full, act));
        }
    }
    __syncthreads();

    // D. tile-level transcript table: slots sorted by transcript
    uint32_t keys[kIpt], vals[kIpt], hf[kIpt], seg[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) { const uint32_t s = tid * kIpt + k; keys[k] = sm.b1[s]; vals[k] = s; }
    __syncthreads();
    BSortKV(sm.tmp.sortkv).Sort(keys, vals);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) sm.b1[tid * kIpt + k] = keys[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t r = tid * kIpt + k;
        hf[k] = (keys[k] != kNoTxp && (r == 0 || sm.b1[r - 1] != keys[k])) ? 1u : 0u;
    }
    uint32_t Dt = 0;
    BScan(sm.tmp.scan).InclusiveSum(hf, seg, Dt);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        if (hf[k]) sm.a1[seg[k] - 1] = keys[k];
        if (keys[k] != kNoTxp) sm.dslot[vals[k]] = (uint16_t)(seg[k] - 1u);
    }
    __syncthreads();

    // E. (group, transcript) segments: stable sort of the slots by group | table index keeps the (j, lane)
    //    order inside a segment, so the alignments one STS of the sweep scatters for a transcript get
    //    consecutive x slots, i.e. consecutive banks
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        keys[k] = sm.b2[s] != kNoTxp ? (((uint32_t)sm.gslot[s] << 12) | sm.dslot[s]) : kNoTxp;
        vals[k] = s;
    }
    __syncthreads();
    BSortKV(sm.tmp.sortkv).Sort(keys, vals, 0, 20);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) sm.b1[tid * kIpt + k] = keys[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t r = tid * kIpt + k;
        hf[k] = (keys[k] != kNoTxp && (r == 0 || sm.b1[r - 1] != keys[k])) ? 1u : 0u;
    }
    uint32_t S = 0;
    BScan(sm.tmp.scan).InclusiveSum(hf, seg, S);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) if (hf[k]) sm.a2[seg[k] - 1] = tid * kIpt + k;
    if (tid == 0) sm.a2[S] = nnz;
    __syncthreads();
    uint32_t nit[kIpt], xsz[kIpt], st0[kIpt], cnt0[kIpt], itx[kIpt], xsx[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        nit[k] = xsz[k] = st0[k] = cnt0[k] = 0;
        if (s < S) {
            st0[k] = sm.a2[s];
            cnt0[k] = sm.a2[s + 1] - st0[k];
            nit[k] = (cnt0[k] + (uint32_t)kItem - 1u) / (uint32_t)kItem;
            xsz[k] = (cnt0[k] + (uint32_t)(kItemStride - kItem) * (nit[k] - 1u) + 1u) & ~1u;
        }
    }
    uint32_t n_items_tile = 0, n_xs_tile = 0;
    BScan(sm.tmp.scan).ExclusiveSum(nit, itx, n_items_tile);
    __syncthreads();
    BScan(sm.tmp.scan).ExclusiveSum(xsz, xsx, n_xs_tile);
    __syncthreads();   // also: every thread has read a2[s], a2[s + 1]
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        if (s < S) {
            const uint32_t g = sm.b1[st0[k]] >> 12;
            const bool first = s == 0 || (sm.b1[st0[k] - 1u] >> 12) != g;
            if (first) { sm.gfs[g] = s; sm.git0[g] = itx[k]; sm.gxs0[g] = xsx[k]; }
        }
    }
    if (tid == 0) { sm.gfs[G] = S; sm.git0[G] = n_items_tile; sm.gxs0[G] = n_xs_tile; }
    __syncthreads();
    // per group: record size and maxima
    if (tid < G) {
        const uint32_t g = tid;
        const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g], ni = sm.git0[g + 1] - sm.git0[g], xs = sm.gxs0[g + 1] - sm.gxs0[g];
        sm.grec[g] = rec_bytes_of(Dg, ni);
        atomicMax(a.cursors + 5, sm.grec[g]); atomicMax(a.cursors + 6, Dg); atomicMax(a.cursors + 7, xs);
        atomicMax(a.cursors + 8, sm.goff[g + 1] - sm.goff[g]);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t g = 0; g < G; ++g) { const uint32_t t = sm.grec[g]; sm.grec[g] = acc; acc += t; }
        sm.grec[G] = acc;
        sm.misc[0] = atomicAdd(a.cursors + 0, acc / 16u);
        sm.misc[2] = atomicAdd(a.cursors + 1, G);
        sm.misc[3] = atomicAdd(a.cursors + 2, T);
        atomicAdd(a.cursors + 3, S); atomicAdd(a.cursors + 4, n_items_tile);
    }
    __syncthreads();
    const uint32_t rec_off = sm.misc[0], grp_base = sm.misc[2], pair_base = sm.misc[3];
    unsigned char *recs = reinterpret_cast<unsigned char *>(a.o_records + rec_off);
    if (tid < G) {
        const uint32_t g = tid;
        const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g], ni = sm.git0[g + 1] - sm.git0[g];
        const uint32_t rows_g = (uint32_t)sm.grow[g + 1] - sm.grow[g];
        uint32_t nnz_g = 0;
        for (uint32_t p = sm.grow[g]; p < sm.grow[g + 1]; ++p) nnz_g += sm.plen[p];
        const uint32_t bytes = sm.grec[g + 1] - sm.grec[g];
        uint32_t *hdr = reinterpret_cast<uint32_t *>(recs + sm.grec[g]);
        hdr[0] = Dg | (ni << 16); hdr[1] = nnz_g | (rows_g << 16); hdr[2] = r0 + sm.grow[g]; hdr[3] = pair_base + sm.goff[g];
        uint32_t *table = reinterpret_cast<uint32_t *>(recs + sm.grec[g] + kRecTable);
        for (uint32_t d = Dg; d < r4(Dg); ++d) table[d] = 0u;
        uint32_t *items = table + r4(Dg);
        for (uint32_t i = ni; i < r4(ni); ++i) items[i] = 0u;
        a.o_groups[grp_base + g] = make_uint4(rec_off + sm.grec[g] / 16u, bytes, pair_base + sm.goff[g], nnz_g);
    }
    for (uint32_t q = tid; q < 32u * G; q += kBuildThreads) {
        const uint32_t g = q >> 5, l = q & 31u, p = sm.grow[g] + l;
        recs[sm.grec[g] + kRecRlen + l] = p < sm.grow[g + 1] ? sm.plen[p] : (uint8_t)0;
    }
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        if (s < S) {
            const uint32_t key = sm.b1[st0[k]];
            const uint32_t g = key >> 12, d = key & 0xFFFu;
            const uint32_t dl = s - sm.gfs[g], xb = xsx[k] - sm.gxs0[g], ib = itx[k] - sm.git0[g];
            const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g];
            uint32_t *table = reinterpret_cast<uint32_t *>(recs + sm.grec[g] + kRecTable);
            uint32_t *items = table + r4(Dg);
            table[dl] = sm.a1[d];
            for (uint32_t v = 0; v < nit[k]; ++v) {
                const uint32_t slots = min((uint32_t)kItem, cnt0[k] - (uint32_t)kItem * v);
                items[ib + v] = (xb + (uint32_t)kItemStride * v) | ((slots - 1u) << 12) | (dl << 17);
            }
        }
    }
    __syncthreads();   // every thread has read b1[st0], a2[]; both are reused below
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        if (s < S) {
            const uint32_t g = sm.b1[st0[k]] >> 12;   // still the sorted keys: b1 is rewritten after the next barrier
            sm.a2[s] = st0[k] | ((xsx[k] - sm.gxs0[g]) << 12);
        }
    }
    __syncthreads();
    // F. per alignment: smem byte offsets of its transcript's prev[] copy and of its x slot
    uint32_t lp[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        lp[k] = 0;
        if (keys[k] != kNoTxp) {
            const uint32_t r = tid * kIpt + k, s = seg[k] - 1u, g = keys[k] >> 12;
            const uint32_t info = sm.a2[s];
            const uint32_t rr = r - (info & 0xFFFu);
            const uint32_t pos = (info >> 12) + (rr / (uint32_t)kItem) * (uint32_t)kItemStride + (rr % (uint32_t)kItem);
            lp[k] = ((s - sm.gfs[g]) * 8u) | ((pos * 8u) << 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) if (keys[k] != kNoTxp) sm.b1[vals[k]] = lp[k];
    __syncthreads();
    for (uint32_t s = tid; s < T; s += kBuildThreads) {
        const uint32_t src = sm.b2[s];
        a.o_pairs[(size_t)pair_base + s] = src != kNoTxp ? make_uint2(__float_as_uint(a.prob[src]), sm.b1[s]) : make_uint2(0u, 0u);
        if (a.aux) a.o_aux[(size_t)pair_base + s] = src != kNoTxp ? a.aux[src] : 1.0;
    }
}

// ---------------------------------------------------------------------------
// the sweep
// ---------------------------------------------------------------------------

// One group = up to 32 rows, one per lane, sorted by length (lane 0 longest, L = its length).  The
// j-th alignments of the rows are stored contiguously for the lanes with len > j (a prefix of the
// lanes), so lane l finds its j-th alignment at slot l + sum_{i<j} c_i, c_i = #lanes with len > i.
// J >= L alignments per read live in registers between the denominator pass and the scatter; the
// code is straight-line (inactive slots load nothing and contribute 0), so the J independent
// load -> gather -> multiply chains overlap and the denominator is a pairwise tree.
template <int J, bool HAS_AUX>
__device__ __forceinline__ void group_run(const unsigned char *stg, const double *gaux, const char *sp, char *xp,
                                          uint32_t lane, uint32_t len, double rw)
{
    const unsigned full = 0xffffffffu;
    uint32_t off[J];
    {
        uint32_t o = lane;
#pragma unroll
        for (int j = 0; j < J; ++j) { off[j] = o; o += __popc(__ballot_sync(full, (uint32_t)j < len)); }
    }
    uint2 pr[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        pr[j] = make_uint2(0u, 0u);                                  // prob 0, table entry 0: contributes nothing
        if ((uint32_t)j < len) pr[j] = *reinterpret_cast<const uint2 *>(stg + 8u * off[j]);
    }
    double w[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const double pv = *reinterpret_cast<const double *>(sp + (pr[j].y & 0xFFFFu));
        w[j] = pv * (double)__uint_as_float(pr[j].x);               // em.rs:107
        if (HAS_AUX) { if ((uint32_t)j < len) w[j] *= gaux[off[j]]; }  // em.rs:108-111
    }
    double t[J];
#pragma unroll
    for (int j = 0; j < J; ++j) t[j] = w[j];
#pragma unroll
    for (int n = J; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int j = 0; j < n / 2; ++j) t[j] = t[2 * j] + t[2 * j + 1];
        if (n & 1) t[n / 2] = t[n - 1];
    }
    const double denom = t[0];
    // reads whose denominator is <= 1e-30 contribute nothing (em.rs:115)
    const double inv = (denom > OAR_EM_DENOM_THRESH ? tiled::fast_rcp(denom) : 0.0) * rw;
#pragma unroll
    for (int j = 0; j < J; ++j)
        if ((uint32_t)j < len) *reinterpret_cast<double *>(xp + (pr[j].y >> 16)) = w[j] * inv;   // em.rs:119-130
}

// Groups with rows longer than the register budget: two predicated passes over shared memory.
template <bool HAS_AUX>
__device__ __noinline__ void group_long(const unsigned char *stg, const double *gaux, const char *sp, char *xp,
                                        uint32_t lane, uint32_t len, uint32_t L, double rw)
{
    const unsigned full = 0xffffffffu;
    double denom = 0.0;
    uint32_t off = lane;
    for (uint32_t j = 0; j < L; ++j) {
        const bool act = j < len;
        if (act) {
            const uint2 pr = *reinterpret_cast<const uint2 *>(stg + 8u * off);
            double ww = *reinterpret_cast<const double *>(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
            if (HAS_AUX) ww *= gaux[off];
            denom += ww;
        }
        off += __popc(__ballot_sync(full, act));
    }
    const double inv = (denom > OAR_EM_DENOM_THRESH ? tiled::fast_rcp(denom) : 0.0) * rw;
    off = lane;
    for (uint32_t j = 0; j < L; ++j) {
        const bool act = j < len;
        if (act) {
            const uint2 pr = *reinterpret_cast<const uint2 *>(stg + 8u * off);
            double ww = *reinterpret_cast<const double *>(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
            if (HAS_AUX) ww *= gaux[off];
            *reinterpret_cast<double *>(xp + (pr.y >> 16)) = ww * inv;
        }
        off += __popc(__ballot_sync(full, act));
    }
}

// m_step (em.rs:87-133): persistent, one independent TMA-fed pipeline per warp.
template <bool HAS_AUX, bool HAS_WTS>
__global__ void __launch_bounds__(kThreads, kMinCtas) em_sweep_lane(View v, Geometry g, const double *__restrict__ prev,
                                                                    double *__restrict__ curr,
                                                                    const uint32_t *__restrict__ wperm,
                                                                    const OarEmState *__restrict__ st, int check_done)
{
    extern __shared__ __align__(128) unsigned char smem_all[];
    if (check_done && st->done) return;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_groups = v.n_groups;
    const uint32_t gw = blockIdx.x * (uint32_t)kWarps + warp, nW = gridDim.x * (uint32_t)kWarps;

    if (gw < n_groups) {
        // this warp's shared memory: [stage 0][stage 1][xs][s_prev][mbarriers]; a stage = pairs | record
        unsigned char *smem = smem_all + warp * g.warp_bytes;
        char *xp = reinterpret_cast<char *>(smem + g.xs_off);
        double *s_prev = reinterpret_cast<double *>(smem + g.prev_off);
        const char *sp = reinterpret_cast<const char *>(s_prev);
        const uint32_t bar0 = tiled::smem_u32(smem + g.bar_off), smem0 = tiled::smem_u32(smem);

        auto issue = [&](uint32_t s, uint4 loc) {   // lane 0 only
            const uint32_t bar = bar0 + 8u * s, dst = smem0 + s * g.stage_bytes;
            const uint32_t dbytes = ((loc.w + 1u) & ~1u) * 8u;
            tiled::mbar_expect_tx(bar, dbytes + loc.y);
            if (dbytes) tiled::bulk_g2s(dst, v.pairs + loc.z, dbytes, bar);
            tiled::bulk_g2s(dst + g.pair_bytes, v.records + loc.x, loc.y, bar);
        };

        uint4 pending = make_uint4(0, 0, 0, 0);   // locator of the group two ahead (lane 0)
        if (lane == 0) {
            tiled::mbar_init(bar0, 1);
            tiled::mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            issue(0, v.groups[gw]);
            if (gw + nW < n_groups) issue(1, v.groups[gw + nW]);
            if (gw + 2 * nW < n_groups) pending = v.groups[gw + 2 * nW];
        }
        __syncwarp();
        tiled::mbar_wait(bar0, 0);
        {
            const uint32_t D = lds<uint32_t>(smem, g.pair_bytes) & 0xFFFFu;
            for (uint32_t d = lane; d < D; d += 32) s_prev[d] = prev[lds<uint32_t>(smem, g.pair_bytes + kRecTable + 4u * d)];
        }
        __syncwarp();

        uint32_t grp = gw;
        for (uint32_t it = 0;; ++it) {
            const uint32_t s = it & 1u;
            const unsigned char *stg = smem + s * g.stage_bytes;
            const uint32_t rec = s * g.stage_bytes + g.pair_bytes;
            const uint4 hdr = lds<uint4>(smem, rec);
            const uint32_t D = hdr.x & 0xFFFFu, NI = hdr.x >> 16;
            const uint32_t table = rec + kRecTable, items = table + 4u * r4(D);

            // ---- E-step, one read per lane; x = w/denom scattered into the transcript order ----
            {
                const uint32_t len = lds<uint8_t>(smem, rec + kRecRlen + lane);
                const uint32_t L = __shfl_sync(full, len, 0);
                const double *gaux = HAS_AUX ? v.aux + hdr.w : nullptr;
                double rw = 1.0;
                // bootstrap: the read's resampling weight scales its contribution (== visiting it that many times)
                if (HAS_WTS) rw = (double)wperm[hdr.z + lane];
                switch ((L + 1u) >> 1) {   // warp-uniform
                case 0: break;
                case 1: group_run<2, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 2: group_run<4, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 3: group_run<6, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 4: group_run<8, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 5: group_run<10, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 6: group_run<12, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
#if OAR_LANE_REG_ROWS >= 16
                case 7: group_run<14, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
                case 8: group_run<16, HAS_AUX>(stg, gaux, sp, xp, lane, len, rw); break;
#endif
                default: group_long<HAS_AUX>(stg, gaux, sp, xp, lane, len, L, rw); break;
                }
            }
            __syncwarp();   // xs complete; s_prev is free again

            // ---- prev[] of the next group: its record has landed long ago; the gather overlaps the M-step ----
            const uint32_t next = grp + nW;
            const bool has_next = next < n_groups;
            double pv = 0.0;
            uint32_t Dn = 0, table_n = 0;
            if (has_next) {
                tiled::mbar_wait(bar0 + 8u * (s ^ 1u), ((it + 1u) >> 1) & 1u);
                const uint32_t rec_n = (s ^ 1u) * g.stage_bytes + g.pair_bytes;
                Dn = lds<uint32_t>(smem, rec_n) & 0xFFFFu;
                table_n = rec_n + kRecTable;
                if (lane < Dn) pv = prev[lds<uint32_t>(smem, table_n + 4u * lane)];
            }

            // ---- M-step: one lane sums one item (<= 16 consecutive x slots of one transcript), one RED ----
            for (uint32_t ib = 0; ib < NI; ib += 32) {
                const uint32_t i = ib + lane;
                if (i < NI) {
                    const uint32_t desc = lds<uint32_t>(smem, items + 4u * i);
                    const uint32_t cnt = ((desc >> 12) & 31u) + 1u, npair = cnt >> 1;
                    const char *pb = xp + 8u * (desc & 0xFFFu);
                    const double2 *p = reinterpret_cast<const double2 *>(pb);
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    if (cnt & 1u) a2 = *reinterpret_cast<const double *>(pb + 8u * (cnt - 1u));
                    for (uint32_t k = 0; k < npair; k += 2) {
                        const double2 u = p[k];
                        a0 += u.x; a1 += u.y;
                        if (k + 1u < npair) { const double2 w = p[k + 1u]; a2 += w.x; a3 += w.y; }
                    }
                    const double acc = (a0 + a2) + (a1 + a3);
                    if (acc != 0.0) atomicAdd(curr + lds<uint32_t>(smem, table + 4u * (desc >> 17)), acc);
                }
            }
            __syncwarp();   // stage s (pairs, record) and xs are free again

            if (!has_next) break;
            if (lane == 0 && next + nW < n_groups) {
                issue(s, pending);
                if (next + 2 * nW < n_groups) pending = v.groups[next + 2 * nW];
            }
            if (lane < Dn) s_prev[lane] = pv;
            for (uint32_t d = lane + 32u; d < Dn; d += 32) s_prev[d] = prev[lds<uint32_t>(smem, table_n + 4u * d)];
            __syncwarp();
            grp = next;
        }
    }
    if (v.n_fb && blockIdx.x == gridDim.x - 1u)
        kern::rowgroup_rows<HAS_AUX, HAS_WTS>(v.csr_row_ptr, v.csr_txp, v.csr_prob, v.csr_aux, v.csr_wts, v.fb_rows, prev, curr,
                                              tid >> 3, kThreads >> 3, v.n_fb);
}

}  // namespace lane
}  // namespace oar
