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
#define OAR_LANE_WARPS 4
#endif
#ifndef OAR_LANE_IPT
#define OAR_LANE_IPT 9
#endif
#ifndef OAR_LANE_MIN_CTAS
#define OAR_LANE_MIN_CTAS 6
#endif
#ifndef OAR_LANE_REG_ROWS
#define OAR_LANE_REG_ROWS 12
#endif
#ifndef OAR_LANE_GROUP_CAP
#define OAR_LANE_GROUP_CAP 288
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
static_assert(kGroupCap >= 2 * kRowCap && kGroupCap <= 512, "group capacity (the record header packs per-class item counts into 8 bits)");
static_assert(kSlotsMax <= 4096 && kGroupsMax <= 128, "sort keys pack (group, transcript index) into 7 + 12 bits");

// Per-group blob (16-byte granules, ONE TMA bulk copy):
//   header u32[4]: D | items << 16,  n16 | n8 << 8 | n4 << 16 | rows << 24 (items per size class),  row_base,
//                  first pair (index into aux)
//   rlen   u8[32]      row lengths in lane order (0 = no row), non-increasing
//   table  u32[D]      distinct transcript ids (padded to 4)
//   items  u32[items]  (slots - 1) << 12 | table index << 17 (padded to 4), ordered by size class (16, 8, 4, 2
//                      slots); a class-c item sits c + 2 doubles behind its predecessor
//   pairs  uint2[nnz]  {prob bits, lpos} in (j, lane) order (padded to an even count)
constexpr uint32_t kRecRlen = 16, kRecTable = 48;
__host__ __device__ inline uint32_t r4(uint32_t x) { return (x + 3u) & ~3u; }
__host__ __device__ inline uint32_t rec_bytes_of(uint32_t D, uint32_t items) { return kRecTable + 4u * r4(D) + 4u * r4(items); }
template <typename T> __device__ __forceinline__ T lds(const unsigned char *smem, uint32_t off)
{ return *reinterpret_cast<const T *>(smem + off); }

// Shared memory of ONE WARP of the sweep, sized for the store at hand (maxima over all groups):
// two stages (one blob each), the x array, prev[] of the group's transcripts, two mbarriers.
struct Geometry {
    uint32_t stage_bytes, xs_off, prev_off, bar_off, warp_bytes;
    uint32_t trash;   // byte offset (inside the x array) of a slot nobody reads: lanes without a j-th alignment store there
};
inline Geometry make_geometry(uint32_t max_blob_bytes, uint32_t max_d, uint32_t max_xs_doubles)
{
    Geometry g;
    g.stage_bytes = (max_blob_bytes + 15u) & ~15u;
    g.xs_off = 2u * g.stage_bytes;
    g.trash = 8u * ((max_xs_doubles + 1u) & ~1u);
    g.prev_off = g.xs_off + g.trash + 16u;
    g.bar_off = g.prev_off + 8u * ((max_d + 1u) & ~1u);
    g.warp_bytes = (g.bar_off + 16u + 127u) & ~127u;
    return g;
}

struct View {
    uint32_t n_groups;
    const uint4 *blobs;        // all group blobs
    const uint2 *groups;       // per group: {blob offset (16 B granules), blob bytes}
    const double *aux;         // per pair (index = header's first pair + slot), or null
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
    double *o_aux; uint2 *o_groups; uint4 *o_blobs;
    uint32_t *o_trow;          // length-sorted tile order row -> original row
    // [0] blob granules, [1] groups, [2] pairs, [3] sum D, [4] sum items, [5] max blob bytes, [6] max D,
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
    uint32_t git0[kGroupsMax + 2];    // group -> 16-slot | 8-slot << 16 items before it
    uint32_t gxs0[kGroupsMax + 2];    // group -> 4-slot | 2-slot << 16 items before it
    uint32_t grec[kGroupsMax + 2];    // group -> blob offset (bytes) inside the tile's block of blobs
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
    // Items: a transcript with cnt alignments in the group gets cnt / 16 full items and one item for the
    // remainder, of the smallest size class (2, 4, 8 or 16 slots) that holds it.  The group's items are
    // ordered by class (16s first); class c items sit at a stride of c + 2 doubles, so the 8 lanes of an
    // LDS.128 phase of the M-step always hit 8 different 16-byte banks and the x array stays < 2.25 nnz.
    uint32_t st0[kIpt], cnt0[kIpt], pa[kIpt], pb[kIpt], xa[kIpt], xb[kIpt];   // packed counts: n16 | n8 << 16, n4 | n2 << 16
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        st0[k] = cnt0[k] = pa[k] = pb[k] = 0;
        if (s < S) {
            st0[k] = sm.a2[s];
            cnt0[k] = sm.a2[s + 1] - st0[k];
            const uint32_t rem = cnt0[k] & 15u;
            pa[k] = (cnt0[k] >> 4) + (rem > 8u ? 1u : 0u) + ((rem > 4u && rem <= 8u) ? 0x10000u : 0u);
            pb[k] = ((rem > 2u && rem <= 4u) ? 1u : 0u) + ((rem >= 1u && rem <= 2u) ? 0x10000u : 0u);
        }
    }
    uint32_t ta = 0, tb = 0;
    BScan(sm.tmp.scan).ExclusiveSum(pa, xa, ta);
    __syncthreads();
    BScan(sm.tmp.scan).ExclusiveSum(pb, xb, tb);
    __syncthreads();   // also: every thread has read a2[s], a2[s + 1]
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        if (s < S) {
            const uint32_t g = sm.b1[st0[k]] >> 12;
            const bool first = s == 0 || (sm.b1[st0[k] - 1u] >> 12) != g;
            if (first) { sm.gfs[g] = s; sm.git0[g] = xa[k]; sm.gxs0[g] = xb[k]; }
        }
    }
    if (tid == 0) { sm.gfs[G] = S; sm.git0[G] = ta; sm.gxs0[G] = tb; }
    __syncthreads();
    // per group: class counts (the packed fields subtract independently: the prefix sums are monotone per field)
    auto n16_of = [&](uint32_t g) { return (sm.git0[g + 1] - sm.git0[g]) & 0xFFFFu; };
    auto n8_of = [&](uint32_t g) { return (sm.git0[g + 1] - sm.git0[g]) >> 16; };
    auto n4_of = [&](uint32_t g) { return (sm.gxs0[g + 1] - sm.gxs0[g]) & 0xFFFFu; };
    auto n2_of = [&](uint32_t g) { return (sm.gxs0[g + 1] - sm.gxs0[g]) >> 16; };
    auto ni_of = [&](uint32_t g) { return n16_of(g) + n8_of(g) + n4_of(g) + n2_of(g); };
    if (tid < G) {
        const uint32_t g = tid;
        const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g], ni = ni_of(g);
        const uint32_t xs = 18u * n16_of(g) + 10u * n8_of(g) + 6u * n4_of(g) + 2u * n2_of(g);
        sm.grec[g] = rec_bytes_of(Dg, ni) + 8u * (sm.goff[g + 1] - sm.goff[g]);
        atomicMax(a.cursors + 5, sm.grec[g]); atomicMax(a.cursors + 6, Dg); atomicMax(a.cursors + 7, xs);
        atomicMax(a.cursors + 8, sm.goff[g + 1] - sm.goff[g]);
        atomicAdd(a.cursors + 4, ni);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t g = 0; g < G; ++g) { const uint32_t t = sm.grec[g]; sm.grec[g] = acc; acc += t; }
        sm.grec[G] = acc;
        sm.misc[0] = atomicAdd(a.cursors + 0, acc / 16u);
        sm.misc[2] = atomicAdd(a.cursors + 1, G);
        sm.misc[3] = atomicAdd(a.cursors + 2, T);
        atomicAdd(a.cursors + 3, S);
    }
    __syncthreads();
    const uint32_t rec_off = sm.misc[0], grp_base = sm.misc[2], pair_base = sm.misc[3];
    unsigned char *recs = reinterpret_cast<unsigned char *>(a.o_blobs + rec_off);
    if (tid < G) {
        const uint32_t g = tid;
        const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g], ni = ni_of(g);
        const uint32_t rows_g = (uint32_t)sm.grow[g + 1] - sm.grow[g];
        const uint32_t bytes = sm.grec[g + 1] - sm.grec[g];
        uint32_t *hdr = reinterpret_cast<uint32_t *>(recs + sm.grec[g]);
        hdr[0] = Dg | (ni << 16); hdr[1] = n16_of(g) | (n8_of(g) << 8) | (n4_of(g) << 16) | (rows_g << 24);
        hdr[2] = r0 + sm.grow[g]; hdr[3] = pair_base + sm.goff[g];
        uint32_t *table = reinterpret_cast<uint32_t *>(recs + sm.grec[g] + kRecTable);
        for (uint32_t d = Dg; d < r4(Dg); ++d) table[d] = 0u;
        uint32_t *items = table + r4(Dg);
        for (uint32_t i = ni; i < r4(ni); ++i) items[i] = 0u;
        a.o_groups[grp_base + g] = make_uint2(rec_off + sm.grec[g] / 16u, bytes);
    }
    for (uint32_t q = tid; q < 32u * G; q += kBuildThreads) {
        const uint32_t g = q >> 5, l = q & 31u, p = sm.grow[g] + l;
        recs[sm.grec[g] + kRecRlen + l] = p < sm.grow[g + 1] ? sm.plen[p] : (uint8_t)0;
    }
    uint32_t info_a[kIpt], info_b[kIpt];   // per segment: start | x base of its 16-slot items << 12;  x base of its remainder item | n16 << 12
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        info_a[k] = info_b[k] = 0;
        if (s < S) {
            const uint32_t key = sm.b1[st0[k]];
            const uint32_t g = key >> 12, d = key & 0xFFFu;
            const uint32_t dl = s - sm.gfs[g];
            const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g];
            const uint32_t N16 = n16_of(g), N8 = n8_of(g), N4 = n4_of(g);
            const uint32_t l16 = (xa[k] & 0xFFFFu) - (sm.git0[g] & 0xFFFFu), l8 = (xa[k] >> 16) - (sm.git0[g] >> 16);
            const uint32_t l4 = (xb[k] & 0xFFFFu) - (sm.gxs0[g] & 0xFFFFu), l2 = (xb[k] >> 16) - (sm.gxs0[g] >> 16);
            uint32_t *table = reinterpret_cast<uint32_t *>(recs + sm.grec[g] + kRecTable);
            uint32_t *items = table + r4(Dg);
            table[dl] = sm.a1[d];
            const uint32_t n16 = pa[k] & 0xFFFFu, rem = cnt0[k] & 15u;
            for (uint32_t v = 0; v < n16; ++v) {
                const uint32_t slots = min(16u, cnt0[k] - 16u * v);
                items[l16 + v] = ((slots - 1u) << 12) | (dl << 17);
            }
            uint32_t rem_item = 0, rem_base = 0;
            if (rem > 4u && rem <= 8u) { rem_item = N16 + l8; rem_base = 18u * N16 + 10u * l8; }
            else if (rem > 2u && rem <= 4u) { rem_item = N16 + N8 + l4; rem_base = 18u * N16 + 10u * N8 + 6u * l4; }
            else if (rem >= 1u && rem <= 2u) { rem_item = N16 + N8 + N4 + l2; rem_base = 18u * N16 + 10u * N8 + 6u * N4 + 2u * l2; }
            if (rem >= 1u && rem <= 8u) items[rem_item] = ((rem - 1u) << 12) | (dl << 17);
            info_a[k] = st0[k] | ((18u * l16) << 12);
            info_b[k] = rem_base | (n16 << 12);
        }
    }
    __syncthreads();   // every thread has read b1[st0], a1[], a2[]; a1 and a2 are reused below
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t s = tid * kIpt + k;
        if (s < S) { sm.a2[s] = info_a[k]; sm.a1[s] = info_b[k]; }
    }
    __syncthreads();
    // F. per alignment: smem byte offsets of its transcript's prev[] copy and of its x slot
    uint32_t lp[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        lp[k] = 0;
        if (keys[k] != kNoTxp) {
            const uint32_t r = tid * kIpt + k, s = seg[k] - 1u, g = keys[k] >> 12;
            const uint32_t ia = sm.a2[s], ib2 = sm.a1[s];
            const uint32_t rr = r - (ia & 0xFFFu), n16 = ib2 >> 12;
            const uint32_t pos = rr < 16u * n16 ? (ia >> 12) + (rr >> 4) * 18u + (rr & 15u) : (ib2 & 0xFFFu) + (rr - 16u * n16);
            lp[k] = ((s - sm.gfs[g]) * 8u) | ((pos * 8u) << 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) if (keys[k] != kNoTxp) sm.b1[vals[k]] = lp[k];
    __syncthreads();
    // pairs go behind their group's record: slot s of the tile belongs to the group with goff[g] <= s < goff[g + 1]
    for (uint32_t s = tid; s < T; s += kBuildThreads) {
        const uint32_t src = sm.b2[s];
        uint32_t lo = 0, hi = G;   // last g with goff[g] <= s
        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (sm.goff[mid] <= s) lo = mid; else hi = mid; }
        const uint32_t g = lo;
        const uint32_t Dg = sm.gfs[g + 1] - sm.gfs[g], ni = ni_of(g);
        uint2 *pairs = reinterpret_cast<uint2 *>(recs + sm.grec[g] + rec_bytes_of(Dg, ni));
        pairs[s - sm.goff[g]] = src != kNoTxp ? make_uint2(__float_as_uint(a.prob[src]), sm.b1[s]) : make_uint2(0u, 0u);
        if (a.aux) a.o_aux[(size_t)pair_base + s] = src != kNoTxp ? a.aux[src] : 1.0;
    }
}

// ---------------------------------------------------------------------------
// the sweep
// ---------------------------------------------------------------------------

// Shared memory is addressed with explicit 32-bit shared-space addresses: every warp works in its own
// region at a runtime offset, and with generic pointers the compiler re-derives the shared window base
// (S2R SR_CgaCtaId + LEA) at most predicated accesses -- 150 times in this kernel.
__device__ __forceinline__ uint32_t ld_s32(uint32_t a) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t ld_s8(uint32_t a) { uint32_t r; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint4 ld_s128(uint32_t a)
{ uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ double ld_sf64(uint32_t a) { double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a)); return r; }
__device__ __forceinline__ double2 ld_s2f64(uint32_t a)
{ double2 r; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a)); return r; }
__device__ __forceinline__ void st_sf64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// {0, dflt} unless `on`
__device__ __forceinline__ uint2 ld_s64_if(uint32_t a, bool on, uint32_t dflt)
{
    uint2 r;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, %4;\n\t"
                 "@p ld.shared.v2.u32 {%0,%1}, [%2];\n\t}" : "=r"(r.x), "=r"(r.y) : "r"(a), "r"((uint32_t)on), "r"(dflt));
    return r;
}
// {0.0, 0.0} unless `on`
__device__ __forceinline__ double2 ld_s2f64_if(uint32_t a, bool on)
{
    double2 r;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
                 "@p ld.shared.v2.f64 {%0,%1}, [%2];\n\t}" : "=d"(r.x), "=d"(r.y) : "r"(a), "r"((uint32_t)on));
    return r;
}
__device__ __forceinline__ double ld_sf64_if(uint32_t a, bool on)
{
    double r;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t"
                 "@p ld.shared.f64 %0, [%1];\n\t}" : "=d"(r) : "r"(a), "r"((uint32_t)on));
    return r;
}
__device__ __forceinline__ void st_sf64_if(uint32_t a, double v, bool on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"((uint32_t)on) : "memory");
}

// One group = up to 32 rows, one per lane, sorted by length (lane 0 longest, L = its length).  The
// j-th alignments of the rows are stored contiguously for the lanes with len > j (a prefix of the
// lanes), so lane l finds its j-th alignment at slot l + sum_{i<j} c_i, c_i = #lanes with len > i.
// J >= L alignments per read live in registers between the denominator pass and the scatter; the
// code is straight-line (inactive slots load nothing and contribute 0), so the J independent
// load -> gather -> multiply chains overlap and the denominator is a pairwise tree.
//   pairs, sp, xp: shared-space addresses of the group's pairs, of prev[] and of the x array
//   idle: lpos of a lane without a j-th alignment (prob 0, table entry 0, the trash x slot): its store needs no predicate
template <int J, bool HAS_AUX>
__device__ __forceinline__ void group_run(uint32_t pairs, const double *gaux, uint32_t sp, uint32_t xp,
                                          uint32_t lane, uint32_t len, double rw, uint32_t idle)
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
    for (int j = 0; j < J; ++j) pr[j] = ld_s64_if(pairs + 8u * off[j], (uint32_t)j < len, idle);
    double w[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const double pv = ld_sf64(sp + (pr[j].y & 0xFFFFu));
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
    for (int j = 0; j < J; ++j) st_sf64(xp + (pr[j].y >> 16), w[j] * inv);   // em.rs:119-130
}

// Groups with rows longer than the register budget: two predicated passes over shared memory.
template <bool HAS_AUX>
__device__ __noinline__ void group_long(uint32_t pairs, const double *gaux, uint32_t sp, uint32_t xp,
                                        uint32_t lane, uint32_t len, uint32_t L, double rw, uint32_t idle)
{
    const unsigned full = 0xffffffffu;
    double denom = 0.0;
    uint32_t off = lane;
    for (uint32_t j = 0; j < L; ++j) {
        const bool act = j < len;
        const uint2 pr = ld_s64_if(pairs + 8u * off, act, idle);
        double ww = ld_sf64(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
        if (HAS_AUX) { if (act) ww *= gaux[off]; }
        denom += ww;
        off += __popc(__ballot_sync(full, act));
    }
    const double inv = (denom > OAR_EM_DENOM_THRESH ? tiled::fast_rcp(denom) : 0.0) * rw;
    off = lane;
    for (uint32_t j = 0; j < L; ++j) {
        const bool act = j < len;
        const uint2 pr = ld_s64_if(pairs + 8u * off, act, idle);
        double ww = ld_sf64(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
        if (HAS_AUX) { if (act) ww *= gaux[off]; }
        st_sf64(xp + (pr.y >> 16), ww * inv);
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
        // this warp's shared memory: [stage 0][stage 1][xs][s_prev][mbarriers]; a stage = one group blob
        const uint32_t sm0 = tiled::smem_u32(smem_all) + warp * g.warp_bytes;
        const uint32_t xp = sm0 + g.xs_off, sp = sm0 + g.prev_off, bar0 = sm0 + g.bar_off;

        auto issue = [&](uint32_t s, uint2 loc) {   // lane 0 only
            const uint32_t bar = bar0 + 8u * s;
            tiled::mbar_expect_tx(bar, loc.y);
            tiled::bulk_g2s(sm0 + s * g.stage_bytes, v.blobs + loc.x, loc.y, bar);
        };

        uint2 pending = make_uint2(0, 0);   // locator of the group two ahead (lane 0)
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
            const uint32_t D = ld_s32(sm0) & 0xFFFFu;
            for (uint32_t d = lane; d < D; d += 32) st_sf64(sp + 8u * d, prev[ld_s32(sm0 + kRecTable + 4u * d)]);
        }
        __syncwarp();

        uint32_t grp = gw;
        for (uint32_t it = 0;; ++it) {
            const uint32_t s = it & 1u;
            const uint32_t rec = sm0 + s * g.stage_bytes;
            const uint4 hdr = ld_s128(rec);
            const uint32_t D = hdr.x & 0xFFFFu, NI = hdr.x >> 16;
            const uint32_t N16 = hdr.y & 0xFFu, N8 = (hdr.y >> 8) & 0xFFu, N4 = (hdr.y >> 16) & 0xFFu;
            const uint32_t table = rec + kRecTable, items = table + 4u * r4(D), pairs = items + 4u * r4(NI);

            // ---- E-step, one read per lane; x = w/denom scattered into the transcript order ----
            {
                const uint32_t len = ld_s8(rec + kRecRlen + lane);
                const uint32_t L = __shfl_sync(full, len, 0);
                const double *gaux = HAS_AUX ? v.aux + hdr.w : nullptr;
                const uint32_t idle = g.trash << 16;
                double rw = 1.0;
                // bootstrap: the read's resampling weight scales its contribution (== visiting it that many times)
                if (HAS_WTS) rw = (double)wperm[hdr.z + lane];
                switch ((L + 1u) >> 1) {   // warp-uniform
                case 0: break;
                case 1: group_run<2, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 2: group_run<4, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 3: group_run<6, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 4: group_run<8, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 5: group_run<10, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 6: group_run<12, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
#if OAR_LANE_REG_ROWS >= 16
                case 7: group_run<14, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
                case 8: group_run<16, HAS_AUX>(pairs, gaux, sp, xp, lane, len, rw, idle); break;
#endif
                default: group_long<HAS_AUX>(pairs, gaux, sp, xp, lane, len, L, rw, idle); break;
                }
            }
            __syncwarp();   // xs complete; s_prev is free again

            // ---- prev[] of the next group (its blob was requested a whole group ago); the L2 gather overlaps the M-step ----
            const uint32_t next = grp + nW;
            const bool has_next = next < n_groups;
            double pv = 0.0;
            uint32_t Dn = 0, table_n = 0;
            if (has_next) {
                tiled::mbar_wait(bar0 + 8u * (s ^ 1u), ((it + 1u) >> 1) & 1u);
                const uint32_t rec_n = sm0 + (s ^ 1u) * g.stage_bytes;
                Dn = ld_s32(rec_n) & 0xFFFFu;
                table_n = rec_n + kRecTable;
                if (lane < Dn) pv = prev[ld_s32(table_n + 4u * lane)];
            }

            // ---- M-step: one lane sums one item (<= 16 consecutive x slots of one transcript), one RED.
            //      Straight-line: 8 predicated LDS.128, four accumulators, no loop control. ----
            for (uint32_t ib = 0; ib < NI; ib += 32) {
                const uint32_t i = ib + lane;
                const uint32_t desc = i < NI ? ld_s32(items + 4u * i) : 0u;
                const uint32_t cnt = i < NI ? ((desc >> 12) & 31u) + 1u : 0u, npair = cnt >> 1;
                uint32_t bd = 18u * i;   // x offset in doubles: size classes 16, 8, 4, 2 at strides 18, 10, 6, 2
                if (i >= N16) bd = 10u * i + 8u * N16;
                if (i >= N16 + N8) bd = 6u * i + 12u * N16 + 4u * N8;
                if (i >= N16 + N8 + N4) bd = 2u * i + 16u * N16 + 8u * N8 + 4u * N4;
                const uint32_t pb = xp + 8u * bd;
                double a0 = ld_sf64_if(pb + 8u * (cnt - 1u), (cnt & 1u) != 0u), a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (uint32_t k = 0; k < (uint32_t)kItem / 2u; k += 2) {
                    const double2 u = ld_s2f64_if(pb + 16u * k, k < npair);
                    const double2 w = ld_s2f64_if(pb + 16u * k + 16u, k + 1u < npair);
                    a0 += u.x; a1 += u.y; a2 += w.x; a3 += w.y;
                }
                const double acc = (a0 + a2) + (a1 + a3);
                if (acc != 0.0) atomicAdd(curr + ld_s32(table + 4u * (desc >> 17)), acc);
            }
            __syncwarp();   // stage s (the blob) and xs are free again

            if (!has_next) break;
            if (lane == 0 && next + nW < n_groups) {
                issue(s, pending);
                if (next + 2 * nW < n_groups) pending = v.groups[next + 2 * nW];
            }
            if (lane < Dn) st_sf64(sp + 8u * lane, pv);
            for (uint32_t d = lane + 32u; d < Dn; d += 32) st_sf64(sp + 8u * d, prev[ld_s32(table_n + 4u * d)]);
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
