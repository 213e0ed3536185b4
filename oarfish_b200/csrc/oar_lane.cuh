// oar_lane.cuh -- the row-per-lane tiled layout and its fused E+M sweep.
//
// Second generation of the locality-tiled design (oar_tiled.cuh).  Same idea:
// rows (reads) ordered by their smallest transcript id so a tile touches a few
// dozen transcripts, prev[] of those gathered once into shared memory, the
// M-step scattered through shared memory in transcript-sorted order and flushed
// with one f64 RED per (warp, transcript).  What changes is the E-step mapping:
//
//   * chunk layout (oar_tiled.cuh): 4 alignment slots per lane, the per-read
//     denominator (em.rs:98-112) is a segmented f64 warp scan driven by lane
//     descriptors -- ~270 warp instructions per 128 alignments, issue-bound.
//   * lane layout (here): ONE READ PER LANE.  The rows of a tile are sorted by
//     length (longest first) and cut into groups of 32; lane l of the warp that
//     owns group g walks row 32g+l serially, so the denominator is a plain
//     per-lane sum in registers: no shuffles, no masks, no descriptors.  The
//     j-th alignments of a group's rows are stored contiguously (lanes with
//     len > j, which is a prefix of the lanes), so every access is a coalesced
//     conflict-free LDS.64 and the HBM stream has no padding at all.
//
// Per alignment the HBM stream is one 8-byte pair {prob f32, lpos u32}; lpos =
// (byte offset of the transcript's prev[] copy in smem) | (byte offset of the
// alignment's slot in the transcript-sorted x array) << 16.
//
// M-step aggregation: a transcript with >= kAggMin alignments in the tile owns
// ceil(cnt/8) 8-slot units (80-byte stride: conflict-free LDS.128); the unused
// slots of its last unit are listed in the tile record and zeroed before the
// scatter.  Alignments of rarer transcripts get a 1-slot "single" each, flushed
// by one RED.  So the scatter never branches: every alignment has a slot.
//
// The sweep is persistent and TMA-fed exactly like the chunk kernel: two stages
// per CTA, cp.async.bulk + mbarrier, issued two tiles ahead by one thread.
// Groups are dealt to warps in snake order (w, 2W-1-w, 2W+w, ...) so that every
// warp gets long and short rows.
#pragma once
#include <cub/cub.cuh>

#include "oar_common.cuh"
#include "oar_kernels.cuh"
#include "oar_tiled.cuh"

namespace oar {
namespace lane {

#ifndef OAR_LANE_THREADS
#define OAR_LANE_THREADS 256
#endif
#ifndef OAR_LANE_IPT
#define OAR_LANE_IPT 9
#endif
#ifndef OAR_LANE_MIN_CTAS
#define OAR_LANE_MIN_CTAS 3
#endif
constexpr int kThreads = OAR_LANE_THREADS;      // sweep CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMinCtas = OAR_LANE_MIN_CTAS;     // register budget: CTAs per SM the sweep is compiled for
constexpr int kBuildThreads = 256;              // layout construction CTA
constexpr int kIpt = OAR_LANE_IPT;              // slots per build thread
constexpr int kSlotsMax = kBuildThreads * kIpt; // alignments a tile can hold (2304)
constexpr int kRowCap = tiled::kChunkCap;       // longer rows are swept from the CSR (fallback list)
constexpr int kSpanMax = kSlotsMax - kRowCap + 1;
constexpr int kSpanDefault = kSpanMax < 2048 ? kSpanMax : 2048;
#ifndef OAR_LANE_REG_ROWS
#define OAR_LANE_REG_ROWS 12
#endif
constexpr int kRegRows = OAR_LANE_REG_ROWS;     // alignments per read the E-step keeps in registers (8, 12 or 16)
static_assert(kRegRows == 8 || kRegRows == 12 || kRegRows == 16, "register rows come in fours");
constexpr int kAggMin = 4;
constexpr int kUnitStride = 10;                 // doubles per 8-slot unit
constexpr uint32_t kNoTxp = 0xFFFFFFFFu;
static_assert(kSlotsMax <= 8192, "row index is packed into 13 bits");
static_assert(kSlotsMax * 20 < 65536, "x-slot byte offsets are 16 bit");

// Per-tile record (16-byte granules, one TMA bulk copy).  Sections, each padded to 16 bytes:
//   header u32[8]: D, U, S1, G, P, row_base, first pair, nnz
//   goff   u32[G]     first slot of group g inside the tile's pair array
//   rlen   u8[32*G]   row lengths in lane order (0 = no row)
//   table  u32[D]     distinct transcript ids
//   units  u32[U]     transcript id of every 8-slot unit
//   single u32[S1]    transcript id of every 1-slot single
//   pads   u16[P]     byte offsets (in the x array) of the unused slots of partial units
constexpr uint32_t kHdrBytes = 32;
struct RecView {   // section positions as byte offsets from the start of the CTA's shared memory (32-bit: cheap to keep)
    uint32_t D, U, S1, G, P, row_base, first;
    uint32_t goff, rlen, table, units, singles, pads;
};
__host__ __device__ inline uint32_t r4(uint32_t x) { return (x + 3u) & ~3u; }
__host__ __device__ inline uint32_t rec_bytes_of(uint32_t D, uint32_t U, uint32_t S1, uint32_t G, uint32_t P)
{ return kHdrBytes + 4u * r4(G) + 32u * G + 4u * r4(D) + 4u * r4(U) + 4u * r4(S1) + 2u * ((P + 7u) & ~7u); }
__device__ __forceinline__ RecView rec_view(const unsigned char *smem, uint32_t rec_off)
{
    const uint4 h0 = *reinterpret_cast<const uint4 *>(smem + rec_off);
    const uint4 h1 = *reinterpret_cast<const uint4 *>(smem + rec_off + 16);
    RecView r;
    r.D = h0.x; r.U = h0.y; r.S1 = h0.z; r.G = h0.w; r.P = h1.x; r.row_base = h1.y; r.first = h1.z;
    r.goff = rec_off + kHdrBytes;
    r.rlen = r.goff + 4u * r4(r.G);
    r.table = r.rlen + 32u * r.G;
    r.units = r.table + 4u * r4(r.D);
    r.singles = r.units + 4u * r4(r.U);
    r.pads = r.singles + 4u * r4(r.S1);
    return r;
}
template <typename T> __device__ __forceinline__ T lds(const unsigned char *smem, uint32_t off)
{ return *reinterpret_cast<const T *>(smem + off); }

// Shared memory of the sweep, sized for the store at hand: two data stages (pairs), a ring of THREE
// record buffers (the record of tile i is still read by phase 2 of tile i while the copies for tile
// i+2 are in flight, so records cannot share the two-deep data ring), the x array, prev[] and two mbarriers.
struct Geometry {
    uint32_t data_bytes;   // pairs of the largest tile, 128 B multiple
    uint32_t rec_bytes;    // largest record, 128 B multiple
    uint32_t rec_off, xs_off, prev_off, bar_off, total;
};
inline Geometry make_geometry(uint32_t max_nnz, uint32_t max_rec_bytes, uint32_t max_d, uint32_t max_xs_bytes)
{
    Geometry g;
    g.data_bytes = (((max_nnz + 1u) & ~1u) * 8u + 127u) & ~127u;
    g.rec_bytes = (max_rec_bytes + 127u) & ~127u;
    g.rec_off = 2u * g.data_bytes;
    g.xs_off = g.rec_off + 3u * g.rec_bytes;
    g.prev_off = g.xs_off + ((max_xs_bytes + 15u) & ~15u);
    g.bar_off = g.prev_off + 8u * ((max_d + 1u) & ~1u);
    g.total = g.bar_off + 16u;
    return g;
}

struct View {
    uint32_t n_tiles;
    const uint2 *pairs;        // {prob bits, lpos}; tile t starts at tiles[t].z (even)
    const double *aux;         // same indexing, or null
    const uint4 *tiles;        // {record offset (16 B granules), record bytes, first pair, nnz}
    const uint4 *records;
    const uint32_t *fb_rows; uint32_t n_fb;
    const uint32_t *csr_row_ptr; const uint32_t *csr_txp; const float *csr_prob; const double *csr_aux;
    const uint32_t *csr_wts;
};

// ---------------------------------------------------------------------------
// layout construction: one CTA per tile
// ---------------------------------------------------------------------------

struct BuildArgs {
    const uint32_t *row_ptr; const uint32_t *txp; const float *prob; const double *aux;
    const uint32_t *srow;      // sorted position -> original row
    const uint32_t *soff;      // exclusive prefix of the sorted rows' lengths (n_tiled + 1)
    const uint32_t *tile_row;  // n_tiles + 1
    uint2 *o_pairs; double *o_aux; uint4 *o_tiles; uint4 *o_records;
    uint32_t *o_trow;          // tile-order row -> original row
    // [0] record granules, [1] sum D, [2] sum U, [3] sum S1, [4] max record bytes, [5] max D, [6] max x bytes, [7] max nnz, [8] sum P
    uint32_t *cursors;
};

using BSortK = cub::BlockRadixSort<uint32_t, kBuildThreads, kIpt>;
using BSortKV = cub::BlockRadixSort<uint32_t, kBuildThreads, kIpt, uint32_t>;
using BScan = cub::BlockScan<uint32_t, kBuildThreads>;
union BuildTemp { typename BSortK::TempStorage sortk; typename BSortKV::TempStorage sortkv; typename BScan::TempStorage scan; };

struct BuildSmem {
    uint32_t a1[kSlotsMax + 4];   // row-contiguous transcript ids -> per-segment info
    uint32_t a2[kSlotsMax + 4];   // row-contiguous source indices -> segment starts -> lpos
    uint32_t b1[kSlotsMax];       // slot -> transcript -> sorted keys
    uint32_t b2[kSlotsMax];       // slot -> source alignment index in the CSR
    uint16_t rord[kSlotsMax];     // length-sorted position -> row of the tile
    uint16_t coff[kSlotsMax];     // row of the tile -> offset of its alignments in a1/a2
    uint8_t plen[kSlotsMax + 32]; // length-sorted position -> row length
    uint32_t goff[kSlotsMax / 32 + 8];
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
    const uint32_t G = (nrows + 31u) >> 5;
    const uint32_t start = (base_off + tile + 1u) & ~1u;   // first pair of this tile (even: TMA needs 16 B alignment)

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

    // C. group offsets, then the (group, j, lane) slot order
    for (uint32_t g = warp; g < G; g += kBuildThreads / 32) {
        uint32_t sum = sm.plen[g * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
        if (lane == 0) sm.goff[g] = sum;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t g = 0; g < G; ++g) { const uint32_t t = sm.goff[g]; sm.goff[g] = acc; acc += t; }
        sm.goff[G] = acc;
    }
    __syncthreads();
    for (uint32_t g = warp; g < G; g += kBuildThreads / 32) {
        const uint32_t p = g * 32 + lane;
        const uint32_t len = sm.plen[p];
        const uint32_t i = p < nrows ? sm.rord[p] : 0u;
        const uint32_t o = sm.coff[i];
        if (p < nrows) a.o_trow[r0 + p] = a.srow[r0 + i];
        const uint32_t L = __shfl_sync(full, len, 0);
        uint32_t off = sm.goff[g] + lane;
        for (uint32_t j = 0; j < L; ++j) {
            const bool act = j < len;
            if (act) { sm.b1[off] = sm.a1[o + j]; sm.b2[off] = sm.a2[o + j]; }
            off += __popc(__ballot_sync(full, act));
        }
    }
    for (uint32_t s = nnz + tid; s < (uint32_t)kSlotsMax; s += kBuildThreads) { sm.b1[s] = kNoTxp; sm.b2[s] = kNoTxp; }
    __syncthreads();

    // D. slots sorted by transcript (stable: the alignments one STS of the sweep scatters for a
    //    transcript get consecutive ranks, i.e. consecutive banks)
    uint32_t keys[kIpt], vals[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) { const uint32_t s = tid * kIpt + k; keys[k] = sm.b1[s]; vals[k] = s; }
    __syncthreads();
    BSortKV(sm.tmp.sortkv).Sort(keys, vals);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) sm.b1[tid * kIpt + k] = keys[k];
    __syncthreads();
    uint32_t hf[kIpt], seg[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t r = tid * kIpt + k;
        hf[k] = (keys[k] != kNoTxp && (r == 0 || sm.b1[r - 1] != keys[k])) ? 1u : 0u;
    }
    uint32_t D = 0;
    BScan(sm.tmp.scan).InclusiveSum(hf, seg, D);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIpt; ++k) if (hf[k]) sm.a2[seg[k] - 1] = tid * kIpt + k;
    if (tid == 0) sm.a2[D] = nnz;
    __syncthreads();

    // E. per transcript: units (cnt >= kAggMin) or singles, and the unused slots of the last unit
    uint32_t nun[kIpt], nsg[kIpt], npd[kIpt], st0[kIpt], ubase[kIpt], sbase[kIpt], pbase[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t d = tid * kIpt + k;
        nun[k] = nsg[k] = npd[k] = st0[k] = 0;
        if (d < D) {
            st0[k] = sm.a2[d];
            const uint32_t cnt = sm.a2[d + 1] - st0[k];
            if (cnt >= (uint32_t)kAggMin) { nun[k] = (cnt + 7u) >> 3; npd[k] = nun[k] * 8u - cnt; }
            else nsg[k] = cnt;
        }
    }
    uint32_t U = 0, S1 = 0, P = 0;
    BScan(sm.tmp.scan).ExclusiveSum(nun, ubase, U);
    __syncthreads();
    BScan(sm.tmp.scan).ExclusiveSum(nsg, sbase, S1);
    __syncthreads();
    BScan(sm.tmp.scan).ExclusiveSum(npd, pbase, P);
    __syncthreads();   // also: every thread has read a2[d], a2[d + 1]
    const uint32_t rec_bytes = rec_bytes_of(D, U, S1, G, P);
    const uint32_t xs_bytes = U * (uint32_t)kUnitStride * 8u + S1 * 8u;
    if (tid == 0) {
        sm.misc[0] = atomicAdd(a.cursors + 0, rec_bytes / 16u);
        atomicAdd(a.cursors + 1, D); atomicAdd(a.cursors + 2, U); atomicAdd(a.cursors + 3, S1); atomicAdd(a.cursors + 8, P);
        atomicMax(a.cursors + 4, rec_bytes); atomicMax(a.cursors + 5, D); atomicMax(a.cursors + 6, xs_bytes);
        atomicMax(a.cursors + 7, nnz);
    }
    __syncthreads();
    const uint32_t rec_off = sm.misc[0];
    unsigned char *rec = reinterpret_cast<unsigned char *>(a.o_records + rec_off);
    uint32_t *hdr = reinterpret_cast<uint32_t *>(rec);
    uint32_t *o_goff = reinterpret_cast<uint32_t *>(rec + kHdrBytes);
    uint8_t *o_rlen = rec + kHdrBytes + 4u * r4(G);
    uint32_t *o_table = reinterpret_cast<uint32_t *>(o_rlen + 32u * G);
    uint32_t *o_units = o_table + r4(D);
    uint32_t *o_singles = o_units + r4(U);
    uint16_t *o_pads = reinterpret_cast<uint16_t *>(o_singles + r4(S1));
    if (tid == 0) {
        hdr[0] = D; hdr[1] = U; hdr[2] = S1; hdr[3] = G; hdr[4] = P; hdr[5] = r0; hdr[6] = start; hdr[7] = nnz;
        a.o_tiles[tile] = make_uint4(rec_off, rec_bytes, start, nnz);
    }
    for (uint32_t g = tid; g < r4(G); g += kBuildThreads) o_goff[g] = g < G ? sm.goff[g] : 0u;
    for (uint32_t p = tid; p < 32u * G; p += kBuildThreads) o_rlen[p] = sm.plen[p];
    for (uint32_t d = D + tid; d < r4(D); d += kBuildThreads) o_table[d] = 0u;
    for (uint32_t u = U + tid; u < r4(U); u += kBuildThreads) o_units[u] = kNoTxp;
    for (uint32_t u = S1 + tid; u < r4(S1); u += kBuildThreads) o_singles[u] = 0u;
    for (uint32_t u = P + tid; u < ((P + 7u) & ~7u); u += kBuildThreads) o_pads[u] = 0;
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        const uint32_t d = tid * kIpt + k;
        if (d < D) {
            const uint32_t key = sm.b1[st0[k]];
            o_table[d] = key;
            for (uint32_t v = 0; v < nun[k]; ++v) o_units[ubase[k] + v] = key;
            for (uint32_t v = 0; v < nsg[k]; ++v) o_singles[sbase[k] + v] = key;
            for (uint32_t v = 0; v < npd[k]; ++v) {
                const uint32_t p = (ubase[k] + nun[k]) * 8u - npd[k] + v;   // slot index in unit space
                o_pads[pbase[k] + v] = (uint16_t)((p + ((p >> 3) << 1)) * 8u);
            }
            // start (13 bits) | unit or single base (13 bits) | kind
            sm.a1[d] = st0[k] | ((nun[k] ? ubase[k] : sbase[k]) << 13) | (nun[k] ? 0u : 1u << 26);
        }
    }
    __syncthreads();

    // F. per alignment: smem byte offsets of its transcript's prev[] copy and of its x slot
#pragma unroll
    for (int k = 0; k < kIpt; ++k) {
        if (keys[k] != kNoTxp) {
            const uint32_t r = tid * kIpt + k, d = seg[k] - 1u;
            const uint32_t pk = sm.a1[d];
            const uint32_t st = pk & 0x1FFFu, base = (pk >> 13) & 0x1FFFu;
            uint32_t posb;
            if ((pk >> 26) == 0u) { const uint32_t p = base * 8u + (r - st); posb = (p + ((p >> 3) << 1)) * 8u; }
            else posb = U * (uint32_t)kUnitStride * 8u + (base + (r - st)) * 8u;
            sm.a2[vals[k]] = (d * 8u) | (posb << 16);
        }
    }
    __syncthreads();
    for (uint32_t s = tid; s < nnz; s += kBuildThreads) {
        const uint32_t src = sm.b2[s];
        a.o_pairs[(size_t)start + s] = make_uint2(__float_as_uint(a.prob[src]), sm.a2[s]);
        if (a.aux) a.o_aux[(size_t)start + s] = a.aux[src];
    }
}

// ---------------------------------------------------------------------------
// the sweep
// ---------------------------------------------------------------------------

// One group = 32 rows, one per lane, sorted by length (lane 0 longest).  Alignments j < nfull exist in
// every lane: they are read at fixed offsets (gl + 256 j), kept in registers between the denominator
// pass and the scatter, and need no lane predicates.  The code is specialised for J-4 < nfull <= J and
// is straight-line (slots j >= nfull re-read slot 0 and contribute 0), so the J independent
// load -> gather -> multiply chains overlap.  The ragged tail nfull <= j < L (and anything beyond the
// register rows) takes two predicated passes over shared memory; its active lanes are a prefix.
template <int J, bool HAS_AUX>
__device__ __forceinline__ void group_run(const unsigned char *gl, const double *gaux_l, const char *sp, char *xp,
                                          uint32_t lane, uint32_t len, uint32_t L, uint32_t nfull, double rw)
{
    const unsigned full = 0xffffffffu;
    double w[J];
    uint32_t qp[J / 2];   // x-slot byte offsets, two per register
    double denom = 0.0;
    if (nfull) {          // warp-uniform
        uint2 pr[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const bool ok = (j <= J - 4) || ((uint32_t)j < nfull);
            pr[j] = *reinterpret_cast<const uint2 *>(gl + (ok ? 256 * j : 0));
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const bool ok = (j <= J - 4) || ((uint32_t)j < nfull);
            const double pv = *reinterpret_cast<const double *>(sp + (pr[j].y & 0xFFFFu));
            double ww = pv * (double)__uint_as_float(pr[j].x);       // em.rs:107
            if (HAS_AUX) ww *= gaux_l[ok ? 32 * j : 0];                // em.rs:108-111
            w[j] = ok ? ww : 0.0;
            if (j & 1) qp[j >> 1] = __byte_perm(qp[j >> 1], pr[j].y, 0x7610);   // {lo: even slot, hi: odd slot}
            else qp[j >> 1] = pr[j].y >> 16;
            denom += w[j];
        }
    }
    if (nfull < L) {
        uint32_t off = nfull * 32u;
        for (uint32_t j = nfull; j < L; ++j) {
            const bool act = j < len;
            if (act) {
                const uint2 pr = *reinterpret_cast<const uint2 *>(gl + 8u * off);
                double ww = *reinterpret_cast<const double *>(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
                if (HAS_AUX) ww *= gaux_l[off];
                denom += ww;
            }
            off += __popc(__ballot_sync(full, act));
        }
    }
    // reads whose denominator is <= 1e-30 contribute nothing (em.rs:115)
    const double inv = (denom > OAR_EM_DENOM_THRESH ? tiled::fast_rcp(denom) : 0.0) * rw;
    if (nfull) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const bool ok = (j <= J - 4) || ((uint32_t)j < nfull);
            const uint32_t q = (j & 1) ? (qp[j >> 1] >> 16) : (qp[j >> 1] & 0xFFFFu);
            if (ok) *reinterpret_cast<double *>(xp + q) = w[j] * inv;  // em.rs:119-130
        }
    }
    if (nfull < L) {
        uint32_t off = nfull * 32u;
        for (uint32_t j = nfull; j < L; ++j) {
            const bool act = j < len;
            if (act) {
                const uint2 pr = *reinterpret_cast<const uint2 *>(gl + 8u * off);
                double ww = *reinterpret_cast<const double *>(sp + (pr.y & 0xFFFFu)) * (double)__uint_as_float(pr.x);
                if (HAS_AUX) ww *= gaux_l[off];
                *reinterpret_cast<double *>(xp + (pr.y >> 16)) = ww * inv;
            }
            off += __popc(__ballot_sync(full, act));
        }
    }
}

// m_step (em.rs:87-133), persistent and TMA-fed.
template <bool HAS_AUX, bool HAS_WTS>
__global__ void __launch_bounds__(kThreads, kMinCtas) em_sweep_lane(View v, Geometry g, const double *__restrict__ prev,
                                                                    double *__restrict__ curr,
                                                                    const uint32_t *__restrict__ wperm,
                                                                    const OarEmState *__restrict__ st, int check_done)
{
    using tiled::mask01;
    extern __shared__ __align__(128) unsigned char smem[];
    // [data 0][data 1][record 0][record 1][record 2][xs][s_prev][mbarriers]
    char *xp = reinterpret_cast<char *>(smem + g.xs_off);
    double *s_prev = reinterpret_cast<double *>(smem + g.prev_off);
    const char *sp = reinterpret_cast<const char *>(s_prev);

    if (check_done && st->done) return;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tiles = v.n_tiles, stride = gridDim.x;
    const uint32_t tile0 = blockIdx.x;
    if (tile0 >= n_tiles) return;
    const uint32_t bar0 = tiled::smem_u32(smem + g.bar_off), smem0 = tiled::smem_u32(smem);

    // data of the k-th tile of this CTA -> data stage k & 1, its record -> record buffer k % 3; one mbarrier per data stage
    auto issue = [&](uint32_t s, uint32_t rb, uint4 ti) {   // thread 0 only
        const uint32_t bar = bar0 + 8u * s;
        const uint32_t dbytes = ((ti.w + 1u) & ~1u) * 8u;
        tiled::mbar_expect_tx(bar, dbytes + ti.y);
        if (dbytes) tiled::bulk_g2s(smem0 + s * g.data_bytes, v.pairs + ti.z, dbytes, bar);
        tiled::bulk_g2s(smem0 + g.rec_off + rb * g.rec_bytes, v.records + ti.x, ti.y, bar);
    };

    uint4 t_pending = make_uint4(0, 0, 0, 0);   // locator of the tile two ahead (thread 0)
    if (tid == 0) {
        tiled::mbar_init(bar0, 1);
        tiled::mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(0, 0, v.tiles[tile0]);
        if (tile0 + stride < n_tiles) issue(1, 1, v.tiles[tile0 + stride]);
        if (tile0 + 2 * stride < n_tiles) t_pending = v.tiles[tile0 + 2 * stride];
    }
    __syncthreads();
    tiled::mbar_wait(bar0, 0);
    {
        const RecView r = rec_view(smem, g.rec_off);
        for (uint32_t d = tid; d < r.D; d += kThreads) s_prev[d] = prev[lds<uint32_t>(smem, r.table + 4u * d)];
    }

    uint32_t tile = tile0, rb = 0;   // rb = it % 3
    for (uint32_t it = 0;; ++it) {
        const uint32_t s = it & 1u;
        const uint32_t rb1 = rb == 2u ? 0u : rb + 1u, rb2 = rb1 == 2u ? 0u : rb1 + 1u;
        const unsigned char *stg = smem + s * g.data_bytes;
        const RecView r = rec_view(smem, g.rec_off + rb * g.rec_bytes);
        __syncthreads();   // s_prev of this tile is in place; phase 2 of the previous tile has left xs and record rb2

        // unused slots of partial units: zero them (disjoint from every scatter target)
        for (uint32_t k = tid; k < r.P; k += kThreads) *reinterpret_cast<double *>(xp + lds<uint16_t>(smem, r.pads + 2u * k)) = 0.0;

        // ---- phase 1: E-step, one read per lane; x = w/denom scattered into transcript order ----
        for (uint32_t gb = 0, round = 0; gb < r.G; gb += kWarps, ++round) {
            const uint32_t grp = gb + ((round & 1u) ? (uint32_t)kWarps - 1u - warp : warp);
            if (grp >= r.G) continue;
            const uint32_t len = lds<uint8_t>(smem, r.rlen + grp * 32u + lane);
            const uint32_t L = __shfl_sync(full, len, 0), Lmin = __shfl_sync(full, len, 31);
            const uint32_t go = lds<uint32_t>(smem, r.goff + 4u * grp);
            const unsigned char *gl = stg + 8u * (go + lane);
            const double *gaux_l = HAS_AUX ? v.aux + r.first + go + lane : nullptr;
            double rw = 1.0;
            // bootstrap: the read's resampling weight scales its contribution (== visiting it that many times)
            if (HAS_WTS) rw = (double)wperm[r.row_base + grp * 32u + lane];
            const uint32_t nfull = min(Lmin, (uint32_t)kRegRows);
            if (nfull <= 4u) group_run<4, HAS_AUX>(gl, gaux_l, sp, xp, lane, len, L, nfull, rw);
            else if (nfull <= 8u) group_run<8, HAS_AUX>(gl, gaux_l, sp, xp, lane, len, L, nfull, rw);
            else if (kRegRows >= 12 && nfull <= 12u) group_run<(kRegRows >= 12 ? 12 : 8), HAS_AUX>(gl, gaux_l, sp, xp, lane, len, L, nfull, rw);
            else group_run<kRegRows, HAS_AUX>(gl, gaux_l, sp, xp, lane, len, L, nfull, rw);
        }
        __syncthreads();   // xs complete; data stage s and s_prev are free again

        // ---- refill two tiles ahead; fetch prev[] of the next tile behind its record ----
        const uint32_t next = tile + stride;
        const bool has_next = next < n_tiles;
        if (tid == 0 && next + stride < n_tiles) {
            issue(s, rb2, t_pending);
            if (next + 2 * stride < n_tiles) t_pending = v.tiles[next + 2 * stride];
        }
        double pv = 0.0;
        uint32_t Dn = 0, table_n = 0;
        if (has_next) {
            tiled::mbar_wait(bar0 + 8u * (s ^ 1u), ((it + 1u) >> 1) & 1u);
            const RecView rn = rec_view(smem, g.rec_off + rb1 * g.rec_bytes);
            Dn = rn.D; table_n = rn.table;
            if (tid < Dn) pv = prev[lds<uint32_t>(smem, table_n + 4u * tid)];
        }

        // ---- phase 2: sum 8-slot units, combine equal transcripts across the warp, one RED per run ----
        for (uint32_t ub = warp * 32u; ub < r.U; ub += kThreads) {
            const uint32_t u = ub + lane;
            const bool valid = u < r.U;
            const uint32_t u_txp = valid ? lds<uint32_t>(smem, r.units + 4u * u) : kNoTxp;
            double acc = 0.0;
            if (valid) {
                const double2 *b = reinterpret_cast<const double2 *>(xp + u * (uint32_t)(kUnitStride * 8));
                const double2 v0 = b[0], v1 = b[1], v2 = b[2], v3 = b[3];
                acc = ((v0.x + v0.y) + (v1.x + v1.y)) + ((v2.x + v2.y) + (v3.x + v3.y));
            }
            const uint32_t up = __shfl_up_sync(full, u_txp, 1);
            const uint32_t dn = __shfl_down_sync(full, u_txp, 1);
            const bool head = (lane == 0) || (up != u_txp);
            const bool tail = (lane == 31) || (dn != u_txp);
            const unsigned hmask = __ballot_sync(full, head);
            const uint32_t dist2 = lane - (31u - __clz(hmask & (full >> (31u - lane))));
            const uint32_t maxd = __reduce_max_sync(full, dist2);     // longest run of one transcript in the warp
            acc = fma(__shfl_up_sync(full, acc, 1), mask01(dist2 >= 1u), acc);
            if (maxd >= 2u) {
                acc = fma(__shfl_up_sync(full, acc, 2), mask01(dist2 >= 2u), acc);
                if (maxd >= 4u) {
                    acc = fma(__shfl_up_sync(full, acc, 4), mask01(dist2 >= 4u), acc);
                    if (maxd >= 8u) {
                        acc = fma(__shfl_up_sync(full, acc, 8), mask01(dist2 >= 8u), acc);
                        acc = fma(__shfl_up_sync(full, acc, 16), mask01(dist2 >= 16u), acc);
                    }
                }
            }
            if (tail && valid) atomicAdd(curr + u_txp, acc);
        }
        {
            const double *x1 = reinterpret_cast<const double *>(xp + r.U * (uint32_t)(kUnitStride * 8));
            for (uint32_t k = tid; k < r.S1; k += kThreads) {
                const double x = x1[k];
                if (x != 0.0) atomicAdd(curr + lds<uint32_t>(smem, r.singles + 4u * k), x);
            }
        }

        if (!has_next) break;
        if (tid < Dn) s_prev[tid] = pv;
        for (uint32_t d = tid + kThreads; d < Dn; d += kThreads) s_prev[d] = prev[lds<uint32_t>(smem, table_n + 4u * d)];
        tile = next;
        rb = rb1;
    }
    if (v.n_fb && blockIdx.x == gridDim.x - 1u)
        kern::rowgroup_rows<HAS_AUX, HAS_WTS>(v.csr_row_ptr, v.csr_txp, v.csr_prob, v.csr_aux, v.csr_wts, v.fb_rows, prev, curr,
                                              tid >> 3, kThreads >> 3, v.n_fb);
}

}  // namespace lane
}  // namespace oar
