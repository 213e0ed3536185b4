// oar_cells.cu -- batched per-cell EM for single-cell mode.
//
// The reference runs one independent em::em(&emi, 1) per cell barcode on a worker
// thread (src/single_cell.rs:91-193, call at :150), i.e. do_em (em.rs:144-255,
// stop rule niter > 50) over the cell's reads with the FULL transcriptome as
// parameter space and the uniform start N_cell / M.  Transcripts without an
// alignment in the cell can never receive mass (their prev only feeds
// denominators of rows that do not contain them), so each cell's EM runs exactly
// on its own compacted transcript set:
//
//   cell_localize  one CTA per cell: bitmap of the cell's transcripts in shared
//                  memory -> sorted distinct list + per-alignment local id
//   cell_em        one CTA per cell, resident for the whole EM: sweep (8 lanes per
//                  read row, f64 RED into the cell's L2-resident count table),
//                  rel-diff reduce, stop rule, final threshold + sweep -- all
//                  iterations of a cell in ONE launch, no host round trips
#include <algorithm>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "oar_store.cuh"

namespace oar {
namespace cells {

constexpr int kThreads = 512;

// ---- localisation ---------------------------------------------------------
// pass A: count distinct transcripts per cell (bitmap in smem or global scratch)
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(kThreads) cell_localize(const uint32_t *__restrict__ row_ptr,
                                                          const uint32_t *__restrict__ txp,
                                                          const uint64_t *__restrict__ cell_rows, uint32_t n_cells,
                                                          uint32_t n_txps, uint32_t words,
                                                          uint32_t *__restrict__ gscratch,  // per-CTA bitmap+prefix if it does not fit smem
                                                          uint64_t *__restrict__ cell_d,    // COUNT_ONLY: out counts (n_cells+1, [c+1]); else: in offsets
                                                          uint32_t *__restrict__ cell_txps, uint32_t *__restrict__ lid,
                                                          int use_smem)
{
    extern __shared__ uint32_t sm[];
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_total;
    uint32_t *bits = use_smem ? sm : gscratch + (size_t)blockIdx.x * 2 * words;
    uint32_t *pre = bits + words;
    for (uint32_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const uint32_t r0 = (uint32_t)cell_rows[c], r1 = (uint32_t)cell_rows[c + 1];
        const uint32_t a0 = row_ptr[r0], a1 = row_ptr[r1];
        for (uint32_t w = threadIdx.x; w < words; w += kThreads) bits[w] = 0;
        __syncthreads();
        for (uint32_t j = a0 + threadIdx.x; j < a1; j += kThreads) {
            const uint32_t t = txp[j];
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        // exclusive prefix of popcounts over words (block scan in chunks of kThreads)
        uint32_t running = 0;
        for (uint32_t base = 0; base < words; base += kThreads) {
            const uint32_t w = base + threadIdx.x;
            const uint32_t v = w < words ? __popc(bits[w]) : 0u;
            uint32_t incl = v;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= (unsigned)o) incl += t; }
            if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
            __syncthreads();
            if (threadIdx.x < 32) {
                uint32_t x = threadIdx.x < kThreads / 32 ? s_warp[threadIdx.x] : 0u;
                uint32_t xi = x;
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (threadIdx.x >= (unsigned)o) xi += t; }
                if (threadIdx.x < kThreads / 32) s_warp[threadIdx.x] = xi - x;
                if (threadIdx.x == kThreads / 32 - 1) s_total = xi;
            }
            __syncthreads();
            if (w < words) pre[w] = running + s_warp[threadIdx.x >> 5] + incl - v;
            running += s_total;
            __syncthreads();
        }
        if (COUNT_ONLY) {
            if (threadIdx.x == 0) cell_d[c + 1] = running;
        } else {
            const uint64_t d0 = cell_d[c];
            for (uint32_t w = threadIdx.x; w < words; w += kThreads) {
                uint32_t b = bits[w];
                uint32_t k = pre[w];
                while (b) { const int bit = __ffs(b) - 1; b &= b - 1; cell_txps[d0 + k++] = (w << 5) + bit; }
            }
            for (uint32_t j = a0 + threadIdx.x; j < a1; j += kThreads) {
                const uint32_t t = txp[j];
                lid[j] = (uint32_t)d0 + pre[t >> 5] + __popc(bits[t >> 5] & ((1u << (t & 31)) - 1u));
            }
        }
        __syncthreads();
    }
    (void)n_txps;
}

__device__ __forceinline__ double ld_cg(const double *p)
{ double v; asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

// one E+M sweep of a cell by the whole CTA (8-lane groups, em.rs:87-133)
template <bool HAS_AUX>
__device__ __forceinline__ void cell_sweep(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ lid,
                                           const float *__restrict__ prob, const double *__restrict__ aux,
                                           uint32_t r0, uint32_t r1, const double *prev, double *curr)
{
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u;
    const unsigned gmask = 0xFFu << (lane & 24u);
    for (uint32_t row = r0 + (threadIdx.x >> 3); row < r1; row += kThreads / 8) {
        const uint32_t s = row_ptr[row], e = row_ptr[row + 1];
        const uint32_t j0 = s + sub;
        uint32_t t0 = 0; double w0 = 0.0;
        if (j0 < e) { t0 = lid[j0]; w0 = ld_cg(prev + t0) * (double)prob[j0]; if (HAS_AUX) w0 *= aux[j0]; }
        double denom = w0;
        for (uint32_t j = j0 + 8; j < e; j += 8) {
            double w = ld_cg(prev + lid[j]) * (double)prob[j];
            if (HAS_AUX) w *= aux[j];
            denom += w;
        }
        denom += __shfl_xor_sync(gmask, denom, 1);
        denom += __shfl_xor_sync(gmask, denom, 2);
        denom += __shfl_xor_sync(gmask, denom, 4);
        if (denom > OAR_EM_DENOM_THRESH) {
            if (j0 < e) atomicAdd(curr + t0, w0 / denom);
            for (uint32_t j = j0 + 8; j < e; j += 8) {
                const uint32_t t = lid[j];
                double w = ld_cg(prev + t) * (double)prob[j];
                if (HAS_AUX) w *= aux[j];
                atomicAdd(curr + t, w / denom);
            }
        }
    }
}

// do_em (em.rs:144-255) for one cell per CTA; every iteration of the cell inside this launch.
template <bool HAS_AUX>
__global__ void __launch_bounds__(kThreads) cell_em(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ lid,
                                                    const float *__restrict__ prob, const double *__restrict__ aux,
                                                    const uint64_t *__restrict__ cell_rows, const uint64_t *__restrict__ cell_d,
                                                    uint32_t n_cells, uint32_t n_txps, double *bufA, double *bufB,
                                                    uint32_t max_iter, double thr, uint32_t min_iter,
                                                    double *__restrict__ out_val, uint32_t *__restrict__ out_niter)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_rel;
    for (uint32_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const uint32_t r0 = (uint32_t)cell_rows[c], r1 = (uint32_t)cell_rows[c + 1];
        const uint64_t d0 = cell_d[c];
        const uint32_t L = (uint32_t)(cell_d[c + 1] - d0);
        double *prev = bufA, *curr = bufB;   // alignments carry global compact ids (cell offset + local index)
        const double avg = (double)(r1 - r0) / (double)n_txps;    // em.rs:154,165: N_cell / M (full transcriptome)
        for (uint32_t i = threadIdx.x; i < L; i += kThreads) { prev[d0 + i] = avg; curr[d0 + i] = 0.0; }
        __threadfence();
        __syncthreads();
        uint32_t niter = 0;
        while (niter < max_iter) {
            cell_sweep<HAS_AUX>(row_ptr, lid, prob, aux, r0, r1, prev, curr);
            __threadfence();
            __syncthreads();
            double m = 0.0;
            for (uint32_t i = threadIdx.x; i < L; i += kThreads) {
                const double pc = ld_cg(prev + d0 + i), cc = ld_cg(curr + d0 + i);
                if (pc > OAR_MIN_READ_THRESH) { const double rd = (cc - pc) / pc; m = rd > m ? rd : m; }   // em.rs:194-201
                prev[d0 + i] = 0.0;                                                                          // swap + fill(0)
            }
            for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                double b = 0.0;
                for (int w = 0; w < kThreads / 32; ++w) b = s_red[w] > b ? s_red[w] : b;
                s_rel = b;
            }
            __syncthreads();
            const double rel = s_rel;
            double *t = prev; prev = curr; curr = t;
            if (rel < thr && niter > min_iter) break;     // em.rs:212
            ++niter;
        }
        for (uint32_t i = threadIdx.x; i < L; i += kThreads)
            if (ld_cg(prev + d0 + i) < OAR_MIN_READ_THRESH) prev[d0 + i] = 0.0;   // em.rs:238-242
        __threadfence();
        __syncthreads();
        cell_sweep<HAS_AUX>(row_ptr, lid, prob, aux, r0, r1, prev, curr);   // em.rs:245-252
        __threadfence();
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < L; i += kThreads) out_val[d0 + i] = ld_cg(curr + d0 + i);
        if (threadIdx.x == 0) out_niter[c] = niter;
        __syncthreads();
    }
}


// ---- batched cells on the tiled sweep -----------------------------------------------------------
// With (cell, transcript) pairs as the parameter space the batch is ONE block-diagonal EM: the tiled sweep
// runs over all reads of all cells at once; only the convergence logic is per cell.
enum : uint32_t { kRun = 0, kFinal = 1, kDone = 2 };
struct CellState { uint32_t niter; uint32_t phase; };
struct Chunk { uint32_t cell; uint32_t first; uint64_t begin, end; };   // a slice of one cell's id range

__global__ void cells_init(const uint64_t *__restrict__ cell_rows, const uint64_t *__restrict__ cell_d, uint32_t n_cells,
                           uint32_t n_txps, uint32_t max_iter, double *__restrict__ A, double *__restrict__ B,
                           CellState *__restrict__ cs)
{
    for (uint32_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const uint64_t d0 = cell_d[c], d1 = cell_d[c + 1];
        double avg = (double)(cell_rows[c + 1] - cell_rows[c]) / (double)n_txps;   // em.rs:154,165
        if (max_iter == 0 && avg < OAR_MIN_READ_THRESH) avg = 0.0;               // no loop: threshold the start (em.rs:238)
        for (uint64_t i = d0 + threadIdx.x; i < d1; i += blockDim.x) { A[i] = avg; B[i] = 0.0; }
        if (threadIdx.x == 0) { cs[c].niter = 0; cs[c].phase = max_iter == 0 ? kFinal : kRun; }
    }
}

// After a sweep prev -> curr, pass 1: per-cell max signed relative difference (em.rs:194-201), one CTA per
// chunk of a cell's ids, combined with an atomic max on the bits of the (non-negative) f64.
__global__ void __launch_bounds__(256) cells_reduce(const double *__restrict__ prev, const double *__restrict__ curr,
                                                    const Chunk *__restrict__ chunks, uint32_t n_chunks,
                                                    const CellState *__restrict__ cs, unsigned long long *__restrict__ rel_bits,
                                                    const OarEmState *__restrict__ st)
{
    if (st->done) return;
    __shared__ double s_red[8];
    for (uint32_t k = blockIdx.x; k < n_chunks; k += gridDim.x) {
        const Chunk ch = chunks[k];
        if (cs[ch.cell].phase != kRun) continue;
        double m = 0.0;
        for (uint64_t i = ch.begin + threadIdx.x; i < ch.end; i += blockDim.x) {
            const double pc = prev[i], cc = curr[i];
            if (pc > OAR_MIN_READ_THRESH) { const double rd = (cc - pc) / pc; m = rd > m ? rd : m; }
        }
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0.0;
            for (int w = 0; w < 8; ++w) b = s_red[w] > b ? s_red[w] : b;
            if (b > 0.0) atomicMax(rel_bits + ch.cell, (unsigned long long)__double_as_longlong(b));
        }
        __syncthreads();
    }
}

// Pass 2: every chunk derives its cell's decision from the (read-only) old state and the reduced rel-diff --
// stop rule and niter (em.rs:212-218, :181) -- and applies it to its ids: swap + zero; on a stop also zero the
// small counts (em.rs:238-242) so that the next sweep is the final one; after that sweep the counts are the
// cell's result.  The first chunk of a cell writes the new state (double buffered, so nothing races).
__global__ void __launch_bounds__(256) cells_apply(double *__restrict__ prev, double *__restrict__ curr,
                                                   const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t n_cells,
                                                   uint32_t max_iter, double thr, uint32_t min_iter,
                                                   const CellState *__restrict__ cs_old, CellState *__restrict__ cs_new,
                                                   unsigned long long *__restrict__ rel_bits, double *__restrict__ result,
                                                   uint32_t *__restrict__ done_count, OarEmState *st)
{
    if (st->done) return;
    for (uint32_t k = blockIdx.x; k < n_chunks; k += gridDim.x) {
        const Chunk ch = chunks[k];
        const CellState old = cs_old[ch.cell];
        CellState neu = old;
        if (old.phase == kDone) {
            if (ch.first && threadIdx.x == 0) cs_new[ch.cell] = neu;
            continue;                                                   // buffers are zero: the sweep adds nothing
        }
        if (old.phase == kFinal) {
            for (uint64_t i = ch.begin + threadIdx.x; i < ch.end; i += blockDim.x) { result[i] = curr[i]; curr[i] = 0.0; prev[i] = 0.0; }
            neu.phase = kDone;
            if (ch.first && threadIdx.x == 0) {
                cs_new[ch.cell] = neu;
                atomicAdd(done_count, 1u);   // published as st->done by cells_clear_rel: CTAs of THIS launch still test st->done
            }
            continue;
        }
        const double rel = __longlong_as_double((long long)rel_bits[ch.cell]);
        bool stop = rel < thr && old.niter > min_iter;                   // em.rs:212
        if (!stop) { neu.niter = old.niter + 1; stop = neu.niter >= max_iter; }   // em.rs:218, :181
        if (stop) neu.phase = kFinal;
        for (uint64_t i = ch.begin + threadIdx.x; i < ch.end; i += blockDim.x) {
            prev[i] = 0.0;                                               // swap + fill(0)
            if (stop && curr[i] < OAR_MIN_READ_THRESH) curr[i] = 0.0;    // em.rs:238-242 on the new prev
        }
        if (ch.first && threadIdx.x == 0) cs_new[ch.cell] = neu;
    }
}

// The tiles the next graph launch sweeps: those with a cell that has not delivered its result yet.  One CTA walks
// the tiles in order (the list stays ascending: neighbouring tiles keep sharing transcripts in L2).
__global__ void __launch_bounds__(1024) cells_active_tiles(const uint2 *__restrict__ ranges, uint32_t n_tiles,
                                                           const CellState *__restrict__ cs, uint32_t *__restrict__ list,
                                                           uint32_t *__restrict__ n_active, const OarEmState *st)
{
    if (st->done) return;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_run;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += blockDim.x) {
        const uint32_t t = base + threadIdx.x;
        bool active = false;
        if (t < n_tiles) {
            const uint2 r = ranges[t];
            for (uint32_t c = r.x; c <= r.y && !active && r.x != 0xFFFFFFFFu; ++c) active = cs[c].phase != kDone;
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) { const uint32_t v = s_warp[w]; if (w < warp) before += v; total += v; }
        const uint32_t run = s_run;
        if (active) list[run + before + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_active = s_run;
}

// rel_bits must be zero before the next reduce; done in its own tiny pass so that no chunk of cells_apply can
// still be reading it.
// Also publishes the end of the batch: once every cell has delivered its result (done_count, incremented by
// cells_apply) st->done is set HERE, in the launch after cells_apply -- set inside cells_apply, a CTA of the same
// launch that starts later would return at its `if (st->done)` test before copying its chunk of a result.
__global__ void cells_clear_rel(unsigned long long *__restrict__ rel_bits, uint32_t n_cells, const uint32_t *__restrict__ done_count,
                                OarEmState *st)
{
    if (st->done) return;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cells) rel_bits[c] = 0ull;
    if (c == 0 && *done_count == n_cells) st->done = 1;
}

}  // namespace cells
}  // namespace oar

using namespace oar;

extern "C" int oar_em_batched(oar_store *s, const uint64_t *cell_row_ptr, uint32_t n_cells, uint32_t max_iter,
                              double conv_thresh, uint32_t min_iter, uint64_t *out_cell_ptr, uint32_t *out_txp,
                              double *out_val, uint64_t capacity, uint64_t *out_nnz, uint32_t *out_niter)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_em_batched: store is null");
    if (!cell_row_ptr || !out_cell_ptr || !out_nnz) return fail(OAR_ERR_INVALID, "oar_em_batched: null argument");
    OAR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    s->counters[0] = s->counters[1] = 0;
    // stream-ordered allocations from the device's pool (kept warm by the device context): no cudaMalloc / cudaFree per call
    struct Scratch { cudaStream_t st; std::vector<void *> p; ~Scratch() { for (void *q : p) dfree(q, st); } } sc{st, {}};
    auto dalloc = [&](void **ptr, size_t bytes) -> cudaError_t {
        cudaError_t e = dmalloc(ptr, std::max<size_t>(bytes, 16), st); if (e == cudaSuccess) sc.p.push_back(*ptr); return e; };
    uint64_t *d_rows = nullptr, *d_cd = nullptr;
    OAR_CUDA(dalloc((void **)&d_rows, sizeof(uint64_t) * ((size_t)n_cells + 1)));
    OAR_CUDA(dalloc((void **)&d_cd, sizeof(uint64_t) * ((size_t)n_cells + 1)));
    OAR_CUDA(cudaMemcpyAsync(d_rows, cell_row_ptr, sizeof(uint64_t) * ((size_t)n_cells + 1), cudaMemcpyDefault, st));
    // validate the partition on the host copy
    std::vector<uint64_t> h_rows((size_t)n_cells + 1);
    OAR_CUDA(cudaMemcpyAsync(h_rows.data(), d_rows, sizeof(uint64_t) * h_rows.size(), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    if (h_rows[0] != 0 || h_rows[n_cells] != s->n_reads) return fail(OAR_ERR_INVALID, "oar_em_batched: cell_row_ptr must span [0, n_reads]");
    for (uint32_t c = 0; c < n_cells; ++c) if (h_rows[c + 1] < h_rows[c]) return fail(OAR_ERR_INVALID, "oar_em_batched: cell_row_ptr is not monotone");
    OAR_CUDA(cudaEventRecord(s->ev[0], st));

    const uint32_t words = (s->n_txps + 31u) / 32u;
    const size_t smem_need = sizeof(uint32_t) * 2 * (size_t)words;
    const int use_smem = smem_need <= 200 * 1024;
    const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(n_cells, (uint32_t)s->sm_count * 2));
    uint32_t *d_scratch = nullptr;
    if (!use_smem) OAR_CUDA(dalloc((void **)&d_scratch, sizeof(uint32_t) * 2 * (size_t)words * grid));
    if (use_smem) {
        OAR_CUDA(cudaFuncSetAttribute(cells::cell_localize<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
        OAR_CUDA(cudaFuncSetAttribute(cells::cell_localize<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
    }
    const size_t dyn = use_smem ? smem_need : 0;
    OAR_CUDA(cudaMemsetAsync(d_cd, 0, sizeof(uint64_t) * ((size_t)n_cells + 1), st));
    if (n_cells > 0) {
        cells::cell_localize<true><<<grid, cells::kThreads, dyn, st>>>(s->d_row_ptr, s->d_txp, d_rows, n_cells, s->n_txps, words,
                                                                       d_scratch, d_cd, nullptr, nullptr, use_smem);
        OAR_CUDA(cudaGetLastError());
    }
    // offsets on the host (n_cells is small compared to the data)
    std::vector<uint64_t> h_cd((size_t)n_cells + 1);
    OAR_CUDA(cudaMemcpyAsync(h_cd.data(), d_cd, sizeof(uint64_t) * h_cd.size(), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    for (uint32_t c = 0; c < n_cells; ++c) h_cd[c + 1] += h_cd[c];
    const uint64_t total = h_cd[n_cells];
    *out_nnz = total;
    OAR_CUDA(cudaMemcpyAsync(out_cell_ptr, h_cd.data(), sizeof(uint64_t) * h_cd.size(), cudaMemcpyDefault, st));
    if (total > capacity || (total > 0 && (!out_txp || !out_val))) {
        OAR_CUDA(cudaStreamSynchronize(st));
        return fail(OAR_ERR_INVALID, "oar_em_batched: output capacity too small (required size returned in out_nnz)");
    }
    OAR_CUDA(cudaMemcpyAsync(d_cd, h_cd.data(), sizeof(uint64_t) * h_cd.size(), cudaMemcpyHostToDevice, st));
    uint32_t *d_ctx = nullptr, *d_lid = nullptr, *d_niter = nullptr;
    double *d_a = nullptr, *d_b = nullptr, *d_val = nullptr;
    OAR_CUDA(dalloc((void **)&d_ctx, sizeof(uint32_t) * total));
    OAR_CUDA(dalloc((void **)&d_lid, sizeof(uint32_t) * s->nnz));
    OAR_CUDA(dalloc((void **)&d_a, sizeof(double) * total));
    OAR_CUDA(dalloc((void **)&d_b, sizeof(double) * total));
    OAR_CUDA(dalloc((void **)&d_val, sizeof(double) * total));
    OAR_CUDA(cudaMemsetAsync(d_val, 0, sizeof(double) * std::max<uint64_t>(total, 2), st));
    OAR_CUDA(dalloc((void **)&d_niter, sizeof(uint32_t) * std::max<uint32_t>(n_cells, 1)));
    if (n_cells > 0) {
        cells::cell_localize<false><<<grid, cells::kThreads, dyn, st>>>(s->d_row_ptr, s->d_txp, d_rows, n_cells, s->n_txps, words,
                                                                        d_scratch, d_cd, d_ctx, d_lid, use_smem);
        OAR_CUDA(cudaGetLastError());
        // preferred: one block-diagonal EM over (cell, transcript) pairs on the tiled sweep
        const char *ct = getenv("OAR_CELLS_TILED");
        bool tiled_done = false;
        if (!(ct && ct[0] == '0') && total > 0 && total < (1ull << 28)) {
            oar_store *sub = nullptr;
            // d_lid belongs to the sub-store from here on (substore_create frees it when it fails)
            sc.p.erase(std::remove(sc.p.begin(), sc.p.end(), (void *)d_lid), sc.p.end());
            const bool trace = getenv("OAR_TRACE") != nullptr;   // development: wall-clock split on stderr
            auto wall = [&]() { cudaStreamSynchronize(st); timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
            const double t_a = trace ? wall() : 0.0;
            int rc = substore_create(s, d_lid, (uint32_t)total, &sub);
            if (rc != OAR_OK) return rc;
            const double t_b = trace ? wall() : 0.0;
            struct SubGuard { oar_store *p; ~SubGuard() { oar_store_destroy(p); } } sg{sub};
            // chunks of at most 4096 ids, never crossing a cell
            std::vector<cells::Chunk> h_chunks;
            for (uint32_t c = 0; c < n_cells; ++c) {
                uint64_t b = h_cd[c]; const uint64_t e = h_cd[c + 1];
                bool first = true;
                do {
                    const uint64_t e2 = std::min<uint64_t>(e, b + 4096);
                    h_chunks.push_back(cells::Chunk{c, first ? 1u : 0u, b, e2});
                    first = false; b = e2;
                } while (b < e);
            }
            const uint32_t n_chunks = (uint32_t)h_chunks.size();
            cells::CellState *d_cs = nullptr; uint32_t *d_done = nullptr; cells::Chunk *d_chunks = nullptr;
            unsigned long long *d_rel = nullptr;
            OAR_CUDA(dalloc((void **)&d_cs, sizeof(cells::CellState) * 2 * n_cells));
            OAR_CUDA(dalloc((void **)&d_done, sizeof(uint32_t) * 4));
            OAR_CUDA(dalloc((void **)&d_chunks, sizeof(cells::Chunk) * n_chunks));
            OAR_CUDA(dalloc((void **)&d_rel, sizeof(unsigned long long) * n_cells));
            OAR_CUDA(cudaMemcpyAsync(d_chunks, h_chunks.data(), sizeof(cells::Chunk) * n_chunks, cudaMemcpyHostToDevice, st));
            OAR_CUDA(cudaMemsetAsync(d_done, 0, sizeof(uint32_t) * 4, st));
            OAR_CUDA(cudaMemsetAsync(d_rel, 0, sizeof(unsigned long long) * n_cells, st));
            OAR_CUDA(cudaMemsetAsync(sub->d_state, 0, sizeof(OarEmState), st));
            // tiles whose cells have all delivered are dropped from the sweep: per tile its range of cells (once), and at
            // the head of every graph launch the list of tiles that still have a live cell
            const uint32_t nt = sub->tl.n_tiles;
            uint2 *d_ranges = nullptr; uint32_t *d_list = nullptr, *d_nact = nullptr;
            OAR_CUDA(dalloc((void **)&d_ranges, sizeof(uint2) * std::max<uint32_t>(nt, 1)));
            OAR_CUDA(dalloc((void **)&d_list, sizeof(uint32_t) * std::max<uint32_t>(nt, 1)));
            OAR_CUDA(dalloc((void **)&d_nact, sizeof(uint32_t) * 4));
            OAR_CUDA(cudaMemsetAsync(d_nact, 0, sizeof(uint32_t) * 4, st));
            const bool listed = sub->kernel == OAR_KERNEL_TILED && nt > 0;
            if (listed) OAR_CUDA(tile_group_ranges_enqueue(sub, d_rows, n_cells, d_ranges));
            const int gridc = (int)std::max<uint32_t>(1, std::min<uint32_t>(n_cells, (uint32_t)s->sm_count * 16));
            const int gridk = (int)std::max<uint32_t>(1, std::min<uint32_t>(n_chunks, (uint32_t)s->sm_count * 8));
            cells::cells_init<<<gridc, 128, 0, st>>>(d_rows, d_cd, n_cells, s->n_txps, max_iter, d_a, d_b, d_cs);
            OAR_CUDA(cudaGetLastError());
            // a CUDA graph of 16 iterations: sweep a->b, update, sweep b->a, update, ...  (cell states ping-pong)
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            OAR_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = cudaSuccess;
            if (listed) {
                cells::cells_active_tiles<<<1, 1024, 0, st>>>(d_ranges, nt, d_cs, d_list, d_nact, sub->d_state);
                e = cudaGetLastError();
            }
            for (int it = 0; it < 16 && e == cudaSuccess; ++it) {
                double *pv = (it & 1) ? d_b : d_a, *cr = (it & 1) ? d_a : d_b;
                cells::CellState *cs_old = d_cs + (it & 1) * n_cells, *cs_new = d_cs + ((it + 1) & 1) * n_cells;
                e = listed ? sweep_enqueue_list(sub, pv, cr, sub->d_state, 1, d_list, d_nact) : sweep_enqueue(sub, pv, cr, sub->d_state, 1);
                if (e != cudaSuccess) break;
                cells::cells_reduce<<<gridk, 256, 0, st>>>(pv, cr, d_chunks, n_chunks, cs_old, d_rel, sub->d_state);
                cells::cells_apply<<<gridk, 256, 0, st>>>(pv, cr, d_chunks, n_chunks, n_cells, max_iter, conv_thresh, min_iter,
                                                         cs_old, cs_new, d_rel, d_val, d_done, sub->d_state);
                cells::cells_clear_rel<<<(n_cells + 255) / 256, 256, 0, st>>>(d_rel, n_cells, d_done, sub->d_state);
                e = cudaGetLastError();
            }
            cudaError_t e2 = cudaStreamEndCapture(st, &graph);
            if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(e, "cells graph capture"); }
            if (e2 != cudaSuccess) return cuda_fail(e2, "cudaStreamEndCapture");
            e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
            struct ExecGuard { cudaGraphExec_t x; ~ExecGuard() { cudaGraphExecDestroy(x); } } eg{exec};
            const double t_c = trace ? wall() : 0.0;
            std::vector<uint32_t> trace_active;
            // two graph launches in flight; the state copied after each is polled (launches after the end are no-ops)
            uint64_t launched = 0;
            int inflight = 0, head = 0;
            for (bool done = false; !done;) {
                while (inflight < 2) {
                    const int slot = (head + inflight) & 1;
                    OAR_CUDA(cudaGraphLaunch(exec, st));
                    launched += 64 + (listed ? 1 : 0);
                    OAR_CUDA(cudaMemcpyAsync(&sub->h_state[slot], sub->d_state, sizeof(OarEmState), cudaMemcpyDeviceToHost, st));
                    OAR_CUDA(cudaEventRecord(sub->slot_ev[slot], st));
                    ++inflight;
                }
                OAR_CUDA(cudaEventSynchronize(sub->slot_ev[head]));
                if (trace && listed) { uint32_t na = 0; cudaMemcpy(&na, d_nact, sizeof(na), cudaMemcpyDeviceToHost); trace_active.push_back(na); }
                done = sub->h_state[head].done != 0;
                head ^= 1; --inflight;
            }
            OAR_CUDA(cudaStreamSynchronize(st));
            if (trace) {
                const double t_d = wall();
                fprintf(stderr, "[oar] cells: sub-store layout %.1f ms (%u tiles), setup + graph %.1f ms, EM loop %.1f ms (%llu graph launches of 16 iterations, %u cells, %llu pairs)\n",
                        t_b - t_a, sub->tl.n_tiles, t_c - t_b, t_d - t_c, (unsigned long long)(launched / 65), n_cells, (unsigned long long)total);
                fprintf(stderr, "[oar] cells: tiles in the sweep after each graph launch:");
                for (size_t i = 0; i < trace_active.size(); i += std::max<size_t>(1, trace_active.size() / 16)) fprintf(stderr, " %u", trace_active[i]);
                fprintf(stderr, "\n");
            }
            s->counters[0] += launched + 4;
            // per-cell iteration counts
            std::vector<cells::CellState> h_cs(n_cells);
            // 16 iterations per graph launch: the live state is back in buffer 0 (a finished batch stops updating both)
            OAR_CUDA(cudaMemcpyAsync(h_cs.data(), d_cs, sizeof(cells::CellState) * n_cells, cudaMemcpyDeviceToHost, st));
            OAR_CUDA(cudaStreamSynchronize(st));
            std::vector<uint32_t> h_nit(n_cells);
            for (uint32_t c = 0; c < n_cells; ++c) h_nit[c] = h_cs[c].niter;
            OAR_CUDA(cudaMemcpyAsync(d_niter, h_nit.data(), sizeof(uint32_t) * n_cells, cudaMemcpyHostToDevice, st));
            OAR_CUDA(cudaStreamSynchronize(st));
            tiled_done = true;
        }
        if (!tiled_done) {
        const int grid2 = (int)std::max<uint32_t>(1, std::min<uint32_t>(n_cells, (uint32_t)s->sm_count * 4));
        if (s->d_aux)
            cells::cell_em<true><<<grid2, cells::kThreads, 0, st>>>(s->d_row_ptr, d_lid, s->d_prob, s->d_aux, d_rows, d_cd, n_cells,
                                                                    s->n_txps, d_a, d_b, max_iter, conv_thresh, min_iter, d_val, d_niter);
        else
            cells::cell_em<false><<<grid2, cells::kThreads, 0, st>>>(s->d_row_ptr, d_lid, s->d_prob, nullptr, d_rows, d_cd, n_cells,
                                                                     s->n_txps, d_a, d_b, max_iter, conv_thresh, min_iter, d_val, d_niter);
        OAR_CUDA(cudaGetLastError());
        s->counters[0] += 3;
        }
    }
    OAR_CUDA(cudaEventRecord(s->ev[1], st));
    if (total > 0) {
        OAR_CUDA(cudaMemcpyAsync(out_txp, d_ctx, sizeof(uint32_t) * total, cudaMemcpyDefault, st));
        OAR_CUDA(cudaMemcpyAsync(out_val, d_val, sizeof(double) * total, cudaMemcpyDefault, st));
    }
    if (out_niter && n_cells > 0) OAR_CUDA(cudaMemcpyAsync(out_niter, d_niter, sizeof(uint32_t) * n_cells, cudaMemcpyDefault, st));
    OAR_CUDA(cudaEventRecord(s->ev[2], st));
    OAR_CUDA(cudaStreamSynchronize(st));
    float a = 0.f, b = 0.f;
    OAR_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    OAR_CUDA(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    s->timings[1] = a; s->timings[2] = b; s->timings[3] = 0;
    return OAR_OK;
}
