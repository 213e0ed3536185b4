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
                lid[j] = pre[t >> 5] + __popc(bits[t >> 5] & ((1u << (t & 31)) - 1u));
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
        double *prev = bufA + d0, *curr = bufB + d0;
        const double avg = (double)(r1 - r0) / (double)n_txps;    // em.rs:154,165: N_cell / M (full transcriptome)
        for (uint32_t i = threadIdx.x; i < L; i += kThreads) { prev[i] = avg; curr[i] = 0.0; }
        __threadfence();
        __syncthreads();
        uint32_t niter = 0;
        while (niter < max_iter) {
            cell_sweep<HAS_AUX>(row_ptr, lid, prob, aux, r0, r1, prev, curr);
            __threadfence();
            __syncthreads();
            double m = 0.0;
            for (uint32_t i = threadIdx.x; i < L; i += kThreads) {
                const double pc = ld_cg(prev + i), cc = ld_cg(curr + i);
                if (pc > OAR_MIN_READ_THRESH) { const double rd = (cc - pc) / pc; m = rd > m ? rd : m; }   // em.rs:194-201
                prev[i] = 0.0;                                                                               // swap + fill(0)
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
            if (ld_cg(prev + i) < OAR_MIN_READ_THRESH) prev[i] = 0.0;   // em.rs:238-242
        __threadfence();
        __syncthreads();
        cell_sweep<HAS_AUX>(row_ptr, lid, prob, aux, r0, r1, prev, curr);   // em.rs:245-252
        __threadfence();
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < L; i += kThreads) out_val[d0 + i] = ld_cg(curr + i);
        if (threadIdx.x == 0) out_niter[c] = niter;
        __syncthreads();
    }
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
    struct Scratch { std::vector<void *> p; ~Scratch() { for (void *q : p) cudaFree(q); } } sc;
    auto dalloc = [&](void **ptr, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(ptr, std::max<size_t>(bytes, 16)); if (e == cudaSuccess) sc.p.push_back(*ptr); return e; };
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
    OAR_CUDA(dalloc((void **)&d_niter, sizeof(uint32_t) * std::max<uint32_t>(n_cells, 1)));
    if (n_cells > 0) {
        cells::cell_localize<false><<<grid, cells::kThreads, dyn, st>>>(s->d_row_ptr, s->d_txp, d_rows, n_cells, s->n_txps, words,
                                                                        d_scratch, d_cd, d_ctx, d_lid, use_smem);
        OAR_CUDA(cudaGetLastError());
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
