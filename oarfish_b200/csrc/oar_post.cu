// oar_post.cu -- the two cheap passes over the resident CSR that follow the EM in
// oarfish's bulk driver:
//   oar_posteriors  <- write_out_prob's inner loop (src/util/write_function.rs:283-332)
//   oar_aux_counts  <- aux_counts::get_aux_counts   (src/util/aux_counts.rs:23-50)
#include <algorithm>

#include "oar_store.cuh"

namespace oar {
namespace post {

// one 8-lane group per read; same traversal as em_sweep_rowgroup, E-step only.  A lane keeps its first alignment's
// weight in a register (reads of <= 8 alignments -- nearly all -- gather counts[] once); f64 division as in the
// reference (this is output formatting, not the EM loop: no reciprocal trick).
template <bool HAS_AUX>
__global__ void __launch_bounds__(256) posteriors(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp,
                                                  const float *__restrict__ prob, const double *__restrict__ aux,
                                                  const double *__restrict__ counts, uint64_t n_rows, double thresh,
                                                  double *__restrict__ out, uint32_t *__restrict__ kept)
{
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u;
    const unsigned gmask = 0xFFu << (lane & 24u);
    const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    auto weight = [&](uint32_t j) {
        double w = counts[txp[j]] * (double)prob[j];              // write_function.rs:286-291
        if (HAS_AUX) w *= aux[j];
        return w;
    };
    auto clamp01 = [](double np) { return np < 0.0 ? 0.0 : (np > 1.0 ? 1.0 : np); };   // NaN stays NaN (:307)
    for (uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < n_rows; row += ngroups) {
        const uint32_t s = row_ptr[row], e = row_ptr[row + 1];
        const uint32_t j0 = s + sub;
        const bool has0 = j0 < e;
        const double w0 = has0 ? weight(j0) : 0.0;
        double denom = w0;
        for (uint32_t j = j0 + 8; j < e; j += 8) denom += weight(j);
        denom += __shfl_xor_sync(gmask, denom, 1);
        denom += __shfl_xor_sync(gmask, denom, 2);
        denom += __shfl_xor_sync(gmask, denom, 4);
        const double np0 = has0 ? clamp01(w0 / denom) : 0.0;
        const bool keep0 = has0 && np0 >= thresh;                  // :309-313
        double denom2 = keep0 ? np0 : 0.0;
        uint32_t k = keep0 ? 1u : 0u;
        for (uint32_t j = j0 + 8; j < e; j += 8) {
            const double np = clamp01(weight(j) / denom);
            if (np >= thresh) { denom2 += np; ++k; }
        }
        denom2 += __shfl_xor_sync(gmask, denom2, 1);
        denom2 += __shfl_xor_sync(gmask, denom2, 2);
        denom2 += __shfl_xor_sync(gmask, denom2, 4);
        k += __shfl_xor_sync(gmask, k, 1);
        k += __shfl_xor_sync(gmask, k, 2);
        k += __shfl_xor_sync(gmask, k, 4);
        if (has0) out[j0] = keep0 ? np0 / denom2 : 0.0;            // :316-318
        for (uint32_t j = j0 + 8; j < e; j += 8) {
            const double np = clamp01(weight(j) / denom);
            out[j] = (np >= thresh) ? np / denom2 : 0.0;
        }
        if (kept && sub == 0) kept[row] = k;
    }
}

__global__ void aux_counts(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp, uint64_t n_rows,
                           uint32_t *__restrict__ unique, uint32_t *__restrict__ total)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        const bool is_unique = (e - s) == 1;                       // aux_counts.rs:34
        for (uint32_t j = s; j < e; ++j) {
            atomicAdd(total + txp[j], 1u);
            if (is_unique) atomicAdd(unique + txp[j], 1u);
        }
    }
}

}  // namespace post
}  // namespace oar

using namespace oar;

extern "C" int oar_posteriors(oar_store *s, const double *counts, double display_thresh, double *out_prob,
                              uint32_t *out_kept_or_null)
{
    if (!s || !counts || (s->nnz && !out_prob)) return fail(OAR_ERR_INVALID, "oar_posteriors: null argument");
    OAR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    double *d_counts = nullptr, *d_out = nullptr; uint32_t *d_kept = nullptr;
    // stream-ordered allocations from the device pool (no cudaMalloc / cudaFree per call)
    struct Guard { cudaStream_t st; void *a = nullptr, *b = nullptr, *c = nullptr; ~Guard() { dfree(a, st); dfree(b, st); dfree(c, st); } } g{st};
    OAR_CUDA(dmalloc(&d_counts, sizeof(double) * s->n_txps, st)); g.a = d_counts;
    OAR_CUDA(dmalloc(&d_out, sizeof(double) * std::max<uint64_t>(s->nnz, 1), st)); g.b = d_out;
    if (out_kept_or_null) { OAR_CUDA(dmalloc(&d_kept, sizeof(uint32_t) * std::max<uint64_t>(s->n_reads, 1), st)); g.c = d_kept; }
    OAR_CUDA(cudaMemcpyAsync(d_counts, counts, sizeof(double) * s->n_txps, cudaMemcpyDefault, st));
    OAR_CUDA(cudaEventRecord(s->ev[0], st));
    if (s->n_reads) {
        const int threads = 256;
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((s->n_reads + 31) / 32, (uint64_t)s->sm_count * 8));
        if (s->d_aux) post::posteriors<true><<<blocks, threads, 0, st>>>(s->d_row_ptr, s->d_txp, s->d_prob, s->d_aux, d_counts, s->n_reads, display_thresh, d_out, d_kept);
        else post::posteriors<false><<<blocks, threads, 0, st>>>(s->d_row_ptr, s->d_txp, s->d_prob, nullptr, d_counts, s->n_reads, display_thresh, d_out, d_kept);
        OAR_CUDA(cudaGetLastError());
        s->counters[0] += 1;
    }
    OAR_CUDA(cudaEventRecord(s->ev[1], st));
    if (s->nnz) OAR_CUDA(cudaMemcpyAsync(out_prob, d_out, sizeof(double) * s->nnz, cudaMemcpyDefault, st));
    if (out_kept_or_null && s->n_reads) OAR_CUDA(cudaMemcpyAsync(out_kept_or_null, d_kept, sizeof(uint32_t) * s->n_reads, cudaMemcpyDefault, st));
    OAR_CUDA(cudaEventRecord(s->ev[2], st));
    OAR_CUDA(cudaStreamSynchronize(st));
    float a = 0.f, b = 0.f;
    OAR_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    OAR_CUDA(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    s->timings[1] = a; s->timings[2] = b;   // kernel, download (oar_store_timings)
    return OAR_OK;
}

extern "C" int oar_aux_counts(oar_store *s, uint32_t *out_unique, uint32_t *out_total)
{
    if (!s || !out_unique || !out_total) return fail(OAR_ERR_INVALID, "oar_aux_counts: null argument");
    OAR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    uint32_t *d = nullptr;
    OAR_CUDA(dmalloc(&d, sizeof(uint32_t) * 2 * (size_t)s->n_txps, st));
    struct Guard { cudaStream_t st; uint32_t *p; ~Guard() { dfree(p, st); } } g{st, d};
    OAR_CUDA(cudaMemsetAsync(d, 0, sizeof(uint32_t) * 2 * (size_t)s->n_txps, st));
    OAR_CUDA(cudaEventRecord(s->ev[0], st));
    if (s->n_reads) {
        const int threads = 256;
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((s->n_reads + threads - 1) / threads, (uint64_t)s->sm_count * 16));
        post::aux_counts<<<blocks, threads, 0, st>>>(s->d_row_ptr, s->d_txp, s->n_reads, d, d + s->n_txps);
        OAR_CUDA(cudaGetLastError());
        s->counters[0] += 1;
    }
    OAR_CUDA(cudaEventRecord(s->ev[1], st));
    OAR_CUDA(cudaMemcpyAsync(out_unique, d, sizeof(uint32_t) * s->n_txps, cudaMemcpyDefault, st));
    OAR_CUDA(cudaMemcpyAsync(out_total, d + s->n_txps, sizeof(uint32_t) * s->n_txps, cudaMemcpyDefault, st));
    OAR_CUDA(cudaEventRecord(s->ev[2], st));
    OAR_CUDA(cudaStreamSynchronize(st));
    float a = 0.f, b = 0.f;
    OAR_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    OAR_CUDA(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    s->timings[1] = a; s->timings[2] = b;
    return OAR_OK;
}
