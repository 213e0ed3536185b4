// oar_cov.cu -- the bulk coverage model (--model-coverage) on the device.
//
// Pre-EM stage of oarfish's bulk driver (bulk.rs:103-108): per-transcript coverage histograms built from
// the alignments (add_interval, src/util/oarfish_types.rs:496-537), the clamped logistic of each bin's
// relative deficit (logistic_prob, src/util/logistic_probability.rs:7-79) and, per alignment, the mean bin
// probability over the bins it starts in / covers, normalised per read (normalize_read_probs,
// src/util/normalize_probability.rs:5-74).  The result is the store's per-alignment f64 factor
// (coverage_probabilities == `aux`), produced and consumed in HBM.
#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "oar_store.cuh"

namespace oar {
namespace cov {

__global__ void n_bins(const uint32_t *__restrict__ txp_len, uint32_t n_txps, uint32_t bin_width, uint32_t *__restrict__ nb)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_txps) nb[t] = (uint32_t)ceil((double)txp_len[t] / (double)bin_width);   // oarfish_types.rs:460-468
    if (t == n_txps) nb[t] = 0;
}

// add_interval (oarfish_types.rs:496-537), weight 1.0 per alignment
__global__ void add_intervals(const uint32_t *__restrict__ txp, const uint32_t *__restrict__ start,
                              const uint32_t *__restrict__ end, uint64_t nnz, const uint32_t *__restrict__ txp_len,
                              const uint32_t *__restrict__ bin_off, double *__restrict__ bins, double *__restrict__ tw)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) {
        const uint32_t t = txp[j];
        const uint32_t o = bin_off[t], nI = bin_off[t + 1] - o;
        const double nIf = (double)nI, tlen = (double)txp_len[t];
        const double bw = round(tlen / nIf);
        const uint32_t st = min(start[j], end[j]);
        const uint32_t sp = max(st, end[j]);
        const uint32_t sb = (uint32_t)floor(((double)st / tlen) * nIf);
        const uint32_t eb = (uint32_t)floor(((double)sp / tlen) * nIf);
        for (uint32_t i = sb; i < eb && i < nI; ++i) {
            const double bf = (double)i;
            const uint32_t cbs = (uint32_t)(bf * bw);
            const uint32_t cbe = (uint32_t)fmin((bf + 1.0) * bw, tlen);
            uint32_t olap = 0;
            if (st <= cbe) olap = min(sp, cbe) - max(st, cbs);
            atomicAdd(bins + o + i, (double)olap / (double)(cbe - cbs));
        }
        atomicAdd(tw + t, 1.0);
    }
}

// logistic_prob (logistic_probability.rs:40-79)
__global__ void logistic(uint32_t n_txps, const uint32_t *__restrict__ bin_off, const double *__restrict__ bins,
                         const double *__restrict__ tw, double growth_rate, double *__restrict__ cov_prob)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_txps) return;
    const uint32_t o = bin_off[t], n = bin_off[t + 1] - o;
    const double min_cov = tw[t] / 100.0;
    double count_sum = 0.0;
    for (uint32_t i = 0; i < n; ++i) count_sum += (double)(float)(bins[o + i] + min_cov);
    if (count_sum <= 1e-8) { for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] = 0.0; return; }
    const double expected = count_sum / (double)n;
    for (uint32_t i = 0; i < n; ++i) {
        const double diff = (expected - (double)(float)(bins[o + i] + min_cov)) / expected;
        double r = 1.0 / (1.0 + exp(-growth_rate * diff));
        r = fmin(fmax(r, 1e-8), 0.99999);
        cov_prob[o + i] = r;
    }
}

// statrs 0.18 ln_gamma (Lanczos, g = 10.900511, 11 coefficients), branch x >= 0.5: the function the reference calls
// (binomial_probability.rs:4); restated so that the device follows the reference's arithmetic rather than CUDA's lgamma
__device__ __forceinline__ double statrs_ln_gamma(double x)
{
    const double dk[11] = {
        2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
        -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
        -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9 };
    double sum = dk[0];
#pragma unroll
    for (int k = 1; k < 11; ++k) sum += dk[k] / (x + (double)k - 1.0);
    return log(sum) + 0.6207822376352452223455184457816472122518527279025978 +
           (x - 0.5) * log((x - 0.5 + 10.900511) / 2.718281828459045235360287471352662497757);
}

// binomial_continuous_prob + binomial_probability (binomial_probability.rs:180-224, :7-178): the coverage model of the
// single-cell driver (single_cell.rs:132-137).  One thread per transcript, bins in order: the f32 sums of the
// reference (count_sum, sum_vec) keep their summation order.
__global__ void binomial(uint32_t n_txps, const uint32_t *__restrict__ txp_len, const uint32_t *__restrict__ bin_off,
                         const double *__restrict__ bins, const double *__restrict__ tw, double *__restrict__ cov_prob)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_txps) return;
    const uint32_t o = bin_off[t], n = bin_off[t + 1] - o;
    if (n == 0) return;
    const double min_cov = tw[t] / 100.0, zero_thresh = 1e-20, max_scale = 709.0;
    const float tlen = (float)(double)txp_len[t];
    const float bwf = (float)round((double)txp_len[t] / (double)n);
    float count_sum = 0.0f, max_count = nanf("");
    double rate = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        const float c = (float)(bins[o + i] + min_cov);
        const float len = fminf(((float)i + 1.0f) * bwf, tlen) - (float)i * bwf;
        rate += (double)c / (double)len;
        count_sum += c;
        max_count = fmaxf(max_count, c);
    }
    if (count_sum == 0.0f || rate == 0.0) { for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] = 0.0; return; }
    float sum_vec = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        const float c = (float)(bins[o + i] + min_cov);
        sum_vec += c == max_count ? (float)max_scale : (float)(((double)c * max_scale) / (double)max_count);
    }
    const double ln1 = statrs_ln_gamma((double)sum_vec + 1.0);
    double total = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        const float c = (float)(bins[o + i] + min_cov);
        const float len = fminf(((float)i + 1.0f) * bwf, tlen) - (float)i * bwf;
        const double prob = (c == 0.0f || len == 0.0f) ? 0.0 : (double)c / ((double)len * rate);
        const float m = c == max_count ? (float)max_scale : (float)(((double)c * max_scale) / (double)max_count);
        const float rest = sum_vec - m;
        const double den = statrs_ln_gamma((double)m + 1.0) + statrs_ln_gamma((double)rest + 1.0);
        const double num2 = (prob > zero_thresh ? log(prob) : log(zero_thresh)) * (double)m;
        const double num3 = ((1.0 - prob) > zero_thresh ? log(1.0 - prob) : log(zero_thresh)) * (double)rest;
        const double res = exp(ln1 - den + num2 + num3);
        cov_prob[o + i] = res; total += res;
    }
    for (uint32_t i = 0; i < n; ++i) cov_prob[o + i] /= total;
}

__device__ __forceinline__ double aln_cov(const double *__restrict__ cp, uint32_t nb, double sa, double ea, double tlen, double bl)
{   // normalize_probability.rs:19-60
    const uint32_t sb = (uint32_t)(sa / bl);
    const uint32_t eb = min((uint32_t)(ea / bl), nb - 1u);
    double tw = 0.0, cpv = 0.0;
    if (sb == eb) {
        const double w = (ea - sa) / bl; tw = w; cpv = w * cp[sb];
    } else {
        for (uint32_t i = sb; i < eb; ++i) {
            const double w = (i == sb) ? (fmin(bl * (double)sb + bl, tlen) - sa) / bl : 1.0;
            tw += w; cpv += w * cp[i];
        }
    }
    return cpv / tw;
}

// normalize_read_probs: 8-lane group per read
__global__ void __launch_bounds__(256) normalize(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp,
                                                 const uint32_t *__restrict__ start, const uint32_t *__restrict__ end,
                                                 uint64_t n_rows, const uint32_t *__restrict__ txp_len,
                                                 const uint32_t *__restrict__ bin_off, const double *__restrict__ cov_prob,
                                                 double bl, double *__restrict__ out)
{
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u;
    const unsigned gmask = 0xFFu << (lane & 24u);
    const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    for (uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < n_rows; row += ngroups) {
        const uint32_t s = row_ptr[row], e = row_ptr[row + 1];
        double nsum = 0.0;
        for (uint32_t j = s + sub; j < e; j += 8) {
            const uint32_t t = txp[j];
            const double v = aln_cov(cov_prob + bin_off[t], bin_off[t + 1] - bin_off[t], (double)start[j], (double)end[j],
                                     (double)txp_len[t], bl);
            out[j] = v; nsum += v;
        }
        nsum += __shfl_xor_sync(gmask, nsum, 1);
        nsum += __shfl_xor_sync(gmask, nsum, 2);
        nsum += __shfl_xor_sync(gmask, nsum, 4);
        const double d = nsum > 0.0 ? nsum : 1.0;
        for (uint32_t j = s + sub; j < e; j += 8) out[j] /= d;
    }
}

}  // namespace cov
}  // namespace oar

using namespace oar;

enum CovModel { kCovLogistic = 0, kCovBinomial = 1 };

static int coverage_model_impl(oar_store *s, const uint32_t *aln_start, const uint32_t *aln_end, const uint32_t *txp_len,
                               uint32_t bin_width, CovModel model, double growth_rate, double *out_aux_or_null)
{
    if (!s || !txp_len || (s->nnz && (!aln_start || !aln_end))) return fail(OAR_ERR_INVALID, "oar_store_coverage_model: null argument");
    if (bin_width == 0) return fail(OAR_ERR_UNSUPPORTED, "oar_store_coverage_model: bin width 0 is not implemented (logistic_probability.rs:55)");
    OAR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    OAR_CUDA(cudaStreamSynchronize(st));
    struct Scratch { cudaStream_t st; std::vector<void *> p; ~Scratch() { for (void *q : p) dfree(q, st); } } sc{st, {}};
    auto alloc = [&](auto **ptr, size_t bytes) -> cudaError_t {
        cudaError_t e = dmalloc(ptr, bytes, st); if (e == cudaSuccess) sc.p.push_back(*ptr); return e; };
    const uint32_t M = s->n_txps;
    uint32_t *d_start = nullptr, *d_end = nullptr, *d_len = nullptr, *d_nb = nullptr, *d_off = nullptr;
    OAR_CUDA(alloc(&d_start, sizeof(uint32_t) * std::max<uint64_t>(s->nnz, 1)));
    OAR_CUDA(alloc(&d_end, sizeof(uint32_t) * std::max<uint64_t>(s->nnz, 1)));
    OAR_CUDA(alloc(&d_len, sizeof(uint32_t) * M));
    OAR_CUDA(alloc(&d_nb, sizeof(uint32_t) * ((size_t)M + 1)));
    OAR_CUDA(alloc(&d_off, sizeof(uint32_t) * ((size_t)M + 1)));
    if (s->nnz) {
        OAR_CUDA(cudaMemcpyAsync(d_start, aln_start, sizeof(uint32_t) * s->nnz, cudaMemcpyDefault, st));
        OAR_CUDA(cudaMemcpyAsync(d_end, aln_end, sizeof(uint32_t) * s->nnz, cudaMemcpyDefault, st));
    }
    OAR_CUDA(cudaMemcpyAsync(d_len, txp_len, sizeof(uint32_t) * M, cudaMemcpyDefault, st));
    const int threads = 256;
    cov::n_bins<<<(M + 1 + threads - 1) / threads, threads, 0, st>>>(d_len, M, bin_width, d_nb);
    OAR_CUDA(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        OAR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_nb, d_off, (int)M + 1, st));
        char *tmp = nullptr;
        OAR_CUDA(alloc(&tmp, tmp_bytes));
        OAR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_nb, d_off, (int)M + 1, st));
    }
    uint32_t total_bins = 0;
    OAR_CUDA(cudaMemcpyAsync(&total_bins, d_off + M, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    double *d_bins = nullptr, *d_tw = nullptr, *d_cp = nullptr;
    OAR_CUDA(alloc(&d_bins, sizeof(double) * std::max<uint32_t>(total_bins, 1)));
    OAR_CUDA(alloc(&d_cp, sizeof(double) * std::max<uint32_t>(total_bins, 1)));
    OAR_CUDA(alloc(&d_tw, sizeof(double) * M));
    OAR_CUDA(cudaMemsetAsync(d_bins, 0, sizeof(double) * std::max<uint32_t>(total_bins, 1), st));
    OAR_CUDA(cudaMemsetAsync(d_tw, 0, sizeof(double) * M, st));
    if (s->nnz) {
        const int blocks = (int)std::min<uint64_t>((s->nnz + threads - 1) / threads, (uint64_t)s->sm_count * 16);
        cov::add_intervals<<<blocks, threads, 0, st>>>(s->d_txp, d_start, d_end, s->nnz, d_len, d_off, d_bins, d_tw);
        OAR_CUDA(cudaGetLastError());
    }
    if (model == kCovBinomial) cov::binomial<<<(M + threads - 1) / threads, threads, 0, st>>>(M, d_len, d_off, d_bins, d_tw, d_cp);
    else cov::logistic<<<(M + threads - 1) / threads, threads, 0, st>>>(M, d_off, d_bins, d_tw, growth_rate, d_cp);
    OAR_CUDA(cudaGetLastError());
    if (!s->d_aux) OAR_CUDA(dmalloc(&s->d_aux, sizeof(double) * (s->nnz + 16), st));
    if (s->n_reads) {
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((s->n_reads + 31) / 32, (uint64_t)s->sm_count * 8));
        cov::normalize<<<blocks, threads, 0, st>>>(s->d_row_ptr, s->d_txp, d_start, d_end, s->n_reads, d_len, d_off, d_cp,
                                                  (double)bin_width, s->d_aux);
        OAR_CUDA(cudaGetLastError());
    }
    s->counters[0] += 4;
    if (out_aux_or_null && s->nnz) OAR_CUDA(cudaMemcpyAsync(out_aux_or_null, s->d_aux, sizeof(double) * s->nnz, cudaMemcpyDefault, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    // the store now carries the coverage factor: rebuild the tiled copy (it embeds aux) and drop stale graphs
    for (auto &g : s->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (s->tl.ready) {
        const uint32_t span = s->tl.span;
        s->kernel = OAR_KERNEL_ROWGROUP;          // whatever happens below, the store stays usable through the CSR kernel
        const int rc = build_tiled_layout(s, span);
        if (rc == OAR_OK) s->kernel = OAR_KERNEL_TILED;
        else if (rc != OAR_ERR_UNSUPPORTED) return rc;
    }
    return OAR_OK;
}

extern "C" int oar_store_coverage_model(oar_store *s, const uint32_t *aln_start, const uint32_t *aln_end,
                                        const uint32_t *txp_len, uint32_t bin_width, double growth_rate,
                                        double *out_aux_or_null)
{
    return coverage_model_impl(s, aln_start, aln_end, txp_len, bin_width, kCovLogistic, growth_rate, out_aux_or_null);
}

extern "C" int oar_store_coverage_model_binomial(oar_store *s, const uint32_t *aln_start, const uint32_t *aln_end,
                                                 const uint32_t *txp_len, uint32_t bin_width, double *out_aux_or_null)
{
    return coverage_model_impl(s, aln_start, aln_end, txp_len, bin_width, kCovBinomial, 0.0, out_aux_or_null);
}
