// oar_ingest.cu -- the alignment filters and the score -> probability formula on the device (SURVEY.md section 8 f-4,
// the half of it that is data-parallel): AlignmentFilters::filter (src/util/oarfish_types.rs:955-1130), applied to
// every read's group of alignment records at once, writes the store the EM runs on straight into HBM.
//
// What stays in Rust: BGZF / BAM decoding and grouping records by read name (alignment_parser.rs:301-437); the caller
// hands over the records as columns (one entry per record, groups delimited by group_ptr) -- the fields filter() reads
// through AlnRecordLike (oarfish_types.rs:186-202).
//
// One thread walks one read group, exactly in the order filter() does: orientation, supplementary, aligned length,
// 3' and 5' clipping (each with its discard counter), best retained score (first maximum), the empty / zero-length /
// non-positive-score rejections, the aligned-fraction test, then per alignment the score-threshold test in f32 and
// prob = exp((score - best) / score_prob_denom) in f32.  Two passes (count, scan, write) produce the CSR.
#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "oar_store.cuh"

namespace oar {
namespace ingest {

enum : int { kDisc5p = 0, kDisc3p, kDiscScore, kDiscAlnFrac, kDiscAlnLen, kDiscOri, kDiscSupp, kNoMapping, kNoValidAln, kValidBestAln, kBadRefId, kCounters };

struct Records {
    const uint64_t *group_ptr;
    const uint32_t *ref_id, *aln_start, *aln_end, *aln_span, *seq_len;
    const int32_t *score;
    const uint8_t *flags;
    const uint32_t *txp_len;
    uint32_t n_txps;
};

// exp() of the reference is f32 (Rust f32::exp -> the platform's expf, correctly rounded in glibc for all practical
// inputs): evaluate in f64 and round once
__device__ __forceinline__ float exp_f32(float x) { return (float)exp((double)x); }

// Walks one group.  WRITE = false: returns the number of retained alignments (0: the read is dropped) and adds to the
// discard counters; WRITE = true: writes txp / prob / source index of the retained alignments at `out`.
template <bool WRITE>
__device__ __forceinline__ uint32_t filter_group(const Records &r, const oar_filter_opts &f, uint64_t g, unsigned long long *counters,
                                                 uint32_t *o_txp, float *o_prob, uint32_t *o_src, uint64_t out)
{
    const uint64_t b = r.group_ptr[g], e = r.group_ptr[g + 1];
    uint32_t seq_len = 0, n_mapped = 0;
    for (uint64_t j = b; j < e; ++j) {
        if (seq_len == 0 && r.seq_len[j] != 0) seq_len = r.seq_len[j];      // first record that carries the sequence (:981-984)
        if (!(r.flags[j] & OAR_REC_UNMAPPED)) ++n_mapped;                   // :976
    }
    int32_t best = INT32_MIN;
    float frac_at_best = 0.f;
    uint32_t len_at_best = 0, kept = 0;
    uint32_t c[kCounters] = {0};
    auto retained = [&](uint64_t j) -> bool {                               // the closure of ag.retain (:987-1062)
        const uint8_t fl = r.flags[j];
        if (fl & OAR_REC_UNMAPPED) return false;
        const bool rc = (fl & OAR_REC_REVERSE) != 0;
        if ((f.which_strand == OAR_STRAND_FORWARD && rc) || (f.which_strand == OAR_STRAND_REVERSE && !rc)) { ++c[kDiscOri]; return false; }
        if (fl & OAR_REC_SUPPLEMENTARY) { ++c[kDiscSupp]; return false; }
        if (r.aln_span[j] < f.min_aligned_len) { ++c[kDiscAlnLen]; return false; }
        const uint32_t t = r.ref_id[j];
        if (t >= r.n_txps) { ++c[kBadRefId]; return false; }
        if ((long long)r.aln_end[j] <= (long long)r.txp_len[t] - (long long)f.three_prime_clip) { ++c[kDisc3p]; return false; }
        if (r.aln_start[j] >= f.five_prime_clip) { ++c[kDisc5p]; return false; }
        return true;
    };
    for (uint64_t j = b; j < e; ++j) {
        if (!retained(j)) continue;
        ++kept;
        if (r.score[j] > best) {                                            // :1047-1058
            best = r.score[j];
            len_at_best = r.aln_span[j];
            frac_at_best = seq_len > 0 ? (float)r.aln_span[j] / (float)seq_len : 0.f;
        }
    }
    uint32_t n_out = 0;
    bool valid = true;
    if (kept == 0 || len_at_best == 0 || best <= 0) {                       // :1064-1076
        if (n_mapped == 0) ++c[kNoMapping]; else ++c[kNoValidAln];
        valid = false;
    } else if (frac_at_best < f.min_aligned_fraction) {                     // :1077-1082
        ++c[kDiscAlnFrac];
        valid = false;
    }
    if (valid) {
        ++c[kValidBestAln];
        const float mscore = (float)best, inv_max = 1.0f / mscore;         // :1088-1089
        for (uint64_t j = b; j < e; ++j) {
            // the second walk must not count the per-record discards again: test without the counters
            const uint8_t fl = r.flags[j];
            if (fl & (OAR_REC_UNMAPPED | OAR_REC_SUPPLEMENTARY)) continue;
            const bool rc = (fl & OAR_REC_REVERSE) != 0;
            if ((f.which_strand == OAR_STRAND_FORWARD && rc) || (f.which_strand == OAR_STRAND_REVERSE && !rc)) continue;
            if (r.aln_span[j] < f.min_aligned_len) continue;
            const uint32_t t = r.ref_id[j];
            if (t >= r.n_txps) continue;
            if ((long long)r.aln_end[j] <= (long long)r.txp_len[t] - (long long)f.three_prime_clip) continue;
            if (r.aln_start[j] >= f.five_prime_clip) continue;
            const float fs = (float)r.score[j];
            if (fs * inv_max >= f.score_threshold) {                        // :1100-1104
                if (WRITE) {
                    o_txp[out + n_out] = t;
                    o_prob[out + n_out] = exp_f32((fs - mscore) / f.score_prob_denom);
                    if (o_src) o_src[out + n_out] = (uint32_t)j;
                }
                ++n_out;
            } else ++c[kDiscScore];
        }
    }
    if (!WRITE)
        for (int k = 0; k < kCounters; ++k) if (c[k]) atomicAdd(counters + k, (unsigned long long)c[k]);
    return n_out;
}

__global__ void __launch_bounds__(256) count_groups(Records r, oar_filter_opts f, uint64_t n_groups, uint32_t *__restrict__ n_kept,
                                                    uint32_t *__restrict__ row_kept, unsigned long long *__restrict__ counters)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint32_t n = filter_group<false>(r, f, g, counters, nullptr, nullptr, nullptr, 0);
        n_kept[g] = n;
        row_kept[g] = n ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) write_groups(Records r, oar_filter_opts f, uint64_t n_groups, const uint32_t *__restrict__ aln_off,
                                                    const uint32_t *__restrict__ row_off, uint32_t *__restrict__ row_ptr,
                                                    uint32_t *__restrict__ o_txp, float *__restrict__ o_prob, uint32_t *__restrict__ o_src,
                                                    uint32_t *__restrict__ o_group, uint32_t n_rows, uint32_t nnz)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint32_t a0 = aln_off[g], a1 = aln_off[g + 1];
        if (a1 == a0) continue;                                             // add_filtered_group drops empty groups (:724)
        const uint32_t row = row_off[g];
        row_ptr[row] = a0;
        if (o_group) o_group[row] = (uint32_t)g;
        filter_group<true>(r, f, g, nullptr, o_txp, o_prob, o_src, a0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) row_ptr[n_rows] = nnz;
}

}  // namespace ingest
}  // namespace oar

using namespace oar;

extern "C" int oar_store_create_filtered(const uint64_t *group_ptr, const uint32_t *ref_id, const uint32_t *aln_start,
                                         const uint32_t *aln_end, const uint32_t *aln_span, const int32_t *score,
                                         const uint8_t *flags, const uint32_t *seq_len, uint64_t n_groups, uint64_t n_records,
                                         const uint32_t *txp_len, uint32_t n_txps, const oar_filter_opts *opts, int device,
                                         oar_store **out, uint64_t out_discard[10], uint32_t *out_src_or_null,
                                         uint32_t *out_group_or_null)
{
    if (!out) return fail(OAR_ERR_INVALID, "oar_store_create_filtered: out is null");
    *out = nullptr;
    if (!group_ptr || !txp_len || !opts || (n_records && (!ref_id || !aln_start || !aln_end || !aln_span || !score || !flags || !seq_len)))
        return fail(OAR_ERR_INVALID, "oar_store_create_filtered: null argument");
    if (n_records >= 0xFFFFFFF0ull || n_groups >= 0xFFFFFFF0ull)
        return fail(OAR_ERR_UNSUPPORTED, "oar_store_create_filtered: >= 2^32 records or groups are not supported");
    oar_store *s = nullptr;
    int rc = new_store(device, 0, 0, n_txps, "oar_store_create_filtered", &s);
    if (rc != OAR_OK) return rc;
    rc = [&]() -> int {
        cudaStream_t st = s->stream;
        struct Scratch { cudaStream_t st; std::vector<void *> p; ~Scratch() { for (void *q : p) dfree(q, st); } } sc{st, {}};
        auto up = [&](auto **dst, const void *src, size_t bytes) -> cudaError_t {
            cudaError_t e = dmalloc(dst, std::max<size_t>(bytes, 16), st);
            if (e != cudaSuccess) return e;
            sc.p.push_back(*dst);
            return bytes ? cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyDefault, st) : cudaSuccess;
        };
        OAR_CUDA(cudaEventRecord(s->ev[0], st));
        ingest::Records r;
        uint64_t *d_gp; uint32_t *d_ref, *d_start, *d_end, *d_span, *d_seq, *d_len; int32_t *d_score; uint8_t *d_flags;
        OAR_CUDA(up(&d_gp, group_ptr, sizeof(uint64_t) * (n_groups + 1)));
        OAR_CUDA(up(&d_ref, ref_id, sizeof(uint32_t) * n_records));
        OAR_CUDA(up(&d_start, aln_start, sizeof(uint32_t) * n_records));
        OAR_CUDA(up(&d_end, aln_end, sizeof(uint32_t) * n_records));
        OAR_CUDA(up(&d_span, aln_span, sizeof(uint32_t) * n_records));
        OAR_CUDA(up(&d_seq, seq_len, sizeof(uint32_t) * n_records));
        OAR_CUDA(up(&d_score, score, sizeof(int32_t) * n_records));
        OAR_CUDA(up(&d_flags, flags, sizeof(uint8_t) * n_records));
        OAR_CUDA(up(&d_len, txp_len, sizeof(uint32_t) * n_txps));
        r.group_ptr = d_gp; r.ref_id = d_ref; r.aln_start = d_start; r.aln_end = d_end; r.aln_span = d_span; r.seq_len = d_seq;
        r.score = d_score; r.flags = d_flags; r.txp_len = d_len; r.n_txps = n_txps;
        uint32_t *d_nk = nullptr, *d_rk = nullptr, *d_aoff = nullptr, *d_roff = nullptr;
        unsigned long long *d_cnt = nullptr;
        OAR_CUDA(dmalloc(&d_nk, sizeof(uint32_t) * (n_groups + 1), st)); sc.p.push_back(d_nk);
        OAR_CUDA(dmalloc(&d_rk, sizeof(uint32_t) * (n_groups + 1), st)); sc.p.push_back(d_rk);
        OAR_CUDA(dmalloc(&d_aoff, sizeof(uint32_t) * (n_groups + 1), st)); sc.p.push_back(d_aoff);
        OAR_CUDA(dmalloc(&d_roff, sizeof(uint32_t) * (n_groups + 1), st)); sc.p.push_back(d_roff);
        OAR_CUDA(dmalloc(&d_cnt, sizeof(unsigned long long) * ingest::kCounters, st)); sc.p.push_back(d_cnt);
        OAR_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * ingest::kCounters, st));
        OAR_CUDA(cudaMemsetAsync(d_nk + n_groups, 0, sizeof(uint32_t), st));
        OAR_CUDA(cudaMemsetAsync(d_rk + n_groups, 0, sizeof(uint32_t), st));
        const int threads = 256;
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_groups + threads - 1) / threads, (uint64_t)s->sm_count * 16));
        if (n_groups) {
            ingest::count_groups<<<blocks, threads, 0, st>>>(r, *opts, n_groups, d_nk, d_rk, d_cnt);
            OAR_CUDA(cudaGetLastError());
        }
        {
            size_t tmp_bytes = 0;
            OAR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_nk, d_aoff, (int)n_groups + 1, st));
            void *tmp = nullptr;
            OAR_CUDA(dmalloc((char **)&tmp, tmp_bytes, st)); sc.p.push_back(tmp);
            OAR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_nk, d_aoff, (int)n_groups + 1, st));
            OAR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_rk, d_roff, (int)n_groups + 1, st));
        }
        uint32_t h_tot[2] = {0, 0};
        unsigned long long h_cnt[ingest::kCounters];
        OAR_CUDA(cudaMemcpyAsync(&h_tot[0], d_aoff + n_groups, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        OAR_CUDA(cudaMemcpyAsync(&h_tot[1], d_roff + n_groups, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        OAR_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        OAR_CUDA(cudaStreamSynchronize(st));
        if (h_cnt[ingest::kBadRefId]) return fail(OAR_ERR_INVALID, "oar_store_create_filtered: ref_id out of range (>= n_txps)");
        const uint32_t nnz = h_tot[0], n_rows = h_tot[1];
        s->n_reads = n_rows; s->nnz = nnz;
        const size_t pad = 16;
        OAR_CUDA(dmalloc(&s->d_row_ptr, sizeof(uint32_t) * ((size_t)n_rows + 1 + pad), st));
        OAR_CUDA(dmalloc(&s->d_txp, sizeof(uint32_t) * ((size_t)nnz + pad), st));
        OAR_CUDA(dmalloc(&s->d_prob, sizeof(float) * ((size_t)nnz + pad), st));
        OAR_CUDA(cudaMemsetAsync(s->d_txp + nnz, 0, sizeof(uint32_t) * pad, st));
        OAR_CUDA(cudaMemsetAsync(s->d_prob + nnz, 0, sizeof(float) * pad, st));
        OAR_CUDA(cudaMemsetAsync(s->d_state, 0, sizeof(OarEmState) * 2, st));
        uint32_t *d_src = nullptr, *d_grp = nullptr;
        if (out_src_or_null) { OAR_CUDA(dmalloc(&d_src, sizeof(uint32_t) * std::max<uint32_t>(nnz, 1), st)); sc.p.push_back(d_src); }
        if (out_group_or_null) { OAR_CUDA(dmalloc(&d_grp, sizeof(uint32_t) * std::max<uint32_t>(n_rows, 1), st)); sc.p.push_back(d_grp); }
        ingest::write_groups<<<blocks, threads, 0, st>>>(r, *opts, n_groups, d_aoff, d_roff, s->d_row_ptr, s->d_txp, s->d_prob, d_src, d_grp, n_rows, nnz);
        OAR_CUDA(cudaGetLastError());
        if (out_src_or_null && nnz) OAR_CUDA(cudaMemcpyAsync(out_src_or_null, d_src, sizeof(uint32_t) * nnz, cudaMemcpyDefault, st));
        if (out_group_or_null && n_rows) OAR_CUDA(cudaMemcpyAsync(out_group_or_null, d_grp, sizeof(uint32_t) * n_rows, cudaMemcpyDefault, st));
        if (out_discard) for (int k = 0; k < 10; ++k) out_discard[k] = h_cnt[k];
        OAR_CUDA(cudaStreamSynchronize(st));
        int rc2 = finish_store(s);
        if (rc2 != OAR_OK) return rc2;
        OAR_CUDA(cudaEventRecord(s->ev[1], st));
        OAR_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        OAR_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        s->timings[0] = ms;
        return OAR_OK;
    }();
    if (rc != OAR_OK) { std::string keep = oar_last_error(); oar_store_destroy(s); return fail(rc, keep); }
    *out = s;
    return OAR_OK;
}

/* The CSR a store holds, for callers that built it on the device (oar_store_create_filtered): row_ptr as u64 like
 * InMemoryAlignmentStore.boundaries, txp_id, prob; any of the three may be NULL. */
extern "C" int oar_store_export(oar_store *s, uint64_t *out_row_ptr, uint32_t *out_txp_id, float *out_prob)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_store_export: store is null");
    OAR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    if (out_row_ptr) {
        std::vector<uint32_t> h((size_t)s->n_reads + 1);
        OAR_CUDA(cudaMemcpyAsync(h.data(), s->d_row_ptr, sizeof(uint32_t) * h.size(), cudaMemcpyDeviceToHost, st));
        OAR_CUDA(cudaStreamSynchronize(st));
        std::vector<uint64_t> w(h.begin(), h.end());
        OAR_CUDA(cudaMemcpyAsync(out_row_ptr, w.data(), sizeof(uint64_t) * w.size(), cudaMemcpyDefault, st));
        OAR_CUDA(cudaStreamSynchronize(st));
    }
    if (out_txp_id && s->nnz) OAR_CUDA(cudaMemcpyAsync(out_txp_id, s->d_txp, sizeof(uint32_t) * s->nnz, cudaMemcpyDefault, st));
    if (out_prob && s->nnz) OAR_CUDA(cudaMemcpyAsync(out_prob, s->d_prob, sizeof(float) * s->nnz, cudaMemcpyDefault, st));
    OAR_CUDA(cudaStreamSynchronize(st));
    return OAR_OK;
}
