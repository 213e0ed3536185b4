// oarfish_em.hpp -- C++ host-side mirror of oarfish's inference interface over the C ABI.
//
// The reference is compiled code (Rust) and no Rust toolchain exists in this image, so the
// reference-shaped host layer is C++: same names, argument meaning and return shapes as
//
//   oarfish::em(em_info, nthreads)            <- em::em        (src/em.rs:262)
//   oarfish::em_par(em_info, nthreads)        <- em::em_par    (src/em.rs:320)
//   oarfish::bootstrap(em_info, n, nthreads)  <- em::bootstrap (src/em.rs:292)
//
// with AlnInfo / TranscriptInfo / InMemoryAlignmentStore / EMInfo mirroring
// src/util/oarfish_types.rs:330-344, :430-437, :547-738 and :408-428 as far as the EM reads them.
// The Rust shim in INTEGRATION.md does exactly what `detail::upload` does here.
// Errors of the C ABI become std::runtime_error (the reference aborts on failure, Cargo.toml:116).
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "oarfish_em.h"

namespace oarfish {

enum class Strand : uint8_t { Forward = 0, Reverse = 1 };

// AlnInfo, oarfish_types.rs:330-337 (24 bytes after rustc's field reordering)
struct AlnInfo {
    double prob = 0.0;  // always 0.0 (oarfish_types.rs:352); unused by the EM
    uint32_t ref_id = 0;
    uint32_t start = 0;
    uint32_t end = 0;
    Strand strand = Strand::Forward;
    uint32_t alignment_span() const { return end - start; }  // :341-343
};
static_assert(sizeof(AlnInfo) == 24, "AlnInfo must stay 24 bytes like the Rust struct");

// TranscriptInfo, oarfish_types.rs:430-437 (the EM reads only lenf, for the KDE hook)
struct TranscriptInfo {
    size_t len = 1;
    double lenf = 1.0;
};

struct AlignmentFilters {
    bool model_coverage = false;  // the only field of AlignmentFilters the EM reads (em.rs:108)
};

namespace detail {
struct StoreDeleter { void operator()(oar_store *s) const { oar_store_destroy(s); } };
inline void check(int rc, const char *what)
{
    if (rc != OAR_OK) throw std::runtime_error(std::string(what) + ": " + oar_last_error());
}
}  // namespace detail

// InMemoryAlignmentStore, oarfish_types.rs:547-558
class InMemoryAlignmentStore {
public:
    AlignmentFilters filter_opts;
    std::vector<AlnInfo> alignments;
    std::vector<float> as_probabilities;
    std::vector<double> coverage_probabilities;

    explicit InMemoryAlignmentStore(AlignmentFilters fo = {}) : filter_opts(fo), boundaries_{0} {}
    // copies share no device state: the copy uploads its own store on first use
    InMemoryAlignmentStore(const InMemoryAlignmentStore &o)
        : filter_opts(o.filter_opts), alignments(o.alignments), as_probabilities(o.as_probabilities),
          coverage_probabilities(o.coverage_probabilities), boundaries_(o.boundaries_) {}
    InMemoryAlignmentStore &operator=(const InMemoryAlignmentStore &o)
    {
        if (this != &o) {
            filter_opts = o.filter_opts; alignments = o.alignments; as_probabilities = o.as_probabilities;
            coverage_probabilities = o.coverage_probabilities; boundaries_ = o.boundaries_;
            invalidate_device();
        }
        return *this;
    }

    // add_filtered_group, oarfish_types.rs:718-738: empty groups are dropped
    bool add_filtered_group(const std::vector<AlnInfo> &alns, const std::vector<float> &as_probs)
    {
        if (alns.empty()) return false;
        alignments.insert(alignments.end(), alns.begin(), alns.end());
        as_probabilities.insert(as_probabilities.end(), as_probs.begin(), as_probs.end());
        coverage_probabilities.insert(coverage_probabilities.end(), alns.size(), 0.0);  // :731-732
        boundaries_.push_back(alignments.size());
        device_.reset();
        return true;
    }
    size_t len() const { return boundaries_.size() - 1; }          // :562-564
    size_t num_aligned_reads() const { return len(); }             // :745-747
    size_t total_len() const { return alignments.size(); }         // :740-742
    const std::vector<size_t> &boundaries() const { return boundaries_; }

    // the device copy (created on first use, reused by em then bootstrap like bulk.rs:155-179).  It is a snapshot:
    // after changing alignments / as_probabilities / coverage_probabilities / filter_opts in place (the reference's
    // normalize_read_probs replaces coverage_probabilities once the store is built) call invalidate_device().
    oar_store *device_store(uint32_t n_txps, int device) const
    {
        if (!device_ || device_txps_ != n_txps || device_id_ != device || device_nnz_ != alignments.size() ||
            device_cov_ != filter_opts.model_coverage)
            upload(n_txps, device);
        return device_.get();
    }
    void invalidate_device() const { device_.reset(); multi_.reset(); }
    // one copy per device for em::bootstrap's fan-out (oar_multi_*)
    oar_multi *multi_store(uint32_t n_txps, const std::vector<int> &devices) const
    {
        if (!multi_ || multi_txps_ != n_txps || multi_devices_ != devices || multi_nnz_ != alignments.size() ||
            multi_cov_ != filter_opts.model_coverage) {
            Flat f = flatten();
            oar_multi *h = nullptr;
            detail::check(oar_multi_create(f.row_ptr.data(), f.txp_id.data(), as_probabilities.data(), f.aux, len(), total_len(),
                                           n_txps, devices.data(), (int)devices.size(), &h),
                          "oar_multi_create");
            multi_.reset(h);
            multi_txps_ = n_txps; multi_devices_ = devices; multi_nnz_ = alignments.size(); multi_cov_ = filter_opts.model_coverage;
        }
        return multi_.get();
    }

private:
    std::vector<size_t> boundaries_;  // private in the reference too (:555)
    mutable std::unique_ptr<oar_store, detail::StoreDeleter> device_;
    mutable uint32_t device_txps_ = 0;
    mutable int device_id_ = -1;
    mutable size_t device_nnz_ = 0;
    mutable bool device_cov_ = false;
    struct MultiDeleter { void operator()(oar_multi *m) const { oar_multi_destroy(m); } };
    mutable std::unique_ptr<oar_multi, MultiDeleter> multi_;
    mutable uint32_t multi_txps_ = 0;
    mutable std::vector<int> multi_devices_;
    mutable size_t multi_nnz_ = 0;
    mutable bool multi_cov_ = false;

    // flatten exactly as the EM reads the store (em.rs:97-131): boundaries -> row_ptr u64,
    // AlnInfo.ref_id -> txp_id, as_probabilities -> prob, coverage_probabilities -> aux iff model_coverage
    struct Flat { std::vector<uint64_t> row_ptr; std::vector<uint32_t> txp_id; const double *aux; };
    Flat flatten() const
    {
        Flat f;
        f.row_ptr.assign(boundaries_.begin(), boundaries_.end());
        f.txp_id.resize(alignments.size());
        for (size_t j = 0; j < alignments.size(); ++j) f.txp_id[j] = alignments[j].ref_id;
        f.aux = filter_opts.model_coverage ? coverage_probabilities.data() : nullptr;
        return f;
    }
    void upload(uint32_t n_txps, int device) const
    {
        Flat f = flatten();
        oar_store *h = nullptr;
        detail::check(oar_store_create(f.row_ptr.data(), f.txp_id.data(), as_probabilities.data(), f.aux, len(), total_len(),
                                       n_txps, device, &h),
                      "oar_store_create");
        device_.reset(h);
        device_txps_ = n_txps;
        device_id_ = device;
        device_nnz_ = alignments.size();
        device_cov_ = filter_opts.model_coverage;
    }
};

// EMInfo, oarfish_types.rs:408-428
struct EMInfo {
    const InMemoryAlignmentStore *eq_map = nullptr;
    const std::vector<TranscriptInfo> *txp_info = nullptr;
    uint32_t max_iter = 1000;           // --max-em-iter default, prog_opts.rs:532
    double convergence_thresh = 1e-3;   // --convergence-thresh default, prog_opts.rs:536
    std::optional<std::vector<double>> init_abundances;
    // kde_model (em.rs:173-178): the density factor comes from the un-vendored `kders` crate; fold it into
    // coverage_probabilities and set filter_opts.model_coverage.
    int device = 0;
    std::vector<int> devices;   // em::bootstrap fans out over these (empty: `device` only)
};

namespace detail {
inline std::vector<double> run(const EMInfo &emi, uint32_t min_iter)
{
    const uint32_t m = (uint32_t)emi.txp_info->size();
    oar_store *st = emi.eq_map->device_store(m, emi.device);
    std::vector<double> out(m);
    uint32_t niter = 0;
    double rel = 0.0;
    const double *init = emi.init_abundances ? emi.init_abundances->data() : nullptr;
    check(oar_em(st, init, emi.max_iter, emi.convergence_thresh, min_iter, out.data(), &niter, &rel), "oar_em");
    return out;
}
}  // namespace detail

// em::em (em.rs:262-271): do_em with the `niter > 50` stop rule (em.rs:212)
inline std::vector<double> em(const EMInfo &em_info, size_t /*nthreads*/) { return detail::run(em_info, 50); }

// em::em_par (em.rs:320-447): same EM, stop rule `niter > 1` (em.rs:399)
inline std::vector<double> em_par(const EMInfo &em_info, size_t /*nthreads*/) { return detail::run(em_info, 1); }

// em::bootstrap (em.rs:292-314): result[replicate][transcript].  The reference draws from the unseeded
// thread RNG (em.rs:274); pass a seed for reproducible replicates.
inline std::vector<std::vector<double>> bootstrap(const EMInfo &em_info, uint32_t num_boot, size_t /*nthreads*/,
                                                  std::optional<uint64_t> seed = std::nullopt)
{
    const uint32_t m = (uint32_t)em_info.txp_info->size();
    const uint64_t sd = seed ? *seed : ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}();
    std::vector<double> flat((size_t)num_boot * m);
    if (em_info.devices.size() > 1) {   // the pool of em.rs:296-300, with devices in place of threads
        oar_multi *mt = em_info.eq_map->multi_store(m, em_info.devices);
        detail::check(oar_multi_bootstrap(mt, num_boot, sd, em_info.max_iter, em_info.convergence_thresh, flat.data(), nullptr),
                      "oar_multi_bootstrap");
    } else {
        oar_store *st = em_info.eq_map->device_store(m, em_info.devices.empty() ? em_info.device : em_info.devices[0]);
        detail::check(oar_bootstrap(st, num_boot, sd, 0, 1, em_info.max_iter, em_info.convergence_thresh, flat.data(), nullptr),
                      "oar_bootstrap");
    }
    std::vector<std::vector<double>> out(num_boot);
    for (uint32_t b = 0; b < num_boot; ++b) out[b].assign(flat.begin() + (size_t)b * m, flat.begin() + (size_t)(b + 1) * m);
    return out;
}

}  // namespace oarfish
