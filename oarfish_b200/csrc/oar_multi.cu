// oar_multi.cu -- the GPUs of one box behind ONE call from ONE host thread.
//
// The reference's em::bootstrap(em_info, num_boot, nthreads) (src/em.rs:292-314) builds its own rayon pool and fans
// the replicates out internally; the single-cell driver (src/single_cell.rs:91-193) does the same with cells.  A
// Rust caller linking this library gets the same shape: one call, and the library runs one host thread per device.
//
//   oar_multi_create     upload once to devices[0], then device-to-device copies of the validated CSR over NVLink
//                        (cudaMemcpyPeerAsync) to the other devices; every device builds its own tiled layout
//   oar_multi_bootstrap  replicates are independent units: the device threads pull global replicate ids from a shared
//                        atomic counter (replicates differ 2x in iteration count, so a static split idles devices);
//                        weights depend on (seed, id) only, so results do not depend on who ran what; results are
//                        copied straight into out[id]; no collective anywhere on the data path
//   oar_em_batched_multi cells are independent units: contiguous cell ranges balanced on alignments, every device gets
//                        ONLY its cells' rows (host-to-device scatter of a slice, no broadcast), results are stitched
//                        into one CSR over cells on the host
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "oar_store.cuh"

using namespace oar;

struct oar_multi {
    std::vector<int> devices;
    std::vector<oar_store *> stores;   // stores[i] lives on devices[i]
    uint64_t n_reads = 0, nnz = 0;
    uint32_t n_txps = 0;
    double ms_upload = 0, ms_replicate = 0;   // host-to-device upload + layout on devices[0]; peer copies + layouts elsewhere (wall)
    std::vector<uint32_t> last_per_device;    // units (replicates) each device ran in the last call
};

namespace {
double now_ms()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

int check_devices(const int *devices, int n_devices, const char *who)
{
    if (!devices || n_devices <= 0) return fail(OAR_ERR_INVALID, std::string(who) + ": devices is empty");
    int ndev = 0;
    OAR_CUDA(cudaGetDeviceCount(&ndev));
    for (int i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(OAR_ERR_INVALID, std::string(who) + ": bad device index");
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(OAR_ERR_INVALID, std::string(who) + ": a device is listed twice");
    }
    return OAR_OK;
}

// Runs fn(i) on one host thread per device and returns the first failure (status + message; the workers' thread-local
// error strings do not reach the caller otherwise).
template <typename F>
int for_each_device(int n, F fn)
{
    std::vector<int> rc(n, OAR_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    th.reserve(n);
    for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            rc[i] = fn(i);
            if (rc[i] != OAR_OK) msg[i] = oar_last_error();
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i] != OAR_OK) return fail(rc[i], msg[i]);
    return OAR_OK;
}
}  // namespace

extern "C" void oar_multi_destroy(oar_multi *m)
{
    if (!m) return;
    for (oar_store *s : m->stores) oar_store_destroy(s);
    delete m;
}

extern "C" int oar_multi_create(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob, const double *aux_or_null,
                                uint64_t n_reads, uint64_t nnz, uint32_t n_txps, const int *devices, int n_devices,
                                oar_multi **out)
{
    if (!out) return fail(OAR_ERR_INVALID, "oar_multi_create: out is null");
    *out = nullptr;
    int rc = check_devices(devices, n_devices, "oar_multi_create");
    if (rc != OAR_OK) return rc;
    oar_multi *m = new (std::nothrow) oar_multi();
    if (!m) return fail(OAR_ERR_OOM, "oar_multi_create: host allocation failed");
    m->devices.assign(devices, devices + n_devices);
    m->stores.assign(n_devices, nullptr);
    m->last_per_device.assign(n_devices, 0);
    m->n_reads = n_reads; m->nnz = nnz; m->n_txps = n_txps;
    double t0 = now_ms();
    rc = oar_store_create(row_ptr, txp_id, prob, aux_or_null, n_reads, nnz, n_txps, devices[0], &m->stores[0]);
    m->ms_upload = now_ms() - t0;
    if (rc == OAR_OK && n_devices > 1) {
        t0 = now_ms();
        rc = for_each_device(n_devices - 1, [&](int k) { return store_clone(m->stores[0], m->devices[k + 1], &m->stores[k + 1]); });
        m->ms_replicate = now_ms() - t0;
    }
    if (rc != OAR_OK) { std::string keep = oar_last_error(); oar_multi_destroy(m); return fail(rc, keep); }
    *out = m;
    return OAR_OK;
}

extern "C" int oar_multi_info(const oar_multi *m, int *n_devices, double out_ms[2], uint32_t *out_last_per_device)
{
    if (!m) return fail(OAR_ERR_INVALID, "oar_multi_info: handle is null");
    if (n_devices) *n_devices = (int)m->devices.size();
    if (out_ms) { out_ms[0] = m->ms_upload; out_ms[1] = m->ms_replicate; }
    if (out_last_per_device) std::memcpy(out_last_per_device, m->last_per_device.data(), sizeof(uint32_t) * m->last_per_device.size());
    return OAR_OK;
}

extern "C" oar_store *oar_multi_store(oar_multi *m, int i)
{
    if (!m || i < 0 || i >= (int)m->stores.size()) return nullptr;
    return m->stores[i];
}

extern "C" int oar_multi_bootstrap(oar_multi *m, uint32_t num_boot, uint64_t seed, uint32_t max_iter, double conv_thresh,
                                   double *out, uint32_t *out_niter)
{
    if (!m) return fail(OAR_ERR_INVALID, "oar_multi_bootstrap: handle is null");
    if (num_boot > 0 && !out) return fail(OAR_ERR_INVALID, "oar_multi_bootstrap: out is null");
    const int G = (int)m->devices.size();
    std::atomic<uint32_t> next{0};
    std::atomic<bool> failed{false};
    std::vector<uint32_t> ran(G, 0);
    const uint64_t M = m->n_txps;
    int rc = for_each_device(G, [&](int i) -> int {
        for (;;) {
            if (failed.load(std::memory_order_relaxed)) return OAR_OK;
            const uint32_t g = next.fetch_add(1, std::memory_order_relaxed);
            if (g >= num_boot) return OAR_OK;
            const int r = oar_bootstrap(m->stores[i], 1, seed, g, 1, max_iter, conv_thresh, out + (uint64_t)g * M,
                                        out_niter ? out_niter + g : nullptr);
            if (r != OAR_OK) { failed.store(true); return r; }
            ++ran[i];
        }
    });
    m->last_per_device = ran;
    return rc;
}

extern "C" int oar_em_batched_multi(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob, const double *aux_or_null,
                                    uint64_t n_reads, uint64_t nnz, uint32_t n_txps, const uint64_t *cell_row_ptr,
                                    uint32_t n_cells, const int *devices, int n_devices, uint32_t max_iter, double conv_thresh,
                                    uint32_t min_iter, uint64_t *out_cell_ptr, uint32_t *out_txp, double *out_val,
                                    uint64_t capacity, uint64_t *out_nnz, uint32_t *out_niter, uint32_t *out_cells_per_device)
{
    if (!row_ptr || !cell_row_ptr || !out_cell_ptr || !out_nnz) return fail(OAR_ERR_INVALID, "oar_em_batched_multi: null argument");
    if (nnz > 0 && (!txp_id || !prob)) return fail(OAR_ERR_INVALID, "oar_em_batched_multi: txp_id/prob is null");
    int rc = check_devices(devices, n_devices, "oar_em_batched_multi");
    if (rc != OAR_OK) return rc;
    if (cell_row_ptr[0] != 0 || cell_row_ptr[n_cells] != n_reads)
        return fail(OAR_ERR_INVALID, "oar_em_batched_multi: cell_row_ptr must span [0, n_reads]");
    for (uint32_t c = 0; c < n_cells; ++c)
        if (cell_row_ptr[c + 1] < cell_row_ptr[c] || cell_row_ptr[c + 1] > n_reads)
            return fail(OAR_ERR_INVALID, "oar_em_batched_multi: cell_row_ptr is not monotone");
    if (row_ptr[0] != 0 || row_ptr[n_reads] != nnz) return fail(OAR_ERR_INVALID, "oar_em_batched_multi: row_ptr does not end at nnz");
    // contiguous cell ranges with about nnz / G alignments each: range i ends at the first cell boundary at or past
    // alignment (i + 1) * nnz / G
    const int G = n_devices;
    std::vector<uint32_t> cut(G + 1, n_cells);
    cut[0] = 0;
    {
        uint32_t c = 0;
        for (int i = 1; i < G; ++i) {
            const uint64_t target = (uint64_t)((__uint128_t)nnz * (uint64_t)i / (uint64_t)G);
            while (c < n_cells && row_ptr[cell_row_ptr[c]] < target) ++c;
            cut[i] = c;
        }
    }
    struct Part { std::vector<uint64_t> cell_ptr; std::vector<uint32_t> txp, niter; std::vector<double> val; uint64_t n = 0; };
    std::vector<Part> parts(G);
    rc = for_each_device(G, [&](int i) -> int {
        const uint32_t c0 = cut[i], c1 = cut[i + 1];
        Part &p = parts[i];
        if (c1 <= c0) return OAR_OK;
        const uint64_t r0 = cell_row_ptr[c0], r1 = cell_row_ptr[c1];
        const uint64_t a0 = row_ptr[r0], a1 = row_ptr[r1];
        oar_store *s = nullptr;
        int r = store_create_slice(row_ptr + r0, a0, txp_id + a0, prob + a0, aux_or_null ? aux_or_null + a0 : nullptr, r1 - r0,
                                   a1 - a0, n_txps, devices[i], &s);
        if (r != OAR_OK) return r;
        std::vector<uint64_t> local(c1 - c0 + 1);
        for (uint32_t c = c0; c <= c1; ++c) local[c - c0] = cell_row_ptr[c] - r0;
        p.cell_ptr.assign(c1 - c0 + 1, 0);
        p.niter.assign(c1 - c0, 0);
        // size query first (capacity 0), then the run: the number of (cell, transcript) pairs is not known up front
        uint64_t need = 0;
        r = oar_em_batched(s, local.data(), c1 - c0, max_iter, conv_thresh, min_iter, p.cell_ptr.data(), nullptr, nullptr, 0, &need, nullptr);
        if (r == OAR_OK || need > 0) {
            p.txp.resize(need); p.val.resize(need);
            r = oar_em_batched(s, local.data(), c1 - c0, max_iter, conv_thresh, min_iter, p.cell_ptr.data(), p.txp.data(),
                               p.val.data(), need, &p.n, p.niter.data());
        }
        std::string keep = r != OAR_OK ? oar_last_error() : "";
        oar_store_destroy(s);
        if (r != OAR_OK) return fail(r, keep);
        return OAR_OK;
    });
    if (rc != OAR_OK) return rc;
    uint64_t total = 0;
    for (int i = 0; i < G; ++i) total += parts[i].n;
    *out_nnz = total;
    out_cell_ptr[0] = 0;
    uint64_t off = 0;
    for (int i = 0; i < G; ++i) {
        const uint32_t c0 = cut[i], c1 = cut[i + 1];
        for (uint32_t c = c0; c < c1; ++c) out_cell_ptr[c + 1] = off + parts[i].cell_ptr[c + 1 - c0];
        off += parts[i].n;
        if (out_cells_per_device) out_cells_per_device[i] = c1 - c0;
    }
    if (total > capacity || (total > 0 && (!out_txp || !out_val)))
        return fail(OAR_ERR_INVALID, "oar_em_batched_multi: output capacity too small (required size returned in out_nnz)");
    off = 0;
    for (int i = 0; i < G; ++i) {
        const Part &p = parts[i];
        if (p.n) {
            std::memcpy(out_txp + off, p.txp.data(), sizeof(uint32_t) * p.n);
            std::memcpy(out_val + off, p.val.data(), sizeof(double) * p.n);
        }
        if (out_niter && !p.niter.empty()) std::memcpy(out_niter + cut[i], p.niter.data(), sizeof(uint32_t) * p.niter.size());
        off += p.n;
    }
    return OAR_OK;
}
