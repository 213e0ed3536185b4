// oar_em.cu -- C ABI + EM driver of the B200-native oarfish EM engine.
//
// Reference semantics implemented here (files under /root/reference/src):
//   em.rs:87-133   m_step      -> em_sweep_* kernels (fused E-step + M-step)
//   em.rs:144-255  do_em       -> run_em(): on-device convergence, CUDA graph
//   em.rs:320-447  em_par      -> same driver with min_iter = 1
//   em.rs:273-314  bootstrap   -> oar_bootstrap*: multinomial read weights
//   bootstrap.rs:7-16          -> boot_weights_kernel (Philox4x32-10)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <vector>

#include "oar_store.cuh"
#include "oar_tiled.cuh"

namespace oar {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) { g_last_error = msg; return code; }
int cuda_fail(cudaError_t e, const char *what)
{
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? OAR_ERR_OOM : OAR_ERR_CUDA;
}

}  // namespace oar

using namespace oar;

// ---------------------------------------------------------------------------
// store
// ---------------------------------------------------------------------------

static const int kGraphIters = 18;  // EM iterations per graph launch (a multiple of 3: three count buffers rotate)

static void destroy_graphs(oar_store *s)
{
    for (auto &g : s->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
    }
}

extern "C" int oar_version(void) { return 1000; }

extern "C" const char *oar_last_error(void) { return g_last_error.c_str(); }

extern "C" int oar_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

extern "C" void oar_store_destroy(oar_store *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    destroy_graphs(s);
    free_tiled_layout(s);
    if (!s->borrowed) { dfree(s->d_row_ptr, s->stream); dfree(s->d_prob, s->stream); dfree(s->d_aux, s->stream); }
    dfree(s->d_txp, s->stream);
    dfree(s->d_counts[0], s->stream); dfree(s->d_counts[1], s->stream); dfree(s->d_counts[2], s->stream); dfree(s->d_state, s->stream); dfree(s->d_weights, s->stream);
    if (s->stream) cudaStreamSynchronize(s->stream);
    // stream, events and the pinned state block go back to the device context's pools (cudaFreeHost / cudaStreamDestroy
    // synchronise the whole context: 0.8-400 ms per store on the GPU box)
    if (s->ctx) {
        ctx_give_host_state(s->ctx, s->h_state);
        if (!s->borrowed) {
            for (auto &e : s->ev) ctx_give_event(s->ctx, true, e);
            for (auto &e : s->slot_ev) ctx_give_event(s->ctx, false, e);
            ctx_give_stream(s->ctx, s->stream);
        }
    }
    delete s;
    (void)cudaGetLastError();
}

namespace oar {
int substore_create(oar_store *parent, uint32_t *d_txp, uint32_t n_txps, oar_store **out)
{
    *out = nullptr;
    oar_store *s = new (std::nothrow) oar_store();
    if (!s) return fail(OAR_ERR_OOM, "substore_create: host allocation failed");
    s->borrowed = true;
    s->device = parent->device; s->sm_count = parent->sm_count; s->stream = parent->stream; s->ctx = parent->ctx;
    s->n_reads = parent->n_reads; s->nnz = parent->nnz; s->n_txps = n_txps;
    s->d_row_ptr = parent->d_row_ptr; s->d_prob = parent->d_prob; s->d_aux = parent->d_aux; s->d_txp = d_txp;
    for (int i = 0; i < 4; ++i) s->ev[i] = parent->ev[i];
    for (int i = 0; i < 2; ++i) s->slot_ev[i] = parent->slot_ev[i];
    s->ctas_per_sm = parent->ctas_per_sm; s->allow_fused = parent->allow_fused;
    int rc = [&]() -> int {
        OAR_CUDA(dmalloc(&s->d_state, sizeof(OarEmState) * 3, s->stream));
        OAR_CUDA(cudaMemsetAsync(s->d_state, 0, sizeof(OarEmState) * 3, s->stream));
        OAR_CUDA(ctx_take_host_state(s->ctx, &s->h_state));
        if (parent->tl.ready) {
            const int rc2 = build_tiled_layout(s, parent->tl.span);
            if (rc2 == OAR_OK) s->kernel = OAR_KERNEL_TILED;
            else if (rc2 != OAR_ERR_UNSUPPORTED) return rc2;
        }
        return OAR_OK;
    }();
    if (rc != OAR_OK) { std::string keep = g_last_error; oar_store_destroy(s); g_last_error = keep; return rc; }
    *out = s;
    return OAR_OK;
}
}  // namespace oar

namespace oar {
// Shared tail of store creation: the CSR arrays are resident (d_row_ptr u32, d_txp, d_prob, d_aux); build the tiled
// layout and read the tuning environment.
int finish_store(oar_store *s)
{
    // OAR_TILED=0 keeps only the CSR (row-group kernel); OAR_TILE_SPAN tunes the tile fill
    const char *env = getenv("OAR_TILED");
    if (!(env && env[0] == '0')) {
        const char *sp = getenv("OAR_TILE_SPAN");
        int rc2 = build_tiled_layout(s, sp ? (uint32_t)atoi(sp) : 0u);
        if (rc2 == OAR_OK) s->kernel = OAR_KERNEL_TILED;
        else if (rc2 != OAR_ERR_UNSUPPORTED) return rc2;   // unsupported shape: keep the CSR kernel
    }
    const char *fu = getenv("OAR_FUSED_UPDATE");   // 0: em_update as its own launch between the sweeps (A/B timing)
    s->allow_fused = !(fu && fu[0] == '0');
    const char *cps = getenv("OAR_CTAS_PER_SM");
    if (cps && atoi(cps) > 0) s->ctas_per_sm = atoi(cps);
    return OAR_OK;
}

// Handle + stream + events + pinned state from the device context; EM work buffers.
int new_store(int device, uint64_t n_reads, uint64_t nnz, uint32_t n_txps, const char *who, oar_store **out)
{
    *out = nullptr;
    if (n_txps == 0) return fail(OAR_ERR_INVALID, std::string(who) + ": n_txps must be > 0");
    if (nnz >= 0xFFFFFFF0ull || n_reads >= 0xFFFFFFF0ull)
        return fail(OAR_ERR_UNSUPPORTED, std::string(who) + ": stores with >= 2^32 alignments or reads are not supported");
    int ndev = 0;
    OAR_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(OAR_ERR_INVALID, std::string(who) + ": bad device index");
    cudaError_t ce = cudaSuccess;
    DeviceCtx *ctx = device_ctx(device, &ce);
    if (!ctx) return cuda_fail(ce, "device context");
    oar_store *s = new (std::nothrow) oar_store();
    if (!s) return fail(OAR_ERR_OOM, std::string(who) + ": host allocation failed");
    s->device = device; s->ctx = ctx; s->sm_count = ctx->sm_count;
    s->n_reads = n_reads; s->nnz = nnz; s->n_txps = n_txps;
    int rc = [&]() -> int {
        OAR_CUDA(ctx_take_stream(ctx, &s->stream));
        for (auto &e : s->ev) OAR_CUDA(ctx_take_event(ctx, true, &e));
        for (auto &e : s->slot_ev) OAR_CUDA(ctx_take_event(ctx, false, &e));
        OAR_CUDA(ctx_take_host_state(ctx, &s->h_state));
        OAR_CUDA(dmalloc(&s->d_counts[0], sizeof(double) * n_txps, s->stream));
        OAR_CUDA(dmalloc(&s->d_counts[1], sizeof(double) * n_txps, s->stream));
        OAR_CUDA(dmalloc(&s->d_counts[2], sizeof(double) * n_txps, s->stream));
        OAR_CUDA(dmalloc(&s->d_state, sizeof(OarEmState) * 2, s->stream));
        return OAR_OK;
    }();
    if (rc != OAR_OK) { std::string keep = g_last_error; oar_store_destroy(s); g_last_error = keep; return rc; }
    *out = s;
    return OAR_OK;
}
}  // namespace oar

namespace oar {
// oar_store_create with the boundaries of a slice: row_ptr[0] == row_base, the slice's alignments start at txp_id[0].
int store_create_slice(const uint64_t *row_ptr, uint64_t row_base, const uint32_t *txp_id, const float *prob,
                       const double *aux_or_null, uint64_t n_reads, uint64_t nnz, uint32_t n_txps, int device, oar_store **out)
{
    if (!out) return fail(OAR_ERR_INVALID, "oar_store_create: out is null");
    *out = nullptr;
    if (!row_ptr) return fail(OAR_ERR_INVALID, "oar_store_create: row_ptr is null");
    if (nnz > 0 && (!txp_id || !prob)) return fail(OAR_ERR_INVALID, "oar_store_create: txp_id/prob is null");
    oar_store *s = nullptr;
    int rc = new_store(device, n_reads, nnz, n_txps, "oar_store_create", &s);
    if (rc != OAR_OK) return rc;
    rc = [&]() -> int {
        OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
        const size_t pad = 16;  // slack so vector loads may over-read safely
        OAR_CUDA(dmalloc(&s->d_row_ptr, sizeof(uint32_t) * (n_reads + 1 + pad), s->stream));
        OAR_CUDA(dmalloc(&s->d_txp, sizeof(uint32_t) * (nnz + pad), s->stream));
        OAR_CUDA(dmalloc(&s->d_prob, sizeof(float) * (nnz + pad), s->stream));
        if (aux_or_null) OAR_CUDA(dmalloc(&s->d_aux, sizeof(double) * (nnz + pad), s->stream));
        // stage the u64 boundaries, narrow to u32 and validate on the device
        uint64_t *d_rp64 = nullptr;
        uint32_t *d_flag = reinterpret_cast<uint32_t *>(s->d_state + 1);
        OAR_CUDA(dmalloc(&d_rp64, sizeof(uint64_t) * (n_reads + 1), s->stream));
        OAR_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t) * 4, s->stream));
        OAR_CUDA(cudaMemcpyAsync(d_rp64, row_ptr, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyDefault, s->stream));
        if (nnz) {
            OAR_CUDA(cudaMemcpyAsync(s->d_txp, txp_id, sizeof(uint32_t) * nnz, cudaMemcpyDefault, s->stream));
            OAR_CUDA(cudaMemcpyAsync(s->d_prob, prob, sizeof(float) * nnz, cudaMemcpyDefault, s->stream));
            if (aux_or_null)
                OAR_CUDA(cudaMemcpyAsync(s->d_aux, aux_or_null, sizeof(double) * nnz, cudaMemcpyDefault, s->stream));
        }
        OAR_CUDA(cudaMemsetAsync(s->d_txp + nnz, 0, sizeof(uint32_t) * pad, s->stream));
        OAR_CUDA(cudaMemsetAsync(s->d_prob + nnz, 0, sizeof(float) * pad, s->stream));
        {
            const int threads = 256;
            const int blocks = (int)std::min<uint64_t>((n_reads + threads) / threads, (uint64_t)s->sm_count * 16);
            kern::narrow_validate_rowptr<<<blocks, threads, 0, s->stream>>>(d_rp64, s->d_row_ptr, n_reads, nnz, row_base, d_flag);
            const int blocks2 = (int)std::max<uint64_t>(1, std::min<uint64_t>((nnz + threads - 1) / threads, (uint64_t)s->sm_count * 16));
            kern::validate_txp<<<blocks2, threads, 0, s->stream>>>(s->d_txp, nnz, n_txps, d_flag + 1);
        }
        uint32_t *h_flag = reinterpret_cast<uint32_t *>(&s->h_state[kHostStateSlots - 1]);   // pinned: the copy does not stage
        OAR_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, s->stream));
        OAR_CUDA(cudaStreamSynchronize(s->stream));
        dfree(d_rp64, s->stream);
        OAR_CUDA(cudaGetLastError());
        if (h_flag[0]) return fail(OAR_ERR_INVALID, "oar_store_create: row_ptr is not a monotone prefix ending at nnz");
        if (h_flag[1]) return fail(OAR_ERR_INVALID, "oar_store_create: txp_id out of range (>= n_txps)");
        int rc2 = finish_store(s);
        if (rc2 != OAR_OK) return rc2;
        OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
        OAR_CUDA(cudaStreamSynchronize(s->stream));
        float ms = 0.f;
        OAR_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        s->timings[0] = ms;
        return OAR_OK;
    }();
    if (rc != OAR_OK) {
        std::string keep = g_last_error;
        oar_store_destroy(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return OAR_OK;
}

// A copy of `src` (its CSR as uploaded and validated) on another device: device-to-device over NVLink when the two
// devices are peers (cudaMemcpyPeerAsync; staged through the host otherwise), then this device's own tiled layout.
int store_clone(const oar_store *src, int device, oar_store **out)
{
    if (!src || !out) return fail(OAR_ERR_INVALID, "store_clone: null argument");
    *out = nullptr;
    oar_store *s = nullptr;
    int rc = new_store(device, src->n_reads, src->nnz, src->n_txps, "store_clone", &s);
    if (rc != OAR_OK) return rc;
    rc = [&]() -> int {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(pe, "cudaDeviceEnablePeerAccess");
            (void)cudaGetLastError();
        }
        OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
        const size_t pad = 16, N = src->n_reads, nnz = src->nnz;
        OAR_CUDA(dmalloc(&s->d_row_ptr, sizeof(uint32_t) * (N + 1 + pad), s->stream));
        OAR_CUDA(dmalloc(&s->d_txp, sizeof(uint32_t) * (nnz + pad), s->stream));
        OAR_CUDA(dmalloc(&s->d_prob, sizeof(float) * (nnz + pad), s->stream));
        if (src->d_aux) OAR_CUDA(dmalloc(&s->d_aux, sizeof(double) * (nnz + pad), s->stream));
        OAR_CUDA(cudaMemsetAsync(s->d_state, 0, sizeof(OarEmState) * 2, s->stream));
        OAR_CUDA(cudaMemcpyPeerAsync(s->d_row_ptr, device, src->d_row_ptr, src->device, sizeof(uint32_t) * (N + 1), s->stream));
        OAR_CUDA(cudaMemcpyPeerAsync(s->d_txp, device, src->d_txp, src->device, sizeof(uint32_t) * (nnz + pad), s->stream));
        OAR_CUDA(cudaMemcpyPeerAsync(s->d_prob, device, src->d_prob, src->device, sizeof(float) * (nnz + pad), s->stream));
        if (src->d_aux) OAR_CUDA(cudaMemcpyPeerAsync(s->d_aux, device, src->d_aux, src->device, sizeof(double) * nnz, s->stream));
        OAR_CUDA(cudaEventRecord(s->ev[2], s->stream));
        int rc2 = finish_store(s);
        if (rc2 != OAR_OK) return rc2;
        OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
        OAR_CUDA(cudaStreamSynchronize(s->stream));
        float ms = 0.f, ms_copy = 0.f;
        OAR_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        OAR_CUDA(cudaEventElapsedTime(&ms_copy, s->ev[0], s->ev[2]));
        s->timings[0] = ms; s->timings[3] = ms_copy;
        return OAR_OK;
    }();
    if (rc != OAR_OK) { std::string keep = g_last_error; oar_store_destroy(s); g_last_error = keep; return rc; }
    *out = s;
    return OAR_OK;
}
}  // namespace oar

extern "C" int oar_store_create(const uint64_t *row_ptr, const uint32_t *txp_id, const float *prob,
                                const double *aux_or_null, uint64_t n_reads, uint64_t nnz,
                                uint32_t n_txps, int device, oar_store **out)
{
    return store_create_slice(row_ptr, 0, txp_id, prob, aux_or_null, n_reads, nnz, n_txps, device, out);
}

extern "C" int oar_store_info(const oar_store *s, uint64_t *n_reads, uint64_t *nnz, uint32_t *n_txps, int *device)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_store_info: store is null");
    if (n_reads) *n_reads = s->n_reads;
    if (nnz) *nnz = s->nnz;
    if (n_txps) *n_txps = s->n_txps;
    if (device) *device = s->device;
    return OAR_OK;
}

extern "C" int oar_store_set_kernel(oar_store *s, int kernel)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_store_set_kernel: store is null");
    if (kernel == OAR_KERNEL_AUTO) kernel = s->tl.ready ? OAR_KERNEL_TILED : OAR_KERNEL_ROWGROUP;
    if (kernel != OAR_KERNEL_ROWGROUP && kernel != OAR_KERNEL_TILED)
        return fail(OAR_ERR_INVALID, "oar_store_set_kernel: unknown kernel");
    if (kernel == OAR_KERNEL_TILED && !s->tl.ready)
        return fail(OAR_ERR_UNSUPPORTED, "oar_store_set_kernel: the tiled layout was not built (OAR_TILED=0, or a store shape it does not support)");
    if (kernel != s->kernel) { cudaSetDevice(s->device); cudaStreamSynchronize(s->stream); destroy_graphs(s); }
    s->kernel = kernel;
    return OAR_OK;
}

extern "C" int oar_store_set_progress(oar_store *s, oar_progress_fn fn, void *user)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_store_set_progress: store is null");
    s->progress = fn; s->progress_user = user;
    return OAR_OK;
}

extern "C" int oar_store_layout_info(const oar_store *s, uint64_t out[8])
{
    if (!s || !out) return fail(OAR_ERR_INVALID, "oar_store_layout_info: null argument");
    const TiledLayout &t = s->tl;
    out[0] = t.ready ? 1 : 0; out[1] = t.n_tiles;
    out[2] = (uint64_t)t.n_tiles * tiled::kTile;   // alignment slots held in HBM
    out[3] = t.n_fallback;
    out[4] = t.sum_d; out[5] = t.sum_u; out[6] = t.span; out[7] = (uint64_t)s->kernel;
    return OAR_OK;
}

namespace {
__global__ void layout_trash_offsets(const uint2 *__restrict__ rec, const uint4 *__restrict__ records, uint32_t first, uint32_t n, uint32_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(records + rec[first + i].x) + tiled::kRecDU)[3];
}
}  // namespace

extern "C" int oar_store_layout_lpos(oar_store *s, uint32_t first_tile, uint32_t n_tiles, uint32_t *out, uint32_t *out_trash_or_null)
{
    if (!s || !out) return fail(OAR_ERR_INVALID, "oar_store_layout_lpos: null argument");
    const TiledLayout &t = s->tl;
    if (!t.ready || (uint64_t)first_tile + n_tiles > t.n_tiles) return fail(OAR_ERR_INVALID, "oar_store_layout_lpos: tile range outside the layout");
    cudaSetDevice(s->device);
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    OAR_CUDA(cudaMemcpy(out, t.lpos + (size_t)first_tile * tiled::kTile, sizeof(uint32_t) * (size_t)n_tiles * tiled::kTile, cudaMemcpyDeviceToHost));
    if (out_trash_or_null && n_tiles > 0) {
        uint32_t *d = nullptr;
        OAR_CUDA(dmalloc(&d, sizeof(uint32_t) * n_tiles, s->stream));
        layout_trash_offsets<<<(n_tiles + 255) / 256, 256, 0, s->stream>>>(t.rec, t.records, first_tile, n_tiles, d);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_trash_or_null, d, sizeof(uint32_t) * n_tiles, cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        dfree(d, s->stream);
        OAR_CUDA(e);
    }
    return OAR_OK;
}

extern "C" int oar_store_timings(const oar_store *s, double out_ms[4])
{
    if (!s || !out_ms) return fail(OAR_ERR_INVALID, "oar_store_timings: null argument");
    for (int i = 0; i < 4; ++i) out_ms[i] = s->timings[i];
    return OAR_OK;
}

extern "C" int oar_store_counters(const oar_store *s, uint64_t out[2])
{
    if (!s || !out) return fail(OAR_ERR_INVALID, "oar_store_counters: null argument");
    out[0] = s->counters[0]; out[1] = s->counters[1];
    return OAR_OK;
}

extern "C" void *oar_store_stream(oar_store *s) { return s ? (void *)s->stream : nullptr; }

// ---------------------------------------------------------------------------
// sweep dispatch
// ---------------------------------------------------------------------------

static const uint32_t kFoldFallbackMax = 65536;   // fallback rows swept inside the tiled kernel (spread over all its CTAs); longer lists get their own launch

static tiled::View tiled_view(const oar_store *s)
{
    const TiledLayout &t = s->tl;
    tiled::View v;
    v.n_tiles = t.n_tiles; v.prob = t.prob; v.lpos = t.lpos; v.aux = t.aux; v.rec = t.rec; v.records = t.records; v.wlane = t.wlane;
    v.tile_list = nullptr; v.n_active = nullptr;
    // a handful of fallback rows rides along in the tiled kernel; a long list gets its own launch
    const bool fold = t.n_fallback <= kFoldFallbackMax;
    v.fb_rows = t.fallback; v.n_fb = fold ? t.n_fallback : 0u;
    v.csr_row_ptr = s->d_row_ptr; v.csr_txp = s->d_txp; v.csr_prob = s->d_prob; v.csr_aux = s->d_aux; v.csr_wts = nullptr;
    return v;
}

template <bool AUX, bool WTS, bool LIST = false, bool FUSED = false>
static cudaError_t launch_tiled(oar_store *s, const tiled::View &v, const double *prev, double *curr,
                                const uint32_t *wperm, const OarEmState *state, int check_done)
{
    auto kfn = tiled::em_sweep_tiled<AUX, WTS, LIST, FUSED>;
    const tiled::Geometry g = tiled::make_geometry(s->tl.max_rec, s->tl.max_d, s->tl.max_u, WTS);
    int static_bytes = 0;
    cudaError_t ae = ctx_ensure_smem(s->ctx, reinterpret_cast<const void *>(kfn), (int)g.total, &static_bytes);
    if (ae != cudaSuccess) return ae;
    // shared-space address of the dynamic window: 1 KB reserved by the system, then the kernel's static shared memory,
    // rounded up to the 128-byte alignment of the extern array (the kernel checks it against its own cvta)
    tiled::Geometry gg = g;
    gg.xs_base = (1024u + (uint32_t)static_bytes + 127u) & ~127u;
    // persistent CTAs: as many per SM as shared memory allows, capped by the register budget (5)
    int per_sm = (int)((227u * 1024u) / (g.total + 1024u));
    per_sm = std::max(1, std::min(per_sm, s->ctas_per_sm > 0 ? std::min(s->ctas_per_sm, tiled::sweep_ctas(AUX, WTS, LIST)) : tiled::sweep_ctas(AUX, WTS, LIST)));
    const uint32_t grid = std::min<uint32_t>(v.n_tiles, (uint32_t)s->sm_count * (uint32_t)per_sm);
    kfn<<<grid, tiled::kThreads, g.total, s->stream>>>(v, gg, prev, curr, wperm, state, check_done);
    return cudaGetLastError();
}

// Row-group sweep over all rows (list == null) or over a list of row ids.
static cudaError_t enqueue_rowgroup(oar_store *s, const uint32_t *list, uint64_t n_rows, const double *prev,
                                    double *curr, const uint32_t *wts, const OarEmState *state, int check_done)
{
    if (n_rows == 0) return cudaSuccess;
    const int threads = 256;
    const uint64_t groups_per_block = threads / 8;
    const uint64_t want = (n_rows + groups_per_block - 1) / groups_per_block;
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)s->sm_count * 8));
#define OAR_LAUNCH(AUX, WTS)                                                                      \
    kern::em_sweep_rowgroup<AUX, WTS><<<blocks, threads, 0, s->stream>>>(                         \
        s->d_row_ptr, s->d_txp, s->d_prob, s->d_aux, wts, list, prev, curr, n_rows, state, check_done)
    if (s->d_aux) { if (wts) OAR_LAUNCH(true, true); else OAR_LAUNCH(true, false); }
    else          { if (wts) OAR_LAUNCH(false, true); else OAR_LAUNCH(false, false); }
#undef OAR_LAUNCH
    s->counters[0] += 1;
    return cudaGetLastError();
}

// Enqueue one fused E+M sweep prev -> curr (curr must already be zero).
// `wts` are per-read weights in read order; the tiled kernel reads the copy
// permuted into tile order (s->tl.wperm, refreshed by refresh_wperm()).
// the sweep can carry the convergence bookkeeping of the PREVIOUS iteration (tiled kernel only, see fused_update())
// Measured on B200: fused wins wherever one launch matters (C2: 41.7 -> 50.2 k it/s, plain and weighted) and, barely, on the
// plain C3 sweep; for the weighted sweep on C3 the head costs what the launch it saves does (5 408 vs 5 417 it/s at five CTAs
// per SM, 5 466 vs 5 480 at six; earlier builds: 4 965 vs 5 035), so long weighted sweeps keep em_update.
static const uint32_t kFusedWeightedMaxTiles = 32768;
static bool fused_update(const oar_store *s, bool weighted)
{
    if (!(s->allow_fused && s->kernel == OAR_KERNEL_TILED && s->tl.ready && s->tl.n_tiles > 0)) return false;
    static const char *env = getenv("OAR_FUSED_WTS_MAX_TILES");   // development: A/B of the policy
    return !weighted || s->tl.n_tiles <= (env ? (uint32_t)strtoul(env, nullptr, 10) : kFusedWeightedMaxTiles);
}
static cudaError_t enqueue_sweep(oar_store *s, const double *prev, double *curr, const uint32_t *wts,
                                 const OarEmState *state, int check_done, bool fused = false)
{
    if (s->n_reads == 0) return cudaSuccess;
    if (s->kernel != OAR_KERNEL_TILED || !s->tl.ready)   // no layout (OAR_TILED=0, unsupported shape, failed rebuild): the CSR kernel
        return enqueue_rowgroup(s, nullptr, s->n_reads, prev, curr, wts, state, check_done);
    const TiledLayout &t = s->tl;
    if (t.n_tiles > 0) {
        tiled::View v = tiled_view(s);
        v.csr_wts = wts;
        const uint32_t *wp = wts ? t.wperm : nullptr;
        cudaError_t le;
        if (fused) {
            if (s->d_aux) le = wts ? launch_tiled<true, true, false, true>(s, v, prev, curr, wp, state, check_done)
                                   : launch_tiled<true, false, false, true>(s, v, prev, curr, wp, state, check_done);
            else          le = wts ? launch_tiled<false, true, false, true>(s, v, prev, curr, wp, state, check_done)
                                   : launch_tiled<false, false, false, true>(s, v, prev, curr, wp, state, check_done);
        } else {
            if (s->d_aux) le = wts ? launch_tiled<true, true>(s, v, prev, curr, wp, state, check_done)
                                   : launch_tiled<true, false>(s, v, prev, curr, wp, state, check_done);
            else          le = wts ? launch_tiled<false, true>(s, v, prev, curr, wp, state, check_done)
                                   : launch_tiled<false, false>(s, v, prev, curr, wp, state, check_done);
        }
        if (le != cudaSuccess) return le;
        s->counters[0] += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (t.n_tiles > 0 && t.n_fallback <= kFoldFallbackMax) return cudaSuccess;   // swept inside the tiled kernel
    return enqueue_rowgroup(s, t.fallback, t.n_fallback, prev, curr, wts, state, check_done);
}

namespace oar {
cudaError_t sweep_enqueue(oar_store *s, const double *prev, double *curr, const OarEmState *state, int check_done)
{ return enqueue_sweep(s, prev, curr, nullptr, state, check_done); }

// The same sweep over the tiles listed in tile_list[0 .. *n_active) (device memory) only; rows outside the tiled
// layout are swept every time.  Needs the tiled layout.
cudaError_t sweep_enqueue_list(oar_store *s, const double *prev, double *curr, const OarEmState *state, int check_done,
                               const uint32_t *tile_list, const uint32_t *n_active)
{
    const TiledLayout &t = s->tl;
    if (s->kernel != OAR_KERNEL_TILED || !t.ready || t.n_tiles == 0) return enqueue_sweep(s, prev, curr, nullptr, state, check_done);
    tiled::View v = tiled_view(s);
    v.tile_list = tile_list; v.n_active = n_active;
    cudaError_t le = s->d_aux ? launch_tiled<true, false, true>(s, v, prev, curr, nullptr, state, check_done)
                              : launch_tiled<false, false, true>(s, v, prev, curr, nullptr, state, check_done);
    if (le != cudaSuccess) return le;
    s->counters[0] += 1;
    if (t.n_fallback <= kFoldFallbackMax) return cudaSuccess;
    return enqueue_rowgroup(s, t.fallback, t.n_fallback, prev, curr, nullptr, state, check_done);
}
}  // namespace oar

// device word set by tiled::lane_weights when a weight does not fit 16 bits (next to the validation flags of store creation)
static uint32_t *weight_flag(oar_store *s) { return reinterpret_cast<uint32_t *>(s->d_state + 1) + 2; }

// 0 if the weights staged since the last reset fit the tiled sweep; enqueues nothing when the CSR kernel is active
static int check_weight_flag(oar_store *s, const char *who)
{
    if (s->kernel != OAR_KERNEL_TILED || s->tl.n_tiles == 0) return OAR_OK;
    uint32_t *h = reinterpret_cast<uint32_t *>(&s->h_state[oar::kHostStateSlots - 1]);
    OAR_CUDA(cudaMemcpyAsync(h, weight_flag(s), sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    if (*h) return fail(OAR_ERR_UNSUPPORTED, std::string(who) + ": a read weight above 65535 does not fit the tiled sweep (OAR_TILED=0 selects the CSR kernel)");
    return OAR_OK;
}

// Bring the tile-order copies of the bootstrap weights up to date.
static cudaError_t refresh_wperm(oar_store *s, const uint32_t *wts)
{
    const TiledLayout &t = s->tl;
    if (s->kernel != OAR_KERNEL_TILED || t.n_tiled_rows == 0) return cudaSuccess;
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((t.n_tiled_rows + threads - 1) / threads, (uint64_t)s->sm_count * 16);
    tiled::permute_weights<<<blocks, threads, 0, s->stream>>>(wts, t.trow, t.n_tiled_rows, t.wperm);
    s->counters[0] += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || t.n_tiles == 0) return e;
    if (!s->tl.wlane) {
        e = dmalloc(&s->tl.wlane, sizeof(uint16_t) * (size_t)t.n_tiles * tiled::kThreads, s->stream);
        if (e != cudaSuccess) return e;
    }
    const int blocks2 = (int)std::min<uint32_t>(t.n_tiles, (uint32_t)s->sm_count * 8);
    tiled::lane_weights<<<blocks2, tiled::kThreads, 0, s->stream>>>(t.rec, t.records, t.wperm, t.n_tiles, s->tl.wlane, weight_flag(s));
    s->counters[0] += 1;
    return cudaGetLastError();
}

static cudaError_t enqueue_update(oar_store *s, double *prev, const double *curr)
{
    const int threads = 256;
    const int blocks = (int)std::max<uint32_t>(1, std::min<uint32_t>((s->n_txps + threads * 4 - 1) / (threads * 4), (uint32_t)s->sm_count * 4));
    kern::em_update<<<blocks, threads, 0, s->stream>>>(prev, curr, s->n_txps, s->d_state);
    s->counters[0] += 1;
    return cudaGetLastError();
}

// Build (once per store and variant) the CUDA graph of kGraphIters iterations:
// even iterations sweep buf0 -> buf1, odd ones buf1 -> buf0.
static int ensure_graph(oar_store *s, bool weighted)
{
    GraphSlot &g = s->graphs[weighted ? 1 : 0];
    if (g.exec && g.kernel == s->kernel && g.fused == fused_update(s, weighted)) return OAR_OK;
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    const uint32_t *wts = weighted ? s->d_weights : nullptr;
    uint64_t saved = s->counters[0];
    cudaGraph_t graph = nullptr;
    OAR_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    const bool fused = fused_update(s, weighted);
    for (int it = 0; it < kGraphIters && e == cudaSuccess; ++it) {
        // sweep k: X[k % 3] -> X[(k + 1) % 3].  Fused: its head judges sweep k-1 (X[(k + 2) % 3] against X[k % 3]) and zeroes
        // X[(k + 2) % 3], the target of sweep k+1.  Otherwise em_update judges sweep k right after it and zeroes X[k % 3].
        // (Tried in round 2: em_update on a side branch of the graph, next to sweep k+1 -- no faster than behind sweep k,
        // 5 350 it/s either way on C3: the update is not what the iteration waits for.)
        double *prev = s->d_counts[it % 3], *curr = s->d_counts[(it + 1) % 3];
        e = enqueue_sweep(s, prev, curr, wts, s->d_state, 1, fused);
        if (e == cudaSuccess && !fused) e = enqueue_update(s, prev, curr);
    }
    cudaError_t e2 = cudaStreamEndCapture(s->stream, &graph);
    s->counters[0] = saved;
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(e, "graph capture (launch)"); }
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { g.exec = nullptr; return cuda_fail(e, "cudaGraphInstantiate"); }
    g.kernel = s->kernel; g.fused = fused;
    return OAR_OK;
}

// One complete EM (do_em, em.rs:144-255) on the store's stream.  `init_dev`
// (device, M) or null for uniform.  Result is left in *result_buf (device).
static int run_em(oar_store *s, const double *init_dev, uint32_t max_iter, double thr, uint32_t min_iter,
                  bool weighted, double **result_buf, uint32_t *out_niter, double *out_rel, uint32_t *out_sweeps)
{
    const uint32_t M = s->n_txps;
    const uint32_t *wts = weighted ? s->d_weights : nullptr;
    // state
    OarEmState init_state;
    memset(&init_state, 0, sizeof(init_state));
    init_state.conv_thresh = thr; init_state.max_iter = max_iter; init_state.min_iter = min_iter;
    init_state.done = (max_iter == 0) ? 1u : 0u;
    init_state.n_txps = M;
    for (int i = 0; i < 3; ++i) init_state.bufs[i] = s->d_counts[i];
    s->h_state[3] = init_state;
    OAR_CUDA(cudaMemcpyAsync(s->d_state, &s->h_state[3], sizeof(OarEmState), cudaMemcpyHostToDevice, s->stream));
    // prev = init or N/M (em.rs:160-167); curr = 0 (em.rs:158)
    {
        const int threads = 256;
        const int blocks = (int)std::max<uint32_t>(1, std::min<uint32_t>((M + threads - 1) / threads, (uint32_t)s->sm_count * 8));
        const double avg = (double)s->n_reads / (double)M;
        kern::em_init<<<blocks, threads, 0, s->stream>>>(s->d_counts[0], s->d_counts[1], s->d_counts[2], init_dev, avg, M);
        s->counters[0] += 1;
        OAR_CUDA(cudaGetLastError());
    }
    uint32_t sweeps = 0, niter = 0;
    double rel = 0.0;
    if (max_iter > 0) {
        int rc = ensure_graph(s, weighted);
        if (rc != OAR_OK) return rc;
        cudaGraphExec_t exec = s->graphs[weighted ? 1 : 0].exec;
        // keep two graph launches in flight; poll the state copied after each
        int inflight = 0, head = 0;
        bool done = false;
        uint64_t launched_iters = 0;
        while (!done) {
            while (inflight < 2) {
                int slot = (head + inflight) & 1;
                OAR_CUDA(cudaGraphLaunch(exec, s->stream));
                launched_iters += kGraphIters;
                OAR_CUDA(cudaMemcpyAsync(&s->h_state[slot], s->d_state, sizeof(OarEmState), cudaMemcpyDeviceToHost, s->stream));
                OAR_CUDA(cudaEventRecord(s->slot_ev[slot], s->stream));
                ++inflight;
            }
            OAR_CUDA(cudaEventSynchronize(s->slot_ev[head]));
            const OarEmState &hs = s->h_state[head];
            if (s->progress) s->progress(hs.niter, hs.last_rel, s->progress_user);
            if (hs.done) { done = true; sweeps = hs.sweeps; niter = hs.niter; rel = hs.last_rel; }
            head ^= 1; --inflight;
        }
        // the launches still in flight are no-ops (done is set); they finish before the final sweep (same stream)
        // every kernel node of every graph launch is a launch of ours (those after convergence exit at once)
        const uint64_t per_iter = (fused_update(s, weighted) ? 1 : 2) + ((s->kernel != OAR_KERNEL_ROWGROUP && s->tl.n_tiles > 0 && s->tl.n_fallback > kFoldFallbackMax) ? 1 : 0);
        s->counters[0] += launched_iters * per_iter;
    }
    // `sweeps` loop sweeps were judged: the last result is X[sweeps % 3]; X[(sweeps + 2) % 3] is zero (zeroed by the head of
    // the sweep after it, or by em_update / em_init), whatever a speculative sweep wrote into X[(sweeps + 1) % 3]
    double *prev = s->d_counts[sweeps % 3], *curr = s->d_counts[(sweeps + 2) % 3];
    {
        const int threads = 256;
        const int blocks = (int)std::max<uint32_t>(1, std::min<uint32_t>((M + threads - 1) / threads, (uint32_t)s->sm_count * 8));
        kern::em_threshold<<<blocks, threads, 0, s->stream>>>(prev, M);  // em.rs:238-242
        s->counters[0] += 1;
        OAR_CUDA(cudaGetLastError());
    }
    OAR_CUDA(enqueue_sweep(s, prev, curr, wts, s->d_state, 0));         // em.rs:245-252
    s->counters[1] += sweeps + 1;
    *result_buf = curr;
    if (out_niter) *out_niter = niter;
    if (out_rel) *out_rel = rel;
    if (out_sweeps) *out_sweeps = sweeps;
    return OAR_OK;
}

static int stage_init(oar_store *s, const double *init_or_null, double **d_init)
{
    *d_init = nullptr;
    if (!init_or_null) return OAR_OK;
    OAR_CUDA(dmalloc(d_init, sizeof(double) * s->n_txps, s->stream));
    OAR_CUDA(cudaMemcpyAsync(*d_init, init_or_null, sizeof(double) * s->n_txps, cudaMemcpyDefault, s->stream));
    return OAR_OK;
}

extern "C" int oar_em(oar_store *s, const double *init_or_null, uint32_t max_iter, double conv_thresh,
                      uint32_t min_iter, double *out_counts, uint32_t *out_niter, double *out_rel_diff)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_em: store is null");
    if (!out_counts) return fail(OAR_ERR_INVALID, "oar_em: out_counts is null");
    OAR_CUDA(cudaSetDevice(s->device));
    s->counters[0] = s->counters[1] = 0;
    double *d_init = nullptr;
    int rc = stage_init(s, init_or_null, &d_init);
    if (rc != OAR_OK) { dfree(d_init, s->stream); return rc; }
    OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
    double *res = nullptr;
    rc = run_em(s, d_init, max_iter, conv_thresh, min_iter, false, &res, out_niter, out_rel_diff, nullptr);
    if (rc != OAR_OK) { dfree(d_init, s->stream); cudaStreamSynchronize(s->stream); return rc; }
    OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
    OAR_CUDA(cudaMemcpyAsync(out_counts, res, sizeof(double) * s->n_txps, cudaMemcpyDefault, s->stream));
    OAR_CUDA(cudaEventRecord(s->ev[2], s->stream));
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    dfree(d_init, s->stream);
    float a = 0.f, b = 0.f;
    OAR_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    OAR_CUDA(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    s->timings[1] = a; s->timings[2] = b; s->timings[3] = 0;
    return OAR_OK;
}

// ---------------------------------------------------------------------------
// bootstrap
// ---------------------------------------------------------------------------

static int ensure_weights(oar_store *s)
{
    if (!s->d_weights) OAR_CUDA(dmalloc(&s->d_weights, sizeof(uint32_t) * (s->n_reads + 16), s->stream));
    return OAR_OK;
}

static int enqueue_sample_weights(oar_store *s, uint64_t seed, uint32_t replicate, uint32_t *d_w)
{
    OAR_CUDA(cudaMemsetAsync(d_w, 0, sizeof(uint32_t) * s->n_reads, s->stream));
    if (s->n_reads == 0) return OAR_OK;
    const int threads = 256;
    const uint64_t nthreads_needed = (s->n_reads + 1) / 2;  // two draws per Philox block
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((nthreads_needed + threads - 1) / threads, (uint64_t)s->sm_count * 16));
    kern::boot_weights_kernel<<<blocks, threads, 0, s->stream>>>(d_w, s->n_reads, seed, replicate);
    s->counters[0] += 1;
    OAR_CUDA(cudaGetLastError());
    return OAR_OK;
}

extern "C" int oar_bootstrap_sample_weights(oar_store *s, uint64_t seed, uint32_t replicate, uint32_t *out_weights)
{
    if (!s || !out_weights) return fail(OAR_ERR_INVALID, "oar_bootstrap_sample_weights: null argument");
    OAR_CUDA(cudaSetDevice(s->device));
    int rc = ensure_weights(s);
    if (rc != OAR_OK) return rc;
    rc = enqueue_sample_weights(s, seed, replicate, s->d_weights);
    if (rc != OAR_OK) return rc;
    OAR_CUDA(cudaMemcpyAsync(out_weights, s->d_weights, sizeof(uint32_t) * s->n_reads, cudaMemcpyDefault, s->stream));
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    return OAR_OK;
}

static int bootstrap_impl(oar_store *s, const uint32_t *weights_or_null, uint32_t n_rep, uint64_t seed,
                          uint32_t first, uint32_t stride, uint32_t max_iter, double thr, uint32_t min_iter,
                          double *out, uint32_t *out_niter)
{
    if (!s) return fail(OAR_ERR_INVALID, "oar_bootstrap: store is null");
    if (n_rep > 0 && !out) return fail(OAR_ERR_INVALID, "oar_bootstrap: out is null");
    OAR_CUDA(cudaSetDevice(s->device));
    s->counters[0] = s->counters[1] = 0;
    int rc = ensure_weights(s);
    if (rc != OAR_OK) return rc;
    OAR_CUDA(cudaMemsetAsync(weight_flag(s), 0, sizeof(uint32_t), s->stream));
    OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
    for (uint32_t b = 0; b < n_rep; ++b) {
        if (weights_or_null) {
            OAR_CUDA(cudaMemcpyAsync(s->d_weights, weights_or_null + (uint64_t)b * s->n_reads,
                                     sizeof(uint32_t) * s->n_reads, cudaMemcpyDefault, s->stream));
        } else {
            rc = enqueue_sample_weights(s, seed, first + b * stride, s->d_weights);
            if (rc != OAR_OK) return rc;
        }
        static const bool trace = getenv("OAR_TRACE") != nullptr;   // development: where a replicate's time goes (adds syncs)
        if (trace) OAR_CUDA(cudaEventRecord(s->ev[2], s->stream));
        OAR_CUDA(refresh_wperm(s, s->d_weights));
        if (trace) OAR_CUDA(cudaEventRecord(s->ev[3], s->stream));
        double *res = nullptr;
        uint32_t niter = 0;
        rc = run_em(s, nullptr, max_iter, thr, min_iter, true, &res, &niter, nullptr, nullptr);
        if (rc != OAR_OK) { cudaStreamSynchronize(s->stream); return rc; }
        if (out_niter) out_niter[b] = niter;
        OAR_CUDA(cudaMemcpyAsync(out + (uint64_t)b * s->n_txps, res, sizeof(double) * s->n_txps, cudaMemcpyDefault, s->stream));
        if (trace) {
            OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
            OAR_CUDA(cudaStreamSynchronize(s->stream));
            float w1 = 0.f, w2 = 0.f, em = 0.f;
            cudaEventElapsedTime(&w1, s->ev[0], s->ev[2]); cudaEventElapsedTime(&w2, s->ev[2], s->ev[3]); cudaEventElapsedTime(&em, s->ev[3], s->ev[1]);
            fprintf(stderr, "[oar] replicate %u: draw weights %.3f ms, tile order + lane weights %.3f ms, EM + download %.3f ms (%u iterations: %.1f us each)\n",
                    first + b * stride, w1, w2, em, niter, em * 1e3f / (float)(niter + 1));
            OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
        }
    }
    OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    float a = 0.f;
    OAR_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    s->timings[1] = a; s->timings[2] = 0; s->timings[3] = 0;
    return check_weight_flag(s, "oar_bootstrap");
}

extern "C" int oar_bootstrap(oar_store *s, uint32_t num_boot, uint64_t seed, uint32_t first_replicate,
                             uint32_t replicate_stride, uint32_t max_iter, double conv_thresh,
                             double *out, uint32_t *out_niter)
{
    // do_bootstrap runs do_em, whose stop rule is niter > 50 (em.rs:212, :287-289)
    return bootstrap_impl(s, nullptr, num_boot, seed, first_replicate, replicate_stride ? replicate_stride : 1,
                          max_iter, conv_thresh, 50, out, out_niter);
}

extern "C" int oar_bootstrap_weights(oar_store *s, const uint32_t *weights, uint32_t n_replicates,
                                     uint32_t max_iter, double conv_thresh, uint32_t min_iter,
                                     double *out, uint32_t *out_niter)
{
    if (n_replicates > 0 && !weights) return fail(OAR_ERR_INVALID, "oar_bootstrap_weights: weights is null");
    return bootstrap_impl(s, weights, n_replicates, 0, 0, 1, max_iter, conv_thresh, min_iter, out, out_niter);
}

// ---------------------------------------------------------------------------
// raw sweep (measurement / tests)
// ---------------------------------------------------------------------------

extern "C" int oar_sweep(oar_store *s, const double *prev_dev, double *curr_dev,
                         const uint32_t *weights_or_null, int sync)
{
    if (!s || !prev_dev || !curr_dev) return fail(OAR_ERR_INVALID, "oar_sweep: null argument");
    OAR_CUDA(cudaSetDevice(s->device));
    OAR_CUDA(cudaMemsetAsync(curr_dev, 0, sizeof(double) * s->n_txps, s->stream));
    if (weights_or_null) {
        OAR_CUDA(cudaMemsetAsync(weight_flag(s), 0, sizeof(uint32_t), s->stream));
        OAR_CUDA(refresh_wperm(s, weights_or_null));
    }
    OAR_CUDA(enqueue_sweep(s, prev_dev, curr_dev, weights_or_null, s->d_state, 0));
    if (weights_or_null && sync) return check_weight_flag(s, "oar_sweep");
    if (sync) OAR_CUDA(cudaStreamSynchronize(s->stream));
    return OAR_OK;
}

extern "C" int oar_sweep_timed(oar_store *s, const double *prev_dev, double *curr_dev,
                               const uint32_t *weights_or_null, int reps, float *out_ms_total)
{
    if (!s || !prev_dev || !curr_dev || !out_ms_total || reps <= 0)
        return fail(OAR_ERR_INVALID, "oar_sweep_timed: bad argument");
    OAR_CUDA(cudaSetDevice(s->device));
    if (weights_or_null) {
        OAR_CUDA(cudaMemsetAsync(weight_flag(s), 0, sizeof(uint32_t), s->stream));
        OAR_CUDA(refresh_wperm(s, weights_or_null));
    }
    OAR_CUDA(cudaMemsetAsync(curr_dev, 0, sizeof(double) * s->n_txps, s->stream));
    // development (OAR_TIMED_MODE): the same sweeps with the EM's early-exit test ("done"), as nodes of a CUDA graph ("graph"), or both
    static const char *tmode = getenv("OAR_TIMED_MODE");
    const int check_done = tmode && strstr(tmode, "done") ? 1 : 0;
    if (check_done) OAR_CUDA(cudaMemsetAsync(s->d_state, 0, sizeof(OarEmState), s->stream));
    if (tmode && strstr(tmode, "graph")) {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        OAR_CUDA(cudaStreamSynchronize(s->stream));
        OAR_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kGraphIters; ++i) OAR_CUDA(enqueue_sweep(s, prev_dev, curr_dev, weights_or_null, s->d_state, check_done));
        OAR_CUDA(cudaStreamEndCapture(s->stream, &graph));
        OAR_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        OAR_CUDA(cudaGraphLaunch(exec, s->stream));
        OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
        const int launches = (reps + kGraphIters - 1) / kGraphIters;
        for (int i = 0; i < launches; ++i) OAR_CUDA(cudaGraphLaunch(exec, s->stream));
        OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
        OAR_CUDA(cudaStreamSynchronize(s->stream));
        OAR_CUDA(cudaEventElapsedTime(out_ms_total, s->ev[0], s->ev[1]));
        *out_ms_total *= (float)reps / (float)(launches * kGraphIters);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        return OAR_OK;
    }
    OAR_CUDA(cudaEventRecord(s->ev[0], s->stream));
    for (int i = 0; i < reps; ++i) OAR_CUDA(enqueue_sweep(s, prev_dev, curr_dev, weights_or_null, s->d_state, check_done));
    OAR_CUDA(cudaEventRecord(s->ev[1], s->stream));
    OAR_CUDA(cudaStreamSynchronize(s->stream));
    OAR_CUDA(cudaEventElapsedTime(out_ms_total, s->ev[0], s->ev[1]));
    return weights_or_null ? check_weight_flag(s, "oar_sweep_timed") : OAR_OK;
}
