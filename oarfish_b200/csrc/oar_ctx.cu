// oar_ctx.cu -- per-device resource pools (see oar_ctx.cuh).
#include "oar_ctx.cuh"

namespace oar {

namespace {
constexpr int kMaxDevices = 64;
DeviceCtx g_ctx[kMaxDevices];
std::once_flag g_once[kMaxDevices];
cudaError_t g_init_err[kMaxDevices];
}  // namespace

DeviceCtx *device_ctx(int device, cudaError_t *err)
{
    if (device < 0 || device >= kMaxDevices) { if (err) *err = cudaErrorInvalidDevice; return nullptr; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { if (err) *err = e; return nullptr; }
    std::call_once(g_once[device], [device] {
        DeviceCtx &c = g_ctx[device];
        c.device = device;
        cudaError_t e2 = cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device);
        // stream-ordered allocations come from the device's default pool; never give memory back to the
        // driver between stores (creating a store right after destroying one reuses the same blocks)
        cudaMemPool_t pool;
        if (e2 == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        (void)cudaGetLastError();
        g_init_err[device] = e2;
    });
    if (g_init_err[device] != cudaSuccess) { if (err) *err = g_init_err[device]; return nullptr; }
    return &g_ctx[device];
}

cudaError_t ctx_take_stream(DeviceCtx *c, cudaStream_t *out)
{
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->streams.empty()) { *out = c->streams.back(); c->streams.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
void ctx_give_stream(DeviceCtx *c, cudaStream_t s)
{
    if (!s) return;
    std::lock_guard<std::mutex> lk(c->mu);
    c->streams.push_back(s);
}

cudaError_t ctx_take_event(DeviceCtx *c, bool timing, cudaEvent_t *out)
{
    {
        std::lock_guard<std::mutex> lk(c->mu);
        auto &v = timing ? c->timing_events : c->plain_events;
        if (!v.empty()) { *out = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return timing ? cudaEventCreate(out) : cudaEventCreateWithFlags(out, cudaEventDisableTiming);
}
void ctx_give_event(DeviceCtx *c, bool timing, cudaEvent_t e)
{
    if (!e) return;
    std::lock_guard<std::mutex> lk(c->mu);
    (timing ? c->timing_events : c->plain_events).push_back(e);
}

cudaError_t ctx_take_host_state(DeviceCtx *c, OarEmState **out)
{
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->host_states.empty()) { *out = c->host_states.back(); c->host_states.pop_back(); return cudaSuccess; }
    }
    return cudaMallocHost(out, sizeof(OarEmState) * kHostStateSlots);
}
void ctx_give_host_state(DeviceCtx *c, OarEmState *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(c->mu);
    c->host_states.push_back(p);
}

cudaError_t ctx_ensure_smem(DeviceCtx *c, const void *fn, int bytes, int *static_bytes_or_null)
{
    std::lock_guard<std::mutex> lk(c->mu);
    if (static_bytes_or_null) {
        auto st = c->smem_static.find(fn);
        if (st == c->smem_static.end()) {
            cudaFuncAttributes a;
            cudaError_t e = cudaFuncGetAttributes(&a, fn);
            if (e != cudaSuccess) return e;
            st = c->smem_static.emplace(fn, (int)a.sharedSizeBytes).first;
        }
        *static_bytes_or_null = st->second;
    }
    auto it = c->smem_attr.find(fn);
    if (it != c->smem_attr.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) c->smem_attr[fn] = bytes;
    return e;
}

}  // namespace oar
