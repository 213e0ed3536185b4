// oar_ctx.cuh -- per-device context shared by all stores of a process.
//
// Why: creating and destroying a store used to create and destroy a stream, six events and a pinned
// host block every time.  On the GPU box cudaFreeHost of that 256-byte block alone took 0.8-400 ms
// (it synchronises the whole context), which was more than half of the end-to-end step the driver
// measured in round 1.  The context keeps those resources in free lists (taken at store creation,
// handed back at destruction, never released before process exit), remembers which kernels already
// have their dynamic shared-memory limit raised, and keeps the device's memory pool warm.  Every
// list is guarded by the context's mutex: distinct handles may be created, used and destroyed from
// different host threads (the single-cell driver of the reference calls em::em from a pool of worker
// threads, single_cell.rs:91-193).
#pragma once
#include <mutex>
#include <unordered_map>
#include <vector>

#include "oar_common.cuh"

namespace oar {

constexpr int kHostStateSlots = 8;   // OarEmState slots in one pinned block (poll slots, staging)

struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    std::mutex mu;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> timing_events, plain_events;
    std::vector<OarEmState *> host_states;
    std::unordered_map<const void *, int> smem_attr;   // kernel -> cudaFuncAttributeMaxDynamicSharedMemorySize granted
    std::unordered_map<const void *, int> smem_static; // kernel -> its static shared memory (cudaFuncGetAttributes)
};

// The context of `device` (current device is switched to it); null + *err on failure.
DeviceCtx *device_ctx(int device, cudaError_t *err);

cudaError_t ctx_take_stream(DeviceCtx *c, cudaStream_t *out);
void ctx_give_stream(DeviceCtx *c, cudaStream_t s);
cudaError_t ctx_take_event(DeviceCtx *c, bool timing, cudaEvent_t *out);
void ctx_give_event(DeviceCtx *c, bool timing, cudaEvent_t e);
cudaError_t ctx_take_host_state(DeviceCtx *c, OarEmState **out);
void ctx_give_host_state(DeviceCtx *c, OarEmState *p);
// Raise the dynamic shared-memory limit of kernel `fn` to at least `bytes` (once per device and size).
// *static_bytes_or_null receives the kernel's static shared-memory size.
cudaError_t ctx_ensure_smem(DeviceCtx *c, const void *fn, int bytes, int *static_bytes_or_null = nullptr);

}  // namespace oar
