// oar_common.cuh -- shared definitions for the oarfish EM engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/oarfish_em.h"

// Constants of the reference EM (src/util/constants.rs:1-2).
#define OAR_MIN_READ_THRESH 1e-5
#define OAR_EM_DENOM_THRESH 1e-30

// Device-resident state of one EM run.  The convergence decision of the
// reference (src/em.rs:194-218) is taken on the device by the last CTA of the
// update kernel, so a CUDA graph of many iterations can run without host
// round-trips: once `done` is set, the remaining kernels of the graph return
// immediately.
struct OarEmState {
    unsigned long long rel_bits;  // max signed rel-diff of this iteration, as the bits of a non-negative f64
    double last_rel;              // rel_diff evaluated by the most recent update
    double conv_thresh;
    uint32_t niter;               // the reference's loop counter (em.rs:170)
    uint32_t sweeps;              // m_step calls completed inside the loop
    uint32_t done;
    uint32_t ticket;
    uint32_t max_iter;
    uint32_t min_iter;
    uint32_t primed;              // fused update: the first sweep of an EM has no predecessor to judge
    uint32_t n_txps;              // fused update: length of the count vectors
    double *bufs[3];              // fused update: the three rotating count vectors
};

namespace oar {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

// Stream-ordered allocation from the device's default memory pool (kept warm: freed blocks are
// cached by the pool, so creating / destroying stores repeatedly does not pay cudaMalloc/cudaFree).
template <typename T>
inline cudaError_t dmalloc(T **p, size_t bytes, cudaStream_t st)
{ return cudaMallocAsync(reinterpret_cast<void **>(p), bytes ? bytes : 16, st); }
inline void dfree(void *p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }
void warm_pool(int device);

#define OAR_CUDA(expr)                                              \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ::oar::cuda_fail(_e, #expr);  \
    } while (0)

}  // namespace oar
