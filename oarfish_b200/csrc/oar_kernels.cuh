// oar_kernels.cuh -- sm_100a kernels of the oarfish EM engine.
#pragma once
#include "oar_common.cuh"

namespace oar {
namespace kern {

// ---------------------------------------------------------------------------
// store validation / narrowing
// ---------------------------------------------------------------------------

// boundaries (Vec<usize>, oarfish_types.rs:555) -> u32 row_ptr; flag[0] != 0 if
// not a monotone prefix array that starts at 0 and ends at nnz.
// `base` = value of the first boundary (0 for a whole store; a slice of a larger store -- one device's cells in
// oar_em_batched_multi -- starts at its first alignment's offset): it is subtracted while narrowing.
static __global__ void narrow_validate_rowptr(const uint64_t *__restrict__ rp64, uint32_t *__restrict__ rp32,
                                       uint64_t n_reads, uint64_t nnz, uint64_t base, uint32_t *flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_reads; i += stride) {
        const uint64_t v = rp64[i] - base;
        if (i == 0 && v != 0) bad = true;
        if (i == n_reads && v != nnz) bad = true;
        if (i < n_reads && rp64[i + 1] - base < v) bad = true;
        if (v > nnz) bad = true;
        rp32[i] = (uint32_t)v;
    }
    if (bad) atomicOr(flag, 1u);
}

static __global__ void validate_txp(const uint32_t *__restrict__ txp, uint64_t nnz, uint32_t n_txps, uint32_t *flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride)
        if (txp[i] >= n_txps) bad = true;
    if (bad) atomicOr(flag, 1u);
}

// ---------------------------------------------------------------------------
// EM bookkeeping kernels
// ---------------------------------------------------------------------------

// prev = init or avg (em.rs:160-167); curr = 0 (em.rs:158)
static __global__ void em_init(double *__restrict__ prev, double *__restrict__ curr, double *__restrict__ third,
                        const double *__restrict__ init, double avg, uint32_t M)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
        prev[i] = init ? init[i] : avg;
        curr[i] = 0.0;
        third[i] = 0.0;
    }
}

// set very small abundances to 0 (em.rs:238-242)
static __global__ void em_threshold(double *__restrict__ prev, uint32_t M)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride)
        if (prev[i] < OAR_MIN_READ_THRESH) prev[i] = 0.0;
}

// The stop rule of do_em / em_par (em.rs:212-218, :399, :181), applied by whoever holds the last ticket of an
// iteration's rel-diff reduction: consumes rel_bits, advances niter / sweeps, sets done.
__device__ __forceinline__ void em_decide(OarEmState *st)
{
    const unsigned long long bits = atomicAdd(&st->rel_bits, 0ull);
    const double rel = __longlong_as_double((long long)bits);
    st->last_rel = rel;
    st->sweeps += 1;
    if (rel < st->conv_thresh && st->niter > st->min_iter) {
        st->done = 1;                       // break (em.rs:212-214)
    } else {
        st->niter += 1;                     // em.rs:218
        if (st->niter >= st->max_iter) st->done = 1;  // while niter < max_iter (em.rs:181)
    }
    st->rel_bits = 0ull;                    // em.rs:234
    st->ticket = 0;
    __threadfence();
}

// After a sweep prev -> curr:  rel_diff = max_i{(curr_i - prev_i)/prev_i : prev_i > 1e-5}
// (signed, starts from 0; em.rs:194-201), then "swap + fill(0)" (em.rs:204-207)
// == zero the old prev, which is the next sweep's target.  The last CTA applies
// the stop rule (em.rs:212 / :399) and advances niter (em.rs:218).
static __global__ void __launch_bounds__(256) em_update(double *__restrict__ prev, const double *__restrict__ curr,
                                                 uint32_t M, OarEmState *st)
{
    if (st->done) return;
    double m = 0.0;
    const uint32_t stride = gridDim.x * blockDim.x;
#ifdef OAR_UPDATE_FAKE   // timing experiment only (wrong results): the launch and the ticket without the pass over the counts
    M = 0;
#endif
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
        const double pc = prev[i];
        const double cc = curr[i];
        if (pc > OAR_MIN_READ_THRESH) {
            const double rd = (cc - pc) / pc;
            m = rd > m ? rd : m;
        }
        prev[i] = 0.0;
    }
    // block max (values are >= 0, so the u64 bit pattern is order preserving)
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    __shared__ double wmax[8];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wmax[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b = wmax[w] > b ? wmax[w] : b;
        if (b > 0.0) atomicMax(&st->rel_bits, (unsigned long long)__double_as_longlong(b));
        __threadfence();
        const uint32_t t = atomicAdd(&st->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        em_decide(st);
    }
}

// The same bookkeeping FUSED into the head of the next sweep (em_sweep_tiled): every CTA of sweep k+1 judges its slice
// of iteration k -- old = prev of sweep k, now = its result (the prev of sweep k+1) -- zeroes old (the target of sweep
// k+2; three buffers rotate) and takes a ticket; the last one applies the stop rule.  Sweep k+1 itself runs on: if
// the rule says stop, its output is simply not used (the result is `now`, thresholded, swept once more into the
// buffer zeroed here).  Saves the em_update launch between every two sweeps (6.6 us of 195 on C3).
// `s_wmax`: 33 doubles of shared memory the caller is not using yet (the sweep's x array).
__device__ __forceinline__ void em_update_slice(double *__restrict__ old, const double *__restrict__ now, uint32_t M, OarEmState *st, double *s_wmax)
{
    const uint32_t per = (M + gridDim.x - 1) / gridDim.x;
    const uint32_t b = blockIdx.x * per, e = min(M, b + per);
    double m = 0.0;
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const double pc = old[i];
        const double cc = now[i];
        if (pc > OAR_MIN_READ_THRESH) {
            const double rd = (cc - pc) / pc;
            m = rd > m ? rd : m;
        }
        old[i] = 0.0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_wmax[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double bm = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) bm = s_wmax[w] > bm ? s_wmax[w] : bm;
        if (bm > 0.0) atomicMax(&st->rel_bits, (unsigned long long)__double_as_longlong(bm));
        __threadfence();
        const uint32_t t = atomicAdd(&st->ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            if (st->primed) em_decide(st);
            else { st->primed = 1; st->rel_bits = 0ull; st->ticket = 0; __threadfence(); }   // first sweep of the EM: nothing to judge
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// fused E-step + M-step, variant ROWGROUP: one 8-lane group per read row
// ---------------------------------------------------------------------------
//
// m_step (em.rs:87-133): per row, denom = sum_j prev[t_j]*prob_j[*aux_j]
// (f64), and if denom > 1e-30 every alignment adds w_j/denom to curr[t_j].
// A group caches its first alignment in registers (rows of <= 8 alignments
// never reload), reduces denom with three shuffles and scatters with
// red.global.add.f64.  Bootstrap replicates scale the increment by the row's
// integer resampling weight (== visiting the row that many times,
// oarfish_types.rs:571-598).
template <bool HAS_AUX, bool HAS_WTS>
__device__ __forceinline__ void rowgroup_rows(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ txp,
                                              const float *__restrict__ prob, const double *__restrict__ aux,
                                              const uint32_t *__restrict__ wts, const uint32_t *__restrict__ list,
                                              const double *__restrict__ prev, double *__restrict__ curr,
                                              uint64_t first_group, uint64_t ngroups, uint64_t n_rows)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sub = lane & 7u;
    const unsigned gmask = 0xFFu << (lane & 24u);
    for (uint64_t k = first_group; k < n_rows; k += ngroups) {
        const uint64_t row = list ? (uint64_t)list[k] : k;  // fallback rows of the tiled layout come as a list
        double scale = 1.0;
        if (HAS_WTS) {
            const uint32_t c = wts[row];
            if (c == 0) continue;
            scale = (double)c;
        }
        const uint32_t s = row_ptr[row], e = row_ptr[row + 1];
        const uint32_t j0 = s + sub;
        uint32_t t0 = 0;
        double w0 = 0.0;
        if (j0 < e) {
            t0 = txp[j0];
            w0 = prev[t0] * (double)prob[j0];
            if (HAS_AUX) w0 *= aux[j0];
        }
        double denom = w0;
        for (uint32_t j = j0 + 8; j < e; j += 8) {
            double w = prev[txp[j]] * (double)prob[j];
            if (HAS_AUX) w *= aux[j];
            denom += w;
        }
        denom += __shfl_xor_sync(gmask, denom, 1);
        denom += __shfl_xor_sync(gmask, denom, 2);
        denom += __shfl_xor_sync(gmask, denom, 4);
        if (denom > OAR_EM_DENOM_THRESH) {
            if (j0 < e) {
                double inc = w0 / denom;
                if (HAS_WTS) inc *= scale;
                atomicAdd(curr + t0, inc);
            }
            for (uint32_t j = j0 + 8; j < e; j += 8) {
                const uint32_t t = txp[j];
                double w = prev[t] * (double)prob[j];
                if (HAS_AUX) w *= aux[j];
                double inc = w / denom;
                if (HAS_WTS) inc *= scale;
                atomicAdd(curr + t, inc);
            }
        }
    }
}

template <bool HAS_AUX, bool HAS_WTS>
static __global__ void __launch_bounds__(256) em_sweep_rowgroup(const uint32_t *__restrict__ row_ptr,
                                                         const uint32_t *__restrict__ txp,
                                                         const float *__restrict__ prob,
                                                         const double *__restrict__ aux,
                                                         const uint32_t *__restrict__ wts,
                                                         const uint32_t *__restrict__ list,
                                                         const double *__restrict__ prev,
                                                         double *__restrict__ curr, uint64_t n_rows,
                                                         const OarEmState *__restrict__ st, int check_done)
{
    if (check_done && st->done) return;
    rowgroup_rows<HAS_AUX, HAS_WTS>(row_ptr, txp, prob, aux, wts, list, prev, curr,
                                    ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3,
                                    ((uint64_t)gridDim.x * blockDim.x) >> 3, n_rows);
}

// ---------------------------------------------------------------------------
// bootstrap resampling weights: histogram of N uniform draws from [0, N)
// (bootstrap.rs:7-16; the sort there only orders the visit, the multiset is
// what matters).  Philox4x32-10 keyed by the seed, counter = (draw pair,
// replicate): weights are a pure function of (seed, replicate, N).
// ---------------------------------------------------------------------------

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static __global__ void __launch_bounds__(256) boot_weights_kernel(uint32_t *__restrict__ w, uint64_t n,
                                                           uint64_t seed, uint32_t replicate)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t npairs = (n + 1) / 2;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
        uint32_t r[4];
        philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), replicate, 0x0A2F15B0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        const uint64_t u0 = ((uint64_t)r[1] << 32) | r[0];
        const uint64_t u1 = ((uint64_t)r[3] << 32) | r[2];
        atomicAdd(w + __umul64hi(u0, n), 1u);
        if (2 * p + 1 < n) atomicAdd(w + __umul64hi(u1, n), 1u);
    }
}

}  // namespace kern
}  // namespace oar
