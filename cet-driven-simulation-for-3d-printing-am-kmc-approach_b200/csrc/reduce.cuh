// reduce.cuh — fixed-order reductions and the block-wide "first prefix >= r" search.
//
// Every sum that feeds the BKL search (row sums, plane-segment sums, the total) is formed
// by exactly one routine below, with an association order that depends only on the data
// layout, never on history.  The incremental update after an event re-runs the same
// routine on the affected rows / planes, so the hierarchy is always bit-identical to what
// a full rebuild would produce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "site_rates.cuh"

namespace cet {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;   // commutative adds: every lane holds the same bits
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// max of non-negative, non-NaN doubles: their order is the order of their bit patterns, so two integer warp
// reductions (REDUX) replace the five shuffle + fmax steps
__device__ __forceinline__ double warp_max_nonneg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// Row sums split by occupancy class; lane l adds k = l, l+32, ... in order, then xor tree.
__device__ __forceinline__ void warp_row_sums(const uint8_t *vox_row, const double *rate_row, int n2,
                                              double *occ, double *emp)
{
    const int lane = threadIdx.x & 31;
    double a = 0.0, b = 0.0;
    for (int k = lane; k < n2; k += 32) {
        const double r = rate_row[k];
        if (vox_state(vox_row[k]) != 0) a += r; else b += r;
    }
    *occ = warp_sum(a);
    *emp = warp_sum(b);
}

// Deposition row: sum of existing (non-NaN) dep rates and their count.
__device__ __forceinline__ void warp_dep_row(const double *dep_row, int n2, double *sum, int *cnt)
{
    const int lane = threadIdx.x & 31;
    double a = 0.0;
    int c = 0;
    for (int k = lane; k < n2; k += 32) {
        const double r = dep_row[k];
        if (r == r) { a += r; ++c; }
    }
    *sum = warp_sum(a);
    *cnt = warp_sum_i(c);
}

// Strided warp sum of n doubles (plane-segment sums over the rows of a plane).
__device__ __forceinline__ double warp_strided_sum(const double *x, int n)
{
    const int lane = threadIdx.x & 31;
    double a = 0.0;
    for (int q = lane; q < n; q += 32) a += x[q];
    return warp_sum(a);
}

// Block-wide sum of n doubles, fixed order (thread t adds t, t+B, ...; warp tree; warp 0 tree).
// All threads of the block must call; result returned to every thread.  smem: >= 33 doubles.
__device__ __forceinline__ double block_sum(const double *x, int n, double *smem)
{
    double a = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) a += x[q];
    a = warp_sum(a);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = a;
    __syncthreads();
    if (w == 0) {
        double v = lane < nw ? smem[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) smem[32] = v;
    }
    __syncthreads();
    return smem[32];
}
__device__ __forceinline__ long long block_sum_i(const int32_t *x, int n, long long *smem)
{
    long long a = 0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) a += x[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = a;
    __syncthreads();
    if (w == 0) {
        long long v = lane < nw ? smem[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) smem[32] = v;
    }
    __syncthreads();
    return smem[32];
}

// Scratch for block_find_first (one per block).
struct FindScratch {
    double warp_tot[32];
    double chunk_tot;
    int first, last_pos;
    double first_excl, last_excl;
};

// First index q in [0,n) with v(q) > 0 and base + (v(0)+...+v(q)) >= r — the hierarchical
// counterpart of the reference's linear scan `cumulative += rate; if cumulative >= r`
// (kmc_simulation.py:266-272).  If rounding leaves no such index the last positive entry is
// taken (the reference's events[-1] fallback, :273-274).  Returns the index (or -1 when no
// entry is positive) and *excl = base + sum of the entries before it.  All threads call.
template <class Load>
__device__ __forceinline__ int block_find_first(int n, double base, double r, Load load,
                                                FindScratch *fs, double *excl, bool *clamped)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (threadIdx.x == 0) { fs->first = 0x7fffffff; fs->last_pos = -1; }
    __syncthreads();
    double run = base;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int q = c0 + threadIdx.x;
        const double v = q < n ? load(q) : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) fs->warp_tot[w] = inc;
        __syncthreads();
        if (w == 0) {
            double t = lane < nw ? fs->warp_tot[lane] : 0.0;
            double ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            if (lane < nw) fs->warp_tot[lane] = ti - t;   // exclusive warp offsets
            if (lane == 31) fs->chunk_tot = ti;
        }
        __syncthreads();
        const double ex = run + (fs->warp_tot[w] + (inc - v));
        const double cum = run + (fs->warp_tot[w] + inc);
        if (v > 0.0) {
            if (cum >= r) atomicMin(&fs->first, q);
            atomicMax(&fs->last_pos, q);
        }
        __syncthreads();
        const int first = fs->first;
        if (first != 0x7fffffff) {
            if (q == first) fs->first_excl = ex;
            __syncthreads();
            *excl = fs->first_excl;
            *clamped = false;
            return first;
        }
        if (q == fs->last_pos) fs->last_excl = ex;
        run += fs->chunk_tot;
        __syncthreads();
    }
    *clamped = true;
    *excl = fs->last_pos >= 0 ? fs->last_excl : base;
    return fs->last_pos;
}

}  // namespace cet
