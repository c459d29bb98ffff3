// pair_walk.cuh — the lane-serial pair loop shared by the dense rate kernel (rates_dense.cu) and the
// neighbour-rate refresh (rates_refresh.cu): a lane owns one site and walks that site's pair mask in slot order
// with the running sum in a register.  The per-pair arithmetic is att_pair_rate_E / diff_pair_rate of
// site_rates.cuh with the range check of fast_exp and the `isfinite` half of keep_rate hoisted out of the loop
// body; the bits are the same (tests/test_gpu_sweep.py compares every variant).
#pragma once
#include <stdint.h>
#include "site_rates.cuh"

namespace cet {

template <int OFF>
__device__ __forceinline__ unsigned lds_u8(uint32_t addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// bits 4*o of a nibble mask -> bits 2*o
__device__ __forceinline__ uint32_t nib_compress2(uint32_t x)
{
    x = (x | (x >> 2)) & 0x05050505u;
    x = (x | (x >> 4)) & 0x00550055u;
    return (x | (x >> 8)) & 0x5555u;
}
// Walk order of a site's pair mask: slot o at bit 30 - 2*o, so the leading set bit is the lowest slot and its
// position is the BYTE offset of the slot's entry in the reversed 16-bit offset table (DenseSmem::dpb).
__device__ __forceinline__ uint32_t pair_walk_mask(uint64_t pm)
{
    return __brev((nib_compress2((uint32_t)pm) | (nib_compress2((uint32_t)(pm >> 32)) << 16)) << 1);
}

// `sum += keep_rate(P, rate)` without the selects: rates are never negative and sum never -0 or NaN, so adding the
// kept rate or nothing gives the same bits; for rate > threshold, `rate < inf` is a test of the high word.
__device__ __forceinline__ void add_kept(const cet_rate_params &P, double &sum, double rate)
{
    asm("{\n"
        ".reg .pred p;\n"
        "setp.gt.f64 p, %1, %2;\n"
        "setp.lt.and.s32 p, %3, 0x7ff00000, p;\n"
        "@p add.rn.f64 %0, %0, %1;\n"
        "}\n"
        : "+d"(sum)
        : "d"(rate), "d"(P.rate_threshold), "r"(__double2hiint(rate)));
}


// One attachment / diffusion pair with operand op (kmc_event_rates.py:147-157 / :102-108), before the threshold
// filter.  xmax collects the largest |Arrhenius argument| (high word) for the range check after the loop.
template <bool ATT>
__device__ __forceinline__ double pair_rate_raw(const cet_rate_params &P, const double *exp_tab, double op, double A, double B, int &xmax)
{
    if (ATT) {
        const double x = -op * A;
        xmax = max(xmax, __double2hiint(x) & 0x7fffffff);
        return fast_exp_core(x, exp_tab) * B;
    }
    const double neighbor_T = pymax(op, 1.0);
    const double dT = fabs(A - neighbor_T);
    const double denom = pymax(P.T_melt - neighbor_T, 1.0);
    return fma(0.1 * dT, rcp1(denom), 1.0) * B;
}
constexpr int EXP_RANGE_HI = 0x4085e000;     // high word of 700.0: xmax >= this <=> !(|x| < 700) for some pair, NaN included

}  // namespace cet
