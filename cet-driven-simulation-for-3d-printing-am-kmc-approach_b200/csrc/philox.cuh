// philox.cuh — counter-based Philox4x32-10 draws of the sweep kernels (sweep.cu, sweep_tile.cu).
// Every draw is keyed by (seed, sweep, GLOBAL site or site group, stream), so a decision does not
// depend on which GPU, CTA or kernel shape evaluates it.
#pragma once
#include <stdint.h>

namespace cet {

struct u32x4 { uint32_t x, y, z, w; };
__device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
// two uniforms in [0,1) with 53 random bits each, keyed by the global site
__device__ __forceinline__ void philox_u2(uint64_t seed, uint64_t site, uint32_t sweep, uint32_t stream,
                                          double *u0, double *u1)
{
    const u32x4 r = philox4x32_10(u32x4{(uint32_t)site, (uint32_t)(site >> 32), sweep, stream},
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    *u0 = (double)(a >> 11) * 1.1102230246251565e-16;
    *u1 = (double)(b >> 11) * 1.1102230246251565e-16;
}
enum { STREAM_FIRE = 0, STREAM_PICK = 1, STREAM_ANGLES = 2, STREAM_SPECIES = 3, STREAM_DEFECT = 4, STREAM_FIRE_REST = 5, STREAM_FIRE_TILE = 6 };

}  // namespace cet
