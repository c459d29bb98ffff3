// sweep_tile.cu — the fused tile kernel of the synchronous-sublattice sweep.
//
// One pass over the lattice does, per 8 x 8 x 32 tile of sites,
//   refresh : re-evaluate the rate sum of every site whose stamp bit is set (the sites an event of the
//             previous sweep changed, and their neighbours) — or of every site after a thermal step —
//             from a tile of cvox + pairop (tile_state.cuh) staged in shared memory with its halo of 2
//             by two 3-D TMA loads (cp.async.bulk.tensor), double-buffered so that the loads of the
//             next tile run under the arithmetic of the current one;
//   stream  : one fire test per site against its (now current) rate sum, p = 1 - exp(-R tau), with a
//             Philox4x32-10 block per 4 sites keyed by (seed, sweep, global plane, row band, k);
//             fired sites are appended to the sweep's list; per-(plane, tile) partial sums and maxima
//             of R feed the next time increment in a fixed order.
// This replaces three kernels and one cache of the first design (stream, stamp scan, list-driven
// re-evaluation with ~15 gathered DRAM sectors per site, and the 8-byte neighbour-class word the
// gathers maintained): neighbour states and pair operands now come from shared memory, and DRAM
// sees one streaming read of cvox (1 B), pairop (8 B) and the rate sums (8 B) per site.
//
// The pair arithmetic is the per-event code of site_rates.cuh (att_pair_rate_E / diff_pair_rate),
// the per-site half is tile_site_prep (tile_state.cuh), and a site adds its pairs in slot order, so
// the result equals site_rate_sum — and the dense kernel of rates.cu — bit for bit (tested).
#include <cuda.h>
#include <algorithm>
#include "ctx.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "tile_state.cuh"

namespace cet {

int rate_tables_ensure(cet_ctx *c);      // rates.cu

constexpr int TL_PAIRS = 256;            // pair slots per warp (16 sites x 14 pairs fit: a fuller tile runs as two halves)

struct TileWarpSmem {
    double rate[TL_PAIRS];
    double A[32], B[32];
    uint32_t desc[TL_PAIRS];             // staged pairop index of the neighbour | owner lane << 14
};
struct TileSmem {
    double po[2][TL_HI * TL_HJ * TL_PK];         // 128-byte aligned TMA destinations first
    uint8_t vx[2][TL_VBYTES];
    double tab[RT_TABLE_DOUBLES];
    TileWarpSmem w[TL_WARPS];
    uint16_t dlist[TL_SITES];
    int dp[16];                                  // staged pairop offset of neighbour slot o
    unsigned long long bar[2];
    unsigned int n_dirty;
};
static_assert(sizeof(TileSmem) + 1024 <= 227 * 1024, "TileSmem exceeds the shared memory of an SM");
static_assert((TL_PBYTES % 128) == 0 && (TL_VBYTES % 128) == 0, "TMA destinations must stay 128-byte aligned");

enum { TM_ALL = 1, TM_STREAM = 2 };

struct TileArgs {
    const uint8_t *cvox;
    const double *pairop, *T;
    double *site_rate, *dep_rate;
    const uint32_t *stamp;
    SweepState *ss;
    int32_t *fired;
    unsigned int cap_fired;
    double *blk_sum, *blk_max;
    const double *tab;
    cet_rate_params P;
    int L, n0, np, i_off;
    int p_lo, p_hi;              // evaluated local planes
    int top_plane;               // local plane of the global top plane, or -1
    int njb, nkb, n_tiles;
    int mode;
    uint64_t seed;
    uint32_t sweep;
};

// ---- mbarrier / TMA (PTX ISA 8.x; sm_90+) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TL_DONE;\n"
        "bra TL_WAIT;\n"
        "TL_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// The rare exact fire test of a site whose 32-bit digit passed the pre-filter (out of line).
__device__ __noinline__ bool tile_fire_exact(double x, double d, uint64_t seed, uint64_t gsite, uint32_t sweep)
{
    const double p32 = -expm1(-x) * 4294967296.0;
    const double f = floor(p32);
    if (d != f) return d < f;
    double u_rest, unused;                              // leading digit ties: the rest of the uniform decides
    philox_u2(seed, gsite, sweep, STREAM_FIRE_REST, &u_rest, &unused);
    return u_rest < p32 - f;
}

// Evaluate one site per lane (all 32 lanes call).  sv / sp: the staged cvox / pairop tile.
__device__ __forceinline__ void tile_eval(const TileArgs &a, const TileSmem &sm, TileWarpSmem &ws, const uint8_t *sv,
                                          const double *sp, int li, int lj, int lk, int p, int j, int k, bool active)
{
    const cet_rate_params &P = a.P;
    const int lane = threadIdx.x & 31;
    const int vidx = ((li + 2) * TL_HJ + (lj + 2)) * TL_VK + lk + TL_VK0;
    const int pidx = ((li + 2) * TL_HJ + (lj + 2)) * TL_PK + lk + TL_PK0;
    const int s = (p * a.L + j) * a.L + k;
    unsigned c = 0;
    uint32_t wlo = 0, whi = 0;
    double T_self = 1.0, T_m = 1.0, T_p = 1.0;
    if (active) {
        c = sv[vidx];
#pragma unroll
        for (int o = 0; o < 8; ++o)
            wlo |= ((unsigned)sv[vidx + (CET_NB_DI(o) * TL_HJ + CET_NB_DJ(o)) * TL_VK + CET_NB_DK(o)] & 15u) << (4 * o);
#pragma unroll
        for (int o = 8; o < 14; ++o)
            whi |= ((unsigned)sv[vidx + (CET_NB_DI(o) * TL_HJ + CET_NB_DJ(o)) * TL_VK + CET_NB_DK(o)] & 15u) << (4 * (o - 8));
        const unsigned code = c & 15u;
        if (code == TC_EMPTY) {
            T_self = sp[pidx];                                       // an empty site's pairop is its temperature
            T_m = T_self; T_p = T_self;
            if ((wlo | whi) & 0x11111111u) {                         // an occupied neighbour: grad_z is needed (:151-153)
                if (k > 0) T_m = (sv[vidx - 1] & 15u) == TC_EMPTY ? sp[pidx - 1] : a.T[s - 1];
                if (k < a.L - 1) T_p = (sv[vidx + 1] & 15u) == TC_EMPTY ? sp[pidx + 1] : a.T[s + 1];
            }
        } else if ((code & 1u) && code != TC_DEFECT) {
            T_self = a.T[s];
        }
    }
    const uint64_t w = (uint64_t)wlo | ((uint64_t)whi << 32);
    const TilePrep q = tile_site_prep(P, sm.tab, w, active ? c : 0u, T_self, T_m, T_p);
    const uint64_t pm = q.pm;
    const bool is_emp = q.is_emp;
    if (pm) { ws.A[lane] = q.A; ws.B[lane] = q.B; }
    const int cnt = popc64(pm);

    // ---- packed counts: attachment pairs in the low half, diffusion pairs in the high half
    const unsigned mine = is_emp ? (unsigned)cnt : (unsigned)cnt << 16;
    const unsigned all = __reduce_add_sync(0xffffffffu, mine);
    double sum = q.sum0;
    const int npass = all == 0 ? 0 : (((all & 0xFFFFu) + (all >> 16) > (unsigned)TL_PAIRS) ? 2 : 1);
    for (int pass = 0; pass < npass; ++pass) {
        const bool part = npass == 1 || (lane >> 4) == pass;
        const unsigned mine_p = part ? mine : 0u;
        unsigned inc = mine_p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, inc, 31), excl = inc - mine_p;
        const int n_att = (int)(total & 0xFFFFu), n_diff = (int)(total >> 16);
        const int cnt_p = part ? cnt : 0;
        const int start = is_emp ? (int)(excl & 0xFFFFu) : TL_PAIRS - (int)(excl >> 16) - cnt_p;
        if (part) {
            int pos = start;
            unsigned lo = (unsigned)pm, hi = (unsigned)(pm >> 32);
            const unsigned base = ((unsigned)lane << 14) + (unsigned)pidx;
            while (lo) {
                const int b = __ffs(lo) - 1;
                lo &= lo - 1;
                ws.desc[pos++] = base + (unsigned)sm.dp[b >> 2];
            }
            while (hi) {
                const int b = __ffs(hi) - 1;
                hi &= hi - 1;
                ws.desc[pos++] = base + (unsigned)sm.dp[8 + (b >> 2)];
            }
        }
        __syncwarp();
        // ---- pairs, 32 at a time: attachment from the front, diffusion from the back
        const int qd0 = TL_PAIRS - n_diff;
        for (int q0 = lane; q0 < n_att; q0 += 32) {                   // kmc_event_rates.py:135-158
            const unsigned d = ws.desc[q0];
            const int ts = d >> 14;
            ws.rate[q0] = att_pair_rate_E(P, sp[d & 0x3FFFu], ws.A[ts], ws.B[ts], sm.tab + RT_EXP2);
        }
        for (int q0 = lane; q0 < n_diff; q0 += 32) {                  // :100-109
            const unsigned d = ws.desc[qd0 + q0];
            const int ts = d >> 14;
            ws.rate[qd0 + q0] = diff_pair_rate(P, ws.A[ts], ws.B[ts], sp[d & 0x3FFFu]);
        }
        __syncwarp();
        // ---- per-site sums in slot order (the association order of site_rate_sum)
        for (int q0 = 0; q0 < cnt_p; ++q0) sum += ws.rate[start + q0];
        __syncwarp();
    }
    if (active) {
        a.site_rate[s] = sum;
        if (p == a.top_plane) {                                          // deposition (:55-72)
            double dep;
            a.dep_rate[j * a.L + k] = (is_emp && dep_rate(P, T_self, &dep)) ? dep : NAN;
        }
    }
}

template <bool TMA>
__global__ void __launch_bounds__(TL_THREADS, 1)
    sweep_tile_kernel(const __grid_constant__ TileArgs a, const __grid_constant__ CUtensorMap tm_vox,
                      const __grid_constant__ CUtensorMap tm_po)
{
    extern __shared__ unsigned char tile_dyn_smem[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(tile_dyn_smem + ((1024u - (smem_u32(tile_dyn_smem) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int L = a.L;

    for (int q = tid; q < RT_TABLE_DOUBLES; q += TL_THREADS) sm.tab[q] = a.tab[q];
    if (tid < 14) sm.dp[tid] = ((int)c_nb_off[tid][0] * TL_HJ + c_nb_off[tid][1]) * TL_PK + c_nb_off[tid][2];
    if (TMA && tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int tiles_per_iblock = a.njb * a.nkb;
    auto issue = [&](int t, int buf) {                 // one thread: both boxes of tile t into buffer buf
        const int kb = t % a.nkb, jb = (t / a.nkb) % a.njb, ib = t / tiles_per_iblock;
        const int p0 = a.p_lo + TL_I * ib, j0 = TL_J * jb, k0 = TL_K * kb;
        mbar_expect_tx(&sm.bar[buf], (unsigned)(TL_VBYTES + TL_PBYTES));
        tma_load_3d(sm.vx[buf], &tm_vox, &sm.bar[buf], k0 - TL_VK0, j0 - 2, p0 - 2);
        tma_load_3d(sm.po[buf], &tm_po, &sm.bar[buf], k0 - TL_PK0, j0 - 2, p0 - 2);
    };
    if (TMA && tid == 0 && (int)blockIdx.x < a.n_tiles) issue((int)blockIdx.x, 0);

    const double tau = (a.mode & TM_STREAM) ? (a.ss->terminated ? 0.0 : a.ss->tau) : 0.0;
    unsigned int n_refreshed = 0;
    int it = 0;
    for (int t = (int)blockIdx.x; t < a.n_tiles; t += (int)gridDim.x, ++it) {
        const int buf = TMA ? (it & 1) : 0;
        const int kb = t % a.nkb, jb = (t / a.nkb) % a.njb, ib = t / tiles_per_iblock;
        const int p0 = a.p_lo + TL_I * ib, j0 = TL_J * jb, k0 = TL_K * kb;
        if (TMA) {
            if (tid == 0 && t + (int)gridDim.x < a.n_tiles) issue(t + (int)gridDim.x, buf ^ 1);
        } else {
            // cooperative loads (lattices whose row stride TMA cannot describe): zero outside the local array
            for (int q = tid; q < TL_VBYTES; q += TL_THREADS) {
                const int x = q % TL_VK, b = (q / TL_VK) % TL_HJ, aa = q / (TL_VK * TL_HJ);
                const int gp = p0 - 2 + aa, gj = j0 - 2 + b, gk = k0 - TL_VK0 + x;
                const bool in = gp >= 0 && gp < a.np && gj >= 0 && gj < L && gk >= 0 && gk < L;
                sm.vx[0][q] = in ? a.cvox[((int64_t)gp * L + gj) * L + gk] : (uint8_t)0;
            }
            for (int q = tid; q < TL_HI * TL_HJ * TL_PK; q += TL_THREADS) {
                const int x = q % TL_PK, b = (q / TL_PK) % TL_HJ, aa = q / (TL_PK * TL_HJ);
                const int gp = p0 - 2 + aa, gj = j0 - 2 + b, gk = k0 - TL_PK0 + x;
                const bool in = gp >= 0 && gp < a.np && gj >= 0 && gj < L && gk >= 0 && gk < L;
                sm.po[0][q] = in ? a.pairop[((int64_t)gp * L + gj) * L + gk] : 0.0;
            }
        }
        // ---- the tile's stamped sites (independent of the staged data: runs under the loads)
        const bool all = (a.mode & TM_ALL) != 0;
        if (!all) {
            if (wid == 0) {
                unsigned bits[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = lane + 32 * h, p = p0 + (row >> 3), j = j0 + (row & 7);
                    unsigned b = 0;
                    if (p < a.p_hi && j < L && k0 < L) {
                        const int s0 = (p * L + j) * L + k0;
                        const unsigned w0 = a.stamp[s0 >> 5], w1 = a.stamp[(s0 >> 5) + 1];
                        b = __funnelshift_r(w0, w1, s0 & 31);
                        if (L - k0 < 32) b &= (1u << (L - k0)) - 1u;
                    }
                    bits[h] = b;
                }
                const int mine = __popc(bits[0]) + __popc(bits[1]);
                int inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += u;
                }
                int pos = inc - mine;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    unsigned b = bits[h];
                    const int row = lane + 32 * h;
                    while (b) {
                        const int e = __ffs(b) - 1;
                        b &= b - 1;
                        sm.dlist[pos++] = (uint16_t)(row * TL_K + e);
                    }
                }
                if (lane == 31) sm.n_dirty = (unsigned)inc;
            }
        }
        if (TMA) mbar_wait(&sm.bar[buf], (unsigned)((it >> 1) & 1));
        __syncthreads();
        const uint8_t *sv = sm.vx[buf];
        const double *sp = sm.po[buf];

        // ---- refresh: 32 sites per warp and trip
        const int n_eval = all ? TL_SITES : (int)sm.n_dirty;
        for (int q0 = 32 * wid; q0 < n_eval; q0 += 32 * TL_WARPS) {
            const int q = q0 + lane;
            bool active = q < n_eval;
            const int e = active ? (all ? q : (int)sm.dlist[q]) : 0;
            const int li = e >> 8, lj = (e >> 5) & 7, lk = e & 31;
            const int p = p0 + li, j = j0 + lj, k = k0 + lk;
            if (all) active = active && p < a.p_hi && j < L && k < L;
            tile_eval(a, sm, sm.w[wid], sv, sp, li, lj, lk, p, j, k, active);
        }
        if (!all && tid == 0) n_refreshed += (unsigned)n_eval;
        if (!(a.mode & TM_STREAM)) {
            __syncthreads();                                   // the buffer may be refilled two tiles on
            continue;
        }
        __syncthreads();                                       // refreshed rate sums are visible to the CTA

        // ---- stream: warp = (plane, band of 4 rows), lane = k; one Philox block per lane
        {
            const int li = wid >> 1, band = wid & 1;
            const int p = p0 + li, k = k0 + lane;
            const int jb0 = j0 + 4 * band;
            const bool col = p < a.p_hi && k < L;
            double R[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = jb0 + r;
                R[r] = (col && j < L) ? __ldcg(a.site_rate + ((int64_t)p * L + j) * L + k) : 0.0;
            }
            if (p == a.top_plane && col) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int j = jb0 + r;
                    if (j < L) {
                        const double d = __ldcg(a.dep_rate + j * L + k);
                        if (d == d) R[r] = d + R[r];
                    }
                }
            }
            double rsum = ((R[0] + R[1]) + R[2]) + R[3];
            double rmax = fmax(fmax(R[0], R[1]), fmax(R[2], R[3]));
            unsigned fmask = 0;
            if (tau > 0.0 && rmax > 0.0) {
                const u32x4 rnd = philox4x32_10(u32x4{(uint32_t)(jb0 >> 2) * (uint32_t)L + (uint32_t)k, (uint32_t)(a.i_off + p),
                                                      a.sweep, (uint32_t)STREAM_FIRE_TILE},
                                                (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                const uint32_t words[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
                const double tau32 = tau * 4294967296.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double d = (double)words[r];
                    if (d <= R[r] * tau32) {              // else digit > floor(2^32 x) >= floor(2^32 p): cannot fire
                        const uint64_t gsite = ((uint64_t)(a.i_off + p) * (uint64_t)L + (uint64_t)(jb0 + r)) * (uint64_t)L + (uint64_t)k;
                        if (tile_fire_exact(R[r] * tau, d, a.seed, gsite, a.sweep)) fmask |= 1u << r;
                    }
                }
            }
            rsum = warp_sum(rsum);
            rmax = warp_max(rmax);
            if (lane == 0 && p < a.p_hi) {
                const int64_t slot = ((int64_t)(p - a.p_lo) * tiles_per_iblock + (jb * a.nkb + kb)) * 2 + band;
                a.blk_sum[slot] = rsum;
                a.blk_max[slot] = rmax;
            }
            // fired sites: one list reservation per warp
            const unsigned any = __ballot_sync(0xffffffffu, fmask != 0);
            if (any) {
                const int mine = __popc(fmask);
                int inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += u;
                }
                unsigned int base = 0;
                if (lane == 31) base = atomicAdd(&a.ss->n_fired, (unsigned)inc);
                base = __shfl_sync(0xffffffffu, base, 31);
                unsigned int pos = base + (unsigned)(inc - mine);
                while (fmask) {
                    const int r = __ffs(fmask) - 1;
                    fmask &= fmask - 1;
                    if (pos < a.cap_fired) a.fired[pos] = (p * L + (jb0 + r)) * L + k;
                    else a.ss->overflow = 1;
                    ++pos;
                }
            }
        }
        __syncthreads();                                       // the buffer may be refilled two tiles on
    }
    if (tid == 0 && n_refreshed) atomicAdd(&a.ss->n_dirty, n_refreshed);
}

// ---- resident tile state -----------------------------------------------------------------------
// cvox / pairop of local sites [s_lo, s_hi) from vox, T and the orientation vectors; planes outside the
// global lattice keep cvox = 0.  flag[0] is set when an empty site carries an orientation.
__global__ void tile_state_build_kernel(const uint8_t *__restrict__ vox, const Vec4 *__restrict__ v, const double *__restrict__ T,
                                        uint8_t *cvox, double *pairop, const cet_rate_params P, uint64_t lut, int64_t s_lo,
                                        int64_t s_hi, int *flag)
{
    for (int64_t s = s_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < s_hi; s += (int64_t)gridDim.x * blockDim.x) {
        const unsigned b = vox[s];
        const unsigned code = (unsigned)(lut >> (4 * (b & 15u))) & 15u;
        cvox[s] = (uint8_t)((b & 0xF0u) | code);
        double po = 0.0;
        if (code == TC_EMPTY) {
            const Vec4 u = v[s];
            if (!(u.x == 0.0 && u.y == 0.0 && u.z == 1.0)) *flag = 1;
            po = T[s];
        } else if ((code & 9u) == 9u) {
            po = tile_pairop(P, code, 0.0, v[s].z);
        }
        pairop[s] = po;
    }
}
// after a thermal step: the temperature half of pairop
__global__ void tile_pairop_T_kernel(const uint8_t *__restrict__ cvox, const double *__restrict__ T, double *pairop, int64_t s_lo,
                                     int64_t s_hi)
{
    for (int64_t s = s_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < s_hi; s += (int64_t)gridDim.x * blockDim.x)
        if ((cvox[s] & 15u) == TC_EMPTY) pairop[s] = T[s];
}
// set the stamp bits of local sites [s_lo, s_hi)
__global__ void stamp_fill_kernel(uint32_t *stamp, int64_t s_lo, int64_t s_hi)
{
    const int64_t w_lo = s_lo >> 5, w_hi = (s_hi + 31) >> 5;
    for (int64_t w = w_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w_hi; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = 0xffffffffu;
        if (w == w_lo) m &= 0xffffffffu << (s_lo & 31);
        if (w == w_hi - 1 && (s_hi & 31)) m &= 0xffffffffu >> (32 - (s_hi & 31));
        if (m == 0xffffffffu) stamp[w] = m;
        else atomicOr(&stamp[w], m);
    }
}

int sm_count(cet_ctx *c)
{
    if (c->n_sm <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || n <= 0) n = 148;
        c->n_sm = n;
    }
    return c->n_sm;
}

// local planes that lie inside the global lattice
static void domain_planes(const cet_ctx *c, int *lo, int *hi)
{
    const int i_off = (int)(c->i_begin - c->halo), np = (int)c->np;
    *lo = i_off < 0 ? -i_off : 0;
    *hi = (i_off + np > c->n0) ? (int)(c->n0 - i_off) : np;
}

int tile_state_alloc(cet_ctx *c)
{
    if (!c->cvox) {
        CET_CUDA(cudaMalloc(&c->cvox, (size_t)c->nloc + 64));
        CET_CUDA(cudaMemsetAsync(c->cvox, 0, (size_t)c->nloc + 64, c->stream));
    }
    if (!c->pairop) {
        CET_CUDA(cudaMalloc(&c->pairop, (size_t)c->nloc * sizeof(double)));
        CET_CUDA(cudaMemsetAsync(c->pairop, 0, (size_t)c->nloc * sizeof(double), c->stream));
    }
    if (!c->tile_flag) CET_CUDA(cudaMalloc(&c->tile_flag, 64));
    return 0;
}

// Rebuild cvox / pairop on local planes [p_lo, p_hi) (clipped to the global lattice).
int tile_state_build(cet_ctx *c, int p_lo, int p_hi)
{
    if (int rc = tile_state_alloc(c)) return rc;
    int d_lo, d_hi;
    domain_planes(c, &d_lo, &d_hi);
    if (p_lo < d_lo) p_lo = d_lo;
    if (p_hi > d_hi) p_hi = d_hi;
    if (p_hi <= p_lo) return 0;
    const int64_t s_lo = (int64_t)p_lo * c->plane, s_hi = (int64_t)p_hi * c->plane;
    const int grid = (int)std::min<int64_t>((s_hi - s_lo + 255) / 256, (int64_t)sm_count(c) * 16);
    tile_state_build_kernel<<<grid, 256, 0, c->stream>>>(c->vox, c->v, c->T, c->cvox, c->pairop, c->rp, tile_code_lut(c->rp), s_lo,
                                                         s_hi, c->tile_flag);
    CET_CUDA(cudaGetLastError());
    return 0;
}

// Full rebuild + the check of the orientation invariant (one host read).
int tile_state_ensure(cet_ctx *c)
{
    if (c->tile_valid) return 0;
    if (int rc = tile_state_alloc(c)) return rc;
    CET_CUDA(cudaMemsetAsync(c->tile_flag, 0, sizeof(int), c->stream));
    if (int rc = tile_state_build(c, 0, (int)c->np)) return rc;
    int h = 0;
    CET_CUDA(cudaMemcpyAsync(&h, c->tile_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    c->emp_canonical = h == 0;
    c->tile_valid = true;
    return 0;
}

int tile_pairop_T_update(cet_ctx *c)
{
    int d_lo, d_hi;
    domain_planes(c, &d_lo, &d_hi);
    const int64_t s_lo = (int64_t)d_lo * c->plane, s_hi = (int64_t)d_hi * c->plane;
    const int grid = (int)std::min<int64_t>((s_hi - s_lo + 255) / 256, (int64_t)sm_count(c) * 16);
    tile_pairop_T_kernel<<<grid, 256, 0, c->stream>>>(c->cvox, c->T, c->pairop, s_lo, s_hi);
    CET_CUDA(cudaGetLastError());
    return 0;
}

int stamp_fill(cet_ctx *c, int p_lo, int p_hi)
{
    if (p_hi <= p_lo) return 0;
    const int64_t s_lo = (int64_t)p_lo * c->plane, s_hi = (int64_t)p_hi * c->plane;
    const int64_t nw = ((s_hi + 31) >> 5) - (s_lo >> 5);
    stamp_fill_kernel<<<(int)std::min<int64_t>((nw + 255) / 256, (int64_t)sm_count(c) * 8), 256, 0, c->stream>>>(c->stamp, s_lo,
                                                                                                                   s_hi);
    CET_CUDA(cudaGetLastError());
    return 0;
}

// ---- TMA descriptors -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}

static bool tile_tma_usable(const cet_ctx *c) { return c->n1 % 16 == 0 && c->n1 >= 64 && !(c->debug_flags & 1); }

// Tensor maps over the local arrays (k fastest, then j, then plane); rebuilt when a pointer changed.
static int tile_maps_ensure(cet_ctx *c)
{
    if (c->tmap_vox_ptr == c->cvox && c->tmap_po_ptr == c->pairop) return 0;
    EncodeTiledFn enc = encode_tiled_fn();
    CET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)c->n2, (cuuint64_t)c->n1, (cuuint64_t)c->np};
    const cuuint32_t estr[3] = {1, 1, 1};
    {
        const cuuint64_t strides[2] = {(cuuint64_t)c->n2, (cuuint64_t)c->plane};
        const cuuint32_t box[3] = {TL_VK, TL_HJ, TL_HI};
        CUresult r = enc((CUtensorMap *)c->tmap_vox, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->cvox, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(cvox) failed with CUresult %d", (int)r);
    }
    {
        const cuuint64_t strides[2] = {(cuuint64_t)c->n2 * 8, (cuuint64_t)c->plane * 8};
        const cuuint32_t box[3] = {TL_PK, TL_HJ, TL_HI};
        CUresult r = enc((CUtensorMap *)c->tmap_po, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, c->pairop, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pairop) failed with CUresult %d", (int)r);
    }
    c->tmap_vox_ptr = c->cvox; c->tmap_po_ptr = c->pairop;
    return 0;
}

// Number of (plane, tile, row band) partial sums per plane written by the stream phase.
int tile_parts_per_plane(const cet_ctx *c)
{
    return (int)(((c->n1 + TL_J - 1) / TL_J) * ((c->n2 + TL_K - 1) / TL_K) * 2);
}

// One pass of the fused kernel over local planes [p_lo, p_hi).  mode: TM_ALL re-evaluates every site
// (else the stamped ones), TM_STREAM adds the fire test and the partial sums.
int tile_pass(cet_ctx *c, int p_lo, int p_hi, int mode, uint64_t seed, uint32_t sweep)
{
    if (p_hi <= p_lo) return 0;
    if (int rc = rate_tables_ensure(c)) return rc;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.cvox = c->cvox; a.pairop = c->pairop; a.T = c->T;
    a.site_rate = c->site_rate; a.dep_rate = c->dep_rate; a.stamp = c->stamp;
    a.ss = c->sweep; a.fired = c->fired; a.cap_fired = (unsigned int)c->cap_fired;
    a.blk_sum = c->blk_sum; a.blk_max = c->blk_max; a.tab = c->rate_tab;
    a.P = c->rp;
    a.L = (int)c->n1; a.n0 = (int)c->n0; a.np = (int)c->np; a.i_off = (int)(c->i_begin - c->halo);
    a.p_lo = p_lo; a.p_hi = p_hi;
    const int top = (int)(c->n0 - 1 - (c->i_begin - c->halo));
    a.top_plane = (top >= p_lo && top < p_hi) ? top : -1;
    a.njb = (int)((c->n1 + TL_J - 1) / TL_J); a.nkb = (int)((c->n2 + TL_K - 1) / TL_K);
    a.n_tiles = ((p_hi - p_lo + TL_I - 1) / TL_I) * a.njb * a.nkb;
    a.mode = mode; a.seed = seed; a.sweep = sweep;
    const int grid = std::min(a.n_tiles, sm_count(c));
    const bool tma = tile_tma_usable(c);
    if (!c->tile_attr_set) {
        CET_CUDA(cudaFuncSetAttribute(sweep_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem) + 1024));
        CET_CUDA(cudaFuncSetAttribute(sweep_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem) + 1024));
        c->tile_attr_set = true;
    }
    if (tma) {
        if (int rc = tile_maps_ensure(c)) return rc;
        sweep_tile_kernel<true><<<grid, TL_THREADS, sizeof(TileSmem) + 1024, c->stream>>>(a, *(const CUtensorMap *)c->tmap_vox,
                                                                                  *(const CUtensorMap *)c->tmap_po);
    } else {
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof(dummy));
        sweep_tile_kernel<false><<<grid, TL_THREADS, sizeof(TileSmem) + 1024, c->stream>>>(a, dummy, dummy);
    }
    CET_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cet
