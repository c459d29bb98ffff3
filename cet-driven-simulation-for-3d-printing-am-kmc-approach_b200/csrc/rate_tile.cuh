// rate_tile.cuh — evaluation of 32 sites by one warp: the core of the dense rate kernel and of the
// neighbour-rate refresh (rates.cu).
//
// Cost model.  The lattice data of a site is 41 B (1 B state, 8 B T, 24 B unit vector in, 8 B rate
// out), but a site next to a solid/empty interface owns up to 14 pair events (diffusion into an
// empty neighbour, attachment of an occupied neighbour) and every attachment costs an fp64 exp: on
// interface-rich lattices the pass is bound by instruction issue / the FP64 pipe, not by HBM.  A
// slot-major loop (14 iterations, lanes without a pair in that slot idle) executed ~1 700 warp
// instructions per 32 sites; this layout executes only the pairs that exist:
//   A  per site, one lane each, coalesced: state byte, cached neighbour-class word (4 bits per
//      neighbour), T.  Shift/AND tests on the word give the pair mask, the bond / impurity counts
//      and the in-bounds count; the site-level Arrhenius factor (nucleation of an empty site,
//      Boltzmann factor of an occupied one) goes through ONE exp call shared by both classes, with
//      K_eff(n_imp, n_in) and E_tot(species, n_bonds) read from shared-memory tables;
//   B  the (site, slot) pairs of the 32 sites are compacted site-major into the warp's slice of
//      shared memory by one packed warp scan (attachment pairs from the front, diffusion pairs
//      from the back) and evaluated 32 at a time, the two classes as separate loops;
//   C  each site adds its own pairs in slot order — the association order of site_rate_sum, so
//      dense rebuild, refresh and the per-event code agree bit for bit.
// Warps are autonomous inside a tile (only __syncwarp): a persistent CTA loads the tables once and
// pulls 2048-site chunks from a global queue (rate_cta_loop).
#pragma once
#if defined(__CUDACC__)
#include "reduce.cuh"
#endif
#include "site_rates.cuh"

namespace cet {

constexpr int RT_THREADS = 256, RT_WARPS = RT_THREADS / 32;
constexpr int RT_CHUNK = 256;                    // sites per queue entry
constexpr int RT_WARP_PAIRS = 256;               // pair slots per warp; a tile with more pairs runs as two half-tiles (<= 16 * 14)

// Cached neighbour-class word of a site: nibble o describes neighbour o,
//   bit 0  occupied (state != 0, inside the lattice)          -> bond count, diffusion targets are the clear ones
//   bit 1  Re, bit 2  C                                       -> impurity count; (nibble >> 1) & 3 = species index
//   bit 3  attachable species (W / Re / C); bit 3 WITHOUT bit 0 marks a neighbour outside the lattice
// bit 56 = the site has k == 0, bit 57 = k == L-1.  The word depends on the state ids of the rate
// parameters (cet_set_rate_params invalidates the cache).
#define CET_NST_K0 (1ull << 56)
#define CET_NST_KTOP (1ull << 57)
CET_HD unsigned nb_code(const cet_rate_params &P, int st)
{
    if (st == 0) return 0u;
    unsigned c = 1u;
    if (st == P.states_w) c |= 8u;
    if (st == P.states_re) c |= 8u | 2u;
    if (st == P.states_c) c |= 8u | 4u;
    return c;
}
// nb_code of the 16 possible state values, 4 bits each
CET_HD uint64_t nb_code_lut(const cet_rate_params &P)
{
    uint64_t lut = 0;
    for (int st = 0; st < 16; ++st) lut |= (uint64_t)nb_code(P, st) << (4 * st);
    return lut;
}

// The 14 state bytes are loaded first, independently of each other (one memory round trip), and
// mapped to classes afterwards.
CET_HD uint64_t nst_word(uint64_t lut, const uint8_t *vox, int64_t s, int i, int j, int k, int n0, int L)
{
    const unsigned inb = inbounds_mask(i, j, k, n0, L);
    unsigned b[14];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int o = 0; o < 14; ++o)
        b[o] = (inb >> o & 1u) ? (unsigned)vox[s + ((int64_t)CET_NB_DI(o) * L + CET_NB_DJ(o)) * L + CET_NB_DK(o)] & 15u : 16u;
    uint64_t w = (k == 0 ? CET_NST_K0 : 0ull) | (k == L - 1 ? CET_NST_KTOP : 0ull);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int o = 0; o < 14; ++o) w |= (b[o] == 16u ? 8ull : ((lut >> (4 * b[o])) & 15ull)) << (4 * o);
    return w;
}

// K_eff[n_imp * 16 + n_in] and E_tot[sp * 16 + n_bonds], built on the device by the same inline
// functions the per-event code calls (rate_tables_kernel), then 2^(j/32) and 0.5 * E_b[ia].
constexpr int RT_KEFF = 0, RT_ETOT = 256, RT_EXP2 = 304, RT_HE = 336, RT_TABLE_DOUBLES = 340;

#if defined(__CUDACC__)

struct RateWarpSmem {
    double A[32], B[32];                     // attachment: 1/(kT T'), nu*gfac; diffusion: T', nu*boltz
    double sx[32], sy[32], sz[32];           // attachment: the empty site's own unit vector
    double pair_rate[RT_WARP_PAIRS];
    int site[32];
    uint16_t pair_desc[RT_WARP_PAIRS];       // slot | species index << 4 | lane << 6
};
struct RateSmem {
    double tab[RT_TABLE_DOUBLES];
    int lin[16];
    unsigned int chunk[2];
    RateWarpSmem w[RT_WARPS];
};

struct RateTileArgs {
    Lat g;
    cet_rate_params P;
    const double *tab;        // RT_TABLE_DOUBLES
    double *site_rate, *dep_rate;
    uint64_t *nst_out;        // GATHER: cache words of the evaluated sites are rewritten
    int top_lo, top_hi;       // local linear index range of the global top plane (empty range if not local)
    int nloc;
    uint64_t lut;             // nb_code_lut(P)
};

// once per CTA; ends with __syncthreads
__device__ __forceinline__ void rate_smem_init(RateSmem &sm, const RateTileArgs &a)
{
    for (int q = threadIdx.x; q < RT_TABLE_DOUBLES; q += blockDim.x) sm.tab[q] = a.tab[q];
    if (threadIdx.x < 14)
        sm.lin[threadIdx.x] = ((int)c_nb_off[threadIdx.x][0] * a.g.L + c_nb_off[threadIdx.x][1]) * a.g.L + c_nb_off[threadIdx.x][2];
    __syncthreads();
}

// All 32 lanes of a warp call; s is this lane's site (local linear index), active false on padding lanes.
template <bool GATHER>
__device__ __forceinline__ void rate_tile(const RateTileArgs &a, const RateSmem &sm, RateWarpSmem &ws, int s, bool active)
{
    const cet_rate_params &P = a.P;
    const int lane = threadIdx.x & 31;

    // ---- A: per-site ------------------------------------------------------------------------
    uint64_t w = 0;
    int st = -1, df = 0;
    double T_self = 1.0, T_m = 1.0, T_p = 1.0;
    if (active) {
        const uint8_t vb = a.g.vox[s];
        T_self = a.g.T[s];
        T_m = a.g.T[s > 0 ? s - 1 : s]; T_p = a.g.T[s + 1 < a.nloc ? s + 1 : s];    // k -+ 1 (replaced below at the row ends)
        st = vb & 15; df = vb >> 4;
        if (GATHER) {
            const int L = a.g.L, LL = L * L;
            const int p = s / LL, r = s - p * LL, j = r / L, k = r - j * L;
            w = nst_word(a.lut, a.g.vox, s, a.g.i_off + p, j, k, a.g.n0, L);
            a.nst_out[s] = w;
        } else {
            w = a.g.nst[s];
        }
    }
    const bool is_emp = st == 0, is_occ = st > 0 && st != P.defect_id;
    const uint64_t w3 = w >> 3;
    const uint64_t m_occ = w & CET_NIB_LSB;                                  // occupied in-bounds neighbours
    const uint64_t m_att = w & w3 & CET_NIB_LSB, m_oob = ~w & w3 & CET_NIB_LSB;
    const double local_T = pymax(T_self, 1.0);
    const double inv_kTT = rcp(P.kT * local_T);
    uint64_t pm = 0;                                                         // pair mask, bit 4*o
    double arg = 0.0;
    bool need_exp = false;
    if (is_occ) {
        pm = ~(w | w3) & CET_NIB_LSB;                                        // empty in-bounds neighbours
        if (pm) {
            arg = occ_exp_arg(df, sm.tab[RT_ETOT + species3(P, st) * 16 + popc64(m_occ)], inv_kTT);
            need_exp = true;
        }
    } else if (is_emp) {
        pm = m_att;
        if (nuc_exists(P, local_T)) {
            const int n_imp = popc64((w >> 1) & m_att) + popc64((w >> 2) & m_att);
            arg = nuc_exp_arg(P, local_T, sm.tab[RT_KEFF + n_imp * 16 + (14 - popc64(m_oob))], inv_kTT);
            need_exp = true;
        }
    }
    const double e = need_exp ? fast_exp_t(arg, sm.tab + RT_EXP2) : 0.0;    // one exp for both site classes
    double sum0 = 0.0;
    if (is_occ && pm) {
        ws.A[lane] = local_T; ws.B[lane] = P.nu * e;
    } else if (is_emp) {
        if (need_exp) sum0 = nuc_from_exp(P, e);
        if (pm) {
            const Vec4 sv = a.g.v[s];
            ws.A[lane] = inv_kTT;
            ws.B[lane] = emp_ng(P, local_T, (w & CET_NST_K0) ? T_self : T_m, (w & CET_NST_KTOP) ? T_self : T_p);
            ws.sx[lane] = sv.x; ws.sy[lane] = sv.y; ws.sz[lane] = sv.z;
        }
    }
    ws.site[lane] = s;
    const int cnt = popc64(pm);

    // ---- packed counts: attachment pairs in the low half, diffusion pairs in the high half
    const unsigned mine = is_emp ? (unsigned)cnt : (unsigned)cnt << 16;
    const unsigned all = __reduce_add_sync(0xffffffffu, mine);
    double sum = sum0;
    const int npass = all == 0 ? 0 : (((all & 0xFFFFu) + (all >> 16) > (unsigned)RT_WARP_PAIRS) ? 2 : 1);
    for (int pass = 0; pass < npass; ++pass) {
        const bool part = npass == 1 || (lane >> 4) == pass;          // two half-tiles when the pairs do not fit
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
        const int start = is_emp ? (int)(excl & 0xFFFFu) : RT_WARP_PAIRS - (int)(excl >> 16) - cnt_p;
        if (part) {
            int pos = start;
            unsigned lo = (unsigned)pm, hi = (unsigned)(pm >> 32);
            const unsigned c_lo = (unsigned)(w >> 1), c_hi = (unsigned)(w >> 33), base = (unsigned)lane << 6;
            while (lo) {
                const int b = __ffs(lo) - 1;
                lo &= lo - 1;
                ws.pair_desc[pos++] = (uint16_t)((unsigned)(b >> 2) | ((c_lo >> b & 3u) << 4) | base);
            }
            while (hi) {
                const int b = __ffs(hi) - 1;
                hi &= hi - 1;
                ws.pair_desc[pos++] = (uint16_t)((unsigned)(8 + (b >> 2)) | ((c_hi >> b & 3u) << 4) | base);
            }
        }
        __syncwarp();

        // ---- B: pairs, one attachment and one diffusion pair per lane and trip; the gathers of the
        // next trip are issued before the arithmetic of the current one
        {
            const int n_trip = ((n_att > n_diff ? n_att : n_diff) + 31) >> 5;
            const int qd0 = RT_WARP_PAIRS - n_diff;
            unsigned da = 0, dd = 0;
            Vec4 nv = Vec4{0.0, 0.0, 1.0, 0.0};
            double Tn = 1.0;
            if (lane < n_att) { da = ws.pair_desc[lane]; nv = a.g.v[ws.site[da >> 6] + sm.lin[da & 15u]]; }      // one 32-byte load
            if (lane < n_diff) { dd = ws.pair_desc[qd0 + lane]; Tn = a.g.T[ws.site[dd >> 6] + sm.lin[dd & 15u]]; }
            for (int t = 0; t < n_trip; ++t) {
                const int q = 32 * t + lane, qn = q + 32;
                const unsigned da_c = da, dd_c = dd;
                const Vec4 nv_c = nv;
                const double Tn_c = Tn;
                if (qn < n_att) { da = ws.pair_desc[qn]; nv = a.g.v[ws.site[da >> 6] + sm.lin[da & 15u]]; }
                if (qn < n_diff) { dd = ws.pair_desc[qd0 + qn]; Tn = a.g.T[ws.site[dd >> 6] + sm.lin[dd & 15u]]; }
                if (q < n_att) {                                         // attachment (kmc_event_rates.py:135-158)
                    const int ts = da_c >> 6;
                    ws.pair_rate[q] = att_pair_rate(P, sm.tab[RT_HE + ((da_c >> 4) & 3u)], ws.A[ts], ws.B[ts], ws.sx[ts],
                                                    ws.sy[ts], ws.sz[ts], nv_c.x, nv_c.y, nv_c.z, sm.tab + RT_EXP2);
                }
                if (q < n_diff) {                                        // diffusion (:100-109)
                    const int ts = dd_c >> 6;
                    ws.pair_rate[qd0 + q] = diff_pair_rate(P, ws.A[ts], ws.B[ts], Tn_c);
                }
            }
        }
        __syncwarp();

        // ---- C: per-site sums in slot order -----------------------------------------------------
        for (int q = 0; q < cnt_p; ++q) sum += ws.pair_rate[start + q];
        __syncwarp();                                                // the slice is reused by the next pass / tile
    }
    if (active) {
        a.site_rate[s] = sum;
        if (s >= a.top_lo && s < a.top_hi) {                             // deposition (:55-72): top plane only
            double dep;
            a.dep_rate[s - a.top_lo] = (is_emp && dep_rate(P, T_self, &dep)) ? dep : NAN;
        }
    }
}

// Queue-driven loop of one CTA over n sites (dense: s = s_lo + index; list != nullptr: s = list[index]).
// A CTA pulls RT_CHUNK * RT_WARPS consecutive sites at a time and deals their 32-site tiles
// round-robin to its warps, so the warps of an SM work on adjacent rows at the same time (their
// neighbour gathers meet in L1) while the queue balances interface-rich against empty regions.
template <bool GATHER>
__device__ __forceinline__ void rate_cta_loop(const RateTileArgs &a, RateSmem &sm, int s_lo, int n, const int32_t *list,
                                              unsigned int *queue)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int CTA_CHUNK = RT_CHUNK * RT_WARPS;
    if (threadIdx.x == 0) sm.chunk[0] = atomicAdd(queue, 1u);
    __syncthreads();
    for (int it = 0;; ++it) {
        const int64_t c0 = (int64_t)sm.chunk[it & 1] * CTA_CHUNK;
        if (c0 >= n) break;
        if (threadIdx.x == 0) sm.chunk[(it + 1) & 1] = atomicAdd(queue, 1u);      // the next chunk, popped ahead of need
        const int hi = (int)(c0 + CTA_CHUNK < n ? c0 + CTA_CHUNK : n);
        for (int q0 = (int)c0 + 32 * wid; q0 < hi; q0 += 32 * RT_WARPS) {
            const bool active = q0 + lane < hi;
            const int s = active ? (list ? list[q0 + lane] : s_lo + q0 + lane) : 0;
            rate_tile<GATHER>(a, sm, sm.w[wid], s, active);
        }
        __syncthreads();
    }
}

#endif  // __CUDACC__

}  // namespace cet
