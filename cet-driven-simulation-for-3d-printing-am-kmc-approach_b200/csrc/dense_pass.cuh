// dense_pass.cuh — the dense evaluation of one lattice row by one warp, shared by the rate
// rebuild (rates.cu) and the sublattice sweep (sweep.cu).
//
// Cost model.  The lattice data of a row is 33 B/site (1 B state, 8 B T, 24 B unit vector), but
// a site next to a solid/empty interface owns up to 14 events and every attachment event costs
// an fp64 exp: on interface-rich lattices the pass is bound by the FP64 pipe, not by HBM.  The
// layout therefore minimises executed fp64 instructions per event while keeping every global
// access coalesced:
//   1. classify: the warp walks the row once and compacts the indices of sites that can own
//      diffusion events (occupied, not a defect) and of empty sites into two shared-memory
//      lists (ballot + popc).  Mixed rows would otherwise run both code paths with half the
//      lanes idle;
//   2. each list is processed 32 sites at a time, slot-major: for neighbour offset o every lane
//      reads its own site's neighbour at the same offset, so consecutive lanes touch ascending
//      addresses of one row (coalesced, served by L1/L2 across the 8 rows of a CTA); slots with
//      no active lane in the warp are skipped with one vote;
//   3. per-site sums are accumulated in slot order by the owning lane — the association order
//      of `site_rate_sum`, so the dense result is bit-identical to the incremental refresh.
#pragma once
#include "reduce.cuh"
#include "site_rates.cuh"

namespace cet {

// Shared-memory workspace of one warp: two index lists of up to L entries.
struct RowLists {
    uint16_t *occ, *emp;
    int n_occ, n_emp;
};

// Pass 1: classify the sites of row (p, j).  zero_row (may be NULL) is a per-warp L-entry
// buffer that is cleared here (defect sites keep rate 0).  With a stamp array only the sites whose
// stamp equals stamp_id are listed (neighbour-rate refresh after a sweep); other(k, st) is then
// called for every selected site that owns no event list entry (defects).
template <class F>
__device__ __forceinline__ void row_classify(const Lat &g, const cet_rate_params &P, int rbase, RowLists &w,
                                             double *zero_row, const uint32_t *stamp, uint32_t stamp_id, F &&other)
{
    const int L = g.L, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int n_occ = 0, n_emp = 0;
    for (int k0 = 0; k0 < L; k0 += 32) {
        const int k = k0 + lane;
        bool sel = k < L;
        if (stamp != nullptr && sel) sel = stamp[rbase + k] == stamp_id;
        if (stamp != nullptr && !__any_sync(0xffffffffu, sel)) continue;
        const int st = sel ? vox_state(g.vox[rbase + k]) : -1;
        const bool is_emp = st == 0, is_occ = st > 0 && st != P.defect_id;
        if (sel && !is_emp && !is_occ) other(k, st);
        const unsigned me = __ballot_sync(0xffffffffu, is_emp), mo = __ballot_sync(0xffffffffu, is_occ);
        if (is_emp) w.emp[n_emp + __popc(me & lt)] = (uint16_t)k;
        if (is_occ) w.occ[n_occ + __popc(mo & lt)] = (uint16_t)k;
        n_emp += __popc(me); n_occ += __popc(mo);
        if (zero_row && k < L) zero_row[k] = 0.0;
    }
    w.n_occ = n_occ; w.n_emp = n_emp;
    __syncwarp();
}

// Neighbour offsets as 32-bit linear index deltas, one table per CTA in shared memory (the
// dense kernels require the local extent to fit 31 bits, checked by the launchers).
struct NbOffsets { int lin[14]; };
__device__ __forceinline__ void nb_offsets_init(NbOffsets *t, int L)
{
    if (threadIdx.x < 14)
        t->lin[threadIdx.x] = ((int)c_nb_off[threadIdx.x][0] * L + c_nb_off[threadIdx.x][1]) * L + c_nb_off[threadIdx.x][2];
    __syncthreads();
}

// Neighbour states of site s, 4 bits per slot, split over two 32-bit words (slots 0-7, 8-13) so
// that the nibble tests below are single 32-bit operations.
struct Nst { unsigned lo, hi; };
__device__ __forceinline__ Nst neighbour_states32(const uint8_t *vox, int s, unsigned inb, const NbOffsets *t)
{
    // fully unrolled on purpose: 14 independent byte loads in flight at once (a rolled loop
    // serialises them behind the OR that consumes each result); the body is a handful of instructions
    unsigned b[14];
#pragma unroll
    for (int o = 0; o < 14; ++o) b[o] = (inb >> o & 1u) ? (unsigned)vox[s + t->lin[o]] & 15u : 0u;
    Nst n;
    n.lo = 0; n.hi = 0;
#pragma unroll
    for (int o = 0; o < 8; ++o) n.lo |= b[o] << (4 * o);
#pragma unroll
    for (int o = 8; o < 14; ++o) n.hi |= b[o] << (4 * (o - 8));
    return n;
}
__device__ __forceinline__ unsigned nib_nonzero32(unsigned x) { return (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u; }
__device__ __forceinline__ unsigned nib_equals32(unsigned x, int v) { return ~nib_nonzero32(x ^ (0x11111111u * (unsigned)v)) & 0x11111111u; }
// bit o of the result <=> bit 4*o of the nibble-LSB mask pair (lo: slots 0-7, hi: slots 8-13)
__device__ __forceinline__ bool slot_bit(unsigned lo, unsigned hi, int o)
{
    return ((o < 8 ? lo >> (4 * o) : hi >> (4 * (o - 8))) & 1u) != 0;
}

// One chunk of up to 32 occupied sites, one per lane (warp-collective: all 32 lanes call it;
// `active` false on padding lanes).  Returns the lane's diffusion-rate sum in slot order.
// Packed neighbour states of site s.  GATHER=false: one load from the cache (dense rebuild).
// GATHER=true: 14 byte loads from the lattice, and the cache entry is repaired (the refresh pass
// visits exactly the sites whose neighbourhood changed, so the cache needs no atomics in apply).
template <bool GATHER>
__device__ __forceinline__ Nst site_nst(const Lat &g, const NbOffsets *nbt, int s, unsigned inb, bool active,
                                        uint64_t *nst_out)
{
    Nst nst;
    if (GATHER) {
        nst = neighbour_states32(g.vox, s, inb, nbt);
        if (active) nst_out[s] = (uint64_t)nst.lo | ((uint64_t)nst.hi << 32);
    } else {
        const uint64_t nw = g.nst[s];
        nst.lo = (unsigned)nw; nst.hi = (unsigned)(nw >> 32);
    }
    return nst;
}

template <bool GATHER>
__device__ __forceinline__ double occ_chunk(const Lat &g, const cet_rate_params &P, const NbOffsets *nbt, int i, int j,
                                            int k, int s, bool active, uint64_t *nst_out = nullptr)
{
    const unsigned inb = inbounds_mask(i, j, k, g.n0, g.L);
    const Nst nst = site_nst<GATHER>(g, nbt, s, inb, active, nst_out);
    const int n_bonds = __popc(nib_nonzero32(nst.lo)) + __popc(nib_nonzero32(nst.hi));
    const bool has_events = active && n_bonds != __popc(inb);        // an in-bounds neighbour is empty
    double sum = 0.0;
    if (__any_sync(0xffffffffu, has_events)) {
        const uint8_t v = g.vox[s];
        OccPrep q;
        q.local_T = 1.0; q.boltz = 0.0;
        if (has_events) q = occ_prep(P, vox_state(v), vox_defects(v), g.T[s], n_bonds);
        unsigned emp_nb = 0;                                          // empty in-bounds neighbours, one bit per slot
        if (has_events) {
            const unsigned zl = ~nib_nonzero32(nst.lo) & 0x11111111u, zh = ~nib_nonzero32(nst.hi) & 0x00111111u;
#pragma unroll 1
            for (int o = 0; o < 14; ++o)
                if (slot_bit(zl, zh, o)) emp_nb |= 1u << o;
            emp_nb &= inb;
        }
#pragma unroll 2
        for (int o = 0; o < 14; ++o) {
            const bool on = emp_nb >> o & 1u;
            if (!__any_sync(0xffffffffu, on)) continue;
            if (on) sum += diff_pair_rate(P, q, g.T[s + nbt->lin[o]]);
        }
    }
    return sum;
}

// One chunk of up to 32 empty sites.  Returns the nucleation + attachment rate sum in slot order;
// the deposition event (global top plane only) is reported separately.
template <bool GATHER>
__device__ __forceinline__ double emp_chunk(const Lat &g, const cet_rate_params &P, const NbOffsets *nbt, int i, int j,
                                            int k, int s, bool active, bool *has_dep, double *dep,
                                            uint64_t *nst_out = nullptr)
{
    const int L = g.L;
    const unsigned inb = inbounds_mask(i, j, k, g.n0, L);
    const Nst nst = site_nst<GATHER>(g, nbt, s, inb, active, nst_out);
    const unsigned re_l = nib_equals32(nst.lo, P.states_re), re_h = nib_equals32(nst.hi, P.states_re) & 0x00111111u;
    const unsigned c_l = nib_equals32(nst.lo, P.states_c), c_h = nib_equals32(nst.hi, P.states_c) & 0x00111111u;
    unsigned att_l = nib_equals32(nst.lo, P.states_w) | re_l | c_l;           // nibble-LSB set: slot offers an attachment
    unsigned att_h = (nib_equals32(nst.hi, P.states_w) & 0x00111111u) | re_h | c_h;
    if (!active) { att_l = 0; att_h = 0; }
    const int km = k - 1 > 0 ? k - 1 : 0, kp = k + 1 < L - 1 ? k + 1 : L - 1;
    const double T_self = g.T[s];
    const EmpPrep q = emp_prep(P, T_self, g.T[s + (km - k)], g.T[s + (kp - k)],
                               __popc(re_l) + __popc(re_h) + __popc(c_l) + __popc(c_h), __popc(inb));
    double sum = q.nuc_rate;
    if (__any_sync(0xffffffffu, (att_l | att_h) != 0)) {
        const Vec4 sv = g.v[s];
        const double sx = sv.x, sy = sv.y, sz = sv.z;
#pragma unroll 2
        for (int o = 0; o < 14; ++o) {
            const bool on = slot_bit(att_l, att_h, o);
            if (!__any_sync(0xffffffffu, on)) continue;
            if (on) {
                const int t = s + nbt->lin[o];
                const int ia = slot_bit(re_l, re_h, o) ? 1 : (slot_bit(c_l, c_h, o) ? 2 : 0);
                const Vec4 nv = g.v[t];
                sum += att_pair_rate(P, q, ia, sx, sy, sz, nv.x, nv.y, nv.z);
            }
        }
    }
    *dep = 0.0;
    *has_dep = false;
    if (i == g.n0 - 1 && active) *has_dep = dep_rate(P, T_self, dep);
    return sum;
}

// Pass 2: occupied sites of a row.  fn(k, rate_sum, active) is called by all 32 lanes once per chunk.
template <class F>
__device__ __forceinline__ void row_occupied(const Lat &g, const cet_rate_params &P, const NbOffsets *nbt, int i,
                                             int j, int rbase, const RowLists &w, F &&fn)
{
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < w.n_occ; c0 += 32) {
        const bool active = c0 + lane < w.n_occ;
        const int k = w.occ[active ? c0 + lane : c0];
        fn(k, occ_chunk<false>(g, P, nbt, i, j, k, rbase + k, active), active);
    }
}

// Pass 3: empty sites of a row.  fn(k, rate_sum, has_dep, dep, active).
template <class F>
__device__ __forceinline__ void row_empty(const Lat &g, const cet_rate_params &P, const NbOffsets *nbt, int i, int j,
                                          int rbase, const RowLists &w, F &&fn)
{
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < w.n_emp; c0 += 32) {
        const bool active = c0 + lane < w.n_emp;
        const int k = w.emp[active ? c0 + lane : c0];
        bool has_dep;
        double dep;
        const double sum = emp_chunk<false>(g, P, nbt, i, j, k, rbase + k, active, &has_dep, &dep);
        fn(k, sum, has_dep, dep, active);
    }
}

}  // namespace cet
