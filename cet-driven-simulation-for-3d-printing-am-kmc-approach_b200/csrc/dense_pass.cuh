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
// buffer that is cleared here (defect sites keep rate 0).
__device__ __forceinline__ void row_classify(const Lat &g, const cet_rate_params &P, int64_t rbase, RowLists &w,
                                             double *zero_row)
{
    const int L = g.L, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int n_occ = 0, n_emp = 0;
    for (int k0 = 0; k0 < L; k0 += 32) {
        const int k = k0 + lane;
        const int st = k < L ? vox_state(g.vox[rbase + k]) : -1;
        const bool is_emp = st == 0, is_occ = st > 0 && st != P.defect_id;
        const unsigned me = __ballot_sync(0xffffffffu, is_emp), mo = __ballot_sync(0xffffffffu, is_occ);
        if (is_emp) w.emp[n_emp + __popc(me & lt)] = (uint16_t)k;
        if (is_occ) w.occ[n_occ + __popc(mo & lt)] = (uint16_t)k;
        n_emp += __popc(me); n_occ += __popc(mo);
        if (zero_row && k < L) zero_row[k] = 0.0;
    }
    w.n_occ = n_occ; w.n_emp = n_emp;
    __syncwarp();
}

// Pass 2: occupied sites.  fn(k, rate_sum, active) is called by all 32 lanes of the warp once
// per chunk (converged), `active` false on the padding lanes of the last chunk.
template <class F>
__device__ __forceinline__ void row_occupied(const Lat &g, const cet_rate_params &P, int i, int j, int64_t rbase,
                                             const RowLists &w, F &&fn)
{
    const int L = g.L, lane = threadIdx.x & 31;
    const unsigned inb_ij = inbounds_mask_ij(i, j, g.n0, L);
    for (int c0 = 0; c0 < w.n_occ; c0 += 32) {
        const bool active = c0 + lane < w.n_occ;
        const int k = w.occ[active ? c0 + lane : c0];
        const int64_t s = rbase + k;
        const unsigned inb = inbounds_mask_k(inb_ij, k, L);
        const uint64_t nst = neighbour_states(g, s, inb);
        const int n_bonds = popc64(nib_nonzero(nst));
        const bool has_events = active && n_bonds != popc32(inb);        // an in-bounds neighbour is empty
        double sum = 0.0;
        if (__any_sync(0xffffffffu, has_events)) {
            const uint8_t v = g.vox[s];
            OccPrep q;
            q.local_T = 1.0; q.boltz = 0.0;
            if (has_events) q = occ_prep(P, vox_state(v), vox_defects(v), g.T[s], n_bonds);
#pragma unroll 1
            for (int o = 0; o < 14; ++o) {
                const bool on = has_events && (inb >> o & 1u) && ((nst >> (4 * o)) & 15) == 0;
                if (!__any_sync(0xffffffffu, on)) continue;
                if (on) sum += diff_pair_rate(P, q, g.T[g.nb(s, o)]);
            }
        }
        fn(k, sum, active);
    }
}

// Pass 3: empty sites.  fn(k, rate_sum, has_dep, dep, active); the deposition event exists only
// on the global top plane (i == n0-1).
template <class F>
__device__ __forceinline__ void row_empty(const Lat &g, const cet_rate_params &P, int i, int j, int64_t rbase,
                                          const RowLists &w, F &&fn)
{
    const int L = g.L, lane = threadIdx.x & 31;
    const bool top = i == g.n0 - 1;
    const unsigned inb_ij = inbounds_mask_ij(i, j, g.n0, L);
    for (int c0 = 0; c0 < w.n_emp; c0 += 32) {
        const bool active = c0 + lane < w.n_emp;
        const int k = w.emp[active ? c0 + lane : c0];
        const int64_t s = rbase + k;
        const unsigned inb = inbounds_mask_k(inb_ij, k, L);
        const uint64_t nst = neighbour_states(g, s, inb);
        const uint64_t m_re = nib_equals(nst, P.states_re), m_c = nib_equals(nst, P.states_c);
        uint64_t att_nb = nib_equals(nst, P.states_w) | m_re | m_c;      // bit 4*o: slot o offers an attachment
        if (!active) att_nb = 0;
        const int km = k - 1 > 0 ? k - 1 : 0, kp = k + 1 < L - 1 ? k + 1 : L - 1;
        const double T_self = g.T[s];
        const EmpPrep q = emp_prep(P, T_self, g.T[s + (km - k)], g.T[s + (kp - k)], popc64(m_re) + popc64(m_c),
                                   popc32(inb));
        double sum = q.nuc_rate;
        if (__any_sync(0xffffffffu, att_nb != 0)) {
            const double sx = g.vx[s], sy = g.vy[s], sz = g.vz[s];
#pragma unroll 1
            for (int o = 0; o < 14; ++o) {
                const bool on = (att_nb >> (4 * o)) & 1u;
                if (!__any_sync(0xffffffffu, on)) continue;
                if (on) {
                    const int64_t t = g.nb(s, o);
                    const int ia = species_index(P, (int)(nst >> (4 * o)) & 15);
                    sum += att_pair_rate(P, q, ia, sx, sy, sz, g.vx[t], g.vy[t], g.vz[t]);
                }
            }
        }
        double dep = 0.0;
        bool has_dep = false;
        if (top && active) has_dep = dep_rate(P, T_self, &dep);
        fn(k, sum, has_dep, dep, active);
    }
}

}  // namespace cet
