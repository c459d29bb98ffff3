// site_rates.cuh — per-site event enumeration and Arrhenius rates.
//
// One function, `site_events`, produces the events of ONE lattice site in exactly the
// order the reference appends them (kmc_event_rates.py:75-109 for an occupied site,
// :112-158 for an empty one) and with the reference's floating-point evaluation order.
// Every consumer — the dense rate kernel, the neighbour-rate update after an event, the
// event-list export and the BKL search inside a site — goes through it, so a rate computed
// incrementally is bit-identical to the one a full rebuild produces.
//
// The file compiles for the device (nvcc) and for the host (g++, tests/hostsim only: the
// host build exists to check this arithmetic against the oracle without a GPU; it is not a
// product code path).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/cetkmc.h"

#if defined(__CUDACC__)
#define CET_HD __host__ __device__ __forceinline__
#else
#define CET_HD inline
#endif

namespace cet {

// Python / Numba max(a,b), min(a,b) on floats: keep `a` unless `b` compares strictly
// greater / smaller, so a NaN first argument propagates (kmc_event_rates.py:60,93,104,...).
CET_HD double pymax(double a, double b) { return (b > a) ? b : a; }
CET_HD double pymin(double a, double b) { return (b < a) ? b : a; }
CET_HD bool finite_f64(double x) { return (x - x) == 0.0; }

// kmc_event_rates.py:29-36 — offset order of get_bcc_neighbors
#define CET_NB_DI(o) ((o) < 2 ? 1 : (o) < 4 ? -1 : (o) == 8 ? 2 : (o) == 9 ? -2 : 0)
#define CET_NB_DJ(o) ((o) == 0 || (o) == 2 || (o) == 4 || (o) == 5 ? 1 : ((o) == 1 || (o) == 3 || (o) == 6 || (o) == 7) ? -1 : (o) == 10 ? 2 : (o) == 11 ? -2 : 0)
#define CET_NB_DK(o) ((o) == 4 || (o) == 6 ? 1 : ((o) == 5 || (o) == 7) ? -1 : (o) == 12 ? 2 : (o) == 13 ? -2 : 0)

// Packed voxel byte: low nibble = state (0 empty, 1 W, 2 Re, 3 C, 4 defect), high nibble =
// defects_mask value.
CET_HD int vox_state(uint8_t v) { return v & 0x0F; }
CET_HD int vox_defects(uint8_t v) { return v >> 4; }

// View of the lattice (or of one slab of it with ghost planes) in device memory.
struct Lat {
    const uint8_t *vox;
    const double *theta, *phi, *T;
    int L;       // global edge length (i, j and k all run over [0, L))
    int i_off;   // global i of local plane 0
    CET_HD int64_t idx(int i, int j, int k) const { return ((int64_t)(i - i_off) * L + j) * L + k; }
};

// kmc_event_rates.py:10-23 compute_misorientation, followed by cos() as used at :155.
// Returns cos(arccos(clamp(v1.v2))).
CET_HD double cos_misorientation(double t1, double p1, double t2, double p2)
{
    double s1 = sin(t1), c1 = cos(t1), s2 = sin(t2), c2 = cos(t2);
    double v1x = s1 * cos(p1), v1y = s1 * sin(p1);
    double v2x = s2 * cos(p2), v2y = s2 * sin(p2);
    double dot = v1x * v2x + v1y * v2y + c1 * c2;
    dot = pymax(pymin(dot, 1.0), -1.0);
    return cos(acos(dot));
}

// Deposition rate of an empty top-plane site (kmc_event_rates.py:60-64).  Returns false when
// the event does not exist (non-finite rate).  Zero rates are kept, as in the reference.
CET_HD bool dep_rate(const cet_rate_params &P, double T, double *rate)
{
    double local_T = pymax(T, 1.0);
    double thermal_factor = exp(-(P.T_melt - local_T) / (P.kT * local_T));
    double eff = P.nu_dep * thermal_factor;
    *rate = eff;
    return finite_f64(eff);
}

// kmc_event_rates.py:65-71 species of a deposited atom from one uniform draw
CET_HD int dep_species(const cet_rate_params &P, double r)
{
    if (r < P.impurity_c) return P.states_c;
    if (r < P.impurity_c + P.impurity_re) return P.states_re;
    return P.states_w;
}

// Enumerate the diff events (occupied site) or nuc+att events (empty site) of site (i,j,k)
// in reference order.  emit(type, slot, rate, atom): slot = index 0..13 into the offset table
// of the target / source neighbour, or -1 for nuc.
template <class F>
CET_HD void site_events(const Lat &g, const cet_rate_params &P, int i, int j, int k, F &&emit)
{
    const int L = g.L;
    const int64_t s = g.idx(i, j, k);
    const int64_t LL = (int64_t)L * L;
    const int self = vox_state(g.vox[s]);

    // in-bounds mask + neighbour states (both loops of the reference walk the same list)
    unsigned inb = 0;
    int nstate[14];
    int n_in = 0;
#pragma unroll
    for (int o = 0; o < 14; ++o) {
        const int ni = i + CET_NB_DI(o), nj = j + CET_NB_DJ(o), nk = k + CET_NB_DK(o);
        const bool ok = ni >= 0 && ni < L && nj >= 0 && nj < L && nk >= 0 && nk < L;
        nstate[o] = 0;
        if (ok) {
            inb |= 1u << o;
            ++n_in;
            nstate[o] = vox_state(g.vox[s + CET_NB_DI(o) * LL + CET_NB_DJ(o) * L + CET_NB_DK(o)]);
        }
    }

    if (self != 0) {
        // ---- occupied: diffusion, kmc_event_rates.py:79-109
        if (self == P.defect_id) return;                         // :80-81
        double E_b_atom, E_diff_atom;
        if (self == P.states_w) { E_b_atom = P.E_b[0]; E_diff_atom = P.E_diff[0]; }
        else if (self == P.states_re) { E_b_atom = P.E_b[1]; E_diff_atom = P.E_diff[1]; }
        else { E_b_atom = P.E_b[2]; E_diff_atom = P.E_diff[2]; }
        const double local_T = pymax(g.T[s], 1.0);
        const double defect_factor = 1.0 + (double)vox_defects(g.vox[s]);
        int n_bonds = 0;
#pragma unroll
        for (int o = 0; o < 14; ++o)
            if ((inb >> o & 1u) && nstate[o] != 0) ++n_bonds;
        if (n_bonds == n_in) return;                             // no empty neighbour: no event
        const double E_tot = pymax(E_diff_atom + 0.1 * (double)n_bonds * E_b_atom, 0.0);
        const double boltz = exp(-defect_factor * E_tot / (P.kT * local_T));
#pragma unroll
        for (int o = 0; o < 14; ++o) {
            if (!(inb >> o & 1u) || nstate[o] != 0) continue;
            const double neighbor_T =
                pymax(g.T[s + CET_NB_DI(o) * LL + CET_NB_DJ(o) * L + CET_NB_DK(o)], 1.0);
            const double dT = fabs(local_T - neighbor_T);
            const double denom = pymax(P.T_melt - neighbor_T, 1.0);
            const double grad_factor = 1.0 + 0.1 * dT / denom;
            const double rate = P.nu * grad_factor * boltz;
            if (rate > P.rate_threshold && finite_f64(rate)) emit((int)CET_EV_DIFF, o, rate, self);
        }
        return;
    }

    // ---- empty: nucleation + attachment, kmc_event_rates.py:116-158
    const double local_T = pymax(g.T[s], 1.0);
    const double dT = P.T_melt - local_T;
    if (dT > P.delta_T_c) {                                      // :120-132
        int n_imp = 0;
#pragma unroll
        for (int o = 0; o < 14; ++o)
            if ((inb >> o & 1u) && (nstate[o] == P.states_re || nstate[o] == P.states_c)) ++n_imp;
        const double f_imp = pymin(P.max_imp_fraction, (double)n_imp / (double)(n_in > 1 ? n_in : 1));
        double K_eff = P.k_nuc * (1.0 - P.beta_imp_nuc * f_imp);
        K_eff = pymax(0.1 * P.k_nuc, pymin(P.k_nuc, K_eff));
        const double barrier = K_eff / pymax((dT + 1e-6) * (dT + 1e-6), 1e-6);
        const double rate = P.i0 * exp(-barrier / (P.kT * local_T));
        if (rate > P.rate_threshold && finite_f64(rate)) emit((int)CET_EV_NUC, -1, rate, P.states_w);
    }
    bool have_self = false;
    double th_s = 0.0, ph_s = 0.0, gfac = 0.0, kTT = 0.0;
#pragma unroll
    for (int o = 0; o < 14; ++o) {                               // :135-158
        if (!(inb >> o & 1u)) continue;
        const int na = nstate[o];
        if (na == 0) continue;
        int ia;
        if (na == P.states_w) ia = 0;
        else if (na == P.states_re) ia = 1;
        else if (na == P.states_c) ia = 2;
        else continue;
        if (!have_self) {
            have_self = true;
            th_s = g.theta[s]; ph_s = g.phi[s];
            const int km = k - 1 > 0 ? k - 1 : 0, kp = k + 1 < L - 1 ? k + 1 : L - 1;
            const double grad_z = (g.T[s + (kp - k)] - g.T[s + (km - k)]) * 0.5;
            gfac = 1.0 + P.anisotropy * (pymax(0.0, grad_z) / pymax(P.T_melt - local_T, 1.0));
            kTT = P.kT * local_T;
        }
        const int64_t t = s + CET_NB_DI(o) * LL + CET_NB_DJ(o) * L + CET_NB_DK(o);
        const double cm = cos_misorientation(th_s, ph_s, g.theta[t], g.phi[t]);
        const double E_att = 0.5 * P.E_b[ia] * (1.0 - cm);
        const double rate = P.nu * exp(-E_att / kTT) * gfac;
        if (rate > P.rate_threshold && finite_f64(rate)) emit((int)CET_EV_ATT, o, rate, na);
    }
}

// Sum of a site's event rates in list order (what the dense kernel stores) and their count.
CET_HD double site_rate_sum(const Lat &g, const cet_rate_params &P, int i, int j, int k, int *count)
{
    double sum = 0.0;
    int n = 0;
    site_events(g, P, i, j, k, [&](int, int, double r, int) { sum += r; ++n; });
    if (count) *count = n;
    return sum;
}

}  // namespace cet
