// site_rates.cuh — per-site event enumeration and Arrhenius rates (kmc_event_rates.py:43-160).
//
// Building blocks, shared by EVERY consumer (dense rate kernel, sweep kernel, incremental
// refresh after an event, event-list export, event pick inside a site) so that a rate computed
// by any of them has the same bits:
//     occ_prep / diff_pair_rate    occupied site  -> diffusion events    (:79-109)
//     emp_prep / att_pair_rate     empty site     -> nucleation + attachment events (:116-158)
//     dep_rate                     empty top-plane site -> deposition event (:55-72)
//     site_events                  the events of one site in the reference's list order
//
// Differences from a literal transcription, all within the stated 1e-12 relative tolerance:
//   * orientations enter through resident unit vectors v = (sin t cos p, sin t sin p, cos t)
//     kept in HBM beside theta/phi (kmc_event_rates.py:11-20 recomputes them per pair: 8
//     trigonometric calls per attachment event); cos(arccos(dot)) (:23,:155) is taken as dot;
//   * x / y inside a rate is x * rcp(y) with a Newton-refined reciprocal (<= 1 ulp);
//   * the Arrhenius factors use fast_exp (table of 2^(j/32) + degree-6 polynomial, <= 1 ulp; 11
//     fp64 instructions instead of libdevice's ~30), and the per-site constant factors are
//     multiplied first: rate = grad * (nu * boltz), rate = exp(..) * (nu * gfac).
//
// The file also compiles for the host (g++, tests/hostsim only — a checker for this arithmetic
// that needs no GPU; not a product code path).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/cetkmc.h"

#if defined(__CUDACC__)
#define CET_HD __host__ __device__ __forceinline__
#define CET_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CET_HD inline
#define CET_HD_NOINLINE inline
#endif

namespace cet {

// Python / Numba max(a,b), min(a,b) on floats: keep `a` unless `b` compares strictly
// greater / smaller, so a NaN first argument propagates (kmc_event_rates.py:60,93,104,...).
CET_HD double pymax(double a, double b) { return (b > a) ? b : a; }
CET_HD double pymin(double a, double b) { return (b < a) ? b : a; }
CET_HD bool finite_f64(double x) { return (x - x) == 0.0; }

// 1/x for normal positive x: hardware seed + two Newton steps on the device.
CET_HD double rcp(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}

// exp(x): x = (32 m + j) ln2/32 + r, |r| <= ln2/64;  exp(x) = 2^m * 2^(j/32) * (1 + p(r)).
// Outside |x| < 700 (overflow, gradual underflow, NaN) the library exp is used.
#define CET_EXP2_TAB_INIT {0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, \
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, \
    0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, \
    0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, \
    0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, \
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, \
    0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0}
static const double h_exp2_tab[32] = CET_EXP2_TAB_INIT;
#if defined(__CUDACC__)
static __device__ const double d_exp2_tab[32] = CET_EXP2_TAB_INIT;
#endif
static CET_HD_NOINLINE double slow_exp(double x) { return exp(x); }   // rare path, kept out of line
#if defined(__CUDACC__)
// fast_exp constants in the constant bank: fp64 instructions take them as operands directly (an
// immediate costs two extra move instructions per constant)
static __constant__ double c_exp_k[9] = {0x1.71547652b82fep+5, 6755399441055744.0, -0x1.62e42fefa39efp-6, -0x1.abc9e3b39803fp-61,
                                         0x1.6c16c16c16c17p-10, 0x1.1111111111111p-7, 0x1.5555555555555p-5,
                                         0x1.5555555555555p-3, 0.5};
#endif
// the in-range part of fast_exp_t (|x| < 700)
CET_HD double fast_exp_core(double x, const double *tab)   // tab: 2^(j/32), j = 0..31 (any address space)
{
#if defined(__CUDA_ARCH__)
    const double tt = __fma_rn(x, c_exp_k[0], c_exp_k[1]);                 // x * 32/ln2, rint by magic add (1.5 * 2^52)
    const int n = __double2loint(tt);
    const double nf = tt - c_exp_k[1];
    double r = __fma_rn(nf, c_exp_k[2], x);                                 // ln2/32 in two parts
    r = __fma_rn(nf, c_exp_k[3], r);
    double q = __fma_rn(c_exp_k[4], r, c_exp_k[5]);                         // 1/720, 1/120
    q = __fma_rn(q, r, c_exp_k[6]);                                         // 1/24
    q = __fma_rn(q, r, c_exp_k[7]);                                         // 1/6
    q = __fma_rn(q, r, c_exp_k[8]);
    const double p = __fma_rn(q, r * r, r);                                 // exp(r) - 1
    const double t = tab[n & 31];
    const double y = __fma_rn(t, p, t);
    return __hiloint2double(__double2hiint(y) + ((n >> 5) << 20), __double2loint(y));
#else
    const double big = 6755399441055744.0;
    const double tt = fma(x, 0x1.71547652b82fep+5, big);
    int64_t bits;
    memcpy(&bits, &tt, 8);
    const int n = (int)(uint32_t)bits;
    const double nf = tt - big;
    double r = fma(nf, -0x1.62e42fefa39efp-6, x);
    r = fma(nf, -0x1.abc9e3b39803fp-61, r);
    double q = fma(0x1.6c16c16c16c17p-10, r, 0x1.1111111111111p-7);
    q = fma(q, r, 0x1.5555555555555p-5);
    q = fma(q, r, 0x1.5555555555555p-3);
    q = fma(q, r, 0.5);
    const double p = fma(q, r * r, r);
    const double t = tab[n & 31];
    const double y = fma(t, p, t);
    int64_t yb;
    memcpy(&yb, &y, 8);
    yb += (int64_t)(n >> 5) << 52;
    double out;
    memcpy(&out, &yb, 8);
    return out;
#endif
}
CET_HD double fast_exp_t(double x, const double *tab)
{
    if (!(fabs(x) < 700.0)) return slow_exp(x);
    return fast_exp_core(x, tab);
}

#if defined(__CUDA_ARCH__)
#define CET_EXP2_TAB d_exp2_tab
#else
#define CET_EXP2_TAB h_exp2_tab
#endif
CET_HD double fast_exp(double x) { return fast_exp_t(x, CET_EXP2_TAB); }

// 1/x with one Newton step (relative error < 2^-45): for quotients that enter a rate as 1 + small * q
CET_HD double rcp1(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return __fma_rn(r, __fma_rn(-x, r, 1.0), r);
#else
    return 1.0 / x;
#endif
}

// kmc_event_rates.py:29-36 — offset order of get_bcc_neighbors
#define CET_NB_DI(o) ((o) < 2 ? 1 : (o) < 4 ? -1 : (o) == 8 ? 2 : (o) == 9 ? -2 : 0)
#define CET_NB_DJ(o) ((o) == 0 || (o) == 2 || (o) == 4 || (o) == 5 ? 1 : ((o) == 1 || (o) == 3 || (o) == 6 || (o) == 7) ? -1 : (o) == 10 ? 2 : (o) == 11 ? -2 : 0)
#define CET_NB_DK(o) ((o) == 4 || (o) == 6 ? 1 : ((o) == 5 || (o) == 7) ? -1 : (o) == 12 ? 2 : (o) == 13 ? -2 : 0)

// The same table as data, for loops that are deliberately NOT unrolled (the fully unrolled form
// of the 14-slot loops with an inlined fp64 exp per slot is ~140 KB of SASS and stalls on
// instruction fetch; see profiles/).
#define CET_NB_TABLE_INIT {{1, 1, 0}, {1, -1, 0}, {-1, 1, 0}, {-1, -1, 0}, {0, 1, 1}, {0, 1, -1}, {0, -1, 1}, \
                           {0, -1, -1}, {2, 0, 0}, {-2, 0, 0}, {0, 2, 0}, {0, -2, 0}, {0, 0, 2}, {0, 0, -2}}
static const int8_t h_nb_off[14][3] = CET_NB_TABLE_INIT;
#if defined(__CUDACC__)
static __constant__ int8_t c_nb_off[14][3] = CET_NB_TABLE_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define CET_NB_TAB c_nb_off
#else
#define CET_NB_TAB h_nb_off
#endif

// slot of the opposite offset: offset[CET_NB_OPP(o)] == -offset[o]
#define CET_NB_OPP(o) ((o) < 4 ? 3 - (o) : (o) < 8 ? 11 - (o) : ((o) ^ 1))

// Packed voxel byte: low nibble = state (0 empty, 1 W, 2 Re, 3 C, 4 defect), high nibble =
// defects_mask value.
CET_HD int vox_state(uint8_t v) { return v & 0x0F; }
CET_HD int vox_defects(uint8_t v) { return v >> 4; }

// Orientation unit vector of a site, one 32-byte record (= one DRAM sector: a gathered neighbour
// costs one sector instead of three with separate x/y/z arrays).  w is unused padding.
struct alignas(32) Vec4 { double x, y, z, w; };

// View of the lattice (or of one slab of it with ghost planes) in device memory.
struct Lat {
    const uint8_t *vox;
    const uint64_t *nst;   // packed states of the 14 neighbours, 4 bits per slot (maintained cache; dense kernels only)
    const Vec4 *v;
    const double *T;
    int L;       // edge length of axes 1 and 2
    int n0;      // global number of planes along axis 0 (== L for the reference's cubic lattices)
    int i_off;   // global i of local plane 0
    CET_HD int64_t idx(int i, int j, int k) const { return ((int64_t)(i - i_off) * L + j) * L + k; }
    CET_HD int64_t nb(int64_t s, int o) const
    {
        return s + ((int64_t)CET_NB_TAB[o][0] * L + CET_NB_TAB[o][1]) * L + CET_NB_TAB[o][2];
    }
};

// In-bounds mask of the 14 neighbours (bit o set <=> get_bcc_neighbors keeps offset o).
CET_HD unsigned inbounds_mask_ij(int i, int j, int n0, int L)      // the part that depends on the row only
{
    unsigned m = 0x3FFFu;
    if (i + 1 >= n0) m &= ~0x0003u;      // slots 0,1: di = +1
    if (i - 1 < 0) m &= ~0x000Cu;        // slots 2,3: di = -1
    if (i + 2 >= n0) m &= ~0x0100u;      // slot 8
    if (i - 2 < 0) m &= ~0x0200u;        // slot 9
    if (j + 1 >= L) m &= ~0x0035u;       // slots 0,2,4,5: dj = +1
    if (j - 1 < 0) m &= ~0x00CAu;        // slots 1,3,6,7: dj = -1
    if (j + 2 >= L) m &= ~0x0400u;       // slot 10
    if (j - 2 < 0) m &= ~0x0800u;        // slot 11
    return m;
}
CET_HD unsigned inbounds_mask_k(unsigned m, int k, int L)          // ... and the per-site part
{
    if (k + 1 >= L) m &= ~0x0050u;       // slots 4,6: dk = +1
    if (k - 1 < 0) m &= ~0x00A0u;        // slots 5,7: dk = -1
    if (k + 2 >= L) m &= ~0x1000u;       // slot 12
    if (k - 2 < 0) m &= ~0x2000u;        // slot 13
    return m;
}
CET_HD unsigned inbounds_mask(int i, int j, int k, int n0, int L)
{
    return inbounds_mask_k(inbounds_mask_ij(i, j, n0, L), k, L);
}

// Orientation unit vector (kmc_event_rates.py:11-15).
CET_HD void unit_vector(double theta, double phi, double *x, double *y, double *z)
{
    const double st = sin(theta);
    *x = st * cos(phi); *y = st * sin(phi); *z = cos(theta);
}
CET_HD Vec4 unit_vec4(double theta, double phi)
{
    Vec4 v;
    unit_vector(theta, phi, &v.x, &v.y, &v.z);
    v.w = 0.0;
    return v;
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

// `rate > RATE_THRESHOLD and isfinite(rate)` (kmc_event_rates.py:108,131,157): the rate, or 0 when filtered
CET_HD double keep_rate(const cet_rate_params &P, double rate)
{
    return (rate > P.rate_threshold && rate < INFINITY) ? rate : 0.0;
}

// ---- occupied site: diffusion (kmc_event_rates.py:79-109) ------------------------------------
// Split into the pieces the dense kernel evaluates separately (per-site exponent argument, one
// exp shared by both site classes, per-pair rates); occ_prep / emp_prep compose the same pieces,
// so every consumer produces the same bits.
struct OccPrep { double local_T, nb; };            // nb = nu * exp(-defect_factor * E_tot / (kT * local_T))

CET_HD int species3(const cet_rate_params &P, int st) { return st == P.states_w ? 0 : st == P.states_re ? 1 : 2; }   // :83-91
CET_HD double occ_E_tot(const cet_rate_params &P, int sp, int n_bonds)                                              // :105
{
    return pymax(P.E_diff[sp] + 0.1 * (double)n_bonds * P.E_b[sp], 0.0);
}
CET_HD double occ_exp_arg(int defects, double E_tot, double inv_kTT) { return -(1.0 + (double)defects) * E_tot * inv_kTT; }

CET_HD OccPrep occ_prep(const cet_rate_params &P, int self_state, int defects, double T_self, int n_bonds)
{
    OccPrep q;
    q.local_T = pymax(T_self, 1.0);
    const double inv_kTT = rcp(P.kT * q.local_T);
    q.nb = P.nu * fast_exp(occ_exp_arg(defects, occ_E_tot(P, species3(P, self_state), n_bonds), inv_kTT));
    return q;
}

// rate of the diffusion event into an empty neighbour whose raw temperature is Tn_raw;
// returns 0 when the event is filtered (:108)
CET_HD double diff_pair_rate(const cet_rate_params &P, double local_T, double nb, double Tn_raw)
{
    const double neighbor_T = pymax(Tn_raw, 1.0);
    const double dT = fabs(local_T - neighbor_T);
    const double denom = pymax(P.T_melt - neighbor_T, 1.0);
    const double grad_factor = fma(0.1 * dT, rcp1(denom), 1.0);
    return keep_rate(P, grad_factor * nb);
}

// ---- empty site: nucleation + attachment (kmc_event_rates.py:116-158) --------------------------
struct EmpPrep {
    double nuc_rate;     // 0 when there is no nucleation event
    double inv_kTT;      // 1 / (kT * local_T)
    double ng;           // nu * (1 + ANISOTROPY * max(0, grad_z) / max(T_melt - local_T, 1))
};

CET_HD double nuc_K_eff(const cet_rate_params &P, int n_imp, int n_in)                      // :126-128
{
    const double f_imp = pymin(P.max_imp_fraction, (double)n_imp / (double)(n_in > 1 ? n_in : 1));
    const double K_eff = P.k_nuc * (1.0 - P.beta_imp_nuc * f_imp);
    return pymax(0.1 * P.k_nuc, pymin(P.k_nuc, K_eff));
}
CET_HD bool nuc_exists(const cet_rate_params &P, double local_T) { return (P.T_melt - local_T) > P.delta_T_c; }   // :120
CET_HD double nuc_exp_arg(const cet_rate_params &P, double local_T, double K_eff, double inv_kTT)                  // :129-130
{
    const double d = (P.T_melt - local_T) + 1e-6;
    const double barrier = K_eff * rcp(pymax(d * d, 1e-6));
    return -barrier * inv_kTT;
}
CET_HD double nuc_from_exp(const cet_rate_params &P, double e)                                                     // :130-131
{
    return keep_rate(P, P.i0 * e);
}
CET_HD double emp_ng(const cet_rate_params &P, double local_T, double T_km, double T_kp)                           // :151-154
{
    const double grad_z = (T_kp - T_km) * 0.5;
    const double gfac = 1.0 + P.anisotropy * (pymax(0.0, grad_z) * rcp1(pymax(P.T_melt - local_T, 1.0)));
    return P.nu * gfac;
}

CET_HD EmpPrep emp_prep(const cet_rate_params &P, double T_self, double T_km, double T_kp, int n_imp, int n_in)
{
    EmpPrep q;
    const double local_T = pymax(T_self, 1.0);
    q.inv_kTT = rcp(P.kT * local_T);
    q.nuc_rate = 0.0;
    if (nuc_exists(P, local_T))
        q.nuc_rate = nuc_from_exp(P, fast_exp(nuc_exp_arg(P, local_T, nuc_K_eff(P, n_imp, n_in), q.inv_kTT)));
    q.ng = emp_ng(P, local_T, T_km, T_kp);
    return q;
}

// rate of the attachment event copying an occupied neighbour of species energy hE = 0.5 * E_b[ia]
// (ia = 0 W, 1 Re, 2 C); returns 0 when filtered (:157)
// E_att = 0.5 * E_b * (1 - cos(mis)) from the clamped dot product of the two unit vectors (:23,:155)
CET_HD double att_E(double hE, double dot) { return hE * (1.0 - pymax(pymin(dot, 1.0), -1.0)); }
CET_HD double att_pair_rate_E(const cet_rate_params &P, double E_att, double inv_kTT, double ng,
                              const double *exp_tab = CET_EXP2_TAB)
{
    return keep_rate(P, fast_exp_t(-E_att * inv_kTT, exp_tab) * ng);
}
CET_HD double att_pair_rate(const cet_rate_params &P, double hE, double inv_kTT, double ng, double sx, double sy,
                            double sz, double nx, double ny, double nz, const double *exp_tab = CET_EXP2_TAB)
{
    // an empty site that carries no orientation has s = (0, 0, 1) and the dot product is nz bit for bit:
    // the tile kernel (sweep_tile.cu) keeps att_E(hE, nz) resident per occupied site for that case
    return att_pair_rate_E(P, att_E(hE, fma(sz, nz, fma(sy, ny, sx * nx))), inv_kTT, ng, exp_tab);
}

CET_HD int species_index(const cet_rate_params &P, int st)
{
    return st == P.states_w ? 0 : st == P.states_re ? 1 : st == P.states_c ? 2 : -1;
}

// Neighbour states of site s packed 4 bits per offset (0 for out-of-bounds offsets).
CET_HD uint64_t neighbour_states(const Lat &g, int64_t s, unsigned inb)
{
    uint64_t nst = 0;
#pragma unroll 1
    for (int o = 0; o < 14; ++o)
        if (inb >> o & 1u) nst |= (uint64_t)vox_state(g.vox[g.nb(s, o)]) << (4 * o);
    return nst;
}

// SWAR helpers on the packed neighbour states: a mask with bit 4*o set for every slot whose
// nibble is non-zero / equals `value` (1..15).
#define CET_NIB_LSB 0x0011111111111111ull
CET_HD uint64_t nib_nonzero(uint64_t x) { return (x | (x >> 1) | (x >> 2) | (x >> 3)) & CET_NIB_LSB; }
CET_HD uint64_t nib_equals(uint64_t x, int value) { return ~nib_nonzero(x ^ (CET_NIB_LSB * (uint64_t)value)) & CET_NIB_LSB; }
CET_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
CET_HD int popc32(unsigned x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// Enumerate the diff events (occupied site) or nuc+att events (empty site) of site (i,j,k)
// in reference order.  emit(type, slot, rate, atom): slot = index 0..13 into the offset table
// of the target / source neighbour, or -1 for nuc.
template <class F>
CET_HD void site_events(const Lat &g, const cet_rate_params &P, int i, int j, int k, F &&emit)
{
    const int L = g.L;
    const int64_t s = g.idx(i, j, k);
    const uint8_t v = g.vox[s];
    const int self = vox_state(v);
    const unsigned inb = inbounds_mask(i, j, k, g.n0, L);
    const uint64_t nst = neighbour_states(g, s, inb);
    const int n_in = popc32(inb), n_bonds = popc64(nib_nonzero(nst));

    if (self != 0) {
        if (self == P.defect_id) return;                         // :80-81
        if (n_bonds == n_in) return;                             // no empty neighbour: no event
        const OccPrep q = occ_prep(P, self, vox_defects(v), g.T[s], n_bonds);
#pragma unroll 1
        for (int o = 0; o < 14; ++o) {
            if (!(inb >> o & 1u) || ((nst >> (4 * o)) & 15) != 0) continue;
            const double rate = diff_pair_rate(P, q.local_T, q.nb, g.T[g.nb(s, o)]);
            if (rate != 0.0) emit((int)CET_EV_DIFF, o, rate, self);
        }
        return;
    }

    const int n_imp = popc64(nib_equals(nst, P.states_re)) + popc64(nib_equals(nst, P.states_c));
    const int km = k - 1 > 0 ? k - 1 : 0, kp = k + 1 < L - 1 ? k + 1 : L - 1;
    const EmpPrep q = emp_prep(P, g.T[s], g.T[s + (km - k)], g.T[s + (kp - k)], n_imp, n_in);
    if (q.nuc_rate != 0.0) emit((int)CET_EV_NUC, -1, q.nuc_rate, P.states_w);
    if (n_bonds == 0) return;
    const Vec4 sv = g.v[s];
    const double sx = sv.x, sy = sv.y, sz = sv.z;
#pragma unroll 1
    for (int o = 0; o < 14; ++o) {                               // :135-158
        const int na = (int)(nst >> (4 * o)) & 15;
        const int ia = species_index(P, na);
        if (na == 0 || ia < 0) continue;
        const int64_t t = g.nb(s, o);
        const Vec4 nv = g.v[t];
        const double rate = att_pair_rate(P, 0.5 * P.E_b[ia], q.inv_kTT, q.ng, sx, sy, sz, nv.x, nv.y, nv.z);
        if (rate != 0.0) emit((int)CET_EV_ATT, o, rate, na);
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
