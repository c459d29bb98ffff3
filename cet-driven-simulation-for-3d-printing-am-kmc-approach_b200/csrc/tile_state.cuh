// tile_state.cuh — the two resident arrays the fused sweep kernel (sweep_tile.cu) stages with TMA,
// and the per-site half of its rate evaluation (shared with the host checker in tests/hostsim).
//
//   cvox    u8   class code of the site in the low nibble, defects_mask value in the high nibble
//                (kmc_event_rates.py:94).  Code bits: 0 occupied, 1 Re, 2 C, 3 "inside the lattice and
//                empty or attachable".  Values: 0 outside the lattice (what a TMA box reads beyond the
//                tensor, so tiles at the lattice faces need no special case), 8 empty, 9 W, 11 Re, 13 C,
//                3 defect (state == defect_id: no events, kmc_event_rates.py:80), 1 any other state.
//   pairop  f64  the one operand a PAIR event needs from the neighbour site:
//                empty neighbour    -> its temperature T            (diffusion target, :102-107)
//                W / Re / C neighbour -> E_att = 0.5 E_b (1 - cos mis) against an unoriented empty site
//                                      (attachment source, :147-155), i.e. att_E(hE, v.z)
//                so a 3-D tile of cvox + pairop with a halo of 2 holds everything the 14-neighbour rate
//                model reads, and an empty site finds its own T in pairop as well.
// pairop's attachment half assumes the reference's invariant that empty sites carry theta = phi = 0
// (lattice_init.py:24-25, kmc_simulation.py:289-290,300-301): then v_self = (0,0,1) and the dot
// product of kmc_event_rates.py:16-21 is the neighbour's z component bit for bit.  tile_state_build
// checks the invariant; lattices that violate it run the general gather kernels (rates.cu).
#pragma once
#include "rate_tile.cuh"
#include "site_rates.cuh"

namespace cet {

constexpr int TL_I = 4, TL_J = 8, TL_K = 32;             // sites of one tile (planes x rows x k)
constexpr int TL_SITES = TL_I * TL_J * TL_K;
constexpr int TL_HI = TL_I + 4, TL_HJ = TL_J + 4;        // staged planes / rows (halo 2)
constexpr int TL_VK = 64, TL_VK0 = 16;                   // staged cvox bytes per row and offset of the tile's first k: a u8 box row of 48 B
                                                         // faults (illegal instruction) on B200 although it is a multiple of 16 B; 64 B works
                                                         // (scripts/tma_probe.cu)
constexpr int TL_PK = TL_K + 4, TL_PK0 = 2;              // staged pairop doubles per row
constexpr int TL_VBYTES = TL_HI * TL_HJ * TL_VK;         // 6 144
constexpr int TL_PBYTES = TL_HI * TL_HJ * TL_PK * 8;     // 27 648
constexpr int TL_THREADS = 256, TL_WARPS = TL_THREADS / 32;

enum { TC_OUTSIDE = 0, TC_OTHER = 1, TC_DEFECT = 3, TC_EMPTY = 8, TC_W = 9, TC_RE = 11, TC_C = 13 };

CET_HD unsigned tile_code(const cet_rate_params &P, int st)
{
    if (st == 0) return TC_EMPTY;
    if (st == P.states_w) return TC_W;
    if (st == P.states_re) return TC_RE;
    if (st == P.states_c) return TC_C;
    if (st == P.defect_id) return TC_DEFECT;
    return TC_OTHER;
}
CET_HD uint64_t tile_code_lut(const cet_rate_params &P)
{
    uint64_t lut = 0;
    for (int st = 0; st < 16; ++st) lut |= (uint64_t)tile_code(P, st) << (4 * st);
    return lut;
}
// pairop of a site (code: its class code; T: its temperature; z: z component of its unit vector)
CET_HD double tile_pairop(const cet_rate_params &P, unsigned code, double T, double z)
{
    if (code == TC_EMPTY) return T;
    if ((code & 9u) == 9u) return att_E(0.5 * P.E_b[(code >> 1) & 3u], z);
    return 0.0;
}

// Per-site half of the evaluation (phase A of rate_tile.cuh with the class codes of this file).
//   w       class codes of the 14 neighbours, 4 bits per slot (0 = outside the lattice)
//   c       the site's own cvox byte
//   T_self  the site's temperature; T_km / T_kp the temperatures at k -+ 1 (the site's own at the row ends)
//   tab     K_eff / E_tot / 2^(j/32) tables of rate_tile.cuh
// Returns the pair mask (bit 4*o per pair slot) and what the pair phase needs.
struct TilePrep {
    uint64_t pm;        // empty site: attachable neighbours; occupied site: empty neighbours
    double A, B;        // empty: 1/(kT T'), nu*gfac; occupied: T', nu*boltz
    double sum0;        // nucleation rate (0 when there is none)
    bool is_emp;
};
// The halves of tile_site_prep by site class: the dense kernel (rates_dense.cu) sorts the sites of a tile by
// class first and calls them separately; tile_site_prep composes the same functions, so both give the same bits.
CET_HD int tile_n_in(uint64_t w) { return popc64((w | (w >> 3)) & CET_NIB_LSB); }     // a class code is non-zero iff bit 0 or bit 3 is set
// nucleation rate of an empty site (kmc_event_rates.py:120-131); 0 when the event does not exist
CET_HD double tile_nuc_rate(const cet_rate_params &P, const double *tab, uint64_t w, uint64_t m_att, double local_T, double inv_kTT)
{
    if (!nuc_exists(P, local_T)) return 0.0;
    const int n_imp = popc64((w >> 1) & m_att) + popc64((w >> 2) & m_att);
    return nuc_from_exp(P, fast_exp_t(nuc_exp_arg(P, local_T, tab[RT_KEFF + n_imp * 16 + tile_n_in(w)], inv_kTT), tab + RT_EXP2));
}
CET_HD TilePrep tile_prep_emp(const cet_rate_params &P, const double *tab, uint64_t w, double T_self, double T_km, double T_kp)
{
    TilePrep r;
    r.is_emp = true; r.A = 0.0; r.B = 0.0;
    const uint64_t m_att = w & (w >> 3) & CET_NIB_LSB;             // occupied neighbours of an attachable species
    const double local_T = pymax(T_self, 1.0);
    const double inv_kTT = rcp(P.kT * local_T);
    r.pm = m_att;
    r.sum0 = tile_nuc_rate(P, tab, w, m_att, local_T, inv_kTT);
    if (r.pm) { r.A = inv_kTT; r.B = emp_ng(P, local_T, T_km, T_kp); }
    return r;
}
// code: TC_W / TC_RE / TC_C; df: the site's defects_mask value
CET_HD TilePrep tile_prep_occ(const cet_rate_params &P, const double *tab, uint64_t w, unsigned code, int df, double T_self)
{
    TilePrep r;
    r.is_emp = false; r.A = 0.0; r.B = 0.0; r.sum0 = 0.0;
    r.pm = ~w & (w >> 3) & CET_NIB_LSB;                            // empty neighbours inside the lattice
    if (r.pm) {
        const double local_T = pymax(T_self, 1.0);
        const double inv_kTT = rcp(P.kT * local_T);
        const int sp = code == TC_W ? 0 : code == TC_RE ? 1 : 2;                           // :83-91
        const double e = fast_exp_t(occ_exp_arg(df, tab[RT_ETOT + sp * 16 + popc64(w & CET_NIB_LSB)], inv_kTT), tab + RT_EXP2);
        r.A = local_T; r.B = P.nu * e;
    }
    return r;
}
CET_HD TilePrep tile_site_prep(const cet_rate_params &P, const double *tab, uint64_t w, unsigned c, double T_self,
                               double T_km, double T_kp)
{
    const unsigned code = c & 15u;
    if (code == TC_EMPTY) return tile_prep_emp(P, tab, w, T_self, T_km, T_kp);
    if ((code & 1u) && code != TC_DEFECT) return tile_prep_occ(P, tab, w, code, (int)(c >> 4), T_self);
    TilePrep r;
    r.pm = 0; r.A = 0.0; r.B = 0.0; r.sum0 = 0.0; r.is_emp = false;
    return r;
}
// one pair: op = pairop of the neighbour
CET_HD double tile_pair_rate(const cet_rate_params &P, const double *tab, bool is_emp, double A, double B, double op)
{
    return is_emp ? att_pair_rate_E(P, op, A, B, tab + RT_EXP2) : diff_pair_rate(P, A, B, op);
}

}  // namespace cet
