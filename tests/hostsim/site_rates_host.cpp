// Test-only host build of csrc/site_rates.cuh: runs the product's per-site event enumeration
// on the CPU so that its ordering and arithmetic can be checked against the oracle without a GPU.
// Not part of the product (nothing in the package links or loads this).
#include <stdint.h>
#include <vector>
#include "../../cet-driven-simulation-for-3d-printing-am-kmc-approach_b200/csrc/site_rates.cuh"
#include "../../cet-driven-simulation-for-3d-printing-am-kmc-approach_b200/csrc/rate_tile.cuh"

using namespace cet;

extern "C" long long hostsim_events(const uint8_t *vox, const double *theta, const double *phi, const double *T,
                                    int L, const cet_rate_params *P, long long cap, uint8_t *type,
                                    long long *pos, double *rate, long long *target, int32_t *atom)
{
    const long long LL = (long long)L * L;
    std::vector<Vec4> v(LL * L);
    for (long long q = 0; q < LL * L; ++q) v[q] = unit_vec4(theta[q], phi[q]);
    Lat g;
    g.vox = vox; g.v = v.data(); g.T = T; g.L = L; g.n0 = L; g.i_off = 0;
    long long n = 0;
    auto put = [&](int ty, long long s, double r, long long t, int a) {
        if (n < cap) { type[n] = (uint8_t)ty; pos[n] = s; rate[n] = r; target[n] = t; atom[n] = a; }
        ++n;
    };
    for (int i = 0; i < L; ++i) {
        if (i == L - 1)
            for (int j = 0; j < L; ++j)
                for (int k = 0; k < L; ++k) {
                    const long long s = g.idx(i, j, k);
                    double r;
                    if (vox_state(vox[s]) == 0 && dep_rate(*P, T[s], &r)) put(CET_EV_DEP, s, r, -1, P->states_w);
                }
        for (int cls = 1; cls >= 0; --cls)     // occupied sites first, then empty ones
            for (int j = 0; j < L; ++j)
                for (int k = 0; k < L; ++k) {
                    const long long s = g.idx(i, j, k);
                    if ((vox_state(vox[s]) != 0) != (cls == 1)) continue;
                    site_events(g, *P, i, j, k, [&](int ty, int slot, double r, int a) {
                        put(ty, s, r, slot < 0 ? -1 : s + CET_NB_DI(slot) * LL + CET_NB_DJ(slot) * L + CET_NB_DK(slot), a);
                    });
                }
    }
    return n;
}

// The cached neighbour-class word of every site (rate_tile.cuh: nb_code_lut, nst_word) — what
// nst_build_kernel and the refresh write — so that its encoding can be checked without a GPU.
extern "C" void hostsim_nst_words(const uint8_t *vox, int L, const cet_rate_params *P, uint64_t *out)
{
    const uint64_t lut = nb_code_lut(*P);
    for (int i = 0; i < L; ++i)
        for (int j = 0; j < L; ++j)
            for (int k = 0; k < L; ++k) {
                const long long s = ((long long)i * L + j) * L + k;
                out[s] = nst_word(lut, vox, s, i, j, k, L, L);
            }
}

// The fused sweep kernel's per-site evaluation (tile_state.cuh: class codes, pairop, tile_site_prep,
// tile_pair_rate) run site by site on the host: `tile` receives its rate sums, `general` those of
// site_rate_sum (the per-event code).  On lattices whose empty sites carry no orientation the two
// must agree bit for bit.  Returns the number of empty sites that carry an orientation.
#include "../../cet-driven-simulation-for-3d-printing-am-kmc-approach_b200/csrc/tile_state.cuh"
extern "C" long long hostsim_tile_rates(const uint8_t *vox, const double *theta, const double *phi, const double *T,
                                        int L, const cet_rate_params *P, double *tile, double *general)
{
    const long long LL = (long long)L * L, N = LL * L;
    std::vector<Vec4> v(N);
    for (long long q = 0; q < N; ++q) v[q] = unit_vec4(theta[q], phi[q]);
    Lat g;
    g.vox = vox; g.v = v.data(); g.T = T; g.L = L; g.n0 = L; g.i_off = 0;
    std::vector<double> tab(RT_TABLE_DOUBLES);
    for (int t = 0; t < 256; ++t) tab[RT_KEFF + t] = nuc_K_eff(*P, t >> 4, t & 15);
    for (int t = 0; t < 48; ++t) tab[RT_ETOT + t] = occ_E_tot(*P, t >> 4, t & 15);
    for (int t = 0; t < 32; ++t) tab[RT_EXP2 + t] = h_exp2_tab[t];
    const uint64_t lut = tile_code_lut(*P);
    std::vector<uint8_t> cvox(N);
    std::vector<double> pairop(N);
    long long oriented = 0;
    for (long long s = 0; s < N; ++s) {
        const unsigned code = (unsigned)(lut >> (4 * (vox[s] & 15))) & 15u;
        cvox[s] = (uint8_t)((vox[s] & 0xF0) | code);
        pairop[s] = tile_pairop(*P, code, T[s], v[s].z);
        if (code == TC_EMPTY && !(v[s].x == 0.0 && v[s].y == 0.0 && v[s].z == 1.0)) ++oriented;
    }
    for (int i = 0; i < L; ++i)
        for (int j = 0; j < L; ++j)
            for (int k = 0; k < L; ++k) {
                const long long s = g.idx(i, j, k);
                general[s] = site_rate_sum(g, *P, i, j, k, nullptr);
                uint64_t w = 0;
                for (int o = 0; o < 14; ++o) {
                    const int ni = i + CET_NB_DI(o), nj = j + CET_NB_DJ(o), nk = k + CET_NB_DK(o);
                    if (ni < 0 || ni >= L || nj < 0 || nj >= L || nk < 0 || nk >= L) continue;      // outside: code 0
                    w |= (uint64_t)(cvox[g.idx(ni, nj, nk)] & 15u) << (4 * o);
                }
                const unsigned c = cvox[s];
                double T_self = 1.0, T_m = 1.0, T_p = 1.0;
                if ((c & 15u) == TC_EMPTY) {
                    T_self = pairop[s];
                    T_m = k > 0 ? ((cvox[s - 1] & 15u) == TC_EMPTY ? pairop[s - 1] : T[s - 1]) : T_self;
                    T_p = k < L - 1 ? ((cvox[s + 1] & 15u) == TC_EMPTY ? pairop[s + 1] : T[s + 1]) : T_self;
                } else {
                    T_self = T[s];
                }
                const TilePrep q = tile_site_prep(*P, tab.data(), w, c, T_self, T_m, T_p);
                double sum = q.sum0;
                for (int o = 0; o < 14; ++o)
                    if (q.pm >> (4 * o) & 1u) {
                        const long long t = s + CET_NB_DI(o) * LL + CET_NB_DJ(o) * L + CET_NB_DK(o);
                        sum += tile_pair_rate(*P, tab.data(), q.is_emp, q.A, q.B, pairop[t]);
                    }
                tile[s] = sum;
            }
    return oriented;
}
