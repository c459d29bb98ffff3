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
