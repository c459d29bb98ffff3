/*
 * oracle.c — CPU restatement of the reference KMC / thermal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * call this file; it is the checker used by tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: the reference ships no tests and no golden vectors ("parity
 * unpinned" by the reference's own suite), so this restatement is pinned against
 * the reference ITSELF, imported from /root/reference in the build container
 * (tests/test_oracle_vs_reference.py, bit-exact) and against fixtures generated
 * from it (oracle/gen_golden.py -> tests/golden/).
 *
 * Every function cites the reference file:line it follows.  Arithmetic is written
 * in the reference's evaluation order; build with -ffp-contract=off (see Makefile)
 * so no multiply-add is fused (Numba/LLVM does not contract either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* Python / Numba `max(a, b)` and `min(a, b)` for floats: keep `a` unless `b`
 * compares strictly greater / smaller (so a NaN first argument propagates). */
static inline double pymax(double a, double b) { return (b > a) ? b : a; }
static inline double pymin(double a, double b) { return (b < a) ? b : a; }

/* kmc_event_rates.py:29-36 — fixed offset order */
static const int OFF[14][3] = {
    {1, 1, 0}, {1, -1, 0}, {-1, 1, 0}, {-1, -1, 0},
    {0, 1, 1}, {0, 1, -1}, {0, -1, 1}, {0, -1, -1},
    {2, 0, 0}, {-2, 0, 0}, {0, 2, 0}, {0, -2, 0}, {0, 0, 2}, {0, 0, -2}};

/* kmc_event_rates.py:26-40 get_bcc_neighbors: in-bounds subset, order kept */
int oracle_bcc_neighbors(int64_t i, int64_t j, int64_t k, int64_t L, int64_t *out /* [14*3] */)
{
    int n = 0;
    for (int o = 0; o < 14; ++o) {
        int64_t ni = i + OFF[o][0], nj = j + OFF[o][1], nk = k + OFF[o][2];
        if (ni >= 0 && ni < L && nj >= 0 && nj < L && nk >= 0 && nk < L) {
            out[3 * n + 0] = ni; out[3 * n + 1] = nj; out[3 * n + 2] = nk;
            ++n;
        }
    }
    return n;
}

/* kmc_event_rates.py:10-23 compute_misorientation */
double oracle_misorientation(double t1, double p1, double t2, double p2)
{
    double v1x = sin(t1) * cos(p1), v1y = sin(t1) * sin(p1), v1z = cos(t1);
    double v2x = sin(t2) * cos(p2), v2y = sin(t2) * sin(p2), v2z = cos(t2);
    double dot = v1x * v2x + v1y * v2y + v1z * v2z;
    dot = pymax(pymin(dot, 1.0), -1.0);
    return acos(dot);
}

typedef struct {
    uint8_t *type; int64_t *pos; double *rate; int64_t *target; int32_t *atom;
    int64_t cap, n;
} evbuf;

static inline void emit(evbuf *b, int type, int64_t pos, double rate, int64_t target, int atom)
{
    if (b->n < b->cap) {
        if (b->type) b->type[b->n] = (uint8_t)type;
        if (b->pos) b->pos[b->n] = pos;
        if (b->rate) b->rate[b->n] = rate;
        if (b->target) b->target[b->n] = target;
        if (b->atom) b->atom[b->n] = atom;
    }
    b->n++;
}

/*
 * kmc_event_rates.py:43-160 compute_row_events for plane i.
 * species_draws: the stream Numba's np.random.random() would produce
 * (kmc_event_rates.py:65); *draw_pos advances by one per finite-rate dep event.
 * If species_draws is NULL every dep event gets atom = states_w and the counter
 * still advances.
 */
static void plane_events(int64_t i, const int64_t *state, const double *theta, const double *phi,
                         const double *T, const int64_t *defects, int64_t L, const oracle_params *P,
                         const double *species_draws, int64_t *draw_pos, evbuf *b)
{
    const int64_t LL = L * L;
    const int64_t base = i * LL;
    const int64_t top_layer = L - 1;              /* kmc_event_rates.py:166 */
    int64_t nb[42];

    /* Deposition — kmc_event_rates.py:55-72 */
    if (i == top_layer) {
        for (int64_t j = 0; j < L; ++j)
            for (int64_t k = 0; k < L; ++k) {
                int64_t s = base + j * L + k;
                if (state[s] != 0) continue;
                double local_T = pymax(T[s], 1.0);
                double thermal_factor = exp(-(P->T_melt - local_T) / (P->kT * local_T));
                double eff = P->nu_dep * thermal_factor;
                if (!isfinite(eff)) continue;
                int atom = P->states_w;
                if (species_draws) {
                    double r = species_draws[*draw_pos];
                    if (r < P->impurity_c) atom = P->states_c;
                    else if (r < P->impurity_c + P->impurity_re) atom = P->states_re;
                    else atom = P->states_w;
                }
                (*draw_pos)++;
                emit(b, ORACLE_EV_DEP, s, eff, -1, atom);
            }
    }

    /* Occupied-site events (diffusion) — kmc_event_rates.py:75-109 */
    for (int64_t j = 0; j < L; ++j)
        for (int64_t k = 0; k < L; ++k) {
            int64_t s = base + j * L + k;
            int64_t atom = state[s];
            if (atom == 0) continue;
            if (atom == 4) continue;                          /* :80-81 */
            double E_b_atom, E_diff_atom;
            if (atom == P->states_w) { E_b_atom = P->E_b[0]; E_diff_atom = P->E_diff[0]; }
            else if (atom == P->states_re) { E_b_atom = P->E_b[1]; E_diff_atom = P->E_diff[1]; }
            else { E_b_atom = P->E_b[2]; E_diff_atom = P->E_diff[2]; }
            double local_T = pymax(T[s], 1.0);
            double defect_factor = 1.0 + (double)defects[s];
            int n = oracle_bcc_neighbors(i, j, k, L, nb);
            int64_t n_bonds = 0;
            for (int q = 0; q < n; ++q)
                if (state[nb[3 * q] * LL + nb[3 * q + 1] * L + nb[3 * q + 2]] != 0) n_bonds++;
            for (int q = 0; q < n; ++q) {
                int64_t t = nb[3 * q] * LL + nb[3 * q + 1] * L + nb[3 * q + 2];
                if (state[t] != 0) continue;
                double neighbor_T = pymax(T[t], 1.0);
                double dT = fabs(local_T - neighbor_T);
                double denom = pymax(P->T_melt - neighbor_T, 1.0);
                double grad_factor = 1.0 + 0.1 * dT / denom;
                double E_tot = pymax(E_diff_atom + 0.1 * (double)n_bonds * E_b_atom, 0.0);
                double rate = P->nu * grad_factor * exp(-defect_factor * E_tot / (P->kT * local_T));
                if (rate > P->rate_threshold && isfinite(rate))
                    emit(b, ORACLE_EV_DIFF, s, rate, t, (int)atom);
            }
        }

    /* Empty-site events (nucleation + attachment) — kmc_event_rates.py:112-158 */
    for (int64_t j = 0; j < L; ++j)
        for (int64_t k = 0; k < L; ++k) {
            int64_t s = base + j * L + k;
            if (state[s] != 0) continue;
            double local_T = pymax(T[s], 1.0);
            double dT = P->T_melt - local_T;
            int n = oracle_bcc_neighbors(i, j, k, L, nb);
            if (dT > P->delta_T_c) {                           /* :120-132 */
                int64_t n_imp = 0;
                for (int q = 0; q < n; ++q) {
                    int64_t sn = state[nb[3 * q] * LL + nb[3 * q + 1] * L + nb[3 * q + 2]];
                    if (sn == P->states_re || sn == P->states_c) n_imp++;
                }
                int64_t len = n > 1 ? n : 1;
                double f_imp = pymin(P->max_imp_fraction, (double)n_imp / (double)len);
                double K_eff = P->k_nuc * (1.0 - P->beta_imp_nuc * f_imp);
                K_eff = pymax(0.1 * P->k_nuc, pymin(P->k_nuc, K_eff));
                double barrier = K_eff / pymax((dT + 1e-6) * (dT + 1e-6), 1e-6);
                double rate = P->i0 * exp(-barrier / (P->kT * local_T));
                if (rate > P->rate_threshold && isfinite(rate))
                    emit(b, ORACLE_EV_NUC, s, rate, -1, P->states_w);
            }
            for (int q = 0; q < n; ++q) {                      /* :135-158 */
                int64_t t = nb[3 * q] * LL + nb[3 * q + 1] * L + nb[3 * q + 2];
                int64_t na = state[t];
                if (na == 0) continue;
                int ia;
                if (na == P->states_w) ia = 0;
                else if (na == P->states_re) ia = 1;
                else if (na == P->states_c) ia = 2;
                else continue;
                double mis = oracle_misorientation(theta[s], phi[s], theta[t], phi[t]);
                int64_t km = k - 1 > 0 ? k - 1 : 0;
                int64_t kp = k + 1 < L - 1 ? k + 1 : L - 1;
                double grad_z = (T[base + j * L + kp] - T[base + j * L + km]) * 0.5;
                double gf = pymax(0.0, grad_z) / pymax(P->T_melt - local_T, 1.0);
                double E_att = 0.5 * P->E_b[ia] * (1.0 - cos(mis));
                double rate = P->nu * exp(-E_att / (P->kT * local_T)) * (1.0 + P->anisotropy * gf);
                if (rate > P->rate_threshold && isfinite(rate))
                    emit(b, ORACLE_EV_ATT, s, rate, t, (int)na);
            }
        }
}

/* kmc_event_rates.py:162-176 get_event_rates — events in the reference's list order. */
int64_t oracle_event_rates(const int64_t *state, const double *theta, const double *phi,
                           const double *T, const int64_t *defects, int64_t L,
                           const oracle_params *P, const double *species_draws,
                           int64_t *draws_used, uint8_t *type, int64_t *pos, double *rate,
                           int64_t *target, int32_t *atom, int64_t cap)
{
    evbuf b = {type, pos, rate, target, atom, cap, 0};
    int64_t dp = 0;
    for (int64_t i = 0; i < L; ++i)
        plane_events(i, state, theta, phi, T, defects, L, P, species_draws, &dp, &b);
    if (draws_used) *draws_used = dp;
    return b.n;
}

/*
 * Per-site rate totals (the quantity the GPU rate kernel stores): for every site
 * the sum of its diff events (occupied) or nuc+att events (empty), added in list
 * order; dep[j*L+k] = dep rate on the top plane or NaN where no dep event exists.
 * Planes are independent -> OpenMP over i.  Returns the number of events.
 */
int64_t oracle_site_rates(const int64_t *state, const double *theta, const double *phi,
                          const double *T, const int64_t *defects, int64_t L,
                          const oracle_params *P, double *site_rate, double *dep_rate,
                          int32_t *site_nev)
{
    const int64_t LL = L * L;
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int64_t i = 0; i < L; ++i) {
        int64_t cap = 16 * LL;
        uint8_t *ty = (uint8_t *)malloc(cap);
        int64_t *po = (int64_t *)malloc(cap * sizeof(int64_t));
        double *ra = (double *)malloc(cap * sizeof(double));
        evbuf b = {ty, po, ra, NULL, NULL, cap, 0};
        int64_t dp = 0;
        plane_events(i, state, theta, phi, T, defects, L, P, NULL, &dp, &b);
        for (int64_t s = i * LL; s < (i + 1) * LL; ++s) { site_rate[s] = 0.0; if (site_nev) site_nev[s] = 0; }
        if (i == L - 1 && dep_rate)
            for (int64_t q = 0; q < LL; ++q) dep_rate[q] = NAN;
        for (int64_t e = 0; e < b.n; ++e) {
            if (ty[e] == ORACLE_EV_DEP) { if (dep_rate) dep_rate[po[e] - i * LL] = ra[e]; }
            else site_rate[po[e]] += ra[e];
            if (site_nev) site_nev[po[e]]++;
        }
        total += b.n;
        free(ty); free(po); free(ra);
    }
    return total;
}

/* CPython >= 3.12 builtin sum() over floats (Neumaier compensation), as used at
 * kmc_simulation.py:259.  Start value is the int 0, i.e. the first float. */
double oracle_pysum(const double *x, int64_t n)
{
    if (n == 0) return 0.0;
    double f = x[0], c = 0.0;
    for (int64_t q = 1; q < n; ++q) {
        double v = x[q];
        double t = f + v;
        if (fabs(f) >= fabs(v)) c += (f - t) + v;
        else c += (v - t) + f;
        f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
}

/*
 * thermal_solver.py:107-117 update_temperature_cet (scipy.ndimage.laplace,
 * mode='reflect' == edge replicate for a 3-tap kernel).  Operation order as
 * verified bit-identical to scipy 1.18.1: per axis  c*(-2) + (l + r); axes summed
 * (t0 + t1) + t2; * inv_dx2; * (dt*ALPHA) (scalar product formed first); T + ...; clip.
 * nan_to_num != 0 applies kmc_simulation.py:249 (np.nan_to_num(T, nan=nan_value)) first.
 */
void oracle_thermal_cet(const double *Tin, double *Tout, int64_t n0, int64_t n1, int64_t n2,
                        double dt_alpha, double inv_dx2, double lo, double hi,
                        int nan_to_num, double nan_value)
{
    const int64_t s0 = n1 * n2, s1 = n2;
    const double *T = Tin;
    double *tmp = NULL;
    if (nan_to_num) {
        tmp = (double *)malloc(sizeof(double) * n0 * n1 * n2);
        for (int64_t q = 0; q < n0 * n1 * n2; ++q) {
            double v = Tin[q];
            if (isnan(v)) v = nan_value;
            else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
            tmp[q] = v;
        }
        T = tmp;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n0; ++i)
        for (int64_t j = 0; j < n1; ++j)
            for (int64_t k = 0; k < n2; ++k) {
                int64_t s = i * s0 + j * s1 + k;
                double c = T[s];
                double l0 = T[(i > 0 ? i - 1 : 0) * s0 + j * s1 + k];
                double r0 = T[(i < n0 - 1 ? i + 1 : n0 - 1) * s0 + j * s1 + k];
                double l1 = T[i * s0 + (j > 0 ? j - 1 : 0) * s1 + k];
                double r1 = T[i * s0 + (j < n1 - 1 ? j + 1 : n1 - 1) * s1 + k];
                double l2 = T[i * s0 + j * s1 + (k > 0 ? k - 1 : 0)];
                double r2 = T[i * s0 + j * s1 + (k < n2 - 1 ? k + 1 : n2 - 1)];
                double t0 = c * -2.0 + (l0 + r0);
                double t1 = c * -2.0 + (l1 + r1);
                double t2 = c * -2.0 + (l2 + r2);
                double lap = (t0 + t1) + t2;
                lap = lap * inv_dx2;
                double v = c + dt_alpha * lap;
                /* np.clip == minimum(maximum(v, lo), hi); NaN propagates */
                if (v < lo) v = lo;
                if (v > hi) v = hi;
                Tout[s] = v;
            }
    free(tmp);
}

/*
 * thermal_solver.py:36-105 update_temperature (laser + latent heat variant).
 *   dTdt = ALPHA*lap + q_vol/(RHO*CP) + (200e3/CP)*dF_dt ; new_T = T + dt*dTdt ; clip
 * q_top[j*n2+k] = I_surface/VOXEL_SIZE (host-computed, thermal_solver.py:80-94),
 * applied on plane i = n0-1 only; dF_dt = mask / max(dt, 1e-12) (:97-99).
 */
void oracle_thermal_full(const double *T, const int64_t *state, const int64_t *prev_state,
                         double *Tout, int64_t n0, int64_t n1, int64_t n2, double dt, double alpha,
                         double inv_dx2, const double *q_top, double rho_cp, double latent_over_cp,
                         double lo, double hi)
{
    const int64_t s0 = n1 * n2, s1 = n2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n0; ++i)
        for (int64_t j = 0; j < n1; ++j)
            for (int64_t k = 0; k < n2; ++k) {
                int64_t s = i * s0 + j * s1 + k;
                double c = T[s];
                double l0 = T[(i > 0 ? i - 1 : 0) * s0 + j * s1 + k];
                double r0 = T[(i < n0 - 1 ? i + 1 : n0 - 1) * s0 + j * s1 + k];
                double l1 = T[i * s0 + (j > 0 ? j - 1 : 0) * s1 + k];
                double r1 = T[i * s0 + (j < n1 - 1 ? j + 1 : n1 - 1) * s1 + k];
                double l2 = T[i * s0 + j * s1 + (k > 0 ? k - 1 : 0)];
                double r2 = T[i * s0 + j * s1 + (k < n2 - 1 ? k + 1 : n2 - 1)];
                double t0 = c * -2.0 + (l0 + r0);
                double t1 = c * -2.0 + (l1 + r1);
                double t2 = c * -2.0 + (l2 + r2);
                double lap = ((t0 + t1) + t2) * inv_dx2;
                double q = (i == n0 - 1) ? q_top[j * n2 + k] : 0.0;
                double mask = (prev_state[s] == 0 && state[s] != 0) ? 1.0 : 0.0;
                double dFdt = mask / pymax(dt, 1e-12);   /* mask.astype(float) / max(dt,1e-12) */
                double dTdt = (alpha * lap + q / rho_cp) + latent_over_cp * dFdt;
                double v = c + dt * dTdt;
                if (v < lo) v = lo;
                if (v > hi) v = hi;
                Tout[s] = v;
            }
}

/*
 * kmc_simulation.py:246-332 — the per-step body of run_kmc with every random
 * draw injected:
 *   py_draws   : Python `random.random()` stream — u1 (:265), u2 iff
 *                defect_fraction > 0 (:323), u3 (:331), in that order per step;
 *   np_draws   : NumPy global stream — theta (:283/:308) then phi (:284/:309),
 *                consumed only by dep and nuc events (theta = pi*u, phi = 2pi*u);
 *   sp_draws   : Numba's stream for the deposited species (kmc_event_rates.py:65),
 *                one per finite-rate dep event per step.
 * Runs steps [step0, step0+n_steps); the caller handles the 200-step metric
 * cadence.  Returns the number of steps completed (a step that terminates at
 * :260-262 is not counted; *terminated is set).
 * log_* (may be NULL): chosen event per step.
 */
int64_t oracle_kmc_run(int64_t *state, int64_t *atom_type, double *theta, double *phi, double *T,
                       const int64_t *defects, int64_t L, const oracle_params *P,
                       int64_t step0, int64_t n_steps, double defect_fraction,
                       const oracle_thermal_params *TP,
                       const double *py_draws, int64_t *py_pos,
                       const double *np_draws, int64_t *np_pos,
                       const double *sp_draws, int64_t *sp_pos,
                       double *total_time, int64_t *nucleation_count, int *terminated,
                       uint8_t *log_type, int64_t *log_pos, int64_t *log_target,
                       int32_t *log_atom, double *log_rate, double *log_total)
{
    const int64_t N = L * L * L;
    const int64_t cap = 16 * N;
    uint8_t *ty = (uint8_t *)malloc(cap);
    int64_t *po = (int64_t *)malloc(cap * sizeof(int64_t));
    double *ra = (double *)malloc(cap * sizeof(double));
    int64_t *ta = (int64_t *)malloc(cap * sizeof(int64_t));
    int32_t *at = (int32_t *)malloc(cap * sizeof(int32_t));
    double *Tn = (double *)malloc(N * sizeof(double));
    const double PI = 3.141592653589793;
    int64_t done = 0;
    *terminated = 0;
    for (int64_t step = step0; step < step0 + n_steps; ++step) {
        if (step % TP->every == 0) {                         /* kmc_simulation.py:248-250 */
            oracle_thermal_cet(T, Tn, L, L, L, TP->dt_alpha, TP->inv_dx2, TP->lo, TP->hi, 1, TP->nan_value);
            memcpy(T, Tn, N * sizeof(double));
        }
        int64_t used = 0;
        int64_t n = oracle_event_rates(state, theta, phi, T, defects, L, P,
                                       sp_draws ? sp_draws + *sp_pos : NULL, &used, ty, po, ra, ta, at, cap);
        *sp_pos += used;
        double total = oracle_pysum(ra, n);                  /* :259 */
        if (n == 0 || total < 1e-25 || !isfinite(total)) {   /* :260-262 */
            *terminated = 1;
            break;
        }
        double r = py_draws[(*py_pos)++] * total;            /* :265 */
        double cum = 0.0;
        int64_t chosen = -1;
        for (int64_t e = 0; e < n; ++e) {                    /* :266-272 */
            cum += ra[e];
            if (cum >= r) { chosen = e; break; }
        }
        if (chosen < 0) chosen = n - 1;                      /* :273-274 */
        int64_t s = po[chosen], t = ta[chosen];
        int et = ty[chosen], atom = at[chosen];
        int64_t upd = s;
        if (et == ORACLE_EV_DEP || et == ORACLE_EV_NUC) {    /* :280-284, :305-310 */
            state[s] = atom; atom_type[s] = atom;
            theta[s] = 0.0 + (PI - 0.0) * np_draws[(*np_pos)++];
            phi[s] = 0.0 + (2 * PI - 0.0) * np_draws[(*np_pos)++];
            if (et == ORACLE_EV_NUC) (*nucleation_count)++;
        } else if (et == ORACLE_EV_DIFF) {                   /* :292-303 */
            state[t] = state[s]; atom_type[t] = atom_type[s];
            theta[t] = theta[s]; phi[t] = phi[s];
            state[s] = 0; atom_type[s] = 0; theta[s] = 0.0; phi[s] = 0.0;
            upd = t;
        } else {                                             /* att :312-317 */
            state[s] = atom; atom_type[s] = atom;
            theta[s] = theta[t]; phi[s] = phi[t];
        }
        if (defect_fraction > 0.0) {                         /* :323-327 */
            double u2 = py_draws[(*py_pos)++];
            if (u2 < defect_fraction) {
                atom_type[upd] = 4; state[upd] = 4; theta[upd] = 0.0; phi[upd] = 0.0;
            }
        }
        double u3 = py_draws[(*py_pos)++];                   /* :331-332 */
        double dt = pymax(-log(pymax(1e-12, u3)) / total, 1e-12);
        *total_time += dt;
        if (log_type) log_type[done] = (uint8_t)et;
        if (log_pos) log_pos[done] = s;
        if (log_target) log_target[done] = t;
        if (log_atom) log_atom[done] = atom;
        if (log_rate) log_rate[done] = ra[chosen];
        if (log_total) log_total[done] = total;
        done++;
    }
    free(ty); free(po); free(ra); free(ta); free(at); free(Tn);
    return done;
}

/* ---------------------------------------------------------------------------------------
 * Grain clustering (utils.py:28-84): DFS with an explicit stack over the 14-offset
 * neighbourhood, joining occupied sites whose misorientation (kmc_event_rates.py:10-23) is
 * below theta_threshold.  Cubic lattice (L, L, L).  visited[s] receives the cluster number
 * 1.. in discovery (raster) order, 0 on empty sites; sizes[q] / box_lo / box_hi [3q+axis]
 * describe cluster q+1 (sizes may be NULL).  Returns the number of clusters; cap bounds the
 * per-cluster outputs.
 * --------------------------------------------------------------------------------------- */
int64_t oracle_clusters(const int64_t *state, const double *theta, const double *phi, int64_t L,
                        double theta_threshold, int32_t *visited, int64_t cap, int32_t *sizes,
                        int32_t *box_lo, int32_t *box_hi)
{
    const int64_t N = L * L * L;
    int64_t *stack = (int64_t *)malloc((size_t)(N > 0 ? N : 1) * sizeof(int64_t));
    int64_t label = 0;
    for (int64_t s = 0; s < N; ++s) visited[s] = 0;
    for (int64_t s0 = 0; s0 < N; ++s0) {                              /* utils.py:35-38 raster scan */
        if (state[s0] == 0 || visited[s0] != 0) continue;
        ++label;
        int64_t top = 0, size = 1;
        int32_t lo[3] = {(int32_t)(s0 / (L * L)), (int32_t)((s0 / L) % L), (int32_t)(s0 % L)};
        int32_t hi[3] = {lo[0], lo[1], lo[2]};
        stack[top++] = s0;
        visited[s0] = (int32_t)label;
        while (top > 0) {                                              /* :44-63 */
            const int64_t c = stack[--top];
            const int64_t ci = c / (L * L), cj = (c / L) % L, ck = c % L;
            int64_t nb[14 * 3];
            const int n = oracle_bcc_neighbors(ci, cj, ck, L, nb);
            for (int q = 0; q < n; ++q) {
                const int64_t t = (nb[3 * q] * L + nb[3 * q + 1]) * L + nb[3 * q + 2];
                if (state[t] == 0 || visited[t] != 0) continue;
                const double mis = phi ? oracle_misorientation(theta[c], phi[c], theta[t], phi[t])
                                       : fabs(theta[c] - theta[t]);
                if (mis < theta_threshold) {
                    visited[t] = (int32_t)label;
                    stack[top++] = t;
                    ++size;
                    for (int ax = 0; ax < 3; ++ax) {
                        const int32_t v = (int32_t)nb[3 * q + ax];
                        if (v < lo[ax]) lo[ax] = v;
                        if (v > hi[ax]) hi[ax] = v;
                    }
                }
            }
        }
        if (label <= cap) {
            if (sizes) sizes[label - 1] = (int32_t)size;
            for (int ax = 0; ax < 3; ++ax) {
                if (box_lo) box_lo[3 * (label - 1) + ax] = lo[ax];
                if (box_hi) box_hi[3 * (label - 1) + ax] = hi[ax];
            }
        }
    }
    free(stack);
    return label;
}
