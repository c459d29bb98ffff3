/* oracle.h — declarations for the CPU restatement (test infrastructure only). */
#ifndef CETKMC_ORACLE_H
#define CETKMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_EV_DEP = 0, ORACLE_EV_DIFF = 1, ORACLE_EV_NUC = 2, ORACLE_EV_ATT = 3 };

/* Values of constants.py as passed by kmc_event_rates.py:164-173 (+ the module
 * globals the jitted code closes over, kmc_event_rates.py:3-7). */
typedef struct {
    double nu, nu_dep;
    double E_b[3], E_diff[3];
    double kT, T_melt, i0, delta_T_c;
    double k_nuc, beta_imp_nuc, max_imp_fraction;
    double rate_threshold, anisotropy, impurity_re, impurity_c;
    int32_t states_w, states_re, states_c, pad_;
} oracle_params;

/* kmc_simulation.py:248-250 + thermal_solver.py:107-117 */
typedef struct {
    double dt_alpha, inv_dx2, lo, hi, nan_value;
    int64_t every;
} oracle_thermal_params;

int oracle_bcc_neighbors(int64_t i, int64_t j, int64_t k, int64_t L, int64_t *out);
double oracle_misorientation(double t1, double p1, double t2, double p2);
int64_t oracle_event_rates(const int64_t *state, const double *theta, const double *phi,
                           const double *T, const int64_t *defects, int64_t L,
                           const oracle_params *P, const double *species_draws,
                           int64_t *draws_used, uint8_t *type, int64_t *pos, double *rate,
                           int64_t *target, int32_t *atom, int64_t cap);
int64_t oracle_site_rates(const int64_t *state, const double *theta, const double *phi,
                          const double *T, const int64_t *defects, int64_t L,
                          const oracle_params *P, double *site_rate, double *dep_rate,
                          int32_t *site_nev);
double oracle_pysum(const double *x, int64_t n);
void oracle_thermal_cet(const double *Tin, double *Tout, int64_t n0, int64_t n1, int64_t n2,
                        double dt_alpha, double inv_dx2, double lo, double hi,
                        int nan_to_num, double nan_value);
void oracle_thermal_full(const double *T, const int64_t *state, const int64_t *prev_state,
                         double *Tout, int64_t n0, int64_t n1, int64_t n2, double dt, double alpha,
                         double inv_dx2, const double *q_top, double rho_cp, double latent_over_cp,
                         double lo, double hi);
int64_t oracle_kmc_run(int64_t *state, int64_t *atom_type, double *theta, double *phi, double *T,
                       const int64_t *defects, int64_t L, const oracle_params *P,
                       int64_t step0, int64_t n_steps, double defect_fraction,
                       const oracle_thermal_params *TP,
                       const double *py_draws, int64_t *py_pos,
                       const double *np_draws, int64_t *np_pos,
                       const double *sp_draws, int64_t *sp_pos,
                       double *total_time, int64_t *nucleation_count, int *terminated,
                       uint8_t *log_type, int64_t *log_pos, int64_t *log_target,
                       int32_t *log_atom, double *log_rate, double *log_total);
int64_t oracle_clusters(const int64_t *state, const double *theta, const double *phi, int64_t L,
                        double theta_threshold, int32_t *visited, int64_t cap, int32_t *sizes,
                        int32_t *box_lo, int32_t *box_hi);
#ifdef __cplusplus
}
#endif
#endif
