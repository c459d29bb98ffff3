// kmc_exact.cu — rejection-free BKL / n-fold-way steps in the reference's event order
// (kmc_simulation.py:246-332), all random draws injected.
//
// The reference rebuilds the full event list (O(L^3)) for every executed event and scans it
// linearly.  Here the per-site rate sums stay resident in HBM with a 3-level sum hierarchy
// (rates.cu); one persistent CTA executes a run of steps back to back:
//     search   : first list position whose prefix sum >= u1*total — block-wide warp-shuffle
//                scans over plane segments -> rows -> sites, then the <= 15 events of the site;
//     apply    : kmc_simulation.py:280-327 on the packed voxel byte and theta/phi;
//     refresh  : only the <= 30 sites within neighbour reach of the changed site(s) are
//                re-evaluated, then their rows, planes and the total, each with the same
//                fixed-order routine the dense rebuild uses (bit-identical hierarchy);
//     time     : dt = max(-ln(max(1e-12,u3))/total, 1e-12).
// The thermal update every `thermal_every` steps (kmc_simulation.py:248-250) and the dense
// rate rebuild it forces are launched between runs by the host loop in cet_kmc_run; no host
// synchronisation happens until the whole call has been enqueued.
#include "ctx.cuh"
#include "reduce.cuh"

namespace cet {

int thermal_cet_step(cet_ctx *c, const cet_thermal_params *p, const int32_t *stop_flag);
int rates_build(cet_ctx *c);

constexpr int KX_THREADS = 512;
constexpr int KX_MAX_AFFECTED = 32;

struct KmcArgs {
    uint8_t *vox;
    double *theta, *phi;
    Vec4 *v;
    const double *T;
    double *site_rate, *dep_rate, *row_occ, *row_emp, *row_dep, *seg, *total;
    int32_t *row_depcnt;
    cet_rate_params P;
    int L;
    KmcState *ks;
    long long n_steps;
    double defect_fraction;
    const double *py, *npd, *sp;
    long long n_py, n_np, n_sp;
    // log (may be NULL), indexed by ks->steps_done at entry of the step
    uint8_t *log_type; long long *log_pos, *log_target; int32_t *log_atom; double *log_rate, *log_total;
    long long log_cap;
};

struct StepShared {
    FindScratch fs;
    double red[40];
    long long redi[40];
    int affected[KX_MAX_AFFECTED];   // linear site indices
    int n_affected;
    int changed[2], n_changed;
    int stop;
    double total, r;
    long long n_dep;
    int ev_type, ev_pos, ev_target, ev_atom;
    double ev_rate;
};

__global__ void __launch_bounds__(KX_THREADS) kmc_steps_kernel(const KmcArgs a)
{
    __shared__ StepShared sh;
    const int L = a.L, LL = L * L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = KX_THREADS / 32;
    Lat g;
    g.vox = a.vox; g.v = a.v; g.T = a.T; g.L = L; g.n0 = L; g.i_off = 0;
    const bool use_u2 = a.defect_fraction > 0.0;
    const int py_per_step = use_u2 ? 3 : 2;

    for (long long it = 0; it < a.n_steps; ++it) {
        // ---- totals, termination, draw budget -------------------------------------------
        if (tid == 0) {
            KmcState *ks = a.ks;
            sh.stop = 0;
            if (ks->terminated || ks->starved) sh.stop = 1;
            else {
                const double total = a.total[0];
                const long long n_dep = ((const long long *)a.total)[1];
                sh.total = total; sh.n_dep = n_dep;
                if (ks->py_pos + py_per_step > a.n_py || ks->np_pos + 2 > a.n_np ||
                    (a.sp != nullptr && ks->sp_pos + n_dep > a.n_sp)) {
                    ks->starved = 1; sh.stop = 1;
                } else {
                    ks->sp_pos += n_dep;                       // drawn inside get_event_rates (:65)
                    ks->last_total_rate = total;
                    if (total < 1e-25 || !finite_f64(total)) {    // kmc_simulation.py:260-262
                        ks->terminated = 1; sh.stop = 1;
                    } else {
                        sh.r = a.py[ks->py_pos++] * total;     // :265
                    }
                }
            }
        }
        __syncthreads();
        if (sh.stop) return;
        const double r = sh.r;

        // ---- search: plane segment -> row -> site ----------------------------------------
        double excl;
        bool clamped;
        const int sidx = block_find_first(3 * L, 0.0, r, [&](int q) { return a.seg[q]; }, &sh.fs, &excl, &clamped);
        __syncthreads();
        const int p = sidx / 3, segm = sidx % 3;   // sidx >= 0 because total > 0
        const double *rowsum = segm == 0 ? a.row_dep : (segm == 1 ? a.row_occ + p * L : a.row_emp + p * L);
        const int j = block_find_first(L, excl, r, [&](int q) { return rowsum[q]; }, &sh.fs, &excl, &clamped);
        __syncthreads();
        const int rbase = (p * L + j) * L;
        int k;
        if (segm == 0) {
            k = block_find_first(L, excl, r, [&](int q) { const double v = a.dep_rate[j * L + q]; return v == v ? v : 0.0; },
                                 &sh.fs, &excl, &clamped);
        } else {
            const bool want_occ = segm == 1;
            k = block_find_first(L, excl, r,
                                 [&](int q) { return ((vox_state(a.vox[rbase + q]) != 0) == want_occ) ? a.site_rate[rbase + q] : 0.0; },
                                 &sh.fs, &excl, &clamped);
        }
        __syncthreads();

        // ---- event inside the site; apply (:280-327); time (:331-332) ---------------------
        if (segm == 0 && warp == 0) {
            // index of this dep event in the species stream = number of dep events before it
            long long before = 0;
            for (int q = lane; q < j; q += 32) before += a.row_depcnt[q];
            for (int q = lane; q < k; q += 32) { const double v = a.dep_rate[j * L + q]; if (v == v) ++before; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
            if (lane == 0) sh.redi[0] = before;
        }
        __syncthreads();
        if (tid == 0) {
            KmcState *ks = a.ks;
            const int s = rbase + k;
            int ety = -1, eslot = -1, eatom = 0;
            double erate = 0.0;
            if (segm == 0) {
                ety = CET_EV_DEP; erate = a.dep_rate[j * L + k];
                eatom = a.P.states_w;
                if (a.sp) eatom = dep_species(a.P, a.sp[ks->sp_pos - sh.n_dep + sh.redi[0]]);
            } else {
                double cum = excl;
                bool found = false;
                site_events(g, a.P, p, j, k, [&](int ty, int slot, double rate, int atom) {
                    if (found) return;
                    cum += rate;
                    ety = ty; eslot = slot; eatom = atom; erate = rate;   // keeps the last one as fallback
                    if (cum >= r) found = true;
                });
                if (!found) ks->fallback_last++;
            }
            int tgt = -1, upd = s;
            if (ety == CET_EV_DEP || ety == CET_EV_NUC) {                  // :280-284, :305-310
                a.vox[s] = (uint8_t)((a.vox[s] & 0xF0) | eatom);
                const double ut = a.npd[ks->np_pos], up = a.npd[ks->np_pos + 1];
                ks->np_pos += 2;
                const double th = __dadd_rn(0.0, __dmul_rn(3.141592653589793 - 0.0, ut));
                const double ph = __dadd_rn(0.0, __dmul_rn(2 * 3.141592653589793 - 0.0, up));
                a.theta[s] = th; a.phi[s] = ph;
                a.v[s] = unit_vec4(th, ph);
                if (ety == CET_EV_NUC) ks->nucleation_count++;
                sh.changed[0] = s; sh.n_changed = 1;
            } else if (ety == CET_EV_DIFF) {                               // :292-303
                tgt = s + CET_NB_DI(eslot) * LL + CET_NB_DJ(eslot) * L + CET_NB_DK(eslot);
                a.vox[tgt] = (uint8_t)((a.vox[tgt] & 0xF0) | (a.vox[s] & 0x0F));
                a.theta[tgt] = a.theta[s]; a.phi[tgt] = a.phi[s];
                a.v[tgt] = a.v[s];
                a.vox[s] = (uint8_t)(a.vox[s] & 0xF0);
                a.theta[s] = 0.0; a.phi[s] = 0.0;
                a.v[s] = Vec4{0.0, 0.0, 1.0, 0.0};
                upd = tgt;
                sh.changed[0] = s; sh.changed[1] = tgt; sh.n_changed = 2;
            } else {                                                       // att :312-317
                tgt = s + CET_NB_DI(eslot) * LL + CET_NB_DJ(eslot) * L + CET_NB_DK(eslot);
                a.vox[s] = (uint8_t)((a.vox[s] & 0xF0) | eatom);
                a.theta[s] = a.theta[tgt]; a.phi[s] = a.phi[tgt];
                a.v[s] = a.v[tgt];
                sh.changed[0] = s; sh.n_changed = 1;
            }
            if (use_u2) {                                                  // :323-327
                const double u2 = a.py[ks->py_pos++];
                if (u2 < a.defect_fraction) {
                    a.vox[upd] = (uint8_t)((a.vox[upd] & 0xF0) | a.P.defect_id);
                    a.theta[upd] = 0.0; a.phi[upd] = 0.0;
                    a.v[upd] = Vec4{0.0, 0.0, 1.0, 0.0};
                }
            }
            const double u3 = a.py[ks->py_pos++];                          // :331-332
            const double dt = pymax(-log(pymax(1e-12, u3)) / sh.total, 1e-12);
            ks->total_time += dt;
            const long long li = ks->steps_done;
            if (li < a.log_cap) {
                if (a.log_type) a.log_type[li] = (uint8_t)ety;
                if (a.log_pos) a.log_pos[li] = s;
                if (a.log_target) a.log_target[li] = tgt;
                if (a.log_atom) a.log_atom[li] = eatom;
                if (a.log_rate) a.log_rate[li] = erate;
                if (a.log_total) a.log_total[li] = sh.total;
            }
            ks->steps_done++;
            // affected sites: the changed ones and everything within neighbour reach of them
            int n = 0;
            for (int c = 0; c < sh.n_changed; ++c) {
                const int cs = sh.changed[c];
                const int ci = cs / LL, cj = (cs / L) % L, ck = cs % L;
                sh.affected[n++] = cs;
                for (int o = 0; o < 14; ++o) {
                    const int ni = ci + CET_NB_DI(o), nj = cj + CET_NB_DJ(o), nk = ck + CET_NB_DK(o);
                    if (ni >= 0 && ni < L && nj >= 0 && nj < L && nk >= 0 && nk < L)
                        sh.affected[n++] = (ni * L + nj) * L + nk;
                }
            }
            sh.n_affected = n;
        }
        __syncthreads();

        // ---- refresh: sites -> rows -> planes -> total -------------------------------------
        const int na = sh.n_affected;
        if (tid < na) {
            const int s = sh.affected[tid];
            const int si = s / LL, sj = (s / L) % L, sk = s % L;
            a.site_rate[s] = site_rate_sum(g, a.P, si, sj, sk, nullptr);
        }
        if (tid < sh.n_changed) {
            const int s = sh.changed[tid];
            if (s / LL == L - 1) {                                         // top plane: deposition event
                double v, rr = NAN;
                if (vox_state(a.vox[s]) == 0 && dep_rate(a.P, a.T[s], &v)) rr = v;
                a.dep_rate[s - (L - 1) * LL] = rr;
            }
        }
        __syncthreads();
        for (int q = warp; q < na; q += nwarps) {
            const int s = sh.affected[q];
            const int row = s / L;                                         // p*L + j
            double occ, emp;
            warp_row_sums(a.vox + (int64_t)row * L, a.site_rate + (int64_t)row * L, L, &occ, &emp);
            if (lane == 0) { a.row_occ[row] = occ; a.row_emp[row] = emp; }
        }
        for (int q = warp; q < sh.n_changed; q += nwarps) {
            const int s = sh.changed[q];
            if (s / LL == L - 1) {
                const int jj = (s / L) % L;
                double ds; int dc;
                warp_dep_row(a.dep_rate + jj * L, L, &ds, &dc);
                if (lane == 0) { a.row_dep[jj] = ds; a.row_depcnt[jj] = dc; }
            }
        }
        __syncthreads();
        for (int q = warp; q < na; q += nwarps) {
            const int pp = sh.affected[q] / LL;
            const double so = warp_strided_sum(a.row_occ + pp * L, L);
            const double se = warp_strided_sum(a.row_emp + pp * L, L);
            const double sd = (pp == L - 1) ? warp_strided_sum(a.row_dep, L) : 0.0;
            if (lane == 0) { a.seg[3 * pp] = sd; a.seg[3 * pp + 1] = so; a.seg[3 * pp + 2] = se; }
        }
        __syncthreads();
        {
            const double t = block_sum(a.seg, 3 * L, sh.red);
            const long long nd = block_sum_i(a.row_depcnt, L, sh.redi);
            if (tid == 0) { a.total[0] = t; ((long long *)a.total)[1] = nd; }
        }
        __syncthreads();
    }
}

}  // namespace cet

using namespace cet;

extern "C" int cet_kmc_run(cet_ctx *c, int64_t step0, int64_t n_steps, double defect_fraction,
                           const cet_thermal_params *tp, int32_t thermal_every,
                           const double *py_draws, int64_t n_py, const double *np_draws, int64_t n_np,
                           const double *sp_draws, int64_t n_sp, double total_time0, cet_kmc_result *res,
                           uint8_t *log_type, int64_t *log_pos, int64_t *log_target, int32_t *log_atom,
                           double *log_rate, double *log_total)
{
    CET_REQUIRE(c && res, "cet_kmc_run: NULL argument");
    CET_REQUIRE(c->cubic && c->have_rp, "cet_kmc_run: needs a cubic context with rate params");
    CET_REQUIRE(c->n0 == c->n1, "cet_kmc_run: the exact BKL path needs a cubic lattice");
    CET_REQUIRE(c->halo == 0 && c->i_begin == 0 && c->i_end == c->n0,
                "cet_kmc_run: the exact BKL path runs on one GPU holding the whole lattice");
    CET_REQUIRE(c->nloc < (1ll << 31), "cet_kmc_run: lattice too large for the exact path");
    CET_REQUIRE(n_steps >= 0 && (thermal_every <= 0 || tp != nullptr), "cet_kmc_run: bad arguments");
    CET_REQUIRE(py_draws && np_draws, "cet_kmc_run: py_draws / np_draws are required");
    cet::DeviceGuard dg(c->device);
    auto grow = [&](double **buf, size_t *cap, size_t n) -> int {
        if (*cap >= n) return 0;
        if (*buf) cudaFree(*buf);
        *buf = nullptr; *cap = 0;
        CET_CUDA(cudaMalloc(buf, (n + 16) * sizeof(double)));
        *cap = n + 16;
        return 0;
    };
    if (int rc = grow(&c->d_py, &c->cap_py, (size_t)n_py)) return rc;
    if (int rc = grow(&c->d_np, &c->cap_np, (size_t)n_np)) return rc;
    if (sp_draws && n_sp > 0) if (int rc = grow(&c->d_sp, &c->cap_sp, (size_t)n_sp)) return rc;
    if (n_py) CET_CUDA(cudaMemcpyAsync(c->d_py, py_draws, (size_t)n_py * 8, cudaMemcpyHostToDevice, c->stream));
    if (n_np) CET_CUDA(cudaMemcpyAsync(c->d_np, np_draws, (size_t)n_np * 8, cudaMemcpyHostToDevice, c->stream));
    if (sp_draws && n_sp > 0)
        CET_CUDA(cudaMemcpyAsync(c->d_sp, sp_draws, (size_t)n_sp * 8, cudaMemcpyHostToDevice, c->stream));
    const bool want_log = log_type || log_pos || log_target || log_atom || log_rate || log_total;
    const size_t per_log = 1 + 8 + 8 + 4 + 8 + 8;
    char *lg = nullptr;
    if (want_log && n_steps > 0) {
        const size_t need = (size_t)n_steps * per_log + 64;
        if (c->cap_log < need) {
            if (c->d_log) cudaFree(c->d_log);
            c->d_log = nullptr; c->cap_log = 0;
            CET_CUDA(cudaMalloc(&c->d_log, need));
            c->cap_log = need;
        }
        lg = (char *)c->d_log;
    }
    KmcState init;
    memset(&init, 0, sizeof(init));
    init.total_time = total_time0;
    CET_CUDA(cudaMemcpyAsync(c->kmc, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));

    KmcArgs a;
    memset(&a, 0, sizeof(a));
    a.P = c->rp; a.L = (int)c->n1; a.ks = c->kmc;
    a.defect_fraction = defect_fraction;
    a.py = c->d_py; a.npd = c->d_np; a.sp = (sp_draws && n_sp > 0) ? c->d_sp : nullptr;
    a.n_py = n_py; a.n_np = n_np; a.n_sp = n_sp;
    if (lg) {
        a.log_pos = (long long *)lg; a.log_target = a.log_pos + n_steps;
        a.log_rate = (double *)(a.log_target + n_steps); a.log_total = a.log_rate + n_steps;
        a.log_atom = (int32_t *)(a.log_total + n_steps); a.log_type = (uint8_t *)(a.log_atom + n_steps);
        a.log_cap = n_steps;
    }
    int64_t step = step0;
    const int64_t end = step0 + n_steps;
    while (step < end) {
        if (thermal_every > 0 && step % thermal_every == 0) {        // kmc_simulation.py:248-250
            if (int rc = thermal_cet_step(c, tp, &c->kmc->terminated)) return rc;
        }
        if (!c->rates_valid) if (int rc = rates_build(c)) return rc;
        int64_t run = end - step;
        if (thermal_every > 0) {
            const int64_t to_next = thermal_every - step % thermal_every;
            if (to_next < run) run = to_next;
        }
        // pointers are re-read every run: the thermal step swaps the T ping-pong buffers
        a.vox = c->vox; a.theta = c->theta; a.phi = c->phi; a.T = c->T;
        a.v = c->v;
        a.site_rate = c->site_rate; a.dep_rate = c->dep_rate; a.row_occ = c->row_occ; a.row_emp = c->row_emp;
        a.row_dep = c->row_dep; a.row_depcnt = c->row_depcnt; a.seg = c->seg; a.total = c->total;
        a.n_steps = run;
        kmc_steps_kernel<<<1, KX_THREADS, 0, c->stream>>>(a);
        CET_CUDA(cudaGetLastError());
        c->nst_valid = false;          // the step kernel edits states without maintaining the neighbour cache
        c->tile_valid = false;
        step += run;
    }
    KmcState out;
    CET_CUDA(cudaMemcpyAsync(&out, c->kmc, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t done = out.steps_done;
    if (lg && done > 0) {
        if (log_pos) CET_CUDA(cudaMemcpyAsync(log_pos, a.log_pos, (size_t)done * 8, cudaMemcpyDeviceToHost, c->stream));
        if (log_target) CET_CUDA(cudaMemcpyAsync(log_target, a.log_target, (size_t)done * 8, cudaMemcpyDeviceToHost, c->stream));
        if (log_rate) CET_CUDA(cudaMemcpyAsync(log_rate, a.log_rate, (size_t)done * 8, cudaMemcpyDeviceToHost, c->stream));
        if (log_total) CET_CUDA(cudaMemcpyAsync(log_total, a.log_total, (size_t)done * 8, cudaMemcpyDeviceToHost, c->stream));
        if (log_atom) CET_CUDA(cudaMemcpyAsync(log_atom, a.log_atom, (size_t)done * 4, cudaMemcpyDeviceToHost, c->stream));
        if (log_type) CET_CUDA(cudaMemcpyAsync(log_type, a.log_type, (size_t)done, cudaMemcpyDeviceToHost, c->stream));
        CET_CUDA(cudaStreamSynchronize(c->stream));
    }
    res->steps_done = out.steps_done; res->py_used = out.py_pos; res->np_used = out.np_pos;
    res->sp_used = out.sp_pos; res->nucleation_count = out.nucleation_count;
    res->fallback_last = out.fallback_last; res->total_time = out.total_time;
    res->last_total_rate = out.last_total_rate; res->terminated = out.terminated; res->starved = out.starved;
    return 0;
}
