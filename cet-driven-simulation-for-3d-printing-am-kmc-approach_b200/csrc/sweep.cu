// sweep.cu — synchronous-sublattice KMC sweeps for large lattices (no reference counterpart:
// the reference executes ONE event per O(L^3) rate rebuild, kmc_simulation.py:246-332).
//
// One sweep visits every site once:
//   decide  (dense, HBM/FP64-bound): per site, the event rates of kmc_event_rates.py:43-160 are
//           evaluated from the sweep-start lattice; the site fires with p = 1-exp(-R_site*tau)
//           and picks one of its events with probability rate/R_site.  Draws come from a
//           counter-based Philox4x32-10 keyed by (seed, sweep, GLOBAL site index), so a site's
//           decision does not depend on which GPU evaluates it.  A fired event claims the
//           sites it will write (itself; a diffusion event also its target).
//   apply   (sparse): conflicts are resolved by sublattice order — the 5x5x5 checkerboard
//           colour of the source site, rotated every sweep, is the claim priority.  Two sites
//           of one colour differ by multiples of 5 per axis, while two events can only collide
//           when their sources differ by a neighbour offset or a difference of two offsets
//           (every coordinate <= 4), so a colour never conflicts with itself and the outcome
//           is deterministic.  Winners are applied exactly as kmc_simulation.py:280-327.
//   tau     for the next sweep from the totals of this one:
//           tau = min(events_per_sweep / R_total, -ln(1-p_max) / R_max).
// Slabs: a context with halo H >= 6 re-evaluates decisions of ghost sites to depth 4 and
// resolves claims to depth 2, which is everything that can write an owned site; one halo
// exchange per sweep (comm.cu) is the only data-path communication.  Plane sums are combined
// in a fixed order so the trajectory is independent of the number of slabs.
#include "ctx.cuh"
#include "dense_pass.cuh"
#include "reduce.cuh"

namespace cet {

int thermal_cet_step(cet_ctx *c, const cet_thermal_params *p, const int32_t *stop_flag);
int comm_sweep_reduce(cet_ctx *c, double *plane_sum, int n, double *max_inout);   // comm.cu
int comm_halo_exchange(cet_ctx *c, int fields);

// ---- Philox4x32-10 ------------------------------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };
__device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
// two uniforms in [0,1) with 53 random bits each
__device__ __forceinline__ void philox_u2(uint64_t seed, uint64_t site, uint32_t sweep, uint32_t stream,
                                          double *u0, double *u1)
{
    const u32x4 r = philox4x32_10(u32x4{(uint32_t)site, (uint32_t)(site >> 32), sweep, stream},
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    *u0 = (double)(a >> 11) * 1.1102230246251565e-16;
    *u1 = (double)(b >> 11) * 1.1102230246251565e-16;
}

struct Record {          // one fired event
    long long src;       // local linear index of the source site
    int32_t info;        // type | (slot+1) << 4 | atom << 12
    int32_t colour_rank;
    double theta, phi;   // orientation the written site receives (att: neighbour's; dep/nuc: drawn)
};

struct SweepArgs {
    Lat g;
    const double *theta, *phi;
    cet_rate_params P;
    SweepState *ss;
    Record *records;
    unsigned int cap_records;
    unsigned long long *claim;
    double *blk_sum, *blk_max;
    int p_lo, p_hi;          // local planes whose sites are evaluated
    int np;
    uint64_t seed;
    uint32_t sweep;
    int rows_per_blk, blks_per_plane;
};

constexpr int SW_WARPS = 8;

__device__ __forceinline__ int colour_rank(int i, int j, int k, uint32_t sweep)
{
    const int c = ((i % 5) * 5 + (j % 5)) * 5 + (k % 5);
    return (int)((c + 37u * sweep) % 125u);     // rotate the sublattice order every sweep
}
__device__ __forceinline__ unsigned long long claim_key(int rank, long long gsite)
{
    return ((unsigned long long)(125 - rank) << 48) | (unsigned long long)(gsite + 1);
}

// A site fired: pick one of its events with probability rate / R (list order: dep first, then the
// site's own events), record it and claim the sites it writes.
__device__ __noinline__ void fire_event(const SweepArgs &a, int i, int j, int k, int64_t s, long long gsite,
                                        double R, double dep, bool has_dep, double u_pick)
{
    const cet_rate_params &P = a.P;
    const int L = a.g.L;
    const int64_t LL = (int64_t)L * L;
    const double x = u_pick * R;
    double cum = 0.0;
    int ety = -1, eslot = -1, eatom = 0;
    bool found = false;
    if (has_dep) {
        cum = dep; ety = CET_EV_DEP; eatom = P.states_w;
        if (cum >= x) found = true;
    }
    if (!found)
        site_events(a.g, P, i, j, k, [&](int ty, int slot, double rate, int atom) {
            if (found) return;
            cum += rate; ety = ty; eslot = slot; eatom = atom;
            if (cum >= x) found = true;
        });
    if (ety < 0) return;
    Record rec;
    rec.src = s;
    rec.theta = 0.0; rec.phi = 0.0;
    int64_t tgt = -1;
    if (ety == CET_EV_DEP || ety == CET_EV_NUC) {
        double ut, up;
        philox_u2(a.seed, (uint64_t)gsite, a.sweep, 1u, &ut, &up);
        rec.theta = __dmul_rn(3.141592653589793, ut);          // np.random.uniform(0, pi)
        rec.phi = __dmul_rn(2 * 3.141592653589793, up);        // np.random.uniform(0, 2pi)
        if (ety == CET_EV_DEP) {                               // kmc_event_rates.py:65-71
            double us, unused;
            philox_u2(a.seed, (uint64_t)gsite, a.sweep, 2u, &us, &unused);
            eatom = dep_species(P, us);
        }
    } else {
        tgt = s + CET_NB_DI(eslot) * LL + CET_NB_DJ(eslot) * L + CET_NB_DK(eslot);
        if (ety == CET_EV_ATT) { rec.theta = a.theta[tgt]; rec.phi = a.phi[tgt]; }
        else { rec.theta = a.theta[s]; rec.phi = a.phi[s]; }
    }
    const int rank = colour_rank(i, j, k, a.sweep);
    rec.info = ety | ((eslot + 1) << 4) | (eatom << 12);
    rec.colour_rank = rank;
    const unsigned int slot = atomicAdd(&a.ss->n_records, 1u);
    if (slot >= a.cap_records) { a.ss->overflow = 1; return; }
    const unsigned long long key = claim_key(rank, gsite);
    atomicMax(&a.claim[s], key);
    if (ety == CET_EV_DIFF) atomicMax(&a.claim[tgt], key);
    a.records[slot] = rec;
}

// __grid_constant__: the argument block stays in constant memory even though fire_event takes
// its address (no per-thread stack copy).
__global__ void __launch_bounds__(SW_WARPS * 32) sweep_decide_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double s_sum[SW_WARPS], s_max[SW_WARPS];
    const int L = a.g.L;
    const int64_t LL = (int64_t)L * L;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pl = blockIdx.x / a.blks_per_plane, jb = blockIdx.x % a.blks_per_plane;
    const int p = a.p_lo + pl, j = jb * SW_WARPS + w;
    const int i = a.g.i_off + p;
    const bool stop = a.ss->terminated != 0;
    const double tau = a.ss->tau;
    double rsum = 0.0, rmax = 0.0;
    if (j < L && !stop) {
        const int64_t rbase = ((int64_t)p * L + j) * L;
        RowLists lists;
        lists.occ = (uint16_t *)(dyn_smem + (size_t)w * 2 * L * sizeof(uint16_t));
        lists.emp = lists.occ + L;
        row_classify(a.g, a.P, rbase, lists, nullptr);
        // one site: accumulate the totals, draw, and (rarely) fire
        auto visit = [&](int k, double R, bool has_dep, double dep, bool active) {
            bool fire = false;
            double u_pick = 0.0;
            const long long gsite = (long long)i * LL + (long long)j * L + k;
            if (active) {
                rsum += R;
                rmax = fmax(rmax, R);
                if (R > 0.0 && tau > 0.0) {
                    double u_fire;
                    philox_u2(a.seed, (uint64_t)gsite, a.sweep, 0u, &u_fire, &u_pick);
                    const double x = R * tau;                 // 1 - exp(-x) <= x: most sites reject here
                    if (u_fire < x) fire = u_fire < -expm1(-x);
                }
            }
            if (fire) fire_event(a, i, j, k, rbase + k, gsite, R, dep, has_dep, u_pick);
            __syncwarp();      // re-converge before the next chunk (keeps the dense part 32-wide)
        };
        row_occupied(a.g, a.P, i, j, rbase, lists,
                     [&](int k, double sum, bool active) { visit(k, sum, false, 0.0, active); });
        row_empty(a.g, a.P, i, j, rbase, lists, [&](int k, double sum, bool has_dep, double dep, bool active) {
            visit(k, has_dep ? dep + sum : sum, has_dep, dep, active);
        });
    }
    rsum = warp_sum(rsum);
    rmax = warp_max(rmax);
    if (lane == 0) { s_sum[w] = rsum; s_max[w] = rmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, m = 0.0;
        for (int q = 0; q < SW_WARPS; ++q) { t += s_sum[q]; m = fmax(m, s_max[q]); }
        a.blk_sum[blockIdx.x] = t;
        a.blk_max[blockIdx.x] = m;
    }
}

// One warp per evaluated plane: fixed-order sum of the plane's block partials.  plane_sum is
// indexed by GLOBAL plane so that every slab count produces the same numbers.
__global__ void sweep_plane_reduce_kernel(const double *blk_sum, const double *blk_max, int blks_per_plane,
                                          int n_planes, int first_global, int own_lo, int own_hi,
                                          double *plane_sum, double *max_out)
{
    const int pl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pl >= n_planes) return;
    const int lane = threadIdx.x & 31;
    const int gi = first_global + pl;
    double s = 0.0, m = 0.0;
    for (int q = lane; q < blks_per_plane; q += 32) {
        s += blk_sum[(int64_t)pl * blks_per_plane + q];
        m = fmax(m, blk_max[(int64_t)pl * blks_per_plane + q]);
    }
    s = warp_sum(s);
    m = warp_max(m);
    if (lane == 0 && gi >= own_lo && gi < own_hi) {   // ghost planes are another slab's to report
        plane_sum[gi] = s;
        atomicMax((unsigned long long *)max_out, (unsigned long long)__double_as_longlong(m));   // m >= 0
    }
}

// Single CTA: total in fixed plane order, tau for the next sweep, time bookkeeping.
__global__ void sweep_finalize_kernel(SweepState *ss, const double *plane_sum, int L, const double *max_in,
                                      double events_per_sweep, double p_max)
{
    __shared__ double sm[40];
    const double total = block_sum(plane_sum, L, sm);
    if (threadIdx.x == 0 && !ss->terminated) {
        const double rmax = *max_in;
        ss->time += ss->tau;                       // the sweep that just ran advanced time by its tau
        ss->sum_rate = total; ss->max_rate = rmax;
        if (total < 1e-25 || !finite_f64(total)) { // kmc_simulation.py:260-262
            ss->terminated = 1; ss->tau = 0.0;
        } else {
            double tau = events_per_sweep / total;
            const double cap = -log1p(-p_max) / rmax;
            if (cap < tau) tau = cap;
            ss->tau = tau;
        }
    }
}

struct ApplyArgs {
    uint8_t *vox;
    double *theta, *phi, *vx, *vy, *vz;
    SweepState *ss;
    const Record *records;
    unsigned int cap_records;
    unsigned long long *claim;
    cet_rate_params P;
    int L, i_off;
    int c_lo, c_hi;        // local planes with complete claims
    int own_lo, own_hi;    // local planes owned by this slab (for the counters)
    uint64_t seed;
    uint32_t sweep;
    double defect_fraction;
};

__global__ void sweep_apply_kernel(const ApplyArgs a)
{
    const unsigned int n = min(a.ss->n_records, a.cap_records);
    const int64_t LL = (int64_t)a.L * a.L;
    unsigned long long fired = 0, applied = 0, nuc = 0;
    for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const Record rec = a.records[q];
        const int ety = rec.info & 15, eslot = ((rec.info >> 4) & 255) - 1, eatom = rec.info >> 12;
        const int64_t s = rec.src;
        const int p = (int)(s / LL), j = (int)((s / a.L) % a.L), k = (int)(s % a.L);
        const long long gsite = (long long)(a.i_off + p) * LL + (long long)j * a.L + k;
        const unsigned long long key = claim_key(rec.colour_rank, gsite);
        int64_t tgt = -1;
        int pt = p;
        if (ety == CET_EV_DIFF || ety == CET_EV_ATT) {
            tgt = s + CET_NB_DI(eslot) * LL + CET_NB_DJ(eslot) * a.L + CET_NB_DK(eslot);
            if (ety == CET_EV_DIFF) pt = p + CET_NB_DI(eslot);
        }
        const bool owned = p >= a.own_lo && p < a.own_hi;
        if (owned) ++fired;
        const bool complete = p >= a.c_lo && p < a.c_hi && pt >= a.c_lo && pt < a.c_hi;
        const unsigned long long cs = a.claim[s];
        const unsigned long long ct = (ety == CET_EV_DIFF) ? a.claim[tgt] : key;
        const bool win = complete && cs == key && ct == key;
        if (win) {
            int64_t upd = s;
            double ux, uy, uz;
            unit_vector(rec.theta, rec.phi, &ux, &uy, &uz);          // same bits as the source's resident vector
            if (ety == CET_EV_DIFF) {                                    // kmc_simulation.py:292-303
                a.vox[tgt] = (uint8_t)((a.vox[tgt] & 0xF0) | (a.vox[s] & 0x0F));
                a.theta[tgt] = rec.theta; a.phi[tgt] = rec.phi;
                a.vx[tgt] = ux; a.vy[tgt] = uy; a.vz[tgt] = uz;
                a.vox[s] = (uint8_t)(a.vox[s] & 0xF0);
                a.theta[s] = 0.0; a.phi[s] = 0.0;
                a.vx[s] = 0.0; a.vy[s] = 0.0; a.vz[s] = 1.0;
                upd = tgt;
            } else {                                                     // dep / nuc / att
                a.vox[s] = (uint8_t)((a.vox[s] & 0xF0) | eatom);
                a.theta[s] = rec.theta; a.phi[s] = rec.phi;
                a.vx[s] = ux; a.vy[s] = uy; a.vz[s] = uz;
                if (ety == CET_EV_NUC && owned) ++nuc;
            }
            if (a.defect_fraction > 0.0) {                               // :323-327
                double u2, unused;
                philox_u2(a.seed, (uint64_t)gsite, a.sweep, 3u, &u2, &unused);
                if (u2 < a.defect_fraction) {
                    a.vox[upd] = (uint8_t)((a.vox[upd] & 0xF0) | a.P.defect_id);
                    a.theta[upd] = 0.0; a.phi[upd] = 0.0;
                    a.vx[upd] = 0.0; a.vy[upd] = 0.0; a.vz[upd] = 1.0;
                }
            }
            if (owned) ++applied;
        }
        // release the claims this event holds (only the top claimant of a site clears it)
        if (cs == key) a.claim[s] = 0ull;
        if (ety == CET_EV_DIFF && ct == key) a.claim[tgt] = 0ull;
    }
    fired = warp_sum_i((int)fired); applied = warp_sum_i((int)applied); nuc = warp_sum_i((int)nuc);
    if ((threadIdx.x & 31) == 0) {
        if (fired) atomicAdd(&a.ss->n_fired, fired);
        if (applied) atomicAdd(&a.ss->n_applied, applied);
        if (nuc) atomicAdd(&a.ss->n_nuc, nuc);
    }
}

__global__ void sweep_reset_kernel(SweepState *ss, double *max_slot)
{
    ss->n_records = 0;
    *max_slot = 0.0;
}

static int sweep_alloc(cet_ctx *c)
{
    if (!c->sweep) {
        CET_CUDA(cudaMalloc(&c->sweep, sizeof(SweepState)));
        CET_CUDA(cudaMemsetAsync(c->sweep, 0, sizeof(SweepState), c->stream));
    }
    if (!c->claim) {
        CET_CUDA(cudaMalloc(&c->claim, (size_t)c->nloc * sizeof(unsigned long long)));
        CET_CUDA(cudaMemsetAsync(c->claim, 0, (size_t)c->nloc * sizeof(unsigned long long), c->stream));
    }
    if (!c->records) {
        size_t cap = (size_t)c->nloc / 4 + 4096;
        if (cap > 0x7fffffffull) cap = 0x7fffffffull;
        CET_CUDA(cudaMalloc(&c->records, cap * sizeof(Record)));
        c->cap_records = cap;
    }
    const int bpp = (int)((c->n1 + SW_WARPS - 1) / SW_WARPS);
    if (!c->blk_sum) {
        c->n_blk = (int64_t)bpp * c->np;
        CET_CUDA(cudaMalloc(&c->blk_sum, (size_t)c->n_blk * sizeof(double)));
        CET_CUDA(cudaMalloc(&c->blk_max, (size_t)c->n_blk * sizeof(double)));
        CET_CUDA(cudaMalloc(&c->plane_sum, (size_t)(c->n0 + 2) * sizeof(double)));
    }
    return 0;
}

// Local plane ranges of a slab (see the header comment).
struct SlabRanges { int eval_lo, eval_hi, claim_lo, claim_hi, own_lo, own_hi; };
static SlabRanges slab_ranges(const cet_ctx *c)
{
    SlabRanges r;
    const int i_off = (int)(c->i_begin - c->halo), np = (int)c->np;
    const int dom_lo = i_off < 0 ? -i_off : 0;                         // first local plane inside the domain
    const int dom_hi = (i_off + np > c->n0) ? (int)(c->n0 - i_off) : np;
    const bool cut_lo = i_off > 0, cut_hi = i_off + np < c->n0;        // slab face with missing neighbours beyond
    r.eval_lo = cut_lo ? 2 : dom_lo;  r.eval_hi = cut_hi ? np - 2 : dom_hi;
    r.claim_lo = cut_lo ? 4 : dom_lo; r.claim_hi = cut_hi ? np - 4 : dom_hi;
    r.own_lo = c->halo; r.own_hi = np - c->halo;
    return r;
}

}  // namespace cet

using namespace cet;

extern "C" int cet_sweep_reset(cet_ctx *c)
{
    CET_REQUIRE(c, "cet_sweep_reset: NULL ctx");
    cet::DeviceGuard dg(c->device);
    if (c->sweep) CET_CUDA(cudaMemsetAsync(c->sweep, 0, sizeof(SweepState), c->stream));
    c->sweep_index = 0;
    return 0;
}

extern "C" int cet_sweep_run(cet_ctx *c, int64_t n_sweeps, const cet_sweep_params *sp,
                             const cet_thermal_params *tp, cet_sweep_result *res)
{
    CET_REQUIRE(c && sp && res, "cet_sweep_run: NULL argument");
    CET_REQUIRE(c->cubic && c->have_rp, "cet_sweep_run: needs a cubic context with rate params");
    CET_REQUIRE(sp->p_max > 0.0 && sp->p_max < 1.0 && sp->events_per_sweep > 0.0,
                "cet_sweep_run: need 0 < p_max < 1 and events_per_sweep > 0");
    CET_REQUIRE(sp->thermal_every <= 0 || tp != nullptr, "cet_sweep_run: thermal_every needs thermal params");
    CET_REQUIRE(c->world == 1 || c->halo >= 6, "cet_sweep_run: slabs need halo >= 6");
    CET_REQUIRE(c->world > 1 || (c->i_begin == 0 && c->i_end == c->n0),
                "cet_sweep_run: a partial slab needs cet_comm_init");
    cet::DeviceGuard dg(c->device);
    if (int rc = sweep_alloc(c)) return rc;
    const SlabRanges R = slab_ranges(c);
    const int bpp = (int)((c->n1 + SW_WARPS - 1) / SW_WARPS);
    const int n_eval = R.eval_hi - R.eval_lo;
    const int i_off = (int)(c->i_begin - c->halo);
    double *max_slot = c->plane_sum + c->n0;      // plane_sum[n0] holds the running max
    CET_REQUIRE(c->n1 <= 65535, "cet_sweep_run: L must fit 16-bit row indices");
    const size_t decide_smem = (size_t)SW_WARPS * 2 * c->n1 * sizeof(uint16_t);
    if (decide_smem > 48 * 1024)
        CET_CUDA(cudaFuncSetAttribute(sweep_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)decide_smem));

    SweepState before;
    CET_CUDA(cudaMemcpyAsync(&before, c->sweep, sizeof(before), cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));

    for (int64_t n = 0; n < n_sweeps; ++n) {
        ProfScope step_scope(c, PROF_STEP);
        if (sp->thermal_every > 0 && c->sweep_index % sp->thermal_every == 0) {
            if (int rc = thermal_cet_step(c, tp, &c->sweep->terminated)) return rc;
            if (c->world > 1) if (int rc = comm_halo_exchange(c, 4)) return rc;
        }
        sweep_reset_kernel<<<1, 1, 0, c->stream>>>(c->sweep, max_slot);
        CET_CUDA(cudaMemsetAsync(c->plane_sum, 0, (size_t)c->n0 * sizeof(double), c->stream));
        SweepArgs a;
        a.g = c->lat(); a.theta = c->theta; a.phi = c->phi; a.P = c->rp; a.ss = c->sweep;
        a.records = (Record *)c->records; a.cap_records = (unsigned int)c->cap_records;
        a.claim = c->claim; a.blk_sum = c->blk_sum; a.blk_max = c->blk_max;
        a.p_lo = R.eval_lo; a.p_hi = R.eval_hi; a.np = (int)c->np;
        a.seed = sp->seed; a.sweep = (uint32_t)c->sweep_index;
        a.rows_per_blk = SW_WARPS; a.blks_per_plane = bpp;
        {
            ProfScope ps(c, PROF_DECIDE);
            sweep_decide_kernel<<<n_eval * bpp, SW_WARPS * 32, decide_smem, c->stream>>>(a);
        }
        CET_CUDA(cudaGetLastError());
        sweep_plane_reduce_kernel<<<(n_eval + 3) / 4, 128, 0, c->stream>>>(
            c->blk_sum, c->blk_max, bpp, n_eval, i_off + R.eval_lo, (int)c->i_begin, (int)c->i_end,
            c->plane_sum, max_slot);
        CET_CUDA(cudaGetLastError());
        if (c->world > 1) if (int rc = comm_sweep_reduce(c, c->plane_sum, (int)c->n0, max_slot)) return rc;
        ApplyArgs b;
        b.vox = c->vox; b.theta = c->theta; b.phi = c->phi; b.vx = c->vx; b.vy = c->vy; b.vz = c->vz;
        b.ss = c->sweep;
        b.records = (const Record *)c->records; b.cap_records = (unsigned int)c->cap_records;
        b.claim = c->claim; b.P = c->rp; b.L = (int)c->n1; b.i_off = i_off;
        b.c_lo = R.claim_lo; b.c_hi = R.claim_hi; b.own_lo = R.own_lo; b.own_hi = R.own_hi;
        b.seed = sp->seed; b.sweep = (uint32_t)c->sweep_index; b.defect_fraction = sp->defect_fraction;
        {
            ProfScope ps(c, PROF_APPLY);
            sweep_apply_kernel<<<148 * 4, 256, 0, c->stream>>>(b);
        }
        CET_CUDA(cudaGetLastError());
        sweep_finalize_kernel<<<1, 256, 0, c->stream>>>(c->sweep, c->plane_sum, (int)c->n0, max_slot,
                                                        sp->events_per_sweep, sp->p_max);
        CET_CUDA(cudaGetLastError());
        if (c->world > 1) {
            ProfScope ps(c, PROF_HALO);
            if (int rc = comm_halo_exchange(c, 1 | 2)) return rc;
        }
        c->sweep_index++;
    }
    c->rates_valid = false;
    SweepState after;
    CET_CUDA(cudaMemcpyAsync(&after, c->sweep, sizeof(after), cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    res->sweeps_done = n_sweeps;
    res->events_fired = (int64_t)(after.n_fired - before.n_fired);
    res->events_applied = (int64_t)(after.n_applied - before.n_applied);
    res->nucleation_count = (int64_t)(after.n_nuc - before.n_nuc);
    res->sweep_index = c->sweep_index;
    res->time = after.time - before.time;
    res->last_total_rate = after.sum_rate; res->last_max_rate = after.max_rate; res->last_tau = after.tau;
    res->terminated = after.terminated;
    res->overflow = (int32_t)after.overflow;
    return 0;
}
