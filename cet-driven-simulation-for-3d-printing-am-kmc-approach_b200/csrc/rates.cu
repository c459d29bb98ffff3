// rates.cu — dense per-voxel event-rate evaluation (kmc_event_rates.py:43-176) and the sum
// hierarchy the BKL search walks, plus the legacy event-list export.
//
// What is stored per site is the SUM of the site's event rates in reference list order
// (diff events of an occupied site, nuc+att events of an empty one); individual events are
// re-enumerated on demand by `site_events` (site_rates.cuh).  On the top plane (global
// i == L-1) empty sites additionally own one deposition event each (kmc_event_rates.py:55-72)
// kept in dep_rate (NaN = no event).
//
// Hierarchy, all in the reference's canonical list order
//   plane i -> [dep segment | occupied-site segment | empty-site segment] -> row j -> site k
//   row_occ/row_emp[p*n1+j] : warp_row_sums over k
//   seg[3p+{0,1,2}]         : warp_strided_sum over j
//   total                   : block_sum over seg
//
// Dense kernel shape: one warp per (plane, j) row — see dense_pass.cuh.
#include "ctx.cuh"
#include "dense_pass.cuh"
#include "reduce.cuh"

namespace cet {

constexpr int RB_WARPS = 8;

struct RatesArgs {
    Lat g;
    cet_rate_params P;
    double *site_rate, *dep_rate, *row_occ, *row_emp, *row_dep;
    int32_t *row_depcnt;
    int p_lo, p_hi;   // local planes to evaluate
    const uint32_t *stamp;   // DIRTY mode: only sites with stamp == stamp_id are re-evaluated
    uint32_t stamp_id;
};

// dynamic shared memory per warp: rate_row[L] doubles, then the two uint16 index lists
__host__ __device__ inline size_t dense_smem_per_warp(int L, bool with_rates)
{
    const size_t lists = (((size_t)2 * L * sizeof(uint16_t)) + 15) & ~(size_t)15;
    return lists + (with_rates ? (size_t)L * sizeof(double) : 0);
}

__global__ void __launch_bounds__(RB_WARPS * 32) rates_rows_kernel(const RatesArgs a)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ NbOffsets nbt;
    const int L = a.g.L;
    nb_offsets_init(&nbt, L);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * RB_WARPS + wid;
    const int nrows = (a.p_hi - a.p_lo) * L;
    if (row >= nrows) return;
    const int p = a.p_lo + row / L, j = row % L;
    const int i = a.g.i_off + p;
    const int rbase = (p * L + j) * L;
    unsigned char *mine = dyn_smem + wid * dense_smem_per_warp(L, true);
    double *rate_row = (double *)mine;
    RowLists w;
    w.occ = (uint16_t *)(mine + (size_t)L * sizeof(double));
    w.emp = w.occ + L;
    const bool top = i == a.g.n0 - 1;
    if (top)
        for (int k = lane; k < L; k += 32) a.dep_rate[j * L + k] = NAN;    // empty sites overwrite below
    row_classify(a.g, a.P, rbase, w, rate_row, nullptr, 0u, [](int, int) {});
    row_occupied(a.g, a.P, &nbt, i, j, rbase, w, [&](int k, double sum, bool active) {
        if (active) rate_row[k] = sum;
    });
    row_empty(a.g, a.P, &nbt, i, j, rbase, w, [&](int k, double sum, bool has_dep, double dep, bool active) {
        if (active) {
            rate_row[k] = sum;
            if (has_dep) a.dep_rate[j * L + k] = dep;
        }
    });
    __syncwarp();
    for (int k = lane; k < L; k += 32) a.site_rate[rbase + k] = rate_row[k];
    double occ, emp;
    warp_row_sums(a.g.vox + rbase, rate_row, L, &occ, &emp);
    if (lane == 0) { a.row_occ[p * L + j] = occ; a.row_emp[p * L + j] = emp; }
    if (top) {
        double ds; int dc;
        warp_dep_row(a.dep_rate + j * L, L, &ds, &dc);
        if (lane == 0) { a.row_dep[j] = ds; a.row_depcnt[j] = dc; }
    }
}

// Neighbour-rate refresh (sweep.cu): the sites stamped in this sweep are re-evaluated with the same
// chunk arithmetic as the dense pass.  Two kernels:
//   scan  streams the stamp array (4 B/site), compacts the stamped sites of each class into two
//         global lists (one atomic per class per 8-row tile, so a list stays in lattice order and a
//         chunk of 32 entries touches a handful of adjacent rows), and settles the sites that own
//         no list entry (defects; occupied sites of the top plane lose their deposition event);
//   eval  evaluates the lists 32 sites per warp with every warp of the GPU busy.
// Results go straight to site_rate / dep_rate; the BKL row sums are not maintained.
struct DirtyArgs {
    Lat g;
    cet_rate_params P;
    double *site_rate, *dep_rate;
    uint64_t *nst;                    // cache entries of the refreshed sites are rewritten
    const uint8_t *stamp;
    uint32_t stamp_id;
    int p_lo, p_hi;
    int32_t *list_occ, *list_emp;     // capacity: one entry per local site
    unsigned int *n_occ, *n_emp;      // list lengths (device counters)
};

__global__ void __launch_bounds__(RB_WARPS * 32) dirty_scan_kernel(const __grid_constant__ DirtyArgs a)
{
    __shared__ int s_occ[RB_WARPS * 64], s_emp[RB_WARPS * 64];     // staging; spills are appended directly
    __shared__ unsigned int c_occ, c_emp, b_occ, b_emp;
    const int L = a.g.L;
    if (threadIdx.x == 0) { c_occ = 0; c_emp = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nrows = (a.p_hi - a.p_lo) * L;
    const int row = blockIdx.x * RB_WARPS + wid;
    if (row < nrows) {
        const int p = a.p_lo + row / L, j = row % L;
        const bool top = a.g.i_off + p == a.g.n0 - 1;
        const int rbase = (p * L + j) * L;
        const bool vec = (L % 16) == 0;                      // rows are then 16-byte aligned in the stamp array
        const uint8_t want = (uint8_t)a.stamp_id;
        for (int k0 = 0; k0 < L; k0 += 512) {
            const int kb = k0 + 16 * lane;                   // this lane's sixteen consecutive sites
            if (kb >= L) continue;
            uint32_t wv[4];
            if (vec) {
                const uint4 v4 = *reinterpret_cast<const uint4 *>(a.stamp + rbase + kb);
                wv[0] = v4.x; wv[1] = v4.y; wv[2] = v4.z; wv[3] = v4.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    wv[q] = 0;
                    for (int b = 0; b < 4; ++b)
                        if (kb + 4 * q + b < L) wv[q] |= (uint32_t)(a.stamp[rbase + kb + 4 * q + b] == want ? want : (uint8_t)~want) << (8 * b);
                        else wv[q] |= (uint32_t)(uint8_t)~want << (8 * b);
                }
            }
            // one bit per matching byte (exact zero-byte test on wv ^ rep), then visit the matches only
            const uint32_t rep = 0x01010101u * want;
            uint32_t match = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t x = wv[q] ^ rep;                                  // zero byte <=> match
                const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte
                match |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * q);
            }
            while (match) {
            const int e = __ffs(match) - 1;
            match &= match - 1;
            const int k = kb + e;
            const int st = vox_state(a.g.vox[rbase + k]);
            if (st != 0) {
                if (top) a.dep_rate[j * L + k] = NAN;                   // occupied: no deposition event
                if (st == a.P.defect_id) {                               // defect: no events; repair its cache word here
                    a.site_rate[rbase + k] = 0.0;
                    const unsigned inb = inbounds_mask(a.g.i_off + p, j, k, a.g.n0, L);
                    a.nst[rbase + k] = neighbour_states(a.g, rbase + k, inb);
                    continue;
                }
                const unsigned int q = atomicAdd(&c_occ, 1u);
                if (q < RB_WARPS * 64) s_occ[q] = rbase + k;
                else a.list_occ[atomicAdd(a.n_occ, 1u)] = rbase + k;
            } else {
                const unsigned int q = atomicAdd(&c_emp, 1u);
                if (q < RB_WARPS * 64) s_emp[q] = rbase + k;
                else a.list_emp[atomicAdd(a.n_emp, 1u)] = rbase + k;
            }
            }
        }
    }
    __syncthreads();
    const unsigned int no = min(c_occ, (unsigned)(RB_WARPS * 64)), ne = min(c_emp, (unsigned)(RB_WARPS * 64));
    if (threadIdx.x == 0) {
        b_occ = no ? atomicAdd(a.n_occ, no) : 0u;
        b_emp = ne ? atomicAdd(a.n_emp, ne) : 0u;
    }
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < no; q += RB_WARPS * 32) a.list_occ[b_occ + q] = s_occ[q];
    for (unsigned int q = threadIdx.x; q < ne; q += RB_WARPS * 32) a.list_emp[b_emp + q] = s_emp[q];
}

__global__ void __launch_bounds__(RB_WARPS * 32) dirty_eval_kernel(const __grid_constant__ DirtyArgs a)
{
    __shared__ NbOffsets nbt;
    const int L = a.g.L, LL = L * L;
    nb_offsets_init(&nbt, L);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int no = (int)*a.n_occ, ne = (int)*a.n_emp;
    // blocked distribution: a CTA owns a contiguous run of each list (lists are in lattice order, so
    // its warps work on adjacent rows and share neighbour rows through L1)
    const int per_o = ((no + 31) / 32 + gridDim.x - 1) / gridDim.x * 32, per_e = ((ne + 31) / 32 + gridDim.x - 1) / gridDim.x * 32;
    const int o_lo = blockIdx.x * per_o, o_hi = min(no, o_lo + per_o);
    const int e_lo = blockIdx.x * per_e, e_hi = min(ne, e_lo + per_e);
    for (int c0 = o_lo + wid * 32; c0 < o_hi; c0 += RB_WARPS * 32) {
        const bool active = c0 + lane < o_hi;
        const int s = a.list_occ[active ? c0 + lane : c0];
        const int p = s / LL, j = (s / L) % L, k = s % L;
        const double sum = occ_chunk<true>(a.g, a.P, &nbt, a.g.i_off + p, j, k, s, active, a.nst);
        if (active) a.site_rate[s] = sum;
    }
    for (int c0 = e_lo + wid * 32; c0 < e_hi; c0 += RB_WARPS * 32) {
        const bool active = c0 + lane < e_hi;
        const int s = a.list_emp[active ? c0 + lane : c0];
        const int p = s / LL, j = (s / L) % L, k = s % L;
        const int i = a.g.i_off + p;
        bool has_dep;
        double dep;
        const double sum = emp_chunk<true>(a.g, a.P, &nbt, i, j, k, s, active, &has_dep, &dep, a.nst);
        if (active) {
            a.site_rate[s] = sum;
            if (i == a.g.n0 - 1) a.dep_rate[j * L + k] = has_dep ? dep : NAN;
        }
    }
}

// One warp per local plane: plane-segment sums in list order.
__global__ void seg_kernel(const double *row_occ, const double *row_emp, const double *row_dep,
                           double *seg, int n1, int p_lo, int p_hi, int i_off, int n0)
{
    const int p = p_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= p_hi) return;
    const double so = warp_strided_sum(row_occ + (int64_t)p * n1, n1);
    const double se = warp_strided_sum(row_emp + (int64_t)p * n1, n1);
    const double sd = (i_off + p == n0 - 1) ? warp_strided_sum(row_dep, n1) : 0.0;
    if ((threadIdx.x & 31) == 0) { seg[3 * p + 0] = sd; seg[3 * p + 1] = so; seg[3 * p + 2] = se; }
}

__global__ void total_kernel(const double *seg, const int32_t *row_depcnt, double *total, int p_lo,
                             int p_hi, int n1, int has_top)
{
    __shared__ double sm[40];
    __shared__ long long smi[40];
    const double t = block_sum(seg + 3 * p_lo, 3 * (p_hi - p_lo), sm);
    const long long nd = has_top ? block_sum_i(row_depcnt, n1, smi) : 0;
    if (threadIdx.x == 0) { total[0] = t; ((long long *)total)[1] = nd; }
}

// Dense evaluation of local planes [p_lo, p_hi): site_rate, dep_rate and the row sums.
int rates_rows(cet_ctx *c, int p_lo, int p_hi)
{
    CET_REQUIRE(c->cubic, "rates: context was created with cet_create_shape (thermal only)");
    CET_REQUIRE(c->have_rp, "rates: cet_set_rate_params has not been called");
    if (p_hi <= p_lo) return 0;
    if (int rc = nst_ensure(c)) return rc;
    RatesArgs a;
    a.g = c->lat();
    a.P = c->rp;
    a.site_rate = c->site_rate; a.dep_rate = c->dep_rate;
    a.row_occ = c->row_occ; a.row_emp = c->row_emp; a.row_dep = c->row_dep; a.row_depcnt = c->row_depcnt;
    a.p_lo = p_lo; a.p_hi = p_hi; a.stamp = nullptr; a.stamp_id = 0;
    const int nrows = (a.p_hi - a.p_lo) * (int)c->n1;
    CET_REQUIRE(c->n1 <= 65535, "rates: L must fit 16-bit row indices");
    CET_REQUIRE(c->nloc < (1ll << 31), "rates: the local lattice must have fewer than 2^31 sites");
    const size_t smem = RB_WARPS * dense_smem_per_warp((int)c->n1, true);
    CET_REQUIRE(smem <= 220 * 1024, "rates: L=%lld needs %zu B of shared memory per CTA", (long long)c->n1, smem);
    if (smem > 40 * 1024)
        CET_CUDA(cudaFuncSetAttribute(rates_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope ps(c, PROF_RATES);
        rates_rows_kernel<<<(nrows + RB_WARPS - 1) / RB_WARPS, RB_WARPS * 32, smem, c->stream>>>(a);
    }
    CET_CUDA(cudaGetLastError());
    return 0;
}

// Re-evaluate, on local planes [p_lo, p_hi), the sites whose stamp equals stamp_id.
// lists: 2 * nloc int32 (occupied list, then empty list); counters: two device unsigned ints (zeroed by the caller).
int rates_rows_dirty(cet_ctx *c, int p_lo, int p_hi, const uint8_t *stamp, uint32_t stamp_id, int32_t *lists,
                     unsigned int *counters)
{
    if (p_hi <= p_lo) return 0;
    if (int rc = nst_ensure(c)) return rc;
    DirtyArgs a;
    a.g = c->lat();
    a.P = c->rp;
    a.site_rate = c->site_rate; a.dep_rate = c->dep_rate; a.nst = c->nst;
    a.stamp = stamp; a.stamp_id = stamp_id; a.p_lo = p_lo; a.p_hi = p_hi;
    a.list_occ = lists; a.list_emp = lists + c->nloc;
    a.n_occ = counters; a.n_emp = counters + 1;
    const int nrows = (p_hi - p_lo) * (int)c->n1;
    dirty_scan_kernel<<<(nrows + RB_WARPS - 1) / RB_WARPS, RB_WARPS * 32, 0, c->stream>>>(a);
    CET_CUDA(cudaGetLastError());
    dirty_eval_kernel<<<148 * 40, RB_WARPS * 32, 0, c->stream>>>(a);
    CET_CUDA(cudaGetLastError());
    return 0;
}

int rates_build(cet_ctx *c)
{
    const int p_lo = c->halo, p_hi = (int)(c->np - c->halo);
    if (int rc = rates_rows(c, p_lo, p_hi)) return rc;
    RatesArgs a;
    a.g = c->lat(); a.p_lo = p_lo; a.p_hi = p_hi;
    const int npl = a.p_hi - a.p_lo;
    seg_kernel<<<(npl + 3) / 4, 128, 0, c->stream>>>(c->row_occ, c->row_emp, c->row_dep, c->seg, (int)c->n1,
                                                     a.p_lo, a.p_hi, a.g.i_off, a.g.n0);
    CET_CUDA(cudaGetLastError());
    total_kernel<<<1, 256, 0, c->stream>>>(c->seg, c->row_depcnt, c->total, a.p_lo, a.p_hi, (int)c->n1,
                                           c->i_end == c->n0 ? 1 : 0);
    CET_CUDA(cudaGetLastError());
    c->rates_valid = true;
    if (c->halo == 0) c->sweep_rates_valid = true;
    return 0;
}

// ---------------------------------------------------------------------------------------
// Legacy event list (kmc_event_rates.py:162-176 return value) as SoA, reference order.
// ---------------------------------------------------------------------------------------
struct ExportArgs {
    Lat g;
    cet_rate_params P;
    int p_lo;
    long long *plane_counts;      // [3 * planes]: dep, occ, emp event counts
    const long long *plane_base;  // [3 * planes]: first list index of each segment
    const double *species;        // device copy of the species draws (may be NULL)
    long long n_species;
    long long cap;
    uint8_t *type; long long *pos; double *rate; long long *target; int32_t *atom;
};

__device__ __forceinline__ int block_excl_scan_i(int v, int *smem, int *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = lane < nw ? smem[lane] : 0, ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < nw) smem[lane] = ti - t;
        if (lane == 31) smem[32] = ti;
    }
    __syncthreads();
    *total = smem[32];
    return smem[w] + inc - v;
}

// One CTA per plane.  WRITE=false counts the events of the three segments; WRITE=true writes
// them at plane_base + rank-in-segment.
template <bool WRITE>
__global__ void __launch_bounds__(256) export_kernel(const ExportArgs a)
{
    __shared__ int sm[40];
    const int L = a.g.L;
    const int pl = blockIdx.x, p = a.p_lo + pl, i = a.g.i_off + p;
    const int64_t LL = (int64_t)L * L;
    const int nsite = L * L;
    for (int segm = 0; segm < 3; ++segm) {
        if (segm == 0 && i != a.g.n0 - 1) {
            if (!WRITE && threadIdx.x == 0) a.plane_counts[3 * pl] = 0;
            continue;
        }
        long long run = 0;
        for (int c0 = 0; c0 < nsite; c0 += blockDim.x) {
            const int q = c0 + threadIdx.x;
            const int j = q / L, k = q % L;
            int cnt = 0;
            double drate = 0.0;
            bool mine = false;
            if (q < nsite) {
                const int st = vox_state(a.g.vox[a.g.idx(i, j, k)]);
                if (segm == 0) {
                    if (st == 0 && dep_rate(a.P, a.g.T[a.g.idx(i, j, k)], &drate)) cnt = 1;
                } else {
                    mine = (segm == 1) ? (st != 0) : (st == 0);
                    if (mine) site_rate_sum(a.g, a.P, i, j, k, &cnt);
                }
            }
            int tot;
            const int off = block_excl_scan_i(cnt, sm, &tot);
            if (WRITE && cnt > 0) {
                long long e = a.plane_base[3 * pl + segm] + run + off;
                const long long self = (long long)i * LL + (long long)j * L + k;
                if (segm == 0) {
                    if (e < a.cap) {
                        const long long di = run + off;   // index into the species stream
                        int atom = a.P.states_w;
                        if (a.species && di < a.n_species) atom = dep_species(a.P, a.species[di]);
                        a.type[e] = CET_EV_DEP; a.pos[e] = self; a.rate[e] = drate; a.target[e] = -1;
                        a.atom[e] = atom;
                    }
                } else {
                    site_events(a.g, a.P, i, j, k, [&](int ty, int slot, double r, int atom) {
                        if (e < a.cap) {
                            a.type[e] = (uint8_t)ty; a.pos[e] = self; a.rate[e] = r;
                            a.target[e] = slot < 0 ? -1
                                                   : self + CET_NB_DI(slot) * LL + CET_NB_DJ(slot) * L + CET_NB_DK(slot);
                            a.atom[e] = atom;
                        }
                        ++e;
                    });
                }
            }
            run += tot;
        }
        if (!WRITE && threadIdx.x == 0) a.plane_counts[3 * pl + segm] = run;
    }
}

}  // namespace cet

using namespace cet;

extern "C" {

int cet_rates_build(cet_ctx *c)
{
    CET_REQUIRE(c, "cet_rates_build: NULL ctx");
    cet::DeviceGuard dg(c->device);
    return rates_build(c);
}

int cet_rates_total(cet_ctx *c, double *total, int64_t *n_dep)
{
    CET_REQUIRE(c, "cet_rates_total: NULL ctx");
    CET_REQUIRE(c->rates_valid, "cet_rates_total: rates are stale (call cet_rates_build)");
    cet::DeviceGuard dg(c->device);
    double h[2];
    CET_CUDA(cudaMemcpyAsync(h, c->total, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    if (total) *total = h[0];
    if (n_dep) memcpy(n_dep, &h[1], 8);
    return 0;
}

int cet_rates_download(cet_ctx *c, double *site_rate, double *dep_rate)
{
    CET_REQUIRE(c, "cet_rates_download: NULL ctx");
    CET_REQUIRE(c->rates_valid || c->sweep_rates_valid, "cet_rates_download: rates are stale (call cet_rates_build)");
    cet::DeviceGuard dg(c->device);
    if (site_rate)
        CET_CUDA(cudaMemcpyAsync(site_rate, c->site_rate + c->owned_offset(), (size_t)c->owned_sites() * 8,
                                 cudaMemcpyDeviceToHost, c->stream));
    if (dep_rate) {
        CET_REQUIRE(c->i_end == c->n0, "cet_rates_download: this slab does not own the top plane");
        CET_CUDA(cudaMemcpyAsync(dep_rate, c->dep_rate, (size_t)c->plane * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CET_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int export_counts(cet_ctx *c, ExportArgs &a, long long *h_counts, int npl)
{
    a.g = c->lat(); a.P = c->rp; a.p_lo = c->halo;
    if (int rc = ensure_stage(c, (size_t)npl * 6 * sizeof(long long))) return rc;
    a.plane_counts = (long long *)c->stage;
    a.plane_base = a.plane_counts + 3 * npl;
    export_kernel<false><<<npl, 256, 0, c->stream>>>(a);
    CET_CUDA(cudaGetLastError());
    CET_CUDA(cudaMemcpyAsync(h_counts, a.plane_counts, (size_t)npl * 3 * sizeof(long long),
                             cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int cet_events_count(cet_ctx *c, int64_t *n_events, int64_t *n_dep)
{
    CET_REQUIRE(c && c->cubic && c->have_rp, "cet_events_count: needs a cubic context with rate params");
    cet::DeviceGuard dg(c->device);
    const int npl = (int)(c->i_end - c->i_begin);
    long long *h = new long long[3 * npl];
    ExportArgs a;
    memset(&a, 0, sizeof(a));
    int rc = export_counts(c, a, h, npl);
    long long tot = 0, nd = 0;
    for (int q = 0; q < 3 * npl; ++q) { tot += h[q]; if (q % 3 == 0) nd += h[q]; }
    delete[] h;
    if (rc) return rc;
    if (n_events) *n_events = tot;
    if (n_dep) *n_dep = nd;
    return 0;
}

int cet_events_export(cet_ctx *c, const double *species_draws, int64_t n_draws, int64_t cap,
                      uint8_t *type, int64_t *pos, double *rate, int64_t *target, int32_t *atom,
                      int64_t *n_written)
{
    CET_REQUIRE(c && c->cubic && c->have_rp, "cet_events_export: needs a cubic context with rate params");
    CET_REQUIRE(cap >= 0, "cet_events_export: negative capacity");
    cet::DeviceGuard dg(c->device);
    const int npl = (int)(c->i_end - c->i_begin);
    long long *h = new long long[6 * npl];
    ExportArgs a;
    memset(&a, 0, sizeof(a));
    int rc = export_counts(c, a, h, npl);
    if (rc) { delete[] h; return rc; }
    long long run = 0;
    for (int q = 0; q < 3 * npl; ++q) { h[3 * npl + q] = run; run += h[q]; }
    const long long n = run < cap ? run : cap;
    if (n_written) *n_written = run;
    void *dbuf = nullptr, *dsp = nullptr;
    rc = 0;
    do {
        if (n == 0) break;
        if (cudaMemcpyAsync((void *)a.plane_base, h + 3 * npl, (size_t)npl * 3 * sizeof(long long),
                            cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = 1; break; }
        const size_t per = 1 + 8 + 8 + 8 + 4;
        if (cudaMalloc(&dbuf, (size_t)n * per + 64) != cudaSuccess) { rc = 1; break; }
        char *q = (char *)dbuf;
        a.pos = (long long *)q; q += (size_t)n * 8;
        a.rate = (double *)q; q += (size_t)n * 8;
        a.target = (long long *)q; q += (size_t)n * 8;
        a.atom = (int32_t *)q; q += (size_t)n * 4;
        a.type = (uint8_t *)q;
        a.cap = n;
        if (species_draws && n_draws > 0) {
            if (cudaMalloc(&dsp, (size_t)n_draws * 8) != cudaSuccess) { rc = 1; break; }
            if (cudaMemcpyAsync(dsp, species_draws, (size_t)n_draws * 8, cudaMemcpyHostToDevice, c->stream) !=
                cudaSuccess) { rc = 1; break; }
            a.species = (const double *)dsp; a.n_species = n_draws;
        }
        export_kernel<true><<<npl, 256, 0, c->stream>>>(a);
        if (cudaGetLastError() != cudaSuccess) { rc = 1; break; }
        cudaError_t e = cudaSuccess;
        if (pos) e = cudaMemcpyAsync(pos, a.pos, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (!e && rate) e = cudaMemcpyAsync(rate, a.rate, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (!e && target) e = cudaMemcpyAsync(target, a.target, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (!e && atom) e = cudaMemcpyAsync(atom, a.atom, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
        if (!e && type) e = cudaMemcpyAsync(type, a.type, (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        if (!e) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = 1; break; }
    } while (0);
    if (rc) set_error("cet_events_export: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
    if (dbuf) cudaFree(dbuf);
    if (dsp) cudaFree(dsp);
    delete[] h;
    return rc;
}

}  // extern "C"
