// poisson_onchip.cu -- the persistent on-chip Poisson kernel (see poisson_onchip.h for the design) and its launcher.
//
// Replaces src/poisson.c:224-285 (poisson_SOR_log: all sweeps, sum |u - u0| after each, stop at the first one below tol)
// for grids whose iterate fits the register files of the GPU.  Control flow is the lagged stop machine of
// poisson_stream.h (PoissonCtl, lag_fold / lag_action / lag_final) evaluated identically by every CTA.
#include "kernels.h"
#include "peer_device.cuh"
#include "poisson_onchip.h"

namespace cnv {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_acq_rel_gpu_add(unsigned long long *p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
// bounded spin until *p >= want (another CTA of this cooperative grid raises it); false on timeout or reported error
__device__ __forceinline__ bool oc_wait(const unsigned long long *p, unsigned long long want, const unsigned long long *err,
                                        unsigned long long timeout_ns)
{
    if (ld_acquire_gpu(p) >= want) return true;
    const unsigned long long t0 = peerdev::globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_gpu(p) < want) {
        if ((++spins & 255u) == 0 && (peerdev::globaltimer_ns() - t0 > timeout_ns || ld_acquire_gpu(err) != 0)) return false;
        __nanosleep(40);  // polls of one flag all go to the same L2 slice: keep them sparse
    }
    return true;
}

}  // namespace

// barrier among the compute warps only (the service warp does not take part in the sweeps)
__device__ __forceinline__ void bar_compute(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }


// The nsw sweeps of one pass.  A function of its own, instantiated per (colour of the tile origin, warp needs the update
// mask), so that the 64 + 64 registers of a patch meet exactly ONE pair of half-sweep bodies: with several bodies behind
// run-time selections in one loop the compiler has to shuffle the whole patch between their register assignments at every
// join (it did: 44 instructions per cell update instead of 15).  All instantiations execute the same barriers.
template <bool POW2, int PAR0, bool SEL>
__device__ __forceinline__ void oc_run_sweeps(OcThread &t, const RelaxConsts &rc, double *sm, const int nsw, double *s_acc, const int ncomp)
{
    for (int s = 0; s < nsw; s++) {
        double acc = 0.0;
        // a cell at distance d from the output region matters only while >= d half-sweeps remain after this one
        const int rem0 = 2 * (nsw - s) - 1;
        if (t.dist <= rem0) oc_half_sweep<POW2, PAR0, SEL>(t, rc, sm, acc);
        bar_compute(ncomp);
        if (t.dist <= rem0 - 1) oc_half_sweep<POW2, PAR0 ^ 1, SEL>(t, rc, sm, acc);
        // this thread's share of the sweep's norm (output threads only); reduced once per pass, off the sweeps' critical path
        s_acc[s * kOcMaxThreads + threadIdx.x] = t.own ? acc : 0.0;
        bar_compute(ncomp);
    }
}

// One CTA = the compute warps (one thread per patch) + ONE SERVICE WARP that owns everything which is latency rather than
// work.  While the compute warps run the sweeps of pass p it publishes the CTA's norm partials of pass p-1 and counts the CTA
// in; the CTA that counts in last sums all records in a fixed order and publishes the totals; every service warp then reads
// those 64 bytes and folds them into the action of pass p+1, so the stop decision never sits between two passes.  Between the passes it releases the CTA's band flag
// (a release is a fence: 1-2.5 k cycles during which the issuing warp stands still -- tools/ubench/fp64_lat.cu) and polls the
// neighbour tiles' flags, while the compute warps reduce their norms and store the interior.  The compute warps only ever wait
// at CTA barriers.  Critical path between two passes: band stores -> barrier -> release -> neighbours' flags -> barrier -> halo
// loads.
// flags: [ncta] band flags (passes whose band rows are released), [kOcNormSlots] norm counters (CTAs that have published their
// partials of the passes using that slot, cumulative), [kOcNormSlots] totals flags, error word -- one 128-byte line each;
// zeroed by the launcher.  partials[slot][ncta + 1][8]: per-sweep norm partials of a pass (slot = pass % kOcNormSlots) + totals.
// prof (optional): per CTA 8 counters of clock64 ticks (compute thread 0: slots 0-3, service lane 0: slots 4-7).
template <bool POW2>
__global__ void __launch_bounds__(kOcMaxThreads + 32, kOcOcc)
k_poisson_onchip(const OnchipGeom g, const RelaxConsts rc, double *__restrict__ buf0, double *__restrict__ buf1,
                 double *__restrict__ buf2, const double *__restrict__ rhs, PoissonCtl *ctl, unsigned long long *flags,
                 double *partials, double *hist, const unsigned long long timeout_ns, unsigned long long *prof)
{
    extern __shared__ double4 sm4[];
    double *sm = reinterpret_cast<double *>(sm4);
    __shared__ double s_acc[8 * kOcMaxThreads];       // [sweep of the pass][thread]: per-thread norm shares
    __shared__ double s_part[8][kOcMaxThreads / 32];  // [sweep of the pass][compute warp]
    __shared__ LagAction s_act;                       // what the upcoming pass does (service warp -> everybody)
    __shared__ PoissonCtl s_ctl;                      // chain state X_p (service lane 0)
    __shared__ int s_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    const int NT = oc_threads(g);
    const int ncw = (NT + 31) >> 5, ncomp = ncw * 32;  // compute warps / threads; warp ncw is the service warp
    const bool service = warp == ncw;
    const bool active = tid < NT;
    const bool first_cta = cta == 0;
    unsigned long long *bflags = flags, *ncount = flags + (size_t)ncta * kOcFlagStride;
    unsigned long long *err = flags + (size_t)(ncta + 2 * kOcNormSlots) * kOcFlagStride;
    double *bufs[3] = {buf0, buf1, buf2};
    long long tick = 0;
    unsigned long long pacc[4] = {0, 0, 0, 0};
    const bool timer = prof && lane == 0 && (tid == 0 || service);
    auto lap = [&](int slot) {
        if (timer) {
            const long long now = clock64();
            pacc[slot] += (unsigned long long)(now - tick);
            tick = now;
        }
    };

    for (int i = tid; i < kOcSlots * kOcPitch; i += blockDim.x) sm[i] = 0.0;
    if (tid == 0) { s_ctl = *ctl; s_bad = 0; }
    __syncthreads();
    if (tid == 0) s_act = lag_action(s_ctl, 0, g.T);
    const int cur0 = s_ctl.cur;

    if (!service) {
        // ================= compute warps: one thread per patch =================
        const OcTile tile = oc_tile(g, blockIdx.x, blockIdx.y);
        OcThread t;
        oc_thread_init(t, g, tile, active ? tid : 0, sm);
        if (!active) { t.upd = t.inr = 0; t.own = t.band = false; t.src = -1; t.fast = true; t.dist = 1 << 30; }
        // warp-uniform: some lane of this warp owns a cell that must not be updated (Dirichlet ring, outside the array)
        const bool warp_sel = __any_sync(0xffffffffu, !t.fast) != 0;
        if (active) {
            oc_load_rhs(t, rhs);
            oc_load_psi(t, bufs[cur0]);
            oc_publish_all(t, sm);
        }
        __syncthreads();  // (start) patches published, s_act of pass 0 written
        if (timer) tick = clock64();
        for (int p = 0;; p++) {
            const LagAction act = s_act;  // (written by the service warp before the barrier that ended the previous pass)
            if (act.kind == 0) break;
            double *__restrict__ out = bufs[act.out];
            const int nsw = act.nsw;
            if (act.kind == 2) {  // recompute the converged pass from its (intact) input with exactly act.nsw sweeps
                if (active) {
                    oc_load_psi(t, bufs[act.in]);
                    oc_publish_all(t, sm);
                }
                bar_compute(ncomp);
            }
            if (tile.par0 == 0) {
                if (warp_sel) oc_run_sweeps<POW2, 0, true>(t, rc, sm, nsw, s_acc, ncomp);
                else oc_run_sweeps<POW2, 0, false>(t, rc, sm, nsw, s_acc, ncomp);
            } else {
                if (warp_sel) oc_run_sweeps<POW2, 1, true>(t, rc, sm, nsw, s_acc, ncomp);
                else oc_run_sweeps<POW2, 1, false>(t, rc, sm, nsw, s_acc, ncomp);
            }
            if (t.band) oc_store(t, out);  // what the neighbours read as halo goes first
            lap(0);
            __syncthreads();  // (A) the band stores of this CTA have been issued
            lap(1);
            {   // per-warp norm sums of all sweeps of the pass: independent shuffle chains, fixed order
                double a[8];
#pragma unroll
                for (int s = 0; s < 8; s++) a[s] = s < nsw ? s_acc[s * kOcMaxThreads + tid] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int s = 0; s < 8; s++) a[s] = xadd(a[s], __shfl_xor_sync(0xffffffffu, a[s], o));
                if (lane < 8) {
                    double mine = a[0];
#pragma unroll
                    for (int s = 1; s < 8; s++) mine = lane == s ? a[s] : mine;
                    s_part[lane][warp] = mine;
                }
            }
            // the rest of the output region is only read back by this CTA itself (a "redo" pass two passes from now) or by the
            // host at the end: stored off the critical path
            if (active && t.own && !t.band) oc_store(t, out);
            lap(2);
            __syncthreads();  // (D) the next action is known, the neighbours' bands have landed
            if (act.kind != 2 && s_act.kind == 1) {
                if (active && !t.own && t.src >= 0) {
                    oc_load_psi(t, out);
                    oc_publish_all(t, sm);
                }
                bar_compute(ncomp);
            }
            lap(3);
        }
        if (timer)
            for (int i = 0; i < 4; i++) prof[(size_t)cta * 8 + i] = pacc[i];
        return;
    }

    // ================= service warp =================
    int nb[8];
    const int nnb = oc_neighbours(g, blockIdx.x, blockIdx.y, nb);
    // Norms of a pass.  Every CTA publishes its eight per-sweep partials (64-byte record) and counts itself in; the CTA that
    // counts in LAST sums all records in a fixed order (lane l takes CTAs l, l+32, ... in order, then a butterfly: the result does
    // not depend on who is last) and publishes the totals; everybody reads the totals.  (All CTAs reading all records -- 144 x 144
    // requests on 72 L2 lines at the same moment -- cost 12 k cycles per pass.)
    // partials: [slot][ncta + 1 records of 8 doubles], record ncta = the totals; ncount[slot] cumulative, tflag[slot] = pass + 1.
    unsigned long long *tflag = ncount + (size_t)kOcNormSlots * kOcFlagStride;
    auto publish_partials = [&](int q, int nsw_q) {
        double *rec = partials + (size_t)(q % kOcNormSlots) * (ncta + 1) * 8;
        if (lane < 8) {
            double e = 0.0;
            for (int w = 0; w < ncw; w++) e = xadd(e, s_part[lane][w]);
            rec[(size_t)cta * 8 + lane] = lane < nsw_q ? e : 0.0;
        }
        __syncwarp();
        unsigned long long old = 0;
        if (lane == 0) old = atom_acq_rel_gpu_add(&ncount[(size_t)(q % kOcNormSlots) * kOcFlagStride], 1ull);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old + 1 != (unsigned long long)ncta * (unsigned long long)(q / kOcNormSlots + 1)) return;
        // last one in: every record of the pass is visible (acquire on the counter)
        double e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = 0.0;
        for (int c0 = 0; c0 < ncta; c0 += 32 * 5) {  // (one round for up to 160 CTAs: all 20 loads of a lane in flight at once)
            double2 v[5][4];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int cc = c0 + 32 * k + lane;
                const double2 *src = reinterpret_cast<const double2 *>(rec + (size_t)(cc < ncta ? cc : 0) * 8);
#pragma unroll
                for (int i = 0; i < 4; i++) v[k][i] = __ldcg(src + i);  // L2: written by other SMs
            }
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (c0 + 32 * k + lane < ncta)
#pragma unroll
                    for (int i = 0; i < 4; i++) { e[2 * i] = xadd(e[2 * i], v[k][i].x); e[2 * i + 1] = xadd(e[2 * i + 1], v[k][i].y); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < 8; i++) e[i] = xadd(e[i], __shfl_xor_sync(0xffffffffu, e[i], o));
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) rec[(size_t)ncta * 8 + i] = e[i];
            st_release_gpu(&tflag[(size_t)(q % kOcNormSlots) * kOcFlagStride], (unsigned long long)(q + 1));
        }
    };
    // the global norms of pass q -> e[0..7] in every lane
    auto gather = [&](int q, double (&e)[8]) {
        if (lane == 0) {
            if (!oc_wait(&tflag[(size_t)(q % kOcNormSlots) * kOcFlagStride], (unsigned long long)(q + 1), err, timeout_ns)) s_bad = 1;
        }
        __syncwarp();
        lap(1);
        const double2 *src = reinterpret_cast<const double2 *>(partials + ((size_t)(q % kOcNormSlots) * (ncta + 1) + ncta) * 8);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 v = __ldcg(src + i);
            e[2 * i] = v.x; e[2 * i + 1] = v.y;
        }
    };

    PoissonCtl c = s_ctl;  // chain state X_p, in registers of the service warp (every lane evolves the same copy)
    __syncthreads();  // (start)
    if (timer) tick = clock64();
    int P = 0, nsw_prev = 0;
    for (int p = 0;; p++) {
        const LagAction act = lag_action(c, p, g.T);  // == s_act
        P = p;
        if (act.kind == 0) break;
        // ---- while the compute warps sweep: this CTA's partials of pass p-1 (s_part is complete since barrier D), then the
        // decision for pass p+1: X_{p+1} = lag_fold(X_p, norms of pass p-1) ----
        if (p >= 1) publish_partials(p - 1, nsw_prev);
        const bool need = p >= 1 && c.state == 0 && c.redo == 0;
        double e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = 0.0;
        lap(0);
        if (need) gather(p - 1, e);
        if (s_bad) {
            c.state = 3;
            if (lane == 0) st_release_gpu(err, 1ull);
        } else if (p + 1 >= 2) {
            lag_fold(c, e, g.T, first_cta && lane == 0 ? hist : nullptr);
        }
        const LagAction next = lag_action(c, p + 1, g.T);
        lap(2);
        __syncthreads();  // (A) the band stores of this CTA have been issued
        // release: the band rows (ordered before this store by barrier A) are visible GPU-wide to whoever acquires the flag
        if (lane == 0) {
            st_release_gpu(&bflags[(size_t)cta * kOcFlagStride], (unsigned long long)(p + 1));
            s_act = next;  // (nobody reads s_act between barriers A and D)
        }
        // halo cells of the next pass = the neighbour tiles' output band of this one
        if (act.kind != 2 && next.kind == 1 && lane >= 8 && lane < 8 + nnb) {
            if (!oc_wait(&bflags[(size_t)nb[lane - 8] * kOcFlagStride], (unsigned long long)(p + 1), err, timeout_ns)) s_bad = 1;
        }
        __syncwarp();
        nsw_prev = act.nsw;
        __syncthreads();  // (D)
        lap(3);
    }
    if (timer)
        for (int i = 0; i < 4; i++) prof[(size_t)cta * 8 + 4 + i] = pacc[i];

    // ---- host-visible state after P passes: X_P, completed with the norms of pass P-1 where they still count ----
    const bool last_counts = c.state != 3 && lag_final_needs_last(c, P);  // identical on every CTA
    if (last_counts) publish_partials(P - 1, nsw_prev);
    if (!first_cta) return;
    if (last_counts) {
        double e[8];
        gather(P - 1, e);
        if (s_bad) c.state = 3;
        else lag_final(c, P, e, g.T, lane == 0 ? hist : nullptr);
    }
    if (lane == 0) {
        c.ticket = 0;
        *ctl = c;
    }
}

template <bool POW2>
static void launch_onchip_t(const OnchipGeom &g, const RelaxConsts &rc, double *b0, double *b1, double *b2, const double *rhs,
                            PoissonCtl *ctl, unsigned long long *flags, double *partials, double *hist, cudaStream_t s,
                            unsigned long long *prof)
{
    const size_t smem = oc_smem_bytes();
    static size_t configured[kMaxDevices] = {};
    size_t &conf = configured[current_device_slot()];
    if (smem > conf) {
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_onchip<POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    const int ncta = g.ntx * g.nty;
    CNV_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(unsigned long long) * kOcFlagStride * (ncta + 2 * kOcNormSlots + 1), s));
    static const unsigned long long timeout_ns = []() {
        const char *e = std::getenv("CNV_ONCHIP_TIMEOUT_MS");
        return (unsigned long long)(e ? std::atoi(e) : 10000) * 1000000ull;
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.ntx, g.nty);
    cfg.blockDim = dim3(round_up(oc_threads(g), 32) + 32);  // compute warps + the service warp
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's flags
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_onchip<POW2>, g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, timeout_ns, prof));
}

void launch_onchip(const OnchipGeom &g, const RelaxConsts &rc, double *b0, double *b1, double *b2, const double *rhs, PoissonCtl *ctl,
                   unsigned long long *flags, double *partials, double *hist, cudaStream_t s, unsigned long long *prof)
{
    if (rc.pow2) launch_onchip_t<true>(g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, s, prof);
    else launch_onchip_t<false>(g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, s, prof);
}

}  // namespace cnv
