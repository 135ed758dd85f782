// poisson.cu -- streamfunction Poisson solve on the B200: temporally blocked red-black SOR.
//
// Replaces src/poisson.c:62-285 (poisson, poisson_SOR, poisson_log, poisson_SOR_log, error) of the
// reference.  The per-thread phases live in poisson_stream.h (shared with the CPU schedule
// checker); this file holds the kernel skeleton (cp.async pipeline, barriers, reductions, the
// device-side stopping logic) and the host-side solver object.
//
// Roofline: the un-blocked algorithm moves 24 B per interior cell per sweep (read psi, read f,
// write psi); one pass of this kernel applies T sweeps for the same 24 B (+ halo overlap), so the
// HBM bound is 24/T B per cell-update and the kernel becomes shared-memory-bandwidth bound
// (~56 B of LDS/STS per cell-update) for T >= 2.
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "nccl_dl.h"
#include "peer_device.cuh"

namespace cnv {

static size_t g_launches = 0;
size_t total_launches() { return g_launches; }
void count_launch(size_t n) { g_launches += n; }

int current_device_slot()
{
    int dev = 0;
    CNV_CUDA_CHECK(cudaGetDevice(&dev));
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Sums `acc` over the threads of each level group in a fixed order; result in s_e[g].
__device__ __forceinline__ void group_sums(double *sm, double acc, int g, int kk, int TPG, double *s_e)
{
    const int tid = threadIdx.x;
    __syncthreads();
    sm[tid] = acc;
    __syncthreads();
    double *red2 = sm + blockDim.x;
    if (kk < 32) {
        double s = 0.0;
        for (int m = kk; m < TPG; m += 32) s = xadd(s, sm[g * TPG + m]);
        red2[g * 32 + kk] = s;
    }
    __syncthreads();
    if (kk == 0) {
        double s = 0.0;
        const int lim = TPG < 32 ? TPG : 32;
        for (int m = 0; m < lim; m++) s = xadd(s, red2[g * 32 + m]);
        s_e[g] = s;
    }
    __syncthreads();
}

// peer-memory helpers (multi-GPU: system-scope flags, bounded spin-waits): peer_device.cuh
using peerdev::globaltimer_ns;
using peerdev::report_error;
using peerdev::st_release_sys;
using peerdev::wait_ge;

// Sum over the ranks (rank order: identical on every rank) of the norms pass `g` published in this rank's mailbox.
// Returns false if a rank's flag did not arrive in time.
__device__ bool gather_norms(const PeerLinks &L, unsigned long long g, double *e)
{
    PeerMailbox *mb = L.mail[L.rank];
    for (int r = 0; r < L.world; r++)
        if (!wait_ge(L, &mb->norm_flag[r], g + 1)) return false;  // pass g publishes the value g+1
    const int slot = (int)(g & (kNormSlots - 1));
    for (int i = 0; i < 8; i++) {
        double sum = 0.0;
        for (int r = 0; r < L.world; r++) sum = xadd(sum, *(volatile double *)&mb->norms[slot][r][i]);
        e[i] = sum;
    }
    return true;
}

// Host-visible state after `L.pidx` passes (peer path), derived by one thread.
// S_0 is what k_reset_ctl wrote; S_p = decide(S_{p-1}, norms of pass p-1); ctlbuf[p & 1] = S_p.
__device__ PoissonCtl peer_state(const PeerLinks &L, int T, double *hist)
{
    if (L.pidx == 0) return L.ctlbuf[0];
    PoissonCtl c = L.ctlbuf[(L.pidx - 1) & 1];
    double e[8];
    if (c.state != 0) return c;
    if (!gather_norms(L, L.gidx - 1, e)) {
        report_error(L);
        c.state = 3;
        return c;
    }
    decide(c, e, pass_sweeps(c, T), hist);
    return c;
}

template <int T, bool POW2, bool PEER>
__global__ void __launch_bounds__(pass_max_threads(T), pass_min_ctas(T))
k_poisson_pass(const PassGeom p, const RelaxConsts rc, double *__restrict__ buf0, double *__restrict__ buf1,
               const double *__restrict__ rhs, PoissonCtl *ctl, double *__restrict__ partials, double *hist,
               double *norms_out, const int fused_decide, const PeerLinks L)
{
    extern __shared__ double4 sm4[];
    double *sm = reinterpret_cast<double *>(sm4);
    __shared__ double s_e[8];
    __shared__ int s_last;

    // Programmatic dependent launch: passes are launched back to back with programmatic stream serialisation, so the
    // next pass may become resident while this one drains (its CTAs then sit in griddepcontrol.wait until this grid
    // has completed and its writes are visible) -- hides the launch latency between dependent passes.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    const bool first_cta = blockIdx.x == 0 && blockIdx.y == 0;
    __shared__ PoissonCtl s_ctl;
    unsigned long long *trace = nullptr;
    if (L.trace && L.pidx < L.trace_passes && tid == 0) {
        trace = L.trace + ((size_t)L.pidx * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x) * 6;
        trace[0] = globaltimer_ns();
    }
    __shared__ LagAction s_act;  // (peer path only)
    if (PEER) {
        // peer path: derive this pass' state from the previous state + every rank's published norms.  The whole CTA
        // cooperates -- one thread per rank waits for that rank's flag, 8 x world threads fetch the norms -- so the
        // start-up cost does not grow with the number of GPUs (64 dependent loads by one thread were ~10 us at 8 GPUs).
        // The norms are those of pass gidx-1: ~2 us of flag latency per CTA (profiles/scale_r2.md).
        __shared__ double s_nrm[kMaxRanks][8];
        __shared__ int s_bad;
        const PoissonCtl prev = L.ctlbuf[L.pidx == 0 ? 0 : (L.pidx - 1) & 1];
        const bool need = peer_needs_norms(prev, L.pidx);  // uniform
        if (tid == 0) s_bad = 0;
        __syncthreads();
        if (need) {
            PeerMailbox *mb = L.mail[L.rank];
            for (int r = tid; r < L.world; r += blockDim.x)
                if (!wait_ge(L, &mb->norm_flag[r], L.gidx)) s_bad = 1;  // pass g publishes the value g+1
            __syncthreads();
            const int slot = (int)((L.gidx - 1) & (kNormSlots - 1));
            for (int i = tid; i < 8 * L.world; i += blockDim.x) s_nrm[i >> 3][i & 7] = *(volatile double *)&mb->norms[slot][i >> 3][i & 7];
            __syncthreads();
        }
        if (tid == 0) {
            PoissonCtl c = prev;
            double e[8];
            for (int g = 0; g < 8; g++) {
                double sum = 0.0;
                if (need)
                    for (int r = 0; r < L.world; r++) sum = xadd(sum, s_nrm[r][g]);  // rank order: identical on every rank
                e[g] = sum;
            }
            if (need && s_bad) report_error(L);
            const LagAction a = peer_advance(c, e, need, s_bad != 0, T, first_cta ? hist : nullptr);
            s_ctl = c;
            s_act = a;
            if (first_cta && L.pidx > 0) L.ctlbuf[L.pidx & 1] = c;
            if (first_cta && a.kind == 0) {
                // nothing to do (solve finished, or the pass in flight reaches itmax): a no-op pass still advances
                // the cumulative counters of the protocol
                if (L.rank > 0) atomicAdd_system(&L.mail[L.rank - 1]->halo_count[1], L.push_low);
                if (L.rank < L.world - 1) atomicAdd_system(&L.mail[L.rank + 1]->halo_count[0], L.push_high);
                for (int r = 0; r < L.world; r++) st_release_sys(&L.mail[r]->norm_flag[L.rank], L.gidx + 1);
            }
        }
        __syncthreads();
    }
    const PoissonCtl c0 = PEER ? s_ctl : *ctl;
    if (trace) trace[1] = globaltimer_ns();
    if (PEER ? s_act.kind == 0 : c0.state != 0) return;  // solve already finished: later passes of a batch are no-ops
    const int nsw = PEER ? s_act.nsw : pass_sweeps(c0, T);
    const int cur = PEER ? s_act.in : c0.cur;
    const int nxt = PEER ? s_act.out : cur ^ 1;
    const double *__restrict__ in = cur ? buf1 : buf0;
    double *__restrict__ out = nxt ? buf1 : buf0;

    const CtaGeom G = cta_geom(p, blockIdx.x, blockIdx.y);
    if (PEER) {
        __shared__ int s_inputs_ok;
        if (tid == 0) {
            PeerMailbox *mb = L.mail[L.rank];
            const bool push_down = L.rank > 0 && G.y0 < p.own_lo + p.HY;
            const bool push_up = L.rank < L.world - 1 && G.y1 > p.own_hi - p.HY;
            bool ok = true;
            // first pass of a solve: the neighbour must have finished zeroing the buffer this CTA pushes into
            if (L.pidx == 0 && push_down) ok = ok && wait_ge(L, &mb->ready[0], L.epoch);
            if (L.pidx == 0 && push_up) ok = ok && wait_ge(L, &mb->ready[1], L.epoch);
            // this CTA streams halo rows -> the neighbour's pushes of the previous pass must have landed
            if (L.rank > 0 && G.ylo < p.own_lo) ok = ok && wait_ge(L, &mb->halo_count[0], L.gidx * L.need_low);
            if (L.rank < L.world - 1 && G.yhi >= p.own_hi) ok = ok && wait_ge(L, &mb->halo_count[1], L.gidx * L.need_high);
            if (!ok) report_error(L);
            s_inputs_ok = ok ? 1 : 0;
            if (trace) trace[2] = globaltimer_ns();
        }
        __syncthreads();
        // a wait failed (a peer stopped responding): neither stream stale halos nor store into buffers the neighbour may be
        // re-initialising.  The error flag is set in every rank's mailbox; the hosts abort at their next state read-back.
        if (!s_inputs_ok) return;
    }
    StreamThread<T> st;
    stream_init<T>(st, p, G, sm, in, rhs, tid, blockDim.x);
    stream_set_sweeps<T>(st, nsw);
    stream_prezero<T>(st, sm);
    const int TPG = p.WS / (2 * kPairs);
    const int kk = tid - st.g * TPG;

    stream_prologue<T>(st, sm);  // kPrefetch rows in flight
    for (int r = st.ybase; r <= st.rend; r += 4) {
        cp_async_wait<kPrefetch - 1 - kLand>();  // row r (+kLand) has landed (this thread's copies) ...
        __syncthreads();                 // ... and everybody's; previous step's updates are visible
        stream_step<T, POW2, 0>(st, rc, sm, out, r, nsw);
        cp_async_wait<kPrefetch - 1 - kLand>();
        __syncthreads();
        stream_step<T, POW2, 1>(st, rc, sm, out, r + 1, nsw);
        cp_async_wait<kPrefetch - 1 - kLand>();
        __syncthreads();
        stream_step<T, POW2, 2>(st, rc, sm, out, r + 2, nsw);
        cp_async_wait<kPrefetch - 1 - kLand>();
        __syncthreads();
        stream_step<T, POW2, 3>(st, rc, sm, out, r + 3, nsw);
    }
    cp_async_wait<0>();
    const double acc = st.acc;
    struct { int g; } t = {st.g};
    if (trace) trace[3] = globaltimer_ns();

    if (PEER) {
        // Slab boundary rows: copy what this CTA has just written (still L2 resident) into the neighbour GPU's halo
        // rows with coalesced 16-byte peer stores over NVLink -- outside the streaming loop, which stays identical to
        // the single-GPU kernel.  Only the CTAs of the two edge chunk rows do this; ONE thread then orders the CTA's
        // pushes system-wide (the CTA barrier makes them precede its fence: the pattern of a grid barrier's arrive) and
        // counts the push in the neighbour's mailbox, while the rest of the CTA goes on with the norms.  Round 2
        // measured 8 us per CTA (every CTA, 576 threads each) for the unconditional all-thread __threadfence_system()
        // this replaces (profiles/scale_r2.md).
        const bool push_down = L.rank > 0 && G.y0 < p.own_lo + p.HY;
        const bool push_up = L.rank < L.world - 1 && G.y1 > p.own_hi - p.HY;
        if (push_down || push_up) {  // uniform over the CTA
            __syncthreads();  // every write-back of this CTA is done and visible to the CTA
            const int c0 = G.gx0 + p.HX, c1 = (c0 + p.Wout < p.ld ? c0 + p.Wout : p.ld);  // this strip's output columns
            const int npair = (c1 - c0) >> 1;
            for (int side = 0; side < 2; side++) {
                if (side == 0 ? !push_down : !push_up) continue;
                const int ra = side == 0 ? (G.y0 > p.own_lo ? G.y0 : p.own_lo) : (G.y0 > p.own_hi - p.HY ? G.y0 : p.own_hi - p.HY);
                const int rb = side == 0 ? (G.y1 < p.own_lo + p.HY ? G.y1 : p.own_lo + p.HY) : (G.y1 < p.own_hi ? G.y1 : p.own_hi);
                double *peer = (side == 0 ? L.down_buf[nxt] + L.down_delta : L.up_buf[nxt] + L.up_delta);
                for (int idx = tid; idx < (rb - ra) * npair; idx += blockDim.x) {
                    const int rr = ra + idx / npair, cc = c0 + 2 * (idx % npair);
                    const size_t off = (size_t)rr * p.ld + cc;
                    const double2 v = __ldcg(reinterpret_cast<const double2 *>(out + off));
                    *reinterpret_cast<double2 *>(peer + off) = v;
                }
            }
            __syncthreads();  // all pushes of the CTA are issued ...
            if (tid == 0) {
                __threadfence_system();  // ... and ordered before the counts below, system-wide (cumulative through the barrier)
                if (push_down) atomicAdd_system(&L.mail[L.rank - 1]->halo_count[1], 1ull);
                if (push_up) atomicAdd_system(&L.mail[L.rank + 1]->halo_count[0], 1ull);
            }
        }
        if (trace) trace[4] = globaltimer_ns();
    }
    // per-CTA L1 update norms, one per sweep of the pass (level g <-> sweep g+1)
    group_sums(sm, acc, t.g, kk, TPG, s_e);
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid < T) {
        partials[(size_t)cta * T + tid] = s_e[tid];
        __threadfence();
    }
    __syncthreads();
    unsigned *ticket = PEER ? &L.mail[L.rank]->ticket : &ctl->ticket;
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)ncta - 1;
    __syncthreads();
    if (trace && !s_last) trace[5] = globaltimer_ns();
    if (!s_last) return;

    // last CTA: grid-wide sums in a fixed order, then the stopping decision (src/poisson.c:272-279)
    __threadfence();
    double part = 0.0;
    for (int c = kk; c < ncta; c += TPG) part = xadd(part, __ldcg(&partials[(size_t)c * T + t.g]));
    group_sums(sm, part, t.g, kk, TPG, s_e);
    if (PEER) {
        // publish this rank's norms of the pass in every rank's mailbox, then raise the flag everywhere: thread
        // (r, g) stores one norm into rank r's mailbox, then -- after the CTA barrier, through which the release below
        // is cumulative over those stores -- one thread per destination rank releases its flag
        const int slot = (int)(L.gidx & (kNormSlots - 1));
        for (int i = tid; i < 8 * L.world; i += blockDim.x) {
            const int r = i >> 3, g = i & 7;
            L.mail[r]->norms[slot][L.rank][g] = (g < T && g < nsw) ? s_e[g] : 0.0;
        }
        __syncthreads();
        for (int r = tid; r < L.world; r += blockDim.x) st_release_sys(&L.mail[r]->norm_flag[L.rank], L.gidx + 1);
        if (tid == 0) {
            *ticket = 0;
            if (trace) trace[5] = globaltimer_ns();
        }
        return;
    }
    if (tid == 0) {
        if (fused_decide) {
            PoissonCtl c = *ctl;
            decide(c, s_e, nsw, hist);
            c.ticket = 0;
            *ctl = c;
        } else {
            for (int g = 0; g < T; g++) norms_out[g] = g < nsw ? s_e[g] : 0.0;
            ctl->ticket = 0;
        }
    }
}

// peer path: state after `L.pidx` passes -> ctlbuf[pidx & 1] (what the host reads back); also reports timeouts
__global__ void k_peer_finalize(const PeerLinks L, int T, double *hist)
{
    PoissonCtl c;
    if (*(volatile unsigned long long *)&L.mail[L.rank]->error) {  // a wait failed somewhere: do not wait for norms that never come
        c = L.ctlbuf[0];
        c.state = 3;
    } else {
        c = peer_state(L, T, hist);
        if (*(volatile unsigned long long *)&L.mail[L.rank]->error) c.state = 3;
    }
    L.ctlbuf[L.pidx & 1] = c;
}
// peer path: this rank's iterate buffers are (re-)initialised for solve epoch L.epoch: tell both neighbours
__global__ void k_peer_ready(const PeerLinks L)
{
    __threadfence_system();
    if (L.rank > 0) st_release_sys(&L.mail[L.rank - 1]->ready[1], L.epoch);          // I am its upper neighbour
    if (L.rank < L.world - 1) st_release_sys(&L.mail[L.rank + 1]->ready[0], L.epoch);  // I am its lower neighbour
}
// peer path: all pushes of the passes launched so far have landed in this rank's halos (called before the
// iterate is re-initialised for the next solve, so that a late push cannot overwrite the new initial guess)
__global__ void k_peer_quiesce(const PeerLinks L)
{
    PeerMailbox *mb = L.mail[L.rank];
    bool ok = true;
    if (L.rank > 0) ok = ok && wait_ge(L, &mb->halo_count[0], L.gidx * L.need_low);
    if (L.rank < L.world - 1) ok = ok && wait_ge(L, &mb->halo_count[1], L.gidx * L.need_high);
    if (!ok) report_error(L);
}

// multi-GPU: stopping decision from all-reduced norms (identical on every rank)
__global__ void k_decide(PoissonCtl *ctl, const double *norms, int T, double *hist)
{
    if (ctl->state != 0) return;
    PoissonCtl c = *ctl;
    decide(c, norms, pass_sweeps(c, T), hist);
    *ctl = c;
}

// multi-GPU, native NCCL path: norms of all ranks were gathered into gather[world][8] (own row included);
// summing them in rank order gives every rank bit-identical totals, hence the same decision
__global__ void k_decide_gather(PoissonCtl *ctl, const double *gather, int world, int T, double *hist)
{
    if (ctl->state != 0) return;
    PoissonCtl c = *ctl;
    double e[8];
    for (int g = 0; g < 8; g++) {
        double s = 0.0;
        for (int r = 0; r < world; r++) s = xadd(s, gather[r * 8 + g]);
        e[g] = s;
    }
    decide(c, e, pass_sweeps(c, T), hist);
    *ctl = c;
}

__global__ void k_reset_ctl(PoissonCtl *ctl, int itmax, double tol, int nbuf)
{
    PoissonCtl c;
    c.state = itmax > 0 ? 0 : 2;
    c.cur = 0; c.sweeps = 0; c.redo = 0; c.itmax = itmax; c.result_k = -1; c.ticket = 0; c.passes = 0;
    c.tol = tol; c.result_e = 0.0; c.last_e = 0.0; c.hit_e = 0.0; c.nbuf = nbuf; c.pad_ = 0;
    *ctl = c;
}

template <int T, bool POW2>
static void launch_pass(const PassGeom &g, const RelaxConsts &rc, double *b0, double *b1, const double *rhs, PoissonCtl *ctl,
                        double *partials, double *hist, double *norms, int fused, int threads, size_t smem, cudaStream_t s,
                        const PeerLinks &L)
{
    // opt in to large dynamic shared memory (static smem counts against the 227 KB).  Function attributes belong to the
    // device context, so the cache is per device (cnv_set_device may switch between solvers of one process).
    static size_t configured[kMaxDevices] = {};
    size_t &conf = configured[current_device_slot()];
    if (smem > 48 * 1024 && smem > conf) {
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_pass<T, POW2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_pass<T, POW2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.nstrips, g.nchunks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = !(std::getenv("CNV_POISSON_PDL") && std::atoi(std::getenv("CNV_POISSON_PDL")) == 0);
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const double *crhs = rhs;
    if (L.enabled)
        CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_pass<T, POW2, true>, g, rc, b0, b1, crhs, ctl, partials, hist, norms, fused, L));
    else
        CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_pass<T, POW2, false>, g, rc, b0, b1, crhs, ctl, partials, hist, norms, fused, L));
}

// ---------------------------------------------------------------------------------------------
constexpr int kOnchipDefault = 1;  // the on-chip kernel wherever it has a plan: measured faster than the streaming kernel on every
                                   // grid it fits (1.8 vs 3.3 us per sweep at 64^2 ... 2.7 vs 6.1 at 1024^2, profiles/onchip_r2.md)

static int env_int(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

PoissonSolver::PoissonSolver(int nrows, int ncols, int T, int grow0, int gnrows, int own_lo, int own_hi)
{
    if (gnrows < 0) gnrows = nrows;
    if (own_hi < 0) own_hi = nrows;
    if (T <= 0) T = env_int("CNV_POISSON_T", 8);  // deepest blocking measured best at every size (profiles/)
    if (T != 1 && T != 2 && T != 4 && T != 6 && T != 8) {
        std::printf("** Error: temporal block depth must be 1, 2, 4, 6 or 8 **\n");
        std::exit(1);
    }
    if (nrows < 3 || ncols < 3) {
        std::printf("** Error: invalid parameter **\n");  // src/linearalg.c:58-62 wording
        std::exit(1);
    }
    T_ = T;
    PlanLimits lim;
    int dev = 0;
    cudaDeviceProp prop;
    CNV_CUDA_CHECK(cudaGetDevice(&dev));
    CNV_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    lim.num_sms = prop.multiProcessorCount;
    lim.smem_per_cta = prop.sharedMemPerBlockOptin - 2048;  // room for the kernel's static shared memory (<= 1.7 KB, ptxas -v)
    smem_optin_ = lim.smem_per_cta;
    lim.smem_per_sm = prop.sharedMemPerMultiprocessor;
    lim.max_threads_per_sm = prop.maxThreadsPerMultiProcessor;
    const int ld = round_up(ncols, 16);
    geom_ = make_plan(nrows, ncols, ld, grow0, gnrows, own_lo, own_hi, T, lim, env_int("CNV_POISSON_WS", 0),
                      env_int("CNV_POISSON_CHUNKS", 0));
    if (geom_.WS == 0) {
        std::printf("** Error: no launch plan for a %dx%d grid **\n", nrows, ncols);
        std::exit(1);
    }
    threads_ = pass_threads(T, geom_.WS);
    smem_ = pass_smem_bytes(T, geom_.WS);
    // CNV_POISSON_ONCHIP: 1 = the persistent on-chip kernel whenever a plan exists (whole-domain solvers only), 0 = never
    const int oc_mode = env_int("CNV_POISSON_ONCHIP", kOnchipDefault);
    if (oc_mode != 0 && grow0 == 0 && gnrows == nrows && own_lo == 0 && own_hi == nrows)
        use_onchip_ = onchip_plan(nrows, ncols, ld, grow0, gnrows, own_lo, own_hi, lim.num_sms, &oc_, env_int("CNV_ONCHIP_T", 0),
                                  env_int("CNV_ONCHIP_NTX", 0), env_int("CNV_ONCHIP_NTY", 0));
    std::memset(&rc_, 0, sizeof rc_);
    const size_t bytes = (size_t)nrows * ld * sizeof(double);
    if (use_onchip_) {  // three iterate buffers rotate (lagged stop decision), per-CTA flags, per-pass norm partials
        CNV_CUDA_CHECK(cudaMalloc(&buf_[2], bytes));
        CNV_CUDA_CHECK(cudaMemset(buf_[2], 0, bytes));
        const size_t ncta = (size_t)oc_.ntx * oc_.nty;
        CNV_CUDA_CHECK(cudaMalloc(&oc_flags_, sizeof(unsigned long long) * kOcFlagStride * (ncta + 2 * kOcNormSlots + 1)));
        CNV_CUDA_CHECK(cudaMalloc(&oc_partials_, sizeof(double) * kOcNormSlots * (ncta + 1) * 8));
        CNV_CUDA_CHECK(cudaMemset(oc_partials_, 0, sizeof(double) * kOcNormSlots * (ncta + 1) * 8));
        if (env_int("CNV_ONCHIP_PROF", 0)) {
            CNV_CUDA_CHECK(cudaMalloc(&oc_prof_, sizeof(unsigned long long) * ncta * 8));
            CNV_CUDA_CHECK(cudaMemset(oc_prof_, 0, sizeof(unsigned long long) * ncta * 8));
        }
    }
    for (int i = 0; i < 2; i++) {
        CNV_CUDA_CHECK(cudaMalloc(&buf_[i], bytes));
        CNV_CUDA_CHECK(cudaMemset(buf_[i], 0, bytes));
    }
    CNV_CUDA_CHECK(cudaMalloc(&rhs_, bytes));
    CNV_CUDA_CHECK(cudaMemset(rhs_, 0, bytes));
    const size_t npart = (size_t)geom_.nstrips * geom_.nchunks * T;
    CNV_CUDA_CHECK(cudaMalloc(&partials_, sizeof(double) * npart));
    CNV_CUDA_CHECK(cudaMalloc(&norms_, sizeof(double) * 8));
    CNV_CUDA_CHECK(cudaMemset(norms_, 0, sizeof(double) * 8));
    CNV_CUDA_CHECK(cudaMalloc(&ctl_, sizeof(PoissonCtl)));
    CNV_CUDA_CHECK(cudaMemset(ctl_, 0, sizeof(PoissonCtl)));
    CNV_CUDA_CHECK(cudaMallocHost(&h_ctl_, sizeof(PoissonCtl)));
    CNV_CUDA_CHECK(cudaEventCreateWithFlags(&ev_, cudaEventDisableTiming));
}

PoissonSolver::~PoissonSolver()
{
    peer_close();  // (a caller that shares buffers with peers quiesces + synchronises all ranks first, see peer_close)
    cudaFree(buf_[0]); cudaFree(buf_[1]); if (buf_[2]) cudaFree(buf_[2]); cudaFree(rhs_); cudaFree(partials_); cudaFree(norms_); cudaFree(ctl_);
    if (hist_) cudaFree(hist_);
    if (gather_) cudaFree(gather_);
    if (mailbox_) cudaFree(mailbox_);
    if (trace_) cudaFree(trace_);
    if (ctlbuf_) cudaFree(ctlbuf_);
    if (oc_flags_) cudaFree(oc_flags_);
    if (oc_partials_) cudaFree(oc_partials_);
    if (oc_prof_) cudaFree(oc_prof_);
    cudaFreeHost(h_ctl_);
    cudaEventDestroy(ev_);
}

int PoissonSolver::onchip_profile(unsigned long long *out, int max_ctas)
{
    if (!oc_prof_) return 0;
    const int ncta = std::min(max_ctas, oc_.ntx * oc_.nty);
    CNV_CUDA_CHECK(cudaDeviceSynchronize());
    CNV_CUDA_CHECK(cudaMemcpy(out, oc_prof_, sizeof(unsigned long long) * 8 * (size_t)ncta, cudaMemcpyDeviceToHost));
    return ncta;
}

void PoissonSolver::set_consts(double dx, double dy, double beta) { rc_ = make_relax_consts(dx, dy, beta); }

void PoissonSolver::enable_history(int cap)
{
    if (cap <= 0) { use_hist_ = false; return; }
    if (hist_cap_ < cap) {
        if (hist_) cudaFree(hist_);
        CNV_CUDA_CHECK(cudaMalloc(&hist_, sizeof(double) * (size_t)cap));
        hist_cap_ = cap;
    }
    CNV_CUDA_CHECK(cudaMemset(hist_, 0, sizeof(double) * (size_t)hist_cap_));
    use_hist_ = true;
}

void PoissonSolver::reset_ctl(int itmax, double tol, cudaStream_t s)
{
    dist_passes_ = 0;
    k_reset_ctl<<<1, 1, 0, s>>>(links_.enabled ? links_.ctlbuf : ctl_, itmax, tol, 2);
    count_launch(1);
}

PoissonCtl PoissonSolver::read_ctl(cudaStream_t s)
{
    const PoissonCtl *src = ctl_;
    if (links_.enabled) {  // the state after the passes enqueued so far is derived by a one-thread kernel
        PeerLinks L = links_;
        L.pidx = dist_passes_;
        L.gidx = peer_gidx_;
        k_peer_finalize<<<1, 1, 0, s>>>(L, T_, use_hist_ ? hist_ : nullptr);
        count_launch(1);
        src = links_.ctlbuf + (dist_passes_ & 1);
    }
    CNV_CUDA_CHECK(cudaMemcpyAsync(h_ctl_, src, sizeof(PoissonCtl), cudaMemcpyDeviceToHost, s));
    CNV_CUDA_CHECK(cudaEventRecord(ev_, s));
    CNV_CUDA_CHECK(cudaEventSynchronize(ev_));
    if (h_ctl_->state == 3 && !links_.enabled) {
        std::printf("** Error: the on-chip Poisson kernel timed out waiting for a neighbour tile (CNV_ONCHIP_TIMEOUT_MS) **\n");
        std::fflush(stdout);
        std::exit(1);
    }
    if (h_ctl_->state == 3) {
        std::printf("** Error: multi-GPU peer exchange timed out (a rank stopped responding for more than %.0f s; "
                    "CNV_PEER_TIMEOUT_MS) **\n", links_.timeout_ns * 1e-9);
        std::fflush(stdout);
        std::exit(1);
    }
    return *h_ctl_;
}

// ---- peer-memory (CUDA IPC) setup ----------------------------------------------------------------
void PoissonSolver::peer_export(unsigned char *out256)
{
    if (!mailbox_) {
        CNV_CUDA_CHECK(cudaMalloc(&mailbox_, sizeof(PeerMailbox)));
        CNV_CUDA_CHECK(cudaMemset(mailbox_, 0, sizeof(PeerMailbox)));
        CNV_CUDA_CHECK(cudaMalloc(&ctlbuf_, 2 * sizeof(PoissonCtl)));
        CNV_CUDA_CHECK(cudaMemset(ctlbuf_, 0, 2 * sizeof(PoissonCtl)));
    }
    cudaIpcMemHandle_t h[4];  // (the fourth slot of the 256-byte record is reserved, zero)
    std::memset(h, 0, sizeof h);
    CNV_CUDA_CHECK(cudaIpcGetMemHandle(&h[0], buf_[0]));
    CNV_CUDA_CHECK(cudaIpcGetMemHandle(&h[1], buf_[1]));
    CNV_CUDA_CHECK(cudaIpcGetMemHandle(&h[2], mailbox_));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    std::memcpy(out256, h, 256);
}

// pushes per pass this slab makes downwards / upwards (CTAs whose output rows reach into the boundary band)
void PoissonSolver::peer_push_counts(int rank, int world, long long *low, long long *high) const
{
    long long lo = 0, hi = 0;
    for (int by = 0; by < geom_.nchunks; by++) {
        const CtaGeom G = cta_geom(geom_, 0, by);
        if (rank > 0 && G.y0 < geom_.own_lo + geom_.HY) lo += geom_.nstrips;
        if (rank < world - 1 && G.y1 > geom_.own_hi - geom_.HY) hi += geom_.nstrips;
    }
    *low = lo; *high = hi;
}

// handles: world x 256 bytes (peer_export of every rank: buffer 0, buffer 1, mailbox, 64 reserved bytes); layout: world x 4 ints (own_lo, own_hi, push_low, push_high)
int PoissonSolver::peer_import(int rank, int world, const unsigned char *handles, const int *layout)
{
    if (world > kMaxRanks || !mailbox_) return 1;
    std::memset(&links_, 0, sizeof links_);
    links_.rank = rank; links_.world = world; links_.ctlbuf = ctlbuf_;
    const int nbuf = 2;
    // every mapping opened here is remembered, so that the error paths below and peer_close() release it again
    auto open = [&](const unsigned char *h64, void **out) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, h64, 64);
        if (cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
        imported_.push_back(*out);
        return true;
    };
    auto fail = [&](int code) {
        cudaGetLastError();
        release_imports();
        std::memset(&links_, 0, sizeof links_);
        return code;
    };
    for (int r = 0; r < world; r++) {
        if (r == rank) { links_.mail[r] = mailbox_; continue; }
        void *ptr = nullptr;
        if (!open(handles + 256 * r + 128, &ptr)) return fail(2);
        links_.mail[r] = (PeerMailbox *)ptr;
    }
    const int *me = layout + 4 * rank;
    if (rank > 0) {
        const int *nb = layout + 4 * (rank - 1);
        for (int b = 0; b < nbuf; b++) {
            void *ptr = nullptr;
            if (!open(handles + 256 * (rank - 1) + 64 * b, &ptr)) return fail(3);
            links_.down_buf[b] = (double *)ptr;
        }
        links_.down_delta = (long long)(nb[1] - me[0]) * geom_.ld;  // my row own_lo + i -> its row own_hi' + i
        links_.need_low = (unsigned long long)nb[3];                // its upward pushes land in my low halo
    }
    if (rank < world - 1) {
        const int *nb = layout + 4 * (rank + 1);
        for (int b = 0; b < nbuf; b++) {
            void *ptr = nullptr;
            if (!open(handles + 256 * (rank + 1) + 64 * b, &ptr)) return fail(4);
            links_.up_buf[b] = (double *)ptr;
        }
        links_.up_delta = (long long)(nb[0] - me[1]) * geom_.ld;    // my row own_hi - HY + i -> its row own_lo'' - HY + i
        links_.need_high = (unsigned long long)nb[2];
    }
    links_.push_low = (unsigned long long)me[2];
    links_.push_high = (unsigned long long)me[3];
    // spin-wait bound: a rank whose host is late (writing output, paused) is not dead; a dead one must not hang the GPU
    links_.timeout_ns = (unsigned long long)std::max(1, env_int("CNV_PEER_TIMEOUT_MS", 300000)) * 1000000ull;
    links_.enabled = 1;
    distributed_ = true;
    peer_gidx_ = 0;
    return 0;
}

void PoissonSolver::release_imports()
{
    for (void *p : imported_) cudaIpcCloseMemHandle(p);
    imported_.clear();
    cudaGetLastError();
}

// End of the peer path for this solver.  Protocol for the caller (SlabPoisson.close, the multi-GPU driver): every rank calls
// peer_quiesce + a device synchronisation (all pushes INTO this rank have landed, all passes OF this rank -- trailing no-op
// passes included, which still write flags into every mailbox -- have drained), then a barrier of all ranks, then
// peer_close(); only after that may any rank free the buffers its peers had mapped.
void PoissonSolver::peer_close()
{
    if (!links_.enabled && imported_.empty()) return;
    cudaDeviceSynchronize();
    release_imports();
    std::memset(&links_, 0, sizeof links_);
}

int PoissonSolver::peer_trace_enable(int passes)
{
    const int ctas = geom_.nstrips * geom_.nchunks;
    if (trace_) cudaFree(trace_);
    const size_t bytes = sizeof(unsigned long long) * 6 * (size_t)ctas * passes;
    CNV_CUDA_CHECK(cudaMalloc(&trace_, bytes));
    CNV_CUDA_CHECK(cudaMemset(trace_, 0, bytes));
    links_.trace = trace_;
    links_.trace_passes = passes;
    trace_pass_ = 0;
    return ctas;
}

void PoissonSolver::peer_trace_read(unsigned long long *out, size_t n)
{
    CNV_CUDA_CHECK(cudaDeviceSynchronize());
    CNV_CUDA_CHECK(cudaMemcpy(out, trace_, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
}

void PoissonSolver::peer_quiesce(cudaStream_t s)
{
    if (!links_.enabled) return;
    PeerLinks L = links_;
    L.gidx = peer_gidx_;
    k_peer_quiesce<<<1, 1, 0, s>>>(L);
    count_launch(1);
}

void PoissonSolver::peer_ready(cudaStream_t s)
{
    if (!links_.enabled) return;
    PeerLinks L = links_;
    L.epoch = ++peer_epoch_;
    k_peer_ready<<<1, 1, 0, s>>>(L);
    count_launch(1);
}

void PoissonSolver::enqueue_passes(int npasses, cudaStream_t s)
{
    double *hist = use_hist_ ? hist_ : nullptr;
    const int fused = distributed_ ? 0 : 1;
    for (int i = 0; i < npasses; i++) {
        PeerLinks L = links_;
        if (L.enabled) {
            L.pidx = dist_passes_++;
            L.gidx = peer_gidx_++;
            L.epoch = peer_epoch_;
        } else if (L.trace) {
            L.pidx = trace_pass_++;  // diagnostics only (tools/peer_trace.py --single)
        }
#define CNV_PASS(TT)                                                                                                     \
    if (T_ == TT) {                                                                                                      \
        if (rc_.pow2)                                                                                                    \
            launch_pass<TT, true>(geom_, rc_, buf_[0], buf_[1], rhs_, ctl_, partials_, hist, norms_, fused, threads_, smem_, s, L); \
        else                                                                                                             \
            launch_pass<TT, false>(geom_, rc_, buf_[0], buf_[1], rhs_, ctl_, partials_, hist, norms_, fused, threads_, smem_, s, L); \
    }
        CNV_PASS(1) CNV_PASS(2) CNV_PASS(4) CNV_PASS(6) CNV_PASS(8)
#undef CNV_PASS
    }
    CNV_CUDA_CHECK(cudaGetLastError());
    launches_ += npasses;
    count_launch(npasses);
}

#define CNV_NCCL_CHECK(expr)                                                                              \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            std::printf("** Error: NCCL failure %s at %s:%d **\n", nccl_api().GetErrorString ? nccl_api().GetErrorString(r__) : "?", \
                        __FILE__, __LINE__);                                                              \
            std::fflush(stdout);                                                                          \
            std::exit(1);                                                                                 \
        }                                                                                                 \
    } while (0)

void PoissonSolver::attach_comm(const SlabComm &c)
{
    comm_ = c;
    if (!gather_) {
        CNV_CUDA_CHECK(cudaMalloc(&gather_, sizeof(double) * 8 * (size_t)c.world));
        CNV_CUDA_CHECK(cudaMemset(gather_, 0, sizeof(double) * 8 * (size_t)c.world));
    }
    distributed_ = true;
}

// send/recv of `depth` boundary rows with both slab neighbours; must be called inside an NCCL group
static void halo_ops(const NcclApi &n, const SlabComm &c, const PassGeom &g, double *f, int depth, cudaStream_t s)
{
    ncclComm_t comm = (ncclComm_t)c.comm;
    const size_t cnt = (size_t)depth * g.ld;
    if (c.rank > 0) {
        CNV_NCCL_CHECK(n.Send(f + (size_t)g.own_lo * g.ld, cnt, ncclFloat64, c.rank - 1, comm, s));
        CNV_NCCL_CHECK(n.Recv(f + (size_t)(g.own_lo - depth) * g.ld, cnt, ncclFloat64, c.rank - 1, comm, s));
    }
    if (c.rank < c.world - 1) {
        CNV_NCCL_CHECK(n.Send(f + (size_t)(g.own_hi - depth) * g.ld, cnt, ncclFloat64, c.rank + 1, comm, s));
        CNV_NCCL_CHECK(n.Recv(f + (size_t)g.own_hi * g.ld, cnt, ncclFloat64, c.rank + 1, comm, s));
    }
}

void PoissonSolver::exchange_halos(double *field, int depth, cudaStream_t s)
{
    const NcclApi &n = nccl_api();
    CNV_NCCL_CHECK(n.GroupStart());
    halo_ops(n, comm_, geom_, field, depth, s);
    CNV_NCCL_CHECK(n.GroupEnd());
}

void PoissonSolver::enqueue_passes_dist(int npasses, cudaStream_t s)
{
    const NcclApi &n = nccl_api();
    ncclComm_t comm = (ncclComm_t)comm_.comm;
    for (int i = 0; i < npasses; i++) {
        enqueue_passes(1, s);  // local norms -> norms_
        CNV_NCCL_CHECK(n.GroupStart());
        // static pattern: pass p writes buffer (p+1)&1 (a "redo" pass only breaks it for the final pass)
        halo_ops(n, comm_, geom_, buf_[(dist_passes_ + 1) & 1], geom_.HY, s);
        for (int r = 0; r < comm_.world; r++) {
            if (r == comm_.rank) continue;
            CNV_NCCL_CHECK(n.Send(norms_, 8, ncclFloat64, r, comm, s));
            CNV_NCCL_CHECK(n.Recv(gather_ + 8 * r, 8, ncclFloat64, r, comm, s));
        }
        CNV_NCCL_CHECK(n.GroupEnd());
        CNV_CUDA_CHECK(cudaMemcpyAsync(gather_ + 8 * comm_.rank, norms_, sizeof(double) * 8, cudaMemcpyDeviceToDevice, s));
        k_decide_gather<<<1, 1, 0, s>>>(ctl_, gather_, comm_.world, T_, use_hist_ ? hist_ : nullptr);
        count_launch(1);
        dist_passes_++;
    }
    CNV_CUDA_CHECK(cudaGetLastError());
}

// continuity diagnostic of a slab run (src/main.c:407-408, maxel / minel): two[0] <- max over the ranks, two[1] <- min
void PoissonSolver::allreduce_max_min(double *two, cudaStream_t s)
{
    const NcclApi &n = nccl_api();
    ncclComm_t comm = (ncclComm_t)comm_.comm;
    CNV_NCCL_CHECK(n.GroupStart());
    CNV_NCCL_CHECK(n.AllReduce(two, two, 1, ncclFloat64, ncclMax, comm, s));
    CNV_NCCL_CHECK(n.AllReduce(two + 1, two + 1, 1, ncclFloat64, ncclMin, comm, s));
    CNV_NCCL_CHECK(n.GroupEnd());
}

// owned rows of a slab-local field -> rank 0 (device to device over NCCL), which assembles the whole field on the host:
// rank 0 passes host_out (gnrows x ncols) and a device staging array of at least (largest slab) x ld doubles
void PoissonSolver::gather_field_to_root(const double *field, double *stage, double *host_out, cudaStream_t s)
{
    const NcclApi &n = nccl_api();
    ncclComm_t comm = (ncclComm_t)comm_.comm;
    const int world = comm_.world, rank = comm_.rank, ld = geom_.ld, gn = geom_.gnrows;
    const int base = gn / world, rem = gn % world;
    if (rank != 0) {
        CNV_NCCL_CHECK(n.Send(field + (size_t)geom_.own_lo * ld, (size_t)(geom_.own_hi - geom_.own_lo) * ld, ncclFloat64, 0, comm, s));
        return;
    }
    if (host_out)
        CNV_CUDA_CHECK(cudaMemcpy2DAsync(host_out, sizeof(double) * geom_.ncols, field + (size_t)geom_.own_lo * ld, sizeof(double) * ld,
                                         sizeof(double) * geom_.ncols, geom_.own_hi - geom_.own_lo, cudaMemcpyDeviceToHost, s));
    for (int r = 1; r < world; r++) {
        const int r0 = r * base + (r < rem ? r : rem), rows = base + (r < rem ? 1 : 0);
        CNV_NCCL_CHECK(n.Recv(stage, (size_t)rows * ld, ncclFloat64, r, comm, s));  // (always: the sender does not know what rank 0 wants)
        if (host_out)
            CNV_CUDA_CHECK(cudaMemcpy2DAsync(host_out + (size_t)r0 * geom_.ncols, sizeof(double) * geom_.ncols, stage, sizeof(double) * ld,
                                             sizeof(double) * geom_.ncols, rows, cudaMemcpyDeviceToHost, s));
    }
    CNV_CUDA_CHECK(cudaStreamSynchronize(s));
}

void PoissonSolver::enqueue_decide(cudaStream_t s)
{
    k_decide<<<1, 1, 0, s>>>(ctl_, norms_, T_, use_hist_ ? hist_ : nullptr);
    count_launch(1);
}

PoissonResult PoissonSolver::solve(int itmax, double tol, cudaStream_t s, int *result_buf, bool keep_history)
{
    if (keep_history && hist_cap_ < itmax) {
        if (hist_) cudaFree(hist_);
        CNV_CUDA_CHECK(cudaMalloc(&hist_, sizeof(double) * (size_t)itmax));
        hist_cap_ = itmax;
    }
    use_hist_ = keep_history;
    reset_ctl(itmax, tol, s);
    // Grids whose iterate fits the register files: the whole solve in one persistent cooperative launch
    // (poisson_onchip.cu).  Same arithmetic, same red-black order -> same bits.
    if (use_onchip_ && !distributed_ && itmax > 0) {
        k_reset_ctl<<<1, 1, 0, s>>>(ctl_, itmax, tol, 3);  // (three buffers in rotation)
        launch_onchip(oc_, rc_, buf_[0], buf_[1], buf_[2], rhs_, ctl_, oc_flags_, oc_partials_, use_hist_ ? hist_ : nullptr, s, oc_prof_);
        launches_ += 1;
        count_launch(2);
        PoissonCtl c = read_ctl(s);
        if (result_buf) *result_buf = c.cur;
        PoissonResult r;
        r.status = c.state == 1 ? 0 : 1;
        r.k = c.result_k;
        r.sweeps = c.sweeps;
        r.passes = c.passes;
        r.e = c.state == 1 ? c.result_e : c.last_e;
        return r;
    }
    // Sweep counts drift slowly from one time step to the next (the shipped logs move by <= ~10
    // sweeps per step), so the first batch covers the previous solve's pass count plus one and is
    // followed by small batches.  Passes enqueued after convergence exit immediately.
    int batch = predicted_passes_ > 0 ? predicted_passes_ + 1 : 16;
    // (+2: the redo pass and the slack of the batching)
    const int max_passes = (itmax + T_ - 1) / T_ + 2;
    int enq = 0;
    PoissonCtl c;
    // slab solvers: the peer path exchanges inside the pass kernel; without it every pass is followed by the NCCL group
    // (halo rows + norm all-gather) and the decide kernel
    const bool nccl_path = distributed_ && !links_.enabled;
    if (nccl_path && !comm_.comm) {
        std::printf("** Error: slab solver without a communicator: attach one (cnv_poisson_attach_comm) or set up the peer path **\n");
        std::exit(1);
    }
    for (;;) {
        batch = std::max(1, std::min(batch, max_passes + 1 - enq));
        if (nccl_path) batch = std::min(batch, 32);  // a surplus (no-op) pass still costs its NCCL group on this path
        if (nccl_path) enqueue_passes_dist(batch, s);
        else enqueue_passes(batch, s);
        enq += batch;
        c = read_ctl(s);
        if (c.state != 0) break;
        if (enq > max_passes + 1) {
            std::printf("** Error: Poisson state machine did not terminate **\n");
            std::exit(1);
        }
        batch = predicted_passes_ > 0 ? 4 : 16;
    }
    predicted_passes_ = c.passes;
    if (result_buf) *result_buf = c.cur;
    PoissonResult r;
    r.status = c.state == 1 ? 0 : 1;
    r.k = c.result_k;
    r.sweeps = c.sweeps;
    r.passes = c.passes;
    r.e = c.state == 1 ? c.result_e : c.last_e;
    return r;
}

PoissonResult PoissonSolver::solve_from(const double *f, int ldf, double fsign, int itmax, double tol, cudaStream_t s,
                                        int *result_buf, bool keep_history)
{
    launch_prep_rhs(f, geom_.nrows, geom_.ncols, ldf, fsign, rc_.pscale, rhs_, buf_[0], buf_[1], geom_.ld, s);
    count_launch(1);
    return solve(itmax, tol, s, result_buf, keep_history);
}

}  // namespace cnv
