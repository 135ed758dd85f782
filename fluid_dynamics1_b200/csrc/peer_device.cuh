// peer_device.cuh -- device-side helpers of the multi-GPU peer-memory protocol (system-scope flags over NVLink, bounded
// spin-waits) for kernels other than the streaming pass kernel, which carries its own copies in poisson.cu (measured code
// is left untouched).  See poisson_stream.h (PeerMailbox, PeerLinks) for the protocol.
#pragma once
#include "poisson_stream.h"

namespace cnv {
namespace peerdev {

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr unsigned long long kTimeoutNs = 4000000000ull;  // a peer that is 4 s late is dead: never hang the GPU
__device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long want)
{
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(p) < want) {
        if (globaltimer_ns() - t0 > kTimeoutNs) return false;
        __nanosleep(64);
    }
    return true;
}

// Start of a pass (whole CTA; contains barriers): derive this pass' state from the previous one and, if needed, every
// rank's norms of the pass before (plain machine only -- the lagged machine is a feature of the streaming kernel), store
// it for the next pass (first CTA), and advance the protocol counters if the pass has nothing to do.
// s_nrm: kMaxRanks x 8 doubles, s_flag: 1 int of shared memory.
__device__ __forceinline__ LagAction begin_pass(const PeerLinks &L, int T, double *hist, bool first_cta, double (*s_nrm)[8], int *s_flag,
                                                PoissonCtl *s_ctl, LagAction *s_act)
{
    const int tid = threadIdx.x;
    const PoissonCtl prev = L.ctlbuf[L.pidx == 0 ? 0 : (L.pidx - 1) & 1];
    const bool need = peer_needs_norms(prev, L.pidx, 0);  // uniform
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (need) {
        PeerMailbox *mb = L.mail[L.rank];
        for (int r = tid; r < L.world; r += blockDim.x)
            if (!wait_ge(&mb->norm_flag[r], L.gidx)) *s_flag = 1;  // pass gidx-1 publishes the value gidx
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
        const bool bad = need && *s_flag != 0;
        if (bad) atomicExch(&L.mail[L.rank]->error, 1ull);
        const LagAction a = peer_advance(c, e, need, bad, L.pidx, 0, T, first_cta ? hist : nullptr);
        *s_ctl = c;
        *s_act = a;
        if (first_cta && L.pidx > 0) L.ctlbuf[L.pidx & 1] = c;
        if (first_cta && a.kind == 0) {  // a no-op pass still advances the cumulative counters of the protocol
            if (L.rank > 0) atomicAdd_system(&L.mail[L.rank - 1]->halo_count[1], L.push_low);
            if (L.rank < L.world - 1) atomicAdd_system(&L.mail[L.rank + 1]->halo_count[0], L.push_high);
            for (int r = 0; r < L.world; r++) st_release_sys(&L.mail[r]->norm_flag[L.rank], L.gidx + 1);
        }
    }
    __syncthreads();
    return *s_act;
}

// Before a CTA touches data: the neighbour has re-initialised the buffers this CTA pushes into (first pass of a solve), and
// the neighbour's pushes of the previous pass have landed in the halo rows this CTA reads.  One thread; follow with a barrier.
__device__ __forceinline__ void wait_inputs(const PeerLinks &L, bool push_down, bool push_up, bool reads_low, bool reads_high)
{
    PeerMailbox *mb = L.mail[L.rank];
    bool ok = true;
    if (L.pidx == 0 && push_down) ok &= wait_ge(&mb->ready[0], L.epoch);
    if (L.pidx == 0 && push_up) ok &= wait_ge(&mb->ready[1], L.epoch);
    if (reads_low) ok &= wait_ge(&mb->halo_count[0], L.gidx * L.need_low);
    if (reads_high) ok &= wait_ge(&mb->halo_count[1], L.gidx * L.need_high);
    if (!ok) atomicExch(&mb->error, 1ull);
}

// Copy rows [ra, rb) x columns [c0, c1) (c0 even, pairs) of `out` (this rank's freshly written iterate, still in L2) into the
// neighbour's halo copy with coalesced 16-byte peer stores.  Whole CTA.
__device__ __forceinline__ void push_rows(const double *out, double *peer, int ld, int ra, int rb, int c0, int c1)
{
    const int npair = (c1 - c0) >> 1;
    for (int idx = threadIdx.x; idx < (rb - ra) * npair; idx += blockDim.x) {
        const int rr = ra + idx / npair, cc = c0 + 2 * (idx % npair);
        const size_t off = (size_t)rr * ld + cc;
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(out + off));
        *reinterpret_cast<double2 *>(peer + off) = v;
    }
}

// Last CTA of the pass: this rank's per-sweep norms into every rank's mailbox, then the flag everywhere.  Whole CTA.
__device__ __forceinline__ void publish_norms(const PeerLinks &L, const double *s_e, int nsw)
{
    const int slot = (int)(L.gidx & (kNormSlots - 1));
    for (int i = threadIdx.x; i < 8 * L.world; i += blockDim.x) {
        const int r = i >> 3, g = i & 7;
        L.mail[r]->norms[slot][L.rank][g] = g < nsw ? s_e[g] : 0.0;
    }
    __threadfence_system();
    __syncthreads();
    for (int r = threadIdx.x; r < L.world; r += blockDim.x) st_release_sys(&L.mail[r]->norm_flag[L.rank], L.gidx + 1);
}

}  // namespace peerdev
}  // namespace cnv
