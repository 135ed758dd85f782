// peer_device.cuh -- device-side helpers of the multi-GPU peer-memory protocol (system-scope flags over NVLink, bounded
// spin-waits), shared by every kernel that takes part in it.  See poisson_stream.h
// (PeerMailbox, PeerLinks) for the protocol.
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
// Bounded spin-wait until *p >= want.  Gives up (returns false) when the peer is L.timeout_ns late (CNV_PEER_TIMEOUT_MS,
// default 5 minutes: a rank whose host is busy writing output or paused must not be taken for dead; a dead peer still
// never hangs the GPU) or as soon as ANY rank has reported a failure in this rank's mailbox (report_error), so that one
// timeout ends the run everywhere instead of being waited out rank by rank.
__device__ __forceinline__ bool wait_ge(const PeerLinks &L, const unsigned long long *p, unsigned long long want)
{
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = globaltimer_ns();
    const unsigned long long *err = &L.mail[L.rank]->error;
    unsigned spins = 0;
    while (ld_acquire_sys(p) < want) {
        if ((++spins & 63u) == 0 && (globaltimer_ns() - t0 > L.timeout_ns || ld_acquire_sys(err) != 0)) return false;
        __nanosleep(64);
    }
    return true;
}
// a wait failed: flag the error in every rank's mailbox (the hosts abort at their next state read-back)
__device__ __forceinline__ void report_error(const PeerLinks &L)
{
    for (int r = 0; r < L.world; r++) st_release_sys(&L.mail[r]->error, 1ull);
}

}  // namespace peerdev
}  // namespace cnv
