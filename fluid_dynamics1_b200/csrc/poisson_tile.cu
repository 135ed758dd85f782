// poisson_tile.cu -- the stationary-tile pass kernel of the Poisson solve (see poisson_tile.h) and its planner.
//
// Replaces the same reference code as the streaming kernel (src/poisson.c:224-285: T red-black SOR sweeps and the
// sum |u - u0| of each of them per launch) for grids small enough that the streaming kernel's pipeline fill and
// y-halo dominate.  Control flow (PoissonCtl, decide(): first sweep with e < tol, "redo" pass, itmax) is shared.
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "peer_device.cuh"

namespace cnv {


// PEER: slab of a multi-GPU run with the in-kernel peer-memory exchange (CNV_TILE_PEER=1; the protocol of the streaming
// kernel, poisson_stream.h PeerMailbox / PeerLinks, plain stop machine): the pass derives its state from every rank's
// published norms, tiles that read halo rows wait for the neighbours' pushes of the previous pass, tiles whose output rows
// lie in the 2T boundary band store them into the neighbour's halo rows over NVLink, the last CTA publishes the norms.
template <int M, bool POW2, bool PEER>
__global__ void __launch_bounds__(tile_max_threads(M), 1)
k_poisson_tile(const TileGeom g, const RelaxConsts rc, double *__restrict__ buf0, double *__restrict__ buf1,
               const double *__restrict__ rhs, PoissonCtl *ctl, double *__restrict__ partials, double *hist, double *norms_out,
               const int fused_decide, const PeerLinks L)
{
    extern __shared__ double4 sm4[];
    double *sm = reinterpret_cast<double *>(sm4);
    __shared__ double s_part[8][32];  // [sweep of the pass][warp]
    __shared__ double s_e[8];
    __shared__ int s_last;

    // programmatic dependent launch (only when launched with the attribute, CNV_TILE_PDL=1; otherwise both are no-ops):
    // the next pass may become resident while this one drains and waits here for its completion
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    __shared__ PoissonCtl s_ctl;
    __shared__ LagAction s_act;
    int cur, nxt, nsw;
    if (PEER) {
        __shared__ double s_nrm[kMaxRanks][8];
        __shared__ int s_flag;
        const LagAction a = peerdev::begin_pass(L, g.T, hist, blockIdx.x == 0 && blockIdx.y == 0, s_nrm, &s_flag, &s_ctl, &s_act);
        if (a.kind == 0) return;  // solve already finished: later passes of a batch are no-ops
        cur = a.in; nxt = a.out; nsw = a.nsw;
    } else {
        const PoissonCtl c0 = *ctl;
        if (c0.state != 0) return;  // solve already finished: later passes of a batch are no-ops
        nsw = pass_sweeps(c0, g.T);
        cur = c0.cur; nxt = cur ^ 1;
    }
    const double *__restrict__ in = cur ? buf1 : buf0;
    double *__restrict__ out = nxt ? buf1 : buf0;
    const TileRows rows = tile_rows_of(g, blockIdx.y);
    const bool push_down = PEER && L.rank > 0 && rows.pa[0] < rows.pb[0];
    const bool push_up = PEER && L.rank < L.world - 1 && rows.pa[1] < rows.pb[1];
    if (PEER) {
        if (tid == 0)
            peerdev::wait_inputs(L, push_down, push_up, L.rank > 0 && rows.rlo < g.own_lo, L.rank < L.world - 1 && rows.rhi > g.own_hi);
        __syncthreads();
    }

    // the block is rounded up to whole warps (shuffles, barriers); the surplus threads own no cells
    const bool active = tid < g.KP * g.NSEG;
    TileThread<M> t;
    t.acc = 0.0;
    if (active) tile_load<M>(t, g, blockIdx.x, blockIdx.y, tid, sm, in, rhs);
    // colour 0 (red) = (global row + column) even; tile column 0 is even, so in tile row 0 the red cell of a pair is
    // its even column iff the global row of tile row 0 is even
    const int par0 = (g.grow0 + g.own_lo + (int)blockIdx.y * g.OH - g.HT) & 1;
    __syncthreads();
    for (int s = 0; s < nsw; s++) {
        if (active) {
            if (par0 == 0) tile_half_sweep<M, POW2, 0>(t, rc, sm); else tile_half_sweep<M, POW2, 1>(t, rc, sm);
        }
        __syncthreads();
        if (active) {
            if (par0 == 0) tile_half_sweep<M, POW2, 1>(t, rc, sm); else tile_half_sweep<M, POW2, 0>(t, rc, sm);
        }
        // this sweep's norm: fixed-order warp sum -> one slot per warp
        double a = t.acc;
        t.acc = 0.0;
        for (int o = 16; o > 0; o >>= 1) a = xadd(a, __shfl_xor_sync(0xffffffffu, a, o));
        if (lane == 0) s_part[s][warp] = a;
        __syncthreads();
    }
    if (active) tile_store<M>(t, out);

    if (PEER && (push_down || push_up)) {
        // boundary rows this tile has just written (still L2 resident) -> the neighbour's halo rows, then count the push
        __syncthreads();  // every write-back of this CTA is done and visible to the CTA
        const int c0 = (int)blockIdx.x * g.OW, c1 = c0 + g.OW < g.ld ? c0 + g.OW : g.ld;
        if (push_down) peerdev::push_rows(out, L.down_buf[nxt] + L.down_delta, g.ld, rows.pa[0], rows.pb[0], c0, c1);
        if (push_up) peerdev::push_rows(out, L.up_buf[nxt] + L.up_delta, g.ld, rows.pa[1], rows.pb[1], c0, c1);
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            if (push_down) atomicAdd_system(&L.mail[L.rank - 1]->halo_count[1], 1ull);
            if (push_up) atomicAdd_system(&L.mail[L.rank + 1]->halo_count[0], 1ull);
        }
    }

    if (tid < nsw) {
        double e = 0.0;
        for (int w = 0; w < nwarps; w++) e = xadd(e, s_part[tid][w]);
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        partials[(size_t)cta * 8 + tid] = e;
        __threadfence();
    }
    __syncthreads();
    const int ncta = gridDim.x * gridDim.y;
    unsigned *ticket = PEER ? &L.mail[L.rank]->ticket : &ctl->ticket;
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)ncta - 1;
    __syncthreads();
    if (!s_last) return;

    // last CTA: grid-wide sums in a fixed order (warp w <-> sweep w), then the stopping decision (src/poisson.c:272-279)
    __threadfence();
    for (int w = warp; w < 8; w += nwarps) {
        double e = 0.0;
        if (w < nsw) {
            for (int c = lane; c < ncta; c += 32) e = xadd(e, __ldcg(&partials[(size_t)c * 8 + w]));
            for (int o = 16; o > 0; o >>= 1) e = xadd(e, __shfl_xor_sync(0xffffffffu, e, o));
        }
        if (lane == 0) s_e[w] = e;
    }
    __syncthreads();
    if (PEER) {
        peerdev::publish_norms(L, s_e, nsw);
        if (tid == 0) *ticket = 0;
        return;
    }
    if (tid == 0) {
        if (fused_decide) {
            PoissonCtl c = *ctl;
            decide(c, s_e, nsw, hist);
            c.ticket = 0;
            *ctl = c;
        } else {
            for (int i = 0; i < 8; i++) norms_out[i] = i < nsw ? s_e[i] : 0.0;
            ctl->ticket = 0;
        }
    }
}

template <int M, bool POW2, bool PEER>
static void launch_tile_t(const TileGeom &g, const RelaxConsts &rc, double *b0, double *b1, const double *rhs, PoissonCtl *ctl,
                          double *partials, double *hist, double *norms, int fused, cudaStream_t s, const PeerLinks &L)
{
    const size_t smem = tile_smem_bytes(g);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_tile<M, POW2, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // CNV_TILE_PDL=1: programmatic stream serialisation between consecutive passes (hides the launch latency, which is a
    // large share of a pass on the small grids this kernel is for); off by default until measured
    static const bool pdl = std::getenv("CNV_TILE_PDL") && std::atoi(std::getenv("CNV_TILE_PDL")) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.ntx, g.nty);
    cfg.blockDim = dim3(round_up(g.KP * g.NSEG, 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const double *crhs = rhs;
    CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_tile<M, POW2, PEER>, g, rc, b0, b1, crhs, ctl, partials, hist, norms, fused, L));
}

void launch_tile_pass(const TileGeom &g, const RelaxConsts &rc, double *b0, double *b1, const double *rhs, PoissonCtl *ctl,
                      double *partials, double *hist, double *norms, int fused, cudaStream_t s, const PeerLinks &L)
{
#define CNV_TILE(MM)                                                                                                       \
    if (g.M == MM) {                                                                                                       \
        if (L.enabled) {                                                                                                   \
            if (rc.pow2) launch_tile_t<MM, true, true>(g, rc, b0, b1, rhs, ctl, partials, hist, norms, fused, s, L);       \
            else launch_tile_t<MM, false, true>(g, rc, b0, b1, rhs, ctl, partials, hist, norms, fused, s, L);              \
        } else {                                                                                                           \
            if (rc.pow2) launch_tile_t<MM, true, false>(g, rc, b0, b1, rhs, ctl, partials, hist, norms, fused, s, L);      \
            else launch_tile_t<MM, false, false>(g, rc, b0, b1, rhs, ctl, partials, hist, norms, fused, s, L);             \
        }                                                                                                                  \
    }
    CNV_TILE(6) CNV_TILE(8) CNV_TILE(10) CNV_TILE(12) CNV_TILE(14) CNV_TILE(16)
#undef CNV_TILE
}

}  // namespace cnv
