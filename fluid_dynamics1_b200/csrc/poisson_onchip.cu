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
// bounded spin until *p >= want (another CTA of this cooperative grid raises it); false on timeout or reported error
__device__ __forceinline__ bool oc_wait(const unsigned long long *p, unsigned long long want, const unsigned long long *err,
                                        unsigned long long timeout_ns)
{
    if (ld_acquire_gpu(p) >= want) return true;
    const unsigned long long t0 = peerdev::globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_gpu(p) < want) {
        if ((++spins & 255u) == 0 && (peerdev::globaltimer_ns() - t0 > timeout_ns || ld_acquire_gpu(err) != 0)) return false;
    }
    return true;
}

}  // namespace

// flags[cta] = passes completed by that CTA (cumulative within the launch); flags[gridDim] = error word.
// partials[slot][cta][8]: per-sweep norm partials of pass (slot = pass % kOcNormSlots).
template <bool POW2>
__global__ void __launch_bounds__(kOcMaxThreads, 1)
k_poisson_onchip(const OnchipGeom g, const RelaxConsts rc, double *__restrict__ buf0, double *__restrict__ buf1,
                 double *__restrict__ buf2, const double *__restrict__ rhs, PoissonCtl *ctl, unsigned long long *flags,
                 double *partials, double *hist, const unsigned long long timeout_ns)
{
    extern __shared__ double4 sm4[];
    double *sm = reinterpret_cast<double *>(sm4);
    __shared__ double s_part[8][kOcMaxThreads / 32];  // [sweep of the pass][warp]
    __shared__ double s_sub[8][8];                    // norm gather: [sweep][sub-sum]
    __shared__ double s_e[8];
    __shared__ LagAction s_act;
    __shared__ PoissonCtl s_ctl;                      // chain state X_p (thread 0 evolves it)
    __shared__ int s_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    const int NT = oc_threads(g);
    const bool active = tid < NT;
    const bool first_cta = cta == 0;
    unsigned long long *err = flags + ncta;
    double *bufs[3] = {buf0, buf1, buf2};

    for (int i = tid; i < kOcSlots * kOcPitch; i += blockDim.x) sm[i] = 0.0;
    if (tid == 0) { s_ctl = *ctl; s_bad = 0; }
    const OcTile tile = oc_tile(g, blockIdx.x, blockIdx.y);
    int nb[8];
    const int nnb = oc_neighbours(g, blockIdx.x, blockIdx.y, nb);
    OcThread t;
    oc_thread_init(t, g, tile, active ? tid : 0, sm);
    if (!active) { t.upd = t.ownp = t.halop = t.inp = 0; t.fast = false; t.dist = 1 << 30; }
    __syncthreads();
    if (active) {
        oc_load_rhs(t, rhs);
        oc_load_psi(t, bufs[s_ctl.cur], t.inp);
        oc_publish_all(t, sm);
    }
    __syncthreads();

    int P = 0;  // passes started
    for (int p = 0;; p++) {
        // ---- X_p = lag_fold(X_{p-1}, norms of pass p-2), summed by every CTA itself in a fixed order ----
        const bool need = p >= 2 && s_ctl.state == 0 && s_ctl.redo == 0;  // uniform (s_ctl is stable here)
        if (need) {
            const unsigned long long want = (unsigned long long)(p - 1);  // every CTA has completed pass p-2
            for (int c = tid; c < ncta; c += blockDim.x)
                if (!oc_wait(&flags[c], want, err, timeout_ns)) s_bad = 1;
            __syncthreads();
            const double *part = partials + (size_t)((p - 2) % kOcNormSlots) * ncta * 8;
            if (tid < 64) {  // thread (s, k): CTAs k, k+8, k+16, ... of sweep s
                const int s = tid >> 3, k = tid & 7;
                double sum = 0.0;
                for (int c = k; c < ncta; c += 8) sum = xadd(sum, __ldcg(&part[(size_t)c * 8 + s]));
                s_sub[s][k] = sum;
            }
            __syncthreads();
            if (tid < 8) {
                double sum = 0.0;
                for (int k = 0; k < 8; k++) sum = xadd(sum, s_sub[tid][k]);
                s_e[tid] = sum;
            }
            __syncthreads();
        }
        if (tid == 0) {
            PoissonCtl c = s_ctl;
            if (s_bad) {
                c.state = 3;
                st_release_gpu(err, 1ull);
            } else if (p >= 2) {
                double e[8];
                for (int i = 0; i < 8; i++) e[i] = need ? s_e[i] : 0.0;
                lag_fold(c, e, g.T, first_cta ? hist : nullptr);
            }
            s_ctl = c;
            s_act = lag_action(c, p, g.T);
        }
        __syncthreads();
        const LagAction act = s_act;
        P = p;
        if (act.kind == 0) break;
        double *__restrict__ out = bufs[act.out];
        if (act.kind == 2) {  // recompute the converged pass from its (intact) input with exactly act.nsw sweeps
            if (active) {
                oc_load_psi(t, bufs[act.in], t.inp);
                oc_publish_all(t, sm);
            }
            __syncthreads();
        }
        const int nsw = act.nsw;
        for (int s = 0; s < nsw; s++) {
            double acc = 0.0;
            // a cell at distance d from the output region matters only while >= d half-sweeps remain after this one
            const int rem0 = 2 * (nsw - s) - 1;
            if (t.dist <= rem0) {
                if (tile.par0 == 0) oc_half_sweep<POW2, 0>(t, rc, sm, acc); else oc_half_sweep<POW2, 1>(t, rc, sm, acc);
            }
            __syncthreads();
            if (t.dist <= rem0 - 1) {
                if (tile.par0 == 0) oc_half_sweep<POW2, 1>(t, rc, sm, acc); else oc_half_sweep<POW2, 0>(t, rc, sm, acc);
            }
            // this sweep's norm: fixed-order warp sum -> one slot per warp (read after the pass' last barrier)
            for (int o = 16; o > 0; o >>= 1) acc = xadd(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if (lane == 0) s_part[s][warp] = acc;
            __syncthreads();
        }
        if (active) oc_store(t, out);
        if (tid < 8) {
            double e = 0.0;
            if (tid < nsw) {
                const int nw = (blockDim.x + 31) >> 5;
                for (int w = 0; w < nw; w++) e = xadd(e, s_part[tid][w]);
            }
            partials[((size_t)(p % kOcNormSlots) * ncta + cta) * 8 + tid] = e;
        }
        __syncthreads();  // every store of this CTA has been issued ...
        if (tid == 0) {
            __threadfence();  // ... and is visible GPU-wide before the flag
            st_release_gpu(&flags[cta], (unsigned long long)(p + 1));
        }
        if (act.kind == 2) continue;  // final pass: the next iteration folds it and stops
        // ---- halo cells of the next pass = the neighbour tiles' output of this one ----
        if (tid < nnb) {
            if (!oc_wait(&flags[nb[tid]], (unsigned long long)(p + 1), err, timeout_ns)) s_bad = 1;
        }
        __syncthreads();
        if (active && t.halop) {
            oc_load_psi(t, out, t.halop);
            oc_publish_all(t, sm);
        }
        __syncthreads();
    }

    // ---- host-visible state after P passes: X_P, completed with the norms of pass P-1 where they still count ----
    if (!first_cta) return;
    if (s_ctl.state != 3 && lag_final_needs_last(s_ctl, P)) {
        for (int c = tid; c < ncta; c += blockDim.x)
            if (!oc_wait(&flags[c], (unsigned long long)P, err, timeout_ns)) s_bad = 1;
        __syncthreads();
        const double *part = partials + (size_t)((P - 1) % kOcNormSlots) * ncta * 8;
        if (tid < 64) {
            const int s = tid >> 3, k = tid & 7;
            double sum = 0.0;
            for (int c = k; c < ncta; c += 8) sum = xadd(sum, __ldcg(&part[(size_t)c * 8 + s]));
            s_sub[s][k] = sum;
        }
        __syncthreads();
        if (tid < 8) {
            double sum = 0.0;
            for (int k = 0; k < 8; k++) sum = xadd(sum, s_sub[tid][k]);
            s_e[tid] = sum;
        }
        __syncthreads();
        if (tid == 0) {
            PoissonCtl c = s_ctl;
            if (s_bad) c.state = 3;
            else lag_final(c, P, s_e, g.T, hist);
            s_ctl = c;
        }
        __syncthreads();
    }
    if (tid == 0) {
        PoissonCtl c = s_ctl;
        c.ticket = 0;
        *ctl = c;
    }
}

template <bool POW2>
static void launch_onchip_t(const OnchipGeom &g, const RelaxConsts &rc, double *b0, double *b1, double *b2, const double *rhs,
                            PoissonCtl *ctl, unsigned long long *flags, double *partials, double *hist, cudaStream_t s)
{
    const size_t smem = oc_smem_bytes();
    static size_t configured[kMaxDevices] = {};
    size_t &conf = configured[current_device_slot()];
    if (smem > conf) {
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_onchip<POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    const int ncta = g.ntx * g.nty;
    CNV_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(unsigned long long) * (ncta + 1), s));
    static const unsigned long long timeout_ns = []() {
        const char *e = std::getenv("CNV_ONCHIP_TIMEOUT_MS");
        return (unsigned long long)(e ? std::atoi(e) : 10000) * 1000000ull;
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.ntx, g.nty);
    cfg.blockDim = dim3(round_up(oc_threads(g), 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's flags
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_onchip<POW2>, g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, timeout_ns));
}

void launch_onchip(const OnchipGeom &g, const RelaxConsts &rc, double *b0, double *b1, double *b2, const double *rhs, PoissonCtl *ctl,
                   unsigned long long *flags, double *partials, double *hist, cudaStream_t s)
{
    if (rc.pow2) launch_onchip_t<true>(g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, s);
    else launch_onchip_t<false>(g, rc, b0, b1, b2, rhs, ctl, flags, partials, hist, s);
}

}  // namespace cnv
