// poisson_resident.cu -- shared-memory-resident red-black SOR for small grids (<= ~360^2 cells).
//
// Measured (tools/probe_small.py, B200): 64^2 2.6 us/sweep vs 4.1 for the streaming kernel; 128^2 4.5 vs 3.9;
// 256^2 6.0 vs 3.7 -- barrier / cluster-barrier latency per half-sweep outweighs the saved launches from 128^2
// on, so by default only single-CTA grids take this path (see resident_plan).
//
// The reference's shipped configurations are 64^2 (config_default.txt) and 128^2 (config_high_re.txt).
// At those sizes a pass of the streaming kernel is pure launch + pipeline-fill latency, so the whole
// Poisson solve (src/poisson.c:224-285: all sweeps AND the convergence test after every sweep) runs in
// ONE launch here:
//   * a thread-block CLUSTER of C = 1, 2, 4 or 8 CTAs splits the rows; each CTA keeps its rows of psi in
//     shared memory (column-parity split, like the streaming kernel) plus one halo row above and below;
//   * after each half-sweep the CTAs refresh their halo rows straight from the neighbour CTA's shared memory
//     (distributed shared memory, cluster.map_shared_rank) -- no global-memory traffic inside the solve;
//   * the right-hand side of a thread's cells lives in registers for the whole solve;
//   * |u - u0| partials are reduced per CTA in a fixed order, exchanged through DSMEM and summed in rank
//     order by every CTA, so all CTAs take the same stop decision (first sweep with e < tol, src/poisson.c:273).
// Cell updates use exact.h::relax, i.e. the same separately rounded operation sequence as everywhere else;
// the red-black order makes the result independent of how the rows are split: fields are bit-identical to the
// streaming kernel and to the reference.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace cnv {

constexpr int kResThreads = 1024;
constexpr int kResMaxM = 8;  // rows per thread (register-resident right-hand sides: 2 colours x kResMaxM doubles)

// block-wide sum in a fixed order; every thread returns the total
__device__ __forceinline__ double block_sum(double v, double *scratch /* 33 doubles */)
{
    for (int o = 16; o > 0; o >>= 1) v = xadd(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double s = scratch[lane];  // 32 warps
        for (int o = 16; o > 0; o >>= 1) s = xadd(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (lane == 0) scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

template <bool POW2>
__global__ void __launch_bounds__(kResThreads, 1)
k_poisson_resident(const ResidentGeom g, const RelaxConsts rc, const double *__restrict__ psi0, const double *__restrict__ rhs,
                   double *__restrict__ out, PoissonCtl *ctl, double *hist, const int itmax, const double tol)
{
    extern __shared__ double4 smraw[];
    double *SE = reinterpret_cast<double *>(smraw);  // [(RPC + 2)][PK]  even columns
    double *SO = SE + (size_t)(g.RPC + 2) * g.PK;     // [(RPC + 2)][PK]  odd columns
    __shared__ double scratch[33];
    __shared__ double part;                           // this CTA's norm partial of the current sweep

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = g.C > 1 ? (int)cluster.block_rank() : 0;
    const int r0 = rank * g.RPC;
    const int r1 = r0 + g.RPC < g.nrows ? r0 + g.RPC : g.nrows;  // owned rows [r0, r1)
    const int tid = threadIdx.x;
    const int kk = tid % g.KP, rsub = tid / g.KP;
    const bool tact = rsub < g.RPB;  // threads beyond KP*RPB idle

    // ---- load psi (owned rows + halo rows) and zero the pads ----
    for (int idx = tid; idx < (g.RPC + 2) * g.PK; idx += kResThreads) {
        const int lr = idx / g.PK, pk = idx - lr * g.PK;
        const int i = r0 - 1 + lr, k = pk - 1;
        double e = 0.0, o = 0.0;
        if (i >= 0 && i < g.nrows && k >= 0 && k < g.KP) {
            const double *row = psi0 + (size_t)i * g.ld;
            e = row[2 * k];
            o = 2 * k + 1 < g.ld ? row[2 * k + 1] : 0.0;
        }
        SE[idx] = e;
        SO[idx] = o;
    }
    // ---- this thread's cells: rows i = r0 + rsub + RPB*m; type (even/odd column) of its red cells is fixed ----
    // colour 0 = red = (i + j) even (src/poisson.c:247): in row i the red cells are the even columns iff i is even
    const int par = (r0 + rsub) & 1;  // RPB is even, so the row parity does not depend on m
    double Pr[kResMaxM], Pb[kResMaxM];
    unsigned vr = 0, vb = 0;  // bit m: the red / black cell of row m is updatable
#pragma unroll
    for (int m = 0; m < kResMaxM; m++) {
        const int i = r0 + rsub + g.RPB * m;
        const int jr = 2 * kk + par, jb = 2 * kk + (par ^ 1);
        const bool rowok = tact && i < r1 && i >= 1 && i <= g.nrows - 2;
        Pr[m] = Pb[m] = 0.0;
        if (rowok && jr >= 1 && jr <= g.ncols - 2) { vr |= 1u << m; Pr[m] = rhs[(size_t)i * g.ld + jr]; }
        if (rowok && jb >= 1 && jb <= g.ncols - 2) { vb |= 1u << m; Pb[m] = rhs[(size_t)i * g.ld + jb]; }
    }
    __syncthreads();

    // E cell (col 2k):   N/S = SE[lr+-1][k], W = SO[lr][k-1], E = SO[lr][k]
    // O cell (col 2k+1): N/S = SO[lr+-1][k], W = SE[lr][k],   E = SE[lr][k+1]
    auto half_sweep = [&](const int colour, double &acc) {
        const int type = par ^ colour;  // 0: this thread's cells of this colour are even-column cells
        double *A = type ? SO : SE;
        const double *B = type ? SE : SO;
        const unsigned valid = colour ? vb : vr;
#pragma unroll
        for (int m = 0; m < kResMaxM; m++) {
            if (!((valid >> m) & 1u)) continue;
            const int c = (1 + rsub + g.RPB * m) * g.PK + 1 + kk;
            const double own = A[c];
            const double W = type ? B[c] : B[c - 1], E = type ? B[c + 1] : B[c];
            const double nv = relax<POW2>(A[c + g.PK], A[c - g.PK], E, W, own, colour ? Pb[m] : Pr[m], rc);
            A[c] = nv;
            acc = xadd(acc, fabs(xsub(nv, own)));
        }
    };
    // refresh the halo rows from the neighbour CTAs' shared memory (both parity arrays of the row)
    auto halo_refresh = [&]() {
        if (g.C == 1) return;
        const int nown = r1 - r0;
        if (rank > 0 && tid < 2 * g.PK) {  // row below my first row = last owned row of rank-1 (always RPC rows there)
            double *mine = (tid < g.PK ? SE : SO) + (tid % g.PK);
            const double *theirs = cluster.map_shared_rank(mine, rank - 1) + (size_t)g.RPC * g.PK;
            *mine = *theirs;
        }
        if (rank < g.C - 1 && r1 < g.nrows && tid >= 2 * g.PK && tid < 4 * g.PK) {  // first owned row of rank+1
            const int t2 = tid - 2 * g.PK;
            double *base = (t2 < g.PK ? SE : SO) + (t2 % g.PK);
            const double *theirs = cluster.map_shared_rank(base, rank + 1) + (size_t)1 * g.PK;
            base[(size_t)(nown + 1) * g.PK] = *theirs;
        }
    };
    auto sync_all = [&]() {
        if (g.C > 1) cluster.sync();
        else __syncthreads();
    };

    int k = 0, state = 2;
    double e = 0.0;
    for (; k < itmax; k++) {
        double acc = 0.0;
        half_sweep(0, acc);
        sync_all();
        halo_refresh();
        __syncthreads();
        half_sweep(1, acc);
        const double mine = block_sum(acc, scratch);
        if (tid == 0) part = mine;
        sync_all();  // black cells and the partials of every CTA are final
        halo_refresh();
        e = 0.0;
        if (g.C > 1) {
            for (int r = 0; r < g.C; r++) e = xadd(e, *cluster.map_shared_rank(&part, r));
        } else {
            e = part;
        }
        __syncthreads();
        if (rank == 0 && tid == 0 && hist) hist[k] = e;
        if (e < tol) { state = 1; k++; break; }
        // (the next write of `part` lies behind the next sweep's first cluster barrier)
    }
    if (g.C > 1) cluster.sync();  // nobody may exit while a neighbour still reads its shared memory

    // ---- write back the owned rows (ring and padding columns included) ----
    const int nown = r1 - r0;
    for (int idx = tid; idx < nown * g.KP; idx += kResThreads) {
        const int lr = idx / g.KP, kq = idx - lr * g.KP;
        const int c = (1 + lr) * g.PK + 1 + kq;
        double *row = out + (size_t)(r0 + lr) * g.ld;
        row[2 * kq] = SE[c];
        if (2 * kq + 1 < g.ld) row[2 * kq + 1] = SO[c];
    }
    if (rank == 0 && tid == 0) {
        PoissonCtl c = *ctl;
        c.state = state;
        c.cur = 1;  // result in buffer 1
        c.sweeps = k;
        c.result_k = k - 1;
        c.result_e = e;
        c.last_e = e;
        c.passes = 1;
        c.redo = 0;
        c.ticket = 0;
        *ctl = c;
    }
}

// ---- host side ----------------------------------------------------------------------------------
bool resident_plan(int nrows, int ncols, int ld, size_t smem_limit, ResidentGeom *g, size_t *smem)
{
    // CNV_POISSON_RESIDENT: 0 never, 1 (default) only where it was measured faster than the streaming kernel
    // (single-CTA grids with <= 2 rows per thread and colour, e.g. the reference's 64^2 default config:
    // 0.29 vs 0.42 ms per time step), 2 whenever a plan exists (clusters up to 8 CTAs; used by the tests).
    int mode = 1;
    if (const char *e = std::getenv("CNV_POISSON_RESIDENT")) mode = std::atoi(e);
    if (mode <= 0) return false;
    const long cells = (long)nrows * ncols;
    if (cells > 8L * kResMaxM * kResThreads * 2) return false;
    g->nrows = nrows; g->ncols = ncols; g->ld = ld;
    g->KP = (ncols + 1) / 2;
    if (g->KP > kResThreads / 2) return false;
    g->PK = g->KP + 2;
    g->RPB = (kResThreads / g->KP) & ~1;  // even
    if (g->RPB < 2) return false;
    // smallest cluster whose threads hold <= 2 rows per colour (latency); else the largest feasible cluster
    int bestC = 0;
    for (int C = 1; C <= 8; C *= 2) {
        const int RPC = (nrows + C - 1) / C;
        if ((long)RPC * (C - 1) >= nrows) continue;  // every CTA but the last owns exactly RPC rows, the last >= 1
        if (RPC > kResMaxM * g->RPB) continue;
        if ((size_t)2 * (RPC + 2) * g->PK * sizeof(double) > smem_limit) continue;
        bestC = C;
        if (RPC <= 2 * g->RPB) break;
    }
    if (!bestC) return false;
    if (mode == 1 && !(bestC == 1 && (nrows + bestC - 1) / bestC <= 2 * g->RPB)) return false;
    g->C = bestC;
    g->RPC = (nrows + bestC - 1) / bestC;
    *smem = (size_t)2 * (g->RPC + 2) * g->PK * sizeof(double);
    return true;
}

template <bool POW2>
static void launch_resident_t(const ResidentGeom &g, size_t smem, const RelaxConsts &rc, const double *psi0, const double *rhs,
                              double *out, PoissonCtl *ctl, double *hist, int itmax, double tol, cudaStream_t s)
{
    static size_t configured[kMaxDevices] = {};  // per device: function attributes belong to the device context
    size_t &conf = configured[current_device_slot()];
    if (smem > 48 * 1024 && smem > conf) {
        CNV_CUDA_CHECK(cudaFuncSetAttribute(k_poisson_resident<POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.C);
    cfg.blockDim = dim3(kResThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g.C > 1 ? 1 : 0;
    CNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_poisson_resident<POW2>, g, rc, psi0, rhs, out, ctl, hist, itmax, tol));
}

void launch_resident(const ResidentGeom &g, size_t smem, const RelaxConsts &rc, const double *psi0, const double *rhs, double *out,
                     PoissonCtl *ctl, double *hist, int itmax, double tol, cudaStream_t s)
{
    if (rc.pow2) launch_resident_t<true>(g, smem, rc, psi0, rhs, out, ctl, hist, itmax, tol, s);
    else launch_resident_t<false>(g, smem, rc, psi0, rhs, out, ctl, hist, itmax, tol, s);
}

}  // namespace cnv
