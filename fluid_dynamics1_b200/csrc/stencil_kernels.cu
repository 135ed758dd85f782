// stencil_kernels.cu -- explicit vorticity-transport stencils of the cnavier time step, matrix-free.
//
// Restates, as CUDA kernels over pitched row-major fp64 fields A[i*ld + j] (i = first index of the
// reference's mtrx, j = second), the parts of the reference time loop that are not the Poisson
// solve (src/main.c:283-395):
//   k_ring_bc_vorticity  Dirichlet BCs on u,v (:283-296) + wall vorticity w = dvdx - dudy on the
//                        ring (:298-320)
//   k_euler_fused        dwdx,dwdy,d2wdx2,d2wdy2 (:323-341) + euler (src/fluiddyn.c:71-102) on all
//                        points + the Poisson right-hand side -w scaled for the solver (:348)
//   k_velocity           u = DY psi, v = -(DX psi) on all points (:366-383)
//   k_continuity         max/min of DX u + DY v (:387-395, :407-408)
//   k_apply / pointwise  single-operator forms behind the drop-in C signatures
// DX differentiates along j with spacing dx, DY along i with spacing dy (kron(I,d_x) / kron(d_y,I),
// src/main.c:149-152).  Arithmetic order follows fd_coeffs.h / exact.h, so results are bit-identical
// to the reference's dense route.
#include <cfloat>
#include <cstdio>

#include "kernels.h"

namespace cnv {

// ---------------------------------------------------------------------------------------------
__global__ void k_apply(const double *__restrict__ A, int nrows, int ncols, int lda, int axis, FdTable t,
                        double *__restrict__ out, int ldo, double scale)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nrows || j >= ncols) return;
    double d;
    if (axis == 1) {
        const double *row = A + (size_t)i * lda;
        d = fd_apply(t, j, [&](int c) { return row[c]; });
    } else {
        const double *col = A + j;
        d = fd_apply(t, i, [&](int r) { return col[(size_t)r * lda]; });
    }
    out[(size_t)i * ldo + j] = scale == 1.0 ? d : xmul(scale, d);  // scale = -1: exact negation
}

// ---------------------------------------------------------------------------------------------
struct BcValues { double u1, u2, u3, u4, v1, v2, v3, v4; };

// value of u / v after the reference's BC loops (rows first, then columns => corners take the
// column values u1/u2, v1/v2; src/main.c:283-296)
__device__ __forceinline__ double bc_u(const double *u, int ld, int nrows, int ncols, const BcValues &b, int i, int j)
{
    if (j == 0) return b.u1;
    if (j == ncols - 1) return b.u2;
    if (i == 0) return b.u3;
    if (i == nrows - 1) return b.u4;
    return u[(size_t)i * ld + j];
}
__device__ __forceinline__ double bc_v(const double *v, int ld, int nrows, int ncols, const BcValues &b, int i, int j)
{
    if (j == 0) return b.v1;
    if (j == ncols - 1) return b.v2;
    if (i == 0) return b.v3;
    if (i == nrows - 1) return b.v4;
    return v[(size_t)i * ld + j];
}

// one thread per ring cell: [0,ncols) bottom row, [ncols,2ncols) top row, then the two columns
__global__ void k_ring_bc_vorticity(double *__restrict__ u, double *__restrict__ v, double *__restrict__ w, int nrows,
                                    int ncols, int ld, BcValues b, FdTable d1x, FdTable d1y)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (idx < ncols) { i = 0; j = idx; }
    else if (idx < 2 * ncols) { i = nrows - 1; j = idx - ncols; }
    else if (idx < 2 * ncols + nrows) { i = idx - 2 * ncols; j = 0; }
    else if (idx < 2 * ncols + 2 * nrows) { i = idx - 2 * ncols - nrows; j = ncols - 1; }
    else return;
    const double dvdx = fd_apply(d1x, j, [&](int c) { return bc_v(v, ld, nrows, ncols, b, i, c); });
    const double dudy = fd_apply(d1y, i, [&](int r) { return bc_u(u, ld, nrows, ncols, b, r, j); });
    const size_t p = (size_t)i * ld + j;
    w[p] = xsub(dvdx, dudy);
    u[p] = bc_u(u, ld, nrows, ncols, b, i, j);
    v[p] = bc_v(v, ld, nrows, ncols, b, i, j);
}

// ---------------------------------------------------------------------------------------------
// w_new = (-u*dwdx - v*dwdy + (1/Re)*(d2wdx2 + d2wdy2))*dt + w on ALL points (ring included),
// src/fluiddyn.c:87.  Also emits rhs = pscale * (-w_new): the reference flips the sign of w in place
// around the Poisson call (invsig, src/main.c:348,363) and the solver multiplies f by
// dx*dx*dy*dy (src/poisson.c:246); the product is rounded once either way.
//
// Tile: 32 x 8 outputs per CTA staged in shared memory with a halo of 3 (order 6), so every w value
// is read from HBM/L2 once per tile instead of 13 times.
constexpr int ETX = 32, ETY = 8, EH = 3;

__global__ void __launch_bounds__(ETX *ETY)
k_euler_fused(const double *__restrict__ w, const double *__restrict__ u, const double *__restrict__ v, int nrows, int ncols,
              int ld, FdTable d1x, FdTable d1y, FdTable d2x, FdTable d2y, double inv_re, double dt, double pscale,
              double *__restrict__ w_new, double *__restrict__ rhs)
{
    __shared__ double tile[ETY + 2 * EH][ETX + 2 * EH + 1];
    const int j0 = blockIdx.x * ETX, i0 = blockIdx.y * ETY;
    for (int t = threadIdx.y * ETX + threadIdx.x; t < (ETY + 2 * EH) * (ETX + 2 * EH); t += ETX * ETY) {
        const int li = t / (ETX + 2 * EH), lj = t - li * (ETX + 2 * EH);
        const int gi = i0 - EH + li, gj = j0 - EH + lj;
        tile[li][lj] = (gi >= 0 && gi < nrows && gj >= 0 && gj < ncols) ? w[(size_t)gi * ld + gj] : 0.0;
    }
    __syncthreads();
    const int j = j0 + threadIdx.x, i = i0 + threadIdx.y;
    if (i >= nrows || j >= ncols) return;
    // closure rows near the walls reach up to 4 cells away from the wall: outside the tile halo only
    // when the cell itself is within 3 cells of the wall, where the tile still covers the operands
    // if they lie inside [i0-3, i0+ETY+3) -- otherwise read global memory.
    auto wx = [&](int c) {
        const int lj = c - (j0 - EH);
        return (lj >= 0 && lj < ETX + 2 * EH) ? tile[threadIdx.y + EH][lj] : w[(size_t)i * ld + c];
    };
    auto wy = [&](int r) {
        const int li = r - (i0 - EH);
        return (li >= 0 && li < ETY + 2 * EH) ? tile[li][threadIdx.x + EH] : w[(size_t)r * ld + j];
    };
    const double dwdx = fd_apply(d1x, j, wx);
    const double dwdy = fd_apply(d1y, i, wy);
    const double d2wdx2 = fd_apply(d2x, j, wx);
    const double d2wdy2 = fd_apply(d2y, i, wy);
    const size_t p = (size_t)i * ld + j;
    const double uu = u[p], vv = v[p], w0 = tile[threadIdx.y + EH][threadIdx.x + EH];
    // (((-u)*dwdx - v*dwdy) + (1/Re)*(d2wdx2+d2wdy2)) * dt + w
    const double conv = xsub(xmul(-uu, dwdx), xmul(vv, dwdy));
    const double diff = xmul(inv_re, xadd(d2wdx2, d2wdy2));
    const double wn = xadd(xmul(xadd(conv, diff), dt), w0);
    w_new[p] = wn;
    if (rhs) rhs[p] = xmul(pscale, -wn);
}

// pointwise form behind the drop-in euler() signature (derivatives supplied by the caller)
__global__ void k_euler_pointwise(double *__restrict__ w, const double *__restrict__ dwdx, const double *__restrict__ dwdy,
                                  const double *__restrict__ d2wdx2, const double *__restrict__ d2wdy2,
                                  const double *__restrict__ u, const double *__restrict__ v, size_t n, double inv_re, double dt)
{
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const double conv = xsub(xmul(-u[p], dwdx[p]), xmul(v[p], dwdy[p]));
        const double diff = xmul(inv_re, xadd(d2wdx2[p], d2wdy2[p]));
        w[p] = xadd(xmul(xadd(conv, diff), dt), w[p]);
    }
}
// out = a + b (continuity, src/fluiddyn.c:141) or b - a (vorticity, :194)
__global__ void k_pointwise_addsub(const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out,
                                   size_t n, int sub)
{
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        out[p] = sub ? xsub(b[p], a[p]) : xadd(a[p], b[p]);
}

// ---------------------------------------------------------------------------------------------
// u = DY psi ; v = -(DX psi) on all points, ring included (src/main.c:366-383)
__global__ void k_velocity(const double *__restrict__ psi, int nrows, int ncols, int ldp, FdTable d1x, FdTable d1y,
                           double *__restrict__ u, double *__restrict__ v, int ld)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nrows || j >= ncols) return;
    const double *row = psi + (size_t)i * ldp;
    const double *col = psi + j;
    const double dpdx = fd_apply(d1x, j, [&](int c) { return row[c]; });
    const double dpdy = fd_apply(d1y, i, [&](int r) { return col[(size_t)r * ldp]; });
    u[(size_t)i * ld + j] = dpdy;
    v[(size_t)i * ld + j] = -dpdx;
}

// ---------------------------------------------------------------------------------------------
// continuity diagnostic: max and min over the grid of DX u + DY v (src/main.c:387-408).
// Persistent grid (a few CTAs per SM) looping over 32x8 tiles; block partials, then the last block
// reduces them in parallel.  max/min are order independent, so the result is deterministic.
__global__ void __launch_bounds__(256)
k_continuity(const double *__restrict__ u, const double *__restrict__ v, int nrows, int ncols, int ld,
             FdTable d1x, FdTable d1y, double *__restrict__ partial, unsigned *__restrict__ ticket,
             double *__restrict__ result)
{
    double mx = -DBL_MAX, mn = DBL_MAX;  // maxel/minel start values, src/linearalg.c:478,514
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_x = (ncols + 31) / 32, tiles_y = (nrows + 7) / 8;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int j = (tile % tiles_x) * 32 + tx, i = (tile / tiles_x) * 8 + ty;
        if (i < nrows && j < ncols) {
            const double *row = u + (size_t)i * ld;
            const double *col = v + j;
            const double dudx = fd_apply(d1x, j, [&](int c) { return row[c]; });
            const double dvdy = fd_apply(d1y, i, [&](int r) { return col[(size_t)r * ld]; });
            const double c = xadd(dudx, dvdy);
            mx = fmax(mx, c);
            mn = fmin(mn, c);
        }
    }
    __shared__ double smx[8], smn[8];
    __shared__ bool last;
    auto block_reduce = [&]() {
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        __syncthreads();
        if (tx == 0) { smx[ty] = mx; smn[ty] = mn; }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 1; k < 8; k++) { mx = fmax(mx, smx[k]); mn = fmin(mn, smn[k]); }
    };
    block_reduce();
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = mx;
        partial[2 * blockIdx.x + 1] = mn;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    mx = -DBL_MAX; mn = DBL_MAX;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
        mx = fmax(mx, __ldcg(&partial[2 * k]));
        mn = fmin(mn, __ldcg(&partial[2 * k + 1]));
    }
    block_reduce();
    if (threadIdx.x == 0) {
        result[0] = mx;
        result[1] = mn;
        *ticket = 0;
    }
}

// rhs = pscale * (sign * f), psi0 = 0: prologue of a stand-alone Poisson solve
__global__ void k_prep_rhs(const double *__restrict__ f, int nrows, int ncols, int ldf, double sign, double pscale,
                           double *__restrict__ rhs, double *__restrict__ psi0, double *__restrict__ psi1, int ld)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nrows || j >= ld) return;
    const size_t p = (size_t)i * ld + j;
    if (f) rhs[p] = j < ncols ? xmul(pscale, sign < 0 ? -f[(size_t)i * ldf + j] : f[(size_t)i * ldf + j]) : 0.0;
    if (psi0) psi0[p] = 0.0;
    if (psi1) psi1[p] = 0.0;
}

// ---- launchers --------------------------------------------------------------------------------
static inline dim3 grid2d(int nrows, int ncols, dim3 b) { return dim3((ncols + b.x - 1) / b.x, (nrows + b.y - 1) / b.y); }

void launch_apply(const double *A, int nrows, int ncols, int lda, int axis, const FdTable &t, double *out, int ldo,
                  double scale, cudaStream_t s)
{
    dim3 b(32, 8);
    k_apply<<<grid2d(nrows, ncols, b), b, 0, s>>>(A, nrows, ncols, lda, axis, t, out, ldo, scale);
}
void launch_ring_bc_vorticity(double *u, double *v, double *w, int nrows, int ncols, int ld, const double bc[8],
                              const FdTable &d1x, const FdTable &d1y, cudaStream_t s)
{
    BcValues b{bc[0], bc[1], bc[2], bc[3], bc[4], bc[5], bc[6], bc[7]};
    const int n = 2 * ncols + 2 * nrows;
    k_ring_bc_vorticity<<<(n + 127) / 128, 128, 0, s>>>(u, v, w, nrows, ncols, ld, b, d1x, d1y);
}
void launch_euler_fused(const double *w, const double *u, const double *v, int nrows, int ncols, int ld, const FdTable &d1x,
                        const FdTable &d1y, const FdTable &d2x, const FdTable &d2y, double inv_re, double dt, double pscale,
                        double *w_new, double *rhs, cudaStream_t s)
{
    dim3 b(ETX, ETY);
    k_euler_fused<<<grid2d(nrows, ncols, b), b, 0, s>>>(w, u, v, nrows, ncols, ld, d1x, d1y, d2x, d2y, inv_re, dt, pscale,
                                                        w_new, rhs);
}
void launch_euler_pointwise(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
                            const double *u, const double *v, size_t n, double inv_re, double dt, cudaStream_t s)
{
    k_euler_pointwise<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(w, dwdx, dwdy, d2wdx2, d2wdy2,
                                                                                                  u, v, n, inv_re, dt);
}
void launch_pointwise_addsub(const double *a, const double *b, double *out, size_t n, int sub, cudaStream_t s)
{
    k_pointwise_addsub<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(a, b, out, n, sub);
}
void launch_velocity(const double *psi, int nrows, int ncols, int ldp, const FdTable &d1x, const FdTable &d1y, double *u,
                     double *v, int ld, cudaStream_t s)
{
    dim3 b(32, 8);
    k_velocity<<<grid2d(nrows, ncols, b), b, 0, s>>>(psi, nrows, ncols, ldp, d1x, d1y, u, v, ld);
}
int continuity_blocks(int nrows, int ncols)
{
    const int tiles = ((ncols + 31) / 32) * ((nrows + 7) / 8);
    return tiles < 148 * 8 ? tiles : 148 * 8;
}
void launch_continuity(const double *u, const double *v, int nrows, int ncols, int ld, const FdTable &d1x, const FdTable &d1y,
                       double *partial, unsigned *ticket, double *result, cudaStream_t s)
{
    k_continuity<<<continuity_blocks(nrows, ncols), 256, 0, s>>>(u, v, nrows, ncols, ld, d1x, d1y, partial, ticket, result);
}
void launch_prep_rhs(const double *f, int nrows, int ncols, int ldf, double sign, double pscale, double *rhs, double *psi0,
                     double *psi1, int ld, cudaStream_t s)
{
    dim3 b(32, 8);
    k_prep_rhs<<<grid2d(nrows, ld, b), b, 0, s>>>(f, nrows, ncols, ldf, sign, pscale, rhs, psi0, psi1, ld);
}

}  // namespace cnv
