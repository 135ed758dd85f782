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
#include <type_traits>

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
// column values u1/u2, v1/v2; src/main.c:283-296).  gi = GLOBAL row; the arrays are slab-local.
__device__ __forceinline__ double bc_u(const double *u, int ld, const RowMap &m, int ncols, const BcValues &b, int gi, int j)
{
    if (j == 0) return b.u1;
    if (j == ncols - 1) return b.u2;
    if (gi == 0) return b.u3;
    if (gi == m.gnrows - 1) return b.u4;
    return u[(size_t)(gi - m.grow0) * ld + j];
}
__device__ __forceinline__ double bc_v(const double *v, int ld, const RowMap &m, int ncols, const BcValues &b, int gi, int j)
{
    if (j == 0) return b.v1;
    if (j == ncols - 1) return b.v2;
    if (gi == 0) return b.v3;
    if (gi == m.gnrows - 1) return b.v4;
    return v[(size_t)(gi - m.grow0) * ld + j];
}

// one thread per ring cell of the GLOBAL grid: [0,ncols) bottom row, [ncols,2ncols) top row, then the two
// columns; a slab handles the cells whose row it owns
__global__ void k_ring_bc_vorticity(double *__restrict__ u, double *__restrict__ v, double *__restrict__ w, RowMap m,
                                    int ncols, int ld, BcValues b, FdTable d1x, FdTable d1y)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrows = m.gnrows;
    int i, j;
    if (idx < ncols) { i = 0; j = idx; }
    else if (idx < 2 * ncols) { i = nrows - 1; j = idx - ncols; }
    else if (idx < 2 * ncols + nrows) { i = idx - 2 * ncols; j = 0; }
    else if (idx < 2 * ncols + 2 * nrows) { i = idx - 2 * ncols - nrows; j = ncols - 1; }
    else return;
    const int li = i - m.grow0;
    if (li < m.own_lo || li >= m.own_hi) return;
    const double dvdx = fd_apply(d1x, j, [&](int c) { return bc_v(v, ld, m, ncols, b, i, c); });
    const double dudy = fd_apply(d1y, i, [&](int r) { return bc_u(u, ld, m, ncols, b, r, j); });
    const size_t p = (size_t)li * ld + j;
    w[p] = xsub(dvdx, dudy);
    u[p] = bc_u(u, ld, m, ncols, b, i, j);
    v[p] = bc_v(v, ld, m, ncols, b, i, j);
}

// ---------------------------------------------------------------------------------------------
// Tiled stencil machinery.  A CTA of 256 threads (64 x 4) owns a 64 x 32 tile; the operand field is
// staged once in shared memory with a halo of 3 (order 6 reach), so HBM sees each value ~1.3 times.
// Thread (tx, ty) walks 8 rows of column tx.  Interior cells (>= half cells away from every wall, the
// overwhelming majority) use the unrolled interior row with its coefficients in registers; cells
// within `half` of a wall take the generic closure rows (fd_apply).  Both accumulate in ascending
// column order from +0.0 with separately rounded products (fd_coeffs.h), so results do not depend on
// the path taken.
constexpr int TW = 64, TH = 32, THALO = 3, TROWS = TH / 4;
constexpr int TPITCH = TW + 2 * THALO + 1;

struct Tile {
    double v[TH + 2 * THALO][TPITCH];
};

__device__ __forceinline__ bool tile_inside(int i0, int j0, int nrows, int ncols);
// Each of the 8 warps takes tile rows warp, warp+8, ...; a lane takes columns lane, lane+32, lane+64.  All
// (up to 15) global loads of a thread are issued before the first shared-memory store, so they are in flight
// together (the loop is fully unrolled; out-of-domain elements read as 0 without touching memory).
// A tile whose halo lies entirely inside the local array (uniform per CTA; all but the perimeter tiles) takes a path
// without per-element bounds tests: the general path spends 4 compares + address arithmetic per element, 13 instructions
// per element in SASS, which was ~40 % of all instructions k_continuity issued (profiles/ncu_stencil_r2.md).
__device__ __forceinline__ void tile_load(Tile &t, const double *__restrict__ A, int i0, int j0, int nrows, int ncols, int ld)
{
    constexpr int NR = (TH + 2 * THALO + 7) / 8, NC = (TW + 2 * THALO + 31) / 32;
    constexpr int LASTR = TH + 2 * THALO - 8 * (NR - 1), LASTC = TW + 2 * THALO - 32 * (NC - 1);  // warps / lanes of the last round
    const int lane = threadIdx.x & 31, warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
    double buf[NR][NC];
    if (tile_inside(i0, j0, nrows, ncols)) {
        const double *base = A + (size_t)(i0 - THALO + warp) * ld + (j0 - THALO + lane);
#pragma unroll
        for (int a = 0; a < NR; a++) {
            const double *row = base + (size_t)(8 * a) * ld;
#pragma unroll
            for (int b = 0; b < NC; b++) {
                const bool ok = (a < NR - 1 || warp < LASTR) && (b < NC - 1 || lane < LASTC);
                buf[a][b] = ok ? row[32 * b] : 0.0;
            }
        }
    } else {
#pragma unroll
        for (int a = 0; a < NR; a++) {
            const int li = warp + 8 * a, gi = i0 - THALO + li;
            const bool rowok = li < TH + 2 * THALO && gi >= 0 && gi < nrows;
            const double *row = A + (size_t)(rowok ? gi : 0) * ld + (j0 - THALO);
#pragma unroll
            for (int b = 0; b < NC; b++) {
                const int lj = lane + 32 * b, gj = j0 - THALO + lj;
                buf[a][b] = (rowok && lj < TW + 2 * THALO && gj >= 0 && gj < ncols) ? row[lj] : 0.0;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NR; a++) {
#pragma unroll
        for (int b = 0; b < NC; b++) {
            if ((a < NR - 1 || warp < LASTR) && (b < NC - 1 || lane < LASTC)) t.v[warp + 8 * a][lane + 32 * b] = buf[a][b];
        }
    }
}

// true when the tile's halo lies entirely inside the local array
__device__ __forceinline__ bool tile_inside(int i0, int j0, int nrows, int ncols)
{
    return i0 >= THALO && i0 + TH + THALO <= nrows && j0 >= THALO && j0 + TW + THALO <= ncols;
}

// interior row: sum_k c[k] * x(k - HALF), ascending
template <int HALF, class Load>
__device__ __forceinline__ double fd_interior(const double (&c)[7], Load x)
{
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k <= 2 * HALF; k++) sum = xadd(sum, xmul(c[k], x(k - HALF)));
    return sum;
}

// Register windows: win[0..6] = the operand at offsets -3..+3 along the differentiated axis.
// Interior cells (>= HALF away from both walls) use the unrolled interior row on the window; cells within
// HALF of a wall take the generic closure rows (fd_apply) reading the tile.
// ZC: the centre coefficient is exactly 0 (interior rows of Diff1, src/finitediff.c:72-145 never set it).  Its term 0 * x = +-0
// cannot change the running sum: that sum starts at +0.0 and (+0) + (-0) = +0, so it is never -0, and s + (+-0) = s for
// every s that is not -0.  Skipping the tap is therefore bit-exact (finite operands) and saves 2 of 14 fp64 operations.
template <int HALF, bool ZC = false>
__device__ __forceinline__ double win_deriv(const double (&c)[7], const double (&win)[7])
{
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k <= 2 * HALF; k++) {
        if (ZC && k == HALF) continue;
        sum = xadd(sum, xmul(c[k], win[THALO - HALF + k]));
    }
    return sum;
}
template <int HALF, bool INTERIOR, bool ZC = false>
__device__ __forceinline__ double tile_dx(const Tile &t, const FdTable &tab, const double (&c)[7], const double (&win)[7],
                                          int li, int j, int j0)
{
    if (INTERIOR || (j >= HALF && j < tab.n - HALF)) return win_deriv<HALF, ZC>(c, win);
    return fd_apply(tab, j, [&](int col) { return t.v[li][col - j0 + THALO]; });
}
// i = GLOBAL row of the cell, gi0 = global row of the tile's first row
template <int HALF, bool INTERIOR, bool ZC = false>
__device__ __forceinline__ double tile_dy(const Tile &t, const FdTable &tab, const double (&c)[7], const double (&win)[7],
                                          int lj, int i, int gi0)
{
    if (INTERIOR || (i >= HALF && i < tab.n - HALF)) return win_deriv<HALF, ZC>(c, win);
    return fd_apply(tab, i, [&](int row) { return t.v[row - gi0 + THALO][lj]; });
}
// a tile whose cells are all >= 3 away from every wall needs no closure rows at all (uniform per CTA):
// the kernels run a branch-free body for those tiles and the general body for the perimeter tiles
__device__ __forceinline__ bool tile_is_interior(int i0, int j0, const RowMap &m, int ncols)
{
    const int gi0 = m.grow0 + i0;
    return gi0 >= THALO && j0 >= THALO && gi0 + TH <= m.gnrows - THALO && i0 + TH <= m.own_hi && j0 + TW <= ncols - THALO;
}
// x window of row li around column lj; y window start (rows li-3..li+3) and slide by one row
__device__ __forceinline__ void win_x(const Tile &t, int li, int lj, double (&w)[7])
{
#pragma unroll
    for (int d = 0; d < 7; d++) w[d] = t.v[li][lj - THALO + d];
}
__device__ __forceinline__ void win_y_init(const Tile &t, int li, int lj, double (&w)[7])
{
#pragma unroll
    for (int d = 0; d < 7; d++) w[d] = t.v[li - THALO + d][lj];
}
__device__ __forceinline__ void win_y_slide(const Tile &t, int li, int lj, double (&w)[7])
{
#pragma unroll
    for (int d = 0; d < 6; d++) w[d] = w[d + 1];
    w[6] = t.v[li + THALO][lj];
}
__device__ __forceinline__ void load_coefs(const FdTable &t, double (&c)[7])
{
#pragma unroll
    for (int k = 0; k < 7; k++) c[k] = t.interior.c[k];
}

// w_new = (-u*dwdx - v*dwdy + (1/Re)*(d2wdx2 + d2wdy2))*dt + w on ALL points (ring included),
// src/fluiddyn.c:87.  Also emits rhs = pscale * (-w_new): the reference flips the sign of w in place
// around the Poisson call (invsig, src/main.c:348,363) and the solver multiplies f by
// dx*dx*dy*dy (src/poisson.c:246); the product is rounded once either way.
template <int HALF>
__global__ void __launch_bounds__(256)
k_euler_fused(const double *__restrict__ w, const double *__restrict__ u, const double *__restrict__ v, RowMap m, int ncols,
              int ld, const FdTable d1x, const FdTable d1y, const FdTable d2x, const FdTable d2y, double inv_re, double dt,
              double pscale, double *__restrict__ w_new, double *__restrict__ rhs)
{
    __shared__ Tile tw;
    const int j0 = blockIdx.x * TW, i0 = m.own_lo + blockIdx.y * TH, gi0 = m.grow0 + i0;
    tile_load(tw, w, i0, j0, m.nloc, ncols, ld);
    double c1x[7], c1y[7], c2x[7], c2y[7];
    load_coefs(d1x, c1x); load_coefs(d1y, c1y); load_coefs(d2x, c2x); load_coefs(d2y, c2y);
    __syncthreads();
    const int j = j0 + threadIdx.x, lj = threadIdx.x + THALO;
    if (j >= ncols) return;
    auto body = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        double wy[7], wx[7];
        win_y_init(tw, threadIdx.y * TROWS + THALO, lj, wy);
#pragma unroll
        for (int rr = 0; rr < TROWS; rr++) {
            const int i = i0 + threadIdx.y * TROWS + rr, li = threadIdx.y * TROWS + rr + THALO;
            if (!INTERIOR && i >= m.own_hi) break;
            if (rr > 0) win_y_slide(tw, li, lj, wy);
            win_x(tw, li, lj, wx);
            const double dwdx = tile_dx<HALF, INTERIOR, true>(tw, d1x, c1x, wx, li, j, j0);
            const double dwdy = tile_dy<HALF, INTERIOR, true>(tw, d1y, c1y, wy, lj, m.grow0 + i, gi0);
            const double d2wdx2 = tile_dx<HALF, INTERIOR>(tw, d2x, c2x, wx, li, j, j0);
            const double d2wdy2 = tile_dy<HALF, INTERIOR>(tw, d2y, c2y, wy, lj, m.grow0 + i, gi0);
            const size_t p = (size_t)i * ld + j;
            const double uu = u[p], vv = v[p], w0 = wx[THALO];
            // (((-u)*dwdx - v*dwdy) + (1/Re)*(d2wdx2+d2wdy2)) * dt + w
            const double conv = xsub(xmul(-uu, dwdx), xmul(vv, dwdy));
            const double diff = xmul(inv_re, xadd(d2wdx2, d2wdy2));
            const double wn = xadd(xmul(xadd(conv, diff), dt), w0);
            w_new[p] = wn;
            if (rhs) rhs[p] = xmul(pscale, -wn);
        }
    };
    if (tile_is_interior(i0, j0, m, ncols)) body(std::true_type{});
    else body(std::false_type{});
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
template <int HALF>
__global__ void __launch_bounds__(256)
k_velocity(const double *__restrict__ psi, RowMap m, int ncols, int ldp, const FdTable d1x, const FdTable d1y,
           double *__restrict__ u, double *__restrict__ v, int ld)
{
    __shared__ Tile tp;
    const int j0 = blockIdx.x * TW, i0 = m.own_lo + blockIdx.y * TH, gi0 = m.grow0 + i0;
    tile_load(tp, psi, i0, j0, m.nloc, ncols, ldp);
    double c1x[7], c1y[7];
    load_coefs(d1x, c1x); load_coefs(d1y, c1y);
    __syncthreads();
    const int j = j0 + threadIdx.x, lj = threadIdx.x + THALO;
    if (j >= ncols) return;
    auto body = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        double wy[7], wx[7];
        win_y_init(tp, threadIdx.y * TROWS + THALO, lj, wy);
#pragma unroll
        for (int rr = 0; rr < TROWS; rr++) {
            const int i = i0 + threadIdx.y * TROWS + rr, li = threadIdx.y * TROWS + rr + THALO;
            if (!INTERIOR && i >= m.own_hi) break;
            if (rr > 0) win_y_slide(tp, li, lj, wy);
            win_x(tp, li, lj, wx);
            const double dpdx = tile_dx<HALF, INTERIOR, true>(tp, d1x, c1x, wx, li, j, j0);
            const double dpdy = tile_dy<HALF, INTERIOR, true>(tp, d1y, c1y, wy, lj, m.grow0 + i, gi0);
            u[(size_t)i * ld + j] = dpdy;
            v[(size_t)i * ld + j] = -dpdx;
        }
    };
    if (tile_is_interior(i0, j0, m, ncols)) body(std::true_type{});
    else body(std::false_type{});
}

// ---------------------------------------------------------------------------------------------
// continuity diagnostic: max and min over the grid of DX u + DY v (src/main.c:387-408).
// One 64 x 32 tile per CTA (u and v staged with halos); block partials, then the last block reduces them
// in parallel.  max/min are order independent, so the result is deterministic.
// per-thread max / min of DX u + DY v over the thread's cells of the tile, u and v staged in tu / tv
template <int HALF>
__device__ __forceinline__ void continuity_from_tiles(const Tile &tu, const Tile &tv, const RowMap &m, int ncols, const FdTable &d1x,
                                                      const FdTable &d1y, const double (&c1x)[7], const double (&c1y)[7], int i0, int j0,
                                                      double &mx, double &mn)
{
    const int gi0 = m.grow0 + i0;
    const int j = j0 + threadIdx.x, lj = threadIdx.x + THALO;
    auto body = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        double wy[7], wx[7];
        win_y_init(tv, threadIdx.y * TROWS + THALO, lj, wy);
#pragma unroll
        for (int rr = 0; rr < TROWS; rr++) {
            const int i = i0 + threadIdx.y * TROWS + rr, li = threadIdx.y * TROWS + rr + THALO;
            if (!INTERIOR && i >= m.own_hi) break;
            if (rr > 0) win_y_slide(tv, li, lj, wy);
            win_x(tu, li, lj, wx);
            const double dudx = tile_dx<HALF, INTERIOR, true>(tu, d1x, c1x, wx, li, j, j0);
            const double dvdy = tile_dy<HALF, INTERIOR, true>(tv, d1y, c1y, wy, lj, m.grow0 + i, gi0);
            const double c = xadd(dudx, dvdy);
            mx = fmax(mx, c);
            mn = fmin(mn, c);
        }
    };
    if (j < ncols) {
        if (tile_is_interior(i0, j0, m, ncols)) body(std::true_type{});
        else body(std::false_type{});
    }
}

// block max/min -> per-block partials -> the last block (atomic ticket) reduces them in parallel.  max/min are
// order independent, so the result is deterministic.
__device__ __forceinline__ void minmax_finish(double mx, double mn, double *__restrict__ partial, unsigned *__restrict__ ticket,
                                              double *__restrict__ result)
{
    __shared__ double smx[8], smn[8];
    __shared__ bool last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31, wid = tid >> 5;
    auto block_reduce = [&]() {
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        __syncthreads();
        if (lane == 0) { smx[wid] = mx; smn[wid] = mn; }
        __syncthreads();
        if (tid == 0)
            for (int k = 1; k < 8; k++) { mx = fmax(mx, smx[k]); mn = fmin(mn, smn[k]); }
    };
    block_reduce();
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        partial[2 * bid] = mx;
        partial[2 * bid + 1] = mn;
        __threadfence();
        last = atomicAdd(ticket, 1u) == (unsigned)nblocks - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    mx = -DBL_MAX; mn = DBL_MAX;
    for (int k = tid; k < nblocks; k += 256) {
        mx = fmax(mx, __ldcg(&partial[2 * k]));
        mn = fmin(mn, __ldcg(&partial[2 * k + 1]));
    }
    block_reduce();
    if (tid == 0) {
        result[0] = mx;
        result[1] = mn;
        *ticket = 0;
    }
}

// Persistent: a CTA walks the tiles bid, bid + grid, ... and keeps its max / min in registers, so the block reduction, the
// fence and the ticket of minmax_finish are paid once per CTA, not once per 64 x 32 tile.
// Interior tiles (no closure row, halo inside the array: all but the perimeter) take a lean path: DX u needs neighbours
// along x only, DY v along y only, so u is staged WITHOUT its y halo (32 x 70 instead of 38 x 70 values) and v is not
// staged at all -- a thread walks 8 rows of one column, so its 14 values of v come straight from global memory into the
// register window (coalesced across the warp, issued before the barrier so that they overlap the staging of u).  Per thread
// and tile that is 12 + 14 loads, 12 shared stores and 48 shared loads instead of 30 / 30 / 62.  Same operations on the same
// values in the same order as the general path, so the bits do not depend on the path.
struct TileX {
    double v[TH][TW + 2 * THALO];
};
static_assert(sizeof(TileX) <= sizeof(Tile), "TileX aliases the u tile");

template <int HALF>
__global__ void __launch_bounds__(256)
k_continuity(const double *__restrict__ u, const double *__restrict__ v, RowMap m, int ncols, int ld,
             const FdTable d1x, const FdTable d1y, double *__restrict__ partial, unsigned *__restrict__ ticket,
             double *__restrict__ result, const int gx, const int ntiles)
{
    __shared__ Tile tu, tv;
    double mx = -DBL_MAX, mn = DBL_MAX;  // maxel/minel start values, src/linearalg.c:478,514
    double c1x[7], c1y[7];
    load_coefs(d1x, c1x); load_coefs(d1y, c1y);
    const int lane = threadIdx.x & 31, warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j0 = (tile % gx) * TW, i0 = m.own_lo + (tile / gx) * TH;
        __syncthreads();  // everybody has finished reading the previous tile
        if (tile_is_interior(i0, j0, m, ncols) && tile_inside(i0, j0, m.nloc, ncols)) {
            TileX &tx = reinterpret_cast<TileX &>(tu);
            constexpr int NC = (TW + 2 * THALO + 31) / 32, LASTC = TW + 2 * THALO - 32 * (NC - 1);
            double win[TROWS + 2 * THALO], ub[TH / 8][NC];
            const double *vp = v + (size_t)(i0 + threadIdx.y * TROWS - THALO) * ld + (j0 + threadIdx.x);
#pragma unroll
            for (int k = 0; k < TROWS + 2 * THALO; k++) win[k] = vp[(size_t)k * ld];
            const double *up = u + (size_t)(i0 + warp) * ld + (j0 - THALO + lane);
#pragma unroll
            for (int a = 0; a < TH / 8; a++)
#pragma unroll
                for (int b = 0; b < NC; b++) ub[a][b] = (b < NC - 1 || lane < LASTC) ? up[(size_t)(8 * a) * ld + 32 * b] : 0.0;
#pragma unroll
            for (int a = 0; a < TH / 8; a++)
#pragma unroll
                for (int b = 0; b < NC; b++)
                    if (b < NC - 1 || lane < LASTC) tx.v[warp + 8 * a][lane + 32 * b] = ub[a][b];
            __syncthreads();
#pragma unroll
            for (int rr = 0; rr < TROWS; rr++) {
                const int li = threadIdx.y * TROWS + rr;
                double wx[7], wy[7];
#pragma unroll
                for (int d = 0; d < 7; d++) { wx[d] = tx.v[li][threadIdx.x + d]; wy[d] = win[rr + d]; }
                const double c = xadd(win_deriv<HALF, true>(c1x, wx), win_deriv<HALF, true>(c1y, wy));
                mx = fmax(mx, c);
                mn = fmin(mn, c);
            }
        } else {
            tile_load(tu, u, i0, j0, m.nloc, ncols, ld);
            tile_load(tv, v, i0, j0, m.nloc, ncols, ld);
            __syncthreads();
            continuity_from_tiles<HALF>(tu, tv, m, ncols, d1x, d1y, c1x, c1y, i0, j0, mx, mn);
        }
    }
    minmax_finish(mx, mn, partial, ticket, result);
}

// ---------------------------------------------------------------------------------------------
// Pressure-Poisson right-hand side, the recipe the reference leaves commented out (src/main.c:421-427, declared
// as pressure() in include/fluiddyn.h:11 but never defined):
//     f = dudx**2 + dvdy**2 + 2*dudy*dvdx          p = poisson(-f)
// evaluated left to right with separately rounded operations: ((dudx*dudx) + (dvdy*dvdy)) + ((2*dudy)*dvdx).
// One 64 x 32 tile of u and v per CTA; the four first derivatives use the same closures as everywhere else.
// Writes f (optional) and the solver's pre-scaled right-hand side pscale * (-f) (optional).
template <int HALF>
__global__ void __launch_bounds__(256)
k_pressure_rhs(const double *__restrict__ u, const double *__restrict__ v, RowMap m, int ncols, int ld, const FdTable d1x,
               const FdTable d1y, double pscale, double *__restrict__ f_out, double *__restrict__ rhs, int ldo)
{
    __shared__ Tile tu, tv;
    const int j0 = blockIdx.x * TW, i0 = m.own_lo + blockIdx.y * TH, gi0 = m.grow0 + i0;
    tile_load(tu, u, i0, j0, m.nloc, ncols, ld);
    tile_load(tv, v, i0, j0, m.nloc, ncols, ld);
    double c1x[7], c1y[7];
    load_coefs(d1x, c1x); load_coefs(d1y, c1y);
    __syncthreads();
    const int j = j0 + threadIdx.x, lj = threadIdx.x + THALO;
    if (j >= ncols) return;
    auto body = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        double uy[7], vy[7], wx[7];
        win_y_init(tu, threadIdx.y * TROWS + THALO, lj, uy);
        win_y_init(tv, threadIdx.y * TROWS + THALO, lj, vy);
#pragma unroll
        for (int rr = 0; rr < TROWS; rr++) {
            const int i = i0 + threadIdx.y * TROWS + rr, li = threadIdx.y * TROWS + rr + THALO;
            if (!INTERIOR && i >= m.own_hi) break;
            if (rr > 0) { win_y_slide(tu, li, lj, uy); win_y_slide(tv, li, lj, vy); }
            win_x(tu, li, lj, wx);
            const double dudx = tile_dx<HALF, INTERIOR, true>(tu, d1x, c1x, wx, li, j, j0);
            win_x(tv, li, lj, wx);
            const double dvdx = tile_dx<HALF, INTERIOR, true>(tv, d1x, c1x, wx, li, j, j0);
            const double dudy = tile_dy<HALF, INTERIOR, true>(tu, d1y, c1y, uy, lj, m.grow0 + i, gi0);
            const double dvdy = tile_dy<HALF, INTERIOR, true>(tv, d1y, c1y, vy, lj, m.grow0 + i, gi0);
            const double f = xadd(xadd(xmul(dudx, dudx), xmul(dvdy, dvdy)), xmul(xmul(2.0, dudy), dvdx));
            if (f_out) f_out[(size_t)i * ldo + j] = f;
            if (rhs) rhs[(size_t)i * ld + j] = xmul(pscale, -f);
        }
    };
    if (tile_is_interior(i0, j0, m, ncols)) body(std::true_type{});
    else body(std::false_type{});
}

// rhs = pscale * (sign * f), psi0 = 0: prologue of a stand-alone Poisson solve
// (f may alias rhs with ldf == ld: the host-buffer upload copies straight into rhs and scales in place)
__global__ void k_prep_rhs(const double *f, int nrows, int ncols, int ldf, double sign, double pscale,
                           double *rhs, double *__restrict__ psi0, double *__restrict__ psi1, int ld)
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
void launch_ring_bc_vorticity(double *u, double *v, double *w, const RowMap &m, int ncols, int ld, const double bc[8],
                              const FdTable &d1x, const FdTable &d1y, cudaStream_t s)
{
    BcValues b{bc[0], bc[1], bc[2], bc[3], bc[4], bc[5], bc[6], bc[7]};
    const int n = 2 * ncols + 2 * m.gnrows;
    k_ring_bc_vorticity<<<(n + 127) / 128, 128, 0, s>>>(u, v, w, m, ncols, ld, b, d1x, d1y);
}
static inline dim3 tile_grid(const RowMap &m, int ncols) { return dim3((ncols + TW - 1) / TW, (m.own_hi - m.own_lo + TH - 1) / TH); }
#define CNV_BY_HALF(half, CALL)             \
    do {                                    \
        if ((half) == 1) { CALL(1); }       \
        else if ((half) == 2) { CALL(2); }  \
        else { CALL(3); }                   \
    } while (0)

void launch_euler_fused(const double *w, const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x,
                        const FdTable &d1y, const FdTable &d2x, const FdTable &d2y, double inv_re, double dt, double pscale,
                        double *w_new, double *rhs, cudaStream_t s)
{
    const dim3 b(TW, 4), g = tile_grid(m, ncols);
#define CALL(H) k_euler_fused<H><<<g, b, 0, s>>>(w, u, v, m, ncols, ld, d1x, d1y, d2x, d2y, inv_re, dt, pscale, w_new, rhs)
    CNV_BY_HALF(d1x.half, CALL);
#undef CALL
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
void launch_velocity(const double *psi, const RowMap &m, int ncols, int ldp, const FdTable &d1x, const FdTable &d1y, double *u,
                     double *v, int ld, cudaStream_t s)
{
    const dim3 b(TW, 4), g = tile_grid(m, ncols);
#define CALL(H) k_velocity<H><<<g, b, 0, s>>>(psi, m, ncols, ldp, d1x, d1y, u, v, ld)
    CNV_BY_HALF(d1x.half, CALL);
#undef CALL
}
void launch_pressure_rhs(const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x, const FdTable &d1y,
                         double pscale, double *f_out, double *rhs, int ldo, cudaStream_t s)
{
    const dim3 b(TW, 4), g = tile_grid(m, ncols);
#define CALL(H) k_pressure_rhs<H><<<g, b, 0, s>>>(u, v, m, ncols, ld, d1x, d1y, pscale, f_out, rhs, ldo)
    CNV_BY_HALF(d1x.half, CALL);
#undef CALL
}

int continuity_blocks(int nrows, int ncols) { return ((ncols + TW - 1) / TW) * ((nrows + TH - 1) / TH); }
void launch_continuity(const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x, const FdTable &d1y,
                       double *partial, unsigned *ticket, double *result, cudaStream_t s)
{
    const dim3 b(TW, 4), g = tile_grid(m, ncols);
    const int ntiles = (int)(g.x * g.y);
    static int slots[kMaxDevices] = {0};  // CTAs the device holds at once (registers and the two staged tiles decide)
    const int dslot = current_device_slot();
    if (!slots[dslot]) {
        int dev = 0, sms = 148, per_sm = 3;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_continuity<3>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
        slots[dslot] = per_sm * sms;
    }
    const int grid = ntiles < slots[dslot] ? ntiles : slots[dslot];
#define CALL(H) k_continuity<H><<<grid, b, 0, s>>>(u, v, m, ncols, ld, d1x, d1y, partial, ticket, result, (int)g.x, ntiles)
    CNV_BY_HALF(d1x.half, CALL);
#undef CALL
}
void launch_prep_rhs(const double *f, int nrows, int ncols, int ldf, double sign, double pscale, double *rhs, double *psi0,
                     double *psi1, int ld, cudaStream_t s)
{
    dim3 b(32, 8);
    k_prep_rhs<<<grid2d(nrows, ld, b), b, 0, s>>>(f, nrows, ncols, ldf, sign, pscale, rhs, psi0, psi1, ld);
}

}  // namespace cnv
