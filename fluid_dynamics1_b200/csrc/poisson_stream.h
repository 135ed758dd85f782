// poisson_stream.h -- temporally blocked red-black SOR: the per-thread phases of the streaming
// kernel, shared between the CUDA kernel (poisson.cu) and the host-side schedule checker
// (tests/emul/stream_emul.cc compiles these same functions with g++ and runs the "threads" of a
// CTA in a loop, so the tiling / ring-buffer / skew logic is tested without a GPU).
//
// Algorithm (restates src/poisson.c:238-262, the OpenMP red-black sweep, T sweeps per HBM pass)
// ---------------------------------------------------------------------------------------------
// One CTA owns a strip of Wout output columns x Hout output rows and streams rows bottom-up
// through a ring buffer of R = 4T+PF+1 rows in shared memory.  Each row lives in column-parity
// split form (SE = even columns, SO = odd columns) so that every access of a half-row update is
// unit stride: for an even-column cell (pair k) N/S are SE[q+-1][k], E/W are SO[q][k], SO[q][k-1].
// The 2T half-sweeps (red L1, black L1, red L2, ..., black LT) are skewed by 2 rows each:
// at step r (row r has just landed) stage sigma updates row r-1-2*sigma.  All inputs of every
// stage were produced in EARLIER steps, so the 2T stages of one step are independent and one
// __syncthreads per step suffices; the update is in place (same dependency structure as the
// reference's in-place sweeps, only the order of independent cell updates changes, so every
// cell sees bit-identical operands).  Row r-4T is final after step r-1 and is written back at
// step r.  Dependencies reach 2 cells per sweep, so strips/chunks overlap by 2T halo cells whose
// (stale-neighbour) results are discarded: the first/last loaded row and column are never
// updated, and the region of influence of that staleness stays inside the halo.
// Thread (g, kk) of the CTA owns level g+1 (its red and black stage) and four adjacent columns
// (pairs 2kk, 2kk+1), so its |u - u0| contributions all belong to sweep g+1: one accumulator.
#pragma once
#include "exact.h"

namespace cnv {

constexpr int kPrefetch = 3;  // rows in flight ahead of the compute front (cp.async groups)
constexpr int ring_rows(int T) { return 4 * T + kPrefetch + 1; }

struct PassGeom {
    // local array: nrows x ncols doubles with pitch ld; row 0 is global row grow0 of gnrows
    int nrows, ncols, ld, grow0, gnrows;
    int own_lo, own_hi;  // local rows [own_lo, own_hi) are produced (stored) by this device
    // tiling
    int WS;    // strip width held in shared memory (multiple of 4)
    int HX;    // x halo (multiple of 4, >= 2T)
    int Wout;  // WS - 2*HX
    int Hout;  // output rows per chunk
    int HY;    // y halo = 2T
    int nstrips, nchunks;
};

struct CtaGeom {
    int gx0;       // global column of shared-memory column 0 (even; may be negative)
    int y0, y1;    // output rows [y0, y1)
    int ylo, yhi;  // streamed rows [ylo, yhi]
};

CNV_HD CtaGeom cta_geom(const PassGeom &p, int bx, int by)
{
    CtaGeom G;
    G.gx0 = bx * p.Wout - p.HX;
    G.y0 = p.own_lo + by * p.Hout;
    G.y1 = G.y0 + p.Hout < p.own_hi ? G.y0 + p.Hout : p.own_hi;
    G.ylo = G.y0 - p.HY > 0 ? G.y0 - p.HY : 0;
    G.yhi = G.y1 - 1 + p.HY < p.nrows - 1 ? G.y1 - 1 + p.HY : p.nrows - 1;
    return G;
}

// shared-memory layout of one ring slot: [SE | SO | PE | PO], each WP+4 doubles (2 pad each side)
CNV_HD int slot_stride(int WS) { return 4 * (WS / 2 + 4); }
CNV_HD int arr_off(int WS, int which) { return which * (WS / 2 + 4) + 2; }  // 0 SE, 1 SO, 2 PE, 3 PO

struct ThreadCtx {
    int g;        // level index (sweep g+1 of the pass)
    int k0;       // first of the two column pairs
    int vmask;    // bit i: smem col 4kk+i is updatable; cols = (E k0, O k0, E k0+1, O k0+1)
    bool colown;  // the four columns belong to the strip's output range
};

CNV_HD ThreadCtx thread_ctx(const PassGeom &p, const CtaGeom &G, int tid)
{
    ThreadCtx t;
    const int TPG = p.WS >> 2;
    t.g = tid / TPG;
    const int kk = tid - t.g * TPG;
    t.k0 = 2 * kk;
    t.vmask = 0;
    for (int i = 0; i < 4; i++) {
        int c = 4 * kk + i, gc = G.gx0 + c;
        if (c >= 1 && c <= p.WS - 2 && gc >= 1 && gc <= p.ncols - 2) t.vmask |= 1 << i;
    }
    t.colown = 4 * kk >= p.HX && 4 * kk < p.HX + p.Wout;
    return t;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
#else
inline void cp_async8(double *dst, const double *src) { *dst = *src; }
#endif

// ---- phase 1: bring local row rl (psi and prepared right-hand side) into its ring slot --------
template <int T>
CNV_HD void phase_load(const PassGeom &p, const CtaGeom &G, double *sm, const double *__restrict__ in,
                       const double *__restrict__ rhs, int tid, int nthreads, int rl)
{
    constexpr int R = ring_rows(T);
    if (rl > G.yhi) return;
    const int WP = p.WS >> 1;
    double *slot = sm + ((rl - G.ylo) % R) * slot_stride(p.WS);
    const size_t rowoff = (size_t)rl * p.ld;
    for (int c = tid; c < p.WS; c += nthreads) {
        const bool isP = c >= WP;
        const int k = isP ? c - WP : c;
        const int gc = G.gx0 + 2 * k;
        double *dE = slot + arr_off(p.WS, isP ? 2 : 0) + k;
        double *dO = slot + arr_off(p.WS, isP ? 3 : 1) + k;
        if (gc >= 0 && gc < p.ld) {
            const double *src = (isP ? rhs : in) + rowoff + gc;
            cp_async8(dE, src);
            cp_async8(dO, src + 1);
        } else {
            *dE = 0.0;
            *dO = 0.0;
        }
    }
}

// ---- phase 2: write back the finished row qs (all 2T half-sweeps applied) ---------------------
template <int T>
CNV_HD void phase_store(const PassGeom &p, const CtaGeom &G, const double *sm, double *__restrict__ out,
                        int tid, int nthreads, int qs)
{
    constexpr int R = ring_rows(T);
    if (qs < G.y0 || qs >= G.y1) return;
    const double *slot = sm + ((qs - G.ylo) % R) * slot_stride(p.WS);
    const double *SE = slot + arr_off(p.WS, 0), *SO = slot + arr_off(p.WS, 1);
    const int kbeg = p.HX >> 1, kend = (p.HX + p.Wout) >> 1;
    double *orow = out + (size_t)qs * p.ld;
    for (int k = kbeg + tid; k < kend; k += nthreads) {
        const int gc = G.gx0 + 2 * k;
        if (gc < p.ld) {
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<double2 *>(orow + gc) = make_double2(SE[k], SO[k]);
#else
            orow[gc] = SE[k];
            orow[gc + 1] = SO[k];
#endif
        }
    }
}

// ---- phase 3: this thread's red and black half-row updates of step r --------------------------
template <int T, bool POW2>
CNV_HD void phase_compute(const PassGeom &p, const CtaGeom &G, const RelaxConsts &rc, double *sm,
                          const ThreadCtx &t, int r, int nsw, double &acc)
{
    constexpr int R = ring_rows(T);
    if (t.g >= nsw) return;  // level not active in a shortened pass
    const int ss = slot_stride(p.WS);
    const int k0 = t.k0;
#pragma unroll
    for (int colour = 0; colour < 2; ++colour) {
        const int q = r - 1 - 4 * t.g - 2 * colour;
        if (q <= G.ylo || q >= G.yhi) continue;  // first/last streamed row: never updated
        const int gq = p.grow0 + q;
        if (gq < 1 || gq > p.gnrows - 2) continue;  // Dirichlet ring rows stay as loaded
        // colour 0 = red = (i+j) even (src/poisson.c:247): column parity == row parity
        const int typeO = (gq + colour) & 1;  // 0: even-column cells, 1: odd-column cells
        double *s0 = sm + ((q - G.ylo) % R) * ss;
        const double *sN = sm + ((q + 1 - G.ylo) % R) * ss;
        const double *sS = sm + ((q - 1 - G.ylo) % R) * ss;
        const int oA = arr_off(p.WS, typeO), oB = arr_off(p.WS, typeO ^ 1), oP = arr_off(p.WS, 2 + typeO);
#if defined(__CUDA_ARCH__)
        const double2 N = *reinterpret_cast<const double2 *>(sN + oA + k0);
        const double2 S = *reinterpret_cast<const double2 *>(sS + oA + k0);
        const double2 own = *reinterpret_cast<const double2 *>(s0 + oA + k0);
        const double2 Pv = *reinterpret_cast<const double2 *>(s0 + oP + k0);
        const double2 b = *reinterpret_cast<const double2 *>(s0 + oB + k0);
#else
        struct d2 { double x, y; };
        const d2 N = {sN[oA + k0], sN[oA + k0 + 1]}, S = {sS[oA + k0], sS[oA + k0 + 1]};
        const d2 own = {s0[oA + k0], s0[oA + k0 + 1]}, Pv = {s0[oP + k0], s0[oP + k0 + 1]};
        const d2 b = {s0[oB + k0], s0[oB + k0 + 1]};
#endif
        const double x = s0[oB + k0 + (typeO ? 2 : -1)];
        // even-column cell of pair k: W = odd[k-1], E = odd[k];  odd-column cell: W = even[k], E = even[k+1]
        const double W0 = typeO ? b.x : x, E0 = typeO ? b.y : b.x;
        const double W1 = typeO ? b.y : b.x, E1 = typeO ? x : b.y;
        double n0 = relax<POW2>(N.x, S.x, E0, W0, own.x, Pv.x, rc);
        double n1 = relax<POW2>(N.y, S.y, E1, W1, own.y, Pv.y, rc);
        n0 = (t.vmask >> typeO) & 1 ? n0 : own.x;
        n1 = (t.vmask >> (2 + typeO)) & 1 ? n1 : own.y;
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<double2 *>(s0 + oA + k0) = make_double2(n0, n1);
#else
        s0[oA + k0] = n0;
        s0[oA + k0 + 1] = n1;
#endif
        if (t.colown && q >= G.y0 && q < G.y1) {
            acc = xadd(acc, fabs(xsub(n0, own.x)));
            acc = xadd(acc, fabs(xsub(n1, own.y)));
        }
    }
}

// first and last step index of a CTA's stream
CNV_HD int first_step(const CtaGeom &G) { return G.ylo; }
template <int T>
CNV_HD int last_step(const CtaGeom &G) { return G.y1 - 1 + 4 * T; }

// ---- solver state machine (one instance per solve, device resident) ---------------------------
// Reference semantics (src/poisson.c:234-284): for k = 0..itmax-1 { sweep; e = sum|u-u0|;
// if (e < tol) return u (log k) }; exit(1).  A pass applies nsw <= T sweeps and records one norm
// per sweep.  If the first norm below tol belongs to the LAST sweep of the pass the output
// buffer is the answer.  If it belongs to an earlier sweep s, the pass input is still intact
// (passes are out of place), so the next pass recomputes exactly s+1 sweeps from it ("redo").
struct PoissonCtl {
    int state;   // 0 running, 1 converged, 2 itmax reached without convergence
    int cur;     // buffer (0/1) holding the current iterate = input of the next pass
    int sweeps;  // sweeps applied to buffer `cur`
    int redo;    // > 0: the next pass applies exactly `redo` sweeps and finishes
    int itmax;
    int result_k;  // reference's logged iteration number (sweeps - 1)
    unsigned ticket;
    int passes;    // passes that did work
    double tol;
    double result_e;
    double last_e;
};

CNV_HD int pass_sweeps(const PoissonCtl &c, int T)
{
    if (c.redo > 0) return c.redo;
    int left = c.itmax - c.sweeps;
    return left < T ? left : T;
}

// e[0..nsw-1]: global L1 update norms of the sweeps of the pass just finished
CNV_HD void decide(PoissonCtl &c, const double *e, int nsw, double *hist)
{
    c.passes++;
    if (c.redo > 0) {  // recomputation up to the converged sweep: done
        c.cur ^= 1;
        c.sweeps += nsw;
        c.result_k = c.sweeps - 1;
        c.result_e = e[nsw - 1];
        c.last_e = e[nsw - 1];
        c.redo = 0;
        c.state = 1;
        return;
    }
    int hit = -1;
    for (int s = 0; s < nsw; s++) {
        if (hist) hist[c.sweeps + s] = e[s];
        if (e[s] < c.tol) { hit = s; break; }
    }
    if (hit == nsw - 1) {
        c.cur ^= 1;
        c.sweeps += nsw;
        c.result_k = c.sweeps - 1;
        c.result_e = e[hit];
        c.last_e = e[hit];
        c.state = 1;
    } else if (hit >= 0) {
        c.redo = hit + 1;  // `cur` untouched: the pass input is recomputed with hit+1 sweeps
    } else {
        c.cur ^= 1;
        c.sweeps += nsw;
        c.last_e = e[nsw - 1];
        c.result_k = c.sweeps - 1;
        c.result_e = e[nsw - 1];
        if (c.sweeps >= c.itmax) c.state = 2;
    }
}

}  // namespace cnv
