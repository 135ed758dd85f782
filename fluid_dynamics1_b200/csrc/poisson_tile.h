// poisson_tile.h -- temporally blocked red-black SOR with a STATIONARY tile per CTA: the pass kernel for grids
// that fit the GPU's shared memory / L2 (<= ~2048^2 per GPU), shared between the CUDA kernel (poisson_tile.cu)
// and the host-side schedule checker (tests/emul/stream_emul.cc).
//
// Why a second pass kernel.  The streaming kernel (poisson_stream.h) pays 2*2T halo rows plus 4T pipeline-fill
// steps per CTA; at 4096^2 that is ~12 % of a CTA's work, at 1024^2 (32 output rows per CTA) it is 70 %.  Here a
// CTA loads a tile of TH x TW cells (output region + a 2T halo on all four sides) once, applies the 2T half-sweeps
// of the pass to the whole tile with one __syncthreads per half-sweep, and writes the output region back: no
// pipeline fill, halo overhead only.  Same sweeps, same operand values (src/poisson.c:238-262): a cell at depth d
// from the tile edge sees a stale neighbour at half-sweep d+1 at the earliest, so after 2T half-sweeps everything at
// depth >= 2T is exact -- that is the output region; what is computed in the halo is discarded.
//
// Thread (kp, seg) owns column pair kp (columns 2kp, 2kp+1 of the tile) in rows seg*M .. seg*M+M-1 and keeps those
// 2M values of psi in REGISTERS for the whole pass.  In a half-sweep it updates its M cells of the current colour
// (one per row, alternating even/odd column).  Of the operands of a cell only three come from shared memory: the
// horizontal neighbour that belongs to the adjacent thread, the right-hand side, and (first/last row only) the
// vertical neighbour in the adjacent segment; the other neighbours are the thread's own registers.  Updated cells
// are published to shared memory for the neighbours (column-parity split arrays, unit-stride across a warp).
#pragma once
#include "poisson_plan.h"

namespace cnv {

struct TileGeom {
    // local array: nrows x ncols doubles with pitch ld; row 0 is global row grow0 of gnrows; rows [own_lo, own_hi) are produced
    int nrows, ncols, ld, grow0, gnrows, own_lo, own_hi;
    int T;       // sweeps per pass
    int HT;      // halo cells on every side = 2T
    int KP;      // column pairs per tile (tile width 2*KP)
    int M;       // rows per thread (even)
    int NSEG;    // row segments (tile height TH = NSEG*M); threads per CTA = KP*NSEG
    int PK;      // shared row pitch in doubles = KP + 1 (one pad: right pad of a row == left pad of the next)
    int OW, OH;  // output columns / rows per tile
    int ntx, nty;
};

// Registers: 2M doubles of psi per thread.  Up to M = 10 the kernel fits the streaming kernel's budget (<= 102
// registers, 640 threads); taller thread columns run with fewer, fatter threads.
constexpr int tile_max_threads(int M) { return M <= 10 ? 640 : 448; }

CNV_HD int tile_rows(const TileGeom &g) { return g.NSEG * g.M; }
// four arrays (SE, SO: psi even/odd columns; PE, PO: right-hand side), each with a pad row below row 0 and above
// row TH-1 plus one pad element at either end: "row -1" and "row TH" are addressable and belong to nobody (the
// first / last segment read them for tile rows that are never updated; no aliasing with another array's cells)
CNV_HD int tile_arr_doubles(const TileGeom &g) { return (tile_rows(g) + 2) * g.PK + 2; }
CNV_HD size_t tile_smem_bytes(const TileGeom &g) { return (size_t)4 * tile_arr_doubles(g) * sizeof(double); }

template <int M>
struct TileThread {
    double E[M], O[M];   // psi of this thread's cells: even / odd column of its pair, rows tr0 .. tr0+M-1
    int base;            // byte offset of SE[tr0][kp]
    int pitch;           // PK * 8
    int aSO, aPE, aPO;   // byte distance from SE to the other arrays
    unsigned rvalid;     // bit i: row tr0+i may be updated (off the Dirichlet ring, inside the array, not a tile edge row)
    unsigned rown;       // bit i: row tr0+i belongs to the tile's output rows
    unsigned rin;        // bit i: row tr0+i exists in the local array
    bool ce, co;         // even / odd column updatable (off the ring, inside the domain)
    bool colin;          // the pair exists in the local array (pitch columns)
    bool colown;         // the pair belongs to the tile's output columns
    bool fast;           // everything updatable: select-free update
    long long gofs;      // element offset of (row tr0, even column) in the global arrays
    int ld;
    double acc;          // |u - u0| of this thread's output cells in the current sweep
};

// ---- load: this thread's cells -> registers + shared memory; right-hand side -> shared memory ----
template <int M>
CNV_HD void tile_load(TileThread<M> &t, const TileGeom &g, int bx, int by, int tid, double *sm, const double *in, const double *rhs)
{
    const int kp = tid % g.KP, seg = tid / g.KP;
    const int tr0 = seg * M, TH = tile_rows(g);
    const int gx0 = bx * g.OW - g.HT;            // column of tile column 0 (even)
    const int ty0 = g.own_lo + by * g.OH - g.HT;  // local row of tile row 0
    const int gc = gx0 + 2 * kp;
    const int arr = tile_arr_doubles(g) * 8;
    t.pitch = g.PK * 8;
    t.base = smem_base(sm) + (g.PK + 1) * 8 + (tr0 * g.PK + kp) * 8;  // (pad element + pad row) in front of row 0
    t.aSO = arr; t.aPE = 2 * arr; t.aPO = 3 * arr;
    t.ld = g.ld;
    t.colin = gc >= 0 && gc < g.ld;
    t.ce = gc >= 1 && gc <= g.ncols - 2;
    t.co = gc + 1 >= 1 && gc + 1 <= g.ncols - 2;
    const int c0 = bx * g.OW;
    t.colown = gc >= c0 && gc < c0 + g.OW && gc < g.ld;
    const int y0 = g.own_lo + by * g.OH, y1 = y0 + g.OH < g.own_hi ? y0 + g.OH : g.own_hi;
    t.rvalid = t.rown = t.rin = 0;
    for (int i = 0; i < M; i++) {
        const int tr = tr0 + i, lr = ty0 + tr, gr = g.grow0 + lr;
        if (lr >= 0 && lr < g.nrows) t.rin |= 1u << i;
        if (lr >= 1 && lr <= g.nrows - 2 && gr >= 1 && gr <= g.gnrows - 2 && tr >= 1 && tr <= TH - 2) t.rvalid |= 1u << i;
        if (lr >= y0 && lr < y1) t.rown |= 1u << i;
    }
    t.fast = t.ce && t.co && t.rvalid == (1u << M) - 1;
    t.gofs = (long long)(ty0 + tr0) * g.ld + gc;
    t.acc = 0.0;
    // all 2M global loads first (they are independent; the shared-memory stores below are ordered asm statements with a
    // memory clobber, so loads issued between them would wait for one L2 round trip per row), then the stores
    dbl2 f[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        dbl2 u = {0.0, 0.0};
        f[i] = u;
        if (t.colin && ((t.rin >> i) & 1u)) {
            const long long off = t.gofs + (long long)i * g.ld;
            u = ldg2(in + off);
            f[i] = ldg2(rhs + off);
        }
        t.E[i] = u.x; t.O[i] = u.y;
    }
#pragma unroll
    for (int i = 0; i < M; i++) {
        const int a = t.base + i * t.pitch;
        sts1(sm, a, t.E[i]); sts1(sm, a + t.aSO, t.O[i]);
        sts1(sm, a + t.aPE, f[i].x); sts1(sm, a + t.aPO, f[i].y);
        if (kp == 0) { sts1(sm, a - 8, 0.0); sts1(sm, a + t.aSO - 8, 0.0); }  // the row pads (deterministic halo garbage)
    }
}

// The M updates of a half-sweep are independent of each other: their shared-memory operands are cells of the OTHER
// colour and right-hand sides (nobody writes those during this half-sweep), and the vertical neighbours of row i are
// cells of the column that rows i-1 / i+1 do NOT update now (the updated column alternates from row to row).  So they
// run in batches of CH rows -- all loads of a batch, then its updates, then its stores -- instead of M
// load -> relax -> store round trips behind ordered shared-memory accesses.
// SEL: apply the per-row / per-column validity selects (threads next to the Dirichlet ring, the array edge or the tile
// edge).  NORM: 0 = the thread owns no output cell, 1 = all its cells are output cells, 2 = per-row test.
template <int M, bool POW2, int PH, bool SEL, int NORM>
CNV_HD void tile_half_sweep_body(TileThread<M> &t, const RelaxConsts &rc, double *sm)
{
    constexpr int p0 = PH & 1, pl = (PH + M - 1) & 1;
    constexpr int CH = 4;  // rows per batch: 4 independent updates in flight (like the streaming kernel's 4 cells per step)
    // vertical neighbours outside the thread's rows: same parity array as the cell that needs them
    const double below = lds1(sm, t.base + (p0 ? t.aSO : 0) - t.pitch);
    const double above = lds1(sm, t.base + (pl ? t.aSO : 0) + M * t.pitch);
#pragma unroll
    for (int c = 0; c < M; c += CH) {
        double X[CH], P[CH], nv[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const int i = c + j;
            if (i < M) {
                const int a = t.base + i * t.pitch;
                const bool odd = ((PH + i) & 1) != 0;
                // even-column cell of pair k: W = odd[k-1] (neighbour thread);  odd-column cell: E = even[k+1] (neighbour thread)
                X[j] = lds1(sm, odd ? a + 8 : a + t.aSO - 8);
                P[j] = lds1(sm, a + (odd ? t.aPO : t.aPE));
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const int i = c + j;
            if (i < M) {
                const bool odd = ((PH + i) & 1) != 0;
                if (!odd) {  // even-column cell: W = odd[k-1] (neighbour thread), E = odd[k] (own)
                    const double S = i == 0 ? below : t.E[i > 0 ? i - 1 : 0], N = i == M - 1 ? above : t.E[i < M - 1 ? i + 1 : M - 1];
                    nv[j] = relax<POW2>(N, S, t.O[i], X[j], t.E[i], P[j], rc);
                } else {     // odd-column cell: W = even[k] (own), E = even[k+1] (neighbour thread)
                    const double S = i == 0 ? below : t.O[i > 0 ? i - 1 : 0], N = i == M - 1 ? above : t.O[i < M - 1 ? i + 1 : M - 1];
                    nv[j] = relax<POW2>(N, S, X[j], t.E[i], t.O[i], P[j], rc);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const int i = c + j;
            if (i < M) {
                const int a = t.base + i * t.pitch;
                const bool odd = ((PH + i) & 1) != 0;
                const double old = odd ? t.O[i] : t.E[i];
                double v = nv[j];
                if (SEL) {
                    const bool ok = ((t.rvalid >> i) & 1u) && (odd ? t.co : t.ce);
                    v = ok ? v : old;
                }
                if (odd) { t.O[i] = v; sts1(sm, a + t.aSO, v); }
                else { t.E[i] = v; sts1(sm, a, v); }
                // L1 update norm over the output cells (cells that were not updated contribute exactly 0)
                if (NORM == 1) t.acc = xadd(t.acc, fabs(xsub(v, old)));
                else if (NORM == 2 && ((t.rown >> i) & 1u)) t.acc = xadd(t.acc, fabs(xsub(v, old)));
            }
        }
    }
}

// ---- one half-sweep.  PH = column parity (0 even, 1 odd) of the cell this thread updates in its row 0; the
// parity alternates from row to row.  M is even and so is every segment start, hence PH is the same for all
// threads of the CTA: PH = (global row of tile row 0 + colour) & 1, colour 0 = red = (i + j) even (src/poisson.c:247).
template <int M, bool POW2, int PH>
CNV_HD void tile_half_sweep(TileThread<M> &t, const RelaxConsts &rc, double *sm)
{
#if defined(__CUDA_ARCH__)
    // Row addresses are base + i * pitch with a run-time pitch.  Left alone, the compiler hoists all of them (4 arrays x M
    // rows) out of the sweep loop, which costs ~3M registers, spills, and leaves one temporary for all loads of a batch
    // (i.e. serialises them).  An opaque base makes them cheap per-use arithmetic instead.
    asm volatile("" : "+r"(t.base));
    // likewise the row masks: evaluated where the (rare) general bodies use them, not packed into predicates at the top of
    // every sweep for all threads
    asm volatile("" : "+r"(t.rvalid), "+r"(t.rown));
#endif
    // interior threads (everything updatable) either own output cells in all their rows or none (tile halo): two
    // select-free bodies; everything else (ring, array edge, tile edge, partial output rows) takes the general one
    const bool allown = t.colown && t.rown == (1u << M) - 1;
    const bool noown = !t.colown || t.rown == 0;
    if (t.fast && allown) tile_half_sweep_body<M, POW2, PH, false, 1>(t, rc, sm);
    else if (t.fast && noown) tile_half_sweep_body<M, POW2, PH, false, 0>(t, rc, sm);
    else if (noown) tile_half_sweep_body<M, POW2, PH, true, 0>(t, rc, sm);
    else tile_half_sweep_body<M, POW2, PH, true, 2>(t, rc, sm);
}

// ---- write back the output cells of this thread ----
template <int M>
CNV_HD void tile_store(const TileThread<M> &t, double *out)
{
    if (!t.colown) return;
#pragma unroll
    for (int i = 0; i < M; i++)
        if ((t.rown >> i) & 1u) stg2(out + t.gofs + (long long)i * t.ld, t.E[i], t.O[i]);
}

// ---- slab geometry of a tile row `by` (multi-GPU peer exchange; shared with the CPU tests) ------------------------
// Output rows [y0, y1) of the tiles in tile row `by`, the local rows the tile reads, and the boundary rows it has to push
// into the lower (side 0) / upper (side 1) neighbour's halo: the 2T owned rows next to that slab edge.
struct TileRows {
    int y0, y1;        // output rows
    int rlo, rhi;      // rows read: [rlo, rhi)
    int pa[2], pb[2];  // rows to push towards side s: [pa[s], pb[s]) (empty if pa >= pb)
};
CNV_HD TileRows tile_rows_of(const TileGeom &g, int by)
{
    TileRows r;
    r.y0 = g.own_lo + by * g.OH;
    r.y1 = r.y0 + g.OH < g.own_hi ? r.y0 + g.OH : g.own_hi;
    const int ty0 = r.y0 - g.HT, TH = tile_rows(g);
    r.rlo = ty0 > 0 ? ty0 : 0;
    r.rhi = ty0 + TH < g.nrows ? ty0 + TH : g.nrows;
    r.pa[0] = r.y0 > g.own_lo ? r.y0 : g.own_lo;
    r.pb[0] = r.y1 < g.own_lo + g.HT ? r.y1 : g.own_lo + g.HT;
    r.pa[1] = r.y0 > g.own_hi - g.HT ? r.y0 : g.own_hi - g.HT;
    r.pb[1] = r.y1 < g.own_hi ? r.y1 : g.own_hi;
    return r;
}

// ---- planner -----------------------------------------------------------------------------------
// Picks the tile shape (KP, M, NSEG) with the lowest estimated time per sweep: waves x (fixed cost + 2T half-sweeps x
// M rows x the warps sharing a scheduler) / T.  Feasible: whole warps <= tile_max_threads(M), shared memory within the
// opt-in limit.  force_* > 0 pin a dimension (tests, tools/probe_poisson.py).
inline bool tile_plan(int nrows, int ncols, int ld, int grow0, int gnrows, int own_lo, int own_hi, int T, int num_sms,
                      size_t smem_limit, TileGeom *best, double *best_cost, int force_kp = 0, int force_m = 0, int force_nseg = 0)
{
    TileGeom b;
    std::memset(&b, 0, sizeof b);
    double bc = 1e300;
    const int own = own_hi - own_lo, HT = 2 * T;
    for (int M = 6; M <= 16; M += 2) {
        if (force_m && M != force_m) continue;
        for (int KP = HT + 4; KP <= 160; KP++) {
            if (force_kp && KP != force_kp) continue;
            const int OW = 2 * KP - 2 * HT;
            if (!force_kp && OW - 2 >= ncols && KP > HT + 4) break;  // wider than the domain needs
            for (int NSEG = 2; round_up(NSEG * KP, 32) <= tile_max_threads(M); NSEG++) {
                if (force_nseg && NSEG != force_nseg) continue;
                const int TH = NSEG * M, OH = TH - 2 * HT;
                if (OH < 2) continue;
                TileGeom g;
                std::memset(&g, 0, sizeof g);
                g.T = T; g.HT = HT; g.KP = KP; g.M = M; g.NSEG = NSEG; g.PK = KP + 1; g.OW = OW; g.OH = OH;
                if (tile_smem_bytes(g) > smem_limit) break;
                if (!force_nseg && OH - M >= own && NSEG > 2) break;  // taller than the slab needs
                const int threads = KP * NSEG;
                if (threads < 128 && !force_kp) continue;
                g.ntx = (ncols + OW - 1) / OW;  // pad columns beyond ncols stay zero in both buffers: no tile of their own
                g.nty = (own + OH - 1) / OH;
                const long tiles = (long)g.ntx * g.nty;
                const long waves = (tiles + num_sms - 1) / num_sms;
                const int wps = (threads / 32 + 3) / 4;  // warps per scheduler
                const double hs = (double)M * (wps < 2 ? 2 : wps);
                const double cost = waves * (2.0 * T * hs + 420.0) / T;
                if (cost < bc) { bc = cost; b = g; }
            }
        }
    }
    if (bc >= 1e300) return false;
    b.nrows = nrows; b.ncols = ncols; b.ld = ld; b.grow0 = grow0; b.gnrows = gnrows; b.own_lo = own_lo; b.own_hi = own_hi;
    *best = b;
    if (best_cost) *best_cost = bc;
    return true;
}

}  // namespace cnv
