// poisson_onchip.h -- the whole Poisson solve in ONE persistent launch with the iterate resident ON CHIP (registers),
// for grids that fit the GPU's register files (<= ~1.5 M cells: BASELINE config 3, 1024^2, and everything below it).
// Shared between the CUDA kernel (poisson_onchip.cu) and the host-side schedule checker (tests/emul/stream_emul.cc).
//
// Why.  At 1024^2 the fields (2 x 8 MB) live in L2 and the streaming pass kernel (poisson_stream.h) is bound by what
// it pays per PASS, not per cell: 98 streamed row-steps per CTA for 31 useful rows, one launch + one grid-wide
// reduction per 8 sweeps (6.4 us per sweep, 0.55 of the 24 B/cell HBM roofline).  Here a CTA per SM keeps one tile of
// psi and of the right-hand side in REGISTERS for the whole solve (src/poisson.c:234-279: all sweeps and the
// convergence test), and only tile-boundary data ever moves:
//   * thread (px, py) owns a patch of kOcPW = 4 columns x kOcPM = 8 rows (32 cells of psi + their 32 right-hand sides);
//     a half-sweep updates its 16 cells of one colour, all independent of each other; of their 64 operands 52 are the
//     thread's own registers, 12 belong to the four adjacent threads and come from shared memory, where every thread
//     publishes the 24 perimeter cells of its patch (arrays indexed [slot][thread]: unit stride across a warp for the
//     owner AND for all four neighbours, so conflict-free for any tile shape); one __syncthreads per half-sweep;
//   * the tile carries a halo of 2T cells on all four sides; T sweeps ("a pass") run without any communication (the
//     stale halo edge spreads one cell per half-sweep and stays inside the halo, whose results are discarded);
//   * after a pass the CTA stores its output cells to the pass' global iterate buffer (L2), raises its flag
//     (release, gpu scope), waits for the flags of its <= 8 neighbour tiles and reloads ONLY its halo cells;
//   * the stop decision lags one pass behind (lag_fold / lag_action of poisson_stream.h, three iterate buffers in
//     rotation): a pass starts knowing the global norms of the pass before the previous one, which every CTA sums
//     itself from all CTAs' partials in a fixed order, so there is no grid-wide barrier on the critical path;
//     if the first sweep below `tol` was not the last sweep of its pass, the tiles are reloaded from that pass' (still
//     intact) input and exactly the converged number of sweeps is recomputed.
// Same sweeps, same operand values, same separately rounded operations (exact.h::relax) as the reference's red-black
// loop: the field is bit-identical; the L1 norm differs only in summation order (as for the other kernels).
#pragma once
#include "poisson_plan.h"

namespace cnv {

constexpr int kOcPW = 4;           // patch columns per thread
constexpr int kOcPM = 8;           // patch rows per thread
#ifndef CNV_OC_OCC
#define CNV_OC_OCC 1
#endif
// CTAs per SM.  Two half-size CTAs per SM (so that one computes while the other sits in the latency chain between two passes)
// were measured SLOWER at 1024^2: 4.2-4.6 vs 3.9 us per sweep (more tiles = more halo cells to recompute, and the sweeps of the
// two CTAs share the fp64 pipe, which is already half busy); profiles/onchip_r2.md.
constexpr int kOcOcc = CNV_OC_OCC;
constexpr int kOcMaxThreads = 384 / kOcOcc - 32;  // compute threads; + one service warp = 384 x 168 registers = 64512 per SM
constexpr int kOcMaxCtas = 148 * kOcOcc;     // (the kernel gathers all CTAs' norm partials in a fixed-size shared array)
constexpr int kOcPad = 64;         // published arrays are addressable for thread ids -kOcPad .. kOcMaxThreads + kOcPad - 1
constexpr int kOcPitch = kOcMaxThreads + 2 * kOcPad;  // doubles per slot
constexpr int kOcSlots = 2 * kOcPM + 2 * kOcPW;       // ColLo[PM], ColHi[PM], RowLo[PW], RowHi[PW]
constexpr int kOcNormSlots = 4;    // rings of per-pass norm partials
constexpr int kOcFlagStride = 16;  // one 128-byte line per CTA flag: pollers of different flags hit different L2 lines
static_assert(kOcPM % 2 == 0 && kOcPW == 4, "colour pattern of a patch is static");

CNV_HD int oc_slot_col_lo(int i) { return i; }                        // cell (i, 0)
CNV_HD int oc_slot_col_hi(int i) { return kOcPM + i; }                // cell (i, PW-1)
CNV_HD int oc_slot_row_lo(int j) { return 2 * kOcPM + j; }            // cell (0, j)
CNV_HD int oc_slot_row_hi(int j) { return 2 * kOcPM + kOcPW + j; }    // cell (PM-1, j)
inline size_t oc_smem_bytes() { return (size_t)kOcSlots * kOcPitch * sizeof(double); }

struct OnchipGeom {
    // local array: nrows x ncols doubles with pitch ld; row 0 is global row grow0 of gnrows; rows [own_lo, own_hi) are produced
    int nrows, ncols, ld, grow0, gnrows, own_lo, own_hi;
    int T;         // sweeps per pass
    int HX, HY;    // halo columns / rows on either side: 2T rounded up to whole patches (multiples of 4 / 8)
    int OW, OH;    // output columns / rows per tile (multiples of 4 / 8): a patch is all output or all halo
    int NPX, NPY;  // patches per tile; tile = (4 NPX) x (8 NPY) cells, threads per CTA = NPX * NPY
    int ntx, nty;  // tiles
};

CNV_HD int oc_threads(const OnchipGeom &g) { return g.NPX * g.NPY; }

struct OcTile {
    int tx0, ty0;  // array column / row of tile cell (0, 0) (may be negative)
    int x0, x1;    // output columns [x0, x1)
    int y0, y1;    // output rows [y0, y1)
    int par0;      // colour of tile cell (0, 0): (global row + column) & 1
};

CNV_HD OcTile oc_tile(const OnchipGeom &g, int bx, int by)
{
    OcTile t;
    t.x0 = bx * g.OW;
    t.x1 = t.x0 + g.OW < g.ld ? t.x0 + g.OW : g.ld;
    t.y0 = g.own_lo + by * g.OH;
    t.y1 = t.y0 + g.OH < g.own_hi ? t.y0 + g.OH : g.own_hi;
    t.tx0 = t.x0 - g.HX;
    t.ty0 = t.y0 - g.HY;
    t.par0 = (g.grow0 + t.ty0 + t.tx0) & 1;
    return t;
}

// Per-thread state.  p / f: the patch of psi and of the (pre-scaled) right-hand side, row i = 0..7 bottom-up, column j = 0..3.
// Output regions are aligned to patches, so a thread is an OUTPUT thread (all its cells inside the array belong to the tile's
// output region: stored after a pass, counted in the norm) or a HALO thread (reloaded after every pass, norm discarded).
struct OcThread {
    double p[kOcPM][kOcPW];
    double f[kOcPM][kOcPW];
    unsigned upd;       // bit 4 i + j: cell may be updated (inside the array, off the Dirichlet ring, off the array's first / last row)
    unsigned inr;       // bit i: patch row i exists in the array (global loads / stores move whole 32-byte patch rows)
    bool band;          // output thread whose patch lies in the boundary band of the output region (a neighbour's halo)
    bool own;           // output thread
    int src;            // halo thread: the tile (CTA index) whose output region its patch lies in, -1 if outside the array
    bool fast;          // all 32 cells updatable
    int dist;           // distance (cells) of the patch from the output region: 0 for output threads
    long long gofs;     // element offset of cell (0, 0) in the global arrays
    int ld;
    int self, west, east, south, north;  // byte offsets of slot 0 for this thread and its four neighbours
};

CNV_HD void oc_thread_init(OcThread &t, const OnchipGeom &g, const OcTile &tl, int tid, const double *sm)
{
    const int px = tid % g.NPX, py = tid / g.NPX;
    const int c0 = tl.tx0 + kOcPW * px, r0 = tl.ty0 + kOcPM * py;  // array coordinates of patch cell (0, 0)
    t.upd = t.inr = 0;
    for (int i = 0; i < kOcPM; i++)
        for (int j = 0; j < kOcPW; j++) {
            const int ar = r0 + i, ac = c0 + j, gr = g.grow0 + ar;
            const bool in = ar >= 0 && ar < g.nrows && ac >= 0 && ac < g.ld;  // (c0 and ld are multiples of 4: a row is in or out)
            if (in && ar >= 1 && ar <= g.nrows - 2 && gr >= 1 && gr <= g.gnrows - 2 && ac >= 1 && ac <= g.ncols - 2) t.upd |= 1u << (4 * i + j);
            if (in && j == 0) t.inr |= 1u << i;
        }
    // patch-aligned regions: the patch origin decides (x0, y0, OW, OH, HX, HY are multiples of the patch size)
    t.own = r0 >= tl.y0 && r0 < tl.y1 && c0 >= tl.x0 && c0 < tl.x1;
    // the part of the output region the neighbour tiles read as their halo: stored before the flag is raised
    t.band = t.own && (r0 < tl.y0 + g.HY || r0 + kOcPM > tl.y1 - g.HY || c0 < tl.x0 + g.HX || c0 + kOcPW > tl.x1 - g.HX);
    t.src = -1;
    // (pad columns beyond the last tile's output region belong to nobody: they hold zeros for ever)
    if (!t.own && r0 >= g.own_lo && r0 < g.own_hi && c0 >= 0 && c0 < g.ld && c0 / g.OW < g.ntx)
        t.src = ((r0 - g.own_lo) / g.OH) * g.ntx + c0 / g.OW;
    // distance to the output region (Chebyshev: information moves one cell per half-sweep in either direction)
    const int dr = r0 + kOcPM - 1 < tl.y0 ? tl.y0 - (r0 + kOcPM - 1) : (r0 >= tl.y1 ? r0 - tl.y1 + 1 : 0);
    const int dc = c0 + kOcPW - 1 < tl.x0 ? tl.x0 - (c0 + kOcPW - 1) : (c0 >= tl.x1 ? c0 - tl.x1 + 1 : 0);
    t.dist = dr > dc ? dr : dc;
    t.fast = t.upd == 0xffffffffu;
    t.gofs = (long long)r0 * g.ld + c0;
    t.ld = g.ld;
    const int base = smem_base(sm) + kOcPad * 8;
    t.self = base + tid * 8;
    t.west = t.self - 8;
    t.east = t.self + 8;
    t.south = t.self - g.NPX * 8;
    t.north = t.self + g.NPX * 8;
    for (int i = 0; i < kOcPM; i++)
        for (int j = 0; j < kOcPW; j++) t.p[i][j] = t.f[i][j] = 0.0;
}

constexpr int kOcSlotBytes = kOcPitch * 8;

// publish the perimeter cells of the patch (all of them, both colours)
CNV_HD void oc_publish_all(const OcThread &t, double *sm)
{
#pragma unroll
    for (int i = 0; i < kOcPM; i++) {
        sts1(sm, t.self + oc_slot_col_lo(i) * kOcSlotBytes, t.p[i][0]);
        sts1(sm, t.self + oc_slot_col_hi(i) * kOcSlotBytes, t.p[i][kOcPW - 1]);
    }
#pragma unroll
    for (int j = 0; j < kOcPW; j++) {
        sts1(sm, t.self + oc_slot_row_lo(j) * kOcSlotBytes, t.p[0][j]);
        sts1(sm, t.self + oc_slot_row_hi(j) * kOcSlotBytes, t.p[kOcPM - 1][j]);
    }
}

// (re)load the patch of psi from `in` (the rows that exist in the array); all loads are issued before the first use
CNV_HD void oc_load_psi(OcThread &t, const double *in)
{
#pragma unroll
    for (int i = 0; i < kOcPM; i++)
        if ((t.inr >> i) & 1u) {
            const dbl4 v = ldg4(in + t.gofs + (long long)i * t.ld);
            t.p[i][0] = v.x; t.p[i][1] = v.y; t.p[i][2] = v.z; t.p[i][3] = v.w;
        }
}
CNV_HD void oc_load_rhs(OcThread &t, const double *rhs)
{
#pragma unroll
    for (int i = 0; i < kOcPM; i++)
        if ((t.inr >> i) & 1u) {
            const dbl4 v = ldg4(rhs + t.gofs + (long long)i * t.ld);
            t.f[i][0] = v.x; t.f[i][1] = v.y; t.f[i][2] = v.z; t.f[i][3] = v.w;
        }
}
CNV_HD void oc_store(const OcThread &t, double *out)
{
#pragma unroll
    for (int i = 0; i < kOcPM; i++)
        if ((t.inr >> i) & 1u) stg4(out + t.gofs + (long long)i * t.ld, t.p[i][0], t.p[i][1], t.p[i][2], t.p[i][3]);
}

// One half-sweep.  PH = (i + j) & 1 of the cells updated (patch coordinates; the patch origin has even tile coordinates, so
// PH = colour ^ tile.par0 for every thread of the CTA).  The 16 updates only read cells of the other colour, which nobody
// writes during this half-sweep: they are independent, the in-place update is safe, and the neighbours' published cells
// read here were written in an earlier half-sweep (one barrier per half-sweep).
// SEL: apply the per-cell update mask (threads next to the Dirichlet ring or the array edge; chosen per WARP by the caller so
// that a warp never runs both bodies).  `acc` collects |new - old| of ALL 16 cells; halo threads discard theirs.
template <bool POW2, int PH, bool SEL>
CNV_HD void oc_half_sweep(OcThread &t, const RelaxConsts &rc, double *sm, double &acc)
{
    constexpr int PM = kOcPM, PW = kOcPW;
    // operands that belong to the adjacent threads: column 0 cells (rows with i & 1 == PH) need W, column 3 cells (the other
    // rows) need E, row 0 cells (columns with j & 1 == PH) need S, row PM-1 cells (the other columns) need N
    double wv[PM / 2], ev[PM / 2], sv[PW / 2], nv[PW / 2];
#pragma unroll
    for (int h = 0; h < PM / 2; h++) {
        wv[h] = lds1(sm, t.west + oc_slot_col_hi(2 * h + PH) * kOcSlotBytes);
        ev[h] = lds1(sm, t.east + oc_slot_col_lo(2 * h + (PH ^ 1)) * kOcSlotBytes);
    }
#pragma unroll
    for (int h = 0; h < PW / 2; h++) {
        sv[h] = lds1(sm, t.south + oc_slot_row_hi(2 * h + PH) * kOcSlotBytes);
        nv[h] = lds1(sm, t.north + oc_slot_row_lo(2 * h + (PH ^ 1)) * kOcSlotBytes);
    }
#pragma unroll
    for (int i = 0; i < PM; i++) {
#pragma unroll
        for (int h = 0; h < PW / 2; h++) {
            const int j = 2 * h + ((i + PH) & 1);
            const double N = i + 1 < PM ? t.p[i + 1 < PM ? i + 1 : i][j] : nv[h];
            const double S = i > 0 ? t.p[i > 0 ? i - 1 : 0][j] : sv[h];
            const double E = j + 1 < PW ? t.p[i][j + 1 < PW ? j + 1 : j] : ev[i / 2];
            const double W = j > 0 ? t.p[i][j > 0 ? j - 1 : 0] : wv[i / 2];
            const double old = t.p[i][j];
            double v = relax<POW2>(N, S, E, W, old, t.f[i][j], rc);
            if (SEL) v = ((t.upd >> (4 * i + j)) & 1u) ? v : old;
            t.p[i][j] = v;
            if (j == 0) sts1(sm, t.self + oc_slot_col_lo(i) * kOcSlotBytes, v);
            if (j == PW - 1) sts1(sm, t.self + oc_slot_col_hi(i) * kOcSlotBytes, v);
            if (i == 0) sts1(sm, t.self + oc_slot_row_lo(j) * kOcSlotBytes, v);
            if (i == PM - 1) sts1(sm, t.self + oc_slot_row_hi(j) * kOcSlotBytes, v);
            acc = xadd(acc, fabs(xsub(v, old)));  // (a cell that was not updated contributes exactly 0)
        }
    }
}

// ---- planner -----------------------------------------------------------------------------------------------------
// One tile per CTA, all CTAs co-resident (ntx * nty <= SMs: the kernel is launched cooperatively and waits on its
// neighbours' flags).  The cost per sweep (clock cycles) is fitted to the B200 measurements in profiles/onchip_r2.md; the
// exchange after every pass (flags through L2 + halo reload) is amortised over T sweeps.
inline bool onchip_plan(int nrows, int ncols, int ld, int grow0, int gnrows, int own_lo, int own_hi, int num_sms, OnchipGeom *best,
                        int force_T = 0, int force_ntx = 0, int force_nty = 0)
{
    OnchipGeom b;
    std::memset(&b, 0, sizeof b);
    double bc = 1e300;
    const int own = own_hi - own_lo;
    for (int T = 2; T <= 8; T += 2) {
        if (force_T && T != force_T) continue;
        const int HX = round_up(2 * T, kOcPW), HY = round_up(2 * T, kOcPM);
        for (int ntx = 1; ntx <= num_sms * kOcOcc; ntx++) {
            if (force_ntx && ntx != force_ntx) continue;
            const int OW = round_up((ncols + ntx - 1) / ntx, kOcPW);
            if (ntx > 1 && OW < HX) break;               // a tile's halo must come from its direct neighbours only
            if ((ncols + OW - 1) / OW != ntx) continue;  // (same tiling as a smaller ntx)
            const int NPX = (OW + 2 * HX) / kOcPW;
            if (NPX > kOcPad - 1) continue;
            for (int nty = 1; ntx * nty <= num_sms * kOcOcc && ntx * nty <= kOcMaxCtas; nty++) {
                if (force_nty && nty != force_nty) continue;
                const int OH = round_up((own + nty - 1) / nty, kOcPM);
                if (nty > 1 && OH < HY) break;
                if ((own + OH - 1) / OH != nty) continue;
                const int NPY = (OH + 2 * HY) / kOcPM;
                const int NT = NPX * NPY;
                if (NT > kOcMaxThreads) continue;
                // measured (profiles/onchip_r2.md): a sweep costs ~2800 cycles + 80 per compute warp (latency of the two
                // half-sweeps, then the fp64 pipe), the exchange between two passes ~6000 cycles
                const double warps = (NT + 31) / 32;
                const double cost = 2800.0 + 80.0 * warps + 6000.0 / T;
                if (cost < bc) {
                    bc = cost;
                    b.T = T; b.HX = HX; b.HY = HY; b.OW = OW; b.OH = OH; b.NPX = NPX; b.NPY = NPY; b.ntx = ntx; b.nty = nty;
                }
            }
        }
    }
    if (bc >= 1e300) return false;
    b.nrows = nrows; b.ncols = ncols; b.ld = ld; b.grow0 = grow0; b.gnrows = gnrows; b.own_lo = own_lo; b.own_hi = own_hi;
    *best = b;
    return true;
}

// neighbour tiles whose output a tile's halo comes from: (bx + dx, by + dy), dx, dy in {-1, 0, 1}, inside the tile grid
CNV_HD int oc_neighbours(const OnchipGeom &g, int bx, int by, int *out /* 8 */)
{
    int n = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            const int x = bx + dx, y = by + dy;
            if (x >= 0 && x < g.ntx && y >= 0 && y < g.nty) out[n++] = y * g.ntx + x;
        }
    return n;
}

}  // namespace cnv
