// poisson_plan.h -- host-side planning for the streaming Poisson kernel: relaxation constants
// (evaluated with the reference's expressions) and the strip/chunk decomposition of one pass.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "poisson_stream.h"

namespace cnv {

inline bool is_pow2_double(double x)
{
    int e;
    return x > 0 && std::frexp(x, &e) == 0.5;
}

// src/poisson.c:246 constants.  `beta` == 1 reproduces the Gauss-Seidel variants (:85, :198):
// beta*A/D == A/D and (1-beta)*u0 == +0.
inline RelaxConsts make_relax_consts(double dx, double dy, double beta)
{
    RelaxConsts c;
    std::memset(&c, 0, sizeof c);
    c.cxx = dx * dx;
    c.cyy = dy * dy;
    c.cf = dx * dx * dy * dy;
    c.D = 2 * (dx * dx + dy * dy);
    c.rD = 1.0 / c.D;
    c.beta = beta;
    c.omb = 1 - beta;
    c.pow2 = (dx == dy) && is_pow2_double(dx);
    c.bb = 0.25 * beta;
    c.pscale = c.pow2 ? c.cxx : c.cf;
    // Markstein's correction is exact unless the divisor's significand is all ones
    uint64_t bits;
    std::memcpy(&bits, &c.D, 8);
    c.true_div = ((bits & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull) || !std::isfinite(c.rD);
    if (const char *e = std::getenv("CNV_POISSON_GENERAL")) {  // force the literal operation sequence
        if (std::atoi(e)) { c.pow2 = 0; c.pscale = c.cf; }
    }
    if (const char *e = std::getenv("CNV_POISSON_TRUE_DIV")) c.true_div = std::atoi(e) != 0;
    return c;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

inline size_t pass_smem_bytes(int T, int WS) { return (size_t)ring_rows(T) * slot_stride(WS) * sizeof(double); }
inline int pass_threads(int T, int WS) { return T * (WS / (2 * kPairs)); }
// Launch bounds of k_poisson_pass<T>: the kernel is latency bound, so registers are capped (<= 102 per
// thread: column history + software-pipelined operands) to keep 20 warps resident per SM: one 640-thread CTA
// for deep blocking, two 320-thread CTAs otherwise.
constexpr int pass_max_threads(int T) { return T >= 6 ? 640 : 320; }
constexpr int pass_min_ctas(int T) { return T >= 6 ? 1 : 2; }

struct PlanLimits {
    int num_sms = 148;
    size_t smem_per_cta = 227 * 1024;
    size_t smem_per_sm = 228 * 1024;
    int max_threads_per_sm = 2048;
};

// Choose WS / Hout for `own` = [own_lo, own_hi) rows of an nrows x ncols local array.
// Cost model (fitted to B200 measurements, profiles/): the kernel is latency bound, so an SM's throughput
// grows with its resident warps until ~24; the work of a CTA is (steps x WS x T) cell-updates including
// halo and pipeline fill/drain; CTAs resident on one SM share it; the pass ends with the last wave, so the
// CTA count should fill the SM slots (148 x occupancy) as exactly as possible.
inline PassGeom make_plan(int nrows, int ncols, int ld, int grow0, int gnrows, int own_lo, int own_hi, int T,
                          const PlanLimits &lim, int force_ws = 0, int force_chunks = 0, int trim = -1)
{
    // rows taken off a boundary chunk whose CTAs exchange with a neighbour slab (see PassGeom::trim_lo)
    if (trim < 0) trim = std::getenv("CNV_POISSON_TRIM") ? std::atoi(std::getenv("CNV_POISSON_TRIM")) : 32;
    // (Making the chunks at the domain boundary 2T rows TALLER -- they have no halo rows to stream on that side -- was
    // tried and measured slower on B200: 4.99e11 vs 5.05e11 cell-updates/s at 4096^2, profiles/ab_r2.md; removed.)
    const int want_lo = own_lo > 0 ? trim : 0, want_hi = own_hi < nrows ? trim : 0;
    PassGeom best;
    std::memset(&best, 0, sizeof best);
    double best_cost = 1e300;
    const int HX = round_up(2 * T, 2 * kPairs), HY = 2 * T;  // a thread's columns are all inside or all outside the halo
    const int own = own_hi - own_lo;
    const int reg_threads = 65536 / 102 / 32 * 32;  // <= 102 registers per thread (launch bounds): 640 threads per SM
    for (int WS = force_ws ? force_ws : 32; WS <= (force_ws ? force_ws : 2048); WS += 32) {
        if (WS % (2 * kPairs)) break;
        const int Wout = WS - 2 * HX;
        if (Wout < 4) continue;
        const int NT = pass_threads(T, WS);
        if (NT > pass_max_threads(T) || (!force_ws && NT < 64 && WS < ncols + 2 * HX)) continue;
        const size_t smem = pass_smem_bytes(T, WS);
        if (smem > lim.smem_per_cta) continue;
        // do not take a strip much wider than the domain needs
        if (!force_ws && WS - 32 >= ncols + 2 * HX) continue;
        const int nstrips = (ncols + Wout - 1) / Wout;
        int occ = (int)(lim.smem_per_sm / (smem + 1024));
        if (occ > reg_threads / NT) occ = reg_threads / NT;
        if (occ > lim.max_threads_per_sm / NT) occ = lim.max_threads_per_sm / NT;
        if (occ < 1) occ = 1;
        if (occ > 8) occ = 8;
        const int slots = lim.num_sms * occ;
        for (int waves = 1; waves <= 3; waves++) {
            for (int slack = 0; slack < (force_chunks ? 1 : 3); slack++) {
                int nchunks = force_chunks ? force_chunks : (slots * waves) / nstrips - slack;
                if (nchunks < 1) nchunks = 1;
                if (nchunks > own) nchunks = own;
                if (!force_chunks) {  // keep the y overhead (halo + pipeline fill: 8T rows per chunk) bounded
                    const int max_chunks = own / (4 * T) > 1 ? own / (4 * T) : 1;
                    if (nchunks > max_chunks) nchunks = max_chunks;
                }
                int tl = want_lo, th = want_hi;
                int Hout = (own + tl + th + nchunks - 1) / nchunks;
                if (nchunks < 3 || Hout - (tl > th ? tl : th) < 4 * T) {  // too few / too small chunks to skew them
                    tl = th = 0;
                    Hout = (own + nchunks - 1) / nchunks;
                }
                nchunks = (own + tl + Hout - 1) / Hout;  // the last chunk absorbs the remainder
                const long ctas = (long)nstrips * nchunks;
                const long nwaves = (ctas + slots - 1) / slots;
                const int steps = Hout + 2 * HY + kSkew * T + 6;  // halo + pipeline fill + fixed per-CTA start-up cost
                // CTAs sharing the busiest SM in the last (or only) wave, and its resident warps
                const long in_last = ctas - (nwaves - 1) * slots;
                const int resident = (int)((in_last + lim.num_sms - 1) / lim.num_sms);
                const double warps = resident * NT / 32.0;
                const double eff = std::pow(warps >= 24.0 ? 1.0 : warps / 24.0, 0.7);
                const double wave_cost = (double)steps * WS * T;
                const double full_warps = occ * NT / 32.0;
                const double full_eff = std::pow(full_warps >= 24.0 ? 1.0 : full_warps / 24.0, 0.7);
                const double cost = (nwaves - 1) * wave_cost * occ / full_eff + wave_cost * resident / eff;
                if (cost < best_cost) {
                    best_cost = cost;
                    best.WS = WS; best.HX = HX; best.Wout = Wout; best.Hout = Hout; best.HY = HY;
                    best.nstrips = nstrips; best.nchunks = nchunks; best.trim_lo = tl; best.trim_hi = th;
                }
            }
            if (force_chunks) break;
        }
    }
    best.nrows = nrows; best.ncols = ncols; best.ld = ld; best.grow0 = grow0; best.gnrows = gnrows;
    best.own_lo = own_lo; best.own_hi = own_hi;
    return best;
}

}  // namespace cnv
