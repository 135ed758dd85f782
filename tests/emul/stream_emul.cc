// stream_emul.cc -- CPU schedule checker for the streaming Poisson kernel.  TEST ONLY.
//
// Compiles the kernel's own per-thread phase functions (fluid_dynamics1_b200/csrc/poisson_stream.h)
// for the host and executes one pass the way the GPU does: CTA by CTA, step by step, with the
// "threads" of a CTA run in a loop between the points where the kernel has a __syncthreads().
// It exists so that the tiling / ring buffer / stage skew / halo logic can be compared with the
// oracle in the CPU test suite (no GPU in the build container).  It is not linked into the
// product library and is never used as a fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../fluid_dynamics1_b200/csrc/poisson_plan.h"
#include "../../fluid_dynamics1_b200/csrc/poisson_onchip.h"

using namespace cnv;

template <int T, bool POW2>
static void run_pass(const PassGeom &p, const RelaxConsts &rc, const double *in, const double *rhs, double *out,
                     int nsw, double *norms)
{
    constexpr int R = ring_rows(T);
    const int NT = pass_threads(T, p.WS);
    std::vector<double> sm((size_t)R * slot_stride(p.WS));
    std::vector<StreamThread<T>> st(NT);
    for (int g = 0; g < T; g++) norms[g] = 0.0;
    for (int by = 0; by < p.nchunks; by++)
        for (int bx = 0; bx < p.nstrips; bx++) {
            // poison shared memory so that any read of a row that was never loaded shows up
            for (auto &x : sm) x = std::nan("");
            const CtaGeom G = cta_geom(p, bx, by);
            for (int t = 0; t < NT; t++) stream_init<T>(st[t], p, G, sm.data(), in, rhs, t, NT);
            for (int t = 0; t < NT; t++) { stream_set_sweeps<T>(st[t], nsw); stream_prezero<T>(st[t], sm.data()); }
            for (int t = 0; t < NT; t++) stream_prologue<T>(st[t], sm.data());
            for (int r = st[0].ybase; r <= st[0].rend; r += 4) {
                // --- barrier before every step ---
                for (int t = 0; t < NT; t++) stream_step<T, POW2, 0>(st[t], rc, sm.data(), out, r, nsw);
                for (int t = 0; t < NT; t++) stream_step<T, POW2, 1>(st[t], rc, sm.data(), out, r + 1, nsw);
                for (int t = 0; t < NT; t++) stream_step<T, POW2, 2>(st[t], rc, sm.data(), out, r + 2, nsw);
                for (int t = 0; t < NT; t++) stream_step<T, POW2, 3>(st[t], rc, sm.data(), out, r + 3, nsw);
            }
            for (int t = 0; t < NT; t++) norms[st[t].g] += st[t].acc;
        }
}

extern "C" {

// plan only: returns WS, HX, Wout, Hout, nstrips, nchunks, threads, smem bytes, trim_lo, trim_hi in out[10]
void emul_plan(int nrows, int ncols, int ld, int grow0, int gnrows, int own_lo, int own_hi, int T, int force_ws,
               int force_chunks, long *out)
{
    PlanLimits lim;
    PassGeom p = make_plan(nrows, ncols, ld, grow0, gnrows, own_lo, own_hi, T, lim, force_ws, force_chunks);
    out[0] = p.WS; out[1] = p.HX; out[2] = p.Wout; out[3] = p.Hout; out[4] = p.nstrips; out[5] = p.nchunks;
    out[6] = pass_threads(T, p.WS); out[7] = (long)pass_smem_bytes(T, p.WS);
    out[8] = p.trim_lo; out[9] = p.trim_hi;
}

// One pass of nsw <= T sweeps.  in/out: nrows x ld; rhs = pscale * f (prepared like the product's
// prep kernel).  mode: 0 auto (pow2 fast path when exact), 1 force the literal general sequence.
int emul_pass(int T, int nrows, int ncols, int ld, int grow0, int gnrows, int own_lo, int own_hi, int force_ws,
              int force_chunks, double dx, double dy, double beta, int mode, const double *in, const double *f,
              double *out, int nsw, double *norms)
{
    PlanLimits lim;
    PassGeom p = make_plan(nrows, ncols, ld, grow0, gnrows, own_lo, own_hi, T, lim, force_ws, force_chunks);
    if (p.WS == 0) return -1;
    RelaxConsts rc = make_relax_consts(dx, dy, beta);
    if (mode == 1) { rc.pow2 = 0; rc.pscale = rc.cf; }
    std::vector<double> rhs((size_t)nrows * ld);
    for (size_t i = 0; i < rhs.size(); i++) rhs[i] = rc.pscale * f[i];
#define RUN(TT)                                                                       \
    if (T == TT) {                                                                    \
        if (rc.pow2) run_pass<TT, true>(p, rc, in, rhs.data(), out, nsw, norms);      \
        else run_pass<TT, false>(p, rc, in, rhs.data(), out, nsw, norms);             \
        return 0;                                                                     \
    }
    RUN(1) RUN(2) RUN(4) RUN(6) RUN(8)
#undef RUN
    return -2;
}

// state machine: feed a sequence of per-pass norm vectors, observe the decisions
void emul_decide(int *ctl_ints, double *ctl_dbls, const double *e, int nsw, double *hist)
{
    PoissonCtl c;
    c.state = ctl_ints[0]; c.cur = ctl_ints[1]; c.sweeps = ctl_ints[2]; c.redo = ctl_ints[3]; c.itmax = ctl_ints[4];
    c.result_k = ctl_ints[5]; c.passes = ctl_ints[6]; c.ticket = 0;
    c.tol = ctl_dbls[0]; c.result_e = ctl_dbls[1]; c.last_e = ctl_dbls[2]; c.hit_e = 0.0; c.nbuf = 2; c.pad_ = 0;
    decide(c, e, nsw, hist);
    ctl_ints[0] = c.state; ctl_ints[1] = c.cur; ctl_ints[2] = c.sweeps; ctl_ints[3] = c.redo; ctl_ints[5] = c.result_k;
    ctl_ints[6] = c.passes;
    ctl_dbls[1] = c.result_e; ctl_dbls[2] = c.last_e;
}
// ---- lagged stop decision (poisson_stream.h lag_fold / lag_action / lag_final) ----
// ints: state, cur, sweeps, redo, itmax, result_k, passes, nbuf;  dbls: tol, result_e, last_e, hit_e
static PoissonCtl ctl_from(const int *i, const double *d)
{
    PoissonCtl c; std::memset(&c, 0, sizeof c);
    c.state = i[0]; c.cur = i[1]; c.sweeps = i[2]; c.redo = i[3]; c.itmax = i[4]; c.result_k = i[5]; c.passes = i[6]; c.nbuf = i[7];
    c.tol = d[0]; c.result_e = d[1]; c.last_e = d[2]; c.hit_e = d[3];
    return c;
}
static void ctl_to(const PoissonCtl &c, int *i, double *d)
{
    i[0] = c.state; i[1] = c.cur; i[2] = c.sweeps; i[3] = c.redo; i[4] = c.itmax; i[5] = c.result_k; i[6] = c.passes; i[7] = c.nbuf;
    d[0] = c.tol; d[1] = c.result_e; d[2] = c.last_e; d[3] = c.hit_e;
}
void emul_lag_fold(int *ints, double *dbls, const double *e, int T, double *hist)
{
    PoissonCtl c = ctl_from(ints, dbls);
    lag_fold(c, e, T, hist);
    ctl_to(c, ints, dbls);
}
// out: kind (0 no-op, 1 run, 2 redo), in, out, nsw
void emul_lag_action(const int *ints, const double *dbls, int pidx, int T, int *out)
{
    const LagAction a = lag_action(ctl_from(ints, dbls), pidx, T);
    out[0] = a.kind; out[1] = a.in; out[2] = a.out; out[3] = a.nsw;
}
// start of a pass in the protocol simulation (tests/test_lag_protocol.py).  lag == 0: the peer kernel's own preamble code
// (peer_needs_norms / peer_advance, plain machine: the pass folds the norms of pass pidx-1).  lag == 1: the lagged machine
// as the on-chip kernel drives it (lag_fold with the norms of pass pidx-2, then lag_action).
int emul_peer_needs_norms(const int *ints, const double *dbls, int pidx, int lag)
{
    const PoissonCtl c = ctl_from(ints, dbls);
    if (!lag) return peer_needs_norms(c, pidx) ? 1 : 0;
    return pidx >= 2 && c.state == 0 && c.redo == 0 ? 1 : 0;
}
void emul_peer_advance(int *ints, double *dbls, const double *e, int need, int bad, int pidx, int lag, int T, int *out)
{
    PoissonCtl c = ctl_from(ints, dbls);
    LagAction a;
    if (!lag) {
        a = peer_advance(c, e, need != 0, bad != 0, T, nullptr);
    } else {
        if (need && bad) c.state = 3;
        else if (pidx >= 2) lag_fold(c, e, T, nullptr);
        a = lag_action(c, pidx, T);
    }
    ctl_to(c, ints, dbls);
    out[0] = a.kind; out[1] = a.in; out[2] = a.out; out[3] = a.nsw;
}
// host-visible state after P launched passes; returns 1 if e_last (norms of pass P-1) was consumed
int emul_lag_final(int *ints, double *dbls, int P, const double *e_last, int T, double *hist)
{
    PoissonCtl c = ctl_from(ints, dbls);
    const bool need = lag_final_needs_last(c, P);
    lag_final(c, P, e_last, T, hist);
    ctl_to(c, ints, dbls);
    return need ? 1 : 0;
}
int emul_pass_sweeps(int sweeps, int redo, int itmax, int T)
{
    PoissonCtl c; std::memset(&c, 0, sizeof c);
    c.sweeps = sweeps; c.redo = redo; c.itmax = itmax;
    return pass_sweeps(c, T);
}


// ---- the persistent on-chip kernel (poisson_onchip.h): a whole solve with the kernel's own per-thread code, tile geometry,
// masks and lagged stop machine.  CTAs run one after the other inside a pass (compute + store), then all of them reload
// their halo cells -- the order the flags of the real kernel enforce.  bufs: 3 x nrows x ld (bufs[0] = initial iterate).
// plan_out: T, H, OW, OH, NPX, NPY, ntx, nty.  Returns 0, -1 (no plan).
}  // extern "C"
template <bool POW2>
static void run_onchip(const OnchipGeom &g, const RelaxConsts &rc, const double *rhs, double *bufs, PoissonCtl &ctl, double *hist)
{
    const int ncta = g.ntx * g.nty, NT = oc_threads(g);
    const size_t fld = (size_t)g.nrows * g.ld;
    double *B[3] = {bufs, bufs + fld, bufs + 2 * fld};
    std::vector<std::vector<double>> sm(ncta, std::vector<double>(kOcSlots * kOcPitch, 0.0));
    std::vector<std::vector<OcThread>> th(ncta, std::vector<OcThread>(NT));
    std::vector<OcTile> tiles(ncta);
    for (int c = 0; c < ncta; c++) {
        tiles[c] = oc_tile(g, c % g.ntx, c / g.ntx);
        for (int t = 0; t < NT; t++) {
            oc_thread_init(th[c][t], g, tiles[c], t, sm[c].data());
            oc_load_rhs(th[c][t], rhs);
            oc_load_psi(th[c][t], B[ctl.cur]);
            oc_publish_all(th[c][t], sm[c].data());
        }
    }
    std::vector<std::vector<double>> norms;  // [pass][8]
    PoissonCtl X = ctl;
    int P = 0;
    for (int p = 0;; p++) {
        const bool need = p >= 2 && X.state == 0 && X.redo == 0;
        if (p >= 2) {
            double e[8];
            for (int i = 0; i < 8; i++) e[i] = need ? norms[p - 2][i] : 0.0;
            lag_fold(X, e, g.T, hist);
        }
        const LagAction act = lag_action(X, p, g.T);
        P = p;
        if (act.kind == 0) break;
        std::vector<double> e(8, 0.0);
        for (int c = 0; c < ncta; c++) {
            double *s = sm[c].data();
            if (act.kind == 2)
                for (int t = 0; t < NT; t++) { oc_load_psi(th[c][t], B[act.in]); oc_publish_all(th[c][t], s); }
            for (int sw = 0; sw < act.nsw; sw++) {
                const int rem0 = 2 * (act.nsw - sw) - 1;
                double acc = 0.0;
                for (int col = 0; col < 2; col++)
                    for (int t = 0; t < NT; t++) {
                        if (th[c][t].dist > rem0 - col) continue;
                        double a = 0.0;
                        const bool sel = !th[c][t].fast;
                        if ((tiles[c].par0 ^ col) == 0) { if (sel) oc_half_sweep<POW2, 0, true>(th[c][t], rc, s, a); else oc_half_sweep<POW2, 0, false>(th[c][t], rc, s, a); }
                        else { if (sel) oc_half_sweep<POW2, 1, true>(th[c][t], rc, s, a); else oc_half_sweep<POW2, 1, false>(th[c][t], rc, s, a); }
                        if (th[c][t].own) acc += a;
                    }
                e[sw] += acc;
            }
            for (int t = 0; t < NT; t++) if (th[c][t].own) oc_store(th[c][t], B[act.out]);
        }
        norms.push_back(e);
        if (act.kind == 2) continue;
        for (int c = 0; c < ncta; c++)
            for (int t = 0; t < NT; t++)
                if (!th[c][t].own) { oc_load_psi(th[c][t], B[act.out]); oc_publish_all(th[c][t], sm[c].data()); }
    }
    if (lag_final_needs_last(X, P)) lag_final(X, P, norms[P - 1].data(), g.T, hist);
    ctl = X;
}
extern "C" {

int emul_onchip_solve(int nrows, int ncols, int ld, int T, int ntx, int nty, int num_sms, double dx, double dy, double beta, int mode,
                      const double *f, double *bufs, int itmax, double tol, int *ints, double *dbls, double *hist, long *plan_out)
{
    OnchipGeom g;
    if (!onchip_plan(nrows, ncols, ld, 0, nrows, 0, nrows, num_sms, &g, T, ntx, nty)) return -1;
    plan_out[0] = g.T; plan_out[1] = g.HX; plan_out[2] = g.OW; plan_out[3] = g.OH; plan_out[4] = g.NPX; plan_out[5] = g.NPY;
    plan_out[6] = g.ntx; plan_out[7] = g.nty;
    RelaxConsts rc = make_relax_consts(dx, dy, beta);
    if (mode == 1) { rc.pow2 = 0; rc.pscale = rc.cf; }
    std::vector<double> rhs((size_t)nrows * ld);
    for (size_t i = 0; i < rhs.size(); i++) rhs[i] = rc.pscale * f[i];
    PoissonCtl c; std::memset(&c, 0, sizeof c);
    c.state = itmax > 0 ? 0 : 2; c.itmax = itmax; c.result_k = -1; c.tol = tol; c.nbuf = 3;
    if (rc.pow2) run_onchip<true>(g, rc, rhs.data(), bufs, c, hist);
    else run_onchip<false>(g, rc, rhs.data(), bufs, c, hist);
    ctl_to(c, ints, dbls);
    return 0;
}

// Markstein division vs hardware division: returns the number of mismatches
long emul_check_div(double d, const double *a, long n)
{
    long bad = 0;
    const double rd = 1.0 / d;
    for (long i = 0; i < n; i++)
        if (xdiv_const(a[i], d, rd) != a[i] / d) bad++;
    return bad;
}
}
