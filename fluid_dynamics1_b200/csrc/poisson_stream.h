// poisson_stream.h -- temporally blocked red-black SOR: the per-thread phases of the streaming
// kernel, shared between the CUDA kernel (poisson.cu) and the host-side schedule checker
// (tests/emul/stream_emul.cc compiles these same functions with g++ and runs the "threads" of a
// CTA in a loop, so the tiling / ring-buffer / skew logic is tested without a GPU).
//
// Algorithm (restates src/poisson.c:238-262, the OpenMP red-black sweep, T sweeps per HBM pass)
// ---------------------------------------------------------------------------------------------
// One CTA owns a strip of Wout output columns x Hout output rows and streams rows bottom-up
// through a ring buffer of R = kSkew*T + kPrefetch rows in shared memory.  Each row lives in column-parity
// split form (SE = even columns, SO = odd columns) so that every access of a half-row update is
// unit stride: for an even-column cell (pair k) N/S are SE[q+-1][k], E/W are SO[q][k], SO[q][k-1].
// The 2T half-sweeps (red L1, black L1, red L2, ..., black LT) are skewed: at step r level g updates
// its red row r-1-kSkew*g and its black row two rows below.  All inputs of every stage were produced in
// EARLIER steps, so the 2T stages of one step are independent and one __syncthreads per step suffices;
// the update is in place (same dependency structure as the reference's in-place sweeps, only the order
// of independent cell updates changes, so every cell sees bit-identical operands).  A row is final once
// the last level's black stage has processed it, and is written back from registers in that same step.
// Dependencies reach 2 cells per sweep, so strips/chunks overlap by 2T halo cells whose (stale-neighbour)
// results are discarded: the first/last loaded row is never updated and the first/last loaded column
// sees a pad value; the region of influence of either stays inside the halo.
// Thread (g, kk) of the CTA owns level g+1 (its red and black stage) and 2*kPairs adjacent columns,
// so its |u - u0| contributions all belong to sweep g+1: one accumulator.
#pragma once
#include "exact.h"

namespace cnv {

constexpr int kPrefetch = 3;  // rows in flight ahead of the compute front (cp.async groups)
// Rows between consecutive levels; must be even (compile-time colour parity).  4 is the minimum (red + black stage,
// 2 rows each): every operand of step r+1 except the row above the red row is final during step r and is loaded one
// step ahead (software pipelining).  With 6 that row could be prefetched too: measured slower on B200 (larger ring,
// fewer threads).
constexpr int kSkew = 4;
constexpr int kLand = kSkew > 4 ? 1 : 0;  // input rows must have landed this many steps early
// Column pairs per thread (a thread owns 2*kPairs adjacent columns).  More pairs = more independent
// cell updates in flight per thread (fp64 latency) and less per-cell overhead, at the price of registers.
constexpr int kPairs = 2;  // 4 measured slower on B200 (128 registers + spills, 12 warps/SM: 47 vs 36 us/sweep at 4096^2)
static_assert(kPairs % 2 == 0, "pairs are moved as 16-byte vectors");
constexpr int ring_rows(int T) { return kSkew * T + kPrefetch - (kSkew > 4 ? 1 : 0); }
struct PassGeom {
    // local array: nrows x ncols doubles with pitch ld; row 0 is global row grow0 of gnrows
    int nrows, ncols, ld, grow0, gnrows;
    int own_lo, own_hi;  // local rows [own_lo, own_hi) are produced (stored) by this device
    // tiling
    int WS;    // strip width held in shared memory (multiple of 4)
    int HX;    // x halo (multiple of 4, >= 2T)
    int Wout;  // WS - 2*HX
    int Hout;  // output rows per chunk
    // The first / last chunk is trim_lo / trim_hi rows shorter than the others (negative: taller).  Next to a neighbour
    // slab its CTAs wait for halo rows at the start and push their boundary rows at the end, so a little less streaming
    // work keeps them off the critical path of the pass; at the domain boundary there are no halo rows to stream on
    // that side, so the chunk takes 2T rows more.
    int trim_lo, trim_hi;
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
    G.y0 = p.own_lo + by * p.Hout - (by > 0 ? p.trim_lo : 0);
    G.y1 = p.own_lo + (by + 1) * p.Hout - p.trim_lo;
    if (G.y1 > p.own_hi || by == p.nchunks - 1) G.y1 = p.own_hi;
    G.ylo = G.y0 - p.HY > 0 ? G.y0 - p.HY : 0;
    G.yhi = G.y1 - 1 + p.HY < p.nrows - 1 ? G.y1 - 1 + p.HY : p.nrows - 1;
    return G;
}

// shared-memory layout of one ring slot: [SE | SO | PE | PO], each WP+4 doubles (2 pad each side)
CNV_HD int slot_stride(int WS) { return 4 * (WS / 2 + 4); }
CNV_HD int arr_off(int WS, int which) { return which * (WS / 2 + 4) + 2; }  // 0 SE, 1 SO, 2 PE, 3 PO

struct ThreadCtx {
    int g;        // level index (sweep g+1 of the pass)
    int k0;       // first of this thread's kPairs column pairs
    int vmask;    // bit 2p / 2p+1: the even / odd column of pair k0+p is updatable
    bool colown;  // the thread's columns belong to the strip's output range
};

CNV_HD ThreadCtx thread_ctx(const PassGeom &p, const CtaGeom &G, int tid)
{
    ThreadCtx t;
    const int TPG = p.WS / (2 * kPairs);
    t.g = tid / TPG;
    const int kk = tid - t.g * TPG;
    t.k0 = kPairs * kk;
    // A column is protected from updates only if it is a Dirichlet ring column or lies outside the domain.
    // The first/last column held in shared memory (halo edge of an interior strip) IS updated, from a pad
    // value: whatever that produces spreads by at most 2 columns per sweep, i.e. stays inside the 2T-column
    // halo whose results are discarded anyway -- exactly as far as the staleness of a frozen edge would reach.
    t.vmask = 0;
    for (int i = 0; i < 2 * kPairs; i++) {
        int gc = G.gx0 + 2 * t.k0 + i;
        if (gc >= 1 && gc <= p.ncols - 2) t.vmask |= 1 << i;
    }
    t.colown = 2 * t.k0 >= p.HX && 2 * t.k0 < p.HX + p.Wout;
    return t;
}

// ---- shared-memory accessors.  All ring offsets are BYTE offsets.  On the device they already include
// the CTA's shared window base, so an access is one 32-bit add + ld.shared/st.shared with no generic
// address arithmetic; on the host (schedule checker) they are offsets from the emulated array.
struct dbl2 { double x, y; };
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int smem_base(const double *sm) { return (int)__cvta_generic_to_shared(sm); }
__device__ __forceinline__ void cp_async8(const double *, int dst, const double *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ dbl2 lds2(const double *, int off)
{
    dbl2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(off));
    return v;
}
__device__ __forceinline__ double lds1(const double *, int off)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(off));
    return v;
}
__device__ __forceinline__ void sts2(double *, int off, double a, double b)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(off), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void sts1(double *, int off, double a)
{
    asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(off), "d"(a) : "memory");
}
__device__ __forceinline__ void stg2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
__device__ __forceinline__ dbl2 ldg2(const double *p)
{
    const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));  // L2 only: the other buffer is written every pass
    return {v.x, v.y};
}
// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): four consecutive doubles, 32-byte aligned
struct dbl4 { double x, y, z, w; };
__device__ __forceinline__ dbl4 ldg4(const double *p)
{
    dbl4 v;
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
#else
struct dbl4 { double x, y, z, w; };
inline dbl4 ldg4(const double *p) { return {p[0], p[1], p[2], p[3]}; }
inline void stg4(double *p, double a, double b, double c, double d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
inline int smem_base(const double *) { return 0; }
inline double *at(const double *sm, int off) { return reinterpret_cast<double *>(reinterpret_cast<char *>(const_cast<double *>(sm)) + off); }
inline void cp_async8(const double *sm, int dst, const double *src) { *at(sm, dst) = *src; }
inline void cp_async_commit() {}
inline dbl2 lds2(const double *sm, int off) { return {at(sm, off)[0], at(sm, off)[1]}; }
inline double lds1(const double *sm, int off) { return *at(sm, off); }
inline void sts2(double *sm, int off, double a, double b) { at(sm, off)[0] = a; at(sm, off)[1] = b; }
inline void sts1(double *sm, int off, double a) { *at(sm, off) = a; }
inline void stg2(double *p, double a, double b) { p[0] = a; p[1] = b; }
inline dbl2 ldg2(const double *p) { return {p[0], p[1]}; }
#endif

// true if the predicate holds for any lane of the calling warp (host schedule checker: the thread itself)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ bool warp_any(bool x) { return __any_sync(__activemask(), x) != 0; }
#else
inline bool warp_any(bool x) { return x; }
#endif
// The value `v` of the lane `delta` above (+1) / below (-1) this one.  Device only: the host schedule checker runs one
// thread at a time and takes the shared-memory form of the same load (stream_step).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double lane_shift(unsigned mask, double v, int up)
{
    return up ? __shfl_down_sync(mask, v, 1) : __shfl_up_sync(mask, v, 1);
}
#endif

// kPairs doubles of one parity array (pairs k0 .. k0+kPairs-1), moved as 16-byte vectors
struct vecP { double v[kPairs]; };
CNV_HD vecP ldsP(const double *sm, int off)
{
    vecP r;
#pragma unroll
    for (int i = 0; i < kPairs; i += 2) { const dbl2 t = lds2(sm, off + 8 * i); r.v[i] = t.x; r.v[i + 1] = t.y; }
    return r;
}
CNV_HD void stsP(double *sm, int off, const vecP &a)
{
#pragma unroll
    for (int i = 0; i < kPairs; i += 2) sts2(sm, off + 8 * i, a.v[i], a.v[i + 1]);
}
CNV_HD vecP zeroP()
{
    vecP r;
#pragma unroll
    for (int i = 0; i < kPairs; i++) r.v[i] = 0.0;
    return r;
}

// Per-thread state of the streaming pass.  Everything that does not change from step to step is
// computed once here, so that one step costs only the loads/stores, the arithmetic and a handful
// of pointer increments: ring-slot offsets advance incrementally (no modulo in the loop), global
// addresses advance by one pitch per step, and the colour type of a stage is a compile-time
// parameter of the step (the step loop is unrolled by two).
//
// Register-resident column history.  At step r the thread's red row is q = r-1-dq and its black row
// q-2.  Of the operands of the two updates only four come from shared memory: the row above the red
// row (N), the red cells themselves (own), the two right-hand sides, plus one neighbour value per
// colour that belongs to the adjacent thread (x).  Everything else is a value this same thread
// loaded or produced 1..3 steps ago and still holds:
//     red:   E/W pair b = N(r-1),  S = N(r-2)
//     black: own = N(r-3),  N = red result(r-1),  E/W pair = red result(r-2),  S = red result(r-3)
// (the cells in question are not touched by any other stage in between, see the skew argument above).
template <int T>
struct StreamThread {
    static constexpr int R = ring_rows(T);
    static constexpr int NCH = (2 * kPairs + T - 1) / T;  // (psi|rhs) column pairs this thread copies per row
    // uniform
    int ss, ringend;     // slot stride (bytes); end of the ring (bytes, base included)
    int ybase, rend;     // first / last step; ring slot of row q is (q - ybase) mod R
    int ylo, yhi, y0, y1;
    int vlo, vhi;        // rows that may be updated: inside the streamed range and off the Dirichlet ring
    int ld;
    int lslot;           // byte offset of the slot of the row being loaded (r + kPrefetch)
    // load
    const double *lptr[NCH];  // global address of this thread's chunk in the row being loaded
    int ldE[NCH], ldO[NCH];   // destination byte offsets inside a slot
    bool lcopy[NCH], lzero[NCH];
    // write back (threads of the last level only): element offset of the thread's first column in the black row
    long long sdst;
    bool sact;
    // compute
    int g, dq, k0;
    // byte offsets of the slots of rows qtop, qtop-1, qtop-2, qtop-3 (qtop = r - dq).  Like the histories below the
    // four entries are a rotating window: the step loop is unrolled by four and in phase PH (= step mod 4) the
    // offset of row qtop-j lives in o[(j - PH) & 3], so that moving down one row renames instead of copying
    int o[4];
    int aSE, aSO, aPE, aPO;  // array byte offsets inside a slot, k0 folded in
    int vmask;           // bit 2p / 2p+1: even / odd column of pair k0+p updatable
    bool colown, allvalid;
    bool mask_path;      // warp-uniform: some lane of this warp owns a column that must not be updated (see stream_step)
    int int_lo; unsigned int_span;  // steps r in [int_lo, int_lo + int_span] have both rows updatable AND inside [y0, y1)
    // h[(PH - a) & 3] = N loaded a steps ago, rr[(PH - a) & 3] = red results of a steps ago (a = 0: this step);
    // compile-time indices, so both stay in registers and no register moves are needed to age them
    vecP h[4], rr[4];
    vecP pf_N, pf_own, pf_Pr, pf_Pb;  // operands of THIS step, loaded during the previous step (software pipelining)
    double pf_x, pf_xb;
    // the two cells beside the thread's columns come from the adjacent lane's registers (stream_step); lanes whose
    // neighbour sits in another warp or another level group load them from shared memory as before
    unsigned wmask;          // the lanes of this warp that exist
    bool nb_lo_far, nb_hi_far;
    double acc;
};

// advance a ring offset by one slot; `ringend` = base + R*ss, so wrapping subtracts R*ss
template <int T>
CNV_HD int wrap_inc_t(int off, int ss, int ringend) { off += ss; return off >= ringend ? off - StreamThread<T>::R * ss : off; }

template <int T>
CNV_HD void stream_init(StreamThread<T> &s, const PassGeom &p, const CtaGeom &G, const double *sm, const double *in,
                        const double *rhs, int tid, int nthreads)
{
    constexpr int R = StreamThread<T>::R;
    const int WP = p.WS >> 1;
    const int base = smem_base(sm);
    s.ss = slot_stride(p.WS) * 8;
    s.ringend = base + R * s.ss;
    // start on a step whose stage-0 row has even (global row + 1): the colour type then alternates
    // with the step parity (PAR in stream_step)
    // (with full operand prefetch the stream starts kLand steps early: the operands of the first streamed row
    // are loaded one step before they are used)
    s.ybase = G.ylo - kLand - ((p.grow0 + G.ylo - kLand - 1) & 1);
    s.rend = G.y1 + 2 + kSkew * (T - 1);  // the last level's black stage reaches row y1-1
    s.rend += (4 - ((s.rend - s.ybase + 1) & 3)) & 3;  // whole quads of steps (the extra ones find nothing to do)
    s.ylo = G.ylo; s.yhi = G.yhi; s.y0 = G.y0; s.y1 = G.y1;
    s.vlo = G.ylo + 1 > 1 - p.grow0 ? G.ylo + 1 : 1 - p.grow0;
    s.vhi = G.yhi - 1 < p.gnrows - 2 - p.grow0 ? G.yhi - 1 : p.gnrows - 2 - p.grow0;
    s.ld = p.ld;
    s.lslot = base + (kPrefetch % R) * s.ss;
    for (int j = 0; j < StreamThread<T>::NCH; j++) {
        const int c = tid + j * nthreads;
        s.lcopy[j] = s.lzero[j] = false; s.lptr[j] = in; s.ldE[j] = 0; s.ldO[j] = 0;
        if (c < p.WS) {
            const bool isP = c >= WP;
            const int k = isP ? c - WP : c;
            const int gc = G.gx0 + 2 * k;
            s.ldE[j] = (arr_off(p.WS, isP ? 2 : 0) + k) * 8;
            s.ldO[j] = (arr_off(p.WS, isP ? 3 : 1) + k) * 8;
            const bool inside = gc >= 0 && gc < p.ld;
            s.lcopy[j] = inside;
            s.lzero[j] = !inside;
            s.lptr[j] = (isP ? rhs : in) + ((long long)(s.ybase + kPrefetch) * p.ld + (inside ? gc : 0));
        }
    }
    const ThreadCtx t = thread_ctx(p, G, tid);
    s.g = t.g; s.dq = kSkew * t.g; s.k0 = t.k0;
    for (int j = 0; j < 4; j++) s.o[j] = base + ((((-s.dq - j) % R) + R) % R) * s.ss;
    s.aSE = (arr_off(p.WS, 0) + t.k0) * 8; s.aSO = (arr_off(p.WS, 1) + t.k0) * 8;
    s.aPE = (arr_off(p.WS, 2) + t.k0) * 8; s.aPO = (arr_off(p.WS, 3) + t.k0) * 8;
    s.vmask = t.vmask;
    s.allvalid = t.vmask == (1 << (2 * kPairs)) - 1;
    s.mask_path = warp_any(!s.allvalid);
    s.colown = t.colown;
    {
        const int TPG = p.WS / (2 * kPairs), kk = tid - t.g * TPG;
#if defined(__CUDA_ARCH__)
        s.wmask = __activemask();
#else
        s.wmask = 0;
#endif
        s.nb_lo_far = kk == 0 || (tid & 31) == 0;
        s.nb_hi_far = kk == TPG - 1 || (tid & 31) == 31;
    }
    s.int_lo = 0x7fffffff; s.int_span = 0;
    const int gc4 = G.gx0 + 2 * t.k0;  // first of this thread's columns
    s.sact = t.g == T - 1 && t.colown && gc4 >= 0 && gc4 < p.ld;
    s.sdst = (long long)(s.ybase - 3 - s.dq) * p.ld + gc4;
    for (int j = 0; j < 4; j++) s.h[j] = s.rr[j] = zeroP();
    s.pf_N = s.pf_own = s.pf_Pr = s.pf_Pb = zeroP();
    s.pf_x = s.pf_xb = 0.0;
    s.acc = 0.0;
}

// The steps in which this thread's red row q = r-1-dq and black row q-2 are both updatable and both inside the
// CTA's output rows [y0, y1) -- there the norm and the write-back need no range checks (nsw = sweeps of this pass)
template <int T>
CNV_HD void stream_set_sweeps(StreamThread<T> &s, int nsw)
{
    const int lo = s.vlo > s.y0 ? s.vlo : s.y0;               // black row q-2 >= lo
    const int hi = s.vhi < s.y1 - 1 ? s.vhi : s.y1 - 1;       // red row q <= hi
    const int r_lo = lo + 3 + s.dq, r_hi = hi + 1 + s.dq;
    if (s.g < nsw && r_hi >= r_lo) { s.int_lo = r_lo; s.int_span = (unsigned)(r_hi - r_lo); }
    else { s.int_lo = 0x7fffffff; s.int_span = 0; }
}

// copy this thread's chunks of one row into the ring slot at byte offset `slot`; `back` = how many rows
// before the row lptr[] currently points at (prologue only)
template <int T>
CNV_HD void stream_load_row(const StreamThread<T> &s, double *sm, int slot, long long back_elems)
{
#pragma unroll
    for (int j = 0; j < StreamThread<T>::NCH; j++) {
        if (s.lcopy[j]) {
            const double *g = s.lptr[j] - back_elems;
            cp_async8(sm, slot + s.ldE[j], g);
            cp_async8(sm, slot + s.ldO[j], g + 1);
        }  // (columns outside the array: zero-filled once for all ring slots, stream_prezero)
    }
}

// Columns outside the local array (strips hanging over the domain edge) hold zeros in every ring slot for the
// whole pass -- nobody ever stores anything else there (their update mask is 0, so a store writes back the loaded 0) --
// so they are filled once here instead of once per streamed row.  Call before stream_prologue; the first CTA barrier of
// the step loop orders these stores before any read.
template <int T>
CNV_HD void stream_prezero(const StreamThread<T> &s, double *sm)
{
    const int base = s.lslot - (kPrefetch % StreamThread<T>::R) * s.ss;
#pragma unroll
    for (int j = 0; j < StreamThread<T>::NCH; j++) {
        if (s.lzero[j]) {
            for (int q = 0; q < StreamThread<T>::R; q++) {
                sts1(sm, base + q * s.ss + s.ldE[j], 0.0);
                sts1(sm, base + q * s.ss + s.ldO[j], 0.0);
            }
        }
    }
}

// rows ybase .. ybase+kPrefetch-1 (those inside the streamed range), one cp.async group each
template <int T>
CNV_HD void stream_prologue(const StreamThread<T> &s, double *sm)
{
    const int base = s.lslot - (kPrefetch % StreamThread<T>::R) * s.ss;
    for (int i = 0; i < kPrefetch; i++) {
        const int rl = s.ybase + i;
        if (rl >= s.ylo && rl <= s.yhi) stream_load_row<T>(s, sm, base + i * s.ss, (long long)(kPrefetch - i) * s.ld);
        cp_async_commit();
    }
}

// One step of the stream (the caller has waited for row r and passed the CTA barrier):
// issue the copy of row r+kPrefetch, then update this thread's red row r-1-dq and black row r-3-dq;
// threads of the last level write the finished black row back to global memory.
// PH = step number mod 4 (selects the rotating register windows); PAR = PH & 1 = parity of (global red row):
// 0 -> the red cells of this step are the even-column cells.
// Loads and arithmetic are unconditional so that the four cell updates of a step interleave; validity
// (Dirichlet ring, halo edges, pipeline fill/drain) is applied by selects only on the slow path -- threads
// whose four columns are all updatable take the select-free path whenever both rows are updatable.
// One range test (int_lo / int_span, stream_set_sweeps) selects the "interior" body: both rows updatable and inside
// the CTA's output rows, so the norm and the write-back need no per-row checks; the L1 norm is accumulated as
// acc += (|d0| + |d1|) per colour (a 2-long dependent chain).  Measured on B200 against the per-row-checked body it
// replaced: 5.20e11 vs 4.99e11 cell-updates/s at 4096^2 (profiles/ab_r2.md).
constexpr int rot4(int j, int ph) { return (j - ph) & 3; }

template <int T, bool POW2, int PH>
CNV_HD void stream_step(StreamThread<T> &s, const RelaxConsts &rc, double *sm, double *__restrict__ out, int r, int nsw)
{
    // ---- load ----
    if (r + kPrefetch <= s.yhi) stream_load_row<T>(s, sm, s.lslot, 0);
    cp_async_commit();
#pragma unroll
    for (int j = 0; j < StreamThread<T>::NCH; j++) s.lptr[j] += s.ld;
    s.lslot = wrap_inc_t<T>(s.lslot, s.ss, s.ringend);

    // ---- relax ----
    constexpr bool typeR = (PH & 1) != 0;  // red cells are the odd-column cells
    constexpr int O0 = rot4(0, PH), O1 = rot4(1, PH), O3 = rot4(3, PH);  // slots of rows qtop, qtop-1, qtop-3
    constexpr int A0 = PH, A1 = (PH + 3) & 3, A2 = (PH + 2) & 3, A3 = (PH + 1) & 3;  // history of 0..3 steps ago
    const int aA = typeR ? s.aSO : s.aSE, aB = typeR ? s.aSE : s.aSO;
    const int qtop = r - s.dq, q = qtop - 1, qb = qtop - 3;
    // every operand but (with kSkew == 4) the row above the red row was loaded during the previous step (see the
    // end of this function); that row is written by the level below during the previous step and is read now
    s.h[A0] = kSkew > 4 ? s.pf_N : ldsP(sm, s.o[O0] + aA);
    const vecP &N = s.h[A0];
    const vecP own = s.pf_own, Pr = s.pf_Pr, Pb = s.pf_Pb;
    const double x = s.pf_x, xb = s.pf_xb;
    const vecP &b = s.h[A1], &S = s.h[A2];
    const vecP &ownb = s.h[A3], &Nb = s.rr[A1], &bb = s.rr[A2], &Sb = s.rr[A3];
    // even-column cell of pair k: W = odd[k-1], E = odd[k];  odd-column cell: W = even[k], E = even[k+1]
    // red cells (type typeR) have their E/W neighbours in b (+ x at the thread's edge); black cells (the other
    // type) have theirs in bb (+ xb)
    vecP &n = s.rr[A0];  // (overwrites the red results of four steps ago, which nobody needs any more)
    vecP m;
#pragma unroll
    for (int i = 0; i < kPairs; i++) {
        const double Wr = typeR ? b.v[i] : (i == 0 ? x : b.v[i - 1]);
        const double Er = typeR ? (i == kPairs - 1 ? x : b.v[i + 1]) : b.v[i];
        n.v[i] = relax<POW2>(N.v[i], S.v[i], Er, Wr, own.v[i], Pr.v[i], rc);
        const double Wb = typeR ? (i == 0 ? xb : bb.v[i - 1]) : bb.v[i];
        const double Eb = typeR ? bb.v[i] : (i == kPairs - 1 ? xb : bb.v[i + 1]);
        m.v[i] = relax<POW2>(Nb.v[i], Sb.v[i], Eb, Wb, ownb.v[i], Pb.v[i], rc);
    }
    if ((unsigned)(r - s.int_lo) <= s.int_span) {
        // interior: both rows updatable and inside the output rows -- no per-row range checks at all
        if (s.mask_path) {
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                const bool vr = (s.vmask >> (2 * i + (typeR ? 1 : 0))) & 1, vb = (s.vmask >> (2 * i + (typeR ? 0 : 1))) & 1;
                n.v[i] = vr ? n.v[i] : own.v[i];
                m.v[i] = vb ? m.v[i] : ownb.v[i];
            }
        }
        stsP(sm, s.o[O1] + aA, n);
        stsP(sm, s.o[O3] + aB, m);
        if (s.colown) {
            double dr = 0.0, db = 0.0;
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                dr = i == 0 ? fabs(xsub(n.v[i], own.v[i])) : xadd(dr, fabs(xsub(n.v[i], own.v[i])));
                db = i == 0 ? fabs(xsub(m.v[i], ownb.v[i])) : xadd(db, fabs(xsub(m.v[i], ownb.v[i])));
            }
            s.acc = xadd(xadd(s.acc, dr), db);
        }
        if (s.sact) {
            double *dst = out + s.sdst;
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                if (typeR) stg2(dst + 2 * i, m.v[i], bb.v[i]);
                else stg2(dst + 2 * i, bb.v[i], m.v[i]);
            }
        }
    } else {
        if (s.g < nsw && qb >= s.vlo && q <= s.vhi) {
            // steady state: both rows updatable (and inside the streamed range by construction).  Warps none of whose
            // lanes owns a protected column (Dirichlet ring, outside the domain) store unconditionally; a warp that has
            // such a lane applies the column mask in ALL its lanes, so it does not diverge into the general path below
            // (the first and last strip would otherwise run ~9 % longer than the others and set the pass time).
            if (s.mask_path) {
#pragma unroll
                for (int i = 0; i < kPairs; i++) {
                    const bool vr = (s.vmask >> (2 * i + (typeR ? 1 : 0))) & 1, vb = (s.vmask >> (2 * i + (typeR ? 0 : 1))) & 1;
                    n.v[i] = vr ? n.v[i] : own.v[i];
                    m.v[i] = vb ? m.v[i] : ownb.v[i];
                }
            }
            stsP(sm, s.o[O1] + aA, n);
            stsP(sm, s.o[O3] + aB, m);
        } else {
            const bool active = s.g < nsw;
            const bool rowr = active && q >= s.vlo && q <= s.vhi;
            const bool rowb = active && qb >= s.vlo && qb <= s.vhi;
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                const bool vr = (s.vmask >> (2 * i + (typeR ? 1 : 0))) & 1, vb = (s.vmask >> (2 * i + (typeR ? 0 : 1))) & 1;
                n.v[i] = (rowr & vr) ? n.v[i] : own.v[i];
                m.v[i] = (rowb & vb) ? m.v[i] : ownb.v[i];
            }
            if (q >= s.ylo && q <= s.yhi) stsP(sm, s.o[O1] + aA, n);
            if (qb >= s.ylo && qb <= s.yhi) stsP(sm, s.o[O3] + aB, m);
        }
        // L1 update norm of this level (non-updated cells contribute exactly 0)
        if (s.colown) {
            // same association as the interior body: acc + (|d0| + |d1|) per colour; acc + 0.0 is exact
            double dr = 0.0, db = 0.0;
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                dr = i == 0 ? fabs(xsub(n.v[i], own.v[i])) : xadd(dr, fabs(xsub(n.v[i], own.v[i])));
                db = i == 0 ? fabs(xsub(m.v[i], ownb.v[i])) : xadd(db, fabs(xsub(m.v[i], ownb.v[i])));
            }
            s.acc = xadd(s.acc, (q >= s.y0 && q < s.y1) ? dr : 0.0);
            s.acc = xadd(s.acc, (qb >= s.y0 && qb < s.y1) ? db : 0.0);
        }
        // ---- write back: the black row of the last level is final; its red cells are r2 ----
        if (s.sact && qb >= s.y0 && qb < s.y1) {
            double *dst = out + s.sdst;
#pragma unroll
            for (int i = 0; i < kPairs; i++) {
                if (typeR) stg2(dst + 2 * i, m.v[i], bb.v[i]);  // black cells are the even-column cells
                else stg2(dst + 2 * i, bb.v[i], m.v[i]);
            }
        }
    }
    s.sdst += s.ld;
    // rows move up by one: the windows rotate by renaming (PH -> PH+1); only the new top slot is computed
    constexpr int PN = (PH + 1) & 3;
    constexpr int N0 = rot4(0, PN), N1 = rot4(1, PN), N3 = rot4(3, PN);
    s.o[N0] = wrap_inc_t<T>(s.o[O0], s.ss, s.ringend);
    // ---- software pipelining: operands of the NEXT step that nobody writes during this one ----
    // next red row = this step's row qtop (its red cells carry the level below since >= 3 steps, its black
    // cells since the previous step; rhs is static); next black row = this step's row q-1 (its red cells were
    // written by this level during the previous step).  The colour type flips with the step parity.
    {
        constexpr bool tN = !typeR;
        const int nA = tN ? s.aSO : s.aSE, nB = tN ? s.aSE : s.aSO;
        const int nPA = tN ? s.aPO : s.aPE, nPB = tN ? s.aPE : s.aPO;
        if (kSkew > 4) s.pf_N = ldsP(sm, s.o[N0] + nA);
        s.pf_own = ldsP(sm, s.o[N1] + nA);
        s.pf_Pr = ldsP(sm, s.o[N1] + nPA);
        s.pf_Pb = ldsP(sm, s.o[N3] + nPB);
        // The cell beside the thread's columns, in the next red row (x) and the next black row (xb).  Both are values the
        // ADJACENT thread of this level holds in registers: x = its N of this step (row qtop, same array, nobody writes it
        // during this step), xb = its red results of the previous step (row q-1; what it stored is what it kept, masks
        // included).  Two shuffles each instead of a shared-memory load whose 16-byte lane stride costs two extra
        // wavefronts per warp -- 8.0 M of the 53.6 M wavefronts per launch on the pipe that bounds the kernel
        // (profiles/ncu_pass_r2.md).  Lanes at a warp or level-group boundary take the load.
#if defined(__CUDA_ARCH__)
        s.pf_x = lane_shift(s.wmask, tN ? N.v[0] : N.v[kPairs - 1], tN);
        s.pf_xb = lane_shift(s.wmask, tN ? s.rr[A1].v[kPairs - 1] : s.rr[A1].v[0], !tN);
        if (tN ? s.nb_hi_far : s.nb_lo_far) s.pf_x = lds1(sm, s.o[N1] + nB + (tN ? 8 * kPairs : -8));
        if (tN ? s.nb_lo_far : s.nb_hi_far) s.pf_xb = lds1(sm, s.o[N3] + nA + (tN ? -8 : 8 * kPairs));
#else
        s.pf_x = lds1(sm, s.o[N1] + nB + (tN ? 8 * kPairs : -8));
        s.pf_xb = lds1(sm, s.o[N3] + nA + (tN ? -8 : 8 * kPairs));
#endif
    }
}

// ---- multi-GPU peer-memory path ------------------------------------------------------------------
// Every rank owns a PeerMailbox in its device memory that the other ranks write over NVLink (CUDA IPC
// mappings).  All counters are cumulative over the life of the communicator (never reset), indexed by the
// GLOBAL pass number g (every launched pass, no-ops included, counts on every rank):
//   halo_count[0|1]  CTAs of the lower|upper neighbour that have finished pushing their boundary rows into my
//                    low|high halo: after pass g it is (g+1) x (pushing CTAs of that neighbour)
//   ready[0|1]       the lower|upper neighbour has finished re-initialising its iterate buffers for solve epoch e:
//                    pass 0 of a solve must not push into a buffer the neighbour is still about to zero
//   norm_flag[r]     g+1 once rank r has published the per-sweep norms of pass g into norms[g % kNormSlots][r][*]
//                    (a no-op pass raises the flag without touching the norms)
// Norm slots.  A rank that reads the norms of pass g-1 can see remote publications of passes g and g+1 at most (a working
// pass g+2 needs this rank's flag of pass g+1).  Across solves a rank that is two hops away from a finished rank may still
// be reading the last slot of the old solve while the finished rank's first pass of the NEXT solve publishes; its later
// passes need the neighbours' re-initialisation, which in turn waits for everybody's old passes.  Eight slots keep all of
// those apart with room to spare.
constexpr int kMaxRanks = 8;
constexpr int kNormSlots = 8;
struct PeerMailbox {
    unsigned long long halo_count[2];
    unsigned long long ready[2];  // epoch of the lower|upper neighbour's latest (re-)initialised iterate buffers
    unsigned long long norm_flag[kMaxRanks];
    double norms[kNormSlots][kMaxRanks][8];
    unsigned long long error;  // a spin-wait timed out (a peer died): the host aborts
    unsigned int ticket;       // local: last-CTA election of the pass kernel
    unsigned int pad_;
};
struct PoissonCtl;
struct PeerLinks {
    int enabled, rank, world;
    int pidx;                  // pass index since the last reset of the state machine
    unsigned long long gidx;   // global pass index
    unsigned long long epoch;  // number of (re-)initialisations of the iterate so far (identical on every rank)
    PeerMailbox *mail[kMaxRanks];
    double *down_buf[2], *up_buf[2];   // neighbours' iterate buffers (peer mappings), null at the ends
    long long down_delta, up_delta;    // element offset: my row -> its halo copy in the neighbour's array
    unsigned long long need_low, need_high;  // pushes per pass arriving in my low / high halo
    unsigned long long push_low, push_high;  // pushes per pass I make downwards / upwards
    PoissonCtl *ctlbuf;        // [2]: ctlbuf[p & 1] = state used by pass p
    // optional trace (tools/peer_trace.py): [pass][cta][6] globaltimer stamps of thread 0 -- start, state known,
    // halos landed, stream done, push done, exit; null in production
    unsigned long long *trace;
    int trace_passes;
    unsigned long long timeout_ns;  // bound of every spin-wait on a peer flag (CNV_PEER_TIMEOUT_MS)
};

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
    double hit_e;  // lagged decision: norm of the converged sweep while its "redo" pass is pending
    int nbuf;      // iterate buffers in rotation: 2 (ping-pong; 0 reads as 2) or 3 (lagged decision)
    int pad_;
};

CNV_HD int next_buf(const PoissonCtl &c, int b, int by = 1)
{
    if (c.nbuf != 3) return (b + by) & 1;
    b += by;
    return b >= 3 ? b - 3 : b;
}

CNV_HD int pass_sweeps(const PoissonCtl &c, int T)
{
    if (c.redo > 0) return c.redo;
    int left = c.itmax - c.sweeps;
    return left < T ? left : T;
}

// e[0..nsw-1]: global L1 update norms of the sweeps of the pass just finished
CNV_HD void decide(PoissonCtl &c, const double *e, int nsw, double *hist)
{
    // (e is only ever indexed by the unrolled loop counter: with a caller's register array nothing here touches local memory --
    // the on-chip kernel's service warp runs this between two passes)
    int hit = -1;
    double ehit = 0.0, elast = 0.0;
    const bool redo = c.redo > 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (s < nsw) {
            if (s == nsw - 1) elast = e[s];
            if (!redo && hit < 0) {
                if (hist) hist[c.sweeps + s] = e[s];
                if (e[s] < c.tol) { hit = s; ehit = e[s]; }
            }
        }
    }
    c.passes++;
    if (redo) {  // recomputation up to the converged sweep: done
        c.cur = next_buf(c, c.cur);
        c.sweeps += nsw;
        c.result_k = c.sweeps - 1;
        c.result_e = elast;
        c.last_e = elast;
        c.redo = 0;
        c.state = 1;
        return;
    }
    if (hit == nsw - 1) {
        c.cur = next_buf(c, c.cur);
        c.sweeps += nsw;
        c.result_k = c.sweeps - 1;
        c.result_e = ehit;
        c.last_e = ehit;
        c.state = 1;
    } else if (hit >= 0) {
        c.redo = hit + 1;  // `cur` untouched: the pass input is recomputed with hit+1 sweeps
        c.hit_e = ehit;
    } else {
        c.cur = next_buf(c, c.cur);
        c.sweeps += nsw;
        c.last_e = elast;
        c.result_k = c.sweeps - 1;
        c.result_e = elast;
        if (c.sweeps >= c.itmax) c.state = 2;
    }
}

// ---- lagged stop decision (persistent on-chip kernel, poisson_onchip.cu) -----------------------------------
// With the plain machine pass p needs the grid-wide norms of pass p-1 before it can start: each pass would be a
// rendezvous of all CTAs.  The lagged machine lets pass p start knowing only the norms of passes <= p-2 and run
// SPECULATIVELY as if pass p-1 did not converge.  Three iterate buffers rotate (pass p reads B[p%3], writes
// B[(p+1)%3]) so that whatever the late decision turns out to be, the data it needs is still intact:
//   * pass h holds the first sweep with e < tol, at its LAST sweep: the answer is its output B[(h+1)%3]; the
//     speculative pass h+1 only read it.  Passes >= h+2 are no-ops.
//   * ... at an EARLIER sweep s: pass h+2 recomputes s+1 sweeps from the input of pass h, B[h%3] (the speculative
//     pass h+1 wrote B[(h+2)%3]), into B[(h+2)%3], and is final -- its norms need not be awaited: they equal those
//     pass h reported (same input, same arithmetic, same summation order), so result_e = e_h[s].
// X_p below is the chain state a pass derives at its start: X_0 = X_1 = reset state, X_p = lag_fold(X_{p-1}, e_{p-2}).
// (Round 2 also ran this machine on the multi-GPU peer path: bit-identical, but no faster at 2 or 8 GPUs -- the per-pass
// norm wait it removes is ~2 us of a 530 us pass, profiles/scale_r2.md -- so the peer path keeps the plain machine.)

// X_{p-1}, norms of pass p-2 (ignored when the solve is finished or pass p-1 was the redo pass) -> X_p
CNV_HD void lag_fold(PoissonCtl &c, const double *e, int T, double *hist)
{
    if (c.state != 0) return;
    if (c.redo > 0) {  // pass p-1 was the redo pass: final (e = norms of the discarded speculative pass)
        c.passes++;
        c.cur = next_buf(c, c.cur, 2);
        c.sweeps += c.redo;
        c.result_k = c.sweeps - 1;
        c.result_e = c.hit_e;
        c.last_e = c.hit_e;
        c.redo = 0;
        c.state = 1;
        return;
    }
    decide(c, e, pass_sweeps(c, T), hist);
}

struct LagAction {
    int kind;  // 0 no-op, 1 run (speculative), 2 redo
    int in, out, nsw;
};

// what pass number `pidx` (since the reset) does, given X_pidx
CNV_HD LagAction lag_action(const PoissonCtl &c, int pidx, int T)
{
    LagAction a = {0, 0, 0, 0};
    if (c.state != 0) return a;
    if (c.redo > 0) {
        a.kind = 2; a.in = c.cur; a.out = next_buf(c, c.cur, 2); a.nsw = c.redo;
        return a;
    }
    // pass pidx-1 (if any) is in flight: it reads B[cur] and applies pass_sweeps(c) sweeps
    const int pending = pidx > 0 ? 1 : 0;
    const int base = c.sweeps + (pending ? pass_sweeps(c, T) : 0);
    int left = c.itmax - base;
    if (left > T) left = T;
    if (left <= 0) return a;  // the pass in flight reaches itmax
    a.kind = 1; a.in = next_buf(c, c.cur, pending); a.out = next_buf(c, c.cur, pending + 1); a.nsw = left;
    return a;
}

// ---- what a pass of the peer path does at its start (shared by the kernel and the CPU protocol simulation) ----
// `prev` = state the previous pass stored (ctlbuf[(pidx-1) & 1]; the reset state for pidx == 0); the pass folds the norms
// of pass pidx-1.
CNV_HD bool peer_needs_norms(const PoissonCtl &prev, int pidx) { return pidx >= 1 && prev.state == 0; }
// c (in: prev, out: the state of this pass, stored for the next one) and the pass' action.  e = the needed norms summed
// over the ranks in rank order (ignored unless `need`); bad = a rank's norm flag timed out.
CNV_HD LagAction peer_advance(PoissonCtl &c, const double *e, bool need, bool bad, int T, double *hist)
{
    if (need && bad) c.state = 3;
    else if (need) decide(c, e, pass_sweeps(c, T), hist);
    LagAction a;
    a.kind = c.state == 0 ? 1 : 0;
    a.in = c.cur; a.out = next_buf(c, c.cur); a.nsw = pass_sweeps(c, T);
    return a;
}

// Host-visible state after P passes were launched: X_P (chain) + the norms of pass P-1 where they count.
// Returns true if e_last (norms of pass P-1) is needed, i.e. the caller must wait for them first.
CNV_HD bool lag_final_needs_last(const PoissonCtl &xP, int P) { return P > 0 && xP.state == 0 && xP.redo == 0; }
CNV_HD void lag_final(PoissonCtl &xP, int P, const double *e_last, int T, double *hist)
{
    if (lag_final_needs_last(xP, P)) decide(xP, e_last, pass_sweeps(xP, T), hist);
}

}  // namespace cnv
