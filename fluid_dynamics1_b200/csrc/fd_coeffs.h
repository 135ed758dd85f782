// fd_coeffs.h -- banded form of the reference's finite-difference operators.
//
// The reference builds dense n x n matrices (Diff1: src/finitediff.c:51-153, Diff2: :178-292) and
// applies them through Kronecker products and a dense mat-vec (src/main.c:149-152, :298-304).  Only
// <= 7 entries per row are non-zero.  FdTable holds those bands: one interior row plus the three
// low and three high closure rows, coefficients evaluated on the HOST with the reference's own
// expression shape  (double)num / den / h  resp.  (double)num / den / (h*h)  so that they are
// bit-identical, and fd_apply() accumulates  sum += c[t]*x[start+t]  in ascending column order
// from +0.0 with separately rounded products -- the order of the dense mat-vec loop
// (src/linearalg.c:258-263), whose out-of-band terms are +-0 and cannot change a finite sum.
#pragma once
#include "exact.h"

namespace cnv {

struct FdRow {
    double c[7];
    int cnt;
    int pad_;
};

struct FdTable {
    FdRow interior;  // row i, columns i-half .. i+half
    FdRow lo[3];     // rows 0,1,2          (columns 0 .. cnt-1)
    FdRow hi[3];     // rows n-1, n-2, n-3  (columns n-cnt .. n-1)
    int half;        // order / 2
    int n;
};

// x(col) returns the operand at position `col` along the differentiated axis
template <class Load>
CNV_HD double fd_apply(const FdTable &t, int i, Load x)
{
    const FdRow *r;
    int start;
    if (i < t.half) { r = &t.lo[i]; start = 0; }
    else if (i >= t.n - t.half) { r = &t.hi[t.n - 1 - i]; start = t.n - r->cnt; }
    else { r = &t.interior; start = i - t.half; }
    double sum = 0.0;
    for (int k = 0; k < r->cnt; k++) sum = xadd(sum, xmul(r->c[k], x(start + k)));
    return sum;
}

// ---- host side: table construction ----
namespace fd_detail {
struct Q { int num2; int den; };  // numerator stored doubled so that +-0.5 is representable
// (num2/2)/den/scale with the "/ den" omitted when den == 1, as the reference writes it
inline double ev(Q q, double scale) { double a = q.num2 / 2.0; return q.den == 1 ? a / scale : a / q.den / scale; }
}  // namespace fd_detail

// Fills the table for derivative `deriv` (1|2) of accuracy `order` (2|4|6) on n points with
// spacing h.  Returns false for any other order (the reference prints
// "** Error: valid orders are 2, 4 or 6 **" and exits, src/finitediff.c:150-151).
inline bool fd_make_table(int n, int order, int deriv, double h, FdTable *t)
{
    using fd_detail::Q;
    using fd_detail::ev;
    if (order != 2 && order != 4 && order != 6) return false;
    // closure / interior stencils by width.  first derivative: 2-pt one-sided, 3-pt, 5-pt, 7-pt central
    static const Q d1w2[] = {{-2, 1}, {2, 1}};
    static const Q d1w3[] = {{-1, 1}, {0, 1}, {1, 1}};
    static const Q d1w5[] = {{2, 12}, {-4, 3}, {0, 1}, {4, 3}, {-2, 12}};
    static const Q d1w7[] = {{-2, 60}, {6, 20}, {-6, 4}, {0, 1}, {6, 4}, {-6, 20}, {2, 60}};
    // second derivative: 4-pt one-sided, 3-pt, 5-pt, 7-pt central
    static const Q d2w4[] = {{4, 1}, {-10, 1}, {8, 1}, {-2, 1}};
    static const Q d2w3[] = {{2, 1}, {-4, 1}, {2, 1}};
    static const Q d2w5[] = {{-2, 12}, {8, 3}, {-10, 2}, {8, 3}, {-2, 12}};
    static const Q d2w7[] = {{2, 90}, {-6, 20}, {6, 2}, {-98, 18}, {6, 2}, {-6, 20}, {2, 90}};
    const double scale = deriv == 1 ? h : h * h;
    const Q *edge = deriv == 1 ? d1w2 : d2w4;
    const int edge_w = deriv == 1 ? 2 : 4;
    const Q *central[3] = {deriv == 1 ? d1w3 : d2w3, deriv == 1 ? d1w5 : d2w5, deriv == 1 ? d1w7 : d2w7};
    auto fill = [&](FdRow &r, const Q *q, int w, bool reversed) {
        r.cnt = w; r.pad_ = 0;
        for (int k = 0; k < 7; k++) r.c[k] = 0.0;
        for (int k = 0; k < w; k++) r.c[k] = ev(q[reversed ? w - 1 - k : k], scale);
    };
    t->n = n;
    t->half = order / 2;
    fill(t->interior, central[order / 2 - 1], order + 1, false);
    for (int c = 0; c < 3; c++) {
        // closure row c: one-sided for c == 0, else the central stencil of width 2c+1.
        // High rows copy the low-row values (src/finitediff.c:136-145, :273-284): in ascending column
        // order the first-derivative rows keep their sequence, the second-derivative rows are mirrored.
        const Q *q = c == 0 ? edge : central[c - 1];
        const int w = c == 0 ? edge_w : 2 * c + 1;
        fill(t->lo[c], q, w, false);
        fill(t->hi[c], q, w, deriv == 2);
    }
    return true;
}

// band of row i: first column and coefficients (used to build the dense Diff1/Diff2 of the drop-in API)
inline void fd_row(const FdTable &t, int i, int *start, const FdRow **row)
{
    if (i < t.half) { *row = &t.lo[i]; *start = 0; }
    else if (i >= t.n - t.half) { *row = &t.hi[t.n - 1 - i]; *start = t.n - (*row)->cnt; }
    else { *row = &t.interior; *start = i - t.half; }
}

}  // namespace cnv
