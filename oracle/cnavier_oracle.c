/*
 * cnavier_oracle.c -- CPU restatement of the cnavier hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (fluid_dynamics1_b200/csrc) never links it.
 *
 * Parity status: PINNED.  The restatement is validated (tests/test_oracle_*.py)
 *   (1) bit-for-bit against the unmodified reference sources compiled into
 *       oracle/_ref/libcnavier_ref.so (Diff1/Diff2 rows, kronecker+mtrxmul operator
 *       application, euler, poisson_SOR_log in both orderings, whole time steps), and
 *   (2) against the reference's shipped run logs (output/logs/testRun*.txt: per-step
 *       Poisson sweep counts and 7-digit residuals), committed as tests/golden/ (json).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  All fields are flat row-major double arrays A[i*ny + j]; i is the
 * first index of the reference's mtrx (A.M[i][j]), j the second.
 *
 * Arithmetic contract (what "bit-for-bit" relies on): every product and sum below is an
 * individually rounded IEEE-754 binary64 operation in the association order of the
 * reference expression.  Build with -ffp-contract=off and without -march=native so gcc
 * never fuses a*b+c (the reference x86-64 -O2 build has no FMA either).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265359 /* include/poisson.h:9 -- truncated on purpose */

/* ------------------------------------------------------------------------------------
 * Finite-difference rows.  src/finitediff.c:51-153 (Diff1) and :178-292 (Diff2).
 *
 * The reference fills a dense n x n matrix; only a band of <= 7 entries per row is
 * non-zero.  orc_diff_row returns that band for row i: first column `*start`, `*cnt`
 * coefficients in ascending column order.  Coefficients are evaluated with the
 * reference's own expressions: (double)num / den / h   (first derivative)
 *                              (double)num / den / (h*h) (second derivative)
 * where den == 1 entries are written in the reference without the "/ den".
 * The mirrored closure rows at the high end are copies of the low-end values
 * (src/finitediff.c:72-73, 99-103, 136-145, 201-204, 232-238, 273-284), reproduced by
 * indexing the same low-row tables.
 * ---------------------------------------------------------------------------------- */
typedef struct { double num, den; } frac;

static double coef1(frac c, double h) { return c.den == 1.0 ? c.num / h : c.num / c.den / h; }
static double coef2(frac c, double h) { return c.den == 1.0 ? c.num / (h * h) : c.num / c.den / (h * h); }

/* first derivative: closure rows 0,1,2 and the interior rows for orders 2,4,6 */
static const frac D1_ROW0[2] = {{-1, 1}, {1, 1}};                              /* :60-61, 79-80, 109-110 */
static const frac D1_ROW1[3] = {{-0.5, 1}, {0, 1}, {0.5, 1}};                  /* :83-85, 112-114 (and order-2 interior :66-68) */
static const frac D1_ROW2[5] = {{1, 12}, {-2, 3}, {0, 1}, {2, 3}, {-1, 12}};   /* :116-120 (and order-4 interior :91-95) */
static const frac D1_INT6[7] = {{-1, 60}, {3, 20}, {-3, 4}, {0, 1}, {3, 4}, {-3, 20}, {1, 60}}; /* :126-132 */
/* second derivative */
static const frac D2_ROW0[4] = {{2, 1}, {-5, 1}, {4, 1}, {-1, 1}};             /* :187-190, 210-213, 244-247 */
static const frac D2_ROW1[3] = {{1, 1}, {-2, 1}, {1, 1}};                      /* :216-218, 249-251 (order-2 interior :195-197) */
static const frac D2_ROW2[5] = {{-1, 12}, {4, 3}, {-5, 2}, {4, 3}, {-1, 12}};  /* :253-257 (order-4 interior :224-228) */
static const frac D2_INT6[7] = {{1, 90}, {-3, 20}, {3, 2}, {-49, 18}, {3, 2}, {-3, 20}, {1, 90}}; /* :263-269 */

/* returns 0 on success, -1 for an unsupported order (reference: exit(1), :150-151, :289-290) */
int orc_diff_row(int n, int o, int deriv, double h, int i, int *start, double *coef, int *cnt)
{
    if (o != 2 && o != 4 && o != 6) return -1;
    int half = o / 2;
    const frac *tab; int w, first; int mirror = 0;
    int lo = (i < half) ? i : -1;               /* low closure row index */
    int hi = (i >= n - half) ? (n - 1 - i) : -1;/* high closure row index (0 = last row) */
    int cls = lo >= 0 ? lo : hi;                /* which closure table */
    if (lo < 0 && hi >= 0) mirror = 1;
    if (cls < 0) {                              /* interior row */
        if (o == 2)      { tab = deriv == 1 ? D1_ROW1 : D2_ROW1; w = 3; }
        else if (o == 4) { tab = deriv == 1 ? D1_ROW2 : D2_ROW2; w = 5; }
        else             { tab = deriv == 1 ? D1_INT6 : D2_INT6; w = 7; }
        first = i - w / 2;
        for (int t = 0; t < w; t++) coef[t] = deriv == 1 ? coef1(tab[t], h) : coef2(tab[t], h);
        *start = first; *cnt = w; return 0;
    }
    if (cls == 0)      { tab = deriv == 1 ? D1_ROW0 : D2_ROW0; w = deriv == 1 ? 2 : 4; }
    else if (cls == 1) { tab = deriv == 1 ? D1_ROW1 : D2_ROW1; w = 3; }
    else               { tab = deriv == 1 ? D1_ROW2 : D2_ROW2; w = 5; }
    if (!mirror) {
        /* low rows: row 0 starts at col 0; rows 1,2 are centred and also start at col 0 */
        *start = 0; *cnt = w;
        for (int t = 0; t < w; t++) coef[t] = deriv == 1 ? coef1(tab[t], h) : coef2(tab[t], h);
        return 0;
    }
    /* high rows end at column n-1.  First derivative: D[n-1-c][n-1-k] = D[c][w-1-k]
     * (same sequence in ascending column order, :136-145).  Second derivative:
     * D[n-1-c][n-1-k] = D[c][k] (reversed sequence, :273-284). */
    *start = n - w; *cnt = w;
    for (int t = 0; t < w; t++) {
        frac c = deriv == 1 ? tab[t] : tab[w - 1 - t];
        coef[t] = deriv == 1 ? coef1(c, h) : coef2(c, h);
    }
    return 0;
}

/* Dense n x n operator exactly as Diff1/Diff2 return it (src/finitediff.c:51,178). */
int orc_diff_dense(int n, int o, int deriv, double h, double *D)
{
    memset(D, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; i++) {
        int s, c; double co[7];
        if (orc_diff_row(n, o, deriv, h, i, &s, co, &c)) return -1;
        for (int t = 0; t < c; t++) D[(size_t)i * n + s + t] = co[t];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Matrix-free operator application.  Restates the dense route
 *   reshape (src/linearalg.c:373-408) -> mtrxmul (:236-285) with DX = kron(I, d_x),
 *   DY = kron(d_y, I) (src/main.c:149-152, src/linearalg.c:353-371).
 * Effective arithmetic: out[i][j] = sum_k d[j][k]*A[i][k] (axis 1, "DX": along j) or
 * sum_k d[i][k]*A[k][j] (axis 0, "DY": along i), k ascending, accumulator starting at
 * +0.0, every product rounded separately (src/linearalg.c:258-263).  The zero entries
 * outside the band contribute +-0.0 and cannot change a finite running sum, so only the
 * band is visited; the leading "0.0 +" is kept so that a first product of -0.0 yields
 * +0.0 exactly as in the dense loop.
 * ---------------------------------------------------------------------------------- */
int orc_apply(const double *A, int nx, int ny, int axis, int deriv, int order, double h, double *out)
{
    int n = axis == 1 ? ny : nx;
    double *rows = (double *)malloc(sizeof(double) * 7 * (size_t)n);
    int *st = (int *)malloc(sizeof(int) * 2 * (size_t)n);
    if (!rows || !st) return -2;
    for (int r = 0; r < n; r++)
        if (orc_diff_row(n, order, deriv, h, r, &st[2 * r], &rows[7 * r], &st[2 * r + 1])) { free(rows); free(st); return -1; }
#pragma omp parallel for schedule(static) if (nx >= 128)
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++) {
            int r = axis == 1 ? j : i;
            int s = st[2 * r], c = st[2 * r + 1];
            const double *co = &rows[7 * r];
            double sum = 0.0;
            for (int t = 0; t < c; t++) {
                double a = axis == 1 ? A[(size_t)i * ny + s + t] : A[(size_t)(s + t) * ny + j];
                sum += co[t] * a;
            }
            out[(size_t)i * ny + j] = sum;
        }
    free(rows); free(st);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Poisson solver.  src/poisson.c:111-173 (poisson_SOR), :224-285 (poisson_SOR_log),
 * :62-109 / :176-222 (poisson / poisson_log = the same with beta == 1 and no u0 term),
 * convergence norm src/poisson.c:34-60.
 *
 *   redblack = 0: lexicographic in-place sweep (serial build, :264-270)
 *   redblack = 1: (i+j) even first, then odd (OpenMP build, :238-262)
 *   sor      = 0: Gauss-Seidel update  u = A / D                      (:85)
 *   sor      = 1: u = beta * A / D + (1 - beta) * u0                  (:148, :246)
 *   with A = dy*dy*(u[i+1][j]+u[i-1][j]) + dx*dx*(u[i][j+1]+u[i][j-1]) - dx*dx*dy*dy*f[i][j]
 *        D = 2*(dx*dx+dy*dy)
 * Zero initial guess, ring stays zero (:229-230).  After each sweep e = sum over the
 * WHOLE grid of sqrt(pow(u-u0,2)) (= |u-u0| for all values whose square neither
 * underflows nor overflows; fabs is used here) and the loop stops at the first e < tol.
 * Returns: 0 converged (*k_out = reference's logged k = sweeps-1, *e_out = e),
 *          1 itmax reached (reference: message + exit(1), :280-284).
 * hist (optional, length itmax) receives e after every sweep.
 * ---------------------------------------------------------------------------------- */
static inline double sor_cell(const double *u, const double *f, size_t p, int ny,
                              double dxx, double dyy, double cf, double D, double beta, double omb,
                              int sor, double u0)
{
    double A = dyy * (u[p + ny] + u[p - ny]) + dxx * (u[p + 1] + u[p - 1]) - cf * f[p];
    if (!sor) return A / D;
    return beta * A / D + omb * u0;
}

int orc_poisson(const double *f, int nx, int ny, double dx, double dy, int itmax, double tol,
                double beta, int redblack, int sor, double *u, int *k_out, double *e_out, double *hist)
{
    size_t ncell = (size_t)nx * ny;
    double *u0 = (double *)malloc(sizeof(double) * ncell);
    if (!u0) return -2;
    memset(u, 0, sizeof(double) * ncell);
    const double dxx = dx * dx, dyy = dy * dy;
    const double cf = dx * dx * dy * dy;          /* ((dx*dx)*dy)*dy, as the C expression parses */
    const double D = 2 * (dx * dx + dy * dy);
    const double omb = 1 - beta;
    int status = 1;
    for (int k = 0; k < itmax; k++) {
        memcpy(u0, u, sizeof(double) * ncell);    /* mtrxcpy(u0,u), :236 */
        if (redblack) {
            for (int colour = 0; colour < 2; colour++) {
#pragma omp parallel for schedule(static) if (nx > 128 && ny > 128)
                for (int i = 1; i < nx - 1; i++)
                    for (int j = 1 + ((i + 1 + colour) & 1); j < ny - 1; j += 2) {
                        size_t p = (size_t)i * ny + j;
                        u[p] = sor_cell(u, f, p, ny, dxx, dyy, cf, D, beta, omb, sor, u0[p]);
                    }
            }
        } else {
            for (int i = 1; i < nx - 1; i++)
                for (int j = 1; j < ny - 1; j++) {
                    size_t p = (size_t)i * ny + j;
                    u[p] = sor_cell(u, f, p, ny, dxx, dyy, cf, D, beta, omb, sor, u0[p]);
                }
        }
        double e = 0;
#pragma omp parallel for reduction(+ : e) schedule(static) if (nx > 128 && ny > 128)
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < ny; j++) e += fabs(u[(size_t)i * ny + j] - u0[(size_t)i * ny + j]);
        if (hist) hist[k] = e;
        if (e < tol) { *k_out = k; *e_out = e; status = 0; break; }
        *k_out = k; *e_out = e;
    }
    free(u0);
    return status;
}

/* K sweeps with no convergence test (throughput baseline; same update as above). */
void orc_poisson_sweeps(const double *f, int nx, int ny, double dx, double dy, int nsweeps,
                        double beta, double *u, double *norms)
{
    const double dxx = dx * dx, dyy = dy * dy, cf = dx * dx * dy * dy, D = 2 * (dx * dx + dy * dy), omb = 1 - beta;
    size_t ncell = (size_t)nx * ny;
    double *u0 = (double *)malloc(sizeof(double) * ncell);
    for (int k = 0; k < nsweeps; k++) {
        memcpy(u0, u, sizeof(double) * ncell);
        for (int colour = 0; colour < 2; colour++) {
#pragma omp parallel for schedule(static)
            for (int i = 1; i < nx - 1; i++)
                for (int j = 1 + ((i + 1 + colour) & 1); j < ny - 1; j += 2) {
                    size_t p = (size_t)i * ny + j;
                    u[p] = sor_cell(u, f, p, ny, dxx, dyy, cf, D, beta, omb, 1, u0[p]);
                }
        }
        double e = 0;
#pragma omp parallel for reduction(+ : e) schedule(static)
        for (size_t p = 0; p < ncell; p++) e += fabs(u[p] - u0[p]);
        if (norms) norms[k] = e;
    }
    free(u0);
}

/* ------------------------------------------------------------------------------------
 * Pointwise fluid-dynamics operators.  src/fluiddyn.c:71-102 (euler), :126-154
 * (continuity), :179-207 (vorticity: returns second argument minus first).
 * ---------------------------------------------------------------------------------- */
void orc_euler(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
               const double *u, const double *v, double Re, double dt, size_t ncell)
{
#pragma omp parallel for schedule(static) if (ncell >= 128 * 128)
    for (size_t p = 0; p < ncell; p++)
        w[p] = (-u[p] * dwdx[p] - v[p] * dwdy[p] + (1. / Re) * (d2wdx2[p] + d2wdy2[p])) * dt + w[p];
}
void orc_continuity(const double *dudx, const double *dvdy, double *out, size_t ncell)
{
    for (size_t p = 0; p < ncell; p++) out[p] = dudx[p] + dvdy[p];
}
void orc_vorticity(const double *a, const double *b, double *out, size_t ncell)
{
    for (size_t p = 0; p < ncell; p++) out[p] = b[p] - a[p];
}

/* SOR relaxation factor, src/main.c:134 (truncated PI). */
/* Pressure-Poisson right-hand side of the commented recipe, src/main.c:421-427:
 *     dudx = DX u, dudy = DY u, dvdx = DX v, dvdy = DY v;   f = dudx**2 + dvdy**2 + 2*dudy*dvdx
 * (DX acts along j with dx, DY along i with dy, src/main.c:146-152).  Left-to-right evaluation, every operation
 * rounded separately; x**2 == x*x.  The pressure itself is p = poisson(-f) with the ordinary solver above. */
int orc_pressure_rhs(const double *u, const double *v, int nx, int ny, int order, double dx, double dy, double *f)
{
    size_t n = (size_t)nx * ny;
    double *d = (double *)malloc(sizeof(double) * 4 * n);
    if (!d) return -2;
    double *dudx = d, *dudy = d + n, *dvdx = d + 2 * n, *dvdy = d + 3 * n;
    int rc = orc_apply(u, nx, ny, 1, 1, order, dx, dudx) | orc_apply(u, nx, ny, 0, 1, order, dy, dudy) |
             orc_apply(v, nx, ny, 1, 1, order, dx, dvdx) | orc_apply(v, nx, ny, 0, 1, order, dy, dvdy);
    if (!rc)
        for (size_t p = 0; p < n; p++) f[p] = dudx[p] * dudx[p] + dvdy[p] * dvdy[p] + 2 * dudy[p] * dvdx[p];
    free(d);
    return rc ? -1 : 0;
}

double orc_beta(int nx, int ny)
{
    return 0.5 * (2 / (1 + sin(ORC_PI / (nx + 1))) + 2 / (1 + sin(ORC_PI / (ny + 1))));
}
/* Number of time steps the reference executes: t = 0 .. it_max, it_max = (int)(tf/dt - 1)
 * (src/main.c:162, :276). */
int orc_num_steps(double tf, double dt) { return (int)((tf / dt) - 1) + 1; }

/* ------------------------------------------------------------------------------------
 * One time step of the reference main loop, src/main.c:283-395, matrix-free.
 * State: u, v, w, psi (nx*ny each).  Parameters mirror the Config fields used there.
 * Returns the Poisson status (0 ok, 1 = itmax hit -> reference would exit(1)).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    double Re, dt, dx, dy, beta, poisson_tol;
    int nx, ny, order, poisson_max_it, poisson_type, redblack;
    double u1, u2, u3, u4, v1, v2, v3, v4;
} orc_params;

int orc_step(const orc_params *P, double *u, double *v, double *w, double *psi,
             int *k_out, double *e_out, double *cont_max, double *cont_min)
{
    const int nx = P->nx, ny = P->ny;
    const size_t nc = (size_t)nx * ny;
    double *t0 = (double *)malloc(sizeof(double) * nc * 4);
    if (!t0) return -2;
    double *t1 = t0 + nc, *t2 = t1 + nc, *t3 = t2 + nc;
    /* Dirichlet BCs: rows first, then columns, so corners take the column values (:283-296) */
    for (int j = 0; j < ny; j++) {
        v[j] = P->v3; v[(size_t)(nx - 1) * ny + j] = P->v4;
        u[j] = P->u3; u[(size_t)(nx - 1) * ny + j] = P->u4;
    }
    for (int i = 0; i < nx; i++) {
        v[(size_t)i * ny] = P->v1; v[(size_t)i * ny + ny - 1] = P->v2;
        u[(size_t)i * ny] = P->u1; u[(size_t)i * ny + ny - 1] = P->u2;
    }
    /* wall vorticity on the ring: w = dvdx - dudy with dudy = DY u, dvdx = DX v (:298-320) */
    orc_apply(u, nx, ny, 0, 1, P->order, P->dy, t0); /* dudy */
    orc_apply(v, nx, ny, 1, 1, P->order, P->dx, t1); /* dvdx */
    for (int j = 0; j < ny; j++) {
        w[j] = t1[j] - t0[j];
        size_t p = (size_t)(nx - 1) * ny + j; w[p] = t1[p] - t0[p];
    }
    for (int i = 0; i < nx; i++) {
        size_t p = (size_t)i * ny; w[p] = t1[p] - t0[p];
        p += ny - 1; w[p] = t1[p] - t0[p];
    }
    /* vorticity derivatives (:323-341) and explicit Euler on ALL points (:344) */
    orc_apply(w, nx, ny, 1, 1, P->order, P->dx, t0); /* dwdx   */
    orc_apply(w, nx, ny, 0, 1, P->order, P->dy, t1); /* dwdy   */
    orc_apply(w, nx, ny, 1, 2, P->order, P->dx, t2); /* d2wdx2 */
    orc_apply(w, nx, ny, 0, 2, P->order, P->dy, t3); /* d2wdy2 */
    orc_euler(w, t0, t1, t2, t3, u, v, P->Re, P->dt, nc);
    /* Poisson for psi with f = -w, zero initial guess (:347-363) */
    for (size_t p = 0; p < nc; p++) t0[p] = -w[p];
    int st = orc_poisson(t0, nx, ny, P->dx, P->dy, P->poisson_max_it, P->poisson_tol,
                         P->poisson_type == 2 ? P->beta : 1.0, P->redblack, P->poisson_type == 2,
                         psi, k_out, e_out, NULL);
    /* velocities on ALL points: u = DY psi, v = -(DX psi) (:366-383) */
    orc_apply(psi, nx, ny, 0, 1, P->order, P->dy, u);
    orc_apply(psi, nx, ny, 1, 1, P->order, P->dx, v);
    for (size_t p = 0; p < nc; p++) v[p] = -v[p];
    /* continuity diagnostic (:387-395, :407-408) */
    orc_apply(u, nx, ny, 1, 1, P->order, P->dx, t0);
    orc_apply(v, nx, ny, 0, 1, P->order, P->dy, t1);
    double mx = -1.7976931348623157e308, mn = 1.7976931348623157e308;
    for (size_t p = 0; p < nc; p++) {
        double c = t0[p] + t1[p];
        if (c > mx) mx = c;
        if (c < mn) mn = c;
    }
    if (cont_max) *cont_max = mx;
    if (cont_min) *cont_min = mn;
    free(t0);
    return st;
}

/* Run nsteps steps from the reference initial condition (interior u=ui, v=vi; w=psi=0;
 * src/main.c:178-181, :214-221).  ks/es (length nsteps) receive the per-step Poisson log
 * values.  Returns the index of the first failing step + 1, or 0. */
int orc_run(const orc_params *P, double ui, double vi, int nsteps,
            double *u, double *v, double *w, double *psi, int *ks, double *es)
{
    const int nx = P->nx, ny = P->ny;
    size_t nc = (size_t)nx * ny;
    memset(u, 0, sizeof(double) * nc); memset(v, 0, sizeof(double) * nc);
    memset(w, 0, sizeof(double) * nc); memset(psi, 0, sizeof(double) * nc);
    for (int i = 1; i < nx - 1; i++)
        for (int j = 1; j < ny - 1; j++) { u[(size_t)i * ny + j] = ui; v[(size_t)i * ny + j] = vi; }
    for (int t = 0; t < nsteps; t++) {
        int k = 0; double e = 0;
        int st = orc_step(P, u, v, w, psi, &k, &e, NULL, NULL);
        if (ks) ks[t] = k;
        if (es) es[t] = e;
        if (st) return t + 1;
    }
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
