/*
 * ref_shim.c -- flat-array entry points around the UNMODIFIED reference functions.
 * TEST INFRASTRUCTURE ONLY; compiled (together with the reference's own .c files, read
 * in place from /root/reference/src) into oracle/_ref/libcnavier_ref.so by oracle/Makefile.
 * Nothing here restates an algorithm: every wrapper marshals a row-major double array
 * into the reference's `mtrx` (include/linearalg.h:6-11), calls the real function and
 * copies the answer back, so that Python (ctypes) can use the reference as ground truth.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "linearalg.h"
#include "finitediff.h"
#include "fluiddyn.h"
#include "poisson.h"
#include "config.h"

static mtrx from_flat(const double *a, int m, int n)
{
    mtrx A = initm(m, n);
    for (int i = 0; i < m; i++) memcpy(A.M[i], a + (size_t)i * n, sizeof(double) * n);
    return A;
}
static void to_flat(mtrx A, double *a)
{
    for (int i = 0; i < A.m; i++) memcpy(a + (size_t)i * A.n, A.M[i], sizeof(double) * A.n);
}

void ref_set_openmp(int enabled)
{
    set_openmp_config(enabled);
    set_poisson_openmp_config(enabled);
    set_fluiddyn_openmp_config(enabled);
}

/* Diff1 / Diff2 dense matrices (src/finitediff.c:51, :178) */
void ref_diff(int n, int o, int deriv, double h, double *D)
{
    mtrx M = deriv == 1 ? Diff1(n, o, h) : Diff2(n, o, h);
    to_flat(M, D);
    M.M = freem(M);
}

/* The reference's dense operator route for one field: eye + kronecker + reshape + mtrxmul
 * (src/main.c:142-152, :298-304).  axis 1 = DX = kron(I, d), axis 0 = DY = kron(d, I). */
void ref_apply_dense(const double *a, int n, int axis, int deriv, int order, double h, double *out)
{
    mtrx d = deriv == 1 ? Diff1(n, order, h) : Diff2(n, order, h);
    mtrx I = eye(n);
    mtrx K = axis == 1 ? kronecker(I, d) : kronecker(d, I);
    mtrx A = from_flat(a, n, n);
    mtrx a0 = reshape(A, n * n, 1);
    mtrx r0 = mtrxmul(K, a0);
    mtrx R = reshape(r0, n, n);
    to_flat(R, out);
    d.M = freem(d); I.M = freem(I); K.M = freem(K); A.M = freem(A);
    a0.M = freem(a0); r0.M = freem(r0); R.M = freem(R);
}

/* poisson_SOR_log / poisson_log (src/poisson.c:224-285 / :176-222); the iteration count and
 * residual are parsed back from the reference's own log line (:275). Converging inputs only:
 * the reference calls exit(1) at itmax. */
int ref_poisson(const double *f, int m, int n, double dx, double dy, int itmax, double tol, double beta,
                int type, double *u, int *k_out, double *e_out)
{
    char *buf = NULL; size_t len = 0;
    FILE *log = open_memstream(&buf, &len);
    mtrx F = from_flat(f, m, n);
    mtrx U = type == 2 ? poisson_SOR_log(F, dx, dy, itmax, tol, beta, log) : poisson_log(F, dx, dy, itmax, tol, log);
    fclose(log);
    int k = -1; double e = -1;
    int ok = buf && sscanf(buf, "Poisson equation solved with %d iterations - root-sum-of-squares error: %lE", &k, &e) == 2;
    free(buf);
    to_flat(U, u);
    *k_out = k; *e_out = e;
    F.M = freem(F); U.M = freem(U);
    return ok ? 0 : 1;
}

double ref_error(const double *a, const double *b, int m, int n)
{
    mtrx A = from_flat(a, m, n), B = from_flat(b, m, n);
    double e = error(A, B);
    A.M = freem(A); B.M = freem(B);
    return e;
}

void ref_euler(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
               const double *u, const double *v, int m, int n, double Re, double dt)
{
    mtrx W = from_flat(w, m, n), A = from_flat(dwdx, m, n), B = from_flat(dwdy, m, n), C = from_flat(d2wdx2, m, n),
         D = from_flat(d2wdy2, m, n), U = from_flat(u, m, n), V = from_flat(v, m, n);
    euler(W, A, B, C, D, U, V, Re, dt);
    to_flat(W, w);
    W.M = freem(W); A.M = freem(A); B.M = freem(B); C.M = freem(C); D.M = freem(D); U.M = freem(U); V.M = freem(V);
}

void ref_continuity(const double *a, const double *b, int m, int n, double *out)
{
    mtrx A = from_flat(a, m, n), B = from_flat(b, m, n);
    mtrx C = continuity(A, B);
    to_flat(C, out);
    A.M = freem(A); B.M = freem(B); C.M = freem(C);
}
void ref_vorticity(const double *a, const double *b, int m, int n, double *out)
{
    mtrx A = from_flat(a, m, n), B = from_flat(b, m, n);
    mtrx C = vorticity(A, B);
    to_flat(C, out);
    A.M = freem(A); B.M = freem(B); C.M = freem(C);
}

/* config system (src/config.c:47, :106) */
int ref_config_size(void) { return (int)sizeof(Config); }
void ref_load_default_config(Config *c) { *c = load_default_config(); }
void ref_load_config_from_file(const char *fn, Config *c) { *c = load_config_from_file(fn); }
void ref_print_config(const Config *c) { print_config(c); fflush(stdout); }
