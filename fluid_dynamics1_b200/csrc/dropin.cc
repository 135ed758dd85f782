// dropin.cc -- libcnavier_dropin.so: the reference's C signatures (include/cnavier_dropin.h) on top of
// the cnv_* C ABI.  Host mtrx arguments are gathered into dense row-major staging buffers, the GPU
// entry point runs, results are scattered into freshly allocated caller-owned mtrx storage.
// Error behaviour mirrors the reference: message, then exit(1).
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cnavier_b200.h"

namespace {

std::vector<double> gather(mtrx A)
{
    std::vector<double> h((size_t)A.m * A.n);
    for (int i = 0; i < A.m; i++) std::memcpy(&h[(size_t)i * A.n], A.M[i], sizeof(double) * A.n);
    return h;
}
void scatter(const std::vector<double> &h, mtrx A)
{
    for (int i = 0; i < A.m; i++) std::memcpy(A.M[i], &h[(size_t)i * A.n], sizeof(double) * A.n);
}
mtrx from_dense(const std::vector<double> &h, int m, int n)
{
    mtrx A;
    A.M = allocm(m, n);
    A.m = m;
    A.n = n;
    scatter(h, A);
    return A;
}
void log_to(FILE *f, const char *fmt, ...)
{
    if (!f) return;
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(f, fmt, ap);
    va_end(ap);
    std::fflush(f);
}
const char kSolved[] = "Poisson equation solved with %d iterations - root-sum-of-squares error: %E\n";
const char kItmax[] = "Error: maximum number of iterations achieved for Poisson equation.\n";

// to_stdout: poisson / poisson_SOR print with printf (src/poisson.c:101,163); the _log variants write
// to the FILE* only, and stay silent when it is NULL (src/poisson.c:24-32)
mtrx solve(mtrx f, double dx, double dy, int itmax, double tol, double beta, FILE *log, bool to_stdout)
{
    std::vector<double> hf = gather(f), hu(hf.size());
    int k = 0;
    double e = 0;
    const int status = cnv_poisson_host(hf.data(), f.m, f.n, dx, dy, itmax, tol, beta, 0, hu.data(), &k, &e, nullptr);
    if (status != 0) {
        if (to_stdout) std::printf("%s", kItmax);
        else log_to(log, "%s", kItmax);
        std::exit(1);
    }
    if (to_stdout) std::printf(kSolved, k, e);
    else log_to(log, kSolved, k, e);
    return from_dense(hu, f.m, f.n);
}

mtrx diff(int n, int o, double h, int deriv)
{
    std::vector<double> D((size_t)n * n);
    if (cnv_diff_dense(n, o, deriv, h, D.data()) != 0) {
        std::printf("** Error: valid orders are 2, 4 or 6 **\n");
        std::exit(1);
    }
    return from_dense(D, n, n);
}

}  // namespace

extern "C" {

// ---- include/linearalg.h helpers (src/linearalg.c) ---------------------------------------------
double **allocm(int m, int n)
{
    if (m < 1 || n < 1) {
        std::printf("** Error: invalid parameter **\n");
        std::exit(1);
    }
    double **rows = static_cast<double **>(std::malloc(sizeof(double *) * m));
    if (!rows) { std::printf("** Error: insufficient memory **"); std::exit(1); }
    for (int i = 0; i < m; i++) {
        rows[i] = static_cast<double *>(std::malloc(sizeof(double) * n));  // one allocation per row: freem() contract
        if (!rows[i]) { std::printf("** Error: insufficient memory **"); std::exit(1); }
    }
    return rows;
}
double **freem(mtrx A)
{
    if (!A.M) return nullptr;
    if (A.m < 1 || A.n < 1) {
        std::printf("** Error: invalid parameter **\n");
        std::exit(1);
    }
    for (int i = 0; i < A.m; i++) std::free(A.M[i]);
    std::free(A.M);
    return nullptr;
}
void zerosm(mtrx A)
{
    for (int i = 0; i < A.m; i++) std::memset(A.M[i], 0, sizeof(double) * A.n);
}
mtrx initm(int m, int n)
{
    mtrx A;
    A.M = allocm(m, n);
    A.m = m;
    A.n = n;
    zerosm(A);
    return A;
}
mtrx eye(int n)
{
    mtrx A = initm(n, n);
    for (int i = 0; i < n; i++) A.M[i][i] = 1;
    return A;
}
mtrx reshape(mtrx A, int m, int n)
{
    if (A.m * A.n != m * n) {
        std::printf("** Error: the reshaped matrix must have the same number of elements **\n");
        std::printf("Number of elements of input matrix: %d\n", A.m * A.n);
        std::printf("Number of elements of output matrix: %d\n", m * n);
        std::exit(1);
    }
    mtrx B = initm(m, n);
    for (long p = 0; p < (long)m * n; p++) B.M[p / n][p % n] = A.M[p / A.n][p % A.n];  // row-major order kept
    return B;
}
mtrx kronecker(mtrx A, mtrx B)
{
    // the reference sizes the result A.n*B.n square and indexes blocks with A.n (src/linearalg.c:356,367)
    const int n = A.n * B.n;
    mtrx C;
    C.M = allocm(n, n);
    C.m = C.n = n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) C.M[i][j] = A.M[i / A.n][j / A.n] * B.M[i % B.n][j % B.n];
    return C;
}
mtrx mtrxmul(mtrx A, mtrx B)
{
    if (A.n != B.m) {
        std::printf("** Error: the first matrix number of columns must be equal to the second matrix number of rows **\n");
        std::printf("Columns of first matrix: %d\n", A.n);
        std::printf("Rows of second matrix: %d\n", B.m);
        std::exit(1);
    }
    mtrx C = initm(A.m, B.n);
    for (int i = 0; i < C.m; i++)
        for (int j = 0; j < C.n; j++) {
            double sum = 0.0;
            for (int k = 0; k < A.n; k++) sum += A.M[i][k] * B.M[k][j];
            C.M[i][j] = sum;
        }
    return C;
}
void invsig(mtrx A)
{
    for (int i = 0; i < A.m; i++)
        for (int j = 0; j < A.n; j++) A.M[i][j] = -A.M[i][j];
}
double maxel(mtrx A)
{
    double best = -DBL_MAX;
    for (int i = 0; i < A.m; i++)
        for (int j = 0; j < A.n; j++)
            if (A.M[i][j] > best) best = A.M[i][j];
    return best;
}
double minel(mtrx A)
{
    double best = DBL_MAX;
    for (int i = 0; i < A.m; i++)
        for (int j = 0; j < A.n; j++)
            if (A.M[i][j] < best) best = A.M[i][j];
    return best;
}
void mtrxcpy(mtrx A, mtrx B)
{
    for (int i = 0; i < A.m; i++) std::memcpy(A.M[i], B.M[i], sizeof(double) * A.n);
}
void set_openmp_config(int) {}
void set_poisson_openmp_config(int) {}
void set_fluiddyn_openmp_config(int) {}

// ---- include/finitediff.h -----------------------------------------------------------------------
mtrx Diff1(int n, int o, double dx) { return diff(n, o, dx, 1); }
mtrx Diff2(int n, int o, double dx) { return diff(n, o, dx, 2); }

// ---- include/fluiddyn.h ---------------------------------------------------------------------------
void euler(mtrx w, mtrx dwdx, mtrx dwdy, mtrx d2wdx2, mtrx d2wdy2, mtrx u, mtrx v, double Re, double dt)
{
    std::vector<double> hw = gather(w), a = gather(dwdx), b = gather(dwdy), c = gather(d2wdx2), d = gather(d2wdy2),
                        hu = gather(u), hv = gather(v);
    cnv_euler_host(hw.data(), a.data(), b.data(), c.data(), d.data(), hu.data(), hv.data(), w.m, w.n, Re, dt);
    scatter(hw, w);
}
mtrx continuity(mtrx dudx, mtrx dvdy)
{
    std::vector<double> a = gather(dudx), b = gather(dvdy), o(a.size());
    cnv_continuity_host(a.data(), b.data(), dudx.m, dudx.n, o.data());
    return from_dense(o, dudx.m, dudx.n);
}
mtrx vorticity(mtrx first, mtrx second)
{
    std::vector<double> a = gather(first), b = gather(second), o(a.size());
    cnv_vorticity_host(a.data(), b.data(), first.m, first.n, o.data());
    return from_dense(o, first.m, first.n);
}

// ---- include/poisson.h ----------------------------------------------------------------------------
double error(mtrx u1, mtrx u2)
{
    std::vector<double> a = gather(u1), b = gather(u2);
    return cnv_error_host(a.data(), b.data(), u1.m, u1.n);
}
mtrx poisson(mtrx f, double dx, double dy, int itmax, double tol) { return solve(f, dx, dy, itmax, tol, 1.0, nullptr, true); }
mtrx poisson_SOR(mtrx f, double dx, double dy, int itmax, double tol, double beta)
{
    return solve(f, dx, dy, itmax, tol, beta, nullptr, true);
}
mtrx poisson_log(mtrx f, double dx, double dy, int itmax, double tol, FILE *log_file)
{
    return solve(f, dx, dy, itmax, tol, 1.0, log_file, false);
}
mtrx poisson_SOR_log(mtrx f, double dx, double dy, int itmax, double tol, double beta, FILE *log_file)
{
    return solve(f, dx, dy, itmax, tol, beta, log_file, false);
}

// ---- include/config.h -----------------------------------------------------------------------------
Config load_default_config(void)
{
    Config c;
    cnv_config_default(&c);
    return c;
}
Config load_config_from_file(const char *filename)
{
    Config c;
    cnv_config_from_file(filename, &c);
    return c;
}
void print_config(const Config *config) { cnv_config_print(config); }
void print_usage(const char *program_name)
{
    std::printf("Usage: %s [config_file] [output_folder]\n", program_name);
    std::printf("  config_file    key = value text file (see --help-config); defaults when omitted\n");
    std::printf("  output_folder  results go to ./output/[output_folder]/\n");
}
void print_openmp_status(const Config *config)
{
    std::printf("\n=== OpenMP Status ===\n");
    std::printf("OpenMP Support: NOT USED (CUDA sm_100a path)\n");
    std::printf("Configuration Setting: %s (ignored)\n", config->openmp_enabled ? "ENABLED" : "DISABLED");
    std::printf("Status: %d CUDA device(s) visible; red-black ordering as in the OpenMP build\n", cnv_device_count());
    std::printf("====================\n\n");
}

}  // extern "C"
