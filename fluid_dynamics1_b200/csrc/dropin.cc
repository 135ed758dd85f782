// dropin.cc -- libcnavier_dropin.so: the reference's C signatures (include/cnavier_dropin.h) on top of
// the cnv_* C ABI.  Host mtrx arguments are passed to the GPU entry points as they lie when they are contiguous (every mtrx
// this library's allocm / initm creates is: one page-locked, recycled block), gathered into a staging block otherwise;
// results are written straight into freshly allocated caller-owned mtrx storage.
// Error behaviour mirrors the reference: message, then exit(1).
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/cnavier_b200.h"

namespace {

// ---- field storage ------------------------------------------------------------------------------------------
// allocm() below backs every mtrx with ONE contiguous block (rows[i] = block + i*n) from cnv_host_alloc: page-locked and
// recycled when large.  A field that main.c builds with initm / receives from this library can therefore be handed to the
// GPU as it lies (one DMA, no row-by-row staging); a mtrx assembled by other means (rows malloc'ed one by one, as the
// reference's own allocm does, src/linearalg.c:66-79) is still accepted and gathered into a staging block.
struct Block {
    double *data;
    size_t bytes;
};
std::mutex g_mu;
std::unordered_map<double **, Block> &blocks()
{
    static auto *b = new std::unordered_map<double **, Block>;
    return *b;
}

bool contiguous(mtrx A)
{
    for (int i = 1; i < A.m; i++)
        if (A.M[i] != A.M[0] + (size_t)i * A.n) return false;
    return true;
}

// dense row-major view of a mtrx: the mtrx's own storage when it is contiguous, a (pooled) staging copy otherwise
struct Dense {
    double *p;
    bool staged;
    mtrx src;
    explicit Dense(mtrx A, bool copy_in = true) : p(nullptr), staged(false), src(A)
    {
        if (contiguous(A)) { p = A.M[0]; return; }
        staged = true;
        p = static_cast<double *>(cnv_host_alloc(sizeof(double) * (size_t)A.m * A.n));
        if (!p) { std::printf("** Error: insufficient memory **"); std::exit(1); }
        if (copy_in)
            for (int i = 0; i < A.m; i++) std::memcpy(p + (size_t)i * A.n, A.M[i], sizeof(double) * A.n);
    }
    void copy_back() const
    {
        if (staged)
            for (int i = 0; i < src.m; i++) std::memcpy(src.M[i], p + (size_t)i * src.n, sizeof(double) * src.n);
    }
    ~Dense() { if (staged) cnv_host_free(p); }
    Dense(const Dense &) = delete;
};

mtrx new_mtrx(int m, int n)
{
    mtrx A;
    A.M = allocm(m, n);
    A.m = m;
    A.n = n;
    return A;
}
mtrx from_dense(const std::vector<double> &h, int m, int n)
{
    mtrx A = new_mtrx(m, n);
    std::memcpy(A.M[0], h.data(), sizeof(double) * (size_t)m * n);
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
    Dense hf(f);
    mtrx U = new_mtrx(f.m, f.n);  // caller-owned result (released with freem), written by the GPU path directly
    int k = 0;
    double e = 0;
    const int status = cnv_poisson_host(hf.p, f.m, f.n, dx, dy, itmax, tol, beta, 0, U.M[0], &k, &e, nullptr);
    if (status != 0) {
        if (to_stdout) std::printf("%s", kItmax);
        else log_to(log, "%s", kItmax);
        std::exit(1);
    }
    if (to_stdout) std::printf(kSolved, k, e);
    else log_to(log, kSolved, k, e);
    return U;
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
    // one contiguous block behind the row table (see "field storage" above); freem() below is its counterpart.  The
    // reference's pair (src/linearalg.c:53-99) mallocs and frees row by row; callers only ever go through
    // allocm / freem and A.M[i][j], so the pair may choose its own backing store.
    double **rows = static_cast<double **>(std::malloc(sizeof(double *) * m));
    const size_t bytes = sizeof(double) * (size_t)m * n;
    double *data = rows ? static_cast<double *>(cnv_host_alloc(bytes)) : nullptr;
    if (!rows || !data) { std::printf("** Error: insufficient memory **"); std::exit(1); }
    for (int i = 0; i < m; i++) rows[i] = data + (size_t)i * n;
    std::lock_guard<std::mutex> lk(g_mu);
    blocks()[rows] = {data, bytes};
    return rows;
}
double **freem(mtrx A)
{
    if (!A.M) return nullptr;
    if (A.m < 1 || A.n < 1) {
        std::printf("** Error: invalid parameter **\n");
        std::exit(1);
    }
    bool ours = false;
    Block b = {nullptr, 0};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = blocks().find(A.M);
        if (it != blocks().end()) { ours = true; b = it->second; blocks().erase(it); }
    }
    if (ours) cnv_host_free(b.data);
    else for (int i = 0; i < A.m; i++) std::free(A.M[i]);  // a mtrx built the reference's way: one malloc per row
    std::free(A.M);
    return nullptr;
}
void zerosm(mtrx A)
{
    for (int i = 0; i < A.m; i++) std::memset(A.M[i], 0, sizeof(double) * A.n);
}
mtrx initm(int m, int n)
{
    mtrx A = new_mtrx(m, n);
    std::memset(A.M[0], 0, sizeof(double) * (size_t)m * n);
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
    mtrx C = new_mtrx(n, n);
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
    Dense hw(w), a(dwdx), b(dwdy), c(d2wdx2), d(d2wdy2), hu(u), hv(v);
    cnv_euler_host(hw.p, a.p, b.p, c.p, d.p, hu.p, hv.p, w.m, w.n, Re, dt);
    hw.copy_back();
}
mtrx continuity(mtrx dudx, mtrx dvdy)
{
    Dense a(dudx), b(dvdy);
    mtrx o = new_mtrx(dudx.m, dudx.n);
    cnv_continuity_host(a.p, b.p, dudx.m, dudx.n, o.M[0]);
    return o;
}
mtrx vorticity(mtrx first, mtrx second)
{
    Dense a(first), b(second);
    mtrx o = new_mtrx(first.m, first.n);
    cnv_vorticity_host(a.p, b.p, first.m, first.n, o.M[0]);
    return o;
}

// ---- include/poisson.h ----------------------------------------------------------------------------
double error(mtrx u1, mtrx u2)
{
    Dense a(u1), b(u2);
    return cnv_error_host(a.p, b.p, u1.m, u1.n);
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
