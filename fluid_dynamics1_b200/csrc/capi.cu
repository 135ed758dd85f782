// capi.cu -- extern "C" surface of libcnavier_b200.so (declared in include/cnavier_b200.h) and the
// device-resident time stepper that restates the loop body of the reference driver
// (src/main.c:283-395) on top of the CUDA kernels.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#include "../../include/cnavier_b200.h"
#include "kernels.h"
#include "nccl_dl.h"

namespace cnv {
size_t total_launches();
void count_launch(size_t n);

static void require_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        std::printf("** Error: no CUDA device visible: the B200 path has no CPU fallback **\n");
        std::fflush(stdout);
        std::exit(1);
    }
}

// RAII device scratch for the host-staging entry points
struct DevArray {
    double *p = nullptr;
    int nrows, ncols, ld;
    DevArray(int nr, int nc) : nrows(nr), ncols(nc), ld(round_up(nc, 16))
    {
        CNV_CUDA_CHECK(cudaMalloc(&p, sizeof(double) * (size_t)nr * ld));
        CNV_CUDA_CHECK(cudaMemset(p, 0, sizeof(double) * (size_t)nr * ld));
    }
    ~DevArray() { cudaFree(p); }
    void upload(const double *h, cudaStream_t s = 0)
    {
        CNV_CUDA_CHECK(cudaMemcpy2DAsync(p, sizeof(double) * ld, h, sizeof(double) * ncols, sizeof(double) * ncols, nrows,
                                         cudaMemcpyHostToDevice, s));
    }
    void download(double *h, cudaStream_t s = 0) const
    {
        CNV_CUDA_CHECK(cudaMemcpy2DAsync(h, sizeof(double) * ncols, p, sizeof(double) * ld, sizeof(double) * ncols, nrows,
                                         cudaMemcpyDeviceToHost, s));
        CNV_CUDA_CHECK(cudaStreamSynchronize(s));
    }
};

// L1 distance kernel for error() (src/poisson.c:34-60): block partials, fixed-order final sum on host
__global__ void k_l1_distance(const double *__restrict__ a, const double *__restrict__ b, int nrows, int ncols, int ld,
                              double *__restrict__ partial)
{
    double s = 0.0;
    const size_t n = (size_t)nrows * ld;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        if ((int)(p % ld) < ncols) s = xadd(s, fabs(xsub(b[p], a[p])));
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = xadd(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---- pinned host memory pool (cnv_host_alloc / cnv_host_free) ------------------------------------------------
// The reference hands fields around as caller-owned mtrx storage that is allocated and freed once per call and per time
// step (allocm / freem, src/linearalg.c:53-99).  Page-locking 134 MB takes longer than solving on it, so large blocks
// are page-locked once and recycled through size-keyed free lists; small ones are plain heap memory.
struct HostPool {
    std::mutex mu;
    std::unordered_map<void *, std::pair<size_t, bool>> live;   // ptr -> (bytes, pinned)
    std::unordered_map<size_t, std::vector<void *>> free_pinned; // bytes -> recycled pinned blocks
    std::unordered_map<void *, size_t> registered;              // pinned blocks that are mmap + cudaHostRegister (ptr -> mapped bytes)
    size_t pooled_bytes = 0;
};
static HostPool &host_pool()
{
    static HostPool *p = new HostPool;  // never destroyed: blocks may outlive static destruction order
    return *p;
}
constexpr size_t kPinThreshold = 1u << 20;        // blocks of >= 1 MiB are page-locked
constexpr size_t kPoolLimit = (size_t)8 << 30;    // recycled (idle) pinned memory kept at most

static bool device_present()
{
    static int n = -1;
    if (n < 0) {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
        n = c;
    }
    return n > 0;
}

// NUMA node of the current device (sysfs entry of its PCI function), -1 if unknown.  On a two-socket box every GPU hangs off
// one socket; page-locked memory on the other socket makes every H2D / D2H cross the socket interconnect, which caps the
// copies of the far GPUs when all GPUs of the box copy at once (8 GPUs: 14 GB/s per GPU against 45 GB/s at 2 GPUs,
// profiles/scale_r2.md).
static int device_numa_node()
{
    static int node[kMaxDevices];
    static bool known[kMaxDevices] = {};
    const int slot = current_device_slot();
    if (known[slot]) return node[slot];
    int dev = 0, n = -1;
    char bus[64] = {0};
    if (const char *e = std::getenv("CNV_HOST_NUMA_NODE")) {  // platforms whose sysfs does not say (-1): name the node by hand
        n = std::atoi(e);
    } else if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetPCIBusId(bus, sizeof bus, dev) == cudaSuccess) {
        for (char *c = bus; *c; c++) *c = (char)std::tolower((unsigned char)*c);
        char path[160];
        std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
        if (FILE *f = std::fopen(path, "r")) {
            if (std::fscanf(f, "%d", &n) != 1) n = -1;
            std::fclose(f);
        }
    } else {
        cudaGetLastError();
    }
    node[slot] = n; known[slot] = true;
    return n;
}
// `bytes` of page-locked memory whose pages live on the current device's NUMA node: anonymous mapping, preferred-node policy
// on the range (mbind; a failure -- no permission, one node only -- just leaves the default placement), first touch, then
// cudaHostRegister.  nullptr if any step fails; the caller then takes cudaHostAlloc.
static void *numa_local_pinned(size_t bytes)
{
    const int node = device_numa_node();
    if (node < 0 || node >= 1024) return nullptr;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    unsigned long mask[16] = {0};
    mask[node / 64] |= 1ul << (node % 64);
    (void)syscall(SYS_mbind, p, bytes, 1 /* MPOL_PREFERRED */, mask, 1024ul, 0u);
    (void)madvise(p, bytes, MADV_HUGEPAGE);
    std::memset(p, 0, bytes);
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, bytes);
        return nullptr;
    }
    return p;
}

// ---- solver objects of the host-buffer entry point, kept between calls ---------------------------------------
// cnv_poisson_host() is what poisson_SOR_log(mtrx ...) of the drop-in library lands on, once per time step with the same
// grid: creating a PoissonSolver per call costs three cudaMalloc + memset of the field size.  The last two shapes per
// device stay alive (the pressure recipe alternates two solves); the pass-count predictor carries over between calls.
struct CachedSolver {
    int dev, nrows, ncols, T;
    size_t env;  // fingerprint of the CNV_* environment the solver was planned under (tuning / A-B switches are read at construction)
    PoissonSolver *s;
};
extern "C" char **environ;
static size_t cnv_env_fingerprint()
{
    size_t h = 1469598103934665603ull;  // FNV-1a over every CNV_* variable
    for (char **e = environ; e && *e; e++) {
        if (std::strncmp(*e, "CNV_", 4) != 0) continue;
        for (const char *c = *e; *c; c++) h = (h ^ (unsigned char)*c) * 1099511628211ull;
        h = (h ^ 0xffu) * 1099511628211ull;
    }
    return h;
}
static std::vector<CachedSolver> &solver_cache()
{
    static std::vector<CachedSolver> *c = new std::vector<CachedSolver>;
    return *c;
}
static PoissonSolver *cached_solver(int nrows, int ncols, int T)
{
    int dev = 0;
    CNV_CUDA_CHECK(cudaGetDevice(&dev));
    auto &c = solver_cache();
    const size_t env = cnv_env_fingerprint();
    for (size_t i = 0; i < c.size(); i++)
        if (c[i].dev == dev && c[i].nrows == nrows && c[i].ncols == ncols && c[i].T == T && c[i].env == env) {
            CachedSolver hit = c[i];
            c.erase(c.begin() + i);
            c.push_back(hit);  // most recently used last
            return hit.s;
        }
    size_t same_dev = 0;
    for (const CachedSolver &e : c) same_dev += e.dev == dev;
    if (same_dev >= 2)
        for (size_t i = 0; i < c.size(); i++)
            if (c[i].dev == dev) { delete c[i].s; c.erase(c.begin() + i); break; }
    c.push_back({dev, nrows, ncols, T, env, new PoissonSolver(nrows, ncols, T)});
    return c.back().s;
}
}  // namespace cnv

using namespace cnv;

struct cnv_poisson {
    PoissonSolver *s;
};

// ---------------------------------------------------------------------------------------------
struct cnv_sim {
    Config cfg;
    int nrows, ncols, ld;  // nrows = rows of the LOCAL arrays (== cfg.nx on one GPU)
    RowMap map;            // slab geometry (single GPU: the whole grid is owned)
    cnv_poisson ph{nullptr};  // non-owning view of `ps` for the C ABI
    double dx, dy, beta, inv_re;
    FdTable d1x, d1y, d2x, d2y;
    double *u = nullptr, *v = nullptr, *w = nullptr, *w2 = nullptr;
    PoissonSolver *ps = nullptr;
    PoissonSolver *pp = nullptr;  // pressure solves (created on first use; psi stays intact in `ps`)
    int psi_buf = 0;
    double *cont_partial = nullptr, *cont_result = nullptr, *h_cont = nullptr;
    unsigned *cont_ticket = nullptr;
    bool diag = true;
    long long sweeps = 0, passes = 0, steps = 0;
    cudaStream_t stream = 0;
};

extern "C" {

const char *cnv_version(void) { return "cnavier-b200 0.1 (sm_100a)"; }

int cnv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// one process per GPU: select the device for this library's (statically linked) CUDA runtime
int cnv_set_device(int device)
{
    require_device();
    CNV_CUDA_CHECK(cudaSetDevice(device));
    return 0;
}
int cnv_get_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}
void cnv_device_synchronize(void) { CNV_CUDA_CHECK(cudaDeviceSynchronize()); }

unsigned long long cnv_launch_count(void) { return (unsigned long long)total_launches(); }

// Host memory for fields: page-locked and recycled when large (see HostPool).  Without a CUDA device it is plain heap
// memory (this is storage, not compute: nothing here computes on the CPU).  Contents are NOT zeroed.
void *cnv_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 8;
    HostPool &hp = host_pool();
    const bool pin = bytes >= kPinThreshold && device_present();
    void *p = nullptr;
    if (pin) {
        {
            std::lock_guard<std::mutex> lk(hp.mu);
            auto it = hp.free_pinned.find(bytes);
            if (it != hp.free_pinned.end() && !it->second.empty()) {
                p = it->second.back();
                it->second.pop_back();
                hp.pooled_bytes -= bytes;
            }
        }
        if (!p) {
            const size_t mapped = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);  // whole 2 MiB pages
            p = numa_local_pinned(mapped);
            if (p) {
                std::lock_guard<std::mutex> lk(hp.mu);
                hp.registered[p] = mapped;
            }
        }
        if (!p && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
    }
    bool pinned = p != nullptr;
    if (!p) p = std::malloc(bytes);
    if (!p) return nullptr;
    std::lock_guard<std::mutex> lk(hp.mu);
    hp.live[p] = {bytes, pinned};
    return p;
}
void cnv_host_free(void *p)
{
    if (!p) return;
    HostPool &hp = host_pool();
    size_t bytes = 0;
    bool pinned = false, known = false;
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        auto it = hp.live.find(p);
        if (it != hp.live.end()) {
            known = true; bytes = it->second.first; pinned = it->second.second;
            hp.live.erase(it);
            if (pinned && hp.pooled_bytes + bytes <= kPoolLimit) {
                hp.free_pinned[bytes].push_back(p);
                hp.pooled_bytes += bytes;
                return;
            }
        }
    }
    if (known && pinned) {
        size_t mapped = 0;
        {
            std::lock_guard<std::mutex> lk(hp.mu);
            auto it = hp.registered.find(p);
            if (it != hp.registered.end()) { mapped = it->second; hp.registered.erase(it); }
        }
        if (mapped) { cudaHostUnregister(p); munmap(p, mapped); }
        else cudaFreeHost(p);
    } else {
        std::free(p);
    }
}
int cnv_host_numa_node(void) { return device_present() ? device_numa_node() : -1; }
/* 1 if p points into a live page-locked block of cnv_host_alloc */
int cnv_host_is_pinned(const void *p)
{
    HostPool &hp = host_pool();
    std::lock_guard<std::mutex> lk(hp.mu);
    for (const auto &kv : hp.live)
        if (kv.second.second && (const char *)p >= (const char *)kv.first && (const char *)p < (const char *)kv.first + kv.second.first)
            return 1;
    return 0;
}

// src/main.c:134 with the truncated PI of include/poisson.h:9
double cnv_sor_beta(int nx, int ny) { return 0.5 * (2 / (1 + sin(PI / (nx + 1))) + 2 / (1 + sin(PI / (ny + 1)))); }
// src/main.c:162, :276: t = 0 .. (int)(tf/dt - 1)
int cnv_num_steps(double tf, double dt) { return (int)((tf / dt) - 1) + 1; }

int cnv_diff_dense(int n, int order, int deriv, double h, double *D)
{
    FdTable t;
    if (n < order || !fd_make_table(n, order, deriv, h, &t)) return 1;
    std::memset(D, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; i++) {
        int start; const FdRow *r;
        fd_row(t, i, &start, &r);
        for (int k = 0; k < r->cnt; k++) D[(size_t)i * n + start + k] = r->c[k];
    }
    return 0;
}

int cnv_apply_host(const double *A, int nrows, int ncols, int axis, int deriv, int order, double h, double *out)
{
    FdTable t;
    const int n = axis == 1 ? ncols : nrows;
    if (n < order || !fd_make_table(n, order, deriv, h, &t)) return 1;
    require_device();
    DevArray a(nrows, ncols), o(nrows, ncols);
    a.upload(A);
    launch_apply(a.p, nrows, ncols, a.ld, axis, t, o.p, o.ld, 1.0, 0);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    o.download(out);
    return 0;
}

int cnv_euler_host(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
                   const double *u, const double *v, int nrows, int ncols, double Re, double dt)
{
    require_device();
    const double *src[7] = {w, dwdx, dwdy, d2wdx2, d2wdy2, u, v};
    std::vector<DevArray *> d;
    for (int i = 0; i < 7; i++) { d.push_back(new DevArray(nrows, ncols)); d[i]->upload(src[i]); }
    launch_euler_pointwise(d[0]->p, d[1]->p, d[2]->p, d[3]->p, d[4]->p, d[5]->p, d[6]->p, (size_t)nrows * d[0]->ld,
                           1. / Re, dt, 0);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    d[0]->download(w);
    for (auto *x : d) delete x;
    return 0;
}

static int addsub_host(const double *a, const double *b, int nrows, int ncols, double *out, int sub)
{
    require_device();
    DevArray da(nrows, ncols), db(nrows, ncols), dc(nrows, ncols);
    da.upload(a); db.upload(b);
    launch_pointwise_addsub(da.p, db.p, dc.p, (size_t)nrows * da.ld, sub, 0);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    dc.download(out);
    return 0;
}
int cnv_continuity_host(const double *dudx, const double *dvdy, int nrows, int ncols, double *out)
{
    return addsub_host(dudx, dvdy, nrows, ncols, out, 0);
}
int cnv_vorticity_host(const double *a, const double *b, int nrows, int ncols, double *out)
{
    return addsub_host(a, b, nrows, ncols, out, 1);
}

// f = dudx^2 + dvdy^2 + 2*dudy*dvdx from host u, v (the commented pressure recipe, src/main.c:421-427)
int cnv_pressure_rhs_host(const double *u, const double *v, int nrows, int ncols, int order, double dx, double dy, double *f_out)
{
    FdTable tx, ty;
    if (ncols < order || nrows < order || !fd_make_table(ncols, order, 1, dx, &tx) || !fd_make_table(nrows, order, 1, dy, &ty)) return 1;
    require_device();
    DevArray du(nrows, ncols), dv(nrows, ncols), df(nrows, ncols);
    du.upload(u); dv.upload(v);
    const RowMap m = {nrows, 0, nrows, 0, nrows};
    launch_pressure_rhs(du.p, dv.p, m, ncols, du.ld, tx, ty, 1.0, df.p, nullptr, df.ld, 0);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    df.download(f_out);
    return 0;
}

double cnv_error_host(const double *a, const double *b, int nrows, int ncols)
{
    require_device();
    DevArray da(nrows, ncols), db(nrows, ncols);
    da.upload(a); db.upload(b);
    const int nb = 296;
    double *partial;
    CNV_CUDA_CHECK(cudaMalloc(&partial, sizeof(double) * nb));
    k_l1_distance<<<nb, 256>>>(da.p, db.p, nrows, ncols, da.ld, partial);
    count_launch(1);
    std::vector<double> h(nb);
    CNV_CUDA_CHECK(cudaMemcpy(h.data(), partial, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    cudaFree(partial);
    double e = 0;
    for (double x : h) e += x;
    return e;
}

// ---- Poisson -----------------------------------------------------------------------------------
cnv_poisson *cnv_poisson_create(int nrows, int ncols, int T)
{
    require_device();
    return new cnv_poisson{new PoissonSolver(nrows, ncols, T)};
}
cnv_poisson *cnv_poisson_create_slab(int nrows, int ncols, int T, int grow0, int gnrows, int own_lo, int own_hi)
{
    require_device();
    return new cnv_poisson{new PoissonSolver(nrows, ncols, T, grow0, gnrows, own_lo, own_hi)};
}
void cnv_poisson_destroy(cnv_poisson *p)
{
    if (!p) return;
    delete p->s;
    delete p;
}
void cnv_poisson_set_consts(cnv_poisson *p, double dx, double dy, double beta) { p->s->set_consts(dx, dy, beta); }
int cnv_poisson_ld(const cnv_poisson *p) { return p->s->ld(); }
double *cnv_poisson_rhs_ptr(cnv_poisson *p) { return p->s->rhs(); }
static int buf_index(cnv_poisson *p, int which) { return which >= 0 && which < p->s->num_buffers() ? which : which & 1; }
double *cnv_poisson_buf_ptr(cnv_poisson *p, int which) { return p->s->buffer(buf_index(p, which)); }
int cnv_poisson_num_buffers(cnv_poisson *p) { return p->s->num_buffers(); }
double *cnv_poisson_norms_ptr(cnv_poisson *p) { return p->s->local_norms(); }
void cnv_poisson_plan_info(const cnv_poisson *p, long long *out)
{
    const PassGeom &g = p->s->geom();
    out[0] = g.WS; out[1] = g.HX; out[2] = g.Wout; out[3] = g.Hout; out[4] = g.nstrips; out[5] = g.nchunks;
    out[6] = pass_threads(p->s->T(), g.WS); out[7] = (long long)pass_smem_bytes(p->s->T(), g.WS);
    out[8] = p->s->T(); out[9] = p->s->consts().pow2;
    const OnchipGeom &o = p->s->onchip_geom();
    out[10] = p->s->onchip() ? 1 : 0;
    out[11] = o.T; out[12] = o.ntx; out[13] = o.nty; out[14] = o.NPX; out[15] = o.NPY; out[16] = o.OW; out[17] = o.OH;
}
// CNV_ONCHIP_PROF=1 (read when the solver is created): clock ticks thread 0 of every CTA of the on-chip kernel spent in the
// phases of its pass loop during the last solve -- out[cta][8]: 0 wait for all norms, 1 fold + decide, 2 sweeps, 3 store + flag,
// 4 wait for the neighbours, 5 halo reload.  Returns the number of CTAs written (0 = profiling off).
int cnv_poisson_onchip_profile(cnv_poisson *p, unsigned long long *out, int max_ctas) { return p->s->onchip_profile(out, max_ctas); }
int cnv_poisson_prepare(cnv_poisson *p, const double *f_dev, int ldf, double fsign, void *stream)
{
    const PassGeom &g = p->s->geom();
    p->s->peer_quiesce((cudaStream_t)stream);  // multi-GPU peer path: no neighbour push may land after the re-initialisation
    launch_prep_rhs(f_dev, g.nrows, g.ncols, ldf, fsign, p->s->consts().pscale, p->s->rhs(), p->s->buffer(0), p->s->buffer(1),
                    g.ld, (cudaStream_t)stream);
    p->s->zero_extra_buffer((cudaStream_t)stream);
    p->s->peer_ready((cudaStream_t)stream);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int cnv_poisson_upload(cnv_poisson *p, const double *f_host, double fsign, void *stream)
{
    // host f -> the solver's rhs array (pitched) with one 2-D copy, then scaled in place; no staging allocation
    const PassGeom &g = p->s->geom();
    cudaStream_t st = (cudaStream_t)stream;
    p->s->peer_quiesce(st);
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(p->s->rhs(), sizeof(double) * g.ld, f_host, sizeof(double) * g.ncols,
                                     sizeof(double) * g.ncols, g.nrows, cudaMemcpyHostToDevice, st));
    launch_prep_rhs(p->s->rhs(), g.nrows, g.ncols, g.ld, fsign, p->s->consts().pscale, p->s->rhs(), p->s->buffer(0),
                    p->s->buffer(1), g.ld, st);
    p->s->zero_extra_buffer(st);
    p->s->peer_ready(st);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    return 0;
}
// Slab solvers: the OWNED rows of f from a host array (own_rows x ncols, dense; page-locked for a true DMA) straight into
// the right-hand-side array, halo rows from the slab neighbours (the library's NCCL communicator, attached with
// cnv_poisson_attach_comm), scaling in place, zero iterate -- no staging array, nothing synchronises.
int cnv_poisson_upload_owned(cnv_poisson *p, const double *f_owned_host, double fsign, void *stream)
{
    const PassGeom &g = p->s->geom();
    cudaStream_t st = (cudaStream_t)stream;
    const bool slab = g.own_lo > 0 || g.own_hi < g.nrows;
    if (slab && !p->s->has_comm()) return 1;
    p->s->peer_quiesce(st);
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(p->s->rhs() + (size_t)g.own_lo * g.ld, sizeof(double) * g.ld, f_owned_host,
                                     sizeof(double) * g.ncols, sizeof(double) * g.ncols, g.own_hi - g.own_lo,
                                     cudaMemcpyHostToDevice, st));
    if (slab) p->s->exchange_halos(p->s->rhs(), g.HY, st);
    launch_prep_rhs(p->s->rhs(), g.nrows, g.ncols, g.ld, fsign, p->s->consts().pscale, p->s->rhs(), p->s->buffer(0),
                    p->s->buffer(1), g.ld, st);
    p->s->zero_extra_buffer(st);
    p->s->peer_ready(st);
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
    return 0;
}
// the owned rows of iterate buffer `which` into a host array (own_rows x ncols); valid once `stream` has drained
int cnv_poisson_download_owned_async(cnv_poisson *p, int which, double *u_owned_host, void *stream)
{
    const PassGeom &g = p->s->geom();
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(u_owned_host, sizeof(double) * g.ncols, p->s->buffer(buf_index(p, which)) + (size_t)g.own_lo * g.ld,
                                     sizeof(double) * g.ld, sizeof(double) * g.ncols, g.own_hi - g.own_lo, cudaMemcpyDeviceToHost,
                                     (cudaStream_t)stream));
    return 0;
}
int cnv_poisson_solve(cnv_poisson *p, int itmax, double tol, void *stream, int *k, double *e, int *sweeps, int *passes,
                      int *result_buf)
{
    PoissonResult r = p->s->solve(itmax, tol, (cudaStream_t)stream, result_buf, false);
    if (k) *k = r.k;
    if (e) *e = r.e;
    if (sweeps) *sweeps = r.sweeps;
    if (passes) *passes = r.passes;
    return r.status;
}
// residual history of the enqueue()-driven paths: hist[k] = global L1 update norm of sweep k
void cnv_poisson_enable_history(cnv_poisson *p, int capacity) { p->s->enable_history(capacity); }
int cnv_poisson_read_history(cnv_poisson *p, double *out, int n)
{
    if (!p->s->history()) return 1;
    CNV_CUDA_CHECK(cudaMemcpy(out, p->s->history(), sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}
void cnv_poisson_reset(cnv_poisson *p, int itmax, double tol, void *stream) { p->s->reset_ctl(itmax, tol, (cudaStream_t)stream); }
void cnv_poisson_enqueue(cnv_poisson *p, int npasses, void *stream) { p->s->enqueue_passes(npasses, (cudaStream_t)stream); }
void cnv_poisson_enqueue_decide(cnv_poisson *p, void *stream) { p->s->enqueue_decide((cudaStream_t)stream); }
void cnv_poisson_set_distributed(cnv_poisson *p, int on) { p->s->set_distributed(on != 0); }
void cnv_poisson_state(cnv_poisson *p, void *stream, int *state, double *e)
{
    PoissonCtl c = p->s->read_ctl((cudaStream_t)stream);
    state[0] = c.state; state[1] = c.cur; state[2] = c.sweeps; state[3] = c.passes; state[4] = c.result_k; state[5] = c.redo;
    if (e) { e[0] = c.result_e; e[1] = c.last_e; }
}
// ---- native NCCL communicator for the slab path (one per process; bootstrap id exchanged by the caller) ----
struct cnv_comm {
    SlabComm c;
};
int cnv_comm_unique_id(unsigned char *out128)
{
    const NcclApi &n = nccl_api();
    if (!n.ok) return 1;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return 2;
    std::memcpy(out128, &id, 128);
    return 0;
}
cnv_comm *cnv_comm_create(int rank, int world, const unsigned char *id128)
{
    const NcclApi &n = nccl_api();
    if (!n.ok) return nullptr;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    if (n.CommInitRank(&comm, world, id, rank) != ncclSuccess) return nullptr;
    cnv_comm *c = new cnv_comm;
    c->c.comm = comm;
    c->c.rank = rank;
    c->c.world = world;
    return c;
}
void cnv_comm_destroy(cnv_comm *c)
{
    if (!c) return;
    if (c->c.comm) nccl_api().CommDestroy((ncclComm_t)c->c.comm);
    delete c;
}
void cnv_poisson_attach_comm(cnv_poisson *p, cnv_comm *c) { p->s->attach_comm(c->c); }
// ---- peer-memory path (CUDA IPC): see PoissonSolver::peer_* ----
void cnv_poisson_peer_export(cnv_poisson *p, unsigned char *out256) { p->s->peer_export(out256); }
void cnv_poisson_peer_push_counts(cnv_poisson *p, int rank, int world, long long *low, long long *high)
{
    p->s->peer_push_counts(rank, world, low, high);
}
int cnv_poisson_peer_import(cnv_poisson *p, int rank, int world, const unsigned char *handles, const int *layout)
{
    return p->s->peer_import(rank, world, handles, layout);
}
void cnv_poisson_peer_disable(cnv_poisson *p) { p->s->peer_disable(); }
// end of the peer path: enqueue the quiesce wait (every push into this rank has landed) ...
void cnv_poisson_peer_quiesce(cnv_poisson *p, void *stream) { p->s->peer_quiesce((cudaStream_t)stream); }
// ... and, after a device synchronisation on every rank and a barrier of all ranks, unmap the peers' buffers
void cnv_poisson_peer_close(cnv_poisson *p) { p->s->peer_close(); }
int cnv_poisson_peer_trace(cnv_poisson *p, int passes) { return p->s->peer_trace_enable(passes); }
void cnv_poisson_peer_trace_read(cnv_poisson *p, unsigned long long *out, long long n) { p->s->peer_trace_read(out, (size_t)n); }
int cnv_poisson_peer_enabled(cnv_poisson *p) { return p->s->peer_enabled() ? 1 : 0; }
void cnv_poisson_enqueue_dist(cnv_poisson *p, int npasses, void *stream) { p->s->enqueue_passes_dist(npasses, (cudaStream_t)stream); }
void cnv_poisson_exchange_halos(cnv_poisson *p, double *field_dev, int depth, void *stream)
{
    p->s->exchange_halos(field_dev, depth, (cudaStream_t)stream);
}

int cnv_poisson_download(cnv_poisson *p, int which, double *u_host, void *stream)
{
    const PassGeom &g = p->s->geom();
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(u_host, sizeof(double) * g.ncols, p->s->buffer(buf_index(p, which)), sizeof(double) * g.ld,
                                     sizeof(double) * g.ncols, g.nrows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CNV_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// same copy without the synchronisation: u_host (pinned) is valid once `stream` has drained; lets a caller keep several
// solves in flight on different streams (copy of one overlapping the sweeps of another)
int cnv_poisson_download_async(cnv_poisson *p, int which, double *u_host, void *stream)
{
    const PassGeom &g = p->s->geom();
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(u_host, sizeof(double) * g.ncols, p->s->buffer(buf_index(p, which)), sizeof(double) * g.ld,
                                     sizeof(double) * g.ncols, g.nrows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

int cnv_poisson_host(const double *f, int nrows, int ncols, double dx, double dy, int itmax, double tol, double beta, int T,
                     double *u, int *k, double *e, double *history)
{
    require_device();
    if (nrows < 3 || ncols < 3) {
        std::printf("** Error: invalid parameter **\n");  // src/linearalg.c:58-62 wording
        std::exit(1);
    }
    // solver object and its device arrays are kept between calls (one call per time step from the drop-in poisson_SOR_log);
    // f goes straight into the solver's right-hand-side array (one pitched copy, DMA when f is page-locked: cnv_host_alloc /
    // the drop-in allocm) and is scaled in place; psi comes straight out of the iterate buffer
    PoissonSolver &s = *cached_solver(nrows, ncols, T);
    s.set_consts(dx, dy, beta);
    const int ld = s.ld();
    cudaStream_t st = 0;
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(s.rhs(), sizeof(double) * ld, f, sizeof(double) * ncols, sizeof(double) * ncols, nrows,
                                     cudaMemcpyHostToDevice, st));
    int buf = 0;
    PoissonResult r = s.solve_from(s.rhs(), ld, 1.0, itmax, tol, st, &buf, history != nullptr);
    CNV_CUDA_CHECK(cudaMemcpy2DAsync(u, sizeof(double) * ncols, s.buffer(buf), sizeof(double) * ld, sizeof(double) * ncols, nrows,
                                     cudaMemcpyDeviceToHost, st));
    if (history) CNV_CUDA_CHECK(cudaMemcpyAsync(history, s.history(), sizeof(double) * (size_t)r.sweeps, cudaMemcpyDeviceToHost, st));
    CNV_CUDA_CHECK(cudaStreamSynchronize(st));
    if (k) *k = r.k;
    if (e) *e = r.e;
    return r.status;
}
// drops the solver objects cnv_poisson_host keeps between calls (frees their device memory)
void cnv_poisson_host_cache_clear(void)
{
    for (CachedSolver &c : solver_cache()) delete c.s;
    solver_cache().clear();
}

// ---- time stepping -----------------------------------------------------------------------------
static cnv_sim *sim_create(const Config *cfg, int T, int rank, int world)
{
    require_device();
    cnv_sim *s = new cnv_sim;
    s->cfg = *cfg;
    const int gn = cfg->nx;  // global rows: first index of every field (initm(nx, ny), src/main.c:178)
    s->ncols = cfg->ny;
    s->dx = (double)cfg->Lx / cfg->nx;  // src/main.c:138-139 (not L/(n-1))
    s->dy = (double)cfg->Ly / cfg->ny;
    s->beta = cnv_sor_beta(cfg->nx, cfg->ny);
    s->inv_re = 1. / cfg->Re;
    if (cfg->poisson_type != 1 && cfg->poisson_type != 2) {
        std::printf("Error - invalid option for Poisson solver\n");  // src/main.c:359
        std::exit(1);
    }
    const int o = cfg->order;
    if (gn < o || s->ncols < o || !fd_make_table(s->ncols, o, 1, s->dx, &s->d1x) ||
        !fd_make_table(gn, o, 1, s->dy, &s->d1y) || !fd_make_table(s->ncols, o, 2, s->dx, &s->d2x) ||
        !fd_make_table(gn, o, 2, s->dy, &s->d2y)) {
        std::printf("** Error: valid orders are 2, 4 or 6 **\n");  // src/finitediff.c:150
        std::exit(1);
    }
    // slab of rows [r0, r1) with 2T halo rows on interior edges (>= 3 = the order-6 stencil reach)
    if (T <= 0) T = std::getenv("CNV_POISSON_T") ? std::atoi(std::getenv("CNV_POISSON_T")) : 8;
    if (world > 1 && T < 2) {
        std::printf("** Error: slab decomposition needs a temporal block depth >= 2 (halo >= 3 rows) **\n");
        std::exit(1);
    }
    const int base = gn / world, rem = gn % world;
    const int r0 = rank * base + (rank < rem ? rank : rem), r1 = r0 + base + (rank < rem ? 1 : 0);
    const int hlo = rank > 0 ? 2 * T : 0, hhi = rank < world - 1 ? 2 * T : 0;
    s->map.grow0 = r0 - hlo;
    s->map.nloc = (r1 - r0) + hlo + hhi;
    s->map.gnrows = gn;
    s->map.own_lo = hlo;
    s->map.own_hi = hlo + (r1 - r0);
    s->nrows = s->map.nloc;
    s->ps = new PoissonSolver(s->map.nloc, s->ncols, T, s->map.grow0, gn, s->map.own_lo, s->map.own_hi);
    s->ps->set_consts(s->dx, s->dy, cfg->poisson_type == 2 ? s->beta : 1.0);
    s->ps->set_distributed(world > 1);
    s->ph.s = s->ps;
    s->ld = s->ps->ld();
    const size_t bytes = sizeof(double) * (size_t)s->nrows * s->ld;
    for (double **p : {&s->u, &s->v, &s->w, &s->w2}) {
        CNV_CUDA_CHECK(cudaMalloc(p, bytes));
        CNV_CUDA_CHECK(cudaMemset(*p, 0, bytes));
    }
    // initial condition: interior u = ui, v = vi; w = psi = 0 (src/main.c:178-181, :214-221); halo rows included
    std::vector<double> h((size_t)s->nrows * s->ncols, 0.0);
    for (int pass = 0; pass < 2; pass++) {
        const double val = pass == 0 ? cfg->ui : cfg->vi;
        for (int li = 0; li < s->nrows; li++) {
            const int gi = s->map.grow0 + li;
            if (gi < 1 || gi > gn - 2) continue;
            for (int j = 1; j < s->ncols - 1; j++) h[(size_t)li * s->ncols + j] = val;
        }
        CNV_CUDA_CHECK(cudaMemcpy2D(pass == 0 ? s->u : s->v, sizeof(double) * s->ld, h.data(), sizeof(double) * s->ncols,
                                    sizeof(double) * s->ncols, s->nrows, cudaMemcpyHostToDevice));
    }
    CNV_CUDA_CHECK(cudaMalloc(&s->cont_partial, sizeof(double) * 2 * continuity_blocks(s->map.own_hi - s->map.own_lo, s->ncols)));
    CNV_CUDA_CHECK(cudaMalloc(&s->cont_result, sizeof(double) * 2));
    CNV_CUDA_CHECK(cudaMalloc(&s->cont_ticket, sizeof(unsigned)));
    CNV_CUDA_CHECK(cudaMemset(s->cont_ticket, 0, sizeof(unsigned)));
    CNV_CUDA_CHECK(cudaMallocHost(&s->h_cont, sizeof(double) * 2));
    return s;
}

cnv_sim *cnv_sim_create(const Config *cfg, int T) { return sim_create(cfg, T, 0, 1); }
cnv_sim *cnv_sim_create_slab(const Config *cfg, int T, int rank, int world) { return sim_create(cfg, T, rank, world); }

// out[0..7] = grow0, nloc, own_lo, own_hi, ld, ncols, T, global rows
void cnv_sim_layout(const cnv_sim *s, int *out)
{
    out[0] = s->map.grow0; out[1] = s->map.nloc; out[2] = s->map.own_lo; out[3] = s->map.own_hi;
    out[4] = s->ld; out[5] = s->ncols; out[6] = s->ps->T(); out[7] = s->map.gnrows;
}
cnv_poisson *cnv_sim_poisson(cnv_sim *s) { return &s->ph; }
// device pointers of the slab-local fields: 0 u, 1 v, 2 w, 3 continuity result (max, min)
double *cnv_sim_field_ptr(cnv_sim *s, int which)
{
    switch (which) {
    case 0: return s->u;
    case 1: return s->v;
    case 2: return s->w;
    case 3: return s->cont_result;
    default: return nullptr;
    }
}
void cnv_sim_set_psi_buf(cnv_sim *s, int which) { s->psi_buf = which >= 0 && which < s->ps->num_buffers() ? which : which & 1; }

// One phase of a time step, asynchronously on `stream` (multi-GPU orchestration interleaves halo exchanges):
//   0  BCs + wall vorticity on owned ring cells        (needs u, v halos)
//   1  derivatives + Euler + Poisson right-hand side    (needs w halos); w <- w_new
//   2  velocities from psi (buffer set by cnv_sim_set_psi_buf; needs psi halos)
//   3  continuity max/min over the owned rows            (needs u, v halos) -> field_ptr(3)
void cnv_sim_phase(cnv_sim *s, int phase, void *stream)
{
    const Config &c = s->cfg;
    const double bc[8] = {c.u1, c.u2, c.u3, c.u4, c.v1, c.v2, c.v3, c.v4};
    cudaStream_t st = (cudaStream_t)stream;
    switch (phase) {
    case 0:
        launch_ring_bc_vorticity(s->u, s->v, s->w, s->map, s->ncols, s->ld, bc, s->d1x, s->d1y, st);
        break;
    case 1:
        launch_euler_fused(s->w, s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->d2x, s->d2y, s->inv_re, c.dt,
                           s->ps->consts().pscale, s->w2, s->ps->rhs(), st);
        std::swap(s->w, s->w2);
        break;
    case 2:
        launch_velocity(s->ps->buffer(s->psi_buf), s->map, s->ncols, s->ld, s->d1x, s->d1y, s->u, s->v, s->ld, st);
        break;
    case 3:
        launch_continuity(s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->cont_partial, s->cont_ticket,
                          s->cont_result, st);
        break;
    default:
        return;
    }
    count_launch(1);
    CNV_CUDA_CHECK(cudaGetLastError());
}

void cnv_sim_destroy(cnv_sim *s)
{
    if (!s) return;
    delete s->ps;
    delete s->pp;
    cudaFree(s->u); cudaFree(s->v); cudaFree(s->w); cudaFree(s->w2);
    cudaFree(s->cont_partial); cudaFree(s->cont_result); cudaFree(s->cont_ticket);
    cudaFreeHost(s->h_cont);
    delete s;
}

void cnv_sim_set_diagnostics(cnv_sim *s, int on) { s->diag = on != 0; }

int cnv_sim_step(cnv_sim *s, int nsteps, int *k, double *e, double *cont_max, double *cont_min)
{
    const Config &c = s->cfg;
    const double bc[8] = {c.u1, c.u2, c.u3, c.u4, c.v1, c.v2, c.v3, c.v4};
    cudaStream_t st = s->stream;
    const size_t bytes = sizeof(double) * (size_t)s->nrows * s->ld;
    for (int t = 0; t < nsteps; t++) {
        // BCs + wall vorticity (src/main.c:283-320)
        launch_ring_bc_vorticity(s->u, s->v, s->w, s->map, s->ncols, s->ld, bc, s->d1x, s->d1y, st);
        // vorticity derivatives + Euler on all points + Poisson right-hand side (:323-348)
        launch_euler_fused(s->w, s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->d2x, s->d2y, s->inv_re, c.dt,
                           s->ps->consts().pscale, s->w2, s->ps->rhs(), st);
        std::swap(s->w, s->w2);
        count_launch(2);
        // zero initial guess every step (fresh initm in the reference, src/poisson.c:229)
        CNV_CUDA_CHECK(cudaMemsetAsync(s->ps->buffer(0), 0, bytes, st));
        PoissonResult r = s->ps->solve(c.poisson_max_it, c.poisson_tol, st, &s->psi_buf, false);
        s->sweeps += r.sweeps;
        s->passes += r.passes;
        if (k) k[t] = r.k;
        if (e) e[t] = r.e;
        if (r.status != 0) return t + 1;  // reference: log the error and exit(1), src/poisson.c:280-284
        // velocities from the streamfunction on all points (:366-383) and the continuity diagnostic (:387-408).  (A fused
        // form that recomputed the halo velocities from a psi tile with a 6-cell halo -- 56 instead of 72 bytes per cell -- was
        // bit-identical but slower, 203 vs 164 us at 4096^2: both kernels are bound by fp64 issue, not HBM.  Removed in round 2.)
        const bool want_cont = s->diag && (cont_max || cont_min);
        launch_velocity(s->ps->buffer(s->psi_buf), s->map, s->ncols, s->ld, s->d1x, s->d1y, s->u, s->v, s->ld, st);
        count_launch(1);
        if (want_cont) {
            launch_continuity(s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->cont_partial, s->cont_ticket,
                              s->cont_result, st);
            count_launch(1);
            CNV_CUDA_CHECK(cudaMemcpyAsync(s->h_cont, s->cont_result, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
            CNV_CUDA_CHECK(cudaStreamSynchronize(st));
            if (cont_max) cont_max[t] = s->h_cont[0];
            if (cont_min) cont_min[t] = s->h_cont[1];
        }
        s->steps++;
    }
    CNV_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// One time step of a slab simulation (world > 1), entirely inside the library: the loop body of src/main.c:283-395 with the halo
// exchanges of the slab decomposition on the library's NCCL communicator (attach it to cnv_sim_poisson() first) and the
// distributed Poisson solve (peer-memory path if it was set up, the NCCL group per pass otherwise).  Every rank calls it
// with the same arguments and gets the same k / e / continuity values.  Returns 0, or (index of the step whose Poisson solve
// hit itmax) + 1 on every rank, or -1 without a communicator.
int cnv_sim_step_slab(cnv_sim *s, int nsteps, int *k, double *e, double *cont_max, double *cont_min)
{
    if (!s->ps->has_comm()) {
        std::printf("** Error: slab time stepping needs the communicator (cnv_poisson_attach_comm) **\n");
        return -1;
    }
    const Config &c = s->cfg;
    const double bc[8] = {c.u1, c.u2, c.u3, c.u4, c.v1, c.v2, c.v3, c.v4};
    cudaStream_t st = s->stream;
    PoissonSolver &P = *s->ps;
    const int H3 = 3;  // reach of the order-6 stencils (orders 2 and 4 need less)
    for (int t = 0; t < nsteps; t++) {
        launch_ring_bc_vorticity(s->u, s->v, s->w, s->map, s->ncols, s->ld, bc, s->d1x, s->d1y, st);
        P.exchange_halos(s->w, H3, st);
        launch_euler_fused(s->w, s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->d2x, s->d2y, s->inv_re, c.dt,
                           P.consts().pscale, s->w2, P.rhs(), st);
        std::swap(s->w, s->w2);
        count_launch(2);
        P.exchange_halos(P.rhs(), P.geom().HY, st);  // 2T rows: the solver recomputes its halo rows
        // zero initial guess every step (fresh initm in the reference, src/poisson.c:229), peer-safe
        P.peer_quiesce(st);
        launch_prep_rhs(nullptr, s->nrows, s->ncols, 0, 1.0, P.consts().pscale, P.rhs(), P.buffer(0), P.buffer(1), s->ld, st);
        P.zero_extra_buffer(st);
        P.peer_ready(st);
        count_launch(1);
        PoissonResult r = P.solve(c.poisson_max_it, c.poisson_tol, st, &s->psi_buf, false);
        s->sweeps += r.sweeps;
        s->passes += r.passes;
        if (k) k[t] = r.k;
        if (e) e[t] = r.e;
        if (r.status != 0) return t + 1;
        P.exchange_halos(P.buffer(s->psi_buf), H3, st);  // (a "redo" pass leaves the result's halo rows stale)
        launch_velocity(P.buffer(s->psi_buf), s->map, s->ncols, s->ld, s->d1x, s->d1y, s->u, s->v, s->ld, st);
        count_launch(1);
        P.exchange_halos(s->u, H3, st);
        P.exchange_halos(s->v, H3, st);
        if (s->diag && (cont_max || cont_min)) {
            launch_continuity(s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->cont_partial, s->cont_ticket, s->cont_result, st);
            count_launch(1);
            P.allreduce_max_min(s->cont_result, st);
            CNV_CUDA_CHECK(cudaMemcpyAsync(s->h_cont, s->cont_result, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
            CNV_CUDA_CHECK(cudaStreamSynchronize(st));
            if (cont_max) cont_max[t] = s->h_cont[0];
            if (cont_min) cont_min[t] = s->h_cont[1];
        }
        s->steps++;
    }
    CNV_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Whole fields of a slab simulation on rank 0 (host arrays of nx*ny doubles there, ignored on the other ranks; a NULL on rank 0
// drops that field after receiving it): every rank sends the owned rows of all four fields over the communicator.
int cnv_sim_gather_fields_slab(cnv_sim *s, double *psi, double *w, double *u, double *v)
{
    if (!s->ps->has_comm()) return -1;
    const int world = s->ps->comm().world, rank = s->ps->comm().rank;
    const double *src[4] = {s->ps->buffer(s->psi_buf), s->w, s->u, s->v};
    double *dst[4] = {psi, w, u, v};
    double *stage = nullptr;
    if (rank == 0) {
        const int rows = (s->map.gnrows + world - 1) / world;
        CNV_CUDA_CHECK(cudaMalloc(&stage, sizeof(double) * (size_t)rows * s->ld));
    }
    for (int i = 0; i < 4; i++) s->ps->gather_field_to_root(src[i], stage, rank == 0 ? dst[i] : nullptr, s->stream);
    CNV_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    if (stage) cudaFree(stage);
    return 0;
}

// Stencil phase only (measurement / profiling): BCs + wall vorticity, fused derivative + Euler, velocity
// recovery from the current psi, continuity diagnostic -- `reps` times, no Poisson solve.
void cnv_sim_stencil_phase(cnv_sim *s, int reps, void *stream)
{
    const Config &c = s->cfg;
    const double bc[8] = {c.u1, c.u2, c.u3, c.u4, c.v1, c.v2, c.v3, c.v4};
    cudaStream_t st = (cudaStream_t)stream;
    for (int t = 0; t < reps; t++) {
        launch_ring_bc_vorticity(s->u, s->v, s->w, s->map, s->ncols, s->ld, bc, s->d1x, s->d1y, st);
        launch_euler_fused(s->w, s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->d2x, s->d2y, s->inv_re, c.dt,
                           s->ps->consts().pscale, s->w2, s->ps->rhs(), st);
        std::swap(s->w, s->w2);
        launch_velocity(s->ps->buffer(s->psi_buf), s->map, s->ncols, s->ld, s->d1x, s->d1y, s->u, s->v, s->ld, st);
        launch_continuity(s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->cont_partial, s->cont_ticket,
                          s->cont_result, st);
        count_launch(4);
    }
    CNV_CUDA_CHECK(cudaGetLastError());
}

int cnv_sim_get_fields(cnv_sim *s, double *psi, double *w, double *u, double *v)
{
    CNV_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    const double *src[4] = {s->ps->buffer(s->psi_buf), s->w, s->u, s->v};
    double *dst[4] = {psi, w, u, v};
    for (int i = 0; i < 4; i++)
        if (dst[i])
            CNV_CUDA_CHECK(cudaMemcpy2D(dst[i], sizeof(double) * s->ncols, src[i], sizeof(double) * s->ld,
                                        sizeof(double) * s->ncols, s->nrows, cudaMemcpyDeviceToHost));
    return 0;
}

int cnv_sim_set_fields(cnv_sim *s, const double *psi, const double *w, const double *u, const double *v)
{
    double *dst[4] = {s->ps->buffer(s->psi_buf), s->w, s->u, s->v};
    const double *src[4] = {psi, w, u, v};
    for (int i = 0; i < 4; i++)
        if (src[i])
            CNV_CUDA_CHECK(cudaMemcpy2D(dst[i], sizeof(double) * s->ld, src[i], sizeof(double) * s->ncols,
                                        sizeof(double) * s->ncols, s->nrows, cudaMemcpyHostToDevice));
    return 0;
}

// Pressure from the current velocity field: p = poisson(-(dudx^2 + dvdy^2 + 2 dudy dvdx)) with the configured
// Poisson variant (src/main.c:421-427 uses an FFT solver that the reference never shipped; here it is the same
// SOR / Gauss-Seidel solve as for psi: zero Dirichlet ring, zero initial guess, first sweep with e < tol).
// itmax <= 0 / tol <= 0 take the configuration's values.  Returns 0, 1 (itmax reached) or -1 (slab simulations).
int cnv_sim_pressure(cnv_sim *s, int itmax, double tol, double *p_host, int *k, double *e)
{
    if (s->map.nloc != s->map.gnrows) {
        std::printf("** Error: pressure is available on single-GPU simulations only **\n");
        return -1;
    }
    const Config &c = s->cfg;
    cudaStream_t st = s->stream;
    if (!s->pp) {
        s->pp = new PoissonSolver(s->nrows, s->ncols, s->ps->T());
        s->pp->set_consts(s->dx, s->dy, c.poisson_type == 2 ? s->beta : 1.0);
    }
    launch_pressure_rhs(s->u, s->v, s->map, s->ncols, s->ld, s->d1x, s->d1y, s->pp->consts().pscale, nullptr, s->pp->rhs(), s->ld, st);
    count_launch(1);
    const size_t bytes = sizeof(double) * (size_t)s->nrows * s->ld;
    CNV_CUDA_CHECK(cudaMemsetAsync(s->pp->buffer(0), 0, bytes, st));
    int buf = 0;
    PoissonResult r = s->pp->solve(itmax > 0 ? itmax : c.poisson_max_it, tol > 0 ? tol : c.poisson_tol, st, &buf, false);
    if (k) *k = r.k;
    if (e) *e = r.e;
    if (p_host) {
        CNV_CUDA_CHECK(cudaStreamSynchronize(st));
        CNV_CUDA_CHECK(cudaMemcpy2D(p_host, sizeof(double) * s->ncols, s->pp->buffer(buf), sizeof(double) * s->ld,
                                    sizeof(double) * s->ncols, s->nrows, cudaMemcpyDeviceToHost));
    }
    return r.status;
}

void cnv_sim_counters(cnv_sim *s, long long *out)
{
    out[0] = s->sweeps; out[1] = s->passes; out[2] = s->steps;
}

}  // extern "C"
