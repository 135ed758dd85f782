// kernels.h -- internal C++ interface between the CUDA translation units of libcnavier_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fd_coeffs.h"
#include "poisson_onchip.h"
#include "poisson_plan.h"

// The reference reports failures with a message and exit(1) (src/poisson.c:280-284,
// src/linearalg.c:58-79); CUDA errors follow the same convention.
#define CNV_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t err__ = (expr);                                                                       \
        if (err__ != cudaSuccess) {                                                                       \
            std::printf("** Error: CUDA failure %s at %s:%d (%s) **\n", cudaGetErrorString(err__), __FILE__, \
                        __LINE__, #expr);                                                                 \
            std::fflush(stdout);                                                                          \
            std::exit(1);                                                                                 \
        }                                                                                                 \
    } while (0)

namespace cnv {

constexpr int kMaxDevices = 64;
int current_device_slot();  // cudaGetDevice, clamped to [0, kMaxDevices): index of per-device caches

// A slab of a row-decomposed grid: the local array has nloc rows, local row 0 is global row grow0 of
// gnrows; rows [own_lo, own_hi) (local) are owned, the others are halo rows.  Single GPU: {n, 0, n, 0, n}.
struct RowMap {
    int nloc, grow0, gnrows, own_lo, own_hi;
};

// ---- stencil_kernels.cu ----
void launch_apply(const double *A, int nrows, int ncols, int lda, int axis, const FdTable &t, double *out, int ldo,
                  double scale, cudaStream_t s);
void launch_ring_bc_vorticity(double *u, double *v, double *w, const RowMap &m, int ncols, int ld, const double bc[8],
                              const FdTable &d1x, const FdTable &d1y, cudaStream_t s);
void launch_euler_fused(const double *w, const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x,
                        const FdTable &d1y, const FdTable &d2x, const FdTable &d2y, double inv_re, double dt, double pscale,
                        double *w_new, double *rhs, cudaStream_t s);
void launch_euler_pointwise(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
                            const double *u, const double *v, size_t n, double inv_re, double dt, cudaStream_t s);
void launch_pointwise_addsub(const double *a, const double *b, double *out, size_t n, int sub, cudaStream_t s);
void launch_velocity(const double *psi, const RowMap &m, int ncols, int ldp, const FdTable &d1x, const FdTable &d1y, double *u,
                     double *v, int ld, cudaStream_t s);
void launch_pressure_rhs(const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x, const FdTable &d1y,
                         double pscale, double *f_out, double *rhs, int ldo, cudaStream_t s);
int continuity_blocks(int nrows, int ncols);
void launch_continuity(const double *u, const double *v, const RowMap &m, int ncols, int ld, const FdTable &d1x, const FdTable &d1y,
                       double *partial, unsigned *ticket, double *result, cudaStream_t s);
void launch_prep_rhs(const double *f, int nrows, int ncols, int ldf, double sign, double pscale, double *rhs, double *psi0,
                     double *psi1, int ld, cudaStream_t s);

// ---- poisson_onchip.cu: the whole solve in one persistent launch, iterate resident in registers (<= ~1.5 M cells) ----
void launch_onchip(const OnchipGeom &g, const RelaxConsts &rc, double *b0, double *b1, double *b2, const double *rhs, PoissonCtl *ctl,
                   unsigned long long *flags, double *partials, double *hist, cudaStream_t s, unsigned long long *prof = nullptr);

// ---- poisson.cu ----
struct PoissonResult {
    int status;  // 0 converged, 1 itmax reached (the reference exits the process here)
    int k;       // logged iteration number = sweeps - 1
    int sweeps;
    int passes;
    double e;    // L1 update norm of the last sweep
};

// Slab communicator of the multi-GPU path (poisson.cu): one NCCL communicator over the ranks of the node.
struct SlabComm {
    void *comm = nullptr;  // ncclComm_t
    int rank = 0, world = 1;
};

class PoissonSolver {
public:
    // local array of nrows x ncols; rows [own_lo, own_hi) are owned (== all rows on one GPU)
    PoissonSolver(int nrows, int ncols, int T, int grow0 = 0, int gnrows = -1, int own_lo = 0, int own_hi = -1);
    ~PoissonSolver();
    PoissonSolver(const PoissonSolver &) = delete;

    int ld() const { return geom_.ld; }
    int T() const { return T_; }
    const PassGeom &geom() const { return geom_; }
    bool onchip() const { return use_onchip_; }
    int onchip_profile(unsigned long long *out, int max_ctas);  // CNV_ONCHIP_PROF=1: 8 phase counters per CTA of the last solve  // solve() runs the persistent on-chip kernel (poisson_onchip.cu)
    const OnchipGeom &onchip_geom() const { return oc_; }
    double *rhs() { return rhs_; }                // device, pitch ld(): pscale * f
    double *buffer(int i) { return buf_[i]; }     // the iterate buffers: 0, 1 (and 2 for the on-chip kernel, whose stop decision lags a pass)
    int num_buffers() const { return use_onchip_ && buf_[2] ? 3 : 2; }
    void zero_extra_buffer(cudaStream_t s)  // third buffer (on-chip kernel): same initial state as 0 and 1
    {
        if (buf_[2]) CNV_CUDA_CHECK(cudaMemsetAsync(buf_[2], 0, (size_t)geom_.nrows * geom_.ld * sizeof(double), s));
    }
    double *history() { return hist_; }
    void enable_history(int cap);  // residual history for the enqueue()-driven paths (hist[k] = norm of sweep k)
    size_t launches() const { return launches_; }
    void set_consts(double dx, double dy, double beta);
    const RelaxConsts &consts() const { return rc_; }

    // rhs() must hold pscale*f and both buffers the initial iterate (zero ring); runs until
    // convergence / itmax with the reference's stopping rule and returns the result.
    // `result_buf` receives the index of the buffer holding psi.
    PoissonResult solve(int itmax, double tol, cudaStream_t s, int *result_buf, bool keep_history = false);
    // f (device, pitch ldf) -> prepared rhs + zero initial guess, then solve()
    PoissonResult solve_from(const double *f, int ldf, double fsign, int itmax, double tol, cudaStream_t s, int *result_buf,
                             bool keep_history = false);
    // enqueue `npasses` passes without any host synchronisation (benchmark / graph building block)
    void enqueue_passes(int npasses, cudaStream_t s);
    void reset_ctl(int itmax, double tol, cudaStream_t s);
    PoissonCtl read_ctl(cudaStream_t s);

    // multi-GPU hooks (NCCL path): when set, the last CTA only publishes local norms to
    // local_norms() and decide_kernel() must be enqueued after the all-reduce.
    void set_distributed(bool on) { distributed_ = on; }
    double *local_norms() { return norms_; }
    void enqueue_decide(cudaStream_t s);
    // Native distributed passes: per pass ONE NCCL group on the compute stream (2T halo rows to/from both slab
    // neighbours + an all-gather of the T norms as 64-byte send/recvs) and a decide kernel that sums the gathered
    // norms in rank order -- no host synchronisation, no second collective launch.
    void attach_comm(const SlabComm &c);
    bool has_comm() const { return comm_.comm != nullptr; }
    void enqueue_passes_dist(int npasses, cudaStream_t s);
    void exchange_halos(double *field, int depth, cudaStream_t s);  // any slab-local field with this solver's layout
    void allreduce_max_min(double *two_dev, cudaStream_t s);        // two[0] <- max, two[1] <- min over the ranks
    void gather_field_to_root(const double *field, double *stage_dev, double *host_out, cudaStream_t s);
    const SlabComm &comm() const { return comm_; }
    void restart_pass_counter() { dist_passes_ = 0; }
    // Peer-memory path (CUDA IPC over NVLink): boundary rows are stored into the neighbours' halos by the pass
    // kernel itself, norms are published in every rank's mailbox, each CTA derives the stop decision: one kernel
    // per pass, no collective launch.  enqueue_passes() takes this path once peer_import() succeeded.
    void peer_export(unsigned char *out256);
    void peer_push_counts(int rank, int world, long long *low, long long *high) const;
    int peer_import(int rank, int world, const unsigned char *handles, const int *layout);
    void peer_quiesce(cudaStream_t s);  // before re-initialising the iterate: every push launched so far has landed
    void peer_ready(cudaStream_t s);    // after re-initialising it: the neighbours may push into the new buffers
    bool peer_enabled() const { return links_.enabled != 0; }
    void peer_disable() { peer_close(); }
    void peer_close();  // unmap the peers' buffers (after quiesce + all-rank barrier, see poisson.cu)
    // diagnostics: record per-CTA timestamps of the first `passes` passes after every reset (tools/peer_trace.py)
    int peer_trace_enable(int passes);                       // returns CTAs per pass
    void peer_trace_read(unsigned long long *out, size_t n);  // n = passes * ctas * 6

private:
    int T_;
    PassGeom geom_;
    OnchipGeom oc_ = {};
    bool use_onchip_ = false;
    unsigned long long *oc_flags_ = nullptr;
    unsigned long long *oc_prof_ = nullptr;   // CNV_ONCHIP_PROF=1: per-CTA phase counters of the last solve
    double *oc_partials_ = nullptr;
    RelaxConsts rc_;
    double *buf_[3] = {nullptr, nullptr, nullptr};
    double *rhs_ = nullptr, *partials_ = nullptr, *hist_ = nullptr, *norms_ = nullptr;
    int hist_cap_ = 0;
    PoissonCtl *ctl_ = nullptr, *h_ctl_ = nullptr;
    cudaEvent_t ev_ = nullptr;
    int predicted_passes_ = 0;
    bool distributed_ = false;
    bool use_hist_ = false;
    size_t launches_ = 0;
    size_t smem_ = 0;
    int threads_ = 0;
    size_t smem_optin_ = 0;
    SlabComm comm_;
    double *gather_ = nullptr;  // [world][8] norms of every rank
    int dist_passes_ = 0;       // passes enqueued since the last reset (static exchange pattern)
    PeerLinks links_ = {};
    std::vector<void *> imported_;  // CUDA IPC mappings opened by peer_import
    void release_imports();
    PeerMailbox *mailbox_ = nullptr;
    PoissonCtl *ctlbuf_ = nullptr;
    unsigned long long peer_gidx_ = 0;  // passes launched since peer_import (global pass index)
    unsigned long long peer_epoch_ = 0; // re-initialisations of the iterate since peer_import
    unsigned long long *trace_ = nullptr;
    int trace_pass_ = 0;
};

}  // namespace cnv
