/*
 * cnavier_b200.h -- C ABI of libcnavier_b200.so, the B200 (sm_100a) implementation of cnavier's hot
 * path.  Plain pointers and sizes only; every entry point names the reference interface it serves.
 *
 * Conventions
 *   - fields are row-major double arrays A[i*ncols + j]; i is the first index of the reference's
 *     mtrx (A.M[i][j]), j the second.  "host" entry points take HOST buffers and stage them.
 *   - return value 0 = success unless stated otherwise.  Conditions for which the reference prints a
 *     message and calls exit(1) (Poisson itmax, invalid order) are reported as non-zero codes here;
 *     the drop-in layer (cnavier_dropin.h) turns them back into message + exit(1).
 *   - CUDA failures print "** Error: CUDA failure ... **" and exit(1) (reference convention).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails loudly.
 */
#ifndef CNAVIER_B200_H
#define CNAVIER_B200_H

#include <stddef.h>
#include "cnavier_dropin.h" /* Config */

#ifdef __cplusplus
extern "C" {
#endif

const char *cnv_version(void);
int cnv_device_count(void);               /* 0 when no CUDA device is visible */
unsigned long long cnv_launch_count(void); /* kernels launched by this library so far */
int cnv_set_device(int device);           /* one process per GPU: bind this library to a device */
int cnv_get_device(void);
void cnv_device_synchronize(void);
/* Host memory for fields (what the drop-in allocm / freem of include/linearalg.h:13-15 are backed by; the reference
 * allocates and frees every field once per call, src/linearalg.c:53-99).  Blocks of >= 1 MiB are page-locked and
 * recycled through size-keyed free lists (page-locking 134 MB costs more than a solve); contents are not zeroed.  Where
 * the platform tells which NUMA node the current device hangs off (sysfs), a page-locked block is placed on that node, so
 * that the copies of a GPU on the second socket do not cross the socket interconnect; cnv_host_numa_node() = that node
 * or -1. */
void *cnv_host_alloc(size_t bytes);
void cnv_host_free(void *p);
int cnv_host_is_pinned(const void *p);
int cnv_host_numa_node(void);

/* ---- scalars of the driver: src/main.c:134 (beta, truncated PI), :162/:276 (step count) ------- */
double cnv_sor_beta(int nx, int ny);
int cnv_num_steps(double tf, double dt);

/* ---- finite-difference operators ------------------------------------------------------------
 * cnv_diff_dense: the dense n x n matrix Diff1 (deriv=1) / Diff2 (deriv=2) returns
 * (src/finitediff.c:51-153, :178-292).  Host only.  Returns 1 for an order other than 2, 4, 6. */
int cnv_diff_dense(int n, int order, int deriv, double h, double *D);
/* cnv_apply_host: matrix-free DX (axis=1, along j) / DY (axis=0, along i) application that replaces
 * kronecker + reshape + mtrxmul (src/main.c:149-152, :298-304; src/linearalg.c:236-285, :353-408),
 * bit-identical to the dense route. */
int cnv_apply_host(const double *A, int nrows, int ncols, int axis, int deriv, int order, double h, double *out);

/* ---- pointwise operators: src/fluiddyn.c:71-102, :126-154, :179-207 -------------------------- */
int cnv_euler_host(double *w, const double *dwdx, const double *dwdy, const double *d2wdx2, const double *d2wdy2,
                   const double *u, const double *v, int nrows, int ncols, double Re, double dt);
int cnv_continuity_host(const double *dudx, const double *dvdy, int nrows, int ncols, double *out);
int cnv_vorticity_host(const double *a, const double *b, int nrows, int ncols, double *out); /* out = b - a */
/* L1 distance sum|a-b| over the grid: error(), src/poisson.c:34-60 */
double cnv_error_host(const double *a, const double *b, int nrows, int ncols);
/* Pressure-Poisson right-hand side f = dudx^2 + dvdy^2 + 2*dudy*dvdx of the recipe the reference leaves commented
 * out (src/main.c:421-427; `pressure()` is declared in include/fluiddyn.h:11 and defined nowhere): the four first
 * derivatives with the Diff1 closures, evaluated left to right with separately rounded operations. */
int cnv_pressure_rhs_host(const double *u, const double *v, int nrows, int ncols, int order, double dx, double dy, double *f_out);

/* ---- Poisson solve: src/poisson.c:62-285 -------------------------------------------------------
 * Solves lap(u) = f, zero Dirichlet ring, zero initial guess, red-black ordering ((i+j) even first,
 * the reference's OpenMP build), stopping at the first sweep whose L1 update norm is < tol.
 * beta = 1 gives the Gauss-Seidel variants (poisson / poisson_log).
 * T = temporal block depth (sweeps per HBM pass: 1, 2, 4, 8; 0 = default).
 * Returns 0 converged, 1 itmax reached (reference: message + exit(1)).
 * *k = the reference's logged iteration number (sweeps-1), *e = final norm. */
int cnv_poisson_host(const double *f, int nrows, int ncols, double dx, double dy, int itmax, double tol, double beta,
                     int T, double *u, int *k, double *e, double *history /* NULL or itmax doubles */);

/* cnv_poisson_host keeps its solver objects (device arrays, launch plan, pass-count predictor) between calls, the last
 * two grid shapes per device; this releases them. */
void cnv_poisson_host_cache_clear(void);

/* Device-resident solver object (benchmarks, time stepping, multi-GPU). */
typedef struct cnv_poisson cnv_poisson;
cnv_poisson *cnv_poisson_create(int nrows, int ncols, int T);
/* slab of a larger grid: local array rows [0,nrows) = global rows [grow0, grow0+nrows) of gnrows;
 * rows [own_lo, own_hi) (local indices) are owned, the rest are halo rows (>= 2T deep). */
cnv_poisson *cnv_poisson_create_slab(int nrows, int ncols, int T, int grow0, int gnrows, int own_lo, int own_hi);
void cnv_poisson_destroy(cnv_poisson *p);
void cnv_poisson_set_consts(cnv_poisson *p, double dx, double dy, double beta);
int cnv_poisson_ld(const cnv_poisson *p);                 /* pitch (doubles) of the device arrays */
double *cnv_poisson_rhs_ptr(cnv_poisson *p);              /* device: prepared right-hand side */
double *cnv_poisson_buf_ptr(cnv_poisson *p, int which);   /* device: iterate buffers 0/1 (2: on-chip kernel) */
int cnv_poisson_num_buffers(cnv_poisson *p);              /* 2, or 3 when the on-chip kernel is in use */
double *cnv_poisson_norms_ptr(cnv_poisson *p);            /* device: T per-sweep norms of the last pass */
/* out[0..9] = WS, HX, Wout, Hout, nstrips, nchunks, threads, smem bytes, T, pow2-path flag of the streaming kernel's plan;
   out[10..17] = 1 if cnv_poisson_solve runs the persistent on-chip kernel, then its T, tile grid ntx x nty, patches per tile
   NPX x NPY, output columns / rows per tile */
void cnv_poisson_plan_info(const cnv_poisson *p, long long *out);
/* diagnostics (CNV_ONCHIP_PROF=1 when the solver is created): clock ticks per phase of the on-chip kernel's pass loop in the last
 * solve, out[cta][8] = wait for all norms, fold + decide, sweeps, store + flag, wait for neighbours, halo reload, -, - */
int cnv_poisson_onchip_profile(cnv_poisson *p, unsigned long long *out, int max_ctas);
/* stage a host right-hand side f (nrows x ncols, dense) and zero the iterate; fsign = -1 solves with -f */
int cnv_poisson_upload(cnv_poisson *p, const double *f_host, double fsign, void *stream);
/* slab solvers: the OWNED rows (own_rows x ncols, dense, ideally page-locked) into the right-hand-side array, halo rows from
 * the slab neighbours over the attached communicator, scaled in place, iterate zeroed; returns 1 without a communicator */
int cnv_poisson_upload_owned(cnv_poisson *p, const double *f_owned_host, double fsign, void *stream);
int cnv_poisson_download_owned_async(cnv_poisson *p, int which, double *u_owned_host, void *stream);
/* same from a device array with pitch ldf */
int cnv_poisson_prepare(cnv_poisson *p, const double *f_dev, int ldf, double fsign, void *stream);
/* run to convergence / itmax (synchronises).  Returns 0 / 1 as cnv_poisson_host. */
int cnv_poisson_solve(cnv_poisson *p, int itmax, double tol, void *stream, int *k, double *e, int *sweeps, int *passes,
                      int *result_buf);
/* asynchronous building blocks: reset the device state machine, enqueue passes (each applies up to T
 * sweeps, no-ops once the state machine has stopped), read the state back (synchronises). */
/* residual history for the asynchronous building blocks: hist[k] = global L1 update norm of sweep k */
void cnv_poisson_enable_history(cnv_poisson *p, int capacity);
int cnv_poisson_read_history(cnv_poisson *p, double *out, int n);
void cnv_poisson_reset(cnv_poisson *p, int itmax, double tol, void *stream);
void cnv_poisson_enqueue(cnv_poisson *p, int npasses, void *stream);
void cnv_poisson_enqueue_decide(cnv_poisson *p, void *stream);
void cnv_poisson_set_distributed(cnv_poisson *p, int on);
/* state[0..5] = state (0 running, 1 converged, 2 itmax), cur buffer, sweeps, passes, k, redo */
void cnv_poisson_state(cnv_poisson *p, void *stream, int *state, double *e);
int cnv_poisson_download(cnv_poisson *p, int which, double *u_host, void *stream);       /* synchronises `stream` */
/* the same copy without the synchronisation (u_host pinned; valid once `stream` has drained): several solver objects
 * on different streams overlap the host copies of one solve with the sweeps of another */
int cnv_poisson_download_async(cnv_poisson *p, int which, double *u_host, void *stream);

/* Native multi-GPU plumbing for slab solvers (one process per GPU).  NCCL is loaded at run time (dlopen), so the
 * library binds to the libnccl the process already holds.  Bootstrap: rank 0 calls cnv_comm_unique_id, the caller
 * broadcasts the 128 bytes (torch.distributed / MPI / a file), every rank calls cnv_comm_create.
 * cnv_poisson_enqueue_dist: per pass one NCCL group on `stream` (2T halo rows with both neighbours + an all-gather of
 * the T per-sweep norms) followed by the device-side stop decision; no host synchronisation. */
typedef struct cnv_comm cnv_comm;
int cnv_comm_unique_id(unsigned char *out128);                           /* 0 ok, else NCCL unavailable */
cnv_comm *cnv_comm_create(int rank, int world, const unsigned char *id128); /* NULL on failure */
void cnv_comm_destroy(cnv_comm *c);
void cnv_poisson_attach_comm(cnv_poisson *p, cnv_comm *c);
void cnv_poisson_enqueue_dist(cnv_poisson *p, int npasses, void *stream);
void cnv_poisson_exchange_halos(cnv_poisson *p, double *field_dev, int depth, void *stream);

/* Peer-memory path (CUDA IPC over NVLink; one process per GPU on one node).  Once set up, cnv_poisson_enqueue() needs
 * no collective at all: the pass kernel stores its boundary rows straight into the neighbours' halo rows, counts its
 * pushes in the neighbours' mailboxes, publishes its per-sweep norms in every rank's mailbox, and every CTA of the next
 * pass derives the stop decision itself.  Setup: every rank exports 256 bytes (IPC handles of iterate buffers 0 and 1,
 * of its mailbox, 64 reserved bytes) and its push counts; the caller all-gathers them (world x 256 bytes;
 * world x 4 ints: own_lo, own_hi, push_low, push_high) and every rank imports.  All spin-waits are bounded
 * (CNV_PEER_TIMEOUT_MS, default 300000; message + exit(1) on every rank once one of them gives up).
 */
void cnv_poisson_peer_export(cnv_poisson *p, unsigned char *out256);
void cnv_poisson_peer_push_counts(cnv_poisson *p, int rank, int world, long long *low, long long *high);
int cnv_poisson_peer_import(cnv_poisson *p, int rank, int world, const unsigned char *handles, const int *layout);
void cnv_poisson_peer_disable(cnv_poisson *p);
/* Tear-down, in this order on EVERY rank: cnv_poisson_peer_quiesce(stream); synchronise the device; barrier of all ranks;
 * cnv_poisson_peer_close() (unmaps the neighbours' buffers); only then cnv_poisson_destroy.  A rank that frees its buffers
 * while a neighbour's trailing passes still write flags or halo rows into them faults that neighbour. */
void cnv_poisson_peer_quiesce(cnv_poisson *p, void *stream);
void cnv_poisson_peer_close(cnv_poisson *p);
int cnv_poisson_peer_enabled(cnv_poisson *p);
/* diagnostics of the peer path: per-CTA globaltimer stamps (start, state known, halos landed, stream done, push done,
 * exit) of the first `passes` passes after a reset; returns the CTAs per pass.  Read passes*ctas*6 values back. */
int cnv_poisson_peer_trace(cnv_poisson *p, int passes);
void cnv_poisson_peer_trace_read(cnv_poisson *p, unsigned long long *out, long long n);

/* ---- device-resident time stepping: the loop body of src/main.c:283-395 ---------------------- */
typedef struct cnv_sim cnv_sim;
cnv_sim *cnv_sim_create(const Config *cfg, int T);
/* Multi-GPU: this process' slab (rank of world) of the row-decomposed grid; local arrays carry 2T halo rows on
 * interior edges.  The caller orchestrates a step as cnv_sim_phase() calls interleaved with halo exchanges
 * (fluid_dynamics1_b200/parallel.py: SlabSimulation); cnv_sim_step() is for world == 1 only. */
cnv_sim *cnv_sim_create_slab(const Config *cfg, int T, int rank, int world);
void cnv_sim_layout(const cnv_sim *s, int *out); /* grow0, nloc, own_lo, own_hi, ld, ncols, T, global rows */
cnv_poisson *cnv_sim_poisson(cnv_sim *s);        /* the simulation's solver object (not owned by the caller) */
double *cnv_sim_field_ptr(cnv_sim *s, int which); /* device: 0 u, 1 v, 2 w, 3 continuity (max, min) */
void cnv_sim_set_psi_buf(cnv_sim *s, int which);
/* phase 0 BCs + wall vorticity | 1 derivatives + Euler + rhs | 2 velocities from psi | 3 continuity max/min */
void cnv_sim_phase(cnv_sim *s, int phase, void *stream);
void cnv_sim_destroy(cnv_sim *s);
/* Advances nsteps steps.  k/e/cont_max/cont_min: NULL or arrays of nsteps (per-step Poisson log values
 * and continuity diagnostic).  Returns 0, or (index of the step whose Poisson solve hit itmax) + 1. */
int cnv_sim_step(cnv_sim *s, int nsteps, int *k, double *e, double *cont_max, double *cont_min);
/* Slab simulations (world > 1; attach the NCCL communicator to cnv_sim_poisson() first): whole time steps inside the library
 * -- the stencil phases with their 3-row halo exchanges, the distributed Poisson solve (peer-memory path if set up, the NCCL
 * group per pass otherwise), the continuity max / min all-reduce.  Every rank calls it with the same arguments and receives
 * the same values.  Returns 0, (index of the step whose Poisson solve hit itmax) + 1, or -1 without a communicator. */
int cnv_sim_step_slab(cnv_sim *s, int nsteps, int *k, double *e, double *cont_max, double *cont_min);
/* whole fields on rank 0 (host arrays of nx*ny there, ignored on the other ranks; a NULL on rank 0 drops that field): every
 * rank sends the owned rows of all four fields over the communicator */
int cnv_sim_gather_fields_slab(cnv_sim *s, double *psi, double *w, double *u, double *v);
int cnv_sim_get_fields(cnv_sim *s, double *psi, double *w, double *u, double *v); /* host, nx*ny each, NULL to skip */
int cnv_sim_set_fields(cnv_sim *s, const double *psi, const double *w, const double *u, const double *v);
void cnv_sim_set_diagnostics(cnv_sim *s, int continuity_on);
/* measurement hook: the stencil phase only (BCs + wall vorticity, derivatives + Euler, velocity recovery,
 * continuity diagnostic), `reps` times, asynchronously on `stream`; no Poisson solve */
void cnv_sim_stencil_phase(cnv_sim *s, int reps, void *stream);
/* cumulative counters: [0] Poisson sweeps, [1] Poisson passes, [2] steps */
void cnv_sim_counters(cnv_sim *s, long long *out);
/* p = poisson(-f), f as in cnv_pressure_rhs_host from the simulation's current u, v; the configured Poisson variant
 * (src/main.c:351/355), zero ring, zero guess.  itmax <= 0 / tol <= 0: the configuration's values.  p_host (nx*ny,
 * may be NULL), *k, *e as logged by poisson_SOR_log.  Returns 0, 1 = itmax reached, -1 = slab simulation. */
int cnv_sim_pressure(cnv_sim *s, int itmax, double tol, double *p_host, int *k, double *e);

/* ---- whole driver: restates src/main.c:27-481 on top of the device-resident path ---------------
 * (same config file, same stdout/log lines, same step order; VTK through printvtk-compatible writer
 * unless CNV_NO_VTK=1).  Returns the process exit code.
 * CNV_GPUS=N (2..8): the grid is slab-decomposed over N GPUs, one process per GPU -- cnv_main forks N-1 children BEFORE its
 * first CUDA call (so call it from a process that has not initialised CUDA: the cnavier_b200 executable does), exchanges the
 * NCCL id and the CUDA IPC handles of the peer-memory path over socket pairs, and rank 0 (the calling process) writes the
 * same log lines and VTK files as a one-GPU run; fields and Poisson iteration counts are bit-identical to it. */
int cnv_main(int argc, char **argv);
/* CPU-only self-test of the fork + socket-pair bootstrap cnv_main uses for CNV_GPUS=N (all-gather, broadcast, barrier,
 * agreement on a failure): 0 = every one of `world` processes saw the expected bytes. */
int cnv_boot_selftest(int world);
/* The driver's VTK writer on its own: replaces printvtk (src/utils.c:38-100) -- ASCII STRUCTURED_POINTS, "%.6lf", file
 * <output_dir>/<title>-1-<count>.vtk opened in append mode, ONE counter across all fields and calls of the process.
 * `values` is row-major m x n (the reference's A.M[i][j]).  Returns the counter value used.  Host only. */
int cnv_vtk_write(const double *values, int m, int n, const char *title, const char *output_dir);

/* configuration system (src/config.c) under cnv_ names for callers that do not want the
 * reference-named symbols of the drop-in library */
void cnv_config_default(Config *out);
void cnv_config_from_file(const char *filename, Config *out);
void cnv_config_print(const Config *cfg);

#ifdef __cplusplus
}
#endif
#endif /* CNAVIER_B200_H */
