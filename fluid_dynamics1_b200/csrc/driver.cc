// driver.cc -- cnv_main(): the reference program flow (src/main.c:27-481) on the device-resident path.
//
// Same command line (`[config_file] [output_folder]`, --help, --help-config), same config system, same
// step order, same log-file lines (they are the reference's golden-vector format: the Poisson line
// src/poisson.c:275, the per-step lines src/main.c:403-414, the header :262-271, the summary :467-473),
// same exit(1) on Poisson non-convergence, same VTK files (src/utils.c:38-100: ASCII STRUCTURED_POINTS,
// "%.6lf", one global file counter, append mode).  Fields never leave HBM between steps; VTK output
// copies them out only at output steps.  GPU-only knobs come from the environment so config files
// stay reference-compatible:  CNV_NO_VTK=1 suppresses VTK files (large grids),
// CNV_POISSON_T=<1|2|4|8> temporal block depth.
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/cnavier_b200.h"

namespace {

void log_line(FILE *log, const char *fmt, ...)
{
    if (!log) return;
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(log, fmt, ap);
    va_end(ap);
    std::fflush(log);
}

void make_dir(const std::string &dir)
{
    struct stat st;
    if (stat(dir.c_str(), &st) == -1) {
        std::string cmd = "mkdir -p \"" + dir + "\"";
        if (std::system(cmd.c_str()) != 0) std::printf("Warning: could not create %s\n", dir.c_str());
        std::printf("Created output directory: %s\n", dir.c_str());
    }
}

int g_vtk_count = 0;  // one counter across all fields, like the static in src/utils.c:42

void write_vtk(const std::vector<double> &A, int m, int n, const char *title, const std::string &dir)
{
    make_dir(dir);
    char name[512];
    std::snprintf(name, sizeof name, "%s/%s-1-%d.vtk", dir.c_str(), title, g_vtk_count);
    FILE *pf = std::fopen(name, "a");
    if (!pf) {
        std::printf("\nError while opening file: %s\n", name);
        std::exit(1);
    }
    std::printf("%s\n", name);
    std::fprintf(pf, "# vtk DataFile Version 2.0\ntest\nASCII\nDATASET STRUCTURED_POINTS\n");
    std::fprintf(pf, "DIMENSIONS %d %d 1\nORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA %d\n", m, n, m * n);
    std::fprintf(pf, "SCALARS values float\nLOOKUP_TABLE default");
    for (int i = 0; i < m; i++) {
        std::fprintf(pf, "\n");
        for (int j = 0; j < n; j++) std::fprintf(pf, j == 0 ? "%.6lf" : " %.6lf", A[(size_t)i * n + j]);
    }
    std::fclose(pf);
    g_vtk_count++;
}

void usage(const char *prog)
{
    std::printf("Usage: %s [config_file] [output_folder]\n\nOptions:\n", prog);
    std::printf("  config_file    Path to configuration file (optional)\n");
    std::printf("                 If not provided, default values will be used\n");
    std::printf("  output_folder  Name of output subfolder (optional)\n");
    std::printf("                 Files will be saved to ./output/[output_folder]/\n");
    std::printf("                 If not provided, files will be saved to ./output/\n");
    std::printf("\nConfiguration file format:\n  # Comments start with # or ;\n  parameter_name = value\n");
    std::printf("\nFor a complete list of parameters, run with --help-config\n");
}

void help_config()
{
    std::printf("Complete list of configuration parameters:\n\n");
    std::printf("Physical Parameters:\n  Re, Lx, Ly\n");
    std::printf("\nNumerical Parameters:\n  nx, ny, dt, tf, max_co, order (2|4|6), poisson_max_it, poisson_tol,\n");
    std::printf("  output_interval, poisson_type (1=no relaxation, 2=SOR)\n");
    std::printf("\nPerformance Parameters:\n  openmp_enabled (parsed, ignored by the GPU path)\n");
    std::printf("\nBoundary Conditions:\n  ui, vi, u1, u2, u3, u4, v1, v2, v3, v4\n");
}

// ---- multi-GPU run: one process per GPU (CNV_GPUS=N), forked by cnv_main before its first CUDA call ------------------
// The ranks only need three host-side collectives to set themselves up (NCCL unique id, CUDA IPC handles of the peer-memory
// path) and to stay in step at the end: a star of socket pairs between rank 0 (the process the user started) and its children.
struct Boot {
    int rank = 0, world = 1;
    std::vector<int> fd;  // rank 0: fd[r] = socket to rank r (fd[0] unused); rank r > 0: fd[0] = socket to rank 0
    static void xfer(int f, void *buf, size_t n, bool wr)
    {
        char *p = static_cast<char *>(buf);
        while (n > 0) {
            const ssize_t k = wr ? ::write(f, p, n) : ::read(f, p, n);
            if (k <= 0) {
                std::printf("** Error: a rank of the multi-GPU run stopped responding **\n");
                std::fflush(stdout);
                std::_Exit(1);
            }
            p += k;
            n -= (size_t)k;
        }
    }
    // every rank contributes `bytes` bytes; everybody receives all of them in rank order
    void allgather(const void *mine, size_t bytes, void *all)
    {
        char *out = static_cast<char *>(all);
        if (rank == 0) {
            std::memcpy(out, mine, bytes);
            for (int r = 1; r < world; r++) xfer(fd[r], out + (size_t)r * bytes, bytes, false);
            for (int r = 1; r < world; r++) xfer(fd[r], out, bytes * world, true);
        } else {
            xfer(fd[0], const_cast<void *>(mine), bytes, true);
            xfer(fd[0], out, bytes * world, false);
        }
    }
    void bcast(void *buf, size_t bytes)  // from rank 0
    {
        if (rank == 0) { for (int r = 1; r < world; r++) xfer(fd[r], buf, bytes, true); }
        else xfer(fd[0], buf, bytes, false);
    }
    void barrier()
    {
        char c = 0;
        std::vector<char> all(world);
        allgather(&c, 1, all.data());
    }
    bool all_ok(bool ok)
    {
        char c = ok ? 1 : 0;
        std::vector<char> all(world);
        allgather(&c, 1, all.data());
        for (char x : all) ok = ok && x;
        return ok;
    }
};

// fork world-1 children (call BEFORE the first CUDA call of the process); returns this process' rank
Boot boot_fork(int world)
{
    Boot b;
    b.world = world;
    b.fd.assign(world, -1);
    std::fflush(stdout);
    for (int r = 1; r < world; r++) {
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { std::printf("** Error: socketpair failed **\n"); std::exit(1); }
        const pid_t pid = fork();
        if (pid < 0) { std::printf("** Error: fork failed **\n"); std::exit(1); }
        if (pid == 0) {  // child = rank r: keeps its end, drops everything inherited from the parent
            for (int q = 1; q < r; q++) ::close(b.fd[q]);
            ::close(sv[0]);
            b.rank = r;
            b.fd.assign(world, -1);
            b.fd[0] = sv[1];
            if (!std::getenv("CNV_RANK_STDOUT")) (void)!std::freopen("/dev/null", "w", stdout);  // rank 0 does the talking
            return b;
        }
        ::close(sv[1]);
        b.fd[r] = sv[0];
    }
    return b;
}

double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

}  // namespace

// The driver's VTK writer on its own (printvtk semantics, src/utils.c:38-100): row-major m x n values, appended to
// <dir>/<title>-1-<count>.vtk, one counter across all calls of the process.  Returns the counter value the file was written
// with.  Host code only (no GPU involved).
extern "C" int cnv_vtk_write(const double *values, int m, int n, const char *title, const char *output_dir)
{
    if (!values || m < 1 || n < 1) {
        std::printf("\n** Error: Invalid parameter **\n");  // src/utils.c:46-55
        std::exit(1);
    }
    const int used = g_vtk_count;
    write_vtk(std::vector<double>(values, values + (size_t)m * n), m, n, title, output_dir);
    return used;
}

// Self-test of the bootstrap layer behind CNV_GPUS=N (no CUDA): forks world-1 children and runs the exchanges cnv_main uses --
// all-gather of 256-byte records (the IPC handle exchange), broadcast (the NCCL id), barrier, all_ok with one dissenting rank.
// Returns 0 in the parent if every rank saw what it should; the children exit with their own verdict and are waited for.
extern "C" int cnv_boot_selftest(int world)
{
    if (world < 1 || world > 64) return 2;
    Boot b = boot_fork(world);
    bool ok = true;
    std::vector<unsigned char> mine(256), all(256 * (size_t)world);
    for (int i = 0; i < 256; i++) mine[i] = (unsigned char)(b.rank * 7 + i);
    b.allgather(mine.data(), 256, all.data());
    for (int r = 0; r < world; r++)
        for (int i = 0; i < 256; i++) ok = ok && all[(size_t)r * 256 + i] == (unsigned char)(r * 7 + i);
    unsigned char id[128];
    for (int i = 0; i < 128; i++) id[i] = b.rank == 0 ? (unsigned char)(i ^ 0x5a) : 0;
    b.bcast(id, sizeof id);
    for (int i = 0; i < 128; i++) ok = ok && id[i] == (unsigned char)(i ^ 0x5a);
    b.barrier();
    ok = ok && b.all_ok(true);
    const bool verdict = b.all_ok(b.rank != world - 1);  // the last rank dissents: everybody must see a failure
    ok = ok && !verdict;
    ok = b.all_ok(ok);
    if (b.rank != 0) std::_Exit(ok ? 0 : 1);
    int rc = ok ? 0 : 1;
    for (int r = 1; r < world; r++) {
        int st = 0;
        if (::wait(&st) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
    }
    for (int r = 1; r < world; r++) ::close(b.fd[r]);
    return rc;
}

extern "C" int cnv_main(int argc, char **argv)
{
    Config cfg;
    std::string output_dir = "./output";
    const double start_time = now_s();
    if (argc == 1) {
        std::printf("No configuration file specified. Using default values.\n");
        cnv_config_default(&cfg);
    } else if (argc == 2 || argc == 3) {
        if (argc == 2 && (!std::strcmp(argv[1], "--help") || !std::strcmp(argv[1], "-h"))) { usage(argv[0]); return 0; }
        if (argc == 2 && !std::strcmp(argv[1], "--help-config")) { help_config(); return 0; }
        cnv_config_from_file(argv[1], &cfg);
        if (argc == 3) {
            output_dir = std::string("./output/") + argv[2];
            std::printf("Using output directory: %s\n", output_dir.c_str());
        }
    } else {
        std::printf("Error: Too many arguments.\n");
        usage(argv[0]);
        return 1;
    }
    cnv_config_print(&cfg);
    // CNV_GPUS=N (N > 1): slab-decomposed run, one process per GPU.  The children are forked here, before this process makes
    // its first CUDA call; rank 0 (this process) keeps the terminal, the log file and the VTK output.
    const int ngpu = std::getenv("CNV_GPUS") ? std::atoi(std::getenv("CNV_GPUS")) : 1;
    Boot boot;
    if (ngpu > 1) boot = boot_fork(ngpu);
    const int rank = boot.rank, world = boot.world;
    if (world > 1) {
        const bool enough = cnv_device_count() >= world;
        if (!boot.all_ok(enough)) {
            if (rank == 0) std::printf("** Error: CNV_GPUS=%d but only %d CUDA device(s) are visible **\n", world, cnv_device_count());
            if (rank != 0) std::_Exit(1);
            return 1;
        }
        cnv_set_device(rank);
    }
    std::printf("\n=== Execution Status ===\nBackend: CUDA (sm_100a), %d device(s) visible, %d in use\n========================\n\n",
                cnv_device_count(), world);

    const double beta = cnv_sor_beta(cfg.nx, cfg.ny);
    std::printf("Poisson SOR parameter: %lf\n", beta);
    const double dx = (double)cfg.Lx / cfg.nx, dy = (double)cfg.Ly / cfg.ny;
    const int it_max = (int)((cfg.tf / cfg.dt) - 1);
    // Courant check exactly as the reference does it: with u1 (src/main.c:165-174)
    const double r1 = cfg.u1 * cfg.dt / dx, r2 = cfg.u1 * cfg.dt / dy;
    if (r1 > cfg.max_co || r2 > cfg.max_co) {
        std::printf("Unstable Solution!\nr1: %lf\nr2: %lf\n", r1, r2);
        return 1;
    }
    if (cfg.poisson_type != 1 && cfg.poisson_type != 2) {
        std::printf("Error - invalid option for Poisson solver\n");
        return 1;
    }

    const char *tenv = std::getenv("CNV_POISSON_T");
    cnv_sim *sim = world > 1 ? cnv_sim_create_slab(&cfg, tenv ? std::atoi(tenv) : 0, rank, world) : cnv_sim_create(&cfg, tenv ? std::atoi(tenv) : 0);
    cnv_comm *comm = nullptr;
    if (world > 1) {
        // NCCL communicator of the slab halo exchanges: rank 0 creates the unique id, everybody joins
        unsigned char id[128] = {0};
        int ok = rank == 0 ? cnv_comm_unique_id(id) == 0 : 1;
        boot.bcast(id, sizeof id);
        if (ok) comm = cnv_comm_create(rank, world, id);
        if (!boot.all_ok(comm != nullptr)) {
            if (rank == 0) std::printf("** Error: could not create the NCCL communicator of the %d ranks **\n", world);
            if (rank != 0) std::_Exit(1);
            return 1;
        }
        cnv_poisson *ps = cnv_sim_poisson(sim);
        cnv_poisson_attach_comm(ps, comm);
        // peer-memory path of the Poisson solve: exchange the CUDA IPC handles and push counts
        const char *be = std::getenv("CNV_DIST_BACKEND");
        if (!be || std::strcmp(be, "peer") == 0) {
            struct Rec { unsigned char h[256]; int layout[4]; } mine, *all = new Rec[world];
            int lay[8];
            cnv_sim_layout(sim, lay);
            long long lo = 0, hi = 0;
            cnv_poisson_peer_export(ps, mine.h);
            cnv_poisson_peer_push_counts(ps, rank, world, &lo, &hi);
            mine.layout[0] = lay[2]; mine.layout[1] = lay[3]; mine.layout[2] = (int)lo; mine.layout[3] = (int)hi;
            boot.allgather(&mine, sizeof mine, all);
            std::vector<unsigned char> handles((size_t)256 * world);
            std::vector<int> layout((size_t)4 * world);
            for (int r = 0; r < world; r++) {
                std::memcpy(&handles[(size_t)256 * r], all[r].h, 256);
                std::memcpy(&layout[(size_t)4 * r], all[r].layout, sizeof(int) * 4);
            }
            delete[] all;
            const bool imported = cnv_poisson_peer_import(ps, rank, world, handles.data(), layout.data()) == 0;
            if (!boot.all_ok(imported)) cnv_poisson_peer_disable(ps);  // all ranks or none: the NCCL group per pass instead
        }
        std::printf("Slab decomposition: %d ranks, Poisson exchange: %s\n", world, cnv_poisson_peer_enabled(ps) ? "peer memory (NVLink)" : "NCCL");
    }
    const bool no_vtk = std::getenv("CNV_NO_VTK") && std::atoi(std::getenv("CNV_NO_VTK"));

    // log file ./output/logs/<run>.txt (src/main.c:229-273)
    std::string run_name = "default";
    const std::string prefix = "./output/";
    if (output_dir.compare(0, prefix.size(), prefix) == 0 && output_dir.size() > prefix.size())
        run_name = output_dir.substr(prefix.size());
    const std::string log_filename = "./output/logs/" + run_name + ".txt";
    if (rank == 0 && std::system("mkdir -p output/logs") != 0) std::printf("Warning: could not create output/logs\n");
    FILE *log = rank == 0 ? std::fopen(log_filename.c_str(), "w") : nullptr;
    if (rank != 0) {
        // (the other ranks compute their slabs; rank 0 logs and writes the output files)
    } else if (!log) {
        std::printf("Warning: Could not create log file %s. Logging to console instead.\n", log_filename.c_str());
    } else {
        std::printf("Logging simulation progress to: %s\n", log_filename.c_str());
        std::fprintf(log, "=== Fluid Dynamics Simulation Log ===\nRun name: %s\nOutput directory: %s\n", run_name.c_str(),
                     output_dir.c_str());
        std::fprintf(log, "Reynolds number: %.2f\nGrid size: %dx%d\nTime step: %.6f\nFinal time: %.2f\n", cfg.Re, cfg.nx, cfg.ny,
                     cfg.dt, cfg.tf);
        std::fprintf(log, "Max iterations: %d\nOpenMP enabled: %s\n", it_max + 1, cfg.openmp_enabled ? "Yes" : "No");
        std::fprintf(log, "=====================================\n\n");
        std::fflush(log);
    }

    const double sim_start = now_s();
    std::vector<double> psi, w, u, v;
    std::vector<int> k_hist;       // metrics file: logged k and residual of every step
    std::vector<double> e_hist;
    int rc = 0;
    for (int t = 0; t <= it_max; t++) {
        const double t0 = now_s();
        int k = 0;
        double e = 0, cmax = 0, cmin = 0;
        const int failed = world > 1 ? cnv_sim_step_slab(sim, 1, &k, &e, &cmax, &cmin) : cnv_sim_step(sim, 1, &k, &e, &cmax, &cmin);
        if (failed) {
            log_line(log, "Error: maximum number of iterations achieved for Poisson equation.\n");  // src/poisson.c:280
            rc = 1;
            break;
        }
        log_line(log, "Poisson equation solved with %d iterations - root-sum-of-squares error: %E\n", k, e);
        k_hist.push_back(k);
        e_hist.push_back(e);
        const double t1 = now_s();
        const double elapsed = t1 - sim_start;
        log_line(log, "Iteration: %d | ", t);
        log_line(log, "Time: %lf | ", (double)t * cfg.dt);
        log_line(log, "Progress: %.2lf%% | ", (double)100 * t / it_max);
        log_line(log, "Iter time: %.3f s\n", t1 - t0);
        log_line(log, "Continuity max: %E | ", cmax);
        log_line(log, "Continuity min: %E | ", cmin);
        log_line(log, "Elapsed: %.1f s | ", elapsed);
        if (t > 0) log_line(log, "Est. remaining: %.1f s\n", elapsed / (t + 1) * (it_max - t));
        else log_line(log, "Est. remaining: -- s\n");
        if (!no_vtk && cfg.output_interval != 0 && t % cfg.output_interval == 0) {
            const size_t n = (size_t)cfg.nx * cfg.ny;
            if (rank == 0) { psi.resize(n); w.resize(n); u.resize(n); v.resize(n); }
            if (world > 1) {
                cnv_sim_gather_fields_slab(sim, psi.data(), w.data(), u.data(), v.data());  // every rank sends its owned rows to rank 0
                if (rank != 0) continue;
            } else {
                cnv_sim_get_fields(sim, psi.data(), w.data(), u.data(), v.data());
            }
            write_vtk(psi, cfg.nx, cfg.ny, "stream-function", output_dir);
            write_vtk(w, cfg.nx, cfg.ny, "vorticity", output_dir);
            write_vtk(u, cfg.nx, cfg.ny, "x-velocity", output_dir);
            write_vtk(v, cfg.nx, cfg.ny, "y-velocity", output_dir);
        }
    }
    long long counters[3];
    cnv_sim_counters(sim, counters);
    if (world > 1) {
        // tear-down in the order the peer path needs: no rank frees a buffer a neighbour has mapped or may still write into
        cnv_poisson *ps = cnv_sim_poisson(sim);
        cnv_poisson_peer_quiesce(ps, nullptr);
        cnv_device_synchronize();
        boot.barrier();
        cnv_poisson_peer_close(ps);
        cnv_device_synchronize();
        boot.barrier();
    }
    cnv_sim_destroy(sim);
    if (comm) cnv_comm_destroy(comm);
    if (world > 1 && rank != 0) std::_Exit(rc);  // children are done; rank 0 reports
    if (world > 1) {
        for (int r = 1; r < world; r++) {
            int st = 0;
            if (wait(&st) > 0 && (!WIFEXITED(st) || WEXITSTATUS(st) != rc) && rc == 0) rc = 1;
        }
    }
    if (rc == 0) std::printf("Simulation complete!\n");
    const double end = now_s();
    log_line(log, "\n=== Timing Summary ===\n");
    log_line(log, "Setup time: %.4f seconds\n", sim_start - start_time);
    log_line(log, "Simulation time: %.2f seconds\n", end - sim_start);
    log_line(log, "Total program time: %.2f seconds\n", end - start_time);
    log_line(log, "Average iteration time: %.4f seconds\n", (end - sim_start) / (it_max + 1));
    log_line(log, "Iterations completed: %d\n", (int)counters[2]);
    log_line(log, "Poisson sweeps: %lld (%lld passes)\n", counters[0], counters[1]);
    log_line(log, "======================\n");
    if (log) {
        std::fclose(log);
        std::printf("Log saved to: %s\n", log_filename.c_str());
    }
    // CNV_METRICS_JSON=<path>: machine-readable run metrics next to the reference-format log (throughput against the
    // algorithmic traffic of SURVEY.md section 8d: 24 B per interior cell per sweep + 72 B per cell per step)
    if (const char *mj = std::getenv("CNV_METRICS_JSON")) {
        if (FILE *f = std::fopen(mj, "w")) {
            const double secs = end - sim_start, cells = (double)cfg.nx * cfg.ny, icells = (double)(cfg.nx - 2) * (cfg.ny - 2);
            const double steps = (double)counters[2], sweeps = (double)counters[0];
            std::fprintf(f, "{\"grid\": [%d, %d], \"steps\": %lld, \"poisson_sweeps\": %lld, \"poisson_passes\": %lld, "
                            "\"simulation_seconds\": %.6f, \"timestep_cell_updates_per_s\": %.6e, \"poisson_cell_updates_per_s\": %.6e, "
                            "\"poisson_sweeps_per_s\": %.6e, \"algorithmic_gb_per_s\": %.3f, \"status\": %d,\n \"poisson_k\": [",
                         cfg.nx, cfg.ny, counters[2], counters[0], counters[1], secs, cells * steps / secs, icells * sweeps / secs,
                         sweeps / secs, (24.0 * icells * sweeps + 72.0 * cells * steps) / secs / 1e9, rc);
            for (size_t i = 0; i < k_hist.size(); i++) std::fprintf(f, "%s%d", i ? ", " : "", k_hist[i]);
            std::fprintf(f, "],\n \"poisson_residual\": [");
            for (size_t i = 0; i < e_hist.size(); i++) std::fprintf(f, "%s%.6E", i ? ", " : "", e_hist[i]);
            std::fprintf(f, "]}\n");
            std::fclose(f);
        }
    }
    return rc;
}
