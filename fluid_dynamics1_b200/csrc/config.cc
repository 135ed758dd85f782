// config.cc -- the reference's config.txt system (src/config.c, include/config.h), kept so the same
// files drive the GPU path.  Behaviour restated: defaults (src/config.c:47-87); `key = value` parser
// where a line is skipped only if its FIRST character is '\n', '#' or ';', the key is everything before
// the first '=', the value is the next whitespace-delimited token (so trailing "# comments" are
// dropped), ints go through atoi and doubles through atof, unknown keys warn, an unreadable file warns
// and yields the defaults (:106-214); the report printers (:230-360).  Table driven.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/cnavier_b200.h"

namespace {

enum Kind { kDouble, kInt };
struct Key {
    const char *name;
    Kind kind;
    size_t off;
};
#define CNV_KEY(field, kind) {#field, kind, offsetof(Config, field)}
const Key kKeys[] = {
    CNV_KEY(Re, kDouble), CNV_KEY(Lx, kInt), CNV_KEY(Ly, kInt), CNV_KEY(nx, kInt), CNV_KEY(ny, kInt),
    CNV_KEY(dt, kDouble), CNV_KEY(tf, kDouble), CNV_KEY(max_co, kDouble), CNV_KEY(order, kInt),
    CNV_KEY(poisson_max_it, kInt), CNV_KEY(poisson_tol, kDouble), CNV_KEY(output_interval, kInt),
    CNV_KEY(poisson_type, kInt), CNV_KEY(openmp_enabled, kInt), CNV_KEY(ui, kDouble), CNV_KEY(vi, kDouble),
    CNV_KEY(u1, kDouble), CNV_KEY(u2, kDouble), CNV_KEY(u3, kDouble), CNV_KEY(u4, kDouble), CNV_KEY(v1, kDouble),
    CNV_KEY(v2, kDouble), CNV_KEY(v3, kDouble), CNV_KEY(v4, kDouble),
};
#undef CNV_KEY

char *trim(char *s)
{
    while (*s == ' ' || *s == '\t') s++;
    char *end = s + std::strlen(s) - 1;
    while (end > s && (*end == ' ' || *end == '\t' || *end == '\n' || *end == '\r')) *end-- = '\0';
    return s;
}

}  // namespace

extern "C" {

void cnv_config_default(Config *c)
{
    std::memset(c, 0, sizeof *c);
    c->Re = 1000.0; c->Lx = 1; c->Ly = 1;
    c->nx = 64; c->ny = 64; c->dt = 0.005; c->tf = 20.0; c->max_co = 1.0; c->order = 6;
    c->poisson_max_it = 10000; c->poisson_tol = 1E-3; c->output_interval = 10; c->poisson_type = 2;
    c->openmp_enabled = 1;  // the GPU path always runs the parallel (red-black) ordering
    c->u4 = 1.0;            // lid-driven cavity: only the top wall moves
}

void cnv_config_from_file(const char *filename, Config *c)
{
    cnv_config_default(c);
    FILE *f = std::fopen(filename, "r");
    if (!f) {
        std::printf("Error: Could not open configuration file '%s'\n", filename);
        std::printf("Using default configuration.\n");
        return;
    }
    std::printf("Loading configuration from: %s\n", filename);
    char line[256], key[64], value[64];
    while (std::fgets(line, sizeof line, f)) {
        if (line[0] == '\n' || line[0] == '#' || line[0] == ';') continue;
        if (std::sscanf(line, "%63[^=]=%63s", key, value) != 2) continue;
        char *k = trim(key), *v = trim(value);
        bool known = false;
        for (const Key &e : kKeys) {
            if (std::strcmp(k, e.name) != 0) continue;
            char *dst = reinterpret_cast<char *>(c) + e.off;
            if (e.kind == kInt) *reinterpret_cast<int *>(dst) = std::atoi(v);
            else *reinterpret_cast<double *>(dst) = std::atof(v);
            known = true;
            break;
        }
        if (!known) std::printf("Warning: Unknown configuration parameter '%s'\n", k);
    }
    std::fclose(f);
    std::printf("Configuration loaded successfully.\n");
}

void cnv_config_print(const Config *c)
{
    const double dx = (double)c->Lx / c->nx, dy = (double)c->Ly / c->ny;
    std::printf("\n=== Simulation Configuration ===\n");
    std::printf("Physical Parameters:\n");
    std::printf("  Reynolds number (Re): %.2f\n  Domain length (Lx): %d\n  Domain width (Ly): %d\n", c->Re, c->Lx, c->Ly);
    std::printf("\nNumerical Parameters:\n");
    std::printf("  Grid points x (nx): %d\n  Grid points y (ny): %d\n", c->nx, c->ny);
    std::printf("  Time step (dt): %.6f\n  Final time (tf): %.2f\n  Max Courant number: %.2f\n", c->dt, c->tf, c->max_co);
    std::printf("  Finite difference order: %d\n  Poisson max iterations: %d\n  Poisson tolerance: %.2E\n", c->order,
                c->poisson_max_it, c->poisson_tol);
    std::printf("  Output interval: %d\n  Poisson solver type: %d\n", c->output_interval, c->poisson_type);
    std::printf("\nPerformance Parameters:\n");
    std::printf("  OpenMP enabled (config): %s\n", c->openmp_enabled ? "Yes" : "No");
    std::printf("  OpenMP compiled support: No (CUDA sm_100a path; the setting is ignored)\n");
    std::printf("\nBoundary Conditions:\n");
    std::printf("  Internal u field (ui): %.2f\n  Internal v field (vi): %.2f\n", c->ui, c->vi);
    std::printf("  u boundaries (u1,u2,u3,u4): %.2f, %.2f, %.2f, %.2f\n", c->u1, c->u2, c->u3, c->u4);
    std::printf("  v boundaries (v1,v2,v3,v4): %.2f, %.2f, %.2f, %.2f\n", c->v1, c->v2, c->v3, c->v4);
    std::printf("\nDerived Parameters:\n");
    std::printf("  Grid spacing dx: %.6f\n  Grid spacing dy: %.6f\n", dx, dy);
    std::printf("  SOR parameter (beta): %.6f\n", cnv_sor_beta(c->nx, c->ny));
    std::printf("  Maximum iterations: %d\n", (int)((c->tf / c->dt) - 1));
    std::printf("================================\n\n");
}

}  // extern "C"
