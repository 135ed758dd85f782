/*
 * cnavier_dropin.h -- the reference's own C signatures for the hot path, served by the B200 library.
 *
 * libcnavier_dropin.so exports exactly the symbols below, with the argument meaning, ownership and
 * error behaviour of the reference (message + exit(1); no return codes), so that the reference's
 * main.c links against it unchanged.  Each declaration cites the reference
 * interface it replaces.  The per-file headers include/{linearalg,finitediff,fluiddyn,poisson,config}.h
 * only forward to this file, so `#include "poisson.h"` etc. in main.c keep working.
 *
 * Every mtrx returned by these functions is caller-owned row-pointer storage that the caller
 * releases with freem(), exactly as with the reference (src/linearalg.c:84-99).
 */
#ifndef CNAVIER_DROPIN_H
#define CNAVIER_DROPIN_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- field container: reference include/linearalg.h:6-17 ------------------------------------ */
typedef struct matrix {
    double **M; /* m row pointers, each row its own allocation (src/linearalg.c:72-74) */
    int m;      /* rows    (first index, "i") */
    int n;      /* columns (second index, "j") */
} mtrx;

typedef struct vector {
    double *v;
    int n;
} vec;

/* reference include/linearalg.h:19-41; host-side helpers the driver uses around the hot path
 * (src/linearalg.c:26-99, :353-573).  mtrxmul/kronecker/reshape keep the dense semantics so an
 * unmodified main.c still runs at small sizes; the GPU driver never calls them. */
void zerosm(mtrx A);
double **allocm(int m, int n);
double **freem(mtrx A);
mtrx initm(int m, int n);
mtrx eye(int n);
mtrx reshape(mtrx A, int m, int n);
mtrx kronecker(mtrx A, mtrx B);
mtrx mtrxmul(mtrx A, mtrx B);
void invsig(mtrx A);
double maxel(mtrx A);
double minel(mtrx A);
void mtrxcpy(mtrx A, mtrx B);
void set_openmp_config(int enabled); /* accepted and ignored: parallelism is the GPU's */

/* ---- finite differences: reference include/finitediff.h:8-9 (src/finitediff.c:51, :178) ------- */
mtrx Diff1(int n, int o, double dx);
mtrx Diff2(int n, int o, double dx);

/* ---- fluid dynamics: reference include/fluiddyn.h:8-14 (src/fluiddyn.c:71, :126, :179) -------- */
void euler(mtrx w, mtrx dwdx, mtrx dwdy, mtrx d2wdx2, mtrx d2wdy2, mtrx u, mtrx v, double Re, double dt);
mtrx continuity(mtrx dudx, mtrx dvdy);
mtrx vorticity(mtrx dvdx, mtrx dudy); /* returns second argument minus first, as the reference does */
void set_fluiddyn_openmp_config(int enabled);

/* ---- Poisson: reference include/poisson.h:9-20 (src/poisson.c:34, :62, :111, :176, :224) -------- */
#ifndef PI
#define PI 3.14159265359
#endif
double error(mtrx u1, mtrx u2);
mtrx poisson(mtrx f, double dx, double dy, int itmax, double tol);
mtrx poisson_SOR(mtrx f, double dx, double dy, int itmax, double tol, double beta);
mtrx poisson_log(mtrx f, double dx, double dy, int itmax, double tol, FILE *log_file);
mtrx poisson_SOR_log(mtrx f, double dx, double dy, int itmax, double tol, double beta, FILE *log_file);
void set_poisson_openmp_config(int enabled);

/* ---- configuration: reference include/config.h:4-43 (src/config.c:47, :106, :230, :294, :334) --- */
typedef struct {
    double Re;
    int Lx, Ly;
    int nx, ny;
    double dt, tf, max_co;
    int order;
    int poisson_max_it;
    double poisson_tol;
    int output_interval;
    int poisson_type; /* 1 Gauss-Seidel | 2 SOR */
    int openmp_enabled; /* parsed, reported, ignored by the GPU path */
    double ui, vi;
    double u1, u2, u3, u4;
    double v1, v2, v3, v4;
} Config;

Config load_default_config(void);
Config load_config_from_file(const char *filename);
void print_config(const Config *config);
void print_usage(const char *program_name);
void print_openmp_status(const Config *config);

#ifdef __cplusplus
}
#endif
#endif /* CNAVIER_DROPIN_H */
