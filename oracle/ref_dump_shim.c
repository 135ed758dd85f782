/*
 * ref_dump_shim.c -- raw fp64 field dumps from the UNMODIFIED reference executable.
 * TEST INFRASTRUCTURE ONLY.  Linked with -Wl,--wrap=printvtk into oracle/_ref/cnavier_{omp,ser}:
 * every printvtk call of the reference main loop (src/main.c:431-434) first appends the
 * field as raw little-endian doubles to $CNAVIER_DUMP_DIR/<title>.f64 (the shipped VTK
 * format keeps 6 decimals only, src/utils.c:88, useless for 1e-8 parity), then runs the real
 * writer unless CNAVIER_DUMP_ONLY is set.
 */
#include <stdio.h>
#include <stdlib.h>
#include "linearalg.h"

void __real_printvtk(mtrx A, char *title, const char *output_dir);

void __wrap_printvtk(mtrx A, char *title, const char *output_dir)
{
    const char *dir = getenv("CNAVIER_DUMP_DIR");
    if (dir) {
        char name[1024];
        snprintf(name, sizeof name, "%s/%s.f64", dir, title);
        FILE *f = fopen(name, "ab");
        if (f) {
            for (int i = 0; i < A.m; i++) fwrite(A.M[i], sizeof(double), (size_t)A.n, f);
            fclose(f);
        }
    }
    if (!getenv("CNAVIER_DUMP_ONLY")) __real_printvtk(A, title, output_dir);
}
