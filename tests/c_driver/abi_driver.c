/* abi_driver.c -- a plain-C caller of libmmsgpu (include/mmsgpu.h), the way the reference's .Call shim
 * (r-binding/src/mmsgpu_shim.c) calls it: structs filled by hand, plain pointers, status codes.
 * tests/test_gpu_chain.py::test_plain_c_caller writes the inputs as raw little-endian arrays into a
 * directory, runs this program on it and compares the outputs with the ctypes route bit for bit.
 *
 *   abi_driver <dir>      reads dims.txt + *.f64 / *.i32, writes mean_state.f64, mean_forcing.f64, wet_hours.i32
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mmsgpu.h"

static void *slurp(const char *dir, const char *name, size_t bytes)
{
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    void *p = malloc(bytes ? bytes : 1);
    if (!f || !p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %zu bytes of %s\n", bytes, path); exit(2); }
    fclose(f);
    return p;
}

static void spill(const char *dir, const char *name, const void *p, size_t bytes)
{
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fclose(f);
}

static mms_spline spline(const char *dir, const char *stem, int n)
{
    char name[64];
    const char *part[5] = {"x", "y", "b", "c", "d"};
    const double *a[5];
    for (int k = 0; k < 5; ++k) {
        snprintf(name, sizeof name, "%s_%s.f64", stem, part[k]);
        a[k] = (const double *)slurp(dir, name, sizeof(double) * n);
    }
    mms_spline s = {n, 0, a[0], a[1], a[2], a[3], a[4]};
    return s;
}

int main(int argc, char **argv)
{
    if (argc != 2) { fprintf(stderr, "usage: abi_driver <dir>\n"); return 2; }
    const char *dir = argv[1];
    char path[4096];
    snprintf(path, sizeof path, "%s/dims.txt", dir);
    FILE *f = fopen(path, "r");
    long long C; int M, D, S, P, n_knots, nx, ny, ns, nf; unsigned mask;
    if (!f || fscanf(f, "%lld %d %d %d %d %d %d %d %u %d %d", &C, &M, &D, &S, &P, &n_knots, &nx, &ny, &mask, &ns, &nf) != 11) {
        fprintf(stderr, "bad dims.txt\n");
        return 2;
    }
    fclose(f);
    if (mms_abi_version() != MMS_ABI_VERSION) { fprintf(stderr, "unexpected ABI version\n"); return 3; }

    mms_chain_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_cells = C; cfg.n_months = M; cfg.n_days = D; cfg.n_settings = S; cfg.n_periods = P;
    cfg.mid_day = (const double *)slurp(dir, "mid_day.f64", sizeof(double) * M);
    cfg.day_predict = (const double *)slurp(dir, "day_predict.f64", sizeof(double) * D);
    cfg.period_of_day = (const int32_t *)slurp(dir, "period_of_day.i32", sizeof(int32_t) * D);
    cfg.habitats = (const mms_habitat *)slurp(dir, "habitats.bin", sizeof(mms_habitat) * S);
    cfg.burnin_years_water = 0;
    cfg.burnin_years_pop = 1;
    cfg.initial_adult = 100.0;
    cfg.initial_aquatic = 100.0;
    cfg.densmod = *(const mms_densmod *)slurp(dir, "densmod.bin", sizeof(mms_densmod));
    cfg.lifehistory.mdr_temp = spline(dir, "mdr", n_knots);
    cfg.lifehistory.efd_temp = spline(dir, "efd", n_knots);
    cfg.lifehistory.das_temp = spline(dir, "das", n_knots);
    cfg.lifehistory.ds_temp_humid.nx = nx;
    cfg.lifehistory.ds_temp_humid.ny = ny;
    cfg.lifehistory.ds_temp_humid.x = (const double *)slurp(dir, "ds_x.f64", sizeof(double) * nx);
    cfg.lifehistory.ds_temp_humid.y = (const double *)slurp(dir, "ds_y.f64", sizeof(double) * ny);
    cfg.lifehistory.ds_temp_humid.z = (const double *)slurp(dir, "ds_z.f64", sizeof(double) * nx * ny);
    cfg.mean_var_mask = mask;
    cfg.want_wet_hours = 1;
    cfg.math_mode = 0;

    const double *monthly = (const double *)slurp(dir, "monthly.f64", sizeof(double) * MMS_NFORCING * M * C);
    const double *altitude = (const double *)slurp(dir, "altitude.f64", sizeof(double) * C);
    const size_t n_state = (size_t)S * ns * P * C, n_forc = (size_t)nf * P * C, n_wet = (size_t)S * P * C;
    mms_chain_outputs out;
    memset(&out, 0, sizeof out);
    out.mean_state = (double *)malloc(sizeof(double) * n_state);
    out.mean_forcing = (double *)malloc(sizeof(double) * n_forc);
    out.wet_hours = (int32_t *)malloc(sizeof(int32_t) * n_wet);

    /* an invalid call first: the status code and the message must come back, nothing may crash */
    mms_chain_config bad = cfg;
    bad.n_settings = 0;
    if (mms_run_chain(&bad, monthly, altitude, &out) != MMS_ERR_INVALID_ARG || !strlen(mms_last_error())) {
        fprintf(stderr, "expected MMS_ERR_INVALID_ARG with a message\n");
        return 4;
    }
    const int rc = mms_run_chain(&cfg, monthly, altitude, &out);
    if (rc != MMS_OK) { fprintf(stderr, "mms_run_chain: %d %s\n", rc, mms_last_error()); return 5; }
    spill(dir, "mean_state.f64", out.mean_state, sizeof(double) * n_state);
    spill(dir, "mean_forcing.f64", out.mean_forcing, sizeof(double) * n_forc);
    spill(dir, "wet_hours.i32", out.wet_hours, sizeof(int32_t) * n_wet);
    printf("ok %d devices\n", mms_device_count());
    return 0;
}
