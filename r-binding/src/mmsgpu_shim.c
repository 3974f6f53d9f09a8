/*
 * mmsgpu_shim.c -- the `.Call` shim a maintainer adds to the reference's src/
 * (the reference has no src/ today: NAMESPACE:1-5, no useDynLib).
 *
 * Responsibilities, and nothing else:
 *   - read-only REAL()/INTEGER() views of R-owned inputs (never modified, never
 *     retained past the call);
 *   - allocate R outputs with PROTECT(Rf_alloc*);
 *   - hand plain pointers and sizes to libmmsgpu (include/mmsgpu.h);
 *   - turn a non-zero status into Rf_error("%s", mms_last_error()), mirroring the
 *     reference's stop(..., call. = FALSE) convention (R/pipeline_functions.R:335-337).
 * There is NO arithmetic here, no Rcpp, and no CPU fallback: without a CUDA
 * device libmmsgpu returns MMS_ERR_NO_DEVICE and the call errors.
 *
 * Build (inside the reference package):  PKG_CPPFLAGS = -I<mmsgpu>/include,
 * PKG_LIBS = -L<mmsgpu> -lmmsgpu  in src/Makevars; add useDynLib(mosmicrosim) to
 * NAMESPACE.  R is not installed in the development image of this repository,
 * so this file is compiled there only against tests/r_stub/Rinternals.h (a
 * syntax / prototype check); see INTEGRATION.md.
 */
#include <stdint.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "mmsgpu.h"

static void check(int status)
{
    if (status != MMS_OK) Rf_error("%s", mms_last_error());
}

/* REAL() of an argument that must be a double vector / matrix / array (REAL() of an integer matrix would read
 * past its end): the R wrappers coerce with storage.mode<- / as.numeric, this is the backstop */
static double *real_arg(SEXP x, const char *name)
{
    if (TYPEOF(x) != REALSXP) Rf_error("mmsgpu: '%s' must be of type double", name);
    return REAL(x);
}

static int *int_arg(SEXP x, const char *name)
{
    if (TYPEOF(x) != INTSXP) Rf_error("mmsgpu: '%s' must be of type integer", name);
    return INTEGER(x);
}

/* an n_times x n_pixels double matrix of the expected dimensions */
static double *matrix_arg(SEXP m, int64_t H, int64_t C, const char *name)
{
    if (TYPEOF(m) != REALSXP || !Rf_isMatrix(m)) Rf_error("mmsgpu: '%s' must be a double matrix", name);
    if (Rf_nrows(m) != H || Rf_ncols(m) != C)
        Rf_error("mmsgpu: '%s' is %d x %d, expected %lld x %lld", name, Rf_nrows(m), Rf_ncols(m), (long long)H, (long long)C);
    return REAL(m);
}

/* an H x C double matrix; H * C may exceed 2^31 - 1 (a long vector), each dimension may not (R's dim is integer) */
static SEXP alloc_matrix(int64_t H, int64_t C)
{
    if (H > INT32_MAX || C > INT32_MAX) Rf_error("mmsgpu: matrix dimensions beyond R's integer dim");
    SEXP m = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)H * (R_xlen_t)C));
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = (int)H;
    INTEGER(dim)[1] = (int)C;
    Rf_setAttrib(m, R_DimSymbol, dim);
    UNPROTECT(2);
    return m;
}

/* list element by name, or R_NilValue */
static SEXP list_get_opt(SEXP list, const char *name)
{
    if (list == R_NilValue) return R_NilValue;
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    if (names == R_NilValue) return R_NilValue;
    for (R_xlen_t i = 0; i < Rf_xlength(list); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}

/* list element by name (R-exts idiom) */
static SEXP list_get(SEXP list, const char *name)
{
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    for (R_xlen_t i = 0; i < Rf_xlength(list); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    Rf_error("mmsgpu: list element '%s' is missing", name);
    return R_NilValue;
}

static mms_spline spline_from(SEXP z)        /* environment(function_raw)$z of a splinefun closure */
{
    mms_spline s;
    s.n = (int32_t)Rf_xlength(list_get(z, "x"));
    s.reserved = 0;
    s.x = REAL(list_get(z, "x"));
    s.y = REAL(list_get(z, "y"));
    s.b = REAL(list_get(z, "b"));
    s.c = REAL(list_get(z, "c"));
    s.d = REAL(list_get(z, "d"));
    return s;
}

static mms_habitat habitat_from(SEXP h)      /* list(shape, max_depth, radius, inflow_multiplier, leak_per_hour, initial_volume) */
{
    mms_habitat hb;
    hb.shape = Rf_asInteger(list_get(h, "shape"));
    hb.reserved = 0;
    hb.max_depth = Rf_asReal(list_get(h, "max_depth"));
    hb.radius = Rf_asReal(list_get(h, "radius"));
    hb.inflow_multiplier = Rf_asReal(list_get(h, "inflow_multiplier"));
    hb.leak_per_hour = Rf_asReal(list_get(h, "leak_per_hour"));
    hb.initial_volume = Rf_asReal(list_get(h, "initial_volume"));
    return hb;
}

static mms_densmod densmod_from(SEXP d)      /* list(type = 0L/1L, dd_effect, surface_area_cm2) */
{
    mms_densmod dm;
    dm.type = Rf_asInteger(list_get(d, "type"));
    dm.reserved = 0;
    dm.dd_effect = Rf_asReal(list_get(d, "dd_effect"));
    dm.surface_area_cm2 = Rf_asReal(list_get(d, "surface_area_cm2"));
    return dm;
}

/* .Call("mms_R_device_count") */
SEXP mms_R_device_count(void)
{
    return Rf_ScalarInteger(mms_device_count());
}

/* simulate_ephemeral_habitat_vectorised(), R/sim_water.R:216-301.
 * Matrices are R's own n_times x n_pixels column-major doubles. */
SEXP mms_R_water_vectorised(SEXP rainfall, SEXP air_temperature, SEXP humidity, SEXP windspeed,
                            SEXP altitude, SEXP habitat, SEXP burnin_years, SEXP math_mode, SEXP device)
{
    if (!Rf_isMatrix(rainfall)) Rf_error("mmsgpu: 'rainfall_matrix' must be a matrix");
    const int64_t H = Rf_nrows(rainfall), C = Rf_ncols(rainfall);
    const double *rn = matrix_arg(rainfall, H, C, "rainfall_matrix");
    const double *ta = matrix_arg(air_temperature, H, C, "air_temperature_matrix");
    const double *rh = matrix_arg(humidity, H, C, "humidity_matrix");
    const double *ws = matrix_arg(windspeed, H, C, "windspeed_matrix");
    if (Rf_xlength(altitude) != C) Rf_error("altitude_vector must have one element per pixel");
    mms_habitat hb = habitat_from(habitat);
    SEXP area = PROTECT(alloc_matrix(H, C));
    check(mms_water_vectorised(rn, ta, rh, ws, real_arg(altitude, "altitude_vector"), H, C, &hb,
                               Rf_asInteger(burnin_years), Rf_asInteger(math_mode),
                               Rf_asInteger(device), REAL(area), NULL, NULL));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 1));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 1));
    SET_VECTOR_ELT(out, 0, area);
    SET_STRING_ELT(names, 0, Rf_mkChar("water_surface_area"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(3);
    return out;                         /* list(water_surface_area = ) like R/sim_water.R:297-299 */
}

/* simulate_population_vectorised(), R/sim_pop_dynamics.R:208-284 */
SEXP mms_R_population_vectorised(SEXP larval_habitat_area, SEXP mdr, SEXP efd, SEXP das_zerodensity, SEXP ds,
                                 SEXP densmod, SEXP initial_adult, SEXP initial_aquatic, SEXP burnin_years,
                                 SEXP math_mode, SEXP device)
{
    if (!Rf_isMatrix(larval_habitat_area)) Rf_error("mmsgpu: 'larval_habitat_area_matrix' must be a matrix");
    const int64_t H = Rf_nrows(larval_habitat_area), C = Rf_ncols(larval_habitat_area);
    /* every matrix: double, and of the dimensions of the first (a mismatch would be an out-of-bounds host read) */
    const double *area_in = matrix_arg(larval_habitat_area, H, C, "larval_habitat_area_matrix");
    const double *mdr_in = matrix_arg(mdr, H, C, "mdr_matrix");
    const double *efd_in = matrix_arg(efd, H, C, "efd_matrix");
    const double *das_in = matrix_arg(das_zerodensity, H, C, "das_zerodensity_matrix");
    const double *ds_in = matrix_arg(ds, H, C, "ds_matrix");
    mms_densmod dm = densmod_from(densmod);
    SEXP aquatic = PROTECT(alloc_matrix(H, C));
    SEXP adult = PROTECT(alloc_matrix(H, C));
    check(mms_population_vectorised(area_in, mdr_in, efd_in, das_in, ds_in,
                                    H, C, &dm, Rf_asReal(initial_adult), Rf_asReal(initial_aquatic),
                                    Rf_asInteger(burnin_years), Rf_asInteger(math_mode), Rf_asInteger(device),
                                    REAL(aquatic), REAL(adult)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, aquatic);
    SET_VECTOR_ELT(out, 1, adult);
    SET_STRING_ELT(names, 0, Rf_mkChar("aquatic"));
    SET_STRING_ELT(names, 1, Rf_mkChar("adult"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(4);
    return out;                         /* list(aquatic = , adult = ) like R/sim_pop_dynamics.R:279-282 */
}

/* simulate_ephemeral_habitat(), R/sim_water.R:136-205: one pixel, plain vectors, whole-series burn-in.
 * Returns the numeric vector of surface areas (the R function's value, :203). */
SEXP mms_R_water_series(SEXP rainfall, SEXP air_temperature, SEXP humidity, SEXP windspeed, SEXP altitude,
                        SEXP habitat, SEXP burnin_years, SEXP math_mode, SEXP device)
{
    const int64_t H = Rf_xlength(rainfall);
    if (Rf_xlength(air_temperature) != H || Rf_xlength(humidity) != H || Rf_xlength(windspeed) != H)
        Rf_error("all hourly_climate columns must have the same length");
    mms_habitat hb = habitat_from(habitat);
    double alt = Rf_asReal(altitude);
    SEXP area = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)H));
    check(mms_water_series(REAL(rainfall), REAL(air_temperature), REAL(humidity), REAL(windspeed), &alt, H, 1, &hb,
                           Rf_asInteger(burnin_years), Rf_asInteger(math_mode), Rf_asInteger(device), REAL(area),
                           NULL, NULL));
    UNPROTECT(1);
    return area;
}

/* simulate_population(), R/sim_pop_dynamics.R:138-196: one pixel; list(aquatic = , adult = ) which the
 * R wrapper turns into the tibble of :190-193 */
SEXP mms_R_population_series(SEXP larval_habitat_area, SEXP mdr, SEXP efd, SEXP das_zerodensity, SEXP ds,
                             SEXP densmod, SEXP initial_adult, SEXP initial_aquatic, SEXP burnin_years,
                             SEXP math_mode, SEXP device)
{
    const int64_t H = Rf_xlength(larval_habitat_area);
    if (Rf_xlength(mdr) != H || Rf_xlength(efd) != H || Rf_xlength(das_zerodensity) != H || Rf_xlength(ds) != H)
        Rf_error("all hourly_lifehistory columns must have the same length");
    mms_densmod dm = densmod_from(densmod);
    SEXP aquatic = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)H));
    SEXP adult = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)H));
    check(mms_population_series(REAL(larval_habitat_area), REAL(mdr), REAL(efd), REAL(das_zerodensity), REAL(ds),
                                H, 1, &dm, Rf_asReal(initial_adult), Rf_asReal(initial_aquatic),
                                Rf_asInteger(burnin_years), Rf_asInteger(math_mode), Rf_asInteger(device),
                                REAL(aquatic), REAL(adult)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, aquatic);
    SET_VECTOR_ELT(out, 1, adult);
    SET_STRING_ELT(names, 0, Rf_mkChar("aquatic"));
    SET_STRING_ELT(names, 1, Rf_mkChar("adult"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(4);
    return out;
}

/* The fused chain: interpolate_daily_climate() ... summarise_vectors() for a whole tile,
 * R/pipeline_functions.R:258-1026.
 *   monthly        double [C, M, 9] array (R column-major = [9][M][C] with the cell index fastest)
 *   altitude       double [C]
 *   mid_day        double [M];  day_predict double [D];  period_of_day integer [D] (-1 = excluded)
 *   habitats       list of habitat lists;  densmod list;  tables list(mdr=z, efd=z, das=z, ds=list(x,y,z))
 *   mean_var_mask  integer;  n_periods integer;  burnin_years_pop integer;  math_mode integer
 *   devices        integer vector of CUDA ordinals (length 0 = all)
 *   options        NULL, or a named list with the extensions:
 *                    water_temperature = list(shade_proportion, albedo, emissivity, convection_w_m2_k,
 *                                             day_of_year = integer [D], month_of_day = integer [D],
 *                                             latitude = double [C], clear_sky = double [C, 12])
 *                    cell_idx = double [C] 0-based raster cell numbers (strictly increasing), n_raster_cells = double
 * Returns list(mean_state = [R, P, ns, S], mean_forcing = [R, P, nf], wet_hours = [R, P, S]) with R = C, or
 * n_raster_cells when cell_idx is given (every [ , p, v, s] column is then a whole raster slice). */
static SEXP run_chain_impl(SEXP monthly, SEXP terraclimate, SEXP altitude, SEXP mid_day, SEXP day_predict,
                           SEXP period_of_day, SEXP habitats, SEXP densmod, SEXP tables, SEXP mean_var_mask,
                           SEXP n_periods, SEXP burnin_years_pop, SEXP math_mode, SEXP devices, SEXP options)
{
    mms_chain_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_cells = Rf_xlength(altitude);
    cfg.n_months = (int32_t)Rf_xlength(mid_day);
    cfg.n_days = (int32_t)Rf_xlength(day_predict);
    cfg.n_settings = (int32_t)Rf_xlength(habitats);
    cfg.n_periods = Rf_asInteger(n_periods);
    if (terraclimate == R_NilValue && Rf_xlength(monthly) != (R_xlen_t)cfg.n_cells * cfg.n_months * MMS_NFORCING)
        Rf_error("monthly must be a [cells, months, 9] array");
    if (cfg.n_settings < 1 || cfg.n_settings > MMS_MAX_SETTINGS) Rf_error("1 to %d habitat settings", MMS_MAX_SETTINGS);
    cfg.mid_day = real_arg(mid_day, "mid_day");
    cfg.day_predict = real_arg(day_predict, "day_predict");
    if (Rf_xlength(period_of_day) != cfg.n_days) Rf_error("period_of_day must have one entry per simulated day");
    cfg.period_of_day = (const int32_t *)int_arg(period_of_day, "period_of_day");
    mms_habitat hab[MMS_MAX_SETTINGS];
    for (int s = 0; s < cfg.n_settings; ++s) hab[s] = habitat_from(VECTOR_ELT(habitats, s));
    cfg.habitats = hab;
    cfg.burnin_years_water = 0;                               /* R/pipeline_functions.R:438 */
    cfg.burnin_years_pop = Rf_asInteger(burnin_years_pop);    /* R/sim_pop_dynamics.R:217 */
    cfg.initial_adult = 100.0;                                /* R/sim_pop_dynamics.R:215-216 */
    cfg.initial_aquatic = 100.0;
    cfg.densmod = densmod_from(densmod);
    cfg.lifehistory.mdr_temp = spline_from(list_get(tables, "mdr"));
    cfg.lifehistory.efd_temp = spline_from(list_get(tables, "efd"));
    cfg.lifehistory.das_temp = spline_from(list_get(tables, "das"));
    SEXP ds = list_get(tables, "ds");
    cfg.lifehistory.ds_temp_humid.nx = (int32_t)Rf_xlength(list_get(ds, "x"));
    cfg.lifehistory.ds_temp_humid.ny = (int32_t)Rf_xlength(list_get(ds, "y"));
    cfg.lifehistory.ds_temp_humid.x = REAL(list_get(ds, "x"));
    cfg.lifehistory.ds_temp_humid.y = REAL(list_get(ds, "y"));
    cfg.lifehistory.ds_temp_humid.z = REAL(list_get(ds, "z"));
    cfg.mean_var_mask = (uint32_t)Rf_asInteger(mean_var_mask);
    cfg.want_wet_hours = 1;
    cfg.math_mode = Rf_asInteger(math_mode);
    cfg.n_devices = (int32_t)Rf_xlength(devices);
    cfg.devices = cfg.n_devices ? (const int32_t *)int_arg(devices, "devices") : NULL;

    /* extensions: the water-temperature model, rows scattered into raster cell order */
    SEXP wt = list_get_opt(options, "water_temperature");
    const double *latitude = NULL, *wt_clear_sky = NULL;
    if (wt != R_NilValue) {
        cfg.water_temperature.enabled = 1;
        cfg.water_temperature.shade_proportion = Rf_asReal(list_get(wt, "shade_proportion"));
        cfg.water_temperature.albedo = Rf_asReal(list_get(wt, "albedo"));
        cfg.water_temperature.emissivity = Rf_asReal(list_get(wt, "emissivity"));
        cfg.water_temperature.convection_w_m2_k = Rf_asReal(list_get(wt, "convection_w_m2_k"));
        SEXP doy = list_get(wt, "day_of_year"), mon = list_get(wt, "month_of_day");
        if (Rf_xlength(doy) != cfg.n_days || Rf_xlength(mon) != cfg.n_days)
            Rf_error("day_of_year and month_of_day must have one entry per simulated day");
        cfg.water_temperature.day_of_year = (const int32_t *)int_arg(doy, "day_of_year");
        cfg.water_temperature.month_of_day = (const int32_t *)int_arg(mon, "month_of_day");
        if (Rf_xlength(list_get(wt, "latitude")) != (R_xlen_t)cfg.n_cells ||
            Rf_xlength(list_get(wt, "clear_sky")) != (R_xlen_t)cfg.n_cells * 12)
            Rf_error("latitude must have one entry per pixel and clear_sky be a [cells, 12] matrix");
        latitude = real_arg(list_get(wt, "latitude"), "latitude");
        wt_clear_sky = real_arg(list_get(wt, "clear_sky"), "clear_sky");
    }
    SEXP cell_idx = list_get_opt(options, "cell_idx");
    R_xlen_t n_out = (R_xlen_t)cfg.n_cells;
    if (cell_idx != R_NilValue) {
        if (Rf_xlength(cell_idx) != (R_xlen_t)cfg.n_cells) Rf_error("cell_idx must have one entry per pixel");
        const double *ci = real_arg(cell_idx, "cell_idx");
        int64_t *idx = (int64_t *)R_alloc((size_t)cfg.n_cells, sizeof(int64_t));      /* freed by R at the end of .Call */
        for (int64_t c = 0; c < cfg.n_cells; ++c) idx[c] = (int64_t)ci[c];
        cfg.raster_cell_idx = idx;
        cfg.n_raster_cells = (int64_t)Rf_asReal(list_get(options, "n_raster_cells"));
        if (cfg.n_raster_cells < cfg.n_cells) Rf_error("n_raster_cells is smaller than the number of pixels");
        n_out = (R_xlen_t)cfg.n_raster_cells;
    }

    int ns = 0, nf = 0;
    for (int v = 0; v < MMS_NVARS; ++v)
        if (cfg.mean_var_mask >> v & 1u) { if (MMS_STATE_VAR_MASK >> v & 1u) ++ns; else ++nf; }
    const R_xlen_t C = n_out, P = cfg.n_periods, S = cfg.n_settings;
    SEXP mean_state = PROTECT(Rf_allocVector(REALSXP, C * P * ns * S));
    SEXP mean_forcing = PROTECT(Rf_allocVector(REALSXP, C * P * nf));
    SEXP wet_hours = PROTECT(Rf_allocVector(INTSXP, C * P * S));
    mms_chain_outputs out;
    memset(&out, 0, sizeof out);
    out.mean_state = REAL(mean_state);
    out.mean_forcing = REAL(mean_forcing);
    out.wet_hours = (int32_t *)INTEGER(wet_hours);
    /* create, [set_location,] run (upload, kernels and download pipelined in chunks), destroy -- the plan is
     * destroyed before an error is raised (Rf_error does not return) */
    mms_chain_plan *plan = NULL;
    check(mms_chain_create(&cfg, &plan));
    int rc = MMS_OK;
    if (wt != R_NilValue) rc = mms_chain_set_location(plan, latitude, wt_clear_sky);
    if (rc != MMS_OK) {
    } else if (terraclimate == R_NilValue) {
        rc = mms_chain_run(plan, real_arg(monthly, "monthly"), real_arg(altitude, "altitude"), &out);
    } else {
        /* list(values = [cells, months, 6], clear_sky = [cells, 12], month_id = integer [months], n_days = double [months]) */
        SEXP values = list_get(terraclimate, "values"), clear_sky = list_get(terraclimate, "clear_sky");
        SEXP month_id = list_get(terraclimate, "month_id"), n_days = list_get(terraclimate, "n_days");
        if (Rf_xlength(values) != (R_xlen_t)cfg.n_cells * cfg.n_months * MMS_NTERRACLIMATE ||
            Rf_xlength(clear_sky) != (R_xlen_t)cfg.n_cells * 12 || Rf_xlength(month_id) != cfg.n_months ||
            Rf_xlength(n_days) != cfg.n_months || TYPEOF(values) != REALSXP || TYPEOF(clear_sky) != REALSXP ||
            TYPEOF(month_id) != INTSXP || TYPEOF(n_days) != REALSXP || TYPEOF(altitude) != REALSXP) {
            mms_chain_destroy(plan);
            Rf_error("terraclimate must be list(values = double [cells, months, 6], clear_sky = double [cells, 12], "
                     "month_id = integer [months], n_days = double [months])");
        }
        rc = mms_chain_run_terraclimate(plan, REAL(values), REAL(clear_sky), (const int32_t *)INTEGER(month_id),
                                        REAL(n_days), REAL(altitude), &out);
    }
    mms_chain_destroy(plan);
    check(rc);

    SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(res, 0, mean_state);
    SET_VECTOR_ELT(res, 1, mean_forcing);
    SET_VECTOR_ELT(res, 2, wet_hours);
    SET_STRING_ELT(names, 0, Rf_mkChar("mean_state"));
    SET_STRING_ELT(names, 1, Rf_mkChar("mean_forcing"));
    SET_STRING_ELT(names, 2, Rf_mkChar("wet_hours"));
    Rf_setAttrib(res, R_NamesSymbol, names);
    UNPROTECT(5);
    return res;
}

SEXP mms_R_run_chain(SEXP monthly, SEXP altitude, SEXP mid_day, SEXP day_predict, SEXP period_of_day,
                     SEXP habitats, SEXP densmod, SEXP tables, SEXP mean_var_mask, SEXP n_periods,
                     SEXP burnin_years_pop, SEXP math_mode, SEXP devices, SEXP options)
{
    return run_chain_impl(monthly, R_NilValue, altitude, mid_day, day_predict, period_of_day, habitats, densmod,
                          tables, mean_var_mask, n_periods, burnin_years_pop, math_mode, devices, options);
}

/* The same from the raw TerraClimate variables (tc_to_vectors() whole, R/pipeline_functions.R:1099-1130):
 * process_terraclimate_pixel_vars(), R/terraclimate.R:581-643, runs on the device. */
SEXP mms_R_run_chain_terraclimate(SEXP terraclimate, SEXP altitude, SEXP mid_day, SEXP day_predict,
                                  SEXP period_of_day, SEXP habitats, SEXP densmod, SEXP tables,
                                  SEXP mean_var_mask, SEXP n_periods, SEXP burnin_years_pop, SEXP math_mode,
                                  SEXP devices, SEXP options)
{
    return run_chain_impl(R_NilValue, terraclimate, altitude, mid_day, day_predict, period_of_day, habitats,
                          densmod, tables, mean_var_mask, n_periods, burnin_years_pop, math_mode, devices, options);
}

/* process_terraclimate_pixel_vars() for a tile: values [C, M, 6], clear_sky [C, 12] -> monthly [C, M, 9] */
SEXP mms_R_terraclimate_to_monthly(SEXP values, SEXP clear_sky, SEXP month_id, SEXP n_days, SEXP n_cells,
                                   SEXP math_mode, SEXP device)
{
    const int64_t C = (int64_t)Rf_asReal(n_cells);
    const int M = (int)Rf_xlength(month_id);
    if (Rf_xlength(values) != (R_xlen_t)C * M * MMS_NTERRACLIMATE) Rf_error("values must be a [cells, months, 6] array");
    if (Rf_xlength(clear_sky) != (R_xlen_t)C * 12) Rf_error("clear_sky must be a [cells, 12] matrix");
    if (Rf_xlength(n_days) != M) Rf_error("month_id and n_days must have one entry per month");
    SEXP monthly = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C * M * MMS_NFORCING));
    check(mms_terraclimate_to_monthly(REAL(values), REAL(clear_sky), (const int32_t *)INTEGER(month_id), REAL(n_days),
                                      C, M, Rf_asInteger(math_mode), Rf_asInteger(device), REAL(monthly)));
    UNPROTECT(1);
    return monthly;
}

/* summarise_vectors() -> aggregate_vectors(), R/pipeline_functions.R:995-1091:
 * hourly [C, 24 D, n_vars] array, period_of_day integer [D] (-1 = excluded) -> means [C, P, n_vars] */
SEXP mms_R_summarise_vectors(SEXP hourly, SEXP period_of_day, SEXP n_periods, SEXP n_cells, SEXP n_vars, SEXP device)
{
    const int64_t C = (int64_t)Rf_asReal(n_cells);
    const int D = (int)Rf_xlength(period_of_day), V = Rf_asInteger(n_vars), P = Rf_asInteger(n_periods);
    if (Rf_xlength(hourly) != (R_xlen_t)C * 24 * D * V) Rf_error("hourly must be a [cells, 24 * days, variables] array");
    SEXP means = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C * P * V));
    check(mms_summarise_vectors(REAL(hourly), C, D, V, (const int32_t *)INTEGER(period_of_day), P, Rf_asInteger(device),
                                REAL(means)));
    UNPROTECT(1);
    return means;
}

/* interpolate_daily_climate(), R/pipeline_functions.R:258-275: monthly [C, M, 9] -> daily [C, D, 9] */
SEXP mms_R_daily_from_monthly(SEXP monthly, SEXP mid_day, SEXP day_predict, SEXP n_cells, SEXP math_mode, SEXP device)
{
    const int64_t C = (int64_t)Rf_asReal(n_cells);
    const int M = (int)Rf_xlength(mid_day), D = (int)Rf_xlength(day_predict);
    if (Rf_xlength(monthly) != (R_xlen_t)C * M * MMS_NFORCING) Rf_error("monthly must be a [cells, months, 9] array");
    SEXP daily = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C * D * MMS_NFORCING));
    check(mms_daily_from_monthly(REAL(monthly), REAL(mid_day), REAL(day_predict), C, M, D, Rf_asInteger(math_mode),
                                 Rf_asInteger(device), REAL(daily)));
    UNPROTECT(1);
    return daily;
}

/* simulate_hourly_microclimate(), R/pipeline_functions.R:284-301: daily [C, D, 9] -> hourly [C, 24 D, 6] */
SEXP mms_R_hourly_from_daily(SEXP daily, SEXP n_cells, SEXP n_days, SEXP math_mode, SEXP device)
{
    const int64_t C = (int64_t)Rf_asReal(n_cells);
    const int D = Rf_asInteger(n_days);
    if (Rf_xlength(daily) != (R_xlen_t)C * D * MMS_NFORCING) Rf_error("daily must be a [cells, days, 9] array");
    SEXP hourly = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C * 24 * D * 6));
    check(mms_hourly_from_daily(REAL(daily), C, D, Rf_asInteger(math_mode), Rf_asInteger(device), REAL(hourly)));
    UNPROTECT(1);
    return hourly;
}

/* simulate_hourly_lifehistory(), R/pipeline_functions.R:634-683: element-wise over the hourly vectors */
SEXP mms_R_lifehistory_vectorised(SEXP water_temperature, SEXP air_temperature, SEXP humidity, SEXP tables,
                                  SEXP math_mode, SEXP device)
{
    const R_xlen_t n = Rf_xlength(water_temperature);
    if (Rf_xlength(air_temperature) != n || Rf_xlength(humidity) != n) Rf_error("inputs must have the same length");
    mms_lifehistory lh;
    lh.mdr_temp = spline_from(list_get(tables, "mdr"));
    lh.efd_temp = spline_from(list_get(tables, "efd"));
    lh.das_temp = spline_from(list_get(tables, "das"));
    SEXP ds = list_get(tables, "ds");
    lh.ds_temp_humid.nx = (int32_t)Rf_xlength(list_get(ds, "x"));
    lh.ds_temp_humid.ny = (int32_t)Rf_xlength(list_get(ds, "y"));
    lh.ds_temp_humid.x = REAL(list_get(ds, "x"));
    lh.ds_temp_humid.y = REAL(list_get(ds, "y"));
    lh.ds_temp_humid.z = REAL(list_get(ds, "z"));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 4));
    const char *nm[4] = {"mdr", "efd", "das_zerodensity", "ds"};
    double *dst[4];
    for (int k = 0; k < 4; ++k) {
        SEXP v = PROTECT(Rf_allocVector(REALSXP, n));
        SET_VECTOR_ELT(out, k, v);
        UNPROTECT(1);
        SET_STRING_ELT(names, k, Rf_mkChar(nm[k]));
        dst[k] = REAL(v);
    }
    Rf_setAttrib(out, R_NamesSymbol, names);
    check(mms_lifehistory_vectorised(REAL(water_temperature), REAL(air_temperature), REAL(humidity), (int64_t)n, &lh,
                                     Rf_asInteger(math_mode), Rf_asInteger(device), dst[0], dst[1], dst[2], dst[3]));
    UNPROTECT(2);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"mms_R_daily_from_monthly", (DL_FUNC)&mms_R_daily_from_monthly, 6},
    {"mms_R_hourly_from_daily", (DL_FUNC)&mms_R_hourly_from_daily, 5},
    {"mms_R_lifehistory_vectorised", (DL_FUNC)&mms_R_lifehistory_vectorised, 6},
    {"mms_R_device_count", (DL_FUNC)&mms_R_device_count, 0},
    {"mms_R_water_vectorised", (DL_FUNC)&mms_R_water_vectorised, 9},
    {"mms_R_population_vectorised", (DL_FUNC)&mms_R_population_vectorised, 11},
    {"mms_R_water_series", (DL_FUNC)&mms_R_water_series, 9},
    {"mms_R_population_series", (DL_FUNC)&mms_R_population_series, 11},
    {"mms_R_run_chain", (DL_FUNC)&mms_R_run_chain, 14},
    {"mms_R_run_chain_terraclimate", (DL_FUNC)&mms_R_run_chain_terraclimate, 14},
    {"mms_R_terraclimate_to_monthly", (DL_FUNC)&mms_R_terraclimate_to_monthly, 7},
    {"mms_R_summarise_vectors", (DL_FUNC)&mms_R_summarise_vectors, 6},
    {NULL, NULL, 0}};

void R_init_mosmicrosim(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
