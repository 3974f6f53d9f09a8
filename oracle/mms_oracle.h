/*
 * mms_oracle.h -- CPU restatement of the mosmicrosim per-cell simulation chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mosmicrosim_b200/,
 * include/, r-binding/) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference is 100 % R and R is not installed here, the
 * reference ships no golden vectors for this path, and its fitted tables
 * (R/sysdata.rda) are missing.  This file restates the reference's R
 * arithmetic line by line (same evaluation order, same literal constants) and
 * is cross-checked against (i) the known-answer constants derivable from the
 * R source and (ii) an independently written, R-style vectorised numpy
 * restatement (oracle/restate_numpy.py).  Third-party arithmetic that is not
 * under /root/reference (base-R stats::approx, stats::splinefun "fmm",
 * fields::interp.surface, base-R mean) is restated from its published
 * algorithm; each function says so.
 *
 * All citations are relative to /root/reference.
 */
#ifndef MMS_ORACLE_H
#define MMS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* forcing variable order = column order of the reference's monthly tibble,
 * R/terraclimate.R:633-642 */
enum {
    MMSO_TMIN = 0, MMSO_TMAX, MMSO_RHMIN, MMSO_RHMAX, MMSO_WSMIN, MMSO_WSMAX,
    MMSO_CCMIN, MMSO_CCMAX, MMSO_RAIN, MMSO_NFORCING
};

/* chain output variables (period means).  The first six are the reference's
 * vector_variables, R/pipeline_functions.R:1041-1048, in that order. */
enum {
    MMSO_V_ADULT = 0, MMSO_V_AQUATIC, MMSO_V_MDR, MMSO_V_EFD, MMSO_V_DAS0, MMSO_V_DS,
    MMSO_V_AREA, MMSO_V_VOLUME, MMSO_V_AIRT, MMSO_V_RH, MMSO_V_WIND, MMSO_V_RAIN,
    MMSO_V_WATERT,            /* water temperature: the air temperature (R/sim_microclimate.R:163) unless modelled */
    MMSO_NVARS
};

typedef struct {
    int    shape;             /* 0 = 90-degree cone (reference, R/sim_water.R:12-31);
                                 1 = cylinder (EXTENSION, absent from the reference) */
    double max_depth;         /* max_cone_depth, R/sim_water.R:224 */
    double radius;            /* cylinder only */
    double inflow_multiplier; /* R/sim_water.R:225 */
    double leak_per_hour;     /* EXTENSION (absent from the reference); 0 = reference */
    double initial_volume;    /* R/sim_water.R:222 */
} mmso_habitat;

typedef struct {
    int    type;              /* 0 = logit, 1 = cox_ph;
                                 R/lifehistory_estimation_functions.R:525-587 */
    double dd_effect;
    double surface_area_cm2;
} mmso_densmod;

typedef struct {              /* stats::splinefun(method = "fmm") coefficient table */
    int n;
    const double *x, *y, *b, *c, *d;
} mmso_spline;

typedef struct {              /* fields::interp.surface object, z column-major [ix + nx*iy] */
    int nx, ny;
    const double *x, *y, *z;
} mmso_surface;

typedef struct {
    mmso_spline  mdr, efd, das;   /* positive_spline closures */
    mmso_surface ds;              /* make_temp_humid_interpolator */
} mmso_lifehistory;

/* ---- S: monthly -> daily ------------------------------------------------ */
double mmso_approx1(double v, const double *x, const double *y, long n);
int mmso_spline_seasonal(const double *values, const double *datenum, int M,
                         const double *datenum_predict, int D, double *out);
int mmso_daily_from_monthly(const double *monthly, const double *mid_day, int M,
                            const double *day_predict, int D, double *daily);

/* ---- M: daily -> hourly ------------------------------------------------- */
void mmso_interpolate_daily_air_temp(const double *daily_max, const double *daily_min,
                                     int D, double *out);
void mmso_vsine_weights(const double *times, long n, double time_sunrise,
                        double time_inflection, double *weight);
void mmso_vsine_interpolate_daily(const double *daily_max, const double *daily_min,
                                  int D, int sunrise_is_max, double *out);
void mmso_hourly_from_daily_ambient(const double *daily, int D,
                                    double *air_temperature, double *humidity,
                                    double *cloudcover, double *windspeed,
                                    double *water_temperature, double *rainfall);

/* ---- W: water body ------------------------------------------------------ */
double mmso_R_pow(double x, double y);
double mmso_cone_depth_to_volume(double depth);
double mmso_cone_volume_to_surface(double volume);
double mmso_saturation_vapour_pressure_kpa(double t);
double mmso_air_pressure_kpa(double altitude_m);
double mmso_evaporation_rate_kg_h(double t, double rh, double alt, double wind, double area);
double mmso_habitat_max_volume(const mmso_habitat *hb);
double mmso_habitat_surface(const mmso_habitat *hb, double volume);
double mmso_habitat_catchment(const mmso_habitat *hb);
double mmso_iterate_volume(const mmso_habitat *hb, double volume, double rainfall_mm_h,
                           double evap_rate_kg_h_m2, double max_volume, double catchment);
int mmso_simulate_ephemeral_habitat_vectorised(
    const double *rainfall, const double *air_temperature, const double *humidity,
    const double *windspeed, const double *altitude, long H, long C,
    const mmso_habitat *hb, int burnin_years,
    double *water_surface_area, double *water_volume, unsigned char *wet);

/* ---- water temperature (EXTENSION: the model sketched in query_terraclimate.R:370-497; the reference itself
 * stop()s on model_water_temperature = TRUE, R/pipeline_functions.R:334-338, and copies the air temperature,
 * R/sim_microclimate.R:163).  Spec authored in DESIGN.md section 5; parity unpinned by construction. ---- */
typedef struct {
    int    enabled;            /* model_water_temperature */
    double shade_proportion;   /* water_shade_proportion in [0, 1]; 1 = fully shaded, no solar gain */
    double albedo;             /* of the water surface ("assume very low albedo for water") */
    double emissivity;         /* "water is ~0.97-0.99" */
    double convection_w_m2_k;  /* sensible heat exchange with the air; 0 = the notes' model (none) */
    const int *day_of_year;    /* [D] doy(date), R/utils.R:46-48 */
    const int *month_of_day;   /* [D] calendar month 1..12 (row of clear_sky) */
    const double *latitude;    /* [C] degrees */
    const double *clear_sky;   /* [12][C] clear-sky radiation by calendar month, W/m2 (clear_sky_radiation_12mo()) */
} mmso_water_temperature;
/* sin / cos of the solar declination of a day of the year (query_terraclimate.R:434-457) */
void mmso_solar_declination(int doy, double *sin_delta, double *cos_delta);
/* cos of the hour angle of hour-of-day k = 1..24: h = 15 (k - 12) degrees (:459-466) */
double mmso_cos_hour_angle(int hour_1_to_24);
/* clear-sky irradiance of each of the 24 hours of a day whose daily MEAN is clear_sky_mean (W/m2): distributed in
 * proportion to max(0, cos Z), cos Z = cos(lat) cos(delta) cos(h) + sin(lat) sin(delta) (:468-472) */
void mmso_clear_sky_hourly(double latitude_deg, int doy, double clear_sky_mean, double *irradiance24);
/* one hourly step of water volume AND water temperature */
void mmso_iterate_volume_temperature(const mmso_habitat *hb, const mmso_water_temperature *wt,
                                     double *volume_io, double *water_temperature_io,
                                     double rainfall_mm_h, double evap_rate_kg_h_m2, double air_temperature,
                                     double cloud_cover_pct, double clear_sky_irradiance,
                                     double max_volume, double catchment_area);

/* ---- T: TerraClimate variables -> monthly forcing (R/terraclimate.R:581-643) ------- */
double mmso_vapprs(double db);
double mmso_vapour_pressure(double tmean, double vpd_kpa);
double mmso_relative_humidity(double temperature, double vapour_pressure);
double mmso_adjust_wind_speed(double wind_speed, double height_required, double height_measured,
                              double exponent);
double mmso_cloud_cover(double solar_radiation, double clear_sky_radiation, double multiplier);
void mmso_process_terraclimate_pixel_vars(const double *terraclimate, const double *clear_sky,
                                          const int *month_id, const double *n_days,
                                          long C, int M, double *monthly);

/* ---- L: life history ---------------------------------------------------- */
void mmso_fmm_spline(int n, const double *x, const double *y, double *b, double *c, double *d);
void mmso_spline_eval(const mmso_spline *s, const double *u, long nu, double *v);
void mmso_positive_spline(const mmso_spline *s, const double *u, long nu, double *v);
void mmso_interp_surface(const mmso_surface *g, const double *tx, const double *ty,
                         long n, double *out);
double mmso_densmod_apply(const mmso_densmod *dm, double surv_prob, double density);

/* ---- P: population ------------------------------------------------------ */
void mmso_iterate_state(double *aquatic, double *adult, const mmso_densmod *dm,
                        double larval_habitat_area, double mdr, double efd,
                        double das_zerodensity, double ds);
int mmso_simulate_population_vectorised(
    const double *larval_habitat_area, const double *mdr, const double *efd,
    const double *das_zerodensity, const double *ds, long H, long C,
    const mmso_densmod *dm, double initial_adult, double initial_aquatic,
    int burnin_years, double *aquatic, double *adult);
/* the one-pixel twins, R/sim_water.R:136-205 and R/sim_pop_dynamics.R:138-196 (whole-series burn-in) */
int mmso_simulate_ephemeral_habitat(
    const double *rainfall, const double *air_temperature, const double *humidity,
    const double *windspeed, double altitude, long n_times,
    const mmso_habitat *hb, int burnin_years, double *surface_areas);
int mmso_simulate_population(
    const double *larval_habitat_area, const double *mdr, const double *efd,
    const double *das_zerodensity, const double *ds, long n_times,
    const mmso_densmod *dm, double initial_adult, double initial_aquatic,
    int burnin_years, double *aquatic, double *adult);
double mmso_R_mean(const double *x, long n);

/* ---- fused chain (composition of the stages above, per cell) ------------ */
typedef struct {
    int  M, D, C, S, P;
    const double *monthly;      /* [9][M][C], cell fastest */
    const double *mid_day;      /* [M] */
    const double *day_predict;  /* [D] */
    const double *altitude;     /* [C] */
    const mmso_habitat *habitats; /* [S] */
    int burnin_years_water;     /* R/sim_water.R:223 */
    int burnin_years_pop;       /* R/sim_pop_dynamics.R:217 */
    double initial_adult, initial_aquatic;
    mmso_densmod densmod;
    mmso_lifehistory lh;
    const int *period_of_day;   /* [D]; -1 = excluded (included_dates filter) */
    mmso_water_temperature wt;  /* EXTENSION; enabled = 0: water temperature = air temperature (the reference) */
    /* outputs, any may be NULL */
    double *out_mean;           /* [S][MMSO_NVARS][P][C] period means */
    int    *out_wet_hours;      /* [S][P][C] hours with volume > 0 */
    double *out_hourly;         /* [S][MMSO_NVARS][24*D][C] full series */
    unsigned char *out_wet;     /* [S][24*D][C] wet flag */
    int n_threads;              /* OpenMP threads over cells; <=0 = all */
} mmso_chain;
int mmso_run_chain(const mmso_chain *ch);
int mmso_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
