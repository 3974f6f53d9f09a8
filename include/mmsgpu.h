/*
 * mmsgpu.h -- C ABI of libmmsgpu, the B200 (sm_100a) backend for the
 * mosmicrosim per-cell simulation chain.
 *
 * The reference (goldingn/mosmicrosim) is 100 % R and has NO native interface
 * (NAMESPACE:1-5, no src/).  The entry points below are therefore exactly what
 * a `.Call` shim in the reference's src/ would bind (see INTEGRATION.md and
 * r-binding/src/mmsgpu_shim.c): plain pointers and sizes, caller-owned host
 * buffers, int status return, no exceptions, no R / torch / C++ types.
 * Each entry point cites the reference function(s) it replaces, relative to
 * the reference root.
 *
 * Conventions
 *   - all floating point data is IEEE double (R numeric); NA/NaN flow through
 *   - "cell" = pixel; every cell is an independent time integration
 *   - fused-chain inputs/outputs are laid out with the CELL index fastest
 *     ([...][cell]); the stage-level twins take R's time-by-pixel matrices
 *     exactly as R stores them (column-major: element (t, c) at [c*H + t])
 *   - every function returns MMS_OK (0) or an mms_status; mms_last_error()
 *     returns a human-readable message for the calling thread
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with MMS_ERR_NO_DEVICE
 */
#ifndef MMSGPU_H
#define MMSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMS_ABI_VERSION 2

#if defined(__GNUC__)
#define MMS_API __attribute__((visibility("default")))
#else
#define MMS_API
#endif

typedef enum mms_status {
    MMS_OK = 0,
    MMS_ERR_INVALID_ARG = 1,
    MMS_ERR_EXTRAPOLATION = 2, /* dates_predict outside the monthly knots: the reference would fit
                                  an mgcv::gam (R/spline_interpolation.R:30-112); not supported */
    MMS_ERR_BURNIN = 3,        /* burn-in requested but fewer than 8760 hourly steps
                                  (R: "subscript out of bounds", R/sim_water.R:240) */
    MMS_ERR_NONFINITE = 4,     /* NA/NaN/Inf in the monthly forcing */
    MMS_ERR_CUDA = 5,
    MMS_ERR_NO_DEVICE = 6,
    MMS_ERR_ALLOC = 7,
    MMS_ERR_UNSUPPORTED = 8
} mms_status;

/* monthly forcing variables, in the column order of the reference's monthly
 * climate tibble (R/terraclimate.R:633-642) */
enum {
    MMS_F_TMIN = 0, MMS_F_TMAX, MMS_F_RHMIN, MMS_F_RHMAX, MMS_F_WSMIN, MMS_F_WSMAX,
    MMS_F_CCMIN, MMS_F_CCMAX, MMS_F_RAINFALL_DAILY, MMS_NFORCING
};

/* raw monthly TerraClimate variables, in the column order of the per-pixel tibble
 * extract_terraclimate_tile() builds (R/pipeline_functions.R:134-136) */
enum {
    MMS_T_TMAX = 0, MMS_T_TMIN, MMS_T_PPT, MMS_T_WS, MMS_T_VPD, MMS_T_SRAD, MMS_NTERRACLIMATE
};

/* output variables.  0-5 are the reference's vector_variables in its order
 * (R/pipeline_functions.R:1041-1048); 6-7 the water body
 * (R/pipeline_functions.R:879-891 carries larval_habitat_area); 8-12 the hourly
 * microclimate (R/sim_microclimate.R:139-169); 13 the water temperature (the
 * air temperature, R/sim_microclimate.R:163, unless it is modelled). */
enum {
    MMS_V_ADULT = 0, MMS_V_AQUATIC, MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS,
    MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME,
    MMS_V_AIR_TEMPERATURE, MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER,
    MMS_V_WATER_TEMPERATURE,
    MMS_NVARS
};
/* variables that depend on the habitat setting (one series per setting);
 * all others depend on the forcing only (one series per cell) */
#define MMS_STATE_VAR_MASK ((1u << MMS_V_ADULT) | (1u << MMS_V_AQUATIC) | \
                            (1u << MMS_V_LARVAL_HABITAT_AREA) | (1u << MMS_V_WATER_VOLUME) | \
                            (1u << MMS_V_WATER_TEMPERATURE))
#define MMS_PIPELINE_VAR_MASK 0x3Fu   /* the six vector_variables */
#define MMS_MAX_SETTINGS 64

/* Water-body ("habitat") setting.  shape 0 with leak_per_hour 0 is the
 * reference's 90-degree cone, R/sim_water.R:12-31,94-132,216-225.  shape 1
 * (cylinder) and leak_per_hour > 0 are EXTENSIONS that the reference only
 * lists as to-dos (README.md:38-40); their definition is in DESIGN.md. */
typedef struct mms_habitat {
    int32_t shape;             /* 0 cone, 1 cylinder */
    int32_t reserved;
    double  max_depth;         /* max_cone_depth / cylinder depth, metres */
    double  radius;            /* cylinder radius, metres (ignored for the cone) */
    double  inflow_multiplier; /* R/sim_water.R:225 */
    double  leak_per_hour;     /* fraction of the current volume lost per hour */
    double  initial_volume;    /* R/sim_water.R:222 */
} mms_habitat;

/* make_surv_densmod_function(), R/lifehistory_estimation_functions.R:525-587 */
typedef struct mms_densmod {
    int32_t type;              /* 0 = "logit", 1 = "cox_ph" (attr(fun, "type")) */
    int32_t reserved;
    double  dd_effect;
    double  surface_area_cm2;
} mms_densmod;

/* stats::splinefun(method = "fmm") table of a positive_spline() closure
 * (R/lifehistory_estimation_functions.R:236-250): environment(function_raw)$z */
typedef struct mms_spline {
    int32_t n;
    int32_t reserved;
    const double *x, *y, *b, *c, *d;
} mms_spline;

/* interpolation_info of make_temp_humid_interpolator()
 * (R/lifehistory_estimation_functions.R:1144-1190): z is the R matrix,
 * column-major, z[ix + nx*iy], x = temperature steps, y = humidity steps */
typedef struct mms_surface {
    int32_t nx, ny;
    const double *x, *y, *z;
} mms_surface;

typedef struct mms_lifehistory {
    mms_spline  mdr_temp;      /* evaluated at water temperature, R/sim_pop_dynamics.R:108 */
    mms_spline  efd_temp;      /* air temperature, :110 */
    mms_spline  das_temp;      /* water temperature, :114 */
    mms_surface ds_temp_humid; /* (air temperature, humidity), :116-119 */
} mms_lifehistory;

/* Water-temperature model -- an EXTENSION.  The reference only sketches it (comments,
 * query_terraclimate.R:370-497), stop()s on simulate_hourly_conditions(model_water_temperature =
 * TRUE) (R/pipeline_functions.R:334-338) and uses the air temperature (R/sim_microclimate.R:163).
 * The spec is authored in DESIGN.md section 5 from that prose: hourly energy balance of the water
 * body -- evaporative loss, long-wave exchange with the air, solar gain from the pixel's clear-sky
 * radiation distributed over the day by the solar zenith angle (latitude, day of year, hour),
 * reduced by the hourly cloud cover (Angstrom) and by the shade.  Parity unpinned by construction.
 * With enabled = 0 (the default) every result is the reference's, bit for bit.  Needs
 * mms_chain_set_location() and n_settings == 1 (the rates that take the water temperature become
 * a property of the water body). */
typedef struct mms_water_temperature {
    int32_t enabled;               /* model_water_temperature */
    int32_t reserved;
    double  shade_proportion;      /* water_shade_proportion in [0, 1]; 1 = fully shaded (no solar gain) */
    double  albedo;                /* of the water surface; the notes assume "very low": 0 */
    double  emissivity;            /* 0.97 */
    double  convection_w_m2_k;     /* sensible heat exchange with the air, W / m2 / K; 0 = the notes' model (none) */
    const int32_t *day_of_year;    /* [D] doy(date), R/utils.R:46-48 */
    const int32_t *month_of_day;   /* [D] calendar month 1..12 of each day (row of clear_sky) */
} mms_water_temperature;

/* Everything about a fused-chain run except the per-cell data. */
typedef struct mms_chain_config {
    int64_t n_cells;           /* C */
    int32_t n_months;          /* M monthly knots */
    int32_t n_days;            /* D simulated days; H = 24*D hourly steps */
    int32_t n_settings;        /* S habitat settings simulated for every cell (1..MMS_MAX_SETTINGS) */
    int32_t n_periods;         /* P summary periods (0 = no summaries) */
    const double  *mid_day;        /* [M] as.numeric(mid_date - min(mid_date)), R/spline_interpolation.R:15 */
    const double  *day_predict;    /* [D] as.numeric(dates_predict - min(mid_date)), :20 */
    const int32_t *period_of_day;  /* [D] period index of each day, -1 = not in included_dates
                                      (R/pipeline_functions.R:1050-1088); non-negative entries
                                      must be non-decreasing; may be NULL when P = 0 */
    const mms_habitat *habitats;   /* [S] */
    int32_t burnin_years_water;    /* R/sim_water.R:223 (pipeline passes 0, R/pipeline_functions.R:438) */
    int32_t burnin_years_pop;      /* R/sim_pop_dynamics.R:217 (default 1) */
    double  initial_adult;         /* R/sim_pop_dynamics.R:215 */
    double  initial_aquatic;       /* R/sim_pop_dynamics.R:216 */
    mms_densmod     densmod;
    mms_lifehistory lifehistory;
    uint32_t mean_var_mask;        /* bit v set: period means of variable v are produced */
    uint32_t hourly_var_mask;      /* bit v set: the full hourly series of variable v is produced */
    int32_t  want_wet_hours;       /* per-period count of hours with water volume > 0 */
    int32_t  want_wet_flags;       /* hourly wet/dry flag series */
    int32_t  math_mode;            /* 0 = fast (FMA contraction, exp/log-based powers; <= 1e-9 rel),
                                      1 = strict (no contraction, literal operation order) */
    int32_t  n_devices;            /* 0 = all visible devices */
    const int32_t *devices;        /* [n_devices] CUDA ordinals, or NULL for 0..n_devices-1 */
    mms_water_temperature water_temperature;   /* EXTENSION (ABI 2); all zero = the reference */
    /* Output rows in RASTER cell order (ABI 2): with raster_cell_idx != NULL every output row has
     * n_raster_cells columns instead of C, the value of pixel c lands in column raster_cell_idx[c]
     * (0-based: terra's cell number minus one) and raster cells that are not pixels of the tile
     * hold NaN (0 in the integer outputs) -- one row is one raster slice, what
     * create_vector_raster_slice() builds with raster[cell_idx] <- values
     * (R/pipeline_functions.R:1283-1325, cell_idx at :1316-1322).  The scatter happens on the device,
     * in the kernel's stores.  raster_cell_idx [C] must be strictly increasing (pixels in raster
     * order, as terra::cells() returns them). */
    const int64_t *raster_cell_idx;
    int64_t n_raster_cells;
} mms_chain_config;

/* ---- library ------------------------------------------------------------ */
MMS_API int         mms_abi_version(void);
MMS_API int         mms_device_count(void);            /* CUDA devices visible, 0 if none */
MMS_API const char *mms_last_error(void);              /* message of the last failure on this thread */

/* ---- fused chain ---------------------------------------------------------
 * Replaces, for a whole tile of pixels in one call, the reference verbs
 *   interpolate_daily_climate()     R/pipeline_functions.R:258-275
 *   simulate_hourly_microclimate()  R/pipeline_functions.R:284-301
 *   simulate_hourly_conditions()    R/pipeline_functions.R:319-538
 *   simulate_hourly_lifehistory()   R/pipeline_functions.R:634-683
 *   simulate_hourly_vectors()       R/pipeline_functions.R:691-906
 *   summarise_vectors()             R/pipeline_functions.R:995-1026
 * i.e. tc_to_vectors() (R/pipeline_functions.R:1099-1130) after
 * terraclimate_to_monthly_climate().
 *
 * monthly  [MMS_NFORCING][M][C] host, cell fastest
 * altitude [C] host
 * Outputs (host, caller-allocated, any may be NULL when not requested).  With
 * ns = number of state variables selected in the mask (ascending variable
 * id), nf = number of forcing-only variables selected, H = 24*D (and C replaced
 * by n_raster_cells when cfg->raster_cell_idx is given):
 *   out_mean_state    [S][ns][P][C]   out_mean_forcing   [nf][P][C]
 *   out_hourly_state  [S][ns][H][C]   out_hourly_forcing [nf][H][C]
 *   out_wet_hours     [S][P][C] int32 out_wet_flags      [S][H][C] uint8
 * Periods with no included day are filled with NaN (0 for wet hours).
 */
typedef struct mms_chain_outputs {
    double  *mean_state;
    double  *mean_forcing;
    double  *hourly_state;
    double  *hourly_forcing;
    int32_t *wet_hours;
    uint8_t *wet_flags;
} mms_chain_outputs;

MMS_API int mms_run_chain(const mms_chain_config *cfg, const double *monthly, const double *altitude,
                  const mms_chain_outputs *out);

/* Resident ("plan") form of the same call, for callers that keep a tile on
 * the device across several launches (species, habitat ensembles), for tile
 * loops that reuse one plan (mms_chain_run per tile), and for benchmarking the
 * kernel with inputs already in HBM.  A plan shards the cells contiguously over
 * cfg->devices and owns a team of host threads per device, so the shards of a
 * multi-GPU plan are fed, run and drained concurrently: one call on one
 * process replaces the reference's furrr fan-out over pixel batches
 * (query_terraclimate.R:117-123,166-175,196-241).
 *
 * Host buffers: every call below RETURNS ONLY WHEN IT NO LONGER NEEDS THE HOST
 * BUFFERS IT WAS GIVEN (uploads have left them, downloads have filled them);
 * mms_chain_launch alone is asynchronous.  Buffers from mms_host_alloc (or
 * cudaHostRegister'ed memory) are copied by DMA in place; ordinary pageable
 * memory -- R vectors, numpy arrays -- goes through the plan's pinned bounce
 * buffers, several host threads per device. */
typedef struct mms_chain_plan mms_chain_plan;
MMS_API int  mms_chain_create(const mms_chain_config *cfg, mms_chain_plan **plan);
MMS_API int  mms_chain_upload(mms_chain_plan *plan, const double *monthly, const double *altitude);
/* The same upload from the six raw TerraClimate variables: the nine forcing rows are derived on
 * the device as process_terraclimate_pixel_vars() derives them (R/terraclimate.R:581-643; see
 * mms_terraclimate_to_monthly below for the arguments). */
MMS_API int  mms_chain_upload_terraclimate(mms_chain_plan *plan, const double *terraclimate,
                                           const double *clear_sky, const int32_t *month_id,
                                           const double *n_days, const double *altitude);
/* upload + launch + download as ONE pipelined call: the shard of every device is cut into
 * chunks of about one round of the persistent kernel, and the host-to-device copy of chunk
 * k+1, the kernel of chunk k and the device-to-host copy of chunk k-1 overlap.  This is what
 * mms_run_chain() does between create and destroy. */
MMS_API int  mms_chain_run(mms_chain_plan *plan, const double *monthly, const double *altitude,
                           const mms_chain_outputs *out);
MMS_API int  mms_chain_run_terraclimate(mms_chain_plan *plan, const double *terraclimate,
                                        const double *clear_sky, const int32_t *month_id,
                                        const double *n_days, const double *altitude,
                                        const mms_chain_outputs *out);
/* one-shot form of that route (create, run_terraclimate, destroy):
 * tc_to_vectors() whole, R/pipeline_functions.R:1099-1130 */
MMS_API int  mms_run_chain_terraclimate(const mms_chain_config *cfg, const double *terraclimate,
                                        const double *clear_sky, const int32_t *month_id,
                                        const double *n_days, const double *altitude,
                                        const mms_chain_outputs *out);
/* per-cell inputs of the water-temperature model: latitude [C] (degrees) and the pixel's clear-sky
 * radiation by calendar month [12][C] (clear_sky_radiation_12mo(), the array the TerraClimate route
 * takes too); resident on the device until the plan is destroyed */
MMS_API int  mms_chain_set_location(mms_chain_plan *plan, const double *latitude, const double *clear_sky);
MMS_API int  mms_chain_launch(mms_chain_plan *plan);          /* asynchronous */
MMS_API int  mms_chain_sync(mms_chain_plan *plan);            /* waits, reports kernel-side errors */
MMS_API int  mms_chain_download(mms_chain_plan *plan, const mms_chain_outputs *out);
MMS_API int  mms_chain_kernel_ms(mms_chain_plan *plan, float *ms);   /* max over shards, last launch (mms_chain_run: sum over its chunks) */
/* how the plan cut the work: shards (devices), chunks of the first shard, host threads per shard */
MMS_API int  mms_chain_plan_info(mms_chain_plan *plan, int32_t *n_shards, int32_t *n_chunks_per_shard,
                                 int32_t *n_workers_per_shard);
MMS_API int  mms_chain_launch_count(mms_chain_plan *plan, int64_t *n);/* kernels launched so far */
MMS_API void mms_chain_destroy(mms_chain_plan *plan);
/* the contiguous cell range [cell0, cell0 + count) that shard `index` of `n_shards` owns
 * (pure host logic; what mms_chain_create uses, and what a one-process-per-GPU caller
 * uses to cut its own tile out of a raster) */
MMS_API int  mms_shard_range(int64_t n_cells, int32_t n_shards, int32_t index, int64_t *cell0, int64_t *count);

/* ---- stage-level twins (R matrices in, R matrices out) -------------------
 * simulate_ephemeral_habitat_vectorised(), R/sim_water.R:216-301.
 * Inputs are H-by-C column-major; water_surface_area is required,
 * water_volume and wet (uint8 H-by-C) may be NULL. */
MMS_API int mms_water_vectorised(const double *rainfall, const double *air_temperature,
                         const double *humidity, const double *windspeed,
                         const double *altitude, int64_t n_times, int64_t n_pixels,
                         const mms_habitat *habitat, int32_t burnin_years,
                         int32_t math_mode, int32_t device,
                         double *water_surface_area, double *water_volume, uint8_t *wet);

/* simulate_population_vectorised(), R/sim_pop_dynamics.R:208-284 */
MMS_API int mms_population_vectorised(const double *larval_habitat_area, const double *mdr,
                              const double *efd, const double *das_zerodensity,
                              const double *ds, int64_t n_times, int64_t n_pixels,
                              const mms_densmod *densmod, double initial_adult,
                              double initial_aquatic, int32_t burnin_years,
                              int32_t math_mode, int32_t device,
                              double *aquatic, double *adult);

/* The scalar twins simulate_ephemeral_habitat(), R/sim_water.R:136-205, and simulate_population(),
 * R/sim_pop_dynamics.R:138-196: the same recursions for one pixel (n_pixels = 1 gives the R functions'
 * vector in / vector out; any number of pixels is accepted, same H-by-C layout), but their burn-in
 * repeats the WHOLE series -- index <- rep(seq_len(n_times), burnin_years + 1), R/sim_water.R:151-152,
 * R/sim_pop_dynamics.R:147-148 -- not the first 8760 rows, so any n_times is valid. */
MMS_API int mms_water_series(const double *rainfall, const double *air_temperature,
                     const double *humidity, const double *windspeed,
                     const double *altitude, int64_t n_times, int64_t n_pixels,
                     const mms_habitat *habitat, int32_t burnin_years,
                     int32_t math_mode, int32_t device,
                     double *water_surface_area, double *water_volume, uint8_t *wet);

MMS_API int mms_population_series(const double *larval_habitat_area, const double *mdr,
                          const double *efd, const double *das_zerodensity,
                          const double *ds, int64_t n_times, int64_t n_pixels,
                          const mms_densmod *densmod, double initial_adult,
                          double initial_aquatic, int32_t burnin_years,
                          int32_t math_mode, int32_t device,
                          double *aquatic, double *adult);

/* summarise_vectors() -> aggregate_vectors(), R/pipeline_functions.R:995-1091: period means of hourly series.
 *   hourly        [n_vars][24 * n_days][C]   cell fastest (the layout of the hourly outputs above)
 *   period_of_day [n_days]                   period of each day, -1 = not in included_dates; non-negative entries
 *                                            non-decreasing (by date, or by floor_date(date, "month"))
 *   means         [n_vars][n_periods][C]     out; NaN for a period without a day
 * sum / n in the order of the hours, in double, exactly as the fused chain accumulates its summaries
 * (the reference: R mean(), long-double two-pass -- the difference is far below the 1e-9 tolerance). */
MMS_API int mms_summarise_vectors(const double *hourly, int64_t n_cells, int32_t n_days, int32_t n_vars,
                                  const int32_t *period_of_day, int32_t n_periods, int32_t device, double *means);

/* ---- state-free verbs, for callers that want the intermediate products ----------------
 * (layouts with the cell index fastest, like the fused chain)
 *
 * terraclimate_to_monthly_climate() -> process_terraclimate_pixel_vars(), R/pipeline_functions.R:179-252,
 * R/terraclimate.R:581-643, for all pixels of a tile:
 *   terraclimate [MMS_NTERRACLIMATE][M][C]  tmax, tmin (deg C), ppt (mm / month), ws (m/s at 10 m), vpd (kPa), srad (W/m2)
 *   clear_sky    [12][C]                    clear_sky_SOLR of the pixel by calendar month (R/pipeline_functions.R:194-222)
 *   month_id     [M]                        lubridate::month(start + (end - start) / 2), 1..12 (R/terraclimate.R:591)
 *   n_days       [M]                        as.numeric(end - start): days in the month MINUS ONE, as the reference has it (:592)
 *   monthly      [MMS_NFORCING][M][C]       out
 * Saturation vapour pressure is NicheMapR's VAPPRS (Goff-Gratch), restated: third-party code that is
 * not part of the reference tree -- parity unpinned, see oracle/mms_oracle.c mmso_vapprs(). */
MMS_API int mms_terraclimate_to_monthly(const double *terraclimate, const double *clear_sky,
                                        const int32_t *month_id, const double *n_days, int64_t n_cells,
                                        int32_t n_months, int32_t math_mode, int32_t device, double *monthly);

/*
 * interpolate_daily_climate() -> daily_from_monthly_climate(), R/pipeline_functions.R:258-275,
 * R/sim_microclimate.R:7-64:  monthly [9][M][C] -> daily [9][D][C] */
MMS_API int mms_daily_from_monthly(const double *monthly, const double *mid_day, const double *day_predict,
                                   int64_t n_cells, int32_t n_months, int32_t n_days, int32_t math_mode,
                                   int32_t device, double *daily);
/* simulate_hourly_microclimate() -> hourly_from_daily_climate_ambient(), R/pipeline_functions.R:284-301,
 * R/sim_microclimate.R:126-173:  daily [9][D][C] -> hourly [6][24 D][C] = air_temperature, humidity,
 * cloudcover, windspeed, water_temperature, rainfall */
MMS_API int mms_hourly_from_daily(const double *daily, int64_t n_cells, int32_t n_days, int32_t math_mode,
                                  int32_t device, double *hourly);
/* simulate_hourly_lifehistory() -> simulate_vector_lifehistory(), R/pipeline_functions.R:634-683,
 * R/sim_pop_dynamics.R:100-135: element-wise over n values (any layout) */
MMS_API int mms_lifehistory_vectorised(const double *water_temperature, const double *air_temperature,
                                       const double *humidity, int64_t n, const mms_lifehistory *lifehistory,
                                       int32_t math_mode, int32_t device, double *mdr, double *efd,
                                       double *das_zerodensity, double *ds);

/* device time (CUDA events) of the kernel of the last stage-level twin call on this thread */
MMS_API int mms_stage_kernel_ms(float *ms);

/* ---- pinned host buffers --------------------------------------------------
 * Optional: host<->device copies of buffers obtained here are DMA'd in place;
 * any other host memory works too (staged through the plan's bounce buffers). */
MMS_API int mms_host_alloc(void **ptr, uint64_t bytes);
MMS_API int mms_host_free(void *ptr);

/* ---- measurement helpers (used by bench.py; not part of the R binding) --- */
MMS_API int mms_measure_fp64_peak(int32_t device, double *tflops);  /* dependent-free DFMA loop */
MMS_API int mms_measure_hbm_copy(int32_t device, double *gbs);      /* device-to-device copy, read+write bytes */
MMS_API int mms_flush_l2(int32_t device);                           /* overwrite a buffer larger than L2 */
MMS_API int mms_device_info(int32_t device, int32_t *n_sm, int32_t *sm_clock_khz_max, uint64_t *memory_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MMSGPU_H */
