# gpu_backend.R -- R side of the GPU backend: what a maintainer adds to the
# reference's R/ directory next to src/mmsgpu_shim.c.  Signatures of the
# reference functions are unchanged; the backend is selected with
#   options(mosmicrosim.backend = "gpu")      # default "r": the existing R code
# There is no arithmetic here beyond date bookkeeping the reference already
# does in R, and no CPU fallback: with the GPU backend selected and no CUDA
# device the .Call errors.
#
# NOT EXECUTED in the development image of the B200 backend (no R there), and
# src/mmsgpu_shim.c has only been compiled against a stub of R's C API
# (tests/r_stub/): UNTESTED UNDER R.  The Python package mosmicrosim_b200/
# mirrors these wrappers on the same C ABI and is what the test-suite drives.

mms_backend <- function() {
  getOption("mosmicrosim.backend", "r")
}

mms_math_mode <- function() {
  # 0 = fast (default), 1 = strict (literal operation order, for debugging parity)
  as.integer(identical(getOption("mosmicrosim.gpu_math", "fast"), "strict"))
}

mms_devices <- function() {
  # integer(0) = all visible devices; a furrr/future worker should set this to
  # its own device, e.g. options(mosmicrosim.gpu_devices = 3L)
  as.integer(getOption("mosmicrosim.gpu_devices", integer(0)))
}

# habitat list for the shim; shape / radius / leak_per_hour are extensions of
# the reference's 90-degree cone (R/sim_water.R:12-31)
mms_habitat <- function(initial_volume = 0, max_cone_depth = 1, inflow_multiplier = 1,
                        shape = c("cone", "cylinder"), radius = 1, leak_per_hour = 0) {
  shape <- match.arg(shape)
  list(shape = if (shape == "cone") 0L else 1L,
       max_depth = as.numeric(max_cone_depth),
       radius = as.numeric(radius),
       inflow_multiplier = as.numeric(inflow_multiplier),
       leak_per_hour = as.numeric(leak_per_hour),
       initial_volume = as.numeric(initial_volume))
}

# plain parameters of a make_surv_densmod_function() closure
# (R/lifehistory_estimation_functions.R:525-587)
mms_densmod <- function(das_densmod_function) {
  env <- environment(das_densmod_function)
  list(type = if (attr(das_densmod_function, "type") == "logit") 0L else 1L,
       dd_effect = as.numeric(env$dd_effect),
       surface_area_cm2 = as.numeric(env$surface_area_cm2))
}

# coefficient table of a positive_spline() closure
# (R/lifehistory_estimation_functions.R:236-250): splinefun keeps it in `z`
mms_spline_table <- function(positive_spline_function) {
  raw <- environment(positive_spline_function)$function_raw
  z <- environment(raw)$z
  list(x = as.numeric(z$x), y = as.numeric(z$y), b = as.numeric(z$b),
       c = as.numeric(z$c), d = as.numeric(z$d))
}

# tables of one element of lifehistory_functions, with ds_temp_humid replaced by
# its bilinear emulator exactly as simulate_hourly_lifehistory() does
# (R/pipeline_functions.R:649-653)
mms_lifehistory_tables <- function(lifehistory) {
  ds_fun <- make_temp_humid_interpolator(lifehistory$ds_temp_humid)
  info <- environment(ds_fun)$interpolation_info
  list(mdr = mms_spline_table(lifehistory$mdr_temp),
       efd = mms_spline_table(lifehistory$efd_temp),
       das = mms_spline_table(lifehistory$das_temp),
       ds = list(x = as.numeric(info$x), y = as.numeric(info$y),
                 z = as.numeric(info$z)))
}

# ---- drop-in numeric cores ---------------------------------------------------

# simulate_ephemeral_habitat_vectorised(), R/sim_water.R:216-301
simulate_ephemeral_habitat_vectorised_gpu <- function(
    rainfall_matrix, air_temperature_matrix, humidity_matrix, windspeed_matrix,
    altitude_vector = rep(0, ncol(rainfall_matrix)),
    initial_volume = 0, burnin_years = 0, max_cone_depth = 1, inflow_multiplier = 1) {
  storage.mode(rainfall_matrix) <- "double"
  storage.mode(air_temperature_matrix) <- "double"
  storage.mode(humidity_matrix) <- "double"
  storage.mode(windspeed_matrix) <- "double"
  .Call("mms_R_water_vectorised", rainfall_matrix, air_temperature_matrix,
        humidity_matrix, windspeed_matrix, as.numeric(altitude_vector),
        mms_habitat(initial_volume, max_cone_depth, inflow_multiplier),
        as.integer(burnin_years), mms_math_mode(),
        if (length(mms_devices())) mms_devices()[1] else 0L,
        PACKAGE = "mosmicrosim")
}

# simulate_population_vectorised(), R/sim_pop_dynamics.R:208-284
simulate_population_vectorised_gpu <- function(
    larval_habitat_area_matrix, mdr_matrix, efd_matrix, das_zerodensity_matrix,
    ds_matrix, das_densmod_function, initial_adult = 100, initial_aquatic = 100,
    burnin_years = 1) {
  # the shim insists on double matrices of identical dimensions (an integer or mis-sized matrix
  # would be an out-of-bounds read in the host-to-device copy)
  storage.mode(larval_habitat_area_matrix) <- "double"
  storage.mode(mdr_matrix) <- "double"
  storage.mode(efd_matrix) <- "double"
  storage.mode(das_zerodensity_matrix) <- "double"
  storage.mode(ds_matrix) <- "double"
  stopifnot(identical(dim(mdr_matrix), dim(larval_habitat_area_matrix)),
            identical(dim(efd_matrix), dim(larval_habitat_area_matrix)),
            identical(dim(das_zerodensity_matrix), dim(larval_habitat_area_matrix)),
            identical(dim(ds_matrix), dim(larval_habitat_area_matrix)))
  .Call("mms_R_population_vectorised", larval_habitat_area_matrix, mdr_matrix,
        efd_matrix, das_zerodensity_matrix, ds_matrix,
        mms_densmod(das_densmod_function), as.numeric(initial_adult),
        as.numeric(initial_aquatic), as.integer(burnin_years), mms_math_mode(),
        if (length(mms_devices())) mms_devices()[1] else 0L,
        PACKAGE = "mosmicrosim")
}

# simulate_ephemeral_habitat(), R/sim_water.R:136-205 (one pixel; the burn-in repeats the whole series)
simulate_ephemeral_habitat_gpu <- function(hourly_climate, altitude, initial_volume = 0,
                                           burnin_years = 0, max_cone_depth = 1,
                                           inflow_multiplier = 1) {
  .Call("mms_R_water_series", as.numeric(hourly_climate$rainfall),
        as.numeric(hourly_climate$air_temperature), as.numeric(hourly_climate$humidity),
        as.numeric(hourly_climate$windspeed), as.numeric(altitude),
        mms_habitat(initial_volume = initial_volume, max_cone_depth = max_cone_depth,
                    inflow_multiplier = inflow_multiplier),
        as.integer(burnin_years), mms_math_mode(),
        if (length(mms_devices())) mms_devices()[1] else 0L,
        PACKAGE = "mosmicrosim")
}

# simulate_population(), R/sim_pop_dynamics.R:138-196 (one pixel; tibble(aquatic, adult) as at :190-193)
simulate_population_gpu <- function(hourly_lifehistory, das_densmod_function, initial_adult = 100,
                                    initial_aquatic = 100, burnin_years = 1) {
  res <- .Call("mms_R_population_series", as.numeric(hourly_lifehistory$larval_habitat_area),
               as.numeric(hourly_lifehistory$mdr), as.numeric(hourly_lifehistory$efd),
               as.numeric(hourly_lifehistory$das_zerodensity), as.numeric(hourly_lifehistory$ds),
               mms_densmod(das_densmod_function), as.numeric(initial_adult),
               as.numeric(initial_aquatic), as.integer(burnin_years), mms_math_mode(),
               if (length(mms_devices())) mms_devices()[1] else 0L,
               PACKAGE = "mosmicrosim")
  dplyr::tibble(aquatic = res$aquatic, adult = res$adult)
}

# ---- state-free verbs (for callers that want the intermediate tibbles) -------------

# daily_from_monthly_climate() for a whole tile, R/sim_microclimate.R:7-64.
# monthly: [cells, months, 9] array; returns a [cells, days, 9] array
daily_from_monthly_gpu <- function(monthly, mid_day, day_predict) {
  d <- .Call("mms_R_daily_from_monthly", monthly, as.numeric(mid_day), as.numeric(day_predict),
             as.numeric(dim(monthly)[1]), mms_math_mode(),
             if (length(mms_devices())) mms_devices()[1] else 0L, PACKAGE = "mosmicrosim")
  array(d, dim = c(dim(monthly)[1], length(day_predict), 9))
}

# hourly_from_daily_climate_ambient() for a whole tile, R/sim_microclimate.R:126-173.
# daily: [cells, days, 9]; returns [cells, 24 * days, 6] = air_temperature, humidity,
# cloudcover, windspeed, water_temperature, rainfall
hourly_from_daily_gpu <- function(daily) {
  h <- .Call("mms_R_hourly_from_daily", daily, as.numeric(dim(daily)[1]), as.integer(dim(daily)[2]),
             mms_math_mode(), if (length(mms_devices())) mms_devices()[1] else 0L, PACKAGE = "mosmicrosim")
  array(h, dim = c(dim(daily)[1], 24 * dim(daily)[2], 6))
}

# simulate_vector_lifehistory(), R/sim_pop_dynamics.R:100-135, on plain vectors
simulate_vector_lifehistory_gpu <- function(hourly_conditions, lifehistory) {
  res <- .Call("mms_R_lifehistory_vectorised", as.numeric(hourly_conditions$water_temperature),
               as.numeric(hourly_conditions$air_temperature), as.numeric(hourly_conditions$humidity),
               mms_lifehistory_tables(lifehistory), mms_math_mode(),
               if (length(mms_devices())) mms_devices()[1] else 0L, PACKAGE = "mosmicrosim")
  dplyr::tibble(date = hourly_conditions$date, hour = hourly_conditions$hour,
                mdr = res$mdr, efd = res$efd, das_zerodensity = res$das_zerodensity, ds = res$ds,
                larval_habitat_area = hourly_conditions$water_surface_area)
}

# summarise_vectors() -> aggregate_vectors() for a whole tile, R/pipeline_functions.R:995-1091.
# hourly: [cells, 24 * days, variables] array; dates_predict: the simulated days; returns
# list(start, end, means = [cells, periods, variables])
summarise_vectors_gpu <- function(hourly, dates_predict, aggregate_by = c("month", "day"),
                                  included_dates = NULL) {
  aggregate_by <- match.arg(aggregate_by)
  if (is.null(included_dates)) included_dates <- dates_predict
  keep <- dates_predict %in% included_dates
  key <- if (aggregate_by == "month") lubridate::floor_date(dates_predict, "month") else dates_predict
  periods <- sort(unique(key[keep]))
  period_of_day <- as.integer(ifelse(keep, match(key, periods) - 1L, -1L))
  m <- .Call("mms_R_summarise_vectors", hourly, period_of_day, length(periods),
             as.numeric(dim(hourly)[1]), as.integer(dim(hourly)[3]),
             if (length(mms_devices())) mms_devices()[1] else 0L, PACKAGE = "mosmicrosim")
  end <- if (aggregate_by == "month") lubridate::ceiling_date(periods + 1, "month") - 1 else periods
  list(start = periods, end = end,
       means = array(m, dim = c(dim(hourly)[1], length(periods), dim(hourly)[3])))
}

# ---- the fused pipeline --------------------------------------------------------

# [pixels, months, variables] array (the pixel index fastest, as the kernel wants it) from a list column of
# per-pixel monthly tibbles: ONE bind_rows() and one reshape -- not an interpreted loop over the pixels, which
# at the 1.5 M pixels of the Africa grid would be the run time
mms_pack_pixels <- function(list_col, vars) {
  n_pixels <- length(list_col)
  n_months <- nrow(list_col[[1]])
  all <- dplyr::bind_rows(list_col)
  if (nrow(all) != n_pixels * n_months) stop("all pixels of a tile must have the same months", call. = FALSE)
  a <- array(as.numeric(unlist(all[vars], use.names = FALSE)), dim = c(n_months, n_pixels, length(vars)))
  aperm(a, c(2, 1, 3))
}

# options list of the shim's extensions (NULL = the reference's behaviour)
mms_chain_options <- function(chain, pixels, model_water_temperature, water_shade_proportion, clear_sky,
                              cell_idx, n_raster_cells, water_temperature_options = list()) {
  # simulate_hourly_conditions() argument handling, R/pipeline_functions.R:324-338
  if (water_shade_proportion != 1 & !model_water_temperature) {
    warning("water temperature is not being modelled, ",
            "so water_shade_proportion argument is ignored", call. = FALSE)
  }
  opts <- list()
  if (model_water_temperature) {
    # EXTENSION: the reference stop()s here; the GPU backend models the water temperature (DESIGN.md section 5)
    if (is.null(clear_sky)) stop("model_water_temperature = TRUE needs the pixels' clear-sky radiation", call. = FALSE)
    wt <- utils::modifyList(list(albedo = 0, emissivity = 0.97, convection_w_m2_k = 0), water_temperature_options)
    opts$water_temperature <- c(wt, list(
      shade_proportion = as.numeric(water_shade_proportion),
      day_of_year = as.integer(format(chain$dates_predict, "%j")),          # doy(), R/utils.R:46-48
      month_of_day = as.integer(lubridate::month(chain$dates_predict)),
      latitude = as.numeric(pixels$latitude),
      clear_sky = clear_sky))                                               # [pixels, 12]
  }
  if (!is.null(cell_idx)) {
    # rows come back as whole raster slices: raster[cell_idx] <- values of create_vector_raster_slice()
    # (R/pipeline_functions.R:1316-1322) done by the kernel's stores; terra cell numbers are 1-based
    opts$cell_idx <- as.numeric(cell_idx) - 1
    opts$n_raster_cells <- as.numeric(n_raster_cells)
  }
  if (length(opts)) opts else NULL
}

# tc_to_vectors() (R/pipeline_functions.R:1099-1130) from interpolate_daily_climate()
# on, for the whole tile in one call.  pixel_monthly_climate is the tibble that
# terraclimate_to_monthly_climate() returns; the result has the schema of
# summarise_vectors(): longitude, latitude, species, pixel_vectors = list(tibble(
# start, end, adult, aquatic, mdr, efd, das_zerodensity, ds)).
#   on_extrapolation: the reference's default dates_predict can overhang the monthly mid-dates by half a day (a
#     30-day first month, a non-leap February last month: R/sim_microclimate.R:14-24,31), where the reference fits an
#     mgcv::gam (R/spline_interpolation.R:30-112) that the GPU backend does not reproduce: "error" (default) stops,
#     "clip" drops the overhanging prediction days with a warning.
#   cell_idx, n_raster_cells: return whole raster slices instead of tile rows (simplify = TRUE implied).
#   na_knots: the library refuses NA in the monthly forcing ("error", default); "interpolate" first replaces every NA
#     knot that has non-NA neighbours on both sides by what stats::approx() -- which drops NA pairs,
#     R/spline_interpolation.R:25-27 -- interpolates at that knot's date (mms_fill_na_knots()).
monthly_climate_to_vectors_gpu <- function(pixel_monthly_climate,
                                           species = c("An. gambiae", "An. stephensi"),
                                           included_dates = NULL,
                                           aggregate_by = c("month", "day"),
                                           model_water_temperature = FALSE,
                                           water_shade_proportion = 1,
                                           clear_sky = NULL,
                                           on_extrapolation = c("error", "clip"),
                                           cell_idx = NULL, n_raster_cells = NULL,
                                           simplify = !is.null(cell_idx),
                                           na_knots = c("error", "interpolate")) {
  species <- match.arg(species)
  aggregate_by <- match.arg(aggregate_by)
  na_knots <- match.arg(na_knots)
  lifehistory <- switch(species,
                        "An. gambiae" = lifehistory_functions$An_gambiae,
                        "An. stephensi" = lifehistory_functions$An_stephensi)
  first <- pixel_monthly_climate$monthly_climate[[1]]
  vars <- c("tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax",
            "rainfall_daily")
  monthly <- mms_pack_pixels(pixel_monthly_climate$monthly_climate, vars)
  chain <- mms_chain_setup(first, included_dates, aggregate_by, match.arg(on_extrapolation))
  if (na_knots == "interpolate" && anyNA(monthly)) monthly <- mms_fill_na_knots(monthly, chain$mid_day)
  opts <- mms_chain_options(chain, pixel_monthly_climate, model_water_temperature, water_shade_proportion,
                            clear_sky, cell_idx, n_raster_cells)
  res <- .Call("mms_R_run_chain", monthly, as.numeric(pixel_monthly_climate$altitude),
               chain$mid_day, chain$day_predict, chain$period_of_day,
               list(mms_habitat()),                # R/pipeline_functions.R:431-441
               mms_densmod(lifehistory$das_densmod), mms_lifehistory_tables(lifehistory),
               63L,                                # the six vector_variables
               length(chain$periods), 1L, mms_math_mode(), mms_devices(), opts,
               PACKAGE = "mosmicrosim")
  mms_vectors_result(res, pixel_monthly_climate, species, chain$periods, aggregate_by, simplify)
}

# NA knots of the packed forcing array [pixels, months, 9]: approx(x, y, xout) drops the (x, y) pairs with an NA, i.e.
# it interpolates between the neighbouring knots that are left; replacing the NA knot by approx()'s own value at the
# knot's date gives the same daily series (rainfall in the log1p space it is interpolated in,
# R/sim_microclimate.R:33,61).  NA runs at either end of a series stay NA (the library then refuses the tile; the
# reference has NA days there).  Only pixels with an NA are touched.
mms_fill_na_knots <- function(monthly, mid_day) {
  rain <- dim(monthly)[3]                          # rainfall_daily is the last of the nine variables
  for (v in seq_len(dim(monthly)[3])) {
    for (p in which(rowSums(is.na(monthly[, , v, drop = FALSE])) > 0)) {
      y <- monthly[p, , v]
      if (v == rain) y <- log1p(y)
      if (sum(!is.na(y)) < 2) next
      f <- stats::approx(mid_day, y, xout = mid_day)$y
      if (v == rain) f <- expm1(f)
      hole <- is.na(monthly[p, , v])
      monthly[p, hole, v] <- f[hole]
    }
  }
  monthly
}

# dates exactly as daily_from_monthly_climate() derives them (R/sim_microclimate.R:12-31) and periods
# exactly as aggregate_vectors() groups them (R/pipeline_functions.R:1050-1088); `first` is any pixel's
# monthly tibble (all pixels of a tile share start / end)
mms_chain_setup <- function(first, included_dates, aggregate_by, on_extrapolation = "error") {
  midpoint <- first$start + round((first$end - first$start) / 2)
  dates_predict <- seq(from = min(midpoint), to = max(midpoint), by = 1)
  mid_date <- first$start + (first$end - first$start) / 2
  outside <- dates_predict < min(mid_date) | dates_predict > max(mid_date)     # R/spline_interpolation.R:30-31
  if (any(outside) && on_extrapolation == "clip") {
    warning(sum(outside), " prediction day(s) outside the monthly mid-dates dropped ",
            "(the reference would extrapolate them with an mgcv::gam)", call. = FALSE)
    dates_predict <- dates_predict[!outside]
  }
  if (is.null(included_dates)) included_dates <- dates_predict
  keep <- dates_predict %in% included_dates
  key <- if (aggregate_by == "month") lubridate::floor_date(dates_predict, "month") else dates_predict
  periods <- sort(unique(key[keep]))
  list(mid_day = as.numeric(mid_date - min(mid_date)),
       day_predict = as.numeric(dates_predict - min(mid_date)),
       period_of_day = as.integer(ifelse(keep, match(key, periods) - 1L, -1L)),
       periods = periods, dates_predict = dates_predict)
}

# the summarise_vectors() schema from the library's [column, period, variable] arrays; simplify = TRUE returns the
# arrays themselves (list(start, end, adult = [columns, periods], ...)) -- at a million pixels building a tibble per
# pixel is slower than the simulation
mms_vectors_result <- function(res, pixels, species, periods, aggregate_by, simplify = FALSE) {
  n_periods <- length(periods)
  n_cols <- length(res$mean_state) / (n_periods * 2)                     # pixels, or raster cells
  state <- array(res$mean_state, dim = c(n_cols, n_periods, 2))          # adult, aquatic
  forcing <- array(res$mean_forcing, dim = c(n_cols, n_periods, 4))      # mdr, efd, das_zerodensity, ds
  end <- if (aggregate_by == "month") lubridate::ceiling_date(periods + 1, "month") - 1 else periods
  if (simplify) {
    return(list(start = periods, end = end, species = species,
                adult = state[, , 1], aquatic = state[, , 2], mdr = forcing[, , 1], efd = forcing[, , 2],
                das_zerodensity = forcing[, , 3], ds = forcing[, , 4]))
  }
  n_pixels <- nrow(pixels)
  dplyr::tibble(
    longitude = pixels$longitude,
    latitude = pixels$latitude,
    species = species,
    pixel_vectors = lapply(seq_len(n_pixels), function(i) {
      dplyr::tibble(start = periods, end = end,
                    adult = state[i, , 1], aquatic = state[i, , 2],
                    mdr = forcing[i, , 1], efd = forcing[i, , 2],
                    das_zerodensity = forcing[i, , 3], ds = forcing[i, , 4])
    })
  )
}

# tc_to_vectors() whole (R/pipeline_functions.R:1099-1130): the raw TerraClimate tibbles of
# extract_terraclimate_tile() go to the device, where process_terraclimate_pixel_vars()
# (R/terraclimate.R:581-643) runs in front of the fused chain.  Altitude and the clear-sky lookup are
# read on the R side exactly as terraclimate_to_monthly_climate() reads them (:183-233).
tc_to_vectors_gpu <- function(pixel_terraclimate_data,
                              species = c("An. gambiae", "An. stephensi"),
                              included_dates = NULL,
                              aggregate_by = c("month", "day"),
                              model_water_temperature = FALSE,
                              water_shade_proportion = 1,
                              on_extrapolation = c("error", "clip"),
                              cell_idx = NULL, n_raster_cells = NULL,
                              simplify = !is.null(cell_idx)) {
  species <- match.arg(species)
  aggregate_by <- match.arg(aggregate_by)
  lifehistory <- switch(species,
                        "An. gambiae" = lifehistory_functions$An_gambiae,
                        "An. stephensi" = lifehistory_functions$An_stephensi)
  tc_elev <- terra::unwrap(tc_elev_wrapped)
  tc_gads <- terra::unwrap(tc_gads_wrapped)
  coords <- pixel_terraclimate_data[, c("longitude", "latitude")]
  altitude <- altitude_m(longitude = coords$longitude, latitude = coords$latitude,
                         altitude_raster = tc_elev)
  cell_id <- terra::extract(tc_gads, coords, ID = FALSE)[, 1]
  # [pixels, 12]: one lookup per DISTINCT GADS cell, not per pixel
  lookup_rows <- match(cell_id, clear_sky_lookup$cell_id)
  solr <- do.call(rbind, lapply(clear_sky_lookup$clear_sky_SOLR, as.numeric))
  clear_sky <- solr[lookup_rows, , drop = FALSE]
  storage.mode(clear_sky) <- "double"
  first <- pixel_terraclimate_data$monthly_terraclimate[[1]]
  vars <- c("tmax", "tmin", "ppt", "ws", "vpd", "srad")
  values <- mms_pack_pixels(pixel_terraclimate_data$monthly_terraclimate, vars)
  terraclimate <- list(
    values = values, clear_sky = clear_sky,
    month_id = as.integer(lubridate::month(first$start + (first$end - first$start) / 2)),  # R/terraclimate.R:591
    n_days = as.numeric(first$end - first$start))                                          # :592
  chain <- mms_chain_setup(first, included_dates, aggregate_by, match.arg(on_extrapolation))
  opts <- mms_chain_options(chain, pixel_terraclimate_data, model_water_temperature, water_shade_proportion,
                            clear_sky, cell_idx, n_raster_cells)
  res <- .Call("mms_R_run_chain_terraclimate", terraclimate, as.numeric(altitude),
               chain$mid_day, chain$day_predict, chain$period_of_day,
               list(mms_habitat()), mms_densmod(lifehistory$das_densmod),
               mms_lifehistory_tables(lifehistory), 63L, length(chain$periods), 1L,
               mms_math_mode(), mms_devices(), opts, PACKAGE = "mosmicrosim")
  mms_vectors_result(res, pixel_terraclimate_data, species, chain$periods, aggregate_by, simplify)
}

# ---- the switch: one line each in the reference's own functions ------------------
# R/sim_water.R:216     if (mms_backend() == "gpu") return(simulate_ephemeral_habitat_vectorised_gpu(...))
# R/sim_pop_dynamics.R:208  if (mms_backend() == "gpu") return(simulate_population_vectorised_gpu(...))
# R/sim_water.R:136     if (mms_backend() == "gpu") return(simulate_ephemeral_habitat_gpu(...))
# R/sim_pop_dynamics.R:138  if (mms_backend() == "gpu") return(simulate_population_gpu(...))
# R/pipeline_functions.R:1099 (tc_to_vectors):
#   if (mms_backend() == "gpu") return(tc_to_vectors_gpu(pixel_terraclimate_data, species, included_dates))
# or, keeping terraclimate_to_monthly_climate() on the R side:
#   monthly <- terraclimate_to_monthly_climate(pixel_terraclimate_data)
#   if (mms_backend() == "gpu") return(monthly_climate_to_vectors_gpu(monthly, species, included_dates))
