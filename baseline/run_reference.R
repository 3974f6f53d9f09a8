# run_reference.R -- runs the REFERENCE's own R functions (goldingn/mosmicrosim) on THIS repository's synthetic
# inputs and writes every stage's output as flat little-endian binaries, so that anyone with R can pin the CPU
# oracle (oracle/) and the CUDA path (libmmsgpu) to the real reference:
#
#   python tests/golden/make_r_inputs.py                              # inputs (already committed)
#   Rscript baseline/run_reference.R --source /path/to/mosmicrosim    # sources R/*.R of a checkout (no sysdata needed)
#   #   or: Rscript baseline/run_reference.R                          # uses the installed package (mosmicrosim:::)
#   python -m pytest tests/test_r_reference.py                        # oracle vs R (CPU) and GPU vs R (-m gpu)
#
# NOT EXECUTED in the development image of this repository: R is not installed there (nor on its GPU boxes), which is
# why parity is "unpinned" (DESIGN.md section 2).  Needs R >= 4.1 with dplyr, tidyr, lubridate and fields.
# A second mode times the same functions cell-parallel with parallel::mclapply for the CPU baseline:
#   Rscript baseline/run_reference.R --source DIR --time [cores]
#
# Stages driven (all internal functions of the reference; citations relative to its root):
#   daily_from_monthly_climate()             R/sim_microclimate.R:7-64      -> daily.f64
#   hourly_from_daily_climate_ambient()      R/sim_microclimate.R:126-173   -> hourly.f64
#   simulate_ephemeral_habitat_vectorised()  R/sim_water.R:216-301          -> water_surface_area.f64
#   positive_spline() closures, make_temp_humid_interpolator(), R/lifehistory_estimation_functions.R:236-250,1144-1190
#                                                                           -> mdr / efd / das_zerodensity / ds .f64
#   make_surv_densmod_function() + simulate_population_vectorised()  R/sim_pop_dynamics.R:208-284 -> aquatic / adult .f64
#   aggregate_vectors(aggregate_by = "month")  R/pipeline_functions.R:1030-1091 -> monthly_means.f64
#
# File layouts (what tests/test_r_reference.py reads): every array is written with writeBin(as.vector(x)), i.e. R's
# column-major order -- the FIRST index varies fastest:
#   daily.f64               dim c(D, 9, C)   variables in the order of `forcing_vars` below
#   hourly.f64              dim c(H, 6, C)   air_temperature, humidity, cloudcover, windspeed, water_temperature, rainfall
#   water_surface_area.f64, mdr.f64, efd.f64, das_zerodensity.f64, ds.f64, aquatic.f64, adult.f64   dim c(H, C)
#   monthly_means.f64       dim c(P, 6, C)   adult, aquatic, mdr, efd, das_zerodensity, ds (the reference's vector_variables)
# for each species directory (An_gambiae: logit density dependence, An_stephensi: cox_ph).

args <- commandArgs(trailingOnly = TRUE)
arg_value <- function(flag, default = NULL) {
  k <- match(flag, args)
  if (is.na(k)) return(default)
  if (k < length(args) && !startsWith(args[k + 1], "--")) args[k + 1] else TRUE
}
script_dir <- dirname(sub("^--file=", "", grep("^--file=", commandArgs(FALSE), value = TRUE)[1]))
repo <- normalizePath(file.path(script_dir, ".."))
in_dir <- arg_value("--inputs", file.path(repo, "tests", "golden", "r_reference", "inputs"))
out_dir <- arg_value("--outputs", file.path(repo, "tests", "golden", "r_reference", "outputs"))
source_dir <- arg_value("--source")
time_mode <- !is.null(arg_value("--time"))

# ---- the reference's functions: from a source checkout, or from the installed package -------------------------
suppressPackageStartupMessages({ library(dplyr); library(tidyr) })
if (!is.null(source_dir)) {
  ref <- new.env()
  for (f in c("utils.R", "spline_interpolation.R", "sim_microclimate.R", "sim_water.R", "sim_pop_dynamics.R",
              "lifehistory_estimation_functions.R", "pipeline_functions.R")) {
    sys.source(file.path(source_dir, "R", f), envir = ref)
  }
} else {
  ref <- asNamespace("mosmicrosim")
}

rd <- function(dir, name, n) readBin(file.path(dir, name), "double", n = n, size = 8, endian = "little")
ri <- function(dir, name, n) readBin(file.path(dir, name), "integer", n = n, size = 4, endian = "little")
wr <- function(dir, name, x) writeBin(as.vector(x), file.path(dir, name), size = 8, endian = "little")

dims <- scan(file.path(in_dir, "dims.txt"), quiet = TRUE)      # C M D n_knots n_grid
C <- dims[1]; M <- dims[2]; D <- dims[3]; n_knots <- dims[4]; n_grid <- dims[5]
H <- 24 * D
as_date <- function(days) as.Date(days, origin = "1970-01-01")
start <- as_date(ri(in_dir, "start_days.i32", M))
end <- as_date(ri(in_dir, "end_days.i32", M))
dates_predict <- as_date(ri(in_dir, "dates_predict_days.i32", D))
# monthly.f64 is [9][M][C] with the cell index fastest == an R array of dim c(C, M, 9)
monthly <- array(rd(in_dir, "monthly.f64", 9 * M * C), dim = c(C, M, 9))
altitude <- rd(in_dir, "altitude.f64", C)
forcing_vars <- c("tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax", "rainfall_daily")
hourly_vars <- c("air_temperature", "humidity", "cloudcover", "windspeed", "water_temperature", "rainfall")
vector_vars <- c("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds")

monthly_tibble <- function(cell) {
  cols <- lapply(seq_along(forcing_vars), function(v) monthly[cell, , v])
  names(cols) <- forcing_vars
  dplyr::bind_cols(dplyr::tibble(start = start, end = end), dplyr::as_tibble(cols))
}

# the life-history closures of one species, built by the reference's own constructors from this repository's tables
lifehistory <- function(species_dir) {
  temps_out <- rd(species_dir, "temps_out.f64", n_knots)
  ds_z <- rd(species_dir, "ds_z.f64", n_grid * n_grid)
  dm <- scan(file.path(species_dir, "densmod.txt"), what = list(type = "", dd_effect = 0, surface_area_cm2 = 0), quiet = TRUE)
  list(
    mdr_temp = ref$positive_spline(rd(species_dir, "mdr_y.f64", n_knots), temps_out),
    efd_temp = ref$positive_spline(rd(species_dir, "efd_y.f64", n_knots), temps_out),
    das_temp = ref$positive_spline(rd(species_dir, "das_y.f64", n_knots), temps_out),
    # make_temp_humid_interpolator() calls its argument ONCE, on expand.grid(temperature, humidity) (temperature
    # fastest): handing back the table in that order reproduces the repository's surface exactly, on R's own grid axes
    ds_temp_humid = ref$make_temp_humid_interpolator(function(temperature, humidity) ds_z),
    das_densmod = ref$make_surv_densmod_function(dm$dd_effect, dm$surface_area_cm2, dm$type)
  )
}

# everything for one block of cells: returns the stage outputs as arrays
run_cells <- function(cells, lh) {
  n <- length(cells)
  daily_out <- array(NA_real_, dim = c(D, 9, n))
  hourly_out <- array(NA_real_, dim = c(H, 6, n))
  for (k in seq_len(n)) {
    daily <- ref$daily_from_monthly_climate(monthly_tibble(cells[k]), dates_predict = dates_predict)
    hourly <- ref$hourly_from_daily_climate_ambient(daily)
    for (v in seq_along(forcing_vars)) daily_out[, v, k] <- daily[[forcing_vars[v]]]
    for (v in seq_along(hourly_vars)) hourly_out[, v, k] <- hourly[[hourly_vars[v]]]
  }
  mat <- function(name) matrix(hourly_out[, match(name, hourly_vars), ], nrow = H, ncol = n)
  air_t <- mat("air_temperature"); water_t <- mat("water_temperature"); rh <- mat("humidity")
  water <- ref$simulate_ephemeral_habitat_vectorised(
    rainfall_matrix = mat("rainfall"), air_temperature_matrix = air_t, humidity_matrix = rh,
    windspeed_matrix = mat("windspeed"), altitude_vector = altitude[cells])
  # simulate_vector_lifehistory(), R/sim_pop_dynamics.R:100-135: mdr and das take the water temperature, efd the air temperature
  mdr <- matrix(lh$mdr_temp(as.vector(water_t)), H, n)
  efd <- matrix(lh$efd_temp(as.vector(air_t)), H, n)
  das <- matrix(lh$das_temp(as.vector(water_t)), H, n)
  ds <- matrix(lh$ds_temp_humid(as.vector(air_t), as.vector(rh)), H, n)
  pop <- ref$simulate_population_vectorised(
    larval_habitat_area_matrix = water$water_surface_area, mdr_matrix = mdr, efd_matrix = efd,
    das_zerodensity_matrix = das, ds_matrix = ds, das_densmod_function = lh$das_densmod)
  hour_dates <- rep(dates_predict, each = 24)
  means <- NULL
  for (k in seq_len(n)) {
    agg <- ref$aggregate_vectors(
      dplyr::tibble(date = hour_dates, adult = pop$adult[, k], aquatic = pop$aquatic[, k], mdr = mdr[, k],
                    efd = efd[, k], das_zerodensity = das[, k], ds = ds[, k]),
      aggregate_by = "month")
    if (is.null(means)) means <- array(NA_real_, dim = c(nrow(agg), 6, n))
    for (v in seq_along(vector_vars)) means[, v, k] <- agg[[vector_vars[v]]]
  }
  list(daily = daily_out, hourly = hourly_out, water_surface_area = water$water_surface_area, mdr = mdr, efd = efd,
       das_zerodensity = das, ds = ds, aquatic = pop$aquatic, adult = pop$adult, monthly_means = means)
}

species <- c("An_gambiae", "An_stephensi")
if (time_mode) {
  # the CPU baseline the north star names: the reference parallelised across cells with parallel::mclapply
  cores <- suppressWarnings(as.integer(arg_value("--time")))
  if (is.na(cores)) cores <- parallel::detectCores()
  lh <- lifehistory(file.path(in_dir, species[1]))
  t0 <- Sys.time()
  res <- parallel::mclapply(seq_len(C), function(cell) mean(run_cells(cell, lh)$adult), mc.cores = cores)
  secs <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
  cat(sprintf('{"impl": "reference-R", "cells": %d, "years": %.3f, "cores": %d, "seconds": %.3f, "cell_years_per_s": %.3f}\n',
              C, H / 8760, cores, secs, C * (H / 8760) / secs))
} else {
  for (sp in species) {
    lh <- lifehistory(file.path(in_dir, sp))
    res <- run_cells(seq_len(C), lh)
    dir.create(file.path(out_dir, sp), recursive = TRUE, showWarnings = FALSE)
    for (name in names(res)) wr(file.path(out_dir, sp), paste0(name, ".f64"), res[[name]])
    writeLines(sprintf("%d %d %d %d", C, D, H, dim(res$monthly_means)[1]), file.path(out_dir, sp, "dims.txt"))
    cat("wrote", file.path(out_dir, sp), "\n")
  }
  writeLines(c(paste("R", R.version.string), paste("source", ifelse(is.null(source_dir), "installed package", source_dir)),
               paste("fields", as.character(utils::packageVersion("fields")))), file.path(out_dir, "provenance.txt"))
}
