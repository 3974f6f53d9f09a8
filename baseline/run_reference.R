# run_reference.R -- times the REFERENCE's own R functions on synthetic inputs, cell-parallel with
# parallel::mclapply, for a side-by-side CPU number (BASELINE.md section 4, item 2).
#
# NOT EXECUTED in the development image of this repository (R is not installed there; bench.py's
# cpu_baseline uses the C restatement in oracle/ instead).  Needs R with the reference installed
# (remotes::install_github("goldingn/mosmicrosim")) and its sysdata; usage:
#   Rscript baseline/run_reference.R [n_cells = 64] [cores = parallel::detectCores()]
args <- commandArgs(trailingOnly = TRUE)
n_cells <- if (length(args) >= 1) as.integer(args[1]) else 64L
cores <- if (length(args) >= 2) as.integer(args[2]) else parallel::detectCores()
set.seed(20251218)

# monthly knots December .. January around one simulated year (no extrapolation branch)
start <- seq(as.Date("2000-12-01"), by = "1 month", length.out = 14)
end <- seq(as.Date("2001-01-01"), by = "1 month", length.out = 14) - 1
moy <- as.integer(format(start, "%m")) - 1

make_cell <- function(i) {
  phase <- runif(1, 0, 2 * pi)
  season <- sin(2 * pi * moy / 12 + phase)
  tmax <- 30 + 5 * season + rnorm(14)
  rhmax <- runif(14, 60, 100)
  wsmax <- runif(14, 0.5, 5)
  cloud <- runif(14, 0, 80)
  dplyr::tibble(
    start = start, end = end,
    tmin = tmax - runif(14, 8, 14), tmax = tmax,
    rhmin = pmax(0.01, rhmax * runif(14, 0.3, 0.8)), rhmax = rhmax,
    wsmin = 0.1 * wsmax, wsmax = wsmax,
    ccmin = pmin(100, 0.8 * cloud), ccmax = pmin(100, 1.2 * cloud),
    rainfall_daily = rgamma(14, shape = 2, scale = 2) * pmax(0, 0.5 + 0.8 * sin(2 * pi * moy / 12 + phase + 1))
  )
}

one_cell <- function(i) {
  monthly <- make_cell(i)
  altitude <- runif(1, 0, 2500)
  daily <- mosmicrosim:::daily_from_monthly_climate(monthly, dates_predict = seq(as.Date("2001-01-01"), as.Date("2001-12-31"), by = 1))
  hourly <- mosmicrosim:::hourly_from_daily_climate_ambient(daily)
  m <- function(x) matrix(x, ncol = 1)
  water <- mosmicrosim:::simulate_ephemeral_habitat_vectorised(
    m(hourly$rainfall), m(hourly$air_temperature), m(hourly$humidity), m(hourly$windspeed), altitude)
  lh <- mosmicrosim:::lifehistory_functions$An_gambiae
  lh$ds_temp_humid <- mosmicrosim:::make_temp_humid_interpolator(lh$ds_temp_humid)
  pop <- mosmicrosim:::simulate_population_vectorised(
    water$water_surface_area, m(lh$mdr_temp(hourly$water_temperature)), m(lh$efd_temp(hourly$air_temperature)),
    m(lh$das_temp(hourly$water_temperature)), m(lh$ds_temp_humid(hourly$air_temperature, hourly$humidity)),
    lh$das_densmod)
  mean(pop$adult)
}

t0 <- Sys.time()
res <- parallel::mclapply(seq_len(n_cells), one_cell, mc.cores = cores)
secs <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
cat(sprintf('{"impl": "reference-R", "cells": %d, "cores": %d, "seconds": %.3f, "cell_years_per_s": %.3f}\n',
            n_cells, cores, secs, n_cells / secs))
