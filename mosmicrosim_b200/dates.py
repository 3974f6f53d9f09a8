"""Date bookkeeping of the reference pipeline, done once on the host.

Mirrors the Date arithmetic in daily_from_monthly_climate()
(R/sim_microclimate.R:12-31), spline_seasonal() (R/spline_interpolation.R:10-21)
and aggregate_vectors() (R/pipeline_functions.R:1050-1088).  Dates are numpy
datetime64[D]; R Dates are day counts too, so the numbers are identical.
"""
from __future__ import annotations

import numpy as np


def _as_days(dates) -> np.ndarray:
    return np.asarray(dates, dtype="datetime64[D]")


def r_round_half_even(x: np.ndarray) -> np.ndarray:
    """base-R round(x) for the x.5 values that occur here (IEC 60559 half-even)."""
    return np.rint(x)


def month_mid_days(start, end):
    """mid_date = start + (end - start) / 2 as (possibly x.5) day numbers since
    the Unix epoch (R/sim_microclimate.R:31)."""
    s = _as_days(start).astype(np.int64).astype(np.float64)
    e = _as_days(end).astype(np.int64).astype(np.float64)
    return s + (e - s) / 2


def default_dates_predict(start, end) -> np.ndarray:
    """dates_predict = seq(min(midpoint), max(midpoint), by = 1) with
    midpoint = start + round((end - start) / 2)  (R/sim_microclimate.R:14-24)."""
    s = _as_days(start).astype(np.int64)
    e = _as_days(end).astype(np.int64)
    midpoint = s + r_round_half_even((e - s) / 2).astype(np.int64)
    return np.arange(midpoint.min(), midpoint.max() + 1).astype("datetime64[D]")


def datenums(start, end, dates_predict):
    """(mid_day[M], day_predict[D]): as.numeric(dates - min(dates)) for the
    monthly knots and for the prediction dates (R/spline_interpolation.R:10-21)."""
    mid = month_mid_days(start, end)
    min_date = mid.min()
    dp = _as_days(dates_predict).astype(np.int64).astype(np.float64)
    return mid - min_date, dp - min_date


def needs_extrapolation(mid_day: np.ndarray, day_predict: np.ndarray) -> bool:
    """any(dates_predict > max(dates) | dates_predict < min(dates)),
    R/spline_interpolation.R:30-34 -- the branch that fits an mgcv::gam."""
    return bool(np.any((day_predict > mid_day.max()) | (day_predict < mid_day.min())))


def period_index(dates_predict, aggregate_by="month", included_dates=None):
    """Map every simulated day to a summary period, as aggregate_vectors() does
    (R/pipeline_functions.R:1030-1091): filter(date %in% included_dates), then
    group by date ("day") or by floor_date(date, "month") ("month"; "year" is an
    extension).  Returns (period_of_day int32[D] with -1 = excluded,
    period_start datetime64[D][P], period_end datetime64[D][P])."""
    if aggregate_by not in ("month", "day", "year"):
        raise ValueError("'arg' should be one of 'month', 'day'")
    d = _as_days(dates_predict)
    keep = np.ones(len(d), dtype=bool)
    if included_dates is not None:
        keep = np.isin(d, _as_days(included_dates))
    if aggregate_by == "day":
        key = d
        unit = "D"
    elif aggregate_by == "month":
        key = d.astype("datetime64[M]")
        unit = "M"
    else:
        key = d.astype("datetime64[Y]")
        unit = "Y"
    uniq = np.unique(key[keep])
    lookup = {k: i for i, k in enumerate(uniq)}
    pod = np.full(len(d), -1, dtype=np.int32)
    for i in np.nonzero(keep)[0]:
        pod[i] = lookup[key[i]]
    start = uniq.astype("datetime64[D]")
    end = (uniq + np.timedelta64(1, unit)).astype("datetime64[D]") - np.timedelta64(1, "D")
    return pod, start, end
