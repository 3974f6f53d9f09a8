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


def clip_dates_predict(start, end, dates_predict):
    """Drop the prediction days that lie outside the monthly mid-dates.  The reference's default dates_predict
    (midpoints rounded half to even) can overhang the knots (unrounded midpoints) by half a day -- a 30-day first
    month, a non-leap February last month (R/sim_microclimate.R:14-24,31) -- and the reference then fits an
    mgcv::gam for the whole series (R/spline_interpolation.R:30-112), which the GPU backend does not reproduce.
    Returns (clipped dates, number of days dropped)."""
    mid = month_mid_days(start, end)
    dp = np.asarray(dates_predict, dtype="datetime64[D]")
    x = _as_days(dp).astype(np.int64).astype(np.float64)
    keep = (x >= mid.min()) & (x <= mid.max())
    return dp[keep], int((~keep).sum())


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


# ---- TerraClimate month bookkeeping --------------------------------------

def day_of_year(dates) -> np.ndarray:
    """doy(): as.integer(format(date, "%j")), R/utils.R:46-48 (1-based)."""
    d = np.asarray(dates, dtype="datetime64[D]")
    return ((d - d.astype("datetime64[Y]")).astype(np.int64) + 1).astype(np.int32)


def month_of_year(dates) -> np.ndarray:
    """lubridate::month(date): 1..12."""
    d = np.asarray(dates, dtype="datetime64[D]")
    return (d.astype("datetime64[M]").astype(np.int64) % 12 + 1).astype(np.int32)


def terraclimate_month_info(start, end):
    """(month_id int32[M], n_days float64[M]) as process_terraclimate_pixel_vars() computes them
    (R/terraclimate.R:591-592): month_id = lubridate::month(start + (end - start) / 2) and
    n_days = as.numeric(end - start), i.e. the days in the month MINUS ONE (replicated, not fixed)."""
    s = _as_days(start).astype(np.int64)
    e = _as_days(end).astype(np.int64)
    mid = (s + np.floor((e - s) / 2).astype(np.int64)).astype("datetime64[D]")   # month() of an x.5 date floors
    month_id = (mid.astype("datetime64[M]").astype(np.int64) % 12 + 1).astype(np.int32)
    return month_id, (e - s).astype(np.float64)


def r_mean(x) -> float:
    """base-R mean() of a double vector (long double sum, then one correction pass)."""
    x = np.asarray(x, dtype=np.longdouble)
    if x.size == 0:
        return float("nan")
    s = x.sum() / x.size
    s += (x - s).sum() / x.size
    return float(s)


def pad_last_month(start, end, values, n_previous=3):
    """pad_last_month(), R/terraclimate.R:519-569: append a synthetic next month whose values are the
    means of that calendar month in the previous n_previous years of the series (those present).
    values: [..., M, ...]-shaped arrays are not needed here -- pass [V][M] or [V][M][C] with the month
    axis second.  Returns (start[M+1], end[M+1], values[V][M+1](...))."""
    s, e = _as_days(start), _as_days(end)
    values = np.asarray(values, dtype=np.float64)
    last_start = s.max()
    next_start = (last_start.astype("datetime64[M]") + 1).astype("datetime64[D]")        # :525-528
    next_end = (next_start.astype("datetime64[M]") + 1).astype("datetime64[D]") - np.timedelta64(1, "D")  # :531-534
    previous = [(next_start.astype("datetime64[M]") - 12 * k).astype("datetime64[D]")
                for k in range(1, n_previous + 1)]                                        # :537-540
    rows = np.flatnonzero(np.isin(s, previous))                                           # :544-546
    sub = values[:, rows]                                                                 # [V][k](...)
    n_cols = int(np.prod(values.shape[2:], dtype=np.int64))
    flat = sub.reshape(sub.shape[0], sub.shape[1], n_cols)
    if len(rows):          # R mean(): long double sum / n, then one correction pass (k <= n_previous terms)
        ld = flat.astype(np.longdouble)
        m = ld.sum(axis=1) / len(rows)
        m += (ld - m[:, None, :]).sum(axis=1) / len(rows)
        pad = m.astype(np.float64)
    else:
        pad = np.full((flat.shape[0], n_cols), np.nan)
    pad = pad.reshape((values.shape[0], 1) + values.shape[2:])                            # :551-556
    return (np.append(s, next_start), np.append(e, next_end), np.concatenate([values, pad], axis=1))
