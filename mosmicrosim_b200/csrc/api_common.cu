// api_common.cu -- shared host-side pieces of libmmsgpu (see api_common.h) and the library-level entry points
// of include/mmsgpu.h.  Citations are relative to the reference root.
#include "api_common.h"

namespace mms_api {

namespace {
thread_local std::string g_error;
thread_local float g_stage_kernel_ms = 0.f;
}  // namespace

std::string &last_error() { return g_error; }
float &stage_kernel_ms() { return g_stage_kernel_ms; }

int fail(int status, const char *fmt, ...)
{
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return status;
}

int fail_str(int status, const std::string &msg)
{
    g_error = msg;
    return status;
}

int device_count_raw()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pick_device(int32_t device)
{
    const int n = device_count_raw();
    if (n < 1) return fail(MMS_ERR_NO_DEVICE, "no CUDA device is visible; libmmsgpu has no CPU fallback");
    if (device < 0 || device >= n) return fail(MMS_ERR_INVALID_ARG, "device %d is not visible (%d devices)", device, n);
    MMS_CUDA(cudaSetDevice(device));
    return MMS_OK;
}

bool host_pointer_is_pinned(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void fill_hour_tables(mms::HourTables &hr)
{
    const double time_sunrise = 6 * 100.0, time_sunset = 18 * 100.0, time_daily_max = 13 * 100.0;
    const double time_reference = (time_daily_max - time_sunrise) / 2 + time_sunrise;          // :204
    const double tau = 3 / (2400 - time_sunset + time_sunrise);                                // :207
    hr.z_night = std::sin((360 * (time_sunset - time_reference) / (2 * (time_daily_max - time_sunrise))) / 57.29577);
    for (int h = 0; h < 24; ++h) {
        const double times = 100.0 * (h + 1);
        hr.z_light[h] = std::sin((360 * (times - time_reference) / (2 * (time_daily_max - time_sunrise))) / 57.29577);
        const double time_since_ss = times - time_sunset + 2400 * ((times < 1800) ? 1.0 : 0.0);  // :260
        hr.exp_e[h] = std::exp(tau * time_since_ss);                                             // :261-262
        hr.predawn[h] = 1 - times / time_sunrise;                                                // :275
        // vsine_weights(times = h + 1, time_sunrise = 6, time_inflection = 13)
        const double t = h + 1.0, sr = 6.0, inf = 13.0;
        const double m1 = (t <= sr) ? 1.0 : 0.0;
        const double w1 = (1 - 0.5 * std::max(0.0, (sr - t) / sr)) * m1;
        const double m2 = (t > sr && t <= inf) ? 1.0 : 0.0;
        const double w2 = (1 - std::min(1.0, std::max(0.0, (t - sr) / (inf - sr)))) * m2;
        const double m3 = (t > inf) ? 1.0 : 0.0;
        const double w3 = 0.5 * std::min(1.0, std::max(0.0, (t - inf) / (24 - inf))) * m3;
        hr.w_vsine[h] = w1 + w2 + w3;                                                            // :451
        hr.w_vsine_inv[h] = 1 - hr.w_vsine[h];                                                   // :498-500
        hr.air_coef[h] = (h < 5) ? hr.predawn[h] : ((h <= 17) ? hr.z_light[h] : 1.0 / hr.exp_e[h]);
    }
}

double r_pow(double x, double y)
{
    if (y == 2.0) return x * x;
    if (x == 1. || y == 0.) return 1.;
    if (x == 0.) return (y > 0.) ? 0. : INFINITY;
    return std::pow(x, y);
}

int derive_habitat(const mms_habitat &in, mms::HabitatDev &out)
{
    if (in.shape != 0 && in.shape != 1) return fail(MMS_ERR_INVALID_ARG, "habitat shape must be 0 (cone) or 1 (cylinder)");
    if (!(in.max_depth > 0)) return fail(MMS_ERR_INVALID_ARG, "habitat max_depth must be positive");
    if (in.shape == 1 && !(in.radius > 0)) return fail(MMS_ERR_INVALID_ARG, "cylinder radius must be positive");
    if (!(in.leak_per_hour >= 0) || !(in.leak_per_hour <= 1)) return fail(MMS_ERR_INVALID_ARG, "leak_per_hour must be in [0, 1]");
    const double pi = 3.141592653589793;
    out.shape = in.shape;
    out.pad = 0;
    out.leak_per_hour = in.leak_per_hour;
    out.initial_volume = in.initial_volume;
    if (in.shape == 0) {
        const double radius = in.max_depth * std::sin(pi / 2);                      // cone_depth_to_radius
        out.max_volume = pi * r_pow(radius, 2) * in.max_depth / 3;                  // cone_depth_to_volume
        out.cyl_area = 0.0;
        out.catchment_area = in.inflow_multiplier * (3.046474 * r_pow(out.max_volume, 0.666667));
    } else {
        out.cyl_area = pi * (in.radius * in.radius);
        out.max_volume = out.cyl_area * in.max_depth;
        out.catchment_area = in.inflow_multiplier * out.cyl_area;
    }
    return MMS_OK;
}

int build_day_table(const double *x, int M, const double *v, int D, std::vector<mms::DayEntry> &days)
{
    for (int m = 1; m < M; ++m)
        if (!(x[m] > x[m - 1])) return fail(MMS_ERR_INVALID_ARG, "mid_day must be strictly increasing");
    days.resize(D);
    for (int d = 0; d < D; ++d) {
        const double vd = v[d];
        if (!(vd >= x[0]) || !(vd <= x[M - 1]))
            return fail(MMS_ERR_EXTRAPOLATION,
                        "prediction day %d (%.1f) lies outside the monthly mid-dates [%.1f, %.1f]: the reference "
                        "would fit an mgcv::gam here (R/spline_interpolation.R:30-112), which the GPU backend "
                        "does not reproduce; add buffer months or shorten dates_predict", d, vd, x[0], x[M - 1]);
        int i = 0, j = M - 1;
        while (i < j - 1) {
            const int ij = (i + j) / 2;
            if (vd < x[ij]) j = ij; else i = ij;
        }
        mms::DayEntry e;
        e.idx = i;
        e.flag = (vd == x[j]) ? 1 : ((vd == x[i]) ? 2 : 0);
        e.w = (vd - x[i]) / (x[j] - x[i]);
        days[d] = e;
    }
    return MMS_OK;
}

int pack_tables(const mms_lifehistory &lh, HostTables &t)
{
    const mms_spline *sp[3] = {&lh.mdr_temp, &lh.efd_temp, &lh.das_temp};
    const int n = sp[0]->n;
    if (n < 2) return fail(MMS_ERR_INVALID_ARG, "spline tables need at least 2 knots");
    for (int f = 0; f < 3; ++f) {
        if (!sp[f]->x || !sp[f]->y || !sp[f]->b || !sp[f]->c || !sp[f]->d)
            return fail(MMS_ERR_INVALID_ARG, "spline table %d has NULL arrays", f);
        if (sp[f]->n != n || std::memcmp(sp[f]->x, sp[0]->x, sizeof(double) * n) != 0)
            return fail(MMS_ERR_UNSUPPORTED,
                        "the mdr/efd/das spline tables must share one knot vector (the reference builds all "
                        "three on temps_out = seq(-20, 80, by = 0.1))");
    }
    if ((size_t)n * 13 * sizeof(double) > 200 * 1024)
        return fail(MMS_ERR_UNSUPPORTED, "spline tables with %d knots do not fit in shared memory", n);
    t.sx.assign(sp[0]->x, sp[0]->x + n);
    for (int k = 1; k < n; ++k)
        if (!(t.sx[k] > t.sx[k - 1])) return fail(MMS_ERR_INVALID_ARG, "spline knots must be strictly increasing");
    t.scoef.resize((size_t)n * 12);
    for (int k = 0; k < n; ++k)
        for (int f = 0; f < 3; ++f) {
            double *o = &t.scoef[((size_t)k * 3 + f) * 4];
            o[0] = sp[f]->y[k]; o[1] = sp[f]->b[k]; o[2] = sp[f]->c[k]; o[3] = sp[f]->d[k];
        }
    t.x0 = t.sx[0];
    t.step = (t.sx[n - 1] - t.sx[0]) / (n - 1);
    t.inv_step = 1.0 / t.step;
    // equally spaced to within 1e-9 of a step?  (the reference's knots are seq(-20, 80, by = 0.1))
    t.uniform = 1;
    for (int k = 0; k < n; ++k)
        if (std::fabs(t.sx[k] - (t.x0 + k * t.step)) > 1e-9 * t.step) t.uniform = 0;
    const mms_surface &g = lh.ds_temp_humid;
    if (g.nx < 2 || g.ny < 2 || !g.x || !g.y || !g.z) return fail(MMS_ERR_INVALID_ARG, "ds surface needs a grid of at least 2 x 2");
    {
        const double dx = (g.x[g.nx - 1] - g.x[0]) / (g.nx - 1), dy = (g.y[g.ny - 1] - g.y[0]) / (g.ny - 1);
        for (int k = 0; k < g.nx; ++k) if (std::fabs(g.x[k] - (g.x[0] + k * dx)) > 1e-9 * dx) t.uniform = 0;
        for (int k = 0; k < g.ny; ++k) if (std::fabs(g.y[k] - (g.y[0] + k * dy)) > 1e-9 * dy) t.uniform = 0;
    }
    for (int k = 1; k < g.nx; ++k) if (!(g.x[k] > g.x[k - 1])) return fail(MMS_ERR_INVALID_ARG, "ds surface x must be strictly increasing");
    for (int k = 1; k < g.ny; ++k) if (!(g.y[k] > g.y[k - 1])) return fail(MMS_ERR_INVALID_ARG, "ds surface y must be strictly increasing");
    t.quad.resize((size_t)(g.nx - 1) * (g.ny - 1) * 4);
    for (int j = 0; j < g.ny - 1; ++j)
        for (int i = 0; i < g.nx - 1; ++i) {
            double *o = &t.quad[((size_t)i + (size_t)(g.nx - 1) * j) * 4];
            o[0] = g.z[i + (size_t)g.nx * j];
            o[1] = g.z[i + 1 + (size_t)g.nx * j];
            o[2] = g.z[i + (size_t)g.nx * (j + 1)];
            o[3] = g.z[i + 1 + (size_t)g.nx * (j + 1)];
        }
    return MMS_OK;
}

int check_month_ids(const int32_t *month_id, const double *n_days, int32_t n_months)
{
    for (int m = 0; m < n_months; ++m) {
        if (month_id[m] < 1 || month_id[m] > 12)
            return fail(MMS_ERR_INVALID_ARG, "month_id[%d] = %d is not in 1..12", m, month_id[m]);
        if (!(n_days[m] == n_days[m])) return fail(MMS_ERR_NONFINITE, "n_days[%d] is NaN", m);
    }
    return MMS_OK;
}

}  // namespace mms_api

using namespace mms_api;

extern "C" {

int mms_abi_version(void) { return MMS_ABI_VERSION; }

int mms_device_count(void) { return device_count_raw(); }

const char *mms_last_error(void) { return last_error().c_str(); }

// Contiguous cell range of shard `index` out of `n_shards`: whole 128-cell blocks, the
// remainder blocks going to the first shards, the last shard taking the ragged tail.
int mms_shard_range(int64_t n_cells, int32_t n_shards, int32_t index, int64_t *cell0, int64_t *count)
{
    if (n_cells < 0 || n_shards < 1 || index < 0 || index >= n_shards || !cell0 || !count)
        return fail(MMS_ERR_INVALID_ARG, "bad shard request");
    const int64_t blocks_total = (n_cells + mms::kBlock - 1) / mms::kBlock;
    const int64_t base = blocks_total / n_shards, extra = blocks_total % n_shards;
    const int64_t b0 = base * index + std::min<int64_t>(index, extra);
    const int64_t nb = base + (index < extra ? 1 : 0);
    const int64_t first = std::min<int64_t>(b0 * mms::kBlock, n_cells);
    const int64_t last = std::min<int64_t>((b0 + nb) * mms::kBlock, n_cells);
    *cell0 = first;
    *count = last - first;
    return MMS_OK;
}

int mms_stage_kernel_ms(float *ms)
{
    if (!ms) return fail(MMS_ERR_INVALID_ARG, "NULL ms");
    *ms = stage_kernel_ms();
    return MMS_OK;
}

// ---- pinned host memory (so that copies run at PCIe speed without the bounce buffers) -------
int mms_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return fail(MMS_ERR_INVALID_ARG, "NULL ptr");
    if (device_count_raw() < 1) return fail(MMS_ERR_NO_DEVICE, "no CUDA device is visible");
    MMS_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return MMS_OK;
}

int mms_host_free(void *ptr)
{
    if (ptr) MMS_CUDA(cudaFreeHost(ptr));
    return MMS_OK;
}

}  // extern "C"
