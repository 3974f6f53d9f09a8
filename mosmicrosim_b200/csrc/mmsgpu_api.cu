// mmsgpu_api.cu -- host side of libmmsgpu: the C ABI declared in
// include/mmsgpu.h.  Validates the call, derives the per-day / per-hour /
// per-setting constants exactly as the reference derives them on the host,
// shards the cells over the requested devices (no inter-device traffic: cells
// never communicate), and launches the sm_100a kernels.  No model arithmetic
// per cell-hour happens here, and there is no CPU fallback.
//
// Citations are relative to the reference root.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mms_common.cuh"

namespace {

thread_local std::string g_error;
thread_local float g_stage_kernel_ms = 0.f;   // device time of the last stage-twin kernel on this thread

// CUDA-event bracket around one kernel launch on a stream
// a stream that is destroyed on every way out of the function that made it
struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) cudaStreamDestroy(s); }
    operator cudaStream_t() const { return s; }
};

struct EventTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    explicit EventTimer(cudaStream_t s) : st(s)
    {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    void stop() { cudaEventRecord(e1, st); }
    float ms()
    {
        float t = 0.f;
        if (cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
        return t;
    }
    ~EventTimer()
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

int fail(int status, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return status;
}

#define MMS_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? MMS_ERR_ALLOC : MMS_ERR_CUDA,      \
                        "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                        cudaGetErrorString(e_));                                             \
    } while (0)

int device_count_raw()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- host-side constants ----------------------------------------------------

// Hour-of-day tables of interpolate_daily_air_temp() (R/sim_microclimate.R:185-285)
// and vsine_weights() (:431-454), evaluated with the host libm like R does.
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

// R's `^` for the two places the host needs it (R_pow: x == 0 -> 0 for y > 0)
double r_pow(double x, double y)
{
    if (y == 2.0) return x * x;
    if (x == 1. || y == 0.) return 1.;
    if (x == 0.) return (y > 0.) ? 0. : INFINITY;
    return std::pow(x, y);
}

// max_cone_volume and catchment_area, R/sim_water.R:12-31,257-261 (+ cylinder EXTENSION)
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

// approx1() interval of every prediction day (base-R approx.c, rule = 1)
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

struct HostTables {
    std::vector<double> sx, scoef;        // spline knots, coefficients [n][3][4]
    int uniform = 0;
    double x0 = 0, step = 0, inv_step = 0;
    std::vector<double> quad;             // ds quad table
};

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

int count_bits(uint32_t m) { return __builtin_popcount(m); }

struct Shard {
    int device = 0;
    int64_t cell0 = 0, n_cells = 0, c_stride = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_monthly = nullptr, *d_altitude = nullptr;
    double *d_terra = nullptr;     // staging of mms_chain_upload_terraclimate(), allocated on first use
    mms::DayEntry *d_days = nullptr;
    int32_t *d_period = nullptr;
    double *d_sx = nullptr, *d_scoef = nullptr, *d_gx = nullptr, *d_gy = nullptr, *d_quad = nullptr;
    double *d_mean_state = nullptr, *d_mean_forcing = nullptr, *d_hourly_state = nullptr, *d_hourly_forcing = nullptr;
    int32_t *d_wet_hours = nullptr;
    uint8_t *d_wet_flags = nullptr;
    int32_t *d_error = nullptr;    // [0] error flag, [1] work-unit counter
    int n_sm = 0;
    bool launched = false;
};

}  // namespace

struct mms_chain_plan {
    int64_t C = 0;
    int M = 0, D = 0, S = 0, P = 0;
    int burnin_water = 0, burnin_pop = 0;
    double initial_adult = 0, initial_aquatic = 0;
    mms_densmod densmod{};
    uint32_t mean_mask = 0, hourly_mask = 0;
    bool want_wet_hours = false, want_wet_flags = false;
    int math_mode = 0;
    int n_mean_state = 0, n_mean_forcing = 0, n_hourly_state = 0, n_hourly_forcing = 0;
    mms::HourTables hr{};
    std::vector<mms::HabitatDev> hab;
    std::vector<mms::DayEntry> days;
    std::vector<int32_t> period;
    HostTables tab;
    int ds_nx = 0, ds_ny = 0;
    std::vector<double> gx, gy;
    std::vector<Shard> shards;
    int64_t launches = 0;
    bool uploaded = false;
};

namespace {

void free_shard(Shard &s)
{
    cudaSetDevice(s.device);
    void *ptrs[] = {s.d_monthly, s.d_altitude, s.d_days, s.d_period, s.d_sx, s.d_scoef, s.d_gx, s.d_gy, s.d_quad,
                    s.d_mean_state, s.d_mean_forcing, s.d_hourly_state, s.d_hourly_forcing, s.d_wet_hours,
                    s.d_wet_flags, s.d_error, s.d_terra};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Shard();
}

template <typename T>
int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) return MMS_OK;
    MMS_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return MMS_OK;
}

template <typename T>
int dev_upload(T **p, const std::vector<T> &v, cudaStream_t st)
{
    int rc = dev_alloc(p, v.size());
    if (rc) return rc;
    if (!v.empty()) MMS_CUDA(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return MMS_OK;
}

int setup_shard(mms_chain_plan *pl, Shard &s)
{
    MMS_CUDA(cudaSetDevice(s.device));
    MMS_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    MMS_CUDA(cudaEventCreate(&s.ev0));
    MMS_CUDA(cudaEventCreate(&s.ev1));
    const size_t cs = (size_t)s.c_stride;
    const size_t H = (size_t)24 * pl->D;
    int rc;
    if ((rc = dev_alloc(&s.d_monthly, (size_t)MMS_NFORCING * pl->M * cs))) return rc;
    if ((rc = dev_alloc(&s.d_altitude, cs))) return rc;
    MMS_CUDA(cudaMemsetAsync(s.d_monthly, 0, sizeof(double) * MMS_NFORCING * pl->M * cs, s.stream));
    MMS_CUDA(cudaMemsetAsync(s.d_altitude, 0, sizeof(double) * cs, s.stream));
    if ((rc = dev_upload(&s.d_days, pl->days, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_period, pl->period, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_sx, pl->tab.sx, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_scoef, pl->tab.scoef, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_gx, pl->gx, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_gy, pl->gy, s.stream))) return rc;
    if ((rc = dev_upload(&s.d_quad, pl->tab.quad, s.stream))) return rc;
    if ((rc = dev_alloc(&s.d_mean_state, (size_t)pl->S * pl->n_mean_state * pl->P * cs))) return rc;
    if ((rc = dev_alloc(&s.d_mean_forcing, (size_t)pl->n_mean_forcing * pl->P * cs))) return rc;
    if ((rc = dev_alloc(&s.d_hourly_state, (size_t)pl->S * pl->n_hourly_state * H * cs))) return rc;
    if ((rc = dev_alloc(&s.d_hourly_forcing, (size_t)pl->n_hourly_forcing * H * cs))) return rc;
    if (pl->want_wet_hours && (rc = dev_alloc(&s.d_wet_hours, (size_t)pl->S * pl->P * cs))) return rc;
    if (pl->want_wet_flags && (rc = dev_alloc(&s.d_wet_flags, (size_t)pl->S * H * cs))) return rc;
    if ((rc = dev_alloc(&s.d_error, 2))) return rc;
    MMS_CUDA(cudaMemsetAsync(s.d_error, 0, 2 * sizeof(int32_t), s.stream));
    MMS_CUDA(cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, s.device));
    return MMS_OK;
}

int copy_in_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t st)
{
    if (width == 0 || height == 0) return MMS_OK;
    MMS_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, st));
    return MMS_OK;
}

int copy_out_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t st)
{
    if (width == 0 || height == 0) return MMS_OK;
    MMS_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, st));
    return MMS_OK;
}

}  // namespace

extern "C" {

int mms_abi_version(void) { return MMS_ABI_VERSION; }

int mms_device_count(void) { return device_count_raw(); }

const char *mms_last_error(void) { return g_error.c_str(); }

void mms_chain_destroy(mms_chain_plan *plan)
{
    if (!plan) return;
    for (Shard &s : plan->shards) free_shard(s);
    delete plan;
}

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

int mms_chain_create(const mms_chain_config *cfg, mms_chain_plan **out)
{
    if (!cfg || !out) return fail(MMS_ERR_INVALID_ARG, "NULL config or plan pointer");
    *out = nullptr;
    if (cfg->n_cells <= 0) return fail(MMS_ERR_INVALID_ARG, "n_cells must be positive");
    if (cfg->n_months < 2) return fail(MMS_ERR_INVALID_ARG, "need at least two monthly knots");
    if (cfg->n_days < 1) return fail(MMS_ERR_INVALID_ARG, "n_days must be positive");
    if (cfg->n_settings < 1 || cfg->n_settings > MMS_MAX_SETTINGS)
        return fail(MMS_ERR_INVALID_ARG, "n_settings must be in 1..%d", MMS_MAX_SETTINGS);
    if (cfg->n_periods < 0) return fail(MMS_ERR_INVALID_ARG, "n_periods must be >= 0");
    if (!cfg->mid_day || !cfg->day_predict || !cfg->habitats) return fail(MMS_ERR_INVALID_ARG, "NULL mid_day, day_predict or habitats");
    if (cfg->burnin_years_water < 0 || cfg->burnin_years_pop < 0) return fail(MMS_ERR_INVALID_ARG, "burn-in years must be >= 0");
    if ((cfg->burnin_years_water > 0 || cfg->burnin_years_pop > 0) && (int64_t)24 * cfg->n_days < mms::kHoursPerYear)
        return fail(MMS_ERR_BURNIN, "burn-in repeats the first %d hourly steps but the series has only %lld "
                    "(R: subscript out of bounds, R/sim_pop_dynamics.R:233)", mms::kHoursPerYear, (long long)24 * cfg->n_days);
    if (cfg->densmod.type != 0 && cfg->densmod.type != 1) return fail(MMS_ERR_INVALID_ARG, "densmod type must be 0 (logit) or 1 (cox_ph)");
    if (cfg->math_mode != 0 && cfg->math_mode != 1) return fail(MMS_ERR_INVALID_ARG, "math_mode must be 0 (fast) or 1 (strict)");
    const uint32_t valid = (1u << MMS_NVARS) - 1u;
    if ((cfg->mean_var_mask | cfg->hourly_var_mask) & ~valid) return fail(MMS_ERR_INVALID_ARG, "unknown variable bit in a mask");
    if ((cfg->mean_var_mask || cfg->want_wet_hours) && (cfg->n_periods < 1 || !cfg->period_of_day))
        return fail(MMS_ERR_INVALID_ARG, "period summaries requested without periods");

    const int n_dev_visible = device_count_raw();
    if (n_dev_visible < 1) return fail(MMS_ERR_NO_DEVICE, "no CUDA device is visible; libmmsgpu has no CPU fallback");
    std::vector<int> devs;
    const int n_dev = cfg->n_devices > 0 ? cfg->n_devices : n_dev_visible;
    for (int k = 0; k < n_dev; ++k) {
        const int d = (cfg->devices && cfg->n_devices > 0) ? cfg->devices[k] : k;
        if (d < 0 || d >= n_dev_visible) return fail(MMS_ERR_INVALID_ARG, "device %d is not visible (%d devices)", d, n_dev_visible);
        devs.push_back(d);
    }

    mms_chain_plan *pl = new mms_chain_plan();
    pl->C = cfg->n_cells; pl->M = cfg->n_months; pl->D = cfg->n_days; pl->S = cfg->n_settings; pl->P = cfg->n_periods;
    pl->burnin_water = cfg->burnin_years_water; pl->burnin_pop = cfg->burnin_years_pop;
    pl->initial_adult = cfg->initial_adult; pl->initial_aquatic = cfg->initial_aquatic;
    pl->densmod = cfg->densmod;
    pl->mean_mask = cfg->mean_var_mask; pl->hourly_mask = cfg->hourly_var_mask;
    pl->want_wet_hours = cfg->want_wet_hours != 0; pl->want_wet_flags = cfg->want_wet_flags != 0;
    pl->math_mode = cfg->math_mode;
    pl->n_mean_state = count_bits(pl->mean_mask & MMS_STATE_VAR_MASK);
    pl->n_mean_forcing = count_bits(pl->mean_mask & ~MMS_STATE_VAR_MASK);
    pl->n_hourly_state = count_bits(pl->hourly_mask & MMS_STATE_VAR_MASK);
    pl->n_hourly_forcing = count_bits(pl->hourly_mask & ~MMS_STATE_VAR_MASK);
    fill_hour_tables(pl->hr);
    int rc = MMS_OK;
    pl->hab.resize(pl->S);
    for (int s = 0; s < pl->S && !rc; ++s) rc = derive_habitat(cfg->habitats[s], pl->hab[s]);
    if (!rc) rc = build_day_table(cfg->mid_day, pl->M, cfg->day_predict, pl->D, pl->days);
    if (!rc && cfg->period_of_day && pl->P > 0) {
        pl->period.assign(cfg->period_of_day, cfg->period_of_day + pl->D);
        int last = -1;
        for (int d = 0; d < pl->D && !rc; ++d) {
            const int q = pl->period[d];
            if (q >= pl->P) rc = fail(MMS_ERR_INVALID_ARG, "period_of_day[%d] = %d >= n_periods", d, q);
            else if (q >= 0 && q < last) rc = fail(MMS_ERR_INVALID_ARG, "period_of_day must be non-decreasing over its non-negative entries");
            if (q >= 0) last = q;
        }
    }
    if (!rc) rc = pack_tables(cfg->lifehistory, pl->tab);
    if (rc) { delete pl; return rc; }
    pl->ds_nx = cfg->lifehistory.ds_temp_humid.nx; pl->ds_ny = cfg->lifehistory.ds_temp_humid.ny;
    pl->gx.assign(cfg->lifehistory.ds_temp_humid.x, cfg->lifehistory.ds_temp_humid.x + pl->ds_nx);
    pl->gy.assign(cfg->lifehistory.ds_temp_humid.y, cfg->lifehistory.ds_temp_humid.y + pl->ds_ny);

    // contiguous cell ranges, one per device; each shard padded to whole CTAs
    const int G = (int)devs.size();
    for (int g = 0; g < G; ++g) {
        int64_t cell0 = 0, n = 0;
        mms_shard_range(pl->C, G, g, &cell0, &n);
        if (n <= 0) continue;
        Shard s;
        s.device = devs[g];
        s.cell0 = cell0;
        s.n_cells = n;
        s.c_stride = ((n + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
        pl->shards.push_back(s);
    }
    for (Shard &s : pl->shards) {
        rc = setup_shard(pl, s);
        if (rc) { mms_chain_destroy(pl); return rc; }
    }
    *out = pl;
    return MMS_OK;
}

int mms_chain_upload(mms_chain_plan *pl, const double *monthly, const double *altitude)
{
    if (!pl || !monthly || !altitude) return fail(MMS_ERR_INVALID_ARG, "NULL plan, monthly or altitude");
    for (Shard &s : pl->shards) {
        MMS_CUDA(cudaSetDevice(s.device));
        int rc = copy_in_2d(s.d_monthly, sizeof(double) * s.c_stride, monthly + s.cell0, sizeof(double) * pl->C,
                            sizeof(double) * s.n_cells, (size_t)MMS_NFORCING * pl->M, s.stream);
        if (rc) return rc;
        MMS_CUDA(cudaMemcpyAsync(s.d_altitude, altitude + s.cell0, sizeof(double) * s.n_cells, cudaMemcpyHostToDevice, s.stream));
        // rainfall_daily = log1p(rainfall_daily) before interpolation, R/sim_microclimate.R:33
        double *rain = s.d_monthly + (size_t)MMS_F_RAINFALL_DAILY * pl->M * s.c_stride;
        const int64_t n = (int64_t)pl->M * s.c_stride;
        MMS_CUDA(pl->math_mode ? mms::launch_log1p_rows_strict(rain, n, s.stream) : mms::launch_log1p_rows_fast(rain, n, s.stream));
        pl->launches += 1;
    }
    pl->uploaded = true;
    return MMS_OK;
}

int mms_chain_launch(mms_chain_plan *pl)
{
    if (!pl) return fail(MMS_ERR_INVALID_ARG, "NULL plan");
    if (!pl->uploaded) return fail(MMS_ERR_INVALID_ARG, "mms_chain_upload has not been called on this plan");
    for (Shard &s : pl->shards) {
        MMS_CUDA(cudaSetDevice(s.device));
        mms::ChainParams p{};
        p.monthly = s.d_monthly; p.altitude = s.d_altitude; p.days = s.d_days;
        p.period_of_day = s.d_period;
        p.c_stride = s.c_stride; p.n_cells = s.n_cells;
        p.M = pl->M; p.D = pl->D; p.S = pl->S; p.P = pl->P;
        p.burnin_years_water = pl->burnin_water; p.burnin_years_pop = pl->burnin_pop;
        p.initial_adult = pl->initial_adult; p.initial_aquatic = pl->initial_aquatic;
        p.densmod_type = pl->densmod.type; p.dd_effect = pl->densmod.dd_effect; p.surface_area_cm2 = pl->densmod.surface_area_cm2;
        p.spline.x = s.d_sx; p.spline.coef = s.d_scoef; p.spline.n = (int)pl->tab.sx.size();
        p.spline.uniform = pl->tab.uniform; p.spline.x0 = pl->tab.x0; p.spline.step = pl->tab.step; p.spline.inv_step = pl->tab.inv_step;
        p.ds.x = s.d_gx; p.ds.y = s.d_gy; p.ds.quad = s.d_quad; p.ds.nx = pl->ds_nx; p.ds.ny = pl->ds_ny;
        p.ds.x0 = pl->gx[0]; p.ds.inv_dx = (pl->ds_nx - 1) / (pl->gx[pl->ds_nx - 1] - pl->gx[0]);
        p.ds.y0 = pl->gy[0]; p.ds.inv_dy = (pl->ds_ny - 1) / (pl->gy[pl->ds_ny - 1] - pl->gy[0]);
        p.ds.x1 = pl->gx[pl->ds_nx - 1]; p.ds.y1 = pl->gy[pl->ds_ny - 1];
        p.mean_state = s.d_mean_state; p.mean_forcing = s.d_mean_forcing;
        p.hourly_state = s.d_hourly_state; p.hourly_forcing = s.d_hourly_forcing;
        p.wet_hours = s.d_wet_hours; p.wet_flags = s.d_wet_flags;
        p.mean_mask = pl->mean_mask; p.hourly_mask = pl->hourly_mask;
        p.error_flag = s.d_error;
        p.unit_counter = s.d_error + 1;
        if (const char *fw = std::getenv("MMS_CHAIN_WARPS")) p.force_warps = std::atoi(fw);
        p.warp_specialised = -1;   // automatic choice; MMS_CHAIN_WS = 0 / 1 forces one form (tests, tuning)
        if (const char *ws = std::getenv("MMS_CHAIN_WS")) p.warp_specialised = std::atoi(ws);
        MMS_CUDA(cudaMemsetAsync(s.d_error + 1, 0, sizeof(int32_t), s.stream));
        p.hr = pl->hr;
        for (int k = 0; k < pl->S; ++k) p.hab[k] = pl->hab[k];
        // periods that never occur stay NaN (all-ones bit pattern) / 0
        const size_t cs = (size_t)s.c_stride;
        if (s.d_mean_state) MMS_CUDA(cudaMemsetAsync(s.d_mean_state, 0xFF, sizeof(double) * pl->S * pl->n_mean_state * pl->P * cs, s.stream));
        if (s.d_mean_forcing) MMS_CUDA(cudaMemsetAsync(s.d_mean_forcing, 0xFF, sizeof(double) * pl->n_mean_forcing * pl->P * cs, s.stream));
        if (s.d_wet_hours) MMS_CUDA(cudaMemsetAsync(s.d_wet_hours, 0, sizeof(int32_t) * pl->S * pl->P * cs, s.stream));
        MMS_CUDA(cudaEventRecord(s.ev0, s.stream));
        // the fast kernels index the tables arithmetically, which needs equally spaced knots /
        // axes (true for every table the reference builds); anything else runs the strict
        // kernels, which search the stored knots like R does
        const bool strict = pl->math_mode != 0 || !pl->tab.uniform;
        MMS_CUDA(strict ? mms::launch_chain_strict(p, s.n_sm, s.stream) : mms::launch_chain_fast(p, s.n_sm, s.stream));
        MMS_CUDA(cudaEventRecord(s.ev1, s.stream));
        s.launched = true;
        pl->launches += 1;
    }
    return MMS_OK;
}

int mms_chain_sync(mms_chain_plan *pl)
{
    if (!pl) return fail(MMS_ERR_INVALID_ARG, "NULL plan");
    for (Shard &s : pl->shards) {
        MMS_CUDA(cudaSetDevice(s.device));
        MMS_CUDA(cudaStreamSynchronize(s.stream));
    }
    for (Shard &s : pl->shards) {
        if (!s.launched) continue;
        MMS_CUDA(cudaSetDevice(s.device));
        int32_t flag = 0;
        MMS_CUDA(cudaMemcpy(&flag, s.d_error, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag)
            return fail(MMS_ERR_NONFINITE, "non-finite (NA/NaN/Inf) monthly forcing in cells %lld..%lld; R's approx() "
                        "would drop those knots (na.rm = TRUE), the GPU backend refuses them",
                        (long long)s.cell0, (long long)(s.cell0 + s.n_cells - 1));
    }
    return MMS_OK;
}

int mms_chain_download(mms_chain_plan *pl, const mms_chain_outputs *out)
{
    if (!pl || !out) return fail(MMS_ERR_INVALID_ARG, "NULL plan or outputs");
    const size_t H = (size_t)24 * pl->D, C = (size_t)pl->C;
    for (Shard &s : pl->shards) {
        MMS_CUDA(cudaSetDevice(s.device));
        const size_t cs = (size_t)s.c_stride, n = (size_t)s.n_cells, c0 = (size_t)s.cell0;
        int rc = MMS_OK;
        if (out->mean_state && s.d_mean_state)
            rc = copy_out_2d(out->mean_state + c0, 8 * C, s.d_mean_state, 8 * cs, 8 * n, (size_t)pl->S * pl->n_mean_state * pl->P, s.stream);
        if (!rc && out->mean_forcing && s.d_mean_forcing)
            rc = copy_out_2d(out->mean_forcing + c0, 8 * C, s.d_mean_forcing, 8 * cs, 8 * n, (size_t)pl->n_mean_forcing * pl->P, s.stream);
        if (!rc && out->hourly_state && s.d_hourly_state)
            rc = copy_out_2d(out->hourly_state + c0, 8 * C, s.d_hourly_state, 8 * cs, 8 * n, (size_t)pl->S * pl->n_hourly_state * H, s.stream);
        if (!rc && out->hourly_forcing && s.d_hourly_forcing)
            rc = copy_out_2d(out->hourly_forcing + c0, 8 * C, s.d_hourly_forcing, 8 * cs, 8 * n, (size_t)pl->n_hourly_forcing * H, s.stream);
        if (!rc && out->wet_hours && s.d_wet_hours)
            rc = copy_out_2d(out->wet_hours + c0, 4 * C, s.d_wet_hours, 4 * cs, 4 * n, (size_t)pl->S * pl->P, s.stream);
        if (!rc && out->wet_flags && s.d_wet_flags)
            rc = copy_out_2d(out->wet_flags + c0, C, s.d_wet_flags, cs, n, (size_t)pl->S * H, s.stream);
        if (rc) return rc;
    }
    return mms_chain_sync(pl);
}

int mms_chain_kernel_ms(mms_chain_plan *pl, float *ms)
{
    if (!pl || !ms) return fail(MMS_ERR_INVALID_ARG, "NULL plan or ms");
    float worst = 0.f;
    for (Shard &s : pl->shards) {
        if (!s.launched) return fail(MMS_ERR_INVALID_ARG, "no launch recorded");
        MMS_CUDA(cudaSetDevice(s.device));
        MMS_CUDA(cudaEventSynchronize(s.ev1));
        float t = 0.f;
        MMS_CUDA(cudaEventElapsedTime(&t, s.ev0, s.ev1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MMS_OK;
}

int mms_chain_launch_count(mms_chain_plan *pl, int64_t *n)
{
    if (!pl || !n) return fail(MMS_ERR_INVALID_ARG, "NULL plan or n");
    *n = pl->launches;
    return MMS_OK;
}

int mms_run_chain(const mms_chain_config *cfg, const double *monthly, const double *altitude,
                  const mms_chain_outputs *out)
{
    mms_chain_plan *pl = nullptr;
    int rc = mms_chain_create(cfg, &pl);
    if (rc) return rc;
    rc = mms_chain_upload(pl, monthly, altitude);
    if (!rc) rc = mms_chain_launch(pl);
    if (!rc) rc = mms_chain_download(pl, out);
    mms_chain_destroy(pl);
    return rc;
}

// ---- stage-level twins -------------------------------------------------------

static int pick_device(int32_t device)
{
    const int n = device_count_raw();
    if (n < 1) return fail(MMS_ERR_NO_DEVICE, "no CUDA device is visible; libmmsgpu has no CPU fallback");
    if (device < 0 || device >= n) return fail(MMS_ERR_INVALID_ARG, "device %d is not visible (%d devices)", device, n);
    MMS_CUDA(cudaSetDevice(device));
    return MMS_OK;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { MMS_CUDA(cudaMalloc(&p, bytes)); return MMS_OK; }
};

// whole_series: the scalar twins' burn-in rule (every pass repeats all n_times rows) instead of the
// vectorised functions' "first 8760 rows"
static int water_twin(const double *rainfall, const double *air_temperature, const double *humidity,
                      const double *windspeed, const double *altitude, int64_t n_times, int64_t n_pixels,
                      const mms_habitat *habitat, int32_t burnin_years, bool whole_series, int32_t math_mode,
                      int32_t device, double *water_surface_area, double *water_volume, uint8_t *wet)
{
    if (!rainfall || !air_temperature || !humidity || !windspeed || !altitude || !habitat || !water_surface_area)
        return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_times < 1 || n_pixels < 1) return fail(MMS_ERR_INVALID_ARG, "matrices must have at least one row and column");
    if (burnin_years < 0) return fail(MMS_ERR_INVALID_ARG, "burnin_years must be >= 0");
    if (!whole_series && burnin_years > 0 && n_times < mms::kHoursPerYear)
        return fail(MMS_ERR_BURNIN, "burn-in repeats the first %d rows but the matrices have %lld "
                    "(R: subscript out of bounds, R/sim_water.R:240)", mms::kHoursPerYear, (long long)n_times);
    int rc = pick_device(device);
    if (rc) return rc;
    mms::WaterStageParams p{};
    if ((rc = derive_habitat(*habitat, p.hab))) return rc;
    const size_t bytes = sizeof(double) * (size_t)n_times * (size_t)n_pixels;
    DevBuf in[4], alt, area, vol, wetb;
    const double *src[4] = {rainfall, air_temperature, humidity, windspeed};
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) {
        if ((rc = in[k].alloc(bytes))) return rc;
        MMS_CUDA(cudaMemcpyAsync(in[k].p, src[k], bytes, cudaMemcpyHostToDevice, st));
    }
    if ((rc = alt.alloc(sizeof(double) * n_pixels)) || (rc = area.alloc(bytes)) ||
        (water_volume && (rc = vol.alloc(bytes))) || (wet && (rc = wetb.alloc((size_t)n_times * n_pixels))))
        return rc;
    MMS_CUDA(cudaMemcpyAsync(alt.p, altitude, sizeof(double) * n_pixels, cudaMemcpyHostToDevice, st));
    p.rainfall = (const double *)in[0].p; p.air_temperature = (const double *)in[1].p;
    p.humidity = (const double *)in[2].p; p.windspeed = (const double *)in[3].p;
    p.altitude = (const double *)alt.p;
    p.water_surface_area = (double *)area.p; p.water_volume = (double *)vol.p; p.wet = (uint8_t *)wetb.p;
    p.H = n_times; p.C = n_pixels; p.burnin_years = burnin_years;
    p.burnin_steps = whole_series ? n_times : (int64_t)mms::kHoursPerYear;
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_water_stage_strict(p, st) : mms::launch_water_stage_fast(p, st));
    timer.stop();
    MMS_CUDA(cudaMemcpyAsync(water_surface_area, area.p, bytes, cudaMemcpyDeviceToHost, st));
    if (water_volume) MMS_CUDA(cudaMemcpyAsync(water_volume, vol.p, bytes, cudaMemcpyDeviceToHost, st));
    if (wet) MMS_CUDA(cudaMemcpyAsync(wet, wetb.p, (size_t)n_times * n_pixels, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

int mms_water_vectorised(const double *rainfall, const double *air_temperature, const double *humidity,
                         const double *windspeed, const double *altitude, int64_t n_times, int64_t n_pixels,
                         const mms_habitat *habitat, int32_t burnin_years, int32_t math_mode, int32_t device,
                         double *water_surface_area, double *water_volume, uint8_t *wet)
{
    return water_twin(rainfall, air_temperature, humidity, windspeed, altitude, n_times, n_pixels, habitat,
                      burnin_years, false, math_mode, device, water_surface_area, water_volume, wet);
}

int mms_water_series(const double *rainfall, const double *air_temperature, const double *humidity,
                     const double *windspeed, const double *altitude, int64_t n_times, int64_t n_pixels,
                     const mms_habitat *habitat, int32_t burnin_years, int32_t math_mode, int32_t device,
                     double *water_surface_area, double *water_volume, uint8_t *wet)
{
    return water_twin(rainfall, air_temperature, humidity, windspeed, altitude, n_times, n_pixels, habitat,
                      burnin_years, true, math_mode, device, water_surface_area, water_volume, wet);
}

static int population_twin(const double *larval_habitat_area, const double *mdr, const double *efd,
                           const double *das_zerodensity, const double *ds, int64_t n_times, int64_t n_pixels,
                           const mms_densmod *densmod, double initial_adult, double initial_aquatic,
                           int32_t burnin_years, bool whole_series, int32_t math_mode, int32_t device,
                           double *aquatic, double *adult)
{
    if (!larval_habitat_area || !mdr || !efd || !das_zerodensity || !ds || !densmod || !aquatic || !adult)
        return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_times < 1 || n_pixels < 1) return fail(MMS_ERR_INVALID_ARG, "matrices must have at least one row and column");
    if (densmod->type != 0 && densmod->type != 1) return fail(MMS_ERR_INVALID_ARG, "densmod type must be 0 (logit) or 1 (cox_ph)");
    if (burnin_years < 0) return fail(MMS_ERR_INVALID_ARG, "burnin_years must be >= 0");
    if (!whole_series && burnin_years > 0 && n_times < mms::kHoursPerYear)
        return fail(MMS_ERR_BURNIN, "burn-in repeats the first %d rows but the matrices have %lld "
                    "(R: subscript out of bounds, R/sim_pop_dynamics.R:233)", mms::kHoursPerYear, (long long)n_times);
    int rc = pick_device(device);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * (size_t)n_times * (size_t)n_pixels;
    DevBuf in[5], aq, ad;
    const double *src[5] = {larval_habitat_area, mdr, efd, das_zerodensity, ds};
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    for (int k = 0; k < 5; ++k) {
        if ((rc = in[k].alloc(bytes))) return rc;
        MMS_CUDA(cudaMemcpyAsync(in[k].p, src[k], bytes, cudaMemcpyHostToDevice, st));
    }
    if ((rc = aq.alloc(bytes)) || (rc = ad.alloc(bytes))) return rc;
    mms::PopStageParams p{};
    p.larval_habitat_area = (const double *)in[0].p; p.mdr = (const double *)in[1].p; p.efd = (const double *)in[2].p;
    p.das_zerodensity = (const double *)in[3].p; p.ds = (const double *)in[4].p;
    p.aquatic = (double *)aq.p; p.adult = (double *)ad.p;
    p.H = n_times; p.C = n_pixels; p.burnin_years = burnin_years;
    p.burnin_steps = whole_series ? n_times : (int64_t)mms::kHoursPerYear;
    p.densmod_type = densmod->type; p.dd_effect = densmod->dd_effect; p.surface_area_cm2 = densmod->surface_area_cm2;
    p.initial_adult = initial_adult; p.initial_aquatic = initial_aquatic;
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_pop_stage_strict(p, st) : mms::launch_pop_stage_fast(p, st));
    timer.stop();
    MMS_CUDA(cudaMemcpyAsync(aquatic, aq.p, bytes, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(cudaMemcpyAsync(adult, ad.p, bytes, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

int mms_population_vectorised(const double *larval_habitat_area, const double *mdr, const double *efd,
                              const double *das_zerodensity, const double *ds, int64_t n_times, int64_t n_pixels,
                              const mms_densmod *densmod, double initial_adult, double initial_aquatic,
                              int32_t burnin_years, int32_t math_mode, int32_t device, double *aquatic, double *adult)
{
    return population_twin(larval_habitat_area, mdr, efd, das_zerodensity, ds, n_times, n_pixels, densmod,
                           initial_adult, initial_aquatic, burnin_years, false, math_mode, device, aquatic, adult);
}

int mms_population_series(const double *larval_habitat_area, const double *mdr, const double *efd,
                          const double *das_zerodensity, const double *ds, int64_t n_times, int64_t n_pixels,
                          const mms_densmod *densmod, double initial_adult, double initial_aquatic,
                          int32_t burnin_years, int32_t math_mode, int32_t device, double *aquatic, double *adult)
{
    return population_twin(larval_habitat_area, mdr, efd, das_zerodensity, ds, n_times, n_pixels, densmod,
                           initial_adult, initial_aquatic, burnin_years, true, math_mode, device, aquatic, adult);
}

// ---- state-free verbs ----------------------------------------------------------

// interpolate_daily_climate() / daily_from_monthly_climate(), R/sim_microclimate.R:7-64
int mms_daily_from_monthly(const double *monthly, const double *mid_day, const double *day_predict,
                           int64_t n_cells, int32_t n_months, int32_t n_days, int32_t math_mode, int32_t device,
                           double *daily)
{
    if (!monthly || !mid_day || !day_predict || !daily) return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_cells < 1 || n_months < 2 || n_days < 1) return fail(MMS_ERR_INVALID_ARG, "need cells >= 1, months >= 2, days >= 1");
    int rc = pick_device(device);
    if (rc) return rc;
    std::vector<mms::DayEntry> days;
    if ((rc = build_day_table(mid_day, n_months, day_predict, n_days, days))) return rc;
    const int64_t cs = ((n_cells + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
    DevBuf d_monthly, d_days, d_out, d_err;
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    if ((rc = d_monthly.alloc(sizeof(double) * MMS_NFORCING * n_months * cs)) || (rc = d_days.alloc(sizeof(mms::DayEntry) * n_days)) ||
        (rc = d_out.alloc(sizeof(double) * MMS_NFORCING * n_days * cs)) || (rc = d_err.alloc(2 * sizeof(int32_t))))
        return rc;
    MMS_CUDA(cudaMemsetAsync(d_monthly.p, 0, sizeof(double) * MMS_NFORCING * n_months * cs, st));
    MMS_CUDA(cudaMemsetAsync(d_err.p, 0, 2 * sizeof(int32_t), st));
    if ((rc = copy_in_2d(d_monthly.p, 8 * cs, monthly, 8 * n_cells, 8 * n_cells, (size_t)MMS_NFORCING * n_months, st))) return rc;
    MMS_CUDA(cudaMemcpyAsync(d_days.p, days.data(), sizeof(mms::DayEntry) * n_days, cudaMemcpyHostToDevice, st));
    double *rain = (double *)d_monthly.p + (size_t)MMS_F_RAINFALL_DAILY * n_months * cs;
    MMS_CUDA(math_mode ? mms::launch_log1p_rows_strict(rain, (int64_t)n_months * cs, st)
                       : mms::launch_log1p_rows_fast(rain, (int64_t)n_months * cs, st));
    mms::ChainParams p{};
    p.monthly = (const double *)d_monthly.p; p.days = (const mms::DayEntry *)d_days.p;
    p.c_stride = cs; p.n_cells = n_cells; p.M = n_months; p.D = n_days;
    p.error_flag = (int32_t *)d_err.p;
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_daily_strict(p, (double *)d_out.p, st) : mms::launch_daily_fast(p, (double *)d_out.p, st));
    timer.stop();
    if ((rc = copy_out_2d(daily, 8 * n_cells, d_out.p, 8 * cs, 8 * n_cells, (size_t)MMS_NFORCING * n_days, st))) return rc;
    int32_t flag = 0;
    MMS_CUDA(cudaMemcpyAsync(&flag, d_err.p, sizeof flag, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    if (flag) return fail(MMS_ERR_NONFINITE, "non-finite (NA/NaN/Inf) monthly forcing; R's approx() would drop those knots, "
                          "the GPU backend refuses them");
    return MMS_OK;
}

// ---- process_terraclimate_pixel_vars(), R/terraclimate.R:581-643 ----------

static int check_month_ids(const int32_t *month_id, const double *n_days, int32_t n_months)
{
    for (int m = 0; m < n_months; ++m) {
        if (month_id[m] < 1 || month_id[m] > 12)
            return fail(MMS_ERR_INVALID_ARG, "month_id[%d] = %d is not in 1..12", m, month_id[m]);
        if (!(n_days[m] == n_days[m])) return fail(MMS_ERR_NONFINITE, "n_days[%d] is NaN", m);
    }
    return MMS_OK;
}

// adjust_wind_speed(ws, height_required = 2, height_measured = 10) on "level grass", R/terraclimate.R:615-617,
// R/utils.R:123-132: the factor is evaluated with the host libm, as R does
static double terraclimate_wind_factor() { return std::pow(2.0 / 10.0, 0.15); }

int mms_terraclimate_to_monthly(const double *terraclimate, const double *clear_sky, const int32_t *month_id,
                                const double *n_days, int64_t n_cells, int32_t n_months, int32_t math_mode,
                                int32_t device, double *monthly)
{
    if (!terraclimate || !clear_sky || !month_id || !n_days || !monthly) return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_cells < 1 || n_months < 1) return fail(MMS_ERR_INVALID_ARG, "need cells >= 1 and months >= 1");
    int rc = check_month_ids(month_id, n_days, n_months);
    if (rc) return rc;
    if ((rc = pick_device(device))) return rc;
    const int64_t cs = ((n_cells + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
    DevBuf d_tc, d_cs, d_mid, d_nd, d_out;
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    if ((rc = d_tc.alloc(sizeof(double) * MMS_NTERRACLIMATE * n_months * cs)) || (rc = d_cs.alloc(sizeof(double) * 12 * cs)) ||
        (rc = d_mid.alloc(sizeof(int32_t) * n_months)) || (rc = d_nd.alloc(sizeof(double) * n_months)) ||
        (rc = d_out.alloc(sizeof(double) * MMS_NFORCING * n_months * cs)))
        return rc;
    if ((rc = copy_in_2d(d_tc.p, 8 * cs, terraclimate, 8 * n_cells, 8 * n_cells, (size_t)MMS_NTERRACLIMATE * n_months, st)) ||
        (rc = copy_in_2d(d_cs.p, 8 * cs, clear_sky, 8 * n_cells, 8 * n_cells, 12, st)))
        return rc;
    MMS_CUDA(cudaMemcpyAsync(d_mid.p, month_id, sizeof(int32_t) * n_months, cudaMemcpyHostToDevice, st));
    MMS_CUDA(cudaMemcpyAsync(d_nd.p, n_days, sizeof(double) * n_months, cudaMemcpyHostToDevice, st));
    mms::TerraParams p{};
    p.terraclimate = (const double *)d_tc.p; p.clear_sky = (const double *)d_cs.p;
    p.month_id = (const int32_t *)d_mid.p; p.n_days = (const double *)d_nd.p; p.monthly = (double *)d_out.p;
    p.c_stride = cs; p.n_cells = n_cells; p.M = n_months; p.log1p_rainfall = 0;
    p.wind_factor = terraclimate_wind_factor();
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_terraclimate_strict(p, st) : mms::launch_terraclimate_fast(p, st));
    timer.stop();
    if ((rc = copy_out_2d(monthly, 8 * n_cells, d_out.p, 8 * cs, 8 * n_cells, (size_t)MMS_NFORCING * n_months, st))) return rc;
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

// mms_chain_upload() for callers that hold the six raw TerraClimate variables: the nine forcing rows
// (rainfall as log1p) are derived on the device, so a third less crosses PCIe
int mms_chain_upload_terraclimate(mms_chain_plan *pl, const double *terraclimate, const double *clear_sky,
                                  const int32_t *month_id, const double *n_days, const double *altitude)
{
    if (!pl || !terraclimate || !clear_sky || !month_id || !n_days || !altitude)
        return fail(MMS_ERR_INVALID_ARG, "NULL plan or argument");
    int rc = check_month_ids(month_id, n_days, pl->M);
    if (rc) return rc;
    for (Shard &s : pl->shards) {
        MMS_CUDA(cudaSetDevice(s.device));
        if (!s.d_terra) {
            // [6][M][cs] raw variables, [12][cs] clear-sky values, then n_days [M] and month_id [M]
            const size_t bytes = sizeof(double) * ((size_t)(MMS_NTERRACLIMATE * pl->M + 12) * s.c_stride + pl->M) +
                                 sizeof(int32_t) * pl->M;
            MMS_CUDA(cudaMalloc(&s.d_terra, bytes));
        }
        double *d_tc = s.d_terra, *d_cs = d_tc + (size_t)MMS_NTERRACLIMATE * pl->M * s.c_stride;
        double *d_nd = d_cs + (size_t)12 * s.c_stride;
        int32_t *d_mid = reinterpret_cast<int32_t *>(d_nd + pl->M);
        if ((rc = copy_in_2d(d_tc, sizeof(double) * s.c_stride, terraclimate + s.cell0, sizeof(double) * pl->C,
                             sizeof(double) * s.n_cells, (size_t)MMS_NTERRACLIMATE * pl->M, s.stream)) ||
            (rc = copy_in_2d(d_cs, sizeof(double) * s.c_stride, clear_sky + s.cell0, sizeof(double) * pl->C,
                             sizeof(double) * s.n_cells, 12, s.stream)))
            return rc;
        MMS_CUDA(cudaMemcpyAsync(d_nd, n_days, sizeof(double) * pl->M, cudaMemcpyHostToDevice, s.stream));
        MMS_CUDA(cudaMemcpyAsync(d_mid, month_id, sizeof(int32_t) * pl->M, cudaMemcpyHostToDevice, s.stream));
        MMS_CUDA(cudaMemcpyAsync(s.d_altitude, altitude + s.cell0, sizeof(double) * s.n_cells, cudaMemcpyHostToDevice, s.stream));
        mms::TerraParams p{};
        p.terraclimate = d_tc; p.clear_sky = d_cs; p.month_id = d_mid; p.n_days = d_nd; p.monthly = s.d_monthly;
        p.c_stride = s.c_stride; p.n_cells = s.n_cells; p.M = pl->M; p.log1p_rainfall = 1;
        p.wind_factor = terraclimate_wind_factor();
        MMS_CUDA(pl->math_mode ? mms::launch_terraclimate_strict(p, s.stream) : mms::launch_terraclimate_fast(p, s.stream));
        pl->launches += 1;
    }
    pl->uploaded = true;
    return MMS_OK;
}

int mms_run_chain_terraclimate(const mms_chain_config *cfg, const double *terraclimate, const double *clear_sky,
                               const int32_t *month_id, const double *n_days, const double *altitude,
                               const mms_chain_outputs *out)
{
    mms_chain_plan *pl = nullptr;
    int rc = mms_chain_create(cfg, &pl);
    if (rc) return rc;
    rc = mms_chain_upload_terraclimate(pl, terraclimate, clear_sky, month_id, n_days, altitude);
    if (!rc) rc = mms_chain_launch(pl);
    if (!rc) rc = mms_chain_download(pl, out);
    mms_chain_destroy(pl);
    return rc;
}

// simulate_hourly_microclimate() / hourly_from_daily_climate_ambient(), R/sim_microclimate.R:126-173
int mms_hourly_from_daily(const double *daily, int64_t n_cells, int32_t n_days, int32_t math_mode, int32_t device,
                          double *hourly)
{
    if (!daily || !hourly) return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_cells < 1 || n_days < 1) return fail(MMS_ERR_INVALID_ARG, "need cells >= 1 and days >= 1");
    int rc = pick_device(device);
    if (rc) return rc;
    const int64_t cs = ((n_cells + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
    DevBuf d_in, d_out;
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    if ((rc = d_in.alloc(sizeof(double) * MMS_NFORCING * n_days * cs)) || (rc = d_out.alloc(sizeof(double) * 6 * 24 * n_days * cs)))
        return rc;
    MMS_CUDA(cudaMemsetAsync(d_in.p, 0, sizeof(double) * MMS_NFORCING * n_days * cs, st));
    if ((rc = copy_in_2d(d_in.p, 8 * cs, daily, 8 * n_cells, 8 * n_cells, (size_t)MMS_NFORCING * n_days, st))) return rc;
    mms::ChainParams p{};
    p.c_stride = cs; p.n_cells = n_cells; p.D = n_days;
    fill_hour_tables(p.hr);
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_hourly_strict(p, (const double *)d_in.p, (double *)d_out.p, st)
                       : mms::launch_hourly_fast(p, (const double *)d_in.p, (double *)d_out.p, st));
    timer.stop();
    if ((rc = copy_out_2d(hourly, 8 * n_cells, d_out.p, 8 * cs, 8 * n_cells, (size_t)6 * 24 * n_days, st))) return rc;
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

// summarise_vectors() / aggregate_vectors(), R/pipeline_functions.R:995-1091
int mms_summarise_vectors(const double *hourly, int64_t n_cells, int32_t n_days, int32_t n_vars,
                          const int32_t *period_of_day, int32_t n_periods, int32_t device, double *means)
{
    if (!hourly || !period_of_day || !means) return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n_cells < 1 || n_days < 1 || n_vars < 1 || n_periods < 1)
        return fail(MMS_ERR_INVALID_ARG, "need cells, days, variables and periods >= 1");
    int last = -1;
    for (int d = 0; d < n_days; ++d) {
        const int q = period_of_day[d];
        if (q >= n_periods) return fail(MMS_ERR_INVALID_ARG, "period_of_day[%d] = %d >= n_periods", d, q);
        if (q >= 0 && q < last) return fail(MMS_ERR_INVALID_ARG, "period_of_day must be non-decreasing over its non-negative entries");
        if (q >= 0) last = q;
    }
    int rc = pick_device(device);
    if (rc) return rc;
    const int64_t cs = ((n_cells + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
    const size_t n_in = (size_t)n_vars * 24 * n_days * cs, n_out = (size_t)n_vars * n_periods * cs;
    DevBuf d_in, d_out, d_pod;
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    if ((rc = d_in.alloc(sizeof(double) * n_in)) || (rc = d_out.alloc(sizeof(double) * n_out)) ||
        (rc = d_pod.alloc(sizeof(int32_t) * n_days)))
        return rc;
    MMS_CUDA(cudaMemsetAsync(d_in.p, 0, sizeof(double) * n_in, st));
    MMS_CUDA(cudaMemsetAsync(d_out.p, 0xFF, sizeof(double) * n_out, st));          // periods without a day: NaN
    if ((rc = copy_in_2d(d_in.p, 8 * cs, hourly, 8 * n_cells, 8 * n_cells, (size_t)n_vars * 24 * n_days, st))) return rc;
    MMS_CUDA(cudaMemcpyAsync(d_pod.p, period_of_day, sizeof(int32_t) * n_days, cudaMemcpyHostToDevice, st));
    EventTimer timer(st);
    MMS_CUDA(mms::launch_summarise_fast((const double *)d_in.p, (const int32_t *)d_pod.p, cs, n_days, n_periods, n_vars,
                                        (double *)d_out.p, st));
    timer.stop();
    if ((rc = copy_out_2d(means, 8 * n_cells, d_out.p, 8 * cs, 8 * n_cells, (size_t)n_vars * n_periods, st))) return rc;
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

// simulate_hourly_lifehistory() / simulate_vector_lifehistory(), R/sim_pop_dynamics.R:100-135
int mms_lifehistory_vectorised(const double *water_temperature, const double *air_temperature, const double *humidity,
                               int64_t n, const mms_lifehistory *lh, int32_t math_mode, int32_t device,
                               double *mdr, double *efd, double *das_zerodensity, double *ds)
{
    if (!water_temperature || !air_temperature || !humidity || !lh || !mdr || !efd || !das_zerodensity || !ds)
        return fail(MMS_ERR_INVALID_ARG, "NULL argument");
    if (n < 1) return fail(MMS_ERR_INVALID_ARG, "need at least one value");
    int rc = pick_device(device);
    if (rc) return rc;
    HostTables tab;
    if ((rc = pack_tables(*lh, tab))) return rc;
    const mms_surface &g = lh->ds_temp_humid;
    DevBuf in[3], out[4], d_sx, d_scoef, d_gx, d_gy, d_quad;
    const double *src[3] = {water_temperature, air_temperature, humidity};
    double *dst[4] = {mdr, efd, das_zerodensity, ds};
    Stream st;
    MMS_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    const size_t bytes = sizeof(double) * (size_t)n;
    for (int k = 0; k < 3; ++k) {
        if ((rc = in[k].alloc(bytes))) return rc;
        MMS_CUDA(cudaMemcpyAsync(in[k].p, src[k], bytes, cudaMemcpyHostToDevice, st));
    }
    for (int k = 0; k < 4; ++k)
        if ((rc = out[k].alloc(bytes))) return rc;
    if ((rc = d_sx.alloc(8 * tab.sx.size())) || (rc = d_scoef.alloc(8 * tab.scoef.size())) || (rc = d_gx.alloc(8 * (size_t)g.nx)) ||
        (rc = d_gy.alloc(8 * (size_t)g.ny)) || (rc = d_quad.alloc(8 * tab.quad.size())))
        return rc;
    MMS_CUDA(cudaMemcpyAsync(d_sx.p, tab.sx.data(), 8 * tab.sx.size(), cudaMemcpyHostToDevice, st));
    MMS_CUDA(cudaMemcpyAsync(d_scoef.p, tab.scoef.data(), 8 * tab.scoef.size(), cudaMemcpyHostToDevice, st));
    MMS_CUDA(cudaMemcpyAsync(d_gx.p, g.x, 8 * (size_t)g.nx, cudaMemcpyHostToDevice, st));
    MMS_CUDA(cudaMemcpyAsync(d_gy.p, g.y, 8 * (size_t)g.ny, cudaMemcpyHostToDevice, st));
    MMS_CUDA(cudaMemcpyAsync(d_quad.p, tab.quad.data(), 8 * tab.quad.size(), cudaMemcpyHostToDevice, st));
    mms::ChainParams p{};
    p.spline.x = (const double *)d_sx.p; p.spline.coef = (const double *)d_scoef.p; p.spline.n = (int)tab.sx.size();
    p.spline.uniform = tab.uniform; p.spline.x0 = tab.x0; p.spline.step = tab.step; p.spline.inv_step = tab.inv_step;
    p.ds.x = (const double *)d_gx.p; p.ds.y = (const double *)d_gy.p; p.ds.quad = (const double *)d_quad.p;
    p.ds.nx = g.nx; p.ds.ny = g.ny;
    p.ds.x0 = g.x[0]; p.ds.inv_dx = (g.nx - 1) / (g.x[g.nx - 1] - g.x[0]); p.ds.x1 = g.x[g.nx - 1];
    p.ds.y0 = g.y[0]; p.ds.inv_dy = (g.ny - 1) / (g.y[g.ny - 1] - g.y[0]); p.ds.y1 = g.y[g.ny - 1];
    int n_sm = 0;
    MMS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const bool strict = math_mode != 0 || !tab.uniform;
    EventTimer timer(st);
    MMS_CUDA(strict ? mms::launch_lifehistory_strict(p, n, (const double *)in[0].p, (const double *)in[1].p, (const double *)in[2].p,
                                                     (double *)out[0].p, (double *)out[1].p, (double *)out[2].p, (double *)out[3].p, n_sm, st)
                    : mms::launch_lifehistory_fast(p, n, (const double *)in[0].p, (const double *)in[1].p, (const double *)in[2].p,
                                                   (double *)out[0].p, (double *)out[1].p, (double *)out[2].p, (double *)out[3].p, n_sm, st));
    timer.stop();
    for (int k = 0; k < 4; ++k) MMS_CUDA(cudaMemcpyAsync(dst[k], out[k].p, bytes, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(cudaStreamSynchronize(st));
    g_stage_kernel_ms = timer.ms();
    return MMS_OK;
}

int mms_stage_kernel_ms(float *ms)
{
    if (!ms) return fail(MMS_ERR_INVALID_ARG, "NULL ms");
    *ms = g_stage_kernel_ms;
    return MMS_OK;
}

// ---- pinned host memory (so that the copies above run at PCIe speed) -------
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
