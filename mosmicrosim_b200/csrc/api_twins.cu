// api_twins.cu -- host side of the stage-level twins (R matrices in, R matrices out) and of the state-free
// verbs of include/mmsgpu.h.  Validates the call, moves the data, launches the sm_100a kernels; no model
// arithmetic happens here and there is no CPU fallback.  Citations are relative to the reference root.
#include "host_pipeline.h"

using namespace mms_api;

namespace {

// One stage-level twin call: n_pixels series of n_times steps per matrix, R's column-major layout (each pixel's
// series contiguous).  The pixels are processed in chunks of about one round of the persistent stage kernel through
// a Pipe (host_pipeline.h): the upload of chunk k+1, the kernel of chunk k and the download of chunk k-1 overlap,
// pageable matrices (R's) go through the workers' pinned bounce buffers.  Matrices that do not fit in the device's
// free memory are walked in several passes of as many pixels as do fit.
struct TwinArrays {
    std::vector<const double *> in;                  // H-by-C inputs
    const double *per_pixel = nullptr;               // [C] (altitude) or nullptr
    std::vector<std::pair<void *, size_t>> out;      // H-by-C outputs: pointer, bytes per value (8, or 1 for the flags)
};
using TwinLaunch = std::function<cudaError_t(int64_t n_pixels, const std::vector<char *> &d_in, const double *d_per_pixel,
                                             const std::vector<char *> &d_out, cudaStream_t)>;

int run_twin(int device, int64_t H, int64_t C, const TwinArrays &a, int round_pixels, const TwinLaunch &launch)
{
    int rc = pick_device(device);
    if (rc) return rc;
    // pixels per pass: what fits in 80 % of the free device memory (MMS_TWIN_PASS_PIXELS: tests)
    size_t bytes_per_pixel = a.per_pixel ? 8 : 0;
    for (size_t k = 0; k < a.in.size(); ++k) bytes_per_pixel += (size_t)H * 8;
    for (auto &o : a.out) bytes_per_pixel += (size_t)H * o.second;
    size_t free_b = 0, total_b = 0;
    MMS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int64_t pass_pixels = (int64_t)std::min<size_t>((size_t)C, (size_t)(0.8 * (double)free_b) / bytes_per_pixel);
    if (const char *e = std::getenv("MMS_TWIN_PASS_PIXELS")) pass_pixels = std::min<int64_t>(C, std::max<int64_t>(1, std::atoll(e)));
    if (pass_pixels < 1) return fail(MMS_ERR_ALLOC, "one pixel's series (%zu bytes) do not fit in the device's free memory", bytes_per_pixel);
    if (pass_pixels < C) pass_pixels = std::max<int64_t>(32, pass_pixels / 32 * 32);

    std::vector<DevBuf> d_in(a.in.size()), d_out(a.out.size());
    DevBuf d_pp;
    for (auto &b : d_in) if ((rc = b.alloc((size_t)pass_pixels * H * 8))) return rc;
    for (size_t k = 0; k < a.out.size(); ++k) if ((rc = d_out[k].alloc((size_t)pass_pixels * H * a.out[k].second))) return rc;
    if (a.per_pixel && (rc = d_pp.alloc((size_t)pass_pixels * 8))) return rc;

    float kernel_ms = 0.f;
    for (int64_t p0 = 0; p0 < C; p0 += pass_pixels) {
        const int64_t np = std::min(pass_pixels, C - p0);
        Pipe pipe;
        pipe.device = device;
        pipe.window_unit = 1;
        // the twins move 40-56 bytes per pixel-step for a few flops: what matters is that the copies of the next and the
        // previous chunk overlap the kernel and that the first upload / last download (not overlapped) are short --
        // chunks of about one GB (all arrays together), at least one per round of the persistent kernel
        int64_t n_chunks = std::max<int64_t>((np + round_pixels - 1) / round_pixels,
                                             (int64_t)(((size_t)np * bytes_per_pixel) >> 30));
        if (const char *e = std::getenv("MMS_TWIN_CHUNKS")) n_chunks = std::max<int64_t>(1, std::atoll(e));
        n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, std::min<int64_t>(64, (np + 31) / 32)));
        const int64_t blocks = (np + 31) / 32;
        for (int64_t k = 0; k < n_chunks; ++k) {
            PipeChunk c;
            c.c0 = blocks * k / n_chunks * 32;
            c.n_cols = c.n_cells = std::min(np, blocks * (k + 1) / n_chunks * 32) - c.c0;
            pipe.chunks.push_back(c);
        }
        PipeJob job;
        job.kind = PIPE_RUN;
        auto mat = [&](const void *host, void *dev, size_t elem) {
            Mat m;
            m.host = (char *)const_cast<void *>(host) + (size_t)p0 * elem;
            m.dev = (char *)dev;
            m.host_pitch = m.dev_pitch = 0;
            m.rows = 1;
            m.elem = elem;
            return m;
        };
        for (size_t k = 0; k < a.in.size(); ++k) job.in.push_back(mat(a.in[k], d_in[k].p, (size_t)H * 8));
        if (a.per_pixel) job.in.push_back(mat(a.per_pixel, d_pp.p, 8));
        for (size_t k = 0; k < a.out.size(); ++k) job.out.push_back(mat(a.out[k].first, d_out[k].p, (size_t)H * a.out[k].second));
        const bool pinned = classify_mats(job.in) & classify_mats(job.out);
        // small calls (tests, single pixels) are not worth a team of workers
        const size_t pass_bytes = (size_t)np * bytes_per_pixel;
        // (pinned matrices need no staging, but several copy streams fill the PCIe link better than one: measured)
        const size_t want = pass_bytes / ((size_t)64 << 20);
        const int n_workers = (int)std::max<size_t>(1, std::min<size_t>(pinned ? 4 : (size_t)default_workers(1), want));
        job.workers = n_workers;
        job.run_chunk = [&](int k) {
            const PipeChunk &c = pipe.chunks[k];
            std::vector<char *> di, dd;
            for (size_t j = 0; j < a.in.size(); ++j) di.push_back((char *)d_in[j].p + (size_t)c.c0 * H * 8);
            for (size_t j = 0; j < a.out.size(); ++j) dd.push_back((char *)d_out[j].p + (size_t)c.c0 * H * a.out[j].second);
            const cudaError_t e = launch(c.n_cells, di, a.per_pixel ? (const double *)d_pp.p + c.c0 : nullptr, dd, pipe.s_compute);
            if (e != cudaSuccess) { pipe.set_error(MMS_ERR_CUDA, std::string("stage kernel launch failed: ") + cudaGetErrorString(e)); return false; }
            return true;
        };
        pipe.start(n_workers);
        PipeJob setup;
        setup.kind = PIPE_CALL;
        setup.call = [&] { return pipe.create_device_objects(); };
        std::string msg;
        int r = pipe.run(&setup, &msg);
        if (!r) r = pipe.run(&job, &msg);
        if (!r) {
            for (PipeChunk &c : pipe.chunks) {
                float t = 0.f;
                if (c.kernel_recorded && cudaEventSynchronize(c.ev_t1) == cudaSuccess && cudaEventElapsedTime(&t, c.ev_t0, c.ev_t1) == cudaSuccess) kernel_ms += t;
            }
        }
        pipe.stop();
        pipe.destroy_device_objects();
        if (r) return fail_str(r, msg);
    }
    stage_kernel_ms() = kernel_ms;
    return MMS_OK;
}

// pixels one round of the persistent stage kernels holds on this device (chain_tu.cu, stage_plan)
int twin_round_pixels(int device, int math_mode)
{
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    return n_sm * (math_mode ? 12 : 16) * 32;
}

}  // namespace

extern "C" {

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
    mms::WaterStageParams base{};
    if ((rc = derive_habitat(*habitat, base.hab))) return rc;
    base.H = n_times; base.burnin_years = burnin_years;
    base.burnin_steps = whole_series ? n_times : (int64_t)mms::kHoursPerYear;
    TwinArrays a;
    a.in = {rainfall, air_temperature, humidity, windspeed};
    a.per_pixel = altitude;
    a.out.push_back({water_surface_area, 8});
    if (water_volume) a.out.push_back({water_volume, 8});
    if (wet) a.out.push_back({wet, 1});
    const bool has_vol = water_volume != nullptr, has_wet = wet != nullptr;
    return run_twin(device, n_times, n_pixels, a, twin_round_pixels(device, math_mode),
                    [&](int64_t n, const std::vector<char *> &di, const double *d_alt, const std::vector<char *> &dd, cudaStream_t st) {
                        mms::WaterStageParams p = base;
                        p.rainfall = (const double *)di[0]; p.air_temperature = (const double *)di[1];
                        p.humidity = (const double *)di[2]; p.windspeed = (const double *)di[3];
                        p.altitude = d_alt;
                        p.water_surface_area = (double *)dd[0];
                        p.water_volume = has_vol ? (double *)dd[1] : nullptr;
                        p.wet = has_wet ? (uint8_t *)dd[has_vol ? 2 : 1] : nullptr;
                        p.C = n;
                        return math_mode ? mms::launch_water_stage_strict(p, st) : mms::launch_water_stage_fast(p, st);
                    });
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
    mms::PopStageParams base{};
    base.H = n_times; base.burnin_years = burnin_years;
    base.burnin_steps = whole_series ? n_times : (int64_t)mms::kHoursPerYear;
    base.densmod_type = densmod->type; base.dd_effect = densmod->dd_effect; base.surface_area_cm2 = densmod->surface_area_cm2;
    base.initial_adult = initial_adult; base.initial_aquatic = initial_aquatic;
    TwinArrays a;
    a.in = {larval_habitat_area, mdr, efd, das_zerodensity, ds};
    a.out = {{aquatic, 8}, {adult, 8}};
    return run_twin(device, n_times, n_pixels, a, twin_round_pixels(device, math_mode),
                    [&](int64_t n, const std::vector<char *> &di, const double *, const std::vector<char *> &dd, cudaStream_t st) {
                        mms::PopStageParams p = base;
                        p.larval_habitat_area = (const double *)di[0]; p.mdr = (const double *)di[1]; p.efd = (const double *)di[2];
                        p.das_zerodensity = (const double *)di[3]; p.ds = (const double *)di[4];
                        p.aquatic = (double *)dd[0]; p.adult = (double *)dd[1];
                        p.C = n;
                        return math_mode ? mms::launch_pop_stage_strict(p, st) : mms::launch_pop_stage_fast(p, st);
                    });
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
    MMS_CUDA(math_mode ? mms::launch_log1p_rows_strict(rain, n_months, cs, cs, st)
                       : mms::launch_log1p_rows_fast(rain, n_months, cs, cs, st));
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
    stage_kernel_ms() = timer.ms();
    if (flag) return fail(MMS_ERR_NONFINITE, "non-finite (NA/NaN/Inf) monthly forcing; R's approx() would drop those knots, "
                          "the GPU backend refuses them");
    return MMS_OK;
}

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
    p.c_stride = cs; p.n_cells = n_cells; p.n_cols = cs; p.M = n_months; p.log1p_rainfall = 0;
    p.wind_factor = terraclimate_wind_factor();
    EventTimer timer(st);
    MMS_CUDA(math_mode ? mms::launch_terraclimate_strict(p, st) : mms::launch_terraclimate_fast(p, st));
    timer.stop();
    if ((rc = copy_out_2d(monthly, 8 * n_cells, d_out.p, 8 * cs, 8 * n_cells, (size_t)MMS_NFORCING * n_months, st))) return rc;
    MMS_CUDA(cudaStreamSynchronize(st));
    stage_kernel_ms() = timer.ms();
    return MMS_OK;
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
    stage_kernel_ms() = timer.ms();
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
    stage_kernel_ms() = timer.ms();
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
    stage_kernel_ms() = timer.ms();
    return MMS_OK;
}

}  // extern "C"
