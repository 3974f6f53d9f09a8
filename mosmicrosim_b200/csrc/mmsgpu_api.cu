// mmsgpu_api.cu -- the fused-chain plan of libmmsgpu (include/mmsgpu.h) and its host pipeline.
//
// A plan shards the cells contiguously over the requested devices (no inter-device traffic: cells never
// communicate).  Every shard owns a small team of host threads ("workers") bound to its device, so that the
// shards of a multi-GPU plan are fed and drained concurrently -- the replacement of the reference's
// multi-process furrr fan-out over pixel batches (query_terraclimate.R:117-123,166-175,196-241) is ONE call on
// ONE process.  Within a shard the cells are cut into chunks (column windows of the shard's [...][cell] arrays,
// each about one round of the persistent kernel); mms_chain_run() pipelines them:
//
//      worker w:   stage + H2D chunk k      | drain chunk k-1        | stage + H2D chunk k+1 | ...
//      device  :   kernel k-1               | kernel k               | kernel k+1            | ...
//
// The last worker to deliver its part of chunk k enqueues that chunk's kernels on the shard's compute stream
// behind the workers' copy events; each worker's download of its columns waits on the chunk's kernel event.
// Host memory that the driver can DMA directly (mms_host_alloc, cudaHostRegister) is copied in place with one
// cudaMemcpy2DAsync per array and chunk; ordinary pageable memory -- what R vectors and numpy arrays are --
// goes through per-worker rings of pinned bounce buffers (worker memcpy of piece j overlaps the DMA of piece
// j +- 1), several workers per device, because one thread's memcpy is slower than the PCIe link.
//
// No model arithmetic per cell-hour happens here, and there is no CPU fallback.
// Citations are relative to the reference root.
#include "host_pipeline.h"

using namespace mms_api;

namespace {

constexpr int kMaxChunks = 64;                    // chunks per shard

struct Shard {
    mms_chain_plan *plan = nullptr;
    int device = 0;
    int64_t cell0 = 0, n_cells = 0, c_stride = 0;   // the shard's cells [cell0, cell0 + n_cells) of the caller's arrays
    // output geometry: rows of out_stride columns; the shard's first column is column out_col0 of the caller's rows.
    // Without a scatter that is the cell geometry again (out_col0 = cell0, out_stride = c_stride).
    int64_t out_col0 = 0, out_cols = 0, out_stride = 0;
    int64_t *d_out_col = nullptr;     // [c_stride] column of every cell relative to out_col0 (scatter only)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // whole-shard launch (mms_chain_launch)
    bool timing_chunks = false;                 // kernel time = sum over chunks (mms_chain_run) or ev0..ev1
    double *d_monthly = nullptr, *d_altitude = nullptr;
    double *d_terra = nullptr;     // staging of the terraclimate route, allocated on first use
    mms::DayEntry *d_days = nullptr;
    int32_t *d_period = nullptr;
    double *d_sx = nullptr, *d_scoef = nullptr, *d_gx = nullptr, *d_gy = nullptr, *d_quad = nullptr;
    double *d_mean_state = nullptr, *d_mean_forcing = nullptr, *d_hourly_state = nullptr, *d_hourly_forcing = nullptr;
    int32_t *d_wet_hours = nullptr;
    uint8_t *d_wet_flags = nullptr;
    int32_t *d_error = nullptr;    // [chunk][2]: error flag, work-unit counter
    // water-temperature EXTENSION: per-day tables, per-cell location inputs (mms_chain_set_location)
    double2 *d_sun_day = nullptr;
    int32_t *d_month_of_day = nullptr;
    double *d_sin_lat = nullptr, *d_cos_lat = nullptr, *d_clear_sky = nullptr;
    int n_sm = 0;
    bool launched = false;
    Pipe pipe;                     // the device's compute stream, chunks and host workers (host_pipeline.h)
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
    bool all_periods_present = false;      // every period has a day: no NaN fill needed
    int math_mode = 0;
    int n_mean_state = 0, n_mean_forcing = 0, n_hourly_state = 0, n_hourly_forcing = 0;
    mms::HourTables hr{};
    std::vector<mms::HabitatDev> hab;
    std::vector<mms::DayEntry> days;
    std::vector<int32_t> period;
    HostTables tab;
    int ds_nx = 0, ds_ny = 0;
    std::vector<double> gx, gy;
    std::vector<std::unique_ptr<Shard>> shards;
    std::atomic<int64_t> launches{0};
    bool uploaded = false;
    int n_workers = 1;
    // outputs scattered into raster cell order (N2): cell c -> column cell_idx[c] of rows R long
    bool scatter = false;
    std::vector<int64_t> cell_idx;
    int64_t R = 0;                         // columns of the caller's output rows (C without a scatter)
    // water-temperature EXTENSION
    mms_water_temperature wt{};
    std::vector<double2> sun_day;          // (sin delta, cos delta) of every day, host libm
    std::vector<int32_t> month_of_day;
    bool location_set = false;
};

namespace {

// First output column owned by the cells from global index b on: a window of cells [a, b) owns the columns
// [out_boundary(a), out_boundary(b)) of the caller's rows -- its cells' own columns plus the gap (raster cells that
// are not pixels of the tile) up to the next window's first cell; the first window also owns the leading gap, the
// last one the trailing gap.  Any partition of the cells into windows therefore tiles the rows exactly.
int64_t out_boundary(const mms_chain_plan *pl, int64_t b)
{
    if (!pl->scatter) return b;
    if (b <= 0) return 0;
    if (b >= pl->C) return pl->R;
    return pl->cell_idx[(size_t)b];
}

#define MMS_CUDA_W(shard, call) MMS_CUDA_P(&(shard)->pipe, call)

template <typename T>
bool dev_alloc(Shard *s, T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) return true;
    MMS_CUDA_W(s, cudaMalloc((void **)p, count * sizeof(T)));
    return true;
}

template <typename T>
bool dev_upload(Shard *s, T **p, const std::vector<T> &v, cudaStream_t st)
{
    if (!dev_alloc(s, p, v.size())) return false;
    if (!v.empty()) MMS_CUDA_W(s, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return true;
}

// device memory, tables, streams and events of one shard (runs on the shard's worker 0)
bool setup_shard(Shard *s)
{
    mms_chain_plan *pl = s->plan;
    if (!s->pipe.create_device_objects()) return false;
    MMS_CUDA_W(s, cudaEventCreate(&s->ev0));
    MMS_CUDA_W(s, cudaEventCreate(&s->ev1));
    const size_t cs = (size_t)s->c_stride, os = (size_t)s->out_stride;
    const size_t H = (size_t)24 * pl->D;
    cudaStream_t st = s->pipe.s_compute;
    if (pl->scatter) {
        std::vector<int64_t> col((size_t)s->c_stride, 0);
        for (int64_t c = 0; c < s->n_cells; ++c) col[(size_t)c] = pl->cell_idx[(size_t)(s->cell0 + c)] - s->out_col0;
        if (!dev_upload(s, &s->d_out_col, col, st)) return false;
        MMS_CUDA_W(s, cudaStreamSynchronize(st));          // `col` goes out of scope
    }
    if (!dev_alloc(s, &s->d_monthly, (size_t)MMS_NFORCING * pl->M * cs)) return false;
    if (!dev_alloc(s, &s->d_altitude, cs)) return false;
    MMS_CUDA_W(s, cudaMemsetAsync(s->d_monthly, 0, sizeof(double) * MMS_NFORCING * pl->M * cs, st));
    MMS_CUDA_W(s, cudaMemsetAsync(s->d_altitude, 0, sizeof(double) * cs, st));
    if (!dev_upload(s, &s->d_days, pl->days, st)) return false;
    if (!dev_upload(s, &s->d_period, pl->period, st)) return false;
    if (!dev_upload(s, &s->d_sx, pl->tab.sx, st)) return false;
    if (!dev_upload(s, &s->d_scoef, pl->tab.scoef, st)) return false;
    if (!dev_upload(s, &s->d_gx, pl->gx, st)) return false;
    if (!dev_upload(s, &s->d_gy, pl->gy, st)) return false;
    if (!dev_upload(s, &s->d_quad, pl->tab.quad, st)) return false;
    if (!dev_alloc(s, &s->d_mean_state, (size_t)pl->S * pl->n_mean_state * pl->P * os)) return false;
    if (!dev_alloc(s, &s->d_mean_forcing, (size_t)pl->n_mean_forcing * pl->P * os)) return false;
    if (!dev_alloc(s, &s->d_hourly_state, (size_t)pl->S * pl->n_hourly_state * H * os)) return false;
    if (!dev_alloc(s, &s->d_hourly_forcing, (size_t)pl->n_hourly_forcing * H * os)) return false;
    if (pl->want_wet_hours && !dev_alloc(s, &s->d_wet_hours, (size_t)pl->S * pl->P * os)) return false;
    if (pl->want_wet_flags && !dev_alloc(s, &s->d_wet_flags, (size_t)pl->S * H * os)) return false;
    if (pl->wt.enabled) {
        if (!dev_upload(s, &s->d_sun_day, pl->sun_day, st)) return false;
        if (!dev_upload(s, &s->d_month_of_day, pl->month_of_day, st)) return false;
        if (!dev_alloc(s, &s->d_sin_lat, cs) || !dev_alloc(s, &s->d_cos_lat, cs) || !dev_alloc(s, &s->d_clear_sky, 12 * cs)) return false;
        MMS_CUDA_W(s, cudaMemsetAsync(s->d_sin_lat, 0, sizeof(double) * cs, st));
        MMS_CUDA_W(s, cudaMemsetAsync(s->d_cos_lat, 0, sizeof(double) * cs, st));
        MMS_CUDA_W(s, cudaMemsetAsync(s->d_clear_sky, 0, sizeof(double) * 12 * cs, st));
    }
    if (!dev_alloc(s, &s->d_error, 2 * s->pipe.chunks.size())) return false;
    MMS_CUDA_W(s, cudaMemsetAsync(s->d_error, 0, 2 * s->pipe.chunks.size() * sizeof(int32_t), st));
    MMS_CUDA_W(s, cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, s->device));
    MMS_CUDA_W(s, cudaStreamSynchronize(st));
    return true;
}

void free_shard_device(Shard *s)
{
    cudaSetDevice(s->device);
    void *ptrs[] = {s->d_monthly, s->d_altitude, s->d_days, s->d_period, s->d_sx, s->d_scoef, s->d_gx, s->d_gy, s->d_quad,
                    s->d_mean_state, s->d_mean_forcing, s->d_hourly_state, s->d_hourly_forcing, s->d_wet_hours,
                    s->d_wet_flags, s->d_error, s->d_terra, s->d_out_col, s->d_sun_day, s->d_month_of_day, s->d_sin_lat, s->d_cos_lat,
                    s->d_clear_sky};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    s->pipe.destroy_device_objects();
}

// ---- kernels of one column window [c0, c0 + n_cols) of a shard ------------------------------------------

// the forcing rows of the window are on the device: derive what the fused kernel reads
bool enqueue_transform(Shard *s, bool terraclimate, int64_t c0, int64_t n_cols, int64_t n_cells)
{
    mms_chain_plan *pl = s->plan;
    if (n_cols <= 0) return true;
    if (terraclimate) {
        // process_terraclimate_pixel_vars() on the device (R/terraclimate.R:581-643), rainfall stored as log1p
        const size_t cs = (size_t)s->c_stride;
        double *d_tc = s->d_terra, *d_cs = d_tc + (size_t)MMS_NTERRACLIMATE * pl->M * cs;
        double *d_nd = d_cs + (size_t)12 * cs;
        int32_t *d_mid = reinterpret_cast<int32_t *>(d_nd + pl->M);
        mms::TerraParams p{};
        p.terraclimate = d_tc + c0; p.clear_sky = d_cs + c0; p.month_id = d_mid; p.n_days = d_nd;
        p.monthly = s->d_monthly + c0;
        p.c_stride = s->c_stride; p.n_cells = n_cells; p.n_cols = n_cols; p.M = pl->M; p.log1p_rainfall = 1;
        p.wind_factor = terraclimate_wind_factor();
        MMS_CUDA_W(s, pl->math_mode ? mms::launch_terraclimate_strict(p, s->pipe.s_compute) : mms::launch_terraclimate_fast(p, s->pipe.s_compute));
    } else {
        // rainfall_daily = log1p(rainfall_daily) before interpolation, R/sim_microclimate.R:33
        double *rain = s->d_monthly + (size_t)MMS_F_RAINFALL_DAILY * pl->M * s->c_stride + c0;
        MMS_CUDA_W(s, pl->math_mode ? mms::launch_log1p_rows_strict(rain, pl->M, n_cols, s->c_stride, s->pipe.s_compute)
                                    : mms::launch_log1p_rows_fast(rain, pl->M, n_cols, s->c_stride, s->pipe.s_compute));
    }
    pl->launches += 1;
    return true;
}

// the fused chain over the window; err_slot: which (flag, counter) pair of d_error the launch uses
bool enqueue_chain(Shard *s, int64_t c0, int64_t n_cols, int64_t n_cells, int err_slot)
{
    mms_chain_plan *pl = s->plan;
    if (n_cols <= 0) return true;
    cudaStream_t st = s->pipe.s_compute;
    const size_t os = (size_t)s->out_stride;
    const size_t H = (size_t)24 * pl->D;
    // output columns of this window, relative to the shard's first output column
    const int64_t ga = s->cell0 + std::min(c0, s->n_cells), gb = s->cell0 + std::min(c0 + n_cells, s->n_cells);
    const int64_t o0 = out_boundary(pl, ga) - s->out_col0;
    const size_t ow = pl->scatter ? (size_t)(out_boundary(pl, gb) - out_boundary(pl, ga)) : (size_t)n_cols;
    // without a scatter the kernel addresses the window's columns (all pointers offset to its first cell); with one,
    // out_col[] holds shard-relative columns, so the output pointers stay at the shard's origin
    const int64_t ob = pl->scatter ? 0 : c0;
    mms::ChainParams p{};
    p.out_stride = s->out_stride;
    p.out_col = pl->scatter ? s->d_out_col + c0 : nullptr;
    p.monthly = s->d_monthly + c0; p.altitude = s->d_altitude + c0; p.days = s->d_days;
    p.period_of_day = s->d_period;
    p.c_stride = s->c_stride; p.n_cells = n_cells; p.n_cols = n_cols;
    p.M = pl->M; p.D = pl->D; p.S = pl->S; p.P = pl->P;
    p.burnin_years_water = pl->burnin_water; p.burnin_years_pop = pl->burnin_pop;
    p.initial_adult = pl->initial_adult; p.initial_aquatic = pl->initial_aquatic;
    p.densmod_type = pl->densmod.type; p.dd_effect = pl->densmod.dd_effect; p.surface_area_cm2 = pl->densmod.surface_area_cm2;
    p.spline.x = s->d_sx; p.spline.coef = s->d_scoef; p.spline.n = (int)pl->tab.sx.size();
    p.spline.uniform = pl->tab.uniform; p.spline.x0 = pl->tab.x0; p.spline.step = pl->tab.step; p.spline.inv_step = pl->tab.inv_step;
    p.ds.x = s->d_gx; p.ds.y = s->d_gy; p.ds.quad = s->d_quad; p.ds.nx = pl->ds_nx; p.ds.ny = pl->ds_ny;
    p.ds.x0 = pl->gx[0]; p.ds.inv_dx = (pl->ds_nx - 1) / (pl->gx[pl->ds_nx - 1] - pl->gx[0]);
    p.ds.y0 = pl->gy[0]; p.ds.inv_dy = (pl->ds_ny - 1) / (pl->gy[pl->ds_ny - 1] - pl->gy[0]);
    p.ds.x1 = pl->gx[pl->ds_nx - 1]; p.ds.y1 = pl->gy[pl->ds_ny - 1];
    p.mean_state = s->d_mean_state ? s->d_mean_state + ob : nullptr;
    p.mean_forcing = s->d_mean_forcing ? s->d_mean_forcing + ob : nullptr;
    p.hourly_state = s->d_hourly_state ? s->d_hourly_state + ob : nullptr;
    p.hourly_forcing = s->d_hourly_forcing ? s->d_hourly_forcing + ob : nullptr;
    p.wet_hours = s->d_wet_hours ? s->d_wet_hours + ob : nullptr;
    p.wet_flags = s->d_wet_flags ? s->d_wet_flags + ob : nullptr;
    p.mean_mask = pl->mean_mask; p.hourly_mask = pl->hourly_mask;
    if (pl->wt.enabled) {
        p.wt.enabled = 1;
        p.wt.solar_gain = (1.0 - pl->wt.shade_proportion) * (1.0 - pl->wt.albedo);
        p.wt.sigma_eps = 5.67e-8 * pl->wt.emissivity;
        p.wt.convection = pl->wt.convection_w_m2_k;
        p.wt.sun_day = s->d_sun_day; p.wt.month_of_day = s->d_month_of_day;
        p.wt.sin_lat = s->d_sin_lat + c0; p.wt.cos_lat = s->d_cos_lat + c0; p.wt.clear_sky = s->d_clear_sky + c0;
    }
    p.error_flag = s->d_error + 2 * err_slot;
    p.unit_counter = s->d_error + 2 * err_slot + 1;
    if (const char *fw = std::getenv("MMS_CHAIN_WARPS")) p.force_warps = std::atoi(fw);
    p.warp_specialised = -1;   // automatic choice; MMS_CHAIN_WS = 0 / 1 forces one form (tests, tuning)
    if (const char *ws = std::getenv("MMS_CHAIN_WS")) p.warp_specialised = std::atoi(ws);
    // both the error flag and the work-unit counter start from zero at every launch: a tile with
    // non-finite forcing does not poison the launches that follow it on the same plan
    MMS_CUDA_W(s, cudaMemsetAsync(s->d_error + 2 * err_slot, 0, 2 * sizeof(int32_t), st));
    p.hr = pl->hr;
    for (int k = 0; k < pl->S; ++k) p.hab[k] = pl->hab[k];
    // periods that never occur, and raster cells that are not pixels of the tile, stay NaN (all-ones bit pattern) / 0
    if ((!pl->all_periods_present || pl->scatter) && ow > 0) {
        if (s->d_mean_state) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_mean_state + o0, 8 * os, 0xFF, 8 * ow, (size_t)pl->S * pl->n_mean_state * pl->P, st));
        if (s->d_mean_forcing) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_mean_forcing + o0, 8 * os, 0xFF, 8 * ow, (size_t)pl->n_mean_forcing * pl->P, st));
        if (s->d_wet_hours) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_wet_hours + o0, 4 * os, 0, 4 * ow, (size_t)pl->S * pl->P, st));
    }
    if (pl->scatter && ow > 0) {
        if (s->d_hourly_state) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_hourly_state + o0, 8 * os, 0xFF, 8 * ow, (size_t)pl->S * pl->n_hourly_state * H, st));
        if (s->d_hourly_forcing) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_hourly_forcing + o0, 8 * os, 0xFF, 8 * ow, (size_t)pl->n_hourly_forcing * H, st));
        if (s->d_wet_flags) MMS_CUDA_W(s, cudaMemset2DAsync(s->d_wet_flags + o0, os, 0, ow, (size_t)pl->S * H, st));
    }
    // the fast kernels index the tables arithmetically, which needs equally spaced knots /
    // axes (true for every table the reference builds); anything else runs the strict
    // kernels, which search the stored knots like R does
    const bool strict = pl->math_mode != 0 || !pl->tab.uniform;
    MMS_CUDA_W(s, strict ? mms::launch_chain_strict(p, s->n_sm, st) : mms::launch_chain_fast(p, s->n_sm, st));
    pl->launches += 1;
    return true;
}

// the output columns (relative to the shard's first output column) that go with the cell window [w0, w0 + wn)
void out_window(const Shard *s, int64_t w0, int64_t wn, int64_t &o0, int64_t &on)
{
    const mms_chain_plan *pl = s->plan;
    const int64_t a = out_boundary(pl, s->cell0 + w0), b = out_boundary(pl, s->cell0 + w0 + wn);
    o0 = a - s->out_col0;
    on = wn > 0 ? b - a : 0;
}

// hand one job per shard to the shards' workers and wait for all of them; returns the first failure
int dispatch(mms_chain_plan *pl, std::vector<PipeJob> &jobs)
{
    for (size_t g = 0; g < pl->shards.size(); ++g) pl->shards[g]->pipe.post(&jobs[g]);
    int rc = MMS_OK;
    for (auto &sp : pl->shards) {
        std::string msg;
        const int r = sp->pipe.wait(&msg);
        if (r != MMS_OK && rc == MMS_OK) rc = fail_str(r, msg);
    }
    return rc;
}

Mat make_mat(const void *host, size_t host_cells, void *dev, const Shard *s, size_t rows, size_t elem)
{
    Mat m;
    m.host = (char *)const_cast<void *>(host) + (size_t)s->cell0 * elem;
    m.host_pitch = host_cells * elem;
    m.dev = (char *)dev;
    m.dev_pitch = (size_t)s->c_stride * elem;
    m.rows = rows;
    m.elem = elem;
    return m;
}

// an output array: rows of R columns on the host (R = C without a scatter), of out_stride on the device
Mat make_out_mat(const mms_chain_plan *pl, void *host, void *dev, const Shard *s, size_t rows, size_t elem)
{
    Mat m;
    m.host = (char *)host + (size_t)s->out_col0 * elem;
    m.host_pitch = (size_t)pl->R * elem;
    m.dev = (char *)dev;
    m.dev_pitch = (size_t)s->out_stride * elem;
    m.rows = rows;
    m.elem = elem;
    return m;
}

void add_outputs(mms_chain_plan *pl, Shard *s, const mms_chain_outputs *out, PipeJob &job)
{
    const size_t H = (size_t)24 * pl->D;
    if (out->mean_state && s->d_mean_state)
        job.out.push_back(make_out_mat(pl, out->mean_state, s->d_mean_state, s, (size_t)pl->S * pl->n_mean_state * pl->P, 8));
    if (out->mean_forcing && s->d_mean_forcing)
        job.out.push_back(make_out_mat(pl, out->mean_forcing, s->d_mean_forcing, s, (size_t)pl->n_mean_forcing * pl->P, 8));
    if (out->hourly_state && s->d_hourly_state)
        job.out.push_back(make_out_mat(pl, out->hourly_state, s->d_hourly_state, s, (size_t)pl->S * pl->n_hourly_state * H, 8));
    if (out->hourly_forcing && s->d_hourly_forcing)
        job.out.push_back(make_out_mat(pl, out->hourly_forcing, s->d_hourly_forcing, s, (size_t)pl->n_hourly_forcing * H, 8));
    if (out->wet_hours && s->d_wet_hours)
        job.out.push_back(make_out_mat(pl, out->wet_hours, s->d_wet_hours, s, (size_t)pl->S * pl->P, 4));
    if (out->wet_flags && s->d_wet_flags)
        job.out.push_back(make_out_mat(pl, out->wet_flags, s->d_wet_flags, s, (size_t)pl->S * H, 1));
    job.out_window = [s](int64_t w0, int64_t wn, int64_t &o0, int64_t &on) { out_window(s, w0, wn, o0, on); };
}

// pinned or pageable?  One driver query per host array; a call whose arrays are all pinned needs one worker per device
void classify(mms_chain_plan *pl, std::vector<PipeJob> &jobs)
{
    bool all_pinned = true;
    for (PipeJob &j : jobs) {
        all_pinned = classify_mats(j.in) && all_pinned;
        all_pinned = classify_mats(j.out) && all_pinned;
    }
    // pageable memory: every worker stages its share.  Pinned memory needs no staging, but a few copy streams fill
    // the PCIe link better than one when the chunks are large
    for (size_t g = 0; g < jobs.size(); ++g) {
        PipeJob &j = jobs[g];
        if (!all_pinned) { j.workers = pl->n_workers; continue; }
        const Shard *s = pl->shards[g].get();
        size_t bytes = 0;
        for (const Mat &m : j.in) bytes += m.rows * m.elem;
        for (const Mat &m : j.out) bytes += m.rows * m.elem;
        bytes = bytes * (size_t)s->n_cells / std::max<size_t>(1, s->pipe.chunks.size());      // per chunk
        j.workers = (int)std::max<size_t>(1, std::min<size_t>(std::min(4, pl->n_workers), bytes / ((size_t)16 << 20)));
    }
}

// non-finite forcing seen by the kernels of the last launches (flags are zeroed at every launch)
int check_error_flags(mms_chain_plan *pl)
{
    for (auto &sp : pl->shards) {
        Shard *s = sp.get();
        if (!s->launched) continue;
        MMS_CUDA(cudaSetDevice(s->device));
        std::vector<int32_t> flags(2 * s->pipe.chunks.size(), 0);
        MMS_CUDA(cudaMemcpy(flags.data(), s->d_error, flags.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < s->pipe.chunks.size(); ++k)
            if (flags[2 * k])
                return fail(MMS_ERR_NONFINITE, "non-finite (NA/NaN/Inf) monthly forcing in cells %lld..%lld; R's approx() "
                            "would drop those knots (na.rm = TRUE), the GPU backend refuses them",
                            (long long)s->cell0, (long long)(s->cell0 + s->n_cells - 1));
    }
    return MMS_OK;
}

int ensure_terra_staging(mms_chain_plan *pl)
{
    for (auto &sp : pl->shards) {
        Shard *s = sp.get();
        if (s->d_terra) continue;
        MMS_CUDA(cudaSetDevice(s->device));
        // [6][M][cs] raw variables, [12][cs] clear-sky values, then n_days [M] and month_id [M]
        const size_t bytes = sizeof(double) * ((size_t)(MMS_NTERRACLIMATE * pl->M + 12) * s->c_stride + pl->M) +
                             sizeof(int32_t) * pl->M;
        MMS_CUDA(cudaMalloc(&s->d_terra, bytes));
    }
    return MMS_OK;
}

void fill_inputs(mms_chain_plan *pl, std::vector<PipeJob> &jobs, PipeKind kind, const double *monthly,
                 const double *terraclimate, const double *clear_sky, const int32_t *month_id, const double *n_days,
                 const double *altitude)
{
    jobs.assign(pl->shards.size(), PipeJob());
    const size_t C = (size_t)pl->C;
    for (size_t g = 0; g < pl->shards.size(); ++g) {
        Shard *s = pl->shards[g].get();
        PipeJob &j = jobs[g];
        j.kind = kind;
        const bool tc = terraclimate != nullptr;
        if (tc) {
            double *d_cs = s->d_terra + (size_t)MMS_NTERRACLIMATE * pl->M * s->c_stride;
            j.in.push_back(make_mat(terraclimate, C, s->d_terra, s, (size_t)MMS_NTERRACLIMATE * pl->M, 8));
            j.in.push_back(make_mat(clear_sky, C, d_cs, s, 12, 8));
            // the two small per-month vectors ride in front of worker 0's first copy (every chunk's kernels wait
            // for worker 0's event of that chunk, recorded later on the same stream)
            j.before_first_copy = [s, pl, n_days, month_id](cudaStream_t st) {
                double *d_nd = s->d_terra + (size_t)(MMS_NTERRACLIMATE * pl->M + 12) * s->c_stride;
                int32_t *d_mid = reinterpret_cast<int32_t *>(d_nd + pl->M);
                MMS_CUDA_W(s, cudaMemcpyAsync(d_nd, n_days, sizeof(double) * pl->M, cudaMemcpyHostToDevice, st));
                MMS_CUDA_W(s, cudaMemcpyAsync(d_mid, month_id, sizeof(int32_t) * pl->M, cudaMemcpyHostToDevice, st));
                return true;
            };
        } else {
            j.in.push_back(make_mat(monthly, C, s->d_monthly, s, (size_t)MMS_NFORCING * pl->M, 8));
        }
        j.in.push_back(make_mat(altitude, C, s->d_altitude, s, 1, 8));
        // what the fused kernel reads is derived from the rows of the chunk as soon as they are on the device ...
        j.inputs_ready = [s, tc](int k) {
            const PipeChunk &c = s->pipe.chunks[k];
            return enqueue_transform(s, tc, c.c0, c.n_cols, c.n_cells);
        };
        // ... and, in a pipelined run, the chunk's fused kernel follows
        j.run_chunk = [s](int k) {
            const PipeChunk &c = s->pipe.chunks[k];
            return enqueue_chain(s, c.c0, c.n_cols, c.n_cells, k);
        };
    }
}

// chunks of one shard: at most one round of the persistent kernel each (every warp slot of the device gets one
// unit; a chunk that needed a second, partly filled round would run at about half the warps per SM), equal numbers
// of 32-cell units, at most kMaxChunks (then each chunk is several even rounds)
void plan_chunks(mms_chain_plan *pl, Shard *s, int n_sm_guess)
{
    const int64_t blocks = s->c_stride / 32;
    const int kns = mms::chain_settings_per_thread(pl->S);
    const int64_t units = blocks * (pl->S / kns);
    const bool strict = pl->math_mode != 0 || !pl->tab.uniform;
    const int64_t round_units = mms::chain_round_units(pl->S, strict, n_sm_guess);
    int64_t n = (units + round_units - 1) / round_units;
    if (const char *e = std::getenv("MMS_CHAIN_CHUNKS")) {
        const int64_t f = std::atoll(e);
        if (f >= 1) n = f;
    }
    n = std::max<int64_t>(1, std::min<int64_t>(n, std::min<int64_t>(kMaxChunks, blocks)));
    s->pipe.chunks.clear();
    for (int64_t k = 0; k < n; ++k) {
        const int64_t b0 = blocks * k / n, b1 = blocks * (k + 1) / n;
        PipeChunk c;
        c.c0 = b0 * 32;
        c.n_cols = (b1 - b0) * 32;
        c.n_cells = std::max<int64_t>(0, std::min<int64_t>(s->n_cells - c.c0, c.n_cols));
        s->pipe.chunks.push_back(c);
    }
}

}  // namespace

extern "C" {

void mms_chain_destroy(mms_chain_plan *plan)
{
    if (!plan) return;
    for (auto &sp : plan->shards) {
        sp->pipe.stop();
        free_shard_device(sp.get());
    }
    delete plan;
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

    std::unique_ptr<mms_chain_plan> plu(new mms_chain_plan());
    mms_chain_plan *pl = plu.get();
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
        std::vector<char> seen(pl->P, 0);
        int last = -1;
        for (int d = 0; d < pl->D && !rc; ++d) {
            const int q = pl->period[d];
            if (q >= pl->P) rc = fail(MMS_ERR_INVALID_ARG, "period_of_day[%d] = %d >= n_periods", d, q);
            else if (q >= 0 && q < last) rc = fail(MMS_ERR_INVALID_ARG, "period_of_day must be non-decreasing over its non-negative entries");
            if (q >= 0 && q < pl->P) { last = q; seen[q] = 1; }
        }
        pl->all_periods_present = std::all_of(seen.begin(), seen.end(), [](char c) { return c != 0; });
    }
    if (!rc) rc = pack_tables(cfg->lifehistory, pl->tab);
    if (!rc && cfg->water_temperature.enabled) {
        // EXTENSION: the water-temperature model
        const mms_water_temperature &w = cfg->water_temperature;
        if (pl->S != 1)
            rc = fail(MMS_ERR_UNSUPPORTED, "model_water_temperature needs n_settings == 1: the rates that take the water "
                      "temperature become a property of the water body (run one habitat setting per call)");
        else if (!w.day_of_year || !w.month_of_day) rc = fail(MMS_ERR_INVALID_ARG, "water temperature: NULL day_of_year or month_of_day");
        else if (!(w.shade_proportion >= 0 && w.shade_proportion <= 1)) rc = fail(MMS_ERR_INVALID_ARG, "water_shade_proportion must be in [0, 1]");
        else if (!(w.albedo >= 0 && w.albedo <= 1) || !(w.emissivity > 0 && w.emissivity <= 1) || !(w.convection_w_m2_k >= 0))
            rc = fail(MMS_ERR_INVALID_ARG, "water temperature: albedo in [0, 1], emissivity in (0, 1], convection >= 0");
        pl->wt = w;
        pl->wt.day_of_year = nullptr; pl->wt.month_of_day = nullptr;     // copied below; the caller's arrays are not kept
        for (int d = 0; d < pl->D && !rc; ++d) {
            const int doy = w.day_of_year[d], mon = w.month_of_day[d];
            if (doy < 1 || doy > 366 || mon < 1 || mon > 12) { rc = fail(MMS_ERR_INVALID_ARG, "water temperature: day_of_year[%d] / month_of_day[%d] out of range", d, d); break; }
            // zeta = omega (doy - 80) + 2 eps (sin(omega doy) - sin(omega 80)), delta = asin(0.39784993 sin zeta)
            // (query_terraclimate.R:434-457), host libm
            const double omega = 2 * 3.141592653589793 / 365, eps = 0.01675;
            const double zeta = omega * (doy - 80) + 2 * eps * (std::sin(omega * doy) - std::sin(omega * 80));
            const double delta = std::asin(0.39784993 * std::sin(zeta));
            pl->sun_day.push_back(make_double2(std::sin(delta), std::cos(delta)));
            pl->month_of_day.push_back(mon);
        }
        for (int k = 1; k <= 24; ++k) pl->hr.cos_hour_angle[k - 1] = std::cos(15.0 * (k - 12) * (3.141592653589793 / 180));   // :459-466
    }
    pl->R = pl->C;
    if (!rc && cfg->raster_cell_idx) {
        // N2: scatter the output rows into raster cell order (cell_idx of R/pipeline_functions.R:1316-1322)
        pl->scatter = true;
        pl->R = cfg->n_raster_cells;
        pl->cell_idx.assign(cfg->raster_cell_idx, cfg->raster_cell_idx + pl->C);
        if (pl->R < pl->C) rc = fail(MMS_ERR_INVALID_ARG, "n_raster_cells (%lld) is smaller than n_cells", (long long)pl->R);
        for (int64_t c = 0; c < pl->C && !rc; ++c) {
            const int64_t k = pl->cell_idx[(size_t)c];
            if (k < 0 || k >= pl->R) rc = fail(MMS_ERR_INVALID_ARG, "raster_cell_idx[%lld] = %lld is outside 0..n_raster_cells-1", (long long)c, (long long)k);
            else if (c > 0 && k <= pl->cell_idx[(size_t)c - 1])
                rc = fail(MMS_ERR_UNSUPPORTED, "raster_cell_idx must be strictly increasing (pixels in raster cell order, as "
                          "terra::cells() returns them); sort the tile's pixels by cell number");
        }
    }
    if (rc) return rc;
    pl->ds_nx = cfg->lifehistory.ds_temp_humid.nx; pl->ds_ny = cfg->lifehistory.ds_temp_humid.ny;
    pl->gx.assign(cfg->lifehistory.ds_temp_humid.x, cfg->lifehistory.ds_temp_humid.x + pl->ds_nx);
    pl->gy.assign(cfg->lifehistory.ds_temp_humid.y, cfg->lifehistory.ds_temp_humid.y + pl->ds_ny);

    // contiguous cell ranges, one per device; each shard padded to whole 128-cell blocks
    const int G = (int)devs.size();
    int n_sm_guess = 148;
    cudaDeviceGetAttribute(&n_sm_guess, cudaDevAttrMultiProcessorCount, devs[0]);
    for (int g = 0; g < G; ++g) {
        int64_t cell0 = 0, n = 0;
        mms_shard_range(pl->C, G, g, &cell0, &n);
        if (n <= 0) continue;
        std::unique_ptr<Shard> s(new Shard());
        s->plan = pl;
        s->device = devs[g];
        s->pipe.device = devs[g];
        s->cell0 = cell0;
        s->n_cells = n;
        s->c_stride = ((n + mms::kBlock - 1) / mms::kBlock) * mms::kBlock;
        s->out_col0 = out_boundary(pl, cell0);
        s->out_cols = out_boundary(pl, cell0 + n) - s->out_col0;
        s->out_stride = pl->scatter ? std::max<int64_t>(32, ((s->out_cols + 31) / 32) * 32) : s->c_stride;
        plan_chunks(pl, s.get(), n_sm_guess);
        pl->shards.push_back(std::move(s));
    }
    pl->n_workers = default_workers((int)pl->shards.size());
    for (auto &sp : pl->shards) sp->pipe.start(pl->n_workers);
    // device memory, tables and events of every shard, in parallel (one CUDA context each)
    std::vector<PipeJob> jobs(pl->shards.size());
    for (size_t g = 0; g < jobs.size(); ++g) {
        Shard *s = pl->shards[g].get();
        jobs[g].kind = PIPE_CALL;
        jobs[g].call = [s] { return setup_shard(s); };
    }
    rc = dispatch(pl, jobs);
    if (rc) { mms_chain_destroy(plu.release()); return rc; }
    *out = plu.release();
    return MMS_OK;
}

int mms_chain_upload(mms_chain_plan *pl, const double *monthly, const double *altitude)
{
    if (!pl || !monthly || !altitude) return fail(MMS_ERR_INVALID_ARG, "NULL plan, monthly or altitude");
    std::vector<PipeJob> jobs;
    fill_inputs(pl, jobs, PIPE_UPLOAD, monthly, nullptr, nullptr, nullptr, nullptr, altitude);
    classify(pl, jobs);
    const int rc = dispatch(pl, jobs);
    if (rc) return rc;
    pl->uploaded = true;
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
    if ((rc = ensure_terra_staging(pl))) return rc;
    std::vector<PipeJob> jobs;
    fill_inputs(pl, jobs, PIPE_UPLOAD, nullptr, terraclimate, clear_sky, month_id, n_days, altitude);
    classify(pl, jobs);
    if ((rc = dispatch(pl, jobs))) return rc;
    pl->uploaded = true;
    return MMS_OK;
}

// latitude and clear-sky radiation of every cell (water-temperature EXTENSION); sin / cos of the latitude are taken
// here with the host libm.  Small (14 doubles per cell): one synchronous copy per shard.
int mms_chain_set_location(mms_chain_plan *pl, const double *latitude, const double *clear_sky)
{
    if (!pl || !latitude || !clear_sky) return fail(MMS_ERR_INVALID_ARG, "NULL plan, latitude or clear_sky");
    if (!pl->wt.enabled) return fail(MMS_ERR_INVALID_ARG, "the plan was not created with the water-temperature model enabled");
    for (auto &sp : pl->shards) {
        Shard *s = sp.get();
        MMS_CUDA(cudaSetDevice(s->device));
        std::vector<double> sl((size_t)s->n_cells), cl((size_t)s->n_cells);
        for (int64_t c = 0; c < s->n_cells; ++c) {
            const double lat = latitude[s->cell0 + c];
            if (!(lat >= -90 && lat <= 90)) return fail(MMS_ERR_INVALID_ARG, "latitude of cell %lld is not in [-90, 90]", (long long)(s->cell0 + c));
            const double phi = lat * (3.141592653589793 / 180);
            sl[(size_t)c] = std::sin(phi);
            cl[(size_t)c] = std::cos(phi);
        }
        MMS_CUDA(cudaMemcpy(s->d_sin_lat, sl.data(), sizeof(double) * s->n_cells, cudaMemcpyHostToDevice));
        MMS_CUDA(cudaMemcpy(s->d_cos_lat, cl.data(), sizeof(double) * s->n_cells, cudaMemcpyHostToDevice));
        MMS_CUDA(cudaMemcpy2D(s->d_clear_sky, sizeof(double) * s->c_stride, clear_sky + s->cell0, sizeof(double) * pl->C,
                              sizeof(double) * s->n_cells, 12, cudaMemcpyHostToDevice));
    }
    pl->location_set = true;
    return MMS_OK;
}

int mms_chain_launch(mms_chain_plan *pl)
{
    if (!pl) return fail(MMS_ERR_INVALID_ARG, "NULL plan");
    if (!pl->uploaded) return fail(MMS_ERR_INVALID_ARG, "mms_chain_upload has not been called on this plan");
    if (pl->wt.enabled && !pl->location_set) return fail(MMS_ERR_INVALID_ARG, "the water-temperature model needs mms_chain_set_location() first");
    for (auto &sp : pl->shards) {
        Shard *s = sp.get();
        MMS_CUDA(cudaSetDevice(s->device));
        s->pipe.clear_error();
        MMS_CUDA(cudaEventRecord(s->ev0, s->pipe.s_compute));
        // resident data: one launch over the whole shard
        if (!enqueue_chain(s, 0, s->c_stride, s->n_cells, 0)) {
            std::string msg;
            const int r = s->pipe.wait(&msg);
            return fail_str(r ? r : MMS_ERR_CUDA, msg);
        }
        MMS_CUDA(cudaEventRecord(s->ev1, s->pipe.s_compute));
        for (PipeChunk &c : s->pipe.chunks) {
            MMS_CUDA(cudaEventRecord(c.ev_kernel, s->pipe.s_compute));
            c.kernel_recorded = true;
        }
        s->timing_chunks = false;
        s->launched = true;
    }
    return MMS_OK;
}

int mms_chain_sync(mms_chain_plan *pl)
{
    if (!pl) return fail(MMS_ERR_INVALID_ARG, "NULL plan");
    for (auto &sp : pl->shards) {
        MMS_CUDA(cudaSetDevice(sp->device));
        MMS_CUDA(cudaStreamSynchronize(sp->pipe.s_compute));
    }
    return check_error_flags(pl);
}

int mms_chain_download(mms_chain_plan *pl, const mms_chain_outputs *out)
{
    if (!pl || !out) return fail(MMS_ERR_INVALID_ARG, "NULL plan or outputs");
    std::vector<PipeJob> jobs(pl->shards.size());
    for (size_t g = 0; g < pl->shards.size(); ++g) {
        jobs[g].kind = PIPE_DOWNLOAD;
        add_outputs(pl, pl->shards[g].get(), out, jobs[g]);
    }
    classify(pl, jobs);
    const int rc = dispatch(pl, jobs);
    if (rc) return rc;
    return mms_chain_sync(pl);
}

static int run_pipelined(mms_chain_plan *pl, std::vector<PipeJob> &jobs, const mms_chain_outputs *out)
{
    if (pl->wt.enabled && !pl->location_set) return fail(MMS_ERR_INVALID_ARG, "the water-temperature model needs mms_chain_set_location() first");
    for (size_t g = 0; g < pl->shards.size(); ++g) add_outputs(pl, pl->shards[g].get(), out, jobs[g]);
    classify(pl, jobs);
    for (auto &sp : pl->shards) { sp->timing_chunks = true; sp->launched = true; }
    const int rc = dispatch(pl, jobs);
    if (rc) return rc;
    pl->uploaded = true;
    return mms_chain_sync(pl);
}

int mms_chain_run(mms_chain_plan *pl, const double *monthly, const double *altitude, const mms_chain_outputs *out)
{
    if (!pl || !monthly || !altitude || !out) return fail(MMS_ERR_INVALID_ARG, "NULL plan, monthly, altitude or outputs");
    std::vector<PipeJob> jobs;
    fill_inputs(pl, jobs, PIPE_RUN, monthly, nullptr, nullptr, nullptr, nullptr, altitude);
    return run_pipelined(pl, jobs, out);
}

int mms_chain_run_terraclimate(mms_chain_plan *pl, const double *terraclimate, const double *clear_sky,
                               const int32_t *month_id, const double *n_days, const double *altitude,
                               const mms_chain_outputs *out)
{
    if (!pl || !terraclimate || !clear_sky || !month_id || !n_days || !altitude || !out)
        return fail(MMS_ERR_INVALID_ARG, "NULL plan or argument");
    int rc = check_month_ids(month_id, n_days, pl->M);
    if (rc) return rc;
    if ((rc = ensure_terra_staging(pl))) return rc;
    std::vector<PipeJob> jobs;
    fill_inputs(pl, jobs, PIPE_RUN, nullptr, terraclimate, clear_sky, month_id, n_days, altitude);
    return run_pipelined(pl, jobs, out);
}

int mms_chain_kernel_ms(mms_chain_plan *pl, float *ms)
{
    if (!pl || !ms) return fail(MMS_ERR_INVALID_ARG, "NULL plan or ms");
    float worst = 0.f;
    for (auto &sp : pl->shards) {
        Shard *s = sp.get();
        if (!s->launched) return fail(MMS_ERR_INVALID_ARG, "no launch recorded");
        MMS_CUDA(cudaSetDevice(s->device));
        float t = 0.f;
        if (s->timing_chunks) {
            for (PipeChunk &c : s->pipe.chunks) {
                if (c.n_cols <= 0 || !c.kernel_recorded) continue;
                float tc = 0.f;
                MMS_CUDA(cudaEventSynchronize(c.ev_t1));
                MMS_CUDA(cudaEventElapsedTime(&tc, c.ev_t0, c.ev_t1));
                t += tc;
            }
        } else {
            MMS_CUDA(cudaEventSynchronize(s->ev1));
            MMS_CUDA(cudaEventElapsedTime(&t, s->ev0, s->ev1));
        }
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MMS_OK;
}

int mms_chain_launch_count(mms_chain_plan *pl, int64_t *n)
{
    if (!pl || !n) return fail(MMS_ERR_INVALID_ARG, "NULL plan or n");
    *n = pl->launches.load();
    return MMS_OK;
}

int mms_chain_plan_info(mms_chain_plan *pl, int32_t *n_shards, int32_t *n_chunks_per_shard, int32_t *n_workers_per_shard)
{
    if (!pl) return fail(MMS_ERR_INVALID_ARG, "NULL plan");
    if (n_shards) *n_shards = (int32_t)pl->shards.size();
    if (n_chunks_per_shard) *n_chunks_per_shard = pl->shards.empty() ? 0 : (int32_t)pl->shards[0]->pipe.chunks.size();
    if (n_workers_per_shard) *n_workers_per_shard = pl->n_workers;
    return MMS_OK;
}

int mms_run_chain(const mms_chain_config *cfg, const double *monthly, const double *altitude,
                  const mms_chain_outputs *out)
{
    mms_chain_plan *pl = nullptr;
    int rc = mms_chain_create(cfg, &pl);
    if (rc) return rc;
    rc = mms_chain_run(pl, monthly, altitude, out);
    mms_chain_destroy(pl);
    return rc;
}

int mms_run_chain_terraclimate(const mms_chain_config *cfg, const double *terraclimate, const double *clear_sky,
                               const int32_t *month_id, const double *n_days, const double *altitude,
                               const mms_chain_outputs *out)
{
    mms_chain_plan *pl = nullptr;
    int rc = mms_chain_create(cfg, &pl);
    if (rc) return rc;
    rc = mms_chain_run_terraclimate(pl, terraclimate, clear_sky, month_id, n_days, altitude, out);
    mms_chain_destroy(pl);
    return rc;
}

}  // extern "C"
