// api_common.h -- what the host-side translation units of libmmsgpu share (mmsgpu_api.cu: the fused-chain plan
// and its host pipeline; api_twins.cu: the stage-level twins and the state-free verbs): error reporting, RAII
// holders, and the per-day / per-hour / per-setting constants, derived exactly as the reference derives them
// on the host.  No model arithmetic per cell-hour happens on the host, and there is no CPU fallback.
//
// Citations are relative to the reference root.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mms_common.cuh"

namespace mms_api {

std::string &last_error();          // message of the last failure on the calling thread (api_common.cu)
float &stage_kernel_ms();           // device time of the last stage-twin kernel on the calling thread

int fail(int status, const char *fmt, ...);
int fail_str(int status, const std::string &msg);

#define MMS_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return ::mms_api::fail(e_ == cudaErrorMemoryAllocation ? MMS_ERR_ALLOC : MMS_ERR_CUDA, \
                        "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                        cudaGetErrorString(e_));                                             \
    } while (0)

int device_count_raw();
int pick_device(int32_t device);    // validates the ordinal and makes it current

// a stream that is destroyed on every way out of the function that made it
struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) cudaStreamDestroy(s); }
    operator cudaStream_t() const { return s; }
};

// CUDA-event bracket around one kernel launch on a stream
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

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { MMS_CUDA(cudaMalloc(&p, bytes)); return MMS_OK; }
};

// page-locked host memory (bounce buffers of the host pipeline)
struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~PinBuf() { release(); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    int alloc(size_t n)
    {
        release();
        MMS_CUDA(cudaHostAlloc(&p, n, cudaHostAllocPortable));
        bytes = n;
        return MMS_OK;
    }
};

inline int copy_in_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t st)
{
    if (width == 0 || height == 0) return MMS_OK;
    MMS_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, st));
    return MMS_OK;
}

inline int copy_out_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t st)
{
    if (width == 0 || height == 0) return MMS_OK;
    MMS_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, st));
    return MMS_OK;
}

// true when the driver can DMA straight from / to this host pointer (cudaHostAlloc, cudaHostRegister or
// managed memory); ordinary (pageable) memory -- what R vectors and numpy arrays are -- goes through the
// pipeline's pinned bounce buffers instead
bool host_pointer_is_pinned(const void *p);

// ---- host-side constants -------------------------------------------------------------------

// Hour-of-day tables of interpolate_daily_air_temp() (R/sim_microclimate.R:185-285)
// and vsine_weights() (:431-454), evaluated with the host libm like R does.
void fill_hour_tables(mms::HourTables &hr);

// R's `^` for the two places the host needs it (R_pow: x == 0 -> 0 for y > 0)
double r_pow(double x, double y);

// max_cone_volume and catchment_area, R/sim_water.R:12-31,257-261 (+ cylinder EXTENSION)
int derive_habitat(const mms_habitat &in, mms::HabitatDev &out);

// approx1() interval of every prediction day (base-R approx.c, rule = 1)
int build_day_table(const double *x, int M, const double *v, int D, std::vector<mms::DayEntry> &days);

struct HostTables {
    std::vector<double> sx, scoef;        // spline knots, coefficients [n][3][4]
    int uniform = 0;
    double x0 = 0, step = 0, inv_step = 0;
    std::vector<double> quad;             // ds quad table
};
int pack_tables(const mms_lifehistory &lh, HostTables &t);

// adjust_wind_speed(ws, height_required = 2, height_measured = 10) on "level grass", R/terraclimate.R:615-617,
// R/utils.R:123-132: the factor is evaluated with the host libm, as R does
inline double terraclimate_wind_factor() { return std::pow(2.0 / 10.0, 0.15); }
int check_month_ids(const int32_t *month_id, const double *n_days, int32_t n_months);

inline int count_bits(uint32_t m) { return __builtin_popcount(m); }

}  // namespace mms_api
