// stage_kernels.cuh -- matrix-in / matrix-out twins of the reference's two
// vectorised simulators, on R's own time-by-pixel column-major layout
// (element (t, c) at [c*H + t], i.e. each pixel's series is contiguous):
//   simulate_ephemeral_habitat_vectorised()  R/sim_water.R:216-301
//   simulate_population_vectorised()         R/sim_pop_dynamics.R:208-284
//
// One thread integrates one pixel.  Because R's layout is [pixel][time], a
// warp reading "the same t for adjacent pixels" would be strided by H*8 bytes;
// instead each CTA stages a tile of kStageCells pixels x kStageT time steps per
// input through shared memory with cp.async (coalesced 64-byte row segments,
// transposed on the shared-memory side), integrates the kStageT steps from
// shared memory, and writes the output tile back the same way.
//
// Included by chain_fast.cu / chain_strict.cu.  Citations: reference root.
#pragma once

#include <cuda_pipeline.h>

#include "mms_math.cuh"

namespace mms {
namespace MMS_MODE {

constexpr int kStageCells = 128;   // pixels per CTA (= threads)
constexpr int kStageT = 8;         // time steps per tile: 64-byte row segments
constexpr int kStagePitch = kStageT + 2;   // doubles; keeps 16-byte alignment, spreads banks
constexpr size_t kStageTileBytes = sizeof(double) * kStageCells * kStagePitch;
constexpr size_t kWaterStageSmem = 4 * kStageTileBytes + kStageCells * kStageT;   // + wet flags
constexpr size_t kPopStageSmem = 5 * kStageTileBytes;

// Copy a [cells x kStageT] tile of a column-major H-by-C matrix into smem
// tile[cell][kStagePitch].  nt <= kStageT valid steps, nc <= kStageCells valid pixels.
template <int kVecDoubles>
__device__ __forceinline__ void stage_load(double *tile, const double *__restrict__ mat, int64_t H,
                                           int64_t cell0, int nc, int64_t t0, int nt)
{
    constexpr int kChunks = kStageT / kVecDoubles;            // chunks per row
    for (int k = threadIdx.x; k < kStageCells * kChunks; k += blockDim.x) {
        const int row = k / kChunks, chunk = k % kChunks;
        const int t = chunk * kVecDoubles;
        if (row < nc && t < nt) {
            const double *src = mat + (cell0 + row) * H + t0 + t;
            double *dst = tile + row * kStagePitch + t;
            if (kVecDoubles == 2 && t + 1 < nt) {
                __pipeline_memcpy_async(dst, src, 16);
            } else {
                __pipeline_memcpy_async(dst, src, 8);
                if (kVecDoubles == 2 && t + 1 < nt) __pipeline_memcpy_async(dst + 1, src + 1, 8);
            }
        }
    }
}

template <int kVecDoubles>
__device__ __forceinline__ void stage_store(const double *tile, double *__restrict__ mat, int64_t H,
                                            int64_t cell0, int nc, int64_t t0, int nt)
{
    constexpr int kChunks = kStageT / kVecDoubles;
    for (int k = threadIdx.x; k < kStageCells * kChunks; k += blockDim.x) {
        const int row = k / kChunks, chunk = k % kChunks;
        const int t = chunk * kVecDoubles;
        if (row < nc && t < nt) {
            double *dst = mat + (cell0 + row) * H + t0 + t;
            const double *src = tile + row * kStagePitch + t;
            if (kVecDoubles == 2 && t + 1 < nt) {
                *reinterpret_cast<double2 *>(dst) = *reinterpret_cast<const double2 *>(src);
            } else {
                dst[0] = src[0];
            }
        }
    }
}

// ---- simulate_ephemeral_habitat_vectorised() --------------------------------
template <int kVecDoubles>
__global__ void __launch_bounds__(kStageCells)
water_stage_kernel(const __grid_constant__ WaterStageParams p)
{
    // dynamic smem: 4 input tiles; the two output tiles alias inputs 0 and 1
    // (each thread reads all inputs of step t before it writes the outputs of step t)
    extern __shared__ __align__(16) unsigned char stage_smem[];
    double (*s_in)[kStageCells * kStagePitch] =
        reinterpret_cast<double (*)[kStageCells * kStagePitch]>(stage_smem);
    double *s_area = s_in[0];
    double *s_vol = s_in[1];
    uint8_t *s_wet = reinterpret_cast<uint8_t *>(s_in[4]);

    const int64_t cell0 = (int64_t)blockIdx.x * kStageCells;
    const int nc = (int)min((int64_t)kStageCells, p.C - cell0);
    const int me = threadIdx.x;
    const bool live = me < nc;
    const double p_a = live ? air_pressure_kpa(p.altitude[cell0 + me]) : 101.325;
    ::mms::fm::math_tables_init();
    __syncthreads();

    double volume = p.hab.initial_volume;                                  // R/sim_water.R:277
    double surface = habitat_surface(p.hab, volume);
    for (int pass = 0; pass <= p.burnin_years; ++pass) {                   // index, :232-233
        const bool final_pass = pass == p.burnin_years;
        const int64_t n_steps = final_pass ? p.H : (int64_t)kHoursPerYear;
        for (int64_t t0 = 0; t0 < n_steps; t0 += kStageT) {
            const int nt = (int)min((int64_t)kStageT, n_steps - t0);
            stage_load<kVecDoubles>(s_in[0], p.rainfall, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[1], p.air_temperature, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[2], p.humidity, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[3], p.windspeed, p.H, cell0, nc, t0, nt);
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncthreads();
            if (live) {
                for (int t = 0; t < nt; ++t) {
                    const int k = me * kStagePitch + t;
                    const double rain = s_in[0][k], ta = s_in[1][k];
                    const double rate = evaporation_rate(ta, s_in[2][k], p_a, s_in[3][k]);           // :264-270
                    const double gain = p.hab.catchment_area * MMS_DIVC(rain, 1000.0);               // :122-123
                    volume = iterate_volume(p.hab, volume, surface, rate, gain);                     // :279-287
                    surface = habitat_surface(p.hab, volume);                                        // :294
                    s_area[k] = surface;
                    s_vol[k] = volume;
                    s_wet[me * kStageT + t] = (volume > 0.0) ? 1 : 0;
                }
            }
            __syncthreads();
            if (final_pass) {                                                                        // :290-291
                stage_store<kVecDoubles>(s_area, p.water_surface_area, p.H, cell0, nc, t0, nt);
                if (p.water_volume) stage_store<kVecDoubles>(s_vol, p.water_volume, p.H, cell0, nc, t0, nt);
                if (p.wet) {
                    for (int k = threadIdx.x; k < kStageCells * kStageT; k += blockDim.x) {
                        const int row = k / kStageT, t = k % kStageT;
                        if (row < nc && t < nt) p.wet[(cell0 + row) * p.H + t0 + t] = s_wet[k];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---- simulate_population_vectorised() ---------------------------------------
template <int kVecDoubles>
__global__ void __launch_bounds__(kStageCells)
pop_stage_kernel(const __grid_constant__ PopStageParams p)
{
    // dynamic smem: 5 input tiles; the two output tiles alias inputs 0 and 1
    extern __shared__ __align__(16) unsigned char stage_smem[];
    double (*s_in)[kStageCells * kStagePitch] =
        reinterpret_cast<double (*)[kStageCells * kStagePitch]>(stage_smem);
    double *s_aq = s_in[0];
    double *s_ad = s_in[1];

    const int64_t cell0 = (int64_t)blockIdx.x * kStageCells;
    const int nc = (int)min((int64_t)kStageCells, p.C - cell0);
    const int me = threadIdx.x;
    const bool live = me < nc;

    double aquatic = p.initial_aquatic, adult = p.initial_adult;           // R/sim_pop_dynamics.R:249-252
    ::mms::fm::math_tables_init();
    __syncthreads();
    for (int pass = 0; pass <= p.burnin_years; ++pass) {                   // :225-226
        const bool final_pass = pass == p.burnin_years;
        const int64_t n_steps = final_pass ? p.H : (int64_t)kHoursPerYear;
        for (int64_t t0 = 0; t0 < n_steps; t0 += kStageT) {
            const int nt = (int)min((int64_t)kStageT, n_steps - t0);
            stage_load<kVecDoubles>(s_in[0], p.larval_habitat_area, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[1], p.mdr, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[2], p.efd, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[3], p.das_zerodensity, p.H, cell0, nc, t0, nt);
            stage_load<kVecDoubles>(s_in[4], p.ds, p.H, cell0, nc, t0, nt);
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncthreads();
            if (live) {
                for (int t = 0; t < nt; ++t) {
                    const int k = me * kStagePitch + t;
                    const double area = s_in[0][k];
                    const PopHour ph = pop_hour(p.densmod_type, s_in[1][k], s_in[2][k], s_in[3][k], s_in[4][k]);
                    iterate_state(aquatic, adult, ph, p.densmod_type, p.dd_effect, p.surface_area_cm2,
                                  area);                                                              // :259-268
                    s_aq[k] = aquatic;
                    s_ad[k] = adult;
                }
            }
            __syncthreads();
            if (final_pass) {                                                                         // :276-282
                stage_store<kVecDoubles>(s_aq, p.aquatic, p.H, cell0, nc, t0, nt);
                stage_store<kVecDoubles>(s_ad, p.adult, p.H, cell0, nc, t0, nt);
            }
            __syncthreads();
        }
    }
}

}  // namespace MMS_MODE
}  // namespace mms
