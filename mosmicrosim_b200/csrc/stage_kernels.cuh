// stage_kernels.cuh -- matrix-in / matrix-out twins of the reference's two
// vectorised simulators, on R's own time-by-pixel column-major layout
// (element (t, c) at [c*H + t], i.e. each pixel's series is contiguous):
//   simulate_ephemeral_habitat_vectorised()  R/sim_water.R:216-301
//   simulate_population_vectorised()         R/sim_pop_dynamics.R:208-284
//
// One thread integrates one pixel; the work unit is a warp = 32 pixels, and warps
// never synchronise with each other.  Because R's layout is [pixel][time], a warp
// reading "the same t for adjacent pixels" would be strided by H*8 bytes; instead
// each warp lands a tile of 32 pixels x kStageT steps per input in shared memory
// with cp.async (whole 32-byte sectors per row), double-buffered so that the
// copies of tile k+1 are in flight while tile k is integrated.  Each lane then
// reads its own row (swizzled, conflict-free LDS.128) and writes its kStageT
// outputs per variable as one 32-byte store = one whole sector of its series.
//
// Each step costs an exp / log / pow or three, so with the copies hidden the
// kernels are FP64-arithmetic-bound, not HBM-bound: the CTA is persistent (one per
// SM) and the host picks the warp count (chain_tu.cu, stage_plan) so that the
// units run in evenly filled rounds with a 128-register budget per thread
// (a 100k-pixel tile: two rounds of 11 warps per SM).
//
// Included by chain_tu.cu.  Citations: reference root.
#pragma once

#include <cuda_pipeline.h>

#include "mms_math.cuh"

namespace mms {
namespace MMS_MODE {

constexpr int kStageT = 4;                       // steps per tile: one 32-byte sector per row
constexpr int kStageTile = 32 * kStageT;         // doubles per (input, buffer) tile of one warp
#if MMS_STRICT
constexpr int kStageMaxWarps = 12;               // CUDA libm paths need 168 registers
#else
constexpr int kStageMaxWarps = 16;               // 128 registers
#endif
constexpr int kStageFewWarps = 12;               // second instantiation: 168 registers (more hoisting, no spills)
constexpr int kWaterInputs = 4, kPopInputs = 5;
constexpr int kPopChunk = 2;                     // steps whose state-independent half is evaluated together (4: 38.3 vs 31.6 ms)
constexpr size_t kStageSmemBudget = 200 * 1024;  // dynamic shared memory of a CTA (227 KB minus the static math tables -- 20.5 KB when replicated -- and a reserve)
constexpr size_t stage_smem_bytes(int n_inputs, int warps)
{
    return sizeof(double) * 2 * (size_t)n_inputs * kStageTile * (size_t)warps;
}

// Row r of a tile holds its 4 steps as two 16-byte chunks; chunk c sits at position c ^ ((r >> 2) & 1),
// so that the 8 lanes of a quarter-warp reading "their" chunk c hit 8 distinct 4-bank groups.
__device__ __forceinline__ int stage_offset(int row, int t)
{
    return row * kStageT + ((((t >> 1) ^ (row >> 2)) & 1) << 1) + (t & 1);
}

// Issue the copies of one warp tile: steps t0 .. t0+nt-1 of pixels cell0 .. cell0+nc-1.
// kVec >= 2: 16-byte copies, two lanes per row (each instruction covers 16 whole sectors);
// kVec == 1: 8-byte copies for series that are not 16-byte aligned (odd H).
template <int kVec>
__device__ __forceinline__ void stage_issue(double *tile, const double *__restrict__ mat, int64_t H, int64_t cell0,
                                            int nc, int64_t t0, int nt, int lane)
{
    if (kVec >= 2) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = j * 16 + (lane >> 1), t = (lane & 1) * 2;
            if (row < nc && t < nt) {
                const double *src = mat + (cell0 + row) * H + t0 + t;
                double *dst = tile + stage_offset(row, t);
                if (t + 1 < nt) __pipeline_memcpy_async(dst, src, 16);
                else __pipeline_memcpy_async(dst, src, 8);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = j * 8 + (lane >> 2), t = lane & 3;
            if (row < nc && t < nt)
                __pipeline_memcpy_async(tile + stage_offset(row, t), mat + (cell0 + row) * H + t0 + t, 8);
        }
    }
}

// A lane's own row of a tile.
__device__ __forceinline__ void stage_row(const double *tile, int lane, double (&v)[kStageT])
{
    const double2 a = *reinterpret_cast<const double2 *>(tile + stage_offset(lane, 0));
    const double2 b = *reinterpret_cast<const double2 *>(tile + stage_offset(lane, 2));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// Store a lane's kStageT outputs: dst = &out[cell*H + t0].
template <int kVec>
__device__ __forceinline__ void stage_put(double *dst, const double (&v)[kStageT], int nt)
{
    if (kVec == 4 && nt == kStageT) {
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                     : "memory");
    } else if (kVec >= 2) {
        if (nt >= 2) *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
        else if (nt == 1) dst[0] = v[0];
        if (nt >= 4) *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
        else if (nt == 3) dst[2] = v[2];
    } else {
#pragma unroll
        for (int t = 0; t < kStageT; ++t)
            if (t < nt) dst[t] = v[t];
    }
}

// ---- simulate_ephemeral_habitat_vectorised() --------------------------------
// kVec: 4 = H % 4 == 0 and all matrices 32-byte aligned; 2 = H even, 16-byte aligned; 1 = anything.
// kShape: the water-body shape (0 cone, 1 cylinder), a template parameter so that the other shape's code and
// the branch around it leave the recursion.
// kPx: pixels per thread (a unit is 32 * kPx adjacent pixels; lane l integrates pixels l, l + 32, ... of it as
// kPx independent recursions).  Fixed at 1: two pixels per thread measured slower (52.2 vs 31.6 ms on the
// 100k-pixel population twin) -- the second chain costs more in registers and shared memory (half the warps)
// than its interleaving gains.
template <int kVec, int kMaxWarps, int kShape>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
water_stage_kernel(const __grid_constant__ WaterStageParams p)
{
    constexpr int kPx = 1;
    extern __shared__ __align__(16) unsigned char stage_smem[];
    constexpr int kBuf = kPx * kWaterInputs * kStageTile;                  // doubles per buffer of one warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    double *tiles = reinterpret_cast<double *>(stage_smem) + (size_t)warp * 2 * kBuf;
    ::mms::fm::math_tables_init();
    __syncthreads();                                                       // the only block-wide barrier

    const int64_t n_units = (p.C + 32 * kPx - 1) / (32 * kPx);
    for (int64_t unit = blockIdx.x + (int64_t)gridDim.x * warp; unit < n_units; unit += (int64_t)gridDim.x * n_warps) {
        int64_t cell0[kPx], cell[kPx];
        int nc[kPx];
        bool live[kPx];
        double p_a[kPx], volume[kPx], surface[kPx];
#pragma unroll
        for (int q = 0; q < kPx; ++q) {
            cell0[q] = (unit * kPx + q) * 32;
            const int64_t left = p.C - cell0[q];
            nc[q] = (int)(left < 0 ? 0 : (left > 32 ? 32 : left));
            live[q] = lane < nc[q];
            cell[q] = live[q] ? cell0[q] + lane : 0;
            p_a[q] = air_pressure_kpa(p.altitude[cell[q]]);
            volume[q] = p.hab.initial_volume;                              // R/sim_water.R:277
            surface[q] = habitat_surface(p.hab, volume[q], kShape);
        }

        auto issue = [&](int buf, int64_t t0, int nt) {
#pragma unroll
            for (int q = 0; q < kPx; ++q) {
                double *b = tiles + buf * kBuf + q * kWaterInputs * kStageTile;
                stage_issue<kVec>(b + 0 * kStageTile, p.rainfall, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 1 * kStageTile, p.air_temperature, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 2 * kStageTile, p.humidity, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 3 * kStageTile, p.windspeed, p.H, cell0[q], nc[q], t0, nt, lane);
            }
        };

        for (int pass = 0; pass <= p.burnin_years; ++pass) {               // index, :232-233
            const bool final_pass = pass == p.burnin_years;
            const int64_t n_steps = final_pass ? p.H : p.burnin_steps;
            const int64_t n_tiles = (n_steps + kStageT - 1) / kStageT;
            if (n_tiles > 0) issue(0, 0, (int)min((int64_t)kStageT, n_steps));
            __pipeline_commit();
            for (int64_t k = 0; k < n_tiles; ++k) {
                const int64_t t0 = k * kStageT;
                const int nt = (int)min((int64_t)kStageT, n_steps - t0);
                if (k + 1 < n_tiles) issue((int)((k + 1) & 1), t0 + kStageT, (int)min((int64_t)kStageT, n_steps - t0 - kStageT));
                __pipeline_commit();
                __pipeline_wait_prior(1);                                  // tile k has landed
                __syncwarp();
                double rate[kPx][kStageT], gain[kPx][kStageT];
#pragma unroll
                for (int q = 0; q < kPx; ++q) {
                    const double *b = tiles + (int)(k & 1) * kBuf + q * kWaterInputs * kStageTile;
                    double rain[kStageT], ta[kStageT], rh[kStageT], ws[kStageT];
                    stage_row(b + 0 * kStageTile, lane, rain);
                    stage_row(b + 1 * kStageTile, lane, ta);
                    stage_row(b + 2 * kStageTile, lane, rh);
                    stage_row(b + 3 * kStageTile, lane, ws);
#pragma unroll
                    for (int t = 0; t < kStageT; ++t) {                    // state independent: 4 * kPx way parallel
                        rate[q][t] = evaporation_rate(ta[t], rh[t], p_a[q], ws[t]);                      // :264-270
                        gain[q][t] = p.hab.catchment_area * MMS_DIVC(rain[t], 1000.0);                   // :122-123
                    }
                }
                double o_area[kPx][kStageT], o_vol[kPx][kStageT];
                uint32_t wet[kPx];
#pragma unroll
                for (int q = 0; q < kPx; ++q) wet[q] = 0;
#pragma unroll
                for (int t = 0; t < kStageT; ++t) {
#pragma unroll
                    for (int q = 0; q < kPx; ++q) {
                        if (t < nt) {
                            volume[q] = iterate_volume(p.hab, volume[q], surface[q], rate[q][t], gain[q][t]);   // :279-287
                            surface[q] = habitat_surface(p.hab, volume[q], kShape);                                     // :294
                        }
                        o_area[q][t] = surface[q];
                        o_vol[q][t] = volume[q];
                        wet[q] |= (volume[q] > 0.0 ? 1u : 0u) << (8 * t);
                    }
                }
#pragma unroll
                for (int q = 0; q < kPx; ++q) {
                    if (final_pass && live[q]) {                                                         // :290-291
                        const int64_t at = cell[q] * p.H + t0;
                        stage_put<kVec>(p.water_surface_area + at, o_area[q], nt);
                        if (p.water_volume) stage_put<kVec>(p.water_volume + at, o_vol[q], nt);
                        if (p.wet) {
                            if (kVec == 4 && nt == kStageT) {
                                *reinterpret_cast<uint32_t *>(p.wet + at) = wet[q];
                            } else {
#pragma unroll
                                for (int t = 0; t < kStageT; ++t)
                                    if (t < nt) p.wet[at + t] = (uint8_t)((wet[q] >> (8 * t)) & 1u);
                            }
                        }
                    }
                }
                __syncwarp();                                              // buffer k & 1 may be refilled
            }
        }
    }
}

// ---- simulate_population_vectorised() ---------------------------------------
// kDens: the density-dependence form (0 logit, 1 cox_ph), compiled in like kShape above
template <int kVec, int kMaxWarps, int kDens>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
pop_stage_kernel(const __grid_constant__ PopStageParams p)
{
    constexpr int kPx = 1;
    extern __shared__ __align__(16) unsigned char stage_smem[];
    constexpr int kBuf = kPx * kPopInputs * kStageTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    double *tiles = reinterpret_cast<double *>(stage_smem) + (size_t)warp * 2 * kBuf;
    ::mms::fm::math_tables_init();
    __syncthreads();                                                       // the only block-wide barrier

    const int64_t n_units = (p.C + 32 * kPx - 1) / (32 * kPx);
    for (int64_t unit = blockIdx.x + (int64_t)gridDim.x * warp; unit < n_units; unit += (int64_t)gridDim.x * n_warps) {
        int64_t cell0[kPx], cell[kPx];
        int nc[kPx];
        bool live[kPx];
        double aquatic[kPx], adult[kPx];
#pragma unroll
        for (int q = 0; q < kPx; ++q) {
            cell0[q] = (unit * kPx + q) * 32;
            const int64_t left = p.C - cell0[q];
            nc[q] = (int)(left < 0 ? 0 : (left > 32 ? 32 : left));
            live[q] = lane < nc[q];
            cell[q] = live[q] ? cell0[q] + lane : 0;
            aquatic[q] = p.initial_aquatic;                                // R/sim_pop_dynamics.R:249-252
            adult[q] = p.initial_adult;
        }

        auto issue = [&](int buf, int64_t t0, int nt) {
#pragma unroll
            for (int q = 0; q < kPx; ++q) {
                double *b = tiles + buf * kBuf + q * kPopInputs * kStageTile;
                stage_issue<kVec>(b + 0 * kStageTile, p.larval_habitat_area, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 1 * kStageTile, p.mdr, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 2 * kStageTile, p.efd, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 3 * kStageTile, p.das_zerodensity, p.H, cell0[q], nc[q], t0, nt, lane);
                stage_issue<kVec>(b + 4 * kStageTile, p.ds, p.H, cell0[q], nc[q], t0, nt, lane);
            }
        };

        for (int pass = 0; pass <= p.burnin_years; ++pass) {               // :225-226
            const bool final_pass = pass == p.burnin_years;
            const int64_t n_steps = final_pass ? p.H : p.burnin_steps;
            const int64_t n_tiles = (n_steps + kStageT - 1) / kStageT;
            if (n_tiles > 0) issue(0, 0, (int)min((int64_t)kStageT, n_steps));
            __pipeline_commit();
            for (int64_t k = 0; k < n_tiles; ++k) {
                const int64_t t0 = k * kStageT;
                const int nt = (int)min((int64_t)kStageT, n_steps - t0);
                if (k + 1 < n_tiles) issue((int)((k + 1) & 1), t0 + kStageT, (int)min((int64_t)kStageT, n_steps - t0 - kStageT));
                __pipeline_commit();
                __pipeline_wait_prior(1);                                  // tile k has landed
                __syncwarp();
                double o_aq[kPx][kStageT], o_ad[kPx][kStageT];
#pragma unroll
                for (int h = 0; h < kStageT; h += kPopChunk) {             // kPopChunk steps at a time: bounded register use
                    PopHour ph[kPx][kPopChunk];
                    double area[kPx][kPopChunk];
#pragma unroll
                    for (int q = 0; q < kPx; ++q) {
                        const double *b = tiles + (int)(k & 1) * kBuf + q * kPopInputs * kStageTile;
#pragma unroll
                        for (int j = 0; j < kPopChunk; j += 2) {
                            const int at = stage_offset(lane, h + j);
                            const double2 a = *reinterpret_cast<const double2 *>(b + 0 * kStageTile + at);
                            const double2 m = *reinterpret_cast<const double2 *>(b + 1 * kStageTile + at);
                            const double2 e = *reinterpret_cast<const double2 *>(b + 2 * kStageTile + at);
                            const double2 d0 = *reinterpret_cast<const double2 *>(b + 3 * kStageTile + at);
                            const double2 d = *reinterpret_cast<const double2 *>(b + 4 * kStageTile + at);
                            area[q][j] = a.x; area[q][j + 1] = a.y;
                            ph[q][j] = pop_hour(kDens, m.x, e.x, d0.x, d.x);          // state independent
                            ph[q][j + 1] = pop_hour(kDens, m.y, e.y, d0.y, d.y);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kPopChunk; ++j) {
#pragma unroll
                        for (int q = 0; q < kPx; ++q) {
                            if (h + j < nt)
                                iterate_state(aquatic[q], adult[q], ph[q][j], kDens, p.dd_effect,
                                              p.surface_area_cm2, area[q][j]);                        // :259-268
                            o_aq[q][h + j] = aquatic[q];
                            o_ad[q][h + j] = adult[q];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kPx; ++q) {
                    if (final_pass && live[q]) {                                                      // :276-282
                        const int64_t at = cell[q] * p.H + t0;
                        stage_put<kVec>(p.aquatic + at, o_aq[q], nt);
                        stage_put<kVec>(p.adult + at, o_ad[q], nt);
                    }
                }
                __syncwarp();                                              // buffer k & 1 may be refilled
            }
        }
    }
}

}  // namespace MMS_MODE
}  // namespace mms
