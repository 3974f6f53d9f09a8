// chain_ws.cuh -- warp-specialised form of the fused chain (fast math mode only).
//
// Why: the monolithic kernel (chain_kernel.cuh) is latency bound.  B200's FP64 pipe wants four
// independent DFMAs in flight per SM sub-partition (8-cycle latency, one issue per 2 cycles),
// ptxas interleaves at most two independent chains, and ~168 registers per cell limit an SM to
// 11 warps.  Here every 32-cell work unit is served by a PAIR of warps:
//   * the forcing warp runs the state-independent half (monthly -> daily -> hourly
//     microclimate, evaporation rate, life-history tables, ds^(1/24), emergence, logit link)
//     and hands kWsK hours at a time to
//   * the state warp, which runs the recursions (water bodies and populations of the kNS
//     habitat settings) and the summaries of the state variables,
// through a double-buffered shared-memory ring guarded by one named barrier per pair
// (bar.sync id, 64).  The hand-off is one-directional, so the forcing warp runs ahead.  Each
// role needs about half the registers, so twice as many warps -- twice as many independent
// instruction streams -- fit per SM for the same number of cells in flight.
//
// Round 2: the state warp's hour loop is itself software-pipelined (state_pass_pipe: the water step of hour h next to
// the population update of hour h - 1, through a four-slot ring), and the settings of a thread go through the
// stage-major N-wide forms of the arithmetic (mms_fastmath.cuh).
//
// Arithmetic is exactly that of the monolithic fast kernel (same device functions).
#pragma once

#include <cuda_pipeline.h>

#include "chain_kernel.cuh"

#if !MMS_STRICT

namespace mms {
namespace MMS_MODE {

// Hours per hand-off (divides 24) and unroll factors of the per-hand-off hour loops of the forcing / state
// warps.  Measured on BASELINE configs[1], ms per launch (K / forcing unroll / state unroll):
//   3/3/3 61.6   3/1/1 68.1   3/3/1 71.3   3/1/3 58.7   2/1/2 57.2   2/2/2 56.5   1/1/1 55.1
// -- the smaller the instruction footprint of the two loops (together they overflow the 32 KB L1.5
// instruction cache when unrolled x3), the better; handing over every hour with static ring offsets wins.
#ifndef MMS_WS_K
#define MMS_WS_K 1
#endif
constexpr int kWsK = MMS_WS_K;
#ifndef MMS_WS_UNROLL_F
#define MMS_WS_UNROLL_F 1
#endif
#ifndef MMS_WS_UNROLL_S
#define MMS_WS_UNROLL_S 1
#endif
constexpr int kWsUnrollF = MMS_WS_UNROLL_F, kWsUnrollS = MMS_WS_UNROLL_S;
constexpr int kWsFields = 6;         // evap, surv_ad, exp(-mdr/24), egg, link0, gain_m_h

// The hand-off: two slots and one named barrier per pair (bar.sync id, 64) at every hour.  ncu (profiles/r01_*,
// per-role stall samples in profiles/r02_pair_kernel_binder.md) shows the state warps waiting 11 % of their time at
// that barrier and the forcing warps 3 %, so round 2 tried to decouple the roles: MMS_WS_RING = 4 replaces the
// barrier by a ring of four slots with a full / empty pair of shared-memory mbarriers per slot (every lane arrives),
// so that the forcing warp can run up to four hours ahead.  MEASURED (BASELINE configs[1], one B200): 54.4 ms against
// 52.8 ms for the barrier form -- the try_wait / arrive pairs and the slot arithmetic cost more issue slots and
// latency on the two in-order instruction streams than the lock-step loses.  Kept behind the switch (parity-tested),
// off by default.
#ifndef MMS_WS_RING
#define MMS_WS_RING 0
#endif
constexpr int kWsRing = MMS_WS_RING;             // slots (a power of two), or 0
// MMS_WS_PIPE: the state warp of a pair runs the water step of hour h next to the population update of hour h - 1
// (state_pass_pipe below); it then reads two consecutive slots while the forcing warp fills a third: four slots
#ifndef MMS_WS_PIPE
#define MMS_WS_PIPE 1
#endif
#ifndef MMS_WS_PIPE_MAX_NS
#define MMS_WS_PIPE_MAX_NS 2       // 4: the fan-out form's state warps (four settings each) as well -- measured
                                   // slower (9472 cells x 64 settings: 102.6 against 90.0 ms: eight chains at 96 registers)
#endif
#if MMS_WS_RING > 0 && MMS_WS_PIPE
#error "the mbarrier ring experiment (MMS_WS_RING) needs -DMMS_WS_PIPE=0: the pipelined state pass uses the barrier ring"
#endif
constexpr int kWsBarSlots = MMS_WS_PIPE ? 4 : 2;
constexpr int kWsSlots = kWsRing > 0 ? kWsRing : kWsBarSlots;
static_assert(kWsRing == 0 || (kWsRing & (kWsRing - 1)) == 0, "the ring needs a power-of-two number of slots");
static_assert(kWsRing == 0 || MMS_WS_K == 1, "the mbarrier ring hands over one hour at a time");
// MMS_WS_QUAD_ASYNC: the forcing warp requests the ds(T, RH) quad of hour h + 1 -- an L2 access, 20-25 % of its stall
// samples at the first use (profiles/r02_ncu_chain_ws.txt) -- while it works on hour h, with cp.async into a
// double-buffered shared-memory record: an asynchronous copy holds no registers (carrying the requested quad in
// registers spilled: MMS_WS_DSPREFETCH above).  The microclimate of hour h + 1 that locates the quad is parked in the
// ring slot of hour h + 1, which nobody reads yet (the state warp is at slots h - 1 and h - 2 at most).
// MEASURED (BASELINE configs[1], one B200): 54.6 ms against 49.6 ms without -- the forcing warp's hour grows from 261 to
// 331 instructions (parking and re-reading six values, the copy's address, commit / wait, one more uniform branch),
// which costs more than the stall it removes.  Parity-tested, off by default.
#ifndef MMS_WS_QUAD_ASYNC
#define MMS_WS_QUAD_ASYNC 0
#endif
constexpr int kWsQuadDoubles = MMS_WS_QUAD_ASYNC ? 2 * 4 * 32 : 0;               // [buffer][half][lane][2]
constexpr int kWsRingDoubles = kWsSlots * kWsK * kWsFields * 32 + 2 * kWsSlots + kWsQuadDoubles;   // per pair: slots, full[] / empty[] mbarriers, quad buffers
static_assert(!MMS_WS_QUAD_ASYNC || (kWsRing == 0 && kWsBarSlots == 4 && MMS_WS_K == 1),
              "the asynchronous quad parks the next hour's microclimate in the fourth ring slot");

// The ds(T, RH) quad table (2 MB) lives in L2: its look-up is the one long-latency load of the forcing warp's hour
// (19 % of the forcing warps' stall samples sit on the first use of the quad).  Round 2 tried to request the quad of
// hour h + 1 at the top of hour h.  MEASURED (same workload): carrying the requested quad in registers 63.4 ms
// (the forcing role is at its 80-register cap -- 22 warps x 2560 registers -- and the carried values spill inside the
// hour loop); a bare prefetch.global.L1 of the record, nothing carried, 57.9 ms (the ~35 instructions that find the
// next hour's record cost more than the latency they hide: both instruction streams are in-order and bound by
// FP64 dependent-issue latency, so every added instruction lengthens the hour).  Off by default.
// (1 = carry the requested quad in registers; 2 = only a prefetch.global.L1 of its two sectors, nothing carried.)
#ifndef MMS_WS_DSPREFETCH
#define MMS_WS_DSPREFETCH 0
#endif
#if MMS_WS_QUAD_ASYNC && MMS_WS_DSPREFETCH
#error "MMS_WS_DSPREFETCH and MMS_WS_QUAD_ASYNC are alternatives"
#endif

// Which role evaluates the last, transcendental step of the three per-cell-hour products that do not depend on the
// setting: bit 0 the density-free link of das0 (division + log), bit 1 exp(-mdr / 24), bit 2 ds^(1/24) (log + exp).
// A set bit moves that step from the forcing warp (which then hands over the raw rate) to the state warp.  The
// forcing warp is the critical role of the pair: ncu (profiles/r02_ncu_chain_ws.txt) has the state warps waiting
// at the hand-off barrier for a quarter of their time.
// MEASURED (BASELINE configs[1], ms per launch; 52.8 with everything in the forcing warp): bit 0 alone 51.4, bit 1 alone
// 52.1, bits 0+1 55.0, bit 2 alone 53.0, bits 0+2 56.9, all three 61.1 -- the two roles are within a few per cent of
// each other, whatever the barrier-stall samples suggest.  Default: the link moves (the pair kernel only; in the
// fan-out kernel sixteen state warps would each repeat it).
#ifndef MMS_WS_SPLIT
#define MMS_WS_SPLIT 1
#endif
constexpr int kWsSplit = MMS_WS_SPLIT;

// hour_forcing() (chain_kernel.cuh) without the steps that MMS_WS_SPLIT leaves to the state warp
template <bool kFull, bool kWaterOnly, int kSplit>
__device__ __forceinline__ HourF hour_forcing_ws(const ChainParams &p, const SmemTables &tb, const DayCtx &dc,
                                                 double p_a, int h, int dens_type)
{
    if (kSplit == 0) return hour_forcing<kFull, kWaterOnly>(p, tb, dc, p_a, h, dens_type);
    HourF f;
    hour_microclimate<kFull>(p.hr, dc, h, f);
    f.mdr = f.efd = f.das0 = f.ds = 0.0;
    double dx = 0.0;
    double4 c0 = make_double4(0, 0, 0, 0), c1 = c0, c2 = c0;
    DsFetch dsf;
    if (!kWaterOnly) {
        dsf = ds_fetch(p.ds, f.air_t, f.rh);
        const int i = spline_interval(p.spline, tb.sx, f.air_t);
        dx = f.air_t - tb.sx[i];
        const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
        c0 = cf[0]; c1 = cf[1]; c2 = cf[2];
    }
    f.evap = evaporation_rate(f.air_t, f.rh, p_a, f.wind);
    if (!kWaterOnly) {
        // the ds quad (an L2 access) is used as late as the hour allows
        f.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
        f.efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
        f.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
        f.ph.egg = f.efd * kLitC[3];
        f.ph.one_m_emf = (kSplit & 2) ? f.mdr : pop_emergence_complement(f.mdr);
        f.ph.link0 = (kSplit & 1) ? f.das0 : pop_link0(dens_type, f.das0);
        f.ds = ds_blend(dsf);
        f.ph.surv_ad = (kSplit & 4) ? f.ds : r_pow_pos(f.ds, kLitC[3]);
    }
    return f;
}

// prefetch.global.L1 of the quad record the ds look-up of (temperature, humidity) will read
__device__ __forceinline__ void ds_prefetch_l1(const SurfaceDev &g, double temperature, double humidity)
{
    const double fx = (temperature - g.x0) * g.inv_dx, fy = (humidity - g.y0) * g.inv_dy;
    int i = __double2int_rd(fx), j = __double2int_rd(fy);
    i = min(max(i, 0), g.nx - 2);
    j = min(max(j, 0), g.ny - 2);
    const double *q = g.quad + (size_t)(i + (g.nx - 1) * j) * 4;       // one 32-byte record = one sector
    asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
}

__device__ __forceinline__ void pair_barrier(int pair, int threads = 64)
{
    asm volatile("bar.sync %0, %1;" ::"r"(pair + 1), "r"(threads) : "memory");   // barrier 0 is __syncthreads()
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}

// wait until the phase with this parity of the barrier has completed (try_wait suspends the warp for a
// hardware-defined time before it reports failure: no busy spin on the issue slots of the sub-partition)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MMS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MMS_DONE_%=;\n"
        "bra MMS_WAIT_%=;\n"
        "MMS_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// one end of a pair's ring: where the slots and the barriers are, and how many hand-offs this warp has made
// (both roles make the same hand-offs in the same order, over all passes and all units of the kernel)
struct RingEnd {
    double *slots;          // [kWsSlots][kWsFields][32], this lane's column
    uint32_t full, empty;   // shared-memory addresses of full[0], empty[0]
    uint32_t it;            // hand-offs so far
    double *quad;           // [2 buffers][2 halves][32][2], this lane's first half
    int pair;               // named barrier of the team (pair + 1)
    int bar_threads;        // threads that meet at it: 64 for a pair, 32 (1 + state warps) for a fan-out team
};

__device__ __forceinline__ RingEnd ring_end(double *ring, int pair, int lane)
{
    RingEnd r;
    r.slots = ring + lane;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + kWsSlots * kWsK * kWsFields * 32);
    r.quad = ring + kWsSlots * kWsK * kWsFields * 32 + 2 * kWsSlots + 2 * lane;
    r.full = smem_u32(bars);
    r.empty = smem_u32(bars + kWsSlots);
    r.it = 0;
    r.pair = pair;
    r.bar_threads = 64;
    return r;
}

// forcing warp: the slot to fill next (waits until the state warp has drained it)
__device__ __forceinline__ double *ring_acquire_empty(RingEnd &r)
{
    if (kWsRing > 0) {
        const uint32_t slot = r.it & (kWsSlots - 1);
        mbar_wait(r.empty + 8 * slot, ((r.it / kWsSlots) & 1) ^ 1);
        return r.slots + (size_t)slot * (kWsFields * 32);
    }
    return r.slots + (size_t)(r.it & (kWsBarSlots - 1)) * (kWsK * kWsFields * 32);
}
__device__ __forceinline__ void ring_publish(RingEnd &r)
{
    if (kWsRing > 0) mbar_arrive(r.full + 8 * (r.it & (kWsSlots - 1)));
    else pair_barrier(r.pair, r.bar_threads);        // the state warp(s) may now read this buffer
    r.it += 1;
}
// state warp: the slot to read next (waits until the forcing warp has filled it)
__device__ __forceinline__ const double *ring_acquire_full(RingEnd &r)
{
    if (kWsRing > 0) {
        const uint32_t slot = r.it & (kWsSlots - 1);
        mbar_wait(r.full + 8 * slot, (r.it / kWsSlots) & 1);
        return r.slots + (size_t)slot * (kWsFields * 32);
    }
    pair_barrier(r.pair, r.bar_threads);             // the forcing warp has filled this buffer
    return r.slots + (size_t)(r.it & (kWsBarSlots - 1)) * (kWsK * kWsFields * 32);
}
__device__ __forceinline__ void ring_release(RingEnd &r)
{
    if (kWsRing > 0) mbar_arrive(r.empty + 8 * (r.it & (kWsSlots - 1)));
    r.it += 1;
}

#if MMS_WS_DSPREFETCH == 1
// the microclimate of an hour and (requested, not yet used) its ds quad
struct HourAhead {
    double air_t, rh, wind, cloud;
    DsFetch dsf;
};

template <bool kFull, bool kWaterOnly>
__device__ __forceinline__ HourAhead hour_ahead(const ChainParams &p, const DayCtx &dc, int h)
{
    HourF f;
    hour_microclimate<kFull>(p.hr, dc, h, f);
    HourAhead a;
    a.air_t = f.air_t; a.rh = f.rh; a.wind = f.wind; a.cloud = f.cloud;
    if (!kWaterOnly) a.dsf = ds_fetch(p.ds, f.air_t, f.rh);
    return a;
}

// hour_forcing() (chain_kernel.cuh) from a microclimate evaluated, and a quad requested, an hour earlier
template <bool kFull, bool kWaterOnly>
__device__ __forceinline__ HourF hour_forcing_from(const ChainParams &p, const SmemTables &tb, const HourAhead &a,
                                                   double p_a, int dens_type)
{
    HourF f;
    f.air_t = a.air_t; f.rh = a.rh; f.wind = a.wind; f.cloud = a.cloud;
    f.mdr = f.efd = f.das0 = f.ds = 0.0;
    double dx = 0.0;
    double4 c0 = make_double4(0, 0, 0, 0), c1 = c0, c2 = c0;
    if (!kWaterOnly) {
        const int i = spline_interval(p.spline, tb.sx, f.air_t);
        dx = f.air_t - tb.sx[i];
        const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
        c0 = cf[0]; c1 = cf[1]; c2 = cf[2];
    }
    f.evap = evaporation_rate(f.air_t, f.rh, p_a, f.wind);      // R/sim_water.R:264-270 (air temperature)
    if (!kWaterOnly) {
        f.ds = ds_blend(a.dsf);
        f.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
        f.efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
        f.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
        f.ph = pop_hour(dens_type, f.mdr, f.efd, f.das0, f.ds);
    }
    return f;
}
#endif

// period bookkeeping shared by both roles (aggregate_vectors(), R/pipeline_functions.R:1050-1088)
struct PeriodCursor {
    int cur_period, acc_n, acc_inc;
    bool acc_on;
    __device__ __forceinline__ void init() { cur_period = -1; acc_n = 0; acc_inc = 0; acc_on = false; }
    // returns the period to flush (or -1) and whether a new period starts on this day
    __device__ __forceinline__ int begin_day(const ChainParams &p, int day, bool &fresh)
    {
        fresh = false;
        int to_flush = -1;
        const int period = p.period_of_day ? p.period_of_day[day] : -1;
        if (period >= 0 && period != cur_period) {
            to_flush = cur_period;
            fresh = true;
            cur_period = period;
        }
        acc_on = period >= 0;
        acc_inc = acc_on ? 1 : 0;
        return to_flush;
    }
};

#if MMS_WS_QUAD_ASYNC
// ds_fetch() (chain_kernel.cuh) in two steps: where the quad is, and -- an hour later -- the look-up from the copy
__device__ __forceinline__ void ds_request(const SurfaceDev &g, double temperature, double humidity, double *buf,
                                           double &ex, double &ey)
{
    const double fx = (temperature - g.x0) * g.inv_dx, fy = (humidity - g.y0) * g.inv_dy;
    int i = __double2int_rd(fx), j = __double2int_rd(fy);     // NaN -> 0
    i = min(max(i, 0), g.nx - 2);
    j = min(max(j, 0), g.ny - 2);
    ex = fx - (double)i;
    ey = fy - (double)j;                                      // == 1 on the last grid line
    const double2 *q = reinterpret_cast<const double2 *>(g.quad) + (i + (g.nx - 1) * j) * 2;
    __pipeline_memcpy_async(buf, q, 16);                      // (z00, z10)
    __pipeline_memcpy_async(buf + 2 * 32, q + 1, 16);         // (z01, z11)
    __pipeline_commit();
}
__device__ __forceinline__ DsFetch ds_arrived(const SurfaceDev &g, double temperature, double humidity, const double *buf,
                                              double ex, double ey)
{
    DsFetch d;
    d.inside = temperature >= g.x0 && temperature <= g.x1 && humidity >= g.y0 && humidity <= g.y1;
    d.ex = ex;
    d.ey = ey;
    d.a = *reinterpret_cast<const double2 *>(buf);
    d.b = *reinterpret_cast<const double2 *>(buf + 2 * 32);
    return d;
}
#endif

// ---- forcing warp ---------------------------------------------------------------------------
template <bool kFull, bool kHourly, bool kWaterOnly, bool kRecord, int kDens, int kSplit>
__device__ __forceinline__ void forcing_pass(const ChainParams &p, const SmemTables &tb, const UnitCtx &u,
                                             RingEnd &ring, int n_days, int &bad)
{
    constexpr int kNF = kFull ? 9 : 4;
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    const int lane = threadIdx.x & 31;
    double acc[kNF];
#pragma unroll
    for (int k = 0; k < kNF; ++k) acc[k] = 0.0;
    PeriodCursor pc;
    pc.init();

    auto flush = [&](int period) {
        if (!u.live || !u.write_forcing) return;
        const int64_t at = (int64_t)period * u.os + u.ocell;
        const double n = (double)pc.acc_n;
        const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                            MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
        for (int k = 0; k < kNF; ++k) {
            if (p.mean_mask & (1u << vid[k])) {
                const int64_t row = out_slot(p.mean_mask, ~kState, vid[k]);
                p.mean_forcing[row * p.P * u.os + at] = acc[k] / n;
            }
        }
    };

    DayCtx dc;
    dc.cur = load_daily<kFull>(p, u.cell, 0, bad);
    dc.nxt = load_daily<kFull>(p, u.cell, (p.D > 1) ? 1 : 0, bad);
    dc.A = (dc.cur.tmax - dc.cur.tmin) * 0.5;                              // R/sim_microclimate.R:217
    dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;                // :220
    dc.t_midnight_prev = (dc.t_sunset - dc.nxt.tmin) / p.hr.exp_e[23] + dc.nxt.tmin;   // :271-272
    dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;

    (void)lane;
    for (int day = 0; day < n_days; ++day) {
        if (kRecord) {
            bool fresh;
            const int to_flush = pc.begin_day(p, day, fresh);
            if (fresh) {
                if (to_flush >= 0) flush(to_flush);
#pragma unroll
                for (int k = 0; k < kNF; ++k) acc[k] = 0.0;
                pc.acc_n = 0;
            }
        }
        const double rain_h = dc.cur.rain * (1.0 / 24.0);                 // R/sim_microclimate.R:132
        const double gain_m_h = rain_h * (1.0 / 1000.0);                  // R/sim_water.R:122
        double air_t_last = 0.0;
#if MMS_WS_DSPREFETCH == 1
        HourAhead ahead = hour_ahead<kFull, kWaterOnly>(p, dc, 0);
#endif
#if MMS_WS_QUAD_ASYNC
        constexpr bool kAsync = !kWaterOnly;
        // the microclimate of hour hn into the ring slot `park` (its own hour's slot, not yet in use), its quad requested
        auto look_ahead = [&](int hn, double *park) {
            HourF fa;
            hour_microclimate<kFull>(p.hr, dc, hn, fa);
            double ex, ey;
            ds_request(p.ds, fa.air_t, fa.rh, ring.quad + (hn & 1) * (4 * 32), ex, ey);
            park[0 * 32] = fa.air_t; park[1 * 32] = fa.rh; park[2 * 32] = fa.wind;
            if (kFull) park[3 * 32] = fa.cloud;
            park[4 * 32] = ex; park[5 * 32] = ey;
        };
        if (kAsync) look_ahead(0, ring.slots + (size_t)(ring.it & 3) * (kWsFields * 32));
#else
        constexpr bool kAsync = false;
#endif
#pragma unroll 1
        for (int h0 = 0; h0 < 24; h0 += kWsK) {
#if MMS_WS_DSPREFETCH == 1
            // the quad of the hour to come is requested before this hour's arithmetic starts
            const HourAhead now = ahead;
            if (h0 + kWsK < 24) ahead = hour_ahead<kFull, kWaterOnly>(p, dc, h0 + kWsK);
#elif MMS_WS_DSPREFETCH == 2
            if (!kWaterOnly && h0 + kWsK < 24) {
                HourF fa;
                hour_microclimate<false>(p.hr, dc, h0 + kWsK, fa);
                ds_prefetch_l1(p.ds, fa.air_t, fa.rh);
            }
#endif
            double *slot = ring_acquire_empty(ring);
#pragma unroll kWsUnrollF
            for (int k = 0; k < kWsK; ++k) {
                const int h = h0 + k;
#if MMS_WS_DSPREFETCH == 1
                const HourF f = hour_forcing_from<kFull, kWaterOnly>(p, tb, now, u.p_a, kDens >= 0 ? (kDens & 1) : p.densmod_type);
#elif MMS_WS_QUAD_ASYNC
                HourF f;
                if (kAsync) {
                    if (h < 23) look_ahead(h + 1, ring.slots + (size_t)((ring.it + 1) & 3) * (kWsFields * 32));
                    const int dens_type = kDens >= 0 ? (kDens & 1) : p.densmod_type;
                    f.air_t = slot[0 * 32]; f.rh = slot[1 * 32]; f.wind = slot[2 * 32];
                    f.cloud = kFull ? slot[3 * 32] : 0.0;
                    const double ex = slot[4 * 32], ey = slot[5 * 32];
                    const int i = spline_interval(p.spline, tb.sx, f.air_t);
                    const double dx = f.air_t - tb.sx[i];
                    const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
                    const double4 c0 = cf[0], c1 = cf[1], c2 = cf[2];
                    f.evap = evaporation_rate(f.air_t, f.rh, u.p_a, f.wind);
                    f.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
                    f.efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
                    f.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
                    f.ph.egg = f.efd * kLitC[3];
                    f.ph.one_m_emf = (kSplit & 2) ? f.mdr : pop_emergence_complement(f.mdr);
                    f.ph.link0 = (kSplit & 1) ? f.das0 : pop_link0(dens_type, f.das0);
                    if (h < 23) __pipeline_wait_prior(1);     // this hour's quad has landed (the next one's may be in flight)
                    else __pipeline_wait_prior(0);
                    f.ds = ds_blend(ds_arrived(p.ds, f.air_t, f.rh, ring.quad + (h & 1) * (4 * 32), ex, ey));
                    f.ph.surv_ad = (kSplit & 4) ? f.ds : r_pow_pos(f.ds, kLitC[3]);
                } else {
                    f = hour_forcing_ws<kFull, kWaterOnly, kSplit>(p, tb, dc, u.p_a, h, kDens >= 0 ? (kDens & 1) : p.densmod_type);
                }
#else
                const HourF f = hour_forcing_ws<kFull, kWaterOnly, kSplit>(p, tb, dc, u.p_a, h, kDens >= 0 ? (kDens & 1) : p.densmod_type);
#endif
                double *rec = slot + (size_t)k * (kWsFields * 32);
                rec[0 * 32] = f.evap;
                rec[1 * 32] = f.ph.surv_ad;
                rec[2 * 32] = f.ph.one_m_emf;          // exp(-timestep * mdr); emf = 1 - it
                rec[3 * 32] = f.ph.egg;
                rec[4 * 32] = f.ph.link0;
                rec[5 * 32] = gain_m_h;
                air_t_last = f.air_t;
                if (kRecord) {
                    acc[0] += (pc.acc_on ? f.mdr : 0.0); acc[1] += (pc.acc_on ? f.efd : 0.0);
                    acc[2] += (pc.acc_on ? f.das0 : 0.0); acc[3] += (pc.acc_on ? f.ds : 0.0);
                    if (kFull) {
                        acc[4] += (pc.acc_on ? f.air_t : 0.0); acc[5] += (pc.acc_on ? f.rh : 0.0);
                        acc[6] += (pc.acc_on ? f.wind : 0.0); acc[7] += (pc.acc_on ? rain_h : 0.0);
                        acc[8] += (pc.acc_on ? f.cloud : 0.0);
                    }
                    pc.acc_n += pc.acc_inc;
                    if (kHourly && u.live && u.write_forcing) {
                        const int64_t tc = (int64_t)(day * 24 + h) * u.os + u.ocell;
                        const double val[9] = {f.mdr, f.efd, f.das0, f.ds, f.air_t, f.rh, f.wind, rain_h, f.cloud};
                        const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                                            MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
                        for (int q = 0; q < kNF; ++q) {
                            if (p.hourly_mask & (1u << vid[q])) {
                                const int64_t row = out_slot(p.hourly_mask, ~kState, vid[q]);
                                p.hourly_forcing[row * u.H * u.os + tc] = val[q];
                            }
                        }
                    }
                }
            }
            ring_publish(ring);
        }
        // ---- roll the day window
        dc.t_midnight_prev = air_t_last;             // hour-24 value, R/sim_microclimate.R:271
        dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;
        dc.cur = dc.nxt;
        const int d2 = day + 2;
        dc.nxt = load_daily<kFull>(p, u.cell, d2 < p.D ? d2 : p.D - 1, bad);
        dc.A = (dc.cur.tmax - dc.cur.tmin) * 0.5;
        dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;
    }
    if (kRecord && pc.cur_period >= 0) flush(pc.cur_period);
}

// ---- state warp ------------------------------------------------------------------------------
// How a state warp receives an hour's state-independent values: from the forcing warp of its pair / fan-out team
// through the ring (PairLink), or from the two forcing roles of a team (TeamLink, chain_team.cuh)
template <int kNS, bool kFull, bool kHourly, bool kRecord, int kDens, int kSplit>
__device__ __forceinline__ void state_pass_pipe(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st,
                                                RingEnd &ring, int n_days);
template <int kNS, bool kFull, bool kHourly, bool kWaterOnly, bool kRecord, int kDens, int kSplit, class Link>
__device__ __forceinline__ void state_pass(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st,
                                           Link &link, int n_days);

struct PairLink {
    RingEnd &ring;
    const double *slot;
    bool pipelined;         // a pair (two settings per state warp at most): the software-pipelined hour loop
    // a pass with the population (burn-in year or the recorded series)
    template <int kNS, bool kFull, bool kHourly, bool kRecord, int kDens, int kSplit>
    __device__ __forceinline__ void pop_pass(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st, int n_days)
    {
        if (MMS_WS_PIPE && kNS <= MMS_WS_PIPE_MAX_NS && pipelined) state_pass_pipe<kNS, kFull, kHourly, kRecord, kDens, kSplit>(p, u, st, ring, n_days);
        else state_pass<kNS, kFull, kHourly, false, kRecord, kDens, kSplit>(p, u, st, *this, n_days);
    }
    __device__ __forceinline__ void begin_pass() {}
    __device__ __forceinline__ void end_pass() {}
    // hour h0 + k of a hand-off of kWsK hours
    __device__ __forceinline__ void fetch(int h, int k, double &evap, PopHour &ph, double &gain_m_h)
    {
        (void)h;
        if (k == 0) slot = ring_acquire_full(ring);
        const double *rec = slot + (size_t)k * (kWsFields * 32);
        evap = rec[0 * 32];
        ph.surv_ad = rec[1 * 32];
        ph.one_m_emf = rec[2 * 32];
        ph.egg = rec[3 * 32];
        ph.link0 = rec[4 * 32];
        gain_m_h = rec[5 * 32];
        if (k == kWsK - 1) ring_release(ring);   // the values are in registers: the forcing warp may refill the slot
    }
};

template <int kNS, bool kFull, bool kHourly, bool kWaterOnly, bool kRecord, int kDens, int kSplit, class Link>
__device__ __forceinline__ void state_pass(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st,
                                           Link &link, int n_days)
{
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    const int lane = threadIdx.x & 31;
    double acc[kNS][4];
    int acc_wet[kNS];
    auto reset = [&]() {
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            acc_wet[s] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[s][k] = 0.0;
        }
    };
    reset();
    PeriodCursor pc;
    pc.init();

    auto flush = [&](int period) {
        if (!u.live) return;
        const int64_t at = (int64_t)period * u.os + u.ocell;
        const double n = (double)pc.acc_n;
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!kFull && k >= 2) continue;
                if (p.mean_mask & (1u << vid[k])) {
                    const int64_t row = (int64_t)(u.set0 + s) * u.n_mean_state + out_slot(p.mean_mask, kState, vid[k]);
                    p.mean_state[row * p.P * u.os + at] = acc[s][k] / n;
                }
            }
            if (p.wet_hours) p.wet_hours[(int64_t)(u.set0 + s) * p.P * u.os + at] = acc_wet[s];
        }
    };

    (void)lane;
    link.begin_pass();
    for (int day = 0; day < n_days; ++day) {
        if (kRecord) {
            bool fresh;
            const int to_flush = pc.begin_day(p, day, fresh);
            if (fresh) {
                if (to_flush >= 0) flush(to_flush);
                reset();
                pc.acc_n = 0;
            }
        }
#pragma unroll 1
        for (int h0 = 0; h0 < 24; h0 += kWsK) {
#pragma unroll kWsUnrollS
            for (int k = 0; k < kWsK; ++k) {
                double evap, gain_m_h;
                PopHour ph;
                link.fetch(h0 + k, k, evap, ph, gain_m_h);
                if (!kWaterOnly) {
                    // the steps MMS_WS_SPLIT moved here: same device functions on the same values, same results
                    if (kSplit & 4) ph.surv_ad = r_pow_pos(ph.surv_ad, kLitC[3]);
                    if (kSplit & 2) ph.one_m_emf = pop_emergence_complement(ph.one_m_emf);
                    if (kSplit & 1) ph.link0 = pop_link0(kDens >= 0 ? (kDens & 1) : p.densmod_type, ph.link0);
                }
                ph.emf = 1.0 - ph.one_m_emf;
                {
                    // all settings of the thread stage by stage (mms_fastmath.cuh: why); the shape of setting s is
                    // compiled in when the launcher knew it (two bits per setting of kDens, 0 = run time)
                    double gain[kNS];
                    int shape[kNS];
                    MMS_EACH(s, kNS) {
                        const HabitatDev &hb = p.hab[u.set0 + s];
                        gain[s] = hb.catchment_area * gain_m_h;                                    // R/sim_water.R:122-123
                        const int shape_bits = kDens >= 0 ? (kDens >> (1 + 2 * s)) & 3 : 0;
                        shape[s] = shape_bits ? shape_bits - 1 : hb.shape;
                    }
                    water_step_n<kNS>(p.hab + u.set0, st.volume, st.surface, evap, gain, shape);   // :279-294
                    if (!kWaterOnly)
                        iterate_state_n<kNS>(st.aquatic, st.adult, ph, kDens >= 0 ? (kDens & 1) : p.densmod_type, p.dd_effect,
                                             p.surface_area_cm2, st.surface);
                }
                if (kRecord) {
#pragma unroll
                    for (int s = 0; s < kNS; ++s) {
                        acc[s][0] += (pc.acc_on ? st.adult[s] : 0.0);
                        acc[s][1] += (pc.acc_on ? st.aquatic[s] : 0.0);
                        if (kFull) {
                            acc[s][2] += (pc.acc_on ? st.surface[s] : 0.0);
                            acc[s][3] += (pc.acc_on ? st.volume[s] : 0.0);
                        }
                        acc_wet[s] += (st.volume[s] > 0.0) ? pc.acc_inc : 0;
                    }
                    pc.acc_n += pc.acc_inc;
                    if (kHourly && u.live) {
                        const int64_t tc = (int64_t)(day * 24 + h0 + k) * u.os + u.ocell;
#pragma unroll
                        for (int s = 0; s < kNS; ++s) {
                            const double val[4] = {st.adult[s], st.aquatic[s], st.surface[s], st.volume[s]};
                            const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (!kFull && q >= 2) continue;
                                if (p.hourly_mask & (1u << vid[q])) {
                                    const int64_t row = (int64_t)(u.set0 + s) * u.n_hourly_state + out_slot(p.hourly_mask, kState, vid[q]);
                                    p.hourly_state[row * u.H * u.os + tc] = val[q];
                                }
                            }
                            if (p.wet_flags) p.wet_flags[(int64_t)(u.set0 + s) * u.H * u.os + tc] = (st.volume[s] > 0.0) ? 1 : 0;
                        }
                    }
                }
            }
        }
    }
    link.end_pass();
    if (kRecord && pc.cur_period >= 0) flush(pc.cur_period);
}

// The state pass of a PAIR with the hour loop software-pipelined: in one loop body the water step of hour h (it needs
// that hour's evaporation rate and rain only) and the population update of hour h - 1 (it needs the surface of hour
// h - 1, kept) -- the population does not feed back into the water, so a thread with kNS settings has 2 kNS independent
// chains in the body instead of kNS.  The pending hour's rates are re-read from the previous ring slot (four slots:
// the forcing warp is filling a third).  Day and period bookkeeping of a day runs at the END of the iteration that
// recorded the last hour of the day before.  Same operations on the same values as state_pass().
template <int kNS, bool kFull, bool kHourly, bool kRecord, int kDens, int kSplit>
__device__ __forceinline__ void state_pass_pipe(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st,
                                                RingEnd &ring, int n_days)
{
    static_assert(kWsK == 1 && kWsRing == 0 && kWsBarSlots == 4, "one hour per hand-off through the four-slot barrier ring");
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    const int dens_type = kDens >= 0 ? (kDens & 1) : p.densmod_type;
    double acc[kNS][4];
    int acc_wet[kNS];
    auto reset = [&]() {
        MMS_EACH(s, kNS) {
            acc_wet[s] = 0;
            MMS_EACH(k, 4) acc[s][k] = 0.0;
        }
    };
    reset();
    PeriodCursor pc;
    pc.init();

    auto flush = [&](int period) {
        if (!u.live) return;
        const int64_t at = (int64_t)period * u.os + u.ocell;
        const double n = (double)pc.acc_n;
        MMS_EACH(s, kNS) {
            const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
            MMS_EACH(k, 4) {
                if (!kFull && k >= 2) continue;
                if (p.mean_mask & (1u << vid[k])) {
                    const int64_t row = (int64_t)(u.set0 + s) * u.n_mean_state + out_slot(p.mean_mask, kState, vid[k]);
                    p.mean_state[row * p.P * u.os + at] = acc[s][k] / n;
                }
            }
            if (p.wet_hours) p.wet_hours[(int64_t)(u.set0 + s) * p.P * u.os + at] = acc_wet[s];
        }
    };
    auto begin_day = [&](int day) {
        if (!kRecord) return;
        bool fresh;
        const int to_flush = pc.begin_day(p, day, fresh);
        if (fresh) {
            if (to_flush >= 0) flush(to_flush);
            reset();
            pc.acc_n = 0;
        }
    };
    // the next slot the forcing warp has filled, and the one before it
    auto next_slot = [&](const double *&cur, const double *&prev) {
        pair_barrier(ring.pair, ring.bar_threads);
        cur = ring.slots + (size_t)(ring.it & 3) * (kWsFields * 32);
        prev = ring.slots + (size_t)((ring.it - 1) & 3) * (kWsFields * 32);
        ring.it += 1;
    };
    auto water = [&](const double *rec) {
        const double evap = rec[0 * 32], gain_m_h = rec[5 * 32];
        double gain[kNS];
        int shape[kNS];
        MMS_EACH(s, kNS) {
            const HabitatDev &hb = p.hab[u.set0 + s];
            gain[s] = hb.catchment_area * gain_m_h;                                    // R/sim_water.R:122-123
            const int shape_bits = kDens >= 0 ? (kDens >> (1 + 2 * s)) & 3 : 0;
            shape[s] = shape_bits ? shape_bits - 1 : hb.shape;
        }
        water_step_n<kNS>(p.hab + u.set0, st.volume, st.surface, evap, gain, shape);   // :279-294
    };
    auto rates = [&](const double *rec) {
        PopHour ph;
        ph.surv_ad = rec[1 * 32];
        ph.one_m_emf = rec[2 * 32];
        ph.egg = rec[3 * 32];
        ph.link0 = rec[4 * 32];
        if (kSplit & 4) ph.surv_ad = r_pow_pos(ph.surv_ad, kLitC[3]);
        if (kSplit & 2) ph.one_m_emf = pop_emergence_complement(ph.one_m_emf);
        if (kSplit & 1) ph.link0 = pop_link0(dens_type, ph.link0);
        ph.emf = 1.0 - ph.one_m_emf;
        return ph;
    };
    // the outputs of hour (day, h): population now, water body as it was at that hour
    auto record = [&](int day, int h, const double (&surface)[kNS], const double (&volume)[kNS]) {
        if (!kRecord) return;
        MMS_EACH(s, kNS) {
            acc[s][0] += (pc.acc_on ? st.adult[s] : 0.0);
            acc[s][1] += (pc.acc_on ? st.aquatic[s] : 0.0);
            if (kFull) {
                acc[s][2] += (pc.acc_on ? surface[s] : 0.0);
                acc[s][3] += (pc.acc_on ? volume[s] : 0.0);
            }
            acc_wet[s] += (volume[s] > 0.0) ? pc.acc_inc : 0;
        }
        pc.acc_n += pc.acc_inc;
        if (kHourly && u.live) {
            const int64_t tc = (int64_t)(day * 24 + h) * u.os + u.ocell;
            MMS_EACH(s, kNS) {
                const double val[4] = {st.adult[s], st.aquatic[s], surface[s], volume[s]};
                const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
                MMS_EACH(q, 4) {
                    if (!kFull && q >= 2) continue;
                    if (p.hourly_mask & (1u << vid[q])) {
                        const int64_t row = (int64_t)(u.set0 + s) * u.n_hourly_state + out_slot(p.hourly_mask, kState, vid[q]);
                        p.hourly_state[row * u.H * u.os + tc] = val[q];
                    }
                }
                if (p.wet_flags) p.wet_flags[(int64_t)(u.set0 + s) * u.H * u.os + tc] = (volume[s] > 0.0) ? 1 : 0;
            }
        }
    };

    const double *cur, *prev;
    begin_day(0);
    next_slot(cur, prev);
    water(cur);                                      // hour 0 of day 0; its population update is pending
    int pday = 0, phour = 0;
    for (int day = 0; day < n_days; ++day) {
#pragma unroll 1
        for (int h = (day == 0) ? 1 : 0; h < 24; ++h) {
            next_slot(cur, prev);
            const PopHour ph = rates(prev);
            double surface_prev[kNS], volume_prev[kNS];
            MMS_EACH(s, kNS) { surface_prev[s] = st.surface[s]; volume_prev[s] = st.volume[s]; }
            water(cur);                                                                                   // hour (day, h)
            iterate_state_n<kNS>(st.aquatic, st.adult, ph, dens_type, p.dd_effect, p.surface_area_cm2, surface_prev);   // hour (pday, phour)
            record(pday, phour, surface_prev, volume_prev);
            if (h == 0) begin_day(day);              // the day before is complete: its period may end here
            pday = day;
            phour = h;
        }
    }
    {   // the last hour's population update
        const double *last = ring.slots + (size_t)((ring.it - 1) & 3) * (kWsFields * 32);
        const PopHour ph = rates(last);
        iterate_state_n<kNS>(st.aquatic, st.adult, ph, dens_type, p.dd_effect, p.surface_area_cm2, st.surface);
        record(pday, phour, st.surface, st.volume);
    }
    if (kRecord && pc.cur_period >= 0) flush(pc.cur_period);
}

// one work unit in the state role: water burn-in years (R/sim_water.R:228-233), population burn-in years
// (R/sim_pop_dynamics.R:221-226), then the full series
template <int kNS, bool kFull, bool kHourly, int kDens, int kSplit, class Link>
__device__ __forceinline__ void state_unit(const ChainParams &p, const UnitCtx &u, Link &link)
{
    UnitState<kNS> st;
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
        st.aquatic[s] = p.initial_aquatic;               // R/sim_pop_dynamics.R:249-252
        st.adult[s] = p.initial_adult;
        st.volume[s] = p.hab[u.set0 + s].initial_volume; // R/sim_water.R:277
        st.surface[s] = habitat_surface(p.hab[u.set0 + s], st.volume[s]);
    }
    for (int y = 0; y < p.burnin_years_water; ++y)
        state_pass<kNS, kFull, kHourly, true, false, kDens, kSplit>(p, u, st, link, kDaysPerYear);
    double vol0[kNS];
#pragma unroll
    for (int s = 0; s < kNS; ++s) vol0[s] = st.volume[s];
    for (int y = 0; y < p.burnin_years_pop; ++y) {
        link.template pop_pass<kNS, kFull, kHourly, false, kDens, kSplit>(p, u, st, kDaysPerYear);
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            st.volume[s] = vol0[s];
            st.surface[s] = habitat_surface(p.hab[u.set0 + s], vol0[s]);
        }
    }
    link.template pop_pass<kNS, kFull, kHourly, true, kDens, kSplit>(p, u, st, p.D);
}

// one work unit in one role -- identical loop structure in both roles (burn-in passes, then the full series), so
// the warps of a team meet at every barrier
template <int kNS, bool kFull, bool kHourly, int kDens, int kSplit>
__device__ __forceinline__ void role_unit(const ChainParams &p, const SmemTables &tb, UnitCtx &u, RingEnd &ring,
                                          bool forcing_role)
{
    if (forcing_role) {
        u.p_a = air_pressure_kpa(__ldg(p.altitude + u.cell));
        int bad = 0;
        for (int y = 0; y < p.burnin_years_water; ++y)
            forcing_pass<kFull, kHourly, true, false, kDens, kSplit>(p, tb, u, ring, kDaysPerYear, bad);
        for (int y = 0; y < p.burnin_years_pop; ++y)
            forcing_pass<kFull, kHourly, false, false, kDens, kSplit>(p, tb, u, ring, kDaysPerYear, bad);
        forcing_pass<kFull, kHourly, false, true, kDens, kSplit>(p, tb, u, ring, p.D, bad);
        if (bad && u.live) atomicOr(p.error_flag, 1);
    } else {
        u.p_a = 0.0;
        PairLink link{ring, nullptr, ring.bar_threads == 64 || MMS_WS_PIPE_MAX_NS >= 4};
        state_unit<kNS, kFull, kHourly, kDens, kSplit>(p, u, link);
    }
}

// ---- the kernel: warp 2i = forcing role, warp 2i + 1 = state role of pair i ---------------------
// kDens: what the launcher knows at compile time, or -1 (read the parameters): bit 0 the density-dependence form
// (0 logit, 1 cox_ph), then two bits per setting of a unit with its water-body shape + 1 (0 = read hab[].shape).
// The other form's / shape's code and the (uniform, but opaque to the compiler) branches around it leave both loops
template <int kNS, bool kFull, bool kHourly, int kMaxPairs, int kDens>
__global__ void __launch_bounds__(kMaxPairs * 64, 1)
chain_kernel_ws(const __grid_constant__ ChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_coef = reinterpret_cast<double *>(smem_raw);              // [n][3][4]
    double *s_x = s_coef + (size_t)p.spline.n * kSplineRow;                     // [n]
    double *s_gx = s_x + p.spline.n;                                    // [nx]
    double *s_gy = s_gx + p.ds.nx;                                      // [ny]
    double *s_ring = s_gy + p.ds.ny;
    s_ring += (reinterpret_cast<uintptr_t>(s_ring) >> 3) & 1;           // 16-byte aligned (cp.async records)                                    // [pairs][kWsRingDoubles]
    const int n_pairs = blockDim.x >> 6;
    int *s_unit = reinterpret_cast<int *>(s_ring + (size_t)n_pairs * kWsRingDoubles);   // [pairs]
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[(k / 12) * kSplineRow + k % 12] = p.spline.coef[k];
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.ds.nx; k += blockDim.x) s_gx[k] = p.ds.x[k];
    for (int k = threadIdx.x; k < p.ds.ny; k += blockDim.x) s_gy[k] = p.ds.y[k];
    ::mms::fm::math_tables_init();
    if (kWsRing > 0 && (int)threadIdx.x < n_pairs) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(s_ring + (size_t)threadIdx.x * kWsRingDoubles + kWsSlots * kWsK * kWsFields * 32);
        for (int k = 0; k < 2 * kWsSlots; ++k) mbar_init(bars + k, 32);    // every lane of the arriving warp arrives
    }
    __syncthreads();
    SmemTables tb{s_x, s_coef, s_gx, s_gy};

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp w runs on SM sub-partition w % 4: sub-partitions 0, 1 host the forcing warps and 2, 3
    // the state warps, so that each sub-partition's L0 instruction cache sees one loop only
    // (an odd last pair takes the two leftover warps)
    // (measured: mixing both roles on every sub-partition instead costs 1.5 %)
    const bool tail = (n_pairs & 1) && warp >= 2 * (n_pairs - 1);
    const int pair = tail ? n_pairs - 1 : (warp & 1) + ((warp >> 2) << 1);
    const bool forcing_role = tail ? (warp & 1) == 0 : (warp & 2) == 0;
    RingEnd ring = ring_end(s_ring + (size_t)pair * kWsRingDoubles, pair, lane);

    for (;;) {
        if (forcing_role && lane == 0) s_unit[pair] = atomicAdd(p.unit_counter, 1);
        pair_barrier(pair);
        const int unit = s_unit[pair];
        pair_barrier(pair);                          // both have read it before it is overwritten
        if (unit >= p.n_units) break;
        const int cb = unit % p.n_cell_blocks, g = unit / p.n_cell_blocks;

        UnitCtx u;
        u.cell = (int64_t)cb * 32 + lane; u.cs = p.c_stride; u.set0 = g * kNS; u.H = 24 * p.D;
        u.live = u.cell < p.n_cells; u.write_forcing = g == 0;
        u.os = p.out_stride;
        u.ocell = (p.out_col && u.live) ? __ldg(p.out_col + u.cell) : u.cell;
        u.n_mean_state = __popc(p.mean_mask & MMS_STATE_VAR_MASK);
        u.n_hourly_state = __popc(p.hourly_mask & MMS_STATE_VAR_MASK);
        role_unit<kNS, kFull, kHourly, kDens, kWsSplit>(p, tb, u, ring, forcing_role);
    }
}


// ---- fan-out form: ONE forcing warp feeds the state warps of ALL settings of its 32 cells -------------------------
// For habitat ensembles (BASELINE configs[4]: 64 settings per cell): everything the forcing warp computes --
// microclimate, evaporation rate, life-history tables, ds^(1/24), emergence, the density-free link -- does not depend
// on the setting, yet the other two forms recompute it for every group of kNS settings (16 times
// per cell-hour for 64 settings).  Here a team = 1 forcing warp + `groups` state warps (kNS settings each) works on
// one cell block; every hour all state warps read the SAME ring slot (a lane reads its own cell's six values: the
// same addresses in every warp, no bank conflict) behind one named barrier of (1 + groups) * 32 threads.
// A CTA holds `teams` teams (one for 64 settings: 17 warps, <= 112 registers per thread).
constexpr int kFanMaxWarps = kChainFanWarps;
template <int kNS, bool kFull, bool kHourly, int kDens>
__global__ void __launch_bounds__(kFanMaxWarps * 32, 1)
chain_kernel_fan(const __grid_constant__ ChainParams p, int groups)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_coef = reinterpret_cast<double *>(smem_raw);              // [n][3][4]
    double *s_x = s_coef + (size_t)p.spline.n * kSplineRow;                     // [n]
    double *s_gx = s_x + p.spline.n;                                    // [nx]
    double *s_gy = s_gx + p.ds.nx;                                      // [ny]
    double *s_ring = s_gy + p.ds.ny;
    s_ring += (reinterpret_cast<uintptr_t>(s_ring) >> 3) & 1;           // 16-byte aligned (cp.async records)                                    // [teams][kWsRingDoubles]
    const int team_warps = groups + 1;
    const int n_teams = (blockDim.x >> 5) / team_warps;
    int *s_unit = reinterpret_cast<int *>(s_ring + (size_t)n_teams * kWsRingDoubles);   // [teams]
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[(k / 12) * kSplineRow + k % 12] = p.spline.coef[k];
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.ds.nx; k += blockDim.x) s_gx[k] = p.ds.x[k];
    for (int k = threadIdx.x; k < p.ds.ny; k += blockDim.x) s_gy[k] = p.ds.y[k];
    ::mms::fm::math_tables_init();
    __syncthreads();
    SmemTables tb{s_x, s_coef, s_gx, s_gy};

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = warp / team_warps, role = warp % team_warps;      // role 0 = forcing, role g + 1 = settings of group g
    if (team >= n_teams) return;
    const bool forcing_role = role == 0;
    RingEnd ring = ring_end(s_ring + (size_t)team * kWsRingDoubles, team, lane);
    ring.bar_threads = team_warps * 32;

    for (;;) {
        if (forcing_role && lane == 0) s_unit[team] = atomicAdd(p.unit_counter, 1);
        pair_barrier(team, ring.bar_threads);
        const int cb = s_unit[team];                 // a unit of this form = one cell block, all settings
        pair_barrier(team, ring.bar_threads);        // everybody has read it before it is overwritten
        if (cb >= p.n_cell_blocks) break;
        UnitCtx u;
        u.cell = (int64_t)cb * 32 + lane; u.cs = p.c_stride; u.set0 = forcing_role ? 0 : (role - 1) * kNS; u.H = 24 * p.D;
        u.live = u.cell < p.n_cells; u.write_forcing = true;
        u.os = p.out_stride;
        u.ocell = (p.out_col && u.live) ? __ldg(p.out_col + u.cell) : u.cell;
        u.n_mean_state = __popc(p.mean_mask & MMS_STATE_VAR_MASK);
        u.n_hourly_state = __popc(p.hourly_mask & MMS_STATE_VAR_MASK);
        role_unit<kNS, kFull, kHourly, kDens, 0>(p, tb, u, ring, forcing_role);
    }
}

}  // namespace MMS_MODE
}  // namespace mms

#endif  // !MMS_STRICT
