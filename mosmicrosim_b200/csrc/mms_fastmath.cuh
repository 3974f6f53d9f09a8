// mms_fastmath.cuh -- double-precision exp / log / division for the "fast" math mode.
//
// Why not CUDA's libm here: (1) its exp()/log() materialise every polynomial
// coefficient with UMOV / IMAD.MOV pairs, which more than doubled the issued
// instruction count of the fused kernel; (2) its special-value handling is a
// branch, and in this model special values are the common case (a dry water
// body gives exp(+huge) and log(0); a cell outside the ds grid gives NaN), so
// some lane of nearly every warp took the slow path and the branches split the
// hour loop into small basic blocks; (3) its evaluations are long dependent
// DFMA chains (B200: 8-cycle DFMA latency, one DFMA per 2 cycles per SM
// sub-partition -- tools/microbench/dfma_latency.cu -- and ptxas interleaves at
// most two independent chains), profiles/r01_ncu_chain_v1.txt.
//
// Everything here is branch free (range and NaN/Inf tests on the integer bit
// pattern: ALU pipe, not the FP64 pipe; selects written in PTX so that the
// compiler cannot turn them back into branches) and table driven (short
// chains, few FP64 operations):
//   exp: x = (32 m + j) ln2/32 + r, exp(x) = 2^m 2^(j/32) exp(r), |r| <= ln2/64,
//        Taylor degree 6 (truncation 3.4e-18 relative): 11 FP64 operations;
//   log: x = 2^e m, k = top 7 mantissa bits, c_k ~ 1/m_k, r = m c_k - 1 (one FMA,
//        |r| <= 2^-8), log(x) = e ln2 - log(c_k) + log1p(r), Taylor degree 6
//        (truncation 2e-18 absolute): 11 FP64 operations, no division.
// Accuracy: exp ~1 ulp; log ~1e-16 absolute (+1 ulp); division 1.5 ulp.
// Saturation policy of the guarded forms (documented in DESIGN.md, never reached
// by sane inputs): exp saturates to 0 / +Inf for |x| >= 708 (libm: -745.1 /
// 709.78); positive denormals count as zero for pow / log.
#pragma once

#include <cuda_runtime.h>

namespace mms {
namespace fm {

// constants live in __constant__ memory: the compiler then feeds them to DFMA from the constant
// bank (LDCU.128 pairs) instead of building each 64-bit literal with two UMOV / IMAD.MOV
static __constant__ double kLn2Hi = 0.6931471803691238;       // 21 trailing zero bits: n * hi is exact
static __constant__ double kLn2Lo = 1.9082149292705877e-10;
static __constant__ double kExpC[8] = {
    46.16624130844683,            // 32 / ln2
    -0.021660849386535119,        // -ln2_hi / 32
    -5.9631716539705865e-12,      // -ln2_lo / 32
    1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 0.0,
};
static __constant__ double kLogC[4] = {1.0 / 3.0, 0.2, -1.0 / 6.0, 0.0};

// The two tables (2.3 KB) live in shared memory; every kernel that uses the fast math calls math_tables_init() once
// per block before its first __syncthreads().  The lanes of a warp look up unrelated entries, and entries that share
// a bank serialise: ncu on the pair form of the fused kernel (profiles/r02_ncu_chain_ws.txt) counts 10.2 G shared-
// memory wavefronts where 4.7 G would do, the shared-memory pipe 68 % busy.  MMS_MATH_TABLE_COPIES=1 REPLICATES the
// tables across the banks -- lane l reads copy l % 16 of the exp table (an 8-byte load is served per half-warp: 16
// lanes, 16 distinct bank pairs, whatever the entries) and copy l % 8 of the log table (a 16-byte load is served per
// quarter-warp) -- 20 KB, conflict free.  MEASURED (one B200, ms per launch, one copy -> replicated): pair form on
// BASELINE configs[1] 50.53 -> 50.79, monolithic form 100k cells x 1 setting 41.03 -> 40.87, fan-out form 9472 cells x 64
// settings 104.5 -> 101.9: the conflicts are not what the kernels wait for (the look-ups sit inside FP64 dependency
// chains that are longer than the replays).  Off by default (the shared memory is better spent elsewhere).
#ifndef MMS_MATH_TABLE_COPIES
#define MMS_MATH_TABLE_COPIES 0    // 1: replicated, conflict-free tables
#endif
constexpr int kExpCopies = MMS_MATH_TABLE_COPIES ? 16 : 1, kLogCopies = MMS_MATH_TABLE_COPIES ? 8 : 1;
struct MathTables {
    double2 logt[128 * kLogCopies];    // [k][copy]: (c_k, -log c_k)
    double expt[32 * kExpCopies];      // [j][copy]: 2^(j/32)
};

__device__ __forceinline__ MathTables &math_tables()
{
    __shared__ MathTables t;
    return t;
}

__device__ __forceinline__ void math_tables_init()
{
    MathTables &t = math_tables();
    for (int i = threadIdx.x; i < 128 * kLogCopies; i += blockDim.x) {
        const int k = i / kLogCopies;
        const double c = 1.0 / (1.0 + (k + 0.5) / 128.0);
        t.logt[i] = make_double2(c, -log(c));
    }
    for (int i = threadIdx.x; i < 32 * kExpCopies; i += blockDim.x) t.expt[i] = exp2((i / kExpCopies) / 32.0);
}

__device__ __forceinline__ double exp_table(int j)
{
    return math_tables().expt[j * kExpCopies + (int)(threadIdx.x & (kExpCopies - 1))];
}
__device__ __forceinline__ double2 log_table(int k)
{
    return math_tables().logt[k * kLogCopies + (int)(threadIdx.x & (kLogCopies - 1))];
}

__device__ __forceinline__ double rcp_approx(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));     // MUFU.RCP64H, ~20 bits
    return r;
}

// a / b, normal finite operands, ~1.5 ulp: MUFU + 5 DFMA, no branch (CUDA's IEEE division
// carries a slow-path branch into the hour loop)
__device__ __forceinline__ double div_fast(double a, double b)
{
    double r = rcp_approx(b);
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    return a * r;
}

// c ? a : b as a real SELP: written in PTX so that the compiler cannot turn the select back
// into a branch around the (expensive) computation of one operand
__device__ __forceinline__ double sel(bool c, double a, double b)
{
    double r;
    asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
    return r;
}

// (x is a positive, normal, finite double) ? a : b -- the test is two integer ops on the high word
__device__ __forceinline__ double sel_pos_normal(double x, double a, double b)
{
    double r;
    asm("{ .reg .pred p; .reg .u32 t; add.u32 t, %3, 0xfff00000; setp.lt.u32 p, t, 0x7fe00000;"
        " selp.f64 %0, %1, %2, p; }"
        : "=d"(r) : "d"(a), "d"(b), "r"(__double2hiint(x)));
    return r;
}

__device__ __forceinline__ double make_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double make_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ bool is_nan_bits(double x)
{
    const unsigned ahi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    return ahi > 0x7ff00000u || (ahi == 0x7ff00000u && __double2loint(x) != 0);
}

// exp(x) for |x| <= 708 (caller guarantees it)
__device__ __forceinline__ double exp_core(double x)
{
    const double magic = 6755399441055744.0;                  // 1.5 * 2^52
    const double t = fma(x, kExpC[0], magic);
    const int n = __double2loint(t);                          // rint(x * 32 / ln2) = 32 m + j
    const double nf = t - magic;
    double r = fma(nf, kExpC[1], x);
    r = fma(nf, kExpC[2], r);
    const double tj = exp_table(n & 31);
    const double r2 = r * r;
    // exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2 (1/24 + r/120 + r^2/720)): two short chains
    const double pa = fma(r, kExpC[3], 0.5);
    const double pb = fma(r2, kExpC[6], fma(r, kExpC[5], kExpC[4]));
    const double p = fma(r2, fma(r2, pb, pa), r);
    const double res = fma(tj, p, tj);
    return __hiloint2double(__double2hiint(res) + ((n >> 5) << 20), __double2loint(res));
}

// exp(x), any x: |x| >= 708 saturates to 0 / +Inf, NaN -> NaN
__device__ __forceinline__ double exp_sat(double x)
{
    const int hi = __double2hiint(x);
    const bool big = ((unsigned)hi & 0x7fffffffu) >= 0x40862000u;   // |x| >= 708, Inf, NaN
    const double r = exp_core(sel(big, 0.0, x));
    const double special = is_nan_bits(x) ? make_nan() : ((hi < 0) ? 0.0 : make_inf());
    return sel(big, special, r);
}

// log(x) for normal, positive, finite x (caller guarantees it)
__device__ __forceinline__ double log_core(double x)
{
    const int hi = __double2hiint(x);
    const double ef = (double)((hi >> 20) - 1023);
    const double2 ct = log_table((hi >> 13) & 127);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
    const double r = fma(m, ct.x, -1.0);                      // |r| <= 2^-8
    const double r2 = r * r;
    // log1p(r) = r + r^2 (-1/2 + r/3 + r^2 (-1/4 + r/5 - r^2/6)): two short chains
    const double pa = fma(r, kLogC[0], -0.5);
    const double pb = fma(r2, kLogC[2], fma(r, kLogC[1], -0.25));
    const double l1p = fma(r2, fma(r2, pb, pa), r);
    return fma(ef, kLn2Hi, ct.y) + fma(ef, kLn2Lo, l1p);
}

// log(x), any x: +-0 -> -Inf, x < 0 or NaN -> NaN, +Inf -> +Inf, positive denormal -> log(DBL_MIN)
__device__ __forceinline__ double log_sat(double x)
{
    const int hi = __double2hiint(x);
    const double r = log_core(sel_pos_normal(x, x, 1.0));
    const bool zero = (((unsigned)hi & 0x7fffffffu) | (unsigned)__double2loint(x)) == 0u;
    const bool pinf = hi == 0x7ff00000 && __double2loint(x) == 0;
    double special = (hi < 0 || is_nan_bits(x)) ? make_nan() : -708.3964185322641;   // negative / NaN, else denormal
    special = pinf ? make_inf() : special;
    special = zero ? -make_inf() : special;
    return sel_pos_normal(x, r, special);
}

// R's x ^ y for a positive, non-integer literal y, any x < Inf: 0 -> 0, x < 0 / NaN -> NaN
__device__ __forceinline__ double pow_pos(double x, double y)
{
    return exp_sat(y * log_sat(x));
}

// ---- cheap, site-specific forms (the hot loop of the fused kernel) ------------------------
// The general *_sat wrappers cost ~12-15 ALU instructions each; where the call site knows the
// range of its argument one integer range test and one or two selects are enough.

// x ^ y for x in [0, 1e300] or NaN and a literal 0 < y < 1 (water volume, survival probabilities):
// 0 and positive denormals -> 0 (R: 0 resp. ~1e-200), NaN -> NaN
__device__ __forceinline__ double pow_unit(double x, double y)
{
    const double xs = sel_pos_normal(x, x, 1.0);
    const double r = exp_core(y * log_core(xs));
    // not positive-normal: keep sign/exponent/quiet bit only -> 0 stays 0, NaN stays NaN
    const double special = __hiloint2double(__double2hiint(x) & 0xfff80000, 0);
    return sel_pos_normal(x, r, special);
}

// log(x) for x in [0, +Inf] or NaN (odds, probabilities): 0 / denormal -> -Inf, +Inf -> +Inf,
// negative or NaN -> NaN
__device__ __forceinline__ double log_nonneg(double x)
{
    const double r = log_core(sel_pos_normal(x, x, 1.0));
    const int hi = __double2hiint(x);
    const double special = (hi >= 0x7ff00000) ? x : ((hi < 0) ? make_nan() : -make_inf());
    return sel_pos_normal(x, r, special);
}

}  // namespace fm
}  // namespace mms
