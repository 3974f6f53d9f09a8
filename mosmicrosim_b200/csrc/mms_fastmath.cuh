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
// chains that are longer than the replays).  Re-measured on the round's final kernels: pair form 49.6 -> 49.8, fan-out
// form 89.8 -> 91.7, monolithic form 37.6 -> 36.4 -- so build.py turns it on for the translation unit of the
// monolithic form (chain_tu.cu) only; off by default.
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

// Every function below comes as an N-wide form over small arrays, written STAGE BY STAGE (operation k of all N
// elements, then operation k + 1 of all N ...), and the scalar form is its N = 1 case.  Why: a thread that integrates
// several habitat settings owns that many independent dependency chains, but ptxas leaves code written setting by
// setting almost serial (tools/sass_deps.py on the fan-out kernel's state loop: 46 % of the FP64 operations sat
// directly behind their producer, 30 % one instruction later) -- and every back-to-back pair waits out the 8-cycle
// FP64 latency.  Stage-major source order puts N - 1 independent operations between producer and consumer.  Same
// operations on the same values: the results do not depend on N.
#define MMS_EACH(i, n) _Pragma("unroll") for (int i = 0; i < (n); ++i)

// a / b, normal finite operands, ~1.5 ulp: MUFU + 5 DFMA, no branch (CUDA's IEEE division
// carries a slow-path branch into the hour loop)
template <int N>
__device__ __forceinline__ void div_fast_n(const double (&a)[N], const double (&b)[N], double (&q)[N])
{
    double r[N], e[N];
    MMS_EACH(i, N) r[i] = rcp_approx(b[i]);
    MMS_EACH(i, N) e[i] = fma(-b[i], r[i], 1.0);
    MMS_EACH(i, N) r[i] = fma(e[i], r[i], r[i]);
    MMS_EACH(i, N) e[i] = fma(-b[i], r[i], 1.0);
    MMS_EACH(i, N) r[i] = fma(e[i], r[i], r[i]);
    MMS_EACH(i, N) q[i] = a[i] * r[i];
}
__device__ __forceinline__ double div_fast(double a, double b)
{
    const double av[1] = {a}, bv[1] = {b};
    double q[1];
    div_fast_n<1>(av, bv, q);
    return q[0];
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
template <int N>
__device__ __forceinline__ void exp_core_n(const double (&x)[N], double (&out)[N])
{
    const double magic = 6755399441055744.0;                  // 1.5 * 2^52
    double t[N], nf[N], r[N], tj[N], r2[N], pa[N], pb[N], p[N], res[N];
    int n[N];
    MMS_EACH(i, N) t[i] = fma(x[i], kExpC[0], magic);
    MMS_EACH(i, N) { n[i] = __double2loint(t[i]); nf[i] = t[i] - magic; }   // rint(x * 32 / ln2) = 32 m + j
    MMS_EACH(i, N) r[i] = fma(nf[i], kExpC[1], x[i]);
    MMS_EACH(i, N) r[i] = fma(nf[i], kExpC[2], r[i]);
    MMS_EACH(i, N) tj[i] = exp_table(n[i] & 31);
    MMS_EACH(i, N) r2[i] = r[i] * r[i];
    // exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2 (1/24 + r/120 + r^2/720)): two short chains
    MMS_EACH(i, N) pa[i] = fma(r[i], kExpC[3], 0.5);
    MMS_EACH(i, N) pb[i] = fma(r[i], kExpC[5], kExpC[4]);
    MMS_EACH(i, N) pb[i] = fma(r2[i], kExpC[6], pb[i]);
    MMS_EACH(i, N) p[i] = fma(r2[i], pb[i], pa[i]);
    MMS_EACH(i, N) p[i] = fma(r2[i], p[i], r[i]);
    MMS_EACH(i, N) res[i] = fma(tj[i], p[i], tj[i]);
    MMS_EACH(i, N) out[i] = __hiloint2double(__double2hiint(res[i]) + ((n[i] >> 5) << 20), __double2loint(res[i]));
}
__device__ __forceinline__ double exp_core(double x)
{
    const double xv[1] = {x};
    double r[1];
    exp_core_n<1>(xv, r);
    return r[0];
}

// exp(x), any x: |x| >= 708 saturates to 0 / +Inf, NaN -> NaN
template <int N>
__device__ __forceinline__ void exp_sat_n(const double (&x)[N], double (&out)[N])
{
    double xs[N], r[N];
    bool big[N];
    MMS_EACH(i, N) {
        big[i] = ((unsigned)__double2hiint(x[i]) & 0x7fffffffu) >= 0x40862000u;   // |x| >= 708, Inf, NaN
        xs[i] = sel(big[i], 0.0, x[i]);
    }
    exp_core_n<N>(xs, r);
    MMS_EACH(i, N) {
        const double special = is_nan_bits(x[i]) ? make_nan() : ((__double2hiint(x[i]) < 0) ? 0.0 : make_inf());
        out[i] = sel(big[i], special, r[i]);
    }
}
__device__ __forceinline__ double exp_sat(double x)
{
    const double xv[1] = {x};
    double r[1];
    exp_sat_n<1>(xv, r);
    return r[0];
}

// log(x) for normal, positive, finite x (caller guarantees it)
template <int N>
__device__ __forceinline__ void log_core_n(const double (&x)[N], double (&out)[N])
{
    double ef[N], m[N], r[N], r2[N], pa[N], pb[N], l1p[N];
    double2 ct[N];
    MMS_EACH(i, N) {
        const int hi = __double2hiint(x[i]);
        ef[i] = (double)((hi >> 20) - 1023);
        ct[i] = log_table((hi >> 13) & 127);
        m[i] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x[i]));   // [1, 2)
    }
    MMS_EACH(i, N) r[i] = fma(m[i], ct[i].x, -1.0);           // |r| <= 2^-8
    MMS_EACH(i, N) r2[i] = r[i] * r[i];
    // log1p(r) = r + r^2 (-1/2 + r/3 + r^2 (-1/4 + r/5 - r^2/6)): two short chains
    MMS_EACH(i, N) pa[i] = fma(r[i], kLogC[0], -0.5);
    MMS_EACH(i, N) pb[i] = fma(r[i], kLogC[1], -0.25);
    MMS_EACH(i, N) pb[i] = fma(r2[i], kLogC[2], pb[i]);
    MMS_EACH(i, N) l1p[i] = fma(r2[i], pb[i], pa[i]);
    MMS_EACH(i, N) l1p[i] = fma(r2[i], l1p[i], r[i]);
    MMS_EACH(i, N) out[i] = fma(ef[i], kLn2Hi, ct[i].y) + fma(ef[i], kLn2Lo, l1p[i]);
}
__device__ __forceinline__ double log_core(double x)
{
    const double xv[1] = {x};
    double r[1];
    log_core_n<1>(xv, r);
    return r[0];
}

// log(x), any x: +-0 -> -Inf, x < 0 or NaN -> NaN, +Inf -> +Inf, positive denormal -> log(DBL_MIN)
template <int N>
__device__ __forceinline__ void log_sat_n(const double (&x)[N], double (&out)[N])
{
    double xs[N], r[N];
    MMS_EACH(i, N) xs[i] = sel_pos_normal(x[i], x[i], 1.0);
    log_core_n<N>(xs, r);
    MMS_EACH(i, N) {
        const int hi = __double2hiint(x[i]);
        const bool zero = (((unsigned)hi & 0x7fffffffu) | (unsigned)__double2loint(x[i])) == 0u;
        const bool pinf = hi == 0x7ff00000 && __double2loint(x[i]) == 0;
        double special = (hi < 0 || is_nan_bits(x[i])) ? make_nan() : -708.3964185322641;   // negative / NaN, else denormal
        special = pinf ? make_inf() : special;
        special = zero ? -make_inf() : special;
        out[i] = sel_pos_normal(x[i], r[i], special);
    }
}
__device__ __forceinline__ double log_sat(double x)
{
    const double xv[1] = {x};
    double r[1];
    log_sat_n<1>(xv, r);
    return r[0];
}

// R's x ^ y for a positive, non-integer literal y, any x < Inf: 0 -> 0, x < 0 / NaN -> NaN
template <int N>
__device__ __forceinline__ void pow_pos_n(const double (&x)[N], double y, double (&out)[N])
{
    double l[N];
    log_sat_n<N>(x, l);
    MMS_EACH(i, N) l[i] = y * l[i];
    exp_sat_n<N>(l, out);
}
__device__ __forceinline__ double pow_pos(double x, double y)
{
    return exp_sat(y * log_sat(x));
}

// ---- cheap, site-specific forms (the hot loop of the fused kernel) ------------------------
// The general *_sat wrappers cost ~12-15 ALU instructions each; where the call site knows the
// range of its argument one integer range test and one or two selects are enough.

// x ^ y for x in [0, 1e300] or NaN and a literal 0 < y < 1 (water volume, survival probabilities):
// 0 and positive denormals -> 0 (R: 0 resp. ~1e-200), NaN -> NaN
template <int N>
__device__ __forceinline__ void pow_unit_n(const double (&x)[N], double y, double (&out)[N])
{
    double xs[N], l[N], r[N];
    MMS_EACH(i, N) xs[i] = sel_pos_normal(x[i], x[i], 1.0);
    log_core_n<N>(xs, l);
    MMS_EACH(i, N) l[i] = y * l[i];
    exp_core_n<N>(l, r);
    MMS_EACH(i, N) {
        // not positive-normal: keep sign/exponent/quiet bit only -> 0 stays 0, NaN stays NaN
        const double special = __hiloint2double(__double2hiint(x[i]) & 0xfff80000, 0);
        out[i] = sel_pos_normal(x[i], r[i], special);
    }
}
__device__ __forceinline__ double pow_unit(double x, double y)
{
    const double xv[1] = {x};
    double r[1];
    pow_unit_n<1>(xv, y, r);
    return r[0];
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
