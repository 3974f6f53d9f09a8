// mms_fastmath.cuh -- double-precision exp / log for the "fast" math mode.
//
// Why not CUDA's libm here: (1) its exp()/log() materialise every polynomial
// coefficient with UMOV / IMAD.MOV pairs, which more than doubled the issued
// instruction count of the fused kernel, and (2) its special-value handling is
// a branch, and in this model special values are the common case (a dry water
// body gives exp(+huge), log(0); a cell outside the ds grid gives NaN), so some
// lane of nearly every warp took the slow path and the branches split the hour
// loop into small basic blocks that the scheduler could not interleave
// (profiles/r01_ncu_chain_v1.txt).  Everything here is branch free: range and
// NaN/Inf tests are done on the integer bit pattern (ALU pipe, not the FP64
// pipe), the polynomial always runs on a harmless argument, and the special
// result is selected at the end.
//
// Accuracy ~1 ulp (polynomial error 1.6e-17 / 4.7e-18 relative,
// tools/gen_math_coeffs.py).  Saturation policy of the *_sat forms, documented
// in DESIGN.md: exp saturates to 0 / +Inf for |x| >= 708 (libm: -745.1 / 709.78),
// log treats positive denormals as DBL_MIN.
#pragma once

#include <cuda_runtime.h>

namespace mms {
namespace fm {

// exp(r) = 1 + r + r^2 q(r), |r| <= ln2/2, q of degree 9
static __constant__ double kExpQ[10] = {
    0.5000000000000001,     0.16666666666666669,    0.04166666666662413,    0.008333333333330062,
    0.0013888888917213717,  0.00019841269863053618, 2.4801521295954376e-05, 2.7557268459997064e-06,
    2.7620088445409746e-07, 2.510038549551032e-08,
};
// log(m) = 2 f + f s Q(s), f = (m-1)/(m+1), s = f^2, m in [sqrt(1/2), sqrt(2)), Q of degree 6
static __constant__ double kLogQ[7] = {
    0.666666666666667,   0.3999999999989889,  0.28571428626253836, 0.22222211089379496,
    0.1818289243253687,  0.15331611722055008, 0.14617807153417325,
};
static __constant__ double kLn2Hi = 0.6931471803691238;       // 21 trailing zero bits: n * hi is exact
static __constant__ double kLn2Lo = 1.9082149292705877e-10;
static __constant__ double kLog2e = 1.4426950408889634;

__device__ __forceinline__ double rcp_approx(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));     // MUFU.RCP64H, ~20 bits
    return r;
}

// a / b for normal, finite b and a (no special-value handling): MUFU + 7 DFMA, no branch
__device__ __forceinline__ double div_core(double a, double b)
{
    double r = rcp_approx(b);
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// c ? a : b as a real SELP: written in PTX so that the compiler cannot turn the select back
// into a branch around the (expensive) computation of one operand
__device__ __forceinline__ double sel(bool c, double a, double b)
{
    double r;
    asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
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
    const double t = fma(x, kLog2e, magic);
    const int n = __double2loint(t);                          // rint(x * log2(e))
    const double nf = t - magic;
    double r = fma(nf, -kLn2Hi, x);
    r = fma(nf, -kLn2Lo, r);
    // even / odd split of the degree-9 polynomial: two independent Horner chains in r^2
    // (the FP64 pipe needs four independent DFMAs in flight per warp; ptxas interleaves two
    // call sites at most, so the other factor of two has to come from inside the evaluation)
    const double r2 = r * r;
    double qe = fma(kExpQ[8], r2, kExpQ[6]);
    double qo = fma(kExpQ[9], r2, kExpQ[7]);
    qe = fma(qe, r2, kExpQ[4]);
    qo = fma(qo, r2, kExpQ[5]);
    qe = fma(qe, r2, kExpQ[2]);
    qo = fma(qo, r2, kExpQ[3]);
    qe = fma(qe, r2, kExpQ[0]);
    qo = fma(qo, r2, kExpQ[1]);
    const double q = fma(qo, r, qe);
    const double p = fma(q, r2, r);                           // exp(r) - 1
    const double res = 1.0 + p;
    return __hiloint2double(__double2hiint(res) + (n << 20), __double2loint(res));
}

// exp(x), any x, branch free: |x| >= 708 saturates to 0 / +Inf, NaN -> NaN
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
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                      // m in [1, 2)
    const bool up = hi >= 0x3ff6a09f;                         // m >= sqrt(2): halve it
    hi = up ? hi - 0x00100000 : hi;
    e = up ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double num = m - 1.0, den = m + 1.0;
    const double f = div_core(num, den);
    const double s = f * f;
    double q = kLogQ[6];
    q = fma(q, s, kLogQ[5]);
    q = fma(q, s, kLogQ[4]);
    q = fma(q, s, kLogQ[3]);
    q = fma(q, s, kLogQ[2]);
    q = fma(q, s, kLogQ[1]);
    q = fma(q, s, kLogQ[0]);
    const double ef = (double)e;
    const double small = fma(f * s, q, ef * kLn2Lo);
    return fma(ef, kLn2Hi, fma(2.0, f, small));
}

// log(x), any x, branch free: +-0 -> -Inf, x < 0 or NaN -> NaN, +Inf -> +Inf,
// positive denormal -> log(DBL_MIN)
__device__ __forceinline__ double log_sat(double x)
{
    const int hi = __double2hiint(x);
    const bool ok = (unsigned)(hi - 0x00100000) < 0x7fe00000u;      // normal, positive, finite
    const double r = log_core(sel(ok, x, 1.0));
    const bool zero = (((unsigned)hi & 0x7fffffffu) | (unsigned)__double2loint(x)) == 0u;
    const bool pinf = hi == 0x7ff00000 && __double2loint(x) == 0;
    double special = (hi < 0 || is_nan_bits(x)) ? make_nan() : -708.3964185322641;   // negative / NaN, else denormal
    special = pinf ? make_inf() : special;
    special = zero ? -make_inf() : special;
    return sel(ok, r, special);
}

// R's x ^ y for a positive, non-integer literal y: 0 -> 0, x < 0 / NaN -> NaN (x < Inf assumed)
__device__ __forceinline__ double pow_pos(double x, double y)
{
    return exp_sat(y * log_sat(x));
}

// ---- cheap, site-specific forms (the hot loop of the fused kernel) ------------------------
// The general *_sat wrappers cost ~12-15 ALU instructions each; where the call site knows the
// range of its argument one integer range test and one or two selects are enough.

// a / b, normal finite operands, ~1.5 ulp: MUFU + 5 DFMA (no final correction step)
__device__ __forceinline__ double div_fast(double a, double b)
{
    double r = rcp_approx(b);
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    return a * r;
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

// log(x) for the kernel's own arguments: like log_core with the 1.5-ulp division
__device__ __forceinline__ double log_fastdiv(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const bool up = hi >= 0x3ff6a09f;
    hi = up ? hi - 0x00100000 : hi;
    e = up ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double f = div_fast(m - 1.0, m + 1.0);
    const double s = f * f;
    const double s2 = s * s;                                  // even / odd split, as in exp_core
    double qe = fma(kLogQ[6], s2, kLogQ[4]);
    double qo = fma(kLogQ[5], s2, kLogQ[3]);
    qe = fma(qe, s2, kLogQ[2]);
    qo = fma(qo, s2, kLogQ[1]);
    qe = fma(qe, s2, kLogQ[0]);
    const double q = fma(qo, s, qe);
    const double ef = (double)e;
    const double small = fma(f * s, q, ef * kLn2Lo);
    return fma(ef, kLn2Hi, fma(2.0, f, small));
}

// x ^ y for x in [0, 1e300] or NaN and a literal 0 < y < 1 (water volume, survival probabilities):
// 0 and positive denormals -> 0 (R: 0 resp. ~1e-200), NaN -> NaN
__device__ __forceinline__ double pow_unit(double x, double y)
{
    const double xs = sel_pos_normal(x, x, 1.0);
    const double r = exp_core(y * log_fastdiv(xs));
    // not positive-normal: keep sign/exponent/quiet bit only -> 0 stays 0, NaN stays NaN
    const double special = __hiloint2double(__double2hiint(x) & 0xfff80000, 0);
    return sel_pos_normal(x, r, special);
}

// log(x) for x in [0, +Inf] or NaN (odds, probabilities): 0 / denormal -> -Inf, +Inf -> +Inf, NaN -> NaN
__device__ __forceinline__ double log_nonneg(double x)
{
    const double r = log_fastdiv(sel_pos_normal(x, x, 1.0));
    const int hi = __double2hiint(x);
    // +Inf / +NaN: keep; negative (incl. -NaN): NaN; zero / denormal: -Inf
    const double special = (hi >= 0x7ff00000) ? x : ((hi < 0) ? make_nan() : -make_inf());
    return sel_pos_normal(x, r, special);
}

}  // namespace fm
}  // namespace mms
