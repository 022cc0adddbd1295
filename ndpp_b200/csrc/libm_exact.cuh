// libm_exact.cuh -- sinh / cosh with the bits of the host's C library, for convert_file6's Law 44
// (src/scattdata_header.F90:822-831:  0.5*A/sinh(A) * (cosh(A*mu) + R*sinh(A*mu))).
//
// Why: the reference is Fortran; its sinh/cosh are the host libm's (gfortran emits plain calls), and the
// closed forms of calc_int_pn_tablelin amplify a last-bit change of a table value into ~1e-12 absolute on a
// moment.  libdevice's sinh/cosh differ from glibc's in the last bit for a few per cent
// of the arguments, so the tables are built with a restatement of the algorithm glibc runs instead.
//
// Third-party dependency restated here (absent from /root/reference, SURVEY section 8c "Third-party
// arithmetic"): GNU libc 2.39 (Ubuntu GLIBC 2.39-0ubuntu8.5), x86-64, on a CPU with FMA + AVX2:
//   sinh   sysdeps/ieee754/dbl-64/e_sinh.c   (fdlibm: expm1-based below 22, exp above; built without FMA)
//   cosh   sysdeps/ieee754/dbl-64/e_cosh.c   (fdlibm; built without FMA)
//   expm1  sysdeps/ieee754/dbl-64/s_expm1.c  (fdlibm, glibc's re-grouped polynomial); the ifunc selects the
//          copy built with -mfma -mavx2, whose contractions are written out below as explicit fma()
//   exp    sysdeps/ieee754/dbl-64/e_exp.c    (table-driven, N = 128, degree-5 polynomial), FMA copy likewise
// The published algorithms are restated from their descriptions; the choice of contractions follows the
// instruction stream of the functions the ifunc resolvers pick on this class of CPU.  The 2^(i/128) table
// is regenerated from its definition (scripts/gen_exp_table.py).  Bit equality with the running libm is
// *measured*, not assumed: tests/test_libm_exact.py (host build of this header, >= 1e8 arguments per
// function) and ndppgpu_check_libm (device against host, include/ndppgpu.h).
//
// Everything is compiled with contraction off (-fmad=false / -ffp-contract=off): only the fma() written
// here is fused.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NDPP_LM_HD __host__ __device__ __forceinline__
#else
#define NDPP_LM_HD static inline
#endif

namespace ndpp {
namespace lm {

#ifdef __CUDACC__
__device__ const uint64_t d_exp_tab[256] = {
#include "exp_table.inc"
};
#endif
static const uint64_t h_exp_tab[256] = {
#include "exp_table.inc"
};

NDPP_LM_HD uint64_t tab(int i)
{
#ifdef __CUDA_ARCH__
    return __ldg(&d_exp_tab[i]);
#else
    return h_exp_tab[i];
#endif
}

#ifdef __CUDACC__
__device__ const uint64_t d_log_tab[274] = {
#include "log_table.inc"
};
#endif
static const uint64_t h_log_tab[274] = {
#include "log_table.inc"
};
NDPP_LM_HD uint64_t ltab(int i)
{
#ifdef __CUDA_ARCH__
    return __ldg(&d_log_tab[i]);
#else
    return h_log_tab[i];
#endif
}

NDPP_LM_HD uint64_t as_u64(double x)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
NDPP_LM_HD double as_f64(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
NDPP_LM_HD double fma_(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
NDPP_LM_HD uint32_t hi_word(double x) { return (uint32_t)(as_u64(x) >> 32); }
NDPP_LM_HD double with_hi_word(double x, uint32_t hi) { return as_f64((as_u64(x) & 0xffffffffull) | ((uint64_t)hi << 32)); }
NDPP_LM_HD double abs_(double x) { return as_f64(as_u64(x) & 0x7fffffffffffffffull); }

// exp(x), finite x; the callers below pass |x| only.
NDPP_LM_HD double exp_(double x)
{
    const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p52;
    const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    uint32_t abstop = (uint32_t)(as_u64(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {                       // |x| < 2^-54 or |x| >= 512 or not finite
        if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;            // tiny
        if (abstop >= 0x409u) {                                        // |x| >= 1024, inf, nan
            if (as_u64(x) == 0xfff0000000000000ull) return 0.0;
            if (abstop >= 0x7ffu) return 1.0 + x;
            return (as_u64(x) >> 63) ? 0.0 : as_f64(0x7ff0000000000000ull);  // underflow / overflow
        }
        abstop = 0;                                                    // 512 <= |x| < 1024: scale in two steps
    }
    double kd = fma_(x, InvLn2N, Shift);
    const uint64_t ki = as_u64(kd);
    kd -= Shift;
    const double r = fma_(kd, NegLn2loN, fma_(kd, NegLn2hiN, x));
    const int idx = 2 * (int)(ki % 128);
    const uint64_t top = ki << 45;
    const double tail = as_f64(tab(idx));
    uint64_t sbits = tab(idx + 1) + top;
    const double r2 = r * r;
    const double tmp = fma_(r2 * r2, fma_(r, C5, C4), fma_(fma_(r, C3, C2), r2, tail + r));
    if (abstop == 0) {
        if ((ki & 0x80000000ull) == 0) {                  // k > 0: the exponent of scale may overflow
            sbits -= 1009ull << 52;
            const double scale = as_f64(sbits);
            return 0x1p1009 * fma_(scale, tmp, scale);
        }
        sbits += 1022ull << 52;                           // k < 0: result may be subnormal
        const double scale = as_f64(sbits);
        double y = scale + scale * tmp;
        if (y < 1.0) {
            double lo = scale - y + scale * tmp;
            const double hi = 1.0 + y;
            lo = 1.0 - hi + y + lo;
            y = (hi + lo) - 1.0;
            if (y == 0.0) y = 0.0;
        }
        return 0x1p-1022 * y;
    }
    const double scale = as_f64(sbits);
    return fma_(scale, tmp, scale);
}

// expm1(x), finite x.
NDPP_LM_HD double expm1_(double x)
{
    const double o_threshold = 7.09782712893383973096e+02;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
                 Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
    uint32_t hx = hi_word(x);
    const uint32_t xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;
    double hi, lo, c = 0.0, t, e, y;
    int k;
    if (hx >= 0x4043687Au) {                              // |x| >= 56 ln2
        if (hx >= 0x40862E42u) {
            if (hx >= 0x7ff00000u) {
                if (((hx & 0xfffffu) | (uint32_t)as_u64(x)) != 0) return x + x;
                return xsb == 0 ? x : -1.0;
            }
            if (x > o_threshold) return as_f64(0x7ff0000000000000ull);
        }
        if (xsb != 0) return 1.0e-300 - 1.0;
    }
    if (hx > 0x3fd62e42u) {                               // |x| > 0.5 ln2
        if (hx < 0x3FF0A2B2u) {                           // |x| < 1.5 ln2
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int)(invln2 * x + (xsb == 0 ? 0.5 : -0.5));
            t = (double)k;
            hi = fma_(-t, ln2_hi, x);                     // t*ln2_hi is exact
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x3c900000u) {                        // |x| < 2^-54
        t = 1.0e300 + x;
        return x - (t - (1.0e300 + x));
    } else {
        k = 0;
    }
    const double hfx = 0.5 * x;
    const double hxs = x * hfx;
    const double R2 = fma_(hxs, Q3, Q2);
    const double R3 = fma_(hxs, Q5, Q4);
    const double h2 = hxs * hxs;
    const double R1 = fma_(hxs, Q1, 1.0);
    const double h4 = h2 * h2;
    const double r1 = fma_(h4, R3, fma_(h2, R2, R1));
    t = fma_(-r1, hfx, 3.0);
    e = hxs * ((r1 - t) / fma_(-x, t, 6.0));
    if (k == 0) return x - fma_(e, x, -hxs);
    e = fma_(e - c, x, -c);
    e -= hxs;
    if (k == -1) return 0.5 * (x - e) - 0.5;
    if (k == 1) {
        if (x < -0.25) return -2.0 * (e - (x + 0.5));
        return 1.0 + 2.0 * (x - e);
    }
    if (k <= -2 || k > 56) {
        y = 1.0 - (e - x);
        y = with_hi_word(y, hi_word(y) + ((uint32_t)k << 20));
        return y - 1.0;
    }
    if (k < 20) {
        t = as_f64((uint64_t)(0x3ff00000u - (0x200000u >> k)) << 32);   // 1 - 2^-k
        y = t - (e - x);
        y = with_hi_word(y, hi_word(y) + ((uint32_t)k << 20));
    } else {
        t = as_f64((uint64_t)((uint32_t)(0x3ff - k) << 20) << 32);      // 2^-k
        y = x - (e + t);
        y += 1.0;
        y = with_hi_word(y, hi_word(y) + ((uint32_t)k << 20));
    }
    return y;
}

NDPP_LM_HD double sinh_(double x)
{
    const uint32_t jx = hi_word(x);
    const uint32_t ix = jx & 0x7fffffffu;
    if (ix >= 0x7ff00000u) return x + x;
    const double h = (jx & 0x80000000u) ? -0.5 : 0.5;
    if (ix < 0x40360000u) {                               // |x| < 22
        if (ix < 0x3e300000u) {                           // |x| < 2^-28
            if (1.0e307 + x > 1.0) return x;
        }
        const double t = expm1_(abs_(x));
        if (ix < 0x3ff00000u) return h * (2.0 * t - t * t / (t + 1.0));
        return h * (t + t / (t + 1.0));
    }
    if (ix < 0x40862e42u) return h * exp_(abs_(x));
    const uint32_t lx = (uint32_t)as_u64(x);
    if (ix < 0x408633ceu || (ix == 0x408633ceu && lx <= 0x8fb9f87du)) {
        const double w = exp_(0.5 * abs_(x));
        const double t = h * w;
        return t * w;
    }
    return x * 1.0e307;
}

NDPP_LM_HD double cosh_(double x)
{
    const uint32_t ix = hi_word(x) & 0x7fffffffu;
    if (ix < 0x40360000u) {                               // |x| < 22
        if (ix < 0x3fd62e43u) {                           // |x| < 0.5 ln2
            if (ix < 0x3c800000u) return 1.0;
            const double t = expm1_(abs_(x));
            const double w = 1.0 + t;
            return 1.0 + (t * t) / (w + w);
        }
        const double t = exp_(abs_(x));
        return 0.5 * t + 0.5 / t;
    }
    if (ix < 0x40862e42u) return 0.5 * exp_(abs_(x));
    if ((as_u64(x) & 0x7fffffffffffffffull) <= 0x408633ce8fb9f87dull) {
        const double w = exp_(0.5 * abs_(x));
        const double t = 0.5 * w;
        return t * w;
    }
    if (ix >= 0x7ff00000u) return x * x;
    return 1.0e300 * 1.0e300;
}

// log(x): sysdeps/ieee754/dbl-64/e_log.c (table-driven, N = 128: z = x / 2^k in [0x1.6p-1, 0x1.6p0), c near the centre
// of z's sub-interval, r = z/c - 1 by one fma with the table's 1/c, degree-5 polynomial A; inside [1 - 2^-4, 1 + 0x1.09p-4)
// the degree-11 polynomial B with a double-double head).  FMA copy: the contractions below are those of the function the
// ifunc resolver picks on this class of CPU, read off its instruction stream.  Constants: log_table.inc
// (scripts/gen_log_table.py).  The incoming-energy grid builders (src/scatt.F90:311-536, src/sab.F90:548-566) place their
// points with log and exp of the host libm.
NDPP_LM_HD double log_(double x)
{
    uint64_t ix = as_u64(x);
    const double Ln2hi = as_f64(ltab(0)), Ln2lo = as_f64(ltab(1));
    if (ix - 0x3fee000000000000ull < 0x0003090000000000ull) {            // 1 - 2^-4 <= x < 1 + 0x1.09p-4
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = x - 1.0;
        const double B0 = as_f64(ltab(7)), B1 = as_f64(ltab(8)), B2 = as_f64(ltab(9)), B3 = as_f64(ltab(10)),
                     B4 = as_f64(ltab(11)), B5 = as_f64(ltab(12)), B6 = as_f64(ltab(13)), B7 = as_f64(ltab(14)),
                     B8 = as_f64(ltab(15)), B9 = as_f64(ltab(16)), B10 = as_f64(ltab(17));
        const double r2 = r * r, r3 = r * r2;
        const double p123 = fma_(r2, B3, fma_(r, B2, B1));
        const double p456 = fma_(r2, B6, fma_(r, B5, B4));
        double p = fma_(r3, B10, fma_(r2, B9, fma_(r, B8, B7)));
        p = fma_(p, r3, p456);
        p = fma_(p, r3, p123);
        const double rw = fma_(r, 0x1p27, r);                             // r + r * 2^27
        const double rhi = fma_(-0x1p27, r, rw);
        const double rlo = r - rhi;
        const double rhi2 = rhi * rhi;
        const double hi = fma_(rhi2, B0, r);
        double lo = fma_(rhi2, B0, r - hi);
        lo = fma_(B0 * rlo, rhi + r, lo);
        return fma_(p, r3, lo) + hi;
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return as_f64(0xfff0000000000000ull);             // log(+-0) = -inf
        if (ix == 0x7ff0000000000000ull) return x;                        // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return as_f64(0x7ff8000000000000ull) + (x - x);   // x < 0 or NaN -> NaN
        ix = as_u64(x * 0x1p52) - (52ull << 52);                          // subnormal: normalise
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const double z = as_f64(ix - (tmp & 0xfff0000000000000ull));
    const double invc = as_f64(ltab(18 + 2 * i)), logc = as_f64(ltab(19 + 2 * i));
    const double A0 = as_f64(ltab(2)), A1 = as_f64(ltab(3)), A2 = as_f64(ltab(4)), A3 = as_f64(ltab(5)), A4 = as_f64(ltab(6));
    const double kd = (double)k;
    const double r = fma_(z, invc, -1.0);
    const double w = fma_(kd, Ln2hi, logc);
    const double hi = w + r;
    const double lo = fma_(kd, Ln2lo, (w - hi) + r);
    const double r2 = r * r;
    const double q = fma_(fma_(r, A4, A3), r2, fma_(r, A2, A1));
    return fma_(r * r2, q, fma_(r2, A0, lo)) + hi;
}

}  // namespace lm
}  // namespace ndpp
