/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's Legendre helpers:
 *   calc_pn               src/legendre.F90:349-432
 *   calc_int_pn_tablelin  src/legendre.F90:22-336
 *
 * Arithmetic notes (all of them matter for parity at the 1e-9 level, because the closed forms
 * below cancel catastrophically when xhigh-xlow is small):
 *   - compile with -O2 -ffp-contract=off (x86-64 gfortran does not contract to FMA);
 *   - x**n is written PW(x,n) = __builtin_powi(x,n): with optimisation on GCC's middle end
 *     expands it with its addition-chain table exactly as it does for gfortran's x**n, so the
 *     multiplication order is the one a `gfortran -O3` build of the reference uses;
 *   - operator order/parentheses follow the Fortran text (left-to-right for equal precedence).
 *
 * Only orders 0..10 are restated: MAX_LEGENDRE_ORDER = 10 (src/constants.F90:113, enforced at
 * src/ndpp.F90:291-294) makes the reference's cases 11..20 unreachable.  Case 9 of
 * calc_int_pn_tablelin is, as in the reference (legendre.F90:117-126), the same expression as
 * case 7 -- a reference quirk reproduced on purpose.
 */
#include "ndpp_oracle.h"

#define PW(x, n) __builtin_powi((x), (n))

/* src/legendre.F90:349-432 */
double ref_calc_pn(int n, double x)
{
    switch (n) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return 1.5 * x * x - 0.5;
    case 3: return 2.5 * x * x * x - 1.5 * x;
    case 4: return 4.375 * PW(x, 4) - 3.75 * x * x + 0.375;
    case 5: return 7.875 * PW(x, 5) - 8.75 * x * x * x + 1.875 * x;
    case 6: return 14.4375 * PW(x, 6) - 19.6875 * PW(x, 4) + 6.5625 * x * x - 0.3125;
    case 7: return 26.8125 * PW(x, 7) - 43.3125 * PW(x, 5) + 19.6875 * x * x * x - 2.1875 * x;
    case 8:
        return 50.2734375 * PW(x, 8) - 93.84375 * PW(x, 6) + 54.140625 * PW(x, 4) -
               9.84375 * x * x + 0.2734375;
    case 9:
        return 94.9609375 * PW(x, 9) - 201.09375 * PW(x, 7) + 140.765625 * PW(x, 5) -
               36.09375 * x * x * x + 2.4609375 * x;
    case 10:
        return 180.42578125 * PW(x, 10) - 427.32421875 * PW(x, 8) + 351.9140625 * PW(x, 6) -
               117.3046875 * PW(x, 4) + 13.53515625 * x * x - 0.24609375;
    default:
        /* orders 11..20 exist in the reference but are unreachable (MAX_LEGENDRE_ORDER=10);
           beyond 20 the reference returns ONE. */
        return 1.0;
    }
}

/* src/legendre.F90:22-336.  integrals[0..n-1] receive l = 0..n-1. */
void ref_calc_int_pn_tablelin(int n, double xlow, double xhigh, double flow, double fhigh,
                              double *integrals)
{
    const double ONE = 1.0, TWO = 2.0;
    int l;
    double values;

    for (l = 0; l < n; ++l) integrals[l] = 0.0;
    /* legendre.F90:44 */
    if (xhigh - xlow < REF_FP_PRECISION) return;

    for (l = 0; l < n; ++l) {
        switch (l) {
        case 0:
            values = 0.5 * ((fhigh + flow) * PW(xlow, 2) - TWO * flow * xhigh * xlow) / (xhigh - xlow) +
                     0.5 * ((fhigh + flow) * PW(xhigh, 2) - TWO * fhigh * xhigh * xlow) / (xhigh - xlow);
            break;
        case 1:
            values = (ONE / 6.0 * ((TWO * fhigh + flow) * PW(xhigh, 3) - 3.0 * fhigh * PW(xhigh, 2) * xlow) /
                          (xhigh - xlow) +
                      ONE / 6.0 * ((fhigh + TWO * flow) * PW(xlow, 3) - 3.0 * flow * xhigh * PW(xlow, 2)) /
                          (xhigh - xlow));
            break;
        case 2:
            values = ONE / 8.0 *
                         ((3.0 * fhigh + flow) * PW(xhigh, 4) - 2.0 * (fhigh + flow) * PW(xhigh, 2) -
                          4.0 * (fhigh * PW(xhigh, 3) - fhigh * xhigh) * xlow) /
                         (xhigh - xlow) +
                     ONE / 8.0 *
                         ((fhigh + 3.0 * flow) * PW(xlow, 4) - 4.0 * flow * xhigh * PW(xlow, 3) -
                          2.0 * (fhigh + flow) * PW(xlow, 2) + 4.0 * flow * xhigh * xlow) /
                         (xhigh - xlow);
            break;
        case 3:
            values = (ONE / 8.0 *
                          ((4.0 * fhigh + flow) * PW(xhigh, 5) - 2.0 * (2.0 * fhigh + flow) * PW(xhigh, 3) -
                           (5.0 * fhigh * PW(xhigh, 4) - 6.0 * fhigh * PW(xhigh, 2)) * xlow) /
                          (xhigh - xlow) +
                      ONE / 8.0 *
                          ((fhigh + 4.0 * flow) * PW(xlow, 5) - 5.0 * flow * xhigh * PW(xlow, 4) -
                           2.0 * (fhigh + 2.0 * flow) * PW(xlow, 3) + 6.0 * flow * xhigh * PW(xlow, 2)) /
                          (xhigh - xlow));
            break;
        case 4:
            values = ONE / 48.0 *
                         (7.0 * (5.0 * fhigh + flow) * PW(xhigh, 6) - 15.0 * (3.0 * fhigh + flow) * PW(xhigh, 4) +
                          9.0 * (fhigh + flow) * PW(xhigh, 2) -
                          6.0 * (7.0 * fhigh * PW(xhigh, 5) - 10.0 * fhigh * PW(xhigh, 3) + 3.0 * fhigh * xhigh) *
                              xlow) /
                         (xhigh - xlow) +
                     ONE / 48.0 *
                         (7.0 * (fhigh + 5.0 * flow) * PW(xlow, 6) - 42.0 * flow * xhigh * PW(xlow, 5) -
                          15.0 * (fhigh + 3.0 * flow) * PW(xlow, 4) + 60.0 * flow * xhigh * PW(xlow, 3) +
                          9.0 * (fhigh + flow) * PW(xlow, 2) - 18.0 * flow * xhigh * xlow) /
                         (xhigh - xlow);
            break;
        case 5:
            values = ONE / 16.0 *
                         (3.0 * (6.0 * fhigh + flow) * PW(xhigh, 7) - 7.0 * (4.0 * fhigh + flow) * PW(xhigh, 5) +
                          5.0 * (2.0 * fhigh + flow) * PW(xhigh, 3) -
                          (21.0 * fhigh * PW(xhigh, 6) - 35.0 * fhigh * PW(xhigh, 4) +
                           15.0 * fhigh * PW(xhigh, 2)) *
                              xlow) /
                         (xhigh - xlow) +
                     ONE / 16.0 *
                         (3.0 * (fhigh + 6.0 * flow) * PW(xlow, 7) - 21.0 * flow * xhigh * PW(xlow, 6) -
                          7.0 * (fhigh + 4.0 * flow) * PW(xlow, 5) + 35.0 * flow * xhigh * PW(xlow, 4) +
                          5.0 * (fhigh + 2.0 * flow) * PW(xlow, 3) - 15.0 * flow * xhigh * PW(xlow, 2)) /
                         (xhigh - xlow);
            break;
        case 6:
            values = (ONE / 128.0 *
                          (33.0 * (7.0 * fhigh + flow) * PW(xhigh, 8) - 84.0 * (5.0 * fhigh + flow) * PW(xhigh, 6) +
                           70.0 * (3.0 * fhigh + flow) * PW(xhigh, 4) - 20.0 * (fhigh + flow) * PW(xhigh, 2) -
                           8.0 *
                               (33.0 * fhigh * PW(xhigh, 7) - 63.0 * fhigh * PW(xhigh, 5) +
                                35.0 * fhigh * PW(xhigh, 3) - 5.0 * fhigh * xhigh) *
                               xlow) /
                          (xhigh - xlow) +
                      ONE / 128.0 *
                          (33.0 * (fhigh + 7.0 * flow) * PW(xlow, 8) - 264.0 * flow * xhigh * PW(xlow, 7) -
                           84.0 * (fhigh + 5.0 * flow) * PW(xlow, 6) + 504.0 * flow * xhigh * PW(xlow, 5) +
                           70.0 * (fhigh + 3.0 * flow) * PW(xlow, 4) - 280.0 * flow * xhigh * PW(xlow, 3) -
                           20.0 * (fhigh + flow) * PW(xlow, 2) + 40.0 * flow * xhigh * xlow) /
                          (xhigh - xlow));
            break;
        case 7:
        case 9: /* reference quirk: legendre.F90:117-126 repeats the l=7 expression for l=9 */
            values = (ONE / 384.0 *
                          (143.0 * (8.0 * fhigh + flow) * PW(xhigh, 9) - 396.0 * (6.0 * fhigh + flow) * PW(xhigh, 7) +
                           378.0 * (4.0 * fhigh + flow) * PW(xhigh, 5) - 140.0 * (2.0 * fhigh + flow) * PW(xhigh, 3) -
                           3.0 *
                               (429.0 * fhigh * PW(xhigh, 8) - 924.0 * fhigh * PW(xhigh, 6) +
                                630.0 * fhigh * PW(xhigh, 4) - 140.0 * fhigh * PW(xhigh, 2)) *
                               xlow) /
                          (xhigh - xlow) +
                      ONE / 384.0 *
                          (143.0 * (fhigh + 8.0 * flow) * PW(xlow, 9) - 1287.0 * flow * xhigh * PW(xlow, 8) -
                           396.0 * (fhigh + 6.0 * flow) * PW(xlow, 7) + 2772.0 * flow * xhigh * PW(xlow, 6) +
                           378.0 * (fhigh + 4.0 * flow) * PW(xlow, 5) - 1890.0 * flow * xhigh * PW(xlow, 4) -
                           140.0 * (fhigh + 2.0 * flow) * PW(xlow, 3) + 420.0 * flow * xhigh * PW(xlow, 2)) /
                          (xhigh - xlow));
            break;
        case 8:
            values = (ONE / 256.0 *
                          (143.0 * (9.0 * fhigh + flow) * PW(xhigh, 10) - 429.0 * (7.0 * fhigh + flow) * PW(xhigh, 8) +
                           462.0 * (5.0 * fhigh + flow) * PW(xhigh, 6) - 210.0 * (3.0 * fhigh + flow) * PW(xhigh, 4) +
                           35.0 * (fhigh + flow) * PW(xhigh, 2) -
                           2.0 *
                               (715.0 * fhigh * PW(xhigh, 9) - 1716.0 * fhigh * PW(xhigh, 7) +
                                1386.0 * fhigh * PW(xhigh, 5) - 420.0 * fhigh * PW(xhigh, 3) + 35.0 * fhigh * xhigh) *
                               xlow) /
                          (xhigh - xlow) +
                      ONE / 256.0 *
                          (143.0 * (fhigh + 9.0 * flow) * PW(xlow, 10) - 1430.0 * flow * xhigh * PW(xlow, 9) -
                           429.0 * (fhigh + 7.0 * flow) * PW(xlow, 8) + 3432.0 * flow * xhigh * PW(xlow, 7) +
                           462.0 * (fhigh + 5.0 * flow) * PW(xlow, 6) - 2772.0 * flow * xhigh * PW(xlow, 5) -
                           210.0 * (fhigh + 3.0 * flow) * PW(xlow, 4) + 840.0 * flow * xhigh * PW(xlow, 3) +
                           35.0 * (fhigh + flow) * PW(xlow, 2) - 70.0 * flow * xhigh * xlow) /
                          (xhigh - xlow));
            break;
        case 10:
            values = (ONE / 3072.0 *
                          (4199.0 * (11.0 * fhigh + flow) * PW(xhigh, 12) -
                           14586.0 * (9.0 * fhigh + flow) * PW(xhigh, 10) +
                           19305.0 * (7.0 * fhigh + flow) * PW(xhigh, 8) -
                           12012.0 * (5.0 * fhigh + flow) * PW(xhigh, 6) +
                           3465.0 * (3.0 * fhigh + flow) * PW(xhigh, 4) - 378.0 * (fhigh + flow) * PW(xhigh, 2) -
                           12.0 *
                               (4199.0 * fhigh * PW(xhigh, 11) - 12155.0 * fhigh * PW(xhigh, 9) +
                                12870.0 * fhigh * PW(xhigh, 7) - 6006.0 * fhigh * PW(xhigh, 5) +
                                1155.0 * fhigh * PW(xhigh, 3) - 63.0 * fhigh * xhigh) *
                               xlow) /
                          (xhigh - xlow) +
                      ONE / 3072.0 *
                          (4199.0 * (fhigh + 11.0 * flow) * PW(xlow, 12) - 50388.0 * flow * xhigh * PW(xlow, 11) -
                           14586.0 * (fhigh + 9.0 * flow) * PW(xlow, 10) + 145860.0 * flow * xhigh * PW(xlow, 9) +
                           19305.0 * (fhigh + 7.0 * flow) * PW(xlow, 8) - 154440.0 * flow * xhigh * PW(xlow, 7) -
                           12012.0 * (fhigh + 5.0 * flow) * PW(xlow, 6) + 72072.0 * flow * xhigh * PW(xlow, 5) +
                           3465.0 * (fhigh + 3.0 * flow) * PW(xlow, 4) - 13860.0 * flow * xhigh * PW(xlow, 3) -
                           378.0 * (fhigh + flow) * PW(xlow, 2) + 756.0 * flow * xhigh * xlow) /
                          (xhigh - xlow));
            break;
        default:
            values = ONE; /* l >= 11: unreachable for order <= MAX_LEGENDRE_ORDER */
            break;
        }
        integrals[l] = integrals[l] + values;
    }
}
