/* ORACLE (test infrastructure only): restatement of the two steps that follow the integrator in the
 * reference's driver (src/ndpp.F90:611-648):
 *   ref_apply_tol_scatt  <- apply_tol_scatt   src/scatt.F90:786-818
 *   ref_thin_grid        <- thin_grid_one / thin_grid_two   src/thin.F90:51-169, 175-320
 * Arrays are Fortran (order, groups, NE) column-major == C [iE][g][l]. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ndpp_oracle.h"

/* data(:, g, iE) <- 0 where 0 < data(1, g, iE) < tol, then renormalised to the original sum_g data(1, g, iE) */
void ref_apply_tol_scatt(double *data, int L, int G, int NE, double tol)
{
    int iE, g, l;
    for (iE = 0; iE < NE; ++iE) {
        double *col = data + (size_t)iE * G * L;
        double orig_total = 0.0, now = 0.0, norm;
        for (g = 0; g < G; ++g) orig_total = orig_total + col[g * L];
        for (g = 0; g < G; ++g)
            if ((col[g * L] > 0.0) && (col[g * L] < tol))
                for (l = 0; l < L; ++l) col[g * L + l] = 0.0;
        if (orig_total > 0.0) {
            for (g = 0; g < G; ++g) now = now + col[g * L];
            norm = orig_total / now;
        } else {
            norm = 0.0;
        }
        for (g = 0; g < G; ++g)
            for (l = 0; l < L; ++l) col[g * L + l] = col[g * L + l] * norm;
    }
}

/* Greedy thinning in E_in with log-x interpolation (src/thin.F90:101-103).  y2 may be NULL (thin_grid_one).
 * keep[] receives the 0-based indices of the points kept (first and last always are); returns their number.
 * *maxerr follows the reference's own update rule -- `if (error > maxerr) maxerr = abs(testval - y)` with
 * `error` the *relative* error (signed: divided by y, not |y|), evaluated in the reference's loop order;
 * *max_abs is the plain maximum of |testval - y| over all accepted tests. */
int ref_thin_grid(const double *x, const double *y1, const double *y2, int NE, int GL, const double *tokeep, int n_tokeep,
                  double tol, int *keep, double *compression, double *maxerr, double *max_abs)
{
    int num_keep = 0, klo = 0, khi = 2, k = 1, e, t;
    double merr = 0.0, mabs = 0.0;
    const int all_ok = GL * (y2 ? 2 : 1);
    if (NE < 1) { *compression = 0.0; *maxerr = 0.0; *max_abs = 0.0; return 0; }
    keep[num_keep++] = 0;
    while (khi <= NE - 1) {
        int remove_it = 0, is_keep = 0;
        const double x1 = x[klo], x2 = x[khi], xx = x[k];
        const double x_frac = 1.0 / log(x2 / x1) * log(xx / x1);
        for (t = 0; t < n_tokeep; ++t)
            if (tokeep[t] == xx) is_keep = 1;
        if (!is_keep) {
            for (e = 0; e < GL; ++e) {
                const double *ys[2] = {y1, y2};
                int m;
                for (m = 0; m < (y2 ? 2 : 1); ++m) {
                    const double a = ys[m][(size_t)klo * GL + e], b = ys[m][(size_t)khi * GL + e],
                                 y = ys[m][(size_t)k * GL + e];
                    const double testval = a + (b - a) * x_frac;
                    double error = fabs(testval - y);
                    if (y != 0.0) error = error / y;
                    if (error <= tol) {
                        remove_it = remove_it + 1;
                        if (error > merr) merr = fabs(testval - y);
                        if (fabs(testval - y) > mabs) mabs = fabs(testval - y);
                    }
                }
            }
        }
        if (remove_it != all_ok) {
            keep[num_keep++] = k;
            klo = k;
        }
        k = k + 1;
        khi = khi + 1;
    }
    if (NE > 1) keep[num_keep++] = NE - 1;
    *compression = ((double)NE - (double)num_keep) / (double)NE;
    *maxerr = merr;
    *max_abs = mabs;
    return num_keep;
}
