/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's free-gas thermal kernel integrator, src/freegas.F90:18-644.
 * Recursive, exactly as the Fortran text; PI is the reference's truncated constant.
 * Parity status: the reference holds no test for this module => "parity unpinned"; held in tests/ by
 * the closed-form free-gas kernel (A = 1 and general A), a double quadrature of the free-gas law for
 * P1..P3, and a literal pure-Python transcription of the Fortran text (tests/freegas_walk.py) that
 * this file reproduces bit for bit.
 * Build: gcc -O2 -ffp-contract=off (oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>

#include "ndpp_oracle.h"

#define ZERO 0.0
#define ONE 1.0
#define TWO 2.0

/* evaluation counters for the algorithmic-flop figure F_E */
static __thread long long t_n_fgk = 0, t_n_sab = 0;
static long long g_n_fgk = 0, g_n_sab = 0;

void ref_freegas_counters_flush(void)
{
#pragma omp critical(ref_fg_cnt)
    {
        g_n_fgk += t_n_fgk;
        g_n_sab += t_n_sab;
    }
    t_n_fgk = 0;
    t_n_sab = 0;
}

void ref_freegas_counters(long long *n_fgk, long long *n_sab, int reset)
{
    ref_freegas_counters_flush();
    if (n_fgk) *n_fgk = g_n_fgk;
    if (n_sab) *n_sab = g_n_sab;
    if (reset) {
        g_n_fgk = 0;
        g_n_sab = 0;
    }
}

/* src/freegas.F90:154-181 */
void ref_calc_FG_Eout_bounds(double A, double kT, double Ein, double *Eout_lo, double *Eout_hi)
{
    double alpha = ((A - ONE) / (A + ONE));
    alpha = alpha * alpha; /* (..)**2 */
    *Eout_lo = 0.001 * alpha * Ein;
    if (Ein > 300.0 * kT / A)
        *Eout_hi = 12.0 * kT * (A + ONE) / A + 1.5 * Ein;
    else
        *Eout_hi = 12.0 * kT * (A + ONE) / A + TWO * Ein;
}

/* src/freegas.F90:188-228 */
double ref_calc_sab(double A, double kT, double Ein, double Eout, double beta, double mu)
{
    const double alpha_min = 1.0E-6, sab_min = -225.0, lterm_min = 2.0E-10;
    double sab, alpha, lterm, t;

    t_n_sab++;
    t = (A + ONE) / A;
    lterm = sqrt(Eout / Ein) / kT * (t * t);
    alpha = (Ein + Eout - TWO * mu * sqrt(Ein * Eout)) / (A * kT);
    if (alpha < alpha_min) alpha = alpha_min;
    t = alpha + beta;
    sab = -(t * t) / (4.0 * alpha);
    if (sab < sab_min) {
        sab = ZERO;
    } else {
        sab = lterm * exp(sab) / (sqrt(4.0 * REF_PI * alpha));
        if (sab < lterm_min) sab = ZERO;
    }
    return sab;
}

/* src/freegas.F90:235-345 */
double ref_brent_mu(double awr, double kT, double Ein, double Eout, double beta, double thresh, double lo, double hi,
                    const ref_params *p)
{
    double a, b, c, d, fa, fb, fc, s, fs, tmpval;
    int i, mflag;
    const double BRENT_MU_THRESH = p->brent_mu_thresh;

    a = lo;
    b = hi;
    c = ZERO;
    d = REF_INFINITY;
    fa = ref_calc_sab(awr, kT, Ein, Eout, beta, a) - thresh;
    fb = ref_calc_sab(awr, kT, Ein, Eout, beta, b) - thresh;
    fc = ZERO;
    s = ZERO;
    fs = ZERO;

    if (fa * fb >= ZERO) return (fa < fb) ? a : b;

    if (fabs(fa) < fabs(fb)) {
        tmpval = a; a = b; b = tmpval;
        tmpval = fa; fa = fb; fb = tmpval;
    }
    c = a;
    fc = fa;
    mflag = 1;
    i = 0;
    while ((fb != ZERO) && (fabs(a - b) > BRENT_MU_THRESH)) {
        if ((fa != fc) && (fb != fc))
            s = a * fb * fc / (fa - fb) / (fa - fc) + b * fa * fc / (fb - fa) / (fb - fc) +
                c * fa * fb / (fc - fa) / (fc - fb);
        else
            s = b - fb * (b - a) / (fb - fa);

        tmpval = (3.0 * a + b) * 0.25;
        if ((!(((s > tmpval) && (s < b)) || ((s < tmpval) && (s > b)))) ||
            (mflag && (fabs(s - b) >= (0.5 * fabs(b - c)))) ||
            (!mflag && (fabs(s - b) >= (fabs(c - d) * 0.5)))) {
            s = 0.5 * (a + b);
            mflag = 1;
        } else {
            if ((mflag && (fabs(b - c) < BRENT_MU_THRESH)) || (!mflag && (fabs(c - d) < BRENT_MU_THRESH))) {
                s = (a + b) * 0.5;
                mflag = 1;
            } else {
                mflag = 0;
            }
        }
        fs = ref_calc_sab(awr, kT, Ein, Eout, beta, s) - thresh;
        d = c;
        c = b;
        fc = fb;
        if (fa * fs < ZERO) {
            b = s;
            fb = fs;
        } else {
            a = s;
            fa = fs;
        }
        if (fabs(fa) < fabs(fb)) {
            tmpval = a; a = b; b = tmpval;
            tmpval = fa; fa = fb; fb = tmpval;
        }
        i = i + 1;
    }
    (void)i;
    return b;
}

/* src/freegas.F90:356-409 */
void ref_find_FG_mu(double A, double kT, double Ein, double Eout, const ref_params *p, double *mu)
{
    double mu_max, beta, alpha_max, sab_max, sab_minthresh, mu_lo, mu_hi;

    beta = (Eout - Ein) / kT;
    alpha_max = sqrt(beta * beta + ONE) - ONE;
    mu_max = (Ein + Eout - alpha_max * A * kT) / (TWO * sqrt(Ein * Eout));
    if (fabs(mu_max) > ONE) {
        mu_lo = -ONE;
        mu_hi = ONE;
    } else {
        sab_max = ref_calc_sab(A, kT, Ein, Eout, beta, mu_max);
        sab_minthresh = sab_max * p->sab_threshold;
        if (ref_calc_sab(A, kT, Ein, Eout, beta, -ONE) > sab_minthresh)
            mu_lo = -ONE;
        else
            mu_lo = ref_brent_mu(A, kT, Ein, Eout, beta, sab_minthresh, -ONE, mu_max, p);
        if (ref_calc_sab(A, kT, Ein, Eout, beta, ONE) > sab_minthresh)
            mu_hi = ONE;
        else
            mu_hi = ref_brent_mu(A, kT, Ein, Eout, beta, sab_minthresh, mu_max, ONE, p);
    }
    mu[0] = mu_lo;
    mu[1] = mu_hi;
}

/* src/freegas.F90:415-473 */
double ref_calc_fgk(double awr, double kT, double Ein, double Eout, int l, double mu, const double *fEmu,
                    const double *global_mu, int M)
{
    double fgk, alpha, beta, lterm, interp, fEmu_val, dmu, t;
    int i;

    t_n_fgk++;
    dmu = A1(global_mu, 2) - A1(global_mu, 1);
    if (mu <= A1(global_mu, 1))
        i = 1;
    else if (mu >= A1(global_mu, M))
        i = M - 1;
    else
        i = (int)((mu + ONE) / dmu) + 1;
    interp = (mu - A1(global_mu, i)) / (A1(global_mu, i + 1) - A1(global_mu, i));
    fEmu_val = (ONE - interp) * A1(fEmu, i) + interp * A1(fEmu, i + 1);

    t = (awr + ONE) / awr;
    lterm = fEmu_val * sqrt(Eout / Ein) / kT * (t * t);
    alpha = (Ein + Eout - TWO * mu * sqrt(Ein * Eout)) / (awr * kT);
    beta = (Eout - Ein) / kT;
    if (alpha < 1.0E-6) alpha = 1.0E-6;
    t = alpha + beta;
    fgk = -(t * t) / (4.0 * alpha);
    if (fgk <= -708.0)
        fgk = ZERO;
    else
        fgk = lterm * exp(fgk) / (sqrt(4.0 * REF_PI * alpha)) * ref_calc_pn(l, mu);
    return fgk;
}

typedef struct {
    double awr, kT, Ein;
    int l, M;
    const double *fEmu, *gmu;
    const ref_params *p;
} fg_ctx;

/* src/freegas.F90:511-553 */
static double simpson_aux_mu(const fg_ctx *c, double Eout, double a, double b, double eps, double S, double fa,
                             double fb, double fc, int bottom)
{
    double cc, d, h, e, fd, fe, Sleft, Sright, S2;
    cc = 0.5 * (a + b);
    h = b - a;
    d = 0.5 * (a + cc);
    e = 0.5 * (cc + b);
    fd = ref_calc_fgk(c->awr, c->kT, c->Ein, Eout, c->l, d, c->fEmu, c->gmu, c->M);
    fe = ref_calc_fgk(c->awr, c->kT, c->Ein, Eout, c->l, e, c->fEmu, c->gmu, c->M);
    Sleft = (h / 12.0) * (fa + 4.0 * fd + fc);
    Sright = (h / 12.0) * (fc + 4.0 * fe + fb);
    S2 = Sleft + Sright;
    if ((bottom <= 0) || (fabs(S2 - S) <= 15.0 * eps)) return S2 + (S2 - S) / 15.0;
    return simpson_aux_mu(c, Eout, a, cc, 0.5 * eps, Sleft, fa, fc, fd, bottom - 1) +
           simpson_aux_mu(c, Eout, cc, b, 0.5 * eps, Sright, fc, fb, fe, bottom - 1);
}

/* src/freegas.F90:482-509 */
static double simpson_mu(const fg_ctx *c, double Eout, double a, double b)
{
    double cc, h, fa, fb, fc, S;
    cc = (a + b) * 0.5;
    h = (b - a);
    fa = ref_calc_fgk(c->awr, c->kT, c->Ein, Eout, c->l, a, c->fEmu, c->gmu, c->M);
    fb = ref_calc_fgk(c->awr, c->kT, c->Ein, Eout, c->l, b, c->fEmu, c->gmu, c->M);
    fc = ref_calc_fgk(c->awr, c->kT, c->Ein, Eout, c->l, cc, c->fEmu, c->gmu, c->M);
    S = (h / 6.0) * (fa + 4.0 * fc + fb);
    return simpson_aux_mu(c, Eout, a, b, c->p->adaptive_mu_tol, S, fa, fb, fc, c->p->adaptive_mu_its);
}

/* inner integral at one Eout: find_FG_mu then adaptiveSimpsons_mu (freegas.F90:582-591,625-631) */
static double inner_at(const fg_ctx *c, double Eout)
{
    double m[2];
    ref_find_FG_mu(c->awr, c->kT, c->Ein, Eout, c->p, m);
    return simpson_mu(c, Eout, m[0], m[1]);
}

/* src/freegas.F90:598-644.  The reference evaluates find_FG_mu(d), find_FG_mu(e), then fd, fe. */
static double simpson_aux_Eout(const fg_ctx *c, double a, double b, double eps, double S, double fa, double fb,
                               double fc, int bottom)
{
    double cc, d, e, h, fd, fe, Sleft, Sright, S2;
    cc = 0.5 * (a + b);
    d = 0.5 * (a + cc);
    e = 0.5 * (cc + b);
    h = b - a;
    fd = inner_at(c, d);
    fe = inner_at(c, e);
    Sleft = (h / 12.0) * (fa + 4.0 * fd + fc);
    Sright = (h / 12.0) * (fc + 4.0 * fe + fb);
    S2 = Sleft + Sright;
    if ((bottom <= 0) || (fabs(S2 - S) <= 15.0 * eps)) return S2 + (S2 - S) / 15.0;
    return simpson_aux_Eout(c, a, cc, 0.5 * eps, Sleft, fa, fc, fd, bottom - 1) +
           simpson_aux_Eout(c, cc, b, 0.5 * eps, Sright, fc, fb, fe, bottom - 1);
}

/* src/freegas.F90:563-596 */
static double simpson_Eout(const fg_ctx *c, double a, double b)
{
    double cc, h, fa, fb, fc, S;
    cc = 0.5 * (a + b);
    h = b - a;
    fa = inner_at(c, a);
    fb = inner_at(c, b);
    fc = inner_at(c, cc);
    S = (h / 6.0) * (fa + 4.0 * fc + fb);
    return simpson_aux_Eout(c, a, b, c->p->adaptive_eout_tol, S, fa, fb, fc, c->p->adaptive_eout_its);
}

/* src/freegas.F90:18-146.  distro is (order x groups) column-major. */
void ref_integrate_freegas_leg(double Ein, double A, double kT, const double *fEmu, const double *mu, int M,
                               const double *E_bins, int nbins, int order, const ref_params *p, double *distro)
{
    int g, l, groups = nbins - 1;
    double p0_1g_norm, Eout_lo, Eout_hi, Elo, Ehi, alphaEin, Ebottom;
    fg_ctx c;
#define D(l, g) distro[((l)-1) + (size_t)order * ((g)-1)]

    c.awr = A;
    c.kT = kT;
    c.Ein = Ein;
    c.M = M;
    c.fEmu = fEmu;
    c.gmu = mu;
    c.p = p;

    alphaEin = (A - ONE) / (A + ONE);
    alphaEin = alphaEin * alphaEin * Ein;
    p0_1g_norm = ZERO;
    ref_calc_FG_Eout_bounds(A, kT, Ein, &Eout_lo, &Eout_hi);

    for (g = 1; g <= groups; ++g) {
        if ((A1(E_bins, g) < Eout_hi) && (A1(E_bins, g + 1) > Eout_lo)) {
            Elo = (Eout_lo > A1(E_bins, g)) ? Eout_lo : A1(E_bins, g);
            Ehi = (Eout_hi < A1(E_bins, g + 1)) ? Eout_hi : A1(E_bins, g + 1);
            if (A1(E_bins, g) == ZERO)
                Ebottom = 0.01 * Elo;
            else
                Ebottom = A1(E_bins, g);
            for (l = 1; l <= order; ++l) {
                c.l = l - 1;
                D(l, g) = simpson_Eout(&c, Ebottom, Elo) + simpson_Eout(&c, Ehi, A1(E_bins, g + 1));
            }
            if ((Elo < alphaEin) && (alphaEin < Ehi)) {
                for (l = 1; l <= order; ++l) {
                    c.l = l - 1;
                    D(l, g) = D(l, g) + simpson_Eout(&c, Elo, alphaEin);
                }
                Elo = alphaEin;
            }
            if ((Elo < Ein) && (Ein < Ehi)) {
                for (l = 1; l <= order; ++l) {
                    c.l = l - 1;
                    D(l, g) = D(l, g) + simpson_Eout(&c, Elo, Ein);
                }
                Elo = Ein;
            }
            for (l = 1; l <= order; ++l) {
                c.l = l - 1;
                D(l, g) = D(l, g) + simpson_Eout(&c, Elo, Ehi);
            }
        } else {
            /* :118-131 -- Ebottom is computed but unused; the integral runs from E_bins(g) */
            for (l = 1; l <= order; ++l) {
                c.l = l - 1;
                D(l, g) = simpson_Eout(&c, A1(E_bins, g), A1(E_bins, g + 1));
            }
        }
        p0_1g_norm = p0_1g_norm + D(1, g);
        for (l = 1; l <= order; ++l)
            if (fabs(D(l, g)) < 1E-18) D(l, g) = ZERO;
    }
    for (g = 1; g <= groups; ++g)
        for (l = 1; l <= order; ++l) D(l, g) = D(l, g) / p0_1g_norm;
#undef D
}
