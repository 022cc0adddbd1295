"""integrate_freegas_leg and its helpers (src/freegas.F90:18-644) walked literally in pure Python, transcribed from the
Fortran text independently of oracle/freegas_ref.c: calc_FG_Eout_bounds, calc_sab (NJOY limits alpha_min / sab_min /
lterm_min), brent_mu, find_FG_mu, calc_fgk (-708 cut, no lterm floor), the two recursive adaptive Simpson integrators and
the group loop with its break points (alpha*E_in, E_in), tails and the final normalisation.  Python floats are IEEE
doubles without contraction and math.exp / math.sqrt are the C library's, so the walk reproduces the restatement bit
for bit (tests/test_oracle_golden.py runs both with loosened adaptive tolerances to keep the pure-Python cost small)."""
import math

import numpy as np

PI = 3.1415926535898
HUGE = 1.7976931348623157e308


def make_walk(A, kT, fEmu, gmu, L, mu_tol, mu_its, eo_tol, eo_its, sab_threshold, brent_thresh):
    dmu = gmu[1] - gmu[0]
    M = len(gmu)
    count = [0]

    def pn(l, x):
        if l == 0: return 1.0
        if l == 1: return x
        if l == 2: return 1.5 * x * x - 0.5
        if l == 3: return 2.5 * x * x * x - 1.5 * x
        raise ValueError

    def calc_sab(Ein, Eout, beta, mu):
        t = (A + 1.0) / A
        lterm = math.sqrt(Eout / Ein) / kT * (t * t)
        alpha = (Ein + Eout - 2.0 * mu * math.sqrt(Ein * Eout)) / (A * kT)
        if alpha < 1.0e-6: alpha = 1.0e-6
        s = -(alpha + beta) ** 2 / (4.0 * alpha)
        if s < -225.0: return 0.0
        s = lterm * math.exp(s) / (math.sqrt(4.0 * PI * alpha))
        return 0.0 if s < 2.0e-10 else s

    def brent_mu(Ein, Eout, beta, thresh, lo, hi):
        a, b, c, d = lo, hi, 0.0, HUGE
        fa = calc_sab(Ein, Eout, beta, a) - thresh
        fb = calc_sab(Ein, Eout, beta, b) - thresh
        if fa * fb >= 0.0:
            return a if fa < fb else b
        if abs(fa) < abs(fb):
            a, b, fa, fb = b, a, fb, fa
        c, fc, mflag = a, fa, True
        while fb != 0.0 and abs(a - b) > brent_thresh:
            if fa != fc and fb != fc:
                s = a * fb * fc / (fa - fb) / (fa - fc) + b * fa * fc / (fb - fa) / (fb - fc) + c * fa * fb / (fc - fa) / (fc - fb)
            else:
                s = b - fb * (b - a) / (fb - fa)
            tmp = (3.0 * a + b) * 0.25
            if (not ((s > tmp and s < b) or (s < tmp and s > b))) or (mflag and abs(s - b) >= 0.5 * abs(b - c)) or \
               ((not mflag) and abs(s - b) >= abs(c - d) * 0.5):
                s = 0.5 * (a + b); mflag = True
            else:
                if (mflag and abs(b - c) < brent_thresh) or ((not mflag) and abs(c - d) < brent_thresh):
                    s = (a + b) * 0.5; mflag = True
                else:
                    mflag = False
            fs = calc_sab(Ein, Eout, beta, s) - thresh
            d, c, fc = c, b, fb
            if fa * fs < 0.0:
                b, fb = s, fs
            else:
                a, fa = s, fs
            if abs(fa) < abs(fb):
                a, b, fa, fb = b, a, fb, fa
        return b

    def find_mu(Ein, Eout):
        beta = (Eout - Ein) / kT
        alpha_max = math.sqrt(beta * beta + 1.0) - 1.0
        den = 2.0 * math.sqrt(Ein * Eout)
        num = Ein + Eout - alpha_max * A * kT
        mu_max = num / den if den != 0.0 else math.copysign(math.inf, num)
        if abs(mu_max) > 1.0:
            return -1.0, 1.0
        sab_max = calc_sab(Ein, Eout, beta, mu_max)
        th = sab_max * sab_threshold
        lo = -1.0 if calc_sab(Ein, Eout, beta, -1.0) > th else brent_mu(Ein, Eout, beta, th, -1.0, mu_max)
        hi = 1.0 if calc_sab(Ein, Eout, beta, 1.0) > th else brent_mu(Ein, Eout, beta, th, mu_max, 1.0)
        return lo, hi

    def fgk(Ein, Eout, l, mu):
        count[0] += 1
        if mu <= gmu[0]: i = 1
        elif mu >= gmu[M - 1]: i = M - 1
        else: i = int((mu + 1.0) / dmu) + 1
        interp = (mu - gmu[i - 1]) / (gmu[i] - gmu[i - 1])
        fv = (1.0 - interp) * fEmu[i - 1] + interp * fEmu[i]
        t = (A + 1.0) / A
        lterm = fv * math.sqrt(Eout / Ein) / kT * (t * t)
        alpha = (Ein + Eout - 2.0 * mu * math.sqrt(Ein * Eout)) / (A * kT)
        beta = (Eout - Ein) / kT
        if alpha < 1.0e-6: alpha = 1.0e-6
        v = -(alpha + beta) ** 2 / (4.0 * alpha)
        if v <= -708.0: return 0.0
        return lterm * math.exp(v) / (math.sqrt(4.0 * PI * alpha)) * pn(l, mu)

    def aux_mu(Ein, Eout, l, a, b, eps, S, fa, fb, fc, bottom):
        c = 0.5 * (a + b); h = b - a; d = 0.5 * (a + c); e = 0.5 * (c + b)
        fd, fe = fgk(Ein, Eout, l, d), fgk(Ein, Eout, l, e)
        Sl = (h / 12.0) * (fa + 4.0 * fd + fc); Sr = (h / 12.0) * (fc + 4.0 * fe + fb); S2 = Sl + Sr
        if bottom <= 0 or abs(S2 - S) <= 15.0 * eps:
            return S2 + (S2 - S) / 15.0
        return aux_mu(Ein, Eout, l, a, c, 0.5 * eps, Sl, fa, fc, fd, bottom - 1) + \
            aux_mu(Ein, Eout, l, c, b, 0.5 * eps, Sr, fc, fb, fe, bottom - 1)

    def simp_mu(Ein, Eout, l, a, b):
        c = (a + b) * 0.5; h = b - a
        fa, fb, fc = fgk(Ein, Eout, l, a), fgk(Ein, Eout, l, b), fgk(Ein, Eout, l, c)
        S = (h / 6.0) * (fa + 4.0 * fc + fb)
        return aux_mu(Ein, Eout, l, a, b, mu_tol, S, fa, fb, fc, mu_its)

    def inner(Ein, Eout, l):
        lo, hi = find_mu(Ein, Eout)
        return simp_mu(Ein, Eout, l, lo, hi)

    def aux_eo(Ein, l, a, b, eps, S, fa, fb, fc, bottom):
        c = 0.5 * (a + b); d = 0.5 * (a + c); e = 0.5 * (c + b); h = b - a
        fd, fe = inner(Ein, d, l), inner(Ein, e, l)
        Sl = (h / 12.0) * (fa + 4.0 * fd + fc); Sr = (h / 12.0) * (fc + 4.0 * fe + fb); S2 = Sl + Sr
        if bottom <= 0 or abs(S2 - S) <= 15.0 * eps:
            return S2 + (S2 - S) / 15.0
        return aux_eo(Ein, l, a, c, 0.5 * eps, Sl, fa, fc, fd, bottom - 1) + \
            aux_eo(Ein, l, c, b, 0.5 * eps, Sr, fc, fb, fe, bottom - 1)

    def simp_eo(Ein, l, a, b):
        c = 0.5 * (a + b); h = b - a
        fa, fb, fc = inner(Ein, a, l), inner(Ein, b, l), inner(Ein, c, l)
        S = (h / 6.0) * (fa + 4.0 * fc + fb)
        return aux_eo(Ein, l, a, b, eo_tol, S, fa, fb, fc, eo_its)

    def integrate(Ein, E_bins):
        G = len(E_bins) - 1
        out = [[0.0] * L for _ in range(G)]
        alphaEin = (A - 1.0) / (A + 1.0)
        alphaEin = alphaEin * alphaEin * Ein
        ar = ((A - 1.0) / (A + 1.0)) ** 2
        Eout_lo = 0.001 * ar * Ein
        Eout_hi = 12.0 * kT * (A + 1.0) / A + (1.5 * Ein if Ein > 300.0 * kT / A else 2.0 * Ein)
        norm = 0.0
        for g in range(G):
            if E_bins[g] < Eout_hi and E_bins[g + 1] > Eout_lo:
                Elo = Eout_lo if Eout_lo > E_bins[g] else E_bins[g]
                Ehi = Eout_hi if Eout_hi < E_bins[g + 1] else E_bins[g + 1]
                Ebottom = 0.01 * Elo if E_bins[g] == 0.0 else E_bins[g]
                for l in range(L):
                    out[g][l] = simp_eo(Ein, l, Ebottom, Elo) + simp_eo(Ein, l, Ehi, E_bins[g + 1])
                if Elo < alphaEin < Ehi:
                    for l in range(L):
                        out[g][l] = out[g][l] + simp_eo(Ein, l, Elo, alphaEin)
                    Elo = alphaEin
                if Elo < Ein < Ehi:
                    for l in range(L):
                        out[g][l] = out[g][l] + simp_eo(Ein, l, Elo, Ein)
                    Elo = Ein
                for l in range(L):
                    out[g][l] = out[g][l] + simp_eo(Ein, l, Elo, Ehi)
            else:
                for l in range(L):
                    out[g][l] = simp_eo(Ein, l, E_bins[g], E_bins[g + 1])
            norm += out[g][0]
            for l in range(L):
                if abs(out[g][l]) < 1e-18: out[g][l] = 0.0
        return np.array(out) / norm, count[0]
    return integrate


