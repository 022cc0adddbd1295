// Host build of ndpp_b200/csrc/libm_exact.cuh, compared bit for bit with the running C library
// (test infrastructure; compiled on the fly by tests/test_libm_exact.py with
//  g++ -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../ndpp_b200/csrc/libm_exact.cuh"

static inline uint64_t splitmix(uint64_t& s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static inline uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }

// fn: 0 exp, 1 expm1, 2 sinh, 3 cosh, 4 log.  Arguments: uniform in [lo, hi] (mode 0) or sign * 10^uniform(lo, hi) (mode 1).
// Returns the number of arguments whose results differ from libm's in any bit; *first_bad = one such argument.
extern "C" long long libm_exact_mismatches(int fn, double lo, double hi, long long n, unsigned long long seed, int mode,
                                           double* first_bad)
{
    long long bad = 0;
    double fb = 0.0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long blk = 0; blk < (n + 65535) / 65536; ++blk) {
        uint64_t s = seed * 0x2545f4914f6cdd1dull + (uint64_t)blk * 0x9e3779b97f4a7c15ull;
        const long long i1 = (blk + 1) * 65536 < n ? (blk + 1) * 65536 : n;
        for (long long i = blk * 65536; i < i1; ++i) {
            const uint64_t r = splitmix(s);
            const double u = (double)(r >> 11) * 0x1p-53;
            double x = lo + (hi - lo) * u;
            if (mode == 1) {
                x = pow(10.0, x);
                if (splitmix(s) & 1) x = -x;
            }
            double a, b;
            switch (fn) {
            case 0: a = ndpp::lm::exp_(x); b = exp(x); break;
            case 1: a = ndpp::lm::expm1_(x); b = expm1(x); break;
            case 2: a = ndpp::lm::sinh_(x); b = sinh(x); break;
            case 4: a = ndpp::lm::log_(x); b = log(x); break;
            default: a = ndpp::lm::cosh_(x); b = cosh(x); break;
            }
            if (bits(a) != bits(b) && !(a != a && b != b)) {
                ++bad;
#pragma omp critical
                fb = x;
            }
        }
    }
    if (first_bad) *first_bad = fb;
    return bad;
}

// element-wise evaluation with the running libm (the GPU test compares the device port with these)
extern "C" void libm_host_eval(int fn, const double* x, double* y, long long n)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) y[i] = fn == 4 ? log(x[i]) : fn == 2 ? sinh(x[i]) : fn == 3 ? cosh(x[i]) : fn == 1 ? expm1(x[i]) : exp(x[i]);
}
extern "C" void libm_port_eval(int fn, const double* x, double* y, long long n)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i)
        y[i] = fn == 4 ? ndpp::lm::log_(x[i]) : fn == 2 ? ndpp::lm::sinh_(x[i]) : fn == 3 ? ndpp::lm::cosh_(x[i]) : fn == 1 ? ndpp::lm::expm1_(x[i]) : ndpp::lm::exp_(x[i]);
}
