// kernels_egrid.cuh -- the incoming-energy grids on the device: create_Ein_grid
// (src/scatt.F90:166-236 with combine_Eins :246, add_elastic_Eins :311, add_one_more_point :426, add_inelastic_Eins
// :456) and sab_egrid (src/sab.F90:460-568).
//
// The reference builds a grid by a chain of two-pointer merges (src/array_merge.F90), one per reaction channel and --
// in add_inelastic_Eins -- one per (level, group edge): 2 760 merges of a growing array for a 40-level nuclide on 70
// groups, ~2 s of a host core (the oracle's literal restatement measures it).  A merge of two ascending arrays that
// drops the values found in both is the sorted union, and the union is associative, so the chain equals
//
//     sort(all candidate points) -> drop repeats
//
// whatever the order of the merges.  Here every candidate is produced by its own thread -- the points the reference
// places with log / exp of the host libm carry those bits (lm::log_, lm::exp_ of libm_exact.cuh) -- appended to one
// array through an atomic counter, sorted with a device-wide radix sort and compacted (cub, part of the CUDA toolkit:
// the sort is plumbing here, not one of the path's kernels); single-thread kernels apply the reference's searches for
// the cuts (iEthresh, E_bins(size)) and add_one_more_point.  The grids stay on the device for ndppgpu_group_set_grids /
// the *_dev entry points; nothing returns to the host but their lengths.
//
// Where this differs from the chain of merges, or meets one of its quirks (reported in the status word):
//   * a zero that meets a larger value becomes MIN_EIN in merge; every zero candidate is replaced before the sort when
//     some merged array does not start with zero (the chain would keep a second MIN_EIN if two arrays held a zero);
//   * values repeated inside one input array survive a merge in some positions; here every repeat is dropped (bit 2);
//   * a critical energy that is NaN (negative discriminant in add_inelastic_Eins, e.g. a positive Q far above a group
//     edge): its points are left out (bit 1).  The reference ends there too -- every comparison of its merge fails on
//     an all-NaN array, the whole other array is copied, and the one NaN its exit branch takes is cut off again
//     (src/array_merge.F90:78-100; the oracle's literal merge shows it) -- so this bit is information, not a difference.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"
#include "kernels_sab.cuh"
#include "libm_exact.cuh"

namespace ndpp {

#define EG_MIN_EIN 1E-14   // src/constants.F90:109
enum { EG_ST_RANGE = 1, EG_ST_NAN = 2, EG_ST_REPEAT = 4, EG_ST_OVERFLOW = 8 };

struct EgOut {
    double* cand;               // candidate points
    unsigned long long* count;  // appended so far
    long long cap;
    int* status;
};

__device__ __forceinline__ void eg_append(const EgOut& o, double v, int zero_to_min)
{
    if (v != v) { atomicOr(o.status, EG_ST_NAN); return; }
    if (zero_to_min && v == 0.0) v = EG_MIN_EIN;
    const unsigned long long p = atomicAdd(o.count, 1ULL);
    if (p < (unsigned long long)o.cap) o.cand[p] = v; else atomicOr(o.status, EG_ST_OVERFLOW);
}

// copies src[0 .. n) to the candidates; a repeat inside the array is flagged
__global__ void k_eg_copy(const double* __restrict__ src, int n, EgOut o, int zero_to_min)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = src[i];
    if (i + 1 < n && src[i + 1] == v) atomicOr(o.status, EG_ST_REPEAT);
    eg_append(o, v, zero_to_min);
}

// the same for several arrays in one launch: element t of the concatenation belongs to array k with
// first[k] <= t < first[k + 1] (the nuclide grid, the group structure and one grid per reaction channel: ~45 arrays)
struct EgSrc { const double* p; int first; int pad; };
__global__ void k_eg_copy_many(const EgSrc* __restrict__ srcs, int n_src, int total, EgOut o, int zero_to_min)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int lo = 0, hi = n_src;              // srcs[n_src].first == total
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (srcs[mid].first <= t) lo = mid; else hi = mid; }
    const int i = t - srcs[lo].first, n = srcs[lo + 1].first - srcs[lo].first;
    const double* __restrict__ src = srcs[lo].p;
    const double v = src[i];
    if (i + 1 < n && src[i + 1] == v) atomicOr(o.status, EG_ST_REPEAT);
    eg_append(o, v, zero_to_min);
}

// add_elastic_Eins (src/scatt.F90:311-419).  Thread t < (nb-1) * extend: the upscatter point i = t % extend - extend of
// group t / extend (:343-376); thread (nb-1) * extend + g: the downscatter points of group g, in order until the first
// one that is not below the group's top (:391-407).
__global__ void k_eg_elastic_pts(const double* __restrict__ E_bins, int nb, double awr, double kT, double cutoff,
                                 int extend, EgOut o)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_up = (nb - 1) * extend;
    if (t < n_up) {
        if (cutoff == 0.0) return;
        const int g = t / extend, i = t % extend - extend;
        const double lo_shift = 2.0 * kT * (awr + 1.0) / awr;
        double Ehi = E_bins[g + 1];
        const double Elo = E_bins[g];
        if (Ehi <= cutoff) {
            double dElo;
            if (Ehi - lo_shift > Elo) dElo = lm::log_(Ehi / (Ehi - lo_shift)) / (double)extend;
            else dElo = lm::log_(Ehi / 1E-11) / (double)extend;
            const double newE = Ehi * lm::exp_((double)i * dElo);
            if (newE >= Elo) eg_append(o, newE, 0);
        } else if (Elo < cutoff) {
            Ehi = cutoff;
            const double dElo = lm::log_(Ehi / (Ehi - lo_shift)) / (double)extend;
            const double newE = Ehi * lm::exp_((double)i * dElo);
            if (newE > Elo) eg_append(o, newE, 0);
        }
    } else if (t < n_up + nb - 1) {
        const int g = t - n_up;
        if (E_bins[g] == 0.0) return;
        double alpha = (awr - 1.0) / (awr + 1.0);
        alpha = alpha * alpha;
        const double dEhi = 7.0 * lm::log_(1.0 / alpha) / (double)extend;
        const double Ehi = E_bins[g + 1];
        for (int i = 1; i <= extend - 1; ++i) {
            const double newE = E_bins[g] * lm::exp_((double)i * dEhi);
            if (newE < Ehi) eg_append(o, newE, 0); else break;
        }
    }
}

// add_inelastic_Eins (src/scatt.F90:456-536): thread per (channel with Q /= 0, group edge g = 2 .. nb-1, point 1 .. pts-1);
// negQ[k] = -Q_value of the channel.
__global__ void k_eg_inelastic_pts(const double* __restrict__ negQ, int nq, const double* __restrict__ E_bins, int nb,
                                   double awr, double thresh, int pts, EgOut o)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = pts - 1, ng = nb - 2;
    if (t >= (long long)nq * ng * per) return;
    const int i = (int)(t % per) + 1, g = (int)((t / per) % ng) + 1, k = (int)(t / ((long long)per * ng));
    const double Q = negQ[k];
    const double Eg = E_bins[g];
    const double Ef = (1.0 + awr) / (awr) * Eg;
    const double D = ((awr * awr) * (1.0 + Ef / Q) - 1.0) * (Ef / Q);
    const double Fp = (1.0 + sqrt(D)) / (1.0 + Ef / Q);
    const double Fm = (1.0 - sqrt(D)) / (1.0 + Ef / Q);
    const double Ecp = ((1.0 + awr) / (awr) * Q) / (1.0 - Fp * Fp / (awr * awr));
    const double Ecm = ((1.0 + awr) / (awr) * Q) / (1.0 - Fm * Fm / (awr * awr));
    double Elo, Ehi;
    if (Ecp > Ecm) { Elo = Ecm; Ehi = Ecp; } else { Elo = Ecp; Ehi = Ecm; }
    if (Elo < thresh) Elo = thresh;
    if (Ehi < thresh) Ehi = thresh;
    if (Elo != Ehi) {
        const double dE = lm::log_(Ehi / Elo) / (double)pts;
        eg_append(o, Elo * lm::exp_((double)i * dE), 0);
    }
}

// binary_search of the reference (1-based lower index, src/search.F90:21-71) for the single-thread kernels below;
// 0 when val lies outside the array (the reference aborts)
__device__ inline int eg_search1(const double* a, int n, double val)
{
    if (n < 1 || val < a[0] || val > a[n - 1]) return 0;
    int L = 1, R = n;
    while (R - L > 1) {
        const int mid = L + (R - L) / 2;
        if (val >= a[mid - 1]) L = mid; else R = mid;
    }
    return L;
}

// res[0] = points of the finished grid, res[1] = auxiliary index.  One thread each.
//   mode 0  add_one_more_point on grid[0 .. *n_unique)                                    (elastic grid, :202-203)
//   mode 1  res[1] = iEthresh - 1, the 0-based start of Ein_el(iEthresh:) (:208-210)
//   mode 2  cut at binary_search(grid, E_bins(size)) and add_one_more_point (:223-235)
__global__ void k_eg_finish(double* __restrict__ grid, const int* __restrict__ n_unique, int n_given, int mode, double val,
                            int* __restrict__ res, int* __restrict__ status)
{
    if (blockIdx.x || threadIdx.x) return;
    int n = n_unique ? *n_unique : n_given;
    if (mode == 1) {
        const int i = eg_search1(grid, n, val);
        if (i == 0) { atomicOr(status, EG_ST_RANGE); res[1] = 0; } else res[1] = i - 1;
        return;
    }
    if (mode == 2) {
        const int i = eg_search1(grid, n, val);
        if (i == 0) atomicOr(status, EG_ST_RANGE); else n = i;
    }
    grid[n] = grid[n - 1] * (1.0 + (double)1.0E-3f);   // the literal 1.0E-3 is single precision (:438)
    res[0] = n + 1;
}

// ---- S(a,b) -----------------------------------------------------------------------------------------------------------
// crossing points (src/sab.F90:497-541): thread per (i, j) = (incoming interval, outgoing energy index)
__global__ void k_eg_sab_cross(SabDev sab, const double* __restrict__ e_bins, int nb, EgOut o)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)(sab.n_in - 1) * sab.n_eout) return;
    const int i = (int)(t / sab.n_eout), j = (int)(t % sab.n_eout);
    const double Ei1 = sab.e_in[i], Ei2 = sab.e_in[i + 1];
    const double Eo1 = sab.e_out[(size_t)i * sab.n_eout + j], Eo2 = sab.e_out[(size_t)(i + 1) * sab.n_eout + j];
    int g1 = eg_search1(e_bins, nb, Eo1), g2 = eg_search1(e_bins, nb, Eo2);
    if (g1 == 0 || g2 == 0) { atomicOr(o.status, EG_ST_RANGE); return; }
    if (Eo2 < Eo1) g2 = g1;
    for (int g = g1 + 1; g <= g2; ++g) eg_append(o, (e_bins[g - 1] - Eo1) / (Eo2 - Eo1) * (Ei2 - Ei1) + Ei1, 0);
}

// res[0] = i_max_ein = binary_search(Ein, max_ein) (:544)
__global__ void k_eg_sab_cut(const double* __restrict__ grid, const int* __restrict__ n_unique, double max_ein,
                             int* __restrict__ res, int* __restrict__ status)
{
    if (blockIdx.x || threadIdx.x) return;
    const int i = eg_search1(grid, *n_unique, max_ein);
    if (i == 0) atomicOr(status, EG_ST_RANGE);
    res[0] = i;
}

// EXTEND_PTS points inside every interval of base[0 .. i_max): Ein(j) = Ein(j-1) * exp(dE), one after the other
// (:552-566); thread per interval
__global__ void k_eg_sab_expand(const double* __restrict__ base, int i_max, int extend, double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i_max) return;
    if (k == i_max - 1) { out[(size_t)k * (extend + 1)] = base[k]; return; }
    const double dE = lm::log_(base[k + 1] / base[k]) / (double)(extend + 1);
    const double step = lm::exp_(dE);
    double v = base[k];
    double* const dst = out + (size_t)k * (extend + 1);
    dst[0] = v;
    for (int i = 1; i <= extend; ++i) { v = v * step; dst[i] = v; }
}

}  // namespace ndpp
