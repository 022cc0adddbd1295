// kernels_file4.cuh -- K2 and the E_in loops.
//
//   file4_cm_warp   <- integrate_file4_cm_leg + the lin-lin blend of integrate_distro
//                      (src/scattdata_header.F90:956-1078, 542-589), one warp per (E_in, slot)
//   k_elastic       <- calc_elastic_grid   (src/scatt.F90:603-675), one warp per E_in
//   k_inelastic     <- calc_inelastic_grid (src/scatt.F90:682-778), one block per E_in, warps over
//                      reactions, contributions summed in the reference's reaction order
//   k_copy_top      <- the "Ein above the top group edge copies the previous column" rule
//                      (src/scatt.F90:669,770)
//
// The reference calls integrate_file4_cm_leg twice per (E_in, reaction) -- once for the table row
// below E_in and once for the row above -- and re-evaluates tolab and every P_l at both ends of
// every mu segment.  Here both rows share one pass: each node of the integration grid is visited
// once, u = tolab(R, w) and P_0..P_{L-1}(u) are evaluated once, and the trapezoid is applied as a
// node weight (half the sum of the two adjacent segment widths).  That changes the order of the
// floating-point additions only (relative effect ~1e-16).
#pragma once
#include "common.cuh"

namespace ndpp {

// Moments of one (E_in, slot) pair.  dst[g*L + l] receives the blended, scaled moments of the
// groups the reaction reaches; other entries are left untouched (the caller pre-zeroes).
// All 32 lanes of the warp must call this together.
__device__ __forceinline__ void file4_cm_warp(const NucDev& nuc, const SlotDev& s, const InterpInfo& info, double Ein,
                                              double* __restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const int M = nuc.M, L = nuc.L, G = nuc.G;
    const double* __restrict__ w = nuc.mu;
    const int iE = info.iE;
    const double* __restrict__ f_lo = s.tab + (size_t)s.row_off[iE] * M;
    const double* __restrict__ f_hi = s.tab + (size_t)s.row_off[iE + 1] * M;
    const double fE = (Ein - s.e_grid[iE]) / (s.e_grid[iE + 1] - s.e_grid[iE]);

    const double awr = nuc.awr, Q = s.Q;
    const double dw = w[1] - w[0];
    const double R = awr * sqrt((1.0 + Q * (awr + 1.0) / (awr * Ein)));
    const double onepawr2 = (1.0 + awr) * (1.0 + awr);
    const double onepR2 = 1.0 + R * R;
    const double inv2REin = 0.5 / (R * Ein);

    for (int g = 0; g < G; ++g) {
        double wlo = (nuc.e_bins[g] * onepawr2 - Ein * onepR2) * inv2REin;
        if (wlo < -1.0) wlo = -1.0; else if (wlo > 1.0) wlo = 1.0;
        const int ilo = (int)((wlo + 1.0) / dw) + 1;  // 1-based, as the reference
        double whi = (nuc.e_bins[g + 1] * onepawr2 - Ein * onepR2) * inv2REin;
        if (whi < -1.0) whi = -1.0; else if (whi > 1.0) whi = 1.0;
        const int ihi = (int)((whi + 1.0) / dw) + 1;
        if (wlo == whi) {
            if (wlo == -1.0) continue;
            else if (wlo == 1.0) break;
        }
        // end-point values of both rows (:1021-1034)
        double flo_a, flo_b, fhi_a, fhi_b;
        if (ilo == M) {
            flo_a = f_lo[M - 1]; flo_b = f_hi[M - 1];
        } else {
            const double t = (wlo - w[ilo - 1]) / (w[ilo] - w[ilo - 1]);
            flo_a = (1.0 - t) * f_lo[ilo - 1] + t * f_lo[ilo];
            flo_b = (1.0 - t) * f_hi[ilo - 1] + t * f_hi[ilo];
        }
        if (ihi == M) {
            fhi_a = f_lo[M - 1]; fhi_b = f_hi[M - 1];
        } else {
            const double t = (whi - w[ihi - 1]) / (w[ihi] - w[ihi - 1]);
            fhi_a = (1.0 - t) * f_lo[ihi - 1] + t * f_lo[ihi];
            fhi_b = (1.0 - t) * f_hi[ihi - 1] + t * f_hi[ihi];
        }
        // node list: x_0 = wlo, x_j = w(ilo + j) (1-based) for j = 1..ihi-ilo, x_last = whi
        const int n_int = ihi - ilo;  // interior nodes
        const int n_nodes = n_int + 2;
        double acc_a[NDPP_MAX_L], acc_b[NDPP_MAX_L];
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) { acc_a[l] = 0.0; acc_b[l] = 0.0; }
        for (int j = lane; j < n_nodes; j += 32) {
            double x, xm, xp, fa, fb;
            if (j == 0) { x = wlo; fa = flo_a; fb = flo_b; }
            else if (j == n_nodes - 1) { x = whi; fa = fhi_a; fb = fhi_b; }
            else { x = w[ilo - 1 + j]; fa = f_lo[ilo - 1 + j]; fb = f_hi[ilo - 1 + j]; }
            // neighbours (segment widths are formed exactly as the reference forms them)
            double wl = 0.0, wr = 0.0;
            if (j > 0) {
                xm = (j == 1) ? wlo : w[ilo - 2 + j];
                wl = x - xm;
            }
            if (j < n_nodes - 1) {
                xp = (j == n_nodes - 2) ? whi : w[ilo + j];
                wr = xp - x;
            }
            const double wgt = wl + wr;
            const double u = tolab(R, x);
            double pn[NDPP_MAX_L];
            calc_pn_all(L, u, pn);
            const double ga = wgt * fa, gb = wgt * fb;
#pragma unroll
            for (int l = 0; l < NDPP_MAX_L; ++l)
                if (l < L) { acc_a[l] += ga * pn[l]; acc_b[l] += gb * pn[l]; }
        }
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) {
            if (l < L) {
                const double a = 0.5 * warp_sum(acc_a[l]);
                const double b = 0.5 * warp_sum(acc_b[l]);
                if (lane == 0) {
                    const double v = a * (1.0 - fE) + b * fE;
                    dst[g * L + l] = apply_scale(info, v);
                }
            }
        }
    }
}

// calc_elastic_grid.  One warp per E_in.  Columns whose E_in is below the free-gas cutoff are
// produced by the free-gas kernels (kernels_freegas.cuh) and skipped here; columns above the top
// group edge are filled by k_copy_top afterwards.
// Register budgets.  Left to itself ptxas gave k_inelastic 146 registers: one block of 8 warps per SM, 12.5 % occupancy
// (ncu: FP64 pipe 21 %, issue slots 32 %) for a latency-bound kernel.  C2, ms of everything but the file-6 pipeline per
// step, same box: 1 block/SM (k_elastic 3) 11.3; 2 (4) at 128 registers 7.8; 3 (5) at 80 / 96 registers 8.2; 4 (6) 9.6.
#ifndef K4_EL_BLOCKS
#define K4_EL_BLOCKS 4
#endif
#ifndef K4_INEL_BLOCKS
#define K4_INEL_BLOCKS 2
#endif
__global__ void __launch_bounds__(128, K4_EL_BLOCKS) k_elastic(NucDev nuc, const SlotDev* __restrict__ slots, const int* __restrict__ el_ids, int n_el,
                          const double* __restrict__ Ein, int NE, double* __restrict__ out)
{
    const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= NE) return;
    const double E = Ein[warp];
    const int GL = nuc.G * nuc.L;
    double* col = out + (size_t)warp * GL;
    if (!(E <= nuc.e_bins[nuc.n_bins - 1])) return;  // k_copy_top
    for (int k = 0; k < n_el; ++k) {
        const SlotDev& s = slots[el_ids[k]];
        // zero first: a free-gas column whose interpolated elastic xs is <= 0 is written by nobody else
        // (scatt_interp_distro returns distro = ZERO there, scattdata_header.F90:414-419)
        for (int e = lane; e < GL; e += 32) col[e] = 0.0;
        __syncwarp();
        if (E < nuc.freegas_cutoff) continue;        // free-gas column: k_freegas_finish overwrites the active ones
        const InterpInfo info = interp_info(nuc, s, E);
        if (info.active) file4_cm_warp(nuc, s, info, E, col);
        __syncwarp();
    }
    if (n_el == 0)
        for (int e = lane; e < GL; e += 32) col[e] = 0.0;
}

// calc_inelastic_grid.  One block per E_in.  Warps take the non-elastic slots round-robin; each
// warp builds its slot's [G][L] contribution in its own shared-memory slab; after every round the
// slabs are added to the block accumulator in slot order, which reproduces the reference's
// summation order over reactions.  File-6 slots were integrated beforehand by their own kernels
// (kernels_file6.cuh); their per-E_in results (already scaled) are read from pre[slot].
// Dynamic shared memory: (2 + nwarps) * G * L doubles + nwarps doubles.
__global__ void __launch_bounds__(256, K4_INEL_BLOCKS) k_inelastic(NucDev nuc, const SlotDev* __restrict__ slots, const int* __restrict__ in_ids, int n_in,
                            const double* const* __restrict__ pre, const double* __restrict__ Ein, int NE,
                            double* __restrict__ out, double* __restrict__ nuout)
{
    extern __shared__ double sm[];
    const int iE = blockIdx.x;
    const int GL = nuc.G * nuc.L;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* acc = sm;
    double* nuacc = sm + GL;
    double* slabs = sm + 2 * GL;
    double* yields = slabs + (size_t)nw * GL;
    const double E = Ein[iE];
    if (!(E <= nuc.e_bins[nuc.n_bins - 1])) return;  // k_copy_top
    for (int e = threadIdx.x; e < GL; e += blockDim.x) { acc[e] = 0.0; nuacc[e] = 0.0; }
    double* slab = slabs + (size_t)warp * GL;
    for (int base = 0; base < n_in; base += nw) {
        const int k = base + warp;
        for (int e = lane; e < GL; e += 32) slab[e] = 0.0;
        __syncwarp();
        if (k < n_in) {
            const int sid = in_ids[k];
            const SlotDev& s = slots[sid];
            if (pre[sid] != nullptr) {
                const double* src = pre[sid] + (size_t)iE * GL;
                for (int e = lane; e < GL; e += 32) slab[e] = src[e];
            } else {
                const InterpInfo info = interp_info(nuc, s, E);
                if (info.active) file4_cm_warp(nuc, s, info, E, slab);
            }
            if (lane == 0) {
                double y = (double)s.multiplicity;
                if (nuout != nullptr && s.yield != nullptr) y = interpolate_tab1(s.yield, E);
                yields[warp] = y;
            }
        }
        __syncthreads();
        const int nk = min(nw, n_in - base);
        for (int e = threadIdx.x; e < GL; e += blockDim.x) {
            double a = acc[e], b = nuacc[e];
            for (int q = 0; q < nk; ++q) {
                const double v = slabs[(size_t)q * GL + e];
                a = a + v;
                b = b + yields[q] * v;
            }
            acc[e] = a; nuacc[e] = b;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < GL; e += blockDim.x) {
        out[(size_t)iE * GL + e] = acc[e];
        if (nuout != nullptr) nuout[(size_t)iE * GL + e] = nuacc[e];
    }
}

// Columns whose E_in lies above the top group edge copy the previous column (src/scatt.F90:669,770).
// Single block; walks E_in in order so that chains of such points resolve as the serial loop does.
__global__ void k_copy_top(const double* __restrict__ Ein, int NE, double e_top, int GL, double* __restrict__ a,
                           double* __restrict__ b)
{
    for (int base = 0; base < NE; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int hit = (i < NE) && !(Ein[i] <= e_top);
        if (!__syncthreads_or(hit)) continue;
        for (int j = base; j < min(NE, base + (int)blockDim.x); ++j) {
            if (!(Ein[j] <= e_top) && j > 0) {
                for (int e = threadIdx.x; e < GL; e += blockDim.x) {
                    a[(size_t)j * GL + e] = a[(size_t)(j - 1) * GL + e];
                    if (b) b[(size_t)j * GL + e] = b[(size_t)(j - 1) * GL + e];
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace ndpp
