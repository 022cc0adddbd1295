// kernels_file6_ws.cuh -- K3, warp-specialised: integrate_file6_cm_leg (src/scattdata_header.F90:1085-1253)
// as a persistent producer / consumer pipeline.
//
// ncu on the one-role kernel (kernels_file6.cuh: k_file6_cm) showed warps spending 61 % of their time
// in the latency-bound point evaluation (E_out search, divisions, a square root, dependent table loads)
// and only 25 % in the closed-form Legendre segment integrals that carry 3/4 of the FP64 work, so the
// FP64 pipe idled at 63 %.  Here every (E_in, outgoing group) pair is worked by a consumer warp and
// F6_NPROD producer warps on the same SM sub-partition:
//
//   producer  evaluates f(mu) = proby * J * pEo at the M lab cosines of every outgoing energy of the
//             group (:1176-1236), 32 points per stage (31 segments + the shared end point), and hands the
//             values over through a ring of shared-memory stages guarded by mbarriers;
//   consumer  integrates the segments with calc_int_pn_tablelin (:1240-1244), reduces over the warp,
//             and accumulates the trapezoid over the NE_PER_GRP outgoing energies in the reference's
//             order (:1246-1252).  A single warp of segment integrals fills 89 % of the FP64 pipe
//             (scratch/cbench.cu), so the pipe stays busy while the producers wait on memory.
//
// Consumers take (E_in, group, outgoing energy) tasks from a global counter (persistent grid) and post them to
// their producers, so there is no block-level barrier and no tail imbalance inside a block.  A task is one of the
// NE_PER_GRP outgoing energies of an (E_in, group) pair (65 windows, ~0.1 ms of a warp pair): with whole pairs as
// tasks (20 x larger) the last wave of a launch left the device a third empty on a nuclide sharded over 8 GPUs
// (16 800 pairs for 1 776 consumer warps).  The moments of a task go to `part`; k_file6_reduce adds them over the
// outgoing energies in the reference's order with the trapezoid weights (:1246-1252), so the bits do not change.
//
// The arithmetic is the one of k_file6_cm, operation for operation (results are bit-identical; the
// GPU tests compare the two kernels).  What differs is bookkeeping:
//   * fEmu(mu, E_out) of the unit-base interpolation (:1679-1702) is materialised once per active E_in
//     (k_f6_femu, as the reference does) instead of being re-blended from four table values at every
//     lookup: 36 FP64 operations and 12 loads less per point, for 16 KB per union-grid point that stay
//     in L2 while the E_in is being worked on;
//   * the union grid is read as 32-byte records (UbRec) through L1;
//   * the E_out interval of a point is first looked for next to the interval of the lane's previous
//     point (E_out,cm falls monotonically with the lab cosine); on a sorted grid the interval
//     a[i] <= v < a[i+1] is unique, so a verified guess is the bisection's answer; anything else runs
//     the reference's bisection;
//   * divisions whose divisor is a property of the grid (E_out interval width, mu spacing) use a
//     reciprocal refined once per interval (FastDiv: nvcc's own division sequence with the Newton
//     refinement hoisted; falls back to `/` outside the range nvcc itself guards).
#pragma once
#include <stdint.h>

#include "kernels_file6.cuh"

namespace ndpp {

struct alignas(16) UbRec {  // interval i of the union grid of one E_in
    double eo;    // E_out(i)
    double pd;    // pdf(i); pdf(NPu-2) is zero when the last point is duplicated (:1127-1130)
    double reo;   // refined reciprocal of E_out(i+1) - E_out(i) (unused when the interval is empty)
    double pad;
};
static_assert(sizeof(UbRec) == 32, "UbRec must be 32 bytes");

// rmu[k] = refined reciprocal of mu[k+1] - mu[k], once per nuclide
__global__ void k_rmu(const double* __restrict__ mu, int M, double* __restrict__ rmu)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < M - 1) rmu[k] = FastDiv::refine(mu[k + 1] - mu[k]);
    else if (k == M - 1) rmu[k] = 0.0;
}

// Ordered compaction of the active E_in (ub.n > 0); one block.
__global__ void k_f6_active(const int* __restrict__ n, int NE, int* __restrict__ act, int* __restrict__ n_act)
{
    __shared__ int wtot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < NE; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool flag = (i < NE) && (n[i] > 0);
        const unsigned b = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) wtot[warp] = __popc(b);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += wtot[w];
        if (flag) act[off + __popc(b & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; ++w) t += wtot[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_act = base;
}

// Records of the active E_in a0 .. a0+na-1 (block a - a0) + the "grid is sorted" flag that licenses the
// guessed search.  rec / sorted are indexed by the position in this batch.
__global__ void k_f6_records(UbDev ub, const int* __restrict__ act, int a0, UbRec* __restrict__ rec,
                             int* __restrict__ sorted)
{
    __shared__ int bad;
    const int b = blockIdx.x, i = act[a0 + b];
    const int n = ub.n[i];
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    const size_t o = (size_t)i * ub.maxU, ro = (size_t)b * ub.maxU;
    const bool dup_last = (n >= 2) && (ub.eout[o + n - 1] == ub.eout[o + n - 2]);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        UbRec r;
        r.eo = ub.eout[o + k];
        r.pd = (dup_last && k == n - 2) ? 0.0 : ub.pdf[o + k];
        r.reo = 0.0; r.pad = 0.0;
        if (k + 1 < n) {
            const double nx = ub.eout[o + k + 1];
            if (nx != r.eo) r.reo = FastDiv::refine(nx - r.eo);
            if (!(r.eo <= nx) || !(r.eo >= 0.0)) bad = 1;
        } else if (!(r.eo >= 0.0)) bad = 1;
        rec[ro + k] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) sorted[b] = bad ? 0 : 1;
}

// fEmu(k, i) = (1 - f) * row1 + f * row2 on the union grid (:1679-1702) for the batch's E_in:
// femu[(b * maxU + i) * M + k].  grid (ceil(M / 256), maxU, na).
__global__ void k_f6_femu(NucDev nuc, SlotDev s, UbDev ub, const int* __restrict__ act, int a0,
                          double* __restrict__ femu)
{
    const int b = blockIdx.z, i = blockIdx.y, iEin = act[a0 + b];
    if (i >= ub.n[iEin]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = nuc.M;
    if (k >= M) return;
    const int iE = ub.info[iEin].iE;
    const size_t o = (size_t)iEin * ub.maxU + i;
    const double f = ub.f[iEin], r1 = ub.r1[o], r2 = ub.r2[o];
    const double* c1 = s.tab + ((size_t)s.row_off[iE] + ub.j1[o]) * M + k;
    const double* c2 = s.tab + ((size_t)s.row_off[iE + 1] + ub.j2[o]) * M + k;
    double v = (1.0 - f) * ((1.0 - r1) * c1[0] + r1 * c1[M]);
    v = v + f * ((1.0 - r2) * c2[0] + r2 * c2[M]);
    femu[((size_t)b * ub.maxU + i) * M + k] = v;
}

// ---- mbarrier wrappers (shared::cta) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "F6_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra F6_DONE;\n"
        " bra F6_WAIT;\n"
        "F6_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// Block = F6_NPROD producer warpgroups + 1 consumer warpgroup: consumer c (warp 4 * F6_NPROD + c) is fed
// by the producers c, c + 4, ... (same SM sub-partition), which take the windows round-robin.
#ifndef F6_NPROD
#define F6_NPROD 1
#endif
#ifndef F6_CHUNK
#define F6_CHUNK 4         // outgoing energies per task; C2 on one B200, kernel ms: 1 -> 162.3, 4 -> 159.1, 20 (whole pair) -> 160.1
                           // (same box); on an eighth of the nuclide (8 GPUs) whole pairs left a 6 % tail
#endif
#ifndef F6_SHFL_POWERS
#define F6_SHFL_POWERS 1   // a segment's xhigh powers come from the next lane by shuffle: 159.1 -> 156.4 ms (0: recomputed)
#endif
#ifndef F6_STAGE_BITS
#define F6_STAGE_BITS 3   // ring of 2^bits windows per producer; A/B on C2 at 3 blocks/SM (same box, +-0.3 ms): 4 stages 161.0 ms, 8: 160.3, 16: 161.8
#endif
#define F6_STAGES (1 << F6_STAGE_BITS)
#define F6_CONS 4
#define F6_THREADS ((F6_NPROD + 1) * 128)
// Measured on B200 (C2, k_file6_cm 203 ms): 1 producer per consumer, 2 blocks/SM -> 173-183 ms; 2 producers with
// setmaxnreg 56/128 -> 189 ms; 3 producers, 1 block/SM -> 238 ms.
// Occupancy and the register split (C2, same box, kernel ms): 2 blocks/SM at 128 registers 172.6;
// 3 blocks/SM at 80 registers 169.8; 3 blocks/SM with the producers handing registers to the consumers
// (setmaxnreg 64/96: 170.2, 56/104: 168.0, 48/112: 169.1, 40/120: 172.7); 4 blocks/SM 64/64: 181.7, 40/88: 174.2.
// With the fused closed forms (legendre_fused.inc), whose live ranges are shorter: 56/104: 160.2, 64/96: 161.3,
// 48/112: 164.8; 4 blocks/SM 40/88: 166.7, 48/80: 165.4.
#ifndef F6_BLOCKS_PER_SM
#define F6_BLOCKS_PER_SM (F6_NPROD == 1 ? 3 : 1)
#endif
// setmaxnreg budgets of the two warpgroups.  They must balance against the launch allocation, 65536 / (256 * 3) rounded
// down to a multiple of 8 = 80 registers: 128 * (80 - 56) == 128 * (104 - 80); an unbalanced pair deadlocks.
#if F6_NPROD == 1 && F6_BLOCKS_PER_SM == 3 && !defined(F6_REG_PROD) && !defined(F6_NO_SETMAXNREG)
#define F6_REG_PROD 56
#define F6_REG_CONS 104
#endif

struct F6Shared {
    double buf[F6_CONS][F6_NPROD][F6_STAGES][32];
    unsigned long long full[F6_CONS][F6_NPROD][F6_STAGES];
    unsigned long long empty[F6_CONS][F6_NPROD][F6_STAGES];
    unsigned long long task_full[F6_CONS][2];   // consumer -> its producers: next (E_in, group) task
    unsigned long long task_empty[F6_CONS][2];
    long long task[F6_CONS][2];
};

// Scalars of one (E_in, group) pair (:1138-1173), computed identically by both roles.
struct F6Pair {
    int b, iEin, g, NPu, active;  // b = position of the E_in in the batch
    double E, Eout_last, Eb_lo, dEo, ap1inv;
};

__device__ __forceinline__ void f6_pair_setup(const NucDev& nuc, const double* __restrict__ Ein, const UbDev& ub,
                                              const int* __restrict__ act, int a0, long long t, F6Pair& P)
{
    const int G = nuc.G, nbins = nuc.n_bins, K = nuc.ne_per_grp;
    P.b = (int)(t / G);
    P.iEin = act[a0 + P.b];
    P.g = (int)(t % G);
    P.NPu = ub.n[P.iEin];
    P.E = Ein[P.iEin];
    P.active = 0;
    const double awr = nuc.awr;
    P.Eout_last = ub.eout[(size_t)P.iEin * ub.maxU + P.NPu - 1];
    P.ap1inv = 1.0 / (awr + 1.0);
    const double Eo_lo = 1E-12;  // the reference overwrites Eo_lo (:1141)
    const double Eo_hi = P.Eout_last + (P.E + 2.0 * (awr + 1.0) * sqrt(P.E * P.Eout_last)) * P.ap1inv * P.ap1inv;
    int g_lo, g_hi;
    double top;
    if (Eo_lo <= nuc.e_bins[0]) g_lo = 0;
    else if (Eo_lo >= nuc.e_bins[nbins - 1]) return;
    else g_lo = binary_search(nuc.e_bins, nbins, Eo_lo);
    if (Eo_hi <= nuc.e_bins[0]) return;
    else if (Eo_hi >= nuc.e_bins[nbins - 1]) { g_hi = nbins - 2; top = nuc.e_bins[g_hi]; }  // :1159 quirk
    else { g_hi = binary_search(nuc.e_bins, nbins, Eo_hi); top = Eo_hi; }
    if (P.g < g_lo || P.g > g_hi) return;
    P.Eb_lo = (P.g == g_lo) ? Eo_lo : nuc.e_bins[P.g];
    const double Eb_hi = (P.g == g_hi) ? top : nuc.e_bins[P.g + 1];
    P.dEo = (Eb_hi - P.Eb_lo) / (double)(K - 1);
    P.active = 1;
}

// Scalars of one outgoing energy (:1176-1187).
struct F6Item {
    double c, mu_l_min, dmu;
    bool skip;
};

__device__ __forceinline__ void f6_item_setup(const F6Pair& P, double Eo, int M, F6Item& I)
{
    I.c = P.ap1inv * sqrt(P.E / Eo);
    I.mu_l_min = (1.0 + I.c * I.c - P.Eout_last / Eo) / (2.0 * I.c);
    I.skip = false;
    if (I.mu_l_min < -1.0) I.mu_l_min = -1.0;
    else if (fabs(I.mu_l_min - 1.0) < 1E-10) I.mu_l_min = 1.0;
    else if (I.mu_l_min > 1.0) I.skip = true;
    I.dmu = (1.0 - I.mu_l_min) / (double)(M - 1);
}

// One lab cosine of one outgoing energy (:1188-1236): f(mu) = proby * J * pEo.
struct F6PointCtx {
    const UbRec* __restrict__ R;      // records of the E_in
    const double* __restrict__ fE;    // fEmu of the E_in: fE[i * M + k]
    const double* __restrict__ mu; const double* __restrict__ rmu;
    double eo0, eoLast;
    int NPu, M;
    bool use_guess;
    FastDiv div_dmu;
};

__device__ __forceinline__ double f6_point(const F6PointCtx& C, double Eo, double c, double x, int& guess)
{
    const double Eo_cm = Eo * (1.0 + c * c - 2.0 * c * x);
    if (!(Eo_cm > 0.0)) return 0.0;
    const UbRec* __restrict__ R = C.R;
    int iEo;
    if (Eo_cm <= C.eo0) iEo = 0;
    else if (Eo_cm >= C.eoLast) iEo = C.NPu - 2;
    else {
        const long long v = __double_as_longlong(Eo_cm);
        bool hit = false;
        iEo = guess;
        if (iEo >= 0) {
            const long long a = __double_as_longlong(__ldg(&R[iEo].eo));
            const long long b = __double_as_longlong(__ldg(&R[iEo + 1].eo));
            hit = (a <= v) && (v < b);
            if (!hit && v < a && iEo > 0) {
                iEo = iEo - 1;
                hit = __double_as_longlong(__ldg(&R[iEo].eo)) <= v;
            }
        }
        if (!hit) {  // the reference's bisection (binary_search_nonneg)
            int Lo = 0, Hi = C.NPu - 1;
            while (Hi - Lo > 1) {
                const int mid = Lo + (Hi - Lo) / 2;
                if (v >= __double_as_longlong(__ldg(&R[mid].eo))) Lo = mid; else Hi = mid;
            }
            iEo = Lo;
        }
    }
    if (C.use_guess) guess = iEo;
    const double2 a01 = __ldg(reinterpret_cast<const double2*>(R + iEo));      // eo, pd
    const double2 b01 = __ldg(reinterpret_cast<const double2*>(R + iEo + 1));
    double fEo, pEo;
    // the reference's INTT after unit-base interpolation is always lin-lin (:1716)
    if (b01.x == a01.x) { fEo = 0.0; pEo = a01.y; }
    else {
        FastDiv dv;
        dv.set(b01.x - a01.x, __ldg(&R[iEo].reo));
        fEo = dv(Eo_cm - a01.x);
        pEo = (1.0 - fEo) * a01.y + fEo * b01.y;
    }
    const double J = sqrt(Eo / Eo_cm);
    double mu_c;
    if (x == -1.0) mu_c = -1.0;
    else if (x == 1.0) mu_c = 1.0;
    else { mu_c = (x - c) * J; if (fabs(mu_c) > 1.0) return 0.0; }
    const int M = C.M;
    int k0; double ff;
    if (fabs(mu_c - 1.0) < 1E-10) { k0 = M - 2; ff = 1.0; }
    else {
        k0 = (int)C.div_dmu(mu_c + 1.0);
        const double m0 = __ldg(C.mu + k0), m1 = __ldg(C.mu + k0 + 1);
        FastDiv dm;
        dm.set(m1 - m0, __ldg(C.rmu + k0));
        ff = dm(mu_c - m0);
    }
    const double* __restrict__ fa = C.fE + (size_t)iEo * M + k0;
    double proby = (1.0 - fEo) * ((1.0 - ff) * __ldg(fa) + ff * __ldg(fa + 1));
    proby = proby + fEo * ((1.0 - ff) * __ldg(fa + M) + ff * __ldg(fa + M + 1));
    return proby * J * pEo;
}

// Eo of outgoing energy `it`, accumulated as the reference accumulates it (:1169-1173)
__device__ __forceinline__ double f6_item_energy(const F6Pair& P, int it)
{
    double Eo = P.Eb_lo - P.dEo;
    for (int i = 0; i <= it; ++i) Eo = Eo + P.dEo;
    return Eo;
}

// tasks t = (b * G + g) * nCh + chunk, b = 0 .. na-1 over the batch's active E_in act[a0 + b]; a task covers F6_CHUNK
// consecutive outgoing energies; part[((b * G + g) * K + it) * L + l] receives the integral over mu of outgoing energy
// `it` (zero-filled by the caller: inactive and skipped energies write nothing)
template <int LT>
__global__ void __launch_bounds__(F6_THREADS, F6_BLOCKS_PER_SM)
k_file6_cm_ws(NucDev nuc, const double* __restrict__ Ein, UbDev ub, const UbRec* __restrict__ rec,
              const int* __restrict__ sorted, const double* __restrict__ femu, const double* __restrict__ rmu,
              const int* __restrict__ act, int a0, int na, unsigned long long* __restrict__ counter,
              double* __restrict__ part)
{
    __shared__ F6Shared sh;
    constexpr int L = LT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp < F6_NPROD * F6_CONS;
    const int cons = warp & (F6_CONS - 1);  // consumer this warp belongs to / is
    const int q = warp / F6_CONS;           // producers: which of the F6_NPROD
    const int M = nuc.M, K = nuc.ne_per_grp, G = nuc.G;

    if (threadIdx.x < F6_CONS * F6_NPROD * F6_STAGES) {
        const int cc = threadIdx.x / (F6_NPROD * F6_STAGES), r = threadIdx.x % (F6_NPROD * F6_STAGES);
        mbar_init(smem_u32(&sh.full[cc][r / F6_STAGES][r % F6_STAGES]), 1);
        mbar_init(smem_u32(&sh.empty[cc][r / F6_STAGES][r % F6_STAGES]), 1);
    }
    if (threadIdx.x < F6_CONS * 2) {
        mbar_init(smem_u32(&sh.task_full[threadIdx.x >> 1][threadIdx.x & 1]), 1);
        mbar_init(smem_u32(&sh.task_empty[threadIdx.x >> 1][threadIdx.x & 1]), F6_NPROD);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int nCh = (K + F6_CHUNK - 1) / F6_CHUNK;   // tasks per (E_in, group) pair
    const long long n_tasks = (long long)na * G * nCh;
    const int nW = (M - 1 + 30) / 31;  // windows per outgoing energy
    const uint32_t tfull0 = smem_u32(&sh.task_full[cons][0]), tempty0 = smem_u32(&sh.task_empty[cons][0]);
    int tslot = 0;
    uint32_t tphase = 0;
#define F6_TASK_ADVANCE() do { if (++tslot == 2) { tslot = 0; tphase ^= 1u; } } while (0)

    if (producer) {
#ifdef F6_REG_PROD
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F6_REG_PROD));
#endif
        double* const ring = &sh.buf[cons][q][0][0];
        const uint32_t full0 = smem_u32(&sh.full[cons][q][0]), empty0 = smem_u32(&sh.empty[cons][q][0]);
        int stage = 0;
        uint32_t phase = 0;
        F6PointCtx C;
        C.mu = nuc.mu; C.rmu = rmu; C.M = M;
        C.div_dmu.set(nuc.mu[1] - nuc.mu[0]);
        int cur_b = -1;
        for (;;) {
            mbar_wait(tfull0 + 8u * tslot, tphase);
            const long long t = sh.task[cons][tslot];
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * tslot);
            F6_TASK_ADVANCE();
            if (t < 0) break;

            F6Pair P;
            f6_pair_setup(nuc, Ein, ub, act, a0, t / nCh, P);
            if (!P.active) continue;
            const int it0 = (int)(t % nCh) * F6_CHUNK, it1 = min(K, it0 + F6_CHUNK);
            double Eo = f6_item_energy(P, it0 - 1);
            if (P.b != cur_b) {   // consecutive tasks mostly share their E_in
                cur_b = P.b;
                C.R = rec + (size_t)P.b * ub.maxU;
                C.fE = femu + (size_t)P.b * ub.maxU * M;
                C.NPu = P.NPu;
                C.use_guess = sorted[P.b] != 0;
                C.eo0 = __ldg(&C.R[0].eo); C.eoLast = __ldg(&C.R[P.NPu - 1].eo);
            }
            int w0 = 0;  // windows of the task before this outgoing energy, modulo F6_NPROD
            for (int it = it0; it < it1; ++it) {
                Eo = Eo + P.dEo;
                F6Item I;
                f6_item_setup(P, Eo, M, I);
                if (I.skip) continue;
                int j = q - w0;   // this producer's windows: j with (w0 + j) % F6_NPROD == q
                if (j < 0) j += F6_NPROD;
                w0 = (w0 + nW) % F6_NPROD;
                int guess = -1;
                for (; j < nW; j += F6_NPROD) {
                    const int p = j * 31 + lane;
                    double fv = 0.0;
                    if (p < M) fv = f6_point(C, Eo, I.c, I.mu_l_min + I.dmu * (double)p, guess);
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    ring[stage * 32 + lane] = fv;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8u * stage);
                    if (++stage == F6_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
#ifdef F6_REG_CONS
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F6_REG_CONS));
#endif
        // ring positions of the F6_NPROD producers, packed: F6_STAGE_BITS of stage + 1 bit of phase each
        uint32_t pos = 0;
        const uint32_t full00 = smem_u32(&sh.full[cons][0][0]), empty00 = smem_u32(&sh.empty[cons][0][0]);
        const uint32_t buf00 = smem_u32(&sh.buf[cons][0][0][0]);
        // Tasks are taken one ahead: the next task is fetched and posted to the producers' two-slot mailbox before the
        // current one is consumed, so the producers run on into the next task's windows while the consumer finishes
        // this one (no pipeline bubble at a task boundary: the counter's round trip and the producers' set-up are hidden).
        auto fetch = [&]() -> long long {
            long long v = 0;
            if (lane == 0) v = (long long)atomicAdd(counter, 1ULL);
            v = __shfl_sync(0xffffffffu, v, 0);
            return v >= n_tasks ? -1 : v;
        };
        auto post = [&](long long v) {
            mbar_wait(tempty0 + 8u * tslot, tphase ^ 1u);
            if (lane == 0) {
                sh.task[cons][tslot] = v;
                mbar_arrive(tfull0 + 8u * tslot);
            }
            F6_TASK_ADVANCE();
        };
        long long t = fetch();
        post(t);
        for (long long t_next = 0; t >= 0; t = t_next) {
            t_next = fetch();
            post(t_next);

            F6Pair P;
            f6_pair_setup(nuc, Ein, ub, act, a0, t / nCh, P);
            if (!P.active) continue;
            const int it0 = (int)(t % nCh) * F6_CHUNK, it1 = min(K, it0 + F6_CHUNK);
            double Eo = f6_item_energy(P, it0 - 1);
            int wq = 0;
            for (int it = it0; it < it1; ++it) {
                Eo = Eo + P.dEo;
                F6Item I;
                f6_item_setup(P, Eo, M, I);
                if (I.skip) continue;
                double fEl[NDPP_MAX_L];
#pragma unroll
                for (int l = 0; l < NDPP_MAX_L; ++l) fEl[l] = 0.0;
                {
                    double pd = (double)lane;   // (double)(base + lane), advanced by the exact 31.0 per window
                    for (int base = 0; base < M - 1; base += 31, pd += 31.0) {
                        const int p = base + lane;
                        constexpr uint32_t PB = F6_STAGE_BITS + 1, PM = (1u << PB) - 1u;
                        const uint32_t sp = (pos >> (PB * wq)) & PM, st = sp & (F6_STAGES - 1u), ph = sp >> F6_STAGE_BITS;
                        const uint32_t slot = (uint32_t)wq * F6_STAGES + st;
                        mbar_wait(full00 + 8u * slot, ph);
                        double fv, fnext;
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fv) : "r"(buf00 + 256u * slot + 8u * lane));
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fnext) : "r"(buf00 + 256u * slot + 8u * ((lane + 1) & 31)));
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty00 + 8u * slot);
                        pos = (pos & ~(PM << (PB * wq))) | (((sp + 1u) & PM) << (PB * wq));  // stage++, phase flips on wrap
                        if (++wq == F6_NPROD) wq = 0;
                        // a segment whose two end values are zero adds exact zeros: skipped
                        const bool mine = lane < 31 && p + 1 < M && (fv != 0.0 || fnext != 0.0);
#if F6_SHFL_POWERS
                        // The right end of a lane's segment is the left end of the next lane's: x = mu_l_min + dmu * pd
                        // with pd an exact integer, so the powers of xhigh are bitwise the next lane's powers of xlow.
                        // Each point's powers are computed once and handed down by shuffles (off the FP64 pipe).
                        if (__any_sync(0xffffffffu, mine)) {
                            const double x = I.mu_l_min + I.dmu * pd;
                            Powers A, B;
                            make_powers(x, A);
                            const double xh = __shfl_down_sync(0xffffffffu, x, 1);
#define F6_DOWN(k) if constexpr (LT + 1 >= k) B.p##k = __shfl_down_sync(0xffffffffu, A.p##k, 1); else B.p##k = 0.0;
                            F6_DOWN(2) F6_DOWN(3) F6_DOWN(4) F6_DOWN(5) F6_DOWN(6) F6_DOWN(7) F6_DOWN(8) F6_DOWN(9) F6_DOWN(10)
                            F6_DOWN(11) F6_DOWN(12)
#undef F6_DOWN
                            if (mine) add_int_pn_tablelin<LT>(L, x, xh, fv, fnext, A, B, fEl);
                        }
#else
                        if (mine) {
                            const double x = I.mu_l_min + I.dmu * pd;
                            const double xh = I.mu_l_min + I.dmu * (pd + 1.0);
                            Powers A, B;
                            make_powers(x, A);
                            make_powers(xh, B);
                            add_int_pn_tablelin<LT>(L, x, xh, fv, fnext, A, B, fEl);
                        }
#endif
                    }
                }
                // the integral over mu of this outgoing energy; k_file6_reduce applies the trapezoid over the energies
#pragma unroll
                for (int l = 0; l < NDPP_MAX_L; ++l) {
                    if (l < L) {
                        const double v = warp_sum(fEl[l]);
                        if (lane == 0) part[((size_t)(t / nCh) * K + it) * L + l] = v;
                    }
                }
            }
        }
    }
#undef F6_TASK_ADVANCE
}

// Trapezoid over the NE_PER_GRP outgoing energies of every (E_in, group) pair of the batch, in the reference's order
// with its weights 1, 2, ..., 2, 1 and the final * dEo * 0.5 (:1246-1252).  One thread per (pair, order).
__global__ void k_file6_reduce(NucDev nuc, const double* __restrict__ Ein, UbDev ub, const int* __restrict__ act, int a0,
                               int na, int L, const double* __restrict__ part, double* __restrict__ raw)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int G = nuc.G, K = nuc.ne_per_grp;
    if (id >= (long long)na * G * L) return;
    const long long pair = id / L;
    const int l = (int)(id % L);
    F6Pair P;
    f6_pair_setup(nuc, Ein, ub, act, a0, pair, P);
    if (!P.active) return;   // raw was zero-filled
    const double* __restrict__ v = part + (size_t)pair * K * L + l;
    double d = 0.0;
    for (int it = 0; it < K; ++it) d = (it != 0 && it != K - 1) ? d + 2.0 * v[(size_t)it * L] : d + v[(size_t)it * L];
    raw[((size_t)P.iEin * G + P.g) * L + l] = d * P.dEo * 0.5;
}

}  // namespace ndpp
