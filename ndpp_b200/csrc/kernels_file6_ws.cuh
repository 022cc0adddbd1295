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
// Consumers take (E_in, group) tasks from a global counter (persistent grid) and post them to their
// producers, so there is no block-level barrier and no tail imbalance inside a block.
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
    double d[F6_CONS][NDPP_MAX_L];  // trapezoid accumulators of the consumers (warp-uniform)
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

// tasks t = b * G + g, b = 0 .. na-1 over the batch's active E_in act[a0 + b]
template <int LT>
__global__ void __launch_bounds__(F6_THREADS, F6_BLOCKS_PER_SM)
k_file6_cm_ws(NucDev nuc, const double* __restrict__ Ein, UbDev ub, const UbRec* __restrict__ rec,
              const int* __restrict__ sorted, const double* __restrict__ femu, const double* __restrict__ rmu,
              const int* __restrict__ act, int a0, int na, unsigned long long* __restrict__ counter,
              double* __restrict__ raw)
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

    const long long n_tasks = (long long)na * G;
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
        for (;;) {
            mbar_wait(tfull0 + 8u * tslot, tphase);
            const long long t = sh.task[cons][tslot];
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * tslot);
            F6_TASK_ADVANCE();
            if (t < 0) break;

            F6Pair P;
            f6_pair_setup(nuc, Ein, ub, act, a0, t, P);
            if (!P.active) continue;
            C.R = rec + (size_t)P.b * ub.maxU;
            C.fE = femu + (size_t)P.b * ub.maxU * M;
            C.NPu = P.NPu;
            C.use_guess = sorted[P.b] != 0;
            C.eo0 = __ldg(&C.R[0].eo); C.eoLast = __ldg(&C.R[P.NPu - 1].eo);

            double Eo = P.Eb_lo - P.dEo;
            int w0 = 0;  // windows of the task before this outgoing energy, modulo F6_NPROD
            for (int it = 0; it < K; ++it) {
                Eo = Eo + P.dEo;  // accumulated as the reference accumulates it (:1169-1173)
                F6Item I;
                f6_item_setup(P, Eo, M, I);
                if (I.skip) continue;
                // this producer's windows: j with (w0 + j) % F6_NPROD == q
                int j = q - w0;
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
        for (;;) {
            long long t = 0;
            if (lane == 0) t = (long long)atomicAdd(counter, 1ULL);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_tasks) t = -1;
            // hand the task to the producers
            mbar_wait(tempty0 + 8u * tslot, tphase ^ 1u);
            if (lane == 0) {
                sh.task[cons][tslot] = t;
                mbar_arrive(tfull0 + 8u * tslot);
            }
            F6_TASK_ADVANCE();
            if (t < 0) break;

            F6Pair P;
            f6_pair_setup(nuc, Ein, ub, act, a0, t, P);
            if (!P.active) continue;
            double* const d = sh.d[cons];
            if (lane < NDPP_MAX_L) d[lane] = 0.0;
            __syncwarp();
            double Eo = P.Eb_lo - P.dEo;
            int wq = 0;
            for (int it = 0; it < K; ++it) {
                Eo = Eo + P.dEo;
                F6Item I;
                f6_item_setup(P, Eo, M, I);
                double fEl[NDPP_MAX_L];
#pragma unroll
                for (int l = 0; l < NDPP_MAX_L; ++l) fEl[l] = 0.0;
                if (!I.skip) {
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
                        if (lane < 31 && p + 1 < M && (fv != 0.0 || fnext != 0.0)) {
                            const double x = I.mu_l_min + I.dmu * pd;
                            const double xh = I.mu_l_min + I.dmu * (pd + 1.0);
                            Powers A, B;
                            make_powers(x, A);
                            make_powers(xh, B);
                            add_int_pn_tablelin<LT>(L, x, xh, fv, fnext, A, B, fEl);
                        }
                    }
                }
                // trapezoid over the outgoing energies, in the reference's order (:1246-1252)
#pragma unroll
                for (int l = 0; l < NDPP_MAX_L; ++l) {
                    if (l < L) {
                        const double v = warp_sum(fEl[l]);
                        if (lane == 0) d[l] = (it != 0 && it != K - 1) ? d[l] + 2.0 * v : d[l] + v;
                    }
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int l = 0; l < NDPP_MAX_L; ++l)
                    if (l < L) raw[((size_t)P.iEin * G + P.g) * L + l] = d[l] * P.dEo * 0.5;
            }
        }
    }
#undef F6_TASK_ADVANCE
}

// The same work with one role per warp (no pipeline): every warp takes (E_in, group) tasks itself and
// alternates point evaluation and segment integrals.  Kept for the A/B measurement in DESIGN.md.
template <int LT>
__global__ void __launch_bounds__(128, 4)
k_file6_cm_solo(NucDev nuc, const double* __restrict__ Ein, UbDev ub, const UbRec* __restrict__ rec,
                const int* __restrict__ sorted, const double* __restrict__ femu, const double* __restrict__ rmu,
                const int* __restrict__ act, int a0, int na, unsigned long long* __restrict__ counter,
                double* __restrict__ raw)
{
    __shared__ double dsh[4][NDPP_MAX_L];
    constexpr int L = LT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = nuc.M, K = nuc.ne_per_grp, G = nuc.G;
    const long long n_tasks = (long long)na * G;
    F6PointCtx C;
    C.mu = nuc.mu; C.rmu = rmu; C.M = M;
    C.div_dmu.set(nuc.mu[1] - nuc.mu[0]);
    double* const d = dsh[warp];
    for (;;) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(counter, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        F6Pair P;
        f6_pair_setup(nuc, Ein, ub, act, a0, t, P);
        if (!P.active) continue;
        C.R = rec + (size_t)P.b * ub.maxU;
        C.fE = femu + (size_t)P.b * ub.maxU * M;
        C.NPu = P.NPu;
        C.use_guess = sorted[P.b] != 0;
        C.eo0 = __ldg(&C.R[0].eo); C.eoLast = __ldg(&C.R[P.NPu - 1].eo);
        if (lane < NDPP_MAX_L) d[lane] = 0.0;
        __syncwarp();
        double Eo = P.Eb_lo - P.dEo;
        for (int it = 0; it < K; ++it) {
            Eo = Eo + P.dEo;
            F6Item I;
            f6_item_setup(P, Eo, M, I);
            double fEl[NDPP_MAX_L];
#pragma unroll
            for (int l = 0; l < NDPP_MAX_L; ++l) fEl[l] = 0.0;
            if (!I.skip) {
                int guess = -1;
                for (int base = 0; base < M - 1; base += 31) {
                    const int p = base + lane;
                    const double x = I.mu_l_min + I.dmu * (double)p;
                    double fv = 0.0;
                    if (p < M) fv = f6_point(C, Eo, I.c, x, guess);
                    const double fnext = __shfl_down_sync(0xffffffffu, fv, 1);
                    if (lane < 31 && p + 1 < M && (fv != 0.0 || fnext != 0.0)) {
                        const double xh = I.mu_l_min + I.dmu * (double)(p + 1);
                        Powers A, B;
                        make_powers(x, A);
                        make_powers(xh, B);
                        add_int_pn_tablelin<LT>(L, x, xh, fv, fnext, A, B, fEl);
                    }
                }
            }
#pragma unroll
            for (int l = 0; l < NDPP_MAX_L; ++l) {
                if (l < L) {
                    const double v = warp_sum(fEl[l]);
                    if (lane == 0) d[l] = (it != 0 && it != K - 1) ? d[l] + 2.0 * v : d[l] + v;
                }
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int l = 0; l < NDPP_MAX_L; ++l)
                if (l < L) raw[((size_t)P.iEin * G + P.g) * L + l] = d[l] * P.dEo * 0.5;
        }
        __syncwarp();
    }
}

}  // namespace ndpp
