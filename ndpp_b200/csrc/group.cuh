// group.cuh -- several B200s behind the C-ABI: device groups, a nuclide sharded by E_in over a group, and the
// work-item planner / runner of a whole library.  Included at the end of ndppgpu.cu (same translation unit: it uses
// the Ctx / Nuclide internals).
//
// What it replaces: the reference distributes a library from its own driver -- partition_work hands every MPI rank a
// contiguous block of nuclides (src/ndpp.F90:934-950), the loop at :549 walks the block, and the results are handed
// back at :839-864.  Here
//   * a *group* is the set of GPUs working together: all GPUs of one process (ndppgpu_group_init: one host thread per
//     device, ncclCommInitAll) or one GPU per process (ndppgpu_group_init_rank: ncclCommInitRank with an id the ranks
//     exchanged over MPI / torch.distributed) -- the same code drives both;
//   * ndppgpu_group_nuclide_* mirror ndppgpu_nuclide_* one to one: the nuclide is replicated on every device of the
//     group and calc_elastic_grid / calc_inelastic_grid are sharded over the E_in grid.  Every (nuclide, E_in) column
//     depends on read-only tables only, so there is no exchange step; the one collective is the NCCL gather of the
//     finished columns to the root device (ncclSend / ncclRecv over NVLink on a side stream, double-buffered so that
//     it overlaps the next call's kernels), followed by the top-of-grid copy rule (src/scatt.F90:669,770) on the root;
//   * the E_in grid is dealt *cyclically* (column i to device i mod N): the cost of a column varies smoothly with
//     E_in (thresholds open one by one, the continuum's group count grows slowly), so neighbouring columns cost the
//     same and a cyclic deal balances to within one column per device without any cost model;
//   * a library (many nuclides) is cut into (nuclide, matrix, E_in tile) work items weighted with the algorithmic-flop
//     formulas of SURVEY 8d and dealt longest-processing-time-first with a set-up cost per (device, nuclide)
//     (ndppgpu_plan_library: pure host code); ndppgpu_library_run integrates a device's items in place into one result
//     buffer per device and gathers the buffers to the root with NCCL.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; the library is loaded on demand (a one-GPU run needs no NCCL)

#include <atomic>
#include <functional>
#include <map>
#include <set>
#include <thread>

namespace {

// ---- NCCL, loaded when the first multi-device group is created --------------------------------------------------
struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string load()
    {
        if (h) return "";
        const char* env = std::getenv("NDPPGPU_NCCL_LIB");
        // a copy that is already mapped (torch brings its own) is preferred: one NCCL per process
        for (const char* name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h)
            for (const char* name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
                h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
                if (h) break;
            }
        if (!h) return std::string("ndppgpu: cannot load NCCL (") + dlerror() + "); set NDPPGPU_NCCL_LIB";
#define NDPP_SYM(field, sym)                                                            \
    field = (decltype(field))dlsym(h, #sym);                                            \
    if (!field) return std::string("ndppgpu: NCCL lacks ") + #sym;
        NDPP_SYM(GetUniqueId, ncclGetUniqueId) NDPP_SYM(CommInitRank, ncclCommInitRank) NDPP_SYM(CommInitAll, ncclCommInitAll)
        NDPP_SYM(CommDestroy, ncclCommDestroy) NDPP_SYM(GroupStart, ncclGroupStart) NDPP_SYM(GroupEnd, ncclGroupEnd)
        NDPP_SYM(Send, ncclSend) NDPP_SYM(Recv, ncclRecv) NDPP_SYM(GetErrorString, ncclGetErrorString)
        NDPP_SYM(GetVersion, ncclGetVersion)
#undef NDPP_SYM
        return "";
    }
};
NcclApi g_nccl;

struct Group {
    int world = 1, n_local = 1, first = 0;   // devices in the group, devices of this process, global rank of local 0
    std::vector<Ctx*> ctx;                   // [n_local], owned
    std::vector<ncclComm_t> comm;            // [n_local], null when world == 1
    std::vector<cudaStream_t> cs;            // [n_local] collective (side) streams
    std::string err;
    long long gathered_bytes = 0;            // received by the root over NCCL
    bool is_root(int li) const { return first + li == 0; }
};

int gfail(Group* g, const std::string& msg)
{
    g_last_error = msg;
    if (g) {
        g->err = msg;
        if (!g->ctx.empty() && g->ctx[0]) g->ctx[0]->err = msg;
    }
    return 1;
}

#define NCK(g, call)                                                                                         \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != ncclSuccess)                                                                              \
            return gfail(g, std::string(#call) + ": " + g_nccl.GetErrorString(r__) + " (" __FILE__ ":" +     \
                                std::to_string(__LINE__) + ")");                                             \
    } while (0)

// f(local index) on one host thread per local device (the launches of one device synchronise with the host here
// and there -- count read-backs, the error latch -- so the devices need a thread each to run side by side).
// Returns 0 or the first failure, whose text becomes the group's.
template <class F> int group_parallel(Group* g, F&& f)
{
    const int n = g->n_local;
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    auto body = [&](int li) {
        cudaSetDevice(g->ctx[li]->device);
        rc[li] = f(li);
        if (rc[li]) msg[li] = g_last_error;   // thread-local text of the failing call
    };
    if (n == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (int li = 0; li < n; ++li) th.emplace_back(body, li);
        for (auto& t : th) t.join();
    }
    for (int li = 0; li < n; ++li)
        if (rc[li]) return gfail(g, msg[li].empty() ? "ndppgpu: device " + std::to_string(g->first + li) + " failed" : msg[li]);
    return 0;
}

int group_make_streams(Group* g)
{
    g->cs.assign(g->n_local, nullptr);
    for (int li = 0; li < g->n_local; ++li) {
        CK(g->ctx[li], cudaSetDevice(g->ctx[li]->device));
        CK(g->ctx[li], cudaStreamCreateWithFlags(&g->cs[li], cudaStreamNonBlocking));
    }
    return 0;
}

// out[i] = stage[i % world][i / world]: the cyclic deal undone.  One block per column.
__global__ void k_interleave(const double* __restrict__ stage, size_t slab, int world, int NE, int GL, double* __restrict__ out)
{
    for (int i = blockIdx.x; i < NE; i += gridDim.x) {
        const double* src = stage + (size_t)(i % world) * slab + (size_t)(i / world) * GL;
        double* dst = out + (size_t)i * GL;
        for (int e = threadIdx.x; e < GL; e += blockDim.x) dst[e] = src[e];
    }
}

// ---- a nuclide replicated over the group, its E_in grids dealt cyclically ----------------------------------------
struct GroupGrid {            // one E_in grid (elastic or inelastic) and its buffers
    int NE = 0, nmat = 1;     // nmat = 2: inelastic with nu-scatter
    std::vector<double> Ein;  // full grid (host)
    std::vector<DevBuf> d_E;                 // [n_local] this device's share
    std::vector<DevBuf> d_out[2];            // [parity][n_local] local result [n_loc][nmat? no: per matrix below]
    std::vector<DevBuf> d_nu[2];
    DevBuf stage[2], stage_nu[2];            // root: [world][max_n][GL]
    DevBuf fin[2], fin_nu[2];                // root: [NE][GL]
    DevBuf d_E_full;                         // root: the whole grid, for the top-of-grid rule
    int last = -1;                           // parity of the latest integrate
    int n_of(int rank, int world) const { return NE > rank ? (NE - rank + world - 1) / world : 0; }
};

struct GroupNuclide {
    Group* g = nullptr;
    std::vector<Nuclide*> nuc;               // [n_local]
    GroupGrid el, inel;
    std::vector<cudaEvent_t> ev_comp, ev_sent[2];   // [n_local]
    int calls = 0;
    bool nuscatter = false;
    int GL = 0;
    double e_top = 0.0;
};

int group_sync(GroupNuclide* gn);

int group_grid_set(GroupNuclide* gn, GroupGrid& gr, const double* Ein, int NE, int nmat)
{
    Group* g = gn->g;
    if (group_sync(gn)) return 1;   // a gather of the previous grid may still read the buffers released below
    gr.NE = std::max(NE, 0);
    gr.nmat = nmat;
    gr.last = -1;
    gr.Ein.assign(Ein ? Ein : nullptr, Ein ? Ein + gr.NE : nullptr);
    const int W = g->world;
    gr.d_E.clear(); gr.d_E.resize(g->n_local);
    for (int p = 0; p < 2; ++p) {
        gr.d_out[p].clear(); gr.d_out[p].resize(g->n_local);
        gr.d_nu[p].clear(); gr.d_nu[p].resize(g->n_local);
    }
    const size_t GL = (size_t)gn->GL;
    return group_parallel(g, [&](int li) -> int {
        Ctx* c = g->ctx[li];
        const int r = g->first + li, n_loc = gr.n_of(r, W);
        std::vector<double> share((size_t)n_loc);
        for (int k = 0; k < n_loc; ++k) share[k] = gr.Ein[(size_t)r + (size_t)k * W];
        if (upload(c, gr.d_E[li], share.data(), share.size())) return 1;
        for (int p = 0; p < 2; ++p) {
            if (dev_alloc(c, gr.d_out[p][li], (size_t)n_loc * GL * sizeof(double))) return 1;
            if (nmat == 2 && dev_alloc(c, gr.d_nu[p][li], (size_t)n_loc * GL * sizeof(double))) return 1;
        }
        if (g->is_root(li)) {
            const size_t slab = (size_t)gr.n_of(0, W) * GL;
            for (int p = 0; p < 2; ++p) {
                if (dev_alloc(c, gr.stage[p], (size_t)W * slab * sizeof(double)) ||
                    dev_alloc(c, gr.fin[p], (size_t)gr.NE * GL * sizeof(double)))
                    return 1;
                if (nmat == 2 && (dev_alloc(c, gr.stage_nu[p], (size_t)W * slab * sizeof(double)) ||
                                  dev_alloc(c, gr.fin_nu[p], (size_t)gr.NE * GL * sizeof(double))))
                    return 1;
            }
            if (upload(c, gr.d_E_full, gr.Ein.data(), gr.Ein.size())) return 1;
        }
        CK(c, cudaStreamSynchronize(c->stream));
        return 0;
    });
}

// compute this device's share into buffer `p`, then (side stream) send it to the root / receive and assemble
int group_integrate_one(GroupNuclide* gn, GroupGrid& gr, bool inelastic, int li, int p)
{
    Group* g = gn->g;
    Ctx* c = g->ctx[li];
    Nuclide* n = gn->nuc[li];
    const int W = g->world, r = g->first + li, n_loc = gr.n_of(r, W);
    const size_t GL = (size_t)gn->GL;
    double* out = gr.d_out[p][li].as<double>();
    double* nu = gr.nmat == 2 ? gr.d_nu[p][li].as<double>() : nullptr;
    // the buffer was last read by the send / assembly of two calls ago
    CK(c, cudaStreamWaitEvent(c->stream, gn->ev_sent[p][li], 0));
    if (n_loc > 0) {
        if (inelastic) { if (inelastic_dev(n, gr.d_E[li].as<double>(), n_loc, out, nu)) return 1; }
        else if (elastic_dev(n, gr.d_E[li].as<double>(), n_loc, out)) return 1;
    }
    CK(c, cudaEventRecord(gn->ev_comp[li], c->stream));
    cudaStream_t cs = g->cs[li];
    CK(c, cudaStreamWaitEvent(cs, gn->ev_comp[li], 0));
    if (!g->is_root(li)) {
        if (n_loc > 0) {
            NCK(g, g_nccl.GroupStart());
            NCK(g, g_nccl.Send(out, (size_t)n_loc * GL, ncclDouble, 0, g->comm[li], cs));
            if (nu) NCK(g, g_nccl.Send(nu, (size_t)n_loc * GL, ncclDouble, 0, g->comm[li], cs));
            NCK(g, g_nccl.GroupEnd());
        }
    } else {
        const size_t slab = (size_t)gr.n_of(0, W) * GL;
        double* st = gr.stage[p].as<double>();
        double* st_nu = gr.nmat == 2 ? gr.stage_nu[p].as<double>() : nullptr;
        if (W > 1) {
            NCK(g, g_nccl.GroupStart());
            for (int q = 1; q < W; ++q) {
                const int nq = gr.n_of(q, W);
                if (nq == 0) continue;
                NCK(g, g_nccl.Recv(st + (size_t)q * slab, (size_t)nq * GL, ncclDouble, q, g->comm[li], cs));
                if (st_nu) NCK(g, g_nccl.Recv(st_nu + (size_t)q * slab, (size_t)nq * GL, ncclDouble, q, g->comm[li], cs));
                g->gathered_bytes += (long long)((size_t)nq * GL * sizeof(double) * (st_nu ? 2 : 1));
            }
            NCK(g, g_nccl.GroupEnd());
        }
        if (n_loc > 0) {
            CK(c, cudaMemcpyAsync(st, out, (size_t)n_loc * GL * sizeof(double), cudaMemcpyDeviceToDevice, cs));
            if (st_nu) CK(c, cudaMemcpyAsync(st_nu, nu, (size_t)n_loc * GL * sizeof(double), cudaMemcpyDeviceToDevice, cs));
        }
        if (gr.NE > 0) {
            const int blocks = std::min(gr.NE, 8 * c->sm_count);
            k_interleave<<<blocks, 128, 0, cs>>>(st, slab, W, gr.NE, (int)GL, gr.fin[p].as<double>());
            if (launch_check(c, "k_interleave")) return 1;
            if (st_nu) {
                k_interleave<<<blocks, 128, 0, cs>>>(st_nu, slab, W, gr.NE, (int)GL, gr.fin_nu[p].as<double>());
                if (launch_check(c, "k_interleave")) return 1;
            }
            // a column above the top group edge copies its predecessor, which another device may have computed
            k_copy_top<<<1, 256, 0, cs>>>(gr.d_E_full.as<double>(), gr.NE, gn->e_top, (int)GL, gr.fin[p].as<double>(),
                                          st_nu ? gr.fin_nu[p].as<double>() : nullptr);
            if (launch_check(c, "k_copy_top")) return 1;
        }
    }
    CK(c, cudaEventRecord(gn->ev_sent[p][li], cs));
    return 0;
}

int group_integrate(GroupNuclide* gn, int what)
{
    Group* g = gn->g;
    const int p = gn->calls & 1;
    gn->calls++;
    if (what & 1) gn->el.last = p;
    if (what & 2) gn->inel.last = p;
    return group_parallel(g, [&](int li) -> int {
        if ((what & 1) && group_integrate_one(gn, gn->el, false, li, p)) return 1;
        if ((what & 2) && group_integrate_one(gn, gn->inel, true, li, p)) return 1;
        return 0;
    });
}

int group_sync(GroupNuclide* gn)
{
    Group* g = gn->g;
    return group_parallel(g, [&](int li) -> int {
        Ctx* c = g->ctx[li];
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaStreamSynchronize(g->cs[li]));
        return 0;
    });
}

// root: latest assembled matrices to host arrays
int group_fetch(GroupNuclide* gn, double* el_mat, double* inel_mat, double* nuinel_mat)
{
    Group* g = gn->g;
    if (g->first != 0) return 0;   // not the root process: nothing arrives here
    Ctx* c = g->ctx[0];
    CK(c, cudaSetDevice(c->device));
    cudaStream_t cs = g->cs[0];
    const size_t GL = (size_t)gn->GL;
    if (el_mat && gn->el.NE > 0) {
        if (gn->el.last < 0) return gfail(g, "ndppgpu_group_fetch: the elastic matrix has not been integrated");
        CK(c, cudaMemcpyAsync(el_mat, gn->el.fin[gn->el.last].p, (size_t)gn->el.NE * GL * sizeof(double), cudaMemcpyDeviceToHost, cs));
        c->stats.d2h_bytes += (double)((size_t)gn->el.NE * GL * sizeof(double));
    }
    if ((inel_mat || nuinel_mat) && gn->inel.NE > 0) {
        if (gn->inel.last < 0) return gfail(g, "ndppgpu_group_fetch: the inelastic matrix has not been integrated");
        const size_t b = (size_t)gn->inel.NE * GL * sizeof(double);
        if (inel_mat) { CK(c, cudaMemcpyAsync(inel_mat, gn->inel.fin[gn->inel.last].p, b, cudaMemcpyDeviceToHost, cs)); c->stats.d2h_bytes += (double)b; }
        if (nuinel_mat) {
            if (gn->inel.nmat != 2) return gfail(g, "ndppgpu_group_fetch: nu-scatter was not requested");
            CK(c, cudaMemcpyAsync(nuinel_mat, gn->inel.fin_nu[gn->inel.last].p, b, cudaMemcpyDeviceToHost, cs));
            c->stats.d2h_bytes += (double)b;
        }
    }
    CK(c, cudaStreamSynchronize(cs));
    return 0;
}

// ---- the planner: (nuclide, matrix, tile) items, cost model of SURVEY 8d, LPT with a set-up cost --------------------
double flops_file4(int G, int L, int M, int g_act = 3) { return 12.0 + 22.0 * G + 40.0 * g_act + (double)(M + g_act) * (13 + 4 * (L - 2) + 10 * L); }
double flops_file6_cm(int G_b, int L, int M, int K, int NPu = 127)
{
    return 11.0 * M * NPu + (double)G_b * K * (17 + 71.0 * M + (double)(15 * L + 11) * (M - 1)) + 3.0 * G_b * L;
}
double flops_freegas(int /*G*/, int L) { return 40.0 * 3.2e7 * (L / 4.0); }   // ~3.2e7 calc_fgk per E_in at L = 4 (C3, counted)

void tile_bounds(int n, int tile, int n_tiles, int& lo, int& hi)
{
    const int base = n / n_tiles, rem = n % n_tiles;
    lo = tile * base + std::min(tile, rem);
    hi = lo + base + (tile < rem ? 1 : 0);
}

void make_items(const ndppgpu_shape* shapes, int n, int G, int L, int M, int K, int tile_rows, int cont_split,
                std::vector<ndppgpu_item>& items)
{
    items.clear();
    for (int k = 0; k < n; ++k) {
        const ndppgpu_shape& s = shapes[k];
        int nt = std::max(1, (s.n_el + tile_rows - 1) / tile_rows);
        for (int t = 0; t < nt; ++t) {
            int lo, hi;
            tile_bounds(s.n_el, t, nt, lo, hi);
            double cost = (double)(hi - lo) * 2 * flops_file4(G, L, M);
            const int fg = std::max(0, std::min(hi, s.freegas_points) - lo);
            cost += fg * flops_freegas(G, L);
            items.push_back({s.index, 0, t, nt, -1, 0, cost});
        }
        if (s.n_inel <= 0) continue;
        std::vector<double> thr(s.level_thresholds, s.level_thresholds + std::max(s.n_levels, 0));
        std::sort(thr.begin(), thr.end());
        double e0 = s.e_lo;
        if (!thr.empty() || s.has_cont) {
            e0 = 1e300;
            if (!thr.empty()) e0 = thr[0];
            if (s.has_cont) e0 = std::min(e0, s.cont_threshold);
        }
        const double a = std::log(std::max(e0, s.e_lo)), b = std::log(s.e_hi);
        nt = std::max(1, (s.n_inel + tile_rows - 1) / tile_rows);
        if (s.has_cont) nt = std::max(nt, std::min(s.n_inel, cont_split * nt));
        for (int t = 0; t < nt; ++t) {
            int lo, hi;
            tile_bounds(s.n_inel, t, nt, lo, hi);
            const double mid = (lo + hi) / 2.0 / std::max(s.n_inel, 1);
            const double E = std::exp(a + mid * (b - a));
            int n_lev = 0;
            for (double x : thr) if (x < E) ++n_lev;
            double cost = (double)(hi - lo) * 2 * n_lev * flops_file4(G, L, M);
            if (s.has_cont && E > s.cont_threshold) cost += (hi - lo) * flops_file6_cm(std::max(1, G - 5), L, M, K);
            items.push_back({s.index, 1, t, nt, -1, 0, cost});
        }
    }
}

// longest-processing-time-first, aware of the cost of opening a nuclide on a device; deterministic
void plan_lpt(std::vector<ndppgpu_item>& items, int world, double setup_cost)
{
    std::vector<int> order(items.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        const ndppgpu_item &a = items[x], &b = items[y];
        if (a.cost != b.cost) return a.cost > b.cost;
        if (a.nuclide != b.nuclide) return a.nuclide < b.nuclide;
        if (a.matrix != b.matrix) return a.matrix < b.matrix;
        return a.tile < b.tile;
    });
    std::vector<double> load(world, 0.0);
    std::vector<std::set<int>> have(world);
    for (int i : order) {
        ndppgpu_item& it = items[i];
        int best = 0;
        double best_t = 0.0;
        for (int r = 0; r < world; ++r) {
            const double t = load[r] + it.cost + (have[r].count(it.nuclide) ? 0.0 : setup_cost);
            if (r == 0 || t < best_t) { best = r; best_t = t; }
        }
        if (!have[best].count(it.nuclide)) { have[best].insert(it.nuclide); load[best] += setup_cost; }
        load[best] += it.cost;
        it.rank = best;
    }
}

// the reference's partition: contiguous blocks of nuclides, first ranks take the remainder (src/ndpp.F90:941-948)
void plan_static(std::vector<ndppgpu_item>& items, const ndppgpu_shape* shapes, int n, int world)
{
    std::map<int, int> owner;
    const int base = n / world, rem = n % world;
    int k = 0;
    for (int r = 0; r < world; ++r) {
        const int cnt = base + (r < rem ? 1 : 0);
        for (int j = 0; j < cnt; ++j) owner[shapes[k + j].index] = r;
        k += cnt;
    }
    for (auto& it : items) it.rank = owner[it.nuclide];
}

double plan_imbalance(const std::vector<ndppgpu_item>& items, int world)
{
    std::vector<double> load(world, 0.0);
    for (auto& it : items) if (it.rank >= 0 && it.rank < world) load[it.rank] += it.cost;
    double mx = 0.0, sum = 0.0;
    for (double l : load) { mx = std::max(mx, l); sum += l; }
    return sum > 0 ? mx / (sum / world) : 1.0;
}

// ---- library run ------------------------------------------------------------------------------------------------------
struct LibPiece { int nuclide, matrix, tile, n_tiles, rank, rows; size_t off_rows; };  // off_rows: row offset in rank's buffer

struct Library {
    Group* g = nullptr;
    int G = 0, L = 0, nuscatter = 0;
    std::vector<ndppgpu_item> items;             // the plan, rows filled in by ndppgpu_library_set_plan
    std::vector<LibPiece> pieces;                // same order as items
    std::vector<size_t> rows_of_rank;            // [world]
    std::vector<DevBuf> flat, flat_nu;           // [n_local] result buffer of each local device
    DevBuf parts, parts_nu;                      // root: [sum rows][GL] received buffers, rank after rank
    std::vector<size_t> part_off;                // [world] row offset of a rank's buffer inside parts
    bool ran = false;
    ndppgpu_library_report rep{};
};

}  // namespace

// ======================================================================================================================
extern "C" {

int ndppgpu_group_unique_id(void* id128)
{
    if (!id128) return fail(nullptr, "ndppgpu_group_unique_id: null argument");
    const std::string e = g_nccl.load();
    if (!e.empty()) return fail(nullptr, e);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    std::memcpy(id128, &id, 128);
    return 0;
}

static int group_finish_init(std::unique_ptr<Group>& g, void** group)
{
    if (group_make_streams(g.get())) return 1;
    *group = g.release();
    return 0;
}

int ndppgpu_group_init(int n_devices, const int* devices, void** group)
{
    if (!group) return fail(nullptr, "ndppgpu_group_init: null group");
    *group = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, std::string("ndppgpu_group_init: no CUDA device (") + cudaGetErrorString(e) +
                                 "); this library has no CPU fallback");
    if (n_devices <= 0) n_devices = count;   // all GPUs of the box
    if (n_devices > count) return fail(nullptr, "ndppgpu_group_init: more devices requested than the box has");
    std::unique_ptr<Group> g(new Group());
    g->world = g->n_local = n_devices;
    g->first = 0;
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; ++i) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < n_devices; ++i) {
        void* c = nullptr;
        if (ndppgpu_init(devs[i], &c)) { for (Ctx* q : g->ctx) ndppgpu_finalize(q); return 1; }
        g->ctx.push_back((Ctx*)c);
    }
    g->comm.assign(n_devices, nullptr);
    if (n_devices > 1) {
        const std::string le = g_nccl.load();
        if (!le.empty()) return gfail(g.get(), le);
        NCK(g.get(), g_nccl.CommInitAll(g->comm.data(), n_devices, devs.data()));
    }
    return group_finish_init(g, group);
}

int ndppgpu_group_init_rank(int device, int rank, int world, const void* id128, void** group)
{
    if (!group) return fail(nullptr, "ndppgpu_group_init_rank: null group");
    *group = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(nullptr, "ndppgpu_group_init_rank: rank outside 0..world-1");
    if (world > 1 && !id128) return fail(nullptr, "ndppgpu_group_init_rank: null NCCL id");
    std::unique_ptr<Group> g(new Group());
    g->world = world; g->n_local = 1; g->first = rank;
    void* c = nullptr;
    if (ndppgpu_init(device, &c)) return 1;
    g->ctx.push_back((Ctx*)c);
    g->comm.assign(1, nullptr);
    if (world > 1) {
        const std::string le = g_nccl.load();
        if (!le.empty()) return gfail(g.get(), le);
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        CK(g->ctx[0], cudaSetDevice(g->ctx[0]->device));
        NCK(g.get(), g_nccl.CommInitRank(&g->comm[0], world, id, rank));
    }
    return group_finish_init(g, group);
}

int ndppgpu_group_info(void* group, int* world, int* n_local, int* first_rank)
{
    Group* g = (Group*)group;
    if (!g) return fail(nullptr, "ndppgpu_group_info: null group");
    if (world) *world = g->world;
    if (n_local) *n_local = g->n_local;
    if (first_rank) *first_rank = g->first;
    return 0;
}

void* ndppgpu_group_ctx(void* group, int local_index)
{
    Group* g = (Group*)group;
    return (g && local_index >= 0 && local_index < g->n_local) ? (void*)g->ctx[local_index] : nullptr;
}

long long ndppgpu_group_gathered_bytes(void* group, int reset)
{
    Group* g = (Group*)group;
    if (!g) return 0;
    const long long v = g->gathered_bytes;
    if (reset) g->gathered_bytes = 0;
    return v;
}

int ndppgpu_group_finalize(void* group)
{
    Group* g = (Group*)group;
    if (!g) return 0;
    for (int li = 0; li < g->n_local; ++li) {
        cudaSetDevice(g->ctx[li]->device);
        if (li < (int)g->cs.size() && g->cs[li]) { cudaStreamSynchronize(g->cs[li]); cudaStreamDestroy(g->cs[li]); }
        if (g->comm[li]) g_nccl.CommDestroy(g->comm[li]);
        ndppgpu_finalize(g->ctx[li]);
    }
    delete g;
    return 0;
}

// ---- the nuclide of a group: ndppgpu_nuclide_* one to one ------------------------------------------------------------
int ndppgpu_group_nuclide_create(void* group, double awr, double kT, double freegas_cutoff, int n_grid, const double* energy,
                                 const double* elastic_xs, const double* e_bins, int n_bins, const ndppgpu_params* params,
                                 void** gnuc)
{
    Group* g = (Group*)group;
    if (!g || !gnuc || !params || !e_bins) return gfail(g, "ndppgpu_group_nuclide_create: null argument");
    *gnuc = nullptr;
    std::unique_ptr<GroupNuclide> gn(new GroupNuclide());
    gn->g = g;
    gn->nuc.assign(g->n_local, nullptr);
    const int rc = group_parallel(g, [&](int li) -> int {
        void* n = nullptr;
        if (ndppgpu_nuclide_create(g->ctx[li], awr, kT, freegas_cutoff, n_grid, energy, elastic_xs, e_bins, n_bins, params, &n))
            return 1;
        gn->nuc[li] = (Nuclide*)n;
        return 0;
    });
    auto cleanup = [&]() { for (Nuclide* n : gn->nuc) if (n) ndppgpu_nuclide_free(n); };
    if (rc) { cleanup(); return 1; }
    gn->GL = gn->nuc[0]->G * gn->nuc[0]->L;
    gn->e_top = e_bins[n_bins - 1];
    gn->nuscatter = params->nuscatter != 0;
    gn->ev_comp.assign(g->n_local, nullptr);
    for (int p = 0; p < 2; ++p) gn->ev_sent[p].assign(g->n_local, nullptr);
    for (int li = 0; li < g->n_local; ++li) {
        cudaSetDevice(g->ctx[li]->device);
        if (cudaEventCreateWithFlags(&gn->ev_comp[li], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&gn->ev_sent[0][li], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&gn->ev_sent[1][li], cudaEventDisableTiming) != cudaSuccess) {
            cleanup();
            return gfail(g, "ndppgpu_group_nuclide_create: cudaEventCreate failed");
        }
    }
    *gnuc = gn.release();
    return 0;
}

int ndppgpu_group_nuclide_add_reaction(void* gnuc, int rxn_index, int MT, double Q_value, int threshold, int scatter_in_cm,
                                       int has_angle_dist, int has_energy_dist, int law, int multiplicity,
                                       const double* yield_tab1, int n_yield, const double* sigma, int n_sigma,
                                       const double* p_valid_tab1, int n_pvalid, const double* adist_energy,
                                       const int* adist_type, const int* adist_loc, int n_adist_e, const double* adist_data,
                                       int n_adist_data, const double* edist_data, int n_edist_data)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_nuclide_add_reaction: null nuclide");
    for (Nuclide* n : gn->nuc)   // host-side bookkeeping only: no need for threads
        if (ndppgpu_nuclide_add_reaction(n, rxn_index, MT, Q_value, threshold, scatter_in_cm, has_angle_dist, has_energy_dist,
                                         law, multiplicity, yield_tab1, n_yield, sigma, n_sigma, p_valid_tab1, n_pvalid,
                                         adist_energy, adist_type, adist_loc, n_adist_e, adist_data, n_adist_data, edist_data,
                                         n_edist_data))
            return gfail(gn->g, g_last_error);
    return 0;
}

int ndppgpu_group_convert_distro(void* gnuc)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_convert_distro: null nuclide");
    return group_parallel(gn->g, [&](int li) -> int { return ndppgpu_convert_distro(gn->nuc[li]); });
}

int ndppgpu_group_set_grids(void* gnuc, const double* Ein_el, int NE_el, const double* Ein_inel, int NE_inel)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_set_grids: null nuclide");
    if ((NE_el > 0 && !Ein_el) || (NE_inel > 0 && !Ein_inel)) return gfail(gn->g, "ndppgpu_group_set_grids: null grid");
    if (NE_el >= 0 && group_grid_set(gn, gn->el, Ein_el, NE_el, 1)) return 1;                 // a negative count leaves
    if (NE_inel >= 0 && group_grid_set(gn, gn->inel, Ein_inel, NE_inel, gn->nuscatter ? 2 : 1)) return 1;   // the grid as it is
    return 0;
}

int ndppgpu_group_integrate(void* gnuc, int what)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_integrate: null nuclide");
    if (!(what & 3)) return gfail(gn->g, "ndppgpu_group_integrate: what must be 1 (elastic), 2 (inelastic) or 3");
    return group_integrate(gn, what & 3);
}

int ndppgpu_group_sync(void* gnuc)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_sync: null nuclide");
    return group_sync(gn);
}

// every device's compute stream waits for its side stream: an event recorded on ndppgpu_stream(ndppgpu_group_ctx(..))
// after this call lies behind the gather and the assembly of every integrate issued so far (device-side timing)
int ndppgpu_group_join(void* gnuc)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_join: null nuclide");
    Group* g = gn->g;
    for (int li = 0; li < g->n_local; ++li) {
        Ctx* c = g->ctx[li];
        CK(c, cudaSetDevice(c->device));
        for (int p = 0; p < 2; ++p) CK(c, cudaStreamWaitEvent(c->stream, gn->ev_sent[p][li], 0));
    }
    CK(g->ctx[0], cudaSetDevice(g->ctx[0]->device));
    return 0;
}

int ndppgpu_group_fetch(void* gnuc, double* el_mat, double* inel_mat, double* nuinel_mat)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_fetch: null nuclide");
    if (group_sync(gn)) return 1;
    return group_fetch(gn, el_mat, inel_mat, nuinel_mat);
}

// device pointer (on the root device) of the latest assembled matrix; 0 elastic, 1 inelastic, 2 nu-inelastic
void* ndppgpu_group_result_dev(void* gnuc, int matrix)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn || gn->g->first != 0) return nullptr;
    GroupGrid& gr = matrix == 0 ? gn->el : gn->inel;
    if (gr.last < 0) return nullptr;
    return matrix == 2 ? gr.fin_nu[gr.last].p : gr.fin[gr.last].p;
}

int ndppgpu_group_elastic(void* gnuc, const double* Ein, int NE, double* el_mat)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_elastic: null nuclide");
    if (NE <= 0) return 0;
    if (!Ein) return gfail(gn->g, "ndppgpu_group_elastic: null argument");
    if (group_grid_set(gn, gn->el, Ein, NE, 1) || group_integrate(gn, 1) || group_sync(gn)) return 1;
    return group_fetch(gn, el_mat, nullptr, nullptr);
}

int ndppgpu_group_inelastic(void* gnuc, const double* Ein, int NE, double* inel_mat, double* nuinel_mat)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return fail(nullptr, "ndppgpu_group_inelastic: null nuclide");
    if (NE <= 0) return 0;
    if (!Ein) return gfail(gn->g, "ndppgpu_group_inelastic: null argument");
    // nu-scatter buffers exist when the nuclide was created with params.nuscatter, as in the reference (scatt.F90:711-716)
    if (group_grid_set(gn, gn->inel, Ein, NE, gn->nuscatter ? 2 : 1) || group_integrate(gn, 2) || group_sync(gn)) return 1;
    return group_fetch(gn, nullptr, inel_mat, gn->nuscatter ? nuinel_mat : nullptr);
}

int ndppgpu_group_nuclide_free(void* gnuc)
{
    GroupNuclide* gn = (GroupNuclide*)gnuc;
    if (!gn) return 0;
    Group* g = gn->g;
    for (int li = 0; li < g->n_local; ++li) {
        cudaSetDevice(g->ctx[li]->device);
        cudaStreamSynchronize(g->ctx[li]->stream);
        cudaStreamSynchronize(g->cs[li]);
        if (gn->ev_comp[li]) cudaEventDestroy(gn->ev_comp[li]);
        for (int p = 0; p < 2; ++p) if (gn->ev_sent[p][li]) cudaEventDestroy(gn->ev_sent[p][li]);
    }
    // device buffers of the grids belong to their devices: release them with the right device current
    auto drop = [&](GroupGrid& gr) {
        for (int li = 0; li < g->n_local; ++li) {
            cudaSetDevice(g->ctx[li]->device);
            if (li < (int)gr.d_E.size()) gr.d_E[li].reset();
            for (int p = 0; p < 2; ++p) {
                if (li < (int)gr.d_out[p].size()) gr.d_out[p][li].reset();
                if (li < (int)gr.d_nu[p].size()) gr.d_nu[p][li].reset();
            }
        }
        cudaSetDevice(g->ctx[0]->device);
    };
    drop(gn->el); drop(gn->inel);
    for (Nuclide* n : gn->nuc) if (n) ndppgpu_nuclide_free(n);
    cudaSetDevice(g->ctx[0]->device);
    delete gn;   // root-device buffers (stage / fin) are freed here, device 0 current
    return 0;
}

// ---- planner (host only: usable without a GPU) -----------------------------------------------------------------------
int ndppgpu_plan_library(const ndppgpu_shape* shapes, int n_shapes, int G, int L, int M, int K, int tile_rows, int world,
                         double setup_cost, int policy, ndppgpu_item* items_out, int max_items, int* n_items,
                         double* imbalance)
{
    if (!shapes || !n_items || n_shapes < 0 || world < 1 || tile_rows < 1)
        return fail(nullptr, "ndppgpu_plan_library: bad argument");
    std::vector<ndppgpu_item> items;
    // Continuum tiles are ~1e3 x heavier per row than the others; they are cut only as fine as balance needs -- the
    // coarsest split (1, 2, 4, 8 x) whose heaviest item stays below 1/16 of a device's share -- because every extra
    // tile is another launch of the persistent file-6 kernel with its own tail and another device that opens the nuclide.
    for (int split : {1, 2, 4, 8}) {
        make_items(shapes, n_shapes, G, L, M, K, tile_rows, split, items);
        double total = 0.0, mx = 0.0;
        for (auto& it : items) { total += it.cost; mx = std::max(mx, it.cost); }
        if (items.empty() || mx <= total / (16.0 * world)) break;
    }
    if (policy == 1) plan_static(items, shapes, n_shapes, world);
    else plan_lpt(items, world, setup_cost);
    // per device: one table upload per nuclide
    std::stable_sort(items.begin(), items.end(), [](const ndppgpu_item& a, const ndppgpu_item& b) {
        if (a.rank != b.rank) return a.rank < b.rank;
        if (a.nuclide != b.nuclide) return a.nuclide < b.nuclide;
        if (a.matrix != b.matrix) return a.matrix < b.matrix;
        return a.tile < b.tile;
    });
    *n_items = (int)items.size();
    if (imbalance) *imbalance = plan_imbalance(items, world);
    if (items_out) {
        if ((int)items.size() > max_items) return fail(nullptr, "ndppgpu_plan_library: items_out too small");
        std::copy(items.begin(), items.end(), items_out);
    }
    return 0;
}

void ndppgpu_tile_bounds(int n, int tile, int n_tiles, int* lo, int* hi)
{
    int a = 0, b = 0;
    if (n_tiles > 0) tile_bounds(n, tile, n_tiles, a, b);
    if (lo) *lo = a;
    if (hi) *hi = b;
}

// ---- library run -------------------------------------------------------------------------------------------------------
int ndppgpu_library_create(void* group, int G, int L, int nuscatter, const ndppgpu_item* items, int n_items, void** lib)
{
    Group* g = (Group*)group;
    if (!g || !lib || (n_items > 0 && !items)) return gfail(g, "ndppgpu_library_create: null argument");
    *lib = nullptr;
    std::unique_ptr<Library> b(new Library());
    b->g = g; b->G = G; b->L = L; b->nuscatter = nuscatter;
    b->items.assign(items, items + n_items);
    b->rows_of_rank.assign(g->world, 0);
    for (auto& it : b->items) {
        if (it.rank < 0 || it.rank >= g->world) return gfail(g, "ndppgpu_library_create: item with a rank outside the group");
        if (it.rows < 0) return gfail(g, "ndppgpu_library_create: item without its number of rows");
        b->pieces.push_back({it.nuclide, it.matrix, it.tile, it.n_tiles, it.rank, it.rows, b->rows_of_rank[it.rank]});
        b->rows_of_rank[it.rank] += (size_t)it.rows;
    }
    b->part_off.assign(g->world + 1, 0);
    for (int r = 0; r < g->world; ++r) b->part_off[r + 1] = b->part_off[r] + b->rows_of_rank[r];
    *lib = b.release();
    return 0;
}

int ndppgpu_library_run(void* lib, ndppgpu_open_fn open, ndppgpu_close_fn close, void* user, ndppgpu_library_report* rep)
{
    Library* b = (Library*)lib;
    if (!b || !open) return fail(nullptr, "ndppgpu_library_run: null argument");
    Group* g = b->g;
    const size_t GL = (size_t)b->G * b->L;
    const int W = g->world;
    b->flat.clear(); b->flat.resize(g->n_local);
    b->flat_nu.clear(); b->flat_nu.resize(g->n_local);
    std::vector<double> t_open(g->n_local, 0.0), t_int(g->n_local, 0.0), t_alloc(g->n_local, 0.0), k_ms(g->n_local, 0.0);
    std::vector<int> opens(g->n_local, 0);
    std::vector<cudaEvent_t> ev0(g->n_local, nullptr), ev1(g->n_local, nullptr);   // device-side clock of the whole run
    std::vector<double> dev_ms(g->n_local, 0.0);
    const auto t_all = std::chrono::steady_clock::now();
    auto secs = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    };
    int rc = group_parallel(g, [&](int li) -> int {
        Ctx* c = g->ctx[li];
        const int r = g->first + li;
        ndppgpu_stats_t s0{};
        ndppgpu_stats(c, &s0, 0);
        CK(c, cudaEventCreate(&ev0[li]));
        CK(c, cudaEventCreate(&ev1[li]));
        CK(c, cudaEventRecord(ev0[li], c->stream));
        const auto t_a = std::chrono::steady_clock::now();
        double *flat = nullptr, *flat_nu = nullptr;
        if (r == 0) {   // the root integrates in place into the buffer the other devices' results are gathered into
            const size_t tot = std::max<size_t>(b->part_off[W], 1);
            if (dev_alloc(c, b->parts, tot * GL * sizeof(double))) return 1;
            if (b->nuscatter && dev_alloc(c, b->parts_nu, tot * GL * sizeof(double))) return 1;
            flat = b->parts.as<double>(); flat_nu = b->parts_nu.as<double>();
        } else {
            if (dev_alloc(c, b->flat[li], std::max<size_t>(b->rows_of_rank[r], 1) * GL * sizeof(double))) return 1;
            if (b->nuscatter && dev_alloc(c, b->flat_nu[li], std::max<size_t>(b->rows_of_rank[r], 1) * GL * sizeof(double))) return 1;
            flat = b->flat[li].as<double>(); flat_nu = b->flat_nu[li].as<double>();
        }
        t_alloc[li] = secs(t_a);
        void* nuc = nullptr;
        int cur = -1;
        const double *Ein_el = nullptr, *Ein_inel = nullptr;
        int NE_el = 0, NE_inel = 0;
        TmpBuf d_Eel, d_Einel;
        auto shut = [&]() { if (nuc) { if (close) close(user, cur, nuc); else ndppgpu_nuclide_free(nuc); nuc = nullptr; } };
        for (const LibPiece& pc : b->pieces) {
            if (pc.rank != r) continue;
            if (pc.nuclide != cur) {
                shut();
                const auto t0 = std::chrono::steady_clock::now();
                cur = pc.nuclide;
                if (open(user, cur, c, &nuc, &Ein_el, &NE_el, &Ein_inel, &NE_inel) || !nuc) {
                    if (g_last_error.empty()) g_last_error = "ndppgpu_library_run: the open callback failed for nuclide " + std::to_string(cur);
                    return 1;
                }
                if (!((Nuclide*)nuc)->converted && ndppgpu_convert_distro(nuc)) { shut(); return 1; }
                if (tmp_upload(c, d_Eel, Ein_el, (size_t)std::max(NE_el, 0)) || tmp_upload(c, d_Einel, Ein_inel, (size_t)std::max(NE_inel, 0))) { shut(); return 1; }
                CK(c, cudaStreamSynchronize(c->stream));   // the grids are the caller's pageable memory
                opens[li]++;
                t_open[li] += secs(t0);
            }
            const auto t0 = std::chrono::steady_clock::now();
            const int NE = pc.matrix == 0 ? NE_el : NE_inel;
            int lo, hi;
            tile_bounds(NE, pc.tile, pc.n_tiles, lo, hi);
            if (hi - lo != pc.rows) {
                shut();
                g_last_error = "ndppgpu_library_run: tile of nuclide " + std::to_string(cur) + " has " + std::to_string(hi - lo) +
                               " rows, the plan says " + std::to_string(pc.rows);
                return 1;
            }
            double* out = flat + pc.off_rows * GL;
            int e = 0;
            if (pc.matrix == 0) e = elastic_dev((Nuclide*)nuc, d_Eel.as<double>() + lo, hi - lo, out);
            else e = inelastic_dev((Nuclide*)nuc, d_Einel.as<double>() + lo, hi - lo, out,
                                   b->nuscatter ? flat_nu + pc.off_rows * GL : nullptr);
            if (e) { shut(); return 1; }
            t_int[li] += secs(t0);
        }
        shut();
        CK(c, cudaStreamSynchronize(c->stream));
        ndppgpu_stats_t s1{};
        ndppgpu_stats(c, &s1, 0);
        k_ms[li] = s1.kernel_ms - s0.kernel_ms;
        return 0;
    });
    if (rc) return 1;
    const double t_compute = secs(t_all);
    // one gather: every device's buffer to the root
    const auto t_g = std::chrono::steady_clock::now();
    rc = group_parallel(g, [&](int li) -> int {
        Ctx* c = g->ctx[li];
        const int r = g->first + li;
        cudaStream_t cs = g->cs[li];
        const size_t n = b->rows_of_rank[r] * GL;
        if (r == 0) {
            if (W > 1) {
                NCK(g, g_nccl.GroupStart());
                for (int q = 1; q < W; ++q) {
                    const size_t nq = b->rows_of_rank[q] * GL;
                    if (!nq) continue;
                    NCK(g, g_nccl.Recv(b->parts.as<double>() + b->part_off[q] * GL, nq, ncclDouble, q, g->comm[li], cs));
                    if (b->nuscatter) NCK(g, g_nccl.Recv(b->parts_nu.as<double>() + b->part_off[q] * GL, nq, ncclDouble, q, g->comm[li], cs));
                    g->gathered_bytes += (long long)(nq * sizeof(double) * (b->nuscatter ? 2 : 1));
                }
                NCK(g, g_nccl.GroupEnd());
            }
        } else if (n) {
            NCK(g, g_nccl.GroupStart());
            NCK(g, g_nccl.Send(b->flat[li].p, n, ncclDouble, 0, g->comm[li], cs));
            if (b->nuscatter) NCK(g, g_nccl.Send(b->flat_nu[li].p, n, ncclDouble, 0, g->comm[li], cs));
            NCK(g, g_nccl.GroupEnd());
        }
        CK(c, cudaEventRecord(ev1[li], cs));
        CK(c, cudaStreamSynchronize(cs));
        float ms = 0.f;
        if (ev0[li] && cudaEventElapsedTime(&ms, ev0[li], ev1[li]) == cudaSuccess) dev_ms[li] = ms;
        return 0;
    });
    for (int li = 0; li < g->n_local; ++li) {
        cudaSetDevice(g->ctx[li]->device);
        if (ev0[li]) cudaEventDestroy(ev0[li]);
        if (ev1[li]) cudaEventDestroy(ev1[li]);
    }
    if (rc) return 1;
    // local buffers are no longer needed
    for (int li = 0; li < g->n_local; ++li) {
        cudaSetDevice(g->ctx[li]->device);
        b->flat[li].reset();
        b->flat_nu[li].reset();
    }
    cudaSetDevice(g->ctx[0]->device);
    b->ran = true;
    ndppgpu_library_report& R = b->rep;
    R = ndppgpu_library_report{};
    R.wall_s = secs(t_all);
    R.compute_s = t_compute;
    R.gather_s = secs(t_g);
    for (int li = 0; li < g->n_local; ++li) {
        R.opens += opens[li];
        R.open_s_max = std::max(R.open_s_max, t_open[li]);
        R.integrate_s_max = std::max(R.integrate_s_max, t_int[li]);
        R.kernel_s_max = std::max(R.kernel_s_max, k_ms[li] * 1e-3);
        R.kernel_s_sum += k_ms[li] * 1e-3;
        R.device_s_max = std::max(R.device_s_max, dev_ms[li] * 1e-3);
        R.alloc_s_max = std::max(R.alloc_s_max, t_alloc[li]);
    }
    R.items = (int)b->items.size();
    for (auto& pc : b->pieces) R.moment_evals += (long long)pc.rows * (long long)GL * (pc.matrix == 1 && b->nuscatter ? 2 : 1);
    if (rep) *rep = R;
    return 0;
}

// root: matrix (0 elastic, 1 inelastic, 2 nu-inelastic) of one nuclide, tiles in order, top-of-grid rule applied
// (src/scatt.F90:669,770: a column above the top group edge copies its predecessor, which may belong to another tile).
int ndppgpu_library_fetch(void* lib, int nuclide, int matrix, const double* Ein, int NE, double e_top, double* mat)
{
    Library* b = (Library*)lib;
    if (!b || !mat || (NE > 0 && !Ein)) return fail(nullptr, "ndppgpu_library_fetch: null argument");
    Group* g = b->g;
    if (!b->ran) return gfail(g, "ndppgpu_library_fetch: ndppgpu_library_run has not completed");
    if (g->first != 0) return gfail(g, "ndppgpu_library_fetch: only the root process holds the gathered matrices");
    if (matrix == 2 && !b->nuscatter) return gfail(g, "ndppgpu_library_fetch: nu-scatter was not requested");
    Ctx* c = g->ctx[0];
    CK(c, cudaSetDevice(c->device));
    const size_t GL = (size_t)b->G * b->L;
    const int want = matrix == 0 ? 0 : 1;
    std::vector<const LibPiece*> ps;
    for (auto& pc : b->pieces) if (pc.nuclide == nuclide && pc.matrix == want) ps.push_back(&pc);
    std::sort(ps.begin(), ps.end(), [](const LibPiece* a, const LibPiece* b2) { return a->tile < b2->tile; });
    if (ps.empty()) return gfail(g, "ndppgpu_library_fetch: no such matrix in the plan");
    size_t row = 0;
    double* base = (matrix == 2 ? b->parts_nu : b->parts).as<double>();
    // first column above the top group edge: it and its successors all take column j (the last one that was integrated)
    int i0 = NE;
    for (int i = 0; i < NE; ++i) if (!(Ein[i] <= e_top)) { i0 = i; break; }
    const int j = i0 > 0 ? i0 - 1 : 0;
    const double* col_j = nullptr;
    for (size_t k = 0; k < ps.size(); ++k) {
        if (ps[k]->tile != (int)k || ps[k]->n_tiles != (int)ps.size()) return gfail(g, "ndppgpu_library_fetch: missing tile");
        const double* src = base + (b->part_off[ps[k]->rank] + ps[k]->off_rows) * GL;
        CK(c, cudaMemcpyAsync(mat + row * GL, src, (size_t)ps[k]->rows * GL * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((size_t)j >= row && (size_t)j < row + (size_t)ps[k]->rows) col_j = src + ((size_t)j - row) * GL;
        row += (size_t)ps[k]->rows;
    }
    if ((int)row != NE) return gfail(g, "ndppgpu_library_fetch: the tiles hold " + std::to_string(row) + " columns, the grid has " + std::to_string(NE));
    for (int i = std::max(i0, 1); i < NE && col_j; ++i)   // device -> host copies of column j, nothing is computed here
        CK(c, cudaMemcpyAsync(mat + (size_t)i * GL, col_j, GL * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += (double)(row * GL * sizeof(double));
    return 0;
}

int ndppgpu_library_report_get(void* lib, ndppgpu_library_report* rep)
{
    Library* b = (Library*)lib;
    if (!b || !rep) return fail(nullptr, "ndppgpu_library_report_get: null argument");
    *rep = b->rep;
    return 0;
}

int ndppgpu_library_free(void* lib)
{
    Library* b = (Library*)lib;
    if (!b) return 0;
    Group* g = b->g;
    for (int li = 0; li < g->n_local && li < (int)b->flat.size(); ++li) {
        cudaSetDevice(g->ctx[li]->device);
        b->flat[li].reset();
        if (li < (int)b->flat_nu.size()) b->flat_nu[li].reset();
    }
    cudaSetDevice(g->ctx[0]->device);
    delete b;
    return 0;
}

}  // extern "C"
