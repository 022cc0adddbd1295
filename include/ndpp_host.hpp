// ndpp_host.hpp -- C++ host side above the C-ABI of libndppgpu.so (include/ndppgpu.h).
//
// The reference is compiled Fortran; no Fortran compiler exists in this project's image, so the host
// layer a maintainer would write with ISO_C_BINDING is mirrored here in C++ with the
// reference's own names, argument meaning and error behaviour:
//
//   Tab1, DistAngle, DistEnergy, Reaction, Nuclide, SAlphaBeta  <- src/endf_header.F90:9-20,
//                                                                  src/ace_header.F90:14-164, 188-235
//   calc_scatt       <- src/scatt.F90:33-157      (rxn_data(:) set-up, convert_distro, calc_elastic_grid,
//                                                  calc_inelastic_grid, rxn_data(i) % clear())
//   calc_scattsab    <- src/scatt.F90:543-596
//   apply_tol_scatt  <- src/scatt.F90:786-818
//   thin_grid        <- src/thin.F90:19-47
//   fatal_error      <- src/error.F90:79-154      (here: throws FatalError carrying the message; the
//                                                  program's top level prints it and stops, as
//                                                  tools/ndpp_calc_scatt.cpp does)
//
// Index-valued fields keep the reference's conventions (threshold is 1-based, adist % location is the
// 0-based `lc` the reference adds 1 to).  Moment arrays are returned as flat vectors in Fortran order
// mat(L, G, NE) == C mat[iE][g][l].  Everything numerical happens in the CUDA library: there is no CPU
// path here, and a missing GPU surfaces as FatalError from the first call.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ndppgpu.h"

namespace ndpp_host {

// src/constants.F90:124-152
enum { HISTOGRAM = 1, LINEAR_LINEAR = 2, LINEAR_LOG = 3, LOG_LINEAR = 4, LOG_LOG = 5 };
enum { ANGLE_ISOTROPIC = 1, ANGLE_32_EQUI = 2, ANGLE_TABULAR = 3 };
enum { SCATT_TYPE_LEGENDRE = 0, SCATT_TYPE_TABULAR = 1 };
enum { SAB_SECONDARY_EQUAL = 0, SAB_SECONDARY_SKEWED = 1, SAB_SECONDARY_CONT = 2 };
enum { SAB_ELASTIC_DISCRETE = 3, SAB_ELASTIC_EXACT = 4 };

struct FatalError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void fatal_error(const std::string& message) { throw FatalError(message); }

// ---- the parsed ACE data model (input of the path) --------------------------------------------------
struct Tab1 {  // src/endf_header.F90:9-20
    std::vector<int> nbt, interp;
    std::vector<double> x, y;
    bool present = false;
    // [NR, NBT(NR), INT(NR), NP, x(NP), y(NP)]: the array interpolate_tab1 reads (src/interpolation.F90:24-60)
    std::vector<double> flatten() const
    {
        std::vector<double> a;
        a.push_back((double)nbt.size());
        for (int v : nbt) a.push_back((double)v);
        for (int v : interp) a.push_back((double)v);
        a.push_back((double)x.size());
        a.insert(a.end(), x.begin(), x.end());
        a.insert(a.end(), y.begin(), y.end());
        return a;
    }
};

struct DistAngle {  // src/ace_header.F90:14-24
    std::vector<double> energy;
    std::vector<int> type, location;
    std::vector<double> data;
};

struct DistEnergy {  // src/ace_header.F90:31-43
    int law = 0;
    std::vector<double> data;
    Tab1 p_valid;
    std::unique_ptr<DistEnergy> next;
};

struct Reaction {  // src/ace_header.F90:50-67
    int MT = 0;
    double Q_value = 0.0;
    int multiplicity = 1;
    int threshold = 1;  // 1-based index into Nuclide::energy
    bool scatter_in_cm = true;
    std::vector<double> sigma;
    bool has_angle_dist = false;
    DistAngle adist;
    std::unique_ptr<DistEnergy> edist;  // has_energy_dist == (edist != nullptr)
    Tab1 multiplicity_E;
};

struct Nuclide {  // src/ace_header.F90:94-164 (the fields calc_scatt reads)
    std::string name = "synthetic";
    double awr = 0.0, kT = 0.0, freegas_cutoff = 0.0;
    std::vector<double> energy, elastic;
    std::vector<Reaction> reactions;
};

struct DistEnergySab {  // src/ace_header.F90:188-194
    std::vector<double> e_out, e_out_pdf, mu;  // mu(n_mu, n_e_out) in Fortran order
};

struct SAlphaBeta {  // src/ace_header.F90:201-235
    std::string name = "synthetic.sab";
    double awr = 0.0, kT = 0.0, threshold_inelastic = 0.0, threshold_elastic = 0.0;
    int n_inelastic_e_in = 0, n_inelastic_e_out = 0, n_inelastic_mu = 0, secondary_mode = SAB_SECONDARY_EQUAL;
    std::vector<double> inelastic_e_in, inelastic_sigma;
    std::vector<double> inelastic_e_out;  // (NEo, NEi) Fortran order
    std::vector<double> inelastic_mu;     // (n_mu, NEo, NEi)
    std::vector<DistEnergySab> inelastic_data;  // continuous representation
    int elastic_mode = SAB_ELASTIC_DISCRETE, n_elastic_e_in = 0, n_elastic_mu = 0;
    std::vector<double> elastic_e_in, elastic_P, elastic_mu;  // elastic_mu(n_mu, NEe)
};

// run-time integration parameters other than the arguments of calc_scatt (src/global.F90:28-59) with the
// defaults of src/constants.F90:69-100
struct Settings {
    int ne_per_grp = 20, adaptive_mu_its = 15, adaptive_eout_its = 15;
    double sab_threshold = 1.0e-6, brent_mu_thresh = 1.0e-6, adaptive_mu_tol = 1.0e-7, adaptive_eout_tol = 1.0e-8;
};

// ---- context ------------------------------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = -1)
    {
        if (ndppgpu_abi_version() != NDPPGPU_ABI_VERSION) fatal_error("libndppgpu.so: ABI version mismatch");
        if (ndppgpu_init(device, &h_) != 0) fatal_error(last_error(nullptr));
    }
    // a context that belongs to a device group (ndppgpu_group_ctx): used, not owned
    Context(void* borrowed, bool) : h_(borrowed), owned_(false) {}
    ~Context() { if (h_ && owned_) ndppgpu_finalize(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    void* handle() const { return h_; }
    static std::string last_error(void* h)
    {
        char buf[1024];
        buf[0] = 0;
        ndppgpu_last_error(h, buf, (int)sizeof buf);
        return std::string(buf);
    }
    void check(int rc) const { if (rc != 0) fatal_error(last_error(h_)); }
    ndppgpu_stats_t stats(bool reset = false) const
    {
        ndppgpu_stats_t s;
        check(ndppgpu_stats(h_, &s, reset ? 1 : 0));
        return s;
    }

private:
    void* h_ = nullptr;
    bool owned_ = true;
};

// ---- several GPUs: the devices of this process working on one nuclide / one library (ndppgpu_group_*) -------------
// Replaces what the reference spreads over MPI ranks in its own driver (partition_work, src/ndpp.F90:934-950).
class DeviceGroup {
public:
    explicit DeviceGroup(int n_devices = 0)   // 0: every GPU of the box
    {
        if (ndppgpu_abi_version() != NDPPGPU_ABI_VERSION) fatal_error("libndppgpu.so: ABI version mismatch");
        if (ndppgpu_group_init(n_devices, nullptr, &g_) != 0) fatal_error(Context::last_error(nullptr));
        ndppgpu_group_info(g_, &world_, &n_local_, nullptr);
        root_.reset(new Context(ndppgpu_group_ctx(g_, 0), false));
    }
    ~DeviceGroup() { root_.reset(); if (g_) ndppgpu_group_finalize(g_); }
    DeviceGroup(const DeviceGroup&) = delete;
    DeviceGroup& operator=(const DeviceGroup&) = delete;
    void* handle() const { return g_; }
    int world() const { return world_; }
    const Context& root() const { return *root_; }       // context of the root device: receives the assembled matrices
    void check(int rc) const { root_->check(rc); }
    ndppgpu_stats_t stats(int local_index) const
    {
        ndppgpu_stats_t s;
        check(ndppgpu_stats(ndppgpu_group_ctx(g_, local_index), &s, 0));
        return s;
    }

private:
    void* g_ = nullptr;
    int world_ = 1, n_local_ = 1;
    std::unique_ptr<Context> root_;
};

namespace detail {
template <class T> inline const T* ptr_or_null(const std::vector<T>& v) { return v.empty() ? nullptr : v.data(); }
}  // namespace detail

// ---- rxn_data(:) of calc_scatt: one ScattData slot per (reaction, energy distribution) ------------------
class ScattDataSet {
public:
    // the two loops of src/scatt.F90:88-126: mySD % init for every slot, then convert_distro
    ScattDataSet(const Context& ctx, const Nuclide& nuc, const std::vector<double>& energy_bins, int scatt_type,
                 int order, int mu_bins, bool nuscatt, const Settings& st = Settings())
        : ctx_(ctx)
    {
        init(nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st);
    }
    // the same set replicated on every device of a group; calc_*_grid then shard the E_in grid over the devices and
    // gather the columns to the root (ndppgpu_group_*): same arguments, same results
    ScattDataSet(const DeviceGroup& group, const Nuclide& nuc, const std::vector<double>& energy_bins, int scatt_type,
                 int order, int mu_bins, bool nuscatt, const Settings& st = Settings())
        : ctx_(group.root()), group_(group.handle())
    {
        init(nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st);
    }
    ~ScattDataSet() { clear(); }
    ScattDataSet(const ScattDataSet&) = delete;
    ScattDataSet& operator=(const ScattDataSet&) = delete;

private:
    void init(const Nuclide& nuc, const std::vector<double>& energy_bins, int scatt_type, int order, int mu_bins,
              bool nuscatt, const Settings& st)
    {
        if (energy_bins.size() < 2) fatal_error("calc_scatt: energy_bins needs at least two edges");
        groups_ = (int)energy_bins.size() - 1;
        order_ = (scatt_type == SCATT_TYPE_LEGENDRE) ? order + 1 : order;  // src/scattdata_header.F90:114-118
        ndppgpu_params p;
        p.scatt_type = scatt_type; p.order = order; p.mu_bins = mu_bins; p.nuscatter = nuscatt ? 1 : 0;
        p.ne_per_grp = st.ne_per_grp; p.adaptive_mu_its = st.adaptive_mu_its;
        p.adaptive_eout_its = st.adaptive_eout_its; p.reserved = 0;
        p.sab_threshold = st.sab_threshold; p.brent_mu_thresh = st.brent_mu_thresh;
        p.adaptive_mu_tol = st.adaptive_mu_tol; p.adaptive_eout_tol = st.adaptive_eout_tol;
        if (nuc.energy.size() != nuc.elastic.size()) fatal_error("calc_scatt: nuc % energy and nuc % elastic differ in size");
        if (group_)
            ctx_.check(ndppgpu_group_nuclide_create(group_, nuc.awr, nuc.kT, nuc.freegas_cutoff, (int)nuc.energy.size(),
                                                    nuc.energy.data(), nuc.elastic.data(), energy_bins.data(),
                                                    (int)energy_bins.size(), &p, &h_));
        else
            ctx_.check(ndppgpu_nuclide_create(ctx_.handle(), nuc.awr, nuc.kT, nuc.freegas_cutoff, (int)nuc.energy.size(),
                                              nuc.energy.data(), nuc.elastic.data(), energy_bins.data(),
                                              (int)energy_bins.size(), &p, &h_));
        try {
            for (size_t i = 0; i < nuc.reactions.size(); ++i) {
                const Reaction& rxn = nuc.reactions[i];
                const DistEnergy* ed = rxn.edist.get();
                do {  // once for the reaction, once more per edist % next (src/scatt.F90:93-104)
                    add_slot((int)i, rxn, ed);
                    ed = ed ? ed->next.get() : nullptr;
                } while (ed != nullptr);
            }
            ctx_.check(group_ ? ndppgpu_group_convert_distro(h_) : ndppgpu_convert_distro(h_));
        } catch (...) {
            clear();
            throw;
        }
    }

public:
    int order() const { return order_; }    // L
    int groups() const { return groups_; }  // G

    // calc_elastic_grid (src/scatt.F90:603-675): allocates and fills el_mat(L, G, NE)
    void calc_elastic_grid(const std::vector<double>& Ein, std::vector<double>& el_mat) const
    {
        el_mat.assign(Ein.size() * (size_t)groups_ * order_, 0.0);
        ctx_.check(group_ ? ndppgpu_group_elastic(h_, Ein.data(), (int)Ein.size(), el_mat.data())
                          : ndppgpu_elastic(h_, Ein.data(), (int)Ein.size(), el_mat.data()));
    }
    // both grids in one call (src/scatt.F90:143-150): the elastic matrices are copied to the host while the inelastic
    // kernels run (ndppgpu_calc_scatt); on a device group the two calls above
    void calc_grids(const std::vector<double>& Ein_el, const std::vector<double>& Ein_inel, bool nuscatt,
                    std::vector<double>& el_mat, std::vector<double>& inel_mat, std::vector<double>& nuinel_mat) const
    {
        if (group_) {
            calc_elastic_grid(Ein_el, el_mat);
            inel_mat.clear(); nuinel_mat.clear();
            if (!Ein_inel.empty()) calc_inelastic_grid(Ein_inel, nuscatt, inel_mat, nuinel_mat);
            return;
        }
        const size_t w = (size_t)groups_ * order_;
        el_mat.assign(Ein_el.size() * w, 0.0);
        inel_mat.assign(Ein_inel.size() * w, 0.0);
        if (nuscatt && !Ein_inel.empty()) nuinel_mat.assign(inel_mat.size(), 0.0); else nuinel_mat.clear();
        auto out = [](std::vector<double>& v) { return v.empty() ? nullptr : v.data(); };
        ctx_.check(ndppgpu_calc_scatt(h_, detail::ptr_or_null(Ein_el), (int)Ein_el.size(), out(el_mat),
                                      detail::ptr_or_null(Ein_inel), (int)Ein_inel.size(), out(inel_mat), out(nuinel_mat)));
    }
    // calc_inelastic_grid (src/scatt.F90:682-778); nuinel_mat stays empty unless nuscatt
    void calc_inelastic_grid(const std::vector<double>& Ein, bool nuscatt, std::vector<double>& inel_mat,
                             std::vector<double>& nuinel_mat) const
    {
        inel_mat.assign(Ein.size() * (size_t)groups_ * order_, 0.0);
        if (nuscatt) nuinel_mat.assign(inel_mat.size(), 0.0); else nuinel_mat.clear();
        ctx_.check(group_ ? ndppgpu_group_inelastic(h_, Ein.data(), (int)Ein.size(), inel_mat.data(),
                                                    nuscatt ? nuinel_mat.data() : nullptr)
                          : ndppgpu_inelastic(h_, Ein.data(), (int)Ein.size(), inel_mat.data(),
                                              nuscatt ? nuinel_mat.data() : nullptr));
    }
    // calc_*_grid + apply_tol_scatt + thin_grid with only the kept columns copied back (src/ndpp.F90:607-648);
    // Ein and the matrices are cut to the points kept, as thin_grid re-allocates them
    void calc_elastic_thinned(std::vector<double>& Ein, double print_tol, double thin_tol,
                              const std::vector<double>& tokeep, std::vector<double>& el_mat, double& compression,
                              double& max_abs_err) const
    {
        const size_t w = (size_t)groups_ * order_;
        if (group_) {   // integrate on the group, then the two steps of src/ndpp.F90:611-648 on the root device
            calc_elastic_grid(Ein, el_mat);
            post(Ein, print_tol, thin_tol, tokeep, el_mat, nullptr, compression, max_abs_err);
            return;
        }
        el_mat.assign(Ein.size() * w, 0.0);
        int kept = 0;
        ctx_.check(ndppgpu_elastic_thinned(h_, Ein.data(), (int)Ein.size(), print_tol, thin_tol, tokeep.data(),
                                           (int)tokeep.size(), el_mat.data(), &kept, &compression, &max_abs_err));
        Ein.resize(kept);
        el_mat.resize(kept * w);
    }
    void calc_inelastic_thinned(std::vector<double>& Ein, bool nuscatt, double print_tol, double thin_tol,
                                const std::vector<double>& tokeep, std::vector<double>& inel_mat,
                                std::vector<double>& nuinel_mat, double& compression, double& max_abs_err) const
    {
        const size_t w = (size_t)groups_ * order_;
        if (group_) {
            calc_inelastic_grid(Ein, nuscatt, inel_mat, nuinel_mat);
            post(Ein, print_tol, thin_tol, tokeep, inel_mat, nuscatt ? &nuinel_mat : nullptr, compression, max_abs_err);
            return;
        }
        inel_mat.assign(Ein.size() * w, 0.0);
        if (nuscatt) nuinel_mat.assign(inel_mat.size(), 0.0); else nuinel_mat.clear();
        int kept = 0;
        ctx_.check(ndppgpu_inelastic_thinned(h_, Ein.data(), (int)Ein.size(), print_tol, thin_tol, tokeep.data(),
                                             (int)tokeep.size(), inel_mat.data(),
                                             nuscatt ? nuinel_mat.data() : nullptr, &kept, &compression, &max_abs_err));
        Ein.resize(kept);
        inel_mat.resize(kept * w);
        if (nuscatt) nuinel_mat.resize(kept * w);
    }
    // create_Ein_grid(rxn_data, E_bins, nuc % energy, awr, kT, cutoff, thresh, Ein_el, Ein_inel) (src/scatt.F90:166-236)
    // on the device; Ein_inel comes back empty for a nuclide with elastic scattering only (unallocated in the reference)
    void create_Ein_grid(std::vector<double>& Ein_el, std::vector<double>& Ein_inel, int extend_pts = 50,
                         int inel_extend_pts = 30, int* status = nullptr) const
    {
        if (group_) fatal_error("create_Ein_grid: call it on the nuclide of one device (the grids are inputs of the group calls)");
        int n_el = 0, n_inel = 0;
        ctx_.check(ndppgpu_nuclide_create_ein_grid(h_, extend_pts, inel_extend_pts, &n_el, &n_inel, status));
        Ein_el.assign((size_t)n_el, 0.0);
        Ein_inel.assign((size_t)n_inel, 0.0);
        ctx_.check(ndppgpu_nuclide_ein_grid(h_, 0, Ein_el.data(), nullptr));
        if (n_inel > 0) ctx_.check(ndppgpu_nuclide_ein_grid(h_, 1, Ein_inel.data(), nullptr));
    }
    // rxn_data(i) % clear() (src/scatt.F90:153-155)
    void clear()
    {
        if (h_) { if (group_) ndppgpu_group_nuclide_free(h_); else ndppgpu_nuclide_free(h_); }
        h_ = nullptr;
    }

private:
    // apply_tol_scatt and thin_grid of an assembled matrix set, on the root device
    void post(std::vector<double>& Ein, double print_tol, double thin_tol, const std::vector<double>& tokeep,
              std::vector<double>& mat, std::vector<double>* nu, double& compression, double& max_abs_err) const
    {
        const int NE = (int)Ein.size();
        compression = 0.0; max_abs_err = 0.0;
        if (NE == 0) return;
        ctx_.check(ndppgpu_apply_tol(ctx_.handle(), mat.data(), NE, groups_, order_, print_tol));
        if (nu) ctx_.check(ndppgpu_apply_tol(ctx_.handle(), nu->data(), NE, groups_, order_, print_tol));
        if (thin_tol > 0.0) {
            int kept = 0;
            ctx_.check(ndppgpu_thin_grid(ctx_.handle(), Ein.data(), mat.data(), nu ? nu->data() : nullptr, NE,
                                         groups_ * order_, tokeep.data(), (int)tokeep.size(), thin_tol, &kept, &compression,
                                         &max_abs_err));
            Ein.resize(kept);
            mat.resize((size_t)kept * groups_ * order_);
            if (nu) nu->resize((size_t)kept * groups_ * order_);
        }
    }
    void add_slot(int i_rxn, const Reaction& rxn, const DistEnergy* ed)
    {
        std::vector<double> yl, pv;
        if (rxn.multiplicity_E.present) yl = rxn.multiplicity_E.flatten();
        if (ed && ed->p_valid.present) pv = ed->p_valid.flatten();
        const DistAngle& ad = rxn.adist;
        const bool ha = rxn.has_angle_dist;
        if (ha && (ad.type.size() != ad.energy.size() || ad.location.size() != ad.energy.size()))
            fatal_error("calc_scatt: inconsistent angular-distribution arrays");
        ctx_.check((group_ ? ndppgpu_group_nuclide_add_reaction : ndppgpu_nuclide_add_reaction)(
            h_, i_rxn, rxn.MT, rxn.Q_value, rxn.threshold, rxn.scatter_in_cm ? 1 : 0, ha ? 1 : 0, ed ? 1 : 0,
            ed ? ed->law : 0, rxn.multiplicity, detail::ptr_or_null(yl), (int)yl.size(), detail::ptr_or_null(rxn.sigma),
            (int)rxn.sigma.size(), detail::ptr_or_null(pv), (int)pv.size(), ha ? detail::ptr_or_null(ad.energy) : nullptr,
            ha ? detail::ptr_or_null(ad.type) : nullptr, ha ? detail::ptr_or_null(ad.location) : nullptr,
            ha ? (int)ad.energy.size() : 0, ha ? detail::ptr_or_null(ad.data) : nullptr, ha ? (int)ad.data.size() : 0,
            ed ? detail::ptr_or_null(ed->data) : nullptr, ed ? (int)ed->data.size() : 0));
    }
    const Context& ctx_;
    void* group_ = nullptr;   // ndppgpu group handle when the set lives on a device group
    void* h_ = nullptr;
    int order_ = 0, groups_ = 0;
};

// calc_scatt (src/scatt.F90:33-157).  The E_in grids are inputs: create_Ein_grid (src/scatt.F90:166-536)
// stays on the caller's side of the seam.  el_mat / inel_mat / nuinel_mat are (re)allocated here and owned by
// the caller, as in the reference; inel_mat and nuinel_mat stay empty when Ein_inel is empty (:146-150).
// `order` is intent(inout) in the reference and is left unchanged, as its body leaves it.
inline void calc_scatt(const Context& ctx, const Nuclide& nuc, const std::vector<double>& energy_bins, int scatt_type,
                       int& order, int mu_bins, bool nuscatt, const std::vector<double>& Ein_el,
                       const std::vector<double>& Ein_inel, std::vector<double>& el_mat, std::vector<double>& inel_mat,
                       std::vector<double>& nuinel_mat, const Settings& st = Settings())
{
    ScattDataSet rxn_data(ctx, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st);
    rxn_data.calc_grids(Ein_el, Ein_inel, nuscatt, el_mat, inel_mat, nuinel_mat);
    rxn_data.clear();
}

// calc_scatt on a device group: the single-process, several-GPU form of the same call
inline void calc_scatt(const DeviceGroup& group, const Nuclide& nuc, const std::vector<double>& energy_bins, int scatt_type,
                       int& order, int mu_bins, bool nuscatt, const std::vector<double>& Ein_el,
                       const std::vector<double>& Ein_inel, std::vector<double>& el_mat, std::vector<double>& inel_mat,
                       std::vector<double>& nuinel_mat, const Settings& st = Settings())
{
    ScattDataSet rxn_data(group, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st);
    rxn_data.calc_elastic_grid(Ein_el, el_mat);
    inel_mat.clear();
    nuinel_mat.clear();
    if (!Ein_inel.empty()) rxn_data.calc_inelastic_grid(Ein_inel, nuscatt, inel_mat, nuinel_mat);
    rxn_data.clear();
}

// calc_scattsab (src/scatt.F90:543-596): scatt_mat(order+1, G, NE) for Legendre output, (order, G, NE)
// cosine bins for tabular output (a TODO in the reference, :579-588).  E_grid comes from
// sab_egrid (src/sab.F90:460).  mu_bins is accepted and unused, as in the reference.
inline void calc_scattsab(const Context& ctx, const SAlphaBeta& sab, const std::vector<double>& energy_bins,
                          int scatt_type, int order, std::vector<double>& scatt_mat, int mu_bins,
                          const std::vector<double>& E_grid)
{
    (void)mu_bins;
    if (energy_bins.size() < 2) fatal_error("calc_scattsab: energy_bins needs at least two edges");
    std::vector<int> cn;
    std::vector<double> ce, cp, cm;
    if (sab.secondary_mode == SAB_SECONDARY_CONT) {
        for (const DistEnergySab& d : sab.inelastic_data) {
            cn.push_back((int)d.e_out.size());
            ce.insert(ce.end(), d.e_out.begin(), d.e_out.end());
            cp.insert(cp.end(), d.e_out_pdf.begin(), d.e_out_pdf.end());
            cm.insert(cm.end(), d.mu.begin(), d.mu.end());
        }
    }
    using detail::ptr_or_null;
    void* h = nullptr;
    ctx.check(ndppgpu_sab_create(ctx.handle(), sab.awr, sab.kT, sab.threshold_inelastic, sab.threshold_elastic,
                                 sab.n_inelastic_e_in, sab.n_inelastic_e_out, sab.n_inelastic_mu, sab.secondary_mode,
                                 ptr_or_null(sab.inelastic_e_in), ptr_or_null(sab.inelastic_sigma),
                                 ptr_or_null(sab.inelastic_e_out), ptr_or_null(sab.inelastic_mu), ptr_or_null(cn),
                                 ptr_or_null(ce), ptr_or_null(cp), ptr_or_null(cm), sab.elastic_mode,
                                 sab.n_elastic_e_in, sab.n_elastic_mu, ptr_or_null(sab.elastic_e_in),
                                 ptr_or_null(sab.elastic_P), ptr_or_null(sab.elastic_mu), &h));
    const int L = (scatt_type == SCATT_TYPE_LEGENDRE) ? order + 1 : order;
    scatt_mat.assign(E_grid.size() * (energy_bins.size() - 1) * (size_t)L, 0.0);
    const int rc = ndppgpu_sab(h, energy_bins.data(), (int)energy_bins.size(), scatt_type, order, E_grid.data(),
                               (int)E_grid.size(), scatt_mat.data(), nullptr, nullptr);
    const std::string msg = rc ? Context::last_error(ctx.handle()) : std::string();
    ndppgpu_sab_free(h);
    if (rc) fatal_error(msg);
}

// apply_tol_scatt(data, tol) (src/scatt.F90:786-818) on data(L, G, NE), in place
inline void apply_tol_scatt(const Context& ctx, std::vector<double>& data, int L, int G, double tol)
{
    const size_t w = (size_t)L * G;
    if (w == 0 || data.size() % w) fatal_error("apply_tol_scatt: array shape");
    ctx.check(ndppgpu_apply_tol(ctx.handle(), data.data(), (int)(data.size() / w), G, L, tol));
}

// thin_grid(xout, yout, tokeep, tol, compression, maxerr [, yout2]) (src/thin.F90:19-47): the arrays are cut to the
// points kept; yout2 may be null
inline void thin_grid(const Context& ctx, std::vector<double>& xout, std::vector<double>& yout,
                      const std::vector<double>& tokeep, double tol, double& compression, double& maxerr,
                      std::vector<double>* yout2 = nullptr)
{
    const int NE = (int)xout.size();
    if (NE == 0 || yout.size() % xout.size()) fatal_error("thin_grid: array shape");
    const int GL = (int)(yout.size() / xout.size());
    if (yout2 && yout2->size() != yout.size()) fatal_error("thin_grid: yout2 shape");
    int kept = 0;
    ctx.check(ndppgpu_thin_grid(ctx.handle(), xout.data(), yout.data(), yout2 ? yout2->data() : nullptr, NE, GL,
                                tokeep.data(), (int)tokeep.size(), tol, &kept, &compression, &maxerr));
    xout.resize(kept);
    yout.resize((size_t)kept * GL);
    if (yout2) yout2->resize((size_t)kept * GL);
}

}  // namespace ndpp_host
