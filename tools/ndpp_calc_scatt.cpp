// ndpp_calc_scatt -- C++ driver of the scattering-moment path: reads a case file (the argument list of
// calc_scatt / calc_scattsab with the nuclide already parsed, written by ndpp_b200/dump.py), calls the C++ host
// layer of include/ndpp_host.hpp exactly where preprocess_ndpp calls the Fortran routines
// (src/ndpp.F90:607-609, 773-775) and writes the moment arrays to a result file.
//
//   ndpp_calc_scatt CASE RESULT [--device N]
//
// Errors end the program the way the reference's fatal_error does (src/error.F90:79-154): " ERROR: <message>"
// on stderr and a non-zero exit status.  All arithmetic runs in libndppgpu.so; there is no CPU path.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ndpp_host.hpp"

using namespace ndpp_host;

namespace {

const double KIND_NUCLIDE = 1.0, KIND_SAB = 2.0;

struct Reader {  // cursor over the stream of little-endian doubles
    std::vector<double> a;
    size_t pos = 0;
    explicit Reader(const char* path)
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) fatal_error(std::string("Cannot open case file ") + path);
        std::fseek(f, 0, SEEK_END);
        const long bytes = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (bytes < 0 || bytes % 8) { std::fclose(f); fatal_error("Case file is not a stream of 8-byte values"); }
        a.resize((size_t)bytes / 8);
        const size_t got = a.empty() ? 0 : std::fread(a.data(), 8, a.size(), f);
        std::fclose(f);
        if (got != a.size()) fatal_error("Short read of the case file");
    }
    double num()
    {
        if (pos >= a.size()) fatal_error("Case file ends early");
        return a[pos++];
    }
    int inum()
    {
        const double v = num();
        if (v != (double)(int)v) fatal_error("Case file: integer field holds a non-integer");
        return (int)v;
    }
    std::vector<double> vec()
    {
        const int n = inum();
        if (n < 0 || pos + (size_t)n > a.size()) fatal_error("Case file: bad vector length");
        std::vector<double> v(a.begin() + pos, a.begin() + pos + n);
        pos += (size_t)n;
        return v;
    }
    std::vector<int> ivec()
    {
        const std::vector<double> v = vec();
        std::vector<int> o(v.size());
        for (size_t i = 0; i < v.size(); ++i) o[i] = (int)v[i];
        return o;
    }
    Tab1 tab1()
    {
        Tab1 t;
        t.present = inum() != 0;
        if (t.present) { t.nbt = ivec(); t.interp = ivec(); t.x = vec(); t.y = vec(); }
        return t;
    }
};

void write_result(const char* path, double kind, int G, int L, const std::vector<double>& el,
                  const std::vector<double>& inel, const std::vector<double>& nu)
{
    FILE* f = std::fopen(path, "wb");
    if (!f) fatal_error(std::string("Cannot open result file ") + path);
    const size_t w = (size_t)G * L;
    const double head[6] = {kind, (double)(el.size() / w), (double)G, (double)L, (double)(inel.size() / w),
                            nu.empty() ? 0.0 : 1.0};
    bool ok = std::fwrite(head, 8, 6, f) == 6;
    for (const std::vector<double>* m : {&el, &inel, &nu})
        if (!m->empty()) ok = ok && std::fwrite(m->data(), 8, m->size(), f) == m->size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) fatal_error(std::string("Cannot write result file ") + path);
}

void run_nuclide(const Context& ctx, Reader& r, const char* out)
{
    Nuclide nuc;
    nuc.awr = r.num(); nuc.kT = r.num(); nuc.freegas_cutoff = r.num();
    nuc.energy = r.vec();
    nuc.elastic = r.vec();
    const int n_rxn = r.inum();
    nuc.reactions.resize(n_rxn);
    for (Reaction& x : nuc.reactions) {
        x.MT = r.inum(); x.Q_value = r.num(); x.multiplicity = r.inum(); x.threshold = r.inum();
        x.scatter_in_cm = r.inum() != 0;
        x.sigma = r.vec();
        x.multiplicity_E = r.tab1();
        x.has_angle_dist = r.inum() != 0;
        if (x.has_angle_dist) {
            x.adist.energy = r.vec(); x.adist.type = r.ivec(); x.adist.location = r.ivec(); x.adist.data = r.vec();
        }
        const int n_ed = r.inum();
        std::unique_ptr<DistEnergy>* tail = &x.edist;
        for (int k = 0; k < n_ed; ++k) {
            tail->reset(new DistEnergy);
            (*tail)->law = r.inum();
            (*tail)->data = r.vec();
            (*tail)->p_valid = r.tab1();
            tail = &(*tail)->next;
        }
    }
    const std::vector<double> energy_bins = r.vec();
    const int scatt_type = r.inum();
    int order = r.inum();
    const int mu_bins = r.inum();
    const bool nuscatt = r.inum() != 0;
    Settings st;
    st.ne_per_grp = r.inum(); st.adaptive_mu_its = r.inum(); st.adaptive_eout_its = r.inum();
    st.sab_threshold = r.num(); st.brent_mu_thresh = r.num(); st.adaptive_mu_tol = r.num(); st.adaptive_eout_tol = r.num();
    const std::vector<double> Ein_el = r.vec(), Ein_inel = r.vec();

    std::vector<double> el_mat, inel_mat, nuinel_mat;
    calc_scatt(ctx, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, Ein_el, Ein_inel, el_mat, inel_mat,
               nuinel_mat, st);
    const int L = (scatt_type == SCATT_TYPE_LEGENDRE) ? order + 1 : order;
    write_result(out, KIND_NUCLIDE, (int)energy_bins.size() - 1, L, el_mat, inel_mat, nuinel_mat);
}

void run_sab(const Context& ctx, Reader& r, const char* out)
{
    SAlphaBeta s;
    s.awr = r.num(); s.kT = r.num(); s.threshold_inelastic = r.num(); s.threshold_elastic = r.num();
    s.n_inelastic_e_in = r.inum(); s.n_inelastic_e_out = r.inum(); s.n_inelastic_mu = r.inum();
    s.secondary_mode = r.inum();
    s.inelastic_e_in = r.vec(); s.inelastic_sigma = r.vec(); s.inelastic_e_out = r.vec(); s.inelastic_mu = r.vec();
    const int n_rows = r.inum();
    s.inelastic_data.resize(n_rows);
    for (DistEnergySab& d : s.inelastic_data) { d.e_out = r.vec(); d.e_out_pdf = r.vec(); d.mu = r.vec(); }
    s.elastic_mode = r.inum(); s.n_elastic_e_in = r.inum(); s.n_elastic_mu = r.inum();
    s.elastic_e_in = r.vec(); s.elastic_P = r.vec(); s.elastic_mu = r.vec();
    const std::vector<double> energy_bins = r.vec();
    const int scatt_type = r.inum(), order = r.inum(), mu_bins = r.inum();
    const std::vector<double> E_grid = r.vec();

    std::vector<double> scatt_mat;
    calc_scattsab(ctx, s, energy_bins, scatt_type, order, scatt_mat, mu_bins, E_grid);
    const int L = (scatt_type == SCATT_TYPE_LEGENDRE) ? order + 1 : order;
    write_result(out, KIND_SAB, (int)energy_bins.size() - 1, L, scatt_mat, {}, {});
}

}  // namespace

int main(int argc, char** argv)
{
    try {
        if (argc < 3) fatal_error("usage: ndpp_calc_scatt CASE RESULT [--device N]");
        int device = -1;
        for (int i = 3; i + 1 < argc; ++i)
            if (!std::strcmp(argv[i], "--device")) device = std::atoi(argv[i + 1]);
        Reader r(argv[1]);
        const double kind = r.num();
        Context ctx(device);
        if (kind == KIND_NUCLIDE) run_nuclide(ctx, r, argv[2]);
        else if (kind == KIND_SAB) run_sab(ctx, r, argv[2]);
        else fatal_error("Case file: unknown kind");
        if (r.pos != r.a.size()) fatal_error("Case file has trailing data");
        const ndppgpu_stats_t s = ctx.stats();
        std::printf(" %lld moment evaluations, %lld kernel launches, %.3f ms on the device\n", s.moment_evals,
                    s.launches, s.kernel_ms);
        return 0;
    } catch (const FatalError& e) {
        std::fprintf(stderr, " ERROR: %s\n", e.what());
        return 255;  // the reference's default error code is -1
    }
}
