// ndpp_calc_scatt -- C++ driver of the scattering-moment path: reads a case file (the argument list of
// calc_scatt / calc_scattsab with the nuclide already parsed, written by ndpp_b200/dump.py), calls the C++ host
// layer of include/ndpp_host.hpp exactly where preprocess_ndpp calls the Fortran routines
// (src/ndpp.F90:607-609, 773-775) and writes the moment arrays to a result file.
//
//   ndpp_calc_scatt CASE RESULT [--device N | --devices N]      --devices: N GPUs of this process (0 = all) work on the
//                                                               nuclide together (ndppgpu_group_*, nuclide cases)
//                   [--library FILE [--ascii] [--name ZAID] [--print-tol P] [--thin-tol T]]   nuclide cases only
//                   [--library-only]
//   --ein-grid (with --library, one device): the E_in grids are built by create_Ein_grid on the device
//   (ndppgpu_nuclide_create_ein_grid) instead of being read from the case file
//
// With --library the program does what the per-nuclide body of preprocess_ndpp does (src/ndpp.F90:560-702):
// calc_scatt, apply_tol_scatt and thin_grid in one device call per matrix set (ndppgpu_*_thinned), then init_library
// and print_scatt (include/ndpp_library.hpp).  --library-only skips the integration and writes the library from the
// matrices of an existing RESULT file on the grids of CASE (no GPU needed: used by the CPU tests of the writer).
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
#include "ndpp_library.hpp"

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

struct Options {
    int device = -1, devices = -1;
    std::string library, name = "synthetic";
    bool ascii = false, library_only = false, ein_grid = false;
    double print_tol = 1.0e-8, thin_tol = 0.0;   // print_tol default of src/constants.F90; thin_tol as a fraction
};

// the matrices of a RESULT file written by write_result
void read_result(const char* path, int G, int L, std::vector<double>& el, std::vector<double>& inel,
                 std::vector<double>& nu)
{
    Reader r(path);
    r.num();
    const int ne_el = r.inum(), g = r.inum(), l = r.inum(), ne_in = r.inum(), has_nu = r.inum();
    if (g != G || l != L) fatal_error("Result file does not match the case (groups / orders)");
    const size_t w = (size_t)G * L;
    auto take = [&](size_t n, std::vector<double>& v) {
        if (r.pos + n > r.a.size()) fatal_error("Result file ends early");
        v.assign(r.a.begin() + r.pos, r.a.begin() + r.pos + n);
        r.pos += n;
    };
    take(ne_el * w, el);
    take(ne_in * w, inel);
    if (has_nu) take(ne_in * w, nu); else nu.clear();
}

void run_nuclide(const Context* ctx, const DeviceGroup* group, Reader& r, const char* out, const Options& opt)
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
    const int L = (scatt_type == SCATT_TYPE_LEGENDRE) ? order + 1 : order;
    const int G = (int)energy_bins.size() - 1;
    std::vector<double> xe = Ein_el, xi = Ein_inel;   // the grids that end up in the library
    if (opt.library_only) {
        read_result(out, G, L, el_mat, inel_mat, nuinel_mat);
        if (el_mat.size() != xe.size() * (size_t)G * L || inel_mat.size() != xi.size() * (size_t)G * L)
            fatal_error("Result file does not match the grids of the case");
    } else if (!opt.library.empty()) {
        // src/ndpp.F90:607-648 with the tolerance and the thinning on the device; tokeep = the group edges (:641-648)
        double compr = 0.0, err = 0.0;
        std::unique_ptr<ScattDataSet> set(group ? new ScattDataSet(*group, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st)
                                                : new ScattDataSet(*ctx, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, st));
        ScattDataSet& rxn_data = *set;
        // create_Ein_grid (src/scatt.F90:139) on the device instead of the grids of the case file
        if (opt.ein_grid) rxn_data.create_Ein_grid(xe, xi);
        rxn_data.calc_elastic_thinned(xe, opt.print_tol, opt.thin_tol, energy_bins, el_mat, compr, err);
        if (!xi.empty())
            rxn_data.calc_inelastic_thinned(xi, nuscatt, opt.print_tol, opt.thin_tol, energy_bins, inel_mat, nuinel_mat,
                                            compr, err);
        rxn_data.clear();
        write_result(out, KIND_NUCLIDE, G, L, el_mat, inel_mat, nuinel_mat);
    } else {
        if (group)
            calc_scatt(*group, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, Ein_el, Ein_inel, el_mat, inel_mat,
                       nuinel_mat, st);
        else
            calc_scatt(*ctx, nuc, energy_bins, scatt_type, order, mu_bins, nuscatt, Ein_el, Ein_inel, el_mat, inel_mat,
                       nuinel_mat, st);
        write_result(out, KIND_NUCLIDE, G, L, el_mat, inel_mat, nuinel_mat);
    }
    if (!opt.library.empty()) {
        if (scatt_type != SCATT_TYPE_LEGENDRE)
            fatal_error("Tabular scattering of ACE nuclides is NOT YET IMPLEMENTED");   // as the reference, scattdata_header.F90:1452-1460
        LibraryWriter w(opt.library, opt.name, nuc.kT, energy_bins, scatt_type, order, nuscatt, mu_bins, opt.thin_tol,
                        opt.ascii ? LibFormat::ASCII : LibFormat::BINARY);
        w.print_scatt(xe, el_mat, xi, inel_mat, nuinel_mat);
        w.close();
        if (!w.ok()) fatal_error("Cannot write library file " + opt.library);
    }
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
        const char* usage = "usage: ndpp_calc_scatt CASE RESULT [--device N | --devices N] [--library FILE [--ascii] [--name ZAID] "
                            "[--print-tol P] [--thin-tol T] [--ein-grid]] [--library-only]";
        if (argc < 3) fatal_error(usage);
        Options opt;
        for (int i = 3; i < argc; ++i) {
            const std::string a = argv[i];
            auto value = [&]() -> const char* { if (i + 1 >= argc) fatal_error(usage); return argv[++i]; };
            if (a == "--device") opt.device = std::atoi(value());
            else if (a == "--devices") opt.devices = std::atoi(value());
            else if (a == "--library") opt.library = value();
            else if (a == "--name") opt.name = value();
            else if (a == "--print-tol") opt.print_tol = std::atof(value());
            else if (a == "--thin-tol") opt.thin_tol = std::atof(value());
            else if (a == "--ascii") opt.ascii = true;
            else if (a == "--library-only") opt.library_only = true;
            else if (a == "--ein-grid") opt.ein_grid = true;
            else fatal_error(usage);
        }
        if (opt.library_only && opt.library.empty()) fatal_error(usage);
        Reader r(argv[1]);
        const double kind = r.num();
        if (opt.library_only) {   // the writer alone: no device context
            if (kind != KIND_NUCLIDE) fatal_error("--library-only needs a nuclide case");
            run_nuclide(nullptr, nullptr, r, argv[2], opt);
            if (r.pos != r.a.size()) fatal_error("Case file has trailing data");
            return 0;
        }
        if (opt.devices >= 0) {   // several GPUs of this process on the one nuclide
            if (kind != KIND_NUCLIDE) fatal_error("--devices is implemented for nuclide cases");
            DeviceGroup group(opt.devices);
            run_nuclide(nullptr, &group, r, argv[2], opt);
            if (r.pos != r.a.size()) fatal_error("Case file has trailing data");
            long long evals = 0, launches = 0;
            double ms = 0.0;
            for (int d = 0; d < group.world(); ++d) {
                const ndppgpu_stats_t s = group.stats(d);
                evals += s.moment_evals; launches += s.launches; ms = s.kernel_ms > ms ? s.kernel_ms : ms;
            }
            std::printf(" %d devices: %lld moment evaluations, %lld kernel launches, %.3f ms on the busiest device, "
                        "%lld bytes gathered over NCCL\n", group.world(), evals, launches, ms,
                        ndppgpu_group_gathered_bytes(group.handle(), 0));
            return 0;
        }
        Context ctx(opt.device);
        if (kind == KIND_NUCLIDE) run_nuclide(&ctx, nullptr, r, argv[2], opt);
        else if (kind == KIND_SAB) {
            if (!opt.library.empty()) fatal_error("--library is implemented for nuclide cases");
            run_sab(ctx, r, argv[2]);
        } else fatal_error("Case file: unknown kind");
        if (r.pos != r.a.size()) fatal_error("Case file has trailing data");
        const ndppgpu_stats_t s = ctx.stats();
        std::printf(" %lld moment evaluations, %lld kernel launches, %.3f ms on the device\n", s.moment_evals,
                    s.launches, s.kernel_ms);
        return 0;
    } catch (const std::runtime_error& e) {   // FatalError and the library writer's errors
        std::fprintf(stderr, " ERROR: %s\n", e.what());
        return 255;  // the reference's default error code is -1
    }
}
