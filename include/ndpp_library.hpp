// ndpp_library.hpp -- the NDPP library file of one nuclide, written from the moment arrays (C++ twin of
// ndpp_b200/output.py).  Header-only, no dependency on the CUDA library.
//
//   group_index    <- src/ndpp.F90:649-683    energy-group locations in an E_in grid (binary_search, 1-based)
//   init_library   <- src/ndpp.F90:1246-1329  header: name(10), kT, NG, E_bins, scatt_type, scatt_order, nuscatter,
//                                             chi_present, mu_bins, thin_tol
//   print_scatt    <- src/scatt.F90:821-997 (ASCII), :1139-1258 (BINARY, Fortran stream access): per matrix NE, Ein(:),
//                     grp_index(NG+1), and per E_in gmin, gmax (1-based; 0, 0 for an all-zero column) followed by the
//                     L moments of every group in the window of positive P0
//
// Matrices are Fortran mat(L, G, NE) == C mat[iE][g][l], as the C-ABI returns them.  The byte stream is the one the
// reference's reader src/utils/ndpp_data.py:141-245 consumes; tests/test_host_cpp.py compares the files with those of
// the Python writer byte for byte.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ndpp_host {

enum class LibFormat { ASCII, BINARY };

// binary_search of src/search.F90:21-71 (1-based lower index; val == last -> n-1); the callers guard the range
inline int binary_search_1based(const std::vector<double>& a, double val)
{
    const int n = (int)a.size();
    if (n < 2 || val < a[0] || val > a[n - 1]) throw std::runtime_error("Value outside of array during binary search");
    int L = 1, R = n;  // 1-based
    while (R - L > 1) {
        const int mid = L + (R - L) / 2;
        if (val >= a[mid - 1]) L = mid; else R = mid;
    }
    return L;
}

// group_index_* of src/ndpp.F90:649-683
inline std::vector<int> group_index(const std::vector<double>& Ein, const std::vector<double>& energy_bins)
{
    const int n = (int)Ein.size();
    std::vector<int> out(energy_bins.size(), 0);
    for (size_t g = 0; g < energy_bins.size(); ++g) {
        const double e = energy_bins[g];
        if (e < Ein.front()) out[g] = 1;
        else if (e >= Ein.back()) out[g] = n;
        else out[g] = binary_search_1based(Ein, e);
    }
    if (!out.empty()) out.back() = n;
    return out;
}

class LibraryWriter {
public:
    LibraryWriter(const std::string& filename, const std::string& name, double kT, const std::vector<double>& energy_bins,
                  int scatt_type, int scatt_order, bool nuscatter, int mu_bins, double thin_tol, LibFormat fmt,
                  bool chi_present = false)
        : fmt_(fmt), eb_(energy_bins), NG_((int)energy_bins.size() - 1),
          L_(scatt_type == 0 ? scatt_order + 1 : scatt_order), nuscatter_(nuscatter)
    {
        if (NG_ < 1) throw std::runtime_error("init_library: energy_bins needs at least two edges");
        f_ = std::fopen(filename.c_str(), fmt == LibFormat::BINARY ? "wb" : "w");
        if (!f_) throw std::runtime_error("Cannot open library file " + filename);
        std::string name10 = (name + std::string(10, ' ')).substr(0, 10);
        const int hdr[4] = {scatt_type, scatt_order, nuscatter ? 1 : 0, chi_present ? 1 : 0};
        if (fmt_ == LibFormat::BINARY) {
            put(name10.data(), 10);
            put(&kT, 8);
            put(&NG_, 4);
            put(eb_.data(), 8 * eb_.size());
            put(hdr, 16);
            put(&mu_bins, 4);
            put(&thin_tol, 8);
        } else {
            // '(A20,1PE20.12,I20,A20)': a character(10) name in an A20 field is right-justified
            line(rstrip(std::string(10, ' ') + name10 + fortran_e(kT) + int20(NG_)));
            ascii_reals(eb_.data(), eb_.size());
            line(int20(hdr[0]) + int20(hdr[1]) + int20(hdr[2]) + int20(hdr[3]));
            line(int20(mu_bins) + fortran_e(thin_tol));
        }
    }
    ~LibraryWriter() { close(); }
    LibraryWriter(const LibraryWriter&) = delete;
    LibraryWriter& operator=(const LibraryWriter&) = delete;

    // print_scatt: the elastic grid and matrix, then the inelastic ones (or a zero count), then nu-inelastic
    void print_scatt(const std::vector<double>& Ein_el, const std::vector<double>& el_mat,
                     const std::vector<double>& Ein_inel, const std::vector<double>& inel_mat,
                     const std::vector<double>& nuinel_mat)
    {
        grid(Ein_el);
        matrix(el_mat, Ein_el.size());
        if (!Ein_inel.empty()) {
            grid(Ein_inel);
            matrix(inel_mat, Ein_inel.size());
            if (nuscatter_) {
                if (nuinel_mat.empty()) throw std::runtime_error("nuscatter is set but no nu-inelastic matrix was given");
                matrix(nuinel_mat, Ein_inel.size());
            }
        } else if (fmt_ == LibFormat::BINARY) {
            const int zero = 0;
            put(&zero, 4);
        } else {
            line(int20(0));
        }
    }
    void close()
    {
        if (f_) {
            const bool bad = std::ferror(f_) != 0;
            const bool bad2 = std::fclose(f_) != 0;
            f_ = nullptr;
            if (bad || bad2) ok_ = false;
        }
    }
    bool ok() const { return ok_; }

    // Fortran edit descriptor 1PE20.12 (a three-digit exponent drops the letter: 1.000000000000+100)
    static std::string fortran_e(double v)
    {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%20.12E", v);
        std::string s(buf);
        const size_t e = s.find('E');
        if (e != std::string::npos && s.size() - e - 1 > 3) {
            std::string mant = s.substr(0, e), ex = s.substr(e + 1);  // sign + digits
            size_t b = mant.find_first_not_of(' ');
            mant = mant.substr(b == std::string::npos ? 0 : b);
            std::string digits = ex.substr(1);
            while (digits.size() < 3) digits = "0" + digits;
            s = mant + ex[0] + digits;
            if (s.size() < 20) s = std::string(20 - s.size(), ' ') + s;
        }
        return s;
    }

private:
    void put(const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f_) != n) ok_ = false; }
    void line(const std::string& s) { put(s.data(), s.size()); put("\n", 1); }
    static std::string int20(long long v)
    {
        char buf[32];
        std::snprintf(buf, sizeof buf, "%20lld", v);
        return buf;
    }
    static std::string rstrip(std::string s)
    {
        while (!s.empty() && s.back() == ' ') s.pop_back();
        return s;
    }
    void ascii_reals(const double* a, size_t n)
    {
        for (size_t i = 0; i < n; i += 4) {
            std::string s;
            for (size_t j = i; j < n && j < i + 4; ++j) s += fortran_e(a[j]);
            line(rstrip(s));
        }
    }
    void ascii_ints(const int* a, size_t n)
    {
        for (size_t i = 0; i < n; i += 4) {
            std::string s;
            for (size_t j = i; j < n && j < i + 4; ++j) s += int20(a[j]);
            line(rstrip(s));
        }
    }
    void grid(const std::vector<double>& Ein)
    {
        const std::vector<int> gi = group_index(Ein, eb_);
        const int NE = (int)Ein.size();
        if (fmt_ == LibFormat::BINARY) {
            put(&NE, 4);
            put(Ein.data(), 8 * Ein.size());
            put(gi.data(), 4 * gi.size());
        } else {
            line(int20(NE));
            ascii_reals(Ein.data(), Ein.size());
            ascii_ints(gi.data(), gi.size());
        }
    }
    // gmin, gmax from the P0 moments (src/scatt.F90:923-931) and the window's moments
    void matrix(const std::vector<double>& mat, size_t NE)
    {
        const size_t w = (size_t)NG_ * L_;
        if (mat.size() != NE * w) throw std::runtime_error("print_scatt: matrix shape");
        for (size_t iE = 0; iE < NE; ++iE) {
            const double* col = mat.data() + iE * w;
            int lo = 0, hi = 0;
            for (int g = 0; g < NG_; ++g)
                if (col[(size_t)g * L_] > 0.0) { if (!lo) lo = g + 1; hi = g + 1; }
            const double* first = col + (size_t)(lo > 0 ? lo - 1 : 0) * L_;
            const size_t n = lo > 0 ? (size_t)(hi - lo + 1) * L_ : 0;
            if (fmt_ == LibFormat::BINARY) {
                const int lh[2] = {lo, hi};
                put(lh, 8);
                put(first, 8 * n);
            } else {
                line(int20(lo) + int20(hi));
                ascii_reals(first, n);
            }
        }
    }

    LibFormat fmt_;
    std::vector<double> eb_;
    int NG_, L_;
    bool nuscatter_, ok_ = true;
    std::FILE* f_ = nullptr;
};

}  // namespace ndpp_host
