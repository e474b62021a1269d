// convert_main.cpp -- `cuspmm_convert <dir>`: MatrixMarket -> the engine's on-disk text formats.
//
// Native counterpart of the reference's offline converter utils/python_utils/convert_mtx.py
// (process_mtx, :63-295): walks <dir> recursively; `dense.mtx` becomes `dense.in`; every other
// `*.mtx` becomes  <base>.csr  <base>.coo  <base>_colind.ell  <base>_values.ell
// <base>_rowind.ell  <base>_values_colmajor.ell  <base>.bsr .
// The output is BYTE-FOR-BYTE what the Python script writes with scipy 1.x (tests/test_convert.py
// compares against tests/golden/*, which that script produced): numbers are printed the way
// Python's str() prints numpy int64 / float64 scalars, coordinate `symmetric` files are mirrored,
// CSR/CSC have sorted indices, COO is (row, col)-sorted, and the .bsr file reproduces the script's
// two-step tobsr() -> tobsr((1,1)) (scipy's estimated block size first, so the explicit zeros of those
// larger blocks survive as 1x1 blocks, convert_mtx.py:22-24,112).
// Deliberate differences: `--bsr-block B` stores real B x B blocks instead of the forced 1x1
// (convert_mtx.py:22 `size = 1`), and the column-ELL width is max(max row nnz, max column nnz) where the
// script uses the max ROW nnz (convert_mtx.py:252) and crashes when a column is longer.
// Host-only tool (no CUDA): the engine itself converts CSR -> BSR / sliced ELL on the device.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Coo {
    int64_t rows = 0, cols = 0;
    bool isInt = false;                 // "integer" field: values print as Python ints
    std::vector<int64_t> r, c;
    std::vector<double> v;
};

// Python's repr(float) (float_repr_style 'short'): shortest round-trip digits; scientific when the
// decimal exponent is < -4 or >= 16, with at least two exponent digits; otherwise fixed with ".0" added
std::string pyFloat(double x) {
    if (x != x) return "nan";
    if (x == 1.0 / 0.0) return "inf";
    if (x == -1.0 / 0.0) return "-inf";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, res.ptr);                 // [-]d[.ddd]e[+-]XX
    const bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    const size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    const int exp10 = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (char ch : mant)
        if (ch != '.') digits += ch;
    std::string out;
    if (exp10 < -4 || exp10 >= 16) {
        out = digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char e[16];
        std::snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
        out += e;
    } else if (exp10 < 0) {
        out = "0." + std::string(-exp10 - 1, '0') + digits;
    } else if ((size_t)exp10 + 1 >= digits.size()) {
        out = digits + std::string(exp10 + 1 - digits.size(), '0') + ".0";
    } else {
        out = digits.substr(0, exp10 + 1) + "." + digits.substr(exp10 + 1);
    }
    return neg ? "-" + out : out;
}

struct Num {          // a matrix entry as Python would print it
    bool isInt;
    std::string str(double x) const { return isInt ? std::to_string((long long)x) : pyFloat(x); }
};

// scipy.io.mmread for coordinate / array files, real | integer | pattern, general | symmetric | skew-symmetric
bool readMtx(const std::string &path, Coo &m, bool &isArray) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    std::string lower = line;
    std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
    std::istringstream hs(lower);
    std::string banner, object, format, field, symmetry;
    hs >> banner >> object >> format >> field >> symmetry;
    if (banner != "%%matrixmarket" || object != "matrix") return false;
    isArray = format == "array";
    const bool pattern = field == "pattern";
    m.isInt = field == "integer";
    const bool sym = symmetry == "symmetric", skew = symmetry == "skew-symmetric";
    while (std::getline(in, line))
        if (!line.empty() && line[0] != '%' && line.find_first_not_of(" \t\r") != std::string::npos) break;
    std::istringstream ss(line);
    int64_t nnz = 0;
    if (isArray) {
        ss >> m.rows >> m.cols;
        // column-major dense listing (lower triangle only for symmetric)
        for (int64_t j = 0; j < m.cols; ++j)
            for (int64_t i = (sym || skew) ? j : 0; i < m.rows; ++i) {
                std::string tok;
                in >> tok;
                const double val = std::strtod(tok.c_str(), nullptr);
                if (skew && i == j) continue;
                m.r.push_back(i); m.c.push_back(j); m.v.push_back(val);
                if ((sym || skew) && i != j) { m.r.push_back(j); m.c.push_back(i); m.v.push_back(skew ? -val : val); }
            }
        return true;
    }
    ss >> m.rows >> m.cols >> nnz;
    std::vector<int64_t> r2, c2;
    std::vector<double> v2;
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t i, j;
        double val = 1.0;
        in >> i >> j;
        if (!pattern) {
            std::string tok;
            in >> tok;
            val = std::strtod(tok.c_str(), nullptr);
        }
        m.r.push_back(i - 1); m.c.push_back(j - 1); m.v.push_back(val);
        if ((sym || skew) && i != j) { r2.push_back(j - 1); c2.push_back(i - 1); v2.push_back(skew ? -val : val); }
    }
    // scipy appends the mirrored entries after the stored ones
    m.r.insert(m.r.end(), r2.begin(), r2.end());
    m.c.insert(m.c.end(), c2.begin(), c2.end());
    m.v.insert(m.v.end(), v2.begin(), v2.end());
    return true;
}

struct Csr {
    int64_t rows = 0, cols = 0;
    std::vector<int64_t> ptr, idx;
    std::vector<double> val;
};

// coo.tocsr(): row-sorted, column-sorted, duplicates summed (byCol = true gives tocsc())
Csr toCompressed(const Coo &m, bool byCol) {
    const std::vector<int64_t> &major = byCol ? m.c : m.r, &minor = byCol ? m.r : m.c;
    const int64_t nMajor = byCol ? m.cols : m.rows;
    std::vector<size_t> order(m.v.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return major[a] != major[b] ? major[a] < major[b] : minor[a] < minor[b];
    });
    Csr out;
    out.rows = m.rows; out.cols = m.cols;
    out.ptr.assign(nMajor + 1, 0);
    for (size_t k = 0; k < order.size(); ++k) {
        const size_t e = order[k];
        if (k > 0 && major[e] == major[order[k - 1]] && minor[e] == minor[order[k - 1]]) {
            out.val.back() += m.v[e];
            continue;
        }
        out.idx.push_back(minor[e]);
        out.val.push_back(m.v[e]);
        out.ptr[major[e] + 1]++;
    }
    for (int64_t i = 0; i < nMajor; ++i) out.ptr[i + 1] += out.ptr[i];
    return out;
}

template <typename T>
std::string joinInts(const std::vector<T> &v) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) {
        if (i) s += ' ';
        s += std::to_string((long long)v[i]);
    }
    return s;
}

int64_t countBlocks(const Csr &a, int64_t R, int64_t C) {      // scipy.sparse._spfuncs.count_blocks
    const int64_t nbc = a.cols / C + 1;
    std::vector<int64_t> mask(nbc, -1);
    int64_t n = 0;
    for (int64_t i = 0; i < a.rows; ++i) {
        const int64_t bi = i / R;
        for (int64_t k = a.ptr[i]; k < a.ptr[i + 1]; ++k) {
            const int64_t bj = a.idx[k] / C;
            if (mask[bj] != bi) { mask[bj] = bi; ++n; }
        }
    }
    return n;
}

int64_t estimateBlocksize(const Csr &a) {                      // scipy.sparse._spfuncs.estimate_blocksize(A, 0.7)
    const double eff = 0.7, high = (1.0 + eff) / 2.0, nnz = (double)a.idx.size();
    if (a.idx.empty()) return 1;
    const int64_t M = a.rows, N = a.cols;
    const double e22 = (M % 2 == 0 && N % 2 == 0) ? nnz / (4.0 * countBlocks(a, 2, 2)) : 0.0;
    const double e33 = (M % 3 == 0 && N % 3 == 0) ? nnz / (9.0 * countBlocks(a, 3, 3)) : 0.0;
    if (e22 > high && e33 > high) return nnz / (36.0 * countBlocks(a, 6, 6)) > eff ? 6 : 3;
    const double e44 = (M % 4 == 0 && N % 4 == 0) ? nnz / (16.0 * countBlocks(a, 4, 4)) : 0.0;
    if (e44 > eff) return 4;
    if (e33 > eff) return 3;
    if (e22 > eff) return 2;
    return 1;
}

struct Bsr {
    int64_t bs = 1;
    std::vector<int64_t> ptr, idx;
    std::vector<double> data;      // [numBlocks][bs*bs]
};

// csr.tobsr((bs, bs)): block columns of a block row in FIRST-TOUCH order (sparsetools csr_tobsr)
Bsr csrToBsr(const Csr &a, int64_t bs, bool sortCols) {
    Bsr b;
    b.bs = bs;
    const int64_t nbr = a.rows / bs, nbc = a.cols / bs;
    b.ptr.assign(nbr + 1, 0);
    std::vector<int64_t> slot(nbc, -1);
    for (int64_t R = 0; R < nbr; ++R) {
        const size_t first = b.idx.size();
        for (int64_t r = 0; r < bs; ++r) {
            const int64_t i = R * bs + r;
            for (int64_t k = a.ptr[i]; k < a.ptr[i + 1]; ++k) {
                const int64_t bj = a.idx[k] / bs, c = a.idx[k] % bs;
                if (slot[bj] < (int64_t)first) {
                    slot[bj] = (int64_t)b.idx.size();
                    b.idx.push_back(bj);
                    b.data.resize(b.data.size() + bs * bs, 0.0);
                }
                b.data[slot[bj] * bs * bs + r * bs + c] += a.val[k];
            }
        }
        if (sortCols) {     // canonical order for real blocks
            const size_t n = b.idx.size() - first;
            std::vector<size_t> ord(n);
            std::iota(ord.begin(), ord.end(), 0);
            std::sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return b.idx[first + x] < b.idx[first + y]; });
            std::vector<int64_t> idx2(n);
            std::vector<double> data2(n * bs * bs);
            for (size_t t = 0; t < n; ++t) {
                idx2[t] = b.idx[first + ord[t]];
                std::copy_n(b.data.begin() + (first + ord[t]) * bs * bs, bs * bs, data2.begin() + t * bs * bs);
            }
            std::copy(idx2.begin(), idx2.end(), b.idx.begin() + first);
            std::copy(data2.begin(), data2.end(), b.data.begin() + first * bs * bs);
        }
        for (size_t t = first; t < b.idx.size(); ++t) slot[b.idx[t]] = -1;
        b.ptr[R + 1] = (int64_t)b.idx.size();
    }
    return b;
}

// bsr.tocsr() keeps every stored block element (explicit zeros included), blocks in stored order
Csr bsrToCsrKeepZeros(const Bsr &b, int64_t rows, int64_t cols) {
    Csr a;
    a.rows = rows; a.cols = cols;
    a.ptr.assign(rows + 1, 0);
    const int64_t bs = b.bs;
    for (int64_t R = 0; R + 1 < (int64_t)b.ptr.size(); ++R)
        for (int64_t r = 0; r < bs; ++r) {
            for (int64_t t = b.ptr[R]; t < b.ptr[R + 1]; ++t)
                for (int64_t c = 0; c < bs; ++c) {
                    a.idx.push_back(b.idx[t] * bs + c);
                    a.val.push_back(b.data[t * bs * bs + r * bs + c]);
                }
            a.ptr[R * bs + r + 1] = (int64_t)a.idx.size();
        }
    return a;
}

void writeSparse(const std::string &dir, const std::string &base, const Coo &m, int64_t bsrBlock) {
    const Num num{m.isInt};
    const Csr csr = toCompressed(m, false), csc = toCompressed(m, true);
    const int64_t nnz = (int64_t)csr.idx.size();
    auto path = [&](const std::string &suffix) { return dir + "/" + base + suffix; };
    auto joinVals = [&](const double *p, size_t n) {
        std::string s;
        for (size_t i = 0; i < n; ++i) { if (i) s += ' '; s += num.str(p[i]); }
        return s;
    };
    {   // .csr (convert_mtx.py:127-143)
        std::ofstream f(path(".csr"));
        f << csr.rows << ' ' << csr.cols << ' ' << nnz << '\n' << joinInts(csr.ptr) << '\n' << joinInts(csr.idx) << '\n'
          << joinVals(csr.val.data(), csr.val.size()) << '\n';
        std::cout << "Saved CSR format to " << path(".csr") << "\n";
    }
    {   // .coo: the COO entries (duplicates NOT summed) lexsorted by (row, col) (convert_mtx.py:173-188)
        std::vector<size_t> order(m.v.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return m.r[a] != m.r[b] ? m.r[a] < m.r[b] : m.c[a] < m.c[b]; });
        std::ofstream f(path(".coo"));
        f << m.rows << ' ' << m.cols << ' ' << m.v.size() << '\n';
        for (size_t e : order) f << m.r[e] << ' ' << m.c[e] << ' ' << num.str(m.v[e]) << '\n';
        std::cout << "Saved COO format to " << path(".coo") << "\n";
    }
    int64_t maxRow = 0, maxCol = 0;
    for (int64_t i = 0; i < csr.rows; ++i) maxRow = std::max(maxRow, csr.ptr[i + 1] - csr.ptr[i]);
    for (int64_t j = 0; j < csc.cols; ++j) maxCol = std::max(maxCol, csc.ptr[j + 1] - csc.ptr[j]);
    auto writeEll = [&](const Csr &a, int64_t lines, int64_t width, const std::string &idxFile, const std::string &valFile) {
        std::ofstream fi(idxFile), fv(valFile);
        fi << m.rows << ' ' << m.cols << ' ' << nnz << ' ' << width << '\n';
        for (int64_t i = 0; i < lines; ++i) {
            std::string si, sv;
            for (int64_t s = 0; s < width; ++s) {
                if (s) { si += ' '; sv += ' '; }
                const int64_t k = a.ptr[i] + s;
                if (k < a.ptr[i + 1]) { si += std::to_string((long long)a.idx[k]); sv += num.str(a.val[k]); }
                else { si += "-1"; sv += "0"; }      // padding: -1 / the int literal 0 (convert_mtx.py:207-208)
            }
            fi << si << '\n';
            fv << sv << '\n';
        }
    };
    writeEll(csr, csr.rows, maxRow, path("_colind.ell"), path("_values.ell"));                       // :198-239
    std::cout << "Saved ELL colind to " << path("_colind.ell") << "\n";
    writeEll(csc, csc.cols, std::max(maxRow, maxCol), path("_rowind.ell"), path("_values_colmajor.ell"));   // :245-286
    std::cout << "Saved ELL rowind to " << path("_rowind.ell") << "\n";
    {   // .bsr (save_bsr_matrix, convert_mtx.py:7-61)
        Bsr out;
        if (bsrBlock > 1 && csr.rows % bsrBlock == 0 && csr.cols % bsrBlock == 0) {
            out = csrToBsr(csr, bsrBlock, /*sortCols=*/true);
        } else {
            const int64_t est = estimateBlocksize(csr);                       // matrix.tobsr()  (:112)
            const Bsr first = csrToBsr(csr, est, false);
            out = csrToBsr(est == 1 ? csr : bsrToCsrKeepZeros(first, csr.rows, csr.cols), 1, false);   // .tobsr((1,1)) (:24)
        }
        std::ofstream f(path(".bsr"));
        f << csr.rows << ' ' << csr.cols << ' ' << out.data.size() << ' ' << out.bs << ' ' << out.bs << ' ' << out.idx.size() << '\n'
          << joinInts(out.ptr) << '\n' << joinInts(out.idx) << '\n';
        const size_t per = (size_t)out.bs * out.bs;
        for (size_t t = 0; t < out.idx.size(); ++t) f << joinVals(out.data.data() + t * per, per) << '\n';
        std::cout << "bsr using shape " << out.bs << "," << out.bs << "\nSaved BSR values to " << path(".bsr") << "\n";
    }
}

void writeDense(const std::string &dir, const Coo &m) {      // convert_mtx.py:66-95
    const Num num{m.isInt};
    std::vector<double> D((size_t)m.rows * m.cols, 0.0);
    for (size_t e = 0; e < m.v.size(); ++e) D[(size_t)m.r[e] * m.cols + m.c[e]] += m.v[e];     // todense() sums duplicates
    int64_t nz = 0;
    for (double x : D) nz += x != 0.0;
    std::ofstream f(dir + "/dense.in");
    f << m.rows << ' ' << m.cols << ' ' << nz << '\n';
    for (int64_t i = 0; i < m.rows; ++i) {
        std::string s;
        for (int64_t j = 0; j < m.cols; ++j) { if (j) s += ' '; s += num.str(D[(size_t)i * m.cols + j]); }
        f << s << '\n';
    }
    std::cout << "Processed " << dir << "/dense.mtx -> " << dir << "/dense.in\n";
}

void walk(const std::string &dir, int64_t bsrBlock) {
    DIR *d = opendir(dir.c_str());
    if (!d) { std::cerr << "cannot open " << dir << "\n"; return; }
    std::vector<std::string> names;
    while (dirent *e = readdir(d)) names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string &n : names) {
        if (n == "." || n == "..") continue;
        const std::string p = dir + "/" + n;
        struct stat st;
        if (stat(p.c_str(), &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) { walk(p, bsrBlock); continue; }
        if (n.size() < 4 || n.compare(n.size() - 4, 4, ".mtx") != 0) continue;
        Coo m;
        bool isArray = false;
        if (!readMtx(p, m, isArray)) { std::cerr << "Error reading " << p << "\n"; continue; }
        if (n == "dense.mtx") writeDense(dir, m);
        else {
            std::cout << "Processing " << p << "...\n";
            writeSparse(dir, n.substr(0, n.size() - 4), m, bsrBlock);
            std::cout << "finish all\n";
        }
    }
}

}  // namespace

int main(int argc, char **argv) {
    int64_t bsrBlock = 1;
    std::string dir;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--bsr-block") && i + 1 < argc) bsrBlock = std::atoll(argv[++i]);
        else dir = argv[i];
    }
    if (dir.empty()) {
        std::cout << "Usage: " << argv[0] << " <directory> [--bsr-block B]\n";     // convert_mtx.py:298-301
        return 1;
    }
    walk(dir, bsrBlock);
    return 0;
}
