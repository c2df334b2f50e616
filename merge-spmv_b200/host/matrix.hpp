// matrix.hpp -- host-side matrix containers, Matrix-Market IO, generators and statistics for the
// gpu_spmv / cpu_spmv drivers.
//
// Replaces the input side of the reference's driver surface (file:line in /root/reference):
//   CooMatrix::InitMarket      sparse_matrix.h:217-380   (quirks kept, see read_matrix_market)
//   CooMatrix::InitDense/Wheel/Grid2d/Grid3d   sparse_matrix.h:386-617
//   CsrMatrix::Init (COO->CSR) sparse_matrix.h:666-728   (stable sort by (row, col), duplicates kept)
//   CsrMatrix::Stats / DisplayHistogram / GraphStats::Display   sparse_matrix.h:59-107,786-956
// plus the builder-defined uniform / power-law / banded families of BASELINE.json
// (same formulas as merge-spmv_b200/generators.py).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <string>
#include <vector>

namespace mspmv_host {

struct CooEntry {
    int row, col;
    double val;
};

struct Coo {
    int num_rows = 0, num_cols = 0;
    std::vector<CooEntry> entries;
};

template <typename V>
struct Csr {
    int num_rows = 0, num_cols = 0, num_nonzeros = 0;
    std::vector<int> row_offsets;
    std::vector<int> column_indices;
    std::vector<V> values;
};

// COO -> CSR exactly as CsrMatrix::Init (sparse_matrix.h:666-728): stable sort by (row, col) so
// duplicates stay separate entries in file order; row_offsets[r] = first entry with row >= r.
template <typename V>
Csr<V> coo_to_csr(Coo& coo)
{
    std::stable_sort(coo.entries.begin(), coo.entries.end(), [](const CooEntry& a, const CooEntry& b) {
        return a.row < b.row || (a.row == b.row && a.col < b.col);
    });
    Csr<V> m;
    m.num_rows = coo.num_rows;
    m.num_cols = coo.num_cols;
    m.num_nonzeros = (int)coo.entries.size();
    m.row_offsets.assign(m.num_rows + 1, 0);
    m.column_indices.resize(m.num_nonzeros);
    m.values.resize(m.num_nonzeros);
    int prev_row = -1;
    for (int k = 0; k < m.num_nonzeros; ++k) {
        const CooEntry& e = coo.entries[k];
        for (int r = prev_row + 1; r <= e.row; ++r) m.row_offsets[r] = k;
        prev_row = e.row;
        m.column_indices[k] = e.col;
        m.values[k] = (V)e.val;
    }
    for (int r = prev_row + 1; r <= m.num_rows; ++r) m.row_offsets[r] = m.num_nonzeros;
    return m;
}

// ---------------------------------------------------------------------------------------------
// Matrix-Market reader with the reference's behaviour (sparse_matrix.h:217-380):
//   * lines longer than 1023 chars end parsing silently; a line starting "%%" is the banner, flags
//     by substring: "symmetric" (also matches skew-symmetric), "skew", "array";
//   * hermitian/complex/pattern/integer are not recognised: pattern entries get default_value,
//     complex keeps the real part;
//   * coordinate entries are 1-based; symmetric off-diagonal entries are mirrored right after the
//     entry (negated when skew); explicit zeros and duplicates are kept; array is column-major.
// ---------------------------------------------------------------------------------------------
inline Coo read_matrix_market(const std::string& path, double default_value = 1.0, bool verbose = false)
{
    if (verbose) {
        std::printf("Reading... ");
        std::fflush(stdout);
    }
    std::ifstream ifs(path.c_str());
    if (!ifs.good()) {
        std::fprintf(stderr, "Error opening file\n");
        std::exit(1);
    }
    Coo coo;
    bool array = false, symmetric = false, skew = false;
    long long declared = -1, current = -1;
    char line[1024];
    if (verbose) {
        std::printf("Parsing... ");
        std::fflush(stdout);
    }
    while (true) {
        ifs.getline(line, 1024);
        if (!ifs.good()) break;
        if (line[0] == '%') {
            if (line[1] == '%') {
                symmetric = std::strstr(line, "symmetric") != nullptr;
                skew = std::strstr(line, "skew") != nullptr;
                array = std::strstr(line, "array") != nullptr;
                if (verbose) {
                    std::printf("(symmetric: %d, skew: %d, array: %d) ", symmetric, skew, array);
                    std::fflush(stdout);
                }
            }
        } else if (current == -1) {
            int r = 0, c = 0, n = 0;
            int parsed = std::sscanf(line, "%d %d %d", &r, &c, &n);
            if (!array && parsed == 3) {
                declared = symmetric ? 2LL * n : n;
            } else if (array && parsed == 2) {
                declared = (long long)r * c;
            } else {
                std::fprintf(stderr, "Error parsing MARKET matrix: invalid problem description: %s\n", line);
                std::exit(1);
            }
            coo.num_rows = r;
            coo.num_cols = c;
            coo.entries.reserve((size_t)declared);
            current = 0;
        } else {
            if (current >= declared) {
                std::fprintf(stderr, "Error parsing MARKET matrix: encountered more than %lld num_nonzeros\n", declared);
                std::exit(1);
            }
            int row, col;
            double val;
            if (array) {
                if (std::sscanf(line, "%lf", &val) != 1) {
                    std::fprintf(stderr, "Error parsing MARKET matrix: badly formed current_nz: '%s' at edge %lld\n", line, current);
                    std::exit(1);
                }
                col = (int)(current / coo.num_rows);
                row = (int)(current - (long long)coo.num_rows * col);
                coo.entries.push_back({row, col, val});
            } else {
                char* l = line;
                char* t = nullptr;
                row = (int)std::strtol(l, &t, 0);
                if (t == l) {
                    std::fprintf(stderr, "Error parsing MARKET matrix: badly formed row at edge %lld\n", current);
                    std::exit(1);
                }
                l = t;
                col = (int)std::strtol(l, &t, 0);
                if (t == l) {
                    std::fprintf(stderr, "Error parsing MARKET matrix: badly formed col at edge %lld\n", current);
                    std::exit(1);
                }
                l = t;
                val = std::strtod(l, &t);
                if (t == l) val = default_value;
                coo.entries.push_back({row - 1, col - 1, val});
            }
            ++current;
            if (symmetric && row != col) {  // the reference compares the 1-based (or array) indices it parsed
                const CooEntry e = coo.entries.back();
                coo.entries.push_back({e.col, e.row, e.val * (skew ? -1 : 1)});
                ++current;
            }
        }
    }
    if (verbose) {
        std::printf("done. ");
        std::fflush(stdout);
    }
    return coo;
}

template <typename V>
void write_matrix_market(const std::string& path, const Csr<V>& m)
{
    std::FILE* f = std::fopen(path.c_str(), "w");
    if (!f) {
        std::fprintf(stderr, "cannot write %s\n", path.c_str());
        std::exit(1);
    }
    std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n", m.num_rows, m.num_cols, m.num_nonzeros);
    for (int r = 0; r < m.num_rows; ++r)
        for (int k = m.row_offsets[r]; k < m.row_offsets[r + 1]; ++k)
            std::fprintf(f, "%d %d %.17g\n", r + 1, m.column_indices[k] + 1, (double)m.values[k]);
    std::fclose(f);
}

// ---------------------------------------------------------------------------------------------
// The reference's generators (all values = 1.0)
// ---------------------------------------------------------------------------------------------
inline Coo gen_dense(int rows, int cols)  // sparse_matrix.h:386-413
{
    Coo coo;
    coo.num_rows = rows;
    coo.num_cols = cols;
    coo.entries.reserve((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) coo.entries.push_back({r, c, 1.0});
    return coo;
}

inline Coo gen_wheel(int spokes)  // sparse_matrix.h:419-452
{
    Coo coo;
    coo.num_rows = coo.num_cols = spokes + 1;
    for (int i = 0; i < spokes; ++i) coo.entries.push_back({0, i + 1, 1.0});
    for (int i = 0; i < spokes; ++i) coo.entries.push_back({i + 1, (i + 1) % spokes + 1, 1.0});
    return coo;
}

inline Coo gen_grid2d(int width)  // sparse_matrix.h:461-526, self_loop = false; W, E, N, S
{
    Coo coo;
    coo.num_rows = coo.num_cols = width * width;
    for (int j = 0; j < width; ++j)
        for (int k = 0; k < width; ++k) {
            int me = j * width + k;
            if (k - 1 >= 0) coo.entries.push_back({me, me - 1, 1.0});
            if (k + 1 < width) coo.entries.push_back({me, me + 1, 1.0});
            if (j - 1 >= 0) coo.entries.push_back({me, me - width, 1.0});
            if (j + 1 < width) coo.entries.push_back({me, me + width, 1.0});
        }
    return coo;
}

inline Coo gen_grid3d(int width)  // sparse_matrix.h:533-617, self_loop = false
{
    Coo coo;
    coo.num_rows = coo.num_cols = width * width * width;
    const int w2 = width * width;
    for (int i = 0; i < width; ++i)
        for (int j = 0; j < width; ++j)
            for (int k = 0; k < width; ++k) {
                int me = i * w2 + j * width + k;
                if (k - 1 >= 0) coo.entries.push_back({me, me - 1, 1.0});
                if (k + 1 < width) coo.entries.push_back({me, me + 1, 1.0});
                if (j - 1 >= 0) coo.entries.push_back({me, me - width, 1.0});
                if (j + 1 < width) coo.entries.push_back({me, me + width, 1.0});
                if (i - 1 >= 0) coo.entries.push_back({me, me - w2, 1.0});
                if (i + 1 < width) coo.entries.push_back({me, me + w2, 1.0});
            }
    return coo;
}

// ---------------------------------------------------------------------------------------------
// BASELINE.json families (builder-defined; formulas shared with generators.py)
// ---------------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double hashed_value(uint64_t index, uint64_t seed)  // U[0.5, 1.5)
{
    uint64_t h = splitmix64(index + seed * 0x100000001B3ull);
    return 0.5 + (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

template <typename V>
void fill_stratified(Csr<V>& m, uint64_t seed, bool random_values)
{
    m.column_indices.resize(m.num_nonzeros);
    m.values.resize(m.num_nonzeros);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int r = 0; r < m.num_rows; ++r) {
        const int64_t start = m.row_offsets[r], len = m.row_offsets[r + 1] - start;
        for (int64_t j = 0; j < len; ++j) {
            const int64_t k = start + j;
            const int64_t lo = j * m.num_cols / len, hi = (j + 1) * m.num_cols / len;
            const uint64_t h = splitmix64((uint64_t)k + seed) >> 1;
            m.column_indices[k] = (int)(lo + (int64_t)(h % (uint64_t)std::max<int64_t>(hi - lo, 1)));
            m.values[k] = random_values ? (V)hashed_value((uint64_t)k, seed ^ 0xABCDEFull) : (V)1;
        }
    }
}

template <typename V>
Csr<V> gen_uniform(int rows, int cols, int nnz_per_row, bool random_values, uint64_t seed = 0x5EED0001ull)
{
    Csr<V> m;
    m.num_rows = rows;
    m.num_cols = cols;
    m.num_nonzeros = rows * nnz_per_row;
    m.row_offsets.resize(rows + 1);
    for (int r = 0; r <= rows; ++r) m.row_offsets[r] = r * nnz_per_row;
    fill_stratified(m, seed, random_values);
    return m;
}

template <typename V>
Csr<V> gen_powerlaw(int rows, int cols, int max_row, int64_t target_nnz, bool random_values,
                    uint64_t seed = 0x5EED0003ull, double* alpha_out = nullptr)
{
    max_row = std::min(max_row, cols);
    auto total = [&](double alpha) {
        int64_t s = 0;
        for (int k = 1; k <= rows; ++k) s += std::max<int64_t>(1, (int64_t)std::floor(max_row / std::pow((double)k, alpha)));
        return s;
    };
    double lo = 0.0, hi = 4.0;
    for (int it = 0; it < 60; ++it) {
        double mid = 0.5 * (lo + hi);
        if (total(mid) > target_nnz) lo = mid;
        else hi = mid;
    }
    const double alpha = hi;
    if (alpha_out) *alpha_out = alpha;
    std::vector<int> perm(rows);
    std::iota(perm.begin(), perm.end(), 0);
    std::vector<int64_t> key(rows);
    for (int i = 0; i < rows; ++i) key[i] = (int64_t)splitmix64((uint64_t)i + seed);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<int64_t> lengths(rows);
    for (int k = 0; k < rows; ++k)  // rank k lands on row perm[k]
        lengths[perm[k]] = std::max<int64_t>(1, (int64_t)std::floor(max_row / std::pow((double)(k + 1), alpha)));
    Csr<V> m;
    m.num_rows = rows;
    m.num_cols = cols;
    m.row_offsets.resize(rows + 1);
    int64_t acc = 0;
    for (int r = 0; r < rows; ++r) {
        m.row_offsets[r] = (int)acc;
        acc += lengths[r];
    }
    if (acc + rows >= (int64_t)INT32_MAX - 65536) {
        std::fprintf(stderr, "rows + nnz must stay below 2^31\n");
        std::exit(1);
    }
    m.row_offsets[rows] = (int)acc;
    m.num_nonzeros = (int)acc;
    fill_stratified(m, seed, random_values);
    return m;
}

template <typename V>
Csr<V> gen_banded(int rows, int half_bandwidth, bool random_values, uint64_t seed = 0x5EED0004ull)
{
    Csr<V> m;
    m.num_rows = m.num_cols = rows;
    m.row_offsets.resize(rows + 1);
    int64_t acc = 0;
    for (int r = 0; r < rows; ++r) {
        m.row_offsets[r] = (int)acc;
        acc += std::min(r + half_bandwidth, rows - 1) - std::max(r - half_bandwidth, 0) + 1;
    }
    m.row_offsets[rows] = (int)acc;
    m.num_nonzeros = (int)acc;
    m.column_indices.resize(acc);
    m.values.resize(acc);
#pragma omp parallel for
    for (int r = 0; r < rows; ++r) {
        int c0 = std::max(r - half_bandwidth, 0);
        for (int k = m.row_offsets[r]; k < m.row_offsets[r + 1]; ++k) {
            m.column_indices[k] = c0 + (k - m.row_offsets[r]);
            m.values[k] = random_values ? (V)hashed_value((uint64_t)k, seed ^ 0xABCDEFull) : (V)1;
        }
    }
    return m;
}

// Binary dump for tests: int32 rows, cols, nnz; int32 row_offsets[rows+1]; int32 col[nnz]; float64 val[nnz]
template <typename V>
void dump_csr(const std::string& path, const Csr<V>& m)
{
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        std::fprintf(stderr, "cannot write %s\n", path.c_str());
        std::exit(1);
    }
    int hdr[3] = {m.num_rows, m.num_cols, m.num_nonzeros};
    std::fwrite(hdr, sizeof(int), 3, f);
    std::fwrite(m.row_offsets.data(), sizeof(int), m.row_offsets.size(), f);
    std::fwrite(m.column_indices.data(), sizeof(int), m.column_indices.size(), f);
    std::vector<double> v(m.values.begin(), m.values.end());
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
}

// ---------------------------------------------------------------------------------------------
// Statistics and their print-out (sparse_matrix.h:59-107, 786-956)
// ---------------------------------------------------------------------------------------------
struct GraphStats {
    int num_rows, num_cols, num_nonzeros;
    double row_length_mean, row_length_std_dev, row_length_variation, row_length_skewness;

    void display(bool show_labels) const
    {
        if (show_labels)
            std::printf("\n\t num_rows: %d\n\t num_cols: %d\n\t num_nonzeros: %d\n\t row_length_mean: %.5f\n"
                        "\t row_length_std_dev: %.5f\n\t row_length_variation: %.5f\n\t row_length_skewness: %.5f\n",
                        num_rows, num_cols, num_nonzeros, row_length_mean, row_length_std_dev,
                        row_length_variation, row_length_skewness);
        else
            std::printf("%d, %d, %d, %.5f, %.5f, %.5f, %.5f, ", num_rows, num_cols, num_nonzeros, row_length_mean,
                        row_length_std_dev, row_length_variation, row_length_skewness);
    }
};

template <typename V>
GraphStats stats(const Csr<V>& m)
{
    GraphStats s;
    s.num_rows = m.num_rows;
    s.num_cols = m.num_cols;
    s.num_nonzeros = m.num_nonzeros;
    s.row_length_mean = double(m.num_nonzeros) / m.num_rows;
    double variance = 0.0, skew = 0.0;
    for (int r = 0; r < m.num_rows; ++r) {
        double delta = double(m.row_offsets[r + 1] - m.row_offsets[r]) - s.row_length_mean;
        variance += delta * delta;
        skew += delta * delta * delta;
    }
    variance /= m.num_rows;
    s.row_length_std_dev = std::sqrt(variance);
    s.row_length_skewness = (skew / m.num_rows) / std::pow(s.row_length_std_dev, 3.0);
    s.row_length_variation = s.row_length_std_dev / s.row_length_mean;
    return s;
}

template <typename V>
void display_histogram(const Csr<V>& m)  // sparse_matrix.h:919-956 (percentages are of num_cols, as there)
{
    int log_counts[11] = {0};
    int max_log_length = -1, max_length = -1;
    for (int r = 0; r < m.num_rows; ++r) {
        int length = m.row_offsets[r + 1] - m.row_offsets[r];
        max_length = std::max(max_length, length);
        int log_length = -1;
        while (length > 0) {
            length /= 10;
            ++log_length;
        }
        max_log_length = std::max(max_log_length, log_length);
        ++log_counts[log_length + 1];
    }
    std::printf("CSR matrix (%d rows, %d columns, %d non-zeros, max-length %d):\n", m.num_rows, m.num_cols,
                m.num_nonzeros, max_length);
    for (int i = -1; i < max_log_length + 1; ++i)
        std::printf("\tDegree 1e%d: \t%d (%.2f%%)\n", i, log_counts[i + 1], (float)log_counts[i + 1] * 100.0 / m.num_cols);
    std::fflush(stdout);
}

// ---------------------------------------------------------------------------------------------
// Verification: SpmvGold (gpu_spmv.cu:72-92) and CompareResults (utils.h:692-742)
// ---------------------------------------------------------------------------------------------
template <typename V>
void spmv_gold(const Csr<V>& a, const V* x, const V* y_in, V* y_out, V alpha, V beta)
{
    for (int r = 0; r < a.num_rows; ++r) {
        V partial = beta * y_in[r];
        for (int k = a.row_offsets[r]; k < a.row_offsets[r + 1]; ++k)
            partial += alpha * a.values[k] * x[a.column_indices[k]];
        y_out[r] = partial;
    }
}

// CsrMatrix::Display (sparse_matrix.h:962-975): the --v2 dump of the input matrix
template <typename V>
void display_matrix(const Csr<V>& m)
{
    std::printf("Input Matrix (%d vertices, %d nonzeros):\n", m.num_rows, m.num_nonzeros);
    for (int row = 0; row < m.num_rows; ++row) {
        std::printf("%d [@%d, #%d]: ", row, m.row_offsets[row], m.row_offsets[row + 1] - m.row_offsets[row]);
        for (int k = m.row_offsets[row]; k < m.row_offsets[row + 1]; ++k)
            std::printf("%d (%f), ", m.column_indices[k], (double)m.values[k]);
        std::printf("\n");
    }
    std::fflush(stdout);
}

template <typename V>
int compare_results(const V* computed, const V* reference, int len, bool verbose = true)
{
    for (int i = 0; i < len; ++i) {
        float a = (float)computed[i], b = (float)reference[i];
        int ia, ib;
        std::memcpy(&ia, &a, 4);
        std::memcpy(&ib, &b, 4);
        float sqrt_diff = std::sqrt((float)std::abs(ia - ib));
        if (sqrt_diff > len) {
            if (verbose)
                std::printf("INCORRECT (sqrt_diff: %g): [%d]: %g != %g", sqrt_diff, i, (double)computed[i], (double)reference[i]);
            return 1;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Command line (utils.h:280-449): --key or --key=value
// ---------------------------------------------------------------------------------------------
struct CommandLineArgs {
    std::vector<std::string> keys, values;
    CommandLineArgs(int argc, char** argv)
    {
        for (int i = 1; i < argc; ++i) {
            std::string arg = argv[i];
            if (arg.size() < 2 || arg[0] != '-' || arg[1] != '-') continue;
            std::string::size_type pos = arg.find('=');
            if (pos == std::string::npos) {
                keys.push_back(arg.substr(2));
                values.push_back("");
            } else {
                keys.push_back(arg.substr(2, pos - 2));
                values.push_back(arg.substr(pos + 1));
            }
        }
    }
    bool CheckCmdLineFlag(const char* name) const
    {
        for (auto& k : keys)
            if (k == name) return true;
        return false;
    }
    template <typename T>
    void GetCmdLineArgument(const char* name, T& val) const;
};
template <>
inline void CommandLineArgs::GetCmdLineArgument<std::string>(const char* name, std::string& val) const
{
    for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i] == name) val = values[i];
}
template <>
inline void CommandLineArgs::GetCmdLineArgument<int>(const char* name, int& val) const
{
    for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i] == name && !values[i].empty()) val = (int)std::strtol(values[i].c_str(), nullptr, 0);
}
template <>
inline void CommandLineArgs::GetCmdLineArgument<long long>(const char* name, long long& val) const
{
    for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i] == name && !values[i].empty()) val = std::strtoll(values[i].c_str(), nullptr, 0);
}
template <>
inline void CommandLineArgs::GetCmdLineArgument<float>(const char* name, float& val) const
{
    for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i] == name && !values[i].empty()) val = std::strtof(values[i].c_str(), nullptr);
}

// Builds the matrix the flags ask for and prints its label like the reference drivers
// (gpu_spmv.cu:611-655, cpu_spmv.cpp:551-595).  New flags: --uniform=<nnz/row> --powerlaw=<max row>
// --banded=<half bandwidth> with --rows --cols --nnz --seed --values=ones|random.
template <typename V>
Csr<V> build_from_args(const CommandLineArgs& args, bool quiet, bool gpu_driver)
{
    std::string mtx, values = "ones";
    int grid2d = -1, grid3d = -1, wheel = -1, dense = -1, uniform = -1, powerlaw = -1, banded = -1;
    int rows = 1 << 20, cols = -1;
    long long nnz = -1, seed = -1;
    args.GetCmdLineArgument("mtx", mtx);
    args.GetCmdLineArgument("grid2d", grid2d);
    args.GetCmdLineArgument("grid3d", grid3d);
    args.GetCmdLineArgument("wheel", wheel);
    args.GetCmdLineArgument("dense", dense);
    args.GetCmdLineArgument("uniform", uniform);
    args.GetCmdLineArgument("powerlaw", powerlaw);
    args.GetCmdLineArgument("banded", banded);
    args.GetCmdLineArgument("rows", rows);
    args.GetCmdLineArgument("cols", cols);
    args.GetCmdLineArgument("nnz", nnz);
    args.GetCmdLineArgument("seed", seed);
    args.GetCmdLineArgument("values", values);
    const bool random_values = values == "random";
    if (cols < 0) cols = rows;

    Coo coo;
    if (!mtx.empty()) {
        coo = read_matrix_market(mtx, 1.0, !quiet);
        if (coo.num_rows == 1 || coo.num_cols == 1 || coo.entries.size() == 1) {
            if (!quiet) std::printf("Trivial dataset\n");
            std::exit(0);
        }
        std::printf("%s, ", mtx.c_str());
    } else if (grid2d > 0) {
        std::printf("grid2d_%d, ", grid2d);
        coo = gen_grid2d(grid2d);
    } else if (grid3d > 0) {
        std::printf("grid3d_%d, ", grid3d);
        coo = gen_grid3d(grid3d);
    } else if (wheel > 0) {
        std::printf("wheel_%d, ", wheel);  // the reference prints grid2d here (gpu_spmv.cu:639), a typo not kept
        coo = gen_wheel(wheel);
    } else if (dense > 0) {
        int size = 1 << 24;
        if (gpu_driver) args.GetCmdLineArgument("size", size);  // cpu_spmv.cpp:584 has no --size
        int drows = size / dense;
        std::printf("dense_%d_x_%d, ", drows, dense);
        coo = gen_dense(drows, dense);
    } else if (uniform > 0) {
        std::printf("uniform_%dx%d_%d, ", rows, cols, uniform);
        std::fflush(stdout);
        return gen_uniform<V>(rows, cols, uniform, random_values, seed < 0 ? 0x5EED0001ull : (uint64_t)seed);
    } else if (powerlaw > 0) {
        if (nnz < 0) nnz = 100LL * rows;
        double alpha = 0;
        Csr<V> m = gen_powerlaw<V>(rows, cols, powerlaw, nnz, random_values, seed < 0 ? 0x5EED0003ull : (uint64_t)seed, &alpha);
        std::printf("powerlaw_%d_max%d_alpha%.4f, ", rows, powerlaw, alpha);
        std::fflush(stdout);
        return m;
    } else if (banded >= 0) {
        std::printf("banded_%d_bw%d, ", rows, 2 * banded + 1);
        std::fflush(stdout);
        return gen_banded<V>(rows, banded, random_values, seed < 0 ? 0x5EED0004ull : (uint64_t)seed);
    } else {
        std::fprintf(stderr, "No graph type specified.\n");
        std::exit(1);
    }
    std::fflush(stdout);
    return coo_to_csr<V>(coo);
}

// --dumpcsr=<path>: write the CSR the driver built (test hook)
template <typename V>
void maybe_dump_csr(const CommandLineArgs& args, const Csr<V>& m)
{
    std::string path;
    args.GetCmdLineArgument("dumpcsr", path);
    if (!path.empty()) dump_csr(path, m);
    // --writemtx=<path>: save the matrix as Matrix-Market, e.g. to feed a synthetic config to the
    // reference's own binaries through --mtx=
    std::string mtx_out;
    args.GetCmdLineArgument("writemtx", mtx_out);
    if (!mtx_out.empty()) write_matrix_market(mtx_out, m);
}

}  // namespace mspmv_host
