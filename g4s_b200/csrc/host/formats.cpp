// Host-side input formats of the mm/ path, behind the C ABI:
//   g4s_csr_read_matrix_market  <- CSR<IT,NT>::construct            (mm/inc/CSR.h:485-669, banner :440-478)
//   g4s_csr_from_edge_list      <- CSR<IT,NT>::CSR(graph&)           (mm/inc/CSR.h:255-329, graph.h:4-25)
//   g4s_csr_submatrix           <- CSR(const CSR&, M_, N_, M_start, N_start) (mm/inc/CSR.h:691-733)
// Same accepted inputs, same resulting CSR (row-major (row, col) order, duplicates KEPT by the MatrixMarket
// reader and SUMMED by the edge-list constructor); malformed input returns G4S_ERR_FORMAT / G4S_ERR_IO where
// the reference throws std::runtime_error.  Outputs are malloc'd (g4s_free).
#include <omp.h>
#include <parallel/algorithm>

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "g4s_b200.h"

namespace g4s {
int fail(int status, const std::string &msg);
}
using g4s::fail;

namespace {

std::vector<std::string> words(const std::string &line) {
    std::vector<std::string> out;
    std::istringstream ss(line);
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
}

template <class T>
T *to_malloc(const std::vector<T> &v) {
    T *p = static_cast<T *>(malloc(sizeof(T) * std::max<size_t>(v.size(), 1)));
    if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

struct Entry {
    long key;  // cols * row + col
    double val;
};

}  // namespace

extern "C" {

int g4s_csr_read_matrix_market(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids,
                               double **values) {
    if (!path || !rows || !cols || !nnz || !rowptr || !colids || !values)
        return fail(G4S_ERR_INVALID, "g4s_csr_read_matrix_market: null argument");
    std::ifstream in(path);
    if (!in) return fail(G4S_ERR_IO, std::string("unable to open file \"") + path + "\" for reading");
    std::string line;
    std::getline(in, line);
    const std::vector<std::string> banner = words(line);
    if (banner.size() != 5 || banner[0] != "%%MatrixMarket" || banner[1] != "matrix")
        return fail(G4S_ERR_FORMAT, "invalid MatrixMarket banner");
    const std::string &storage = banner[2], &type = banner[3], &symmetry = banner[4];
    if (storage != "array" && storage != "coordinate")
        return fail(G4S_ERR_FORMAT, "invalid MatrixMarket storage format [" + storage + "]");
    if (storage == "array") return fail(G4S_ERR_FORMAT, "not impl storage type array");
    const bool pattern = type == "pattern", complex_ = type == "complex";
    if (!pattern && !complex_ && type != "real" && type != "integer")
        return fail(G4S_ERR_FORMAT, "invalid MatrixMarket data type [" + type + "]");
    const bool general = symmetry == "general", skew = symmetry == "skew-symmetric";
    if (!general && !skew && symmetry != "symmetric" && symmetry != "hermitian")
        return fail(G4S_ERR_FORMAT, "invalid MatrixMarket symmetry [" + symmetry + "]");
    if (symmetry == "hermitian") return fail(G4S_ERR_FORMAT, "not impl matrix type: hermitian");

    do {
        if (!std::getline(in, line)) line.clear();
    } while (!line.empty() && line[0] == '%');
    const std::vector<std::string> size = words(line);
    if (size.size() != 3) return fail(G4S_ERR_FORMAT, "invalid MatrixMarket coordinate format");
    const long R = atol(size[0].c_str()), C = atol(size[1].c_str()), E = atol(size[2].c_str());
    if (R < 0 || C <= 0 || R > 2147483646L || C > 2147483647L)
        return fail(G4S_ERR_FORMAT, "MatrixMarket dimensions out of int32 range");
    if (E <= 0) return fail(G4S_ERR_FORMAT, "something wrong: nnz is 0");

    // ---- body: a whitespace-separated TOKEN stream, T tokens per entry (the reference reads it with `in >> i >> j >> v`,
    // so line breaks carry no meaning).  The rest of the file is read in one piece and tokenised / parsed by all cores:
    // chunks are cut at token starts, a first pass counts the tokens of every chunk so that each chunk knows which
    // entry its first token belongs to, a second pass parses the entries whose FIRST token lies in the chunk.
    const int T = pattern ? 2 : (complex_ ? 4 : 3);
    std::vector<char> text;
    {
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        in.seekg(here);
        text.resize((size_t)(end - here));
        if (!text.empty() && !in.read(text.data(), (std::streamsize)text.size()))
            return fail(G4S_ERR_IO, std::string("read error on \"") + path + "\"");
    }
    const char *buf = text.data();
    const size_t len = text.size();
    auto is_space = [](char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    const int nchunks = (int)std::max<size_t>(1, std::min<size_t>((size_t)omp_get_max_threads() * 8, len / 65536 + 1));
    std::vector<size_t> cut(nchunks + 1);
    for (int c = 0; c <= nchunks; ++c) {
        size_t p = len / nchunks * c;
        if (c == nchunks) p = len;
        while (p > 0 && p < len && !(is_space(buf[p - 1]) && !is_space(buf[p]))) ++p;  // move on to the next token start
        cut[c] = p;
    }
    std::vector<long> tok0(nchunks + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < nchunks; ++c) {
        long n = 0;
        bool in_tok = false;
        for (size_t p = cut[c]; p < cut[c + 1]; ++p) {
            const bool sp_ = is_space(buf[p]);
            n += (!sp_ && !in_tok);
            in_tok = !sp_;
        }
        tok0[c + 1] = n;
    }
    for (int c = 0; c < nchunks; ++c) tok0[c + 1] += tok0[c];
    const long avail = std::min<long>(E, tok0[nchunks] / T);  // complete entries present in the file
    std::vector<long> I(avail), J(avail);
    std::vector<double> V(pattern ? 0 : avail);
    // integer token: [+-]digits, the whole token; value token: what `in >> double` accepts of the usual notations
    auto parse_long = [](const char *b, const char *e, long &out) {
        bool neg = false;
        if (b < e && (*b == '+' || *b == '-')) neg = *b++ == '-';
        if (b == e || e - b > 18) return false;  // 18 digits cannot overflow a long
        long v = 0;
        for (; b < e; ++b) {
            if (*b < '0' || *b > '9') return false;
            v = v * 10 + (*b - '0');
        }
        out = neg ? -v : v;
        return true;
    };
    auto parse_double = [](const char *b, const char *e, double &out) {
        if (b < e && *b == '+') ++b;
        const auto r = std::from_chars(b, e, out);
        return r.ec == std::errc() && r.ptr == e;
    };
    long first_bad = avail, first_range = avail;  // first entry that does not parse / lies outside the matrix
#pragma omp parallel for schedule(dynamic, 1) reduction(min : first_bad, first_range)
    for (int c = 0; c < nchunks; ++c) {
        size_t p = cut[c];
        long g = tok0[c];
        while (p < cut[c + 1]) {
            while (p < len && is_space(buf[p])) ++p;
            if (p >= cut[c + 1]) break;
            if (g % T != 0 || g / T >= avail) {  // not the first token of an entry we own: skip it
                while (p < len && !is_space(buf[p])) ++p;
                ++g;
                continue;
            }
            const long k = g / T;
            const char *tb[4], *te[4];
            size_t q = p;
            for (int t = 0; t < T; ++t) {  // the entry's tokens may run past this chunk's end
                while (q < len && is_space(buf[q])) ++q;
                tb[t] = buf + q;
                while (q < len && !is_space(buf[q])) ++q;
                te[t] = buf + q;
            }
            long i = 0, j = 0;
            double v = 1.0, imag = 0.0;
            bool ok = parse_long(tb[0], te[0], i) && parse_long(tb[1], te[1], j);
            if (ok && !pattern) ok = parse_double(tb[2], te[2], v);
            if (ok && complex_) ok = parse_double(tb[3], te[3], imag);
            if (!ok) {
                first_bad = std::min(first_bad, k);
            } else {
                --i;
                --j;
                if (i < 0 || i >= R || j < 0 || j >= C) first_range = std::min(first_range, k);
                I[k] = i;
                J[k] = j;
                if (!pattern) V[k] = v;
            }
            while (p < len && !is_space(buf[p])) ++p;  // on to the token after the entry's first one
            ++g;
        }
    }
    if (first_range < first_bad) return fail(G4S_ERR_FORMAT, "MatrixMarket entry out of range");
    if (first_bad != E) return fail(G4S_ERR_FORMAT, "read nnz not equal to declared nnz " + std::to_string(first_bad));
    text = std::vector<char>();
    // entries in file order, the mirror image of an off-diagonal entry of a symmetric file right after its source
    std::vector<Entry> ent;
    if (general) {
        ent.resize(E);
#pragma omp parallel for schedule(static)
        for (long k = 0; k < E; ++k) ent[k] = {C * I[k] + J[k], pattern ? 1.0 : V[k]};
    } else {
        ent.reserve(2 * E);
        for (long k = 0; k < E; ++k) {
            const double v = pattern ? 1.0 : V[k];
            ent.push_back({C * I[k] + J[k], v});
            if (I[k] != J[k]) ent.push_back({C * J[k] + I[k], skew ? -v : v});
        }
    }
    if (ent.size() > 2147483647UL) return fail(G4S_ERR_FORMAT, "more than 2^31-1 entries after symmetric expansion");
    // (row, col) order.  The reference's std::sort leaves duplicates in unspecified order; file order is kept here.
    __gnu_parallel::stable_sort(ent.begin(), ent.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });

    std::vector<int> rp(R + 1, 0), ci(ent.size());
    std::vector<double> va(ent.size());
    for (size_t e = 0; e < ent.size(); ++e) {
        rp[ent[e].key / C + 1]++;
        ci[e] = static_cast<int>(ent[e].key % C);
        va[e] = ent[e].val;
    }
    for (long r = 0; r < R; ++r) rp[r + 1] += rp[r];
    *rows = static_cast<int>(R);
    *cols = static_cast<int>(C);
    *nnz = static_cast<int>(ent.size());
    *rowptr = to_malloc(rp);
    *colids = to_malloc(ci);
    *values = to_malloc(va);
    if (!*rowptr || !*colids || !*values) return fail(G4S_ERR_ALLOC, "g4s_csr_read_matrix_market: out of memory");
    return G4S_OK;
}

int g4s_csr_from_edge_list(long m, long n, const long *start, const long *end, const double *w, int *nnz,
                           int **rowptr, int **colids, double **values) {
    if (m < 0 || n < 0 || (m > 0 && (!start || !end || !w)) || !nnz || !rowptr || !colids || !values)
        return fail(G4S_ERR_INVALID, "g4s_csr_from_edge_list: bad arguments");
    if (n > 2147483646L) return fail(G4S_ERR_INVALID, "g4s_csr_from_edge_list: more than 2^31-2 vertices");
    bool in_range = true;
#pragma omp parallel for schedule(static) reduction(&& : in_range)
    for (long e = 0; e < m; ++e) in_range = in_range && start[e] >= 0 && start[e] < n && end[e] >= 0 && end[e] < n;
    if (!in_range) return fail(G4S_ERR_FORMAT, "g4s_csr_from_edge_list: vertex id out of range");
    // Runs of equal start vertex are sorted by (end, weight) and equal (start, end) pairs are summed left to
    // right; a start vertex that shows up again later opens a new, unmerged run — as in the reference.  The runs are
    // independent, so they are sorted and merged by all cores (in place, in a copy of the (end, weight) pairs); one
    // sequential pass over the RUNS then places them row by row in order of appearance.
    typedef std::pair<long, double> Half;  // (end, weight): the start vertex is the same throughout a run
    std::vector<long> run_begin;
    for (long e = 0; e < m; ++e)
        if (e == 0 || start[e] != start[e - 1]) run_begin.push_back(e);
    const long nruns = (long)run_begin.size();
    run_begin.push_back(m);
    std::vector<Half> half(static_cast<size_t>(m));
#pragma omp parallel for schedule(static)
    for (long e = 0; e < m; ++e) half[e] = Half(end[e], w[e]);
    std::vector<long> kept(nruns + 1, 0);  // entries of a run after merging
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < nruns; ++r) {
        Half *b = half.data() + run_begin[r], *e = half.data() + run_begin[r + 1];
        std::sort(b, e);
        Half *out = b;
        for (Half *q = b + 1; q < e; ++q) {
            if (q->first == out->first) out->second += q->second;
            else *++out = *q;
        }
        kept[r] = out - b + 1;
    }
    std::vector<int> rp(n + 1, 0);
    std::vector<long> place(nruns);  // position of the run inside its row
    long total = 0;
    for (long r = 0; r < nruns; ++r) {
        const long row = start[run_begin[r]];
        place[r] = rp[row + 1];
        rp[row + 1] += (int)kept[r];
        total += kept[r];
        if (total > 2147483647L) return fail(G4S_ERR_INVALID, "g4s_csr_from_edge_list: nnz exceeds int32");
    }
    for (long r = 0; r < n; ++r) rp[r + 1] += rp[r];
    struct Merged {
        size_t n;
        size_t size() const { return n; }
    } merged{static_cast<size_t>(total)};
    std::vector<int> ci(merged.size());
    std::vector<double> va(merged.size());
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < nruns; ++r) {
        const Half *b = half.data() + run_begin[r];
        const long at = rp[start[run_begin[r]]] + place[r];
        for (long k = 0; k < kept[r]; ++k) {
            ci[at + k] = static_cast<int>(b[k].first);
            va[at + k] = b[k].second;
        }
    }
    *nnz = static_cast<int>(merged.size());
    *rowptr = to_malloc(rp);
    *colids = to_malloc(ci);
    *values = to_malloc(va);
    if (!*rowptr || !*colids || !*values) return fail(G4S_ERR_ALLOC, "g4s_csr_from_edge_list: out of memory");
    return G4S_OK;
}

int g4s_csr_submatrix(int rows, int cols, const int *rowptr, const int *colids, const double *values, int M_,
                      int N_, int M_start, int N_start, int *nnz, int **orpt, int **ocol, double **oval) {
    if (!rowptr || !nnz || !orpt || !ocol || !oval || M_ < 0 || N_ < 0 || M_start < 0 || N_start < 0)
        return fail(G4S_ERR_INVALID, "g4s_csr_submatrix: bad arguments");
    if ((long)M_ + M_start > rows) return fail(G4S_ERR_SHAPE, "matrix subsect error M");
    if ((long)N_ + N_start > cols) return fail(G4S_ERR_SHAPE, "matrix subsect error N");
    std::vector<int> rp(M_ + 1, 0), ci;
    std::vector<double> va;
    for (int i = 0; i < M_; ++i) {
        for (int j = rowptr[i + M_start]; j < rowptr[i + M_start + 1]; ++j) {
            const int c = colids[j];
            if (c >= N_start && c < N_start + N_) {
                ci.push_back(c - N_start);
                va.push_back(values[j]);
            }
        }
        rp[i + 1] = static_cast<int>(ci.size());
    }
    *nnz = static_cast<int>(ci.size());
    *orpt = to_malloc(rp);
    *ocol = to_malloc(ci);
    *oval = to_malloc(va);
    if (!*orpt || !*ocol || !*oval) return fail(G4S_ERR_ALLOC, "g4s_csr_submatrix: out of memory");
    return G4S_OK;
}

}  // extern "C"
