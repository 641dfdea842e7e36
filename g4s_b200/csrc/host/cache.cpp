// Binary CSR cache.  The reference parses MatrixMarket text with `ifstream >>` on every run
// (mm/inc/CSR.h:526-553, called from mm/src/mkl_spgemm.cpp:41,46); at the sizes of BASELINE.json's configs (1.7e9
// entries) that is tens of minutes per start.  g4s_csr_read_cached keeps the parsed CSR next to the text file as
// `<file>.g4scsr` and reloads it with three bulk reads; the arrays are exactly the ones construct() produced
// (int32 rowptr / colids, fp64 values, 0-based), so everything downstream is unchanged.
//
// File layout (little-endian, 64-byte header, arrays back to back):
//   char  magic[8] = "G4SCSR1\0";  u32 version = 1, index_bytes = 4, value_bytes = 8, zerobased = 1;
//   i64   rows, cols, nnz;  u64 checksum (of the three arrays, in file order);  u64 reserved
//   i32 rowptr[rows+1];  i32 colids[nnz];  f64 values[nnz]
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "g4s_b200.h"

namespace g4s {
int fail(int status, const std::string &msg);
}
using g4s::fail;

namespace {

struct Header {
    char magic[8];
    uint32_t version, index_bytes, value_bytes, zerobased;
    int64_t rows, cols, nnz;
    uint64_t checksum, reserved;
};
static_assert(sizeof(Header) == 64, "header is 64 bytes");
const char kMagic[8] = {'G', '4', 'S', 'C', 'S', 'R', '1', '\0'};

// order-dependent 64-bit mix over 8-byte words (tail bytes zero-padded); cheap enough for multi-GB arrays
uint64_t mix(uint64_t h, const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h = (h << 29) | (h >> 35);
    }
    if (i < bytes) {
        uint64_t w = 0;
        memcpy(&w, p + i, bytes - i);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h = (h << 29) | (h >> 35);
    }
    return h;
}
uint64_t checksum_of(int rows, long long nnz, const int *rowptr, const int *colids, const double *values) {
    uint64_t h = 0x6734735f62323030ull;
    h = mix(h, rowptr, sizeof(int) * ((size_t)rows + 1));
    h = mix(h, colids, sizeof(int) * (size_t)nnz);
    h = mix(h, values, sizeof(double) * (size_t)nnz);
    return h;
}

bool read_all(FILE *f, void *dst, size_t bytes) {
    unsigned char *p = static_cast<unsigned char *>(dst);
    while (bytes) {
        const size_t chunk = bytes < ((size_t)1 << 30) ? bytes : ((size_t)1 << 30);
        if (fread(p, 1, chunk, f) != chunk) return false;
        p += chunk;
        bytes -= chunk;
    }
    return true;
}
bool write_all(FILE *f, const void *src, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(src);
    while (bytes) {
        const size_t chunk = bytes < ((size_t)1 << 30) ? bytes : ((size_t)1 << 30);
        if (fwrite(p, 1, chunk, f) != chunk) return false;
        p += chunk;
        bytes -= chunk;
    }
    return true;
}

}  // namespace

extern "C" {

int g4s_csr_write_binary(const char *path, int rows, int cols, int nnz, const int *rowptr, const int *colids,
                         const double *values) {
    if (!path || rows < 0 || cols < 0 || nnz < 0 || !rowptr || (nnz && (!colids || !values)))
        return fail(G4S_ERR_INVALID, "g4s_csr_write_binary: bad arguments");
    if (rowptr[0] != 0 || rowptr[rows] != nnz) return fail(G4S_ERR_INVALID, "g4s_csr_write_binary: rowptr does not span [0, nnz]");
    // write to a temporary name (per process: the ranks of a multi-GPU job may all miss the cache at once) and rename:
    // a reader never sees a half-written cache
    const std::string tmp = std::string(path) + ".tmp." + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(G4S_ERR_IO, "unable to open file \"" + tmp + "\" for writing");
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.version = 1;
    h.index_bytes = 4;
    h.value_bytes = 8;
    h.zerobased = 1;
    h.rows = rows;
    h.cols = cols;
    h.nnz = nnz;
    h.checksum = checksum_of(rows, nnz, rowptr, colids, values);
    bool ok = write_all(f, &h, sizeof(h)) && write_all(f, rowptr, sizeof(int) * ((size_t)rows + 1)) &&
              write_all(f, colids, sizeof(int) * (size_t)nnz) && write_all(f, values, sizeof(double) * (size_t)nnz);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(G4S_ERR_IO, std::string("writing \"") + path + "\" failed");
    }
    return G4S_OK;
}

int g4s_csr_read_binary(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids, double **values) {
    if (!path || !rows || !cols || !nnz || !rowptr || !colids || !values)
        return fail(G4S_ERR_INVALID, "g4s_csr_read_binary: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(G4S_ERR_IO, std::string("unable to open file \"") + path + "\" for reading");
    Header h;
    if (!read_all(f, &h, sizeof(h))) {
        fclose(f);
        return fail(G4S_ERR_FORMAT, std::string(path) + ": shorter than the 64-byte header");
    }
    const char *why = nullptr;
    if (memcmp(h.magic, kMagic, 8) != 0) why = "not a G4SCSR1 file";
    else if (h.version != 1) why = "unknown version";
    else if (h.index_bytes != 4 || h.value_bytes != 8 || h.zerobased != 1) why = "not CSR<int32, fp64, 0-based>";
    else if (h.rows < 0 || h.cols < 0 || h.nnz < 0 || h.rows > 2147483646LL || h.cols > 2147483647LL || h.nnz > 2147483647LL)
        why = "dimensions out of the int32 range";
    if (!why) {  // the file must hold exactly the arrays the header promises
        struct stat st;
        const long long want = 64 + 4 * (h.rows + 1) + 12 * h.nnz;
        if (fstat(fileno(f), &st) != 0 || (long long)st.st_size != want) why = "file size does not match the header";
    }
    if (why) {
        fclose(f);
        return fail(G4S_ERR_FORMAT, std::string(path) + ": " + why);
    }
    int *rp = static_cast<int *>(malloc(sizeof(int) * ((size_t)h.rows + 1)));
    int *ci = static_cast<int *>(malloc(sizeof(int) * (size_t)(h.nnz ? h.nnz : 1)));
    double *va = static_cast<double *>(malloc(sizeof(double) * (size_t)(h.nnz ? h.nnz : 1)));
    auto drop = [&]() {
        free(rp);
        free(ci);
        free(va);
        fclose(f);
    };
    if (!rp || !ci || !va) {
        drop();
        return fail(G4S_ERR_ALLOC, "host allocation failed");
    }
    if (!read_all(f, rp, sizeof(int) * ((size_t)h.rows + 1)) || !read_all(f, ci, sizeof(int) * (size_t)h.nnz) ||
        !read_all(f, va, sizeof(double) * (size_t)h.nnz)) {
        drop();
        return fail(G4S_ERR_IO, std::string(path) + ": read failed");
    }
    bool sane = rp[0] == 0 && rp[h.rows] == (int)h.nnz;
    for (int64_t i = 0; sane && i < h.rows; ++i) sane = rp[i] <= rp[i + 1];
    if (sane) {  // a stale, foreign or crafted file must not hand out-of-range columns to the GPU kernels
        int bad = 0;
#pragma omp parallel for reduction(| : bad) schedule(static)
        for (int64_t k = 0; k < h.nnz; ++k) bad |= (ci[k] < 0) | (ci[k] >= (int)h.cols);
        sane = bad == 0;
    }
    if (sane && checksum_of((int)h.rows, h.nnz, rp, ci, va) != h.checksum) sane = false;
    if (!sane) {
        drop();
        return fail(G4S_ERR_FORMAT, std::string(path) + ": contents do not match the header (corrupt cache)");
    }
    fclose(f);
    *rows = (int)h.rows;
    *cols = (int)h.cols;
    *nnz = (int)h.nnz;
    *rowptr = rp;
    *colids = ci;
    *values = va;
    return G4S_OK;
}

int g4s_csr_read_cached(const char *mtx_path, int *rows, int *cols, int *nnz, int **rowptr, int **colids, double **values,
                        int *cache_hit) {
    if (!mtx_path) return fail(G4S_ERR_INVALID, "g4s_csr_read_cached: null argument");
    if (cache_hit) *cache_hit = 0;
    const std::string cache = std::string(mtx_path) + ".g4scsr";
    struct stat sm, sc;
    const bool have_mtx = stat(mtx_path, &sm) == 0;
    // a cache older than its text file is stale (compared at the file system's nanosecond resolution: a text file rewritten
    // within the same second as its cache must not keep it); a cache without a text file is used as it is
    auto not_older = [](const struct stat &c, const struct stat &m) {
        return c.st_mtim.tv_sec > m.st_mtim.tv_sec || (c.st_mtim.tv_sec == m.st_mtim.tv_sec && c.st_mtim.tv_nsec >= m.st_mtim.tv_nsec);
    };
    if (stat(cache.c_str(), &sc) == 0 && (!have_mtx || not_older(sc, sm))) {
        if (g4s_csr_read_binary(cache.c_str(), rows, cols, nnz, rowptr, colids, values) == G4S_OK) {
            if (cache_hit) *cache_hit = 1;
            return G4S_OK;
        }  // unreadable cache: fall through to the text file and rewrite it
    }
    const int rc = g4s_csr_read_matrix_market(mtx_path, rows, cols, nnz, rowptr, colids, values);
    if (rc != G4S_OK) return rc;
    g4s_csr_write_binary(cache.c_str(), *rows, *cols, *nnz, *rowptr, *colids, *values);  // best effort
    return G4S_OK;
}

}  // extern "C"
