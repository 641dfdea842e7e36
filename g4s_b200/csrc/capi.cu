// C-ABI glue: error state, device checks, CSR handles, SpMV entry points, Timings, partitioner.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace g4s {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }
int fail(int status, const std::string &msg) {
    t_error = msg;
    return status;
}

static int g_sm_count[64];
static int g_dev_ok[64];

int ensure_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(G4S_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (dev < 0 || dev >= 64) return fail(G4S_ERR_CUDA, "device ordinal out of range");
    if (g_dev_ok[dev] == 1) return G4S_OK;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(G4S_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(G4S_ERR_CUDA, "g4s_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor) + " (there is no fallback path)");
    g_sm_count[dev] = prop.multiProcessorCount;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {  // keep freed scratch cached across calls
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_dev_ok[dev] = 1;
    return G4S_OK;
}

int sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64 && g_sm_count[dev] > 0) ? g_sm_count[dev] : 148;
}

struct XParts;
int spmv_run(g4s_csr *h, const double *x, double *y, const int *row_map, bool accum, cudaStream_t stream,
             const XParts *parts = nullptr, int chunk_begin = 0, int chunk_end = -1);
int spmv_host_pipelined(g4s_csr *h, const double *x, double *y);
int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host, cudaStream_t stream);
void spmv_free_host_pipe(g4s_csr *h);
int spmv_run_partitioned(g4s_csr *h, int world, int self, const double *const *x_parts, const int *cuts, double *y,
                         const unsigned long long *flags, unsigned long long epoch,
                         unsigned long long *const *signal_arrays, cudaStream_t stream);
int peer_signal(unsigned long long *const *flag_arrays, int world, int self, unsigned long long epoch, cudaStream_t stream);
int spmv_build_plan(g4s_csr *h, cudaStream_t stream);
void spmv_free_plan(g4s_csr *h);

// Device arrays carry 64 bytes of slack so that 16-byte bulk copies may round the last element up.
int alloc_csr(g4s_csr **out, int rows, int cols, long long nnz) {
    g4s_csr *h = new (std::nothrow) g4s_csr();
    if (!h) return fail(G4S_ERR_ALLOC, "host allocation failed");
    h->rows = rows;
    h->cols = cols;
    h->nnz = nnz;
    h->owns = true;
    // stream-ordered pool (release threshold = keep): repeated create/destroy cycles, as in the reference's
    // benchmark loop (mm/src/mkl_spgemm.cpp:67-79), reuse the same device memory instead of paying cudaMalloc/cudaFree
    h->pooled = true;
    cudaError_t e = cudaMallocAsync(&h->rowptr, sizeof(int) * ((size_t)rows + 1) + 64, 0);
    if (e == cudaSuccess) e = cudaMallocAsync(&h->colids, sizeof(int) * (size_t)nnz + 64, 0);
    if (e == cudaSuccess) e = cudaMallocAsync(&h->values, sizeof(double) * (size_t)nnz + 64, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g4s_csr_destroy(h);
        return fail(G4S_ERR_ALLOC, std::string("device allocation of CSR: ") + cudaGetErrorString(e));
    }
    *out = h;
    return G4S_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace g4s


// ---- pageable host memory <-> device through pinned staging ----------------------------------------------------
// cudaMemcpy on pageable memory runs at a few GB/s (driver-side staging on one thread, plus first-touch page faults
// on freshly malloc'd destinations).  Large copies go through two pinned 32 MB buffers instead: the PCIe transfer of
// one piece overlaps the OpenMP-parallel memcpy of the next.  Pinned (or small) host buffers are copied directly.
namespace {
constexpr size_t STAGE_BYTES = 32u << 20;
struct Staging {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    int device = -1;
    bool ensure() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (buf[0] && dev == device) return true;
        device = dev;
        for (int i = 0; i < 2; ++i) {
            if (cudaMallocHost(&buf[i], STAGE_BYTES) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        return cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
    }
};
thread_local Staging t_stage;

void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    const size_t piece = 1u << 20;
    const long long n = (long long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        const size_t off = (size_t)i * piece;
        memcpy((char *)dst + off, (const char *)src + off, std::min(piece, bytes - off));
    }
}
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
}  // namespace

namespace g4s {
int copy_h2d(void *dst_dev, const void *src_host, size_t bytes) {
    if (!bytes) return G4S_OK;
    if (bytes < (4u << 20) || is_pinned(src_host) || !t_stage.ensure()) {
        G4S_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
        return G4S_OK;
    }
    Staging &st = t_stage;
    int i = 0;
    for (size_t off = 0; off < bytes; off += STAGE_BYTES, ++i) {
        const size_t n = std::min(STAGE_BYTES, bytes - off);
        const int b = i & 1;
        if (i >= 2) G4S_CUDA(cudaEventSynchronize(st.ev[b]));  // the transfer that last used this buffer is done
        parallel_memcpy(st.buf[b], (const char *)src_host + off, n);
        G4S_CUDA(cudaMemcpyAsync((char *)dst_dev + off, st.buf[b], n, cudaMemcpyHostToDevice, st.stream));
        G4S_CUDA(cudaEventRecord(st.ev[b], st.stream));
    }
    G4S_CUDA(cudaStreamSynchronize(st.stream));
    return G4S_OK;
}
int copy_d2h(void *dst_host, const void *src_dev, size_t bytes) {
    if (!bytes) return G4S_OK;
    if (bytes < (4u << 20) || is_pinned(dst_host) || !t_stage.ensure()) {
        G4S_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
        return G4S_OK;
    }
    Staging &st = t_stage;
    G4S_CUDA(cudaDeviceSynchronize());  // the source was produced on other streams
    const int pieces = (int)((bytes + STAGE_BYTES - 1) / STAGE_BYTES);
    auto issue = [&](int i) -> cudaError_t {
        const size_t off = (size_t)i * STAGE_BYTES, n = std::min(STAGE_BYTES, bytes - off);
        cudaError_t e = cudaMemcpyAsync(st.buf[i & 1], (const char *)src_dev + off, n, cudaMemcpyDeviceToHost, st.stream);
        return e != cudaSuccess ? e : cudaEventRecord(st.ev[i & 1], st.stream);
    };
    G4S_CUDA(issue(0));
    for (int i = 0; i < pieces; ++i) {
        if (i + 1 < pieces) G4S_CUDA(issue(i + 1));  // the other buffer was drained in the previous iteration
        G4S_CUDA(cudaEventSynchronize(st.ev[i & 1]));
        const size_t off = (size_t)i * STAGE_BYTES, n = std::min(STAGE_BYTES, bytes - off);
        parallel_memcpy((char *)dst_host + off, st.buf[i & 1], n);
    }
    return G4S_OK;
}
}  // namespace g4s

// ---- partitioner (BIN::set_rows_offset, mm/inc/BIN.h:100-122) ------------------------------------------
template <class T>
static int partition_rows(const T *prefix, int rows, int parts, int *cuts) {
    if (!prefix || !cuts || rows < 0 || parts < 1) return g4s::fail(G4S_ERR_INVALID, "g4s_partition_rows: bad arguments");
    const long long total = (long long)prefix[rows] - (long long)prefix[0];
    const long long avg = (total + parts - 1) / parts;
    cuts[0] = 0;
    for (int p = 0; p < parts; ++p) {
        const long long target = (long long)prefix[0] + avg * (p + 1);
        const T *it = std::lower_bound(prefix, prefix + rows + 1, target,
                                       [](const T &a, long long b) { return (long long)a < b; });
        long long idx = it - prefix;
        cuts[p + 1] = (int)std::min<long long>(idx, rows);
    }
    cuts[parts] = rows;
    return G4S_OK;
}
using namespace g4s;

extern "C" {

const char *g4s_last_error(void) { return t_error.c_str(); }
const char *g4s_version(void) { return "g4s_b200 0.1 (sm_100a)"; }
long long g4s_kernel_launch_count(void) { return g_launches.load(); }

// scan(in, out, N) of the reference (mm/inc/utility.h:166-209) on device arrays; the implementation lives in spmv.cu
int g4s_exclusive_scan_i32_device(const int *in_dev, int *out_dev, long long n, int write_total, long long *total_host,
                                  void *stream) {
    if (n < 0 || (n > 0 && (!in_dev || !out_dev)) || (write_total && !out_dev))
        return fail(G4S_ERR_INVALID, "g4s_exclusive_scan_i32_device: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    return exclusive_scan_i32(in_dev, out_dev, n, write_total, total_host, (cudaStream_t)stream);
}

int g4s_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int g4s_set_device(int device) {
    G4S_CUDA(cudaSetDevice(device));
    return ensure_device();
}

// ---- Timings (mm/src/Timings.cpp) -----------------------------------------------------------------------
void g4s_timings_init(g4s_timings *t) {
    if (!t) return;
    memset(t, 0, sizeof(*t));
    t->measure_separate = 1;
    t->measure_total = 1;
}
void g4s_timings_add(g4s_timings *a, const g4s_timings *b) {
    a->create += b->create;
    a->spmm += b->spmm;
    a->convert += b->convert;
    a->order += b->order;
    a->export_csr += b->export_csr;
    a->destroy += b->destroy;
    a->total += b->total;
}
void g4s_timings_div(g4s_timings *t, double x) {
    t->create /= x;
    t->spmm /= x;
    t->convert /= x;
    t->order /= x;
    t->export_csr /= x;
    t->destroy /= x;
    t->total /= x;
}
void g4s_timings_print(const g4s_timings *t, double total_flop) {
    const double g = total_flop / 1e9;
    const double sum = t->create + t->spmm + t->convert + t->order + t->export_csr + t->destroy;
    printf("total flop %lf\n", total_flop);
    if (!t->measure_separate) return;
    const char *names[6] = {"create", "spmm", "convert", "order", "export_csr", "destroy"};
    const double v[6] = {t->create, t->spmm, t->convert, t->order, t->export_csr, t->destroy};
    printf("time(ms):\n");
    for (int i = 0; i < 6; ++i) printf("    %-18s %8.3lfms %6.2lf%%\n", names[i], 1000 * v[i], v[i] / t->total * 100);
    printf("    %-18s %8.3lfms %6.2lf%%\n", "sum_total", 1000 * sum, sum / t->total * 100);
    printf("perf(Gflops):\n");
    for (int i = 0; i < 6; ++i) printf("    %-18s %6.2lf\n", names[i], g / v[i]);
    printf("    %-18s %6.2lf\n", "total", g / t->total);
}

// ---- CSR handles ----------------------------------------------------------------------------------------
int g4s_csr_create_host(g4s_csr_t *out, int rows, int cols, const int *rowptr, const int *colids,
                        const double *values) {
    if (!out || rows < 0 || cols < 0 || !rowptr) return fail(G4S_ERR_INVALID, "g4s_csr_create_host: bad arguments");
    const long long nnz = rowptr[rows];
    if (nnz < 0 || (nnz > 0 && (!colids || !values)))
        return fail(G4S_ERR_INVALID, "g4s_csr_create_host: null colids/values");
    int rc = ensure_device();
    if (rc) return rc;
    g4s_csr *h = nullptr;
    rc = alloc_csr(&h, rows, cols, nnz);
    if (rc) return rc;
    rc = g4s::copy_h2d(h->rowptr, rowptr, sizeof(int) * ((size_t)rows + 1));
    if (rc == G4S_OK && nnz) rc = g4s::copy_h2d(h->colids, colids, sizeof(int) * (size_t)nnz);
    if (rc == G4S_OK && nnz) rc = g4s::copy_h2d(h->values, values, sizeof(double) * (size_t)nnz);
    if (rc != G4S_OK) {
        g4s_csr_destroy(h);
        return rc;
    }
    *out = h;
    return G4S_OK;
}

int g4s_csr_create_device(g4s_csr_t *out, int rows, int cols, const int *rowptr_dev, const int *colids_dev,
                          const double *values_dev, void *stream) {
    if (!out || rows < 0 || cols < 0 || !rowptr_dev) return fail(G4S_ERR_INVALID, "g4s_csr_create_device: bad arguments");
    if (!aligned16(rowptr_dev) || !aligned16(colids_dev) || !aligned16(values_dev))
        return fail(G4S_ERR_INVALID, "g4s_csr_create_device: device arrays must be 16-byte aligned");
    int rc = ensure_device();
    if (rc) return rc;
    int nnz = 0;
    G4S_CUDA(cudaMemcpyAsync(&nnz, rowptr_dev + rows, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    G4S_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (nnz < 0) return fail(G4S_ERR_INVALID, "g4s_csr_create_device: negative nnz");
    g4s_csr *h = new (std::nothrow) g4s_csr();
    if (!h) return fail(G4S_ERR_ALLOC, "host allocation failed");
    h->rows = rows;
    h->cols = cols;
    h->nnz = nnz;
    h->rowptr = const_cast<int *>(rowptr_dev);
    h->colids = const_cast<int *>(colids_dev);
    h->values = const_cast<double *>(values_dev);
    h->owns = false;
    *out = h;
    return G4S_OK;
}

int g4s_csr_destroy(g4s_csr_t h) {
    if (!h) return G4S_OK;
    spmv_free_host_pipe(h);
    spmv_free_plan(h);
    if (h->owns && h->pooled) {
        cudaDeviceSynchronize();  // cudaFree's implicit guarantee: nothing in flight still uses the arrays
        if (h->pool_base) {
            cudaFreeAsync(h->pool_base, 0);
            h->rowptr = h->colids = nullptr;
            h->values = nullptr;
        }
        if (h->rowptr) cudaFreeAsync(h->rowptr, 0);
        if (h->colids) cudaFreeAsync(h->colids, 0);
        if (h->values) cudaFreeAsync(h->values, 0);
    } else if (h->owns) {
        if (h->rowptr) cudaFree(h->rowptr);
        if (h->colids) cudaFree(h->colids);
        if (h->values) cudaFree(h->values);
    }
    if (h->row_map) cudaFree(h->row_map);
    if (h->x_dev) cudaFree(h->x_dev);
    if (h->y_dev) cudaFree(h->y_dev);
    delete h;
    return G4S_OK;
}

int g4s_csr_shape(g4s_csr_t h, int *rows, int *cols, long long *nnz) {
    if (!h) return fail(G4S_ERR_INVALID, "null handle");
    if (rows) *rows = h->rows;
    if (cols) *cols = h->cols;
    if (nnz) *nnz = h->nnz;
    return G4S_OK;
}

int g4s_csr_device_arrays(g4s_csr_t h, const int **rowptr_dev, const int **colids_dev, const double **values_dev) {
    if (!h) return fail(G4S_ERR_INVALID, "null handle");
    if (rowptr_dev) *rowptr_dev = h->rowptr;
    if (colids_dev) *colids_dev = h->colids;
    if (values_dev) *values_dev = h->values;
    return G4S_OK;
}

int g4s_csr_download(g4s_csr_t h, int *rowptr, int *colids, double *values) {
    if (!h) return fail(G4S_ERR_INVALID, "null handle");
    int rc = G4S_OK;
    if (rowptr) rc = g4s::copy_d2h(rowptr, h->rowptr, sizeof(int) * ((size_t)h->rows + 1));
    if (rc == G4S_OK && colids && h->nnz) rc = g4s::copy_d2h(colids, h->colids, sizeof(int) * (size_t)h->nnz);
    if (rc == G4S_OK && values && h->nnz) rc = g4s::copy_d2h(values, h->values, sizeof(double) * (size_t)h->nnz);
    return rc;
}

// ---- SpMV -----------------------------------------------------------------------------------------------
int g4s_spmv_device(g4s_csr_t A, const double *x_dev, double *y_dev, void *stream) {
    if (!A || (!x_dev && A->cols) || (!y_dev && A->rows)) return fail(G4S_ERR_INVALID, "g4s_spmv_device: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    return spmv_run(A, x_dev, y_dev, nullptr, false, (cudaStream_t)stream);
}

int g4s_spmv_device_ex(g4s_csr_t A, const double *x_dev, double *y_dev, const int *row_map_dev, int accumulate,
                       void *stream) {
    if (!A || (!x_dev && A->cols) || (!y_dev && A->rows)) return fail(G4S_ERR_INVALID, "g4s_spmv_device_ex: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (!row_map_dev) row_map_dev = A->row_map;  // a row-compressed handle scatters through its own map
    return spmv_run(A, x_dev, y_dev, row_map_dev, accumulate != 0, (cudaStream_t)stream);
}

int g4s_spmv_partitioned_device(g4s_csr_t A, int world, int self, const double *const *x_parts, const int *cuts,
                                double *y_dev, const unsigned long long *ready_flags_dev, unsigned long long epoch,
                                unsigned long long *const *signal_arrays, void *stream) {
    if (!A || !x_parts || !cuts || (!y_dev && A->rows)) return fail(G4S_ERR_INVALID, "g4s_spmv_partitioned_device: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    return spmv_run_partitioned(A, world, self, x_parts, cuts, y_dev, ready_flags_dev, epoch, signal_arrays,
                                (cudaStream_t)stream);
}

int g4s_spmv_partition_info(g4s_csr_t A, int *n_remote_columns, unsigned *owner_mask) {
    if (!A) return fail(G4S_ERR_INVALID, "null handle");
    if (n_remote_columns) *n_remote_columns = A->plan.part_colids ? A->plan.part_n_needed : -1;
    if (owner_mask) *owner_mask = A->plan.part_owner_mask;
    return G4S_OK;
}
int g4s_spmv_partition_set_wait_mask(g4s_csr_t A, unsigned mask) {
    if (!A) return fail(G4S_ERR_INVALID, "null handle");
    A->plan.part_owner_mask = mask;
    return G4S_OK;
}

int g4s_peer_signal(unsigned long long *const *flag_arrays, int world, int self, unsigned long long epoch, void *stream) {
    if (!flag_arrays) return fail(G4S_ERR_INVALID, "g4s_peer_signal: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    return peer_signal(flag_arrays, world, self, epoch, (cudaStream_t)stream);
}

// ---- peer memory (CUDA IPC): buffers that other ranks' kernels on the same box read over NVLink ------------------
int g4s_peer_alloc(size_t bytes, void **ptr_dev, unsigned char *handle64) {
    if (!ptr_dev || !handle64) return fail(G4S_ERR_INVALID, "g4s_peer_alloc: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    G4S_CUDA(cudaMalloc(ptr_dev, bytes ? bytes : 8));
    cudaIpcMemHandle_t h;
    G4S_CUDA(cudaIpcGetMemHandle(&h, *ptr_dev));
    memcpy(handle64, &h, 64);
    return G4S_OK;
}
int g4s_peer_open(const unsigned char *handle64, void **ptr_dev) {
    if (!ptr_dev || !handle64) return fail(G4S_ERR_INVALID, "g4s_peer_open: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    G4S_CUDA(cudaIpcOpenMemHandle(ptr_dev, h, cudaIpcMemLazyEnablePeerAccess));
    return G4S_OK;
}
int g4s_peer_close(void *ptr_dev) {
    if (ptr_dev) G4S_CUDA(cudaIpcCloseMemHandle(ptr_dev));
    return G4S_OK;
}
int g4s_peer_free(void *ptr_dev) {
    if (ptr_dev) G4S_CUDA(cudaFree(ptr_dev));
    return G4S_OK;
}

int g4s_spmv_host(g4s_csr_t A, const double *x, double *y) {
    if (!A || (!x && A->cols) || (!y && A->rows)) return fail(G4S_ERR_INVALID, "g4s_spmv_host: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (A->rows == 0) return G4S_OK;
    return spmv_host_pipelined(A, x, y);
}

int g4s_spmv_csr_f64(int rows, int cols, const int *rowptr, const int *colids, const double *values,
                     const double *x, double *y) {
    g4s_csr_t h = nullptr;
    int rc = g4s_csr_create_host(&h, rows, cols, rowptr, colids, values);
    if (rc) return rc;
    rc = g4s_spmv_host(h, x, y);
    g4s_csr_destroy(h);
    return rc;
}

int g4s_spmv_cost(g4s_csr_t A, double *bytes, double *flops) {
    if (!A) return fail(G4S_ERR_INVALID, "null handle");
    if (bytes) *bytes = 12.0 * (double)A->nnz + 4.0 * ((double)A->rows + 1) + 8.0 * (double)A->cols + 8.0 * (double)A->rows;
    if (flops) *flops = 2.0 * (double)A->nnz;
    return G4S_OK;
}

int g4s_spmv_set_tuning(g4s_csr_t A, int lanes_per_row, int variant) {
    if (!A) return fail(G4S_ERR_INVALID, "null handle");
    if (lanes_per_row < 0 || lanes_per_row > 32 || (lanes_per_row & (lanes_per_row - 1)))
        return fail(G4S_ERR_INVALID, "lanes_per_row must be 0 or a power of two <= 32");
    A->plan.lanes_per_row = lanes_per_row;
    A->plan.variant = variant;
    return G4S_OK;
}

int g4s_partition_rows_i32(const int *work_prefix, int rows, int parts, int *cuts) {
    return partition_rows(work_prefix, rows, parts, cuts);
}
int g4s_partition_rows_i64(const long long *work_prefix, int rows, int parts, int *cuts) {
    return partition_rows(work_prefix, rows, parts, cuts);
}

void g4s_free(void *p) { free(p); }

}  // extern "C"
