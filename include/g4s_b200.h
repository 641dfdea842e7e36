/* g4s_b200 — C ABI of the B200-native mv/ (matrix-vector) and mm/ (matrix-matrix) hot path of G4S.
 *
 * Plain C: pointers, sizes and opaque handles only; no C++ or torch types cross this boundary.
 * Every entry names the reference interface it replaces (paths relative to the reference tree).
 * All functions return G4S_OK (0) or a negative g4s_status and never throw or abort; the text of the last
 * error on the calling thread is available from g4s_last_error().  There is NO CPU fallback: every compute
 * entry fails with G4S_ERR_CUDA when no sm_100 device is usable.
 *
 * Conventions kept from the reference (mm/inc/CSR.h:22-100): CSR<int,double>, 0-based, rowptr[rows+1],
 * colids[nnz], values[nnz]; SpGEMM output has column ids sorted ascending inside each row (the reference's
 * mkl_sparse_order, mm/inc/mkl_mult.h:70, and HashSpGEMM<sortOutput=true>, mm/inc/hash_mult.h:525-553).
 *
 * Pointer spaces: `_host` entries take host memory and do their own H2D/D2H copies; `_device` entries take
 * device memory on the current CUDA device (16-byte aligned bases) and a cudaStream_t passed as void*.
 */
#ifndef G4S_B200_H
#define G4S_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum g4s_status {
    G4S_OK = 0,
    G4S_ERR_INVALID = -1, /* bad argument (null pointer, negative size, misaligned device pointer) */
    G4S_ERR_CUDA = -2,    /* CUDA runtime error or no usable device */
    G4S_ERR_ALLOC = -3,   /* host or device allocation failed */
    G4S_ERR_IO = -4,      /* file missing or unreadable */
    G4S_ERR_FORMAT = -5,  /* malformed MatrixMarket / edge-list input (the reference throws std::runtime_error) */
    G4S_ERR_SHAPE = -6    /* non-conformable operands */
} g4s_status;

const char *g4s_last_error(void);
const char *g4s_version(void);
/* Number of kernels this library has launched on behalf of the calling process (bench.py's gpu_launches). */
long long g4s_kernel_launch_count(void);
int g4s_device_count(void);
int g4s_set_device(int device);

/* ------------------------------------------------------------------------------------------------------
 * Timings — C twin of class Timings (mm/inc/Timings.h:4-22): two bools then seven doubles, seconds.
 * Layout-compatible with the reference class (standard layout, no virtuals), so a `Timings&` can be passed
 * as `g4s_timings*` from C++.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct g4s_timings {
    unsigned char measure_separate;
    unsigned char measure_total;
    double create;     /* upload / handle creation            (mkl_sparse_d_create_csr, mkl_mult.h:50-53) */
    double spmm;       /* symbolic + numeric product           (mkl_sparse_spmm,        mkl_mult.h:58)    */
    double convert;    /* always 0: output is CSR already      (mkl_sparse_convert_csr, mkl_mult.h:64)    */
    double order;      /* in-kernel column sort, counted in spmm; 0 here (mkl_sparse_order, mkl_mult.h:70) */
    double export_csr; /* D2H of crpt/ccol/cval                (mkl_sparse_d_export_csr, mkl_mult.h:79)   */
    double destroy;    /* device frees                          (mkl_sparse_destroy,     mkl_mult.h:102-106) */
    double total;
} g4s_timings;
void g4s_timings_init(g4s_timings *t);                           /* Timings::Timings,  mm/src/Timings.cpp:4-14  */
void g4s_timings_add(g4s_timings *acc, const g4s_timings *t);    /* operator+=,        Timings.cpp:16-24        */
void g4s_timings_div(g4s_timings *t, double x);                  /* operator/=,        Timings.cpp:26-34        */
void g4s_timings_print(const g4s_timings *t, double total_flop); /* Timings::print,    Timings.cpp:36-60        */

/* ------------------------------------------------------------------------------------------------------
 * CSR handle — device-resident CSR<int,double> plus the SpMV inspector data (merge-path tile table).
 * Plays the role of MKL's sparse_matrix_t in the reference (mkl_sparse_d_create_csr / mkl_sparse_destroy,
 * mm/inc/mkl_mult.h:50-53,102-106).  NEW API SURFACE: the reference has no sparse mat-vec (mv/mv.c is dense).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct g4s_csr *g4s_csr_t;

/* copies host arrays to the device (owned by the handle) */
int g4s_csr_create_host(g4s_csr_t *out, int rows, int cols, const int *rowptr, const int *colids,
                        const double *values);
/* borrows device arrays (caller keeps them alive); nnz = rowptr[rows] is read back once */
int g4s_csr_create_device(g4s_csr_t *out, int rows, int cols, const int *rowptr_dev, const int *colids_dev,
                          const double *values_dev, void *stream);
int g4s_csr_destroy(g4s_csr_t h);
int g4s_csr_shape(g4s_csr_t h, int *rows, int *cols, long long *nnz);
/* device pointers of the handle's arrays (for callers that want to run their own kernels / collectives) */
int g4s_csr_device_arrays(g4s_csr_t h, const int **rowptr_dev, const int **colids_dev, const double **values_dev);
/* copy the handle's arrays back to host buffers sized rows+1 / nnz / nnz */
int g4s_csr_download(g4s_csr_t h, int *rowptr, int *colids, double *values);

/* y = A x  (alpha = 1, beta = 0 — the only combination mv/mv.c:6-27 uses).
 * Replaces, for sparse A, matrix_multiply_dgemv (mv/mv.c:23-27). */
int g4s_spmv_device(g4s_csr_t A, const double *x_dev, double *y_dev, void *stream);
int g4s_spmv_host(g4s_csr_t A, const double *x, double *y); /* H2D x, kernel, D2H y, synchronous */
/* one-shot: create + spmv + destroy, all arrays on the host */
int g4s_spmv_csr_f64(int rows, int cols, const int *rowptr, const int *colids, const double *values,
                     const double *x, double *y);
/* Algorithmic bytes and flops of one SpMV: 12 nnz + 4(rows+1) + 8 cols + 8 rows; 2 nnz. */
int g4s_spmv_cost(g4s_csr_t A, double *bytes, double *flops);
/* Tuning knobs (0 = automatic): lanes cooperating on one row (1..32, power of two), kernel variant. */
int g4s_spmv_set_tuning(g4s_csr_t A, int lanes_per_row, int variant);

/* Extended form used by the multi-GPU path, where each GPU holds its row block split into a diagonal block
 * (columns it owns) and a row-compressed off-diagonal block: local row r of A is written to y[row_map[r]]
 * (row_map_dev NULL = identity) and accumulate != 0 turns the store into y += A x. */
int g4s_spmv_device_ex(g4s_csr_t A, const double *x_dev, double *y_dev, const int *row_map_dev, int accumulate,
                       void *stream);

/* ------------------------------------------------------------------------------------------------------
 * Dense mv entry points, same names and signatures as the reference (mv/mv.c:6-27); host buffers,
 * `dim x dim` doubles.  A is interpreted exactly as the reference's CBLAS calls do (column-major view of
 * the row-major-filled buffer, upper triangle):
 *   dgemv : C = A_cm B            dsymv : C = sym(upper(A_cm)) B
 *   dtrmv : B <- upper(A_cm)^T B  (C untouched)      sspmv : C = sympacked(A) B  (cblas_dspmv on A as packed)
 * They upload A, run one HBM-bound kernel and download the result; failures are reported on stderr
 * (the reference signatures return void) and leave outputs untouched.
 * ---------------------------------------------------------------------------------------------------- */
void matrix_multiply_dgemv(double *A, double *B, double *C, int dim);
void matrix_multiply_dsymv(double *A, double *B, double *C, int dim);
void matrix_multiply_dtrmv(double *A, double *B, double *C, int dim);
void matrix_multiply_sspmv(double *A, double *B, double *C, int dim);
/* status-returning twins on device memory */
int g4s_dense_mv_device(int op /*0 dgemv,1 dsymv,2 dtrmv,3 dspmv*/, const double *A_dev, double *B_dev,
                        double *C_dev, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * SpGEMM  C = A B, CSR x CSR -> CSR, sorted columns.
 * ---------------------------------------------------------------------------------------------------- */
/* Allocator the caller's free routine matches (the reference frees mkl()'s outputs with my_free,
 * mm/inc/CSR.h:50-62 / utility.h:141-153).  NULL selects malloc (release with g4s_free). */
typedef void *(*g4s_alloc_fn)(size_t bytes, void *ctx);

/* C twin of  void mkl(int*arpt,int*acol,double*aval,int*brpt,int*bcol,double*bval,
 *                     int**crpt_,int**ccol_,double**cval_,int M,int K,int N,int*cnnz_,Timings&timing)
 * (mm/inc/mkl_mult.h:40-110): host CSR in, callee-allocated host CSR out, crpt has M+1 entries with
 * crpt[M] = cnnz, phases reported in `timing` (may be NULL). */
int g4s_mkl(const int *arpt, const int *acol, const double *aval, const int *brpt, const int *bcol,
            const double *bval, int **crpt_, int **ccol_, double **cval_, int M, int K, int N, int *cnnz_,
            g4s_timings *timing);
int g4s_mkl_alloc(const int *arpt, const int *acol, const double *aval, const int *brpt, const int *bcol,
                  const double *bval, int **crpt_, int **ccol_, double **cval_, int M, int K, int N, int *cnnz_,
                  g4s_timings *timing, g4s_alloc_fn alloc_index, g4s_alloc_fn alloc_value, void *alloc_ctx);
void g4s_free(void *p);

/* Device-resident SpGEMM on handles (HashSpGEMM(a,b,c,...), mm/inc/hash_mult.h:1028-1057): returns a new
 * handle that owns C.  Symbolic and numeric phases run on `stream`; the call synchronises once, after the
 * symbolic phase, to size C. */
int g4s_spgemm_device(g4s_csr_t A, g4s_csr_t B, g4s_csr_t *C, void *stream);
/* Per-phase device milliseconds of the last g4s_spgemm_device on this thread: binning, symbolic, scan+alloc,
 * numeric(+sort). */
int g4s_spgemm_last_phase_ms(double *ms4);
/* phase timing costs five event records per product and is off by default: switch it on for the products you want reported */
int g4s_spgemm_set_phase_timing(int on);
/* The same product by expand - sort - compress, the GPU form of the reference's OuterSpGEMM "join"
 * (mm/inc/outer_mult.h:271-542: (row, col, value) triples, radix sort on (row << 32) | col, equal keys summed).  Same CSR
 * as g4s_spgemm_device; the values are summed in the reference's sequential order for EVERY row (bit-identical to
 * HashSpGEMM<false,true>).  Slower and 32 bytes of scratch per intermediate product (at most 2^31-1 of them): the
 * library's second, independent SpGEMM, for cross-checks and for the reference's algorithm inventory. */
int g4s_spgemm_esc_device(g4s_csr_t A, g4s_csr_t B, g4s_csr_t *C, void *stream);
/* HeapSpGEMM (mm/inc/heap_mult.h:47-223): C = A B by a k-way heap merge of the sorted rows of B, rows of A of any length;
 * same sorted CSR, values bit-identical to HashSpGEMM<false,true> (ties on a column leave the heap in A's stored order).
 * G4S_ERR_INVALID when a row of B is not sorted by column. */
int g4s_spgemm_heap_device(g4s_csr_t A, g4s_csr_t B, g4s_csr_t *C, void *stream);

/* compute_flop / get_flop (mm/inc/mkl_mult.h:8-38, hash_mult.h:45-62): intermediate products, 64-bit.
 * row_work_dev may be NULL; else int32[rows] on the device (BIN::set_intprod_num, BIN.h:77-95). */
int g4s_compute_flop_device(g4s_csr_t A, g4s_csr_t B, long long *total, int *row_work_dev, void *stream);
long long compute_flop_host(const int *arpt, const int *acol, const int *brpt, int M); /* host twin */

/* ------------------------------------------------------------------------------------------------------
 * Partitioner — BIN::set_rows_offset (mm/inc/BIN.h:100-122) generalised from threads to GPUs: contiguous
 * row ranges of (nearly) equal work.  work_prefix = exclusive prefix sum with rows+1 entries (rowptr itself
 * for an nnz balance).  Cut p = lower_bound(prefix, ceil(total/parts)*(p+1)), last cut = rows; 64-bit
 * arithmetic, cuts clamped to `rows` (the reference's int arithmetic overflows, and its cuts can pass the
 * last row for tiny inputs).
 * ---------------------------------------------------------------------------------------------------- */
int g4s_partition_rows_i32(const int *work_prefix, int rows, int parts, int *cuts /* parts+1 */);
int g4s_partition_rows_i64(const long long *work_prefix, int rows, int parts, int *cuts);

/* Multi-GPU building blocks.  A GPU's row block A[r0:r1, :] (global column ids) is split into
 * the DIAGONAL block (columns [c0,c1) it owns, rebased to 0) and the OFF-DIAGONAL block (every other column),
 * which is row-compressed: only rows holding an off-diagonal entry are kept; g4s_csr_row_map gives the map back to
 * local rows and g4s_spmv_device_ex on that handle scatters through it.  g4s_csr_compact_columns returns the
 * sorted distinct columns a block references (the x entries to fetch from other GPUs; release with
 * g4s_device_free) and rewrites the block's column ids to positions in that list.  g4s_gather_f64 packs
 * dst[k] = src[idx[k]].  The exchange itself (NCCL) is the caller's, see g4s_b200/dist.py. */
int g4s_csr_split_columns(g4s_csr_t A, int c0, int c1, g4s_csr_t *diag, g4s_csr_t *off, void *stream);
/* Fused alternative (one NVSwitch box, world <= 8): y = A x in ONE kernel, where A is the rank's row block with
 * GLOBAL column ids and x is partitioned: x_parts[q] (host array of `world` device pointers this GPU can
 * dereference: its own slice for q == self, CUDA-IPC-mapped peer memory otherwise) holds x[cuts[q] .. cuts[q+1]).
 * Entries owned by other GPUs are loaded over NVLink from inside the SpMV kernel; no exchange step, no second
 * kernel.  Ordering against the owners' writes is either the caller's (ready_flags_dev == NULL: run a barrier first)
 * or in-kernel: ready_flags_dev points at this rank's array of `world` 64-bit flags in peer-shared memory, into
 * which rank q stores `epoch` with g4s_peer_signal once its slice for this product is written; only the chunks
 * that touch another GPU's slice wait (acquire) for flags >= epoch, the rest of the matrix streams meanwhile.
 * signal_arrays (host array of `world` peer-mapped flag arrays, as for g4s_peer_signal) non-NULL folds the signal into
 * the product: the kernel publishes `epoch` for this rank's slice when it starts, so a step is ONE launch. */
int g4s_spmv_partitioned_device(g4s_csr_t A, int world, int self, const double *const *x_parts, const int *cuts,
                                double *y_dev, const unsigned long long *ready_flags_dev, unsigned long long epoch,
                                unsigned long long *const *signal_arrays, void *stream);
/* After the first partitioned product on an owned handle (which numbers the remote columns and rewrites their ids in
 * place): how many x entries the rank pulls from peers per product (-1: handle not localized) and the bit mask of the
 * ranks that own them.  A product waits for the flags of the ranks in the mask only; the caller widens it to the ranks
 * that read ITS slice when the pattern is not structurally symmetric (g4s_b200/dist.py does, with one all-gather). */
int g4s_spmv_partition_info(g4s_csr_t A, int *n_remote_columns, unsigned *owner_mask);
int g4s_spmv_partition_set_wait_mask(g4s_csr_t A, unsigned mask);
/* flag_arrays[q] = rank q's flag array (own or IPC-mapped).  Stream-ordered after the writes of the x slice:
 * stores `epoch` into slot `self` of every rank's array (release, system scope). */
int g4s_peer_signal(unsigned long long *const *flag_arrays, int world, int self, unsigned long long epoch,
                    void *stream);
/* Peer memory: cudaMalloc + CUDA IPC handle (64 bytes) / open a peer's handle / close / free. */
int g4s_peer_alloc(size_t bytes, void **ptr_dev, unsigned char *handle64);
int g4s_peer_open(const unsigned char *handle64, void **ptr_dev);
int g4s_peer_close(void *ptr_dev);
int g4s_peer_free(void *ptr_dev);
int g4s_csr_row_map(g4s_csr_t h, const int **row_map_dev, int *full_rows);
int g4s_csr_compact_columns(g4s_csr_t A, int **needed_cols_dev, int *n_needed, void *stream);
int g4s_gather_f64(double *dst_dev, const double *src_dev, const int *idx_dev, long long n, void *stream);
int g4s_device_free(void *p);

/* ------------------------------------------------------------------------------------------------------
 * Loaders (host) — CSR::construct (mm/inc/CSR.h:485-669) and CSR(graph&) (mm/inc/CSR.h:255-329).
 * Outputs are malloc'd (g4s_free).
 * ---------------------------------------------------------------------------------------------------- */
int g4s_csr_read_matrix_market(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids,
                               double **values);
/* Binary CSR cache (no counterpart in the reference, whose construct() re-parses the text file with
 * `ifstream >>` on every run, mm/inc/CSR.h:526-553): a 64-byte header (magic "G4SCSR1", shape, checksum) followed by the
 * int32 rowptr / colids and fp64 values exactly as construct() produced them.  read_binary validates magic, sizes, row
 * pointers and checksum and returns G4S_ERR_FORMAT for anything else.  read_cached is the drop-in for construct(): it
 * loads `<mtx_path>.g4scsr` when that file exists and is not older than the text file, otherwise parses the text file
 * and (best effort) writes the cache; *cache_hit (optional) reports which happened. */
int g4s_csr_write_binary(const char *path, int rows, int cols, int nnz, const int *rowptr, const int *colids,
                         const double *values);
int g4s_csr_read_binary(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids, double **values);
int g4s_csr_read_cached(const char *mtx_path, int *rows, int *cols, int *nnz, int **rowptr, int **colids, double **values,
                        int *cache_hit);
/* edge list laid out as class graph (mm/inc/graph.h:4-25): long start[m], end[m]; double w[m]; n vertices */
int g4s_csr_from_edge_list(long m, long n, const long *start, const long *end, const double *w, int *nnz,
                           int **rowptr, int **colids, double **values);
/* CSR(const CSR&, M_, N_, M_start, N_start) (mm/inc/CSR.h:691-733) */
int g4s_csr_submatrix(int rows, int cols, const int *rowptr, const int *colids, const double *values, int M_,
                      int N_, int M_start, int N_start, int *nnz, int **orpt, int **ocol, double **oval);

/* the same with this rank's block rows taken in the order row_order_dev[0..mb_local) (local row numbers; see
 * g4s_bsr3_spmm64_ordered_device) */
int g4s_bsr3_spmm64_partitioned_ordered_device(int mb_local, const int *browptr_dev, const int *bcolids_dev,
                                               const double *bvalues_dev, int world, const double *const *B_parts,
                                               const int *cuts, double *C_dev, const int *row_order_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * Synthetic inputs of BASELINE.json's configs, generated directly in CSR on the device.
 * Natural ordering, row = (k*n + j)*n + i.   2-D 5-point: diag 4, off-diag -1.  3-D 27-point: diag 26,
 * off-diag -1.  Rows [row0, row1) only (row-partitioned generation for multi-GPU); column ids stay global.
 * ---------------------------------------------------------------------------------------------------- */
long long g4s_laplacian2d_nnz(int n, long long row0, long long row1);
long long g4s_laplacian3d27_nnz(int n, long long row0, long long row1);
int g4s_csr_generate_laplacian2d(g4s_csr_t *out, int n, long long row0, long long row1, void *stream);
int g4s_csr_generate_laplacian3d27(g4s_csr_t *out, int n, long long row0, long long row1, void *stream);
/* Graph500 R-MAT (a,b,c,d = .57,.19,.19,.05), 2^scale vertices, edge_factor * 2^scale generated edges,
 * weights U(0,1), duplicates summed (CSR(graph&) semantics); counter-based generator keyed by `seed`. */
int g4s_csr_generate_rmat(g4s_csr_t *out, int scale, int edge_factor, unsigned long long seed, void *stream);
/* the edge list alone, laid out as class graph (long start[m], end[m]; double w[m]) in device memory */
int g4s_rmat_edges_device(int scale, long long m, unsigned long long seed, long *start_dev, long *end_dev,
                          double *w_dev, void *stream);
/* Device graph-to-CSR loader: CSR(graph&) (mm/inc/CSR.h:255-329) for an edge list already on the GPU.  Edges are
 * ordered by (start, end) and equal pairs summed; the input need not be grouped by start vertex. */
int g4s_csr_from_edges_device(long m, long n, const long *start_dev, const long *end_dev, const double *w_dev,
                              g4s_csr_t *out, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * BSR SpMM (BASELINE config 5): C[mb*bs, ncol] = A_bsr B, bs x bs row-major blocks, row-major dense B, C.
 * No counterpart in the reference; FP64 tensor-core (DMMA) path for bs = 3.
 * ---------------------------------------------------------------------------------------------------- */
int g4s_bsr_spmm_device(int mb, int kb, int bs, const int *browptr_dev, const int *bcolids_dev,
                        const double *bvalues_dev, int ncol, const double *B_dev, double *C_dev, void *stream);
/* CitcomS's assembled stiffness operator -> BSR 3x3, host.  Input is the reference's half-stored
 * node format: node_map[nno*42] (construct_node_maps, citcoms/lib/Construct_arrays.c:264-310) and the coefficient arrays
 * Eqn_k1/2/3[nno*42] (construct_node_ks, :330-456; value_bytes = 4 for the reference's `higher_precision` float,
 * global_defs.h:116-120, or 8 for double).  Output (malloc'd, g4s_free): both triangles of the symmetric matrix as
 * nno block rows sorted by block column, fp64 row-major 3x3 blocks, such that g4s_bsr_spmm_device applies the operator
 * of n_assemble_del2_u (citcoms/lib/Element_calculations.c:516-565).  G4S_ERR_FORMAT for a map that is not in that
 * format (a node not owning equations 3(node-1)+d, a slot that is not a lower-numbered node, a neighbour listed twice). */
int g4s_bsr_from_citcoms_nodes(int nno, const int *node_map, const void *eqn_k1, const void *eqn_k2, const void *eqn_k3,
                               int value_bytes, int *nnzb, int **browptr, int **bcolids, double **bvalues);
/* kernel choice for bs = 3, ncol = 64: 0 automatic, 1 DFMA, 2 DMMA (FP64 tensor cores), 3 generic, 4 K-packed DFMA */
int g4s_bsr_spmm_set_variant(int variant);
/* Ordered form for bs = 3, ncol = 64: block rows are processed in the order row_order_dev[0..mb) (a permutation of the
 * block rows; the result is the same C, only the schedule changes), cut into tiles [tile_ptr_dev[t], tile_ptr_dev[t+1])
 * that are handed to the SMs round-robin (tile_ptr_dev may be null: one contiguous stretch of the order per SM).  With a
 * tile-major order of a structured mesh the block rows in flight on one SM share their rows of B, which are then served
 * by that SM's L1, and neighbouring tiles are in flight on other SMs at the same time. */
int g4s_bsr3_spmm64_ordered_device(int mb, int kb, const int *browptr_dev, const int *bcolids_dev, const double *bvalues_dev,
                                   const double *B_dev, double *C_dev, const int *row_order_dev, const int *tile_ptr_dev,
                                   int ntiles, void *stream);
/* Tile-major order of an n0 x n1 x n2 grid numbered n0 fastest (node = (k*n1 + j)*n0 + i): p0 x p1 patches of the first
 * two axes, each swept along the third axis (one tile).  order (host) receives n0*n1*n2 node numbers; tile_ptr (host,
 * optional) ceil(n0/p0)*ceil(n1/p1) + 1 offsets into it; *ntiles (optional) the number of tiles. */
int g4s_grid_pencil_order(int n0, int n1, int n2, int p0, int p1, int *order, int *tile_ptr, int *ntiles);
/* Multi-GPU form for bs = 3, ncol = 64 (one NVSwitch box, world <= 8): this rank's mb_local block rows with GLOBAL block
 * column ids; B is row-partitioned by `cuts` (block rows) and B_parts[q] points at rank q's slice (own memory or
 * CUDA-IPC peer memory, see g4s_peer_alloc): rows of B owned by other GPUs are read over NVLink inside the kernel. */
int g4s_bsr3_spmm64_partitioned_device(int mb_local, const int *browptr_dev, const int *bcolids_dev,
                                       const double *bvalues_dev, int world, const double *const *B_parts,
                                       const int *cuts, double *C_dev, void *stream);

/* Inspector / executor form for bs = 3, ncol = 64 (the role of MKL's mkl_sparse_set_mm_hint + mkl_sparse_optimize):
 * a sliding-window sweep (csrc/bsr_sweep.cu).  The block rows are handed over as STRIPS — lists of block rows that one
 * lane group sweeps in order while it keeps three consecutive rows of the strip in registers, so that a row of B
 * is loaded once for all rows of the window that reference it.  Good strips are grid lines of a mesh
 * (g4s_grid_pencil_strips) or, for any banded matrix in its natural order, runs of consecutive rows (nstrips = 0:
 * runs of 64).  16 consecutive strips form the tile of one CTA and should be neighbours in the mesh.
 *   create      : browptr_dev / bcolids_dev are DEVICE arrays (as everywhere in this section); strip_ptr[nstrips+1] and
 *                 strip_rows[mb] are HOST arrays and must cover every block row exactly once.  world > 1: this rank's
 *                 rows with GLOBAL block-column ids, B row-partitioned by cuts[world+1] (see
 *                 g4s_bsr3_spmm64_partitioned_device).  Builds the schedule on the host (OpenMP) and allocates the plan
 *                 stream on the device (about 81 bytes per block).
 *   set_values  : (re)packs the 3x3 blocks into the plan stream; call after create and whenever the values change.
 *   spmm64      : C = A B.  B must be finite: blocks absent from a step are stored as explicit zeros.
 *   info        : slot_fill = blocks / block slots moved (1.0: every loaded row of B feeds three blocks; a plan with a
 *                 low fill is slower than g4s_bsr_spmm_device and should not be used). */
typedef struct g4s_bsr_plan *g4s_bsr_plan_t;
int g4s_bsr3_plan_create(g4s_bsr_plan_t *out, int mb, int kb, const int *browptr_dev, const int *bcolids_dev, int nstrips,
                         const int *strip_ptr, const int *strip_rows, int world, const int *cuts);
int g4s_bsr3_plan_set_values(g4s_bsr_plan_t plan, const double *bvalues_dev, void *stream);
int g4s_bsr3_plan_spmm64_device(g4s_bsr_plan_t plan, const double *B_dev, double *C_dev, void *stream);
int g4s_bsr3_plan_spmm64_partitioned_device(g4s_bsr_plan_t plan, int world, const double *const *B_parts, const int *cuts,
                                            double *C_dev, void *stream);
int g4s_bsr3_plan_info(g4s_bsr_plan_t plan, double *slot_fill, long long *stream_bytes, int *nstages, int *ntiles,
                       int *stage_smem_bytes);
int g4s_bsr3_plan_destroy(g4s_bsr_plan_t plan);
/* The schedule alone, built from HOST arrays without touching a device (what the CPU tests replay), for a launch of
 * `grid` CTAs at ctas_per_sm (1 or 2) per SM: per stage, in execution order, the table entry (byte offset in the plan
 * stream, chunk bytes | expected bytes << 32) and the 64-int row of copy descriptors, per CTA its first stage, the
 * concatenated stage headers + position lists (`meta`, stage q at meta_off[q]), and per block the index (in doubles) of its
 * slot in the plan stream.  Output arrays are malloc'd (g4s_free); any of the pointers may be null. */
int g4s_bsr3_plan_inspect_host(int mb, int kb, const int *browptr, const int *bcolids, int nstrips, const int *strip_ptr,
                               const int *strip_rows, int world, const int *cuts, int grid, int ctas_per_sm, int *nstages,
                               int *ntiles, long long *stream_bytes, int *stage_smem_bytes, double *slot_fill,
                               long long **stage_table, int **cta_ptr, int **prod, int **meta, long long **meta_off,
                               long long **base);
/* Strips for the nodes k_begin <= k < k_end of an n0 x n1 x n2 grid numbered n0 fastest, as LOCAL row numbers
 * ((k - k_begin)*n1 + j)*n0 + i: one strip per grid line along the third axis, the lines of a p0 x p1 patch consecutive
 * (p0 = p1 = 4: one patch per tile).  strip_ptr (host) receives n0*n1 + 1 offsets, strip_rows n0*n1*(k_end-k_begin) rows. */
int g4s_grid_pencil_strips(int n0, int n1, int k_begin, int k_end, int p0, int p1, int *strip_ptr, int *strip_rows);

/* ------------------------------------------------------------------------------------------------------
 * The G4S graph engine ABI.  `spmm_dense` is the engine CitcomS calls through E->spmm_dense
 * (citcoms/bin/Citcom.c:45-48,93; citcoms/lib/global_defs.h:48-49,854-857); the reference declares it but does not
 * ship its body.  Semantics per GraphProcess (deepmd/source/op/graph.h:21-32): for every vertex, gather() for each of
 * its `degree` neighbours, then apply().  The callbacks are host functions, so this entry runs on the CPU by
 * construction; elapsed seconds are added to *time.
 * g4s_ebe_matvec_device is the callback-free device form of the engine's one concrete instance, CitcomS's
 * element-by-element operator (gather at citcoms/lib/Element_calculations.c:453-471): Au[dofs[e][r]] +=
 * sum_c elt_k[e][r*ndof + c] * u[dofs[e][c]] for every element e; ndof = 24 / 8 / 4 (loc_mat_size); Au is accumulated
 * into (the caller zeroes it, as e_assemble_del2_u does at Element_calculations.c:494-495).
 * ---------------------------------------------------------------------------------------------------- */
typedef void (*g4s_fun_gather)(int, int, const double **, const double *, double *);
typedef void (*g4s_fun_apply)(int, const double **, const double *, double *);
void spmm_dense(uint32_t numNodes, uint32_t degree, const double **edgeWeight, const double *vertexStates,
                double *temp, double *result, g4s_fun_gather gather, g4s_fun_apply apply, double *time,
                int threadNum);
int g4s_ebe_matvec_device(int nel, int ndof, const double *elt_k_dev, const int *elem_dofs_dev, const double *u_dev,
                          double *Au_dev, void *stream);


/* Stable radix sort of n (64-bit key, 8-byte value) pairs by the low key_bits bits of the key, on the device, in place
 * (the *_tmp arrays are scratch of the same size).  The library's own sort: counterpart of radix_sort(begin, end, buf, key)
 * (mm/inc/radix_sort.h:701-705) as the join SpGEMM calls it (mm/inc/outer_mult.h:427-442); g4s_spgemm_esc_device and
 * g4s_csr_from_edges_device sort with it. */
int g4s_radix_sort_pairs_device(unsigned long long *keys_dev, unsigned long long *keys_tmp_dev, void *values_dev,
                                void *values_tmp_dev, long long n, int key_bits, void *stream);

/* Exclusive prefix sum of n int32 counts on the device: out[k] = in[0] + ... + in[k-1], and out[n] = the total when
 * write_total is non-zero (out then has n + 1 entries); out may alias in.  The reference's scan(in, out, N)
 * (mm/inc/utility.h:166-209), as hash_symbolic calls it to turn row_nz into C's row pointers (mm/inc/hash_mult.h:506-507).
 * Three launches (tile sums, their scan, apply) over 2048-entry tiles.  *total_host, when non-NULL, receives the 64-bit total and the call
 * synchronises `stream`; the int32 outputs are only meaningful while the total fits. */
int g4s_exclusive_scan_i32_device(const int *in_dev, int *out_dev, long long n, int write_total, long long *total_host,
                                  void *stream);

/* ------------------------------------------------------------------------------------------------------
 * OptMatmul: the dense fp64 product DeePMD-kit routes through the G4S engine,
 * res[M,K] = xx[M,N] w[N,K], all row-major (deepmd/source/op/opt_matmul.cc:24-62; engine loop
 * deepmd/source/op/graph.h:21-32: one vertex per row of xx, degree K, gather = a dot product of length N), and the two
 * products of its registered gradient (deepmd/source/op/_opt_matmul_grad.py): dxx[M,N] = grad[M,K] w^T,
 * dw[N,K] = xx^T grad (either output may be NULL).  FP64 tensor cores (DMMA) on the GPU.  g4s_opt_matmul takes host
 * pointers (what OptMatmulOp::Compute holds: tensor.flat<double>().data()) and copies inside the call.
 * ---------------------------------------------------------------------------------------------------- */
int g4s_opt_matmul_device(int M, int N, int K, const double *xx_dev, const double *w_dev, double *res_dev, void *stream);
int g4s_opt_matmul_grad_device(int M, int N, int K, const double *xx_dev, const double *w_dev, const double *grad_dev,
                               double *dxx_dev, double *dw_dev, void *stream);
int g4s_opt_matmul(int M, int N, int K, const double *xx, const double *w, double *res);

#ifdef __cplusplus
}
#endif
#endif /* G4S_B200_H */
