// CitcomS's assembled stiffness operator ("node format") -> BSR with 3x3 blocks, so that
// BASELINE config 5 (g4s_bsr_spmm_device) runs on a real CitcomS matrix instead of a synthetic one.
//
// The reference keeps HALF of the symmetric matrix, 14 slots of 3 equations per node
// (construct_node_maps, citcoms/lib/Construct_arrays.c:264-310; coefficients Eqn_k1/2/3 filled by construct_node_ks,
// :330-456; the product n_assemble_del2_u, citcoms/lib/Element_calculations.c:516-565, applies every stored
// coefficient twice, once as K[e][c] and once as K[c][e]):
//   slot 0           the node itself:            K[3b+i][3b+k]  = Eqn_k{k+1}[i]        (i, k = 0..2)
//   slot ia = 1..13  lower-numbered neighbour n: K[3b+k][3n+d]  = Eqn_k{k+1}[3 ia + d] = K[3n+d][3b+k]
//   unused slots carry the dummy equation neq = 3 nno.
// Equation numbers are 3*(node-1)+direction (construct_id, Construct_arrays.c:156-160), so node b is block row b.
// The BSR built here stores BOTH triangles (the GPU kernel streams whole block rows and has no scatter), rows sorted by
// block column, values widened to fp64 (the reference's higher_precision is float, global_defs.h:116-120).
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "g4s_b200.h"

namespace g4s {
int fail(int status, const std::string &msg);
}
using g4s::fail;

namespace {

constexpr int kSlots = 14, kMaxEqn = 42;

template <class T>
int convert(int nno, const int *node_map, const T *k1, const T *k2, const T *k3, int *nnzb, int **browptr_out,
            int **bcolids_out, double **bvalues_out) {
    const long long neq = 3LL * nno;
    // pass 1 (all cores): validate the map, count the lower neighbours of every node and how often it is somebody's
    // neighbour, and find the largest distance between a node and a neighbour it names (bounds the scan of pass 2)
    std::vector<int> lower(nno, 0), upper(nno, 0);
    int first_bad = nno, reach = 0;
#pragma omp parallel for schedule(static) reduction(min : first_bad) reduction(max : reach)
    for (int b = 0; b < nno; ++b) {
        const int *c = node_map + (size_t)b * kMaxEqn;
        bool ok = true;
        for (int d = 0; d < 3; ++d) ok = ok && c[d] == 3 * b + d;
        int nl = 0;
        for (int ia = 1; ok && ia < kSlots; ++ia) {
            const int e0 = c[3 * ia];
            if (e0 == neq) {
                ok = c[3 * ia + 1] == neq && c[3 * ia + 2] == neq;
                continue;
            }
            ok = e0 >= 0 && e0 % 3 == 0 && c[3 * ia + 1] == e0 + 1 && c[3 * ia + 2] == e0 + 2 && e0 / 3 < b;
            if (!ok) break;
            ++nl;
            reach = std::max(reach, b - e0 / 3);
#pragma omp atomic
            ++upper[e0 / 3];
        }
        if (!ok) first_bad = std::min(first_bad, b);
        lower[b] = nl;
    }
    if (first_bad < nno) {  // name the defect of the first offending node (sequential re-check of that node only)
        const int b = first_bad;
        const int *c = node_map + (size_t)b * kMaxEqn;
        for (int d = 0; d < 3; ++d)
            if (c[d] != 3 * b + d)
                return fail(G4S_ERR_FORMAT, "node map: node " + std::to_string(b + 1) + " does not own equations 3(node-1)+d");
        for (int ia = 1; ia < kSlots; ++ia) {
            const int e0 = c[3 * ia];
            if (e0 == neq && (c[3 * ia + 1] != neq || c[3 * ia + 2] != neq))
                return fail(G4S_ERR_FORMAT, "node map: half-used slot at node " + std::to_string(b + 1));
            if (e0 != neq && (e0 < 0 || e0 % 3 != 0 || c[3 * ia + 1] != e0 + 1 || c[3 * ia + 2] != e0 + 2 || e0 / 3 >= b))
                return fail(G4S_ERR_FORMAT, "node map: slot " + std::to_string(ia) + " of node " + std::to_string(b + 1) +
                                                " is not the three equations of a lower-numbered node");
        }
        return fail(G4S_ERR_FORMAT, "node map: malformed entry at node " + std::to_string(b + 1));
    }
    long long total = 0;
    for (int b = 0; b < nno; ++b) total += 1 + lower[b] + upper[b];
    if (total > 2147483647LL) return fail(G4S_ERR_INVALID, "more than 2^31-1 blocks");
    int *rp = static_cast<int *>(malloc(sizeof(int) * ((size_t)nno + 1)));
    int *ci = static_cast<int *>(malloc(sizeof(int) * (size_t)std::max<long long>(total, 1)));
    double *va = static_cast<double *>(malloc(sizeof(double) * 9 * (size_t)std::max<long long>(total, 1)));
    if (!rp || !ci || !va) {
        free(rp);
        free(ci);
        free(va);
        return fail(G4S_ERR_ALLOC, "host allocation failed");
    }
    rp[0] = 0;
    for (int b = 0; b < nno; ++b) rp[b + 1] = rp[b] + 1 + lower[b] + upper[b];
    // pass 2 (all cores): block rows are dealt out in contiguous stretches.  The owner of rows [r0, r1) writes, for each of
    // them, [its lower neighbours, sorted][itself], and then walks the nodes e in (r0, r1 + reach) in ascending order and
    // appends the transposed block of every neighbour of e that lies in [r0, r1) — ascending e, so the upper part of a row
    // comes out sorted without a sort.
    const int nchunks = std::max(1, std::min(nno / 4096 + 1, omp_get_max_threads() * 4));
    int twice = nno;  // first node that lists a neighbour twice
#pragma omp parallel for schedule(dynamic, 1) reduction(min : twice)
    for (int ch = 0; ch < nchunks; ++ch) {
        const int r0 = (int)((long long)nno * ch / nchunks), r1 = (int)((long long)nno * (ch + 1) / nchunks);
        std::vector<int> cursor(r1 - r0);  // next free upper position of every owned row
        int order[kSlots];
        for (int b = r0; b < r1; ++b) {
            cursor[b - r0] = rp[b] + lower[b] + 1;
            const int *c = node_map + (size_t)b * kMaxEqn;
            const T *B[3] = {k1 + (size_t)b * kMaxEqn, k2 + (size_t)b * kMaxEqn, k3 + (size_t)b * kMaxEqn};
            int nl = 0;
            for (int ia = 1; ia < kSlots; ++ia)
                if (c[3 * ia] != neq) order[nl++] = ia;
            std::sort(order, order + nl, [&](int x, int y) { return c[3 * x] < c[3 * y]; });
            int pos = rp[b];
            for (int q = 0; q < nl; ++q, ++pos) {
                const int ia = order[q];
                if (q && c[3 * ia] == c[3 * order[q - 1]]) twice = std::min(twice, b);
                ci[pos] = c[3 * ia] / 3;
                double *blk = va + (size_t)pos * 9;  // K[3b+k][3n+d]
                for (int k = 0; k < 3; ++k)
                    for (int d = 0; d < 3; ++d) blk[3 * k + d] = (double)B[k][3 * ia + d];
            }
            ci[pos] = b;
            double *diag = va + (size_t)pos * 9;  // K[3b+i][3b+k] = Eqn_k{k+1}[i]
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) diag[3 * i + k] = (double)B[k][i];
        }
        const int e_end = (int)std::min<long long>(nno, (long long)r1 + reach);
        for (int e = r0 + 1; e < e_end; ++e) {
            const int *c = node_map + (size_t)e * kMaxEqn;
            const T *B[3] = {k1 + (size_t)e * kMaxEqn, k2 + (size_t)e * kMaxEqn, k3 + (size_t)e * kMaxEqn};
            for (int ia = 1; ia < kSlots; ++ia) {
                if (c[3 * ia] == neq) continue;
                const int n = c[3 * ia] / 3;
                if (n < r0 || n >= r1) continue;
                const int up = cursor[n - r0]++;
                ci[up] = e;
                double *tblk = va + (size_t)up * 9;  // K[3n+d][3e+k] = K[3e+k][3n+d]
                for (int k = 0; k < 3; ++k)
                    for (int d = 0; d < 3; ++d) tblk[3 * d + k] = (double)B[k][3 * ia + d];
            }
        }
    }
    if (twice < nno) {
        free(rp);
        free(ci);
        free(va);
        return fail(G4S_ERR_FORMAT, "node map: node " + std::to_string(twice + 1) + " lists a neighbour twice");
    }
    *nnzb = (int)total;
    *browptr_out = rp;
    *bcolids_out = ci;
    *bvalues_out = va;
    return G4S_OK;
}

}  // namespace

extern "C" {

int g4s_bsr_from_citcoms_nodes(int nno, const int *node_map, const void *eqn_k1, const void *eqn_k2, const void *eqn_k3,
                               int value_bytes, int *nnzb, int **browptr, int **bcolids, double **bvalues) {
    if (nno < 0 || !node_map || !eqn_k1 || !eqn_k2 || !eqn_k3 || !nnzb || !browptr || !bcolids || !bvalues)
        return fail(G4S_ERR_INVALID, "g4s_bsr_from_citcoms_nodes: bad arguments");
    if (nno > 715827882) return fail(G4S_ERR_INVALID, "g4s_bsr_from_citcoms_nodes: 3*nno exceeds int32 equation numbers");
    if (value_bytes == 4)
        return convert<float>(nno, node_map, static_cast<const float *>(eqn_k1), static_cast<const float *>(eqn_k2),
                              static_cast<const float *>(eqn_k3), nnzb, browptr, bcolids, bvalues);
    if (value_bytes == 8)
        return convert<double>(nno, node_map, static_cast<const double *>(eqn_k1), static_cast<const double *>(eqn_k2),
                               static_cast<const double *>(eqn_k3), nnzb, browptr, bcolids, bvalues);
    return fail(G4S_ERR_INVALID, "g4s_bsr_from_citcoms_nodes: value_bytes must be 4 (float, the reference's higher_precision) or 8");
}

}  // extern "C"
