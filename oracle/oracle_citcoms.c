/* TEST INFRASTRUCTURE ONLY — CPU restatement of CitcomS's assembled ("node format") stiffness operator, the
 * reference-side data format of SURVEY.md §8f row 2.  PARITY UNPINNED: the citcoms tree needs MPI and its own build
 * system and is not buildable here, and it ships no test for these routines; every function restates the cited
 * lines literally (1-based node numbers kept inside, 0-based equation numbers as the reference's ID array) and the
 * set is cross-checked against an independently assembled sparse matrix in tests/test_citcoms_cpu.py.
 *
 * Storage (citcoms/lib/Construct_arrays.c:264-312): per node nn, max_eqn = 14*dims = 42 slots.
 *   Node_map[(nn-1)*42 + 0..2]        equations of node nn itself
 *   Node_map[(nn-1)*42 + 3*ia + d]    equation d of the ia-th LOWER-numbered neighbour ja < nn (ia = 1..13)
 *   unused slots hold neq (a dummy equation: u[neq] = 0)
 *   Eqn_k1/2/3[slot] = K[eqn_1/2/3 of nn][Node_map[slot]]   (half of the symmetric matrix, higher_precision = float,
 *                                                            citcoms/lib/global_defs.h:116-120)
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* integer coordinates of element node rr relative to node 1, (z, x, y) (citcoms/lib/element_definitions.h:198-208) */
static const int citcoms_offset[9][3] = {{0, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {0, 0, 1},
                                         {1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}};

/* construct_ien (citcoms/lib/Construct_arrays.c:45-81): ien[(element-1)*8 + rr-1] = 1-based node number; nodes are
 * numbered z fastest, then x, then y.  Returns the number of elements. */
int oracle_citcoms_ien(int nox, int noy, int noz, int *ien) {
    const int elx = nox - 1, ely = noy - 1, elz = noz - 1;
    for (int r = 1; r <= ely; r++)
        for (int q = 1; q <= elx; q++)
            for (int p = 1; p <= elz; p++) {
                const int element = (r - 1) * elx * elz + (q - 1) * elz + p;
                const int start = (r - 1) * noz * nox + (q - 1) * noz + p;
                for (int rr = 1; rr <= 8; rr++)
                    ien[(element - 1) * 8 + rr - 1] =
                        start + citcoms_offset[rr][0] + citcoms_offset[rr][1] * noz + citcoms_offset[rr][2] * noz * nox;
            }
    return elx * ely * elz;
}

/* construct_node_maps (citcoms/lib/Construct_arrays.c:264-310) with ID[node].doff[d] = 3*(node-1) + d-1
 * (construct_id, :156-160).  node_map has nno*42 entries. */
void oracle_citcoms_node_maps(int nox, int noy, int noz, int *node_map) {
    const int dims = 3, dims2 = 2, max_eqn = 14 * dims;
    const int nno = nox * noy * noz, neq = 3 * nno, noxz = nox * noz;
    for (int i = 0; i < max_eqn * nno; i++) node_map[i] = neq; /* neq indicates an invalid eqn # */
    for (int ii = 1; ii <= noy; ii++)
        for (int jj = 1; jj <= nox; jj++)
            for (int kk = 1; kk <= noz; kk++) {
                const int nn = kk + (jj - 1) * noz + (ii - 1) * noxz;
                for (int doff = 1; doff <= dims; doff++) node_map[(nn - 1) * max_eqn + doff - 1] = 3 * (nn - 1) + doff - 1;
                int ia = 0;
                int is = 1, ie = dims2, js = 1, je = dims, ks = 1, ke = dims;
                if (kk == 1) ks = 2;
                if (kk == noz) ke = 2;
                if (jj == 1) js = 2;
                if (jj == nox) je = 2;
                if (ii == 1) is = 2;
                if (ii == noy) ie = 2;
                for (int i = is; i <= ie; i++)
                    for (int j = js; j <= je; j++)
                        for (int k = ks; k <= ke; k++) {
                            const int ja = nn - ((2 - i) * noxz + (2 - j) * noz + 2 - k);
                            if (ja < nn) {
                                ia++;
                                for (int doff = 1; doff <= dims; doff++)
                                    node_map[(nn - 1) * max_eqn + ia * dims + doff - 1] = 3 * (ja - 1) + doff - 1;
                            }
                        }
            }
}

/* construct_node_ks (citcoms/lib/Construct_arrays.c:330-456) without boundary-condition weights (w = ww = 1): the lower
 * half of every element matrix elt_k[e] (24 x 24 row-major, rows/columns 3*(a-1)+direction) is added into the slots of
 * the higher-numbered node.  Returns 0, or -1 when a slot is missing (the reference asserts). */
int oracle_citcoms_node_ks(int nel, int nno, const int *ien, const double *elt_k, const int *node_map, float *k1,
                           float *k2, float *k3) {
    const int dims = 3, ends = 8, lms = 24, max_eqn = 14 * dims;
    memset(k1, 0, sizeof(float) * (size_t)max_eqn * nno);
    memset(k2, 0, sizeof(float) * (size_t)max_eqn * nno);
    memset(k3, 0, sizeof(float) * (size_t)max_eqn * nno);
    for (int element = 1; element <= nel; element++) {
        const double *elt_K = elt_k + (size_t)(element - 1) * lms * lms;
        for (int i = 1; i <= ends; i++) { /* i, is the node we are storing to */
            const int node = ien[(element - 1) * 8 + i - 1];
            const int pp = (i - 1) * dims;
            const int loc0 = (node - 1) * max_eqn;
            for (int j = 1; j <= ends; j++) { /* j is the node we are receiving from */
                const int node1 = ien[(element - 1) * 8 + j - 1];
                if (node1 <= node) { /* only for half of the matrix, because of the symmetry */
                    const int qq = (j - 1) * dims;
                    for (int d = 0; d < dims; d++) { /* search for direction d+1 */
                        const int eqn = 3 * (node1 - 1) + d;
                        int index = -1;
                        for (int k = 0; k < max_eqn; k++)
                            if (node_map[loc0 + k] == eqn) {
                                index = k;
                                break;
                            }
                        if (index < 0) return -1;
                        k1[loc0 + index] += elt_K[pp * lms + qq + d];
                        k2[loc0 + index] += elt_K[(pp + 1) * lms + qq + d];
                        k3[loc0 + index] += elt_K[(pp + 2) * lms + qq + d];
                    }
                }
            }
        }
    }
    return 0;
}

/* n_assemble_del2_u (citcoms/lib/Element_calculations.c:516-565), one cap, without the parallel exchange and the
 * boundary-condition strip: u and Au have neq + 1 entries (entry neq is the dummy equation). */
void oracle_citcoms_n_assemble_del2_u(int nno, const int *node_map, const float *k1, const float *k2, const float *k3,
                                      double *u, double *Au) {
    const int neq = 3 * nno, max_eqn = 42;
    for (int e = 0; e <= neq; e++) Au[e] = 0.0;
    u[neq] = 0.0;
    for (int e = 1; e <= nno; e++) {
        const int eqn1 = 3 * (e - 1), eqn2 = eqn1 + 1, eqn3 = eqn1 + 2;
        const double U1 = u[eqn1], U2 = u[eqn2], U3 = u[eqn3];
        const int *C = node_map + (e - 1) * max_eqn;
        const float *B1 = k1 + (e - 1) * max_eqn, *B2 = k2 + (e - 1) * max_eqn, *B3 = k3 + (e - 1) * max_eqn;
        for (int i = 3; i < max_eqn; i++) {
            const double UU = u[C[i]];
            Au[eqn1] += B1[i] * UU;
            Au[eqn2] += B2[i] * UU;
            Au[eqn3] += B3[i] * UU;
        }
        for (int i = 0; i < max_eqn; i++) Au[C[i]] += B1[i] * U1 + B2[i] * U2 + B3[i] * U3;
    }
}
