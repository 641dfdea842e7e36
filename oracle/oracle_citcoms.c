/* TEST INFRASTRUCTURE ONLY — CPU restatement of CitcomS's assembled ("node format") stiffness operator, the
 * reference-side data format of SURVEY.md §8f row 2.  PARITY UNPINNED: the citcoms tree needs MPI and its own build
 * system and is not buildable here, and it ships no test for these routines; every function restates the
 * arithmetic and the ORDER of the cited lines (same slot numbering, same accumulation order, float coefficients) and the
 * set is cross-checked against an independently assembled sparse matrix in tests/test_citcoms.py.
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
 * (construct_id, :156-160).  node_map has nno*42 entries.  The reference enumerates, for node (y, x, z), the offsets
 * dy in {-1, 0}, dx and dz in {-1, 0, +1} (clipped at the mesh faces) in that nesting order and keeps those that land on
 * a LOWER node number; slot ia (1-based, in order of discovery) receives that neighbour's three equations. */
void oracle_citcoms_node_maps(int nox, int noy, int noz, int *node_map) {
    const int slots = 14, nno = nox * noy * noz, neq = 3 * nno;
    for (int q = 0; q < 3 * slots * nno; q++) node_map[q] = neq; /* the dummy equation marks an unused slot */
    for (int y = 0; y < noy; y++)
        for (int x = 0; x < nox; x++)
            for (int z = 0; z < noz; z++) {
                const int node0 = z + x * noz + y * nox * noz; /* 0-based node number: z fastest, then x, then y */
                int *slot = node_map + 3 * slots * node0;
                for (int d = 0; d < 3; d++) slot[d] = 3 * node0 + d;
                int ia = 0;
                for (int dy = -1; dy <= 0; dy++) {
                    if (y + dy < 0) continue;
                    for (int dx = -1; dx <= 1; dx++) {
                        if (x + dx < 0 || x + dx >= nox) continue;
                        for (int dz = -1; dz <= 1; dz++) {
                            if (z + dz < 0 || z + dz >= noz) continue;
                            const int other = node0 + dz + dx * noz + dy * nox * noz;
                            if (other >= node0) continue;
                            ++ia;
                            for (int d = 0; d < 3; d++) slot[3 * ia + d] = 3 * other + d;
                        }
                    }
                }
            }
}

/* construct_node_ks (citcoms/lib/Construct_arrays.c:330-456) without boundary-condition weights (w = ww = 1): for every
 * element and every ordered pair (storing node, contributing node) with contributing <= storing, the 3 x 3 sub-block of
 * the 24 x 24 element matrix (row-major, row / column 3*(a-1)+direction) is added into the storing node's slot that holds
 * the contributing node's equations: Eqn_k{r}[slot + d] += elt_k[3a + r][3b + d].  Accumulation is in float, element by
 * element, as the reference's `higher_precision +=`.  Returns 0, or -1 when a slot is missing (the reference asserts). */
int oracle_citcoms_node_ks(int nel, int nno, const int *ien, const double *elt_k, const int *node_map, float *k1,
                           float *k2, float *k3) {
    const int slots = 14, lms = 24;
    float *k[3] = {k1, k2, k3};
    for (int r = 0; r < 3; r++) memset(k[r], 0, sizeof(float) * 3 * slots * (size_t)nno);
    for (int e = 0; e < nel; e++) {
        const double *K = elt_k + (size_t)e * lms * lms;
        const int *nodes = ien + 8 * e; /* 1-based */
        for (int a = 0; a < 8; a++) {
            const int store = nodes[a] - 1;
            const int *map = node_map + 3 * slots * store;
            for (int b = 0; b < 8; b++) {
                const int from = nodes[b] - 1;
                if (from > store) continue; /* only half of the matrix is stored */
                int where = -1; /* the reference searches the 42 slots for each of the three equations */
                for (int q = 0; q < 3 * slots; q += 3)
                    if (map[q] == 3 * from) {
                        where = q;
                        break;
                    }
                if (where < 0 || map[where + 1] != 3 * from + 1 || map[where + 2] != 3 * from + 2) return -1;
                for (int d = 0; d < 3; d++)
                    for (int r = 0; r < 3; r++)
                        k[r][3 * slots * store + where + d] += K[(3 * a + r) * lms + 3 * b + d];
            }
        }
    }
    return 0;
}

/* n_assemble_del2_u (citcoms/lib/Element_calculations.c:516-565), one cap, without the parallel exchange and the
 * boundary-condition strip: u and Au have neq + 1 entries (entry neq is the dummy equation that unused slots point at).
 * Every stored coefficient is applied twice, in the reference's order: first the node's own three rows gather from its
 * lower neighbours (slots 1..13, ascending), then all 14 slots receive the node's three unknowns (the transposed half
 * and the diagonal block). */
void oracle_citcoms_n_assemble_del2_u(int nno, const int *node_map, const float *k1, const float *k2, const float *k3,
                                      double *u, double *Au) {
    const int neq = 3 * nno, width = 42;
    memset(Au, 0, sizeof(double) * ((size_t)neq + 1));
    u[neq] = 0.0;
    for (int node0 = 0; node0 < nno; node0++) {
        const int *eq = node_map + (size_t)width * node0;
        const float *c1 = k1 + (size_t)width * node0, *c2 = k2 + (size_t)width * node0, *c3 = k3 + (size_t)width * node0;
        const int row = 3 * node0;
        const double own1 = u[row], own2 = u[row + 1], own3 = u[row + 2];
        for (int q = 3; q < width; q++) {
            const double other = u[eq[q]];
            Au[row] += c1[q] * other;
            Au[row + 1] += c2[q] * other;
            Au[row + 2] += c3[q] * other;
        }
        for (int q = 0; q < width; q++) Au[eq[q]] += c1[q] * own1 + c2[q] * own2 + c3[q] * own3;
    }
}
