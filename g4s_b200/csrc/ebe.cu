// Element-by-element operator  Au += sum_e  scatter_e( K_e · gather_e(u) )  on the GPU — the device form of the one
// concrete G4S-engine instance in the reference: CitcomS's e_assemble_del2_u, whose per-(element, node) callback is
// `gather` at citcoms/lib/Element_calculations.c:453-471 (8 nodes x 3 dof: a dense 24 x 24 block per element,
// row-major, row = 3(a-1)+(i-1), column = 3(b-1)+(j-1)).  The IEN / ID indirection of the callback is flattened once
// into elem_dofs[e][c] = equation number of local dof c of element e.
//
// One warp per element, lane r owns row r of K_e: it streams its 8*ndof-byte row with 128-bit loads (every sector of
// K_e is fetched from HBM exactly once; the kernel is bound by 8 ndof^2 bytes per element), multiplies by the element's
// u entries, which lanes fetch once and broadcast with shuffles, and scatter-adds with RED.F64.  Elements that share a
// node add into the same entries, so the order of those additions is not fixed (rounding-level differences only).
#include "common.cuh"

namespace g4s {

template <int NDOF>
__global__ void __launch_bounds__(256) ebe_matvec_kernel(int nel, const double *__restrict__ elt_k,
                                                         const int *__restrict__ elem_dofs,
                                                         const double *__restrict__ u, double *__restrict__ Au) {
    const int lane = threadIdx.x & 31;
    for (long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; e < nel;
         e += ((long long)gridDim.x * blockDim.x) >> 5) {
        int dof = -1;
        double ue = 0.0;
        if (lane < NDOF) {
            dof = __ldg(elem_dofs + e * NDOF + lane);
            ue = __ldg(u + dof);
        }
        const double *row = elt_k + (size_t)e * NDOF * NDOF + (size_t)(lane < NDOF ? lane : 0) * NDOF;
        double acc = 0.0;
        if (NDOF % 2 == 0) {
            const double2 *row2 = reinterpret_cast<const double2 *>(row);
#pragma unroll
            for (int c = 0; c < NDOF; c += 2) {
                const double2 k = __ldg(row2 + c / 2);
                acc = fma(k.x, __shfl_sync(0xffffffffu, ue, c), acc);
                acc = fma(k.y, __shfl_sync(0xffffffffu, ue, c + 1), acc);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NDOF; ++c) acc = fma(__ldg(row + c), __shfl_sync(0xffffffffu, ue, c), acc);
        }
        if (lane < NDOF) atomicAdd(Au + dof, acc);
    }
}

}  // namespace g4s

using namespace g4s;

extern "C" int g4s_ebe_matvec_device(int nel, int ndof, const double *elt_k_dev, const int *elem_dofs_dev,
                                     const double *u_dev, double *Au_dev, void *stream) {
    if (nel < 0 || !elt_k_dev || !elem_dofs_dev || !u_dev || !Au_dev)
        return fail(G4S_ERR_INVALID, "g4s_ebe_matvec_device: bad arguments");
    if (ndof != 24 && ndof != 8 && ndof != 4)
        return fail(G4S_ERR_INVALID, "g4s_ebe_matvec_device: ndof must be 24 (3-D), 8 (2-D) or 4 (1-D): loc_mat_size of "
                                     "citcoms/lib/element_definitions.h:108");
    if (nel == 0) return G4S_OK;
    int rc = ensure_device();
    if (rc) return rc;
    const int grid = (int)std::min<long long>(((long long)nel * 32 + 255) / 256, (long long)sm_count() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (ndof == 24) ebe_matvec_kernel<24><<<grid, 256, 0, st>>>(nel, elt_k_dev, elem_dofs_dev, u_dev, Au_dev);
    else if (ndof == 8) ebe_matvec_kernel<8><<<grid, 256, 0, st>>>(nel, elt_k_dev, elem_dofs_dev, u_dev, Au_dev);
    else ebe_matvec_kernel<4><<<grid, 256, 0, st>>>(nel, elt_k_dev, elem_dofs_dev, u_dev, Au_dev);
    G4S_CHECK_LAUNCH("ebe_matvec_kernel");
    return G4S_OK;
}
