// The G4S graph engine ABI (first "next" row).
//
// The reference ships only the DECLARATION of the engine that its applications call through a function pointer:
//     void spmm_dense(u_int32_t numNodes, u_int32_t degree, const double** edgeWeight, const double* vertexStates,
//                     double* temp, double* result, fun_gather gather, fun_apply apply, double* time, int threadNum);
//     (citcoms/bin/Citcom.c:45-48, assigned at :93; citcoms/lib/global_defs.h:48-49, :854-857)
// and one spelled-out copy of its loop, GraphProcess (deepmd/source/op/graph.h:21-32):
//     for every vertex vi: for every neighbour nb < degree: gather(vi, nb, ...); then apply(vi, ...).
// This file supplies that missing body with the declared signature, so that CitcomS links against the library.
// gather/apply are HOST callbacks and therefore run on the CPU by construction; the one concrete instance that
// matters (CitcomS's element-by-element operator, whose gather is citcoms/lib/Element_calculations.c:453-471) has a
// device implementation that needs no callbacks: g4s_ebe_matvec_device (ebe.cu).
#include <omp.h>

#include <chrono>

#include "g4s_b200.h"

extern "C" void spmm_dense(uint32_t numNodes, uint32_t degree, const double **edgeWeight, const double *vertexStates,
                           double *temp, double *result, g4s_fun_gather gather, g4s_fun_apply apply, double *time,
                           int threadNum) {
    (void)temp;  // the caller passes its result buffer here too (Element_calculations.c:500); the engine never reads it
    const auto t0 = std::chrono::steady_clock::now();
    const int threads = threadNum > 0 ? threadNum : 1;
    // schedule(dynamic, 1) as GraphProcess; a gather that scatter-adds into shared entries of `result` (CitcomS's
    // does) is only safe with threadNum == 1, which is what its caller passes
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (long long vi = 0; vi < (long long)numNodes; ++vi) {
        for (uint32_t nb = 0; nb < degree; ++nb) gather((int)vi, (int)nb, edgeWeight, vertexStates, result);
        if (apply) apply((int)vi, edgeWeight, vertexStates, result);
    }
    if (time) *time += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
