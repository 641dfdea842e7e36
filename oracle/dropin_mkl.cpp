// TEST INFRASTRUCTURE ONLY.  The reference's own mm/ headers (unmodified, compiled where they lie) with
// include/g4s_b200.hpp in place of mkl_mult.h: the reference's call shape mkl(A, B, C, timing) runs on the GPU
// library and the result is compared, with the reference's own CSR::operator==, against the reference's own
// HashSpGEMM<false,true> on the CPU.  Built by oracle/Makefile into oracle/_ref/dropin_mkl.
#include "all.h"
#include "Timings.h"
#include "g4s_b200.hpp"
#include <cstdio>
int main() {
    const int n = 48;                       // 2-D 5-point Laplacian built with the reference's own containers
    const int rows = n * n;
    std::vector<int> rp(rows + 1, 0), ci; std::vector<double> va;
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
        const int r = j * n + i;
        if (j > 0) { ci.push_back(r - n); va.push_back(-1); }
        if (i > 0) { ci.push_back(r - 1); va.push_back(-1); }
        ci.push_back(r); va.push_back(4);
        if (i < n - 1) { ci.push_back(r + 1); va.push_back(-1); }
        if (j < n - 1) { ci.push_back(r + n); va.push_back(-1); }
        rp[r + 1] = (int)ci.size();
    }
    CSR<int, double> A(rp.data(), ci.data(), va.data(), rows, rows, (int)ci.size(), 0), B(A), C, Cref;
    Timings timing;
    mkl(A, B, C, timing);                   // GPU, reference signature (mm/inc/mkl_mult.h:113-117)
    HashSpGEMM<false, true>(A, B, Cref, std::multiplies<double>(), std::plus<double>());  // reference CPU
    const bool same = (C == Cref);          // CSR::operator== (mm/inc/CSR.h:343-408)
    std::printf("nnzC %d ref %d equal %d total %.6f\n", C.nnz, Cref.nnz, (int)same, timing.total);
    return (same && C.nnz == 13 * n * n - 20 * n + 4) ? 0 : 1;
}
