// Host-side micro-benchmark (no GPU): what does filling a freshly allocated 670 MB output array cost on this box?
// Mirrors parallel_memcpy / default_alloc of the library's mkl() twin: 1 MB pieces over all OpenMP threads into
// posix_memalign(2 MB) + MADV_HUGEPAGE memory.   g++ -O2 -fopenmp tools/micro/host_copy.cpp -o tools/micro/host_copy
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void pcopy(void *dst, const void *src, size_t bytes) {
    const size_t piece = 1u << 20;
    const long long n = (long long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        const size_t off = (size_t)i * piece;
        memcpy((char *)dst + off, (const char *)src + off, std::min(piece, bytes - off));
    }
}
int main() {
    const size_t bytes = (size_t)670 << 20;
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    char line[256] = "?";
    if (f) {
        if (!fgets(line, sizeof line, f)) line[0] = 0;
        fclose(f);
    }
    printf("threads %d, THP: %s", omp_get_max_threads(), line);
    void *src = nullptr;
    posix_memalign(&src, 2 << 20, bytes);
    memset(src, 1, bytes);
    for (int mode = 0; mode < 4; ++mode) {
        double best = 1e9, alloc_t = 0;
        for (int rep = 0; rep < 3; ++rep) {
            void *dst = nullptr;
            double t0 = now();
            posix_memalign(&dst, 2 << 20, bytes);
            if (mode != 3) madvise(dst, bytes, MADV_HUGEPAGE);
            if (mode == 1) {  // populate in parallel first
                const size_t piece = 16u << 20;
                const long long n = (long long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static)
                for (long long i = 0; i < n; ++i) madvise((char *)dst + i * piece, std::min(piece, bytes - (size_t)i * piece), 23 /*MADV_POPULATE_WRITE*/);
            }
            if (mode == 2) pcopy(dst, src, bytes);  // warm: second copy into the same pages is what a reused buffer costs
            double t1 = now();
            pcopy(dst, src, bytes);
            double t2 = now();
            best = std::min(best, t2 - t1);
            alloc_t = t1 - t0;
            free(dst);
        }
        const char *names[] = {"fresh pages (MADV_HUGEPAGE)", "after parallel MADV_POPULATE_WRITE", "pages already touched", "fresh pages, no madvise"};
        printf("%-40s copy %.1f ms = %.1f GB/s (allocation + preparation before it: %.1f ms)\n", names[mode], best * 1e3, bytes / best / 1e9, alloc_t * 1e3);
    }
    return 0;
}
