/*
 * oracle/spmm_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, OpenMP over rows) of the one third-party kernel the
 * reference's hot path calls:  DGL  update_all(fn.copy_u('h','m'), fn.sum('m','h'))
 * at /root/reference/module/layer.py:35-37, 88-90 (training), 43-44, 96-97 (eval) and
 * /root/reference/train.py:196-198, 203-205 (precompute).  DGL 0.9 (requirements.txt:5,
 * README.md:41 -- not vendored) lowers that call to an unweighted CSR SpMM:
 *     Y[v, :] = sum over in-edges (u -> v) of X[u, :]            (f32, sum in edge order)
 * and autograd runs the same kernel on the reversed graph.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file (oracle/Makefile -> oracle/_build/libspmm_ref.so).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* COO (dst[e], src[e]) -> CSR by destination; stable (edge order kept inside a row), like the
 * COO->CSR conversion DGL performs lazily on a freshly built heterograph (train.py:276).  Parallel stable
 * counting sort: edges are cut into T contiguous blocks, each block histograms its rows, a scan turns the
 * histograms into per-(block,row) write cursors, then every block scatters its own edges in order. */
#include <omp.h>

/* Threads per call (0 = OpenMP default).  Set explicitly by the caller: torchrun exports OMP_NUM_THREADS=1, which
 * would otherwise turn the CPU baseline single-threaded behind our back. */
static int g_threads = 0;
void bns_ref_set_threads(int n) { g_threads = n > 0 ? n : 0; }
static int nthreads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

int bns_ref_coo_to_csr(int64_t n_dst, int64_t nnz, const int64_t *dst, const int64_t *src,
                       int64_t *indptr, int64_t *cols)
{
    int T = nthreads();
    if (T > 32) T = 32;
    if (nnz < (1 << 16)) T = 1;
    const int64_t rows = n_dst > 0 ? n_dst : 1;
    int64_t *cur = (int64_t *)calloc((size_t)T * (size_t)rows, sizeof(int64_t));
    if (!cur) return -2;
    int bad = 0;
    const int64_t blk = (nnz + T - 1) / T;
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int b = 0; b < T; ++b) {
        int64_t *h = cur + (size_t)b * rows;
        const int64_t e0 = b * blk, e1 = (e0 + blk < nnz) ? e0 + blk : nnz;
        for (int64_t e = e0; e < e1; ++e) {
            if (dst[e] < 0 || dst[e] >= n_dst) { bad = 1; continue; }
            h[dst[e]]++;
        }
    }
    if (bad) { free(cur); return -1; }
    indptr[0] = 0;
    for (int64_t v = 0; v < n_dst; ++v) {          /* row totals -> indptr; histograms -> write cursors */
        int64_t run = indptr[v];
        for (int b = 0; b < T; ++b) {
            const int64_t c = cur[(size_t)b * rows + v];
            cur[(size_t)b * rows + v] = run;
            run += c;
        }
        indptr[v + 1] = run;
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int b = 0; b < T; ++b) {
        int64_t *h = cur + (size_t)b * rows;
        const int64_t e0 = b * blk, e1 = (e0 + blk < nnz) ? e0 + blk : nnz;
        for (int64_t e = e0; e < e1; ++e) cols[h[dst[e]]++] = src[e];
    }
    free(cur);
    return 0;
}

/* Y[v,:] = sum_{k in [indptr[v], indptr[v+1])} X[cols[k], :] */
int bns_ref_spmm_sum_f32(int64_t n_dst, const int64_t *indptr, const int64_t *cols,
                         const float *X, int64_t ldx, int64_t F, float *Y, int64_t ldy)
{
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads())
    for (int64_t v = 0; v < n_dst; ++v) {
        float *y = Y + v * ldy;
        for (int64_t f = 0; f < F; ++f) y[f] = 0.0f;
        for (int64_t k = indptr[v]; k < indptr[v + 1]; ++k) {
            const float *x = X + cols[k] * ldx;
            for (int64_t f = 0; f < F; ++f) y[f] += x[f];
        }
    }
    return 0;
}

/* dX[u,:] = sum over out-edges (u -> v) of dY[v,:]: the same sum on the reversed graph.  Rows of the
 * forward CSR are scattered; to stay race-free without atomics the caller passes the CSR of the
 * reversed graph (built with bns_ref_coo_to_csr on swapped endpoints) to bns_ref_spmm_sum_f32. */
