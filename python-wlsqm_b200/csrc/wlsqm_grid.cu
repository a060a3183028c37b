// wlsqm_grid.cu -- uniform-grid spatial search on the device: k nearest neighbours of every point of a cloud
// (neighbourhood construction) and of arbitrary query points (nearest local model).
//
// This is the step *before* the fitting path in the reference's pipeline, done there on the host with
// scipy.spatial.cKDTree (examples/expertsolver_example.py:51-66: hoods = tree.query(x, k+1)[1][:,1:];
// wlsqm/fitter/expert.pyx:676-681,837: tree.query(x, k=1)).  Results are the exact k nearest neighbours,
// ordered by (distance, index); they coincide with cKDTree's wherever distances are distinct (ties are
// ordered arbitrarily by cKDTree -- tests/test_gpu_grid.py compares index for index on random clouds).
//
// Build: bounding box -> cell size for ~2 points per cell -> cell id per point -> stable radix sort of
// (cell, index) pairs (cub) -> per-cell start offsets (histogram + exclusive scan) -> coordinates gathered
// into sorted order.  Query: one thread per query walks the cell shells around its own cell, keeps the k best
// candidates in a sorted list, and stops when the k-th distance is below the distance to the unvisited region.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <new>
#include <string>

#include <cub/cub.cuh>

#include "../../include/wlsqm_b200.h"
#include "wlsqm_grid.h"
#include "wlsqm_kernels.h"
#include "wlsqm_mem.h"

using namespace wlsqm;

struct wlsqm_grid {
    int dim = 0, device = 0;
    long long n = 0, ncells = 0;
    GridView v{};
    int* cell_start = nullptr;
    int* sorted_idx = nullptr;
    double* sorted_x = nullptr;
    cudaStream_t stream = nullptr;
    long long bytes = 0;
};

namespace wlsqm {
GridView grid_view(const wlsqm_grid* g) { return g->v; }
int grid_device(const wlsqm_grid* g) { return g->device; }
}  // namespace wlsqm

namespace {

int gfail(int code, const char* msg) {
    wlsqm::set_last_error(msg);
    return code;
}

#define GCU(expr)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            std::string m__ = std::string(#expr) + ": " + cudaGetErrorString(e__);                  \
            return gfail(e__ == cudaErrorMemoryAllocation ? WLSQM_E_MEMORY : WLSQM_E_CUDA, m__.c_str()); \
        }                                                                                           \
    } while (0)

// order-preserving map double -> uint64 (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long enc(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
inline double dec(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
    double x;
    memcpy(&x, &b, 8);
    return x;
}

__global__ void bbox_kernel(const double* __restrict__ x, long long s0, long long n, int dim, unsigned long long* mm) {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        for (int d = 0; d < dim; ++d) {
            const double v = x[i * s0 + d];
            if (fabs(v) <= 1.79769313486231570e308) {     // NaN and +/-inf coordinates do not shape the box
                lo[d] = fmin(lo[d], v);
                hi[d] = fmax(hi[d], v);
            }
        }
    for (int d = 0; d < dim; ++d) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mm[d], enc(lo[d]));
            atomicMax(&mm[3 + d], enc(hi[d]));
        }
    }
}

__global__ void cell_id_kernel(GridView g, const double* __restrict__ x, long long s0, int* __restrict__ cid,
                               int* __restrict__ idx, int* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int c = 0;
    for (int d = g.dim - 1; d >= 0; --d) c = c * g.dims[d] + grid_cell_coord(g, d, x[i * s0 + d]);
    cid[i] = c;
    idx[i] = (int)i;
    atomicAdd(&counts[c], 1);
}

__global__ void gather_sorted_kernel(long long n, int dim, const double* __restrict__ x, long long s0,
                                     const int* __restrict__ sorted_idx, double* __restrict__ xs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * dim) return;
    const long long p = t / dim;
    const int d = (int)(t - p * dim);
    xs[t] = x[(long long)sorted_idx[p] * s0 + d];
}

// k nearest grid points of every query.  SELF: the queries are the grid's own points, visited in sorted (cell)
// order so that the threads of a warp walk the same cells; the point itself is skipped when exclude_self.
template <int DIM, int KMAX, bool SELF>
__global__ void __launch_bounds__(128) knn_kernel(GridView g, const double* __restrict__ xq, long long xq_s0, long long nq,
                                                  int k, int exclude_self, int32_t* __restrict__ out32,
                                                  int64_t* __restrict__ out64, double* __restrict__ out_d2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    double q[DIM];
    long long qi;      // row of the output
    if (SELF) {
        qi = g.sorted_idx[t];
#pragma unroll
        for (int d = 0; d < DIM; ++d) q[d] = g.sorted_x[t * DIM + d];
    } else {
        qi = t;
#pragma unroll
        for (int d = 0; d < DIM; ++d) q[d] = xq[t * xq_s0 + d];
    }
    int cq[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cq[d] = grid_cell_coord(g, d, q[d]);

    double bd[KMAX];
    int bi[KMAX];
    int cnt = 0;
    double worst = INFINITY;
    int worst_i = 0x7fffffff;
    const int self = (SELF && exclude_self) ? (int)qi : -1;
    int maxr = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) maxr = max(maxr, max(cq[d], g.dims[d] - 1 - cq[d]));

    for (int r = 0; r <= maxr; ++r) {
        // cells whose Chebyshev distance from the query's cell is exactly r
        const int z0 = DIM >= 3 ? max(cq[DIM >= 3 ? 2 : 0] - r, 0) : 0;
        const int z1 = DIM >= 3 ? min(cq[DIM >= 3 ? 2 : 0] + r, g.dims[2] - 1) : 0;
        const int y0 = DIM >= 2 ? max(cq[DIM >= 2 ? 1 : 0] - r, 0) : 0;
        const int y1 = DIM >= 2 ? min(cq[DIM >= 2 ? 1 : 0] + r, g.dims[1] - 1) : 0;
        const int x0 = max(cq[0] - r, 0), x1 = min(cq[0] + r, g.dims[0] - 1);
        for (int cz = z0; cz <= z1; ++cz) {
            const bool fz = DIM >= 3 && abs(cz - cq[DIM >= 3 ? 2 : 0]) == r;
            for (int cy = y0; cy <= y1; ++cy) {
                const bool fy = fz || (DIM >= 2 && abs(cy - cq[DIM >= 2 ? 1 : 0]) == r);
                // on a face of the shell the whole x range belongs to it, otherwise only its two end cells
                const int step = fy ? 1 : max(x1 - x0, 1);
                for (int cx = x0; cx <= x1; cx += step) {
                    if (!fy && abs(cx - cq[0]) != r) continue;
                    const long long cell = ((long long)cz * g.dims[1] + cy) * g.dims[0] + cx;
                    const int p1 = g.cell_start[cell + 1];
                    for (int p = g.cell_start[cell]; p < p1; ++p) {
                        double d2 = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) {
                            const double dd = g.sorted_x[(long long)p * DIM + d] - q[d];
                            d2 += dd * dd;
                        }
                        const int id = g.sorted_idx[p];
                        if (id == self || !(d2 < INFINITY)) continue;     // (points with NaN / inf coordinates are nobody's neighbour)
                        if (cnt == k && !(d2 < worst || (d2 == worst && id < worst_i))) continue;
                        int j = cnt < k ? cnt : k - 1;
                        while (j > 0 && (bd[j - 1] > d2 || (bd[j - 1] == d2 && bi[j - 1] > id))) {
                            bd[j] = bd[j - 1];
                            bi[j] = bi[j - 1];
                            --j;
                        }
                        bd[j] = d2;
                        bi[j] = id;
                        if (cnt < k) ++cnt;
                        if (cnt == k) { worst = bd[k - 1]; worst_i = bi[k - 1]; }
                    }
                }
            }
        }
        if (cnt == k) {
            // every unvisited point lies outside the block of visited cells: at least `margin` away
            double margin = INFINITY;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                if (cq[d] - r > 0) margin = fmin(margin, q[d] - (g.lo[d] + (double)(cq[d] - r) * g.h));
                if (cq[d] + r < g.dims[d] - 1) margin = fmin(margin, (g.lo[d] + (double)(cq[d] + r + 1) * g.h) - q[d]);
            }
            // (the cell edges carry rounding of order ulp(extent): keep one more shell unless clearly inside)
            if (margin > 0.0 && worst < margin * margin * (1.0 - 1e-12)) break;
        }
    }
    for (int j = 0; j < k; ++j) {
        const int id = j < cnt ? bi[j] : (int)g.n;          // missing neighbours are reported as n (like cKDTree)
        if (out32) out32[qi * k + j] = id;
        if (out64) out64[qi * k + j] = id;
        if (out_d2) out_d2[qi * k + j] = j < cnt ? bd[j] : INFINITY;
    }
}

template <int DIM, bool SELF>
cudaError_t launch_knn_k(const GridView& g, const double* xq, long long s0, long long nq, int k, int excl, int32_t* o32,
                         int64_t* o64, double* od2, cudaStream_t st) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((nq + threads - 1) / threads);
    if (k <= 8) knn_kernel<DIM, 8, SELF><<<blocks, threads, 0, st>>>(g, xq, s0, nq, k, excl, o32, o64, od2);
    else if (k <= 32) knn_kernel<DIM, 32, SELF><<<blocks, threads, 0, st>>>(g, xq, s0, nq, k, excl, o32, o64, od2);
    else if (k <= 64) knn_kernel<DIM, 64, SELF><<<blocks, threads, 0, st>>>(g, xq, s0, nq, k, excl, o32, o64, od2);
    else knn_kernel<DIM, 128, SELF><<<blocks, threads, 0, st>>>(g, xq, s0, nq, k, excl, o32, o64, od2);
    return cudaGetLastError();
}

template <int DIM>
cudaError_t launch_knn_d(const GridView& g, const double* xq, long long s0, long long nq, int k, int excl, int32_t* o32,
                         int64_t* o64, double* od2, cudaStream_t st) {
    if (xq) return launch_knn_k<DIM, false>(g, xq, s0, nq, k, excl, o32, o64, od2, st);
    return launch_knn_k<DIM, true>(g, nullptr, 0, nq, k, excl, o32, o64, od2, st);
}

int is_dev(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" {

int wlsqm_grid_destroy(wlsqm_grid_t* g) {
    if (!g) return WLSQM_OK;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    dev_free(g->cell_start); dev_free(g->sorted_idx); dev_free(g->sorted_x);
    if (g->stream) cudaStreamDestroy(g->stream);
    cudaGetLastError();
    delete g;
    return WLSQM_OK;
}

int wlsqm_grid_create(int dimension, int64_t n, const double* x, int64_t x_s0, int device, wlsqm_grid_t** out) {
    if (!out) return gfail(WLSQM_E_VALUE, "out is NULL");
    *out = nullptr;
    if (dimension < 1 || dimension > 3) return gfail(WLSQM_E_VALUE, "dimension must be 1, 2 or 3");
    if (n < 1 || n > 2000000000LL) return gfail(WLSQM_E_VALUE, "the grid needs between 1 and 2e9 points");
    if (!x) return gfail(WLSQM_E_VALUE, "x is NULL");
    if (wlsqm_device_count() < 1) return gfail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    GCU(cudaSetDevice(device));
    wlsqm_grid* g = new (std::nothrow) wlsqm_grid();
    if (!g) return gfail(WLSQM_E_MEMORY, "out of host memory");
    g->dim = dimension; g->device = device; g->n = n;
    cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete g; return gfail(WLSQM_E_CUDA, cudaGetErrorString(e)); }
    cudaStream_t st = g->stream;
    auto bail = [&](int rc) { wlsqm_grid_destroy(g); return rc; };
    // the points may still be being produced on the caller's stream (device arrays)
    if (order_after_caller(st) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(caller_stream()); }

    // the points on the device (a host array is staged; only the sorted copy is kept)
    double* xd = nullptr;
    const double* xdev = x;
    long long s0 = x_s0;
    if (!is_dev(x)) {
        if (dev_alloc((void**)&xd, (size_t)n * dimension * 8) != cudaSuccess) return bail(gfail(WLSQM_E_MEMORY, "cudaMalloc failed (grid points)"));
        e = cudaMemcpy2DAsync(xd, (size_t)dimension * 8, x, (size_t)x_s0 * 8, (size_t)dimension * 8, (size_t)n, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { dev_free(xd); return bail(gfail(WLSQM_E_CUDA, cudaGetErrorString(e))); }
        xdev = xd;
        s0 = dimension;
    }
    unsigned long long* mm = nullptr;
    int *cid = nullptr, *idx = nullptr, *cid_s = nullptr, *counts = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() { dev_free(xd); dev_free(mm); dev_free(cid); dev_free(idx); dev_free(cid_s); dev_free(counts); dev_free(tmp); };
    auto fail_here = [&](int code, const char* msg) { cudaStreamSynchronize(st); cleanup(); return bail(gfail(code, msg)); };

    // ---- bounding box -> cell size ----------------------------------------------------------------------
    if (dev_alloc((void**)&mm, 6 * 8) != cudaSuccess) return fail_here(WLSQM_E_MEMORY, "cudaMalloc failed");
    unsigned long long init[6] = {~0ULL, ~0ULL, ~0ULL, 0ULL, 0ULL, 0ULL};
    cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, st);
    bbox_kernel<<<296, 256, 0, st>>>(xdev, s0, n, dimension, mm);
    unsigned long long hm[6];
    e = cudaMemcpyAsync(hm, mm, sizeof hm, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail_here(WLSQM_E_CUDA, cudaGetErrorString(e));
    double ext[3] = {0, 0, 0};
    GridView& v = g->v;
    v.dim = dimension; v.n = n;
    int deff = 0;
    double vol = 1.0;
    for (int d = 0; d < 3; ++d) {
        v.lo[d] = 0.0; v.dims[d] = 1;
        if (d < dimension) {
            const double lo = dec(hm[d]), hi = dec(hm[3 + d]);
            if (!(lo <= hi)) return fail_here(WLSQM_E_VALUE, "the point set has no finite coordinates");
            v.lo[d] = lo;
            ext[d] = hi - lo;
            if (ext[d] > 0.0) { vol *= ext[d]; ++deff; }
        }
    }
    const double ppc = 2.0;                                       // target points per cell
    double h = 1.0;
    if (deff > 0) {
        h = std::pow(vol / std::max(1.0, (double)n / ppc), 1.0 / deff);
        // (a product of huge finite extents overflows: without this check the loop below never ends)
        if (!std::isfinite(vol) || !std::isfinite(h) || !(h > 0.0))
            return fail_here(WLSQM_E_VALUE, "the extent of the point set is too large for a search grid (overflow)");
        for (int guard = 0;; ++guard) {
            if (guard > 4096) return fail_here(WLSQM_E_VALUE, "no cell size found for the search grid");
            double total = 1.0;
            bool ok = true;
            for (int d = 0; d < dimension; ++d) {
                const double c = std::floor(ext[d] / h) + 1.0;
                if (c > 1048576.0) ok = false;
                total *= c;
            }
            if (ok && total <= std::max(64.0, 4.0 * (double)n) && total < 1.0e9) break;
            h *= 1.25;
        }
    }
    v.h = h; v.inv_h = 1.0 / h;
    long long ncells = 1;
    for (int d = 0; d < dimension; ++d) {
        v.dims[d] = (int)std::floor(ext[d] / h) + 1;
        ncells *= v.dims[d];
    }
    g->ncells = ncells;

    // ---- cell ids, stable sort by cell, start offsets ------------------------------------------------------
    bool okm = dev_alloc((void**)&cid, (size_t)n * 4) == cudaSuccess && dev_alloc((void**)&idx, (size_t)n * 4) == cudaSuccess &&
               dev_alloc((void**)&cid_s, (size_t)n * 4) == cudaSuccess && dev_alloc((void**)&counts, (size_t)(ncells + 1) * 4) == cudaSuccess &&
               dev_alloc((void**)&g->cell_start, (size_t)(ncells + 1) * 4) == cudaSuccess &&
               dev_alloc((void**)&g->sorted_idx, (size_t)n * 4) == cudaSuccess &&
               dev_alloc((void**)&g->sorted_x, (size_t)n * dimension * 8) == cudaSuccess;
    if (!okm) return fail_here(WLSQM_E_MEMORY, "cudaMalloc failed (grid)");
    g->bytes = (ncells + 1) * 4 + n * 4 + n * dimension * 8;
    cudaMemsetAsync(counts, 0, (size_t)(ncells + 1) * 4, st);
    const unsigned nb = (unsigned)((n + 255) / 256);
    cell_id_kernel<<<nb, 256, 0, st>>>(v, xdev, s0, cid, idx, counts);
    size_t tb1 = 0, tb2 = 0;
    int bits = 1;
    while ((1LL << bits) < ncells) ++bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tb1, cid, cid_s, idx, g->sorted_idx, (int)n, 0, bits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, counts, g->cell_start, (int)(ncells + 1), st);
    if (dev_alloc((void**)&tmp, std::max(tb1, tb2) + 16) != cudaSuccess) return fail_here(WLSQM_E_MEMORY, "cudaMalloc failed (sort)");
    size_t tb = std::max(tb1, tb2) + 16;
    cub::DeviceRadixSort::SortPairs(tmp, tb, cid, cid_s, idx, g->sorted_idx, (int)n, 0, bits, st);
    tb = std::max(tb1, tb2) + 16;
    cub::DeviceScan::ExclusiveSum(tmp, tb, counts, g->cell_start, (int)(ncells + 1), st);
    gather_sorted_kernel<<<(unsigned)((n * dimension + 255) / 256), 256, 0, st>>>(n, dimension, xdev, s0, g->sorted_idx, g->sorted_x);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail_here(WLSQM_E_CUDA, cudaGetErrorString(e));
    cleanup();
    v.cell_start = g->cell_start; v.sorted_idx = g->sorted_idx; v.sorted_x = g->sorted_x;
    *out = g;
    return WLSQM_OK;
}

int wlsqm_grid_knn(wlsqm_grid_t* g, const double* xq, int64_t xq_s0, int64_t nq, int k, int exclude_self, int32_t* idx32,
                   int64_t* idx64, double* d2) {
    if (!g) return gfail(WLSQM_E_VALUE, "NULL grid");
    if (k < 1 || k > 128) return gfail(WLSQM_E_VALUE, "k must be between 1 and 128");
    if (!xq) nq = g->n;
    if (nq == 0) return WLSQM_OK;
    if (!idx32 && !idx64 && !d2) return gfail(WLSQM_E_VALUE, "no output array given");
    GCU(cudaSetDevice(g->device));
    cudaStream_t st = g->stream;
    // queries / freshly allocated output tensors of the caller: order after what is queued on the caller's stream
    if (order_after_caller(st) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(caller_stream()); }
    const size_t cnt = (size_t)nq * k;
    // host arrays are staged
    double* xqd = nullptr;
    int32_t* o32 = idx32; int64_t* o64 = idx64; double* od2 = d2;
    void *b32 = nullptr, *b64 = nullptr, *bd2 = nullptr;
    auto done = [&](int rc) { dev_free(xqd); dev_free(b32); dev_free(b64); dev_free(bd2); return rc; };
    const double* xqv = xq;
    long long s0 = xq_s0;
    if (xq && !is_dev(xq)) {
        if (dev_alloc((void**)&xqd, (size_t)nq * g->dim * 8) != cudaSuccess) return done(gfail(WLSQM_E_MEMORY, "cudaMalloc failed (queries)"));
        cudaError_t ec = cudaMemcpy2DAsync(xqd, (size_t)g->dim * 8, xq, (size_t)xq_s0 * 8, (size_t)g->dim * 8, (size_t)nq, cudaMemcpyHostToDevice, st);
        if (ec != cudaSuccess) return done(gfail(WLSQM_E_CUDA, cudaGetErrorString(ec)));
        xqv = xqd; s0 = g->dim;
    }
    if (idx32 && !is_dev(idx32)) { if (dev_alloc((void**)&b32, cnt * 4) != cudaSuccess) return done(gfail(WLSQM_E_MEMORY, "cudaMalloc failed")); o32 = (int32_t*)b32; }
    if (idx64 && !is_dev(idx64)) { if (dev_alloc((void**)&b64, cnt * 8) != cudaSuccess) return done(gfail(WLSQM_E_MEMORY, "cudaMalloc failed")); o64 = (int64_t*)b64; }
    if (d2 && !is_dev(d2)) { if (dev_alloc((void**)&bd2, cnt * 8) != cudaSuccess) return done(gfail(WLSQM_E_MEMORY, "cudaMalloc failed")); od2 = (double*)bd2; }
    cudaError_t e;
    if (g->dim == 1) e = launch_knn_d<1>(g->v, xqv, s0, nq, k, exclude_self, o32, o64, od2, st);
    else if (g->dim == 2) e = launch_knn_d<2>(g->v, xqv, s0, nq, k, exclude_self, o32, o64, od2, st);
    else e = launch_knn_d<3>(g->v, xqv, s0, nq, k, exclude_self, o32, o64, od2, st);
    if (e != cudaSuccess) return done(gfail(WLSQM_E_CUDA, cudaGetErrorString(e)));
    if (b32) cudaMemcpyAsync(idx32, b32, cnt * 4, cudaMemcpyDeviceToHost, st);
    if (b64) cudaMemcpyAsync(idx64, b64, cnt * 8, cudaMemcpyDeviceToHost, st);
    if (bd2) cudaMemcpyAsync(d2, bd2, cnt * 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(gfail(WLSQM_E_CUDA, cudaGetErrorString(e)));
    return done(WLSQM_OK);
}

int wlsqm_grid_info(wlsqm_grid_t* g, int64_t* ncells, double* cell_size, int64_t* bytes) {
    if (!g) return gfail(WLSQM_E_VALUE, "NULL grid");
    if (ncells) *ncells = g->ncells;
    if (cell_size) *cell_size = g->v.h;
    if (bytes) *bytes = g->bytes;
    return WLSQM_OK;
}

}  // extern "C"
