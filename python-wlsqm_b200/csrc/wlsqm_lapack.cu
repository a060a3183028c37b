// wlsqm_lapack.cu -- K4: batched small dense LU (factor / solve) and condition numbers.
//
// Replaces the batched general drivers of wlsqm/utils/lapackdrivers.pyx:
//   mgeneralfactor[p]_c   :1612-1692  (dgetrf per system)
//   mgeneralfactored[p]_c :1638-1723  (dgetrs per system)
//   mgeneral[p]_c         :1551-1609  (dgesv = both)
// with the reference's memory layout: A (n,n,nlhs) Fortran-contiguous, b (n,nlhs) Fortran,
// ipiv (n,nlhs) int32 Fortran, 1-based pivots, everything in place, `info` never reported
// (lapackdrivers.pyx:1575,1606,1633,1663 ignore it too: a singular system yields inf/NaN).
// One warp per system; the matrix lives in the warp's shared-memory slice during factorisation.
//
// cond_kernel: 2-norm condition number of the scaled problem matrices kept by prepare(debug=True)
// (svd_c -> dgesvd, lapackdrivers.pyx:1756-1774, called from impl.pyx:662-682), by one-sided
// (Hestenes) Jacobi SVD, one warp per matrix.
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

// LU of the n x n column-major matrix A (leading dimension lda) held in shared memory; warp-cooperative.
__device__ __forceinline__ void warp_getrf(int n, double* A, int lda, int* ipiv, int lane) {
    for (int p = 0; p < n; ++p) {
        double best = -1.0;
        int bi = p;
        for (int i = p + lane; i < n; i += 32) {
            const double v = fabs(A[i + lda * p]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) ipiv[p] = bi;
        if (bi != p)
            for (int m = lane; m < n; m += 32) {
                const double t = A[p + lda * m];
                A[p + lda * m] = A[bi + lda * m];
                A[bi + lda * m] = t;
            }
        __syncwarp();
        const double rp = 1.0 / A[p + lda * p];
        for (int i = p + 1 + lane; i < n; i += 32) A[i + lda * p] *= rp;
        __syncwarp();
        for (int i = p + 1 + lane; i < n; i += 32) {
            const double l = A[i + lda * p];
            for (int m = p + 1; m < n; ++m) A[i + lda * m] -= l * A[p + lda * m];
        }
        __syncwarp();
    }
}

__global__ void getrf_kernel(int n, long long nlhs, double* __restrict__ Ag, int* __restrict__ ipivg, int warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lda = n | 1;
    double* A = smem + (size_t)warp * warp_doubles;
    int* piv = reinterpret_cast<int*>(A + (size_t)lda * n);
    for (long long l = (long long)blockIdx.x * nwarps + warp; l < nlhs; l += (long long)gridDim.x * nwarps) {
        double* g = Ag + l * (long long)n * n;
        for (int t = lane; t < n * n; t += 32) A[(t % n) + lda * (t / n)] = g[t];
        __syncwarp();
        warp_getrf(n, A, lda, piv, lane);
        for (int t = lane; t < n * n; t += 32) g[t] = A[(t % n) + lda * (t / n)];
        for (int t = lane; t < n; t += 32) ipivg[l * n + t] = piv[t] + 1;   // 1-based, like LAPACK
        __syncwarp();
    }
}

// dgetrs 'N', one right-hand side per system; one thread per system would be uncoalesced on LU, so a
// warp stages LU in shared memory and lanes split the axpy updates.
__global__ void getrs_kernel(int n, long long nlhs, const double* __restrict__ LUg, const int* __restrict__ ipivg,
                             double* __restrict__ bg, int warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lda = n | 1;
    double* A = smem + (size_t)warp * warp_doubles;
    double* b = A + (size_t)lda * n;
    for (long long l = (long long)blockIdx.x * nwarps + warp; l < nlhs; l += (long long)gridDim.x * nwarps) {
        const double* g = LUg + l * (long long)n * n;
        for (int t = lane; t < n * n; t += 32) A[(t % n) + lda * (t / n)] = g[t];
        for (int t = lane; t < n; t += 32) b[t] = bg[l * n + t];
        __syncwarp();
        if (lane == 0)
            for (int p = 0; p < n; ++p) {
                const int ip = ipivg[l * n + p] - 1;
                if (ip != p && ip >= 0 && ip < n) { const double t = b[p]; b[p] = b[ip]; b[ip] = t; }
            }
        __syncwarp();
        for (int p = 0; p < n; ++p) {
            const double xp = b[p];
            __syncwarp();
            for (int i = p + 1 + lane; i < n; i += 32) b[i] -= A[i + lda * p] * xp;
            __syncwarp();
        }
        for (int p = n - 1; p >= 0; --p) {
            const double xp = b[p] / A[p + lda * p];
            __syncwarp();
            if (lane == 0) b[p] = xp;
            for (int i = lane; i < p; i += 32) b[i] -= A[i + lda * p] * xp;
            __syncwarp();
        }
        for (int t = lane; t < n; t += 32) bg[l * n + t] = b[t];
        __syncwarp();
    }
}

static int lapack_cfg(int n, int& warps, size_t& smem, int& warp_doubles, int extra) {
    const int lda = n | 1;
    warp_doubles = (lda * n + extra + 1) & ~1;
    const size_t per_warp = (size_t)warp_doubles * 8;
    if (per_warp > 200 * 1024) return -1;
    warps = (int)((96 * 1024) / per_warp);
    if (warps < 1) warps = 1;
    if (warps > 8) warps = 8;
    smem = per_warp * warps;
    return 0;
}

cudaError_t launch_getrf(int n, long long nlhs, double* A, int* ipiv, cudaStream_t st) {
    if (nlhs == 0 || n == 0) return cudaSuccess;
    int warps, wd;
    size_t smem;
    if (lapack_cfg(n, warps, smem, wd, (n + 1) / 2 + 1)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(getrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (nlhs + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    getrf_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(n, nlhs, A, ipiv, wd);
    return cudaGetLastError();
}

cudaError_t launch_getrs(int n, long long nlhs, const double* LU, const int* ipiv, double* b, cudaStream_t st) {
    if (nlhs == 0 || n == 0) return cudaSuccess;
    int warps, wd;
    size_t smem;
    if (lapack_cfg(n, warps, smem, wd, n + 1)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(getrs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (nlhs + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    getrs_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(n, nlhs, LU, ipiv, b, wd);
    return cudaGetLastError();
}

// ---- condition numbers ---------------------------------------------------------------------------
__global__ void cond_kernel(long long ncases, const CaseMeta* meta, CaseMeta uni, const double* __restrict__ As,
                            int as_stride, double* __restrict__ cond, int warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* A = smem + (size_t)warp * warp_doubles;
    for (long long c = (long long)blockIdx.x * nwarps + warp; c < ncases; c += (long long)gridDim.x * nwarps) {
        const int n = meta ? meta[c].nr : uni.nr;
        if (n < 1) { if (lane == 0) cond[c] = __longlong_as_double(0x7ff8000000000000LL); continue; }
        const int lda = n | 1;
        const double* g = As + c * (long long)as_stride;
        for (int t = lane; t < n * n; t += 32) A[(t % n) + lda * (t / n)] = g[t];
        __syncwarp();
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0.0;
            for (int p = 0; p < n - 1; ++p)
                for (int q = p + 1; q < n; ++q) {
                    double a = 0.0, b = 0.0, g2 = 0.0;
                    for (int i = lane; i < n; i += 32) {
                        const double x = A[i + lda * p], y = A[i + lda * q];
                        a = fma(x, x, a); b = fma(y, y, b); g2 = fma(x, y, g2);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        b += __shfl_xor_sync(0xffffffffu, b, o);
                        g2 += __shfl_xor_sync(0xffffffffu, g2, o);
                    }
                    const double denom = sqrt(a * b);
                    if (denom > 0.0 && fabs(g2) > 1e-15 * denom) {
                        off = fmax(off, fabs(g2) / denom);
                        const double zeta = (b - a) / (2.0 * g2);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                        for (int i = lane; i < n; i += 32) {
                            const double x = A[i + lda * p], y = A[i + lda * q];
                            A[i + lda * p] = cs * x - sn * y;
                            A[i + lda * q] = sn * x + cs * y;
                        }
                    }
                    __syncwarp();
                }
            if (off < 1e-15) break;
        }
        double smax = 0.0, smin = 1.79769313486231570e308;
        for (int m = 0; m < n; ++m) {
            double a = 0.0;
            for (int i = lane; i < n; i += 32) a = fma(A[i + lda * m], A[i + lda * m], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            a = sqrt(a);
            smax = fmax(smax, a);
            smin = fmin(smin, a);
        }
        if (lane == 0) cond[c] = smax / smin;
        __syncwarp();
    }
}

cudaError_t launch_cond(int n_max, long long ncases, const CaseMeta* meta, const CaseMeta& uni, const double* As,
                        int as_stride, double* cond, cudaStream_t st) {
    if (ncases == 0) return cudaSuccess;
    int warps, wd;
    size_t smem;
    if (lapack_cfg(n_max < 1 ? 1 : n_max, warps, smem, wd, 0)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(cond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (ncases + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cond_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(ncases, meta, uni, As, as_stride, cond, wd);
    return cudaGetLastError();
}

}  // namespace wlsqm
