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
#include <algorithm>
#include <cstdlib>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

cudaError_t launch_getrf(int n, long long nlhs, double* A, int* ipiv, cudaStream_t st);
cudaError_t launch_getrs(int n, long long nlhs, const double* LU, const int* ipiv, double* b, cudaStream_t st);

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

// ---- n <= 32: rows in registers, several systems per warp ---------------------------------------------------
// LPF lanes per system (LPF = 4, 8, 16, 32 >= n), FPW = 32 / LPF systems side by side in a warp, lane = matrix row held
// in LPF registers (every loop over columns is unrolled).  Partial pivoting without moving rows: the pivot lane
// publishes its row to shared memory at the pivot's position -- which is the LAPACK-layout LU that is written back --
// and the other lanes eliminate in registers (the scheme of prepare_reg_kernel's LU phase).  dgesv solves its right-hand
// side in the same pass: lane = row keeps its multipliers and its U row, the substitutions broadcast one value per
// step by shuffle (no second read of the factors).
//   mode 0: dgetrf (A -> LU, ipiv)     mode 1: dgesv (A -> LU, ipiv; b -> x)     mode 2: dgetrs (LU, ipiv given; b -> x)
// max / min over the LPF consecutive lanes of a lane's group: one warp-wide REDUX per group (independent, so their
// latencies overlap) instead of a dependent shuffle butterfly
template <int LPF>
__device__ __forceinline__ unsigned lu_group_max(unsigned v, int grp) {
    if constexpr (LPF == 32) return __reduce_max_sync(0xffffffffu, v);
    if constexpr (LPF <= 8) {      // many small groups: a short butterfly inside the group is cheaper
#pragma unroll
        for (int o = 1; o < LPF; o <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    unsigned r = 0u;
#pragma unroll
    for (int g = 0; g < 32 / LPF; ++g) {
        const unsigned m = __reduce_max_sync(0xffffffffu, grp == g ? v : 0u);
        if (grp == g) r = m;
    }
    return r;
}
template <int LPF>
__device__ __forceinline__ unsigned lu_group_min(unsigned v, int grp) {
    if constexpr (LPF == 32) return __reduce_min_sync(0xffffffffu, v);
    if constexpr (LPF <= 8) {
#pragma unroll
        for (int o = 1; o < LPF; o <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    unsigned r = 0u;
#pragma unroll
    for (int g = 0; g < 32 / LPF; ++g) {
        const unsigned m = __reduce_min_sync(0xffffffffu, grp == g ? v : 0xffffffffu);
        if (grp == g) r = m;
    }
    return r;
}

template <int LPF>
__global__ void __launch_bounds__(256, LPF == 32 ? 2 : 4) lu_reg_kernel(int n, long long nlhs, double* __restrict__ Ag, int* __restrict__ ipivg,
                                                     double* __restrict__ bg, int mode) {
    constexpr int FPW = 32 / LPF;
    constexpr int LDG = LPF + 2;                      // row stride of the published LU (even: 16 B aligned rows)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane / LPF, jl = lane % LPF;
    double* G = smem + ((size_t)warp * FPW + grp) * (LPF * LDG + LPF);   // [LPF][LDG] LU rows in pivot order
    double* bs = G + LPF * LDG;                                            // [LPF] right-hand side (mode 2)
    const long long nn = (long long)n * n;
    const long long nsets = (nlhs + FPW - 1) / FPW;
    for (long long set = (long long)blockIdx.x * nwarps + warp; set < nsets; set += (long long)gridDim.x * nwarps) {
        const long long sys = set * FPW + grp;
        const bool live = sys < nlhs;
        const bool row = live && jl < n;
        double* ag = Ag + (live ? sys : 0) * nn;
        double a[LPF];
#pragma unroll
        for (int m = 0; m < LPF; ++m) a[m] = (row && m < n) ? ag[jl + (long long)n * m] : 0.0;
        int mystep = jl;          // position of this lane's row in the pivot order
        if (mode != 2) {
            bool act = row;
            int pos = jl;         // current position of this lane's row in LAPACK's (swapped) row order
            int myipiv = 0;       // ipiv[jl]
            mystep = -1;
#pragma unroll
            for (int p = 0; p < LPF; ++p) {
                if (p < n) {      // warp-uniform
                    // pivot = first row of largest |a[.][p]| in current row order (idamax).  The magnitude of a
                    // non-negative double orders like its bit pattern: REDUX on the high words, the low words and the
                    // position only among the lanes that tie (a NaN entry counts as 0 so that every lane agrees)
                    double mag = fabs(a[p]);
                    if (!(mag == mag)) mag = 0.0;
                    const unsigned khi = (unsigned)__double2hiint(mag), klo = (unsigned)__double2loint(mag);
                    const unsigned hi = lu_group_max<LPF>(act ? khi : 0u, grp);
                    const bool c1 = act && khi == hi;
                    const unsigned lo = lu_group_max<LPF>(c1 ? klo : 0u, grp);
                    const bool c2 = c1 && klo == lo;
                    const int bpos = (int)lu_group_min<LPF>(c2 ? (unsigned)pos : 0xffffu, grp);
                    const bool win = c2 && pos == bpos;
                    if (jl == p) myipiv = bpos + 1;
                    const bool piv = live && win;
                    if (piv) {
#pragma unroll
                        for (int m = 0; m < LPF; m += 2) *reinterpret_cast<double2*>(G + p * LDG + m) = make_double2(a[m], a[m + 1]);
                        mystep = p;
                        act = false;
                    } else if (act && pos == p) {
                        pos = bpos;             // the row that sat at position p takes the pivot row's old place
                    }
                    __syncwarp();
                    if (act) {
                        const double rp = 1.0 / G[p * LDG + p];
                        const double l = a[p] * rp;
                        a[p] = l;
#pragma unroll
                        for (int m = ((p + 1) & ~1); m < LPF; m += 2) {
                            const double2 u2 = *reinterpret_cast<const double2*>(G + p * LDG + m);
                            if (m > p) a[m] = fma(-l, u2.x, a[m]);
                            a[m + 1] = fma(-l, u2.y, a[m + 1]);
                        }
                    }
                }
            }
            __syncwarp();
            // LU back to the caller, LAPACK layout: row p of the factors is G[p]
            if (row) {
#pragma unroll
                for (int m = 0; m < LPF; ++m)
                    if (m < n) ag[jl + (long long)n * m] = G[jl * LDG + m];
                if (ipivg) ipivg[sys * n + jl] = myipiv;
            }
        }
        if (mode != 0) {
            // lane = row jl of the pivot order from here on: the factors are re-read from G (mode 1), the right-hand
            // side is permuted through shared memory; each substitution step broadcasts one value from a fixed lane
            double y = 0.0;
            if (mode == 2) {
                if (row) bs[jl] = bg[sys * n + jl];
                __syncwarp();
                if (live && jl == 0)
                    for (int p = 0; p < n; ++p) {
                        const int ip = ipivg[sys * n + p] - 1;
                        if (ip != p && ip >= 0 && ip < n) { const double t = bs[p]; bs[p] = bs[ip]; bs[ip] = t; }
                    }
            } else {
                if (row && mystep >= 0) bs[mystep] = bg[sys * n + jl];   // the row's own entry travels with the row
#pragma unroll
                for (int m = 0; m < LPF; m += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(G + jl * LDG + m);
                    a[m] = v.x;
                    a[m + 1] = v.y;
                }
            }
            __syncwarp();
            if (row) y = bs[jl];
            // L y = P b (unit lower)
#pragma unroll
            for (int p = 0; p < LPF; ++p) {
                if (p < n) {
                    const double yp = __shfl_sync(FULL, y, grp * LPF + p);
                    if (jl > p) y = fma(-a[p], yp, y);
                }
            }
            // U x = y
#pragma unroll
            for (int m = LPF - 1; m >= 0; --m) {
                if (m < n) {
                    if (jl == m) y = y / a[m];
                    const double xm = __shfl_sync(FULL, y, grp * LPF + m);
                    if (jl < m) y = fma(-a[m], xm, y);
                }
            }
            if (row) bg[sys * n + jl] = y;
        }
        __syncwarp();
    }
}

template <int LPF>
static cudaError_t launch_lu_reg_t(int n, long long nlhs, double* A, int* ipiv, double* b, int mode, cudaStream_t st) {
    constexpr int FPW = 32 / LPF;
    const int warps = 8;
    const size_t smem = (size_t)warps * FPW * (LPF * (LPF + 2) + LPF) * 8;
    cudaError_t e = cudaFuncSetAttribute(lu_reg_kernel<LPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long nsets = (nlhs + FPW - 1) / FPW;
    long long blocks = (nsets + warps - 1) / warps;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lu_reg_kernel<LPF><<<(unsigned)blocks, warps * 32, smem, st>>>(n, nlhs, A, ipiv, b, mode);
    return cudaGetLastError();
}

// n <= 32 (and not disabled by WLSQM_LU_SMEM=1, the A/B switch of the tests)
static bool lu_reg_ok(int n) {
    static const bool off = [] { const char* v = getenv("WLSQM_LU_SMEM"); return v && v[0] == '1'; }();
    return n <= 32 && !off;
}
static cudaError_t launch_lu_reg(int n, long long nlhs, double* A, int* ipiv, double* b, int mode, cudaStream_t st) {
    if (n <= 4) return launch_lu_reg_t<4>(n, nlhs, A, ipiv, b, mode, st);
    if (n <= 8) return launch_lu_reg_t<8>(n, nlhs, A, ipiv, b, mode, st);
    if (n <= 16) return launch_lu_reg_t<16>(n, nlhs, A, ipiv, b, mode, st);
    return launch_lu_reg_t<32>(n, nlhs, A, ipiv, b, mode, st);
}

static int lapack_cfg(int n, int& warps, size_t& smem, int& warp_doubles, int extra) {
    const int lda = n | 1;
    warp_doubles = (lda * n + extra + 1) & ~1;
    const size_t per_warp = (size_t)warp_doubles * 8;
    if (per_warp > 200 * 1024) return -1;
    // CTA size that keeps the most warps resident (these kernels are latency-bound: one system per warp, a chain of n
    // pivot steps), whole multiples of the four schedulers preferred: n = 36 -> five CTAs of 4 warps instead of two of 8
    int best_w = 1, best_total = 0;
    for (int w = 1; w <= 8; ++w) {
        if ((size_t)w * per_warp > 200 * 1024) break;
        const int c = std::min(64 / w, (int)((227 * 1024) / ((size_t)w * per_warp + 1024)));
        const int total = (w * c) & ~3;
        if (total > best_total || (total == best_total && total > 0)) { best_total = total; best_w = w; }
    }
    warps = best_w;
    smem = per_warp * warps;
    return 0;
}

cudaError_t launch_gesv(int n, long long nlhs, double* A, int* ipiv, double* b, cudaStream_t st) {
    if (nlhs == 0 || n == 0) return cudaSuccess;
    if (lu_reg_ok(n)) return launch_lu_reg(n, nlhs, A, ipiv, b, 1, st);
    cudaError_t e = launch_getrf(n, nlhs, A, ipiv, st);
    if (e == cudaSuccess) e = launch_getrs(n, nlhs, A, ipiv, b, st);
    return e;
}

cudaError_t launch_getrf(int n, long long nlhs, double* A, int* ipiv, cudaStream_t st) {
    if (nlhs == 0 || n == 0) return cudaSuccess;
    if (lu_reg_ok(n)) return launch_lu_reg(n, nlhs, A, ipiv, nullptr, 0, st);
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
    if (lu_reg_ok(n)) return launch_lu_reg(n, nlhs, const_cast<double*>(LU), const_cast<int*>(ipiv), b, 2, st);
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

// ---- symmetric indefinite systems: Bunch-Kaufman U D U^T (dsytrf / dsytrs / dsysv, uplo = 'U') ---------------
// Replaces the batched symmetric drivers of wlsqm/utils/lapackdrivers.pyx:
//   msymmetricfactor[p]_c   :1199-1233, 1275-1314   (dsytrf 'U' per system)
//   msymmetricfactored[p]_c :1236-1272, 1317-1354   (dsytrs 'U' per system, one right-hand side each)
//   msymmetric[p]_c         :1107-1196              (dsysv = both)
//   msymmetrize[p]_c        :204-278                (A <- (A + A^T)/2)
// Same layout as the general drivers; ipiv is LAPACK's: ipiv[k] > 0 = 1x1 block, rows/columns k and ipiv[k]
// interchanged; ipiv[k] = ipiv[k-1] < 0 = 2x2 block in (k-1, k), rows/columns k-1 and -ipiv[k] interchanged.
// Only the upper triangle is read and written.  The factorisation is LAPACK's unblocked dsytf2 (the algorithm
// dsytrf runs for matrices up to its block size), diagonal pivoting with alpha = (1 + sqrt(17)) / 8.

// index of the first entry of largest magnitude among v(i), i = lo .. hi-1 (idamax); v by callable; warp-wide result
// LANES threads work on one system: 32 = a warp (the matrix in shared memory, lanes split the vector operations), 1 = one
// thread per system (small n: no cross-lane traffic at all, and the threads of a warp may take different pivoting paths)
template <int LANES>
__device__ __forceinline__ void group_sync() {
    if constexpr (LANES > 1) __syncwarp();
}

template <int LANES, typename V>
__device__ __forceinline__ int warp_idamax(int lo, int hi, int lane, V&& v, double& vmax) {
    double best = -1.0;
    int bi = lo;
    for (int i = lo + lane; i < hi; i += LANES) {
        const double a = fabs(v(i));
        if (a > best) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    vmax = best < 0.0 ? 0.0 : best;
    return bi;
}

// A: n x n column-major in shared memory (leading dimension lda), upper triangle; ipiv 0-based row indices with
// LAPACK's sign convention applied on store (see sytrf_kernel); w: 2n doubles of scratch
template <int LANES>
__device__ __forceinline__ void warp_sytf2_upper(int n, double* A, int lda, int* ipiv, double* w, int lane) {
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    int k = n - 1;                      // 0-based index of the current column
    while (k >= 0) {
        int kstep = 1, kp = k;
        const double absakk = fabs(A[k + lda * k]);
        double colmax = 0.0;
        int imax = 0;
        if (k > 0) imax = warp_idamax<LANES>(0, k, lane, [&](int i) { return A[i + lda * k]; }, colmax);
        if (fmax(absakk, colmax) == 0.0 || absakk != absakk) {
            kp = k;                     // singular (or NaN) column: no interchange, LAPACK sets info and goes on
        } else {
            if (absakk >= alpha * colmax) {
                kp = k;
            } else {
                // largest off-diagonal entry in row imax of the leading (k+1) x (k+1) block
                double rowmax = 0.0, r2 = 0.0;
                warp_idamax<LANES>(imax + 1, k + 1, lane, [&](int j) { return A[imax + lda * j]; }, rowmax);
                if (imax > 0) {
                    warp_idamax<LANES>(0, imax, lane, [&](int i) { return A[i + lda * imax]; }, r2);
                    rowmax = fmax(rowmax, r2);
                }
                if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
                else if (fabs(A[imax + lda * imax]) >= alpha * rowmax) kp = imax;
                else { kp = imax; kstep = 2; }
            }
            const int kk = k - kstep + 1;
            if (kp != kk) {
                // interchange rows and columns kk and kp in the leading block
                for (int i = lane; i < kp; i += LANES) {
                    const double t = A[i + lda * kk]; A[i + lda * kk] = A[i + lda * kp]; A[i + lda * kp] = t;
                }
                for (int j = kp + 1 + lane; j < kk; j += LANES) {
                    const double t = A[j + lda * kk]; A[j + lda * kk] = A[kp + lda * j]; A[kp + lda * j] = t;
                }
                if (lane == 0) {
                    double t = A[kk + lda * kk]; A[kk + lda * kk] = A[kp + lda * kp]; A[kp + lda * kp] = t;
                    if (kstep == 2) { t = A[k - 1 + lda * k]; A[k - 1 + lda * k] = A[kp + lda * k]; A[kp + lda * k] = t; }
                }
                group_sync<LANES>();
            }
            if (kstep == 1) {
                // A := A - U(k) D(k) U(k)^T = A - x x^T / d  (dsyr), then x := x / d
                const double r1 = 1.0 / A[k + lda * k];
                for (int j = 0; j < k; ++j) {
                    const double xj = A[j + lda * k];
                    if (xj != 0.0) {
                        const double t = -r1 * xj;
                        for (int i = lane; i <= j; i += LANES) A[i + lda * j] = fma(A[i + lda * k], t, A[i + lda * j]);
                    }
                }
                group_sync<LANES>();
                for (int i = lane; i < k; i += LANES) A[i + lda * k] *= r1;
                group_sync<LANES>();
            } else if (k > 1) {
                double d12 = A[k - 1 + lda * k];
                const double d22 = A[k - 1 + lda * (k - 1)] / d12, d11 = A[k + lda * k] / d12;
                const double t = 1.0 / (d11 * d22 - 1.0);
                d12 = t / d12;
                for (int j = lane; j < k - 1; j += LANES) {
                    w[j] = d12 * (d11 * A[j + lda * (k - 1)] - A[j + lda * k]);        // wkm1
                    w[n + j] = d12 * (d22 * A[j + lda * k] - A[j + lda * (k - 1)]);    // wk
                }
                group_sync<LANES>();
                for (int j = k - 2; j >= 0; --j) {
                    const double wkm1 = w[j], wk = w[n + j];
                    for (int i = lane; i <= j; i += LANES)
                        A[i + lda * j] = A[i + lda * j] - A[i + lda * k] * wk - A[i + lda * (k - 1)] * wkm1;
                }
                group_sync<LANES>();
                for (int j = lane; j < k - 1; j += LANES) {
                    A[j + lda * k] = w[n + j];
                    A[j + lda * (k - 1)] = w[j];
                }
                group_sync<LANES>();
            }
        }
        if (lane == 0) {
            if (kstep == 1) ipiv[k] = kp + 1;
            else { ipiv[k] = -(kp + 1); ipiv[k - 1] = -(kp + 1); }
        }
        k -= kstep;
    }
    group_sync<LANES>();
}

// dsytrs 'U', one right-hand side: b := A^-1 b with A = U D U^T from warp_sytf2_upper (ipiv 1-based, LAPACK signs)
template <int LANES>
__device__ __forceinline__ void warp_sytrs_upper(int n, const double* A, int lda, const int* ipiv, double* b, int lane) {
    // U D x = b
    int k = n - 1;
    while (k >= 0) {
        if (ipiv[k] > 0) {
            const int kp = ipiv[k] - 1;
            if (lane == 0 && kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            group_sync<LANES>();
            const double bk = b[k];
            for (int i = lane; i < k; i += LANES) b[i] -= A[i + lda * k] * bk;
            group_sync<LANES>();
            if (lane == 0) b[k] = bk * (1.0 / A[k + lda * k]);
            group_sync<LANES>();
            k -= 1;
        } else {
            const int kp = -ipiv[k] - 1;
            if (lane == 0 && kp != k - 1) { const double t = b[k - 1]; b[k - 1] = b[kp]; b[kp] = t; }
            group_sync<LANES>();
            const double bk = b[k], bkm = b[k - 1];
            for (int i = lane; i < k - 1; i += LANES) b[i] = (b[i] - A[i + lda * k] * bk) - A[i + lda * (k - 1)] * bkm;
            group_sync<LANES>();
            if (lane == 0) {
                const double akm1k = A[k - 1 + lda * k];
                const double akm1 = A[k - 1 + lda * (k - 1)] / akm1k, ak = A[k + lda * k] / akm1k;
                const double denom = akm1 * ak - 1.0;
                const double bkm1 = bkm / akm1k, bkk = bk / akm1k;
                b[k - 1] = (ak * bkm1 - bkk) / denom;
                b[k] = (akm1 * bkk - bkm1) / denom;
            }
            group_sync<LANES>();
            k -= 2;
        }
    }
    // U^T x = b
    k = 0;
    while (k < n) {
        const int nb = ipiv[k] > 0 ? 1 : 2;
        double d0 = 0.0, d1 = 0.0;
        for (int i = lane; i < k; i += LANES) {
            d0 = fma(A[i + lda * k], b[i], d0);
            if (nb == 2) d1 = fma(A[i + lda * (k + 1)], b[i], d1);
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o);
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        }
        group_sync<LANES>();
        if (lane == 0) {
            b[k] -= d0;
            if (nb == 2) b[k + 1] -= d1;
            const int kp = (ipiv[k] > 0 ? ipiv[k] : -ipiv[k]) - 1;
            if (kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
        }
        group_sync<LANES>();
        k += nb;
    }
}

// factor (and, with bg != nullptr, solve) nlhs symmetric systems; do_factor = 0: A and ipiv hold factors already
__global__ void sy_kernel(int n, long long nlhs, double* __restrict__ Ag, int* __restrict__ ipivg, double* __restrict__ bg,
                          int do_factor, int warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lda = n | 1;
    double* A = smem + (size_t)warp * warp_doubles;
    double* w = A + (size_t)lda * n;          // 2n scratch, the first n double as the right-hand side
    int* piv = reinterpret_cast<int*>(w + 2 * n);
    for (long long l = (long long)blockIdx.x * nwarps + warp; l < nlhs; l += (long long)gridDim.x * nwarps) {
        double* g = Ag + l * (long long)n * n;
        for (int t = lane; t < n * n; t += 32) A[(t % n) + lda * (t / n)] = g[t];
        if (!do_factor)
            for (int t = lane; t < n; t += 32) piv[t] = ipivg[l * n + t];
        __syncwarp();
        if (do_factor) {
            warp_sytf2_upper<32>(n, A, lda, piv, w, lane);
            // only the upper triangle (incl. diagonal) is written back: the strict lower triangle is not referenced
            for (int t = lane; t < n * n; t += 32) {
                const int i = t % n, j = t / n;
                if (i <= j) g[t] = A[i + lda * j];
            }
            if (ipivg)
                for (int t = lane; t < n; t += 32) ipivg[l * n + t] = piv[t];
        }
        if (bg) {
            for (int t = lane; t < n; t += 32) w[t] = bg[l * n + t];
            __syncwarp();
            warp_sytrs_upper<32>(n, A, lda, piv, w, lane);
            for (int t = lane; t < n; t += 32) bg[l * n + t] = w[t];
        }
        __syncwarp();
    }
}

// n <= 12: ONE THREAD per system.  A warp-per-system pass leaves 29 of 32 lanes idle on a 3 x 3 matrix and pays a warp
// synchronisation per vector operation; here the 32 systems of a warp are staged into shared memory with coalesced loads
// (they are contiguous in global memory), every thread runs the same Bunch-Kaufman code on its own region -- threads may
// take different pivoting paths, there is no cross-lane traffic -- and the results go back coalesced.
__global__ void __launch_bounds__(128) sy_thread_kernel(int n, long long nlhs, double* __restrict__ Ag, int* __restrict__ ipivg,
                                                        double* __restrict__ bg, int do_factor, int thr_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nn = n * n;
    double* wbase = smem + (size_t)warp * 32 * thr_doubles;
    double* A = wbase + (size_t)lane * thr_doubles;       // this thread's system: [n][n] column-major, lda = n
    double* w = A + nn;                                    // 2n scratch, the first n double as the right-hand side
    int* piv = reinterpret_cast<int*>(w + 2 * n);
    const long long nsets = (nlhs + 31) / 32;
    for (long long set = (long long)blockIdx.x * nwarps + warp; set < nsets; set += (long long)gridDim.x * nwarps) {
        const long long l0 = set * 32;
        const int cnt = (int)min(32LL, nlhs - l0);
        const bool live = lane < cnt;
        const double* g0 = Ag + l0 * nn;
        for (int f = lane; f < cnt * nn; f += 32) wbase[(size_t)(f / nn) * thr_doubles + (f % nn)] = g0[f];
        if (!do_factor)
            for (int f = lane; f < cnt * n; f += 32)
                reinterpret_cast<int*>(wbase + (size_t)(f / n) * thr_doubles + nn + 2 * n)[f % n] = ipivg[l0 * n + f];
        __syncwarp();
        if (live && do_factor) warp_sytf2_upper<1>(n, A, n, piv, w, 0);
        if (bg) {
            // (the right-hand side shares its place with the factorisation's scratch: it comes in afterwards)
            __syncwarp();
            for (int f = lane; f < cnt * n; f += 32) wbase[(size_t)(f / n) * thr_doubles + nn + (f % n)] = bg[l0 * n + f];
            __syncwarp();
            if (live) warp_sytrs_upper<1>(n, A, n, piv, w, 0);
        }
        __syncwarp();
        if (do_factor) {
            double* gw = Ag + l0 * nn;
            for (int f = lane; f < cnt * nn; f += 32) {
                const int e = f % nn;
                if ((e % n) <= (e / n)) gw[f] = wbase[(size_t)(f / nn) * thr_doubles + e];     // upper triangle only
            }
            if (ipivg)
                for (int f = lane; f < cnt * n; f += 32)
                    ipivg[l0 * n + f] = reinterpret_cast<const int*>(wbase + (size_t)(f / n) * thr_doubles + nn + 2 * n)[f % n];
        }
        if (bg)
            for (int f = lane; f < cnt * n; f += 32) bg[l0 * n + f] = wbase[(size_t)(f / n) * thr_doubles + nn + (f % n)];
        __syncwarp();
    }
}

__global__ void symmetrize_kernel(int n, long long nlhs, double* __restrict__ Ag) {
    const long long per = (long long)n * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nlhs * per; t += (long long)gridDim.x * blockDim.x) {
        const long long l = t / per;
        const int r = (int)(t - l * per), i = r % n, j = r / n;
        if (i < j) {   // strict upper triangle: this thread owns the pair (i,j), (j,i)
            double* g = Ag + l * per;
            const double v = 0.5 * (g[i + (long long)n * j] + g[j + (long long)n * i]);
            g[i + (long long)n * j] = v;
            g[j + (long long)n * i] = v;
        }
    }
}

cudaError_t launch_sy(int n, long long nlhs, double* A, int* ipiv, double* b, int do_factor, cudaStream_t st) {
    if (nlhs == 0 || n == 0) return cudaSuccess;
    static const bool warp_only = [] { const char* v = getenv("WLSQM_SY_WARP"); return v && v[0] == '1'; }();   // A/B switch
    static const int thread_max_n = [] { const char* v = getenv("WLSQM_SY_THREAD_MAX_N"); return (v && *v) ? atoi(v) : 12; }();
    // (measured, ns per system, thread / warp per system: n = 3 0.20 / 2.26, n = 8 2.1 / ~8, n = 10 5.9 / 9.6, n = 12 8.3 / 10.8, n = 15 16.8 / 14.7)
    if (n <= thread_max_n && n <= 16 && !warp_only) {
        const int thr = (n * n + 2 * n + (n + 1) / 2) | 1;          // odd stride: the threads' regions spread over the banks
        const int threads = std::max(32, std::min(128, (int)((200 * 1024) / ((size_t)thr * 8)) & ~31));
        const size_t smem = (size_t)threads * thr * 8;
        cudaError_t e = cudaFuncSetAttribute(sy_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        long long blocks = (nlhs + threads - 1) / threads;
        if (blocks > 148 * 16) blocks = 148 * 16;
        sy_thread_kernel<<<(unsigned)blocks, threads, smem, st>>>(n, nlhs, A, ipiv, b, do_factor, thr);
        return cudaGetLastError();
    }
    int warps, wd;
    size_t smem;
    if (lapack_cfg(n, warps, smem, wd, 2 * n + (n + 1) / 2 + 1)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(sy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (nlhs + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sy_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(n, nlhs, A, ipiv, b, do_factor, wd);
    return cudaGetLastError();
}

cudaError_t launch_symmetrize(int n, long long nlhs, double* A, cudaStream_t st) {
    if (nlhs == 0 || n < 2) return cudaSuccess;
    long long blocks = ((long long)n * n * nlhs + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    symmetrize_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, nlhs, A);
    return cudaGetLastError();
}

// ---- tridiag (lapackdrivers.pyx:854-877 -> LAPACK DGTSV, one right-hand side) --------------------------------------
// The reference's minimal example driver: Gaussian elimination with partial pivoting on a tridiagonal system; dl / d / du
// are overwritten by the factorisation like DGTSV does, b by the solution.  One thread: the recurrence is sequential.
__global__ void gtsv_kernel(int n, double* __restrict__ dl, double* __restrict__ d, double* __restrict__ du,
                            double* __restrict__ b) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < n - 1; ++i) {
        if (fabs(d[i]) >= fabs(dl[i])) {
            if (d[i] == 0.0) return;                      // singular: DGTSV stops with info = i + 1 (the reference ignores info)
            const double fact = dl[i] / d[i];
            d[i + 1] -= fact * du[i];
            b[i + 1] -= fact * b[i];
            if (i < n - 2) dl[i] = 0.0;                   // (DGTSV clears DL only in its main loop, not in the last step)
        } else {                                          // interchange rows i and i + 1
            const double fact = d[i] / dl[i];
            d[i] = dl[i];
            double temp = d[i + 1];
            d[i + 1] = du[i] - fact * temp;
            if (i < n - 2) {
                dl[i] = du[i + 1];
                du[i + 1] = -fact * dl[i];
            }
            du[i] = temp;
            temp = b[i];
            b[i] = b[i + 1];
            b[i + 1] = temp - fact * b[i + 1];
        }
    }
    if (d[n - 1] == 0.0) return;
    b[n - 1] /= d[n - 1];
    if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
    for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - dl[i] * b[i + 2]) / d[i];
}

cudaError_t launch_gtsv(int n, double* dl, double* d, double* du, double* b, cudaStream_t st) {
    if (n < 1) return cudaSuccess;
    gtsv_kernel<<<1, 32, 0, st>>>(n, dl, d, du, b);
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
