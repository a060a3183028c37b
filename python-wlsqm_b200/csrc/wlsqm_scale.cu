// wlsqm_scale.cu -- batched matrix equilibration: the scaling algorithms of wlsqm.utils.lapackdrivers as one launch
// over nlhs independent matrices ("the other scalers as batched GPU ops").
//
// Replaces, per matrix of the batch, do_rescale (wlsqm/utils/lapackdrivers.pyx:319-385) and the routines it
// dispatches to:
//   rescale_columns_c    :412-424   column scaling by the Euclidean norm
//   rescale_rows_c       :441-453   row scaling by the Euclidean norm
//   rescale_twopass_c    :474-495   columns, then rows of the column-scaled matrix
//   rescale_dgeequ_c     :517-523   LAPACK DGEEQU (inf-norm rows, then columns of the row-scaled matrix; a zero row
//                                   or column fails: the reference raises LinAlgError, here the matrix's status is 0)
//   rescale_ruiz2001_c   :553-623   simultaneous inf-norm scaling, iterated to 1e-15 (<= 100 sweeps)
//   rescale_scalgm_c     :626-847   SCALGM of Chiang & Chandler (scale-up / scale-down sweeps)
//   apply_scaling_c      :293-299   A <- diag(row) A diag(col)
//
// One warp per matrix, the matrix in shared memory.  A lane owns whole rows (row passes) or whole columns (column
// passes) and walks them in the reference's loop order with separately rounded multiplications, additions, divisions
// and square roots (no FMA contraction), so that every scale factor is the reference's bit for bit.
#include <cfloat>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

namespace {

constexpr double SCALE_EPS = 1e-15;     // lapackdrivers.pyx:87

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqr(double a) { return __dsqrt_rn(a); }

struct Mat {
    const double* A;   // [ncols][lda] in shared memory (Fortran order, padded leading dimension)
    int nrows, ncols, lda;
    __device__ __forceinline__ double at(int j, int m) const { return A[j + lda * m]; }
};

// Euclidean column pass (lapackdrivers.pyx:412-424): cs[m] /= sqrt(sum_j (A[j,m] * (cs[m] * rs[j]))^2)
__device__ void cols_eucl(const Mat& M, const double* rs, double* cs, int lane) {
    for (int m = lane; m < M.ncols; m += 32) {
        const double c = cs[m];
        double acc = 0.0;
        for (int j = 0; j < M.nrows; ++j) {
            const double tmp = mul(M.at(j, m), mul(c, rs[j]));
            acc = add(acc, mul(tmp, tmp));
        }
        cs[m] = dvd(cs[m], sqr(acc));
    }
}
// Euclidean row pass (:441-453)
__device__ void rows_eucl(const Mat& M, double* rs, const double* cs, int lane) {
    for (int j = lane; j < M.nrows; j += 32) {
        const double r = rs[j];
        double acc = 0.0;
        for (int m = 0; m < M.ncols; ++m) {
            const double tmp = mul(M.at(j, m), mul(r, cs[m]));
            acc = add(acc, mul(tmp, tmp));
        }
        rs[j] = dvd(rs[j], sqr(acc));
    }
}

// SCALGM building blocks (:664-760).  UP: reciprocal of the smallest non-zero magnitude; DOWN: of the largest.
// `mod` (may be null) is a multiplicative modifier of the other side's scaling that has not been applied yet.
template <bool UP>
__device__ void scalgm_rows(const Mat& M, const double* rs, const double* cs, const double* mod_cs, double* new_rs, int lane) {
    for (int j = lane; j < M.nrows; j += 32) {
        const double r = rs[j];
        double acc = 0.0;
        for (int m = 0; m < M.ncols; ++m) {
            const double s = mod_cs ? mul(mul(r, cs[m]), mod_cs[m]) : mul(r, cs[m]);
            const double tmp = fabs(mul(M.at(j, m), s));
            if (UP ? (acc == 0.0 || (tmp > 0.0 && tmp < acc)) : (tmp > acc)) acc = tmp;
        }
        new_rs[j] = dvd(1.0, acc);
    }
}
template <bool UP>
__device__ void scalgm_cols(const Mat& M, const double* rs, const double* mod_rs, const double* cs, double* new_cs, int lane) {
    for (int m = lane; m < M.ncols; m += 32) {
        const double c = cs[m];
        double acc = 0.0;
        for (int j = 0; j < M.nrows; ++j) {
            const double s = mod_rs ? mul(mul(c, rs[j]), mod_rs[j]) : mul(c, rs[j]);
            const double tmp = fabs(mul(M.at(j, m), s));
            if (UP ? (acc == 0.0 || (tmp > 0.0 && tmp < acc)) : (tmp > acc)) acc = tmp;
        }
        new_cs[m] = dvd(1.0, acc);
    }
}

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

}  // namespace

// algo: the reference's ScalingAlgo values (lapackdrivers.pyx:305-317)
__global__ void __launch_bounds__(256) rescale_kernel(int nrows, int ncols, long long nlhs, double* __restrict__ Ag, int algo,
                                                      double* __restrict__ rsg, double* __restrict__ csg, int* __restrict__ okg,
                                                      int warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lda = nrows | 1;
    double* A = smem + (size_t)warp * warp_doubles;
    double* rs = A + (size_t)lda * ncols;       // row_scale
    double* cs = rs + nrows;                     // col_scale
    double* R1 = cs + ncols;                     // DR / DR1 / DRprev ...
    double* C1 = R1 + nrows;
    double* R2 = C1 + ncols;
    double* C2 = R2 + nrows;
    Mat M{A, nrows, ncols, lda};
    const long long per = (long long)nrows * ncols;
    for (long long l = (long long)blockIdx.x * nwarps + warp; l < nlhs; l += (long long)gridDim.x * nwarps) {
        double* g = Ag + l * per;
        for (int t = lane; t < nrows * ncols; t += 32) A[(t % nrows) + lda * (t / nrows)] = g[t];
        // init_scaling_c (:285-290)
        for (int j = lane; j < nrows; j += 32) rs[j] = R1[j] = R2[j] = 1.0;
        for (int m = lane; m < ncols; m += 32) cs[m] = C1[m] = C2[m] = 1.0;
        __syncwarp();
        int ok = 1;
        if (algo == 1) {
            cols_eucl(M, rs, cs, lane);
        } else if (algo == 2) {
            rows_eucl(M, rs, cs, lane);
        } else if (algo == 3) {
            cols_eucl(M, rs, cs, lane);
            __syncwarp();
            rows_eucl(M, rs, cs, lane);
        } else if (algo == 6) {
            // DGEEQU (LAPACK): r_i = 1 / max_j |a_ij|, then c_j = 1 / max_i |a_ij| r_i, both clipped to [smlnum, bignum]
            const double smlnum = DBL_MIN, bignum = dvd(1.0, smlnum);
            double lo = DBL_MAX, hi = 0.0;
            for (int j = lane; j < nrows; j += 32) {
                double r = 0.0;
                for (int m = 0; m < ncols; ++m) r = fmax(r, fabs(M.at(j, m)));
                rs[j] = r;
                lo = fmin(lo, r);
                hi = fmax(hi, r);
            }
            lo = warp_min_d(lo);
            if (lo == 0.0) ok = 0;       // a zero row: info > 0
            __syncwarp();
            if (ok) {
                for (int j = lane; j < nrows; j += 32) rs[j] = dvd(1.0, fmin(fmax(rs[j], smlnum), bignum));
                __syncwarp();
                lo = DBL_MAX;
                for (int m = lane; m < ncols; m += 32) {
                    double c = 0.0;
                    for (int j = 0; j < nrows; ++j) c = fmax(c, mul(fabs(M.at(j, m)), rs[j]));
                    cs[m] = c;
                    lo = fmin(lo, c);
                }
                lo = warp_min_d(lo);
                if (lo == 0.0) ok = 0;   // a zero column
                __syncwarp();
                if (ok)
                    for (int m = lane; m < ncols; m += 32) cs[m] = dvd(1.0, fmin(fmax(cs[m], smlnum), bignum));
            }
        } else if (algo == 4) {
            // Ruiz (2001), :553-623.  R1 / C1 = DRprev / DCprev, R2 / C2 = DR / DC of the sweep.
            for (int k = 0; k < 100; ++k) {
                for (int j = lane; j < nrows; j += 32) {
                    const double r = R1[j];
                    double acc = 0.0;
                    for (int m = 0; m < ncols; ++m) {
                        const double tmp = fabs(dvd(M.at(j, m), mul(r, C1[m])));
                        if (tmp > acc) acc = tmp;
                    }
                    R2[j] = sqr(acc);
                }
                for (int m = lane; m < ncols; m += 32) {
                    const double c = C1[m];
                    double acc = 0.0;
                    for (int j = 0; j < nrows; ++j) {
                        const double tmp = fabs(dvd(M.at(j, m), mul(c, R1[j])));
                        if (tmp > acc) acc = tmp;
                    }
                    C2[m] = sqr(acc);
                }
                __syncwarp();
                double er = 0.0, ec = 0.0;
                for (int j = lane; j < nrows; j += 32) {
                    R1[j] = mul(R1[j], R2[j]);
                    rs[j] = dvd(rs[j], R2[j]);
                    er = fmax(er, fabs(add(1.0, -mul(R2[j], R2[j]))));
                }
                for (int m = lane; m < ncols; m += 32) {
                    C1[m] = mul(C1[m], C2[m]);
                    cs[m] = dvd(cs[m], C2[m]);
                    ec = fmax(ec, fabs(add(1.0, -mul(C2[m], C2[m]))));
                }
                er = warp_max_d(er);
                ec = warp_max_d(ec);
                __syncwarp();
                if (er < SCALE_EPS && ec < SCALE_EPS) break;
            }
        } else if (algo == 5) {
            // SCALGM, :762-847
            int mode = 1;
            for (int k = 0; k < 100; ++k) {
                if (mode == 1) {
                    scalgm_rows<true>(M, rs, cs, nullptr, R1, lane);
                    __syncwarp();
                    scalgm_cols<true>(M, rs, R1, cs, C1, lane);
                    scalgm_cols<true>(M, rs, nullptr, cs, C2, lane);
                    __syncwarp();
                    scalgm_rows<true>(M, rs, cs, C2, R2, lane);
                    __syncwarp();
                    for (int j = lane; j < nrows; j += 32) rs[j] = mul(rs[j], sqr(mul(R1[j], R2[j])));
                    for (int m = lane; m < ncols; m += 32) cs[m] = mul(cs[m], sqr(mul(C1[m], C2[m])));
                    __syncwarp();
                }
                scalgm_rows<false>(M, rs, cs, nullptr, R1, lane);
                __syncwarp();
                scalgm_cols<false>(M, rs, R1, cs, C1, lane);
                scalgm_cols<false>(M, rs, nullptr, cs, C2, lane);
                __syncwarp();
                scalgm_rows<false>(M, rs, cs, C2, R2, lane);
                __syncwarp();
                for (int j = lane; j < nrows; j += 32) rs[j] = mul(rs[j], sqr(mul(R1[j], R2[j])));
                for (int m = lane; m < ncols; m += 32) cs[m] = mul(cs[m], sqr(mul(C1[m], C2[m])));
                __syncwarp();
                // convergence: every row and column norm (inf) within epsilon of 1
                double er = 0.0, ec = 0.0;
                for (int j = lane; j < nrows; j += 32) {
                    const double r = rs[j];
                    double acc = 0.0;
                    for (int m = 0; m < ncols; ++m) {
                        const double tmp = fabs(mul(M.at(j, m), mul(r, cs[m])));
                        if (tmp > acc) acc = tmp;
                    }
                    er = fmax(er, fabs(add(1.0, -acc)));
                }
                for (int m = lane; m < ncols; m += 32) {
                    const double c = cs[m];
                    double acc = 0.0;
                    for (int j = 0; j < nrows; ++j) {
                        const double tmp = fabs(mul(M.at(j, m), mul(c, rs[j])));
                        if (tmp > acc) acc = tmp;
                    }
                    ec = fmax(ec, fabs(add(1.0, -acc)));
                }
                er = warp_max_d(er);
                ec = warp_max_d(ec);
                if (er < SCALE_EPS && ec < SCALE_EPS) {
                    if (mode == 1) mode = 2;
                    else break;
                }
            }
        }
        __syncwarp();
        // apply_scaling_c (:293-299) and write back; a failed DGEEQU leaves the matrix as it was
        if (ok) {
            for (int t = lane; t < nrows * ncols; t += 32) {
                const int j = t % nrows, m = t / nrows;
                g[t] = mul(A[j + lda * m], mul(rs[j], cs[m]));
            }
        }
        for (int j = lane; j < nrows; j += 32) rsg[l * nrows + j] = rs[j];
        for (int m = lane; m < ncols; m += 32) csg[l * ncols + m] = cs[m];
        if (okg && lane == 0) okg[l] = ok;
        __syncwarp();
    }
}

cudaError_t launch_rescale(int nrows, int ncols, long long nlhs, double* A, int algo, double* row_scale, double* col_scale,
                           int* ok, cudaStream_t st) {
    if (nlhs == 0 || nrows == 0 || ncols == 0) return cudaSuccess;
    const int lda = nrows | 1;
    int wd = lda * ncols + 3 * (nrows + ncols);
    wd = (wd + 1) & ~1;
    const size_t per_warp = (size_t)wd * 8;
    if (per_warp > 227 * 1024) return cudaErrorInvalidValue;
    int warps = 8;
    while (warps > 1 && warps * per_warp > 227 * 1024) --warps;
    const size_t smem = warps * per_warp;
    cudaError_t e = cudaFuncSetAttribute(rescale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (nlhs + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    rescale_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(nrows, ncols, nlhs, A, algo, row_scale, col_scale, ok, wd);
    return cudaGetLastError();
}

}  // namespace wlsqm
