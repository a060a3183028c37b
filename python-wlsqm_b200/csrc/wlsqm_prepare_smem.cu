// wlsqm_prepare_smem.cu -- K1 (generic variant): per-case assembly, equilibration, pivoted LU and
// solution-operator build with the whole fit in shared memory.  Kept as the A/B baseline of the
// register/DMMA kernel in wlsqm_prepare.cu (WLSQM_PREP_KERNEL=smem selects it).
//
// One warp owns one fit; everything between the gather of xk and the store of the operator
// lives in that warp's slice of shared memory.  Replaces, for a whole batch in one launch:
//   make_c_{1,2,3}D      wlsqm/fitter/impl.pyx:449-544 / 286-432 / 70-269   (monomials, d^2)
//   Case_make_weights    wlsqm/fitter/infra.pyx:668-702                       (UNIFORM / CENTER)
//   remap                wlsqm/fitter/infra.pyx:145-200                       (knowns -> r2o)
//   make_A               wlsqm/fitter/impl.pyx:566-602                        (A = C_r^T W C_r)
//   preprocess_A         wlsqm/fitter/impl.pyx:620-689
//     rescale_ruiz2001_c wlsqm/utils/lapackdrivers.pyx:553-623  (inf-norm sqrt scaling, tol 1e-15, <=100 its)
//     apply_scaling_c    wlsqm/utils/lapackdrivers.pyx:293-299
//     generalfactor_c    wlsqm/utils/lapackdrivers.pyx:1433,1628-1635 -> dgetrf (partial pivoting)
// and then, instead of keeping (c, w, LU, ipiv, scales) for every later solve (impl.pyx:731-846),
// forms the dense solution operator of the case once:
//   Op[q][j], q < nk      : d fi[r2o[j]] / d fk[q]       (exactly the reference's `sens`, impl.pyx:769-779,831-846)
//   Op[nk + m][j]         : d fi[r2o[j]] / d fi[known m] (the knowns elimination of impl.pyx:792-818, solved once)
// Both are obtained with the *scaled* LU (dgetrs order: permute, unit-lower forward, upper backward),
// then multiplied by col_scale, i.e. column by column what the reference does for `sens`.
#include <type_traits>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

// lower-triangle pair index t -> (j, m) with j >= m, t = j(j+1)/2 + m
__device__ __forceinline__ void tri_decode(int t, int& j, int& m) {
    int jj = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((jj + 1) * (jj + 2) / 2 <= t) ++jj;
    while (jj * (jj + 1) / 2 > t) --jj;
    j = jj;
    m = t - jj * (jj + 1) / 2;
}

template <int DIM>
__global__ void __launch_bounds__(PREP_MAX_THREADS) prepare_smem_kernel(PrepareParams P) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double* wb = smem + (size_t)warp * P.warp_doubles;
    double* C = wb;                      // [nk][cs]   monomials, cs odd
    double* W = C + P.off_w;             // [nk]
    double* A = wb + P.off_a;            // [nr][lda]  column-major A[j + lda*m], lda odd
    double* RS = wb + P.off_rs;          // [nr] row (= column) scale
    double* S = wb + P.off_s;            // [nr][sq]   right-hand sides / solution, sq odd
    int* R2O = (int*)(wb + P.off_i);     // [no] reduced -> original slot
    int* IPIV = R2O + 36;                // [nr]

    const long long gw = (long long)blockIdx.x * nwarps + warp;
    const long long GW = (long long)gridDim.x * nwarps;

    for (long long c = gw; c < P.ncases; c += GW) {
        CaseMeta mt;
        if (P.meta) {
            mt = P.meta[c];
        } else {
            mt = P.uni;
            mt.op_off = c * P.op_stride;
        }
        const int nk = mt.nk, no = mt.no, nr = mt.nr, nkn = mt.nkn;
        const long long knowns = mt.knowns;
        const int cs = no | 1, lda = nr | 1, nq = nk + nkn, sq = nq | 1;
        if (nr < 1) continue;   // everything known: silent no-op (impl.pyx:574,636,742)

        // ---- 1. monomials and squared distances (lane = neighbour) -------------------------
        double xi0 = P.xi[c * P.xi_s0], xi1 = 0.0, xi2 = 0.0;
        if (DIM >= 2) xi1 = P.xi[c * P.xi_s0 + 1];
        if (DIM >= 3) xi2 = P.xi[c * P.xi_s0 + 2];
        double max_d2 = 0.0;
        for (int k = lane; k < nk; k += 32) {
            const double* xp = P.xk + c * P.xk_s0 + (long long)k * P.xk_s1;
            double dx = xp[0] - xi0, dy = 0.0, dz = 0.0;
            if (DIM >= 2) dy = xp[1] - xi1;
            if (DIM >= 3) dz = xp[2] - xi2;
            double d2 = dx * dx;
            if (DIM >= 2) d2 += dy * dy;
            if (DIM >= 3) d2 += dz * dz;
            max_d2 = fmax(max_d2, d2);
            W[k] = d2;
            const Pow5 px = scaled_powers(dx), py = scaled_powers(dy), pz = scaled_powers(dz);
            static_for<0, max_no<DIM>()>([&](auto I) {
                constexpr int s = decltype(I)::value;
                if (s < no) C[k * cs + s] = monomial<DIM, s>(px, py, pz);
            });
        }
        max_d2 = warp_max(max_d2);
        // ---- 2. weights (infra.pyx:679-702) and the reduced->original map --------------------
        for (int k = lane; k < nk; k += 32) {
            double w = 1.0;
            if (mt.wm == WLSQM_WEIGHT_CENTER) {
                double t = 1.0 - sqrt(W[k] / max_d2);
                w = 1e-4 + (1.0 - 1e-4) * (t * t);
            }
            W[k] = w;
        }
        for (int o = lane; o < no; o += 32) {
            const long long below = knowns & ((1LL << o) - 1);
            if (!((knowns >> o) & 1LL)) R2O[o - __popcll(below)] = o;        // unknown: reduced index
            else R2O[nr + __popcll(below)] = o;                              // known slots, ascending, after the unknowns
        }
        __syncwarp();

        // ---- 3. A = C_r^T W C_r (lower triangle, mirrored) and the knowns columns -------------
        const int ntri = nr * (nr + 1) / 2;
        for (int t = lane; t < ntri; t += 32) {
            int j, m;
            tri_decode(t, j, m);
            const int oj = R2O[j], om = R2O[m];
            double acc = 0.0;
            for (int k = 0; k < nk; ++k) acc += (W[k] * C[k * cs + om]) * C[k * cs + oj];
            A[j + lda * m] = acc;
            A[m + lda * j] = acc;
        }
        for (int t = lane; t < nr * nkn; t += 32) {     // raw A[oj, known om]; scaled by -row_j below
            const int j = t % nr, mk = t / nr;
            const int oj = R2O[j], om = R2O[nr + mk];
            double acc = 0.0;
            for (int k = 0; k < nk; ++k) acc += (W[k] * C[k * cs + om]) * C[k * cs + oj];
            S[j * sq + nk + mk] = acc;
        }
        for (int j = lane; j < nr; j += 32) RS[j] = 1.0;
        __syncwarp();

        // ---- 4. Ruiz equilibration.  A is exactly symmetric here, so the reference's row and
        //         column passes coincide (DR == DC) and one pass per sweep suffices; the running
        //         reciprocal products row_j = 1/DRp_j are kept instead of dividing every entry.
        for (int it = 0; it < 100; ++it) {
            double dr0 = 1.0, dr1 = 1.0, dev = 0.0;
            {
                const int j = lane;
                if (j < nr) {
                    double mx = 0.0;
                    const double rj = RS[j];
                    for (int m = 0; m < nr; ++m) mx = fmax(mx, fabs(A[j + lda * m]) * (rj * RS[m]));
                    dr0 = sqrt(mx);
                    dev = fabs(1.0 - mx);   // == |1 - DR^2| up to one rounding of sqrt
                }
            }
            if (nr > 32) {
                const int j = lane + 32;
                if (j < nr) {
                    double mx = 0.0;
                    const double rj = RS[j];
                    for (int m = 0; m < nr; ++m) mx = fmax(mx, fabs(A[j + lda * m]) * (rj * RS[m]));
                    dr1 = sqrt(mx);
                    dev = fmax(dev, fabs(1.0 - mx));
                }
            }
            __syncwarp();
            if (lane < nr) RS[lane] = RS[lane] / dr0;
            if (lane + 32 < nr) RS[lane + 32] = RS[lane + 32] / dr1;
            dev = warp_max(dev);
            __syncwarp();
            if (dev < 1e-15) break;
        }

        // ---- 5. A <- diag(row) A diag(col)  (lapackdrivers.pyx:293-299) -----------------------
        for (int t = lane; t < nr * nr; t += 32) {
            const int j = t % nr, m = t / nr;
            A[j + lda * m] *= RS[j] * RS[m];
        }
        __syncwarp();
        if (P.As) {   // debug=True: keep the scaled matrix for conds() (impl.pyx:662-682)
            double* as = P.As + c * (long long)P.as_stride;
            for (int t = lane; t < nr * nr; t += 32) as[t] = A[(t % nr) + lda * (t / nr)];
        }

        // ---- 6. LU with partial pivoting, in place (dgetf2 semantics: first maximal |entry|) ----
        for (int p = 0; p < nr; ++p) {
            // pivot search in column p, rows p..nr-1
            double best = -1.0;
            int bi = p;
            for (int i = p + lane; i < nr; i += 32) {
                const double v = fabs(A[i + lda * p]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) IPIV[p] = bi;
            if (bi != p) {
                for (int m = lane; m < nr; m += 32) {
                    const double t = A[p + lda * m];
                    A[p + lda * m] = A[bi + lda * m];
                    A[bi + lda * m] = t;
                }
            }
            __syncwarp();
            const double rp = 1.0 / A[p + lda * p];
            double l0 = 0.0, l1 = 0.0;
            const int i0 = p + 1 + lane, i1 = i0 + 32;
            if (i0 < nr) { l0 = A[i0 + lda * p] * rp; A[i0 + lda * p] = l0; }
            if (i1 < nr) { l1 = A[i1 + lda * p] * rp; A[i1 + lda * p] = l1; }
            for (int m = p + 1; m < nr; ++m) {
                const double u = A[p + lda * m];
                if (i0 < nr) A[i0 + lda * m] -= l0 * u;
                if (i1 < nr) A[i1 + lda * m] -= l1 * u;
            }
            __syncwarp();
        }

        // ---- 7. right-hand sides: row_j w_q c[q,oj] (impl.pyx:769-779) and -row_j A[oj,known] --
        for (int t = lane; t < nr * nq; t += 32) {
            const int q = t % nq, j = t / nq;
            double v;
            if (q < nk) v = RS[j] * (W[q] * C[q * cs + R2O[j]]);
            else v = -RS[j] * S[j * sq + q];
            S[j * sq + q] = v;
        }
        __syncwarp();
        // ---- 8. dgetrs per column (lane = column): permute, L forward, U backward -------------
        for (int q = lane; q < nq; q += 32) {
            for (int p = 0; p < nr; ++p) {
                const int ip = IPIV[p];
                if (ip != p) {
                    const double t = S[p * sq + q];
                    S[p * sq + q] = S[ip * sq + q];
                    S[ip * sq + q] = t;
                }
            }
            for (int p = 0; p < nr; ++p) {
                const double xp = S[p * sq + q];
                for (int i = p + 1; i < nr; ++i) S[i * sq + q] -= A[i + lda * p] * xp;
            }
            for (int p = nr - 1; p >= 0; --p) {
                const double xp = S[p * sq + q] / A[p + lda * p];
                S[p * sq + q] = xp;
                for (int i = 0; i < p; ++i) S[i * sq + q] -= A[i + lda * p] * xp;
            }
        }
        __syncwarp();
        // ---- 9. Op[q][j] = S[j][q] * col_j, streamed out as one contiguous block ---------------
        double* op = P.op + mt.op_off;
        for (int t = lane; t < nr * nq; t += 32) {
            const int j = t % nr, q = t / nr;
            op[t] = S[j * sq + q] * RS[j];
        }
        __syncwarp();
    }
}

template <int DIM>
static cudaError_t launch_dim(const PrepareParams& P, int blocks, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(prepare_smem_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prepare_smem_kernel<DIM><<<blocks, threads, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_prepare_smem(int dim, const PrepareParams& P, int blocks, int threads, size_t smem, cudaStream_t st) {
    if (dim == 1) return launch_dim<1>(P, blocks, threads, smem, st);
    if (dim == 2) return launch_dim<2>(P, blocks, threads, smem, st);
    return launch_dim<3>(P, blocks, threads, smem, st);
}

}  // namespace wlsqm
