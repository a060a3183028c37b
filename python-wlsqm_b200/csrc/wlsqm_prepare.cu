// wlsqm_prepare.cu -- K1: per-case assembly, equilibration, pivoted LU and solution-operator build,
// with the matrix in registers and the contractions on the FP64 tensor cores (DMMA m8n8k4).
//
// One warp owns one fit.  Replaces, for a whole batch in one launch:
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
//
// Phases (per fit, one warp):
//   P1  lane = neighbour k: monomials c[k][s] -> shared table CT[s][k] (32-column blocks, row stride 36
//       doubles so that DMMA fragment loads are bank-conflict free), d^2 -> weights.
//   P2  Gram matrix G = C^T W C over all `no` slots on the FP64 tensor cores: for every 4 neighbours one
//       fragment load per 8-slot tile feeds T(T+1)/2 DMMAs (lower tile triangle, mirrored => exactly
//       symmetric).  The knowns columns -G[:, known] are appended to CT as extra right-hand sides.
//   P3  lane = row: the reduced matrix row lives in registers; Ruiz sweeps read the running scale vector
//       by broadcast LDS.128, one vote decides convergence; scaling in registers.
//   P4  LU with partial pivoting (first maximal |entry| in current row order, like dgetf2/idamax) without
//       moving rows: the pivot lane publishes its row to shared memory (which is also the final LAPACK-
//       layout LU storage), the others eliminate in registers.  Pivot search = three REDUX.
//   P5  lane = right-hand side q: y = P (row o (w_q c_q)) gathered through the pivot order, unit-lower
//       forward and upper backward substitution in registers against broadcast LU rows, column scale,
//       staged in place of the consumed CT block and written with one bulk (TMA) store per 32 rows.
#include <type_traits>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

template <int DIM, int ORD>
struct PK {
    static constexpr int NO = DIM == 1 ? ORD + 1
                                       : (DIM == 2 ? (ORD + 1) * (ORD + 2) / 2 : (ORD + 1) * (ORD + 2) * (ORD + 3) / 6);
    static constexpr int NOP = (NO + 7) & ~7;       // Gram tiles of 8 slots
    static constexpr int NRP = (NO + 3) & ~3;       // padded row length held in registers
    static constexpr int RPL = NRP > 32 ? 2 : 1;    // matrix rows per lane
    static constexpr int T = NOP / 8;
    static constexpr int LDA = NOP + 2;             // row stride of G / LU: LDS.128 by lane = row is conflict free
    static constexpr int BLK = NOP * PREP_CB;       // doubles per 32-column block of CT
};

__host__ __device__ constexpr int prep_no(int dim, int ord) {
    return dim == 1 ? ord + 1 : (dim == 2 ? (ord + 1) * (ord + 2) / 2 : (ord + 1) * (ord + 2) * (ord + 3) / 6);
}

int prep_reg_warp_doubles(int dim, int maxorder, int nb) {
    const int nop = (prep_no(dim, maxorder) + 7) & ~7;
    const int d = nb * nop * PREP_CB + nop * (nop + 2) + nb * 32 + 40 + 40 + 80 + 20;
    return (d + 15) & ~15;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1])
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

template <int DIM, int ORD>
__global__ void __launch_bounds__(PREP_REG_THREADS) prepare_reg_kernel(PrepRegParams P) {
    using K = PK<DIM, ORD>;
    constexpr int NOP = K::NOP, NRP = K::NRP, RPL = K::RPL, T = K::T, LDA = K::LDA, CB = PREP_CB, BLK = K::BLK;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double* wb = smem + (size_t)warp * P.warp_doubles;
    double* CT = wb;                                  // [nb][NOP][CB]  monomials (transposed), later the staged operator
    double* G = CT + P.nb * BLK;                      // [NOP][LDA]     Gram matrix, later the LU factors (pivot order)
    double* W = G + NOP * LDA;                        // [nb*32]        weights of the right-hand-side columns
    double* RS = W + P.nb * 32;                       // [40]           row (= column) scale, natural order
    double* DINV = RS + 40;                           // [40]           reciprocal pivots
    double2* REC = reinterpret_cast<double2*>(DINV + 40);   // [40]     pivot order: {row scale, CT row offset}
    int* R2O = reinterpret_cast<int*>(DINV + 40 + 80);      // [40]     reduced -> original slot, then the known slots

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
        const int nk = mt.nk, no = mt.no, nr = mt.nr, nkn = mt.nkn, nq = nk + nkn;
        const long long knowns = mt.knowns;
        if (nr < 1) continue;   // everything known: silent no-op (impl.pyx:574,636,742)
        const int nkp = (nk + 3) & ~3;
        const int nblk = (max(nkp, nq) + 31) >> 5;

        // the previous fit's bulk stores read this warp's CT blocks
        if (lane == 0) tma_store_wait_read();
        __syncwarp();

        // ---- P1. monomials and squared distances (lane = neighbour) ----------------------------
        double xi0 = P.xi[c * P.xi_s0], xi1 = 0.0, xi2 = 0.0;
        if (DIM >= 2) xi1 = P.xi[c * P.xi_s0 + 1];
        if (DIM >= 3) xi2 = P.xi[c * P.xi_s0 + 2];
        double max_d2 = 0.0;
        for (int b = 0; b < nblk; ++b) {
            const int k = b * 32 + lane;
            const bool in = k < nk;
            double dx = 0.0, dy = 0.0, dz = 0.0;
            if (in) {
                const double* xp = P.xk + c * P.xk_s0 + (long long)k * P.xk_s1;
                dx = xp[0] - xi0;
                if (DIM >= 2) dy = xp[1] - xi1;
                if (DIM >= 3) dz = xp[2] - xi2;
            }
            double d2 = dx * dx;
            if (DIM >= 2) d2 += dy * dy;
            if (DIM >= 3) d2 += dz * dz;
            max_d2 = fmax(max_d2, d2);
            W[k] = d2;
            const Pow5 px = scaled_powers(dx), py = scaled_powers(dy), pz = scaled_powers(dz);
            double* ctb = CT + b * BLK + lane;
            static_for<0, NOP>([&](auto I) {
                constexpr int s = decltype(I)::value;
                double v = 0.0;
                if constexpr (s < K::NO) {
                    if (in && s < no) v = monomial<DIM, s>(px, py, pz);
                }
                ctb[s * CB] = v;
            });
        }
        max_d2 = warp_max(max_d2);
        // ---- weights (infra.pyx:679-702); columns >= nk carry weight 0 through the Gram phase ----
        for (int b = 0; b < nblk; ++b) {
            const int k = b * 32 + lane;
            double w = 0.0;
            if (k < nk) {
                w = 1.0;
                if (mt.wm == WLSQM_WEIGHT_CENTER) {
                    const double t = 1.0 - sqrt(W[k] / max_d2);
                    w = 1e-4 + (1.0 - 1e-4) * (t * t);
                }
            }
            W[k] = w;
        }
        for (int o = lane; o < no; o += 32) {
            const long long below = knowns & ((1LL << o) - 1);
            if (!((knowns >> o) & 1LL)) R2O[o - __popcll(below)] = o;        // unknown: reduced index
            else R2O[nr + __popcll(below)] = o;                              // known slots, ascending, after the unknowns
        }
        __syncwarp();

        // ---- P2. G = C^T W C on the FP64 tensor cores --------------------------------------------
        {
            double acc[T * (T + 1) / 2][2];
#pragma unroll
            for (int t = 0; t < T * (T + 1) / 2; ++t) acc[t][0] = acc[t][1] = 0.0;
            const int kk = lane & 3, jj = lane >> 2;
            for (int k0 = 0; k0 < nkp; k0 += 4) {
                const int k = k0 + kk;
                const double* ctk = CT + (k >> 5) * BLK + (k & 31) + jj * CB;
                const double w = W[k];
                double cf[T], wf[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    cf[t] = ctk[8 * t * CB];
                    wf[t] = w * cf[t];
                }
#pragma unroll
                for (int tj = 0; tj < T; ++tj)
#pragma unroll
                    for (int tm = 0; tm <= tj; ++tm) dmma884(acc[tj * (tj + 1) / 2 + tm], cf[tj], wf[tm]);
            }
            // lower triangle, mirrored (make_A computes (w c_m) c_j for the full square; the mirror makes
            // the matrix exactly symmetric, which the single-pass Ruiz sweep below relies on)
#pragma unroll
            for (int tj = 0; tj < T; ++tj)
#pragma unroll
                for (int tm = 0; tm <= tj; ++tm)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = 8 * tj + jj, col = 8 * tm + 2 * kk + e;
                        const double v = acc[tj * (tj + 1) / 2 + tm][e];
                        if (row >= col) {
                            G[row * LDA + col] = v;
                            G[col * LDA + row] = v;
                        }
                    }
        }
        __syncwarp();
        // knowns elimination columns: right-hand sides -A[oj, known om] (impl.pyx:792-818), weight 1
        if (nkn) {
            for (int t = lane; t < no * nkn; t += 32) {
                const int s = t % no, mk = t / no;
                const int q = nk + mk;
                CT[(q >> 5) * BLK + s * CB + (q & 31)] = -G[s * LDA + R2O[nr + mk]];
            }
            if (lane < nkn) W[nk + lane] = 1.0;
            __syncwarp();
        }

        // ---- P3. reduced matrix rows -> registers; Ruiz equilibration ------------------------------
        double a[RPL][NRP];
        double rj[RPL];
        int roff[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int j = lane + 32 * t;
            const bool valid = j < nr;
            const int oj = valid ? R2O[j] : 0;
            roff[t] = oj * CB;
            rj[t] = 1.0;
            const double* g = G + oj * LDA;
            if (knowns == 0) {
#pragma unroll
                for (int m = 0; m < NRP; m += 2) {
                    const double2 v = ld2(g + m);
                    a[t][m] = valid ? v.x : 0.0;
                    a[t][m + 1] = valid ? v.y : 0.0;
                }
            } else {
#pragma unroll
                for (int m = 0; m < NRP; ++m) {
                    const int om = m < nr ? R2O[m] : 0;
                    const double v = g[om];
                    a[t][m] = (valid && m < nr) ? v : 0.0;
                }
            }
        }
        for (int i = lane; i < 40; i += 32) RS[i] = 1.0;
        __syncwarp();
        // A is exactly symmetric, so the reference's row and column passes coincide (DR == DC); the running
        // reciprocal products row_j = 1/DRp_j are kept instead of dividing every entry.
        for (int it = 0; it < 100; ++it) {
            double mx[RPL];
#pragma unroll
            for (int t = 0; t < RPL; ++t) mx[t] = 0.0;
#pragma unroll
            for (int m = 0; m < NRP; m += 2) {
                const double2 r2 = ld2(RS + m);
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    mx[t] = fmax(mx[t], fabs(a[t][m]) * r2.x);
                    mx[t] = fmax(mx[t], fabs(a[t][m + 1]) * r2.y);
                }
            }
            __syncwarp();
            bool conv = true;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int j = lane + 32 * t;
                if (j < nr) {
                    const double m2 = mx[t] * rj[t];            // = DR_j^2, the scaled inf-norm of row j
                    conv = conv && (fabs(1.0 - m2) < 1e-15);
                    rj[t] *= rsqrt(m2);
                    RS[j] = rj[t];
                }
            }
            __syncwarp();
            if (__all_sync(FULL, conv)) break;
        }
        // ---- A <- diag(row) A diag(col)  (lapackdrivers.pyx:293-299) -----------------------------
#pragma unroll
        for (int m = 0; m < NRP; m += 2) {
            const double2 r2 = ld2(RS + m);
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                a[t][m] *= rj[t] * r2.x;
                a[t][m + 1] *= rj[t] * r2.y;
            }
        }
        if (P.As) {   // debug=True: keep the scaled matrix for conds() (impl.pyx:662-682)
            double* as = P.As + c * (long long)P.as_stride;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int j = lane + 32 * t;
#pragma unroll
                for (int m = 0; m < NRP; ++m)
                    if (j < nr && m < nr) as[j + nr * m] = a[t][m];
            }
        }

        // ---- P4. LU with partial pivoting; rows stay in their lanes ----------------------------------
        bool act[RPL];
        int pos[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            pos[t] = lane + 32 * t;
            act[t] = pos[t] < nr;
        }
#pragma unroll
        for (int p = 0; p < NRP; ++p) {
            if (p < nr) {
                // this lane's candidate: largest |a[.][p]| among its active rows, first in row order on ties
                unsigned khi = 0u, klo = 0u;
                int kpos = 0xffff, kt = 0;
                bool any = false;
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a[t][p]));
                    const unsigned hi_t = (unsigned)(bits >> 32), lo_t = (unsigned)bits;
                    const bool better = act[t] && (!any || hi_t > khi || (hi_t == khi && (lo_t > klo || (lo_t == klo && pos[t] < kpos))));
                    if (better) { khi = hi_t; klo = lo_t; kpos = pos[t]; kt = t; any = true; }
                }
                const unsigned hi = __reduce_max_sync(FULL, any ? khi : 0u);
                const bool c1 = any && khi == hi;
                const unsigned lo = __reduce_max_sync(FULL, c1 ? klo : 0u);
                const bool c2 = c1 && klo == lo;
                const unsigned ppos = __reduce_min_sync(FULL, c2 ? (unsigned)kpos : 0xffffu);
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    const bool piv = c2 && kt == t && (unsigned)pos[t] == ppos && act[t];
                    if (piv) {
                        double* row = G + p * LDA;
#pragma unroll
                        for (int m = 0; m < NRP; m += 2) st2(row + m, a[t][m], a[t][m + 1]);
                        REC[p] = make_double2(rj[t], __hiloint2double(0, roff[t]));
                        pos[t] = p;
                        act[t] = false;
                    } else if (pos[t] == p) {
                        pos[t] = (int)ppos;     // the row that sat at position p takes the pivot row's old place
                    }
                }
                __syncwarp();
                double u[NRP];
#pragma unroll
                for (int m = (p & ~1); m < NRP; m += 2) {
                    const double2 v = ld2(G + p * LDA + m);
                    u[m] = v.x;
                    u[m + 1] = v.y;
                }
                const double rp = 1.0 / u[p];
                if (lane == 0) DINV[p] = rp;
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    if (act[t]) {
                        const double l = a[t][p] * rp;
                        a[t][p] = l;
#pragma unroll
                        for (int m = p + 1; m < NRP; ++m) a[t][m] = fma(-l, u[m], a[t][m]);
                    }
                }
            }
        }
        __syncwarp();

        // ---- P5. operator columns: dgetrs per right-hand side (lane = column) ----------------------
        double* opg = P.op + mt.op_off;
        const int qblk = (nq + 31) >> 5;
        for (int b = 0; b < qblk; ++b) {
            const int q = b * 32 + lane;
            const double wq = W[q];
            const double* ctq = CT + b * BLK + lane;
            double y[NRP];
            // y = P (row o (w_q c_q)): right-hand side row_j w_q c[q,oj] (impl.pyx:769-779), in pivot order
#pragma unroll
            for (int i = 0; i < NRP; ++i) {
                y[i] = 0.0;
                if (i < nr) {
                    const double2 rec = REC[i];
                    y[i] = rec.x * (wq * ctq[__double2loint(rec.y)]);
                }
            }
            __syncwarp();   // every lane has consumed this CT block: it becomes the staging buffer
            // unit-lower forward substitution
#pragma unroll
            for (int i = 1; i < NRP; ++i) {
                if (i < nr) {
#pragma unroll
                    for (int p = 0; p < i; p += 2) {
                        const double2 l2 = ld2(G + i * LDA + p);
                        y[i] = fma(-l2.x, y[p], y[i]);
                        if (p + 1 < i) y[i] = fma(-l2.y, y[p + 1], y[i]);
                    }
                }
            }
            // upper backward substitution (padded columns hold zeros)
#pragma unroll
            for (int i = NRP - 1; i >= 0; --i) {
                if (i < nr) {
#pragma unroll
                    for (int m = ((i + 1) & ~1); m < NRP; m += 2) {
                        const double2 u2 = ld2(G + i * LDA + m);
                        if (m > i) y[i] = fma(-u2.x, y[m], y[i]);
                        y[i] = fma(-u2.y, y[m + 1], y[i]);
                    }
                    y[i] *= DINV[i];
                }
            }
            // Op[q][j] = x_j * col_j, staged as 32 consecutive operator rows and stored in one bulk copy
            double* ops = CT + b * BLK;
            if (q < nq) {
#pragma unroll
                for (int i = 0; i < NRP; ++i)
                    if (i < nr) ops[lane * nr + i] = y[i] * RS[i];
            }
            const int rows = min(32, nq - b * 32);
            if (lane == 0 && ((rows * nr) & 1)) ops[rows * nr] = 0.0;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_1d(opg + (long long)b * 32 * nr, ops, (uint32_t)(((rows * nr + 1) & ~1) * 8));
                tma_store_commit();
            }
        }
    }
    if (lane == 0) tma_store_wait_all();
}

// ---- launch -----------------------------------------------------------------------------------------
template <int DIM, int ORD>
static cudaError_t launch_t(const PrepRegParams& P, int blocks, int threads, size_t smem, cudaStream_t st,
                            int* occupancy) {
    cudaError_t e = cudaFuncSetAttribute(prepare_reg_kernel<DIM, ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    if (occupancy) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, prepare_reg_kernel<DIM, ORD>, threads, smem);
    prepare_reg_kernel<DIM, ORD><<<blocks, threads, smem, st>>>(P);
    return cudaGetLastError();
}

template <int DIM>
static cudaError_t launch_d(int ord, const PrepRegParams& P, int blocks, int threads, size_t smem, cudaStream_t st,
                            int* occupancy) {
    switch (ord) {
        case 0: return launch_t<DIM, 0>(P, blocks, threads, smem, st, occupancy);
        case 1: return launch_t<DIM, 1>(P, blocks, threads, smem, st, occupancy);
        case 2: return launch_t<DIM, 2>(P, blocks, threads, smem, st, occupancy);
        case 3: return launch_t<DIM, 3>(P, blocks, threads, smem, st, occupancy);
        default: return launch_t<DIM, 4>(P, blocks, threads, smem, st, occupancy);
    }
}

static cudaError_t dispatch(int dim, int ord, const PrepRegParams& P, int blocks, int threads, size_t smem,
                            cudaStream_t st, int* occupancy) {
    if (dim == 1) return launch_d<1>(ord, P, blocks, threads, smem, st, occupancy);
    if (dim == 2) return launch_d<2>(ord, P, blocks, threads, smem, st, occupancy);
    return launch_d<3>(ord, P, blocks, threads, smem, st, occupancy);
}

cudaError_t prepare_reg_occupancy(int dim, int maxorder, int threads, size_t smem, int* ctas_per_sm) {
    PrepRegParams P{};
    return dispatch(dim, maxorder, P, 0, threads, smem, nullptr, ctas_per_sm);
}

cudaError_t launch_prepare_reg(int dim, int maxorder, const PrepRegParams& P, int blocks, int threads, size_t smem,
                               cudaStream_t st) {
    return dispatch(dim, maxorder, P, blocks, threads, smem, st, nullptr);
}

}  // namespace wlsqm
