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
    static constexpr int LPF = NRP <= 8 ? 8 : (NRP <= 16 ? 16 : 32);   // lanes per fit in the row phases (P3, P4)
    static constexpr int FPW = 32 / LPF;            // fits a warp carries through the row phases together
    static constexpr int T = NOP / 8;
    static constexpr int LDA = NOP + 2;             // row stride of G / LU: LDS.128 by lane = row is conflict free
    static constexpr int BLK = NOP * PREP_CB;       // doubles per 32-column block of CT
    // launch bound: 640 threads = 96 registers (at most 24 B of spills in any instantiation) -- the kernel is latency-bound
    // and 20 resident warps instead of 16 is what shared memory allows for 2D order 4 (measured: 3.96 -> 3.74 ms per 1M
    // fits); models of at most 8 DOFs (four fits per warp, ~7 KB of shared memory per warp): 1024 threads = 64 registers
    // without spills, 28 resident warps (2M fits 1D order 3: 2.29 ms at 16 warps, 1.96 at 20, 1.77 at 24, 1.60 at 28);
    // two rows per lane (3D order 4): 256 threads = 255 registers
    static constexpr int MAXT = RPL == 2 ? 256 : (NRP <= 8 ? PREP_REG_SMALL_THREADS : PREP_REG_MAX_THREADS);
    static constexpr int MINB = 1;
};

__host__ __device__ constexpr int prep_no(int dim, int ord) {
    return dim == 1 ? ord + 1 : (dim == 2 ? (ord + 1) * (ord + 2) / 2 : (ord + 1) * (ord + 2) * (ord + 3) / 6);
}

// the launch bound of the instantiation for (dim, order), for the host's choice of the CTA size
int prep_reg_max_threads(int dim, int order) {
    const int nrp = (prep_no(dim, order) + 3) & ~3;
    return nrp > 32 ? 256 : (nrp <= 8 ? PREP_REG_SMALL_THREADS : PREP_REG_MAX_THREADS);
}

int prep_reg_fits_per_warp(int dim, int maxorder) {
    const int nrp = (prep_no(dim, maxorder) + 3) & ~3;
    return nrp <= 8 ? 4 : (nrp <= 16 ? 2 : 1);
}

// shared-memory doubles that stay with ONE fit from the Gram phase to the operator store:
//   G [NOP][NOP+2] | W [nb*32] | RS [NRP] | DINV [NRP] | REC [NRP x 2] | R2O [NOP ints] | KN [nkn_max][NOP]
int prep_reg_fit_doubles(int dim, int maxorder, int nb, int nkn_max) {
    const int no = prep_no(dim, maxorder), nop = (no + 7) & ~7, nrp = (no + 3) & ~3;
    const int d = nop * (nop + 2) + nb * 32 + 4 * nrp + nop / 2 + nkn_max * nop;
    return (d + 1) & ~1;
}
// per warp: ONE block of the monomial table CT [NOP][PREP_CB] (transient: the Gram phase walks the neighbours 32 at a
// time; then right-hand-side gather and operator staging) followed by prep_reg_fits_per_warp() fit regions.  The
// staging of 32 operator rows needs 32 * nr <= NOP * PREP_CB doubles.
int prep_reg_warp_doubles(int dim, int maxorder, int nb, int nkn_max) {
    const int nop = (prep_no(dim, maxorder) + 7) & ~7;
    const int d = nop * PREP_CB + prep_reg_fits_per_warp(dim, maxorder) * prep_reg_fit_doubles(dim, maxorder, nb, nkn_max);
    return (d + 15) & ~15;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1])
        : "d"(a), "d"(b));
}

// max of two non-negative, non-NaN doubles: one DSETP + two selects (fmax() costs ~8 instructions for its NaN rules)
__device__ __forceinline__ double max_nn(double a, double b) { return b > a ? b : a; }

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

// max / min over the LPF consecutive lanes of this lane's group (grp = lane / LPF): one warp-wide REDUX per
// group (they are independent, so their latencies overlap) instead of a dependent shuffle butterfly
template <int LPF>
__device__ __forceinline__ unsigned group_max(unsigned v, int grp) {
    if constexpr (LPF == 32) return __reduce_max_sync(0xffffffffu, v);
    unsigned r = 0u;
#pragma unroll
    for (int g = 0; g < 32 / LPF; ++g) {
        const unsigned m = __reduce_max_sync(0xffffffffu, grp == g ? v : 0u);
        if (grp == g) r = m;
    }
    return r;
}
template <int LPF>
__device__ __forceinline__ unsigned group_min(unsigned v, int grp) {
    if constexpr (LPF == 32) return __reduce_min_sync(0xffffffffu, v);
    unsigned r = 0u;
#pragma unroll
    for (int g = 0; g < 32 / LPF; ++g) {
        const unsigned m = __reduce_min_sync(0xffffffffu, grp == g ? v : 0xffffffffu);
        if (grp == g) r = m;
    }
    return r;
}

// DIRECT = one-shot fit (fit_*_many without sens, simple.pyx:731-1170): no operator is formed or stored.  The data
// fk rides through the Gram phase in the unused padding slot NO of the monomial table, which makes
// b_s = sum_k w_k f_k c[k][s] row NO of the Gram matrix; b is eliminated along with the matrix in the LU phase
// (one extra column) and back-substituted across the lanes, and fi is written directly.
template <int DIM, int ORD, bool DIRECT>
__global__ void __launch_bounds__(PK<DIM, ORD>::MAXT, PK<DIM, ORD>::MINB) prepare_reg_kernel(PrepRegParams P) {
    using K = PK<DIM, ORD>;
    static_assert(!DIRECT || K::NOP > K::NO, "the one-shot path needs a free padding slot in the monomial table");
    constexpr int NOP = K::NOP, NRP = K::NRP, RPL = K::RPL, T = K::T, LDA = K::LDA, CB = PREP_CB, BLK = K::BLK;
    constexpr int LPF = K::LPF, FPW = K::FPW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned GMASK = LPF == 32 ? 0xffffffffu : ((1u << (LPF & 31)) - 1u);
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double* CT = smem + (size_t)warp * P.warp_doubles;   // [NOP][CB] monomials of 32 neighbours (transposed); P5: gather + staging
    double* fits = CT + BLK;
    // carve-up of one fit's region (offsets in doubles)
    constexpr int oG = 0;                      // [NOP][LDA]  Gram matrix, later the LU factors (pivot order)
    constexpr int oW = oG + NOP * LDA;         // [nb*32]     weights of the right-hand-side columns
    const int oRS = oW + P.nb * 32;            // [NRP]       row (= column) scale, natural order
    const int oDINV = oRS + NRP;               // [NRP]       reciprocal pivots (Ruiz: second scale buffer)
    const int oREC = oDINV + NRP;              // [NRP] x 16 B pivot order: {row scale, CT row offset}
    const int oR2O = oREC + 2 * NRP;           // [NOP] ints  reduced -> original slot, then the known slots
    const int oKN = oR2O + NOP / 2;            // [nkn_max][NOP] knowns right-hand sides -A[:, known]
    // row phases: this lane's fit (group) and row inside it
    const int grp = lane / LPF, jl = lane % LPF;
    double* gb = fits + grp * P.fit_doubles;
    double* gG = gb + oG;
    double* gRS = gb + oRS;
    double* gDINV = gb + oDINV;
    double2* gREC = reinterpret_cast<double2*>(gb + oREC);
    const int* gR2O = reinterpret_cast<const int*>(gb + oR2O);

    // launches over a case list (batches of mixed orders run one instantiation per order): entry -> case
    auto case_of = [&](long long v) -> long long { return P.perm ? (long long)P.perm[v] : v; };
    auto get_meta = [&](long long c) {
        CaseMeta m;
        if (P.meta && !P.geom_uniform) {
            m = P.meta[c];
        } else {
            m = P.uni;
            m.op_off = c * P.op_stride;
            if (P.meta) m.knowns = P.meta[c].knowns;   // same sizes everywhere, only the knowns pattern varies
        }
        return m;
    };
    // monomials of neighbour k of case c into column `lane` of one CT block; returns d^2 (0 outside the hood)
    auto monomial_column = [&](long long c, int k, int nk, int no, double* ctb) -> double {
        const bool in = k < nk;
        double dx = 0.0, dy = 0.0, dz = 0.0;
        if (in) {
            const double* xp = P.xk + c * P.xk_s0 + (long long)k * P.xk_s1;
            const double* xo = P.xi + c * P.xi_s0;
            dx = xp[0] - xo[0];
            if (DIM >= 2) dy = xp[DIM >= 2 ? 1 : 0] - xo[DIM >= 2 ? 1 : 0];
            if (DIM >= 3) dz = xp[DIM >= 3 ? 2 : 0] - xo[DIM >= 3 ? 2 : 0];
        }
        double d2 = dx * dx;
        if (DIM >= 2) d2 += dy * dy;
        if (DIM >= 3) d2 += dz * dz;
        const Pow5 px = scaled_powers(dx), py = scaled_powers(dy), pz = scaled_powers(dz);
        if (no == K::NO) {
            // the model of this instantiation's own order (every uniform batch, every per-order launch): no test per
            // slot.  Columns outside the neighbourhood hold the monomials of a zero offset -- finite values that their
            // zero weight removes from every sum.
            static_for<0, NOP>([&](auto I) {
                constexpr int s = decltype(I)::value;
                double v = 0.0;
                if constexpr (s < K::NO) v = monomial<DIM, s>(px, py, pz);
                if constexpr (DIRECT && s == K::NO) {
                    if (in) v = P.fk[c * P.fk_s0 + (long long)k * P.fk_s1];
                }
                ctb[s * CB] = v;
            });
        } else {
            static_for<0, NOP>([&](auto I) {
                constexpr int s = decltype(I)::value;
                double v = 0.0;
                if constexpr (s < K::NO) {
                    if (in && s < no) v = monomial<DIM, s>(px, py, pz);
                }
                if constexpr (DIRECT && s == K::NO) {
                    if (in) v = P.fk[c * P.fk_s0 + (long long)k * P.fk_s1];
                }
                ctb[s * CB] = v;
            });
        }
        return d2;
    };

    const long long w0 = ((long long)blockIdx.x * nwarps + warp) * FPW;
    const long long WS = (long long)gridDim.x * nwarps * FPW;
    bool pending = false;    // a bulk store may still be reading CT

    // Phase-synchronous mode (P.phase_sync): the warps of a CTA walk through the phases together (a CTA barrier at
    // every phase boundary), so that they execute the same few KB of this kernel's ~75 KB of straight-line code at
    // any time and share instruction-cache lines.  Every warp of the CTA runs the same number of iterations; a
    // warp past the end of the batch recomputes the last fits (identical values are stored twice).
    const long long cta_w0 = (long long)blockIdx.x * nwarps * FPW;
    const long long n_it = cta_w0 < P.ncases ? (P.ncases - cta_w0 + WS - 1) / WS : 0;
    const long long c_last = ((P.ncases - 1) / FPW) * FPW;
    auto phase_barrier = [&]() { if (P.phase_sync == 1) __syncthreads(); };
    for (long long it = 0; it < n_it; ++it) {
        long long c0 = w0 + it * WS;
        if (c0 >= P.ncases) {
            if (!P.phase_sync) break;
            c0 = c_last;
        }
        if (P.phase_sync) __syncthreads();      // (mode 2: once per iteration only)
        if (pending) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            pending = false;
        }

        // ================= per fit, whole warp: P1 monomials + weights, P2 Gram matrix =================
#pragma unroll 1
        for (int f = 0; f < FPW; ++f) {
            if (c0 + f >= P.ncases) break;
            const long long c = case_of(c0 + f);
            const CaseMeta mt = get_meta(c);
            const int nk = mt.nk, no = mt.no, nr = mt.nr, nkn = mt.nkn, nq = nk + nkn;
            const long long knowns = mt.knowns;
            if (nr < 1) continue;   // everything known: silent no-op (impl.pyx:574,636,742)
            const int nkp = (nk + 3) & ~3;
            const int nblk = (max(nkp, nq) + 31) >> 5;
            double* fb = fits + f * P.fit_doubles;
            double* G = fb + oG;
            double* W = fb + oW;
            int* R2O = reinterpret_cast<int*>(fb + oR2O);

            // ---- P1a. squared distances and weights (lane = neighbour; infra.pyx:679-702) ---------------
            // columns >= nk carry weight 0 through the Gram phase
            double max_d2 = 0.0;
#pragma unroll 1
            for (int b = 0; b < nblk; ++b) {
                const int k = b * 32 + lane;
                double d2 = 0.0;
                if (k < nk) {
                    const double* xp = P.xk + c * P.xk_s0 + (long long)k * P.xk_s1;
                    const double* xo = P.xi + c * P.xi_s0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        const double dd = xp[d] - xo[d];
                        d2 += dd * dd;
                    }
                }
                max_d2 = max_nn(max_d2, d2);
                W[k] = d2;
            }
            if (mt.wm == WLSQM_WEIGHT_CENTER) {
                max_d2 = warp_max(max_d2);
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    const int k = b * 32 + lane;
                    double w = 0.0;
                    if (k < nk) {
                        const double t = 1.0 - sqrt(W[k] / max_d2);
                        w = 1e-4 + (1.0 - 1e-4) * (t * t);
                    }
                    W[k] = w;
                }
            } else {
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    const int k = b * 32 + lane;
                    W[k] = k < nk ? 1.0 : 0.0;
                }
            }
            for (int o = lane; o < no; o += 32) {
                const long long below = knowns & ((1LL << o) - 1);
                if (!((knowns >> o) & 1LL)) R2O[o - __popcll(below)] = o;        // unknown: reduced index
                else R2O[nr + __popcll(below)] = o;                              // known slots, ascending, after the unknowns
            }

            // ---- P1b + P2, one block of 32 neighbours at a time: monomials c[k][s] into the (single) CT block, then
            //      G += C^T W C on the FP64 tensor cores; the accumulators stay in registers across the blocks ----------
            {
                double acc[T * (T + 1) / 2][2];
#pragma unroll
                for (int t = 0; t < T * (T + 1) / 2; ++t) acc[t][0] = acc[t][1] = 0.0;
                const int kk = lane & 3, jj = lane >> 2;
                const int nkblk = (nkp + 31) >> 5;
#pragma unroll 1
                for (int b = 0; b < nkblk; ++b) {
                    __syncwarp();                                   // the previous block's fragments have been read
                    monomial_column(c, b * 32 + lane, nk, no, CT + lane);
                    __syncwarp();
                    const int kend = min(32, nkp - b * 32);
                    for (int k0 = 0; k0 < kend; k0 += 4) {
                        const int k = k0 + kk;
                        const double* ctk = CT + k + jj * CB;
                        const double w = W[b * 32 + k];
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
                }
                // lower triangle, mirrored (make_A computes (w c_m) c_j for the full square; the mirror makes
                // the matrix exactly symmetric, which the equilibration below relies on)
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
            // knowns elimination: right-hand sides -A[oj, known om] (impl.pyx:792-818), weight 1
            if (!DIRECT && nkn) {
                double* KN = fb + oKN;
                for (int t = lane; t < no * nkn; t += 32) {
                    const int s = t % no, mk = t / no;
                    KN[mk * NOP + s] = -G[s * LDA + R2O[nr + mk]];
                }
                for (int m = lane; m < nkn; m += 32) W[nk + m] = 1.0;      // (up to 34 known slots: 3D order 4)
                __syncwarp();
            }
        }

        // ================= row phases: LPF lanes per fit, FPW fits side by side ========================
        // ---- P3. reduced matrix rows -> registers; Ruiz equilibration ------------------------------
        long long cg = c0 + grp;
        int nr = 0;
        long long knowns = 0;
        if (cg < P.ncases) {
            cg = case_of(cg);
            const CaseMeta mg = get_meta(cg);
            nr = mg.nr;
            knowns = mg.knowns;
        }
        const int nrmax = FPW == 1 ? nr : (int)__reduce_max_sync(FULL, (unsigned)nr);
        phase_barrier();
        if (nrmax < 1) { phase_barrier(); phase_barrier(); continue; }
        double a[RPL][NRP];
        double rj[RPL];
        int roff[RPL], ojs[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int j = jl + 32 * t;
            const bool valid = j < nr;
            const int oj = valid ? gR2O[j] : 0;
            ojs[t] = oj;
            roff[t] = oj * CB;
            rj[t] = 1.0;
            const double* g = gG + oj * LDA;
            if (knowns == 0) {
#pragma unroll
                for (int m = 0; m < NRP; m += 2) {
                    const double2 v = ld2(g + m);
                    a[t][m] = valid ? v.x : 0.0;
                    a[t][m + 1] = valid ? v.y : 0.0;
                }
                // one-shot fit: the padding slot carries the right-hand side, it is not a matrix column
                if constexpr (DIRECT && K::NO < NRP) a[t][K::NO] = 0.0;
            } else {
#pragma unroll
                for (int m = 0; m < NRP; ++m) {
                    const int om = m < nr ? gR2O[m] : 0;
                    const double v = g[om];
                    a[t][m] = (valid && m < nr) ? v : 0.0;
                }
            }
        }
        // one-shot fit: this row's right-hand side, knowns eliminated (impl.pyx:792-818): b_j = G[oj][NO] - sum_m fi[om] A[oj][om]
        double bj[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) bj[t] = 0.0;
        if constexpr (DIRECT) {
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                if (jl + 32 * t < nr) {
                    const double* g = gG + ojs[t] * LDA;
                    bj[t] = g[K::NO];
                    const int nkn_g = __popcll(knowns);
                    for (int mk = 0; mk < nkn_g; ++mk) {
                        const int om = gR2O[nr + mk];
                        bj[t] = fma(-P.fi[cg * P.fi_s0 + om], g[om], bj[t]);
                    }
                }
            }
        }
        // Equilibration (rescale_ruiz2001_c, lapackdrivers.pyx:553-623): the reference iterates r_j <- r_j / sqrt(scaled
        // inf-norm of row j) until every row norm is within 1e-15 of 1 (8-11 sweeps over the matrix).  A is symmetric (row
        // and column passes coincide) and positive definite, and for such a matrix that iteration has exactly ONE fixed
        // point: |s_jm| < sqrt(s_jj s_mm) (Cauchy-Schwarz) means a row maximum of 1 can only sit on the diagonal, so
        // s_jj = 1 for all j, i.e. r_j = a_jj^(-1/2).  The fixed point is evaluated directly (the reference's converged
        // scaled matrices have diag = 1 +- 2e-15 on every seeded and golden case).  The sweeps themselves live on in the
        // shared-memory variant of this kernel (wlsqm_prepare_smem.cu, WLSQM_PREP_KERNEL=smem), against which
        // tests/test_gpu_variants.py compares this one.  Singular matrices (nk < nr) have no unique fixed point -- and no
        // meaningful fit in the reference either.
        {
            for (int i = jl; i < NRP; i += LPF) gRS[i] = 1.0;       // (padding columns: finite scale for the zero entries)
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int j = jl + 32 * t;                          // (the same lane wrote gRS[j] just above)
                if (j < nr) {
                    rj[t] = rsqrt(gG[ojs[t] * LDA + ojs[t]]);
                    gRS[j] = rj[t];
                }
            }
            __syncwarp();
        }
        // ---- A <- diag(row) A diag(col)  (lapackdrivers.pyx:293-299) -----------------------------
#pragma unroll
        for (int m = 0; m < NRP; m += 2) {
            const double2 r2 = ld2(gRS + m);
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                a[t][m] *= rj[t] * r2.x;
                a[t][m + 1] *= rj[t] * r2.y;
            }
        }
        if constexpr (DIRECT) {
#pragma unroll
            for (int t = 0; t < RPL; ++t) bj[t] *= rj[t];
        }
        if (P.As && nr > 0) {   // debug=True: keep the scaled matrix for conds() (impl.pyx:662-682)
            double* as = P.As + cg * (long long)P.as_stride;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int j = jl + 32 * t;
#pragma unroll
                for (int m = 0; m < NRP; ++m)
                    if (j < nr && m < nr) as[j + nr * m] = a[t][m];
            }
        }

        phase_barrier();
        // ---- P4. LU with partial pivoting; rows stay in their lanes ----------------------------------
        bool act[RPL];
        int pos[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            pos[t] = jl + 32 * t;
            act[t] = pos[t] < nr;
        }
#pragma unroll
        for (int p = 0; p < NRP; ++p) {
            if (p < nrmax) {
                // every row's reciprocal of its entry in this column, BEFORE the pivot is known: the division overlaps the
                // pivot search instead of following the publication of the pivot row (the LU is a chain of such steps)
                double rcp[RPL];
#pragma unroll
                for (int t = 0; t < RPL; ++t) rcp[t] = 1.0 / a[t][p];
                // pivot = largest |a[.][p]| among the active rows, first in current row order on ties (idamax)
                bool piv[RPL];
                bool has_piv;
                unsigned ppos;
                if constexpr (RPL == 1) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a[0][p]));
                    const unsigned khi = (unsigned)(bits >> 32), klo = (unsigned)bits;
                    const unsigned hi = group_max<LPF>(act[0] ? khi : 0u, grp);
                    const bool c1 = act[0] && khi == hi;
                    unsigned cand = (__ballot_sync(FULL, c1) >> (grp * LPF)) & GMASK;
                    if (__any_sync(FULL, __popc(cand) > 1)) {   // rare: rows share the leading 32 bits -> exact comparison
                        const unsigned lo = group_max<LPF>(c1 ? klo : 0u, grp);
                        const bool c2 = c1 && klo == lo;
                        const unsigned best = group_min<LPF>(c2 ? (unsigned)pos[0] : 0xffffu, grp);
                        cand = (__ballot_sync(FULL, c2 && (unsigned)pos[0] == best) >> (grp * LPF)) & GMASK;
                    }
                    has_piv = cand != 0u;
                    const int pl = grp * LPF + (has_piv ? __ffs(cand) - 1 : 0);
                    ppos = (unsigned)__shfl_sync(FULL, pos[0], pl);
                    piv[0] = has_piv && lane == pl;
                } else {
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
                    ppos = __reduce_min_sync(FULL, c2 ? (unsigned)kpos : 0xffffu);
                    has_piv = true;
#pragma unroll
                    for (int t = 0; t < RPL; ++t) piv[t] = c2 && kt == t && (unsigned)pos[t] == ppos && act[t];
                }
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    if (piv[t]) {
                        double* row = gG + p * LDA;
#pragma unroll
                        for (int m = 0; m < NRP; m += 2) st2(row + m, a[t][m], a[t][m + 1]);
                        if constexpr (DIRECT) row[NOP] = bj[t];
                        gREC[p] = make_double2(rj[t], __hiloint2double(0, roff[t]));
                        gDINV[p] = rcp[t];
                        pos[t] = p;
                        act[t] = false;
                    } else if (has_piv && pos[t] == p) {
                        pos[t] = (int)ppos;     // the row that sat at position p takes the pivot row's old place
                    }
                }
                __syncwarp();
                const double rp = gDINV[p];      // = 1 / (the pivot), published with its row
                double l[RPL];
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    l[t] = a[t][p] * rp;
                    if (act[t]) a[t][p] = l[t];
                    if constexpr (DIRECT) {
                        if (act[t]) bj[t] = fma(-l[t], gG[p * LDA + NOP], bj[t]);
                    }
                }
#pragma unroll
                for (int m = ((p + 1) & ~1); m < NRP; m += 2) {
                    const double2 u2 = ld2(gG + p * LDA + m);
#pragma unroll
                    for (int t = 0; t < RPL; ++t) {
                        // unconditional: a row that has been the pivot lives in shared memory from then on, its
                        // registers are never read again (no predicate / branch per update keeps the code small)
                        if (m > p) a[t][m] = fma(-l[t], u2.x, a[t][m]);
                        a[t][m + 1] = fma(-l[t], u2.y, a[t][m + 1]);
                    }
                }
            }
        }
        __syncwarp();
        if constexpr (!DIRECT) {
            // rows nr .. NRP-1 of the pivot-order LU storage and their reciprocal pivots: zeros (they held Gram entries),
            // so that the operator phase can run its substitutions over the padded size without a branch per row
            if (nr > 0) {
                for (int t = jl; t < (NRP - nr) * NRP; t += LPF) gG[(nr + t / NRP) * LDA + (t % NRP)] = 0.0;
                for (int t = nr + jl; t < NRP; t += LPF) gDINV[t] = 0.0;
            }
            __syncwarp();
        }

        phase_barrier();
        if constexpr (DIRECT && RPL == 2) {
            // ---- U x = y for more rows than lanes (3D order 4): the right-hand side sits in column NOP of the published
            //      LU (pivot order); from the bottom row up, the whole warp forms the dot product of a row with the part
            //      of the solution already known ----
            for (int i = nr - 1; i >= 0; --i) {
                double part = 0.0;
                for (int m = i + 1 + lane; m < nr; m += 32) part = fma(gG[i * LDA + m], gG[m * LDA + NOP], part);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                if (lane == 0) gG[i * LDA + NOP] = (gG[i * LDA + NOP] - part) / gG[i * LDA + i];     // dgetrs divides by the pivot
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int j = jl + 32 * t;
                if (j < nr) P.fi[cg * P.fi_s0 + gR2O[j]] = gG[j * LDA + NOP] * gRS[j];
            }
            __syncwarp();
            continue;
        } else if constexpr (DIRECT) {
            // ---- U x = y across the lanes: lane jl takes row jl of the pivot order (re-read from the published LU) ----
            double u[NRP];
            double y = 0.0;
            if (jl < nr) {
#pragma unroll
                for (int m = 0; m < NRP; m += 2) {
                    const double2 v = ld2(gG + jl * LDA + m);
                    u[m] = v.x;
                    u[m + 1] = v.y;
                }
                y = gG[jl * LDA + NOP];
            } else {
#pragma unroll
                for (int m = 0; m < NRP; ++m) u[m] = 0.0;
            }
#pragma unroll
            for (int m = NRP - 1; m >= 0; --m) {
                if (m < nrmax) {
                    if (jl == m && m < nr) y = y / u[m];      // dgetrs divides by the pivot (exact where the data are)
                    const double xm = __shfl_sync(FULL, y, grp * LPF + m);
                    if (jl < m && m < nr) y = fma(-u[m], xm, y);
                }
            }
            // fi[r2o[j]] = x_j * col_j; known slots and columns beyond the model are not touched
            if (jl < nr) P.fi[cg * P.fi_s0 + gR2O[jl]] = y * gRS[jl];
            __syncwarp();
            continue;
        }

        // ================= whole warp: P5 operator columns (dgetrs per right-hand side) =================
        // One fit after the other (a runtime loop: the unrolled substitution code exists once, not FPW times).
#pragma unroll 1
        for (int f = 0; f < FPW; ++f) {
            if (c0 + f >= P.ncases) break;
            const long long cf = case_of(c0 + f);
            const CaseMeta mt = get_meta(cf);
            if (mt.nr < 1) continue;
            const int nrf = mt.nr, nkf = mt.nk, nqf = mt.nk + mt.nkn, nof = mt.no;
            const double* fb = fits + f * P.fit_doubles;
            const double* G = fb + oG;
            const double* DI = fb + oDINV;
            const double* RS = fb + oRS;
            const double2* REC = reinterpret_cast<const double2*>(fb + oREC);
            const int qblk = (nqf + 31) >> 5;
            for (int b = 0; b < qblk; ++b) {
                const int q = b * 32 + lane;
                double y[NRP];
                // y = P (row o (w_q c_q)): right-hand side row_j w_q c[q,oj] (impl.pyx:769-779), in pivot order.
                // The lane recomputes its column of monomials into its own CT column and reads it back through
                // the pivot-order row offsets (no other lane touches that column: no barrier needed).
                double* ctq = CT + lane;
                if (pending) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                    pending = false;
                }
                monomial_column(cf, q, nkf, nof, ctq);
                if (q >= nkf && q < nqf) {
                    const double* kn = fb + oKN + (q - nkf) * NOP;
                    for (int s2 = 0; s2 < nof; ++s2) ctq[s2 * CB] = kn[s2];
                }
                const double wq = q < nqf ? fb[oW + q] : 0.0;
#pragma unroll
                for (int i = 0; i < NRP; ++i) {
                    y[i] = 0.0;
                    if (i < nrf) {
                        const double2 rec = REC[i];
                        y[i] = rec.x * (wq * ctq[__double2loint(rec.y)]);
                    }
                }
                if constexpr (NRP > 16) {
                // (3D orders 3 and 4: 20 / 36 solution components per lane leave no registers for overlapped rows -- row by row)
                // unit-lower forward substitution
#pragma unroll
                for (int i = 1; i < NRP; ++i) {
                    if (i < nrf) {
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
                    if (i < nrf) {
#pragma unroll
                        for (int m = ((i + 1) & ~1); m < NRP; m += 2) {
                            const double2 u2 = ld2(G + i * LDA + m);
                            if (m > i) y[i] = fma(-u2.x, y[m], y[i]);
                            y[i] = fma(-u2.y, y[m + 1], y[i]);
                        }
                        y[i] *= DI[i];
                    }
                }
                } else {
                // Both substitutions in column (axpy) order, two columns per step, WITHOUT a branch per row: rows and
                // columns beyond the fit's size hold zeros (cleared after the LU), so the whole block is one stretch of
                // straight-line code in which the updates of different rows are independent and overlap -- with a
                // branch per row (one dependent FMA chain per basic block) the substitutions were latency-bound.
                // unit-lower forward substitution: y_i -= l_ip y_p, p ascending (dtrsm's order)
#pragma unroll
                for (int p = 0; p + 1 < NRP; p += 2) {
                    {
                        const double2 l2 = ld2(G + (p + 1) * LDA + p);
                        y[p + 1] = fma(-l2.x, y[p], y[p + 1]);
                    }
#pragma unroll
                    for (int i = p + 2; i < NRP; ++i) {
                        const double2 l2 = ld2(G + i * LDA + p);
                        y[i] = fma(-l2.x, y[p], y[i]);
                        y[i] = fma(-l2.y, y[p + 1], y[i]);
                    }
                }
                // upper backward substitution: x_m = y_m / u_mm, then y_i -= u_im x_m for i < m, m descending (dtrsm's order)
#pragma unroll
                for (int m = NRP - 2; m >= 0; m -= 2) {
                    y[m + 1] *= DI[m + 1];
                    {
                        const double2 u2 = ld2(G + m * LDA + m);      // (u_mm, u_m,m+1)
                        y[m] = fma(-u2.y, y[m + 1], y[m]);
                        y[m] *= DI[m];
                    }
#pragma unroll
                    for (int i = 0; i < m; ++i) {
                        const double2 u2 = ld2(G + i * LDA + m);
                        y[i] = fma(-u2.y, y[m + 1], y[i]);
                        y[i] = fma(-u2.x, y[m], y[i]);
                    }
                }
                }
                // Op[q][j] = x_j * col_j: 32 consecutive operator rows staged in CT, one bulk store per block
                const int rows = min(32, nqf - b * 32);
                __syncwarp();    // gathers done
                if (lane < rows) {
#pragma unroll
                    for (int i = 0; i < NRP; ++i)
                        if (i < nrf) CT[lane * nrf + i] = y[i] * RS[i];
                }
                if (lane == 0 && ((rows * nrf) & 1)) CT[rows * nrf] = 0.0;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_1d(P.op + mt.op_off + (long long)b * 32 * nrf, CT, (uint32_t)(((rows * nrf + 1) & ~1) * 8));
                    tma_store_commit();
                }
                pending = true;
            }
        }
    }
    if (lane == 0) tma_store_wait_all();
}

// ---- launch -----------------------------------------------------------------------------------------
template <int DIM, int ORD, bool DIRECT>
static cudaError_t launch_k(const PrepRegParams& P, int blocks, int threads, size_t smem, cudaStream_t st,
                            int* occupancy) {
    cudaError_t e = cudaFuncSetAttribute(prepare_reg_kernel<DIM, ORD, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    if (occupancy)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, prepare_reg_kernel<DIM, ORD, DIRECT>, threads, smem);
    prepare_reg_kernel<DIM, ORD, DIRECT><<<blocks, threads, smem, st>>>(P);
    return cudaGetLastError();
}

template <int DIM, int ORD>
static cudaError_t launch_t(const PrepRegParams& P, int blocks, int threads, size_t smem, cudaStream_t st,
                            int* occupancy) {
    if (P.fk) return launch_k<DIM, ORD, true>(P, blocks, threads, smem, st, occupancy);
    return launch_k<DIM, ORD, false>(P, blocks, threads, smem, st, occupancy);
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

cudaError_t prepare_reg_occupancy(int dim, int maxorder, int threads, size_t smem, int* ctas_per_sm, bool direct) {
    PrepRegParams P{};
    if (direct) P.fk = reinterpret_cast<const double*>(8);    // selects the one-shot instantiation; nothing is launched
    return dispatch(dim, maxorder, P, 0, threads, smem, nullptr, ctas_per_sm);
}

bool prepare_reg_direct_ok(int dim, int maxorder) {
    const int no = prep_no(dim, maxorder);
    return ((no + 7) & ~7) > no;      // a free padding slot carries the data through the Gram phase (true for every model)
}

cudaError_t launch_prepare_reg(int dim, int maxorder, const PrepRegParams& P, int blocks, int threads, size_t smem,
                               cudaStream_t st) {
    return dispatch(dim, maxorder, P, blocks, threads, smem, st, nullptr);
}

}  // namespace wlsqm
