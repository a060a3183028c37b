// wlsqm_solve.cu -- K2: the per-time-step solve as ONE streaming pass over the stored operators.
//
// Replaces ExpertSolver.solve's per-case work (wlsqm/fitter/expert.pyx:467-655):
//   Case_set_fi / Case_get_fi   wlsqm/fitter/infra.pyx:780-795
//   impl.solve / solve_contig   wlsqm/fitter/impl.pyx:731-846 / 861-974   (RHS, knowns elimination, dgetrs, sens)
//   impl.solve_iterative        wlsqm/fitter/impl.pyx:986-1083             (data-space refinement)
//
// Layout / dataflow.  prepare() left, for every case, one contiguous block Op[(nk+nkn)][nr] in HBM
// (wlsqm_prepare.cu).  Each warp owns a ring of shared-memory stages; one elected lane streams the
// next cases' blocks into the ring with 1D bulk TMA (cp.async.bulk + mbarrier complete_tx) while
// the warp works on the current one, so HBM sees long, fully used, 16 B-aligned bursts and no
// register is spent on loads in flight.  With the block resident in shared memory:
//   fi[r2o[j]]  = sum_q Op[q][j] * fext[q],   fext = (fk[0..nk), known fi values)
//   sens[k][o]  = Op[k][j(o)]  (NaN for known o)                      -- no extra solves
//   ALGO_ITERATIVE: r_k = fk_k - model(xk_k); stop if max|r| == previous max|r| exactly
//                   (impl.pyx:1057-1060); fi[unknown] += Op[:nk]^T r    -- operator is not re-read from HBM
// Results go to the solver-owned fi copy (the reference's Case.fi, read later by interpolate) and,
// when the caller's fi is device memory that cannot alias fk, straight to it as well; otherwise the
// deferred write-back of expert.pyx:548-557 is a second tiny kernel (scatter_fi_kernel).
#include <type_traits>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

template <int DIM, bool ITER, bool SENS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS) solve_kernel(SolveParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int S = P.stages;
    double* wb = reinterpret_cast<double*>(smem_raw) + (size_t)warp * P.warp_doubles;
    double* ring = wb;                      // S * stage_doubles
    double* fext = wb + P.off_f;            // nk + nkn data values
    double* fis = wb + P.off_fi;            // current solution of the case (no values)
    double* rs = wb + P.off_r;              // ITER: residual at the neighbours
    double* xks = wb + P.off_xk;            // ITER: neighbour coordinates
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.bar_off_bytes) + warp * S;

    const long long gw = (long long)blockIdx.x * nwarps + warp;
    const long long GW = (long long)gridDim.x * nwarps;
    const long long n_my = gw < P.ncases ? (P.ncases - gw + GW - 1) / GW : 0;

    auto get_meta = [&](long long c) {
        CaseMeta m;
        if (P.meta) {
            m = P.meta[c];
        } else {
            m = P.uni;
            m.op_off = c * P.op_stride;
        }
        return m;
    };
    auto issue = [&](int s, long long c) {   // lane 0 only
        const CaseMeta m = get_meta(c);
        const uint32_t bytes = ((uint32_t)((m.nk + m.nkn) * (int)m.nr) * 8u + 15u) & ~15u;
        mbar_expect_tx(&bars[s], bytes);
        if (bytes) tma_load_1d(ring + (size_t)s * P.stage_doubles, P.op + m.op_off, bytes, &bars[s]);
    };

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
        for (int s = 0; s < S - 1 && s < n_my; ++s) issue(s, gw + (long long)s * GW);
    }
    __syncwarp();

    int stage = 0, itmax = 0;
    uint32_t phase = 0;
    for (long long i = 0; i < n_my; ++i) {
        const long long c = gw + i * GW;
        if (lane == 0 && i + S - 1 < n_my) {
            int sn = stage + S - 1;
            if (sn >= S) sn -= S;
            issue(sn, gw + (i + S - 1) * GW);
        }
        const CaseMeta mt = get_meta(c);
        const int nk = mt.nk, no = mt.no, nr = mt.nr, nkn = mt.nkn, nq = nk + nkn;
        const long long knowns = mt.knowns;

        // ---- gather the data of this case while its operator block is in flight ---------------
        {
            const double* f = P.fk + c * P.fk_s0;
            for (int k = lane; k < nk; k += 32) fext[k] = ld_stream(f + (long long)k * P.fk_s1);
        }
        for (int o = lane; o < no; o += 32) {
            if ((knowns >> o) & 1LL) {
                const double g = P.fi_in[c * P.fi_in_s0 + o];
                fext[nk + __popcll(knowns & ((1LL << o) - 1))] = g;
                fis[o] = g;
            }
        }
        if (ITER) {
            const double* xp = P.xk + c * P.xk_s0;
            if (P.xk_s1 == DIM) {
                for (int t = lane; t < nk * DIM; t += 32) xks[t] = xp[t];
            } else {
                for (int t = lane; t < nk * DIM; t += 32) xks[t] = xp[(long long)(t / DIM) * P.xk_s1 + (t % DIM)];
            }
        }
        __syncwarp();
        mbar_wait(&bars[stage], phase);
        const double* op = ring + (size_t)stage * P.stage_doubles;

        // ---- fi[unknown] = Op^T fext (lane = DOF slot) ------------------------------------------
        for (int o = lane; o < no; o += 32) {
            if (!((knowns >> o) & 1LL)) {
                const int j = o - __popcll(knowns & ((1LL << o) - 1));
                const double* col = op + j;
                double a0 = 0.0, a1 = 0.0;
                int q = 0;
                for (; q + 1 < nq; q += 2) {
                    const double2 f2 = *reinterpret_cast<const double2*>(fext + q);
                    a0 = fma(col[q * nr], f2.x, a0);
                    a1 = fma(col[(q + 1) * nr], f2.y, a1);
                }
                if (q < nq) a0 = fma(col[q * nr], fext[q], a0);
                fis[o] = a0 + a1;
            }
        }
        __syncwarp();

        // ---- sensitivities: the operator itself, re-indexed by DOF slot (impl.pyx:838-846) ------
        if (SENS) {
            double* sn = P.sens + c * P.sens_s0;
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            for (int t = lane; t < nk * no; t += 32) {
                const int k = t / no, o = t - k * no;
                double v = qnan;
                if (!((knowns >> o) & 1LL)) v = op[k * nr + (o - __popcll(knowns & ((1LL << o) - 1)))];
                st_stream(sn + (long long)k * P.sens_s1 + o, v);
            }
        }

        // ---- ALGO_ITERATIVE: refinement against the data (impl.pyx:1010-1083) --------------------
        int it = 0;
        if (ITER) {
            double xi0 = P.xi[c * P.xi_s0], xi1 = 0.0, xi2 = 0.0;
            if (DIM >= 2) xi1 = P.xi[c * P.xi_s0 + 1];
            if (DIM >= 3) xi2 = P.xi[c * P.xi_s0 + 2];
            double prev = -1.0;
            bool broke = false;
            for (it = 0; it < P.max_iter; ++it) {
                double nrm = 0.0;
                for (int k = lane; k < nk; k += 32) {
                    const double dx = xks[k * DIM] - xi0;
                    const double dy = DIM >= 2 ? xks[k * DIM + (DIM >= 2 ? 1 : 0)] - xi1 : 0.0;
                    const double dz = DIM >= 3 ? xks[k * DIM + (DIM >= 3 ? 2 : 0)] - xi2 : 0.0;
                    const double r = fext[k] - eval_taylor<DIM>(no, fis, dx, dy, dz);
                    rs[k] = r;
                    nrm = fmax(nrm, fabs(r));
                }
                nrm = warp_max(nrm);
                if (nrm == prev) { broke = true; break; }
                prev = nrm;
                __syncwarp();
                if (nr > 0) {
                    for (int o = lane; o < no; o += 32) {
                        if (!((knowns >> o) & 1LL)) {
                            const double* col = op + (o - __popcll(knowns & ((1LL << o) - 1)));
                            double a0 = 0.0, a1 = 0.0;
                            int q = 0;
                            for (; q + 1 < nk; q += 2) {
                                a0 = fma(col[q * nr], rs[q], a0);
                                a1 = fma(col[(q + 1) * nr], rs[q + 1], a1);
                            }
                            if (q < nk) a0 = fma(col[q * nr], rs[q], a0);
                            fis[o] += a0 + a1;
                        }
                    }
                }
                __syncwarp();
            }
            // for/else quirk of impl.pyx:1080-1083: a loop that runs to completion reports max_iter,
            // except max_iter <= 0, which reports 1.
            if (!broke) it = P.max_iter > 0 ? P.max_iter : 1;
            if (P.iters_case && lane == 0) P.iters_case[c] = it;
            itmax = max(itmax, it);
        }

        // ---- write-back: solver-owned copy (all `no` entries) and, if allowed, the caller's fi ---
        for (int o = lane; o < no; o += 32) {
            const double v = fis[o];
            P.fi_case[c * P.fi_case_ld + o] = v;
            if (P.fi_out && !((knowns >> o) & 1LL)) P.fi_out[c * P.fi_out_s0 + o] = v;
        }
        __syncwarp();   // every lane is done with this stage before lane 0 re-arms it
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (ITER && lane == 0 && itmax > 0) atomicMax(P.iters_max, itmax);
}

// Deferred write-back (expert.pyx:548-557): caller's fi <- solver-owned copy, first no_j columns.
__global__ void scatter_fi_kernel(const CaseMeta* meta, CaseMeta uni, long long ncases, const double* fi_case,
                                  int ld, double* fi_out, long long s0) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c = t / ld;
    const int o = (int)(t - c * ld);
    if (c >= ncases) return;
    const int no = meta ? meta[c].no : uni.no;
    if (o < no) fi_out[c * s0 + o] = fi_case[c * ld + o];
}

template <int DIM, bool ITER, bool SENS>
static cudaError_t launch_one(const SolveParams& P, int blocks, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e =
        cudaFuncSetAttribute(solve_kernel<DIM, ITER, SENS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    solve_kernel<DIM, ITER, SENS><<<blocks, threads, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_solve(int dim, const SolveParams& P, int blocks, int threads, size_t smem, cudaStream_t st) {
    const bool iter = P.algorithm == WLSQM_ALGO_ITERATIVE, sens = P.sens != nullptr;
    if (!iter) return sens ? launch_one<1, false, true>(P, blocks, threads, smem, st)
                           : launch_one<1, false, false>(P, blocks, threads, smem, st);
    if (dim == 1) return sens ? launch_one<1, true, true>(P, blocks, threads, smem, st)
                              : launch_one<1, true, false>(P, blocks, threads, smem, st);
    if (dim == 2) return sens ? launch_one<2, true, true>(P, blocks, threads, smem, st)
                              : launch_one<2, true, false>(P, blocks, threads, smem, st);
    return sens ? launch_one<3, true, true>(P, blocks, threads, smem, st)
                : launch_one<3, true, false>(P, blocks, threads, smem, st);
}

cudaError_t launch_scatter_fi(const CaseMeta* meta, const CaseMeta& uni, long long ncases, const double* fi_case,
                              int fi_case_ld, double* fi_out, long long fi_out_s0, cudaStream_t st) {
    const long long total = ncases * fi_case_ld;
    if (total == 0) return cudaSuccess;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    scatter_fi_kernel<<<(unsigned)blocks, threads, 0, st>>>(meta, uni, ncases, fi_case, fi_case_ld, fi_out, fi_out_s0);
    return cudaGetLastError();
}

}  // namespace wlsqm
