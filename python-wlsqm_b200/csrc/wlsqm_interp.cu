// wlsqm_interp.cu -- K3: evaluate fitted local models (or any of their derivatives) at query points.
//
// Replaces
//   interpolate_{1,2,3}D / interpolate_nD  wlsqm/fitter/interp.pyx:860-937 / 670-846 / 274-655 / 252-258
//   taylor_* / general_*                   wlsqm/fitter/polyeval.pyx:82-1000
//   expert_interpolate_nearest             wlsqm/fitter/expert.pyx:874-895   (the prange over queries)
//   interpolate_fit                        wlsqm/fitter/interp.pyx:34-143    (single model, I == nullptr)
//
// The reference hard-codes, per (dimension, derivative), a table that shifts coefficients into a
// scratch polynomial `fi2` and evaluates that with a nested Horner form.  Here the derivative
// d^(p,q,r) of  sum_s fi[s] x^a y^b z^c/(a! b! c!)  is evaluated directly as
//   sum_{a>=p, b>=q, c>=r} fi[s] x^(a-p) y^(b-q) z^(c-r) / ((a-p)! (b-q)! (c-r)!)
// by shifting the per-axis scaled-power tables; slots whose exponents fall below (p,q,r) multiply
// a zero.  A derivative slot >= no gives 0, as interp.pyx:690-694.
// One thread per query; queries that share a model hit the same coefficient row in L1/L2.
#include <type_traits>
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"

namespace wlsqm {

__device__ __forceinline__ Pow5 shifted_powers(double d, int p) {
    const Pow5 b = scaled_powers(d);
    Pow5 r;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        double v = 0.0;
#pragma unroll
        for (int pp = 0; pp <= a; ++pp)
            if (p == pp) v = b.p[a - pp];
        r.p[a] = v;
    }
    return r;
}

template <int DIM>
__device__ __forceinline__ double eval_diff(int no, const double* __restrict__ fi, double dx, double dy, double dz,
                                            int p, int q, int r) {
    const Pow5 px = shifted_powers(dx, p);
    const Pow5 py = shifted_powers(DIM >= 2 ? dy : 0.0, DIM >= 2 ? q : 0);
    const Pow5 pz = shifted_powers(DIM >= 3 ? dz : 0.0, DIM >= 3 ? r : 0);
    double acc = 0.0;
    static_for<0, max_no<DIM>()>([&](auto I) {
        constexpr int S = max_no<DIM>() - 1 - decltype(I)::value;
        if (S < no) acc = fma(__ldg(fi + S), monomial<DIM, S, false>(px, py, pz), acc);
    });
    return acc;
}

template <int DIM>
__device__ __forceinline__ void slot_exponents(int s, int& p, int& q, int& r) {
    p = q = r = 0;
    static_for<0, max_no<DIM>()>([&](auto I) {
        constexpr int S = decltype(I)::value;
        constexpr SlotExp e = slot_exp<DIM>(S);
        if (s == S) { p = e.a; q = e.b; r = e.c; }
    });
}

template <int DIM>
__global__ void __launch_bounds__(256) interpolate_kernel(InterpParams P) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.nx) return;
    const long long i = P.I ? P.I[m] : 0;
    const int order = P.order ? (int)P.order[i] : P.order_uniform;
    const int no = number_of_dofs(DIM, order);
    const double* xq = P.x + m * P.x_s0;
    const double* xo = P.xi + i * P.xi_s0;
    const double dx = xq[0] - xo[0];
    const double dy = DIM >= 2 ? xq[DIM >= 2 ? 1 : 0] - xo[DIM >= 2 ? 1 : 0] : 0.0;
    const double dz = DIM >= 3 ? xq[DIM >= 3 ? 2 : 0] - xo[DIM >= 3 ? 2 : 0] : 0.0;
    const double* fi = P.fi + i * P.fi_s0;
    if (P.diff >= 0) {
        int p, q, r;
        slot_exponents<DIM>(P.diff, p, q, r);
        st_stream(P.out + m, eval_diff<DIM>(no, fi, dx, dy, dz, p, q, r));
    } else {   // extension: every derivative slot of the model in one pass, out[m][0..no)
        double* o = P.out + m * P.out_s0;
        for (int d = 0; d < no; ++d) {
            int p, q, r;
            slot_exponents<DIM>(d, p, q, r);
            st_stream(o + d, eval_diff<DIM>(no, fi, dx, dy, dz, p, q, r));
        }
    }
}

cudaError_t launch_interpolate(const InterpParams& P, cudaStream_t st) {
    if (P.nx == 0) return cudaSuccess;
    const int threads = 256;
    const unsigned blocks = (unsigned)((P.nx + threads - 1) / threads);
    if (P.dim == 1) interpolate_kernel<1><<<blocks, threads, 0, st>>>(P);
    else if (P.dim == 2) interpolate_kernel<2><<<blocks, threads, 0, st>>>(P);
    else interpolate_kernel<3><<<blocks, threads, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace wlsqm
