// wlsqm_kernels.h -- parameter blocks and launchers shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include "wlsqm_common.cuh"

namespace wlsqm {

constexpr int PREP_MAX_THREADS = 512;        // shared-memory variant (wlsqm_prepare_smem.cu)
constexpr int PREP_REG_SMALL_THREADS = 1024; // ... models of at most 8 DOFs: 64 registers, as many warps as shared memory holds
constexpr int PREP_REG_MAX_THREADS = 640;    // largest launch bound among the instantiations (2D order 4: 20 warps at 96 registers)
constexpr int PREP_REG_THREADS = 512;        // register/DMMA variant (wlsqm_prepare.cu): launch bound; the host picks the CTA size
constexpr int WLSQM_MAX_PEERS = 8;           // GPUs of one NVSwitch domain that can receive the fused gather
constexpr int PREP_CB = 36;                  // row stride (doubles) inside a 32-column block of the monomial table
constexpr int SOLVE_MAX_THREADS = 1024;       // ALGO_BASIC variants (<= 64 registers per thread)
constexpr int SOLVE_MAX_THREADS_ITER = 512;   // ALGO_ITERATIVE variants carry the Taylor evaluator
constexpr int SOLVE_ITER12_THREADS = 768;     // ... in 1D / 2D: one 24-warp CTA per SM (80 registers)

// All strides are in elements (doubles), all offsets into per-warp shared memory in doubles.
struct PrepareParams {
    const CaseMeta* meta;   // per-case records, or nullptr when the batch is uniform
    CaseMeta uni;           // the uniform record (op_off ignored)
    long long op_stride;    // uniform batches: doubles per operator block
    long long ncases;
    const double* xi; long long xi_s0;               // [ncases][dim]
    const double* xk; long long xk_s0, xk_s1;        // [ncases][nk][dim], last axis contiguous
    double* op;                                      // operator blocks (output)
    double* As; int as_stride;                       // debug: scaled matrices [ncases][as_stride] or nullptr
    int warp_doubles, off_w, off_a, off_rs, off_s, off_i;
};

// register/DMMA prepare kernel (wlsqm_prepare.cu)
struct PrepRegParams {
    const CaseMeta* meta;   // per-case records, or nullptr when the batch is uniform
    CaseMeta uni;
    long long op_stride;
    long long ncases;
    const double* xi; long long xi_s0;               // [ncases][dim]
    const double* xk; long long xk_s0, xk_s1;        // [ncases][nk][dim], last axis contiguous
    double* op;                                      // operator blocks (output)
    double* As; int as_stride;                       // debug: scaled matrices [ncases][as_stride] or nullptr
    int geom_uniform;                                // every case has the sizes of `uni`; only meta[c].knowns is read
    int nb;                                          // 32-column blocks of the monomial table
    int fit_doubles;                                 // shared-memory doubles per fit region (prep_reg_fit_doubles)
    int warp_doubles;                                // per warp: monomial table + prep_reg_fits_per_warp() fit regions
    // one-shot fit (no operator): data and in/out solution; fk == nullptr selects the operator-building kernel
    const double* fk; long long fk_s0, fk_s1;        // [ncases][nk]
    double* fi; long long fi_s0;                     // [ncases][>= no]: knowns read, unknowns written
    int phase_sync;                                  // CTA barrier at every phase boundary (instruction-cache sharing)
    const int* perm;                                 // launch over a case list: entry v of [0, ncases) is case perm[v] (nullptr: v)
};

struct SolveParams {
    const CaseMeta* meta;
    CaseMeta uni;
    int geom_uniform;                                // every case has the sizes of `uni`; only meta[c].knowns is read
    long long op_stride;
    long long ncases;                                // cases [case_lo, ncases) are processed by this launch
    long long case_lo;
    const double* op;
    const double* fk; long long fk_s0, fk_s1;        // [ncases][nk] fully strided (simple.pyx:149-159)
    const double* fi_in; long long fi_in_s0;         // caller's fi (known values are read from it)
    double* fi_case; int fi_case_ld;                 // solver-owned copy of the solution (Case.fi)
    double* fi_out; long long fi_out_s0;             // caller's fi on the device, or nullptr (deferred write-back)
    double* sens; long long sens_s0, sens_s1;        // [ncases][nk][no], last axis contiguous, or nullptr
    int algorithm, max_iter;
    const double* xi; long long xi_s0;               // ALGO_ITERATIVE: geometry kept at prepare()
    const double* xk; long long xk_s0, xk_s1;
    int* iters_max;                                  // max refinement iterations over all cases
    int* iters_case;                                 // optional per-case iteration counts
    int stages, stage_doubles;                       // TMA ring per warp; a stage is [operator | fext | xk]
    int off_f, off_xk;                               // offsets of fext / xk inside a stage (doubles)
    int warp_doubles, off_fi, off_r;                 // per-warp smem carve-up (doubles)
    int bar_off_bytes;                               // start of the mbarrier array
    int f_tma, xk_tma;                               // fk / xk rows qualify for bulk copies (alignment, unit stride)
    int pack_lw;                                     // packed variant: log2 of the lanes per case
    // fused result gather (multi-GPU): every solved row is also stored into the global solution array of
    // each of `ngather` GPUs (this one included) through peer memory over NVLink -- the all-gather of fi without a
    // collective: row c of this solver is row gather_row0 + c there
    int ngather;
    double* gather[WLSQM_MAX_PEERS];
    long long gather_row0, gather_s0;
    int gather_full;                                 // rows are whole 128 B lines (gather_s0 a multiple of 16, <= 32): the padding is stored too
};

struct InterpParams {
    int dim;
    long long nx;
    const double* x; long long x_s0;                 // [nx][dim]
    const long long* I;                              // [nx] model index, or nullptr (single model 0)
    long long nmodels;                               // indices outside [0, nmodels) give NaN (0: not checked)
    const double* xi; long long xi_s0;               // model origins [nmodels][dim]
    const double* fi; long long fi_s0;               // model coefficients [nmodels][fi_s0]
    const signed char* order; int order_uniform;     // per-model order, or uniform
    int diff;                                        // DOF slot to differentiate to, or -1: all slots
    double* out; long long out_s0;                   // [nx] (diff >= 0) or [nx][out_s0]
    int stage_no;                                    // all slots, every model has stage_no DOFs and out_s0 == stage_no:
                                                     // results are transposed through shared memory (else 0)
};

// thread-local message behind wlsqm_last_error() (wlsqm_capi.cu)
void set_last_error(const char* msg);

cudaError_t launch_prepare_smem(int dim, const PrepareParams& P, int blocks, int threads, size_t smem, cudaStream_t st);
int prep_reg_fit_doubles(int dim, int maxorder, int nb, int nkn_max);
int prep_reg_warp_doubles(int dim, int maxorder, int nb, int nkn_max);
int prep_reg_fits_per_warp(int dim, int maxorder);
int prep_reg_max_threads(int dim, int order);
cudaError_t prepare_reg_occupancy(int dim, int maxorder, int threads, size_t smem, int* ctas_per_sm, bool direct = false);
bool prepare_reg_direct_ok(int dim, int maxorder);
cudaError_t launch_prepare_reg(int dim, int maxorder, const PrepRegParams& P, int blocks, int threads, size_t smem,
                               cudaStream_t st);
cudaError_t launch_solve(int dim, const SolveParams& P, int blocks, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_solve_pack(const SolveParams& P, int blocks, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_scatter_fi(const CaseMeta* meta, const CaseMeta& uni, long long ncases, const double* fi_case,
                              int fi_case_ld, double* fi_out, long long fi_out_s0, cudaStream_t st);
cudaError_t launch_interpolate(const InterpParams& P, cudaStream_t st);
struct GridView;
cudaError_t launch_interpolate_continuous(const InterpParams& P, const GridView& g, double r, cudaStream_t st);
cudaError_t launch_getrf(int n, long long nlhs, double* A, int* ipiv, cudaStream_t st);
cudaError_t launch_getrs(int n, long long nlhs, const double* LU, const int* ipiv, double* b, cudaStream_t st);
cudaError_t launch_gesv(int n, long long nlhs, double* A, int* ipiv, double* b, cudaStream_t st);
cudaError_t launch_sy(int n, long long nlhs, double* A, int* ipiv, double* b, int do_factor, cudaStream_t st);
cudaError_t launch_symmetrize(int n, long long nlhs, double* A, cudaStream_t st);
cudaError_t launch_gtsv(int n, double* dl, double* d, double* du, double* b, cudaStream_t st);
cudaError_t launch_rescale(int nrows, int ncols, long long nlhs, double* A, int algo, double* row_scale, double* col_scale,
                           int* ok, cudaStream_t st);
cudaError_t launch_cond(int n_max, long long ncases, const CaseMeta* meta, const CaseMeta& uni, const double* As,
                        int as_stride, double* cond, cudaStream_t st);

}  // namespace wlsqm
