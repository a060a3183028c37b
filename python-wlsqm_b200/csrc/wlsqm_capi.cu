// wlsqm_capi.cu -- the C ABI declared in include/wlsqm_b200.h: solver handle, device state, staging of
// host arrays, launch configuration.  No numerics live here; see wlsqm_prepare.cu / wlsqm_solve.cu /
// wlsqm_interp.cu / wlsqm_lapack.cu.  There is deliberately no CPU code path: every compute entry
// point needs a CUDA device and fails with WLSQM_E_CUDA otherwise.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/wlsqm_b200.h"
#include "wlsqm_common.cuh"
#include "wlsqm_kernels.h"
#include "wlsqm_grid.h"
#include "wlsqm_mem.h"
#include "wlsqm_host.h"

namespace wlsqm {
GridView grid_view(const wlsqm_grid* g);
}

using namespace wlsqm;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

}  // namespace
namespace wlsqm {
void set_last_error(const char* msg) { g_err = msg ? msg : ""; }
}  // namespace wlsqm
namespace {

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? WLSQM_E_MEMORY : WLSQM_E_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(e__));                                                  \
    } while (0)

constexpr size_t SMEM_PER_SM = 228 * 1024;      // B200: 228 KB per SM, 227 KB per CTA, 1 KB reserved per CTA
constexpr size_t SMEM_PER_CTA = 227 * 1024;

inline int even(int v) { return (v + 1) & ~1; }

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// One look at the four metadata arrays of a batch (20 bytes per case): which of them hold one value throughout, and --
// only for those that do not -- the reductions the callers need.  The pass is bound by host memory bandwidth (about
// 1 ms per million cases on one thread), so large batches are split over the library's host threads.
struct MetaScan {
    bool same_nk = true, same_order = true, same_knowns = true, same_wm = true;
    int32_t max_nk = 0, min_order = 0, max_order = 0;
    bool uniform() const { return same_nk && same_order && same_knowns && same_wm; }
};
MetaScan scan_meta(long long n, const int32_t* nk, const int32_t* order, const int64_t* knowns, const int32_t* wm) {
    MetaScan r;
    if (n < 1) return r;
    r.max_nk = nk[0]; r.min_order = r.max_order = order[0];
    if (n < 2) return r;
    const int32_t nk0 = nk[0], od0 = order[0], wm0 = wm[0];
    const int64_t kn0 = knowns[0];
    std::atomic<uint32_t> d_nk{0}, d_od{0}, d_wm{0};
    std::atomic<uint64_t> d_kn{0};
    auto piece = [&](size_t lo, size_t hi) {
        uint32_t a = 0, b = 0, c = 0;
        uint64_t d = 0;
        for (size_t i = lo; i < hi; ++i) a |= (uint32_t)(nk[i] ^ nk0);
        for (size_t i = lo; i < hi; ++i) b |= (uint32_t)(order[i] ^ od0);
        for (size_t i = lo; i < hi; ++i) c |= (uint32_t)(wm[i] ^ wm0);
        for (size_t i = lo; i < hi; ++i) d |= (uint64_t)(knowns[i] ^ kn0);
        if (a) d_nk.fetch_or(a, std::memory_order_relaxed);
        if (b) d_od.fetch_or(b, std::memory_order_relaxed);
        if (c) d_wm.fetch_or(c, std::memory_order_relaxed);
        if (d) d_kn.fetch_or(d, std::memory_order_relaxed);
    };
    const bool par = n >= 262144 && env_int("WLSQM_META_THREADS", 1) != 0;
    if (par) par_for((size_t)n, piece);
    else piece(0, (size_t)n);
    r.same_nk = d_nk.load() == 0; r.same_order = d_od.load() == 0; r.same_wm = d_wm.load() == 0; r.same_knowns = d_kn.load() == 0;
    if (!r.same_nk || !r.same_order) {
        std::atomic<int32_t> mk{nk0}, lo_o{od0}, hi_o{od0};
        auto atomic_max = [](std::atomic<int32_t>& t, int32_t v) { int32_t c = t.load(); while (v > c && !t.compare_exchange_weak(c, v)) {} };
        auto atomic_min = [](std::atomic<int32_t>& t, int32_t v) { int32_t c = t.load(); while (v < c && !t.compare_exchange_weak(c, v)) {} };
        auto red = [&](size_t lo, size_t hi) {
            int32_t m = nk0, l = od0, h = od0;
            if (!r.same_nk)
                for (size_t i = lo; i < hi; ++i) m = nk[i] > m ? nk[i] : m;
            if (!r.same_order)
                for (size_t i = lo; i < hi; ++i) { l = order[i] < l ? order[i] : l; h = order[i] > h ? order[i] : h; }
            atomic_max(mk, m); atomic_min(lo_o, l); atomic_max(hi_o, h);
        };
        if (par) par_for((size_t)n, red);
        else red(0, (size_t)n);
        r.max_nk = mk.load(); r.min_order = lo_o.load(); r.max_order = hi_o.load();
    }
    return r;
}

// 1 = device-accessible (device or managed), 0 = host (pinned or pageable)
int is_device_ptr(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// A device buffer that only grows.  Blocks come from the library's memory pool (wlsqm_mem.h).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return WLSQM_OK;
        if (p) dev_free_sync(p);     // growing a live buffer: earlier work may still read the old block
        p = nullptr;
        cap = 0;
        if (bytes == 0) return WLSQM_OK;
        cudaError_t e = dev_alloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(WLSQM_E_MEMORY, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        cap = bytes;
        return WLSQM_OK;
    }
    // the caller has synchronised the streams that used the buffer
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        cap = 0;
    }
};

// dst[i][k][0..w) = src[idx[i][k]][0..w).  Only the first nk_i neighbours of case i are read (nk_i from the per-case
// records, else nk_uni; negative: all k columns): the padding of a ragged hood array (cKDTree and wlsqm_grid_knn both
// report missing neighbours as the index n) is never dereferenced and its slots are zero-filled.  An index outside
// [0, nsrc) among the USED slots -- where the reference's x[hoods] raises IndexError -- gives NaN and raises *err.
__global__ void gather_hoods_kernel(const double* __restrict__ src, long long src_s0, int w, const int32_t* __restrict__ idx,
                                    long long idx_s0, long long n, int k, double* __restrict__ dst, long long nsrc,
                                    const CaseMeta* __restrict__ meta, int nk_uni, int* __restrict__ err) {
    const long long per = (long long)k * w;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n * per;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / per;
        const int r = (int)(t - i * per);
        const int kk = r / w, d = r - kk * w;
        const int nki = meta ? meta[i].nk : (nk_uni >= 0 ? nk_uni : k);
        double v = 0.0;
        if (kk < nki) {
            const long long j = idx[i * idx_s0 + kk];
            if (j >= 0 && j < nsrc) v = src[j * src_s0 + d];
            else {
                v = __longlong_as_double(0x7ff8000000000000LL);
                if (err) *err = 1;
            }
        }
        dst[t] = v;
    }
}

// strided 3-level copy  dst[i][a][b] <- src[i*s0 + a*s1 + b*s2], dst dense
__global__ void gather3_kernel(double* __restrict__ dst, const double* __restrict__ src, long long n, int na, int nb,
                               long long s0, long long s1, long long s2) {
    const long long per = (long long)na * nb;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n * per;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / per;
        const int r = (int)(t - i * per);
        const int a = r / nb, b = r - a * nb;
        dst[t] = src[i * s0 + a * s1 + b * s2];
    }
}

}  // namespace

struct wlsqm_solver {
    int dim = 0, device = 0, algorithm = 1, do_sens = 0, max_iter = 0, debug = 0;
    long long ncases = 0;
    int maxnk = 0, maxno = 1, maxnr = 0, maxnq = 0, maxorder = 0, maxnkn = 0;
    bool uniform = true, any_knowns = false, uniform_no = true;
    bool last_nk_odd = false;   // nk of the LAST case is odd (see the bulk copies of fk rows in wlsqm_solver_solve)
    bool geom_uniform = true;   // same nk / order / weighting / number of knowns everywhere (the knowns pattern may vary)
    CaseMeta uni{};
    long long op_stride = 0, op_total = 0;
    std::vector<CaseMeta> hmeta;
    CaseMeta* dmeta = nullptr;
    signed char* dorder = nullptr;
    double* op = nullptr;
    double* fi_case = nullptr;
    // wlsqm_solver_keep_solution(s, 0): solve() with a device fi writes the caller's array only; interpolate() needs the copy
    bool keep_solution = true, fi_case_valid = true;
    double* xi_dev = nullptr;
    double* As = nullptr;
    int as_stride = 0;
    int* iters_dev = nullptr;   // [0] = max, [1..ncases] per case (ITERATIVE only)
    bool ready = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // H2D / D2H streams of the staged (host-pointer) pipeline
    std::vector<cudaEvent_t> events;
    int sm_count = 148;
    DevBuf xk_keep;             // dense copy of xk (ALGO_ITERATIVE needs the geometry at solve time)
    DevBuf st_xk, st_fk, st_fi, st_sens, st_x, st_I, st_out;
    DevBuf hoods_dev, hood_x, hood_f, hood_fk;   // prepare_hoods / solve_hoods: neighbour lists and gathered data
    long long hood_points = 0;
    int* hood_err = nullptr;                     // device flag: a used hood index fell outside [0, npoints)
    // fused result gather: global solution arrays of the GPUs of the job (peer memory), see wlsqm_solver_set_gather
    int ngather = 0;
    double* gather[WLSQM_MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    long long gather_row0 = 0, gather_s0 = 0;
    wlsqm_grid* models_grid = nullptr;           // search grid over the model origins (index_models)
    wlsqm_solver* lender = nullptr;              // guest mode: op / dmeta / dorder / xi_dev / As / xk_keep belong to this solver
    long long bytes_state = 0;
    // batches of mixed orders: case lists per order (device, ascending case index inside an order), so that prepare()
    // runs one kernel instantiation per order instead of the maximum order's for every case
    int* order_perm = nullptr;
    long long order_lo[6] = {0, 0, 0, 0, 0, 0};
};

namespace {

long long staging_bytes(const wlsqm_solver* s) {
    return (long long)(s->xk_keep.cap + s->st_xk.cap + s->st_fk.cap + s->st_fi.cap + s->st_sens.cap + s->st_x.cap +
                       s->st_I.cap + s->st_out.cap + s->hoods_dev.cap + s->hood_x.cap + s->hood_f.cap + s->hood_fk.cap);
}

int use_device(const wlsqm_solver* s) {
    CU(cudaSetDevice(s->device));
    return WLSQM_OK;
}

// ---- launch configuration ------------------------------------------------------------------------
struct LaunchCfg { int blocks, threads; size_t smem; };

int config_prepare(const wlsqm_solver* s, PrepareParams& P, LaunchCfg& L) {
    const int nk = std::max(s->maxnk, 1), no = s->maxno, nr = std::max(s->maxnr, 1), nq = std::max(s->maxnq, 1);
    const int cs = no | 1, lda = nr | 1, sq = nq | 1;
    int off = 0;
    off += even(nk * cs);       P.off_w = off;
    off += even(nk);            P.off_a = off;
    off += even(nr * lda);      P.off_rs = off;
    off += even(nr);            P.off_s = off;
    off += even(nr * sq);       P.off_i = off;
    off += 36;
    P.warp_doubles = off;
    const size_t per_warp = (size_t)off * 8;
    if (per_warp > SMEM_PER_CTA)
        return fail(WLSQM_E_VALUE, "case too large for shared memory (nk=%d, no=%d): %zu bytes per fit", nk, no, per_warp);
    int warps = env_int("WLSQM_PREP_WARPS", 8);
    warps = std::max(1, std::min(warps, PREP_MAX_THREADS / 32));
    while (warps > 1 && warps * per_warp > SMEM_PER_CTA) --warps;
    L.threads = warps * 32;
    L.smem = warps * per_warp;
    int ctas = (int)(SMEM_PER_SM / (L.smem + 1024));
    ctas = std::max(1, std::min(ctas, std::max(1, 48 / warps)));
    long long need = (s->ncases + warps - 1) / warps;
    L.blocks = (int)std::max<long long>(1, std::min<long long>((long long)s->sm_count * ctas, need));
    return WLSQM_OK;
}

// register/DMMA kernel: false if the fit does not fit its shared-memory carve-up (then the smem kernel runs)
// (order_sel / count: the launch covers `count` cases of order `order_sel` only -- batches of mixed orders)
bool config_prepare_reg(const wlsqm_solver* s, PrepRegParams& P, LaunchCfg& L, bool direct = false, int order_sel = -1,
                        long long count = -1) {
    const int kord = order_sel >= 0 ? order_sel : s->maxorder;
    const long long ncases = count >= 0 ? count : s->ncases;
    const int nkp = (std::max(s->maxnk, 1) + 3) & ~3;
    P.nb = (std::max(nkp, s->maxnq) + 31) / 32;
    const int nkn_max = s->maxnq > 0 ? s->maxnkn : 0;
    P.fit_doubles = prep_reg_fit_doubles(s->dim, kord, P.nb, nkn_max);
    P.warp_doubles = prep_reg_warp_doubles(s->dim, kord, P.nb, nkn_max);
    const int fpw = prep_reg_fits_per_warp(s->dim, kord);
    const size_t per_warp = (size_t)P.warp_doubles * 8;
    if (per_warp > SMEM_PER_CTA) return false;
    // CTA size: the one that keeps most warps resident (shared memory and registers both limit it).  The answer depends
    // only on (kernel instantiation, bytes per warp): remembered, so that a small one-shot fit does not pay sixteen
    // occupancy queries (tens of microseconds) on every call.
    int best_w = 0, best_ctas = 0;
    const int force_w = env_int("WLSQM_PREP_WARPS", 0);
    struct OccKey { int dim, ord, direct, force; size_t per_warp; int w, ctas; };
    static std::mutex occ_mu;
    static std::vector<OccKey> occ_cache;
    {
        std::lock_guard<std::mutex> lock(occ_mu);
        for (const OccKey& k : occ_cache)
            if (k.dim == s->dim && k.ord == kord && k.direct == (int)direct && k.force == force_w && k.per_warp == per_warp) {
                best_w = k.w; best_ctas = k.ctas;
                break;
            }
    }
    const bool cached = best_w > 0;
    // (whole multiples of the four schedulers first -- 18 warps measured slower than 16 for 3D order 3 -- then any size)
    for (int pass = 0; !cached && pass < 2 && best_w < 1; ++pass)
    for (int w = 1; w <= prep_reg_max_threads(s->dim, kord) / 32; ++w) {
        if (force_w > 0 && w != force_w) continue;
        if (pass == 0 && force_w <= 0 && (w & 3)) continue;
        if (w * per_warp > SMEM_PER_CTA) break;
        int c = 0;
        if (prepare_reg_occupancy(s->dim, kord, w * 32, w * per_warp, &c, direct) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        // among the CTA sizes that keep the most warps resident, the LARGEST: the warps of one CTA start together and
        // stay roughly in phase, so they share instruction-cache lines of this kernel's ~75 KB of straight-line code
        // (measured, 2D order 4: one 16-warp CTA per SM 4.71 ms, four 4-warp CTAs 5.26 ms)
        if (w * c >= best_w * best_ctas) { best_w = w; best_ctas = c; }
    }
    if (best_w < 1 || best_ctas < 1) return false;
    if (!cached) {
        std::lock_guard<std::mutex> lock(occ_mu);
        occ_cache.push_back(OccKey{s->dim, kord, (int)direct, force_w, per_warp, best_w, best_ctas});
    }
    const int warps = best_w;
    int ctas = best_ctas;
    L.threads = warps * 32;
    L.smem = warps * per_warp;
    P.phase_sync = env_int("WLSQM_PREP_PHASE_SYNC", 0);
    const int cap = env_int("WLSQM_PREP_CTAS", 0);
    if (cap > 0) ctas = std::min(ctas, cap);
    long long need = (ncases + (long long)warps * fpw - 1) / ((long long)warps * fpw);
    L.blocks = (int)std::max<long long>(1, std::min<long long>((long long)s->sm_count * ctas, need));
    return true;
}

int config_solve(const wlsqm_solver* s, SolveParams& P, LaunchCfg& L, long long ncases_launch) {
    const bool iter = s->algorithm == WLSQM_ALGO_ITERATIVE;
    // one stage of the per-warp ring: [operator block | fext = fk + known fi | xk (ALGO_ITERATIVE)]
    const int op_doubles = std::max(2, even(s->maxnq * s->maxnr));
    const int f_doubles = std::max(2, even(s->maxnq));
    const int x_doubles = iter ? std::max(2, even(s->maxnk * s->dim)) : 0;
    const int stage_doubles = op_doubles + f_doubles + x_doubles;
    const size_t stage_bytes = (size_t)stage_doubles * 8;
    // Measured on B200 (profiles/README.md): a ring of 2 stages and ONE CTA per SM with as many warps as
    // fit (16 for the headline block size) streams best; deeper rings or several CTAs per SM put more
    // distant operator blocks in flight at once and lose HBM efficiency.  Blocks under 3 KB are bound by
    // per-case issue overhead instead and want two such CTAs per SM.
    // ALGO_ITERATIVE keeps a warp busy on one block for thousands of cycles: more resident warps (one stage
    // each) beat prefetch depth there (measured: 3D order 4 k=60 with sens, 16.7 -> 9.6 ms per 1M points).
    // Batches whose cases differ in size (ALGO_BASIC): when two stages of the LARGEST block leave room for fewer than
    // 16 warps, one stage and twice the warps win -- the typical block is smaller than the stage, so resident warps put
    // more bytes in flight than a prefetch slot sized for the worst case (measured, 400k cases 3D orders 2-4 nk 40-60:
    // 2 stages x 6 warps 1.21 ms, 1 stage x 12 warps 0.73 ms = 0.78 of the HBM peak)
    const bool hetero_shallow = !iter && !s->geom_uniform && 16 * (2 * stage_bytes + 512) > SMEM_PER_CTA;
    int S = env_int("WLSQM_SOLVE_STAGES", ((iter && stage_bytes >= 8192) || hetero_shallow) ? 1 : 2);
    S = std::max(1, std::min(S, 8));
    // ALGO_ITERATIVE in 1D / 2D: one 24-warp CTA per SM at 80 registers (the refinement loop is latency-bound: 24 resident
    // warps beat 16, and one CTA whose warps start together beats two 12-warp CTAs: 1.59 -> 1.54 ms)
    const bool iter_small = iter && s->dim < 3;
    const int max_warps = (iter ? (iter_small ? SOLVE_ITER12_THREADS : SOLVE_MAX_THREADS_ITER) : SOLVE_MAX_THREADS) / 32;
    // Batches whose cases differ in size: the ring is sized for the largest block but the typical block is smaller, so the
    // pass is bound by per-case issue overhead and bytes in flight like any small-block batch -- as many resident warps as
    // shared memory holds (measured, 1M cases 2D orders 2-4 nk 22-30: 16 warps 0.770 ms, 24 warps 0.589 ms)
    const size_t mean_block_bytes = s->ncases > 0 ? (size_t)(s->op_total * 8 / s->ncases) : stage_bytes;
    const bool small_mean = !iter && !s->geom_uniform && (mean_block_bytes < 3072 || hetero_shallow);
    int warps = env_int("WLSQM_SOLVE_WARPS", iter_small ? 24 : (small_mean ? 32 : 16));
    warps = std::max(1, std::min(warps, max_warps));
    size_t per_warp = 0;
    int off_fi, off_r, wd;
    for (;;) {
        off_fi = S * stage_doubles;
        off_r = off_fi + 36;
        wd = off_r + (iter ? std::max(2, even(s->maxnk)) : 0);
        wd = (wd + 15) & ~15;                      // 128 B granularity per warp slice
        per_warp = (size_t)wd * 8 + (size_t)S * 8;
        if (warps * per_warp <= SMEM_PER_CTA) break;
        if (warps > 4) --warps;
        else if (S > 2) --S;
        else if (warps > 1) --warps;
        else if (S > 1) --S;
        else return fail(WLSQM_E_VALUE, "operator block too large for shared memory (%zu bytes)", stage_bytes);
    }
    P.stages = S;
    P.stage_doubles = stage_doubles;
    P.off_f = op_doubles; P.off_xk = op_doubles + f_doubles;
    P.off_fi = off_fi; P.off_r = off_r;
    P.warp_doubles = wd;
    // bulk-copy eligibility of the data arrays (16 B alignment of every row start, unit stride)
    P.f_tma = (P.fk_s1 == 1 && ((uintptr_t)P.fk % 16 == 0) && (P.fk_s0 % 2 == 0)) ? 1 : 0;
    P.xk_tma = (iter && P.xk_s1 == s->dim && ((uintptr_t)P.xk % 16 == 0) && (P.xk_s0 % 2 == 0)) ? 1 : 0;
    if (env_int("WLSQM_SOLVE_NO_FTMA", 0)) P.f_tma = P.xk_tma = 0;
    P.bar_off_bytes = warps * wd * 8;
    L.threads = warps * 32;
    L.smem = (size_t)P.bar_off_bytes + (size_t)warps * S * 8;
    int ctas = (int)(SMEM_PER_SM / (L.smem + 1024));
    ctas = std::max(1, std::min(ctas, std::max(1, env_int("WLSQM_SOLVE_MAXWARPS_SM", iter_small ? 24 : ((stage_bytes >= 3072 && !small_mean) ? 16 : 32)) / warps)));
    long long need = (ncases_launch + warps - 1) / warps;
    L.blocks = (int)std::max<long long>(1, std::min<long long>((long long)s->sm_count * ctas, need));
    return WLSQM_OK;
}

// packed solve (several small cases per warp pass): launch configuration; the stage holds
// [CPW operator blocks | CPW fk rows | CPW x nkn known values]
bool pack_eligible(const wlsqm_solver* s) {
    // blocks of 3 KB and more already stream at the HBM roofline with one case per warp (measured)
    return s->geom_uniform && s->algorithm != WLSQM_ALGO_ITERATIVE && s->maxno <= 16 && s->uni.nr > 0 &&
           s->op_stride * 8 < env_int("WLSQM_SOLVE_PACK_BELOW", 3072) && !env_int("WLSQM_SOLVE_NOPACK", 0);
}
int config_solve_pack(const wlsqm_solver* s, SolveParams& P, LaunchCfg& L, long long ncases_launch) {
    const int no = s->uni.no, nk = s->uni.nk, nkn = s->uni.nkn;
    int lw = 0;
    while ((1 << lw) < no) ++lw;
    P.pack_lw = lw;
    const int cpw = 32 >> lw;
    const int op_doubles = cpw * (int)s->op_stride;
    const int f_doubles = std::max(2, even(cpw * nk));
    const int k_doubles = std::max(2, even(cpw * nkn));
    const int stage_doubles = op_doubles + f_doubles + k_doubles;
    const size_t stage_bytes = (size_t)stage_doubles * 8;
    int S = env_int("WLSQM_SOLVE_STAGES", 2);
    S = std::max(2, std::min(S, 8));
    int warps = env_int("WLSQM_SOLVE_WARPS", 32);    // one 32-warp CTA per SM (measured: 0.692 ms against 0.712 ms for 8-warp CTAs, cfg4 2D)
    warps = std::max(1, std::min(warps, SOLVE_MAX_THREADS / 32));
    int wd = 0;
    for (;;) {
        wd = (S * stage_doubles + 15) & ~15;
        if ((size_t)warps * (wd * 8 + S * 8) <= SMEM_PER_CTA) break;
        if (warps > 1) --warps;
        else return fail(WLSQM_E_VALUE, "operator pack too large for shared memory (%zu bytes)", stage_bytes);
    }
    P.stages = S;
    P.stage_doubles = stage_doubles;
    P.off_f = op_doubles; P.off_xk = op_doubles + f_doubles;
    P.warp_doubles = wd;
    // one bulk copy for the CPW fk rows of a pack: dense rows, 16 B aligned pack starts
    P.f_tma = (P.fk_s1 == 1 && P.fk_s0 == nk && ((uintptr_t)P.fk % 16 == 0) && ((cpw * nk) % 2 == 0)) ? 1 : 0;
    if (env_int("WLSQM_SOLVE_NO_FTMA", 0)) P.f_tma = 0;
    P.xk_tma = 0;
    P.bar_off_bytes = warps * wd * 8;
    L.threads = warps * 32;
    L.smem = (size_t)P.bar_off_bytes + (size_t)warps * S * 8;
    int ctas = (int)(SMEM_PER_SM / (L.smem + 1024));
    ctas = std::max(1, std::min(ctas, std::max(1, env_int("WLSQM_SOLVE_MAXWARPS_SM", 32) / warps)));
    const long long packs = (ncases_launch + cpw - 1) / cpw;
    long long need = (packs + warps - 1) / warps;
    L.blocks = (int)std::max<long long>(1, std::min<long long>((long long)s->sm_count * ctas, need));
    return WLSQM_OK;
}

// copy a pitched host/device 2-D array of doubles into a dense device buffer (rows x width)
int to_dense(double* dst, const double* src, long long rows, long long width, long long pitch, cudaStream_t st) {
    if (rows == 0 || width == 0) return WLSQM_OK;
    if (rows == 1 || pitch == width) {
        CU(cudaMemcpyAsync(dst, src, (size_t)rows * width * 8, cudaMemcpyDefault, st));
        return WLSQM_OK;
    }
    CU(cudaMemcpy2DAsync(dst, (size_t)width * 8, src, (size_t)pitch * 8, (size_t)width * 8, (size_t)rows,
                         cudaMemcpyDefault, st));
    return WLSQM_OK;
}
// below this size the driver's own staging of pageable copies is as fast (measured: 24 MB at 25 GB/s, 240 MB at 12 GB/s)
constexpr long long BOUNCE_MIN_BYTES = 64ll << 20;
// to_dense / from_dense for arrays that may be ordinary pageable host memory: large ones go through the page-locked
// rings with threaded host copies (wlsqm_host.h); the call returns when the host side is done with the array
int stage_in(int device, double* dst, const double* src, long long rows, long long width, long long pitch, cudaStream_t st) {
    if (rows * width * 8 >= BOUNCE_MIN_BYTES && (size_t)width * 8 <= BOUNCE_SLOT_BYTES && env_int("WLSQM_BOUNCE", 1) != 0 &&
        !is_device_ptr(src) && is_pageable_host(src)) {
        bounce_lock();
        const cudaError_t e = h2d_bounced(bounce_rings(device).in, dst, src, rows, width, pitch, st);
        bounce_unlock();
        if (e != cudaSuccess) return fail(WLSQM_E_CUDA, "host -> device staging: %s", cudaGetErrorString(e));
        return WLSQM_OK;
    }
    return to_dense(dst, src, rows, width, pitch, st);
}
int from_dense(double* dst, long long pitch, const double* src, long long src_pitch, long long rows, long long width,
               cudaStream_t st);
int stage_out(int device, double* dst, long long pitch, const double* src, long long src_pitch, long long rows, long long width,
              cudaStream_t st) {
    if (rows * width * 8 >= BOUNCE_MIN_BYTES && (size_t)width * 8 <= BOUNCE_SLOT_BYTES && env_int("WLSQM_BOUNCE", 1) != 0 &&
        !is_device_ptr(dst) && is_pageable_host(dst)) {
        bounce_lock();
        const cudaError_t e = d2h_bounced(bounce_rings(device).out, dst, pitch, src, src_pitch, rows, width, st);
        bounce_unlock();
        if (e != cudaSuccess) return fail(WLSQM_E_CUDA, "device -> host staging: %s", cudaGetErrorString(e));
        return WLSQM_OK;
    }
    return from_dense(dst, pitch, src, src_pitch, rows, width, st);
}
int from_dense(double* dst, long long pitch, const double* src, long long src_pitch, long long rows, long long width,
               cudaStream_t st) {
    if (rows == 0 || width == 0) return WLSQM_OK;
    if (rows == 1 || (pitch == width && src_pitch == width)) {
        CU(cudaMemcpyAsync(dst, src, (size_t)rows * width * 8, cudaMemcpyDefault, st));
        return WLSQM_OK;
    }
    CU(cudaMemcpy2DAsync(dst, (size_t)pitch * 8, src, (size_t)src_pitch * 8, (size_t)width * 8, (size_t)rows,
                         cudaMemcpyDefault, st));
    return WLSQM_OK;
}

bool ranges_overlap(const void* a, long long abytes, const void* b, long long bbytes) {
    const char* a0 = (const char*)a;
    const char* b0 = (const char*)b;
    return a0 < b0 + bbytes && b0 < a0 + abytes;
}

}  // namespace

extern "C" {

int wlsqm_b200_abi_version(void) { return 1; }
const char* wlsqm_last_error(void) { return g_err.c_str(); }

int wlsqm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int wlsqm_number_of_dofs(int dimension, int order) { return number_of_dofs(dimension, order); }

void* wlsqm_pinned_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
// write-combined page-locked memory: faster for the device to read over PCIe on some hosts, slow for the CPU to read back
// -- for buffers the host only WRITES (the per-step input fk)
void* wlsqm_pinned_alloc_wc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocWriteCombined) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void wlsqm_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

// max nk, min / max order and whether every case has the same (nk, order, knowns, weighting), in one vectorised pass
// (the Python mirror needs them for its shape checks; three numpy reductions over 1M cases cost ~1 ms per call)
int wlsqm_meta_summary(int64_t ncases, const int32_t* nk, const int32_t* order, const int64_t* knowns, const int32_t* wm,
                       int32_t* max_nk, int32_t* min_order, int32_t* max_order, int32_t* uniform) {
    if (ncases < 0 || (ncases > 0 && (!nk || !order || !knowns || !wm))) return fail(WLSQM_E_VALUE, "NULL metadata array");
    const MetaScan m = scan_meta(ncases, nk, order, knowns, wm);
    if (max_nk) *max_nk = m.max_nk;
    if (min_order) *min_order = m.min_order;
    if (max_order) *max_order = m.max_order;
    if (uniform) *uniform = m.uniform() ? 1 : 0;
    return WLSQM_OK;
}

int wlsqm_pool_stats(int device, int64_t* reserved, int64_t* used) {
    long long r = -1, u = -1;
    dev_pool_stats(device, &r, &u);
    if (reserved) *reserved = r;
    if (used) *used = u;
    return WLSQM_OK;
}

int wlsqm_pool_trim(int device) {
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    dev_pool_trim(device);
    return WLSQM_OK;
}

int wlsqm_solver_create(int dimension, int64_t ncases, const int32_t* nk, const int32_t* order, const int64_t* knowns,
                        const int32_t* wm, int algorithm, int do_sens, int max_iter, int debug, int device,
                        wlsqm_solver_t** out) {
    if (!out) return fail(WLSQM_E_VALUE, "out is NULL");
    *out = nullptr;
    if (dimension < 1 || dimension > 3) return fail(WLSQM_E_VALUE, "Dimension must be 1, 2 or 3, got %d", dimension);
    if (algorithm != WLSQM_ALGO_BASIC && algorithm != WLSQM_ALGO_ITERATIVE)
        return fail(WLSQM_E_VALUE, "Unknown algorithm specifier %d; see wlsqm.fitter.defs for valid specifiers ALGO_*",
                    algorithm);
    if (ncases < 0) return fail(WLSQM_E_VALUE, "ncases must be >= 0");
    if (ncases > 0 && (!nk || !order || !knowns || !wm)) return fail(WLSQM_E_VALUE, "NULL metadata array");
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= wlsqm_device_count()) return fail(WLSQM_E_VALUE, "bad device ordinal %d", device);

    wlsqm_solver* s = new (std::nothrow) wlsqm_solver();
    if (!s) return fail(WLSQM_E_MEMORY, "out of host memory");
    s->dim = dimension; s->device = device; s->algorithm = algorithm; s->do_sens = do_sens ? 1 : 0;
    s->max_iter = max_iter; s->debug = debug ? 1 : 0; s->ncases = ncases;
    // Uniform batches (every case with the same nk / order / knowns / weighting -- the usual ExpertSolver and
    // fit_*_many call) are recognised by one pass of comparisons and keep no per-case records at all.
    const bool all_same = scan_meta(ncases, nk, order, knowns, wm).uniform();
    const long long nrec = all_same ? std::min<long long>(ncases, 1) : ncases;
    try {
        s->hmeta.resize((size_t)nrec);
    } catch (...) {
        delete s;
        return fail(WLSQM_E_MEMORY, "out of host memory");
    }
    long long off = 0;
    for (long long i = 0; i < nrec; ++i) {
        const int no = number_of_dofs(dimension, order[i]);
        if (no < 0) { delete s; return fail(WLSQM_E_VALUE, "case %lld: order must be 0..4, got %d", i, order[i]); }
        if (nk[i] < 0) { delete s; return fail(WLSQM_E_VALUE, "case %lld: nk must be >= 0, got %d", i, nk[i]); }
        if (wm[i] != WLSQM_WEIGHT_UNIFORM && wm[i] != WLSQM_WEIGHT_CENTER) {
            delete s;
            return fail(WLSQM_E_VALUE, "case %lld: unknown weighting method %d", i, wm[i]);
        }
        CaseMeta& m = s->hmeta[(size_t)i];
        memset(&m, 0, sizeof m);
        m.knowns = knowns[i] & ((1LL << no) - 1);
        m.nkn = (signed char)__builtin_popcountll((unsigned long long)m.knowns);
        m.nk = nk[i]; m.no = (short)no; m.nr = (short)(no - m.nkn);
        m.order = (signed char)order[i]; m.wm = (signed char)wm[i];
        m.op_off = off;
        off += even((m.nk + m.nkn) * (int)m.nr);
        s->maxnk = std::max(s->maxnk, m.nk);
        s->maxno = std::max(s->maxno, no);
        s->maxnr = std::max(s->maxnr, (int)m.nr);
        s->maxnq = std::max(s->maxnq, m.nk + m.nkn);
        s->maxorder = std::max(s->maxorder, (int)m.order);
        s->maxnkn = std::max(s->maxnkn, (int)m.nkn);
        if (m.knowns) s->any_knowns = true;
        if (i > 0) {
            const CaseMeta& f = s->hmeta[0];
            if (m.nk != f.nk || m.order != f.order || m.knowns != f.knowns || m.wm != f.wm) s->uniform = false;
            if (m.no != f.no) s->uniform_no = false;
            if (m.nk != f.nk || m.order != f.order || m.nkn != f.nkn || m.wm != f.wm) s->geom_uniform = false;
        }
    }
    if (ncases > 0) s->last_nk_odd = (nk[ncases - 1] & 1) != 0;
    if (all_same) off *= ncases;
    s->op_total = off;
    if (ncases > 0) {
        s->uni = s->hmeta[0];
        s->uni.op_off = 0;
        s->op_stride = even((s->uni.nk + s->uni.nkn) * (int)s->uni.nr);
    }

    int rc = use_device(s);
    if (rc) { delete s; return rc; }
    {
        int smc = 0;   // (cudaGetDeviceProperties costs milliseconds; the one attribute is all that is needed)
        if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && smc > 0) s->sm_count = smc;
        else cudaGetLastError();
    }
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; return fail(WLSQM_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    s->own_stream = true;

    auto alloc = [&](void** p, size_t bytes) -> int {
        *p = nullptr;
        if (bytes == 0) return WLSQM_OK;
        cudaError_t e2 = dev_alloc(p, bytes);
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            return fail(WLSQM_E_MEMORY, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e2));
        }
        s->bytes_state += (long long)bytes;
        return WLSQM_OK;
    };
    rc = WLSQM_OK;
    if (!rc && !s->uniform) rc = alloc((void**)&s->dmeta, sizeof(CaseMeta) * (size_t)ncases);
    if (!rc && !s->uniform_no) rc = alloc((void**)&s->dorder, (size_t)ncases);
    if (!rc) rc = alloc((void**)&s->op, (size_t)s->op_total * 8);
    if (!rc) rc = alloc((void**)&s->fi_case, (size_t)ncases * s->maxno * 8);
    if (!rc) rc = alloc((void**)&s->xi_dev, (size_t)ncases * dimension * 8);
    if (!rc && s->debug) {
        s->as_stride = s->maxnr * s->maxnr;
        rc = alloc((void**)&s->As, (size_t)ncases * s->as_stride * 8);
    }
    if (!rc && algorithm == WLSQM_ALGO_ITERATIVE) rc = alloc((void**)&s->iters_dev, ((size_t)ncases + 1) * 4);
    std::vector<int> perm;
    if (!rc && !s->uniform_no && ncases >= env_int("WLSQM_PREP_BUCKET_MIN", 8192) && ncases < (1LL << 31)) {
        long long cnt[5] = {0, 0, 0, 0, 0};
        for (long long i = 0; i < ncases; ++i) ++cnt[s->hmeta[(size_t)i].order];
        for (int o = 0; o < 5; ++o) s->order_lo[o + 1] = s->order_lo[o] + cnt[o];
        long long at[5];
        for (int o = 0; o < 5; ++o) at[o] = s->order_lo[o];
        perm.resize((size_t)ncases);
        for (long long i = 0; i < ncases; ++i) perm[(size_t)at[s->hmeta[(size_t)i].order]++] = (int)i;
        rc = alloc((void**)&s->order_perm, (size_t)ncases * 4);
    }
    if (rc) { wlsqm_solver_destroy(s); return rc; }
    if (s->order_perm) cudaMemcpyAsync(s->order_perm, perm.data(), (size_t)ncases * 4, cudaMemcpyHostToDevice, s->stream);
    if (s->dmeta)
        cudaMemcpyAsync(s->dmeta, s->hmeta.data(), sizeof(CaseMeta) * (size_t)ncases, cudaMemcpyHostToDevice, s->stream);
    if (s->dorder) {
        std::vector<signed char> ho((size_t)ncases);
        for (long long i = 0; i < ncases; ++i) ho[(size_t)i] = s->hmeta[(size_t)i].order;
        cudaMemcpyAsync(s->dorder, ho.data(), (size_t)ncases, cudaMemcpyHostToDevice, s->stream);
        cudaStreamSynchronize(s->stream);
    }
    if (s->fi_case) cudaMemsetAsync(s->fi_case, 0, (size_t)ncases * s->maxno * 8, s->stream);
    e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        wlsqm_solver_destroy(s);
        return fail(WLSQM_E_CUDA, "solver initialisation: %s", cudaGetErrorString(e));
    }
    *out = s;
    return WLSQM_OK;
}

int wlsqm_solver_destroy(wlsqm_solver_t* s) {
    if (!s) return WLSQM_OK;
    cudaSetDevice(s->device);
    // every stream that may still touch the solver's buffers, then the blocks go back to the pool unsynchronised
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->s_in) cudaStreamSynchronize(s->s_in);
    if (s->s_out) cudaStreamSynchronize(s->s_out);
    if (!s->lender) {
        dev_free(s->dmeta); dev_free(s->dorder); dev_free(s->op); dev_free(s->xi_dev); dev_free(s->As); dev_free(s->order_perm);
        s->xk_keep.release();
    } else {
        s->xk_keep.p = nullptr; s->xk_keep.cap = 0;
    }
    dev_free(s->fi_case); dev_free(s->iters_dev); s->st_xk.release(); s->st_fk.release(); s->st_fi.release(); s->st_sens.release();
    s->st_x.release(); s->st_I.release(); s->st_out.release();
    s->hoods_dev.release(); s->hood_x.release(); s->hood_f.release(); s->hood_fk.release();
    dev_free(s->hood_err);
    if (s->models_grid) { wlsqm_grid_destroy(s->models_grid); s->models_grid = nullptr; }
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    if (s->s_in) cudaStreamDestroy(s->s_in);
    if (s->s_out) cudaStreamDestroy(s->s_out);
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    cudaGetLastError();
    delete s;
    return WLSQM_OK;
}

// ExpertSolver(..., host=other) (expert.pyx:163-189, 243-263; infra.pyx:528-544): the guest borrows the host's
// geometry-dependent state -- here the operator blocks, per-case records, origins (and the kept xk of an iterative
// host) -- and owns only its solution copy, iteration counters and staging.  Several fields on one geometry then
// cost one set of operators.  The host must outlive the guest (the Python class keeps a reference).
int wlsqm_solver_create_guest(wlsqm_solver_t* host, int algorithm, int do_sens, int max_iter, wlsqm_solver_t** out) {
    if (!out) return fail(WLSQM_E_VALUE, "out is NULL");
    *out = nullptr;
    if (!host) return fail(WLSQM_E_VALUE, "NULL host solver");
    if (algorithm != WLSQM_ALGO_BASIC && algorithm != WLSQM_ALGO_ITERATIVE)
        return fail(WLSQM_E_VALUE, "Unknown algorithm specifier %d; see wlsqm.fitter.defs for valid specifiers ALGO_*", algorithm);
    if (algorithm == WLSQM_ALGO_ITERATIVE && host->algorithm != WLSQM_ALGO_ITERATIVE)
        return fail(WLSQM_E_VALUE, "an ALGO_ITERATIVE guest needs an ALGO_ITERATIVE host (the host keeps the geometry xk)");
    const wlsqm_solver* root = host->lender ? host->lender : host;
    wlsqm_solver* s = new (std::nothrow) wlsqm_solver();
    if (!s) return fail(WLSQM_E_MEMORY, "out of host memory");
    s->dim = root->dim; s->device = root->device; s->algorithm = algorithm; s->do_sens = do_sens ? 1 : 0;
    s->max_iter = max_iter; s->debug = root->debug; s->ncases = root->ncases;
    s->maxnk = root->maxnk; s->maxno = root->maxno; s->maxnr = root->maxnr; s->maxnq = root->maxnq;
    s->maxorder = root->maxorder; s->maxnkn = root->maxnkn;
    s->uniform = root->uniform; s->any_knowns = root->any_knowns; s->uniform_no = root->uniform_no; s->last_nk_odd = root->last_nk_odd;
    s->geom_uniform = root->geom_uniform;
    s->uni = root->uni; s->op_stride = root->op_stride; s->op_total = root->op_total;
    try {
        s->hmeta = root->hmeta;
    } catch (...) {
        delete s;
        return fail(WLSQM_E_MEMORY, "out of host memory");
    }
    s->sm_count = root->sm_count;
    s->lender = const_cast<wlsqm_solver*>(root);
    s->dmeta = root->dmeta; s->dorder = root->dorder; s->op = root->op; s->xi_dev = root->xi_dev;
    s->As = root->As; s->as_stride = root->as_stride;
    s->xk_keep.p = root->xk_keep.p; s->xk_keep.cap = 0;
    int rc = use_device(s);
    if (rc) { s->lender = nullptr; s->dmeta = nullptr; s->dorder = nullptr; s->op = nullptr; s->xi_dev = nullptr; s->As = nullptr; s->xk_keep.p = nullptr; delete s; return rc; }
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { wlsqm_solver_destroy(s); return fail(WLSQM_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    s->own_stream = true;
    const size_t fb = (size_t)s->ncases * s->maxno * 8;
    if (fb && dev_alloc((void**)&s->fi_case, fb) != cudaSuccess) { cudaGetLastError(); wlsqm_solver_destroy(s); return fail(WLSQM_E_MEMORY, "cudaMalloc(%zu bytes) failed", fb); }
    s->bytes_state += (long long)fb;
    if (algorithm == WLSQM_ALGO_ITERATIVE) {
        const size_t ib = ((size_t)s->ncases + 1) * 4;
        if (dev_alloc((void**)&s->iters_dev, ib) != cudaSuccess) { cudaGetLastError(); wlsqm_solver_destroy(s); return fail(WLSQM_E_MEMORY, "cudaMalloc(%zu bytes) failed", ib); }
        s->bytes_state += (long long)ib;
    }
    if (s->fi_case) cudaMemsetAsync(s->fi_case, 0, fb, s->stream);
    e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { wlsqm_solver_destroy(s); return fail(WLSQM_E_CUDA, "guest initialisation: %s", cudaGetErrorString(e)); }
    s->ready = root->ready;
    *out = s;
    return WLSQM_OK;
}

// guest.prepare(): nothing to compute -- the operators are the host's; the guest is ready when its host is
int wlsqm_solver_prepare_guest(wlsqm_solver_t* s) {
    if (!s || !s->lender) return fail(WLSQM_E_VALUE, "not a guest solver");
    if (!s->lender->ready) return fail(WLSQM_E_NOTREADY, "In guest mode, host must be in the ready state");
    int rc = use_device(s);
    if (rc) return rc;
    // the host may have re-prepared (and re-allocated its kept geometry) since the guest was created
    s->xk_keep.p = s->lender->xk_keep.p;
    if (s->models_grid) { wlsqm_grid_destroy(s->models_grid); s->models_grid = nullptr; }
    CU(cudaStreamSynchronize(s->lender->stream));     // the host's prepare may still be in flight on its stream
    s->ready = true;
    return WLSQM_OK;
}

int wlsqm_solver_set_stream(wlsqm_solver_t* s, void* cuda_stream) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    cudaStream_t next = (cudaStream_t)cuda_stream;
    if (!s->own_stream && s->stream == next) return WLSQM_OK;
    cudaSetDevice(s->device);
    if (s->own_stream && s->stream) {
        cudaStreamSynchronize(s->stream);
        cudaStreamDestroy(s->stream);
    } else {
        // work queued on the previous caller stream (a prepare, say) must be visible to what follows on the new one
        CU(order_streams(s->stream, next));
    }
    s->stream = next;
    s->own_stream = false;
    return WLSQM_OK;
}

// ---- fused result gather over peer memory (multi-GPU) -----------------------------------------------
// One process per GPU.  Every rank allocates its copy of the GLOBAL solution array with wlsqm_peer_alloc and publishes
// the 64-byte IPC handle; the others open it (wlsqm_peer_open).  wlsqm_solver_set_gather hands the solver the base
// pointers of all copies: from then on solve() stores every row it computes into each of them (its own included),
// i.e. the all-gather of fi happens inside the solve kernel, tile by tile, over NVLink.  The ranks still need one
// synchronisation per step before they read rows written by their peers (any stream-ordered barrier).
int wlsqm_peer_alloc(int device, int64_t bytes, void** ptr, void* ipc_handle_64) {
    if (!ptr || !ipc_handle_64 || bytes <= 0) return fail(WLSQM_E_VALUE, "wlsqm_peer_alloc: bad argument");
    *ptr = nullptr;
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);      // (IPC needs a plain allocation, not a block of the stream-ordered pool)
    if (e != cudaSuccess) { cudaGetLastError(); return fail(WLSQM_E_MEMORY, "cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(p); return fail(WLSQM_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_64, &h, 64);
    CU(cudaMemset(p, 0, (size_t)bytes));
    *ptr = p;
    return WLSQM_OK;
}
int wlsqm_peer_open(int device, const void* ipc_handle_64, void** ptr) {
    if (!ptr || !ipc_handle_64) return fail(WLSQM_E_VALUE, "wlsqm_peer_open: bad argument");
    *ptr = nullptr;
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return fail(WLSQM_E_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
    return WLSQM_OK;
}
int wlsqm_peer_close(void* ptr) {
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) cudaGetLastError();
    return WLSQM_OK;
}
int wlsqm_peer_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) cudaGetLastError();
    return WLSQM_OK;
}
int wlsqm_solver_keep_solution(wlsqm_solver_t* s, int keep) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    s->keep_solution = keep != 0;
    return WLSQM_OK;
}

int wlsqm_solver_set_gather(wlsqm_solver_t* s, int ntargets, void* const* bases, int64_t row0, int64_t row_stride) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (ntargets < 0 || ntargets > WLSQM_MAX_PEERS) return fail(WLSQM_E_VALUE, "between 0 and %d gather targets", WLSQM_MAX_PEERS);
    if (ntargets > 0 && (!bases || row_stride < s->maxno || row0 < 0)) return fail(WLSQM_E_VALUE, "wlsqm_solver_set_gather: bad layout");
    for (int p = 0; p < ntargets; ++p)
        if (!bases[p]) return fail(WLSQM_E_VALUE, "wlsqm_solver_set_gather: NULL target");
    s->ngather = ntargets;
    for (int p = 0; p < WLSQM_MAX_PEERS; ++p) s->gather[p] = p < ntargets ? (double*)bases[p] : nullptr;
    s->gather_row0 = row0; s->gather_s0 = row_stride;
    return WLSQM_OK;
}

int wlsqm_set_caller_stream(void* cuda_stream) {
    set_caller_stream((cudaStream_t)cuda_stream);
    return WLSQM_OK;
}

int wlsqm_solver_synchronize(wlsqm_solver_t* s) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    int rc = use_device(s);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return WLSQM_OK;
}

int wlsqm_solver_prepare(wlsqm_solver_t* s, const double* xi, int64_t xi_s0, const double* xk, int64_t xk_s0,
                         int64_t xk_s1) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (s->lender) return wlsqm_solver_prepare_guest(s);
    s->ready = false;
    if (s->models_grid) { wlsqm_grid_destroy(s->models_grid); s->models_grid = nullptr; }   // the origins may move
    if (s->ncases == 0) { s->ready = true; return WLSQM_OK; }
    if (!xi || !xk) return fail(WLSQM_E_VALUE, "xi and xk must not be NULL");
    int rc = use_device(s);
    if (rc) return rc;
    const int dim = s->dim;
    const long long n = s->ncases;
    const bool iter = s->algorithm == WLSQM_ALGO_ITERATIVE;

    // origins: always kept (interpolate and ALGO_ITERATIVE need them)
    rc = to_dense(s->xi_dev, xi, n, dim, xi_s0, s->stream);
    if (rc) return rc;

    const double* xk_use = xk;
    long long s0 = xk_s0, s1 = xk_s1;
    const bool xk_dev = is_device_ptr(xk);
    if (!xk_dev || iter) {
        DevBuf& buf = iter ? s->xk_keep : s->st_xk;
        const long long row = (long long)s->maxnk * dim;
        rc = buf.reserve((size_t)n * row * 8);
        if (rc) return rc;
        if (xk_s1 == dim) {
            rc = stage_in(s->device, (double*)buf.p, xk, n, row, xk_s0, s->stream);
            if (rc) return rc;
        } else if (xk_dev) {
            gather3_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>((double*)buf.p, xk, n, s->maxnk, dim, xk_s0, xk_s1, 1);
            CU(cudaGetLastError());
        } else {
            return fail(WLSQM_E_VALUE, "host xk must have contiguous neighbour rows (stride %lld != %d)", (long long)xk_s1, dim);
        }
        xk_use = (const double*)buf.p;
        s0 = row;
        s1 = dim;
    }

    LaunchCfg L;
    PrepRegParams R{};
    R.meta = s->dmeta; R.uni = s->uni; R.op_stride = s->op_stride; R.ncases = n;
    R.geom_uniform = s->geom_uniform ? 1 : 0;
    R.xi = s->xi_dev; R.xi_s0 = dim;
    R.xk = xk_use; R.xk_s0 = s0; R.xk_s1 = s1;
    R.op = s->op; R.As = s->As; R.as_stride = s->as_stride;
    const char* ksel = getenv("WLSQM_PREP_KERNEL");
    const bool want_smem = ksel && !strcmp(ksel, "smem");
    if (!want_smem && config_prepare_reg(s, R, L)) {
        if (s->order_perm && env_int("WLSQM_PREP_BUCKETS", 1) != 0) {
            // mixed orders: one launch per order over its case list, each with the kernel instantiated for that order
            // (an order-2 fit in the order-4 kernel pays for 16-wide rows; measured in profiles/README.md)
            for (int o = 0; o < 5; ++o) {
                const long long cnt = s->order_lo[o + 1] - s->order_lo[o];
                if (cnt < 1) continue;
                PrepRegParams Ro = R;
                LaunchCfg Lo;
                Ro.perm = s->order_perm + s->order_lo[o];
                Ro.ncases = cnt;
                if (!config_prepare_reg(s, Ro, Lo, false, o, cnt)) return fail(WLSQM_E_VALUE, "prepare: no launch shape for order %d", o);
                CU(launch_prepare_reg(dim, o, Ro, Lo.blocks, Lo.threads, Lo.smem, s->stream));
            }
        } else {
            CU(launch_prepare_reg(dim, s->maxorder, R, L.blocks, L.threads, L.smem, s->stream));
        }
    } else {
        PrepareParams P{};
        P.meta = s->dmeta; P.uni = s->uni; P.op_stride = s->op_stride; P.ncases = n;
        P.xi = s->xi_dev; P.xi_s0 = dim;
        P.xk = xk_use; P.xk_s0 = s0; P.xk_s1 = s1;
        P.op = s->op; P.As = s->As; P.as_stride = s->as_stride;
        rc = config_prepare(s, P, L);
        if (rc) return rc;
        CU(launch_prepare_smem(dim, P, L.blocks, L.threads, L.smem, s->stream));
    }
    if (!xk_dev) {
        // the call owns the host array only until it returns
        CU(cudaStreamSynchronize(s->stream));
        if (!iter) s->st_xk.release();
    }
    s->ready = true;
    return WLSQM_OK;
}

int wlsqm_solver_solve(wlsqm_solver_t* s, const double* fk, int64_t fk_s0, int64_t fk_s1, double* fi, int64_t fi_s0,
                       double* sens, int64_t sens_s0, int64_t sens_s1, int32_t* iters_out) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called before solve()");
    if (iters_out) *iters_out = 0;
    if (s->ncases == 0) return WLSQM_OK;
    if (!fk || !fi) return fail(WLSQM_E_VALUE, "fk and fi must not be NULL");
    if (s->do_sens && !sens) return fail(WLSQM_E_VALUE, "sens must be given when do_sens is set");
    int rc = use_device(s);
    if (rc) return rc;
    const long long n = s->ncases;
    const bool iter = s->algorithm == WLSQM_ALGO_ITERATIVE;
    const bool fk_dev = is_device_ptr(fk), fi_dev = is_device_ptr(fi);
    const bool sens_dev = s->do_sens ? (is_device_ptr(sens) != 0) : true;
    cudaStream_t st = s->stream;

    SolveParams P{};
    P.meta = s->dmeta; P.uni = s->uni; P.op_stride = s->op_stride; P.ncases = n; P.op = s->op;
    P.geom_uniform = s->geom_uniform ? 1 : 0;
    P.algorithm = s->algorithm; P.max_iter = s->max_iter;
    P.fi_case = s->fi_case; P.fi_case_ld = s->maxno;
    P.xi = s->xi_dev; P.xi_s0 = s->dim;
    P.xk = (const double*)s->xk_keep.p; P.xk_s0 = (long long)s->maxnk * s->dim; P.xk_s1 = s->dim;
    if (iter) {
        CU(cudaMemsetAsync(s->iters_dev, 0, 4, st));
        P.iters_max = s->iters_dev;
        P.iters_case = s->iters_dev + 1;
    }
    P.ngather = s->ngather;
    for (int p = 0; p < s->ngather; ++p) P.gather[p] = s->gather[p];
    P.gather_row0 = s->gather_row0; P.gather_s0 = s->gather_s0;
    P.gather_full = (s->ngather > 0 && s->gather_s0 % 16 == 0 && s->gather_s0 <= 32 && s->gather_s0 > s->maxno) ? 1 : 0;

    // ---- device-side views of the arguments (host arrays get dense device mirrors) -------------------
    const bool staged = !fk_dev || !fi_dev || !sens_dev;
    if (fk_dev) {
        P.fk = fk; P.fk_s0 = fk_s0; P.fk_s1 = fk_s1;
    } else {
        if (fk_s1 != 1 && s->maxnk > 1) return fail(WLSQM_E_VALUE, "host fk must have unit stride on its last axis");
        rc = s->st_fk.reserve((size_t)n * s->maxnk * 8);
        if (rc) return rc;
        P.fk = (const double*)s->st_fk.p; P.fk_s0 = s->maxnk; P.fk_s1 = 1;
    }
    bool deferred = false;
    s->fi_case_valid = true;
    if (fi_dev) {
        P.fi_in = fi; P.fi_in_s0 = fi_s0;
        // the caller's fk may be a view into fi (expert.pyx:548-555): then write back only after all cases are solved
        bool alias = false;
        if (fk_dev) {
            if (fk_s0 < 0 || fk_s1 < 0 || fi_s0 < 0) alias = true;
            else
                alias = ranges_overlap(fk, ((n - 1) * fk_s0 + (long long)(s->maxnk - 1) * fk_s1 + 1) * 8, fi,
                                       ((n - 1) * fi_s0 + s->maxno) * 8);
        }
        if (alias) deferred = true;
        else { P.fi_out = fi; P.fi_out_s0 = fi_s0; }
        if (!s->keep_solution && !deferred) P.fi_case = nullptr;
        s->fi_case_valid = P.fi_case != nullptr;
    } else if (s->any_knowns) {
        rc = s->st_fi.reserve((size_t)n * s->maxno * 8);
        if (rc) return rc;
        P.fi_in = (const double*)s->st_fi.p; P.fi_in_s0 = s->maxno;
    }
    const long long plane = (long long)s->maxnk * s->maxno;
    // host sens: the kernel writes each chunk of cases into one of two rotating device buffers (2 x chunk x nk x no
    // doubles -- not a mirror of the whole array, which is 67 GB for BASELINE.json configs[2]); a chunk goes back to the
    // host while the next one is being computed
    const bool sens_host = s->do_sens && !sens_dev;
    if (s->do_sens && sens_dev) { P.sens = sens; P.sens_s0 = sens_s0; P.sens_s1 = sens_s1; }
    // page-locked, dense, uniform: straight asynchronous copies; anything else (ordinary numpy memory, pitched rows,
    // per-case nk / no) goes through the page-locked ring and is scattered by the host threads
    const bool sens_direct = sens_host && s->uniform && s->uni.nk == s->maxnk && sens_s1 == s->maxno && sens_s0 == plane &&
                             !is_pageable_host(sens);
    const bool sens_drain = sens_host && !sens_direct;
    constexpr int SENS_BUFS = 2;

    // ---- chunked pipeline: H2D of chunk i+1, kernel of chunk i and D2H of chunk i-1 overlap ------------
    // (device-resident arguments: one chunk, everything on the solver's stream, no synchronisation)
    long long chunk = n;
    cudaStream_t s_in = st, s_out = st;
    if (staged) {
        chunk = std::max<long long>(1024, env_int("WLSQM_SOLVE_CHUNK", 65536));
        if (sens_host) {
            // a chunk of sens must fit a share of the ring slot granularity and keep the buffers modest
            const long long cap = std::max<long long>(256, (long long)((1ull << 30) / ((size_t)std::max<long long>(plane, 1) * 8)));
            chunk = std::min(chunk, cap);
            rc = s->st_sens.reserve((size_t)SENS_BUFS * (size_t)std::min(chunk, n) * plane * 8);
            if (rc) return rc;
        }
        if (!s->s_in) CU(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
        if (!s->s_out) CU(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
        s_in = s->s_in; s_out = s->s_out;
        const size_t nev = 2 * (size_t)((n + chunk - 1) / chunk) + 1 + SENS_BUFS;
        while (s->events.size() < nev) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->events.push_back(e);
        }
        // the copy streams must not run ahead of work already queued on the solver's stream
        CU(cudaEventRecord(s->events[nev - 1], st));
        CU(cudaStreamWaitEvent(s_in, s->events[nev - 1], 0));
    }
    // ordinary (pageable) numpy arrays: host threads copy through rings of page-locked slots (wlsqm_host.h)
    const bool fk_bounce = !fk_dev && n * s->maxnk * 8 >= BOUNCE_MIN_BYTES && is_pageable_host(fk) &&
                           (size_t)s->maxnk * 8 <= BOUNCE_SLOT_BYTES && env_int("WLSQM_BOUNCE", 1) != 0;
    const bool fi_bounce = !fi_dev && s->uniform_no && (fk_bounce || n * s->maxno * 8 >= BOUNCE_MIN_BYTES) &&
                           is_pageable_host(fi) && env_int("WLSQM_BOUNCE", 1) != 0 &&
                           (size_t)chunk * s->uni.no * 8 <= BOUNCE_SLOT_BYTES;
    struct RingGuard {
        bool on = false;
        ~RingGuard() { if (on) bounce_unlock(); }
    } ring_guard;
    const bool fi_scatter = !fi_dev && !s->uniform_no;       // per-case row lengths: unpacked by the host threads
    BouncePair* rings = nullptr;
    if (fk_bounce || fi_bounce || sens_drain || fi_scatter) {
        bounce_lock();
        ring_guard.on = true;
        rings = &bounce_rings(s->device);
    }
    struct PendingOut { int slot; long long c0, rows; };
    std::vector<PendingOut> pend;
    auto finish_oldest = [&]() -> int {
        const PendingOut p = pend.front();
        pend.erase(pend.begin());
        CU(d2h_bounced_finish(rings->out, p.slot, fi + p.c0 * fi_s0, fi_s0, p.rows, s->uni.no));
        return WLSQM_OK;
    };
    // one chunk of sens from its device buffer to the caller's host array (first nk_j rows x no_j columns of every case)
    struct SensChunk { long long c0 = 0, rows = 0; const double* buf = nullptr; bool live = false; } sens_prev;
    auto drain_sens = [&](const SensChunk& ch) -> int {
        while (!pend.empty()) { int r2 = finish_oldest(); if (r2) return r2; }      // (the ring is shared with the fi pieces)
        const cudaError_t e2 = d2h_pieces(rings->out, ch.buf, ch.rows, plane, s_out, [&](const double* piece, long long r0, long long nr) {
            par_for((size_t)nr, [&](size_t lo, size_t hi) {
                for (size_t r = lo; r < hi; ++r) {
                    const long long i = ch.c0 + r0 + (long long)r;
                    const CaseMeta& m = s->hmeta.size() > (size_t)i ? s->hmeta[(size_t)i] : s->uni;
                    if (m.nr < 1) continue;     // silent no-op case: sens untouched (impl.pyx:742)
                    const double* src = piece + r * plane;
                    double* dst = sens + i * sens_s0;
                    if (m.no == s->maxno && sens_s1 == s->maxno) memcpy(dst, src, (size_t)m.nk * m.no * 8);
                    else
                        for (int k = 0; k < m.nk; ++k) memcpy(dst + (long long)k * sens_s1, src + (size_t)k * s->maxno, (size_t)m.no * 8);
                }
            });
        });
        if (e2 != cudaSuccess) return fail(WLSQM_E_CUDA, "device -> host staging of sens: %s", cudaGetErrorString(e2));
        return WLSQM_OK;
    };
    LaunchCfg L;
    size_t ev = 0;
    long long chunk_index = 0;
    const size_t sens_ev0 = s->events.size() >= (size_t)SENS_BUFS ? s->events.size() - 1 - SENS_BUFS : 0;   // (staged: reserved above)
    for (long long c0 = 0; c0 < n; c0 += chunk, ++chunk_index) {
        const long long c1 = std::min(n, c0 + chunk), rows = c1 - c0;
        const int sb = (int)(chunk_index % SENS_BUFS);
        double* sens_buf = nullptr;
        if (sens_host) {
            sens_buf = (double*)s->st_sens.p + (size_t)sb * (size_t)std::min(chunk, n) * plane;
            // the kernel indexes sens by the global case number: bias the base so that case c0 lands at the buffer's start
            P.sens = reinterpret_cast<double*>(reinterpret_cast<uintptr_t>(sens_buf) - (uintptr_t)c0 * (uintptr_t)plane * 8u);
            P.sens_s0 = plane; P.sens_s1 = s->maxno;
            if (sens_direct && chunk_index >= SENS_BUFS) CU(cudaStreamWaitEvent(st, s->events[sens_ev0 + sb], 0));   // its last copy out is done
        }
        if (fk_bounce) {
            CU(h2d_bounced(rings->in, (double*)s->st_fk.p + c0 * s->maxnk, fk + c0 * fk_s0, rows, s->maxnk, fk_s0, s_in));
        } else if (!fk_dev) {
            rc = to_dense((double*)s->st_fk.p + c0 * s->maxnk, fk + c0 * fk_s0, rows, s->maxnk, fk_s0, s_in);
            if (rc) return rc;
        }
        if (!fi_dev && s->any_knowns) {
            rc = to_dense((double*)s->st_fi.p + c0 * s->maxno, fi + c0 * fi_s0, rows, s->maxno, fi_s0, s_in);
            if (rc) return rc;
        }
        if (staged) {
            CU(cudaEventRecord(s->events[ev], s_in));
            CU(cudaStreamWaitEvent(st, s->events[ev], 0));
            ++ev;
        }
        P.case_lo = c0;
        P.ncases = c1;
        if (pack_eligible(s)) {
            rc = config_solve_pack(s, P, L, rows);
            if (rc) return rc;
            CU(launch_solve_pack(P, L.blocks, L.threads, L.smem, st));
        } else {
            rc = config_solve(s, P, L, rows);
            if (rc) return rc;
            // Rows of fk with an odd element count are fetched by bulk copies of one element more (inside the even row
            // pitch, wlsqm_solve.cu).  The storage of a CALLER's strided device array may end with the last element of
            // its last row: that one case runs in a launch of its own with plain loads.
            if (fk_dev && P.f_tma && s->last_nk_odd && c1 == n) {
                if (rows > 1) {
                    P.ncases = c1 - 1;
                    rc = config_solve(s, P, L, rows - 1);
                    if (rc) return rc;
                    CU(launch_solve(s->dim, P, L.blocks, L.threads, L.smem, st));
                }
                P.case_lo = c1 - 1;
                P.ncases = c1;
                rc = config_solve(s, P, L, 1);
                if (rc) return rc;
                P.f_tma = 0;
            }
            CU(launch_solve(s->dim, P, L.blocks, L.threads, L.smem, st));
        }
        if (sens_drain) {
            // the previous chunk goes to the host while this one is being computed (its copies were queued on s_out
            // before s_out is told to wait for this chunk's kernel)
            if (sens_prev.live) { rc = drain_sens(sens_prev); if (rc) return rc; }
            sens_prev.c0 = c0; sens_prev.rows = rows; sens_prev.buf = sens_buf; sens_prev.live = true;
        }
        if (staged) {
            CU(cudaEventRecord(s->events[ev], st));
            CU(cudaStreamWaitEvent(s_out, s->events[ev], 0));
            ++ev;
        }
        if (fi_bounce) {
            if ((int)pend.size() >= BOUNCE_SLOTS - 1) { rc = finish_oldest(); if (rc) return rc; }
            const int slot = d2h_bounced_begin(rings->out, s->fi_case + c0 * s->maxno, s->maxno, rows, s->uni.no, s_out);
            if (slot < 0) return fail(WLSQM_E_CUDA, "device -> host staging: %s", cudaGetErrorString((cudaError_t)(-slot)));
            pend.push_back(PendingOut{slot, c0, rows});
        } else if (!fi_dev && s->uniform_no) {
            rc = from_dense(fi + c0 * fi_s0, fi_s0, s->fi_case + c0 * s->maxno, s->maxno, rows, s->uni.no, s_out);
            if (rc) return rc;
        }
        if (sens_direct) {
            CU(cudaMemcpyAsync(sens + c0 * plane, sens_buf, (size_t)rows * plane * 8, cudaMemcpyDeviceToHost, s_out));
            CU(cudaEventRecord(s->events[sens_ev0 + sb], s_out));
        }
    }
    if (sens_drain && sens_prev.live) { rc = drain_sens(sens_prev); if (rc) return rc; }
    while (!pend.empty()) { rc = finish_oldest(); if (rc) return rc; }
    if (deferred) CU(launch_scatter_fi(s->dmeta, s->uni, n, s->fi_case, s->maxno, fi, fi_s0, st));

    // ---- fi with per-case row lengths: pieces through the page-locked ring, scattered by the host threads ----------
    if (fi_scatter) {
        const int ld = s->maxno;
        const cudaError_t e2 = d2h_pieces(rings->out, s->fi_case, n, ld, s_out, [&](const double* piece, long long r0, long long nr) {
            par_for((size_t)nr, [&](size_t lo, size_t hi) {
                for (size_t r = lo; r < hi; ++r) {
                    const long long i = r0 + (long long)r;
                    memcpy(fi + i * fi_s0, piece + r * ld, (size_t)s->hmeta[(size_t)i].no * 8);
                }
            });
        });
        if (e2 != cudaSuccess) return fail(WLSQM_E_CUDA, "device -> host staging of fi: %s", cudaGetErrorString(e2));
    }
    int32_t it = 0;
    if (iter) CU(cudaMemcpyAsync(&it, s->iters_dev, 4, cudaMemcpyDeviceToHost, st));
    if (staged) CU(cudaStreamSynchronize(s_out));
    if (iter || staged) CU(cudaStreamSynchronize(st));
    if (iters_out) *iters_out = it;
    return WLSQM_OK;
}

int wlsqm_solver_iterations(wlsqm_solver_t* s, int32_t* out) {
    if (!s || !out) return fail(WLSQM_E_VALUE, "NULL argument");
    if (s->algorithm != WLSQM_ALGO_ITERATIVE) {
        memset(out, 0, (size_t)s->ncases * 4);
        return WLSQM_OK;
    }
    int rc = use_device(s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, s->iters_dev + 1, (size_t)s->ncases * 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return WLSQM_OK;
}

int wlsqm_solver_interpolate(wlsqm_solver_t* s, const double* x, int64_t x_s0, const int64_t* I, int64_t nx, int diff,
                             double* out, int64_t out_s0) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called first");
    if (nx == 0) return WLSQM_OK;
    if (!x || !I || !out) return fail(WLSQM_E_VALUE, "x, I and out must not be NULL");
    if (!s->fi_case_valid) return fail(WLSQM_E_NOTREADY, "the solver keeps no copy of the solution (wlsqm_solver_keep_solution(s, 0)); solve() again with the copy enabled");
    const int size = number_of_dofs(s->dim, 4);
    if (diff != WLSQM_DIFF_ALL && (diff < 0 || diff >= size)) return fail(WLSQM_E_VALUE, "invalid diff %d", diff);
    int rc = use_device(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const bool x_dev = is_device_ptr(x), I_dev = is_device_ptr(I), out_dev = is_device_ptr(out);
    const long long ow = diff == WLSQM_DIFF_ALL ? s->maxno : 1;
    InterpParams P{};
    P.dim = s->dim; P.nx = nx; P.diff = diff; P.nmodels = s->ncases;
    P.xi = s->xi_dev; P.xi_s0 = s->dim; P.fi = s->fi_case; P.fi_s0 = s->maxno;
    P.order = s->dorder; P.order_uniform = s->uni.order;
    if (x_dev) { P.x = x; P.x_s0 = x_s0; }
    else {
        rc = s->st_x.reserve((size_t)nx * s->dim * 8);
        if (rc) return rc;
        rc = stage_in(s->device, (double*)s->st_x.p, x, nx, s->dim, x_s0, st);
        if (rc) return rc;
        P.x = (const double*)s->st_x.p; P.x_s0 = s->dim;
    }
    if (I_dev) P.I = (const long long*)I;
    else {
        rc = s->st_I.reserve((size_t)nx * 8);
        if (rc) return rc;
        rc = stage_in(s->device, (double*)s->st_I.p, reinterpret_cast<const double*>(I), nx, 1, 1, st);   // (8-byte items)
        if (rc) return rc;
        P.I = (const long long*)s->st_I.p;
    }
    if (out_dev) { P.out = out; P.out_s0 = diff == WLSQM_DIFF_ALL ? out_s0 : 1; }
    else {
        rc = s->st_out.reserve((size_t)nx * ow * 8);
        if (rc) return rc;
        if (diff == WLSQM_DIFF_ALL && !s->uniform_no) CU(cudaMemsetAsync(s->st_out.p, 0, (size_t)nx * ow * 8, st));
        P.out = (double*)s->st_out.p; P.out_s0 = ow;
    }
    P.stage_no = (diff == WLSQM_DIFF_ALL && s->uniform_no && P.out_s0 == s->uni.no) ? s->uni.no : 0;
    CU(launch_interpolate(P, st));
    if (!out_dev) {
        if (diff == WLSQM_DIFF_ALL) rc = stage_out(s->device, out, out_s0, (const double*)s->st_out.p, ow, nx, ow, st);
        else rc = stage_out(s->device, out, 1, (const double*)s->st_out.p, 1, nx, 1, st);
        if (rc) return rc;
    }
    if (!x_dev || !I_dev || !out_dev) CU(cudaStreamSynchronize(st));
    return WLSQM_OK;
}

int wlsqm_solver_conds(wlsqm_solver_t* s, double* out) {
    if (!s || !out) return fail(WLSQM_E_VALUE, "NULL argument");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called before conds()");
    if (!s->debug) return fail(WLSQM_E_NOTREADY, "Not in debug mode; condition number data has not been computed");
    if (s->ncases == 0) return WLSQM_OK;
    int rc = use_device(s);
    if (rc) return rc;
    const bool out_dev = is_device_ptr(out);
    double* d = out;
    if (!out_dev) {
        rc = s->st_out.reserve((size_t)s->ncases * 8);
        if (rc) return rc;
        d = (double*)s->st_out.p;
    }
    CU(launch_cond(s->maxnr, s->ncases, s->dmeta, s->uni, s->As, s->as_stride, d, s->stream));
    if (!out_dev) CU(cudaMemcpyAsync(out, d, (size_t)s->ncases * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return WLSQM_OK;
}

int wlsqm_solver_memory(wlsqm_solver_t* s, int64_t* used, int64_t* total) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (used) *used = s->bytes_state;
    if (total) *total = s->bytes_state + staging_bytes(s);
    return WLSQM_OK;
}

int wlsqm_solver_get_fi(wlsqm_solver_t* s, double* out, int64_t out_s0) {
    if (!s || !out) return fail(WLSQM_E_VALUE, "NULL argument");
    if (s->ncases == 0) return WLSQM_OK;
    if (!s->fi_case_valid) return fail(WLSQM_E_NOTREADY, "the solver keeps no copy of the solution (wlsqm_solver_keep_solution(s, 0)); solve() again with the copy enabled");
    int rc = use_device(s);
    if (rc) return rc;
    rc = from_dense(out, out_s0, s->fi_case, s->maxno, s->ncases, s->maxno, s->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return WLSQM_OK;
}

int wlsqm_gather_hoods(const double* src, int64_t src_s0, int w, int64_t nsrc, const int32_t* idx, int64_t idx_s0, int64_t n,
                       int k, double* dst, int device, void* cuda_stream) {
    if (n == 0 || k == 0 || w == 0) return WLSQM_OK;
    if (!src || !idx || !dst) return fail(WLSQM_E_VALUE, "NULL argument");
    if (!is_device_ptr(src) || !is_device_ptr(idx) || !is_device_ptr(dst))
        return fail(WLSQM_E_VALUE, "wlsqm_gather_hoods works on device arrays");
    CU(cudaSetDevice(device));
    const long long total = (long long)n * k * w;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 32);
    // asynchronous on the caller's stream: an index outside [0, nsrc) yields NaN (no flag can be reported without a sync)
    gather_hoods_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(src, src_s0, w, idx, idx_s0, n, k, dst, nsrc, nullptr,
                                                                      -1, nullptr);
    CU(cudaGetLastError());
    return WLSQM_OK;
}

// ExpertSolver.prepare with the neighbourhoods given as index lists into a point array: xk = x[hoods] is
// gathered on the device (the caller-side gather of examples/expertsolver_example.py:59-66).
int wlsqm_solver_prepare_hoods(wlsqm_solver_t* s, const double* x, int64_t x_s0, int64_t npoints, const int32_t* hoods,
                               int64_t hoods_s0, const double* xi, int64_t xi_s0) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (s->ncases == 0) return wlsqm_solver_prepare(s, x, x_s0, x, 0, 0);
    if (!x || !hoods) return fail(WLSQM_E_VALUE, "x and hoods must not be NULL");
    if (!xi && npoints < s->ncases) return fail(WLSQM_E_VALUE, "without xi, x must hold one point per case");
    int rc = use_device(s);
    if (rc) return rc;
    const int dim = s->dim;
    const long long n = s->ncases;
    cudaStream_t st = s->stream;
    const double* xd = x;
    long long xs0 = x_s0;
    if (!is_device_ptr(x)) {
        rc = s->hood_x.reserve((size_t)npoints * dim * 8);
        if (rc) return rc;
        rc = to_dense((double*)s->hood_x.p, x, npoints, dim, x_s0, st);
        if (rc) return rc;
        xd = (const double*)s->hood_x.p;
        xs0 = dim;
    }
    rc = s->hoods_dev.reserve((size_t)n * s->maxnk * 4);
    if (rc) return rc;
    CU(cudaMemcpy2DAsync(s->hoods_dev.p, (size_t)s->maxnk * 4, hoods, (size_t)hoods_s0 * 4, (size_t)s->maxnk * 4, (size_t)n,
                         cudaMemcpyDefault, st));
    s->hood_points = npoints;
    rc = s->st_xk.reserve((size_t)n * s->maxnk * dim * 8);
    if (rc) return rc;
    const long long total = n * s->maxnk * dim;
    if (!s->hood_err) {
        if (dev_alloc((void**)&s->hood_err, 4) != cudaSuccess) { cudaGetLastError(); return fail(WLSQM_E_MEMORY, "device allocation failed"); }
    }
    CU(cudaMemsetAsync(s->hood_err, 0, 4, st));
    gather_hoods_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, (long long)s->sm_count * 32), 256, 0, st>>>(
        xd, xs0, dim, (const int32_t*)s->hoods_dev.p, s->maxnk, n, s->maxnk, (double*)s->st_xk.p, npoints, s->dmeta,
        s->uni.nk, s->hood_err);
    CU(cudaGetLastError());
    const double* xi_use = xi ? xi : xd;
    const long long xi_use_s0 = xi ? xi_s0 : xs0;
    rc = wlsqm_solver_prepare(s, xi_use, xi_use_s0, (const double*)s->st_xk.p, (long long)s->maxnk * dim, dim);
    if (rc) return rc;
    int herr = 0;
    CU(cudaMemcpyAsync(&herr, s->hood_err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (herr) {
        // the reference's caller-side gather x[hoods] raises IndexError here (examples/expertsolver_example.py:66)
        s->ready = false;
        s->hoods_dev.release();
        return fail(WLSQM_E_VALUE, "hoods: a neighbour index among the first nk[i] of some case lies outside [0, %lld)", (long long)npoints);
    }
    s->st_xk.release();      // (ALGO_ITERATIVE keeps its own copy of the geometry)
    s->hood_x.release();
    return WLSQM_OK;
}

// ExpertSolver.solve with the data given per point: fk = f[hoods] is gathered on the device, so that a time step
// moves one value per point to the GPU instead of one per (point, neighbour).
int wlsqm_solver_solve_hoods(wlsqm_solver_t* s, const double* f, int64_t f_s0, double* fi, int64_t fi_s0, double* sens,
                             int64_t sens_s0, int64_t sens_s1, int32_t* iters_out) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called before solve()");
    if (s->ncases == 0) { if (iters_out) *iters_out = 0; return WLSQM_OK; }
    if (!s->hoods_dev.p) return fail(WLSQM_E_NOTREADY, "solve_hoods() needs the neighbour lists of prepare_hoods()");
    if (!f) return fail(WLSQM_E_VALUE, "f must not be NULL");
    int rc = use_device(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const long long n = s->ncases;
    const double* fd = f;
    long long fs0 = f_s0;
    if (!is_device_ptr(f)) {
        rc = s->hood_f.reserve((size_t)s->hood_points * 8);
        if (rc) return rc;
        rc = to_dense((double*)s->hood_f.p, f, s->hood_points, 1, f_s0, st);
        if (rc) return rc;
        fd = (const double*)s->hood_f.p;
        fs0 = 1;
    }
    // rows of the gathered fk get an EVEN pitch, so that the solve kernel can fetch them with bulk copies whatever nk is
    const long long kp = ((long long)s->maxnk + 1) & ~1ll;
    rc = s->hood_fk.reserve((size_t)n * kp * 8);
    if (rc) return rc;
    const long long total = n * kp;
    // (the indices were validated against [0, hood_points) by prepare_hoods; padding slots are not read)
    gather_hoods_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, (long long)s->sm_count * 32), 256, 0, st>>>(
        fd, fs0, 1, (const int32_t*)s->hoods_dev.p, s->maxnk, n, (int)kp, (double*)s->hood_fk.p, s->hood_points, s->dmeta,
        s->uni.nk, nullptr);
    CU(cudaGetLastError());
    return wlsqm_solver_solve(s, (const double*)s->hood_fk.p, kp, 1, fi, fi_s0, sens, sens_s0, sens_s1, iters_out);
}

int wlsqm_solver_index_models(wlsqm_solver_t* s) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called before prep_interpolate()");
    if (s->models_grid || s->ncases == 0) return WLSQM_OK;
    int rc = use_device(s);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return wlsqm_grid_create(s->dim, s->ncases, s->xi_dev, s->dim, s->device, &s->models_grid);
}

int wlsqm_solver_nearest_models(wlsqm_solver_t* s, const double* x, int64_t x_s0, int64_t nx, int64_t* I_out) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->models_grid) return fail(WLSQM_E_NOTREADY, "Points xi have not been indexed; prep_interpolate() must be called before interpolate()");
    if (nx == 0) return WLSQM_OK;
    if (!x || !I_out) return fail(WLSQM_E_VALUE, "NULL argument");
    return wlsqm_grid_knn(s->models_grid, x, x_s0, nx, 1, 0, nullptr, I_out, nullptr);
}

int wlsqm_solver_interpolate_continuous(wlsqm_solver_t* s, const double* x, int64_t x_s0, int64_t nx, double r, int diff,
                                        double* out) {
    if (!s) return fail(WLSQM_E_VALUE, "NULL solver");
    if (!s->ready) return fail(WLSQM_E_NOTREADY, "Solver is not in the ready state; prepare() must be called first");
    if (!s->models_grid) return fail(WLSQM_E_NOTREADY, "Points xi have not been indexed; prep_interpolate() must be called before interpolate()");
    if (nx == 0) return WLSQM_OK;
    if (!x || !out) return fail(WLSQM_E_VALUE, "x and out must not be NULL");
    if (!(r > 0.0)) return fail(WLSQM_E_VALUE, "r must be positive");
    if (!s->fi_case_valid) return fail(WLSQM_E_NOTREADY, "the solver keeps no copy of the solution (wlsqm_solver_keep_solution(s, 0)); solve() again with the copy enabled");
    const int size = number_of_dofs(s->dim, 4);
    if (diff < 0 || diff >= size) return fail(WLSQM_E_VALUE, "invalid diff %d", diff);
    int rc = use_device(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const bool x_dev = is_device_ptr(x), out_dev = is_device_ptr(out);
    InterpParams P{};
    P.dim = s->dim; P.nx = nx; P.diff = diff;
    P.xi = s->xi_dev; P.xi_s0 = s->dim; P.fi = s->fi_case; P.fi_s0 = s->maxno;
    P.order = s->dorder; P.order_uniform = s->uni.order;
    if (x_dev) { P.x = x; P.x_s0 = x_s0; }
    else {
        rc = s->st_x.reserve((size_t)nx * s->dim * 8);
        if (rc) return rc;
        rc = stage_in(s->device, (double*)s->st_x.p, x, nx, s->dim, x_s0, st);
        if (rc) return rc;
        P.x = (const double*)s->st_x.p; P.x_s0 = s->dim;
    }
    if (out_dev) P.out = out;
    else {
        rc = s->st_out.reserve((size_t)nx * 8);
        if (rc) return rc;
        P.out = (double*)s->st_out.p;
    }
    P.out_s0 = 1;
    CU(launch_interpolate_continuous(P, grid_view(s->models_grid), r, st));
    if (!out_dev) CU(cudaMemcpyAsync(out, s->st_out.p, (size_t)nx * 8, cudaMemcpyDeviceToHost, st));
    if (!x_dev || !out_dev) CU(cudaStreamSynchronize(st));
    return WLSQM_OK;
}

// One-shot fits without sensitivities on a uniform batch: one launch of the fused kernel (assemble, equilibrate,
// factor, solve; no operator is formed, nothing is kept).  Returns 1 if the call was handled (rc_out has the result),
// 0 if the batch does not qualify and the general path (create + prepare + solve + destroy) must run.
static int fit_many_direct(int dimension, int64_t ncases, const double* xk, int64_t xk_s0, int64_t xk_s1, const double* fk,
                           int64_t fk_s0, int64_t fk_s1, const int32_t* nk, const double* xi, int64_t xi_s0, double* fi,
                           int64_t fi_s0, const int32_t* order, const int64_t* knowns, const int32_t* wm, int device,
                           int* rc_out) {
    if (env_int("WLSQM_FIT_DIRECT", 1) == 0) return 0;
    if (ncases < 1 || dimension < 1 || dimension > 3 || !nk || !order || !knowns || !wm || !xk || !fk || !xi || !fi) return 0;
    if (!scan_meta(ncases, nk, order, knowns, wm).uniform()) return 0;
    const int no = number_of_dofs(dimension, order[0]);
    if (no < 0 || nk[0] < 0 || (wm[0] != WLSQM_WEIGHT_UNIFORM && wm[0] != WLSQM_WEIGHT_CENTER)) return 0;   // general path reports it
    if (!prepare_reg_direct_ok(dimension, order[0])) return 0;
    if (wlsqm_device_count() < 1 || device < 0 || device >= wlsqm_device_count()) return 0;
    const bool xk_dev = is_device_ptr(xk), fk_dev = is_device_ptr(fk), xi_dev = is_device_ptr(xi), fi_dev = is_device_ptr(fi);
    const int dim = dimension, k = nk[0];
    if (!xk_dev && xk_s1 != dim && k > 1) return 0;
    if (!fk_dev && fk_s1 != 1 && k > 1) return 0;
    if (xk_dev && xk_s1 != dim) return 0;                      // (the kernel wants contiguous neighbour rows)
    if (fk_dev && fi_dev) {
        // fk may be a view of fi (the deferred write-back of expert.pyx:548-557 exists for that): general path
        const char *a0 = (const char*)fk, *a1 = a0 + ((ncases - 1) * fk_s0 + (long long)(k - 1) * fk_s1 + 1) * 8;
        const char *b0 = (const char*)fi, *b1 = b0 + ((ncases - 1) * fi_s0 + no) * 8;
        if (a0 < b1 && b0 < a1) return 0;
    }
    // a solver shell for the launch configuration only (no device state)
    wlsqm_solver sh;
    sh.dim = dim; sh.device = device; sh.ncases = ncases;
    CaseMeta m;
    memset(&m, 0, sizeof m);
    m.knowns = knowns[0] & ((1LL << no) - 1);
    m.nkn = (signed char)__builtin_popcountll((unsigned long long)m.knowns);
    m.nk = k; m.no = (short)no; m.nr = (short)(no - m.nkn); m.order = (signed char)order[0]; m.wm = (signed char)wm[0];
    sh.uni = m;
    sh.maxnk = k; sh.maxno = no; sh.maxnr = m.nr; sh.maxnq = k + m.nkn; sh.maxorder = m.order; sh.maxnkn = m.nkn;
    auto run = [&]() -> int {
        CU(cudaSetDevice(device));
        int smc = 0;
        if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && smc > 0) sh.sm_count = smc;
        else cudaGetLastError();
        cudaStream_t st = caller_stream();   // the caller's stream (default: the legacy default stream); synchronous result
        DevBuf bxk, bfk, bxi, bfi;
        auto done = [&](int r) { bxk.release(); bfk.release(); bxi.release(); bfi.release(); return r; };
        int rc = WLSQM_OK;
        PrepRegParams R{};
        R.meta = nullptr; R.uni = m; R.op_stride = 0; R.ncases = ncases; R.geom_uniform = 1; R.op = nullptr;
        R.xi = xi; R.xi_s0 = xi_s0; R.xk = xk; R.xk_s0 = xk_s0; R.xk_s1 = xk_s1;
        R.fk = fk; R.fk_s0 = fk_s0; R.fk_s1 = fk_s1; R.fi = fi; R.fi_s0 = fi_s0;
        if (!xk_dev) {
            rc = bxk.reserve((size_t)ncases * k * dim * 8);
            if (!rc) rc = stage_in(device, (double*)bxk.p, xk, ncases, (long long)k * dim, xk_s0, st);
            R.xk = (const double*)bxk.p; R.xk_s0 = (long long)k * dim; R.xk_s1 = dim;
        }
        if (!rc && !fk_dev) {
            rc = bfk.reserve((size_t)ncases * k * 8);
            if (!rc) rc = stage_in(device, (double*)bfk.p, fk, ncases, k, fk_s0, st);
            R.fk = (const double*)bfk.p; R.fk_s0 = k; R.fk_s1 = 1;
        }
        if (!rc && !xi_dev) {
            rc = bxi.reserve((size_t)ncases * dim * 8);
            if (!rc) rc = to_dense((double*)bxi.p, xi, ncases, dim, xi_s0, st);
            R.xi = (const double*)bxi.p; R.xi_s0 = dim;
        }
        if (!rc && !fi_dev) {
            rc = bfi.reserve((size_t)ncases * no * 8);
            if (!rc) rc = stage_in(device, (double*)bfi.p, fi, ncases, no, fi_s0, st);      // the known values travel in
            R.fi = (double*)bfi.p; R.fi_s0 = no;
        }
        if (rc) return done(rc);
        LaunchCfg L;
        if (!config_prepare_reg(&sh, R, L, true)) return done(-1000);          // does not fit: general path
        cudaError_t e = launch_prepare_reg(dim, m.order, R, L.blocks, L.threads, L.smem, st);
        if (e == cudaSuccess && !fi_dev) {
            rc = stage_out(device, fi, fi_s0, (const double*)bfi.p, no, ncases, no, st);
            if (rc) return done(rc);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return done(fail(WLSQM_E_CUDA, "one-shot fit: %s", cudaGetErrorString(e)));
        return done(WLSQM_OK);
    };
    const int rc = run();
    if (rc == -1000) return 0;
    *rc_out = rc;
    return 1;
}

int wlsqm_fit_many(int dimension, int64_t ncases, const double* xk, int64_t xk_s0, int64_t xk_s1, const double* fk,
                   int64_t fk_s0, int64_t fk_s1, const int32_t* nk, const double* xi, int64_t xi_s0, double* fi,
                   int64_t fi_s0, double* sens, int64_t sens_s0, int64_t sens_s1, int do_sens, const int32_t* order,
                   const int64_t* knowns, const int32_t* wm, int algorithm, int max_iter, int device,
                   int32_t* iters_out) {
    if (algorithm == WLSQM_ALGO_BASIC && !do_sens) {
        int rc = WLSQM_OK;
        if (fit_many_direct(dimension, ncases, xk, xk_s0, xk_s1, fk, fk_s0, fk_s1, nk, xi, xi_s0, fi, fi_s0, order, knowns,
                            wm, device, &rc)) {
            if (iters_out) *iters_out = 0;
            return rc;
        }
    }
    wlsqm_solver_t* s = nullptr;
    int rc = wlsqm_solver_create(dimension, ncases, nk, order, knowns, wm, algorithm, do_sens, max_iter, 0, device, &s);
    if (rc) return rc;
    // the temporary solver works on its own stream: order it after what the caller has queued (device-array arguments
    // may still be being produced, or their blocks still be in use, on the caller's stream)
    if (order_after_caller(s->stream) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(caller_stream()); }
    rc = wlsqm_solver_prepare(s, xi, xi_s0, xk, xk_s0, xk_s1);
    if (!rc) rc = wlsqm_solver_solve(s, fk, fk_s0, fk_s1, fi, fi_s0, sens, sens_s0, sens_s1, iters_out);
    if (!rc) rc = wlsqm_solver_synchronize(s);
    std::string keep = g_err;
    wlsqm_solver_destroy(s);
    g_err = keep;
    return rc;
}

int wlsqm_interpolate_fit(int dimension, int order, const double* xi, const double* fi, const double* x, int64_t x_s0,
                          int64_t nx, int diff, double* out, int device) {
    if (dimension < 1 || dimension > 3) return fail(WLSQM_E_VALUE, "dimension must be 1, 2 or 3; got %d", dimension);
    const int no = number_of_dofs(dimension, order);
    if (no < 0) return fail(WLSQM_E_VALUE, "order must be 0, 1, 2, 3 or 4; got %d", order);
    const int size = number_of_dofs(dimension, 4);
    if (diff != WLSQM_DIFF_ALL && (diff < 0 || diff >= size)) return fail(WLSQM_E_VALUE, "invalid diff %d", diff);
    if (nx == 0) return WLSQM_OK;
    if (!xi || !fi || !x || !out) return fail(WLSQM_E_VALUE, "NULL argument");
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    const long long ow = diff == WLSQM_DIFF_ALL ? no : 1;
    DevBuf bm, bx, bo;
    int rc = bm.reserve((size_t)(dimension + no) * 8);
    if (!rc && !is_device_ptr(x)) rc = bx.reserve((size_t)nx * dimension * 8);
    if (!rc && !is_device_ptr(out)) rc = bo.reserve((size_t)nx * ow * 8);
    auto done = [&](int r) { bm.release(); bx.release(); bo.release(); return r; };
    if (rc) return done(rc);
    cudaStream_t st = caller_stream();
    double* dm = (double*)bm.p;
    cudaError_t e = cudaMemcpyAsync(dm, xi, (size_t)dimension * 8, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dm + dimension, fi, (size_t)no * 8, cudaMemcpyDefault, st);
    if (e != cudaSuccess) return done(fail(WLSQM_E_CUDA, "interpolate_fit upload: %s", cudaGetErrorString(e)));
    InterpParams P{};
    P.dim = dimension; P.nx = nx; P.diff = diff; P.I = nullptr;
    P.xi = dm; P.xi_s0 = dimension; P.fi = dm + dimension; P.fi_s0 = no; P.order = nullptr; P.order_uniform = order;
    if (bx.p) {
        rc = to_dense((double*)bx.p, x, nx, dimension, x_s0, st);
        if (rc) return done(rc);
        P.x = (const double*)bx.p; P.x_s0 = dimension;
    } else { P.x = x; P.x_s0 = x_s0; }
    P.out = bo.p ? (double*)bo.p : out; P.out_s0 = ow;
    P.stage_no = diff == WLSQM_DIFF_ALL ? no : 0;
    e = launch_interpolate(P, st);
    if (e == cudaSuccess && bo.p) e = cudaMemcpyAsync(out, bo.p, (size_t)nx * ow * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(WLSQM_E_CUDA, "interpolate_fit: %s", cudaGetErrorString(e)));
    return done(WLSQM_OK);
}

static int lapack_common(int n, int64_t nlhs, double* A, int32_t* ipiv, double* b, int device, int do_factor, int do_solve,
                         int sym = 0) {
    if (n < 0 || nlhs < 0) return fail(WLSQM_E_VALUE, "n and nlhs must be >= 0");
    if (n == 0 || nlhs == 0) return WLSQM_OK;
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t na = (size_t)n * n * nlhs * 8, np = (size_t)n * nlhs * 4, nb = (size_t)n * nlhs * 8;
    DevBuf ba, bp, bb;
    // ipiv may be NULL for the symmetric dsysv driver (the reference's msymmetric keeps its pivots to itself)
    const bool a_dev = is_device_ptr(A), p_dev = ipiv ? is_device_ptr(ipiv) != 0 : true, b_dev = b ? is_device_ptr(b) != 0 : true;
    int rc = WLSQM_OK;
    if (!a_dev) rc = ba.reserve(na);
    if (!rc && !p_dev) rc = bp.reserve(np);
    if (!rc && !b_dev) rc = bb.reserve(nb);
    auto done = [&](int r) { ba.release(); bp.release(); bb.release(); return r; };
    if (rc) return done(rc);
    cudaStream_t st = caller_stream();
    double* dA = a_dev ? A : (double*)ba.p;
    int* dP = p_dev ? ipiv : (int*)bp.p;
    double* dB = b_dev ? b : (double*)bb.p;
    cudaError_t e = cudaSuccess;
    if (!a_dev) e = cudaMemcpyAsync(dA, A, na, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !p_dev && !do_factor) e = cudaMemcpyAsync(dP, ipiv, np, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && do_solve && !b_dev) e = cudaMemcpyAsync(dB, b, nb, cudaMemcpyHostToDevice, st);
    if (sym) {
        if (e == cudaSuccess) e = launch_sy(n, nlhs, dA, dP, do_solve ? dB : nullptr, do_factor, st);
    } else {
        if (e == cudaSuccess && do_factor && do_solve) e = launch_gesv(n, nlhs, dA, dP, dB, st);
        else if (e == cudaSuccess && do_factor) e = launch_getrf(n, nlhs, dA, dP, st);
        else if (e == cudaSuccess && do_solve) e = launch_getrs(n, nlhs, dA, dP, dB, st);
    }
    if (e == cudaSuccess && do_factor && !a_dev) e = cudaMemcpyAsync(A, dA, na, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && do_factor && !p_dev && ipiv) e = cudaMemcpyAsync(ipiv, dP, np, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && do_solve && !b_dev) e = cudaMemcpyAsync(b, dB, nb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaErrorInvalidValue) return done(fail(WLSQM_E_VALUE, "n = %d is too large for the shared-memory LU", n));
    if (e != cudaSuccess) return done(fail(WLSQM_E_CUDA, "batched LU: %s", cudaGetErrorString(e)));
    return done(WLSQM_OK);
}

int wlsqm_mgetrf(int n, int64_t nlhs, double* A, int32_t* ipiv, int device) {
    if (!A || !ipiv) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, A, ipiv, nullptr, device, 1, 0);
}
int wlsqm_mgetrs(int n, int64_t nlhs, const double* LU, const int32_t* ipiv, double* b, int device) {
    if (!LU || !ipiv || !b) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, const_cast<double*>(LU), const_cast<int32_t*>(ipiv), b, device, 0, 1);
}
int wlsqm_mgesv(int n, int64_t nlhs, double* A, int32_t* ipiv, double* b, int device) {
    if (!A || !ipiv || !b) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, A, ipiv, b, device, 1, 1);
}

int wlsqm_msytrf(int n, int64_t nlhs, double* A, int32_t* ipiv, int device) {
    if (!A || !ipiv) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, A, ipiv, nullptr, device, 1, 0, 1);
}
int wlsqm_msytrs(int n, int64_t nlhs, const double* UDU, const int32_t* ipiv, double* b, int device) {
    if (!UDU || !ipiv || !b) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, const_cast<double*>(UDU), const_cast<int32_t*>(ipiv), b, device, 0, 1, 1);
}
int wlsqm_msysv(int n, int64_t nlhs, double* A, int32_t* ipiv, double* b, int device) {
    if (!A || !b) return fail(WLSQM_E_VALUE, "NULL argument");
    return lapack_common(n, nlhs, A, ipiv, b, device, 1, 1, 1);
}
// tridiag (lapackdrivers.pyx:854-877): DGTSV with one right-hand side; a = DL (n-1 used), b = D (n), c = DU (n-1 used), x = RHS
int wlsqm_gtsv(int n, double* dl, double* d, double* du, double* b, int device) {
    if (n < 0) return fail(WLSQM_E_VALUE, "n must be >= 0");
    if (n == 0) return WLSQM_OK;
    if (!dl || !d || !du || !b) return fail(WLSQM_E_VALUE, "NULL argument");
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    const bool dev = is_device_ptr(dl) && is_device_ptr(d) && is_device_ptr(du) && is_device_ptr(b);
    const bool host = !is_device_ptr(dl) && !is_device_ptr(d) && !is_device_ptr(du) && !is_device_ptr(b);
    if (!dev && !host) return fail(WLSQM_E_VALUE, "tridiag: the four arrays must live on the same side (host or device)");
    cudaStream_t st = caller_stream();
    const size_t nb = (size_t)n * 8;
    DevBuf buf;
    double *p0 = dl, *p1 = d, *p2 = du, *p3 = b;
    if (host) {
        int rc = buf.reserve(4 * nb);
        if (rc) return rc;
        p0 = (double*)buf.p; p1 = p0 + n; p2 = p1 + n; p3 = p2 + n;
        cudaError_t e = cudaMemcpyAsync(p0, dl, nb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p1, d, nb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p2, du, nb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p3, b, nb, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { buf.release(); return fail(WLSQM_E_CUDA, "tridiag upload: %s", cudaGetErrorString(e)); }
    }
    cudaError_t e = launch_gtsv(n, p0, p1, p2, p3, st);
    if (e == cudaSuccess && host) {
        e = cudaMemcpyAsync(dl, p0, nb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, p1, nb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(du, p2, nb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b, p3, nb, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    buf.release();
    if (e != cudaSuccess) return fail(WLSQM_E_CUDA, "tridiag: %s", cudaGetErrorString(e));
    return WLSQM_OK;
}

// do_rescale (lapackdrivers.pyx:319-385) over a batch: scale every matrix in place, return the scale vectors
int wlsqm_mrescale(int nrows, int ncols, int64_t nlhs, double* A, int algo, double* row_scale, double* col_scale, int32_t* ok,
                   int device) {
    if (nrows < 0 || ncols < 0 || nlhs < 0) return fail(WLSQM_E_VALUE, "nrows, ncols and nlhs must be >= 0");
    if (algo < 1 || algo > 6) return fail(WLSQM_E_VALUE, "Unknown algorithm identifier, got %d", algo);
    if (nrows == 0 || ncols == 0 || nlhs == 0) return WLSQM_OK;
    if (!A || !row_scale || !col_scale) return fail(WLSQM_E_VALUE, "NULL argument");
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t na = (size_t)nrows * ncols * nlhs * 8, nr = (size_t)nrows * nlhs * 8, nc = (size_t)ncols * nlhs * 8, nk = (size_t)nlhs * 4;
    const bool a_dev = is_device_ptr(A), r_dev = is_device_ptr(row_scale), c_dev = is_device_ptr(col_scale);
    const bool k_dev = ok ? is_device_ptr(ok) != 0 : true;
    DevBuf ba, br, bc, bk;
    int rc = WLSQM_OK;
    if (!a_dev) rc = ba.reserve(na);
    if (!rc && !r_dev) rc = br.reserve(nr);
    if (!rc && !c_dev) rc = bc.reserve(nc);
    if (!rc && !k_dev) rc = bk.reserve(nk);
    auto done = [&](int r) { ba.release(); br.release(); bc.release(); bk.release(); return r; };
    if (rc) return done(rc);
    cudaStream_t st = caller_stream();
    double* dA = a_dev ? A : (double*)ba.p;
    double* dR = r_dev ? row_scale : (double*)br.p;
    double* dC = c_dev ? col_scale : (double*)bc.p;
    int* dK = ok ? (k_dev ? ok : (int*)bk.p) : nullptr;
    cudaError_t e = cudaSuccess;
    if (!a_dev) e = cudaMemcpyAsync(dA, A, na, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_rescale(nrows, ncols, nlhs, dA, algo, dR, dC, dK, st);
    if (e == cudaSuccess && !a_dev) e = cudaMemcpyAsync(A, dA, na, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !r_dev) e = cudaMemcpyAsync(row_scale, dR, nr, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !c_dev) e = cudaMemcpyAsync(col_scale, dC, nc, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && ok && !k_dev) e = cudaMemcpyAsync(ok, dK, nk, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaErrorInvalidValue) return done(fail(WLSQM_E_VALUE, "a %d x %d matrix is too large for the shared-memory scaler", nrows, ncols));
    if (e != cudaSuccess) return done(fail(WLSQM_E_CUDA, "batched rescale: %s", cudaGetErrorString(e)));
    return done(WLSQM_OK);
}

int wlsqm_msymmetrize(int n, int64_t nlhs, double* A, int device) {
    if (n < 0 || nlhs < 0) return fail(WLSQM_E_VALUE, "n and nlhs must be >= 0");
    if (n == 0 || nlhs == 0) return WLSQM_OK;
    if (!A) return fail(WLSQM_E_VALUE, "NULL argument");
    if (wlsqm_device_count() < 1) return fail(WLSQM_E_CUDA, "no CUDA device available (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t na = (size_t)n * n * nlhs * 8;
    const bool a_dev = is_device_ptr(A);
    DevBuf ba;
    if (!a_dev) { int rc = ba.reserve(na); if (rc) return rc; }
    cudaStream_t st = caller_stream();
    double* dA = a_dev ? A : (double*)ba.p;
    cudaError_t e = cudaSuccess;
    if (!a_dev) e = cudaMemcpyAsync(dA, A, na, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_symmetrize(n, nlhs, dA, st);
    if (e == cudaSuccess && !a_dev) e = cudaMemcpyAsync(A, dA, na, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    ba.release();
    if (e != cudaSuccess) return fail(WLSQM_E_CUDA, "msymmetrize: %s", cudaGetErrorString(e));
    return WLSQM_OK;
}

}  // extern "C"
