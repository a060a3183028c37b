// wlsqm_mem.h -- device memory for the library: one stream-ordered CUDA memory pool per device.
//
// Every buffer the library owns (operators, solution copies, staging, search grids) comes from here.  The
// reference allocates its whole per-solver arena with one malloc in CaseManager_commit (wlsqm/fitter/infra.pyx:
// 545-632) and the one-shot drivers build and drop such an arena on every call (simple.pyx:731-1170); on a GPU a
// cudaMalloc / cudaFree pair per call costs more than the fits themselves (10k fits: 0.1 ms of kernels, 0.8-20 ms
// of allocation calls), so freed blocks are kept in the pool (up to WLSQM_POOL_KEEP_MB of FREE memory, default 8192;
// anything beyond that goes back to the driver when it is freed) and the next call reuses them.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace wlsqm {

// Allocate `bytes` on the CURRENT device.  The pointer is valid on every stream when the call returns.
cudaError_t dev_alloc(void** p, size_t bytes);
// Return a block.  The caller guarantees that all work touching it has completed (synchronise the streams that
// used it first); no synchronisation happens here.
void dev_free(void* p);
// dev_free for callers that cannot give that guarantee: waits for the device first (what cudaFree did implicitly).
void dev_free_sync(void* p);
// Bytes currently held by the pool of `device` (reserved from the driver, in use by the library); -1 if no pool yet.
void dev_pool_stats(int device, long long* reserved, long long* used);
// Give cached blocks back to the driver.
void dev_pool_trim(int device);

// The CUDA stream of the CALLER, for entry points that have no handle to carry one (one-shot fits, interpolate_fit,
// the search grid, the batched dense drivers): thread-local, set through wlsqm_set_caller_stream(); the default is the
// legacy default stream (which is torch's default stream).  Work of such a call runs on that stream, or -- where the
// library uses a private stream -- is ordered after everything already queued on it (order_after_caller), so that a
// kernel never reads an input before the torch op producing it has run, nor writes into a freshly allocated torch
// tensor whose block the caching allocator still has in use on the caller's stream.
void set_caller_stream(cudaStream_t s);
cudaStream_t caller_stream();
cudaError_t order_after_caller(cudaStream_t priv);
// make `later` wait for everything queued so far on `earlier` (no-op when they are the same stream)
cudaError_t order_streams(cudaStream_t earlier, cudaStream_t later);

}  // namespace wlsqm
