// wlsqm_mem.cu -- see wlsqm_mem.h
#include <cstdint>
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "wlsqm_mem.h"

namespace wlsqm {
namespace {

constexpr int MAX_DEVICES = 64;

struct DevPool {
    cudaMemPool_t pool = nullptr;
    cudaStream_t stream = nullptr;   // the allocations' own stream: alloc and free are ordered on it
    bool ok = false;
    uint64_t keep = 0;              // free bytes the pool may hold on to
    std::atomic<bool> tried{false};
};
DevPool g_pools[MAX_DEVICES];
std::mutex g_mu;

DevPool* pool_of(int device) {
    if (device < 0 || device >= MAX_DEVICES) return nullptr;
    DevPool& d = g_pools[device];
    if (d.tried.load(std::memory_order_acquire)) return d.ok ? &d : nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    if (d.tried.load(std::memory_order_acquire)) return d.ok ? &d : nullptr;
    const char* off = getenv("WLSQM_POOL");
    if (!(off && off[0] == '0')) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&d.pool, &props) == cudaSuccess &&
            cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess) {
            // The pool never gives memory back on its own (the built-in release threshold counts live blocks too: a
            // solver holding 3.6 GB of operators would make every freed staging block go back to the driver at the
            // next synchronisation).  dev_free() trims it instead, keeping up to WLSQM_POOL_KEEP_MB of FREE blocks.
            const char* keep = getenv("WLSQM_POOL_KEEP_MB");
            d.keep = (uint64_t)((keep && *keep) ? atoll(keep) : 8192) << 20;
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &thr);
            d.ok = true;
        } else {
            cudaGetLastError();
        }
    }
    d.tried.store(true, std::memory_order_release);
    return d.ok ? &d : nullptr;
}

}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    DevPool* d = pool_of(device);
    if (!d) return cudaMalloc(p, bytes);
    e = cudaMallocFromPoolAsync(p, bytes, d->pool, d->stream);
    if (e == cudaErrorMemoryAllocation) {      // cached blocks of the wrong sizes may be in the way
        cudaGetLastError();
        cudaStreamSynchronize(d->stream);
        cudaMemPoolTrimTo(d->pool, 0);
        e = cudaMallocFromPoolAsync(p, bytes, d->pool, d->stream);
    }
    if (e != cudaSuccess) { *p = nullptr; return e; }
    // the block may be used on any stream once the allocation has completed on its own
    e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { *p = nullptr; return e; }
    return cudaSuccess;
}

void dev_free(void* p) {
    if (!p) return;
    cudaPointerAttributes a{};
    int device = 0;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess) device = a.device;
    else { cudaGetLastError(); cudaGetDevice(&device); }
    DevPool* d = (device >= 0 && device < MAX_DEVICES && g_pools[device].ok) ? &g_pools[device] : nullptr;
    if (!d || cudaFreeAsync(p, d->stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        cudaGetLastError();
        return;
    }
    // keep at most d->keep bytes of free blocks cached
    uint64_t reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(d->pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(d->pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess) {
        if (reserved > used && reserved - used > d->keep) {
            cudaStreamSynchronize(d->stream);          // the frees queued so far are complete
            cudaMemPoolTrimTo(d->pool, (size_t)(used + d->keep));
        }
    } else {
        cudaGetLastError();
    }
}

void dev_free_sync(void* p) {
    if (!p) return;
    cudaDeviceSynchronize();
    dev_free(p);
}

void dev_pool_stats(int device, long long* reserved, long long* used) {
    if (reserved) *reserved = -1;
    if (used) *used = -1;
    if (device < 0 || device >= MAX_DEVICES || !g_pools[device].ok) return;
    uint64_t r = 0, u = 0;
    cudaMemPoolGetAttribute(g_pools[device].pool, cudaMemPoolAttrReservedMemCurrent, &r);
    cudaMemPoolGetAttribute(g_pools[device].pool, cudaMemPoolAttrUsedMemCurrent, &u);
    if (reserved) *reserved = (long long)r;
    if (used) *used = (long long)u;
}

namespace {
thread_local cudaStream_t g_caller_stream = nullptr;
}
void set_caller_stream(cudaStream_t s) { g_caller_stream = s; }
cudaStream_t caller_stream() { return g_caller_stream; }

cudaError_t order_streams(cudaStream_t earlier, cudaStream_t later) {
    if (earlier == later) return cudaSuccess;
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(ev, earlier);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(later, ev, 0);
    cudaEventDestroy(ev);      // (released by the runtime once the recorded work has completed)
    return e;
}
cudaError_t order_after_caller(cudaStream_t priv) { return order_streams(g_caller_stream, priv); }

void dev_pool_trim(int device) {
    if (device < 0 || device >= MAX_DEVICES || !g_pools[device].ok) return;
    cudaStreamSynchronize(g_pools[device].stream);
    cudaMemPoolTrimTo(g_pools[device].pool, 0);
}

}  // namespace wlsqm
