// wlsqm_host.cu -- see wlsqm_host.h
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "wlsqm_host.h"

namespace wlsqm {

bool is_pageable_host(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

namespace {

// a small persistent pool: parallel_for over [0, n) in contiguous pieces
class HostPool {
public:
    static HostPool& get() {
        static HostPool p;
        return p;
    }
    int threads() const { return (int)workers_.size() + 1; }
    void parallel_for(size_t n, const std::function<void(size_t, size_t)>& fn) {
        const int T = threads();
        if (n == 0) return;
        if (T == 1 || n < 2) { fn(0, n); return; }
        std::unique_lock<std::mutex> call(call_mu_);          // one parallel region at a time
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; parts_ = T; pending_ = T - 1; ++epoch_;
        }
        cv_.notify_all();
        run_part(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        int T = 0;
        const char* e = getenv("WLSQM_HOST_THREADS");
        if (e && *e) T = atoi(e);
        if (T <= 0) T = std::min(8, std::max(1, (int)std::thread::hardware_concurrency() / 2));
        for (int i = 1; i < T; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run_part(int i) {
        const size_t per = (n_ + parts_ - 1) / parts_;
        const size_t lo = std::min(n_, per * i), hi = std::min(n_, lo + per);
        if (lo < hi) (*fn_)(lo, hi);
    }
    void loop(int i) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            run_part(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t, size_t)>* fn_ = nullptr;
    size_t n_ = 0;
    int parts_ = 1, pending_ = 0;
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};

}  // namespace

void par_for(size_t n, const std::function<void(size_t, size_t)>& fn) { HostPool::get().parallel_for(n, fn); }

void par_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows) {
    if (rows == 0 || row_bytes == 0) return;
    char* d = (char*)dst;
    const char* s = (const char*)src;
    if (dst_pitch == row_bytes && src_pitch == row_bytes) {
        // one contiguous run: split by bytes (64 KB granules)
        const size_t total = rows * row_bytes, gran = 64u << 10, pieces = (total + gran - 1) / gran;
        if (total < (1u << 20)) { memcpy(d, s, total); return; }
        HostPool::get().parallel_for(pieces, [&](size_t lo, size_t hi) {
            const size_t b0 = lo * gran, b1 = std::min(total, hi * gran);
            memcpy(d + b0, s + b0, b1 - b0);
        });
        return;
    }
    if (rows * row_bytes < (1u << 20)) {
        for (size_t r = 0; r < rows; ++r) memcpy(d + r * dst_pitch, s + r * src_pitch, row_bytes);
        return;
    }
    HostPool::get().parallel_for(rows, [&](size_t lo, size_t hi) {
        for (size_t r = lo; r < hi; ++r) memcpy(d + r * dst_pitch, s + r * src_pitch, row_bytes);
    });
}

cudaError_t BounceRing::init() {
    for (int i = 0; i < BOUNCE_SLOTS; ++i) {
        if (!slot[i]) {
            cudaError_t e = cudaHostAlloc(&slot[i], BOUNCE_SLOT_BYTES, cudaHostAllocPortable);
            if (e != cudaSuccess) { slot[i] = nullptr; return e; }
        }
        if (!ev[i]) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) { ev[i] = nullptr; return e; }
        }
    }
    return cudaSuccess;
}

void BounceRing::destroy() {
    for (int i = 0; i < BOUNCE_SLOTS; ++i) {
        if (ev[i]) { if (busy[i]) cudaEventSynchronize(ev[i]); cudaEventDestroy(ev[i]); ev[i] = nullptr; }
        if (slot[i]) { cudaFreeHost(slot[i]); slot[i] = nullptr; }
        busy[i] = false;
    }
}

cudaError_t h2d_bounced(BounceRing& ring, double* dst, const double* src, long long rows, long long width, long long pitch,
                        cudaStream_t stream) {
    if (rows <= 0 || width <= 0) return cudaSuccess;
    cudaError_t e = ring.init();
    if (e != cudaSuccess) return e;
    const size_t row_bytes = (size_t)width * 8;
    const long long per = std::max<long long>(1, (long long)(BOUNCE_SLOT_BYTES / row_bytes));
    if (row_bytes > BOUNCE_SLOT_BYTES) return cudaErrorInvalidValue;
    for (long long r0 = 0; r0 < rows; r0 += per) {
        const long long nr = std::min(per, rows - r0);
        const int s = ring.next;
        ring.next = (ring.next + 1) % BOUNCE_SLOTS;
        if (ring.busy[s]) {
            e = cudaEventSynchronize(ring.ev[s]);
            if (e != cudaSuccess) return e;
            ring.busy[s] = false;
        }
        par_copy_rows(ring.slot[s], row_bytes, src + r0 * pitch, (size_t)pitch * 8, row_bytes, (size_t)nr);
        e = cudaMemcpyAsync(dst + r0 * width, ring.slot[s], (size_t)nr * row_bytes, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaEventRecord(ring.ev[s], stream);
        if (e != cudaSuccess) return e;
        ring.busy[s] = true;
    }
    return cudaSuccess;
}

namespace {
BouncePair g_rings[64];
std::mutex g_ring_mu;
}  // namespace
BouncePair& bounce_rings(int device) { return g_rings[(device >= 0 && device < 64) ? device : 0]; }
void bounce_lock() { g_ring_mu.lock(); }
void bounce_unlock() { g_ring_mu.unlock(); }

int d2h_bounced_begin(BounceRing& ring, const double* src, long long src_pitch, long long rows, long long width,
                      cudaStream_t stream) {
    cudaError_t e = ring.init();
    if (e != cudaSuccess) return -(int)e;
    const size_t row_bytes = (size_t)width * 8;
    if ((size_t)rows * row_bytes > BOUNCE_SLOT_BYTES) return -(int)cudaErrorInvalidValue;
    const int s = ring.next;
    ring.next = (ring.next + 1) % BOUNCE_SLOTS;
    if (ring.busy[s]) {        // (a piece abandoned by a failed call)
        cudaEventSynchronize(ring.ev[s]);
        ring.busy[s] = false;
    }
    if (src_pitch == width)
        e = cudaMemcpyAsync(ring.slot[s], src, (size_t)rows * row_bytes, cudaMemcpyDeviceToHost, stream);
    else
        e = cudaMemcpy2DAsync(ring.slot[s], row_bytes, src, (size_t)src_pitch * 8, row_bytes, (size_t)rows,
                              cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaEventRecord(ring.ev[s], stream);
    if (e != cudaSuccess) return -(int)e;
    ring.busy[s] = true;
    return s;
}

cudaError_t d2h_bounced_finish(BounceRing& ring, int slot, double* dst, long long pitch, long long rows, long long width) {
    cudaError_t e = cudaEventSynchronize(ring.ev[slot]);
    ring.busy[slot] = false;
    if (e != cudaSuccess) return e;
    par_copy_rows(dst, (size_t)pitch * 8, ring.slot[slot], (size_t)width * 8, (size_t)width * 8, (size_t)rows);
    return cudaSuccess;
}

cudaError_t d2h_pieces(BounceRing& ring, const double* src, long long rows, long long width, cudaStream_t stream,
                       const std::function<void(const double*, long long, long long)>& unpack) {
    if (rows <= 0 || width <= 0) return cudaSuccess;
    cudaError_t e = ring.init();
    if (e != cudaSuccess) return e;
    const size_t row_bytes = (size_t)width * 8;
    if (row_bytes > BOUNCE_SLOT_BYTES) return cudaErrorInvalidValue;
    const long long per = std::max<long long>(1, (long long)(BOUNCE_SLOT_BYTES / row_bytes));
    struct Piece { long long r0, nr; int s; };
    Piece prev{0, 0, -1};
    auto finish = [&](const Piece& p) -> cudaError_t {
        cudaError_t e2 = cudaEventSynchronize(ring.ev[p.s]);
        if (e2 != cudaSuccess) return e2;
        ring.busy[p.s] = false;
        unpack((const double*)ring.slot[p.s], p.r0, p.nr);
        return cudaSuccess;
    };
    for (long long r0 = 0; r0 < rows; r0 += per) {
        const long long nr = std::min(per, rows - r0);
        const int s = ring.next;
        ring.next = (ring.next + 1) % BOUNCE_SLOTS;
        if (ring.busy[s]) {
            e = cudaEventSynchronize(ring.ev[s]);
            if (e != cudaSuccess) return e;
            ring.busy[s] = false;
        }
        e = cudaMemcpyAsync(ring.slot[s], src + r0 * width, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaEventRecord(ring.ev[s], stream);
        if (e != cudaSuccess) return e;
        ring.busy[s] = true;
        if (prev.s >= 0) {
            e = finish(prev);
            if (e != cudaSuccess) return e;
        }
        prev = Piece{r0, nr, s};
    }
    if (prev.s >= 0) e = finish(prev);
    return e;
}

cudaError_t d2h_bounced(BounceRing& ring, double* dst, long long pitch, const double* src, long long src_pitch, long long rows,
                        long long width, cudaStream_t stream) {
    if (rows <= 0 || width <= 0) return cudaSuccess;
    cudaError_t e = ring.init();
    if (e != cudaSuccess) return e;
    const size_t row_bytes = (size_t)width * 8;
    if (row_bytes > BOUNCE_SLOT_BYTES) return cudaErrorInvalidValue;
    const long long per = std::max<long long>(1, (long long)(BOUNCE_SLOT_BYTES / row_bytes));
    // software pipeline over the ring: the device copy of piece i+1 runs while the host threads unpack piece i
    struct Piece { long long r0, nr; int s; };
    Piece prev{0, 0, -1};
    auto unpack = [&](const Piece& p) -> cudaError_t {
        cudaError_t e2 = cudaEventSynchronize(ring.ev[p.s]);
        if (e2 != cudaSuccess) return e2;
        ring.busy[p.s] = false;
        par_copy_rows(dst + p.r0 * pitch, (size_t)pitch * 8, ring.slot[p.s], row_bytes, row_bytes, (size_t)p.nr);
        return cudaSuccess;
    };
    for (long long r0 = 0; r0 < rows; r0 += per) {
        const long long nr = std::min(per, rows - r0);
        const int s = ring.next;
        ring.next = (ring.next + 1) % BOUNCE_SLOTS;
        if (ring.busy[s]) {        // (only a slot left busy by an earlier h2d use of the same ring)
            e = cudaEventSynchronize(ring.ev[s]);
            if (e != cudaSuccess) return e;
            ring.busy[s] = false;
        }
        if (src_pitch == width)
            e = cudaMemcpyAsync(ring.slot[s], src + r0 * src_pitch, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, stream);
        else
            e = cudaMemcpy2DAsync(ring.slot[s], row_bytes, src + r0 * src_pitch, (size_t)src_pitch * 8, row_bytes, (size_t)nr,
                                  cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaEventRecord(ring.ev[s], stream);
        if (e != cudaSuccess) return e;
        ring.busy[s] = true;
        if (prev.s >= 0) {
            e = unpack(prev);
            if (e != cudaSuccess) return e;
        }
        prev = Piece{r0, nr, s};
    }
    if (prev.s >= 0) e = unpack(prev);
    return e;
}

}  // namespace wlsqm
