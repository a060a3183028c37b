// wlsqm_host.h -- host side of the staged (host-pointer) paths when the caller's arrays are ordinary pageable memory.
//
// cudaMemcpyAsync from pageable memory is synchronous and runs at 10-12 GB/s (the driver bounces it through one
// small pinned buffer on one thread).  The reference's users hand over plain numpy arrays (expert.pyx:467-655 takes
// memoryviews of whatever the caller has), so that path matters: here the rows are copied by a few host threads into
// a ring of page-locked slots and go to the device with asynchronous copies, and results come back the same way.
#pragma once
#include <cstddef>
#include <functional>
#include <cuda_runtime.h>

namespace wlsqm {

// 1 = ordinary pageable host memory (neither device memory nor page-locked / registered)
bool is_pageable_host(const void* p);

// fn(lo, hi) over contiguous pieces of [0, n) on the library's host threads (the caller's thread takes one piece);
// pieces run concurrently: fn must only touch its own range or synchronise
void par_for(size_t n, const std::function<void(size_t, size_t)>& fn);

// dst[r][0..row_bytes) = src[r][0..row_bytes) for r < rows (pitches in bytes), split over the library's host threads
void par_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows);

constexpr size_t BOUNCE_SLOT_BYTES = 16u << 20;
constexpr int BOUNCE_SLOTS = 3;

// A ring of page-locked slots with one event per slot.
struct BounceRing {
    void* slot[BOUNCE_SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[BOUNCE_SLOTS] = {nullptr, nullptr, nullptr};
    bool busy[BOUNCE_SLOTS] = {false, false, false};
    int next = 0;
    cudaError_t init();          // lazily allocates the slots (cudaHostAlloc) and events
    void destroy();
};

// Host (pageable) -> device: dense device rows dst[rows][width] <- src rows of `pitch` doubles.  Returns when the
// source has been consumed; the device copies are queued on `stream` (wait on the stream / an event for the data).
cudaError_t h2d_bounced(BounceRing& ring, double* dst, const double* src, long long rows, long long width, long long pitch,
                        cudaStream_t stream);

// The two rings (host -> device, device -> host) of a device, and the lock a staged call holds while it uses them.
struct BouncePair {
    BounceRing in, out;
};
BouncePair& bounce_rings(int device);
void bounce_lock();
void bounce_unlock();

// Two-phase device -> host for pipelines: begin queues the device copy of one piece (<= one slot) into a ring slot and
// returns its index (or a negative cudaError_t); finish waits for it and unpacks into the caller's rows.
int d2h_bounced_begin(BounceRing& ring, const double* src, long long src_pitch, long long rows, long long width,
                      cudaStream_t stream);
cudaError_t d2h_bounced_finish(BounceRing& ring, int slot, double* dst, long long pitch, long long rows, long long width);

// Device -> host in pieces of at most one ring slot, unpacked by the caller: for every piece, rows [r0, r0 + nr) of the
// dense device array `src` (rows of `width` doubles) arrive in a page-locked slot and `unpack(slot, r0, nr)` runs on the
// calling thread (it may use par_for) while the device copy of the next piece is already in flight.  For results whose
// host layout is not a plain strided array (per-case row lengths).  Copies already queued on `stream` are honoured.
cudaError_t d2h_pieces(BounceRing& ring, const double* src, long long rows, long long width, cudaStream_t stream,
                       const std::function<void(const double*, long long, long long)>& unpack);

// Device -> host (pageable), rows of `width` doubles from dense device rows of `src_pitch` doubles into host rows of
// `pitch` doubles.  Copies already queued on `stream` are honoured; returns when the host array holds the data.
cudaError_t d2h_bounced(BounceRing& ring, double* dst, long long pitch, const double* src, long long src_pitch, long long rows,
                        long long width, cudaStream_t stream);

}  // namespace wlsqm
