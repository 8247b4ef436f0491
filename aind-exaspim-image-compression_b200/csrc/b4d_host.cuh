// b4d_host.cuh — copies between ordinary (pageable) host arrays and the device.
//
// The callers of the reference hand over NumPy arrays, i.e. pageable memory.  A plain
// cudaMemcpy of such a buffer is staged by the driver through one bounce buffer on one
// host thread (measured on the B200 box: ~5 GB/s, against ~50 GB/s for pinned memory),
// which costs more than the whole denoise at 6 B/voxel.  HostMover keeps a small ring of
// pinned pieces and a few copy threads: the DMA of piece i overlaps the host memcpy of
// piece i-1, and each host memcpy is split across the threads (first-touch page faults
// of a fresh output array are the slow part and they parallelise).
// Pinned or registered user buffers bypass it (one cudaMemcpyAsync, truly asynchronous).
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstddef>
#include <mutex>
#include <thread>
#include <vector>

class HostMover {
  public:
    static constexpr size_t PIECE = size_t(16) << 20;
    static constexpr int RING = 4;

    HostMover() = default;
    ~HostMover();
    HostMover(const HostMover &) = delete;
    HostMover &operator=(const HostMover &) = delete;

    // true when `p` is host memory the driver cannot DMA directly (neither cudaHostAlloc'ed nor registered)
    static bool pageable(const void *p);

    // device -> host.  Pageable destination: returns when `dst` is complete (like cudaMemcpy would),
    // the pieces are issued on `cs` and therefore ordered after the work already queued there.
    // Pinned destination: one cudaMemcpyAsync on `cs`, returns at once.
    cudaError_t d2h(void *dst, const void *src_dev, size_t bytes, cudaStream_t cs);
    // host -> device.  Pageable source: returns when `src` has been read completely and the last
    // piece is queued on `cs`.  Pinned source: one cudaMemcpyAsync on `cs`.
    cudaError_t h2d(void *dst_dev, const void *src, size_t bytes, cudaStream_t cs);

  private:
    cudaError_t ensure();
    void parallel_copy(char *dst, const char *src, size_t bytes);
    void worker(int id);

    char *pin_[RING] = {};
    cudaEvent_t ev_[RING] = {};
    bool ready_ = false;

    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    unsigned long long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    char *job_dst_ = nullptr;
    const char *job_src_ = nullptr;
    size_t job_bytes_ = 0;
    int job_parts_ = 1;
};
