// b4d_host.cu — HostMover: pipelined, multi-threaded staging of pageable host arrays (see b4d_host.cuh).
#include "b4d_host.cuh"

#include <algorithm>
#include <cstring>

namespace {
constexpr size_t PAGE = 4096;
inline size_t part_begin(size_t bytes, int parts, int k) {
    if (k >= parts) return bytes;
    const size_t per = ((bytes + parts - 1) / parts + PAGE - 1) & ~(PAGE - 1);
    return std::min(bytes, per * (size_t)k);
}
}  // namespace

HostMover::~HostMover() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto &t : threads_) t.join();
    for (int i = 0; i < RING; ++i) {
        if (ev_[i]) {
            cudaEventSynchronize(ev_[i]);
            cudaEventDestroy(ev_[i]);
        }
        if (pin_[i]) cudaFreeHost(pin_[i]);
    }
}

bool HostMover::pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

cudaError_t HostMover::ensure() {
    if (ready_) return cudaSuccess;
    for (int i = 0; i < RING; ++i) {
        cudaError_t e = cudaHostAlloc((void **)&pin_[i], PIECE, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    const int workers = (int)std::min(5u, std::max(1u, hc / 3));
    for (int i = 0; i < workers; ++i) threads_.emplace_back(&HostMover::worker, this, i);
    ready_ = true;
    return cudaSuccess;
}

void HostMover::worker(int id) {
    unsigned long long seen = 0;
    for (;;) {
        char *dst;
        const char *src;
        size_t b0, b1;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            dst = job_dst_;
            src = job_src_;
            b0 = part_begin(job_bytes_, job_parts_, id + 1);
            b1 = part_begin(job_bytes_, job_parts_, id + 2);
        }
        if (b1 > b0) std::memcpy(dst + b0, src + b0, b1 - b0);
        {
            std::lock_guard<std::mutex> lk(mu_);
            --pending_;
        }
        cv_done_.notify_one();
    }
}

// the caller copies part 0, worker k part k + 1; returns when all parts are done
void HostMover::parallel_copy(char *dst, const char *src, size_t bytes) {
    const int parts = (int)threads_.size() + 1;
    if (bytes < (size_t(1) << 20) || parts == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_dst_ = dst;
        job_src_ = src;
        job_bytes_ = bytes;
        job_parts_ = parts;
        pending_ = parts - 1;
        ++generation_;
    }
    cv_work_.notify_all();
    const size_t b1 = part_begin(bytes, parts, 1);
    std::memcpy(dst, src, b1);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
}

cudaError_t HostMover::d2h(void *dst, const void *src_dev, size_t bytes, cudaStream_t cs) {
    if (bytes == 0) return cudaSuccess;
    if (!pageable(dst)) return cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, cs);
    cudaError_t e = ensure();
    if (e != cudaSuccess) return e;
    // pieces still queued by an earlier h2d own their ring slots until their events fire
    for (int i = 0; i < RING; ++i)
        if ((e = cudaEventSynchronize(ev_[i])) != cudaSuccess) return e;
    const size_t np = (bytes + PIECE - 1) / PIECE;
    auto len = [&](size_t i) { return std::min(PIECE, bytes - i * PIECE); };
    size_t issued = 0, drained = 0;
    while (drained < np) {
        while (issued < np && issued - drained < (size_t)RING) {
            const int slot = (int)(issued % RING);
            e = cudaMemcpyAsync(pin_[slot], (const char *)src_dev + issued * PIECE, len(issued), cudaMemcpyDeviceToHost, cs);
            if (e != cudaSuccess) return e;
            if ((e = cudaEventRecord(ev_[slot], cs)) != cudaSuccess) return e;
            ++issued;
        }
        const int slot = (int)(drained % RING);
        if ((e = cudaEventSynchronize(ev_[slot])) != cudaSuccess) return e;
        parallel_copy((char *)dst + drained * PIECE, pin_[slot], len(drained));
        ++drained;
    }
    return cudaSuccess;
}

cudaError_t HostMover::h2d(void *dst_dev, const void *src, size_t bytes, cudaStream_t cs) {
    if (bytes == 0) return cudaSuccess;
    if (!pageable(src)) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, cs);
    cudaError_t e = ensure();
    if (e != cudaSuccess) return e;
    const size_t np = (bytes + PIECE - 1) / PIECE;
    for (size_t i = 0; i < np; ++i) {
        const int slot = (int)(i % RING);
        const size_t n = std::min(PIECE, bytes - i * PIECE);
        if ((e = cudaEventSynchronize(ev_[slot])) != cudaSuccess) return e;  // slot free again (no-op when never recorded)
        parallel_copy(pin_[slot], (const char *)src + i * PIECE, n);
        e = cudaMemcpyAsync((char *)dst_dev + i * PIECE, pin_[slot], n, cudaMemcpyHostToDevice, cs);
        if (e != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev_[slot], cs)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}
