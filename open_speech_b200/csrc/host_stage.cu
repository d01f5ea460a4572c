// Pageable host buffers on the batch entries (osb_stt_frontend_host, osb_stt_full_host).
//
// A drop-in caller holds numpy arrays / bytes, i.e. pageable memory.  cudaMemcpyAsync from or to pageable memory is staged by the driver
// through its own bounce buffer by ONE host thread, synchronously: 1.28 GB per 256 x 60 s step took 96 ms against 15.9 ms from pinned
// buffers.  Here the staging is ours: a small pool of helper threads copies a clip group between the caller's pages and a ring of pinned
// slots (two on the way in, three on the way out) while the DMA engines and the kernels work on the neighbouring groups.
//   StageIn::src(g)   copy group g into its pinned slot (after the H2D that last used the slot has finished) and return the slot
//   StageOut::dst(g)  pinned slot the D2H of group g should write to (after the group that last used it has been handed to the caller)
//   StageOut::done(g) the D2H of group g has been enqueued: remember where it goes; hand over whatever has already landed
//   StageOut::finish  hand over the rest
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_stage.cuh"

namespace osb {

namespace {

// one copy job at a time, split into 1 MiB chunks that the helpers and the calling thread pull from the job's own counter
// (a job is its own object, kept alive by whoever still works on it: a helper that wakes late can never touch the next job's counters)
class CopyPool {
    static constexpr size_t kChunk = 1u << 20;
    struct Job {
        char* dst;
        const char* src;
        size_t bytes, n_chunks;
        std::atomic<size_t> next{0}, done{0};
    };

public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();  // never destroyed: helper threads may outlive static destructors at interpreter exit
        return *p;
    }
    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < 4 * kChunk || workers_ == 0 || !busy_.try_lock()) {  // small copy, or another caller owns the pool: plain memcpy
            memcpy(dst, src, bytes);
            return;
        }
        auto j = std::make_shared<Job>();
        j->dst = (char*)dst; j->src = (const char*)src; j->bytes = bytes; j->n_chunks = (bytes + kChunk - 1) / kChunk;
        {
            std::lock_guard<std::mutex> lk(mu_);
            cur_ = j;
            ++generation_;
        }
        cv_.notify_all();
        work(*j);
        while (j->done.load(std::memory_order_acquire) < j->n_chunks) std::this_thread::yield();
        busy_.unlock();
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw >= 16 ? 7 : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0));
        if (const char* e = getenv("OSB_COPY_THREADS")) n = atoi(e) > 0 ? atoi(e) - 1 : 0;
        if (n > 31) n = 31;
        workers_ = n;
        for (int i = 0; i < n; ++i) std::thread([this] { loop(); }).detach();
    }
    static void work(Job& j) {
        for (;;) {
            const size_t c = j.next.fetch_add(1, std::memory_order_relaxed);
            if (c >= j.n_chunks) return;
            const size_t off = c * kChunk, len = (j.bytes - off) < kChunk ? (j.bytes - off) : kChunk;
            memcpy(j.dst + off, j.src + off, len);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                j = cur_;
            }
            work(*j);
        }
    }
    std::mutex mu_, busy_;
    std::condition_variable cv_;
    unsigned long long generation_ = 0;
    std::shared_ptr<Job> cur_;
    int workers_ = 0;
};

struct PinRing {
    void* slot[3] = {nullptr, nullptr, nullptr};
    size_t cap = 0;
    int device = -1;
    int ensure(int n, size_t bytes, int dev) {
        if (device != dev || bytes > cap) {
            for (auto& s : slot) { if (s) cudaFreeHost(s); s = nullptr; }
            cap = 0;
            size_t c = 1u << 20;
            while (c < bytes) c <<= 1;
            for (int i = 0; i < n; ++i) OSB_CUDA(cudaMallocHost(&slot[i], c));
            cap = c;
            device = dev;
        }
        for (int i = 0; i < n; ++i)
            if (!slot[i]) OSB_CUDA(cudaMallocHost(&slot[i], cap));
        return OSB_OK;
    }
};

}  // namespace

bool host_is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

void host_parallel_copy(void* dst, const void* src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }

int StageIn::open(bool pageable, size_t max_group_bytes, int device) {
    on = pageable && !getenv("OSB_NO_HOST_STAGING");
    if (!on) return OSB_OK;
    static thread_local PinRing ring;
    int rc = ring.ensure(2, max_group_bytes, device);
    if (rc) return rc;
    slot[0] = ring.slot[0]; slot[1] = ring.slot[1];
    return OSB_OK;
}

int StageIn::src(int g, const void* user, size_t bytes, const cudaEvent_t* h2d_done, const void** out) {
    if (!on) { *out = user; return OSB_OK; }
    if (g >= 2) OSB_CUDA(cudaEventSynchronize(h2d_done[g - 2]));  // the H2D that read this slot last
    host_parallel_copy(slot[g & 1], user, bytes);
    *out = slot[g & 1];
    return OSB_OK;
}

int StageOut::open(bool pageable, size_t max_group_bytes, int device, cudaEvent_t* events) {
    on = pageable && !getenv("OSB_NO_HOST_STAGING");
    ev = events;
    n = 0; handed = 0;
    if (!on) return OSB_OK;
    static thread_local PinRing ring;
    int rc = ring.ensure(3, max_group_bytes, device);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) slot[i] = ring.slot[i];
    return OSB_OK;
}

int StageOut::hand_over(int g) {
    OSB_CUDA(cudaEventSynchronize(ev[g]));
    host_parallel_copy(user[g], slot[g % 3], bytes[g]);
    return OSB_OK;
}

int StageOut::dst(int g, void* user_dst, void** out) {
    if (!on) { *out = user_dst; return OSB_OK; }
    while (handed + 3 <= g) {  // group g - 3 used this slot
        int rc = hand_over(handed);
        if (rc) return rc;
        ++handed;
    }
    *out = slot[g % 3];
    return OSB_OK;
}

int StageOut::done(int g, void* user_dst, size_t nbytes, cudaStream_t s_out) {
    if (!on) return OSB_OK;
    user[g] = user_dst; bytes[g] = nbytes; n = g + 1;
    OSB_CUDA(cudaEventRecord(ev[g], s_out));
    while (handed < g && cudaEventQuery(ev[handed]) == cudaSuccess) {  // whatever has landed meanwhile
        int rc = hand_over(handed);
        if (rc) return rc;
        ++handed;
    }
    cudaGetLastError();  // cudaErrorNotReady of the query is not an error
    return OSB_OK;
}

int StageOut::finish() {
    if (!on) return OSB_OK;
    while (handed < n) {
        int rc = hand_over(handed);
        if (rc) return rc;
        ++handed;
    }
    return OSB_OK;
}

}  // namespace osb
