// Host-side staging of PAGEABLE caller buffers for the host-pointer FD call (ilqg_fd_batch_host, ilqg.cu): plain C++ (threads, no
// CUDA), so that tests/test_host_copy_pool.py can drive it on a machine without a GPU.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#ifndef ILQG_HOST_MAXCHUNKS
#define ILQG_HOST_MAXCHUNKS 32
#endif

// Pageable caller buffers in the host-pointer FD call.  A cudaMemcpyAsync from / to pageable memory is staged by the driver through
// its own bounce buffers by ONE thread at ~10 GB/s and blocks the calling thread: 86,016 hopper knots take 8 ms instead of 1.8.
// What a calcMJDerivatives-shaped caller hands over is malloc'ed memory, so the call does that staging itself: a handle-owned pinned
// mirror of the staging block, and a crew of K host threads that copy slice t of K of every chunk — the inputs of chunk c into the
// mirror before its upload is issued, its deriv / qacc blocks out of the mirror as soon as its download has landed — while the copy
// engines and the kernels work on the other chunks.  (ilqg_set_host_pinning page-locks the caller's buffers instead: faster still when
// the same buffers come back call after call, but it changes the caller's pages; this path changes nothing the caller can see.)
struct BounceCrew {
    struct Job { char* dst; const char* src; size_t bytes; };
    int K = 0, nchunks = 0;
    std::vector<Job> in_jobs[ILQG_HOST_MAXCHUNKS], out_jobs[ILQG_HOST_MAXCHUNKS];
    std::atomic<int> in_done[ILQG_HOST_MAXCHUNKS], out_ready[ILQG_HOST_MAXCHUNKS], left{0}, abort{0};
    bool started = false;
    BounceCrew() { for (int i = 0; i < ILQG_HOST_MAXCHUNKS; i++) { in_done[i].store(0); out_ready[i].store(0); } }
    static void wait(const std::atomic<int>& a, int want, const std::atomic<int>& abort) {
        for (int spin = 0; a.load(std::memory_order_acquire) < want && !abort.load(std::memory_order_relaxed);) {
            if (spin < 100000) spin++;
            if (spin >= 100000) std::this_thread::sleep_for(std::chrono::microseconds(50));   // (a pass of seconds: stop burning the core)
            else if (spin > 2000) std::this_thread::yield();
        }
    }
    // a copy that does not pull the destination through the cache (each byte is written once and read by somebody else later):
    // streaming stores where the pointers allow, memcpy for the rest
    static void copy_stream(char* dst, const char* src, size_t bytes) {
#if defined(__x86_64__)
        if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
            const size_t body = bytes & ~(size_t)63;
            for (size_t o = 0; o < body; o += 64) {
                const __m128i a = _mm_load_si128((const __m128i*)(src + o)), b = _mm_load_si128((const __m128i*)(src + o + 16)),
                              c = _mm_load_si128((const __m128i*)(src + o + 32)), d = _mm_load_si128((const __m128i*)(src + o + 48));
                _mm_stream_si128((__m128i*)(dst + o), a);
                _mm_stream_si128((__m128i*)(dst + o + 16), b);
                _mm_stream_si128((__m128i*)(dst + o + 32), c);
                _mm_stream_si128((__m128i*)(dst + o + 48), d);
            }
            _mm_sfence();
            dst += body; src += body; bytes -= body;
        }
#endif
        if (bytes) memcpy(dst, src, bytes);
    }
    static void slice(const Job& j, int t, int K) {   // thread t's part of a job, cut at 4 KB boundaries of the byte range
        const size_t per = ((j.bytes + K - 1) / K + 4095) & ~(size_t)4095, lo = per * t;
        if (lo >= j.bytes) return;
        copy_stream(j.dst + lo, j.src + lo, lo + per <= j.bytes ? per : j.bytes - lo);
    }
    void run(int t) {
        for (int c = 0; c < nchunks && !abort.load(std::memory_order_relaxed); c++) {
            for (const Job& j : in_jobs[c]) slice(j, t, K);
            in_done[c].fetch_add(1, std::memory_order_release);
        }
        for (int c = 0; c < nchunks; c++) {
            wait(out_ready[c], 1, abort);
            if (abort.load(std::memory_order_relaxed)) break;
            for (const Job& j : out_jobs[c]) slice(j, t, K);
        }
        left.fetch_add(1, std::memory_order_release);
    }
    void inputs_of(int c) { if (K) wait(in_done[c], K, abort); }
    void landed(int c) { if (K) out_ready[c].store(1, std::memory_order_release); }
    void finish() {   // every worker has left run(): nothing points into this object any more
        if (!started) return;
        const std::atomic<int> never{0};
        wait(left, K, never);
        started = false;
    }
    ~BounceCrew() { abort.store(1); finish(); }   // (an early error return: the workers leave at their next wait)
};

// The crew's threads belong to the handle and sleep between calls (creating eight threads costs ~0.2 ms, a tenth of a call).
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    BounceCrew* job = nullptr;
    unsigned gen = 0;
    bool stop = false;
    void worker(int t, unsigned seen) {   // `seen`: the generation at the thread's creation (a thread created for a later crew must not
                                          // take the hand-over of an earlier one for news)
        for (;;) {
            BounceCrew* j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                j = job;
            }
            j->run(t);
        }
    }
    void shutdown() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& x : th) if (x.joinable()) x.join();
        th.clear();
        stop = false;
    }
    void launch(BounceCrew* c) {
        if ((int)th.size() != c->K) {
            shutdown();
            const unsigned g = gen;   // (only this thread writes gen)
            for (int t = 0; t < c->K; t++) th.emplace_back([this, t, g] { worker(t, g); });
        }
        c->started = true;
        { std::lock_guard<std::mutex> lk(mu); job = c; gen++; }
        cv.notify_all();
    }
    ~CopyPool() { shutdown(); }
};
