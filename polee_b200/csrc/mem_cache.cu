// mem_cache.cu -- process-wide caching device allocator.
//
// One prep-sample fit allocates ~6 GB at C3 (layouts, sort scratch, work buffers) and frees it again; cudaMalloc /
// cudaFree of GB-sized blocks cost tens of milliseconds each way, as much as 100 ADAM steps.  `polee prep`
// (src/main.jl:590-631) fits sample after sample of similar size in one process, so freed blocks are kept per device
// and handed out again (best fit, at most 25 % larger than the request).  Rules:
//   * dfree() synchronises the device before a block becomes reusable -- the same implicit barrier cudaFree has, so
//     callers keep cudaFree's semantics;
//   * blocks are never zeroed: callers clear what they need (as with cudaMalloc);
//   * when cudaMalloc fails the cache of that device is released and the allocation retried;
//   * the cache of a device is released when it grows past POLEE_CACHE_MAX_GB (default 48 GB: samples of very
//     different sizes would otherwise pile up blocks nobody asks for again);
//   * polee_trim_memory() returns everything to the driver; POLEE_NO_CACHE=1 turns the cache off.
#include "common.cuh"

#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <unordered_map>

namespace polee {

namespace {

struct Live {
    size_t bytes;
    int device;
};

std::mutex g_mu;
std::unordered_map<void *, Live> g_live;                 // blocks handed out
std::map<int, std::multimap<size_t, void *>> g_free;     // device -> size -> cached block
std::map<int, size_t> g_cached_bytes;
// POLEE_SETUP_TIMING: what the driver allocator cost since the last report
size_t g_stat_calls = 0, g_stat_bytes = 0, g_stat_hits = 0;
double g_stat_seconds = 0.0;

}  // namespace

std::shared_mutex &capture_mutex() {
    static std::shared_mutex mu;
    return mu;
}

namespace {

bool cache_enabled() {
    static const bool on = getenv("POLEE_NO_CACHE") == nullptr;
    return on;
}

void trim_locked(int device) {
    for (auto &dev : g_free) {
        if (device >= 0 && dev.first != device) continue;
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(dev.first);
        {
            std::unique_lock<std::shared_mutex> cap(capture_mutex());
            for (auto &blk : dev.second) cudaFree(blk.second);
        }
        dev.second.clear();
        g_cached_bytes[dev.first] = 0;
        cudaSetDevice(cur);
    }
}

}  // namespace

cudaError_t dmalloc(void **p, size_t bytes) {
    *p = nullptr;
    bytes = std::max<size_t>(bytes, 1);
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    const size_t want = (bytes + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(g_mu);
    if (cache_enabled()) {
        auto &fl = g_free[device];
        auto it = fl.lower_bound(want);
        if (it != fl.end() && it->first <= want + want / 4) {
            *p = it->second;
            g_live[*p] = Live{it->first, device};
            g_cached_bytes[device] -= it->first;
            fl.erase(it);
            ++g_stat_hits;
            return cudaSuccess;
        }
    }
    {
        std::unique_lock<std::shared_mutex> cap(capture_mutex());
        const auto t0 = std::chrono::steady_clock::now();
        e = cudaMalloc(p, want);
        g_stat_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ++g_stat_calls;
        g_stat_bytes += want;
    }
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();  // clear the sticky-free error, release the cache, retry once
        trim_locked(device);
        std::unique_lock<std::shared_mutex> cap(capture_mutex());
        e = cudaMalloc(p, want);
    }
    if (e == cudaSuccess) g_live[*p] = Live{want, device};
    return e;
}

cudaError_t dfree(void *p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_live.find(p);
    if (it == g_live.end() || !cache_enabled()) {
        if (it != g_live.end()) g_live.erase(it);
        std::unique_lock<std::shared_mutex> cap(capture_mutex());
        return cudaFree(p);
    }
    const Live blk = it->second;
    g_live.erase(it);
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != blk.device) cudaSetDevice(blk.device);
    cudaError_t e;
    {
        std::unique_lock<std::shared_mutex> cap(capture_mutex());
        e = cudaDeviceSynchronize();  // nothing in flight may still touch the block once it is reusable
    }
    if (cur != blk.device) cudaSetDevice(cur);
    g_free[blk.device].emplace(blk.bytes, p);
    g_cached_bytes[blk.device] += blk.bytes;
    static const size_t cap = (size_t)(getenv("POLEE_CACHE_MAX_GB") ? atof(getenv("POLEE_CACHE_MAX_GB")) : 48.0) << 30;
    if (g_cached_bytes[blk.device] > cap) trim_locked(blk.device);
    return e;
}

void dtrim(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    trim_locked(device);
}

void dreport(const char *what) {
    std::lock_guard<std::mutex> lk(g_mu);
    fprintf(stderr, "[polee alloc] %-28s cudaMalloc: %zu calls, %.3f GB, %.1f ms; served from the cache: %zu\n", what, g_stat_calls,
            g_stat_bytes / 1e9, g_stat_seconds * 1e3, g_stat_hits);
    g_stat_calls = g_stat_bytes = g_stat_hits = 0;
    g_stat_seconds = 0.0;
}

size_t dcached_bytes(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_cached_bytes[device];
}

}  // namespace polee
