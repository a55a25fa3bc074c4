// Host-side state shared by the translation units of libb200mp.so.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <string>

#include <cuda_runtime.h>

#include "../../include/b200mp.h"

namespace b200mp {

constexpr unsigned kCounterRing = 256, kCounterStride = 16;

int fail(int code, const char *fmt, ...);

#define B200_CUDA(expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            cudaGetLastError();                                                                                 \
            return ::b200mp::fail(B200MP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                                       \
    } while (0)

// grow-only device scratch buffer
struct Scratch {
    char *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200MP_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        B200_CUDA(cudaMalloc(&ptr, want));
        cap = want;
        return B200MP_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace b200mp

// Small host calls (a planner thread's single isValid / checkMotion) do not queue behind the context mutex: each takes one of a few
// "lanes" -- its own stream and its own mapped pinned buffer -- so several planner threads are in flight at once.
struct b200mp_small_lane {
    cudaStream_t stream = nullptr;
    void *h = nullptr;  // host address of the mapped buffer
    void *d = nullptr;  // the same memory as the device sees it
    std::atomic<bool> busy{false};
};

struct b200mp_ctx {
    static constexpr int kLanes = 8;
    static constexpr size_t kLaneBytes = size_t(16) << 10;
    b200mp_small_lane lanes[kLanes];
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int cc_major = 0;
    cudaStream_t stream = nullptr;
    void *d_counters = nullptr;                // ring of kCounterRing work-queue heads, one per launch in flight
    std::atomic<unsigned> counter_next{0};
    // small-call path (one checkMotion / isValid at a time, the way OMPL planners call): a mapped pinned buffer the kernels
    // read and write directly, so a call is one memset + one launch + one synchronise
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // copy streams of the chunked host entry points (created on first use)
    cudaStream_t stream2 = nullptr;                           // second kernel stream of the chunked host entry points (odd chunks)
    void *h_small = nullptr;   // host address
    void *d_small = nullptr;   // the same memory as the device sees it
    size_t small_cap = 0;
    void *h_stage = nullptr;   // pinned two-slot ring the chunked fp64 entry points cast into
    size_t stage_cap = 0;
    b200mp::Scratch scratch;
    std::mutex mu;
    std::mutex build_mu;        // one roadmap build at a time per context
    void *self_comm = nullptr;  // group of one used by b200mp_prm_build (a b200mp_comm, created on first use)
    std::atomic<long long> launches{0};
    std::atomic<int> first_look{1};  // B200MP_OPT_FIRST_LOOK
};
