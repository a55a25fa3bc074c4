// A resident validation kernel for single calls (a serial planner's checkMotion / isValid: RRT-Connect, BASELINE config 1).
// One small block stays on an SM with the scene staged in shared memory and polls a mailbox in mapped host memory; a request is a
// pair of configurations written by the host, the answer one word written back -- no launch, no stream synchronise per call.
// Decisions are those of validate_motions_kernel: the same check_state on the same states (start + v * pct, sequential back-steps),
// the warps of the block splitting a pass's rakes between them. The kernel leaves by itself after `idle_ns` without a request (the
// host relaunches it on the next call), so a forgotten session cannot hold the GPU.
#pragma once
#include "validate.cuh"

namespace b200mp {

constexpr int kSessionWarps = 4;  // 16 of the reference's rakes per pass
constexpr int kSessionThreads = 32 * kSessionWarps;
enum { kSessionMotion = 1, kSessionQuit = 2 };

// mapped host memory; the request and the answer sit on different 128-byte lines
struct SessionMailbox {
    // host -> device
    unsigned seq_req;      // incremented by the host AFTER the request below is written
    unsigned pad0[31];
    unsigned cmd;          // the request proper: 64 words, read by one warp with two coalesced loads
    int resolution;
    float q0[16], q1[16];
    unsigned pad1[30];
    // device -> host
    unsigned seq_ack;      // = seq_req once `result` is valid
    unsigned result;
    unsigned alive;        // 1 while the kernel polls, 0 once it has left
    unsigned pad2[29];
};
static_assert(sizeof(SessionMailbox) == 4 * 128, "mailbox layout");

__device__ __forceinline__ unsigned ld_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f(const float *p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned *p, unsigned v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class R, bool MASKED, bool GRID>
__global__ void __launch_bounds__(kSessionThreads, 1)
session_kernel(SceneDesc sd, SessionMailbox *mb, unsigned first_seq, unsigned long long idle_ns) {
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long bar;
    __shared__ SphereLevels<R> lv;
    __shared__ SharedViews sv;
    __shared__ float s_q0[R::kDim], s_q1[R::kDim];
    __shared__ int s_res;
    __shared__ unsigned s_cmd, s_seq, s_bad;
    fill_levels<R>(sd, lv);
    stage_scene(smem, sd.blob, sd.blob_floats, &bar);
    const Scene sc = scene_view(sd, smem);
    float *slots = smem + sd.blob_floats + threadIdx.x;
    fill_views(sd, sc, sv);
    __syncthreads();
    constexpr int D = R::kDim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float pct = static_cast<float>((lane & 7) + 1) / 8.0f;
    unsigned last = first_seq;
    unsigned long long t_last = now_ns();
    for (;;) {
        if (warp == 0) {
            // lane 0 polls the sequence word (one 4-byte read over the bus per look); the request is then fetched by the whole warp at
            // once -- a lane per word -- instead of one dependent bus round trip per value
            unsigned s = last;
            bool quit = false;
            if (lane == 0) {
                for (;;) {
                    s = ld_sys(&mb->seq_req);
                    if (s != last) break;
                    if (now_ns() - t_last > idle_ns) { quit = true; break; }  // nobody called for a while: leave
                }
            }
            s = __shfl_sync(kFull, s, 0);
            quit = __shfl_sync(kFull, quit ? 1 : 0, 0) != 0;
            const unsigned *req = &mb->cmd;
            const unsigned w0 = quit ? 0u : ld_sys(req + lane), w1 = quit ? 0u : ld_sys(req + 32 + lane);
            // words: 0 cmd, 1 resolution, 2 .. 17 q0, 18 .. 33 q1
            if (lane == 0) { s_seq = s; s_cmd = quit ? static_cast<unsigned>(kSessionQuit) : w0; s_bad = 0u; }
            if (lane == 1) s_res = static_cast<int>(w0);
            if (lane >= 2 && lane - 2 < D) s_q0[lane - 2] = __uint_as_float(w0);
            if (lane >= 18 && lane - 18 < D) s_q1[lane - 18] = __uint_as_float(w0);
            if (lane + 32 >= 18 && lane + 32 - 18 < D) s_q1[lane + 32 - 18] = __uint_as_float(w1);
        }
        __syncthreads();
        if (s_cmd != kSessionMotion) {
            if (threadIdx.x == 0) {
                st_sys(&mb->alive, 0u);
                __threadfence_system();
                if (s_seq != last) st_sys(&mb->seq_ack, s_seq);  // an explicit quit is acknowledged
            }
            return;
        }
        // the motion, exactly as validate_motions_kernel walks it; warp w of a pass takes rakes base + 4 w .. base + 4 w + 3
        float a[D], v[D], q[D], back[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { a[j] = s_q0[j]; v[j] = s_q1[j] - a[j]; }
        const float dist = vamp_l2_norm<D>(v);
        const float nf = fmaxf(ceilf(dist / 8.0f * static_cast<float>(s_res)), 1.0f);
        const int n = static_cast<int>(nf);
        const float denom = static_cast<float>(8 * n);
#pragma unroll
        for (int j = 0; j < D; ++j) { q[j] = fmaf(v[j], pct, a[j]); back[j] = v[j] / denom; }
        // Latency, not throughput: a short motion is checked by ALL the warps together -- each takes a share of the links and of the
        // self-collision pairs of the same states (check_state<..., SLICED>) -- a longer one spreads its rakes over the warps.
        const int nslice = n <= 4 ? kSessionWarps : (n <= 8 ? kSessionWarps / 2 : 1);
        const int slice = warp % nslice, groups = kSessionWarps / nslice;
        const int my_rake = 4 * (warp / nslice) + (lane >> 3);
        int at = 0;  // back-steps applied so far (sequential, as the reference accumulates them)
        for (int base = 0; base < n; base += 4 * groups) {
            const int rake = base + my_rake;
            for (; at < min(rake, n); ++at) {
#pragma unroll
                for (int j = 0; j < D; ++j) q[j] = q[j] - back[j];
            }
            const bool active = rake < n;
            if (__any_sync(kFull, active)) {
                const bool good = check_state<R, 8, true, MASKED, GRID, kSessionThreads, false, true>(q, sd, sc, slots, lv, sv, active, 0u, Share<true>{slice, nslice});
                if (__any_sync(kFull, active && !good) && lane == 0) s_bad = 1u;
            }
            __syncthreads();
            if (s_bad) break;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            st_sys(&mb->result, s_bad ? 0u : 1u);
            __threadfence_system();
            st_sys(&mb->seq_ack, s_seq);
            last = s_seq;
            t_last = now_ns();
        }
        last = s_seq;  // (every thread keeps its copy; only thread 0's matters)
        __syncthreads();
    }
}

}  // namespace b200mp
