// Host-side launches of the validation kernels, one explicit instantiation per robot (validate_<robot>.cu) so that the robots
// compile in parallel. b200mp.cu declares them `extern template`.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "validate.cuh"
#include "session.cuh"
#include "host_common.hpp"

namespace b200mp {

// resident blocks per SM of a kernel at a given dynamic shared-memory size (cached per device; defined in b200mp.cu)
int blocks_per_sm(int device, const void *kern, int threads, size_t smem, int *per_sm);
// this launch's own work-queue head from the context's ring (defined in b200mp.cu)
LaunchCounters *next_counter(b200mp_ctx *c);

template <class R, int KIND>  // 0: warp-per-motion kernel, 1: packed rakes of the short buckets, 2: first look at the long motions
inline auto pick_motion_kernel(const SceneDesc &d) {
    if constexpr (KIND == 1)
        return d.n_grid > 0 ? validate_rakes_kernel<R, false, true, false>
                            : (d.sorted ? validate_rakes_kernel<R, false, false, false> : validate_rakes_kernel<R, true, false, false>);
    else if constexpr (KIND == 2)
        return d.n_grid > 0 ? validate_rakes_kernel<R, false, true, true>
                            : (d.sorted ? validate_rakes_kernel<R, false, false, true> : validate_rakes_kernel<R, true, false, true>);
    else
        return d.n_grid > 0 ? validate_motions_kernel<R, false, true>
                            : (d.sorted ? validate_motions_kernel<R, false, false> : validate_motions_kernel<R, true, false>);
}

template <class R>
int launch_motions(b200mp_ctx *c, const SceneDesc &desc, const float *d_q0, const float *d_q1, const int2 *d_pairs, int64_t m, int resolution,
                   uint8_t *d_valid, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(desc.blob_floats) * 4 + lane_store_bytes<R>(motion_threads<R>());
    auto kern = pick_motion_kernel<R, 0>(desc);
    int per_sm = 0, rc = 0;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(kern), motion_threads<R>(), smem, &per_sm))) return rc;
    if (per_sm < 1) return fail(B200MP_ERR_CAPACITY, "validate_motions: kernel does not fit on an SM with %zu B of scene", smem);
    const long long wpb = motion_threads<R>() / 32;
    const long long blocks = std::max<long long>(1, std::min<long long>(static_cast<long long>(c->sm_count) * per_sm, (m + wpb - 1) / wpb));
    LaunchCounters *cnt = next_counter(c);
    B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(LaunchCounters), st));
    kern<<<static_cast<unsigned>(blocks), motion_threads<R>(), smem, st>>>(desc, d_q0, d_q1, d_pairs, m, resolution, d_valid, cnt, nullptr, nullptr);
    B200_CUDA(cudaGetLastError());
    c->launches.fetch_add(1);
    return B200MP_OK;
}

// Motions grouped by their number of rakes first (pre-pass on the device, no host round trip): the short ones (<= 4 rakes, e.g.
// roadmap edges) go through validate_rakes_kernel, four rakes of any motions per warp pass, the rest through validate_motions_kernel.
template <class R>
int launch_motions_bucketed(b200mp_ctx *c, const SceneDesc &desc, const float *d_q0, const float *d_q1, const int2 *d_pairs, int64_t m,
                            int resolution, uint8_t *d_valid, cudaStream_t st) {
    if (m > 0xfffffff0LL) return fail(B200MP_ERR_CAPACITY, "validate (bucketed): at most 2^32 - 16 motions per call");
    const size_t smem = static_cast<size_t>(desc.blob_floats) * 4 + lane_store_bytes<R>(motion_threads<R>());
    auto k_rakes = pick_motion_kernel<R, 1>(desc);
    auto k_first = pick_motion_kernel<R, 2>(desc);
    auto k_long = pick_motion_kernel<R, 0>(desc);
    int per_rakes = 0, per_first = 0, per_long = 0, rc = 0;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(k_rakes), motion_threads<R>(), smem, &per_rakes))) return rc;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(k_first), motion_threads<R>(), smem, &per_first))) return rc;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(k_long), motion_threads<R>(), smem, &per_long))) return rc;
    if (per_rakes < 1 || per_first < 1 || per_long < 1) return fail(B200MP_ERR_CAPACITY, "validate (bucketed): kernel does not fit on an SM with %zu B of scene", smem);
    // stream-ordered temporaries: keys[m] | list[m] | per-block counts[nb][8] | plan
    const unsigned nb = static_cast<unsigned>((m + kBucketThreads - 1) / kBucketThreads);
    const size_t o_list = (static_cast<size_t>(m) + 255) / 256 * 256, o_blk = o_list + (static_cast<size_t>(m) * 4 + 255) / 256 * 256;
    const size_t o_plan = o_blk + static_cast<size_t>(nb) * 8 * 4;
    char *tmp = nullptr;
    B200_CUDA(cudaMallocAsync(&tmp, o_plan + sizeof(BucketPlan), st));
    unsigned char *keys = reinterpret_cast<unsigned char *>(tmp);
    unsigned *list = reinterpret_cast<unsigned *>(tmp + o_list);
    unsigned *blk = reinterpret_cast<unsigned *>(tmp + o_blk);
    BucketPlan *plan = reinterpret_cast<BucketPlan *>(tmp + o_plan);
    cudaMemsetAsync(plan, 0, sizeof(BucketPlan), st);
    bucket_count_kernel<R><<<nb, kBucketThreads, 0, st>>>(d_q0, d_q1, d_pairs, m, resolution, keys, d_valid, blk, plan);
    static const bool first_look_env = !(getenv("B200MP_FIRST_LOOK") && atoi(getenv("B200MP_FIRST_LOOK")) == 0);
    const bool first_look = first_look_env && c->first_look.load() != 0;
    bucket_plan_kernel<<<1, 1024, 0, st>>>(plan, blk, static_cast<int>(nb), first_look ? 1 : 0);
    bucket_scatter_kernel<<<nb, kBucketThreads, 0, st>>>(keys, m, blk, list, plan);
    const long long wpb = motion_threads<R>() / 32;
    const long long cap = std::max<long long>(1, (m + wpb - 1) / wpb);
    const unsigned b_rakes = static_cast<unsigned>(std::min<long long>(static_cast<long long>(c->sm_count) * per_rakes, cap));
    const unsigned b_long = static_cast<unsigned>(std::min<long long>(static_cast<long long>(c->sm_count) * per_long, cap));
    const unsigned b_first = static_cast<unsigned>(std::min<long long>(static_cast<long long>(c->sm_count) * per_first, cap));
    k_rakes<<<b_rakes, motion_threads<R>(), smem, st>>>(desc, d_q0, d_q1, d_pairs, d_valid, list, plan);
    if (first_look) k_first<<<b_first, motion_threads<R>(), smem, st>>>(desc, d_q0, d_q1, d_pairs, d_valid, list, plan);
    k_long<<<b_long, motion_threads<R>(), smem, st>>>(desc, d_q0, d_q1, d_pairs, m, resolution, d_valid, &plan->long_queue, list, plan);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return fail(B200MP_ERR_CUDA, "validate (bucketed) launch failed: %s", cudaGetErrorString(e));
    c->launches.fetch_add(first_look ? 6 : 5);
    return B200MP_OK;
}

template <class R>
int launch_states(b200mp_ctx *c, const SceneDesc &desc, const float *d_q, int64_t m, uint8_t *d_valid, cudaStream_t st, const unsigned *d_order) {
    const size_t smem = static_cast<size_t>(desc.blob_floats) * 4 + lane_store_bytes<R>(state_threads<R>());
    auto kern = desc.n_grid > 0 ? validate_states_kernel<R, false, true>
                                : (desc.sorted ? validate_states_kernel<R, false, false> : validate_states_kernel<R, true, false>);
    int per_sm = 0, rc = 0;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(kern), state_threads<R>(), smem, &per_sm))) return rc;
    if (per_sm < 1) return fail(B200MP_ERR_CAPACITY, "validate_states: kernel does not fit on an SM with %zu B of scene", smem);
    const long long warps_needed = (m + 31) / 32, wpb = state_threads<R>() / 32;
    const long long blocks = std::max<long long>(1, std::min<long long>(static_cast<long long>(c->sm_count) * per_sm, (warps_needed + wpb - 1) / wpb));
    // Large batches of unrelated states (the roadmap sampler's candidates: most are invalid, yet nearly every warp holds a valid one and
    // runs the whole check with most lanes idle) are SIFTED first: a pass over the distal half of the links only, then the whole check
    // for the survivors, packed. Same decisions (a state that fails part of the check is invalid).
    static const bool sift_env = !(getenv("B200MP_SIFT") && atoi(getenv("B200MP_SIFT")) == 0);
    if (desc.n_grid > 0 && m >= (1 << 16) && m < 0xfffffff0LL && sift_env) {
        auto pre = validate_states_kernel<R, false, true, true>;
        int per_pre = 0;
        if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(pre), state_threads<R>(), smem, &per_pre))) return rc;
        if (per_pre >= 1) {
            char *tmp = nullptr;
            B200_CUDA(cudaMallocAsync(&tmp, static_cast<size_t>(m) * 4 + 256, st));
            unsigned *count = reinterpret_cast<unsigned *>(tmp), *list = reinterpret_cast<unsigned *>(tmp + 256);
            LaunchCounters *cnt1 = next_counter(c), *cnt2 = next_counter(c);
            cudaMemsetAsync(cnt1, 0, sizeof(LaunchCounters), st);
            cudaMemsetAsync(cnt2, 0, sizeof(LaunchCounters), st);
            cudaMemsetAsync(count, 0, sizeof(unsigned), st);
            const long long pre_blocks = std::max<long long>(1, std::min<long long>(static_cast<long long>(c->sm_count) * per_pre, (warps_needed + wpb - 1) / wpb));
            pre<<<static_cast<unsigned>(pre_blocks), state_threads<R>(), smem, st>>>(desc, d_q, m, d_valid, cnt1, d_order, nullptr);
            sift_survivors_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, st>>>(d_valid, d_order, m, list, count);
            kern<<<static_cast<unsigned>(blocks), state_threads<R>(), smem, st>>>(desc, d_q, m, d_valid, cnt2, list, count);
            cudaError_t e = cudaGetLastError();
            cudaFreeAsync(tmp, st);
            if (e != cudaSuccess) return fail(B200MP_ERR_CUDA, "validate_states (sifted) launch failed: %s", cudaGetErrorString(e));
            c->launches.fetch_add(3);
            return B200MP_OK;
        }
    }
    LaunchCounters *cnt = next_counter(c);
    B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(LaunchCounters), st));
    kern<<<static_cast<unsigned>(blocks), state_threads<R>(), smem, st>>>(desc, d_q, m, d_valid, cnt, d_order, nullptr);
    B200_CUDA(cudaGetLastError());
    c->launches.fetch_add(1);
    return B200MP_OK;
}

// the resident single-call kernel (session.cuh) on `st`; returns once it is launched
template <class R>
int launch_session(b200mp_ctx *c, const SceneDesc &desc, SessionMailbox *d_mailbox, unsigned first_seq, unsigned long long idle_ns, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(desc.blob_floats) * 4 + lane_store_bytes<R>(kSessionThreads);
    auto kern = desc.n_grid > 0 ? session_kernel<R, false, true> : (desc.sorted ? session_kernel<R, false, false> : session_kernel<R, true, false>);
    int per_sm = 0, rc = 0;
    if ((rc = blocks_per_sm(c->device, reinterpret_cast<const void *>(kern), kSessionThreads, smem, &per_sm))) return rc;
    if (per_sm < 1) return fail(B200MP_ERR_CAPACITY, "session: kernel does not fit on an SM with %zu B of scene", smem);
    kern<<<1, kSessionThreads, smem, st>>>(desc, d_mailbox, first_seq, idle_ns);
    B200_CUDA(cudaGetLastError());
    c->launches.fetch_add(1);
    return B200MP_OK;
}

#define B200MP_INSTANTIATE_ROBOT(R)                                                                                                        \
    template int launch_motions<R>(b200mp_ctx *, const SceneDesc &, const float *, const float *, const int2 *, int64_t, int, uint8_t *, cudaStream_t);          \
    template int launch_motions_bucketed<R>(b200mp_ctx *, const SceneDesc &, const float *, const float *, const int2 *, int64_t, int, uint8_t *, cudaStream_t); \
    template int launch_states<R>(b200mp_ctx *, const SceneDesc &, const float *, int64_t, uint8_t *, cudaStream_t, const unsigned *);                          \
    template int launch_session<R>(b200mp_ctx *, const SceneDesc &, SessionMailbox *, unsigned, unsigned long long, cudaStream_t);
#define B200MP_EXTERN_ROBOT(R)                                                                                                                    \
    extern template int launch_motions<R>(b200mp_ctx *, const SceneDesc &, const float *, const float *, const int2 *, int64_t, int, uint8_t *, cudaStream_t);          \
    extern template int launch_motions_bucketed<R>(b200mp_ctx *, const SceneDesc &, const float *, const float *, const int2 *, int64_t, int, uint8_t *, cudaStream_t); \
    extern template int launch_states<R>(b200mp_ctx *, const SceneDesc &, const float *, int64_t, uint8_t *, cudaStream_t, const unsigned *);                          \
    extern template int launch_session<R>(b200mp_ctx *, const SceneDesc &, SessionMailbox *, unsigned, unsigned long long, cudaStream_t);

}  // namespace b200mp
