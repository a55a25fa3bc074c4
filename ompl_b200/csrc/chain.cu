// Planar kinematic chain (BASELINE config 5): state validity, discrete motion validation and the chain metric, fp64.
//
// Reference behaviour reproduced (paths under the reference checkout):
//   demos/KinematicChain.h:105-121   KinematicChainSpace::distance      -> chain_distance_kernel, knn chain metric (knn.cu)
//   demos/KinematicChain.h:150-176   KinematicChainSpace::interpolate   -> interp_joint()
//   demos/KinematicChain.h:209-276   KinematicChainValidityChecker      -> chain_state_valid()
//   demos/KinematicChain.h:279-315   createHornEnvironment              -> b200mp_chain_horn_environment (host)
//   src/ompl/base/src/DiscreteMotionValidator.cpp:48-145 checkMotion x2 -> chain_motions_kernel
//   src/ompl/base/src/StateSpace.cpp:851-854 validSegmentCount          -> nd = (unsigned)ceil(distance / longestValidSegment)
//
// Mapping: one warp per state / per motion. Lane i owns link i: it accumulates theta_i = q_0 + ... + q_i in the
// reference's order, evaluates cos/sin once, and the link end points are accumulated link by link with warp shuffles so
// every sum is formed in the reference's sequence; all products/sums use explicit round-to-nearest intrinsics (no FMA
// contraction), i.e. the arithmetic of the reference's x86-64 build. The n+1 robot segments go to a per-warp slice of
// shared memory and the (n+1)n/2 self pairs + (n+1) x E environment pairs are tested 32 at a time with the reference's
// float-epsilon thresholds. A motion walks its interior states j/nd, j = 1..nd-1, then s2, in order and stops at the
// first invalid one, which yields both checkMotion forms: the boolean and lastValid.second = (j-1)/nd.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200mp.h"
#include "host_common.hpp"

using namespace b200mp;

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kChainWarps = 8;
constexpr double kPi = 3.141592653589793238462643383279502884;

struct ChainView {
    int n;            // links (<= 31)
    int n_env;        // environment segments
    int n_self;       // (n + 1) n / 2
    double link;      // link length
    double lvs;       // longest valid segment (metric units)
    const double4 *env;      // (x0, y0, x1, y1)
    const uchar2 *self_pairs;  // i < j over the n + 1 robot segments
};

// a*b - c*d without contraction
__device__ __forceinline__ double det2(double a, double b, double c, double d) { return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d)); }

// KinematicChainValidityChecker::intersectionTest (KinematicChain.h:251-276)
__device__ __forceinline__ bool seg_intersect(const double4 s0, const double4 s1) {
    const double s10_x = __dsub_rn(s0.z, s0.x), s10_y = __dsub_rn(s0.w, s0.y);
    const double s32_x = __dsub_rn(s1.z, s1.x), s32_y = __dsub_rn(s1.w, s1.y);
    const double denom = det2(s10_x, s32_y, s32_x, s10_y);
    if (fabs(denom) < DBL_EPSILON) return false;
    const bool pos = denom > 0;
    const double s02_x = __dsub_rn(s0.x, s1.x), s02_y = __dsub_rn(s0.y, s1.y);
    const double s_numer = det2(s10_x, s02_y, s10_y, s02_x);
    if ((s_numer < static_cast<double>(FLT_EPSILON)) == pos) return false;
    const double t_numer = det2(s32_x, s02_y, s32_y, s02_x);
    if ((t_numer < static_cast<double>(FLT_EPSILON)) == pos) return false;
    if (((__dsub_rn(s_numer, denom) > -static_cast<double>(FLT_EPSILON)) == pos) ||
        ((__dsub_rn(t_numer, denom) > static_cast<double>(FLT_EPSILON)) == pos))
        return false;
    return true;
}

// theta_lane = ((q_0 + q_1) + ...) + q_lane, formed in index order (lane i holds q_i; lanes >= n hold 0)
__device__ __forceinline__ double prefix_angle(double q_mine, int n, int lane) {
    double theta = 0.0;
    for (int j = 0; j < n; ++j) {
        const double qj = __shfl_sync(kFullMask, q_mine, j);
        if (j <= lane) theta = __dadd_rn(theta, qj);
    }
    return theta;
}

// Validity of the state whose joint i is held by lane i. seg = this warp's shared-memory slice [32]. Warp-uniform result.
__device__ bool chain_state_valid(const ChainView &c, double q_mine, int lane, double4 *seg) {
    const int n = c.n;
    const double theta = prefix_angle(q_mine, n, lane);
    double cs, sn;
    sincos(theta, &sn, &cs);
    double x = 0.0, y = 0.0;
    double4 mine = make_double4(0, 0, 0, 0);
    for (int j = 0; j < n; ++j) {
        const double cj = __shfl_sync(kFullMask, cs, j), sj = __shfl_sync(kFullMask, sn, j);
        const double xn = __dadd_rn(x, __dmul_rn(cj, c.link)), yn = __dadd_rn(y, __dmul_rn(sj, c.link));
        if (lane == j) mine = make_double4(x, y, xn, yn);
        x = xn;
        y = yn;
    }
    {   // the 1 mm tip segment along the last link (KinematicChain.h:225-227)
        const double cl = __shfl_sync(kFullMask, cs, n - 1), sl = __shfl_sync(kFullMask, sn, n - 1);
        if (lane == n) mine = make_double4(x, y, __dadd_rn(x, __dmul_rn(cl, 0.001)), __dadd_rn(y, __dmul_rn(sl, 0.001)));
    }
    __syncwarp();
    if (lane <= n) seg[lane] = mine;
    __syncwarp();
    bool hit = false;
    for (int t = lane; t < c.n_self; t += 32) {
        const uchar2 p = c.self_pairs[t];
        hit |= seg_intersect(seg[p.x], seg[p.y]);
    }
    if (__any_sync(kFullMask, hit)) return false;
    const int total = (n + 1) * c.n_env;
    for (int t0 = 0; t0 < total; t0 += 256) {
#pragma unroll 1
        for (int t = t0 + lane; t < min(t0 + 256, total); t += 32) {
            const int i = t / c.n_env, e = t - i * c.n_env;
            hit |= seg_intersect(seg[i], c.env[e]);
        }
        if (__any_sync(kFullMask, hit)) return false;
    }
    return true;
}

// KinematicChainSpace::interpolate for one joint (KinematicChain.h:150-176)
__device__ __forceinline__ double interp_joint(double from, double to, double t) {
    double diff = __dsub_rn(to, from);
    if (fabs(diff) <= kPi) return __dadd_rn(from, __dmul_rn(diff, t));
    if (diff > 0.0)
        diff = __dsub_rn(2.0 * kPi, diff);
    else
        diff = __dsub_rn(-2.0 * kPi, diff);
    double v = __dsub_rn(from, __dmul_rn(diff, t));
    if (v > kPi)
        v = __dsub_rn(v, 2.0 * kPi);
    else if (v < -kPi)
        v = __dadd_rn(v, 2.0 * kPi);
    return v;
}

// KinematicChainSpace::distance; lane i holds joint i of both states; every lane returns the full distance
__device__ __forceinline__ double chain_distance(const ChainView &c, double a_mine, double b_mine, int lane) {
    const double t1 = prefix_angle(a_mine, c.n, lane), t2 = prefix_angle(b_mine, c.n, lane);
    double c1, s1, c2, s2;
    sincos(t1, &s1, &c1);
    sincos(t2, &s2, &c2);
    const double dc = __dsub_rn(c1, c2), ds = __dsub_rn(s1, s2);
    double dx = 0.0, dy = 0.0, dist = 0.0;
    for (int j = 0; j < c.n; ++j) {
        dx = __dadd_rn(dx, __shfl_sync(kFullMask, dc, j));
        dy = __dadd_rn(dy, __shfl_sync(kFullMask, ds, j));
        dist = __dadd_rn(dist, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    }
    return __dmul_rn(dist, c.link);
}

__global__ void __launch_bounds__(kChainWarps * 32) chain_states_kernel(ChainView c, const double *__restrict__ q, long long m, uint8_t *__restrict__ valid) {
    __shared__ double4 segs[kChainWarps][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long i = static_cast<long long>(blockIdx.x) * kChainWarps + w; i < m; i += static_cast<long long>(gridDim.x) * kChainWarps) {
        const double qm = lane < c.n ? q[i * c.n + lane] : 0.0;
        const bool ok = chain_state_valid(c, qm, lane, segs[w]);
        if (lane == 0) valid[i] = ok ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kChainWarps * 32)
chain_motions_kernel(ChainView c, const double *__restrict__ a, const double *__restrict__ b, long long m, uint8_t *__restrict__ valid,
                     double *__restrict__ last_t, int *__restrict__ nd_out) {
    __shared__ double4 segs[kChainWarps][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long e = static_cast<long long>(blockIdx.x) * kChainWarps + w; e < m; e += static_cast<long long>(gridDim.x) * kChainWarps) {
        const double am = lane < c.n ? a[e * c.n + lane] : 0.0, bm = lane < c.n ? b[e * c.n + lane] : 0.0;
        const double d = chain_distance(c, am, bm, lane);
        const int nd = static_cast<int>(static_cast<unsigned int>(ceil(__ddiv_rn(d, c.lvs))));
        bool ok = true;
        double lt = -1.0;
        for (int j = 1; j < nd; ++j) {
            const double qm = interp_joint(am, bm, __ddiv_rn(static_cast<double>(j), static_cast<double>(nd)));
            if (!chain_state_valid(c, qm, lane, segs[w])) {
                ok = false;
                lt = __ddiv_rn(static_cast<double>(j - 1), static_cast<double>(nd));
                break;
            }
        }
        if (ok && !chain_state_valid(c, bm, lane, segs[w])) {
            ok = false;
            lt = __ddiv_rn(static_cast<double>(nd - 1), static_cast<double>(nd));
        }
        if (lane == 0) {
            valid[e] = ok ? 1 : 0;
            if (last_t) last_t[e] = lt;
            if (nd_out) nd_out[e] = nd;
        }
    }
}

__global__ void __launch_bounds__(kChainWarps * 32)
chain_distance_kernel(ChainView c, const double *__restrict__ a, const double *__restrict__ b, long long m, double *__restrict__ out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long e = static_cast<long long>(blockIdx.x) * kChainWarps + w; e < m; e += static_cast<long long>(gridDim.x) * kChainWarps) {
        const double am = lane < c.n ? a[e * c.n + lane] : 0.0, bm = lane < c.n ? b[e * c.n + lane] : 0.0;
        const double d = chain_distance(c, am, bm, lane);
        if (lane == 0) out[e] = d;
    }
}

__global__ void chain_interpolate_kernel(int n, const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ t, long long m,
                                         double *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < m * n) out[i] = interp_joint(a[i], b[i], t[i / n]);
}

}  // namespace

struct b200mp_chain {
    b200mp_ctx *ctx = nullptr;
    ChainView view{};
    void *d_mem = nullptr;
};

namespace {

// host buffers -> scratch -> kernel -> host; `fn` launches on the context stream with device pointers
template <typename Launch>
int chain_host_call(b200mp_chain *ch, const double *const *in, const size_t *in_bytes, int n_in, void *const *out, const size_t *out_bytes, int n_out,
                    Launch fn) {
    b200mp_ctx *c = ch->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA(cudaSetDevice(c->device));
    size_t total = 0;
    std::vector<size_t> off;
    for (int i = 0; i < n_in; ++i) { off.push_back(total); total += (in_bytes[i] + 255) / 256 * 256; }
    for (int i = 0; i < n_out; ++i) { off.push_back(total); total += (out_bytes[i] + 255) / 256 * 256; }
    int rc = c->scratch.reserve(total + 256);
    if (rc) return rc;
    std::vector<void *> dp;
    for (size_t o : off) dp.push_back(c->scratch.ptr + o);
    for (int i = 0; i < n_in; ++i) B200_CUDA(cudaMemcpyAsync(dp[i], in[i], in_bytes[i], cudaMemcpyHostToDevice, c->stream));
    fn(dp.data(), c->stream);
    B200_CUDA(cudaGetLastError());
    c->launches.fetch_add(1);
    for (int i = 0; i < n_out; ++i)
        if (out[i]) B200_CUDA(cudaMemcpyAsync(out[i], dp[n_in + i], out_bytes[i], cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return B200MP_OK;
}

inline unsigned chain_grid(const b200mp_chain *ch, long long m) {
    const long long want = (m + kChainWarps - 1) / kChainWarps;
    return static_cast<unsigned>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(ch->ctx->sm_count) * 8)));
}

}  // namespace

extern "C" {

int b200mp_chain_horn_environment(int d, double eps, double *segments, int *n_segments) {
    if (d < 2 || !n_segments) return fail(B200MP_ERR_INVALID, "b200mp_chain_horn_environment: bad argument");
    *n_segments = 2 * (d - 1);
    if (!segments) return B200MP_OK;
    // two polylines that start a small gap apart and curl into a horn (KinematicChain.h:279-315)
    int k = 0;
    for (int side = 0; side < 2; ++side) {
        const double w = 1. / static_cast<double>(d);
        double x = w, y = side == 0 ? -eps : eps, theta = 0.;
        const double scale = side == 0 ? w * (1. + kPi * eps) : w * (1.0 - kPi * eps);
        for (int i = 0; i < d - 1; ++i) {
            theta += kPi / static_cast<double>(d);
            const double xn = x + std::cos(theta) * scale, yn = y + std::sin(theta) * scale;
            segments[4 * k] = x; segments[4 * k + 1] = y; segments[4 * k + 2] = xn; segments[4 * k + 3] = yn;
            ++k;
            x = xn;
            y = yn;
        }
    }
    return B200MP_OK;
}

int b200mp_chain_create(b200mp_ctx *ctx, int n_links, double link_length, const double *segments, int n_seg, double lvs_fraction,
                        b200mp_chain **out) {
    if (!ctx || !out || n_links < 1 || n_seg < 0 || (n_seg > 0 && !segments) || !(lvs_fraction > 0.0))
        return fail(B200MP_ERR_INVALID, "b200mp_chain_create: bad argument");
    if (n_links > 31) return fail(B200MP_ERR_CAPACITY, "b200mp_chain_create: at most 31 links (one lane per segment)");
    B200_CUDA(cudaSetDevice(ctx->device));
    auto *ch = new b200mp_chain();
    ch->ctx = ctx;
    std::vector<unsigned char> pairs;
    for (int i = 0; i <= n_links; ++i)
        for (int j = i + 1; j <= n_links; ++j) { pairs.push_back(static_cast<unsigned char>(i)); pairs.push_back(static_cast<unsigned char>(j)); }
    const size_t env_bytes = (static_cast<size_t>(std::max(n_seg, 1)) * 32 + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&ch->d_mem, env_bytes + pairs.size() + 256);
    if (e == cudaSuccess && n_seg) e = cudaMemcpy(ch->d_mem, segments, static_cast<size_t>(n_seg) * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(static_cast<char *>(ch->d_mem) + env_bytes, pairs.data(), pairs.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (ch->d_mem) cudaFree(ch->d_mem);
        delete ch;
        return fail(B200MP_ERR_CUDA, "b200mp_chain_create: %s", cudaGetErrorString(e));
    }
    ChainView &v = ch->view;
    v.n = n_links;
    v.n_env = n_seg;
    v.n_self = static_cast<int>(pairs.size() / 2);
    v.link = link_length;
    double ext = 0.0;  // RealVectorStateSpace::getMaximumExtent over [-pi, pi]^n (RealVectorStateSpace.cpp:167-176)
    for (int i = 0; i < n_links; ++i) {
        const double d = kPi - (-kPi);
        ext += d * d;
    }
    v.lvs = std::sqrt(ext) * lvs_fraction;  // StateSpace.cpp:236-244
    v.env = static_cast<const double4 *>(ch->d_mem);
    v.self_pairs = reinterpret_cast<const uchar2 *>(static_cast<char *>(ch->d_mem) + env_bytes);
    *out = ch;
    return B200MP_OK;
}

void b200mp_chain_destroy(b200mp_chain *ch) {
    if (!ch) return;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    cudaFree(ch->d_mem);
    delete ch;
}

double b200mp_chain_longest_valid_segment(const b200mp_chain *ch) { return ch ? ch->view.lvs : -1.0; }

int b200mp_chain_validate_states_device(b200mp_chain *ch, const double *d_q, int64_t m, uint8_t *d_valid, void *stream) {
    if (!ch || m < 0 || (m > 0 && (!d_q || !d_valid))) return fail(B200MP_ERR_INVALID, "b200mp_chain_validate_states_device: bad argument");
    if (m == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(ch->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ch->ctx->stream;
    chain_states_kernel<<<chain_grid(ch, m), kChainWarps * 32, 0, st>>>(ch->view, d_q, m, d_valid);
    B200_CUDA(cudaGetLastError());
    ch->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

int b200mp_chain_check_motions_device(b200mp_chain *ch, const double *d_a, const double *d_b, int64_t m, uint8_t *d_valid, double *d_last_t,
                                      int *d_nd, void *stream) {
    if (!ch || m < 0 || (m > 0 && (!d_a || !d_b || !d_valid))) return fail(B200MP_ERR_INVALID, "b200mp_chain_check_motions_device: bad argument");
    if (m == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(ch->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ch->ctx->stream;
    chain_motions_kernel<<<chain_grid(ch, m), kChainWarps * 32, 0, st>>>(ch->view, d_a, d_b, m, d_valid, d_last_t, d_nd);
    B200_CUDA(cudaGetLastError());
    ch->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

int b200mp_chain_validate_states(b200mp_chain *ch, const double *q, int64_t m, uint8_t *valid) {
    if (!ch || m < 0 || (m > 0 && (!q || !valid))) return fail(B200MP_ERR_INVALID, "b200mp_chain_validate_states: bad argument");
    if (m == 0) return B200MP_OK;
    const double *in[1] = {q};
    const size_t ib[1] = {static_cast<size_t>(m) * ch->view.n * 8};
    void *out[1] = {valid};
    const size_t ob[1] = {static_cast<size_t>(m)};
    return chain_host_call(ch, in, ib, 1, out, ob, 1, [&](void **d, cudaStream_t st) {
        chain_states_kernel<<<chain_grid(ch, m), kChainWarps * 32, 0, st>>>(ch->view, static_cast<const double *>(d[0]), m, static_cast<uint8_t *>(d[1]));
    });
}

int b200mp_chain_check_motions(b200mp_chain *ch, const double *a, const double *b, int64_t m, uint8_t *valid, double *last_valid_t, int *nd) {
    if (!ch || m < 0 || (m > 0 && (!a || !b || !valid))) return fail(B200MP_ERR_INVALID, "b200mp_chain_check_motions: bad argument");
    if (m == 0) return B200MP_OK;
    const size_t sb = static_cast<size_t>(m) * ch->view.n * 8;
    const double *in[2] = {a, b};
    const size_t ib[2] = {sb, sb};
    void *out[3] = {valid, last_valid_t, nd};
    const size_t ob[3] = {static_cast<size_t>(m), static_cast<size_t>(m) * 8, static_cast<size_t>(m) * 4};
    return chain_host_call(ch, in, ib, 2, out, ob, 3, [&](void **d, cudaStream_t st) {
        chain_motions_kernel<<<chain_grid(ch, m), kChainWarps * 32, 0, st>>>(ch->view, static_cast<const double *>(d[0]), static_cast<const double *>(d[1]), m,
                                                                            static_cast<uint8_t *>(d[2]), static_cast<double *>(d[3]),
                                                                            static_cast<int *>(d[4]));
    });
}

int b200mp_chain_distance(b200mp_chain *ch, const double *a, const double *b, int64_t m, double *out_d) {
    if (!ch || m < 0 || (m > 0 && (!a || !b || !out_d))) return fail(B200MP_ERR_INVALID, "b200mp_chain_distance: bad argument");
    if (m == 0) return B200MP_OK;
    const size_t sb = static_cast<size_t>(m) * ch->view.n * 8;
    const double *in[2] = {a, b};
    const size_t ib[2] = {sb, sb};
    void *out[1] = {out_d};
    const size_t ob[1] = {static_cast<size_t>(m) * 8};
    return chain_host_call(ch, in, ib, 2, out, ob, 1, [&](void **d, cudaStream_t st) {
        chain_distance_kernel<<<chain_grid(ch, m), kChainWarps * 32, 0, st>>>(ch->view, static_cast<const double *>(d[0]), static_cast<const double *>(d[1]), m,
                                                                             static_cast<double *>(d[2]));
    });
}

int b200mp_chain_interpolate(b200mp_chain *ch, const double *a, const double *b, const double *t, int64_t m, double *out_q) {
    if (!ch || m < 0 || (m > 0 && (!a || !b || !t || !out_q))) return fail(B200MP_ERR_INVALID, "b200mp_chain_interpolate: bad argument");
    if (m == 0) return B200MP_OK;
    const size_t sb = static_cast<size_t>(m) * ch->view.n * 8;
    const double *in[3] = {a, b, t};
    const size_t ib[3] = {sb, sb, static_cast<size_t>(m) * 8};
    void *out[1] = {out_q};
    const size_t ob[1] = {sb};
    const int n = ch->view.n;
    return chain_host_call(ch, in, ib, 3, out, ob, 1, [&](void **d, cudaStream_t st) {
        const long long cnt = static_cast<long long>(m) * n;
        chain_interpolate_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(n, static_cast<const double *>(d[0]), static_cast<const double *>(d[1]),
                                                                                         static_cast<const double *>(d[2]), m, static_cast<double *>(d[3]));
    });
}

}  // extern "C"
