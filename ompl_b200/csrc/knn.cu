// Brute-force / tiled k-nearest-neighbour engine (sm_100a) behind the NearestNeighbors<_T> contract
// (reference: src/ompl/datastructures/NearestNeighbors.h:46-115; ground-truth semantics of
// src/ompl/datastructures/NearestNeighborsLinear.h:65-142; metric src/ompl/base/spaces/src/RealVectorStateSpace.cpp:230-242).
//
// Layout in HBM: the database is kept twice -- exact fp64 rows [N][dim] (what OMPL stores) and an fp32 "feature" row
// per point [N][4*NV] = (c_0 .. c_{dim-1}, 0.., |c|^2) used by the filter. One thread owns one query (its -2q in
// registers); a block streams the feature rows through shared memory in tiles fetched by the bulk async-copy engine
// (cp.async.bulk / UBLKCP, double buffered on mbarriers) and every lane reads the same candidate as a broadcast float4.
// Per pair the hot loop is `dim` FFMAs + one compare: s = |c|^2 - 2 q.c  <  thr, a conservative fp32 filter whose
// threshold carries a rigorous rounding bound. The rare survivors are re-evaluated exactly as the reference does
// (fp64, sum of squared differences in index order) and inserted into the thread's top-k list, so the returned
// neighbours and distances are the reference's doubles, not fp32 approximations. The database is split over
// blockIdx.y; a second kernel merges the partial lists per query with a warp-shuffle arg-min selection.
#include <algorithm>
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
constexpr int kKnnThreads = 256;
constexpr int kTile = 512;       // candidates per shared-memory tile
constexpr int kMaxK = 160;       // list slots per query (k + guard) supported by the tiled kernel

struct Partial {
    double d2;
    int idx;
    int pad;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

// NV float4 per feature row; the last float of the row is |c|^2 (+inf for removed points).
template <int NV>
__global__ void __launch_bounds__(kKnnThreads)
knn_tiled_kernel(const float4 *__restrict__ feat, const double *__restrict__ pts64, long long n, int dim, const double *__restrict__ q64,
                 long long nq, const long long *__restrict__ self_ids, long long self_base, int use_self, int kk, float err_scale,
                 float cmax, Partial *__restrict__ partial, int splits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long bars[2];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);                                        // [2][kTile][NV]
    double *ld2 = reinterpret_cast<double *>(smem_raw + 2 * kTile * NV * sizeof(float4));       // [kk][threads]
    int *lidx = reinterpret_cast<int *>(ld2 + static_cast<size_t>(kk) * kKnnThreads);            // [kk][threads]

    const int tid = threadIdx.x;
    const long long qi = static_cast<long long>(blockIdx.x) * kKnnThreads + tid;
    const bool has_q = qi < nq;
    // database slice of this block
    const long long per = (n + splits - 1) / splits;
    const long long c_begin = min(static_cast<long long>(blockIdx.y) * per, n);
    const long long c_end = min(c_begin + per, n);
    const int n_tiles = static_cast<int>((c_end - c_begin + kTile - 1) / kTile);

    // query registers: m2q = -2 q (fp32), qn = |q|^2
    float m2q[4 * NV];
    float qn = 0.f;
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j) m2q[j] = 0.f;
    double qd[4 * NV - 1];
#pragma unroll
    for (int j = 0; j < 4 * NV - 1; ++j) {
        qd[j] = (has_q && j < dim) ? q64[qi * dim + j] : 0.0;
        const float qf = static_cast<float>(qd[j]);
        m2q[j] = -2.0f * qf;
        qn = fmaf(qf, qf, qn);
    }
    m2q[4 * NV - 1] = 1.0f;  // picks up |c|^2 stored in the last slot
    const long long self = !use_self ? -1 : (self_ids ? (has_q ? self_ids[qi] : -1) : self_base + qi);
    // rounding bound of the fp32 filter: |approx - exact| <= err_scale * (|q| + |c|max)^2
    const float qnorm = sqrtf(qn);
    const float err = err_scale * (qnorm + cmax) * (qnorm + cmax);

    for (int s = 0; s < kk; ++s) {
        ld2[s * kKnnThreads + tid] = INFINITY;
        lidx[s * kKnnThreads + tid] = -1;
    }
    double worst = INFINITY;
    int worst_idx = 0x7fffffff;
    int worst_slot = 0;
    float thr = has_q ? INFINITY : -INFINITY;  // filter threshold on s = |c|^2 - 2 q.c

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long long b = c_begin + static_cast<long long>(t) * kTile;
        const int cnt = static_cast<int>(min(static_cast<long long>(kTile), c_end - b));
        const uint32_t bytes = static_cast<uint32_t>(cnt) * NV * sizeof(float4);
        mbar_expect_tx(&bars[t & 1], bytes);
        bulk_g2s(tiles + static_cast<size_t>(t & 1) * kTile * NV, feat + b * NV, bytes, &bars[t & 1]);
    };
    if (tid == 0 && n_tiles > 0) issue(0);

    for (int t = 0; t < n_tiles; ++t) {
        if (tid == 0 && t + 1 < n_tiles) issue(t + 1);
        mbar_wait(&bars[t & 1], (t >> 1) & 1);
        const float4 *tile = tiles + static_cast<size_t>(t & 1) * kTile * NV;
        const long long base = c_begin + static_cast<long long>(t) * kTile;
        const int cnt = static_cast<int>(min(static_cast<long long>(kTile), c_end - base));
        for (int c0 = 0; c0 < cnt; c0 += 4) {
            float sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 *row = tile + static_cast<size_t>(min(c0 + u, cnt - 1)) * NV;
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 c = row[v];
                    acc = fmaf(m2q[4 * v + 0], c.x, acc);
                    acc = fmaf(m2q[4 * v + 1], c.y, acc);
                    acc = fmaf(m2q[4 * v + 2], c.z, acc);
                    acc = fmaf(m2q[4 * v + 3], c.w, acc);
                }
                sv[u] = acc;
            }
            const bool any = (sv[0] < thr) | (sv[1] < thr) | (sv[2] < thr) | (sv[3] < thr);
            if (any) {
#pragma unroll 1
                for (int u = 0; u < 4; ++u) {
                    if (!(sv[u] < thr) || c0 + u >= cnt) continue;
                    const long long c = base + c0 + u;
                    if (c == self) continue;
                    // exact distance, the reference's arithmetic (RealVectorStateSpace.cpp:230-242)
                    double d2 = 0.0;
#pragma unroll
                    for (int j = 0; j < 4 * NV - 1; ++j) {
                        if (j < dim) {
                            const double diff = pts64[c * dim + j] - qd[j];
                            d2 = __dadd_rn(d2, __dmul_rn(diff, diff));  // no FMA contraction: bit-identical to the reference's x86-64 doubles
                        }
                    }
                    if (d2 < worst || (d2 == worst && static_cast<int>(c) < worst_idx)) {
                        ld2[worst_slot * kKnnThreads + tid] = d2;
                        lidx[worst_slot * kKnnThreads + tid] = static_cast<int>(c);
                        // new worst = lexicographic max of (d2, idx) over the list
                        worst = -1.0;
                        worst_idx = -1;
                        for (int s2 = 0; s2 < kk; ++s2) {
                            const double v = ld2[s2 * kKnnThreads + tid];
                            const int vi = lidx[s2 * kKnnThreads + tid];
                            const int key = vi < 0 ? 0x7fffffff : vi;
                            if (v > worst || (v == worst && key > worst_idx)) { worst = v; worst_idx = key; worst_slot = s2; }
                        }
                        // s_exact = d2 - |q|^2 ; pass while s_approx < worst - qn + err (rounded up)
                        thr = (worst == INFINITY) ? INFINITY : __double2float_ru(worst - static_cast<double>(qn)) + err;
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled two tiles later
    }
    if (has_q) {
        Partial *out = partial + (qi * splits + blockIdx.y) * kk;
        for (int s = 0; s < kk; ++s) {
            Partial p;
            p.d2 = ld2[s * kKnnThreads + tid];
            p.idx = lidx[s * kKnnThreads + tid];
            p.pad = 0;
            out[s] = p;
        }
    }
}

// One warp per query: emit the k smallest (d2, idx) among the splits*kk partial entries in ascending order. Each round
// finds the smallest entry strictly greater than the last one emitted (entries are unique by idx) with a lane-local
// scan followed by a warp-shuffle arg-min, so nothing has to be retired or sorted in memory. dist = sqrt(d2).
__global__ void knn_merge_kernel(const Partial *__restrict__ partial, long long nq, int n_cand, int k, long long *__restrict__ idx,
                                 double *__restrict__ dist, int *__restrict__ counts) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    const Partial *p = partial + w * n_cand;
    constexpr int kPer = 8;  // first 256 candidates live in registers; any further ones are re-read each round
    double d[kPer];
    int id[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int c = lane + 32 * i;
        id[i] = (c < n_cand) ? p[c].idx : -1;
        d[i] = (c < n_cand && id[i] >= 0) ? p[c].d2 : INFINITY;
    }
    double last_d = -1.0;
    int last_i = -1;
    int found = 0;
    for (int r = 0; r < k; ++r) {
        double bd = INFINITY;
        int bi = 0x7fffffff;
        auto consider = [&](double cd, int ci) {
            if (ci < 0) return;
            const bool after = cd > last_d || (cd == last_d && ci > last_i);
            const bool better = cd < bd || (cd == bd && ci < bi);
            if (after && better) { bd = cd; bi = ci; }
        };
#pragma unroll
        for (int i = 0; i < kPer; ++i) consider(d[i], id[i]);
        for (int c = lane + 32 * kPer; c < n_cand; c += 32) consider(p[c].d2, p[c].idx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(kFullMask, bd, o);
            const int oi = __shfl_xor_sync(kFullMask, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (bi == 0x7fffffff) break;
        if (lane == 0) {
            idx[w * k + r] = bi;
            dist[w * k + r] = sqrt(bd);
        }
        last_d = bd;
        last_i = bi;
        ++found;
    }
    if (lane == 0) {
        for (int r = found; r < k; ++r) { idx[w * k + r] = -1; dist[w * k + r] = INFINITY; }
        if (counts) counts[w] = found;
    }
}

// feature rows from fp64 points: (c_0..c_{dim-1}, 0.., |c|^2)
__global__ void knn_features_kernel(const double *__restrict__ pts, long long first, long long count, int dim, int fw, float *__restrict__ feat) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double *p = pts + (first + i) * dim;
    float *f = feat + (first + i) * fw;
    float nrm = 0.f;
    for (int j = 0; j < fw - 1; ++j) {
        const float v = j < dim ? static_cast<float>(p[j]) : 0.f;
        f[j] = v;
        nrm = fmaf(v, v, nrm);
    }
    f[fw - 1] = nrm;
}

__global__ void knn_f32_to_f64_kernel(const float *__restrict__ src, long long count, double *__restrict__ dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = static_cast<double>(src[i]);
}

// exact distances of one query to every point (nearestR / k >= n paths)
__global__ void knn_all_distances_kernel(const double *__restrict__ pts, const float *__restrict__ feat, int fw, long long n, int dim,
                                         const double *__restrict__ q, double *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d2 = 0.0;
    for (int j = 0; j < dim; ++j) {
        const double diff = pts[i * dim + j] - q[j];
        d2 = __dadd_rn(d2, __dmul_rn(diff, diff));  // no FMA contraction: bit-identical to the reference's x86-64 doubles
    }
    out[i] = isinf(feat[i * fw + fw - 1]) ? INFINITY : sqrt(d2);  // removed points report +inf
}

}  // namespace

struct b200mp_knn {
    b200mp_ctx *ctx = nullptr;
    int dim = 0;
    int nv = 0;  // float4 per feature row
    int metric = 0;
    long long n = 0, cap = 0, n_removed = 0;
    double *d_pts = nullptr;
    float *d_feat = nullptr;
    float cmax = 0.f;  // largest |c| seen (filter error bound)
    std::vector<double> h_pts;      // host mirror (nearestR / list / small paths)
    std::vector<unsigned char> alive;
    Scratch scratch;
    Scratch qscratch;  // fp64 copies of fp32 query blocks
};

namespace {

int knn_reserve(b200mp_knn *k, long long want) {
    if (want <= k->cap) return B200MP_OK;
    long long ncap = std::max<long long>(want, std::max<long long>(1024, k->cap * 2));
    double *np = nullptr;
    float *nf = nullptr;
    B200_CUDA(cudaMalloc(&np, static_cast<size_t>(ncap) * k->dim * sizeof(double)));
    B200_CUDA(cudaMalloc(&nf, static_cast<size_t>(ncap) * k->nv * 4 * sizeof(float)));
    if (k->n > 0) {
        B200_CUDA(cudaMemcpyAsync(np, k->d_pts, static_cast<size_t>(k->n) * k->dim * sizeof(double), cudaMemcpyDeviceToDevice, k->ctx->stream));
        B200_CUDA(cudaMemcpyAsync(nf, k->d_feat, static_cast<size_t>(k->n) * k->nv * 4 * sizeof(float), cudaMemcpyDeviceToDevice, k->ctx->stream));
        B200_CUDA(cudaStreamSynchronize(k->ctx->stream));
    }
    if (k->d_pts) cudaFree(k->d_pts);
    if (k->d_feat) cudaFree(k->d_feat);
    k->d_pts = np;
    k->d_feat = nf;
    k->cap = ncap;
    return B200MP_OK;
}

int knn_build_features(b200mp_knn *k, long long first, long long count, cudaStream_t st) {
    if (count <= 0) return B200MP_OK;
    const int thr = 256;
    knn_features_kernel<<<static_cast<unsigned>((count + thr - 1) / thr), thr, 0, st>>>(k->d_pts, first, count, k->dim, k->nv * 4, k->d_feat);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

template <int NV>
int launch_tiled(b200mp_knn *k, const double *d_q, long long nq, const long long *d_self, long long self_base, int use_self, int kk,
                 Partial *d_partial, int splits, cudaStream_t st) {
    const size_t smem = 2 * kTile * NV * sizeof(float4) + static_cast<size_t>(kk) * kKnnThreads * (sizeof(double) + sizeof(int));
    auto kern = knn_tiled_kernel<NV>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(static_cast<unsigned>((nq + kKnnThreads - 1) / kKnnThreads), static_cast<unsigned>(splits));
    // |approx - exact| of an fp32 dot product of <= 4*NV terms plus the stored norm: (4*NV + 2) ulp-level terms, x4 margin
    const float err_scale = 4.0f * (4 * NV + 2) * 5.96e-8f;
    kern<<<grid, kKnnThreads, smem, st>>>(reinterpret_cast<const float4 *>(k->d_feat), k->d_pts, k->n, k->dim, d_q, nq, d_self, self_base,
                                         use_self, kk, err_scale, k->cmax, d_partial, splits);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

// exact distance of one query to every point, filtered by radius, sorted by (distance, id); caller holds ctx->mu
int knn_radius_impl(b200mp_knn *k, const double *q, double radius, int64_t self, std::vector<std::pair<double, int64_t>> &sel) {
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = k->ctx->stream;
    int rc = k->ctx->scratch.reserve(static_cast<size_t>(k->n) * sizeof(double) + k->dim * sizeof(double) + 256);
    if (rc) return rc;
    double *d_all = reinterpret_cast<double *>(k->ctx->scratch.ptr);
    double *d_q = d_all + k->n;
    B200_CUDA(cudaMemcpyAsync(d_q, q, k->dim * sizeof(double), cudaMemcpyHostToDevice, st));
    knn_all_distances_kernel<<<static_cast<unsigned>((k->n + 255) / 256), 256, 0, st>>>(k->d_pts, k->d_feat, k->nv * 4, k->n, k->dim, d_q, d_all);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    std::vector<double> all(static_cast<size_t>(k->n));
    B200_CUDA(cudaMemcpyAsync(all.data(), d_all, static_cast<size_t>(k->n) * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    sel.clear();
    for (int64_t i = 0; i < k->n; ++i)
        if (all[i] <= radius && !std::isinf(all[i]) && i != self) sel.emplace_back(all[i], i);
    std::sort(sel.begin(), sel.end());
    return B200MP_OK;
}

}  // namespace

extern "C" {

int b200mp_knn_create(b200mp_ctx *ctx, int dim, int metric, b200mp_knn **out) {
    if (!ctx || !out || dim < 1) return fail(B200MP_ERR_INVALID, "b200mp_knn_create: bad argument");
    if (metric != B200MP_METRIC_L2) return fail(B200MP_ERR_INVALID, "b200mp_knn_create: unknown metric %d", metric);
    if (dim > 23) return fail(B200MP_ERR_CAPACITY, "b200mp_knn_create: dim %d > 23 not supported by the tiled kernel", dim);
    auto *k = new b200mp_knn();
    k->ctx = ctx;
    k->dim = dim;
    k->metric = metric;
    k->nv = dim <= 7 ? 2 : (dim <= 11 ? 3 : 6);
    *out = k;
    return B200MP_OK;
}

void b200mp_knn_destroy(b200mp_knn *k) {
    if (!k) return;
    cudaSetDevice(k->ctx->device);
    cudaStreamSynchronize(k->ctx->stream);
    if (k->d_pts) cudaFree(k->d_pts);
    if (k->d_feat) cudaFree(k->d_feat);
    k->scratch.release();
    k->qscratch.release();
    delete k;
}

int b200mp_knn_clear(b200mp_knn *k) {
    if (!k) return fail(B200MP_ERR_INVALID, "b200mp_knn_clear: NULL");
    k->n = 0;
    k->n_removed = 0;
    k->cmax = 0.f;
    k->h_pts.clear();
    k->alive.clear();
    return B200MP_OK;
}

int64_t b200mp_knn_size(const b200mp_knn *k) { return k ? k->n - k->n_removed : -1; }

int b200mp_knn_add(b200mp_knn *k, const double *pts, int64_t n, int64_t *first_id) {
    if (!k || (n > 0 && !pts) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_knn_add: bad argument");
    if (first_id) *first_id = k->n;
    if (n == 0) return B200MP_OK;
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    B200_CUDA(cudaSetDevice(k->ctx->device));
    int rc = knn_reserve(k, k->n + n);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(k->d_pts + k->n * k->dim, pts, static_cast<size_t>(n) * k->dim * sizeof(double), cudaMemcpyHostToDevice,
                              k->ctx->stream));
    if ((rc = knn_build_features(k, k->n, n, k->ctx->stream))) return rc;
    k->h_pts.insert(k->h_pts.end(), pts, pts + n * k->dim);
    k->alive.insert(k->alive.end(), static_cast<size_t>(n), 1);
    for (int64_t i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < k->dim; ++j) s += pts[i * k->dim + j] * pts[i * k->dim + j];
        k->cmax = std::max(k->cmax, static_cast<float>(std::sqrt(s)) * 1.0001f);
    }
    k->n += n;
    B200_CUDA(cudaStreamSynchronize(k->ctx->stream));
    return B200MP_OK;
}

int b200mp_knn_add_device_f32(b200mp_knn *k, const float *d_pts, int64_t n, float max_norm, void *stream) {
    if (!k || (n > 0 && !d_pts) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_knn_add_device_f32: bad argument");
    if (n == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : k->ctx->stream;
    int rc = knn_reserve(k, k->n + n);
    if (rc) return rc;
    const long long cnt = static_cast<long long>(n) * k->dim;
    knn_f32_to_f64_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(d_pts, cnt, k->d_pts + k->n * k->dim);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    if ((rc = knn_build_features(k, k->n, n, st))) return rc;
    k->cmax = std::max(k->cmax, max_norm * 1.0001f);
    k->h_pts.clear();  // host mirror not maintained for device-resident databases
    k->alive.clear();
    k->n += n;
    return B200MP_OK;
}

int b200mp_knn_remove(b200mp_knn *k, int64_t id) {
    if (!k) return fail(B200MP_ERR_INVALID, "b200mp_knn_remove: NULL");
    if (id < 0 || id >= k->n || k->alive.empty() || !k->alive[id]) return fail(B200MP_ERR_INVALID, "b200mp_knn_remove: id %lld not present", (long long)id);
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    B200_CUDA(cudaSetDevice(k->ctx->device));
    const float inf = INFINITY;  // tombstone: |c|^2 = +inf never passes the filter
    B200_CUDA(cudaMemcpy(k->d_feat + id * k->nv * 4 + (k->nv * 4 - 1), &inf, sizeof(float), cudaMemcpyHostToDevice));
    k->alive[id] = 0;
    k->n_removed++;
    return B200MP_OK;
}

int b200mp_knn_query_device(b200mp_knn *k, const double *d_q, int64_t nq, int kq, const int64_t *d_self_ids, int64_t self_base,
                            int exclude_self, int64_t *d_idx, double *d_dist, int *d_counts, void *stream) {
    if (!k || nq < 0 || kq < 1 || (nq > 0 && (!d_q || !d_idx || !d_dist))) return fail(B200MP_ERR_INVALID, "b200mp_knn_query_device: bad argument");
    if (nq == 0) return B200MP_OK;
    if (kq > kMaxK) return fail(B200MP_ERR_CAPACITY, "b200mp_knn_query_device: k %d > %d (use b200mp_knn_radius / host query for larger k)", kq, kMaxK);
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : k->ctx->stream;
    const int kk = kq;  // survivors are evaluated exactly, so no guard band is needed
    // split the database so that at least ~2 waves of blocks exist even for a single query
    const long long qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
    long long splits = std::max<long long>(1, (2LL * k->ctx->sm_count + qblocks - 1) / qblocks);
    splits = std::min<long long>(splits, std::max<long long>(1, (k->n + kTile - 1) / kTile));
    splits = std::min<long long>(splits, 1024);
    // chunk queries so the partial buffer stays bounded (<= 1 GiB)
    const long long max_q = std::max<long long>(kKnnThreads, (1LL << 30) / (static_cast<long long>(sizeof(Partial)) * kk * splits) / kKnnThreads * kKnnThreads);
    for (long long q0 = 0; q0 < nq; q0 += max_q) {
        const long long cnt = std::min<long long>(max_q, nq - q0);
        int rc = k->scratch.reserve(static_cast<size_t>(cnt) * splits * kk * sizeof(Partial));
        if (rc) return rc;
        Partial *d_partial = reinterpret_cast<Partial *>(k->scratch.ptr);
        const long long *selfp = d_self_ids ? reinterpret_cast<const long long *>(d_self_ids) + q0 : nullptr;
        switch (k->nv) {
            case 2: rc = launch_tiled<2>(k, d_q + q0 * k->dim, cnt, selfp, self_base + q0, exclude_self, kk, d_partial, (int)splits, st); break;
            case 3: rc = launch_tiled<3>(k, d_q + q0 * k->dim, cnt, selfp, self_base + q0, exclude_self, kk, d_partial, (int)splits, st); break;
            default: rc = launch_tiled<6>(k, d_q + q0 * k->dim, cnt, selfp, self_base + q0, exclude_self, kk, d_partial, (int)splits, st); break;
        }
        if (rc) return rc;
        const int thr = 256;
        const long long warps = cnt;
        knn_merge_kernel<<<static_cast<unsigned>((warps * 32 + thr - 1) / thr), thr, 0, st>>>(
            d_partial, cnt, static_cast<int>(splits) * kk, kq, reinterpret_cast<long long *>(d_idx) + q0 * kq, d_dist + q0 * kq,
            d_counts ? d_counts + q0 : nullptr);
        B200_CUDA(cudaGetLastError());
        k->ctx->launches.fetch_add(1);
    }
    return B200MP_OK;
}

int b200mp_knn_query_device_f32(b200mp_knn *k, const float *d_q, int64_t nq, int kq, int64_t self_base, int exclude_self, int64_t *d_idx,
                                double *d_dist, int *d_counts, void *stream) {
    if (!k || nq < 0 || (nq > 0 && !d_q)) return fail(B200MP_ERR_INVALID, "b200mp_knn_query_device_f32: bad argument");
    if (nq == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : k->ctx->stream;
    const long long cnt = static_cast<long long>(nq) * k->dim;
    int rc = k->qscratch.reserve(static_cast<size_t>(cnt) * sizeof(double));
    if (rc) return rc;
    double *d_q64 = reinterpret_cast<double *>(k->qscratch.ptr);
    knn_f32_to_f64_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(d_q, cnt, d_q64);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return b200mp_knn_query_device(k, d_q64, nq, kq, nullptr, self_base, exclude_self, d_idx, d_dist, d_counts, st);
}

int b200mp_knn_query(b200mp_knn *k, const double *q, int64_t nq, int kq, const int64_t *self_ids, int64_t *idx, double *dist, int *counts) {
    if (!k || nq < 0 || kq < 0 || (nq > 0 && (!q || (kq > 0 && (!idx || !dist))))) return fail(B200MP_ERR_INVALID, "b200mp_knn_query: bad argument");
    if (nq == 0 || kq == 0) {
        if (counts) for (int64_t i = 0; i < nq; ++i) counts[i] = 0;
        return B200MP_OK;
    }
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    if (kq > kMaxK) {
        // large k (RRT* asks for k ~ k_rrt log n, often >= size): exhaustive sorted list per query
        std::vector<std::pair<double, int64_t>> sel;
        for (int64_t i = 0; i < nq; ++i) {
            int rc = B200MP_OK;
            sel.clear();
            if (k->n > 0) rc = knn_radius_impl(k, q + i * k->dim, INFINITY, self_ids ? self_ids[i] : -1, sel);
            if (rc) return rc;
            const int64_t c = std::min<int64_t>(kq, static_cast<int64_t>(sel.size()));
            for (int64_t t = 0; t < kq; ++t) {
                idx[i * kq + t] = t < c ? sel[t].second : -1;
                dist[i * kq + t] = t < c ? sel[t].first : INFINITY;
            }
            if (counts) counts[i] = static_cast<int>(c);
        }
        return B200MP_OK;
    }
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = k->ctx->stream;
    const size_t qb = static_cast<size_t>(nq) * k->dim * sizeof(double);
    const size_t ib = static_cast<size_t>(nq) * kq * sizeof(int64_t);
    const size_t db = static_cast<size_t>(nq) * kq * sizeof(double);
    const size_t sb = self_ids ? static_cast<size_t>(nq) * sizeof(int64_t) : 0;
    const size_t cb = static_cast<size_t>(nq) * sizeof(int);
    int rc = k->ctx->scratch.reserve(qb + ib + db + sb + cb + 1024);
    if (rc) return rc;
    char *p = k->ctx->scratch.ptr;
    double *d_q = reinterpret_cast<double *>(p); p += qb;
    int64_t *d_idx = reinterpret_cast<int64_t *>(p); p += ib;
    double *d_dist = reinterpret_cast<double *>(p); p += db;
    int64_t *d_self = self_ids ? reinterpret_cast<int64_t *>(p) : nullptr; p += sb;
    int *d_cnt = reinterpret_cast<int *>(p);
    B200_CUDA(cudaMemcpyAsync(d_q, q, qb, cudaMemcpyHostToDevice, st));
    if (self_ids) B200_CUDA(cudaMemcpyAsync(d_self, self_ids, sb, cudaMemcpyHostToDevice, st));
    rc = b200mp_knn_query_device(k, d_q, nq, kq, d_self, 0, self_ids ? 1 : 0, d_idx, d_dist, d_cnt, st);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(idx, d_idx, ib, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(dist, d_dist, db, cudaMemcpyDeviceToHost, st));
    if (counts) B200_CUDA(cudaMemcpyAsync(counts, d_cnt, cb, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200MP_OK;
}

// nearestR and exhaustive sorted queries (k >= size): exact distances to every point on the GPU, selection and
// ordering of the (usually short) result on the host. Returns the number of neighbours in *count; writes at most cap.
int b200mp_knn_radius(b200mp_knn *k, const double *q, double radius, int64_t *idx, double *dist, int64_t cap, int64_t *count) {
    if (!k || !q || !count || cap < 0 || (cap > 0 && (!idx || !dist))) return fail(B200MP_ERR_INVALID, "b200mp_knn_radius: bad argument");
    *count = 0;
    if (k->n == 0) return B200MP_OK;
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    std::vector<std::pair<double, int64_t>> sel;
    int rc = knn_radius_impl(k, q, radius, -1, sel);
    if (rc) return rc;
    *count = static_cast<int64_t>(sel.size());
    for (int64_t t = 0; t < std::min<int64_t>(cap, *count); ++t) {
        idx[t] = sel[t].second;
        dist[t] = sel[t].first;
    }
    return B200MP_OK;
}

}  // extern "C"
