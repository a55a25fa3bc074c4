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
#include <cstdlib>
#include <cstring>
#include <map>
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
                // rows past cnt (same buffer, stale or unwritten) may be read: whatever they yield is dropped below by index
                const float4 *row = tile + (c0 + u) * NV;
                float acc = 0.f;  // s = |c|^2 - 2 q.c: the stored norm starts the sum
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 c = row[NV - 1 - v];
                    if (v == 0) acc = c.w;
                    else acc = fmaf(m2q[4 * (NV - 1 - v) + 3], c.w, acc);
                    acc = fmaf(m2q[4 * (NV - 1 - v) + 0], c.x, acc);
                    acc = fmaf(m2q[4 * (NV - 1 - v) + 1], c.y, acc);
                    acc = fmaf(m2q[4 * (NV - 1 - v) + 2], c.z, acc);
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
                                 double *__restrict__ dist, int *__restrict__ counts, int take_sqrt, const int *__restrict__ qpart = nullptr,
                                 int parts = 1, int part = 0) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    if (qpart != nullptr && qpart[w] != part) return;  // row belongs to another part: left untouched
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
            dist[w * k + r] = take_sqrt ? sqrt(bd) : bd;
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


// ================================================================================================ cell index
// Exact search with spatial pruning for large batches (roadmap construction: every state queries the whole set): a balanced k-d
// partition of the database built on the device, rows stored leaf by leaf, tight boxes per leaf; the search itself is the
// query-major kernel further down. Pruning uses lower bounds only, so the result is the same exact (distance, id)-ordered top-k
// as the exhaustive scan.
// ---- partition: a balanced k-d tree of depth L = log2(P) built level by level on the device. Every level splits all of its
// nodes along one dimension (the one whose current cell extent is largest) at the approximate median of the node's rows:
// a 256-bin histogram over the node's extent along that dimension, split at the upper edge of the bin where the count
// passes one half. Leaves are axis-aligned boxes with ~n / P rows each; the pruning boxes are then taken from the rows
// themselves (knn_cell_bounds_kernel), so they are at least as tight as the leaves. Any partition gives the same result
// (pruning uses lower bounds only); this one is what makes the boxes small: the Voronoi cells of random pivots used before
// have bounding boxes so fat in 7-D that every block visited every cell.
constexpr int kKdBins = 256;
constexpr int kKdMaxDim = 24;

__device__ __forceinline__ unsigned kd_fkey(float f) {  // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float kd_unfkey(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void knn_kd_extent_init_kernel(unsigned *__restrict__ ext, int dim) {
    const int j = threadIdx.x;
    if (j < dim) { ext[2 * j] = 0xffffffffu; ext[2 * j + 1] = 0u; }
}

// ext[2 * j] = min key, ext[2 * j + 1] = max key of coordinate j over the rows (grid-stride; one pair of atomics per block and coordinate)
__global__ void __launch_bounds__(256) knn_kd_extent_kernel(const float *__restrict__ rows, long long n, int fw, int dim, unsigned *__restrict__ ext) {
    __shared__ unsigned s_lo[kKdMaxDim], s_hi[kKdMaxDim];
    if (threadIdx.x < kKdMaxDim) { s_lo[threadIdx.x] = 0xffffffffu; s_hi[threadIdx.x] = 0u; }
    __syncthreads();
    for (int j = 0; j < dim; ++j) {
        unsigned lo = 0xffffffffu, hi = 0u;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const float v = rows[i * fw + j];
            if (isfinite(v)) { const unsigned key = kd_fkey(v); lo = min(lo, key); hi = max(hi, key); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(kFullMask, lo, o));
            hi = max(hi, __shfl_xor_sync(kFullMask, hi, o));
        }
        if ((threadIdx.x & 31) == 0 && lo <= hi) {
            atomicMin(&s_lo[j], lo);
            atomicMax(&s_hi[j], hi);
        }
    }
    __syncthreads();
    if (threadIdx.x < dim && s_lo[threadIdx.x] <= s_hi[threadIdx.x]) {
        atomicMin(ext + 2 * threadIdx.x, s_lo[threadIdx.x]);
        atomicMax(ext + 2 * threadIdx.x + 1, s_hi[threadIdx.x]);
    }
}

// one thread: split dimension of every level (greedy: the dimension whose cell extent is currently largest) and the root box
__global__ void knn_kd_plan_kernel(const unsigned *__restrict__ ext, int dim, int L, int *__restrict__ dims, float *__restrict__ root_box) {
    float w[kKdMaxDim];
    for (int j = 0; j < dim; ++j) {
        const float lo = kd_unfkey(ext[2 * j]), hi = kd_unfkey(ext[2 * j + 1]);
        root_box[2 * j] = lo;
        root_box[2 * j + 1] = hi;
        w[j] = hi - lo;
    }
    for (int l = 0; l < L; ++l) {
        int best = 0;
        for (int j = 1; j < dim; ++j)
            if (w[j] > w[best]) best = j;
        dims[l] = best;
        w[best] *= 0.5f;
    }
}

// histogram of level l: node[i] is the row's node inside the level (0 .. 2^l - 1); boxes[node][2 * dim] are the level's cells
__global__ void knn_kd_hist_kernel(const float *__restrict__ rows, long long n, int fw, int dim, const int *__restrict__ node, const int *__restrict__ dims,
                                   int l, const float *__restrict__ boxes, unsigned *__restrict__ hist) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nd = node[i], d = dims[l];
    const float lo = boxes[(static_cast<size_t>(nd) * dim + d) * 2], hi = boxes[(static_cast<size_t>(nd) * dim + d) * 2 + 1];
    const float x = rows[i * fw + d];
    int bin = 0;
    if (hi > lo) bin = min(max(static_cast<int>((x - lo) / (hi - lo) * kKdBins), 0), kKdBins - 1);
    atomicAdd(hist + static_cast<size_t>(nd) * kKdBins + bin, 1u);
}

// the same for the shallow levels (<= 32 nodes: the whole histogram fits shared memory): grid-stride blocks count in shared memory and
// flush once -- a million rows hammering a few hundred global counters is what made these levels the slowest
constexpr int kKdSmemLevels = 5;
__global__ void __launch_bounds__(512) knn_kd_hist_smem_kernel(const float *__restrict__ rows, long long n, int fw, int dim, const int *__restrict__ node,
                                                               const int *__restrict__ dims, int l, const float *__restrict__ boxes,
                                                               unsigned *__restrict__ hist) {
    __shared__ unsigned s_h[(1 << kKdSmemLevels) * kKdBins];
    const int cells = (1 << l) * kKdBins;
    for (int t = threadIdx.x; t < cells; t += blockDim.x) s_h[t] = 0u;
    __syncthreads();
    const int d = dims[l];
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int nd = node[i];
        const float lo = boxes[(static_cast<size_t>(nd) * dim + d) * 2], hi = boxes[(static_cast<size_t>(nd) * dim + d) * 2 + 1];
        const float x = rows[i * fw + d];
        int bin = 0;
        if (hi > lo) bin = min(max(static_cast<int>((x - lo) / (hi - lo) * kKdBins), 0), kKdBins - 1);
        atomicAdd(&s_h[nd * kKdBins + bin], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cells; t += blockDim.x)
        if (s_h[t]) atomicAdd(hist + t, s_h[t]);
}

// one thread per node of level l: split value + the two child boxes (level l + 1)
__global__ void knn_kd_split_kernel(const unsigned *__restrict__ hist, int l, int dim, const int *__restrict__ dims, const float *__restrict__ boxes,
                                    float *__restrict__ child_boxes, float *__restrict__ splits) {
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= (1 << l)) return;
    const int d = dims[l];
    const unsigned *h = hist + static_cast<size_t>(nd) * kKdBins;
    unsigned total = 0;
    for (int b = 0; b < kKdBins; ++b) total += h[b];
    const float lo = boxes[(static_cast<size_t>(nd) * dim + d) * 2], hi = boxes[(static_cast<size_t>(nd) * dim + d) * 2 + 1];
    int cut = kKdBins / 2;
    if (total > 0) {
        unsigned acc = 0;
        for (int b = 0; b < kKdBins; ++b) {
            acc += h[b];
            if (2u * acc >= total) { cut = b + 1; break; }
        }
    }
    const float split = hi > lo ? lo + (hi - lo) * (static_cast<float>(cut) / kKdBins) : lo;
    splits[(1 << l) - 1 + nd] = split;
    for (int j = 0; j < dim; ++j) {
        const float bl = boxes[(static_cast<size_t>(nd) * dim + j) * 2], bh = boxes[(static_cast<size_t>(nd) * dim + j) * 2 + 1];
        float *c0 = child_boxes + (static_cast<size_t>(2 * nd) * dim + j) * 2, *c1 = child_boxes + (static_cast<size_t>(2 * nd + 1) * dim + j) * 2;
        c0[0] = bl; c0[1] = j == d ? split : bh;
        c1[0] = j == d ? split : bl; c1[1] = bh;
    }
}

// node[i] <- child of node[i] at level l (right when x >= split)
__global__ void knn_kd_descend_kernel(const float *__restrict__ rows, long long n, int fw, int *__restrict__ node, const int *__restrict__ dims, int l,
                                      const float *__restrict__ splits) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nd = node[i];
    node[i] = 2 * nd + (rows[i * fw + dims[l]] >= splits[(1 << l) - 1 + nd] ? 1 : 0);
}

// leaf of every row (queries, or the database itself): L comparisons down the split planes
__global__ void knn_kd_assign_kernel(const float *__restrict__ rows, long long n, int fw, int L, const int *__restrict__ dims,
                                     const float *__restrict__ splits, int *__restrict__ cell) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int nd = 0;
    for (int l = 0; l < L; ++l) nd = 2 * nd + (rows[i * fw + dims[l]] >= splits[(1 << l) - 1 + nd] ? 1 : 0);
    cell[i] = nd;
}

__global__ void knn_cell_count_kernel(const int *__restrict__ cell, long long n, int *__restrict__ counts) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(counts + cell[i], 1);
}

// one block: counts[P] -> start[P + 1] (exclusive), cursor[P] = start
__global__ void __launch_bounds__(1024) knn_cell_scan_kernel(const int *__restrict__ counts, int P, int *__restrict__ start, int *__restrict__ cursor) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < P; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < P ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFullMask, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int ex = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
        if (i < P) { start[i] = ex; cursor[i] = ex; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) start[P] = carry_s;
}

__global__ void knn_cell_slot_kernel(const int *__restrict__ cell, long long n, int *__restrict__ cursor, int *__restrict__ pos) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = atomicAdd(cursor + cell[i], 1);
}

__global__ void knn_permute_kernel(const float4 *__restrict__ feat, const double *__restrict__ pts, const int *__restrict__ pos, long long n, int dim,
                                   int nv, float4 *__restrict__ feat_s, double *__restrict__ pts_s, int *__restrict__ orig) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long p = pos[i];
    for (int v = 0; v < nv; ++v) feat_s[p * nv + v] = feat[i * nv + v];
    for (int j = 0; j < dim; ++j) pts_s[p * dim + j] = pts[i * dim + j];
    orig[p] = static_cast<int>(i);
}

__global__ void knn_query_order_kernel(const int *__restrict__ cell, const int *__restrict__ pos, long long nq, int *__restrict__ order,
                                       int *__restrict__ cell_sorted) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    order[pos[i]] = static_cast<int>(i);
    cell_sorted[pos[i]] = cell[i];
}

// one warp per cell: bounding box of its rows (tombstoned rows skipped), widened by a few ulps so that it also bounds the
// exact fp64 points the fp32 rows were rounded from
__global__ void knn_cell_bounds_kernel(const float *__restrict__ feat_s, const int *__restrict__ start, int P, int fw, int dim,
                                       float *__restrict__ lo, float *__restrict__ hi) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= P) return;
    for (int j = 0; j < dim; ++j) {
        float mn = INFINITY, mx = -INFINITY;
        for (int r = start[c] + lane; r < start[c + 1]; r += 32) {
            if (isinf(feat_s[static_cast<long long>(r) * fw + fw - 1])) continue;
            const float v = feat_s[static_cast<long long>(r) * fw + j];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(kFullMask, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
        }
        if (lane == 0) {
            const bool empty = mn > mx;  // no live rows: a box at infinite distance from everything
            // (1e-6 relative: 4e-7 for the fp32 rounding of the rows, the rest for the fp32 rounding of a query tested against the box)
            lo[c * 24 + j] = empty ? INFINITY : mn - (fabsf(mn) * 1e-6f + 1e-30f);
            hi[c * 24 + j] = empty ? -INFINITY : mx + (fabsf(mx) * 1e-6f + 1e-30f);
        }
    }
}


// ================================================================================================ query-major cell search
// Box-to-box pruning (above) is weak in 7-D: 2^14 leaves are only ~4 cells per dimension, every leaf TOUCHES ~3^7 others, so a
// block (or warp) of queries that shares one bound and one box still evaluates 8-15 % of all pairs, while a single query's ball
// reaches a few dozen leaves. Here the unit of work is a batch of QB queries adjacent in k-d order, worked on by ONE WARP with the
// roles of the lanes turned around:
//   * search: lanes span CELLS. Three levels of boxes (256-leaf subtrees, 16-leaf subtrees, leaves; consecutive in heap order) are
//     tested 32 at a time: the upper levels against the batch's box, the leaves against each query POINT (point-to-box distance vs
//     that query's current k-th distance);
//   * evaluation: lanes span candidate ROWS of a visited leaf (coalesced LDG.128 of the k-d ordered feature rows), every row is
//     filtered against the QB queries (fp32, rigorous error bound, as in the kernels above); survivors are re-evaluated by their lane
//     in the reference's fp64 operation order;
//   * the k-list of a query is spread over the lanes (entry i in lane i & 31); the worst entry is kept warp-uniform and found again
//     after an insertion with three warp reductions (REDUX) on the bit pattern of the non-negative doubles;
//   * nearest first: ring s visits the leaves whose smallest point gap^2 lies in (tau_{s-1}, tau_s]; ring 0 is the queries' own
//     leaves, later rings grow geometrically while the bounds shrink.
// Results are written sorted, straight to the caller's tables (no partial lists, no merge kernel). Work per query is proportional
// to the leaves its own ball reaches, so pairs evaluated fall from 14.5 % of N^2 to well under 1 %.
constexpr int kSuper = 16;           // leaves per super-cell
constexpr int kMega = 256;           // leaves per mega-cell
constexpr int kPointThreads = 256;
constexpr int kF2Cap = 32 * kSuper;  // super-cells of one round of 32 mega-cells

// box4[(c * 2 + 0) * NV + v] = lower corner, [(c * 2 + 1) * NV + v] = upper corner of leaf c, padded coordinates 0; sbox4 / mbox4 the
// same for the unions over groups of kSuper / kMega leaves. Empty boxes are (+inf, -inf).
__global__ void knn_pack_boxes_kernel(const float *__restrict__ lo, const float *__restrict__ hi, int P, int dim, int fw, int group, int write_leaves,
                                      float *__restrict__ box, float *__restrict__ gbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (group, coordinate)
    const int G = (P + group - 1) / group;
    if (i >= G * fw) return;
    const int s = i / fw, j = i % fw;
    float slo = INFINITY, shi = -INFINITY;
    for (int t = 0; t < group; ++t) {
        const int c = s * group + t;
        if (c >= P) break;
        const float l = j < dim ? lo[c * 24 + j] : 0.f, h = j < dim ? hi[c * 24 + j] : 0.f;
        if (write_leaves) {
            box[(static_cast<size_t>(c) * 2 + 0) * fw + j] = l;
            box[(static_cast<size_t>(c) * 2 + 1) * fw + j] = h;
        }
        slo = fminf(slo, l);
        shi = fmaxf(shi, h);
    }
    gbox[(static_cast<size_t>(s) * 2 + 0) * fw + j] = slo;
    gbox[(static_cast<size_t>(s) * 2 + 1) * fw + j] = shi;
}

// Share of part `part` of `parts` in the k-d ordered queries: whole leaves, contiguous, balanced by query count. Depends only on the
// leaves' query counts, so every GPU holding the same database and the same queries computes the same shares.
__device__ __forceinline__ void knn_part_range(const int *__restrict__ q_start, int P, long long nq, int parts, int part, int *lo, int *hi) {
    auto first_leaf_at = [&](long long target) {  // first leaf whose queries start at or after `target`
        int a = 0, b = P;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (q_start[mid] < target) a = mid + 1; else b = mid;
        }
        return a;
    };
    *lo = part == 0 ? 0 : q_start[first_leaf_at(nq * part / parts)];
    *hi = part == parts - 1 ? static_cast<int>(nq) : q_start[first_leaf_at(nq * (part + 1) / parts)];
}

// KPL list entries per lane (k <= 32 KPL); QB queries per batch
template <int NV, int QB, int KPL, int MINB>
__global__ void __launch_bounds__(kPointThreads, MINB)
knn_point_kernel(const float4 *__restrict__ feat_s, const double *__restrict__ pts_s, const int *__restrict__ orig, const int *__restrict__ cell_start,
                 const float4 *__restrict__ box4, const float4 *__restrict__ sbox4, const float4 *__restrict__ mbox4, int P, int dim,
                 const double *__restrict__ q64, const int *__restrict__ q_order, const int *__restrict__ q_cell, const int *__restrict__ q_start,
                 long long nq, const long long *__restrict__ self_ids, long long self_base, int use_self, int kk, float err_scale, float cmax,
                 float tau0_frac, float tau_growth, int parts, int part, int chunk, long long *__restrict__ out_idx, double *__restrict__ out_dist,
                 int *__restrict__ out_idx32, int *__restrict__ out_counts, unsigned long long *__restrict__ work_head, unsigned long long *__restrict__ stats) {
    constexpr int W = kPointThreads / 32;
    constexpr int FW = 4 * NV;
    __shared__ unsigned short s_f2[W][kF2Cap];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int S = (P + kSuper - 1) / kSuper, M = (P + kMega - 1) / kMega;
    int p_lo = 0, p_hi = static_cast<int>(nq);
    if (parts > 1) knn_part_range(q_start, P, nq, parts, part, &p_lo, &p_hi);
    unsigned long long st_rows = 0, st_exact = 0, st_ins = 0, st_cells = 0;

    for (;;) {
        unsigned long long w0 = 0;
        if (lane == 0) w0 = atomicAdd(work_head, static_cast<unsigned long long>(chunk));
        w0 = __shfl_sync(kFullMask, w0, 0);
        if (static_cast<long long>(w0) >= p_hi - p_lo) break;
        const int c_end = static_cast<int>(min(static_cast<long long>(p_hi), static_cast<long long>(p_lo) + static_cast<long long>(w0) + chunk));
        for (int p0 = p_lo + static_cast<int>(w0); p0 < c_end; p0 += QB) {
            // ---- the batch: QB queries adjacent in k-d order (all lanes hold all of them)
            bool has_q[QB];
            long long qi[QB];
            int self[QB], own[QB];
            float m2q[QB][FW], qn[QB], err[QB], thr[QB], rw2[QB];  // query b's coordinates are -0.5 * m2q[b][.] (exact)
            double wd2[QB];            // worst entry of the list: value, key, holder (warp-uniform)
            int wkey[QB], wlane[QB], wslot[QB];
            double ld2[QB][KPL];       // this lane's entries of the QB lists
            int lid[QB][KPL];          // id, -1 = empty, -2 = beyond k (never the worst, never reported)
            float blo[FW], bhi[FW];    // box of the batch (upper levels of the search)
#pragma unroll
            for (int j = 0; j < FW; ++j) { blo[j] = INFINITY; bhi[j] = -INFINITY; }
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                has_q[b] = p0 + b < c_end;
                qi[b] = q_order[has_q[b] ? p0 + b : p0];
                own[b] = q_cell[has_q[b] ? p0 + b : p0];
                qn[b] = 0.f;
#pragma unroll
                for (int j = 0; j < FW - 1; ++j) {
                    const float v = j < dim ? static_cast<float>(q64[qi[b] * dim + j]) : 0.f;
                    m2q[b][j] = -2.0f * v;
                    qn[b] = fmaf(v, v, qn[b]);
                    blo[j] = fminf(blo[j], v);
                    bhi[j] = fmaxf(bhi[j], v);
                }
                m2q[b][FW - 1] = 1.0f;  // picks up |c|^2 stored in the last slot
                const long long sf = !use_self ? -1 : (self_ids ? self_ids[qi[b]] : self_base + qi[b]);
                self[b] = static_cast<int>(sf);
                const float qnorm = sqrtf(qn[b]);
                err[b] = err_scale * (qnorm + cmax) * (qnorm + cmax);
                thr[b] = has_q[b] ? INFINITY : -INFINITY;
                rw2[b] = has_q[b] ? INFINITY : -INFINITY;
#pragma unroll
                for (int e = 0; e < KPL; ++e) {
                    const bool real = e * 32 + lane < kk;
                    ld2[b][e] = real ? INFINITY : 0.0;
                    lid[b][e] = real ? -1 : -2;
                }
                wd2[b] = INFINITY;
                wkey[b] = 0x7fffffff;
                wlane[b] = 0;
                wslot[b] = 0;
            }
#pragma unroll
            for (int j = 0; j < FW; ++j) {  // widened like the cell boxes: the box also bounds the exact fp64 queries
                blo[j] = j < FW - 1 ? blo[j] - (fabsf(blo[j]) * 4e-7f + 1e-30f) : 0.f;
                bhi[j] = j < FW - 1 ? bhi[j] + (fabsf(bhi[j]) * 4e-7f + 1e-30f) : 0.f;
            }
            // squared distance from the batch's box to a box (upper levels) / from query b to a box (leaves)
            auto box_gap2 = [&](const float4 *__restrict__ bx) -> float {
                float g2 = 0.f;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 lo = __ldg(bx + v), hi = __ldg(bx + NV + v);
                    float g;
                    g = fmaxf(fmaxf(lo.x - bhi[4 * v + 0], blo[4 * v + 0] - hi.x), 0.f); g2 = fmaf(g, g, g2);
                    g = fmaxf(fmaxf(lo.y - bhi[4 * v + 1], blo[4 * v + 1] - hi.y), 0.f); g2 = fmaf(g, g, g2);
                    g = fmaxf(fmaxf(lo.z - bhi[4 * v + 2], blo[4 * v + 2] - hi.z), 0.f); g2 = fmaf(g, g, g2);
                    if (v < NV - 1) { g = fmaxf(fmaxf(lo.w - bhi[4 * v + 3], blo[4 * v + 3] - hi.w), 0.f); g2 = fmaf(g, g, g2); }
                }
                return g2;
            };
            // list update after the bound of query b changed
            auto refresh = [&](int b) {
                thr[b] = (wd2[b] == INFINITY) ? INFINITY : __double2float_ru(wd2[b] - static_cast<double>(qn[b])) + err[b];
                rw2[b] = (wd2[b] == INFINITY) ? INFINITY : __double2float_ru(wd2[b]) * 1.000001f;
            };
            // lexicographic maximum of (d2, key) over the list of query b: three reductions on the bits of the non-negative doubles
            auto find_worst = [&](int b) {
                unsigned hi_u = 0, lo_u = 0, key_u = 0;
                int slot = 0;
#pragma unroll
                for (int e = 0; e < KPL; ++e) {
                    const unsigned h = static_cast<unsigned>(__double2hiint(ld2[b][e])), l = static_cast<unsigned>(__double2loint(ld2[b][e]));
                    // keys: real id + 1; empty = above every id (distinct per entry); beyond-k entries = 0 (with d2 = 0: the minimum)
                    const unsigned ky = lid[b][e] >= 0 ? static_cast<unsigned>(lid[b][e]) + 1u
                                                       : (lid[b][e] == -1 ? 0xffffffffu - static_cast<unsigned>(e * 32 + lane) : 0u);
                    if (h > hi_u || (h == hi_u && (l > lo_u || (l == lo_u && ky >= key_u)))) { hi_u = h; lo_u = l; key_u = ky; slot = e; }
                }
                const unsigned mh = __reduce_max_sync(kFullMask, hi_u);
                bool cand = hi_u == mh;
                const unsigned ml = __reduce_max_sync(kFullMask, cand ? lo_u : 0u);
                cand = cand && lo_u == ml;
                const unsigned mk = __reduce_max_sync(kFullMask, cand ? key_u : 0u);
                const unsigned who = __ballot_sync(kFullMask, cand && key_u == mk);
                wlane[b] = __ffs(who) - 1;
                wslot[b] = __shfl_sync(kFullMask, slot, wlane[b]);
                wd2[b] = __hiloint2double(static_cast<int>(mh), static_cast<int>(ml));
                wkey[b] = mk >= 0x80000000u ? 0x7fffffff : static_cast<int>(mk) - 1;  // an empty entry ranks above every id
                refresh(b);
            };
            // rows [rb, re) of the k-d ordered database against the batch, 64 at a time (two per lane, all loads issued first)
            auto eval_rows = [&](int rb, int re) {
                for (int r0 = rb; r0 < re; r0 += 64) {
                    int r[2];
                    bool in[2];
                    float4 c[2][NV];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        r[h] = r0 + 32 * h + lane;
                        in[h] = r[h] < re;
                        const float4 *row = feat_s + static_cast<long long>(in[h] ? r[h] : re - 1) * NV;
#pragma unroll
                        for (int v = 0; v < NV; ++v) c[h][v] = __ldg(row + v);
                    }
                    bool pass[2][QB];
                    bool any = false;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int b = 0; b < QB; ++b) {
                            float sv = c[h][NV - 1].w;  // s = |c|^2 - 2 q.c: the stored norm starts the sum
#pragma unroll
                            for (int v = 0; v < NV; ++v) {
                                sv = fmaf(m2q[b][4 * v + 0], c[h][v].x, sv);
                                sv = fmaf(m2q[b][4 * v + 1], c[h][v].y, sv);
                                sv = fmaf(m2q[b][4 * v + 2], c[h][v].z, sv);
                                if (v < NV - 1) sv = fmaf(m2q[b][4 * v + 3], c[h][v].w, sv);
                            }
                            pass[h][b] = in[h] && sv < thr[b];
                            any |= pass[h][b];
                        }
                    }
                    if (!__any_sync(kFullMask, any)) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        bool anyh = false;
#pragma unroll
                        for (int b = 0; b < QB; ++b) anyh |= pass[h][b];
                        if (!__any_sync(kFullMask, anyh)) continue;
                        const int rr = r[h];
                        const int id = anyh ? __ldg(orig + rr) : -1;
                        const bool causal = use_self == 2;  // only points inserted before the query (ids below its own)
#pragma unroll
                        for (int b = 0; b < QB; ++b) {
                            bool cand = pass[h][b];
                            if (!__any_sync(kFullMask, cand)) continue;
                            // exact distance, the reference's arithmetic (RealVectorStateSpace.cpp:230-242), by the lane that holds the row
                            double d2 = 0.0;
                            if (cand && (causal ? id < self[b] : id != self[b])) {
                                for (int j = 0; j < dim; ++j) {
                                    const double diff = pts_s[static_cast<long long>(rr) * dim + j] - q64[qi[b] * dim + j];
                                    d2 = __dadd_rn(d2, __dmul_rn(diff, diff));  // no FMA contraction: bit-identical to the reference's x86-64 doubles
                                }
                            } else {
                                cand = false;
                            }
                            if (stats) st_exact += __popc(__ballot_sync(kFullMask, cand));
                            for (;;) {  // insert the candidates one at a time, each against the bound the previous one left
                                const bool better = cand && (d2 < wd2[b] || (d2 == wd2[b] && id < wkey[b]));
                                const unsigned m = __ballot_sync(kFullMask, better);
                                if (!m) break;
                                const int src = __ffs(m) - 1;
                                const double nd = __shfl_sync(kFullMask, d2, src);
                                const int ni = __shfl_sync(kFullMask, id, src);
                                if (lane == wlane[b]) {
#pragma unroll
                                    for (int e = 0; e < KPL; ++e)
                                        if (e == wslot[b]) { ld2[b][e] = nd; lid[b][e] = ni; }
                                }
                                if (lane == src) cand = false;
                                find_worst(b);
                                if (stats) ++st_ins;
                            }
                        }
                    }
                }
                if (stats) { st_rows += static_cast<unsigned long long>(re - rb) * QB; ++st_cells; }
            };
            // ---- ring 0: the leaves the queries fall into
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                bool first = has_q[b];
#pragma unroll
                for (int b2 = 0; b2 < b; ++b2) first = first && !(has_q[b2] && own[b2] == own[b]);
                if (first) eval_rows(__ldg(cell_start + own[b]), __ldg(cell_start + own[b] + 1));
            }
            // ---- rings: leaves whose smallest point gap^2 lies in (tau_prev, tau], nearest first
            auto sweep = [&](float tau_prev, float tau) {
                float rmax = -INFINITY;
#pragma unroll
                for (int b = 0; b < QB; ++b) rmax = fmaxf(rmax, rw2[b]);
                for (int mb = 0; mb < M; mb += 32) {
                    const int mg = mb + lane;
                    const float g = mg < M ? box_gap2(mbox4 + static_cast<size_t>(mg) * 2 * NV) : INFINITY;
                    unsigned mm = __ballot_sync(kFullMask, g <= tau && g * 0.99999f <= rmax);
                    int nf2 = 0;
                    while (mm) {  // super-cells of two passing mega-cells at a time, one lane each
                        const int m1 = mb + __ffs(mm) - 1;
                        mm &= mm - 1;
                        int m2 = -1;
                        if (mm) { m2 = mb + __ffs(mm) - 1; mm &= mm - 1; }
                        const int mym = lane < 16 ? m1 : m2;
                        const int sc = mym * (kMega / kSuper) + (lane & 15);
                        const float sg = (mym >= 0 && sc < S) ? box_gap2(sbox4 + static_cast<size_t>(sc) * 2 * NV) : INFINITY;
                        const bool ok = sg <= tau && sg * 0.99999f <= rmax;
                        const unsigned sm = __ballot_sync(kFullMask, ok);
                        if (ok) s_f2[wid][nf2 + __popc(sm & ((1u << lane) - 1u))] = static_cast<unsigned short>(sc);
                        nf2 += __popc(sm);
                    }
                    __syncwarp();
                    for (int i = 0; i < nf2; i += 2) {  // leaves of two passing super-cells at a time, one lane each
                        const int s1 = s_f2[wid][i], s2 = i + 1 < nf2 ? static_cast<int>(s_f2[wid][i + 1]) : -1;
                        const int mys = lane < 16 ? s1 : s2;
                        const int c = mys * kSuper + (lane & 15);
                        bool take = false;
                        int rb = 0, re = 0;
                        if (mys >= 0 && c < P) {
                            rb = __ldg(cell_start + c);  // row range of the leaf, fetched by its lane alongside the box
                            re = __ldg(cell_start + c + 1);
                            const float4 *bx = box4 + static_cast<size_t>(c) * 2 * NV;
                            float4 lo[NV], hi[NV];
#pragma unroll
                            for (int v = 0; v < NV; ++v) { lo[v] = __ldg(bx + v); hi[v] = __ldg(bx + NV + v); }
                            float gmin = INFINITY;
                            bool reach = false;
#pragma unroll
                            for (int b = 0; b < QB; ++b) {
                                float g2 = 0.f;
#pragma unroll
                                for (int v = 0; v < NV; ++v) {
                                    const float qc0 = -0.5f * m2q[b][4 * v + 0], qc1 = -0.5f * m2q[b][4 * v + 1], qc2 = -0.5f * m2q[b][4 * v + 2];
                                    const float qc3 = v < NV - 1 ? -0.5f * m2q[b][4 * v + 3] : 0.f;
                                    float d;
                                    d = fmaxf(fmaxf(lo[v].x - qc0, qc0 - hi[v].x), 0.f); g2 = fmaf(d, d, g2);
                                    d = fmaxf(fmaxf(lo[v].y - qc1, qc1 - hi[v].y), 0.f); g2 = fmaf(d, d, g2);
                                    d = fmaxf(fmaxf(lo[v].z - qc2, qc2 - hi[v].z), 0.f); g2 = fmaf(d, d, g2);
                                    if (v < NV - 1) { d = fmaxf(fmaxf(lo[v].w - qc3, qc3 - hi[v].w), 0.f); g2 = fmaf(d, d, g2); }
                                }
                                if (has_q[b]) {
                                    gmin = fminf(gmin, g2);
                                    // 0.9999: the fp32 gap of the fp32-rounded query stays below the true gap of the fp64 query
                                    reach = reach || g2 * 0.9999f <= rw2[b];
                                    take = take || c == own[b];
                                }
                            }
                            take = !take && reach && gmin > tau_prev && gmin <= tau;  // own leaves were ring 0
                        }
                        unsigned cm = __ballot_sync(kFullMask, take);
                        while (cm) {
                            const int l = __ffs(cm) - 1;
                            cm &= cm - 1;
                            eval_rows(__shfl_sync(kFullMask, rb, l), __shfl_sync(kFullMask, re, l));
                        }
                    }
                    __syncwarp();
                }
            };
            // first radius when the own leaves hold fewer than k rows (no finite bound yet): diagonal of the own leaf's box, at least
            float seed_tau = 1e-12f;
            {
                const float4 *bx = box4 + static_cast<size_t>(own[0]) * 2 * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 lo = __ldg(bx + v), hi = __ldg(bx + NV + v);
                    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z, dw = v < NV - 1 ? hi.w - lo.w : 0.f;
                    if (dx > 0.f && dx < INFINITY) seed_tau = fmaf(dx, dx, seed_tau);
                    if (dy > 0.f && dy < INFINITY) seed_tau = fmaf(dy, dy, seed_tau);
                    if (dz > 0.f && dz < INFINITY) seed_tau = fmaf(dz, dz, seed_tau);
                    if (dw > 0.f && dw < INFINITY) seed_tau = fmaf(dw, dw, seed_tau);
                }
            }
            float tau_prev = -1.f, tau = 0.f;
            for (int it = 0;; ++it) {
                float rmax = -INFINITY;
#pragma unroll
                for (int b = 0; b < QB; ++b) rmax = fmaxf(rmax, rw2[b]);
                if (it > 0 && tau * 0.9999f >= rmax) break;  // every leaf within the final bounds has been visited
                tau_prev = it == 0 ? -1.f : tau;
                tau = it == 0 ? (rmax < INFINITY ? tau0_frac * rmax : seed_tau) : tau * tau_growth;
                if (it > 48) tau = INFINITY;
                sweep(tau_prev, tau);
            }
            // ---- sorted output: rank of every entry = number of entries before it in (d2, id) order
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                if (!has_q[b]) continue;
                int rank[KPL];
                unsigned mykey[KPL];
#pragma unroll
                for (int e = 0; e < KPL; ++e) {
                    rank[e] = 0;
                    mykey[e] = lid[b][e] >= 0 ? static_cast<unsigned>(lid[b][e]) : 0xffffffffu - static_cast<unsigned>(e * 32 + lane);
                }
                for (int s2 = 0; s2 < kk; ++s2) {
                    double od = 0.0;
                    unsigned ok = 0;
#pragma unroll
                    for (int e = 0; e < KPL; ++e)
                        if (e == s2 / 32) { od = __shfl_sync(kFullMask, ld2[b][e], s2 & 31); ok = __shfl_sync(kFullMask, mykey[e], s2 & 31); }
#pragma unroll
                    for (int e = 0; e < KPL; ++e) rank[e] += (od < ld2[b][e] || (od == ld2[b][e] && ok < mykey[e])) ? 1 : 0;
                }
                int found = 0;
#pragma unroll
                for (int e = 0; e < KPL; ++e) {
                    const bool real = lid[b][e] != -2;
                    if (real) {
                        if (out_idx) out_idx[qi[b] * kk + rank[e]] = lid[b][e];
                        if (out_dist) out_dist[qi[b] * kk + rank[e]] = lid[b][e] >= 0 ? sqrt(ld2[b][e]) : INFINITY;
                        if (out_idx32) out_idx32[qi[b] * kk + rank[e]] = lid[b][e];
                    }
                    found += __popc(__ballot_sync(kFullMask, real && lid[b][e] >= 0));
                }
                if (out_counts && lane == 0) out_counts[qi[b]] = found;
            }
        }
    }
    if (stats && lane == 0) {
        atomicAdd(stats + 0, st_rows);
        atomicAdd(stats + 1, st_exact);
        atomicAdd(stats + 2, st_ins);
        atomicAdd(stats + 3, st_cells);
    }
}


// ================================================================================================ chain metric
// KinematicChainSpace::distance (demos/KinematicChain.h:105-121): L * sum_i |p_i(a) - p_i(b)| over the joint end points
// p_i = (sum_{j<=i} cos theta_j, sum_{j<=i} sin theta_j), theta_j the cumulative angle. Per point two tables are kept:
// fp32 cumulative end points [2 NL] for the filter and fp64 per-link (cos, sin) [2 n] for the exact re-evaluation, which
// accumulates dx += c1 - c2, dy += s1 - s2, dist += sqrt(dx^2 + dy^2) in the reference's order without contraction.
__global__ void knn_chain_tables_kernel(const double *__restrict__ pts, long long first, long long count, int n, int nl, float *__restrict__ cum32,
                                        double *__restrict__ link64) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double *p = pts + (first + i) * n;
    float *f = cum32 + (first + i) * 2 * nl;
    double *t = link64 + (first + i) * 2 * n;
    double theta = 0.0, cx = 0.0, sy = 0.0;
    for (int j = 0; j < nl; ++j) {
        if (j < n) {
            theta = __dadd_rn(theta, p[j]);
            double c, s;
            sincos(theta, &s, &c);
            t[2 * j] = c;
            t[2 * j + 1] = s;
            cx += c;
            sy += s;
        }
        f[2 * j] = static_cast<float>(cx);      // padding links repeat the last end point: they add 0 to every distance
        f[2 * j + 1] = static_cast<float>(sy);
    }
}

constexpr int kChainTile = 128;

template <int NL>
__global__ void __launch_bounds__(kKnnThreads)
knn_chain_kernel(const float2 *__restrict__ cum32, const double *__restrict__ link64, const unsigned char *__restrict__ dead, long long n, int nlinks,
                 double link_len, const float2 *__restrict__ qcum32, const double *__restrict__ qlink64, long long nq,
                 const long long *__restrict__ self_ids, long long self_base, int use_self, int kk, Partial *__restrict__ partial, int splits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long bars[2];
    float2 *tiles = reinterpret_cast<float2 *>(smem_raw);                                          // [2][kChainTile][NL]
    double *ld = reinterpret_cast<double *>(smem_raw + 2 * kChainTile * NL * sizeof(float2));     // [kk][threads] distances
    int *lidx = reinterpret_cast<int *>(ld + static_cast<size_t>(kk) * kKnnThreads);
    const int tid = threadIdx.x;
    const long long qi = static_cast<long long>(blockIdx.x) * kKnnThreads + tid;
    const bool has_q = qi < nq;
    const long long per = (n + splits - 1) / splits;
    const long long c_begin = min(static_cast<long long>(blockIdx.y) * per, n);
    const long long c_end = min(c_begin + per, n);
    const int n_tiles = static_cast<int>((c_end - c_begin + kChainTile - 1) / kChainTile);
    float2 q[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) q[j] = has_q ? qcum32[qi * NL + j] : make_float2(0.f, 0.f);
    const long long self = !use_self ? -1 : (self_ids ? (has_q ? self_ids[qi] : -1) : self_base + qi);
    // fp32 filter error, in units of link lengths: every end-point coordinate (|.| <= n) carries <= 1 ulp(n) of rounding on
    // both sides and the accumulation a few more; 16 ulp(n) per term is a comfortable bound
    const float err = static_cast<float>(nlinks) * 16.0f * 1.1920929e-7f * static_cast<float>(nlinks);
    for (int s = 0; s < kk; ++s) {
        ld[s * kKnnThreads + tid] = INFINITY;
        lidx[s * kKnnThreads + tid] = -1;
    }
    double worst = INFINITY;
    int worst_idx = 0x7fffffff, worst_slot = 0;
    float thr = has_q ? INFINITY : -INFINITY;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long long b = c_begin + static_cast<long long>(t) * kChainTile;
        const int cnt = static_cast<int>(min(static_cast<long long>(kChainTile), c_end - b));
        const uint32_t bytes = static_cast<uint32_t>(cnt) * NL * sizeof(float2);
        mbar_expect_tx(&bars[t & 1], bytes);
        bulk_g2s(tiles + static_cast<size_t>(t & 1) * kChainTile * NL, cum32 + b * NL, bytes, &bars[t & 1]);
    };
    if (tid == 0 && n_tiles > 0) issue(0);
    for (int t = 0; t < n_tiles; ++t) {
        if (tid == 0 && t + 1 < n_tiles) issue(t + 1);
        mbar_wait(&bars[t & 1], (t >> 1) & 1);
        const float2 *tile = tiles + static_cast<size_t>(t & 1) * kChainTile * NL;
        const long long base = c_begin + static_cast<long long>(t) * kChainTile;
        const int cnt = static_cast<int>(min(static_cast<long long>(kChainTile), c_end - base));
        for (int c0 = 0; c0 < cnt; ++c0) {
            const float2 *row = tile + static_cast<size_t>(c0) * NL;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const float2 c = row[j];
                const float dx = c.x - q[j].x, dy = c.y - q[j].y;
                acc += sqrtf(fmaf(dy, dy, dx * dx));
            }
            if (!(acc < thr)) continue;
            const long long c = base + c0;
            if (c == self || dead[c]) continue;
            double dx = 0.0, dy = 0.0, dist = 0.0;
            for (int j = 0; j < nlinks; ++j) {
                dx = __dadd_rn(dx, __dsub_rn(link64[c * 2 * nlinks + 2 * j], qlink64[qi * 2 * nlinks + 2 * j]));
                dy = __dadd_rn(dy, __dsub_rn(link64[c * 2 * nlinks + 2 * j + 1], qlink64[qi * 2 * nlinks + 2 * j + 1]));
                dist = __dadd_rn(dist, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
            }
            const double d = __dmul_rn(dist, link_len);
            if (d < worst || (d == worst && static_cast<int>(c) < worst_idx)) {
                ld[worst_slot * kKnnThreads + tid] = d;
                lidx[worst_slot * kKnnThreads + tid] = static_cast<int>(c);
                worst = -1.0;
                worst_idx = -1;
                for (int s2 = 0; s2 < kk; ++s2) {
                    const double v = ld[s2 * kKnnThreads + tid];
                    const int vi = lidx[s2 * kKnnThreads + tid];
                    const int key = vi < 0 ? 0x7fffffff : vi;
                    if (v > worst || (v == worst && key > worst_idx)) { worst = v; worst_idx = key; worst_slot = s2; }
                }
                thr = (worst == INFINITY) ? INFINITY : __double2float_ru(worst / link_len) * 1.000001f + err;
            }
        }
        __syncthreads();
    }
    if (has_q) {
        Partial *out = partial + (qi * splits + blockIdx.y) * kk;
        for (int s = 0; s < kk; ++s) {
            Partial p;
            p.d2 = ld[s * kKnnThreads + tid];
            p.idx = lidx[s * kKnnThreads + tid];
            p.pad = 0;
            out[s] = p;
        }
    }
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
    double link_len = 1.0;          // chain metric: link length
    int nl = 0;                     // chain metric: links padded to the kernel's template size
    float *d_cum32 = nullptr;       // chain metric: fp32 cumulative end points [cap][2 nl]
    double *d_link64 = nullptr;     // chain metric: fp64 per-link (cos, sin) [cap][2 dim]
    unsigned char *d_dead = nullptr;  // chain metric: tombstones [cap]
    std::vector<double> h_pts;      // host mirror (nearestR / list / small paths)
    std::vector<unsigned char> alive;
    Scratch scratch;
    Scratch qscratch;  // fp64 copies of fp32 query blocks
    // cell index over rows [0, index_n): rebuilt lazily when the database changed since it was built
    bool use_index = true;
    bool index_valid = false;
    long long index_n = 0;
    int index_P = 0;
    Scratch index_mem;   // feat_sorted | pts_sorted | orig | cell_start | lo | hi
    Scratch index_tmp;   // build / query-ordering temporaries
    float4 *ix_feat = nullptr;
    double *ix_pts = nullptr;
    int *ix_orig = nullptr, *ix_start = nullptr;
    float *ix_lo = nullptr, *ix_hi = nullptr, *ix_splits = nullptr;
    float4 *ix_box = nullptr, *ix_sbox = nullptr, *ix_mbox = nullptr;  // packed boxes of the leaves and of their 16- and 256-leaf subtrees
    unsigned long long *d_point_ctl = nullptr;  // [0] work-queue head, [1..4] statistics of the query-major search
    int *ix_dims = nullptr;
    int index_L = 0;
    unsigned long long *d_rows_visited = nullptr;  // candidate rows streamed by the cell kernel since creation
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
    if (k->metric == B200MP_METRIC_CHAIN) {
        float *nc = nullptr;
        double *nl64 = nullptr;
        unsigned char *nd = nullptr;
        B200_CUDA(cudaMalloc(&nc, static_cast<size_t>(ncap) * 2 * k->nl * sizeof(float)));
        B200_CUDA(cudaMalloc(&nl64, static_cast<size_t>(ncap) * 2 * k->dim * sizeof(double)));
        B200_CUDA(cudaMalloc(&nd, static_cast<size_t>(ncap)));
        B200_CUDA(cudaMemsetAsync(nd, 0, static_cast<size_t>(ncap), k->ctx->stream));
        if (k->n > 0) {
            B200_CUDA(cudaMemcpyAsync(nc, k->d_cum32, static_cast<size_t>(k->n) * 2 * k->nl * sizeof(float), cudaMemcpyDeviceToDevice, k->ctx->stream));
            B200_CUDA(cudaMemcpyAsync(nl64, k->d_link64, static_cast<size_t>(k->n) * 2 * k->dim * sizeof(double), cudaMemcpyDeviceToDevice, k->ctx->stream));
            B200_CUDA(cudaMemcpyAsync(nd, k->d_dead, static_cast<size_t>(k->n), cudaMemcpyDeviceToDevice, k->ctx->stream));
        }
        B200_CUDA(cudaStreamSynchronize(k->ctx->stream));
        if (k->d_cum32) cudaFree(k->d_cum32);
        if (k->d_link64) cudaFree(k->d_link64);
        if (k->d_dead) cudaFree(k->d_dead);
        k->d_cum32 = nc;
        k->d_link64 = nl64;
        k->d_dead = nd;
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
    if (k->metric == B200MP_METRIC_CHAIN) {
        knn_chain_tables_kernel<<<static_cast<unsigned>((count + 127) / 128), 128, 0, st>>>(k->d_pts, first, count, k->dim, k->nl, k->d_cum32, k->d_link64);
        B200_CUDA(cudaGetLastError());
        k->ctx->launches.fetch_add(1);
        return B200MP_OK;
    }
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
    // largest dynamic size this kernel has been opted in for ON THIS DEVICE (one attribute call per growth, not per query)
    static std::mutex opt_mu;
    static std::map<int, size_t> opted;
    {
        std::lock_guard<std::mutex> lk(opt_mu);
        if (smem > opted[k->ctx->device]) {
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            opted[k->ctx->device] = smem;
        }
    }
    dim3 grid(static_cast<unsigned>((nq + kKnnThreads - 1) / kKnnThreads), static_cast<unsigned>(splits));
    // |approx - exact| of an fp32 dot product of <= 4*NV terms plus the stored norm: (4*NV + 2) ulp-level terms, x4 margin
    const float err_scale = 4.0f * (4 * NV + 2) * 5.96e-8f;
    kern<<<grid, kKnnThreads, smem, st>>>(reinterpret_cast<const float4 *>(k->d_feat), k->d_pts, k->n, k->dim, d_q, nq, d_self, self_base,
                                         use_self, kk, err_scale, k->cmax, d_partial, splits);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

__global__ void knn_chain_all_distances_kernel(const double *__restrict__ link64, const unsigned char *__restrict__ dead, long long n, int nlinks,
                                               double link_len, const double *__restrict__ qlink64, double *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dx = 0.0, dy = 0.0, dist = 0.0;
    for (int j = 0; j < nlinks; ++j) {
        dx = __dadd_rn(dx, __dsub_rn(link64[i * 2 * nlinks + 2 * j], qlink64[2 * j]));
        dy = __dadd_rn(dy, __dsub_rn(link64[i * 2 * nlinks + 2 * j + 1], qlink64[2 * j + 1]));
        dist = __dadd_rn(dist, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    }
    out[i] = dead[i] ? INFINITY : __dmul_rn(dist, link_len);
}

// exact distance of one host query to every point (removed points: +inf) into the context's scratch buffer; caller holds ctx->mu
int knn_all_distances(b200mp_knn *k, const double *q, double **d_all_out, size_t extra_bytes = 0) {
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = k->ctx->stream;
    const size_t head = (static_cast<size_t>(k->n) * sizeof(double) + static_cast<size_t>(k->dim) * 3 * sizeof(double) +
                         static_cast<size_t>(k->nl) * 2 * sizeof(float) + 512 + 255) / 256 * 256;
    int rc = k->ctx->scratch.reserve(head + extra_bytes);
    if (rc) return rc;
    double *d_all = reinterpret_cast<double *>(k->ctx->scratch.ptr);
    double *d_q = d_all + k->n;
    B200_CUDA(cudaMemcpyAsync(d_q, q, k->dim * sizeof(double), cudaMemcpyHostToDevice, st));
    if (k->metric == B200MP_METRIC_CHAIN) {
        double *d_ql = d_q + k->dim;
        float *d_qc = reinterpret_cast<float *>(d_ql + 2 * k->dim);
        knn_chain_tables_kernel<<<1, 128, 0, st>>>(d_q, 0, 1, k->dim, k->nl, d_qc, d_ql);
        knn_chain_all_distances_kernel<<<static_cast<unsigned>((k->n + 255) / 256), 256, 0, st>>>(k->d_link64, k->d_dead, k->n, k->dim, k->link_len, d_ql, d_all);
        k->ctx->launches.fetch_add(1);
    } else
    knn_all_distances_kernel<<<static_cast<unsigned>((k->n + 255) / 256), 256, 0, st>>>(k->d_pts, k->d_feat, k->nv * 4, k->n, k->dim, d_q, d_all);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    *d_all_out = d_all;
    return B200MP_OK;
}

// ---- large k on the device: radix select of the k-th smallest distance, then a sort of what lies at or below it -------------------
// (RRT* asks nearestK with k = ceil(k_rrt log n), planners like PRM* with KStarStrategy grow k with the roadmap: past the list sizes of
// the tiled / cell kernels a query used to copy all n distances to the host and sort them there)
struct SelectState {
    unsigned long long prefix;   // bits of the k-th smallest key fixed so far (from the top)
    unsigned long long want;     // rank still to find among the keys that share the prefix
    unsigned long long n_out;    // entries emitted by the compaction
    unsigned hist[256];
};

__device__ __forceinline__ unsigned long long dist_key(double d) { return static_cast<unsigned long long>(__double_as_longlong(d)); }  // d >= 0 or +inf: order preserving

__global__ void knn_select_init_kernel(SelectState *s, unsigned long long want, double *__restrict__ d_all, long long self) {
    if (threadIdx.x == 0) { s->prefix = 0; s->want = want; s->n_out = 0; }
    s->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0 && self >= 0) d_all[self] = INFINITY;  // the excluded point can never be selected
}

// histogram of the byte at `shift` over the keys whose higher bits equal the prefix
__global__ void __launch_bounds__(256) knn_select_hist_kernel(const double *__restrict__ d_all, long long n, int shift, SelectState *s) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = s->prefix;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const unsigned long long key = dist_key(d_all[i]);
        if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&h[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&s->hist[threadIdx.x], h[threadIdx.x]);
}

// one thread: the byte in which the wanted rank falls; histogram cleared for the next pass
__global__ void knn_select_step_kernel(SelectState *s, int shift) {
    unsigned long long want = s->want, acc = 0;
    int b = 0;
    for (; b < 255; ++b) {
        if (acc + s->hist[b] >= want) break;
        acc += s->hist[b];
    }
    s->prefix |= static_cast<unsigned long long>(b) << shift;
    s->want = want - acc;
    for (int i = 0; i < 256; ++i) s->hist[i] = 0;
}

// everything at or below the k-th smallest key (ties included), as (key, id) pairs in arbitrary order
__global__ void knn_select_compact_kernel(const double *__restrict__ d_all, long long n, SelectState *s, unsigned long long *__restrict__ keys,
                                          int *__restrict__ ids, long long cap) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = dist_key(d_all[i]);
    if (key <= s->prefix && key < 0x7ff0000000000000ull) {
        const unsigned long long p = atomicAdd(&s->n_out, 1ull);
        if (static_cast<long long>(p) < cap) { keys[p] = key; ids[p] = static_cast<int>(i); }
    }
}

__global__ void knn_select_pad_kernel(unsigned long long *__restrict__ keys, int *__restrict__ ids, const SelectState *s, long long padded) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < padded && i >= static_cast<long long>(s->n_out)) { keys[i] = ~0ull; ids[i] = 0x7fffffff; }
}

// one compare-exchange stage of a bitonic network over (key, id) pairs, ascending by (key, id)
__global__ void knn_bitonic_stage_kernel(unsigned long long *__restrict__ keys, int *__restrict__ ids, long long padded, long long kk, long long j) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    const long long l = i ^ j;
    if (l <= i) return;
    const unsigned long long a = keys[i], b = keys[l];
    const int ia = ids[i], ib = ids[l];
    const bool up = (i & kk) == 0;
    const bool gt = a > b || (a == b && ia > ib);
    if (gt == up) { keys[i] = b; keys[l] = a; ids[i] = ib; ids[l] = ia; }
}

__global__ void knn_select_emit_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ ids, const SelectState *s, int k,
                                       long long *__restrict__ idx, double *__restrict__ dist, int *__restrict__ count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const bool have = static_cast<unsigned long long>(t) < s->n_out;
    idx[t] = have ? ids[t] : -1;
    dist[t] = have ? __longlong_as_double(static_cast<long long>(keys[t])) : INFINITY;
    if (t == 0 && count) *count = static_cast<int>(min(static_cast<unsigned long long>(k), s->n_out));
}

// k nearest of one host query, any k, selected and ordered on the device; results to host buffers. caller holds ctx->mu
int knn_query_large_k(b200mp_knn *k, const double *q, int kq, int64_t self, int64_t *idx, double *dist, int *count) {
    // scratch behind the distances: select state | keys[cap] | ids[cap] | idx[kq] | dist[kq] | count ; cap = padded (k + room for ties)
    long long padded = 1;
    while (padded < static_cast<long long>(kq) + 1024) padded *= 2;
    padded = std::min<long long>(padded, 1LL << 31);
    const size_t o_keys = 2048, o_ids = o_keys + static_cast<size_t>(padded) * 8, o_idx = o_ids + static_cast<size_t>(padded) * 4;
    const size_t o_dist = o_idx + static_cast<size_t>(kq) * 8, o_cnt = o_dist + static_cast<size_t>(kq) * 8, extra = o_cnt + 64;
    double *d_all = nullptr;
    int rc = knn_all_distances(k, q, &d_all, extra);
    if (rc) return rc;
    cudaStream_t st = k->ctx->stream;
    const size_t head = (static_cast<size_t>(k->n) * sizeof(double) + static_cast<size_t>(k->dim) * 3 * sizeof(double) +
                         static_cast<size_t>(k->nl) * 2 * sizeof(float) + 512 + 255) / 256 * 256;
    char *base = k->ctx->scratch.ptr + head;
    SelectState *s = reinterpret_cast<SelectState *>(base);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(base + o_keys);
    int *ids = reinterpret_cast<int *>(base + o_ids);
    long long *d_idx = reinterpret_cast<long long *>(base + o_idx);
    double *d_dist = reinterpret_cast<double *>(base + o_dist);
    int *d_cnt = reinterpret_cast<int *>(base + o_cnt);
    const long long n = k->n;
    const unsigned sel_blocks = static_cast<unsigned>(std::min<long long>((n + 255) / 256, static_cast<long long>(k->ctx->sm_count) * 8));
    const unsigned long long want = static_cast<unsigned long long>(std::min<long long>(kq, n));
    knn_select_init_kernel<<<1, 256, 0, st>>>(s, want, d_all, self);
    for (int shift = 56; shift >= 0; shift -= 8) {
        knn_select_hist_kernel<<<sel_blocks, 256, 0, st>>>(d_all, n, shift, s);
        knn_select_step_kernel<<<1, 1, 0, st>>>(s, shift);
    }
    knn_select_compact_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d_all, n, s, keys, ids, padded);
    knn_select_pad_kernel<<<static_cast<unsigned>((padded + 255) / 256), 256, 0, st>>>(keys, ids, s, padded);
    long long launches = 20;
    for (long long kk2 = 2; kk2 <= padded; kk2 <<= 1)
        for (long long j = kk2 >> 1; j > 0; j >>= 1) {
            knn_bitonic_stage_kernel<<<static_cast<unsigned>((padded + 255) / 256), 256, 0, st>>>(keys, ids, padded, kk2, j);
            ++launches;
        }
    knn_select_emit_kernel<<<(kq + 255) / 256, 256, 0, st>>>(keys, ids, s, kq, d_idx, d_dist, d_cnt);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(launches + 1);
    B200_CUDA(cudaMemcpyAsync(idx, d_idx, static_cast<size_t>(kq) * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(dist, d_dist, static_cast<size_t>(kq) * 8, cudaMemcpyDeviceToHost, st));
    int h_cnt = 0;
    B200_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned long long n_out = 0;
    B200_CUDA(cudaMemcpyAsync(&n_out, &s->n_out, sizeof(n_out), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (static_cast<long long>(n_out) > padded)  // more than 1024 points tie with the k-th distance: not reachable with real data
        return fail(B200MP_ERR_CAPACITY, "knn (large k): %llu points at or below the k-th distance exceed the selection buffer", n_out);
    if (count) *count = h_cnt;
    return B200MP_OK;
}

// exact distance of one query to every point, filtered by radius, sorted by (distance, id); caller holds ctx->mu
int knn_radius_impl(b200mp_knn *k, const double *q, double radius, int64_t self, std::vector<std::pair<double, int64_t>> &sel) {
    double *d_all = nullptr;
    int rc = knn_all_distances(k, q, &d_all);
    if (rc) return rc;
    cudaStream_t st = k->ctx->stream;
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


namespace {

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// development switches (A/B measurements): B200MP_KNN_QPT = queries per batch of the search, B200MP_KNN_CELL_ROWS = rows per leaf
inline int knn_warp_qpt() {  // queries per batch of the query-major search
    const char *e = getenv("B200MP_KNN_QPT");
    const int v = e ? atoi(e) : 1;
    return (v == 2 || v == 4) ? v : 1;
}
inline float knn_env_float(const char *name, float dflt) {
    const char *e = getenv(name);
    return e ? static_cast<float>(atof(e)) : dflt;
}

// leaf (cell) of every feature row under the index's split planes
int knn_assign(b200mp_knn *k, const float *rows, long long n, int *cell, cudaStream_t st) {
    knn_kd_assign_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(rows, n, k->nv * 4, k->index_L, k->ix_dims, k->ix_splits, cell);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

// counting sort of `n` items by cell: start[P + 1] and the slot of every item (order inside a cell is arbitrary)
int knn_cell_sort(b200mp_knn *k, const int *cell, long long n, int *counts, int *start, int *cursor, int *pos, cudaStream_t st) {
    const int P = k->index_P;
    B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * P, st));
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
    knn_cell_count_kernel<<<blocks, 256, 0, st>>>(cell, n, counts);
    knn_cell_scan_kernel<<<1, 1024, 0, st>>>(counts, P, start, cursor);
    knn_cell_slot_kernel<<<blocks, 256, 0, st>>>(cell, n, cursor, pos);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(3);
    return B200MP_OK;
}

int knn_build_index(b200mp_knn *k, cudaStream_t st) {
    const long long n = k->n;
    const int fw = k->nv * 4, dim = k->dim;
    int P = 64, L = 6;
    const long long rows_per_cell = getenv("B200MP_KNN_CELL_ROWS") ? atoll(getenv("B200MP_KNN_CELL_ROWS")) : 64;
    while (P < 65536 && static_cast<long long>(P) * rows_per_cell < n) { P *= 2; ++L; }
    k->index_P = P;
    k->index_L = L;
    size_t off = 0;
    const size_t o_feat = off; off += align256(static_cast<size_t>(n) * fw * sizeof(float));
    const size_t o_pts = off;  off += align256(static_cast<size_t>(n) * dim * sizeof(double));
    const size_t o_orig = off; off += align256(static_cast<size_t>(n) * sizeof(int));
    const size_t o_start = off; off += align256(static_cast<size_t>(P + 1) * sizeof(int));
    const size_t o_lo = off; off += align256(static_cast<size_t>(P) * 24 * sizeof(float));
    const size_t o_hi = off; off += align256(static_cast<size_t>(P) * 24 * sizeof(float));
    const size_t o_split = off; off += align256(static_cast<size_t>(P) * sizeof(float));
    const size_t o_dims = off; off += align256(static_cast<size_t>(L + 1) * sizeof(int));
    const size_t o_box = off; off += align256(static_cast<size_t>(P) * 2 * fw * sizeof(float));
    const size_t o_sbox = off; off += align256(static_cast<size_t>(P / kSuper) * 2 * fw * sizeof(float));
    const size_t o_mbox = off; off += align256(static_cast<size_t>((P + kMega - 1) / kMega) * 2 * fw * sizeof(float));
    int rc = k->index_mem.reserve(off);
    if (rc) return rc;
    char *m = k->index_mem.ptr;
    k->ix_feat = reinterpret_cast<float4 *>(m + o_feat);
    k->ix_pts = reinterpret_cast<double *>(m + o_pts);
    k->ix_orig = reinterpret_cast<int *>(m + o_orig);
    k->ix_start = reinterpret_cast<int *>(m + o_start);
    k->ix_lo = reinterpret_cast<float *>(m + o_lo);
    k->ix_hi = reinterpret_cast<float *>(m + o_hi);
    k->ix_splits = reinterpret_cast<float *>(m + o_split);
    k->ix_dims = reinterpret_cast<int *>(m + o_dims);
    k->ix_box = reinterpret_cast<float4 *>(m + o_box);
    k->ix_sbox = reinterpret_cast<float4 *>(m + o_sbox);
    k->ix_mbox = reinterpret_cast<float4 *>(m + o_mbox);
    // temporaries: cell[n] | pos[n] | counts[P] | cursor[P] | hist[P / 2][bins] | boxes A, B [P][2 dim] | extent[2 dim]
    const size_t a_n = align256(static_cast<size_t>(n) * 4), a_p = align256(static_cast<size_t>(P) * 4);
    const size_t a_hist = align256(static_cast<size_t>(P / 2) * kKdBins * 4), a_box = align256(static_cast<size_t>(P) * 2 * dim * 4);
    rc = k->index_tmp.reserve(2 * a_n + 2 * a_p + a_hist + 2 * a_box + 256);
    if (rc) return rc;
    char *t = k->index_tmp.ptr;
    int *cell = reinterpret_cast<int *>(t);
    int *pos = reinterpret_cast<int *>(t + a_n);
    int *counts = reinterpret_cast<int *>(t + 2 * a_n);
    int *cursor = reinterpret_cast<int *>(t + 2 * a_n + a_p);
    unsigned *hist = reinterpret_cast<unsigned *>(t + 2 * a_n + 2 * a_p);
    float *box_a = reinterpret_cast<float *>(t + 2 * a_n + 2 * a_p + a_hist);
    float *box_b = reinterpret_cast<float *>(t + 2 * a_n + 2 * a_p + a_hist + a_box);
    unsigned *ext = reinterpret_cast<unsigned *>(t + 2 * a_n + 2 * a_p + a_hist + 2 * a_box);
    const unsigned nb = static_cast<unsigned>((n + 255) / 256);
    // ---- split planes, level by level (cell[] doubles as the per-row node index while the tree grows)
    const unsigned nb_stride = static_cast<unsigned>(std::min<long long>(nb, 4LL * k->ctx->sm_count));
    knn_kd_extent_init_kernel<<<1, 32, 0, st>>>(ext, dim);  // (no host round trip: the build stays asynchronous)
    knn_kd_extent_kernel<<<nb_stride, 256, 0, st>>>(k->d_feat, n, fw, dim, ext);
    knn_kd_plan_kernel<<<1, 1, 0, st>>>(ext, dim, L, k->ix_dims, box_a);
    B200_CUDA(cudaMemsetAsync(cell, 0, static_cast<size_t>(n) * 4, st));
    long long launches = 2;
    for (int l = 0; l < L; ++l) {
        B200_CUDA(cudaMemsetAsync(hist, 0, (static_cast<size_t>(1) << l) * kKdBins * 4, st));
        if (l <= kKdSmemLevels)
            knn_kd_hist_smem_kernel<<<std::min<unsigned>((nb + 1) / 2, 2u * k->ctx->sm_count), 512, 0, st>>>(k->d_feat, n, fw, dim, cell, k->ix_dims, l, box_a, hist);
        else
            knn_kd_hist_kernel<<<nb, 256, 0, st>>>(k->d_feat, n, fw, dim, cell, k->ix_dims, l, box_a, hist);
        knn_kd_split_kernel<<<((1 << l) + 63) / 64, 64, 0, st>>>(hist, l, dim, k->ix_dims, box_a, box_b, k->ix_splits);
        knn_kd_descend_kernel<<<nb, 256, 0, st>>>(k->d_feat, n, fw, cell, k->ix_dims, l, k->ix_splits);
        std::swap(box_a, box_b);
        launches += 3;
    }
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(launches);
    if ((rc = knn_cell_sort(k, cell, n, counts, k->ix_start, cursor, pos, st))) return rc;
    knn_permute_kernel<<<nb, 256, 0, st>>>(reinterpret_cast<const float4 *>(k->d_feat), k->d_pts, pos, n, k->dim, k->nv, k->ix_feat, k->ix_pts,
                                           k->ix_orig);
    knn_cell_bounds_kernel<<<(P * 32 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float *>(k->ix_feat), k->ix_start, P, fw, k->dim, k->ix_lo,
                                                                 k->ix_hi);
    knn_pack_boxes_kernel<<<(P / kSuper * fw + 255) / 256, 256, 0, st>>>(k->ix_lo, k->ix_hi, P, k->dim, fw, kSuper, 1, reinterpret_cast<float *>(k->ix_box),
                                                                        reinterpret_cast<float *>(k->ix_sbox));
    knn_pack_boxes_kernel<<<((P + kMega - 1) / kMega * fw + 255) / 256, 256, 0, st>>>(k->ix_lo, k->ix_hi, P, k->dim, fw, kMega, 0, nullptr,
                                                                                     reinterpret_cast<float *>(k->ix_mbox));
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(4);
    k->index_valid = true;
    k->index_n = n;
    return B200MP_OK;
}

template <int NV, int QB, int KPL, int MINB>
int launch_point(b200mp_knn *k, const double *d_q, const int *q_order, const int *q_cell, const int *q_start, long long nq, const long long *d_self,
                 long long self_base, int use_self, int kk, long long *d_idx, double *d_dist, int *d_idx32, int *d_counts, cudaStream_t st, int parts, int part) {
    auto kern = knn_point_kernel<NV, QB, KPL, MINB>;
    int per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPointThreads, 0));
    const float err_scale = 4.0f * (4 * NV + 2) * 5.96e-8f;
    const int chunk = std::max(QB, getenv("B200MP_KNN_CHUNK") ? atoi(getenv("B200MP_KNN_CHUNK")) : 8) / QB * QB;
    const long long warps_needed = (nq / std::max(parts, 1) + chunk - 1) / chunk + 1;
    const long long blocks = std::max<long long>(1, std::min<long long>(static_cast<long long>(k->ctx->sm_count) * std::max(per_sm, 1),
                                                                         (warps_needed + kPointThreads / 32 - 1) / (kPointThreads / 32)));
    const bool want_stats = getenv("B200MP_KNN_STAT_PAIRS") != nullptr;
    B200_CUDA(cudaMemsetAsync(k->d_point_ctl, 0, sizeof(unsigned long long), st));  // the work-queue head; statistics accumulate
    kern<<<static_cast<unsigned>(blocks), kPointThreads, 0, st>>>(
        k->ix_feat, k->ix_pts, k->ix_orig, k->ix_start, k->ix_box, k->ix_sbox, k->ix_mbox, k->index_P, k->dim, d_q, q_order, q_cell, q_start, nq, d_self,
        self_base, use_self, kk, err_scale, k->cmax, knn_env_float("B200MP_KNN_TAU0", 0.3f), knn_env_float("B200MP_KNN_TAUG", 2.0f), parts, part, chunk,
        d_idx, d_dist, d_idx32, d_counts, k->d_point_ctl, want_stats ? k->d_point_ctl + 1 : nullptr);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

template <int NV>
int launch_point_nv(b200mp_knn *k, int qb, const double *d_q, const int *q_order, const int *q_cell, const int *q_start, long long nq,
                    const long long *d_self, long long self_base, int use_self, int kk, long long *d_idx, double *d_dist, int *d_idx32, int *d_counts,
                    cudaStream_t st, int parts, int part) {
#define B200_POINT(QB_, KPL_, MINB_) launch_point<NV, QB_, KPL_, MINB_>(k, d_q, q_order, q_cell, q_start, nq, d_self, self_base, use_self, kk, d_idx, d_dist, d_idx32, d_counts, st, parts, part)
    const int minb = getenv("B200MP_KNN_MINB") ? atoi(getenv("B200MP_KNN_MINB")) : 4;
    if (kk <= 32) {
        if (qb == 1) return minb == 2 ? B200_POINT(1, 1, 2) : (minb == 3 ? B200_POINT(1, 1, 3) : B200_POINT(1, 1, 4));
        return qb == 4 ? B200_POINT(4, 1, 1) : B200_POINT(2, 1, 2);
    }
    return qb == 1 ? B200_POINT(1, 2, 3) : B200_POINT(2, 2, 2);
#undef B200_POINT
}

// large query batches against a large database: cell-pruned exact search
int knn_query_indexed(b200mp_knn *k, const double *d_q, long long nq, int kq, const long long *d_self, long long self_base, int use_self,
                      long long *d_idx, double *d_dist, int *d_counts, cudaStream_t st, int parts = 1, int part = 0, int *d_idx32 = nullptr) {
    int rc;
    if (!k->index_valid || k->index_n != k->n)
        if ((rc = knn_build_index(k, st))) return rc;
    const int fw = k->nv * 4, P = k->index_P;
    // temporaries: qfeat[nq][fw] | cell[nq] | pos[nq] | order[nq] | cell_sorted[nq] | counts[P] | start[P+1] | cursor[P]
    const size_t a_q = align256(static_cast<size_t>(nq) * 4);
    size_t off = 0;
    const size_t o_feat = off; off += align256(static_cast<size_t>(nq) * fw * sizeof(float));
    const size_t o_cell = off; off += a_q;
    const size_t o_pos = off; off += a_q;
    const size_t o_order = off; off += a_q;
    const size_t o_csort = off; off += a_q;
    const size_t o_counts = off; off += align256(static_cast<size_t>(P) * 4);
    const size_t o_start = off; off += align256(static_cast<size_t>(P + 1) * 4);
    const size_t o_cursor = off; off += align256(static_cast<size_t>(P) * 4);
    if ((rc = k->scratch.reserve(off))) return rc;
    char *m = k->scratch.ptr;
    float *qfeat = reinterpret_cast<float *>(m + o_feat);
    int *cell = reinterpret_cast<int *>(m + o_cell), *pos = reinterpret_cast<int *>(m + o_pos);
    int *order = reinterpret_cast<int *>(m + o_order), *csort = reinterpret_cast<int *>(m + o_csort);
    int *counts = reinterpret_cast<int *>(m + o_counts), *start = reinterpret_cast<int *>(m + o_start), *cursor = reinterpret_cast<int *>(m + o_cursor);
    const unsigned qblocks = static_cast<unsigned>((nq + 255) / 256);
    knn_features_kernel<<<qblocks, 256, 0, st>>>(d_q, 0, nq, k->dim, fw, qfeat);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    if ((rc = knn_assign(k, qfeat, nq, cell, st))) return rc;
    if ((rc = knn_cell_sort(k, cell, nq, counts, start, cursor, pos, st))) return rc;
    knn_query_order_kernel<<<qblocks, 256, 0, st>>>(cell, pos, nq, order, csort);
    // query-major search: writes the sorted rows of its part's queries itself
    const int qb = knn_warp_qpt();
    switch (k->nv) {
        case 2: rc = launch_point_nv<2>(k, qb, d_q, order, csort, start, nq, d_self, self_base, use_self, kq, d_idx, d_dist, d_idx32, d_counts, st, parts, part); break;
        case 3: rc = launch_point_nv<3>(k, qb, d_q, order, csort, start, nq, d_self, self_base, use_self, kq, d_idx, d_dist, d_idx32, d_counts, st, parts, part); break;
        default: rc = launch_point_nv<6>(k, qb, d_q, order, csort, start, nq, d_self, self_base, use_self, kq, d_idx, d_dist, d_idx32, d_counts, st, parts, part); break;
    }
    return rc;
}

}  // namespace


namespace {

template <int NL>
int launch_chain(b200mp_knn *k, const float *qcum, const double *qlink, long long nq, const long long *d_self, long long self_base, int use_self,
                 int kk, Partial *d_partial, int splits, cudaStream_t st) {
    const size_t smem = 2 * kChainTile * NL * sizeof(float2) + static_cast<size_t>(kk) * kKnnThreads * (sizeof(double) + sizeof(int));
    auto kern = knn_chain_kernel<NL>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(static_cast<unsigned>((nq + kKnnThreads - 1) / kKnnThreads), static_cast<unsigned>(splits));
    kern<<<grid, kKnnThreads, smem, st>>>(reinterpret_cast<const float2 *>(k->d_cum32), k->d_link64, k->d_dead, k->n, k->dim, k->link_len,
                                         reinterpret_cast<const float2 *>(qcum), qlink, nq, d_self, self_base, use_self, kk, d_partial, splits);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return B200MP_OK;
}

int knn_query_chain(b200mp_knn *k, const double *d_q, long long nq, int kq, const long long *d_self, long long self_base, int use_self,
                    long long *d_idx, double *d_dist, int *d_counts, cudaStream_t st) {
    const long long qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
    long long splits = std::max<long long>(1, (2LL * k->ctx->sm_count + qblocks - 1) / qblocks);
    splits = std::min<long long>(splits, std::max<long long>(1, (k->n + kChainTile - 1) / kChainTile));
    splits = std::min<long long>(splits, 1024);
    const size_t o_cum = 0;
    const size_t o_link = align256(static_cast<size_t>(nq) * 2 * k->nl * sizeof(float));
    const size_t o_part = o_link + align256(static_cast<size_t>(nq) * 2 * k->dim * sizeof(double));
    int rc = k->scratch.reserve(o_part + static_cast<size_t>(nq) * splits * kq * sizeof(Partial) + 256);
    if (rc) return rc;
    float *qcum = reinterpret_cast<float *>(k->scratch.ptr + o_cum);
    double *qlink = reinterpret_cast<double *>(k->scratch.ptr + o_link);
    Partial *partial = reinterpret_cast<Partial *>(k->scratch.ptr + o_part);
    knn_chain_tables_kernel<<<static_cast<unsigned>((nq + 127) / 128), 128, 0, st>>>(d_q, 0, nq, k->dim, k->nl, qcum, qlink);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    switch (k->nl) {
        case 8: rc = launch_chain<8>(k, qcum, qlink, nq, d_self, self_base, use_self, kq, partial, static_cast<int>(splits), st); break;
        case 20: rc = launch_chain<20>(k, qcum, qlink, nq, d_self, self_base, use_self, kq, partial, static_cast<int>(splits), st); break;
        default: rc = launch_chain<32>(k, qcum, qlink, nq, d_self, self_base, use_self, kq, partial, static_cast<int>(splits), st); break;
    }
    if (rc) return rc;
    knn_merge_kernel<<<static_cast<unsigned>((nq * 32 + 255) / 256), 256, 0, st>>>(partial, nq, static_cast<int>(splits) * kq, kq, d_idx, d_dist, d_counts, 0);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
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
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&k->d_rows_visited, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(k->d_rows_visited, 0, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&k->d_point_ctl, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(k->d_point_ctl, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        if (k->d_rows_visited) cudaFree(k->d_rows_visited);
        delete k;
        return fail(B200MP_ERR_CUDA, "b200mp_knn_create: device allocation failed");
    }
    *out = k;
    return B200MP_OK;
}

int b200mp_knn_create_chain(b200mp_ctx *ctx, int n_links, double link_length, b200mp_knn **out) {
    if (!ctx || !out || n_links < 1 || !(link_length > 0.0)) return fail(B200MP_ERR_INVALID, "b200mp_knn_create_chain: bad argument");
    if (n_links > 32) return fail(B200MP_ERR_CAPACITY, "b200mp_knn_create_chain: at most 32 links");
    auto *k = new b200mp_knn();
    k->ctx = ctx;
    k->dim = n_links;
    k->metric = B200MP_METRIC_CHAIN;
    k->nv = 1;  // the L2 feature rows are not used by this metric
    k->nl = n_links <= 8 ? 8 : (n_links <= 20 ? 20 : 32);
    k->link_len = link_length;
    k->use_index = false;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&k->d_rows_visited, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(k->d_rows_visited, 0, sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        delete k;
        return fail(B200MP_ERR_CUDA, "b200mp_knn_create_chain: device allocation failed");
    }
    *out = k;
    return B200MP_OK;
}

void b200mp_knn_destroy(b200mp_knn *k) {
    if (!k) return;
    cudaSetDevice(k->ctx->device);
    cudaStreamSynchronize(k->ctx->stream);
    if (k->d_pts) cudaFree(k->d_pts);
    if (k->d_feat) cudaFree(k->d_feat);
    if (k->d_cum32) cudaFree(k->d_cum32);
    if (k->d_link64) cudaFree(k->d_link64);
    if (k->d_dead) cudaFree(k->d_dead);
    k->scratch.release();
    k->qscratch.release();
    k->index_mem.release();
    k->index_tmp.release();
    if (k->d_rows_visited) cudaFree(k->d_rows_visited);
    if (k->d_point_ctl) cudaFree(k->d_point_ctl);
    delete k;
}

int b200mp_knn_set_index(b200mp_knn *k, int enable) {
    if (!k) return fail(B200MP_ERR_INVALID, "b200mp_knn_set_index: NULL");
    k->use_index = enable != 0;
    return B200MP_OK;
}

int64_t b200mp_knn_rows_visited(b200mp_knn *k) {
    if (!k) return -1;
    unsigned long long v = 0;
    cudaSetDevice(k->ctx->device);
    cudaStreamSynchronize(k->ctx->stream);
    if (cudaMemcpy(&v, k->d_rows_visited, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    unsigned long long p = 0;
    if (k->d_point_ctl && cudaMemcpy(&p, k->d_point_ctl + 1, sizeof(p), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<int64_t>(v + p);
}

int b200mp_knn_search_stats(b200mp_knn *k, int64_t *out) {
    if (!k || !out) return fail(B200MP_ERR_INVALID, "b200mp_knn_search_stats: NULL argument");
    unsigned long long v[4] = {0, 0, 0, 0};
    B200_CUDA(cudaSetDevice(k->ctx->device));
    B200_CUDA(cudaStreamSynchronize(k->ctx->stream));
    if (k->d_point_ctl) B200_CUDA(cudaMemcpy(v, k->d_point_ctl + 1, sizeof(v), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = static_cast<int64_t>(v[i]);
    return B200MP_OK;
}

int b200mp_knn_clear(b200mp_knn *k) {
    if (!k) return fail(B200MP_ERR_INVALID, "b200mp_knn_clear: NULL");
    k->n = 0;
    k->n_removed = 0;
    k->cmax = 0.f;
    k->index_valid = false;
    if (k->d_dead && k->cap > 0) {
        cudaSetDevice(k->ctx->device);
        cudaMemset(k->d_dead, 0, static_cast<size_t>(k->cap));
    }
    k->h_pts.clear();
    k->alive.clear();
    return B200MP_OK;
}

int64_t b200mp_knn_size(const b200mp_knn *k) { return k ? k->n - k->n_removed : -1; }

int b200mp_knn_add(b200mp_knn *k, const double *pts, int64_t n, int64_t *first_id) {
    if (!k || (n > 0 && !pts) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_knn_add: bad argument");
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    if (first_id) *first_id = k->n;
    if (n == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(k->ctx->device));
    int rc = knn_reserve(k, k->n + n);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(k->d_pts + k->n * k->dim, pts, static_cast<size_t>(n) * k->dim * sizeof(double), cudaMemcpyHostToDevice,
                              k->ctx->stream));
    if ((rc = knn_build_features(k, k->n, n, k->ctx->stream))) return rc;
    k->h_pts.insert(k->h_pts.end(), pts, pts + n * k->dim);
    k->alive.resize(static_cast<size_t>(k->n + n), 1);
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
    k->n += n;
    k->alive.resize(static_cast<size_t>(k->n), 1);  // liveness stays indexed by global id
    return B200MP_OK;
}

int b200mp_knn_remove(b200mp_knn *k, int64_t id) {
    if (!k) return fail(B200MP_ERR_INVALID, "b200mp_knn_remove: NULL");
    if (id < 0 || id >= k->n || static_cast<size_t>(id) >= k->alive.size() || !k->alive[id]) return fail(B200MP_ERR_INVALID, "b200mp_knn_remove: id %lld not present", (long long)id);
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    B200_CUDA(cudaSetDevice(k->ctx->device));
    if (k->metric == B200MP_METRIC_CHAIN) {
        const unsigned char one = 1;
        B200_CUDA(cudaMemcpy(k->d_dead + id, &one, 1, cudaMemcpyHostToDevice));
    } else {
        const float inf = INFINITY;  // tombstone: |c|^2 = +inf never passes the filter
        B200_CUDA(cudaMemcpy(k->d_feat + id * k->nv * 4 + (k->nv * 4 - 1), &inf, sizeof(float), cudaMemcpyHostToDevice));
    }
    k->alive[id] = 0;
    k->n_removed++;
    k->index_valid = false;
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
    if (k->metric == B200MP_METRIC_CHAIN)
        return knn_query_chain(k, d_q, nq, kq, reinterpret_cast<const long long *>(d_self_ids), self_base, exclude_self,
                               reinterpret_cast<long long *>(d_idx), d_dist, d_counts, st);
    if (k->use_index && nq >= 4096 && k->n >= 32768 && kq <= 64 && k->n <= 0x7fffffffLL)
        return knn_query_indexed(k, d_q, nq, kq, reinterpret_cast<const long long *>(d_self_ids), self_base, exclude_self,
                                 reinterpret_cast<long long *>(d_idx), d_dist, d_counts, st);
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
            d_counts ? d_counts + q0 : nullptr, k->metric == B200MP_METRIC_L2 ? 1 : 0);
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

int b200mp_knn_query_part_device_f32(b200mp_knn *k, const float *d_q, int64_t nq, int kq, int64_t self_base, int exclude_self, int part, int n_parts,
                                     int64_t *d_idx, double *d_dist, void *stream) {
    if (!k || nq < 0 || kq < 1 || n_parts < 1 || part < 0 || part >= n_parts || (nq > 0 && (!d_q || !d_idx || !d_dist)))
        return fail(B200MP_ERR_INVALID, "b200mp_knn_query_part_device_f32: bad argument");
    if (nq == 0) return B200MP_OK;
    if (kq > kMaxK) return fail(B200MP_ERR_CAPACITY, "b200mp_knn_query_part_device_f32: k %d > %d", kq, kMaxK);
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : k->ctx->stream;
    const bool indexed = k->metric == B200MP_METRIC_L2 && k->use_index && nq >= 4096 && k->n >= 32768 && kq <= 64 && k->n <= 0x7fffffffLL;
    if (!indexed) {
        // contiguous share of the queries through the ordinary path
        const int64_t per = (nq + n_parts - 1) / n_parts, b = std::min<int64_t>(per * part, nq), cnt = std::min<int64_t>(per, nq - b);
        if (cnt <= 0) return B200MP_OK;
        return b200mp_knn_query_device_f32(k, d_q + b * k->dim, cnt, kq, self_base + b, exclude_self, d_idx + b * kq, d_dist + b * kq, nullptr, st);
    }
    const long long cnt = static_cast<long long>(nq) * k->dim;
    int rc = k->qscratch.reserve(static_cast<size_t>(cnt) * sizeof(double));
    if (rc) return rc;
    double *d_q64 = reinterpret_cast<double *>(k->qscratch.ptr);
    knn_f32_to_f64_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(d_q, cnt, d_q64);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    // queries are shared out by the index cell they fall into (cell c -> part c mod n_parts): every part works on whole cells
    return knn_query_indexed(k, d_q64, nq, kq, nullptr, self_base, exclude_self, reinterpret_cast<long long *>(d_idx), d_dist, nullptr, st, n_parts, part);
}

int b200mp_knn_query_part_ids_device_f32(b200mp_knn *k, const float *d_q, int64_t nq, int kq, int64_t self_base, int exclude_self, int part, int n_parts,
                                         int32_t *d_ids, void *stream) {
    if (!k || nq < 0 || kq < 1 || n_parts < 1 || part < 0 || part >= n_parts || (nq > 0 && (!d_q || !d_ids)))
        return fail(B200MP_ERR_INVALID, "b200mp_knn_query_part_ids_device_f32: bad argument");
    if (nq == 0) return B200MP_OK;
    if (kq > 64) return fail(B200MP_ERR_CAPACITY, "b200mp_knn_query_part_ids_device_f32: k %d > 64", kq);
    if (k->metric != B200MP_METRIC_L2 || k->n > 0x7fffffffLL) return fail(B200MP_ERR_INVALID, "b200mp_knn_query_part_ids_device_f32: L2 containers only");
    B200_CUDA(cudaSetDevice(k->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : k->ctx->stream;
    const long long cnt = static_cast<long long>(nq) * k->dim;
    int rc = k->qscratch.reserve(static_cast<size_t>(cnt) * sizeof(double));
    if (rc) return rc;
    double *d_q64 = reinterpret_cast<double *>(k->qscratch.ptr);
    knn_f32_to_f64_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(d_q, cnt, d_q64);
    B200_CUDA(cudaGetLastError());
    k->ctx->launches.fetch_add(1);
    return knn_query_indexed(k, d_q64, nq, kq, nullptr, self_base, exclude_self, nullptr, nullptr, nullptr, st, n_parts, part, d_ids);
}

int b200mp_knn_query(b200mp_knn *k, const double *q, int64_t nq, int kq, const int64_t *self_ids, int64_t *idx, double *dist, int *counts) {
    if (!k || nq < 0 || kq < 0 || (nq > 0 && (!q || (kq > 0 && (!idx || !dist))))) return fail(B200MP_ERR_INVALID, "b200mp_knn_query: bad argument");
    if (nq == 0 || kq == 0) {
        if (counts) for (int64_t i = 0; i < nq; ++i) counts[i] = 0;
        return B200MP_OK;
    }
    std::lock_guard<std::mutex> lk(k->ctx->mu);
    if (kq > kMaxK) {
        // large k (RRT* asks for k ~ k_rrt log n, k >= size means "everything, sorted"): per query all exact distances, then a radix
        // select + sort on the device; only the k results cross to the host
        for (int64_t i = 0; i < nq; ++i) {
            int c = 0;
            if (k->n > 0) {
                const int rc = knn_query_large_k(k, q + i * k->dim, kq, self_ids ? self_ids[i] : -1, idx + i * kq, dist + i * kq, &c);
                if (rc) return rc;
            } else {
                for (int64_t t = 0; t < kq; ++t) { idx[i * kq + t] = -1; dist[i * kq + t] = INFINITY; }
            }
            if (counts) counts[i] = c;
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
    if (qb + ib + db + sb + cb + 64 <= k->ctx->small_cap && nq <= 64) {
        // a planner's single nearest / nearestK: query written into the mapped pinned buffer, results read back from it --
        // no copies to enqueue, the kernels address the host memory directly
        char *h = static_cast<char *>(k->ctx->h_small), *d = static_cast<char *>(k->ctx->d_small);
        size_t o_q = 0, o_i = qb, o_d = qb + ib, o_s = qb + ib + db, o_c = qb + ib + db + sb;
        o_c = (o_c + 7) / 8 * 8;
        std::memcpy(h + o_q, q, qb);
        if (self_ids) std::memcpy(h + o_s, self_ids, sb);
        int rc2 = b200mp_knn_query_device(k, reinterpret_cast<const double *>(d + o_q), nq, kq, self_ids ? reinterpret_cast<const int64_t *>(d + o_s) : nullptr, 0,
                                          self_ids ? 1 : 0, reinterpret_cast<int64_t *>(d + o_i), reinterpret_cast<double *>(d + o_d),
                                          reinterpret_cast<int *>(d + o_c), st);
        if (rc2) return rc2;
        B200_CUDA(cudaStreamSynchronize(st));
        std::memcpy(idx, h + o_i, ib);
        std::memcpy(dist, h + o_d, db);
        if (counts) std::memcpy(counts, h + o_c, cb);
        return B200MP_OK;
    }
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
