// Point-cloud tree built on the device (b200mp_env_add_pointcloud_device). Same structure as the reference's CAPT
// (external/vamp/src/impl/vamp/collision/capt.hh: implicit k-d tree over 2^L leaves with axes cycling x, y, z; per leaf the
// list of every point within r_max + r_point of the leaf's cell; per leaf the bounding box of the listed points; the cloud's
// bounding box) and the same query, but not the same arrays: the reference splits at exact medians found by sorting every
// level (O(n log^2 n), serial recursion); here every level takes the approximate median of each node from a 64-bin histogram,
// and the lists are gathered by one block per leaf. A leaf's list holds every point that a sphere of radius <= r_max centred
// anywhere in the leaf's cell can touch, so the query `exists p: |p - c| <= r + r_point` is answered exactly for any set of
// split planes (tests/test_capt_gpu.py: equal to brute force over the points). The reference's recursive build hands candidate
// points down through slab tests with an early break and occasionally drops a point a leaf needs; the host build reproduces
// that (its arrays equal the reference's), this one does not, so this tree rejects a few states (0.02 % in the test) that the
// reference-identical tree accepts and never the other way round. The reference's shortcut for cells inside the r_min ball
// of their own point is not needed for exactness and is not applied.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "validate.cuh"

namespace b200mp {

constexpr int kCgBins = 64;

__device__ __forceinline__ unsigned cg_fkey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float cg_unfkey(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ext[0..2] = min keys, ext[3..5] = max keys of x, y, z
__global__ void cg_extent_kernel(const float *__restrict__ xyz, long long n, unsigned *__restrict__ ext) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int a = 0; a < 3; ++a) {
        unsigned lo = 0xffffffffu, hi = 0u;
        if (i < n) { lo = cg_fkey(xyz[3 * i + a]); hi = lo; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0 && lo <= hi) {
            atomicMin(ext + a, lo);
            atomicMax(ext + 3 + a, hi);
        }
    }
}

__global__ void cg_root_kernel(const unsigned *__restrict__ ext, float *__restrict__ box, float *__restrict__ top) {
    if (threadIdx.x < 3) {
        const float lo = cg_unfkey(ext[threadIdx.x]), hi = cg_unfkey(ext[3 + threadIdx.x]);
        box[2 * threadIdx.x] = lo;
        box[2 * threadIdx.x + 1] = hi;
        top[threadIdx.x] = lo;
        top[3 + threadIdx.x] = hi;
    }
}

// level l: histogram of the points of every node along axis l % 3 over the node's current extent
__global__ void cg_hist_kernel(const float *__restrict__ xyz, long long n, const int *__restrict__ node, int axis, const float *__restrict__ boxes,
                               unsigned *__restrict__ hist) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nd = node[i];
    const float lo = boxes[(static_cast<size_t>(nd) * 3 + axis) * 2], hi = boxes[(static_cast<size_t>(nd) * 3 + axis) * 2 + 1];
    int bin = 0;
    if (hi > lo) bin = min(max(static_cast<int>((xyz[3 * i + axis] - lo) / (hi - lo) * kCgBins), 0), kCgBins - 1);
    atomicAdd(hist + static_cast<size_t>(nd) * kCgBins + bin, 1u);
}

__global__ void cg_split_kernel(const unsigned *__restrict__ hist, int l, int axis, const float *__restrict__ boxes, float *__restrict__ child_boxes,
                                float *__restrict__ tests) {
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= (1 << l)) return;
    const unsigned *h = hist + static_cast<size_t>(nd) * kCgBins;
    unsigned total = 0;
    for (int b = 0; b < kCgBins; ++b) total += h[b];
    const float lo = boxes[(static_cast<size_t>(nd) * 3 + axis) * 2], hi = boxes[(static_cast<size_t>(nd) * 3 + axis) * 2 + 1];
    int cut = kCgBins / 2;
    if (total > 0) {
        unsigned acc = 0;
        for (int b = 0; b < kCgBins; ++b) {
            acc += h[b];
            if (2u * acc >= total) { cut = b + 1; break; }
        }
    }
    const float split = hi > lo ? lo + (hi - lo) * (static_cast<float>(cut) / kCgBins) : lo;
    tests[(1 << l) - 1 + nd] = split;
    for (int j = 0; j < 3; ++j) {
        const float bl = boxes[(static_cast<size_t>(nd) * 3 + j) * 2], bh = boxes[(static_cast<size_t>(nd) * 3 + j) * 2 + 1];
        float *c0 = child_boxes + (static_cast<size_t>(2 * nd) * 3 + j) * 2, *c1 = child_boxes + (static_cast<size_t>(2 * nd + 1) * 3 + j) * 2;
        c0[0] = bl; c0[1] = j == axis ? split : bh;
        c1[0] = j == axis ? split : bl; c1[1] = bh;
    }
}

__global__ void cg_descend_kernel(const float *__restrict__ xyz, long long n, int *__restrict__ node, int l, int axis, const float *__restrict__ tests) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nd = node[i];
    node[i] = 2 * nd + (xyz[3 * i + axis] >= tests[(1 << l) - 1 + nd] ? 1 : 0);  // the query's own comparison (capt.hh:444-455)
}

__device__ __forceinline__ float cg_box_distsq(const float *bx, float x, float y, float z) {
    const float dx = fmaxf(fmaxf(bx[0] - x, x - bx[3]), 0.f), dy = fmaxf(fmaxf(bx[1] - y, y - bx[4]), 0.f), dz = fmaxf(fmaxf(bx[2] - z, z - bx[5]), 0.f);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// One block per leaf. FILL == false: number of points within reach of the leaf's cell (-> records of 8, rounded up).
// FILL == true: write them (any order: the lists are ordered afterwards), pad with +inf, and the bounding box of the list.
template <bool FILL>
__global__ void __launch_bounds__(128) cg_lists_kernel(const float *__restrict__ xyz, int n, const float *__restrict__ boxes, float reach_sq,
                                                       uint32_t *__restrict__ counts8, const uint32_t *__restrict__ aff_starts, float *__restrict__ aff,
                                                       long long n_aff8, float *__restrict__ aabbs) {
    __shared__ unsigned s_cnt;
    __shared__ unsigned s_bb[6];
    const unsigned leaf = blockIdx.x;
    float bx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) bx[k] = boxes[6ull * leaf + k];
    if (threadIdx.x == 0) s_cnt = 0;
    if (threadIdx.x < 3) { s_bb[threadIdx.x] = 0xffffffffu; s_bb[3 + threadIdx.x] = 0u; }
    __syncthreads();
    const unsigned long long base = FILL ? 8ull * aff_starts[leaf] : 0ull;
    const unsigned long long stride = 8ull * static_cast<unsigned long long>(n_aff8);
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool in = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < n) {
            x = xyz[3 * i]; y = xyz[3 * i + 1]; z = xyz[3 * i + 2];
            in = cg_box_distsq(bx, x, y, z) <= reach_sq;
        }
        if (in) {
            const unsigned at = atomicAdd(&s_cnt, 1u);
            if (FILL) {
                aff[base + at] = x;
                aff[stride + base + at] = y;
                aff[2 * stride + base + at] = z;
                atomicMin(&s_bb[0], cg_fkey(x)); atomicMin(&s_bb[1], cg_fkey(y)); atomicMin(&s_bb[2], cg_fkey(z));
                atomicMax(&s_bb[3], cg_fkey(x)); atomicMax(&s_bb[4], cg_fkey(y)); atomicMax(&s_bb[5], cg_fkey(z));
            }
        }
    }
    __syncthreads();
    if (!FILL) {
        if (threadIdx.x == 0) counts8[leaf] = (s_cnt + 7u) / 8u;
        return;
    }
    const unsigned cnt = s_cnt, padded = (cnt + 7u) / 8u * 8u;
    for (unsigned i = cnt + threadIdx.x; i < padded; i += blockDim.x) {
        aff[base + i] = INFINITY;
        aff[stride + base + i] = INFINITY;
        aff[2 * stride + base + i] = INFINITY;
    }
    if (threadIdx.x < 6) {
        // an empty list gets an inverted box: every query is rejected by the leaf's box test
        aabbs[6ull * leaf + threadIdx.x] = cnt ? cg_unfkey(s_bb[threadIdx.x]) : (threadIdx.x < 3 ? INFINITY : -INFINITY);
    }
}

// exclusive scan of counts8[0..m) into starts[0..m] by one block (m up to 2^20 leaves)
__global__ void __launch_bounds__(1024) cg_scan_kernel(const uint32_t *__restrict__ counts8, unsigned m, uint32_t *__restrict__ starts) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < m; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < m ? counts8[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const unsigned carry = s_carry;
        if (i < m) starts[i] = carry + warp_off + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) starts[m] = s_carry;
}

}  // namespace b200mp
