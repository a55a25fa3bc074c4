// Device-side preparation of an uploaded point-cloud tree (see capt_collides_warp in validate.cuh): every leaf's
// affordance list is re-ordered by distance from the leaf's cell and a table of prefix lengths for the ladder of radii
// capt_reach(k) is built. The cell of a leaf is the box cut out by the split planes on the path to it
// (external/vamp/src/impl/vamp/collision/capt.hh:444-455: child = 2 idx + 1 + (c[axis] >= tests[idx]), axes cycling x, y, z).
// The arrays a caller passes in / exports keep the reference's order; only the device copy is re-ordered.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "validate.cuh"

namespace b200mp {

constexpr int kLeafSortMax = 4096;  // longest list ordered in shared memory; longer lists stay as they are (prefix = whole list)

__global__ void capt_cell_boxes_kernel(const float *__restrict__ tests, int nlog2, unsigned n_leaves, float *__restrict__ boxes) {
    const unsigned leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    float lo[3] = {-INFINITY, -INFINITY, -INFINITY}, hi[3] = {INFINITY, INFINITY, INFINITY};
    unsigned idx = 0;
    for (int i = 0; i < nlog2; ++i) {
        const int axis = i % 3;
        const unsigned right = (leaf >> (nlog2 - 1 - i)) & 1u;
        const float t = tests[idx];
        if (t == t) {  // NaN: plane never written, every query goes left and nothing is cut off
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (a == axis) {
                    if (right) lo[a] = fmaxf(lo[a], t); else hi[a] = fminf(hi[a], t);
                }
            }
        }
        idx = 2u * idx + 1u + right;
    }
    float *b = boxes + 6ull * leaf;
    b[0] = lo[0]; b[1] = lo[1]; b[2] = lo[2];
    b[3] = hi[0]; b[4] = hi[1]; b[5] = hi[2];
}

// one block per leaf
__global__ void __launch_bounds__(256) capt_order_leaf_kernel(const float *__restrict__ aff_in, float *__restrict__ aff_out, long long n_aff8,
                                                              const uint32_t *__restrict__ aff_starts, const float *__restrict__ boxes,
                                                              unsigned short *__restrict__ pref) {
    __shared__ unsigned long long keys[kLeafSortMax];
    const unsigned leaf = blockIdx.x;
    const unsigned long long b = 8ull * aff_starts[leaf];
    const unsigned n = 8u * (aff_starts[leaf + 1] - aff_starts[leaf]);
    const unsigned long long stride = 8ull * static_cast<unsigned long long>(n_aff8);
    if (n > kLeafSortMax) {
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
            for (int a = 0; a < 3; ++a) aff_out[a * stride + b + i] = aff_in[a * stride + b + i];
        if (threadIdx.x < kCaptBuckets) pref[static_cast<size_t>(leaf) * kCaptBuckets + threadIdx.x] = 0xffffu;
        return;
    }
    unsigned npad = 2;
    while (npad < n) npad <<= 1;
    const float *bx = boxes + 6ull * leaf;
    const float lo0 = bx[0], lo1 = bx[1], lo2 = bx[2], hi0 = bx[3], hi1 = bx[4], hi2 = bx[5];
    for (unsigned i = threadIdx.x; i < npad; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            const float x = aff_in[b + i], y = aff_in[stride + b + i], z = aff_in[2 * stride + b + i];
            float d2 = INFINITY;  // padding points (+inf) go last
            if (isfinite(x) && isfinite(y) && isfinite(z)) {
                const float dx = fmaxf(fmaxf(lo0 - x, x - hi0), 0.f), dy = fmaxf(fmaxf(lo1 - y, y - hi1), 0.f), dz = fmaxf(fmaxf(lo2 - z, z - hi2), 0.f);
                d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            }
            key = (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | i;  // d2 >= 0: the bit pattern orders like the value
        }
        keys[i] = key;
    }
    __syncthreads();
    for (unsigned k = 2; k <= npad; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = threadIdx.x; t < npad / 2; t += blockDim.x) {
                const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const unsigned l = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long a = keys[i], c = keys[l];
                if ((a > c) == up) { keys[i] = c; keys[l] = a; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned j = static_cast<unsigned>(keys[i] & 0xffffffffu);
        for (int a = 0; a < 3; ++a) aff_out[a * stride + b + i] = aff_in[a * stride + b + j];
    }
    if (threadIdx.x < kCaptBuckets) {
        const float rr = capt_reach(threadIdx.x) * capt_reach(threadIdx.x);
        const unsigned lim = __float_as_uint(rr);
        unsigned lo = 0, hi = n;  // number of keys with d2 <= rr
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (static_cast<unsigned>(keys[mid] >> 32) <= lim) lo = mid + 1; else hi = mid;
        }
        const unsigned rec = (lo + 7u) / 8u;
        pref[static_cast<size_t>(leaf) * kCaptBuckets + threadIdx.x] = static_cast<unsigned short>(rec >= 0xffffu ? 0xffffu : rec);
    }
}

}  // namespace b200mp
