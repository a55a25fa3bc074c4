// Point-cloud filter on the device (sm_100a): the reference's vamp::collision::filter_pointcloud
// (external/vamp/src/impl/vamp/collision/filter.hh:176-274) -- cull to a range ball + workspace box, then six passes of
// "sort along a Morton curve, drop every point within min_dist of the last point kept", one pass per axis permutation.
//
// What the reference does serially is split into data-parallel steps per pass:
//   codes      one thread per list entry: remap + bit-interleave (filter.hh:144-170) -> 64-bit key (code << 32 | position),
//              plus a block reduction of the coordinate extrema that shrink the curve range for the next pass;
//   sort       bitonic network on the keys (shared-memory stages for spans <= 2048, one launch per larger span);
//              the position in the low half makes the order total, i.e. a stable sort by code;
//   next       nxt[i] = first j > i farther than min_dist from entry i (short forward scans over the sorted points);
//   chain      the kept set is {0, nxt[0], nxt[nxt[0]], ...}: segments of 1024 entries compute where a chain entering at
//              each of their entries leaves them (one thread per segment, back to front), one thread hops segment to
//              segment recording the entry point of each, then every segment walks its own piece, counts, and after
//              a scan of the counts writes its kept entries in order.
// Quirks of the reference that are reproduced because they change the output (listed with file:line in DESIGN.md): both ends of
// the curve range are std::min over the axes; the work list keeps its size n after culling, so its tail entries all refer
// to point 0 (one such entry is equivalent to the whole tail: the others sit right behind it in the order and are at
// distance 0); negative / > 1000 remapped coordinates wrap like the x86-64 float -> uint32 conversion.
// Tie order: the reference sorts with an unstable sort, so among equal codes its order is unspecified; here ties keep
// their previous relative order. Outputs equal the reference's whenever codes are distinct (tests/test_filter_*.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200mp.h"
#include "host_common.hpp"

using namespace b200mp;

namespace {

constexpr int kSeg = 1024;          // entries per chain segment
constexpr int kSortBlock = 2048;    // keys sorted per block in shared memory
constexpr unsigned long long kPadKey = ~0ull;

struct FilterParams {
    float origin[3], ws_min[3], ws_max[3];
    float sqrange;
    int cull;
};

__device__ __forceinline__ float sq3(float ax, float ay, float az, float bx, float by, float bz) {
    const float xs = ax - bx, ys = ay - by, zs = az - bz;
    return fmaf(zs, zs, fmaf(ys, ys, xs * xs));
}

// keep flag of the cull step (filter.hh:203-218)
__global__ void filter_cull_kernel(const float *__restrict__ xyz, long long n, FilterParams p, unsigned char *__restrict__ flag) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    bool keep = true;
    if (p.cull)
        keep = sq3(x, y, z, p.origin[0], p.origin[1], p.origin[2]) < p.sqrange && p.ws_min[0] <= x && x <= p.ws_max[0] &&
               p.ws_min[1] <= y && y <= p.ws_max[1] && p.ws_min[2] <= z && z <= p.ws_max[2];
    flag[i] = keep ? 1 : 0;
}

// order-preserving compaction of the kept indices: per-block counts, one-block scan, scatter
__global__ void filter_count_kernel(const unsigned char *__restrict__ flag, long long n, unsigned *__restrict__ block_count) {
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, i < n && flag[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, __popc(b));
    __syncthreads();
    if (threadIdx.x == 0) block_count[blockIdx.x] = s_cnt;
}

// exclusive scan of `count[0..m)` by one block; total -> *total
__global__ void __launch_bounds__(1024) filter_scan_kernel(unsigned *__restrict__ count, long long m, unsigned *__restrict__ total) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < m; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned v = i < m ? count[i] : 0u;
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
        if (i < m) count[i] = carry + warp_off + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void filter_scatter_kernel(const unsigned char *__restrict__ flag, long long n, const unsigned *__restrict__ block_off,
                                      unsigned *__restrict__ idx) {
    __shared__ unsigned s_woff[32];
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool keep = i < n && flag[i];
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_woff[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) { const unsigned c = s_woff[k]; s_woff[k] = acc; acc += c; }
    }
    __syncthreads();
    if (keep) idx[block_off[blockIdx.x] + s_woff[w] + __popc(b & ((1u << lane) - 1u))] = static_cast<unsigned>(i);
}

// _pdep_u32(v, mask) for the three Morton masks: bit k of v goes to bit 3k + shift
__device__ __forceinline__ unsigned spread3(unsigned v, int shift, int bits) {
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k)
        if (k < bits) out |= ((v >> k) & 1u) << (3 * k + shift);
    return out;
}

__device__ __forceinline__ unsigned remap_point(float x, float mn, float mx) {
    const float v = __fmul_rn(__fdiv_rn(x - mn, mx - mn), 1000.0f);
    return static_cast<unsigned>(__float2ll_rz(v));  // float -> int64 -> low 32 bits, as the x86-64 conversion does
}

__device__ __forceinline__ unsigned fkey(float f) {  // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// keys of one pass + extrema of the coordinates of every listed point (encoded with fkey, atomicMin / atomicMax)
__global__ void filter_codes_kernel(const float *__restrict__ xyz, const unsigned *__restrict__ idx, unsigned m, unsigned m_pad, int a0, int a1,
                                    int a2, float mn, float mx, unsigned long long *__restrict__ keys, unsigned *__restrict__ extrema) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned lo = 0xffffffffu, hi = 0u;
    if (e < m) {
        const float *p = xyz + 3ull * idx[e];
        const float c[3] = {p[0], p[1], p[2]};
        const unsigned code = spread3(remap_point(c[a0], mn, mx), 0, 11) | spread3(remap_point(c[a1], mn, mx), 1, 11) |
                              spread3(remap_point(c[a2], mn, mx), 2, 10);
        keys[e] = (static_cast<unsigned long long>(code) << 32) | e;
        lo = fkey(fminf(fminf(c[0], c[1]), c[2]));
        hi = fkey(fmaxf(fmaxf(c[0], c[1]), c[2]));
    } else if (e < m_pad) {
        keys[e] = kPadKey;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(&extrema[0], lo);
        atomicMax(&extrema[1], hi);
    }
}

// bitonic network. Stage (k, j): element i is compared with i ^ j, ascending when (i & k) == 0.
// Shared-memory kernel: runs, for one block of kSortBlock keys, either the whole network up to k = kSortBlock (k_first == 2)
// or the tail j = kSortBlock / 2 .. 1 of one stage k > kSortBlock.
__global__ void __launch_bounds__(kSortBlock / 2) filter_sort_local_kernel(unsigned long long *__restrict__ keys, unsigned k_first, unsigned k_last) {
    __shared__ unsigned long long s[kSortBlock];
    const unsigned base = blockIdx.x * kSortBlock;
    s[threadIdx.x] = keys[base + threadIdx.x];
    s[threadIdx.x + kSortBlock / 2] = keys[base + threadIdx.x + kSortBlock / 2];
    __syncthreads();
    for (unsigned k = k_first; k <= k_last; k <<= 1) {
        for (unsigned j = min(k >> 1, static_cast<unsigned>(kSortBlock / 2)); j > 0; j >>= 1) {
            const unsigned t = threadIdx.x;
            const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // the lower index of the t-th pair at distance j
            const unsigned l = i | j;
            const bool up = ((base + i) & k) == 0;
            const unsigned long long a = s[i], b = s[l];
            if ((a > b) == up) { s[i] = b; s[l] = a; }
            __syncthreads();
        }
    }
    keys[base + threadIdx.x] = s[threadIdx.x];
    keys[base + threadIdx.x + kSortBlock / 2] = s[threadIdx.x + kSortBlock / 2];
}

__global__ void filter_sort_global_kernel(unsigned long long *__restrict__ keys, unsigned m_pad, unsigned k, unsigned j) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m_pad / 2) return;
    const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const unsigned l = i | j;
    const bool up = (i & k) == 0;
    const unsigned long long a = keys[i], b = keys[l];
    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
}

// sorted entry -> (x, y, z, point index)
__global__ void filter_gather_kernel(const float *__restrict__ xyz, const unsigned *__restrict__ idx, const unsigned long long *__restrict__ keys,
                                     unsigned m, float4 *__restrict__ sp) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned id = idx[static_cast<unsigned>(keys[i] & 0xffffffffu)];
    const float *p = xyz + 3ull * id;
    sp[i] = make_float4(p[0], p[1], p[2], __uint_as_float(id));
}

// nxt[i] = first j > i with |sp[j] - sp[i]|^2 > sqdist, or m (filter.hh:249: `> sqdist` keeps)
__global__ void filter_next_kernel(const float4 *__restrict__ sp, unsigned m, float sqdist, unsigned *__restrict__ nxt) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 a = sp[i];
    unsigned j = i + 1;
    while (j < m) {
        const float4 b = __ldg(sp + j);
        if (sq3(b.x, b.y, b.z, a.x, a.y, a.z) > sqdist) break;
        ++j;
    }
    nxt[i] = j;
}

// one thread per segment, back to front: where does a chain that enters at entry i leave the segment?
__global__ void filter_exit_kernel(const unsigned *__restrict__ nxt, unsigned m, unsigned *__restrict__ exit_) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = s * kSeg;
    if (b >= m) return;
    const unsigned e = min(b + kSeg, m);
    for (unsigned i = e; i-- > b;) {
        const unsigned j = nxt[i];
        exit_[i] = j >= e ? j : exit_[j];
    }
}

// one thread: the chain from entry 0 visits the segments in order; record where it enters each (or m: not visited)
__global__ void filter_entry_kernel(const unsigned *__restrict__ exit_, unsigned m, unsigned n_seg, unsigned *__restrict__ entry) {
    for (unsigned s = threadIdx.x; s < n_seg; s += blockDim.x) entry[s] = m;
    __syncthreads();
    if (threadIdx.x != 0) return;
    unsigned cur = 0;
    while (cur < m) {
        entry[cur / kSeg] = cur;
        cur = exit_[cur];
    }
}

// one thread per segment: walk the segment's piece of the chain; WRITE == false counts, WRITE == true emits the kept
// point indices and coordinates in order at the segment's offset
template <bool WRITE>
__global__ void filter_walk_kernel(const unsigned *__restrict__ nxt, const unsigned *__restrict__ entry, const float4 *__restrict__ sp, unsigned m,
                                   unsigned n_seg, unsigned *__restrict__ seg_count, unsigned *__restrict__ idx_out) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const unsigned e = min((s + 1) * kSeg, m);
    unsigned cur = entry[s], c = 0;
    const unsigned off = WRITE ? seg_count[s] : 0u;
    while (cur < e) {
        if (WRITE) idx_out[off + c] = __float_as_uint(sp[cur].w);
        ++c;
        cur = nxt[cur];
    }
    if (!WRITE) seg_count[s] = c;
}

__global__ void filter_emit_kernel(const float *__restrict__ xyz, const unsigned *__restrict__ idx, unsigned m, float *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float *p = xyz + 3ull * idx[i];
    out[3ull * i] = p[0];
    out[3ull * i + 1] = p[1];
    out[3ull * i + 2] = p[2];
}

inline float unfkey(unsigned u) {
    const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}

inline size_t up256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace

extern "C" int b200mp_filter_pointcloud(b200mp_ctx *c, const float *xyz, int64_t n, float min_dist, float max_range, const float *origin,
                                        const float *ws_min, const float *ws_max, int cull, float *out_xyz, int64_t *out_n) {
    if (!c || !out_n || n < 0 || (n > 0 && (!xyz || !out_xyz)) || !origin || !ws_min || !ws_max)
        return fail(B200MP_ERR_INVALID, "b200mp_filter_pointcloud: null argument or negative size");
    if (n >= (int64_t(1) << 31)) return fail(B200MP_ERR_CAPACITY, "b200mp_filter_pointcloud: at most 2^31 - 1 points");
    *out_n = 0;
    if (n == 0) return B200MP_OK;  // filter.hh:185-188
    std::lock_guard<std::mutex> lock(c->mu);
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    // list capacity: every point + the single stand-in for the reference's unshrunk list tail; padded to a power of two >= kSortBlock
    const unsigned cap = static_cast<unsigned>(n) + 1u;
    unsigned cap_pad = kSortBlock;
    while (cap_pad < cap) cap_pad <<= 1;
    const unsigned n_blocks = static_cast<unsigned>((n + 255) / 256);
    const unsigned seg_cap = (cap + kSeg - 1) / kSeg;
    size_t off = 0;
    const size_t o_xyz = off; off += up256(static_cast<size_t>(n) * 12);
    const size_t o_flag = off; off += up256(static_cast<size_t>(n));
    const size_t o_bcnt = off; off += up256(static_cast<size_t>(std::max(n_blocks, seg_cap) + 1) * 4);
    const size_t o_idx = off; off += up256(static_cast<size_t>(cap) * 4);
    const size_t o_idx2 = off; off += up256(static_cast<size_t>(cap) * 4);
    const size_t o_keys = off; off += up256(static_cast<size_t>(cap_pad) * 8);
    const size_t o_sp = off; off += up256(static_cast<size_t>(cap) * 16);
    const size_t o_nxt = off; off += up256(static_cast<size_t>(cap) * 4);
    const size_t o_exit = off; off += up256(static_cast<size_t>(cap) * 4);
    const size_t o_entry = off; off += up256(static_cast<size_t>(seg_cap) * 4);
    const size_t o_misc = off; off += 256;  // [0] min key, [1] max key, [2] total
    const size_t o_out = off; off += up256(static_cast<size_t>(cap) * 12);
    int rc = c->scratch.reserve(off);
    if (rc) return rc;
    char *m8 = c->scratch.ptr;
    float *d_xyz = reinterpret_cast<float *>(m8 + o_xyz);
    unsigned char *d_flag = reinterpret_cast<unsigned char *>(m8 + o_flag);
    unsigned *d_cnt = reinterpret_cast<unsigned *>(m8 + o_bcnt);
    unsigned *d_idx = reinterpret_cast<unsigned *>(m8 + o_idx), *d_idx2 = reinterpret_cast<unsigned *>(m8 + o_idx2);
    unsigned long long *d_keys = reinterpret_cast<unsigned long long *>(m8 + o_keys);
    float4 *d_sp = reinterpret_cast<float4 *>(m8 + o_sp);
    unsigned *d_nxt = reinterpret_cast<unsigned *>(m8 + o_nxt), *d_exit = reinterpret_cast<unsigned *>(m8 + o_exit);
    unsigned *d_entry = reinterpret_cast<unsigned *>(m8 + o_entry), *d_misc = reinterpret_cast<unsigned *>(m8 + o_misc);
    float *d_out = reinterpret_cast<float *>(m8 + o_out);
    long long launches = 0;

    B200_CUDA(cudaMemcpyAsync(d_xyz, xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, st));
    FilterParams fp;
    for (int k = 0; k < 3; ++k) { fp.origin[k] = origin[k]; fp.ws_min[k] = ws_min[k]; fp.ws_max[k] = ws_max[k]; }
    fp.sqrange = max_range * max_range;
    fp.cull = cull;
    filter_cull_kernel<<<n_blocks, 256, 0, st>>>(d_xyz, n, fp, d_flag);
    filter_count_kernel<<<n_blocks, 256, 0, st>>>(d_flag, n, d_cnt);
    filter_scan_kernel<<<1, 1024, 0, st>>>(d_cnt, n_blocks, d_misc + 2);
    filter_scatter_kernel<<<n_blocks, 256, 0, st>>>(d_flag, n, d_cnt, d_idx);
    launches += 4;
    unsigned kept = 0;
    B200_CUDA(cudaMemcpyAsync(&kept, d_misc + 2, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    unsigned m = kept;
    if (static_cast<int64_t>(kept) < n) {  // the reference's list tail: entries that all refer to point 0
        const unsigned zero = 0;
        B200_CUDA(cudaMemcpyAsync(d_idx + m, &zero, 4, cudaMemcpyHostToDevice, st));
        ++m;
    }
    const float sqdist = min_dist * min_dist;
    float mn = std::min({origin[0] - max_range, origin[1] - max_range, origin[2] - max_range});
    float mx = std::min({origin[0] + max_range, origin[1] + max_range, origin[2] + max_range});  // sic (filter.hh:196)
    int perm[3] = {0, 1, 2};
    do {
        unsigned m_pad = kSortBlock;
        while (m_pad < m) m_pad <<= 1;
        const unsigned init[2] = {0xffffffffu, 0u};
        B200_CUDA(cudaMemcpyAsync(d_misc, init, 8, cudaMemcpyHostToDevice, st));
        filter_codes_kernel<<<(m_pad + 255) / 256, 256, 0, st>>>(d_xyz, d_idx, m, m_pad, perm[0], perm[1], perm[2], mn, mx, d_keys, d_misc);
        filter_sort_local_kernel<<<m_pad / kSortBlock, kSortBlock / 2, 0, st>>>(d_keys, 2u, kSortBlock);
        launches += 2;
        for (unsigned k = kSortBlock * 2; k <= m_pad; k <<= 1) {
            for (unsigned j = k >> 1; j >= kSortBlock; j >>= 1) {
                filter_sort_global_kernel<<<(m_pad / 2 + 255) / 256, 256, 0, st>>>(d_keys, m_pad, k, j);
                ++launches;
            }
            filter_sort_local_kernel<<<m_pad / kSortBlock, kSortBlock / 2, 0, st>>>(d_keys, k, k);
            ++launches;
        }
        const unsigned n_seg = (m + kSeg - 1) / kSeg;
        filter_gather_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_xyz, d_idx, d_keys, m, d_sp);
        filter_next_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_sp, m, sqdist, d_nxt);
        filter_exit_kernel<<<(n_seg + 63) / 64, 64, 0, st>>>(d_nxt, m, d_exit);
        filter_entry_kernel<<<1, 256, 0, st>>>(d_exit, m, n_seg, d_entry);
        filter_walk_kernel<false><<<(n_seg + 63) / 64, 64, 0, st>>>(d_nxt, d_entry, d_sp, m, n_seg, d_cnt, nullptr);
        filter_scan_kernel<<<1, 1024, 0, st>>>(d_cnt, n_seg, d_misc + 2);
        filter_walk_kernel<true><<<(n_seg + 63) / 64, 64, 0, st>>>(d_nxt, d_entry, d_sp, m, n_seg, d_cnt, d_idx2);
        launches += 7;
        unsigned back[3];
        B200_CUDA(cudaMemcpyAsync(back, d_misc, 12, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        B200_CUDA(cudaGetLastError());
        const float new_min = std::min(mx, unfkey(back[0])), new_max = std::max(mn, unfkey(back[1]));
        m = back[2];
        std::swap(d_idx, d_idx2);
        mx = static_cast<float>((new_max + mx) / 2.);
        mn = static_cast<float>((new_min + mn) / 2.);
    } while (std::next_permutation(perm, perm + 3));
    filter_emit_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_xyz, d_idx, m, d_out);
    ++launches;
    B200_CUDA(cudaMemcpyAsync(out_xyz, d_out, static_cast<size_t>(m) * 12, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_CUDA(cudaGetLastError());
    c->launches.fetch_add(launches);
    *out_n = m;
    return B200MP_OK;
}
