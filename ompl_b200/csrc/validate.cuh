// Motion / state validation kernels (sm_100a).
//
// One warp per motion: lane l of a pass holds the state of rake (l >> 3), sample (l & 7), i.e. the warp covers 4 of
// the reference's 8-wide "rakes" per pass (external/vamp/src/impl/vamp/planning/validate.hh:23-67). Every lane runs
// the robot's forward kinematics in registers, link by link (generated robots/<robot>_gen.cuh), and tests the link's
// bounding sphere, then its fine spheres, against the scene staged in shared memory with the reference's decision
// semantics (external/vamp/src/impl/vamp/collision/validity.hh:46-150):
//   * obstacles of each class are scanned in stored order; the scan of a sphere stops at the first obstacle whose
//     min_distance is >= the sphere's max_extent in all 8 lanes of its rake  -> here: prefix length b = #{prefmax < ext}
//     with ext taken as the max over the 8-lane group, found by binary search instead of a per-obstacle test;
//   * collision <=> sign bit of (squared distance - squared radius sum)  -> strict `< 0`;
//   * a coarse hit in any lane of a rake triggers the fine tests for the whole rake; a fine hit invalidates the motion.
// Control flow is warp-uniform: loop bounds are warp maxima, lanes mask their own results, exits are warp votes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "robot_common.cuh"

namespace b200mp {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxPointClouds = 4;
constexpr int kMaxHeightFields = 2;

// Height field (external/vamp/src/impl/vamp/collision/shapes.hh:250-312): centre, RECIPROCAL scales, cell counts (as floats,
// the reference does this arithmetic in float), heights row-major in global memory.
struct HeightFieldView {
    float x, y, z, xs, ys, zs;
    float xd, yd, xd2, yd2;
    const float *data;
    unsigned n;
};

// Point-cloud tree in global memory (arrays as external/vamp/src/impl/vamp/collision/capt.hh:571-613 stores them).
struct CaptView {
    const float *tests;
    const float *aabbs;           // [leaf][6]
    const uint32_t *aff_starts;   // [leaves + 1], units of 8-point vectors
    const float *aff;             // [3][n_aff8 * 8]; on the device every leaf's list is ordered by distance from the leaf's cell
    const unsigned short *pref;   // [leaf][kCaptBuckets]: records of the list within capt_reach(k) of the cell (0xffff: whole list), or null
    long long n_aff8;
    float top[6];
    float r_point;
    int nlog2;
    int n_tests;
};

// Uniform grid over the obstacles of a SORTED scene (built at commit, lives in global memory / L2). Level L lists, per
// cell and shape class, the records a sphere of radius <= reach centred in that cell can possibly touch (every obstacle
// whose bounding box inflated by reach overlaps the cell; border cells extend to infinity, so any centre maps to a cell
// that is conservative for it). In a sorted scene the reference's ordered scan stops only at obstacles that cannot
// touch the sphere (min_distance >= max_extent), so "any listed obstacle has a negative test value" is the same
// decision as the full scan, found with a few dozen tests instead of a thousand.
constexpr int kGridLevels = 24;
constexpr int kShapeClasses = 5;  // spheres, capsules, z-capsules, cuboids, z-cuboids
struct GridView {
    float ox, oy, oz, inv_h;
    int nx, ny, nz;
    float reach;
    // per cell and class 8 bytes: x = (where the list goes on << 8) | entries, y = the first two entries (u16 record indices into the
    // staged class arrays; the class's first padding record where the list is shorter) -- most lists end there, so one load answers
    // a cell; entries 2, 3, ... of longer lists follow in `ids` from position x >> 8
    const uint2 *cellw[kShapeClasses];
    const unsigned short *ids[kShapeClasses];
};

// Scene as one packed float blob (16-byte aligned sections). Offsets are in floats.
struct SceneDesc {
    const float *blob;
    int blob_floats;  // multiple of 4
    int n_sph, n_cap, n_zcap, n_cub, n_zcub;
    int off_sph, off_sph_pm;      // float4 (x,y,z,r)                                  ; prefmax[n]
    int off_cap, off_cap_pm;      // 2 x float4 (x1,y1,z1,r) (xv,yv,zv,rdv)
    int off_zcap, off_zcap_pm;    // 2 x float4 (x1,y1,z1,r) (zv,rdv,0,0)
    int off_cub, off_cub_pm;      // 4 x float4 (cx,cy,cz,h1) (a1x,a1y,a1z,h2) (a2x,a2y,a2z,h3) (a3x,a3y,a3z,0)
    int off_zcub, off_zcub_pm;    // 3 x float4 (cx,cy,cz,h3) (a1x,a1y,h1,0) (a2x,a2y,h2,0)
    int sorted;  // 1 when every class is in non-decreasing min_distance order (prefix masks are then redundant)
    int n_pc;
    CaptView pc[kMaxPointClouds];
    int n_hf;
    HeightFieldView hf[kMaxHeightFields];
    int n_grid;  // 0: no grid (unsorted or small scene) -> ordered scans
    GridView grid[kGridLevels];
    // Attachment (collision/attachments.hh): global memory, 16 floats = row-major 3x4 pose of the attachment relative to the robot's
    // end-effector frame (+ 4 pad), then n_att records (x, y, z, r) in the attachment's own frame. n_att == 0: none.
    const float *att;
    int n_att;
};

// Pointers into the staged copy.
struct Scene {
    const float4 *sph; const float *sph_pm; int n_sph;
    const float4 *cap; const float *cap_pm; int n_cap;
    const float4 *zcap; const float *zcap_pm; int n_zcap;
    const float4 *cub; const float *cub_pm; int n_cub;
    const float4 *zcub; const float *zcub_pm; int n_zcub;
};

__device__ __forceinline__ Scene scene_view(const SceneDesc &d, const float *base) {
    Scene s;
    s.sph = reinterpret_cast<const float4 *>(base + d.off_sph); s.sph_pm = base + d.off_sph_pm; s.n_sph = d.n_sph;
    s.cap = reinterpret_cast<const float4 *>(base + d.off_cap); s.cap_pm = base + d.off_cap_pm; s.n_cap = d.n_cap;
    s.zcap = reinterpret_cast<const float4 *>(base + d.off_zcap); s.zcap_pm = base + d.off_zcap_pm; s.n_zcap = d.n_zcap;
    s.cub = reinterpret_cast<const float4 *>(base + d.off_cub); s.cub_pm = base + d.off_cub_pm; s.n_cub = d.n_cub;
    s.zcub = reinterpret_cast<const float4 *>(base + d.off_zcub); s.zcub_pm = base + d.off_zcub_pm; s.n_zcub = d.n_zcub;
    return s;
}

// ------------------------------------------------------------------------------------------------ TMA bulk staging
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Copy the scene blob global -> shared with the bulk async-copy engine (cp.async.bulk, SASS UBLKCP), completion
// tracked by an mbarrier. Called by the whole block; returns after the data has landed and is visible to all threads.
__device__ __forceinline__ void stage_scene(float *dst, const float *src, int n_floats, unsigned long long *bar) {
    const uint32_t bar_a = smem_u32(bar);
    const uint32_t bytes = static_cast<uint32_t>(n_floats) * 4u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && bytes > 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        constexpr uint32_t kChunk = 32768u;
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            const uint32_t n = min(kChunk, bytes - off);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(reinterpret_cast<char *>(dst) + off)), "l"(reinterpret_cast<const char *>(src) + off), "r"(n),
                "r"(bar_a)
                : "memory");
        }
    }
    if (bytes > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a)
                : "memory");
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ helpers
// number of leading entries of the non-decreasing array pm[0..n) that are < ext
__device__ __forceinline__ int count_lt(const float *pm, int n, float ext) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pm[mid] < ext) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int G>
__device__ __forceinline__ float group_max(float v) {
    if (G >= 2) v = fmaxf(v, __shfl_xor_sync(kFull, v, 1));
    if (G >= 4) v = fmaxf(v, __shfl_xor_sync(kFull, v, 2));
    if (G >= 8) v = fmaxf(v, __shfl_xor_sync(kFull, v, 4));
    return v;
}

template <int G>
__device__ __forceinline__ bool group_any(bool p) {
    if (G == 1) return p;
    const unsigned b = __ballot_sync(kFull, p);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gm = ((1u << G) - 1u) << (lane & ~(static_cast<unsigned>(G) - 1u));
    return (b & gm) != 0u;
}

// Exit policy of a scan loop, evaluated every 8 obstacles:
//  kUntilAllHit : stop once every lane has a hit or is not participating (coarse tests; state kernel)
//  kAnyHit      : stop as soon as any participating lane has a hit (fine tests of the motion kernel: the motion is dead)
enum ExitPolicy { kUntilAllHit = 0, kAnyHit = 1 };

template <int POLICY>
__device__ __forceinline__ bool scan_should_stop(bool hit, bool participating) {
    if (POLICY == kAnyHit) return __any_sync(kFull, hit);
    return __all_sync(kFull, hit || !participating);
}

// T robot spheres per lane against every obstacle class.
template <int T>
struct Tile {
    float x[T], y[T], z[T], r[T];
    float ext[T];  // group max_extent, -inf for idle slots / idle lanes
};

// Primitive tests. Each returns d with sign bit set <=> collision. `o` points at the obstacle's float4 records.
struct SphereTest {   // collision/sphere_sphere.hh:10-23
    static constexpr int kRec = 1;
    __device__ static __forceinline__ float test(const float4 *o, float x, float y, float z, float r, float) {
        const float4 c = o[0];
        const float xs = c.x - x, ys = c.y - y, zs = c.z - z;
        const float sum = fmaf(zs, zs, fmaf(ys, ys, xs * xs));
        const float rs = c.w + r;
        return fmaf(-rs, rs, sum);
    }
};
struct CapsuleTest {  // collision/sphere_capsule.hh:9-23; records (x1,y1,z1,r) (xv,yv,zv,rdv)
    static constexpr int kRec = 2;
    __device__ static __forceinline__ float test(const float4 *o, float x, float y, float z, float r, float) {
        const float4 p = o[0], v = o[1];
        const float dot = fmaf(z - p.z, v.z, fmaf(y - p.y, v.y, (x - p.x) * v.x));
        const float cdf = fminf(fmaxf(dot * v.w, 0.0f), 1.0f);
        const float xs = x - fmaf(v.x, cdf, p.x), ys = y - fmaf(v.y, cdf, p.y), zs = z - fmaf(v.z, cdf, p.z);
        const float sum = fmaf(zs, zs, fmaf(ys, ys, xs * xs));
        const float rs = r + p.w;
        return fmaf(-rs, rs, sum);
    }
};
struct ZCapsuleTest {  // collision/sphere_capsule.hh:31-44; records (x1,y1,z1,r) (zv,rdv,0,0)
    static constexpr int kRec = 2;
    __device__ static __forceinline__ float test(const float4 *o, float x, float y, float z, float r, float) {
        const float4 p = o[0], v = o[1];
        const float dot = (z - p.z) * v.x;
        const float cdf = fminf(fmaxf(dot * v.y, 0.0f), 1.0f);
        const float xs = x - p.x, ys = y - p.y, zs = z - fmaf(v.x, cdf, p.z);
        const float sum = fmaf(zs, zs, fmaf(ys, ys, xs * xs));
        const float rs = r + p.w;
        return fmaf(-rs, rs, sum);
    }
};
struct CuboidTest {  // collision/sphere_cuboid.hh:9-27; records (cx,cy,cz,h1) (a1,h2) (a2,h3) (a3,0)
    static constexpr int kRec = 4;
    __device__ static __forceinline__ float test(const float4 *o, float x, float y, float z, float, float rsq) {
        const float4 c = o[0], a1 = o[1], a2 = o[2], a3 = o[3];
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float e1 = fmaxf(fabsf(fmaf(a1.z, zs, fmaf(a1.y, ys, a1.x * xs))) - c.w, 0.0f);
        const float e2 = fmaxf(fabsf(fmaf(a2.z, zs, fmaf(a2.y, ys, a2.x * xs))) - a1.w, 0.0f);
        const float e3 = fmaxf(fabsf(fmaf(a3.z, zs, fmaf(a3.y, ys, a3.x * xs))) - a2.w, 0.0f);
        return fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, -rsq)));
    }
};
struct ZCuboidTest {  // collision/sphere_cuboid.hh:35-52; records (cx,cy,cz,h3) (a1x,a1y,h1,0) (a2x,a2y,h2,0)
    static constexpr int kRec = 3;
    __device__ static __forceinline__ float test(const float4 *o, float x, float y, float z, float, float rsq) {
        const float4 c = o[0], a1 = o[1], a2 = o[2];
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float e1 = fmaxf(fabsf(fmaf(a1.y, ys, a1.x * xs)) - a1.z, 0.0f);
        const float e2 = fmaxf(fabsf(fmaf(a2.y, ys, a2.x * xs)) - a2.z, 0.0f);
        const float e3 = fmaxf(fabsf(zs) - c.w, 0.0f);
        return fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, -rsq)));
    }
};

// One obstacle class. The collision flag is the sign bit of `acc` (OR of the raw bits of every test value, exactly the
// reference's testz sign-bit predicate). Arrays are padded to a multiple of kObsUnroll with far-away dummies, so the
// loop runs whole blocks; one warp vote per block decides whether to stop.
//   MASKED = false (sorted scenes): entries beyond a sphere's own prefix lie farther from the origin than the sphere
//            reaches, their tests are non-negative by construction, so no per-test mask is needed;
//   MASKED = true  (scenes given unsorted): the reference's `break` truncates the scan, so tests at i >= b are dropped.
constexpr int kObsUnroll = 4;
// position given to idle sphere slots: far from every obstacle and from the +1e9 padding records
constexpr float kIdle = -1.0e9f;

template <class Test, int T, int POLICY, bool MASKED>
__device__ __forceinline__ int scan_class(const float4 *__restrict__ obs, const float *__restrict__ pm, int n, const Tile<T> &s,
                                          const float (&rsq)[T], bool participating, int acc) {
    int b[T];
    int bl = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) { b[t] = count_lt(pm, n, s.ext[t]); bl = max(bl, b[t]); }
    const int B = __reduce_max_sync(kFull, bl);
#pragma unroll 1
    for (int i = 0; i < B; i += kObsUnroll) {
#pragma unroll
        for (int u = 0; u < kObsUnroll; ++u) {
            const float4 *o = obs + (i + u) * Test::kRec;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int bits = __float_as_int(Test::test(o, s.x[t], s.y[t], s.z[t], s.r[t], rsq[t]));
                acc |= MASKED ? ((i + u < b[t]) ? bits : 0) : bits;
            }
        }
        if (scan_should_stop<POLICY>(acc < 0, participating)) break;
    }
    return acc;
}

template <int T, int POLICY, bool MASKED>
__device__ __forceinline__ bool scan_scene(const Scene &sc, const Tile<T> &s, bool participating) {
    int acc = 0;
    float rsq[T];
#pragma unroll
    for (int t = 0; t < T; ++t) rsq[t] = s.r[t] * s.r[t];
    if (sc.n_sph > 0) {
        acc = scan_class<SphereTest, T, POLICY, MASKED>(sc.sph, sc.sph_pm, sc.n_sph, s, rsq, participating, acc);
        if (scan_should_stop<POLICY>(acc < 0, participating)) return acc < 0;
    }
    if (sc.n_cap > 0) {
        acc = scan_class<CapsuleTest, T, POLICY, MASKED>(sc.cap, sc.cap_pm, sc.n_cap, s, rsq, participating, acc);
        if (scan_should_stop<POLICY>(acc < 0, participating)) return acc < 0;
    }
    if (sc.n_zcap > 0) {
        acc = scan_class<ZCapsuleTest, T, POLICY, MASKED>(sc.zcap, sc.zcap_pm, sc.n_zcap, s, rsq, participating, acc);
        if (scan_should_stop<POLICY>(acc < 0, participating)) return acc < 0;
    }
    if (sc.n_cub > 0) {
        acc = scan_class<CuboidTest, T, POLICY, MASKED>(sc.cub, sc.cub_pm, sc.n_cub, s, rsq, participating, acc);
        if (scan_should_stop<POLICY>(acc < 0, participating)) return acc < 0;
    }
    if (sc.n_zcub > 0) {
        acc = scan_class<ZCuboidTest, T, POLICY, MASKED>(sc.zcub, sc.zcub_pm, sc.n_zcub, s, rsq, participating, acc);
    }
    return acc < 0;
}

// ------------------------------------------------------------------------------------------------ grid-culled scans
__device__ __forceinline__ int grid_cell(const GridView &g, float x, float y, float z) {
    // monotone in each coordinate; the host maps inflated obstacle boxes to cell ranges with the same expression
    const int cx = min(max(__float2int_rd((x - g.ox) * g.inv_h), 0), g.nx - 1);
    const int cy = min(max(__float2int_rd((y - g.oy) * g.inv_h), 0), g.ny - 1);
    const int cz = min(max(__float2int_rd((z - g.oz) * g.inv_h), 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}

// smallest level whose reach covers radius r, or -1 (radii are per-sphere constants: warp-uniform)
__device__ __forceinline__ int grid_level(const SceneDesc &d, float r) {
    int lv = -1;
#pragma unroll
    for (int l = kGridLevels - 1; l >= 0; --l)
        if (l < d.n_grid && r <= d.grid[l].reach) lv = l;
    return lv;
}

// Exit test of a grid walk, evaluated every 8 candidates. any_mode (warp-uniform): stop as soon as any lane has a hit;
// otherwise stop once every participating lane has a hit. One vote either way: any(hit) = !all(!hit).
__device__ __forceinline__ bool grid_should_stop(bool hit, bool on, bool any_mode) {
    const bool a = __all_sync(kFull, any_mode ? !hit : (hit || !on));
    return a != any_mode;
}

// One shape class: every lane walks the candidate list of its own cell; record `n` (the first padding record) stands in
// once a lane's list is exhausted, so the loop body needs no predication.
template <class Test>
__device__ __forceinline__ int grid_scan_class(const uint2 *__restrict__ cellw, const unsigned short *__restrict__ ids, int cell,
                                               const float4 *__restrict__ obs, int n, float x, float y, float z, float r, float rsq,
                                               bool on, bool any_mode, int acc) {
    const uint2 w = on ? __ldg(cellw + cell) : make_uint2(0u, static_cast<unsigned>(n) * 0x10001u);
    const unsigned cnt = w.x & 255u, b = w.x >> 8;  // (entry i >= 2 of the list sits at ids[b + i - 2])
    const unsigned wmax = __reduce_max_sync(kFull, cnt);
#pragma unroll 1
    for (unsigned i = 0; i < wmax; i += 2) {
        int id0 = static_cast<int>(w.y & 0xffffu), id1 = static_cast<int>(w.y >> 16);
        if (i > 0) {
            id0 = (i < cnt) ? static_cast<int>(__ldg(ids + (b + i - 2u))) : n;
            id1 = (i + 1 < cnt) ? static_cast<int>(__ldg(ids + (b + i - 1u))) : n;
        }
        acc |= __float_as_int(Test::test(obs + id0 * Test::kRec, x, y, z, r, rsq));
        acc |= __float_as_int(Test::test(obs + id1 * Test::kRec, x, y, z, r, rsq));
        if ((i & 6u) == 6u && grid_should_stop(acc < 0, on, any_mode)) break;
    }
    return acc;
}

__device__ __forceinline__ bool grid_scan_scene(const GridView &g, const Scene &sc, float x, float y, float z, float r, bool on, bool any_mode) {
    const int cell = grid_cell(g, x, y, z);
    const float rsq = r * r;
    int acc = 0;
    if (sc.n_sph > 0) {
        acc = grid_scan_class<SphereTest>(g.cellw[0], g.ids[0], cell, sc.sph, sc.n_sph, x, y, z, r, rsq, on, any_mode, acc);
        if (grid_should_stop(acc < 0, on, any_mode)) return acc < 0;
    }
    if (sc.n_cap > 0) {
        acc = grid_scan_class<CapsuleTest>(g.cellw[1], g.ids[1], cell, sc.cap, sc.n_cap, x, y, z, r, rsq, on, any_mode, acc);
        if (grid_should_stop(acc < 0, on, any_mode)) return acc < 0;
    }
    if (sc.n_zcap > 0) {
        acc = grid_scan_class<ZCapsuleTest>(g.cellw[2], g.ids[2], cell, sc.zcap, sc.n_zcap, x, y, z, r, rsq, on, any_mode, acc);
        if (grid_should_stop(acc < 0, on, any_mode)) return acc < 0;
    }
    if (sc.n_cub > 0) {
        acc = grid_scan_class<CuboidTest>(g.cellw[3], g.ids[3], cell, sc.cub, sc.n_cub, x, y, z, r, rsq, on, any_mode, acc);
        if (grid_should_stop(acc < 0, on, any_mode)) return acc < 0;
    }
    if (sc.n_zcub > 0) acc = grid_scan_class<ZCuboidTest>(g.cellw[4], g.ids[4], cell, sc.zcub, sc.n_zcub, x, y, z, r, rsq, on, any_mode, acc);
    return acc < 0;
}

// Coarse walk that REMEMBERS what it touches. A link's fine spheres lie inside its bounding sphere (to 1e-8 m, checked when the
// tables are generated), so a fine sphere can only touch obstacles the bounding sphere touches: the walk tests every candidate of
// the cell against the bounding sphere inflated by kHitMargin (far above any rounding) and lists the hits -- (class << 16) | record
// -- in the lane's shared-memory slots; the fine spheres are then tested against that short list directly, without grid walks of
// their own. A bounding sphere within kHitMargin of an obstacle counts as a coarse hit here where the reference's exact test may
// not; the fine tests it triggers are exact and cannot hit unless a fine sphere really intersects, so decisions are unchanged.
// Returns the number of hits, kHitCap + 1 meaning "more than fit": the caller then falls back to per-sphere walks.
constexpr float kHitMargin = 1.0e-4f;
constexpr int kHitCap = 8;  // list entries per lane

template <class Test, int BT>
__device__ __forceinline__ int grid_collect_class(uint2 w, const unsigned short *__restrict__ ids, const float4 *__restrict__ obs, int n, int cls,
                                                  float x, float y, float z, float r, float rsq, float *__restrict__ hits, int cnt) {
    const unsigned c = w.x & 255u, b = w.x >> 8;  // (entry i >= 2 of the list sits at ids[b + i - 2])
    const unsigned wmax = __reduce_max_sync(kFull, c);
    int id0 = static_cast<int>(w.y & 0xffffu), id1 = static_cast<int>(w.y >> 16);  // the first two candidates came with the cell word
#pragma unroll 1
    for (unsigned i = 0; i < wmax; i += 2) {
        const float v0 = Test::test(obs + id0 * Test::kRec, x, y, z, r, rsq);
        const float v1 = Test::test(obs + id1 * Test::kRec, x, y, z, r, rsq);
        if (v0 < 0.0f) {  // (record n, the far-away stand-in of exhausted lists, never hits)
            if (cnt < kHitCap) hits[cnt * BT] = __int_as_float((cls << 16) | id0);
            cnt = min(cnt + 1, kHitCap + 1);
        }
        if (v1 < 0.0f) {
            if (cnt < kHitCap) hits[cnt * BT] = __int_as_float((cls << 16) | id1);
            cnt = min(cnt + 1, kHitCap + 1);
        }
        id0 = (i + 2 < c) ? static_cast<int>(__ldg(ids + (b + i))) : n;
        id1 = (i + 3 < c) ? static_cast<int>(__ldg(ids + (b + i + 1u))) : n;
    }
    return cnt;
}

template <int BT>
__device__ __forceinline__ int grid_collect_scene(const GridView &g, const Scene &sc, float x, float y, float z, float r, bool on,
                                                  float *__restrict__ hits) {
    const int cell = grid_cell(g, x, y, z);
    r += kHitMargin;
    const float rsq = r * r;
    // the cell words of every class first (independent loads in flight together), then class by class
    uint2 w[kShapeClasses];
    const int ns[kShapeClasses] = {sc.n_sph, sc.n_cap, sc.n_zcap, sc.n_cub, sc.n_zcub};
#pragma unroll
    for (int c = 0; c < kShapeClasses; ++c) {
        w[c] = make_uint2(0u, static_cast<unsigned>(ns[c]) * 0x10001u);
        if (ns[c] > 0 && on) w[c] = __ldg(g.cellw[c] + cell);
    }
    int cnt = 0;
    if (sc.n_sph > 0) cnt = grid_collect_class<SphereTest, BT>(w[0], g.ids[0], sc.sph, sc.n_sph, 0, x, y, z, r, rsq, hits, cnt);
    if (sc.n_cap > 0) cnt = grid_collect_class<CapsuleTest, BT>(w[1], g.ids[1], sc.cap, sc.n_cap, 1, x, y, z, r, rsq, hits, cnt);
    if (sc.n_zcap > 0) cnt = grid_collect_class<ZCapsuleTest, BT>(w[2], g.ids[2], sc.zcap, sc.n_zcap, 2, x, y, z, r, rsq, hits, cnt);
    if (sc.n_cub > 0) cnt = grid_collect_class<CuboidTest, BT>(w[3], g.ids[3], sc.cub, sc.n_cub, 3, x, y, z, r, rsq, hits, cnt);
    if (sc.n_zcub > 0) cnt = grid_collect_class<ZCuboidTest, BT>(w[4], g.ids[4], sc.zcub, sc.n_zcub, 4, x, y, z, r, rsq, hits, cnt);
    return cnt;
}

// TWO fine spheres against the lane's list of obstacles (`mine` <= kHitCap entries, hmax = the warp's longest list): an obstacle's
// record is read once for both. acc0 / acc1 collect the raw bits of the test values (sign bit <=> collision).
template <int BT>
__device__ __forceinline__ void hits_scan2(const Scene &sc, const float *__restrict__ hits, int mine, int hmax, float x0, float y0, float z0, float r0,
                                           float x1, float y1, float z1, float r1, bool any_mode, int &acc0, int &acc1) {
    const float rsq0 = r0 * r0, rsq1 = r1 * r1;
#pragma unroll 1
    for (int h = 0; h < hmax; ++h) {
        if (h < mine) {
            const int ent = __float_as_int(hits[h * BT]);
            const int id = ent & 0xffff;
            float v0, v1;
            switch (ent >> 16) {
                case 0: { const float4 *o = sc.sph + id * SphereTest::kRec; v0 = SphereTest::test(o, x0, y0, z0, r0, rsq0); v1 = SphereTest::test(o, x1, y1, z1, r1, rsq1); break; }
                case 1: { const float4 *o = sc.cap + id * CapsuleTest::kRec; v0 = CapsuleTest::test(o, x0, y0, z0, r0, rsq0); v1 = CapsuleTest::test(o, x1, y1, z1, r1, rsq1); break; }
                case 2: { const float4 *o = sc.zcap + id * ZCapsuleTest::kRec; v0 = ZCapsuleTest::test(o, x0, y0, z0, r0, rsq0); v1 = ZCapsuleTest::test(o, x1, y1, z1, r1, rsq1); break; }
                case 3: { const float4 *o = sc.cub + id * CuboidTest::kRec; v0 = CuboidTest::test(o, x0, y0, z0, r0, rsq0); v1 = CuboidTest::test(o, x1, y1, z1, r1, rsq1); break; }
                default: { const float4 *o = sc.zcub + id * ZCuboidTest::kRec; v0 = ZCuboidTest::test(o, x0, y0, z0, r0, rsq0); v1 = ZCuboidTest::test(o, x1, y1, z1, r1, rsq1); break; }
            }
            acc0 |= __float_as_int(v0);
            acc1 |= __float_as_int(v1);
        }
        if (any_mode && __any_sync(kFull, (acc0 | acc1) < 0)) break;
    }
}

// ------------------------------------------------------------------------------------------------ point clouds
__device__ __forceinline__ float aabb_distsq(const float *bb, float x, float y, float z) {
    const float d0 = x - fminf(fmaxf(x, bb[0]), bb[3]);
    const float d1 = y - fminf(fmaxf(y, bb[1]), bb[4]);
    const float d2 = z - fminf(fmaxf(z, bb[2]), bb[5]);
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// A leaf's affordance list holds every point within r_max + r_point of the leaf's cell (capt.hh:123-285). A query sphere
// whose centre lies in the cell can only touch points no farther from the cell than its own radius, so the device copy of
// the tree keeps each list ordered by distance from the cell and a small table of prefix lengths for a ladder of radii:
// a fine sphere of 3 cm then tests a few dozen points instead of the ~1000 the list holds for the 24 cm bounding spheres.
// Same decisions as the full scan (the skipped points cannot be within reach).
constexpr int kCaptBuckets = 8;
__host__ __device__ constexpr float capt_reach(int k) {
    float v = 0.015f;
    for (int i = 0; i < k; ++i) v *= 1.5f;
    return v;
}

// One sphere per lane against one tree (collision/capt.hh:366-407 / :425-506). Per lane: top AABB reject, nlog2 dependent
// loads down the implicit tree, leaf AABB reject. The lists of the lanes that are left are then scanned by the whole warp,
// one list at a time with coalesced loads (lane j takes points j, j + 32, ...); hit <=> distsq <= (r + r_point)^2.
// Must be called by all 32 lanes; r is warp-uniform. any_mode: stop at the first lane found in collision.
__device__ __forceinline__ bool capt_collides_warp(const CaptView &t, float x, float y, float z, float r, bool on, bool any_mode) {
    r += t.r_point;
    const float rsq = r * r;
    int bucket = kCaptBuckets;
    const float reach = r * 1.0001f + 1.0e-6f;  // slack for the rounding of the stored cell distances
#pragma unroll
    for (int k = kCaptBuckets - 1; k >= 0; --k)
        if (reach <= capt_reach(k)) bucket = k;
    bool pend = on && !(aabb_distsq(t.top, x, y, z) > rsq);
    unsigned pb = 0, pe = 0;
    if (pend) {
        unsigned idx = 0;
        float c0 = x, c1 = y, c2 = z;
        for (int i = 0; i < t.nlog2; ++i) {
            idx = 2u * idx + 1u + (c0 >= __ldg(t.tests + idx) ? 1u : 0u);
            const float tmp = c0; c0 = c1; c1 = c2; c2 = tmp;  // cycle x, y, z
        }
        const unsigned leaf = idx - static_cast<unsigned>(t.n_tests);
        const float *bb = t.aabbs + 6u * leaf;
        float lb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) lb[k] = __ldg(bb + k);
        if (aabb_distsq(lb, x, y, z) > rsq) {
            pend = false;
        } else {
            const unsigned beg = __ldg(t.aff_starts + leaf);
            unsigned n = __ldg(t.aff_starts + leaf + 1) - beg;
            if (t.pref != nullptr && bucket < kCaptBuckets) {
                const unsigned pf = __ldg(t.pref + static_cast<size_t>(leaf) * kCaptBuckets + bucket);
                if (pf != 0xffffu) n = min(n, pf);
            }
            pb = beg * 8u;
            pe = pb + n * 8u;
            pend = n > 0;
        }
    }
    const float *ax = t.aff, *ay = t.aff + t.n_aff8 * 8, *az = t.aff + 2 * t.n_aff8 * 8;
    const unsigned lane = threadIdx.x & 31u;
    bool hit = false;
    unsigned todo = __ballot_sync(kFull, pend);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float qx = __shfl_sync(kFull, x, src), qy = __shfl_sync(kFull, y, src), qz = __shfl_sync(kFull, z, src);
        const unsigned sb = __shfl_sync(kFull, pb, src), se = __shfl_sync(kFull, pe, src);
        bool found = false;
#pragma unroll 1
        for (unsigned base = sb; base < se; base += 128u) {
            bool h = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned i = base + 32u * u + lane;
                if (i < se) {
                    const float dx = __ldg(ax + i) - qx, dy = __ldg(ay + i) - qy, dz = __ldg(az + i) - qz;
                    h |= fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= rsq;
                }
            }
            if (__any_sync(kFull, h)) { found = true; break; }
        }
        if (static_cast<int>(lane) == src) hit = found;
        if (any_mode && found) break;
    }
    return hit;
}

// every tree for one sphere, inlined: the ordered-scan variant has exactly two sites of it (a link's bounding sphere, and ONE rolled
// loop over its fine spheres) -- the tree query is ~250 instructions, five inlined copies of it (a bounding sphere + a tile of four fine
// spheres) had pushed the Fetch kernel's hot code to 40 KB, past the instruction cache. `pcs`: the block's shared copy of the views.
__device__ __forceinline__ bool pointclouds_inline(const CaptView *pcs, int n_pc, float x, float y, float z, float r, bool on, bool any_mode) {
    bool hit = false;
    for (int p = 0; p < n_pc; ++p) hit |= capt_collides_warp(pcs[p], x, y, z, r, on && !hit, any_mode);
    return hit;
}

// ------------------------------------------------------------------------------------------------ height fields
// collision/sphere_heightfield.hh:9-30: the cell under the sphere's centre (index clamped to [0, xd] x [0, yd] as the reference
// does; a read past the array, undefined there, takes the last element here); collision <=> z - r below the scaled height.
__device__ __forceinline__ bool heightfields_hit(const HeightFieldView *hf, int n_hf, float x, float y, float z, float r) {
    bool hit = false;
    for (int i = 0; i < n_hf; ++i) {
        const HeightFieldView &a = hf[i];
        const float xo = a.x - x, yo = a.y - y;
        const float xs = floorf(fminf(fmaxf(fmaf(a.xs, xo, a.xd2), 0.f), a.xd));
        const float ys = floorf(fminf(fmaxf(fmaf(a.ys, yo, a.yd2), 0.f), a.yd));
        const unsigned idx = min(static_cast<unsigned>(max(__float2int_rn(fmaf(ys, a.xd, xs)), 0)), a.n - 1u);
        const float zhs = fmaf(a.zs, __ldg(a.data + idx), a.z);
        hit |= ((z - r) - zhs) < 0.0f;
    }
    return hit;
}
__device__ __forceinline__ bool heightfields_hit(const SceneDesc &sd, float x, float y, float z, float r) {
    return heightfields_hit(sd.hf, sd.n_hf, x, y, z, r);
}

// ------------------------------------------------------------------------------------------------ one state per lane
// Per-thread scratch in shared memory: slot s of thread t lives at base[s * BT + t] (bank-conflict free). Slots
// 3l .. 3l+2 hold the world-space centre of link l's bounding sphere, then sin[dim] and cos[dim] of the joint angles.
// ... followed by kHitCap slots listing the obstacles the current link's bounding sphere touches (see grid_collect_scene).
// ... followed by kScratchSlots slots: the hit list (kHitCap entries) while a link is tested against the environment, the saved frame
// of the first link while a self-collision pair is refined.
constexpr int kScratchSlots = 12;
static_assert(kHitCap <= kScratchSlots, "the hit list lives in the scratch slots");
template <class R>
__host__ __device__ constexpr int lane_slots() { return 3 * R::kLinks + 2 * R::kDim + kScratchSlots; }
template <class R>
__host__ __device__ constexpr int trig_slot() { return 3 * R::kLinks; }
template <class R>
__host__ __device__ constexpr int hit_slot() { return 3 * R::kLinks + 2 * R::kDim; }
template <class R>
__host__ __device__ constexpr unsigned revolute_mask() {
    unsigned m = 0;
    for (int j = 0; j < R::kDim; ++j) m |= R::joint_revolute(j) ? (1u << j) : 0u;
    return m;
}

// Grid level of every robot sphere radius for this scene (computed once per block after staging).
template <class R>
struct SphereLevels {
    signed char coarse[R::kLinks];
    signed char fine[R::kFine];
};

template <class R>
__device__ __forceinline__ void fill_levels(const SceneDesc &sd, SphereLevels<R> &lv) {
    const typename R::Tables &tb = R::tables();
    for (int i = threadIdx.x; i < R::kLinks; i += blockDim.x) lv.coarse[i] = static_cast<signed char>(grid_level(sd, tb.coarse[i].w));
    for (int i = threadIdx.x; i < R::kFine; i += blockDim.x) lv.fine[i] = static_cast<signed char>(grid_level(sd, tb.fine[i].w));
}

// sin / cos of every joint angle with the reference's polynomials, left in the thread's shared-memory slots. The 2 * dim
// evaluations run as one rolled loop (the code footprint of the kernel matters more than the few LDS / STS).
template <class R, int BT>
__device__ __forceinline__ void trig_to_slots(const float (&q)[R::kDim], float *__restrict__ slots) {
    constexpr int D = R::kDim, T0 = trig_slot<R>();
    constexpr float PI = 3.14159265359;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        // cos(x) = sin(wrap(x + pi/2)) with the reference's float constants (vector/interface.hh:451-459)
        const float v = q[j] + static_cast<float>(PI / 2.);
        slots[(T0 + j) * BT] = q[j];
        slots[(T0 + D + j) * BT] = v - ((v >= PI) ? static_cast<float>(2 * PI) : 0.0f);
    }
#pragma unroll 1
    for (int i = 0; i < 2 * D; ++i) {
        const int j = i < D ? i : i - D;
        const float q_of_slot = slots[(T0 + i) * BT];
        float r = vamp_sin(q_of_slot);
        if (!((revolute_mask<R>() >> j) & 1u)) r = i < D ? q_of_slot : 1.0f;  // prismatic joints: the frame code uses q itself (kept in the sin slot)
        slots[(T0 + i) * BT] = r;
    }
}

// one step of the streamed frame, k - 1 -> k, through the robot's single out-of-line copy of the frame code (R::step); the joint of
// frame k is joint k - 1, its sin / cos (or, for a prismatic joint, its value) come from the thread's slots
// INL: the inlined copy (the ordered-scan variant, whose hot loops are the obstacle scans: its three sites may each have one)
template <class R, int BT, bool INL = false>
__device__ __forceinline__ void frame_step(Frame &F, int k, const float *__restrict__ slots) {
    const float s = slots[(trig_slot<R>() + k - 1) * BT], c = slots[(trig_slot<R>() + R::kDim + k - 1) * BT];
    if (INL) F = R::step_inline(k, F, s, c);
    else F = R::step(k, F, s, c);
}

// bring the streamed frame F (currently frame `have`, -1 = nothing) to frame k by walking k's chain (arm): either on
// from `have` when that is an earlier frame of the same chain, or from the world frame at the chain's first joint
template <class R, int BT, bool INL = false>
__device__ __forceinline__ void frame_seek(Frame &F, int &have, int k, const float *__restrict__ slots) {
    if (have == k) return;
    const int first = R::tables().frame_branch[k];
    int from = have + 1;
    if (have < first || have > k) { frame_identity(F); from = first; }
#pragma unroll 1
    for (int j = max(from, 1); j <= k; ++j) frame_step<R, BT, INL>(F, j, slots);
    have = k;
}

// Allowed self-collision pairs on the stored centres (collision/validity.hh:13-44), run from the SAME task table as the
// candidate-grid variant (check_state_grid holds the commented original of this loop): pair groups ruled out by one test, the frame of the
// link with fewer fine spheres parked in the scratch slots, the fine tests in the second link's frame. Used by the ordered-scan variant.
// Updates `alive`; returns true when the caller should return `result` immediately.
// KILL (lane-per-state form only): lanes are grouped into motions by `kill` (the lanes that belong to the same motion as this
// one); a hit in any of them ends the work of all of them.
// SLICED (the resident single-call kernel, session.cuh): several warps check the SAME states, each a share of the links and pairs
// (index % nslice == slice); a state is valid when every share says so. The unsliced kernels carry no trace of it (empty type).
template <bool SLICED>
struct Share {
    __device__ __forceinline__ bool mine(unsigned) const { return true; }
};
template <>
struct Share<true> {
    int slice, nslice;
    __device__ __forceinline__ bool mine(unsigned i) const { return static_cast<int>(i % static_cast<unsigned>(nslice)) == slice; }
};
template <class R, int G, bool EDGE, bool KILL, int BT, bool SLICED = false>
__device__ __forceinline__ bool self_collision(float *__restrict__ slots, bool &alive, bool &result, unsigned kill, Share<SLICED> share = {}) {
    const typename R::Tables &tb = R::tables();
    float *scratch = slots + hit_slot<R>() * BT;
    Frame F;
    int have = -1;
    bool need = false;
#pragma unroll 1
    for (int t = R::kFrames + R::kLinks; t < R::kTasks; ++t) {
        const RobotTask tk = tb.tasks[t];
        const int kind = static_cast<int>(tk.kind);
        if (kind == kTaskGroup) {
            const int la = tk.a, g = tk.b, lr = tb.group_ref[g];
            const float xs = slots[(3 * la + 0) * BT] - slots[(3 * lr + 0) * BT], ys = slots[(3 * la + 1) * BT] - slots[(3 * lr + 1) * BT],
                        zs = slots[(3 * la + 2) * BT] - slots[(3 * lr + 2) * BT];
            const float rs = tb.coarse[la].w + tb.group_r[g];
            const float d = fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs)));
            if (!__any_sync(kFull, (d < 0.0f) && alive)) t += tb.group_skip[g];
            continue;
        }
        if (SLICED && !share.mine(tk.a + 3u * tk.b)) continue;
        const int la = tk.a, lb = tk.b;
        if (kind == kTaskSelfA) {
            const float xs = slots[(3 * la + 0) * BT] - slots[(3 * lb + 0) * BT], ys = slots[(3 * la + 1) * BT] - slots[(3 * lb + 1) * BT],
                        zs = slots[(3 * la + 2) * BT] - slots[(3 * lb + 2) * BT];
            const float rs = tb.coarse[la].w + tb.coarse[lb].w;
            const float d = fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs)));
            need = group_any<G>((d < 0.0f) && alive) && alive;
            if (!__any_sync(kFull, need)) { ++t; continue; }
        }
        frame_seek<R, BT, true>(F, have, static_cast<int>(tk.frame), slots);
        if (kind == kTaskSelfA) {
#pragma unroll
            for (int i = 0; i < 9; ++i) scratch[i * BT] = F.r[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) scratch[(9 + i) * BT] = F.t[i];
            continue;
        }
        Frame Rel;
        {
            Frame A;
#pragma unroll
            for (int i = 0; i < 9; ++i) A.r[i] = scratch[i * BT];
#pragma unroll
            for (int i = 0; i < 3; ++i) A.t[i] = scratch[(9 + i) * BT] - F.t[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) Rel.r[3 * i + j] = fmaf(F.r[6 + i], A.r[6 + j], fmaf(F.r[3 + i], A.r[3 + j], F.r[i] * A.r[j]));
                Rel.t[i] = fmaf(F.r[6 + i], A.t[2], fmaf(F.r[3 + i], A.t[1], F.r[i] * A.t[0]));
            }
        }
        const float4 cb = tb.coarse[lb];
        const int ae = tb.link_begin[la + 1], bb = tb.link_begin[lb], be = tb.link_begin[lb + 1];
        bool fhit = false;
#pragma unroll 1
        for (int i = tb.link_begin[la]; i < ae; ++i) {
            const float4 al = tb.fine[i];
            float qx, qy, qz;
            frame_apply(Rel, al.x, al.y, al.z, qx, qy, qz);
            const float xs0 = qx - cb.x, ys0 = qy - cb.y, zs0 = qz - cb.z;
            const float rs0 = al.w + cb.w + 1.0e-4f;
            const bool cand = (fmaf(-rs0, rs0, fmaf(zs0, zs0, fmaf(ys0, ys0, xs0 * xs0))) < 0.0f) && need && !fhit;
            if (!__any_sync(kFull, cand)) continue;
            int acc = 0;
#pragma unroll 2
            for (int j = bb; j < be; ++j) {
                const float4 fl = tb.fine[j];
                const float xs2 = qx - fl.x, ys2 = qy - fl.y, zs2 = qz - fl.z;
                const float rs2 = al.w + fl.w;
                acc |= __float_as_int(fmaf(-rs2, rs2, fmaf(zs2, zs2, fmaf(ys2, ys2, xs2 * xs2))));
            }
            fhit |= (acc < 0) && cand;
            if (EDGE) {
                if (__any_sync(kFull, fhit)) { result = !fhit; return true; }
            } else {
                if (KILL) fhit = (__ballot_sync(kFull, fhit) & kill) != 0u;
                if (__all_sync(kFull, fhit || !need)) break;
            }
        }
        if (fhit) alive = false;
        if (!EDGE && !__any_sync(kFull, alive)) { result = false; return true; }
    }
    return false;
}

// phase A of the ordered-scan variant: trig into the slots, then one unrolled base -> tip pass that parks the bounding-sphere centres
// and leaves the last frame in `tip` (the candidate-grid variant runs the same steps from its task table instead, check_state_grid)
template <class R, int BT>
__device__ __forceinline__ void fk_to_slots(const float (&q)[R::kDim], float *__restrict__ slots, Frame &tip) {
    // (the ordered-scan variant: its hot loops are the obstacle scans, so the forward pass may be the generated straight-line code)
    trig_to_slots<R, BT>(q, slots);
    float sn[R::kDim], cs[R::kDim];
#pragma unroll
    for (int j = 0; j < R::kDim; ++j) {
        sn[j] = slots[(trig_slot<R>() + j) * BT];
        cs[j] = slots[(trig_slot<R>() + R::kDim + j) * BT];
    }
    // (prismatic joints keep q in their sin slot; the generated frames read q itself)
    R::fk_coarse(q, sn, cs, tip, [&](int l, float x, float y, float z) {
        slots[(3 * l + 0) * BT] = x;
        slots[(3 * l + 1) * BT] = y;
        slots[(3 * l + 2) * BT] = z;
    });
}

// Returns true when this lane's state is collision free. `active` lanes take part; idle lanes only follow the
// warp-uniform control flow. G = lanes per reference rake (8 for motions, 1 for single states).
// EDGE: all lanes belong to one motion, so the first hit anywhere ends the whole warp's work.
//
// Phase A runs the forward kinematics once, base -> tip, and parks the bounding-sphere centres in shared memory.
// Phase B tests them against the scene tip -> base (the reference's order: distal links collide first, so colliding
// states exit early); only when a rake has a coarse hit is the link's frame rebuilt (chain walk from the cached
// sin / cos) and its fine spheres tested. Phase C: self_collision().
// This is the ordered-scan variant (scenes without a candidate grid: small or unsorted).
template <class R, int G, bool EDGE, bool KILL, bool MASKED, int BT, bool SLICED = false>
__device__ bool check_state_ordered(const float (&q)[R::kDim], const SceneDesc &sd, const Scene &sc, float *__restrict__ slots, const CaptView *pcs,
                                    bool active, unsigned kill, Share<SLICED> share = {}) {
    constexpr int FT = 4;  // fine spheres processed together per lane
    constexpr int POLICY = EDGE ? kAnyHit : kUntilAllHit;
    const typename R::Tables &tb = R::tables();
    bool alive = active;  // false once this lane's state is known to collide (or lane idle)
    Frame F;
    fk_to_slots<R, BT>(q, slots, F);
    int have = R::kFrames - 1;  // frame held in F (warp-uniform)
#pragma unroll 1
    for (int l = R::kLinks - 1; l >= 0; --l) {
        if (!share.mine(static_cast<unsigned>(l))) continue;
        const float4 cl = tb.coarse[l];
        Tile<1> c;
        c.x[0] = slots[(3 * l + 0) * BT];
        c.y[0] = slots[(3 * l + 1) * BT];
        c.z[0] = slots[(3 * l + 2) * BT];
        c.r[0] = cl.w;
        const float ext = alive ? sqrtf(fmaf(c.z[0], c.z[0], fmaf(c.y[0], c.y[0], c.x[0] * c.x[0]))) + cl.w : -INFINITY;
        c.ext[0] = group_max<G>(ext);
        if (!alive) { c.x[0] = kIdle; c.y[0] = kIdle; c.z[0] = kIdle; }  // idle lanes must never raise the sign bit
        bool chit = scan_scene<1, kUntilAllHit, MASKED>(sc, c, alive);
        if (sd.n_hf > 0 && alive && !chit) chit = heightfields_hit(sd, c.x[0], c.y[0], c.z[0], cl.w);
        if (sd.n_pc > 0) chit |= pointclouds_inline(pcs, sd.n_pc, c.x[0], c.y[0], c.z[0], cl.w, alive && !chit, false);
        const bool need = group_any<G>(chit && alive) && alive;
        if (!__any_sync(kFull, need)) continue;
        frame_seek<R, BT, true>(F, have, tb.link_frame[l], slots);
        const int fb = tb.link_begin[l], fe = tb.link_begin[l + 1];
        bool fhit = false, settled = false;  // settled: every lane that takes part has its hit -- nothing left to find for this link
        // fine spheres against the primitives and height fields, four at a time (skipped when the scene has neither: point clouds only)
        const bool prims = (sc.n_sph | sc.n_cap | sc.n_zcap | sc.n_cub | sc.n_zcub) != 0 || sd.n_hf > 0;
#pragma unroll 1
        for (int f0 = fb; prims && f0 < fe; f0 += FT) {
            Tile<FT> s;
#pragma unroll
            for (int t = 0; t < FT; ++t) {
                const int fi = min(f0 + t, fe - 1);
                const float4 fl = tb.fine[fi];
                frame_apply(F, fl.x, fl.y, fl.z, s.x[t], s.y[t], s.z[t]);
                s.r[t] = fl.w;
                const bool on = need && (f0 + t < fe);
                const float e = on ? sqrtf(fmaf(s.z[t], s.z[t], fmaf(s.y[t], s.y[t], s.x[t] * s.x[t]))) + fl.w : -INFINITY;
                s.ext[t] = group_max<G>(e);
                if (!on) { s.ext[t] = -INFINITY; s.x[t] = kIdle; s.y[t] = kIdle; s.z[t] = kIdle; }
            }
            fhit |= scan_scene<FT, POLICY, MASKED>(sc, s, need && !fhit);
            if (sd.n_hf > 0) {
#pragma unroll
                for (int t = 0; t < FT; ++t)
                    if (need && !fhit && s.ext[t] > -1e30f) fhit = heightfields_hit(sd, s.x[t], s.y[t], s.z[t], s.r[t]);
            }
            if (EDGE) {
                if (__any_sync(kFull, fhit)) return !fhit;
            } else {
                if (KILL) fhit = (__ballot_sync(kFull, fhit) & kill) != 0u;
                if (__all_sync(kFull, fhit || !need)) { settled = true; break; }
            }
        }
        // ... and against the point-cloud trees, one at a time (one inlined copy of the tree query)
        if (sd.n_pc > 0 && !settled) {
#pragma unroll 1
            for (int f = fb; f < fe; ++f) {
                const float4 fl = tb.fine[f];
                float x, y, z;
                frame_apply(F, fl.x, fl.y, fl.z, x, y, z);
                const bool on = need && !fhit;
                fhit |= pointclouds_inline(pcs, sd.n_pc, x, y, z, fl.w, on, POLICY == kAnyHit) && on;
                if (EDGE) {
                    if (__any_sync(kFull, fhit)) return !fhit;
                } else {
                    if (KILL) fhit = (__ballot_sync(kFull, fhit) & kill) != 0u;
                    if (__all_sync(kFull, fhit || !need)) break;
                }
            }
        }
        if (fhit) alive = false;
        if (!EDGE && !__any_sync(kFull, alive)) return false;
    }
    bool result = false;
    if (self_collision<R, G, EDGE, KILL, BT, SLICED>(slots, alive, result, kill, share)) return result;
    return alive;
}

// Rare fall-backs of the grid variant, kept out of line so that they do not sit between the hot loops in the instruction
// cache: robot spheres larger than the widest grid reach take the ordered scan; point clouds.
template <int G>
__device__ __noinline__ bool ordered_scan_one(const Scene *scp, float x, float y, float z, float r, bool on, bool any_mode) {
    const Scene sc = *scp;
    Tile<1> s1;
    s1.x[0] = x; s1.y[0] = y; s1.z[0] = z; s1.r[0] = r;
    const float e = on ? sqrtf(fmaf(z, z, fmaf(y, y, x * x))) + r : -INFINITY;
    s1.ext[0] = group_max<G>(e);
    if (!on) { s1.x[0] = kIdle; s1.y[0] = kIdle; s1.z[0] = kIdle; }
    return any_mode ? scan_scene<1, kAnyHit, false>(sc, s1, on) : scan_scene<1, kUntilAllHit, false>(sc, s1, on);
}

static __device__ __noinline__ bool pointclouds_one(const CaptView *pcs, int n_pc, float x, float y, float z, float r, bool on, bool any_mode) {
    bool hit = false;
    for (int p = 0; p < n_pc; ++p) hit |= capt_collides_warp(pcs[p], x, y, z, r, on && !hit, any_mode);
    return hit;
}

// Block-shared copies of what the out-of-line fall-backs need (taking the address of a kernel parameter or of a register
// struct would force a stack copy): the staged scene view and the point-cloud views.
struct SharedViews {
    Scene sc;
    CaptView pc[kMaxPointClouds];
    GridView grid[kGridLevels];
    HeightFieldView hf[kMaxHeightFields];
    int n_pc, n_grid, n_hf, n_att;
    const float *att;
};
__device__ __forceinline__ void fill_views(const SceneDesc &sd, const Scene &sc, SharedViews &sv) {
    if (threadIdx.x == 0) {
        sv.sc = sc;
        for (int p = 0; p < sd.n_pc; ++p) sv.pc[p] = sd.pc[p];
        for (int l = 0; l < sd.n_grid; ++l) sv.grid[l] = sd.grid[l];
        for (int h = 0; h < sd.n_hf; ++h) sv.hf[h] = sd.hf[h];
        sv.n_pc = sd.n_pc; sv.n_grid = sd.n_grid; sv.n_hf = sd.n_hf; sv.n_att = sd.n_att;
        sv.att = sd.att;
    }
}

__device__ __forceinline__ bool capt_any(const SharedViews &sv, float x, float y, float z, float r, bool on) {
    return pointclouds_one(sv.pc, sv.n_pc, x, y, z, r, on, false);
}

// the per-sphere grid walk as a rare fall-back of the hit-list path (a lane's list overflowed): out of line for the same reason
static __device__ __noinline__ bool grid_scan_scene_cold(const GridView *g, const Scene *scp, float x, float y, float z, float r, bool on, bool any_mode) {
    return grid_scan_scene(*g, *scp, x, y, z, r, on, any_mode);
}

// Everything rare about a link's bounding sphere, out of line: height fields and point clouds (obstacles that are not in the hit list)
static __device__ __noinline__ bool coarse_other_cold(const SharedViews *svp, float x, float y, float z, float r, bool on) {
    const SharedViews &sv = *svp;
    bool other = false;
    if (sv.n_hf > 0 && on) other = heightfields_hit(sv.hf, sv.n_hf, x, y, z, r);
    if (sv.n_pc > 0) other |= pointclouds_one(sv.pc, sv.n_pc, x, y, z, r, on && !other, false);
    return other;
}

// Everything rare about a pair of fine spheres, out of line: lists that overflowed (the spheres walk the grid themselves, as the
// bounding sphere did; lvl < 0: radius beyond the widest grid reach -> ordered scan), height fields, point clouds.
template <int G, int BT>
__device__ __noinline__ bool fine_pair_cold(const Scene *scp, const SharedViews *svp, const float *__restrict__ hit_list, int mine, int hmax, bool walk,
                                            int l0, int l1, float x0, float y0, float z0, float r0, float x1, float y1, float z1, float r1, bool on,
                                            bool any_mode) {
    const SharedViews &sv = *svp;
    bool hh;
    if (!walk) {
        int a0 = 0, a1 = 0;
        hits_scan2<BT>(*scp, hit_list, on ? mine : 0, hmax, x0, y0, z0, r0, x1, y1, z1, r1, any_mode, a0, a1);
        hh = (a0 | a1) < 0;
    } else {
        hh = l0 >= 0 ? grid_scan_scene(sv.grid[l0], sv.sc, x0, y0, z0, r0, on, any_mode) : ordered_scan_one<G>(&sv.sc, x0, y0, z0, r0, on, any_mode);
        const bool on1 = on && !hh;
        hh |= l1 >= 0 ? grid_scan_scene(sv.grid[l1], sv.sc, x1, y1, z1, r1, on1, any_mode) : ordered_scan_one<G>(&sv.sc, x1, y1, z1, r1, on1, any_mode);
    }
    if (sv.n_hf > 0 && on && !hh) hh = heightfields_hit(sv.hf, sv.n_hf, x0, y0, z0, r0) || heightfields_hit(sv.hf, sv.n_hf, x1, y1, z1, r1);
    if (sv.n_pc > 0) {
        hh |= pointclouds_one(sv.pc, sv.n_pc, x0, y0, z0, r0, on && !hh, any_mode);
        hh |= pointclouds_one(sv.pc, sv.n_pc, x1, y1, z1, r1, on && !hh, any_mode);
    }
    return hh;
}

// Candidate-grid variant (sorted scenes with a grid): the whole state check as ONE loop over the robot's task table (RobotTask,
// robot_common.cuh) -- the forward pass over the frames that parks the bounding-sphere centres, the links against the environment
// tip -> base (the reference's order: distal links collide first), the allowed self-collision pairs. A task first decides whether it
// needs its frame at all (a link: does the bounding sphere touch anything -- the candidate-list walk that also remembers WHAT it
// touches; a pair: do the bounding spheres touch); the streamed frame is then brought to the task's frame at the ONE place this
// function steps frames; then the fine spheres are tested (a link: two at a time against the remembered obstacles). One copy of the
// frame code and of every loop: the hot code of these kernels has to fit the 32 KB instruction cache.
// PRE: only the first part of the table -- the forward pass and the DISTAL half of the links against the environment. A state that
// fails it is invalid; one that passes needs the whole check (see launch_states: large batches of unrelated states are sifted first).
// QUIT (a first look only): the check gives up once no more than QUIT lanes are still undecided -- they count as passed. A first look
// gains nothing from states that pass (the motion kernel tests them again), and running the last few lanes of a pass to the end of
// the task table is what its passes spent most of their time on.
template <class R, int G, bool EDGE, bool KILL, int BT, bool SLICED = false, bool PRE = false, int QUIT = 0>
__device__ bool check_state_grid(const float (&q)[R::kDim], const SceneDesc &sd, const Scene &sc, float *__restrict__ slots,
                                 const SphereLevels<R> &lv, const SharedViews &sv, bool active, unsigned kill, Share<SLICED> share = {}) {
    constexpr int kTaskEnd = PRE ? R::kFrames + (R::kLinks + 1) / 2 : R::kTasks;
    const typename R::Tables &tb = R::tables();
    bool alive = active;
    trig_to_slots<R, BT>(q, slots);
    Frame F;
    int have = -1;  // frame held in F (warp-uniform)
    float *scratch = slots + hit_slot<R>() * BT;
    bool need = false;  // lanes taking part in the fine tests of the current link / pair
    int hits = 0;
#pragma unroll 1
    for (int t = 0; t < kTaskEnd; ++t) {
        const RobotTask tk = tb.tasks[t];
        const int kind = static_cast<int>(tk.kind);
        // (a sliced check: this warp's share of the links and pairs; both halves of a pair fall to the same share)
        if (kind == kTaskGroup) {
            // a group of pairs ruled out by one test: link a against the sphere that contains the bounding spheres of all the group's
            // other links (centred on one of them)
            const int la = tk.a, g = tk.b, lr = tb.group_ref[g];
            const float xs = slots[(3 * la + 0) * BT] - slots[(3 * lr + 0) * BT], ys = slots[(3 * la + 1) * BT] - slots[(3 * lr + 1) * BT],
                        zs = slots[(3 * la + 2) * BT] - slots[(3 * lr + 2) * BT];
            const float rs = tb.coarse[la].w + tb.group_r[g];
            const float d = fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs)));
            if (!__any_sync(kFull, (d < 0.0f) && alive)) t += tb.group_skip[g];
            continue;
        }
        if (SLICED && kind != kTaskFk && !share.mine(tk.a + 3u * (kind == kTaskEnv ? 0u : tk.b))) continue;
        if (kind == kTaskEnv) {
            const int l = tk.a;
            const float x = slots[(3 * l + 0) * BT], y = slots[(3 * l + 1) * BT], z = slots[(3 * l + 2) * BT];
            const float r = tb.coarse[l].w;
            const int lvl = lv.coarse[l];
            // bounding sphere: list what it touches (hits > 0 <=> coarse hit); radii beyond the widest grid reach take the ordered scan
            bool h;
            if (lvl >= 0) {
                hits = grid_collect_scene<BT>(sd.grid[lvl], sc, x, y, z, r, alive, scratch);
                h = hits > 0;
            } else {
                h = ordered_scan_one<G>(&sv.sc, x, y, z, r, alive, false);
                hits = kHitCap + 1;
            }
            // touches something that is not in the list (height field, point cloud)? Those are tested per fine sphere
            if (sd.n_hf > 0 || sd.n_pc > 0) h |= coarse_other_cold(&sv, x, y, z, r, alive && !h);
            need = group_any<G>(h && alive) && alive;
            if (!__any_sync(kFull, need)) continue;
        } else if (kind == kTaskSelfA) {
            // allowed self-collision pair (collision/validity.hh:13-44): bounding spheres on the stored centres
            const int la = tk.a, lb = tk.b;
            const float xs = slots[(3 * la + 0) * BT] - slots[(3 * lb + 0) * BT], ys = slots[(3 * la + 1) * BT] - slots[(3 * lb + 1) * BT],
                        zs = slots[(3 * la + 2) * BT] - slots[(3 * lb + 2) * BT];
            const float rs = tb.coarse[la].w + tb.coarse[lb].w;
            const float d = fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs)));
            need = group_any<G>((d < 0.0f) && alive) && alive;
            if (!__any_sync(kFull, need)) { ++t; continue; }  // (skips the pair's second half as well)
        }
        // ---- the one place frames are stepped: bring F to the task's frame by walking its chain (arm), either on from the frame held
        // when that is an earlier frame of the same chain, or from the world frame at the chain's first joint
        {
            const int k = static_cast<int>(tk.frame);
            if (have != k) {
                const int first = tb.frame_branch[k];
                int from = have + 1;
                if (have < first || have > k) { frame_identity(F); from = first; }
#pragma unroll 1
                for (int j = max(from, 1); j <= k; ++j)
                    F = R::step_inline(j, F, slots[(trig_slot<R>() + j - 1) * BT], slots[(trig_slot<R>() + R::kDim + j - 1) * BT]);
                have = k;
            }
        }
        if (kind == kTaskFk) {
#pragma unroll 1
            for (int l = static_cast<int>(tk.a); l < static_cast<int>(tk.b); ++l) {
                const float4 c = tb.coarse[l];
                float x, y, z;
                frame_apply(F, c.x, c.y, c.z, x, y, z);
                slots[(3 * l + 0) * BT] = x;
                slots[(3 * l + 1) * BT] = y;
                slots[(3 * l + 2) * BT] = z;
            }
        } else if (kind == kTaskEnv) {
            const int l = tk.a;
            const int fe = tb.link_begin[l + 1];
            // a list that overflowed (or was never made): the link's fine spheres walk the grid themselves, as the bounding sphere did
            const bool walk = __any_sync(kFull, need && hits > kHitCap);
            const bool rare = walk || sd.n_hf > 0 || sd.n_pc > 0;  // (warp-uniform)
            const int mine = need ? min(hits, kHitCap) : 0;
            const int hmax = __reduce_max_sync(kFull, mine);
            bool fhit = false;
#pragma unroll 1
            for (int f = tb.link_begin[l]; f < fe; f += 2) {
                const float4 fa = tb.fine[f], fb = tb.fine[min(f + 1, fe - 1)];  // (an odd last sphere is tested twice)
                float x0, y0, z0, x1, y1, z1;
                frame_apply(F, fa.x, fa.y, fa.z, x0, y0, z0);
                frame_apply(F, fb.x, fb.y, fb.z, x1, y1, z1);
                const bool on = need && !fhit;
                bool hh;
                if (!rare) {
                    int a0 = 0, a1 = 0;
                    hits_scan2<BT>(sc, scratch, on ? mine : 0, hmax, x0, y0, z0, fa.w, x1, y1, z1, fb.w, EDGE, a0, a1);
                    hh = (a0 | a1) < 0;
                } else {
                    hh = fine_pair_cold<G, BT>(&sv.sc, &sv, scratch, mine, hmax, walk, lv.fine[f], lv.fine[min(f + 1, fe - 1)], x0, y0, z0, fa.w, x1, y1, z1,
                                               fb.w, on, EDGE);
                }
                fhit |= hh && on;
                if (EDGE) {
                    if (__any_sync(kFull, fhit)) return !fhit;  // a fine hit of an edge: the motion is dead, stop at once
                } else {
                    if (KILL) fhit = (__ballot_sync(kFull, fhit) & kill) != 0u;
                    if (__all_sync(kFull, fhit || !need)) break;
                }
            }
            if (fhit) alive = false;
            if (!EDGE && (QUIT > 0 ? __popc(__ballot_sync(kFull, alive)) <= QUIT : !__any_sync(kFull, alive))) return alive;
        } else if (kind == kTaskSelfA) {
            // the frame of the pair's first link waits in the scratch slots while F moves on to the second link's frame
#pragma unroll
            for (int i = 0; i < 9; ++i) scratch[i * BT] = F.r[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) scratch[(9 + i) * BT] = F.t[i];
        } else {
            // The pair's fine spheres, in the SECOND link's frame: with Rel = F^-1 * A the first link's spheres (the link with fewer of
            // them; outer loop) are carried into that frame once, and the second link's spheres are then the table's frame-local
            // constants -- a rigid transform keeps distances (to rounding, ~1e-7 m; the reference measures the same distances in the
            // world frame). Each outer sphere is first tested against the second link's bounding sphere, inflated by 0.1 mm: the
            // bounding sphere contains every fine sphere of that link (to 1e-8 m, checked when the tables are generated), so a fine-fine
            // contact implies this test is negative and skipping on a non-negative value cannot change a decision.
            const int la = tk.a, lb = tk.b;
            Frame Rel;
            {
                Frame A;
#pragma unroll
                for (int i = 0; i < 9; ++i) A.r[i] = scratch[i * BT];
#pragma unroll
                for (int i = 0; i < 3; ++i) A.t[i] = scratch[(9 + i) * BT] - F.t[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) Rel.r[3 * i + j] = fmaf(F.r[6 + i], A.r[6 + j], fmaf(F.r[3 + i], A.r[3 + j], F.r[i] * A.r[j]));
                    Rel.t[i] = fmaf(F.r[6 + i], A.t[2], fmaf(F.r[3 + i], A.t[1], F.r[i] * A.t[0]));
                }
            }
            const float4 cb = tb.coarse[lb];
            const int ae = tb.link_begin[la + 1], bb = tb.link_begin[lb], be = tb.link_begin[lb + 1];
            bool fhit = false;
#pragma unroll 1
            for (int i = tb.link_begin[la]; i < ae; ++i) {
                const float4 al = tb.fine[i];
                float qx, qy, qz;
                frame_apply(Rel, al.x, al.y, al.z, qx, qy, qz);
                const float xs0 = qx - cb.x, ys0 = qy - cb.y, zs0 = qz - cb.z;
                const float rs0 = al.w + cb.w + 1.0e-4f;
                const bool cand = (fmaf(-rs0, rs0, fmaf(zs0, zs0, fmaf(ys0, ys0, xs0 * xs0))) < 0.0f) && need && !fhit;
                if (!__any_sync(kFull, cand)) continue;
                int acc = 0;
#pragma unroll 2
                for (int j = bb; j < be; ++j) {
                    const float4 fl = tb.fine[j];
                    const float xs2 = qx - fl.x, ys2 = qy - fl.y, zs2 = qz - fl.z;
                    const float rs2 = al.w + fl.w;
                    acc |= __float_as_int(fmaf(-rs2, rs2, fmaf(zs2, zs2, fmaf(ys2, ys2, xs2 * xs2))));
                }
                fhit |= (acc < 0) && cand;
                if (EDGE) {
                    if (__any_sync(kFull, fhit)) return !fhit;
                } else {
                    if (KILL) fhit = (__ballot_sync(kFull, fhit) & kill) != 0u;
                    if (__all_sync(kFull, fhit || !need)) break;
                }
            }
            if (fhit) alive = false;
            if (!EDGE && (QUIT > 0 ? __popc(__ballot_sync(kFull, alive)) <= QUIT : !__any_sync(kFull, alive))) return alive;
        }
    }
    return alive;
}

// ------------------------------------------------------------------------------------------------ attachments
// Robot::fkcc_attach (robots/panda.hh:10264-15469): after the robot's own tests, the attachment -- spheres rigidly fixed to the end
// effector, collision/attachments.hh:13-56 -- is posed at eefk(q) * tf and tested against the environment
// (attachment_environment_collision, validity.hh:251-268) and against the links the robot's table lists
// (attachment_sphere_collision, validity.hh:270-293: bounding sphere first, then the link's fine spheres). Rare path: out of
// line, per-lane results, warp-uniform loops; the caller applies its exit policy to the returned flag.
template <int G, bool MASKED, bool GRID>
__device__ __forceinline__ bool env_sphere_hit(const SharedViews &sv, float x, float y, float z, float r, bool on) {
    bool h;
    int lvl = -1;
    if (GRID) {
        for (int l = kGridLevels - 1; l >= 0; --l)
            if (l < sv.n_grid && r <= sv.grid[l].reach) lvl = l;
    }
    if (GRID && lvl >= 0) {
        h = grid_scan_scene(sv.grid[lvl], sv.sc, x, y, z, r, on, false);
    } else {
        Tile<1> s1;
        s1.x[0] = x; s1.y[0] = y; s1.z[0] = z; s1.r[0] = r;
        const float e = on ? sqrtf(fmaf(z, z, fmaf(y, y, x * x))) + r : -INFINITY;
        s1.ext[0] = group_max<G>(e);
        if (!on) { s1.x[0] = kIdle; s1.y[0] = kIdle; s1.z[0] = kIdle; }
        h = scan_scene<1, kUntilAllHit, MASKED>(sv.sc, s1, on);
    }
    if (sv.n_hf > 0 && on && !h) h = heightfields_hit(sv.hf, sv.n_hf, x, y, z, r);
    if (sv.n_pc > 0) h |= capt_any(sv, x, y, z, r, on && !h);
    return h && on;
}

template <class R, int G, bool MASKED, bool GRID, int BT>
__device__ __noinline__ bool attachment_hit(const float *__restrict__ slots, const SharedViews *__restrict__ svp, bool alive) {
    const SharedViews &sv = *svp;
    const typename R::Tables &tb = R::tables();
    // N = eefk(q) * tf: the end-effector frame is a constant transform away from one of the moving frames
    Frame F, N;
    int have = -1;
    frame_seek<R, BT>(F, have, R::kEeFrame, slots);
    {
        Frame E;
        R::ee_pose(F, E);
        float t[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) t[i] = __ldg(sv.att + i);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) N.r[3 * r + c] = fmaf(E.r[3 * r + 2], t[8 + c], fmaf(E.r[3 * r + 1], t[4 + c], E.r[3 * r] * t[c]));
            N.t[r] = fmaf(E.r[3 * r + 2], t[11], fmaf(E.r[3 * r + 1], t[7], fmaf(E.r[3 * r], t[3], E.t[r])));
        }
    }
    const float4 *sph = reinterpret_cast<const float4 *>(sv.att + 16);
    const int n = sv.n_att;
    bool hit = false;
    // attachment vs. environment
    for (int i = 0; i < n; ++i) {
        if (!__any_sync(kFull, alive && !hit)) break;
        const float4 a = __ldg(sph + i);
        float x, y, z;
        frame_apply(N, a.x, a.y, a.z, x, y, z);
        hit |= env_sphere_hit<G, MASKED, GRID>(sv, x, y, z, a.w, alive && !hit);
    }
    // attachment vs. the robot's links
    have = -1;
    for (int k = 0; k < R::kAttLinks; ++k) {
        if (!__any_sync(kFull, alive && !hit)) break;
        const int l = tb.att_links[k];
        const float bx = slots[(3 * l + 0) * BT], by = slots[(3 * l + 1) * BT], bz = slots[(3 * l + 2) * BT];
        const float rb = tb.coarse[l].w;
        int acc = 0;
        for (int i = 0; i < n; ++i) {
            const float4 a = __ldg(sph + i);
            float x, y, z;
            frame_apply(N, a.x, a.y, a.z, x, y, z);
            const float xs = bx - x, ys = by - y, zs = bz - z, rs = rb + a.w;
            acc |= __float_as_int(fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs))));
        }
        const bool need = group_any<G>(acc < 0 && alive && !hit) && alive && !hit;
        if (!__any_sync(kFull, need)) continue;
        frame_seek<R, BT>(F, have, tb.link_frame[l], slots);
        int facc = 0;
        for (int f = tb.link_begin[l]; f < tb.link_begin[l + 1]; ++f) {
            const float4 fl = tb.fine[f];
            float px, py, pz;
            frame_apply(F, fl.x, fl.y, fl.z, px, py, pz);
            for (int i = 0; i < n; ++i) {
                const float4 a = __ldg(sph + i);
                float x, y, z;
                frame_apply(N, a.x, a.y, a.z, x, y, z);
                const float xs = px - x, ys = py - y, zs = pz - z, rs = fl.w + a.w;
                facc |= __float_as_int(fmaf(-rs, rs, fmaf(zs, zs, fmaf(ys, ys, xs * xs))));
            }
        }
        hit |= (facc < 0) && need;
    }
    return hit;
}

template <class R, int G, bool EDGE, bool MASKED, bool GRID, int BT, bool KILL = false, bool SLICED = false, bool PRE = false, int QUIT = 0>
__device__ __forceinline__ bool check_state(const float (&q)[R::kDim], const SceneDesc &sd, const Scene &sc, float *__restrict__ slots,
                                            const SphereLevels<R> &lv, const SharedViews &sv, bool active, unsigned kill = 0u,
                                            Share<SLICED> share = {}) {
    static_assert(!PRE || GRID, "the sifting pass exists for the candidate-grid variant only");
    static_assert(QUIT == 0 || (GRID && KILL && !EDGE), "only a first look on a grid scene may give up early");
    bool good = GRID ? check_state_grid<R, G, EDGE, KILL, BT, SLICED, PRE, QUIT>(q, sd, sc, slots, lv, sv, active, kill, share)
                     : check_state_ordered<R, G, EDGE, KILL, MASKED, BT, SLICED>(q, sd, sc, slots, sv.pc, active, kill, share);
    if (!PRE && sd.n_att > 0 && share.mine(0u)) {
        // (an EDGE pass that already has a hit is over: its caller drops the motion)
        if (EDGE && __any_sync(kFull, active && !good)) return good;
        const bool alive = active && good;
        if (__any_sync(kFull, alive)) {
            bool ahit = attachment_hit<R, G, MASKED, GRID, BT>(slots, &sv, alive);
            if (KILL) ahit = (__ballot_sync(kFull, ahit) & kill) != 0u;
            good = good && !ahit;
        }
    }
    return good;
}

// ------------------------------------------------------------------------------------------------ kernels
struct LaunchCounters {
    unsigned long long next;  // dynamic work queue head
};

// Work description of a bucketed batch of motions (written on the device by the pre-pass below, read by the two kernels that follow
// it): the motions are grouped by their number of rakes n = max(ceil(|q1 - q0| / 8 * resolution), 1).
//   bucket b = 1 .. 4: motions with exactly b rakes -> validate_rakes_kernel packs their rakes four to a warp pass
//   bucket 5         : longer motions               -> validate_motions_kernel, one warp per motion, passes in sequence with early exit
constexpr int kRakeBuckets = 6;  // index 0 unused
// states of rake 0 a first look tests per motion: 8 = the whole rake, four motions per warp pass; 4 = every other state, eight motions;
// 2 = the middle and the end state, sixteen motions. Measured on C2 (2^20 random motions, 98.8 % invalid): 2.83 / 2.46 / 2.30 ms.
#ifndef B200MP_FIRST_LANES
#define B200MP_FIRST_LANES 2
#endif
constexpr int kFirstLanes = B200MP_FIRST_LANES;
// undecided lanes at which a first-look pass on a grid scene gives up (check_state_grid's QUIT)
#ifndef B200MP_LOOK_QUIT
#define B200MP_LOOK_QUIT 4
#endif
constexpr int kLookQuit = B200MP_LOOK_QUIT;
struct BucketPlan {
    unsigned cnt[kRakeBuckets];       // motions per bucket
    unsigned off[kRakeBuckets];       // start of each bucket in the list
    unsigned cursor[kRakeBuckets];    // fill positions of the scatter pass
    unsigned pass_end[kRakeBuckets];  // warp passes up to and including bucket b (b <= 4)
    unsigned long long next_pass;     // work queue head of validate_rakes_kernel (short buckets)
    unsigned long long next_first;    // work queue head of its first-look instantiation
    LaunchCounters long_queue;        // work queue head of validate_motions_kernel over bucket 5
    // first look at the long motions (bucket 5): rake 0 of four motions per warp pass, see validate_rakes_kernel
    unsigned pre_passes;              // ceil(cnt[5] / motions per first-look pass)
    unsigned pre_sample;              // passes that always run; their outcome decides about the rest
    unsigned pre_done, pre_killed, pre_tested;
    unsigned pre_decision;            // 0 undecided, 1 run the remaining first-look passes, 2 skip them
};

// Threads per block / blocks per SM of the validation kernels (measured on C2: the motion kernel gains ~4 % from 20 warps
// at 96 registers, the state kernel prefers 16 warps at 128). Shared memory per block = scene blob + lane_slots floats per thread.
constexpr int kMotionBlocksPerSM = 1;
constexpr int kStateThreads = 256, kStateBlocksPerSM = 2;
// (measured on the 2^20 roadmap, Panda: 640 threads at 96 registers 100 ms; 704 at 80 registers 126 ms -- spills --; 576 at 96: 107 ms)
#ifndef B200MP_MOTION_THREADS
#define B200MP_MOTION_THREADS 640
#endif
// the motion kernel's block: as many warps as keep the per-thread slots within ~160 KB (Baxter: 33 links + 14 joints)
template <class R>
__host__ __device__ constexpr int motion_threads() {
    return lane_slots<R>() * B200MP_MOTION_THREADS * 4 <= 172 * 1024 ? B200MP_MOTION_THREADS
           : (lane_slots<R>() * 512 * 4 <= 160 * 1024 ? 512 : (lane_slots<R>() * 384 * 4 <= 160 * 1024 ? 384 : 256));
}

// the state kernel's block: one block of 20 warps where the robot's slots allow it, else two blocks of 8 warps
template <class R>
__host__ __device__ constexpr int state_threads() { return motion_threads<R>() == B200MP_MOTION_THREADS ? B200MP_MOTION_THREADS : kStateThreads; }
template <class R>
__host__ __device__ constexpr int state_blocks_per_sm() { return state_threads<R>() == kStateThreads ? kStateBlocksPerSM : 1; }

template <class R>
__host__ __device__ constexpr size_t lane_store_bytes(int threads) { return static_cast<size_t>(lane_slots<R>()) * threads * sizeof(float); }

// horizontal sum in the reference's order (external/vamp/src/impl/vamp/vector/avx.hh:442-452): lanes padded to 8
template <int DIM>
__device__ __forceinline__ float vamp_l2_norm(const float (&v)[DIM]) {
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < DIM; ++i) s[i & 7] = fmaf(v[i], v[i], s[i & 7]);  // rows are summed first (DIM <= 8: one row)
    const float s0 = s[4] + s[0], s1 = s[5] + s[1], s2 = s[6] + s[2], s3 = s[7] + s[3];
    return sqrtf((s0 + s2) + (s1 + s3));
}

// One warp per motion; motions are taken from a global queue (edge cost varies by 100x with early exit).
// `list` / `plan` (both or neither): the motions to check are list[plan->off[5] .. + plan->cnt[5]) instead of 0 .. m - 1.
template <class R, bool MASKED, bool GRID>
__global__ void __launch_bounds__(motion_threads<R>(), kMotionBlocksPerSM)
validate_motions_kernel(SceneDesc sd, const float *__restrict__ q0, const float *__restrict__ q1, const int2 *__restrict__ pairs,
                        long long m, int resolution, uint8_t *__restrict__ valid, LaunchCounters *cnt, const unsigned *__restrict__ list,
                        const BucketPlan *__restrict__ plan) {
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long bar;
    __shared__ SphereLevels<R> lv;
    fill_levels<R>(sd, lv);
    stage_scene(smem, sd.blob, sd.blob_floats, &bar);
    const Scene sc = scene_view(sd, smem);
    float *slots = smem + sd.blob_floats + threadIdx.x;
    __shared__ SharedViews sv;
    fill_views(sd, sc, sv);
    __syncthreads();
    constexpr int D = R::kDim;
    const int lane = threadIdx.x & 31;
    const int rake_in_pass = lane >> 3;
    const float pct = static_cast<float>((lane & 7) + 1) / 8.0f;
    if (plan != nullptr) {
        m = plan->cnt[kRakeBuckets - 1];
        list += plan->off[kRakeBuckets - 1];
        if (m == 0) return;
    }

    // Queue entries are taken `take` at a time. After a first look that went on (pre_decision == 1) most entries of the list may already
    // be invalid: a warp then takes several, each lane looks at one (list entry -> valid flag: two dependent loads) and the warp works
    // through the ones still open -- one queue round trip and one load round trip per take instead of per entry (at 3 % open entries the
    // skip loop was latency-bound: 2^20 entries x two round trips over 2960 warps). The take is sized from the first look's sample for
    // about one open entry each, and so that a warp still gets eight takes or more (a larger one would only unbalance the tail);
    // without a first look it is one entry.
    unsigned take = 1u;
    if (plan != nullptr && plan->pre_decision == 1u) {
        const unsigned tested = plan->pre_tested, killed = min(plan->pre_killed, tested);
        const unsigned per_open = tested / max(tested - killed, 1u);
        const unsigned long long fair = static_cast<unsigned long long>(m) / (8ull * gridDim.x * (blockDim.x >> 5));
        take = static_cast<unsigned>(min(static_cast<unsigned long long>(min(max(per_open, 1u), 32u)), max(fair, 1ull)));
    }
    for (;;) {
        unsigned long long head = 0;
        if (lane == 0) head = atomicAdd(&cnt->next, static_cast<unsigned long long>(take));
        head = __shfl_sync(kFull, head, 0);
        if (head >= static_cast<unsigned long long>(m)) break;
        unsigned long long e_mine = head + static_cast<unsigned>(lane);
        bool open = static_cast<unsigned>(lane) < take && e_mine < static_cast<unsigned long long>(m);
        if (open && list != nullptr) {
            e_mine = __ldg(list + e_mine);
            open = valid[e_mine] != 0;  // else the first look (validate_rakes_kernel) already found it in collision
        }
        unsigned pending = __ballot_sync(kFull, open);
        while (pending != 0u) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1u;
            const unsigned long long e = __shfl_sync(kFull, e_mine, src);

            float a[D], v[D], q[D], back[D];
            // roadmap edges address their end points in one state table (pairs != nullptr: q0 == q1 == the table)
            unsigned long long ia = e, ib = e;
            if (pairs != nullptr) {
                const int2 pr = __ldg(pairs + e);
                ia = static_cast<unsigned long long>(pr.x);
                ib = static_cast<unsigned long long>(pr.y);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {
                a[j] = __ldg(q0 + ia * D + j);
                v[j] = __ldg(q1 + ib * D + j) - a[j];
            }
            const float dist = vamp_l2_norm<D>(v);
            // n = max(ceil(dist / rake * resolution), 1)  (planning/validate.hh:41)
            const float nf = fmaxf(ceilf(dist / 8.0f * static_cast<float>(resolution)), 1.0f);
            const int n = static_cast<int>(nf);
            const float denom = static_cast<float>(8 * n);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                q[j] = fmaf(v[j], pct, a[j]);  // block = start + vector * percents
                back[j] = v[j] / denom;        // backstep = vector / (rake * n)
            }
            // this lane's first rake: `rake_in_pass` sequential back-steps, exactly as the reference accumulates them
            for (int r = 0; r < rake_in_pass; ++r) {
#pragma unroll
                for (int j = 0; j < D; ++j) q[j] = q[j] - back[j];
            }
            bool ok = true;
            for (int base = 0; base < n; base += 4) {
                const bool active = (base + rake_in_pass) < n;
                const bool good = check_state<R, 8, true, MASKED, GRID, motion_threads<R>()>(q, sd, sc, slots, lv, sv, active);
                if (__any_sync(kFull, active && !good)) { ok = false; break; }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int j = 0; j < D; ++j) q[j] = q[j] - back[j];
                }
            }
            if (lane == 0) valid[e] = ok ? 1 : 0;
        }
    }
}

// ---- short motions: rakes of different motions share a warp pass ----------------------------------------------------------------
// A roadmap edge is 2-3 rakes long (its length is a k-th neighbour distance), so the warp-per-motion kernel above runs with half to
// three quarters of its lanes. Here a warp pass takes four RAKES from the bucket's task list -- task t of bucket b is rake t % b of
// motion t / b -- whatever motions they belong to. States, arithmetic and decisions are those of the kernel above (start + v * pct,
// then `rake` sequential back-steps). A fine hit ends the work of every lane of the same motion in the warp (KILL mask); rakes of
// that motion running in other warps finish on their own, later passes skip it (valid[e] is already 0).
// valid[] must be 1 on entry (the pre-pass sets it); a pass only ever writes 0.
// FIRST: the instantiation that runs the first-look passes over the long motions (below); the other one runs the short buckets.
// (Two kernels rather than one loop with both kinds of pass: the short-bucket loop is the roadmap's hot loop and stays lean.)
template <class R, bool MASKED, bool GRID, bool FIRST>
__global__ void __launch_bounds__(motion_threads<R>(), kMotionBlocksPerSM)
validate_rakes_kernel(SceneDesc sd, const float *__restrict__ q0, const float *__restrict__ q1, const int2 *__restrict__ pairs,
                      uint8_t *__restrict__ valid, const unsigned *__restrict__ list, BucketPlan *__restrict__ plan) {
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long bar;
    __shared__ SphereLevels<R> lv;
    const unsigned short_passes = FIRST ? 0u : plan->pass_end[kRakeBuckets - 2];
    const unsigned total = FIRST ? plan->pre_passes : short_passes;
    if (total == 0) return;
    fill_levels<R>(sd, lv);
    stage_scene(smem, sd.blob, sd.blob_floats, &bar);
    const Scene sc = scene_view(sd, smem);
    float *slots = smem + sd.blob_floats + threadIdx.x;
    __shared__ SharedViews sv;
    fill_views(sd, sc, sv);
    // the plan's constants (final before this kernel starts) in shared memory: a pass looks three of them up before it can load anything
    __shared__ unsigned s_pass_end[kRakeBuckets], s_cnt[kRakeBuckets], s_off[kRakeBuckets], s_pre_sample;
    if (threadIdx.x < kRakeBuckets) {
        s_pass_end[threadIdx.x] = plan->pass_end[threadIdx.x];
        s_cnt[threadIdx.x] = plan->cnt[threadIdx.x];
        s_off[threadIdx.x] = plan->off[threadIdx.x];
    }
    if (threadIdx.x == 0) s_pre_sample = plan->pre_sample;
    __syncthreads();
    constexpr int D = R::kDim;
    const int lane = threadIdx.x & 31;
    // lanes per motion / rake: the short buckets run whole rakes (8 states); a first look takes kFirstLanes of rake 0's states per motion
    constexpr int LPM = FIRST ? kFirstLanes : 8;
    constexpr unsigned kGroupMask = (1u << LPM) - 1u;
    const int group = lane / LPM;
    // (a first look at roadmap edges -- both ends are milestones, known to be valid -- with two lanes: 3/8 and 6/8 instead of 4/8 and 8/8)
    const float pct = FIRST ? (LPM == 2 && pairs != nullptr ? static_cast<float>(3 * ((lane & 1) + 1)) / 8.0f
                                                            : static_cast<float>((8 / LPM) * ((lane & (LPM - 1)) + 1)) / 8.0f)
                            : static_cast<float>((lane & 7) + 1) / 8.0f;
    // the short buckets take their passes four at a time from the queue (one atomic round trip per four passes; passes still run in
    // increasing order inside a warp); the first look, whose sample / verdict logic counts single passes, one at a time
    constexpr unsigned kTake = FIRST ? 1u : 4u;
    unsigned long long next = 0;
    unsigned left = 0;
    for (;;) {
        if (left == 0u) {
            if (lane == 0) next = atomicAdd(FIRST ? &plan->next_first : &plan->next_pass, static_cast<unsigned long long>(kTake));
            next = __shfl_sync(kFull, next, 0);
            left = kTake;
        }
        const unsigned long long p64 = next++;
        --left;
        if (p64 >= total) break;
        const unsigned p = static_cast<unsigned>(p64);
        // First look at the LONG motions (the passes after the short buckets): rake 0 -- eight states spread over the whole motion -- of
        // four motions per pass. Where most motions are in collision (random end points) this settles them at a quarter of a warp pass
        // each; validate_motions_kernel then skips what is already invalid. Where most survive it would only add work, so the kill rate of
        // the first pre_sample passes decides whether the first look goes on (decided on the device, no host round trip).
        constexpr bool first_look = FIRST;
        int b = 1;
        unsigned t;
        if (first_look) {
            const unsigned lp = p - short_passes;
            if (lp >= s_pre_sample) {
                // beyond the sample: run unless the verdict is already "skip". A warp does not idle while the sample is being
                // finished by others -- it takes one more pass on trust -- but no more than that: past one pass per warp of the grid
                // it waits for the verdict (where passes are slow, e.g. ordered scans of a large scene, optimism would cost too much)
                unsigned dec = 0;
                if (lane == 0) {
                    const volatile unsigned *flag = &plan->pre_decision;
                    dec = *flag;
                    if (lp >= s_pre_sample + gridDim.x * (blockDim.x >> 5))
                        while ((dec = *flag) == 0u) __nanosleep(200);
                }
                dec = __shfl_sync(kFull, dec, 0);
                if (dec == 2u) continue;
            }
            b = kRakeBuckets - 1;
            t = static_cast<unsigned>(32 / LPM) * lp + static_cast<unsigned>(group);
        } else {
            while (p >= s_pass_end[b]) ++b;  // (warp-uniform; at most three steps)
            t = 4u * (p - s_pass_end[b - 1]) + static_cast<unsigned>(group);
        }
        const unsigned ei = first_look ? t : t / static_cast<unsigned>(b);
        const int rake = first_look ? 0 : static_cast<int>(t - ei * static_cast<unsigned>(b));
        bool active = ei < s_cnt[b];
        const unsigned e = active ? __ldg(list + s_off[b] + ei) : 0xffffffffu;
        if (active && valid[e] == 0) active = false;  // another pass already found this motion in collision
        float a[D], v[D], q[D], back[D];
        unsigned long long ia = e, ib = e;
        if (active && pairs != nullptr) {
            const int2 pr = __ldg(pairs + e);
            ia = static_cast<unsigned long long>(pr.x);
            ib = static_cast<unsigned long long>(pr.y);
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            a[j] = active ? __ldg(q0 + ia * D + j) : 0.0f;
            v[j] = active ? __ldg(q1 + ib * D + j) - a[j] : 0.0f;
        }
        // n = max(ceil(dist / rake * resolution), 1)  (planning/validate.hh:41) = b: the pre-pass computed it with the motion kernel's expression
        const float denom = static_cast<float>(8 * b);  // (not used by a first look: rake 0 takes no back-step)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            q[j] = fmaf(v[j], pct, a[j]);  // block = start + vector * percents
            back[j] = v[j] / denom;        // backstep = vector / (rake * n)
        }
        for (int r = 0; r < rake; ++r) {   // sequential back-steps, exactly as the reference accumulates them
#pragma unroll
            for (int j = 0; j < D; ++j) q[j] = q[j] - back[j];
        }
        // lanes of the same motion in this warp
        unsigned kill = 0u;
#pragma unroll
        for (int g2 = 0; g2 < 32 / LPM; ++g2) {
            const unsigned eg = __shfl_sync(kFull, e, LPM * g2);
            const bool ag = __shfl_sync(kFull, active ? 1 : 0, LPM * g2) != 0;
            if (ag && eg == e) kill |= kGroupMask << (LPM * g2);
        }
        bool bad = false;
        if (__any_sync(kFull, active)) {
            // (G = LPM: a first look's group of lanes is its motion; for scenes given unsorted the scan of a sphere then reaches no
            // further than the reference's own -- the group maximum is over a subset of the rake --, so it can only find what the
            // reference finds as well)
            const bool good = check_state<R, LPM, false, MASKED, GRID, motion_threads<R>(), true, false, false, (FIRST && GRID) ? kLookQuit : 0>(
                q, sd, sc, slots, lv, sv, active, kill);
            bad = group_any<LPM>(active && !good);
            if (bad && (lane & (LPM - 1)) == 0) valid[e] = 0;
        }
        if (first_look && p - short_passes < s_pre_sample) {
            const unsigned tested = __popc(__ballot_sync(kFull, active && (lane & (LPM - 1)) == 0));
            const unsigned killed = __popc(__ballot_sync(kFull, bad && (lane & (LPM - 1)) == 0));
            if (lane == 0) {
                atomicAdd(&plan->pre_tested, tested);
                atomicAdd(&plan->pre_killed, killed);
                __threadfence();
                if (atomicAdd(&plan->pre_done, 1u) + 1u == s_pre_sample) {
                    // the sample is complete: go on with the first look only if it settles most motions
                    const unsigned k = atomicAdd(&plan->pre_killed, 0u), n = atomicAdd(&plan->pre_tested, 0u);
                    __threadfence();
                    // (a look costs LPM / 32 of a pass per motion and saves a killed motion at least half a pass of the other kernel:
                    // with whole rakes the break-even kill rate was tuned at 1 / 2; it scales with the lanes a look spends)
                    atomicExch(&plan->pre_decision, 16u * k > static_cast<unsigned>(LPM) * n ? 1u : 2u);
                }
            }
        }
    }
}

// Pre-pass of a bucketed batch: number of rakes of every motion -> bucket key; per-bucket counts (whole batch and per block);
// valid[] = 1. The list is then filled in a STABLE order: consecutive motions of a bucket stay consecutive, so the rakes that share
// a warp pass belong to neighbouring motions of the batch (roadmap edges of one milestone: nearby states, the same links near the
// same obstacles), which is what keeps the lock-step lanes of a pass doing the same work.
constexpr int kBucketThreads = 1024;  // motions per block of the pre-pass (one row of per-block counts each: fewer rows for the plan kernel)
template <class R>
__global__ void __launch_bounds__(kBucketThreads)
bucket_count_kernel(const float *__restrict__ q0, const float *__restrict__ q1, const int2 *__restrict__ pairs, long long m, int resolution,
                    unsigned char *__restrict__ keys, uint8_t *__restrict__ valid, unsigned *__restrict__ blk_cnt, BucketPlan *__restrict__ plan) {
    constexpr int D = R::kDim;
    __shared__ unsigned s_cnt[8];
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    int key = 0;
    if (e < m) {
        long long ia = e, ib = e;
        if (pairs != nullptr) {
            const int2 pr = __ldg(pairs + e);
            ia = pr.x;
            ib = pr.y;
        }
        float v[D];
#pragma unroll
        for (int j = 0; j < D; ++j) v[j] = __ldg(q1 + ib * D + j) - __ldg(q0 + ia * D + j);
        const float dist = vamp_l2_norm<D>(v);
        const float nf = fmaxf(ceilf(dist / 8.0f * static_cast<float>(resolution)), 1.0f);
        key = nf <= 4.0f ? static_cast<int>(nf) : kRakeBuckets - 1;  // (NaN lengths go to the long bucket as well)
        keys[e] = static_cast<unsigned char>(key);
        valid[e] = 1;
    }
#pragma unroll
    for (int k = 1; k < kRakeBuckets; ++k) {  // (one shared-memory atomic per warp and key)
        const unsigned mask = __ballot_sync(kFull, key == k);
        if ((threadIdx.x & 31) == 0 && mask != 0u) atomicAdd(&s_cnt[k], static_cast<unsigned>(__popc(mask)));
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        blk_cnt[static_cast<size_t>(blockIdx.x) * 8 + threadIdx.x] = s_cnt[threadIdx.x];
        if (threadIdx.x < kRakeBuckets && s_cnt[threadIdx.x]) atomicAdd(&plan->cnt[threadIdx.x], s_cnt[threadIdx.x]);
    }
}

// one block: bucket offsets and pass counts; blk_cnt[block][key] -> exclusive prefix over the blocks (in place). 1024 rows per round,
// one row per thread (coalesced 32-byte rows), block-wide scan per key, carry from round to round.
static __global__ void __launch_bounds__(1024) bucket_plan_kernel(BucketPlan *__restrict__ plan, unsigned *__restrict__ blk_cnt, int nb, int first_look) {
    __shared__ unsigned s_warp[32][8], s_tot[8], s_carry[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t < 8) s_carry[t] = 0;
    __syncthreads();
    uint4 *rows = reinterpret_cast<uint4 *>(blk_cnt);
    for (int base = 0; base < nb; base += 1024) {
        const int r = base + t;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (r < nb) {
            lo = rows[2 * static_cast<size_t>(r)];
            hi = rows[2 * static_cast<size_t>(r) + 1];
        }
        const unsigned v[kRakeBuckets] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y};
        unsigned incl[kRakeBuckets];
#pragma unroll
        for (int k = 0; k < kRakeBuckets; ++k) {
            unsigned x = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(kFull, x, o);
                if (lane >= o) x += up;
            }
            incl[k] = x;
            if (lane == 31) s_warp[w][k] = x;
        }
        __syncthreads();
        if (w < kRakeBuckets) {  // warp k: exclusive scan of the 32 warp totals of key k
            const unsigned mine = s_warp[lane][w];
            unsigned x = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(kFull, x, o);
                if (lane >= o) x += up;
            }
            s_warp[lane][w] = x - mine;
            if (lane == 31) s_tot[w] = x;
        }
        __syncthreads();
        if (r < nb) {
            unsigned ex[kRakeBuckets];
#pragma unroll
            for (int k = 0; k < kRakeBuckets; ++k) ex[k] = s_carry[k] + s_warp[w][k] + incl[k] - v[k];
            rows[2 * static_cast<size_t>(r)] = make_uint4(ex[0], ex[1], ex[2], ex[3]);
            rows[2 * static_cast<size_t>(r) + 1] = make_uint4(ex[4], ex[5], 0u, 0u);
        }
        __syncthreads();
        if (t < kRakeBuckets) s_carry[t] += s_tot[t];
        __syncthreads();
    }
    if (t == 0) {
        unsigned off = 0, passes = 0;
        plan->off[0] = 0; plan->cursor[0] = 0; plan->pass_end[0] = 0;
        for (int b = 1; b < kRakeBuckets; ++b) {
            plan->off[b] = off;
            plan->cursor[b] = 0;
            off += plan->cnt[b];
            if (b < kRakeBuckets - 1) passes += (plan->cnt[b] * static_cast<unsigned>(b) + 3u) / 4u;
            plan->pass_end[b] = passes;
        }
        plan->next_pass = 0;
        plan->next_first = 0;
        plan->long_queue.next = 0;
        const unsigned per_pass = 32u / static_cast<unsigned>(kFirstLanes);  // motions per first-look pass
        const unsigned pre = first_look ? (plan->cnt[kRakeBuckets - 1] + per_pass - 1u) / per_pass : 0u;
        plan->pre_passes = pre;
        plan->pre_sample = pre <= 512u ? pre : min(max(512u, pre / 32u), 4096u);
        plan->pre_done = 0;
        plan->pre_killed = 0;
        plan->pre_tested = 0;
        plan->pre_decision = pre == 0u ? 2u : 0u;
    }
}

static __global__ void __launch_bounds__(kBucketThreads) bucket_scatter_kernel(const unsigned char *__restrict__ keys, long long m,
                                                                               const unsigned *__restrict__ blk_base, unsigned *__restrict__ list,
                                                                               const BucketPlan *__restrict__ plan) {
    __shared__ unsigned s_w[kBucketThreads / 32][8];  // [warp][key]: motions of that key in the warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int key = e < m ? keys[e] : 0;
    unsigned rank = 0;
#pragma unroll
    for (int k = 1; k < kRakeBuckets; ++k) {
        const unsigned mask = __ballot_sync(kFull, key == k);
        if (key == k) rank = __popc(mask & ((1u << lane) - 1u));
        if (lane == 0) s_w[w][k] = __popc(mask);
    }
    __syncthreads();
    if (key != 0) {
        unsigned base = blk_base[static_cast<size_t>(blockIdx.x) * 8 + key];
        for (int w2 = 0; w2 < w; ++w2) base += s_w[w2][key];
        list[plan->off[key] + base + rank] = static_cast<unsigned>(e);
    }
}

// One lane per state: VampStateValidityChecker::isValid = validate_motion<Robot, rake, 1>(q, q), i.e. one fkcc on
// 8 identical lanes (src/ompl/vamp/VampStateValidityChecker.h:37-41) -> group size 1 here.
// PRE: the sifting pass (check_state_grid<..., PRE>). m_dev (optional): the number of states to visit, read on the device (the survivors
// of a sifting pass are counted there).
template <class R, bool MASKED, bool GRID, bool PRE = false>
__global__ void __launch_bounds__(state_threads<R>(), state_blocks_per_sm<R>())
validate_states_kernel(SceneDesc sd, const float *__restrict__ qs, long long m, uint8_t *__restrict__ valid, LaunchCounters *cnt,
                       const unsigned *__restrict__ order, const unsigned *__restrict__ m_dev) {
    if (m_dev != nullptr) m = static_cast<long long>(*m_dev);
    if (m == 0) return;
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long bar;
    __shared__ SphereLevels<R> lv;
    fill_levels<R>(sd, lv);
    stage_scene(smem, sd.blob, sd.blob_floats, &bar);
    const Scene sc = scene_view(sd, smem);
    float *slots = smem + sd.blob_floats + threadIdx.x;
    __shared__ SharedViews sv;
    fill_views(sd, sc, sv);
    __syncthreads();
    constexpr int D = R::kDim;
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(&cnt->next, 1ull);
        w = __shfl_sync(kFull, w, 0);
        const long long first = static_cast<long long>(w) * 32;
        if (first >= m) break;
        const bool active = first + lane < m;
        // `order` (optional): the states are visited in this order (a warp's 32 states then come from one bucket of similar
        // configurations, see prm_sample_async) -- results still go to the states' own slots
        const long long i = active ? (order != nullptr ? static_cast<long long>(__ldg(order + first + lane)) : first + lane) : 0;
        float q[D];
#pragma unroll
        for (int j = 0; j < D; ++j) q[j] = active ? __ldg(qs + i * D + j) : 0.0f;
        // validate_vector with start == goal: block = q + 0 * percents
        const bool good = check_state<R, 1, false, MASKED, GRID, state_threads<R>(), false, false, PRE>(q, sd, sc, slots, lv, sv, active);
        if (active) valid[i] = good ? 1 : 0;
    }
}

// survivors of a sifting pass, in the order they were visited (block by block: neighbours stay neighbours, so the warps of the full
// pass still hold similar configurations); out[0 .. *count)
static __global__ void __launch_bounds__(256) sift_survivors_kernel(const uint8_t *__restrict__ flags, const unsigned *__restrict__ order, long long m,
                                                                    unsigned *__restrict__ out, unsigned *__restrict__ count) {
    __shared__ unsigned s_warp[8], s_base;
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned id = 0;
    bool keep = false;
    if (p < m) {
        id = order != nullptr ? __ldg(order + p) : static_cast<unsigned>(p);
        keep = flags[id] != 0;
    }
    const unsigned b = __ballot_sync(kFull, keep);
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < 8; ++i) { const unsigned c = s_warp[i]; s_warp[i] = tot; tot += c; }
        s_base = tot ? atomicAdd(count, tot) : 0u;
    }
    __syncthreads();
    if (keep) out[s_base + s_warp[w] + __popc(b & ((1u << lane) - 1u))] = id;
}

}  // namespace b200mp
