// b200mp C ABI implementation: contexts, environments, validation launches (see include/b200mp.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200mp.h"
#include "validate_launch.cuh"
#include "capt_order.cuh"
#include "capt_gpu_build.cuh"
#include "robots/panda_gen.cuh"
#include "robots/fetch_gen.cuh"
#include "robots/ur5_gen.cuh"
#include "robots/baxter_gen.cuh"
#include "host_common.hpp"

using namespace b200mp;

namespace b200mp {
thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// resident blocks per SM of a kernel at a given dynamic shared-memory size; the opt-in attribute and the occupancy query
// are made once per (device, kernel, size) and remembered (they cost more than a small launch). Both are per-device state.
int blocks_per_sm(int device, const void *kern, int threads, size_t smem, int *per_sm) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, size_t>, int> cache;
    static std::map<std::pair<int, const void *>, size_t> opted;  // largest dynamic size each kernel has been opted in for (only ever raised)
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({device, kern, smem});
    if (it != cache.end()) { *per_sm = it->second; return B200MP_OK; }
    if (smem > opted[{device, kern}]) {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        opted[{device, kern}] = smem;
    }
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, threads, smem));
    cache[{device, kern, smem}] = *per_sm;
    return B200MP_OK;
}

// Every launch gets its own work-queue head from a ring in the context, zeroed on the launch stream just before the kernel, so
// launches in flight on different streams (the _device entry points take any caller stream) never share or reset each other's queue.
LaunchCounters *next_counter(b200mp_ctx *c) {
    return static_cast<LaunchCounters *>(c->d_counters) + (c->counter_next.fetch_add(1) % kCounterRing) * kCounterStride;  // own 128-byte line
}


B200MP_EXTERN_ROBOT(Panda)
B200MP_EXTERN_ROBOT(Fetch)
B200MP_EXTERN_ROBOT(Ur5)
B200MP_EXTERN_ROBOT(Baxter)
}  // namespace b200mp

// ---------------------------------------------------------------------------------------------------- environment
namespace {

struct HSphere { float x, y, z, r, min_d; };
struct HCapsule { float x1, y1, z1, xv, yv, zv, r, rdv, min_d; };
struct HCuboid { float c[3], a1[3], a2[3], a3[3], h[3], min_d; };

struct DeviceCapt {
    float *tests = nullptr, *aabbs = nullptr, *aff = nullptr;
    uint32_t *aff_starts = nullptr;
    unsigned short *pref = nullptr;
    CaptView view{};
};

inline float dot3(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }
inline float clampf(float v, float lo, float hi) { return std::max(std::min(v, hi), lo); }

}  // namespace

struct b200mp_env {
    b200mp_ctx *ctx = nullptr;
    std::vector<HSphere> spheres;
    std::vector<HCapsule> capsules, zcapsules;
    std::vector<HCuboid> cuboids, zcuboids;
    std::vector<DeviceCapt> pcs;
    std::vector<HeightFieldView> hfs;  // data pointers are device allocations owned by the environment
    float *d_blob = nullptr;
    void *d_grid = nullptr;
    std::vector<float> att;  // attachment: 12 floats of pose + 4 pad, then (x, y, z, r) per sphere; empty = none
    float *d_att = nullptr;
    bool use_grid = true;
    SceneDesc desc{};
    bool committed = false;
    mutable std::atomic<int> open_sessions{0};  // resident kernels reading this environment's device arrays (b200mp_session_*)
};

namespace {

struct Box {
    float lo[3], hi[3];
};

// axis-aligned bounds of every obstacle, per class, in record order
void obstacle_boxes(const b200mp_env *e, std::vector<Box> (&boxes)[kShapeClasses], bool *regular) {
    *regular = true;
    for (const auto &s : e->spheres) boxes[0].push_back({{s.x - s.r, s.y - s.r, s.z - s.r}, {s.x + s.r, s.y + s.r, s.z + s.r}});
    auto capsule = [](const HCapsule &c) {
        Box b;
        const float p[3] = {c.x1, c.y1, c.z1}, v[3] = {c.xv, c.yv, c.zv};
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = std::min(p[k], p[k] + v[k]) - c.r;
            b.hi[k] = std::max(p[k], p[k] + v[k]) + c.r;
        }
        return b;
    };
    for (const auto &c : e->capsules) boxes[1].push_back(capsule(c));
    for (auto c : e->zcapsules) {
        c.xv = c.yv = 0.f;  // the z-aligned test only uses zv
        boxes[2].push_back(capsule(c));
    }
    auto unit = [](const float *a) { return std::fabs(dot3(a[0], a[1], a[2], a[0], a[1], a[2]) - 1.f) < 1e-3f; };
    auto ortho = [](const float *a, const float *b) { return std::fabs(dot3(a[0], a[1], a[2], b[0], b[1], b[2])) < 1e-3f; };
    for (const auto &c : e->cuboids) {
        if (!unit(c.a1) || !unit(c.a2) || !unit(c.a3) || !ortho(c.a1, c.a2) || !ortho(c.a1, c.a3) || !ortho(c.a2, c.a3)) *regular = false;
        Box b;
        for (int k = 0; k < 3; ++k) {
            const float ext = std::fabs(c.a1[k]) * c.h[0] + std::fabs(c.a2[k]) * c.h[1] + std::fabs(c.a3[k]) * c.h[2];
            b.lo[k] = c.c[k] - ext;
            b.hi[k] = c.c[k] + ext;
        }
        boxes[3].push_back(b);
    }
    for (const auto &c : e->zcuboids) {
        // the z-aligned test uses a1.xy, a2.xy and the world z axis
        const float n1 = c.a1[0] * c.a1[0] + c.a1[1] * c.a1[1], n2 = c.a2[0] * c.a2[0] + c.a2[1] * c.a2[1];
        if (std::fabs(n1 - 1.f) > 1e-3f || std::fabs(n2 - 1.f) > 1e-3f || std::fabs(c.a1[0] * c.a2[0] + c.a1[1] * c.a2[1]) > 1e-3f) *regular = false;
        Box b;
        for (int k = 0; k < 2; ++k) {
            const float ext = std::fabs(c.a1[k]) * c.h[0] + std::fabs(c.a2[k]) * c.h[1];
            b.lo[k] = c.c[k] - ext;
            b.hi[k] = c.c[k] + ext;
        }
        b.lo[2] = c.c[2] - c.h[2];
        b.hi[2] = c.c[2] + c.h[2];
        boxes[4].push_back(b);
    }
}

// Candidate grids of a sorted scene (see GridView). Returns B200MP_OK with d->n_grid = 0 when no grid is used.
int build_grids(b200mp_env *e, SceneDesc *d) {
    d->n_grid = 0;
    const size_t total = e->spheres.size() + e->capsules.size() + e->zcapsules.size() + e->cuboids.size() + e->zcuboids.size();
    if (!e->use_grid || !d->sorted || total < 32) return B200MP_OK;  // short lists: the ordered scan is already cheap
    std::vector<Box> boxes[kShapeClasses];
    bool regular = true;
    obstacle_boxes(e, boxes, &regular);
    if (!regular) return B200MP_OK;  // a cuboid with non-orthonormal axes has no simple bounding box: keep the full scan
    for (const auto &cls : boxes)
        if (cls.size() > 65535) return B200MP_OK;
    // reach ladder: x 1.12 per level from 2.5 cm, the last level 36 cm -- a robot sphere walks the level just above its radius, so the
    // lists it sees hold little more than what it can actually touch
    float reach[kGridLevels];
    for (int l = 0; l < kGridLevels; ++l) reach[l] = l == kGridLevels - 1 ? 0.36f : 0.025f * std::pow(1.12f, static_cast<float>(l));
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (const auto &cls : boxes)
        for (const auto &b : cls)
            for (int k = 0; k < 3; ++k) {
                if (!std::isfinite(b.lo[k]) || !std::isfinite(b.hi[k])) return B200MP_OK;
                lo[k] = std::min(lo[k], b.lo[k]);
                hi[k] = std::max(hi[k], b.hi[k]);
            }
    // cell size: at most ~32k cells, never below 6 cm
    float h = 0.06f;
    for (;; h *= 1.15f) {
        double cells = 1;
        for (int k = 0; k < 3; ++k) cells *= std::floor((hi[k] - lo[k]) / h) + 1;
        if (cells <= 32768) break;
    }
    int n[3];
    for (int k = 0; k < 3; ++k) n[k] = static_cast<int>(std::floor((hi[k] - lo[k]) / h)) + 1;
    const size_t cells = static_cast<size_t>(n[0]) * n[1] * n[2];
    const float inv_h = 1.0f / h;
    auto cell_of = [&](float v, int k) {  // the device's grid_cell(), same float operations
        const int c = static_cast<int>(std::floor((v - lo[k]) * inv_h));
        return std::min(std::max(c, 0), n[k] - 1);
    };
    // layout of the device buffer: for level, class: one 8-byte record per cell (count, where the list goes on, its first two
    // entries), then the further entries of the longer lists (u16 record indices); 16-byte aligned pieces
    struct Piece { size_t start_off, ids_off; std::vector<unsigned> start; std::vector<unsigned short> ids; };
    std::vector<Piece> pieces(kGridLevels * kShapeClasses);
    size_t bytes = 0;
    for (int l = 0; l < kGridLevels; ++l)
        for (int t = 0; t < kShapeClasses; ++t) {
            Piece &p = pieces[l * kShapeClasses + t];
            const auto &cls = boxes[t];
            if (cls.empty()) continue;
            const float grow = reach[l] * 1.0001f + 3e-4f;  // slack covers kHitMargin (bounding spheres are inflated by 0.1 mm) and rounding in the device-side tests
            // A cell lists obstacle i when the obstacle can be within `grow` of SOME point of the cell: Euclidean distance between the
            // cell's box (unbounded on the outer side of border cells, which stand in for everything beyond the grid) and the
            // obstacle's bounding box -- for spheres the distance to the centre minus the radius. Candidate cells = the box range.
            const bool is_sphere = t == 0;
            auto gap = [&](int c, int k, float blo, float bhi) {  // distance along axis k between cell c and [blo, bhi]
                const float clo = c == 0 ? -INFINITY : lo[k] + static_cast<float>(c) * h;
                const float chi = c == n[k] - 1 ? INFINITY : lo[k] + static_cast<float>(c + 1) * h;
                return std::max(0.0f, std::max(blo - chi, clo - bhi));
            };
            p.start.assign(cells + 1, 0u);
            std::vector<int> rng(cls.size() * 6);
            std::vector<std::pair<unsigned, unsigned short>> hits;  // (cell, obstacle)
            for (size_t i = 0; i < cls.size(); ++i) {
                const Box &b = cls[i];
                const float rad = is_sphere ? 0.5f * (b.hi[0] - b.lo[0]) : 0.0f;
                float blo[3], bhi[3];
                for (int k = 0; k < 3; ++k) {
                    rng[6 * i + k] = cell_of(b.lo[k] - grow, k);
                    rng[6 * i + 3 + k] = cell_of(b.hi[k] + grow, k);
                    blo[k] = is_sphere ? 0.5f * (b.lo[k] + b.hi[k]) : b.lo[k];
                    bhi[k] = is_sphere ? blo[k] : b.hi[k];
                }
                const float lim = (grow + rad) * 1.0001f + 1e-5f;
                for (int z = rng[6 * i + 2]; z <= rng[6 * i + 5]; ++z) {
                    const float gz = gap(z, 2, blo[2], bhi[2]);
                    for (int y = rng[6 * i + 1]; y <= rng[6 * i + 4]; ++y) {
                        const float gy = gap(y, 1, blo[1], bhi[1]);
                        for (int x = rng[6 * i]; x <= rng[6 * i + 3]; ++x) {
                            const float gx = gap(x, 0, blo[0], bhi[0]);
                            if (gx * gx + gy * gy + gz * gz > lim * lim) continue;
                            const unsigned cell = static_cast<unsigned>((static_cast<size_t>(z) * n[1] + y) * n[0] + x);
                            ++p.start[cell + 1];
                            hits.emplace_back(cell, static_cast<unsigned short>(i));
                        }
                    }
                }
            }
            // counts -> the per-cell record: x = (position in `ids` of the list's THIRD entry << 8) | entries, y = the first two entries
            // (the padding record's index where the list is shorter); `ids` holds entries 2, 3, ... of every list
            {
                const unsigned pad = static_cast<unsigned>(cls.size());
                std::vector<unsigned> off(cells + 1, 0u);
                for (size_t c = 0; c < cells; ++c) {
                    if (p.start[c + 1] > 255u) return B200MP_OK;  // a cell this crowded: keep the ordered scan for the whole scene
                    off[c + 1] = off[c] + (p.start[c + 1] > 2u ? p.start[c + 1] - 2u : 0u);
                }
                if (off[cells] >= (1u << 24)) return B200MP_OK;
                std::vector<unsigned> cnt(cells);
                for (size_t c = 0; c < cells; ++c) cnt[c] = p.start[c + 1];
                p.ids.assign(off[cells], 0);
                p.start.assign(cells * 2, 0u);
                for (size_t c = 0; c < cells; ++c) {
                    p.start[2 * c] = (off[c] << 8) | cnt[c];  // (the device reads entry i >= 2 from ids[(x >> 8) + i - 2])
                    p.start[2 * c + 1] = pad * 0x10001u;
                }
                std::vector<unsigned> seen(cells, 0u);
                for (const auto &hc : hits) {  // (obstacles were visited in ascending order: so is every cell's list)
                    const unsigned k = seen[hc.first]++;
                    if (k == 0) p.start[2 * hc.first + 1] = (p.start[2 * hc.first + 1] & 0xffff0000u) | hc.second;
                    else if (k == 1) p.start[2 * hc.first + 1] = (p.start[2 * hc.first + 1] & 0x0000ffffu) | (static_cast<unsigned>(hc.second) << 16);
                    else p.ids[off[hc.first] + k - 2] = hc.second;
                }
            }
            p.start_off = bytes;
            bytes += (cells * 2 * sizeof(unsigned) + 15) / 16 * 16;
            p.ids_off = bytes;
            bytes += (p.ids.size() * sizeof(unsigned short) + 15) / 16 * 16 + 16;
        }
    if (bytes > (size_t(512) << 20)) return B200MP_OK;
    std::vector<unsigned char> host(bytes, 0);
    for (const auto &p : pieces) {
        if (p.start.empty()) continue;
        std::memcpy(host.data() + p.start_off, p.start.data(), p.start.size() * sizeof(unsigned));
        if (!p.ids.empty()) std::memcpy(host.data() + p.ids_off, p.ids.data(), p.ids.size() * sizeof(unsigned short));
    }
    B200_CUDA(cudaMalloc(&e->d_grid, std::max<size_t>(bytes, 16)));
    B200_CUDA(cudaMemcpy(e->d_grid, host.data(), bytes, cudaMemcpyHostToDevice));
    const unsigned char *base = static_cast<const unsigned char *>(e->d_grid);
    for (int l = 0; l < kGridLevels; ++l) {
        GridView &g = d->grid[l];
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
        g.inv_h = inv_h;
        g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
        g.reach = reach[l];
        for (int t = 0; t < kShapeClasses; ++t) {
            const Piece &p = pieces[l * kShapeClasses + t];
            g.cellw[t] = p.start.empty() ? nullptr : reinterpret_cast<const uint2 *>(base + p.start_off);
            g.ids[t] = p.start.empty() ? nullptr : reinterpret_cast<const unsigned short *>(base + p.ids_off);
        }
    }
    d->n_grid = kGridLevels;
    return B200MP_OK;
}

}  // namespace

extern "C" {

int b200mp_version(void) { return B200MP_VERSION; }
const char *b200mp_last_error(void) { return g_last_error.c_str(); }

int b200mp_create(int device, b200mp_ctx **out) {
    if (!out) return fail(B200MP_ERR_INVALID, "b200mp_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200MP_ERR_NO_DEVICE, "b200mp_create: no CUDA device (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(B200MP_ERR_INVALID, "b200mp_create: device %d out of range [0,%d)", device, n);
    struct Destroy { void operator()(b200mp_ctx *p) const { b200mp_destroy(p); } };
    std::unique_ptr<b200mp_ctx, Destroy> guard(new b200mp_ctx());  // a failing step below releases what was created so far
    b200mp_ctx *c = guard.get();
    c->device = device;
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    c->cc_major = prop.major;
    if (prop.major < 10) {
        return fail(B200MP_ERR_NO_DEVICE, "b200mp_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        // stream-ordered temporaries (bucketed validation, roadmap phases) come from the device's default pool: keep what it has
        // acquired across synchronisations instead of handing it back to the driver after every call
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    B200_CUDA(cudaMalloc(&c->d_counters, sizeof(LaunchCounters) * kCounterRing * kCounterStride));
    c->small_cap = size_t(256) << 10;
    B200_CUDA(cudaHostAlloc(&c->h_small, c->small_cap, cudaHostAllocMapped));
    B200_CUDA(cudaHostGetDevicePointer(&c->d_small, c->h_small, 0));
    for (auto &ln : c->lanes) {
        B200_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        B200_CUDA(cudaHostAlloc(&ln.h, b200mp_ctx::kLaneBytes, cudaHostAllocMapped));
        B200_CUDA(cudaHostGetDevicePointer(&ln.d, ln.h, 0));
    }
    *out = guard.release();
    return B200MP_OK;
}

void b200mp_destroy(b200mp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->self_comm) b200mp_comm_destroy(static_cast<b200mp_comm *>(c->self_comm));
    c->scratch.release();
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto &ln : c->lanes) {
        if (ln.stream) { cudaStreamSynchronize(ln.stream); cudaStreamDestroy(ln.stream); }
        if (ln.h) cudaFreeHost(ln.h);
    }
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int b200mp_device(const b200mp_ctx *c, int *device, int *sm_count) {
    if (!c) return fail(B200MP_ERR_INVALID, "b200mp_device: ctx is NULL");
    if (device) *device = c->device;
    if (sm_count) *sm_count = c->sm_count;
    return B200MP_OK;
}

int b200mp_synchronize(b200mp_ctx *c) {
    if (!c) return fail(B200MP_ERR_INVALID, "b200mp_synchronize: ctx is NULL");
    B200_CUDA(cudaSetDevice(c->device));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return B200MP_OK;
}

int64_t b200mp_launch_count(const b200mp_ctx *c) { return c ? c->launches.load() : -1; }

int b200mp_set_option(b200mp_ctx *c, int option, int value) {
    if (!c) return fail(B200MP_ERR_INVALID, "b200mp_set_option: ctx is NULL");
    switch (option) {
        case B200MP_OPT_FIRST_LOOK: c->first_look.store(value != 0 ? 1 : 0); return B200MP_OK;
        default: return fail(B200MP_ERR_INVALID, "b200mp_set_option: unknown option %d", option);
    }
}

int b200mp_robot_info(int robot, int *dim, int *n_spheres, int *resolution, float *lo, float *hi) {
    auto fill = [&](auto tag) {
        using R = decltype(tag);
        if (dim) *dim = R::kDim;
        if (n_spheres) *n_spheres = R::kFine;
        if (resolution) *resolution = R::kResolution;
        for (int j = 0; j < R::kDim; ++j) {
            if (lo) lo[j] = R::kLo[j];
            if (hi) hi[j] = R::kHi[j];
        }
    };
    switch (robot) {
        case B200MP_ROBOT_PANDA: fill(Panda{}); return B200MP_OK;
        case B200MP_ROBOT_FETCH: fill(Fetch{}); return B200MP_OK;
        case B200MP_ROBOT_UR5: fill(Ur5{}); return B200MP_OK;
        case B200MP_ROBOT_BAXTER: fill(Baxter{}); return B200MP_OK;
        default: return fail(B200MP_ERR_INVALID, "b200mp_robot_info: unknown robot %d", robot);
    }
}

// ------------------------------------------------------------------------------------------------ environment
int b200mp_env_create(b200mp_ctx *ctx, b200mp_env **out) {
    if (!ctx || !out) return fail(B200MP_ERR_INVALID, "b200mp_env_create: NULL argument");
    auto *e = new b200mp_env();
    e->ctx = ctx;
    *out = e;
    return B200MP_OK;
}

static void env_free_device(b200mp_env *e) {
    if (e->d_blob) cudaFree(e->d_blob);
    if (e->d_grid) cudaFree(e->d_grid);
    e->d_blob = nullptr;
    e->d_grid = nullptr;
    e->committed = false;
}

void b200mp_env_destroy(b200mp_env *e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    env_free_device(e);
    for (auto &p : e->pcs) {
        cudaFree(p.tests);
        cudaFree(p.aabbs);
        cudaFree(p.aff);
        cudaFree(p.aff_starts);
        cudaFree(p.pref);
    }
    for (auto &h : e->hfs) cudaFree(const_cast<float *>(h.data));
    if (e->d_att) cudaFree(e->d_att);
    delete e;
}

// vamp::collision::Attachment (collision/attachments.hh:13-56) + Environment::attachments (environment.hh:25). May be called before
// or after commit: the attachment lives in its own device buffer, so grasping / releasing an object does not re-upload the scene.
int b200mp_env_set_attachment(b200mp_env *e, const float *tf, const float *xyzr, int n) {
    if (!e || n < 0 || (n > 0 && (!tf || !xyzr))) return fail(B200MP_ERR_INVALID, "b200mp_env_set_attachment: bad argument");
    if (e->open_sessions.load() > 0)
        return fail(B200MP_ERR_STATE, "b200mp_env_set_attachment: %d session(s) are open on this environment (close them first: a resident kernel reads these arrays)",
                    e->open_sessions.load());
    B200_CUDA(cudaSetDevice(e->ctx->device));
    if (e->d_att) B200_CUDA(cudaFree(e->d_att));  // (cudaFree waits for launches that still read it)
    e->d_att = nullptr;
    e->att.clear();
    if (n > 0) {
        e->att.assign(tf, tf + 12);
        e->att.resize(16, 0.f);
        e->att.insert(e->att.end(), xyzr, xyzr + static_cast<size_t>(n) * 4);
        B200_CUDA(cudaMalloc(&e->d_att, e->att.size() * 4));
        B200_CUDA(cudaMemcpy(e->d_att, e->att.data(), e->att.size() * 4, cudaMemcpyHostToDevice));
    }
    e->desc.att = e->d_att;
    e->desc.n_att = n;
    return B200MP_OK;
}

int b200mp_env_add_spheres(b200mp_env *e, const float *s, int n) {
    if (!e || (!s && n > 0) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_env_add_spheres: bad argument");
    for (int i = 0; i < n; ++i) {
        HSphere h{s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3], 0};
        h.min_d = std::sqrt(h.x * h.x + h.y * h.y + h.z * h.z) - h.r;  // shapes.hh:236-239
        e->spheres.push_back(h);
    }
    e->committed = false;
    return B200MP_OK;
}

int b200mp_env_add_cuboids(b200mp_env *e, const float *c, int n, int z_aligned) {
    if (!e || (!c && n > 0) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_env_add_cuboids: bad argument");
    for (int i = 0; i < n; ++i) {
        const float *p = c + 15 * i;
        HCuboid h{};
        std::memcpy(h.c, p, 12);
        std::memcpy(h.a1, p + 3, 12);
        std::memcpy(h.a2, p + 6, 12);
        std::memcpy(h.a3, p + 9, 12);
        std::memcpy(h.h, p + 12, 12);
        // closest point of the box to the world origin (shapes.hh:52-67)
        const float d1 = dot3(-h.c[0], -h.c[1], -h.c[2], h.a1[0], h.a1[1], h.a1[2]);
        const float d2 = dot3(-h.c[0], -h.c[1], -h.c[2], h.a2[0], h.a2[1], h.a2[2]);
        const float d3 = dot3(-h.c[0], -h.c[1], -h.c[2], h.a3[0], h.a3[1], h.a3[2]);
        const float v1 = clampf(d1, -h.h[0], h.h[0]), v2 = clampf(d2, -h.h[1], h.h[1]), v3 = clampf(d3, -h.h[2], h.h[2]);
        const float xn = h.c[0] + h.a1[0] * v1 + h.a2[0] * v2 + h.a3[0] * v3;
        const float yn = h.c[1] + h.a1[1] * v1 + h.a2[1] * v2 + h.a3[1] * v3;
        const float zn = h.c[2] + h.a1[2] * v1 + h.a2[2] * v2 + h.a3[2] * v3;
        h.min_d = std::sqrt(xn * xn + yn * yn + zn * zn);
        (z_aligned ? e->zcuboids : e->cuboids).push_back(h);
    }
    e->committed = false;
    return B200MP_OK;
}

int b200mp_env_add_capsules(b200mp_env *e, const float *c, int n, int z_aligned) {
    if (!e || (!c && n > 0) || n < 0) return fail(B200MP_ERR_INVALID, "b200mp_env_add_capsules: bad argument");
    for (int i = 0; i < n; ++i) {
        const float *p = c + 7 * i;
        HCapsule h{p[0], p[1], p[2], p[3], p[4], p[5], p[6], 0, 0};
        h.rdv = static_cast<float>(1.0 / dot3(h.xv, h.yv, h.zv, h.xv, h.yv, h.zv));  // factory.hh:110-126
        // closest point of the capsule to the world origin (shapes.hh:165-189)
        const float t = clampf(dot3(-h.x1, -h.y1, -h.z1, h.xv, h.yv, h.zv) * h.rdv, 0.F, 1.F);
        const float xp = h.x1 + h.xv * t, yp = h.y1 + h.yv * t, zp = h.z1 + h.zv * t;
        float xo = -xp, yo = -yp, zo = -zp;
        const float ol = std::sqrt(dot3(xo, yo, zo, xo, yo, zo));
        xo = xo / ol;
        yo = yo / ol;
        zo = zo / ol;
        const float ro = clampf(ol, 0.F, h.r);
        const float xn = xp + ro * xo, yn = yp + ro * yo, zn = zp + ro * zo;
        h.min_d = std::sqrt(xn * xn + yn * yn + zn * zn);
        (z_aligned ? e->zcapsules : e->capsules).push_back(h);
    }
    e->committed = false;
    return B200MP_OK;
}

int b200mp_env_add_heightfield(b200mp_env *e, float cx, float cy, float cz, float sx, float sy, float sz, int64_t nx, int64_t ny,
                               const float *heights) {
    if (!e || !heights || nx < 1 || ny < 1 || nx * ny > (int64_t(1) << 30) || !(sx != 0.f) || !(sy != 0.f) || !(sz != 0.f))
        return fail(B200MP_ERR_INVALID, "b200mp_env_add_heightfield: bad argument");
    if (e->hfs.size() >= kMaxHeightFields) return fail(B200MP_ERR_CAPACITY, "b200mp_env_add_heightfield: at most %d height fields", kMaxHeightFields);
    B200_CUDA(cudaSetDevice(e->ctx->device));
    HeightFieldView h{};
    h.x = cx; h.y = cy; h.z = cz;
    h.xs = 1.F / sx; h.ys = 1.F / sy; h.zs = 1.F / sz;  // factory::heightfield::flat stores reciprocals (factory.hh:376-385)
    h.xd = static_cast<float>(nx); h.yd = static_cast<float>(ny);
    h.xd2 = static_cast<float>(nx / 2); h.yd2 = static_cast<float>(ny / 2);
    h.n = static_cast<unsigned>(nx * ny);
    float *d = nullptr;
    B200_CUDA(cudaMalloc(&d, static_cast<size_t>(nx * ny) * sizeof(float)));
    B200_CUDA(cudaMemcpy(d, heights, static_cast<size_t>(nx * ny) * sizeof(float), cudaMemcpyHostToDevice));
    h.data = d;
    e->hfs.push_back(h);
    e->committed = false;
    return B200MP_OK;
}

// d.tests / d.aabbs / d.aff_starts / d.aff are device arrays in the reference's layout (ownership passes to the environment):
// order every leaf's list by distance from its cell, build the prefix table, attach
static int env_attach_capt(b200mp_env *e, DeviceCapt d, int nlog2, int64_t n_aff8, const float *top, float r_point) {
    const size_t nl = size_t(1) << nlog2, nt = nl - 1;
    if (n_aff8 > 0) {
        // device copy: lists ordered by distance from the leaf's cell + prefix table (capt_order.cuh)
        float *boxes = nullptr, *sorted = nullptr;
        B200_CUDA(cudaMalloc(&boxes, nl * 6 * 4));
        B200_CUDA(cudaMalloc(&sorted, static_cast<size_t>(3) * n_aff8 * 8 * 4));
        B200_CUDA(cudaMalloc(&d.pref, nl * kCaptBuckets * sizeof(unsigned short)));
        capt_cell_boxes_kernel<<<static_cast<unsigned>((nl + 255) / 256), 256>>>(d.tests, nlog2, static_cast<unsigned>(nl), boxes);
        capt_order_leaf_kernel<<<static_cast<unsigned>(nl), 256>>>(d.aff, sorted, n_aff8, d.aff_starts, boxes, d.pref);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaDeviceSynchronize());
        e->ctx->launches.fetch_add(2);
        cudaFree(boxes);
        cudaFree(d.aff);
        d.aff = sorted;
    }
    d.view.tests = d.tests;
    d.view.aabbs = d.aabbs;
    d.view.aff_starts = d.aff_starts;
    d.view.aff = d.aff;
    d.view.pref = d.pref;
    d.view.n_aff8 = n_aff8;
    std::memcpy(d.view.top, top, 24);
    d.view.r_point = r_point;
    d.view.nlog2 = nlog2;
    d.view.n_tests = static_cast<int>(nt);
    e->pcs.push_back(d);
    e->committed = false;
    return B200MP_OK;
}

int b200mp_env_add_capt(b200mp_env *e, int nlog2, const float *tests, const float *aabbs, const uint32_t *aff_starts,
                        const float *aff, int64_t n_aff8, const float *top, float r_min, float r_max, float r_point) {
    (void)r_min;
    (void)r_max;
    if (!e || (!tests && nlog2 > 0) || !aabbs || !aff_starts || !aff || !top || nlog2 < 0 || nlog2 > 30 || n_aff8 < 0)
        return fail(B200MP_ERR_INVALID, "b200mp_env_add_capt: bad argument");
    if (e->pcs.size() >= kMaxPointClouds) return fail(B200MP_ERR_CAPACITY, "b200mp_env_add_capt: at most %d point clouds", kMaxPointClouds);
    B200_CUDA(cudaSetDevice(e->ctx->device));
    const size_t nl = size_t(1) << nlog2, nt = nl - 1;
    DeviceCapt d;
    B200_CUDA(cudaMalloc(&d.tests, std::max<size_t>(nt, 1) * 4));
    B200_CUDA(cudaMalloc(&d.aabbs, nl * 6 * 4));
    B200_CUDA(cudaMalloc(&d.aff_starts, (nl + 1) * 4));
    B200_CUDA(cudaMalloc(&d.aff, std::max<size_t>(3 * n_aff8 * 8, 4) * 4));
    if (nt) B200_CUDA(cudaMemcpy(d.tests, tests, nt * 4, cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(d.aabbs, aabbs, nl * 6 * 4, cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(d.aff_starts, aff_starts, (nl + 1) * 4, cudaMemcpyHostToDevice));
    if (n_aff8) B200_CUDA(cudaMemcpy(d.aff, aff, 3 * n_aff8 * 8 * 4, cudaMemcpyHostToDevice));
    return env_attach_capt(e, d, nlog2, n_aff8, top, r_point);
}

int b200mp_env_add_pointcloud_device(b200mp_env *e, const float *xyz, int64_t n, float r_min, float r_max, float r_point) {
    (void)r_min;
    if (!e || n < 1 || !xyz || !(r_max > 0.f) || !(r_point >= 0.f))
        return fail(B200MP_ERR_INVALID, "b200mp_env_add_pointcloud_device: bad argument");
    if (n > (int64_t(1) << 17))
        return fail(B200MP_ERR_CAPACITY, "b200mp_env_add_pointcloud_device: at most 2^17 points (filter the cloud first, or use the host build)");
    if (e->pcs.size() >= kMaxPointClouds) return fail(B200MP_ERR_CAPACITY, "b200mp_env_add_pointcloud_device: at most %d point clouds", kMaxPointClouds);
    B200_CUDA(cudaSetDevice(e->ctx->device));
    int nlog2 = 0;
    while ((int64_t(1) << nlog2) < n) ++nlog2;
    const size_t nl = size_t(1) << nlog2, nt = nl - 1;
    const unsigned nb = static_cast<unsigned>((n + 255) / 256);
    float *d_xyz = nullptr, *box_a = nullptr, *box_b = nullptr, *cells = nullptr, *d_top = nullptr;
    int *node = nullptr;
    unsigned *hist = nullptr, *ext = nullptr;
    uint32_t *counts8 = nullptr;
    DeviceCapt d;
    B200_CUDA(cudaMalloc(&d_xyz, static_cast<size_t>(n) * 12));
    B200_CUDA(cudaMemcpy(d_xyz, xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice));
    B200_CUDA(cudaMalloc(&node, static_cast<size_t>(n) * 4));
    B200_CUDA(cudaMemset(node, 0, static_cast<size_t>(n) * 4));
    B200_CUDA(cudaMalloc(&box_a, nl * 6 * 4));
    B200_CUDA(cudaMalloc(&box_b, nl * 6 * 4));
    B200_CUDA(cudaMalloc(&cells, nl * 6 * 4));
    B200_CUDA(cudaMalloc(&hist, std::max<size_t>(nl / 2, 1) * kCgBins * 4));
    B200_CUDA(cudaMalloc(&ext, 6 * 4));
    B200_CUDA(cudaMalloc(&d_top, 6 * 4));
    B200_CUDA(cudaMalloc(&counts8, nl * 4));
    B200_CUDA(cudaMalloc(&d.tests, std::max<size_t>(nt, 1) * 4));
    B200_CUDA(cudaMalloc(&d.aabbs, nl * 6 * 4));
    B200_CUDA(cudaMalloc(&d.aff_starts, (nl + 1) * 4));
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    B200_CUDA(cudaMemcpy(ext, init, sizeof(init), cudaMemcpyHostToDevice));
    cg_extent_kernel<<<nb, 256>>>(d_xyz, n, ext);
    cg_root_kernel<<<1, 32>>>(ext, box_a, d_top);
    long long launches = 2;
    for (int l = 0; l < nlog2; ++l) {
        const int axis = l % 3;
        B200_CUDA(cudaMemset(hist, 0, (static_cast<size_t>(1) << l) * kCgBins * 4));
        cg_hist_kernel<<<nb, 256>>>(d_xyz, n, node, axis, box_a, hist);
        cg_split_kernel<<<((1 << l) + 63) / 64, 64>>>(hist, l, axis, box_a, box_b, d.tests);
        cg_descend_kernel<<<nb, 256>>>(d_xyz, n, node, l, axis, d.tests);
        std::swap(box_a, box_b);
        launches += 3;
    }
    // the cells the QUERY sees: cut out by the planes alone, unbounded at the rim (a sphere centre can be anywhere)
    capt_cell_boxes_kernel<<<static_cast<unsigned>((nl + 255) / 256), 256>>>(d.tests, nlog2, static_cast<unsigned>(nl), cells);
    const float reach = (r_max + r_point) * 1.00001f + 1.0e-6f;
    cg_lists_kernel<false><<<static_cast<unsigned>(nl), 128>>>(d_xyz, static_cast<int>(n), cells, reach * reach, counts8, nullptr, nullptr, 0, nullptr);
    cg_scan_kernel<<<1, 1024>>>(counts8, static_cast<unsigned>(nl), d.aff_starts);
    uint32_t total8 = 0;
    B200_CUDA(cudaMemcpy(&total8, d.aff_starts + nl, 4, cudaMemcpyDeviceToHost));
    const int64_t n_aff8 = total8;
    B200_CUDA(cudaMalloc(&d.aff, std::max<size_t>(static_cast<size_t>(3) * n_aff8 * 8, 4) * 4));
    cg_lists_kernel<true><<<static_cast<unsigned>(nl), 128>>>(d_xyz, static_cast<int>(n), cells, reach * reach, nullptr, d.aff_starts, d.aff, n_aff8, d.aabbs);
    float top[6];
    B200_CUDA(cudaMemcpy(top, d_top, sizeof(top), cudaMemcpyDeviceToHost));
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaDeviceSynchronize());
    e->ctx->launches.fetch_add(launches + 4);
    cudaFree(d_xyz); cudaFree(node); cudaFree(box_a); cudaFree(box_b); cudaFree(cells); cudaFree(hist); cudaFree(ext); cudaFree(d_top); cudaFree(counts8);
    return env_attach_capt(e, d, nlog2, n_aff8, top, r_point);
}

int b200mp_env_clone(const b200mp_env *src, b200mp_ctx *dst_ctx, b200mp_env **out) {
    if (!src || !dst_ctx || !out) return fail(B200MP_ERR_INVALID, "b200mp_env_clone: NULL argument");
    *out = nullptr;
    if (!src->pcs.empty() || !src->hfs.empty())
        return fail(B200MP_ERR_CAPACITY, "b200mp_env_clone: scenes with point clouds or height fields own device arrays and are not cloned");
    auto *e = new b200mp_env();
    e->ctx = dst_ctx;
    e->spheres = src->spheres;
    e->capsules = src->capsules;
    e->zcapsules = src->zcapsules;
    e->cuboids = src->cuboids;
    e->zcuboids = src->zcuboids;
    e->use_grid = src->use_grid;
    if (!src->att.empty()) {
        const int rc = b200mp_env_set_attachment(e, src->att.data(), src->att.data() + 16, static_cast<int>((src->att.size() - 16) / 4));
        if (rc) { b200mp_env_destroy(e); return rc; }
    }
    if (src->committed) {
        const int rc = b200mp_env_commit(e);
        if (rc) { b200mp_env_destroy(e); return rc; }
    }
    *out = e;
    return B200MP_OK;
}

int b200mp_env_set_grid(b200mp_env *e, int enable) {
    if (!e) return fail(B200MP_ERR_INVALID, "b200mp_env_set_grid: env is NULL");
    e->use_grid = enable != 0;
    e->committed = false;
    return B200MP_OK;
}

int b200mp_env_sort(b200mp_env *e) {
    if (!e) return fail(B200MP_ERR_INVALID, "b200mp_env_sort: env is NULL");
    auto by = [](const auto &a, const auto &b) { return a.min_d < b.min_d; };  // environment.hh:43-69
    std::sort(e->spheres.begin(), e->spheres.end(), by);
    std::sort(e->capsules.begin(), e->capsules.end(), by);
    std::sort(e->zcapsules.begin(), e->zcapsules.end(), by);
    std::sort(e->cuboids.begin(), e->cuboids.end(), by);
    std::sort(e->zcuboids.begin(), e->zcuboids.end(), by);
    e->committed = false;
    return B200MP_OK;
}

int64_t b200mp_env_min_distances(const b200mp_env *e, int kind, float *out) {
    if (!e) return -1;
    auto dump = [&](const auto &v) -> int64_t {
        if (out) for (size_t i = 0; i < v.size(); ++i) out[i] = v[i].min_d;
        return static_cast<int64_t>(v.size());
    };
    switch (kind) {
        case 0: return dump(e->spheres);
        case 1: return dump(e->capsules);
        case 2: return dump(e->zcapsules);
        case 3: return dump(e->cuboids);
        case 4: return dump(e->zcuboids);
        default: return -1;
    }
}

int b200mp_env_commit(b200mp_env *e) {
    if (!e) return fail(B200MP_ERR_INVALID, "b200mp_env_commit: env is NULL");
    if (e->open_sessions.load() > 0)
        return fail(B200MP_ERR_STATE, "b200mp_env_commit: %d session(s) are open on this environment (close them first: a resident kernel reads these arrays)",
                    e->open_sessions.load());
    B200_CUDA(cudaSetDevice(e->ctx->device));
    std::vector<float> blob;
    auto align4 = [&]() { while (blob.size() % 4) blob.push_back(0.f); };
    // prefmax[i] = max(min_d[0..i]): the ordered scan of a sphere with extent `ext` reaches obstacle i
    // iff prefmax[i-1] < ext ... i.e. it tests exactly the first #{prefmax < ext} obstacles (validity.hh:61-67).
    auto push_pm = [&](const auto &v) {
        float m = -INFINITY;
        for (const auto &o : v) {
            m = std::max(m, o.min_d);
            blob.push_back(m);
        }
        align4();
    };
    SceneDesc d{};
    d.n_sph = static_cast<int>(e->spheres.size());
    d.n_cap = static_cast<int>(e->capsules.size());
    d.n_zcap = static_cast<int>(e->zcapsules.size());
    d.n_cub = static_cast<int>(e->cuboids.size());
    d.n_zcub = static_cast<int>(e->zcuboids.size());
    // every class is padded to a multiple of kObsUnroll records with far-away dummies that can never collide,
    // so the kernels scan whole blocks without a tail
    constexpr float kFar = 1.0e9f;
    // (at least one per non-empty class: record n is the stand-in the grid-culled scans use for exhausted lists)
    auto pad_count = [](size_t n) { return n == 0 ? 0 : kObsUnroll - static_cast<int>(n % kObsUnroll); };
    bool sorted = true;
    auto is_sorted = [&](const auto &v) {
        for (size_t i = 1; i < v.size(); ++i) if (v[i].min_d < v[i - 1].min_d) sorted = false;
    };
    is_sorted(e->spheres); is_sorted(e->capsules); is_sorted(e->zcapsules); is_sorted(e->cuboids); is_sorted(e->zcuboids);
    d.sorted = sorted ? 1 : 0;

    d.off_sph = static_cast<int>(blob.size());
    for (const auto &s : e->spheres) blob.insert(blob.end(), {s.x, s.y, s.z, s.r});
    for (int i = 0; i < pad_count(e->spheres.size()); ++i) blob.insert(blob.end(), {kFar, kFar, kFar, 0.f});
    d.off_sph_pm = static_cast<int>(blob.size());
    push_pm(e->spheres);
    d.off_cap = static_cast<int>(blob.size());
    for (const auto &c : e->capsules) blob.insert(blob.end(), {c.x1, c.y1, c.z1, c.r, c.xv, c.yv, c.zv, c.rdv});
    for (int i = 0; i < pad_count(e->capsules.size()); ++i) blob.insert(blob.end(), {kFar, kFar, kFar, 0.f, 0.f, 0.f, 0.f, 0.f});
    d.off_cap_pm = static_cast<int>(blob.size());
    push_pm(e->capsules);
    d.off_zcap = static_cast<int>(blob.size());
    for (const auto &c : e->zcapsules) blob.insert(blob.end(), {c.x1, c.y1, c.z1, c.r, c.zv, c.rdv, 0.f, 0.f});
    for (int i = 0; i < pad_count(e->zcapsules.size()); ++i) blob.insert(blob.end(), {kFar, kFar, kFar, 0.f, 0.f, 0.f, 0.f, 0.f});
    d.off_zcap_pm = static_cast<int>(blob.size());
    push_pm(e->zcapsules);
    d.off_cub = static_cast<int>(blob.size());
    for (const auto &c : e->cuboids)
        blob.insert(blob.end(), {c.c[0], c.c[1], c.c[2], c.h[0], c.a1[0], c.a1[1], c.a1[2], c.h[1], c.a2[0], c.a2[1], c.a2[2], c.h[2],
                                 c.a3[0], c.a3[1], c.a3[2], 0.f});
    for (int i = 0; i < pad_count(e->cuboids.size()); ++i)
        blob.insert(blob.end(), {kFar, kFar, kFar, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f});
    d.off_cub_pm = static_cast<int>(blob.size());
    push_pm(e->cuboids);
    d.off_zcub = static_cast<int>(blob.size());
    for (const auto &c : e->zcuboids)
        blob.insert(blob.end(), {c.c[0], c.c[1], c.c[2], c.h[2], c.a1[0], c.a1[1], c.h[0], 0.f, c.a2[0], c.a2[1], c.h[1], 0.f});
    for (int i = 0; i < pad_count(e->zcuboids.size()); ++i)
        blob.insert(blob.end(), {kFar, kFar, kFar, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f});
    d.off_zcub_pm = static_cast<int>(blob.size());
    push_pm(e->zcuboids);
    align4();
    d.blob_floats = static_cast<int>(blob.size());
    const size_t bytes = blob.size() * 4;
    // dynamic shared memory budget: opt-in maximum minus static use (barrier) and a margin
    // and the per-thread bounding-sphere centres the kernels keep beside the scene
    // (checked here for the leanest robot; a launch for a robot whose slots do not fit beside this scene fails with B200MP_ERR_CAPACITY)
    const size_t budget = static_cast<size_t>(e->ctx->max_smem_optin) - 1024 -
                          std::min({lane_store_bytes<Panda>(kStateThreads), lane_store_bytes<Fetch>(kStateThreads), lane_store_bytes<Ur5>(kStateThreads)});
    if (bytes > budget)
        return fail(B200MP_ERR_CAPACITY, "b200mp_env_commit: scene needs %zu bytes of shared memory, limit %zu", bytes, budget);
    env_free_device(e);
    B200_CUDA(cudaMalloc(&e->d_blob, std::max<size_t>(bytes, 16)));
    if (bytes) B200_CUDA(cudaMemcpy(e->d_blob, blob.data(), bytes, cudaMemcpyHostToDevice));
    d.blob = e->d_blob;
    {
        const int grc = build_grids(e, &d);
        if (grc) return grc;
    }
    d.n_hf = static_cast<int>(e->hfs.size());
    for (int i = 0; i < d.n_hf; ++i) d.hf[i] = e->hfs[i];
    d.n_pc = static_cast<int>(e->pcs.size());
    for (int i = 0; i < d.n_pc; ++i) d.pc[i] = e->pcs[i].view;
    d.att = e->d_att;
    d.n_att = e->att.empty() ? 0 : static_cast<int>((e->att.size() - 16) / 4);
    e->desc = d;
    e->committed = true;
    return B200MP_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------- launches
namespace {

int check_common(const char *who, b200mp_ctx *c, const b200mp_env *e, const void *a, const void *out, int64_t m) {
    if (!c || !e) return fail(B200MP_ERR_INVALID, "%s: NULL handle", who);
    if (e->ctx != c) return fail(B200MP_ERR_INVALID, "%s: environment belongs to another context", who);
    if (!e->committed) return fail(B200MP_ERR_STATE, "%s: environment not committed (call b200mp_env_commit)", who);
    if (m < 0 || (m > 0 && (!a || !out))) return fail(B200MP_ERR_INVALID, "%s: bad buffer or count", who);
    return B200MP_OK;
}

// a free small-call lane, or nullptr when all are taken (the caller then goes through the context mutex like a large call)
b200mp_small_lane *acquire_lane(b200mp_ctx *c) {
    for (auto &ln : c->lanes) {
        bool expected = false;
        if (ln.busy.compare_exchange_strong(expected, true, std::memory_order_acquire)) return &ln;
    }
    return nullptr;
}
struct LaneGuard {
    b200mp_small_lane *ln;
    ~LaneGuard() { if (ln) ln->busy.store(false, std::memory_order_release); }
};

// host float -> device float (direct) or host double -> pinned float staging -> device
template <typename T>
int upload_states(b200mp_ctx *c, const T *h, int64_t count, float *d, cudaStream_t st) {
    if (std::is_same<T, float>::value) {
        B200_CUDA(cudaMemcpyAsync(d, h, count * sizeof(float), cudaMemcpyHostToDevice, st));
        return B200MP_OK;
    }
    std::vector<float> stage(static_cast<size_t>(std::max<int64_t>(count, 1)));
    for (int64_t i = 0; i < count; ++i) stage[i] = static_cast<float>(h[i]);  // ompl_to_vamp: static_cast<float>(values[i])
    B200_CUDA(cudaMemcpyAsync(d, stage.data(), count * sizeof(float), cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaStreamSynchronize(st));  // the staging vector dies with this scope
    return B200MP_OK;
}

template <typename T>
int validate_motions_host(b200mp_ctx *c, int robot, const b200mp_env *e, const T *q0, const T *q1, int64_t m, int resolution,
                          uint8_t *valid) {
    int rc = check_common("b200mp_validate_motions", c, e, q0, valid, m);
    if (rc) return rc;
    if (m > 0 && !q1) return fail(B200MP_ERR_INVALID, "b200mp_validate_motions: q1 is NULL");
    int dim = 0, res = 0;
    rc = b200mp_robot_info(robot, &dim, nullptr, &res, nullptr, nullptr);
    if (rc) return rc;
    if (m == 0) return B200MP_OK;
    const size_t qbytes = static_cast<size_t>(m) * dim * sizeof(float);
    if (2 * qbytes + static_cast<size_t>(m) + 64 <= b200mp_ctx::kLaneBytes) {
        // a planner's single checkMotion (or a handful): own lane, no context lock; inputs cast straight into the mapped buffer, result
        // read back from it -- one launch and one stream synchronise per call
        LaneGuard g{acquire_lane(c)};
        if (g.ln) {
            B200_CUDA(cudaSetDevice(c->device));
            float *h0 = static_cast<float *>(g.ln->h), *h1 = h0 + static_cast<size_t>(m) * dim;
            for (int64_t i = 0; i < m * dim; ++i) { h0[i] = static_cast<float>(q0[i]); h1[i] = static_cast<float>(q1[i]); }
            const float *d0 = static_cast<const float *>(g.ln->d);
            uint8_t *dv = reinterpret_cast<uint8_t *>(g.ln->d) + 2 * qbytes;
            rc = b200mp_validate_motions_device(c, robot, e, d0, d0 + static_cast<size_t>(m) * dim, m, resolution, dv, g.ln->stream);
            if (rc) return rc;
            B200_CUDA(cudaStreamSynchronize(g.ln->stream));
            std::memcpy(valid, static_cast<const uint8_t *>(g.ln->h) + 2 * qbytes, static_cast<size_t>(m));
            return B200MP_OK;
        }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA(cudaSetDevice(c->device));
    if (2 * qbytes + static_cast<size_t>(m) + 64 <= c->small_cap) {
        // few motions (a planner's single checkMotion): inputs converted straight into the mapped buffer, result read from it
        float *h0 = static_cast<float *>(c->h_small), *h1 = h0 + static_cast<size_t>(m) * dim;
        for (int64_t i = 0; i < m * dim; ++i) { h0[i] = static_cast<float>(q0[i]); h1[i] = static_cast<float>(q1[i]); }
        const float *d0 = static_cast<const float *>(c->d_small);
        uint8_t *dv = reinterpret_cast<uint8_t *>(c->d_small) + 2 * qbytes;
        rc = b200mp_validate_motions_device(c, robot, e, d0, d0 + static_cast<size_t>(m) * dim, m, resolution, dv, c->stream);
        if (rc) return rc;
        B200_CUDA(cudaStreamSynchronize(c->stream));
        std::memcpy(valid, static_cast<const uint8_t *>(c->h_small) + 2 * qbytes, static_cast<size_t>(m));
        return B200MP_OK;
    }
    rc = c->scratch.reserve(2 * qbytes + static_cast<size_t>(m) + 256);
    if (rc) return rc;
    float *d_q0 = reinterpret_cast<float *>(c->scratch.ptr);
    float *d_q1 = reinterpret_cast<float *>(c->scratch.ptr + qbytes);
    uint8_t *d_valid = reinterpret_cast<uint8_t *>(c->scratch.ptr + 2 * qbytes);
    // Large batches go through in chunks: the upload of chunk i + 1 and the download of chunk i - 1 overlap the kernel of chunk i (three
    // streams, one event pair per chunk). fp32 states are copied straight from the caller's buffers (the overlap needs them pinned,
    // pageable ones just serialise); fp64 states (what OMPL stores) are cast chunk by chunk into a two-slot pinned staging ring, so the
    // cast of chunk i + 1 on the host also overlaps the copies and the kernel of chunk i.
    constexpr bool kIsFloat = std::is_same<T, float>::value;
    constexpr int64_t kChunk = kIsFloat ? (int64_t(1) << 18) : (int64_t(1) << 16);
    if (m >= 2 * kChunk) {
        if (!c->h2d_stream) B200_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        if (!c->d2h_stream) B200_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        // chunks alternate between two kernel streams: the kernels are persistent (one block per SM), so a chunk's kernels end in a tail
        // of SMs going idle one by one -- the next chunk's kernels, on the other stream, move in as they do
        static const bool two_streams = !(getenv("B200MP_CHUNK_STREAMS") && atoi(getenv("B200MP_CHUNK_STREAMS")) == 1);
        if (two_streams && !c->stream2) B200_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        const size_t slot_floats = static_cast<size_t>(kChunk) * dim;  // one array of one slot
        if (!kIsFloat && c->stage_cap < 4 * slot_floats * sizeof(float)) {
            if (c->h_stage) cudaFreeHost(c->h_stage);
            c->h_stage = nullptr;
            c->stage_cap = 0;
            B200_CUDA(cudaHostAlloc(&c->h_stage, 4 * slot_floats * sizeof(float), cudaHostAllocDefault));
            c->stage_cap = 4 * slot_floats * sizeof(float);
        }
        // chunk boundaries: the first chunk's upload is the only one nothing overlaps, every chunk costs a kernel sequence with its own
        // tail, so few chunks win (measured on 2^20 C2 motions: 2^18 then doubling 4.87 ms, 2 x 2^19 4.86, 2^16 then doubling 5.34, one chunk 5.31)
        std::vector<int64_t> bounds{0};
        constexpr int64_t kMaxChunk = kIsFloat ? 2 * kChunk : kChunk;  // (fp64: a chunk must fit a slot of the staging ring)
        const char *e0 = getenv("B200MP_CHUNK0"), *e1 = getenv("B200MP_CHUNKMAX");  // (development switches, fp32 only)
        const int64_t first_sz = e0 && kIsFloat ? atoll(e0) : kChunk, max_sz = e1 && kIsFloat ? atoll(e1) : kMaxChunk;
        for (int64_t sz = first_sz; bounds.back() < m; sz = std::min<int64_t>(2 * sz, max_sz)) bounds.push_back(std::min<int64_t>(m, bounds.back() + sz));
        const int n_chunks = static_cast<int>(bounds.size()) - 1;
        std::vector<cudaEvent_t> up(n_chunks), done(n_chunks);
        for (int i = 0; i < n_chunks; ++i) {
            B200_CUDA(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        int err = B200MP_OK;
        for (int i = 0; i < n_chunks && !err; ++i) {
            const int64_t first = bounds[i], cnt = bounds[i + 1] - bounds[i];
            const size_t off = static_cast<size_t>(first) * dim, bytes = static_cast<size_t>(cnt) * dim * sizeof(float);
            const float *s0 = reinterpret_cast<const float *>(q0) + off, *s1 = reinterpret_cast<const float *>(q1) + off;
            cudaError_t ce = cudaSuccess;
            if (!kIsFloat) {
                float *h0 = static_cast<float *>(c->h_stage) + static_cast<size_t>(i & 1) * 2 * slot_floats, *h1 = h0 + slot_floats;
                if (i >= 2) ce = cudaEventSynchronize(up[i - 2]);  // the slot's previous contents have been uploaded
                // ompl_to_vamp: static_cast<float>(values[i]); four host threads share the cast of a chunk (it would otherwise be the
                // slowest stage of the pipeline)
                const size_t total = static_cast<size_t>(cnt) * dim;
                auto cast = [&](size_t b, size_t e2) {
                    for (size_t t = b; t < e2; ++t) {
                        h0[t] = static_cast<float>(q0[off + t]);
                        h1[t] = static_cast<float>(q1[off + t]);
                    }
                };
                constexpr size_t kCastThreads = 4;
                std::thread helpers[kCastThreads - 1];
                for (size_t w = 1; w < kCastThreads; ++w) helpers[w - 1] = std::thread(cast, total * w / kCastThreads, total * (w + 1) / kCastThreads);
                cast(0, total / kCastThreads);
                for (auto &th : helpers) th.join();
                s0 = h0;
                s1 = h1;
            }
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_q0 + off, s0, bytes, cudaMemcpyHostToDevice, c->h2d_stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_q1 + off, s1, bytes, cudaMemcpyHostToDevice, c->h2d_stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(up[i], c->h2d_stream);
            cudaStream_t ks = (two_streams && (i & 1)) ? c->stream2 : c->stream;
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ks, up[i], 0);
            if (ce != cudaSuccess) { err = fail(B200MP_ERR_CUDA, "b200mp_validate_motions: upload failed: %s", cudaGetErrorString(ce)); break; }
            err = b200mp_validate_motions_device(c, robot, e, d_q0 + off, d_q1 + off, cnt, resolution, d_valid + first, ks);
            if (err) break;
            ce = cudaEventRecord(done[i], ks);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(c->d2h_stream, done[i], 0);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(valid + first, d_valid + first, static_cast<size_t>(cnt), cudaMemcpyDeviceToHost, c->d2h_stream);
            if (ce != cudaSuccess) err = fail(B200MP_ERR_CUDA, "b200mp_validate_motions: download failed: %s", cudaGetErrorString(ce));
        }
        cudaStreamSynchronize(c->h2d_stream);
        cudaStreamSynchronize(c->stream);
        if (c->stream2) cudaStreamSynchronize(c->stream2);
        cudaError_t ce = cudaStreamSynchronize(c->d2h_stream);
        for (int i = 0; i < n_chunks; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(done[i]); }
        if (!err && ce != cudaSuccess) err = fail(B200MP_ERR_CUDA, "b200mp_validate_motions: %s", cudaGetErrorString(ce));
        return err;
    }
    if ((rc = upload_states(c, q0, m * dim, d_q0, c->stream))) return rc;
    if ((rc = upload_states(c, q1, m * dim, d_q1, c->stream))) return rc;
    rc = b200mp_validate_motions_device(c, robot, e, d_q0, d_q1, m, resolution, d_valid, c->stream);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(valid, d_valid, static_cast<size_t>(m), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return B200MP_OK;
}

template <typename T>
int validate_states_host(b200mp_ctx *c, int robot, const b200mp_env *e, const T *q, int64_t m, uint8_t *valid) {
    int rc = check_common("b200mp_validate_states", c, e, q, valid, m);
    if (rc) return rc;
    int dim = 0;
    rc = b200mp_robot_info(robot, &dim, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (m == 0) return B200MP_OK;
    const size_t qbytes = static_cast<size_t>(m) * dim * sizeof(float);
    if (qbytes + static_cast<size_t>(m) + 64 <= b200mp_ctx::kLaneBytes) {
        LaneGuard g{acquire_lane(c)};  // a planner's single isValid: own lane, no context lock
        if (g.ln) {
            B200_CUDA(cudaSetDevice(c->device));
            float *h0 = static_cast<float *>(g.ln->h);
            for (int64_t i = 0; i < m * dim; ++i) h0[i] = static_cast<float>(q[i]);
            uint8_t *dv = reinterpret_cast<uint8_t *>(g.ln->d) + qbytes;
            rc = b200mp_validate_states_device(c, robot, e, static_cast<const float *>(g.ln->d), m, dv, g.ln->stream);
            if (rc) return rc;
            B200_CUDA(cudaStreamSynchronize(g.ln->stream));
            std::memcpy(valid, static_cast<const uint8_t *>(g.ln->h) + qbytes, static_cast<size_t>(m));
            return B200MP_OK;
        }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA(cudaSetDevice(c->device));
    if (qbytes + static_cast<size_t>(m) + 64 <= c->small_cap) {
        float *h0 = static_cast<float *>(c->h_small);
        for (int64_t i = 0; i < m * dim; ++i) h0[i] = static_cast<float>(q[i]);
        uint8_t *dv = reinterpret_cast<uint8_t *>(c->d_small) + qbytes;
        rc = b200mp_validate_states_device(c, robot, e, static_cast<const float *>(c->d_small), m, dv, c->stream);
        if (rc) return rc;
        B200_CUDA(cudaStreamSynchronize(c->stream));
        std::memcpy(valid, static_cast<const uint8_t *>(c->h_small) + qbytes, static_cast<size_t>(m));
        return B200MP_OK;
    }
    rc = c->scratch.reserve(qbytes + static_cast<size_t>(m) + 256);
    if (rc) return rc;
    float *d_q = reinterpret_cast<float *>(c->scratch.ptr);
    uint8_t *d_valid = reinterpret_cast<uint8_t *>(c->scratch.ptr + qbytes);
    if ((rc = upload_states(c, q, m * dim, d_q, c->stream))) return rc;
    rc = b200mp_validate_states_device(c, robot, e, d_q, m, d_valid, c->stream);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(valid, d_valid, static_cast<size_t>(m), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return B200MP_OK;
}

// dependent FFMA chains: 8 independent accumulators per thread
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

extern "C" {

int b200mp_validate_motions_device(b200mp_ctx *c, int robot, const b200mp_env *e, const float *d_q0, const float *d_q1, int64_t m,
                                   int resolution, uint8_t *d_valid, void *stream) {
    int rc = check_common("b200mp_validate_motions_device", c, e, d_q0, d_valid, m);
    if (rc) return rc;
    if (m == 0) return B200MP_OK;
    if (!d_q1) return fail(B200MP_ERR_INVALID, "b200mp_validate_motions_device: d_q1 is NULL");
    if (resolution < 0) return fail(B200MP_ERR_INVALID, "b200mp_validate_motions_device: negative resolution");
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    // batches are bucketed by length on the device: short motions share warp passes, long ones get a first look (validate_launch.cuh)
    static const bool bucketed_ok = !(getenv("B200MP_MOTIONS_BUCKETED") && atoi(getenv("B200MP_MOTIONS_BUCKETED")) == 0);
    const bool bucketed = bucketed_ok && m >= 4096;
#define B200_MOTIONS(R)                                                                                                        \
    (bucketed ? launch_motions_bucketed<R>(c, e->desc, d_q0, d_q1, nullptr, m, resolution ? resolution : R::kResolution, d_valid, st) \
              : launch_motions<R>(c, e->desc, d_q0, d_q1, nullptr, m, resolution ? resolution : R::kResolution, d_valid, st))
    switch (robot) {
        case B200MP_ROBOT_PANDA: return B200_MOTIONS(Panda);
        case B200MP_ROBOT_FETCH: return B200_MOTIONS(Fetch);
        case B200MP_ROBOT_UR5: return B200_MOTIONS(Ur5);
        case B200MP_ROBOT_BAXTER: return B200_MOTIONS(Baxter);
        default: return fail(B200MP_ERR_INVALID, "unknown robot %d", robot);
    }
#undef B200_MOTIONS
}

}  // extern "C"

// ---- sessions: the resident single-call kernel (session.cuh) ---------------------------------------------------------------------
struct b200mp_session {
    b200mp_ctx *ctx = nullptr;
    const b200mp_env *env = nullptr;
    int robot = 0, dim = 0, resolution = 0;
    cudaStream_t stream = nullptr;
    b200mp::SessionMailbox *h = nullptr, *d = nullptr;  // the same mapped memory as the host / the device sees it
    unsigned seq = 0;
    unsigned long long idle_ns = 50ull * 1000 * 1000;
    long long relaunches = 0;
    bool counted = false;  // registered in env->open_sessions
    std::mutex mu;
};

namespace {
int session_launch(b200mp_session *s) {
    volatile b200mp::SessionMailbox *mb = s->h;
    mb->alive = 1u;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    const SceneDesc &desc = s->env->desc;
    int rc = B200MP_ERR_INVALID;
    switch (s->robot) {
        case B200MP_ROBOT_PANDA: rc = launch_session<Panda>(s->ctx, desc, s->d, s->seq, s->idle_ns, s->stream); break;
        case B200MP_ROBOT_FETCH: rc = launch_session<Fetch>(s->ctx, desc, s->d, s->seq, s->idle_ns, s->stream); break;
        case B200MP_ROBOT_UR5: rc = launch_session<Ur5>(s->ctx, desc, s->d, s->seq, s->idle_ns, s->stream); break;
        case B200MP_ROBOT_BAXTER: rc = launch_session<Baxter>(s->ctx, desc, s->d, s->seq, s->idle_ns, s->stream); break;
        default: return fail(B200MP_ERR_INVALID, "b200mp_session: unknown robot %d", s->robot);
    }
    if (rc == B200MP_OK) ++s->relaunches;
    return rc;
}

template <typename T>
int session_check(b200mp_session *s, const T *q0, const T *q1, int resolution, uint8_t *valid) {
    if (!s || !q0 || !q1 || !valid) return fail(B200MP_ERR_INVALID, "b200mp_session_check_motion: NULL argument");
    if (resolution < 0) return fail(B200MP_ERR_INVALID, "b200mp_session_check_motion: negative resolution");
    std::lock_guard<std::mutex> lk(s->mu);
    volatile b200mp::SessionMailbox *mb = s->h;
    if (mb->alive == 0u) {  // the kernel left after its idle time (or was never started): start it again
        B200_CUDA(cudaSetDevice(s->ctx->device));
        const int rc = session_launch(s);
        if (rc) return rc;
    }
    for (int j = 0; j < s->dim; ++j) {
        mb->q0[j] = static_cast<float>(q0[j]);  // ompl_to_vamp: static_cast<float>(values[i])
        mb->q1[j] = static_cast<float>(q1[j]);
    }
    mb->resolution = resolution == 0 ? s->resolution : resolution;
    mb->cmd = b200mp::kSessionMotion;
    const unsigned want = ++s->seq;
    std::atomic_thread_fence(std::memory_order_release);
    mb->seq_req = want;
    // spin on the answer; a kernel that left between our look at `alive` and the request is restarted once
    const auto t0 = std::chrono::steady_clock::now();
    bool restarted = false;
    for (unsigned spins = 0;; ++spins) {
        if (mb->seq_ack == want) break;
        if ((spins & 1023u) == 1023u) {
            if (mb->alive == 0u && !restarted && mb->seq_ack != want) {
                B200_CUDA(cudaStreamSynchronize(s->stream));
                if (mb->seq_ack == want) break;
                --s->seq;  // the relaunched kernel starts from the sequence number before this request and sees it as new
                const int rc = session_launch(s);
                ++s->seq;
                if (rc) return rc;
                restarted = true;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) {
                const cudaError_t ce = cudaStreamQuery(s->stream);
                return fail(B200MP_ERR_CUDA, "b200mp_session_check_motion: no answer from the resident kernel (%s)",
                            ce == cudaSuccess || ce == cudaErrorNotReady ? "timeout" : cudaGetErrorString(ce));
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *valid = static_cast<uint8_t>(mb->result);
    return B200MP_OK;
}
}  // namespace

extern "C" {

int b200mp_session_open(b200mp_ctx *c, int robot, const b200mp_env *e, b200mp_session **out) {
    if (!out) return fail(B200MP_ERR_INVALID, "b200mp_session_open: out is NULL");
    *out = nullptr;
    float dummy = 0;
    uint8_t dummy_out = 0;
    int rc = check_common("b200mp_session_open", c, e, &dummy, &dummy_out, 1);
    if (rc) return rc;
    int dim = 0, res = 0;
    if ((rc = b200mp_robot_info(robot, &dim, nullptr, &res, nullptr, nullptr))) return rc;
    B200_CUDA(cudaSetDevice(c->device));
    auto *s = new b200mp_session();
    s->ctx = c;
    s->env = e;
    s->robot = robot;
    s->dim = dim;
    s->resolution = res;
    if (const char *v = getenv("B200MP_SESSION_IDLE_MS")) s->idle_ns = static_cast<unsigned long long>(std::max(1, atoi(v))) * 1000ull * 1000ull;
    cudaError_t ce = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaHostAlloc(reinterpret_cast<void **>(&s->h), sizeof(b200mp::SessionMailbox), cudaHostAllocMapped);
    if (ce == cudaSuccess) ce = cudaHostGetDevicePointer(reinterpret_cast<void **>(&s->d), s->h, 0);
    if (ce != cudaSuccess) {
        b200mp_session_close(s);
        return fail(B200MP_ERR_CUDA, "b200mp_session_open: %s", cudaGetErrorString(ce));
    }
    std::memset(s->h, 0, sizeof(b200mp::SessionMailbox));
    if ((rc = session_launch(s))) { b200mp_session_close(s); return rc; }
    e->open_sessions.fetch_add(1);
    s->counted = true;
    *out = s;
    return B200MP_OK;
}

void b200mp_session_close(b200mp_session *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->h && s->stream) {
        volatile b200mp::SessionMailbox *mb = s->h;
        if (mb->alive != 0u) {  // ask the kernel to leave (it would also do so by itself after its idle time)
            mb->cmd = b200mp::kSessionQuit;
            std::atomic_thread_fence(std::memory_order_release);
            mb->seq_req = ++s->seq;
        }
        cudaStreamSynchronize(s->stream);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->h) cudaFreeHost(s->h);
    if (s->counted) s->env->open_sessions.fetch_sub(1);
    cudaGetLastError();
    delete s;
}

int b200mp_session_check_motion(b200mp_session *s, const float *q0, const float *q1, int resolution, uint8_t *valid) {
    return session_check<float>(s, q0, q1, resolution, valid);
}
int b200mp_session_check_motion_f64(b200mp_session *s, const double *q0, const double *q1, int resolution, uint8_t *valid) {
    return session_check<double>(s, q0, q1, resolution, valid);
}
/* VampStateValidityChecker::isValid = validate_motion<Robot, rake, 1>(q, q) */
int b200mp_session_is_valid(b200mp_session *s, const float *q, uint8_t *valid) { return session_check<float>(s, q, q, 1, valid); }
int b200mp_session_is_valid_f64(b200mp_session *s, const double *q, uint8_t *valid) { return session_check<double>(s, q, q, 1, valid); }
int b200mp_session_stats(const b200mp_session *s, int64_t *calls, int64_t *launches) {
    if (!s) return fail(B200MP_ERR_INVALID, "b200mp_session_stats: session is NULL");
    if (calls) *calls = s->seq;
    if (launches) *launches = s->relaunches;
    return B200MP_OK;
}

int b200mp_validate_edges_device(b200mp_ctx *c, int robot, const b200mp_env *e, const float *d_states, const int32_t *d_edges, int64_t m,
                                 int resolution, uint8_t *d_valid, void *stream) {
    int rc = check_common("b200mp_validate_edges_device", c, e, d_states, d_valid, m);
    if (rc) return rc;
    if (m == 0) return B200MP_OK;
    if (!d_edges) return fail(B200MP_ERR_INVALID, "b200mp_validate_edges_device: d_edges is NULL");
    if (resolution < 0) return fail(B200MP_ERR_INVALID, "b200mp_validate_edges_device: negative resolution");
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    const int2 *pairs = reinterpret_cast<const int2 *>(d_edges);
    // roadmap edges are a few rakes long: batches go through the bucketed path (rakes of different edges share a warp pass)
    static const bool packed_ok = !(getenv("B200MP_EDGES_PACKED") && atoi(getenv("B200MP_EDGES_PACKED")) == 0);
    const bool packed = packed_ok && m >= 1024;
#define B200_EDGES(R)                                                                                                               \
    (packed ? launch_motions_bucketed<R>(c, e->desc, d_states, d_states, pairs, m, resolution ? resolution : R::kResolution, d_valid, st) \
            : launch_motions<R>(c, e->desc, d_states, d_states, pairs, m, resolution ? resolution : R::kResolution, d_valid, st))
    switch (robot) {
        case B200MP_ROBOT_PANDA: return B200_EDGES(Panda);
        case B200MP_ROBOT_FETCH: return B200_EDGES(Fetch);
        case B200MP_ROBOT_UR5: return B200_EDGES(Ur5);
        case B200MP_ROBOT_BAXTER: return B200_EDGES(Baxter);
        default: return fail(B200MP_ERR_INVALID, "unknown robot %d", robot);
    }
#undef B200_EDGES
}

}  // extern "C"

namespace b200mp {
// states visited in the order of d_order (may be null); results in the states' own slots. Used by the roadmap sampler.
int validate_states_ordered(b200mp_ctx *c, int robot, const b200mp_env *e, const float *d_q, int64_t m, uint8_t *d_valid, const unsigned *d_order,
                            cudaStream_t st) {
    switch (robot) {
        case B200MP_ROBOT_PANDA: return launch_states<Panda>(c, e->desc, d_q, m, d_valid, st, d_order);
        case B200MP_ROBOT_FETCH: return launch_states<Fetch>(c, e->desc, d_q, m, d_valid, st, d_order);
        case B200MP_ROBOT_UR5: return launch_states<Ur5>(c, e->desc, d_q, m, d_valid, st, d_order);
        case B200MP_ROBOT_BAXTER: return launch_states<Baxter>(c, e->desc, d_q, m, d_valid, st, d_order);
        default: return fail(B200MP_ERR_INVALID, "unknown robot %d", robot);
    }
}
}  // namespace b200mp

extern "C" {

int b200mp_validate_states_device(b200mp_ctx *c, int robot, const b200mp_env *e, const float *d_q, int64_t m, uint8_t *d_valid,
                                  void *stream) {
    int rc = check_common("b200mp_validate_states_device", c, e, d_q, d_valid, m);
    if (rc) return rc;
    if (m == 0) return B200MP_OK;
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return validate_states_ordered(c, robot, e, d_q, m, d_valid, nullptr, st);
}

int b200mp_validate_motions(b200mp_ctx *c, int robot, const b200mp_env *e, const float *q0, const float *q1, int64_t m, int resolution,
                            uint8_t *valid) {
    return validate_motions_host<float>(c, robot, e, q0, q1, m, resolution, valid);
}
int b200mp_validate_motions_f64(b200mp_ctx *c, int robot, const b200mp_env *e, const double *q0, const double *q1, int64_t m,
                                int resolution, uint8_t *valid) {
    return validate_motions_host<double>(c, robot, e, q0, q1, m, resolution, valid);
}
int b200mp_validate_states(b200mp_ctx *c, int robot, const b200mp_env *e, const float *q, int64_t m, uint8_t *valid) {
    return validate_states_host<float>(c, robot, e, q, m, valid);
}
int b200mp_validate_states_f64(b200mp_ctx *c, int robot, const b200mp_env *e, const double *q, int64_t m, uint8_t *valid) {
    return validate_states_host<double>(c, robot, e, q, m, valid);
}

int b200mp_fp32_peak(b200mp_ctx *c, int iters, double *tflops) {
    if (!c || !tflops || iters < 1) return fail(B200MP_ERR_INVALID, "b200mp_fp32_peak: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256;
    int rc = c->scratch.reserve(static_cast<size_t>(blocks) * threads * 4);
    if (rc) return rc;
    cudaEvent_t a, b;
    B200_CUDA(cudaEventCreate(&a));
    B200_CUDA(cudaEventCreate(&b));
    fp32_peak_kernel<<<blocks, threads, 0, c->stream>>>(reinterpret_cast<float *>(c->scratch.ptr), 16, 0.999f, 1e-3f);  // warm up
    B200_CUDA(cudaEventRecord(a, c->stream));
    fp32_peak_kernel<<<blocks, threads, 0, c->stream>>>(reinterpret_cast<float *>(c->scratch.ptr), iters, 0.999f, 1e-3f);
    B200_CUDA(cudaEventRecord(b, c->stream));
    B200_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    B200_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    c->launches.fetch_add(2);
    const double flops = 2.0 * 8 * 16 * static_cast<double>(iters) * blocks * threads;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return B200MP_OK;
}

}  // extern "C"
