// Host-side construction of the collision-affording point tree (CAPT) in the exact array layout the reference
// queries (external/vamp/src/impl/vamp/collision/capt.hh:104-365): an implicit balanced k-d tree over the points
// padded to 2^L with +inf, split planes at the median of x, y, z in turn, and per leaf the list of points that a
// sphere of radius <= r_max centred anywhere in the leaf's cell could touch ("affordances"), stored 8 points per
// record and padded with +inf.
//
// Decisions downstream depend on WHICH points a leaf lists, so the construction rules are reproduced faithfully,
// including two reference peculiarities:
//   * candidates are carried down the tree with a per-axis slab filter of width r_max (not r_max + r_point) and
//     only the final leaf test uses the Euclidean distance to the cell with (r_max + r_point)^2;
//   * when a node is split, the upper child inherits points of the lower half only through a scan that starts at
//     the lower half's SMALLEST coordinate (capt.hh:226-232), i.e. all of the lower half or none of it, while the
//     lower child inherits the proper prefix of the upper half (capt.hh:234-240).
// No GPU is needed for the build; b200mp_env_add_pointcloud uploads the result.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <atomic>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

#include "../../include/b200mp.h"
#include "host_common.hpp"

using namespace b200mp;

struct b200mp_capt {
    int nlog2 = 0;
    float r_min = 0, r_max = 0, r_point = 0;
    std::vector<float> tests;        // 2^L - 1 split planes, NaN where never written
    std::vector<float> aabbs;        // [leaf][6] lower xyz, upper xyz of the leaf's listed points
    std::vector<uint32_t> aff_starts;  // [leaves + 1], in records of 8 points
    std::vector<float> ax, ay, az;   // affordance coordinates, 8 per record
    float top[6];
};

namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

struct Cell {
    float lo[3], hi[3];
};

inline float cell_distsq(const Cell &c, const float *p) {
    const float d0 = p[0] - std::min(std::max(p[0], c.lo[0]), c.hi[0]);
    const float d1 = p[1] - std::min(std::max(p[1], c.lo[1]), c.hi[1]);
    const float d2 = p[2] - std::min(std::max(p[2], c.lo[2]), c.hi[2]);
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// One contiguous run of leaves (a whole tree, or the subtree a worker thread builds): the same arrays as b200mp_capt, with
// aff_starts relative to the run.
struct LeafRun {
    std::vector<float> aabbs, ax, ay, az;
    std::vector<uint32_t> starts;  // per leaf: records (of 8 points) listed so far in this run, AFTER the leaf
    float top[6] = {kInf, kInf, kInf, -kInf, -kInf, -kInf};
};

struct Task {
    uint32_t begin, count, node;
    std::vector<uint32_t> cand;
    Cell cell;
    int axis;
};

struct Builder {
    const std::vector<float> &pts;  // [n2][3]
    std::vector<uint32_t> &order;   // shared; every subtree sorts its own segment
    std::vector<float> &tests;      // shared; every node is written once
    float slab, afford_sq, inner_sq;
    LeafRun *out = nullptr;         // where leaves go (one run per worker task)
    int task_depth = -1;            // >= 0: subtrees rooted at this depth are queued instead of built
    std::vector<Task> *tasks = nullptr;

    const float *p(uint32_t id) const { return pts.data() + 3 * static_cast<size_t>(id); }

    void leaf(uint32_t at, const std::vector<uint32_t> &cand, const Cell &cell) {
        LeafRun &o = *out;
        const float *rep = p(order[at]);
        float bb[6] = {rep[0], rep[1], rep[2], rep[0], rep[1], rep[2]};
        if (std::isfinite(rep[0])) {
            for (int k = 0; k < 3; ++k) {
                o.top[k] = std::min(o.top[k], rep[k]);
                o.top[3 + k] = std::max(o.top[3 + k], rep[k]);
            }
            size_t listed = 1;
            o.ax.push_back(rep[0]);
            o.ay.push_back(rep[1]);
            o.az.push_back(rep[2]);
            // a cell that lies inside the ball of radius r_min + r_point around its own point needs nothing else:
            // every admissible query centred in the cell already hits that point
            const float e0 = std::max(rep[0] - cell.lo[0], cell.hi[0] - rep[0]);
            const float e1 = std::max(rep[1] - cell.lo[1], cell.hi[1] - rep[1]);
            const float e2 = std::max(rep[2] - cell.lo[2], cell.hi[2] - rep[2]);
            if (!((e0 * e0 + e1 * e1 + e2 * e2) <= inner_sq)) {
                for (const uint32_t id : cand) {
                    const float *q = p(id);
                    if (cell_distsq(cell, q) <= afford_sq) {
                        for (int k = 0; k < 3; ++k) {
                            bb[k] = std::min(bb[k], q[k]);
                            bb[3 + k] = std::max(bb[3 + k], q[k]);
                        }
                        o.ax.push_back(q[0]);
                        o.ay.push_back(q[1]);
                        o.az.push_back(q[2]);
                        ++listed;
                    }
                }
            }
            for (; listed % 8; ++listed) {  // padding that can never satisfy distsq <= r^2
                o.ax.push_back(kInf);
                o.ay.push_back(kInf);
                o.az.push_back(kInf);
            }
        }
        o.aabbs.insert(o.aabbs.end(), bb, bb + 6);
        o.starts.push_back(static_cast<uint32_t>(o.ax.size() / 8));
    }

    void split(uint32_t begin, uint32_t count, uint32_t node, std::vector<uint32_t> cand, const Cell &cell, int axis, int depth) {
        if (count == 1) {
            leaf(begin, cand, cell);
            return;
        }
        if (depth == task_depth) {  // hand the subtree to a worker (tasks stay in left-to-right order)
            tasks->push_back(Task{begin, count, node, std::move(cand), cell, axis});
            return;
        }
        std::sort(order.begin() + begin, order.begin() + begin + count,
                  [this, axis](uint32_t a, uint32_t b) { return p(a)[axis] < p(b)[axis]; });
        const uint32_t half = count / 2, mid = begin + half, end = begin + count;
        const float plane = static_cast<float>((p(order[mid - 1])[axis] + p(order[mid])[axis]) / 2.0);
        tests[node] = plane;
        Cell lo_cell = cell, hi_cell = cell;
        lo_cell.hi[axis] = plane;
        hi_cell.lo[axis] = plane;

        std::vector<uint32_t> lo_cand, hi_cand;
        lo_cand.reserve(cand.size());
        hi_cand.reserve(cand.size());
        for (const uint32_t id : cand) {
            const float v = p(id)[axis];
            if (v <= plane + slab) lo_cand.push_back(id);
            if (v >= plane - slab) hi_cand.push_back(id);
        }
        // points of the lower half handed to the upper child: scan from the lower half's first (smallest) entry
        for (uint32_t i = begin; i < mid; ++i) {
            const float v = p(order[i])[axis];
            if (!(v >= plane - slab) || !std::isfinite(v)) break;
            hi_cand.push_back(order[i]);
        }
        // points of the upper half handed to the lower child: its prefix within the slab
        for (uint32_t i = mid; i < end; ++i) {
            const float v = p(order[i])[axis];
            if (!(v <= plane + slab) || !std::isfinite(v)) break;
            lo_cand.push_back(order[i]);
        }
        cand = std::vector<uint32_t>();  // release before descending
        const int next = (axis + 1) % 3;
        split(begin, half, 2 * node + 1, std::move(lo_cand), lo_cell, next, depth + 1);
        split(mid, half, 2 * node + 2, std::move(hi_cand), hi_cell, next, depth + 1);
    }
};

}  // namespace

extern "C" {

int b200mp_capt_build(const float *xyz, int64_t n, float r_min, float r_max, float r_point, b200mp_capt **out) {
    if (!out || n < 0 || (n > 0 && !xyz) || n > (int64_t(1) << 30))
        return fail(B200MP_ERR_INVALID, "b200mp_capt_build: bad argument");
    auto *t = new b200mp_capt();
    t->r_min = r_min;
    t->r_max = r_max;
    t->r_point = r_point;
    while ((int64_t(1) << t->nlog2) < n) ++t->nlog2;
    const size_t n2 = size_t(1) << t->nlog2;
    std::vector<float> pts(3 * n2, kInf);
    if (n > 0) std::memcpy(pts.data(), xyz, sizeof(float) * 3 * static_cast<size_t>(n));
    for (int k = 0; k < 3; ++k) {
        t->top[k] = kInf;
        t->top[3 + k] = -kInf;
    }
    t->tests.assign(n2 - 1, std::numeric_limits<float>::quiet_NaN());
    t->aabbs.reserve(6 * n2);
    t->aff_starts.reserve(n2 + 1);
    t->aff_starts.push_back(0);
    const float reach = r_max + r_point;
    std::vector<uint32_t> order(n2);
    std::iota(order.begin(), order.end(), 0u);
    Cell root;
    for (int k = 0; k < 3; ++k) {
        root.lo[k] = -kInf;
        root.hi[k] = kInf;
    }
    // The top kTaskDepth levels are split on this thread; the 2^kTaskDepth subtrees below are independent (disjoint segments of
    // `order`, disjoint nodes of `tests`, own leaf runs) and are built by worker threads, then concatenated in order -- the
    // arrays are those of the serial recursion.
    constexpr int kTaskDepth = 5;
    std::vector<Task> tasks;
    LeafRun serial;
    Builder top{pts, order, t->tests, r_max, reach * reach, (r_min + r_point) * (r_min + r_point)};
    top.out = &serial;
    if (t->nlog2 > kTaskDepth + 2 && !getenv("B200MP_CAPT_SERIAL")) {
        top.task_depth = kTaskDepth;
        top.tasks = &tasks;
    }
    top.split(0, static_cast<uint32_t>(n2), 0, {}, root, 0, 0);
    std::vector<LeafRun> runs(tasks.size());
    if (!tasks.empty()) {
        const unsigned hw = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= tasks.size()) return;
                Builder b{pts, order, t->tests, r_max, reach * reach, (r_min + r_point) * (r_min + r_point)};
                b.out = &runs[i];
                Task &k = tasks[i];
                b.split(k.begin, k.count, k.node, std::move(k.cand), k.cell, k.axis, kTaskDepth + 1);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < hw; ++w) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
    } else {
        runs.push_back(std::move(serial));
    }
    size_t n_rec = 0, n_leaf = 0;
    for (const auto &r : runs) { n_rec += r.ax.size(); n_leaf += r.starts.size(); }
    t->ax.reserve(n_rec); t->ay.reserve(n_rec); t->az.reserve(n_rec);
    t->aabbs.reserve(6 * n_leaf);
    for (const auto &r : runs) {
        const uint32_t base = static_cast<uint32_t>(t->ax.size() / 8);
        t->ax.insert(t->ax.end(), r.ax.begin(), r.ax.end());
        t->ay.insert(t->ay.end(), r.ay.begin(), r.ay.end());
        t->az.insert(t->az.end(), r.az.begin(), r.az.end());
        t->aabbs.insert(t->aabbs.end(), r.aabbs.begin(), r.aabbs.end());
        for (const uint32_t sre : r.starts) t->aff_starts.push_back(base + sre);
        for (int k = 0; k < 3; ++k) {
            t->top[k] = std::min(t->top[k], r.top[k]);
            t->top[3 + k] = std::max(t->top[3 + k], r.top[3 + k]);
        }
    }
    *out = t;
    return B200MP_OK;
}

void b200mp_capt_destroy(b200mp_capt *t) { delete t; }

int b200mp_capt_sizes(const b200mp_capt *t, int64_t *sizes) {
    if (!t || !sizes) return fail(B200MP_ERR_INVALID, "b200mp_capt_sizes: NULL argument");
    sizes[0] = t->nlog2;
    sizes[1] = static_cast<int64_t>(t->tests.size());
    sizes[2] = static_cast<int64_t>(t->aabbs.size() / 6);
    sizes[3] = static_cast<int64_t>(t->ax.size() / 8);
    return B200MP_OK;
}

int b200mp_capt_export(const b200mp_capt *t, float *tests, float *aabbs, uint32_t *aff_starts, float *aff, float *top) {
    if (!t) return fail(B200MP_ERR_INVALID, "b200mp_capt_export: NULL handle");
    const size_t na = t->ax.size();
    if (tests) std::memcpy(tests, t->tests.data(), t->tests.size() * 4);
    if (aabbs) std::memcpy(aabbs, t->aabbs.data(), t->aabbs.size() * 4);
    if (aff_starts) std::memcpy(aff_starts, t->aff_starts.data(), t->aff_starts.size() * 4);
    if (aff) {
        std::memcpy(aff, t->ax.data(), na * 4);
        std::memcpy(aff + na, t->ay.data(), na * 4);
        std::memcpy(aff + 2 * na, t->az.data(), na * 4);
    }
    if (top) std::memcpy(top, t->top, 24);
    return B200MP_OK;
}

int b200mp_env_add_pointcloud(b200mp_env *env, const float *xyz, int64_t n, float r_min, float r_max, float r_point) {
    if (!env) return fail(B200MP_ERR_INVALID, "b200mp_env_add_pointcloud: env is NULL");
    b200mp_capt *t = nullptr;
    int rc = b200mp_capt_build(xyz, n, r_min, r_max, r_point, &t);
    if (rc) return rc;
    const size_t na = t->ax.size();
    std::vector<float> aff(3 * std::max<size_t>(na, 1));
    b200mp_capt_export(t, nullptr, nullptr, nullptr, aff.data(), nullptr);
    rc = b200mp_env_add_capt(env, t->nlog2, t->tests.data(), t->aabbs.data(), t->aff_starts.data(), aff.data(),
                             static_cast<int64_t>(na / 8), t->top, r_min, r_max, r_point);
    b200mp_capt_destroy(t);
    return rc;
}

}  // extern "C"
