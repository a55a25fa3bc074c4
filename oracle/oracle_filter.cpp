// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's point-cloud filter
// (external/vamp/src/impl/vamp/collision/filter.hh:176-274, VAMP pin aa4f4b53), scalar C++, one function.
// Pinned on the compiled reference (oracle/_ref, refv_filter_pointcloud): identical output whenever no two surviving
// entries share a Morton code in a pass (tests/test_filter_cpu.py); with ties the reference's order is whatever its
// unstable sort (pdqsort_branchless, un-vendored) leaves, here ties keep their previous relative order (stable).
//
// Behaviour restated, including the parts that look accidental:
//  * `min` AND `max` of the curve range are both std::min over the three coordinates (filter.hh:195-196);
//  * the work list is resized to n BEFORE culling and never shrunk (filter.hh:198-219): the n - kept trailing entries
//    all refer to point 0, take part in the first pass, and can put point 0 into the output even when it was culled;
//  * remap_point converts a float that may be negative or > 1000 to uint32 (filter.hh:144-147): on x86-64 that is
//    cvttss2si to int64 followed by truncation to 32 bits, and pdep keeps the low 11 / 11 / 10 bits per axis;
//  * six passes, one per permutation of the axes in std::next_permutation order, the range shrinking between passes
//    (filter.hh:224-262).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

inline uint32_t remap(float x, float mn, float mx) {
    const float v = ((x - mn) / (mx - mn)) * 1000.0f;
    return static_cast<uint32_t>(static_cast<int64_t>(v));  // x86-64 float -> uint32 conversion
}

inline uint32_t spread(uint32_t v, uint32_t mask) {  // _pdep_u32
    uint32_t out = 0;
    for (uint32_t bit = 0, m = mask; m; ++bit) {
        const uint32_t low = m & (~m + 1u);
        if (v & (1u << bit)) out |= low;
        m ^= low;
    }
    return out;
}

inline float sq3(const float *a, const float *b) {
    const float xs = a[0] - b[0], ys = a[1] - b[1], zs = a[2] - b[2];
    return std::fmaf(zs, zs, std::fmaf(ys, ys, xs * xs));  // (xs*xs + ys*ys) + zs*zs under -ffp-contract=fast
}

}  // namespace

extern "C" int64_t orf_filter_pointcloud(const float *xyz, int64_t n, float min_dist, float max_range, const float *origin,
                                         const float *ws_min, const float *ws_max, int cull, float *out, int64_t cap) {
    if (n == 0) return 0;
    const float sqdist = min_dist * min_dist, sqrange = max_range * max_range;
    float mn = std::min({origin[0] - max_range, origin[1] - max_range, origin[2] - max_range});
    float mx = std::min({origin[0] + max_range, origin[1] + max_range, origin[2] + max_range});
    std::vector<uint32_t> idx(static_cast<size_t>(n), 0u);
    size_t hi = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = xyz + 3 * i;
        const bool keep = !cull || (sq3(p, origin) < sqrange && ws_min[0] <= p[0] && p[0] <= ws_max[0] && ws_min[1] <= p[1] &&
                                    p[1] <= ws_max[1] && ws_min[2] <= p[2] && p[2] <= ws_max[2]);
        if (keep) idx[hi++] = static_cast<uint32_t>(i);
    }
    int perm[3] = {0, 1, 2};
    std::vector<uint32_t> code(idx.size());
    std::vector<uint32_t> order, kept;
    do {
        float new_min = mx, new_max = mn;
        code.resize(idx.size());
        for (size_t e = 0; e < idx.size(); ++e) {
            const float *p = xyz + 3 * static_cast<size_t>(idx[e]);
            const uint32_t c0 = remap(p[perm[0]], mn, mx), c1 = remap(p[perm[1]], mn, mx), c2 = remap(p[perm[2]], mn, mx);
            new_min = std::min({new_min, p[0], p[1], p[2]});
            new_max = std::max({new_max, p[0], p[1], p[2]});
            code[e] = spread(c0, 0x49249249u) | spread(c1, 0x92492492u) | spread(c2, 0x24924924u);
        }
        order.resize(idx.size());
        for (size_t e = 0; e < idx.size(); ++e) order[e] = static_cast<uint32_t>(e);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
        kept.clear();
        kept.push_back(idx[order[0]]);
        for (size_t i = 1; i < order.size(); ++i) {
            const uint32_t h1 = idx[order[i]], h2 = kept.back();
            if (sq3(xyz + 3 * static_cast<size_t>(h1), xyz + 3 * static_cast<size_t>(h2)) > sqdist) kept.push_back(h1);
        }
        idx = kept;
        mx = static_cast<float>((new_max + mx) / 2.);
        mn = static_cast<float>((new_min + mn) / 2.);
    } while (std::next_permutation(perm, perm + 3));
    for (size_t i = 0; i < idx.size() && static_cast<int64_t>(i) < cap; ++i)
        for (int k = 0; k < 3; ++k) out[3 * i + k] = xyz[3 * static_cast<size_t>(idx[i]) + k];
    return static_cast<int64_t>(idx.size());
}
