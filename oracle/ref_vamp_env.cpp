// TEST INFRASTRUCTURE ONLY -- see ref_vamp_impl.hh. Environment marshalling for the compiled reference.
#include "ref_vamp_impl.hh"

using namespace refv;
namespace vc = vamp::collision;

extern "C"
{
    void *refv_env_create() { return new Env(); }
    void refv_env_destroy(void *e) { delete static_cast<Env *>(e); }

    void refv_env_add_sphere(void *e, float x, float y, float z, float r)
    {
        static_cast<Env *>(e)->f.spheres.emplace_back(vc::Sphere<float>(x, y, z, r));
    }

    // c = centre(3), axis1(3), axis2(3), axis3(3), half extents(3)  -- shapes.hh:31-127
    void refv_env_add_cuboid(void *e, const float *c, int z_aligned)
    {
        vc::Cuboid<float> s(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12], c[13], c[14]);
        auto &env = static_cast<Env *>(e)->f;
        (z_aligned ? env.z_aligned_cuboids : env.cuboids).emplace_back(s);
    }

    // c = p1(3), v = p2 - p1 (3), r; rdv = float(1.0 / dot(v, v)) as factory.hh:110-126 does
    void refv_env_add_capsule(void *e, const float *c, int z_aligned)
    {
        const float dot = vc::dot_3(c[3], c[4], c[5], c[3], c[4], c[5]);
        vc::Capsule<float> s(c[0], c[1], c[2], c[3], c[4], c[5], c[6], static_cast<float>(1.0 / dot));
        auto &env = static_cast<Env *>(e)->f;
        (z_aligned ? env.z_aligned_capsules : env.capsules).emplace_back(s);
    }

    void refv_env_add_pointcloud(void *e, const float *xyz, std::int64_t n, float r_min, float r_max, float r_point)
    {
        std::vector<vc::Point> pts(static_cast<std::size_t>(n));
        for (std::int64_t i = 0; i < n; ++i)
        {
            pts[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        }
        static_cast<Env *>(e)->f.pointclouds.emplace_back(pts, r_min, r_max, r_point);
    }

    void refv_env_add_heightfield(void *e, float cx, float cy, float cz, float sx, float sy, float sz, std::int64_t nx, std::int64_t ny, const float *data)
    {
        std::vector<float> d(data, data + nx * ny);
        // factory::heightfield::flat (factory.hh:363-386; that header needs real Eigen, so its one constructor call is made here):
        // the shape stores the reciprocal scales
        static_cast<Env *>(e)->f.heightfields.emplace_back(
            vc::HeightField<float>(cx, cy, cz, 1.F / sx, 1.F / sy, 1.F / sz, static_cast<std::size_t>(nx), static_cast<std::size_t>(ny), d));
    }

    // Attachment (collision/attachments.hh:13-56): tf = row-major 3x4 [R | t] of the attachment relative to the end effector,
    // xyzr[n][4] its spheres in the attachment's frame. n == 0 removes it.
    void refv_env_set_attachment(void *e, const float *tf, const float *xyzr, std::int64_t n)
    {
        auto &env = static_cast<Env *>(e)->f;
        env.attachments.reset();
        if (n <= 0)
        {
            return;
        }
        Eigen::Transform<float, 3, Eigen::Isometry> t;
        for (int r = 0; r < 3; ++r)
        {
            for (int c = 0; c < 3; ++c)
            {
                t.linear()(r, c) = tf[4 * r + c];
            }
            t.translation()[r] = tf[4 * r + 3];
        }
        vc::Attachment<float> a(t);
        for (std::int64_t i = 0; i < n; ++i)
        {
            a.spheres.emplace_back(vc::Sphere<float>(xyzr[4 * i], xyzr[4 * i + 1], xyzr[4 * i + 2], xyzr[4 * i + 3]));
        }
        env.attachments.emplace(a);
    }

    void refv_env_sort(void *e) { static_cast<Env *>(e)->f.sort(); }

    void refv_env_finalize(void *e)
    {
        auto *env = static_cast<Env *>(e);
        delete env->v;
        env->v = new EnvV(env->f);
    }

    // kind: 0 spheres, 1 capsules, 2 z-capsules, 3 cuboids, 4 z-cuboids. Returns count; fills
    // min_distance in stored order when out != nullptr (environment.hh:43-69, shapes.hh ctors).
    std::int64_t refv_env_min_distances(void *e, int kind, float *out)
    {
        auto &env = static_cast<Env *>(e)->f;
        auto dump = [&](const auto &v) -> std::int64_t
        {
            if (out != nullptr)
            {
                for (std::size_t i = 0; i < v.size(); ++i)
                {
                    out[i] = v[i].min_distance;
                }
            }
            return static_cast<std::int64_t>(v.size());
        };
        switch (kind)
        {
            case 0: return dump(env.spheres);
            case 1: return dump(env.capsules);
            case 2: return dump(env.z_aligned_capsules);
            case 3: return dump(env.cuboids);
            case 4: return dump(env.z_aligned_cuboids);
            default: return -1;
        }
    }

    // Flat export of point cloud tree `idx` so a caller can hand the reference-built CAPT to another
    // engine. sizes = {nlog2, n_tests, n_leaves, n_aff_vectors(8 points each)}.
    int refv_capt_sizes(void *e, int idx, std::int64_t *sizes)
    {
        auto &pcs = static_cast<Env *>(e)->f.pointclouds;
        if (idx < 0 || idx >= static_cast<int>(pcs.size()))
        {
            return -1;
        }
        const auto &t = pcs[idx];
        sizes[0] = t.nlog2;
        sizes[1] = static_cast<std::int64_t>(t.tests.size());
        sizes[2] = static_cast<std::int64_t>(t.aabbs.size());
        sizes[3] = static_cast<std::int64_t>(t.affordances[0].size());
        return 0;
    }

    // tests[n_tests]; aabbs[n_leaves][6] = lower(3), upper(3); aff_starts[n_leaves + 1];
    // aff[3][n_aff_vectors * 8]; top[6]; radii[3] = r_min, r_max, r_point
    int refv_capt_export(
        void *e, int idx, float *tests, float *aabbs, std::uint32_t *aff_starts, float *aff, float *top, float *radii)
    {
        auto &pcs = static_cast<Env *>(e)->f.pointclouds;
        if (idx < 0 || idx >= static_cast<int>(pcs.size()))
        {
            return -1;
        }
        const auto &t = pcs[idx];
        std::memcpy(tests, t.tests.data(), t.tests.size() * sizeof(float));
        for (std::size_t i = 0; i < t.aabbs.size(); ++i)
        {
            for (int k = 0; k < 3; ++k)
            {
                aabbs[6 * i + k] = t.aabbs[i].lower[k];
                aabbs[6 * i + 3 + k] = t.aabbs[i].upper[k];
            }
        }
        std::memcpy(aff_starts, t.aff_starts.data(), t.aff_starts.size() * sizeof(std::uint32_t));
        const std::size_t nv = t.affordances[0].size();
        for (int k = 0; k < 3; ++k)
        {
            for (std::size_t i = 0; i < nv; ++i)
            {
                t.affordances[k][i].to_array_unaligned(aff + (k * nv + i) * 8);
            }
        }
        for (int k = 0; k < 3; ++k)
        {
            top[k] = t.aabb_top.lower[k];
            top[3 + k] = t.aabb_top.upper[k];
        }
        radii[0] = t.r_min;
        radii[1] = t.r_max;
        radii[2] = t.r_point;
        return 0;
    }

    // Scalar CAPT query straight through the reference (capt.hh:366-407).
    int refv_capt_collides(void *e, int idx, const float *xyzr, std::int64_t n, std::uint8_t *out)
    {
        auto &pcs = static_cast<Env *>(e)->f.pointclouds;
        if (idx < 0 || idx >= static_cast<int>(pcs.size()))
        {
            return -1;
        }
        for (std::int64_t i = 0; i < n; ++i)
        {
            out[i] = pcs[idx].collides({xyzr[4 * i], xyzr[4 * i + 1], xyzr[4 * i + 2]}, xyzr[4 * i + 3]) ? 1 : 0;
        }
        return 0;
    }

    int refv_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }
}

// ---- point-cloud filter (collision/filter.hh:176-274), unmodified; returns the number of points kept, writes at most cap
#include <vamp/collision/filter.hh>
// 1: equal curve codes keep their previous order (std::stable_sort in place of the absent pdqsort); 0: std::sort (default)
extern "C" void refv_set_stable_sort(int on) { refv_sort_mode() = on ? 1 : 0; }
extern "C" std::int64_t refv_filter_pointcloud(const float *xyz, std::int64_t n, float min_dist, float max_range, const float *origin,
                                               const float *ws_min, const float *ws_max, int cull, float *out, std::int64_t cap)
{
    std::vector<vc::Point> pts(static_cast<std::size_t>(n));
    for (std::int64_t i = 0; i < n; ++i)
    {
        pts[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    }
    const auto kept = vc::filter_pointcloud(pts, min_dist, max_range, vc::Point{origin[0], origin[1], origin[2]},
                                            vc::Point{ws_min[0], ws_min[1], ws_min[2]}, vc::Point{ws_max[0], ws_max[1], ws_max[2]}, cull != 0);
    for (std::size_t i = 0; i < kept.size() && static_cast<std::int64_t>(i) < cap; ++i)
    {
        out[3 * i] = kept[i][0];
        out[3 * i + 1] = kept[i][1];
        out[3 * i + 2] = kept[i][2];
    }
    return static_cast<std::int64_t>(kept.size());
}
