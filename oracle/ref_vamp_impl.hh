// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C wrapper around the UNMODIFIED reference implementation (VAMP headers under
// /root/reference/external/vamp/src/impl, pinned aa4f4b53), compiled where they lie by
// oracle/Makefile into oracle/_ref/libref_vamp.so. Nothing here restates the algorithm: it
// only marshals flat arrays into the reference's own types and calls
//   vamp::planning::validate_motion<Robot, 8, res>      (planning/validate.hh:69-77)
//   Robot::sphere_fk<8>                                 (robots/panda.hh:116-463)
//   collision::CAPT                                     (collision/capt.hh:61-618)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load the resulting library.
#pragma once

#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <vamp/vector.hh>
#include <vamp/collision/environment.hh>
#include <vamp/collision/shapes.hh>
#include <vamp/collision/validity.hh>
#include <vamp/planning/validate.hh>

namespace refv
{
    static constexpr std::size_t rake = vamp::FloatVectorWidth;
    using EnvF = vamp::collision::Environment<float>;
    using EnvV = vamp::collision::Environment<vamp::FloatVector<rake>>;

    struct Env
    {
        EnvF f;
        EnvV *v = nullptr;  // rebuilt by finalize()
        ~Env() { delete v; }
    };

    template <typename Robot>
    inline auto make_config(const float *q) -> typename Robot::Configuration
    {
        // Same construction the OMPL bridge uses, with the pad lanes defined as zero
        // (src/ompl/vamp/Utils.h:16-30; SURVEY Appendix A "ompl_to_vamp over-read").
        std::array<float, Robot::Configuration::num_scalars> a{};
        for (std::size_t i = 0; i < Robot::dimension; ++i)
        {
            a[i] = q[i];
        }
        return typename Robot::Configuration(a);
    }

    template <typename Robot, std::size_t res>
    inline void run_motions(
        const EnvV &env,
        const float *q0,
        const float *q1,
        std::int64_t begin,
        std::int64_t end,
        std::uint8_t *out)
    {
        constexpr std::size_t d = Robot::dimension;
        for (std::int64_t i = begin; i < end; ++i)
        {
            auto a = make_config<Robot>(q0 + i * d);
            auto b = make_config<Robot>(q1 + i * d);
            out[i] = vamp::planning::validate_motion<Robot, rake, res>(a, b, env) ? 1 : 0;
        }
    }

    template <typename Robot>
    inline int validate_motions(
        const Env *e,
        const float *q0,
        const float *q1,
        std::int64_t m,
        int resolution,
        std::uint8_t *out,
        int nthreads)
    {
        if (e == nullptr || e->v == nullptr)
        {
            return -1;
        }
        if (nthreads < 1)
        {
            nthreads = 1;
        }
        // fkcc_attach re-poses the attachment inside the (const) environment (collision/attachments.hh:37-41, a `mutable`
        // member): threads sharing one environment would race, so with an attachment every thread works on its own copy
        const bool own_copy = e->v->attachments.has_value() && nthreads > 1;
        auto body_on = [&](const EnvV &env, std::int64_t b, std::int64_t en)
        {
            if (resolution == 1)
            {
                run_motions<Robot, 1>(env, q0, q1, b, en, out);
            }
            else
            {
                run_motions<Robot, Robot::resolution>(env, q0, q1, b, en, out);
            }
        };
        auto body = [&](std::int64_t b, std::int64_t en) { body_on(*e->v, b, en); };
        if (resolution != 1 && resolution != 0 && resolution != static_cast<int>(Robot::resolution))
        {
            return -2;  // the reference fixes resolution at compile time
        }
        if (nthreads == 1)
        {
            body(0, m);
            return 0;
        }
        // dynamic chunks: edge cost varies by 100x with early exit
        std::atomic<std::int64_t> next{0};
        const std::int64_t chunk = 256;
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
        {
            pool.emplace_back(
                [&]()
                {
                    std::unique_ptr<EnvV> mine;
                    if (own_copy)
                    {
                        mine = std::make_unique<EnvV>(*e->v);
                    }
                    const EnvV &env = own_copy ? *mine : *e->v;
                    for (;;)
                    {
                        const std::int64_t b = next.fetch_add(chunk);
                        if (b >= m)
                        {
                            break;
                        }
                        body_on(env, b, std::min(b + chunk, m));
                    }
                });
        }
        for (auto &t : pool)
        {
            t.join();
        }
        return 0;
    }

    // out[m][n_spheres][4] = (x, y, z, r) of the FINE spheres from the reference's own traced FK.
    template <typename Robot>
    inline int sphere_fk(const float *q, std::int64_t m, float *out)
    {
        constexpr std::size_t d = Robot::dimension;
        constexpr std::size_t ns = Robot::n_spheres;
        for (std::int64_t base = 0; base < m; base += rake)
        {
            typename Robot::template ConfigurationBlock<rake> block;
            for (std::size_t j = 0; j < d; ++j)
            {
                alignas(32) std::array<float, rake> lane{};
                for (std::size_t l = 0; l < rake; ++l)
                {
                    const std::int64_t i = std::min<std::int64_t>(base + l, m - 1);
                    lane[l] = q[i * d + j];
                }
                block[j] = vamp::FloatVector<rake, 1>(lane.data());
            }
            typename Robot::template Spheres<rake> s;
            Robot::template sphere_fk<rake>(block, s);
            for (std::size_t k = 0; k < ns; ++k)
            {
                const auto xs = s.x[k].to_array();
                const auto ys = s.y[k].to_array();
                const auto zs = s.z[k].to_array();
                const auto rs = s.r[k].to_array();
                for (std::size_t l = 0; l < rake && base + static_cast<std::int64_t>(l) < m; ++l)
                {
                    float *o = out + ((base + l) * ns + k) * 4;
                    o[0] = xs[l];
                    o[1] = ys[l];
                    o[2] = zs[l];
                    o[3] = rs[l];
                }
            }
        }
        return 0;
    }
    // Robot::eefk (robots/panda.hh:15472-15552): the end-effector pose for scalar configurations
    template <typename Robot>
    inline int eefk(const float *q, std::int64_t m, float *out)
    {
        constexpr std::size_t d = Robot::dimension;
        for (std::int64_t i = 0; i < m; ++i)
        {
            std::array<float, d> x;
            for (std::size_t j = 0; j < d; ++j)
            {
                x[j] = q[i * d + j];
            }
            const auto tf = Robot::eefk(x);
            for (int r = 0; r < 3; ++r)
            {
                for (int c = 0; c < 3; ++c)
                {
                    out[i * 12 + 4 * r + c] = tf.linear()(r, c);
                }
                out[i * 12 + 4 * r + 3] = tf.translation()[r];
            }
        }
        return 0;
    }
}  // namespace refv
