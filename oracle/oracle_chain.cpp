// TEST INFRASTRUCTURE ONLY -- CPU restatement of the planar kinematic-chain path (BASELINE config 5).
// Reference followed (paths under /root/reference):
//   demos/KinematicChain.h:105-121    KinematicChainSpace::distance (sum of joint end-point distances, cumulative angles)
//   demos/KinematicChain.h:150-176    KinematicChainSpace::interpolate (shortest way round per joint, wrap to [-pi, pi])
//   demos/KinematicChain.h:209-230    KinematicChainValidityChecker::isValidImpl (n links + a 0.001 tip segment)
//   demos/KinematicChain.h:233-276    self / environment intersection tests, float-epsilon thresholds
//   demos/KinematicChain.h:279-315    createHornEnvironment
//   src/ompl/base/src/DiscreteMotionValidator.cpp:48-145   both checkMotion forms
//   src/ompl/base/src/StateSpace.cpp:851-854               validSegmentCount
// Pinned against oracle/_ref/libref_chain.so (those reference files compiled unmodified) by tests/test_chain_cpu.py and
// the golden vectors tests/golden/chain.npz generated from it. Compiled with -ffp-contract=off like the reference demo.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <utility>
#include <vector>

namespace
{
    constexpr double kPi = 3.141592653589793238462643383279502884;

    struct Seg { double x0, y0, x1, y1; };

    struct Chain
    {
        int n;
        double link, lvs;
        std::vector<Seg> env;
    };

    double distance(const Chain &c, const double *a, const double *b)
    {
        double theta1 = 0., theta2 = 0., dx = 0., dy = 0., dist = 0.;
        for (int i = 0; i < c.n; ++i)
        {
            theta1 += a[i];
            theta2 += b[i];
            dx += std::cos(theta1) - std::cos(theta2);
            dy += std::sin(theta1) - std::sin(theta2);
            dist += std::sqrt(dx * dx + dy * dy);
        }
        return dist * c.link;
    }

    void interpolate(const Chain &c, const double *from, const double *to, double t, double *out)
    {
        for (int i = 0; i < c.n; ++i)
        {
            double diff = to[i] - from[i];
            if (std::fabs(diff) <= kPi)
                out[i] = from[i] + diff * t;
            else
            {
                if (diff > 0.0)
                    diff = 2.0 * kPi - diff;
                else
                    diff = -2.0 * kPi - diff;
                out[i] = from[i] - diff * t;
                if (out[i] > kPi)
                    out[i] -= 2.0 * kPi;
                else if (out[i] < -kPi)
                    out[i] += 2.0 * kPi;
            }
        }
    }

    bool intersects(const Seg &s0, const Seg &s1)
    {
        const double s10_x = s0.x1 - s0.x0, s10_y = s0.y1 - s0.y0;
        const double s32_x = s1.x1 - s1.x0, s32_y = s1.y1 - s1.y0;
        const double denom = s10_x * s32_y - s32_x * s10_y;
        if (std::fabs(denom) < std::numeric_limits<double>::epsilon()) return false;
        const bool pos = denom > 0;
        const double s02_x = s0.x0 - s1.x0, s02_y = s0.y0 - s1.y0;
        const double s_numer = s10_x * s02_y - s10_y * s02_x;
        if ((s_numer < std::numeric_limits<float>::epsilon()) == pos) return false;
        const double t_numer = s32_x * s02_y - s32_y * s02_x;
        if ((t_numer < std::numeric_limits<float>::epsilon()) == pos) return false;
        if (((s_numer - denom > -std::numeric_limits<float>::epsilon()) == pos) ||
            ((t_numer - denom > std::numeric_limits<float>::epsilon()) == pos))
            return false;
        return true;
    }

    bool valid(const Chain &c, const double *q)
    {
        std::vector<Seg> segs;
        double theta = 0., x = 0., y = 0., xN, yN;
        for (int i = 0; i < c.n; ++i)
        {
            theta += q[i];
            xN = x + std::cos(theta) * c.link;
            yN = y + std::sin(theta) * c.link;
            segs.push_back({x, y, xN, yN});
            x = xN;
            y = yN;
        }
        xN = x + std::cos(theta) * 0.001;
        yN = y + std::sin(theta) * 0.001;
        segs.push_back({x, y, xN, yN});
        for (std::size_t i = 0; i < segs.size(); ++i)
            for (std::size_t j = i + 1; j < segs.size(); ++j)
                if (intersects(segs[i], segs[j])) return false;
        for (const auto &s : segs)
            for (const auto &e : c.env)
                if (intersects(s, e)) return false;
        return true;
    }

    int segment_count(const Chain &c, const double *a, const double *b)
    {
        return static_cast<int>(static_cast<unsigned int>(std::ceil(distance(c, a, b) / c.lvs)));
    }
}  // namespace

extern "C"
{
    int orc_horn(int d, double eps, double *out)
    {
        std::vector<Seg> env;
        double w = 1. / (double)d, x = w, y = -eps, xN, yN, theta = 0., scale = w * (1. + kPi * eps);
        for (int i = 0; i < d - 1; ++i)
        {
            theta += kPi / (double)d;
            xN = x + std::cos(theta) * scale;
            yN = y + std::sin(theta) * scale;
            env.push_back({x, y, xN, yN});
            x = xN;
            y = yN;
        }
        theta = 0.;
        x = w;
        y = eps;
        scale = w * (1.0 - kPi * eps);
        for (int i = 0; i < d - 1; ++i)
        {
            theta += kPi / d;
            xN = x + std::cos(theta) * scale;
            yN = y + std::sin(theta) * scale;
            env.push_back({x, y, xN, yN});
            x = xN;
            y = yN;
        }
        for (std::size_t i = 0; i < env.size(); ++i)
        {
            out[4 * i] = env[i].x0; out[4 * i + 1] = env[i].y0; out[4 * i + 2] = env[i].x1; out[4 * i + 3] = env[i].y1;
        }
        return static_cast<int>(env.size());
    }

    void *orc_create(int n_links, double link_length, const double *segments, int n_seg, double lvs_fraction)
    {
        auto *c = new Chain{n_links, link_length, 0.0, {}};
        for (int i = 0; i < n_seg; ++i) c->env.push_back({segments[4 * i], segments[4 * i + 1], segments[4 * i + 2], segments[4 * i + 3]});
        double e = 0.0;  // RealVectorStateSpace::getMaximumExtent with bounds [-pi, pi]^n
        for (int i = 0; i < n_links; ++i)
        {
            const double d = kPi - (-kPi);
            e += d * d;
        }
        c->lvs = std::sqrt(e) * lvs_fraction;
        return c;
    }
    void orc_destroy(void *h) { delete static_cast<Chain *>(h); }
    double orc_longest_valid_segment(void *h) { return static_cast<Chain *>(h)->lvs; }

    void orc_distance(void *h, const double *a, const double *b, std::int64_t m, double *out)
    {
        const Chain &c = *static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i) out[i] = distance(c, a + i * c.n, b + i * c.n);
    }
    void orc_interpolate(void *h, const double *a, const double *b, const double *t, std::int64_t m, double *out)
    {
        const Chain &c = *static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i) interpolate(c, a + i * c.n, b + i * c.n, t[i], out + i * c.n);
    }
    void orc_valid_states(void *h, const double *q, std::int64_t m, std::uint8_t *out)
    {
        const Chain &c = *static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i) out[i] = valid(c, q + i * c.n) ? 1 : 0;
    }
    void orc_check_motions(void *h, const double *a, const double *b, std::int64_t m, int form, std::uint8_t *out, double *last_t, int *nd_out)
    {
        const Chain &c = *static_cast<Chain *>(h);
        std::vector<double> test(c.n);
        for (std::int64_t e = 0; e < m; ++e)
        {
            const double *s1 = a + e * c.n, *s2 = b + e * c.n;
            const int nd = segment_count(c, s1, s2);
            if (nd_out) nd_out[e] = nd;
            bool result = true;
            double lv = -1.0;
            if (form == 0)
            {
                if (!valid(c, s2))
                    result = false;
                else if (nd >= 2)
                {
                    std::queue<std::pair<int, int>> pos;
                    pos.emplace(1, nd - 1);
                    while (!pos.empty())
                    {
                        const std::pair<int, int> x = pos.front();
                        const int mid = (x.first + x.second) / 2;
                        interpolate(c, s1, s2, (double)mid / (double)nd, test.data());
                        if (!valid(c, test.data()))
                        {
                            result = false;
                            break;
                        }
                        pos.pop();
                        if (x.first < mid) pos.emplace(x.first, mid - 1);
                        if (x.second > mid) pos.emplace(mid + 1, x.second);
                    }
                }
            }
            else
            {
                if (nd > 1)
                    for (int j = 1; j < nd; ++j)
                    {
                        interpolate(c, s1, s2, (double)j / (double)nd, test.data());
                        if (!valid(c, test.data()))
                        {
                            lv = (double)(j - 1) / (double)nd;
                            result = false;
                            break;
                        }
                    }
                if (result && !valid(c, s2))
                {
                    lv = (double)(nd - 1) / (double)nd;
                    result = false;
                }
            }
            out[e] = result ? 1 : 0;
            if (last_t) last_t[e] = lv;
        }
    }
}
