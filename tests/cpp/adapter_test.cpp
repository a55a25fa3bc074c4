// Exercises the OMPL adapters of include/b200mp_ompl.hpp through the base-class interfaces, the way the reference's own
// tests drive their implementations:
//   * validators: the known answers of the VAMP demo/test (demos/vamp/VAMPPlanning.cpp:43-47,87-88,
//     tests/vamp/vamp_test.cpp:29-88): start and goal valid in the 14-sphere cage, straight edge invalid; plus agreement
//     of single calls with the batch extensions;
//   * NearestNeighborsB200<State *>: the body of tests/datastructures/nearestneighbors.cpp:stateSpaceTest against
//     NearestNeighborsLinear (distances equal within 1e-6), nearestR(inf), k > size, remove down to empty, and the
//     "No elements found" exception.
// Usage: adapter_test            -> runs everything on GPU 0 (exit 0 on success)
//        adapter_test --no-gpu   -> expects engine construction to fail with ompl::Exception (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <vector>

#include "b200mp_ompl.hpp"
#include <ompl/datastructures/NearestNeighborsLinear.h>

#include <atomic>
#include <thread>

namespace ob = ompl::base;
namespace b2 = ompl::b200;
using RVState = ob::RealVectorStateSpace::StateType;

static int failures = 0;
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

struct OwnedState
{
    RVState s;
    std::vector<double> v;
    explicit OwnedState(std::vector<double> x) : v(std::move(x)) { s.values = v.data(); }
    OwnedState(const OwnedState &) = delete;
};

static double l2(const ob::State *a, const ob::State *b, std::size_t dim)
{
    const double *x = a->as<RVState>()->values, *y = b->as<RVState>()->values;
    double d = 0;
    for (std::size_t i = 0; i < dim; ++i) d += (x[i] - y[i]) * (x[i] - y[i]);
    return std::sqrt(d);
}

static void test_validators(const b2::EnginePtr &eng)
{
    auto env = std::make_shared<b2::Environment>(eng);
    const float c[14][3] = {{0.55, 0, 0.25}, {0.35, 0.35, 0.25}, {0, 0.55, 0.25}, {-0.55, 0, 0.25}, {-0.35, -0.35, 0.25},
                            {0, -0.55, 0.25}, {0.35, -0.35, 0.25}, {0.35, 0.35, 0.8}, {0, 0.55, 0.8}, {-0.35, 0.35, 0.8},
                            {-0.55, 0, 0.8}, {-0.35, -0.35, 0.8}, {0, -0.55, 0.8}, {0.35, -0.35, 0.8}};
    for (const auto &p : c) env->addSphere({p[0], p[1], p[2]}, 0.2F);
    env->sort();
    env->commit();

    ob::SpaceInformation *si = nullptr;  // the validators only store it (OMPL core cannot be built in this image)
    std::shared_ptr<ob::StateValidityChecker> svc = std::make_shared<b2::B200StateValidityChecker<b2::Panda>>(si, env);
    std::shared_ptr<ob::MotionValidator> mv = std::make_shared<b2::B200MotionValidator<b2::Panda>>(si, env);

    OwnedState start({0., -0.785, 0., -2.356, 0., 1.571, 0.785}), goal({2.35, 1., 0., -0.8, 0., 2.5, 0.785});
    EXPECT(svc->isValid(&start.s));
    EXPECT(svc->isValid(&goal.s));
    EXPECT(!mv->checkMotion(&start.s, &goal.s));
    EXPECT(mv->checkMotion(&start.s, &start.s));
    std::pair<ob::State *, double> last(&goal.s, 0.5);
    EXPECT(!mv->checkMotion(&start.s, &goal.s, last));
    EXPECT(last.first == nullptr && last.second == 0.0);  // the reference bridge's behaviour (VampMotionValidator.h:43-52)
    EXPECT(mv->getValidMotionCount() == 0 && mv->getInvalidMotionCount() == 0);  // counters untouched, as in the bridge

    // batch extensions agree with single calls
    std::array<double, 7> lo, hi;
    b2::getRobotBounds<b2::Panda>(lo, hi);
    EXPECT(std::fabs(lo[0] + 2.9671) < 1e-6 && std::fabs(hi[3] - 0.0873) < 1e-4);
    std::mt19937_64 rng(5);
    std::vector<std::unique_ptr<OwnedState>> st;
    for (int i = 0; i < 400; ++i)
    {
        std::vector<double> q(7);
        for (int j = 0; j < 7; ++j) q[j] = std::uniform_real_distribution<double>(lo[j], hi[j])(rng);
        st.emplace_back(new OwnedState(q));
    }
    std::vector<const ob::State *> a, b;
    for (int i = 0; i < 400; ++i) { a.push_back(&st[i]->s); b.push_back(&st[(i * 7 + 1) % 400]->s); }
    std::vector<std::uint8_t> sv(400), mvb(400);
    auto *svc_b = dynamic_cast<b2::B200StateValidityChecker<b2::Panda> *>(svc.get());
    auto *mv_b = dynamic_cast<b2::B200MotionValidator<b2::Panda> *>(mv.get());
    svc_b->isValidBatch(a.data(), a.size(), sv.data());
    mv_b->checkMotions(a.data(), b.data(), a.size(), mvb.data());
    int nvalid = 0;
    for (int i = 0; i < 400; i += 7)
    {
        EXPECT(svc->isValid(a[i]) == (sv[i] != 0));
        EXPECT(mv->checkMotion(a[i], b[i]) == (mvb[i] != 0));
    }
    for (auto v : sv) nvalid += v;
    EXPECT(nvalid > 10 && nvalid < 390);

    // an environment that was not committed is an error surfaced as ompl::Exception
    auto env2 = std::make_shared<b2::Environment>(eng);
    env2->addSphere({1, 0, 0}, 0.1F);
    b2::B200StateValidityChecker<b2::Panda> bad(si, env2);
    bool threw = false;
    try { bad.isValid(&start.s); } catch (const ompl::Exception &) { threw = true; }
    EXPECT(threw);
}

// "must be thread safe" (StateValidityChecker.h:87-89, MotionValidator.h:60-63): four planner threads share one validator pair
static void test_thread_safety(const b2::EnginePtr &eng)
{
    auto env = std::make_shared<b2::Environment>(eng);
    for (int i = 0; i < 40; ++i)
        env->addSphere({0.35F + 0.02F * static_cast<float>(i % 9), 0.1F * static_cast<float>(i % 7) - 0.3F, 0.1F + 0.03F * static_cast<float>(i)}, 0.06F);
    env->sort();
    env->commit();
    ob::SpaceInformation *si = nullptr;
    b2::B200StateValidityChecker<b2::Panda> svc(si, env);
    b2::B200MotionValidator<b2::Panda> mv(si, env);
    std::array<double, 7> lo, hi;
    b2::getRobotBounds<b2::Panda>(lo, hi);
    std::mt19937_64 rng(11);
    std::vector<std::unique_ptr<OwnedState>> st;
    for (int i = 0; i < 256; ++i)
    {
        std::vector<double> q(7);
        for (int j = 0; j < 7; ++j) q[j] = std::uniform_real_distribution<double>(lo[j], hi[j])(rng);
        st.emplace_back(new OwnedState(q));
    }
    std::vector<int> want_s(256), want_m(256);
    for (int i = 0; i < 256; ++i)
    {
        want_s[i] = svc.isValid(&st[i]->s) ? 1 : 0;
        want_m[i] = mv.checkMotion(&st[i]->s, &st[(i + 1) % 256]->s) ? 1 : 0;
    }
    std::atomic<int> bad{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < 4; ++t)
        pool.emplace_back([&, t]() {
            for (int rep = 0; rep < 3; ++rep)
                for (int i = t; i < 256; i += 4)
                {
                    if ((svc.isValid(&st[i]->s) ? 1 : 0) != want_s[i]) ++bad;
                    if ((mv.checkMotion(&st[i]->s, &st[(i + 1) % 256]->s) ? 1 : 0) != want_m[i]) ++bad;
                }
        });
    for (auto &th : pool) th.join();
    EXPECT(bad.load() == 0);
}

// the other robots, a height field, the point-cloud filter and both tree builds through the adapters
static void test_more_shapes_and_robots(const b2::EnginePtr &eng)
{
    ob::SpaceInformation *si = nullptr;
    // terrain 32 x 32 cells of 0.125 m, 0.5 m below the base: the straight-up UR5 is free, one folded down into it is not
    std::vector<float> h(32 * 32, 0.1F);
    auto env = std::make_shared<b2::Environment>(eng);
    env->addHeightfield({0.F, 0.F, -0.5F}, {0.125F, 0.125F, 1.F}, 32, 32, h);
    env->commit();
    b2::B200StateValidityChecker<b2::UR5> ur5(si, env);
    OwnedState up({0., -1.5708, 0., 0., 0., 0.}), down({0., 0.9, 0.9, 0., 0., 0.});
    EXPECT(ur5.isValid(&up.s));
    EXPECT(!ur5.isValid(&down.s));
    b2::B200StateValidityChecker<b2::Baxter> baxter(si, env);
    OwnedState zero14(std::vector<double>(14, 0.0));
    (void)baxter.isValid(&zero14.s);  // 14 joints, two arms: must run (the decision is covered by the Python parity tests)

    // cloud: a wall of points at x = 0.6; filter it, build the tree both ways, the Panda stretched into it collides either way
    std::vector<float> wall;
    for (int i = 0; i < 60; ++i)
        for (int j = 0; j < 60; ++j) { wall.push_back(0.6F); wall.push_back(-0.3F + 0.01F * static_cast<float>(i)); wall.push_back(0.2F + 0.01F * static_cast<float>(j)); }
    const auto kept = b2::filterPointcloud(eng, wall, 0.015F, 1.5F, {0.F, 0.F, 0.4F}, {-1.5F, -1.5F, -1.1F}, {1.5F, 1.5F, 1.9F}, true);
    EXPECT(!kept.empty() && kept.size() < wall.size() && kept.size() % 3 == 0);
    OwnedState home({0., -0.785, 0., -2.356, 0., 1.571, 0.785}), reach({0., 1.2, 0., -0.6, 0., 1.9, 0.785});
    for (int device_build = 0; device_build < 2; ++device_build)
    {
        auto e2 = std::make_shared<b2::Environment>(eng);
        if (device_build) e2->addPointcloudDevice(kept.data(), kept.size() / 3, 0.012F, 0.18F, 0.0025F);
        else e2->addPointcloud(kept.data(), kept.size() / 3, 0.012F, 0.18F, 0.0025F);
        e2->commit();
        b2::B200StateValidityChecker<b2::Panda> panda(si, e2);
        EXPECT(panda.isValid(&home.s));
        EXPECT(!panda.isValid(&reach.s));
    }
}

static void test_nearest_neighbors(const b2::EnginePtr &eng, std::size_t dim, int n, int k)
{
    std::mt19937_64 rng(dim * 100 + n);
    std::vector<std::unique_ptr<OwnedState>> store;
    std::vector<ob::State *> states;
    for (int i = 0; i < n; ++i)
    {
        std::vector<double> q(dim);
        for (auto &x : q) x = std::uniform_real_distribution<double>(-3, 3)(rng);
        store.emplace_back(new OwnedState(q));
        states.push_back(&store.back()->s);
    }
    b2::NearestNeighborsB200<ob::State *> nn;
    ompl::NearestNeighbors<ob::State *> &proximity = nn;  // everything below goes through the abstract interface
    ompl::NearestNeighborsLinear<ob::State *> linear;
    auto df = [dim](ob::State *const &x, ob::State *const &y) { return l2(x, y, dim); };
    proximity.setDistanceFunction(df);
    linear.setDistanceFunction(df);
    bool threw = false;
    try { proximity.add(states[0]); } catch (const ompl::Exception &) { threw = true; }
    EXPECT(threw);  // coordinates must be bound first
    nn.bind(dim, [](ob::State *const &s) { return static_cast<const double *>(s->as<RVState>()->values); }, eng);
    EXPECT(proximity.reportsSortedResults());
    proximity.add(states);
    linear.add(states);
    EXPECT(static_cast<int>(proximity.size()) == n);
    std::vector<ob::State *> nghbr, truth;
    proximity.list(nghbr);
    EXPECT(nghbr.size() == proximity.size());
    const double eps = 1e-6;
    for (int i = 0; i < n; i += std::max(1, n / 40))
    {
        ob::State *s = proximity.nearest(states[i]);
        EXPECT(l2(s, states[i], dim) < eps);
        proximity.nearestK(states[i], k, nghbr);
        EXPECT(static_cast<int>(nghbr.size()) == std::min(k, n));
        linear.nearestK(states[i], nghbr.size(), truth);
        for (std::size_t t = 0; t < nghbr.size(); ++t) EXPECT(std::fabs(l2(states[i], truth[t], dim) - l2(states[i], nghbr[t], dim)) < eps);
        for (std::size_t t = 1; t < nghbr.size(); ++t) EXPECT(l2(states[i], nghbr[t - 1], dim) <= l2(states[i], nghbr[t], dim));
        proximity.nearestR(states[i], std::numeric_limits<double>::infinity(), nghbr);
        EXPECT(l2(nghbr[0], states[i], dim) < eps);
        EXPECT(nghbr.size() == proximity.size());
        linear.nearestR(states[i], std::numeric_limits<double>::infinity(), truth);
        for (std::size_t t = 0; t < nghbr.size(); ++t) EXPECT(std::fabs(l2(states[i], truth[t], dim) - l2(states[i], nghbr[t], dim)) < eps);
        proximity.nearestR(states[i], 2.0, nghbr);
        linear.nearestR(states[i], 2.0, truth);
        EXPECT(nghbr.size() == truth.size());
        proximity.nearestK(states[i], 2 * n, nghbr);  // k > size returns everything (NearestNeighborsGNAT.h:222-233)
        EXPECT(static_cast<int>(nghbr.size()) == n);
        bool found = false;
        for (auto *p : nghbr) found |= (p == states[i]);
        EXPECT(found);
    }
    proximity.nearestK(states[0], 0, nghbr);
    EXPECT(nghbr.empty());
    // a query element that is not in the container (PRM queries a milestone before adding it; RRT queries a temp motion)
    OwnedState probe(std::vector<double>(dim, 0.25));
    proximity.nearestK(&probe.s, 5, nghbr);
    linear.nearestK(&probe.s, 5, truth);
    for (std::size_t t = 0; t < nghbr.size(); ++t) EXPECT(nghbr[t] == truth[t]);
    // batch extension
    std::vector<std::vector<ob::State *>> many;
    nn.nearestKBatch(states, k, many);
    for (int i = 0; i < n; i += std::max(1, n / 25))
    {
        proximity.nearestK(states[i], k, nghbr);
        EXPECT(many[i] == nghbr);
    }
    // incremental removal down to empty
    for (int i = n - 1; i >= 0; --i)
    {
        if (i % 16 == 0 || i < 40)
        {
            EXPECT(proximity.remove(states[i]));
        }
        else
        {
            // bulk of the removals without the per-step checks (each is a device update)
            EXPECT(proximity.remove(states[i]));
            continue;
        }
        EXPECT(static_cast<int>(proximity.size()) == i);
        proximity.list(nghbr);
        EXPECT(static_cast<int>(nghbr.size()) == i);
        if (i > 0)
        {
            ob::State *s = proximity.nearest(states[i]);  // the removed element must not come back
            EXPECT(s != states[i]);
        }
    }
    EXPECT(!proximity.remove(states[0]));
    threw = false;
    try { proximity.nearest(states[0]); }
    catch (const std::exception &e) { threw = std::string(e.what()).substr(0, 17) == "No elements found"; }
    EXPECT(threw);
    proximity.clear();
    proximity.add(states[3]);
    EXPECT(proximity.size() == 1 && proximity.nearest(states[5]) == states[3]);
}

// A distance function that is not the Euclidean one must be rejected rather than silently replaced.
static void test_metric_guard(const b2::EnginePtr &eng)
{
    std::vector<std::unique_ptr<OwnedState>> store;
    b2::NearestNeighborsB200<ob::State *> nn;
    nn.setDistanceFunction([](ob::State *const &x, ob::State *const &y) { return 2.0 * l2(x, y, 3) + 1.0; });
    nn.bind(3, [](ob::State *const &s) { return static_cast<const double *>(s->as<RVState>()->values); }, eng);
    for (int i = 0; i < 10; ++i)
    {
        store.emplace_back(new OwnedState({0.1 * i, 0.2 * i, -0.1 * i}));
        nn.add(&store.back()->s);
    }
    std::vector<ob::State *> out;
    bool threw = false;
    try { nn.nearestK(&store[0]->s, 3, out); } catch (const ompl::Exception &) { threw = true; }
    EXPECT(threw);
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0)
    {
        try
        {
            b2::Engine e(0);
        }
        catch (const ompl::Exception &e)
        {
            std::printf("no-gpu ok: %s\n", e.what());
            return 0;
        }
        std::printf("FAIL: engine construction succeeded without a GPU\n");
        return 1;
    }
    try
    {
        auto eng = std::make_shared<b2::Engine>(0);
        test_validators(eng);
        test_thread_safety(eng);
        test_more_shapes_and_robots(eng);
        test_nearest_neighbors(eng, 7, 1000, 10);
        test_nearest_neighbors(eng, 3, 60, 30);
        test_nearest_neighbors(eng, 20, 300, 10);
        test_metric_guard(eng);
    }
    catch (const std::exception &e)
    {
        std::printf("FAIL: unexpected exception: %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "adapter_test: %d FAILURES\n" : "adapter_test: all passed\n", failures);
    return failures ? 1 : 0;
}
