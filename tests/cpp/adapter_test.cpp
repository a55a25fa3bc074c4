// Exercises the OMPL adapters of include/b200mp_ompl.hpp through the base-class interfaces, the way the reference's own
// tests drive their implementations:
//   * validators: the known answers of the VAMP demo/test (demos/vamp/VAMPPlanning.cpp:43-47,87-88,
//     tests/vamp/vamp_test.cpp:29-88): start and goal valid in the 14-sphere cage, straight edge invalid; plus agreement
//     of single calls with the batch extensions;
//   * NearestNeighborsB200<State *>: the body of tests/datastructures/nearestneighbors.cpp:stateSpaceTest against
//     NearestNeighborsLinear (distances equal within 1e-6), nearestR(inf), k > size, remove down to empty, and the
//     "No elements found" exception; randomAccessPatternTest (:205-280: rounds of add / nearestK / nearestR / remove);
//   * the planar chain (config 5): validity checker, DiscreteMotionValidator (both forms, counters, lastValid) and the
//     container under the chain metric, through the OMPL base classes;
//   * batched call sites (rewireBatch, checkMotionsFrom, shortcutBatch) against single calls;
//   * roadmap hand-off: buildRoadmap -> PlannerData -> the loop of PRM(const PlannerData &) (PRM.cpp:96-156), file round trip, GraphML.
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
#include "b200mp_ompl_roadmap.hpp"
#include <ompl/datastructures/NearestNeighborsLinear.h>

#include <map>
#include <numeric>
#include <sstream>
#include <unordered_set>

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

    // a grasped object (collision/attachments.hh): a 6 cm ball 5 cm beyond the hand leaves the start state valid,
    // a 60 cm one reaches the cage; clearing it restores the bare robot. No re-commit.
    const std::array<float, 12> tf = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0.05F};
    env->setAttachment(tf, {{0.F, 0.F, 0.F, 0.03F}});
    EXPECT(svc->isValid(&start.s));
    env->setAttachment(tf, {{0.F, 0.F, 0.F, 0.03F}, {0.F, 0.F, 0.1F, 0.3F}});
    EXPECT(!svc->isValid(&start.s));
    EXPECT(!mv->checkMotion(&start.s, &start.s));
    env->clearAttachment();
    EXPECT(svc->isValid(&start.s));

    // single calls through the resident kernel (b200mp_session_*): same answers, no launch per call
    mv_b->useSession(true);
    EXPECT(!mv->checkMotion(&start.s, &goal.s));
    EXPECT(mv->checkMotion(&start.s, &start.s));
    for (int i = 0; i < 400; i += 3) EXPECT(mv->checkMotion(a[i], b[i]) == (mvb[i] != 0));
    mv_b->useSession(false);

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


// tests/datastructures/nearestneighbors.cpp:205-280 (randomAccessPatternTest), with the reference's sizes scaled down
// (n = 200 states -> 40 rounds of 10 adds, 10 queries with k in [1, 30] and r in [0, 3], then each stored state removed with p = 1/2)
static void test_random_access_pattern(const b2::EnginePtr &eng, std::size_t dim)
{
    std::mt19937_64 rng(77 + dim);
    std::uniform_real_distribution<double> coord(-3, 3), u01(0, 1);
    const int m = 40, n = 10, maxk = 30;
    const double eps = 1e-6;
    std::vector<std::unique_ptr<OwnedState>> store;
    auto fresh = [&]() {
        std::vector<double> q(dim);
        for (auto &x : q) x = coord(rng);
        store.emplace_back(new OwnedState(q));
        return static_cast<ob::State *>(&store.back()->s);
    };
    b2::NearestNeighborsB200<ob::State *> nn;
    nn.bind(dim, [](ob::State *const &s) { return static_cast<const double *>(s->as<RVState>()->values); }, eng);
    ompl::NearestNeighbors<ob::State *> &proximity = nn;
    ompl::NearestNeighborsLinear<ob::State *> linear;
    auto df = [dim](ob::State *const &x, ob::State *const &y) { return l2(x, y, dim); };
    proximity.setDistanceFunction(df);
    linear.setDistanceFunction(df);
    std::unordered_set<ob::State *> states;
    std::vector<ob::State *> nghbr, truth;
    for (int i = 0; i < m; ++i)
    {
        for (int j = 0; j < n; ++j)
        {
            ob::State *s = fresh();
            states.insert(s);
            proximity.add(s);
            linear.add(s);
        }
        for (int j = 0; j < n; ++j)
        {
            ob::State *s = fresh();
            const int k = 1 + static_cast<int>(u01(rng) * maxk) % maxk;
            linear.nearestK(s, k, truth);
            proximity.nearestK(s, k, nghbr);
            EXPECT(nghbr.size() == truth.size());
            for (std::size_t p = 0; p < std::min(nghbr.size(), truth.size()); ++p) EXPECT(std::fabs(l2(s, truth[p], dim) - l2(s, nghbr[p], dim)) < eps);
            const double r = 3.0 * u01(rng);
            linear.nearestR(s, r, truth);
            proximity.nearestR(s, r, nghbr);
            EXPECT(nghbr.size() == truth.size());
            for (std::size_t p = 0; p < std::min(nghbr.size(), truth.size()); ++p) EXPECT(std::fabs(l2(s, truth[p], dim) - l2(s, nghbr[p], dim)) < eps);
        }
        for (auto it = states.begin(); it != states.end();)
        {
            if (u01(rng) < .5)
            {
                const std::size_t szLinear = linear.size(), sz = proximity.size();
                linear.remove(*it);
                const bool removed = proximity.remove(*it);
                it = states.erase(it);
                EXPECT(removed);
                EXPECT(proximity.size() == sz - 1);
                EXPECT(sz == szLinear);
                EXPECT(proximity.size() == states.size());
                proximity.list(nghbr);
                linear.list(truth);
                EXPECT(nghbr.size() == truth.size());
            }
            else
                ++it;
        }
    }
}

// config 5 through the OMPL interfaces: 20-link chain in the horn environment of demos/KinematicChainBenchmark.cpp:48-49
static void test_chain_adapters(const b2::EnginePtr &eng)
{
    const unsigned int links = 20;
    const double len = 1.0 / links;
    auto chain = std::make_shared<b2::KinematicChain>(eng, links, len, b2::KinematicChain::hornEnvironment(links, std::log(static_cast<double>(links)) / links));
    ob::SpaceInformation *si = nullptr;
    std::shared_ptr<ob::StateValidityChecker> svc = std::make_shared<b2::B200ChainValidityChecker>(si, chain);
    std::shared_ptr<ob::MotionValidator> mv = std::make_shared<b2::B200DiscreteMotionValidator>(si, chain);
    auto *svc_b = dynamic_cast<b2::B200ChainValidityChecker *>(svc.get());
    auto *mv_b = dynamic_cast<b2::B200DiscreteMotionValidator *>(mv.get());

    // known answers, taken from the compiled reference (oracle/_ref/libref_chain.so = demos/KinematicChain.h): the chain stretched
    // along +x runs into the horn, the same chain pointing up (first joint at pi/2) is free
    OwnedState straight(std::vector<double>(links, 0.0));
    std::vector<double> upq(links, 0.0);
    upq[0] = 1.5708;
    OwnedState up(upq);
    EXPECT(!svc->isValid(&straight.s));
    EXPECT(svc->isValid(&up.s));

    std::mt19937_64 rng(3);
    std::uniform_real_distribution<double> first(1.5708 - 1.0, 1.5708 + 1.0), small(-0.35, 0.35);
    std::vector<std::unique_ptr<OwnedState>> st;
    for (int i = 0; i < 300; ++i)
    {
        std::vector<double> q(links);
        for (auto &x : q) x = small(rng);  // gently curved chains around the upright pose: about half of them are valid
        q[0] = first(rng);
        st.emplace_back(new OwnedState(q));
    }
    std::vector<const ob::State *> a, b;
    for (int i = 0; i < 300; ++i) { a.push_back(&st[i]->s); b.push_back(&st[(i * 7 + 3) % 300]->s); }
    std::vector<std::uint8_t> sv(300), mvb(300);
    svc_b->isValidBatch(a.data(), a.size(), sv.data());
    int nvalid = std::accumulate(sv.begin(), sv.end(), 0);
    EXPECT(nvalid > 20 && nvalid < 290);
    mv_b->checkMotions(a.data(), b.data(), a.size(), mvb.data());
    const unsigned int v0 = mv->getValidMotionCount(), i0 = mv->getInvalidMotionCount();
    EXPECT(v0 + i0 == 300);
    EXPECT(static_cast<int>(v0) == std::accumulate(mvb.begin(), mvb.end(), 0));
    OwnedState scratch(std::vector<double>(links, 0.0));
    int checked = 0;
    for (int i = 0; i < 300; i += 5)
    {
        EXPECT(svc->isValid(a[i]) == (sv[i] != 0));
        if (!sv[i]) continue;  // checkMotion assumes a valid start (MotionValidator.h:72-77)
        const bool ok = mv->checkMotion(a[i], b[i]);
        EXPECT(ok == (mvb[i] != 0));
        std::pair<ob::State *, double> last(&scratch.s, -7.0);
        const bool ok2 = mv->checkMotion(a[i], b[i], last);
        EXPECT(ok2 == ok);
        if (!ok)
        {
            // lastValid.second = (j - 1) / nd of the first invalid state j; the state handed back is the interpolation at that parameter
            // and, the start being valid, it is itself valid (DiscreteMotionValidator.cpp:60-71)
            EXPECT(last.second >= 0.0 && last.second < 1.0);
            OwnedState want(std::vector<double>(links, 0.0));
            chain->interpolate(a[i], b[i], last.second, &want.s);
            for (unsigned int j = 0; j < links; ++j) EXPECT(scratch.v[j] == want.v[j]);
            EXPECT(svc->isValid(&scratch.s));
        }
        else
            EXPECT(last.second == -7.0);  // untouched on success, like the reference
        ++checked;
    }
    EXPECT(checked > 10);
    EXPECT(mv->getValidMotionCount() + mv->getInvalidMotionCount() == 300 + 2u * static_cast<unsigned>(checked));

    // nearest neighbours under KinematicChainSpace::distance, against NearestNeighborsLinear with the same metric
    b2::NearestNeighborsB200<ob::State *> nn;
    auto df = [&chain](ob::State *const &x, ob::State *const &y) { return chain->distance(x, y); };
    nn.setDistanceFunction(df);
    nn.bindChain(links, len, [](ob::State *const &s) { return static_cast<const double *>(s->as<RVState>()->values); }, eng);
    ompl::NearestNeighborsLinear<ob::State *> linear;
    linear.setDistanceFunction(df);
    std::vector<ob::State *> all;
    for (auto &p : st) all.push_back(&p->s);
    nn.add(all);
    linear.add(all);
    std::vector<ob::State *> got, want;
    for (int i = 0; i < 300; i += 17)
    {
        nn.nearestK(all[i], 10, got);
        linear.nearestK(all[i], 10, want);
        EXPECT(got.size() == want.size());
        for (std::size_t t = 0; t < std::min(got.size(), want.size()); ++t) EXPECT(std::fabs(df(all[i], got[t]) - df(all[i], want[t])) < 1e-9);
    }
    // the guard still rejects a functor that is not the bound metric
    b2::NearestNeighborsB200<ob::State *> wrong;
    wrong.setDistanceFunction([](ob::State *const &x, ob::State *const &y) { return l2(x, y, 20); });
    wrong.bindChain(links, len, [](ob::State *const &s) { return static_cast<const double *>(s->as<RVState>()->values); }, eng);
    wrong.add(all);
    bool threw = false;
    try { wrong.nearestK(all[0], 3, got); } catch (const ompl::Exception &) { threw = true; }
    EXPECT(threw);
}

// batched call sites of PRM::addMilestone, RRT* and PathSimplifier against the single calls they replace
static void test_batched_call_sites(const b2::EnginePtr &eng)
{
    auto env = std::make_shared<b2::Environment>(eng);
    const float cage[14][3] = {{0.55, 0, 0.25}, {0.35, 0.35, 0.25}, {0, 0.55, 0.25}, {-0.55, 0, 0.25}, {-0.35, -0.35, 0.25},
                               {0, -0.55, 0.25}, {0.35, -0.35, 0.25}, {0.35, 0.35, 0.8}, {0, 0.55, 0.8}, {-0.35, 0.35, 0.8},
                               {-0.55, 0, 0.8}, {-0.35, -0.35, 0.8}, {0, -0.55, 0.8}, {0.35, -0.35, 0.8}};
    for (const auto &p : cage) env->addSphere({p[0], p[1], p[2]}, 0.2F);
    env->sort();
    env->commit();
    ob::SpaceInformation *si = nullptr;
    b2::B200MotionValidator<b2::Panda> mv(si, env);
    std::array<double, 7> lo, hi;
    b2::getRobotBounds<b2::Panda>(lo, hi);
    std::mt19937_64 rng(21);
    std::vector<std::unique_ptr<OwnedState>> st;
    OwnedState centre({0., -0.785, 0., -2.356, 0., 1.571, 0.785});
    for (int i = 0; i < 48; ++i)
    {
        std::vector<double> q(7);
        for (int j = 0; j < 7; ++j) q[j] = std::min(hi[j], std::max(lo[j], centre.v[j] + std::uniform_real_distribution<double>(-1.2, 1.2)(rng)));
        st.emplace_back(new OwnedState(q));
    }
    std::vector<const ob::State *> nbh;
    for (auto &p : st) nbh.push_back(&p->s);
    std::vector<std::uint8_t> to(48), from(48), fwd(48), rev(48);
    mv.rewireBatch(&centre.s, nbh.data(), nbh.size(), to.data(), from.data());
    mv.checkMotionsFrom(&centre.s, nbh.data(), nbh.size(), fwd.data());
    mv.checkMotionsFrom(&centre.s, nbh.data(), nbh.size(), rev.data(), true);
    int nv = 0;
    for (int i = 0; i < 48; ++i)
    {
        EXPECT((to[i] != 0) == mv.checkMotion(nbh[i], &centre.s));
        EXPECT((from[i] != 0) == mv.checkMotion(&centre.s, nbh[i]));
        EXPECT(fwd[i] == from[i] && rev[i] == to[i]);
        nv += to[i];
    }
    EXPECT(nv > 0 && nv < 48);
    std::vector<std::uint8_t> only(48);
    mv.rewireBatch(&centre.s, nbh.data(), nbh.size(), nullptr, only.data());
    EXPECT(only == from);
    // shortcuts over a path of 12 way points
    std::vector<const ob::State *> path(nbh.begin(), nbh.begin() + 12);
    std::vector<std::pair<std::size_t, std::size_t>> pairs;
    for (std::size_t i = 0; i < 12; ++i)
        for (std::size_t j = i + 2; j < 12; j += 3) pairs.emplace_back(i, j);
    std::vector<std::uint8_t> sc(pairs.size());
    mv.shortcutBatch(path, pairs, sc.data());
    for (std::size_t c = 0; c < pairs.size(); ++c) EXPECT((sc[c] != 0) == mv.checkMotion(path[pairs[c].first], path[pairs[c].second]));
}

// GPU-built roadmap -> ompl::base::PlannerData -> the loop of PRM(const PlannerData &) (PRM.cpp:96-156), restated over the same calls
static void test_roadmap_handoff(const b2::EnginePtr &eng)
{
    auto env = std::make_shared<b2::Environment>(eng);
    const float c[14][3] = {{0.55, 0, 0.25}, {0.35, 0.35, 0.25}, {0, 0.55, 0.25}, {-0.55, 0, 0.25}, {-0.35, -0.35, 0.25},
                            {0, -0.55, 0.25}, {0.35, -0.35, 0.25}, {0.35, 0.35, 0.8}, {0, 0.55, 0.8}, {-0.35, 0.35, 0.8},
                            {-0.55, 0, 0.8}, {-0.35, -0.35, 0.8}, {0, -0.55, 0.8}, {0.35, -0.35, 0.8}};
    for (const auto &p : c) env->addSphere({p[0], p[1], p[2]}, 0.2F);
    env->sort();
    env->commit();
    const std::size_t n = 6000;
    const int k = 10;
    const b2::Roadmap rm = b2::buildRoadmap<b2::Panda>(env, 42, n, k);
    EXPECT(rm.numStates() == n && rm.dim == 7);
    EXPECT(rm.numEdges() >= n * k / 2 && rm.numEdges() <= n * k);
    EXPECT(static_cast<std::size_t>(rm.stats.n_valid_edges) == static_cast<std::size_t>(std::accumulate(rm.valid.begin(), rm.valid.end(), 0)));

    b2::RoadmapStates states(rm);
    ob::PlannerData data;
    const std::size_t added = b2::toPlannerData(rm, states, data);
    EXPECT(added == static_cast<std::size_t>(rm.stats.n_valid_edges));
    EXPECT(data.numVertices() == n && data.numEdges() == added);

    // PRM(const PlannerData &): every vertex cloned, every edge read with its weight, components united
    std::vector<std::vector<double>> cloned(data.numVertices());
    std::vector<unsigned int> parent(data.numVertices());
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](unsigned int x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
    std::size_t graph_edges = 0;
    double weight_err = 0;
    for (unsigned int v = 0; v < data.numVertices(); ++v)
    {
        const double *q = data.getVertex(v).getState()->as<RVState>()->values;
        cloned[v].assign(q, q + 7);
    }
    for (unsigned int v = 0; v < data.numVertices(); ++v)
    {
        std::vector<unsigned int> nb;
        data.getEdges(v, nb);
        for (unsigned int w : nb)
        {
            ob::Cost cost;
            EXPECT(data.getEdgeWeight(v, w, &cost));
            double d = 0;
            for (int j = 0; j < 7; ++j) d += (cloned[v][j] - cloned[w][j]) * (cloned[v][j] - cloned[w][j]);
            weight_err = std::max(weight_err, std::fabs(std::sqrt(d) - cost.value()));
            parent[find(v)] = find(w);
            ++graph_edges;
        }
    }
    EXPECT(graph_edges == added);
    EXPECT(weight_err < 1e-12);
    for (std::size_t i = 0; i < n; i += 997)
        for (int j = 0; j < 7; ++j) EXPECT(cloned[i][j] == static_cast<double>(rm.states[i * 7 + j]));
    std::map<unsigned int, std::size_t> comp;
    for (unsigned int v = 0; v < data.numVertices(); ++v) ++comp[find(v)];
    std::size_t biggest = 0;
    for (const auto &kv : comp) biggest = std::max(biggest, kv.second);
    EXPECT(biggest > n / 2);  // a connected roadmap of the cage's free space

    // file round trip and GraphML
    const std::string path = "/tmp/b200mp_adapter_test_roadmap.b2rm";
    b2::saveRoadmap(path, rm);
    const b2::Roadmap back = b2::loadRoadmap(path);
    EXPECT(back == rm);
    EXPECT(back.stats.n_valid_edges == rm.stats.n_valid_edges);
    std::remove(path.c_str());
    bool threw = false;
    try { b2::loadRoadmap("/tmp/b200mp_no_such_file.b2rm"); } catch (const ompl::Exception &) { threw = true; }
    EXPECT(threw);
    std::ostringstream xml;
    b2::writeGraphML(xml, rm);
    const std::string doc = xml.str();
    std::size_t nodes = 0, edges = 0;
    for (std::size_t p = doc.find("<node id="); p != std::string::npos; p = doc.find("<node id=", p + 1)) ++nodes;
    for (std::size_t p = doc.find("<edge id="); p != std::string::npos; p = doc.find("<edge id=", p + 1)) ++edges;
    EXPECT(nodes == n && edges == added);
    EXPECT(doc.find("attr.name=\"coords\"") != std::string::npos && doc.find("attr.name=\"weight\"") != std::string::npos);

    // several GPUs when the box has them: same roadmap, one thread per GPU
    std::shared_ptr<b2::Engine> eng1;
    try { eng1 = std::make_shared<b2::Engine>(1); } catch (const ompl::Exception &) {}
    if (eng1)
    {
        const b2::Roadmap two = b2::buildRoadmapMulti<b2::Panda>({env, env->cloneTo(eng1)}, 42, n, k);
        EXPECT(two == rm);
        std::printf("buildRoadmapMulti on 2 GPUs: identical roadmap (%zu edges)\n", two.numEdges());
    }
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
        test_random_access_pattern(eng, 3);
        test_random_access_pattern(eng, 7);
        test_chain_adapters(eng);
        test_batched_call_sites(eng);
        test_roadmap_handoff(eng);
    }
    catch (const std::exception &e)
    {
        std::printf("FAIL: unexpected exception: %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "adapter_test: %d FAILURES\n" : "adapter_test: all passed\n", failures);
    return failures ? 1 : 0;
}
