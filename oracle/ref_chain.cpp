// TEST INFRASTRUCTURE ONLY -- C wrapper around the UNMODIFIED reference files demos/KinematicChain.h
// (KinematicChainSpace::distance / interpolate, KinematicChainValidityChecker, createHornEnvironment) and
// src/ompl/base/src/DiscreteMotionValidator.cpp (both checkMotion forms), compiled where they lie under /root/reference
// against the shim headers in oracle/shim_ompl (see oracle/Makefile). Output: oracle/_ref/libref_chain.so.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unistd.h>
#include <vector>

#include "KinematicChain.h"
#include "ompl/base/DiscreteMotionValidator.h"

namespace ob = ompl::base;

namespace
{
    struct Chain
    {
        Environment env;
        std::shared_ptr<KinematicChainSpace> space;
        std::shared_ptr<ob::SpaceInformation> si;
        std::shared_ptr<ob::DiscreteMotionValidator> mv;
        unsigned int n;
    };
    struct Tmp
    {
        ob::RealVectorStateSpace::StateType s;
        explicit Tmp(const double *v) { s.values = const_cast<double *>(v); }
    };
}  // namespace

extern "C"
{
    // createHornEnvironment(d, eps) -> segments[2(d-1)][4]; the demo also writes environment_<d>.dat, sent to /tmp here
    int refc_horn(int d, double eps, double *out)
    {
        char cwd[4096];
        if (!getcwd(cwd, sizeof(cwd))) return -1;
        if (chdir("/tmp") != 0) return -1;
        Environment e = createHornEnvironment(d, eps);
        if (chdir(cwd) != 0) return -1;
        for (std::size_t i = 0; i < e.size(); ++i)
        {
            out[4 * i] = e[i].x0; out[4 * i + 1] = e[i].y0; out[4 * i + 2] = e[i].x1; out[4 * i + 3] = e[i].y1;
        }
        return static_cast<int>(e.size());
    }

    void *refc_create(int n_links, double link_length, const double *segments, int n_seg, double lvs_fraction)
    {
        auto *c = new Chain();
        c->n = n_links;
        for (int i = 0; i < n_seg; ++i) c->env.emplace_back(segments[4 * i], segments[4 * i + 1], segments[4 * i + 2], segments[4 * i + 3]);
        c->space = std::make_shared<KinematicChainSpace>(n_links, link_length, &c->env);
        c->space->setLongestValidSegmentFraction(lvs_fraction);
        c->space->setup();
        c->si = std::make_shared<ob::SpaceInformation>(c->space);
        c->si->setStateValidityChecker(std::make_shared<KinematicChainValidityChecker>(c->si));
        c->mv = std::make_shared<ob::DiscreteMotionValidator>(c->si.get());
        return c;
    }
    void refc_destroy(void *h) { delete static_cast<Chain *>(h); }
    double refc_longest_valid_segment(void *h) { return static_cast<Chain *>(h)->space->getLongestValidSegmentLength(); }

    void refc_distance(void *h, const double *a, const double *b, std::int64_t m, double *out)
    {
        auto *c = static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i)
        {
            Tmp x(a + i * c->n), y(b + i * c->n);
            out[i] = c->space->distance(&x.s, &y.s);
        }
    }
    void refc_interpolate(void *h, const double *a, const double *b, const double *t, std::int64_t m, double *out)
    {
        auto *c = static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i)
        {
            Tmp x(a + i * c->n), y(b + i * c->n), o(out + i * c->n);
            c->space->interpolate(&x.s, &y.s, t[i], &o.s);
        }
    }
    void refc_valid_states(void *h, const double *q, std::int64_t m, std::uint8_t *out)
    {
        auto *c = static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i)
        {
            Tmp x(q + i * c->n);
            out[i] = c->si->isValid(&x.s) ? 1 : 0;
        }
    }
    // form 0: checkMotion(s1, s2) (bisection order); form 1: checkMotion(s1, s2, lastValid) (linear order), last_t = lastValid.second
    void refc_check_motions(void *h, const double *a, const double *b, std::int64_t m, int form, std::uint8_t *out, double *last_t, int *nd)
    {
        auto *c = static_cast<Chain *>(h);
        for (std::int64_t i = 0; i < m; ++i)
        {
            Tmp x(a + i * c->n), y(b + i * c->n);
            if (nd) nd[i] = static_cast<int>(c->space->validSegmentCount(&x.s, &y.s));
            if (form == 0)
                out[i] = c->mv->checkMotion(&x.s, &y.s) ? 1 : 0;
            else
            {
                std::pair<ob::State *, double> lv(nullptr, -1.0);
                out[i] = c->mv->checkMotion(&x.s, &y.s, lv) ? 1 : 0;
                if (last_t) last_t[i] = lv.second;
            }
        }
    }
}

// the genuine implementation file of the validator under test
#include "ompl/base/src/DiscreteMotionValidator.cpp"
