// TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference NearestNeighborsLinear<_T>
// (/root/reference/src/ompl/datastructures/NearestNeighborsLinear.h) behind a C wrapper, driven exactly as
// OMPL's planners drive it: _T = unsigned long vertex id, distance through a std::function that looks the
// states up on the side (PRM.h:345-348) and evaluates RealVectorStateSpace::distance semantics.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include <ompl/datastructures/NearestNeighborsLinear.h>

namespace
{
    struct Holder
    {
        ompl::NearestNeighborsLinear<unsigned long> nn;
        std::vector<double> pts;  // [n+1][dim]; slot n holds the current query
        int dim = 0;
        std::int64_t n = 0;
    };
}  // namespace

extern "C"
{
    void *refnn_create(const double *db, std::int64_t n, int dim)
    {
        auto *h = new Holder();
        h->dim = dim;
        h->n = n;
        h->pts.assign(db, db + n * dim);
        h->pts.resize((n + 1) * dim);
        h->nn.setDistanceFunction(
            [h](const unsigned long &a, const unsigned long &b)
            {
                const double *s1 = &h->pts[a * h->dim], *s2 = &h->pts[b * h->dim];
                double dist = 0.0;  // RealVectorStateSpace.cpp:230-242
                for (int i = 0; i < h->dim; ++i)
                {
                    const double diff = s1[i] - s2[i];
                    dist += diff * diff;
                }
                return std::sqrt(dist);
            });
        for (std::int64_t i = 0; i < n; ++i) h->nn.add(static_cast<unsigned long>(i));
        return h;
    }
    void refnn_destroy(void *p) { delete static_cast<Holder *>(p); }

    int refnn_remove(void *p, std::int64_t id) { return static_cast<Holder *>(p)->nn.remove(static_cast<unsigned long>(id)) ? 1 : 0; }
    std::int64_t refnn_size(void *p) { return static_cast<std::int64_t>(static_cast<Holder *>(p)->nn.size()); }

    // returns number of neighbours written (<= k)
    std::int64_t refnn_nearest_k(void *p, const double *q, int k, std::int64_t *idx, double *dist)
    {
        auto *h = static_cast<Holder *>(p);
        for (int i = 0; i < h->dim; ++i) h->pts[h->n * h->dim + i] = q[i];
        std::vector<unsigned long> nbh;
        h->nn.nearestK(static_cast<unsigned long>(h->n), static_cast<std::size_t>(k), nbh);
        for (std::size_t t = 0; t < nbh.size(); ++t)
        {
            idx[t] = static_cast<std::int64_t>(nbh[t]);
            dist[t] = h->nn.getDistanceFunction()(nbh[t], static_cast<unsigned long>(h->n));
        }
        return static_cast<std::int64_t>(nbh.size());
    }

    // nq independent nearestK queries on `nthreads` host threads; every thread drives its OWN container instance (the reference
    // container is not thread safe and the query lives in a side slot). idx / dist are [nq][k], padded with -1 / +inf.
    void refnn_nearest_k_batch(const double *db, std::int64_t n, int dim, const double *q, std::int64_t nq, int k, std::int64_t *idx,
                               double *dist, int nthreads)
    {
        std::atomic<std::int64_t> next{0};
        auto work = [&]()
        {
            void *h = refnn_create(db, n, dim);
            for (;;)
            {
                const std::int64_t i = next.fetch_add(1);
                if (i >= nq)
                    break;
                const std::int64_t c = refnn_nearest_k(h, q + i * dim, k, idx + i * k, dist + i * k);
                for (std::int64_t t = c; t < k; ++t)
                {
                    idx[i * k + t] = -1;
                    dist[i * k + t] = INFINITY;
                }
            }
            refnn_destroy(h);
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }

    std::int64_t refnn_nearest_r(void *p, const double *q, double radius, std::int64_t *idx, double *dist, std::int64_t cap)
    {
        auto *h = static_cast<Holder *>(p);
        for (int i = 0; i < h->dim; ++i) h->pts[h->n * h->dim + i] = q[i];
        std::vector<unsigned long> nbh;
        h->nn.nearestR(static_cast<unsigned long>(h->n), radius, nbh);
        for (std::size_t t = 0; t < nbh.size() && static_cast<std::int64_t>(t) < cap; ++t)
        {
            idx[t] = static_cast<std::int64_t>(nbh[t]);
            dist[t] = h->nn.getDistanceFunction()(nbh[t], static_cast<unsigned long>(h->n));
        }
        return static_cast<std::int64_t>(nbh.size());
    }

    // -1 and sets *threw when the container is empty ("No elements found ...", NearestNeighborsLinear.h:115)
    std::int64_t refnn_nearest(void *p, const double *q, int *threw)
    {
        auto *h = static_cast<Holder *>(p);
        for (int i = 0; i < h->dim; ++i) h->pts[h->n * h->dim + i] = q[i];
        *threw = 0;
        try
        {
            return static_cast<std::int64_t>(h->nn.nearest(static_cast<unsigned long>(h->n)));
        }
        catch (const ompl::Exception &)
        {
            *threw = 1;
            return -1;
        }
    }
}
