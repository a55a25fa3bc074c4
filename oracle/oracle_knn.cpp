// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's nearest-neighbour ground truth.
//
// Follows /root/reference/src/ompl/datastructures/NearestNeighborsLinear.h:119-132 (nearestK: order all
// elements by distance to the query, keep the first k; fewer than k elements -> all of them, sorted)
// and :135-142 (nearestR: keep d <= radius, sorted), with the metric of
// /root/reference/src/ompl/base/spaces/src/RealVectorStateSpace.cpp:230-242 (double, sqrt of the sum of
// squared differences accumulated in index order). Ties are ordered by database index so results are
// deterministic (the reference's partial_sort leaves tie order unspecified).
// Pinned against oracle/_ref/libref_nn.so (the unmodified NearestNeighborsLinear) in tests/test_oracle.py.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

namespace
{
    inline double rv_distance(const double *a, const double *b, int dim)
    {
        double dist = 0.0;
        for (int i = 0; i < dim; ++i)
        {
            const double diff = a[i] - b[i];
            dist += diff * diff;
        }
        return std::sqrt(dist);
    }

    // demos/KinematicChain.h:105-121: sum over joints of the 2-D distance between the joint end points,
    // angles accumulated along the chain, scaled by the link length.
    inline double chain_distance(const double *a, const double *b, int dim, double link_length)
    {
        double theta1 = 0., theta2 = 0., dx = 0., dy = 0., dist = 0.;
        for (int i = 0; i < dim; ++i)
        {
            theta1 += a[i];
            theta2 += b[i];
            dx += std::cos(theta1) - std::cos(theta2);
            dy += std::sin(theta1) - std::sin(theta2);
            dist += std::sqrt(dx * dx + dy * dy);
        }
        return dist * link_length;
    }
}  // namespace

extern "C"
{
    // metric 0 = RealVectorStateSpace L2, 1 = KinematicChainSpace (param = link length).
    // exclude_self != 0: query i skips database element i (PRM queries before inserting, PRM.cpp:583,601).
    // idx[nq][k] (-1 padded), dist[nq][k] (+inf padded). Returns 0.
    int or_knn(const double *db, std::int64_t n, const double *q, std::int64_t nq, int dim, int k, int metric, double param,
               int exclude_self, std::int64_t *idx, double *dist, int nthreads)
    {
        auto body = [&](std::int64_t b, std::int64_t e)
        {
            std::vector<std::pair<double, std::int64_t>> all;
            for (std::int64_t i = b; i < e; ++i)
            {
                all.clear();
                for (std::int64_t j = 0; j < n; ++j)
                {
                    if (exclude_self && j == i) continue;
                    const double d = metric == 0 ? rv_distance(db + j * dim, q + i * dim, dim)
                                                 : chain_distance(db + j * dim, q + i * dim, dim, param);
                    all.emplace_back(d, j);
                }
                const std::size_t kk = std::min<std::size_t>(k, all.size());
                std::partial_sort(all.begin(), all.begin() + kk, all.end());
                for (int t = 0; t < k; ++t)
                {
                    idx[i * k + t] = t < static_cast<int>(kk) ? all[t].second : -1;
                    dist[i * k + t] = t < static_cast<int>(kk) ? all[t].first : INFINITY;
                }
            }
        };
        if (nthreads <= 1)
        {
            body(0, nq);
            return 0;
        }
        std::atomic<std::int64_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
            pool.emplace_back(
                [&]()
                {
                    for (;;)
                    {
                        const std::int64_t b = next.fetch_add(16);
                        if (b >= nq) break;
                        body(b, std::min<std::int64_t>(b + 16, nq));
                    }
                });
        for (auto &t : pool) t.join();
        return 0;
    }

    // nearestR: returns count; idx/dist hold up to cap entries sorted ascending.
    std::int64_t or_knn_radius(const double *db, std::int64_t n, const double *q, int dim, double radius, int metric, double param,
                               std::int64_t *idx, double *dist, std::int64_t cap)
    {
        std::vector<std::pair<double, std::int64_t>> all;
        for (std::int64_t j = 0; j < n; ++j)
        {
            const double d = metric == 0 ? rv_distance(db + j * dim, q, dim) : chain_distance(db + j * dim, q, dim, param);
            if (d <= radius) all.emplace_back(d, j);
        }
        std::sort(all.begin(), all.end());
        for (std::size_t t = 0; t < all.size() && static_cast<std::int64_t>(t) < cap; ++t)
        {
            idx[t] = all[t].second;
            dist[t] = all[t].first;
        }
        return static_cast<std::int64_t>(all.size());
    }
}
