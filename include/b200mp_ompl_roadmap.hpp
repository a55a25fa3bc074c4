// b200mp_ompl_roadmap.hpp -- the batched PRM roadmap of the b200mp engine in OMPL's terms: build (one or several GPUs), hand-off
// to stock OMPL through ompl::base::PlannerData, and persistence without Boost.
//
// What it connects to in the reference (paths relative to the reference checkout):
//   buildRoadmap / buildRoadmapMulti   replace the growRoadmap / addMilestone loop of ompl::geometric::PRM
//                                      (src/ompl/geometric/planners/prm/src/PRM.cpp:240-270, 570-604) by b200mp_prm_build[_multi];
//   toPlannerData                      fills an ompl::base::PlannerData (src/ompl/base/PlannerData.h:173-430) exactly the way
//                                      PRM(const PlannerData &) reads it back (PRM.cpp:96-156: getVertex(i).getState(), getEdges(i, list),
//                                      getEdgeWeight(i, j, &cost)), so `ompl::geometric::PRM prm(data)` starts from the GPU-built roadmap;
//   saveRoadmap / loadRoadmap          a versioned little-endian binary file (ompl::base::PlannerDataStorage, PlannerDataStorage.h:175-263,
//                                      needs Boost.Serialization archives, which tie the file to a Boost version; this format does not);
//   writeGraphML                       the document PlannerData::printGraphML produces (src/ompl/base/src/PlannerData.cpp:262-305:
//                                      boost::write_graphml with the properties "coords" on nodes and "weight" on edges).
// NOT the graph stock PRM grows from the same states: see the note at b200mp_prm_build in b200mp.h (every state is connected to its k
// nearest among ALL states, stock PRM only sees the milestones inserted earlier).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "b200mp_ompl.hpp"
#include <ompl/base/Cost.h>
#ifndef B200MP_OMPL_PLANNERDATA_HEADER
#define B200MP_OMPL_PLANNERDATA_HEADER <ompl/base/PlannerData.h>
#endif
#include B200MP_OMPL_PLANNERDATA_HEADER

namespace ompl::b200
{
    /// A roadmap as the engine returns it. Edge e joins milestones edges[2e] < edges[2e+1] (older -> newer, the direction PRM checks,
    /// PRM.cpp:585-598); weights[e] is its length (RealVectorStateSpace::distance, the motionCost of path-length objectives).
    struct Roadmap
    {
        std::size_t dim{0};
        std::vector<float> states;           // [n][dim], every one valid
        std::vector<std::int32_t> edges;     // [E][2]
        std::vector<std::uint8_t> valid;     // [E] 1 = collision free
        std::vector<double> weights;         // [E]
        std::vector<std::int32_t> neighbors; // [n][k] nearest-neighbour table behind the edges (only when requested)
        b200mp_roadmap_stats stats{};

        std::size_t numStates() const { return dim ? states.size() / dim : 0; }
        std::size_t numEdges() const { return valid.size(); }
        bool operator==(const Roadmap &o) const
        {
            return dim == o.dim && states == o.states && edges == o.edges && valid == o.valid && weights == o.weights;
        }
    };

    namespace detail
    {
        inline void sizeRoadmap(Roadmap &r, std::size_t dim, std::size_t n, int k, bool neighbors)
        {
            r.dim = dim;
            r.states.resize(n * dim);
            r.edges.resize(2 * n * static_cast<std::size_t>(k));
            r.valid.resize(n * static_cast<std::size_t>(k));
            r.weights.resize(n * static_cast<std::size_t>(k));
            r.neighbors.resize(neighbors && n > 1 ? n * std::min<std::size_t>(static_cast<std::size_t>(k), n - 1) : 0);
        }
        inline void trimRoadmap(Roadmap &r)
        {
            const auto e = static_cast<std::size_t>(r.stats.n_edges);
            r.edges.resize(2 * e);
            r.valid.resize(e);
            r.weights.resize(e);
        }
    }  // namespace detail

    /// n valid samples of `Robot` in `env`, each connected to its k nearest, every edge checked: one GPU (b200mp_prm_build).
    template <typename Robot>
    inline Roadmap buildRoadmap(const EnvironmentPtr &env, std::uint64_t seed, std::size_t n, int k, int resolution = 0)
    {
        Roadmap r;
        detail::sizeRoadmap(r, Robot::dimension, n, k, false);
        check(b200mp_prm_build(env->engine()->get(), Robot::id, env->get(), seed, static_cast<std::int64_t>(n), k, resolution, r.states.data(),
                               r.edges.data(), r.valid.data(), r.weights.data(), static_cast<std::int64_t>(r.valid.size()), &r.stats),
              "buildRoadmap");
        detail::trimRoadmap(r);
        return r;
    }

    /// The same roadmap built by several GPUs of this node, one thread per GPU (b200mp_comm_create_all + b200mp_prm_build_multi; NCCL
    /// inside the library). envs[r] holds the scene on GPU r's engine (Environment::cloneTo copies one across engines); one environment
    /// is the single-GPU case. The result is identical to buildRoadmap's for any number of GPUs. incremental: the graph stock PRM grows
    /// from the same states inserted in index order (B200MP_PRM_INCREMENTAL), for graph-for-graph comparison with ompl::geometric::PRM.
    template <typename Robot>
    inline Roadmap buildRoadmapMulti(const std::vector<EnvironmentPtr> &envs, std::uint64_t seed, std::size_t n, int k, int resolution = 0,
                                     bool wantNeighbors = false, bool incremental = false)
    {
        const int g = static_cast<int>(envs.size());
        if (g < 1)
            throw ompl::Exception("b200mp::buildRoadmapMulti", "no environments");
        std::vector<b200mp_ctx *> ctxs(g);
        for (int r = 0; r < g; ++r)
            ctxs[r] = envs[r]->engine()->get();
        std::vector<b200mp_comm *> comms(g, nullptr);
        check(b200mp_comm_create_all(ctxs.data(), g, comms.data()), "buildRoadmapMulti");
        Roadmap out;
        detail::sizeRoadmap(out, Robot::dimension, n, k, wantNeighbors);
        std::vector<int> rc(g, B200MP_OK);
        std::vector<std::string> err(g);
        std::vector<std::thread> pool;
        for (int r = 0; r < g; ++r)
            pool.emplace_back([&, r]() {
                // rank 0 receives the roadmap; the other ranks take part in every exchange and drop their copy
                const bool root = r == 0;
                b200mp_roadmap_stats st{};
                rc[r] = b200mp_prm_build_multi(comms[r], Robot::id, envs[r]->get(), seed, static_cast<std::int64_t>(n), k, resolution,
                                               incremental ? B200MP_PRM_INCREMENTAL : 0, root ? out.states.data() : nullptr, root ? out.edges.data() : nullptr,
                                               root ? out.valid.data() : nullptr, root ? out.weights.data() : nullptr,
                                               root && wantNeighbors ? out.neighbors.data() : nullptr, static_cast<std::int64_t>(out.valid.size()),
                                               &st);
                if (rc[r] != B200MP_OK)
                    err[r] = b200mp_last_error();
                if (root)
                    out.stats = st;
            });
        for (auto &t : pool)
            t.join();
        for (auto *c : comms)
            b200mp_comm_destroy(c);
        for (int r = 0; r < g; ++r)
            if (rc[r] != B200MP_OK)
                throw ompl::Exception("b200mp::buildRoadmapMulti", err[r]);
        detail::trimRoadmap(out);
        return out;
    }

    /// Owns the ompl::base::State objects a PlannerData made from a Roadmap points to (PlannerData stores pointers; PRM(const PlannerData &)
    /// clones what it reads, PRM.cpp:108, so the block only has to outlive that constructor -- or call data.decoupleFromPlanner()).
    class RoadmapStates
    {
    public:
        RoadmapStates(const Roadmap &r) : values_(r.states.begin(), r.states.end()), states_(r.numStates())
        {
            for (std::size_t i = 0; i < states_.size(); ++i)
                states_[i].values = values_.data() + i * r.dim;  // ompl_to_vamp's cast, reversed: float -> double is exact
        }
        RoadmapStates(const RoadmapStates &) = delete;
        RoadmapStates &operator=(const RoadmapStates &) = delete;
        const ob::State *state(std::size_t i) const { return &states_[i]; }
        std::size_t size() const { return states_.size(); }

    private:
        std::vector<double> values_;
        std::vector<ob::RealVectorStateSpace::StateType> states_;
    };

    /// Milestones and collision-free edges of `r` into `data`: vertex i of the roadmap becomes the i-th vertex added, every valid edge one
    /// directed PlannerData edge older -> newer with Cost(length). Returns the number of edges added.
    inline std::size_t toPlannerData(const Roadmap &r, const RoadmapStates &states, ob::PlannerData &data)
    {
        std::vector<unsigned int> id(r.numStates());
        for (std::size_t i = 0; i < r.numStates(); ++i)
            id[i] = data.addVertex(ob::PlannerDataVertex(states.state(i)));
        std::size_t added = 0;
        for (std::size_t e = 0; e < r.numEdges(); ++e)
            if (r.valid[e] && data.addEdge(id[r.edges[2 * e]], id[r.edges[2 * e + 1]], ob::PlannerDataEdge(), ob::Cost(r.weights[e])))
                ++added;
        return added;
    }

    // ---- persistence: "B2RM" | u32 version | u32 dim | u64 n | u64 E | u32 k (0 = no table) | states f32 | edges i32 | valid u8 | weights f64
    //      | neighbours i32; little endian, no padding
    inline void saveRoadmap(const std::string &path, const Roadmap &r)
    {
        std::FILE *f = std::fopen(path.c_str(), "wb");
        if (!f)
            throw ompl::Exception("b200mp::saveRoadmap", "cannot open " + path);
        const std::uint32_t version = 1, dim = static_cast<std::uint32_t>(r.dim);
        const std::uint64_t n = r.numStates(), e = r.numEdges();
        const std::uint32_t k = n ? static_cast<std::uint32_t>(r.neighbors.size() / n) : 0;
        bool ok = std::fwrite("B2RM", 1, 4, f) == 4 && std::fwrite(&version, 4, 1, f) == 1 && std::fwrite(&dim, 4, 1, f) == 1 &&
                  std::fwrite(&n, 8, 1, f) == 1 && std::fwrite(&e, 8, 1, f) == 1 && std::fwrite(&k, 4, 1, f) == 1;
        auto put = [&](const void *p, std::size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
        put(r.states.data(), r.states.size() * sizeof(float));
        put(r.edges.data(), r.edges.size() * sizeof(std::int32_t));
        put(r.valid.data(), r.valid.size());
        put(r.weights.data(), r.weights.size() * sizeof(double));
        put(r.neighbors.data(), r.neighbors.size() * sizeof(std::int32_t));
        ok = (std::fclose(f) == 0) && ok;
        if (!ok)
            throw ompl::Exception("b200mp::saveRoadmap", "write to " + path + " failed");
    }

    inline Roadmap loadRoadmap(const std::string &path)
    {
        std::FILE *f = std::fopen(path.c_str(), "rb");
        if (!f)
            throw ompl::Exception("b200mp::loadRoadmap", "cannot open " + path);
        char magic[4];
        std::uint32_t version = 0, dim = 0, k = 0;
        std::uint64_t n = 0, e = 0;
        bool ok = std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "B2RM", 4) == 0 && std::fread(&version, 4, 1, f) == 1 &&
                  version == 1 && std::fread(&dim, 4, 1, f) == 1 && std::fread(&n, 8, 1, f) == 1 && std::fread(&e, 8, 1, f) == 1 &&
                  std::fread(&k, 4, 1, f) == 1 && dim >= 1 && dim <= 64 && n < (1ull << 40) && e < (1ull << 40);
        Roadmap r;
        if (ok)
        {
            r.dim = dim;
            r.states.resize(n * dim);
            r.edges.resize(2 * e);
            r.valid.resize(e);
            r.weights.resize(e);
            r.neighbors.resize(n * k);
            auto get = [&](void *p, std::size_t bytes) { ok = ok && (bytes == 0 || std::fread(p, 1, bytes, f) == bytes); };
            get(r.states.data(), r.states.size() * sizeof(float));
            get(r.edges.data(), r.edges.size() * sizeof(std::int32_t));
            get(r.valid.data(), r.valid.size());
            get(r.weights.data(), r.weights.size() * sizeof(double));
            get(r.neighbors.data(), r.neighbors.size() * sizeof(std::int32_t));
            for (std::size_t i = 0; ok && i < r.edges.size(); ++i)
                ok = r.edges[i] >= 0 && static_cast<std::uint64_t>(r.edges[i]) < n;
        }
        std::fclose(f);
        if (!ok)
            throw ompl::Exception("b200mp::loadRoadmap", path + " is not a version-1 b200mp roadmap file (or is truncated)");
        r.stats.n_states = static_cast<std::int64_t>(n);
        r.stats.n_edges = static_cast<std::int64_t>(e);
        for (auto v : r.valid)
            r.stats.n_valid_edges += v;
        return r;
    }

    /// The GraphML document of PlannerData::printGraphML for the roadmap's collision-free part: nodes n<i> with the comma-separated
    /// coordinates in "coords" (default stream precision, as vertexCoords writes them), directed edges with their length in "weight"
    /// (17 significant digits, boost::lexical_cast's rendering of a double).
    inline void writeGraphML(std::ostream &out, const Roadmap &r)
    {
        out << "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n"
               "<graphml xmlns=\"http://graphml.graphdrawing.org/xmlns\" xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\" "
               "xsi:schemaLocation=\"http://graphml.graphdrawing.org/xmlns http://graphml.graphdrawing.org/xmlns/1.0/graphml.xsd\">\n"
               "  <key id=\"key0\" for=\"node\" attr.name=\"coords\" attr.type=\"string\" />\n"
               "  <key id=\"key1\" for=\"edge\" attr.name=\"weight\" attr.type=\"double\" />\n"
               "  <graph id=\"G\" edgedefault=\"directed\" parse.nodeids=\"free\" parse.edgeids=\"canonical\" parse.order=\"nodesfirst\">\n";
        for (std::size_t i = 0; i < r.numStates(); ++i)
        {
            std::ostringstream c;
            for (std::size_t j = 0; j < r.dim; ++j)
                c << (j ? "," : "") << static_cast<double>(r.states[i * r.dim + j]);
            out << "    <node id=\"n" << i << "\">\n      <data key=\"key0\">" << c.str() << "</data>\n    </node>\n";
        }
        std::size_t id = 0;
        char w[40];
        for (std::size_t e = 0; e < r.numEdges(); ++e)
        {
            if (!r.valid[e])
                continue;
            std::snprintf(w, sizeof(w), "%.17g", r.weights[e]);
            out << "    <edge id=\"e" << id++ << "\" source=\"n" << r.edges[2 * e] << "\" target=\"n" << r.edges[2 * e + 1]
                << "\">\n      <data key=\"key1\">" << w << "</data>\n    </edge>\n";
        }
        out << "  </graph>\n</graphml>\n";
    }
}  // namespace ompl::b200
