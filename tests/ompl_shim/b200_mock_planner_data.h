// TEST SHIM: interface-identical stand-in for ompl::base::PlannerData / PlannerDataVertex / PlannerDataEdge
// (ompl/base/PlannerData.h:56-170, 173-430 of the reference). The genuine header includes SpaceInformation.h, which needs Boost
// and Eigen; neither exists in this image. Only the members the roadmap hand-off and PRM(const PlannerData &) (PRM.cpp:96-156)
// use are provided, with the reference's signatures and semantics (directed edges, duplicate edges and self loops refused,
// vertices unique by state pointer).
#pragma once
#include <map>
#include <vector>

#include <ompl/base/Cost.h>
#include <ompl/base/State.h>

namespace ompl::base
{
    class PlannerDataVertex
    {
    public:
        PlannerDataVertex(const State *st, int tag = 0) : state_(st), tag_(tag) {}
        virtual ~PlannerDataVertex() = default;
        virtual int getTag() const { return tag_; }
        virtual const State *getState() const { return state_; }
        virtual bool operator==(const PlannerDataVertex &rhs) const { return state_ == rhs.state_; }

    protected:
        const State *state_;
        int tag_;
    };

    class PlannerDataEdge
    {
    public:
        PlannerDataEdge() = default;
        virtual ~PlannerDataEdge() = default;
    };

    class PlannerData
    {
    public:
        static constexpr unsigned int INVALID_INDEX = 0xffffffffu;

        unsigned int addVertex(const PlannerDataVertex &st)
        {
            auto it = index_.find(st.getState());
            if (it != index_.end())
                return it->second;
            const auto id = static_cast<unsigned int>(vertices_.size());
            vertices_.push_back(st);
            out_.emplace_back();
            index_[st.getState()] = id;
            return id;
        }
        virtual bool addEdge(unsigned int v1, unsigned int v2, const PlannerDataEdge & = PlannerDataEdge(), Cost weight = Cost(1.0))
        {
            if (v1 >= vertices_.size() || v2 >= vertices_.size() || v1 == v2 || out_[v1].count(v2))
                return false;
            out_[v1][v2] = weight;
            ++n_edges_;
            return true;
        }
        unsigned int numVertices() const { return static_cast<unsigned int>(vertices_.size()); }
        unsigned int numEdges() const { return n_edges_; }
        const PlannerDataVertex &getVertex(unsigned int i) const { return vertices_.at(i); }
        unsigned int vertexIndex(const PlannerDataVertex &v) const
        {
            auto it = index_.find(v.getState());
            return it == index_.end() ? INVALID_INDEX : it->second;
        }
        unsigned int getEdges(unsigned int v, std::vector<unsigned int> &edgeList) const
        {
            edgeList.clear();
            for (const auto &kv : out_.at(v))
                edgeList.push_back(kv.first);
            return static_cast<unsigned int>(edgeList.size());
        }
        bool getEdgeWeight(unsigned int v1, unsigned int v2, Cost *weight) const
        {
            auto it = out_.at(v1).find(v2);
            if (it == out_.at(v1).end())
                return false;
            *weight = it->second;
            return true;
        }

    private:
        std::vector<PlannerDataVertex> vertices_;
        std::vector<std::map<unsigned int, Cost>> out_;
        std::map<const State *, unsigned int> index_;
        unsigned int n_edges_{0};
    };
}  // namespace ompl::base
