// TEST MOCK: exhaustive-search container with the semantics of ompl/datastructures/NearestNeighborsLinear.h (ground
// truth of the reference's own nearest-neighbour test); used only where the reference checkout is absent.
#pragma once
#include <algorithm>
#include "ompl/datastructures/NearestNeighbors.h"
#include "ompl/util/Exception.h"
namespace ompl
{
    template <typename _T>
    class NearestNeighborsLinear : public NearestNeighbors<_T>
    {
    public:
        bool reportsSortedResults() const override { return true; }
        void clear() override { items_.clear(); }
        void add(const _T &d) override { items_.push_back(d); }
        void add(const std::vector<_T> &d) override { items_.insert(items_.end(), d.begin(), d.end()); }
        bool remove(const _T &d) override
        {
            for (std::size_t i = items_.size(); i-- > 0;)
                if (items_[i] == d) { items_.erase(items_.begin() + i); return true; }
            return false;
        }
        _T nearest(const _T &q) const override
        {
            if (items_.empty()) throw Exception("No elements found in nearest neighbors data structure");
            return *std::min_element(items_.begin(), items_.end(), [&](const _T &a, const _T &b) { return this->distFun_(a, q) < this->distFun_(b, q); });
        }
        void nearestK(const _T &q, std::size_t k, std::vector<_T> &out) const override
        {
            out = items_;
            std::stable_sort(out.begin(), out.end(), [&](const _T &a, const _T &b) { return this->distFun_(a, q) < this->distFun_(b, q); });
            if (out.size() > k) out.resize(k);
        }
        void nearestR(const _T &q, double r, std::vector<_T> &out) const override
        {
            out.clear();
            for (const auto &e : items_) if (this->distFun_(e, q) <= r) out.push_back(e);
            std::stable_sort(out.begin(), out.end(), [&](const _T &a, const _T &b) { return this->distFun_(a, q) < this->distFun_(b, q); });
        }
        std::size_t size() const override { return items_.size(); }
        void list(std::vector<_T> &out) const override { out = items_; }
    private:
        std::vector<_T> items_;
    };
}  // namespace ompl
