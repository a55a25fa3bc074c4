// TEST MOCK of ompl/base/Cost.h (interface subset: value wrapper).
#pragma once
#include <iostream>
namespace ompl::base
{
    class Cost
    {
    public:
        explicit Cost(double v = 0.0) : v_(v) {}
        double value() const { return v_; }

    private:
        double v_;
    };
    inline std::ostream &operator<<(std::ostream &o, const Cost &c) { return o << c.value(); }
}  // namespace ompl::base
