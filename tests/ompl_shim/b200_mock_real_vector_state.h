// TEST SHIM: layout-compatible stand-in for ompl::base::RealVectorStateSpace::StateType (a State with a `double *values`
// member, ompl/base/spaces/RealVectorStateSpace.h:77-97). The genuine header needs Boost and Eigen, which this image lacks.
#pragma once
#include <ompl/base/State.h>
namespace ompl::base
{
    class RealVectorStateSpace
    {
    public:
        class StateType : public State
        {
        public:
            StateType() = default;
            double operator[](unsigned int i) const { return values[i]; }
            double &operator[](unsigned int i) { return values[i]; }
            double *values;
        };
    };
}  // namespace ompl::base
