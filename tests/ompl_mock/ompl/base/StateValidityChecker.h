// TEST MOCK: the virtual interface of ompl/base/StateValidityChecker.h:90-173 that adapters must satisfy.
#pragma once
#include "ompl/base/State.h"
#include "ompl/util/ClassForward.h"
namespace ompl::base
{
    OMPL_CLASS_FORWARD(SpaceInformation);
    OMPL_CLASS_FORWARD(StateValidityChecker);
    struct StateValidityCheckerSpecs
    {
        enum ClearanceComputationType { NONE = 0, EXACT, APPROXIMATE, BOUNDED_APPROXIMATE };
        ClearanceComputationType clearanceComputationType{NONE};
        bool hasValidDirectionComputation{false};
    };
    class StateValidityChecker
    {
    public:
        StateValidityChecker(SpaceInformation *si) : si_(si) {}
        StateValidityChecker(const SpaceInformationPtr &si) : si_(si.get()) {}
        virtual ~StateValidityChecker() = default;
        virtual bool isValid(const State *state) const = 0;
        virtual bool isValid(const State *state, double &dist) const { dist = clearance(state); return isValid(state); }
        virtual double clearance(const State *) const { return 0.0; }
        const StateValidityCheckerSpecs &getSpecs() const { return specs_; }
    protected:
        SpaceInformation *si_;
        StateValidityCheckerSpecs specs_;
    };
}  // namespace ompl::base
