// TEST MOCK: the virtual interface of ompl/base/MotionValidator.h:64-140 that adapters must satisfy.
#pragma once
#include <utility>
#include "ompl/base/State.h"
#include "ompl/util/ClassForward.h"
namespace ompl::base
{
    OMPL_CLASS_FORWARD(SpaceInformation);
    OMPL_CLASS_FORWARD(MotionValidator);
    class MotionValidator
    {
    public:
        MotionValidator(SpaceInformation *si) : si_(si), valid_(0), invalid_(0) {}
        MotionValidator(const SpaceInformationPtr &si) : si_(si.get()), valid_(0), invalid_(0) {}
        virtual ~MotionValidator() = default;
        virtual bool checkMotion(const State *s1, const State *s2) const = 0;
        virtual bool checkMotion(const State *s1, const State *s2, std::pair<State *, double> &lastValid) const = 0;
        unsigned int getValidMotionCount() const { return valid_; }
        unsigned int getInvalidMotionCount() const { return invalid_; }
    protected:
        SpaceInformation *si_;
        mutable unsigned int valid_;
        mutable unsigned int invalid_;
    };
}  // namespace ompl::base
