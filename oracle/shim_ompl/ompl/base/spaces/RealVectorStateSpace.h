// TEST SHIM (oracle build only): the slice of ompl::base::{StateSpace, RealVectorStateSpace, RealVectorBounds,
// ProjectionEvaluator, SpaceInformation} that the reference's demos/KinematicChain.h and
// src/ompl/base/src/DiscreteMotionValidator.cpp touch, so those two UNMODIFIED reference files compile without
// Boost/Eigen. The genuine State.h / StateValidityChecker.h / MotionValidator.h / DiscreteMotionValidator.h are used.
// Behavioural one-liners restated here: StateSpace::validSegmentCount (StateSpace.cpp:851-854), setup's
// longestValidSegment_ = maxExtent * fraction (StateSpace.cpp:236-244), RealVectorStateSpace::getMaximumExtent
// (RealVectorStateSpace.cpp:167-176), allocState/freeState (:272-284).
#pragma once
#include <cmath>
#include <memory>
#include <vector>
#include "ompl/base/State.h"
#include "ompl/base/StateValidityChecker.h"
#include "ompl/util/ClassForward.h"

namespace Eigen
{
    struct VectorXd { std::vector<double> v; };
    template <typename T> struct Ref { T *p; };
}

namespace ompl::base
{
    OMPL_CLASS_FORWARD(StateSpace);

    class RealVectorBounds
    {
    public:
        explicit RealVectorBounds(unsigned int dim) : low(dim, 0.0), high(dim, 0.0) {}
        void setLow(double v) { std::fill(low.begin(), low.end(), v); }
        void setHigh(double v) { std::fill(high.begin(), high.end(), v); }
        std::vector<double> low, high;
    };

    class ProjectionMatrix
    {
    public:
        struct Mat { std::size_t r = 0; std::size_t rows() const { return r; } } mat;
        void computeRandom(unsigned int, unsigned int to) { mat.r = to; }
        void project(const double *, Eigen::Ref<Eigen::VectorXd>) const {}
    };

    class ProjectionEvaluator
    {
    public:
        explicit ProjectionEvaluator(const StateSpace *space) : space_(space) {}
        virtual ~ProjectionEvaluator() = default;
        virtual unsigned int getDimension() const = 0;
        virtual void project(const State *state, Eigen::Ref<Eigen::VectorXd> projection) const = 0;
    protected:
        const StateSpace *space_;
    };

    class StateSpace
    {
    public:
        virtual ~StateSpace() = default;
        template <class T> T *as() { return static_cast<T *>(this); }
        template <class T> const T *as() const { return static_cast<const T *>(this); }
        virtual unsigned int getDimension() const = 0;
        virtual double getMaximumExtent() const = 0;
        virtual double distance(const State *a, const State *b) const = 0;
        virtual void interpolate(const State *from, const State *to, double t, State *state) const = 0;
        virtual State *allocState() const = 0;
        virtual void freeState(State *s) const = 0;
        virtual void copyToReals(std::vector<double> &reals, const State *source) const = 0;
        virtual void registerProjections() {}
        void registerDefaultProjection(const std::shared_ptr<ProjectionEvaluator> &) {}
        void setLongestValidSegmentFraction(double f) { longestValidSegmentFraction_ = f; }
        void setup() { longestValidSegment_ = getMaximumExtent() * longestValidSegmentFraction_; }
        double getLongestValidSegmentLength() const { return longestValidSegment_; }
        virtual unsigned int validSegmentCount(const State *state1, const State *state2) const
        {
            return longestValidSegmentCountFactor_ * (unsigned int)ceil(distance(state1, state2) / longestValidSegment_);
        }
    protected:
        double longestValidSegment_{0.0};
        double longestValidSegmentFraction_{0.01};
        unsigned int longestValidSegmentCountFactor_{1};
    };

    class RealVectorStateSpace : public StateSpace
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
        explicit RealVectorStateSpace(unsigned int dim = 0) : dimension_(dim), bounds_(dim) {}
        void setBounds(const RealVectorBounds &b) { bounds_ = b; }
        const RealVectorBounds &getBounds() const { return bounds_; }
        unsigned int getDimension() const override { return dimension_; }
        double getMaximumExtent() const override
        {
            double e = 0.0;
            for (unsigned int i = 0; i < dimension_; ++i)
            {
                double d = bounds_.high[i] - bounds_.low[i];
                e += d * d;
            }
            return sqrt(e);
        }
        State *allocState() const override
        {
            auto *rstate = new StateType();
            rstate->values = new double[dimension_];
            return rstate;
        }
        void freeState(State *state) const override
        {
            auto *rstate = static_cast<StateType *>(state);
            delete[] rstate->values;
            delete rstate;
        }
        void copyToReals(std::vector<double> &reals, const State *source) const override
        {
            reals.assign(source->as<StateType>()->values, source->as<StateType>()->values + dimension_);
        }
        virtual void enforceBounds(State *) const {}
        virtual bool equalStates(const State *, const State *) const { return false; }
    protected:
        unsigned int dimension_;
        RealVectorBounds bounds_;
    };

    class SpaceInformation
    {
    public:
        explicit SpaceInformation(StateSpacePtr space) : space_(std::move(space)) {}
        const StateSpacePtr &getStateSpace() const { return space_; }
        unsigned int getStateDimension() const { return space_->getDimension(); }
        void setStateValidityChecker(const StateValidityCheckerPtr &svc) { svc_ = svc; }
        bool isValid(const State *s) const { return svc_->isValid(s); }
        State *allocState() const { return space_->allocState(); }
        void freeState(State *s) const { space_->freeState(s); }
    private:
        StateSpacePtr space_;
        StateValidityCheckerPtr svc_;
    };
}  // namespace ompl::base
