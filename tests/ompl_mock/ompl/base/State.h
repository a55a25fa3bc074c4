// TEST MOCK: abstract state with the as<T>() casts of ompl/base/State.h.
#pragma once
namespace ompl::base
{
    class State
    {
    public:
        State(const State &) = delete;
        const State &operator=(const State &) = delete;
        template <class T> const T *as() const { return static_cast<const T *>(this); }
        template <class T> T *as() { return static_cast<T *>(this); }
    protected:
        State() = default;
        virtual ~State() = default;
    };
}  // namespace ompl::base
