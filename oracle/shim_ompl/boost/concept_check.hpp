// TEST SHIM: the only Boost facility the OMPL base-class headers used by the adapters need (ompl/base/State.h).
#pragma once
#include <type_traits>
namespace boost
{
    template <class From, class To>
    struct Convertible
    {
        static_assert(std::is_convertible<From, To>::value, "not convertible");
    };
}  // namespace boost
#define BOOST_CONCEPT_ASSERT(x) static_assert(true, "")
