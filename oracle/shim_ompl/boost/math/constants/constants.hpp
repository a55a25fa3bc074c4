// TEST SHIM: boost::math::constants::pi<T>() for compiling the reference's demos/KinematicChain.h without Boost.
#pragma once
namespace boost::math::constants
{
    template <typename T>
    constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
}
