// TEST MOCK (used only where the reference checkout is absent): same macro contract as ompl/util/ClassForward.h.
#pragma once
#include <memory>
#define OMPL_CLASS_FORWARD(C) \
    class C;                  \
    using C##Ptr = std::shared_ptr<C>
