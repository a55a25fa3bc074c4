// TEST SHIM: see ompl/base/spaces/RealVectorStateSpace.h in this directory.
#pragma once
#include "ompl/base/spaces/RealVectorStateSpace.h"
