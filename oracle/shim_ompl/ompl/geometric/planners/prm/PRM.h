// TEST SHIM: the chain oracle does not run planners; only the state space and validity checker of the demo header are used.
#pragma once
#include "ompl/base/SpaceInformation.h"
