// Joint / motor / servo / limit constraint functions for the sm_100a solver kernels.
#pragma once
#include "bepu_device_math.cuh"

namespace BEPU_NS {

#define BEPU_JOINT_TYPES(X)

}  // namespace BEPU_NS
