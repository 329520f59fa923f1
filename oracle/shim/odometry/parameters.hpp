// stand-in for ../odometry/parameters.hpp (absent parent-project header)
#pragma once
#include "../codegen/output/cmd_parameters.hpp"
namespace odometry {
using ParametersSlam = cmd::ParametersSlam;
using ParametersTracker = cmd::ParametersTracker;
struct Parameters {
    ParametersSlam slam;
    ParametersTracker tracker;
};
}  // namespace odometry
