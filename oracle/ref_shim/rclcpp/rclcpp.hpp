// ORACLE build shim (test infrastructure).  The reference's ORBextractor.hpp:25 includes
// rclcpp/rclcpp.hpp but uses nothing from it; an empty header satisfies it.
#pragma once
