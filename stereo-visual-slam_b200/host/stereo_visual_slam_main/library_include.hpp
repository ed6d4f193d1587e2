// Umbrella include of the drop-in host layer.  Same role and include path as the reference's
// include/stereo_visual_slam_main/library_include.hpp (ROS + Eigen + Sophus + OpenCV), but resolving to the minimal
// stand-ins in ../compat/vslam_compat.hpp unless VSLAM_USE_REAL_DEPS is defined.
#ifndef VSLAM_B200_LIBRARY_INCLUDE_HPP
#define VSLAM_B200_LIBRARY_INCLUDE_HPP

#include "../compat/vslam_compat.hpp"

typedef Sophus::SE3d SE3;  // reference: library_include.hpp:18
typedef Sophus::SO3d SO3;  // reference: library_include.hpp:19

#endif
