// rviz / tf publishing is ROS plumbing and out of scope; this stub keeps the members that
// VO and Map hold (reference: include/stereo_visual_slam_main/visualization.hpp:19-73) so their layouts and
// constructors stay the same.
#ifndef VSLAM_B200_VISUALIZATION_HPP
#define VSLAM_B200_VISUALIZATION_HPP

#include <stereo_visual_slam_main/library_include.hpp>
#include <stereo_visual_slam_main/types_def.hpp>

namespace vslam {

class VslamVisual {
public:
    VslamVisual() = default;
    explicit VslamVisual(ros::NodeHandle&) {}
    int publish_feature_map(const std::vector<cv::Point3f>&) { return 0; }
    int publish_transform(const SE3&) { return 0; }
    void publish_fixed_pose(const Frame&) {}
};

}  // namespace vslam

#endif
