// Sliding-window map: 10 keyframes + the landmark table.  Same public surface as the reference's
// include/stereo_visual_slam_main/map.hpp:15-81 (host bookkeeping on <= 10 frames; no kernels involved).
#ifndef VSLAM_B200_MAP_HPP
#define VSLAM_B200_MAP_HPP

#include <stereo_visual_slam_main/library_include.hpp>
#include <stereo_visual_slam_main/types_def.hpp>
#include <stereo_visual_slam_main/visualization.hpp>

#include <unordered_map>

namespace vslam {

struct Map {
    std::unordered_map<unsigned long, Frame> keyframes_;
    std::unordered_map<unsigned long, Landmark> landmarks_;

    // window size.  The reference fixes it at 10 (`const int`, map.hpp:22); here it can be raised up to the BA kernels'
    // limit of 64 keyframes before the first frame.
    int num_keyframes_ = 10;
    int current_keyframe_id_ = 0;

    VslamVisual my_visual_;
    bool if_write_pose_ = false, if_rviz_ = false;  // parameter server: /if_write_pose, /if_rviz

    explicit Map(ros::NodeHandle& nh);

    // window maintenance (all return 0, like the reference)
    int insert_keyframe(Frame frame_to_add);
    int insert_landmark(Landmark landmark_to_add);
    int remove_keyframe();
    int clean_map();
    // output
    void publish_keyframes() {}  // rviz only in the reference
    void write_pose(const Frame& frame);
    void write_remaining_pose();
};

}  // namespace vslam

#endif
