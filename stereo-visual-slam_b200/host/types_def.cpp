// Frame helpers (reference: src/stereo_visual_slam_main/types_def.cpp:9-28).
#include <stereo_visual_slam_main/types_def.hpp>

namespace vslam {

Eigen::Vector3d Frame::find_3d(const cv::KeyPoint& kp, Eigen::Vector3d& relative_pt3d) {
    // the disparity is looked up at the TRUNCATED pixel, like cv::Mat::at<float>(float, float) in the reference
    const float d = disparity_.at<float>((int)kp.pt.y, (int)kp.pt.x);
    const double depth = fx_ * b_ / d;
    relative_pt3d = Eigen::Vector3d((kp.pt.x - cx_) / fx_ * depth, (kp.pt.y - cy_) / fy_ * depth, depth);
    return T_c_w_.inverse() * relative_pt3d;
}

void Frame::fill_frame(SE3 T_c_w, bool is_keyframe, int keyframe_id) {
    T_c_w_ = T_c_w;
    is_keyframe_ = is_keyframe;
    if (is_keyframe) keyframe_id_ = keyframe_id;
}

}  // namespace vslam
