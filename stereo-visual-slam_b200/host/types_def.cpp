// Frame helpers (reference: src/stereo_visual_slam_main/types_def.cpp:9-28).
#include <stereo_visual_slam_main/types_def.hpp>

namespace vslam {

Feature::Feature(int feature_id, int frame_id, cv::KeyPoint keypoint, cv::Mat descriptor)
    : feature_id_(feature_id), frame_id_(frame_id), keypoint_(keypoint), descriptor_(descriptor) {}

Frame::Frame(int frame_id, double, const cv::Mat& left, const cv::Mat& right)
    : frame_id_(frame_id), left_img_(left), right_img_(right) {}

Observation::Observation(int keyframe_id, int feature_id) : keyframe_id_(keyframe_id), feature_id_(feature_id) {}

Landmark::Landmark(int landmark_id, cv::Point3f pt_3d, cv::Mat descriptor, bool reliable_depth, Observation observation)
    : landmark_id_(landmark_id), pt_3d_(pt_3d), descriptor_(descriptor), reliable_depth_(reliable_depth) {
    observations_.push_back(observation);
}

Eigen::Vector3d Landmark::to_vector_3d() { return Eigen::Vector3d(pt_3d_.x, pt_3d_.y, pt_3d_.z); }

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
