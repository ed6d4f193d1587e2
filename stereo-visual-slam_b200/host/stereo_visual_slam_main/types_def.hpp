// Frame / Feature / Observation / Landmark -- the data contract of the hot path.
// Field names, types, defaults and order mirror /root/reference/include/stereo_visual_slam_main/types_def.hpp:17-121
// one for one (the structs ARE the drop-in boundary); everything around them is new.
#ifndef VSLAM_B200_TYPES_DEF_HPP
#define VSLAM_B200_TYPES_DEF_HPP

#include <stereo_visual_slam_main/library_include.hpp>

#include <vector>

namespace vslam {

struct Frame;
struct Landmark;
struct Feature;

// One 2-D observation of a frame.  descriptor_ is a 1x32 CV_8U row VIEW into the frame's descriptor matrix.
struct Feature {
    int feature_id_, frame_id_, landmark_id_ = -1;  // index in the frame / owning frame / landmark (-1: none yet)
    cv::KeyPoint keypoint_;
    cv::Mat descriptor_;
    bool is_inlier = false;

    Feature() = default;
    Feature(int feature_id, int frame_id, cv::KeyPoint keypoint, cv::Mat descriptor);
};

struct Frame {
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;

    int frame_id_;
    cv::Mat left_img_, right_img_, disparity_;  // CV_8U pair; CV_32F disparity, -1 where no depth is known
    SE3 T_c_w_ = SE3();
    bool is_keyframe_;
    int keyframe_id_;
    std::vector<Feature> features_;
    // KITTI-00 intrinsics and baseline, hard-coded as in the reference
    double fx_ = 718.856, fy_ = 718.856, cx_ = 607.1928, cy_ = 185.2157, b_ = 0.573;

    Frame() = default;
    Frame(int frame_id, double /*timestamp*/, const cv::Mat& left, const cv::Mat& right);

    // world position of a keypoint from disparity_ (float->int truncation of the pixel, as the reference);
    // relative_pt3d receives the camera-frame point
    Eigen::Vector3d find_3d(const cv::KeyPoint& kp, Eigen::Vector3d& relative_pt3d);
    void fill_frame(SE3 T_c_w, bool is_keyframe, int keyframe_id);
};

struct Observation {
    int keyframe_id_, feature_id_;
    bool to_delete = false;

    Observation(int keyframe_id, int feature_id);
};

struct Landmark {
    int landmark_id_;
    cv::Point3f pt_3d_;  // world frame, float32 (BA widens it, optimization.cpp:284 narrows it back)
    cv::Mat descriptor_;
    int observed_times_ = 1;
    std::vector<Observation> observations_;
    bool is_inlier = true, reliable_depth_ = false;

    Landmark() = default;
    Landmark(int landmark_id, cv::Point3f pt_3d, cv::Mat descriptor, bool reliable_depth, Observation observation);
    Eigen::Vector3d to_vector_3d();
};

}  // namespace vslam

#endif
