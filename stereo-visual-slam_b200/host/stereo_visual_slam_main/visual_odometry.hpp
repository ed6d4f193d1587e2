// Stereo visual odometry front end with the reference's class surface
// (/root/reference/include/stereo_visual_slam_main/visual_odometry.hpp:27-185).  Every compute stage forwards to the
// CUDA library through include/vslam_b200.h; the state machine around the stages is host C++ as in the reference.
#ifndef VSLAM_B200_VISUAL_ODOMETRY_HPP
#define VSLAM_B200_VISUAL_ODOMETRY_HPP

#include <stereo_visual_slam_main/library_include.hpp>
#include <stereo_visual_slam_main/map.hpp>
#include <stereo_visual_slam_main/optimization.hpp>
#include <stereo_visual_slam_main/types_def.hpp>
#include <stereo_visual_slam_main/visualization.hpp>

#include <functional>
#include <string>
#include <vector>

struct vslam_ctx;

namespace vslam {

enum TrackState { Init, Track, Lost };

class VO {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    Frame frame_last_;
    Frame frame_current_;
    Map& my_map_;

    std::string dataset_;
    // The reference holds cv::Ptr<ORB> x2 and a cv::Ptr<BFMatcher> here (visual_odometry.hpp:37-39).  Their
    // parameters are the only state they carry; the kernels live in the context.
    int detector_nfeatures_ = 3000;  // cv::ORB::create(3000), visual_odometry.cpp:22,31
    int anms_keep_ = 500;            // adaptive_non_maximal_suppresion(keypoints, 500), visual_odometry.cpp:82
    vslam_ctx* ctx_ = nullptr;       // owned
    // depth source of disparity_map: true (default) = the reference's dense StereoSGBM (visual_odometry.cpp:163-168)
    // on the GPU, bit-exact; false = opt-in fast mode, sparse stereo (ORB on both images + L<->R matching on the same
    // row band with positive disparity + DLT, the north star's triangulation) -- NOT what the reference computes
    bool dense_stereo_ = true;

    int num_inliers_ = 0;
    SE3 T_c_l_ = SE3();
    SE3 T_c_w_ = SE3();
    int seq_ = 1;

    VslamVisual my_visual_;

    TrackState state_ = Init;
    int num_lost_ = 0;
    int curr_keyframe_id_ = 0;
    int curr_landmark_id_ = 0;
    bool if_rviz_ = false;

    // image source: the reference reads KITTI PNGs with cv::imread (file I/O, out of scope).  read_img() calls this
    // hook when set, else reads binary PGM files dataset_/image_{0,1}/%06d.pgm.
    std::function<int(int, cv::Mat&, cv::Mat&)> image_source_;

public:
    VO(ros::NodeHandle& nh, Map& map);
    VO(std::string dataset, ros::NodeHandle& nh, Map& map);
    ~VO();
    VO(const VO&) = delete;
    VO& operator=(const VO&) = delete;

    int read_img(int id, cv::Mat& left_img, cv::Mat& right_img);
    // reference: dense SGBM (the default, dense_stereo_ = true, reproduces it bit-exactly).  With dense_stereo_ = false:
    // SPARSE disparity -- ORB on both images, L<->R feature_matching, matches gated to |dy| <= 2 px and positive
    // disparity, per-match DLT; `disparity` is CV_32F, -1 everywhere except at the (truncated) pixel of every matched
    // left keypoint, where it holds fx*b/Z, so Frame::find_3d and set_ref_3d_position work unchanged in both modes.
    int disparity_map(const Frame& frame, cv::Mat& disparity);
    bool initialization();
    bool tracking(bool& if_insert_keyframe);
    int feature_detection(const cv::Mat& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors);
    int feature_matching(const cv::Mat& descriptors_1, const cv::Mat& descriptors_2,
                         std::vector<cv::DMatch>& feature_matches);
    std::vector<bool> set_ref_3d_position(std::vector<cv::Point3f>& pts_3d, std::vector<cv::KeyPoint>& keypoints,
                                          cv::Mat& descriptors, Frame& frame);
    void motion_estimation(Frame& frame);
    bool check_motion_estimation();
    void move_frame();
    void write_pose(const Frame& frame);
    void rviz_visualize() {}
    void adaptive_non_maximal_suppresion(std::vector<cv::KeyPoint>& keypoints, const int num);
    bool pipeline(bool& if_insert_keyframe);
    bool insert_key_frame(bool check, std::vector<cv::Point3f>& pts_3d, std::vector<cv::KeyPoint>& keypoints,
                          cv::Mat& descriptors);

private:
    void create_context();
};

}  // namespace vslam

#endif
