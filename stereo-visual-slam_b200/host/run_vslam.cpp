// Main loop of the reference's node without the ROS plumbing
// (/root/reference/src/run_vslam.cpp:17-92): per frame VO::pipeline(), and after every keyframe insertion with a full
// window optimize_map(5) x2 (outlier relabel only), optimize_map(10) with pose write-back and optimize_pose_only(10).
//
//   run_vslam <dataset_dir/> <n_frames> [--no-ba] [--nfeatures N] [--anms K] [--sparse] [--window W] [--update-landmarks]
// (depth source: the reference's dense StereoSGBM unless --sparse; --dense is accepted and redundant)
// reads <dataset_dir>/image_{0,1}/%06d.pgm, appends evicted / remaining keyframe poses to ./estimated_traj.txt
// (the reference's format) and prints one "frame <id> <12 numbers of T_w_c> <inliers> <is_keyframe> <#keyframes>
// <#landmarks> <frame wall-clock ms>" line per frame.
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <iostream>
#include <string>

#include "stereo_visual_slam_main/map.hpp"
#include "stereo_visual_slam_main/optimization.hpp"
#include "stereo_visual_slam_main/visual_odometry.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: run_vslam <dataset_dir/> <n_frames> [--no-ba] [--nfeatures N] [--anms K] [--sparse] [--window W] [--update-landmarks]\n");
        return 2;
    }
    const std::string dataset = argv[1];
    const int n_frames = std::atoi(argv[2]);
    bool do_ba = true, dense = true;
    int nfeatures = 3000, anms = 500, window = 10;
    bool update_landmarks = false;  // the reference never writes landmarks back (run_vslam.cpp:61-64)
    for (int i = 3; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--no-ba")) do_ba = false;
        else if (!std::strcmp(argv[i], "--dense")) dense = true;  // the reference's StereoSGBM depth source (default)
        else if (!std::strcmp(argv[i], "--sparse")) dense = false;  // opt-in fast mode: sparse L<->R matching + DLT
        else if (!std::strcmp(argv[i], "--window") && i + 1 < argc) window = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--update-landmarks")) update_landmarks = true;
        else if (!std::strcmp(argv[i], "--nfeatures") && i + 1 < argc) nfeatures = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--anms") && i + 1 < argc) anms = std::atoi(argv[++i]);
    }
    ros::NodeHandle nh;
    nh.setParam("/dataset", dataset);
    nh.setParam("/if_write_pose", true);
    nh.setParam("/if_rviz", false);

    vslam::Map my_map(nh);
    if (window < 2 || window > 64) {
        std::fprintf(stderr, "--window must be in [2, 64]\n");
        return 2;
    }
    my_map.num_keyframes_ = window;
    vslam::VO my_VO(dataset, nh, my_map);
    my_VO.detector_nfeatures_ = nfeatures;
    my_VO.anms_keep_ = anms;
    my_VO.dense_stereo_ = dense;

    const double fx = 718.856, fy = 718.856, cx = 607.1928, cy = 185.2157;
    cv::Mat K = (cv::Mat_<double>(3, 3) << fx, 0, cx, 0, fy, cy, 0, 0, 1);

    for (int ite = 0; ite < n_frames; ++ite) {
        const auto t_frame = std::chrono::steady_clock::now();
        bool if_insert_keyframe = false;
        const bool not_lost = my_VO.pipeline(if_insert_keyframe);
        if (if_insert_keyframe && do_ba && (int)my_map.keyframes_.size() >= my_map.num_keyframes_) {
            vslam::optimize_map(my_map.keyframes_, my_map.landmarks_, K, false, false, 5);
            vslam::optimize_map(my_map.keyframes_, my_map.landmarks_, K, false, false, 5);
            vslam::optimize_map(my_map.keyframes_, my_map.landmarks_, K, true, update_landmarks, 10);
            vslam::optimize_pose_only(my_map.keyframes_, my_map.landmarks_, K, true, 10);
        }
        const vslam::Frame& f = ite == 0 ? my_VO.frame_last_ : my_VO.frame_current_;
        const SE3 T_w_c = f.T_c_w_.inverse();
        std::printf("frame %d", f.frame_id_);
        for (int r = 0; r < 3; ++r)
            std::printf(" %.9g %.9g %.9g %.9g", T_w_c.rotationMatrix()(r, 0), T_w_c.rotationMatrix()(r, 1),
                        T_w_c.rotationMatrix()(r, 2), T_w_c.translation()(r));
        // wall-clock of this frame: pipeline (+ BA when it ran); the reference's README quotes these per (non-)keyframe
        const double frame_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_frame).count();
        std::printf(" %d %d %zu %zu %.3f\n", my_VO.num_inliers_, (int)if_insert_keyframe, my_map.keyframes_.size(),
                    my_map.landmarks_.size(), frame_ms);
        if (!not_lost) break;
    }
    my_map.write_remaining_pose();
    return 0;
}
