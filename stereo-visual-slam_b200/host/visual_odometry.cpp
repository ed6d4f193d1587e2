// VO front end: the reference's stage methods and state machine
// (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:20-706) over the CUDA library.
// Compute stages (feature_detection, adaptive_non_maximal_suppresion, feature_matching, disparity_map,
// motion_estimation) are one C-ABI call each; everything else is the host bookkeeping the reference does, expressed
// with index maps instead of its O(N*M) scans where the result is identical.
#include <stereo_visual_slam_main/visual_odometry.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <unordered_map>

#include "../../include/vslam_b200.h"

namespace vslam {

namespace {
const int kMaxW = 1920, kMaxH = 1200, kMaxKp = 8192;

void check(int st, const char* what) {
    if (st != VSLAM_OK) throw std::runtime_error(std::string(what) + ": " + vslam_status_string(st));
}
}  // namespace

void VO::create_context() {
    vslam_config cfg;
    cfg.device = 0;
    cfg.max_images = 2;
    cfg.max_width = kMaxW;
    cfg.max_height = kMaxH;
    cfg.max_keypoints = kMaxKp;
    cfg.max_ba_poses = 64;
    cfg.max_ba_points = 65536;
    cfg.max_ba_obs = 262144;
    check(vslam_ctx_create(&cfg, &ctx_), "vslam_ctx_create");  // throws without a B200: there is no CPU path
    set_optimization_context(ctx_);
    if (!std::getenv("VSLAM_NO_PINNED_MAT")) {  // image-sized cv::Mat buffers from the library's pinned pool
        cv::Mat::allocator().alloc = vslam_host_alloc;
        cv::Mat::allocator().free = vslam_host_free;
    }
    // runtime knobs for callers that cannot reach the members (the reference's unmodified main, run_vslam.cpp:17-92)
    if (const char* e = std::getenv("VSLAM_NFEATURES")) detector_nfeatures_ = std::atoi(e);
    if (const char* e = std::getenv("VSLAM_ANMS_KEEP")) anms_keep_ = std::atoi(e);
    if (const char* e = std::getenv("VSLAM_SPARSE_STEREO")) dense_stereo_ = std::atoi(e) == 0;
}

VO::VO(ros::NodeHandle& nh, Map& map) : my_map_(map), my_visual_(nh) {
    create_context();
    nh.getParam("/if_rviz", if_rviz_);
}

VO::VO(std::string dataset, ros::NodeHandle& nh, Map& map) : my_map_(map), my_visual_(nh) {
    dataset_ = dataset;
    create_context();
    nh.getParam("/if_rviz", if_rviz_);
}

VO::~VO() {
    release_optimization_context(ctx_);
    vslam_ctx_destroy(ctx_);
}

// ---- I/O -------------------------------------------------------------------------------------------------------
static bool read_pgm(const std::string& path, cv::Mat& img) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    f >> magic >> w >> h >> maxv;
    if (magic != "P5" || w <= 0 || h <= 0 || maxv != 255) return false;
    f.get();
    img.create(h, w, cv::CV_8U);
    f.read(reinterpret_cast<char*>(img.data), (std::streamsize)w * h);
    return (bool)f;
}

bool read_png_gray8(const std::string& path, std::vector<uint8_t>& grey, int& w, int& h);  // png_reader.cpp

static bool read_png(const std::string& path, cv::Mat& img) {
    std::vector<uint8_t> g;
    int w = 0, h = 0;
    if (!read_png_gray8(path, g, w, h)) return false;
    img.create(h, w, cv::CV_8U);
    std::memcpy(img.data, g.data(), g.size());
    return true;
}

// image_0/%06d.png as the reference reads it (visual_odometry.cpp:42-51, cv::imread GRAYSCALE); binary PGM files of the
// same name are accepted as well
static bool read_gray(const std::string& stem, cv::Mat& img) { return read_png(stem + ".png", img) || read_pgm(stem + ".pgm", img); }

int VO::read_img(int id, cv::Mat& left_img, cv::Mat& right_img) {
    if (image_source_) return image_source_(id, left_img, right_img);
    char name[16];
    std::snprintf(name, sizeof(name), "%06d", id);
    const bool ok = read_gray(dataset_ + "image_0/" + name, left_img) && read_gray(dataset_ + "image_1/" + name, right_img);
    if (!ok || !left_img.data) {
        std::cout << "Could not open or find the image" << std::endl;
        return -1;
    }
    return 0;
}

// ---- compute stages --------------------------------------------------------------------------------------------
int VO::feature_detection(const cv::Mat& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {
    if (!img.data) {
        std::cout << "Could not open or find the image" << std::endl;
        return -1;
    }
    const int cap = vslam_orb_keypoint_capacity(ctx_);
    std::vector<cv::KeyPoint> kp(cap);
    cv::Mat desc(cap, 32, cv::CV_8U);
    int32_t n = 0;
    const int st = vslam_orb_detect_compute(ctx_, img.data, img.cols, img.rows, (int)img.step, detector_nfeatures_,
                                            anms_keep_, 1.11f, reinterpret_cast<vslam_keypoint*>(kp.data()), desc.data, &n);
    if (st == VSLAM_E_INVALID) return -1;
    if (st == VSLAM_E_OVERFLOW || st == VSLAM_E_CAPACITY) {
        // a device work list overflowed / the image exceeds the context: degrade like the reference does on a bad frame
        // (no keypoints -> no matches -> "Rejected - inliers not enough") instead of throwing out of VO::pipeline
        std::cout << "feature_detection: " << vslam_status_string(st) << " -- frame dropped" << std::endl;
        keypoints.clear();
        descriptors = cv::Mat(0, 32, cv::CV_8U);
        return -1;
    }
    check(st, "vslam_orb_detect_compute");
    kp.resize(n);
    keypoints.swap(kp);
    descriptors = cv::Mat(n, 32, cv::CV_8U);  // owning N x 32 matrix, rows are handed out as views (Feature::descriptor_)
    if (n > 0) std::memcpy(descriptors.data, desc.data, (size_t)n * 32);
    return 0;
}

void VO::adaptive_non_maximal_suppresion(std::vector<cv::KeyPoint>& keypoints, const int num) {
    if ((int)keypoints.size() < num) return;
    std::vector<int32_t> keep(keypoints.size());
    int32_t n = 0;
    check(vslam_anms(ctx_, reinterpret_cast<const vslam_keypoint*>(keypoints.data()), (int)keypoints.size(), num, 1.11f,
                     keep.data(), &n),
          "vslam_anms");
    std::vector<cv::KeyPoint> out;
    out.reserve(n);
    for (int i = 0; i < n; ++i) out.push_back(keypoints[keep[i]]);
    // the reference returns the survivors in descending-response order (it sorts before suppressing)
    std::stable_sort(out.begin(), out.end(), [](const cv::KeyPoint& a, const cv::KeyPoint& b) { return a.response > b.response; });
    keypoints.swap(out);
}

int VO::feature_matching(const cv::Mat& descriptors_1, const cv::Mat& descriptors_2,
                         std::vector<cv::DMatch>& feature_matches) {
    feature_matches.clear();
    const int nq = descriptors_1.rows, nt = descriptors_2.rows;
    if (nq == 0 || nt == 0) return 0;  // the reference dereferences an empty range here (UB); we return no matches
    const double frame_gap = frame_current_.frame_id_ - frame_last_.frame_id_;
    std::vector<cv::DMatch> out(std::min(nq, nt));
    int n = 0;
    check(vslam_match_hamming(ctx_, descriptors_1.data, nq, descriptors_2.data, nt, 1, 2.0, 30.0 * frame_gap,
                              reinterpret_cast<vslam_dmatch*>(out.data()), &n),
          "vslam_match_hamming");
    out.resize(n);
    feature_matches.swap(out);
    return 0;
}

int VO::disparity_map(const Frame& frame, cv::Mat& disparity) {
    const cv::Mat& L = frame.left_img_;
    const cv::Mat& R = frame.right_img_;
    if (!L.data || !R.data) return -1;
    if (dense_stereo_) {
        // the reference's own path (visual_odometry.cpp:163-168): StereoSGBM(0, 96, 9, 8*81, 32*81, 1, 63, 10, 100, 32)
        // followed by convertTo(CV_32F, 1/16); both happen on the device, bit-exact against cv2 4.13
        disparity.create(L.rows, L.cols, cv::CV_32F);
        check(vslam_sgbm_compute(ctx_, L.data, R.data, 1, L.cols, L.rows, (int)L.step, (long long)L.step * L.rows, nullptr,
                                 nullptr, reinterpret_cast<float*>(disparity.data)),
              "vslam_sgbm_compute");
        return 0;
    }
    const int cap = vslam_orb_keypoint_capacity(ctx_);
    std::vector<vslam_keypoint> kp(2 * (size_t)cap);
    std::vector<uint8_t> desc(2 * (size_t)cap * 32), flags(cap);
    std::vector<vslam_dmatch> matches(cap);
    std::vector<float> xyz(3 * (size_t)cap);
    int32_t n_kp[2] = {0, 0}, n_m = 0;
    // rectified pair: P1 = K [I|0], P2 = K [I|(-b,0,0)]
    const double P1[12] = {frame.fx_, 0, frame.cx_, 0, 0, frame.fy_, frame.cy_, 0, 0, 0, 1, 0};
    const double P2[12] = {frame.fx_, 0, frame.cx_, -frame.fx_ * frame.b_, 0, frame.fy_, frame.cy_, 0, 0, 0, 1, 0};
    check(vslam_stereo_frontend_batch(ctx_, L.data, R.data, 1, L.cols, L.rows, (int)L.step, (long long)L.step * L.rows,
                                      detector_nfeatures_, anms_keep_, 1.11f, 2.0, 30.0, P1, P2, nullptr, kp.data(),
                                      desc.data(), n_kp, matches.data(), &n_m, xyz.data(), flags.data()),
          "vslam_stereo_frontend_batch");
    disparity.create(L.rows, L.cols, cv::CV_32F);
    disparity.setTo(-1.0);  // "no depth", as SGBM's invalid value after the /16 conversion
    for (int i = 0; i < n_m; ++i) {
        const vslam_keypoint& k = kp[matches[i].queryIdx];
        const vslam_keypoint& kr = kp[(size_t)cap + matches[i].trainIdx];
        const float Z = xyz[3 * i + 2];
        // rectified pair: a true correspondence lies on the same row (within the ORB localisation error of the coarser
        // octaves) and to the left in the right image
        if (!(Z > 0) || !(k.x - kr.x > 0.f) || std::fabs(k.y - kr.y) > 2.f) continue;
        disparity.at<float>((int)k.y, (int)k.x) = (float)(frame.fx_ * frame.b_ / Z);
    }
    return 0;
}

std::vector<bool> VO::set_ref_3d_position(std::vector<cv::Point3f>& pts_3d, std::vector<cv::KeyPoint>& keypoints,
                                          cv::Mat& descriptors, Frame& frame) {
    pts_3d.clear();
    std::vector<cv::KeyPoint> kept_kp;
    std::vector<int> kept_rows;
    std::vector<bool> reliable_depth;
    for (size_t i = 0; i < keypoints.size(); ++i) {
        Eigen::Vector3d rel;
        const Eigen::Vector3d w = frame.find_3d(keypoints[i], rel);
        if (rel(2) > 10 && rel(2) < 400) {  // usable depth (visual_odometry.cpp:194)
            pts_3d.push_back(cv::Point3f((float)w(0), (float)w(1), (float)w(2)));
            kept_kp.push_back(keypoints[i]);
            kept_rows.push_back((int)i);
            reliable_depth.push_back(rel(2) < 40);  // visual_odometry.cpp:201
        }
    }
    cv::Mat filtered((int)kept_rows.size(), 32, cv::CV_8U);
    for (size_t r = 0; r < kept_rows.size(); ++r) std::memcpy(filtered.ptr<uint8_t>((int)r), descriptors.ptr<uint8_t>(kept_rows[r]), 32);
    descriptors = filtered;
    keypoints.swap(kept_kp);
    return reliable_depth;
}

void VO::motion_estimation(Frame& frame) {
    const size_t n = frame.features_.size();
    std::vector<float> xyz(3 * n), uv(2 * n);
    for (size_t i = 0; i < n; ++i) {
        const int landmark_id = frame.features_[i].landmark_id_;
        if (landmark_id == -1) std::cout << "No landmark associated!" << std::endl;
        const cv::Point3f& p = my_map_.landmarks_.at(landmark_id).pt_3d_;  // throws like the reference
        xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
        uv[2 * i] = frame.features_[i].keypoint_.pt.x;
        uv[2 * i + 1] = frame.features_[i].keypoint_.pt.y;
    }
    const double K[9] = {frame.fx_, 0, frame.cx_, 0, frame.fy_, frame.cy_, 0, 0, 1};
    double rvec[3] = {0, 0, 0}, tvec[3] = {0, 0, 0}, T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    std::vector<int32_t> inliers(n ? n : 1);
    int32_t n_inl = 0;
    check(vslam_pnp_ransac(ctx_, xyz.data(), uv.data(), (int)n, K, 100, 4.0f, 0.99, rvec, tvec, T, inliers.data(), &n_inl),
          "vslam_pnp_ransac");
    num_inliers_ = n_inl;
    if (n_inl > 0) {
        Eigen::Matrix3d R;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R(r, c) = T[r * 4 + c];
        T_c_w_ = SE3(R, Eigen::Vector3d(T[3], T[7], T[11]));
    } else {
        T_c_w_ = SE3();  // cv::solvePnPRansac leaves rvec/tvec empty on failure; the reference then reads zeros
    }
    for (int i = 0; i < n_inl; ++i) frame.features_[inliers[i]].is_inlier = true;
    frame.features_.erase(std::remove_if(frame.features_.begin(), frame.features_.end(),
                                         [](const Feature& f) { return !f.is_inlier; }),
                          frame.features_.end());
}

// ---- state machine ---------------------------------------------------------------------------------------------
bool VO::check_motion_estimation() {
    if (num_inliers_ < 10) {
        std::cout << "Frame id: " << frame_last_.frame_id_ << " and " << frame_current_.frame_id_ << std::endl;
        std::cout << "Rejected - inliers not enough: " << num_inliers_ << std::endl;
        return false;
    }
    const double frame_gap = frame_current_.frame_id_ - frame_last_.frame_id_;
    const double motion = T_c_l_.log().norm();
    if (motion > 5.0 * frame_gap) {
        std::cout << "Frame id: " << frame_last_.frame_id_ << " and " << frame_current_.frame_id_ << std::endl;
        std::cout << "Rejected - motion is too large: " << motion << std::endl;
        return false;
    }
    return true;
}

bool VO::insert_key_frame(bool check_ok, std::vector<cv::Point3f>& pts_3d, std::vector<cv::KeyPoint>& keypoints,
                          cv::Mat& descriptors) {
    // keyframe unless tracking is comfortable (>= 80 inliers and little yaw -- signed test, as the reference) or rejected
    if ((num_inliers_ >= 80 && T_c_l_.angleY() < 0.03) || !check_ok) return false;

    frame_current_.is_keyframe_ = true;
    frame_current_.keyframe_id_ = curr_keyframe_id_;
    for (Feature& f : frame_current_.features_) {
        Landmark& lm = my_map_.landmarks_.at(f.landmark_id_);
        lm.observed_times_++;
        lm.observations_.push_back(Observation(frame_current_.keyframe_id_, f.feature_id_));
    }

    disparity_map(frame_current_, frame_current_.disparity_);
    std::vector<bool> reliable_depth = set_ref_3d_position(pts_3d, keypoints, descriptors, frame_current_);

    // which detected keypoints are already tracked features?  (the reference compares pt.x / pt.y exactly in a nested
    // loop, visual_odometry.cpp:385-401; a hash of the bit patterns gives the same answer in O(N))
    auto key = [](const cv::Point2f& p) {
        uint32_t a, b;
        std::memcpy(&a, &p.x, 4);
        std::memcpy(&b, &p.y, 4);
        return ((uint64_t)a << 32) | b;
    };
    std::unordered_multimap<uint64_t, size_t> tracked;
    for (size_t j = 0; j < frame_current_.features_.size(); ++j) tracked.emplace(key(frame_current_.features_[j].keypoint_.pt), j);

    int feature_id = (int)frame_current_.features_.size();
    for (size_t i = 0; i < keypoints.size(); ++i) {
        bool exist = false;
        auto range = tracked.equal_range(key(keypoints[i].pt));
        for (auto it = range.first; it != range.second; ++it) {
            exist = true;
            Landmark& lm = my_map_.landmarks_.at(frame_current_.features_[it->second].landmark_id_);
            if (!lm.reliable_depth_ && reliable_depth[i]) {  // upgrade an unreliable depth
                lm.pt_3d_ = pts_3d[i];
                lm.reliable_depth_ = true;
            }
        }
        if (!exist) {
            Feature f(feature_id, frame_current_.frame_id_, keypoints[i], descriptors.row((int)i));
            f.landmark_id_ = curr_landmark_id_;
            frame_current_.features_.push_back(f);
            tracked.emplace(key(keypoints[i].pt), frame_current_.features_.size() - 1);  // later duplicates see it, as in the reference
            my_map_.insert_landmark(Landmark(curr_landmark_id_, pts_3d[i], descriptors.row((int)i), reliable_depth[i],
                                             Observation(frame_current_.keyframe_id_, feature_id)));
            curr_landmark_id_++;
            feature_id++;
        }
    }
    curr_keyframe_id_++;
    my_map_.insert_keyframe(frame_current_);
    return true;
}

void VO::move_frame() { frame_last_ = frame_current_; }

void VO::write_pose(const Frame& frame) { my_map_.write_pose(frame); }

bool VO::initialization() {
    frame_last_ = Frame();
    if (read_img(0, frame_last_.left_img_, frame_last_.right_img_) != 0) return false;
    frame_last_.frame_id_ = 0;

    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    std::vector<cv::Point3f> pts_3d;
    if (feature_detection(frame_last_.left_img_, keypoints, descriptors) != 0) return false;
    disparity_map(frame_last_, frame_last_.disparity_);
    std::vector<bool> reliable_depth = set_ref_3d_position(pts_3d, keypoints, descriptors, frame_last_);

    for (size_t i = 0; i < keypoints.size(); ++i) {
        Feature f((int)i, 0, keypoints[i], descriptors.row((int)i));
        f.landmark_id_ = curr_landmark_id_;
        frame_last_.features_.push_back(f);
        my_map_.insert_landmark(Landmark(curr_landmark_id_, pts_3d[i], descriptors.row((int)i), reliable_depth[i], Observation(0, (int)i)));
        curr_landmark_id_++;
    }
    frame_last_.fill_frame(SE3(), true, curr_keyframe_id_);
    curr_keyframe_id_++;
    my_map_.insert_keyframe(frame_last_);
    return true;
}

bool VO::tracking(bool& if_insert_keyframe) {
    // VSLAM_VO_TRACE=1: per-stage wall clock of this frame on stderr
    static const bool trace = std::getenv("VSLAM_VO_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    double t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto lap = [&](int k) {
        const auto t1 = std::chrono::steady_clock::now();
        t_ms[k] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
    };
    frame_current_ = Frame();
    // pick up the BA-optimised pose of the last keyframe
    if (frame_last_.is_keyframe_) {
        auto it = my_map_.keyframes_.find(frame_last_.keyframe_id_);
        if (it != my_map_.keyframes_.end()) frame_last_ = it->second;
    }
    if (read_img(seq_, frame_current_.left_img_, frame_current_.right_img_) != 0) {
        seq_++;
        return false;
    }
    frame_current_.frame_id_ = seq_;
    lap(0);

    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    feature_detection(frame_current_.left_img_, keypoints, descriptors);
    lap(1);

    // descriptors of the last frame's features, gathered into one matrix
    cv::Mat descriptors_last((int)frame_last_.features_.size(), 32, cv::CV_8U);
    for (size_t i = 0; i < frame_last_.features_.size(); ++i)
        std::memcpy(descriptors_last.ptr<uint8_t>((int)i), frame_last_.features_[i].descriptor_.data, 32);
    std::vector<cv::DMatch> matches;
    feature_matching(descriptors_last, descriptors, matches);  // query = last frame, train = current frame
    lap(2);

    for (size_t i = 0; i < matches.size(); ++i) {
        Feature f((int)i, seq_, keypoints[matches[i].trainIdx], descriptors.row(matches[i].trainIdx));
        f.landmark_id_ = frame_last_.features_[matches[i].queryIdx].landmark_id_;
        frame_current_.features_.push_back(f);
    }

    lap(5);
    motion_estimation(frame_current_);
    lap(3);
    frame_current_.T_c_w_ = T_c_w_;
    T_c_l_ = frame_current_.T_c_w_ * frame_last_.T_c_w_.inverse();

    const bool ok = check_motion_estimation();
    lap(6);
    std::vector<cv::Point3f> pts_3d;
    if_insert_keyframe = insert_key_frame(ok, pts_3d, keypoints, descriptors);
    lap(4);
    if (ok) move_frame();
    lap(7);
    if (trace)
        std::fprintf(stderr, "[vo] frame %d: read %.3f detect %.3f match %.3f pnp %.3f check %.3f keyframe %.3f move %.3f bookkeeping %.3f ms\n",
                     seq_, t_ms[0], t_ms[1], t_ms[2], t_ms[3], t_ms[6], t_ms[4], t_ms[7], t_ms[5]);
    seq_++;
    return ok;
}

// VSLAM_FRAME_LOG=<path>: one "frame <id> <12 numbers of T_w_c> <inliers> <is_keyframe> <#keyframes> <#landmarks>"
// line per processed frame (the same record run_vslam prints), so that any main loop over VO::pipeline can be compared
static void log_frame(const Frame& f, int inliers, bool kf, const Map& map) {
    static const char* path = std::getenv("VSLAM_FRAME_LOG");
    if (!path) return;
    FILE* fp = std::fopen(path, "a");
    if (!fp) return;
    const SE3 T_w_c = f.T_c_w_.inverse();
    std::fprintf(fp, "frame %d", f.frame_id_);
    for (int r = 0; r < 3; ++r)
        std::fprintf(fp, " %.9g %.9g %.9g %.9g", T_w_c.rotationMatrix()(r, 0), T_w_c.rotationMatrix()(r, 1),
                     T_w_c.rotationMatrix()(r, 2), T_w_c.translation()(r));
    std::fprintf(fp, " %d %d %zu %zu\n", inliers, (int)kf, map.keyframes_.size(), map.landmarks_.size());
    std::fclose(fp);
}

bool VO::pipeline(bool& if_insert_keyframe) {
    switch (state_) {
        case Init:
            if (initialization()) {
                state_ = Track;
                log_frame(frame_last_, num_inliers_, false, my_map_);
            } else if (++num_lost_ > 10) {
                state_ = Lost;
            }
            break;
        case Track:
            if (tracking(if_insert_keyframe)) {
                num_lost_ = 0;
            } else if (++num_lost_ > 10) {
                state_ = Lost;
            }
            if (frame_current_.left_img_.data) log_frame(frame_current_, num_inliers_, if_insert_keyframe, my_map_);
            break;
        case Lost:
            std::cout << "VO IS LOST" << std::endl;
            return false;
        default:
            std::cout << "Invalid state" << std::endl;
            return false;
    }
    return true;
}

}  // namespace vslam
