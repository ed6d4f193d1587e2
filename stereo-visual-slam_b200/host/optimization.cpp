// optimize_map / optimize_pose_only with the reference's signatures (src/stereo_visual_slam_main/optimization.cpp:103-436).
// The host side only walks the containers exactly like the reference's graph construction (same filters, same edge
// insertion order) and marshals them into flat arrays; the optimisation itself is one call into the CUDA library.
#include <stereo_visual_slam_main/optimization.hpp>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include <stdexcept>

#include "../../include/vslam_b200.h"

namespace vslam {

static std::vector<vslam_ctx*>& opt_ctx_stack() {
    static std::vector<vslam_ctx*> s;
    return s;
}
void set_optimization_context(vslam_ctx* ctx) {
    if (ctx) opt_ctx_stack().push_back(ctx);
}
void release_optimization_context(vslam_ctx* ctx) {
    auto& s = opt_ctx_stack();
    s.erase(std::remove(s.begin(), s.end(), ctx), s.end());
}
vslam_ctx* optimization_context() { return opt_ctx_stack().empty() ? nullptr : opt_ctx_stack().back(); }

namespace {

struct Graph {
    std::vector<unsigned long> kf_ids;  // dense pose index -> keyframe id
    std::vector<unsigned long> lm_ids;  // dense point index -> landmark id
    std::vector<double> poses, points, uv;
    std::vector<int32_t> obs_pose, obs_point;
};

// Same traversal as the reference: keyframes in container order become poses; landmarks in container order, each with
// its observations in stored order, become edges (optimization.cpp:127-140, 158-214 / 307-316, 332-373).
Graph build_graph(std::unordered_map<unsigned long, Frame>& keyframes,
                  std::unordered_map<unsigned long, Landmark>& landmarks, bool need_reliable_depth) {
    Graph g;
    std::map<unsigned long, int> pose_index;
    for (auto& kf : keyframes) {
        pose_index[kf.first] = (int)g.kf_ids.size();
        g.kf_ids.push_back(kf.first);
        const Eigen::Matrix3d R = kf.second.T_c_w_.rotationMatrix();
        const Eigen::Vector3d t = kf.second.T_c_w_.translation();
        for (int r = 0; r < 3; ++r) {
            g.poses.push_back(R(r, 0)); g.poses.push_back(R(r, 1)); g.poses.push_back(R(r, 2)); g.poses.push_back(t(r));
        }
    }
    for (auto& lm : landmarks) {
        Landmark& L = lm.second;
        if (!L.is_inlier || (need_reliable_depth && !L.reliable_depth_)) continue;
        int point_index = -1;
        for (const Observation& obs : L.observations_) {
            const Feature& feat = keyframes.at(obs.keyframe_id_).features_.at(obs.feature_id_);  // throws like the reference
            if (point_index < 0) {
                point_index = (int)g.lm_ids.size();
                g.lm_ids.push_back(L.landmark_id_);
                g.points.push_back(L.pt_3d_.x); g.points.push_back(L.pt_3d_.y); g.points.push_back(L.pt_3d_.z);
            }
            g.obs_pose.push_back(pose_index.at(obs.keyframe_id_));
            g.obs_point.push_back(point_index);
            g.uv.push_back(feat.keypoint_.pt.x);
            g.uv.push_back(feat.keypoint_.pt.y);
        }
    }
    return g;
}

void run(std::unordered_map<unsigned long, Frame>& keyframes, std::unordered_map<unsigned long, Landmark>& landmarks,
         const cv::Mat& K, bool pose_only, bool if_update_map, bool if_update_landmark, int num_ite) {
    vslam_ctx* g_opt_ctx = optimization_context();
    if (!g_opt_ctx) throw std::runtime_error("vslam::optimize_*: no library context (set_optimization_context)");
    static const bool trace = std::getenv("VSLAM_VO_TRACE") != nullptr;  // per-stage wall clock on stderr
    const auto t0 = std::chrono::steady_clock::now();
    Graph g = build_graph(keyframes, landmarks, !pose_only);
    if (g.kf_ids.empty()) return;
    const auto t1 = std::chrono::steady_clock::now();
    double Kc[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Kc[r * 3 + c] = K.at<double>(r, c);
    vslam_ba_options opt;
    opt.huber_delta = 5.991;      // chi2_th used as the Huber delta (optimization.cpp:154,205)
    opt.chi2_threshold = 5.991;
    opt.num_iterations = num_ite;
    opt.pose_only = pose_only ? 1 : 0;
    opt.max_trials = 10;
    opt.reserved = 0;
    opt.tau = 1e-5;
    vslam_ba_result res;
    std::vector<uint8_t> inlier(g.lm_ids.size() ? g.lm_ids.size() : 1, 1);
    const int st = vslam_ba_optimize(g_opt_ctx, (int)g.kf_ids.size(), g.poses.data(), (int)g.lm_ids.size(),
                                     g.points.data(), (int)g.obs_pose.size(), g.obs_pose.data(), g.obs_point.data(),
                                     g.uv.data(), Kc, &opt, &res, nullptr, inlier.data());
    if (st != VSLAM_OK) throw std::runtime_error(std::string("vslam_ba_optimize: ") + vslam_status_string(st));
    const auto t2 = std::chrono::steady_clock::now();

    // relabel: every landmark that contributed an edge gets the verdict of its last edge (optimization.cpp:254-266)
    for (size_t i = 0; i < g.lm_ids.size(); ++i) landmarks.at(g.lm_ids[i]).is_inlier = inlier[i] != 0;

    if (if_update_map) {  // optimization.cpp:272-287 / 429-435
        for (size_t k = 0; k < g.kf_ids.size(); ++k) {
            const double* T = &g.poses[12 * k];
            Eigen::Matrix3d R;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) R(r, c) = T[r * 4 + c];
            keyframes.at(g.kf_ids[k]).T_c_w_ = SE3(R, Eigen::Vector3d(T[3], T[7], T[11]));
        }
        if (if_update_landmark && !pose_only)
            for (size_t i = 0; i < g.lm_ids.size(); ++i)
                landmarks.at(g.lm_ids[i]).pt_3d_ = cv::Point3f((float)g.points[3 * i], (float)g.points[3 * i + 1], (float)g.points[3 * i + 2]);
    }
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[ba] %s K %zu L %zu obs %zu: graph %.3f optimize %.3f (%d iterations, %d trials) write-back %.3f ms\n",
                     pose_only ? "pose-only" : "map", g.kf_ids.size(), g.lm_ids.size(), g.obs_pose.size(), ms(t0, t1), ms(t1, t2),
                     res.iterations, res.trials, ms(t2, t3));
    }
}

}  // namespace

void optimize_map(std::unordered_map<unsigned long, Frame>& keyframes,
                  std::unordered_map<unsigned long, Landmark>& landmarks, const cv::Mat& K, bool if_update_map,
                  bool if_update_landmark, int num_ite) {
    run(keyframes, landmarks, K, false, if_update_map, if_update_landmark, num_ite);
}

void optimize_pose_only(std::unordered_map<unsigned long, Frame>& keyframes,
                        std::unordered_map<unsigned long, Landmark>& landmarks, const cv::Mat& K, bool if_update_map,
                        int num_ite) {
    run(keyframes, landmarks, K, true, if_update_map, false, num_ite);
}

}  // namespace vslam
