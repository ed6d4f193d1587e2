// Sliding-window bundle adjustment entry points with the reference's signatures
// (/root/reference/include/stereo_visual_slam_main/optimization.hpp:137-152).  The g2o vertex/edge classes of the
// reference (optimization.hpp:37-124) have no counterpart here: their arithmetic (oplus, residual, Jacobians) runs
// inside the CUDA kernel behind vslam_ba_optimize (csrc/ba.cu).
#ifndef VSLAM_B200_OPTIMIZATION_HPP
#define VSLAM_B200_OPTIMIZATION_HPP

#include <stereo_visual_slam_main/library_include.hpp>
#include <stereo_visual_slam_main/map.hpp>
#include <stereo_visual_slam_main/types_def.hpp>

#include <unordered_map>

struct vslam_ctx;

namespace vslam {

// The library context used by optimize_map / optimize_pose_only (the reference's free functions take no handle).
// Contexts are kept on a small stack: every VO registers its context on construction and withdraws it on destruction,
// the most recently registered one still alive is used -- destroying one VO never leaves another VO's optimize_* calls
// without a context.  Stand-alone callers can register their own.  Not owned.
void set_optimization_context(vslam_ctx* ctx);       // register (push); nullptr is ignored
void release_optimization_context(vslam_ctx* ctx);   // withdraw every registration of ctx
vslam_ctx* optimization_context();                   // the context optimize_* will use, or nullptr

// optimize_map: LM(Schur) over every keyframe pose and every landmark with is_inlier && reliable_depth_, `num_ite`
// iterations, then the adaptive chi2 relabel of Landmark::is_inlier; poses written back if if_update_map, landmarks
// if additionally if_update_landmark.
void optimize_map(std::unordered_map<unsigned long, Frame>& keyframes,
                  std::unordered_map<unsigned long, Landmark>& landmarks, const cv::Mat& K, bool if_update_map,
                  bool if_update_landmark, int num_ite);

// optimize_pose_only: poses only, landmarks fixed, every is_inlier landmark contributes.
void optimize_pose_only(std::unordered_map<unsigned long, Frame>& keyframes,
                        std::unordered_map<unsigned long, Landmark>& landmarks, const cv::Mat& K, bool if_update_map,
                        int num_ite);

}  // namespace vslam

#endif
