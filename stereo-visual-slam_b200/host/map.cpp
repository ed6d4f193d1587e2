// Sliding-window bookkeeping (reference: src/stereo_visual_slam_main/map.cpp:13-204).  Same observable behaviour:
// at most num_keyframes_ keyframes; when one more arrives, drop the keyframe nearest to the newest one if it is closer
// than 0.2 (norm of the SE3 log of the relative motion), otherwise the farthest; landmarks without observations go.
#include <stereo_visual_slam_main/map.hpp>

#include <fstream>

namespace vslam {

Map::Map(ros::NodeHandle& nh) : my_visual_(nh) {
    nh.getParam("/if_write_pose", if_write_pose_);
    nh.getParam("/if_rviz", if_rviz_);
}

int Map::insert_keyframe(Frame frame_to_add) {
    current_keyframe_id_ = frame_to_add.keyframe_id_;
    keyframes_[frame_to_add.keyframe_id_] = frame_to_add;
    if ((int)keyframes_.size() > num_keyframes_) remove_keyframe();
    return 0;
}

int Map::insert_landmark(Landmark landmark_to_add) {
    landmarks_[landmark_to_add.landmark_id_] = landmark_to_add;
    return 0;
}

int Map::remove_keyframe() {
    const SE3 T_w_c = keyframes_.at(current_keyframe_id_).T_c_w_.inverse();
    double far_d = 0, near_d = 1000000;
    unsigned long far_id = 0, near_id = 0;
    for (auto& kf : keyframes_) {
        if ((int)kf.first == current_keyframe_id_) continue;
        const double d = (kf.second.T_c_w_ * T_w_c).log().norm();
        if (d > far_d) { far_d = d; far_id = kf.first; }
        if (d < near_d) { near_d = d; near_id = kf.first; }
    }
    const unsigned long victim = near_d < 0.2 ? near_id : far_id;

    // detach the victim's observations from their landmarks
    for (const Feature& feat : keyframes_.at(victim).features_) {
        auto it = landmarks_.find(feat.landmark_id_);
        if (it == landmarks_.end()) continue;
        std::vector<Observation>& obs = it->second.observations_;
        obs.erase(std::remove_if(obs.begin(), obs.end(),
                                 [&](const Observation& o) {
                                     return o.keyframe_id_ == (int)victim && o.feature_id_ == feat.feature_id_;
                                 }),
                  obs.end());
        it->second.observed_times_--;
    }
    if (if_write_pose_) write_pose(keyframes_.at(victim));
    keyframes_.erase(victim);
    clean_map();
    return 0;
}

int Map::clean_map() {
    for (auto it = landmarks_.begin(); it != landmarks_.end();) {
        if (it->second.observed_times_ == 0) it = landmarks_.erase(it);
        else ++it;
    }
    return 0;
}

// KITTI pose row prefixed by the frame id: id r00 r01 r02 x r10 r11 r12 y r20 r21 r22 z of T_w_c (map.cpp:168-196)
void Map::write_pose(const Frame& frame) {
    const SE3 T_w_c = frame.T_c_w_.inverse();
    const Eigen::Matrix3d R = T_w_c.rotationMatrix();
    const Eigen::Vector3d t = T_w_c.translation();
    std::ofstream f("estimated_traj.txt", std::ios_base::app);
    f << frame.frame_id_;
    for (int r = 0; r < 3; ++r) f << " " << R(r, 0) << " " << R(r, 1) << " " << R(r, 2) << " " << t(r);
    f << std::endl;
}

void Map::write_remaining_pose() {
    for (auto& kf : keyframes_) write_pose(kf.second);
}

}  // namespace vslam
