// Minimal stand-ins for the third-party types that appear in the reference's public signatures
// (cv::Mat / cv::KeyPoint / cv::DMatch / cv::Point*, Eigen::Vector3d / Matrix3d, Sophus::SE3d, ros::NodeHandle),
// just large enough that the drop-in headers in stereo_visual_slam_main/ keep the reference's signatures
// (/root/reference/include/stereo_visual_slam_main/{types_def,visual_odometry,optimization,map}.hpp) in a container
// that has none of OpenCV-C++, Eigen, Sophus or ROS.  On a machine that has them, define VSLAM_USE_REAL_DEPS and the
// real headers are used instead -- the class and struct definitions that follow compile against either.
//
// Layout facts relied upon by the C-ABI marshalling: cv::KeyPoint is 28 bytes {pt.x, pt.y, size, angle, response,
// octave, class_id} and cv::DMatch is 16 bytes {queryIdx, trainIdx, imgIdx, distance} -- identical to vslam_keypoint /
// vslam_dmatch in include/vslam_b200.h, so vectors of them are passed to the library without conversion.
#pragma once

#ifdef VSLAM_USE_REAL_DEPS
#include "ros/ros.h"
#include <Eigen/Core>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include "sophus/se3.hpp"
#include "sophus/so3.hpp"
#else

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

// ----------------------------------------------------------------------------------------------------------------
namespace Eigen {

template <int N>
struct VecN {
    double v[N];
    VecN() { for (int i = 0; i < N; ++i) v[i] = 0; }
    VecN(double a, double b, double c) { static_assert(N == 3, "3-vector ctor"); v[0] = a; v[1] = b; v[2] = c; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double squaredNorm() const { double s = 0; for (int i = 0; i < N; ++i) s += v[i] * v[i]; return s; }
    double norm() const { return std::sqrt(squaredNorm()); }
    VecN operator+(const VecN& o) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    VecN operator-(const VecN& o) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
    VecN operator*(double s) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * s; return r; }
    static VecN Zero() { return VecN(); }
};
typedef VecN<3> Vector3d;
typedef VecN<2> Vector2d;

struct Matrix3d {
    double m[9];
    Matrix3d() { std::memset(m, 0, sizeof(m)); }
    double& operator()(int r, int c) { return m[r * 3 + c]; }
    double operator()(int r, int c) const { return m[r * 3 + c]; }
    static Matrix3d Identity() { Matrix3d I; I.m[0] = I.m[4] = I.m[8] = 1; return I; }
    Matrix3d operator*(const Matrix3d& o) const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = m[i * 3] * o.m[j] + m[i * 3 + 1] * o.m[3 + j] + m[i * 3 + 2] * o.m[6 + j];
        return r;
    }
    Vector3d operator*(const Vector3d& p) const {
        return Vector3d(m[0] * p[0] + m[1] * p[1] + m[2] * p[2], m[3] * p[0] + m[4] * p[1] + m[5] * p[2],
                        m[6] * p[0] + m[7] * p[1] + m[8] * p[2]);
    }
    Matrix3d transpose() const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = m[j * 3 + i];
        return r;
    }
};

}  // namespace Eigen

// ----------------------------------------------------------------------------------------------------------------
namespace Sophus {

typedef Eigen::VecN<6> Vector6d;

// SE3 as rotation matrix + translation (Sophus stores a unit quaternion; conventions per SURVEY.md §A.6:
// tangent = [upsilon; omega], exp/log with the V matrix, left-multiplicative updates, angleY()).
class SE3d {
public:
    SE3d() : R_(Eigen::Matrix3d::Identity()) {}
    SE3d(const Eigen::Matrix3d& R, const Eigen::Vector3d& t) : R_(R), t_(t) {}
    const Eigen::Matrix3d& rotationMatrix() const { return R_; }
    const Eigen::Vector3d& translation() const { return t_; }
    Eigen::Vector3d& translation() { return t_; }
    SE3d inverse() const {
        Eigen::Matrix3d Rt = R_.transpose();
        Eigen::Vector3d ti = Rt * t_;
        return SE3d(Rt, Eigen::Vector3d(-ti[0], -ti[1], -ti[2]));
    }
    SE3d operator*(const SE3d& o) const { return SE3d(R_ * o.R_, R_ * o.t_ + t_); }
    Eigen::Vector3d operator*(const Eigen::Vector3d& p) const { return R_ * p + t_; }
    double angleY() const { return std::atan2(-R_(2, 0), std::hypot(R_(2, 1), R_(2, 2))); }

    static SE3d exp(const Vector6d& a) {
        const double w0 = a[3], w1 = a[4], w2 = a[5];
        const double th2 = w0 * w0 + w1 * w1 + w2 * w2, th = std::sqrt(th2);
        double A, B, C;  // R = I + A W + B W^2, V = I + B W + C W^2
        if (th < 1e-10) {
            A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
        } else {
            A = std::sin(th) / th; B = (1 - std::cos(th)) / th2; C = (th - std::sin(th)) / (th2 * th);
        }
        Eigen::Matrix3d W, W2, R, V;
        W(0, 1) = -w2; W(0, 2) = w1; W(1, 0) = w2; W(1, 2) = -w0; W(2, 0) = -w1; W(2, 1) = w0;
        W2 = W * W;
        for (int i = 0; i < 9; ++i) {
            R.m[i] = A * W.m[i] + B * W2.m[i];
            V.m[i] = B * W.m[i] + C * W2.m[i];
        }
        R.m[0] += 1; R.m[4] += 1; R.m[8] += 1;
        V.m[0] += 1; V.m[4] += 1; V.m[8] += 1;
        return SE3d(R, V * Eigen::Vector3d(a[0], a[1], a[2]));
    }

    Vector6d log() const {
        // SO3 log
        const double tr = R_(0, 0) + R_(1, 1) + R_(2, 2);
        double c = std::min(1.0, std::max(-1.0, 0.5 * (tr - 1.0)));
        const double th = std::acos(c);
        double w[3] = {R_(2, 1) - R_(1, 2), R_(0, 2) - R_(2, 0), R_(1, 0) - R_(0, 1)};
        double f = (th < 1e-10) ? 0.5 + th * th / 12.0 : th / (2.0 * std::sin(th));
        if (M_PI - th < 1e-6) {  // near pi: recover the axis from the diagonal
            double ax[3];
            for (int i = 0; i < 3; ++i) ax[i] = std::sqrt(std::max(0.0, (R_(i, i) - c) / (1 - c)));
            if (w[0] < 0) ax[0] = -ax[0];
            if (w[1] < 0) ax[1] = -ax[1];
            if (w[2] < 0) ax[2] = -ax[2];
            for (int i = 0; i < 3; ++i) w[i] = ax[i] * th;
            f = 1.0;
        }
        for (int i = 0; i < 3; ++i) w[i] *= f;
        const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], t = std::sqrt(th2);
        Eigen::Matrix3d W, W2, Vi;
        W(0, 1) = -w[2]; W(0, 2) = w[1]; W(1, 0) = w[2]; W(1, 2) = -w[0]; W(2, 0) = -w[1]; W(2, 1) = w[0];
        W2 = W * W;
        // V^-1 = I - W/2 + k W^2, k = (1 - t cos(t/2) / (2 sin(t/2))) / t^2
        const double k = (t < 1e-10) ? 1.0 / 12.0 : (1.0 - t * std::cos(0.5 * t) / (2.0 * std::sin(0.5 * t))) / th2;
        for (int i = 0; i < 9; ++i) Vi.m[i] = -0.5 * W.m[i] + k * W2.m[i];
        Vi.m[0] += 1; Vi.m[4] += 1; Vi.m[8] += 1;
        Eigen::Vector3d u = Vi * t_;
        Vector6d r;
        r[0] = u[0]; r[1] = u[1]; r[2] = u[2]; r[3] = w[0]; r[4] = w[1]; r[5] = w[2];
        return r;
    }

private:
    Eigen::Matrix3d R_;
    Eigen::Vector3d t_;
};
typedef SE3d SO3d;  // only named in a typedef by the reference (library_include.hpp:19)

}  // namespace Sophus

// ----------------------------------------------------------------------------------------------------------------
namespace cv {

template <typename T>
using Ptr = std::shared_ptr<T>;

struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float a, float b) : x(a), y(b) {}
    Point2f operator-(const Point2f& o) const { return Point2f(x - o.x, y - o.y); }
};
inline double norm(const Point2f& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

struct Point3f {
    float x = 0, y = 0, z = 0;
    Point3f() {}
    Point3f(float a, float b, float c) : x(a), y(b), z(c) {}
};

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0;
};
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");

enum { CV_8U = 0, CV_32F = 5, CV_64F = 6 };

// 2-D, single-channel, reference-counted matrix (the subset of cv::Mat the hot path touches)
class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        type_ = type;
        rows = r; cols = c;
        step = (size_t)c * elemSize();
        const size_t bytes = std::max<size_t>(step * r, 1);
        // image-sized buffers come from the library's pinned pool when the drop-in layer installed it (VO's
        // constructor): frames and disparity images then move over PCIe as plain DMAs
        uint8_t* pinned = (bytes >= kPinnedMin && allocator().alloc) ? static_cast<uint8_t*>(allocator().alloc(bytes)) : nullptr;
        if (pinned) {
            void (*fr)(void*) = allocator().free;
            buf_ = std::shared_ptr<uint8_t>(pinned, [fr](uint8_t* p) { fr(p); });
        } else {
            buf_ = std::shared_ptr<uint8_t>(new uint8_t[bytes], std::default_delete<uint8_t[]>());
        }
        data = buf_.get();
    }
    struct Allocator {
        void* (*alloc)(size_t) = nullptr;
        void (*free)(void*) = nullptr;
    };
    static Allocator& allocator() { static Allocator a; return a; }
    static constexpr size_t kPinnedMin = 256 * 1024;
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_8U ? 1 : type_ == CV_32F ? 4 : 8; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T* ptr(int r) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    // row view sharing the buffer (Feature::descriptor_ holds such views, visual_odometry.cpp:513,593)
    Mat row(int r) const {
        Mat m;
        m.type_ = type_; m.rows = 1; m.cols = cols; m.step = step; m.buf_ = buf_;
        m.data = data + (size_t)r * step;
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step, data + r * step, (size_t)cols * elemSize());
        return m;
    }
    // append rows of `m` (cv::Mat::push_back(const Mat&), used row by row at visual_odometry.cpp:197,572)
    void push_back(const Mat& m) {
        if (m.empty()) return;
        if (empty()) { *this = m.clone(); return; }
        Mat n(rows + m.rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(n.data + r * n.step, data + r * step, (size_t)cols * elemSize());
        for (int r = 0; r < m.rows; ++r) std::memcpy(n.data + (rows + r) * n.step, m.data + r * m.step, (size_t)cols * elemSize());
        *this = n;
    }
    void setTo(double v) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (type_ == CV_8U) at<uint8_t>(r, c) = (uint8_t)v;
                else if (type_ == CV_32F) at<float>(r, c) = (float)v;
                else at<double>(r, c) = v;
            }
    }

protected:
    int type_ = CV_8U;
    std::shared_ptr<uint8_t> buf_;
};

// (cv::Mat_<double>(3, 3) << a, b, c, ...) as used at visual_odometry.cpp:272 and run_vslam.cpp:36
template <typename T>
class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : sizeof(T) == 4 ? CV_32F : CV_8U) {}
    struct Init {
        Mat_* m;
        int i;
        Init& operator,(T v) { m->template ptr<T>(0)[i++] = v; return *this; }
        operator Mat() const { return *m; }
    };
    Init operator<<(T v) { this->template ptr<T>(0)[0] = v; return Init{this, 1}; }
};

}  // namespace cv

// ----------------------------------------------------------------------------------------------------------------
namespace ros {

// parameter-server stub (the reference reads /dataset, /if_write_pose, /if_rviz from the ROS parameter server,
// run_vslam.cpp:20-28).  Values come from setParam() on the handle, or from the process-wide store that ros::init()
// fills from "<name>:=<value>" arguments (e.g. /dataset:=/data/kitti/00/ /if_write_pose:=true) and from the
// environment (VSLAM_PARAM_dataset, VSLAM_PARAM_if_write_pose, VSLAM_PARAM_if_rviz).
struct ParamStore {
    std::map<std::string, bool> b;
    std::map<std::string, std::string> s;
    static ParamStore& global() { static ParamStore g; return g; }
    static bool parse_bool(const std::string& v) { return v == "true" || v == "1" || v == "True"; }
    void set(const std::string& key, const std::string& value) {
        s[key] = value;
        b[key] = parse_bool(value);
    }
};
class NodeHandle {
public:
    void setParam(const std::string& k, bool v) { own_.b[k] = v; }
    void setParam(const std::string& k, const std::string& v) { own_.s[k] = v; }
    bool getParam(const std::string& k, bool& v) const {
        for (const ParamStore* p : {&own_, (const ParamStore*)&ParamStore::global()}) {
            auto i = p->b.find(k);
            if (i != p->b.end()) { v = i->second; return true; }
        }
        v = false;
        return false;
    }
    bool getParam(const std::string& k, std::string& v) const {
        for (const ParamStore* p : {&own_, (const ParamStore*)&ParamStore::global()}) {
            auto i = p->s.find(k);
            if (i != p->s.end()) { v = i->second; return true; }
        }
        return false;
    }
private:
    ParamStore own_;
};
// ros::init / ros::spin / ros::spinOnce (run_vslam.cpp:19,89; visual_odometry.cpp:458,703): ROS plumbing, no-ops here
// apart from the parameter arguments described above
inline void init(int& argc, char** argv, const std::string& /*node_name*/) {
    ParamStore& g = ParamStore::global();
    for (const char* key : {"dataset", "if_write_pose", "if_rviz"}) {
        const char* e = std::getenv((std::string("VSLAM_PARAM_") + key).c_str());
        if (e) g.set(std::string("/") + key, e);
    }
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const size_t p = a.find(":=");
        if (p == std::string::npos) continue;
        std::string key = a.substr(0, p);
        if (!key.empty() && key[0] == '_') key = key.substr(1);
        if (key.empty() || key[0] != '/') key = "/" + key;
        g.set(key, a.substr(p + 2));
    }
}
inline void spin() {}
inline void spinOnce() {}

}  // namespace ros

#endif  // VSLAM_USE_REAL_DEPS
