// LidarCornersEst.hpp -- header-only C++ adaptor over the C ABI (ilcc2_b200.h) with the public surface of the
// reference class ilcc2/include/ilcc2/LidarCornersEst.h:12-84 as driven by ilcc2/test/get_lidar_corners.cpp:183-198.
// A node that today does
//     lidar_corners_est->setROI(pointcloud.makeShared(), click_point);
//     if (lidar_corners_est->EuclideanCluster()) { lidar_corners_est->PCA(); lidar_corners_est->get_corners(v); ... }
// keeps compiling against this class; the PCL viewer key presses become the public `keys` member.
// Point type: ilcc2::PointXYZI has the byte layout of pcl::PointXYZI (x,y,z @0/4/8, intensity @16, 32 bytes), so a
// pcl::PointCloud<pcl::PointXYZI>::points buffer or a velodyne PointCloud2 `data` buffer is passed without a copy.
#pragma once

#include <array>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ilcc2_b200.h"

namespace ilcc2 {

struct alignas(16) PointXYZI {
  float x, y, z, pad0;
  float intensity, pad1[3];
};
static_assert(sizeof(PointXYZI) == 32, "must match pcl::PointXYZI");

struct PointCloud {
  typedef std::shared_ptr<PointCloud> Ptr;
  std::vector<PointXYZI> points;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); }
  void push_back(const PointXYZI& p) { points.push_back(p); }
  Ptr makeShared() const { return Ptr(new PointCloud(*this)); }
};

typedef PointXYZI myPoint;       // ilcc2/include/ilcc2/config.h:8-9
typedef PointCloud myPointCloud;

// What the user would press in the two viewer windows (ilcc2/src/Visualization.cpp:20-57).
struct KeyPresses {
  int plane_index_delta = 0;   // net 'w' (+1) / 's' (-1) presses before 'o'
  bool topleft_white = false;  // one press of 'd'
  bool inverse = false;        // one press of 'a'
  bool auto_colour = false;    // unattended: pick the colour phase with the lower cost (not in the reference)
};

class LidarCornersEst {
 public:
  typedef std::shared_ptr<LidarCornersEst> Ptr;

  explicit LidarCornersEst(int device = 0, int max_points = 150000) : device_(device), max_points_(max_points) {
    ilcc2_default_params(&params_);
    m_cloud_ROI.reset(new myPointCloud);
    m_cloud_chessboard.reset(new myPointCloud);
    m_cloud_PCA.reset(new myPointCloud);
    m_cloud_optim.reset(new myPointCloud);
    m_cloud_corners.reset(new myPointCloud);
  }
  ~LidarCornersEst() {
    if (h_) ilcc2_destroy(h_);
    if (h_online_) ilcc2_destroy(h_online_);
  }
  LidarCornersEst(const LidarCornersEst&) = delete;
  LidarCornersEst& operator=(const LidarCornersEst&) = delete;

  void register_viewer() {}  // LidarCornersEst.h:26-37: no GUI here; see `keys`

  // LidarCornersEst.cpp:20-46
  bool set_chessboard_param(std::string cam_yaml) {
    if (ilcc2_set_chessboard_param(&params_, cam_yaml.c_str()) != ILCC2_SUCCESS) return false;
    if (h_) {
      ilcc2_destroy(h_);
      h_ = nullptr;
    }
    if (h_online_) {
      ilcc2_destroy(h_online_);
      h_online_ = nullptr;
    }
    return true;
  }

  // LidarCornersEst.cpp:48-70 (the ROI filter itself runs on the GPU together with the rest of the path)
  void setROI(myPointCloud::Ptr cloud, myPoint point) {
    cloud_ = cloud;
    m_click_point = point;
    ran_ = false;
  }

  // LidarCornersEst.cpp:124-186: true when a chessboard plane was found and "confirmed"
  bool EuclideanCluster() {
    run();
    return status_ != ILCC2_NO_POINTS_IN_ROI && status_ != ILCC2_NO_CLUSTER && status_ != ILCC2_PLANE_TOO_SMALL &&
           status_ != ILCC2_RANSAC_FAILED;
  }

  // LidarCornersEst.cpp:366-372
  void PCA() {
    if (!ran_) run();
    m_cloud_PCA->points.resize(m_cloud_chessboard->size());
    const float* T = diag_.pca;
    for (size_t i = 0; i < m_cloud_chessboard->size(); ++i) {
      const myPoint& p = m_cloud_chessboard->points[i];
      myPoint q = p;
      q.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
      q.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
      q.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
      m_cloud_PCA->points[i] = q;
    }
  }

  // LidarCornersEst.cpp:374-450
  bool get_corners(std::vector<std::array<double, 3>>& corners) {
    if (!ran_) run();
    if (status_ != ILCC2_OK) return false;
    m_cloud_corners->points.clear();
    for (size_t i = 0; i < corners_.size() / 3; ++i) {
      myPoint p{};
      p.x = corners_[3 * i];
      p.y = corners_[3 * i + 1];
      p.z = corners_[3 * i + 2];
      p.intensity = 50;
      m_cloud_corners->push_back(p);
      corners.push_back({{p.x, p.y, p.z}});
    }
    // m_cloud_optim = m_cloud_PCA moved by the optimised in-plane pose (LidarCornersEst.cpp:412-413)
    PCA();
    const double* th = diag_.theta_t[diag_.hyp_chosen];
    const float c = std::cos((float)th[0]), s = std::sin((float)th[0]);
    m_cloud_optim->points = m_cloud_PCA->points;
    for (myPoint& p : m_cloud_optim->points) {
      const float y = p.y, z = p.z;
      p.y = c * y - s * z + (float)th[1];
      p.z = s * y + c * z + (float)th[2];
    }
    return true;
  }

  // LidarCornersEst.cpp:72-115, the streaming variant driven by ilcc2/test/lidar_chessboard_online.cpp:91-99: no
  // box ROI (the whole cloud is clustered), cluster tolerance 0.10, the plane of the cluster that holds the nearest
  // neighbour of `point`; false when that plane has fewer than 500 points or the neighbour lies in no kept cluster
  // (find_board == false, :111-112) -- outcloud then still holds the plane of the largest cluster, as in the
  // reference.  Deviation: points farther than online_extent (default 25 m per axis) from `point` are ignored (the
  // cell hash of the clustering kernel spans at most 1023 cells of 0.055 m per axis).
  bool get_chessboard_by_point(myPointCloud::Ptr incloud, myPointCloud::Ptr& outcloud, myPoint point) {
    if (!incloud) throw std::logic_error("get_chessboard_by_point: null cloud");
    if (!h_online_) {
      ilcc2_params p = params_;
      p.roi_half_x = p.roi_half_y = p.roi_half_z = online_extent;
      p.cluster_tolerance = 0.1f;   // :79
      p.min_board_points = 500;     // :111
      if (ilcc2_create(&p, device_, 1, max_points_, &h_online_) != ILCC2_SUCCESS)
        throw std::runtime_error(std::string("ilcc2_create: ") + ilcc2_last_error(nullptr));
    }
    const int ncorn = (params_.cells_x - 1) * (params_.cells_y - 1);
    std::vector<float> corners((size_t)ncorn * 3);
    const float click[3] = {point.x, point.y, point.z};
    int32_t st = 0;
    ilcc2_diag dg{};
    const int rc = ilcc2_extract_corners(h_online_, incloud->points.data(), 1, (int)incloud->size(), nullptr, 32, 0, 4, 8, 16,
                                         ILCC2_MEM_HOST, click, nullptr, ILCC2_FLAG_AUTO, nullptr, corners.data(), &st, &dg,
                                         nullptr);
    if (rc != ILCC2_SUCCESS) throw std::runtime_error(std::string("ilcc2_extract_corners: ") + ilcc2_last_error(h_online_));
    online_diag_ = dg;
    outcloud.reset(new myPointCloud);
    std::vector<int32_t> idx(incloud->size());
    const int n = ilcc2_fetch_indices(h_online_, 0, 2, idx.data(), (int)idx.size());
    for (int i = 0; i < n && i < (int)idx.size(); ++i) outcloud->points.push_back(incloud->points[idx[i]]);
    const bool no_plane = st == ILCC2_NO_POINTS_IN_ROI || st == ILCC2_NO_CLUSTER || st == ILCC2_RANSAC_FAILED ||
                          st == ILCC2_PLANE_TOO_SMALL || st == ILCC2_ROI_CAPACITY;
    return !no_plane && outcloud->size() >= 500 && dg.click_in_cluster != 0;
  }
  const ilcc2_diag& online_diag() const { return online_diag_; }
  float online_extent = 25.0f;

  int status() const { return status_; }
  const ilcc2_diag& diag() const { return diag_; }
  ilcc2_params& params() { return params_; }
  double m_grid_length() const { return params_.grid_length; }

  KeyPresses keys;
  myPoint m_click_point{};
  myPointCloud::Ptr m_cloud_ROI, m_cloud_chessboard, m_cloud_PCA, m_cloud_optim, m_cloud_corners;

 private:
  void run() {
    if (!cloud_) throw std::logic_error("LidarCornersEst: setROI() must be called first");
    if (!h_ && ilcc2_create(&params_, device_, 1, max_points_, &h_) != ILCC2_SUCCESS)
      throw std::runtime_error(std::string("ilcc2_create: ") + ilcc2_last_error(nullptr));
    const int n = (int)cloud_->size();
    const int ncorn = (params_.cells_x - 1) * (params_.cells_y - 1);
    corners_.assign((size_t)ncorn * 3, 0.f);
    const float click[3] = {m_click_point.x, m_click_point.y, m_click_point.z};
    const uint8_t flags = (uint8_t)((keys.topleft_white ? ILCC2_FLAG_TOPLEFT_WHITE : 0) |
                                    (keys.inverse ? ILCC2_FLAG_INVERSE : 0) | (keys.auto_colour ? ILCC2_FLAG_AUTO : 0));
    const int8_t delta = (int8_t)keys.plane_index_delta;
    const int rc = ilcc2_extract_corners(h_, cloud_->points.data(), 1, n, nullptr, 32, 0, 4, 8, 16, ILCC2_MEM_HOST, click,
                                         &flags, 0, &delta, corners_.data(), &status_, &diag_, nullptr);
    if (rc != ILCC2_SUCCESS) throw std::runtime_error(std::string("ilcc2_extract_corners: ") + ilcc2_last_error(h_));
    fetch(0, *m_cloud_ROI);
    fetch(2, *m_cloud_chessboard);
    ran_ = true;
  }
  void fetch(int which, myPointCloud& out) {
    std::vector<int32_t> idx(cloud_->size());
    const int n = ilcc2_fetch_indices(h_, 0, which, idx.data(), (int)idx.size());
    out.points.clear();
    for (int i = 0; i < n && i < (int)idx.size(); ++i) out.points.push_back(cloud_->points[idx[i]]);
  }

  int device_, max_points_;
  ilcc2_params params_;
  ilcc2_handle* h_ = nullptr;
  ilcc2_handle* h_online_ = nullptr;
  ilcc2_diag online_diag_{};
  myPointCloud::Ptr cloud_;
  bool ran_ = false;
  int32_t status_ = ILCC2_NO_POINTS_IN_ROI;
  ilcc2_diag diag_{};
  std::vector<float> corners_;
};

}  // namespace ilcc2
