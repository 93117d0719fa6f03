// Calibration.hpp -- header-only C++ adaptor over the C ABI (ilcc2_b200.h) for the calibration node
// ilcc2/test/calib_lidar_cam.cpp:73-190, with the names the node uses today:
//     ImageCornersEst::read_cam_corners / read_lidar_corners / check_order_cam / check_order_lidar / extrinsic2txt
//         (ilcc2/src/ImageCornersEst.cpp:213-306, 430-505), members camK / m_board_size of getRectifyParam (:15-31)
//     Optimization::solvePose3d2dError (ilcc2/src/Optimization.cpp:13-91)
//     get_lidar2cam_axis_roughly (calib_lidar_cam.cpp:50-70)
// so that the node's main() keeps its shape (Eigen / OpenCV types replaced by plain structs).  Ordering and the PnP
// run on the GPU (ilcc2_pnp_*); nothing here computes.
#pragma once

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "ilcc2_b200.h"

namespace ilcc2 {

typedef std::array<double, 3> Vector3d;
typedef std::array<double, 2> Point2d;
struct Size {
  int width = 0, height = 0;
};
struct Isometry3d {  // row-major 4x4
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};

inline Isometry3d get_lidar2cam_axis_roughly(const std::string& cam_name) {
  Isometry3d T;
  if (ilcc2_axis_roughly(cam_name.c_str(), T.m) != ILCC2_SUCCESS)
    throw std::invalid_argument("get_lidar2cam_axis_roughly: unknown camera name " + cam_name);
  return T;
}

inline Vector3d operator*(const Isometry3d& T, const Vector3d& p) {
  Vector3d o;
  for (int r = 0; r < 3; ++r) o[r] = ((T.m[4 * r] * p[0] + T.m[4 * r + 1] * p[1]) + T.m[4 * r + 2] * p[2]) + T.m[4 * r + 3];
  return o;
}

class ImageCornersEst {
 public:
  double camK[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // row-major 3x3
  Size m_board_size;                             // (corner_in_x, corner_in_y)

  ImageCornersEst() {}
  explicit ImageCornersEst(const std::string& cam_yaml, bool /*stereo*/ = false) {
    if (!getRectifyParam(cam_yaml)) throw std::runtime_error("can not open " + cam_yaml);
  }
  bool getRectifyParam(const std::string& cam_yaml) {
    return ilcc2_read_camera_yaml(cam_yaml.c_str(), camK, &m_board_size.width, &m_board_size.height) == ILCC2_SUCCESS;
  }
  void read_cam_corners(const std::string& filename, int num, std::vector<Point2d>& point2d) const {
    std::vector<double> buf(8 * (size_t)num);
    const int n = ilcc2_read_cam_corners(filename.c_str(), num, m_board_size.height, buf.data(), 4 * num);
    for (int i = 0; i < n && i < 4 * num; ++i) point2d.push_back({buf[2 * i], buf[2 * i + 1]});
  }
  void read_lidar_corners(const std::string& filename, int num, std::vector<Vector3d>& point3d) const {
    std::vector<double> buf(3 * (size_t)num);
    const int n = ilcc2_read_lidar_corners(filename.c_str(), num, buf.data());
    for (int i = 0; i < n; ++i) point3d.push_back({buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]});
  }
  void extrinsic2txt(const std::string& savefile, const Isometry3d& lidar2cam) const {
    if (ilcc2_extrinsic2bin(savefile.c_str(), lidar2cam.m) != ILCC2_SUCCESS) throw std::runtime_error("can not write " + savefile);
  }
};

// The data-ready loop (calib_lidar_cam.cpp:101-135) and the solve (:149-152) for all boards at once: lidar_corners /
// image_corners hold bag_num x (board.width * board.height) corners in file order; on return they are ordered like
// check_order_lidar / check_order_cam leave them (the 3-D ones in the rough camera axes), and (r, t) is the solution.
class Optimization {
 public:
  explicit Optimization(int device = 0) {
    if (ilcc2_pnp_create(device, &h_) != ILCC2_SUCCESS) throw std::runtime_error(ilcc2_pnp_last_error(nullptr));
  }
  ~Optimization() { ilcc2_pnp_destroy(h_); }
  Optimization(const Optimization&) = delete;
  Optimization& operator=(const Optimization&) = delete;

  void order_corners(std::vector<Vector3d>& lidar_corners, std::vector<Point2d>& image_corners, Size board,
                     const Isometry3d& T_axis_roughly) {
    const int nc = board.width * board.height, frames = nc > 0 ? (int)(lidar_corners.size() / nc) : 0;
    if (nc <= 0 || lidar_corners.size() != image_corners.size() || lidar_corners.size() != (size_t)frames * nc)
      throw std::invalid_argument("order_corners: lidar / image corner counts must be equal and a multiple of the board size");
    std::vector<float> lf(lidar_corners.size() * 3);
    for (size_t i = 0; i < lidar_corners.size(); ++i)
      for (int k = 0; k < 3; ++k) lf[3 * i + k] = (float)lidar_corners[i][k];  // the files hold floats
    std::vector<double> p3(lidar_corners.size() * 3), p2(image_corners.size() * 2);
    std::vector<uint8_t> valid(lidar_corners.size());
    check(ilcc2_pnp_order_corners(h_, lf.data(), nullptr, &image_corners[0][0], frames, board.width, board.height,
                                  T_axis_roughly.m, p3.data(), p2.data(), valid.data(), ILCC2_MEM_HOST, nullptr));
    for (size_t i = 0; i < lidar_corners.size(); ++i) {
      lidar_corners[i] = {p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]};
      image_corners[i] = {p2[2 * i], p2[2 * i + 1]};
    }
  }

  Isometry3d solvePose3d2dError(const std::vector<Vector3d>& pts3d, const std::vector<Point2d>& pts2d, const double K[9],
                                Vector3d& r_ceres, Vector3d& t_ceres) {
    const double cam[4] = {K[0], K[2], K[4], K[5]};
    check(ilcc2_pnp_solve(h_, &pts3d[0][0], &pts2d[0][0], nullptr, (int)pts3d.size(), cam, 0.1, 50, r_ceres.data(),
                          t_ceres.data(), &summary, ILCC2_MEM_HOST, nullptr));
    Isometry3d T, I;
    ilcc2_compose_extrinsic(r_ceres.data(), t_ceres.data(), I.m, T.m);
    return T;
  }

  ilcc2_pnp_summary summary{};  // what summary.BriefReport() printed

 private:
  void check(int rc) {
    if (rc != ILCC2_SUCCESS) throw std::runtime_error(ilcc2_pnp_last_error(h_));
  }
  ilcc2_pnp* h_ = nullptr;
};

inline Isometry3d operator*(const Isometry3d& A, const Isometry3d& B) {
  Isometry3d C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A.m[4 * i + k] * B.m[4 * k + j];
      C.m[4 * i + j] = s;
    }
  return C;
}

}  // namespace ilcc2
