// calib_lidar_cam_node -- the main() of ilcc2/test/calib_lidar_cam.cpp:73-165 written against the header-only adaptor
// include/ilcc2_b200/Calibration.hpp (same class and method names as the reference: ImageCornersEst, Optimization,
// get_lidar2cam_axis_roughly), to show that the node keeps its shape.  calib_lidar_cam.cpp is the same program on the
// bare C ABI.     calib_lidar_cam_node <package dir> [bag_num]
#include <cstdio>
#include <cstdlib>

#include "ilcc2_b200/Calibration.hpp"

int main(int argc, char** argv) {
  using namespace ilcc2;
  const std::string package_path = argc > 1 ? argv[1] : ".", camera_name = "pointgrey";
  const int bag_num = argc > 2 ? atoi(argv[2]) : 6;
  ImageCornersEst image_corners_est(package_path + "/config/pointgrey.yaml", true);
  const int corner_num = image_corners_est.m_board_size.width * image_corners_est.m_board_size.height;
  const Isometry3d T_axis = get_lidar2cam_axis_roughly(camera_name);
  std::vector<Vector3d> lidar_corners;
  std::vector<Point2d> image_corners;
  for (int bag_idx = 1; bag_idx <= bag_num; ++bag_idx) {
    image_corners_est.read_cam_corners(package_path + "/process_data/" + camera_name + std::to_string(bag_idx) + ".txt", corner_num, image_corners);
    image_corners_est.read_lidar_corners(package_path + "/process_data/" + camera_name + "_lidar_" + std::to_string(bag_idx) + ".txt", corner_num, lidar_corners);
  }
  Optimization optim;
  optim.order_corners(lidar_corners, image_corners, image_corners_est.m_board_size, T_axis);
  Vector3d r_ceres{0, 0, 0}, t_ceres{0, 0, 0};
  Isometry3d T_lidar2cam = optim.solvePose3d2dError(lidar_corners, image_corners, image_corners_est.camK, r_ceres, t_ceres);
  T_lidar2cam = T_lidar2cam * T_axis;
  image_corners_est.extrinsic2txt(package_path + "/config/pointgrey.bin", T_lidar2cam);
  std::printf("%d iterations\n", optim.summary.iterations);
  return 0;
}
