// calib_lidar_cam -- host tool with the input/output contract of the reference node
// ilcc2/test/calib_lidar_cam.cpp:73-190, without ROS / OpenCV windows: for bag_idx = 1..bag_num it reads
// <package>/process_data/<camera_name>_lidar_<bag_idx>.txt and <package>/process_data/<camera_name><bag_idx>.txt
// (:103-117), brings the 3-D corners into the rough camera axes and normalises both orders (:119-123), solves
// Optimization::solvePose3d2dError from r = t = 0 (:149-152) and writes <package>/config/<pose_file> (:165).
// Ordering and the solve run on the GPU through the C ABI (ilcc2_pnp_*); this file only moves text and bytes.
//
//   calib_lidar_cam --package /path/to/ilcc2 --bag-num 6 --yaml pointgrey.yaml --camera pointgrey
//                   --pose-file pointgrey.bin [--device 0]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ilcc2_b200.h"

int main(int argc, char** argv) {
  std::string package = ".", yaml = "front.yaml", camera = "front", pose_file = "front.bin";  // defaults of :82-87
  int bag_num = 1, device = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--package") package = v;
    else if (k == "--bag-num") bag_num = atoi(v.c_str());
    else if (k == "--yaml") yaml = v;
    else if (k == "--camera") camera = v;
    else if (k == "--pose-file") pose_file = v;
    else if (k == "--device") device = atoi(v.c_str());
    else {
      fprintf(stderr, "unknown option %s\n", k.c_str());
      return 2;
    }
  }
  double K[9];
  int bw = 0, bh = 0;
  const std::string yaml_path = package + "/config/" + yaml;
  if (ilcc2_read_camera_yaml(yaml_path.c_str(), K, &bw, &bh) != ILCC2_SUCCESS || bw < 2 || bh < 2) {
    fprintf(stderr, "can not open %s\n", yaml_path.c_str());
    return 1;
  }
  const int nc = bw * bh;
  double T_axis[16];
  if (ilcc2_axis_roughly(camera.c_str(), T_axis) != ILCC2_SUCCESS) {
    fprintf(stderr, "no rough LiDAR->camera axes for camera_name %s\n", camera.c_str());
    return 1;
  }
  std::vector<float> lidar((size_t)bag_num * nc * 3);
  std::vector<double> cam((size_t)bag_num * nc * 2), tmp3((size_t)nc * 3);
  for (int b = 1; b <= bag_num; ++b) {
    const std::string lp = package + "/process_data/" + camera + "_lidar_" + std::to_string(b) + ".txt";
    const std::string cp = package + "/process_data/" + camera + std::to_string(b) + ".txt";
    printf("%s\n%s\n", lp.c_str(), cp.c_str());
    if (ilcc2_read_lidar_corners(lp.c_str(), nc, tmp3.data()) != nc ||
        ilcc2_read_cam_corners(cp.c_str(), nc, bh, cam.data() + (size_t)(b - 1) * nc * 2, nc) != nc) {
      fprintf(stderr, "corner files of bag %d are missing or short\n", b);
      return 1;
    }
    for (int i = 0; i < nc * 3; ++i) lidar[(size_t)(b - 1) * nc * 3 + i] = (float)tmp3[i];  // the reader parses floats
  }
  ilcc2_pnp* h = nullptr;
  if (ilcc2_pnp_create(device, &h) != ILCC2_SUCCESS) {
    fprintf(stderr, "ilcc2_pnp_create: %s\n", ilcc2_pnp_last_error(nullptr));
    return 1;
  }
  const size_t M = (size_t)bag_num * nc;
  std::vector<double> p3(M * 3), p2(M * 2);
  std::vector<uint8_t> valid(M);
  int rc = ilcc2_pnp_order_corners(h, lidar.data(), nullptr, cam.data(), bag_num, bw, bh, T_axis, p3.data(), p2.data(),
                                   valid.data(), ILCC2_MEM_HOST, nullptr);
  double r[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  const double camera4[4] = {K[0], K[2], K[4], K[5]};
  ilcc2_pnp_summary sm;
  if (rc == ILCC2_SUCCESS)
    rc = ilcc2_pnp_solve(h, p3.data(), p2.data(), valid.data(), (int)M, camera4, 0.1, 50, r, t, &sm, ILCC2_MEM_HOST, nullptr);
  if (rc != ILCC2_SUCCESS) {
    fprintf(stderr, "pnp failed: %s\n", ilcc2_pnp_last_error(h));
    ilcc2_pnp_destroy(h);
    return 1;
  }
  printf("%zu\n%zu\n%d\n", M, M, bag_num);
  printf("iterations: %d, initial cost: %.6e, final cost: %.6e, termination: %d\n", sm.iterations, sm.initial_cost,
         sm.final_cost, sm.termination);
  printf("r_ceres: %g %g %g\nt_ceres: %g %g %g\n", r[0], r[1], r[2], t[0], t[1], t[2]);
  double T[16];
  ilcc2_compose_extrinsic(r, t, T_axis, T);
  printf("lidar2cam:\n");
  for (int i = 0; i < 4; ++i) printf("%g %g %g %g\n", T[4 * i], T[4 * i + 1], T[4 * i + 2], T[4 * i + 3]);
  const std::string out = package + "/config/" + pose_file;
  if (ilcc2_extrinsic2bin(out.c_str(), T) != ILCC2_SUCCESS) {
    fprintf(stderr, "can not write %s\n", out.c_str());
    ilcc2_pnp_destroy(h);
    return 1;
  }
  ilcc2_pnp_destroy(h);
  return 0;
}
