// get_lidar_corners -- host tool with the input/output contract of the reference node
// ilcc2/test/get_lidar_corners.cpp:89-216, without ROS: for bag_idx = 1..bag_num it opens
// <bag_path_prefix><bag_idx>.bag, takes the FIRST message on the LiDAR topic (:144-154), runs the corner path on
// the GPU through the LidarCornersEst adaptor, and writes <out_dir>/<camera_name>_lidar_<bag_idx>.txt (:197-198).
// The /clicked_point messages of the reference become a text file with one "x y z [keys]" line per bag, where the
// optional keys are the viewer key presses: d (topleftWhite), a (inverse), w/s (plane index +-1), A (auto colour).
//
//   get_lidar_corners --bag-prefix /data/2018-12-03- --bag-num 6 --yaml pointgrey.yaml --camera pointgrey
//                     --clicks clicks.txt --out-dir process_data [--topic /velodyne_points] [--device 0]
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "ilcc2_b200/LidarCornersEst.hpp"

namespace {

// ---- minimal ROS bag v2.0 reader: uncompressed chunks, record ops 2 (message), 5 (chunk), 7 (connection) ----
struct Reader {
  const std::vector<unsigned char>& b;
  size_t pos;
  bool ok(size_t n) const { return pos + n <= b.size(); }
  uint32_t u32() {
    uint32_t v;
    memcpy(&v, &b[pos], 4);
    pos += 4;
    return v;
  }
};

typedef std::map<std::string, std::string> Header;

Header parse_header(const unsigned char* p, size_t len) {
  Header h;
  size_t pos = 0;
  while (pos + 4 <= len) {
    uint32_t flen;
    memcpy(&flen, p + pos, 4);
    pos += 4;
    if (pos + flen > len) break;
    const std::string field(reinterpret_cast<const char*>(p + pos), flen);
    pos += flen;
    const size_t eq = field.find('=');
    if (eq != std::string::npos) h[field.substr(0, eq)] = field.substr(eq + 1);
  }
  return h;
}

uint32_t le32(const std::string& s) {
  uint32_t v = 0;
  if (s.size() >= 4) memcpy(&v, s.data(), 4);
  return v;
}

// Serialized sensor_msgs/PointCloud2 -> 32-byte XYZI records (what pcl::fromROSMsg does by field name, :163-164)
bool decode_cloud(const unsigned char* m, size_t len, ilcc2::PointCloud& out) {
  size_t pos = 0;
  auto u32 = [&]() {
    uint32_t v = 0;
    if (pos + 4 <= len) memcpy(&v, m + pos, 4);
    pos += 4;
    return v;
  };
  auto str = [&]() {
    const uint32_t n = u32();
    std::string s;
    if (pos + n <= len) s.assign(reinterpret_cast<const char*>(m + pos), n);
    pos += n;
    return s;
  };
  u32(); u32(); u32();
  str();
  const uint32_t height = u32(), width = u32(), nfields = u32();
  int off[4] = {-1, -1, -1, -1};
  for (uint32_t i = 0; i < nfields; ++i) {
    const std::string name = str();
    const uint32_t o = u32();
    const unsigned char dt = pos < len ? m[pos] : 0;
    pos += 1;
    u32();
    if (dt != 7) continue;  // FLOAT32
    if (name == "x") off[0] = (int)o;
    if (name == "y") off[1] = (int)o;
    if (name == "z") off[2] = (int)o;
    if (name == "intensity") off[3] = (int)o;
  }
  pos += 1;  // is_bigendian
  const uint32_t point_step = u32();
  u32();  // row_step
  const uint32_t dlen = u32();
  if (off[0] < 0 || off[1] < 0 || off[2] < 0 || off[3] < 0 || pos + dlen > len) return false;
  const size_t n = (size_t)height * width;
  if ((size_t)dlen < n * point_step) return false;
  out.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const unsigned char* r = m + pos + i * point_step;
    ilcc2::PointXYZI p{};
    memcpy(&p.x, r + off[0], 4);
    memcpy(&p.y, r + off[1], 4);
    memcpy(&p.z, r + off[2], 4);
    memcpy(&p.intensity, r + off[3], 4);
    out.points[i] = p;
  }
  return true;
}

bool first_cloud(const std::string& path, const std::string& topic, ilcc2::PointCloud& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const char magic[] = "#ROSBAG V2.0\n";
  if (buf.size() < 13 || memcmp(buf.data(), magic, 13) != 0) return false;
  std::map<uint32_t, std::string> conns;
  Reader r{buf, 13};
  while (r.ok(4)) {
    const uint32_t hlen = r.u32();
    if (!r.ok(hlen)) break;
    const Header h = parse_header(&buf[r.pos], hlen);
    r.pos += hlen;
    if (!r.ok(4)) break;
    const uint32_t dlen = r.u32();
    if (!r.ok(dlen)) break;
    const size_t dpos = r.pos;
    r.pos += dlen;
    const auto op = h.find("op");
    if (op == h.end() || op->second.empty()) continue;
    if ((unsigned char)op->second[0] == 7) conns[le32(h.at("conn"))] = h.at("topic");
    if ((unsigned char)op->second[0] != 5) continue;
    if (h.count("compression") && h.at("compression") != "none") {
      std::cerr << path << ": compressed chunks are not supported\n";
      return false;
    }
    size_t p = dpos;
    const size_t end = dpos + dlen;
    while (p + 4 <= end) {
      uint32_t hl;
      memcpy(&hl, &buf[p], 4);
      p += 4;
      if (p + hl + 4 > end) break;
      const Header h2 = parse_header(&buf[p], hl);
      p += hl;
      uint32_t dl;
      memcpy(&dl, &buf[p], 4);
      p += 4;
      if (p + dl > end) break;
      const unsigned char op2 = h2.count("op") && !h2.at("op").empty() ? (unsigned char)h2.at("op")[0] : 0xff;
      if (op2 == 7) conns[le32(h2.at("conn"))] = h2.at("topic");
      if (op2 == 2 && conns[le32(h2.at("conn"))] == topic) return decode_cloud(&buf[p], dl, out);
      p += dl;
    }
  }
  return false;
}

}  // namespace

int main(int argc, char** argv) {
  std::map<std::string, std::string> a = {{"--topic", "/velodyne_points"}, {"--camera", "front"}, {"--bag-num", "1"},
                                          {"--device", "0"},               {"--out-dir", "."},
                                          {"--online", "0"}};
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  if (!a.count("--bag-prefix") || !a.count("--yaml") || !a.count("--clicks")) {
    std::cerr << "usage: get_lidar_corners --bag-prefix P --bag-num N --yaml Y --camera C --clicks F --out-dir D\n";
    return 2;
  }
  ilcc2::LidarCornersEst est(atoi(a["--device"].c_str()));
  if (!est.set_chessboard_param(a["--yaml"])) {
    std::cerr << "can not open " << a["--yaml"] << std::endl;
    return 1;
  }
  std::ifstream clicks(a["--clicks"]);
  const int bag_num = atoi(a["--bag-num"].c_str());
  int written = 0;
  for (int bag_idx = 1; bag_idx <= bag_num; ++bag_idx) {
    std::string line;
    if (!std::getline(clicks, line)) break;
    std::istringstream ls(line);
    ilcc2::myPoint click{};
    std::string keys;
    if (!(ls >> click.x >> click.y >> click.z)) continue;  // empty line: skip this bag
    ls >> keys;
    est.keys = ilcc2::KeyPresses();
    for (char k : keys) {
      if (k == 'd') est.keys.topleft_white = !est.keys.topleft_white;
      if (k == 'a') est.keys.inverse = !est.keys.inverse;
      if (k == 'w') est.keys.plane_index_delta++;
      if (k == 's') est.keys.plane_index_delta--;
      if (k == 'A') est.keys.auto_colour = true;
    }
    const std::string bag_path = a["--bag-prefix"] + std::to_string(bag_idx) + ".bag";
    ilcc2::PointCloud cloud;
    if (!first_cloud(bag_path, a["--topic"], cloud)) {
      std::cerr << "can't read lidar topic in " << bag_path << std::endl;  // get_lidar_corners.cpp:157-161
      continue;
    }
    if (a["--online"] != "0") {  // lidar_chessboard_online.cpp:91-99: the plane of the clicked cluster, no corner fit
      ilcc2::myPointCloud::Ptr board;
      const bool found = est.get_chessboard_by_point(cloud.makeShared(), board, click);
      std::cout << bag_idx << ": online found " << (found ? 1 : 0) << " plane " << board->size() << " click_in_cluster "
                << est.online_diag().click_in_cluster << std::endl;
      ++written;
      continue;
    }
    est.setROI(cloud.makeShared(), click);
    if (!est.EuclideanCluster()) {
      std::cerr << bag_idx << ": no chessboard plane (status " << est.status() << ")\n";
      continue;
    }
    est.PCA();
    std::vector<std::array<double, 3>> corners;
    if (!est.get_corners(corners)) {
      std::cerr << bag_idx << ": reject this scan (status " << est.status() << ")\n";
      continue;
    }
    std::vector<float> xyz;
    for (const auto& c : corners) {
      xyz.push_back((float)c[0]);
      xyz.push_back((float)c[1]);
      xyz.push_back((float)c[2]);
    }
    const std::string save = a["--out-dir"] + "/" + a["--camera"] + "_lidar_" + std::to_string(bag_idx) + ".txt";
    if (ilcc2_save_corners2txt(xyz.data(), (int)corners.size(), save.c_str()) != ILCC2_SUCCESS) {
      std::cerr << "can not write " << save << std::endl;
      return 1;
    }
    std::cout << bag_idx << ": chessboard plane size " << est.m_cloud_chessboard->size() << ", "
              << corners.size() << " corners -> " << save << std::endl;
    ++written;
  }
  return written > 0 ? 0 : 1;
}
