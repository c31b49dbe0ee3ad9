// kitti_pipeline.cpp — minimal driver over host/lvo_handlers.hpp: reads KITTI-odometry velodyne sweeps (N x 4 float32
// records x, y, z, reflectance — the layout read by the reference's src/kittiHelper.cpp:25-35), runs the three stages
// the way the three ROS nodes chain them (mapping_skip_frame = 1), and writes the mapped poses in KITTI's 3x4 row-major
// format (one line of 12 numbers per frame, as src/kittiHelper.cpp:95-111 parses the ground truth).
//
//   kitti_pipeline <n_scans: 16|32|64> <out_poses.txt> [--distortion 0|1|2] [--map-out map.bin] <sweep0.bin> [sweep1.bin ...]
//
// --distortion selects laserOdometry.cpp:67 at run time (lvo_config::distortion); --map-out writes the whole-map cloud
// (laserMapping.cpp:823-836, N x 4 float32) every time the reference would publish it (every 20th frame), last one wins.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "lvo_handlers.hpp"

struct PointXYZI {  // memory layout of pcl::PointXYZI (32 bytes, intensity at 16)
  float x, y, z, pad0;
  float intensity, pad1, pad2, pad3;
};
struct Cloud { std::vector<PointXYZI> points; };

static bool read_bin(const char* path, Cloud& c) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  const size_t n = (size_t)f.tellg() / (4 * sizeof(float));
  f.seekg(0);
  std::vector<float> raw(n * 4);
  f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)(raw.size() * sizeof(float)));
  c.points.resize(n);
  for (size_t i = 0; i < n; ++i) { c.points[i] = PointXYZI{raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], 1.f, raw[4 * i + 3], 0, 0, 0}; }
  return true;
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <n_scans> <out_poses.txt> <sweep.bin>...\n", argv[0]); return 2; }
  try {
    const int n_scans = std::atoi(argv[1]);
    lvo_config cfg = n_scans == 16 ? lvo::Context::vlp16() : lvo::Context::hdl64();
    cfg.n_scans = n_scans;
    cfg.max_map_corner = 1 << 18; cfg.max_map_surf = 1 << 19;
    int first = 3;
    const char* map_out = nullptr;
    while (first + 1 < argc && argv[first][0] == '-' && argv[first][1] == '-') {
      const std::string opt = argv[first];
      if (opt == "--distortion") cfg.distortion = std::atoi(argv[first + 1]);
      else if (opt == "--map-out") map_out = argv[first + 1];
      else { std::fprintf(stderr, "unknown option %s\n", argv[first]); return 2; }
      first += 2;
    }
    lvo::Context ctx(cfg);
    lvo::ScanRegistration reg(ctx);
    lvo::LaserOdometry odo(ctx);
    lvo::LaserMapping map(ctx);
    std::FILE* out = std::fopen(argv[2], "w");
    if (!out) { std::perror(argv[2]); return 1; }
    Cloud in, full, sharp, lessSharp, flat, lessFlat, surround, whole;
    for (int k = first; k < argc; ++k) {
      if (!read_bin(argv[k], in)) { std::fprintf(stderr, "cannot read %s\n", argv[k]); return 1; }
      reg.laserCloudHandler(in, full, sharp, lessSharp, flat, lessFlat);
      lvo::Pose rel, wodom, wmap;
      odo.process(sharp, lessSharp, flat, lessFlat, rel, wodom);
      if (cfg.distortion == 2) odo.publishedClouds(lessSharp, lessFlat, full);
      map.process(lessSharp, lessFlat, full, wodom, wmap);
      if (map.surroundDue()) map.laserCloudSurround(surround);
      if (map.mapDue()) {
        map.laserCloudMap(whole);
        if (map_out) {
          std::ofstream mf(map_out, std::ios::binary);
          for (const PointXYZI& p : whole.points) { const float r[4] = {p.x, p.y, p.z, p.intensity}; mf.write(reinterpret_cast<const char*>(r), sizeof(r)); }
        }
      }
      const double x = wmap.q[0], y = wmap.q[1], z = wmap.q[2], w = wmap.q[3];
      const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
      std::fprintf(out, "%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", R[0], R[1], R[2], wmap.t[0], R[3], R[4], R[5], wmap.t[1], R[6],
                   R[7], R[8], wmap.t[2]);
      std::fprintf(out, "# q %.17g %.17g %.17g %.17g t %.17g %.17g %.17g\n", x, y, z, w, wmap.t[0], wmap.t[1], wmap.t[2]);
    }
    std::fclose(out);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "kitti_pipeline: %s\n", e.what());
    return 1;
  }
  return 0;
}
