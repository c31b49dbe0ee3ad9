// oracle/capi.cpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).  extern "C" surface for ctypes.
#include "scan_registration.hpp"
#include "laser_odometry.hpp"
#include "laser_mapping.hpp"
#include "depth_assoc.hpp"
#include <chrono>

using namespace lvo_oracle;

namespace {
template <class T>
long copy_out(const std::vector<T>& v, void* out, long cap_elems) {
  long n = (long)v.size();
  if (out && cap_elems >= n && n > 0) memcpy(out, v.data(), n * sizeof(T));
  return n;
}
struct Pipeline {
  int n_scans;
  double min_range;
  ScanRegOut reg;
  LaserOdometry odo;
  LaserMapping map;
  Cloud registered;
  double ms_reg = 0, ms_odo = 0, ms_map = 0;
  int skip_frame = 1;   // mapping_skip_frame (laserOdometry.cpp:274)
  int frame_count = 0;  // frameCount of laserOdometry.cpp:307,643-645,669
  int last_mapped = 0;
};
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

int lvo_oracle_is_reference_build() {
#ifdef LVO_ORACLE_USE_REFERENCE
  return 1;
#else
  return 0;
#endif
}

float lvo_oracle_atanf(float x) { return lvo_atanf(x); }
float lvo_oracle_atan2f(float y, float x) { return lvo_atan2f(y, x); }

// ---- stand-alone operators --------------------------------------------------------------------------------
// voxel grid; returns n_out (out may be null to size). idx/order optional [n].
long lvo_oracle_voxel_grid(const Pt* in, long n, float leaf, Pt* out, long cap, int* idx, int* order) {
  Cloud c(in, in + n), o;
  std::vector<int> vi, vo;
  voxel_grid(c, leaf, o, &vi, &vo);
  if (idx && !vi.empty()) memcpy(idx, vi.data(), vi.size() * sizeof(int));
  if (order && !vo.empty()) memcpy(order, vo.data(), vo.size() * sizeof(int));
  copy_out(o, out, cap);
  return (long)o.size();
}

// K-NN with gate; ind/sq [nq][K]; rows with fewer than K points inside the gate are -1 / +inf.
// method: 0 brute force, 1 kd-tree
void lvo_oracle_knn(const Pt* cloud, long n, const Pt* q, long nq, int K, float max_sq, int method, int* ind, float* sq) {
  Cloud c(cloud, cloud + n);
  KdTree kd;
  if (method == 1) kd.build(c);
  Neighbor nb[8];
  for (long i = 0; i < nq; ++i) {
    int cnt = method == 1 ? kd.knn(q[i].x, q[i].y, q[i].z, K, nb) : knn_brute(c, q[i].x, q[i].y, q[i].z, K, nb);
    bool ok = cnt == K && (double)nb[K - 1].d < (double)max_sq;
    for (int k = 0; k < K; ++k) {
      ind[i * K + k] = ok ? nb[k].i : -1;
      sq[i * K + k] = ok ? nb[k].d : INFINITY;
    }
  }
}

// residual + local Jacobian of one factor: f = [type, c(3), a(3), b(3), m(3), d] as 14 doubles
int lvo_oracle_eval_factor(const double* f, const double* x, double* r, double* J) {
  Factor F;
  F.type = (int)f[0];
  F.c = Vec3{f[1], f[2], f[3]}; F.a = Vec3{f[4], f[5], f[6]}; F.b = Vec3{f[7], f[8], f[9]}; F.m = Vec3{f[10], f[11], f[12]};
  F.d = f[13];
  return eval_factor(F, x, r, J);
}
// same with the interpolation ratio s of LidarEdgeFactor / LidarPlaneFactor (DISTORTION 1)
int lvo_oracle_eval_factor_s(const double* f, double s, const double* x, double* r, double* J) {
  Factor F;
  F.type = (int)f[0];
  F.c = Vec3{f[1], f[2], f[3]}; F.a = Vec3{f[4], f[5], f[6]}; F.b = Vec3{f[7], f[8], f[9]}; F.m = Vec3{f[10], f[11], f[12]};
  F.d = f[13]; F.s = s;
  return eval_factor(F, x, r, J);
}
// TransformToStart / TransformToEnd (laserOdometry.cpp:154-191) of n points under (q_last_curr, t_last_curr) = qt[7]
void lvo_oracle_transform(const Pt* in, long n, const double* qt, int distortion, int to_end, Pt* out) {
  const Quat q{qt[0], qt[1], qt[2], qt[3]};
  const Vec3 t{qt[4], qt[5], qt[6]};
  for (long i = 0; i < n; ++i) {
    if (to_end) LaserOdometry::transform_to_end(q, t, distortion, in[i], out[i]);
    else LaserOdometry::transform_to_start(q, t, distortion, in[i], out[i]);
  }
}

// LM solve on nf factors (14 doubles each); x[7] in/out; trace rows of 10 doubles (x7,cost,radius,flags); returns rows
int lvo_oracle_solve(const double* f, long nf, double* x, int max_iters, double huber, double* trace, int trace_cap) {
  std::vector<Factor> F(nf);
  for (long i = 0; i < nf; ++i) {
    const double* p = f + i * 14;
    F[i].type = (int)p[0];
    F[i].c = Vec3{p[1], p[2], p[3]}; F[i].a = Vec3{p[4], p[5], p[6]}; F[i].b = Vec3{p[7], p[8], p[9]}; F[i].m = Vec3{p[10], p[11], p[12]};
    F[i].d = p[13];
  }
  LmOptions o; o.max_num_iterations = max_iters; o.huber = huber;
  std::vector<LmTraceRow> tr;
  solve(F, x, o, &tr);
  int n = 0;
  for (const LmTraceRow& r : tr) {
    if (n >= trace_cap) break;
    for (int k = 0; k < 7; ++k) trace[n * 10 + k] = r.x[k];
    trace[n * 10 + 7] = r.cost; trace[n * 10 + 8] = r.radius; trace[n * 10 + 9] = r.flags;
    ++n;
  }
  return (int)tr.size();
}

void lvo_oracle_sym_eigen3(const double* M, double* w, double* V) { sym_eigen3(M, w, V); }
void lvo_oracle_plane_fit5(const double* A, double* n) { plane_fit5(A, n); }

// depth association (config 5).  Returns the number of depth-cloud points; dc_out [cap] / src_out [cap] optional.
long lvo_oracle_depth(const Pt* sweep, long n, const float* extr12, const float* uv, long nkp, Pt* dc_out, int* src_out, long cap, float* depth, int* valid,
                      int* nn) {
  Cloud dc;
  std::vector<int> src;
  depth_cloud(sweep, (size_t)n, extr12, dc, &src);
  if (dc_out && cap >= (long)dc.size() && !dc.empty()) memcpy(dc_out, dc.data(), dc.size() * sizeof(Pt));
  if (src_out && cap >= (long)src.size() && !src.empty()) memcpy(src_out, src.data(), src.size() * sizeof(int));
  if (nkp > 0) depth_associate(dc, uv, (size_t)nkp, depth, valid, nn);
  return (long)dc.size();
}

// ---- the three stages as one stateful pipeline -------------------------------------------------------------
void* lvo_oracle_create(int n_scans, double min_range, double line_res, double plane_res, int outer_iters, int lm_iters,
                        double huber, int use_kdtree) {
  Pipeline* p = new Pipeline();
  p->n_scans = n_scans; p->min_range = min_range;
  p->odo.outer_iters = outer_iters; p->odo.lm.max_num_iterations = lm_iters; p->odo.lm.huber = huber; p->odo.use_kdtree = use_kdtree != 0;
  p->map.outer_iters = outer_iters; p->map.lm.max_num_iterations = lm_iters; p->map.lm.huber = huber; p->map.use_kdtree = use_kdtree != 0;
  p->map.lineRes = (float)line_res; p->map.planeRes = (float)plane_res;
  return p;
}
void lvo_oracle_destroy(void* h) { delete (Pipeline*)h; }
// 0 = DISTORTION 0 (laserOdometry.cpp:67 as shipped), 1 = DISTORTION 1 as written, 2 = DISTORTION 1 + the TransformToEnd block :610-625
void lvo_oracle_set_distortion(void* h, int mode) { ((Pipeline*)h)->odo.distortion = mode; }

int lvo_oracle_extract(void* h, const Pt* in, long n) {
  Pipeline* p = (Pipeline*)h;
  double t0 = now_ms();
  int r = scan_registration(in, (size_t)n, p->n_scans, p->min_range, p->reg);
  p->ms_reg = now_ms() - t0;
  return r;
}
// what: 0 full 1 sharp 2 lessSharp 3 flat 4 lessFlat (Pt) ; 10 curvature(float) ; 11 sortInd 12 picked 13 label 14 scanStart 15 scanEnd (int)
long lvo_oracle_extract_get(void* h, int what, void* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  switch (what) {
    case 0: return copy_out(p->reg.full, out, cap);
    case 1: return copy_out(p->reg.sharp, out, cap);
    case 2: return copy_out(p->reg.lessSharp, out, cap);
    case 3: return copy_out(p->reg.flat, out, cap);
    case 4: return copy_out(p->reg.lessFlat, out, cap);
    case 10: return copy_out(p->reg.curvature, out, cap);
    case 11: return copy_out(p->reg.sortInd, out, cap);
    case 12: return copy_out(p->reg.picked, out, cap);
    case 13: return copy_out(p->reg.label, out, cap);
    case 14: return copy_out(p->reg.scanStart, out, cap);
    case 15: return copy_out(p->reg.scanEnd, out, cap);
  }
  return -1;
}

// odometry on explicit feature clouds; pose_out = [para_q(4), para_t(3), q_w(4), t_w(3)]
int lvo_oracle_odometry(void* h, const Pt* sharp, long ns, const Pt* lsharp, long nls, const Pt* flat, long nf, const Pt* lflat,
                        long nlf, double* pose_out, int keep_log) {
  Pipeline* p = (Pipeline*)h;
  Cloud a(sharp, sharp + ns), b(lsharp, lsharp + nls), c(flat, flat + nf), d(lflat, lflat + nlf);
  double t0 = now_ms();
  int r = p->odo.process(a, b, c, d, keep_log != 0);
  p->ms_odo = now_ms() - t0;
  if (pose_out) {
    for (int k = 0; k < 4; ++k) pose_out[k] = p->odo.para_q[k];
    for (int k = 0; k < 3; ++k) pose_out[4 + k] = p->odo.para_t[k];
    pose_out[7] = p->odo.q_w_curr.x; pose_out[8] = p->odo.q_w_curr.y; pose_out[9] = p->odo.q_w_curr.z; pose_out[10] = p->odo.q_w_curr.w;
    pose_out[11] = p->odo.t_w_curr.x; pose_out[12] = p->odo.t_w_curr.y; pose_out[13] = p->odo.t_w_curr.z;
  }
  if (r == 0 && p->odo.few_corr) r = 2;
  return r;
}
// the "last" clouds after the frame (what :646-656 publishes): which 0 corner, 1 surf, 2 full (distortion 2 only)
long lvo_oracle_odometry_last(void* h, int which, Pt* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  return copy_out(which == 0 ? p->odo.laserCloudCornerLast : (which == 1 ? p->odo.laserCloudSurfLast : p->odo.fullOut), out, cap);
}
// what: 0 corner_corr(int) 1 plane_corr(int) 2 lm trace (double rows of 10) ; 3 counts [n_corner,n_plane,lm_iters] (int) ; 4 final cost (double[1])
long lvo_oracle_odometry_log(void* h, int outer, int what, void* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  if (outer < 0 || outer >= (int)p->odo.log.size()) return -1;
  const OdoOuterLog& lg = p->odo.log[outer];
  if (what == 0) return copy_out(lg.corner_corr, out, cap);
  if (what == 1) return copy_out(lg.plane_corr, out, cap);
  if (what == 2) {
    std::vector<double> v;
    for (const LmTraceRow& r : lg.lm) { for (int k = 0; k < 7; ++k) v.push_back(r.x[k]); v.push_back(r.cost); v.push_back(r.radius); v.push_back(r.flags); }
    return copy_out(v, out, cap);
  }
  if (what == 3) { std::vector<int> v{lg.n_corner, lg.n_plane, lg.lm_iters}; return copy_out(v, out, cap); }
  if (what == 4) { std::vector<double> v{lg.final_cost}; return copy_out(v, out, cap); }
  return -1;
}

// mapping; odom = [q(4), t(3)]; pose_out = [q_w_curr(4), t_w_curr(3), q_wmap_wodom(4), t_wmap_wodom(3)]
int lvo_oracle_mapping(void* h, const Pt* corner, long nc, const Pt* surf, long nsf, const Pt* full, long nfull, const double* odom,
                       double* pose_out, int keep_log) {
  Pipeline* p = (Pipeline*)h;
  Cloud a(corner, corner + nc), b(surf, surf + nsf), f;
  if (full) f.assign(full, full + nfull);
  double t0 = now_ms();
  p->map.process(a, b, full ? &f : nullptr, Quat{odom[0], odom[1], odom[2], odom[3]}, Vec3{odom[4], odom[5], odom[6]},
                 full ? &p->registered : nullptr, keep_log != 0);
  p->ms_map = now_ms() - t0;
  if (pose_out) {
    for (int k = 0; k < 7; ++k) pose_out[k] = p->map.parameters[k];
    pose_out[7] = p->map.q_wmap_wodom.x; pose_out[8] = p->map.q_wmap_wodom.y; pose_out[9] = p->map.q_wmap_wodom.z; pose_out[10] = p->map.q_wmap_wodom.w;
    pose_out[11] = p->map.t_wmap_wodom.x; pose_out[12] = p->map.t_wmap_wodom.y; pose_out[13] = p->map.t_wmap_wodom.z;
  }
  return p->map.map_too_small ? 3 : 0;
}
// outer-independent: what 0 cornerStack 1 surfStack 2 cornerFromMap 3 surfFromMap 4 registered (Pt);
//   5 info int[8] = centerCube(3), cen(3), total corner, total surf
// per outer: 10 corner_knn 11 surf_knn 12 corner_valid 13 surf_valid (int) ; 14 lm trace (double) ; 15 counts int[3] ; 16 final cost
long lvo_oracle_mapping_log(void* h, int outer, int what, void* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  LaserMapping& m = p->map;
  switch (what) {
    case 0: return copy_out(m.laserCloudCornerStack, out, cap);
    case 1: return copy_out(m.laserCloudSurfStack, out, cap);
    case 2: return copy_out(m.laserCloudCornerFromMap, out, cap);
    case 3: return copy_out(m.laserCloudSurfFromMap, out, cap);
    case 4: return copy_out(p->registered, out, cap);
    case 5: {
      std::vector<int> v{m.centerCube[0], m.centerCube[1], m.centerCube[2], m.laserCloudCenWidth, m.laserCloudCenHeight, m.laserCloudCenDepth,
                         (int)m.total_points(m.laserCloudCornerArray), (int)m.total_points(m.laserCloudSurfArray)};
      return copy_out(v, out, cap);
    }
  }
  if (outer < 0 || outer >= (int)m.log.size()) return -1;
  const MapOuterLog& lg = m.log[outer];
  switch (what) {
    case 10: return copy_out(lg.corner_knn, out, cap);
    case 11: return copy_out(lg.surf_knn, out, cap);
    case 12: return copy_out(lg.corner_valid, out, cap);
    case 13: return copy_out(lg.surf_valid, out, cap);
    case 14: {
      std::vector<double> v;
      for (const LmTraceRow& r : lg.lm) { for (int k = 0; k < 7; ++k) v.push_back(r.x[k]); v.push_back(r.cost); v.push_back(r.radius); v.push_back(r.flags); }
      return copy_out(v, out, cap);
    }
    case 15: { std::vector<int> v{lg.n_corner, lg.n_surf, lg.lm_iters}; return copy_out(v, out, cap); }
    case 16: { std::vector<double> v{lg.final_cost}; return copy_out(v, out, cap); }
  }
  return -1;
}
// map export/import: which 0 corner 1 surf. export returns n; cube[i] = array index of point i (cube-major order)
long lvo_oracle_map_export(void* h, int which, Pt* out, int* cube, long cap) {
  Pipeline* p = (Pipeline*)h;
  const std::vector<Cloud>& arr = which == 0 ? p->map.laserCloudCornerArray : p->map.laserCloudSurfArray;
  long n = 0;
  for (const Cloud& c : arr) n += (long)c.size();
  if (out && cap >= n) {
    long k = 0;
    for (int ci = 0; ci < (int)arr.size(); ++ci)
      for (const Pt& pt : arr[ci]) { out[k] = pt; if (cube) cube[k] = ci; ++k; }
  }
  return n;
}
// laserMapping.cpp:806-836: which 0 = surround cloud (valid cubes of the last frame, corner then surf per cube), 1 = whole map
long lvo_oracle_map_cloud(void* h, int which, Pt* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  Cloud c;
  if (which == 0) p->map.surround_cloud(c); else p->map.whole_map_cloud(c);
  return copy_out(c, out, cap);
}
void lvo_oracle_map_import(void* h, int which, const Pt* pts, const int* cube, long n) {
  Pipeline* p = (Pipeline*)h;
  std::vector<Cloud>& arr = which == 0 ? p->map.laserCloudCornerArray : p->map.laserCloudSurfArray;
  for (Cloud& c : arr) c.clear();
  for (long i = 0; i < n; ++i) arr[cube[i]].push_back(pts[i]);
}
void lvo_oracle_set_map_correction(void* h, const double* qt) {
  Pipeline* p = (Pipeline*)h;
  p->map.q_wmap_wodom = Quat{qt[0], qt[1], qt[2], qt[3]};
  p->map.t_wmap_wodom = Vec3{qt[4], qt[5], qt[6]};
}

void lvo_oracle_set_skip_frame(void* h, int skip) { ((Pipeline*)h)->skip_frame = skip < 1 ? 1 : skip; }
int lvo_oracle_last_frame_mapped(void* h) { return ((Pipeline*)h)->last_mapped; }

// Full per-frame chain as the three ROS nodes would run it: extract -> odometry -> (every mapping_skip_frame-th frame,
// laserOdometry.cpp:643-663) mapping.  poses_out[21] = [q_wodom(4), t_wodom(3)], [/aft_mapped_to_init of this frame if it was mapped,
// else the high-frequency pose], [the high-frequency pose /aft_mapped_to_init_high_frec that laserOdometryHandler publishes when the
// odometry message arrives, i.e. with the correction of the PREVIOUS mapped frame (laserMapping.cpp:197-229)].
// Returns the odometry status.
int lvo_oracle_step(void* h, const Pt* in, long n, double* poses_out, int keep_log) {
  Pipeline* p = (Pipeline*)h;
  double t0 = now_ms();
  scan_registration(in, (size_t)n, p->n_scans, p->min_range, p->reg);
  double t1 = now_ms();
  int r = p->odo.process(p->reg.sharp, p->reg.lessSharp, p->reg.flat, p->reg.lessFlat, keep_log != 0, &p->reg.full);
  double t2 = now_ms();
  // laserMapping.cpp:214-215
  const Quat hq = qmul(p->map.q_wmap_wodom, p->odo.q_w_curr);
  const Vec3 hr = rotate(p->map.q_wmap_wodom, p->odo.t_w_curr);
  const double hf[7] = {hq.x, hq.y, hq.z, hq.w, hr.x + p->map.t_wmap_wodom.x, hr.y + p->map.t_wmap_wodom.y, hr.z + p->map.t_wmap_wodom.z};
  const bool do_map = (p->frame_count % p->skip_frame) == 0;   // :643
  if (do_map) {
    p->frame_count = 0;                                        // :644
    // laserOdometry publishes laserCloudCornerLast/SurfLast (= this frame's less-sharp / less-flat) and the full cloud (:646-662)
    p->map.process(p->odo.laserCloudCornerLast, p->odo.laserCloudSurfLast, p->odo.distortion == 2 ? &p->odo.fullOut : &p->reg.full, p->odo.q_w_curr,
                   p->odo.t_w_curr, &p->registered, keep_log != 0);
  }
  p->frame_count++;                                            // :669
  p->last_mapped = do_map ? 1 : 0;
  double t3 = now_ms();
  p->ms_reg = t1 - t0; p->ms_odo = t2 - t1; p->ms_map = t3 - t2;
  if (poses_out) {
    poses_out[0] = p->odo.q_w_curr.x; poses_out[1] = p->odo.q_w_curr.y; poses_out[2] = p->odo.q_w_curr.z; poses_out[3] = p->odo.q_w_curr.w;
    poses_out[4] = p->odo.t_w_curr.x; poses_out[5] = p->odo.t_w_curr.y; poses_out[6] = p->odo.t_w_curr.z;
    for (int k = 0; k < 7; ++k) poses_out[7 + k] = do_map ? p->map.parameters[k] : hf[k];
    for (int k = 0; k < 7; ++k) poses_out[14 + k] = hf[k];
  }
  if (r == 0 && p->odo.few_corr) r = 2;
  return r;
}
void lvo_oracle_timings(void* h, double* ms3) {
  Pipeline* p = (Pipeline*)h;
  ms3[0] = p->ms_reg; ms3[1] = p->ms_odo; ms3[2] = p->ms_map;
}

}  // extern "C"
