// oracle/laser_odometry.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates the odometry frame body, reference src/laserOdometry.cpp:353-641, with TransformToStart :154-172 and
// TransformToEnd :176-191.  `distortion` selects the compile-time switch of :67: 0 = DISTORTION 0 (the shipped value),
// 1 = DISTORTION 1 as the file is written (per-point interpolation ratio in TransformToStart and in the factors; the
// TransformToEnd block :610-625 stays disabled by its `if (0)`), 2 = DISTORTION 1 with that block enabled.
// State that the reference keeps in globals (:95-137) lives in the class.
#pragma once
#include "common.hpp"
#include "knn.hpp"
#include "ceres_lm.hpp"

namespace lvo_oracle {

struct OdoOuterLog {
  std::vector<int> corner_corr;  // [n_sharp][2]  closestPointInd, minPointInd2 (-1,-1 if no factor)
  std::vector<int> plane_corr;   // [n_flat][3]
  std::vector<LmTraceRow> lm;
  int n_corner = 0, n_plane = 0, lm_iters = 0;
  double final_cost = 0;
};

class LaserOdometry {
 public:
  int outer_iters = 10;  // :364
  LmOptions lm;          // max_num_iterations 4 (:573), HuberLoss(0.1) (:369)
  bool use_kdtree = true;  // false: brute-force 1-NN (validation of the kd-tree itself)
  int distortion = 0;      // see the header comment
  Cloud lessSharpOut, lessFlatOut, fullOut;  // distortion == 2: the clouds after TransformToEnd (what :646-656 would publish)

  // interpolation ratio :157-161 / :455-459: float subtraction, then a double division by SCAN_PERIOD
  static double ratio(const Pt& p, int distortion) { return distortion ? (double)(p.i - (float)int(p.i)) / 0.1 : 1.0; }
  static void transform_to_start(const Quat& q_last_curr, const Vec3& t_last_curr, int distortion, const Pt& pi, Pt& po) {  // :154-172
    const double s = ratio(pi, distortion);
    const Quat q_point_last = distortion ? slerp_identity(q_last_curr, s) : q_last_curr;  // Identity.slerp(1, q) == q exactly
    const Vec3 t_point_last{s * t_last_curr.x, s * t_last_curr.y, s * t_last_curr.z};
    Vec3 un = rotate(q_point_last, Vec3{pi.x, pi.y, pi.z});
    po.x = (float)(un.x + t_point_last.x); po.y = (float)(un.y + t_point_last.y); po.z = (float)(un.z + t_point_last.z);
    po.i = pi.i;
  }
  static void transform_to_end(const Quat& q_last_curr, const Vec3& t_last_curr, int distortion, const Pt& pi, Pt& po) {  // :176-191
    Pt un_point_tmp;
    transform_to_start(q_last_curr, t_last_curr, distortion, pi, un_point_tmp);
    const Vec3 d{(double)un_point_tmp.x - t_last_curr.x, (double)un_point_tmp.y - t_last_curr.y, (double)un_point_tmp.z - t_last_curr.z};
    const Vec3 e = rotate(qinv(q_last_curr), d);
    po.x = (float)e.x; po.y = (float)e.y; po.z = (float)e.z;
    po.i = (float)int(pi.i);
  }

  bool systemInited = false;
  Quat q_w_curr{0, 0, 0, 1};
  Vec3 t_w_curr{0, 0, 0};
  double para_q[4] = {0, 0, 0, 1};  // :131
  double para_t[3] = {0, 0, 0};     // :133
  Cloud laserCloudCornerLast, laserCloudSurfLast;
  std::vector<OdoOuterLog> log;     // per outer iteration of the last frame
  bool few_corr = false;

  // Returns 1 on the first frame (initialisation only), else 0.
  int process(const Cloud& cornerPointsSharp, const Cloud& cornerPointsLessSharp, const Cloud& surfPointsFlat,
              const Cloud& surfPointsLessFlat, bool keep_log = true, const Cloud* full = nullptr) {
    int ret = 0;
    log.clear();
    few_corr = false;
    if (!systemInited) {
      systemInited = true;
      ret = 1;
    } else {
      const double DISTANCE_SQ_THRESHOLD = 25, NEARBY_SCAN = 2.5;  // :73-77
      int cornerPointsSharpNum = (int)cornerPointsSharp.size();
      int surfPointsFlatNum = (int)surfPointsFlat.size();
      const Cloud& CL = laserCloudCornerLast;
      const Cloud& SL = laserCloudSurfLast;
      for (int opti_counter = 0; opti_counter < outer_iters; ++opti_counter) {
        OdoOuterLog lg;
        std::vector<Factor> factors;
        Quat q_last_curr{para_q[0], para_q[1], para_q[2], para_q[3]};
        Vec3 t_last_curr{para_t[0], para_t[1], para_t[2]};
        auto TransformToStart = [&](const Pt& pi, Pt& po) { transform_to_start(q_last_curr, t_last_curr, distortion, pi, po); };
        Neighbor nb;
        for (int i = 0; i < cornerPointsSharpNum; ++i) {  // :384-465
          Pt pointSel;
          TransformToStart(cornerPointsSharp[i], pointSel);
          int found = nn1(kdCorner_, CL, pointSel, nb);
          int closestPointInd = -1, minPointInd2 = -1;
          if (found && nb.d < DISTANCE_SQ_THRESHOLD) {
            closestPointInd = nb.i;
            int closestPointScanID = int(CL[closestPointInd].i);
            double minPointSqDis2 = DISTANCE_SQ_THRESHOLD;
            for (int j = closestPointInd + 1; j < (int)CL.size(); ++j) {
              if (int(CL[j].i) <= closestPointScanID) continue;
              if (int(CL[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
              double pointSqDis = (CL[j].x - pointSel.x) * (CL[j].x - pointSel.x) + (CL[j].y - pointSel.y) * (CL[j].y - pointSel.y) +
                                  (CL[j].z - pointSel.z) * (CL[j].z - pointSel.z);
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
            for (int j = closestPointInd - 1; j >= 0; --j) {
              if (int(CL[j].i) >= closestPointScanID) continue;
              if (int(CL[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
              double pointSqDis = (CL[j].x - pointSel.x) * (CL[j].x - pointSel.x) + (CL[j].y - pointSel.y) * (CL[j].y - pointSel.y) +
                                  (CL[j].z - pointSel.z) * (CL[j].z - pointSel.z);
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
          if (minPointInd2 >= 0) {
            Factor f;
            f.type = F_EDGE;
            f.c = Vec3{cornerPointsSharp[i].x, cornerPointsSharp[i].y, cornerPointsSharp[i].z};
            f.a = Vec3{CL[closestPointInd].x, CL[closestPointInd].y, CL[closestPointInd].z};
            f.b = Vec3{CL[minPointInd2].x, CL[minPointInd2].y, CL[minPointInd2].z};
            f.m = Vec3{0, 0, 0}; f.d = 0;
            f.s = ratio(cornerPointsSharp[i], distortion);  // :455-459
            factors.push_back(f);
            lg.n_corner++;
            if (keep_log) { lg.corner_corr.push_back(closestPointInd); lg.corner_corr.push_back(minPointInd2); }
          } else if (keep_log) { lg.corner_corr.push_back(-1); lg.corner_corr.push_back(-1); }
        }
        for (int i = 0; i < surfPointsFlatNum; ++i) {  // :468-561
          Pt pointSel;
          TransformToStart(surfPointsFlat[i], pointSel);
          int found = nn1(kdSurf_, SL, pointSel, nb);
          int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
          if (found && nb.d < DISTANCE_SQ_THRESHOLD) {
            closestPointInd = nb.i;
            int closestPointScanID = int(SL[closestPointInd].i);
            double minPointSqDis2 = DISTANCE_SQ_THRESHOLD, minPointSqDis3 = DISTANCE_SQ_THRESHOLD;
            for (int j = closestPointInd + 1; j < (int)SL.size(); ++j) {
              if (int(SL[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
              double pointSqDis = (SL[j].x - pointSel.x) * (SL[j].x - pointSel.x) + (SL[j].y - pointSel.y) * (SL[j].y - pointSel.y) +
                                  (SL[j].z - pointSel.z) * (SL[j].z - pointSel.z);
              if (int(SL[j].i) <= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
              else if (int(SL[j].i) > closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
            for (int j = closestPointInd - 1; j >= 0; --j) {
              if (int(SL[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
              double pointSqDis = (SL[j].x - pointSel.x) * (SL[j].x - pointSel.x) + (SL[j].y - pointSel.y) * (SL[j].y - pointSel.y) +
                                  (SL[j].z - pointSel.z) * (SL[j].z - pointSel.z);
              if (int(SL[j].i) >= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
              else if (int(SL[j].i) < closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
          }
          if (minPointInd2 >= 0 && minPointInd3 >= 0) {
            Factor f;
            f.type = F_PLANE;
            f.c = Vec3{surfPointsFlat[i].x, surfPointsFlat[i].y, surfPointsFlat[i].z};
            f.a = Vec3{SL[closestPointInd].x, SL[closestPointInd].y, SL[closestPointInd].z};
            f.b = Vec3{SL[minPointInd2].x, SL[minPointInd2].y, SL[minPointInd2].z};
            f.m = Vec3{SL[minPointInd3].x, SL[minPointInd3].y, SL[minPointInd3].z};
            f.d = 0;
            f.s = ratio(surfPointsFlat[i], distortion);  // :549-553
            factors.push_back(f);
            lg.n_plane++;
            if (keep_log) { lg.plane_corr.push_back(closestPointInd); lg.plane_corr.push_back(minPointInd2); lg.plane_corr.push_back(minPointInd3); }
          } else if (keep_log) { lg.plane_corr.push_back(-1); lg.plane_corr.push_back(-1); lg.plane_corr.push_back(-1); }
        }
        if (lg.n_corner + lg.n_plane < 10) few_corr = true;  // :566-568
        double x[7] = {para_q[0], para_q[1], para_q[2], para_q[3], para_t[0], para_t[1], para_t[2]};
        LmSummary s = solve(factors, x, lm, keep_log ? &lg.lm : nullptr);  // :571-576
        for (int k = 0; k < 4; ++k) para_q[k] = x[k];
        for (int k = 0; k < 3; ++k) para_t[k] = x[4 + k];
        lg.lm_iters = s.iterations;
        lg.final_cost = s.final_cost;
        log.push_back(std::move(lg));
      }
      Quat q_last_curr{para_q[0], para_q[1], para_q[2], para_q[3]};
      Vec3 t_last_curr{para_t[0], para_t[1], para_t[2]};
      Vec3 dt = rotate(q_w_curr, t_last_curr);  // :581
      t_w_curr = Vec3{t_w_curr.x + dt.x, t_w_curr.y + dt.y, t_w_curr.z + dt.z};
      q_w_curr = qmul(q_w_curr, q_last_curr);   // :582
    }
    laserCloudCornerLast = cornerPointsLessSharp;  // :627-633 (pointer swap)
    laserCloudSurfLast = surfPointsLessFlat;
    if (distortion == 2) {  // the block :610-625 with its `if (0)` lifted; runs on every frame, the first included (para = identity)
      const Quat q_last_curr{para_q[0], para_q[1], para_q[2], para_q[3]};
      const Vec3 t_last_curr{para_t[0], para_t[1], para_t[2]};
      for (Pt& p : laserCloudCornerLast) transform_to_end(q_last_curr, t_last_curr, 1, p, p);
      for (Pt& p : laserCloudSurfLast) transform_to_end(q_last_curr, t_last_curr, 1, p, p);
      lessSharpOut = laserCloudCornerLast; lessFlatOut = laserCloudSurfLast;
      if (full) { fullOut = *full; for (Pt& p : fullOut) transform_to_end(q_last_curr, t_last_curr, 1, p, p); }
    }
    if (use_kdtree) { kdCorner_.build(laserCloudCornerLast); kdSurf_.build(laserCloudSurfLast); }  // :640-641
    return ret;
  }

 private:
  KdTree kdCorner_, kdSurf_;
  int nn1(const KdTree& kd, const Cloud& c, const Pt& q, Neighbor& nb) const {
    if (use_kdtree) return kd.knn(q.x, q.y, q.z, 1, &nb);
    return knn_brute(c, q.x, q.y, q.z, 1, &nb);
  }
};

}  // namespace lvo_oracle
