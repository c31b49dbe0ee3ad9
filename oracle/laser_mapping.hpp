// oracle/laser_mapping.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates process(), reference src/laserMapping.cpp:307-848, with the pose helpers :142-163.  The cube
// pointer rotations of :323-507 are restated as rotations of a std::vector<Cloud> (same observable result).
#pragma once
#include "common.hpp"
#include "knn.hpp"
#include "voxel_grid.hpp"
#include "ceres_lm.hpp"

namespace lvo_oracle {

struct MapOuterLog {
  std::vector<int> corner_knn, surf_knn;      // [n_stack][5], -1 row when d5^2 >= 1.0
  std::vector<int> corner_valid, surf_valid;  // [n_stack]
  std::vector<LmTraceRow> lm;
  int n_corner = 0, n_surf = 0, lm_iters = 0;
  double final_cost = 0;
};

class LaserMapping {
 public:
  static const int laserCloudWidth = 21, laserCloudHeight = 21, laserCloudDepth = 11;  // :77-79
  static const int laserCloudNum = laserCloudWidth * laserCloudHeight * laserCloudDepth;  // 4851
  int laserCloudCenWidth = 10, laserCloudCenHeight = 10, laserCloudCenDepth = 5;  // :74-76
  float lineRes = 0.4f, planeRes = 0.8f;  // :902-906 (setLeafSize takes floats)
  int outer_iters = 10;                   // :562
  LmOptions lm;
  bool use_kdtree = true;

  std::vector<Cloud> laserCloudCornerArray, laserCloudSurfArray;
  double parameters[7] = {0, 0, 0, 1, 0, 0, 0};  // :110 q_w_curr, t_w_curr
  Quat q_wmap_wodom{0, 0, 0, 1};
  Vec3 t_wmap_wodom{0, 0, 0};
  Quat q_wodom_curr{0, 0, 0, 1};
  Vec3 t_wodom_curr{0, 0, 0};

  // logs of the last frame
  Cloud laserCloudCornerStack, laserCloudSurfStack, laserCloudCornerFromMap, laserCloudSurfFromMap;
  std::vector<MapOuterLog> log;
  int centerCube[3] = {0, 0, 0};
  bool map_too_small = false;
  int laserCloudValidInd[125];     // :85; in this fork laserCloudSurroundInd (:86) is filled with the same cubes (:524-525)
  int laserCloudValidNum = 0;

  LaserMapping() : laserCloudCornerArray(laserCloudNum), laserCloudSurfArray(laserCloudNum) {}

  Quat q_w_curr() const { return Quat{parameters[0], parameters[1], parameters[2], parameters[3]}; }
  Vec3 t_w_curr() const { return Vec3{parameters[4], parameters[5], parameters[6]}; }

  void pointAssociateToMap(const Pt& pi, Pt& po) const {  // :154-163
    Vec3 w = rotate(q_w_curr(), Vec3{pi.x, pi.y, pi.z});
    po.x = (float)(w.x + parameters[4]); po.y = (float)(w.y + parameters[5]); po.z = (float)(w.z + parameters[6]);
    po.i = pi.i;
  }

  // One frame.  full may be null; registered (optional) gets the transformed full-res cloud (:838-842).
  void process(const Cloud& laserCloudCornerLast, const Cloud& laserCloudSurfLast, const Cloud* full, const Quat& q_odom,
               const Vec3& t_odom, Cloud* registered, bool keep_log = true) {
    q_wodom_curr = q_odom; t_wodom_curr = t_odom;
    log.clear();
    // transformAssociateToMap :142-146
    {
      Quat q = qmul(q_wmap_wodom, q_wodom_curr);
      Vec3 r = rotate(q_wmap_wodom, t_wodom_curr);
      parameters[0] = q.x; parameters[1] = q.y; parameters[2] = q.z; parameters[3] = q.w;
      parameters[4] = r.x + t_wmap_wodom.x; parameters[5] = r.y + t_wmap_wodom.y; parameters[6] = r.z + t_wmap_wodom.z;
    }
    int centerCubeI = int((parameters[4] + 25.0) / 50.0) + laserCloudCenWidth;   // :312-321
    int centerCubeJ = int((parameters[5] + 25.0) / 50.0) + laserCloudCenHeight;
    int centerCubeK = int((parameters[6] + 25.0) / 50.0) + laserCloudCenDepth;
    if (parameters[4] + 25.0 < 0) centerCubeI--;
    if (parameters[5] + 25.0 < 0) centerCubeJ--;
    if (parameters[6] + 25.0 < 0) centerCubeK--;

    auto at = [&](int i, int j, int k) { return i + laserCloudWidth * j + laserCloudWidth * laserCloudHeight * k; };
    auto shift = [&](std::vector<Cloud>& A, int axis, int dir) {
      // dir=+1: contents move towards higher index along `axis`, the last slab is recycled (cleared) into index 0
      const int dims[3] = {laserCloudWidth, laserCloudHeight, laserCloudDepth};
      int n = dims[axis], a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
      for (int u = 0; u < dims[a1]; ++u)
        for (int v = 0; v < dims[a2]; ++v) {
          int idx[3];
          idx[a1] = u; idx[a2] = v;
          if (dir > 0) {
            for (int w = n - 1; w >= 1; --w) { idx[axis] = w; int dst = at(idx[0], idx[1], idx[2]); idx[axis] = w - 1; A[dst].swap(A[at(idx[0], idx[1], idx[2])]); }
            idx[axis] = 0; A[at(idx[0], idx[1], idx[2])].clear();
          } else {
            for (int w = 0; w < n - 1; ++w) { idx[axis] = w; int dst = at(idx[0], idx[1], idx[2]); idx[axis] = w + 1; A[dst].swap(A[at(idx[0], idx[1], idx[2])]); }
            idx[axis] = n - 1; A[at(idx[0], idx[1], idx[2])].clear();
          }
        }
    };
    while (centerCubeI < 3) { shift(laserCloudCornerArray, 0, +1); shift(laserCloudSurfArray, 0, +1); centerCubeI++; laserCloudCenWidth++; }   // :323-352
    while (centerCubeI >= laserCloudWidth - 3) { shift(laserCloudCornerArray, 0, -1); shift(laserCloudSurfArray, 0, -1); centerCubeI--; laserCloudCenWidth--; }  // :354-383
    while (centerCubeJ < 3) { shift(laserCloudCornerArray, 1, +1); shift(laserCloudSurfArray, 1, +1); centerCubeJ++; laserCloudCenHeight++; }  // :385-414
    while (centerCubeJ >= laserCloudHeight - 3) { shift(laserCloudCornerArray, 1, -1); shift(laserCloudSurfArray, 1, -1); centerCubeJ--; laserCloudCenHeight--; }  // :416-445
    while (centerCubeK < 3) { shift(laserCloudCornerArray, 2, +1); shift(laserCloudSurfArray, 2, +1); centerCubeK++; laserCloudCenDepth++; }   // :447-476
    while (centerCubeK >= laserCloudDepth - 3) { shift(laserCloudCornerArray, 2, -1); shift(laserCloudSurfArray, 2, -1); centerCubeK--; laserCloudCenDepth--; }  // :478-507
    centerCube[0] = centerCubeI; centerCube[1] = centerCubeJ; centerCube[2] = centerCubeK;

    laserCloudValidNum = 0;
    for (int i = centerCubeI - 2; i <= centerCubeI + 2; i++)      // :512-529
      for (int j = centerCubeJ - 2; j <= centerCubeJ + 2; j++)
        for (int k = centerCubeK - 1; k <= centerCubeK + 1; k++)
          if (i >= 0 && i < laserCloudWidth && j >= 0 && j < laserCloudHeight && k >= 0 && k < laserCloudDepth)
            laserCloudValidInd[laserCloudValidNum++] = at(i, j, k);

    laserCloudCornerFromMap.clear(); laserCloudSurfFromMap.clear();
    for (int i = 0; i < laserCloudValidNum; i++) {  // :533-537
      const Cloud& c = laserCloudCornerArray[laserCloudValidInd[i]];
      const Cloud& s = laserCloudSurfArray[laserCloudValidInd[i]];
      laserCloudCornerFromMap.insert(laserCloudCornerFromMap.end(), c.begin(), c.end());
      laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), s.begin(), s.end());
    }
    int laserCloudCornerFromMapNum = (int)laserCloudCornerFromMap.size();
    int laserCloudSurfFromMapNum = (int)laserCloudSurfFromMap.size();

    voxel_grid(laserCloudCornerLast, lineRes, laserCloudCornerStack);   // :542-545
    voxel_grid(laserCloudSurfLast, planeRes, laserCloudSurfStack);      // :547-550
    int laserCloudCornerStackNum = (int)laserCloudCornerStack.size();
    int laserCloudSurfStackNum = (int)laserCloudSurfStack.size();

    map_too_small = !(laserCloudCornerFromMapNum > 10 && laserCloudSurfFromMapNum > 50);  // :554
    if (!map_too_small) {
      if (use_kdtree) { kdCorner_.build(laserCloudCornerFromMap); kdSurf_.build(laserCloudSurfFromMap); }  // :558-559
      for (int iterCount = 0; iterCount < outer_iters; iterCount++) {  // :562
        MapOuterLog lg;
        std::vector<Factor> factors;
        Neighbor nb[5];
        for (int i = 0; i < laserCloudCornerStackNum; i++) {  // :577-640
          Pt pointOri = laserCloudCornerStack[i], pointSel;
          pointAssociateToMap(pointOri, pointSel);
          int cnt = nn5(kdCorner_, laserCloudCornerFromMap, pointSel, nb);
          bool added = false;
          bool gate = cnt == 5 && nb[4].d < 1.0;  // :584
          if (gate) {
            Vec3 nearCorners[5];
            Vec3 center{0, 0, 0};
            for (int j = 0; j < 5; j++) {
              const Pt& p = laserCloudCornerFromMap[nb[j].i];
              nearCorners[j] = Vec3{p.x, p.y, p.z};
              center = Vec3{center.x + p.x, center.y + p.y, center.z + p.z};
            }
            center = Vec3{center.x / 5.0, center.y / 5.0, center.z / 5.0};
            double covMat[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < 5; j++) {
              double z[3] = {nearCorners[j].x - center.x, nearCorners[j].y - center.y, nearCorners[j].z - center.z};
              for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) covMat[r * 3 + c] = covMat[r * 3 + c] + z[r] * z[c];
            }
            double w[3], V[9];
            sym_eigen3(covMat, w, V);  // :605
            Vec3 unit_direction{V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2]};
            if (w[2] > 3 * w[1]) {  // :611
              Factor f;
              f.type = F_EDGE;
              f.c = Vec3{pointOri.x, pointOri.y, pointOri.z};
              f.a = Vec3{0.1 * unit_direction.x + center.x, 0.1 * unit_direction.y + center.y, 0.1 * unit_direction.z + center.z};
              f.b = Vec3{-0.1 * unit_direction.x + center.x, -0.1 * unit_direction.y + center.y, -0.1 * unit_direction.z + center.z};
              f.m = Vec3{0, 0, 0}; f.d = 0;
              factors.push_back(f);
              lg.n_corner++;
              added = true;
            }
          }
          if (keep_log) {
            for (int j = 0; j < 5; ++j) lg.corner_knn.push_back(gate ? nb[j].i : -1);
            lg.corner_valid.push_back(added ? 1 : 0);
          }
        }
        for (int i = 0; i < laserCloudSurfStackNum; i++) {  // :643-705
          Pt pointOri = laserCloudSurfStack[i], pointSel;
          pointAssociateToMap(pointOri, pointSel);
          int cnt = nn5(kdSurf_, laserCloudSurfFromMap, pointSel, nb);
          bool added = false;
          bool gate = cnt == 5 && nb[4].d < 1.0;  // :652
          if (gate) {
            double matA0[15];
            for (int j = 0; j < 5; j++) {
              const Pt& p = laserCloudSurfFromMap[nb[j].i];
              matA0[j * 3 + 0] = p.x; matA0[j * 3 + 1] = p.y; matA0[j * 3 + 2] = p.z;
            }
            double norm[3];
            plane_fit5(matA0, norm);  // :663
            double nn = sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
            double negative_OA_dot_norm = 1 / nn;   // :664
            norm[0] /= nn; norm[1] /= nn; norm[2] /= nn;  // :665
            bool planeValid = true;
            for (int j = 0; j < 5; j++) {  // :669-679
              const Pt& p = laserCloudSurfFromMap[nb[j].i];
              if (fabs(norm[0] * p.x + norm[1] * p.y + norm[2] * p.z + negative_OA_dot_norm) > 0.2) { planeValid = false; break; }
            }
            if (planeValid) {
              Factor f;
              f.type = F_PLANE_NORM;
              f.c = Vec3{pointOri.x, pointOri.y, pointOri.z};
              f.a = Vec3{norm[0], norm[1], norm[2]};
              f.b = Vec3{0, 0, 0}; f.m = Vec3{0, 0, 0};
              f.d = negative_OA_dot_norm;
              factors.push_back(f);
              lg.n_surf++;
              added = true;
            }
          }
          if (keep_log) {
            for (int j = 0; j < 5; ++j) lg.surf_knn.push_back(gate ? nb[j].i : -1);
            lg.surf_valid.push_back(added ? 1 : 0);
          }
        }
        LmSummary s = solve(factors, parameters, lm, keep_log ? &lg.lm : nullptr);  // :712-720
        lg.lm_iters = s.iterations;
        lg.final_cost = s.final_cost;
        log.push_back(std::move(lg));
      }
    }
    // transformUpdate :148-152
    q_wmap_wodom = qmul(q_w_curr(), qinv(q_wodom_curr));
    {
      Vec3 r = rotate(q_wmap_wodom, t_wodom_curr);
      t_wmap_wodom = Vec3{parameters[4] - r.x, parameters[5] - r.y, parameters[6] - r.z};
    }
    auto insert = [&](const Cloud& stack, std::vector<Cloud>& arr) {  // :737-783
      for (size_t i = 0; i < stack.size(); i++) {
        Pt pointSel;
        pointAssociateToMap(stack[i], pointSel);
        int cubeI = int((pointSel.x + 25.0) / 50.0) + laserCloudCenWidth;
        int cubeJ = int((pointSel.y + 25.0) / 50.0) + laserCloudCenHeight;
        int cubeK = int((pointSel.z + 25.0) / 50.0) + laserCloudCenDepth;
        if (pointSel.x + 25.0 < 0) cubeI--;
        if (pointSel.y + 25.0 < 0) cubeJ--;
        if (pointSel.z + 25.0 < 0) cubeK--;
        if (cubeI >= 0 && cubeI < laserCloudWidth && cubeJ >= 0 && cubeJ < laserCloudHeight && cubeK >= 0 && cubeK < laserCloudDepth)
          arr[at(cubeI, cubeJ, cubeK)].push_back(pointSel);
      }
    };
    insert(laserCloudCornerStack, laserCloudCornerArray);
    insert(laserCloudSurfStack, laserCloudSurfArray);
    for (int i = 0; i < laserCloudValidNum; i++) {  // :788-801
      int ind = laserCloudValidInd[i];
      Cloud tmpCorner, tmpSurf;
      voxel_grid(laserCloudCornerArray[ind], lineRes, tmpCorner);
      laserCloudCornerArray[ind].swap(tmpCorner);
      voxel_grid(laserCloudSurfArray[ind], planeRes, tmpSurf);
      laserCloudSurfArray[ind].swap(tmpSurf);
    }
    if (full && registered) {  // :838-842
      registered->resize(full->size());
      for (size_t i = 0; i < full->size(); ++i) pointAssociateToMap((*full)[i], (*registered)[i]);
    }
  }

  // "publish surround map for every 5 frame" :806-815: the valid cubes of the last frame, corner then surf per cube
  void surround_cloud(Cloud& out) const {
    out.clear();
    for (int i = 0; i < laserCloudValidNum; i++) {
      const int ind = laserCloudValidInd[i];
      out.insert(out.end(), laserCloudCornerArray[ind].begin(), laserCloudCornerArray[ind].end());
      out.insert(out.end(), laserCloudSurfArray[ind].begin(), laserCloudSurfArray[ind].end());
    }
  }
  // whole-map dump every 20 frames :823-836
  void whole_map_cloud(Cloud& out) const {
    out.clear();
    for (int i = 0; i < laserCloudNum; i++) {
      out.insert(out.end(), laserCloudCornerArray[i].begin(), laserCloudCornerArray[i].end());
      out.insert(out.end(), laserCloudSurfArray[i].begin(), laserCloudSurfArray[i].end());
    }
  }

  size_t total_points(const std::vector<Cloud>& arr) const { size_t s = 0; for (const Cloud& c : arr) s += c.size(); return s; }

 private:
  KdTree kdCorner_, kdSurf_;
  int nn5(const KdTree& kd, const Cloud& c, const Pt& q, Neighbor* nb) const {
    if (use_kdtree) return kd.knn(q.x, q.y, q.z, 5, nb);
    return knn_brute(c, q.x, q.y, q.z, 5, nb);
  }
};

}  // namespace lvo_oracle
