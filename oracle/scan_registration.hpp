// oracle/scan_registration.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates laserCloudHandler, reference src/scanRegistration.cpp:127-411, and removeClosedPointCloud :85-112.
#pragma once
#include "common.hpp"
#include "voxel_grid.hpp"

namespace lvo_oracle {

struct ScanRegOut {
  int n_in = 0, n_kept = 0;
  Cloud full;                       // laserCloud, ring-ordered            :246-252
  std::vector<float> curvature;     // cloudCurvature                      :262
  std::vector<int> sortInd, picked, label;  // cloudSortInd / cloudNeighborPicked / cloudLabel
  std::vector<int> scanStart, scanEnd;      // scanStartInd / scanEndInd   :249,251
  Cloud sharp, lessSharp, flat, lessFlat;   // :271-274
};

// Ring id of one point (:166-200).  Returns -1 when the point is rejected (`continue` at :175/:184/:198).
// Overload resolution as in the deviation register: atan/sqrt resolve to the float overloads.
inline int ring_of(float x, float y, float z, int N_SCANS) {
  float angle = (float)((double)(lvo_atanf(z / sqrtf(x * x + y * y)) * 180) / M_PI);  // :166
  int scanID = 0;
  if (N_SCANS == 16) {
    scanID = int((angle + 15) / 2 + 0.5);                       // :171 float, float, then double
    if (scanID > (N_SCANS - 1) || scanID < 0) return -1;
  } else if (N_SCANS == 32) {
    scanID = int((angle + 92.0 / 3.0) * 3.0 / 4.0);             // :180 double throughout
    if (scanID > (N_SCANS - 1) || scanID < 0) return -1;
  } else if (N_SCANS == 64) {
    if (angle >= -8.83) scanID = int((2 - angle) * 3.0 + 0.5);  // :189-190 (2-angle) is float
    else scanID = N_SCANS / 2 + int((-8.83 - angle) * 2.0 + 0.5);  // :192
    if (angle > 2 || angle < -24.33 || scanID > 50 || scanID < 0) return -1;  // :195
  } else {
    return -2;  // "wrong scan number", ROS_BREAK :203
  }
  return scanID;
}

// Returns 0, or -1 for an unsupported N_SCANS.
inline int scan_registration(const Pt* in, size_t n, int N_SCANS, double MINIMUM_RANGE, ScanRegOut& o) {
  const double scanPeriod = 0.1;  // :60
  if (N_SCANS != 16 && N_SCANS != 32 && N_SCANS != 64) return -1;
  o = ScanRegOut();
  o.n_in = (int)n;
  o.scanStart.assign(N_SCANS, 0);
  o.scanEnd.assign(N_SCANS, 0);

  // :136 pcl::removeNaNFromPointCloud (drops non-finite xyz) ; :137 removeClosedPointCloud (:85-112)
  std::vector<Pt> laserCloudIn;
  laserCloudIn.reserve(n);
  float thres = (float)MINIMUM_RANGE;
  for (size_t i = 0; i < n; ++i) {
    const Pt& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    if (p.x * p.x + p.y * p.y + p.z * p.z < thres * thres) continue;  // :99
    laserCloudIn.push_back(p);
  }
  int cloudSize = (int)laserCloudIn.size();
  if (cloudSize == 0) {  // the reference would index points[0] of an empty cloud; we return empty outputs
    return 0;
  }
  float startOri = -lvo_atan2f(laserCloudIn[0].y, laserCloudIn[0].x);                                  // :141
  float endOri = (float)((double)(-lvo_atan2f(laserCloudIn[cloudSize - 1].y, laserCloudIn[cloudSize - 1].x)) + 2 * M_PI);  // :142-144
  if (endOri - startOri > 3 * M_PI) endOri = (float)((double)endOri - 2 * M_PI);       // :146-149
  else if (endOri - startOri < M_PI) endOri = (float)((double)endOri + 2 * M_PI);      // :150-153

  bool halfPassed = false;
  int count = cloudSize;
  std::vector<Cloud> laserCloudScans(N_SCANS);
  for (int i = 0; i < cloudSize; i++) {
    Pt point;
    point.x = laserCloudIn[i].x; point.y = laserCloudIn[i].y; point.z = laserCloudIn[i].z;
    int scanID = ring_of(point.x, point.y, point.z, N_SCANS);
    if (scanID < 0) { count--; continue; }

    float ori = -lvo_atan2f(point.y, point.x);  // :208
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) ori = (float)((double)ori + 2 * M_PI);             // :211-214
      else if (ori > startOri + M_PI * 3 / 2) ori = (float)((double)ori - 2 * M_PI);    // :215-218
      if (ori - startOri > M_PI) halfPassed = true;                                      // :220-223
    } else {
      ori = (float)((double)ori + 2 * M_PI);                                             // :227
      if (ori < endOri - M_PI * 3 / 2) ori = (float)((double)ori + 2 * M_PI);            // :228-231
      else if (ori > endOri + M_PI / 2) ori = (float)((double)ori - 2 * M_PI);           // :232-235
    }
    float relTime = (ori - startOri) / (endOri - startOri);   // :238
    point.i = (float)(scanID + scanPeriod * relTime);         // :239
    laserCloudScans[scanID].push_back(point);
  }
  cloudSize = count;
  o.n_kept = cloudSize;

  Cloud& laserCloud = o.full;
  laserCloud.reserve(cloudSize);
  for (int i = 0; i < N_SCANS; i++) {   // :247-252
    o.scanStart[i] = (int)laserCloud.size() + 5;
    laserCloud.insert(laserCloud.end(), laserCloudScans[i].begin(), laserCloudScans[i].end());
    o.scanEnd[i] = (int)laserCloud.size() - 6;
  }

  std::vector<float>& cloudCurvature = o.curvature;
  std::vector<int>& cloudSortInd = o.sortInd;
  std::vector<int>& cloudNeighborPicked = o.picked;
  std::vector<int>& cloudLabel = o.label;
  cloudCurvature.assign(cloudSize, 0.f);
  cloudSortInd.assign(cloudSize, 0);
  cloudNeighborPicked.assign(cloudSize, 0);
  cloudLabel.assign(cloudSize, 0);
  const Cloud& P = laserCloud;
  for (int i = 5; i < cloudSize - 5; i++) {  // :256-266, evaluated strictly left to right in float
    float diffX = P[i - 5].x + P[i - 4].x + P[i - 3].x + P[i - 2].x + P[i - 1].x - 10 * P[i].x + P[i + 1].x + P[i + 2].x + P[i + 3].x + P[i + 4].x + P[i + 5].x;
    float diffY = P[i - 5].y + P[i - 4].y + P[i - 3].y + P[i - 2].y + P[i - 1].y - 10 * P[i].y + P[i + 1].y + P[i + 2].y + P[i + 3].y + P[i + 4].y + P[i + 5].y;
    float diffZ = P[i - 5].z + P[i - 4].z + P[i - 3].z + P[i - 2].z + P[i - 1].z - 10 * P[i].z + P[i + 1].z + P[i + 2].z + P[i + 3].z + P[i + 4].z + P[i + 5].z;
    cloudCurvature[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
    cloudSortInd[i] = i;
    cloudNeighborPicked[i] = 0;
    cloudLabel[i] = 0;
  }

  // gap test of :321-324 etc.
  auto gapBig = [&](int a, int b) {
    float diffX = P[a].x - P[b].x, diffY = P[a].y - P[b].y, diffZ = P[a].z - P[b].z;
    return diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05;
  };

  for (int i = 0; i < N_SCANS; i++) {  // :277
    if (o.scanEnd[i] - o.scanStart[i] < 6) continue;
    Cloud surfPointsLessFlatScan;
    for (int j = 0; j < 6; j++) {
      int sp = o.scanStart[i] + (o.scanEnd[i] - o.scanStart[i]) * j / 6;            // :284
      int ep = o.scanStart[i] + (o.scanEnd[i] - o.scanStart[i]) * (j + 1) / 6 - 1;  // :285
      // :288 std::sort by curvature; tie order fixed by contract to (curvature, index) ascending
      std::sort(cloudSortInd.begin() + sp, cloudSortInd.begin() + ep + 1, [&](int a, int b) {
        if (cloudCurvature[a] != cloudCurvature[b]) return cloudCurvature[a] < cloudCurvature[b];
        return a < b;
      });

      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {  // :292-344
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > 0.1) {
          largestPickedNum++;
          if (largestPickedNum <= 2) {
            cloudLabel[ind] = 2;
            o.sharp.push_back(P[ind]);
            o.lessSharp.push_back(P[ind]);
          } else if (largestPickedNum <= 20) {
            cloudLabel[ind] = 1;
            o.lessSharp.push_back(P[ind]);
          } else {
            break;
          }
          cloudNeighborPicked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gapBig(ind + l, ind + l - 1)) break;
            cloudNeighborPicked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gapBig(ind + l, ind + l + 1)) break;
            cloudNeighborPicked[ind + l] = 1;
          }
        }
      }

      int smallestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {  // :347-390
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < 0.1) {
          cloudLabel[ind] = -1;
          o.flat.push_back(P[ind]);
          smallestPickedNum++;
          if (smallestPickedNum >= 4) break;   // :359-362 — the 4th flat point is not neighbour-suppressed
          cloudNeighborPicked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gapBig(ind + l, ind + l - 1)) break;
            cloudNeighborPicked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gapBig(ind + l, ind + l + 1)) break;
            cloudNeighborPicked[ind + l] = 1;
          }
        }
      }

      for (int k = sp; k <= ep; k++)   // :392-398
        if (cloudLabel[k] <= 0) surfPointsLessFlatScan.push_back(P[k]);
    }
    Cloud surfPointsLessFlatScanDS;    // :401-407
    voxel_grid(surfPointsLessFlatScan, 0.2f, surfPointsLessFlatScanDS);
    o.lessFlat.insert(o.lessFlat.end(), surfPointsLessFlatScanDS.begin(), surfPointsLessFlatScanDS.end());
  }
  return 0;
}

}  // namespace lvo_oracle
