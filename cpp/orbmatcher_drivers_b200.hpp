// orbmatcher_drivers_b200.hpp -- the two per-frame drivers of the reference's ORBmatcher on top of liborbx, for use INSIDE
// the reference tree (it includes the reference's own orbmatcher.hpp / orbframe.hpp / orbmappoint.hpp):
//
//     ORBmatcherB200 matcher(0.9f, true);          // instead of  ORBmatcher matcher(0.9, true);   (src/tracking.cpp)
//     matcher.SearchByProjection(m_currentFrame, m_lastFrame, th, sensor == MONOCULAR);            // TrackWithMotionModel
//     matcher.SearchByProjection(m_currentFrame, m_localMapPoints, th);                            // SearchLocalPoints
//     matcher.SearchByBoW(m_referenceKeyFrame, m_currentFrame, matches);                           // TrackReferenceKeyFrame
//
// ORBmatcherB200 derives from ORBmatcher, keeps its constructor arguments, thresholds and protected helpers
// (RadiusByViewingCos, ComputeThreeMaxima) and hides all eleven drivers:
//   * SearchByProjection(frame, map points, th)            src/orbmatcher.cpp:42-124   -> one orbm_search_by_projection call
//     (frame grid, GetFeaturesInArea, candidate loop and acceptance on the device; the sequential rule "a key point that an
//     observed map point was stored on earlier in the call is skipped", :87-89 after :121, is iterated to its fixpoint);
//   * SearchByProjection(CurrentFrame, LastFrame, th, mono) src/orbmatcher.cpp:1337-1483 -> the projections are computed on
//     the host with the reference's own matrix expressions, ONE orbm_area_distances call replaces every
//     GetFeaturesInArea + DescriptorDistance of the loop, and the loop itself (its exclusion rule depends on the matches
//     made earlier in the same call, :1412-1414) runs on the host over the returned lists.
//   * SearchByBoW(keyFrame, frame, matches)                 src/orbmatcher.cpp:164-292   -> one orbm_distance_csr call for all
//     (key-frame feature, frame feature of the same vocabulary node) pairs, the sequential loop on the host;
//   * SearchByBoW(keyFrame1, keyFrame2, matches12)          src/orbmatcher.cpp:531-663   -> the same between two key frames (loop
//     closing), candidates restricted to the features of key frame 2 that carry a usable map point;
//   * SearchForTriangulation(keyFrame1, keyFrame2, F12, pairs, onlyStereo) src/orbmatcher.cpp:665-831 -> the same node walk between
//     the untracked features of two key frames, epipolar tests on the host;
//   * SearchByProjection(keyFrame, Scw, points, matched, th) :294-409, Fuse(keyFrame, points, th) :833-982 and
//     Fuse(keyFrame, Scw, points, th, replace) :984-1108 -> projections on the host, ONE orbm_area_distances call over the key
//     frame's grid, the sequential loops (with their Replace / AddObservingKeyframe side effects) on the host;
//   * SearchBySim3(keyFrame1, keyFrame2, matches12, s12, R12, t12, th) :1110-1335 -> one orbm_area_distances call per direction;
//   * SearchByProjection(CurrentFrame, keyFrame, found, th, d) src/orbmatcher.cpp:1485-1616 -> the same with the key frame's
//     map points and PredictScale (relocalisation);
//   * SearchForInitialization(F1, F2, prevMatched, matches)  src/orbmatcher.cpp:411-528   -> one orbm_area_distances call.
// Results are identical to the base class: tests/test_gpu_drivers.py runs both classes on the same reference frames.
#ifndef ORBMATCHER_DRIVERS_B200_HPP
#define ORBMATCHER_DRIVERS_B200_HPP

#include <climits>
#include <cstring>
#include <cmath>
#include <memory>
#include <set>
#include <utility>
#include <vector>

#include <orbmatcher.hpp>
#include <orbkeyframe.hpp>

#include "orbmatcher_b200.hpp"

class ORBmatcherB200 : public ORBmatcher {
 public:
  ORBmatcherB200(float nnratio = 0.6f, bool checkOri = true, int device = 0)
      : ORBmatcher(nnratio, checkOri), device_(device), cap_(8192), gpu_(new orbslam_b200::HammingMatcher(8192, 8192, device)) {}

  // ---- src/orbmatcher.cpp:42-124
  int SearchByProjection(std::shared_ptr<OrbFrame> &F, const std::vector<std::shared_ptr<OrbMapPoint>> &pMP, const float th = 3)
  {
      const bool bFactor = std::abs(th - 1.0) < 0.0000000001f;
      std::vector<int> who, level, pointMatch, assigned;
      std::vector<float> x, y, radius;
      std::vector<unsigned char> observed;          // :87-89 sees the map points stored earlier in this call (:121)
      cv::Mat pointDesc((int)pMP.size() > 0 ? (int)pMP.size() : 1, 32, CV_8U);
      for (size_t i = 0; i < pMP.size(); i++) {
          const std::shared_ptr<OrbMapPoint> &mp = pMP[i];
          if (!mp->GetTrackInView() || mp->IsCorrupt()) continue;
          float r = RadiusByViewingCos(mp->GTrackViewCos());
          if (bFactor) r *= th;
          const int lv = mp->GetTrackScaleLevel();
          mp->GetDescriptor().copyTo(pointDesc.row((int)who.size()));
          x.push_back(mp->getTrackProjX()); y.push_back(mp->getTrackProjY());
          level.push_back(lv); radius.push_back(r * F->m_scaleFactors[lv]);
          observed.push_back(mp->GetObservingKeyFrameCount() > 0);
          who.push_back((int)i);
      }
      if (who.empty() || F->N == 0) return 0;
      std::vector<unsigned char> occupied(F->N, 0);
      for (int k = 0; k < F->N; k++)
          occupied[k] = F->m_mapPoints[k] && F->m_mapPoints[k]->GetObservingKeyFrameCount() > 0;
      const int n = gpu_->SearchByProjection(F->m_undistortedKeys, F->mvuRight, occupied, F->m_descriptors, OrbFrame::m_minX,
                                             OrbFrame::m_minY, OrbFrame::m_maxX, OrbFrame::m_maxY,
                                             pointDesc.rowRange(0, (int)who.size()), x, y, level, radius, observed, mfNNratio,
                                             TH_HIGH, pointMatch, assigned);
      for (int k = 0; k < F->N; k++)
          if (assigned[k] >= 0) F->m_mapPoints[k] = pMP[who[assigned[k]]];
      return n;
  }

  // ---- src/orbmatcher.cpp:1337-1483
  int SearchByProjection(std::shared_ptr<OrbFrame> &CurrentFrame, const std::shared_ptr<OrbFrame> &LastFrame, const float th, const bool bMono)
  {
      const cv::Mat Rcw = CurrentFrame->mTcw.rowRange(0, 3).colRange(0, 3);
      const cv::Mat tcw = CurrentFrame->mTcw.rowRange(0, 3).col(3);
      const cv::Mat twc = -Rcw.t() * tcw;
      const cv::Mat Rlw = LastFrame->mTcw.rowRange(0, 3).colRange(0, 3);
      const cv::Mat tlw = LastFrame->mTcw.rowRange(0, 3).col(3);
      const cv::Mat tlc = Rlw * twc + tlw;
      const bool bForward = tlc.at<float>(2) > CurrentFrame->mb && !bMono;
      const bool bBackward = -tlc.at<float>(2) > CurrentFrame->mb && !bMono;

      // pass 1: the windows of all last-frame map points that project into the image (:1360-1399)
      struct Query { int i; float u, invzc, radius; };
      std::vector<Query> q;
      std::vector<float> qx, qy, qr;
      std::vector<int> l0, l1;
      for (int i = 0; i < LastFrame->N; i++) {
          const std::shared_ptr<OrbMapPoint> &pMP = LastFrame->m_mapPoints[i];
          if (!pMP || LastFrame->m_outliers[i]) continue;
          const cv::Mat x3Dc = Rcw * pMP->GetWorldPosition() + tcw;
          const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
          const float invzc = static_cast<float>(1.0 / x3Dc.at<float>(2));
          if (invzc < 0) continue;
          const float u = CurrentFrame->fx * xc * invzc + CurrentFrame->cx;
          const float v = CurrentFrame->fy * yc * invzc + CurrentFrame->cy;
          if (u < CurrentFrame->m_minX || u > CurrentFrame->m_maxX || v < CurrentFrame->m_minY || v > CurrentFrame->m_maxY) continue;
          const int oct = LastFrame->m_keys[i].octave;
          const float radius = th * CurrentFrame->m_scaleFactors[oct];
          q.push_back(Query{i, u, invzc, radius});
          qx.push_back(u); qy.push_back(v); qr.push_back(radius);
          // GetFeaturesInArea(u, v, r, minLevel = -1, maxLevel = -1) defaults, :1392-1397
          if (bForward) { l0.push_back(oct); l1.push_back(-1); }
          else if (bBackward) { l0.push_back(0); l1.push_back(oct); }
          else { l0.push_back(oct - 1); l1.push_back(oct + 1); }
      }
      if (q.empty() || CurrentFrame->N == 0) return 0;

      // one device call: every window's feature list in the reference's order + the distance of every feature.  A map point
      // without a descriptor takes the reference's repair path (:1402-1407) and is skipped; its row stays zero.
      cv::Mat qd((int)q.size(), 32, CV_8U);
      std::vector<char> hasDesc(q.size(), 1);
      for (size_t k = 0; k < q.size(); k++) {
          cv::Mat d = LastFrame->m_mapPoints[q[k].i]->GetDescriptor();
          if (d.rows == 0) { hasDesc[k] = 0; for (int b = 0; b < 32; b++) qd.ptr(static_cast<int>(k))[b] = 0; }
          else d.copyTo(qd.row(static_cast<int>(k)));
      }
      std::vector<int> offsets, indices, dist;
      gpu_->AreaDistances(CurrentFrame->m_undistortedKeys, CurrentFrame->m_descriptors, OrbFrame::m_minX, OrbFrame::m_minY,
                          OrbFrame::m_maxX, OrbFrame::m_maxY, qd, qx, qy, qr, l0, l1, offsets, indices, dist);

      // pass 2: the reference's loop over the lists (:1399-1450)
      int nmatches = 0;
      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      for (size_t k = 0; k < q.size(); k++) {
          if (offsets[k] == offsets[k + 1]) continue;
          const std::shared_ptr<OrbMapPoint> &pMP = LastFrame->m_mapPoints[q[k].i];
          if (!hasDesc[k]) {
              // no descriptor when the windows were gathered: the reference repairs the point and skips it (:1402-1407); if an
              // earlier entry of this very call already repaired it, the reference would match it, so its distances are
              // evaluated here on the host
              const cv::Mat d = pMP->GetDescriptor();
              if (d.rows == 0) { pMP->ComputeDistinctiveDescriptors(); continue; }
              for (int e = offsets[k]; e < offsets[k + 1]; e++) dist[e] = DescriptorDistance(d, CurrentFrame->m_descriptors.row(indices[e]));
          }
          int bestDist = 256, bestIdx2 = -1;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              const int i2 = indices[e];
              if (CurrentFrame->m_mapPoints[i2] && CurrentFrame->m_mapPoints[i2]->GetObservingKeyFrameCount() > 0) continue;
              if (CurrentFrame->mvuRight[i2] > 0) {
                  const float ur = q[k].u - CurrentFrame->mbf * q[k].invzc;
                  const float er = static_cast<float>(fabs(ur - CurrentFrame->mvuRight[i2]));
                  if (er > q[k].radius) continue;
              }
              if (dist[e] < bestDist) { bestDist = dist[e]; bestIdx2 = i2; }
          }
          if (bestDist <= TH_HIGH) {
              CurrentFrame->m_mapPoints[bestIdx2] = pMP;
              nmatches++;
              if (mbCheckOrientation) {
                  float rot = LastFrame->m_undistortedKeys[q[k].i].angle - CurrentFrame->m_undistortedKeys[bestIdx2].angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back(bestIdx2);
              }
          }
      }
      if (mbCheckOrientation) {                             // :1454-1474
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++) {
                  CurrentFrame->m_mapPoints[rotHist[b][j]] = std::shared_ptr<OrbMapPoint>();
                  nmatches--;
              }
          }
      }
      return nmatches;
  }

  // ---- src/orbmatcher.cpp:1485-1616 (relocalisation): the key frame's map points, minus those already found, projected
  // into the current frame; window levels [predicted - 1, predicted + 1] from the reference's own PredictScale.  Projections
  // on the host, ONE orbm_area_distances call, then the loop: a key point that has received a map point earlier in the
  // same call is skipped (:1553-1554).
  int SearchByProjection(std::shared_ptr<OrbFrame> &CurrentFrame, std::shared_ptr<OrbKeyFrame> pKF,
                         const std::set<std::shared_ptr<OrbMapPoint>> &sAlreadyFound, const float th, const int ORBdist)
  {
      const cv::Mat Rcw = CurrentFrame->mTcw.rowRange(0, 3).colRange(0, 3);
      const cv::Mat tcw = CurrentFrame->mTcw.rowRange(0, 3).col(3);
      const cv::Mat Ow = -Rcw.t() * tcw;
      const std::vector<std::shared_ptr<OrbMapPoint>> vpMPs = pKF->GetMapPointMatches();
      std::vector<int> who, l0, l1;
      std::vector<float> qx, qy, qr;
      for (size_t i = 0; i < vpMPs.size(); i++) {
          const std::shared_ptr<OrbMapPoint> &pMP = vpMPs[i];
          if (!pMP || pMP->IsCorrupt() || sAlreadyFound.count(pMP)) continue;
          const cv::Mat x3Dw = pMP->GetWorldPosition();
          const cv::Mat x3Dc = Rcw * x3Dw + tcw;
          const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
          const float invzc = static_cast<float>(1.0 / x3Dc.at<float>(2));
          const float u = CurrentFrame->fx * xc * invzc + CurrentFrame->cx;
          const float v = CurrentFrame->fy * yc * invzc + CurrentFrame->cy;
          if (u < CurrentFrame->m_minX || u > CurrentFrame->m_maxX || v < CurrentFrame->m_minY || v > CurrentFrame->m_maxY) continue;
          const cv::Mat PO = x3Dw - Ow;
          const float dist3D = static_cast<float>(cv::norm(PO));
          if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
          const int level = pMP->PredictScale(dist3D, CurrentFrame);
          who.push_back((int)i);
          qx.push_back(u); qy.push_back(v); qr.push_back(th * CurrentFrame->m_scaleFactors[level]);
          l0.push_back(level - 1); l1.push_back(level + 1);
      }
      if (who.empty() || CurrentFrame->N == 0) return 0;
      cv::Mat qd((int)who.size(), 32, CV_8U);
      for (size_t k = 0; k < who.size(); k++) vpMPs[who[k]]->GetDescriptor().copyTo(qd.row((int)k));
      std::vector<int> offsets, indices, dist;
      gpu_->AreaDistances(CurrentFrame->m_undistortedKeys, CurrentFrame->m_descriptors, OrbFrame::m_minX, OrbFrame::m_minY,
                          OrbFrame::m_maxX, OrbFrame::m_maxY, qd, qx, qy, qr, l0, l1, offsets, indices, dist);
      int nmatches = 0;
      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      for (size_t k = 0; k < who.size(); k++) {
          int bestDist = 256, bestIdx2 = -1;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              if (CurrentFrame->m_mapPoints[indices[e]]) continue;
              if (dist[e] < bestDist) { bestDist = dist[e]; bestIdx2 = indices[e]; }
          }
          if (bestDist <= ORBdist) {
              CurrentFrame->m_mapPoints[bestIdx2] = vpMPs[who[k]];
              nmatches++;
              if (mbCheckOrientation) {
                  float rot = pKF->mvKeysUn[who[k]].angle - CurrentFrame->m_undistortedKeys[bestIdx2].angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back(bestIdx2);
              }
          }
      }
      if (mbCheckOrientation) {
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++) {
                  CurrentFrame->m_mapPoints[rotHist[b][j]] = std::shared_ptr<OrbMapPoint>();
                  nmatches--;
              }
          }
      }
      return nmatches;
  }

  // ---- src/orbmatcher.cpp:164-292: matching inside the vocabulary nodes the key frame and the frame share.  The node walk
  // and the map-point tests are the reference's; the DescriptorDistance of every (key-frame feature, frame feature of the
  // same node) pair comes from ONE orbm_distance_csr call; the best / second-best loop with its exclusion of frame features
  // matched earlier in the same call (:213-214), the ratio test and the orientation histogram run on the host.
  int SearchByBoW(std::shared_ptr<OrbKeyFrame> keyFrame, std::shared_ptr<OrbFrame> &F, std::vector<std::shared_ptr<OrbMapPoint>> &vpMapPointMatches)
  {
      const std::vector<std::shared_ptr<OrbMapPoint>> mapPointsKeyFrame = keyFrame->GetMapPointMatches();
      vpMapPointMatches = std::vector<std::shared_ptr<OrbMapPoint>>(F->N, std::shared_ptr<OrbMapPoint>());
      const OrbFeatureVector &kfVec = keyFrame->m_features;

      // pass 1: the (key-frame feature, node list) pairs in the reference's visiting order
      struct Query { unsigned int idxKF; const std::vector<unsigned int> *listF; };
      std::vector<Query> q;
      std::vector<int> offsets(1, 0), indices;
      OrbFeatureVector::const_iterator kIt = kfVec.begin(), fIt = F->mFeatVec.begin();
      const OrbFeatureVector::const_iterator kEnd = kfVec.end(), fEnd = F->mFeatVec.end();
      while (kIt != kEnd && fIt != fEnd) {
          if (kIt->first == fIt->first) {
              for (size_t iKF = 0; iKF < kIt->second.size(); iKF++) {
                  const unsigned int realIdxKF = kIt->second[iKF];
                  const std::shared_ptr<OrbMapPoint> &mp = mapPointsKeyFrame[realIdxKF];
                  if (!mp || mp->IsCorrupt()) continue;
                  q.push_back(Query{realIdxKF, &fIt->second});
                  for (size_t iF = 0; iF < fIt->second.size(); iF++) indices.push_back((int)fIt->second[iF]);
                  offsets.push_back((int)indices.size());
              }
              ++kIt; ++fIt;
          } else if (kIt->first < fIt->first) {
              kIt = kfVec.lower_bound(fIt->first);
          } else {
              fIt = F->mFeatVec.lower_bound(kIt->first);
          }
      }
      if (q.empty() || indices.empty()) return 0;
      cv::Mat qd((int)q.size(), 32, CV_8U);
      for (size_t k = 0; k < q.size(); k++) keyFrame->mDescriptors.row((int)q[k].idxKF).copyTo(qd.row((int)k));
      reserve((int)q.size(), F->N);
      std::vector<int> dist;
      gpu_->CandidateDistances(qd, F->m_descriptors, offsets, indices, dist);

      // pass 2: :205-252 over the precomputed distances
      int nmatches = 0;
      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      for (size_t k = 0; k < q.size(); k++) {
          int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              const int realIdxF = indices[e];
              if (vpMapPointMatches[realIdxF]) continue;
              const int d = dist[e];
              if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = realIdxF; }
              else if (d < bestDist2) bestDist2 = d;
          }
          if (bestDist1 <= TH_LOW && static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
              vpMapPointMatches[bestIdxF] = mapPointsKeyFrame[q[k].idxKF];
              if (mbCheckOrientation) {
                  float rot = keyFrame->mvKeysUn[q[k].idxKF].angle - F->m_keys[bestIdxF].angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back(bestIdxF);
              }
              nmatches++;
          }
      }
      if (mbCheckOrientation) {
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++) {
                  vpMapPointMatches[rotHist[b][j]] = std::shared_ptr<OrbMapPoint>();
                  nmatches--;
              }
          }
      }
      return nmatches;
  }

  // ---- src/orbmatcher.cpp:531-663 (loop closing: two key frames).  As above with a key frame on both sides: the candidates of
  // a feature of key frame 1 are the features of key frame 2 under the same vocabulary node that carry a usable map point
  // (:586-592, static); the DescriptorDistance of every such pair comes from ONE orbm_distance_csr call; the exclusion of
  // features of key frame 2 matched earlier in the call (vbMatched2, :588), the TH_LOW / ratio test (:609-611) and the
  // orientation histogram run on the host.
  int SearchByBoW(std::shared_ptr<OrbKeyFrame> pKF1, std::shared_ptr<OrbKeyFrame> pKF2, std::vector<std::shared_ptr<OrbMapPoint>> &vpMatches12)
  {
      const std::vector<std::shared_ptr<OrbMapPoint>> vpMapPoints1 = pKF1->GetMapPointMatches();
      const std::vector<std::shared_ptr<OrbMapPoint>> vpMapPoints2 = pKF2->GetMapPointMatches();
      vpMatches12 = std::vector<std::shared_ptr<OrbMapPoint>>(vpMapPoints1.size(), std::shared_ptr<OrbMapPoint>());
      std::vector<unsigned char> usable2(vpMapPoints2.size(), 0), vbMatched2(vpMapPoints2.size(), 0);
      for (size_t i = 0; i < vpMapPoints2.size(); i++) usable2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->IsCorrupt();

      // pass 1: the (feature of key frame 1, usable features of key frame 2 under the same node) lists in the reference's order
      std::vector<unsigned int> q;
      std::vector<int> offsets(1, 0), indices;
      OrbFeatureVector::const_iterator f1it = pKF1->m_features.begin(), f2it = pKF2->m_features.begin();
      const OrbFeatureVector::const_iterator f1end = pKF1->m_features.end(), f2end = pKF2->m_features.end();
      while (f1it != f1end && f2it != f2end) {
          if (f1it->first == f2it->first) {
              for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                  const unsigned int idx1 = f1it->second[i1];
                  const std::shared_ptr<OrbMapPoint> &mp1 = vpMapPoints1[idx1];
                  if (!mp1 || mp1->IsCorrupt()) continue;
                  q.push_back(idx1);
                  for (size_t i2 = 0; i2 < f2it->second.size(); i2++)
                      if (usable2[f2it->second[i2]]) indices.push_back((int)f2it->second[i2]);
                  offsets.push_back((int)indices.size());
              }
              ++f1it; ++f2it;
          } else if (f1it->first < f2it->first) {
              f1it = pKF1->m_features.lower_bound(f2it->first);
          } else {
              f2it = pKF2->m_features.lower_bound(f1it->first);
          }
      }
      if (q.empty() || indices.empty()) return 0;
      cv::Mat qd((int)q.size(), 32, CV_8U);
      for (size_t k = 0; k < q.size(); k++) pKF1->mDescriptors.row((int)q[k]).copyTo(qd.row((int)k));
      reserve((int)q.size(), pKF2->mDescriptors.rows);
      std::vector<int> dist;
      gpu_->CandidateDistances(qd, pKF2->mDescriptors, offsets, indices, dist);

      // pass 2: :580-630 over the precomputed distances
      int nmatches = 0;
      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      for (size_t k = 0; k < q.size(); k++) {
          int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              const int idx2 = indices[e];
              if (vbMatched2[idx2]) continue;
              const int d = dist[e];
              if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = idx2; }
              else if (d < bestDist2) bestDist2 = d;
          }
          if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
              vpMatches12[q[k]] = vpMapPoints2[bestIdx2];
              vbMatched2[bestIdx2] = 1;
              if (mbCheckOrientation) {
                  float rot = pKF1->mvKeysUn[q[k]].angle - pKF2->mvKeysUn[bestIdx2].angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back((int)q[k]);
              }
              nmatches++;
          }
      }
      if (mbCheckOrientation) {
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++) {
                  vpMatches12[rotHist[b][j]] = std::shared_ptr<OrbMapPoint>();
                  nmatches--;
              }
          }
      }
      return nmatches;
  }

  // ---- src/orbmatcher.cpp:665-831 (local mapping: matches between the untracked key points of two key frames, to be
  // triangulated).  Candidates of a feature of key frame 1: the features of key frame 2 under the same vocabulary node that own
  // no map point and pass the stereo filter (:719-728; vbMatched2 is never set in this fork, so the list is static).  ONE
  // orbm_distance_csr call gives every pair's DescriptorDistance; the loop with its running bestDist (:735), the epipole
  // distance (:740-746) and CheckDistEpipolarLine (:748) runs on the host in the reference's order, then the orientation
  // histogram.
  int SearchForTriangulation(std::shared_ptr<OrbKeyFrame> pKF1, std::shared_ptr<OrbKeyFrame> pKF2, cv::Mat F12,
                             std::vector<std::pair<size_t, size_t>> &vMatchedPairs, const bool bOnlyStereo)
  {
      const cv::Mat Cw = pKF1->GetCameraCenter();
      const cv::Mat R2w = pKF2->GetRotation();
      const cv::Mat t2w = pKF2->GetTranslation();
      const cv::Mat C2 = R2w * Cw + t2w;
      const float invz = 1.0f / C2.at<float>(2);
      const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
      const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;

      std::vector<unsigned char> free2((size_t)pKF2->N, 0);
      for (int i = 0; i < pKF2->N; i++)
          free2[i] = !pKF2->GetMapPoint((size_t)i) && (!bOnlyStereo || pKF2->mvuRight[i] >= 0);

      std::vector<unsigned int> q;
      std::vector<int> offsets(1, 0), indices;
      OrbFeatureVector::const_iterator f1it = pKF1->m_features.begin(), f2it = pKF2->m_features.begin();
      const OrbFeatureVector::const_iterator f1end = pKF1->m_features.end(), f2end = pKF2->m_features.end();
      while (f1it != f1end && f2it != f2end) {
          if (f1it->first == f2it->first) {
              for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                  const unsigned int idx1 = f1it->second[i1];
                  if (pKF1->GetMapPoint(idx1)) continue;
                  if (bOnlyStereo && !(pKF1->mvuRight[idx1] >= 0)) continue;
                  q.push_back(idx1);
                  for (size_t i2 = 0; i2 < f2it->second.size(); i2++)
                      if (free2[f2it->second[i2]]) indices.push_back((int)f2it->second[i2]);
                  offsets.push_back((int)indices.size());
              }
              ++f1it; ++f2it;
          } else if (f1it->first < f2it->first) {
              f1it = pKF1->m_features.lower_bound(f2it->first);
          } else {
              f2it = pKF2->m_features.lower_bound(f1it->first);
          }
      }
      vMatchedPairs.clear();
      if (q.empty() || indices.empty()) return 0;
      cv::Mat qd((int)q.size(), 32, CV_8U);
      for (size_t k = 0; k < q.size(); k++) pKF1->mDescriptors.row((int)q[k]).copyTo(qd.row((int)k));
      reserve((int)q.size(), pKF2->mDescriptors.rows);
      std::vector<int> dist;
      gpu_->CandidateDistances(qd, pKF2->mDescriptors, offsets, indices, dist);

      int nmatches = 0;
      std::vector<int> vMatches12((size_t)pKF1->N, -1);
      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      for (size_t k = 0; k < q.size(); k++) {
          const unsigned int idx1 = q[k];
          const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
          const cv::KeyPoint &kp1 = pKF1->mvKeysUn[idx1];
          int bestDist = TH_LOW, bestIdx2 = -1;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              const int idx2 = indices[e];
              const int d = dist[e];
              if (d > TH_LOW || d > bestDist) continue;
              const cv::KeyPoint &kp2 = pKF2->mvKeysUn[idx2];
              const bool bStereo2 = pKF2->mvuRight[idx2] >= 0;
              if (!bStereo1 && !bStereo2) {
                  const float distex = ex - kp2.pt.x;
                  const float distey = ey - kp2.pt.y;
                  if (distex * distex + distey * distey < 100 * pKF2->mvScaleFactors[kp2.octave]) continue;
              }
              if (CheckDistEpipolarLine(kp1, kp2, F12, pKF2)) { bestIdx2 = idx2; bestDist = d; }
          }
          if (bestIdx2 >= 0) {
              const cv::KeyPoint &kp2 = pKF2->mvKeysUn[bestIdx2];
              vMatches12[idx1] = bestIdx2;
              nmatches++;
              if (mbCheckOrientation) {
                  float rot = kp1.angle - kp2.angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back((int)idx1);
              }
          }
      }
      if (mbCheckOrientation) {
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++) {
                  vMatches12[rotHist[b][j]] = -1;
                  nmatches--;
              }
          }
      }
      vMatchedPairs.reserve(nmatches > 0 ? nmatches : 0);
      for (size_t i = 0; i < vMatches12.size(); i++)
          if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
      return nmatches;
  }

  // ---- the three drivers that project map points into a KEY FRAME and search its feature grid: SearchByProjection(keyFrame,
  // Scw, points, matched, th) :294-409, Fuse(keyFrame, points, th) :833-982 and Fuse(keyFrame, Scw, points, th, replace)
  // :984-1108.  Pass 1 (host): the reference's own projection and its geometric tests (depth, IsInImage, distance invariance,
  // viewing angle), PredictScale and the window radius -- none of which changes while the loop runs.  ONE orbm_area_distances
  // call then evaluates OrbKeyFrame::GetFeaturesInArea (orbkeyframe.cpp:625-675: the frame's grid walk without a level window)
  // for all windows, restricted to the levels [predicted - 1, predicted] the loops keep anyway, with every DescriptorDistance.
  // Pass 2 (host): the reference's loop in its own order over the lists -- the tests that DO change during the call
  // (IsCorrupt, KeyFrameInObservingKeyFrames, vpMatched, GetMapPoint) are evaluated live, and the same Replace /
  // AddObservingKeyframe / AddMapPoint calls are made.
  // A key frame indexes its windows from (int)m_minX (orbkeyframe.cpp:630) while its grid was filled with the float bounds
  // (orbframe.cpp:381-393): with image bounds that are not whole numbers (distorted input) the base class runs instead.
  struct KfWindows {
      std::vector<int> slot;                      // point i -> window, or -1 when a geometric test drops it
      std::vector<float> u, v, ur;                // per window
      std::vector<int> level, offsets, indices, dist;
  };
  static bool wholeBounds(const std::shared_ptr<OrbKeyFrame> &kf)
  {
      return (float)kf->mnMinX == OrbFrame::m_minX && (float)kf->mnMinY == OrbFrame::m_minY;
  }
  void keyFrameWindows(const std::shared_ptr<OrbKeyFrame> &pKF, const cv::Mat &Rcw, const cv::Mat &tcw, const cv::Mat &Ow, const float bf,
                       const bool doubleInverse, const std::vector<std::shared_ptr<OrbMapPoint>> &pts, const float th, KfWindows &W)
  {
      const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
      W.slot.assign(pts.size(), -1);
      std::vector<float> qr;
      std::vector<int> l0;
      std::vector<size_t> who;
      for (size_t i = 0; i < pts.size(); i++) {
          const std::shared_ptr<OrbMapPoint> &pMP = pts[i];
          if (!pMP) continue;
          cv::Mat p3Dw = pMP->GetWorldPosition();
          if (p3Dw.empty()) continue;
          cv::Mat p3Dc = Rcw * p3Dw + tcw;
          if (p3Dc.at<float>(2) < 0.0f) continue;
          const float invz = doubleInverse ? static_cast<const float>(1.0 / p3Dc.at<float>(2)) : 1 / p3Dc.at<float>(2);
          const float x = p3Dc.at<float>(0) * invz;
          const float y = p3Dc.at<float>(1) * invz;
          const float u = fx * x + cx;
          const float v = fy * y + cy;
          if (!pKF->IsInImage(u, v)) continue;
          const float maxDistance = pMP->GetMaxDistanceInvariance();
          const float minDistance = pMP->GetMinDistanceInvariance();
          cv::Mat PO = p3Dw - Ow;
          const float dist3D = static_cast<const float>(cv::norm(PO));
          if (dist3D < minDistance || dist3D > maxDistance) continue;
          cv::Mat Pn = pMP->GetMeanViewingDirection();
          if (PO.dot(Pn) < 0.5 * dist3D) continue;
          const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
          W.slot[i] = (int)who.size();
          who.push_back(i);
          W.u.push_back(u); W.v.push_back(v); W.ur.push_back(u - bf * invz);
          W.level.push_back(nPredictedLevel);
          qr.push_back(th * pKF->mvScaleFactors[nPredictedLevel]);
          l0.push_back(nPredictedLevel - 1);
      }
      W.offsets.assign(1, 0);
      if (who.empty() || pKF->N == 0) { W.offsets.assign(who.size() + 1, 0); return; }
      cv::Mat qd((int)who.size(), 32, CV_8U);
      for (size_t k = 0; k < who.size(); k++) {
          const cv::Mat d = pts[who[k]]->GetDescriptor();
          if (d.empty()) std::memset(qd.ptr((int)k), 0, 32);
          else d.copyTo(qd.row((int)k));
      }
      reserve((int)who.size(), pKF->N);
      gpu_->AreaDistances(pKF->mvKeysUn, pKF->mDescriptors, OrbFrame::m_minX, OrbFrame::m_minY, OrbFrame::m_maxX, OrbFrame::m_maxY,
                          qd, W.u, W.v, qr, l0, W.level, W.offsets, W.indices, W.dist);
  }

  // ---- src/orbmatcher.cpp:294-409 (loop closing)
  int SearchByProjection(std::shared_ptr<OrbKeyFrame> pKF, cv::Mat Scw, const std::vector<std::shared_ptr<OrbMapPoint>> &vpPoints,
                         std::vector<std::shared_ptr<OrbMapPoint>> &vpMatched, int th)
  {
      if (!wholeBounds(pKF)) return ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th);
      cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
      const float scw = static_cast<const float>(sqrt(sRcw.row(0).dot(sRcw.row(0))));
      cv::Mat Rcw = sRcw / scw;
      cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
      cv::Mat Ow = -Rcw.t() * tcw;
      std::set<std::shared_ptr<OrbMapPoint>> spAlreadyFound(vpMatched.begin(), vpMatched.end());
      spAlreadyFound.erase(std::shared_ptr<OrbMapPoint>());
      KfWindows W;
      keyFrameWindows(pKF, Rcw, tcw, Ow, 0.0f, false, vpPoints, (float)th, W);
      int nmatches = 0;
      for (size_t iMP = 0; iMP < vpPoints.size(); iMP++) {
          const std::shared_ptr<OrbMapPoint> &pMP = vpPoints[iMP];
          if (pMP->IsCorrupt() || spAlreadyFound.count(pMP)) continue;
          const int k = W.slot[iMP];
          if (k < 0) continue;
          int bestDist = 256, bestIdx = -1;
          for (int e = W.offsets[k]; e < W.offsets[k + 1]; e++) {
              const int idx = W.indices[e];
              if (vpMatched[idx]) continue;
              if (W.dist[e] < bestDist) { bestDist = W.dist[e]; bestIdx = idx; }
          }
          if (bestDist <= TH_LOW) {
              vpMatched[bestIdx] = pMP;
              nmatches++;
          }
      }
      return nmatches;
  }

  // ---- src/orbmatcher.cpp:833-982 (local mapping: SearchInNeighbors)
  int Fuse(std::shared_ptr<OrbKeyFrame> pKF, const std::vector<std::shared_ptr<OrbMapPoint>> &vpMapPoints, const float th = 3.0)
  {
      if (!wholeBounds(pKF)) return ORBmatcher::Fuse(pKF, vpMapPoints, th);
      cv::Mat Rcw = pKF->GetRotation();
      cv::Mat tcw = pKF->GetTranslation();
      cv::Mat Ow = pKF->GetCameraCenter();
      KfWindows W;
      keyFrameWindows(pKF, Rcw, tcw, Ow, pKF->mbf, false, vpMapPoints, th, W);
      int nFused = 0;
      for (size_t i = 0; i < vpMapPoints.size(); i++) {
          std::shared_ptr<OrbMapPoint> pMP = vpMapPoints[i];
          if (!pMP) continue;
          if (pMP->IsCorrupt() || pMP->KeyFrameInObservingKeyFrames(pKF)) continue;
          const int k = W.slot[i];
          if (k < 0) continue;
          const float u = W.u[k], v = W.v[k], ur = W.ur[k];
          int bestDist = 256, bestIdx = -1;
          for (int e = W.offsets[k]; e < W.offsets[k + 1]; e++) {
              const int idx = W.indices[e];
              const cv::KeyPoint &kp = pKF->mvKeysUn[idx];
              const int &kpLevel = kp.octave;
              if (pKF->mvuRight[idx] >= 0) {
                  const float &kpx = kp.pt.x;
                  const float &kpy = kp.pt.y;
                  const float &kpr = pKF->mvuRight[idx];
                  const float ex = u - kpx;
                  const float ey = v - kpy;
                  const float er = ur - kpr;
                  const float e2 = ex * ex + ey * ey + er * er;
                  if (e2 * pKF->mvInvLevelSigma2[kpLevel] > 7.8) continue;
              } else {
                  const float &kpx = kp.pt.x;
                  const float &kpy = kp.pt.y;
                  const float ex = u - kpx;
                  const float ey = v - kpy;
                  const float e2 = ex * ex + ey * ey;
                  if (e2 * pKF->mvInvLevelSigma2[kpLevel] > 5.99) continue;
              }
              if (W.dist[e] < bestDist) { bestDist = W.dist[e]; bestIdx = idx; }
          }
          if (bestDist <= TH_LOW) {
              std::shared_ptr<OrbMapPoint> pMPinKF = pKF->GetMapPoint(bestIdx);
              if (pMPinKF) {
                  if (!pMPinKF->IsCorrupt()) {
                      if (pMPinKF->GetObservingKeyframes() > pMP->GetObservingKeyframes()) pMP->Replace(pMPinKF);
                      else pMPinKF->Replace(pMP);
                  }
              } else {
                  pMP->AddObservingKeyframe(pKF, bestIdx);
                  pKF->AddMapPoint(pMP, bestIdx);
              }
              nFused++;
          }
      }
      return nFused;
  }

  // ---- src/orbmatcher.cpp:984-1108 (loop closing: SearchAndFuse)
  int Fuse(std::shared_ptr<OrbKeyFrame> pKF, cv::Mat Scw, const std::vector<std::shared_ptr<OrbMapPoint>> &vpPoints, float th,
           std::vector<std::shared_ptr<OrbMapPoint>> &vpReplacePoint)
  {
      if (!wholeBounds(pKF)) return ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint);
      cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
      const float scw = static_cast<const float>(sqrt(sRcw.row(0).dot(sRcw.row(0))));
      cv::Mat Rcw = sRcw / scw;
      cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
      cv::Mat Ow = -Rcw.t() * tcw;
      const std::set<std::shared_ptr<OrbMapPoint>> spAlreadyFound = pKF->GetMapPoints();
      KfWindows W;
      keyFrameWindows(pKF, Rcw, tcw, Ow, 0.0f, true, vpPoints, th, W);
      int nFused = 0;
      for (size_t iMP = 0; iMP < vpPoints.size(); iMP++) {
          std::shared_ptr<OrbMapPoint> pMP = vpPoints[iMP];
          if (pMP->IsCorrupt() || spAlreadyFound.count(pMP)) continue;
          const int k = W.slot[iMP];
          if (k < 0) continue;
          int bestDist = INT_MAX, bestIdx = -1;
          for (int e = W.offsets[k]; e < W.offsets[k + 1]; e++)
              if (W.dist[e] < bestDist) { bestDist = W.dist[e]; bestIdx = W.indices[e]; }
          if (bestDist <= TH_LOW) {
              std::shared_ptr<OrbMapPoint> pMPinKF = pKF->GetMapPoint(bestIdx);
              if (pMPinKF) {
                  if (!pMPinKF->IsCorrupt()) vpReplacePoint[iMP] = pMPinKF;
              } else {
                  pMP->AddObservingKeyframe(pKF, bestIdx);
                  pKF->AddMapPoint(pMP, bestIdx);
              }
              nFused++;
          }
      }
      return nFused;
  }

  // ---- src/orbmatcher.cpp:1110-1335 (loop closing: matches between the map points of two key frames under a similarity).
  // Both directions are static loops (nothing found earlier in the call excludes anything later): the map points of key frame 1
  // are projected into key frame 2 and vice versa on the host with the reference's expressions, each direction is ONE
  // orbm_area_distances call, the first least distance of every window is taken (:1207-1218, strict '<'), and the agreement
  // check (:1315-1330) follows.
  int SearchBySim3(std::shared_ptr<OrbKeyFrame> pKF1, std::shared_ptr<OrbKeyFrame> pKF2, std::vector<std::shared_ptr<OrbMapPoint>> &vpMatches12,
                   const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
  {
      if (!wholeBounds(pKF1) || !wholeBounds(pKF2)) return ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th);
      const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
      cv::Mat R1w = pKF1->GetRotation();
      cv::Mat t1w = pKF1->GetTranslation();
      cv::Mat R2w = pKF2->GetRotation();
      cv::Mat t2w = pKF2->GetTranslation();
      cv::Mat sR12 = s12 * R12;
      cv::Mat sR21 = (1.0 / s12) * R12.t();
      cv::Mat t21 = -sR21 * t12;
      const std::vector<std::shared_ptr<OrbMapPoint>> vpMapPoints1 = pKF1->GetMapPointMatches();
      const int N1 = (int)vpMapPoints1.size();
      const std::vector<std::shared_ptr<OrbMapPoint>> vpMapPoints2 = pKF2->GetMapPointMatches();
      const int N2 = (int)vpMapPoints2.size();
      std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
      for (int i = 0; i < N1; i++) {
          std::shared_ptr<OrbMapPoint> pMP = vpMatches12[i];
          if (pMP) {
              vbAlreadyMatched1[i] = true;
              int idx2 = pMP->GetObeservationIndexOfKeyFrame(pKF2);
              if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
          }
      }
      // one direction: the map points `from` (already-matched ones skipped) through (Rw, tw) and (sR, t) into `into`
      auto direction = [&](const std::vector<std::shared_ptr<OrbMapPoint>> &from, const std::vector<bool> &already, const cv::Mat &Rw,
                           const cv::Mat &tw, const cv::Mat &sR, const cv::Mat &t, const std::shared_ptr<OrbKeyFrame> &into,
                           std::vector<int> &match) {
          match.assign(from.size(), -1);
          std::vector<int> who, l0, l1, offsets, indices, dist;
          std::vector<float> qx, qy, qr;
          for (size_t i = 0; i < from.size(); i++) {
              const std::shared_ptr<OrbMapPoint> &pMP = from[i];
              if (!pMP || already[i]) continue;
              if (pMP->IsCorrupt()) continue;
              cv::Mat p3Dw = pMP->GetWorldPosition();
              cv::Mat p3Da = Rw * p3Dw + tw;
              cv::Mat p3Db = sR * p3Da + t;
              if (p3Db.at<float>(2) < 0.0) continue;
              const float invz = static_cast<const float>(1.0 / p3Db.at<float>(2));
              const float x = p3Db.at<float>(0) * invz;
              const float y = p3Db.at<float>(1) * invz;
              const float u = fx * x + cx;
              const float v = fy * y + cy;
              if (!into->IsInImage(u, v)) continue;
              const float maxDistance = pMP->GetMaxDistanceInvariance();
              const float minDistance = pMP->GetMinDistanceInvariance();
              const float dist3D = static_cast<const float>(cv::norm(p3Db));
              if (dist3D < minDistance || dist3D > maxDistance) continue;
              const int nPredictedLevel = pMP->PredictScale(dist3D, into);
              who.push_back((int)i);
              qx.push_back(u); qy.push_back(v); qr.push_back(th * into->mvScaleFactors[nPredictedLevel]);
              l0.push_back(nPredictedLevel - 1); l1.push_back(nPredictedLevel);
          }
          if (who.empty() || into->N == 0) return;
          cv::Mat qd((int)who.size(), 32, CV_8U);
          for (size_t k = 0; k < who.size(); k++) {
              const cv::Mat d = from[who[k]]->GetDescriptor();
              if (d.empty()) std::memset(qd.ptr((int)k), 0, 32);
              else d.copyTo(qd.row((int)k));
          }
          reserve((int)who.size(), into->N);
          gpu_->AreaDistances(into->mvKeysUn, into->mDescriptors, OrbFrame::m_minX, OrbFrame::m_minY, OrbFrame::m_maxX, OrbFrame::m_maxY,
                              qd, qx, qy, qr, l0, l1, offsets, indices, dist);
          for (size_t k = 0; k < who.size(); k++) {
              int bestDist = INT_MAX, bestIdx = -1;
              for (int e = offsets[k]; e < offsets[k + 1]; e++)
                  if (dist[e] < bestDist) { bestDist = dist[e]; bestIdx = indices[e]; }
              if (bestDist <= TH_HIGH) match[who[k]] = bestIdx;
          }
      };
      std::vector<int> vnMatch1, vnMatch2;
      direction(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1);
      direction(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2);
      int nFound = 0;
      for (int i1 = 0; i1 < N1; i1++) {
          const int idx2 = vnMatch1[i1];
          if (idx2 >= 0) {
              const int idx1 = vnMatch2[idx2];
              if (idx1 == i1) {
                  vpMatches12[i1] = vpMapPoints2[idx2];
                  nFound++;
              }
          }
      }
      return nFound;
  }

  // ---- src/orbmatcher.cpp:411-528 (monocular initialisation): windows around the previously matched positions of the level-0
  // key points of F1, searched in F2 at level 0.  ONE orbm_area_distances call returns every window's features and
  // distances; the loop with its vMatchedDistance / vnMatches21 bookkeeping (a later, closer match takes a key point of F2
  // away from an earlier one) and the orientation histogram run on the host.
  int SearchForInitialization(std::shared_ptr<OrbFrame> &F1, std::shared_ptr<OrbFrame> &F2, std::vector<cv::Point2f> &vbPrevMatched,
                              std::vector<int> &vnMatches12, int windowSize = 10)
  {
      int nmatches = 0;
      const size_t n1 = F1->m_undistortedKeys.size(), n2 = F2->m_undistortedKeys.size();
      vnMatches12 = std::vector<int>(n1, -1);
      std::vector<int> who, l0;
      std::vector<float> qx, qy, qr;
      for (size_t i1 = 0; i1 < n1; i1++) {
          if (F1->m_undistortedKeys[i1].octave > 0) continue;
          who.push_back((int)i1);
          qx.push_back(vbPrevMatched[i1].x); qy.push_back(vbPrevMatched[i1].y); qr.push_back((float)windowSize);
          l0.push_back(0);
      }
      if (who.empty() || n2 == 0) return 0;
      cv::Mat qd((int)who.size(), 32, CV_8U);
      for (size_t k = 0; k < who.size(); k++) F1->m_descriptors.row(who[k]).copyTo(qd.row((int)k));
      std::vector<int> offsets, indices, dist;
      gpu_->AreaDistances(F2->m_undistortedKeys, F2->m_descriptors, OrbFrame::m_minX, OrbFrame::m_minY, OrbFrame::m_maxX,
                          OrbFrame::m_maxY, qd, qx, qy, qr, l0, l0, offsets, indices, dist);

      std::vector<int> rotHist[64];
      const int H = HISTO_LENGTH;
      const float factor = 1.0f / H;
      std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1);
      for (size_t k = 0; k < who.size(); k++) {
          const int i1 = who[k];
          int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
          for (int e = offsets[k]; e < offsets[k + 1]; e++) {
              const int i2 = indices[e], d = dist[e];
              if (vMatchedDistance[i2] <= d) continue;
              if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
              else if (d < bestDist2) bestDist2 = d;
          }
          if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio) {
              if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
              vnMatches12[i1] = bestIdx2;
              vnMatches21[bestIdx2] = i1;
              vMatchedDistance[bestIdx2] = bestDist;
              nmatches++;
              if (mbCheckOrientation) {
                  float rot = F1->m_undistortedKeys[i1].angle - F2->m_undistortedKeys[bestIdx2].angle;
                  if (rot < 0.0) rot += 360.0f;
                  int bin = static_cast<int>(round(rot * factor));
                  if (bin == H) bin = 0;
                  rotHist[bin].push_back(i1);
              }
          }
      }
      if (mbCheckOrientation) {
          int ind1 = -1, ind2 = -1, ind3 = -1;
          ComputeThreeMaxima(rotHist, H, ind1, ind2, ind3);
          for (int b = 0; b < H; b++) {
              if (b == ind1 || b == ind2 || b == ind3) continue;
              for (size_t j = 0; j < rotHist[b].size(); j++)
                  if (vnMatches12[rotHist[b][j]] >= 0) { vnMatches12[rotHist[b][j]] = -1; nmatches--; }
          }
      }
      for (size_t i1 = 0; i1 < n1; i1++)
          if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2->m_undistortedKeys[vnMatches12[i1]].pt;
      return nmatches;
  }

 private:
  // orbm_distance_csr works inside the matcher's query / train capacity: grow it when a frame is larger
  void reserve(int nq, int nt)
  {
      if (nq <= cap_ && nt <= cap_) return;
      while (cap_ < nq || cap_ < nt) cap_ *= 2;
      gpu_.reset(new orbslam_b200::HammingMatcher(cap_, cap_, device_));
  }
  int device_, cap_;
  std::shared_ptr<orbslam_b200::HammingMatcher> gpu_;
};

#endif
