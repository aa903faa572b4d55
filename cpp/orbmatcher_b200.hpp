// orbmatcher_b200.hpp -- batched Hamming matching for the loops around
// ORBmatcher::DescriptorDistance (reference include/orbmatcher.hpp:48, src/orbmatcher.cpp:1662-1677).
//
// The reference's static DescriptorDistance(a, b) stays what it is -- one 32-byte pair per call is
// below the granularity of a kernel launch and this library has no CPU path.  What moves to the
// GPU are the LOOPS that call it: brute-force best / second-best search over a descriptor set
// (orbmatcher.cpp:208-232 pattern) and bulk pair evaluation.  The acceptance rules stay on the host
// exactly as written in the reference (TH_LOW / TH_HIGH, mfNNratio; orbmatcher.cpp:36-37, :234-236).
#ifndef ORBMATCHER_B200_HPP
#define ORBMATCHER_B200_HPP

#include <opencv2/core/core.hpp>

#include <stdexcept>
#include <string>
#include <vector>

#include "orbx.h"

namespace orbslam_b200 {

class HammingMatcher {
 public:
  HammingMatcher(int maxQueries, int maxTrain, int device = 0) : m_(nullptr)
  {
      int rc = orbm_create(device, maxQueries, maxTrain, &m_);
      if (rc != ORBX_OK) {
          std::string msg = m_ ? orbm_last_error(m_) : "invalid configuration";
          if (m_) orbm_destroy(m_);
          m_ = nullptr;
          throw std::runtime_error("liborbx: " + msg);
      }
  }
  ~HammingMatcher() { if (m_) orbm_destroy(m_); }
  HammingMatcher(const HammingMatcher &) = delete;
  HammingMatcher &operator=(const HammingMatcher &) = delete;

  // map-point (train) descriptors: N x 32 CV_8U, continuous; stays resident on the GPU
  void SetTrain(const cv::Mat &train) { check(orbm_set_train(m_, train.ptr(0), train.rows)); }

  // for every row of `query`: bestIdx (lowest train index attaining the minimum, -1 if none < 256),
  // bestDist1, bestDist2 -- the values the loop at orbmatcher.cpp:208-232 leaves behind
  void KnnMatch2(const cv::Mat &query, std::vector<int> &bestIdx, std::vector<int> &bestDist1, std::vector<int> &bestDist2)
  {
      const int n = query.rows;
      bestIdx.resize(n); bestDist1.resize(n); bestDist2.resize(n);
      if (n == 0) return;
      check(orbm_knn2_resident(m_, query.ptr(0), n, bestIdx.data(), bestDist1.data(), bestDist2.data()));
  }

  // DescriptorDistance for n independent pairs (row i of a vs row i of b)
  void DescriptorDistanceBatch(const cv::Mat &a, const cv::Mat &b, std::vector<int> &dist)
  {
      dist.resize(a.rows);
      if (a.rows == 0) return;
      check(orbm_distance_pairs(m_, a.ptr(0), b.ptr(0), a.rows, dist.data()));
  }

  // OrbMapPoint::ComputeDistinctiveDescriptors (orbmappoint.cpp:314-383) for many map points at once.  pool = all
  // observed descriptors (N x 32 CV_8U, continuous); point p observes rows indices[offsets[p] .. offsets[p+1]).
  // best[p] = position in the point's list of the descriptor the reference would clone into m_descriptor, -1 for
  // a point without (good) observations.
  void DistinctiveDescriptors(const cv::Mat &pool, const std::vector<int> &offsets, const std::vector<int> &indices, std::vector<int> &best)
  {
      const int np = (int)offsets.size() - 1;
      best.assign(np > 0 ? np : 0, -1);
      if (np <= 0 || pool.rows == 0) return;
      check(orbm_distinctive(m_, pool.ptr(0), pool.rows, offsets.data(), indices.data(), np, best.data(), nullptr));
  }

  // DescriptorDistance of every (query row, candidate) entry of CSR candidate lists, in list order -- for drivers whose
  // exclusion rules depend on earlier matches of the same call (SearchByProjection, orbmatcher.cpp:42-128): evaluate all
  // distances in one launch, then replay the reference's loop on the host with ReplayBestTwo.
  void CandidateDistances(const cv::Mat &query, const cv::Mat &train, const std::vector<int> &offsets, const std::vector<int> &indices,
                          std::vector<int> &dist)
  {
      dist.assign(indices.size(), 0);
      if (indices.empty() || query.rows == 0) return;
      check(orbm_distance_csr(m_, query.ptr(0), query.rows, train.ptr(0), train.rows, offsets.data(), indices.data(), dist.data()));
  }

  // The best / second-best loop of orbmatcher.cpp:76-114 over one query's candidates with precomputed distances;
  // skip(idx) is the caller's exclusion test (:87-97), level(idx) the candidate's octave (:105, :110).
  struct BestTwo { int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1; };
  template <class Skip, class Level>
  static BestTwo ReplayBestTwo(const int *cand, const int *dist, int n, Skip skip, Level level)
  {
      BestTwo b;
      for (int k = 0; k < n; k++) {
          const int idx = cand[k];
          if (skip(idx)) continue;
          const int d = dist[k];
          if (d < b.bestDist) { b.bestDist2 = b.bestDist; b.bestDist = d; b.bestLevel2 = b.bestLevel; b.bestLevel = level(idx); b.bestIdx = idx; }
          else if (d < b.bestDist2) { b.bestLevel2 = level(idx); b.bestDist2 = d; }
      }
      return b;
  }

  // ORBmatcher::SearchByProjection(frame, map points, th) (orbmatcher.cpp:42-124) in one call, the frame's grid
  // (AssignFeaturesToGrid / GetFeaturesInArea, orbframe.cpp:192-211, :308-380) included.  The caller walks pMP once,
  // keeps the map points that pass GetTrackInView / IsCorrupt (:51-55) and appends per point: its descriptor row,
  // getTrackProjX/Y, GetTrackScaleLevel and radius = r * F->m_scaleFactors[level] with r from :58-62.  occupied[i] != 0
  // where F->m_mapPoints[i] is set and has observations (:87-89) on entry; observed[p] != 0 where map point p itself has
  // observations -- the reference stores an accepted map point at once (:121), so such a point hides its key point from the
  // map points after it; the device iterates to the fixpoint of that rule.  On return assigned[i] is the position (in the
  // arrays handed in) of the map point the reference's loop would leave in F->m_mapPoints[i], -1 for none; the return
  // value is the reference's nmatches.
  int SearchByProjection(const std::vector<cv::KeyPoint> &keysUn, const std::vector<float> &uRight,
                         const std::vector<unsigned char> &occupied, const cv::Mat &descriptors, float minX, float minY,
                         float maxX, float maxY, const cv::Mat &pointDescriptors, const std::vector<float> &projX,
                         const std::vector<float> &projY, const std::vector<int> &level, const std::vector<float> &radius,
                         const std::vector<unsigned char> &observed, float nnRatio, int thHigh, std::vector<int> &pointMatch,
                         std::vector<int> &assigned)
  {
      static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
      const int n = (int)keysUn.size(), np = pointDescriptors.rows;
      pointMatch.assign(np, -1);
      assigned.assign(n, -1);
      if (n == 0 || np == 0) return 0;
      orbm_frame_view fv;
      fv.keys = reinterpret_cast<const orbx_keypoint *>(keysUn.data());
      fv.u_right = uRight.data();
      fv.occupied = occupied.empty() ? nullptr : occupied.data();
      fv.desc = descriptors.ptr(0);
      fv.n = n;
      fv.min_x = minX; fv.min_y = minY; fv.max_x = maxX; fv.max_y = maxY;
      int32_t nmatches = 0;
      check(orbm_search_by_projection(m_, &fv, pointDescriptors.ptr(0), projX.data(), projY.data(), level.data(), radius.data(),
                                      observed.empty() ? nullptr : observed.data(), np, nnRatio, thHigh, pointMatch.data(),
                                      assigned.data(), &nmatches));
      return nmatches;
  }

  // OrbFrame::AssignFeaturesToGrid (orbframe.cpp:192-211): fills grid[ix][iy] (the reference's m_grid, 64 x 48 vectors)
  // with the key-point indices in the order the reference pushes them.
  template <class Grid>
  void AssignFeaturesToGrid(const std::vector<cv::KeyPoint> &keysUn, float minX, float minY, float maxX, float maxY, Grid &grid)
  {
      const int n = (int)keysUn.size();
      std::vector<int32_t> start(64 * 48 + 1, 0), items(n > 0 ? n : 1, 0);
      check(orbm_assign_grid(m_, reinterpret_cast<const orbx_keypoint *>(keysUn.data()), n, minX, minY, maxX, maxY, start.data(), items.data()));
      for (int ix = 0; ix < 64; ix++)
          for (int iy = 0; iy < 48; iy++) {
              auto &cell = grid[ix][iy];
              cell.clear();
              for (int e = start[ix * 48 + iy]; e < start[ix * 48 + iy + 1]; e++) cell.push_back(items[e]);
          }
  }

  // OrbFrame::GetFeaturesInArea (orbframe.cpp:308-380) for many windows of one frame at once, with DescriptorDistance of
  // every feature found -- for the window-based drivers whose acceptance is sequential (SearchForInitialization,
  // orbmatcher.cpp:411-528; SearchByProjection(CurrentFrame, LastFrame), :1337-1483; Fuse).  Window i: centre (x, y)[i],
  // half size r[i], levels [minLevel[i], maxLevel[i]] with the reference's meaning of -1.  The features of window i are
  // indices[offsets[i] .. offsets[i+1]) in the order the reference returns them; dist is filled when queryDescriptors is
  // not empty (row i against those features).
  void AreaDistances(const std::vector<cv::KeyPoint> &keysUn, const cv::Mat &descriptors, float minX, float minY, float maxX,
                     float maxY, const cv::Mat &queryDescriptors, const std::vector<float> &x, const std::vector<float> &y,
                     const std::vector<float> &r, const std::vector<int> &minLevel, const std::vector<int> &maxLevel,
                     std::vector<int> &offsets, std::vector<int> &indices, std::vector<int> &dist)
  {
      const int n = (int)keysUn.size(), nq = (int)x.size();
      offsets.assign(nq + 1, 0); indices.clear(); dist.clear();
      if (n == 0 || nq == 0) return;
      orbm_frame_view fv;
      fv.keys = reinterpret_cast<const orbx_keypoint *>(keysUn.data());
      fv.u_right = nullptr; fv.occupied = nullptr;
      fv.desc = descriptors.ptr(0);
      fv.n = n;
      fv.min_x = minX; fv.min_y = minY; fv.max_x = maxX; fv.max_y = maxY;
      const bool withDist = queryDescriptors.rows == nq;
      int cap = 16 * nq + 1024;
      for (;;) {
          indices.assign(cap, 0); dist.assign(withDist ? cap : 0, 0);
          int32_t total = 0;
          const int rc = orbm_area_distances(m_, &fv, withDist ? queryDescriptors.ptr(0) : nullptr, x.data(), y.data(), r.data(),
                                             minLevel.data(), maxLevel.data(), nq, offsets.data(), indices.data(),
                                             withDist ? dist.data() : nullptr, cap, &total);
          if (rc == ORBX_ERR_CAPACITY && total > cap) { cap = total; continue; }
          check(rc);
          indices.resize(total); if (withDist) dist.resize(total);
          return;
      }
  }

  // the reference's acceptance test (orbmatcher.cpp:234-236)
  static bool Accept(int bestDist1, int bestDist2, int th, float nnRatio)
  {
      return bestDist1 <= th && static_cast<float>(bestDist1) < nnRatio * static_cast<float>(bestDist2);
  }

 private:
  void check(int rc) { if (rc != ORBX_OK) throw std::runtime_error(std::string("liborbx: ") + orbm_last_error(m_)); }
  orbm_matcher *m_;
};

}  // namespace orbslam_b200
#endif
