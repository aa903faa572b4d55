// orbvocabulary_b200.hpp -- OrbVocabulary::transform4 (reference include/orbvocabulary.hpp:64,
// src/orbvocabulary.cpp:168-201) with the per-feature descent (transform5, :203-242) on the GPU.
//
// The caller hands over the tree it already holds in OrbVocabulary::m_nodes as flat arrays (see
// include/orbx.h, orbv_create).  transform4 keeps the reference's signature and its bookkeeping:
// features whose word weight is > 0 are added to the bag-of-words vector and to the feature vector
// in feature order, then the bag-of-words vector is normalised -- exactly the loop at :186-200, with
// the call to transform5 replaced by one batched device call in front of it.  BowVector / FeatureVector
// are the reference's OrbBowVector / OrbFeatureVector (anything with clear / addWeight / normalize,
// resp. clear / addFeature).
#ifndef ORBVOCABULARY_B200_HPP
#define ORBVOCABULARY_B200_HPP

#include <opencv2/core/core.hpp>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbx.h"

namespace orbslam_b200 {

class VocabularyTransform {
 public:
  // child_off[n_nodes+1], child_ids[child_off[n_nodes]], node_desc[n_nodes*32], word_id[n_nodes] (-1 for inner nodes),
  // weight[n_nodes], L = m_L
  VocabularyTransform(const std::vector<int32_t> &childOff, const std::vector<int32_t> &childIds, const std::vector<uint8_t> &nodeDesc,
                      const std::vector<int32_t> &wordId, const std::vector<double> &weight, int L, int device = 0)
      : v_(nullptr)
  {
      int rc = orbv_create(device, (int)wordId.size(), childOff.data(), childIds.data(), nodeDesc.data(), wordId.data(), weight.data(), L, &v_);
      if (rc != ORBX_OK) {
          std::string msg = v_ ? orbv_last_error(v_) : "invalid vocabulary";
          if (v_) orbv_destroy(v_);
          v_ = nullptr;
          throw std::runtime_error("liborbx: " + msg);
      }
  }
  ~VocabularyTransform() { if (v_) orbv_destroy(v_); }
  VocabularyTransform(const VocabularyTransform &) = delete;
  VocabularyTransform &operator=(const VocabularyTransform &) = delete;

  // void OrbVocabulary::transform4(const std::vector<cv::Mat> features, OrbBowVector&, OrbFeatureVector&, int levelsUp) const
  template <class BowVector, class FeatureVector>
  void transform4(const std::vector<cv::Mat> &features, BowVector &bowVector, FeatureVector &featureVector, int levelsUp)
  {
      bowVector.clear();
      featureVector.clear();
      const int n = (int)features.size();
      if (n == 0) return;
      rows_.resize((size_t)n * 32);
      for (int i = 0; i < n; i++) std::memcpy(&rows_[(size_t)i * 32], features[i].ptr(0), 32);
      word_.resize(n); node_.resize(n); w_.resize(n);
      int rc = orbv_transform(v_, rows_.data(), n, levelsUp, word_.data(), w_.data(), node_.data());
      if (rc != ORBX_OK) throw std::runtime_error(std::string("liborbx: ") + orbv_last_error(v_));
      for (int i = 0; i < n; i++) {
          if (w_[i] > 0) {   // not stopped, :193
              bowVector.addWeight((uint32_t)word_[i], w_[i]);
              featureVector.addFeature((uint32_t)node_[i], (uint32_t)i);
          }
      }
      bowVector.normalize();
  }

  // the raw per-feature results of the last transform4 call
  const std::vector<int32_t> &words() const { return word_; }
  const std::vector<int32_t> &nodes() const { return node_; }

 private:
  orbv_vocab *v_;
  std::vector<uint8_t> rows_;
  std::vector<int32_t> word_, node_;
  std::vector<double> w_;
};

}  // namespace orbslam_b200
#endif
