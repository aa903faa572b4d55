// orbextractor_b200.hpp -- drop-in OrbExtractor for chalmers-revere/opendlv-perception-vision-orbslam2
// backed by liborbx.so (include/orbx.h) on a B200.
//
// Same class name, constructor, member functions and public data member as the reference's
// include/orbextractor.hpp:90-109, so Tracking (src/tracking.cpp:121-127), OrbFrame
// (src/orbframe.cpp:65-78,152-158,213-223,518,618-641) and everything above them compile and
// behave unchanged.  To switch: put this header's directory before the reference's include/ on
// the include path under the name orbextractor.hpp (or include it from there), drop
// src/orbextractor.cpp from the build and link -lorbx.  See INTEGRATION.md.
//
// Header-only on purpose: it needs the OpenCV headers of whatever build includes it; liborbx.so
// itself has no OpenCV dependency.  (In this repository it is compiled against oracle/cvshim in
// tests/cpp/adapter_main.cpp, because the image has no OpenCV C++ headers.)
//
// Behavioural notes, each mirroring the reference:
//  * ExtractFeatures clears `keypoints` and create()s `descriptors` (N x 32, CV_8U) or release()s
//    it when N == 0                                         (orbextractor.cpp:599-609)
//  * m_vImagePyramid[l] are host cv::Mat headers over a pinned buffer owned by the handle, valid
//    until the next ExtractFeatures on this instance -- the reference overwrites its pyramid on
//    the next call too (orbextractor.cpp:654-678).  They are ROI-free level images; the 19-px
//    REFLECT_101 frame the reference keeps around each level is never read by any caller
//    (SURVEY.md A.2).  A level leaves HBM only when somebody reads it: the member is an
//    OrbExtractor::Pyramid (indexing, at(), size(), iteration and conversion to
//    std::vector<cv::Mat>& like the reference's vector) whose element access downloads that level
//    on first use after an extraction -- monocular tracking never pays for the pyramid, the
//    reference's ComputeStereoMatches (orbframe.cpp:518, 618-641) pays for the levels it touches.
//    Define ORBX_ADAPTER_EAGER_PYRAMID to download all levels in every call instead.
//  * key points and descriptors are DMA'd into page-locked scratch of the adapter and leave it with
//    one memcpy each (cv::KeyPoint is the library's 28-byte record, checked at compile time).
//  * one instance must not be entered by two threads at once; different instances may run
//    concurrently (the stereo path does, orbframe.cpp:73-76).
//  * errors: the reference prints and continues into undefined behaviour; this adapter throws
//    std::runtime_error with the library's message.
#ifndef ORBEXTRACTOR_HPP
#define ORBEXTRACTOR_HPP

#include <opencv2/core/core.hpp>

#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbx.h"

class OrbExtractor {
 public:
  // std::vector<cv::Mat> m_vImagePyramid of the reference (orbextractor.hpp:109) with the download deferred to the first read
  class Pyramid {
   public:
    typedef cv::Mat value_type;
    typedef std::vector<cv::Mat>::iterator iterator;
    Pyramid() : m_owner(nullptr) {}
    size_t size() const { return m_mats.size(); }
    bool empty() const { return m_mats.empty(); }
    void resize(size_t n) { m_mats.resize(n); m_have.assign(n, 0); }
    cv::Mat &operator[](size_t l) { fetch(l); return m_mats[l]; }
    cv::Mat &at(size_t l) { if (l >= m_mats.size()) throw std::out_of_range("OrbExtractor::m_vImagePyramid"); fetch(l); return m_mats[l]; }
    cv::Mat &front() { return (*this)[0]; }
    cv::Mat &back() { return (*this)[m_mats.size() - 1]; }
    iterator begin() { fetchAll(); return m_mats.begin(); }
    iterator end() { fetchAll(); return m_mats.end(); }
    operator std::vector<cv::Mat> &() { fetchAll(); return m_mats; }
    void fetchAll() { for (size_t l = 0; l < m_mats.size(); l++) fetch(l); }

   private:
    friend class OrbExtractor;
    void invalidate() { std::fill(m_have.begin(), m_have.end(), 0); for (size_t l = 0; l < m_mats.size(); l++) m_mats[l] = cv::Mat(); }
    void fetch(size_t l)
    {
        if (l >= m_mats.size() || m_have[l] || !m_owner || !m_owner->m_handle || !m_owner->m_extracted) return;
        const uint8_t *p = nullptr; int w = 0, h = 0; size_t pitch = 0;
        m_owner->check(orbx_get_level(m_owner->m_handle, 0, (int)l, &p, &w, &h, &pitch));
        m_mats[l] = cv::Mat(h, w, CV_8UC1, (void *)p, pitch);
        m_have[l] = 1;
    }
    OrbExtractor *m_owner;
    std::vector<cv::Mat> m_mats;
    std::vector<char> m_have;
  };

  OrbExtractor(int nFeatures, float scaleFactor, int nLevels, int initialFastTh, int minFastTh)
      : m_vImagePyramid(), m_handle(nullptr), m_nFeatures(nFeatures), m_scaleFactor(scaleFactor), m_nLevels(nLevels),
        m_initialFastTh(initialFastTh), m_minFastTh(minFastTh), m_maxW(0), m_maxH(0), m_device(0), m_extracted(false),
        m_kps(nullptr), m_desc(nullptr), m_cap(0)
  {
      std::cout << "Making extractor" << std::endl;   // orbextractor.cpp:491
      m_vImagePyramid.m_owner = this;
      m_vImagePyramid.resize(nLevels);
  }
  ~OrbExtractor()
  {
      orbx_host_free(m_kps); orbx_host_free(m_desc);
      if (m_handle) orbx_destroy(m_handle);
  }
  OrbExtractor(const OrbExtractor &) = delete;
  OrbExtractor &operator=(const OrbExtractor &) = delete;

  /*Called by other classes to get keypoints and descriptors*/
  void ExtractFeatures(cv::InputArray image, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors)
  {
      if (image.empty()) {
          std::cout << "No Image for ORB features" << std::endl;   // orbextractor.cpp:583-585
          throw std::runtime_error("OrbExtractor::ExtractFeatures: empty image");
      }
      cv::Mat img = image.getMat();
      if (img.type() != CV_8UC1) throw std::runtime_error("OrbExtractor::ExtractFeatures: image must be CV_8UC1");  // :588
      ensureHandle(img.cols, img.rows);
      m_vImagePyramid.invalidate();
      int n = 0;
      // kp_cap == orbx_max_keypoints and page-locked arrays: the library's D2H lands directly in m_kps / m_desc
      check(orbx_extract(m_handle, img.ptr(0), img.cols, img.rows, (size_t)img.step, m_kps, m_cap, m_desc, &n));
      m_extracted = true;
      if (n == 0) {
          descriptors.release();
      } else {
          descriptors.create(n, 32, CV_8U);
          cv::Mat d = descriptors.getMat();
          if ((size_t)d.step == 32u) std::memcpy(d.ptr(0), m_desc, (size_t)n * 32);   // rows back to back
          else for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), m_desc + (size_t)i * 32, 32);
      }
      copyKeyPoints(keypoints, n);
#ifdef ORBX_ADAPTER_EAGER_PYRAMID
      fetchPyramid();
#endif
  }

  /*Getters for scale factors and other properties*/
  int getLevels() { return m_nLevels; }
  double getScaleFactor() { return m_scaleFactor; }
  std::vector<float> getScaleFactors() { return table(0); }
  std::vector<float> getInverseScaleFactors() { return table(1); }
  std::vector<float> getScaleSigmaSquares() { return table(2); }
  std::vector<float> getInverseScaleSigmaSquares() { return table(3); }

  Pyramid m_vImagePyramid;   // std::vector<cv::Mat> m_vImagePyramid; in the reference (orbextractor.hpp:109)

  // additions (not in the reference): explicit pyramid download, device selection
  void fetchPyramid() { m_vImagePyramid.fetchAll(); }
  void setDevice(int device) { m_device = device; }
  orbx_extractor *handle() { return m_handle; }   // for orbx_stereo_match and other device-resident consumers

 private:
  void check(int rc)
  {
      if (rc != ORBX_OK) throw std::runtime_error(std::string("liborbx: ") + orbx_last_error(m_handle));
  }
  void ensureHandle(int w, int h)
  {
      if (m_handle && w <= m_maxW && h <= m_maxH) return;
      if (m_handle) { orbx_destroy(m_handle); m_handle = nullptr; }
      m_extracted = false;
      orbx_config cfg = orbx_config();
      cfg.nfeatures = m_nFeatures; cfg.scale_factor = (float)m_scaleFactor; cfg.nlevels = m_nLevels;
      cfg.ini_th_fast = m_initialFastTh; cfg.min_th_fast = m_minFastTh;
      cfg.max_width = m_maxW = std::max(w, m_maxW); cfg.max_height = m_maxH = std::max(h, m_maxH);
      cfg.max_batch = 1; cfg.device = m_device;
      int rc = orbx_create(&cfg, &m_handle);
      if (rc != ORBX_OK) {
          std::string msg = m_handle ? orbx_last_error(m_handle) : "invalid configuration";
          if (m_handle) { orbx_destroy(m_handle); m_handle = nullptr; }
          m_maxW = m_maxH = 0;
          throw std::runtime_error("liborbx: " + msg);
      }
      const int cap = orbx_max_keypoints(m_handle);
      if (cap != m_cap) {
          orbx_host_free(m_kps); orbx_host_free(m_desc);
          m_kps = nullptr; m_desc = nullptr; m_cap = 0;
          void *a = nullptr, *b = nullptr;
          if (orbx_host_alloc((size_t)cap * sizeof(orbx_keypoint), &a) != ORBX_OK || orbx_host_alloc((size_t)cap * 32, &b) != ORBX_OK) {
              orbx_host_free(a); orbx_host_free(b);
              throw std::runtime_error("liborbx: no page-locked memory for the result scratch");
          }
          m_kps = (orbx_keypoint *)a; m_desc = (uint8_t *)b; m_cap = cap;
      }
  }
  // cv::KeyPoint {Point2f pt; float size, angle, response; int octave, class_id;} is the library's record field for field
  // (orbx_keypoint, include/orbx.h): one memcpy when the layouts agree, a per-field copy otherwise
  void copyKeyPoints(std::vector<cv::KeyPoint> &keypoints, int n)
  {
      keypoints.clear();
      if (n <= 0) return;
      if (sizeof(cv::KeyPoint) == sizeof(orbx_keypoint) && keyPointLayoutMatches()) {
          keypoints.resize((size_t)n);
          std::memcpy((void *)keypoints.data(), m_kps, (size_t)n * sizeof(orbx_keypoint));
          return;
      }
      keypoints.reserve((size_t)n);
      for (int i = 0; i < n; i++) {
          const orbx_keypoint &k = m_kps[i];
          keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id));
      }
  }
  static bool keyPointLayoutMatches()
  {
      const cv::KeyPoint probe(1.0f, 2.0f, 3.0f, 4.0f, 5.0f, 6, 7);
      orbx_keypoint k;
      if (sizeof(probe) != sizeof(k)) return false;
      std::memcpy(&k, &probe, sizeof k);
      return k.x == 1.0f && k.y == 2.0f && k.size == 3.0f && k.angle == 4.0f && k.response == 5.0f && k.octave == 6 && k.class_id == 7;
  }
  // the constructor tables of orbextractor.cpp:492-508 are a pure function of (scaleFactor, nLevels):
  // computed here with the same float chain so the getters work before the first image arrives
  std::vector<float> table(int which)
  {
      std::vector<float> sf(m_nLevels), s2(m_nLevels), isf(m_nLevels), is2(m_nLevels);
      sf[0] = 1.0f; s2[0] = 1.0f;
      for (int i = 1; i < m_nLevels; i++) { sf[i] = sf[i - 1] * (float)m_scaleFactor; s2[i] = sf[i] * sf[i]; }
      for (int i = 0; i < m_nLevels; i++) { isf[i] = 1.0f / sf[i]; is2[i] = 1.0f / s2[i]; }
      return which == 0 ? sf : which == 1 ? isf : which == 2 ? s2 : is2;
  }

  orbx_extractor *m_handle;
  int m_nFeatures;
  double m_scaleFactor;
  int m_nLevels;
  int m_initialFastTh;
  int m_minFastTh;
  int m_maxW, m_maxH, m_device;
  bool m_extracted;
  orbx_keypoint *m_kps;   // page-locked scratch, orbx_max_keypoints records / x 32 bytes
  uint8_t *m_desc;
  int m_cap;
};

#endif  // ORBEXTRACTOR_HPP
