// frame_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own OrbFrame::ComputeStereoMatches (src/orbframe.cpp:511-705, compiled UNMODIFIED by the `ref`
// target of oracle/Makefile against the cv:: header shim) on an image pair given from Python.  The frame is built by the
// reference's own stereo constructor (orbframe.cpp:61-88): two reference OrbExtractors (src/orbextractor.cpp, unmodified;
// cv:: primitives from the C oracle, ref_glue.cpp) extract left and right in two threads, then ComputeStereoMatches runs.
// The constructor calls it while mb is still 0 (mb is assigned afterwards, :86), so the glue sets mb and calls the
// reference's ComputeStereoMatches once more; keys, descriptors, both pyramids and the results are handed back so that
// the restatement can be run on exactly the same inputs.
// Stand-ins defined here because their own translation units pull in the whole SLAM system:
//   ORBmatcher::TH_HIGH / TH_LOW / HISTO_LENGTH (values of src/orbmatcher.cpp:36-38), OrbVocabulary::transform4 (unused:
//   ComputeBoW is never called), Orbconverter::toDescriptorVector (declared in frame_pre.hpp).
// OrbKeyFrame / OrbMap stand-ins and ORBmatcher::DescriptorDistance come from mappoint_glue.cpp.
#include <orbframe.hpp>
#include <orbmatcher.hpp>

#include <cstring>
#include <sstream>

extern "C" {
#include "orb_oracle.h"
}

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
void OrbVocabulary::transform4(const std::vector<cv::Mat>, OrbBowVector &, OrbFeatureVector &, int) const { abort(); }
std::vector<cv::Mat> Orbconverter::toDescriptorVector(const cv::Mat &d)
{
    std::vector<cv::Mat> v;
    for (int j = 0; j < d.rows; j++) v.push_back(d.row(j));
    return v;
}

extern "C" {

struct frameref_cfg { int nfeatures; float scale; int nlevels, ini_th, min_th; };

// Left / right: w x h, tightly packed.  cap = rows available in every per-keypoint output.  pyrL / pyrR: nlevels caller
// buffers receiving the pyramid levels tightly packed (level sizes in lw / lh).  Returns the number of left keypoints
// (nr_out = right), or -1 when cap is too small.  canonical != 0: heap addresses grow with allocation order (the tie order
// the oracle and the GPU path implement), otherwise the stock malloc order.
void orbref_canonical(int on);          // ref_glue.cpp: monotone bump allocator = canonical tie order in DistributeOctTree

int frameref_stereo(const frameref_cfg *c, int canonical, const uint8_t *left, const uint8_t *right, int w, int h, float mbf, float mb,
                    orbo_keypoint *kl, uint8_t *dl, orbo_keypoint *kr, uint8_t *dr, int cap, int *nr_out,
                    uint8_t **pyrL, uint8_t **pyrR, int *lw, int *lh, float *uRight, float *depth)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());                       // the extractor constructor prints (orbextractor.cpp:491)
    int n = -1;
    if (canonical) orbref_canonical(1);
    {
        auto exL = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
        auto exR = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
        cv::Mat imL(h, w, CV_8UC1, (void *)left, (size_t)w), imR(h, w, CV_8UC1, (void *)right, (size_t)w);
        cv::Mat K(3, 3, CV_32F), dist(4, 1, CV_32F);
        for (int i = 0; i < 9; i++) K.ptr<float>(i / 3)[i % 3] = (i % 4 == 0) ? 1.f : 0.f;
        K.at<float>(0, 0) = 700.f; K.at<float>(1, 1) = 700.f; K.at<float>(0, 2) = w * 0.5f; K.at<float>(1, 2) = h * 0.5f;
        for (int i = 0; i < 4; i++) dist.ptr<float>(i)[0] = 0.f;               // no distortion: UndistortKeyPoints returns early
        std::array<float, 4> box = {0.f, 0.f, 0.f, 0.f};                      // FilterKeyPoints is a no-op (:405)
        OrbFrame::m_initialComputations = true;
        OrbFrame frame(imL, imR, 0.0, exL, exR, std::shared_ptr<OrbVocabulary>(), K, dist, mbf, 35.f * mbf / 700.f, box);
        frame.mb = mb;
        frame.ComputeStereoMatches();                    // the reference's own function, now with the baseline the caller gave
        const int nl = frame.N, nr = (int)frame.m_keysRight.size();
        *nr_out = nr;
        if (nl <= cap && nr <= cap) {
            n = nl;
            for (int i = 0; i < nl; i++) {
                memcpy(&kl[i], &frame.m_keys[i], sizeof(orbo_keypoint));
                memcpy(dl + (size_t)i * 32, frame.m_descriptors.ptr(i), 32);
                uRight[i] = frame.mvuRight[i]; depth[i] = frame.m_depths[i];
            }
            for (int i = 0; i < nr; i++) {
                memcpy(&kr[i], &frame.m_keysRight[i], sizeof(orbo_keypoint));
                memcpy(dr + (size_t)i * 32, frame.m_descriptorsRight.ptr(i), 32);
            }
            for (int l = 0; l < c->nlevels; l++) {
                const cv::Mat &a = exL->m_vImagePyramid[l], &b = exR->m_vImagePyramid[l];
                lw[l] = a.cols; lh[l] = a.rows;
                for (int r = 0; r < a.rows; r++) {
                    memcpy(pyrL[l] + (size_t)r * a.cols, a.ptr(r), a.cols);
                    memcpy(pyrR[l] + (size_t)r * b.cols, b.ptr(r), b.cols);
                }
            }
        }
    }
    if (canonical) orbref_canonical(0);
    std::cout.rdbuf(old);
    return n;
}

} // extern "C"
