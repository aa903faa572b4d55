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
//   OrbVocabulary::transform4 (unused: ComputeBoW is never called), Orbconverter::toDescriptorVector (frame_pre.hpp).
// OrbKeyFrame / OrbMap stand-ins come from mappoint_glue.cpp; ORBmatcher (DescriptorDistance, TH_HIGH / TH_LOW, the
// search functions) is the reference's own src/orbmatcher.cpp, compiled unmodified into the same library.
#include <orbframe.hpp>
#include <orbmatcher.hpp>

#include <cstring>
#include <ctime>
#include <sstream>

extern "C" {
#include "orb_oracle.h"
}

#ifdef ORBREF_DROPIN_STEREO
// libdropin2ref.so only: the body of OrbFrame::ComputeStereoMatches replaced as INTEGRATION.md section 2b says.  The
// reference's own definition in orbframe.o is weakened by the Makefile (objcopy --weaken-symbol), so this one is linked and
// the reference's unmodified stereo constructor calls it.  With a bounding box the host-side FilterKeyPoints has already
// thinned m_keys / m_descriptors at this point (CommonSetup); the device-resident results are thinned the same way first.
#include "orbframe_stereo_b200.hpp"
void OrbFrame::ComputeStereoMatches()
{
    orbslam_b200::FilterKeyPoints(*m_ORBextractorLeft, *m_ORBextractorRight, m_boundingBox);
    orbslam_b200::ComputeStereoMatches(*m_ORBextractorLeft, *m_ORBextractorRight, mbf, mb, mvuRight, m_depths);
}
#endif

void OrbVocabulary::transform4(const std::vector<cv::Mat>, OrbBowVector &, OrbFeatureVector &, int) const { abort(); }
std::vector<cv::Mat> Orbconverter::toDescriptorVector(const cv::Mat &d)
{
    std::vector<cv::Mat> v;
    for (int j = 0; j < d.rows; j++) v.push_back(d.row(j));
    return v;
}

extern "C" {

struct frameref_cfg { int nfeatures; float scale; int nlevels, ini_th, min_th; };
#define FRAMEREF_CFG_DEFINED

// Left / right: w x h, tightly packed.  cap = rows available in every per-keypoint output.  pyrL / pyrR: nlevels caller
// buffers receiving the pyramid levels tightly packed (level sizes in lw / lh).  Returns the number of left keypoints
// (nr_out = right), or -1 when cap is too small.  bbox (optional): the bounding box handed to the constructor, applied by the
// reference's FilterKeyPoints (orbframe.cpp:403-445) before the stereo matching.  grid_start[3073] / grid_items (optional): m_grid.  canonical != 0: heap addresses grow with allocation order (the tie order
// the oracle and the GPU path implement), otherwise the stock malloc order.
void orbref_canonical(int on);          // ref_glue.cpp: monotone bump allocator = canonical tie order in DistributeOctTree

int frameref_stereo(const frameref_cfg *c, int canonical, const uint8_t *left, const uint8_t *right, int w, int h, float mbf, float mb,
                    orbo_keypoint *kl, uint8_t *dl, orbo_keypoint *kr, uint8_t *dr, int cap, int *nr_out,
                    uint8_t **pyrL, uint8_t **pyrR, int *lw, int *lh, float *uRight, float *depth,
                    const float *bbox, int32_t *grid_start, int32_t *grid_items)
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
        std::array<float, 4> box = {0.f, 0.f, 0.f, 0.f};                      // FilterKeyPoints is a no-op (:405) ...
        if (bbox) box = {bbox[0], bbox[1], bbox[2], bbox[3]};                  // ... unless the caller gives a bounding box
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
            if (grid_start && grid_items) {              // m_grid as AssignFeaturesToGrid left it, cells in ix * 48 + iy order
                int t = 0;
                for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
                    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
                        grid_start[ix * FRAME_GRID_ROWS + iy] = t;
                        for (size_t k : frame.m_grid[ix][iy]) grid_items[t++] = (int32_t)k;
                    }
                grid_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = t;
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

extern "C++" std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe(int rows);   // mappoint_glue.cpp

std::shared_ptr<OrbFrame> frameref_make_frame(const frameref_cfg *c, const uint8_t *left, const uint8_t *right, int w, int h, float mbf, float mb)
{
    auto exL = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
    auto exR = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
    cv::Mat imL(h, w, CV_8UC1, (void *)left, (size_t)w), imR(h, w, CV_8UC1, (void *)right, (size_t)w);
    cv::Mat K(3, 3, CV_32F), dist(4, 1, CV_32F);
    for (int i = 0; i < 9; i++) K.ptr<float>(i / 3)[i % 3] = (i % 4 == 0) ? 1.f : 0.f;
    K.at<float>(0, 0) = 700.f; K.at<float>(1, 1) = 700.f; K.at<float>(0, 2) = w * 0.5f; K.at<float>(1, 2) = h * 0.5f;
    for (int i = 0; i < 4; i++) dist.ptr<float>(i)[0] = 0.f;
    std::array<float, 4> box = {0.f, 0.f, 0.f, 0.f};
    OrbFrame::m_initialComputations = true;
    auto f = std::make_shared<OrbFrame>(imL, imR, 0.0, exL, exR, std::shared_ptr<OrbVocabulary>(), K, dist, mbf, 35.f * mbf / 700.f, box);
    f->mb = mb;
    f->ComputeStereoMatches();
    return f;
}

// The reference's own ORBmatcher::SearchByProjection(frame, map points, th) (src/orbmatcher.cpp:42-124, unmodified).
// Frame A (leftA / rightA) supplies the map points: one per key point i with i % mp_step == 0, built by the reference's
// OrbMapPoint(position, frame, map, i) constructor (descriptor = A's row i), marked in view at (x_i + dx, y_i + dy), level =
// octave_i, viewing cosine alternating between 0.9995 and 0.9.  Frame B (leftB / rightB) is searched.  Key points of B
// with idx % 7 == 3 already carry a map point that has an observation (excluded by :87-89), those with idx % 11 == 5 one
// without (not excluded).  Handed back: the map points' descriptors and tracking fields, B's descriptors / octaves /
// mvuRight, the candidate lists the reference's own GetFeaturesInArea returns for every map point (CSR, in its order),
// the radius r * scaleFactor[level] of :64-67, B's undistorted key points and image bounds, and the result: assigned[idx] = index of the map point the reference
// stored in B.m_mapPoints[idx] (-1 none, -2 the pre-assigned ones left in place), return value = its nmatches.
int frameref_search_by_projection(const frameref_cfg *c, int canonical, const uint8_t *leftA, const uint8_t *rightA, const uint8_t *leftB,
                                  const uint8_t *rightB, int w, int h, float mbf, float mb, float th, float nnratio,
                                  int mp_step, float dx, float dy, int cap, int list_cap,
                                  int *n_mp_out, uint8_t *mp_desc, float *mp_x, float *mp_radius,
                                  int *n_b_out, uint8_t *b_desc, int32_t *b_octave, float *b_uright, int32_t *b_occupied,
                                  int32_t *offsets, int32_t *indices, int32_t *assigned,
                                  orbo_keypoint *b_keys, float *mp_y, int32_t *mp_level, float *bounds,
                                  int mp_dup, int obs_mod, uint8_t *mp_observed)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    std::shared_ptr<OrbKeyFrame> kf = mpref_standin_keyframe(8);   // before the arena: its static pool outlives this call
    if (canonical) orbref_canonical(1);
    int nmatches = -1;
    {
        std::shared_ptr<OrbFrame> A = frameref_make_frame(c, leftA, rightA, w, h, mbf, mb);
        cv::Mat Tcw(4, 4, CV_32F);
        for (int i = 0; i < 16; i++) Tcw.ptr<float>(i / 4)[i % 4] = (i % 5 == 0) ? 1.f : 0.f;
        A->SetPose(Tcw);
        std::vector<std::shared_ptr<OrbMapPoint>> mps;
        for (int i = 0; i < A->N; i += mp_step)
            for (int d = 0; d < (mp_dup > 1 ? mp_dup : 1); d++) {
                cv::Mat pos(3, 1, CV_32F);
                pos.ptr<float>(0)[0] = 0.f; pos.ptr<float>(1)[0] = 0.f; pos.ptr<float>(2)[0] = 5.f;
                auto mp = std::make_shared<OrbMapPoint>(pos, A, std::shared_ptr<OrbMap>(), i);
                mp->SetTrackInView(true);
                mp->SetTrackProjX(A->m_undistortedKeys[i].pt.x + dx);
                mp->SetTrackProjY(A->m_undistortedKeys[i].pt.y + dy);
                mp->SetnTrackScaleLevel(A->m_undistortedKeys[i].octave);
                mp->SetTrackViewCos(((mps.size() / (mp_dup > 1 ? mp_dup : 1)) & 1) ? 0.9f : 0.9995f);
                // observed map points (every local map point of Tracking::SearchLocalPoints is one): the rule of :87-89 is
                // live INSIDE the call -- once stored at :121 they hide their key point from the map points after them
                if (obs_mod > 0 && mps.size() % (size_t)obs_mod != 0) mp->AddObservingKeyframe(kf, 1);
                mps.push_back(mp);
            }
        std::shared_ptr<OrbFrame> B = frameref_make_frame(c, leftB, rightB, w, h, mbf, mb);
        const int nb = B->N, nmp = (int)mps.size();
        *n_mp_out = nmp; *n_b_out = nb;
        if (nb <= cap && nmp <= cap) {
            cv::Mat pos(3, 1, CV_32F);
            for (int k = 0; k < 3; k++) pos.ptr<float>(k)[0] = 1.f;
            for (int idx = 0; idx < nb; idx++) {
                b_occupied[idx] = 0;
                if (idx % 7 == 3 || idx % 11 == 5) {
                    auto held = std::make_shared<OrbMapPoint>(pos, kf, std::shared_ptr<OrbMap>());
                    if (idx % 7 == 3) { held->AddObservingKeyframe(kf, 1); b_occupied[idx] = 1; }
                    B->m_mapPoints[idx] = held;
                }
                memcpy(b_desc + (size_t)idx * 32, B->m_descriptors.ptr(idx), 32);
                b_octave[idx] = B->m_undistortedKeys[idx].octave;
                memcpy(&b_keys[idx], &B->m_undistortedKeys[idx], sizeof(orbo_keypoint));
                b_uright[idx] = B->mvuRight[idx];
            }
            bounds[0] = OrbFrame::m_minX; bounds[1] = OrbFrame::m_minY; bounds[2] = OrbFrame::m_maxX; bounds[3] = OrbFrame::m_maxY;
            const bool bFactor = std::abs(th - 1.0) < 0.0000000001f;            // :47
            int total = 0;
            offsets[0] = 0;
            bool fits = true;
            for (int i = 0; i < nmp && fits; i++) {
                cv::Mat d = mps[i]->GetDescriptor();
                memcpy(mp_desc + (size_t)i * 32, d.ptr(0), 32);
                const int level = mps[i]->GetTrackScaleLevel();
                float r = mps[i]->GTrackViewCos() > 0.998 ? 2.5f : 4.0f;         // RadiusByViewingCos, :126-131
                if (bFactor) r *= th;                                            // :61-62
                mp_x[i] = mps[i]->getTrackProjX(); mp_y[i] = mps[i]->getTrackProjY(); mp_level[i] = level;
                mp_observed[i] = mps[i]->GetObservingKeyFrameCount() > 0;
                mp_radius[i] = r * B->m_scaleFactors[level];
                const std::vector<size_t> v = B->GetFeaturesInArea(mps[i]->getTrackProjX(), mps[i]->getTrackProjY(),
                                                                   r * B->m_scaleFactors[level], level - 1, level);
                if (total + (int)v.size() > list_cap) { fits = false; break; }
                for (size_t k = 0; k < v.size(); k++) indices[total++] = (int32_t)v[k];
                offsets[i + 1] = total;
            }
            if (fits) {
                std::vector<std::shared_ptr<OrbMapPoint>> before = B->m_mapPoints;
                ORBmatcher matcher(nnratio, true);
                nmatches = matcher.SearchByProjection(B, mps, th);
                for (int idx = 0; idx < nb; idx++) {
                    assigned[idx] = -1;
                    if (!B->m_mapPoints[idx]) continue;
                    if (B->m_mapPoints[idx] == before[idx]) { assigned[idx] = -2; continue; }
                    for (int i = 0; i < nmp; i++) if (mps[i] == B->m_mapPoints[idx]) { assigned[idx] = i; break; }
                }
            }
        }
    }
    if (canonical) orbref_canonical(0);
    std::cout.rdbuf(old);
    return nmatches;
}

// The reference's own ORBmatcher::DescriptorDistance (src/orbmatcher.cpp:1662-1677) on n pairs of 32-byte rows.
void frameref_descriptor_distance(const uint8_t *a, const uint8_t *b, int n, int32_t *out)
{
    for (int i = 0; i < n; i++) {
        const cv::Mat ma(1, 32, CV_8U, (void *)(a + (size_t)i * 32), 32), mb(1, 32, CV_8U, (void *)(b + (size_t)i * 32), 32);
        out[i] = ORBmatcher::DescriptorDistance(ma, mb);
    }
}

// Wall time of the reference's own stereo OrbFrame constructor (src/orbframe.cpp:61-88: two extraction threads, CommonSetup,
// ComputeStereoMatches, AssignFeaturesToGrid) with two long-lived extractors, as Tracking holds them: one warm-up
// construction, then the mean of `reps` constructions in microseconds.  In libframeref.so everything is the reference's CPU
// code; in libdropinref.so the extractor is the liborbx drop-in; in libdropin2ref.so ComputeStereoMatches is too.
double frameref_time_constructor(const frameref_cfg *c, const uint8_t *left, const uint8_t *right, int w, int h, float mbf, int reps)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    double us = -1;
    {
        auto exL = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
        auto exR = std::make_shared<OrbExtractor>(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
        cv::Mat imL(h, w, CV_8UC1, (void *)left, (size_t)w), imR(h, w, CV_8UC1, (void *)right, (size_t)w);
        cv::Mat K(3, 3, CV_32F), dist(4, 1, CV_32F);
        for (int i = 0; i < 9; i++) K.ptr<float>(i / 3)[i % 3] = (i % 4 == 0) ? 1.f : 0.f;
        K.at<float>(0, 0) = 700.f; K.at<float>(1, 1) = 700.f; K.at<float>(0, 2) = w * 0.5f; K.at<float>(1, 2) = h * 0.5f;
        for (int i = 0; i < 4; i++) dist.ptr<float>(i)[0] = 0.f;
        std::array<float, 4> box = {0.f, 0.f, 0.f, 0.f};
        OrbFrame::m_initialComputations = true;
        { OrbFrame warm(imL, imR, 0.0, exL, exR, std::shared_ptr<OrbVocabulary>(), K, dist, mbf, 35.f * mbf / 700.f, box); }
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        for (int r = 0; r < reps; r++) { OrbFrame f(imL, imR, 0.0, exL, exR, std::shared_ptr<OrbVocabulary>(), K, dist, mbf, 35.f * mbf / 700.f, box); }
        clock_gettime(CLOCK_MONOTONIC, &b);
        us = ((b.tv_sec - a.tv_sec) * 1e6 + (b.tv_nsec - a.tv_nsec) * 1e-3) / reps;
    }
    std::cout.rdbuf(old);
    return us;
}

} // extern "C"
