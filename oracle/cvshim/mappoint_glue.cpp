// mappoint_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own OrbMapPoint::ComputeDistinctiveDescriptors (src/orbmappoint.cpp:314-383, compiled UNMODIFIED
// by the `ref` target of oracle/Makefile against the cv:: header shim) on descriptor lists given from Python.
// orbmappoint.cpp needs the class definitions of OrbKeyFrame / OrbMap (reference headers, included as they are) but only
// six of their functions; those are defined HERE as inert stand-ins because their own translation units pull in the
// whole SLAM system (g2o, Eigen, DBoW vocabulary, OpenCV calib3d):
//   OrbKeyFrame::OrbKeyFrame     fills the members ComputeDistinctiveDescriptors reads (mDescriptors, mvuRight, bad flag)
//   OrbKeyFrame::isBad / GetCameraCenter / EraseMapPointMatch / ReplaceMapPointMatch, OrbMap::DeleteOrbMapPoint
// ORBmatcher::DescriptorDistance (src/orbmatcher.cpp:1662-1677) is forwarded to the reference's twin
// OrbDescriptor::distance (src/orbdescriptor.cpp:75-95, compiled from the reference, the same popcount).
// The include guard of orbconverter.hpp (Eigen / g2o conversions, unused here) is pre-defined by the Makefile.
//
// The reference keeps the observations in a std::map keyed by shared_ptr<OrbKeyFrame>, i.e. ordered by the key
// frames' ADDRESSES (orbmappoint.cpp:321, :328).  The glue places the key frames of one point at increasing addresses
// in one buffer, so that the reference's iteration order is the order of the list handed in.
#include <orbmappoint.hpp>
#include <orbmatcher.hpp>
#include <orbdescriptor.hpp>

#include <cstdlib>
#include <new>
#include <vector>

static cv::Mat g_pool;                 // row 0 unused: AddObservingKeyframe ignores keypoint index 0 (orbmappoint.cpp:171)
static std::vector<float> g_uright;
static bool g_bad = false;
// data of the stand-in key frame for ORBmatcher::SearchByBoW (libframeref / libdriverref only)
static std::vector<cv::KeyPoint> g_keysun;
static std::vector<std::shared_ptr<OrbMapPoint>> g_kf_mappoints;

#ifndef ORBREF_REAL_KEYFRAME
long unsigned int OrbKeyFrame::nNextId = 0;

OrbKeyFrame::OrbKeyFrame(std::shared_ptr<OrbFrame>, std::shared_ptr<OrbMap> map, std::shared_ptr<OrbKeyFrameDatabase>)
    : mnFrameId(0), mTimeStamp(0), mnGridCols(0), mnGridRows(0), mfGridElementWidthInv(0), mfGridElementHeightInv(0),
      mnTrackReferenceForFrame(0), mnFuseTargetForKF(0), mnBALocalForKF(0), mnBAFixedForKF(0), m_loopQuery(0), m_loopWords(0),
      mnRelocQuery(0), mnRelocWords(0), mnBAGlobalForKF(0),
      fx(0), fy(0), cx(0), cy(0), invfx(0), invfy(0), mbf(0), mb(0), mThDepth(0), N(g_pool.rows),
      mvKeys(), mvKeysUn(g_keysun), mvuRight(g_uright), mvDepth(), mDescriptors(g_pool), m_bagOfWords(), m_features(),
      mnScaleLevels(0), mfScaleFactor(0), mfLogScaleFactor(0), mvScaleFactors(), mvLevelSigma2(), mvInvLevelSigma2(),
      mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0), mK(),
      m_mapPoints(g_kf_mappoints), m_keyFrameDatabase(), m_orbVocabulary(), m_isFirstConnection(true), m_parent(),
      m_shoulNotBeErased(false), m_shouldBeErased(false), m_isBad(g_bad), mHalfBaseline(0), m_map(map)
{
    m_id = nNextId++;
}
bool OrbKeyFrame::isBad() { return m_isBad; }
cv::Mat OrbKeyFrame::GetCameraCenter() { return cv::Mat(); }
void OrbKeyFrame::EraseMapPointMatch(const size_t &) {}
void OrbKeyFrame::ReplaceMapPointMatch(const size_t &, std::shared_ptr<OrbMapPoint>) {}
#endif
void OrbMap::DeleteOrbMapPoint(std::shared_ptr<OrbMapPoint>) {}
#ifdef ORBREF_WITH_MATCHER
#ifndef ORBREF_REAL_KEYFRAME
// libframeref.so links the reference's own src/orbmatcher.cpp (the real DescriptorDistance); the key-frame members that
// translation unit references but the functions driven here never reach are inert as well
void OrbKeyFrame::AddMapPoint(std::shared_ptr<OrbMapPoint>, const size_t &) {}
std::shared_ptr<OrbMapPoint> OrbKeyFrame::GetMapPoint(const size_t &) { return std::shared_ptr<OrbMapPoint>(); }
cv::Mat OrbKeyFrame::GetRotation() { return cv::Mat(); }
cv::Mat OrbKeyFrame::GetTranslation() { return cv::Mat(); }
std::set<std::shared_ptr<OrbMapPoint>> OrbKeyFrame::GetMapPoints() { return std::set<std::shared_ptr<OrbMapPoint>>(); }
std::vector<std::shared_ptr<OrbMapPoint>> OrbKeyFrame::GetMapPointMatches() { return m_mapPoints; }   // per key frame: a copy of what was registered when it was built
std::vector<size_t> OrbKeyFrame::GetFeaturesInArea(const float &, const float &, const float &) const { return std::vector<size_t>(); }
bool OrbKeyFrame::IsInImage(const float &, const float &) const { return false; }
// a stand-in key frame with `rows` key points (no stereo coordinate), for map points that need an observation
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe(int rows)
{
    g_pool.create(rows, 32, CV_8U);
    g_uright.assign((size_t)rows, -1.0f);
    g_bad = false;
    return std::make_shared<OrbKeyFrame>(std::shared_ptr<OrbFrame>(), std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
}
// a stand-in key frame carrying key points, descriptors and map points (what SearchByBoW reads, orbmatcher.cpp:167-205);
// GetMapPointMatches() hands out the map points registered here until the next call
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe_with(const std::vector<cv::KeyPoint> &keysUn, const cv::Mat &descriptors,
                                                         const std::vector<std::shared_ptr<OrbMapPoint>> &mapPoints)
{
    g_pool = descriptors.clone();           // a buffer of its own: two stand-in key frames may be alive at once (SearchByBoW(KF, KF))
    g_uright.assign((size_t)descriptors.rows, -1.0f);
    g_keysun = keysUn;
    g_kf_mappoints = mapPoints;
    g_bad = false;
    std::shared_ptr<OrbKeyFrame> kf = std::make_shared<OrbKeyFrame>(std::shared_ptr<OrbFrame>(), std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
    g_keysun.clear();
    g_kf_mappoints.clear();
    return kf;
}
void mpref_standin_clear() { g_kf_mappoints.clear(); g_pool = cv::Mat(); }
#else    // ORBREF_REAL_KEYFRAME: libdriverref.so links the reference's own src/orbkeyframe.cpp, UNMODIFIED.  Key frames are built by
         // the reference's constructor from reference OrbFrames; what that translation unit needs beyond the classes already
         // linked (OrbMap / OrbKeyFrameDatabase / OrbVocabulary members only SetBadFlag and ComputeBoW reach) is inert here.
#include <orbkeyframedatabase.hpp>
#include <orbvocabulary.hpp>
void OrbMap::DeleteOrbKeyFrame(std::shared_ptr<OrbKeyFrame>) {}
void OrbKeyFrameDatabase::Erase(std::shared_ptr<OrbKeyFrame>) {}
static std::shared_ptr<OrbFrame> g_template;      // a reference frame whose calibration, scale tables and image bounds the key frames share
void mpref_set_template_frame(const std::shared_ptr<OrbFrame> &f) { g_template = f; }
// a key frame with the template frame's own content, for map points that need an observation
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe(int)
{
    return std::make_shared<OrbKeyFrame>(std::make_shared<OrbFrame>(g_template), std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
}
// a key frame carrying the given key points, descriptors and map points: a copy of the template frame with those members
// replaced goes through the reference's own key-frame constructor (orbkeyframe.cpp:29-60)
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe_with(const std::vector<cv::KeyPoint> &keysUn, const cv::Mat &descriptors,
                                                         const std::vector<std::shared_ptr<OrbMapPoint>> &mapPoints)
{
    std::shared_ptr<OrbFrame> f = std::make_shared<OrbFrame>(g_template);
    f->N = (int)keysUn.size();
    f->m_keys = keysUn; f->m_undistortedKeys = keysUn;
    f->m_descriptors = descriptors.clone();
    f->m_mapPoints = mapPoints;
    f->mvuRight.assign(keysUn.size(), -1.0f);
    f->m_depths.assign(keysUn.size(), -1.0f);
    f->mBowVec.clear(); f->mFeatVec.clear();
    return std::make_shared<OrbKeyFrame>(f, std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
}
void mpref_standin_clear() {}
#endif   // ORBREF_REAL_KEYFRAME
#else
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) { return OrbDescriptor::distance(a, b); }
#endif

#ifndef ORBREF_REAL_KEYFRAME
extern "C" {

// desc: n_desc x 32; point p observes rows indices[offsets[p] .. offsets[p+1]) (one key frame per observation; bad[k] != 0
// marks that key frame bad).  out: n_points x 32, the descriptor the reference leaves in m_descriptor; has[p] = 0 when the
// reference returned early (no usable observation) and m_descriptor stayed empty.
int mpref_distinctive(const uint8_t *desc, int n_desc, const int32_t *offsets, const int32_t *indices, const uint8_t *bad,
                      int n_points, uint8_t *out, int32_t *has)
{
    g_pool.create(n_desc + 1, 32, CV_8U);
    memset(g_pool.ptr(0), 0, 32);
    memcpy(g_pool.ptr(1), desc, (size_t)n_desc * 32);
    g_uright.assign((size_t)n_desc + 1, -1.0f);
    const std::shared_ptr<OrbFrame> noFrame;                    // the stand-in constructor ignores its frame
    for (int p = 0; p < n_points; p++) {
        const int n = offsets[p + 1] - offsets[p];
        void *buf = malloc(sizeof(OrbKeyFrame) * (size_t)(n + 2) + alignof(OrbKeyFrame));
        char *base = (char *)(((uintptr_t)buf + alignof(OrbKeyFrame) - 1) & ~(uintptr_t)(alignof(OrbKeyFrame) - 1));
        std::vector<std::shared_ptr<OrbKeyFrame>> kfs;
        g_bad = false;
        OrbKeyFrame *refKf = new (base) OrbKeyFrame(noFrame, nullptr, nullptr);
        std::shared_ptr<OrbKeyFrame> ref(refKf, [](OrbKeyFrame *k) { k->~OrbKeyFrame(); });
        cv::Mat pos(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) pos.ptr<float>(k)[0] = 0.f;
        {
            std::shared_ptr<OrbMapPoint> mp = std::make_shared<OrbMapPoint>(pos, ref, std::shared_ptr<OrbMap>());
            for (int k = 0; k < n; k++) {
                g_bad = bad && bad[offsets[p] + k];
                OrbKeyFrame *kf = new (base + sizeof(OrbKeyFrame) * (size_t)(k + 1)) OrbKeyFrame(noFrame, nullptr, nullptr);
                kfs.emplace_back(kf, [](OrbKeyFrame *q) { q->~OrbKeyFrame(); });
                mp->AddObservingKeyframe(kfs.back(), (size_t)indices[offsets[p] + k] + 1);
            }
            mp->ComputeDistinctiveDescriptors();
            cv::Mat d = mp->GetDescriptor();
            has[p] = !d.empty();
            if (!d.empty()) memcpy(out + (size_t)p * 32, d.ptr(0), 32);
        }
        kfs.clear(); ref.reset();
        free(buf);
    }
    return 0;
}

} // extern "C"
#endif   // !ORBREF_REAL_KEYFRAME (mpref_distinctive places stand-in key frames by hand)
