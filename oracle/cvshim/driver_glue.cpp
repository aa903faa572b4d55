// driver_glue.cpp -- TEST INFRASTRUCTURE ONLY (GPU box).
//
// Runs the reference's own ORBmatcher drivers and their liborbx-backed replacements (cpp/orbmatcher_drivers_b200.hpp:
// class ORBmatcherB200, the code a maintainer drops into the reference tree) side by side INSIDE the reference's own data
// model: reference OrbFrames built by the reference's stereo constructor (frame_glue.cpp), reference OrbMapPoints, the
// reference's orbmatcher.cpp compiled unmodified.  Both classes get identical copies of the searched frame; what they leave
// in m_mapPoints and what they return is compared element by element.
//   part 1  SearchByProjection(frame, map points, th)             src/orbmatcher.cpp:42-124
//   part 2  SearchByProjection(CurrentFrame, LastFrame, th, mono)  src/orbmatcher.cpp:1337-1483, in its three level modes
//           (forward / backward / neither, :1357-1358) and with the orientation histogram
//   part 3  SearchByBoW(keyFrame, frame, matches)                  src/orbmatcher.cpp:164-292
//   part 4  SearchForInitialization(F1, F2, prevMatched, matches)  src/orbmatcher.cpp:411-528
//   part 5  SearchByProjection(CurrentFrame, keyFrame, found, th, d) src/orbmatcher.cpp:1485-1616
//   part 6  SearchByBoW(keyFrame1, keyFrame2, matches12)           src/orbmatcher.cpp:531-663
//   part 8  SearchByProjection(keyFrame, Scw, points, matched, th)   src/orbmatcher.cpp:294-409
//   part 9  Fuse(keyFrame, Scw, points, th, replace)                src/orbmatcher.cpp:984-1108
//   part 10 Fuse(keyFrame, points, th)                              src/orbmatcher.cpp:833-982
//   part 11 SearchBySim3(keyFrame1, keyFrame2, matches12, s12, R12, t12, th) src/orbmatcher.cpp:1110-1335
//   part 7  SearchForTriangulation(keyFrame1, keyFrame2, F12, ...)  src/orbmatcher.cpp:665-831 (real key frames: the reference's orbkeyframe.cpp)
// Built by `make -C oracle ref` into oracle/_ref/libdriverref.so (links liborbx.so); used by tests/test_gpu_drivers.py.
#include <orbframe.hpp>
#include <orbmatcher.hpp>
#include <orbkeyframe.hpp>

#include <chrono>
#include <cstdlib>
#include <new>
#include <cstring>
#include <sstream>

#include "orbmatcher_drivers_b200.hpp"

extern "C" {
#include "orb_oracle.h"
struct frameref_cfg { int nfeatures; float scale; int nlevels, ini_th, min_th; };
std::shared_ptr<OrbFrame> frameref_make_frame(const frameref_cfg *c, const uint8_t *left, const uint8_t *right, int w, int h, float mbf, float mb);
void orbref_canonical(int on);
}
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe(int rows);
std::shared_ptr<OrbKeyFrame> mpref_standin_keyframe_with(const std::vector<cv::KeyPoint> &keysUn, const cv::Mat &descriptors,
                                                         const std::vector<std::shared_ptr<OrbMapPoint>> &mapPoints);
void mpref_standin_clear();
void mpref_set_template_frame(const std::shared_ptr<OrbFrame> &f);

static int count_mismatches(const std::shared_ptr<OrbFrame> &a, const std::shared_ptr<OrbFrame> &b, int *assigned)
{
    int bad = 0, set = 0;
    for (int k = 0; k < a->N; k++) {
        if (a->m_mapPoints[k] != b->m_mapPoints[k]) bad++;
        if (a->m_mapPoints[k]) set++;
    }
    if (assigned) *assigned = set;
    return bad;
}

extern "C" {

// out[0..3]   part 1: nmatches reference, nmatches ORBmatcherB200, differing m_mapPoints entries, entries set
// out[4+4m..] part 2, mode m = 0 forward, 1 backward, 2 neither: the same four numbers
// out[16..19] part 3: SearchByBoW(key frame, frame): the same four numbers
// out[28..31] wall microseconds of one warm call: part 1 reference / ORBmatcherB200, part 3 reference / ORBmatcherB200
// out[24..27] part 5: SearchByProjection(CurrentFrame, key frame, found): the four numbers of part 1
// out[20..23] part 4: SearchForInitialization(F1, F2): nmatches x 2, differing vnMatches12 / vbPrevMatched entries, matches
// out[36..43] part 7: SearchForTriangulation(key frame 1, key frame 2), all features / stereo only: nmatches x 2, differing pairs, pairs
// out[44..47] part 8: SearchByProjection(key frame, Scw, ...): nmatches x 2, differing vpMatched entries, entries set
// out[48..52] part 9: Fuse(key frame, Scw, ...): nFused x 2, differences (key-frame matches, corrupt flags, observation counts, vpReplacePoint), matches held, replacements
// out[53..57] part 10: Fuse(key frame, points): nFused x 2, differences, matches held, corrupt points afterwards
// out[58..61] part 11: SearchBySim3(key frame 1, key frame 2, ...): nFound x 2, differing vpMatches12 entries, entries set
// out[32..35] part 6: SearchByBoW(key frame 1, key frame 2): nmatches x 2, differing vpMatches12 entries, entries set
int driverref_check(const frameref_cfg *c, const uint8_t *leftA, const uint8_t *rightA, const uint8_t *leftB, const uint8_t *rightB,
                    int w, int h, float mbf, float mb, float th_points, float th_frames, float nnratio, float dx, float dy, int32_t *out)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    std::shared_ptr<OrbKeyFrame> kf;
    int rc = 0;
    try {
        std::shared_ptr<OrbFrame> A = frameref_make_frame(c, leftA, rightA, w, h, mbf, mb);
        std::shared_ptr<OrbFrame> B = frameref_make_frame(c, leftB, rightB, w, h, mbf, mb);
        cv::Mat I(4, 4, CV_32F);
        for (int i = 0; i < 16; i++) I.ptr<float>(i / 4)[i % 4] = (i % 5 == 0) ? 1.f : 0.f;
        A->SetPose(I);
        mpref_set_template_frame(A);                 // the key frames below are built by the reference's own constructor from copies of A
        kf = mpref_standin_keyframe(8);
        cv::Mat one(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) one.ptr<float>(k)[0] = 1.f;

        // ---------------- part 1
        {
            std::vector<std::shared_ptr<OrbMapPoint>> mps;
            // As in Tracking::SearchLocalPoints the map points carry observations (three out of four here), so the rule of
            // :87-89 is live INSIDE the call: a map point stored at :121 hides its key point from the ones after it.  Every
            // third key point of A yields a second, identical map point right behind the first: the two collide on one key
            // point of B, and the twin has to settle for another key point or for none.
            for (int i = 0; i < A->N; i++)
                for (int twin = 0; twin < (i % 3 == 0 ? 2 : 1); twin++) {
                    cv::Mat pos(3, 1, CV_32F);
                    pos.ptr<float>(0)[0] = 0.f; pos.ptr<float>(1)[0] = 0.f; pos.ptr<float>(2)[0] = 5.f;
                    auto mp = std::make_shared<OrbMapPoint>(pos, A, std::shared_ptr<OrbMap>(), i);
                    mp->SetTrackInView(i % 13 != 0);                              // some points are not in view (:51)
                    mp->SetTrackProjX(A->m_undistortedKeys[i].pt.x + dx);
                    mp->SetTrackProjY(A->m_undistortedKeys[i].pt.y + dy);
                    mp->SetnTrackScaleLevel(A->m_undistortedKeys[i].octave);
                    mp->SetTrackViewCos((i & 1) ? 0.9f : 0.9995f);
                    if (mps.size() % 4 != 1) mp->AddObservingKeyframe(kf, 1);
                    mps.push_back(mp);
                }
            for (int idx = 0; idx < B->N; idx++)
                if (idx % 7 == 3 || idx % 11 == 5) {
                    auto held = std::make_shared<OrbMapPoint>(one, kf, std::shared_ptr<OrbMap>());
                    if (idx % 7 == 3) held->AddObservingKeyframe(kf, 1);
                    B->m_mapPoints[idx] = held;
                }
            std::shared_ptr<OrbFrame> B1 = std::make_shared<OrbFrame>(B), B2 = std::make_shared<OrbFrame>(B);
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            out[0] = ref.SearchByProjection(B1, mps, th_points);
            out[1] = gpu.SearchByProjection(B2, mps, th_points);
            out[2] = count_mismatches(B1, B2, &out[3]);
            {   // wall time of one more call each on fresh copies (the first GPU call above allocated the workspace)
                std::shared_ptr<OrbFrame> T1 = std::make_shared<OrbFrame>(B), T2 = std::make_shared<OrbFrame>(B);
                auto t0 = std::chrono::steady_clock::now();
                ref.SearchByProjection(T1, mps, th_points);
                auto t1 = std::chrono::steady_clock::now();
                gpu.SearchByProjection(T2, mps, th_points);
                auto t2 = std::chrono::steady_clock::now();
                out[28] = (int32_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
                out[29] = (int32_t)std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
            }
            for (int idx = 0; idx < B->N; idx++) B->m_mapPoints[idx] = std::shared_ptr<OrbMapPoint>();
        }

        // ---------------- part 2: the last frame's key points become observed map points at their stereo depth (8 m where
        // the frame has none), every 17th is an outlier; the current frame moves along z (forward / backward) or barely
        for (int i = 0; i < A->N; i++) {
            const float z = A->m_depths[i] > 0 ? A->m_depths[i] : 8.0f;
            cv::Mat pos(3, 1, CV_32F);
            pos.ptr<float>(0)[0] = (A->m_undistortedKeys[i].pt.x - OrbFrame::cx) * z * OrbFrame::invfx;
            pos.ptr<float>(1)[0] = (A->m_undistortedKeys[i].pt.y - OrbFrame::cy) * z * OrbFrame::invfy;
            pos.ptr<float>(2)[0] = z;
            if (i % 5 == 4) continue;                                             // key points without a map point
            auto mp = std::make_shared<OrbMapPoint>(pos, A, std::shared_ptr<OrbMap>(), i);
            mp->AddObservingKeyframe(kf, 1);                                      // observed: the exclusion of :1412-1414 is live
            A->m_mapPoints[i] = mp;
            A->m_outliers[i] = (i % 17 == 0);
        }
        const float tz[3] = {-0.9f, 0.9f, 0.05f};
        for (int m = 0; m < 3; m++) {
            cv::Mat T = I.clone();
            T.ptr<float>(0)[3] = 0.02f; T.ptr<float>(1)[3] = -0.01f; T.ptr<float>(2)[3] = tz[m];
            std::shared_ptr<OrbFrame> C1 = std::make_shared<OrbFrame>(B), C2 = std::make_shared<OrbFrame>(B);
            C1->SetPose(T); C2->SetPose(T);
            C1->mb = mb; C2->mb = mb;
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            out[4 + 4 * m] = ref.SearchByProjection(C1, A, th_frames, false);
            out[5 + 4 * m] = gpu.SearchByProjection(C2, A, th_frames, false);
            out[6 + 4 * m] = count_mismatches(C1, C2, &out[7 + 4 * m]);
        }

        // ---------------- part 3: SearchByBoW(key frame, frame).  The key frame carries frame A's key points, descriptors and
        // (for four key points out of five) observed map points; "vocabulary nodes" are a function of the descriptor
        // (its first byte's low six bits), so equal descriptors of A and B share a node as under a real vocabulary.
        {
            std::vector<std::shared_ptr<OrbMapPoint>> kfPoints(A->m_mapPoints);
            std::shared_ptr<OrbKeyFrame> KF = mpref_standin_keyframe_with(A->m_undistortedKeys, A->m_descriptors, kfPoints);
            for (int i = 0; i < A->N; i++) KF->m_features.addFeature(A->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            std::shared_ptr<OrbFrame> D1 = std::make_shared<OrbFrame>(B), D2 = std::make_shared<OrbFrame>(B);
            for (int i = 0; i < B->N; i++) {
                D1->mFeatVec.addFeature(B->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
                D2->mFeatVec.addFeature(B->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            }
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            std::vector<std::shared_ptr<OrbMapPoint>> m1, m2;
            out[16] = ref.SearchByBoW(KF, D1, m1);
            out[17] = gpu.SearchByBoW(KF, D2, m2);
            {
                std::vector<std::shared_ptr<OrbMapPoint>> t1v, t2v;
                auto t0 = std::chrono::steady_clock::now();
                ref.SearchByBoW(KF, D1, t1v);
                auto t1 = std::chrono::steady_clock::now();
                gpu.SearchByBoW(KF, D2, t2v);
                auto t2 = std::chrono::steady_clock::now();
                out[30] = (int32_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
                out[31] = (int32_t)std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
            }
            int bad = (m1.size() != m2.size()), set = 0;
            for (size_t k = 0; k < m1.size() && k < m2.size(); k++) { if (m1[k] != m2[k]) bad++; if (m1[k]) set++; }
            out[18] = bad; out[19] = set;
            mpref_standin_clear();
        }

        // ---------------- part 5: SearchByProjection(CurrentFrame, key frame, already found, th, ORBdist): the stand-in key frame
        // carries frame A's map points (world positions from part 2); every ninth is "already found"
        {
            std::vector<std::shared_ptr<OrbMapPoint>> kfPoints(A->m_mapPoints);
            std::shared_ptr<OrbKeyFrame> KF = mpref_standin_keyframe_with(A->m_undistortedKeys, A->m_descriptors, kfPoints);
            std::set<std::shared_ptr<OrbMapPoint>> found;
            for (size_t i = 0; i < kfPoints.size(); i += 9) if (kfPoints[i]) found.insert(kfPoints[i]);
            cv::Mat T = I.clone();
            T.ptr<float>(0)[3] = 0.03f; T.ptr<float>(1)[3] = 0.01f; T.ptr<float>(2)[3] = -0.2f;
            std::shared_ptr<OrbFrame> E1 = std::make_shared<OrbFrame>(B), E2 = std::make_shared<OrbFrame>(B);
            E1->SetPose(T); E2->SetPose(T);
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            out[24] = ref.SearchByProjection(E1, KF, found, th_frames, 100);
            out[25] = gpu.SearchByProjection(E2, KF, found, th_frames, 100);
            out[26] = count_mismatches(E1, E2, &out[27]);
            mpref_standin_clear();
        }

        // ---------------- part 4: SearchForInitialization(F1, F2): windows around F1's own level-0 key points shifted by (dx, dy)
        {
            std::shared_ptr<OrbFrame> F1 = A, F2 = B;
            std::vector<cv::Point2f> prev1, prev2;
            for (size_t i = 0; i < A->m_undistortedKeys.size(); i++)
                prev1.push_back(cv::Point2f(A->m_undistortedKeys[i].pt.x + dx, A->m_undistortedKeys[i].pt.y + dy));
            prev2 = prev1;
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            std::vector<int> m1, m2;
            out[20] = ref.SearchForInitialization(F1, F2, prev1, m1, 20);
            out[21] = gpu.SearchForInitialization(F1, F2, prev2, m2, 20);
            int bad = (m1.size() != m2.size()), set = 0;
            for (size_t k = 0; k < m1.size() && k < m2.size(); k++) {
                if (m1[k] != m2[k] || prev1[k].x != prev2[k].x || prev1[k].y != prev2[k].y) bad++;
                if (m1[k] >= 0) set++;
            }
            out[22] = bad; out[23] = set;
        }
        // ---------------- part 6: SearchByBoW(key frame 1, key frame 2) (loop closing, :531-663).  Two stand-in key frames carry the
        // key points and descriptors of frames A and B; three of four features of A and every second of B own a map point, some
        // of them corrupt (:571, :591); nodes as in part 3
        {
            std::vector<std::shared_ptr<OrbMapPoint>> p1((size_t)A->N), p2((size_t)B->N);
            for (int i = 0; i < A->N; i++)
                if (i % 4 != 2) {
                    p1[i] = std::make_shared<OrbMapPoint>(one, kf, std::shared_ptr<OrbMap>());
                    if (i % 11 == 4) p1[i]->SetCorruptFlag();
                }
            for (int i = 0; i < B->N; i++)
                if (i % 2 == 0) {
                    p2[i] = std::make_shared<OrbMapPoint>(one, kf, std::shared_ptr<OrbMap>());
                    if (i % 14 == 6) p2[i]->SetCorruptFlag();
                }
            std::shared_ptr<OrbKeyFrame> K1 = mpref_standin_keyframe_with(A->m_undistortedKeys, A->m_descriptors, p1);
            std::shared_ptr<OrbKeyFrame> K2 = mpref_standin_keyframe_with(B->m_undistortedKeys, B->m_descriptors, p2);
            for (int i = 0; i < A->N; i++) K1->m_features.addFeature(A->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            for (int i = 0; i < B->N; i++) K2->m_features.addFeature(B->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            std::vector<std::shared_ptr<OrbMapPoint>> m1, m2;
            out[32] = ref.SearchByBoW(K1, K2, m1);
            out[33] = gpu.SearchByBoW(K1, K2, m2);
            int bad = (m1.size() != m2.size()), set = 0;
            for (size_t k = 0; k < m1.size() && k < m2.size(); k++) { if (m1[k] != m2[k]) bad++; if (m1[k]) set++; }
            out[34] = bad; out[35] = set;
            mpref_standin_clear();
        }
        // ---------------- part 7: SearchForTriangulation(key frame 1, key frame 2, F12, pairs, onlyStereo) (:665-831) on two REAL key
        // frames built by the reference's own constructor from frames A and B: every third feature of A and every fourth of B keeps
        // a map point (those are skipped, :703-707 / :722-724), key frame 2 sits 3 cm to the right and 20 cm back, F12 has
        // horizontal epipolar lines (l = x1' F12 = [0, 1, -y1]); nodes as in part 3; once with all features, once stereo only
        {
            std::shared_ptr<OrbFrame> FA = std::make_shared<OrbFrame>(A), FB = std::make_shared<OrbFrame>(B);
            for (int i = 0; i < FA->N; i++) FA->m_mapPoints[i] = (i % 3 == 1) ? std::make_shared<OrbMapPoint>(one, kf, std::shared_ptr<OrbMap>()) : std::shared_ptr<OrbMapPoint>();
            for (int i = 0; i < FB->N; i++) FB->m_mapPoints[i] = (i % 4 == 2) ? std::make_shared<OrbMapPoint>(one, kf, std::shared_ptr<OrbMap>()) : std::shared_ptr<OrbMapPoint>();
            FA->mFeatVec.clear(); FB->mFeatVec.clear();
            for (int i = 0; i < FA->N; i++) FA->mFeatVec.addFeature(FA->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            for (int i = 0; i < FB->N; i++) FB->mFeatVec.addFeature(FB->m_descriptors.ptr(i)[0] & 63u, (uint32_t)i);
            cv::Mat T = I.clone();
            T.ptr<float>(0)[3] = 0.03f; T.ptr<float>(1)[3] = 0.01f; T.ptr<float>(2)[3] = -0.2f;
            FA->SetPose(I); FB->SetPose(T);
            std::shared_ptr<OrbKeyFrame> K1 = std::make_shared<OrbKeyFrame>(FA, std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
            std::shared_ptr<OrbKeyFrame> K2 = std::make_shared<OrbKeyFrame>(FB, std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>());
            cv::Mat F12(3, 3, CV_32F);
            for (int i = 0; i < 9; i++) F12.ptr<float>(i / 3)[i % 3] = 0.f;
            F12.ptr<float>(2)[1] = 1.f; F12.ptr<float>(1)[2] = -1.f;
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            for (int only = 0; only < 2; only++) {
                std::vector<std::pair<size_t, size_t>> p1, p2;
                out[36 + 4 * only] = ref.SearchForTriangulation(K1, K2, F12, p1, only != 0);
                out[37 + 4 * only] = gpu.SearchForTriangulation(K1, K2, F12, p2, only != 0);
                int bad = (p1.size() != p2.size());
                for (size_t k = 0; k < p1.size() && k < p2.size(); k++) if (p1[k] != p2[k]) bad++;
                out[38 + 4 * only] = bad; out[39 + 4 * only] = (int)p1.size();
            }
        }
        // ---------------- parts 8-10: the drivers that project map points into a KEY FRAME and walk its grid.  The key frame is built by
        // the reference's constructor from a copy of frame B (pose: 3 cm right, 1 cm down, 20 cm back; every fifth feature owns a map
        // point that observes the key frame), the candidates are frame A's key points at their stereo depth (8 m without one), built
        // by the reference's OrbMapPoint(position, frame, map, index) constructor; every 13th is corrupt.  Each run gets a model of
        // its own (the Fuse drivers store observations and replace map points).  The observer `k0` and the key frame live side by
        // side in one buffer, k0 first: Fuse(keyFrame, points) compares std::maps keyed by key-frame ADDRESS (:962), and with k0 the
        // lowest key of every map the comparison is decided by k0's feature index in both models alike.
        {
            struct Model {
                void *raw = nullptr;
                std::shared_ptr<OrbKeyFrame> k0, K;
                std::vector<std::shared_ptr<OrbMapPoint>> inKF, cand;
                ~Model() { inKF.clear(); cand.clear(); K.reset(); k0.reset(); free(raw); }
            };
            cv::Mat T = I.clone();
            T.ptr<float>(0)[3] = 0.03f; T.ptr<float>(1)[3] = 0.01f; T.ptr<float>(2)[3] = -0.2f;
            auto build = [&](Model &M) {
                M.raw = malloc(sizeof(OrbKeyFrame) * 2 + alignof(OrbKeyFrame));
                char *base = (char *)(((uintptr_t)M.raw + alignof(OrbKeyFrame) - 1) & ~(uintptr_t)(alignof(OrbKeyFrame) - 1));
                std::shared_ptr<OrbFrame> F0 = std::make_shared<OrbFrame>(A), FB = std::make_shared<OrbFrame>(B);
                for (int i = 0; i < F0->N; i++) F0->m_mapPoints[i] = std::shared_ptr<OrbMapPoint>();
                for (int i = 0; i < FB->N; i++) FB->m_mapPoints[i] = std::shared_ptr<OrbMapPoint>();
                F0->SetPose(I); FB->SetPose(T);
                M.k0 = std::shared_ptr<OrbKeyFrame>(new (base) OrbKeyFrame(F0, std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>()),
                                                    [](OrbKeyFrame *k) { k->~OrbKeyFrame(); });
                M.K = std::shared_ptr<OrbKeyFrame>(new (base + sizeof(OrbKeyFrame)) OrbKeyFrame(FB, std::shared_ptr<OrbMap>(), std::shared_ptr<OrbKeyFrameDatabase>()),
                                                   [](OrbKeyFrame *k) { k->~OrbKeyFrame(); });
                M.inKF.assign((size_t)B->N, std::shared_ptr<OrbMapPoint>());
                for (int i = 0; i < B->N; i += 5) {
                    const float z = B->m_depths[i] > 0 ? B->m_depths[i] : 8.0f;
                    cv::Mat pos(3, 1, CV_32F);
                    pos.ptr<float>(0)[0] = (B->m_undistortedKeys[i].pt.x - OrbFrame::cx) * z * OrbFrame::invfx;
                    pos.ptr<float>(1)[0] = (B->m_undistortedKeys[i].pt.y - OrbFrame::cy) * z * OrbFrame::invfy;
                    pos.ptr<float>(2)[0] = z;
                    auto mp = std::make_shared<OrbMapPoint>(pos, FB, std::shared_ptr<OrbMap>(), i);
                    mp->AddObservingKeyframe(M.k0, (size_t)(1 + i % 2));
                    mp->AddObservingKeyframe(M.K, (size_t)i);
                    M.K->AddMapPoint(mp, (size_t)i);
                    M.inKF[i] = mp;
                }
                for (int i = 0; i < A->N; i++) {
                    const float z = A->m_depths[i] > 0 ? A->m_depths[i] : 8.0f;
                    cv::Mat pos(3, 1, CV_32F);
                    pos.ptr<float>(0)[0] = (A->m_undistortedKeys[i].pt.x - OrbFrame::cx) * z * OrbFrame::invfx;
                    pos.ptr<float>(1)[0] = (A->m_undistortedKeys[i].pt.y - OrbFrame::cy) * z * OrbFrame::invfy;
                    pos.ptr<float>(2)[0] = z;
                    auto mp = std::make_shared<OrbMapPoint>(pos, F0, std::shared_ptr<OrbMap>(), i);
                    mp->AddObservingKeyframe(M.k0, (size_t)(1 + i % 3));
                    if (i % 13 == 7) mp->SetCorruptFlag();
                    M.cand.push_back(mp);
                    if (i % 29 == 3) M.cand.push_back(mp);                       // a point listed twice
                }
            };
            // what the two models hold afterwards, position by position
            auto label = [](const Model &M, const std::shared_ptr<OrbMapPoint> &p) -> int {
                if (!p) return -1;
                for (size_t k = 0; k < M.cand.size(); k++) if (M.cand[k] == p) return (int)k;
                for (size_t k = 0; k < M.inKF.size(); k++) if (M.inKF[k] == p) return 1000000 + (int)k;
                return -2;
            };
            auto differences = [&](Model &X, Model &Y, int *held) {
                int bad = 0, set = 0;
                for (int i = 0; i < X.K->N; i++) {
                    const int a = label(X, X.K->GetMapPoint((size_t)i)), b = label(Y, Y.K->GetMapPoint((size_t)i));
                    if (a != b) bad++;
                    if (a != -1) set++;
                }
                for (size_t k = 0; k < X.cand.size(); k++)
                    if (X.cand[k]->IsCorrupt() != Y.cand[k]->IsCorrupt() ||
                        X.cand[k]->GetObservingKeyFrameCount() != Y.cand[k]->GetObservingKeyFrameCount()) bad++;
                for (size_t k = 0; k < X.inKF.size(); k++)
                    if (X.inKF[k] && (X.inKF[k]->IsCorrupt() != Y.inKF[k]->IsCorrupt() ||
                                      X.inKF[k]->GetObservingKeyFrameCount() != Y.inKF[k]->GetObservingKeyFrameCount())) bad++;
                *held = set;
                return bad;
            };
            cv::Mat Scw = T.clone();                                              // a similarity with scale 1.25: [s R | s t]
            for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) Scw.ptr<float>(r)[c] *= 1.25f;
            ORBmatcher ref(nnratio, true);
            ORBmatcherB200 gpu(nnratio, true);
            {   // part 8: SearchByProjection(key frame, Scw, points, matched, th) (:294-409): vpMatched starts as the key frame's own matches
                Model X, Y; build(X); build(Y);
                std::vector<std::shared_ptr<OrbMapPoint>> m1 = X.K->GetMapPointMatches(), m2 = Y.K->GetMapPointMatches();
                out[44] = ref.SearchByProjection(X.K, Scw, X.cand, m1, 10);
                out[45] = gpu.SearchByProjection(Y.K, Scw, Y.cand, m2, 10);
                int bad = (m1.size() != m2.size()), set = 0;
                for (size_t k = 0; k < m1.size() && k < m2.size(); k++) { if (label(X, m1[k]) != label(Y, m2[k])) bad++; if (m1[k]) set++; }
                out[46] = bad; out[47] = set;
            }
            {   // part 9: Fuse(key frame, Scw, points, th, replace) (:984-1108)
                Model X, Y; build(X); build(Y);
                std::vector<std::shared_ptr<OrbMapPoint>> r1(X.cand.size()), r2(Y.cand.size());
                out[48] = ref.Fuse(X.K, Scw, X.cand, 4.0f, r1);
                out[49] = gpu.Fuse(Y.K, Scw, Y.cand, 4.0f, r2);
                int held = 0, bad = differences(X, Y, &held), repl = 0;
                for (size_t k = 0; k < r1.size(); k++) { if (label(X, r1[k]) != label(Y, r2[k])) bad++; if (r1[k]) repl++; }
                out[50] = bad; out[51] = held; out[52] = repl;
            }
            {   // part 11: SearchBySim3(key frame 1, key frame 2, matches12, s12, R12, t12, th) (:1110-1335).  Key frame 1 = k0 (frame A at
                // the origin) with a map point on three of four features, key frame 2 = K with its own points; the similarity is the
                // true relative pose (scale 1); every 19th entry of vpMatches12 is set beforehand (:1139-1149)
                Model X; build(X);
                for (int i = 0; i < A->N; i++)
                    if (i % 4 != 1) {
                        const float z = A->m_depths[i] > 0 ? A->m_depths[i] : 8.0f;
                        cv::Mat pos(3, 1, CV_32F);
                        pos.ptr<float>(0)[0] = (A->m_undistortedKeys[i].pt.x - OrbFrame::cx) * z * OrbFrame::invfx;
                        pos.ptr<float>(1)[0] = (A->m_undistortedKeys[i].pt.y - OrbFrame::cy) * z * OrbFrame::invfy;
                        pos.ptr<float>(2)[0] = z;
                        std::shared_ptr<OrbFrame> F0 = std::make_shared<OrbFrame>(A);
                        F0->SetPose(I);
                        auto mp = std::make_shared<OrbMapPoint>(pos, F0, std::shared_ptr<OrbMap>(), i);
                        mp->AddObservingKeyframe(X.k0, (size_t)i);
                        X.k0->AddMapPoint(mp, (size_t)i);
                    }
                // T12 = T1w * Tw2 with T1w = I: the inverse of K's pose
                cv::Mat R12 = T.rowRange(0, 3).colRange(0, 3).t();
                cv::Mat t12 = -R12 * T.rowRange(0, 3).col(3);
                const float s12 = 1.0f;
                std::vector<std::shared_ptr<OrbMapPoint>> m1((size_t)X.k0->N), m2;
                for (int i = 0; i < X.k0->N; i += 19) m1[i] = X.inKF[(size_t)(5 * (i % 40))];
                m2 = m1;
                out[58] = ref.SearchBySim3(X.k0, X.K, m1, s12, R12, t12, 7.5f);
                out[59] = gpu.SearchBySim3(X.k0, X.K, m2, s12, R12, t12, 7.5f);
                int bad = (m1.size() != m2.size()), set = 0;
                for (size_t k = 0; k < m1.size() && k < m2.size(); k++) { if (m1[k] != m2[k]) bad++; if (m1[k]) set++; }
                out[60] = bad; out[61] = set;
            }
            {   // part 10: Fuse(key frame, points, th) (:833-982), with its Replace calls
                Model X, Y; build(X); build(Y);
                out[53] = ref.Fuse(X.K, X.cand, 3.0f);
                out[54] = gpu.Fuse(Y.K, Y.cand, 3.0f);
                int held = 0;
                out[55] = differences(X, Y, &held);
                out[56] = held;
                int corrupt = 0;
                for (size_t k = 0; k < X.cand.size(); k++) corrupt += X.cand[k]->IsCorrupt();
                for (size_t k = 0; k < X.inKF.size(); k++) if (X.inKF[k]) corrupt += X.inKF[k]->IsCorrupt();
                out[57] = corrupt;
            }
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "driverref_check: %s\n", e.what());
        rc = -1;
    }
    std::cout.rdbuf(old);
    return rc;
}

} // extern "C"
