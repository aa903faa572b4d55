/*
 * orb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the ORB front-end hot path of
 * chalmers-revere/opendlv-perception-vision-orbslam2 (OrbExtractor +
 * ORBmatcher::DescriptorDistance and the best/second-best loop).  It is the
 * parity checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (liborbx.so) never links or calls anything in this directory.
 *
 * Pinning: the reference ships no golden vectors for this path (its only built
 * test is test/tests-opendlv-perception-vision-orbslam2.cpp:29-38, a converter
 * smoke test).  The restatement is therefore pinned two ways:
 *   (1) the five OpenCV primitives below are checked bit-for-bit against
 *       cv2 4.13 (tests/test_oracle_primitives.py), and
 *   (2) the whole extractor is checked against oracle/_ref/liborbref.so, which
 *       is the reference's own src/orbextractor.cpp compiled UNMODIFIED against
 *       a header shim (oracle/cvshim) whose cv:: primitives call the functions
 *       below (tests/test_oracle_vs_ref.py + tests/golden/).
 * References in comments are file:line under /root/reference.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBO_MAX_LEVELS 16

/* 28-byte record, layout-compatible with cv::KeyPoint. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbo_keypoint;

typedef struct {
    int nfeatures;
    float scale_factor;
    int nlevels;
    int ini_th, min_th;
    int taps[7];                       /* Gaussian 7-tap integer kernel, sum ~256 */
    float sf[ORBO_MAX_LEVELS];         /* orbextractor.cpp:494-500 */
    float inv_sf[ORBO_MAX_LEVELS];     /* :502-508 */
    float sigma2[ORBO_MAX_LEVELS];
    float inv_sigma2[ORBO_MAX_LEVELS];
    int quota[ORBO_MAX_LEVELS];        /* :512-523 */
    int umax[16];                      /* :532-547 */
} orbo_params;

/* ---- constructor tables (orbextractor.cpp:476-548) ---- */
int orbo_params_init(orbo_params *p, int nfeatures, float scale_factor, int nlevels,
                     int ini_th, int min_th, const int *taps7 /* NULL -> cv2-4.13 taps */);
/* level size: orbextractor.cpp:659 */
void orbo_level_size(const orbo_params *p, int w0, int h0, int level, int *w, int *h);

/* ---- OpenCV primitives restated ---- */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstride,
                           uint8_t *dst, int dw, int dh, size_t dstride);
void orbo_border_reflect101_u8(const uint8_t *src, int w, int h, size_t sstride,
                               uint8_t *dst, size_t dstride, int top, int bottom, int left, int right);
/* FAST-9/16 with 3x3 NMS on an isolated sub-image; returns count, raster order.
 * xs/ys/score arrays must hold `cap` entries; returns -1 on overflow. */
int orbo_fast9_nms(const uint8_t *img, int w, int h, size_t stride, int threshold,
                   int *xs, int *ys, int *score, int cap);
/* score map of the interior (0 where not a corner at `threshold`) -- helper for tests */
void orbo_fast9_score_map(const uint8_t *img, int w, int h, size_t stride, int threshold, uint8_t *score /* w*h */);
void orbo_gaussian7_u8(const uint8_t *src, int w, int h, size_t sstride,
                       uint8_t *dst, size_t dstride, const int taps[7]);
float orbo_fast_atan2(float y, float x);

/* ---- reference-specific stages ---- */
/* gridded FAST with per-cell threshold fallback, orbextractor.cpp:906-970.
 * Output coords are relative to (16,16) like vToDistributeKeys. Returns count or -1. */
int orbo_grid_fast(const uint8_t *lvl, int w, int h, size_t stride, int ini_th, int min_th,
                   int *xs, int *ys, int *score, int cap);
/* DistributeOctTree, orbextractor.cpp:680-904 + DivideNode :72-128, canonical
 * tie rule (equal size -> most recently created node first).  in: n candidates in
 * reference order; out: indices of the selected candidates in list order. Returns count.
 * tie_rule: 0 = newest-first (canonical / bump allocator), 1 = oldest-first. */
int orbo_distribute(const int *xs, const int *ys, const int *score, int n,
                    int minX, int maxX, int minY, int maxY, int N, int tie_rule,
                    int *out_idx, int out_cap);
/* IC_Angle, orbextractor.cpp:136-163 */
float orbo_ic_angle(const uint8_t *lvl, size_t stride, int cx, int cy, const int umax[16]);
/* computeOrbDescriptor, orbextractor.cpp:166-203 */
void orbo_rbrief(const uint8_t *blurred, size_t stride, int cx, int cy, float angle_deg, uint8_t desc[32]);
const int *orbo_pattern(void); /* 1024 ints = 512 (x,y) points */

/* ---- whole extractor (ExtractFeatures, orbextractor.cpp:582-642) ---- */
typedef struct orbo_extractor orbo_extractor;
orbo_extractor *orbo_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                            const int *taps7);
void orbo_destroy(orbo_extractor *e);
const orbo_params *orbo_get_params(const orbo_extractor *e);
/* returns number of keypoints, or <0 on error. kps[cap], desc[cap*32]. */
int orbo_extract(orbo_extractor *e, const uint8_t *img, int w, int h, size_t stride,
                 orbo_keypoint *kps, uint8_t *desc, int cap);
/* frame-partitioned multi-threaded driver (CPU baseline): kps[nimg*cap], desc[nimg*cap*32], counts[nimg] */
int orbo_extract_batch_mt(const orbo_extractor *cfg, const uint8_t *const *imgs, int nimg, int w, int h, size_t stride,
                          orbo_keypoint *kps, uint8_t *desc, int cap, int *counts, int nthreads);
/* pyramid level of the last orbo_extract call (ROI pointer inside the bordered buffer) */
const uint8_t *orbo_level(const orbo_extractor *e, int level, int *w, int *h, size_t *stride);
/* blurred level of the last call (NULL if the level had no keypoints) */
const uint8_t *orbo_blurred(const orbo_extractor *e, int level, int *w, int *h, size_t *stride);
/* per-level candidate list of the last call (coords relative to (16,16)) */
int orbo_candidates(const orbo_extractor *e, int level, const int **xs, const int **ys, const int **score);
void orbo_set_tie_rule(orbo_extractor *e, int tie_rule);

/* ---- OrbFrame::ComputeStereoMatches, orbframe.cpp:511-705 (pinned against the reference's own orbframe.cpp: oracle/_ref/libframeref.so, tests/golden/ref_stereo.npz) ----
 * pyrL/pyrR[l] = level ROI pointers of the two extractors, lw/lh/lstride per level; uRight/depth get nl floats
 * (-1 where there is no match).  Returns the number of matches before the median filter. */
int orbo_stereo_matches(const orbo_keypoint *kl, const uint8_t *dl, int nl,
                        const orbo_keypoint *kr, const uint8_t *dr, int nr,
                        const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *lw, const int *lh,
                        const size_t *lstride, const float *sf, const float *inv_sf,
                        float mbf, float mb, float *uRight, float *depth);

/* ---- matcher (orbmatcher.cpp:1662-1677, loop :208-232) ---- */
int orbo_descriptor_distance(const uint8_t a[32], const uint8_t b[32]);
void orbo_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt,
               int32_t *idx, int32_t *d1, int32_t *d2);
/* per-query candidate lists (CSR): inner loop of SearchByProjection, orbmatcher.cpp:76-114 */
void orbo_knn2_csr(const uint8_t *q, int nq, const uint8_t *t, const int32_t *offsets, const int32_t *indices,
                   int32_t *idx1, int32_t *d1, int32_t *idx2, int32_t *d2);
/* ORBmatcher::SearchByProjection(frame, map points, th) (orbmatcher.cpp:42-124) including the frame's grid
 * (orbframe.cpp:192-211, :308-393); mp_radius[i] = r * scaleFactor[level] as the caller forms it.  mp_observed[i] != 0: map
 * point i has GetObservingKeyFrameCount() > 0, so once accepted it hides its key point from the later map points of the same
 * call (the loop is sequential: :121 stores it, :87-89 tests it); NULL = none has.  Returns nmatches. */
int orbo_search_by_projection(const orbo_keypoint *keys, const float *uright, const uint8_t *occupied, const uint8_t *desc, int n,
                              float min_x, float min_y, float max_x, float max_y,
                              const uint8_t *mp_desc, const float *mp_x, const float *mp_y, const int32_t *mp_level,
                              const float *mp_radius, const uint8_t *mp_observed, int n_mp, float nnratio, int th_high,
                              int32_t *mp_match, int32_t *assigned);
/* OrbFrame::FilterKeyPoints (orbframe.cpp:403-445), in place; returns the new count */
int orbo_filter_keypoints(orbo_keypoint *keys, uint8_t *desc, int n, const float box[4]);
/* OrbFrame::AssignFeaturesToGrid (orbframe.cpp:192-211) as CSR over 64 x 48 cells (cell = ix * 48 + iy) */
void orbo_assign_grid(const orbo_keypoint *keys, int n, float min_x, float min_y, float max_x, float max_y,
                      int32_t *cell_start, int32_t *cell_items);
/* OrbFrame::GetFeaturesInArea (orbframe.cpp:308-380) for nq windows + DescriptorDistance of every feature found */
int orbo_area_distances(const orbo_keypoint *keys, const uint8_t *desc, int n, float min_x, float min_y, float max_x, float max_y,
                        const uint8_t *q_desc, const float *q_x, const float *q_y, const float *q_r, const int32_t *q_min_level,
                        const int32_t *q_max_level, int nq, int32_t *offsets, int32_t *indices, int32_t *dist, int cap);
/* OrbMapPoint::ComputeDistinctiveDescriptors (orbmappoint.cpp:314-383), batched over map points (CSR lists of rows of desc) */
void orbo_distinctive(const uint8_t *desc, const int32_t *offsets, const int32_t *indices, int n_points,
                      int32_t *best, int32_t *median);
/* OrbVocabulary::transform5 (orbvocabulary.cpp:203-242) for n features over a tree given as arrays */
void orbo_voc_transform(const int32_t *child_off, const int32_t *child_ids, const uint8_t *node_desc, const int32_t *word_id,
                        int L, int levels_up, const uint8_t *feat, int n, int32_t *word, int32_t *node);
/* multi-threaded variants for the CPU baseline (query-/frame-partitioned) */
void orbo_knn2_mt(const uint8_t *q, int nq, const uint8_t *t, int nt,
                  int32_t *idx, int32_t *d1, int32_t *d2, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
